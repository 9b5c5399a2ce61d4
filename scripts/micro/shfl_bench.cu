// Throughput of SHFL against shared-memory loads on one SM (eight warps, independent instructions back to back):
// cycles of the SM per warp-level instruction for SHFL.BFLY / SHFL.IDX alone, LDS.128 alone, and the two interleaved
// 1 : 1 -- do shuffles and shared loads share a pipe?  (The load address moves with the round: NVVM hoists an
// asm volatile load with loop-invariant operands out of the loop, memory clobber or not.)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ld2(uint32_t a) {
  unsigned long long x, y;
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "r"(a) : "memory");
  return x ^ y;
}

// mode 0: SHFL.BFLY x 16; 1: SHFL.IDX x 16; 2: LDS.128 x 16; 3: 8 LDS.128 + 8 SHFL.BFLY; 4: 16 LDS.128 + 16 SHFL (sum of both)
template <int mode>
__global__ void __launch_bounds__(256) k(long long* out) {
  __shared__ __align__(16) double buf[4096];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 4096; i += 256) buf[i] = i;
  __syncthreads();
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(buf) + 2048 * (warp & 1) + 16 * lane;
  unsigned v[8];
  unsigned long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 8; ++i) v[i] = tid * 17 + i;
  const int src = (lane * 7 + 3) & 31;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < 256; ++r) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (mode == 0) v[i & 7] += __shfl_xor_sync(0xffffffffu, v[(i + 1) & 7], 1 + (i & 3));
      if (mode == 1) v[i & 7] += __shfl_sync(0xffffffffu, v[(i + 1) & 7], src);
      if (mode == 2) acc[i & 7] ^= ld2(a + 512 * (i & 3) + ((r & 7) << 4));
      if (mode == 3) {
        if (i & 1) v[i & 7] += __shfl_xor_sync(0xffffffffu, v[(i + 1) & 7], 1 + (i & 3));
        else acc[i & 7] ^= ld2(a + 512 * (i & 3) + ((r & 7) << 4));
      }
      if (mode == 4) {
        v[i & 7] += __shfl_xor_sync(0xffffffffu, v[(i + 1) & 7], 1 + (i & 3));
        acc[i & 7] ^= ld2(a + 512 * (i & 3) + ((r & 7) << 4));
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (lane == 0) out[warp] = t1 - t0;
  unsigned s = 0;
  for (int i = 0; i < 8; ++i) s += v[i] + (unsigned)acc[i];
  if (s == 12345u) out[9] = 1;
}

int main() {
  long long* d;
  cudaMalloc(&d, 16 * sizeof(long long));
  // (the 16 / 8 loads of a round have 4 / 2 distinct addresses and the compiler keeps one load per address)
  const char* names[5] = {"16 SHFL.BFLY", "16 SHFL.IDX", "4 LDS.128", "2 LDS.128 + 8 SHFL.BFLY", "4 LDS.128 + 16 SHFL.BFLY"};
  for (int m = 0; m < 5; ++m) {
    if (m == 0) k<0><<<1, 256>>>(d);
    if (m == 1) k<1><<<1, 256>>>(d);
    if (m == 2) k<2><<<1, 256>>>(d);
    if (m == 3) k<3><<<1, 256>>>(d);
    if (m == 4) k<4><<<1, 256>>>(d);
    cudaDeviceSynchronize();
    long long h[8];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int i = 0; i < 8; ++i) mx = h[i] > mx ? h[i] : mx;
    // eight warps x 256 rounds through one SM
    printf("%-28s %.2f cycles of the SM per round of one warp (8 warps)\n", names[m], mx / (8.0 * 256.0));
  }
  return 0;
}
