// Shared-memory load throughput by lane pattern (cycles per warp-level LDS.128 / LDS.64, eight warps issuing
// back-to-back independent loads): what a partially active warp costs the LSU data pipe.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ld2(uint32_t a) {
  unsigned long long x, y;
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "r"(a) : "memory");
  return x ^ y;
}
__device__ __forceinline__ unsigned long long ld1(uint32_t a) {
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
  return v;
}

// pattern: 0 all lanes distinct; 1 broadcast one address; 2 one address per quarter-warp (teams of 8);
// 3 lanes 0..15 active, distinct; 4 four active lanes per quarter (16 active), distinct; 5 lanes 0..23 active;
// 6 six active lanes per quarter (24 active); 7 lanes 0..15 active, one address per 4 lanes
template <int kWide>
__global__ void __launch_bounds__(256) k(long long* out, int pattern) {
  __shared__ __align__(16) double buf[4096];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 4096; i += 256) buf[i] = i;
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(buf) + 2048 * (warp & 1);
  const int w = kWide ? 16 : 8;
  bool act = true;
  uint32_t a = base + w * lane;
  if (pattern == 1) a = base;
  if (pattern == 2) a = base + 112 * (lane >> 3);
  if (pattern == 3) act = lane < 16;
  if (pattern == 4) act = (lane & 7) < 4;
  if (pattern == 5) act = lane < 24;
  if (pattern == 6) act = (lane & 7) < 6;
  if (pattern == 7) { act = lane < 16; a = base + 112 * (lane >> 2); }
  unsigned long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  __syncthreads();
  const long long t0 = clock64();
  if (act) {
#pragma unroll 1
    for (int r = 0; r < 256; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i & 7] ^= kWide ? ld2(a + 512 * (i & 3)) : ld1(a + 512 * (i & 3));
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (lane == 0) out[warp] = t1 - t0;
  if ((acc[0] ^ acc[1] ^ acc[2] ^ acc[3] ^ acc[4] ^ acc[5] ^ acc[6] ^ acc[7]) == 12345ull) out[9] = 1;
}

int main() {
  long long* d;
  cudaMalloc(&d, 16 * sizeof(long long));
  const char* names[8] = {"32 lanes distinct", "broadcast 1 address", "1 address per quarter-warp", "lanes 0-15 distinct",
                          "4 lanes per quarter distinct", "lanes 0-23 distinct", "6 lanes per quarter distinct",
                          "lanes 0-15, 1 address per 4 lanes"};
  for (int wide = 1; wide >= 0; --wide)
    for (int p = 0; p < 8; ++p) {
      if (wide) k<1><<<1, 256>>>(d, p); else k<0><<<1, 256>>>(d, p);
      cudaDeviceSynchronize();
      long long h[8];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double mx = 0;
      for (int i = 0; i < 8; ++i) mx = h[i] > mx ? h[i] : mx;
      // eight warps x 4096 loads each through one SM's pipe
      printf("%s  %-34s %.2f cycles per warp-load (pipe), 8 warps\n", wide ? "LDS.128" : "LDS.64 ", names[p], mx / (8.0 * 4096.0));
    }
  return 0;
}
