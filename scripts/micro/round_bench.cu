// Micro-benchmark of one recursion ROUND of wrench_riccati_kernel.cuh in the spread layout: eight warps, position
// r = warp / 2 works in round r (all four teams of the warp at once: 12-vector in, 12 FMAs on four accumulators,
// two stores), then a block barrier hands over to the next position.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ double2 ld2(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ double ld1(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void st1(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

template <int kSync>
__global__ void __launch_bounds__(256, 2) k(long long* out, double seed, int reps) {
  __shared__ __align__(16) double vec[40][12];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, team = lane >> 3, t8 = lane & 7;
  const int c = t8 < 6 ? t8 : t8 - 6;
  const int pos = warp >> 1, grp = 4 * (warp & 1) + team;
  double fc[12];
  for (int i = 0; i < 12; ++i) fc[i] = seed * 1e-3 * (i + 1 + c);
  for (int i = tid; i < 480; i += 256) (&vec[0][0])[i] = seed + i;
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(&vec[0][0]);
  const uint32_t src = base + 96 * (8 * pos + grp), dst = base + 96 * (8 * (pos + 1) + grp);
  const bool wr = t8 < 6;
  double tpos = seed, tvel = -seed;
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
      if (pos == s) {
        const double2 v01 = ld2(src + 48), v23 = ld2(src + 64), v45 = ld2(src + 80), p01 = ld2(src);
        const double op = ld1(src + 8 * c), ov = ld1(src + 48 + 8 * c);
        double s0 = fma(fc[0], v01.x, tpos), s1 = fc[1] * v01.y, s2 = fma(fc[6], v01.x, tvel), s3 = fc[7] * v01.y;
        s0 = fma(fc[2], v23.x, s0); s1 = fma(fc[3], v23.y, s1); s2 = fma(fc[8], v23.x, s2); s3 = fma(fc[9], v23.y, s3);
        s0 = fma(fc[4], v45.x, s0); s1 = fma(fc[5], v45.y, s1); s2 = fma(fc[10], v45.x, s2); s3 = fma(fc[11], v45.y, s3);
        const double rot = fma(0.3, p01.x, fma(0.7, p01.y, op));
        if (wr) { st1(dst + 8 * c, op + (s0 + s1)); st1(dst + 48 + 8 * c, fma(0.01, rot, ov) + (s2 + s3)); }
      }
      if (kSync == 0) __syncthreads();
      else asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  }
  long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 8 + warp] = t1 - t0;
  if (tpos == 1234.5) out[0] = (long long)vec[0][0];
}

int main() {
  long long* d;
  cudaMalloc(&d, 8 * 400 * sizeof(long long));
  const int reps = 2000;
  for (int blocks : {1, 148, 296}) {
    k<0><<<blocks, 256>>>(d, 1.0, reps);
    k<0><<<blocks, 256>>>(d, 1.0, reps);
    cudaDeviceSynchronize();
    long long h[8];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("rounds with __syncthreads, blocks %d: %.1f cycles per round\n", blocks, h[0] / (4.0 * reps));
  }
  return 0;
}
