// Latency micro-benchmarks behind the design of wrench_riccati_kernel.cuh (cycles, one warp / one CTA):
// dependent DFMA, dependent DADD, LDS, STS -> __syncwarp -> LDS round trip, __syncthreads of 8 warps,
// a divergent-team hand-over (if (team == s) {...} __syncwarp()), shuffle.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k(long long* out, double seed) {
  __shared__ double sm[4096];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < 4096; i += blockDim.x) sm[i] = seed + i;
  __syncthreads();
  long long t0, t1;
  double a = seed, b = 1.0000001, c = 1e-9;
  // 1. dependent DFMA
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) a = fma(a, b, c);
  t1 = clock64();
  if (tid == 0) out[0] = (t1 - t0);
  // 2. dependent DADD
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) a = a + c;
  t1 = clock64();
  if (tid == 0) out[1] = (t1 - t0);
  // 3. dependent LDS (pointer chase through indices)
  __shared__ int idx[1024];
  for (int i = tid; i < 1024; i += blockDim.x) idx[i] = (i * 7 + 3) & 1023;
  __syncthreads();
  int p = lane;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) p = idx[p];
  t1 = clock64();
  if (tid == 0) out[2] = (t1 - t0);
  a += p;
  // 4. STS -> __syncwarp -> LDS (other lane's value) round trip, dependent
  double v = a;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 128; ++i) {
    sm[tid] = v;
    __syncwarp();
    v = sm[tid ^ 1] + c;
    __syncwarp();
  }
  t1 = clock64();
  if (tid == 0) out[3] = (t1 - t0);
  a += v;
  // 5. __syncthreads back to back (all warps)
  __syncthreads();
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 128; ++i) __syncthreads();
  t1 = clock64();
  if (tid == 0) out[4] = (t1 - t0);
  // 6. divergent team hand-over: 4 teams of 8 lanes take turns, value passes through shared memory
  const int team = lane >> 3;
  v = a;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 32; ++i) {
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
      if (team == s) {
        const double x = sm[(tid + 24) & 255];
        sm[tid] = fma(x, b, c);
      }
      __syncwarp();
    }
  }
  t1 = clock64();
  if (tid == 0) out[5] = (t1 - t0);
  // 7. shuffle dependent
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 128; ++i) v = __shfl_xor_sync(0xffffffffu, v, 1) + c;
  t1 = clock64();
  if (tid == 0) out[6] = (t1 - t0);
  // 8. STS -> __syncthreads -> LDS round trip
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    sm[tid] = v;
    __syncthreads();
    v = sm[(tid + 32) & 255] + c;
    __syncthreads();
  }
  t1 = clock64();
  if (tid == 0) out[7] = (t1 - t0);
  if (a + v == 12345.678) out[8] = 1;
}

int main() {
  long long* d;
  cudaMalloc(&d, 16 * sizeof(long long));
  for (int threads : {32, 256}) {
    cudaMemset(d, 0, 16 * sizeof(long long));
    k<<<1, threads>>>(d, 1.5);
    k<<<1, threads>>>(d, 1.5);
    cudaDeviceSynchronize();
    long long h[16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("threads %d: DFMA dep %.1f  DADD dep %.1f  LDS chase %.1f  STS-syncwarp-LDS(+DADD, 2 syncwarps) %.1f  "
           "__syncthreads %.1f  team hand-over (LDS+DFMA+STS+syncwarp) %.1f  SHFL+DADD %.1f  STS-syncthreads-LDS(+DADD) x2bar %.1f\n",
           threads, h[0] / 256.0, h[1] / 256.0, h[2] / 256.0, h[3] / 128.0, h[4] / 128.0, h[5] / 128.0, h[6] / 128.0, h[7] / 64.0);
  }
  return 0;
}
