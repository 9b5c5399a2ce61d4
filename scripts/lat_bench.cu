// Dependent-chain latencies and small-CTA throughput on this GPU (cycles, clock64).
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__global__ void lat_kernel(double* out, long long* cyc, double a, double b) {
  __shared__ double sm[512];
  __shared__ int idx[512];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) { sm[i] = 1.0 + i * 1e-9; idx[i] = (i * 2 + 2) & 511; }
  __syncthreads();
  double x = threadIdx.x * 1e-3 + 1.0;
  long long t0, t1;
  // DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, a, b);
  t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // DADD chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x + b;
  t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
  // DMUL chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x * a;
  t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // fmax/fmin clamp chain (2 ops per iteration)
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fmin(fmax(x + b, 0.5), 1e30);
  t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
  // 64-bit shuffle + add chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x + __shfl_xor_sync(0xffffffffu, x, 1);
  t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
  x = x * 1e-300 + 1.0;
  // LDS.64 pointer chase
  int j = threadIdx.x & 511;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) j = idx[j];
  t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
  x += j;
  // __syncthreads
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) __syncthreads();
  t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
  // FSEL-style select chain on doubles
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = (x > 2.0) ? x * 0.5 : x + b;
  t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
  // drcp chain
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N / 8; ++i) x = __drcp_rn(x) + 1.0;
  t1 = clock64(); if (threadIdx.x == 0) cyc[8] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
// throughput of the tile update shape: 48 independent DFMAs per thread per step, operands from smem
__global__ void tile_kernel(double* out, long long* cyc, int steps, int with_bar) {
  __shared__ __align__(16) double v[3][128];
  __shared__ __align__(16) double w[20][24];
  for (int i = threadIdx.x; i < 384; i += blockDim.x) v[i / 128][i % 128] = 1e-3 * i;
  for (int i = threadIdx.x; i < 480; i += blockDim.x) w[i / 24][i % 24] = 1e-4 * i;
  __syncthreads();
  const int rg = threadIdx.x >> 4, cg = threadIdx.x & 15;
  double a[6][8];
  for (int r = 0; r < 6; ++r) for (int c = 0; c < 8; ++c) a[r][c] = r + c;
  long long t0 = clock64();
  for (int s = 0; s < steps; ++s) {
    if (with_bar) __syncthreads();
#pragma unroll
    for (int s3 = 0; s3 < 3; ++s3) {
      double vv[8], ww[6];
      const double2* p = reinterpret_cast<const double2*>(&v[s3][2 * cg]);
#pragma unroll
      for (int i = 0; i < 4; ++i) { double2 t = p[16 * i]; vv[2 * i] = t.x; vv[2 * i + 1] = t.y; }
      const double2* q = reinterpret_cast<const double2*>(&w[rg][6 * s3]);
#pragma unroll
      for (int i = 0; i < 3; ++i) { double2 t = q[i]; ww[2 * i] = t.x; ww[2 * i + 1] = t.y; }
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) a[r][c] = fma(ww[r], vv[c], a[r][c]);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  double sum = 0; for (int r = 0; r < 6; ++r) for (int c = 0; c < 8; ++c) sum += a[r][c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 4096 * 8);
  long long h[16];
  const char* names[] = {"DFMA dep", "DADD dep", "DMUL dep", "DADD+fmax+fmin dep (3 ops)", "DADD + shfl64 dep", "LDS pointer chase", "__syncthreads", "select(DSETP+DMUL/DADD) dep", "drcp + DADD dep"};
  for (int threads : {32, 320}) {
    lat_kernel<<<1, threads>>>(out, cyc, 1.0000001, 1e-9); cudaDeviceSynchronize();
    lat_kernel<<<1, threads>>>(out, cyc, 1.0000001, 1e-9); cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 9 * 8, cudaMemcpyDeviceToHost);
    printf("--- %d threads (1 CTA): cycles per dependent op ---\n", threads);
    for (int i = 0; i < 9; ++i) printf("  %-32s %.1f\n", names[i], double(h[i]) / (i == 8 ? N / 8 : N));
  }
  for (int bar = 0; bar < 2; ++bar) {
    tile_kernel<<<148, 320>>>(out, cyc, 2000, bar); cudaDeviceSynchronize();
    tile_kernel<<<148, 320>>>(out, cyc, 2000, bar); cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("tile update 6x8, 320 threads/SM, rank-3 step (144 DFMA/thread), barrier=%d: %.0f cycles/step (fp64 pipe floor 720)\n", bar, double(h[0]) / 2000);
  }
  return 0;
}
