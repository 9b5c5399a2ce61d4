// Microbenchmark: how fast do the two FP64 register-tile patterns of the solver issue on one SM?
//   matvec : s[r] = fma(a[r][c], v[c], s[r])          15 x 4 tile, 60 DFMA, a[][] all distinct registers
//   rank1  : a[r][c] = fma(w[r], v[c], a[r][c])       15 x 4 tile, 60 DFMA per rank-1 term
// One CTA per SM, W warps (W / 4 per sub-partition); cycles per 60-DFMA block as seen by a warp.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256, 1) tile_kernel(double* out, const double* in, int iters, long long* cyc) {
  __shared__ double sv[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sv[i] = in[i & 255];
  double a[15][4], s[15];
#pragma unroll
  for (int r = 0; r < 15; ++r) {
    s[r] = 0.0;
#pragma unroll
    for (int c = 0; c < 4; ++c) a[r][c] = in[(threadIdx.x + 7 * r + c) & 255];
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const double2 v01 = *reinterpret_cast<const double2*>(&sv[(4 * it) & 1020]);
    const double2 v23 = *reinterpret_cast<const double2*>(&sv[(4 * it + 2) & 1020]);
    const double v[4] = {v01.x, v01.y, v23.x, v23.y};
    if (MODE == 0) {
#pragma unroll
      for (int r = 0; r < 15; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) s[r] = fma(a[r][c], v[c], s[r]);
    } else {
      double w[16];
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        const double2 t = *reinterpret_cast<const double2*>(&sv[(16 * it + 2 * h + 64) & 1022]);
        w[2 * h] = t.x; w[2 * h + 1] = t.y;
      }
#pragma unroll
      for (int r = 0; r < 15; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) a[r][c] = fma(w[r], v[c], a[r][c]);
    }
  }
  const long long t1 = clock64();
  double t = 0;
#pragma unroll
  for (int r = 0; r < 15; ++r) { t += s[r]; 
#pragma unroll
    for (int c = 0; c < 4; ++c) t += a[r][c]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  double *out, *in; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&in, 256 * 8); cudaMalloc(&cyc, 148 * 8);
  double h[256]; for (int i = 0; i < 256; ++i) h[i] = 1e-3 * (i % 17) - 5e-3;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {1, 2, 4, 8}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) tile_kernel<0><<<148, 32 * warps>>>(out, in, iters, cyc);
        else tile_kernel<1><<<148, 32 * warps>>>(out, in, iters, cyc);
        cudaDeviceSynchronize();
      }
      long long hc[148]; cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
      double m = 0; for (int i = 0; i < 148; ++i) m += double(hc[i]) / 148;
      printf("%s warps/SM %2d: %.1f cycles per 60-DFMA block per warp  (%.2f cycles per DFMA; SM rate %.1f DFMA lanes/clk)\n",
             mode == 0 ? "matvec" : "rank1 ", warps, m / iters, m / iters / 60, 60.0 * 32 * warps / (m / iters));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
