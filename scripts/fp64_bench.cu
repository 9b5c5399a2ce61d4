// Microbenchmark: DFMA / FFMA / F2F(f32->f64) issue rates and LDS.128 bandwidth on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
template <typename T, int ILP>
__global__ void fma_kernel(T* out, int iters, T a, T b) {
  T acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = T(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = acc[i] * a + b;
  }
  T s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void cvt_kernel(double* out, const float* in, int iters) {
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = in[threadIdx.x + i];
  double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] += (double)f[i]; f[i] += 1.0f; }
  }
  double t = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
__global__ void lds_kernel(double* out, int iters) {
  __shared__ double2 sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_double2(i, -i);
  __syncthreads();
  double s = 0;
  int idx = threadIdx.x & 3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) { double2 v = sm[(idx * 16 + i + it) & 1023]; s += v.x + v.y; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F>
float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  void* out; cudaMalloc(&out, 148 * 8 * 1024 * 8); float* in; cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
  const int blocks = 148 * 4, threads = 512, iters = 20000;
  for (int rep = 0; rep < 2; ++rep) {
    float ms = timeit([&] { fma_kernel<double, 8><<<blocks, threads>>>((double*)out, iters, 1.0000001, 1e-9); });
    double fl = double(blocks) * threads * iters * 8;
    printf("DFMA: %.3f ms  %.2f T fma/s = %.2f TFLOP/s  (%.1f fma/clk/SM @1.9GHz)\n", ms, fl / ms / 1e9, 2 * fl / ms / 1e9, fl / ms / 1e-3 / 148 / 1.9e9);
    ms = timeit([&] { fma_kernel<float, 8><<<blocks, threads>>>((float*)out, iters, 1.0000001f, 1e-9f); });
    printf("FFMA: %.3f ms  %.2f T fma/s = %.2f TFLOP/s  (%.1f fma/clk/SM @1.9GHz)\n", ms, fl / ms / 1e9, 2 * fl / ms / 1e9, fl / ms / 1e-3 / 148 / 1.9e9);
    ms = timeit([&] { cvt_kernel<<<blocks, threads>>>((double*)out, in, iters / 10); });
    double cv = double(blocks) * threads * (iters / 10) * 8;
    printf("F2F+DADD: %.3f ms  %.2f T cvt/s (%.1f /clk/SM)\n", ms, cv / ms / 1e9, cv / ms / 1e-3 / 148 / 1.9e9);
    ms = timeit([&] { lds_kernel<<<blocks, threads>>>((double*)out, iters / 10); });
    double by = double(blocks) * threads * (iters / 10) * 16.0 * 16;
    printf("LDS.128 (4 addr/warp): %.3f ms  %.1f B/clk/SM requested\n", ms, by / ms / 1e-3 / 148 / 1.9e9);
  }
  return 0;
}
