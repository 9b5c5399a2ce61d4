// Experiment: does the FP64 tensor path (legacy mma.sync DMMA m8n8k4) buy anything over
// DFMA on B200?  Same arithmetic volume through both: N independent accumulator tiles per warp, long loops.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/dmma_bench scripts/dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int TILES>
__global__ void dmma_kernel(double* out, int iters, double a0, double b0) {
  // m8n8k4: A 8x4 (one double per lane), B 4x8 (one per lane), C/D 8x8 (two per lane)
  double c[TILES][2];
#pragma unroll
  for (int t = 0; t < TILES; ++t) { c[t][0] = threadIdx.x * 1e-3 + t; c[t][1] = -c[t][0]; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < TILES; ++t) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < TILES; ++t) s += c[t][0] + c[t][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = double(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
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
  printf("%s SMs %d\n", p.name, p.multiProcessorCount);
  double* out; cudaMalloc(&out, 148 * 8 * 1024 * 8);
  const int blocks = p.multiProcessorCount * 4, threads = 512, iters = 20000;
  for (int rep = 0; rep < 2; ++rep) {
    float ms = timeit([&] { dfma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double fl = 2.0 * blocks * threads * double(iters) * 8;
    printf("DFMA x8 chains : %.3f ms  %.2f TFLOP/s, %.1f warp-instr/clk/SM issued\n", ms, fl / ms / 1e9,
           double(blocks) * (threads / 32) * double(iters) * 8 / (ms * 1e-3 * 1.965e9 * p.multiProcessorCount));
    ms = timeit([&] { dmma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-3); });
    // one m8n8k4 = 8*8*4 = 256 FMA = 512 flop per warp instruction
    double fm = 512.0 * blocks * (threads / 32) * double(iters) * 8;
    printf("DMMA m8n8k4 x8 : %.3f ms  %.2f TFLOP/s, %.2f warp-instr/clk/SM issued\n", ms, fm / ms / 1e9,
           double(blocks) * (threads / 32) * double(iters) * 8 / (ms * 1e-3 * 1.965e9 * p.multiProcessorCount));
    ms = timeit([&] { dmma_kernel<2><<<blocks, threads>>>(out, iters, 1.0000001, 1e-3); });
    fm = 512.0 * blocks * (threads / 32) * double(iters) * 2;
    printf("DMMA m8n8k4 x2 : %.3f ms  %.2f TFLOP/s (two dependent chains per warp: latency %.1f cycles per mma)\n", ms,
           fm / ms / 1e9, ms * 1e-3 * 1.965e9 / iters / 2 * 2);
  }
  return 0;
}
