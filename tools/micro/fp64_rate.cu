// Microbenchmark: sustained DFMA / FFMA / F2F throughput of one GPU (build: nvcc -arch=sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
template <typename T>
__global__ void fma_kernel(T* out, int iters, T a, T b) {
  T x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void cvt_kernel(float* out, int iters, float a) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  for (int i = 0; i < iters; ++i) {
    double d0 = (double)x0, d1 = (double)x1, d2 = (double)x2, d3 = (double)x3;
    x0 = (float)d0 + a; x1 = (float)d1 + a; x2 = (float)d2 + a; x3 = (float)d3 + a;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}
template <typename F> float time_ms(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
  const int blocks = 148 * 8, threads = 256, iters = 4096;
  double* d; float* f; cudaMalloc(&d, blocks * threads * 8); cudaMalloc(&f, blocks * threads * 4);
  float t64 = time_ms([&] { fma_kernel<double><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9); });
  float t32 = time_ms([&] { fma_kernel<float><<<blocks, threads>>>(f, iters, 1.0000001f, 1e-9f); });
  float tc = time_ms([&] { cvt_kernel<<<blocks, threads>>>(f, iters, 1e-9f); });
  double n = (double)blocks * threads * iters * 8;
  printf("DFMA: %.2f TFLOP/s (%.3f ms)\nFFMA: %.2f TFLOP/s (%.3f ms)\n", 2 * n / t64 * 1e-9, t64, 2 * n / t32 * 1e-9, t32);
  printf("F2F pairs (f32->f64->f32): %.2f Tconv-pairs/s (%.3f ms)\n", (double)blocks * threads * iters * 4 / tc * 1e-9, tc);
  return 0;
}
