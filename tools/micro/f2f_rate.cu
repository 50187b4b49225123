// Throughput of fp32<->fp64 conversions (F2F) vs FFMA on one GPU, full occupancy.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void widen(double* out, const float* in, int iters) {
  float x0 = in[threadIdx.x], x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int i = 0; i < iters; ++i) {
    a0 += (double)x0; a1 += (double)x1; a2 += (double)x2; a3 += (double)x3;
    x0 = __int_as_float(__float_as_int(x0) ^ i); x1 = __int_as_float(__float_as_int(x1) ^ i);
    x2 = __int_as_float(__float_as_int(x2) ^ i); x3 = __int_as_float(__float_as_int(x3) ^ i);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}
__global__ void narrow(float* out, const double* in, int iters) {
  double x0 = in[threadIdx.x], x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int i = 0; i < iters; ++i) {
    a0 += (float)x0; a1 += (float)x1; a2 += (float)x2; a3 += (float)x3;
    x0 += 1.0; x1 += 1.0; x2 += 1.0; x3 += 1.0;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
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
  cudaMemset(d, 0, blocks * threads * 8); cudaMemset(f, 0, blocks * threads * 4);
  float tw = time_ms([&] { widen<<<blocks, threads>>>(d, f, iters); });
  float tn = time_ms([&] { narrow<<<blocks, threads>>>(f, d, iters); });
  double n = (double)blocks * threads * iters * 4;
  double clk = 1.96e9 * 148;
  printf("widen  F2F.F64.F32 (+DADD,+LOP): %.3f ms -> %.1f conv/clk/SM\n", tw, n / (tw * 1e-3) / clk);
  printf("narrow F2F.F32.F64 (+FADD,+DADD): %.3f ms -> %.1f conv/clk/SM\n", tn, n / (tn * 1e-3) / clk);
  return 0;
}
