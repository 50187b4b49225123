// Latency / throughput of the fp64-pipe instructions the Thomas sweep uses, at LOW occupancy
// (grid = 148*5 blocks of 64 threads, like the real kernel).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chain_dfma(double* out, int n, double a) {
  double x = threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = fma(-a, x, 1.0);
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + (double)(t1 - t0) * 1e-30;
  if (blockIdx.x == 0 && threadIdx.x == 0) printf("dependent DFMA: %.1f cycles each\n", (double)(t1 - t0) / n);
}
__global__ void chain_ffma(float* out, int n, float a) {
  float x = threadIdx.x * 1e-3f;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = fmaf(-a, x, 1.0f);
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + (float)(t1 - t0) * 1e-30f;
  if (blockIdx.x == 0 && threadIdx.x == 0) printf("dependent FFMA: %.1f cycles each\n", (double)(t1 - t0) / n);
}
// the real loop body: 8 rows: cvt, dmul, dfma chain, cvt back (register-only)
__global__ void body(float* buf, int n, double c, double k) {
  float v[8];
  for (int r = 0; r < 8; ++r) v[r] = buf[(blockIdx.x * 8 + r) * 64 + threadIdx.x];
  double carry = 0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) { carry = fma(-c, carry, k * (double)v[r]); v[r] = (float)carry; }
  }
  long long t1 = clock64();
  for (int r = 0; r < 8; ++r) buf[(blockIdx.x * 8 + r) * 64 + threadIdx.x] = v[r];
  if (blockIdx.x == 0 && threadIdx.x == 0) printf("loop body (8 rows): %.1f cycles per row\n", (double)(t1 - t0) / n / 8);
}
__global__ void cvt_tp(float* out, int n) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  double acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    acc += (double)x0; acc += (double)x1; acc += (double)x2; acc += (double)x3;
    x0 += 1.f; x1 += 1.f; x2 += 1.f; x3 += 1.f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
  if (blockIdx.x == 0 && threadIdx.x == 0) printf("4x(F2F.F64.F32 + DADD) per iter: %.1f cycles per iter\n", (double)(t1 - t0) / n);
}
int main() {
  double* d; float* f;
  cudaMalloc(&d, 1 << 24); cudaMalloc(&f, 1 << 24); cudaMemset(f, 0, 1 << 24);
  chain_dfma<<<148 * 5, 64>>>(d, 4096, 0.5); cudaDeviceSynchronize();
  chain_ffma<<<148 * 5, 64>>>(f, 4096, 0.5f); cudaDeviceSynchronize();
  body<<<148 * 5, 64>>>(f, 512, 0.5, 1e-3); cudaDeviceSynchronize();
  cvt_tp<<<148 * 5, 64>>>(f, 4096); cudaDeviceSynchronize();
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a); body<<<148 * 5, 64>>>(f, 512, 0.5, 1e-3); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("body kernel: %.3f ms for %d row-iterations/thread\n", ms, 512 * 8);
  return 0;
}
