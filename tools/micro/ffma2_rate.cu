// FFMA vs FFMA2 (fma.rn.f32x2) issue rate on sm_100a.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  if (MODE == 0) {          // scalar FFMA: 16 independent chains
    float x[16];
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    float s = 0; for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else if (MODE == 1) {   // FFMA2: 8 independent packed chains (same flops)
    unsigned long long x[8], pa = pk(a, a), pb = pk(b, b);
    for (int i = 0; i < 8; ++i) x[i] = pk(threadIdx.x + i, threadIdx.x - i);
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fma2(x[i], pa, pb);
    unsigned long long s = x[0]; for (int i = 1; i < 8; ++i) s = add2(s, x[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(s & 0xffffffffu)) + __uint_as_float((unsigned)(s >> 32));
  } else {                  // FADD2: 8 packed chains
    unsigned long long x[8], pb = pk(b, b);
    for (int i = 0; i < 8; ++i) x[i] = pk(threadIdx.x + i, threadIdx.x - i);
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = add2(x[i], pb);
    unsigned long long s = x[0]; for (int i = 1; i < 8; ++i) s = add2(s, x[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(s & 0xffffffffu));
  }
}
template <int MODE> void run(const char* name, int flops_per_iter) {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  k<MODE><<<148 * 8, 256>>>(out, 100, 1.0001f, 0.5f);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fl = (double)148 * 8 * 256 * iters * flops_per_iter;
  printf("%s: %.3f ms, %.1f TFLOP/s\n", name, ms, fl / ms / 1e9);
}
int main() { run<0>("FFMA  x16", 32); run<1>("FFMA2 x8 ", 32); run<2>("FADD2 x8 ", 16); return 0; }
