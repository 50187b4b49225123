// Shared declarations for libsomax_b200: error handling, the padded HBM layout, launch
// accounting and the Tsit5 stage descriptor.  sm_100a only; there is no CPU fallback.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/somax_b200.h"

namespace sb {

// ---------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
extern std::atomic<uint64_t> g_launches;

#define SB_CUDA(expr)                                                                     \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return ::sb::fail(SOMAX_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// Optional per-kernel CUDA-event timing (somax_b200_profile_*): prof_begin() before a launch,
// SB_LAUNCH_CHECK() closes the record.  No-ops unless profiling is enabled.
void prof_begin(const char* name, cudaStream_t s);
void prof_end();
bool prof_enabled();

// CUDA-graph replay of the stepping loop for launch-bound (small) grids.  The stage buffers
// rotate with period 2 steps, so two consecutive full steps (12 RHS evaluations, including the
// two-stream fork/join of the y-sweeps) are captured once per (dt, params) and replayed.
struct StepGraph {
  cudaGraphExec_t exec = nullptr;
  double key[6] = {0, 0, 0, 0, 0, 0};
  uint64_t nlaunch = 0;            // kernels per replay
  cudaStream_t stream = nullptr;   // capture / replay stream (the legacy stream cannot capture)
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  int init();
  void destroy();
};
// grids at or below this many cells per handle use the graph path
constexpr size_t GRAPH_MAX_CELLS = (size_t)1 << 21;

#define SB_LAUNCH_CHECK()                                                                 \
  do {                                                                                    \
    ::sb::prof_end();                                                                     \
    ::sb::g_launches.fetch_add(1, std::memory_order_relaxed);                             \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess)                                                                \
      return ::sb::fail(SOMAX_B200_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(_e)); \
  } while (0)

int require_device();
// Opt a kernel in to `bytes` of dynamic shared memory on the CURRENT device (per device, mutex-guarded).
int ensure_dyn_smem(const void* fn, size_t bytes);

// ---------------------------------------------------------------------------------------
// HBM layout of a field inside a handle ("padded layout").
//
// Reference layout is dense (batch, nl, Ny, Nx).  Internally every row is stored with a left
// pad of OFF = 3 elements so that interior column i = 1 sits on a 16-byte boundary, and the
// row pitch is a multiple of 4 elements: element (b,l,j,i) lives at
//   ((b*nl + l)*Ny + j)*pitch + OFF + i.
// 128-bit vector loads/stores therefore work on every row for fp32 and fp64.  Pad slots
// are kept at zero.
// ---------------------------------------------------------------------------------------
constexpr int OFF = 3;

struct Layout {
  int batch, nl, Ny, Nx, pitch;
  __host__ __device__ size_t plane() const { return (size_t)Ny * pitch; }
  __host__ __device__ size_t count() const { return (size_t)batch * nl * plane(); }
  __host__ __device__ int groups() const { return pitch / 4; }
};

inline Layout make_layout(int batch, int nl, int ny, int nx) {
  Layout L;
  L.batch = batch; L.nl = nl; L.Ny = ny + 2; L.Nx = nx + 2;
  L.pitch = ((L.Nx + OFF + 3) / 4) * 4;
  return L;
}

// ---------------------------------------------------------------------------------------
// Tsit5 stage descriptor (reference core/model.py:75-88 via diffrax).
// One RHS kernel evaluates F = f(BC(Yin)) and, in its epilogue, forms the next stage state
//   Yout = base + sum_j a[j]*(dt*Fprev[j]) + a_new*(dt*F)
// where base is `y` (the step's start state) or Yin itself (first/last evaluation of a step).
// F (not dt*F) is stored so that a clipped last step may change dt.
// ---------------------------------------------------------------------------------------
constexpr int MAX_FIELDS = 3;  // QG: q ; SWM: h,u,v
constexpr int MAX_PREV = 5;

template <typename T>
struct Stage {
  int nfields;
  int nprev;                       // number of previously stored F arrays entering Yout
  const T* Yin[MAX_FIELDS];        // stage state the RHS is evaluated at (BC applied on load)
  const T* y[MAX_FIELDS];          // step start state; nullptr => base is Yin
  const T* Fprev[MAX_PREV][MAX_FIELDS];
  T* Fout[MAX_FIELDS];             // nullptr => do not store F
  T* Yout[MAX_FIELDS];             // nullptr => RHS only
  T a[MAX_PREV];
  T a_new;
  T dt;
  T adt[MAX_PREV];                 // a[j]*dt and a_new*dt, rounded once on the host: the fast
  T adt_new;                       // kernels use them (<= 1 ulp from a*(dt*k))
};

template <typename T>
inline void stage_finalize(Stage<T>& st, double dt) {
  for (int j = 0; j < MAX_PREV; ++j) st.adt[j] = (T)((double)st.a[j] * dt);
  st.adt_new = (T)((double)st.a_new * dt);
}

// Tsit5 tableau rows (a_{s,1..s-1}), s = 2..7.
extern const double TSIT5_A[6][6];

template <typename T>
__device__ __forceinline__ T rk_combine(const Stage<T>& st, int f, size_t idx, T yin, T F) {
  // mimics oracle.tsit5.tree_axpy: acc = base; acc += a_j * (dt*F_j) in order; then the new one
  T acc = st.y[f] ? st.y[f][idx] : yin;
#pragma unroll
  for (int j = 0; j < MAX_PREV; ++j)
    if (j < st.nprev) acc = acc + st.a[j] * (st.dt * st.Fprev[j][f][idx]);
  acc = acc + st.a_new * (st.dt * F);
  return acc;
}

// vector (4-wide) variant used by the stencil kernels
template <typename T> struct Vec4 { T x, y, z, w; };
template <> struct __align__(16) Vec4<float> { float x, y, z, w; };
template <> struct __align__(32) Vec4<double> { double x, y, z, w; };

template <typename T>
__device__ __forceinline__ Vec4<T> ld4(const T* p) { return *reinterpret_cast<const Vec4<T>*>(p); }
template <typename T>
__device__ __forceinline__ void st4(T* p, const Vec4<T>& v) { *reinterpret_cast<Vec4<T>*>(p) = v; }

template <typename T>
__device__ __forceinline__ void rk_epilogue4(const Stage<T>& st, int f, size_t idx,
                                             const Vec4<T>& yin, const Vec4<T>& F) {
  if (st.Fout[f]) st4(st.Fout[f] + idx, F);
  if (!st.Yout[f]) return;
  Vec4<T> acc = st.y[f] ? ld4(st.y[f] + idx) : yin;
#pragma unroll
  for (int j = 0; j < MAX_PREV; ++j) {
    if (j < st.nprev) {
      Vec4<T> k = ld4(st.Fprev[j][f] + idx);
      acc.x = acc.x + st.a[j] * (st.dt * k.x);
      acc.y = acc.y + st.a[j] * (st.dt * k.y);
      acc.z = acc.z + st.a[j] * (st.dt * k.z);
      acc.w = acc.w + st.a[j] * (st.dt * k.w);
    }
  }
  acc.x = acc.x + st.a_new * (st.dt * F.x);
  acc.y = acc.y + st.a_new * (st.dt * F.y);
  acc.z = acc.z + st.a_new * (st.dt * F.z);
  acc.w = acc.w + st.a_new * (st.dt * F.w);
  st4(st.Yout[f] + idx, acc);
}

// Split variant: rk_prefetch4 issues every global load of the epilogue (base state and the
// previous stage derivatives) up front so their latency overlaps the stencil work.
template <typename T>
struct RkRegs { Vec4<T> base; Vec4<T> k[MAX_PREV]; };

template <typename T>
__device__ __forceinline__ RkRegs<T> rk_prefetch4(const Stage<T>& st, int f, size_t idx, bool valid) {
  RkRegs<T> R;
  const Vec4<T> z{0, 0, 0, 0};
  R.base = z;
#pragma unroll
  for (int j = 0; j < MAX_PREV; ++j) R.k[j] = z;
  if (valid && st.Yout[f]) {
    R.base = ld4((st.y[f] ? st.y[f] : st.Yin[f]) + idx);
#pragma unroll
    for (int j = 0; j < MAX_PREV; ++j)
      if (j < st.nprev) R.k[j] = ld4(st.Fprev[j][f] + idx);
  }
  return R;
}

// Fast epilogue: loads batched, products with the host-rounded a*dt.
template <typename T>
__device__ __forceinline__ void rk_epilogue4_fast(const Stage<T>& st, int f, size_t idx, const Vec4<T>& F) {
  if (st.Fout[f]) st4(st.Fout[f] + idx, F);
  if (!st.Yout[f]) return;
  Vec4<T> acc = ld4((st.y[f] ? st.y[f] : st.Yin[f]) + idx);
  Vec4<T> k[MAX_PREV];
#pragma unroll
  for (int j = 0; j < MAX_PREV; ++j)
    if (j < st.nprev) k[j] = ld4(st.Fprev[j][f] + idx);
#pragma unroll
  for (int j = 0; j < MAX_PREV; ++j) {
    if (j < st.nprev) {
      acc.x = fma(st.adt[j], k[j].x, acc.x); acc.y = fma(st.adt[j], k[j].y, acc.y);
      acc.z = fma(st.adt[j], k[j].z, acc.z); acc.w = fma(st.adt[j], k[j].w, acc.w);
    }
  }
  acc.x = fma(st.adt_new, F.x, acc.x); acc.y = fma(st.adt_new, F.y, acc.y);
  acc.z = fma(st.adt_new, F.z, acc.z); acc.w = fma(st.adt_new, F.w, acc.w);
  st4(st.Yout[f] + idx, acc);
}

template <typename T>
__device__ __forceinline__ void rk_finish4(const Stage<T>& st, int f, size_t idx, const RkRegs<T>& R,
                                           const Vec4<T>& F) {
  if (st.Fout[f]) st4(st.Fout[f] + idx, F);
  if (!st.Yout[f]) return;
  Vec4<T> acc = R.base;
#pragma unroll
  for (int j = 0; j < MAX_PREV; ++j) {
    if (j < st.nprev) {
      acc.x = acc.x + st.a[j] * (st.dt * R.k[j].x);
      acc.y = acc.y + st.a[j] * (st.dt * R.k[j].y);
      acc.z = acc.z + st.a[j] * (st.dt * R.k[j].z);
      acc.w = acc.w + st.a[j] * (st.dt * R.k[j].w);
    }
  }
  acc.x = acc.x + st.a_new * (st.dt * F.x);
  acc.y = acc.y + st.a_new * (st.dt * F.y);
  acc.z = acc.z + st.a_new * (st.dt * F.z);
  acc.w = acc.w + st.a_new * (st.dt * F.w);
  st4(st.Yout[f] + idx, acc);
}

// ---------------------------------------------------------------------------------------
// layout conversion kernels (layout.cu)
// ---------------------------------------------------------------------------------------
template <typename T>
int pack_field(const T* ref, T* pad, const Layout& L, cudaStream_t s);     // reference -> padded
template <typename T>
int unpack_field(const T* pad, T* ref, const Layout& L, cudaStream_t s);   // padded -> reference
template <typename T>
int upload_coef(const double* host, T** dev, int Ny, int Nx, bool* is_1d);  // (Ny,Nx) host double

}  // namespace sb
