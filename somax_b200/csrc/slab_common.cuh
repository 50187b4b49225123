// Transport shared by the slab-distributed models (qg_slab.cuh, swm_slab.cuh): segment-table copy
// kernel (peer-memory stores over NVLink), flag barrier in peer memory, CUDA IPC export / attach.
#pragma once
#include <cuda.h>

#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace sb {

constexpr int QGS_MAX_RANKS = 16;

struct Seg {
  const char* src; char* dst;
  unsigned rows, row_bytes;
  size_t spitch, dpitch;
};

template <typename V>
__global__ void __launch_bounds__(256) seg_copy_kernel(const Seg* __restrict__ segs) {
  const Seg sg = segs[blockIdx.y];
  const unsigned vpr = sg.row_bytes / (unsigned)sizeof(V);
  const size_t total = (size_t)sg.rows * vpr;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const unsigned r = (unsigned)(e / vpr), c = (unsigned)(e - (size_t)r * vpr);
    *reinterpret_cast<V*>(sg.dst + (size_t)r * sg.dpitch + (size_t)c * sizeof(V)) =
        *reinterpret_cast<const V*>(sg.src + (size_t)r * sg.spitch + (size_t)c * sizeof(V));
  }
}

struct SegTable {
  Seg* dev = nullptr;
  int n = 0, vec = 16, gx = 1;
};

struct FlagPtrs { unsigned* p[QGS_MAX_RANKS]; };

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Barrier across the ranks of one slab group: thread t publishes `epoch` in rank t's slot for
// this rank and waits for rank t's epoch in its own slot.  Everything this rank stored to peer
// memory earlier in the stream is complete (stream order) and fenced before the flag is released.
static __global__ void slab_barrier_kernel(FlagPtrs F, int me, int nranks, unsigned epoch, unsigned* err,
                                    unsigned long long timeout_ns) {
  const int t = threadIdx.x;
  if (t >= nranks || t == me) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(F.p[t] + me), "r"(epoch) : "memory");
  if (*reinterpret_cast<volatile unsigned*>(err)) return;      // a peer went missing before: do not wait again
  const unsigned long long t0 = gtimer();
  for (;;) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(F.p[me] + t) : "memory");
    if ((int)(v - epoch) >= 0) break;
    if (gtimer() - t0 > timeout_ns) { *err = 1u; break; }   // report instead of hanging the GPU
  }
}


inline int seg_upload(SegTable& t, const std::vector<Seg>& v, size_t* bytes) {
  t.n = (int)v.size();
  if (t.n == 0) return 0;
  t.vec = 16;
  size_t maxb = 0;
  for (const Seg& s : v) {
    const size_t m = (size_t)s.src | (size_t)s.dst | s.row_bytes | s.spitch | s.dpitch;
    if (m & 15) t.vec = std::min(t.vec, (m & 7) ? 4 : 8);
    maxb = std::max(maxb, (size_t)s.rows * s.row_bytes);
  }
  t.gx = (int)std::min<size_t>(64, std::max<size_t>(1, maxb / t.vec / (256 * 8)));
  SB_CUDA(cudaMalloc((void**)&t.dev, v.size() * sizeof(Seg)));
  SB_CUDA(cudaMemcpy(t.dev, v.data(), v.size() * sizeof(Seg), cudaMemcpyHostToDevice));
  *bytes += v.size() * sizeof(Seg);
  return 0;
}

inline int seg_launch(const char* tag, const SegTable& t, cudaStream_t s) {
  if (t.n == 0) return 0;
  prof_begin(tag, s);
  const dim3 grid(t.gx, t.n);
  if (t.vec == 16) seg_copy_kernel<int4><<<grid, 256, 0, s>>>(t.dev);
  else if (t.vec == 8) seg_copy_kernel<unsigned long long><<<grid, 256, 0, s>>>(t.dev);
  else seg_copy_kernel<unsigned><<<grid, 256, 0, s>>>(t.dev);
  SB_LAUNCH_CHECK();
  return 0;
}


// ---- CUDA IPC: a rank exports the buffers its peers store into; the blobs are plain bytes ----
constexpr int SLAB_MAX_BUF = 12;
struct SlabIpcBlob {
  cudaIpcMemHandle_t handle[SLAB_MAX_BUF];
  unsigned long long offset[SLAB_MAX_BUF];
};

inline int slab_ipc_export(void* const* ptrs, int n, void* blob) {
  SlabIpcBlob b;
  memset(&b, 0, sizeof(b));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  typedef CUresult (*range_fn)(CUdeviceptr*, size_t*, CUdeviceptr);
  range_fn get_range = nullptr;
  if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess && fn)
    get_range = reinterpret_cast<range_fn>(fn);
  else
    cudaGetLastError();
  for (int i = 0; i < n; ++i) {
    cudaError_t e = cudaIpcGetMemHandle(&b.handle[i], ptrs[i]);
    if (e != cudaSuccess) return fail(SOMAX_B200_ERR_COMM, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    CUdeviceptr base = 0; size_t sz = 0;
    if (get_range && get_range(&base, &sz, (CUdeviceptr)ptrs[i]) == CUDA_SUCCESS)
      b.offset[i] = (unsigned long long)((CUdeviceptr)ptrs[i] - base);
  }
  memcpy(blob, &b, sizeof(b));
  return 0;
}

// peers[r][i]: buffer i of rank r as mapped in this process (rank `me` = the local pointers)
inline int slab_ipc_attach(const void* blobs, int nranks, int me, int n, void* const* mine,
                           void* (*peers)[SLAB_MAX_BUF], std::vector<void*>& opened) {
  const SlabIpcBlob* B = reinterpret_cast<const SlabIpcBlob*>(blobs);
  for (int r = 0; r < nranks; ++r) {
    if (r == me) { for (int i = 0; i < n; ++i) peers[r][i] = mine[i]; continue; }
    for (int i = 0; i < n; ++i) {
      void* p = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&p, B[r].handle[i], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess)
        return fail(SOMAX_B200_ERR_COMM, std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(r) + "): " + cudaGetErrorString(e));
      opened.push_back(p);
      peers[r][i] = (char*)p + B[r].offset[i];
    }
  }
  return 0;
}

}  // namespace sb
