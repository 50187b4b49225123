// Error plumbing, launch accounting and reference<->padded layout conversion.
#include "common.cuh"

#include <map>
#include <mutex>
#include <utility>

#include <string.h>

#include <vector>

namespace sb {

static thread_local std::string t_last_error;
std::atomic<uint64_t> g_launches{0};

void set_error(const std::string& msg) { t_last_error = msg; }
int fail(int code, const std::string& msg) {
  t_last_error = msg;
  return code;
}

const double TSIT5_A[6][6] = {
    {0.161, 0, 0, 0, 0, 0},
    {-0.008480655492356989, 0.335480655492357, 0, 0, 0, 0},
    {2.8971530571054935, -6.359448489975075, 4.3622954328695815, 0, 0, 0},
    {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 0, 0},
    {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401,
     -0.028269050394068383, 0},
    {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081,
     2.324710524099774}};

// ---------------------------------------------------------------------------------------
// per-kernel event timing
// ---------------------------------------------------------------------------------------
namespace {
struct ProfRec { const char* name; cudaEvent_t a, b; };
struct Prof {
  bool on = false;
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  bool open = false;
  cudaStream_t stream = nullptr;
  cudaEvent_t get() {
    if (used == pool.size()) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
      pool.push_back(e);
    }
    return pool[used++];
  }
};
Prof g_prof;
}  // namespace

void prof_begin(const char* name, cudaStream_t s) {
  if (!g_prof.on || g_prof.recs.size() >= 200000) return;
  ProfRec r{name, g_prof.get(), g_prof.get()};
  if (!r.a || !r.b) return;
  cudaEventRecord(r.a, s);
  g_prof.recs.push_back(r);
  g_prof.open = true;
  g_prof.stream = s;
}

bool prof_enabled() { return g_prof.on; }

int StepGraph::init() {
  if (stream) return 0;
  if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ev_out, cudaEventDisableTiming) != cudaSuccess)
    return fail(SOMAX_B200_ERR_CUDA, "graph stream/event creation failed");
  return 0;
}

void StepGraph::destroy() {
  if (exec) cudaGraphExecDestroy(exec);
  if (stream) cudaStreamDestroy(stream);
  if (ev_in) cudaEventDestroy(ev_in);
  if (ev_out) cudaEventDestroy(ev_out);
  exec = nullptr; stream = nullptr; ev_in = ev_out = nullptr;
}

void prof_end() {
  if (!g_prof.open) return;
  cudaEventRecord(g_prof.recs.back().b, g_prof.stream);
  g_prof.open = false;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of (kernel, DEVICE): a process that
// drives several devices (one host thread per device, as jax.ffi does) must set it on each of
// them, and two threads may arrive here at once.  Largest size set so far per (kernel, device).
int ensure_dyn_smem(const void* fn, size_t bytes) {
  if (bytes <= 48 * 1024) return 0;
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> done;
  int dev = 0;
  SB_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  size_t& cur = done[std::make_pair(fn, dev)];
  if (cur >= bytes) return 0;
  SB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  cur = bytes;
  return 0;
}

int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(SOMAX_B200_ERR_NO_DEVICE,
                "no CUDA device visible: somax_b200 has no CPU fallback");
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess)
    return fail(SOMAX_B200_ERR_NO_DEVICE, "cudaGetDeviceProperties failed");
  if (p.major != 10)
    return fail(SOMAX_B200_ERR_NO_DEVICE,
                std::string("device is sm_") + std::to_string(p.major) + std::to_string(p.minor) +
                    "; this library is built for sm_100a only");
  return 0;
}

// One thread per reference element; x fastest.
template <typename T>
__global__ void pack_kernel(const T* __restrict__ ref, T* __restrict__ pad, int planes, int Ny,
                            int Nx, int pitch) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y;
  int p = blockIdx.z;
  if (i >= pitch) return;
  int c = i - OFF;
  T v = (c >= 0 && c < Nx) ? ref[((size_t)p * Ny + j) * Nx + c] : T(0);
  pad[((size_t)p * Ny + j) * pitch + i] = v;
}

template <typename T>
__global__ void unpack_kernel(const T* __restrict__ pad, T* __restrict__ ref, int planes, int Ny,
                              int Nx, int pitch) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y;
  int p = blockIdx.z;
  if (c >= Nx) return;
  ref[((size_t)p * Ny + j) * Nx + c] = pad[((size_t)p * Ny + j) * pitch + OFF + c];
}

template <typename T>
int pack_field(const T* ref, T* pad, const Layout& L, cudaStream_t s) {
  int planes = L.batch * L.nl;
  dim3 b(256), g((L.pitch + 255) / 256, L.Ny, planes);
  if (g.z > 65535) return fail(SOMAX_B200_ERR_UNSUPPORTED, "batch*nl > 65535");
  prof_begin("pack_kernel", s);
  pack_kernel<T><<<g, b, 0, s>>>(ref, pad, planes, L.Ny, L.Nx, L.pitch);
  SB_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int unpack_field(const T* pad, T* ref, const Layout& L, cudaStream_t s) {
  int planes = L.batch * L.nl;
  dim3 b(256), g((L.Nx + 255) / 256, L.Ny, planes);
  if (g.z > 65535) return fail(SOMAX_B200_ERR_UNSUPPORTED, "batch*nl > 65535");
  prof_begin("unpack_kernel", s);
  unpack_kernel<T><<<g, b, 0, s>>>(pad, ref, planes, L.Ny, L.Nx, L.pitch);
  SB_LAUNCH_CHECK();
  return 0;
}

// Coefficient fields (beta_y, wind, f): stored as a (Ny) profile when x-independent (what
// every reference factory produces), else as a dense (Ny, Nx) array.
template <typename T>
int upload_coef(const double* host, T** dev, int Ny, int Nx, bool* is_1d) {
  bool one = true;
  for (int j = 0; j < Ny && one; ++j)
    for (int i = 1; i < Nx; ++i)
      if (host[(size_t)j * Nx + i] != host[(size_t)j * Nx]) { one = false; break; }
  std::vector<T> tmp;
  if (one) {
    tmp.resize(Ny);
    for (int j = 0; j < Ny; ++j) tmp[j] = (T)host[(size_t)j * Nx];
  } else {
    tmp.resize((size_t)Ny * Nx);
    for (size_t k = 0; k < tmp.size(); ++k) tmp[k] = (T)host[k];
  }
  SB_CUDA(cudaMalloc((void**)dev, tmp.size() * sizeof(T)));
  SB_CUDA(cudaMemcpy(*dev, tmp.data(), tmp.size() * sizeof(T), cudaMemcpyHostToDevice));
  *is_1d = one;
  return 0;
}

template int pack_field<float>(const float*, float*, const Layout&, cudaStream_t);
template int pack_field<double>(const double*, double*, const Layout&, cudaStream_t);
template int unpack_field<float>(const float*, float*, const Layout&, cudaStream_t);
template int unpack_field<double>(const double*, double*, const Layout&, cudaStream_t);
template int upload_coef<float>(const double*, float**, int, int, bool*);
template int upload_coef<double>(const double*, double**, int, int, bool*);

}  // namespace sb

extern "C" {
const char* somax_b200_last_error(void) { return sb::t_last_error.c_str(); }
int somax_b200_abi_version(void) { return SOMAX_B200_ABI_VERSION; }
uint64_t somax_b200_launch_count(void) { return sb::g_launches.load(); }

void somax_b200_profile_enable(int on) { sb::g_prof.on = on != 0; }
void somax_b200_profile_reset(void) { sb::g_prof.recs.clear(); sb::g_prof.used = 0; sb::g_prof.open = false; }

// Aggregates the recorded launches by kernel name into JSON:
// [{"kernel": "...", "launches": n, "total_ms": t}, ...].  Synchronises the device.
int somax_b200_profile_report(char* buf, size_t cap) {
  cudaDeviceSynchronize();
  struct Agg { const char* name; long n; double ms; };
  std::vector<Agg> aggs;
  for (auto& r : sb::g_prof.recs) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
    Agg* a = nullptr;
    for (auto& x : aggs) if (std::string(x.name) == r.name) { a = &x; break; }
    if (!a) { aggs.push_back({r.name, 0, 0.0}); a = &aggs.back(); }
    a->n += 1; a->ms += ms;
  }
  std::string out = "[";
  for (size_t i = 0; i < aggs.size(); ++i) {
    char tmp[256];
    snprintf(tmp, sizeof tmp, "%s{\"kernel\": \"%s\", \"launches\": %ld, \"total_ms\": %.6f}",
             i ? ", " : "", aggs[i].name, aggs[i].n, aggs[i].ms);
    out += tmp;
  }
  out += "]";
  if (out.size() + 1 > cap) return sb::fail(SOMAX_B200_ERR_INVALID, "profile buffer too small");
  memcpy(buf, out.c_str(), out.size() + 1);
  return 0;
}
}
