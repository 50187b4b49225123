// jax.ffi custom-call shim over the C ABI of libsomax_b200.so: one XLA FFI handler per entry point a
// jax-side `SomaxModel` needs (core/model.py:47-88 is the call site they serve).
//
// The XLA FFI headers (`jax.ffi.include_dir()`) are not in this image (no jax / jaxlib, SURVEY.md
// section 0-5), so the shim cannot be BUILT here; tests/test_abi.py compile-checks it against a
// stand-in header (tests/stubs/xla/ffi/api/ffi.h, same names and call shapes).  Build where JAX is:
//   g++ -O2 -fPIC -shared -std=c++17 -I$(python -c "import jax; print(jax.ffi.include_dir())")
//       -Iinclude -I$CUDA_HOME/include somax_b200/csrc/jax_ffi_shim.cc -Lsomax_b200/lib -lsomax_b200
//       -lcudart -o somax_b200/lib/libsomax_b200_jax.so
// and register the handlers as shown in INTEGRATION.md.
//
// A handler unpacks buffers / attributes / the CUDA stream and forwards to the C ABI.  Setup arrays
// (mode matrices, coefficient fields) arrive as host-resident f64 attributes - exactly what
// *_create takes.  Handles are cached per (device, shape, dtype, flags, coefficient hash) under a
// mutex: XLA calls handlers from one host thread per device, concurrently.  Handlers never throw;
// errors come back as ffi::Error with the library's thread-local message.
#include <cuda_runtime_api.h>

#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <tuple>

#include "somax_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using F64s = ffi::Span<const double>;

namespace {

struct Key {
  int kind, device, dtype, batch, nl, ny, nx, bc;
  unsigned spec;
  double dx, dy;
  uint64_t coef_hash;
  bool operator<(const Key& o) const {
    return std::tie(kind, device, dtype, batch, nl, ny, nx, bc, spec, dx, dy, coef_hash) <
           std::tie(o.kind, o.device, o.dtype, o.batch, o.nl, o.ny, o.nx, o.bc, o.spec, o.dx, o.dy, o.coef_hash);
  }
};

std::mutex g_mu;
std::map<Key, void*> g_handles;      // somax_b200_qg_t / somax_b200_swm_t

uint64_t fnv(F64s a, uint64_t h = 1469598103934665603ull) {
  const unsigned char* c = reinterpret_cast<const unsigned char*>(a.begin());
  for (size_t i = 0; i < a.size() * sizeof(double); ++i) h = (h ^ c[i]) * 1099511628211ull;
  return h;
}

ffi::Error fail(const char* what) {
  return ffi::Error(ffi::ErrorCode::kInternal, std::string(what) + ": " + somax_b200_last_error());
}

ffi::Error bad(const char* msg) { return ffi::Error(ffi::ErrorCode::kInvalidArgument, msg); }

// (.., nl, Ny, Nx) or, for the single-layer models, (.., Ny, Nx) with `layered` = 0
bool shape_key(const ffi::AnyBuffer& a, int64_t layered, Key* k) {
  auto d = a.dimensions();
  const int nd = static_cast<int>(d.size()), base = layered ? 3 : 2;
  if (nd != base && nd != base + 1) return false;
  if (a.element_type() != ffi::DataType::F32 && a.element_type() != ffi::DataType::F64) return false;
  k->dtype = a.element_type() == ffi::DataType::F32 ? SOMAX_B200_F32 : SOMAX_B200_F64;
  k->nx = static_cast<int>(d[nd - 1]) - 2;
  k->ny = static_cast<int>(d[nd - 2]) - 2;
  k->nl = layered ? static_cast<int>(d[nd - 3]) : 1;
  k->batch = nd == base + 1 ? static_cast<int>(d[0]) : 1;
  cudaGetDevice(&k->device);
  return true;
}

somax_b200_qg_t get_qg(Key key, double dx, double dy, int64_t spec, F64s Cl2m, F64s Cm2l, F64s lambdas,
                       F64s beta_y, F64s wind) {
  key.kind = 0; key.bc = 0; key.spec = static_cast<unsigned>(spec); key.dx = dx; key.dy = dy;
  key.coef_hash = fnv(wind, fnv(beta_y, fnv(lambdas, fnv(Cm2l, fnv(Cl2m)))));
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_handles.find(key);
  if (it != g_handles.end()) return static_cast<somax_b200_qg_t>(it->second);
  somax_b200_qg_t h = nullptr;
  if (somax_b200_qg_create(&h, key.dtype, key.batch, key.nl, key.ny, key.nx, dx, dy, Cl2m.begin(), Cm2l.begin(),
                           lambdas.begin(), beta_y.begin(), wind.begin(), SOMAX_B200_SOLVER_AUTO, key.spec) != 0)
    return nullptr;
  g_handles[key] = h;
  return h;
}

somax_b200_swm_t get_swm(Key key, double dx, double dy, int64_t bc, int64_t spec, F64s g_prime, F64s f_field,
                         F64s wind_x, F64s wind_y) {
  key.kind = 1; key.bc = static_cast<int>(bc); key.spec = static_cast<unsigned>(spec); key.dx = dx; key.dy = dy;
  key.coef_hash = fnv(wind_y, fnv(wind_x, fnv(f_field, fnv(g_prime))));
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_handles.find(key);
  if (it != g_handles.end()) return static_cast<somax_b200_swm_t>(it->second);
  somax_b200_swm_t h = nullptr;
  if (somax_b200_swm_create(&h, key.dtype, key.batch, key.nl, key.ny, key.nx, dx, dy, key.bc, g_prime.begin(),
                            f_field.begin(), wind_x.begin(), wind_y.begin(), key.spec) != 0)
    return nullptr;
  g_handles[key] = h;
  return h;
}

bool copy_d2d(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  return dst == src || cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s) == cudaSuccess;
}

#define SB_QG_SETUP                                                                                       \
  Key key{};                                                                                              \
  if (!shape_key(q, layered, &key)) return bad("q must be ([members,] [nl,] Ny, Nx) in f32 or f64");       \
  somax_b200_qg_t h = get_qg(key, dx, dy, spec, Cl2m, Cm2l, lambdas, beta_y, wind);                       \
  if (!h) return fail("somax_b200_qg_create")

// SomaxModel.integrate (core/model.py:53-88) for BaroclinicQG / BarotropicQG: q after n_steps Tsit5
// steps (+ one clipped step of dt_last); resume != 0 continues from a saved state (no BC on it).
ffi::Error QgStepsImpl(cudaStream_t stream, ffi::AnyBuffer q, ffi::Result<ffi::AnyBuffer> out, int64_t n_steps,
                       double dt, double dt_last, int64_t resume, double nu, double kappa, double tau0, double H0,
                       double dx, double dy, int64_t layered, int64_t spec, F64s Cl2m, F64s Cm2l, F64s lambdas,
                       F64s beta_y, F64s wind) {
  SB_QG_SETUP;
  if (!copy_d2d(out->untyped_data(), q.untyped_data(), q.size_bytes(), stream)) return bad("cudaMemcpyAsync failed");
  somax_b200_params p{nu, kappa, tau0, H0};
  const int rc = resume ? somax_b200_qg_resume(h, out->untyped_data(), n_steps, dt, dt_last, &p, stream)
                        : somax_b200_qg_steps(h, out->untyped_data(), n_steps, dt, dt_last, &p, stream);
  return rc ? fail("somax_b200_qg_steps") : ffi::Error::Success();
}

// vector_field(q), or with apply_bc != 0 the `_rhs` diffrax sees: vector_field(BC(q)) (core/model.py:47-51).
ffi::Error QgRhsImpl(cudaStream_t stream, ffi::AnyBuffer q, ffi::Result<ffi::AnyBuffer> dq, int64_t apply_bc,
                     double nu, double kappa, double tau0, double H0, double dx, double dy, int64_t layered,
                     int64_t spec, F64s Cl2m, F64s Cm2l, F64s lambdas, F64s beta_y, F64s wind) {
  SB_QG_SETUP;
  somax_b200_params p{nu, kappa, tau0, H0};
  if (somax_b200_qg_rhs(h, q.untyped_data(), dq->untyped_data(), nullptr, &p, static_cast<int>(apply_bc), stream))
    return fail("somax_b200_qg_rhs");
  return ffi::Error::Success();
}

// BaroclinicQG._invert_pv / BarotropicQG._invert_pv (qg/baroclinic.py:135-159, qg/barotropic.py:113-121).
ffi::Error QgInvertImpl(cudaStream_t stream, ffi::AnyBuffer q, ffi::Result<ffi::AnyBuffer> psi, double dx, double dy,
                        int64_t layered, int64_t spec, F64s Cl2m, F64s Cm2l, F64s lambdas, F64s beta_y, F64s wind) {
  SB_QG_SETUP;
  if (somax_b200_qg_invert(h, q.untyped_data(), psi->untyped_data(), stream)) return fail("somax_b200_qg_invert");
  return ffi::Error::Success();
}

// Scalars of diagnose() (qg/baroclinic.py:197-228): out f64 [members, 2 nl + 1] = KE, enstrophy, non-finite count.
ffi::Error QgDiagImpl(cudaStream_t stream, ffi::AnyBuffer q, ffi::Result<ffi::AnyBuffer> out, double dx, double dy,
                      int64_t layered, int64_t spec, F64s Cl2m, F64s Cm2l, F64s lambdas, F64s beta_y, F64s wind) {
  SB_QG_SETUP;
  if (out->element_type() != ffi::DataType::F64) return bad("diagnostics buffer must be f64");
  if (somax_b200_qg_diag(h, q.untyped_data(), static_cast<double*>(out->untyped_data()), stream))
    return fail("somax_b200_qg_diag");
  return ffi::Error::Success();
}

#define SB_SWM_SETUP                                                                                      \
  Key key{};                                                                                              \
  if (!shape_key(hh, layered, &key)) return bad("h, u, v must be ([members,] [nl,] Ny, Nx) in f32 or f64"); \
  if (u.size_bytes() != hh.size_bytes() || v.size_bytes() != hh.size_bytes()) return bad("h, u, v differ in size"); \
  somax_b200_swm_t h = get_swm(key, dx, dy, bc, spec, g_prime, f_field, wind_x, wind_y);                  \
  if (!h) return fail("somax_b200_swm_create")

// SomaxModel.integrate for MultilayerShallowWater2D / NonlinearShallowWater2D.
ffi::Error SwmStepsImpl(cudaStream_t stream, ffi::AnyBuffer hh, ffi::AnyBuffer u, ffi::AnyBuffer v,
                        ffi::Result<ffi::AnyBuffer> ho, ffi::Result<ffi::AnyBuffer> uo, ffi::Result<ffi::AnyBuffer> vo,
                        int64_t n_steps, double dt, double dt_last, int64_t resume, double nu, double kappa,
                        double tau0, double H0, double dx, double dy, int64_t layered, int64_t bc, int64_t spec,
                        F64s g_prime, F64s f_field, F64s wind_x, F64s wind_y) {
  SB_SWM_SETUP;
  const size_t nb = hh.size_bytes();
  if (!copy_d2d(ho->untyped_data(), hh.untyped_data(), nb, stream) ||
      !copy_d2d(uo->untyped_data(), u.untyped_data(), nb, stream) ||
      !copy_d2d(vo->untyped_data(), v.untyped_data(), nb, stream))
    return bad("cudaMemcpyAsync failed");
  somax_b200_params p{nu, kappa, tau0, H0};
  const int rc = resume ? somax_b200_swm_resume(h, ho->untyped_data(), uo->untyped_data(), vo->untyped_data(), n_steps,
                                                dt, dt_last, &p, stream)
                        : somax_b200_swm_steps(h, ho->untyped_data(), uo->untyped_data(), vo->untyped_data(), n_steps,
                                               dt, dt_last, &p, stream);
  return rc ? fail("somax_b200_swm_steps") : ffi::Error::Success();
}

// MultilayerShallowWater2D.vector_field (swm/multilayer.py:150-201), optionally of BC(state).
ffi::Error SwmRhsImpl(cudaStream_t stream, ffi::AnyBuffer hh, ffi::AnyBuffer u, ffi::AnyBuffer v,
                      ffi::Result<ffi::AnyBuffer> dh, ffi::Result<ffi::AnyBuffer> du, ffi::Result<ffi::AnyBuffer> dv,
                      int64_t apply_bc, double nu, double kappa, double tau0, double H0, double dx, double dy,
                      int64_t layered, int64_t bc, int64_t spec, F64s g_prime, F64s f_field, F64s wind_x,
                      F64s wind_y) {
  SB_SWM_SETUP;
  somax_b200_params p{nu, kappa, tau0, H0};
  if (somax_b200_swm_rhs(h, hh.untyped_data(), u.untyped_data(), v.untyped_data(), dh->untyped_data(),
                         du->untyped_data(), dv->untyped_data(), &p, static_cast<int>(apply_bc), stream))
    return fail("somax_b200_swm_rhs");
  return ffi::Error::Success();
}

// Scalars of diagnose() (swm/multilayer.py:225-256): out f64 [members, 3 nl + 1].
ffi::Error SwmDiagImpl(cudaStream_t stream, ffi::AnyBuffer hh, ffi::AnyBuffer u, ffi::AnyBuffer v,
                       ffi::Result<ffi::AnyBuffer> out, double dx, double dy, int64_t layered, int64_t bc,
                       int64_t spec, F64s g_prime, F64s f_field, F64s wind_x, F64s wind_y) {
  SB_SWM_SETUP;
  if (out->element_type() != ffi::DataType::F64) return bad("diagnostics buffer must be f64");
  if (somax_b200_swm_diag(h, hh.untyped_data(), u.untyped_data(), v.untyped_data(),
                          static_cast<double*>(out->untyped_data()), stream))
    return fail("somax_b200_swm_diag");
  return ffi::Error::Success();
}

}  // namespace

#define SB_STREAM .Ctx<ffi::PlatformStream<cudaStream_t>>()
#define SB_BUF_IN .Arg<ffi::AnyBuffer>()
#define SB_BUF_OUT .Ret<ffi::AnyBuffer>()
#define SB_PARAMS .Attr<double>("nu").Attr<double>("kappa").Attr<double>("tau0").Attr<double>("H0")
#define SB_QG_MODEL                                                                                  \
  .Attr<double>("dx").Attr<double>("dy").Attr<int64_t>("layered").Attr<int64_t>("spec")              \
  .Attr<F64s>("Cl2m").Attr<F64s>("Cm2l").Attr<F64s>("lambdas").Attr<F64s>("beta_y").Attr<F64s>("wind")
#define SB_SWM_MODEL                                                                                 \
  .Attr<double>("dx").Attr<double>("dy").Attr<int64_t>("layered").Attr<int64_t>("bc").Attr<int64_t>("spec") \
  .Attr<F64s>("g_prime").Attr<F64s>("f_field").Attr<F64s>("wind_x").Attr<F64s>("wind_y")

XLA_FFI_DEFINE_HANDLER_SYMBOL(SomaxB200QgSteps, QgStepsImpl,
    ffi::Ffi::Bind() SB_STREAM SB_BUF_IN SB_BUF_OUT
        .Attr<int64_t>("n_steps").Attr<double>("dt").Attr<double>("dt_last").Attr<int64_t>("resume")
        SB_PARAMS SB_QG_MODEL);
XLA_FFI_DEFINE_HANDLER_SYMBOL(SomaxB200QgRhs, QgRhsImpl,
    ffi::Ffi::Bind() SB_STREAM SB_BUF_IN SB_BUF_OUT .Attr<int64_t>("apply_bc") SB_PARAMS SB_QG_MODEL);
XLA_FFI_DEFINE_HANDLER_SYMBOL(SomaxB200QgInvert, QgInvertImpl,
    ffi::Ffi::Bind() SB_STREAM SB_BUF_IN SB_BUF_OUT SB_QG_MODEL);
XLA_FFI_DEFINE_HANDLER_SYMBOL(SomaxB200QgDiag, QgDiagImpl,
    ffi::Ffi::Bind() SB_STREAM SB_BUF_IN SB_BUF_OUT SB_QG_MODEL);
XLA_FFI_DEFINE_HANDLER_SYMBOL(SomaxB200SwmSteps, SwmStepsImpl,
    ffi::Ffi::Bind() SB_STREAM SB_BUF_IN SB_BUF_IN SB_BUF_IN SB_BUF_OUT SB_BUF_OUT SB_BUF_OUT
        .Attr<int64_t>("n_steps").Attr<double>("dt").Attr<double>("dt_last").Attr<int64_t>("resume")
        SB_PARAMS SB_SWM_MODEL);
XLA_FFI_DEFINE_HANDLER_SYMBOL(SomaxB200SwmRhs, SwmRhsImpl,
    ffi::Ffi::Bind() SB_STREAM SB_BUF_IN SB_BUF_IN SB_BUF_IN SB_BUF_OUT SB_BUF_OUT SB_BUF_OUT
        .Attr<int64_t>("apply_bc") SB_PARAMS SB_SWM_MODEL);
XLA_FFI_DEFINE_HANDLER_SYMBOL(SomaxB200SwmDiag, SwmDiagImpl,
    ffi::Ffi::Bind() SB_STREAM SB_BUF_IN SB_BUF_IN SB_BUF_IN SB_BUF_OUT SB_SWM_MODEL);
// The slab-distributed model (somax_b200_qgs_*) is driven from the host wrapper
// (somax_b200.parallel.SlabQG), one process per GPU: its set-up exchanges CUDA IPC handles between
// processes, which is outside what a single-process XLA custom call can express.
