// jax.ffi custom-call shim over the C ABI of libsomax_b200.so.
//
// NOT compiled in this image: it needs the XLA FFI headers (`jax.ffi.include_dir()`), and no
// jax / jaxlib is installed here.  Build where JAX is available with
//   g++ -O2 -fPIC -shared -std=c++17 -I$(python -c "import jax; print(jax.ffi.include_dir())") \
//       -Iinclude somax_b200/csrc/jax_ffi_shim.cc -Lsomax_b200/lib -lsomax_b200 \
//       -o somax_b200/lib/libsomax_b200_jax.so
// and register the handlers as shown in INTEGRATION.md.  The shim only unpacks buffers /
// attributes / the CUDA stream and forwards to the C ABI; handles are cached per
// (shape, dtype, coefficient pointer identity) because XLA may call the handler from several host
// threads (one per device).
#include <cuda_runtime_api.h>

#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "somax_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

struct QgKey {
  int dtype, batch, nl, ny, nx;
  double dx, dy;
  uint64_t coef_hash;
  bool operator<(const QgKey& o) const {
    return std::tie(dtype, batch, nl, ny, nx, dx, dy, coef_hash) <
           std::tie(o.dtype, o.batch, o.nl, o.ny, o.nx, o.dx, o.dy, o.coef_hash);
  }
};

std::mutex g_mu;
std::map<QgKey, somax_b200_qg_t> g_qg;

uint64_t fnv(const void* p, size_t n, uint64_t h = 1469598103934665603ull) {
  const unsigned char* c = static_cast<const unsigned char*>(p);
  for (size_t i = 0; i < n; ++i) h = (h ^ c[i]) * 1099511628211ull;
  return h;
}

ffi::Error fail(const char* what) {
  return ffi::Error(ffi::ErrorCode::kInternal, std::string(what) + ": " + somax_b200_last_error());
}

// Setup arrays arrive as HOST-resident attributes (spans of f64), exactly what *_create takes.
somax_b200_qg_t get_qg(const QgKey& key, ffi::Span<const double> Cl2m, ffi::Span<const double> Cm2l,
                       ffi::Span<const double> lambdas, ffi::Span<const double> beta_y,
                       ffi::Span<const double> wind) {
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_qg.find(key);
  if (it != g_qg.end()) return it->second;
  somax_b200_qg_t h = nullptr;
  if (somax_b200_qg_create(&h, key.dtype, key.batch, key.nl, key.ny, key.nx, key.dx, key.dy,
                           Cl2m.begin(), Cm2l.begin(), lambdas.begin(), beta_y.begin(), wind.begin(),
                           SOMAX_B200_SOLVER_AUTO, SOMAX_B200_SPEC_ADVECTION_REGION2) != 0)
    return nullptr;
  g_qg[key] = h;
  return h;
}

// q: (batch?, nl, Ny, Nx) -> q after n_steps Tsit5 steps.  Replaces SomaxModel.integrate
// (core/model.py:53-88) for BaroclinicQG / BarotropicQG.
ffi::Error QgStepsImpl(cudaStream_t stream, ffi::AnyBuffer q, ffi::Result<ffi::AnyBuffer> out,
                       int64_t n_steps, double dt, double dt_last, double nu, double kappa,
                       double tau0, double H0, double dx, double dy,
                       ffi::Span<const double> Cl2m, ffi::Span<const double> Cm2l,
                       ffi::Span<const double> lambdas, ffi::Span<const double> beta_y,
                       ffi::Span<const double> wind) {
  auto dims = q.dimensions();
  const int nd = static_cast<int>(dims.size());
  if (nd < 3) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "q must be (.., nl, Ny, Nx)");
  QgKey key;
  key.dtype = q.element_type() == ffi::DataType::F32 ? SOMAX_B200_F32 : SOMAX_B200_F64;
  key.nx = static_cast<int>(dims[nd - 1]) - 2;
  key.ny = static_cast<int>(dims[nd - 2]) - 2;
  key.nl = static_cast<int>(dims[nd - 3]);
  key.batch = nd == 4 ? static_cast<int>(dims[0]) : 1;
  key.dx = dx; key.dy = dy;
  key.coef_hash = fnv(beta_y.begin(), beta_y.size() * 8, fnv(lambdas.begin(), lambdas.size() * 8));
  somax_b200_qg_t h = get_qg(key, Cl2m, Cm2l, lambdas, beta_y, wind);
  if (!h) return fail("somax_b200_qg_create");
  const size_t bytes = q.size_bytes();
  if (cudaMemcpyAsync(out->untyped_data(), q.untyped_data(), bytes, cudaMemcpyDeviceToDevice, stream) !=
      cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "cudaMemcpyAsync failed");
  somax_b200_params p{nu, kappa, tau0, H0};
  if (somax_b200_qg_steps(h, out->untyped_data(), n_steps, dt, dt_last, &p, stream) != 0)
    return fail("somax_b200_qg_steps");
  return ffi::Error::Success();
}

// dq = vector_field(apply_boundary_conditions(q)): the `_rhs` diffrax sees (core/model.py:47-51).
ffi::Error QgRhsImpl(cudaStream_t stream, ffi::AnyBuffer q, ffi::Result<ffi::AnyBuffer> dq,
                     int64_t apply_bc, double nu, double kappa, double tau0, double H0, double dx,
                     double dy, ffi::Span<const double> Cl2m, ffi::Span<const double> Cm2l,
                     ffi::Span<const double> lambdas, ffi::Span<const double> beta_y,
                     ffi::Span<const double> wind) {
  auto dims = q.dimensions();
  const int nd = static_cast<int>(dims.size());
  QgKey key;
  key.dtype = q.element_type() == ffi::DataType::F32 ? SOMAX_B200_F32 : SOMAX_B200_F64;
  key.nx = static_cast<int>(dims[nd - 1]) - 2;
  key.ny = static_cast<int>(dims[nd - 2]) - 2;
  key.nl = static_cast<int>(dims[nd - 3]);
  key.batch = nd == 4 ? static_cast<int>(dims[0]) : 1;
  key.dx = dx; key.dy = dy;
  key.coef_hash = fnv(beta_y.begin(), beta_y.size() * 8, fnv(lambdas.begin(), lambdas.size() * 8));
  somax_b200_qg_t h = get_qg(key, Cl2m, Cm2l, lambdas, beta_y, wind);
  if (!h) return fail("somax_b200_qg_create");
  somax_b200_params p{nu, kappa, tau0, H0};
  if (somax_b200_qg_rhs(h, q.untyped_data(), dq->untyped_data(), nullptr, &p, static_cast<int>(apply_bc),
                        stream) != 0)
    return fail("somax_b200_qg_rhs");
  return ffi::Error::Success();
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    SomaxB200QgSteps, QgStepsImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::AnyBuffer>().Ret<ffi::AnyBuffer>()
        .Attr<int64_t>("n_steps").Attr<double>("dt").Attr<double>("dt_last")
        .Attr<double>("nu").Attr<double>("kappa").Attr<double>("tau0").Attr<double>("H0")
        .Attr<double>("dx").Attr<double>("dy")
        .Attr<ffi::Span<const double>>("Cl2m").Attr<ffi::Span<const double>>("Cm2l")
        .Attr<ffi::Span<const double>>("lambdas").Attr<ffi::Span<const double>>("beta_y")
        .Attr<ffi::Span<const double>>("wind"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    SomaxB200QgRhs, QgRhsImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::AnyBuffer>().Ret<ffi::AnyBuffer>()
        .Attr<int64_t>("apply_bc")
        .Attr<double>("nu").Attr<double>("kappa").Attr<double>("tau0").Attr<double>("H0")
        .Attr<double>("dx").Attr<double>("dy")
        .Attr<ffi::Span<const double>>("Cl2m").Attr<ffi::Span<const double>>("Cm2l")
        .Attr<ffi::Span<const double>>("lambdas").Attr<ffi::Span<const double>>("beta_y")
        .Attr<ffi::Span<const double>>("wind"));
// The shallow-water handlers (SomaxB200SwmSteps / SomaxB200SwmRhs) follow the same pattern with
// three Arg / three Ret buffers and somax_b200_swm_{create,steps,rhs}.
