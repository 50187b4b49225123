/* somax_b200.h - C ABI of the B200-native somax time-stepping hot path.
 *
 * The reference (jejjohnson/somax v0.0.6) has no FFI layer: the interface this path sits
 * behind is the equinox `SomaxModel` contract (somax/_src/core/model.py:12-95).  This header
 * is the boundary a `jax.ffi` custom call (or any other host language) binds; every entry
 * point names the reference method it replaces.  See INTEGRATION.md for the binding stubs.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / jax types.
 *   - Field arrays are caller-owned DEVICE pointers in the reference layout: dense row-major
 *     (batch, nl, Ny, Nx), Ny = ny + 2, Nx = nx + 2 (one ghost ring), x contiguous, dtype
 *     float32 or float64 as fixed at create().  `batch` is the ensemble-member axis (1 for a
 *     single model; it is what `jax.vmap` over members would add).
 *   - Arrays passed to *_create are HOST pointers (double precision); they are setup data
 *     (somax `create()` factories, reference qg/baroclinic.py:277-332, swm/multilayer.py:313-377).
 *   - All compute calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL =
 *     legacy default stream), never allocate and never synchronise the host.  Allocation
 *     happens only in *_create / *_destroy.
 *   - Return value: 0 on success, negative somax_b200_status otherwise;
 *     somax_b200_last_error() returns a thread-local message.
 *   - Handles are not thread-safe per handle; distinct handles may be used concurrently.
 */
#ifndef SOMAX_B200_H
#define SOMAX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOMAX_B200_ABI_VERSION 1

typedef enum {
  SOMAX_B200_OK = 0,
  SOMAX_B200_ERR_INVALID = -1,     /* bad argument (shape, dtype, null pointer) */
  SOMAX_B200_ERR_UNSUPPORTED = -2, /* valid request the CUDA path does not implement */
  SOMAX_B200_ERR_CUDA = -3,        /* CUDA runtime error (message has the cudaError string) */
  SOMAX_B200_ERR_NO_DEVICE = -4,   /* no sm_100 device: there is NO CPU fallback */
  SOMAX_B200_ERR_COMM = -5         /* peer mapping (CUDA IPC) / slab-group error */
} somax_b200_status;

typedef enum { SOMAX_B200_F32 = 0, SOMAX_B200_F64 = 1 } somax_b200_dtype;

/* Boundary condition of the shallow-water model (swm/multilayer.py:203-223). */
typedef enum { SOMAX_B200_BC_PERIODIC = 0, SOMAX_B200_BC_WALL = 1 } somax_b200_swm_bc;

/* Elliptic-solver selection for the PV inversion (qg/baroclinic.py:146-152, bc="dst"). */
typedef enum {
  SOMAX_B200_SOLVER_AUTO = 0,   /* FFT path when nx is a power of two >= 8, else dense */
  SOMAX_B200_SOLVER_FFT = 1,    /* radix FFT DST-I(nx-1) in x + three border columns (Schur) + Thomas in y */
  SOMAX_B200_SOLVER_DENSE = 2   /* dense DST-I(nx+2) matrix in x + Thomas in y (any nx <= 2046) */
} somax_b200_solver;

/* finitevolx / spectraldiffx conventions; same bits as
 * oracle.OperatorSpec.flags().  The reference's behaviour, pinned by the outputs it printed in its
 * executed tutorials (tests/golden/reference_notebook_outputs.json), is
 * ADVECTION_REGION2 | DIFFUSION_FLUX with the DST bits clear. */
#define SOMAX_B200_SPEC_ADVECTION_REGION2 1u /* Advection2D writes [2:-2,2:-2] */
#define SOMAX_B200_SPEC_DIFFUSION_FLUX 2u    /* Diffusion2D in flux form with zero ghost fluxes */
#define SOMAX_B200_SPEC_DST_CONTINUOUS 4u    /* continuous DST eigenvalues: NOT the reference, ERR_UNSUPPORTED */
#define SOMAX_B200_SPEC_DST_INTERIOR 8u      /* DST on the interior with a zero ring: NOT the reference, ERR_UNSUPPORTED */
/* BarotropicQG._invert_pv keeps the ring of psi the solver returns (qg/barotropic.py:113-121);
 * without this bit the ring of psi is zero, as BaroclinicQG._invert_pv leaves it (qg/baroclinic.py:157-158). */
#define SOMAX_B200_SPEC_KEEP_PSI_RING 16u

typedef struct somax_b200_qg_s* somax_b200_qg_t;
typedef struct somax_b200_swm_s* somax_b200_swm_t;

/* Differentiable scalars of the reference `Params` pytrees plus the top-layer thickness the
 * wind term divides by (BaroclinicQGParams qg/baroclinic.py:37-49 and `strat.H[0]` :181;
 * MultilayerSW2DParams swm/multilayer.py:43-55 and :190-191).  Barotropic / single-layer
 * models pass H0 = 1 (qg/barotropic.py:142, swm/nonlinear_2d.py:167-168). */
typedef struct {
  double lateral_viscosity; /* nu    */
  double bottom_drag;       /* kappa */
  double wind_amplitude;    /* tau0  */
  double H0;
} somax_b200_params;

const char* somax_b200_last_error(void);
int somax_b200_abi_version(void);
/* Number of kernels this library has launched in the calling process (bench `gpu_launches`). */
uint64_t somax_b200_launch_count(void);

/* Per-kernel device timing for bench.py's roofline: when enabled every kernel launch of this
 * library is bracketed by CUDA events on its stream; report() synchronises and writes a JSON
 * array [{"kernel", "launches", "total_ms"}] into buf. */
void somax_b200_profile_enable(int on);
void somax_b200_profile_reset(void);
int somax_b200_profile_report(char* buf, size_t cap);

/* ------------------------------------------------------------------------------------------
 * Quasi-geostrophic model (barotropic = nl 1 with Cl2m = Cm2l = [[1]], lambda = [0], H0 = 1).
 * ------------------------------------------------------------------------------------------ */

/* Replaces BaroclinicQG.create / BarotropicQG.create as far as device state is concerned
 * (qg/baroclinic.py:277-332, qg/barotropic.py:216-248).  Cl2m, Cm2l (nl*nl row-major),
 * lambdas (nl) = `helmholtz_lambdas` are taken AS GIVEN from the model object, never
 * recomputed.  beta_y and wind are (Ny, Nx). */
int somax_b200_qg_create(somax_b200_qg_t* out, int dtype, int batch, int nl, int ny, int nx,
                         double dx, double dy, const double* Cl2m, const double* Cm2l,
                         const double* lambdas, const double* beta_y, const double* wind,
                         int solver, unsigned spec_flags);
int somax_b200_qg_destroy(somax_b200_qg_t h);
/* Bytes of device memory owned by the handle. */
size_t somax_b200_qg_device_bytes(somax_b200_qg_t h);

/* BaroclinicQG.apply_boundary_conditions (qg/baroclinic.py:192-195): ring := 0.  out may alias q. */
int somax_b200_qg_apply_bc(somax_b200_qg_t h, const void* q, void* out, void* stream);

/* BaroclinicQG._invert_pv (qg/baroclinic.py:135-159): psi = ring0(Cm2l . Helm^-1 . Cl2m . q); the
 * Helmholtz solve takes every point of the (Ny, Nx) array, ring included, as an unknown, exactly as
 * finitevolx.pv_inversion(bc="dst") does.  With SPEC_KEEP_PSI_RING: BarotropicQG._invert_pv. */
int somax_b200_qg_invert(somax_b200_qg_t h, const void* q, void* psi, void* stream);

/* BaroclinicQG.vector_field (qg/baroclinic.py:161-190).  apply_bc != 0 evaluates
 * vector_field(apply_boundary_conditions(q)), i.e. `_rhs` of SomaxModel.build_terms
 * (core/model.py:47-51).  psi_out may be NULL. */
int somax_b200_qg_rhs(somax_b200_qg_t h, const void* q, void* dq, void* psi_out,
                      const somax_b200_params* p, int apply_bc, void* stream);

/* SomaxModel.integrate with the default Tsit5 / ConstantStepSize / SaveAt(t1=True)
 * (core/model.py:53-88): BC on the initial state, then n_steps steps of dt and, if
 * dt_last > 0, one clipped step of dt_last.  q is updated in place (ghost ring NOT
 * re-projected, as in the reference). */
int somax_b200_qg_steps(somax_b200_qg_t h, void* q, long n_steps, double dt, double dt_last,
                        const somax_b200_params* p, void* stream);
/* The same integration CONTINUED from a state an earlier *_steps / *_resume call returned: no
 * boundary conditions on the initial state.  diffrax projects state0 once (core/model.py:62) and
 * the states it saves at intermediate `SaveAt(ts=...)` times are un-projected (core/model.py:75-88);
 * stepping from one save time to the next is one resume call. */
int somax_b200_qg_resume(somax_b200_qg_t h, void* q, long n_steps, double dt, double dt_last,
                         const somax_b200_params* p, void* stream);

/* Scalars of BaroclinicQG.diagnose (qg/baroclinic.py:197-228) plus the non-finite guard of the
 * runner (cli/_run.py:671-683).  out (DEVICE, double): batch * (2*nl + 1) values per member:
 * KE[nl], enstrophy[nl], count of non-finite q cells. */
int somax_b200_qg_diag(somax_b200_qg_t h, const void* q, double* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Slab-distributed QG: ONE grid partitioned in y-slabs over `nranks` GPUs (one process per GPU).
 * Same model as somax_b200_qg_* with batch = 1 and the FFT solver; replaces the same reference
 * functions (qg/baroclinic.py:135-195 under core/model.py:47-88).  Rank r owns rows
 * [r*ny/nranks, (r+1)*ny/nranks) and works on the window (nl, ny/nranks + 2, Nx) of the global
 * (nl, Ny, Nx) array that starts at global row r*ny/nranks (its first / last row is the physical
 * ring on the edge ranks, the neighbour's row elsewhere).  Exchanges (distributed-DST transposes,
 * border partials, halo rows) are device kernels that store into the peers' memory, mapped with
 * CUDA IPC over NVLink; ranks are ordered by flag barriers in peer memory.  Needs nx = 2^p,
 * (nx / 64) % nranks == 0, ny % nranks == 0.
 *
 * nlocal == 1: this process holds slab `rank_first`; call export() on every rank, all-gather
 * the blobs by any means (they are plain bytes), then attach().  nlocal == nranks: every slab
 * lives in this process on the current device (validation of the decomposition on one GPU).
 * ------------------------------------------------------------------------------------------ */
typedef struct somax_b200_qgs_s* somax_b200_qgs_t;

/* beta_y, wind: the GLOBAL (Ny, Nx) host arrays (as for somax_b200_qg_create). */
int somax_b200_qgs_create(somax_b200_qgs_t* out, int dtype, int nl, int ny, int nx, double dx,
                          double dy, const double* Cl2m, const double* Cm2l, const double* lambdas,
                          const double* beta_y, const double* wind, int nranks, int rank_first,
                          int nlocal, unsigned spec_flags);
int somax_b200_qgs_destroy(somax_b200_qgs_t g);
size_t somax_b200_qgs_device_bytes(somax_b200_qgs_t g);
/* Size of one rank's export blob (CUDA IPC handles of the buffers its peers store into). */
size_t somax_b200_qgs_export_bytes(void);
int somax_b200_qgs_export(somax_b200_qgs_t g, void* blob);
/* blobs: nranks * export_bytes, in rank order (this rank's own entry is ignored). */
int somax_b200_qgs_attach(somax_b200_qgs_t g, const void* blobs);
/* SomaxModel.integrate on the slabs, in place: q_slabs[v] (DEVICE) is local slab v's window
 * (nl, ny/nranks + 2, Nx), halo rows valid on entry and on return.  Collective: every rank of
 * the group must make the same call. */
int somax_b200_qgs_steps(somax_b200_qgs_t g, void* const* q_slabs, long n_steps, double dt,
                         double dt_last, const somax_b200_params* p, void* stream);
/* After synchronising the stream: number of barriers that gave up waiting for a peer (0 = ok). */
int somax_b200_qgs_status(somax_b200_qgs_t g, int* barrier_timeouts);

/* ------------------------------------------------------------------------------------------
 * Shallow-water model (NonlinearShallowWater2D = nl 1 with g_prime = [g], H0 = 1).
 * ------------------------------------------------------------------------------------------ */

/* MultilayerShallowWater2D.create (swm/multilayer.py:313-377).  g_prime (nl); f_field, wind_x,
 * wind_y are (Ny, Nx) at T points. */
int somax_b200_swm_create(somax_b200_swm_t* out, int dtype, int batch, int nl, int ny, int nx,
                          double dx, double dy, int bc, const double* g_prime,
                          const double* f_field, const double* wind_x, const double* wind_y,
                          unsigned spec_flags);
int somax_b200_swm_destroy(somax_b200_swm_t h);
/* Turns the handle into the reparameterized QG model (ReparameterizedQG, qg/reparameterized.py:66-189:
 * a MultilayerShallowWater2D whose apply_boundary_conditions also projects the state onto the
 * geostrophic manifold, P = G (Q G)^-1 Q).  From then on every entry point that applies the boundary
 * conditions (apply_bc, rhs with apply_bc != 0, steps) applies shallow-water BCs + projection; rhs
 * with apply_bc == 0 stays the plain shallow-water vector field (reparameterized.py:179-183).
 * H (nl) layer thicknesses; Cl2m, Cm2l (nl*nl), eigenvalues (nl) = ModalTransform fields;
 * lambdas (nl) = helmholtz_lambdas = f0^2 * eigenvalues as the model holds them.  Needs BC_WALL. */
int somax_b200_swm_set_projection(somax_b200_swm_t h, double f0, const double* H, const double* Cl2m,
                                  const double* Cm2l, const double* eigenvalues, const double* lambdas,
                                  int solver);
size_t somax_b200_swm_device_bytes(somax_b200_swm_t h);

/* MultilayerShallowWater2D.apply_boundary_conditions (swm/multilayer.py:203-223, 383-410). */
int somax_b200_swm_apply_bc(somax_b200_swm_t h, const void* hh, const void* u, const void* v,
                            void* hh_out, void* u_out, void* v_out, void* stream);

/* ReparameterizedQG.project (qg/reparameterized.py:142-177): (h, u, v) -> the geostrophically balanced
 * state, no boundary conditions applied first.  Needs somax_b200_swm_set_projection. */
int somax_b200_swm_project(somax_b200_swm_t h, const void* hh, const void* u, const void* v,
                           void* hh_out, void* u_out, void* v_out, void* stream);

/* MultilayerShallowWater2D.vector_field (swm/multilayer.py:150-201). */
int somax_b200_swm_rhs(somax_b200_swm_t h, const void* hh, const void* u, const void* v,
                       void* dh, void* du, void* dv, const somax_b200_params* p, int apply_bc,
                       void* stream);

/* SomaxModel.integrate (core/model.py:53-88) for the shallow-water state, in place. */
int somax_b200_swm_steps(somax_b200_swm_t h, void* hh, void* u, void* v, long n_steps, double dt,
                         double dt_last, const somax_b200_params* p, void* stream);
/* As somax_b200_qg_resume: continue without boundary conditions on the initial state. */
int somax_b200_swm_resume(somax_b200_swm_t h, void* hh, void* u, void* v, long n_steps, double dt,
                          double dt_last, const somax_b200_params* p, void* stream);

/* Scalars of MultilayerShallowWater2D.diagnose (swm/multilayer.py:225-256): out (DEVICE,
 * double) batch * (3*nl + 1): ke_sum[nl], sum(h^2)[nl] (PE = 0.5 g'_k * this), potential
 * enstrophy[nl] (all already times dx*dy), count of non-finite cells. */
int somax_b200_swm_diag(somax_b200_swm_t h, const void* hh, const void* u, const void* v,
                        double* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Slab-distributed shallow water: ONE grid partitioned in y-slabs over `nranks` GPUs, halo exchange
 * only (no elliptic solve).  Same model as somax_b200_swm_* with batch = 1; replaces the same
 * reference functions (swm/multilayer.py:150-223 under core/model.py:47-88).  Rank r owns rows
 * [r*ny/nranks, (r+1)*ny/nranks) and works on the window (nl, ny/nranks + 2, Nx) of the global
 * arrays starting at global row r*ny/nranks.  Periodic basins: ranks 0 and nranks-1 are neighbours
 * (the y wrap-around of enforce_periodic is the halo exchange).  Needs ny % nranks == 0.
 * nlocal as for somax_b200_qgs_create.  f_field, wind_x, wind_y: the GLOBAL (Ny, Nx) host arrays.
 * ------------------------------------------------------------------------------------------ */
typedef struct somax_b200_swms_s* somax_b200_swms_t;

int somax_b200_swms_create(somax_b200_swms_t* out, int dtype, int nl, int ny, int nx, double dx, double dy,
                           int bc, const double* g_prime, const double* f_field, const double* wind_x,
                           const double* wind_y, int nranks, int rank_first, int nlocal, unsigned spec_flags);
int somax_b200_swms_destroy(somax_b200_swms_t g);
size_t somax_b200_swms_device_bytes(somax_b200_swms_t g);
size_t somax_b200_swms_export_bytes(void);
int somax_b200_swms_export(somax_b200_swms_t g, void* blob);
int somax_b200_swms_attach(somax_b200_swms_t g, const void* blobs);
/* SomaxModel.integrate on the slabs, in place: h/u/v_slabs[v] (DEVICE) are local slab v's windows
 * (nl, ny/nranks + 2, Nx), halo rows valid on entry and on return.  Collective over the group. */
int somax_b200_swms_steps(somax_b200_swms_t g, void* const* h_slabs, void* const* u_slabs, void* const* v_slabs,
                          long n_steps, double dt, double dt_last, const somax_b200_params* p, void* stream);
int somax_b200_swms_status(somax_b200_swms_t g, int* barrier_timeouts);

#ifdef __cplusplus
}
#endif
#endif /* SOMAX_B200_H */
