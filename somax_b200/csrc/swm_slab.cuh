// Slab-distributed shallow-water model: ONE grid partitioned in y-slabs over several B200s
// (last bullet: halo pattern only - there is no elliptic solve).
//
// Included at the end of swm.cu (it drives the same fused right-hand-side kernels).  Rank r of P
// owns rows [r*ny/P, (r+1)*ny/P) of every layer as a window (ny/P + 2 rows) of the global padded
// arrays; its first / last row is the physical ghost row on the edge ranks and a halo row of the
// neighbour's data elsewhere.  The single-pass kernels need one halo row of the stage state
// (h, u, v); they are told through SwmArgs::ylo / yhi which of their first / last rows are physical,
// so the interior-only operators and the boundary conditions see the global picture.  Periodic
// basins: ranks 0 and P-1 are neighbours too - the y wrap-around of enforce_periodic IS a halo
// exchange - but their ghost rows stay physical ghost rows (stepped with the terms the reference
// leaves there, the state that comes back is the reference's un-projected one): the kernels read the
// boundary-conditioned content of those rows from the inbox (SwmArgs::lo_src / hi_src) instead of
// overwriting them.  One right-hand-side evaluation:
//
//   barrier    flag barrier in peer memory (the neighbours' pushes of the previous evaluation landed)
//   halo in    inbox -> ghost rows of the stage state about to be read
//   rhs        the fused single-GPU kernel (mass fluxes, PV, Bernoulli, Coriolis, diffusion + Tsit5
//              epilogue) on the window
//   halo out   first / last owned row of the new stage state -> the neighbours' inbox (peer-memory
//              stores over NVLink; two inbox sets alternate, so one barrier per evaluation suffices)
//
// Replaces, for this configuration, MultilayerShallowWater2D.{apply_boundary_conditions,
// vector_field} under SomaxModel.integrate (swm/multilayer.py:150-223, core/model.py:47-88).
#include "slab_common.cuh"

namespace sb {

constexpr int SWMS_NBUF = 2;   // exported buffers: inbox, flags

struct SwmSlabRank {
  somax_b200_swm_t core = nullptr;
  int rank = 0, lower = -1, upper = -1;   // neighbour ranks (-1: none; periodic basins wrap around)
  bool phys_lo = false, phys_hi = false;  // row 0 / Ny-1 of the window is the physical ghost row
  SegTable xbc0;                          // periodic: inbox -> physical ghost rows of y (the BC of state0)
  void* inbox = nullptr;                  // [2 sets][2: from below / from above][3 fields][planes][pitch]
  unsigned* flags = nullptr;
  SegTable xout[3][2], xin[3][2];         // [stage buffer y / Ya / Yb][inbox set]
};

}  // namespace sb

struct somax_b200_swms_s {
  int dtype = 0, nl = 0, ny = 0, nx = 0, bc = 0, nranks = 0, rank_first = 0, nlocal = 0, ny_loc = 0;
  std::vector<sb::SwmSlabRank> local;
  void* peers[sb::QGS_MAX_RANKS][sb::SLAB_MAX_BUF] = {};
  bool attached = false;
  std::vector<void*> ipc_opened;
  unsigned epoch = 0;
  size_t bytes = 0;
};

namespace {

inline size_t swms_es(const somax_b200_swms_s* g) { return g->dtype == SOMAX_B200_F32 ? 4 : 8; }

void swms_local_ptrs(const SwmSlabRank& R, void** out) { out[0] = R.inbox; out[1] = R.flags; }

int swms_build_tables(somax_b200_swms_s* g, SwmSlabRank& R) {
  const size_t es = swms_es(g);
  const Layout& L = R.core->L;
  const int planes = L.batch * L.nl, nyl = g->ny_loc;
  const size_t rowb = (size_t)L.pitch * es, planeb = L.plane() * es;
  const size_t fieldb = (size_t)planes * rowb;            // one field's rows in an inbox slot
  const size_t slotb = 3 * fieldb, setb = 2 * slotb;      // slot = 3 fields; set = from below + from above
  void** bufs[3] = {R.core->y, R.core->Ya, R.core->Yb};
  for (int b = 0; b < 3; ++b)
    for (int set = 0; set < 2; ++set) {
      std::vector<Seg> xo, xi;
      for (int f = 0; f < 3; ++f) {
        const char* mine = (const char*)bufs[b][f];
        if (R.lower >= 0)      // my first owned row -> the lower neighbour's "from above" slot
          xo.push_back(Seg{mine + 1 * rowb, (char*)g->peers[R.lower][0] + set * setb + slotb + f * fieldb,
                           (unsigned)planes, (unsigned)rowb, planeb, rowb});
        if (R.upper >= 0)      // my last owned row -> the upper neighbour's "from below" slot
          xo.push_back(Seg{mine + (size_t)nyl * rowb, (char*)g->peers[R.upper][0] + set * setb + f * fieldb,
                           (unsigned)planes, (unsigned)rowb, planeb, rowb});
        char* own = (char*)bufs[b][f];
        // halo rows only: a physical ghost row is never overwritten (periodic: read from the inbox)
        if (R.lower >= 0 && !R.phys_lo)
          xi.push_back(Seg{(const char*)R.inbox + set * setb + f * fieldb, own, (unsigned)planes, (unsigned)rowb, rowb, planeb});
        if (R.upper >= 0 && !R.phys_hi)
          xi.push_back(Seg{(const char*)R.inbox + set * setb + slotb + f * fieldb, own + (size_t)(L.Ny - 1) * rowb,
                           (unsigned)planes, (unsigned)rowb, rowb, planeb});
      }
      if (int rc = seg_upload(R.xout[b][set], xo, &g->bytes)) return rc;
      if (int rc = seg_upload(R.xin[b][set], xi, &g->bytes)) return rc;
    }
  // apply_boundary_conditions(state0) of a periodic basin: ghost row := the opposite rank's edge row
  // (inbox set 1, filled by the initial push)
  std::vector<Seg> x0;
  for (int f = 0; f < 3; ++f) {
    char* own = (char*)R.core->y[f];
    if (R.lower >= 0 && R.phys_lo)
      x0.push_back(Seg{(const char*)R.inbox + setb + f * fieldb, own, (unsigned)planes, (unsigned)rowb, rowb, planeb});
    if (R.upper >= 0 && R.phys_hi)
      x0.push_back(Seg{(const char*)R.inbox + setb + slotb + f * fieldb, own + (size_t)(L.Ny - 1) * rowb,
                       (unsigned)planes, (unsigned)rowb, rowb, planeb});
  }
  return seg_upload(R.xbc0, x0, &g->bytes);
}

int swms_barrier(somax_b200_swms_s* g, cudaStream_t s) {
  if (g->nlocal == g->nranks) return 0;      // one process, one stream: stream order is the barrier
  SwmSlabRank& R = g->local[0];
  FlagPtrs F;
  for (int r = 0; r < QGS_MAX_RANKS; ++r) F.p[r] = r < g->nranks ? (unsigned*)g->peers[r][1] : nullptr;
  ++g->epoch;
  static const unsigned long long timeout_ns = [] {
    const char* e = getenv("SOMAX_B200_SLAB_TIMEOUT_S");
    const double sec = e ? atof(e) : 20.0;
    return (unsigned long long)((sec > 0 ? sec : 20.0) * 1e9);
  }();
  prof_begin("slab_barrier", s);
  slab_barrier_kernel<<<1, 32, 0, s>>>(F, R.rank, g->nranks, g->epoch, R.flags + QGS_MAX_RANKS, timeout_ns);
  SB_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int swms_steps_impl(somax_b200_swms_s* g, void* const* hs, void* const* us, void* const* vs, long n_steps,
                    double dt, double dt_last, const somax_b200_params* p, cudaStream_t s) {
  const long total = n_steps + (dt_last > 0 ? 1 : 0);
  auto bufp = [](SwmSlabRank& R, int b) -> void** { return b == 0 ? R.core->y : (b == 1 ? R.core->Ya : R.core->Yb); };
  for (size_t v = 0; v < g->local.size(); ++v) {
    SwmSlabRank& R = g->local[v];
    void* ext[3] = {hs[v], us[v], vs[v]};
    for (int f = 0; f < 3; ++f)
      if (int rc = pack_field<T>((const T*)ext[f], (T*)R.core->y[f], R.core->L, s)) return rc;
  }
  // edge rows of state0 -> the neighbours' inbox (set 1): the halo rows of the first evaluation and,
  // in a periodic basin, the boundary condition of state0 on the physical ghost rows
  for (SwmSlabRank& R : g->local)
    if (int rc = seg_launch("slab_halo_state", R.xout[0][1], s)) return rc;
  if (int rc = swms_barrier(g, s)) return rc;
  for (SwmSlabRank& R : g->local) {
    if (int rc = seg_launch("slab_halo_in", R.xbc0, s)) return rc;      // rows first ...
    const int ylo = R.core->bc_ylo, yhi = R.core->bc_yhi;
    if (g->bc == SOMAX_B200_BC_PERIODIC && g->nranks > 1) { R.core->bc_ylo = 0; R.core->bc_yhi = 0; }
    const int rc = bc_inplace<T>(R.core, R.core->y, s);                  // ... then columns (enforce_periodic order)
    R.core->bc_ylo = ylo; R.core->bc_yhi = yhi;
    if (rc) return rc;
  }
  int y = 0, Yc = 1, Yn = 2;
  long ev = 0;        // evaluation counter: evaluation e reads inbox set (e + 1) & 1 and pushes into set e & 1
  if (total > 0) {
    auto eval = [&](int in_b, int y_b, int out_b, int e, double hd, bool store_f, int f_slot) -> int {
      if (int rc = swms_barrier(g, s)) return rc;
      for (SwmSlabRank& R : g->local)
        if (int rc = seg_launch("slab_halo_in", R.xin[in_b][(ev + 1) & 1], s)) return rc;
      for (SwmSlabRank& R : g->local) {
        SwmArgs<T> A = make_args<T>(R.core, p, 1);
        if (g->bc == SOMAX_B200_BC_PERIODIC) {
          const Layout& L = R.core->L;
          const size_t fieldb = (size_t)L.batch * L.nl * L.pitch * sizeof(T), slotb = 3 * fieldb, setb = 2 * slotb;
          const char* set = (const char*)R.inbox + ((ev + 1) & 1) * setb;
          for (int f = 0; f < 3; ++f) {
            if (R.phys_lo && R.lower >= 0) A.lo_src[f] = (const T*)(set + f * fieldb);
            if (R.phys_hi && R.upper >= 0) A.hi_src[f] = (const T*)(set + slotb + f * fieldb);
          }
        }
        Stage<T> st = empty_stage<T>();
        st.dt = (T)hd; st.a_new = (T)TSIT5_A[e][e]; st.nprev = e;
        for (int jj = 0; jj < e; ++jj) st.a[jj] = (T)TSIT5_A[e][jj];
        for (int f = 0; f < 3; ++f) {
          st.Yin[f] = (const T*)bufp(R, in_b)[f];
          st.Yout[f] = (T*)bufp(R, out_b)[f];
          if (e > 0) st.y[f] = (const T*)bufp(R, y_b)[f];
          for (int jj = 0; jj < e; ++jj) st.Fprev[jj][f] = (const T*)R.core->F[jj][f];
          st.Fout[f] = store_f ? (T*)R.core->F[f_slot][f] : nullptr;
        }
        if (int rc = launch_rhs<T>(R.core, A, st, s)) return rc;
      }
      for (SwmSlabRank& R : g->local)
        if (int rc = seg_launch("slab_halo_state", R.xout[out_b][ev & 1], s)) return rc;
      ++ev;
      return 0;
    };
    auto step_dt = [&](long i) { return (i < n_steps) ? dt : dt_last; };
    if (int rc = eval(y, y, Yc, 0, step_dt(0), true, 0)) return rc;
    auto stages = [&](double hd) -> int {
      for (int e = 1; e <= 5; ++e) {
        if (int rc = eval(Yc, y, Yn, e, hd, e <= 4, e)) return rc;
        std::swap(Yc, Yn);
      }
      return 0;
    };
    for (long i = 0; i + 1 < total; ++i) {
      if (int rc = stages(step_dt(i))) return rc;
      if (int rc = eval(Yc, Yc, Yn, 0, step_dt(i + 1), true, 0)) return rc;
      const int oy = y; y = Yc; Yc = Yn; Yn = oy;
    }
    if (int rc = stages(step_dt(total - 1))) return rc;
    std::swap(y, Yc);
    // halo rows of the final state: pushed after its kernel; land them before handing the slab back
    if (int rc = swms_barrier(g, s)) return rc;
    for (SwmSlabRank& R : g->local)
      if (int rc = seg_launch("slab_halo_in", R.xin[y][(ev + 1) & 1], s)) return rc;
    if (int rc = swms_barrier(g, s)) return rc;      // the inboxes may be primed again by the next call
  }
  for (size_t v = 0; v < g->local.size(); ++v) {
    SwmSlabRank& R = g->local[v];
    void* ext[3] = {hs[v], us[v], vs[v]};
    for (int f = 0; f < 3; ++f)
      if (int rc = unpack_field<T>((const T*)bufp(R, y)[f], (T*)ext[f], R.core->L, s)) return rc;
  }
  return 0;
}

}  // namespace

extern "C" {

int somax_b200_swms_create(somax_b200_swms_t* out, int dtype, int nl, int ny, int nx, double dx, double dy,
                           int bc, const double* g_prime, const double* f_field, const double* wind_x,
                           const double* wind_y, int nranks, int rank_first, int nlocal, unsigned spec_flags) {
  if (!out) return fail(SOMAX_B200_ERR_INVALID, "out is null");
  *out = nullptr;
  if (nranks < 1 || nranks > QGS_MAX_RANKS || rank_first < 0 || nlocal < 1 || rank_first + nlocal > nranks)
    return fail(SOMAX_B200_ERR_INVALID, "need 1 <= nranks <= 16 and local ranks inside [0, nranks)");
  if (nlocal != nranks && nlocal != 1)
    return fail(SOMAX_B200_ERR_UNSUPPORTED, "a process holds either one slab or all of them");
  if (ny % nranks != 0 || ny / nranks < 3)
    return fail(SOMAX_B200_ERR_UNSUPPORTED, "slab decomposition needs ny divisible by nranks (>= 3 rows each)");
  if (!g_prime || !f_field || !wind_x || !wind_y) return fail(SOMAX_B200_ERR_INVALID, "null coefficient pointer");
  if (int rc = require_device()) return rc;
  auto* g = new somax_b200_swms_s();
  g->dtype = dtype; g->nl = nl; g->ny = ny; g->nx = nx; g->bc = bc;
  g->nranks = nranks; g->rank_first = rank_first; g->nlocal = nlocal; g->ny_loc = ny / nranks;
  const int Nx = nx + 2, nyl = g->ny_loc;
  const size_t es = swms_es(g);
  const bool periodic = bc == SOMAX_B200_BC_PERIODIC && nranks > 1;
  int rc = 0;
  g->local.resize(nlocal);
  for (int v = 0; v < nlocal && !rc; ++v) {
    SwmSlabRank& R = g->local[v];
    R.rank = rank_first + v;
    R.lower = R.rank > 0 ? R.rank - 1 : (periodic ? nranks - 1 : -1);
    R.upper = R.rank < nranks - 1 ? R.rank + 1 : (periodic ? 0 : -1);
    R.phys_lo = R.rank == 0; R.phys_hi = R.rank == nranks - 1;
    const size_t row0 = (size_t)R.rank * nyl;      // the window starts at global row row0
    rc = somax_b200_swm_create(&R.core, dtype, 1, nl, nyl, nx, dx, dy, bc, g_prime, f_field + row0 * Nx,
                               wind_x + row0 * Nx, wind_y + row0 * Nx, spec_flags);
    if (rc) break;
    R.core->bc_ylo = R.phys_lo; R.core->bc_yhi = R.phys_hi;
    const size_t ib = 2 * 2 * 3 * (size_t)nl * R.core->L.pitch * es;
    if (cudaMalloc(&R.inbox, ib) != cudaSuccess || cudaMemset(R.inbox, 0, ib) != cudaSuccess ||
        cudaMalloc((void**)&R.flags, 2 * QGS_MAX_RANKS * sizeof(unsigned)) != cudaSuccess ||
        cudaMemset(R.flags, 0, 2 * QGS_MAX_RANKS * sizeof(unsigned)) != cudaSuccess) {
      rc = fail(SOMAX_B200_ERR_CUDA, "cudaMalloc (slab inbox / flags) failed");
      break;
    }
    g->bytes += somax_b200_swm_device_bytes(R.core) + ib;
  }
  if (!rc && nlocal == nranks) {
    for (int v = 0; v < nlocal; ++v) swms_local_ptrs(g->local[v], g->peers[v]);
    for (int v = 0; v < nlocal && !rc; ++v) rc = swms_build_tables(g, g->local[v]);
    g->attached = !rc;
  }
  if (!rc) rc = (cudaDeviceSynchronize() == cudaSuccess) ? 0 : fail(SOMAX_B200_ERR_CUDA, "slab create: device error");
  if (rc) { somax_b200_swms_destroy(g); return rc; }
  *out = g;
  return 0;
}

int somax_b200_swms_destroy(somax_b200_swms_t g) {
  if (!g) return 0;
  cudaDeviceSynchronize();
  for (void* p : g->ipc_opened) cudaIpcCloseMemHandle(p);
  for (SwmSlabRank& R : g->local) {
    for (int b = 0; b < 3; ++b)
      for (int set = 0; set < 2; ++set) { cudaFree(R.xout[b][set].dev); cudaFree(R.xin[b][set].dev); }
    cudaFree(R.xbc0.dev);
    cudaFree(R.inbox); cudaFree(R.flags);
    somax_b200_swm_destroy(R.core);
  }
  delete g;
  return 0;
}

size_t somax_b200_swms_device_bytes(somax_b200_swms_t g) { return g ? g->bytes : 0; }
size_t somax_b200_swms_export_bytes(void) { return sizeof(SlabIpcBlob); }

int somax_b200_swms_export(somax_b200_swms_t g, void* blob) {
  if (!g || !blob) return fail(SOMAX_B200_ERR_INVALID, "null argument");
  if (g->nlocal != 1) return fail(SOMAX_B200_ERR_INVALID, "export is for one-slab-per-process groups");
  void* ptrs[SWMS_NBUF];
  swms_local_ptrs(g->local[0], ptrs);
  return slab_ipc_export(ptrs, SWMS_NBUF, blob);
}

int somax_b200_swms_attach(somax_b200_swms_t g, const void* blobs) {
  if (!g || !blobs) return fail(SOMAX_B200_ERR_INVALID, "null argument");
  if (g->nlocal != 1) return fail(SOMAX_B200_ERR_INVALID, "attach is for one-slab-per-process groups");
  if (g->attached) return fail(SOMAX_B200_ERR_INVALID, "already attached");
  void* mine[SWMS_NBUF];
  swms_local_ptrs(g->local[0], mine);
  if (int rc = slab_ipc_attach(blobs, g->nranks, g->local[0].rank, SWMS_NBUF, mine, g->peers, g->ipc_opened)) return rc;
  if (int rc = swms_build_tables(g, g->local[0])) return rc;
  g->attached = true;
  return 0;
}

int somax_b200_swms_steps(somax_b200_swms_t g, void* const* h_slabs, void* const* u_slabs, void* const* v_slabs,
                          long n_steps, double dt, double dt_last, const somax_b200_params* p, void* stream) {
  if (!g || !h_slabs || !u_slabs || !v_slabs || !p) return fail(SOMAX_B200_ERR_INVALID, "null argument");
  if (!g->attached) return fail(SOMAX_B200_ERR_INVALID, "slab group not attached to its peers");
  if (n_steps < 0 || !(dt > 0) || dt_last < 0) return fail(SOMAX_B200_ERR_INVALID, "need n_steps>=0, dt>0, dt_last>=0");
  for (int v = 0; v < g->nlocal; ++v)
    if (!h_slabs[v] || !u_slabs[v] || !v_slabs[v]) return fail(SOMAX_B200_ERR_INVALID, "null slab pointer");
  return g->dtype == SOMAX_B200_F32
             ? swms_steps_impl<float>(g, h_slabs, u_slabs, v_slabs, n_steps, dt, dt_last, p, (cudaStream_t)stream)
             : swms_steps_impl<double>(g, h_slabs, u_slabs, v_slabs, n_steps, dt, dt_last, p, (cudaStream_t)stream);
}

int somax_b200_swms_status(somax_b200_swms_t g, int* barrier_timeouts) {
  if (!g || !barrier_timeouts) return fail(SOMAX_B200_ERR_INVALID, "null argument");
  *barrier_timeouts = 0;
  for (SwmSlabRank& R : g->local) {
    unsigned e = 0;
    SB_CUDA(cudaMemcpy(&e, R.flags + QGS_MAX_RANKS, sizeof(e), cudaMemcpyDeviceToHost));
    *barrier_timeouts += (int)e;
  }
  return 0;
}

}  // extern "C"
