// Slab-distributed quasi-geostrophic model: ONE grid partitioned in y-slabs over several B200s
// (BASELINE config "3-layer QG double-gyre 8192^2 slab-decomposed over 2/4/8 B200").
//
// Included at the end of qg.cu (it drives the same stencil / solver stages).  Rank r of P owns
// rows [r*ny/P, (r+1)*ny/P) of every layer, stored as a window (ny/P + 2 rows) of the global
// padded array: its first / last row is the physical ring on the edge ranks and a halo row of
// the neighbour's data elsewhere.  The PV inversion takes EVERY row of the global array as an
// unknown (oracle/elliptic.py), so for the solve the edge ranks also own their ring row: rank r
// holds the solver rows [slab_r0(r), slab_r1(r)) of the Ny = ny + 2.  One right-hand-side evaluation:
//
//   rows_fwd   layer->mode mix + DST-I in x of the slab's rows; the transform kernel stores every
//   + X1       spectral row segment STRAIGHT into the column array S of the rank that owns its strip
//              of x-wavenumbers (peer memory over NVLink): compute and transpose are one kernel
//   cols(1)    first Thomas solve in y on the rank's x-wavenumber strips
//   X2         border row-sum partials -> every rank                       (peer stores)
//   border     fixed-order reduction of all partials (replicated, cheap), then the Schur solve
//              (two dense fp64 DSTs in y) with the OUTPUT rows split over the ranks: slices of
//              ghat, then of gvec / the border column, are pushed to the peers (XG1, XG2);
//              bitwise the single-GPU result
//   cols(2)    second Thomas solve on the rank's strips; its last sweep stores every finished tile
//   + X3       (cp.async.bulk, shared -> peer memory) STRAIGHT into the row array R of the rank that
//              owns the rows: again no transpose kernel
//   rows_inv   inverse transform + mode->layer mix -> psi slab
//   X4         psi halo rows -> neighbours                                 (peer stores)
//   stencil    fused Arakawa / viscosity / wind / drag + Tsit5 epilogue on the slab
//   X5         new stage state halo rows -> neighbours' inbox              (peer stores)
//
// All exchanges are device kernels storing straight into the peer's memory (CUDA IPC mappings
// over NVLink / NVSwitch); ordering between ranks is a flag barrier in peer memory (one small
// kernel, release/acquire at system scope), six per evaluation.  There is no host
// synchronisation and no staging buffer.  With every slab in one process (nlocal == nranks, used
// to validate the decomposition on one device) the same kernels run on one stream and the
// barriers are not needed.
//
// Replaces, for this configuration, the same reference functions as qg.cu
// (qg/baroclinic.py:135-195, core/model.py:47-88).

#include "slab_common.cuh"

namespace sb {

constexpr int QGS_NBUF = 10;   // exported buffers: cols.S, cols.part, R, psi, inbox, flags, ghat, gvec, gvecf, cols.bext

struct SlabRank {
  somax_b200_qg_t core = nullptr;   // stencil buffers + row-transform solver of the slab (ny_loc rows)
  QgSolver* cols = nullptr;         // column solver of the whole grid (ny rows); only strips [s0, s1) are used
  int rank = 0, s0 = 0, s1 = 0;
  void* inbox = nullptr;            // [2][planes][pitch]: halo rows of the newest stage state from below / above
  unsigned* flags = nullptr;        // [QGS_MAX_RANKS] barrier slots + [QGS_MAX_RANKS] error word
  SegTable xb, x2, xpsi, xstate[3], xin[3], xg1, xg2, xg2f, xring;
};

struct SlabPeer { void* buf[QGS_NBUF]; };

}  // namespace sb

struct somax_b200_qgs_s {
  int dtype = 0, nl = 0, ny = 0, nx = 0, nranks = 0, rank_first = 0, nlocal = 0, ny_loc = 0, spr = 0, nseg = 1;
  double dx = 0, dy = 0;
  std::vector<sb::SlabRank> local;
  sb::SlabPeer peers[sb::QGS_MAX_RANKS];
  bool attached = false;
  std::vector<void*> ipc_opened;
  unsigned epoch = 0;
  size_t bytes = 0;
};

namespace {

struct IpcBlob {
  cudaIpcMemHandle_t handle[QGS_NBUF];
  unsigned long long offset[QGS_NBUF];
};

inline size_t qgs_es(const somax_b200_qgs_s* g) { return g->dtype == SOMAX_B200_F32 ? 4 : 8; }
// solver rows (= global field rows) owned by rank r: the edge ranks include their ring row
inline int slab_r0(const somax_b200_qgs_s* g, int r) { return r == 0 ? 0 : r * g->ny_loc + 1; }
inline int slab_r1(const somax_b200_qgs_s* g, int r) { return r == g->nranks - 1 ? g->ny + 2 : (r + 1) * g->ny_loc + 1; }

// Exchange tables of local rank R against the peer pointer table (all pointers valid in this process).
int qgs_build_tables(somax_b200_qgs_s* g, SlabRank& R) {
  const size_t es = qgs_es(g);
  const int P = g->nranks, p = R.rank, nyl = g->ny_loc, spr = g->spr;
  const QgSolverView vr = qg_solver_view(R.core->solver), vc = qg_solver_view(R.cols);
  const int planes = vc.planes, np = vc.np, nstrip = vc.nstrip;
  const int NyS = vc.ny;                                 // solver rows of the whole grid (ny + 2)
  const int my0 = slab_r0(g, p), my1 = slab_r1(g, p), myrows = my1 - my0;      // == vr.ny
  const Layout& L = R.core->L;
  char* Sl = (char*)vc.S;
  std::vector<Seg> xb, x2, xpsi, xg1, xg2, xg2f, xring;
  // BarotropicQG keeps the ring of psi: the solved border columns 0 and nx+1 of my rows (my slice of
  // the border-system outputs) go from the column solver's bext to the row solver's (local copy)
  for (int pl = 0; pl < planes; ++pl)
    for (int w = 0; w < 2; ++w)
      xring.push_back(Seg{(const char*)vc.bext + (((size_t)pl * 3 + w) * NyS + my0) * es,
                          (char*)vr.bext + ((size_t)pl * 3 + w) * myrows * es, 1u, (unsigned)(myrows * es), 0, 0});
  // fused exchanges: rows_fwd scatters into the peers' column arrays, cols(2) pushes into their row arrays
  {
    void* peerS[QGS_MAX_RANKS]; void* peerR[QGS_MAX_RANKS]; int row0[QGS_MAX_RANKS + 1];
    for (int r = 0; r < P; ++r) { peerS[r] = g->peers[r].buf[0]; peerR[r] = g->peers[r].buf[2]; row0[r] = slab_r0(g, r); }
    row0[P] = NyS;
    qg_solver_set_scatter(R.core->solver, peerS, P, spr, my0, NyS);
    qg_solver_set_push(R.cols, peerR, P, row0);
  }
  const int last_owner = (nstrip - 1) / spr;
  const int a0 = my0, na = myrows;                       // my slice of the border-system outputs
  for (int r = 0; r < P; ++r) {
    char* Sr = (char*)g->peers[r].buf[0];
    char* partr = (char*)g->peers[r].buf[1];
    char* bextr = (char*)g->peers[r].buf[9];
    for (int pl = 0; pl < planes; ++pl) {
      // raw border columns 0, nx+1 and nx of my rows (bext of the row stage) -> every rank (each one
      // reduces the border system)
      for (int w = 0; w < 3; ++w)
        xb.push_back(Seg{(const char*)vr.bext + ((size_t)pl * 3 + w) * myrows * es,
                         bextr + (((size_t)pl * 3 + w) * NyS + my0) * es, 1u, (unsigned)(myrows * es), 0, 0});
      // X2: the pre-summed border partials of my strips (my slot of pvec: planes x 2 vectors) -> every
      // other rank
      if (r != p && pl == 0) {
        const size_t o = (size_t)p * planes * 2 * NyS * sizeof(double);
        x2.push_back(Seg{(const char*)vc.pvec + o, partr + o, 1u, (unsigned)((size_t)planes * 2 * NyS * sizeof(double)), 0, 0});
      }
      // XG1 / XG2: my slice of ghat (3 columns), then of gvec (2 combinations; + float copy, + border
      // column nx of S for the rank that owns the last strip) -> every other rank
      if (r != p) {
        if (pl == 0) {      // ghat is interleaved [row][3 nl (padded)]: my rows are one contiguous block
          const size_t nvp = (size_t)((3 * planes + 1) & ~1);
          const size_t go = (size_t)a0 * nvp * sizeof(double);
          xg1.push_back(Seg{(const char*)vc.ghat + go, (char*)g->peers[r].buf[6] + go, 1u, (unsigned)(na * nvp * sizeof(double)), 0, 0});
        }
        for (int i = 0; i < 2; ++i) {
          const size_t go = (((size_t)pl * 2 + i) * NyS + a0) * sizeof(double);
          xg2.push_back(Seg{(const char*)vc.gvec + go, (char*)g->peers[r].buf[7] + go, 1u, (unsigned)(na * sizeof(double)), 0, 0});
          if (es == 4) {
            const size_t fo = (((size_t)pl * 2 + i) * NyS + a0) * sizeof(float);
            xg2f.push_back(Seg{(const char*)vc.gvecf + fo, (char*)g->peers[r].buf[8] + fo, 1u, (unsigned)(na * sizeof(float)), 0, 0});
          }
        }
        if (r == last_owner) {
          const size_t so = ((size_t)pl * NyS * np + sp_off(NyS, a0, vc.nx - 1)) * es;
          xg2f.push_back(Seg{Sl + so, Sr + so, (unsigned)na, (unsigned)es, SP_W * es, SP_W * es});
        }
      }
    }
  }
  // halo rows: my first owned row (1) -> rank p-1's top halo row; my last owned row -> rank p+1's row 0
  const size_t rowb = (size_t)L.pitch * es, planeb = L.plane() * es;
  const size_t inbox_slot = (size_t)planes * rowb;
  auto halo = [&](const char* mine, int buf, bool to_inbox, std::vector<Seg>& out) {
    if (p > 0) {
      char* d = (char*)g->peers[p - 1].buf[buf];
      out.push_back(to_inbox ? Seg{mine + 1 * rowb, d + inbox_slot, (unsigned)planes, (unsigned)rowb, planeb, rowb}
                             : Seg{mine + 1 * rowb, d + (size_t)(L.Ny - 1) * rowb, (unsigned)planes, (unsigned)rowb, planeb, planeb});
    }
    if (p < P - 1) {
      char* d = (char*)g->peers[p + 1].buf[buf];
      out.push_back(to_inbox ? Seg{mine + (size_t)nyl * rowb, d, (unsigned)planes, (unsigned)rowb, planeb, rowb}
                             : Seg{mine + (size_t)nyl * rowb, d, (unsigned)planes, (unsigned)rowb, planeb, planeb});
    }
  };
  halo((const char*)R.core->psi, 3, false, xpsi);
  void* st[3] = {R.core->y, R.core->Ya, R.core->Yb};
  if (int rc = seg_upload(R.xb, xb, &g->bytes)) return rc;
  if (int rc = seg_upload(R.x2, x2, &g->bytes)) return rc;
  if (int rc = seg_upload(R.xpsi, xpsi, &g->bytes)) return rc;
  if (int rc = seg_upload(R.xg1, xg1, &g->bytes)) return rc;
  if (int rc = seg_upload(R.xg2, xg2, &g->bytes)) return rc;
  if (int rc = seg_upload(R.xg2f, xg2f, &g->bytes)) return rc;
  if (int rc = seg_upload(R.xring, xring, &g->bytes)) return rc;
  for (int b = 0; b < 3; ++b) {
    std::vector<Seg> xs, xi;
    halo((const char*)st[b], 4, true, xs);
    // inbox -> ghost rows of my own buffer b (slot 0: from below -> row 0; slot 1: from above -> row Ny-1)
    if (p > 0) xi.push_back(Seg{(const char*)R.inbox, (char*)st[b], (unsigned)planes, (unsigned)rowb, rowb, planeb});
    if (p < P - 1)
      xi.push_back(Seg{(const char*)R.inbox + inbox_slot, (char*)st[b] + (size_t)(L.Ny - 1) * rowb,
                       (unsigned)planes, (unsigned)rowb, rowb, planeb});
    if (int rc = seg_upload(R.xstate[b], xs, &g->bytes)) return rc;
    if (int rc = seg_upload(R.xin[b], xi, &g->bytes)) return rc;
  }
  return 0;
}

void qgs_local_ptrs(const SlabRank& R, void** out) {
  const QgSolverView vr = qg_solver_view(R.core->solver), vc = qg_solver_view(R.cols);
  out[0] = vc.S; out[1] = vc.pvec; out[2] = vr.S; out[3] = R.core->psi; out[4] = R.inbox; out[5] = R.flags;
  out[6] = vc.ghat; out[7] = vc.gvec; out[8] = vc.gvecf ? (void*)vc.gvecf : (void*)vc.gvec; out[9] = vc.bext;
}

int qgs_barrier(somax_b200_qgs_s* g, cudaStream_t s) {
  if (g->nlocal == g->nranks) return 0;      // one process, one stream: stream order is the barrier
  SlabRank& R = g->local[0];
  FlagPtrs F;
  for (int r = 0; r < QGS_MAX_RANKS; ++r) F.p[r] = r < g->nranks ? (unsigned*)g->peers[r].buf[5] : nullptr;
  ++g->epoch;
  prof_begin("slab_barrier", s);
  // a peer that does not show up within the watchdog time (default 20 s, SOMAX_B200_SLAB_TIMEOUT_S)
  // sets the error word: the data of this call is then invalid and somax_b200_qgs_status reports it
  // (the host wrapper checks it after every call)
  static const unsigned long long timeout_ns = [] {
    const char* e = getenv("SOMAX_B200_SLAB_TIMEOUT_S");
    const double sec = e ? atof(e) : 20.0;
    return (unsigned long long)((sec > 0 ? sec : 20.0) * 1e9);
  }();
  slab_barrier_kernel<<<1, 32, 0, s>>>(F, R.rank, g->nranks, g->epoch, R.flags + QGS_MAX_RANKS, timeout_ns);
  SB_LAUNCH_CHECK();
  return 0;
}

int buf_index(const SlabRank& R, const void* p) {
  return p == R.core->y ? 0 : (p == R.core->Ya ? 1 : 2);
}

// One evaluation on every local slab.  role[]: indices (0 = y, 1 = Ya, 2 = Yb) of the buffers
// holding Yin / the step start state / Yout for this stage, identical on all ranks.
template <typename T>
int qgs_eval(somax_b200_qgs_s* g, const somax_b200_params* p, int in_b, int y_b, int out_b, int e,
             double hd, bool store_f, int f_slot, cudaStream_t s) {
  auto bufp = [](SlabRank& R, int b) -> void* { return b == 0 ? R.core->y : (b == 1 ? R.core->Ya : R.core->Yb); };
  for (SlabRank& R : g->local)
    if (int rc = qg_solver_rows_fwd<T>(R.core->solver, (const T*)bufp(R, in_b), 1, s)) return rc;
  for (SlabRank& R : g->local)
    if (int rc = seg_launch("slab_x1_border", R.xb, s)) return rc;
  if (int rc = qgs_barrier(g, s)) return rc;
  for (SlabRank& R : g->local) {
    // halo rows of Yin pushed by the neighbours after the previous evaluation
    if (int rc = seg_launch("slab_halo_in", R.xin[in_b], s)) return rc;
    if (int rc = qg_solver_cols<T>(R.cols, 1, R.s0, R.s1, s)) return rc;
  }
  for (SlabRank& R : g->local)
    if (int rc = qg_solver_border_stage<T>(R.cols, -1, 0, -1, s)) return rc;
  for (SlabRank& R : g->local)
    if (int rc = seg_launch("slab_x2_partials", R.x2, s)) return rc;
  if (int rc = qgs_barrier(g, s)) return rc;
  for (SlabRank& R : g->local) {
    if (int rc = qg_solver_border_stage<T>(R.cols, 0, 0, -1, s)) return rc;
    if (int rc = qg_solver_border_stage<T>(R.cols, 1, slab_r0(g, R.rank), slab_r1(g, R.rank), s)) return rc;
  }
  for (SlabRank& R : g->local)
    if (int rc = seg_launch("slab_xg_border", R.xg1, s)) return rc;
  if (int rc = qgs_barrier(g, s)) return rc;
  for (SlabRank& R : g->local)
    if (int rc = qg_solver_border_stage<T>(R.cols, 2, slab_r0(g, R.rank), slab_r1(g, R.rank), s)) return rc;
  for (SlabRank& R : g->local) {
    if (int rc = seg_launch("slab_xg_border", R.xg2, s)) return rc;
    if (int rc = seg_launch("slab_xg_border", R.xg2f, s)) return rc;
  }
  if (int rc = qgs_barrier(g, s)) return rc;
  for (SlabRank& R : g->local)
    if (int rc = qg_solver_cols<T>(R.cols, 2, R.s0, R.s1, s)) return rc;
  if (int rc = qgs_barrier(g, s)) return rc;
  for (SlabRank& R : g->local) {
    const int keep = keep_psi_ring(R.core);
    if (keep)
      if (int rc = seg_launch("slab_ring_cols", R.xring, s)) return rc;
    if (int rc = qg_solver_rows_inv<T>(R.core->solver, (T*)R.core->psi, keep, s)) return rc;
  }
  for (SlabRank& R : g->local)
    if (int rc = seg_launch("slab_halo_psi", R.xpsi, s)) return rc;
  if (int rc = qgs_barrier(g, s)) return rc;
  for (SlabRank& R : g->local) {
    QgArgs<T> A = make_qargs<T>(R.core, p, 1);
    Stage<T> st = qstage<T>();
    st.Yin[0] = (const T*)bufp(R, in_b);
    st.Yout[0] = (T*)bufp(R, out_b);
    st.dt = (T)hd; st.a_new = (T)TSIT5_A[e][e];
    if (e > 0) {
      st.nprev = e; st.y[0] = (const T*)bufp(R, y_b);
      for (int jj = 0; jj < e; ++jj) { st.a[jj] = (T)TSIT5_A[e][jj]; st.Fprev[jj][0] = (const T*)R.core->F[jj]; }
    }
    st.Fout[0] = store_f ? (T*)R.core->F[f_slot] : nullptr;
    if (int rc = launch_stencil<T>(R.core, A, st, hd, s)) return rc;
  }
  for (SlabRank& R : g->local)
    if (int rc = seg_launch("slab_halo_state", R.xstate[out_b], s)) return rc;
  return 0;
}

template <typename T>
int qgs_steps_impl(somax_b200_qgs_s* g, void* const* q, long n_steps, double dt, double dt_last,
                   const somax_b200_params* p, cudaStream_t s) {
  const long total = n_steps + (dt_last > 0 ? 1 : 0);
  int y = 0, Yc = 1, Yn = 2;
  auto bufp = [](SlabRank& R, int b) -> void* { return b == 0 ? R.core->y : (b == 1 ? R.core->Ya : R.core->Yb); };
  for (size_t v = 0; v < g->local.size(); ++v) {
    SlabRank& R = g->local[v];
    if (int rc = pack_field<T>((const T*)q[v], (T*)R.core->y, R.core->L, s)) return rc;
    if (int rc = qg_bc_inplace<T>(R.core, R.core->y, s)) return rc;     // x ring everywhere, y ring on the edge ranks
  }
  int newest = 0;      // buffer whose halo rows are still in the neighbours' inboxes (0: the caller's are valid)
  bool pending = false;
  if (total > 0) {
    auto step_dt = [&](long i) { return (i < n_steps) ? dt : dt_last; };
    // the very first evaluation reads the caller's halo rows: make the inbox agree with them
    for (SlabRank& R : g->local) {
      // inbox <- own ghost rows of y (so that the generic "inbox -> ghost rows" copy is a no-op)
      const Layout& L = R.core->L;
      const size_t es = sizeof(T), rowb = (size_t)L.pitch * es, planeb = L.plane() * es;
      const int planes = L.batch * L.nl;
      SB_CUDA(cudaMemcpy2DAsync(R.inbox, rowb, R.core->y, planeb, rowb, planes, cudaMemcpyDeviceToDevice, s));
      SB_CUDA(cudaMemcpy2DAsync((char*)R.inbox + (size_t)planes * rowb, rowb,
                                (char*)R.core->y + (size_t)(L.Ny - 1) * rowb, planeb, rowb, planes,
                                cudaMemcpyDeviceToDevice, s));
    }
    if (int rc = qgs_barrier(g, s)) return rc;      // nobody pushes into an inbox before it is primed
    if (int rc = qgs_eval<T>(g, p, y, y, Yc, 0, step_dt(0), true, 0, s)) return rc;
    auto stages = [&](double hd) -> int {
      for (int e = 1; e <= 5; ++e) {
        if (int rc = qgs_eval<T>(g, p, Yc, y, Yn, e, hd, e <= 4, e, s)) return rc;
        std::swap(Yc, Yn);
      }
      return 0;
    };
    auto full_step = [&](double hd, double hnext) -> int {
      if (int rc = stages(hd)) return rc;
      if (int rc = qgs_eval<T>(g, p, Yc, Yc, Yn, 0, hnext, true, 0, s)) return rc;
      const int oy = y; y = Yc; Yc = Yn; Yn = oy;
      return 0;
    };
    for (long i = 0; i + 1 < total; ++i)
      if (int rc = full_step(step_dt(i), step_dt(i + 1))) return rc;
    if (int rc = stages(step_dt(total - 1))) return rc;
    std::swap(y, Yc);
    newest = y; pending = true;
  }
  if (pending) {
    // halo rows of the final state: pushed after its stencil; land them before handing the slab back
    if (int rc = qgs_barrier(g, s)) return rc;
    for (SlabRank& R : g->local)
      if (int rc = seg_launch("slab_halo_in", R.xin[newest], s)) return rc;
    if (int rc = qgs_barrier(g, s)) return rc;      // the inboxes may be primed again by the next call
  }
  for (size_t v = 0; v < g->local.size(); ++v) {
    SlabRank& R = g->local[v];
    if (int rc = unpack_field<T>((const T*)bufp(R, y), (T*)q[v], R.core->L, s)) return rc;
  }
  return 0;
}

}  // namespace

extern "C" {

int somax_b200_qgs_create(somax_b200_qgs_t* out, int dtype, int nl, int ny, int nx, double dx,
                          double dy, const double* Cl2m, const double* Cm2l, const double* lambdas,
                          const double* beta_y, const double* wind, int nranks, int rank_first,
                          int nlocal, unsigned spec_flags) {
  if (!out) return fail(SOMAX_B200_ERR_INVALID, "out is null");
  *out = nullptr;
  if (nranks < 1 || nranks > QGS_MAX_RANKS || rank_first < 0 || nlocal < 1 || rank_first + nlocal > nranks)
    return fail(SOMAX_B200_ERR_INVALID, "need 1 <= nranks <= 16 and local ranks inside [0, nranks)");
  if (nlocal != nranks && nlocal != 1)
    return fail(SOMAX_B200_ERR_UNSUPPORTED, "a process holds either one slab or all of them");
  const bool pow2 = nx >= 64 && (nx & (nx - 1)) == 0;
  if (!pow2 || (nx / 64) % nranks != 0 || ny % nranks != 0 || ny / nranks < 3)
    return fail(SOMAX_B200_ERR_UNSUPPORTED,
                "slab decomposition needs nx = 2^p with nx/64 divisible by nranks, and ny divisible by nranks (>= 3 rows each)");
  if (!Cl2m || !Cm2l || !lambdas || !beta_y || !wind) return fail(SOMAX_B200_ERR_INVALID, "null coefficient pointer");
  if (int rc = require_device()) return rc;
  auto* g = new somax_b200_qgs_s();
  g->dtype = dtype; g->nl = nl; g->ny = ny; g->nx = nx; g->dx = dx; g->dy = dy;
  g->nranks = nranks; g->rank_first = rank_first; g->nlocal = nlocal;
  g->ny_loc = ny / nranks; g->spr = (nx / 64) / nranks;
  // y-sweeps: with few strips per rank a sweep CTA is a bare serial chain over ny/2 rows, so the
  // chains are cut into segments (two-pass, see ThomasTab) once the strips no longer fill the GPU
  g->nseg = nranks >= 4 ? 8 : 1;      // measured at 4 GPUs: 8 segments 12.4 ms/step, 4: 12.7, 1: 12.7; at 8 GPUs 9.2 vs 10.5
  if (const char* e = getenv("SOMAX_B200_SLAB_NSEG")) g->nseg = std::max(1, std::min(atoi(e), 16));
  const int Nx = nx + 2, nyl = g->ny_loc;
  const size_t es = qgs_es(g);
  int rc = 0;
  g->local.resize(nlocal);
  for (int v = 0; v < nlocal && !rc; ++v) {
    SlabRank& R = g->local[v];
    R.rank = rank_first + v; R.s0 = R.rank * g->spr; R.s1 = R.s0 + g->spr;
    const size_t row0 = (size_t)R.rank * nyl;      // the slab's window starts at global row row0
    // the slab's row stage covers the solver rows it owns: window rows [jo, jo + rows)
    const int jo = R.rank == 0 ? 0 : 1, rows = slab_r1(g, R.rank) - slab_r0(g, R.rank);
    rc = qg_create_impl(&R.core, dtype, 1, nl, nyl, nx, dx, dy, Cl2m, Cm2l, lambdas,
                        beta_y + row0 * Nx, wind + row0 * Nx, SOMAX_B200_SOLVER_FFT, spec_flags,
                        rows, jo, R.rank == 0, R.rank == nranks - 1);
    if (rc) break;
    R.core->bc_ylo = R.rank == 0; R.core->bc_yhi = R.rank == nranks - 1;
    rc = qg_solver_create(&R.cols, dtype, 1, nl, ny, nx, dx, dy, Cl2m, Cm2l, lambdas, SOMAX_B200_SOLVER_FFT, g->nseg);
    if (rc) break;
    rc = qg_solver_set_rank_reduce(R.cols, nranks, R.rank, R.s0, R.s1);
    if (rc) break;
    const size_t ib = 2 * (size_t)nl * R.core->L.pitch * es;
    if (cudaMalloc(&R.inbox, ib) != cudaSuccess || cudaMemset(R.inbox, 0, ib) != cudaSuccess ||
        cudaMalloc((void**)&R.flags, 2 * QGS_MAX_RANKS * sizeof(unsigned)) != cudaSuccess ||
        cudaMemset(R.flags, 0, 2 * QGS_MAX_RANKS * sizeof(unsigned)) != cudaSuccess) {
      rc = fail(SOMAX_B200_ERR_CUDA, "cudaMalloc (slab inbox / flags) failed");
      break;
    }
    g->bytes += somax_b200_qg_device_bytes(R.core) + qg_solver_bytes(R.cols) + ib;
  }
  if (!rc && nlocal == nranks) {
    for (int v = 0; v < nlocal; ++v) qgs_local_ptrs(g->local[v], g->peers[v].buf);
    for (int v = 0; v < nlocal && !rc; ++v) rc = qgs_build_tables(g, g->local[v]);
    g->attached = !rc;
  }
  if (!rc) rc = (cudaDeviceSynchronize() == cudaSuccess) ? 0 : fail(SOMAX_B200_ERR_CUDA, "slab create: device error");
  if (rc) { somax_b200_qgs_destroy(g); return rc; }
  *out = g;
  return 0;
}

int somax_b200_qgs_destroy(somax_b200_qgs_t g) {
  if (!g) return 0;
  cudaDeviceSynchronize();
  for (void* p : g->ipc_opened) cudaIpcCloseMemHandle(p);
  for (SlabRank& R : g->local) {
    SegTable* ts[] = {&R.xb, &R.x2, &R.xpsi, &R.xstate[0], &R.xstate[1], &R.xstate[2],
                      &R.xin[0], &R.xin[1], &R.xin[2], &R.xg1, &R.xg2, &R.xg2f, &R.xring};
    for (SegTable* t : ts) cudaFree(t->dev);
    cudaFree(R.inbox); cudaFree(R.flags);
    qg_solver_destroy(R.cols);
    somax_b200_qg_destroy(R.core);
  }
  delete g;
  return 0;
}

size_t somax_b200_qgs_device_bytes(somax_b200_qgs_t g) { return g ? g->bytes : 0; }
size_t somax_b200_qgs_export_bytes(void) { return sizeof(IpcBlob); }

int somax_b200_qgs_export(somax_b200_qgs_t g, void* blob) {
  if (!g || !blob) return fail(SOMAX_B200_ERR_INVALID, "null argument");
  if (g->nlocal != 1) return fail(SOMAX_B200_ERR_INVALID, "export is for one-slab-per-process groups");
  void* ptrs[QGS_NBUF];
  qgs_local_ptrs(g->local[0], ptrs);
  IpcBlob b;
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
  for (int i = 0; i < QGS_NBUF; ++i) {
    cudaError_t e = cudaIpcGetMemHandle(&b.handle[i], ptrs[i]);
    if (e != cudaSuccess) return fail(SOMAX_B200_ERR_COMM, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    CUdeviceptr base = 0; size_t sz = 0;
    if (get_range && get_range(&base, &sz, (CUdeviceptr)ptrs[i]) == CUDA_SUCCESS)
      b.offset[i] = (unsigned long long)((CUdeviceptr)ptrs[i] - base);
  }
  memcpy(blob, &b, sizeof(b));
  return 0;
}

int somax_b200_qgs_attach(somax_b200_qgs_t g, const void* blobs) {
  if (!g || !blobs) return fail(SOMAX_B200_ERR_INVALID, "null argument");
  if (g->nlocal != 1) return fail(SOMAX_B200_ERR_INVALID, "attach is for one-slab-per-process groups");
  if (g->attached) return fail(SOMAX_B200_ERR_INVALID, "already attached");
  const IpcBlob* B = reinterpret_cast<const IpcBlob*>(blobs);
  const int me = g->local[0].rank;
  for (int r = 0; r < g->nranks; ++r) {
    if (r == me) { qgs_local_ptrs(g->local[0], g->peers[r].buf); continue; }
    for (int i = 0; i < QGS_NBUF; ++i) {
      void* p = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&p, B[r].handle[i], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess)
        return fail(SOMAX_B200_ERR_COMM, std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(r) + "): " + cudaGetErrorString(e));
      g->ipc_opened.push_back(p);
      g->peers[r].buf[i] = (char*)p + B[r].offset[i];
    }
  }
  if (int rc = qgs_build_tables(g, g->local[0])) return rc;
  g->attached = true;
  return 0;
}

int somax_b200_qgs_steps(somax_b200_qgs_t g, void* const* q_slabs, long n_steps, double dt,
                         double dt_last, const somax_b200_params* p, void* stream) {
  if (!g || !q_slabs || !p) return fail(SOMAX_B200_ERR_INVALID, "null argument");
  if (!g->attached) return fail(SOMAX_B200_ERR_INVALID, "slab group not attached to its peers");
  if (n_steps < 0 || !(dt > 0) || dt_last < 0) return fail(SOMAX_B200_ERR_INVALID, "need n_steps>=0, dt>0, dt_last>=0");
  for (int v = 0; v < g->nlocal; ++v)
    if (!q_slabs[v]) return fail(SOMAX_B200_ERR_INVALID, "null slab pointer");
  return g->dtype == SOMAX_B200_F32
             ? qgs_steps_impl<float>(g, q_slabs, n_steps, dt, dt_last, p, (cudaStream_t)stream)
             : qgs_steps_impl<double>(g, q_slabs, n_steps, dt, dt_last, p, (cudaStream_t)stream);
}

int somax_b200_qgs_status(somax_b200_qgs_t g, int* barrier_timeouts) {
  if (!g || !barrier_timeouts) return fail(SOMAX_B200_ERR_INVALID, "null argument");
  *barrier_timeouts = 0;
  for (SlabRank& R : g->local) {
    unsigned e = 0;
    SB_CUDA(cudaMemcpy(&e, R.flags + QGS_MAX_RANKS, sizeof(e), cudaMemcpyDeviceToHost));
    *barrier_timeouts += (int)e;
  }
  return 0;
}

}  // extern "C"
