// Internal interface of the PV-inversion solver (qg_solver.cu) used by qg.cu.
#pragma once
#include "common.cuh"

namespace sb {

struct QgSolver;

// Spectral arrays are stored BLOCKED in 64-column strips: within a plane, element (row j,
// x-wavenumber k) lives at ((k / 64) * ny + j) * 64 + (k % 64); the plane stride is ny * np.
constexpr int SP_W = 64;
__host__ __device__ __forceinline__ size_t sp_off(int ny, int j, int k) {
  return ((size_t)(k >> 6) * ny + j) * SP_W + (k & (SP_W - 1));
}

// ny, nx: INTERIOR sizes of the field arrays (Ny = ny + 2, Nx = nx + 2).  Every point of the array
// is an unknown of the solve (the reference's DST acts on the whole array it is given,
// oracle/elliptic.py), so by default the solver has ny + 2 rows.  rows > 0 (slab-distributed
// model): the solver covers `rows` rows starting at field row jo; ylo / yhi say whether its first /
// last row is a physical ring row.  nseg > 1: segmented y-sweeps (see ThomasTab).
int qg_solver_create(QgSolver** out, int dtype, int batch, int nl, int ny, int nx, double dx,
                     double dy, const double* Cl2m, const double* Cm2l, const double* lambdas,
                     int solver_kind, int nseg = 1, int rows = -1, int jo = 0, int ylo = 1, int yhi = 1);
void qg_solver_destroy(QgSolver* s);
size_t qg_solver_bytes(const QgSolver* s);
// Slab-distributed model, fused compute + exchange (no transpose kernels):
//   scatter: rows_fwd of this (row-stage) solver stores spectral element (row j, strip st) into
//            peerS[st / spr], the column array (ny_cols rows per strip) of the rank that owns the
//            strip, at row row0 + j;
//   push:    the last sweep of cols(2) of this (column-stage) solver stores its tiles into peerR[r],
//            the row array of the rank that owns rows [row0[r], row0[r+1]).
// Pointers may be peer memory (CUDA IPC over NVLink); spr must be a power of two.
void qg_solver_set_scatter(QgSolver* s, void* const* peerS, int nranks, int spr, int row0, int ny_cols);
void qg_solver_set_push(QgSolver* s, void* const* peerR, int nranks, const int* row0);
// rank reduce: border stage -1 pre-sums the partials of strips [s0, s1) into slot `me` of
// pvec[nranks][planes][2][ny] (fp64); the slots of the other ranks arrive by exchange; border stage 0
// then adds the slots in rank order instead of walking every strip's partials.
int qg_solver_set_rank_reduce(QgSolver* s, int nranks, int me, int s0, int s1);
int qg_solver_kind(const QgSolver* s);

// psi = Cm2l . Helm^-1 . Cl2m . q on padded planes (batch, nl, Ny, pitch), every point of the
// array an unknown.  ring_zero: read the ghost ring of q as zero (boundary condition applied on
// load).  keep_ring = 0: the ring of psi is not written (BaroclinicQG zeroes it; it stays at its
// initial zero); keep_ring = 1: the solver's ring values are stored (BarotropicQG).
template <typename T>
int qg_solver_run(QgSolver* s, const T* q, T* psi, int ring_zero, int keep_ring, cudaStream_t stream);

// Stages of the FFT-path inversion, for the slab-distributed model (qg_slab.cuh):
//   rows_fwd: layer->mode mix + DST-I in x of every row of q into the solver's spectral array S
//   cols(1, sa, sb): first Thomas solve in y on strips [sa, sb) (64 x-wavenumbers each) of S,
//                    border row-sum partials into `part`
//   border:  reduction of all partials + Schur solve of the border column (needs every strip's
//            partials and the raw border column of S)
//   cols(2, sa, sb): second Thomas solve on the strips
//   rows_inv: inverse transform + mode->layer mix of S into psi
template <typename T> int qg_solver_rows_fwd(QgSolver* s, const T* q, int ring_zero, cudaStream_t stream);
template <typename T> int qg_solver_rows_inv(QgSolver* s, T* psi, int keep_ring, cudaStream_t stream);
template <typename T> int qg_solver_cols(QgSolver* s, int phase, int sa, int sb, cudaStream_t stream);
template <typename T> int qg_solver_border(QgSolver* s, cudaStream_t stream);
// border in three stages (0 reduce; 1 first DST -> ghat[a0, a1); 2 second DST -> gvec / gvecf /
// border column of S for rows [a0, a1)); a1 < 0 = all rows
template <typename T> int qg_solver_border_stage(QgSolver* s, int stage, int a0, int a1, cudaStream_t stream);

// Raw view of the spectral storage: S is [plane][strip][ny][64] (strip = 64 x-wavenumbers, ny =
// solver rows), part is [plane][2][2 * nstrip][ny], bext is [plane][3][ny] (border columns 0, nx+1
// and the raw column nx); ncols is the slot of border column nx in a row of S.
struct QgSolverView {
  void* S; void* part; double* pvec; void* bext;
  double* ghat; double* gvec; float* gvecf;   // border system: ghat [plane][3][ny], gvec / gvecf [plane][2][ny]
  int ny, nx, np, planes, nstrip, ncols, kind, nheavy;
};
QgSolverView qg_solver_view(const QgSolver* s);

}  // namespace sb
