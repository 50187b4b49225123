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

int qg_solver_create(QgSolver** out, int dtype, int batch, int nl, int ny, int nx, double dx,
                     double dy, const double* Cl2m, const double* Cm2l, const double* lambdas,
                     int solver_kind, int nseg = 1);   // nseg > 1: segmented y-sweeps (see ThomasTab)
void qg_solver_destroy(QgSolver* s);
size_t qg_solver_bytes(const QgSolver* s);
int qg_solver_kind(const QgSolver* s);

// psi = ring0(Cm2l . Helm^-1 . Cl2m . q) on padded planes (batch, nl, Ny, pitch).
// Only the interior of q is read; only the interior of psi is written (its ring stays 0).
template <typename T>
int qg_solver_run(QgSolver* s, const T* q, T* psi, cudaStream_t stream);

// Stages of the FFT-path inversion, for the slab-distributed model (qg_slab.cuh):
//   rows_fwd: layer->mode mix + DST-I in x of every row of q into the solver's spectral array S
//   cols(1, sa, sb): first Thomas solve in y on strips [sa, sb) (64 x-wavenumbers each) of S,
//                    border row-sum partials into `part`
//   border:  reduction of all partials + Schur solve of the border column (needs every strip's
//            partials and the raw border column of S)
//   cols(2, sa, sb): second Thomas solve on the strips
//   rows_inv: inverse transform + mode->layer mix of S into psi
template <typename T> int qg_solver_rows_fwd(QgSolver* s, const T* q, cudaStream_t stream);
template <typename T> int qg_solver_rows_inv(QgSolver* s, T* psi, cudaStream_t stream);
template <typename T> int qg_solver_cols(QgSolver* s, int phase, int sa, int sb, cudaStream_t stream);
template <typename T> int qg_solver_border(QgSolver* s, cudaStream_t stream);
// border in three stages (0 reduce; 1 first DST -> ghat[a0, a1); 2 second DST -> gvec / gvecf /
// border column of S for rows [a0, a1)); a1 < 0 = all rows
template <typename T> int qg_solver_border_stage(QgSolver* s, int stage, int a0, int a1, cudaStream_t stream);

// Raw view of the spectral storage: S is [plane][strip][ny][64] (strip = 64 x-wavenumbers),
// part is [plane][2 * nstrip][ny]; ncols is the x index of the border column.
struct QgSolverView {
  void* S; void* part;
  double* ghat; double* gvec; float* gvecf;   // border system: [plane][ny]
  int ny, nx, np, planes, nstrip, ncols, kind, nheavy;
};
QgSolverView qg_solver_view(const QgSolver* s);

}  // namespace sb
