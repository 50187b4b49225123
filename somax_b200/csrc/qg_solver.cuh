// Internal interface of the PV-inversion solver (qg_solver.cu) used by qg.cu.
#pragma once
#include "common.cuh"

namespace sb {

struct QgSolver;

int qg_solver_create(QgSolver** out, int dtype, int batch, int nl, int ny, int nx, double dx,
                     double dy, const double* Cl2m, const double* Cm2l, const double* lambdas,
                     int solver_kind);
void qg_solver_destroy(QgSolver* s);
size_t qg_solver_bytes(const QgSolver* s);
int qg_solver_kind(const QgSolver* s);

// psi = ring0(Cm2l . Helm^-1 . Cl2m . q) on padded planes (batch, nl, Ny, pitch).
// Only the interior of q is read; only the interior of psi is written (its ring stays 0).
template <typename T>
int qg_solver_run(QgSolver* s, const T* q, T* psi, cudaStream_t stream);

}  // namespace sb
