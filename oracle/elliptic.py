"""DST-I Helmholtz solve: numpy/scipy restatement of ``finitevolx.pv_inversion`` /
``streamfunction_from_vorticity`` with ``bc="dst"`` (spectraldiffx v0.0.10).

Test infrastructure only.  PINNED by the reference's own executed tutorials: the outputs stored
in /root/reference/content/tutorials/step09_laplace_2d.ipynb (cells 7, 11, 13),
step10_poisson_2d.ipynb (cells 4, 12) and step11_helmholtz_2d.ipynb (cell 6) - L2 / Linf errors
and solution maxima printed by ``PoissonSolver2D`` / ``HelmholtzSolver2D``
(somax/_src/models/pde2d/poisson.py:36-38,129-133, the same finitevolx call the QG models
make) - are reproduced to 6-7 digits by exactly one convention
(tests/test_oracle_reference_pins.py, tests/golden/reference_notebook_outputs.json):

  * DST-I in both directions with the 5-point FINITE-DIFFERENCE eigenvalues
    (the continuous ones of the tutorial prose, step09_laplace_2d.py:246, miss in the 3rd digit);
  * on the WHOLE array that is passed in, ghost ring included: Ny x Nx unknowns, the ring values
    of the right-hand side are used, and the homogeneous Dirichlet condition sits one cell
    OUTSIDE the array.  (The interior-only solve with a zero ring that the docstring of
    poisson.py:15-17,31 suggests gives 2.208e-02 where the reference printed 5.585729e-02.)

ref call sites: somax/_src/models/qg/baroclinic.py:146-152, qg/barotropic.py:119-121.
"""
from __future__ import annotations

import numpy as np
import scipy.fft

from .operators import DEFAULT_SPEC, OperatorSpec


def dst1_eigenvalues(n: int, d: float, spec: OperatorSpec = DEFAULT_SPEC) -> np.ndarray:
    """Eigenvalues of the 1-D Dirichlet Laplacian on n interior points (float64)."""
    k = np.arange(1, n + 1, dtype=np.float64)
    if spec.dst_fd_eigenvalues:
        return -(4.0 / (d * d)) * np.sin(np.pi * k / (2.0 * (n + 1))) ** 2
    return -((np.pi * k / ((n + 1) * d)) ** 2)


def helmholtz_dst(rhs, dx, dy, lambdas, spec: OperatorSpec = DEFAULT_SPEC, workers=None):
    """Solve (laplacian - lambda_m) psi_m = rhs_m.  Default (the reference's behaviour, see the
    module docstring): every point of the (Ny, Nx) array is an unknown, psi = 0 one cell outside
    the array.  ``spec.dst_full_array = False``: interior unknowns only, zero ghost ring.

    ``rhs``: (..., Ny, Nx); ``lambdas``: scalar or (nl,) matched to axis -3.
    Arithmetic is done in ``rhs.dtype`` (float32 or float64).
    """
    dt = rhs.dtype
    if spec.dst_full_array:
        ny, nx = rhs.shape[-2], rhs.shape[-1]
        r = np.ascontiguousarray(rhs)
    else:
        ny, nx = rhs.shape[-2] - 2, rhs.shape[-1] - 2
        r = np.ascontiguousarray(rhs[..., 1:-1, 1:-1])
    rh = scipy.fft.dstn(r, type=1, axes=(-2, -1), workers=workers)
    lx = dst1_eigenvalues(nx, dx, spec)
    ly = dst1_eigenvalues(ny, dy, spec)
    lam = np.asarray(lambdas, dtype=np.float64)
    denom = ly[:, None] + lx[None, :]
    if lam.ndim == 1:
        denom = denom[None] - lam[:, None, None]
    else:
        denom = denom - lam
    rh = rh / denom.astype(dt)
    psi_i = scipy.fft.idstn(rh, type=1, axes=(-2, -1), workers=workers)
    if spec.dst_full_array:
        return psi_i.astype(dt, copy=False)
    out = np.zeros_like(rhs)
    out[..., 1:-1, 1:-1] = psi_i.astype(dt, copy=False)
    return out
