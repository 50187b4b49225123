"""DST-I Helmholtz solve: numpy/scipy restatement of ``finitevolx.pv_inversion`` /
``streamfunction_from_vorticity`` with ``bc="dst"`` (spectraldiffx v0.0.10).

Test infrastructure only.  PARITY UNPINNED (dependency absent; see oracle/__init__.py).
ref call sites: somax/_src/models/qg/baroclinic.py:146-152, qg/barotropic.py:119-121;
ghost-cell convention: somax/_src/models/pde2d/poisson.py:15-17,31; recipe: SURVEY App. B.4.
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
    """Solve (laplacian - lambda_m) psi_m = rhs_m on the interior, homogeneous
    Dirichlet imposed at the ghost ring; returns the full array with a zero ring.

    ``rhs``: (..., Ny, Nx); ``lambdas``: scalar or (nl,) matched to axis -3.
    Arithmetic is done in ``rhs.dtype`` (float32 or float64).
    """
    dt = rhs.dtype
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
    out = np.zeros_like(rhs)
    out[..., 1:-1, 1:-1] = psi_i.astype(dt, copy=False)
    return out
