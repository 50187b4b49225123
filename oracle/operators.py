"""Arakawa C-grid operators: numpy restatement of the finitevolx calls somax makes.

Test infrastructure only (see ``oracle/__init__.py``).  The conventions that used to be open
(App. E) are PINNED by numbers the reference itself printed in its executed tutorials
(tests/golden/reference_notebook_outputs.json, tests/test_oracle_reference_pins.py).

Arrays are ``[..., j, i] = [..., y, x]`` with shape ``(..., Ny, Nx)``, one ghost
ring (``Ny = ny + 2``).  Co-located indexing: ``T[j,i]`` cell centre, ``U[j,i]``
its east face, ``V[j,i]`` its north face, ``X[j,i]`` its NE corner.  Every
finitevolx operator writes ``out[..., 1:-1, 1:-1]`` and leaves the ghost ring
ZERO; second-level operators near the boundary therefore
read zeros.  All functions broadcast over leading (layer) axes, which is what
``finitevolx.multilayer`` (vmap over axis 0) does.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class OperatorSpec:
    """The finitevolx/spectraldiffx conventions SURVEY.md App. E left open.  The DEFAULTS are the
    ones that reproduce the reference's own printed tutorial outputs (module docstring); the
    other settings are kept so that the tests can show they do NOT.

    One switch flips oracle and CUDA kernels together (the same bits travel
    through the C ABI as ``spec_flags``).
    """

    #: Advection2D writes [2:-2,2:-2] (True) or [1:-1,1:-1] (False).  Pinned True by
    #: step17_shallow_water_2d.ipynb cells 9 and 18 (False: max|u| 0.909 vs 0.8879 printed).
    advection_region2: bool = True
    #: Diffusion2D = flux form with interior-only (zero-ghost) face fluxes (True) or
    #: nu * 5-point laplacian on [1:-1,1:-1] (False).  Pinned True by step17 cell 18:
    #: d(sum u^2)/d(nu) = -3.126971e-04 printed; flux form -3.1268e-04, laplacian -3.0798e-04.
    diffusion_flux_form: bool = True
    #: DST Helmholtz eigenvalues: 5-point finite-difference (True) or continuous.  Pinned True
    #: by step09 cell 7 (5.585729e-02 printed; FD 5.585726e-02, continuous 5.5778e-02).  Only
    #: the FD flavour exists on the GPU (it is what makes the transform-in-x /
    #: tridiagonal-in-y factorisation exact): the C ABI refuses the other one.
    dst_fd_eigenvalues: bool = True
    #: The DST solve takes every point of the array it is given as an unknown (True) or the
    #: interior only with a zero ring (False).  Pinned True by the same cells; the C ABI
    #: refuses False.
    dst_full_array: bool = True

    def flags(self) -> int:
        return ((1 if self.advection_region2 else 0) | (2 if self.diffusion_flux_form else 0)
                | (0 if self.dst_fd_eigenvalues else 4) | (0 if self.dst_full_array else 8))


DEFAULT_SPEC = OperatorSpec()

_I = (Ellipsis, slice(1, -1), slice(1, -1))


def _zeros_like(a):
    return np.zeros_like(a)


# --- boundary helpers ------------------------------------------------------
def zero_boundaries(x):
    """ring := 0.  ref: somax/_src/models/qg/baroclinic.py:158,194 (finitevolx.zero_boundaries);
    pinned by tests/models/test_qg_baroclinic.py:166-183."""
    out = x.copy()
    out[..., 0, :] = 0
    out[..., -1, :] = 0
    out[..., :, 0] = 0
    out[..., :, -1] = 0
    return out


def enforce_periodic(x):
    """Ghost ring := opposite interior row/column (rows first, then columns, so the
    corners take the diagonally opposite interior cell).
    ref: somax/_src/models/swm/multilayer.py:214-216; tests/models/test_navier_stokes.py:72-78."""
    out = x.copy()
    out[..., 0, :] = out[..., -2, :]
    out[..., -1, :] = out[..., 1, :]
    out[..., :, 0] = out[..., :, -2]
    out[..., :, -1] = out[..., :, 1]
    return out


def wall_bc_u(u):
    """ref: somax/_src/models/swm/multilayer.py:386-392 (in-tree, exact order)."""
    u = u.copy()
    u[..., :, 0] = 0
    u[..., :, -2] = 0
    u[..., :, -1] = 0
    u[..., 0, :] = u[..., 1, :]
    u[..., -1, :] = u[..., -2, :]
    return u


def wall_bc_v(v):
    """ref: somax/_src/models/swm/multilayer.py:395-401."""
    v = v.copy()
    v[..., 0, :] = 0
    v[..., -2, :] = 0
    v[..., -1, :] = 0
    v[..., :, 0] = v[..., :, 1]
    v[..., :, -1] = v[..., :, -2]
    return v


def wall_bc_h(h):
    """ref: somax/_src/models/swm/multilayer.py:404-408."""
    h = h.copy()
    h[..., 0, :] = h[..., 1, :]
    h[..., -1, :] = h[..., -2, :]
    h[..., :, 0] = h[..., :, 1]
    h[..., :, -1] = h[..., :, -2]
    return h


# --- differences ------------------------------------------
def diff_x_T_to_U(h, dx):
    """ref call sites: qg/baroclinic.py:204, swm/multilayer.py:186."""
    out = _zeros_like(h)
    out[_I] = (h[..., 1:-1, 2:] - h[..., 1:-1, 1:-1]) / dx
    return out


def diff_y_T_to_V(h, dy):
    """ref call sites: qg/baroclinic.py:203, swm/multilayer.py:187."""
    out = _zeros_like(h)
    out[_I] = (h[..., 2:, 1:-1] - h[..., 1:-1, 1:-1]) / dy
    return out


def laplacian(h, dx, dy):
    """5-point Laplacian.  ref call sites: qg/baroclinic.py:184,188,216."""
    out = _zeros_like(h)
    c = h[..., 1:-1, 1:-1]
    out[_I] = (h[..., 1:-1, 2:] - 2 * c + h[..., 1:-1, :-2]) / (dx * dx) + (
        h[..., 2:, 1:-1] - 2 * c + h[..., :-2, 1:-1]
    ) / (dy * dy)
    return out


def curl(u, v, dx, dy):
    """zeta at X points = dv/dx - du/dy.  ref: Vorticity2D.relative_vorticity, swm/multilayer.py:232."""
    out = _zeros_like(u)
    out[_I] = (v[..., 1:-1, 2:] - v[..., 1:-1, 1:-1]) / dx - (
        u[..., 2:, 1:-1] - u[..., 1:-1, 1:-1]
    ) / dy
    return out


# --- interpolations (App. B.2) ----------------------------------------------
def T_to_U(h):
    out = _zeros_like(h)
    out[_I] = 0.5 * (h[..., 1:-1, 1:-1] + h[..., 1:-1, 2:])
    return out


def T_to_V(h):
    out = _zeros_like(h)
    out[_I] = 0.5 * (h[..., 1:-1, 1:-1] + h[..., 2:, 1:-1])
    return out


def T_to_X(h):
    out = _zeros_like(h)
    out[_I] = 0.25 * (
        h[..., 1:-1, 1:-1] + h[..., 1:-1, 2:] + h[..., 2:, 1:-1] + h[..., 2:, 2:]
    )
    return out


def X_to_U(q):
    out = _zeros_like(q)
    out[_I] = 0.5 * (q[..., 1:-1, 1:-1] + q[..., :-2, 1:-1])
    return out


def X_to_V(q):
    out = _zeros_like(q)
    out[_I] = 0.5 * (q[..., 1:-1, 1:-1] + q[..., 1:-1, :-2])
    return out


def U_to_T(u):
    out = _zeros_like(u)
    out[_I] = 0.5 * (u[..., 1:-1, 1:-1] + u[..., 1:-1, :-2])
    return out


def V_to_T(v):
    out = _zeros_like(v)
    out[_I] = 0.5 * (v[..., 1:-1, 1:-1] + v[..., :-2, 1:-1])
    return out


def V_to_U(v):
    out = _zeros_like(v)
    out[_I] = 0.25 * (
        v[..., 1:-1, 1:-1] + v[..., 1:-1, 2:] + v[..., :-2, 1:-1] + v[..., :-2, 2:]
    )
    return out


def U_to_V(u):
    out = _zeros_like(u)
    out[_I] = 0.25 * (
        u[..., 1:-1, 1:-1] + u[..., 2:, 1:-1] + u[..., 1:-1, :-2] + u[..., 2:, :-2]
    )
    return out


# --- composite operators ------------------------------------------------------
def arakawa_jacobian(f, g, dx, dy):
    """Arakawa (1966) 9-point Jacobian J(f,g)=f_x g_y - f_y g_x, INTERIOR ONLY
    (shape ``(..., Ny-2, Nx-2)``).  ref: somax/_src/models/qg/baroclinic.py:175-177,
    qg/barotropic.py:137-139; formula SURVEY App. B.3."""
    c = slice(1, -1)
    fE, fW = f[..., c, 2:], f[..., c, :-2]
    fN, fS = f[..., 2:, c], f[..., :-2, c]
    fNE, fNW = f[..., 2:, 2:], f[..., 2:, :-2]
    fSE, fSW = f[..., :-2, 2:], f[..., :-2, :-2]
    gE, gW = g[..., c, 2:], g[..., c, :-2]
    gN, gS = g[..., 2:, c], g[..., :-2, c]
    gNE, gNW = g[..., 2:, 2:], g[..., 2:, :-2]
    gSE, gSW = g[..., :-2, 2:], g[..., :-2, :-2]
    jpp = (fE - fW) * (gN - gS) - (fN - fS) * (gE - gW)
    jpx = fE * (gNE - gSE) - fW * (gNW - gSW) - fN * (gNE - gNW) + fS * (gSE - gSW)
    jxp = gN * (fNE - fNW) - gS * (fSE - fSW) - gE * (fNE - fSE) + gW * (fNW - fSW)
    return (jpp + jpx + jxp) / (12.0 * dx * dy)


def potential_vorticity(u, v, h, f, dx, dy):
    """q = (zeta + f_X) / h_X at X points, interior only.
    ref: Vorticity2D.potential_vorticity, swm/multilayer.py:164; App. B.5."""
    zeta = curl(u, v, dx, dy)
    fX = T_to_X(np.broadcast_to(f, h.shape))
    hX = T_to_X(h)
    out = _zeros_like(h)
    out[_I] = (zeta[_I] + fX[_I]) / hX[_I]
    return out


def kinetic_energy(u, v):
    """ke = 0.5*(U_to_T(u^2) + V_to_T(v^2)) at T points.  ref: swm/multilayer.py:178."""
    out = _zeros_like(u)
    out[_I] = 0.5 * (U_to_T(u * u)[_I] + V_to_T(v * v)[_I])
    return out


def advection_upwind1(h, u, v, dx, dy, spec: OperatorSpec = DEFAULT_SPEC):
    """-div(h u) with first-order upwind face values.
    ref: Advection2D(h,u,v,method="upwind1"), swm/multilayer.py:160; App. B.5."""
    fe = _zeros_like(h)
    fn = _zeros_like(h)
    uc, vc = u[_I], v[_I]
    fe[_I] = uc * np.where(uc > 0, h[..., 1:-1, 1:-1], h[..., 1:-1, 2:])
    fn[_I] = vc * np.where(vc > 0, h[..., 1:-1, 1:-1], h[..., 2:, 1:-1])
    out = _zeros_like(h)
    if spec.advection_region2:
        s = (Ellipsis, slice(2, -2), slice(2, -2))
        out[s] = -(
            (fe[..., 2:-2, 2:-2] - fe[..., 2:-2, 1:-3]) / dx
            + (fn[..., 2:-2, 2:-2] - fn[..., 1:-3, 2:-2]) / dy
        )
    else:
        out[_I] = -(
            (fe[..., 1:-1, 1:-1] - fe[..., 1:-1, :-2]) / dx
            + (fn[..., 1:-1, 1:-1] - fn[..., :-2, 1:-1]) / dy
        )
    return out


def diffusion(f, nu, dx, dy, spec: OperatorSpec = DEFAULT_SPEC):
    """div(nu grad f).  ref: Diffusion2D(f, nu), swm/multilayer.py:194-195."""
    if not spec.diffusion_flux_form:
        return nu * laplacian(f, dx, dy)
    fx = nu * diff_x_T_to_U(f, dx)
    fy = nu * diff_y_T_to_V(f, dy)
    out = _zeros_like(f)
    out[_I] = (fx[..., 1:-1, 1:-1] - fx[..., 1:-1, :-2]) / dx + (
        fy[..., 1:-1, 1:-1] - fy[..., :-2, 1:-1]
    ) / dy
    return out
