"""Layer coupling matrix and vertical-mode transform (setup-time, host only).

Test infrastructure only.  UNPINNED (mode ordering / normalisation): the mode matrices are inputs on both sides.
ref: somax/_src/core/transforms.py:172-224 (ModalTransform.from_physics, to_modal, to_layer);
finitevolx.build_coupling_matrix / decompose_vertical_modes restated per SURVEY App. B.7
(MQGeometry convention).
"""
from __future__ import annotations

import numpy as np


def build_coupling_matrix(H, g_prime) -> np.ndarray:
    """MQGeometry-style (non-symmetric) tridiagonal stretching matrix A (float64)."""
    H = np.asarray(H, dtype=np.float64)
    g = np.asarray(g_prime, dtype=np.float64)
    nl = H.shape[0]
    if nl == 1:
        return np.array([[1.0 / (H[0] * g[0])]])
    A = np.zeros((nl, nl))
    A[0, 0] = 1.0 / (H[0] * g[0]) + 1.0 / (H[0] * g[1])
    A[0, 1] = -1.0 / (H[0] * g[1])
    for k in range(1, nl - 1):
        A[k, k - 1] = -1.0 / (H[k] * g[k])
        A[k, k] = (1.0 / g[k] + 1.0 / g[k + 1]) / H[k]
        A[k, k + 1] = -1.0 / (H[k] * g[k + 1])
    A[-1, -2] = -1.0 / (H[-1] * g[-1])
    A[-1, -1] = 1.0 / (H[-1] * g[-1])
    return A


def decompose_vertical_modes(A, f0):
    """(rossby_radii, Cl2m, Cm2l): right eigenvectors as Cm2l, bi-orthonormalised left
    eigenvectors as Cl2m, modes sorted by ascending eigenvalue (barotropic first)."""
    A = np.asarray(A, dtype=np.float64)
    wr, R = np.linalg.eig(A)
    wl, L = np.linalg.eig(A.T)
    wr, R, wl, L = wr.real, R.real, wl.real, L.real
    R = R[:, np.argsort(wr)]
    L = L[:, np.argsort(wl)]
    w = np.sort(wr)
    Cl2m = np.diag(1.0 / np.diag(L.T @ R)) @ L.T
    Cm2l = R
    with np.errstate(divide="ignore", invalid="ignore"):
        radii = 1.0 / (abs(f0) * np.sqrt(np.abs(w)))
    return radii, Cl2m, Cm2l


def reference_eigenvalues(A) -> np.ndarray:
    """``jnp.linalg.eigh(A)`` as called at somax/_src/core/transforms.py:194: jax's eigh
    symmetrises its input ((A+A^T)/2) by default.  For the non-symmetric A above these are
    NOT the eigenvalues Cl2m/Cm2l diagonalise; reproduced, not fixed."""
    A = np.asarray(A, dtype=np.float64)
    return np.linalg.eigvalsh(0.5 * (A + A.T))


class ModalTransform:
    """ref: somax/_src/core/transforms.py:151-224."""

    def __init__(self, Cl2m, Cm2l, eigenvalues, rossby_radii):
        self.Cl2m, self.Cm2l = Cl2m, Cm2l
        self.eigenvalues, self.rossby_radii = eigenvalues, rossby_radii

    @staticmethod
    def from_physics(H, g_prime, f0):
        A = build_coupling_matrix(H, g_prime)
        radii, Cl2m, Cm2l = decompose_vertical_modes(A, f0)
        return ModalTransform(Cl2m, Cm2l, reference_eigenvalues(A), radii)

    def to_modal(self, x):
        return np.einsum("lm,m...->l...", self.Cl2m.astype(x.dtype), x)

    def to_layer(self, x):
        return np.einsum("lm,m...->l...", self.Cm2l.astype(x.dtype), x)
