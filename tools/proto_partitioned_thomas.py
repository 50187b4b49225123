"""CPU prototype (numpy / scipy, no GPU) of the partitioned Thomas solve for the y-slab model
("next"): every rank solves its rows of a column with homogeneous Dirichlet
ends, then corrects with two tabulated homogeneous solutions times the neighbours' interface
values, which come from a 2P x 2P reduced system per column.  Measures, on the columns of the
3 x 8192^2 double-gyre configuration, the error against the unpartitioned solve - including the
low-k (ill-conditioned) and the indefinite columns that the sweep kernels treat separately.

  normalised column system:  x_{j-1} + delta x_j + x_{j+1} = d_j,  j = 0..ny-1, x_{-1} = x_{ny} = 0
  slab p (rows r0..r1-1):     T_p x_p = d_p - e_first x_{r0-1} - e_last x_{r1}
  =>  x_p = xloc_p - x_{r0-1} u_p - x_{r1} v_p,   u_p = T_p^-1 e_first,  v_p = T_p^-1 e_last
"""
import sys

import numpy as np
from scipy.linalg import solve_banded

sys.path.insert(0, ".")
from somax_b200.core import ModalTransform  # noqa: E402  (setup-time host maths, no GPU)


def tri_solve(delta, d, dtype=np.float64):
    n = d.shape[0]
    ab = np.ones((3, n), dtype)
    ab[1] = delta
    return solve_banded((1, 1), ab, d.astype(dtype))


def partitioned(delta, d, P, dtype=np.float64):
    ny = d.shape[0]
    nl = ny // P
    xloc, U, V = [], [], []
    e0, e1 = np.zeros(nl), np.zeros(nl)
    e0[0] = e1[-1] = 1.0
    u = tri_solve(delta, e0)            # tables: fp64, computed once per column
    v = u[::-1]
    for p in range(P):
        xloc.append(tri_solve(delta, d[p * nl:(p + 1) * nl], dtype).astype(np.float64))
        U.append(u)
        V.append(v)
    # unknowns: z[2p] = first row of slab p, z[2p+1] = last row of slab p
    A = np.eye(2 * P)
    b = np.zeros(2 * P)
    for p in range(P):
        for row, k in ((0, 2 * p), (nl - 1, 2 * p + 1)):
            b[k] = xloc[p][row]
            if p > 0:
                A[k, 2 * (p - 1) + 1] += U[p][row]        # x_{r0-1} = last row of slab p-1
            if p < P - 1:
                A[k, 2 * (p + 1)] += V[p][row]            # x_{r1} = first row of slab p+1
    z = np.linalg.solve(A, b)
    out = np.empty(ny)
    for p in range(P):
        below = z[2 * (p - 1) + 1] if p > 0 else 0.0
        above = z[2 * (p + 1)] if p < P - 1 else 0.0
        out[p * nl:(p + 1) * nl] = xloc[p] - below * U[p] - above * V[p]
    return out, np.linalg.cond(A)


def main():
    n = 8192
    Lx = 4e6
    dx = dy = Lx / n
    f0 = 9.375e-5
    modal = ModalTransform.from_physics((400.0, 1100.0, 2600.0), (9.81, 0.025, 0.0125), f0)
    lambdas = f0 ** 2 * modal.eigenvalues
    rng = np.random.default_rng(0)
    d = rng.standard_normal(n) * dy * dy
    print("mode  column  delta+2        class       P   rel.err fp64   rel.err fp32-local   cond(reduced)")
    for m, lam in enumerate(lambdas):
        for c in (0, 1, 7, 63, 319, 320, 1023, 4095, 8190):
            sn = np.sin(np.pi * (c + 1) / (2.0 * n))
            lam_x = -(4.0 / dx ** 2) * sn * sn
            eps = (lam - lam_x) * dy * dy
            delta = -2.0 - eps
            cls = "indefinite" if eps <= 0 else ("low-k" if c < 320 else "plain")
            ref = tri_solve(delta, d)
            for P in (8,):
                x64, cond = partitioned(delta, d, P)
                x32, _ = partitioned(delta, d, P, np.float32)
                r32 = tri_solve(delta, d, np.float32)
                e64 = np.linalg.norm(x64 - ref) / np.linalg.norm(ref)
                e32 = np.linalg.norm(x32 - ref) / np.linalg.norm(ref)
                b32 = np.linalg.norm(r32 - ref) / np.linalg.norm(ref)
                print(f"{m:4d} {c:7d}  {-eps:+.3e}  {cls:10s} {P:3d}   {e64:.2e}       {e32:.2e} (unpartitioned fp32: {b32:.2e})   {cond:.2e}")


if __name__ == "__main__":
    main()
