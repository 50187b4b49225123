"""CPU prototype (numpy) of the PV inversion on y-slabs WITHOUT transposes: the bordered
DST-I(n-1) + Thomas algorithm of tools/proto_bordered.py with both y-solves done as partitioned
Thomas solves ("next").  Rank p keeps its rows for every x-wavenumber;
what crosses ranks per inversion is, per plane: 2 interface rows of each rank for each of the two
solves (all-gather of 2 (n-1) values per rank), the border right-hand side (all-gather of ny/P
values per rank) - O(n P) values instead of two transposes of the whole spectral array.

  python tools/proto_slab_partitioned.py        (checks against oracle.elliptic.helmholtz_dst)
"""
import sys

import numpy as np
import scipy.fft

sys.path.insert(0, ".")
from oracle.elliptic import helmholtz_dst  # noqa: E402
from tools.proto_bordered import thomas  # noqa: E402


def homogeneous(d, a, nl):
    """U[:, k] = T^-1 e_first (scaled by a) for the nl-row local system of every column."""
    E = np.zeros((nl, d.shape[0]))
    E[0] = a
    U = thomas(d, a, E)
    return U, U[::-1]


def thomas_partitioned(d, a, F, P, tables=None):
    """tridiag(a, d_k, a) x = F along axis 0, rows split over P ranks.  Returns (x, words)."""
    ny, n = F.shape
    nl = ny // P
    U, V = tables if tables is not None else homogeneous(d, a, nl)
    xloc = [thomas(d, a, F[p * nl:(p + 1) * nl].copy()) for p in range(P)]          # local, no communication
    # all-gather of the first / last local row of every rank: 2 n words per rank
    first = np.stack([x[0] for x in xloc])
    last = np.stack([x[-1] for x in xloc])
    # reduced system, per column, for z = (first_0, last_0, first_1, last_1, ...): solved redundantly by every rank
    A = np.zeros((n, 2 * P, 2 * P))
    b = np.zeros((n, 2 * P))
    A[:, np.arange(2 * P), np.arange(2 * P)] = 1.0
    for p in range(P):
        for row, k in ((0, 2 * p), (nl - 1, 2 * p + 1)):
            b[:, k] = (first if row == 0 else last)[p]
            if p > 0:
                A[:, k, 2 * (p - 1) + 1] += U[row]
            if p < P - 1:
                A[:, k, 2 * (p + 1)] += V[row]
    z = np.linalg.solve(A, b[..., None])[..., 0]
    x = np.empty_like(F)
    for p in range(P):
        below = z[:, 2 * (p - 1) + 1] if p > 0 else 0.0
        above = z[:, 2 * (p + 1)] if p < P - 1 else 0.0
        x[p * nl:(p + 1) * nl] = xloc[p] - U * below - V * above                       # local correction pass
    return x, 2 * n * P


def solve_slabs(r, dx, dy, lam, P):
    ny, n = r.shape
    b, a = 1.0 / dx ** 2, 1.0 / dy ** 2
    k = np.arange(1, n)
    Lk = -(4.0 / dx ** 2) * np.sin(np.pi * k / (2.0 * n)) ** 2
    sig = ((-1.0) ** (k + 1)) * np.sin(np.pi * k / n)
    d = Lk - 2.0 / dy ** 2 - lam
    tables = homogeneous(d, a, ny // P)                     # set-up time, per (mode, column): one slab-sized table
    fh = scipy.fft.dst(r[:, :n - 1], type=1, axis=1) * 0.5  # rows: local to the slab
    vh, w1 = thomas_partitioned(d, a, fh, P, tables)
    vn1 = (2.0 / n) * (vh * sig[None]).sum(axis=1)          # row sums: local rows only
    l = np.arange(1, ny + 1)
    mu = -(4.0 / dy ** 2) * np.sin(np.pi * l / (2.0 * (ny + 1))) ** 2 - lam
    s = mu - 2.0 / dx ** 2 - (1.0 / dx ** 4) * (2.0 / n) * (sig[None, :] ** 2 / (Lk[None, :] + mu[:, None])).sum(axis=1)
    rhs_g = r[:, n - 1] - b * vn1                            # all-gather of ny / P values per rank
    g = scipy.fft.dst(scipy.fft.dst(rhs_g, type=1) * 0.5 / s, type=1) * 0.5 * (2.0 / (ny + 1))
    wh, w2 = thomas_partitioned(d, a, np.broadcast_to(g[:, None], (ny, n - 1)).copy(), P, tables)
    uh = vh - b * sig[None] * wh
    u = np.zeros((ny, n))
    u[:, :n - 1] = scipy.fft.dst(uh, type=1, axis=1) * 0.5 * (2.0 / n)
    u[:, n - 1] = g
    return u, w1 + w2 + ny


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    print("   n   ny   P   lambda        rel-L2 vs oracle   words exchanged / plane   (two transposes would move)")
    for n, ny, P in [(64, 48, 2), (256, 256, 4), (512, 512, 8), (1024, 2048, 8)]:
        dx, dy = 4e6 / n, 4e6 / ny
        r = np.zeros((1, ny + 2, n + 2))
        r[0, 1:-1, 1:-1] = rng.standard_normal((ny, n)) * 1e-6
        for lam in (0.0, 5.6e-10, -5.66e-11):
            ref = helmholtz_dst(r, dx, dy, np.array([lam]))[0, 1:-1, 1:-1]
            u, words = solve_slabs(r[0, 1:-1, 1:-1], dx, dy, lam, P)
            err = np.linalg.norm(u - ref) / np.linalg.norm(ref)
            print(f"{n:5d} {ny:5d} {P:3d}  {lam:+.2e}    {err:.2e}          {words:10d}               {2 * n * ny * (P - 1) // P:12d}")
