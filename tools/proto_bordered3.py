"""Prototype (numpy) of the GPU inversion for the reference's FULL-ARRAY DST solve
(oracle/elliptic.py): Ny x Nx unknowns, Nx = n + 2 with n = 2^p.

  x: array columns 1..n-1 form a DST-I(n-1) block (power-of-two FFT); columns 0, n, n+1 are
     border unknowns g0, g1, g2 (functions of y) handled by a Schur complement that is a
     3 x 3 system per y-mode (host-inverted);
  y: Thomas per x-wavenumber over all Ny rows.

The two weighted row sums the border system needs, v(1, j) and v(n-1, j), are the sum and the
difference of the odd-k and even-k partial sums (sin(pi k (n-1)/n) = (-1)^(k+1) sin(pi k/n));
the second solve's right-hand side is b sin(pi k/n) (g0 +- g1)(y) with the sign by parity of k.
Run: PYTHONPATH=. python tools/proto_bordered3.py
"""
import numpy as np
import scipy.fft

from oracle.elliptic import helmholtz_dst
from tools.proto_bordered import thomas


def dst1(x, axis=-1):
    return scipy.fft.dst(x, type=1, axis=axis) * 0.5      # sum_t x_t sin(pi k t / N)


def solve_bordered3(r, dx, dy, lam):
    Ny, Nx = r.shape
    n = Nx - 2
    b, a = 1.0 / dx ** 2, 1.0 / dy ** 2
    k = np.arange(1, n)
    Lk = -(4.0 * b) * np.sin(np.pi * k / (2.0 * n)) ** 2
    sg = np.sin(np.pi * k / n)                        # block column 1 weights
    par = np.where(k % 2 == 1, 1.0, -1.0)             # block column n-1 = par * sg
    d = Lk - 2.0 * a - lam
    fh = dst1(r[:, 1:n])
    vh = thomas(d, a, fh)
    w = (2.0 / n) * vh * sg[None]
    s_odd, s_even = w[:, k % 2 == 1].sum(axis=1), w[:, k % 2 == 0].sum(axis=1)
    v1, vn1 = s_odd + s_even, s_odd - s_even
    # Schur system per y-mode l
    l = np.arange(1, Ny + 1)
    mu = -(4.0 * a) * np.sin(np.pi * l / (2.0 * (Ny + 1))) ** 2 - lam
    den = Lk[None, :] + mu[:, None]
    alpha = (2.0 / n) * (sg[None] ** 2 / den).sum(axis=1)
    beta = (2.0 / n) * (par[None] * sg[None] ** 2 / den).sum(axis=1)
    M = np.zeros((Ny, 3, 3))
    M[:, 0, 0] = M[:, 1, 1] = mu - 2 * b - b * b * alpha
    M[:, 0, 1] = M[:, 1, 0] = -b * b * beta
    M[:, 1, 2] = M[:, 2, 1] = b
    M[:, 2, 2] = mu - 2 * b
    Minv = np.linalg.inv(M)
    R = np.stack([r[:, 0] - b * v1, r[:, n] - b * vn1, r[:, n + 1]], axis=1)     # (Ny, 3)
    Rh = dst1(R, axis=0)
    Gh = np.einsum("lij,lj->li", Minv, Rh)
    G = dst1(Gh, axis=0) * (2.0 / (Ny + 1))
    g0, g1, g2 = G[:, 0], G[:, 1], G[:, 2]
    gsel = np.where((k % 2 == 1)[None, :], (g0 + g1)[:, None], (g0 - g1)[:, None])
    wh = thomas(d, a, gsel.copy())
    uh = vh - b * sg[None] * wh
    u = np.zeros_like(r)
    u[:, 1:n] = dst1(uh) * (2.0 / n)
    u[:, 0], u[:, n], u[:, n + 1] = g0, g1, g2
    return u


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n, ny in [(16, 14), (64, 46), (256, 254), (512, 100)]:
        dx, dy = 4e6 / n, 4e6 / ny
        r = rng.standard_normal((ny + 2, n + 2)) * 1e-6
        for lam in [0.0, 5.6e-10, -5.66e-11]:
            ref = helmholtz_dst(r[None], dx, dy, np.array([lam]))[0]
            u = solve_bordered3(r, dx, dy, lam)
            print(n, ny, lam, "relL2 vs oracle:", np.linalg.norm(u - ref) / np.linalg.norm(ref))
