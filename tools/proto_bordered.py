"""Prototype (numpy) of the GPU algorithm: DST-I(n-1) in x (n = 2^p), bordered last column,
Thomas in y.  Used to validate the maths before writing CUDA."""
import numpy as np, scipy.fft, sys
from oracle.elliptic import helmholtz_dst

def thomas(d, a, F):
    """solve tridiag(a, d_k, a) along axis 0 for each column k; d: (n,), F: (ny, n)"""
    ny = F.shape[0]
    dt = F.dtype
    cp = np.zeros_like(F); dp = np.zeros_like(F)
    m = d.astype(dt).copy()
    cp[0] = a / m; dp[0] = F[0] / m
    for j in range(1, ny):
        m = d - a * cp[j-1]
        cp[j] = a / m
        dp[j] = (F[j] - a * dp[j-1]) / m
    x = np.zeros_like(F)
    x[-1] = dp[-1]
    for j in range(ny-2, -1, -1):
        x[j] = dp[j] - cp[j] * x[j+1]
    return x

def solve_bordered(r, dx, dy, lam, dt=np.float64):
    ny, n = r.shape
    r = r.astype(dt)
    b = dt(1.0/dx**2); a = dt(1.0/dy**2)
    k = np.arange(1, n)
    Lk = (-(4.0/dx**2)*np.sin(np.pi*k/(2.0*n))**2)
    sig = ((-1.0)**(k+1))*np.sin(np.pi*k/n)
    fh = scipy.fft.dst(r[:, :n-1], type=1, axis=1) * dt(0.5)   # sum x sin
    d = (Lk - 2.0/dy**2 - lam).astype(dt)
    vh = thomas(d, a, fh.astype(dt))
    # u(n-1, j)
    vn1 = (2.0/n) * (vh * sig.astype(dt)[None]).sum(axis=1)
    # border solve in y-spectral space (float64 host-side diag)
    l = np.arange(1, ny+1)
    mu = -(4.0/dy**2)*np.sin(np.pi*l/(2.0*(ny+1)))**2 - lam
    s = mu - 2.0/dx**2 - (1.0/dx**4)*(2.0/n)*(sig[None,:]**2/(Lk[None,:] + mu[:,None])).sum(axis=1)
    rhs_g = r[:, n-1] - b*vn1
    gh = scipy.fft.dst(rhs_g.astype(np.float64), type=1) * 0.5
    g = (scipy.fft.dst(gh / s, type=1) * 0.5 * (2.0/(ny+1))).astype(dt)
    wh = thomas(d, a, np.broadcast_to(g[:, None], (ny, n-1)).astype(dt).copy())
    uh = vh - b*sig.astype(dt)[None]*wh
    u = np.zeros((ny, n), dt)
    u[:, :n-1] = scipy.fft.dst(uh, type=1, axis=1) * dt(0.5) * dt(2.0/n)
    u[:, n-1] = g
    return u

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n, ny in [(16, 16), (64, 48), (256, 256)]:
        Lx = Ly = 4e6
        dx, dy = Lx/n, Ly/ny
        r = np.zeros((1, ny+2, n+2)); r[0, 1:-1, 1:-1] = rng.standard_normal((ny, n))*1e-6
        for lam in [0.0, 5.6e-10, -5.66e-11]:
            ref = helmholtz_dst(r, dx, dy, np.array([lam]))[0, 1:-1, 1:-1]
            for dt in (np.float64, np.float32):
                u = solve_bordered(r[0, 1:-1, 1:-1], dx, dy, lam, dt)
                err = np.linalg.norm(u - ref)/np.linalg.norm(ref)
                ref32 = helmholtz_dst(r.astype(np.float32), dx, dy, np.array([lam]))[0,1:-1,1:-1]
                e32 = np.linalg.norm(ref32 - ref)/np.linalg.norm(ref)
                print(n, ny, lam, dt.__name__, "relL2 bordered:", err, " oracle-fp32:", e32)
