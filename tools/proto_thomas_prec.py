import numpy as np, scipy.fft
from oracle.elliptic import helmholtz_dst
def thomas_mixed(d, a, F, store=np.float32, arith=np.float64):
    ny = F.shape[0]
    d = d.astype(arith); a = arith(a)
    cp = np.zeros(F.shape, arith); dp = np.zeros(F.shape, store)
    m = d.copy(); cp[0] = a/m; prev = (F[0].astype(arith)/m); dp[0] = prev.astype(store)
    for j in range(1, ny):
        m = d - a*cp[j-1]; cp[j] = a/m
        prev = (F[j].astype(arith) - a*prev)/m     # carry in arith precision
        dp[j] = prev.astype(store)
    x = np.zeros(F.shape, store)
    xc = dp[-1].astype(arith); x[-1] = xc.astype(store)
    for j in range(ny-2, -1, -1):
        xc = dp[j].astype(arith) - cp[j]*xc
        x[j] = xc.astype(store)
    return x
rng = np.random.default_rng(0)
for n in (256, 1024):
    ny = n; Lx = Ly = 4e6; dx, dy = Lx/n, Ly/ny
    r = np.zeros((1, ny+2, n+2)); 
    # smooth-ish rhs + noise
    r[0,1:-1,1:-1] = rng.standard_normal((ny,n))*1e-6
    for lam in (0.0, -5.66e-11):
        # pure 'DST-x (full n, scipy) + thomas-y' to isolate thomas precision
        k = np.arange(1, n+1); Lk = -(4/dx**2)*np.sin(np.pi*k/(2*(n+1)))**2
        d = Lk - 2/dy**2 - lam; a = 1/dy**2
        ref = helmholtz_dst(r, dx, dy, np.array([lam]))[0,1:-1,1:-1]
        fh = (scipy.fft.dst(r[0,1:-1,1:-1].astype(np.float32), type=1, axis=1)*np.float32(0.5))
        for store, arith in ((np.float32,np.float32),(np.float32,np.float64),(np.float64,np.float64)):
            vh = thomas_mixed(d, a, fh, store, arith)
            u = scipy.fft.dst(vh.astype(np.float32), type=1, axis=1)*np.float32(1.0/(n+1))
            print(n, lam, store.__name__, arith.__name__, np.linalg.norm(u-ref)/np.linalg.norm(ref))
