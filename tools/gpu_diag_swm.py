import numpy as np, sys
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
from test_gpu_parity import swm_pair, swm_state, rel
import somax_b200 as sb
for dtype in (np.float32, np.float64):
    for bc in ("periodic", "wall"):
        om, gm = swm_pair(64, 40, dtype, bc)
        h,u,v = swm_state(64, 40, dtype)
        st = sb.MultilayerSW2DState(h=h,u=u,v=v)
        t = gm.vector_field(0.0, st)
        f64 = [a.astype(np.float64) for a in (h,u,v)]
        r64 = om.rhs(*f64)
        r32 = om.rhs(h,u,v)
        print("rhs", dtype.__name__, bc, [f"{rel(a,r):.2e}" for a,r in zip((t.h,t.u,t.v), r64)],
              "oracle-same-dtype vs f64:", [f"{rel(a,r):.2e}" for a,r in zip(r32, r64)])
        for nx, steps in ((32,1),(32,100)):
            om, gm = swm_pair(nx, nx, dtype, bc)
            h,u,v = swm_state(nx, nx, dtype, noise=False)
            dt = 20.0*64/max(nx,64)
            sol = gm.integrate(sb.MultilayerSW2DState(h=h,u=u,v=v), 0.0, steps*dt, dt)
            ref = om.integrate(*[a.astype(np.float64) for a in (h,u,v)], 0.0, steps*dt, dt)
            same = om.integrate(h,u,v, 0.0, steps*dt, dt)
            print("int", dtype.__name__, bc, nx, steps, [f"{rel(getattr(sol.ys,n)[0], r):.2e}" for n,r in zip("huv",ref)],
                  "oracle-same-dtype vs f64:", [f"{rel(a,r):.2e}" for a,r in zip(same, ref)],
                  "max|v|", float(np.abs(ref[2]).max()))
