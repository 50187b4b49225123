"""Regenerate the golden vectors from the REAL reference (needs jax + somax installed).

Run on a machine with JAX:  python tools/capture_reference.py out_dir
It writes the same keys as tools/make_golden.py (which freezes the numpy oracle); comparing the
two files audits the oracle restatement of finitevolx / spectraldiffx / diffrax.
This script cannot run in the graft image (no jax wheels, no network).
"""
import sys
from pathlib import Path

import numpy as np


def main(out_dir):
    import jax
    jax.config.update("jax_enable_x64", True)
    import jax.numpy as jnp
    from somax._src.models.qg.baroclinic import BaroclinicQG, BaroclinicQGState
    from somax._src.models.swm.multilayer import MultilayerShallowWater2D, MultilayerSW2DState

    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    from oracle.testcases import baroclinic_instability_swm, synthetic_qg_state

    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    m = BaroclinicQG.create(nx=32, ny=32, lateral_viscosity=15.0, bottom_drag=1e-7,
                            wind_amplitude=1.3e-10)
    q0 = jnp.asarray(synthetic_qg_state(3, 32, 32, dtype=np.float64))
    dt, n = 600.0, 20
    sol = m.integrate(BaroclinicQGState(q=q0), 0.0, n * dt, dt)
    d = m.diagnose(BaroclinicQGState(q=sol.ys.q[0]))
    st0 = m.apply_boundary_conditions(BaroclinicQGState(q=q0))
    np.savez_compressed(out / "qg3_32x32_f64.npz", q0=np.asarray(q0), q1=np.asarray(sol.ys.q[0]),
                        psi0=np.asarray(m._invert_pv(q0)), dq0=np.asarray(m.vector_field(0.0, st0).q),
                        t1=n * dt, dt=dt, ke=np.asarray(d.kinetic_energy), ens=np.asarray(d.enstrophy),
                        Cl2m=np.asarray(m.modal.Cl2m), Cm2l=np.asarray(m.modal.Cm2l),
                        lambdas=np.asarray(m.helmholtz_lambdas))
    sm = MultilayerShallowWater2D.create(nx=32, ny=32, Lx=1e6, Ly=1e6, f0=1e-4, beta=1.6e-11,
                                         n_layers=2, H=(500.0, 4500.0), g_prime=(9.81, 0.025),
                                         lateral_viscosity=100.0, bottom_drag=1e-7,
                                         wind_amplitude=1e-6, bc="periodic")
    _, (h0, u0, v0) = baroclinic_instability_swm(nx=32, ny=32, dtype=np.float64)
    s0 = MultilayerSW2DState(h=jnp.asarray(h0), u=jnp.asarray(u0), v=jnp.asarray(v0))
    dt, n = 40.0, 20
    sol = sm.integrate(s0, 0.0, n * dt, dt)
    t = sm.vector_field(0.0, sm.apply_boundary_conditions(s0))
    last = MultilayerSW2DState(h=sol.ys.h[0], u=sol.ys.u[0], v=sol.ys.v[0])
    d = sm.diagnose(last)
    np.savez_compressed(out / "swm2_32x32_f64.npz", h0=h0, u0=u0, v0=v0, h1=np.asarray(last.h),
                        u1=np.asarray(last.u), v1=np.asarray(last.v), dh0=np.asarray(t.h),
                        du0=np.asarray(t.u), dv0=np.asarray(t.v), t1=n * dt, dt=dt,
                        energy=np.asarray(d.energy), ens=np.asarray(d.enstrophy))
    print("wrote", sorted(p.name for p in out.glob("*.npz")))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "reference_golden")
