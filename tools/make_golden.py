"""Generate tests/golden/*.npz from the CPU oracle (fp64).

The reference cannot run in this image (no JAX), so these vectors freeze the ORACLE; the oracle's
conventions are pinned by the reference's own printed tutorial outputs
(tests/golden/reference_notebook_outputs.json, tests/test_oracle_reference_pins.py).
tools/capture_reference.py regenerates the same vectors from the real reference on a machine
with somax + JAX.
Run: PYTHONPATH=. python tools/make_golden.py
"""
from pathlib import Path

import numpy as np

from oracle import qg, swm
from oracle.testcases import baroclinic_instability_swm, synthetic_qg_state

out = Path(__file__).resolve().parents[1] / "tests" / "golden"
out.mkdir(parents=True, exist_ok=True)

m = qg.create_baroclinic(nx=32, ny=32, lateral_viscosity=15.0, bottom_drag=1e-7,
                         wind_amplitude=1.3e-10)
q0 = synthetic_qg_state(3, 32, 32, dtype=np.float64)
dt, n = 600.0, 20
q1 = m.integrate(q0, 0.0, n * dt, dt)
d = m.diagnose(q1)
np.savez_compressed(out / "qg3_32x32_f64.npz", q0=q0, q1=q1, psi0=m.invert_pv(q0), dq0=m.rhs(m.bc(q0)),
                    t1=n * dt, dt=dt, ke=d["kinetic_energy"], ens=d["enstrophy"],
                    Cl2m=m.Cl2m, Cm2l=m.Cm2l, lambdas=m.lambdas)

m = swm.create_multilayer(nx=32, ny=32, Lx=1e6, Ly=1e6, f0=1e-4, beta=1.6e-11, n_layers=2,
                          H=(500.0, 4500.0), g_prime=(9.81, 0.025), lateral_viscosity=100.0,
                          bottom_drag=1e-7, wind_amplitude=1e-6, bc="periodic")
_, (h0, u0, v0) = baroclinic_instability_swm(nx=32, ny=32, dtype=np.float64)
dt, n = 40.0, 20
h1, u1, v1 = m.integrate(h0, u0, v0, 0.0, n * dt, dt)
dh, du, dv = m.rhs(*m.bc(h0, u0, v0))
d = m.diagnose(h1, u1, v1)
np.savez_compressed(out / "swm2_32x32_f64.npz", h0=h0, u0=u0, v0=v0, h1=h1, u1=u1, v1=v1,
                    dh0=dh, du0=du, dv0=dv, t1=n * dt, dt=dt, energy=d["energy"],
                    ens=d["enstrophy"])
print("wrote", sorted(p.name for p in out.glob("*.npz")))
