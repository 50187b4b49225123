"""Small PV inversion + a few steps, for compute-sanitizer runs (tools/san_invert.py [nx ny])."""
import sys
import numpy as np
sys.path.insert(0, ".")
import somax_b200 as sb
from oracle.testcases import synthetic_qg_state
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 100
m = sb.BaroclinicQG.create(nx=nx, ny=ny, lateral_viscosity=15.0, bottom_drag=1e-7, wind_amplitude=1.3e-10)
q = synthetic_qg_state(3, nx, ny, dtype=np.float32)
psi = m._invert_pv(q)
print("psi", float(np.abs(psi).max()))
out = m.integrate(sb.BaroclinicQGState(q=q), 0.0, 1200.0, 600.0).ys.q[0]
print("q", float(np.abs(out).max()))
