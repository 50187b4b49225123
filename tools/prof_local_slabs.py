"""Per-kernel times of the slab model with all P slabs on ONE GPU (validation mode): the kernels
and launch parameters of a P-GPU run, minus the NVLink.  python tools/prof_local_slabs.py [nx] [P] [steps]"""
import ctypes as C
import json
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
import somax_b200 as sb
from somax_b200 import _lib
from somax_b200.parallel import SlabQG

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
P = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
m = sb.BaroclinicQG.create(nx=nx, ny=nx, dtype="float32", lateral_viscosity=15.0, bottom_drag=1e-7, wind_amplitude=1.3e-10)
g = torch.Generator(device="cuda").manual_seed(3)
q = 1e-6 * torch.randn((3, nx + 2, nx + 2), generator=g, device="cuda", dtype=torch.float32)
dt = 0.25 * m.grid.dx / 2.0
sl = SlabQG(m, P, local=True)
lib = _lib.lib()
sl.integrate(q, 0.0, 2 * dt, dt)                       # warm-up
lib.somax_b200_profile_reset(); lib.somax_b200_profile_enable(1)
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(); sl.integrate(q, 0.0, steps * dt, dt); t1.record(); torch.cuda.synchronize()
lib.somax_b200_profile_enable(0)
buf = C.create_string_buffer(1 << 20)
_lib.check(lib.somax_b200_profile_report(buf, len(buf)))
rep = json.loads(buf.value.decode())
print("ms/step (profiled, P slabs on one GPU):", t0.elapsed_time(t1) / steps)
for k in sorted(rep if isinstance(rep, list) else rep.get("kernels", []), key=lambda k: -k["total_ms"])[:14]:
    print(f"  {k['kernel']:22s} n={k['launches']:5d} per-launch {k['total_ms'] / max(k['launches'], 1):.4f} ms")
sl.close()
