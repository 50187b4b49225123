"""Tiny driver for ncu: N Tsit5 steps of a bench workload (no timing, no CPU baseline)."""
import ctypes as C
import sys
sys.path.insert(0, ".")
import torch
import bench
from somax_b200 import _lib
import somax_b200 as sb

wl = sys.argv[1] if len(sys.argv) > 1 else "qg3_8192"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
kind, nl, nx, ny, members = bench.WORKLOADS[wl]
model, st0, dt = bench.build_gpu_model(kind, nl, nx, ny, range(members))
nb = members
lib = _lib.lib()
stream = torch.cuda.current_stream().cuda_stream
if kind == "qg":
    q = torch.as_tensor(st0.q).cuda()
    p = sb.models.qg._params_struct(model.params, model._H0)
    _lib.check(lib.somax_b200_qg_steps(model._engine.handle(nb), q.data_ptr(), steps, dt, 0.0, C.byref(p), stream))
else:
    h, u, v = (torch.as_tensor(getattr(st0, f)).cuda() for f in "huv")
    p = model._pstruct()
    _lib.check(lib.somax_b200_swm_steps(model._handle(nb), h.data_ptr(), u.data_ptr(), v.data_ptr(), steps, dt, 0.0, C.byref(p), stream))
torch.cuda.synchronize()
print("done", wl, steps)
