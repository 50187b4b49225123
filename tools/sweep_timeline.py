"""Debug build only (SOMAX_B200_NVCC_EXTRA=-DSB_TH_DEBUG): CTA start/end times of the sweep kernels."""
import ctypes as C, sys
sys.path.insert(0, ".")
import numpy as np, torch, bench
from somax_b200 import _lib
import somax_b200 as sb
kind, nl, nx, ny, members = bench.WORKLOADS["qg3_8192"]
model, st0, dt = bench.build_gpu_model(kind, nl, nx, ny, range(1))
lib = _lib.lib(); raw = C.CDLL(str(_lib.LIB_PATH))
q = torch.as_tensor(st0.q).cuda(); p = sb.models.qg._params_struct(model.params, model._H0)
stream = torch.cuda.current_stream().cuda_stream
h = model._engine.handle(1)
buf = (C.c_ulonglong * (4096 * 4))(); n = C.c_uint(0)
_lib.check(lib.somax_b200_qg_steps(h, q.data_ptr(), 2, dt, 0.0, C.byref(p), stream))
raw.somax_b200_debug_dump(buf, C.byref(n))
_lib.check(lib.somax_b200_qg_steps(h, q.data_ptr(), 1, dt, 0.0, C.byref(p), stream))
raw.somax_b200_debug_dump(buf, C.byref(n))
a = np.frombuffer(buf, dtype=np.uint64).reshape(-1, 4)[: min(n.value, 4096)]
t0 = a[:, 2].min()
first = a[(a[:, 2] - t0) < 4.5e6]
for code in sorted(set(first[:, 0].tolist()), key=lambda c: first[first[:, 0] == c][:, 2].min()):
    v = first[first[:, 0] == code]
    print("%04d n=%3d start %.3f..%.3f end %.3f..%.3f ms" % (code, len(v), (v[:, 2].min() - t0) / 1e6, (v[:, 2].max() - t0) / 1e6,
                                                        (v[:, 3].min() - t0) / 1e6, (v[:, 3].max() - t0) / 1e6))
