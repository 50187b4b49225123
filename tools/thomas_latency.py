"""Sweep kernels in the latency-bound regime of the slab model (a rank that owns a few strips: one
sweep CTA per SM): PV inversion of a narrow, tall grid (nx = 512, ny = 8192: 8 strips) on one GPU.
Prints the average time of every solver kernel; with a -DSB_TH_PHASES build of the library
(SOMAX_B200_LIB=path) also where a sweep CTA spends its cycles (wait for tiles / recurrence /
fence + barrier).  usage: thomas_latency.py [nx] [ny] [reps]"""
import ctypes as C
import json
import os
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
from somax_b200 import _lib
if os.environ.get("SOMAX_B200_LIB"):
    from pathlib import Path
    _lib.LIB_PATH = Path(os.environ["SOMAX_B200_LIB"])
import somax_b200 as sb
import bench

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
model = sb.BaroclinicQG.create(nx=nx, ny=ny, **bench.QG_PARAMS)
q = torch.as_tensor(sb.gfd_testcases.synthetic_qg_state(3, nx, ny, dtype="float32")).cuda()
lib = _lib.lib()
model._invert_pv(q)
lib.somax_b200_profile_reset(); lib.somax_b200_profile_enable(1)
for _ in range(reps):
    model._invert_pv(q)
lib.somax_b200_profile_enable(0)
buf = C.create_string_buffer(1 << 16)
_lib.check(lib.somax_b200_profile_report(buf, len(buf)))
for r in sorted(json.loads(buf.value.decode()), key=lambda r: -r["total_ms"]):
    print(f"{r['kernel']:24s} {r['total_ms'] / r['launches'] * 1e3:9.1f} us")
raw = C.CDLL(str(_lib.LIB_PATH))
if hasattr(raw, "somax_b200_debug_dump"):
    out = (C.c_ulonglong * (4096 * 4))(); n = C.c_uint(0)
    raw.somax_b200_debug_dump(out, C.byref(n))
    model._invert_pv(q)
    raw.somax_b200_debug_dump(out, C.byref(n))
    a = np.frombuffer(out, dtype=np.uint64).reshape(-1, 4)[: n.value]
    print("code(SUBST,FROM_VEC,KIND,TAB)half  total  wait  compute  fence+sync   [kcycles, thread 0 of strip 0 plane 0]")
    for row in a:
        code, tot = int(row[0]) & 0xffffffff, int(row[0]) >> 32
        print(f"{code:06d} {tot / 1e3:8.1f} {int(row[1]) / 1e3:8.1f} {int(row[2]) / 1e3:8.1f} {int(row[3]) / 1e3:8.1f}")
