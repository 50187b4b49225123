"""Tiny driver for ncu: one Tsit5 step of a QG grid split in P y-slabs held in ONE process on one
GPU (`SlabQG(local=True)`): the slab kernels (`seg_copy_kernel` transposes / halo pushes) with the
peer stores landing in the same device.  usage: prof_slab_local.py [n] [P] [steps]"""
import sys
sys.path.insert(0, ".")
import torch
import bench
import somax_b200 as sb
from somax_b200 import gfd_testcases as g
from somax_b200.parallel import SlabQG, split_slabs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
P = int(sys.argv[2]) if len(sys.argv) > 2 else 4
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
model = sb.BaroclinicQG.create(nx=n, ny=n, **bench.QG_PARAMS)
sl = SlabQG(model, P, local=True)
q = torch.as_tensor(g.synthetic_qg_state(3, n, n, dtype="float32")).cuda()
slabs = [s.contiguous().clone() for s in split_slabs(q, P)]
sl._steps(slabs, steps, bench.qg_dt(n), 0.0)
sl.check_peers()
print("done", n, P, steps)
