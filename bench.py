#!/usr/bin/env python
"""Headline benchmark: Tsit5 stepping of the 3-layer quasi-geostrophic double gyre on B200.

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

A "step" is one full Tsit5 step (6 fresh RHS evaluations: PV inversion + fused stencil/RK
epilogue each) of the whole grid.  One cell-update = one interior cell of one layer advanced by
one step.  N = 1 runs BASELINE.json's target configuration, the 3-layer QG double gyre at 8192^2
in fp32 (it fits one B200).  For N > 1 the same ONE 8192^2 grid is partitioned in N y-slabs
(BASELINE config 4, "strong" scaling; halo rows and the transposes of the distributed DST are
peer-memory stores over NVLink); `--decomp members` instead gives every rank its own grid / shards
an ensemble by member (no data-path collective; the default for the ensemble and shallow-water
workloads) - see DESIGN.md section "multi-GPU".

Rank 0 prints ONE JSON line (keys: see the graft bench contract; `roofline` is for the kernel
with the largest share of the step, `cpu_baseline` times the numpy/scipy oracle port on the
host cores, `e2e` goes through the public `model.integrate` API with host numpy buffers).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (kind, nl, nx, ny, ensemble members)
    "qg3_8192": ("qg", 3, 8192, 8192, 1),        # the headline: BASELINE target configuration
    "qg3_4096": ("qg", 3, 4096, 4096, 1),
    "qg3_1024": ("qg", 3, 1024, 1024, 1),
    "qg3_128": ("qg", 3, 128, 128, 1),           # BASELINE config 2 (launch-latency bound)
    "qg1_64": ("qgbt", 1, 64, 64, 1),            # BASELINE config 1: barotropic double gyre (sim-bt-qg)
    "qg3_8192_f64": ("qg", 3, 8192, 8192, 1),    # the headline grid in fp64 (north-star 1e-12 pipeline)
    "qg3_256_ens1024": ("qg", 3, 256, 256, 1024),  # BASELINE config 5: 1024 members x 3 x 256^2
    "swm2_4096": ("swm", 2, 4096, 4096, 1),      # BASELINE config 3 at its largest size
    "swm2_1024": ("swm", 2, 1024, 1024, 1),
    "swm2_64": ("swm", 2, 64, 64, 1),
}
QG_PARAMS = dict(Lx=4e6, Ly=4e6, f0=9.375e-5, beta=1.754e-11, n_layers=3,
                 H=(400.0, 1100.0, 2600.0), g_prime=(9.81, 0.025, 0.0125), lateral_viscosity=15.0,
                 bottom_drag=1e-7, wind_amplitude=1.3e-10)           # configs/_authoring/doublegyre_bc_qg.py:26-36
SWM_PARAMS = dict(Lx=1e6, Ly=1e6, f0=1e-4, beta=1.6e-11, H=(500.0, 4500.0), g_prime=(9.81, 0.025),
                  lateral_viscosity=100.0, bottom_drag=1e-7)        # configs/_authoring/swm_jet.py:21-40
# algorithmic state-sized transfers per Tsit5 step
TRANSFERS = {"qg": 79, "qgbt": 79, "swm": 111}
DTYPES = {"qg3_8192_f64": "float64"}             # every other workload runs the reference's default, fp32
BT_PARAMS = dict(Lx=1e6, Ly=1e6, f0=1e-4, beta=1.6e-11, lateral_viscosity=500.0, bottom_drag=1e-7,
                 wind_amplitude=1e-12)                               # configs/_authoring/doublegyre_bt_qg.py:22-29
# algorithmic transfers of one launch of each kernel (DESIGN.md "kernels"): arrays it must read+write
KERNEL_TRANSFERS = {
    "rowdst_fwd_fft": 2.0, "rowdst_inv_fft": 2.0, "thomas_fwd_0": 2.0, "thomas_bwd_0": 1.0,
    "thomas_fwd_1": 1.0, "thomas_bwd_1": 3.0,
    "qg_rhs_kernel": (37.0 + 6.0) / 6.0, "swm_rhs_kernel": 111.0 / 6.0 / 3.0 * 3.0,
}


def ncu_traffic(workload, tag):
    """Measured DRAM bytes per launch of kernel `tag` (dram__bytes_read + write, averaged over the
    launches of one `ncu --set full` capture summarised in profiles/); None if not captured."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_{workload}_ncu_full*.csv")))
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    best = []
    for fn in files:       # the capture holding the most launches of this kernel (all stages of a step)
        try:
            rows = list(csv.reader(open(fn)))
            hdr = rows[0]
            cols = [(i, unit[h.split("[")[1].rstrip("]")]) for i, h in enumerate(hdr) if h.startswith("dram__bytes_")]
            vals = [sum(float(r[i]) * u for i, u in cols) for r in rows[1:] if tag in r[0]]
        except Exception:
            continue
        if len(vals) >= len(best):
            best = vals
    return sum(best) / len(best) if best else None


def qg_dt(nx):
    return 600.0 * (128.0 / nx) if nx >= 128 else 600.0


def swm_dt(nx):
    return 20.0 * (64.0 / nx)


class ClockSampler(threading.Thread):
    """SM clock and clock-event (throttle) reasons sampled during the timed region: NVML every
    20 ms when `pynvml` is importable (a 10-step timed region lasts ~0.25 s), else `nvidia-smi`
    every 200 ms."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self._halt = index, threading.Event()
        self.sm, self.mx, self.reasons, self.source = [], [], set(), "nvidia-smi"
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.source = "nvml"
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)))
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        mask = int(get(self.handle))
        for name, bit in self.BITS.items():
            if mask & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                              "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
        for line in out.strip().splitlines():
            r = [c.strip() for c in line.split(",")]
            if len(r) >= 8 and r[1].replace(".", "").isdigit():
                self.sm.append(float(r[1]))
                self.mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)

    def run(self):
        while not self._halt.is_set():
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._halt.wait(0.02 if self.nvml is not None else 0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None,
                "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}


def oracle_model(kind, nl, nx, ny, workers):
    from oracle import qg as oqg
    from oracle import testcases as ot
    if kind in ("qg", "qgbt"):
        m = (oqg.create_baroclinic(nx=nx, ny=ny, **QG_PARAMS) if kind == "qg"
             else oqg.create_barotropic(nx=nx, ny=ny, **BT_PARAMS))
        m.workers = workers
        q0 = ot.synthetic_qg_state(nl, nx, ny, dtype=np.float32)
        return m, (q0,), qg_dt(nx)
    m, st = ot.baroclinic_instability_swm(nx=nx, ny=ny, **SWM_PARAMS)
    return m, st, swm_dt(nx)


def time_oracle(kind, nl, nx, ny, steps, warmup, workers):
    """Oracle port (numpy/scipy) on the host cores.  Returns Gcell-steps/s."""
    m, st, dt = oracle_model(kind, nl, nx, ny, workers)
    if warmup:
        m.integrate(*st, 0.0, warmup * dt, dt)
    t = time.perf_counter()
    m.integrate(*st, 0.0, steps * dt, dt)
    el = time.perf_counter() - t
    return nl * nx * ny * steps / el / 1e9, el


def run_reference(args):
    """--impl reference: the reference's algorithm on the host CPU.  The reference itself (JAX)
    cannot be installed in this image (no jax/jaxlib wheels), so this is the
    oracle port - numpy/scipy with all host threads - on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, nl, nx, ny, _members = WORKLOADS[args.workload]
    sn = min(nx, 1024)   # bounded sample: same model and parameters on a 1024^2 grid
    cores = os.cpu_count()
    K, W = max(1, min(args.steps, 8)), min(args.warmup, 1)
    v, el = time_oracle(kind, nl, sn, sn, K, W, cores)
    sample = (f"{K} Tsit5 steps of the {nl}-layer {kind.upper()} at {sn}^2 fp32 (per-cell cost of the "
              f"workload's {nx}^2 grid is >= this: FFT work grows as n log n)")
    line = {
        "impl": "reference", "metric": "cell_updates_per_s", "value": v, "unit": "Gcell-steps/s",
        "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": el / K * 1e3,
        "higher_is_better": True, "scaling": scaling_label(args.workload, args.decomp),
        "vs_baseline": None, "dtype": "f64" if DTYPES.get(args.workload) == "float64" else "f32",
        "data": "synthetic", "config": workload_config(args.workload, sn),
        "cpu_baseline": {"value": v, "unit": "Gcell-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Gcell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def scaling_label(workload, decomp):
    """'strong' when the total work is the same at every N (one grid in slabs; a fixed member set),
    'weak' when every GPU gets its own grid - the same label at N = 1 and N > 1."""
    kind, _nl, _nx, _ny, members = WORKLOADS[workload]
    return "strong" if (members > 1 or (kind in ("qg", "swm") and decomp == "slab")) else "weak"


def workload_config(workload, sample_grid):
    """The `config` object both arms print: the workload by name plus the grid of the bounded
    sample the CPU arm / cpu_baseline is timed on."""
    kind, nl, nx, ny, members = WORKLOADS[workload]
    return {"workload": workload, "model": f"{nl}-layer {kind}", "grid": [ny, nx], "members": members,
            "dt": qg_dt(nx) if kind != "swm" else swm_dt(nx), "cpu_sample_grid": [sample_grid, sample_grid]}


def build_gpu_model(kind, nl, nx, ny, members=range(1), dtype="float32"):
    """Model + initial state; for ensembles the state gets a leading member axis and member e
    uses seed 10_000 + e."""
    import somax_b200 as sb
    from somax_b200 import gfd_testcases as g
    members = list(members)
    if kind == "qgbt":
        model = sb.BarotropicQG.create(nx=nx, ny=ny, dtype=dtype, **BT_PARAMS)
        return model, sb.BarotropicQGState(q=g.synthetic_qg_state(1, nx, ny, dtype=dtype)[0]), qg_dt(nx)
    if kind == "qg":
        model = sb.BaroclinicQG.create(nx=nx, ny=ny, dtype=dtype, **QG_PARAMS)
        if len(members) == 1 and members[0] == 0:
            q0 = g.synthetic_qg_state(nl, nx, ny, dtype=dtype)
        else:
            q0 = np.stack([g.synthetic_qg_state(nl, nx, ny, seed=10_000 + e, dtype=dtype) for e in members])
        return model, sb.BaroclinicQGState(q=q0), qg_dt(nx)
    model, st = g.baroclinic_instability_swm(nx=nx, ny=ny, **SWM_PARAMS)
    return model, st, swm_dt(nx)


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from somax_b200 import _lib
    import somax_b200 as sb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    quiet_stdout()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kind, nl, nx, ny, members = WORKLOADS[args.workload]
    K, W = args.steps, max(args.warmup, 3)
    lib = _lib.lib()
    from somax_b200.parallel import shard_members
    # ensembles: strong scaling over a fixed member set; single grids: one member per rank (weak)
    mine = shard_members(members, rank, world) if members > 1 else range(1)
    nb = len(mine)
    dtype = DTYPES.get(args.workload, "float32")
    model, st0, dt = build_gpu_model(kind, nl, nx, ny, mine, dtype)
    isqg = kind in ("qg", "qgbt")
    fields = [f for f in ("q", "h", "u", "v") if hasattr(st0, f)]
    dev = {f: torch.as_tensor(getattr(st0, f)).cuda() for f in fields}
    state_bytes = sum(t.numel() * t.element_size() for t in dev.values())
    handle = model._engine.handle(nb) if isqg else model._handle(nb)
    p = (sb.models.qg._params_struct(model.params, model._H0) if isqg else model._pstruct())
    stream = torch.cuda.current_stream().cuda_stream

    def steps_dev(n):
        if isqg:
            _lib.check(lib.somax_b200_qg_steps(handle, dev["q"].data_ptr(), n, dt, 0.0, C.byref(p), stream))
        else:
            _lib.check(lib.somax_b200_swm_steps(handle, dev["h"].data_ptr(), dev["u"].data_ptr(),
                                                dev["v"].data_ptr(), n, dt, 0.0, C.byref(p), stream))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (untimed) ----
    steps_dev(W)
    barrier()

    # ---- timed region: K steps, device resident, CUDA events on the launching stream ----
    # Timed CLEAN (no per-launch events); the per-kernel profile comes from a second, identical pass
    # of K steps with the library's per-launch CUDA events on (small grids replay a captured CUDA
    # graph in the timed pass and run eagerly in the profiled one).
    graphed = nb * nl * (ny + 2) * (nx + 2) <= (1 << 21) and K >= 9
    lib.somax_b200_profile_reset()
    lib.somax_b200_profile_enable(0)
    sampler = ClockSampler(local)
    sampler.start()
    n0 = lib.somax_b200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    steps_dev(K)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.somax_b200_launch_count() - n0
    clocks = sampler.stop()
    lib.somax_b200_profile_enable(1)
    steps_dev(K)
    barrier()
    lib.somax_b200_profile_enable(0)
    buf = C.create_string_buffer(1 << 16)
    _lib.check(lib.somax_b200_profile_report(buf, len(buf)))
    prof = json.loads(buf.value.decode())
    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    cells = nl * nx * ny * (members if members > 1 else world)   # whole job
    value = cells * K / (ms_max * 1e-3) / 1e9

    # ---- e2e: the public API, host buffers in -> host buffers out.  The state sits in pinned host
    # memory (allocated outside the timed region); one model.integrate() call copies it H2D, runs
    # K steps and reads the final state back D2H; the diagnostics scalars of the result (a few
    # doubles) are read back as well.  One untimed 1-step call warms the staging buffers up.
    host_state = type(st0)(**{f: torch.as_tensor(np.ascontiguousarray(getattr(st0, f))).pin_memory()
                              for f in fields})
    model.integrate(host_state, 0.0, dt, dt, max_steps=None)
    barrier()
    t0 = time.perf_counter()
    sol = model.integrate(host_state, 0.0, K * dt, dt, max_steps=None)
    out_state = type(st0)(**{f: getattr(sol.ys, f)[0] for f in fields})
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    nonfinite = float(sum(int((~torch.isfinite(getattr(out_state, f))).sum()) for f in fields))
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = cells * K / float(t_e.item()) / 1e9
    io = model.last_io

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
    w = 8 if dtype == "float64" else 4
    padded = nb * nl * (ny + 2) * (nx + 2) * w     # one state-sized array (one field) on this rank
    prof.sort(key=lambda r: -r["total_ms"])
    total_prof = sum(r["total_ms"] for r in prof) or 1.0
    top = prof[0]
    k_tr = KERNEL_TRANSFERS.get(top["kernel"], 2.0)
    per_launch_ms = top["total_ms"] / top["launches"]
    achieved = k_tr * padded / (per_launch_ms * 1e-3) / 1e9
    step_alg_bytes = TRANSFERS[kind] * padded
    step_frac = step_alg_bytes / (ms_max / K * 1e-3) / 1e9 / peak      # per GPU
    roofline = {
        "bound": "hbm", "kernel": top["kernel"], "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": ncu_traffic(args.workload, top["kernel"]), "peak_source": peak_src,
        "algorithmic_bytes_per_launch": k_tr * padded, "avg_launch_ms": per_launch_ms,
        "share_of_step": top["total_ms"] / total_prof,
        "events": "per-launch CUDA events in a second, identical pass of K steps (the timed region runs "
                  "without them" + ("; it replays a CUDA graph)" if graphed else ")"),
        "step": {"algorithmic_bytes": step_alg_bytes, "achieved_gbs": step_alg_bytes / (ms_max / K * 1e-3) / 1e9,
                 "frac": step_frac, "transfers_per_step": TRANSFERS[kind]},
        "kernels": [{"kernel": r["kernel"], "launches": r["launches"], "total_ms": round(r["total_ms"], 3),
                     "share": round(r["total_ms"] / total_prof, 4)} for r in prof],
    }

    # ---- CPU baseline: oracle port on the host cores, bounded sample ----
    cores = os.cpu_count()
    sn = min(nx, 1024)
    ksteps = 4
    cpu_v, cpu_el = time_oracle(kind, nl, sn, sn, ksteps, 0, cores)
    cpu_baseline = {"value": cpu_v, "unit": "Gcell-steps/s", "cores": cores, "kind": "port",
                    "sample": f"{ksteps} Tsit5 steps of the same model at {sn}^2 fp32, numpy/scipy oracle, "
                              f"scipy.fft workers={cores} ({cpu_el:.1f} s)"}

    line = {
        "metric": "cell_updates_per_s", "value": value, "unit": "Gcell-steps/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True,
        "scaling": scaling_label(args.workload, args.decomp), "vs_baseline": None,
        "dtype": "f64" if dtype == "float64" else "f32", "data": "synthetic",
        "config": {**workload_config(args.workload, sn), "l2": "working set (>= 7 GB) far larger than the 126 MB L2; no flush needed"
                   if state_bytes > 5e8 else "small working set: L2 resident by nature of the workload",
                   "parallelism": (f"{members} members sharded by member over {world} GPU(s)" if members > 1 else
                                   ("1 member per GPU (ensemble sharded by member)" if world > 1 else "single GPU")),
                   "solver": "fft + 3 border columns + thomas (whole-array DST-I solve)" if isqg else "n/a",
                   "device_bytes": int(lib.somax_b200_qg_device_bytes(handle) if isqg
                                       else lib.somax_b200_swm_device_bytes(handle))},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "Gcell-steps/s", "h2d_bytes_per_step": io.h2d_bytes / K,
                "d2h_bytes_per_step": io.d2h_bytes / K, "steps_per_call": K,
                "note": "one model.integrate() call of K steps on pinned host tensors: H2D of the state, "
                        "K steps, D2H of the final state", "nonfinite": nonfinite},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_gpu_slab(args):
    """--decomp slab: ONE grid (the workload's) partitioned in y-slabs over the N GPUs (BASELINE
    config 4; strong scaling).  Exchanges are peer-memory stores over NVLink inside the library
    (somax_b200_qgs_*); torch.distributed only all-gathers the IPC handles at set-up and reduces
    the timing."""
    import torch
    import torch.distributed as dist
    from somax_b200 import _lib
    import somax_b200 as sb
    from somax_b200.parallel import SlabQG, slab_window

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    quiet_stdout()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kind, nl, nx, ny, members = WORKLOADS[args.workload]
    if kind == "swm" and members == 1:
        return run_gpu_slab_swm(args, world, rank, local)
    if kind != "qg" or members != 1:
        raise SystemExit("--decomp slab is for single-grid workloads")
    K, W = args.steps, max(args.warmup, 3)
    lib = _lib.lib()
    model = sb.BaroclinicQG.create(nx=nx, ny=ny, **QG_PARAMS)
    dt = qg_dt(nx)
    slab_model = SlabQG(model, world, rank=rank)
    win = slab_window(ny, rank, world)
    from somax_b200 import gfd_testcases as g
    q0 = g.synthetic_qg_state(nl, nx, ny, dtype="float32")[:, win, :]
    host = torch.as_tensor(np.ascontiguousarray(q0)).pin_memory()
    dev = host.cuda()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    slab_model.advance_slab(dev, W, dt)
    barrier()
    # ---- correctness carried by the line: the same W steps on ONE device (the single-GPU model on the
    # whole grid, rank 0), compared on rank 0's window - outside the timed region.  The run fails above
    # the fp32 tolerance of BASELINE.json.
    rel = torch.zeros(1, dtype=torch.float64, device="cuda")
    if rank == 0:
        full0 = torch.as_tensor(g.synthetic_qg_state(nl, nx, ny, dtype="float32")).cuda()
        ref = model.integrate(sb.BaroclinicQGState(q=full0), 0.0, W * dt, dt, max_steps=None).ys.q[0][:, win, :]
        rel[0] = (torch.linalg.vector_norm((dev.double() - ref.double()).flatten()) /
                  torch.linalg.vector_norm(ref.double().flatten()))
        del full0, ref
        model._engine.close()
        torch.cuda.empty_cache()
    if world > 1:
        dist.broadcast(rel, 0)
    slab_rel = float(rel.item())
    if not (slab_rel <= 1e-5):
        sys.stderr.write(f"slab decomposition disagrees with the single-GPU model: relL2 = {slab_rel}\n")
        raise SystemExit(3)
    barrier()
    lib.somax_b200_profile_reset()
    lib.somax_b200_profile_enable(0)
    sampler = ClockSampler(local)
    sampler.start()
    n0 = lib.somax_b200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    slab_model.advance_slab(dev, K, dt, check=False)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.somax_b200_launch_count() - n0
    clocks = sampler.stop()
    slab_model.check_peers()
    # per-kernel profile: a second, identical pass of K steps with per-launch events on
    lib.somax_b200_profile_enable(1)
    slab_model.advance_slab(dev, K, dt, check=False)
    barrier()
    lib.somax_b200_profile_enable(0)
    slab_model.check_peers()
    buf = C.create_string_buffer(1 << 16)
    _lib.check(lib.somax_b200_profile_report(buf, len(buf)))
    prof = json.loads(buf.value.decode())
    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    cells = nl * nx * ny
    value = cells * K / (ms_max * 1e-3) / 1e9

    # e2e: this rank's window from pinned host memory -> device, K steps, back to pinned host memory
    out_host = torch.empty_like(host).pin_memory()
    barrier()
    t0 = time.perf_counter()
    dev.copy_(host, non_blocking=True)
    slab_model.advance_slab(dev, K, dt, check=False)
    out_host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    nonfinite = float((~torch.isfinite(out_host)).sum())
    t_e = torch.tensor([e2e_s, nonfinite], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    slab_model.check_peers()
    e2e_value = cells * K / float(t_e[0].item()) / 1e9
    hb = host.numel() * host.element_size()
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        padded = nl * (ny // world + 2) * (nx + 2) * 4          # one state-sized array of this rank's slab
        prof.sort(key=lambda r: -r["total_ms"])
        total_prof = sum(r["total_ms"] for r in prof) or 1.0
        compute = [r for r in prof if not r["kernel"].startswith("slab_")] or prof
        top = compute[0]
        k_tr = KERNEL_TRANSFERS.get(top["kernel"], 2.0)
        per_launch_ms = top["total_ms"] / top["launches"]
        achieved = k_tr * padded / (per_launch_ms * 1e-3) / 1e9
        step_alg = TRANSFERS["qg"] * padded
        line = {
            "metric": "cell_updates_per_s", "value": value, "unit": "Gcell-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "slab_vs_single_relL2": slab_rel,
            "config": {**workload_config(args.workload, min(nx, 1024)),
                       "parallelism": f"one grid in {world} y-slab(s), distributed DST by peer-memory transposes over NVLink",
                       "l2": "per-rank working set far larger than the 126 MB L2; no flush needed",
                       "solver": "fft + 3 border columns + thomas (row stages on the slab, column stages on wavenumber strips)",
                       "slab_check": f"{W} steps, slabs vs the single-GPU model on rank 0's window, fails above 1e-5"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Gcell-steps/s", "h2d_bytes_per_step": hb / K, "d2h_bytes_per_step": hb / K,
                    "steps_per_call": K, "nonfinite": float(t_e[1].item()),
                    "note": "per rank: pinned host window -> device, K steps (somax_b200_qgs_steps), device -> pinned host"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": top["kernel"], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "avg_launch_ms": per_launch_ms,
                         "share_of_step": top["total_ms"] / total_prof,
                         "step": {"algorithmic_bytes": step_alg, "achieved_gbs": step_alg / (ms_max / K * 1e-3) / 1e9,
                                  "frac": step_alg / (ms_max / K * 1e-3) / 1e9 / peak, "transfers_per_step": TRANSFERS["qg"]},
                         "kernels": [{"kernel": r["kernel"], "launches": r["launches"], "total_ms": round(r["total_ms"], 3),
                                      "share": round(r["total_ms"] / total_prof, 4)} for r in prof]},
            "cpu_baseline": None,
        }
        emit(line)
    slab_model.close()
    if world > 1:
        dist.destroy_process_group()


def run_gpu_slab_swm(args, world, rank, local):
    """--decomp slab for a shallow-water workload: ONE grid in y-slabs, halo exchange only
    (somax_b200_swms_*).  Same protocol as the QG slab line: W warm-up steps checked against the
    single-GPU model on rank 0's window, K steps timed clean, K steps profiled."""
    import torch
    import torch.distributed as dist
    from somax_b200 import _lib
    import somax_b200 as sb
    from somax_b200.parallel import SlabSWM, slab_window
    kind, nl, nx, ny, members = WORKLOADS[args.workload]
    K, W = args.steps, max(args.warmup, 3)
    lib = _lib.lib()
    model, st0, dt = build_gpu_model(kind, nl, nx, ny)
    slab_model = SlabSWM(model, world, rank=rank)
    win = slab_window(ny, rank, world)
    host = [torch.as_tensor(np.ascontiguousarray(getattr(st0, f)[:, win, :])).pin_memory() for f in "huv"]
    dev = [t.cuda() for t in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    slab_model.advance_slab(*dev, W, dt)
    barrier()
    rel = torch.zeros(1, dtype=torch.float64, device="cuda")
    if rank == 0:
        full = type(st0)(**{f: torch.as_tensor(getattr(st0, f)).cuda() for f in "huv"})
        ref = model.integrate(full, 0.0, W * dt, dt, max_steps=None).ys
        worst = 0.0
        for f, t in zip("huv", dev):
            r = getattr(ref, f)[0][:, win, :].double()
            worst = max(worst, float(torch.linalg.vector_norm((t.double() - r).flatten()) /
                                     torch.linalg.vector_norm(r.flatten())))
        rel[0] = worst
        del full, ref
        model.close()
        torch.cuda.empty_cache()
    if world > 1:
        dist.broadcast(rel, 0)
    slab_rel = float(rel.item())
    if not (slab_rel <= 1e-5):
        sys.stderr.write(f"slab decomposition disagrees with the single-GPU model: relL2 = {slab_rel}\n")
        raise SystemExit(3)
    barrier()
    lib.somax_b200_profile_reset()
    lib.somax_b200_profile_enable(0)
    sampler = ClockSampler(local)
    sampler.start()
    n0 = lib.somax_b200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    slab_model.advance_slab(*dev, K, dt, check=False)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.somax_b200_launch_count() - n0
    clocks = sampler.stop()
    slab_model.check_peers()
    lib.somax_b200_profile_enable(1)
    slab_model.advance_slab(*dev, K, dt, check=False)
    barrier()
    lib.somax_b200_profile_enable(0)
    slab_model.check_peers()
    buf = C.create_string_buffer(1 << 16)
    _lib.check(lib.somax_b200_profile_report(buf, len(buf)))
    prof = json.loads(buf.value.decode())
    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    cells = nl * nx * ny
    value = cells * K / (ms_max * 1e-3) / 1e9
    out_host = [torch.empty_like(t).pin_memory() for t in host]
    barrier()
    t0 = time.perf_counter()
    for d, hsrc in zip(dev, host):
        d.copy_(hsrc, non_blocking=True)
    slab_model.advance_slab(*dev, K, dt, check=False)
    for o, d in zip(out_host, dev):
        o.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    nonfinite = float(sum(int((~torch.isfinite(o)).sum()) for o in out_host))
    t_e = torch.tensor([e2e_s, nonfinite], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    slab_model.check_peers()
    hb = sum(t.numel() * t.element_size() for t in host)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        padded = nl * (ny // world + 2) * (nx + 2) * 4
        prof.sort(key=lambda r: -r["total_ms"])
        total_prof = sum(r["total_ms"] for r in prof) or 1.0
        top = ([r for r in prof if not r["kernel"].startswith("slab_")] or prof)[0]
        k_tr = KERNEL_TRANSFERS.get(top["kernel"], 2.0)
        per_launch_ms = top["total_ms"] / top["launches"]
        step_alg = TRANSFERS["swm"] * padded
        emit({
            "metric": "cell_updates_per_s", "value": value, "unit": "Gcell-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "slab_vs_single_relL2": slab_rel,
            "config": {**workload_config(args.workload, min(nx, 1024)),
                       "parallelism": f"one grid in {world} y-slab(s), one halo row of (h, u, v) per neighbour and "
                                      "evaluation pushed over NVLink",
                       "l2": "per-rank working set larger than the 126 MB L2; no flush needed",
                       "slab_check": f"{W} steps, slabs vs the single-GPU model on rank 0's window, fails above 1e-5"},
            "clocks": clocks,
            "e2e": {"value": cells * K / float(t_e[0].item()) / 1e9, "unit": "Gcell-steps/s",
                    "h2d_bytes_per_step": hb / K, "d2h_bytes_per_step": hb / K, "steps_per_call": K,
                    "nonfinite": float(t_e[1].item()),
                    "note": "per rank: pinned host windows -> device, K steps (somax_b200_swms_steps), device -> pinned host"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": top["kernel"], "achieved": k_tr * padded / (per_launch_ms * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": k_tr * padded / (per_launch_ms * 1e-3) / 1e9 / peak,
                         "traffic": None, "avg_launch_ms": per_launch_ms, "share_of_step": top["total_ms"] / total_prof,
                         "step": {"algorithmic_bytes": step_alg, "achieved_gbs": step_alg / (ms_max / K * 1e-3) / 1e9,
                                  "frac": step_alg / (ms_max / K * 1e-3) / 1e9 / peak, "transfers_per_step": TRANSFERS["swm"]},
                         "kernels": [{"kernel": r["kernel"], "launches": r["launches"], "total_ms": round(r["total_ms"], 3),
                                      "share": round(r["total_ms"] / total_prof, 4)} for r in prof]},
            "cpu_baseline": None,
        })
    slab_model.close()
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """Route fd 1 to stderr while libraries initialise (NCCL prints its version banner on stdout);
    `emit` writes the one JSON line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())
    else:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="qg3_8192", choices=sorted(WORKLOADS))
    ap.add_argument("--decomp", default=None, choices=["members", "slab"],
                    help="N > 1: 'slab' = ONE grid in y-slabs (strong scaling; the default for single-grid QG "
                         "workloads: BASELINE config 4); 'members' = one grid per GPU / members sharded by rank "
                         "(the default for ensembles and shallow water); with one GPU both are the single-GPU path")
    args = ap.parse_args()
    if args.decomp is None:
        kind, _nl, _nx, _ny, members = WORKLOADS[args.workload]
        args.decomp = "slab" if (kind == "qg" and members == 1) else "members"
    if args.impl == "reference":
        run_reference(args)
    elif args.decomp == "slab" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        run_gpu_slab(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
