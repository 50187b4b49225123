"""ctypes binding of ``libsomax_b200.so`` (the C ABI declared in ``include/somax_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``somax_b200._lib.build_library()``
with ``nvcc -gencode arch=compute_100a,code=sm_100a``.  Loading it needs no GPU; every compute
entry point fails loudly (``SomaxB200Error``) without one.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "lib" / "libsomax_b200.so"
SOURCES = ["layout.cu", "swm.cu", "swm_f64.cu", "qg_solver.cu", "qg.cu"]
# per-source extra flags: the fp64 shallow-water kernel keeps the reference's rounding (no FMA contraction)
EXTRA_FLAGS = {"swm_f64.cu": ["-fmad=false"]}
NVCC_FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
              "-lineinfo", "-O3", "-std=c++17"]

F32, F64 = 0, 1
BC_PERIODIC, BC_WALL = 0, 1
SOLVER_AUTO, SOLVER_FFT, SOLVER_DENSE = 0, 1, 2
SPEC_ADVECTION_REGION2, SPEC_DIFFUSION_FLUX, SPEC_DST_CONTINUOUS, SPEC_DST_INTERIOR, SPEC_KEEP_PSI_RING = 1, 2, 4, 8, 16
# the reference's conventions (pinned by its printed tutorial outputs, see oracle/operators.py)
DEFAULT_SPEC = SPEC_ADVECTION_REGION2 | SPEC_DIFFUSION_FLUX


class SomaxB200Error(RuntimeError):
    pass


class ParamsStruct(C.Structure):
    _fields_ = [("lateral_viscosity", C.c_double), ("bottom_drag", C.c_double),
                ("wind_amplitude", C.c_double), ("H0", C.c_double)]


# name -> (restype, argtypes); kept in one table so tests can check every header symbol.
_P, _I, _D, _L, _U = C.c_void_p, C.c_int, C.c_double, C.c_long, C.c_uint
_PP = C.POINTER(ParamsStruct)
SIGNATURES = {
    "somax_b200_last_error": (C.c_char_p, []),
    "somax_b200_abi_version": (_I, []),
    "somax_b200_launch_count": (C.c_uint64, []),
    "somax_b200_profile_enable": (None, [_I]),
    "somax_b200_profile_reset": (None, []),
    "somax_b200_profile_report": (_I, [C.c_char_p, C.c_size_t]),
    "somax_b200_qg_create": (_I, [C.POINTER(_P), _I, _I, _I, _I, _I, _D, _D, _P, _P, _P, _P, _P, _I, _U]),
    "somax_b200_qg_destroy": (_I, [_P]),
    "somax_b200_qg_device_bytes": (C.c_size_t, [_P]),
    "somax_b200_qg_apply_bc": (_I, [_P, _P, _P, _P]),
    "somax_b200_qg_invert": (_I, [_P, _P, _P, _P]),
    "somax_b200_qg_rhs": (_I, [_P, _P, _P, _P, _PP, _I, _P]),
    "somax_b200_qg_steps": (_I, [_P, _P, _L, _D, _D, _PP, _P]),
    "somax_b200_qg_resume": (_I, [_P, _P, _L, _D, _D, _PP, _P]),
    "somax_b200_qg_diag": (_I, [_P, _P, _P, _P]),
    "somax_b200_qgs_create": (_I, [C.POINTER(_P), _I, _I, _I, _I, _D, _D, _P, _P, _P, _P, _P, _I, _I, _I, _U]),
    "somax_b200_qgs_destroy": (_I, [_P]),
    "somax_b200_qgs_device_bytes": (C.c_size_t, [_P]),
    "somax_b200_qgs_export_bytes": (C.c_size_t, []),
    "somax_b200_qgs_export": (_I, [_P, _P]),
    "somax_b200_qgs_attach": (_I, [_P, _P]),
    "somax_b200_qgs_steps": (_I, [_P, _P, _L, _D, _D, _PP, _P]),
    "somax_b200_qgs_status": (_I, [_P, C.POINTER(_I)]),
    "somax_b200_swms_create": (_I, [C.POINTER(_P), _I, _I, _I, _I, _D, _D, _I, _P, _P, _P, _P, _I, _I, _I, _U]),
    "somax_b200_swms_destroy": (_I, [_P]),
    "somax_b200_swms_device_bytes": (C.c_size_t, [_P]),
    "somax_b200_swms_export_bytes": (C.c_size_t, []),
    "somax_b200_swms_export": (_I, [_P, _P]),
    "somax_b200_swms_attach": (_I, [_P, _P]),
    "somax_b200_swms_steps": (_I, [_P, _P, _P, _P, _L, _D, _D, _PP, _P]),
    "somax_b200_swms_status": (_I, [_P, C.POINTER(_I)]),
    "somax_b200_swm_create": (_I, [C.POINTER(_P), _I, _I, _I, _I, _I, _D, _D, _I, _P, _P, _P, _P, _U]),
    "somax_b200_swm_destroy": (_I, [_P]),
    "somax_b200_swm_set_projection": (_I, [_P, _D, _P, _P, _P, _P, _P, _I]),
    "somax_b200_swm_device_bytes": (C.c_size_t, [_P]),
    "somax_b200_swm_apply_bc": (_I, [_P, _P, _P, _P, _P, _P, _P, _P]),
    "somax_b200_swm_project": (_I, [_P, _P, _P, _P, _P, _P, _P, _P]),
    "somax_b200_swm_rhs": (_I, [_P, _P, _P, _P, _P, _P, _P, _PP, _I, _P]),
    "somax_b200_swm_steps": (_I, [_P, _P, _P, _P, _L, _D, _D, _PP, _P]),
    "somax_b200_swm_resume": (_I, [_P, _P, _P, _P, _L, _D, _D, _PP, _P]),
    "somax_b200_swm_diag": (_I, [_P, _P, _P, _P, _P, _P]),
}

_lib = None


def build_library(verbose: bool = False) -> Path:
    """Compile the CUDA sources in-tree for sm_100a (nvcc cross-compiles without a GPU): one
    object per source, compiled in parallel, then one shared library."""
    LIB_PATH.parent.mkdir(parents=True, exist_ok=True)
    srcs = [CSRC / s for s in SOURCES]
    hdrs = [str(p) for p in CSRC.glob("*.cuh")] + [str(PKG_DIR.parent / "include" / "somax_b200.h")]
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("SOMAX_B200_NVCC_EXTRA", "").split()     # e.g. -DSB_TH_DEBUG (experiments)
    objdir = PKG_DIR / "lib" / "obj"
    objdir.mkdir(parents=True, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"]
    jobs = []
    for src in srcs:
        obj = objdir / (src.stem + ".o")
        deps = [str(src)] + hdrs
        if extra or not obj.exists() or any(os.path.getmtime(d) > os.path.getmtime(obj) for d in deps):
            cmd = [nvcc] + flags + EXTRA_FLAGS.get(src.name, []) + extra + ["-c", "-o", str(obj), str(src)]
            if verbose:
                print(" ".join(cmd))
            jobs.append((cmd, subprocess.Popen(cmd)))
    for cmd, pr in jobs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, cmd)
    objs = [str(objdir / (src.stem + ".o")) for src in srcs]
    if jobs or not LIB_PATH.exists() or any(os.path.getmtime(o) > os.path.getmtime(LIB_PATH) for o in objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH)] + objs
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded library; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise SomaxB200Error(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (somax_b200 has no CPU fallback)")
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().somax_b200_last_error().decode("utf-8", "replace")
        raise SomaxB200Error(f"libsomax_b200 error {rc}: {msg}")
