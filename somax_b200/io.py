"""Snapshot / restart IO: somax States <-> self-describing datasets <-> zarr v3 stores.

Mirror of somax/_src/io/xarray.py:39-286 (same function names, dimension conventions, attrs and
error messages; pinned by the reference's tests/test_io_xarray.py).  xarray and zarr are not in
this image, so the dataset is a small container (`Dataset` / `Variable`, the subset of the xarray
API the runner uses) and the store is written directly in the zarr v3 layout xarray produces
with ``to_zarr(zarr_format=3, consolidated=False)``: one group ``zarr.json`` carrying the dataset
attrs, one array directory per variable / coordinate with ``dimension_names`` in its
``zarr.json``, uncompressed little-endian ``bytes`` chunks under ``c/``, one chunk per time
slice so that appending along ``time`` only adds chunk files.  ``xarray.open_zarr(path,
consolidated=False)`` opens these stores elsewhere.

CUDA tensors are read back through pinned host buffers (`_host`); nothing here is on the
stepping hot path.
"""
from __future__ import annotations

import dataclasses
import importlib
import json
import shutil
from pathlib import Path
from typing import Any

import numpy as np

_DIMS = {1: ("x",), 2: ("y", "x"), 3: ("layer", "y", "x")}      # io/xarray.py:39-43


def _infer_dims(rank: int, *, time_axis: bool):
    base = _DIMS.get(rank) or tuple(f"dim{i}" for i in range(rank))
    return ("time",) + base if time_axis else base


def _host(x) -> np.ndarray:
    """numpy view of a leaf (numpy array, CPU or CUDA torch tensor)."""
    if isinstance(x, np.ndarray):
        return x
    if hasattr(x, "detach") and hasattr(x, "cpu"):
        return x.detach().cpu().numpy()
    return np.asarray(x)


@dataclasses.dataclass
class Variable:
    dims: tuple
    values: np.ndarray

    def isel(self, **idx):
        v, dims = self.values, list(self.dims)
        for name, i in idx.items():
            ax = dims.index(name)
            v = np.take(v, i, axis=ax)
            dims.pop(ax)
        return Variable(tuple(dims), v)

    @property
    def shape(self):
        return self.values.shape

    @property
    def dtype(self):
        return self.values.dtype


class Dataset:
    """Data variables + coordinates + attrs (the part of ``xarray.Dataset`` this path uses)."""

    def __init__(self, data_vars=None, coords=None, attrs=None):
        self.data_vars = {k: Variable(tuple(d), np.asarray(v)) for k, (d, v) in (data_vars or {}).items()}
        self.coords = {k: Variable(tuple(d), np.asarray(v)) for k, (d, v) in (coords or {}).items()}
        self.attrs = dict(attrs or {})

    @property
    def variables(self):
        return {**self.coords, **self.data_vars}

    def __getitem__(self, name) -> Variable:
        return self.variables[name]

    def __contains__(self, name):
        return name in self.data_vars or name in self.coords

    @property
    def sizes(self):
        out = {}
        for v in self.variables.values():
            out.update(dict(zip(v.dims, v.shape)))
        return out

    def load(self):
        return self


def _state_attrs(state_class: type):
    return {"state_class": state_class.__name__, "state_module": state_class.__module__}


def state_to_dataset(state, *, time=None, attrs=None) -> Dataset:
    """One State -> Dataset (io/xarray.py:71-116): a variable per field, dims by rank, a singleton
    ``time`` axis + float64 coordinate when ``time`` is given."""
    data = {}
    for f in dataclasses.fields(state):
        leaf = _host(getattr(state, f.name))
        dims = _infer_dims(leaf.ndim, time_axis=time is not None)
        data[f.name] = (dims, leaf[None] if time is not None else leaf)
    coords = {"time": (("time",), np.asarray([time], dtype=np.float64))} if time is not None else {}
    ds = Dataset(data, coords)
    ds.attrs.update(_state_attrs(type(state)))
    ds.attrs.update(attrs or {})
    return ds


def snapshots_to_dataset(snapshots, ts, *, state_class=None, attrs=None) -> Dataset:
    """Time-stacked states (``sol.ys``) -> Dataset with a ``time`` coordinate (io/xarray.py:119-169)."""
    state_class = state_class or type(snapshots)
    ts_np = _host(ts)
    if ts_np.ndim != 1:
        raise ValueError(f"ts must be 1-D, got shape {ts_np.shape}")
    data = {}
    for f in dataclasses.fields(state_class):
        leaf = _host(getattr(snapshots, f.name))
        if leaf.shape[0] != ts_np.shape[0]:
            raise ValueError(f"snapshot leaf {f.name!r} has leading dim {leaf.shape[0]} "
                             f"but ts has length {ts_np.shape[0]}")
        data[f.name] = (_infer_dims(leaf.ndim - 1, time_axis=True), leaf)
    ds = Dataset(data, {"time": (("time",), ts_np)})
    ds.attrs.update(_state_attrs(state_class))
    ds.attrs.update(attrs or {})
    return ds


# reference state classes -> the classes of this package (stores written by somax itself load too)
_ALLOWED_ROOTS = ("somax_b200", "somax")


def _resolve_state_class(module_name: str, class_name: str):
    root = module_name.split(".")[0]
    if root not in _ALLOWED_ROOTS or not (module_name == root or module_name.startswith(root + ".")):
        raise ValueError(f"refusing to auto-import state_module {module_name!r}: only modules under "
                         "'somax.' / 'somax_b200.' are allowlisted for auto-recovery. "
                         "Pass state_class explicitly to load this dataset.")
    if root == "somax":                       # the reference's module tree is not importable here
        module = importlib.import_module("somax_b200")
    else:
        module = importlib.import_module(module_name)
    return getattr(module, class_name)


def dataset_to_state(ds: Dataset, state_class=None, *, time_index: int = -1):
    """Dataset -> State (io/xarray.py:172-230); last time slice by default."""
    if state_class is None:
        try:
            module_name, class_name = ds.attrs["state_module"], ds.attrs["state_class"]
        except KeyError as exc:
            raise ValueError("dataset_to_state called without state_class and the Dataset is "
                             "missing 'state_class'/'state_module' attrs") from exc
        state_class = _resolve_state_class(module_name, class_name)
    kw = {}
    for f in dataclasses.fields(state_class):
        if f.name not in ds.variables:
            raise ValueError(f"Dataset is missing variable {f.name!r} required by {state_class.__name__}")
        var = ds[f.name]
        if "time" in var.dims:
            var = var.isel(time=time_index)
        kw[f.name] = np.ascontiguousarray(var.values)
    return state_class(**kw)


# ----------------------------------------------------------------------------------------
# zarr v3 directory stores
# ----------------------------------------------------------------------------------------
_ZTYPES = {"float32": "float32", "float64": "float64", "int32": "int32", "int64": "int64",
           "uint8": "uint8", "bool": "bool"}


def _json_attr(v):
    if isinstance(v, (np.floating,)):
        return float(v)
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, np.ndarray):
        return v.tolist()
    return v


def _chunk_shape(var: Variable):
    return tuple(1 if d == "time" else max(int(n), 1) for d, n in zip(var.dims, var.shape))


def _array_meta(var: Variable, shape=None):
    dt = np.dtype(var.dtype)
    if dt.name not in _ZTYPES:
        raise ValueError(f"unsupported dtype {dt} for the zarr store")
    return {
        "zarr_format": 3, "node_type": "array", "shape": [int(n) for n in (shape or var.shape)],
        "data_type": _ZTYPES[dt.name],
        "chunk_grid": {"name": "regular", "configuration": {"chunk_shape": list(_chunk_shape(var))}},
        "chunk_key_encoding": {"name": "default", "configuration": {"separator": "/"}},
        "fill_value": "NaN" if dt.kind == "f" else 0,
        "codecs": [{"name": "bytes", "configuration": {"endian": "little"}}],
        "attributes": {}, "dimension_names": list(var.dims),
    }


def _write_chunks(adir: Path, var: Variable, t_offset: int = 0):
    """One chunk per index of the time axis (or a single chunk c/0/0... without one)."""
    vals = np.ascontiguousarray(var.values).astype(var.dtype.newbyteorder("<"), copy=False)
    if "time" in var.dims:
        ax = var.dims.index("time")
        for t in range(vals.shape[ax]):
            key = ["0"] * vals.ndim
            key[ax] = str(t + t_offset)
            p = adir.joinpath("c", *key)
            p.parent.mkdir(parents=True, exist_ok=True)
            np.ascontiguousarray(np.take(vals, [t], axis=ax)).tofile(p)
    else:
        p = adir.joinpath("c", *(["0"] * vals.ndim)) if vals.ndim else adir / "c"
        p.parent.mkdir(parents=True, exist_ok=True)
        vals.tofile(p)


def save_dataset(ds: Dataset, path, *, mode: str = "w") -> None:
    """Persist as a zarr v3 store (io/xarray.py:233-249).  mode: "w" overwrite, "w-" fail if the
    store exists, "a" add / replace variables in an existing store."""
    path = Path(path)
    if mode not in ("w", "w-", "a"):
        raise ValueError(f"mode must be 'w', 'w-' or 'a', got {mode!r}")
    if path.exists():
        if mode == "w-":
            raise FileExistsError(f"zarr store {path} already exists (mode='w-')")
        if mode == "w":
            shutil.rmtree(path)
    path.mkdir(parents=True, exist_ok=True)
    attrs = {}
    if mode == "a" and (path / "zarr.json").exists():
        attrs = json.loads((path / "zarr.json").read_text()).get("attributes", {})
    attrs.update({k: _json_attr(v) for k, v in ds.attrs.items()})
    (path / "zarr.json").write_text(json.dumps(
        {"zarr_format": 3, "node_type": "group", "attributes": attrs}, indent=1))
    for name, var in ds.variables.items():
        adir = path / name
        if adir.exists():
            shutil.rmtree(adir)
        adir.mkdir()
        (adir / "zarr.json").write_text(json.dumps(_array_meta(var), indent=1))
        _write_chunks(adir, var)


def _gunzip(raw: bytes, nbytes: int) -> bytes:
    import zlib
    return zlib.decompress(raw, wbits=31)


def _unzstd(raw: bytes, nbytes: int) -> bytes:
    """zstd frames as zarr-python's default v3 compressor writes them (what `ds.to_zarr(...,
    zarr_format=3)` produces when no encoding is given).  The standard library has no zstd before
    Python 3.14; pyarrow (in this image) or the `zstandard` package decode it."""
    try:
        from compression import zstd          # Python >= 3.14
        return zstd.decompress(raw)
    except ImportError:
        pass
    try:
        import zstandard
        return zstandard.ZstdDecompressor().decompress(raw, max_output_size=nbytes)
    except ImportError:
        pass
    try:
        import pyarrow as pa
        return pa.decompress(raw, decompressed_size=nbytes, codec="zstd").to_pybytes()
    except ImportError as e:       # pragma: no cover
        raise ValueError("this store is zstd-compressed and neither compression.zstd, zstandard nor pyarrow "
                         "is importable; rewrite it with encoding={var: {'compressors': None}}") from e


# bytes -> bytes codecs of the zarr v3 chain the reader can undo
_DECODERS = {"gzip": _gunzip, "zstd": _unzstd}


def load_dataset(path) -> Dataset:
    """Open a zarr v3 store as a Dataset (io/xarray.py:252-266): the ones `save_dataset` writes
    (uncompressed) and the ones xarray / zarr-python write with their default zstd (or a gzip)
    compressor, e.g. a `final_state.zarr` of the reference's own somax-sim."""
    path = Path(path)
    meta_p = path / "zarr.json"
    if not meta_p.exists():
        raise FileNotFoundError(f"{path} is not a zarr v3 store (no zarr.json)")
    group = json.loads(meta_p.read_text())
    if group.get("zarr_format") != 3 or group.get("node_type") != "group":
        raise ValueError(f"{path}: expected a zarr v3 group")
    data, coords = {}, {}
    for adir in sorted(p for p in path.iterdir() if p.is_dir() and (p / "zarr.json").exists()):
        m = json.loads((adir / "zarr.json").read_text())
        if m.get("node_type") != "array":
            continue
        codecs = [c["name"] for c in m.get("codecs", [])]
        if not codecs or codecs[0] != "bytes" or any(c not in _DECODERS for c in codecs[1:]):
            raise ValueError(
                f"{adir}: codec chain {codecs} is not supported (supported: 'bytes' optionally followed by "
                f"{sorted(_DECODERS)}).  Write the store with xarray's "
                "`ds.to_zarr(path, zarr_format=3, encoding={var: {'compressors': None} for var in ds.variables})`, "
                "or with a zstd / gzip compressor")
        endian = m["codecs"][0].get("configuration", {}).get("endian", "little")
        dt = np.dtype(m["data_type"]).newbyteorder("<" if endian == "little" else ">")
        shape = tuple(m["shape"])
        cshape = tuple(m["chunk_grid"]["configuration"]["chunk_shape"])
        sep = m.get("chunk_key_encoding", {}).get("configuration", {}).get("separator", "/")
        out = np.full(shape, np.nan if dt.kind == "f" else 0, dtype=dt.newbyteorder("="))
        nchunks = [(-(-n // c)) for n, c in zip(shape, cshape)]
        for idx in np.ndindex(*nchunks) if shape else [()]:
            key = sep.join(["c"] + [str(i) for i in idx]) if shape else "c"
            p = adir / key
            if not p.exists():
                continue
            raw = p.read_bytes()
            nbytes = int(np.prod(cshape, dtype=np.int64)) * dt.itemsize
            for cname in reversed(codecs[1:]):      # bytes -> bytes codecs are undone last to first
                raw = _DECODERS[cname](raw, nbytes)
            block = np.frombuffer(raw, dtype=dt).reshape(cshape)
            sl = tuple(slice(i * c, min((i + 1) * c, n)) for i, c, n in zip(idx, cshape, shape))
            out[sl] = block[tuple(slice(0, s.stop - s.start) for s in sl)]
        dims = tuple(m.get("dimension_names") or [f"dim{i}" for i in range(len(shape))])
        (coords if (len(dims) == 1 and dims[0] == adir.name) else data)[adir.name] = (dims, out)
    return Dataset(data, coords, group.get("attributes", {}))


def append_to_dataset(ds: Dataset, path, *, append_dim: str = "time") -> None:
    """Append along ``append_dim`` to an existing store (io/xarray.py:269-286): new chunk files
    plus the grown ``shape`` in each array's metadata."""
    path = Path(path)
    if append_dim != "time":
        raise NotImplementedError("append is implemented along 'time' (one chunk per time slice)")
    if not (path / "zarr.json").exists():
        raise FileNotFoundError(f"{path} is not an existing zarr store")
    for name, var in ds.variables.items():
        if append_dim not in var.dims:
            continue
        adir = path / name
        if not (adir / "zarr.json").exists():
            raise ValueError(f"store {path} has no variable {name!r} to append to")
        m = json.loads((adir / "zarr.json").read_text())
        ax = var.dims.index(append_dim)
        old = list(m["shape"])
        if list(m.get("dimension_names", [])) != list(var.dims) or \
                [n for i, n in enumerate(old) if i != ax] != [n for i, n in enumerate(var.shape) if i != ax]:
            raise ValueError(f"variable {name!r}: dims / non-append shape differ from the store")
        if np.dtype(m["data_type"]) != np.dtype(var.dtype):
            raise ValueError(f"variable {name!r}: dtype {var.dtype} differs from the store's {m['data_type']}")
        _write_chunks(adir, var, t_offset=old[ax])
        old[ax] += var.shape[ax]
        m["shape"] = old
        (adir / "zarr.json").write_text(json.dumps(m, indent=1))


class AsyncSnapshotWriter:
    """Snapshot sink for the runner: device -> pinned host copies are issued on a side stream and
    the zarr chunks are written by a worker thread, so stepping continues while a snapshot leaves
    the device.  `put` is called with the state at a save time; `close`
    drains.  The pinned staging buffers form a ring of `depth + 1` sets that is allocated once
    (`cudaHostAlloc` of a state-sized buffer costs more than the copy it serves): `put` blocks only
    when every set is still waiting to be written."""

    def __init__(self, path, state_class, attrs=None, depth: int = 2):
        import queue
        import threading
        self.path, self.state_class, self.attrs = Path(path), state_class, dict(attrs or {})
        self._q = queue.Queue()
        self._free = queue.Queue()
        for _ in range(depth + 1):
            self._free.put({})                  # one set of pinned leaves per slot, filled on first use
        self._err = None
        self._first = True
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        self._side = None

    def _run(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            t, leaves, ev, slot = item
            try:
                if ev is not None:
                    ev.synchronize()
                st = self.state_class(**{k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in leaves.items()})
                ds = state_to_dataset(st, time=float(t), attrs=self.attrs)
                if self._first:
                    save_dataset(ds, self.path, mode="w")
                    self._first = False
                else:
                    append_to_dataset(ds, self.path)
            except Exception as exc:      # surfaced by close()
                self._err = exc
            finally:
                self._free.put(slot)

    def put(self, t, state):
        leaves, ev = {}, None
        try:
            import torch
        except Exception:   # pragma: no cover
            torch = None
        slot = self._free.get()
        cuda_leaves = torch is not None and any(
            isinstance(getattr(state, f.name), torch.Tensor) and getattr(state, f.name).is_cuda
            for f in dataclasses.fields(state))
        if cuda_leaves:
            if self._side is None:
                self._side = torch.cuda.Stream()
            self._side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                for f in dataclasses.fields(state):
                    x = getattr(state, f.name)
                    host = slot.get(f.name)
                    if host is None or host.shape != x.shape or host.dtype != x.dtype:
                        host = slot[f.name] = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
                    host.copy_(x, non_blocking=True)
                    x.record_stream(self._side)
                    leaves[f.name] = host
                ev = torch.cuda.Event()
                ev.record(self._side)
        else:
            for f in dataclasses.fields(state):
                leaves[f.name] = np.array(_host(getattr(state, f.name)))
        self._q.put((t, leaves, ev, slot))

    def close(self):
        self._q.put(None)
        self._th.join()
        if self._err is not None:
            raise self._err


__all__ = ["Dataset", "Variable", "state_to_dataset", "snapshots_to_dataset", "dataset_to_state",
           "save_dataset", "load_dataset", "append_to_dataset", "AsyncSnapshotWriter"]
