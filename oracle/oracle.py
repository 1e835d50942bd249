"""ctypes wrapper of oracle/libmpas_oracle.so -- TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs, never by the product package.  Exposes the same TaskAPI as
mpas_regent_b200.dynamics.Dynamics so parity tests run one call sequence on both.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

from mpas_regent_b200 import _abi
from mpas_regent_b200._abi import FIELD_ID, MpasConfig, MpasDims
from mpas_regent_b200.dynamics import TaskAPI, _declare

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmpas_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mpas_oracle.cpp")
    deps = [src, os.path.join(_abi.INCLUDE_DIR, "mpas_b200.h"), os.path.join(_abi.INCLUDE_DIR, "mpas_b200_fields.def")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in deps):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(_SO)
        _declare(lib, "oracle_")
        lib.oracle_upload_field.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.oracle_download_field.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.oracle_download_pad.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.oracle_set_threads.argtypes = [C.c_void_p, C.c_int]
        lib.oracle_set_class.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.oracle_max_threads.restype = C.c_int
        _lib = lib
    return _lib


class Oracle(TaskAPI):
    def __init__(self, dims: MpasDims, cfg: Optional[MpasConfig] = None, threads: int = 1):
        self._lib = load()
        self.dims = dims
        self.cfg = cfg if cfg is not None else _abi.default_config()
        self._h = C.c_void_p()
        rc = self._lib.oracle_create(C.byref(self.dims), C.byref(self.cfg), C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"oracle_create failed ({rc})")
        self._lib.oracle_set_threads(self._h, threads)

    def _call(self, name: str, *args):
        rc = getattr(self._lib, "oracle_" + name)(self._h, *args)
        if rc != 0:
            raise RuntimeError(f"oracle_{name} failed ({rc})")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.oracle_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def restrict(self, entity: int, cls=None):
        """TaskAPI.restrict: the acoustic step / divergence damping run on one launch class only (None = everything)."""
        self._lib.oracle_set_class(self._h, int(entity), -1 if cls is None else int(cls))

    def set_threads(self, n: int):
        self._lib.oracle_set_threads(self._h, n)

    @staticmethod
    def max_threads() -> int:
        return int(load().oracle_max_threads())

    def upload_mesh(self, static: Dict[str, np.ndarray]):
        m, keep = _abi.mesh_ptrs(static, self.dims)
        rc = self._lib.oracle_upload_mesh(self._h, C.byref(m))
        del keep
        if rc != 0:
            raise RuntimeError(f"oracle_upload_mesh failed ({rc})")

    def upload_field(self, name: str, a: np.ndarray):
        a = np.ascontiguousarray(a, dtype=np.float64)
        shp = self.field_shape(name)
        if a.shape != shp:
            raise ValueError(f"{name}: shape {a.shape}, expected {shp}")
        rc = self._lib.oracle_upload_field(self._h, FIELD_ID[name], a.ctypes.data)
        assert rc == 0

    def download_field(self, name: str, out: Optional[np.ndarray] = None) -> np.ndarray:
        shp = self.field_shape(name)
        a = out if out is not None else np.empty(shp, dtype=np.float64)
        rc = self._lib.oracle_download_field(self._h, FIELD_ID[name], a.ctypes.data)
        assert rc == 0
        return a

    def download_pad(self, name: str) -> np.ndarray:
        shp = self.field_shape(name)[1:]
        a = np.empty(shp, dtype=np.float64)
        self._lib.oracle_download_pad(self._h, FIELD_ID[name], a.ctypes.data)
        return a

    def sync(self):
        pass

    # ---- summarize_timestep (rk_timestep.rg:29-359): numpy restatement of mpasb200_summarize_field ---------------
    def set_global_ids(self, entity: int, gid):
        self._gid = getattr(self, "_gid", {})
        self._gid[entity] = None if gid is None else np.asarray(gid, dtype=np.int64)

    def summarize_field(self, name: str, n_first=None, nlevels=None) -> dict:
        ent = _abi.FIELD_ENTITY[name]
        a = self.download_field(name)
        n = a.shape[0] if n_first is None else int(n_first)
        nl = a.shape[1] if nlevels is None else int(nlevels)
        gid = getattr(self, "_gid", {}).get(ent)
        return summarize_np(a[:n, :nl], None if gid is None else gid[:n])


_M64 = (1 << 64) - 1


def _mix64(z: np.ndarray) -> np.ndarray:
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xbf58476d1ce4e5b9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94d049bb133111eb)
    return z ^ (z >> np.uint64(31))


def summarize_np(a: np.ndarray, gid=None) -> dict:
    """the definition in include/mpas_b200.h (MpasFieldSummary), array-at-a-time"""
    a = np.ascontiguousarray(a, dtype=np.float64)
    n, nl = a.shape
    ids = (np.arange(n, dtype=np.uint64) if gid is None else np.asarray(gid).astype(np.uint64))[:, None] * np.uint64(nl) \
        + np.arange(nl, dtype=np.uint64)[None, :]
    bits = a.view(np.uint64).copy()
    nan = np.isnan(a)
    bits[nan] = np.uint64(0x7ff8000000000000)
    with np.errstate(over="ignore"):
        cs = int(_mix64(bits + np.uint64(0x9e3779b97f4a7c15) * (ids + np.uint64(1))).sum(dtype=np.uint64)) & _M64
    raw = a.view(np.uint64)
    key = np.where(raw >> np.uint64(63) != 0, ~raw, raw | np.uint64(1 << 63))        # monotone double -> uint64
    ok = ~nan
    out = {"n_nan": int(nan.sum()), "n_inf": int(np.isinf(a).sum()), "count": int(a.size), "checksum": f"{cs:016x}"}
    if ok.any():
        kmin, kmax = key[ok].min(), key[ok].max()
        pmin, pmax = ids[ok & (key == kmin)].min(), ids[ok & (key == kmax)].min()
        unkey = lambda k: np.array([k & np.uint64(_M64 >> 1) if k >> np.uint64(63) else ~k], dtype=np.uint64).view(np.float64)[0]
        out.update(min=float(unkey(kmin)), max=float(unkey(kmax)), min_at=[int(pmin) // nl, int(pmin) % nl],
                   max_at=[int(pmax) // nl, int(pmax) % nl])
    else:
        out.update(min=float("inf"), max=float("-inf"), min_at=[-1, -1], max_at=[-1, -1])
    return out
