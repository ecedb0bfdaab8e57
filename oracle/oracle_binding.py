"""TEST INFRASTRUCTURE ONLY — ctypes view of oracle/libmas_oracle.so
(the plain-C restatement in oracle/mas_oracle.c; float and double variants).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmas_oracle.so")
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def morton_encode(x: float, y: float, z: float) -> int:
    f = _load().maso_f_morton_encode
    f.restype = C.c_ulonglong
    f.argtypes = [C.c_float] * 3
    return int(f(x, y, z))


class OraclePreconditioner:
    """Restated MAS preconditioner; precision 'f' (float) or 'd' (FP64 arbiter)."""

    def __init__(self, precision: str = "f", prolong_all_levels: bool = False):
        assert precision in ("f", "d")
        self.lib = _load()
        self.pfx = f"maso_{precision}_"
        self.real = np.float32 if precision == "f" else np.float64
        create = self._fn("create", C.c_void_p, [])
        self.h = create()
        self._fn("set_option", None, [C.c_void_p, C.c_int, C.c_int])(self.h, 0, int(prolong_all_levels))
        self.mesh = None

    def _fn(self, name, restype, argtypes):
        f = getattr(self.lib, self.pfx + name)
        f.restype = restype
        f.argtypes = argtypes
        return f

    def close(self):
        if self.h:
            self._fn("destroy", None, [C.c_void_p])(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def allocate(self, mesh):
        self.mesh = mesh
        vp = C.c_void_p
        f = self._fn("allocate", C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp])
        pos = np.ascontiguousarray(mesh.positions, np.float32)
        rc = f(self.h, mesh.nv, mesh.ne, mesh.nf, _p(pos), _p(np.ascontiguousarray(mesh.edges, np.int32)),
               _p(np.ascontiguousarray(mesh.faces, np.int32)), _p(mesh.nbr_starts), _p(mesh.nbr_idx))
        if rc != 0:
            raise RuntimeError(f"oracle allocate failed: {rc}")

    def prepare(self, mesh=None):
        m = mesh or self.mesh
        vp = C.c_void_p
        f = self._fn("prepare", C.c_int, [vp, vp, vp, vp, vp, vp, vp, C.c_uint, C.c_uint, C.c_uint])
        dummy = np.zeros(64, np.uint8)
        ef = m.ef if m.ef.size else dummy
        ee = m.ee if m.ee.size else dummy
        vf = m.vf if m.vf.size else dummy
        rc = f(self.h, _p(m.diag), _p(m.offdiag), _p(m.nbr_starts), _p(ef), _p(ee), _p(vf),
               m.ef_total, m.ee_total, m.vf_total)
        if rc != 0:
            raise RuntimeError(f"oracle prepare failed: {rc}")

    def apply(self, r: np.ndarray, out=None) -> np.ndarray:
        r = np.ascontiguousarray(r, np.float32)
        z = np.zeros_like(r) if out is None else out
        self._fn("apply", C.c_int, [C.c_void_p] * 3)(self.h, _p(z), _p(r))
        return z

    # ---- introspection
    @property
    def nv(self):
        return self.mesh.nv

    @property
    def num_level(self):
        return self._fn("num_level", C.c_int, [C.c_void_p])(self.h)

    @property
    def total_clusters(self):
        return self._fn("total_clusters", C.c_int, [C.c_void_p])(self.h)

    @property
    def stencil_num(self):
        return self._fn("stencil_num", C.c_int, [C.c_void_p])(self.h)

    def _get(self, name, shape, dtype, *extra):
        out = np.zeros(shape, dtype)
        argt = [C.c_void_p] + [C.c_int] * len(extra) + [C.c_void_p]
        self._fn(name, None, argt)(self.h, *extra, _p(out))
        return out

    def aabb(self):
        lo, hi = np.zeros(4, np.float32), np.zeros(4, np.float32)
        self._fn("get_aabb", None, [C.c_void_p] * 3)(self.h, _p(lo), _p(hi))
        return lo, hi

    def level_size(self):
        return self._get("get_level_size", (self.num_level + 1, 2), np.int32)

    def morton(self):
        return self._get("get_morton", self.nv, np.uint64)

    def sorted_get_original(self):
        return self._get("get_sorted_get_original", self.nv, np.int32)

    def original_get_sorted(self):
        return self._get("get_original_get_sorted", self.nv, np.int32)

    def going_next(self, count=None):
        count = self.total_clusters if count is None else count
        out = np.zeros(count, np.int32)
        self._fn("get_going_next", None, [C.c_void_p, C.c_void_p, C.c_int])(self.h, _p(out), count)
        return out

    def coarse_tables(self):
        return self._get("get_coarse_tables", (self.nv, 4), np.int32)

    def coarse_space_table(self, level: int):
        return self._get("get_coarse_space_table", self.nv, np.int32, level)

    def fine_connect_mask(self):
        return self._get("get_fine_connect_mask", self.nv, np.uint32)

    def sorted_adjacency(self):
        starts = np.zeros(self.nv + 1, np.int32)
        idx = np.zeros(self.mesh.nnz, np.int32)
        self._fn("get_sorted_adjacency", None, [C.c_void_p] * 3)(self.h, _p(starts), _p(idx))
        return starts, idx

    def stencils(self):
        n = self.stencil_num
        st = np.zeros(max(n, 1), np.dtype((np.void, 80)))
        mapped = np.zeros((max(n, 1), 5), np.int32)
        self._fn("get_stencils", None, [C.c_void_p] * 3)(self.h, _p(st), _p(mapped))
        return st[:n], mapped[:n]

    def dense_hessian(self, block: int):
        return self._get("get_dense_hessian", (96, 96), self.real, block)

    def dense_inverse(self, block: int):
        return self._get("get_dense_inverse", (96, 96), self.real, block)

    def packed_inverses(self):
        return self._get("get_packed_inverses", (self.total_clusters // 32, 4656), self.real)

    def mapped_r(self):
        return self._get("get_mapped_r", (self.total_clusters, 3), self.real)

    def mapped_z(self):
        return self._get("get_mapped_z", (self.total_clusters, 3), self.real)
