"""TEST INFRASTRUCTURE ONLY — ctypes view of oracle/_ref/libmas_ref*.so.

The shared object is the reference's own SeSchwarzPreconditioner.cpp compiled
by oracle/build_ref.sh behind oracle/ref_harness.cpp.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path(q5fix: bool = False) -> str:
    return os.path.join(_HERE, "_ref", "libmas_ref_q5fix.so" if q5fix else "libmas_ref.so")


def available(q5fix: bool = False) -> bool:
    return os.path.exists(lib_path(q5fix))


_libs = {}


def _load(q5fix: bool):
    if q5fix in _libs:
        return _libs[q5fix]
    # RTLD_LOCAL: the two variants define the same symbols
    lib = C.CDLL(lib_path(q5fix), mode=os.RTLD_LOCAL | os.RTLD_NOW)
    vp, ip, fp, up = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_uint)
    lib.ref_create.restype = vp
    lib.ref_destroy.argtypes = [vp]
    lib.ref_morton_encode.restype = C.c_ulonglong
    lib.ref_morton_encode.argtypes = [C.c_float] * 3
    lib.ref_allocate.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    lib.ref_prepare.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.c_uint, C.c_uint, C.c_uint]
    lib.ref_apply.argtypes = [vp, vp, vp]
    for name in ("ref_num_level", "ref_total_sz", "ref_total_clusters", "ref_stencil_num", "ref_max_neighbours"):
        getattr(lib, name).argtypes = [vp]
        getattr(lib, name).restype = C.c_int
    lib.ref_get_aabb.argtypes = [vp, vp, vp]
    lib.ref_get_level_size.argtypes = [vp, vp]
    lib.ref_get_morton.argtypes = [vp, vp]
    lib.ref_get_sorted_get_original.argtypes = [vp, vp]
    lib.ref_get_original_get_sorted.argtypes = [vp, vp]
    lib.ref_get_going_next.argtypes = [vp, vp, C.c_int]
    lib.ref_get_coarse_tables.argtypes = [vp, vp]
    lib.ref_get_coarse_space_table.argtypes = [vp, C.c_int, vp]
    lib.ref_get_fine_connect_mask.argtypes = [vp, vp]
    lib.ref_get_mapped_neighbors.argtypes = [vp, vp, vp]
    lib.ref_get_stencils.argtypes = [vp, vp, vp]
    lib.ref_get_dense_hessian.argtypes = [vp, C.c_int, vp]
    lib.ref_get_dense_inverse.argtypes = [vp, C.c_int, vp]
    lib.ref_get_mapped_r.argtypes = [vp, vp, C.c_int]
    lib.ref_get_mapped_z.argtypes = [vp, vp, C.c_int]
    lib.ref_set_threads.argtypes = [C.c_int]
    lib.ref_get_threads.restype = C.c_int
    lib.ref_sizeof.argtypes = [C.c_int]
    lib.ref_sizeof.restype = C.c_int
    _libs[q5fix] = lib
    return lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def morton_encode(x: float, y: float, z: float) -> int:
    return int(_load(False).ref_morton_encode(x, y, z))


def sizeof(what: int) -> int:
    return int(_load(False).ref_sizeof(what))


class RefPreconditioner:
    """The reference class behind its three public calls (h:56-63)."""

    def __init__(self, threads: int = 1, q5fix: bool = False):
        self.lib = _load(q5fix)
        self.lib.ref_set_threads(threads)
        self.h = self.lib.ref_create()
        self.mesh = None

    def close(self):
        if self.h:
            self.lib.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_threads(self, n: int):
        self.lib.ref_set_threads(n)

    def allocate(self, mesh):
        self.mesh = mesh
        self._keep = [np.ascontiguousarray(mesh.positions, np.float32),
                      np.ascontiguousarray(mesh.edges, np.int32),
                      np.ascontiguousarray(mesh.faces, np.int32),
                      np.ascontiguousarray(mesh.nbr_starts, np.int32),
                      np.ascontiguousarray(mesh.nbr_idx, np.int32)]
        k = self._keep
        self.lib.ref_allocate(self.h, mesh.nv, mesh.ne, mesh.nf, _p(k[0]), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]))

    def prepare(self, mesh=None):
        m = mesh or self.mesh
        dummy = np.zeros(64, np.uint8)
        ef = m.ef if m.ef.size else dummy
        ee = m.ee if m.ee.size else dummy
        vf = m.vf if m.vf.size else dummy
        self.lib.ref_prepare(self.h, _p(m.diag), _p(m.offdiag), _p(m.nbr_starts), _p(ef), _p(ee), _p(vf),
                             m.ef_total, m.ee_total, m.vf_total)

    def apply(self, r: np.ndarray, out=None) -> np.ndarray:
        r = np.ascontiguousarray(r, np.float32)
        z = np.zeros_like(r) if out is None else out
        self.lib.ref_apply(self.h, _p(z), _p(r))
        return z

    # ---- introspection
    @property
    def nv(self):
        return self.mesh.nv

    @property
    def num_level(self):
        return self.lib.ref_num_level(self.h)

    @property
    def total_sz(self):
        return self.lib.ref_total_sz(self.h)

    @property
    def total_clusters(self):
        return self.lib.ref_total_clusters(self.h)

    @property
    def stencil_num(self):
        return self.lib.ref_stencil_num(self.h)

    def level_size(self):
        out = np.zeros((self.num_level + 1, 2), np.int32)
        self.lib.ref_get_level_size(self.h, _p(out))
        return out

    def aabb(self):
        lo, hi = np.zeros(4, np.float32), np.zeros(4, np.float32)
        self.lib.ref_get_aabb(self.h, _p(lo), _p(hi))
        return lo, hi

    def morton(self):
        out = np.zeros(self.nv, np.uint64)
        self.lib.ref_get_morton(self.h, _p(out))
        return out

    def sorted_get_original(self):
        out = np.zeros(self.nv, np.int32)
        self.lib.ref_get_sorted_get_original(self.h, _p(out))
        return out

    def original_get_sorted(self):
        out = np.zeros(self.nv, np.int32)
        self.lib.ref_get_original_get_sorted(self.h, _p(out))
        return out

    def going_next(self, count=None):
        count = self.total_clusters if count is None else count
        count = min(count, self.num_level * self.nv)
        out = np.zeros(count, np.int32)
        self.lib.ref_get_going_next(self.h, _p(out), count)
        return out

    def coarse_tables(self):
        out = np.zeros((self.nv, 4), np.int32)
        self.lib.ref_get_coarse_tables(self.h, _p(out))
        return out

    def coarse_space_table(self, level: int):
        out = np.zeros(self.nv, np.int32)
        self.lib.ref_get_coarse_space_table(self.h, level, _p(out))
        return out

    def fine_connect_mask(self):
        out = np.zeros(self.nv, np.uint32)
        self.lib.ref_get_fine_connect_mask(self.h, _p(out))
        return out

    def mapped_neighbors(self):
        rows = self.lib.ref_max_neighbours(self.h)
        num = np.zeros(self.nv, np.int32)
        table = np.zeros((rows, self.nv), np.int32)
        self.lib.ref_get_mapped_neighbors(self.h, _p(num), _p(table))
        return num, table

    def stencils(self):
        n = self.stencil_num
        st = np.zeros(max(n, 1), np.dtype((np.void, 80)))
        mapped = np.zeros((max(n, 1), 5), np.int32)
        self.lib.ref_get_stencils(self.h, _p(st), _p(mapped))
        return st[:n], mapped[:n]

    def dense_hessian(self, block: int):
        out = np.zeros((96, 96), np.float32)
        self.lib.ref_get_dense_hessian(self.h, block, _p(out))
        return out

    def dense_inverse(self, block: int):
        out = np.zeros((96, 96), np.float32)
        self.lib.ref_get_dense_inverse(self.h, block, _p(out))
        return out

    def mapped_r(self, count=None):
        count = self.total_clusters if count is None else count
        out = np.zeros((count, 4), np.float32)
        self.lib.ref_get_mapped_r(self.h, _p(out), count)
        return out

    def mapped_z(self, count=None):
        count = self.total_clusters if count is None else count
        out = np.zeros((count, 4), np.float32)
        self.lib.ref_get_mapped_z(self.h, _p(out), count)
        return out
