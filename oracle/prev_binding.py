"""TEST INFRASTRUCTURE ONLY — ctypes view of oracle/_ref/libmas_prev.so: the reference's OLDER variant
(SeSchwarzPreconditionerPreviousVersion.h, header-only) behind oracle/prev_harness.cpp, used as a second cross-check
of z.  Only tests/ may import this."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libmas_prev.so")
_lib = None


def available() -> bool:
    return os.path.exists(LIB)


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        vp = C.c_void_p
        lib.prev_create.restype = vp
        lib.prev_destroy.argtypes = [vp]
        lib.prev_allocate.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp]
        lib.prev_prepare.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.c_uint, C.c_uint, C.c_uint]
        lib.prev_apply.argtypes = [vp, vp, vp]
        lib.prev_get_sorted_get_original.argtypes = [vp, vp]
        lib.prev_total_clusters.argtypes = [vp]
        lib.prev_total_clusters.restype = C.c_int
        lib.prev_set_threads.argtypes = [C.c_int]
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class PrevPreconditioner:
    """AllocatePrecoditioner / PreparePreconditioner / Preconditioning of the previous-version class."""

    def __init__(self, threads: int = 1):
        self.lib = _load()
        self.lib.prev_set_threads(threads)
        self.h = self.lib.prev_create()
        self.nv = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.prev_destroy(self.h)
            self.h = None

    def setup(self, mesh):
        self._keep = [np.ascontiguousarray(a) for a in (mesh.positions, mesh.edges, mesh.faces, mesh.nbr_starts, mesh.nbr_idx,
                                                        mesh.diag, mesh.offdiag)]
        pos, edges, faces, starts, idx, diag, off = self._keep
        self.nv = mesh.nv
        self.lib.prev_allocate(self.h, mesh.nv, mesh.ne, mesh.nf, _p(pos), _p(edges), _p(faces), _p(starts), _p(idx))
        ef, ee, vf = (np.frombuffer(a.tobytes(), np.uint8) if a.size else None for a in (mesh.ef, mesh.ee, mesh.vf))
        self._keep += [ef, ee, vf]
        self.lib.prev_prepare(self.h, _p(diag), _p(off), _p(starts), _p(ef), _p(ee), _p(vf), mesh.ef_total, mesh.ee_total,
                              mesh.vf_total)
        return self

    def apply(self, r):
        r = np.ascontiguousarray(r, np.float32)
        z = np.zeros_like(r)
        self.lib.prev_apply(self.h, _p(z), _p(r))
        return z

    def sorted_get_original(self):
        out = np.zeros(self.nv, np.int32)
        self.lib.prev_get_sorted_get_original(self.h, _p(out))
        return out

    @property
    def total_clusters(self):
        return int(self.lib.prev_total_clusters(self.h))
