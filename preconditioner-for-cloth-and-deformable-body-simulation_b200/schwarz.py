"""Host-side mirror of SE::SeSchwarzPreconditioner over the C ABI of libmas_b200.so.

Same three calls, same argument meaning and order as the reference class
(SeSchwarzPreconditioner.h:55-63) -- AllocatePrecoditioner [sic], PreparePreconditioner,
Preconditioning -- so the parity tests read like calls into the reference.  Inputs may be
numpy arrays (HOST memory, as in the reference) or torch CUDA tensors (DEVICE memory, no copies).
PyTorch is used only for device memory, streams and torch.distributed plumbing.

There is no CPU fallback: if the shared library or a B200 is missing this raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MAS_B200_LIB") or os.path.join(_HERE, "libmas_b200.so")   # override: development builds only

MAS_MEM_HOST, MAS_MEM_DEVICE = 0, 1
(OPT_PROLONG_ALL_LEVELS, OPT_APPLY_VARIANT, OPT_USE_GRAPH, OPT_TIME_KERNELS, OPT_ALIGN_CUTS, OPT_STENCIL_FIX,
 OPT_RESORT_PERIOD, OPT_HOST_PULL, OPT_INVERT_VARIANT, OPT_REGISTER_HOST, OPT_CACHE_HIERARCHY, OPT_STRICT_PUBLISH,
 OPT_PCG_PERSIST_L2) = range(13)
(INT_NUM_VERTS, INT_NUM_LEVEL, INT_TOTAL_CLUSTERS, INT_NUM_BLOCKS, INT_STENCIL_NUM, INT_NNZ, INT_APPLY_LAUNCHES,
 INT_PACKED_FLOATS_PER_BLOCK, INT_OWNED_BLOCK_BEGIN, INT_OWNED_BLOCK_END, INT_PREPARE_LAUNCHES, INT_PCG_LAUNCHES_PER_ITER,
 INT_PCG_CONVERGED, INT_PEER_ERROR, INT_ALIGNED_CUTS, INT_HOST_PULL_CHOICE, INT_HOST_BYTES_IN, INT_HOST_BYTES_OUT) = range(18)
(ARR_MORTON, ARR_SORTED_GET_ORIGINAL, ARR_ORIGINAL_GET_SORTED, ARR_GOING_NEXT, ARR_LEVEL_SIZE, ARR_FINE_CONNECT_MASK,
 ARR_COARSE_SPACE_TABLE, ARR_COARSE_TABLES, ARR_SORTED_ADJ_STARTS, ARR_SORTED_ADJ_IDX, ARR_STENCILS,
 ARR_STENCIL_INDEX_MAPPED, ARR_DENSE_INVERSE, ARR_MAPPED_R, ARR_MAPPED_Z, ARR_AABB) = range(16)

# every symbol include/mas_b200.h declares; tests check the library exports all of them
EXPORTS = [
    "mas_create", "mas_destroy", "mas_last_error", "mas_set_stream", "mas_set_option", "mas_set_partition",
    "mas_allocate", "mas_prepare", "mas_apply", "mas_prepare_begin", "mas_prepare_end", "mas_apply_begin",
    "mas_apply_end", "mas_exchange_buffer", "mas_get_int", "mas_get_array", "mas_morton_encode", "mas_get_timing",
    "mas_pcg_solve", "mas_peer_export", "mas_peer_local", "mas_peer_attach", "mas_synchronize",
]

_lib = None


class MasError(RuntimeError):
    pass


def load_library() -> C.CDLL:
    """dlopen libmas_b200.so (built in-tree by __graft_entry__.build()); fails loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MasError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i, u = C.c_void_p, C.c_int, C.c_uint
    lib.mas_create.argtypes = [C.POINTER(vp), i]
    lib.mas_destroy.argtypes = [vp]
    lib.mas_last_error.argtypes = [vp]
    lib.mas_last_error.restype = C.c_char_p
    lib.mas_set_stream.argtypes = [vp, vp]
    lib.mas_set_option.argtypes = [vp, i, i]
    lib.mas_set_partition.argtypes = [vp, i, i]
    lib.mas_allocate.argtypes = [vp, i, i, i, vp, vp, vp, vp, vp, i]
    lib.mas_prepare.argtypes = [vp, vp, vp, vp, vp, vp, vp, u, u, u, i]
    lib.mas_prepare_begin.argtypes = [vp, vp, vp, vp, vp, vp, vp, u, u, u, i]
    lib.mas_prepare_end.argtypes = [vp]
    lib.mas_apply.argtypes = [vp, vp, vp, i]
    lib.mas_apply_begin.argtypes = [vp, vp, i]
    lib.mas_apply_end.argtypes = [vp, vp, i]
    lib.mas_exchange_buffer.argtypes = [vp, i, C.POINTER(vp), C.POINTER(C.c_size_t)]
    lib.mas_get_int.argtypes = [vp, i, C.POINTER(C.c_longlong)]
    lib.mas_get_array.argtypes = [vp, i, i, vp, C.c_size_t]
    lib.mas_morton_encode.argtypes = [vp, vp, i, vp]
    lib.mas_get_timing.argtypes = [vp, i, C.POINTER(C.c_float)]
    lib.mas_peer_export.argtypes = [vp, vp]
    lib.mas_peer_local.argtypes = [vp, C.POINTER(vp)]
    lib.mas_peer_attach.argtypes = [vp, vp, C.POINTER(vp)]
    lib.mas_synchronize.argtypes = [vp]
    lib.mas_pcg_solve.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.c_float, i, i, i, C.POINTER(i), C.POINTER(C.c_float)]
    for name in EXPORTS:
        if name != "mas_last_error":
            getattr(lib, name).restype = i
    _lib = lib
    return lib


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _ptr(x, dtype=None):
    """(pointer, mem kind, keepalive) for a numpy array or a torch tensor."""
    if x is None:
        return None, None, None
    if _is_torch(x):
        if not x.is_contiguous():
            x = x.contiguous()
        kind = MAS_MEM_DEVICE if x.is_cuda else MAS_MEM_HOST
        return C.c_void_p(x.data_ptr()), kind, x
    a = np.ascontiguousarray(x) if dtype is None else np.ascontiguousarray(x, dtype)
    return a.ctypes.data_as(C.c_void_p), MAS_MEM_HOST, a


class SeSchwarzPreconditioner:
    """Drop-in for the reference class on one B200 (or one shard of a multi-GPU partition).

    Public input members keep the reference names (h:44-51): m_positions, m_edges, m_faces, m_neighbours
    (a (starts, idxs) pair, SeCsr<int> without the unused values)."""

    def __init__(self, device: int = 0, rank: int = 0, world: int = 1, stream=None):
        self.lib = load_library()
        self.h = C.c_void_p()
        rc = self.lib.mas_create(C.byref(self.h), device)
        if rc != 0:
            raise MasError(f"mas_create(device={device}) failed with {rc}: no usable sm_100 GPU (no CPU fallback)")
        self.device = device
        self.invert_variant = 0          # MAS_OPT_INVERT_VARIANT: 0 tensor cores (default), 1 FP32 CUDA cores
        self.rank, self.world = rank, world
        if world > 1:
            self._ck(self.lib.mas_set_partition(self.h, rank, world))
        if stream is not None:
            self.set_stream(stream)
        self.m_positions = None
        self.m_edges = None
        self.m_faces = None
        self.m_neighbours = None
        self._keep = []

    # ---- plumbing
    def _ck(self, rc: int):
        if rc != 0:
            raise MasError(f"libmas_b200 error {rc}: {self.lib.mas_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.mas_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream):
        """stream: torch.cuda.Stream, raw cudaStream_t integer, or None for the default stream."""
        raw = 0 if stream is None else (stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream))
        self._ck(self.lib.mas_set_stream(self.h, C.c_void_p(raw)))

    def synchronize(self):
        """Waits for the context's stream; raises if a peer exchange of a sharded apply ever timed out."""
        self._ck(self.lib.mas_synchronize(self.h))

    def set_option(self, key: int, value: int):
        self._ck(self.lib.mas_set_option(self.h, key, value))
        if key == OPT_INVERT_VARIANT:
            self.invert_variant = value

    # ---- the reference's three calls
    def AllocatePrecoditioner(self, numVerts: int, numEdges: int, numFaces: int):  # noqa: N802 (reference spelling)
        starts, idxs = self.m_neighbours
        pp, kind, k0 = _ptr(self.m_positions, np.float32 if not _is_torch(self.m_positions) else None)
        ep, _, k1 = _ptr(self.m_edges if numEdges > 0 else None, np.int32 if not _is_torch(self.m_edges) else None)
        fp, _, k2 = _ptr(self.m_faces if numFaces > 0 else None, np.int32 if not _is_torch(self.m_faces) else None)
        sp, _, k3 = _ptr(starts, np.int32 if not _is_torch(starts) else None)
        ip, _, k4 = _ptr(idxs, np.int32 if not _is_torch(idxs) else None)
        self._keep = [k0, k1, k2, k3, k4]
        self._ck(self.lib.mas_allocate(self.h, numVerts, numEdges, numFaces, pp, ep, fp, sp, ip, kind))

    def PreparePreconditioner(self, diagonal, csrOffDiagonals, csrRanges, efSets=None, eeSets=None, vfSets=None,  # noqa: N802
                              efCounts=0, eeCounts=0, vfCounts=0, phase: Optional[str] = None):
        """efCounts/eeCounts/vfCounts: the totals the reference reads from the last element of the caller's
        prefix arrays (cpp:306-308); an array is accepted too and its last element is used."""
        def total(x):
            return int(x if np.isscalar(x) else np.asarray(x).ravel()[-1])
        ef_n, ee_n, vf_n = total(efCounts), total(eeCounts), total(vfCounts)
        dp, kind, k0 = _ptr(diagonal, np.float32 if not _is_torch(diagonal) else None)
        op, _, k1 = _ptr(csrOffDiagonals, np.float32 if not _is_torch(csrOffDiagonals) else None)
        rp, _, k2 = _ptr(csrRanges, np.int32 if not _is_torch(csrRanges) else None)
        efp, _, k3 = _ptr(efSets if ef_n + ee_n + vf_n > 0 else None)
        eep, _, k4 = _ptr(eeSets if ef_n + ee_n + vf_n > 0 else None)
        vfp, _, k5 = _ptr(vfSets if ef_n + ee_n + vf_n > 0 else None)
        keep = [k0, k1, k2, k3, k4, k5]
        if phase == "begin":
            self._ck(self.lib.mas_prepare_begin(self.h, dp, op, rp, efp, eep, vfp, ef_n, ee_n, vf_n, kind))
        else:
            self._ck(self.lib.mas_prepare(self.h, dp, op, rp, efp, eep, vfp, ef_n, ee_n, vf_n, kind))
        del keep

    def prepare_end(self):
        self._ck(self.lib.mas_prepare_end(self.h))

    def Preconditioning(self, z, residual, dim: int = 0):  # noqa: N802
        """z = M^-1 residual; `dim` is accepted and ignored exactly like the reference (cpp:100)."""
        if not _is_torch(z) and not (isinstance(z, np.ndarray) and z.dtype == np.float32 and z.flags.c_contiguous):
            raise MasError("z must be a C-contiguous float32 array of xyzw vectors")
        if _is_torch(z) and not (z.is_contiguous() and str(z.dtype) == "torch.float32"):
            raise MasError("z must be a contiguous float32 tensor of xyzw vectors")
        zp, zk, z_keep = _ptr(z)
        rp, rk, r_keep = _ptr(residual, np.float32 if not _is_torch(residual) else None)
        if zk != rk:
            raise MasError("z and residual must live in the same memory space")
        self._ck(self.lib.mas_apply(self.h, zp, rp, zk))
        return z

    # multi-GPU phase split (device pointers only)
    def apply_begin(self, residual):
        rp, rk, _ = _ptr(residual)
        self._ck(self.lib.mas_apply_begin(self.h, rp, rk))

    def apply_end(self, z):
        zp, zk, _ = _ptr(z)
        self._ck(self.lib.mas_apply_end(self.h, zp, zk))

    def exchange_buffer(self, which: int):
        """(device pointer, element count) of the buffer to all-reduce between *_begin and *_end."""
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.lib.mas_exchange_buffer(self.h, which, C.byref(p), C.byref(n)))
        return int(p.value or 0), int(n.value)

    def exchange_tensor(self, which: int):
        """torch view (no copy) of the exchange buffer: float64 for prepare (0), float32 for apply (1)."""
        import torch
        ptr, n = self.exchange_buffer(which)
        dtype, np_t, size = (torch.float64, "<f8", 8) if which == 0 else (torch.float32, "<f4", 4)

        class _Wrap:
            pass
        w = _Wrap()
        w.__cuda_array_interface__ = {"shape": (n,), "typestr": np_t, "data": (ptr, False), "version": 3, "strides": None}
        t = torch.as_tensor(w, device=f"cuda:{self.device}")
        assert t.dtype == dtype and t.data_ptr() == ptr
        return t

    # ---- peer-memory exchange (multi-GPU, NVLink)
    def peer_export(self) -> bytes:
        """64-byte CUDA IPC handle of this rank's exchange arena (send it to the other processes)."""
        buf = C.create_string_buffer(64)
        self._ck(self.lib.mas_peer_export(self.h, buf))
        return buf.raw

    def peer_local(self) -> int:
        p = C.c_void_p()
        self._ck(self.lib.mas_peer_local(self.h, C.byref(p)))
        return int(p.value)

    def peer_attach(self, handles=None, pointers=None):
        """handles: list of `world` 64-byte handles in rank order; pointers: list of `world` device pointers (0/None = use
        the handle).  After this, Preconditioning() on a sharded context is a single call again."""
        hb = b"".join(handles) if handles is not None else None
        hp = C.c_char_p(hb) if hb is not None else None
        pp = None
        if pointers is not None:
            pp = (C.c_void_p * len(pointers))(*[C.c_void_p(int(p) if p else 0) for p in pointers])
        self._ck(self.lib.mas_peer_attach(self.h, hp, pp))

    @property
    def peer_error(self): return self.get_int(INT_PEER_ERROR)

    @property
    def aligned_cuts(self): return bool(self.get_int(INT_ALIGNED_CUTS))

    # ---- introspection (parity tests)
    def get_int(self, key: int) -> int:
        out = C.c_longlong()
        self._ck(self.lib.mas_get_int(self.h, key, C.byref(out)))
        return int(out.value)

    def get_array(self, key: int, shape, dtype, index: int = 0) -> np.ndarray:
        out = np.zeros(shape, dtype)
        if out.nbytes:
            self._ck(self.lib.mas_get_array(self.h, key, index, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    @property
    def nv(self): return self.get_int(INT_NUM_VERTS)
    @property
    def num_level(self): return self.get_int(INT_NUM_LEVEL)
    @property
    def total_clusters(self): return self.get_int(INT_TOTAL_CLUSTERS)
    @property
    def num_blocks(self): return self.get_int(INT_NUM_BLOCKS)
    @property
    def stencil_num(self): return self.get_int(INT_STENCIL_NUM)
    @property
    def apply_launches(self): return self.get_int(INT_APPLY_LAUNCHES)
    @property
    def prepare_launches(self): return self.get_int(INT_PREPARE_LAUNCHES)
    @property
    def owned_fine_blocks(self): return self.get_int(INT_OWNED_BLOCK_BEGIN), self.get_int(INT_OWNED_BLOCK_END)

    def morton(self): return self.get_array(ARR_MORTON, self.nv, np.uint64)
    def sorted_get_original(self): return self.get_array(ARR_SORTED_GET_ORIGINAL, self.nv, np.int32)
    def original_get_sorted(self): return self.get_array(ARR_ORIGINAL_GET_SORTED, self.nv, np.int32)
    def going_next(self): return self.get_array(ARR_GOING_NEXT, self.total_clusters, np.int32)
    def level_size(self): return self.get_array(ARR_LEVEL_SIZE, (self.num_level + 1, 2), np.int32)
    def fine_connect_mask(self): return self.get_array(ARR_FINE_CONNECT_MASK, self.nv, np.uint32)
    def coarse_space_table(self, level): return self.get_array(ARR_COARSE_SPACE_TABLE, self.nv, np.int32, level)
    def coarse_tables(self): return self.get_array(ARR_COARSE_TABLES, (self.nv, 4), np.int32)
    def aabb(self):
        a = self.get_array(ARR_AABB, 8, np.float32)
        return a[:4], a[4:]

    def sorted_adjacency(self):
        return (self.get_array(ARR_SORTED_ADJ_STARTS, self.nv + 1, np.int32),
                self.get_array(ARR_SORTED_ADJ_IDX, self.get_int(INT_NNZ), np.int32))

    def stencils(self):
        n = self.stencil_num
        return (self.get_array(ARR_STENCILS, n, np.dtype((np.void, 80))),
                self.get_array(ARR_STENCIL_INDEX_MAPPED, (n, 5), np.int32))

    def dense_inverse(self, block: int): return self.get_array(ARR_DENSE_INVERSE, (96, 96), np.float32, block)
    def mapped_r(self): return self.get_array(ARR_MAPPED_R, (self.total_clusters, 4), np.float32)
    def mapped_z(self): return self.get_array(ARR_MAPPED_Z, (self.total_clusters, 4), np.float32)

    def morton_encode(self, xyz) -> np.ndarray:
        pts = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        out = np.zeros(pts.shape[0], np.uint64)
        self._ck(self.lib.mas_morton_encode(self.h, pts.ctypes.data_as(C.c_void_p), pts.shape[0], out.ctypes.data_as(C.c_void_p)))
        return out

    def timing_ms(self, which: int) -> float:
        out = C.c_float()
        self._ck(self.lib.mas_get_timing(self.h, which, C.byref(out)))
        return float(out.value)

    # ---- convenience over a synth.Mesh
    def setup_from_mesh(self, mesh, device_inputs: bool = False):
        """AllocatePrecoditioner + PreparePreconditioner from a synth.Mesh; optionally with device-resident inputs."""
        if device_inputs:
            import torch
            dev = f"cuda:{self.device}"
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            tb = lambda a: torch.from_numpy(np.frombuffer(np.ascontiguousarray(a).tobytes(), np.uint8).copy()).to(dev)
            self.m_positions = t(mesh.positions)
            self.m_edges = t(mesh.edges) if mesh.ne else None
            self.m_faces = t(mesh.faces) if mesh.nf else None
            self.m_neighbours = (t(mesh.nbr_starts), t(mesh.nbr_idx))
            self.AllocatePrecoditioner(mesh.nv, mesh.ne, mesh.nf)
            self._dev_inputs = (t(mesh.diag), t(mesh.offdiag), t(mesh.nbr_starts),
                                tb(mesh.ef) if mesh.ef.size else None, tb(mesh.ee) if mesh.ee.size else None,
                                tb(mesh.vf) if mesh.vf.size else None)
            d = self._dev_inputs
            self.PreparePreconditioner(d[0], d[1], d[2], d[3], d[4], d[5], mesh.ef_total, mesh.ee_total, mesh.vf_total)
        else:
            self.m_positions = mesh.positions
            self.m_edges = mesh.edges
            self.m_faces = mesh.faces
            self.m_neighbours = (mesh.nbr_starts, mesh.nbr_idx)
            self.AllocatePrecoditioner(mesh.nv, mesh.ne, mesh.nf)
            self.PreparePreconditioner(mesh.diag, mesh.offdiag, mesh.nbr_starts, mesh.ef, mesh.ee, mesh.vf,
                                       mesh.ef_total, mesh.ee_total, mesh.vf_total)
        return self
