"""Deterministic synthetic inputs for the MAS preconditioner path.

The reference ships no mesh/Hessian assembly and no collision detection
(SURVEY §0); the caller owns them.  These generators play the caller for the
five BASELINE.json configs.  Layouts follow the reference's boundary types:

* positions  float32 [nv,4]  xyzw, w = 0          (SeVec3fSimd, SeVectorSimd.h:45-103)
* 3x3 blocks float32 [n,9]   column-major m[3j+i]  (SeMatrix3f, SeMatrix.h:681-682)
* adjacency  int32 starts[nv+1], idx[nnz]          (SeCsr<int>, SeCsr.h:35-173)
* edges/faces int32 [n,4]                          (Int4, SeSchwarzPreconditioner.h:48-49)
* EfSet/EeSet/VfSet 48-byte records                (SeCollisionElements.h:33-58)
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

EF_DTYPE = np.dtype(
    {"names": ["eId", "fId", "stiff", "bary", "normal"],
     "formats": ["<i4", "<i4", "<f4", ("<f4", 3), ("<f4", 4)],
     "offsets": [0, 4, 8, 12, 32], "itemsize": 48})
EE_DTYPE = np.dtype(
    {"names": ["eId0", "eId1", "stiff", "bary", "normal"],
     "formats": ["<i4", "<i4", "<f4", ("<f4", 2), ("<f4", 4)],
     "offsets": [0, 4, 8, 16, 32], "itemsize": 48})
# `pad` is the float at byte 24 that the reference reads as m_bary[2] (cpp:399)
VF_DTYPE = np.dtype(
    {"names": ["vId", "fId", "stiff", "bary", "pad", "normal"],
     "formats": ["<i4", "<i4", "<f4", ("<f4", 2), "<f4", ("<f4", 4)],
     "offsets": [0, 4, 8, 16, 24, 32], "itemsize": 48})
STENCIL_DTYPE = np.dtype(
    {"names": ["n", "nFirst", "index", "weight", "stiff", "direction"],
     "formats": ["<i4", "<i4", ("<i4", 5), ("<f4", 5), "<f4", ("<f4", 4)],
     "offsets": [0, 4, 8, 28, 48, 64], "itemsize": 80})


@dataclass
class Mesh:
    """Everything AllocatePrecoditioner + PreparePreconditioner consume."""
    name: str
    nv: int
    positions: np.ndarray          # [nv,4] f32
    nbr_starts: np.ndarray         # [nv+1] i32
    nbr_idx: np.ndarray            # [nnz] i32
    diag: np.ndarray               # [nv,9] f32 column-major 3x3
    offdiag: np.ndarray            # [nnz,9] f32 column-major 3x3
    edges: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), np.int32))
    faces: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), np.int32))
    ef: np.ndarray = field(default_factory=lambda: np.zeros(0, EF_DTYPE))
    ee: np.ndarray = field(default_factory=lambda: np.zeros(0, EE_DTYPE))
    vf: np.ndarray = field(default_factory=lambda: np.zeros(0, VF_DTYPE))
    ef_total: int = 0
    ee_total: int = 0
    vf_total: int = 0

    @property
    def nnz(self) -> int:
        return int(self.nbr_idx.shape[0])

    @property
    def ne(self) -> int:
        return int(self.edges.shape[0])

    @property
    def nf(self) -> int:
        return int(self.faces.shape[0])


def _csr_from_edge_sequence(nv: int, a: np.ndarray, b: np.ndarray):
    """adj[a].push(b); adj[b].push(a) for each edge in sequence order."""
    m = a.shape[0]
    src = np.empty(2 * m, np.int64)
    dst = np.empty(2 * m, np.int64)
    src[0::2], dst[0::2] = a, b
    src[1::2], dst[1::2] = b, a
    order = np.argsort(src, kind="stable")          # time order is the array order
    src, dst = src[order], dst[order]
    counts = np.bincount(src, minlength=nv)
    starts = np.zeros(nv + 1, np.int64)
    np.cumsum(counts, out=starts[1:])
    return starts.astype(np.int32), dst.astype(np.int32), src.astype(np.int32)


def _spring_hessian(positions, starts, idx, src, k, m, skew=0.0):
    """K = k d d^T + 0.1 k I per directed edge; offdiag = -K; diag = m I + sum K.

    `skew` adds an antisymmetric part S to the off-diagonal blocks
    (A(v,u) = -(K+S), A(u,v) = -(K+S)^T) to exercise the column-major
    convention with non-symmetric 3x3 blocks (FEM-like fill)."""
    nv = positions.shape[0]
    p = positions[:, :3].astype(np.float32)
    d = p[idx] - p[src]
    ln = np.sqrt((d * d).sum(1, dtype=np.float32)).astype(np.float32)
    d = (d / ln[:, None]).astype(np.float32)
    K = (np.float32(k) * d[:, :, None] * d[:, None, :]).astype(np.float32)
    K += (np.float32(0.1 * k) * np.eye(3, dtype=np.float32))[None]
    off = -K                                         # [nnz, i, j]
    if skew != 0.0:
        lo = np.minimum(src, idx).astype(np.int64)
        hi = np.maximum(src, idx).astype(np.int64)
        h = ((lo * 2654435761 + hi * 40503) % 1000).astype(np.float32) / np.float32(1000.0)
        s = (np.float32(skew * k) * (h - np.float32(0.5))).astype(np.float32)
        sign = np.where(src < idx, np.float32(1), np.float32(-1)).astype(np.float32)
        S = np.zeros_like(K)
        S[:, 0, 1] = s * sign; S[:, 1, 0] = -s * sign
        S[:, 0, 2] = -0.5 * s * sign; S[:, 2, 0] = 0.5 * s * sign
        S[:, 1, 2] = 0.25 * s * sign; S[:, 2, 1] = -0.25 * s * sign
        off = off - S
    diag = np.zeros((nv, 3, 3), np.float32)
    np.add.at(diag, src, K)
    diag += (np.float32(m) * np.eye(3, dtype=np.float32))[None]
    # column-major: m[3j+i] = M(i,j)  => store transpose flattened
    off_cm = np.ascontiguousarray(off.transpose(0, 2, 1)).reshape(-1, 9)
    diag_cm = np.ascontiguousarray(diag.transpose(0, 2, 1)).reshape(-1, 9)
    return diag_cm.astype(np.float32), off_cm.astype(np.float32)


def cloth(n: int, k: float = 1000.0, m: float = 1.0, with_topology: bool = False,
          skew: float = 0.0, spacing: float = 0.01) -> Mesh:
    """Planar N x N cloth, 8-neighbour springs (BASELINE.md §3 recipe)."""
    return cloth_rect(n, n, k, m, with_topology, skew, spacing)


def cloth_rect(nx: int, ny: int, k: float = 1000.0, m: float = 1.0, with_topology: bool = False,
               skew: float = 0.0, spacing: float = 0.01) -> Mesh:
    """Planar nx x ny cloth (v = j*nx + i), same springs and insertion order as the square recipe.  Power-of-two sides
    keep the Morton banks 8x4 patches (the per-axis normalisation of cpp:225 stretches the shorter side)."""
    n = nx
    nv = nx * ny
    v = np.arange(nv, dtype=np.int64)
    i, j = v % n, v // n
    positions = np.zeros((nv, 4), np.float32)
    positions[:, 0] = (np.float32(spacing) * i.astype(np.float32))
    positions[:, 1] = (np.float32(spacing) * j.astype(np.float32))
    right, down = i + 1 < nx, j + 1 < ny
    # per v, in order: (v,v+1), (v,v+N), (v,v+N+1), (v+1,v+N)
    a = np.stack([v, v, v, v + 1], 1)
    b = np.stack([v + 1, v + n, v + n + 1, v + n], 1)
    ok = np.stack([right, down, right & down, right & down], 1)
    a, b = a[ok], b[ok]                               # row-major boolean mask keeps scan order
    starts, idx, src = _csr_from_edge_sequence(nv, a, b)
    diag, off = _spring_hessian(positions, starts, idx, src, k, m, skew)
    mesh = Mesh(f"cloth{nx}x{ny}", nv, positions, starts, idx, diag, off)
    if with_topology:
        mesh.edges, mesh.faces = _cloth_topology(nx, ny)
    return mesh


def cloth_rect_device(nx: int, ny: int, device, k: float = 1000.0, m: float = 1.0, spacing: float = 0.01):
    """The same mesh as cloth_rect(nx, ny, k, m), generated with torch ON `device` (bit-identical arrays; checked by
    tests/test_synth.py): multi-million-vertex meshes for the sharded bench without a 13 GB / 45 s numpy pass per rank.
    Returns a Mesh whose array fields are torch tensors on `device`.

    adj[u] in insertion order of the square recipe is the fixed candidate list
    u-N-1, u-N, u-N+1, u-1, u+N-1, u+1, u+N, u+N+1 filtered by the grid bounds."""
    import torch
    nv = nx * ny
    v = torch.arange(nv, dtype=torch.int64, device=device)
    i, j = v % nx, v // nx
    positions = torch.zeros((nv, 4), dtype=torch.float32, device=device)
    positions[:, 0] = torch.tensor(spacing, dtype=torch.float32, device=device) * i.to(torch.float32)
    positions[:, 1] = torch.tensor(spacing, dtype=torch.float32, device=device) * j.to(torch.float32)
    left, right, up, down = i >= 1, i + 1 < nx, j >= 1, j + 1 < ny
    offs = [-nx - 1, -nx, -nx + 1, -1, nx - 1, 1, nx, nx + 1]
    oks = [left & up, up, up & right, left, left & down, right, down, right & down]
    valid = torch.stack(oks, 1)                                          # [nv, 8]
    cand = torch.stack([v + o for o in offs], 1)
    starts = torch.zeros(nv + 1, dtype=torch.int64, device=device)
    torch.cumsum(valid.sum(1), 0, out=starts[1:])
    idx = cand[valid]                                                    # row-major mask keeps the candidate order
    del cand
    p = positions[:, :3]
    kk = torch.tensor(k, dtype=torch.float32, device=device)
    eye = torch.eye(3, dtype=torch.float32, device=device)
    kdiag = torch.tensor(0.1 * k, dtype=torch.float32, device=device) * eye
    diag = torch.zeros((nv, 3, 3), dtype=torch.float32, device=device)
    off_full = torch.empty((nv, 8, 3, 3), dtype=torch.float32, device=device)
    for c, (o, ok) in enumerate(zip(offs, oks)):
        dst = torch.where(ok, v + o, v)
        d = p[dst] - p
        ln = torch.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
        ln = torch.where(ok, ln, torch.ones_like(ln))
        d = d / ln[:, None]
        K = (kk * d[:, :, None]) * d[:, None, :]
        K = K + kdiag[None]
        K = torch.where(ok[:, None, None], K, torch.zeros_like(K))
        diag = diag + K
        off_full[:, c] = -K
    diag = diag + (torch.tensor(m, dtype=torch.float32, device=device) * eye)[None]
    off = off_full[valid]                                                # [nnz, 3, 3]
    del off_full
    off_cm = off.transpose(1, 2).contiguous().reshape(-1, 9)
    diag_cm = diag.transpose(1, 2).contiguous().reshape(-1, 9)
    return Mesh(f"cloth{nx}x{ny}", nv, positions, starts.to(torch.int32), idx.to(torch.int32), diag_cm, off_cm)


def _cloth_topology(n: int, ny: int = 0):
    """Triangles (v,v+1,v+N),(v+1,v+N+1,v+N) and their unique edges as Int4 rows.
    Only [0],[1] of an edge and [0..2] of a face are read (cpp:338-342)."""
    ny = ny or n
    v = np.arange(n * ny, dtype=np.int64)
    i, j = v % n, v // n
    q = v[(i + 1 < n) & (j + 1 < ny)]
    f0 = np.stack([q, q + 1, q + n, np.zeros_like(q)], 1)
    f1 = np.stack([q + 1, q + n + 1, q + n, np.zeros_like(q)], 1)
    faces = np.empty((2 * q.shape[0], 4), np.int64)
    faces[0::2], faces[1::2] = f0, f1
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], 0)
    e = np.sort(e, 1)
    e = np.unique(e, axis=0)
    edges = np.zeros((e.shape[0], 4), np.int64)
    edges[:, :2] = e
    edges[:, 2:] = -1
    return edges.astype(np.int32), faces.astype(np.int32)


def add_collisions(mesh: Mesh, n_ef: int, n_ee: int, n_vf: int, seed: int = 7,
                   local_fraction: float = 0.5, invalid_fraction: float = 0.02) -> Mesh:
    """Synthetic EF/EE/VF stencils laid out as the reference literally reads them.

    Q2 (cpp:357, 383): eeSets/vfSets are indexed by the GLOBAL stencil index, so
    all three arrays get ef+ee+vf records and each kind is written at its
    global slot.  Q3 (cpp:399): VfSet.m_bary[2] reads the padding float at
    byte 24; we store x+y there so the weight is the intended -(1-x-y).
    `local_fraction` of the pairs are between nearby primitives so that
    BuildCollisionConnection (cpp:514-563) actually fires; a few records get
    negative ids to exercise the skip at cpp:330/359/385."""
    assert mesh.ne > 0 and mesh.nf > 0, "need edges/faces (with_topology=True)"
    rng = np.random.RandomState(seed)
    total = n_ef + n_ee + n_vf
    ef = np.zeros(total, EF_DTYPE)
    ee = np.zeros(total, EE_DTYPE)
    vf = np.zeros(total, VF_DTYPE)

    def unit(k):
        x = rng.normal(size=(k, 3)).astype(np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True).astype(np.float32)
        out = np.zeros((k, 4), np.float32)
        out[:, :3] = x
        return out

    def near(base_count, other_count, k):
        """index into `other` close (in array order) to a random base index"""
        base = rng.randint(0, base_count, size=k)
        far = rng.randint(0, other_count, size=k)
        scaled = (base.astype(np.float64) * other_count / base_count).astype(np.int64)
        loc = np.clip(scaled + rng.randint(-3, 4, size=k), 0, other_count - 1)
        use_local = rng.uniform(size=k) < local_fraction
        return base, np.where(use_local, loc, far)

    e, f = near(mesh.ne, mesh.nf, n_ef)
    s = slice(0, n_ef)
    ef["eId"][s], ef["fId"][s] = e, f
    ef["stiff"][s] = rng.uniform(10, 1000, n_ef)
    ef["bary"][s] = np.stack([rng.uniform(0, 1, n_ef), rng.uniform(0, .5, n_ef), rng.uniform(0, .5, n_ef)], 1)
    ef["normal"][s] = unit(n_ef)

    e0, e1 = near(mesh.ne, mesh.ne, n_ee)
    s = slice(n_ef, n_ef + n_ee)
    ee["eId0"][s], ee["eId1"][s] = e0, e1
    ee["stiff"][s] = rng.uniform(10, 1000, n_ee)
    ee["bary"][s] = rng.uniform(0, 1, size=(n_ee, 2))
    ee["normal"][s] = unit(n_ee)

    vv, ff = near(mesh.nv, mesh.nf, n_vf)
    s = slice(n_ef + n_ee, total)
    vf["vId"][s], vf["fId"][s] = vv, ff
    vf["stiff"][s] = rng.uniform(10, 1000, n_vf)
    bary = rng.uniform(0, .5, size=(n_vf, 2)).astype(np.float32)
    vf["bary"][s] = bary
    vf["pad"][s] = bary[:, 0] + bary[:, 1]
    vf["normal"][s] = unit(n_vf)

    if invalid_fraction > 0:
        bad = rng.uniform(size=total) < invalid_fraction
        ef["eId"][bad[:total] & (np.arange(total) < n_ef)] = -1
        ee["eId1"][bad & (np.arange(total) >= n_ef) & (np.arange(total) < n_ef + n_ee)] = -1
        vf["fId"][bad & (np.arange(total) >= n_ef + n_ee)] = -1

    mesh.ef, mesh.ee, mesh.vf = ef, ee, vf
    mesh.ef_total, mesh.ee_total, mesh.vf_total = n_ef, n_ee, n_vf
    mesh.name += f"+coll{total}"
    return mesh


def tet_cube(nx: int, ny: int, nz: int, k: float = 1000.0, m: float = 1.0,
             skew: float = 0.05, spacing: float = 0.01) -> Mesh:
    """Lattice cube with the 6-tet Kuhn split (<=14 neighbours per vertex).

    Kuhn edges of a cell are the 7 offsets with non-negative components
    (1,0,0),(0,1,0),(0,0,1),(1,1,0),(1,0,1),(0,1,1),(1,1,1); off-diagonal
    blocks get a small antisymmetric part (arbitrary FEM-like fill)."""
    nv = nx * ny * nz
    v = np.arange(nv, dtype=np.int64)
    i, j, l = v % nx, (v // nx) % ny, v // (nx * ny)
    positions = np.zeros((nv, 4), np.float32)
    positions[:, 0] = np.float32(spacing) * i.astype(np.float32)
    positions[:, 1] = np.float32(spacing) * j.astype(np.float32)
    positions[:, 2] = np.float32(spacing) * l.astype(np.float32)
    offs = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1)]
    a_cols, b_cols, ok_cols = [], [], []
    for (di, dj, dl) in offs:
        ok = (i + di < nx) & (j + dj < ny) & (l + dl < nz)
        a_cols.append(v)
        b_cols.append(v + di + dj * nx + dl * nx * ny)
        ok_cols.append(ok)
    a, b, ok = np.stack(a_cols, 1), np.stack(b_cols, 1), np.stack(ok_cols, 1)
    a, b = a[ok], b[ok]
    starts, idx, src = _csr_from_edge_sequence(nv, a, b)
    diag, off = _spring_hessian(positions, starts, idx, src, k, m, skew)
    return Mesh(f"tet{nx}x{ny}x{nz}", nv, positions, starts, idx, diag, off)


def from_edges(positions_xyz: np.ndarray, a: np.ndarray, b: np.ndarray, k: float = 1000.0, m: float = 1.0,
               name: str = "custom") -> Mesh:
    """A mass-spring mesh over arbitrary undirected edges (a[i], b[i]) in insertion order (both directions pushed, as in
    the cloth recipe); vertices may have no edges at all.  Edge-case meshes for the parity tests."""
    nv = int(positions_xyz.shape[0])
    positions = np.zeros((nv, 4), np.float32)
    positions[:, :3] = np.asarray(positions_xyz, np.float32)[:, :3]
    a = np.asarray(a, np.int64).reshape(-1)
    b = np.asarray(b, np.int64).reshape(-1)
    starts, idx, src = _csr_from_edge_sequence(nv, a, b)
    if idx.shape[0]:
        diag, off = _spring_hessian(positions, starts, idx, src, k, m)
    else:
        diag = np.tile((np.float32(m) * np.eye(3, dtype=np.float32)).T.reshape(1, 9), (nv, 1)).astype(np.float32)
        off = np.zeros((0, 9), np.float32)
    return Mesh(name, nv, positions, starts, idx, diag, off)


def chain(n: int, spacing: float = 0.01) -> Mesh:
    """n vertices on a slightly bent line, consecutive ones connected (n = 1: a single free vertex)."""
    t = np.arange(n, dtype=np.float32)
    pos = np.stack([np.float32(spacing) * t, np.float32(0.3 * spacing) * np.sin(t), np.float32(0.2 * spacing) * np.cos(2 * t)], 1)
    v = np.arange(max(n - 1, 0), dtype=np.int64)
    return from_edges(pos, v, v + 1, name=f"chain{n}")


def cloth_with_isolated_vertices(n: int = 20, extra: int = 7) -> Mesh:
    """An n x n cloth plus `extra` vertices without any edge, dropped into the sheet's bounding box (they sort into the
    cloth's banks and must come out as one-vertex clusters)."""
    base = cloth(n)
    rng = np.random.RandomState(5)
    lo, hi = base.positions[:, :3].min(0), base.positions[:, :3].max(0)
    iso = rng.uniform(0.1, 0.9, size=(extra, 3)).astype(np.float32) * (hi - lo) + lo
    pos = np.concatenate([base.positions[:, :3], iso], 0)
    src = np.repeat(np.arange(base.nv), np.diff(base.nbr_starts))
    keep = src < base.nbr_idx                       # every undirected edge once, in first-appearance order
    return from_edges(pos, src[keep], base.nbr_idx[keep], name=f"cloth{n}+iso{extra}")


def _undirected_edges(mesh: Mesh):
    """Every undirected edge of `mesh` once, in first-appearance order."""
    src = np.repeat(np.arange(mesh.nv), np.diff(mesh.nbr_starts))
    keep = src < mesh.nbr_idx
    return src[keep].astype(np.int64), mesh.nbr_idx[keep].astype(np.int64)


def stacked_cloth(n: int = 24, layers: int = 2, k: float = 1000.0) -> Mesh:
    """`layers` n x n sheets lying exactly on top of each other (a folded garment at rest): vertex (i,j) of every layer has
    the same position, hence the same Morton code, so the order inside each group of coincident vertices is decided by the
    tie rule alone (ascending original index; the reference's std::sort leaves it unspecified, cpp:238-243).  Layers are
    stitched with springs between (i,j) of one layer and (i+1,j) of the next (never zero length)."""
    base = cloth(n, k=k)
    a0, b0 = _undirected_edges(base)
    nv1 = base.nv
    pos = np.concatenate([base.positions[:, :3]] * layers, 0)
    a = [a0 + l * nv1 for l in range(layers)]
    b = [b0 + l * nv1 for l in range(layers)]
    v = np.arange(nv1, dtype=np.int64)
    stitch = v[(v % n) + 1 < n]
    for l in range(layers - 1):
        a.append(stitch + l * nv1)
        b.append(stitch + 1 + (l + 1) * nv1)
    return from_edges(pos, np.concatenate(a), np.concatenate(b), k=k, name=f"stacked{layers}x{n}")


def cloth_with_duplicate_edges(n: int = 24, every: int = 5) -> Mesh:
    """An n x n cloth whose adjacency lists every `every`-th spring twice (parallel springs / a caller that does not
    deduplicate): the neighbour lists then hold repeated indices, each with its own off-diagonal block, and the blocks
    must add up (the reference's += at cpp:1292-1298)."""
    base = cloth(n)
    a0, b0 = _undirected_edges(base)
    dup = np.arange(0, a0.shape[0], every)
    return from_edges(base.positions[:, :3], np.concatenate([a0, a0[dup]]), np.concatenate([b0, b0[dup]]),
                      name=f"cloth{n}+dup")


def random_cloud(n: int = 1500, k_nearest: int = 5, seed: int = 3, box=(1.0, 0.7, 0.4)) -> Mesh:
    """An irregular 3-D mesh: n points uniform in a box, every point connected to its `k_nearest` nearest neighbours
    (symmetrised, each undirected edge once, ascending (a, b)).  Unlike the lattices, degrees vary from vertex to vertex,
    Morton banks cut through the connectivity anywhere and several components per bank are the rule — the shape of a
    mesher's tetrahedral output.  scipy is test-side tooling only."""
    from scipy.spatial import cKDTree
    rng = np.random.RandomState(seed)
    pos = (rng.uniform(0.0, 1.0, size=(n, 3)) * np.asarray(box)).astype(np.float32)
    _, nbr = cKDTree(pos.astype(np.float64)).query(pos.astype(np.float64), k=k_nearest + 1)
    a = np.repeat(np.arange(n), k_nearest)
    b = nbr[:, 1:].reshape(-1)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    pairs = np.unique(np.stack([lo, hi], 1), axis=0)
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    return from_edges(pos, pairs[:, 0], pairs[:, 1], name=f"cloud{n}k{k_nearest}s{seed}")


def rippled_cloth(n: int = 64, amplitude: float = 1e-4, k: float = 1000.0) -> Mesh:
    """The n x n cloth with a faint out-of-plane ripple.  The Morton code normalises every axis by its own extent
    (cpp:225), so a ripple of 1e-4 gets the full weight of the z bits: banks stop being 8 x 4 patches, each holds several
    connected components and level 1 is 4-5 times larger than on the flat sheet (554 instead of 128 nodes at n = 64).  At
    512^2 the reference overruns its fixed 1.5x allocation on such input; here buffers follow the real counts."""
    base = cloth(n, k=k)
    a, b = _undirected_edges(base)
    pos = base.positions[:, :3].copy()
    v = np.arange(base.nv)
    pos[:, 2] = (np.float32(amplitude) * np.sin(0.7 * (v % n)) * np.cos(0.45 * (v // n))).astype(np.float32)
    return from_edges(pos, a, b, k=k, name=f"rippled{n}")


def folded_cloth(nx: int = 64, ny: int = 64, gap: float = 0.004, wobble: float = 0.006, k: float = 1000.0, m: float = 1.0,
                 spacing: float = 0.01) -> Mesh:
    """An nx x ny sheet folded in half along x: the second half lies back over the first at height gap + wobble sin(...),
    shifted by a fraction of a cell so that no two vertices share their (x, y).  Where the wobble exceeds the gap the layers
    interpenetrate (edges pierce faces: EF stencils); elsewhere they are within a fraction of an edge length of each other
    (VF / EE stencils).  Geometry for collide.proximity_stencils; springs and topology as in cloth_rect."""
    mesh = cloth_rect(nx, ny, k, m, with_topology=True, spacing=spacing)
    s = np.float32(spacing)
    i = (np.arange(mesh.nv) % nx).astype(np.float32)
    j = (np.arange(mesh.nv) // nx).astype(np.float32)
    half = np.float32(nx // 2)
    upper = i > half
    pos = mesh.positions.copy()
    pos[upper, 0] = s * (2 * half - i[upper]) + np.float32(0.37) * s
    pos[upper, 1] = s * j[upper] + np.float32(0.21) * s
    pos[upper, 2] = np.float32(gap) + np.float32(wobble) * np.sin(np.float32(0.35) * i[upper]) * np.cos(np.float32(0.27) * j[upper])
    # the fold itself: a short ramp so that the crease column does not stretch to the full gap in one edge
    crease = i == half + 1
    pos[crease, 2] *= np.float32(0.5)
    mesh.positions = pos.astype(np.float32)
    src = np.repeat(np.arange(mesh.nv), np.diff(mesh.nbr_starts))
    mesh.diag, mesh.offdiag = _spring_hessian(mesh.positions, mesh.nbr_starts, mesh.nbr_idx, src, k, m, 0.0)
    mesh.name = f"folded{nx}x{ny}"
    return mesh


def dust(n: int = 6000, seed: int = 2, m: float = 1.0) -> Mesh:
    """n free particles, no edges at all: nothing ever aggregates, so EVERY level keeps n one-vertex clusters (three
    times the reference's whole fixed allocation at n = 6000, Q6) and every domain matrix is m I.  Known answer without
    any reference: z = min(numLevel, 4) r / m."""
    rng = np.random.RandomState(seed)
    pos = rng.uniform(0.0, 1.0, size=(n, 3)).astype(np.float32)
    return from_edges(pos, np.zeros(0, np.int64), np.zeros(0, np.int64), m=m, name=f"dust{n}")


def residual(nv: int, seed: int = 1) -> np.ndarray:
    """r ~ U(-1,1) per component, MT19937(seed), xyz per vertex in order; w = 0."""
    rng = np.random.RandomState(seed)
    r = np.zeros((nv, 4), np.float32)
    r[:, :3] = rng.uniform(-1.0, 1.0, size=(nv, 3)).astype(np.float32)
    return r


def fnv1a_i32(values: np.ndarray) -> int:
    """h = 2166136261; h = (h ^ x) * 16777619 per int (SURVEY §8c known answers)."""
    x = np.ascontiguousarray(values).astype(np.uint32, copy=False).ravel()
    h = np.uint64(2166136261)
    # vectorising a serial hash is not possible; chunked python loop on uint32
    h = int(h)
    for val in x.tolist():
        h = ((h ^ val) * 16777619) & 0xFFFFFFFF
    return h


def config(index: int, proximity: bool = False) -> Mesh:
    """BASELINE.json configs[index].  proximity (config 1 only): the 512x512 sheet folded in half, with the EF / EE / VF
    stencils collide.proximity_stencils finds between the two layers instead of the random synthetic ones."""
    if index == 0:
        return cloth(64)
    if index == 1 and proximity:
        from .collide import proximity_stencils
        return proximity_stencils(folded_cloth(512, 512), radius=0.006)
    if index == 1:
        mesh = cloth(512, with_topology=True)
        return add_collisions(mesh, mesh.nv // 16, mesh.nv // 16, mesh.nv // 8)
    if index == 2:
        return cloth(1024)
    if index == 3:
        return tet_cube(128, 128, 64)
    if index == 4:
        return cloth(2048)
    raise ValueError(index)


def weak_scaling_cloth(n_gpus: int) -> Mesh:
    """One sharded cloth with 1,048,576 vertices PER GPU (bench.py --scaling weak): 1024x1024, 2048x1024, 2048x2048,
    4096x2048 for 1/2/4/8 GPUs; in general 1024*a x 1024*b with a*b = n_gpus, a >= b, both powers of two when possible."""
    if n_gpus < 1:
        raise ValueError(n_gpus)
    b = 1
    while (b * 2) * (b * 2) <= n_gpus and n_gpus % (b * 2) == 0:
        b *= 2
    a = n_gpus // b
    return cloth_rect(1024 * a, 1024 * b)
