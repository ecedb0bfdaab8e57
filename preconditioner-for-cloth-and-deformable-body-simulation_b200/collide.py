"""Proximity-based producer of the EF / EE / VF collision stencils PreparePreconditioner consumes.

The reference ships no collision detection: its caller hands over `EfSet / EeSet / VfSet` arrays
(SeCollisionElements.h:33-58) and PrepareCollisionStencils (cpp:304-413) turns them into 4- or 5-vertex stencils.  The random
stencils of `synth.add_collisions` exercise that code but mean nothing geometrically; this module makes the records a
simulator would: from positions, edges and faces it finds

  VF  vertex p within `radius` of a triangle it does not belong to, closest point inside the triangle:
      bary = barycentrics of the closest point (weights of face[0], face[1]; face[2] gets 1 - b0 - b1), normal = unit(p - q)
  EE  two edges without a common vertex whose closest points are interior and within `radius`:
      bary = (weight of edge0[0], weight of edge1[0]), normal = unit(c0 - c1)
  EF  an edge that pierces a triangle it shares no vertex with:
      bary = (weight of edge[0] at the intersection, weights of face[0], face[1]), normal = face normal towards edge[0]

with a uniform grid (cell = radius + longest edge, 27-cell neighbourhoods, sorted cell keys + searchsorted), entirely in torch
tensor ops so that it runs on the GPU for the 262k-vertex BASELINE config and on the CPU for the small test meshes.  It is
workload generation on the caller's side of the boundary, not part of the preconditioner.

Records are laid out the way the reference literally reads them (Q2: all three arrays span every global stencil index, each
kind at its own slots; Q3: VfSet's padding float at byte 24 holds b0 + b1)."""
from __future__ import annotations

import numpy as np

from .synth import EE_DTYPE, EF_DTYPE, VF_DTYPE, Mesh


def _grid(points, cell, origin, dims):
    import torch
    c = torch.clamp(((points - origin) / cell).floor().to(torch.int64), min=0)
    c = torch.minimum(c, dims - 1)
    return c


def _key(c, dims):
    return (c[:, 2] * dims[1] + c[:, 1]) * dims[0] + c[:, 0]


def _candidates(query_pts, target_pts, cell, chunk=65536):
    """Yields (query index, target index) pairs whose cells are neighbours (27-neighbourhood), chunked over the queries."""
    import torch
    lo = torch.minimum(query_pts.min(0).values, target_pts.min(0).values) - cell
    hi = torch.maximum(query_pts.max(0).values, target_pts.max(0).values) + cell
    dims = ((hi - lo) / cell).ceil().to(torch.int64) + 1
    tkey = _key(_grid(target_pts, cell, lo, dims), dims)
    order = torch.argsort(tkey, stable=True)
    skey = tkey[order]
    offs = torch.tensor([(a, b, c) for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)], device=query_pts.device, dtype=torch.int64)
    for q0 in range(0, query_pts.shape[0], chunk):
        qc = _grid(query_pts[q0:q0 + chunk], cell, lo, dims)
        n = qc.shape[0]
        cells = qc[:, None, :] + offs[None, :, :]                             # [n, 27, 3]
        ok = ((cells >= 0) & (cells < dims)).all(-1)
        keys = (cells[..., 2] * dims[1] + cells[..., 1]) * dims[0] + cells[..., 0]
        keys = torch.where(ok, keys, torch.full_like(keys, -1))
        a = torch.searchsorted(skey, keys.reshape(-1))
        b = torch.searchsorted(skey, keys.reshape(-1), right=True)
        cnt = torch.where(keys.reshape(-1) >= 0, b - a, torch.zeros_like(a))
        total = int(cnt.sum())
        if total == 0:
            continue
        owner = torch.repeat_interleave(torch.arange(n * 27, device=cnt.device), cnt)
        first = torch.cumsum(cnt, 0) - cnt
        within = torch.arange(total, device=cnt.device) - first[owner]
        yield q0 + owner // 27, order[a[owner] + within]


def _unit(v):
    return v / v.norm(dim=1, keepdim=True).clamp_min(1e-30)


def proximity_stencils(mesh: Mesh, radius: float, stiff: float = 500.0, device=None) -> Mesh:
    """Fills mesh.ef / ee / vf (+ totals) from the geometry and returns the mesh."""
    import torch
    assert mesh.ne > 0 and mesh.nf > 0, "need edges/faces (with_topology=True)"
    dev = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    P = torch.as_tensor(np.ascontiguousarray(mesh.positions[:, :3]), device=dev, dtype=torch.float32)
    E = torch.as_tensor(np.ascontiguousarray(mesh.edges[:, :2]).astype(np.int64), device=dev)
    F = torch.as_tensor(np.ascontiguousarray(mesh.faces[:, :3]).astype(np.int64), device=dev)
    ea, eb = P[E[:, 0]], P[E[:, 1]]
    fa, fb, fc = P[F[:, 0]], P[F[:, 1]], P[F[:, 2]]
    longest = float((eb - ea).norm(dim=1).max())
    cell = float(radius) + longest
    centroid = (fa + fb + fc) / 3
    mid = (ea + eb) / 2

    # ---- VF: point - triangle
    vf = []
    for vi, fi in _candidates(P, centroid, cell):
        keep = (F[fi] != vi[:, None]).all(1)
        vi, fi = vi[keep], fi[keep]
        p, a, b, c = P[vi], fa[fi], fb[fi], fc[fi]
        e0, e1, d = a - c, b - c, p - c                                   # q = c + b0 e0 + b1 e1
        g00, g01, g11 = (e0 * e0).sum(1), (e0 * e1).sum(1), (e1 * e1).sum(1)
        r0, r1 = (d * e0).sum(1), (d * e1).sum(1)
        det = (g00 * g11 - g01 * g01).clamp_min(1e-30)
        b0, b1 = (r0 * g11 - r1 * g01) / det, (r1 * g00 - r0 * g01) / det
        q = c + b0[:, None] * e0 + b1[:, None] * e1
        dist = (p - q).norm(dim=1)
        inside = (b0 > 0) & (b1 > 0) & (b0 + b1 < 1)
        ok = inside & (dist < radius) & (dist > 1e-7)
        vf.append((vi[ok], fi[ok], b0[ok], b1[ok], _unit((p - q)[ok])))
    # ---- EE: segment - segment
    ee = []
    for e0i, e1i in _candidates(mid, mid, cell):
        keep = e0i < e1i
        e0i, e1i = e0i[keep], e1i[keep]
        keep = (E[e0i][:, :, None] != E[e1i][:, None, :]).all(2).all(1)
        e0i, e1i = e0i[keep], e1i[keep]
        p1, q1, p2, q2 = ea[e0i], eb[e0i], ea[e1i], eb[e1i]
        d1, d2, r = q1 - p1, q2 - p2, p1 - p2
        a, e, f = (d1 * d1).sum(1), (d2 * d2).sum(1), (d2 * r).sum(1)
        b, c = (d1 * d2).sum(1), (d1 * r).sum(1)
        den = (a * e - b * b)
        s = torch.where(den > 1e-20, (b * f - c * e) / den.clamp_min(1e-30), torch.zeros_like(den))
        t = (b * s + f) / e.clamp_min(1e-30)
        c0, c1 = p1 + s[:, None] * d1, p2 + t[:, None] * d2
        dist = (c0 - c1).norm(dim=1)
        ok = (den > 1e-20) & (s > 0) & (s < 1) & (t > 0) & (t < 1) & (dist < radius) & (dist > 1e-7)
        # weights: edge0[0] gets bary0 -> c0 = bary0 p1 + (1 - bary0) q1, i.e. bary0 = 1 - s
        ee.append((e0i[ok], e1i[ok], (1 - s)[ok], (1 - t)[ok], _unit((c0 - c1)[ok])))
    # ---- EF: segment pierces triangle
    ef = []
    for ei, fi in _candidates(mid, centroid, cell):
        keep = (E[ei][:, :, None] != F[fi][:, None, :]).all(2).all(1)
        ei, fi = ei[keep], fi[keep]
        p, q, a, b, c = ea[ei], eb[ei], fa[fi], fb[fi], fc[fi]
        n = torch.linalg.cross(b - a, c - a)
        dp, dq = ((p - a) * n).sum(1), ((q - a) * n).sum(1)
        cross = (dp * dq < 0)
        t = dp / (dp - dq).where(cross, torch.ones_like(dp))
        x = p + t[:, None] * (q - p)
        e0, e1, d = a - c, b - c, x - c
        g00, g01, g11 = (e0 * e0).sum(1), (e0 * e1).sum(1), (e1 * e1).sum(1)
        r0, r1 = (d * e0).sum(1), (d * e1).sum(1)
        det = (g00 * g11 - g01 * g01).clamp_min(1e-30)
        b0, b1 = (r0 * g11 - r1 * g01) / det, (r1 * g00 - r0 * g01) / det
        ok = cross & (b0 > 0) & (b1 > 0) & (b0 + b1 < 1)
        nn = _unit(n) * torch.sign(dp)[:, None]
        ef.append((ei[ok], fi[ok], (1 - t)[ok], b0[ok], b1[ok], nn[ok]))

    def cat(parts, k):
        if not parts:
            return [torch.zeros(0, device=dev)] * k
        return [torch.cat([p[j] for p in parts]) for j in range(k)]

    def ordered(cols):
        if cols[0].numel() == 0:
            return [c.cpu().numpy() for c in cols]
        key = cols[0].to(torch.int64) * (int(cols[1].max()) + 1) + cols[1].to(torch.int64)
        o = torch.argsort(key, stable=True)
        return [c[o].cpu().numpy() for c in cols]

    efc, eec, vfc = ordered(cat(ef, 6)), ordered(cat(ee, 5)), ordered(cat(vf, 5))
    n_ef, n_ee, n_vf = len(efc[0]), len(eec[0]), len(vfc[0])
    total = n_ef + n_ee + n_vf
    A, Bq, C = np.zeros(total, EF_DTYPE), np.zeros(total, EE_DTYPE), np.zeros(total, VF_DTYPE)

    def normal4(nrm):
        out = np.zeros((nrm.shape[0], 4), np.float32)
        out[:, :3] = nrm
        return out
    if n_ef:
        s = slice(0, n_ef)
        A["eId"][s], A["fId"][s], A["stiff"][s] = efc[0], efc[1], stiff
        A["bary"][s] = np.stack([efc[2], efc[3], efc[4]], 1)
        A["normal"][s] = normal4(efc[5])
    if n_ee:
        s = slice(n_ef, n_ef + n_ee)
        Bq["eId0"][s], Bq["eId1"][s], Bq["stiff"][s] = eec[0], eec[1], stiff
        Bq["bary"][s] = np.stack([eec[2], eec[3]], 1)
        Bq["normal"][s] = normal4(eec[4])
    if n_vf:
        s = slice(n_ef + n_ee, total)
        C["vId"][s], C["fId"][s], C["stiff"][s] = vfc[0], vfc[1], stiff
        C["bary"][s] = np.stack([vfc[2], vfc[3]], 1)
        C["pad"][s] = (vfc[2] + vfc[3]).astype(np.float32)
        C["normal"][s] = normal4(vfc[4])
    mesh.ef, mesh.ee, mesh.vf = A, Bq, C
    mesh.ef_total, mesh.ee_total, mesh.vf_total = n_ef, n_ee, n_vf
    mesh.name += f"+prox(ef{n_ef},ee{n_ee},vf{n_vf})"
    return mesh
