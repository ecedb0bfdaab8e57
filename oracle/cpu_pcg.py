"""TEST INFRASTRUCTURE ONLY - CPU-side PCG loop used as the checker for mas_pcg_solve and as bench.py's CPU baseline leg.

The reference ships no solver (its caller owns the loop, SeSchwarzPreconditioner.h:55-63); this is the plain textbook
loop with the conventions SURVEY 8d fixes: x0 = 0, stop at ||r||/||b|| < tol, FP64 dot products, FP32 vectors."""
import numpy as np


def bsr_matrix(mesh, dtype=np.float32, stencils=None):
    """The synthetic Hessian as scipy CSR (diag + off-diagonal 3x3 blocks; SeMatrix3f is column-major).  `stencils`: the
    80-byte Stencil records of a preconditioner built on this mesh (oracle / reference / GPU `stencils()[0]`): their collision
    Hessians stiff (w (x) w) (x) (d d^T) (cpp:1201-1227) are added, as the caller's system matrix contains them."""
    import scipy.sparse as sp
    nv = mesh.nv
    off = np.asarray(mesh.offdiag, dtype).reshape(-1, 3, 3).transpose(0, 2, 1)
    dia = np.asarray(mesh.diag, dtype).reshape(-1, 3, 3).transpose(0, 2, 1)
    A = sp.bsr_matrix((off, mesh.nbr_idx, mesh.nbr_starts), shape=(3 * nv, 3 * nv))
    D = sp.bsr_matrix((dia, np.arange(nv, dtype=np.int32), np.arange(nv + 1, dtype=np.int32)), shape=(3 * nv, 3 * nv))
    M = (A + D).tocsr()
    if stencils is not None and len(stencils):
        raw = np.frombuffer(np.ascontiguousarray(stencils).tobytes(), np.uint8).reshape(-1, 80)
        n = raw[:, 0:4].copy().view(np.int32).ravel()
        idx = raw[:, 8:28].copy().view(np.int32).reshape(-1, 5)
        w = raw[:, 28:48].copy().view(np.float32).reshape(-1, 5).astype(np.float64)
        stiff = raw[:, 48:52].copy().view(np.float32).ravel().astype(np.float64)
        d = raw[:, 64:76].copy().view(np.float32).reshape(-1, 3).astype(np.float64)
        valid = np.arange(5)[None, :] < n[:, None]
        w = np.where(valid, w, 0.0)
        idx = np.where(valid, idx, 0)
        # J_s = [w_0 d^T ... w_4 d^T] (1 x 3nv row), H_s = stiff J_s^T J_s
        rows = np.repeat(np.arange(len(n)), 15)
        cols = (3 * idx[:, :, None] + np.arange(3)[None, None, :]).reshape(-1)
        vals = (w[:, :, None] * d[:, None, :]).reshape(-1)
        J = sp.csr_matrix((vals, (rows, cols)), shape=(len(n), 3 * nv))
        M = (M + (J.T @ sp.diags(stiff) @ J).astype(dtype)).tocsr()
    return M


def cpu_pcg(A, b, apply_precond, rel_tol=1e-5, max_iter=5000):
    """Reference PCG loop for the iteration-count parity tests: FP32 vectors, FP64 dot products, x0 = 0, stop when
    ||r||/||b|| < rel_tol.  apply_precond(r[nv,4] f32) -> z[nv,4] f32, or None for plain CG."""
    nv = b.shape[0]
    dot = lambda u, v: float(np.dot(u.astype(np.float64), v.astype(np.float64)))
    def M(r):
        if apply_precond is None:
            return r.copy()
        r4 = np.zeros((nv, 4), np.float32)
        r4[:, :3] = r.reshape(nv, 3)
        return np.ascontiguousarray(apply_precond(r4)[:, :3]).reshape(-1).astype(np.float32)
    x = np.zeros(3 * nv, np.float32)
    r = np.ascontiguousarray(b[:, :3], np.float32).reshape(-1).copy()
    rr0 = dot(r, r)
    z = M(r)
    p = z.copy()
    rz = dot(r, z)
    it = 0
    while it < max_iter:
        Ap = (A @ p).astype(np.float32)
        alpha = np.float32(rz / dot(p, Ap))
        x += alpha * p
        r -= alpha * Ap
        it += 1
        if dot(r, r) < rel_tol * rel_tol * rr0:
            break
        z = M(r)
        rz_new = dot(r, z)
        p = z + np.float32(rz_new / rz) * p
        rz = rz_new
    return x.reshape(nv, 3), it
