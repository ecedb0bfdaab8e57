"""The torch (device) cloth generator used by the sharded bench produces the same arrays as the numpy recipe."""
import numpy as np
import pytest


@pytest.mark.parametrize("nx,ny", [(8, 8), (40, 24), (64, 32), (1, 5), (2, 2)])
def test_device_cloth_generator_matches_numpy(nx, ny, synth):
    import torch
    a = synth.cloth_rect(nx, ny)
    b = synth.cloth_rect_device(nx, ny, torch.device("cpu"))
    assert np.array_equal(a.positions, b.positions.numpy())
    assert np.array_equal(a.nbr_starts, b.nbr_starts.numpy())
    assert np.array_equal(a.nbr_idx, b.nbr_idx.numpy())
    assert np.array_equal(a.offdiag, b.offdiag.numpy())
    assert np.array_equal(a.diag, b.diag.numpy())


def test_weak_scaling_mesh_shapes(synth, monkeypatch):
    seen = []
    monkeypatch.setattr(synth, "cloth_rect", lambda nx, ny, *a, **k: seen.append((nx, ny)))
    for n in (1, 2, 4, 8):
        synth.weak_scaling_cloth(n)
    assert seen == [(1024, 1024), (2048, 1024), (2048, 2048), (4096, 2048)]


@pytest.mark.gpu
def test_device_cloth_generator_on_gpu_matches_numpy(synth):
    import torch
    a = synth.cloth_rect(96, 40)
    b = synth.cloth_rect_device(96, 40, torch.device("cuda:0"))
    for x, y in ((a.positions, b.positions), (a.nbr_starts, b.nbr_starts), (a.nbr_idx, b.nbr_idx), (a.offdiag, b.offdiag),
                 (a.diag, b.diag)):
        assert np.array_equal(x, y.cpu().numpy())


def test_oracle_known_answer_on_edge_free_particles(oracle_lib, synth):
    """No edges: nothing aggregates, every level keeps nv one-vertex clusters (far beyond the reference's fixed 1.5x
    allocation, so the compiled reference cannot run this) and every domain matrix is the mass block: z = numLevel * r."""
    from helpers import make_oracle
    mesh = synth.dust(6000)
    o = make_oracle(oracle_lib, mesh)
    assert o.num_level == 3
    assert o.level_size().tolist() == [[0, 0], [6000, 6016], [6000, 12032], [6000, 18048]]
    r = synth.residual(mesh.nv)
    z = o.apply(r)
    assert np.array_equal(z[:, :3], np.float32(3) * r[:, :3]) and np.all(z[:, 3] == 0)
