"""GPU parity of the z-buffer projection (m4d_project_points through the C ABI) — integer
outputs, compared bit for bit with the CPU oracle; the oracle itself is pinned to the real
reference function in tests/test_project_cpu.py."""
import numpy as np
import pytest
import torch

from more4d_b200 import synth
from oracle import project_oracle as P

pytestmark = pytest.mark.gpu


def _run(pts, col, ext, K, H, W):
    from more4d_b200 import render
    return render.render_with_project(pts.cuda(), ext, K, col.cuda(), H, W)


@pytest.mark.parametrize("H,W,seed,tilt", [(48, 64, 0, 0.0), (37, 53, 1, 0.15), (90, 160, 2, -0.3),
                                           (64, 64, 3, 1.2), (17, 300, 4, 0.05)])
def test_projection_matches_oracle_exactly(H, W, seed, tilt):
    pts, col, ext, K = synth.point_cloud(H, W, seed, tilt)
    img, mask = _run(pts, col, ext, K, H, W)
    ri, rm = P.render_with_project(pts.numpy(), torch.linalg.inv(ext).numpy(), K.numpy(), col.numpy(), H, W)
    assert img.dtype == np.uint8 and mask.dtype == bool
    assert np.array_equal(img, ri) and np.array_equal(mask, rm)


def test_projection_golden_and_edges(golden):
    g = golden("project")
    pts, col, ext, K = synth.point_cloud(48, 64, 0, 0.0)
    img, mask = _run(pts, col, ext, K, 48, 64)
    assert np.array_equal(img, g["a.image"].numpy()) and np.array_equal(mask, g["a.mask"].numpy().astype(bool))
    # nothing visible -> zeros + all holes; ragged N (not a multiple of the block size)
    E, Kn = torch.eye(4), torch.tensor([[1, 0, 0.5], [0, 1, 0.5], [0, 0, 1.0]])
    p = torch.tensor([[0, 0, -1.0], [5, 5, 1.0], [0, 0, float("nan")]])
    img, mask = _run(p, torch.full((3, 3), 200.0), E, Kn, 5, 7)
    assert img.sum() == 0 and mask.all()
    p = torch.tensor([[0, 0, 2.0], [0, 0, 1.0], [0, 0, 1.0], [0.2, 0.2, 1.0]])
    c = torch.tensor([[9, 9, 9.0], [100, 0, 50], [200, 0, 51], [0, 0, 0]])
    img, mask = _run(p, c, E, Kn, 4, 4)
    assert tuple(img[1, 1]) == (150, 0, 50) and mask.sum() == 15
    with pytest.raises(RuntimeError):
        from more4d_b200 import render
        render.render_with_project(p, E, Kn, c, 4, 4)            # CPU tensor: no fallback


def test_projection_fullsize_properties():
    """720 x 1280 (one point per pixel, BASELINE frame size): size-independent properties."""
    from more4d_b200 import render
    H, W = 720, 1280
    K = render.get_intrinsic_matrix(H, W)
    g = torch.Generator().manual_seed(7)
    v, u = torch.meshgrid((torch.arange(H) + 0.5) / H, (torch.arange(W) + 0.5) / W, indexing="ij")
    z = 1.0 + 4.0 * torch.rand(H, W, generator=g)
    pts = torch.stack([(u - 0.5) / K[0, 0] * z, (v - 0.5) / K[1, 1] * z, z], -1).reshape(-1, 3)
    col = torch.randint(1, 256, (H * W, 3), generator=g).float()
    # (1) identity camera: every point lands on its own pixel -> the colour image comes back
    img, mask = render.render_with_project(pts.cuda(), torch.eye(4), K, col.cuda(), H, W)
    assert np.array_equal(img, col.reshape(H, W, 3).numpy().astype(np.uint8)) and not mask.any()
    # (2) order independence: a permutation of the points renders the same image (moved camera)
    ext = torch.eye(4)
    ext[0, 3], ext[2, 3] = 0.3, -0.5
    a_img, a_mask = render.render_with_project(pts.cuda(), ext, K, col.cuda(), H, W)
    perm = torch.randperm(H * W, generator=g)
    b_img, b_mask = render.render_with_project(pts[perm].cuda(), ext, K, col[perm].cuda(), H, W)
    assert np.array_equal(a_img, b_img) and np.array_equal(a_mask, b_mask)
    assert np.array_equal(a_mask, a_img.astype(np.int64).sum(-1) == 0) and 0 < a_mask.mean() < 1
    # (3) against the oracle at full size
    ri, rm = P.render_with_project(pts.numpy(), torch.linalg.inv(ext).numpy(), K.numpy(), col.numpy(), H, W)
    assert np.array_equal(a_img, ri) and np.array_equal(a_mask, rm)


def test_project_views_batched_equals_per_frame_calls():
    """m4d_project_views: the V frames of a trajectory (moving points, moving camera) in one launch
    sequence are bit-identical to V calls of render_with_project's kernel path."""
    from more4d_b200 import render
    H, W, V = 37, 53, 4
    pts, col, ext, K = synth.point_cloud(H, W, 5, 0.1)
    exts = []
    for k in range(V):
        e = ext.clone()
        e[0, 3] += 0.04 * k
        e[2, 3] -= 0.02 * k
        exts.append(e)
    exts = torch.stack(exts)
    moving = torch.stack([pts + 0.01 * k for k in range(V)]).cuda()
    img, mask = render.project_views(moving, exts, K, col.cuda(), H, W)
    for k in range(V):
        i1, m1 = render.project_points(moving[k], exts[k], K, col.cuda(), H, W)
        assert torch.equal(img[k], i1) and torch.equal(mask[k], m1)
    # shared points, several cameras
    img2, _ = render.project_views(pts.cuda(), exts, K, col.cuda(), H, W)
    i0, _ = render.project_points(pts.cuda(), exts[2], K, col.cuda(), H, W)
    assert torch.equal(img2[2], i0)
