"""Oracle of the hand-off renderer (oracle/project_oracle.py) against outputs of the REAL
`render_with_project` (scripts/inference/infer.py:222-258; tests/golden/make_golden.py execs its
source).  Integer outputs: compared exactly."""
import numpy as np
import torch

from more4d_b200 import synth
from oracle import project_oracle as P
from tests.helpers import checksum

CASES = {"a": (48, 64, 0, 0.0), "b": (37, 53, 1, 0.15)}


def test_oracle_matches_reference_render_with_project(golden):
    g = golden("project")
    for name, (H, W, seed, tilt) in CASES.items():
        pts, col, ext, K = synth.point_cloud(H, W, seed, tilt)
        assert torch.allclose(checksum(pts), g[f"{name}.pts_sum"], rtol=1e-6), "RNG drift: regenerate goldens"
        img, mask = P.render_with_project(pts.numpy(), torch.linalg.inv(ext).numpy(), K.numpy(), col.numpy(), H, W)
        assert img.dtype == np.uint8 and img.shape == (H, W, 3) and mask.shape == (H, W)
        assert np.array_equal(img, g[f"{name}.image"].numpy())
        assert np.array_equal(mask, g[f"{name}.mask"].numpy().astype(bool))
        assert 0.0 < mask.mean() < 1.0                      # the case has both hits and holes


def test_oracle_edge_cases():
    K = np.array([[1, 0, 0.5], [0, 1, 0.5], [0, 0, 1]], np.float32)
    E = np.eye(4, dtype=np.float32)
    # nothing in the frustum: all-zero image, all-hole mask (infer.py:253-254)
    pts = np.array([[0, 0, -1], [5, 5, 1]], np.float32)
    img, mask = P.render_with_project(pts, E, K, np.full((2, 3), 200, np.float32), 4, 6)
    assert img.sum() == 0 and mask.all()
    # z-buffer: nearer point wins; exact ties are averaged; a black point still counts as a hole
    pts = np.array([[0, 0, 2], [0, 0, 1], [0, 0, 1], [0.2, 0.2, 1]], np.float32)
    col = np.array([[9, 9, 9], [100, 0, 50], [200, 0, 51], [0, 0, 0]], np.float32)
    img, mask = P.render_with_project(pts, E, K, col, 4, 4)
    assert tuple(img[1, 1]) == (150, 0, 50)              # u = 0.5 / (1 + eps) -> pixel 1, not 2
    assert not mask[1, 1] and mask[2, 2] and mask.sum() == 15
