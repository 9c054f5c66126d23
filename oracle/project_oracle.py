"""TEST INFRASTRUCTURE — CPU restatement (numpy, float32) of the reference's point-cloud
projection with a z-buffer: `render_with_project`, scripts/inference/infer.py:222-258, with
`project` / `project_camera_space` from MoRe4D/utils/project_utils.py:47-71.

Only tests/, __graft_entry__.smoke() and bench tools may import this module; the product path is
more4d_b200/render.py -> m4d_project_points (CUDA).

Pinning: tests/golden/project.safetensors holds outputs of the REAL reference function (its source
is exec'd from /root/reference by tests/golden/make_golden.py with a torch_scatter `scatter(mean)`
stub — the reference's own dependency is absent here).  Products and sums are taken in the order
the reference writes them, one float32 rounding per operation; torch's einsum may contract in a
different order / with FMA, which can move a point across a pixel border: the golden comparison
allows a stated mismatch budget, the CUDA kernel is bit-exact against THIS restatement.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
EPS = F32(np.finfo(np.float32).eps)


def project(points: np.ndarray, world2cam: np.ndarray, intrinsic: np.ndarray):
    """project_utils.py:59-71 with extrinsics.inverse() already applied: returns (uv [N,2], depth [N])."""
    p = points.astype(F32)
    e = world2cam.astype(F32)
    k = intrinsic.astype(F32)
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    cam = [((e[r, 0] * x + e[r, 1] * y) + e[r, 2] * z) + e[r, 3] for r in range(3)]
    depth = cam[2]
    den = depth + EPS
    with np.errstate(divide="ignore", invalid="ignore"):
        q = [np.nan_to_num(c / den, nan=0.0, posinf=1e8, neginf=-1e8).astype(F32) for c in cam]
    u = (k[0, 0] * q[0] + k[0, 1] * q[1]) + k[0, 2] * q[2]
    v = (k[1, 0] * q[0] + k[1, 1] * q[1]) + k[1, 2] * q[2]
    return np.stack([u, v], -1).astype(F32), depth.astype(F32)


def render_with_project(points, world2cam, intrinsic, colors, H: int, W: int):
    """infer.py:222-258.  points fp32 [N,3], colors fp32 [N,3] in 0..255.  Returns
    (image uint8 [H,W,3], mask bool [H,W])."""
    uv, depth = project(points, world2cam, intrinsic)
    u, v = uv[:, 0], uv[:, 1]
    ok = (u >= 0) & (u <= 1) & (v >= 0) & (v <= 1) & (depth >= 0)
    image = np.zeros((H, W, 3), np.uint8)
    if ok.any():
        fx = np.clip(np.floor(u[ok] * F32(W)), 0, W - 1).astype(F32)
        fy = np.clip(np.floor(v[ok] * F32(H)), 0, H - 1).astype(F32)
        idx = (fx * F32(H) + fy).astype(np.int64)                     # COLUMN-major pixel index
        d = depth[ok] + F32(0.0)
        zmin = np.full(H * W, np.inf, F32)
        np.minimum.at(zmin, idx, d)                                   # z-buffer (unique + index_reduce_ amin)
        front = d == zmin[idx]                                        # ties kept (infer.py:241)
        acc = np.zeros((H * W, 3), np.float64)
        cnt = np.zeros(H * W, np.float64)
        np.add.at(acc, idx[front], colors[ok][front].astype(np.float64))
        np.add.at(cnt, idx[front], 1.0)
        mean = np.zeros((H * W, 3), F32)
        nz = cnt > 0
        mean[nz] = (acc[nz].astype(F32) / cnt[nz, None].astype(F32)).astype(F32)   # scatter(mean)
        image = mean.reshape(W, H, 3).transpose(1, 0, 2).astype(np.uint8)
    mask = image.astype(np.uint64).sum(-1) == 0
    return image, mask
