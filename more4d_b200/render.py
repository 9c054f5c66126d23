"""Host mirror of the stage-1 -> stage-2 hand-off renderer: `render_with_project` and
`get_intrinsic_matrix` of scripts/inference/infer.py:161-176,222-258.
Same argument order and return types as the reference (numpy uint8 image, bool hole mask); the
arithmetic is the CUDA z-buffer in csrc/project.cu."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from . import _lib

DEFAULT_H_ORI, DEFAULT_W_ORI = 540, 960          # scripts/inference/infer.py:53


def get_intrinsic_matrix(H: int, W: int, device="cpu") -> torch.Tensor:
    """infer.py:161-176."""
    H_ori, W_ori = DEFAULT_H_ORI, DEFAULT_W_ORI
    if W_ori / W > H_ori / H:
        fx, fy = 1.0, W_ori / H_ori / (W / H)
    else:
        fy, fx = 1.0, H_ori / W_ori / (H / W)
    return torch.tensor([[fx, 0, 0.5], [0, fy, 0.5], [0, 0, 1]], dtype=torch.float32, device=device)


def project_points(world_points: torch.Tensor, extrinsic: torch.Tensor, intrinsic: torch.Tensor,
                   colors: torch.Tensor, H: int, W: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Device-side result: (image uint8 [H, W, 3], mask uint8 [H, W]) on world_points' device."""
    _lib.require_device()
    if not world_points.is_cuda:
        raise RuntimeError("more4d_b200.render: world_points must be a CUDA tensor (no CPU fallback)")
    dev = world_points.device
    pts = world_points.to(torch.float32).contiguous()
    col = colors.to(device=dev, dtype=torch.float32).contiguous()
    if pts.dim() != 2 or pts.shape[1] != 3 or col.shape != pts.shape:
        raise ValueError("more4d_b200.render: world_points / colors must be [N, 3]")
    # 16 + 9 floats of camera parameters: inverted on the host like the reference's
    # extrinsics.inverse() (project_utils.py:44), passed as kernel arguments
    w2c = torch.linalg.inv(extrinsic.detach().to("cpu", torch.float32)).contiguous()
    K = intrinsic.detach().to("cpu", torch.float32).contiguous()
    N = pts.shape[0]
    lib = _lib.lib()
    ws_bytes = lib.m4d_project_points_workspace(N, H, W)
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    image = torch.empty(H, W, 3, device=dev, dtype=torch.uint8)
    mask = torch.empty(H, W, device=dev, dtype=torch.uint8)
    from . import ops
    rc = lib.m4d_project_points(pts.data_ptr(), col.data_ptr(), w2c.data_ptr(), K.data_ptr(), N, H, W,
                                image.data_ptr(), mask.data_ptr(), ws.data_ptr(), ws_bytes, ops._stream())
    _lib.check(rc, "m4d_project_points")
    ops._Stats.launches += 2                      # three kernels behind one C-ABI call
    return image, mask


def render_with_project(world_points: torch.Tensor, extrinsic: torch.Tensor, intrinsic: torch.Tensor,
                        colors: torch.Tensor, H: int, W: int, device=None) -> Tuple[np.ndarray, np.ndarray]:
    """infer.py:222-258: returns (image_proj uint8 [H, W, 3], mask bool [H, W]) as numpy arrays."""
    if device is not None:
        world_points = world_points.to(device)
    image, mask = project_points(world_points, extrinsic, intrinsic, colors, H, W)
    return image.cpu().numpy(), mask.cpu().numpy().astype(bool)
