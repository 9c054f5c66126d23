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


def project_views(world_points: torch.Tensor, extrinsics: torch.Tensor, intrinsic: torch.Tensor,
                  colors: torch.Tensor, H: int, W: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """All V frames of a camera trajectory in one launch sequence (render_trajectory, infer.py:398-444,
    calls render_with_project once per frame).  world_points / colors: [N, 3] shared by all views or
    [V, N, 3] per view; extrinsics [V, 4, 4].  Returns device tensors (images uint8 [V, H, W, 3], masks
    uint8 [V, H, W]), bit-identical to V calls of project_points."""
    _lib.require_device()
    if not world_points.is_cuda:
        raise RuntimeError("more4d_b200.render: world_points must be a CUDA tensor (no CPU fallback)")
    dev = world_points.device
    pts = world_points.to(torch.float32).contiguous()
    col = colors.to(device=dev, dtype=torch.float32).contiguous()
    V = extrinsics.shape[0]
    N = pts.shape[-2]
    for name, t in (("world_points", pts), ("colors", col)):
        if t.shape[-1] != 3 or t.shape[-2] != N or t.dim() not in (2, 3) or (t.dim() == 3 and t.shape[0] != V):
            raise ValueError(f"more4d_b200.render: {name} must be [N, 3] or [V, N, 3]")
    w2c = torch.linalg.inv(extrinsics.detach().to("cpu", torch.float32)).contiguous()
    K = intrinsic.detach().to("cpu", torch.float32).contiguous()
    lib = _lib.lib()
    ws_bytes = lib.m4d_project_views_workspace(N, V, H, W)
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    image = torch.empty(V, H, W, 3, device=dev, dtype=torch.uint8)
    mask = torch.empty(V, H, W, device=dev, dtype=torch.uint8)
    from . import ops
    rc = lib.m4d_project_views(pts.data_ptr(), 3 * N if pts.dim() == 3 else 0, col.data_ptr(),
                               3 * N if col.dim() == 3 else 0, w2c.data_ptr(), K.data_ptr(), N, V, H, W,
                               image.data_ptr(), mask.data_ptr(), ws.data_ptr(), ws_bytes, ops._stream())
    _lib.check(rc, "m4d_project_views")
    ops._Stats.launches += 2
    return image, mask


def render_with_project(world_points: torch.Tensor, extrinsic: torch.Tensor, intrinsic: torch.Tensor,
                        colors: torch.Tensor, H: int, W: int, device=None) -> Tuple[np.ndarray, np.ndarray]:
    """infer.py:222-258: returns (image_proj uint8 [H, W, 3], mask bool [H, W]) as numpy arrays."""
    if device is not None:
        world_points = world_points.to(device)
    image, mask = project_points(world_points, extrinsic, intrinsic, colors, H, W)
    return image.cpu().numpy(), mask.cpu().numpy().astype(bool)


# --------------------------------------------------------------------------------------
# 3D Gaussian splatting (render_with_gs, infer.py:260-273; gs_render, gaussian_splatting.py:13-43)
# --------------------------------------------------------------------------------------
def _quaternion_to_matrix(q: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """MoRe4D/utils/gaussian_splatting.py:115-137 (xyzw)."""
    i, j, k, r = q.unbind(-1)
    two_s = 2 / ((q * q).sum(-1) + eps)
    return torch.stack([1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)], -1).reshape(3, 3)


def _camera_rows(extrinsics: torch.Tensor, intrinsic: torch.Tensor, H: int, W: int, near=0.5, far=1000.0) -> torch.Tensor:
    """Per-view camera records of csrc/gsplat.cu (GsCam) from cam->world extrinsics [V, 4, 4] and the
    normalised intrinsic [3, 3]: host-side scalar work, like get_fov / get_projection_matrix /
    extrinsics.inverse() in render_cuda (gaussian_splatting.py:179-243)."""
    ext = extrinsics.detach().to("cpu", torch.float32)
    K = intrinsic.detach().to("cpu", torch.float32)
    Kinv = torch.linalg.inv(K)

    def unit(v):
        v = Kinv @ torch.tensor(v, dtype=torch.float32)
        return v / v.norm()
    fov_x = torch.acos((unit([0, 0.5, 1]) * unit([1, 0.5, 1])).sum())
    fov_y = torch.acos((unit([0.5, 0, 1]) * unit([0.5, 1, 1])).sum())
    tx, ty = torch.tan(0.5 * fov_x), torch.tan(0.5 * fov_y)
    n, f = torch.tensor(near), torch.tensor(far)
    view = torch.linalg.inv(ext)[:, :3, :].reshape(-1, 12)
    rest = torch.stack([W / (2 * tx), H / (2 * ty), tx, ty, 2 * n / (2 * tx * n), 2 * n / (2 * ty * n),
                        f / (f - n), -(f * n) / (f - n)]).to(torch.float32)
    return torch.cat([view, rest.expand(view.shape[0], 8)], dim=1).contiguous()


def gs_render_views(means: torch.Tensor, colors: torch.Tensor, opacities: torch.Tensor, scale: torch.Tensor,
                    rotation: torch.Tensor, extrinsics: torch.Tensor, intrinsic: torch.Tensor, H: int, W: int,
                    want_uint8: bool = False, dup_per_gaussian: float = 6.0):
    """V views in ONE launch sequence.  means [V, N, 3] (or [N, 3], shared by all views); colors [N, 3]
    in 0..1; opacities [N]; scale [3], rotation xyzw [4] shared by all gaussians; extrinsics
    [V, 4, 4] cam->world.  Returns the float32 image [V, 3, H, W] (and uint8 [V, H, W, 3])."""
    _lib.require_device()
    from . import ops
    if not means.is_cuda:
        raise RuntimeError("more4d_b200.render: means must be a CUDA tensor (no CPU fallback)")
    dev = means.device
    V = extrinsics.shape[0]
    m = means.to(torch.float32).contiguous()
    shared = m.dim() == 2
    N = m.shape[-2]
    if (not shared and m.shape[0] != V) or m.shape[-1] != 3:
        raise ValueError("more4d_b200.render: means must be [V, N, 3] or [N, 3]")
    col = colors.to(device=dev, dtype=torch.float32).contiguous()
    op = opacities.to(device=dev, dtype=torch.float32).contiguous()
    if col.shape != (N, 3) or op.shape != (N,):
        raise ValueError("more4d_b200.render: colors must be [N, 3] and opacities [N]")
    S = torch.diag(scale.detach().to("cpu", torch.float32))
    R = _quaternion_to_matrix(rotation.detach().to("cpu", torch.float32))
    cov = R @ S @ S.T @ R.T                                               # build_covariance, :140-151
    cov6 = torch.stack([cov[0, 0], cov[0, 1], cov[0, 2], cov[1, 1], cov[1, 2], cov[2, 2]]).to(dev)
    cams = _camera_rows(extrinsics, intrinsic, H, W).to(dev)
    image = torch.empty(V, 3, H, W, device=dev, dtype=torch.float32)
    image_u8 = torch.empty(V, H, W, 3, device=dev, dtype=torch.uint8) if want_uint8 else None
    lib = _lib.lib()
    import ctypes
    cap = int(dup_per_gaussian * N * V) + 1024
    for _ in range(2):
        ws_bytes = lib.m4d_gs_render_workspace(N, V, H, W, cap)
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        needed = ctypes.c_longlong(0)
        rc = lib.m4d_gs_render(m.data_ptr(), 0 if shared else N * 3, col.data_ptr(), 0, op.data_ptr(), cov6.data_ptr(),
                               cams.data_ptr(), N, V, H, W, 0.0, 0.0, 0.0, image.data_ptr(),
                               None if image_u8 is None else image_u8.data_ptr(), ws.data_ptr(), ws_bytes, cap,
                               ctypes.byref(needed), ops._stream())
        if rc == -4 and needed.value > cap:          # M4D_ERR_WORKSPACE: more (tile, gaussian) pairs than planned for
            cap = needed.value + 1024
            continue
        break
    _lib.check(rc, "m4d_gs_render")
    ops._Stats.launches += 5                         # preprocess, scan, scatter, sort, render (+ uint8) behind one call
    return (image, image_u8) if want_uint8 else image


def gs_render(intrinsic, extrinsic, image_shape, means, scale, rotation, color, opacities):
    """Signature of the reference's gs_render (gaussian_splatting.py:13-43): one view, returns the
    float image [1, 3, H, W] on the device."""
    H, W = image_shape
    return gs_render_views(means, color, opacities, scale, rotation, extrinsic.unsqueeze(0), intrinsic, H, W)


def render_with_gs(world_points: torch.Tensor, extrinsic: torch.Tensor, intrinsic: torch.Tensor,
                   colors: torch.Tensor, H: int, W: int, device=None, scale: float = 0.0001) -> np.ndarray:
    """infer.py:260-273: uint8 [H, W, 3] numpy image of one view."""
    if device is not None:
        world_points = world_points.to(device)
    c = colors.float() / 255.0 if colors.max() > 1.0 else colors.float()
    N = world_points.shape[0]
    _, u8 = gs_render_views(world_points, c, torch.ones(N), torch.tensor([scale] * 3), torch.tensor([0.0, 0.0, 0.0, 1.0]),
                            extrinsic.unsqueeze(0), intrinsic, H, W, want_uint8=True)
    return u8[0].cpu().numpy()
