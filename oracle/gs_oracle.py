"""CPU oracle of the 3D-Gaussian-splatting forward render used by the stage-1 -> stage-2 hand-off
(`render_with_gs`, scripts/inference/infer.py:260-273 -> `gs_render`, MoRe4D/utils/
gaussian_splatting.py:13-43 -> `render_cuda` :201-281 -> third-party `GaussianRasterizer`).

TEST INFRASTRUCTURE — see oracle/__init__.py for who may import this.

PARITY UNPINNED.  The rasteriser itself is the third-party CUDA extension
`diff_gaussian_rasterization` (graphdeco-inria, commit 8064f52ca233942bdec2d1a1451c026deedd320b per
README.md:60); it is neither in /root/reference nor installed here, and the reference has no test or
golden vector for this path.  This file restates (a) the reference's own wrapper arithmetic — which IS
in the tree: build_covariance / quaternion_to_matrix (:115-176), get_fov (:179-195),
get_projection_matrix (:198-226), the settings passed at :253-266 (bg 0, near 0.5, far 1000,
colors_precomp, cov3D_precomp, scale_modifier 1, prefiltered False) — and (b) the PUBLISHED forward
algorithm of that extension (Kerbl et al. 2023; cuda_rasterizer/forward.cu preprocessCUDA +
renderCUDA, rasterizer_impl.cu duplicateWithKeys / radix sort / identifyTileRanges):

  per gaussian   p_view = V p ; cull p_view.z <= 0.2 ; p_proj = (P V p).xyz / (w + 1e-7)
                 cov2D = (W J)^T Sigma (W J) with t.xy/t.z clamped to 1.3 tan(fov/2), + 0.3 on the diagonal
                 det = 0 -> cull ; conic = cov2D^-1 ; radius = ceil(3 sqrt(lambda_max)), lambda from
                 mid +- sqrt(max(0.1, mid^2 - det)) ; centre pixel = ((ndc + 1) * size - 1) / 2
                 tiles = 16x16-pixel blocks overlapped by centre +- radius ; none -> cull
  per tile       gaussians sorted by view depth (stable: ties keep gaussian order)
  per pixel      front to back: power = -1/2 d^T conic d ; skip if power > 0 ; alpha = min(0.99,
                 opacity * exp(power)) ; skip if alpha < 1/255 ; stop BEFORE a gaussian that would
                 push T below 1e-4 ; C += color * alpha * T ; T *= 1 - alpha ; out = C + T * bg
All arithmetic in float32, like the extension.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
TILE = 16


def quaternion_to_matrix(q: np.ndarray, eps: float = 1e-8) -> np.ndarray:
    """gaussian_splatting.py:115-137 (xyzw order)."""
    i, j, k, r = (f32(v) for v in q)
    two_s = f32(2.0) / (f32(np.sum(np.asarray(q, f32) ** 2)) + f32(eps))
    return np.array([[1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r)],
                     [two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r)],
                     [two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)]], dtype=f32)


def build_covariance(scale: np.ndarray, rotation_xyzw: np.ndarray) -> np.ndarray:
    """gaussian_splatting.py:140-151: R S S^T R^T."""
    S = np.diag(np.asarray(scale, f32))
    R = quaternion_to_matrix(rotation_xyzw)
    return (R @ S @ S.T @ R.T).astype(f32)


def get_fov(intrinsic: np.ndarray):
    """gaussian_splatting.py:179-195 (normalised intrinsics)."""
    inv = np.linalg.inv(intrinsic.astype(f32)).astype(f32)

    def unit(v):
        v = inv @ np.asarray(v, f32)
        return v / np.linalg.norm(v)
    fov_x = np.arccos(np.dot(unit([0, 0.5, 1]), unit([1, 0.5, 1])))
    fov_y = np.arccos(np.dot(unit([0.5, 0, 1]), unit([0.5, 1, 1])))
    return f32(fov_x), f32(fov_y)


def camera(extrinsic: np.ndarray, intrinsic: np.ndarray, near: float = 0.5, far: float = 1000.0):
    """The matrices render_cuda hands to the rasteriser (:236-243), as row-major world->camera `view`
    [4,4], camera->clip `proj` [4,4], and tan(fov/2)."""
    fov_x, fov_y = get_fov(intrinsic)
    tx, ty = f32(np.tan(f32(0.5) * fov_x)), f32(np.tan(f32(0.5) * fov_y))
    n, f = f32(near), f32(far)
    top, right = ty * n, tx * n
    proj = np.zeros((4, 4), f32)                                   # get_projection_matrix :198-226
    proj[0, 0] = 2 * n / (2 * right)
    proj[1, 1] = 2 * n / (2 * top)
    proj[3, 2] = 1
    proj[2, 2] = f / (f - n)
    proj[2, 3] = -(f * n) / (f - n)
    view = np.linalg.inv(extrinsic.astype(f32)).astype(f32)
    return view, proj, tx, ty


def preprocess(means, cov3d, opacity, view, proj, tx, ty, H, W):
    """forward.cu preprocessCUDA for all gaussians (vectorised).  Returns a dict of per-gaussian arrays;
    `radius == 0` marks culled gaussians."""
    N = means.shape[0]
    m = means.astype(f32)
    R, t = view[:3, :3], view[:3, 3]
    pv = (m @ R.T + t).astype(f32)                                  # p_view
    ok = pv[:, 2] > f32(0.2)
    full = (proj @ view).astype(f32)
    ph = (np.concatenate([m, np.ones((N, 1), f32)], 1) @ full.T).astype(f32)
    pw = f32(1.0) / (ph[:, 3] + f32(1e-7))
    pp = ph[:, :3] * pw[:, None]
    fx, fy = f32(W) / (f32(2.0) * tx), f32(H) / (f32(2.0) * ty)
    # computeCov2D
    limx, limy = f32(1.3) * tx, f32(1.3) * ty
    tz = pv[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        txz = np.clip(pv[:, 0] / tz, -limx, limx) * tz
        tyz = np.clip(pv[:, 1] / tz, -limy, limy) * tz
        J = np.zeros((N, 3, 3), f32)
        J[:, 0, 0] = fx / tz
        J[:, 0, 2] = -(fx * txz) / (tz * tz)
        J[:, 1, 1] = fy / tz
        J[:, 1, 2] = -(fy * tyz) / (tz * tz)
    T = J @ R                                                       # [N, 3, 3]: rows 0,1 matter
    cov = T @ cov3d.astype(f32) @ np.transpose(T, (0, 2, 1))
    a = cov[:, 0, 0] + f32(0.3)
    b = cov[:, 0, 1]
    c = cov[:, 1, 1] + f32(0.3)
    det = a * c - b * b
    ok &= det != 0
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = f32(1.0) / det
        conic = np.stack([c * inv, -b * inv, a * inv], 1).astype(f32)
        mid = f32(0.5) * (a + c)
        lam = mid + np.sqrt(np.maximum(f32(0.1), mid * mid - det))
        lam2 = mid - np.sqrt(np.maximum(f32(0.1), mid * mid - det))
        radius = np.ceil(f32(3.0) * np.sqrt(np.maximum(lam, lam2)))
    px = ((pp[:, 0] + f32(1.0)) * f32(W) - f32(1.0)) * f32(0.5)
    py = ((pp[:, 1] + f32(1.0)) * f32(H) - f32(1.0)) * f32(0.5)
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    radius = np.where(ok & np.isfinite(radius), radius, 0).astype(np.int64)
    pxs, pys = np.where(np.isfinite(px), px, 0), np.where(np.isfinite(py), py, 0)
    x0 = np.clip(((pxs - radius) / TILE).astype(np.int64), 0, gx)
    x1 = np.clip(((pxs + radius + TILE - 1) / TILE).astype(np.int64), 0, gx)
    y0 = np.clip(((pys - radius) / TILE).astype(np.int64), 0, gy)
    y1 = np.clip(((pys + radius + TILE - 1) / TILE).astype(np.int64), 0, gy)
    touched = (x1 - x0) * (y1 - y0)
    ok &= touched > 0
    radius = np.where(ok, radius, 0)
    return dict(depth=tz.astype(f32), xy=np.stack([px, py], 1).astype(f32), conic=conic,
                opacity=opacity.astype(f32), radius=radius, rect=np.stack([x0, y0, x1, y1], 1), grid=(gx, gy))


def render(means, colors, opacity, scale, rotation_xyzw, extrinsic, intrinsic, H, W, bg=(0.0, 0.0, 0.0)):
    """gs_render (gaussian_splatting.py:13-43): returns the float32 image [3, H, W]."""
    cov3d = build_covariance(np.asarray(scale, f32), np.asarray(rotation_xyzw, f32))
    view, proj, tx, ty = camera(extrinsic, intrinsic)
    g = preprocess(means, cov3d, opacity, view, proj, tx, ty, H, W)
    gx, gy = g["grid"]
    vis = np.nonzero(g["radius"] > 0)[0]
    # duplicateWithKeys + stable sort by (tile, depth): ties keep gaussian order
    tiles, ids = [], []
    for i in vis:
        x0, y0, x1, y1 = g["rect"][i]
        for y in range(y0, y1):
            for x in range(x0, x1):
                tiles.append(y * gx + x)
                ids.append(i)
    tiles, ids = np.asarray(tiles, np.int64), np.asarray(ids, np.int64)
    depth_bits = g["depth"][ids].view(np.uint32).astype(np.int64) if len(ids) else np.zeros(0, np.int64)
    order = np.lexsort((ids, depth_bits, tiles)) if len(ids) else np.zeros(0, np.int64)
    tiles, ids = tiles[order], ids[order]
    out = np.zeros((3, H, W), f32)
    bgc = np.asarray(bg, f32)
    col = colors.astype(f32)
    starts = np.searchsorted(tiles, np.arange(gx * gy), "left")
    ends = np.searchsorted(tiles, np.arange(gx * gy), "right")
    for tile in range(gx * gy):
        ty0, tx0 = (tile // gx) * TILE, (tile % gx) * TILE
        ys, xs = np.meshgrid(np.arange(ty0, min(ty0 + TILE, H)), np.arange(tx0, min(tx0 + TILE, W)), indexing="ij")
        pxf, pyf = xs.astype(f32), ys.astype(f32)
        Tr = np.ones(xs.shape, f32)
        C = np.zeros((3,) + xs.shape, f32)
        done = np.zeros(xs.shape, bool)
        for i in ids[starts[tile]:ends[tile]]:
            if done.all():
                break
            dx, dy = g["xy"][i, 0] - pxf, g["xy"][i, 1] - pyf
            con = g["conic"][i]
            power = f32(-0.5) * (con[0] * dx * dx + con[2] * dy * dy) - con[1] * dx * dy
            alpha = np.minimum(f32(0.99), g["opacity"][i] * np.exp(power.astype(f32)))
            skip = (power > 0) | (alpha < f32(1.0 / 255.0))
            test_T = Tr * (f32(1.0) - alpha)
            stop = ~skip & ~done & (test_T < f32(1e-4))
            done |= stop
            act = ~skip & ~done
            C += np.where(act, alpha * Tr, f32(0))[None] * col[i][:, None, None]
            Tr = np.where(act, test_T, Tr)
        out[:, ys, xs] = C + Tr[None] * bgc[:, None, None]
    return out


def render_with_gs(world_points, extrinsic, intrinsic, colors, H, W, scale=0.0001):
    """scripts/inference/infer.py:260-273: uint8 [H, W, 3] (colours 0..255 in, truncated out)."""
    c = colors.astype(f32) / f32(255.0) if colors.max() > 1.0 else colors.astype(f32)
    img = render(world_points, c, np.ones(world_points.shape[0], f32), [scale] * 3, [0.0, 0.0, 0.0, 1.0],
                 extrinsic, intrinsic, H, W)
    return (np.transpose(img, (1, 2, 0)) * f32(255.0)).astype(np.uint8)
