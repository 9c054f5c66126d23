"""Synthetic weights and inputs for the hot path (no checkpoints ship with the reference and
there is no network — SURVEY.md F5).

Every tensor is drawn from its own generator seeded by (seed, crc32(key)), so a single key
can be regenerated anywhere (CPU for parity tests, CUDA for the bench) without materialising
the rest.  Key names are the reference's state-dict keys so the same dict
loads into the reference modules, the CPU oracle and the CUDA host mirror.

The reference zero-initialises several layers (head.head.weight t4d:1390, the MPM's last
linear and gate t4d:750-755, VAE AttentionBlock.proj vae:242, VAEEncoderadaptor.conv_out
traj:170); all of them get non-zero values here, otherwise parity would be vacuous (F6).
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterator, Tuple

import torch

from .config import DiTConfig


def _gen(seed: int, key: str, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFFFFFFFFFF)
    return g


def _randn(seed, key, shape, std, device, dtype, mean=0.0):
    g = _gen(seed, key, device)
    t = torch.randn(shape, generator=g, device=device, dtype=torch.float32) * std + mean
    return t.to(dtype)


def dit_param_specs(cfg: DiTConfig) -> Iterator[Tuple[str, tuple, float, float]]:
    """Yield (key, shape, std, mean) for every parameter of the reference DiT."""
    C, F = cfg.dim, cfg.ffn_dim
    w = 0.02
    pt, ph, pw = cfg.patch_size
    yield "patch_embedding.weight", (C, cfg.in_dim, pt, ph, pw), 0.05, 0.0
    yield "patch_embedding.bias", (C,), w, 0.0
    yield "text_embedding.0.weight", (C, cfg.text_dim), w, 0.0
    yield "text_embedding.0.bias", (C,), w, 0.0
    yield "text_embedding.2.weight", (C, C), w, 0.0
    yield "text_embedding.2.bias", (C,), w, 0.0
    yield "time_embedding.0.weight", (C, cfg.freq_dim), w, 0.0
    yield "time_embedding.0.bias", (C,), w, 0.0
    yield "time_embedding.2.weight", (C, C), w, 0.0
    yield "time_embedding.2.bias", (C,), w, 0.0
    yield "time_projection.1.weight", (6 * C, C), w, 0.0
    yield "time_projection.1.bias", (6 * C,), w, 0.0
    for i in range(cfg.num_layers):
        p = f"blocks.{i}."
        yield p + "modulation", (1, 6, C), C ** -0.5, 0.0
        for name in ("q", "k", "v", "o"):
            yield p + f"self_attn.{name}.weight", (C, C), w, 0.0
            yield p + f"self_attn.{name}.bias", (C,), w, 0.0
        yield p + "self_attn.norm_q.weight", (C,), 0.1, 1.0
        yield p + "self_attn.norm_k.weight", (C,), 0.1, 1.0
        if cfg.cross_attn_norm:
            yield p + "norm3.weight", (C,), 0.1, 1.0
            yield p + "norm3.bias", (C,), w, 0.0
        names = ("q", "k", "v", "o") + (("k_img", "v_img") if cfg.model_type == "i2v" else ())
        for name in names:
            yield p + f"cross_attn.{name}.weight", (C, C), w, 0.0
            yield p + f"cross_attn.{name}.bias", (C,), w, 0.0
        yield p + "cross_attn.norm_q.weight", (C,), 0.1, 1.0
        yield p + "cross_attn.norm_k.weight", (C,), 0.1, 1.0
        if cfg.model_type == "i2v":
            yield p + "cross_attn.norm_k_img.weight", (C,), 0.1, 1.0
        yield p + "ffn.0.weight", (F, C), w, 0.0
        yield p + "ffn.0.bias", (F,), w, 0.0
        yield p + "ffn.2.weight", (C, F), w, 0.0
        yield p + "ffn.2.bias", (C,), w, 0.0
        if cfg.use_spatial_guidance:
            for s in ("spatial_guidance_self", "spatial_guidance_ffn"):
                yield p + f"{s}.spatial_guide.1.weight", (2 * C, cfg.guidance_dim), w, 0.0
                yield p + f"{s}.spatial_guide.1.bias", (2 * C,), w, 0.0
                yield p + f"{s}.gate", (C,), 0.5, 0.0
    yield "head.head.weight", (cfg.out_dim * pt * ph * pw, C), w, 0.0
    yield "head.head.bias", (cfg.out_dim * pt * ph * pw,), w, 0.0
    yield "head.modulation", (1, 2, C), C ** -0.5, 0.0
    if cfg.model_type == "i2v":
        D = cfg.clip_dim
        yield "img_emb.proj.0.weight", (D,), 0.1, 1.0
        yield "img_emb.proj.0.bias", (D,), w, 0.0
        yield "img_emb.proj.1.weight", (D, D), w, 0.0
        yield "img_emb.proj.1.bias", (D,), w, 0.0
        yield "img_emb.proj.3.weight", (C, D), w, 0.0
        yield "img_emb.proj.3.bias", (C,), w, 0.0
        yield "img_emb.proj.4.weight", (C,), 0.1, 1.0
        yield "img_emb.proj.4.bias", (C,), w, 0.0
    if cfg.add_ref_conv:
        yield "ref_conv.weight", (C, cfg.in_dim_ref_conv, ph, pw), 0.05, 0.0
        yield "ref_conv.bias", (C,), w, 0.0
    if cfg.use_omnimae_guidance:
        G = cfg.guidance_dim
        for i in (0, 2):
            yield f"feature_adapter.{i}.weight", (G, G, 3, 3), (9 * G) ** -0.5, 0.0
            yield f"feature_adapter.{i}.bias", (G,), 0.1, 0.0


def dit_state_dict(cfg: DiTConfig, seed: int = 0, device="cpu", dtype=torch.bfloat16,
                   prefix_filter: str | None = None) -> Dict[str, torch.Tensor]:
    """Synthetic state dict (bf16-valued) for the reference DiT layout."""
    out = {}
    for key, shape, std, mean in dit_param_specs(cfg):
        if prefix_filter is not None and not key.startswith(prefix_filter):
            continue
        out[key] = _randn(seed, key, shape, std, device, dtype, mean)
    return out


def block_state_dict(cfg: DiTConfig, layer: int = 0, seed: int = 0, device="cpu",
                     dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    """State dict of one WanAttentionBlock with the ``blocks.N.`` prefix stripped."""
    p = f"blocks.{layer}."
    sd = dit_state_dict(cfg, seed, device, dtype, prefix_filter=p)
    return {k[len(p):]: v for k, v in sd.items()}


def dit_inputs(cfg: DiTConfig, grid: Tuple[int, int, int], batch: int = 1, seed: int = 0,
               device="cpu", dtype=torch.bfloat16, text_tokens=(7, 11), with_ref: bool = True):
    """Synthetic forward inputs with the reference call-surface layouts (t4d:1046-1060).

    `grid` is the token grid (F, H, W) *including* the reference frame when `with_ref`.
    Returns dict(x, y, t, context, clip_fea, full_ref, seq_len)."""
    f, h, w = grid
    lat_t = (f - 1 if with_ref else f) * cfg.patch_size[0]
    lh, lw = h * cfg.patch_size[1], w * cfg.patch_size[2]
    x = _randn(seed, "in.x", (batch, cfg.out_dim, lat_t, lh, lw), 1.0, device, dtype)
    y = _randn(seed, "in.y", (batch, cfg.in_dim - cfg.out_dim, lat_t, lh, lw), 1.0, device, dtype)
    context = [
        _randn(seed, f"in.ctx{i}", (text_tokens[i % len(text_tokens)], cfg.text_dim), 1.0,
               device, dtype) for i in range(batch)
    ]
    clip_fea = _randn(seed, "in.clip", (batch, cfg.clip_tokens, cfg.clip_dim), 1.0, device, dtype)
    full_ref = _randn(seed, "in.ref", (batch, cfg.in_dim_ref_conv, lh, lw), 1.0, device, dtype) \
        if with_ref else None
    t = torch.full((batch,), 500.0, device=device)
    seq_len = lat_t // cfg.patch_size[0] * h * w
    return dict(x=x, y=y, t=t, context=context, clip_fea=clip_fea, full_ref=full_ref,
                seq_len=seq_len)


def vae_state_dict(cfg=None, seed: int = 0, device="cpu", dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    """Synthetic state dict of the reference AutoencoderKLWan (keys with the `model.` prefix)."""
    from .vae_arch import WAN_VAE, vae_param_specs
    return {k: _randn(seed, k, shape, std, device, dtype, mean)
            for k, shape, std, mean in vae_param_specs(cfg or WAN_VAE)}


def adaptor_state_dict(kind: str, seed: int = 0, device="cpu", dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    """Synthetic state dict of VAEEncoderadaptor ("encoder") / VAEDecoderadaptor ("decoder")."""
    from .vae_arch import adaptor_param_specs
    return {k: _randn(seed, f"{kind}.{k}", shape, std, device, dtype, mean)
            for k, shape, std, mean in adaptor_param_specs(kind)}


def trajectory_video(frames: int, height: int, width: int, seed: int = 0, device="cpu",
                     dtype=torch.bfloat16) -> torch.Tensor:
    """Synthetic normalised xyz-displacement tensor [1, 3, F, H, W] shaped like the output of
    scripts/inference/infer_vae.py:98-134 (smooth, frame 0 = 0, grows with t)."""
    g = _gen(seed, "traj.video", "cpu")
    low = torch.randn(1, 3, max(2, frames // 4 + 1), max(2, height // 16), max(2, width // 16),
                      generator=g)
    smooth = torch.nn.functional.interpolate(low, size=(frames, height, width), mode="trilinear",
                                             align_corners=True)
    ramp = torch.linspace(0, 1, frames).view(1, 1, frames, 1, 1)
    return (smooth * ramp * 0.5).to(device=device, dtype=dtype)


def fill_module_(module, cfg: DiTConfig, seed: int = 0) -> None:
    """Fill a more4d_b200.dit.WanTransformer4DModel in place, one parameter at a time, on the
    parameter's own device (no second copy of a 14B state dict)."""
    sd = module.state_dict()
    specs = {k: (shape, std, mean) for k, shape, std, mean in dit_param_specs(cfg)}
    missing = set(sd) - set(specs)
    if missing:
        raise KeyError(f"no synthetic spec for parameters: {sorted(missing)[:5]} ...")
    with torch.no_grad():
        for key, p in sd.items():
            shape, std, mean = specs[key]
            p.copy_(_randn(seed, key, shape, std, p.device, p.dtype, mean))


def point_cloud(H: int, W: int, seed: int = 0, tilt: float = 0.0):
    """Synthetic stage-1 output for the hand-off renderer (scripts/inference/infer.py:398-444): one
    3-D point per pixel of an H x W frame, back-projected with random depth, integer-valued
    colours (uint8 image), a camera translated / rotated by `tilt` so that points collide in the
    z-buffer, leave the frustum and land behind the camera.  Returns (points [N,3] fp32,
    colors [N,3] fp32 in 0..255, extrinsic [4,4] cam->world, intrinsic [3,3])."""
    import math
    g = torch.Generator().manual_seed(1000 + seed)
    from .render import get_intrinsic_matrix
    K = get_intrinsic_matrix(H, W)
    v, u = torch.meshgrid((torch.arange(H) + 0.5) / H, (torch.arange(W) + 0.5) / W, indexing="ij")
    z = 1.0 + 4.0 * torch.rand(H, W, generator=g)
    z[torch.rand(H, W, generator=g) < 0.02] *= -1.0                   # a few points behind the camera
    x = (u - K[0, 2]) / K[0, 0] * z
    y = (v - K[1, 2]) / K[1, 1] * z
    pts = torch.stack([x, y, z], -1).reshape(-1, 3).float()
    pts[: W] = pts[W: 2 * W]                                          # exact depth ties on shared pixels
    colors = torch.randint(0, 256, (H * W, 3), generator=g).float()
    c, s_ = math.cos(tilt), math.sin(tilt)
    ext = torch.tensor([[c, 0, s_, 0.3 * tilt], [0, 1, 0, -0.1 * tilt], [-s_, 0, c, 0.2 * tilt], [0, 0, 0, 1]],
                       dtype=torch.float32)
    return pts, colors, ext, K

