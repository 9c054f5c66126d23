"""Shape/config records for the 4D-STraG hot path.

The reference never states the checkpoint dims in-tree (SURVEY.md §8a "dims"); the presets
below are the standard Wan2.1 sizes it is built for.  Field names follow the constructor of
the reference's ``WanTransformer4DModel`` (MoRe4D/models/wan_transformer4d.py:792-821).
"""
from __future__ import annotations

from dataclasses import dataclass, replace


@dataclass(frozen=True)
class DiTConfig:
    model_type: str = "i2v"
    patch_size: tuple = (1, 2, 2)
    text_len: int = 512
    in_dim: int = 64            # 16 noisy + 16 control + 16 start-image + 16 depth (pctl:762-777)
    dim: int = 5120
    ffn_dim: int = 13824
    freq_dim: int = 256
    text_dim: int = 4096
    out_dim: int = 16
    num_heads: int = 40
    num_layers: int = 40
    qk_norm: bool = True
    cross_attn_norm: bool = True
    eps: float = 1e-6
    add_ref_conv: bool = True
    in_dim_ref_conv: int = 16
    clip_dim: int = 1280
    clip_tokens: int = 257
    use_spatial_guidance: bool = False   # Motion-Perception-Module branch (t4d:739-783)
    use_omnimae_guidance: bool = False   # + its front end: feature_adapter / resize / repeat (t4d:883-892,1127-1156)
    guidance_dim: int = 768

    @property
    def head_dim(self) -> int:
        return self.dim // self.num_heads

    def with_(self, **kw) -> "DiTConfig":
        return replace(self, **kw)


WAN_14B = DiTConfig()
WAN_1_3B = DiTConfig(dim=1536, ffn_dim=8960, num_heads=12, num_layers=30)
# small shapes for CPU tests; head_dim stays 128 (the one invariant the kernels specialise on)
# (clip_dim / clip_tokens stay 1280 / 257: the reference hard-codes both, t4d:520,938)
WAN_TINY = DiTConfig(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128,
                     text_len=32)
# 4D-ViSM stage (Wan2.1-Fun-InP backbone, WanTransformer3DModel): x 16 + mask 4 + masked-video 16
# input channels (pipeline_wan_fun_inpaint.py:707-713), no reference-frame conv
WAN_14B_INP = DiTConfig(in_dim=36, add_ref_conv=False)
WAN_TINY_INP = WAN_TINY.with_(in_dim=36, add_ref_conv=False)


def token_grid(frames: int, height: int, width: int, cfg: DiTConfig = WAN_14B,
               with_ref: bool = True):
    """(F, H, W) token grid for a `frames`x`height`x`width` video: VAE 4x/8x compression
    (wan_vae.py:753-756) then the (1,2,2) patch embed; +1 frame of reference tokens
    (t4d:1086-1090)."""
    lat_t = (frames - 1) // 4 + 1
    f = lat_t // cfg.patch_size[0] + (1 if with_ref else 0)
    return f, height // 8 // cfg.patch_size[1], width // 8 // cfg.patch_size[2]
