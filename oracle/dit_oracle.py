"""CPU oracle: torch restatement of the reference's Wan2.1-DiT forward (4D-STraG denoiser).

TEST INFRASTRUCTURE — see oracle/__init__.py for who may import this.

Parity status: the reference ships no tests or golden vectors so parity is
"unpinned by the reference's own tests"; this restatement is pinned against outputs of the
reference modules themselves (tests/golden/make_golden.py, run in the authoring container
through oracle/ref_import.py) — see tests/test_oracle_vs_golden.py.

Written functionally over a flat state dict with the reference's key names.  Citations are
to /root/reference/MoRe4D/models/wan_transformer4d.py ("t4d").

Two arithmetic modes:
  * ``emulate_bf16=False``  everything in fp32/fp64 — the gold oracle (SURVEY.md §8c mode A).
  * ``emulate_bf16=True``   rounds to bf16 exactly where the reference's CUDA-autocast path
    does (Linear/conv/attention outputs, RMSNorm's rstd and product, the ``.to(dtype)`` casts;
    residual stream, LayerNorm and modulation stay fp32 — SURVEY.md F7).  Used to compare the
    CUDA kernels tightly; the ≤1e-3 north-star tolerance is checked against the gold mode.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


class Arith:
    """Rounding policy."""

    def __init__(self, emulate_bf16: bool):
        self.emulate = emulate_bf16

    def r(self, x: Tensor) -> Tensor:
        """Round to bf16 (emulation mode) and come back to fp32."""
        return x.to(torch.bfloat16).to(torch.float32) if self.emulate else x

    def linear(self, x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
        """nn.Linear as CUDA autocast runs it: bf16 operands, fp32 accumulate, bf16 result."""
        y = F.linear(self.r(x.float()), w.float(), None if b is None else b.float())
        return self.r(y)


# --------------------------------------------------------------------------------------
# embeddings
# --------------------------------------------------------------------------------------
def sinusoidal_embedding(dim: int, position: Tensor) -> Tensor:
    """t4d:239-249 — float64 cos‖sin table of `position`."""
    half = dim // 2
    pos = position.to(torch.float64)
    inv = torch.pow(torch.tensor(10000.0, dtype=torch.float64),
                    -torch.arange(half, dtype=torch.float64) / half)
    ang = pos[:, None] * inv[None, :]
    return torch.cat([ang.cos(), ang.sin()], dim=1)


def rope_angles(max_len: int, head_dim: int, theta: float = 10000.0) -> Tensor:
    """Per-position rotation angles [max_len, head_dim/2] (float64).

    t4d:252-260 builds complex tables for three axes and t4d:928-935 concatenates them along
    the pair axis with widths d/2-2*(d/6), d/6, d/6 (22/21/21 pairs for d=128); each axis uses
    its own `dim` in the frequency exponent."""
    d = head_dim
    widths = [d - 4 * (d // 6), 2 * (d // 6), 2 * (d // 6)]
    cols = []
    pos = torch.arange(max_len, dtype=torch.float64)
    for wdt in widths:
        inv = 1.0 / torch.pow(torch.tensor(theta, dtype=torch.float64),
                              torch.arange(0, wdt, 2, dtype=torch.float64) / wdt)
        cols.append(pos[:, None] * inv[None, :])
    return torch.cat(cols, dim=1)


def rope_token_angles(grid: Tuple[int, int, int], head_dim: int) -> Tensor:
    """[F*H*W, head_dim/2] angles: frame table ‖ row table ‖ col table for each token in
    frame-major, then row, then column order (t4d:357-361)."""
    f, h, w = grid
    d = head_dim
    ang = rope_angles(1024, d)
    nf = d // 2 - 2 * (d // 6)
    nh = d // 6
    af = ang[:f, :nf].view(f, 1, 1, nf).expand(f, h, w, nf)
    ah = ang[:h, nf:nf + nh].view(1, h, 1, nh).expand(f, h, w, nh)
    aw = ang[:w, nf + nh:].view(1, 1, w, nh).expand(f, h, w, nh)
    return torch.cat([af, ah, aw], dim=-1).reshape(f * h * w, d // 2)


def rope_apply(x: Tensor, grids: Sequence[Tuple[int, int, int]], ar: Arith) -> Tensor:
    """t4d:340-369 — rotate adjacent pairs (x[2i], x[2i+1]) by the token's angle in float64;
    tokens beyond f*h*w pass through (t4d:365).  x: [B, L, N, D]."""
    B, L, N, D = x.shape
    out = []
    for b in range(B):
        grid = tuple(int(v) for v in grids[b])
        n_tok = grid[0] * grid[1] * grid[2]
        ang = rope_token_angles(grid, D)                       # [n_tok, D/2] f64
        c, s = ang.cos()[:, None, :], ang.sin()[:, None, :]
        xb = x[b, :n_tok].to(torch.float64).reshape(n_tok, N, D // 2, 2)
        re, im = xb[..., 0], xb[..., 1]
        rot = torch.stack([re * c - im * s, re * s + im * c], dim=-1).reshape(n_tok, N, D)
        out.append(torch.cat([rot.to(torch.float32), x[b, n_tok:].float()], dim=0))
    return ar.r(torch.stack(out))


# --------------------------------------------------------------------------------------
# norms
# --------------------------------------------------------------------------------------
def rms_norm(x: Tensor, weight: Tensor, eps: float, ar: Arith) -> Tensor:
    """WanRMSNorm t4d:378-394: x * rsqrt(mean(x^2)+eps).to(x.dtype) * weight over the FULL
    channel dim (all heads jointly).  Under CUDA autocast `pow` runs in fp32, the rstd is cast
    back to bf16 before the multiply, and both multiplies produce bf16."""
    xf = x.float()
    rstd = torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)
    return ar.r(ar.r(xf * ar.r(rstd)) * weight.float())


def layer_norm(x: Tensor, weight: Optional[Tensor], bias: Optional[Tensor], eps: float) -> Tensor:
    """WanLayerNorm t4d:397-407 — fp32 under autocast."""
    return F.layer_norm(x.float(), (x.shape[-1],),
                        None if weight is None else weight.float(),
                        None if bias is None else bias.float(), eps)


# --------------------------------------------------------------------------------------
# attention
# --------------------------------------------------------------------------------------
# The reference has two behaviours for `k_lens`: its default flash-attn varlen branch trims the
# keys of each sample to k_lens[b] (t4d:122-124), its SDPA branch ignores k_lens with a warning
# (t4d:222-225).  The CUDA kernels follow the default (flash) semantics; the golden vectors are
# produced through the SDPA branch (the only one that runs on CPU), so the golden tests flip
# this switch off for the cases where k_lens < Lk.
RESPECT_K_LENS = True


def attention(q: Tensor, k: Tensor, v: Tensor, k_lens: Optional[Sequence[int]], ar: Arith) -> Tensor:
    """t4d:175-236 — non-causal softmax(q k^T / sqrt(D)) v, layout [B, L, N, D] in and out.
    `k_lens` trims keys per sample as the flash varlen branch does (t4d:122-124); the SDPA
    branch ignores it (t4d:222-225) — identical whenever k_lens == Lk, which is how the golden
    vectors are generated."""
    B, Lq, N, D = q.shape
    scale = 1.0 / math.sqrt(D)
    outs = []
    for b in range(B):
        lk = k.shape[1] if (k_lens is None or not RESPECT_K_LENS) else int(k_lens[b])
        qb = q[b].float().transpose(0, 1)                    # [N, Lq, D]
        kb = k[b, :lk].float().transpose(0, 1)
        vb = v[b, :lk].float().transpose(0, 1)
        s = torch.matmul(qb, kb.transpose(1, 2)) * scale
        p = torch.softmax(s, dim=-1)
        outs.append(torch.matmul(p, vb).transpose(0, 1))     # [Lq, N, D]
    return ar.r(torch.stack(outs))


def self_attention(x: Tensor, sd: Dict[str, Tensor], p: str, num_heads: int, eps: float,
                   seq_lens: Sequence[int], grids, ar: Arith) -> Tensor:
    """WanSelfAttention.forward t4d:434-466.  x: [B, L, C] already cast to the compute dtype."""
    B, L, C = x.shape
    d = C // num_heads
    q = rms_norm(ar.linear(x, sd[p + "q.weight"], sd[p + "q.bias"]), sd[p + "norm_q.weight"], eps, ar)
    k = rms_norm(ar.linear(x, sd[p + "k.weight"], sd[p + "k.bias"]), sd[p + "norm_k.weight"], eps, ar)
    v = ar.linear(x, sd[p + "v.weight"], sd[p + "v.bias"])
    q = rope_apply(q.view(B, L, num_heads, d), grids, ar)
    k = rope_apply(k.view(B, L, num_heads, d), grids, ar)
    o = attention(q, k, v.view(B, L, num_heads, d), seq_lens, ar)
    return ar.linear(o.flatten(2), sd[p + "o.weight"], sd[p + "o.bias"])


def cross_attention(x: Tensor, context: Tensor, sd: Dict[str, Tensor], p: str, num_heads: int,
                    eps: float, ar: Arith, clip_tokens: int = 257) -> Tensor:
    """WanI2VCrossAttention.forward t4d:515-554 when `k_img` weights are present (two
    independently normalised attentions over the first 257 image tokens and the remaining
    text tokens, summed), else WanT2VCrossAttention t4d:471-497.  No key mask: the model
    passes context_lens=None (t4d:1174)."""
    B, L, C = x.shape
    d = C // num_heads
    q = rms_norm(ar.linear(x, sd[p + "q.weight"], sd[p + "q.bias"]), sd[p + "norm_q.weight"], eps, ar)
    q = q.view(B, L, num_heads, d)
    has_img = (p + "k_img.weight") in sd
    ctx_txt = context[:, clip_tokens:] if has_img else context
    k = rms_norm(ar.linear(ctx_txt, sd[p + "k.weight"], sd[p + "k.bias"]), sd[p + "norm_k.weight"], eps, ar)
    v = ar.linear(ctx_txt, sd[p + "v.weight"], sd[p + "v.bias"])
    o = attention(q, k.view(B, -1, num_heads, d), v.view(B, -1, num_heads, d), None, ar)
    if has_img:
        ctx_img = context[:, :clip_tokens]
        ki = rms_norm(ar.linear(ctx_img, sd[p + "k_img.weight"], sd[p + "k_img.bias"]),
                      sd[p + "norm_k_img.weight"], eps, ar)
        vi = ar.linear(ctx_img, sd[p + "v_img.weight"], sd[p + "v_img.bias"])
        oi = attention(q, ki.view(B, -1, num_heads, d), vi.view(B, -1, num_heads, d), None, ar)
        o = ar.r(o + oi)                                     # bf16 + bf16 (t4d:552)
    return ar.linear(o.flatten(2), sd[p + "o.weight"], sd[p + "o.bias"])


def spatial_guidance(x: Tensor, feats: Tensor, cls: Optional[Tensor], sd, p: str, ar: Arith,
                     use_cls_token: bool = False) -> Tensor:
    """SpatialGuidanceModule.forward t4d:757-783 (Motion-Perception-Module injection):
    x*(1+scale*gate)+shift*gate with (scale, shift) = Linear(SiLU(features)), zero-padded to L."""
    src = cls if (use_cls_token and cls is not None) else feats
    sp = ar.linear(ar.r(F.silu(src.float())), sd[p + "spatial_guide.1.weight"],
                   sd[p + "spatial_guide.1.bias"])
    scale, shift = sp.chunk(2, dim=-1)
    if use_cls_token and cls is not None:
        scale = scale.repeat(1, feats.size(1), 1)
        shift = shift.repeat(1, feats.size(1), 1)
    if scale.size(1) < x.size(1):
        pad = x.size(1) - scale.size(1)
        scale = F.pad(scale, (0, 0, 0, pad))
        shift = F.pad(shift, (0, 0, 0, pad))
    gate = sd[p + "gate"].float()[None, None, :]
    # scale/shift/gate are bf16 tensors on the reference's autocast path, so the two products
    # and the (1 + .) are rounded to bf16 before they meet the fp32 activations
    return x * ar.r(1 + ar.r(scale * gate)) + ar.r(shift * gate)


def mpm_front_end(tokens: Tensor, cls: Tensor, sd, target_hw: Tuple[int, int], latent_T: int,
                  ar: Arith, return_adapter: bool = False):
    """The Motion-Perception front end between the (out-of-scope) OmniMAE trunk and the blocks,
    t4d:1146-1152: tokens [B, 196, 768] -> view [B, 768, 14, 14] -> feature_adapter
    (Conv2d 3x3 -> SiLU -> Conv2d 3x3, t4d:888-892) -> F.interpolate(size, 'bilinear',
    align_corners=False) -> repeat over latent_T -> [B, T*h*w, 768]; cls -> [B, 1, 768].
    bf16 emulation rounds where the reference's bf16 tensors would (each conv / SiLU /
    interpolate output)."""
    B = tokens.shape[0]
    G = tokens.shape[-1]
    x = ar.r(tokens.float()).view(B, 14, 14, G).permute(0, 3, 1, 2)
    x = ar.r(F.conv2d(x, sd["feature_adapter.0.weight"].float(), sd["feature_adapter.0.bias"].float(), padding=1))
    x = ar.r(F.silu(x))
    x = ar.r(F.conv2d(x, sd["feature_adapter.2.weight"].float(), sd["feature_adapter.2.bias"].float(), padding=1))
    adapter = x
    x = ar.r(F.interpolate(x, size=tuple(target_hw), mode="bilinear", align_corners=False))
    x = x.unsqueeze(2).repeat(1, 1, latent_T, 1, 1).flatten(2).transpose(1, 2)
    feats = (x.contiguous(), cls.float().view(B, 1, G))
    return (feats, adapter) if return_adapter else feats


# --------------------------------------------------------------------------------------
# block / head / model
# --------------------------------------------------------------------------------------
def block_forward(x: Tensor, e0: Tensor, sd: Dict[str, Tensor], num_heads: int, eps: float,
                  seq_lens, grids, context: Tensor, emulate_bf16: bool = False,
                  guidance: Optional[Tuple[Tensor, Optional[Tensor]]] = None,
                  use_cls_token: bool = False, prefix: str = "") -> Tensor:
    """WanAttentionBlock.forward t4d:633-688.  x: [B, L, C]; e0: [B, 6, C] fp32; returns fp32
    (the residual stream is fp32 from the first gated add on — SURVEY.md F7)."""
    ar = Arith(emulate_bf16)
    p = prefix
    x = x.float()
    e = (sd[p + "modulation"].float() + e0.float()).chunk(6, dim=1)      # t4d:659
    t = layer_norm(x, None, None, eps) * (1 + e[1]) + e[0]                 # t4d:662
    if guidance is not None and (p + "spatial_guidance_self.gate") in sd:
        t = spatial_guidance(t, guidance[0], guidance[1], sd, p + "spatial_guidance_self.", ar,
                             use_cls_token)
    y = self_attention(ar.r(t), sd, p + "self_attn.", num_heads, eps, seq_lens, grids, ar)
    x = x + y * e[2]                                                        # t4d:669
    if (p + "norm3.weight") in sd:
        n3 = layer_norm(x, sd[p + "norm3.weight"], sd[p + "norm3.bias"], eps)
    else:
        n3 = x
    x = x + cross_attention(ar.r(n3), context, sd, p + "cross_attn.", num_heads, eps, ar)  # t4d:674
    t = layer_norm(x, None, None, eps) * (1 + e[4]) + e[3]                 # t4d:677
    if guidance is not None and (p + "spatial_guidance_ffn.gate") in sd:
        t = spatial_guidance(t, guidance[0], guidance[1], sd, p + "spatial_guidance_ffn.", ar,
                             use_cls_token)
    h = ar.linear(ar.r(t), sd[p + "ffn.0.weight"], sd[p + "ffn.0.bias"])
    h = ar.r(F.gelu(h, approximate="tanh"))
    y = ar.linear(h, sd[p + "ffn.2.weight"], sd[p + "ffn.2.bias"])
    return x + y * e[5]                                                     # t4d:684


def head_forward(x: Tensor, e: Tensor, sd, eps: float, ar: Arith) -> Tensor:
    """Head.forward t4d:708-721 with e: [B, C]."""
    m = (sd["head.modulation"].float() + e.float().unsqueeze(1)).chunk(2, dim=1)
    t = layer_norm(x, None, None, eps) * (1 + m[1]) + m[0]
    return ar.linear(t, sd["head.head.weight"], sd["head.head.bias"])


def time_embed(t: Tensor, sd, freq_dim: int, dim: int) -> Tuple[Tensor, Tensor]:
    """t4d:1160-1171 — runs under autocast(float32): fp32 GEMMs on the (bf16-valued) weights."""
    s = sinusoidal_embedding(freq_dim, t).float()
    h = F.silu(F.linear(s, sd["time_embedding.0.weight"].float(), sd["time_embedding.0.bias"].float()))
    e = F.linear(h, sd["time_embedding.2.weight"].float(), sd["time_embedding.2.bias"].float())
    e0 = F.linear(F.silu(e), sd["time_projection.1.weight"].float(),
                  sd["time_projection.1.bias"].float())
    return e, e0.unflatten(1, (6, dim))


def context_embed(context: List[Tensor], clip_fea: Optional[Tensor], sd, text_len: int,
                  ar: Arith) -> Tensor:
    """text_embedding over zero-padded prompts (t4d:1175-1180) and MLPProj over CLIP tokens
    (t4d:724-736, 1182-1184); image tokens first."""
    ctx = torch.stack([F.pad(u.float(), (0, 0, 0, text_len - u.size(0))) for u in context])
    h = ar.linear(ctx, sd["text_embedding.0.weight"], sd["text_embedding.0.bias"])
    h = ar.r(F.gelu(h, approximate="tanh"))
    ctx = ar.linear(h, sd["text_embedding.2.weight"], sd["text_embedding.2.bias"])
    if clip_fea is not None and "img_emb.proj.1.weight" in sd:
        c = layer_norm(clip_fea, sd["img_emb.proj.0.weight"], sd["img_emb.proj.0.bias"], 1e-5)
        c = ar.linear(c, sd["img_emb.proj.1.weight"], sd["img_emb.proj.1.bias"])
        c = ar.r(F.gelu(c))                                   # exact (erf) GELU, t4d:731
        c = ar.linear(c, sd["img_emb.proj.3.weight"], sd["img_emb.proj.3.bias"])
        c = ar.r(layer_norm(c, sd["img_emb.proj.4.weight"], sd["img_emb.proj.4.bias"], 1e-5))
        ctx = torch.cat([c, ctx], dim=1)
    return ctx


def patch_embed(u: Tensor, w: Tensor, b: Tensor, ar: Arith) -> Tensor:
    """patch_embedding Conv3d with kernel=stride=(1,2,2) (t4d:898-899,1073) as the GEMM it is:
    u [Cin, T, H, W] -> tokens [T*(H/2)*(W/2), C], token order frame, row, col."""
    cin, T, H, W = u.shape
    pt, ph, pw = w.shape[2:]
    cols = u.float().reshape(cin, T // pt, pt, H // ph, ph, W // pw, pw)
    cols = cols.permute(1, 3, 5, 0, 2, 4, 6).reshape(-1, cin * pt * ph * pw)
    return ar.linear(cols, w.reshape(w.shape[0], -1), b)


def unpatchify(tok: Tensor, grid: Tuple[int, int, int], patch, out_dim: int) -> Tensor:
    """t4d:1343-1366 — [L, pt*ph*pw*c] -> [c, F*pt, H*ph, W*pw]."""
    f, h, w = grid
    pt, ph, pw = patch
    u = tok[: f * h * w].reshape(f, h, w, pt, ph, pw, out_dim)
    return u.permute(6, 0, 3, 1, 4, 2, 5).reshape(out_dim, f * pt, h * ph, w * pw)


def dit_forward(sd: Dict[str, Tensor], cfg, x: Tensor, t: Tensor, context: List[Tensor],
                seq_len: int, clip_fea: Optional[Tensor] = None, y: Optional[Tensor] = None,
                full_ref: Optional[Tensor] = None, emulate_bf16: bool = False,
                guidance=None, return_tokens: bool = False) -> Tensor:
    """WanTransformer4DModel.forward t4d:1046-1340 (no TeaCache / SP / control adapter /
    subject_ref).  x: [B, 16, T, h, w]; y: [B, 48, T, h, w]; returns [B, 16, T, h, w]."""
    ar = Arith(emulate_bf16)
    B = x.shape[0]
    if y is not None:
        x = torch.cat([x, y], dim=1)                                   # t4d:1069-1070
    pt, ph, pw = cfg.patch_size
    T, H, W = x.shape[2:]
    grid = (T // pt, H // ph, W // pw)
    toks = [patch_embed(x[b], sd["patch_embedding.weight"], sd["patch_embedding.bias"], ar)
            for b in range(B)]
    ref_len = 0
    if full_ref is not None and "ref_conv.weight" in sd:               # t4d:1086-1090
        wr = sd["ref_conv.weight"]
        ref = [patch_embed(full_ref[b].unsqueeze(1), wr.unsqueeze(2), sd["ref_conv.bias"], ar)
               for b in range(B)]
        ref_len = ref[0].shape[0]
        toks = [torch.cat([r, u], dim=0) for r, u in zip(ref, toks)]
        grid = (grid[0] + 1, grid[1], grid[2])
        seq_len = seq_len + ref_len
    seq_lens = [u.shape[0] for u in toks]
    assert max(seq_lens) <= seq_len
    xs = torch.stack([F.pad(u, (0, 0, 0, seq_len - u.shape[0])) for u in toks])   # t4d:1103-1106
    e, e0 = time_embed(t, sd, cfg.freq_dim, cfg.dim)
    ctx = context_embed(context, clip_fea, sd, cfg.text_len, ar)
    grids = [grid] * B
    for i in range(cfg.num_layers):
        xs = block_forward(xs, e0, sd, cfg.num_heads, cfg.eps, seq_lens, grids, ctx,
                           emulate_bf16, guidance=guidance, prefix=f"blocks.{i}.")
    out = head_forward(xs, e, sd, cfg.eps, ar)
    if return_tokens:
        return out
    out = out[:, ref_len:]
    grid = (grid[0] - (1 if ref_len else 0), grid[1], grid[2])
    return torch.stack([unpatchify(out[b], grid, cfg.patch_size, cfg.out_dim) for b in range(B)])
