"""CPU oracle: torch restatement of the reference's causal Wan2.1 3D-VAE and the
Motion-Sensitive trajectory adaptors.

TEST INFRASTRUCTURE — see oracle/__init__.py for who may import this.

Parity status: the reference has no tests or golden vectors — "unpinned by the
reference's own tests".  This restatement is pinned against outputs of the reference modules
themselves (tests/golden/make_golden.py → tests/test_vae_oracle_vs_golden.py).

Formulation.  The reference runs the VAE chunk by chunk (1 frame, then 4-frame chunks on the
encoder side, 1 latent frame at a time on the decoder side) and threads a 2-frame cache
through every CausalConv3d (MoRe4D/models/wan_vae.py "vae":520-547, 678-703, 190-224).  That
is algebraically a causal convolution over the WHOLE sequence, with three special rules which
this oracle (and the CUDA path, which uses the same formulation) states explicitly:
  * CausalConv3d (vae:21-40): zero padding of 2 frames in front, none behind.
  * downsample3d (vae:148-163): the first frame passes through unchanged (first chunk only
    seeds the cache); output frame k >= 1 is the (3,1,1)/stride-2 conv over input frames
    2k-2, 2k-1, 2k.
  * upsample3d (vae:105-141, the 'Rep' sentinel): the first frame passes through un-doubled;
    for i >= 1 the (3,1,1) conv sees the sequence with frame 0 REPLACED BY ZEROS (the second
    chunk runs time_conv without a cache), and its 2C output channels are interleaved into
    frames 2i-1, 2i (vae:138-141).
Tensors are [1, C, T, H, W] like the reference.  `emulate_bf16` rounds after every op the way
the reference's pure-bf16 inference path does (pipeline_wan_fun_control.py:356-386 calls the
VAE outside autocast), with fp32 statistics inside norms.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

from more4d_b200.vae_arch import VAEConfig, WAN_VAE, decoder_layers, encoder_layers

Tensor = torch.Tensor


class Arith:
    def __init__(self, emulate_bf16: bool):
        self.emulate = emulate_bf16

    def r(self, x: Tensor) -> Tensor:
        return x.to(torch.bfloat16).to(torch.float32) if self.emulate else x


def causal_conv3d(x: Tensor, w: Tensor, b: Tensor, ar: Arith, stride_t: int = 1) -> Tensor:
    """CausalConv3d.forward (vae:32-40) over a whole sequence."""
    kt, kh, kw = w.shape[2:]
    x = F.pad(x, (kw // 2, kw // 2, kh // 2, kh // 2, kt - 1 if stride_t == 1 else 0, 0))
    return ar.r(F.conv3d(x, w.float(), b.float(), stride=(stride_t, 1, 1)))


def conv2d_frames(x: Tensor, w: Tensor, b: Tensor, ar: Arith, stride: int = 1, pad=(1, 1, 1, 1)) -> Tensor:
    """nn.Conv2d applied per frame ('b c t h w -> (b t) c h w', vae:143-145)."""
    B, C, T, H, W = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W)
    y = F.conv2d(F.pad(y, pad), w.float(), b.float(), stride=stride)
    return ar.r(y.reshape(B, T, *y.shape[1:]).permute(0, 2, 1, 3, 4))


def rms_norm(x: Tensor, gamma: Tensor, ar: Arith) -> Tensor:
    """RMS_norm.forward (vae:56-58): F.normalize over channels * sqrt(C) * gamma."""
    C = x.shape[1]
    n = ar.r(x.float().pow(2).sum(dim=1, keepdim=True).sqrt())
    y = ar.r(x / n.clamp_min(1e-12))
    y = ar.r(y * (C ** 0.5))
    return ar.r(y * gamma.float().reshape(1, C, 1, 1, 1))


def silu(x: Tensor, ar: Arith) -> Tensor:
    return ar.r(F.silu(x))


def residual_block(x: Tensor, sd, p: str, ar: Arith) -> Tensor:
    """ResidualBlock.forward (vae:206-224)."""
    h = x
    if (p + ".shortcut.weight") in sd:
        h = causal_conv3d(x, sd[p + ".shortcut.weight"], sd[p + ".shortcut.bias"], ar)
    y = silu(rms_norm(x, sd[p + ".residual.0.gamma"], ar), ar)
    y = causal_conv3d(y, sd[p + ".residual.2.weight"], sd[p + ".residual.2.bias"], ar)
    y = silu(rms_norm(y, sd[p + ".residual.3.gamma"], ar), ar)
    y = causal_conv3d(y, sd[p + ".residual.6.weight"], sd[p + ".residual.6.bias"], ar)
    return ar.r(y + h)


def attention_block(x: Tensor, sd, p: str, ar: Arith) -> Tensor:
    """AttentionBlock.forward (vae:244-266): per-frame single-head attention, head_dim = C."""
    B, C, T, H, W = x.shape
    y = rms_norm(x, sd[p + ".norm.gamma"].reshape(C, 1, 1, 1), ar)
    qkv = conv2d_frames(y, sd[p + ".to_qkv.weight"], sd[p + ".to_qkv.bias"], ar, pad=(0, 0, 0, 0))
    qkv = qkv.permute(0, 2, 3, 4, 1).reshape(B * T, H * W, 3 * C)
    q, k, v = qkv.chunk(3, dim=-1)
    s = torch.matmul(q, k.transpose(1, 2)) / math.sqrt(C)
    o = ar.r(torch.matmul(torch.softmax(s, dim=-1), v))                   # [BT, HW, C]
    o = o.reshape(B, T, H, W, C).permute(0, 4, 1, 2, 3)
    o = conv2d_frames(o, sd[p + ".proj.weight"], sd[p + ".proj.bias"], ar, pad=(0, 0, 0, 0))
    return ar.r(o + x)


def downsample(x: Tensor, sd, p: str, kind: str, ar: Arith) -> Tensor:
    """Resample.forward, modes downsample2d / downsample3d (vae:142-163)."""
    y = conv2d_frames(x, sd[p + ".resample.1.weight"], sd[p + ".resample.1.bias"], ar, stride=2,
                      pad=(0, 1, 0, 1))
    if kind == "down3d" and y.shape[2] > 1:
        w, b = sd[p + ".time_conv.weight"], sd[p + ".time_conv.bias"]
        rest = causal_conv3d(y, w, b, ar, stride_t=2)        # windows (0,1,2), (2,3,4), ...
        y = torch.cat([y[:, :, :1], rest], dim=2)
    return y


def upsample(x: Tensor, sd, p: str, kind: str, ar: Arith) -> Tensor:
    """Resample.forward, modes upsample2d / upsample3d (vae:105-145)."""
    B, C, T, H, W = x.shape
    if kind == "up3d" and T > 1:
        w, b = sd[p + ".time_conv.weight"], sd[p + ".time_conv.bias"]
        u = causal_conv3d(x[:, :, 1:], w, b, ar)             # frame 0 is invisible to time_conv
        u = u.reshape(B, 2, C, T - 1, H, W)
        u = torch.stack((u[:, 0], u[:, 1]), 3).reshape(B, C, 2 * (T - 1), H, W)
        x = torch.cat([x[:, :, :1], u], dim=2)
    T2 = x.shape[2]
    y = x.permute(0, 2, 1, 3, 4).reshape(B * T2, C, H, W)
    y = ar.r(F.interpolate(y.float(), scale_factor=(2.0, 2.0), mode="nearest-exact"))   # vae:61-67
    y = y.reshape(B, T2, C, 2 * H, 2 * W).permute(0, 2, 1, 3, 4)
    return conv2d_frames(y, sd[p + ".resample.1.weight"], sd[p + ".resample.1.bias"], ar)


def _run(x: Tensor, layers, sd, prefix: str, ar: Arith) -> Tensor:
    for kind, name, cin, cout in layers:
        p = prefix + name
        if kind == "conv":
            x = causal_conv3d(x, sd[p + ".weight"], sd[p + ".bias"], ar)
        elif kind == "res":
            x = residual_block(x, sd, p, ar)
        elif kind == "attn":
            x = attention_block(x, sd, p, ar)
        elif kind in ("down2d", "down3d"):
            x = downsample(x, sd, p, kind, ar)
        elif kind in ("up2d", "up3d"):
            x = upsample(x, sd, p, kind, ar)
        elif kind == "head":
            x = silu(rms_norm(x, sd[p + ".0.gamma"], ar), ar)
            x = causal_conv3d(x, sd[p + ".2.weight"], sd[p + ".2.bias"], ar)
    return x


def encode(x: Tensor, sd: Dict[str, Tensor], cfg: VAEConfig = WAN_VAE, emulate_bf16: bool = False,
           prefix: str = "model.") -> Tensor:
    """AutoencoderKLWan_.encode (vae:520-547) for one sample: x [1, 3, F, H, W] (F = 1 + 4k) ->
    [1, 2*z, 1 + k, H/8, W/8] = (normalised mu | log-variance), i.e. the parameters of the
    DiagonalGaussianDistribution whose .mode() is the latent."""
    ar = Arith(emulate_bf16)
    h = _run(ar.r(x.float()), encoder_layers(cfg), sd, prefix, ar)
    h = causal_conv3d(h, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"], ar)
    mu, logvar = h.chunk(2, dim=1)
    mean = ar.r(torch.tensor(cfg.mean)).view(1, -1, 1, 1, 1)
    inv_std = ar.r(1.0 / torch.tensor(cfg.std)).view(1, -1, 1, 1, 1)
    mu = ar.r(ar.r(mu - mean) * inv_std)
    return torch.cat([mu, logvar], dim=1)


def decode(z: Tensor, sd: Dict[str, Tensor], cfg: VAEConfig = WAN_VAE, emulate_bf16: bool = False,
           prefix: str = "model.") -> Tensor:
    """AutoencoderKLWan_.decode (vae:678-703) + the wrapper's clamp (vae:827): z [1, z, T, h, w]
    -> [1, 3, 4T-3, 8h, 8w] in [-1, 1]."""
    ar = Arith(emulate_bf16)
    mean = ar.r(torch.tensor(cfg.mean)).view(1, -1, 1, 1, 1)
    inv_std = ar.r(1.0 / torch.tensor(cfg.std)).view(1, -1, 1, 1, 1)
    z = ar.r(ar.r(ar.r(z.float()) / inv_std) + mean)
    h = causal_conv3d(z, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"], ar)
    h = _run(h, decoder_layers(cfg), sd, prefix, ar)
    return h.clamp(-1, 1)


# --------------------------------------------------------------------------------------
# trajectory adaptors (MoRe4D/models/trajectory_module.py "traj")
# --------------------------------------------------------------------------------------
def _group_norm_swish(x: Tensor, w: Tensor, b: Tensor, ar: Arith) -> Tensor:
    """Normalize = GroupNorm(32, eps 1e-6, affine) (traj:59-60) then x*sigmoid(x) (traj:54-56)."""
    y = ar.r(F.group_norm(x.float(), 32, w.float(), b.float(), 1e-6))
    return ar.r(y * ar.r(torch.sigmoid(y)))


def _conv2(x: Tensor, sd, p: str, ar: Arith) -> Tensor:
    return ar.r(F.conv2d(x, sd[p + ".weight"].float(), sd[p + ".bias"].float(), padding=1))


def _resnet_block(x: Tensor, sd, p: str, ar: Arith) -> Tensor:
    """ResnetBlock.forward with temb=None, in == out channels (traj:104-122)."""
    h = _conv2(_group_norm_swish(x, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], ar), sd, p + ".conv1", ar)
    h = _conv2(_group_norm_swish(h, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], ar), sd, p + ".conv2", ar)
    return ar.r(x + h)


def encoder_adaptor(x: Tensor, sd: Dict[str, Tensor], emulate_bf16: bool = False) -> Tensor:
    """VAEEncoderadaptor.forward (traj:177-196): [B, 3, F, H, W] -> pseudo-RGB in (0, 1)."""
    ar = Arith(emulate_bf16)
    B, C, Fr, H, W = x.shape
    x2 = ar.r(x.float()).permute(0, 2, 1, 3, 4).reshape(B * Fr, C, H, W)
    h = _conv2(x2, sd, "conv_in", ar)
    h = _resnet_block(h, sd, "down.0.block.0", ar)
    h = _group_norm_swish(h, sd["norm_out.weight"], sd["norm_out.bias"], ar)
    h = _conv2(h, sd, "conv_out", ar)
    h = ar.r(torch.sigmoid(ar.r(h + x2)))
    return h.view(B, Fr, C, H, W).permute(0, 2, 1, 3, 4)


def decoder_adaptor(z: Tensor, sd: Dict[str, Tensor], emulate_bf16: bool = False) -> Tensor:
    """VAEDecoderadaptor.forward (traj:260-279): reconstructed RGB -> xyz displacement."""
    ar = Arith(emulate_bf16)
    B, C, Fr, H, W = z.shape
    h = _conv2(ar.r(z.float()).permute(0, 2, 1, 3, 4).reshape(B * Fr, C, H, W), sd, "conv_in", ar)
    h = _resnet_block(h, sd, "up.0.block.0", ar)
    h = _resnet_block(h, sd, "up.0.block.1", ar)
    h = _group_norm_swish(h, sd["norm_out.weight"], sd["norm_out.bias"], ar)
    h = _conv2(h, sd, "conv_out", ar)
    return h.view(B, Fr, C, H, W).permute(0, 2, 1, 3, 4)


def roundtrip(x: Tensor, vae_sd, enc_sd, dec_sd, cfg: VAEConfig = WAN_VAE, emulate_bf16: bool = False):
    """The Motion-Sensitive VAE round trip of scripts/inference/infer_vae.py:276-281 with
    .mode() in place of .sample(): adaptor -> *2-1 -> encode -> decode -> adaptor."""
    ar = Arith(emulate_bf16)
    pseudo = ar.r(ar.r(encoder_adaptor(x, enc_sd, emulate_bf16) * 2) - 1)
    params = encode(pseudo, vae_sd, cfg, emulate_bf16)
    latent = params[:, :cfg.z_dim]
    recon = decode(latent, vae_sd, cfg, emulate_bf16)
    return decoder_adaptor(recon, dec_sd, emulate_bf16), latent, recon
