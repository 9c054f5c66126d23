"""Layer program of the Wan2.1 causal 3D-VAE and the trajectory adaptors.

The reference builds these as nested nn.Modules (MoRe4D/models/wan_vae.py:269-476 Encoder3d /
Decoder3d, :727-745 `_video_vae`; MoRe4D/models/trajectory_module.py:125-279).  Here the same
architecture is a flat list of (kind, state-dict prefix, channels...) records shared by the
synthetic-weight factory, the CPU oracle and the CUDA host mirror, so the three cannot drift.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator, List, Tuple


@dataclass(frozen=True)
class VAEConfig:
    dim: int = 96
    z_dim: int = 16
    dim_mult: tuple = (1, 2, 4, 4)
    num_res_blocks: int = 2
    temporal_downsample: tuple = (False, True, True)     # `_video_vae` vae:737
    # latent normalisation constants of AutoencoderKLWan (vae:758-768)
    mean: tuple = (-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508,
                   0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921)
    std: tuple = (2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743,
                  3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160)


WAN_VAE = VAEConfig()
# Each record: (kind, prefix, cin, cout)
#   conv      CausalConv3d 3x3x3                         res      ResidualBlock (vae:190-224)
#   attn      AttentionBlock (vae:227-266)               down2d / down3d / up2d / up3d  Resample
#   head      RMS_norm + SiLU + CausalConv3d 3x3x3
Layer = Tuple[str, str, int, int]


def encoder_layers(cfg: VAEConfig = WAN_VAE) -> List[Layer]:
    dims = [cfg.dim * u for u in (1,) + tuple(cfg.dim_mult)]
    out: List[Layer] = [("conv", "encoder.conv1", 3, dims[0])]
    idx = 0
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        for _ in range(cfg.num_res_blocks):
            out.append(("res", f"encoder.downsamples.{idx}", cin, cout))
            cin = cout
            idx += 1
        if i != len(cfg.dim_mult) - 1:
            kind = "down3d" if cfg.temporal_downsample[i] else "down2d"
            out.append((kind, f"encoder.downsamples.{idx}", cout, cout))
            idx += 1
    c = dims[-1]
    out += [("res", "encoder.middle.0", c, c), ("attn", "encoder.middle.1", c, c),
            ("res", "encoder.middle.2", c, c), ("head", "encoder.head", c, cfg.z_dim * 2)]
    return out


def decoder_layers(cfg: VAEConfig = WAN_VAE) -> List[Layer]:
    dm = tuple(cfg.dim_mult)
    dims = [cfg.dim * u for u in (dm[-1],) + dm[::-1]]
    temporal_up = tuple(cfg.temporal_downsample[::-1])
    c0 = dims[0]
    out: List[Layer] = [("conv", "decoder.conv1", cfg.z_dim, c0),
                        ("res", "decoder.middle.0", c0, c0), ("attn", "decoder.middle.1", c0, c0),
                        ("res", "decoder.middle.2", c0, c0)]
    idx = 0
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        if i in (1, 2, 3):
            cin = cin // 2                                  # vae:409-410: upsample halves channels
        for _ in range(cfg.num_res_blocks + 1):
            out.append(("res", f"decoder.upsamples.{idx}", cin, cout))
            cin = cout
            idx += 1
        if i != len(dm) - 1:
            kind = "up3d" if temporal_up[i] else "up2d"
            out.append((kind, f"decoder.upsamples.{idx}", cout, cout // 2))
            idx += 1
    out.append(("head", "decoder.head", dims[-1], 3))
    return out


def vae_param_specs(cfg: VAEConfig = WAN_VAE, prefix: str = "model.") -> Iterator[Tuple[str, tuple, float, float]]:
    """(key, shape, std, mean) for every parameter of the reference AutoencoderKLWan
    (keys as in its state dict, i.e. with the `model.` prefix of vae:771)."""
    def conv3(name, cout, cin, k=(3, 3, 3)):
        fan = cin * k[0] * k[1] * k[2]
        yield prefix + name + ".weight", (cout, cin) + tuple(k), fan ** -0.5, 0.0
        yield prefix + name + ".bias", (cout,), 0.05, 0.0

    def conv2(name, cout, cin, k=3):
        yield prefix + name + ".weight", (cout, cin, k, k), (cin * k * k) ** -0.5, 0.0
        yield prefix + name + ".bias", (cout,), 0.05, 0.0

    def layer(kind, name, cin, cout):
        if kind == "conv":
            yield from conv3(name, cout, cin)
        elif kind == "res":
            yield prefix + name + ".residual.0.gamma", (cin, 1, 1, 1), 0.1, 1.0
            yield from conv3(name + ".residual.2", cout, cin)
            yield prefix + name + ".residual.3.gamma", (cout, 1, 1, 1), 0.1, 1.0
            yield from conv3(name + ".residual.6", cout, cout)
            if cin != cout:
                yield from conv3(name + ".shortcut", cout, cin, (1, 1, 1))
        elif kind == "attn":
            yield prefix + name + ".norm.gamma", (cin, 1, 1), 0.1, 1.0
            yield from conv2(name + ".to_qkv", cin * 3, cin, 1)
            yield from conv2(name + ".proj", cin, cin, 1)      # zero-init in the reference (F6)
        elif kind in ("down2d", "down3d"):
            yield from conv2(name + ".resample.1", cin, cin)
            if kind == "down3d":
                yield from conv3(name + ".time_conv", cin, cin, (3, 1, 1))
        elif kind in ("up2d", "up3d"):
            yield from conv2(name + ".resample.1", cin // 2, cin)
            if kind == "up3d":
                yield from conv3(name + ".time_conv", cin * 2, cin, (3, 1, 1))
        elif kind == "head":
            yield prefix + name + ".0.gamma", (cin, 1, 1, 1), 0.1, 1.0
            yield from conv3(name + ".2", cout, cin)

    for rec in encoder_layers(cfg):
        yield from layer(*rec)
    yield from conv3("conv1", cfg.z_dim * 2, cfg.z_dim * 2, (1, 1, 1))
    yield from conv3("conv2", cfg.z_dim, cfg.z_dim, (1, 1, 1))
    for rec in decoder_layers(cfg):
        yield from layer(*rec)


def adaptor_param_specs(kind: str, ch: int = 128, in_channels: int = 3) -> Iterator[Tuple[str, tuple, float, float]]:
    """VAEEncoderadaptor (traj:125-196, 1 ResnetBlock under `down.0.block`) or VAEDecoderadaptor
    (traj:200-279, 2 ResnetBlocks under `up.0.block`)."""
    def conv2(name, cout, cin):
        yield name + ".weight", (cout, cin, 3, 3), (cin * 9) ** -0.5, 0.0
        yield name + ".bias", (cout,), 0.05, 0.0

    def norm(name, c):
        yield name + ".weight", (c,), 0.1, 1.0
        yield name + ".bias", (c,), 0.05, 0.0

    yield from conv2("conv_in", ch, in_channels)
    group, n = ("down", 1) if kind == "encoder" else ("up", 2)
    for b in range(n):
        p = f"{group}.0.block.{b}"
        yield from norm(p + ".norm1", ch)
        yield from conv2(p + ".conv1", ch, ch)
        yield from norm(p + ".norm2", ch)
        yield from conv2(p + ".conv2", ch, ch)
    yield from norm("norm_out", ch)
    yield from conv2("conv_out", in_channels, ch)            # zero-init in the encoder adaptor (F6)
