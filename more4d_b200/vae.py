"""Host-side mirror of the Motion-Sensitive 3D-VAE: the causal Wan2.1 VAE
(MoRe4D/models/wan_vae.py, "vae") and the trajectory adaptors
(MoRe4D/models/trajectory_module.py, "traj") over the sm_100a kernels.

Call surface and state-dict keys follow the reference (SURVEY.md §8b seam 6/7):
``AutoencoderKLWan().encode(x)[0].mode()/.sample()``, ``.encode(x).latent_dist``,
``.decode(z).sample``, ``.model.*`` parameters under the same keys; ``VAEEncoderadaptor()(x)``,
``VAEDecoderadaptor()(x)`` on ``[B, 3, F, H, W]``.

B200-first execution: activations are channels-last bf16 ``[T, H, W, C]``; every
convolution is one TMA-fed tcgen05 implicit GEMM over the WHOLE sequence — the reference's
chunk loop, 2-frame feature caches, ``torch.cat``/``clone``/``F.pad`` copies and the O(T^2)
growing output concat (vae:524-538, 689-701) disappear, because the chunked computation is
algebraically a causal convolution with three first-frame rules (oracle/vae_oracle.py).
180 GB of HBM hold the full-resolution sequence (49x720x1280x96 bf16 = 8.7 GB per activation).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import ops
from .vae_arch import (VAEConfig, WAN_VAE, adaptor_param_specs, decoder_layers, encoder_layers,
                       vae_param_specs)

Tensor = torch.Tensor
BF16 = torch.bfloat16


def _no_grad_only(what: str) -> None:
    if torch.is_grad_enabled():
        raise RuntimeError(f"more4d_b200.{what}: forward-only kernels; wrap the call in "
                           "torch.no_grad() (training paths are out of scope)")


def _tree_from_specs(specs, device) -> nn.Module:
    """Nested parameter containers whose state_dict keys equal the dotted spec keys."""
    root = nn.Module()
    for key, shape, _std, _mean in specs:
        parts = key.split(".")
        node = root
        for name in parts[:-1]:
            if name not in node._modules:
                node.add_module(name, nn.Module())
            node = node._modules[name]
        node.register_parameter(parts[-1], nn.Parameter(torch.empty(*shape, device=device, dtype=BF16),
                                                        requires_grad=False))
    return root


class _PackedWeights:
    """Implicit-GEMM weight operands, packed lazily and re-packed when the parameter changes
    (load_state_dict / in-place updates bump ``_version``)."""

    def __init__(self):
        self._cache: Dict[str, tuple] = {}

    def get(self, key: str, w: Tensor, cin_multiple: int = 32) -> Tensor:
        tag = (w.data_ptr(), w._version, tuple(w.shape), cin_multiple)
        hit = self._cache.get(key)
        if hit is None or hit[0] != tag:
            hit = (tag, ops.pack_conv_weight(w.detach(), cin_multiple))
            self._cache[key] = hit
        return hit[1]


class DiagonalGaussianDistribution:
    """Semantics of diffusers' class as the reference uses it (vae:797): parameters =
    (mean | logvar) on dim 1, logvar clamped to [-30, 20]."""

    def __init__(self, parameters: Tensor):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def mode(self) -> Tensor:
        return self.mean

    def sample(self, generator=None) -> Tensor:
        eps = torch.randn(self.mean.shape, generator=generator, device=self.mean.device,
                          dtype=self.mean.dtype)
        return self.mean + self.std * eps


class AutoencoderKLOutput:
    def __init__(self, latent_dist):
        self.latent_dist = latent_dist

    def __getitem__(self, i):
        return (self.latent_dist,)[i]


class DecoderOutput:
    def __init__(self, sample):
        self.sample = sample


class _Cfg(dict):
    __getattr__ = dict.get


class AutoencoderKLWan(nn.Module):
    """vae:748-871 (wrapper) + vae:487-724 (AutoencoderKLWan_) — forward-only, B200-native."""

    def __init__(self, latent_channels=16, temporal_compression_ratio=4, spatial_compression_ratio=8,
                 device=None, cfg: VAEConfig = WAN_VAE):
        super().__init__()
        assert latent_channels == cfg.z_dim
        self.cfg = cfg
        self.config = _Cfg(latent_channels=latent_channels,
                           temporal_compression_ratio=temporal_compression_ratio,
                           spatial_compression_ratio=spatial_compression_ratio)
        self.latent_channels = latent_channels
        self.temporal_compression_ratio = temporal_compression_ratio
        self.spatial_compression_ratio = spatial_compression_ratio
        self.model = _tree_from_specs(vae_param_specs(cfg, prefix=""), device)
        self.mean = torch.tensor(cfg.mean, dtype=torch.float32)
        self.std = torch.tensor(cfg.std, dtype=torch.float32)
        self.scale = [self.mean, 1.0 / self.std]
        self._packed = _PackedWeights()
        self._consts = None
        self.fuse_norms = True      # RMS_norm+SiLU in the producing conv's epilogue (96/192 channels)

    @classmethod
    def from_pretrained(cls, pretrained_model_path, additional_kwargs=None, device=None):
        """vae:849-871: a `.safetensors` / torch checkpoint of the inner model; keys get the `model.`
        prefix; constructor arguments are filtered from `additional_kwargs`; non-strict load."""
        import inspect
        accepted = set(inspect.signature(cls.__init__).parameters) - {"self", "device", "cfg"}
        model = cls(**{k: v for k, v in (additional_kwargs or {}).items() if k in accepted}, device=device)
        if str(pretrained_model_path).endswith(".safetensors"):
            from safetensors.torch import load_file
            state = load_file(pretrained_model_path)
        else:
            state = torch.load(pretrained_model_path, map_location="cpu")
        missing, unexpected = model.load_state_dict({"model." + k: v for k, v in state.items()}, strict=False)
        model.load_report = {"missing": list(missing), "unexpected": list(unexpected)}
        return model

    @property
    def dtype(self):
        return self.model.conv1.weight.dtype

    @property
    def device(self):
        return self.model.conv1.weight.device

    # --------------------------------------------------------------------------- helpers
    def _p(self, key: str) -> Tensor:
        node = self.model
        for name in key.split("."):
            node = node._modules[name] if name in node._modules else node._parameters[name]
        return node

    def _affine_consts(self):
        """mean and 1/std as the reference sees them: cast to the latent dtype (vae:681)."""
        if self._consts is None or self._consts[0].device != self.device:
            mean = self.mean.to(BF16).float().to(self.device)
            inv_std = (1.0 / self.std).to(BF16).float().to(self.device)
            self._consts = (mean, inv_std)
        return self._consts

    def _input_cl(self, x: Tensor, in_scale: float, in_shift: float) -> Tensor:
        if in_scale == 1.0 and in_shift == 0.0:
            return ops.planar_to_cl(x, 16)
        dev = x.device
        div = torch.full((3,), 1.0 / in_scale, device=dev, dtype=torch.float32)
        add = torch.full((3,), in_shift, device=dev, dtype=torch.float32)
        return ops.planar_to_cl(x, 16, div, add)

    def _conv(self, x: Tensor, name: str, kernel, cout: int, pad, **kw) -> Tensor:
        w = self._p(name + ".weight")
        if tuple(kernel) == (1, 1, 1) and not (set(kw) - {"out"}) and w.shape[1] == x.shape[-1] \
                and x.is_contiguous() and x.shape[-1] % 8 == 0:
            # a 1x1x1 convolution over channels-last pixels IS a row-major GEMM [pixels, Cin] x [Cout, Cin]^T
            # (shortcuts vae:204-205, conv1 / conv2 vae:509-510): the 2-CTA GEMM with its double-buffered,
            # 256-bit epilogue instead of the per-tap conv kernel (output-bound at K = 96: 6.0 -> 2 ms)
            out = kw.get("out")
            T, H, W, _ = x.shape
            if out is None:
                out = torch.empty(T, H, W, cout, device=x.device, dtype=BF16)
            if out.is_contiguous() and tuple(out.shape) == (T, H, W, cout):
                ops.linear(x.view(T * H * W, -1), w.view(cout, -1), self._p(name + ".bias"),
                           out=out.view(T * H * W, cout))
                return out
        mult = 16 if x.shape[-1] % 32 else 32               # thin (3 -> 16 channel) inputs
        return ops.conv_cl(x, self._packed.get(name, w, mult), self._p(name + ".bias"), cout, kernel, pad=pad, **kw)

    def _conv_norm(self, x: Tensor, name: str, kt: int, cout: int, gamma: Tensor, want_raw: bool = True,
                   residual: Optional[Tensor] = None):
        """3x3 conv whose epilogue also emits [SiLU](RMS_norm(out) * gamma) — the consumer's first
        op (vae:198-202) — for the channel counts the fused epilogue supports."""
        w = self._p(name + ".weight")
        mult = 16 if x.shape[-1] % 32 else 32
        return ops.conv3x3_rmsnorm_cl(x, self._packed.get(name, w, mult), self._p(name + ".bias"), cout, kt, gamma,
                                      silu=True, want_raw=want_raw, residual=residual)

    def _res(self, x: Tensor, name: str, cin: int, cout: int, x_normed: Optional[Tensor] = None,
             next_gamma: Optional[Tensor] = None):
        """ResidualBlock (vae:190-224).  `x_normed` = SiLU(RMS_norm(x)) if the producer of x already
        emitted it; `next_gamma` asks this block to emit the same for its consumer.  Returns
        (out, out_normed or None)."""
        h = x if cin == cout else self._conv(x, name + ".shortcut", (1, 1, 1), cout, (0, 0, 0))
        y = x_normed if x_normed is not None else ops.rmsnorm_silu_cl(x, self._p(name + ".residual.0.gamma"))
        fuse = self.fuse_norms and cout in ops.FUSED_NORM_CHANNELS
        if fuse:
            _, c1 = self._conv_norm(y, name + ".residual.2", 3, cout, self._p(name + ".residual.3.gamma"),
                                    want_raw=False)
        else:
            c1 = self._conv(y, name + ".residual.2", (3, 3, 3), cout, (2, 1, 1))
            ops.rmsnorm_silu_cl(c1, self._p(name + ".residual.3.gamma"), inplace=True)
        del y
        if fuse and next_gamma is not None:
            return self._conv_norm(c1, name + ".residual.6", 3, cout, next_gamma, residual=h)
        return self._conv(c1, name + ".residual.6", (3, 3, 3), cout, (2, 1, 1), residual=h), None

    def _attn(self, x: Tensor, name: str) -> Tensor:
        T, H, W, C = x.shape
        L = H * W
        y = ops.rmsnorm_silu_cl(x, self._p(name + ".norm.gamma"), silu=False)
        wq = self._p(name + ".to_qkv.weight").view(3 * C, C)
        qkv = ops.linear(y.view(T * L, C), wq, self._p(name + ".to_qkv.bias"))          # [T*L, 3C]
        o = torch.empty(T * L, C, device=x.device, dtype=BF16)
        # the key dimension is the K of the second GEMM: padded to a multiple of 8 tokens (16-byte
        # rows for TMA) with zero probabilities / zero V columns when H/8 * W/8 is not one
        Lp = (L + 7) // 8 * 8
        s = torch.empty(L, Lp, device=x.device, dtype=torch.float32)
        p = torch.zeros(L, Lp, device=x.device, dtype=BF16) if Lp != L else torch.empty(L, L, device=x.device, dtype=BF16)
        vt = torch.zeros(C, Lp, device=x.device, dtype=BF16) if Lp != L else torch.empty(C, L, device=x.device, dtype=BF16)
        for f in range(T):
            blk = qkv[f * L:(f + 1) * L]
            ops.linear(blk[:, :C], blk[:, C:2 * C], None, ops.EPI_F32_RAW, out=s[:, :L])   # [L, L] fp32 logits
            ops.softmax_rows(s[:, :L], 1.0 / math.sqrt(C), out=p[:, :L])
            ops.transpose_bf16(blk[:, 2 * C:], out=vt[:, :L])                               # [C, L]
            ops.linear(p, vt, None, out=o[f * L:(f + 1) * L])
        wp = self._p(name + ".proj.weight").view(C, C)
        out = ops.linear(o, wp, self._p(name + ".proj.bias"), ops.EPI_ADD_BF16, residual=x.view(T * L, C))
        return out.view(T, H, W, C)

    def _down(self, x: Tensor, name: str, kind: str, c: int) -> Tensor:
        T = x.shape[0]
        y = self._conv(x, name + ".resample.1", (1, 3, 3), c, (0, 0, 0), stride=(1, 2, 2))
        if kind == "down3d" and T > 1:
            t_rest = (T - 1) // 2
            out = torch.empty(1 + t_rest, *y.shape[1:], device=x.device, dtype=BF16)
            out[0].copy_(y[0])                                                            # first-chunk rule
            self._conv(y, name + ".time_conv", (3, 1, 1), c, (0, 0, 0), stride=(2, 1, 1), t_out=t_rest,
                       out=out, t_off=1)
            y = out
        return y

    def _up(self, x: Tensor, name: str, kind: str, c: int, next_gamma: Optional[Tensor] = None):
        T, H, W, _ = x.shape
        if kind == "up3d" and T > 1:
            u = torch.empty(1 + 2 * (T - 1), H, W, c, device=x.device, dtype=BF16)
            u[0].copy_(x[0])                                                              # 'Rep' rule
            self._conv(x[1:], name + ".time_conv", (3, 1, 1), 2 * c, (2, 0, 0), t_out=T - 1, out=u,
                       t_mul=2, t_off=1, n_split=c)
            x = u
        up = ops.upsample2x_cl(x)
        if next_gamma is not None and self.fuse_norms and c // 2 in ops.FUSED_NORM_CHANNELS:
            return self._conv_norm(up, name + ".resample.1", 1, c // 2, next_gamma)
        return self._conv(up, name + ".resample.1", (1, 3, 3), c // 2, (0, 1, 1)), None

    def _next_gamma(self, layers, i: int, tail_gamma: Optional[Tensor]) -> Optional[Tensor]:
        """gamma of the RMS_norm that consumes layer i's output first, if that is how it is
        consumed (a ResidualBlock's residual.0, or the head norm after the last layer)."""
        if i + 1 < len(layers):
            kind, name = layers[i + 1][0], layers[i + 1][1]
            return self._p(name + ".residual.0.gamma") if kind == "res" else None
        return tail_gamma

    def _run(self, x: Tensor, layers, x_normed: Optional[Tensor] = None,
             tail_gamma: Optional[Tensor] = None):
        """Runs the layer program; returns (x, SiLU(RMS_norm(x) * tail_gamma) or None)."""
        for i, (kind, name, cin, cout) in enumerate(layers):
            ng = self._next_gamma(layers, i, tail_gamma)
            xn = None
            if kind == "conv":
                x = self._conv(x, name, (3, 3, 3), cout, (2, 1, 1))
            elif kind == "res":
                x, xn = self._res(x, name, cin, cout, x_normed, ng)
            elif kind == "attn":
                x = self._attn(x, name)
            elif kind in ("down2d", "down3d"):
                x = self._down(x, name, kind, cin)
            elif kind in ("up2d", "up3d"):
                x, xn = self._up(x, name, kind, cin, ng)
            x_normed = xn
        return x, x_normed

    # --------------------------------------------------------------------------- encode / decode
    def _encode_one(self, x: Tensor, in_scale: float = 1.0, in_shift: float = 0.0) -> Tensor:
        """x [3, F, H, W] (F = 1 + 4k) -> params [2z, 1 + k, H/8, W/8]  (vae:520-547)."""
        cfg = self.cfg
        layers = encoder_layers(cfg)
        # 3-channel input: zero-padded to 32 channels so the first conv also runs on the tensor
        # cores (K = 27 x 32); the optional `x*2-1` is fused into the layout pass
        xc = self._input_cl(x, in_scale, in_shift)
        body = layers[1:-1]
        _, hname, cin, cout = layers[-1]
        c0, hn = layers[0][3], None
        g0 = self._next_gamma(body, -1, None)
        if g0 is not None and self.fuse_norms and c0 in ops.FUSED_NORM_CHANNELS:
            h, hn = self._conv_norm(xc, "encoder.conv1", 3, c0, g0)
        else:
            h = self._conv(xc, "encoder.conv1", (3, 3, 3), c0, (2, 1, 1))
        del xc
        h, hn = self._run(h, body, hn, self._p(hname + ".0.gamma"))
        if hn is None:
            hn = ops.rmsnorm_silu_cl(h, self._p(hname + ".0.gamma"), inplace=True)
        del h
        h = self._conv(hn, hname + ".2", (3, 3, 3), cout, (2, 1, 1))
        h = self._conv(h, "conv1", (1, 1, 1), cout, (0, 0, 0))
        mean, inv_std = self._affine_consts()
        return ops.cl_to_planar(h, cfg.z_dim, mean, inv_std)

    def _decode_one(self, z: Tensor) -> Tensor:
        """z [z, T, h, w] -> video [3, 4T-3, 8h, 8w] clamped to [-1, 1]  (vae:678-703, 827)."""
        cfg = self.cfg
        layers = decoder_layers(cfg)
        mean, inv_std = self._affine_consts()
        zc = ops.planar_to_cl(z, 32, inv_std, mean)
        T, h, w, _ = zc.shape
        z2 = torch.zeros(T, h, w, 32, device=z.device, dtype=BF16)       # channels 16..31 stay zero
        self._conv(zc, "conv2", (1, 1, 1), cfg.z_dim, (0, 0, 0), out=z2)
        _, hname, cin, cout = layers[-1]
        x, xn = self._run(z2, layers[:-1], None, self._p(hname + ".0.gamma"))
        if xn is None:
            xn = ops.rmsnorm_silu_cl(x, self._p(hname + ".0.gamma"), inplace=True)
        del x
        x = xn
        To, H, W, _ = x.shape
        video = torch.empty(cout, To, H, W, device=z.device, dtype=BF16)
        self._conv(x, hname + ".2", (3, 3, 3), cout, (2, 1, 1), planar_out=video, act=1)
        return video

    def _encode(self, x: Tensor) -> Tensor:
        return torch.stack([self._encode_one(u) for u in x])

    def encode(self, x: Tensor, return_dict: bool = True):
        _no_grad_only("AutoencoderKLWan.encode")
        dist = DiagonalGaussianDistribution(self._encode(x.to(device=self.device, dtype=BF16)))
        return AutoencoderKLOutput(dist) if return_dict else (dist,)

    def encode_scaled(self, x: Tensor, in_scale: float, in_shift: float):
        """encode(x * in_scale + in_shift) with the affine fused into the first conv
        (the `pseudo*2-1` of scripts/inference/infer_vae.py:278)."""
        _no_grad_only("AutoencoderKLWan.encode")
        x = x.to(device=self.device, dtype=BF16)
        return AutoencoderKLOutput(DiagonalGaussianDistribution(
            torch.stack([self._encode_one(u, in_scale, in_shift) for u in x])))

    def decode(self, z: Tensor, return_dict: bool = True):
        _no_grad_only("AutoencoderKLWan.decode")
        z = z.to(device=self.device, dtype=BF16)
        out = torch.stack([self._decode_one(u) for u in z])
        return DecoderOutput(out) if return_dict else (out,)

    def _decode(self, zs: Tensor) -> DecoderOutput:
        """vae:822-830 (already clamped to [-1, 1])."""
        return self.decode(zs)

    # The reference's "memory saver" entry points (vae:783-789,803-821,842-847 -> encode_full /
    # decode_full, vae:549-676) run the same chunk loop under gradient checkpointing and are
    # "functionally identical" forward (vae:551-552); here every path is the whole-sequence kernel
    # formulation, which holds no per-chunk cache at all.
    _encode_memory_saver = _encode
    encode_memory_saver = encode
    _decode_memory_saver = _decode
    decode_memory_saver = decode


class _Adaptor(nn.Module):
    kind = "encoder"

    def __init__(self, device=None, **ignored):
        super().__init__()
        tree = _tree_from_specs(adaptor_param_specs(self.kind), device)
        for name, mod in tree._modules.items():
            self.add_module(name, mod)
        self.in_channels, self.ch = 3, 128
        self._packed = _PackedWeights()

    def _p(self, key: str) -> Tensor:
        node = self
        for name in key.split("."):
            node = node._modules[name] if name in node._modules else node._parameters[name]
        return node

    def _conv(self, x, name, cout, **kw):
        w = self._p(name + ".weight")
        mult = 16 if x.shape[-1] % 32 else 32
        return ops.conv_cl(x, self._packed.get(name, w, mult), self._p(name + ".bias"), cout, (1, 3, 3),
                           pad=(0, 1, 1), **kw)

    def _conv_stats(self, x, name, residual=None):
        """A 128-channel conv whose output feeds a Normalize: the GroupNorm statistics come out of
        the conv's epilogue (ops.conv3x3_gnstats_cl).  Returns (out, stats)."""
        w = self._p(name + ".weight")
        mult = 16 if x.shape[-1] % 32 else 32
        return ops.conv3x3_gnstats_cl(x, self._packed.get(name, w, mult), self._p(name + ".bias"), self.ch,
                                      residual=residual)

    def _resnet(self, x: Tensor, x_stats: Tensor, p: str):
        """ResnetBlock.forward, temb = None, in == out channels (traj:104-122).  Takes and returns
        (tensor, GroupNorm statistics of that tensor)."""
        g = ops.groupnorm_swish_cl(x, self._p(p + ".norm1.weight"), self._p(p + ".norm1.bias"), stats=x_stats)
        c1, s1 = self._conv_stats(g, p + ".conv1")
        del g
        ops.groupnorm_swish_cl(c1, self._p(p + ".norm2.weight"), self._p(p + ".norm2.bias"), inplace=True, stats=s1)
        return self._conv_stats(c1, p + ".conv2", residual=x)

    def _forward_one(self, x: Tensor) -> Tensor:
        raise NotImplementedError

    def forward(self, x: Tensor) -> Tensor:
        _no_grad_only(type(self).__name__)
        dev = self._p("conv_in.weight").device
        x = x.to(device=dev, dtype=BF16)
        return torch.stack([self._forward_one(u.contiguous()) for u in x])


class VAEEncoderadaptor(_Adaptor):
    """traj:125-196: xyz trajectories -> pseudo-RGB in (0, 1)."""
    kind = "encoder"

    def _forward_one(self, x: Tensor) -> Tensor:          # x [3, F, H, W]
        h, st = self._conv_stats(ops.planar_to_cl(x, 16), "conv_in")
        h, st = self._resnet(h, st, "down.0.block.0")
        ops.groupnorm_swish_cl(h, self._p("norm_out.weight"), self._p("norm_out.bias"), inplace=True, stats=st)
        out = torch.empty_like(x)
        self._conv(h, "conv_out", 3, planar_out=out, act=2, skip=x)      # sigmoid(h + x), traj:194
        return out


class VAEDecoderadaptor(_Adaptor):
    """traj:200-279: reconstructed pseudo-RGB -> xyz trajectories."""
    kind = "decoder"

    def _forward_one(self, z: Tensor) -> Tensor:
        h, st = self._conv_stats(ops.planar_to_cl(z, 16), "conv_in")
        h, st = self._resnet(h, st, "up.0.block.0")
        h, st = self._resnet(h, st, "up.0.block.1")
        ops.groupnorm_swish_cl(h, self._p("norm_out.weight"), self._p("norm_out.bias"), inplace=True, stats=st)
        out = torch.empty_like(z)
        self._conv(h, "conv_out", 3, planar_out=out, act=0)
        return out


@torch.no_grad()
def motion_vae_roundtrip(x: Tensor, vae: AutoencoderKLWan, enc: VAEEncoderadaptor,
                         dec: VAEDecoderadaptor):
    """scripts/inference/infer_vae.py:276-281 with .mode(): adaptor -> *2-1 -> encode -> decode
    -> adaptor.  Returns (reconstructed trajectories, latent, reconstructed video)."""
    pseudo = enc(x)
    latent = vae.encode_scaled(pseudo, 2.0, -1.0).latent_dist.mode()
    recon = vae.decode(latent).sample
    return dec(recon), latent, recon
