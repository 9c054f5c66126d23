"""Host logic of the VAE mirror without a GPU: the layer program + fusion plan.  `ops` is replaced
by shape-propagating recorders, so this checks WHICH kernels the host would launch and on what
shapes — in particular that every RMS_norm whose producer is a 96/192-channel 3x3 conv is fused
into that conv's epilogue (m4d_conv3x3_rmsnorm_cl) and that the rest still run as kernels."""
import collections

import pytest
import torch

from more4d_b200 import ops as real_ops, vae as vae_mod


class _Recorder:
    def __init__(self):
        self.calls = collections.Counter()
        self.fused = []          # (cin, cout, want_raw, has_residual)
        self.norm_channels = []
        self.stats_of = {}

    def t(self, *shape):
        return torch.empty(*shape, device="meta", dtype=torch.bfloat16)

    # ---- fakes with the real signatures (shapes only)
    def planar_to_cl(self, x, cpad, div=None, add=None):
        self.calls["planar_to_cl"] += 1
        C, T, H, W = x.shape
        return self.t(T, H, W, cpad)

    def cl_to_planar(self, x, n_affine=0, sub=None, mul=None):
        self.calls["cl_to_planar"] += 1
        T, H, W, C = x.shape
        return self.t(C, T, H, W)

    def pack_conv_weight(self, w, cin_multiple=32):
        if w.dim() == 4:
            w = w.unsqueeze(2)
        cout, cin, kt, kh, kw = w.shape
        cin_p, cout_p = -(-cin // cin_multiple) * cin_multiple, -(-cout // 16) * 16
        return self.t(cout_p, kt * kh * kw * cin_p)

    def conv_cl(self, x, w_packed, bias, cout, kernel, stride=(1, 1, 1), pad=(0, 0, 0), t_out=None, out=None,
                t_mul=1, t_off=0, n_split=None, residual=None, planar_out=None, act=0, skip=None):
        self.calls["conv_cl"] += 1
        T, H, W, Cin = x.shape
        assert w_packed.shape[1] == kernel[0] * kernel[1] * kernel[2] * Cin
        if planar_out is not None:
            return planar_out
        if out is not None:
            return out
        Ho = H // stride[1] if stride[1] > 1 else H
        Wo = W // stride[2] if stride[2] > 1 else W
        return self.t(T if t_out is None else t_out, Ho, Wo, cout)

    def conv3x3_rmsnorm_cl(self, x, w_packed, bias, cout, kt, gamma, silu=True, want_raw=True, residual=None):
        self.calls["conv3x3_rmsnorm_cl"] += 1
        assert cout in real_ops.FUSED_NORM_CHANNELS and gamma.shape[0] == cout and silu
        T, H, W, Cin = x.shape
        assert w_packed.shape == (cout, kt * 9 * Cin)
        self.fused.append((Cin, cout, want_raw, residual is not None))
        return (self.t(T, H, W, cout) if want_raw else None), self.t(T, H, W, cout)

    def conv3x3_gnstats_cl(self, x, w_packed, bias, cout, residual=None):
        self.calls["conv3x3_gnstats_cl"] += 1
        T, H, W, Cin = x.shape
        assert cout == 128 and w_packed.shape == (cout, 9 * Cin)
        out = self.t(T, H, W, cout)
        stats = torch.empty(T * real_ops.GN_SLICES * 64, device="meta")
        self.stats_of[id(stats)] = out
        return out, stats

    def groupnorm_swish_cl(self, x, weight, bias, eps=1e-6, groups=32, inplace=False, stats=None):
        self.calls["groupnorm_swish_cl"] += 1
        # the statistics handed over must be the ones of THIS tensor (left by its producing conv)
        assert stats is not None and self.stats_of[id(stats)] is x
        return x if inplace else self.t(*x.shape)

    def rmsnorm_silu_cl(self, x, gamma, silu=True, inplace=False):
        self.calls["rmsnorm_silu_cl"] += 1
        self.norm_channels.append(x.shape[-1])
        return x

    def upsample2x_cl(self, x):
        self.calls["upsample2x_cl"] += 1
        T, H, W, C = x.shape
        return self.t(T, 2 * H, 2 * W, C)

    def linear(self, x, weight, bias=None, epilogue=0, out=None, residual=None, **kw):
        self.calls["linear"] += 1
        return out if out is not None else self.t(*x.shape[:-1], weight.shape[0])

    def softmax_rows(self, s, scale, out=None):
        self.calls["softmax_rows"] += 1
        return out if out is not None else self.t(*s.shape)

    def transpose_bf16(self, m, out=None):
        self.calls["transpose_bf16"] += 1
        return out if out is not None else self.t(m.shape[1], m.shape[0])

    FUSED_NORM_CHANNELS = real_ops.FUSED_NORM_CHANNELS
    EPI_F32_RAW, EPI_ADD_BF16 = real_ops.EPI_F32_RAW, real_ops.EPI_ADD_BF16


@pytest.fixture
def rec(monkeypatch):
    r = _Recorder()
    monkeypatch.setattr(vae_mod, "ops", r)
    monkeypatch.setattr(vae_mod, "_no_grad_only", lambda what: None)
    return r


def _model():
    m = vae_mod.AutoencoderKLWan(device="meta")
    m._affine_consts = lambda: (torch.empty(16, device="meta"), torch.empty(16, device="meta"))
    return m


def test_encoder_plan_fuses_every_96_192_norm(rec):
    m = _model()
    out = m._encode_one(torch.empty(3, 9, 64, 96, device="meta", dtype=torch.bfloat16))
    assert tuple(out.shape) == (32, 3, 8, 12)                         # 1 + 8/4 latent frames, /8 spatially
    # 4 residual blocks at 96 / 192 channels: conv1 of each emits only the normalised tensor;
    # encoder.conv1 and the first block's conv2 also emit the next block's norm1 input
    assert rec.calls["conv3x3_rmsnorm_cl"] == len(rec.fused) >= 6
    assert (16, 96, True, False) in rec.fused                         # encoder.conv1 on the 16-channel padded video
    assert any(res and want_raw for _, _, want_raw, res in rec.fused)
    # what is left for the stand-alone kernel: 384-channel blocks, the attention norm, the head
    assert set(rec.norm_channels) <= {96, 192, 384}
    assert rec.norm_channels.count(96) <= 1 and rec.calls["rmsnorm_silu_cl"] < 20
    fused = rec.calls["conv3x3_rmsnorm_cl"]
    m.fuse_norms = False
    rec.calls.clear(); rec.fused.clear(); rec.norm_channels.clear()
    m._encode_one(torch.empty(3, 9, 64, 96, device="meta", dtype=torch.bfloat16))
    assert rec.calls["conv3x3_rmsnorm_cl"] == 0
    assert rec.calls["rmsnorm_silu_cl"] >= fused                       # every fused norm is a kernel again


def test_decoder_plan(rec):
    m = _model()
    video = m._decode_one(torch.empty(16, 3, 8, 12, device="meta", dtype=torch.bfloat16))
    assert tuple(video.shape) == (3, 9, 64, 96)
    # upsample-branch convs (384 -> 192, 192 -> 96) feed residual blocks: fused too; the head norm
    # is fused into the last block's conv2, so no 96-channel stand-alone norm remains
    assert any(cin == 384 and cout == 192 for cin, cout, _, _ in rec.fused)
    assert any(cin == 192 and cout == 96 for cin, cout, _, _ in rec.fused)
    assert 96 not in rec.norm_channels and 192 not in rec.norm_channels
    assert rec.calls["upsample2x_cl"] == 3


@pytest.mark.parametrize("cls,gn,convs", [(vae_mod.VAEEncoderadaptor, 3, 3), (vae_mod.VAEDecoderadaptor, 5, 5)])
def test_adaptor_plan_every_groupnorm_gets_its_statistics_from_the_producing_conv(rec, cls, gn, convs):
    """trajectory_module.py:104-122,125-279: conv_in and every ResnetBlock conv feed a Normalize; the
    host hands each GroupNorm the statistics its producer's epilogue left (no statistics pass)."""
    a = cls(device="meta")
    x = torch.empty(3, 5, 32, 48, device="meta", dtype=torch.bfloat16)
    out = a._forward_one(x)
    assert tuple(out.shape) == (3, 5, 32, 48)
    assert rec.calls["groupnorm_swish_cl"] == gn and rec.calls["conv3x3_gnstats_cl"] == convs
    assert rec.calls["conv_cl"] == 1                                   # conv_out (planar 3-channel output)
