"""GPU parity of the Motion-Sensitive VAE path (through the C ABI) against the CPU oracle and the
golden vectors produced by the real reference VAE (chunked, with its feature cache).

Tolerances (relative Frobenius error):
  * single kernels vs the oracle's bf16-emulation: <= 3e-3 (bf16 output rounding ~1e-3 plus
    summation-order disagreements of 1 bf16 ulp);
  * whole encoder / decoder / adaptors: ~30-60 bf16 convolutions deep, every layer's output is
    re-rounded to bf16, so 1-ulp disagreements compound: <= 3e-2 vs the bf16-emulating oracle
    and vs the fp32 reference golden.  (The reference's own bf16 CUDA path differs from its fp32
    path by the same order: the oracle's emulation mode measures that on CPU.)
"""
import pytest
import torch

from more4d_b200 import synth
from oracle import vae_oracle as V
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16
SEED = 11


def _rand(shape, seed, scale=1.0, dtype=BF16):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype)


def _cl(x):            # [1, C, T, H, W] -> channels-last [T, H, W, C] on the GPU
    return x[0].permute(1, 2, 3, 0).contiguous().cuda()


def _ncthw(y):         # channels-last GPU -> [1, C, T, H, W] fp32 CPU
    return y.permute(3, 0, 1, 2).unsqueeze(0).float().cpu()


@pytest.fixture(scope="module")
def ops():
    from more4d_b200 import ops as _ops
    return _ops


AR = V.Arith(True)


@pytest.mark.parametrize("cin,cout,T,H,W", [
    (96, 96, 3, 12, 20),      # ksub = 3, NT = 96, ragged H/W tiles
    (32, 192, 2, 8, 16),      # exactly one tile, NT = 192
    (192, 384, 2, 9, 17),     # ksub = 2, two N tiles of 192
    (384, 32, 1, 16, 32),     # encoder head shape (Cout 32)
    (16, 96, 3, 20, 24),      # thin input, 4 taps per weight box (3-channel video padded to 16)
    (32, 128, 2, 17, 40),     # thin input, 2 taps per weight box
    (48, 64, 2, 16, 16),      # one 64-channel block with 3 k-slices
    (80, 64, 2, 16, 16),      # 64 + 16 channels
])
def test_conv3d_causal(ops, cin, cout, T, H, W):
    x = _rand((1, cin, T, H, W), 1)
    w = _rand((cout, cin, 3, 3, 3), 2, (cin * 27) ** -0.5)
    b = _rand((cout,), 3, 0.1)
    ref = V.causal_conv3d(x.float(), w, b, AR)
    y = ops.conv_cl(_cl(x), ops.pack_conv_weight(w.cuda(), 16 if cin % 32 else 32), b.cuda(), cout, (3, 3, 3),
                    pad=(2, 1, 1))
    assert rel_err(_ncthw(y), ref) < 3e-3
    from more4d_b200 import _lib
    if cin % 32 == 0 and _lib.is_dev_build():   # development builds: the per-tap kernel must agree
        _lib.dev_set_flags(0x10000)
        try:
            y2 = ops.conv_cl(_cl(x), ops.pack_conv_weight(w.cuda()), b.cuda(), cout, (3, 3, 3), pad=(2, 1, 1))
        finally:
            _lib.dev_set_flags(0)
        assert rel_err(y2.float(), y.float()) < 2e-3


def test_conv_residual_and_pointwise(ops):
    cin, cout, T, H, W = 96, 192, 2, 10, 18
    x = _rand((1, cin, T, H, W), 4)
    w1 = _rand((cout, cin, 1, 1, 1), 5, cin ** -0.5)
    b1 = _rand((cout,), 6, 0.1)
    h = V.causal_conv3d(x.float(), w1, b1, AR)
    hg = ops.conv_cl(_cl(x), ops.pack_conv_weight(w1.cuda()), b1.cuda(), cout, (1, 1, 1))
    assert rel_err(_ncthw(hg), h) < 3e-3
    w = _rand((cout, cin, 3, 3, 3), 7, (cin * 27) ** -0.5)
    ref = AR.r(V.causal_conv3d(x.float(), w, b1, AR) + _ncthw(hg))
    y = ops.conv_cl(_cl(x), ops.pack_conv_weight(w.cuda()), b1.cuda(), cout, (3, 3, 3), pad=(2, 1, 1),
                    residual=hg)
    assert rel_err(_ncthw(y), ref) < 3e-3


@pytest.mark.parametrize("cin,cout,T,H,W,kt,res", [
    (96, 96, 3, 20, 37, 3, True),       # ragged tiles, residual, two 64-channel blocks (64 + 32)
    (192, 192, 2, 16, 16, 3, False),    # exactly one tile, single accumulator set
    (384, 192, 2, 9, 33, 1, False),     # upsample-branch Conv2d shape (kt = 1)
    (32, 96, 2, 18, 18, 3, False),      # half-filled channel block (encoder.conv1)
])
def test_conv3x3_fused_rmsnorm(ops, cin, cout, T, H, W, kt, res):
    """m4d_conv3x3_rmsnorm_cl == m4d_conv_cl followed by m4d_rmsnorm_silu_cl, and == the oracle."""
    x = _rand((1, cin, T, H, W), 31)
    w = _rand((cout, cin, kt, 3, 3), 32, (cin * 9 * kt) ** -0.5)
    b = _rand((cout,), 33, 0.1)
    g = _rand((cout, 1, 1, 1), 34, 0.1) + 1
    r = _rand((1, cout, T, H, W), 35) if res else None
    wp = ops.pack_conv_weight(w.cuda())
    rg = _cl(r) if res else None
    raw, normed = ops.conv3x3_rmsnorm_cl(_cl(x), wp, b.cuda(), cout, kt, g.cuda().reshape(-1), residual=rg)
    sep = ops.conv_cl(_cl(x), wp, b.cuda(), cout, (kt, 3, 3), pad=(kt - 1, 1, 1), residual=rg)
    assert torch.equal(raw, sep)                                   # same kernel, same arithmetic
    sep_n = ops.rmsnorm_silu_cl(sep, g.cuda().reshape(-1))
    assert rel_err(normed.float(), sep_n.float()) < 1e-3           # sum-of-squares order differs
    none_raw, normed2 = ops.conv3x3_rmsnorm_cl(_cl(x), wp, b.cuda(), cout, kt, g.cuda().reshape(-1),
                                               want_raw=False, residual=rg)
    assert none_raw is None and torch.equal(normed2, normed)
    if kt == 3:
        y = V.causal_conv3d(x.float(), w, b, AR)
    else:
        y = V.conv2d_frames(x.float(), w[:, :, 0], b, AR, stride=1, pad=(1, 1, 1, 1))
    if res:
        y = AR.r(y + r.float())
    ref = V.silu(V.rms_norm(y, g, AR), AR)
    assert rel_err(_ncthw(normed), ref) < 4e-3


def test_vae_fused_norms_match_unfused():
    """The fused-epilogue path and the separate-kernel path of the VAE agree."""
    x = synth._randn(SEED, "vae.x", (1, 3, 5, 32, 48), 0.5, "cpu", BF16)
    z = synth._randn(SEED, "vae.z", (1, 16, 2, 4, 6), 1.0, "cpu", BF16)
    m = _vae()
    outs = {}
    with torch.no_grad():
        for fuse in (True, False):
            m.fuse_norms = fuse
            outs[fuse] = (m.encode(x.cuda())[0].parameters.float().cpu(), m.decode(z.cuda()).sample.float().cpu())
    assert rel_err(outs[True][0], outs[False][0]) < 1e-2
    assert rel_err(outs[True][1], outs[False][1]) < 1e-2


def test_conv2d_stride2_downsample(ops):
    c, T, H, W = 96, 2, 12, 20
    x = _rand((1, c, T, H, W), 8)
    w = _rand((c, c, 3, 3), 9, (c * 9) ** -0.5)
    b = _rand((c,), 10, 0.1)
    ref = V.conv2d_frames(x.float(), w, b, AR, stride=2, pad=(0, 1, 0, 1))
    y = ops.conv_cl(_cl(x), ops.pack_conv_weight(w.cuda()), b.cuda(), c, (1, 3, 3), stride=(1, 2, 2))
    assert tuple(y.shape) == (T, 6, 10, c)
    assert rel_err(_ncthw(y), ref) < 3e-3


def test_time_conv_down_and_up(ops):
    c, T, H, W = 64, 5, 8, 16
    x = _rand((1, c, T, H, W), 11)
    w = _rand((c, c, 3, 1, 1), 12, (c * 3) ** -0.5)
    b = _rand((c,), 13, 0.1)
    sd = {"p.resample.1.weight": torch.zeros(c, c, 3, 3), "p.resample.1.bias": torch.zeros(c),
          "p.time_conv.weight": w, "p.time_conv.bias": b}
    # downsample3d temporal part: frame 0 passes through, then stride-2 windows
    rest = V.causal_conv3d(x.float(), w, b, AR, stride_t=2)
    out = torch.empty(3, H, W, c, device="cuda", dtype=BF16)
    xg = _cl(x)
    out[0].copy_(xg[0])
    ops.conv_cl(xg, ops.pack_conv_weight(w.cuda()), b.cuda(), c, (3, 1, 1), stride=(2, 1, 1), t_out=2,
                out=out, t_off=1)
    assert rel_err(_ncthw(out)[:, :, 1:], rest) < 3e-3
    assert torch.equal(out[0], xg[0])
    # upsample3d temporal part: frame 0 invisible, 2C channels interleaved into frames
    w2 = _rand((2 * c, c, 3, 1, 1), 14, (c * 3) ** -0.5)
    b2 = _rand((2 * c,), 15, 0.1)
    u = V.causal_conv3d(x[:, :, 1:].float(), w2, b2, AR).reshape(1, 2, c, T - 1, H, W)
    ref = torch.stack((u[:, 0], u[:, 1]), 3).reshape(1, c, 2 * (T - 1), H, W)
    ug = torch.empty(1 + 2 * (T - 1), H, W, c, device="cuda", dtype=BF16)
    ug[0].copy_(xg[0])
    ops.conv_cl(xg[1:], ops.pack_conv_weight(w2.cuda()), b2.cuda(), 2 * c, (3, 1, 1), pad=(2, 0, 0),
                t_out=T - 1, out=ug, t_mul=2, t_off=1, n_split=c)
    assert rel_err(_ncthw(ug)[:, :, 1:], ref) < 3e-3


def test_thin_input_conv_and_planar_out(ops):
    """Encoder3d.conv1 (vae:289) the way the host mirror runs it: the planar 3-channel video is
    laid out channels-last, zero-padded to 16 channels with `x*2-1` fused (m4d_planar_to_cl), and
    takes the halo kernel's thin-input mode."""
    T, H, W = 3, 10, 150
    x = _rand((1, 3, T, H, W), 16, 0.5)
    w = _rand((96, 3, 3, 3, 3), 17, 81 ** -0.5)
    b = _rand((96,), 18, 0.1)
    ref = V.causal_conv3d(AR.r(AR.r(x.float() * 2) - 1), w, b, AR)
    div = torch.full((3,), 0.5, device="cuda", dtype=torch.float32)
    add = torch.full((3,), -1.0, device="cuda", dtype=torch.float32)
    xcl = ops.planar_to_cl(x[0].cuda(), 16, div, add)
    y = ops.conv_cl(xcl, ops.pack_conv_weight(w.cuda(), 16), b.cuda(), 96, (3, 3, 3), pad=(2, 1, 1))
    assert rel_err(_ncthw(y), ref) < 3e-3
    # 96 -> 3 head with planar NCTHW output and clamp
    wh = _rand((3, 96, 3, 3, 3), 19, 0.05)
    bh = _rand((3,), 20, 0.1)
    refh = V.causal_conv3d(ref, wh, bh, AR).clamp(-1, 1)
    vid = torch.empty(3, T, H, W, device="cuda", dtype=BF16)
    ops.conv_cl(y, ops.pack_conv_weight(wh.cuda()), bh.cuda(), 3, (3, 3, 3), pad=(2, 1, 1), planar_out=vid, act=1)
    assert rel_err(vid.float().cpu().unsqueeze(0), refh) < 3e-3


def test_row_kernels(ops):
    T, H, W, C = 2, 5, 7, 192
    x = _rand((1, C, T, H, W), 21, 2.0)
    g = _rand((C, 1, 1, 1), 22, 0.1) + 1
    ref = V.silu(V.rms_norm(x.float(), g, AR), AR)
    y = ops.rmsnorm_silu_cl(_cl(x), g.cuda())
    assert rel_err(_ncthw(y), ref) < 3e-3
    up = ops.upsample2x_cl(_cl(x))
    refu = torch.nn.functional.interpolate(x[0].permute(1, 0, 2, 3).float(), scale_factor=2.0, mode="nearest-exact")
    assert torch.equal(up.float().cpu().permute(0, 3, 1, 2), refu)
    s = _rand((37, 24), 23, 3.0, torch.float32)
    p = ops.softmax_rows(s.cuda(), 0.25)
    assert rel_err(p.float().cpu(), torch.softmax(s * 0.25, -1)) < 3e-3
    for n, tail in ((2052, 0), (14400, 0), (330, 8), (27, 0)):     # register-resident / 3-pass kernel, strided rows
        s = _rand((9, n + tail), 28, 3.0, torch.float32)
        s[3, 5] = 40.0                                              # one dominant logit
        p = ops.softmax_rows(s.cuda()[:, :n], 0.125)
        ref = torch.softmax(s[:, :n] * 0.125, -1)
        assert rel_err(p.float().cpu(), ref) < 3e-3
        assert (p.float().sum(-1).cpu() - 1).abs().max() < 1e-2
    m = _rand((45, 70), 24)
    assert torch.equal(ops.transpose_bf16(m.cuda()).cpu(), m.t().contiguous())
    xa = _rand((1, 128, 3, 6, 10), 25, 2.0)
    w, b = _rand((128,), 26, 0.1) + 1, _rand((128,), 27, 0.1)
    refg = V._group_norm_swish(xa[0].permute(1, 0, 2, 3).float(), w, b, AR)            # [F, C, H, W]
    yg = ops.groupnorm_swish_cl(_cl(xa), w.cuda(), b.cuda())
    assert rel_err(yg.float().cpu().permute(0, 3, 1, 2), refg) < 4e-3


@pytest.mark.parametrize("T,H,W,cin,res", [(3, 33, 47, 128, True), (2, 16, 32, 16, False), (1, 50, 20, 128, False)])
def test_conv_with_fused_groupnorm_statistics(ops, T, H, W, cin, res):
    """m4d_conv3x3_gnstats_cl: same conv output as m4d_conv_cl, and the statistics its epilogue
    leaves reproduce groupnorm_swish_cl's own statistics pass (trajectory_module.py:54-60);
    odd tile counts exercise the CTA pair's idle half, cin 16 the thin-input kernel."""
    x = _rand((1, cin, T, H, W), 71, 1.5)
    w = _rand((128, cin, 3, 3), 72, 0.05)
    b = _rand((128,), 73, 0.2)
    r = _rand((1, 128, T, H, W), 74, 1.0) if res else None
    gw, gb = _rand((128,), 75, 0.1) + 1, _rand((128,), 76, 0.1)
    xc = _cl(x)
    wp = ops.pack_conv_weight(w.cuda(), 16 if cin % 32 else 32)
    rc = _cl(r) if res else None
    ref = ops.conv_cl(xc, wp, b.cuda(), 128, (1, 3, 3), pad=(0, 1, 1), residual=rc)
    out, stats = ops.conv3x3_gnstats_cl(xc, wp, b.cuda(), 128, residual=rc)
    assert torch.equal(out, ref)
    g_ref = ops.groupnorm_swish_cl(ref, gw.cuda(), gb.cuda())
    g = ops.groupnorm_swish_cl(out, gw.cuda(), gb.cuda(), stats=stats)
    assert rel_err(g.float().cpu(), g_ref.float().cpu()) < 2e-3
    # the statistics themselves: per frame and group, sum and sum of squares of the bf16 output
    st = stats.view(T, ops.GN_SLICES, 32, 2).sum(1).cpu()
    o = out.float().cpu().view(T, H * W, 32, 4)
    assert torch.allclose(st[..., 0], o.sum((1, 3)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(st[..., 1], (o * o).sum((1, 3)), rtol=1e-4, atol=1e-2)
    out2, stats2 = ops.conv3x3_gnstats_cl(xc, wp, b.cuda(), 128, residual=rc)
    assert torch.equal(stats, stats2)                                  # no atomics: reproducible


def _vae():
    from more4d_b200.vae import AutoencoderKLWan
    m = AutoencoderKLWan(device="cuda")
    m.load_state_dict(synth.vae_state_dict(seed=SEED), strict=True)
    return m


def test_vae_encode_decode(golden):
    g = golden("vae")
    sd = synth.vae_state_dict(seed=SEED)
    x = synth._randn(SEED, "vae.x", (1, 3, 13, 32, 48), 0.5, "cpu", BF16)
    z = synth._randn(SEED, "vae.z", (1, 16, 4, 4, 6), 1.0, "cpu", BF16)
    m = _vae()
    with torch.no_grad():
        dist = m.encode(x.cuda())[0]
        params = dist.parameters.float().cpu()
        rec = m.decode(z.cuda()).sample.float().cpu()
        p1 = m.encode(x[:, :, :1].cuda()).latent_dist.parameters.float().cpu()
        r1 = m.decode(z[:, :, :1].cuda()).sample.float().cpu()
    torch.cuda.synchronize()
    assert params.shape == (1, 32, 4, 4, 6) and rec.shape == (1, 3, 13, 32, 48)
    assert torch.equal(dist.mode().float().cpu(), params[:, :16])
    pe, re_ = V.encode(x.float(), sd, emulate_bf16=True), V.decode(z.float(), sd, emulate_bf16=True)
    errs = dict(enc_vs_emul=rel_err(params, pe), enc_vs_ref=rel_err(params, g["enc_params"]),
                dec_vs_emul=rel_err(rec, re_), dec_vs_ref=rel_err(rec, g["dec"]),
                enc1_vs_ref=rel_err(p1, g["enc_params_T1"]), dec1_vs_ref=rel_err(r1, g["dec_T1"]),
                oracle_emul_vs_ref_enc=rel_err(pe, g["enc_params"]), oracle_emul_vs_ref_dec=rel_err(re_, g["dec"]))
    print({k: f"{v:.3e}" for k, v in errs.items()})
    for k in ("enc_vs_emul", "enc_vs_ref", "dec_vs_emul", "dec_vs_ref", "enc1_vs_ref", "dec1_vs_ref"):
        assert errs[k] < 3e-2, (k, errs)


def test_adaptors_and_roundtrip(golden):
    from more4d_b200.vae import VAEDecoderadaptor, VAEEncoderadaptor, motion_vae_roundtrip
    g = golden("vae")
    tv = synth.trajectory_video(5, 32, 48, SEED)
    esd, dsd = synth.adaptor_state_dict("encoder", SEED), synth.adaptor_state_dict("decoder", SEED)
    ea, da = VAEEncoderadaptor(device="cuda"), VAEDecoderadaptor(device="cuda")
    ea.load_state_dict(esd, strict=True)
    da.load_state_dict(dsd, strict=True)
    with torch.no_grad():
        ye = ea(tv.cuda()).float().cpu()
        yd = da(tv.cuda()).float().cpu()
    errs = dict(enc_vs_ref=rel_err(ye, g["adaptor_enc"]), dec_vs_ref=rel_err(yd, g["adaptor_dec"]),
                enc_vs_emul=rel_err(ye, V.encoder_adaptor(tv.float(), esd, True)),
                dec_vs_emul=rel_err(yd, V.decoder_adaptor(tv.float(), dsd, True)))
    print({k: f"{v:.3e}" for k, v in errs.items()})
    assert errs["enc_vs_ref"] < 5e-3 and errs["enc_vs_emul"] < 5e-3
    assert errs["dec_vs_ref"] < 3e-2 and errs["dec_vs_emul"] < 3e-2
    # whole Motion-Sensitive round trip (infer_vae.py:276-281 with .mode())
    vsd = synth.vae_state_dict(seed=SEED)
    with torch.no_grad():
        out, lat, rec = motion_vae_roundtrip(tv.cuda(), _vae(), ea, da)
    ro, rl, rr = V.roundtrip(tv.float(), vsd, esd, dsd, emulate_bf16=True)
    assert out.shape == tv.shape and lat.shape == (1, 16, 2, 4, 6)
    e = dict(latent=rel_err(lat.float().cpu(), rl), video=rel_err(rec.float().cpu(), rr),
             out=rel_err(out.float().cpu(), ro))
    print({k: f"{v:.3e}" for k, v in e.items()})
    assert e["latent"] < 3e-2 and e["video"] < 5e-2 and e["out"] < 8e-2


# --------------------------------------------------------------------------------------------
# per-layer, teacher-forced parity (VERDICT r1 next #1c): every layer of the encoder and decoder
# programs receives the ORACLE's input of that layer and must reproduce the oracle's output of that
# layer to kernel-level accuracy — compounding over depth is taken out of the picture.
# --------------------------------------------------------------------------------------------
LAYER_TOL = 3e-3


def _oracle_layer(kind, x, sd, p):
    if kind == "conv":
        return V.causal_conv3d(x, sd[p + ".weight"], sd[p + ".bias"], AR)
    if kind == "res":
        return V.residual_block(x, sd, p, AR)
    if kind == "attn":
        return V.attention_block(x, sd, p, AR)
    if kind in ("down2d", "down3d"):
        return V.downsample(x, sd, p, kind, AR)
    if kind in ("up2d", "up3d"):
        return V.upsample(x, sd, p, kind, AR)
    if kind == "head":
        y = V.silu(V.rms_norm(x, sd[p + ".0.gamma"], AR), AR)
        return V.causal_conv3d(y, sd[p + ".2.weight"], sd[p + ".2.bias"], AR)
    raise ValueError(kind)


def _our_layer(m, kind, name, cin, cout, xg, next_gamma):
    """One layer of the host mirror on a channels-last GPU tensor; returns (out, normed or None)."""
    from more4d_b200 import ops
    if kind == "conv":
        return m._conv(xg, name, (3, 3, 3), cout, (2, 1, 1)), None
    if kind == "res":
        return m._res(xg, name, cin, cout, None, next_gamma)
    if kind == "attn":
        return m._attn(xg, name), None
    if kind in ("down2d", "down3d"):
        return m._down(xg, name, kind, cin), None
    if kind in ("up2d", "up3d"):
        return m._up(xg, name, kind, cin, next_gamma)
    if kind == "head":
        xn = ops.rmsnorm_silu_cl(xg, m._p(name + ".0.gamma"))
        cpad = max(cout, 16)
        return m._conv(xn, name + ".2", (3, 3, 3), cout, (2, 1, 1))[..., :cout] if cpad == cout else \
            m._conv(xn, name + ".2", (3, 3, 3), cout, (2, 1, 1)), None
    raise ValueError(kind)


@pytest.mark.parametrize("side", ["encoder", "decoder"])
def test_vae_layerwise_teacher_forced(side):
    from more4d_b200.vae_arch import WAN_VAE, decoder_layers, encoder_layers
    sd = synth.vae_state_dict(seed=SEED)
    m = _vae()
    if side == "encoder":
        layers = encoder_layers(WAN_VAE)
        x = AR.r(synth._randn(SEED, "vae.x", (1, 3, 9, 32, 48), 0.5, "cpu", BF16).float())
        # first layer consumes the 3-channel video: run it through the mirror's own input path
        x = V.causal_conv3d(x, sd["model.encoder.conv1.weight"], sd["model.encoder.conv1.bias"], AR)
        layers = layers[1:]
    else:
        layers = decoder_layers(WAN_VAE)
        x = AR.r(synth._randn(SEED, "vae.z2", (1, 16, 3, 4, 6), 1.0, "cpu", BF16).float())
        x = torch.cat([x, torch.zeros(1, 16, 3, 4, 6)], dim=1)        # the mirror keeps the latent in 32 channels
    worst = {}
    with torch.no_grad():
        for i, (kind, name, cin, cout) in enumerate(layers):
            p = "model." + name
            xin = x[:, :cin] if kind == "conv" and x.shape[1] != cin else x
            y = _oracle_layer(kind, xin, sd, p)
            ng = m._next_gamma(layers, i, None)
            if kind == "head" and cout < 16:
                # 3-channel video head: the mirror writes planar output with the clamp fused (vae:827)
                from more4d_b200 import ops
                xn = ops.rmsnorm_silu_cl(_cl(x.to(BF16)), m._p(name + ".0.gamma"))
                vid = torch.empty(cout, *xn.shape[:3], device="cuda", dtype=BF16)
                m._conv(xn, name + ".2", (3, 3, 3), cout, (2, 1, 1), planar_out=vid, act=1)
                got, want = vid.float().cpu().unsqueeze(0), y.clamp(-1, 1)
                gn = None
            else:
                out, gn = _our_layer(m, kind, name, cin, cout, _cl(x.to(BF16)), ng)
                got, want = _ncthw(out)[:, :y.shape[1]], y
            e = rel_err(got, want)
            worst[kind] = max(worst.get(kind, 0.0), e)
            assert e < LAYER_TOL, (side, i, kind, name, e)
            if gn is not None:          # fused epilogue: the consumer's SiLU(RMS_norm(.)) of the same output
                nxt = "model." + layers[i + 1][1]
                wn = V.silu(V.rms_norm(y, sd[nxt + ".residual.0.gamma"], AR), AR)
                en = rel_err(_ncthw(gn), wn)
                worst[kind + "+norm"] = max(worst.get(kind + "+norm", 0.0), en)
                assert en < 5e-3, (side, i, kind, name, "fused norm", en)
            x = y                                                    # teacher forcing
    print(side, {k: f"{v:.2e}" for k, v in worst.items()})


def test_vae_attention_block_isolated():
    """AttentionBlock (vae:227-266): RMS_norm -> to_qkv 1x1 -> per-frame
    single-head attention with head_dim = C = 384 (GEMM -> row softmax -> GEMM on the validated GEMM
    kernel) -> proj 1x1 (zero-initialised in the reference, non-zero here, F6) + residual."""
    sd = synth.vae_state_dict(seed=SEED)
    m = _vae()
    for (T, H, W) in [(1, 4, 6), (3, 9, 13), (2, 16, 24)]:       # 24 / 117 / 384 tokens per frame
        x = AR.r(_rand((1, 384, T, H, W), 50 + T, 1.5).float())
        for name in ("encoder.middle.1", "decoder.middle.1"):
            want = V.attention_block(x, sd, "model." + name, AR)
            with torch.no_grad():
                got = _ncthw(m._attn(_cl(x.to(BF16)), name))
            assert rel_err(got, want) < 3e-3, (name, T, H, W)
            # the attention branch itself, without the pass-through: five bf16 rounding stages, and the kernels hand P
            # to the second GEMM in bf16 (as flash / mem-efficient SDPA do) where the oracle keeps it in fp32
            assert rel_err(got - x, want - x) < 1.5e-2
