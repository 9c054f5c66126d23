"""GPU parity tests: each CUDA kernel (through the C ABI via more4d_b200.ops) against the CPU
oracle on the same seeded inputs.  Tolerances: bf16-output kernels are compared with the
oracle's bf16-emulation mode at a relative Frobenius error of a few bf16 ulps (stated per
test); the north-star bound (<= 1e-3 on the block output vs the fp32 gold) is asserted in
test_block_gpu.py."""
import math

import pytest
import torch

from oracle import dit_oracle as O
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


def _rand(shape, seed, scale=1.0, dtype=BF16):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype)


@pytest.fixture(scope="module")
def ops():
    from more4d_b200 import ops as _ops
    return _ops


# ------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [
    (128, 256, 64),        # exactly one tile, one k-block
    (128, 256, 512),       # pipeline wraps (8 k-blocks, 4 stages)
    (300, 320, 192),       # M and N tails
    (1000, 64, 256),       # N < tile (head projection shape)
    (577, 1536, 1536),     # many tiles per CTA? no: 5x6 tiles, multi k
    (2 * 4096, 1024, 256), # 64 x 4 tiles: persistent loop + accumulator double buffering
])
def test_gemm_bias_bf16(ops, M, N, K):
    a, w, b = _rand((M, K), 1), _rand((N, K), 2, 0.05), _rand((N,), 3, 0.1)
    ref = O.Arith(True).linear(a.float(), w, b)
    out = ops.linear(a.cuda(), w.cuda(), b.cuda()).float().cpu()
    assert rel_err(out, ref) < 2e-3          # <= ~1 bf16 ulp disagreements from summation order
    assert (out - ref).abs().max() <= 0.02 * ref.abs().max()


def test_gemm_persistent_many_tiles(ops):
    # more tiles than SMs so every CTA loops: 160 x 2 tiles of 128x256
    M, N, K = 160 * 128, 512, 128
    a, w = _rand((M, K), 4), _rand((N, K), 5, 0.05)
    ref = O.Arith(True).linear(a.float(), w, None)
    out = ops.linear(a.cuda(), w.cuda(), None).float().cpu()
    assert rel_err(out, ref) < 2e-3


@pytest.mark.parametrize("epi", ["gelu_tanh", "gelu_erf", "f32"])
def test_gemm_epilogues(ops, epi):
    M, N, K = 333, 512, 256
    a, w, b = _rand((M, K), 6), _rand((N, K), 7, 0.05), _rand((N,), 8, 0.1)
    y = O.Arith(True).linear(a.float(), w, b)
    if epi == "gelu_tanh":
        ref = O.Arith(True).r(torch.nn.functional.gelu(y, approximate="tanh"))
        out = ops.linear(a.cuda(), w.cuda(), b.cuda(), ops.EPI_GELU_TANH)
    elif epi == "gelu_erf":
        ref = O.Arith(True).r(torch.nn.functional.gelu(y))
        out = ops.linear(a.cuda(), w.cuda(), b.cuda(), ops.EPI_GELU_ERF)
    else:
        ref = y
        out = ops.linear(a.cuda(), w.cuda(), b.cuda(), ops.EPI_F32)
        assert out.dtype == torch.float32
    assert rel_err(out.float().cpu(), ref) < 2e-3


def test_gemm_gate_residual_inplace(ops):
    B, L, N, K = 2, 200, 512, 256
    a, w, b = _rand((B * L, K), 9), _rand((N, K), 10, 0.05), _rand((N,), 11, 0.1)
    x = _rand((B, L, N), 12, dtype=torch.float32)
    em = _rand((B, 6, N), 13, dtype=torch.float32)
    y = O.Arith(True).linear(a.float(), w, b).view(B, L, N)
    ref = x + y * em[:, 2:3]
    xg, emg = x.cuda().clone(), em.cuda()
    out = ops.linear(a.cuda().view(B, L, K), w.cuda(), b.cuda(), ops.EPI_GATE_RESIDUAL_F32, out=xg,
                     residual=xg, gate=emg[:, 2], gate_batch_stride=6 * N, rows_per_batch=L)
    assert out.data_ptr() == xg.data_ptr()
    assert rel_err(xg.cpu(), ref) < 5e-4
    # gate = None  ->  plain residual add (cross-attention)
    xg2 = x.cuda().clone()
    ops.linear(a.cuda().view(B, L, K), w.cuda(), b.cuda(), ops.EPI_GATE_RESIDUAL_F32, out=xg2, residual=xg2)
    assert rel_err(xg2.cpu(), x + y) < 5e-4


@pytest.mark.parametrize("epi,M,N,K", [
    ("bf16", 1024, 256, 512),            # smallest shape the 2-CTA kernel takes (M >= 1024, N >= 256): one cluster tile column
    ("bf16", 1500, 5120, 5120),          # q/k/v/o shape at 14B dims, M tail inside a 256-row cluster tile
    ("gelu_tanh", 1280, 13824, 5120),    # ffn.0 + GELU(tanh) at 14B dims (t4d:620-622)
    ("gelu_tanh", 1100, 704, 256),       # N tail (704 = 2.75 tiles) with the GELU epilogue
    ("gate_residual", 1152, 5120, 13824),  # ffn.2 at 14B dims: K = 13 824 accumulation depth, x += y * e5 (t4d:683-684)
    ("gate_residual", 2048, 640, 5120),  # o-proj style, two batches of 1024 rows, N tail
    ("residual", 1024, 5120, 5120),      # cross-attention o-proj: gate = NULL (t4d:674)
])
def test_gemm2_two_cta_kernel_epilogues(ops, epi, M, N, K):
    """`gemm2_bf16_tn_kernel<EPI>` — the cta_group::2 kernel every block GEMM of the benchmark runs
    (csrc/gemm.cu routes M >= 1024, N >= 256 to it) — against the oracle's autocast Linear for all
    three of its epilogues, incl. the FFN shapes of the 14B config."""
    a, w, b = _rand((M, K), 14), _rand((N, K), 15, 0.02), _rand((N,), 16, 0.1)
    y = O.Arith(True).linear(a.float(), w, b)                     # bf16-rounded, like autocast nn.Linear
    if epi == "bf16":
        out = ops.linear(a.cuda(), w.cuda(), b.cuda()).float().cpu()
        ref = y
        tol = 2e-3
    elif epi == "gelu_tanh":
        out = ops.linear(a.cuda(), w.cuda(), b.cuda(), ops.EPI_GELU_TANH).float().cpu()
        ref = O.Arith(True).r(torch.nn.functional.gelu(y, approximate="tanh"))
        tol = 2e-3
    else:
        B = 2
        L = M // B
        x = _rand((B, L, N), 17, dtype=torch.float32)
        em = _rand((B, 6, N), 18, dtype=torch.float32)
        xg, emg = x.cuda().clone(), em.cuda()
        if epi == "gate_residual":
            ref = (x + y.view(B, L, N) * em[:, 5:6]).view(M, N)
            ops.linear(a.cuda().view(B, L, K), w.cuda(), b.cuda(), ops.EPI_GATE_RESIDUAL_F32, out=xg, residual=xg,
                       gate=emg[:, 5], gate_batch_stride=6 * N, rows_per_batch=L)
        else:
            ref = (x + y.view(B, L, N)).view(M, N)
            ops.linear(a.cuda().view(B, L, K), w.cuda(), b.cuda(), ops.EPI_GATE_RESIDUAL_F32, out=xg, residual=xg)
        out = xg.cpu().view(M, N)
        tol = 5e-4
    assert torch.isfinite(out).all()
    assert rel_err(out, ref) < tol
    # no column / row block is systematically off (a mis-addressed half tile would hide in the Frobenius norm)
    blk = (out - ref).abs().view(M // 4, 4, N).amax(1)
    assert float(blk.max()) <= 0.03 * float(ref.abs().max()) + 1e-3


def test_gemm_rejects_bad_args(ops):
    a, w = _rand((16, 60), 1).cuda(), _rand((8, 60), 2).cuda()      # K % 8 != 0
    with pytest.raises(RuntimeError):
        ops.linear(a, w, None)
    with pytest.raises(RuntimeError):
        ops.linear(_rand((4, 64), 1), _rand((8, 64), 2), None)       # CPU tensors: no fallback


# ------------------------------------------------------------------------------- attention
def _attn_case(ops, B, Lq, Lk, N, k_lens=None, qscale=1.0, seed=0):
    q = _rand((B, Lq, N, 128), seed + 1, qscale)
    k = _rand((B, Lk, N, 128), seed + 2)
    v = _rand((B, Lk, N, 128), seed + 3)
    ref = O.attention(q, k, v, k_lens, O.Arith(True))
    kl = None if k_lens is None else torch.tensor(k_lens, dtype=torch.int32, device="cuda")
    out = ops.attention(q.cuda(), k.cuda(), v.cuda(), kl).float().cpu()
    return out, ref


@pytest.mark.parametrize("B,Lq,Lk,N", [
    (1, 128, 128, 1),      # single tile
    (1, 256, 384, 2),      # full tiles, 3 kv tiles
    (2, 288, 288, 3),      # BASELINE config-1 sequence (2x9x16), q and kv tails
    (2, 300, 257, 2),      # cross-attention: CLIP tokens
    (1, 77, 512, 2),       # cross-attention: text tokens, Lq < one tile
    (1, 1300, 1300, 1),    # 11 kv tiles
])
def test_attention_matches_oracle(ops, B, Lq, Lk, N):
    out, ref = _attn_case(ops, B, Lq, Lk, N)
    assert torch.isfinite(out).all()
    # bf16 P and bf16 output: a few 1e-3; the oracle rounds only the output
    assert rel_err(out, ref) < 6e-3


def test_attention_k_lens_masks_keys(ops):
    out, ref = _attn_case(ops, 2, 200, 300, 2, k_lens=[300, 131])
    assert rel_err(out, ref) < 6e-3


def test_attention_large_logits_rescale_path(ops):
    # logits with std ~ 8 and a growing running max exercise the lazy O rescale
    out, ref = _attn_case(ops, 1, 256, 1024, 1, qscale=8.0, seed=5)
    assert torch.isfinite(out).all()
    assert rel_err(out, ref) < 1e-2


@pytest.mark.parametrize("gain", [3.0, 8.0, 25.0])
def test_attention_reference_guard_paths(ops, gain):
    """The exponentials run against the reference left by earlier key tiles and the half-tile row
    sums guard it (csrc/attention.cu MODE 1).  Planted keys aligned with single query rows make
    the guard fire in the FIRST half of a tile (exact path before anything is stored) and in the
    SECOND half (the first half is already with the tensor pipe: wait for its MMAs, rescale O) —
    with logit jumps of ~34 / ~90 / ~280 nats, i.e. below the guard, above it, and far beyond
    fp32 overflow of the speculative 2^x."""
    B, Lq, Lk, N = 1, 256, 1024, 2
    q = _rand((B, Lq, N, 128), 41)
    k = _rand((B, Lk, N, 128), 42)
    v = _rand((B, Lk, N, 128), 43)
    plant = [(3, 2 * 128 + 5, 0), (77, 2 * 128 + 100, 0), (130, 5 * 128 + 64, 1), (200, 7 * 128 + 127, 1),
             (201, 1 * 128 + 0, 0), (201, 6 * 128 + 90, 0)]          # (row, key, head); row 201 jumps twice
    for i, (r, kk, h) in enumerate(plant):
        k[0, kk, h] = (q[0, r, h].float() * gain * (1.0 + 0.3 * (i == len(plant) - 1))).to(BF16)
    ref = O.attention(q, k, v, None, O.Arith(True))
    out = ops.attention(q.cuda(), k.cuda(), v.cuda()).float().cpu()
    assert torch.isfinite(out).all()
    assert rel_err(out, ref) < 1e-2
    rows = [r for r, _, _ in plant]
    assert rel_err(out[0, rows], ref[0, rows]) < 1e-2                # the rows that took the exact paths


def test_attention_strided_views_and_accumulate(ops):
    B, L, N = 2, 150, 2
    qkv = _rand((B, L, 3, N, 128), 21)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    ar = O.Arith(True)
    ref1 = O.attention(q, k, v, None, ar)
    g = qkv.cuda()
    out = ops.attention(g[:, :, 0], g[:, :, 1], g[:, :, 2])
    assert rel_err(out.float().cpu(), ref1) < 6e-3
    k2, v2 = _rand((B, 257, N, 128), 22), _rand((B, 257, N, 128), 23)
    ref2 = ar.r(ref1 + O.attention(q, k2, v2, None, ar))
    ops.attention(g[:, :, 0], k2.cuda(), v2.cuda(), out=out, accumulate=True)
    assert rel_err(out.float().cpu(), ref2) < 6e-3


@pytest.mark.parametrize("B,Lq,seg,Lk", [(2, 300, 512, 769), (1, 513, 128, 129), (1, 90, 256, 512)])
def test_attention_two_segments_one_launch(ops, B, Lq, seg, Lk):
    """m4d_attention_fwd_seg2 (text + image cross-attention in one launch, t4d:533-552) is
    BIT-IDENTICAL to the two-launch form (attention, then attention(accumulate=True)) and matches
    the oracle's sum of two separately normalised attentions."""
    N = 2
    q = _rand((B, Lq, N, 128), 61)
    k = _rand((B, Lk, N, 128), 62)
    v = _rand((B, Lk, N, 128), 63)
    k[0, Lk - 1, 0] = (q[0, 7, 0].float() * 6.0).to(BF16)         # second segment far above the first
    k[0, 5, 1] = (q[0, 9, 1].float() * 6.0).to(BF16)              # and the other way round
    ar = O.Arith(True)
    ref = ar.r(O.attention(q, k[:, :seg], v[:, :seg], None, ar) + O.attention(q, k[:, seg:], v[:, seg:], None, ar))
    qg, kg, vg = q.cuda(), k.cuda(), v.cuda()
    one = ops.attention_seg2(qg, kg, vg, seg)
    two = ops.attention(qg, kg[:, :seg], vg[:, :seg])
    ops.attention(qg, kg[:, seg:], vg[:, seg:], out=two, accumulate=True)
    assert torch.equal(one, two)
    assert rel_err(one.float().cpu(), ref) < 6e-3
    with pytest.raises(ValueError):
        ops.attention_seg2(qg, kg, vg, seg + 1)


@pytest.mark.parametrize("Lk,seg", [(769, 512), (300, 0), (128, 0), (2048, 1024)])
def test_attention_persistent_ctas_short_key_ranges(ops, Lk, seg):
    """Key ranges up to 2048 with more work items than SMs run on PERSISTENT CTAs (one per SM looping
    over (q block, head, batch) items, barrier phases on running counters).  Rows are independent,
    so the result must be BIT-IDENTICAL to the same queries launched in chunks small enough for the
    one-CTA-per-item kernel (<= 148 items), and match the oracle."""
    B, Lq, N = 2, 2560 + 77, 8                       # 11 q blocks x 8 heads x 2 = 176 items > 148 SMs
    q = _rand((B, Lq, N, 128), 81)
    k = _rand((B, Lk, N, 128), 82)
    v = _rand((B, Lk, N, 128), 83)
    qg, kg, vg = q.cuda(), k.cuda(), v.cuda()
    run = (lambda qq: ops.attention_seg2(qq, kg, vg, seg)) if seg else (lambda qq: ops.attention(qq, kg, vg))
    big = run(qg)
    parts = torch.cat([run(qg[:, a:a + 1024].contiguous()) for a in range(0, Lq, 1024)], dim=1)   # 4 x 8 x 2 = 64 items
    assert torch.equal(big, parts)
    ar = O.Arith(True)
    rows = slice(1500, 1500 + 300)                   # oracle on a slice of the queries (cost)
    if seg:
        ref = ar.r(O.attention(q[:, rows], k[:, :seg], v[:, :seg], None, ar) +
                   O.attention(q[:, rows], k[:, seg:], v[:, seg:], None, ar))
    else:
        ref = O.attention(q[:, rows], k, v, None, ar)
    assert rel_err(big[:, rows].float().cpu(), ref) < 6e-3
    if not seg:                                      # accumulate form on persistent CTAs
        acc = big.clone()
        ops.attention(qg, kg, vg, out=acc, accumulate=True)
        two = ar.r(big.float().cpu() * 2)
        assert rel_err(acc.float().cpu(), two) < 1e-2


def test_attention_rejects_head_dim(ops):
    q = _rand((1, 64, 2, 64), 1).cuda()
    with pytest.raises(RuntimeError):
        ops.attention(q, q, q)


# ------------------------------------------------------------------------------ row kernels
@pytest.mark.parametrize("C", [256, 1536, 5120])
def test_layernorm_modulate(ops, C):
    B, L = 2, 37
    x = _rand((B, L, C), 31, 2.0, torch.float32) + 0.5
    em = _rand((B, 6, C), 32, 0.3, torch.float32)
    ref = O.Arith(True).r(O.layer_norm(x, None, None, 1e-6) * (1 + em[:, 1:2]) + em[:, 0:1])
    emg = em.cuda()
    out = ops.layernorm_modulate(x.cuda(), None, None, emg[:, 0], emg[:, 1], 6 * C, L, 1e-6)
    assert rel_err(out.float().cpu(), ref) < 2e-3
    w, b = _rand((C,), 33, 0.1) + 1, _rand((C,), 34, 0.1)
    ref = O.Arith(True).r(O.layer_norm(x, w, b, 1e-6))
    out = ops.layernorm_modulate(x.cuda(), w.cuda(), b.cuda(), eps=1e-6)
    assert rel_err(out.float().cpu(), ref) < 2e-3
    xb = x.to(BF16)
    ref = O.layer_norm(xb, w, b, 1e-5)
    out = ops.layernorm_modulate(xb.cuda(), w.cuda(), b.cuda(), eps=1e-5, out_dtype=torch.float32)
    assert rel_err(out.cpu(), ref) < 1e-5


@pytest.mark.parametrize("N", [2, 12])
def test_rmsnorm_rope(ops, N):
    from more4d_b200.dit import build_freqs
    B, grid, L = 2, (2, 3, 4), 30
    x = _rand((B, L, N * 128), 41)
    w = _rand((N * 128,), 42, 0.1) + 1
    ar = O.Arith(True)
    ref = O.rope_apply(O.rms_norm(x, w, 1e-6, ar).view(B, L, N, 128), [grid] * B, ar)
    f = build_freqs(128)
    cos, sin = f.real.float().cuda(), f.imag.float().cuda()
    g = torch.tensor([grid] * B, dtype=torch.int32, device="cuda")
    xg = x.cuda().clone()
    ops.rmsnorm_rope_(xg, w.cuda(), N, 1e-6, cos, sin, g)
    out = xg.float().cpu().view(B, L, N, 128)
    assert rel_err(out, ref) < 2e-3
    # pass-through rows (>= F*H*W) are normed but not rotated
    assert rel_err(out[:, 24:], O.rms_norm(x, w, 1e-6, ar).view(B, L, N, 128)[:, 24:]) < 2e-3
    # norm only (cross-attention q/k)
    xg = x.cuda().clone()
    ops.rmsnorm_rope_(xg, w.cuda(), N, 1e-6)
    assert rel_err(xg.float().cpu(), O.rms_norm(x, w, 1e-6, ar)) < 2e-3


def test_time_embedding_path(ops):
    C, fd = 256, 256
    t = torch.tensor([500.0, 999.0, 3.0])
    sd = {"time_embedding.0.weight": _rand((C, fd), 51, 0.05), "time_embedding.0.bias": _rand((C,), 52, 0.05),
          "time_embedding.2.weight": _rand((C, C), 53, 0.05), "time_embedding.2.bias": _rand((C,), 54, 0.05),
          "time_projection.1.weight": _rand((6 * C, C), 55, 0.05), "time_projection.1.bias": _rand((6 * C,), 56, 0.05)}
    e_ref, e0_ref = O.time_embed(t, sd, fd, C)
    s = ops.timestep_embedding(t.cuda(), fd)
    assert rel_err(s.cpu(), O.sinusoidal_embedding(fd, t).float()) < 1e-6
    g = {k: v.cuda() for k, v in sd.items()}
    h = ops.small_linear_f32(s, g["time_embedding.0.weight"], g["time_embedding.0.bias"], silu_out=True)
    e = ops.small_linear_f32(h, g["time_embedding.2.weight"], g["time_embedding.2.bias"])
    e0 = ops.small_linear_f32(e, g["time_projection.1.weight"], g["time_projection.1.bias"], silu_in=True)
    assert rel_err(e.cpu(), e_ref) < 1e-5
    assert rel_err(e0.cpu().view(3, 6, C), e0_ref) < 1e-5


def test_patchify_unpatchify_roundtrip(ops):
    B, T, H, W = 2, 3, 8, 12
    x, y = _rand((B, 16, T, H, W), 61), _rand((B, 48, T, H, W), 62)
    w = _rand((32, 64, 1, 2, 2), 63, 0.05)
    b = _rand((32,), 64, 0.05)
    ar = O.Arith(True)
    ref = torch.stack([O.patch_embed(torch.cat([x[i], y[i]]), w, b, ar) for i in range(B)])
    cols = ops.patchify(x.cuda(), y.cuda())
    out = ops.linear(cols, w.cuda().view(32, -1), b.cuda())
    assert rel_err(out.float().cpu(), ref) < 2e-3
    tok = _rand((B, 5 + T * (H // 2) * (W // 2), 64), 65)
    ref = torch.stack([O.unpatchify(tok[i, 5:].float(), (T, H // 2, W // 2), (1, 2, 2), 16) for i in range(B)])
    out = ops.unpatchify(tok.cuda(), 5, 16, T, H, W)
    assert torch.equal(out.float().cpu(), ref)


def test_cfg_euler_step(ops):
    n = (1, 16, 3, 8, 12)
    u, tx, lat = _rand(n, 71), _rand(n, 72), _rand(n, 73)
    g, dt = 6.0, -0.0123
    d = (tx.float() - u.float()).to(BF16)
    npred = (u.float() + (g * d.float()).to(BF16).float()).to(BF16)
    ref = (lat.float() + dt * npred.float()).to(BF16)
    out = ops.cfg_euler_step_(lat.cuda().clone(), u.cuda(), tx.cuda(), g, dt)
    assert rel_err(out.float().cpu(), ref.float()) < 1e-3


def test_scatter_epilogues_of_the_sequence_parallel_exchange():
    """m4d_rmsnorm_scatter and m4d_attention_fwd_scatter (the fused Ulysses exchange) with LOCAL
    destination buffers: bit-identical to the plain kernels followed by slicing."""
    from more4d_b200 import ops
    B, Ll, H, D, P = 2, 75, 4, 128, 2
    C, gc = H * D, H * D // P
    x = _rand((B, Ll, C), 41, 2.0).cuda()
    w = (_rand((C,), 42, 0.1) + 1).cuda()
    L = Ll * P
    for r in range(P):                                    # this rank's tokens land at rows r*Ll..
        dst = [torch.zeros(B, L, gc, device="cuda", dtype=BF16) for _ in range(P)]
        ops.rmsnorm_scatter(x, w, dst, r * Ll, 1e-6)
        ref = ops.rmsnorm_rope_(x.clone(), w, H, 1e-6)
        for g in range(P):
            assert torch.equal(dst[g][:, r * Ll:(r + 1) * Ll], ref[:, :, g * gc:(g + 1) * gc])
            assert float(dst[g][:, :r * Ll].abs().sum()) == 0 and float(dst[g][:, (r + 1) * Ll:].abs().sum()) == 0
    dst = [torch.zeros(B, L, gc, device="cuda", dtype=BF16) for _ in range(P)]
    ops.rmsnorm_scatter(x, None, dst, 0, 1e-6)            # weight None: plain scatter copy
    assert torch.equal(dst[1][:, :Ll], x[:, :, gc:])
    # attention: 3 destinations of 100 rows cover Lq = 300 (ragged last tile), head columns 2..3 of 6
    Bq, Lq, Lk, h, n_all = 2, 300, 333, 2, 6
    q, k, v = _rand((Bq, Lq, h, D), 43).cuda(), _rand((Bq, Lk, h, D), 44).cuda(), _rand((Bq, Lk, h, D), 45).cuda()
    kl = torch.tensor([333, 200], dtype=torch.int32, device="cuda")
    ref = ops.attention(q, k, v, kl)
    bufs = [torch.zeros(Bq, 100, n_all, D, device="cuda", dtype=BF16) for _ in range(3)]
    ops.attention_scatter(q, k, v, [b_[:, :, 2:4] for b_ in bufs], kl)
    for i, b_ in enumerate(bufs):
        assert torch.equal(b_[:, :, 2:4], ref[:, i * 100:(i + 1) * 100])
        assert float(b_[:, :, :2].abs().sum()) == 0 and float(b_[:, :, 4:].abs().sum()) == 0
