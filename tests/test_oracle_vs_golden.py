"""Pin the CPU oracle (oracle/dit_oracle.py) against golden vectors produced by the real
reference modules (tests/golden/make_golden.py).  fp32-vs-fp32, so the tolerance is tight."""
import math

import pytest
import torch

from more4d_b200 import synth
from more4d_b200.config import WAN_1_3B, WAN_TINY
from oracle import dit_oracle as O
from tests.helpers import checksum, rel_err

TOL = 2e-5


@pytest.fixture(autouse=True)
def _sdpa_semantics():
    # goldens come from the reference's SDPA branch, which ignores k_lens (see dit_oracle.py)
    O.RESPECT_K_LENS = False
    yield
    O.RESPECT_K_LENS = True


def test_leaf_ops(golden):
    g = golden("dit_ops")
    seed, B, N, D, L, grid = 3, 2, 2, 128, 30, (2, 3, 4)
    q = synth._randn(seed, "op.q", (B, L, N, D), 1.0, "cpu", torch.float32)
    k = synth._randn(seed, "op.k", (B, L, N, D), 1.0, "cpu", torch.float32)
    v = synth._randn(seed, "op.v", (B, L, N, D), 1.0, "cpu", torch.float32)
    assert torch.allclose(checksum(q), g["q_sum"], rtol=1e-6), "RNG drift: regenerate goldens"
    ar = O.Arith(False)
    assert rel_err(O.rope_apply(q, [grid] * B, ar), g["rope"]) < 1e-6
    w = synth._randn(seed, "op.w", (N * D,), 0.1, "cpu", torch.float32, mean=1.0)
    assert rel_err(O.rms_norm(q.flatten(2), w, 1e-6, ar), g["rms"]) < 1e-6
    assert rel_err(O.attention(q, k, v, [L, L], ar), g["attn"]) < 1e-5
    sin = O.sinusoidal_embedding(256, torch.tensor([500.0, 999.0, 0.0])).float()
    assert rel_err(sin, g["sinus"]) < 1e-6


def _block_inputs(cfg, seed, seq_len, grid, guidance):
    C = cfg.dim
    sd = synth.block_state_dict(cfg, 0, seed)
    x = synth._randn(seed, "blk.x", (1, seq_len, C), 1.0, "cpu", torch.bfloat16)
    ctx = synth._randn(seed, "blk.ctx", (1, 257 + cfg.text_len, C), 1.0, "cpu", torch.bfloat16)
    tsd = synth.dit_state_dict(cfg, seed, prefix_filter="time_")
    _, e0 = O.time_embed(torch.tensor([500.0]), tsd, cfg.freq_dim, C)
    n_tok = math.prod(grid)
    feats = None
    if guidance:
        feats = (synth._randn(seed, "blk.dino", (1, n_tok, cfg.guidance_dim), 1.0, "cpu", torch.float32),
                 synth._randn(seed, "blk.cls", (1, 1, cfg.guidance_dim), 1.0, "cpu", torch.float32))
    return sd, x, ctx, e0, n_tok, feats


@pytest.mark.parametrize("name,cfg,grid,seq_len,seed,guidance", [
    ("block_tiny", WAN_TINY, (2, 3, 4), 30, 1, False),
    ("block_tiny_mpm", WAN_TINY.with_(use_spatial_guidance=True), (2, 3, 4), 30, 2, True),
    ("block_config1", WAN_1_3B.with_(num_layers=1), (2, 9, 16), 288, 0, False),
])
def test_block(golden, name, cfg, grid, seq_len, seed, guidance):
    g = golden(name)
    sd, x, ctx, e0, n_tok, feats = _block_inputs(cfg, seed, seq_len, grid, guidance)
    assert torch.allclose(checksum(x), g["x_sum"], rtol=1e-6), "RNG drift: regenerate goldens"
    assert torch.allclose(e0, g["e0"], rtol=1e-5, atol=1e-6)
    y = O.block_forward(x, e0, sd, cfg.num_heads, cfg.eps, [n_tok], [grid], ctx.float(),
                        emulate_bf16=False, guidance=feats)
    assert rel_err(y, g["y"]) < TOL
    # increment-level error (the residual pass-through masks errors at the output level)
    assert rel_err(y - x.float(), g["y"] - x.float()) < 5 * TOL
    # the bf16-emulation mode (what the CUDA path is compared with) stays within the
    # north-star tolerance of the gold result
    yb = O.block_forward(x, e0, sd, cfg.num_heads, cfg.eps, [n_tok], [grid], ctx.float(),
                         emulate_bf16=True, guidance=feats)
    assert rel_err(yb, g["y"]) < 2e-3


def test_model_tiny(golden):
    g = golden("dit_tiny")
    cfg, grid, batch, seed = WAN_TINY, (3, 4, 6), 2, 4
    sd = synth.dit_state_dict(cfg, seed)
    inp = synth.dit_inputs(cfg, grid, batch, seed)
    assert torch.allclose(checksum(inp["x"]), g["x_sum"], rtol=1e-6), "RNG drift: regenerate goldens"
    y = O.dit_forward(sd, cfg, inp["x"].float(), inp["t"], [c.float() for c in inp["context"]],
                      inp["seq_len"], clip_fea=inp["clip_fea"].float(), y=inp["y"].float(),
                      full_ref=inp["full_ref"].float())
    assert y.shape == g["y"].shape
    assert rel_err(y, g["y"]) < TOL


def test_model3d_tiny(golden):
    """The 4D-ViSM backbone (WanTransformer3DModel, in_dim 36, no reference conv): the same oracle
    functions reproduce the REAL 3D class (golden made by tests/golden/make_golden.py)."""
    from more4d_b200.config import WAN_TINY_INP
    g = golden("dit3d_tiny")
    cfg, grid, batch, seed = WAN_TINY_INP, (3, 4, 6), 2, 5
    sd = synth.dit_state_dict(cfg, seed)
    assert not any("ref_conv" in k or "spatial_guidance" in k for k in sd)
    inp = synth.dit_inputs(cfg, grid, batch, seed, with_ref=False)
    assert torch.allclose(checksum(inp["x"]), g["x_sum"], rtol=1e-6), "RNG drift: regenerate goldens"
    y = O.dit_forward(sd, cfg, inp["x"].float(), inp["t"], [c.float() for c in inp["context"]],
                      inp["seq_len"], clip_fea=inp["clip_fea"].float(), y=inp["y"].float(), full_ref=None)
    assert y.shape == g["y"].shape
    assert rel_err(y, g["y"]) < TOL


def test_block_14b_dims(golden):
    """The oracle at the headline dims (C 5120, F 13824, 40 heads, L 1152) against the real
    WanAttentionBlock's sampled rows (tests/golden/make_golden.py::block14b_case)."""
    from more4d_b200.config import WAN_14B
    g = golden("block_14b")
    cfg, seed, grid, L = WAN_14B.with_(num_layers=1), 6, (2, 24, 24), 1152
    sd = synth.block_state_dict(cfg, 0, seed)
    x = synth._randn(seed, "blk.x", (1, L, cfg.dim), 1.0, "cpu", torch.bfloat16)
    ctx = synth._randn(seed, "blk.ctx", (1, 257 + cfg.text_len, cfg.dim), 1.0, "cpu", torch.bfloat16)
    e0 = synth._randn(seed, "blk.e0", (1, 6, cfg.dim), 0.3, "cpu", torch.float32)
    assert torch.allclose(checksum(x), g["x_sum"], rtol=1e-6), "RNG drift: regenerate goldens"
    w_sum = checksum(torch.cat([v.flatten().float() for v in sd.values()]))
    assert torch.allclose(w_sum, g["w_sum"], rtol=1e-5), "RNG drift in the 14B block weights"
    y = O.block_forward(x, e0, sd, cfg.num_heads, cfg.eps, [L], [grid], ctx.float())
    rows = g["rows"].long()
    assert rel_err(y[0, rows], g["y_rows"]) < TOL
    assert rel_err(y[0, rows] - x[0, rows].float(), g["inc_rows"]) < 5 * TOL
    assert torch.allclose(checksum(y), g["y_sum"], rtol=1e-4)


def test_model_with_motion_perception_front_end(golden):
    """Oracle front end (feature_adapter -> bilinear -> repeat, t4d:1146-1152) + SpatialGuidance in
    every block against the REAL model forward with `first_frame` (golden dit_tiny_mpm; the OmniMAE
    trunk is the deterministic stub of oracle/ref_import.py on both sides)."""
    from oracle.ref_import import StubOmniMAE
    g = golden("dit_tiny_mpm")
    cfg = WAN_TINY.with_(use_spatial_guidance=True, use_omnimae_guidance=True)
    seed, grid, batch = 8, (3, 4, 6), 2
    sd = synth.dit_state_dict(cfg, seed)
    inp = synth.dit_inputs(cfg, grid, batch, seed)
    ff = synth._randn(seed, "in.first_frame", (batch, 3, 40, 56), 0.25, "cpu", torch.float32, mean=0.5).clamp(0, 1)
    assert torch.allclose(checksum(ff), g["ff_sum"], rtol=1e-6), "RNG drift: regenerate goldens"
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    trunk = StubOmniMAE().trunk
    toks, cls = zip(*[trunk.forward_patch_features(((ff - mean) / std)[i:i + 1], None) for i in range(batch)])
    ar = O.Arith(False)
    feats, adapter = O.mpm_front_end(torch.cat(toks), torch.cat(cls), {k: v.float() for k, v in sd.items()},
                                     grid[1:], grid[0] - 1, ar, return_adapter=True)
    assert rel_err(adapter, g["adapter_out"]) < TOL
    y = O.dit_forward(sd, cfg, inp["x"].float(), inp["t"], [c.float() for c in inp["context"]], inp["seq_len"],
                      clip_fea=inp["clip_fea"].float(), y=inp["y"].float(), full_ref=inp["full_ref"].float(),
                      guidance=feats)
    assert rel_err(y, g["y"]) < TOL


def test_teacache_decisions_match_the_reference(golden):
    """`cache_utils.teacache_decide / teacache_step_done` (the decision block the CUDA mirror runs,
    dit._forward) on the oracle's e0 sequence reproduce the skip pattern of the REAL model with the
    reference's own TeaCache (golden dit_tiny_teacache) — with our TeaCache class and with a
    duck-typed object carrying only the reference's fields."""
    import types
    import numpy as np
    from more4d_b200.cache_utils import TeaCache, teacache_decide, teacache_step_done
    from tests.golden.make_golden import TEACACHE_KW, TEACACHE_TS
    g = golden("dit_tiny_teacache")
    sd = synth.dit_state_dict(WAN_TINY, 4)
    want = [bool(v) for v in g["should_calc"].tolist()]
    assert True in want and False in want                       # the case exercises both branches
    ours = TeaCache(TEACACHE_KW["coefficients"], len(TEACACHE_TS), rel_l1_thresh=TEACACHE_KW["rel_l1_thresh"],
                    num_skip_start_steps=TEACACHE_KW["num_skip_start_steps"], offload=False)
    duck = types.SimpleNamespace(cnt=0, num_steps=len(TEACACHE_TS), num_skip_start_steps=TEACACHE_KW["num_skip_start_steps"],
                                 rel_l1_thresh=TEACACHE_KW["rel_l1_thresh"], accumulated_rel_l1_distance=0,
                                 rescale_func=np.poly1d(TEACACHE_KW["coefficients"]), previous_modulated_input=None,
                                 should_calc=True, previous_residual=None, previous_residual_cond=None,
                                 previous_residual_uncond=None)
    for tc in (ours, duck):
        got = []
        for t in TEACACHE_TS:
            _, e0 = O.time_embed(torch.tensor([t, t]), sd, WAN_TINY.freq_dim, WAN_TINY.dim)
            got.append(teacache_decide(tc, e0, True))
            teacache_step_done(tc, True)
        assert got == want
        assert tc.cnt == 0 and tc.previous_modulated_input is None      # reset after num_steps, t4d:1336-1339
