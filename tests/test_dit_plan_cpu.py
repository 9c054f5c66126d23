"""Host logic of the DiT mirror without a GPU: `ops` is replaced by shape-propagating recorders
(meta tensors), so this checks WHICH kernels the forward launches, on what shapes and in which
order — per block, with hoisted conditioning, under cfg_skip, and on the sequence-parallel path
with a fake 2-rank exchange."""
import collections

import pytest
import torch

from more4d_b200 import dit as dit_mod, ops as real_ops, synth
from more4d_b200.config import WAN_TINY

BF16 = torch.bfloat16


class _Rec:
    EPI_BF16, EPI_GELU_TANH, EPI_GELU_ERF, EPI_F32 = (real_ops.EPI_BF16, real_ops.EPI_GELU_TANH,
                                                      real_ops.EPI_GELU_ERF, real_ops.EPI_F32)
    EPI_GATE_RESIDUAL_F32, EPI_ADD_BF16, EPI_F32_RAW = (real_ops.EPI_GATE_RESIDUAL_F32, real_ops.EPI_ADD_BF16,
                                                        real_ops.EPI_F32_RAW)

    def __init__(self):
        self.calls = collections.Counter()
        self.log = []
        self.attn_shapes = []

    def _t(self, *shape, dtype=BF16):
        return torch.empty(*shape, device="meta", dtype=dtype)

    def _hit(self, name):
        self.calls[name] += 1
        self.log.append(name)

    def linear(self, x, weight, bias=None, epilogue=0, out=None, residual=None, gate=None, gate_batch_stride=0,
               rows_per_batch=0):
        self._hit("linear")
        assert x.shape[-1] == weight.reshape(weight.shape[0], -1).shape[1]
        if epilogue == self.EPI_GATE_RESIDUAL_F32:
            assert residual is not None and residual.dtype == torch.float32 and out is residual
        f32 = epilogue in (self.EPI_F32, self.EPI_GATE_RESIDUAL_F32, self.EPI_F32_RAW)
        return out if out is not None else self._t(*x.shape[:-1], weight.shape[0], dtype=torch.float32 if f32 else BF16)

    def attention(self, q, k, v, k_lens=None, softmax_scale=None, out=None, accumulate=False):
        self._hit("attention")
        assert q.shape[-1] == 128 and k.shape == v.shape and q.shape[2] == k.shape[2]
        self.attn_shapes.append((tuple(q.shape), tuple(k.shape), bool(accumulate)))
        return out if out is not None else self._t(*q.shape)

    def layernorm_modulate(self, x, weight=None, bias=None, shift=None, scale=None, mod_batch_stride=0,
                           rows_per_batch=None, eps=1e-6, out_dtype=BF16, guidance=None, guidance_gate=None):
        self._hit("layernorm_modulate")
        return self._t(*x.shape, dtype=out_dtype)

    def rmsnorm_rope_(self, x, weight, heads, eps=1e-6, rope_cos=None, rope_sin=None, grid_fhw=None):
        self._hit("rmsnorm_rope_" + ("+rope" if rope_cos is not None else "") + ("+norm" if weight is not None else ""))
        assert x.shape[-1] == heads * 128
        return x

    def small_linear_f32(self, x, weight, bias, silu_in=False, silu_out=False):
        self._hit("small_linear_f32")
        return self._t(x.shape[0], weight.shape[0], dtype=torch.float32)

    def timestep_embedding(self, t, dim):
        self._hit("timestep_embedding")
        return self._t(t.shape[0], dim, dtype=torch.float32)

    def add_bcast(self, a_bf16, e):
        self._hit("add_bcast")
        return self._t(e.shape[0], a_bf16.numel() if a_bf16.numel() >= e.shape[1] else e.shape[1], dtype=torch.float32)

    def patchify(self, x, y=None):
        self._hit("patchify")
        B, Cx, T, H, W = x.shape
        Cy = 0 if y is None else y.shape[1]
        return self._t(B, T * (H // 2) * (W // 2), (Cx + Cy) * 4)

    def unpatchify(self, tokens, skip_tokens, cout, T, H, W):
        self._hit("unpatchify")
        return self._t(tokens.shape[0], cout, T, H, W)

    def silu_bf16(self, x):
        self._hit("silu_bf16")
        return self._t(*x.shape)


@pytest.fixture
def rec(monkeypatch):
    r = _Rec()
    monkeypatch.setattr(dit_mod, "ops", r)
    monkeypatch.setattr(dit_mod, "_no_grad_only", lambda what: None)
    monkeypatch.setattr(dit_mod._rope_cache, "get", lambda freqs, device: (torch.empty(1024, 64, device="meta"),
                                                                         torch.empty(1024, 64, device="meta")))
    return r


def _model_and_inputs(layers=2):
    cfg = WAN_TINY.with_(num_layers=layers)
    m = dit_mod.WanTransformer4DModel.from_config(cfg, device="meta")
    inp = synth.dit_inputs(cfg, (3, 4, 6), 2, 0)
    mt = lambda t: t.to("meta")
    kw = dict(x=mt(inp["x"]), t=mt(inp["t"]), context=[mt(c) for c in inp["context"]], seq_len=inp["seq_len"],
              clip_fea=mt(inp["clip_fea"]), y=mt(inp["y"]), full_ref=mt(inp["full_ref"]))
    return cfg, m, kw


PER_BLOCK = {"linear": 12, "attention": 3, "layernorm_modulate": 3, "add_bcast": 1}


def test_forward_launch_plan(rec):
    cfg, m, kw = _model_and_inputs(2)
    y = m(**kw)
    assert tuple(y.shape) == (2, 16, 2, 8, 12)
    L = 3 * 4 * 6
    # per block: q,k,v,o + cross q,k,v,k_img,v_img,o + ffn0,ffn2 = 12 GEMMs; self + text + image attention
    for name, n in PER_BLOCK.items():
        per_model = {"linear": 2 * 2 + 2 + 2 + 1, "attention": 0, "layernorm_modulate": 2 + 1, "add_bcast": 1}[name]
        assert rec.calls[name] == cfg.num_layers * n + per_model, (name, rec.calls[name])
    assert rec.calls["rmsnorm_rope_+rope+norm"] == 2 * cfg.num_layers          # self-attention q, k
    assert rec.calls["rmsnorm_rope_+norm"] == 3 * cfg.num_layers               # cross q, k, k_img: no RoPE
    self_attn = [s for s in rec.attn_shapes if s[0][1] == s[1][1] == L]
    assert len(self_attn) == cfg.num_layers and self_attn[0][0] == (2, L, cfg.num_heads, 128)
    img = [s for s in rec.attn_shapes if s[1][1] == 257]
    assert len(img) == cfg.num_layers and all(acc for _, _, acc in img)         # summed into the text branch
    assert rec.log[-1] == "unpatchify" and rec.log.index("patchify") < rec.log.index("attention")


def test_hoisted_conditioning_removes_step_invariant_launches(rec):
    cfg, m, kw = _model_and_inputs(2)
    m(**kw)
    base = sum(rec.calls.values())
    rec.calls.clear()
    pre = m.precompute_conditioning(kw["context"], kw["clip_fea"])
    once = sum(rec.calls.values())
    rec.calls.clear()
    m(conditioning=pre, **kw)
    hoisted = sum(rec.calls.values())
    # context embedding (2 text + 2 image GEMMs, 2 LayerNorms) and per block 4 K/V GEMMs + 2 norms
    assert once == 4 + 2 + cfg.num_layers * 6
    assert hoisted == base - once
    assert len(pre.cross_kv) == cfg.num_layers and pre.tail(1).context.shape[0] == 1


def test_cfg_skip_halves_the_batch(rec):
    cfg, m, kw = _model_and_inputs(1)
    m.enable_cfg_skip(0.5, 10)
    m.current_steps = 9
    y = m(**kw)
    assert y.shape[0] == 2
    assert all(q[0] == 1 for q, _, _ in rec.attn_shapes)                        # only the conditional half ran


class _FakeSP:
    """A 2-rank exchange that fabricates the peers' halves (shapes only)."""
    world, rank, peer_memory = 2, 0, False

    def shard_tokens(self, x):
        return torch.empty(x.shape[0], x.shape[1] // 2, *x.shape[2:], device="meta", dtype=x.dtype)

    def seq_to_heads(self, x):
        S, B, n, H, D = x.shape
        return torch.empty(S, B, 2 * n, H // 2, D, device="meta", dtype=x.dtype)

    def heads_to_seq(self, x):
        B, L, h, D = x.shape
        return torch.empty(B, L // 2, 2 * h, D, device="meta", dtype=x.dtype)

    def gather_tokens(self, x):
        return torch.empty(x.shape[0], 2 * x.shape[1], *x.shape[2:], device="meta", dtype=x.dtype)


def test_sequence_parallel_plan(rec):
    cfg, m, kw = _model_and_inputs(1)
    m.sp, m.sp_world_size, m.sp_world_rank = _FakeSP(), 2, 0
    y = m(**kw)
    assert tuple(y.shape) == (2, 16, 2, 8, 12)
    L = 3 * 4 * 6
    q, k, _ = [s for s in rec.attn_shapes if s[1][1] == L][0]
    assert q == (2, L, cfg.num_heads // 2, 128) and k == q                     # all tokens, half the heads
    assert rec.calls["rmsnorm_rope_+norm"] == 2 + 3                            # norm before the exchange ...
    assert rec.calls["rmsnorm_rope_+rope"] == 2                                # ... RoPE after it (global positions)
    cross = [s for s in rec.attn_shapes if s[1][1] in (257, cfg.text_len)]
    assert all(s[0][1] == L // 2 for s in cross)                               # cross-attention stays token-local


def test_riflex_changes_only_the_frame_axis_table():
    cfg = WAN_TINY.with_(num_layers=1)
    m = dit_mod.WanTransformer4DModel.from_config(cfg, device="meta")
    base = m.freqs.clone()
    m.enable_riflex(k=6, L_test=66, L_test_scale=4.886)
    n_f = 128 // 2 - 2 * (128 // 6)                    # 22 frame-axis pairs, then 21 + 21
    assert m.freqs.shape == base.shape and m.freqs.dtype == torch.complex128
    diff = (m.freqs != base).any(dim=0)
    assert diff[5] and int(diff.sum()) == 1 and 5 < n_f   # only intrinsic frequency k = 6 moved
    import math
    assert torch.allclose(m.freqs[1, 5].angle(), torch.tensor(0.9 * 2 * math.pi / 66 / 4.886, dtype=torch.float64))
    m.disable_riflex()
    assert torch.equal(m.freqs, base)


class _FakePeerBuffers:
    def __init__(self, B, L, C, P):
        self.qkv = torch.empty(3, B, L, C // P, device="meta", dtype=BF16)
        self.o = torch.empty(B, L // P, C, device="meta", dtype=BF16)
        self.qkv_peers = [torch.empty(3, B, L, C // P, device="meta", dtype=BF16) for _ in range(P)]
        self.o_peers = [torch.empty(B, L // P, C, device="meta", dtype=BF16) for _ in range(P)]
        self.barriers = 0

    def barrier(self):
        self.barriers += 1


def test_fused_peer_exchange_addressing(rec):
    """The fused sequence-parallel exchange (dit._attend_sp_peer) on meta tensors: which kernel
    writes which slice of which rank's buffer, for rank 1 of 2."""
    cfg = WAN_TINY.with_(num_layers=1)
    m = dit_mod.WanTransformer4DModel.from_config(cfg, device="meta")
    attn = m.blocks[0].self_attn
    B, n_loc, C, P, r = 2, 36, cfg.dim, 2, 1
    H, d = cfg.num_heads, 128
    h, gc, L = H // P, C // P, n_loc * P
    pb = _FakePeerBuffers(B, L, C, P)

    class SP:
        world, rank, peer_memory = P, r, True

        def peer_buffers(self, B_, L_, C_, device):
            assert (B_, L_, C_) == (B, L, C)
            return pb

    v_outs, scat, att = [], [], []
    orig_linear = rec.linear

    def linear(x, weight, bias=None, epilogue=0, out=None, **kw):
        if out is not None and out.shape == (n_loc, gc):
            v_outs.append((tuple(weight.shape), out.storage_offset(), tuple(out.stride())))
        return orig_linear(x, weight, bias, epilogue, out=out, **kw)

    rec.linear = linear
    rec.rmsnorm_scatter = lambda x, w, dst, row0, eps=1e-6: scat.append((tuple(x.shape), len(dst), tuple(dst[0].shape), row0))
    rec.attention_scatter = lambda q, k, v, outs, k_lens=None: att.append(
        (tuple(q.shape), tuple(k.shape), [(tuple(o.shape), tuple(o.stride()), o.storage_offset()) for o in outs]))
    x = torch.empty(B, n_loc, C, device="meta", dtype=BF16)
    cos = sin = torch.empty(1024, 64, device="meta")
    out = attn.attend(x, torch.empty(B, device="meta", dtype=torch.int32),
                      torch.empty(B, 3, device="meta", dtype=torch.int32), cos, sin, SP())
    assert out is pb.o and pb.barriers == 2
    # v: P destinations x B batches, each an [n_loc, C/P] GEMM into rows r*n_loc.. of slot 2
    assert len(v_outs) == P * B
    per_b = L * gc
    want = sorted(2 * B * per_b + b * per_b + r * n_loc * gc for b in range(B))
    assert sorted(o for _, o, _ in v_outs[:B]) == want and all(w == (gc, C) and s == (gc, 1) for w, _, s in v_outs)
    # q, k: normalised locally, scattered by head group at row offset r*n_loc
    assert scat == [((B, n_loc, C), P, (B, L, gc), r * n_loc)] * 2
    # attention: all L queries of my h heads; chunk s goes to rank s at my head columns
    (qs, ks, outs), = att
    assert qs == (B, L, h, d) and ks == qs
    assert outs == [((B, n_loc, h, d), (n_loc * C, C, d, 1), r * h * d)] * P
    assert rec.calls["rmsnorm_rope_+rope"] == 2                      # RoPE after the exchange, on my heads


def test_cross_attention_key_buffer_adjacency():
    """WanI2VCrossAttention keeps text keys then image keys in ONE buffer and hands out views; `attend` runs
    both softmaxes in one launch (ops.attention_seg2) only when the views really are adjacent slices of one
    [B, Lt + Li, C] buffer with Lt a multiple of the 128-key tile — also after the cfg_skip batch slicing —
    and falls back to two launches for anything else."""
    adj = dit_mod.WanI2VCrossAttention._adjacent
    kc = torch.randn(2, 512 + 257, 256).to(BF16)
    k, ki = kc[:, :512], kc[:, 512:]
    assert adj(k, ki)
    assert adj(k[1:], ki[1:])                                    # conditional half of a CFG batch (Conditioning.tail)
    assert not adj(k.contiguous(), ki.contiguous())              # separate tensors
    assert not adj(ki, k)                                        # wrong order
    k77, ki77 = torch.randn(2, 77 + 257, 256).to(BF16).split([77, 257], dim=1)
    assert not adj(k77, ki77)                                    # first segment not a multiple of 128 keys
    other = torch.randn(2, 512 + 257, 256).to(BF16)
    assert not adj(k, other[:, 512:])                            # views of different buffers
