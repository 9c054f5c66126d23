"""CPU baseline for bench.py: the oracle port of one DiT block timed on the host cores on a
BOUNDED sample of the benchmark workload.

TEST/BENCH INFRASTRUCTURE — only bench.py's ``cpu_baseline`` / ``--impl reference`` legs call
this; it is never the product path.

A full latent-step at 49x720x1280 on the 14B config is 6.6e15 FLOP — hours on
a CPU — so the sample is: ONE of the `num_layers` blocks, ONE of the two CFG branches, and
`sample_rows` of the L query rows (attending to ALL L keys, so the quadratic attention term
keeps its true per-row cost).  Every op of the block is linear in the number of query rows
once K/V exist, and the K/V projections of the unsampled rows are covered by the same
per-row scaling, so

    t_latent_step ~= t_sample * (L / sample_rows) * num_layers * 2 (CFG)

The arithmetic is the oracle's (oracle/dit_oracle.py, fp32, i.e. the reference's own CPU
code path with its SDPA attention branch, wan_transformer4d.py:228-235).
"""
from __future__ import annotations

import time

import torch
import torch.nn.functional as F

from oracle import dit_oracle as O


@torch.no_grad()
def time_block_sample(cfg, L: int, sample_rows: int, grid, seed: int = 0, repeats: int = 1,
                      threads: int | None = None):
    """Returns ([seconds of each sampled block pass], threads used)."""
    if threads:
        torch.set_num_threads(threads)
    threads = torch.get_num_threads()
    from more4d_b200 import synth
    sd = {k: v.float() for k, v in synth.block_state_dict(cfg, 0, seed).items()}
    C, N, d = cfg.dim, cfg.num_heads, cfg.head_dim
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, sample_rows, C, generator=g)
    e0 = torch.randn(1, 6, C, generator=g) * 0.1
    ctx = torch.randn(1, 257 + cfg.text_len, C, generator=g)
    k_all = torch.randn(1, N, L, d, generator=g)          # keys/values of the full sequence
    v_all = torch.randn(1, N, L, d, generator=g)
    ar = O.Arith(False)
    p = "self_attn."

    def one_pass():
        e = (sd["modulation"] + e0).chunk(6, dim=1)
        t = O.layer_norm(x, None, None, cfg.eps) * (1 + e[1]) + e[0]
        q = O.rms_norm(ar.linear(t, sd[p + "q.weight"], sd[p + "q.bias"]), sd[p + "norm_q.weight"], cfg.eps, ar)
        k = O.rms_norm(ar.linear(t, sd[p + "k.weight"], sd[p + "k.bias"]), sd[p + "norm_k.weight"], cfg.eps, ar)
        v = ar.linear(t, sd[p + "v.weight"], sd[p + "v.bias"])
        # sampled rows are the first rows of the (F,H,W) token grid
        hw = grid[1] * grid[2]
        if sample_rows >= hw:
            sub = (sample_rows // hw, grid[1], grid[2])
        elif sample_rows >= grid[2]:
            sub = (1, sample_rows // grid[2], grid[2])
        else:
            sub = (1, 1, sample_rows)
        n_rope = sub[0] * sub[1] * sub[2]
        q = q.view(1, sample_rows, N, d)
        k = k.view(1, sample_rows, N, d)
        q = torch.cat([O.rope_apply(q[:, :n_rope], [sub], ar), q[:, n_rope:]], dim=1)
        k = torch.cat([O.rope_apply(k[:, :n_rope], [sub], ar), k[:, n_rope:]], dim=1)
        k_all[:, :, :sample_rows] = k.transpose(1, 2)
        v_all[:, :, :sample_rows] = v.view(1, sample_rows, N, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k_all, v_all)      # t4d:232
        y = ar.linear(o.transpose(1, 2).flatten(2), sd[p + "o.weight"], sd[p + "o.bias"])
        xx = x + y * e[2]
        n3 = O.layer_norm(xx, sd["norm3.weight"], sd["norm3.bias"], cfg.eps)
        xx = xx + O.cross_attention(n3, ctx, sd, "cross_attn.", N, cfg.eps, ar)
        t = O.layer_norm(xx, None, None, cfg.eps) * (1 + e[4]) + e[3]
        h = F.gelu(ar.linear(t, sd["ffn.0.weight"], sd["ffn.0.bias"]), approximate="tanh")
        return xx + ar.linear(h, sd["ffn.2.weight"], sd["ffn.2.bias"]) * e[5]

    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        one_pass()
        times.append(time.perf_counter() - t0)
    return times, threads


def latent_steps_per_s(t_sample: float, L: int, sample_rows: int, num_layers: int) -> float:
    return 1.0 / (t_sample * (L / sample_rows) * num_layers * 2)
