"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sample sharding + final all-gather
must reproduce the 1-rank result bitwise; CFG sharding must hand both branches to both ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from more4d_b200 import dist as mdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _step(lat, s, cond):            # deterministic stand-in for StraGDenoiser.step
    return (lat * 0.9 + cond * (s + 1)).to(lat.dtype)


def _inputs(n):
    g = torch.Generator().manual_seed(0)
    lats = [torch.randn(1, 4, 3, 5, generator=g).to(torch.bfloat16) for _ in range(n)]
    conds = [torch.randn(1, 4, 3, 5, generator=g).to(torch.bfloat16) for _ in range(n)]
    return lats, conds


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lats, conds = _inputs(n)
        out = mdist.denoise_sharded(_step, lats, conds, steps=[0, 1, 2])
        u, t = mdist.cfg_split_noise(lambda br: torch.full((2, 3), float(br + 1)))
        # plain numpy through the queue: torch tensors travel as file descriptors served by the
        # child, which may have exited before the parent reads them
        q.put((rank, [o.float().numpy() for o in out], u.numpy(), t.numpy(), mdist.shard_indices(n)))
    finally:
        dist.destroy_process_group()


def _run(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res, key=lambda r: r[0])


def test_sample_sharding_matches_single_rank():
    for n in (4, 5):                                    # even and ragged shard sizes
        lats, conds = _inputs(n)
        ref = mdist.denoise_sharded(_step, lats, conds, steps=[0, 1, 2])     # world = 1 path
        res = _run(n)
        assert res[0][4] == list(range(0, n, 2)) and res[1][4] == list(range(1, n, 2))
        for rank, out, u, t, _ in res:
            assert len(out) == n
            for a, b in zip(out, ref):
                assert torch.equal(torch.from_numpy(a), b.float())
            assert torch.equal(torch.from_numpy(u), torch.full((2, 3), 1.0))
            assert torch.equal(torch.from_numpy(t), torch.full((2, 3), 2.0))


def _sp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sp = mdist.SequenceParallel()
        g = torch.Generator().manual_seed(0)
        S, B, L, H, D = 3, 2, 12, 4, 5
        full = torch.randn(S, B, L, H, D, generator=g)                     # identical on every rank
        n = L // world
        mine = full[:, :, rank * n:(rank + 1) * n].contiguous()            # token shard, all heads
        heads = sp.seq_to_heads(mine)                                      # all tokens, my heads
        h = H // world
        ok1 = torch.equal(heads, full[:, :, :, rank * h:(rank + 1) * h])
        back = sp.heads_to_seq(heads[0])                                   # [B, n, H, D]
        ok2 = torch.equal(back, mine[0])
        tok = sp.shard_tokens(full[0].reshape(B, L, H * D))
        ok3 = torch.equal(tok, mine[0].reshape(B, n, H * D))
        ok4 = torch.equal(sp.gather_tokens(tok), full[0].reshape(B, L, H * D))
        q.put((rank, bool(ok1), bool(ok2), bool(ok3), bool(ok4)))
    finally:
        dist.destroy_process_group()


def test_sequence_parallel_layout_roundtrip():
    """Ulysses all-to-alls (world 2, gloo): token shard x all heads -> all tokens x head shard is
    exactly the corresponding slice of the full tensor, and the inverse restores the shard."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert r[1:] == (True, True, True, True), r
