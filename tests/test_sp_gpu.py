"""Sequence-parallel (Ulysses) forward on TWO real GPUs: both exchange forms — NCCL all-to-all and
the fused peer-memory stores — must reproduce the unsharded forward bit for bit.  Skipped on a
single-GPU box (tools/sp_check.py runs the same check under torchrun, with timing)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from more4d_b200 import config as mcfg, synth
    from more4d_b200.dit import WanTransformer4DModel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        torch.set_grad_enabled(False)
        cfg = mcfg.WAN_TINY.with_(num_heads=4, dim=512, ffn_dim=1024)
        model = WanTransformer4DModel.from_config(cfg, device=dev)
        model.load_state_dict(synth.dit_state_dict(cfg, 21), strict=True)
        inp = synth.dit_inputs(cfg, (3, 5, 6), 2, 21)
        kw = dict(x=inp["x"].to(dev), t=inp["t"].to(dev), context=[c.to(dev) for c in inp["context"]],
                  seq_len=inp["seq_len"], clip_fea=inp["clip_fea"].to(dev), y=inp["y"].to(dev),
                  full_ref=inp["full_ref"].to(dev))
        ref = model(**kw)
        res = {}
        for mode, peer in (("nccl", False), ("peer", True)):
            model.enable_multi_gpus_inference()
            model.sp.peer_memory = peer
            y1, y2 = model(**kw), model(**kw)
            model.disable_multi_gpus_inference()
            res[mode] = bool(torch.equal(ref, y1) and torch.equal(ref, y2))
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sequence_parallel_forward_is_bit_identical():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in out:
        assert res == {"nccl": True, "peer": True}, (rank, res)
