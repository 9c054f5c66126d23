"""Ulysses sequence parallelism on real GPUs: parity against the unsharded forward, then timing.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/sp_check.py [--layers 4]"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from more4d_b200 import config as mcfg, synth          # noqa: E402
from more4d_b200.dit import WanTransformer4DModel      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=4)
    a = ap.parse_args()
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.set_grad_enabled(False)
    out = {"world": world}

    # ---- parity on a small model whose head count divides the world size
    cfg = mcfg.WAN_TINY.with_(num_heads=4, dim=512, ffn_dim=1024)
    grid, seed = (3, 5, 6), 21
    model = WanTransformer4DModel.from_config(cfg, device=dev)
    model.load_state_dict(synth.dit_state_dict(cfg, seed), strict=True)
    inp = synth.dit_inputs(cfg, grid, 2, seed)
    kw = dict(x=inp["x"].to(dev), t=inp["t"].to(dev), context=[c.to(dev) for c in inp["context"]],
              seq_len=inp["seq_len"], clip_fea=inp["clip_fea"].to(dev), y=inp["y"].to(dev),
              full_ref=inp["full_ref"].to(dev))
    y_ref = model(**kw)
    for mode, peer in (("nccl", False), ("peer", True)):
        model.enable_multi_gpus_inference()
        model.sp.peer_memory = peer
        y_sp = model(**kw)
        y_sp2 = model(**kw)                                   # buffers re-used across calls
        model.disable_multi_gpus_inference()
        out[f"parity_bit_exact_{mode}"] = bool(torch.equal(y_ref, y_sp) and torch.equal(y_ref, y_sp2))
        out[f"parity_rel_{mode}"] = float((y_ref.float() - y_sp.float()).norm() / y_ref.float().norm())
    del model

    # ---- timing: 720p, 14B dims, `layers` blocks, CFG batch 2: unsharded vs sharded forward
    cfg = mcfg.WAN_14B.with_(num_layers=a.layers)
    grid = mcfg.token_grid(49, 720, 1280, cfg)
    model = WanTransformer4DModel.from_config(cfg, device=dev)
    synth.fill_module_(model, cfg, seed=0)
    inp = synth.dit_inputs(cfg, grid, 2, 1, device=dev)
    kw = dict(x=inp["x"], t=inp["t"], context=inp["context"], seq_len=inp["seq_len"], clip_fea=inp["clip_fea"],
              y=inp["y"], full_ref=inp["full_ref"])

    def timed(n=2):
        model(**kw)
        dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        for _ in range(n):
            y = model(**kw)
        e.record()
        dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e) / n], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), y

    t1, y1 = timed()
    model.enable_multi_gpus_inference()
    model.sp.peer_memory = False
    t2, y2 = timed()
    model.sp.peer_memory = True
    t3, y3 = timed()
    out.update(layers=a.layers, ms_forward_unsharded=t1, ms_forward_sp_nccl=t2, ms_forward_sp_peer=t3,
               speedup_nccl=t1 / t2, speedup_peer=t1 / t3,
               fullsize_bit_exact_nccl=bool(torch.equal(y1, y2)), fullsize_bit_exact_peer=bool(torch.equal(y1, y3)),
               fullsize_rel_peer=float((y1.float() - y3.float()).norm() / y1.float().norm()),
               fullsize_mismatch_frac_peer=float((y1 != y3).float().mean()))
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
