#!/usr/bin/env python
"""Benchmark of BASELINE.json's metric: 4D-STraG denoise-step latents/sec at 49x720x1280.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One *step* = one latent-step of the reference's denoising loop (pipeline_wan_fun_control.py:
741-840): one DiT forward on the CFG-doubled batch + CFG combine + Euler update, on synthetic
latents / conditioning and seeded random weights of the Wan2.1-14B-Control architecture
(no checkpoints or datasets are reachable).  N > 1: one process per GPU (torchrun), every rank
denoises its own sample (the batch axis of the north-star; weak scaling), no data-path
collective inside a step, one NCCL all-gather of the final latents inside the timed region.

JSON keys follow the driver contract; see DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (config preset, frames, height, width)
    "720p-14b": ("WAN_14B", 49, 720, 1280),
    "480p-14b": ("WAN_14B", 49, 480, 832),
    "720p-1.3b": ("WAN_1_3B", 49, 720, 1280),
    "480p-1.3b": ("WAN_1_3B", 49, 480, 832),
    "tiny": ("WAN_TINY", 9, 64, 96),
    # 4D-ViSM (Wan-InP backbone) denoise part of BASELINE config 5 — NOT the headline metric
    "visim-368p-14b": ("WAN_14B_INP", 49, 368, 512),
}
METRIC = "4D-STraG denoise-step latents/sec @49x720p"
UNIT = "latent-steps/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p.get("bf16_tflops_sustained", 1424.9), p.get("bf16_tflops", 1679.0), "measured"
    return 1400.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(gpu_index)], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for line in self.f:
            c = [v.strip() for v in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            try:
                pw.append(float(c[3]))
            except ValueError:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        busy = sorted(sm)[len(sm) // 4:] if sm else []
        return {"sm_mhz": statistics.median(busy) if busy else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w": statistics.median(pw) if pw else None}


def cpu_reference_arm(args, cfg, L, grid, rank, repeats=1, skip=0):
    """`--impl reference` and the cpu_baseline leg: the oracle port of the block on host cores."""
    import torch
    from oracle import cpu_baseline
    rows = args.cpu_sample_rows
    # all host cores this process may use — torchrun exports OMP_NUM_THREADS=1, which would
    # otherwise time the reference arm on a single thread
    try:
        host_threads = len(os.sched_getaffinity(0))
    except AttributeError:
        host_threads = os.cpu_count() or 1
    times, threads = cpu_baseline.time_block_sample(cfg, L, rows, grid, seed=0, repeats=max(1, repeats),
                                                    threads=host_threads)
    times = times[skip:] if len(times) > skip else times
    t_best = min(times)
    value = cpu_baseline.latent_steps_per_s(t_best, L, rows, cfg.num_layers)
    sample = (f"1 of {cfg.num_layers} blocks, {rows} of {L} query rows vs all {L} keys, 1 of 2 CFG "
              f"branches, fp32 oracle port; extrapolated linearly to a full latent-step "
              f"(x{L / rows:.1f} x{cfg.num_layers} x2); best of {len(times)}: {t_best:.2f} s")
    return {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}, times


def sequence_parallel_legs(args, cfg, model, den, dev, rank, world, frames, height, width, dp_ms_per_step):
    """N > 1 only.  (1) bit-exactness of the Ulysses forward — both exchange forms, NCCL all-to-all and
    the fused peer-memory stores — against the unsharded forward on a small model whose head count
    divides N; (2) ONE sample of the headline workload sequence-sharded over the N GPUs (strong
    scaling, single-sample latency): ms per latent-step, device-timed, max over ranks."""
    import torch
    import torch.distributed as dist
    from more4d_b200 import config as mcfg, synth
    from more4d_b200.dit import WanTransformer4DModel
    from more4d_b200.pipeline import synthetic_conditioning
    rec = {}
    tiny = mcfg.WAN_TINY.with_(num_heads=8, dim=1024, ffn_dim=2048)
    m = WanTransformer4DModel.from_config(tiny, device=dev)
    m.load_state_dict(synth.dit_state_dict(tiny, 21), strict=True)
    inp = synth.dit_inputs(tiny, (3, 5, 8), 2, 21)
    kw = dict(x=inp["x"].to(dev), t=inp["t"].to(dev), context=[c.to(dev) for c in inp["context"]],
              seq_len=inp["seq_len"], clip_fea=inp["clip_fea"].to(dev), y=inp["y"].to(dev),
              full_ref=inp["full_ref"].to(dev))
    ref = m(**kw)
    flags = []
    for mode, peer in (("nccl", False), ("peer", True)):
        m.enable_multi_gpus_inference()
        m.sp.peer_memory = peer
        y1, y2 = m(**kw), m(**kw)                          # second call re-uses the exchange buffers
        m.disable_multi_gpus_inference()
        ok = torch.tensor([int(torch.equal(ref, y1) and torch.equal(ref, y2))], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        rec[f"bit_exact_{mode}"] = bool(ok.item())
        flags.append(bool(ok.item()))
    rec["bit_exact"] = all(flags)
    rec["bit_exact_model"] = "tiny DiT (C 1024, 8 heads, 2 layers, grid 3x5x8, CFG batch 2), both exchange forms, on every rank"
    del m
    # (2) headline workload, one sample over all ranks
    lat_t = (frames - 1) // 4 + 1
    lat_host, cond_host = synthetic_conditioning((1, 16, lat_t, height // 8, width // 8), seed=0, device="cpu",
                                                 text_dim=cfg.text_dim, clip_dim=cfg.clip_dim)
    lat, cond = lat_host.to(dev), cond_host.to(dev)
    model.enable_multi_gpus_inference()
    for s in range(2):
        den.step(lat, s, cond)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.sp_steps):
        den.step(lat, (2 + s) % den.num_inference_steps, cond)
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.sp_steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    model.disable_multi_gpus_inference()
    rec.update(ms_per_step=float(ms.item()), steps=args.sp_steps, scaling="strong",
               latent_steps_per_s=1000.0 / float(ms.item()),
               speedup_vs_one_gpu_step=dp_ms_per_step / float(ms.item()),
               exchange="fused peer-memory stores (symmetric memory) + barrier; 2 exchanges per block",
               note="one sample sequence-sharded over all ranks; `speedup` is against this run's per-GPU "
                    "sample-sharded step time (= the 1-GPU step time under weak scaling)")
    return rec


_JSON_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout: keep the real stdout for it and point fd 1 at stderr, so
    that banners printed by C libraries (NCCL's version line under torchrun) cannot land in front of it."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="720p-14b", choices=sorted(WORKLOADS))
    ap.add_argument("--layers", type=int, default=0, help="debug only: truncate the block stack (number is then INVALID)")
    ap.add_argument("--cpu-sample-rows", type=int, default=2048)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parallelism", default="sample", choices=["sample", "sp"],
                    help="sample: one sample per GPU (weak scaling, the default / headline); sp: ONE sample, "
                         "sequence sharded over the GPUs (Ulysses all-to-alls; strong scaling, single-sample latency)")
    ap.add_argument("--no-vae", action="store_true", help="skip the Motion-Sensitive VAE round-trip sub-record (N=1)")
    ap.add_argument("--no-sp", action="store_true", help="skip the sequence-parallel legs (N>1)")
    ap.add_argument("--sp-steps", type=int, default=10)
    ap.add_argument("--hoist-conditioning", action="store_true",
                    help="compute the step-invariant context embedding / cross-attention K/V once instead of "
                         "per step (bit-identical; NOT the default: the reference recomputes them every step)")
    args = ap.parse_args()

    import torch
    from more4d_b200 import config as mcfg

    preset, frames, height, width = WORKLOADS[args.workload]
    cfg = getattr(mcfg, preset)
    if args.layers:
        cfg = cfg.with_(num_layers=args.layers)
    visim = args.workload.startswith("visim")
    grid = mcfg.token_grid(frames, height, width, cfg, with_ref=not visim)
    L = grid[0] * grid[1] * grid[2]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"{args.workload}: Wan2.1-14B-{'InP' if visim else 'Control'} dims (C{cfg.dim} F{cfg.ffn_dim} "
                          f"{cfg.num_heads}h x{cfg.num_layers}L) {frames}x{height}x{width}, "
                          f"L={L} tokens, CFG batch 2, 1 sample/GPU",
              "parallelism": (f"sp{world} (one sample, Ulysses sequence sharding: 2 all-to-alls per block)"
                              if args.parallelism == "sp" and world > 1 else
                              f"dp{world} (sample-sharded replicas, all-gather of final latents)"),
              "l2": "inputs (28 GB weights + GB-scale activations) exceed the 126 MB L2 every step",
              "layers_override": args.layers or None,
              "hoist_conditioning": bool(args.hoist_conditioning)}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        torch.set_grad_enabled(False)
        cb, times = cpu_reference_arm(args, cfg, L, grid, rank, repeats=args.warmup + args.steps,
                                      skip=args.warmup)
        t_mean = sum(times) / len(times)
        from oracle import cpu_baseline
        value = cpu_baseline.latent_steps_per_s(t_mean, L, args.cpu_sample_rows, cfg.num_layers)
        cb["value"] = value
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / value,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32",
                "data": "synthetic", "config": config, "cpu_baseline": cb,
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        _emit(line)
        return

    # ------------------------------------------------------------------ our arm (B200)
    import torch.distributed as dist
    from more4d_b200 import dist as mdist, ops, synth
    from more4d_b200.dit import WanTransformer3DModel, WanTransformer4DModel
    from more4d_b200.pipeline import (StraGDenoiser, ViSMDenoiser, synthetic_conditioning,
                                      synthetic_visim_conditioning)

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.set_grad_enabled(False)

    model = (WanTransformer3DModel if visim else WanTransformer4DModel).from_config(cfg, device=dev)
    synth.fill_module_(model, cfg, seed=0)
    den = (ViSMDenoiser if visim else StraGDenoiser)(model, guidance_scale=6.0, shift=5.0,
                                                     num_inference_steps=50,
                                                     hoist_conditioning=args.hoist_conditioning)
    sp_mode = args.parallelism == "sp" and world > 1
    if sp_mode:
        model.enable_multi_gpus_inference()
    data_seed = 0 if sp_mode else rank                       # sp: every rank works on the SAME sample
    lat_t = (frames - 1) // 4 + 1
    latent_shape = (1, 16, lat_t, height // 8, width // 8)
    if visim:
        lat_host, cond_host = synthetic_visim_conditioning(latent_shape, seed=data_seed, device="cpu",
                                                           text_dim=cfg.text_dim, clip_dim=cfg.clip_dim)
        lat_host = lat_host.pin_memory()
    else:
        lat_host, cond_host = synthetic_conditioning(latent_shape, seed=data_seed, device="cpu", pin=True,
                                                     text_dim=cfg.text_dim, clip_dim=cfg.clip_dim)
    lat = lat_host.to(dev)
    cond = cond_host.to(dev)
    gathered = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(steps, latents, first_step=0):
        for s in range(steps):
            den.step(latents, (first_step + s) % den.num_inference_steps, cond)

    # warm-up (also JITs nothing: all kernels are prebuilt; this pages weights and warms clocks)
    run(args.warmup, lat)
    barrier()

    # ---- device-resident timing ("value") with per-launch timing of the dominant kernel
    sampler = ClockSampler(local_rank)
    ops.start_kernel_timing()
    l0 = ops.launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.nvtx.range_push("m4d_timed")          # ncu --nvtx --nvtx-include "m4d_timed/"
    ev0.record()
    run(args.steps, lat, args.warmup)
    if world > 1 and not sp_mode:
        gathered = mdist.gather_latents([lat], world)     # the north-star's one collective
    ev1.record()
    barrier()
    torch.cuda.nvtx.range_pop()
    clocks = sampler.stop()
    timed = ops.stop_kernel_timing()
    launches = ops.launches() - l0
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    units = 1 if sp_mode else world                          # samples advanced per step by the whole job
    value = units * args.steps / (ms_total / 1000.0)

    # ---- end to end through the public API with HOST buffers (pinned H2D in, D2H out, every step)
    h2d = lat_host.numel() * 2 + cond_host.nbytes()
    d2h = lat_host.numel() * 2
    barrier()
    t0 = time.perf_counter()
    cur = lat_host
    for s in range(args.steps):
        cur = den(cur, cond_host, steps=[(args.warmup + s) % den.num_inference_steps], device=dev)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = units * args.steps / float(e2e_s.item())

    # ---- per-kernel-class device time: ONE extra step outside the timed region with every launch of the
    # block kernels bracketed by events (name what grows with N); reported for the slowest rank
    ops.start_kernel_timing(("attention", "cross_attention", "gemm", "rows"))
    barrier()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    run(1, lat, args.warmup + args.steps)
    pe1.record()
    torch.cuda.synchronize()
    prof = ops.stop_kernel_timing()
    cls_ms = {k: sum(a.elapsed_time(b) for a, b, _ in v) for k, v in prof.items()}
    cls_ms["step"] = pe0.elapsed_time(pe1)
    cls_ms["other"] = cls_ms["step"] - sum(v for k, v in cls_ms.items() if k != "step")
    keys = sorted(cls_ms)
    cm = torch.tensor([cls_ms[k] for k in keys], device=dev)
    per_rank = None
    if world > 1:
        # every rank's figures (identical work on every rank under weak scaling): the spread between
        # chips is what separates the N-GPU step time (max over ranks) from the 1-GPU one
        mine = torch.tensor([cls_ms[k] for k in keys] + [ev0.elapsed_time(ev1) / args.steps,
                                                         float(clocks.get("sm_mhz") or 0.0),
                                                         float(clocks.get("power_w") or 0.0)], device=dev)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        allr = torch.stack(allr).cpu()
        per_rank = {k: [round(float(v), 1) for v in allr[:, i]] for i, k in enumerate(keys)}
        per_rank["timed_ms_per_step"] = [round(float(v), 1) for v in allr[:, len(keys)]]
        per_rank["sm_mhz"] = [float(v) for v in allr[:, len(keys) + 1]]
        per_rank["power_w"] = [round(float(v)) for v in allr[:, len(keys) + 2]]
        dist.all_reduce(cm, op=dist.ReduceOp.MAX)
    kernel_class_ms = {k: float(v) for k, v in zip(keys, cm.tolist())}
    if per_rank is not None:
        kernel_class_ms["per_rank"] = per_rank
    kernel_class_ms["gemm_tflops"] = sum(w for _, _, w in prof["gemm"]) / (cls_ms["gemm"] / 1e3) / 1e12 if cls_ms["gemm"] else None
    kernel_class_ms["note"] = ("one step with per-launch CUDA events on every block kernel (max over ranks per class); "
                               "`other` = embeddings, head, CFG/Euler, host gaps")

    # ---- N > 1: the sequence-parallel path under the driver (VERDICT r1 #1d, #5)
    sp_record = None
    if world > 1 and not args.no_sp and cfg.num_heads % world == 0 and not visim:
        sp_record = sequence_parallel_legs(args, cfg, model, den, dev, rank, world, frames, height, width,
                                           ms_total / args.steps)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    sust, burst, src = _peaks()
    att = timed["attention"]
    att_ms = [a.elapsed_time(b) for a, b, _ in att]
    att_fl = att[0][2] if att else 0.0
    avg_ms = sum(att_ms) / max(1, len(att_ms))
    achieved = att_fl / (avg_ms / 1e3) / 1e12 if att else 0.0
    # DRAM bytes per launch of the SHIPPED kernel at this workload's shape, from its committed ncu
    # summary (profiles/attention_traffic.json names the capture it was read from)
    traffic, traffic_source = None, None
    tpath = os.path.join(ROOT, "profiles", "attention_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_source = tj.get(args.workload), tj.get("source")
    roofline = {"kernel": "attn_fwd_d128_kernel (self-attention launches, Lq=Lk=L)", "bound": "tensor",
                "achieved": achieved, "peak": sust, "unit": "TFLOP/s", "frac": achieved / sust,
                "frac_of_burst_peak": achieved / burst, "peak_source": f"{src} bf16 sustained (kernel timed inside a long step)",
                "launches_timed": len(att), "avg_launch_ms": avg_ms,
                "share_of_step": sum(att_ms) / ms_total if ms_total else None,
                "algorithmic_flops_per_launch": att_fl, "traffic": traffic, "traffic_source": traffic_source}

    cb = None
    if world == 1 and not args.no_cpu_baseline:
        cb, _ = cpu_reference_arm(args, cfg, L, grid, rank)

    # ---- BASELINE.json configs[3]: Motion-Sensitive VAE encode+decode at 49x720x1280 (N=1 only)
    vae_record = None
    if world == 1 and not args.no_vae and args.workload.startswith("720p"):
        del model, den
        torch.cuda.empty_cache()
        from tools import bench_vae
        vae_record = bench_vae.run(frames, height, width, iters=2, dev=dev)
        vae_record["frac_of_sustained_tensor_peak"] = vae_record["conv_tflops_total"] / sust
        vae_record["roofline_note"] = ("conv FLOPs 6.92e14 (BASELINE.md §3) / total_ms against the measured sustained "
                                       "bf16 peak; the convolutions are tensor-bound, norms / resampling HBM-bound")
        if not args.no_cpu_baseline:
            vae_record["cpu_baseline"] = bench_vae.cpu_sample(threads=len(os.sched_getaffinity(0)))

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong" if sp_mode else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "dit_forwards_per_s": value * 2,
            "roofline": roofline, "cpu_baseline": cb, "kernel_class_ms": kernel_class_ms,
            "vae_roundtrip": vae_record, "sp": sp_record,
            "sp_bit_exact": None if sp_record is None else sp_record.get("bit_exact")}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
