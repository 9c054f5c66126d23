"""Motion-Sensitive VAE round trip at BASELINE.json configs[3]: 49x720x1280 trajectory tensor
(adaptor -> encode -> decode -> adaptor) on one B200, CUDA-event timed per stage.

    python tools/bench_vae.py [--frames 49 --height 720 --width 1280 --iters 2]

FLOP numerators are the conv FLOPs of SURVEY.md §8a/BASELINE.md §3 scaled to the chosen size.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from more4d_b200 import ops, synth                                            # noqa: E402
from more4d_b200.vae import AutoencoderKLWan, VAEDecoderadaptor, VAEEncoderadaptor  # noqa: E402

# conv FLOPs at 49x720x1280
FLOPS_FULL = {"enc_adaptor": 2.73e13, "encode": 2.27e14, "decode": 3.84e14, "dec_adaptor": 5.39e13}


def run(frames=49, height=720, width=1280, iters=2, dev="cuda", e2e=True):
    """Round trip on `dev`; returns the record bench.py embeds as `vae_roundtrip`.  Device-resident
    stage times (CUDA events, best of `iters` after one warm-up pass) and, with e2e, one pass
    through the public modules from a PINNED HOST tensor to a host tensor (H2D of the trajectory
    video and D2H of the reconstruction inside the timed region)."""
    import time
    torch.set_grad_enabled(False)
    vae = AutoencoderKLWan(device=dev)
    vae.load_state_dict(synth.vae_state_dict(seed=0, device=dev), strict=True)
    ea, da = VAEEncoderadaptor(device=dev), VAEDecoderadaptor(device=dev)
    ea.load_state_dict(synth.adaptor_state_dict("encoder", 0, device=dev), strict=True)
    da.load_state_dict(synth.adaptor_state_dict("decoder", 0, device=dev), strict=True)
    x_host = synth.trajectory_video(frames, height, width, 0).pin_memory()
    x = x_host.to(dev)
    scale = frames * height * width / (49 * 720 * 1280)

    def stage(fn, *args):
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        out = fn(*args)
        e.record()
        torch.cuda.synchronize()
        return out, s.elapsed_time(e)

    res, launches, finite, lat = {}, 0, True, None
    torch.cuda.reset_peak_memory_stats()
    for it in range(iters + 1):
        l0 = ops.launches()
        pseudo, t0 = stage(ea, x)
        lat, t1 = stage(lambda p: vae.encode_scaled(p, 2.0, -1.0).latent_dist.mode(), pseudo)
        del pseudo
        rec, t2 = stage(lambda z: vae.decode(z).sample, lat)
        out, t3 = stage(da, rec)
        launches = ops.launches() - l0
        if it > 0:           # first pass = warm-up (weight packing, allocator growth)
            for k, t in zip(FLOPS_FULL, (t0, t1, t2, t3)):
                res.setdefault(k, []).append(t)
        finite = bool(torch.isfinite(out.float()).all())
        del rec, out
    line = {"workload": f"Motion-Sensitive VAE round trip {frames}x{height}x{width}, bf16, 1xB200 "
                        "(VAEEncoderadaptor -> *2-1 -> encode -> mode -> decode -> VAEDecoderadaptor)",
            "latent_shape": list(lat.shape), "finite": finite, "kernel_launches_per_roundtrip": launches,
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30, "stages": {}}
    total = 0.0
    for k, ts in res.items():
        ms = min(ts)
        total += ms
        line["stages"][k] = {"ms": ms, "conv_tflops": FLOPS_FULL[k] * scale / ms / 1e9}
    line["total_ms"] = total
    line["roundtrips_per_s"] = 1000.0 / total if total else None
    line["conv_flops"] = sum(FLOPS_FULL.values()) * scale
    line["conv_tflops_total"] = sum(FLOPS_FULL.values()) * scale / total / 1e9 if total else None
    if e2e:
        from more4d_b200.vae import motion_vae_roundtrip
        out_host = torch.empty(x_host.shape, dtype=x_host.dtype).pin_memory()     # the caller's result buffer
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out, _, _ = motion_vae_roundtrip(x_host.to(dev, non_blocking=True), vae, ea, da)
        out_host.copy_(out.reshape(out_host.shape), non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        line["e2e"] = {"ms": dt * 1e3, "roundtrips_per_s": 1.0 / dt, "h2d_bytes": x_host.numel() * 2,
                       "d2h_bytes": out_host.numel() * 2}
    del vae, ea, da, x
    torch.cuda.empty_cache()
    return line


def cpu_sample(threads=None, latent_frames=1, height=720, width=1280):
    """CPU baseline of the VAE ("ONE VAE chunk at full resolution"): the oracle port
    (fp32) decoding `latent_frames` latent frames (= 1 + 4*(n-1) video frames) at full resolution on
    the host cores; extrapolated to the whole round trip by conv FLOPs (stated in the record)."""
    import time
    from oracle import vae_oracle as V
    if threads:
        torch.set_num_threads(threads)
    sd = {k: v.float() for k, v in synth.vae_state_dict(seed=0).items()}
    g = torch.Generator().manual_seed(0)
    z = torch.randn(1, 16, latent_frames, height // 8, width // 8, generator=g)
    t0 = time.perf_counter()
    V.decode(z, sd)
    dt = time.perf_counter() - t0
    frames = 1 + 4 * (latent_frames - 1)
    flops = FLOPS_FULL["decode"] * frames / 49 * (height * width) / (720 * 1280)
    total = sum(FLOPS_FULL.values())
    est = dt * total / flops
    return {"value": 1.0 / est, "unit": "round trips/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"oracle port (fp32) decode of {latent_frames} latent frames -> {frames} frames at "
                      f"{height}x{width}: {dt:.1f} s for {flops:.2e} conv FLOP; extrapolated by conv FLOPs to the "
                      f"whole 49-frame round trip ({total:.2e}): x{total / flops:.1f}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=49)
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--cpu", action="store_true", help="also time the CPU port sample")
    a = ap.parse_args()
    line = run(a.frames, a.height, a.width, a.iters)
    if a.cpu:
        line["cpu_baseline"] = cpu_sample()
    print(json.dumps(line))


if __name__ == "__main__":
    main()
