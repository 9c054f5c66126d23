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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=49)
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--iters", type=int, default=2)
    a = ap.parse_args()
    torch.set_grad_enabled(False)
    dev = "cuda"
    vae = AutoencoderKLWan(device=dev)
    vae.load_state_dict(synth.vae_state_dict(seed=0, device=dev), strict=True)
    ea, da = VAEEncoderadaptor(device=dev), VAEDecoderadaptor(device=dev)
    ea.load_state_dict(synth.adaptor_state_dict("encoder", 0, device=dev), strict=True)
    da.load_state_dict(synth.adaptor_state_dict("decoder", 0, device=dev), strict=True)
    x = synth.trajectory_video(a.frames, a.height, a.width, 0).to(dev)
    scale = a.frames * a.height * a.width / (49 * 720 * 1280)

    def stage(fn, *args):
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        out = fn(*args)
        e.record()
        torch.cuda.synchronize()
        return out, s.elapsed_time(e)

    res = {}
    for it in range(a.iters + 1):
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
    line = {"workload": f"Motion-Sensitive VAE round trip {a.frames}x{a.height}x{a.width}, bf16, 1xB200",
            "latent_shape": list(lat.shape), "finite": finite, "kernel_launches_per_roundtrip": launches,
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30, "stages": {}}
    total = 0.0
    if not res:
        print(json.dumps(line))
        return
    for k, ts in res.items():
        ms = min(ts)
        total += ms
        line["stages"][k] = {"ms": ms, "conv_tflops": FLOPS_FULL[k] * scale / ms / 1e9}
    line["total_ms"] = total
    line["conv_tflops_total"] = sum(FLOPS_FULL.values()) * scale / total / 1e9
    print(json.dumps(line))


if __name__ == "__main__":
    main()
