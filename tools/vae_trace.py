"""Per-op, per-shape CUDA-event trace of the Motion-Sensitive VAE round trip (development tool).

    python tools/vae_trace.py [--frames 49 --height 720 --width 1280] > gpurun_out/vae_trace.md

Every public function of `more4d_b200.ops` is wrapped with a pair of CUDA events on the launching
stream (no ncu: warm caches, real overlap of host and device); rows are grouped by (op, tensor
argument shapes).  The table is what DESIGN.md's VAE account and `profiles/vae_trace_r02.md` quote:
which layer shapes carry the round trip's time and at what tensor throughput.  Nested ops (an op
that calls another op) are counted in both rows; the stage totals come from events around the
stage, not from the sum.
"""
import argparse
import collections
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from more4d_b200 import ops, synth                                            # noqa: E402
from more4d_b200.vae import AutoencoderKLWan, VAEDecoderadaptor, VAEEncoderadaptor  # noqa: E402

RECORDS = []
STAGE = ["-"]


def _key(args, kwargs):
    parts = []
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, torch.Tensor) and a.dim() >= 2:
            parts.append("x".join(str(s) for s in a.shape))
        elif isinstance(a, (tuple, list)) and a and all(isinstance(i, int) for i in a):
            parts.append("(" + ",".join(str(i) for i in a) + ")")
    return " ".join(parts[:4])


def _wrap(name, fn):
    def inner(*args, **kwargs):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        out = fn(*args, **kwargs)
        e1.record()
        RECORDS.append((STAGE[0], name, _key(args, kwargs), e0, e1))
        return out
    inner.__name__ = name
    return inner


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=49)
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--width", type=int, default=1280)
    a = ap.parse_args()
    torch.set_grad_enabled(False)
    dev = "cuda"
    vae = AutoencoderKLWan(device=dev)
    vae.load_state_dict(synth.vae_state_dict(seed=0, device=dev), strict=True)
    ea, da = VAEEncoderadaptor(device=dev), VAEDecoderadaptor(device=dev)
    ea.load_state_dict(synth.adaptor_state_dict("encoder", 0, device=dev), strict=True)
    da.load_state_dict(synth.adaptor_state_dict("decoder", 0, device=dev), strict=True)
    x = synth.trajectory_video(a.frames, a.height, a.width, 0).to(dev)

    def roundtrip():
        STAGE[0] = "enc_adaptor"
        pseudo = ea(x)
        STAGE[0] = "encode"
        lat = vae.encode_scaled(pseudo, 2.0, -1.0).latent_dist.mode()
        del pseudo
        STAGE[0] = "decode"
        rec = vae.decode(lat).sample
        STAGE[0] = "dec_adaptor"
        out = da(rec)
        del rec, out

    roundtrip()                                          # warm-up: weight packing, allocator growth
    torch.cuda.synchronize()
    for name, fn in list(vars(ops).items()):
        if isinstance(fn, types.FunctionType) and not name.startswith("_") and fn.__module__ == ops.__name__ \
                and name not in ("launches", "start_kernel_timing", "stop_kernel_timing"):
            setattr(ops, name, _wrap(name, fn))
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    s.record()
    roundtrip()
    e.record()
    torch.cuda.synchronize()
    total = s.elapsed_time(e)
    agg = collections.OrderedDict()
    for stage, name, key, e0, e1 in RECORDS:
        k = (stage, name, key)
        n, ms = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, ms + e0.elapsed_time(e1))
    by_op = collections.Counter()
    for (stage, name, key), (n, ms) in agg.items():
        by_op[name] += ms
    print(f"# VAE round trip {a.frames}x{a.height}x{a.width}: per-op CUDA-event trace (tools/vae_trace.py)\n")
    print(f"round trip with per-op events: {total:.1f} ms (events add a few % of host gaps)\n")
    print("| op | total ms | share |\n|---|---:|---:|")
    for name, ms in by_op.most_common():
        print(f"| `{name}` | {ms:.1f} | {ms / total:.3f} |")
    print("\n| stage | op | tensor args | launches | total ms | ms / launch |\n|---|---|---|---:|---:|---:|")
    for (stage, name, key), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
        print(f"| {stage} | `{name}` | {key} | {n} | {ms:.2f} | {ms / n:.3f} |")


if __name__ == "__main__":
    main()
