"""The "beat this" number of SURVEY.md §8(d): the strongest LIBRARY attention kernels the reference
could dispatch to on this box (wan_transformer4d.py:138-169 flash-attn varlen, :232 SDPA) timed next
to `attn_fwd_d128_kernel` at the headline shape B=2, L=50 400, 40 heads, d=128 (bf16, non-causal).

    python tools/attn_library_bar.py [--L 50400] [--heads 40] [--sustain-s 4] > gpurun_out/attn_library_bar.json

For every candidate: a burst figure (best of 5 after 2 warm-ups, CUDA events) and a sustained figure
(back-to-back launches for `--sustain-s` seconds — the regime a kernel sees inside a long denoise
step under the 1 kW power cap), plus the nvidia-smi clocks during the sustained loop.  Development /
measurement tool: library kernels are NEVER on the product path.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from more4d_b200 import ops          # noqa: E402

BF16 = torch.bfloat16


class Clocks:
    def __init__(self):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,"
                                   "clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown",
                                   "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"], stdout=self.f,
                                  stderr=subprocess.DEVNULL)

    def stop(self):
        self.p.terminate()
        self.p.wait(timeout=5)
        self.f.flush()
        self.f.seek(0)
        sm, pw, cap, bad = [], [], 0, 0
        for line in self.f:
            c = [v.strip() for v in line.split(",")]
            try:
                sm.append(float(c[0])); pw.append(float(c[1]))
            except (ValueError, IndexError):
                continue
            cap += c[2].lower().startswith("active")
            bad += c[3].lower().startswith("active") or c[4].lower().startswith("active")
        self.f.close()
        os.unlink(self.f.name)
        sm.sort(); pw.sort()
        n = len(sm)
        return {"sm_mhz_median": sm[n // 2] if n else None, "power_w_median": pw[n // 2] if n else None,
                "sw_power_cap_samples": cap, "slowdown_samples": bad, "samples": n}


def timed(fn, flops, sustain_s):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    # sustained: events around every launch, back to back, no host sync in between
    n = max(3, int(sustain_s * 1000 / best))
    evs = [torch.cuda.Event(True) for _ in range(n + 1)]
    clk = Clocks()
    time.sleep(0.3)
    evs[0].record()
    for i in range(n):
        fn()
        evs[i + 1].record()
    torch.cuda.synchronize()
    clocks = clk.stop()
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(n)]
    tail = per[len(per) // 2:]                       # second half: clocks have settled
    sus = sum(tail) / len(tail)
    return {"burst_ms": best, "burst_tflops": flops / best / 1e9, "sustained_ms": sus,
            "sustained_tflops": flops / sus / 1e9, "launches": n, "clocks": clocks}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=50400)
    ap.add_argument("--heads", type=int, default=40)
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--sustain-s", type=float, default=4.0)
    args = ap.parse_args()
    B, L, N, D = args.B, args.L, args.heads, 128
    flops = 4.0 * B * N * L * L * D
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v = (torch.randn(B, L, N, D, device="cuda", dtype=BF16, generator=g) for _ in range(3))
    out = torch.empty_like(q)
    res = {"shape": {"B": B, "L": L, "heads": N, "d": D}, "flops_per_launch": flops,
           "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0), "kernels": {}}

    def record(name, fn, check=None):
        try:
            r = timed(fn, flops, args.sustain_s)
            if check is not None:
                o = check()
                r["rel_err_vs_ours"] = float((o.float() - out.float()).norm() / out.float().norm())
            res["kernels"][name] = r
        except Exception as e:                                  # a backend that refuses the shape is a result too
            res["kernels"][name] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
        print(name, json.dumps(res["kernels"][name]), file=sys.stderr, flush=True)

    record("more4d_b200.attn_fwd_d128_kernel", lambda: ops.attention(q, k, v, out=out))
    ours = out.clone()

    qt, kt, vt = (t.transpose(1, 2) for t in (q, k, v))            # [B, N, L, D] views, as t4d:228-235 builds them
    from torch.nn.attention import SDPBackend, sdpa_kernel
    for name, be in (("torch.sdpa.CUDNN_ATTENTION", SDPBackend.CUDNN_ATTENTION),
                     ("torch.sdpa.FLASH_ATTENTION", SDPBackend.FLASH_ATTENTION),
                     ("torch.sdpa.EFFICIENT_ATTENTION", SDPBackend.EFFICIENT_ATTENTION)):
        def fn(be=be):
            with sdpa_kernel(be):
                return F.scaled_dot_product_attention(qt, kt, vt)
        record(name, fn, check=lambda fn=fn: fn().transpose(1, 2))
    try:
        from flash_attn import flash_attn_func
        record("flash_attn.flash_attn_func (2.8.3)", lambda: flash_attn_func(q, k, v),
               check=lambda: flash_attn_func(q, k, v))
    except Exception as e:
        res["kernels"]["flash_attn.flash_attn_func"] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
    try:                                                            # flashinfer's Blackwell FMHA, if it can be had offline
        import signal

        def _alarm(*_):
            raise TimeoutError("flashinfer JIT/compile exceeded 300 s")
        signal.signal(signal.SIGALRM, _alarm)
        signal.alarm(300)
        import flashinfer
        qf, kf, vf = q[0], k[0], v[0]
        f1 = 4.0 * N * L * L * D

        def fi():
            return flashinfer.single_prefill_with_kv_cache(qf, kf, vf, causal=False)
        r = timed(fi, f1, args.sustain_s / 2)
        r["note"] = "single sample (B=1) per call"
        res["kernels"]["flashinfer.single_prefill_with_kv_cache"] = r
        signal.alarm(0)
    except BaseException as e:
        signal.alarm(0)
        res["kernels"]["flashinfer.single_prefill_with_kv_cache"] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
    out.copy_(ours)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
