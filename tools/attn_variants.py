"""A/B timing of the attention kernel's compile-time variants (development build only).

    python more4d_b200/build.py --force -DM4D_DEV && python tools/attn_variants.py > gpurun_out/attn_variants.json

For every (MODE, PP): burst (best of 5) and sustained (back-to-back for ~2.5 s, second half averaged)
TFLOP/s at B=2, L=50 400, 40 heads, d=128, the nvidia-smi clock / power during the sustained loop,
and the rel. error against the exact-max variant (MODE 0, PP 0).  MODE 1 = sum-guarded speculative
reference, PP = pairs of every 8 whose 2^x runs on the FMA pipe."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from more4d_b200 import _lib, ops                        # noqa: E402
from tools.attn_library_bar import timed                 # noqa: E402

BF16 = torch.bfloat16


def main():
    L = int(os.environ.get("ATTN_L", "50400"))
    B, N, D = 2, 40, 128
    flops = 4.0 * B * N * L * L * D
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v = (torch.randn(B, L, N, D, device="cuda", dtype=BF16, generator=g) for _ in range(3))
    out = torch.empty_like(q)
    res = {"shape": {"B": B, "L": L, "heads": N}, "variants": {}}
    base = None
    variants = [(0, 0), (1, 2), (2, 2), (2, 1), (2, 3), (2, 0), (2, 4), (1, 3)] if _lib.is_dev_build() else [None]
    for var in variants:
        if var is not None:
            _lib.dev_set_flags(0x100 | (var[0] << 4) | var[1])
        name = "product default" if var is None else f"MODE{var[0]}_PP{var[1]}"
        r = timed(lambda: ops.attention(q, k, v, out=out), flops, 2.5)
        if base is None:
            base = out.clone()
        r["rel_err_vs_first"] = float((out.float() - base.float()).norm() / base.float().norm())
        res["variants"][name] = r
        print(name, json.dumps(r), file=sys.stderr, flush=True)
    if _lib.is_dev_build():
        _lib.dev_set_flags(0)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
