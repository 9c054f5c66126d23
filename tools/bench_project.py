"""Stage hand-off renderer (render_with_project, infer.py:222-258): frames/s of the CUDA z-buffer
at BASELINE config 5's shape (368 x 512 = 188 416 points per frame, 49 frames) and at 720 x 1280,
with the numpy oracle timed beside it on the host.
    python tools/bench_project.py [--frames 49]"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from more4d_b200 import render, synth          # noqa: E402
from oracle import project_oracle as P         # noqa: E402

# algorithmic bytes per point: points 12 + colours 12 + index/depth scratch 8 w + 8 r + z-buffer
# atomic 4 + accumulator atomics 16, per pixel: accumulator 16 r + image 3 w + mask 1 w
BYTES_PER_POINT, BYTES_PER_PIXEL = 60, 20


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=49)
    a = ap.parse_args()
    out = []
    for (H, W) in ((368, 512), (720, 1280)):
        frames = [synth.point_cloud(H, W, s % 4, 0.05 * (s % 7)) for s in range(4)]
        dev = [(p.cuda(), c.cuda(), e, k) for p, c, e, k in frames]
        for p, c, e, k in dev:                                   # warm-up
            render.project_points(p, e, k, c, H, W)
        torch.cuda.synchronize()
        s, t = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        for f in range(a.frames):
            p, c, e, k = dev[f % 4]
            render.project_points(p, e, k, c, H, W)
        t.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(t) / a.frames
        # all frames of the trajectory in one launch sequence (m4d_project_views)
        pts_v = torch.stack([dev[f % 4][0] for f in range(a.frames)])
        col_v = torch.stack([dev[f % 4][1] for f in range(a.frames)])
        ext_v = torch.stack([dev[f % 4][2] for f in range(a.frames)])
        render.project_views(pts_v, ext_v, dev[0][3], col_v, H, W)
        torch.cuda.synchronize()
        s.record()
        render.project_views(pts_v, ext_v, dev[0][3], col_v, H, W)
        t.record()
        torch.cuda.synchronize()
        ms_batched = s.elapsed_time(t) / a.frames
        t0 = time.perf_counter()
        p, c, e, k = frames[0]
        P.render_with_project(p.numpy(), torch.linalg.inv(e).numpy(), k.numpy(), c.numpy(), H, W)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        nbytes = H * W * (BYTES_PER_POINT + BYTES_PER_PIXEL)
        out.append({"frame": f"{H}x{W}", "points": H * W, "gpu_ms_per_frame": ms, "frames_per_s": 1e3 / ms,
                    "algorithmic_GBps": nbytes / ms / 1e6,
                    "batched_gpu_ms_per_frame": ms_batched, "batched_frames_per_s": 1e3 / ms_batched,
                    "batched_algorithmic_GBps": nbytes / ms_batched / 1e6, "cpu_oracle_ms_per_frame": cpu_ms,
                    "note": "includes the host-side 4x4 inverse and workspace allocation of the public call"})
    print(json.dumps({"workload": "render_with_project z-buffer, synthetic point clouds", "results": out}))


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    main()
