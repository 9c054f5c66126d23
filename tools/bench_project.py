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
    # 3DGS forward rasteriser at BASELINE config 5: 368 x 512 = 188 416 gaussians (scale 1e-4, identity
    # rotation, opacity 1: infer.py:263-270), all frames of the trajectory in one launch sequence
    gs = []
    for (H, W) in ((368, 512),):
        pts, col, ext, K = synth.point_cloud(H, W, 0, 0.05)
        V = a.frames
        exts = torch.stack([ext.clone() for _ in range(V)])
        for k in range(V):
            exts[k, 0, 3] += 0.002 * k
        moving = torch.stack([pts + 0.001 * k for k in range(V)]).cuda()
        c = (col / 255.0).cuda()
        args = (torch.ones(len(pts)), torch.tensor([1e-4] * 3), torch.tensor([0.0, 0.0, 0.0, 1.0]))
        render.gs_render_views(moving, c, *args, exts, K, H, W)
        torch.cuda.synchronize()
        s, t = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        img = render.gs_render_views(moving, c, *args, exts, K, H, W)
        t.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(t)
        from oracle import gs_oracle as G
        import numpy as np
        t0 = time.perf_counter()
        G.render(pts.numpy()[:20000], (col / 255.0).numpy()[:20000], np.ones(20000, np.float32), [1e-4] * 3, [0, 0, 0, 1],
                 ext.numpy(), K.numpy(), H, W)
        cpu_ms = (time.perf_counter() - t0) * 1e3 * len(pts) / 20000
        gs.append({"frame": f"{H}x{W}", "gaussians": len(pts), "views": V, "gpu_ms_total": ms,
                   "gpu_ms_per_frame": ms / V, "frames_per_s": V * 1e3 / ms, "finite": bool(torch.isfinite(img).all()),
                   "cpu_oracle_ms_per_frame_extrapolated": cpu_ms,
                   "note": "cpu: numpy oracle on 20 000 of the gaussians, scaled linearly"})
    print(json.dumps({"workload": "render_with_project z-buffer, synthetic point clouds", "results": out,
                      "gs_render": gs}))


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    main()
