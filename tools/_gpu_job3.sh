mkdir -p gpurun_out
timeout 300 python tools/vae_trace.py > gpurun_out/vae_trace_r02j.md 2> gpurun_out/vae_trace.err
head -8 gpurun_out/vae_trace_r02j.md | tail -4; grep "49x720x1280x96 96x2592 96x1x1x1\|49x360x640x192 192x5184 192x1x1x1" gpurun_out/vae_trace_r02j.md
