mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_vae_gpu.py tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/pytest_k.log
cat gpurun_out/pytest_k.log
timeout 120 python tools/prof_conv_thin.py 16 2>&1 | tail -6
timeout 120 python tools/prof_conv_fused.py 96 24 2>&1 | tail -3
timeout 300 python tools/vae_trace.py > gpurun_out/vae_trace_r02h.md 2> gpurun_out/vae_trace.err
head -18 gpurun_out/vae_trace_r02h.md; tail -3 gpurun_out/vae_trace.err
timeout 200 python tools/rows_probe.py 2>&1 >/dev/null | grep -i "rmsnorm_silu"
