mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_vae_gpu.py tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -25) > gpurun_out/pytest_vae_r02h.log
cat gpurun_out/pytest_vae_r02h.log
timeout 300 python tools/vae_trace.py > gpurun_out/vae_trace_r02e.md 2> gpurun_out/vae_trace.err
head -34 gpurun_out/vae_trace_r02e.md; tail -3 gpurun_out/vae_trace.err
timeout 200 python tools/rows_probe.py 2>&1 >/dev/null | grep -i "softmax\|groupnorm\|rmsnorm_silu"
