mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_vae_gpu.py tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -15) > gpurun_out/pytest_vae_r02e.log
cat gpurun_out/pytest_vae_r02e.log
timeout 300 python tools/prof_conv_fused.py 96 24 > gpurun_out/conv_fused_96b.log 2>&1
timeout 300 python tools/prof_conv_fused.py 192 49 > gpurun_out/conv_fused_192b.log 2>&1
cat gpurun_out/conv_fused_96b.log gpurun_out/conv_fused_192b.log
timeout 300 python tools/vae_trace.py > gpurun_out/vae_trace_r02b.md 2> gpurun_out/vae_trace.err
head -40 gpurun_out/vae_trace_r02b.md
timeout 200 python tools/rows_probe.py 2>&1 >/dev/null | grep -i "rmsnorm_silu\|groupnorm"
