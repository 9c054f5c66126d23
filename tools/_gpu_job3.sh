mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_vae_gpu.py tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -15) > gpurun_out/pytest_vae_r02g.log
cat gpurun_out/pytest_vae_r02g.log
timeout 120 python tools/prof_conv_fused.py 96 24 > gpurun_out/conv_fused_96d.log 2>&1
timeout 120 python tools/prof_conv_fused.py 192 49 > gpurun_out/conv_fused_192d.log 2>&1
cat gpurun_out/conv_fused_96d.log gpurun_out/conv_fused_192d.log
timeout 300 python tools/vae_trace.py > gpurun_out/vae_trace_r02d.md 2> gpurun_out/vae_trace.err
head -30 gpurun_out/vae_trace_r02d.md; tail -3 gpurun_out/vae_trace.err
