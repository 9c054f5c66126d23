mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_vae_gpu.py tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -4) > gpurun_out/pytest_k.log
cat gpurun_out/pytest_k.log
timeout 300 python tools/vae_trace.py > gpurun_out/vae_trace_r02k.md 2> gpurun_out/vae_trace.err
head -8 gpurun_out/vae_trace_r02k.md | tail -4; grep "(1,3,3) (0,0,0)\|(3,1,1)" gpurun_out/vae_trace_r02k.md
