mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_vae_gpu.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/pytest_k.log
cat gpurun_out/pytest_k.log
timeout 300 python tools/vae_trace.py > gpurun_out/vae_trace_r02i.md 2> gpurun_out/vae_trace.err
head -18 gpurun_out/vae_trace_r02i.md; grep "linear\|(1,1,1)" gpurun_out/vae_trace_r02i.md | head; tail -3 gpurun_out/vae_trace.err
