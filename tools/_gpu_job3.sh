mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_block_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -4) > gpurun_out/pytest_k.log
cat gpurun_out/pytest_k.log
timeout 200 python tools/prof_cross_attn.py 2>&1 | tail -4
