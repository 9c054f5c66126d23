mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_block_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/pytest_k.log
cat gpurun_out/pytest_k.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r02_n1d.json 2> gpurun_out/bench_r02_n1d.err
tail -c 2600 gpurun_out/bench_r02_n1d.json
