mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | cut -c1-300 | grep -v "^E    .*where\|^E    .*tensor" | tail -40) > gpurun_out/pytest_gpu_r02e.log
tail -5 gpurun_out/pytest_gpu_r02e.log
timeout 600 python bench.py > gpurun_out/bench_r02_n1c.json 2> gpurun_out/bench_r02_n1c.err
tail -c 2500 gpurun_out/bench_r02_n1c.json
