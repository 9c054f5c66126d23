mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_project_gpu.py tests/test_gsplat_gpu.py tests/test_kernels_gpu.py tests/test_vae_gpu.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/pytest_k.log
cat gpurun_out/pytest_k.log
timeout 300 python tools/bench_project.py > gpurun_out/project_r02.json 2> gpurun_out/project_r02.err
cat gpurun_out/project_r02.json; tail -3 gpurun_out/project_r02.err
