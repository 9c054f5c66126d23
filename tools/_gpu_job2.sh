mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err
tail -c 3000 gpurun_out/bench_r02_n2.json
tail -5 gpurun_out/bench_r02_n2.err
(timeout 300 python -m pytest tests/test_sp_gpu.py tests/test_kernels_gpu.py -q -k "sequence or layernorm or rmsnorm" 2>&1 | tail -5) > gpurun_out/pytest_sp.log
cat gpurun_out/pytest_sp.log
