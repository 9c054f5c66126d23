mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_r02_n4.json 2> gpurun_out/bench_r02_n4.err
head -c 400 gpurun_out/bench_r02_n4.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r02_n2b.json 2> gpurun_out/bench_r02_n2b.err
head -c 400 gpurun_out/bench_r02_n2b.json; echo
