mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_probe tools/pipe_probe.cu && /tmp/pipe_probe > gpurun_out/pipe_probe_r02b.log 2>&1
grep -i "f16\|bf16\|MUFU\|mix" gpurun_out/pipe_probe_r02b.log
