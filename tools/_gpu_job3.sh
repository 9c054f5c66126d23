mkdir -p gpurun_out
timeout 1200 ncu --nvtx --nvtx-include "m4d_timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 1 --no-vae --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
wc -l gpurun_out/launches_r02.csv; tail -2 gpurun_out/launches_bench.log | cut -c1-300
