mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | cut -c1-300 | grep -v "^E    .*where\|^E    .*tensor" | tail -80) > gpurun_out/pytest_gpu_r02d.log
tail -5 gpurun_out/pytest_gpu_r02d.log
timeout 300 python tools/vae_trace.py > gpurun_out/vae_trace_r02.md 2> gpurun_out/vae_trace.err
timeout 200 python tools/rows_probe.py > gpurun_out/rows_r02_events.json 2> gpurun_out/rows_r02_events.log
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__block_size --clock-control none --csv --log-file gpurun_out/rows_r02_ncu.csv python tools/rows_probe.py --once > /dev/null 2>&1
timeout 600 python bench.py > gpurun_out/bench_r02_n1b.json 2> gpurun_out/bench_r02_n1b.err
tail -c 3000 gpurun_out/bench_r02_n1b.json
head -40 gpurun_out/vae_trace_r02.md
