mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | cut -c1-300 | grep -v "^E    .*where\|^E    .*tensor" | tail -60) > gpurun_out/pytest_gpu_r02b.log
timeout 200 python tools/rows_probe.py > gpurun_out/rows_r02_events.json 2> gpurun_out/rows_r02_events.log
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__block_size --clock-control none --csv --log-file gpurun_out/rows_r02_ncu.csv python tools/rows_probe.py --once > /dev/null 2>&1
ATTN_L=50400 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:attn_fwd -s 1 -c 1 --csv --log-file gpurun_out/attn_traffic_50400.csv python tools/prof_attn.py > /dev/null 2>&1
timeout 500 python tools/attn_library_bar.py > gpurun_out/attn_library_bar_r02b.json 2> gpurun_out/attn_library_bar_r02b.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err
tail -5 gpurun_out/pytest_gpu_r02b.log
cut -c1-230 gpurun_out/attn_library_bar_r02b.log | head -3
cat gpurun_out/rows_r02_events.log
cut -c1-1500 gpurun_out/bench_r02_n1.json
tail -3 gpurun_out/bench_r02_n1.err
