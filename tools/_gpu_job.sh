mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | cut -c1-300 | grep -v "^E    .*where\|^E    .*tensor" | tail -40) > gpurun_out/pytest_gpu_r02f.log
tail -5 gpurun_out/pytest_gpu_r02f.log
timeout 200 python tools/rows_probe.py > gpurun_out/rows_r02_events.json 2> gpurun_out/rows_r02_events.log
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__block_size --clock-control none --csv --log-file gpurun_out/rows_r02_ncu.csv python tools/rows_probe.py --once > /dev/null 2>&1
timeout 700 python bench.py > gpurun_out/bench_r02_n1e.json 2> gpurun_out/bench_r02_n1e.err
tail -c 1500 gpurun_out/bench_r02_n1e.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
