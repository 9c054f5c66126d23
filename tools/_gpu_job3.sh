mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_block_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/pytest_k.log
cat gpurun_out/pytest_k.log
timeout 600 python bench.py --no-cpu-baseline --no-vae > gpurun_out/bench_r02_n1f.json 2> gpurun_out/bench_r02_n1f.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_n1f.json'))
print(d['value'], d['ms_per_step'], d['clocks'])
print(d['kernel_class_ms'])
print(d['roofline']['frac'], d['roofline']['avg_launch_ms'])
PY
