mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | cut -c1-300 | grep -v "^E    .*where\|^E    .*tensor" | tail -40) > gpurun_out/pytest_gpu_r02g.log
tail -3 gpurun_out/pytest_gpu_r02g.log
timeout 700 python bench.py > gpurun_out/bench_r02_n1g.json 2> gpurun_out/bench_r02_n1g.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_n1g.json'))
print(d['value'], d['ms_per_step'], d['clocks'], d['e2e']['value'])
print(d['kernel_class_ms'])
print(d['roofline']['frac'], d['roofline']['avg_launch_ms'])
v=d['vae_roundtrip']; print(v['total_ms'], v['frac_of_sustained_tensor_peak'], v['e2e'], {k:round(s['ms'],1) for k,s in v['stages'].items()})
PY
