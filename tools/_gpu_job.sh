mkdir -p gpurun_out
timeout 700 python bench.py > gpurun_out/bench_r02_n1i.json 2> gpurun_out/bench_r02_n1i.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_n1i.json'))
print(d['value'], d['ms_per_step'], d['clocks'], d['e2e']['value'])
print(d['kernel_class_ms'])
print(d['roofline']['frac'], d['roofline']['avg_launch_ms'])
v=d['vae_roundtrip']; print(v['total_ms'], v['frac_of_sustained_tensor_peak'], v['e2e']['ms'])
PY
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
