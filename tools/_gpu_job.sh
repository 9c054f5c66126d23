mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4) > gpurun_out/pytest_gpu_r02i.log
cat gpurun_out/pytest_gpu_r02i.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 700 python bench.py > gpurun_out/bench_r02_n1j.json 2> gpurun_out/bench_r02_n1j.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_n1j.json'))
print(d['value'], d['ms_per_step'], d['clocks'], d['e2e']['value'])
print({k:v for k,v in d['kernel_class_ms'].items() if k!='note'})
print(d['roofline']['frac'], d['roofline']['avg_launch_ms'])
v=d['vae_roundtrip']; print(v['total_ms'], v['frac_of_sustained_tensor_peak'], v['e2e']['ms'])
PY
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | head -c 300
