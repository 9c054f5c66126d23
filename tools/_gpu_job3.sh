mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4) > gpurun_out/pytest_gpu_r02h.log
cat gpurun_out/pytest_gpu_r02h.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r02_n2c.json 2> gpurun_out/bench_r02_n2c.err
python - <<'PY'
import json
s=open('gpurun_out/bench_r02_n2c.json').read(); d=json.loads(s[s.index('{'):])
print(d['value'], d['ms_per_step'], d['sp_bit_exact'], d['sp']['ms_per_step'], d['sp']['speedup_vs_one_gpu_step'])
print({k:v for k,v in d['kernel_class_ms'].items() if k not in ('per_rank','note')})
PY
