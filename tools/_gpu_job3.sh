mkdir -p gpurun_out
rm -f gpurun_out/bench_r02_other_workloads.jsonl
for w in 480p-14b 720p-1.3b 480p-1.3b visim-368p-14b; do
  timeout 400 python bench.py --workload $w --no-cpu-baseline --no-vae 2> gpurun_out/other_$w.err >> gpurun_out/bench_r02_other_workloads.jsonl
  tail -1 gpurun_out/other_$w.err | cut -c1-200
done
python - <<'PY'
import json
for l in open('gpurun_out/bench_r02_other_workloads.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:40], round(d['value'],4), round(d['ms_per_step'],1), round(d['roofline']['frac'],3), round(d['kernel_class_ms']['cross_attention'],1))
PY
