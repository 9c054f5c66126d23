mkdir -p gpurun_out
timeout 400 python tools/bench_project.py > gpurun_out/project_r02.json 2> gpurun_out/project_r02.err
cat gpurun_out/project_r02.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['gs_render']); print([ (r['frame'], r['gpu_ms_per_frame'], r['batched_gpu_ms_per_frame']) for r in d['results']])"
tail -3 gpurun_out/project_r02.err
