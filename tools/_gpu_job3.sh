mkdir -p gpurun_out
(timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_vae_gpu.py -m gpu -q -x -k "fused_groupnorm or row_kernels or thin or conv" 2>&1 | tail -15) > gpurun_out/sanitizer_vae.log
tail -15 gpurun_out/sanitizer_vae.log
(timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "two_segments or guard_paths or rmsnorm or layernorm" 2>&1 | tail -15) > gpurun_out/sanitizer_k.log
tail -15 gpurun_out/sanitizer_k.log
