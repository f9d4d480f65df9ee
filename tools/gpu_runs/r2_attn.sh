set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -q -m gpu -x 2>&1 | grep -E "^E|passed|failed" | head -20 > gpurun_out/t_attn.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_attn.json 2>gpurun_out/bench_attn.err
