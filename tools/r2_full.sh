set -x
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/gputest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2c.json 2>gpurun_out/bench_r2c.err
