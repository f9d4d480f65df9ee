set -x
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/gputest.log
timeout 300 python bench_train.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/train_r2c.json 2>gpurun_out/train_r2c.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/bench_r2b.json 2>gpurun_out/bench_r2b.err
