timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2h.json 2>gpurun_out/bench_r2h.err
