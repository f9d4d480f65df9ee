set -x
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | grep -E "^E  |passed|failed|Error" | head -30 > gpurun_out/gputest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | grep -v -i warn | tail -5 > gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2e.json 2>gpurun_out/bench_r2e.err
