set -x
timeout 300 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" > gpurun_out/c1_chain.log
SDVG_PK=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 1300 --csv --log-file gpurun_out/c1_launches.csv python tools/c1_chain.py mixed 0 > gpurun_out/c1_ncu.log 2>&1
timeout 600 python -m pytest tests/test_gpu_train.py -q -m gpu -k "c5_bench_shape" 2>&1 | grep -E "^E|passed|failed" | head -20 > gpurun_out/t1.log
