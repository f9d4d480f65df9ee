timeout 600 python -m pytest tests/test_gpu_train.py -q -m gpu -k "c5_bench_shape" 2>&1 | grep -v Warning | tail -60 > gpurun_out/t1.log
