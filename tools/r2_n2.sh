set -x
timeout 600 python -m pytest tests/test_gpu_train.py -q -m gpu -k "c5_bench_shape" 2>&1 | grep -E "^E|passed|failed" | head -20 > gpurun_out/t1.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2_n2.json 2>gpurun_out/bench_r2_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2>gpurun_out/bench_ref_n2.err
