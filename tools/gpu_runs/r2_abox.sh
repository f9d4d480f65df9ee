timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_persistent.py -q -m gpu 2>&1 | grep -E "^E  |passed|failed|Error" | head -20 > gpurun_out/t_abox.log
for i in 1 2; do
SDVG_PK=0 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" >> gpurun_out/t_abox.log
done
SDVG_PK=0 timeout 200 python tools/c1_chain.py fp32 2>&1 | grep "us per pass" >> gpurun_out/t_abox.log
timeout 300 python bench_train.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train', d['ms_per_step'])" >> gpurun_out/t_abox.log
