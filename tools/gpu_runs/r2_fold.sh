set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | grep -E "^E  |passed|failed|Error" | head -30 > gpurun_out/t_fold.log
for i in 1 2; do
SDVG_LN_FOLD_ACC=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ACC=1 ', round(d['value']), d['ms_per_step'], d['roofline']['classes_ms'], d['clocks']['sm_mhz'], d['roofline']['frac'], d['roofline']['step_frac_of_sustained'])" >> gpurun_out/ab_fold.log
SDVG_LN_FOLD_ACC=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ACC=0 ', round(d['value']), d['ms_per_step'], d['roofline']['classes_ms'], d['clocks']['sm_mhz'], d['roofline']['frac'], d['roofline']['step_frac_of_sustained'])" >> gpurun_out/ab_fold.log
done
SDVG_LN_FOLD=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('FOLD=0', round(d['value']), d['ms_per_step'], d['roofline']['classes_ms'], d['clocks']['sm_mhz'], d['roofline']['frac'], d['roofline']['step_frac_of_sustained'])" >> gpurun_out/ab_fold.log
