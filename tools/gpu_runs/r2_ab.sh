for i in 1 2; do
SDVG_LN_PRE=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PRE=0', round(d['value']), d['ms_per_step'], d['roofline']['classes_ms'], d['clocks']['sm_mhz'])" >> gpurun_out/ab.log
SDVG_LN_PRE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PRE=1', round(d['value']), d['ms_per_step'], d['roofline']['classes_ms'], d['clocks']['sm_mhz'])" >> gpurun_out/ab.log
done
