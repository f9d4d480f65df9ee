: > gpurun_out/t_ab2.log
for i in 1 2; do
for v in old new; do
cp tools/scratch/variants/libsdvg_$v.so sd-video-gen_b200/libsdvg.so; touch sd-video-gen_b200/libsdvg.so
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$v',d['value'],d['ms_per_step'],d['roofline']['classes_ms'])" >> gpurun_out/t_ab2.log
done
done
