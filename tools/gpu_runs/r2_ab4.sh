: > gpurun_out/t_ab4.log
for i in 1 2; do
for v in old new; do
cp tools/scratch/variants/libsdvg_$v.so sd-video-gen_b200/libsdvg.so; touch sd-video-gen_b200/libsdvg.so
SDVG_PK=0 C1_W=10 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" | sed "s/^/$v W=10 /" >> gpurun_out/t_ab4.log
SDVG_PK=0 C1_B=24 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" | sed "s/^/$v /" >> gpurun_out/t_ab4.log
timeout 300 python bench_train.py --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$v train',d['ms_per_step'],d.get('roofline',{}).get('classes_ms'))" >> gpurun_out/t_ab4.log
done
done
cp tools/scratch/variants/libsdvg_new.so sd-video-gen_b200/libsdvg.so; touch sd-video-gen_b200/libsdvg.so
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -q -m gpu -k "compact or inline" 2>&1 | grep -E "^E  |passed|failed|Error" | head -20 >> gpurun_out/t_ab4.log
