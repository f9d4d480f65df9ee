: > gpurun_out/t_ab8.log
for i in 1 2; do
for v in base skip; do
cp tools/scratch/variants/libsdvg_$v.so sd-video-gen_b200/libsdvg.so; touch sd-video-gen_b200/libsdvg.so
SDVG_PK=0 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" | sed "s/^/$v /" >> gpurun_out/t_ab8.log
SDVG_PK=0 C1_B=1 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" | sed "s/^/$v /" >> gpurun_out/t_ab8.log
done
done
cp tools/scratch/variants/libsdvg_skip.so sd-video-gen_b200/libsdvg.so; touch sd-video-gen_b200/libsdvg.so
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x 2>&1 | grep -E "^E  |passed|failed|Error" | head -5 >> gpurun_out/t_ab8.log
