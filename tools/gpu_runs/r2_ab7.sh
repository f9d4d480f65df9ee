: > gpurun_out/t_ab7.log
for i in 1 2; do
for v in new new2; do
cp tools/scratch/variants/libsdvg_$v.so sd-video-gen_b200/libsdvg.so; touch sd-video-gen_b200/libsdvg.so
SDVG_PK=0 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" | sed "s/^/$v /" >> gpurun_out/t_ab7.log
SDVG_PK=0 C1_B=1 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" | sed "s/^/$v /" >> gpurun_out/t_ab7.log
done
done
cp tools/scratch/variants/libsdvg_new2.so sd-video-gen_b200/libsdvg.so; touch sd-video-gen_b200/libsdvg.so
timeout 100 python tools/gemm_trace.py 40 2048 2048 fp16 32 2>&1 | grep -v -i warn >> gpurun_out/t_ab7.log
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -q -m gpu -x 2>&1 | grep -E "^E  |passed|failed|Error" | head -20 >> gpurun_out/t_ab7.log
