: > gpurun_out/t_pkvar.log
for v in "$@"; do
cp tools/scratch/variants/libsdvg_$v.so sd-video-gen_b200/libsdvg.so; touch sd-video-gen_b200/libsdvg.so
echo "== $v" >> gpurun_out/t_pkvar.log
SDVG_PK=1 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" >> gpurun_out/t_pkvar.log
SDVG_PK=1 C1_B=1 timeout 200 python tools/c1_chain.py fp32 2>&1 | grep "us per pass" >> gpurun_out/t_pkvar.log
done
