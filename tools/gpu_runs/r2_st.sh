: > gpurun_out/t_st.log
for i in 1 2; do
for v in st8 st9; do
cp tools/scratch/variants/libsdvg_$v.so sd-video-gen_b200/libsdvg.so; touch sd-video-gen_b200/libsdvg.so
SDVG_PK=0 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" | sed "s/^/$v /" >> gpurun_out/t_st.log
SDVG_PK=0 C1_B=1 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" | sed "s/^/$v /" >> gpurun_out/t_st.log
done
done
