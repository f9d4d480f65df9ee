: > gpurun_out/t_ks.log
for i in 1 2; do
SDVG_PK=0 SDVG_LN_FOLD=0 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" | sed 's/$/ nofold/' >> gpurun_out/t_ks.log
SDVG_PK=0 SDVG_LN_FOLD=0 SDVG_KSPLIT=1 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass\|rror" | sed 's/$/ nofold ksplit/' >> gpurun_out/t_ks.log
done
