for i in 1 2; do
SDVG_PK=0 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" | sed 's/^/fold>=512: /' >> gpurun_out/c1fold.log
SDVG_PK=0 SDVG_LN_FOLD_MIN=1 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" | sed 's/^/fold>=1:   /' >> gpurun_out/c1fold.log
done
