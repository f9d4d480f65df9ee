timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | grep -E "^E  |passed|failed|Error" | head -30 > gpurun_out/t_inline.log
for i in 1 2; do
SDVG_PK=0 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" >> gpurun_out/t_inline.log
SDVG_PK=0 SDVG_STATS_INLINE_MAX=0 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" | sed 's/$/ (finalize kernel)/' >> gpurun_out/t_inline.log
SDVG_PK=0 C1_B=1 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" >> gpurun_out/t_inline.log
SDVG_PK=0 C1_B=1 SDVG_STATS_INLINE_MAX=0 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" | sed 's/$/ (finalize kernel)/' >> gpurun_out/t_inline.log
done
