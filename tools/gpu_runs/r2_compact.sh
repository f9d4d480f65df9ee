timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | grep -E "^E  |passed|failed|Error" | head -30 > gpurun_out/t_compact.log
timeout 100 python tools/gemm_trace.py 40 2048 2048 fp16 32 2>&1 | grep -v -i warn | grep -E "event-timed|tile 0|exit" >> gpurun_out/t_compact.log
for i in 1 2; do
SDVG_PK=0 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" >> gpurun_out/t_compact.log
SDVG_PK=0 C1_B=1 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" >> gpurun_out/t_compact.log
done
