timeout 600 python -m pytest tests/test_gpu_persistent.py -q -m gpu 2>&1 | grep -E "^E  |passed|failed|Error" | head -20 > gpurun_out/t_pkbar.log
for i in 1 2; do
SDVG_PK=1 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" >> gpurun_out/t_pkbar.log
SDVG_PK=1 timeout 200 python tools/c1_chain.py fp32 2>&1 | grep "us per pass" >> gpurun_out/t_pkbar.log
SDVG_PK=1 C1_B=1 timeout 200 python tools/c1_chain.py fp32 2>&1 | grep "us per pass" >> gpurun_out/t_pkbar.log
done
SDVG_PK=0 C1_B=1 timeout 200 python tools/c1_chain.py fp32 2>&1 | grep "us per pass" >> gpurun_out/t_pkbar.log
