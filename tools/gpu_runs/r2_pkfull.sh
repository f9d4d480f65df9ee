timeout 600 python -m pytest tests/test_gpu_persistent.py -q -m gpu 2>&1 | grep -E "^E  |passed|failed|Error" | head -20 > gpurun_out/t_pkfull.log
timeout 200 python tools/pk_check.py gemm 2>&1 | grep -v -i warn | grep -E "pk gemm|worst|Error|error|timed out|Assert" | tail -5 >> gpurun_out/t_pkfull.log
timeout 300 python tools/pk_check.py forward 2>&1 | grep -v -i warn | grep -E "forward d|Error|error|timed out" | tail -8 >> gpurun_out/t_pkfull.log
timeout 400 python tools/pk_check.py rollout 2>&1 | grep -v -i warn | grep -E "rollout|max-rel|Error|error|timed out" | tail -8 >> gpurun_out/t_pkfull.log
for i in 1 2; do
SDVG_PK=1 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" >> gpurun_out/t_pkfull.log
SDVG_PK=1 C1_B=1 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" >> gpurun_out/t_pkfull.log
SDVG_PK=1 timeout 200 python tools/c1_chain.py fp32 2>&1 | grep "us per pass" >> gpurun_out/t_pkfull.log
SDVG_PK=1 C1_B=1 timeout 200 python tools/c1_chain.py fp32 2>&1 | grep "us per pass" >> gpurun_out/t_pkfull.log
done
SDVG_PK=0 C1_B=1 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" >> gpurun_out/t_pkfull.log
SDVG_PK=0 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" >> gpurun_out/t_pkfull.log
