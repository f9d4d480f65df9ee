set -x
timeout 200 python tools/pk_check.py gemm 2>&1 | grep -v -i warn | grep -E "pk gemm|worst|Error|error|timed out|Assert" > gpurun_out/pk_gemm_v2.log
timeout 300 python tools/pk_check.py forward 2>&1 | grep -v -i warn | grep -E "forward d|Error|error|timed out" > gpurun_out/pk_fwd_v2.log
timeout 400 python tools/pk_check.py rollout 2>&1 | grep -v -i warn | grep -E "rollout|max-rel|Error|error|timed out" > gpurun_out/pk_roll_v2.log
SDVG_PK=1 SDVG_PK_TRACE=14,0,533 timeout 120 python tools/pk_trace_rollout.py mixed 5 2>&1 | grep -v -i warn | grep -v TransformerEnc > gpurun_out/pk_trace_v2.log
SDVG_PK=1 SDVG_PK_V2=0 timeout 300 python tools/pk_check.py rollout 2>&1 | grep -v -i warn | grep -E "persistent=True" > gpurun_out/pk_roll_v1.log
timeout 600 python -m pytest tests/test_gpu_train.py -q -m gpu -k "c5_bench_shape" 2>&1 | grep -E "^E|passed|failed" | head -20 > gpurun_out/t1.log
