timeout 120 python tools/pk_check.py gemm 2>&1 | grep -v -i warn | grep -E "fp16|fp32|worst|Error|error" 
timeout 200 python tools/pk_check.py forward 2>&1 | grep -v -i warn | grep -E "forward d256|Error|error" 
timeout 300 python tools/pk_check.py rollout 2>&1 | grep -v -i warn | grep -E "rollout|max-rel|Error|error"
SDVG_PK_TRACE=8,0,533 timeout 100 python tools/pk_trace_rollout.py mixed 5 2>&1 | grep -v -i warn | grep -v TransformerEnc
