SDVG_PK=1 SDVG_PK_TRACE=8,0,533 timeout 120 python tools/pk_trace_rollout.py mixed 5 2>&1 | grep -v -i warn | grep -v TransformerEnc > gpurun_out/pk_trace_mixed8.log
C1_B=1 SDVG_PK=1 SDVG_PK_TRACE=8,0,533 timeout 120 python tools/pk_trace_rollout.py fp32 5 2>&1 | grep -v -i warn | grep -v TransformerEnc > gpurun_out/pk_trace_fp32_1.log
