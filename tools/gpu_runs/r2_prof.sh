set -x
timeout 600 python tools/c5_diag.py 0.0 2>&1 | grep -v -i warn > gpurun_out/c5_diag.log
timeout 120 python tools/gemm_trace.py 5120 2048 2048 fp16 -192 2>&1 | grep -v -i warn > gpurun_out/gemm_trace_192.log
timeout 120 python tools/gemm_trace.py 5120 6144 2048 fp16 -256 2>&1 | grep -v -i warn > gpurun_out/gemm_trace_256.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2.json 2>gpurun_out/bench_ref_r2.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_mma_kernel|layernorm_block_kernel" --launch-skip 900 -c 8 -o gpurun_out/r2_attn_ln_full python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
ncu -i gpurun_out/r2_attn_ln_full.ncu-rep --page raw --csv > gpurun_out/r2_attn_ln_ncu_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_attn_ln_full.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_attn_ln_source.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
