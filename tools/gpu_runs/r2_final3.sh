set -x
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | grep -E "^E  |passed|failed|Error" | head -30 > gpurun_out/gputest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | grep -v -i warn | tail -5 > gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2g.json 2>gpurun_out/bench_r2g.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel --launch-skip 700 -c 8 -o gpurun_out/r2_gemm_full python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
ncu -i gpurun_out/r2_gemm_full.ncu-rep --page raw --csv > gpurun_out/r2_gemm_ncu_full_raw.csv 2>/dev/null
SDVG_PK=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/r2_c1_launches.csv python tools/c1_chain.py mixed > gpurun_out/ncu_c1.log 2>&1
SDVG_PK=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" --launch-skip 1200 -c 6 -o gpurun_out/r2_c1_gemm_full python tools/c1_chain.py mixed > gpurun_out/ncu_c1g.log 2>&1
ncu -i gpurun_out/r2_c1_gemm_full.ncu-rep --page raw --csv > gpurun_out/r2_c1_gemm_ncu_full_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
