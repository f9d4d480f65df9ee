: > gpurun_out/t_gtrace.log
for bn in 32 64; do
timeout 100 python tools/gemm_trace.py 40 2048 2048 fp16 $bn 2>&1 | grep -v -i warn >> gpurun_out/t_gtrace.log
done
timeout 100 python tools/gemm_trace.py 40 6144 2048 fp16 64 2>&1 | grep -v -i warn >> gpurun_out/t_gtrace.log
timeout 100 python tools/gemm_trace.py 5 2048 2048 fp16 32 2>&1 | grep -v -i warn >> gpurun_out/t_gtrace.log
