: > gpurun_out/t_gap.log
for i in 1 2; do
SDVG_PK=0 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" | sed "s/^/default /" >> gpurun_out/t_gap.log
SDVG_PK=0 C1_MAXTOK=5 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" | sed "s/^/maxtok5 /" >> gpurun_out/t_gap.log
SDVG_PK=0 C1_PE=1 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" | sed "s/^/pe /" >> gpurun_out/t_gap.log
SDVG_PK=0 C1_PE=1 C1_MAXTOK=5 timeout 200 python tools/c1_chain.py mixed 2>&1 | grep "us per pass" | sed "s/^/pe+maxtok5 /" >> gpurun_out/t_gap.log
done
