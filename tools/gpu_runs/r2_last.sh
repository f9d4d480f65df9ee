timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | grep -E "^E  |passed|failed|Error" | head -10 > gpurun_out/gputest_last.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v -i warn | tail -4 >> gpurun_out/gputest_last.log
