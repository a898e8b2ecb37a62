#!/bin/bash
# round 2, first GPU call: the whole GPU suite (incl. the cfg5-shape, depth-64 and tba tests), a baseline bench line,
# racecheck on the explicit-barrier build
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $O/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2a_pytest_gpu.txt 2>&1
tail -5 $O/r2a_pytest_gpu.txt
timeout 600 python bench.py > $O/r2a_bench.json 2> $O/r2a_bench.err
cat $O/r2a_bench.json
for tool in racecheck; do
  YAMA_B200_LIB=$PWD/multiz_b200/libyama_b200_sync.so timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_probe.py > $O/r2a_sanitize_${tool}_sync.log 2>&1
  tail -4 $O/r2a_sanitize_${tool}_sync.log
done
