#!/bin/bash
# ncu --set full capture of one launch of the bulk fill kernel on cfg2 (run under gpurun); TAG names the report
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
$NCU -k 'regex:yb_fill2_kernel<.int.128' -s 3 -c 1 -f -o $O/fill2_$TAG python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_fill2_$TAG.log 2>&1
tail -3 $O/ncu_fill2_$TAG.log
