#!/bin/bash
# Evidence for profiles/: run ON THE GPU BOX (gpurun -- 'bash tools/capture_profiles.sh TAG').  Everything lands in
# gpurun_out/; tools/ncu_summary.py turns the .ncu-rep files into the markdown summaries kept under profiles/.
#   bench line + reference arm (never under a profiler), the ncu launch list of the same bench command, and one
#   `ncu --set full` capture per kernel: K2 bulk bin, K1, K3 (both traceback kernels), block scoring.
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err
python bench.py --impl reference > $O/bench_ref_$TAG.json 2>> $O/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 > $O/ncu_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
$NCU -k 'regex:fill_kernel_w<.int.128' -s 4 -c 1 -f -o $O/fill_$TAG python bench.py --steps 2 --warmup 1 > $O/ncu_fill.log 2>&1
$NCU -k 'regex:yb_profile_kernel' -s 8 -c 1 -f -o $O/prof_$TAG python bench.py --steps 2 --warmup 1 > $O/ncu_prof.log 2>&1
$NCU -k 'regex:yb_traceback' -s 8 -c 2 -f -o $O/tb_$TAG python bench.py --steps 2 --warmup 1 > $O/ncu_tb.log 2>&1
$NCU -k 'regex:yb_score_kernel' -s 8 -c 1 -f -o $O/score_$TAG python tools/score_bench.py deep > $O/ncu_score.log 2>&1
python tools/score_bench.py all --json $O/score_bench_$TAG.json > /dev/null 2>> $O/bench_$TAG.err
cat $O/bench_$TAG.json
