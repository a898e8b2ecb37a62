#!/bin/bash
# round 2 evidence set (run under gpurun): GPU test suite, bench line + reference arm, ncu launch list of the bench command,
# `ncu --set full` of the bulk fill kernel / K1 / K3 / the CTA-per-pair kernel on cfg5, bench lines of cfg3 / cfg5,
# racecheck on the explicit-barrier build and on the shipped build.  Everything lands in gpurun_out/.
TAG=${1:-r2b}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.txt 2>&1
tail -5 $O/${TAG}_pytest_gpu.txt
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
cat $O/${TAG}_bench.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
NCU="timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
$NCU -k 'regex:yb_fill2_kernel<.int.128' -s 3 -c 1 -f -o $O/${TAG}_fill2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu_fill2.log 2>&1
$NCU -k 'regex:yb_profile_kernel' -s 3 -c 1 -f -o $O/${TAG}_prof python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu_prof.log 2>&1
$NCU -k 'regex:yb_traceback' -s 6 -c 2 -f -o $O/${TAG}_tb python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu_tb.log 2>&1
timeout 600 python bench.py --workload cfg5 --no-cpu-baseline > $O/${TAG}_bench_cfg5.json 2>> $O/${TAG}_bench.err
cat $O/${TAG}_bench_cfg5.json
$NCU -k 'regex:yb_fill_kernel_w<.int.2048' -s 3 -c 1 -f -o $O/${TAG}_fill_cta python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu_fill_cta.log 2>&1
timeout 600 python bench.py --workload cfg3 --no-cpu-baseline > $O/${TAG}_bench_cfg3.json 2>> $O/${TAG}_bench.err
cat $O/${TAG}_bench_cfg3.json
YAMA_B200_LIB=$PWD/multiz_b200/libyama_b200_sync.so timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_probe.py > $O/${TAG}_racecheck_sync.log 2>&1
tail -4 $O/${TAG}_racecheck_sync.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_probe.py > $O/${TAG}_racecheck_shipped.log 2>&1
tail -4 $O/${TAG}_racecheck_shipped.log
for tool in memcheck initcheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_probe.py > $O/${TAG}_${tool}.log 2>&1
  tail -2 $O/${TAG}_${tool}.log
done
