#!/bin/bash
# round 2, final evidence set on one B200 (every command with its own timeout); outputs in gpurun_out/r2z_*
O=gpurun_out
mkdir -p $O
L=$PWD/multiz_b200
[ -f $L/libyama_b200_sync.so ] || make -C multiz_b200/csrc sanitize > $O/r2z_make_sanitize.log 2>&1      # (same image on the box: nvcc is there)
timeout 600 python -m pytest tests -m gpu -x -q > $O/r2z_pytest_gpu.txt 2>&1
tail -4 $O/r2z_pytest_gpu.txt
timeout 300 python bench.py > $O/r2z_bench.json 2> $O/r2z_bench.err
cut -c1-300 $O/r2z_bench.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2z_bench_reference.json 2>> $O/r2z_bench.err
timeout 240 python bench.py --workload cfg2real --steps 5 --no-cpu-baseline > $O/r2z_bench_cfg2real.json 2> $O/r2z_bench_cfg2real.err
cut -c1-200 $O/r2z_bench_cfg2real.json; tail -2 $O/r2z_bench_cfg2real.err
timeout 200 python bench.py --workload cfg5 --no-cpu-baseline > $O/r2z_bench_cfg5.json 2>> $O/r2z_bench.err
NCU="timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
$NCU -k 'regex:yb_fill2_kernel<.int.128' -s 3 -c 1 -f -o $O/r2z_fill2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2z_ncu_fill2.log 2>&1
$NCU -k 'regex:yb_fill_kernel_w<.int.1024' -s 3 -c 1 -f -o $O/r2z_fill_cta python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu-baseline > $O/r2z_ncu_fill_cta.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2z_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2z_ncu_bench.log 2>&1
YAMA_B200_LIB=$L/libyama_b200_sync.so timeout 400 compute-sanitizer --tool racecheck --print-limit 100 python tools/sanitize_probe.py > $O/r2z_racecheck_sync.log 2>&1
tail -3 $O/r2z_racecheck_sync.log
timeout 300 compute-sanitizer --tool racecheck --print-limit 100 python tools/sanitize_probe.py > $O/r2z_racecheck_shipped.log 2>&1
tail -3 $O/r2z_racecheck_shipped.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_probe.py > $O/r2z_memcheck.log 2>&1
tail -2 $O/r2z_memcheck.log
