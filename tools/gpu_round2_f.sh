#!/bin/bash
# round 2, call f: fill_body3 variants on cfg5 (spin without sleep, a lag of two groups between warps) + ncu of the shipped one
O=gpurun_out
mkdir -p $O
L=$PWD/multiz_b200
for v in f3a f3b; do
  YAMA_B200_LIB=$L/libyama_b200_$v.so timeout 600 python bench.py --workload cfg5 --no-cpu-baseline > $O/r2f_bench_cfg5_$v.json 2> $O/r2f_bench_$v.err
  python -c "import json;d=json.load(open('$O/r2f_bench_cfg5_$v.json'));print('$v',d['value'],d['kernel_split_ms'],d['e2e']['value'],d['e2e']['failed_pairs'])"
done
YAMA_B200_LIB=$L/libyama_b200_f3b.so timeout 900 python -m pytest tests/test_yama_gpu.py -m gpu -x -q -k "wide or cfg5 or deep or golden" > $O/r2f_pytest_f3b.txt 2>&1
tail -3 $O/r2f_pytest_f3b.txt
NCU="timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
$NCU -k 'regex:yb_fill3_kernel' -s 3 -c 1 -f -o $O/r2f_fill3 python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu-baseline > $O/r2f_ncu_fill3.log 2>&1
tail -2 $O/r2f_ncu_fill3.log
