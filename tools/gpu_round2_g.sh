#!/bin/bash
# round 2, call g: fill_body3 with mbarrier hand-shakes -- parity tests, cfg5 bench line, ncu capture
O=gpurun_out
mkdir -p $O
timeout 240 python -m pytest tests/test_yama_gpu.py -m gpu -x -q -k "wide or cfg5 or deep or golden" > $O/r2g_pytest_wide.txt 2>&1
tail -5 $O/r2g_pytest_wide.txt
timeout 240 python bench.py --workload cfg5 --no-cpu-baseline > $O/r2g_bench_cfg5.json 2> $O/r2g_bench.err
python -c "import json;d=json.load(open('$O/r2g_bench_cfg5.json'));print(d['value'],d['kernel_split_ms'],d['e2e']['value'],d['e2e']['failed_pairs'])"
NCU="timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
$NCU -k 'regex:yb_fill3_kernel' -s 3 -c 1 -f -o $O/r2g_fill3 python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu-baseline > $O/r2g_ncu_fill3.log 2>&1
tail -2 $O/r2g_ncu_fill3.log
