#!/bin/bash
# round 2, call i: wide-band bins back on fill_body, 1 024-entry ring (4 CTAs per SM)
O=gpurun_out
mkdir -p $O
timeout 200 python -m pytest tests/test_yama_gpu.py -m gpu -x -q -k "wide or cfg5 or deep or golden" > $O/r2i_pytest_wide.txt 2>&1
tail -3 $O/r2i_pytest_wide.txt
for s in 1.0; do
  timeout 150 python bench.py --workload cfg5 --scale $s --steps 5 --warmup 2 --no-cpu-baseline > $O/r2i_bench_cfg5_$s.json 2>/dev/null
  python -c "import json,sys
try:
    d=json.load(open('$O/r2i_bench_cfg5_$s.json')); print('scale', '$s', d['config']['pairs_per_gpu'], d['value'], d['kernel_split_ms'], 'e2e', d['e2e']['value'], d['e2e']['failed_pairs'])
except Exception as e: print('scale', '$s', 'no line', e)"
done
