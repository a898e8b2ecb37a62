#!/bin/bash
# round 2, call e: the wide-band kernel (fill_body3) -- parity tests, cfg5 bench line, then the rest of the GPU suite
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_yama_gpu.py -m gpu -x -q -k "wide or cfg5 or deep or golden" > $O/r2e_pytest_wide.txt 2>&1
tail -15 $O/r2e_pytest_wide.txt
timeout 600 python bench.py --workload cfg5 --no-cpu-baseline > $O/r2e_bench_cfg5.json 2> $O/r2e_bench.err
cat $O/r2e_bench_cfg5.json | cut -c1-1400
tail -3 $O/r2e_bench.err
timeout 600 python bench.py --no-cpu-baseline > $O/r2e_bench.json 2>> $O/r2e_bench.err
cat $O/r2e_bench.json | cut -c1-2200
