#!/bin/bash
# round 2, call c: delta-coded bands (parity test, A/B of the end-to-end call), per-wave timeline, 10 Mb tool-level pipeline
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_yama_gpu.py -m gpu -x -q -k "delta or random_small or batch_api or resident or tiny" > $O/r2c_pytest.txt 2>&1
tail -5 $O/r2c_pytest.txt
timeout 900 python tools/gpu_ab.py cfg2 1.0 YB_BAND_PACK=0 YB_BAND_PACK=1 > $O/r2c_ab_cfg2.txt 2>&1
cat $O/r2c_ab_cfg2.txt
timeout 900 python tools/gpu_ab.py cfg3 0.25 YB_BAND_PACK=0 YB_BAND_PACK=1 > $O/r2c_ab_cfg3.txt 2>&1
grep -E "^\[|e2e" $O/r2c_ab_cfg3.txt
YB_PROFILE=2 timeout 600 python tools/gpu_ab.py cfg2 1.0 YB_BAND_PACK=1 > $O/r2c_timeline.txt 2>&1
timeout 900 python tools/pipeline_bench.py --ref-len 10000000 --species 5 --v0 > $O/r2c_pipeline_10mb.json 2> $O/r2c_pipeline.err
cat $O/r2c_pipeline_10mb.json | head -60
