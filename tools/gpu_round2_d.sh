#!/bin/bash
# round 2, call d: A/B of the bulk fill kernel's row-switch batching (YB_F2_SW builds), stream priorities, hardware queues
O=gpurun_out
mkdir -p $O
L=$PWD/multiz_b200
timeout 600 python tools/gpu_ab.py cfg2 1.0 "" YB_BAND_PACK=0 YB_PRIO=0 > $O/r2d_ab_base.txt 2>&1
cat $O/r2d_ab_base.txt
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 600 python tools/gpu_ab.py cfg2 1.0 "" > $O/r2d_ab_conn32.txt 2>&1
cat $O/r2d_ab_conn32.txt
YAMA_B200_LIB=$L/libyama_b200_sw2.so timeout 600 python tools/gpu_ab.py cfg2 1.0 "" > $O/r2d_ab_sw2.txt 2>&1
cat $O/r2d_ab_sw2.txt
YAMA_B200_LIB=$L/libyama_b200_sw4.so timeout 600 python tools/gpu_ab.py cfg2 1.0 "" YB_SLACK=5 > $O/r2d_ab_sw4.txt 2>&1
cat $O/r2d_ab_sw4.txt
YAMA_B200_LIB=$L/libyama_b200_sw4.so timeout 600 python -m pytest tests/test_yama_gpu.py -m gpu -x -q -k "random_small or tiny or synthetic or resident or golden" > $O/r2d_pytest_sw4.txt 2>&1
tail -3 $O/r2d_pytest_sw4.txt
