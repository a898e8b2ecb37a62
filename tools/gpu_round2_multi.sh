#!/bin/bash
# round 2, the multi-GPU call (gpurun --gpus 8): the in-process all-device path (test + cfg3 strong scaling with oracle sample),
# then one rank per GPU under torchrun on cfg2 and cfg5.  Every command has its own timeout.
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r2m_gpus.txt 2>&1
timeout 150 python -m pytest tests/test_yama_gpu.py -m gpu -x -q -k "two_devices" > $O/r2m_pytest_two_devices.txt 2>&1
tail -3 $O/r2m_pytest_two_devices.txt
timeout 200 python tools/multi_dev_check.py --workload cfg3 --scale 0.25 > $O/r2m_multidev_cfg3.json 2> $O/r2m_multidev.err
cat $O/r2m_multidev_cfg3.json | tr -d '\n' | cut -c1-1200; echo
N=$(nvidia-smi -L | wc -l)
for wl in cfg2 cfg5; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > $O/r2m_bench_${wl}_n$N.json 2> $O/r2m_bench_${wl}.err
  python -c "import json
try:
    d=json.load(open('$O/r2m_bench_${wl}_n$N.json')); print('$wl', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['limiter_ms'], d['e2e']['failed_pairs'])
except Exception as e: print('$wl', 'no line', e)"
done
