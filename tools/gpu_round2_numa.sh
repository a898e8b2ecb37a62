#!/bin/bash
# round 2: one rank per GPU with each rank on its GPU's NUMA node (bench.py bind_to_gpu_numa_node), cfg2 at N = all GPUs
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/r2n_topo.txt 2>&1
lscpu | grep -i -E "numa|socket|^CPU\(s\)" > $O/r2n_lscpu.txt 2>&1
N=$(nvidia-smi -L | wc -l)
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline > $O/r2n_bench_cfg2_n$N.json 2> $O/r2n_bench_cfg2.err
python -c "import json
try:
    d=json.load(open('$O/r2n_bench_cfg2_n$N.json')); print('cfg2', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['limiter_ms'], d['config'].get('numa_node_of_rank0'))
except Exception as e: print('no line', e)"
cat $O/r2n_lscpu.txt
