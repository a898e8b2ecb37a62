#!/bin/bash
# Development aid (run on the GPU box): the 10 Mb two-species merge through the drop-in, v=1 and v=0, batch vs
# streamed replay, with a context per invocation and behind the resident server; outputs compared byte for byte.
cd /root/repo
python - <<'PY'
import sys; sys.path.insert(0,'/root/repo')
from tools.mafsynth import make_dataset
make_dataset('/tmp/ds', ref_len=int(__import__("os").environ.get("REF_LEN", "10000000")), n_species=2, seed=3, lower=0.01)
PY
cd /tmp/ds
export YB_DROPIN_STATS=1
/root/repo/integration/_ref/bin/multiz ref.sp1.maf ref.sp2.maf 1 u1 u2 >/dev/null 2>&1   # warm the box
for v in 1 0; do
 for srv in "" auto; do
  for m in batch stream stream; do
    YB_SERVER=$srv YB_DROPIN=$m /root/repo/integration/_ref/bin/multiz ref.sp1.maf ref.sp2.maf $v u1 u2 2>/tmp/err.txt >/tmp/out.$m.$v
    echo "v=$v server=${srv:-no} $m $(grep -o 'passes=[0-9]* batches=[0-9]*' /tmp/err.txt) $(grep -o 'misses=[0-9]*' /tmp/err.txt) $(grep -o 'gpu_ms.*stream_waits=[0-9]*' /tmp/err.txt | sed 's/score_calls=0 score_ms=0.0 //')"
  done
 done
 cmp /tmp/out.batch.$v /tmp/out.stream.$v && echo "v=$v same"
done
for j in 32 512; do YB_STREAM_MIN_JOBS=$j YB_SERVER=auto YB_DROPIN=stream /root/repo/integration/_ref/bin/multiz ref.sp1.maf ref.sp2.maf 0 u1 u2 2>/tmp/err.txt >/dev/null; echo "minjobs=$j $(grep -o 'passes=[0-9]* batches=[0-9]*' /tmp/err.txt) $(grep -o 'wall_ms=[0-9]*' /tmp/err.txt)"; done
