#!/bin/bash
# Development aid (run under gpurun): one 10 Mb two-species merge through the drop-in with the library's profile lines
set -e
D=$(mktemp -d)
python - <<PY
import sys; sys.path.insert(0, "$PWD")
from tools.mafsynth import make_dataset
make_dataset("$D", ref_len=10_000_000, n_species=2, seed=1)
PY
cd $D
for i in 1 2 3; do
  YB_PROFILE=1 YB_DROPIN_STATS=1 $GRAFT_REPO_ROOT/integration/_ref/bin/multiz ref.sp1.maf ref.sp2.maf 1 u1 u2 2>err.txt >out.maf
  grep -E "profile|yama_b200:" err.txt | cut -c1-420
done
