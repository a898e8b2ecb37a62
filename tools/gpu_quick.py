"""Quick GPU sanity + throughput probe (development aid; run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiz_b200 import YamaB200
from tools.synth import SynthBatch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rng = np.random.default_rng(0)
Ks = rng.integers(2, 6, size=n); Ls = np.ones(n, dtype=np.int32); Ms = rng.integers(100, 900, size=n)
t0 = time.time(); sb = SynthBatch(1, Ks, Ls, Ms, R=R); print("synth s", time.time() - t0, "cells", sb.cells)
ctx = YamaB200(devices=[0])
for it in range(3):
    res, st = ctx.run_batch(sb.jobs)
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.as_dict().items()})
    print("GCUPS kernel", sb.cells / st.kernel_ms / 1e6, "fill-only", sb.cells / st.fill_ms / 1e6, "e2e", sb.cells / st.total_ms / 1e6)
ctx.resident_load(sb.jobs)
for it in range(3):
    st = ctx.resident_step()
    print("resident GCUPS", sb.cells / st.kernel_ms / 1e6, "fill", st.fill_ms, "prof", st.profile_ms, "tb", st.traceback_ms)
