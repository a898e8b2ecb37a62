"""Development aid (run under gpurun): end-to-end time of the real merge's batches (cfg2real, one yb_run_batch per merge step) under
different wave sizes.   python tools/gpu_wave_ab.py "YB_WAVE_MIN_MB=4,YB_WAVE_TAIL_MB=4,YB_WAVE_MB=16" ..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import record_real_merge
from multiz_b200 import YamaB200, RESULT_DTYPE

batches = record_real_merge(1.0, 1)
for var in [""] + sys.argv[1:]:
    keys = []
    for kv in var.split(","):
        if "=" in kv:
            k, v = kv.split("=", 1); os.environ[k] = v; keys.append(k)
    ctx = YamaB200(devices=[0])
    line = []
    for sb in batches:
        res = np.zeros(len(sb.jobs), dtype=RESULT_DTYPE)
        pinned = ctx.pin_pools(sb.jobs, (sb.A, sb.B, sb.LB, sb.RB))
        walls = []
        for it in range(7):
            t0 = time.perf_counter()
            _, st = ctx.run_batch(pinned, out=res)
            walls.append((time.perf_counter() - t0) * 1e3)
        line.append("%.2f ms (%d waves)" % (float(np.mean(walls[2:])), st.waves if hasattr(st, "waves") else -1))
    print("[%s]" % (var or "default"), " | ".join(line), flush=True)
    ctx.close()
    for k in keys:
        del os.environ[k]
