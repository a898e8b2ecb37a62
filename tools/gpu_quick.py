"""Quick GPU sanity + throughput probe (development aid; run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiz_b200 import YamaB200
from bench import make_batch

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
waves = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [64]
t0 = time.time(); sb, desc = make_batch(wl, 1234, scale); print("synth s", round(time.time() - t0, 2), "pairs", sb.n, "cells", sb.cells)
for w in waves:
    os.environ["YB_WAVE_MB"] = str(w)
    ctx = YamaB200(devices=[0])
    for it in range(4):
        t0 = time.perf_counter()
        res, st = ctx.run_batch(sb.jobs)
        wall = (time.perf_counter() - t0) * 1e3
        d = st.as_dict()
        if it >= 2:
            print(f"wave {w} MB: wall {wall:.1f} ms  e2e {sb.cells / wall / 1e6:.1f} GCUPS | kernel {d['kernel_ms']:.1f} (fill {d['fill_ms']:.1f} prof {d['profile_ms']:.1f} tb {d['traceback_ms']:.1f}) h2d {d['h2d_ms']:.1f} d2h {d['d2h_ms']:.1f} pack+unpack {d['pack_ms']:.1f} | h2d {d['h2d_bytes'] / 1e6:.0f} MB d2h {d['d2h_bytes'] / 1e6:.0f} MB launches {d['kernel_launches']}")
    ctx.close()
ctx = YamaB200(devices=[0])
ctx.resident_load(sb.jobs)
for it in range(3):
    st = ctx.resident_step()
    print("resident GCUPS", round(sb.cells / st.kernel_ms / 1e6, 1), "fill", round(st.fill_ms, 3), "prof", round(st.profile_ms, 3), "tb", round(st.traceback_ms, 3))
