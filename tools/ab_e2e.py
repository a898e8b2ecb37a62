"""Development aid: end-to-end time of yb_run_batch for several builds of the library, alternating (same box, same call):
    python tools/ab_e2e.py cfg2 variants/lib_a.so variants/lib_b.so
"""
import os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wl = sys.argv[1]
code = r'''
import sys, os, time
sys.path.insert(0, %r)
import numpy as np
from multiz_b200 import YamaB200
from multiz_b200.yama import RESULT_DTYPE
from bench import make_batch
sb, _ = make_batch(%r, 1234, 1.0)
ctx = YamaB200(devices=[0])
res = np.zeros(len(sb.jobs), dtype=RESULT_DTYPE)
ts = []
for it in range(9):
    t0 = time.perf_counter(); ctx.run_batch(sb.jobs, out=res); ts.append((time.perf_counter() - t0) * 1e3)
ts = sorted(ts[2:])
print(os.path.basename(os.environ["YAMA_B200_LIB"]), "median %%.2f ms best %%.2f -> %%.1f GCUPS" %% (ts[len(ts)//2], ts[0], sb.cells / ts[len(ts)//2] / 1e6))
''' % (ROOT, wl)
for lib in sys.argv[2:] * 2:
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, YAMA_B200_LIB=os.path.join(ROOT, lib)))
