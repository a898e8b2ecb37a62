"""Development aid: time the resident kernels of several builds of the library (YAMA_B200_LIB)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
scale = sys.argv[2] if len(sys.argv) > 2 else "1.0"
libs = sys.argv[3:] or ["multiz_b200/libyama_b200.so"]
code = r'''
import sys, os
sys.path.insert(0, %r)
from multiz_b200 import YamaB200
from bench import make_batch
sb, _ = make_batch(%r, 1234, float(%r))
ctx = YamaB200(devices=[0]); ctx.resident_load(sb.jobs)
for it in range(4):
    st = ctx.resident_step()
print(os.environ.get("YAMA_B200_LIB"), "GCUPS", round(sb.cells / st.kernel_ms / 1e6, 1), "fill", round(st.fill_ms, 3), "prof", round(st.profile_ms, 3), "tb", round(st.traceback_ms, 3))
''' % (ROOT, wl, scale)
for lib in libs:
    env = dict(os.environ, YAMA_B200_LIB=os.path.join(ROOT, lib))
    subprocess.run([sys.executable, "-c", code], env=env)
