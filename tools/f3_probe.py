"""Development aid (run under gpurun): each wide-band case of the GPU suite in a process of its own with a short timeout,
so that a kernel that does not terminate costs seconds, not the call's limit."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [(2, 1, 1200, 400), (3, 1, 255, 400), (2, 2, 513, 380), (2, 1, 2300, 1100), (1, 1, 1030, 1500), (2, 1, 600, 150), (99, 1, 3000, 300)]
SNIPPET = """
import sys; sys.path.insert(0, {root!r})
import numpy as np
from multiz_b200 import YamaB200
from tools.synth import SynthBatch
from oracle.oracle_py import Oracle
K, L, M, R = {case}
sb = SynthBatch(40 + M, [K, K], [L, L], [M, max(1, M - 37)], R=R, indel=0.02)
ctx = YamaB200(devices=[0])
res, st = ctx.run_batch(sb.jobs)
orc = Oracle(70)
ok = True
for i in range(sb.n):
    o = orc.yama(*sb.problem(i), want_tback=False)
    r = res[i]
    ok &= r["status"] == 0 and (int(r["C"]), int(r["D"]), int(r["I"])) == tuple(int(x) for x in o["cdi"]) and np.array_equal(ctx.script_of(r), o["script"])
print("case", (K, L, M, R), "parity", bool(ok), "kernel_ms %.3f" % st.kernel_ms, flush=True)
"""
for case in CASES:
    try:
        p = subprocess.run([sys.executable, "-c", SNIPPET.format(root=ROOT, case=case)], capture_output=True, text=True, timeout=float(os.environ.get("F3_TIMEOUT", "45")))
        print(p.stdout.strip() or ("case %s rc=%d %s" % (case, p.returncode, p.stderr.strip()[-300:])), flush=True)
    except subprocess.TimeoutExpired:
        print("case", case, "TIMEOUT", flush=True)
