"""Development aid (run under gpurun): many wide-band pairs per CTA of the wide-band fill kernel, with a short timeout."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SNIPPET = """
import sys, time; sys.path.insert(0, {root!r})
import numpy as np
from multiz_b200 import YamaB200
from tools.synth import SynthBatch
from oracle.oracle_py import Oracle
n = {n}
sb = SynthBatch(7, [2] * n, [1] * n, [520] * n, R=300, indel=0.02)
ctx = YamaB200(devices=[0])
for it in range(3):
    t0 = time.perf_counter()
    res, st = ctx.run_batch(sb.jobs)
    print("n", n, "run", it, "wall %.1f ms" % ((time.perf_counter() - t0) * 1e3), "fill %.3f ms" % st.fill_ms, "cells", st.cells, "failed", int((res["status"] != 0).sum()), flush=True)
orc = Oracle(70)
ok = True
for i in (0, 1, n // 2, n - 1):
    o = orc.yama(*sb.problem(i), want_tback=False)
    r = res[i]
    ok &= r["status"] == 0 and (int(r["C"]), int(r["D"]), int(r["I"])) == tuple(int(x) for x in o["cdi"]) and np.array_equal(ctx.script_of(r), o["script"])
print("parity", bool(ok), flush=True)
"""
for lib in sys.argv[1:] or [""]:
    for n in (100, 1500):
        env = dict(os.environ)
        if lib:
            env["YAMA_B200_LIB"] = os.path.join(ROOT, "multiz_b200", lib)
        try:
            p = subprocess.run([sys.executable, "-c", SNIPPET.format(root=ROOT, n=n)], capture_output=True, text=True, timeout=60, env=env)
            print("[%s]" % (lib or "default"), p.stdout.strip().replace("\n", " | ") or ("rc=%d %s" % (p.returncode, p.stderr.strip()[-300:])), flush=True)
        except subprocess.TimeoutExpired as e:
            print("[%s]" % (lib or "default"), "n", n, "TIMEOUT", (e.stdout or b"").decode()[-300:] if isinstance(e.stdout, bytes) else e.stdout, flush=True)
