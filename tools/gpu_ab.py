"""Development aid (run under gpurun): A/B of fill-kernel variants selected by environment variables, each checked
against the oracle on a sample before it is timed.   python tools/gpu_ab.py [workload] [scale] VAR=VAL,VAR=VAL ..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import make_batch
from tools.synth import SynthBatch, random_problem

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
variants = sys.argv[3:] or [""]
from oracle.oracle_py import Oracle
orc = Oracle(70)
sb, desc = make_batch(wl, 1234, scale)
rng = np.random.default_rng(7)
# parity sample: synthetic pre_yama-like pairs + adversarial random bands, a few hundred pairs, oracle time ~ seconds
Ks = [2, 3, 4, 5, 2, 8, 1, 12] * 12
Ls = [1, 1, 1, 1, 2, 3, 1, 2] * 12
Ms = list(rng.integers(1, 900, len(Ks)))
chk = SynthBatch(4242, Ks, Ls, Ms, R=30, lower=0.03)
probs = [tuple(np.array(x) for x in chk.problem(i)) for i in range(chk.n)]
for it in range(150):
    K, L = int(rng.integers(1, 9)), int(rng.integers(1, 6))
    M, N = int(rng.integers(1, 120)), int(rng.integers(1, 120))
    band = ("smooth", "full", "ragged")[it % 3]
    if band == "full" and M * N > 3000:
        M = max(1, 3000 // N)
    probs.append(random_problem(rng, K, L, M, N, band=band, alphabet=("acgt", "mixed", "weird")[it % 3]))
want = [orc.yama(*p, want_tback=False) for p in probs]
from multiz_b200 import YamaB200
for var in variants:
    keys = []
    for kv in var.split(","):
        if "=" in kv:
            k, v = kv.split("=", 1); os.environ[k] = v; keys.append(k)
    ctx = YamaB200(devices=[0])
    jobs, keep = ctx.make_jobs(probs)
    res, st = ctx.run_batch(jobs)
    bad = 0
    for i, o in enumerate(want):
        r = res[i]
        ok = (r["status"] == 0 and (int(r["C"]), int(r["D"]), int(r["I"])) == tuple(int(x) for x in o["cdi"])
              and int(r["m_new"]) == o["m_new"] and np.array_equal(ctx.script_of(r), o["script"]))
        bad += 0 if ok else 1
        if not ok and bad <= 3:
            print("  MISMATCH pair", i, probs[i][0].shape, probs[i][1].shape, "status", int(r["status"]), "cdi", (int(r["C"]), int(r["D"]), int(r["I"])), "want", list(o["cdi"]), "m_new", int(r["m_new"]), o["m_new"])
    ctx.resident_load(sb.jobs)
    fills = []
    for it in range(6):
        st = ctx.resident_step()
        if it >= 2:
            fills.append((st.fill_ms, st.profile_ms, st.traceback_ms, st.kernel_ms))
    f = np.array(fills).mean(axis=0)
    rr = ctx.resident_fetch()
    nfail = int((rr["status"] != 0).sum())
    csum = int(rr["C"].astype(np.int64).sum()), int(rr["m_new"].astype(np.int64).sum())
    print(f"[{var or 'default'}] parity {'OK' if bad == 0 else 'FAIL %d/%d' % (bad, len(want))} | fill {f[0]:.3f} ms = {sb.cells / f[0] / 1e6:.1f} GCUPS | prof {f[1]:.3f} tb {f[2]:.3f} all {f[3]:.3f} ms = {sb.cells / f[3] / 1e6:.1f} GCUPS | failed {nfail} checksum {csum}", flush=True)
    # end to end through yb_run_batch: inputs in ordinary memory (one staging copy on the host), then in yb_host_alloc memory
    resbuf = np.zeros(len(sb.jobs), dtype=res.dtype)
    for tag, jb in (("staged", sb.jobs), ("pinned", ctx.pin_pools(sb.jobs, (sb.A, sb.B, sb.LB, sb.RB)))):
        walls = []
        for it in range(5):
            t0 = time.perf_counter()
            r2, st2 = ctx.run_batch(jb, out=resbuf, check=False)
            walls.append((time.perf_counter() - t0) * 1e3)
        w = float(np.mean(walls[2:]))
        same = bool((r2["C"] == rr["C"]).all() and (r2["m_new"] == rr["m_new"]).all() and (r2["status"] == 0).all())
        dd = st2.as_dict()
        print(f"   e2e {tag}: {w:.2f} ms = {sb.cells / w / 1e6:.1f} GCUPS | host prepare {dd['pack_ms']:.2f} h2d {dd['h2d_ms']:.2f} plan {dd['plan_ms']:.2f} kernels {dd['kernel_ms']:.2f} d2h {dd['d2h_ms']:.2f} | "
              f"h2d {dd['h2d_bytes'] / 1e6:.0f} MB staged {dd['staged_bytes'] / 1e6:.0f} MB | same results {same}", flush=True)
    ctx.close()
    for k in keys:
        del os.environ[k]
