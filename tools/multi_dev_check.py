#!/usr/bin/env python
"""The drop-in's own multi-GPU mode, measured and checked (run on a multi-GPU box): ONE process, `yb_create(NULL, 0)` over all
visible devices, one `yb_run_batch` of a BASELINE workload -- waves go to whichever device is free, no collective.  The
same batch runs on device 0 alone first; every result of the all-device run must equal the one-device run's, a stratified
sample is compared with the CPU oracle, and the strong-scaling ratio is reported.

    python tools/multi_dev_check.py --workload cfg3 --scale 0.25 > gpurun_out/multidev.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--scale", type=float, default=0.25)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--sample", type=int, default=60)
    a = ap.parse_args()
    from bench import make_batch
    from multiz_b200 import YamaB200, RESULT_DTYPE
    from oracle.oracle_py import Oracle
    sb, desc = make_batch(a.workload, 1234, a.scale)
    out = {"workload": desc, "pairs": int(sb.n), "cells": int(sb.cells), "runs": []}
    keep = None
    for devices in ([0], None):
        ctx = YamaB200(devices=devices)
        jobs = ctx.pin_pools(sb.jobs, (sb.A, sb.B, sb.LB, sb.RB))
        res = np.zeros(sb.n, dtype=RESULT_DTYPE)
        walls = []
        for _ in range(a.reps):
            t0 = time.perf_counter()
            _, st = ctx.run_batch(jobs, out=res)
            walls.append((time.perf_counter() - t0) * 1e3)
        w = float(np.mean(walls[1:]))
        run = {"devices": int(st.n_devices), "ms_per_call": w, "gcups_e2e": sb.cells / w / 1e6, "failed": int((res["status"] != 0).sum()),
               "h2d_bytes": int(st.h2d_bytes), "host_prepare_ms": float(st.pack_ms)}
        fields = ("status", "m_new", "C", "D", "I", "cells")
        if keep is None:
            keep = ({f: res[f].copy() for f in fields}, ctx, res.copy())
            one = ctx
        else:
            run["equal_to_one_device"] = bool(all((keep[0][f] == res[f]).all() for f in fields))
            idx = np.unique(np.linspace(0, sb.n - 1, a.sample).astype(int))
            run["scripts_equal_on_sample"] = bool(all(np.array_equal(one.script_of(keep[2][i]), ctx.script_of(res[i])) for i in idx))
            orc = Oracle(70)
            order = np.argsort(sb.cells_per_pair(), kind="stable")
            ok = True
            for i in order[np.unique(np.linspace(0, sb.n - 1, min(a.sample, 24)).astype(int))]:
                o = orc.yama(*sb.problem(int(i)), want_tback=False)
                r = res[int(i)]
                ok &= (int(r["C"]), int(r["D"]), int(r["I"])) == tuple(int(x) for x in o["cdi"]) and np.array_equal(ctx.script_of(r), o["script"])
            run["oracle_sample_ok"] = bool(ok)
            run["speedup_over_one_device"] = out["runs"][0]["ms_per_call"] / w
            ctx.close()
        out["runs"].append(run)
    keep[1].close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
