#!/usr/bin/env python
"""BASELINE.json configs[0]/[1] at the tool level: the reference's `multiz` binary (CPU, oracle/_ref/bin) and the
same host linked against libyama_b200.so (integration/_ref/bin/multiz) run the same progressive merge

    acc = multiz ref.sp1.maf ref.sp2.maf 1 ;  acc = multiz acc ref.sp3.maf 1 ; ...

on synthetic MAFs (tools/mafsynth.py).  Every output byte must agree; wall times and the drop-in's own
statistics line are reported as one JSON object.  Run on the GPU box:

    python tools/pipeline_bench.py --ref-len 10000000 --species 5 [--v0]
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.mafsynth import make_dataset  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "bin", "multiz")
GPU = os.path.join(ROOT, "integration", "_ref", "bin", "multiz")


def run(tool, argv, cwd, env=None):
    e = dict(os.environ, **(env or {}))
    t0 = time.perf_counter()
    p = subprocess.run([tool] + argv, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
    dt = time.perf_counter() - t0
    if p.returncode != 0:
        raise SystemExit(f"{tool} {argv} failed: {p.stderr.decode()[-400:]}")
    return p.stdout, p.stderr.decode(), dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-len", type=int, default=1_000_000)
    ap.add_argument("--species", type=int, default=5)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--v0", action="store_true", help="also run the two-species merge with v=0")
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--server", action="store_true",
                    help="also time the drop-in with the resident server (yama_b200d, YB_SERVER) as its backend")
    ap.add_argument("--skip-reference", action="store_true", help="do not run the reference binary (no identity check)")
    a = ap.parse_args()
    tmp = tempfile.mkdtemp(prefix="yb_pipe_")
    da, db = os.path.join(tmp, "ref"), os.path.join(tmp, "gpu")
    t0 = time.perf_counter()
    make_dataset(da, ref_len=a.ref_len, n_species=a.species - 1, seed=a.seed)
    shutil.copytree(da, db)
    out = {"ref_len": a.ref_len, "species": a.species, "synth_s": round(time.perf_counter() - t0, 2), "steps": []}
    srv_env = None
    if a.server:
        srv_env = {"YB_DROPIN_STATS": "1", "YB_SERVER": os.path.join(tmp, "yb.sock"), "YB_SERVER_IDLE_S": "30"}
        if os.environ.get("YB_PROFILE"):
            srv_env["YB_PROFILE"] = "1"
        dc = os.path.join(tmp, "srv")
        shutil.copytree(da, dc)
        # first contact starts the server (CUDA start-up is paid here, once)
        t0 = time.perf_counter()
        run(GPU, ["ref.sp1.maf", "ref.sp2.maf", "1", "w1", "w2"], dc, srv_env)
        out["server_first_call_s"] = round(time.perf_counter() - t0, 2)
    acc = "ref.sp1.maf"
    plans = [("v1", i) for i in range(2, a.species)]
    for mode, i in plans:
        argv = [acc, f"ref.sp{i}.maf", "1", f"u1.{i}", f"u2.{i}"]
        go, ge, t_gpu = run(GPU, argv, db, {"YB_DROPIN_STATS": "1"})
        so, t_ref = (go, float("nan")) if a.skip_reference else run(REF, argv, da)[::2]
        if a.skip_reference:
            for f in (f"u1.{i}", f"u2.{i}"):
                shutil.copy(os.path.join(db, f), os.path.join(da, f))
        same = so == go and all(open(os.path.join(da, f), "rb").read() == open(os.path.join(db, f), "rb").read()
                                for f in (f"u1.{i}", f"u2.{i}"))
        acc = f"acc{i}.maf"
        for d, o in ((da, so), (db, go)):
            open(os.path.join(d, acc), "wb").write(o)
        stats = ge.strip().splitlines()[-1] if ge.strip() else ""
        out["steps"].append({"step": f"multiz(acc, ref.sp{i}, v=1)", "byte_identical": same, "reference_s": round(t_ref, 2),
                             "b200_s": round(t_gpu, 2), "speedup": round(t_ref / t_gpu, 2), "output_bytes": len(so),
                             "dropin": stats})
        if srv_env:
            if i > 2:
                open(os.path.join(dc, f"acc{i - 1}.maf"), "wb").write(open(os.path.join(db, f"acc{i - 1}.maf"), "rb").read())
            po, pe, t_srv = run(GPU, argv, dc, srv_env)
            out["steps"][-1].update({"b200_server_s": round(t_srv, 2), "server_identical": po == go,
                                     "dropin_server": pe.strip().splitlines()[-1] if pe.strip() else ""})
        if not same:
            break
    if a.v0:
        argv = ["ref.sp1.maf", "ref.sp2.maf", "0", "z1", "z2"]
        so, _, t_ref = run(REF, argv, da)
        go, ge, t_gpu = run(GPU, argv, db, {"YB_DROPIN_STATS": "1"})
        out["steps"].append({"step": "multiz(ref.sp1, ref.sp2, v=0)", "byte_identical": so == go, "reference_s": round(t_ref, 2),
                             "b200_s": round(t_gpu, 2), "speedup": round(t_ref / t_gpu, 2),
                             "dropin": ge.strip().splitlines()[-1] if ge.strip() else ""})
    out["all_identical"] = all(s["byte_identical"] for s in out["steps"])
    out["reference_total_s"] = round(sum(s["reference_s"] for s in out["steps"]), 2)
    out["b200_total_s"] = round(sum(s["b200_s"] for s in out["steps"]), 2)
    if srv_env:
        out["b200_server_total_s"] = round(sum(s.get("b200_server_s", 0.0) for s in out["steps"]), 2)
    print(json.dumps(out, indent=1))
    if not a.keep:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
