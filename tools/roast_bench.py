#!/usr/bin/env python
"""BASELINE.json configs[3] scaled to what a benchmark box finishes in minutes: the reference's own `roast` driver
(auto_mz.c) aligns a synthetic species tree by exec'ing `multiz` and `maf_project` from PATH once per tree node.  Three
arms on identical inputs: the reference's multiz (CPU), the drop-in (one CUDA context per multiz invocation), and the
drop-in behind the resident server (yama_b200d, one context for the whole tree).  The outputs must agree byte for
byte apart from '#' provenance lines.

    python tools/roast_bench.py --ref-len 2000000 --species 8            (on the GPU box)
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dropin_util import GPU_MULTIZ, GPU_SERVER, REF_MULTIZ, make_roast_dataset, run_roast, server_env, stop_server  # noqa: E402

TREES = {
    4: "((ref sp1) (sp2 sp3))",
    5: "((ref sp1) ((sp2 sp3) sp4))",
    8: "((((ref sp1) sp2) (sp3 sp4)) ((sp5 sp6) sp7))",
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-len", type=int, default=2_000_000)
    ap.add_argument("--species", type=int, default=8, choices=sorted(TREES))
    ap.add_argument("--seed", type=int, default=4)
    ap.add_argument("--skip-reference", action="store_true")
    a = ap.parse_args()
    tmp = tempfile.mkdtemp(prefix="yb_roast_")
    base = os.path.join(tmp, "data")
    t0 = time.perf_counter()
    make_roast_dataset(base, a.ref_len, a.species - 1, seed=a.seed)
    out = {"ref_len": a.ref_len, "species": a.species, "tree": TREES[a.species], "synth_s": round(time.perf_counter() - t0, 2)}
    outputs = {}

    def arm(name, tool, env=None):
        d = os.path.join(tmp, name)
        shutil.copytree(base, d)
        t = time.perf_counter()
        outputs[name] = run_roast(tool, d, TREES[a.species], env=env)
        out[name + "_s"] = round(time.perf_counter() - t, 2)
        shutil.rmtree(d, ignore_errors=True)

    arm("b200", GPU_MULTIZ)
    senv = server_env(__import__("pathlib").Path(tmp), GPU_SERVER, idle_s=30)
    try:
        arm("b200_server", GPU_MULTIZ, senv)
    finally:
        stop_server(senv)
    if not a.skip_reference:
        arm("reference", REF_MULTIZ)
        out["byte_identical"] = outputs["reference"] == outputs["b200"] == outputs["b200_server"]
        out["speedup"] = round(out["reference_s"] / out["b200_s"], 2)
        out["speedup_server"] = round(out["reference_s"] / out["b200_server_s"], 2)
    else:
        out["byte_identical_between_gpu_arms"] = outputs["b200"] == outputs["b200_server"]
    out["output_bytes"] = len(outputs["b200"])
    print(json.dumps(out, indent=1))
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
