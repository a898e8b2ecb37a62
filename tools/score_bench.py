"""Throughput of the block-scoring path (yb_score_blocks == mafScoreRange over a batch of blocks), next to the
reference's own mafScoreRange on one host core.  Not the headline benchmark (bench.py); evidence for SURVEY 8(f) rank 1.

    python tools/score_bench.py [shallow|deep|all] [--json out.json]
Units: text GB/s (rows x columns bytes scored per second) and G pair-columns/s (the reference's work: row pairs x
columns).  `kernel` is the device time of yb_score_kernel (CUDA events around the launches), `e2e` the wall time of
the ABI call from host buffers (pack + H2D + kernel + D2H).
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.score_cases import alignment_block  # noqa: E402


def workload(kind, rng):
    if kind == "shallow":      # what a 5-way progressive merge writes: many blocks, 2..6 rows, a few hundred columns
        texts = [alignment_block(rng, int(rng.integers(2, 7)), int(rng.integers(60, 900))) for _ in range(2000)]
        reps = 60
    else:                      # deep alignments (roast, tens of species): 100 rows, 10^4 columns
        texts = [alignment_block(rng, 100, 10000, gap_open=0.01) for _ in range(8)]
        reps = 50
    return [(t, 0, t.shape[1]) for t in texts] * reps


def main():
    kinds = ["shallow", "deep"] if len(sys.argv) < 2 or sys.argv[1] == "all" else [sys.argv[1]]
    from multiz_b200 import YamaB200
    ctx = YamaB200(devices=[0])
    out = []
    for kind in kinds:
        rng = np.random.default_rng(11)
        cases = workload(kind, rng)
        blocks, keep = ctx.make_blocks(cases)
        nbytes = sum(t.size for t, _, _ in cases)
        for _ in range(3):
            sc, st = ctx.score_blocks(blocks)
        ks, es = [], []
        for _ in range(5):
            t0 = time.perf_counter()
            sc, st = ctx.score_blocks(blocks)
            es.append(time.perf_counter() - t0)
            ks.append(st.kernel_ms / 1e3)
        k, e = float(np.median(ks)), float(np.median(es))
        rec = {"workload": kind, "blocks": len(cases), "text_bytes": nbytes, "pair_columns": int(st.cells),
               "kernel_ms": k * 1e3, "e2e_ms": e * 1e3, "kernel_text_GBps": nbytes / k / 1e9, "e2e_text_GBps": nbytes / e / 1e9,
               "kernel_Gpaircols": st.cells / k / 1e9, "e2e_Gpaircols": st.cells / e / 1e9,
               "pack_ms": st.pack_ms, "h2d_ms": st.h2d_ms, "launches": int(st.kernel_launches)}
        # the reference (or the oracle port) on one host core, bounded sample
        try:
            from oracle.oracle_py import Oracle, Reference
            cpu = Reference(70) if Reference.available() else Oracle(70)
            kindc = "reference" if Reference.available() else "port"
            n, work, t0 = 0, 0, time.perf_counter()
            for (t, s, z) in cases:
                v = cpu.score_range(t, s, z)
                assert v == sc[n], (n, v, sc[n])
                work += t.shape[0] * (t.shape[0] - 1) // 2 * z
                n += 1
                if time.perf_counter() - t0 > 5:
                    break
            dt = time.perf_counter() - t0
            rec["cpu"] = {"kind": kindc, "cores": 1, "blocks": n, "Gpaircols": work / dt / 1e9}
        except Exception as ex:  # the checker is optional here
            rec["cpu"] = {"error": str(ex)}
        out.append(rec)
        print(json.dumps(rec))
        del keep
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(out, f, indent=1)
    ctx.close()


if __name__ == "__main__":
    main()
