"""Small batch through every kernel bin, for compute-sanitizer (memcheck / racecheck / initcheck) runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiz_b200 import YamaB200
from tools.synth import SynthBatch, random_problem

ctx = YamaB200(devices=[0])
rng = np.random.default_rng(1)
probs = [random_problem(rng, int(rng.integers(1, 6)), int(rng.integers(1, 4)), int(rng.integers(1, 80)), int(rng.integers(1, 80)),
                        band=("smooth", "full", "ragged")[i % 3]) for i in range(60)]
for (K, L, M, R) in ((2, 1, 300, 30), (3, 1, 150, 150), (2, 1, 400, 150), (2, 1, 700, 400), (1, 1, 1200, 1100), (90, 10, 70, 30)):
    sb = SynthBatch(K * 7 + M, [K, K], [L, L], [M, max(1, M - 13)], R=R)
    probs += [tuple(np.array(x) for x in sb.problem(i)) for i in range(sb.n)]
jobs, keep = ctx.make_jobs(probs)
res, st = ctx.run_batch(jobs)
assert int((res["status"] != 0).sum()) == 0
print("pairs", len(jobs), "cells", st.cells, "launches", st.kernel_launches, "checksum", int(res["m_new"].sum()), int(res["C"].astype(np.int64).sum()))
# a wave of small pairs with connected bands only: the bulk kernels' variant without existence multipliers (GATED = false)
sb = SynthBatch(99, [2, 3, 4, 5, 2, 3] * 4, [1, 1, 1, 1, 2, 1] * 4, [40, 70, 130, 260, 33, 64] * 4, R=30)
res2, st2 = ctx.run_batch(sb.jobs)
assert int((res2["status"] != 0).sum()) == 0
print("ungated wave: pairs", sb.n, "cells", st2.cells, "launches", st2.kernel_launches, "checksum", int(res2["m_new"].sum()), int(res2["C"].astype(np.int64).sum()))
ctx.close()

# block scoring (yb_score_kernel): unit seams, ranges starting inside the text, > 255 rows, many small blocks
from tools.score_cases import score_cases  # noqa: E402
ctx = YamaB200(devices=[0])
cases = score_cases()
blocks, keep2 = ctx.make_blocks(cases)
scores, st2 = ctx.score_blocks(blocks)
print("blocks", len(cases), "launches", st2.kernel_launches, "checksum", float(scores.sum()))
ctx.close()
