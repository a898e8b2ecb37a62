"""GPU parity: libyama_b200.so (through the C ABI) against the CPU oracle, bit for bit.

Reference behaviour under test: mz_yama.c:50-320 -- final C/D/I (int32), every traceback decision on
the optimal path (edit script, incl. tie-breaking :138-154 and :262-267) and the assembled columns
(:293-313).  Integer work: the bar is exact equality, no tolerance.
"""
import numpy as np
import pytest

from tools.synth import SynthBatch, random_problem

pytestmark = pytest.mark.gpu


def _check(ctx, oracle, problems, tag=""):
    jobs, keep = ctx.make_jobs(problems)
    res, st = ctx.run_batch(jobs)
    for i, (A, B, LB, RB) in enumerate(problems):
        o = oracle.yama(A, B, LB, RB, want_tback=False)
        r = res[i]
        assert r["status"] == 0, (tag, i)
        assert (int(r["C"]), int(r["D"]), int(r["I"])) == tuple(int(x) for x in o["cdi"]), (tag, i, A.shape, B.shape)
        assert int(r["m_new"]) == o["m_new"], (tag, i)
        assert np.array_equal(ctx.script_of(r), o["script"]), (tag, i)
        assert np.array_equal(ctx.assemble(jobs[i], r), o["al"]), (tag, i)
        assert int(r["cells"]) == o["cells"]
    return st


@pytest.mark.parametrize("band", ["smooth", "full", "ragged"])
@pytest.mark.parametrize("alphabet", ["acgt", "mixed", "weird"])
def test_random_small(yama_ctx, oracle, band, alphabet):
    rng = np.random.default_rng(hash((band, alphabet)) & 0xffff)
    probs = []
    for it in range(120):
        K, L = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        M, N = int(rng.integers(1, 90)), int(rng.integers(1, 90))
        if band == "full" and M * N > 3000:
            M = max(1, 3000 // N)
        probs.append(random_problem(rng, K, L, M, N, band=band, alphabet=alphabet))
    _check(yama_ctx, oracle, probs, f"{band}/{alphabet}")


def test_tiny_shapes(yama_ctx, oracle):
    rng = np.random.default_rng(5)
    probs = []
    for M in (1, 2, 3, 31, 32, 33, 64, 65):
        for N in (1, 2, 9, 10, 11, 40):
            probs.append(random_problem(rng, 2, 1, M, N, band="full"))
            probs.append(random_problem(rng, 1, 3, M, N, band="ragged"))
    _check(yama_ctx, oracle, probs, "tiny")


def test_synthetic_pre_yama_like(yama_ctx, oracle):
    Ks = [2, 3, 4, 5, 8, 1, 2, 6] * 6
    Ls = [1, 1, 1, 1, 2, 1, 3, 6] * 6
    Ms = [400, 150, 700, 60, 90, 333, 20, 250] * 6
    for R in (30, 100):
        sb = SynthBatch(11 + R, Ks, Ls, Ms, R=R, lower=0.03)
        probs = [sb.problem(i) for i in range(sb.n)]
        _check(yama_ctx, oracle, probs, f"synth R={R}")


def test_deep_profiles(yama_ctx, oracle):
    sb = SynthBatch(3, [63, 32, 99, 90], [1, 32, 1, 10], [120, 100, 80, 80], R=30)
    _check(yama_ctx, oracle, [sb.problem(i) for i in range(sb.n)], "deep")


def test_single_call_mirror(yama_ctx, oracle):
    rng = np.random.default_rng(9)
    A, B, LB, RB = random_problem(rng, 3, 2, 70, 66, band="smooth")
    al, m = yama_ctx.yama(A, 3, 70, B, 2, 66, LB, RB)
    o = oracle.yama(A, B, LB, RB)
    assert m == o["m_new"] and np.array_equal(al, o["al"])


@pytest.mark.parametrize("name", ["yama_small.npz", "yama_deep.npz"])
def test_golden_fixtures(yama_ctx, name):
    """The CUDA path against the committed outputs of the compiled reference (tools/make_golden.py) --
    no oracle in the loop.  yama_deep.npz includes K+L=100 profiles and the int32 wrap case."""
    from golden_util import GoldenYama
    g = GoldenYama(name)
    jobs, keep = yama_ctx.make_jobs(g.problems())
    res, _ = yama_ctx.run_batch(jobs)
    for i in range(g.n):
        r = res[i]
        assert r["status"] == 0
        g.check(i, dict(cdi=(r["C"], r["D"], r["I"]), m_new=r["m_new"], script=yama_ctx.script_of(r),
                        al=yama_ctx.assemble(jobs[i], r), cells=r["cells"]))


def test_wide_bands_cta_per_pair(yama_ctx, oracle):
    """Wide bands run one CTA per pair (4 or 8 warps form one wavefront, bins 2-4 of yama_b200.cu); short wide pairs
    stay one warp per pair on the 512-entry ring (bin 1).  Rows straddle the 128- and 256-lane block sizes."""
    cases = [  # (K, L, M, R)
        (2, 1, 600, 150), (3, 2, 191, 150), (2, 1, 192, 150), (1, 1, 257, 120), (4, 1, 129, 200),    # bins 1 / 2
        (2, 1, 1200, 400), (3, 1, 255, 400), (2, 2, 513, 380),                                       # bin 3
        (2, 1, 2300, 1100), (1, 1, 1030, 1500),                                                      # bin 4
    ]
    for seed, (K, L, M, R) in enumerate(cases):
        sb = SynthBatch(40 + seed, [K, K], [L, L], [M, max(1, M - 37)], R=R, indel=0.02)
        _check(yama_ctx, oracle, [sb.problem(i) for i in range(sb.n)], f"wide K={K} L={L} M={M} R={R}")
