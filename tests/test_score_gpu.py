"""GPU parity of the block-scoring kernel (yb_score_blocks, through the C ABI) against the CPU oracle and the
committed reference scores.

Reference behaviour under test: mafScoreRange, mz_scores.c:124-152 -- the sum over columns and row pairs of
SS - GAP2, an integer-valued double.  The bar is exact equality of the doubles.
"""
import numpy as np
import pytest

from golden_util import GoldenScores
from tools.score_cases import alignment_block

pytestmark = pytest.mark.gpu


def _scores(ctx, cases):
    blocks, keep = ctx.make_blocks(cases)
    sc, st = ctx.score_blocks(blocks)
    del keep
    return sc, st


def test_golden_scores_hox70(yama_ctx):
    g = GoldenScores()
    sc, st = _scores(yama_ctx, g.blocks())
    assert np.array_equal(sc, g.expected(70)), np.nonzero(sc != g.expected(70))
    assert st.kernel_launches >= 1 and st.pairs == g.n


def test_golden_scores_hox85():
    from multiz_b200 import YamaB200
    g = GoldenScores()
    ctx = YamaB200(devices=[0], scores=85)
    try:
        sc, _ = _scores(ctx, g.blocks())
    finally:
        ctx.close()
    assert np.array_equal(sc, g.expected(85))


def test_two_contexts_with_different_tables_do_not_interfere(yama_ctx, oracle):
    """Score tables belong to a context (kernel arguments), not to the process: a HOXD85 context created and used in
    between must not change what the HOXD70 context computes -- for scoring and for yama itself."""
    from multiz_b200 import YamaB200
    from oracle.oracle_py import Oracle
    from tools.synth import random_problem
    rng = np.random.default_rng(31)
    blocks = [(alignment_block(rng, 6, 150), 0, 150) for _ in range(20)]
    probs = [random_problem(rng, 3, 2, 60, 70, band="smooth") for _ in range(20)]
    o85 = Oracle(85)
    other = YamaB200(devices=[0], scores=85)
    try:
        for ctx, orc in ((other, o85), (yama_ctx, oracle), (other, o85), (yama_ctx, oracle)):
            sc, _ = _scores(ctx, blocks)
            assert [float(x) for x in sc] == [orc.score_range(t, s, n) for t, s, n in blocks]
            jobs, keep = ctx.make_jobs(probs)
            res, _ = ctx.run_batch(jobs)
            for i, (A, B, LB, RB) in enumerate(probs):
                want = orc.yama(A, B, LB, RB, want_tback=False)
                assert (int(res[i]["C"]), int(res[i]["D"]), int(res[i]["I"])) == tuple(int(x) for x in want["cdi"])
                assert np.array_equal(ctx.script_of(res[i]), want["script"])
    finally:
        other.close()


def test_random_blocks_against_oracle(yama_ctx, oracle):
    rng = np.random.default_rng(123)
    cases = []
    for it in range(300):
        rows, cols = int(rng.integers(1, 48)), int(rng.integers(1, 700))
        text = alignment_block(rng, rows, cols, sub=float(rng.random()) * 0.3, gap_open=float(rng.random()) * 0.12,
                               lower=0.1, other=0.03)
        start = int(rng.integers(0, cols))
        cases.append((text, start, int(rng.integers(1, cols - start + 1))))
    sc, _ = _scores(yama_ctx, cases)
    for i, (t, s, n) in enumerate(cases):
        assert sc[i] == oracle.score_range(t, s, n), (i, t.shape, s, n)


def test_every_start_and_size_of_one_block(yama_ctx, oracle):
    """All ranges of a 3-unit block: exercises the lead byte, the unit seams (columns 127/128, 255/256) and the tail."""
    rng = np.random.default_rng(4)
    text = alignment_block(rng, 5, 300, gap_open=0.1)
    cases = [(text, s, n) for s in (0, 1, 2, 3, 4, 5, 126, 127, 128, 129, 255, 256, 299) for n in (1, 2, 3, 4, 5, 127, 128, 129, 300)
             if s + n <= 300]
    sc, _ = _scores(yama_ctx, cases)
    for i, (t, s, n) in enumerate(cases):
        assert sc[i] == oracle.score_range(t, s, n), (s, n)


def test_deep_blocks_flush_byte_counters_and_use_64_bit_forms(yama_ctx, oracle):
    rng = np.random.default_rng(8)
    cases = []
    for rows, cols in ((255, 33), (256, 33), (511, 20), (700, 12), (3000, 5), (4100, 3)):
        cases.append((alignment_block(rng, rows, cols, gap_open=0.05), 0, cols))
    sc, _ = _scores(yama_ctx, cases)
    for i, (t, s, n) in enumerate(cases):
        assert sc[i] == oracle.score_range(t, s, n), (i, t.shape)


def test_empty_and_degenerate_blocks(yama_ctx):
    sc, _ = _scores(yama_ctx, [])
    assert len(sc) == 0
    one = np.frombuffer(b"ACGT-ACGT", dtype=np.uint8).reshape(1, 9)
    sc, _ = _scores(yama_ctx, [(one, 0, 9), (one, 3, 2)])
    assert list(sc) == [0.0, 0.0]                               # a single row has no pairs
    assert yama_ctx.mafScoreRange(np.frombuffer(b"AA", dtype=np.uint8).reshape(2, 1), 0, 1) == 91.0   # HOXD70 A/A


def test_bad_range_fails_like_the_reference(yama_ctx):
    from multiz_b200.yama import YamaError
    text = alignment_block(np.random.default_rng(1), 3, 20)
    for start, size in ((-1, 5), (0, 0), (10, 11), (20, 1)):
        with pytest.raises(YamaError, match="mafScoreRange: start = %d, size = %d, textSize = 20" % (start, size)) as e:
            yama_ctx.mafScoreRange(text, start, size)
        assert e.value.code == -6


def test_many_waves_and_additivity_at_size(yama_ctx, oracle):
    """A batch far beyond one 64 MB wave, checked by sampling against the oracle and -- for every block -- by the
    size-independent property score(a, n1+n2) == score(a, n1) + score(a+n1, n2)."""
    rng = np.random.default_rng(2026)
    texts = [alignment_block(rng, int(rng.integers(2, 12)), int(rng.integers(2000, 6000)), gap_open=0.02) for _ in range(24)]
    cases, split = [], []
    for k in range(12000):
        t = texts[k % len(texts)]
        cols = t.shape[1]
        a = int(rng.integers(0, cols - 2)); n = int(rng.integers(2, cols - a + 1)); n1 = int(rng.integers(1, n))
        cases += [(t, a, n), (t, a, n1), (t, a + n1, n - n1)]
        split.append((a, n, n1))
    sc, st = _scores(yama_ctx, cases)
    assert st.h2d_bytes > 2 * (64 << 20)                        # three waves or more
    whole, left, right = sc[0::3], sc[1::3], sc[2::3]
    assert np.array_equal(whole, left + right)
    for i in rng.choice(len(cases), 60, replace=False):
        t, s, n = cases[i]
        assert sc[i] == oracle.score_range(t, s, n), i
