"""CPU side of the block-scoring path (mafScoreRange, mz_scores.c:124-152): the oracle restatement against the
committed golden scores and the live reference, the generator's determinism, and the ABI shim's error behaviour."""
import ctypes as C
import os

import numpy as np
import pytest

from golden_util import GoldenScores
from tools.score_cases import alignment_block, score_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_matches_golden_scores_both_tables():
    from oracle.oracle_py import Oracle
    g = GoldenScores()
    assert g.n >= 90
    for which in (70, 85):
        o = Oracle(which)
        want = g.expected(which)
        for i in range(g.n):
            text, start, size = g.block(i)
            assert o.score_range(text, start, size) == want[i], (which, i, text.shape, start, size)


def test_generator_is_the_one_that_made_the_fixture():
    g = GoldenScores()
    cases = score_cases()
    assert len(cases) == g.n
    for i, (text, start, size) in enumerate(cases):
        t2, s2, n2 = g.block(i)
        assert np.array_equal(text, t2) and (start, size) == (s2, n2), i


def test_oracle_matches_live_reference(reference, oracle):
    rng = np.random.default_rng(77)
    for it in range(150):
        rows, cols = int(rng.integers(1, 40)), int(rng.integers(1, 300))
        text = alignment_block(rng, rows, cols, sub=0.2, gap_open=0.08)
        start = int(rng.integers(0, cols))
        size = int(rng.integers(1, cols - start + 1))
        assert oracle.score_range(text, start, size) == reference.score_range(text, start, size), (it, rows, cols)


def test_oracle_range_errors(oracle):
    text = alignment_block(np.random.default_rng(1), 3, 20)
    for start, size in ((-1, 5), (0, 0), (5, -2), (10, 11), (20, 1)):
        with pytest.raises(ValueError, match="mafScoreRange: start = %d, size = %d, textSize = 20" % (start, size)):
            oracle.score_range(text, start, size)


def test_score_is_additive_over_column_ranges(oracle):
    """mafScoreRange(a, n1+n2) == mafScoreRange(a, n1) + mafScoreRange(a+n1, n2): the gap term of a column looks one
    column back whether or not that column is inside the range (mz_scores.c:143-147)."""
    rng = np.random.default_rng(9)
    text = alignment_block(rng, 7, 500)
    for _ in range(20):
        a = int(rng.integers(0, 400)); n1 = int(rng.integers(1, 50)); n2 = int(rng.integers(1, 50))
        assert oracle.score_range(text, a, n1 + n2) == oracle.score_range(text, a, n1) + oracle.score_range(text, a + n1, n2)


def test_shim_answers_the_score_abi_like_the_oracle(oracle):
    """oracle/libyama_shim.so implements yb_score_blocks for the drop-in's CPU tests: same struct, same errors."""
    from multiz_b200.yama import BLOCK_DTYPE, YamaB200, yb_stats
    path = os.path.join(ROOT, "oracle", "libyama_shim.so")
    if not os.path.exists(path):
        from oracle.oracle_py import build
        build()
    shim = C.CDLL(path)
    shim.yb_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    shim.yb_set_scores.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    shim.yb_score_blocks.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    shim.yb_last_error.argtypes = [C.c_void_p]
    shim.yb_last_error.restype = C.c_char_p
    shim.yb_destroy.argtypes = [C.c_void_p]
    h = C.c_void_p()
    assert shim.yb_create(None, 0, C.byref(h)) == 0
    cases = score_cases()[:30]
    blocks, keep = YamaB200.make_blocks(cases)
    assert blocks.dtype == BLOCK_DTYPE
    scores = np.zeros(len(cases))
    assert shim.yb_score_blocks(h, len(cases), blocks.ctypes.data, scores.ctypes.data, None) == -3      # YB_ERR_SCORES
    assert b"scores not initialized" in shim.yb_last_error(h)
    shim.yb_set_scores(h, oracle.ss.ctypes.data, oracle.gop.ctypes.data, oracle.gap_ext)
    st = yb_stats()
    assert shim.yb_score_blocks(h, len(cases), blocks.ctypes.data, scores.ctypes.data, C.byref(st)) == 0
    for i, (t, s, n) in enumerate(cases):
        assert scores[i] == oracle.score_range(t, s, n)
    blocks[3]["size"] = blocks[3]["text_size"] + 1
    assert shim.yb_score_blocks(h, len(cases), blocks.ctypes.data, scores.ctypes.data, None) == -6      # YB_ERR_ARG
    assert b"mafScoreRange: start = " in shim.yb_last_error(h)
    shim.yb_destroy(h)


def _count_form_score(text, start, size, S6, gap_open):
    """The kernel's derivation (score_kernels.cuh) in numpy: per column ten counts and a quadratic form."""
    t = np.asarray(text, dtype=np.uint8)
    low = t | 0x20
    cls = np.full(t.shape, 4, dtype=np.int64)
    for k, ch in enumerate(b"acgt"):
        cls[low == ch] = k
    cls[t == ord("-")] = 5
    total = 0
    for i in range(start, start + size):
        n = np.bincount(cls[:, i], minlength=6).astype(np.int64)
        s = 0
        for k in range(6):
            s += int(S6[k][k]) * int(n[k] * (n[k] - 1) // 2)
            for l in range(k + 1, 6):
                s += int(S6[k][l]) * int(n[k] * n[l])
        if i > 0:
            d, p = t[:, i] == ord("-"), t[:, i - 1] == ord("-")
            t00, t01, t10, t11 = int((~p & ~d).sum()), int((~p & d).sum()), int((p & ~d).sum()), int((p & d).sum())
            s -= gap_open * (t00 * t01 + t01 * t10 + t10 * t11)
        total += s
    return float(total)


@pytest.mark.parametrize("which", [70, 85])
def test_count_vector_form_equals_the_pair_loop(which):
    """The O(rows) closed form the CUDA kernel uses (six class counts + three dash-transition counts per column) against
    the oracle's O(rows^2) pair loop, on the golden blocks and on random ones, for both score sets."""
    from oracle.oracle_py import Oracle
    o = Oracle(which)
    reps = b"ACGTN-"
    S6 = [[int(o.ss[a, b]) for b in reps] for a in reps]
    GO = int(o.gop[1])
    assert all(S6[a][b] == S6[b][a] for a in range(6) for b in range(6))
    g = GoldenScores()
    want = g.expected(which)
    for i in range(0, g.n, 3):
        text, start, size = g.block(i)
        if text.shape[0] * size > 40000:
            continue
        assert _count_form_score(text, start, size, S6, GO) == want[i], i
    rng = np.random.default_rng(which)
    for _ in range(40):
        rows, cols = int(rng.integers(1, 25)), int(rng.integers(1, 120))
        text = alignment_block(rng, rows, cols, sub=0.25, gap_open=0.1, lower=0.1, other=0.05)
        start = int(rng.integers(0, cols)); size = int(rng.integers(1, cols - start + 1))
        assert _count_form_score(text, start, size, S6, GO) == o.score_range(text, start, size)
