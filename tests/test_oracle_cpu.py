"""CPU suite, part 1: the oracle (oracle/yama_oracle.c) is pinned against the reference.

 * tests/golden/*.npz were produced by tools/make_golden.py from the UNMODIFIED reference compiled into
   oracle/_ref (mz_yama.c:50-320, mz_scores.c:94-122, mz_preyama.c:17-35).  They travel; the reference does not.
 * where oracle/_ref is present (the build container) the oracle is also diffed live against it on fresh
   random problems, byte for byte: traceback matrix, final C/D/I, edit script, assembled columns.
"""
import numpy as np
import pytest

from golden_util import GOLD, GoldenYama
from tools.synth import SynthBatch, random_problem

import os


def test_scores_match_reference_tables(oracle):
    z = np.load(os.path.join(GOLD, "scores.npz"))
    for which in (70, 85):
        oracle.set_scores(which)
        assert np.array_equal(oracle.ss, z[f"ss{which}"])
        assert np.array_equal(oracle.gop, z[f"gop{which}"])
        assert oracle.gap_ext == int(z[f"ge{which}"][0])
    oracle.set_scores(70)


def test_product_score_tables_match_reference():
    """multiz_b200.hox70_tables (what the ctypes mirror feeds yb_set_scores) == init_scores70/85."""
    from multiz_b200 import hox70_tables
    z = np.load(os.path.join(GOLD, "scores.npz"))
    for which in (70, 85):
        ss, gop, ge = hox70_tables(which)
        assert np.array_equal(ss, z[f"ss{which}"]) and np.array_equal(gop, z[f"gop{which}"]) and ge == int(z[f"ge{which}"][0])


def test_smooth_golden(oracle):
    z = np.load(os.path.join(GOLD, "smooth.npz"))
    off = z["off"]
    for i in range(len(z["M"])):
        s = slice(off[i], off[i + 1])
        LB, RB = oracle.smooth(z["LB_in"][s], z["RB_in"][s], int(z["M"][i]), int(z["N"][i]), int(z["R"][i]))
        assert np.array_equal(LB, z["LB_out"][s]) and np.array_equal(RB, z["RB_out"][s]), i


@pytest.mark.parametrize("name", ["yama_small.npz", "yama_deep.npz"])
def test_oracle_golden(oracle, name):
    g = GoldenYama(name)
    assert g.n > 5
    for i in range(g.n):
        A, B, LB, RB = g.problem(i)
        g.check(i, oracle.yama(A, B, LB, RB, want_tback=True))


def test_oracle_band_validation_messages(oracle):
    """The wording of mz_yama.c:59,:64 (these reach stderr through fatalf in the reference)."""
    A = np.full((3, 1), ord("A"), np.uint8)
    B = np.full((12, 1), ord("A"), np.uint8)
    with pytest.raises(ValueError, match="LB and RB not terminated properly: 1 12 12"):
        oracle.yama(A, B, [1, 1, 1, 1], [12, 12, 12, 12])
    with pytest.raises(ValueError, match=r"RB\[1\] - LB\[1\] < 10, 5 0 12"):
        oracle.yama(A, B, [0, 0, 0, 0], [12, 5, 12, 12])
    with pytest.raises(ValueError, match="LB not monotonic"):
        oracle.yama(A, B, [0, 1, 0, 0], [12, 12, 12, 12])
    with pytest.raises(ValueError, match="RB not monotonic"):
        oracle.yama(A, B, [0, 0, 0, 0], [12, 12, 11, 12])


@pytest.mark.parametrize("band", ["smooth", "full", "ragged"])
def test_oracle_vs_live_reference(oracle, reference, band):
    rng = np.random.default_rng({"smooth": 1, "full": 2, "ragged": 3}[band])
    for it in range(150):
        K, L = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        M, N = int(rng.integers(1, 80)), int(rng.integers(1, 80))
        if band == "full" and M * N > 2500:
            M = max(1, 2500 // N)
        A, B, LB, RB = random_problem(rng, K, L, M, N, band=band, alphabet=("acgt", "mixed", "weird")[it % 3])
        o = oracle.yama(A, B, LB, RB)
        r = reference.yama(A, B, LB, RB)
        assert np.array_equal(o["cdi"], r["cdi"]), it
        assert np.array_equal(o["tback"], r["tback"]), it
        assert np.array_equal(o["script"], r["script"]) and np.array_equal(o["al"], r["al"]), it


def test_oracle_vs_live_reference_synth(oracle, reference):
    """Problems shaped like pre_yama()'s (mz_preyama.c:174-259), both radii of BASELINE.json, HOX85 too."""
    for which, R in ((70, 30), (70, 100), (85, 30)):
        oracle.set_scores(which)
        reference.lib.ref_init_scores(which)
        sb = SynthBatch(100 + R + which, [2, 3, 5, 8, 16, 1], [1, 1, 1, 8, 16, 1], [400, 250, 130, 90, 70, 300], R=R, lower=0.03)
        for i in range(sb.n):
            A, B, LB, RB = sb.problem(i)
            o = oracle.yama(A, B, LB, RB)
            r = reference.yama(A, B, LB, RB)
            assert np.array_equal(o["cdi"], r["cdi"]) and np.array_equal(o["tback"], r["tback"]), (which, R, i)
            assert np.array_equal(o["al"], r["al"]), (which, R, i)
    oracle.set_scores(70)
    reference.lib.ref_init_scores(70)
