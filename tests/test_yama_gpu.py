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


def test_batch_api_edges(yama_ctx, oracle):
    """Empty batch, per-pair status for an invalid band / an over-deep profile inside a good batch (the good pairs
    still align), and the queue form yb_submit / yb_flush / yb_fetch giving the same answers as yb_run_batch."""
    import ctypes as C
    from multiz_b200 import yama as ym
    res, st = yama_ctx.run_batch(np.zeros(0, dtype=ym.JOB_DTYPE))
    assert len(res) == 0 and st.pairs == 0

    rng = np.random.default_rng(77)
    good = [random_problem(rng, 2, 1, 40, 44, band="smooth"), random_problem(rng, 3, 2, 25, 30, band="ragged")]
    A, B, LB, RB = random_problem(rng, 2, 1, 30, 35, band="smooth")
    bad_band = (A, B, LB, RB.copy())
    bad_band[3][10] = max(0, int(bad_band[2][10]) + 2)                # row narrower than min(N,10): mz_yama.c:63-65
    deep = random_problem(rng, 300, 1, 12, 12, band="full")           # K > 255
    jobs, keep = yama_ctx.make_jobs([good[0], bad_band, good[1], deep])
    res, st = yama_ctx.run_batch(jobs, check=False)
    assert [int(r["status"]) for r in res] == [0, -2, 0, -4]
    msg = yama_ctx.lib.yb_last_error(yama_ctx.h).decode()
    assert msg.startswith("RB[10] - LB[10] < 10"), msg                # the first failure, in the reference's wording
    for i, k in ((0, 0), (2, 1)):
        o = oracle.yama(*good[k], want_tback=False)
        assert np.array_equal(yama_ctx.script_of(res[i]), o["script"])

    lib = yama_ctx.lib
    lib.yb_clear(yama_ctx.h)
    ids = []
    for (A, B, LB, RB) in good:
        j, k2 = yama_ctx.make_jobs([(A, B, LB, RB)])
        job = ym.yb_job(*[int(j[0][f]) for f in ("K", "M", "L", "N", "A", "B", "LB", "RB")])
        ids.append(lib.yb_submit(yama_ctx.h, C.byref(job)))
        del k2                                                        # submit copies: the inputs may go away
    assert ids == [0, 1]
    stq = ym.yb_stats()
    assert lib.yb_flush(yama_ctx.h, C.byref(stq)) == 0 and stq.pairs == 2
    for i, (A, B, LB, RB) in enumerate(good):
        r = ym.yb_result()
        assert lib.yb_fetch(yama_ctx.h, i, C.byref(r)) == 0
        o = oracle.yama(A, B, LB, RB, want_tback=False)
        assert (r.C, r.D, r.I) == tuple(int(x) for x in o["cdi"]) and r.m_new == o["m_new"]
    lib.yb_clear(yama_ctx.h)


def test_resident_path_matches_batch(yama_ctx, oracle):
    """yb_resident_load/step/fetch (what bench.py times as `value`) returns exactly what yb_run_batch returns."""
    sb = SynthBatch(321, [2, 3, 4, 5, 2, 8] * 20, [1, 1, 1, 1, 2, 8] * 20, list(np.random.default_rng(5).integers(5, 700, 120)), R=30)
    res_b, _ = yama_ctx.run_batch(sb.jobs)
    scripts_b = [yama_ctx.script_of(r) for r in res_b]
    yama_ctx.resident_load(sb.jobs)
    for _ in range(2):
        st = yama_ctx.resident_step()
    assert st.cells == sb.cells
    res_r = yama_ctx.resident_fetch()
    for i in range(sb.n):
        assert tuple(res_r[i][f] for f in ("status", "m_new", "C", "D", "I", "cells")) == \
               tuple(res_b[i][f] for f in ("status", "m_new", "C", "D", "I", "cells")), i
        assert np.array_equal(yama_ctx.script_of(res_r[i]), scripts_b[i]), i
    A, B, LB, RB = sb.problem(7)
    o = oracle.yama(A, B, LB, RB, want_tback=False)
    assert np.array_equal(scripts_b[7], o["script"])


def test_full_size_batch_sampled_and_wave_invariant(oracle):
    """BASELINE.json configs[1] at full size (119 490 pairs, 2.29 G cells): (i) a size-stratified sample of pairs is
    compared with the oracle bit for bit, (ii) every script consumes exactly its two alignments, (iii) the results do
    not depend on how the batch is cut into waves (8 MB waves vs the default) nor on a second run."""
    import os
    from bench import make_batch
    from multiz_b200 import YamaB200
    sb, _ = make_batch("cfg2", 1234, 1.0)
    ctx = YamaB200(devices=[0])
    res, st = ctx.run_batch(sb.jobs)
    assert st.cells == sb.cells and int((res["status"] != 0).sum()) == 0
    ops = np.zeros(0, np.uint8)
    order = np.argsort(sb.M, kind="stable")
    sample = np.unique(np.concatenate([order[:: max(1, sb.n // 150)], order[-20:], order[:20]]))
    for i in sample:
        A, B, LB, RB = sb.problem(int(i))
        o = oracle.yama(A, B, LB, RB, want_tback=False)
        r = res[int(i)]
        assert (int(r["C"]), int(r["D"]), int(r["I"])) == tuple(int(x) for x in o["cdi"]), int(i)
        assert np.array_equal(ctx.script_of(r), o["script"]), int(i)
    # (ii) on every pair: #C + #D == M, #C + #I == N  (what yb_assemble / mz_yama.c:310-312 insist on)
    packed_len = (res["m_new"].astype(np.int64) + 3) // 4
    import ctypes as C
    nC = np.zeros(sb.n, np.int64); nI = np.zeros(sb.n, np.int64); nD = np.zeros(sb.n, np.int64)
    for i in range(0, sb.n, 97):                      # every 97th pair: ~1200 pairs
        s = ctx.script_of(res[i])
        nC[i], nI[i], nD[i] = (s == 0).sum(), (s == 1).sum(), (s == 2).sum()
        assert nC[i] + nD[i] == sb.M[i] and nC[i] + nI[i] == sb.N[i], i
    del ops, packed_len, C
    # (iii) wave-size and run-to-run invariance
    keep = [(int(r["m_new"]), int(r["C"]), int(r["D"]), int(r["I"])) for r in res]
    scripts = {int(i): ctx.script_of(res[int(i)]).tobytes() for i in sample}
    ctx.close()
    os.environ["YB_WAVE_MB"] = "8"
    try:
        ctx2 = YamaB200(devices=[0])
    finally:
        del os.environ["YB_WAVE_MB"]
    for _ in range(2):
        res2, _st = ctx2.run_batch(sb.jobs)
        assert [(int(r["m_new"]), int(r["C"]), int(r["D"]), int(r["I"])) for r in res2] == keep
        for i in sample:
            assert ctx2.script_of(res2[int(i)]).tobytes() == scripts[int(i)]
    ctx2.close()


def test_two_devices_in_one_context_match_one_device(oracle):
    """inside one process the waves of a batch go to whichever device is free; results must not depend on
    it.  Needs two visible GPUs (skipped on a one-GPU box): batch, resident path and block scores with devices [0, 1]
    against devices [0]."""
    from multiz_b200 import YamaB200
    try:
        import torch
        ndev = torch.cuda.device_count()
    except Exception:
        ndev = 1
    if ndev < 2:
        pytest.skip("one visible GPU")
    Ms = list(np.random.default_rng(9).integers(5, 900, 600))
    sb = SynthBatch(77, [2, 3, 4, 5, 2, 8] * 100, [1, 1, 1, 1, 2, 8] * 100, Ms, R=30)
    one = YamaB200(devices=[0])
    two = YamaB200(devices=[0, 1])
    try:
        assert two.n_devices == 2
        import os
        os.environ["YB_WAVE_MB"] = "1"                      # many waves, so both devices get some
        try:
            two_small = YamaB200(devices=[0, 1])
        finally:
            del os.environ["YB_WAVE_MB"]
        ra, _ = one.run_batch(sb.jobs)
        for ctx in (two, two_small):
            rb, st = ctx.run_batch(sb.jobs)
            assert st.n_devices == 2 and st.cells == sb.cells
            for i in range(sb.n):
                assert tuple(ra[i][f] for f in ("status", "m_new", "C", "D", "I", "cells")) == \
                       tuple(rb[i][f] for f in ("status", "m_new", "C", "D", "I", "cells")), i
                assert np.array_equal(one.script_of(ra[i]), ctx.script_of(rb[i])), i
        two.resident_load(sb.jobs)
        two.resident_step()
        rr = two.resident_fetch()
        for i in range(0, sb.n, 7):
            assert np.array_equal(one.script_of(ra[i]), two.script_of(rr[i])), i
        A, B, LB, RB = sb.problem(11)
        assert np.array_equal(one.script_of(ra[11]), oracle.yama(A, B, LB, RB, want_tback=False)["script"])
        two_small.close()
    finally:
        one.close()
        two.close()


def test_fill_variant_without_existence_multipliers_is_exact(oracle):
    """Waves of small pairs run a fill kernel that charges candidates from non-existent nodes like any other (DESIGN
    section 2: they are exactly MININT and can never win).  Same batch with the variant forced off (YB_UNGATED=0): every
    result field and script must agree, and both must agree with the oracle; disconnected bands and pairs beyond the
    size bound must fall back to the gated kernel by themselves."""
    import os
    from multiz_b200 import YamaB200
    rng = np.random.default_rng(404)
    probs = []
    for it in range(400):
        K, L = int(rng.integers(1, 9)), int(rng.integers(1, 5))
        M, N = int(rng.integers(1, 200)), int(rng.integers(1, 200))
        band = ("smooth", "ragged", "full")[it % 3]
        if band == "full" and M * N > 4000:
            M = max(1, 4000 // N)
        probs.append(random_problem(rng, K, L, M, N, band=band, alphabet=("acgt", "mixed", "weird")[it % 3]))
    sb = SynthBatch(9, [2, 3, 5, 8] * 30, [1, 1, 2, 4] * 30, list(rng.integers(30, 1500, 120)), R=30, lower=0.05)
    probs += [tuple(np.array(x) for x in sb.problem(i)) for i in range(sb.n)]
    on = YamaB200(devices=[0])
    os.environ["YB_UNGATED"] = "0"
    try:
        off = YamaB200(devices=[0])
    finally:
        del os.environ["YB_UNGATED"]
    try:
        jobs, keep = on.make_jobs(probs)
        ra, _ = on.run_batch(jobs)
        rb, _ = off.run_batch(jobs)
        for i in range(len(probs)):
            assert tuple(ra[i][f] for f in ("status", "m_new", "C", "D", "I", "cells")) == \
                   tuple(rb[i][f] for f in ("status", "m_new", "C", "D", "I", "cells")), i
            assert np.array_equal(on.script_of(ra[i]), off.script_of(rb[i])), i
        for i in range(0, len(probs), 5):
            A, B, LB, RB = probs[i]
            o = oracle.yama(A, B, LB, RB, want_tback=False)
            assert np.array_equal(on.script_of(ra[i]), o["script"]), i
            assert (int(ra[i]["C"]), int(ra[i]["D"]), int(ra[i]["I"])) == tuple(int(x) for x in o["cdi"]), i
    finally:
        on.close()
        off.close()


def _oracle_many(oracle, problems, threads=8):
    """The oracle on several large pairs at once (ctypes releases the GIL: one host core per pair)."""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=threads) as ex:
        return list(ex.map(lambda p: oracle.yama(*p, want_tback=False), problems))


def test_cfg5_shape_full_size(yama_ctx, oracle):
    """BASELINE.json configs[4] at its stated shape: K=99,L=1 and K=90,L=10 blocks of M ~ 10^4 columns, R=300
    (~6 M cells per pair, band rows of ~601 cells -> the CTA-per-pair kernels with existence multipliers and the
    warp-per-path traceback).  Column scores reach 9*10^4, final scores 10^8..10^9 -- the upper end of the int32
    range the reference computes in (mz_yama.c:29 MININT = INT_MIN/2).  Two pairs of each kind against the oracle:
    final C/D/I, the whole edit script, the assembled columns."""
    sb = SynthBatch(505, [99, 90, 99, 90], [1, 10, 1, 10], [10000, 10000, 9473, 10240], R=300)
    probs = [tuple(np.array(x) for x in sb.problem(i)) for i in range(sb.n)]
    # the largest scores the shape allows: 90 x 10 identical residues per column, 91 each -> C(M,N) = 8.19e8 of 2^30
    M = 10000
    LB, RB = oracle.smooth(np.concatenate([[0], np.arange(1, M + 1)]), np.concatenate([[M], np.arange(1, M + 1)]), M, M, 300)
    probs.append((np.full((M, 90), ord("A"), np.uint8), np.full((M, 10), ord("A"), np.uint8), LB, RB))
    jobs, keep = yama_ctx.make_jobs(probs)
    res, st = yama_ctx.run_batch(jobs)
    outs = _oracle_many(oracle, probs)
    assert int(outs[-1]["cdi"][0]) == 91 * 900 * M
    for i, o in enumerate(outs):
        r = res[i]
        assert r["status"] == 0, i
        assert int(r["cells"]) == o["cells"] and o["cells"] > 5_000_000, i
        assert (int(r["C"]), int(r["D"]), int(r["I"])) == tuple(int(x) for x in o["cdi"]), i
        assert int(r["m_new"]) == o["m_new"], i
        assert np.array_equal(yama_ctx.script_of(r), o["script"]), i
        assert np.array_equal(yama_ctx.assemble(jobs[i], r), o["al"]), i


def test_cfg3_depth64_r100(yama_ctx, oracle):
    """BASELINE.json configs[2] at its deepest point: profile depth 64 split K ~ L (32 x 32) and K = 63, L = 1,
    M = 500, band R = 100 (201-cell rows: the 512-entry ring bins, CTA-per-pair for M >= 192), plus R = 30."""
    for R in (100, 30):
        sb = SynthBatch(640 + R, [32, 63, 32, 16, 31], [32, 1, 32, 16, 1], [500, 500, 191, 500, 500], R=R)
        probs = [sb.problem(i) for i in range(sb.n)]
        res, _ = yama_ctx.run_batch(sb.jobs)
        for i, o in enumerate(_oracle_many(oracle, probs)):
            r = res[i]
            assert r["status"] == 0, (R, i)
            assert (int(r["C"]), int(r["D"]), int(r["I"])) == tuple(int(x) for x in o["cdi"]), (R, i)
            assert np.array_equal(yama_ctx.script_of(r), o["script"]), (R, i)
            assert np.array_equal(yama_ctx.assemble(sb.jobs[i], r), o["al"]), (R, i)


def test_delta_coded_bands_match_raw_bands(oracle):
    """YB_BAND_PACK=1: the host ships a band as one byte per row (the step LB[r]-LB[r-1]) and yb_band_expand restores the
    caller's int arrays on the device.  Steps of 255 and more travel as exceptions; an invalid band (a negative step, a row
    that is too narrow) must still be reported per pair in the reference's words; results equal the raw-band run and the
    oracle.  Rows around the kernel's 128-row chunks and 4-row words."""
    import os
    from multiz_b200 import YamaB200
    rng = np.random.default_rng(2025)
    probs = []
    for M in (1, 2, 3, 4, 5, 127, 128, 129, 130, 255, 256, 257, 700):
        probs.append(random_problem(rng, 2, 1, M, int(rng.integers(1, 90)), band=("smooth", "ragged", "full")[M % 3]))
    # long jumps: a band that follows a diagonal, leaps 255 / 256 / 1000 columns, and goes on (connected: RB leaps first)
    for jump in (253, 254, 255, 1000):                   # steps of 254 (a byte), 255, 256, 1001 (exceptions)
        M, N = 300, 300 + jump
        diag = np.arange(M + 1)
        diag = np.where(diag > 150, diag + jump, diag)
        LB = np.maximum(0, np.where(np.arange(M + 1) > 160, diag - 20, np.minimum(diag, np.arange(M + 1)) - 20)).astype(np.int32)
        RB = np.minimum(N, np.where(np.arange(M + 1) > 140, np.maximum(diag, np.arange(M + 1) + jump) + 20, diag + 20)).astype(np.int32)
        LB = np.maximum.accumulate(LB); RB = np.maximum.accumulate(RB)
        LB[0] = 0; RB[M] = N
        A, B, _, _ = random_problem(rng, 3, 1, M, N, band="full")
        probs.append((A, B, LB, RB))
    good = len(probs)
    A, B, LB, RB = random_problem(rng, 2, 1, 60, 70, band="smooth")
    bad1 = (A, B, LB.copy(), RB.copy()); bad1[2][40] = bad1[2][39] + 2; bad1[2][41] = bad1[2][39] + 1   # LB decreases: mz_yama.c:67-68
    bad2 = (A, B, LB.copy(), RB.copy()); bad2[3][10] = max(0, int(bad2[2][10]) + 2)          # too narrow: mz_yama.c:63-65
    probs += [bad1, bad2]
    out = {}
    for mode in ("0", "1"):
        os.environ["YB_BAND_PACK"] = mode
        try:
            ctx = YamaB200(devices=[0])
            jobs, keep = ctx.make_jobs(probs)
            res, st = ctx.run_batch(jobs, check=False)
            out[mode] = ([tuple(int(r[f]) for f in ("status", "m_new", "C", "D", "I", "cells")) for r in res],
                         [ctx.script_of(r).tobytes() if r["status"] == 0 else b"" for r in res],
                         ctx.lib.yb_last_error(ctx.h).decode(), st.h2d_bytes)
            ctx.close()
        finally:
            del os.environ["YB_BAND_PACK"]
    assert out["0"][:3] == out["1"][:3]
    assert out["1"][3] < out["0"][3]                     # fewer bytes went over the bus
    assert [t[0] for t in out["1"][0][good:]] == [-2, -2]
    for i in range(good):
        o = oracle.yama(*probs[i], want_tback=False)
        t = out["1"][0][i]
        assert t[0] == 0 and t[1] == o["m_new"] and t[2:5] == tuple(int(x) for x in o["cdi"]), i
        assert out["1"][1][i] == o["script"].tobytes(), i
