#!/usr/bin/env python
"""Generate tests/golden/ from the UNMODIFIED reference compiled into oracle/_ref (run in the build
container, where /root/reference exists; the fixtures travel, the reference does not).

  yama_small.npz   ~90 yama() problems (three band kinds x three alphabets + shapes around the
                   32-row / 4-step granularities of the CUDA wavefront) with the reference's own
                   final C/D/I, full traceback matrix, edit script and assembled columns
                   (mz_yama.c:50-320 through oracle/ref_hook.c).
  yama_deep.npz    deep profiles (K+L up to 100) and an int32 wrap case, outputs hashed.
  scores.npz       ss[128][128], gop[16], gap_extend as init_scores70/85 leave them (mz_scores.c:94-122).
  smooth.npz       LB/RB before/after the reference's smooth() (mz_preyama.c:17-35).

  maf/             a 12 kb three-species synthetic data set (tools/mafsynth.py) and the byte-exact outputs of
                   the reference's own `multiz` on it: v=1, v=0, and the second step of a progressive merge
                   (stdout + out1 + out2 each).

  python tools/make_golden.py            # rewrites tests/golden/ deterministically
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.oracle_py import Reference, build  # noqa: E402
from tools.synth import SynthBatch, random_problem  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def digest(a: np.ndarray) -> np.ndarray:
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8).copy()


def pack_problems(ref, problems, full: bool):
    """Flatten inputs and the reference's outputs into a dict of arrays (npz-friendly)."""
    d = {k: [] for k in ("K", "M", "L", "N", "A", "B", "LB", "RB", "cdi", "m_new", "script", "al", "tback", "cells")}
    for (A, B, LB, RB) in problems:
        M, K = A.shape
        N, L = B.shape
        r = ref.yama(A, B, LB, RB, want_tback=True)
        d["K"].append(K); d["M"].append(M); d["L"].append(L); d["N"].append(N)
        d["A"].append(A.ravel()); d["B"].append(B.ravel())
        d["LB"].append(np.asarray(LB, np.int32)); d["RB"].append(np.asarray(RB, np.int32))
        d["cdi"].append(r["cdi"]); d["m_new"].append(r["m_new"]); d["cells"].append(r["cells"])
        d["script"].append(r["script"])
        if full:
            d["al"].append(r["al"].ravel()); d["tback"].append(r["tback"])
        else:
            d["al"].append(digest(r["al"])); d["tback"].append(digest(r["tback"]))
    out = {}
    for k in ("K", "M", "L", "N", "m_new"):
        out[k] = np.asarray(d[k], np.int32)
    out["cells"] = np.asarray(d["cells"], np.int64)
    out["cdi"] = np.asarray(d["cdi"], np.int32)
    for k in ("A", "B", "script", "al", "tback"):
        out[k] = np.concatenate(d[k]).astype(np.uint8)
        out[k + "_off"] = np.concatenate([[0], np.cumsum([len(x) for x in d[k]])]).astype(np.int64)
    for k in ("LB", "RB"):
        out[k] = np.concatenate(d[k]).astype(np.int32)
    out["band_off"] = np.concatenate([[0], np.cumsum([len(x) for x in d["LB"]])]).astype(np.int64)
    out["full"] = np.asarray([1 if full else 0], np.int32)
    return out


def small_problems():
    rng = np.random.default_rng(20260117)
    probs = []
    for band in ("smooth", "full", "ragged"):
        for alphabet in ("acgt", "mixed", "weird"):
            for _ in range(6):
                K, L = int(rng.integers(1, 7)), int(rng.integers(1, 7))
                M, N = int(rng.integers(1, 70)), int(rng.integers(1, 70))
                if band == "full" and M * N > 1500:
                    M = max(1, 1500 // N)
                probs.append(random_problem(rng, K, L, M, N, band=band, alphabet=alphabet))
    # shapes straddling the wavefront granularities (32 rows per block, 4 steps per traceback word)
    for M in (1, 2, 31, 32, 33, 64, 65):
        for N in (1, 9, 10, 11, 37):
            probs.append(random_problem(rng, 2, 1, M, N, band="ragged" if (M + N) % 2 else "full"))
    # what pre_yama() really hands over (tools/synth.c mirrors mz_preyama.c:174-259)
    sb = SynthBatch(77, [2, 3, 5, 1, 4], [1, 1, 1, 1, 3], [90, 64, 33, 120, 50], R=12, lower=0.05)
    probs += [tuple(np.array(x) for x in sb.problem(i)) for i in range(sb.n)]
    return probs


def deep_problems():
    sb = SynthBatch(5, [63, 32, 99, 90, 16, 50], [1, 32, 1, 10, 16, 50], [150, 120, 100, 100, 300, 90], R=30, lower=0.02)
    probs = [tuple(np.array(x) for x in sb.problem(i)) for i in range(sb.n)]
    # deliberate int32 wrap: 100*K*L*M beyond 2^31 with -fwrapv semantics
    sb2 = SynthBatch(6, [120], [120], [1800], R=10, sub=0.0, dash=0.0, indel=0.0)
    probs += [tuple(np.array(x) for x in sb2.problem(0))]
    return probs


def smooth_cases(host):
    rng = np.random.default_rng(42)
    cases = []
    for _ in range(40):
        M, N = int(rng.integers(1, 200)), int(rng.integers(1, 200))
        R = int(rng.choice([0, 1, 5, 12, 30, 100, 300]))
        LB = np.zeros(M + 1, np.int32)
        RB = np.full(M + 1, N, np.int32)
        for i in range(1, M + 1):
            if rng.random() < 0.7:
                j = int(np.clip(round(i * N / M + rng.normal(0, 4)), 0, N))
                LB[i] = RB[i] = j
        lo, ro = LB.copy(), RB.copy()
        host.smooth(lo.ctypes.data_as(C.c_void_p), ro.ctypes.data_as(C.c_void_p), M, N, R)
        cases.append((M, N, R, LB, RB, lo, ro))
    return cases


MAF_CASES = (   # name, argv after the tool name (run with cwd = tests/golden/maf)
    ("v1", ["ref.sp1.maf", "ref.sp2.maf", "1", "v1.out1", "v1.out2"]),
    ("v0", ["ref.sp1.maf", "ref.sp2.maf", "0", "v0.out1", "v0.out2"]),
    ("v1r12", ["R=12", "M=50", "ref.sp1.maf", "ref.sp2.maf", "1", "v1r12.out1", "v1r12.out2"]),
    ("step2", ["v1.stdout", "ref.sp3.maf", "1", "step2.out1", "step2.out2"]),
    ("mixed", ["ref.sp1.maf", "ref.sp3.maf", "1"]),          # no out files: unused pieces interleave on stdout
)


def make_maf_golden():
    import subprocess
    from tools.mafsynth import make_dataset
    d = os.path.join(GOLD, "maf")
    os.makedirs(d, exist_ok=True)
    for f in os.listdir(d):
        os.remove(os.path.join(d, f))
    make_dataset(d, ref_len=12_000, n_species=3, seed=11, lower=0.02, blk=(150, 900))
    tool = os.path.join(ROOT, "oracle", "_ref", "bin", "multiz")
    for name, argv in MAF_CASES:
        with open(os.path.join(d, name + ".stdout"), "wb") as f:
            subprocess.run([tool] + argv, cwd=d, stdout=f, check=True)


def make_score_golden(ref):
    # block scoring: the reference's mafScoreRange under both score sets
    from tools.score_cases import score_cases
    cases = score_cases()
    refs = {70: ref, 85: Reference(85)}      # (init_scores70/85 switch the globals of the one library: score per set in turn)
    out = {}
    for which in (70, 85):
        refs[which].lib.ref_init_scores(which)
        out[which] = np.asarray([refs[which].score_range(t, s, n) for (t, s, n) in cases], dtype=np.float64)
    ref.lib.ref_init_scores(70)
    np.savez_compressed(
        os.path.join(GOLD, "score_small.npz"),
        rows=np.asarray([t.shape[0] for t, _, _ in cases], np.int32), cols=np.asarray([t.shape[1] for t, _, _ in cases], np.int32),
        start=np.asarray([s for _, s, _ in cases], np.int32), size=np.asarray([n for _, _, n in cases], np.int32),
        text=np.concatenate([t.ravel() for t, _, _ in cases]),
        text_off=np.concatenate([[0], np.cumsum([t.size for t, _, _ in cases])]).astype(np.int64),
        score70=out[70], score85=out[85])


def main():
    build(quiet=True)
    if not Reference.available():
        raise SystemExit("oracle/_ref/libyama_ref.so missing: needs /root/reference")
    os.makedirs(GOLD, exist_ok=True)
    ref = Reference(70)
    if sys.argv[1:] == ["score"]:            # only the block-scoring fixture
        make_score_golden(ref)
        return
    np.savez_compressed(os.path.join(GOLD, "yama_small.npz"), **pack_problems(ref, small_problems(), full=True))
    np.savez_compressed(os.path.join(GOLD, "yama_deep.npz"), **pack_problems(ref, deep_problems(), full=False))

    sc = {}
    for which in (70, 85):
        r = Reference(which)
        ss, gop, ge = r.tables()
        sc[f"ss{which}"], sc[f"gop{which}"], sc[f"ge{which}"] = ss, gop, np.asarray([ge], np.int32)
    Reference(70)   # leave the shared library's globals on HOXD70
    np.savez_compressed(os.path.join(GOLD, "scores.npz"), **sc)

    make_score_golden(ref)

    host = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libhost_ref.so"))
    host.smooth.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    host.smooth.restype = None
    cases = smooth_cases(host)
    np.savez_compressed(
        os.path.join(GOLD, "smooth.npz"),
        M=np.asarray([c[0] for c in cases], np.int32), N=np.asarray([c[1] for c in cases], np.int32),
        R=np.asarray([c[2] for c in cases], np.int32),
        off=np.concatenate([[0], np.cumsum([c[0] + 1 for c in cases])]).astype(np.int64),
        LB_in=np.concatenate([c[3] for c in cases]), RB_in=np.concatenate([c[4] for c in cases]),
        LB_out=np.concatenate([c[5] for c in cases]), RB_out=np.concatenate([c[6] for c in cases]))
    make_maf_golden()
    for dp, _, fs in sorted(os.walk(GOLD)):
        for f in sorted(fs):
            print(os.path.relpath(os.path.join(dp, f), GOLD), os.path.getsize(os.path.join(dp, f)))


if __name__ == "__main__":
    main()
