"""Synthetic yama() problems (measurement / test infrastructure).

`SynthBatch` wraps tools/synth.c: block pairs shaped the way pre_yama() hands them to yama()
(mz_preyama.c:174-259), deterministic in (seed, index, shape).  `random_problem` is a slow numpy
generator used by the tests for adversarial band shapes the simulator never produces.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
JOB_DTYPE = np.dtype([("K", "<i4"), ("M", "<i4"), ("L", "<i4"), ("N", "<i4"),
                      ("A", "<u8"), ("B", "<u8"), ("LB", "<u8"), ("RB", "<u8")])


class _SB(C.Structure):
    _fields_ = [("n", C.c_int64), ("K", C.c_void_p), ("M", C.c_void_p), ("L", C.c_void_p), ("N", C.c_void_p),
                ("offA", C.c_void_p), ("offB", C.c_void_p), ("offBand", C.c_void_p),
                ("A", C.c_void_p), ("B", C.c_void_p), ("LB", C.c_void_p), ("RB", C.c_void_p),
                ("bytesA", C.c_int64), ("bytesB", C.c_int64), ("nBand", C.c_int64), ("cells", C.c_int64)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "libsynth.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        _lib = C.CDLL(path)
        _lib.synth_make.restype = C.POINTER(_SB)
        _lib.synth_make.argtypes = [C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_double, C.c_double, C.c_double, C.c_double]
        _lib.synth_free.argtypes = [C.POINTER(_SB)]
    return _lib


def _view(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    ct = {np.int32: C.c_int32, np.int64: C.c_int64, np.uint8: C.c_uint8}[dtype]
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(int(n),))


class SynthBatch:
    def __init__(self, seed, Ks, Ls, Ms, R=30, sub=0.10, dash=0.05, indel=0.01, lower=0.0):
        lib = _load()
        Ks = np.ascontiguousarray(Ks, dtype=np.int32)
        Ls = np.ascontiguousarray(Ls, dtype=np.int32)
        Ms = np.ascontiguousarray(Ms, dtype=np.int32)
        n = len(Ks)
        self._p = lib.synth_make(int(seed), n, Ks.ctypes.data, Ls.ctypes.data, Ms.ctypes.data, int(R),
                                 float(sub), float(dash), float(indel), float(lower))
        s = self._p.contents
        self.n = n
        self.R = R
        self.K = _view(s.K, n, np.int32)
        self.M = _view(s.M, n, np.int32)
        self.L = _view(s.L, n, np.int32)
        self.N = _view(s.N, n, np.int32)
        self.offA = _view(s.offA, n, np.int64)
        self.offB = _view(s.offB, n, np.int64)
        self.offBand = _view(s.offBand, n, np.int64)
        self.A = _view(s.A, s.bytesA, np.uint8)
        self.B = _view(s.B, s.bytesB, np.uint8)
        self.LB = _view(s.LB, s.nBand, np.int32)
        self.RB = _view(s.RB, s.nBand, np.int32)
        self.cells = int(s.cells)
        jobs = np.zeros(n, dtype=JOB_DTYPE)
        jobs["K"], jobs["M"], jobs["L"], jobs["N"] = self.K, self.M, self.L, self.N
        jobs["A"] = np.uint64(s.A or 0) + self.offA.astype(np.uint64)
        jobs["B"] = np.uint64(s.B or 0) + self.offB.astype(np.uint64)
        jobs["LB"] = np.uint64(s.LB or 0) + (self.offBand * 4).astype(np.uint64)
        jobs["RB"] = np.uint64(s.RB or 0) + (self.offBand * 4).astype(np.uint64)
        self.jobs = jobs

    def problem(self, i):
        """(A[M,K], B[N,L], LB, RB) views of pair i."""
        K, M, L, N = int(self.K[i]), int(self.M[i]), int(self.L[i]), int(self.N[i])
        A = self.A[self.offA[i]:self.offA[i] + K * M].reshape(M, K)
        B = self.B[self.offB[i]:self.offB[i] + L * N].reshape(N, L)
        LB = self.LB[self.offBand[i]:self.offBand[i] + M + 1]
        RB = self.RB[self.offBand[i]:self.offBand[i] + M + 1]
        return A, B, LB, RB

    def cells_per_pair(self):
        w = (self.RB.astype(np.int64) - self.LB + 1)
        idx = np.concatenate([self.offBand, [len(self.LB)]])
        cs = np.concatenate([[0], np.cumsum(w)])
        return cs[idx[1:]] - cs[idx[:-1]]

    def close(self):
        if self._p:
            _load().synth_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


ALPHABETS = {
    "acgt": b"ACGT",
    "mixed": b"ACGTacgtNn",
    "weird": b"ACGTacgtNnXRY.*",
}


def _valid_band(LB, RB, M, N):
    need = min(N, 10)
    return (LB[0] == 0 and RB[M] == N and np.all(RB - LB >= need) and np.all(np.diff(LB) >= 0)
            and np.all(np.diff(RB) >= 0))


def random_band(rng, M, N, kind):
    """A band that passes mz_yama.c:58-71: LB[0]=0, RB[M]=N, monotone, width >= min(N,10)."""
    for _ in range(50):
        LB, RB = _random_band(rng, M, N, kind)
        if _valid_band(LB.astype(np.int64), RB.astype(np.int64), M, N):
            return LB, RB
    return np.zeros(M + 1, np.int32), np.full(M + 1, N, np.int32)


def _random_band(rng, M, N, kind):
    need = min(N, 10)
    if kind == "full":
        return np.zeros(M + 1, np.int32), np.full(M + 1, N, np.int32)
    if kind == "smooth":
        R = int(rng.integers(max(need, 1), 40))
        LB = np.zeros(M + 1, np.int32)
        RB = np.full(M + 1, N, np.int32)
        for i in range(1, M + 1):
            if rng.random() < 0.8:
                j = int(round(i * N / M + rng.normal(0, 3)))
                LB[i] = RB[i] = min(max(j, 1), N)
        # same steps as mz_preyama.c:17-35 (host logic, restated for the tests)
        rad = min(M, R)
        LB = np.maximum.accumulate(LB)
        RB = np.minimum.accumulate(RB[::-1])[::-1].copy()
        L2, R2 = LB.copy(), RB.copy()
        for i in range(M, rad, -1):
            L2[i] = min(max(LB[i] - rad, 0), LB[i - rad])
        L2[:rad + 1] = 0
        for i in range(0, M - rad):
            R2[i] = max(min(RB[i] + rad, N), RB[i + rad])
        R2[max(M - rad, 0):] = N
        return L2.astype(np.int32), R2.astype(np.int32)
    # "ragged": random monotone staircase with minimal legal widths and repeated bounds
    LB = np.zeros(M + 1, np.int64)
    RB = np.zeros(M + 1, np.int64)
    lo = 0
    for r in range(M + 1):
        if r > 0 and rng.random() < 0.6:
            lo = min(lo + int(rng.integers(0, 4)), max(N - need, 0))
        LB[r] = lo
    hi = N
    for r in range(M, -1, -1):
        RB[r] = max(hi, LB[r] + need)
        if rng.random() < 0.6:
            hi = max(hi - int(rng.integers(0, 4)), 0)
    RB = np.minimum(np.maximum.accumulate(RB), N)
    RB[M] = N
    RB = np.maximum(RB, LB + need)
    RB = np.minimum(RB, N)
    return LB.astype(np.int32), RB.astype(np.int32)


def random_problem(rng, K, L, M, N, band="smooth", alphabet="mixed", dash=0.15):
    alpha = np.frombuffer(ALPHABETS[alphabet], dtype=np.uint8)
    A = alpha[rng.integers(0, len(alpha), size=(M, K))].copy()
    B = alpha[rng.integers(0, len(alpha), size=(N, L))].copy()
    A[rng.random((M, K)) < dash] = ord("-")
    B[rng.random((N, L)) < dash] = ord("-")
    # correlated columns so that diagonal moves actually win sometimes
    n = min(M, N)
    copy = rng.random(n) < 0.6
    for i in np.nonzero(copy)[0]:
        B[i, :] = A[i, rng.integers(0, K, size=L)]
    LB, RB = random_band(rng, M, N, band)
    return A, B, LB, RB


class RecordedBatch:
    """The yama() jobs of one real multiz invocation, as the drop-in dumped them (YB_DUMP_JOBS, integration/yama_dropin.cpp):
    same attributes as SynthBatch (jobs, A, B, LB, RB, n, cells, K, M, L, N, problem())."""

    def __init__(self, path):
        raw = np.fromfile(path, dtype=np.uint8)
        assert raw[:4].tobytes() == b"YBJ1", path
        n = int(raw[4:12].view(np.uint64)[0])
        dims = raw[12:12 + 16 * n].view(np.int32).reshape(n, 4)
        self.n = n
        self.K, self.M, self.L, self.N = (np.ascontiguousarray(dims[:, k]) for k in range(4))
        K, M, L, N = (x.astype(np.int64) for x in (self.K, self.M, self.L, self.N))
        szA, szB, szBand = K * M, L * N, M + 1
        at = 12 + 16 * n
        self.A = raw[at:at + int(szA.sum())].copy(); at += int(szA.sum())
        self.B = raw[at:at + int(szB.sum())].copy(); at += int(szB.sum())
        nb = int(szBand.sum())
        self.LB = raw[at:at + 4 * nb].copy().view(np.int32); at += 4 * nb
        self.RB = raw[at:at + 4 * nb].copy().view(np.int32); at += 4 * nb
        assert at == len(raw), (at, len(raw))
        self.offA = np.concatenate([[0], np.cumsum(szA)[:-1]]).astype(np.int64)
        self.offB = np.concatenate([[0], np.cumsum(szB)[:-1]]).astype(np.int64)
        self.offBand = np.concatenate([[0], np.cumsum(szBand)[:-1]]).astype(np.int64)
        self.cells = int((self.RB.astype(np.int64) - self.LB + 1).sum())
        jobs = np.zeros(n, dtype=JOB_DTYPE)
        jobs["K"], jobs["M"], jobs["L"], jobs["N"] = self.K, self.M, self.L, self.N
        jobs["A"] = np.uint64(self.A.ctypes.data) + self.offA.astype(np.uint64)
        jobs["B"] = np.uint64(self.B.ctypes.data) + self.offB.astype(np.uint64)
        jobs["LB"] = np.uint64(self.LB.ctypes.data) + (self.offBand * 4).astype(np.uint64)
        jobs["RB"] = np.uint64(self.RB.ctypes.data) + (self.offBand * 4).astype(np.uint64)
        self.jobs = jobs

    problem = SynthBatch.problem
    cells_per_pair = SynthBatch.cells_per_pair

    def close(self):
        pass
