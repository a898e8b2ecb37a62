"""Readers for tests/golden/*.npz (written by tools/make_golden.py from the compiled reference)."""
import hashlib
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def digest(a) -> np.ndarray:
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


class GoldenYama:
    """problems[i] = (A[M,K], B[N,L], LB, RB); expected(i) = dict(cdi, m_new, script, al, tback, cells).
    For hashed sets (`full` False) `al` and `tback` hold sha256 digests."""

    def __init__(self, name):
        self.z = np.load(os.path.join(GOLD, name))
        self.n = len(self.z["K"])
        self.full = bool(self.z["full"][0])

    def _slice(self, key, i):
        off = self.z[key + "_off"]
        return self.z[key][off[i]:off[i + 1]]

    def problem(self, i):
        z = self.z
        K, M, L, N = (int(z[k][i]) for k in ("K", "M", "L", "N"))
        A = self._slice("A", i).reshape(M, K)
        B = self._slice("B", i).reshape(N, L)
        bo = z["band_off"]
        return A, B, z["LB"][bo[i]:bo[i + 1]], z["RB"][bo[i]:bo[i + 1]]

    def problems(self):
        return [self.problem(i) for i in range(self.n)]

    def expected(self, i):
        z = self.z
        K, L = int(z["K"][i]), int(z["L"][i])
        m = int(z["m_new"][i])
        al = self._slice("al", i)
        return dict(cdi=z["cdi"][i], m_new=m, script=self._slice("script", i),
                    al=al.reshape(m, K + L) if self.full else al, tback=self._slice("tback", i),
                    cells=int(z["cells"][i]))

    def check(self, i, got, tback=True):
        """got: dict with cdi, m_new, script, al and optionally tback (as oracle_py returns)."""
        e = self.expected(i)
        assert tuple(int(x) for x in got["cdi"]) == tuple(int(x) for x in e["cdi"]), ("cdi", i)
        assert int(got["m_new"]) == e["m_new"], ("m_new", i)
        assert np.array_equal(got["script"], e["script"]), ("script", i)
        if self.full:
            assert np.array_equal(got["al"], e["al"]), ("al", i)
            if tback and "tback" in got:
                assert np.array_equal(got["tback"], e["tback"]), ("tback", i)
        else:
            assert np.array_equal(digest(got["al"]), e["al"]), ("al", i)
            if tback and "tback" in got:
                assert np.array_equal(digest(got["tback"]), e["tback"]), ("tback", i)
        if "cells" in got:
            assert int(got["cells"]) == e["cells"]


class GoldenScores:
    """tests/golden/score_small.npz: blocks[i] = (text[rows, cols], start, size); score70/score85 = the reference's
    mafScoreRange (mz_scores.c:124-152) under init_scores70 / init_scores85."""

    def __init__(self, name="score_small.npz"):
        self.z = np.load(os.path.join(GOLD, name))
        self.n = len(self.z["rows"])

    def block(self, i):
        z = self.z
        off = z["text_off"]
        return (z["text"][off[i]:off[i + 1]].reshape(int(z["rows"][i]), int(z["cols"][i])), int(z["start"][i]), int(z["size"][i]))

    def blocks(self):
        return [self.block(i) for i in range(self.n)]

    def expected(self, which=70):
        return self.z[f"score{which}"]
