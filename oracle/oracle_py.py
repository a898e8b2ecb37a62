"""ctypes loaders for the CHECKERS in oracle/ (test infrastructure -- never imported by multiz_b200).

  Oracle     liboracle.so         our plain-C restatements (yama_oracle.c, score_oracle.c)
  Reference  _ref/libyama_ref.so  the unmodified reference yama() (mz_yama.c) and mafScoreRange (mz_scores.c)
                                  behind ref_hook.c / ref_score_hook.c
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def build(quiet=True):
    subprocess.run(["make", "-C", HERE] + (["-s"] if quiet else []), check=True)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _row_table(block):
    """block: uint8 [nrows, text_size] -> (contiguous copy, ctypes array of row pointers)."""
    blk = _u8(block)
    if blk.ndim != 2:
        raise ValueError("a block is a [rows, columns] byte matrix")
    ptrs = (C.c_void_p * max(1, blk.shape[0]))(*[blk.ctypes.data + j * blk.strides[0] for j in range(blk.shape[0])])
    return blk, ptrs


def cells_of(LB, RB) -> int:
    return int((np.asarray(RB, dtype=np.int64) - np.asarray(LB, dtype=np.int64) + 1).sum())


class Oracle:
    def __init__(self, which: int = 70):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        self.lib.oracle_yama.restype = C.c_int
        self.lib.oracle_yama.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        self.lib.oracle_smooth.restype = None
        self.lib.oracle_smooth.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        self.lib.oracle_scores.restype = None
        self.lib.oracle_scores.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        self.lib.oracle_check_band.restype = C.c_long
        self.lib.oracle_check_band.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        self.lib.oracle_score_range.restype = C.c_int
        self.lib.oracle_score_range.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                C.POINTER(C.c_double)]
        self.set_scores(which)

    def score_range(self, block, start, size):
        """mafScoreRange of a [rows, textSize] block; ValueError(message) for a bad range."""
        blk, ptrs = _row_table(block)
        out = C.c_double()
        rc = self.lib.oracle_score_range(blk.shape[0], ptrs, blk.shape[1], start, size, self.ss.ctypes.data,
                                         self.gop.ctypes.data, C.byref(out))
        if rc:
            raise ValueError("mafScoreRange: start = %d, size = %d, textSize = %d\n" % (start, size, blk.shape[1]))
        return out.value

    def set_scores(self, which):
        self.ss = np.zeros((128, 128), dtype=np.int32)
        self.gop = np.zeros(16, dtype=np.int32)
        ge = C.c_int()
        self.lib.oracle_scores(which, self.ss.ctypes.data, self.gop.ctypes.data, C.byref(ge))
        self.gap_ext = ge.value

    def smooth(self, LB, RB, M, N, radius):
        LB, RB = _i32(LB).copy(), _i32(RB).copy()
        self.lib.oracle_smooth(LB.ctypes.data, RB.ctypes.data, M, N, radius)
        return LB, RB

    def yama(self, A, B, LB, RB, want_tback=True):
        """A [M,K], B [N,L] -> dict(al, m_new, tback, cdi, script) or raises ValueError(msg)."""
        A, B, LB, RB = _u8(A), _u8(B), _i32(LB), _i32(RB)
        M, K = A.shape
        N, L = B.shape
        msg = C.create_string_buffer(512)
        ncell = self.lib.oracle_check_band(M, N, LB.ctypes.data, RB.ctypes.data, msg, 512)
        if ncell < 0:
            raise ValueError(msg.value.decode())
        out = np.zeros(((M + N), K + L), dtype=np.uint8)
        tb = np.zeros(ncell if want_tback else 1, dtype=np.uint8)
        cdi = np.zeros(3, dtype=np.int32)
        script = np.zeros(M + N + 1, dtype=np.uint8)
        m = self.lib.oracle_yama(A.ctypes.data, K, M, B.ctypes.data, L, N, LB.ctypes.data, RB.ctypes.data,
                                 self.ss.ctypes.data, self.gop.ctypes.data, self.gap_ext, out.ctypes.data,
                                 tb.ctypes.data if want_tback else None, cdi.ctypes.data, script.ctypes.data,
                                 msg, 512)
        if m < 0:
            raise ValueError(msg.value.decode())
        return dict(al=out[:m].copy(), m_new=m, tback=tb, cdi=cdi, script=script[:m].copy(), cells=ncell)


class Reference:
    """The reference's own yama(), compiled from /root/reference into oracle/_ref (travels to the GPU box)."""

    def __init__(self, which: int = 70):
        path = os.path.join(HERE, "_ref", "libyama_ref.so")
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.lib.ref_yama.restype = C.c_int
        self.lib.ref_yama.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib.ref_init_scores.argtypes = [C.c_int]
        self.lib.ref_get_tables.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        self.lib.ref_score_range.restype = C.c_double
        self.lib.ref_score_range.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        self.lib.ref_init_scores(which)

    def score_range(self, block, start, size):
        """The reference's mafScoreRange; the range must be valid (the reference exit(1)s otherwise)."""
        blk, ptrs = _row_table(block)
        assert 0 <= start and size > 0 and start + size <= blk.shape[1]
        return float(self.lib.ref_score_range(blk.shape[0], ptrs, blk.shape[1], start, size))

    @staticmethod
    def available() -> bool:
        return os.path.exists(os.path.join(HERE, "_ref", "libyama_ref.so"))

    def tables(self):
        ss = np.zeros((128, 128), dtype=np.int32)
        gop = np.zeros(16, dtype=np.int32)
        ge = C.c_int()
        self.lib.ref_get_tables(ss.ctypes.data, gop.ctypes.data, C.byref(ge))
        return ss, gop, ge.value

    def yama(self, A, B, LB, RB, want_tback=True):
        """Inputs must pass the band validation (the reference exit(1)s otherwise)."""
        A, B, LB, RB = _u8(A), _u8(B), _i32(LB).copy(), _i32(RB).copy()
        M, K = A.shape
        N, L = B.shape
        ncell = cells_of(LB, RB)
        out = np.zeros(((M + N), K + L), dtype=np.uint8)
        tb = np.zeros(ncell if want_tback else 1, dtype=np.uint8)
        cdi = np.zeros(3, dtype=np.int32)
        script = np.zeros(M + N + 1, dtype=np.uint8)
        m = self.lib.ref_yama(A.ctypes.data, K, M, B.ctypes.data, L, N, LB.ctypes.data, RB.ctypes.data,
                              out.ctypes.data, tb.ctypes.data if want_tback else None, cdi.ctypes.data,
                              script.ctypes.data)
        return dict(al=out[:m].copy(), m_new=m, tback=tb, cdi=cdi, script=script[:m].copy(), cells=ncell)
