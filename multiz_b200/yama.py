"""ctypes mirror of the yama boundary (reference: mz_yama.h:4-22, mz_scores.h:8-15).

Nothing here computes an alignment: every entry point forwards to libyama_b200.so, and the
module raises if that library (the CUDA build) is missing -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

ABI_SYMBOLS = (
    "yb_create", "yb_destroy", "yb_last_error", "yb_device_count", "yb_set_scores",
    "yb_run_batch", "yb_resident_load", "yb_resident_step", "yb_resident_fetch",
    "yb_submit", "yb_flush", "yb_fetch", "yb_clear", "yb_assemble", "yb_check_band", "yb_plan_split", "yb_pair_facts", "yb_script_unpack",
    "yb_score_blocks", "yb_host_alloc", "yb_host_free",
)


class YamaError(RuntimeError):
    """Raised where the reference would fatal()/fatalf() (util.c:17-32); .code is the YB_ERR_* value."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"{msg} (code {code})")
        self.code = code
        self.msg = msg


class yb_job(C.Structure):
    _fields_ = [("K", C.c_int32), ("M", C.c_int32), ("L", C.c_int32), ("N", C.c_int32),
                ("A", C.c_void_p), ("B", C.c_void_p), ("LB", C.c_void_p), ("RB", C.c_void_p)]


class yb_result(C.Structure):
    _fields_ = [("status", C.c_int32), ("m_new", C.c_int32), ("C", C.c_int32), ("D", C.c_int32),
                ("I", C.c_int32), ("reserved", C.c_int32), ("cells", C.c_int64), ("script", C.c_void_p)]


class yb_stats(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
                ("pack_ms", C.c_double), ("total_ms", C.c_double), ("cells", C.c_int64),
                ("pairs", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("kernel_launches", C.c_int32), ("n_devices", C.c_int32), ("fill_ms", C.c_double),
                ("profile_ms", C.c_double), ("traceback_ms", C.c_double), ("plan_ms", C.c_double),
                ("staged_bytes", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


JOB_DTYPE = np.dtype([("K", "<i4"), ("M", "<i4"), ("L", "<i4"), ("N", "<i4"),
                      ("A", "<u8"), ("B", "<u8"), ("LB", "<u8"), ("RB", "<u8")])
RESULT_DTYPE = np.dtype([("status", "<i4"), ("m_new", "<i4"), ("C", "<i4"), ("D", "<i4"), ("I", "<i4"),
                         ("reserved", "<i4"), ("cells", "<i8"), ("script", "<u8")])
assert JOB_DTYPE.itemsize == C.sizeof(yb_job) and RESULT_DTYPE.itemsize == C.sizeof(yb_result)


class yb_block(C.Structure):
    _fields_ = [("nrows", C.c_int32), ("text_size", C.c_int32), ("start", C.c_int32), ("size", C.c_int32),
                ("rows", C.c_void_p)]


BLOCK_DTYPE = np.dtype([("nrows", "<i4"), ("text_size", "<i4"), ("start", "<i4"), ("size", "<i4"), ("rows", "<u8")])
assert BLOCK_DTYPE.itemsize == C.sizeof(yb_block)


# see yb_create: more hardware work queues than the default 8, set before anything in this process initialises CUDA
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def lib_path() -> str:
    return os.environ.get("YAMA_B200_LIB", os.path.join(_HERE, "libyama_b200.so"))


_LIB = None


def load_library():
    """Load libyama_b200.so and declare the prototypes of include/yama_b200.h.  Fails loudly."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} not found: build the CUDA extension first "
                          f"(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
    lib = C.CDLL(path)
    P = C.POINTER
    lib.yb_create.argtypes = [P(C.c_int), C.c_int, P(C.c_void_p)]
    lib.yb_create.restype = C.c_int
    lib.yb_destroy.argtypes = [C.c_void_p]
    lib.yb_destroy.restype = None
    lib.yb_last_error.argtypes = [C.c_void_p]
    lib.yb_last_error.restype = C.c_char_p
    lib.yb_device_count.argtypes = [C.c_void_p]
    lib.yb_device_count.restype = C.c_int
    lib.yb_host_alloc.argtypes = [C.c_void_p, C.c_size_t]
    lib.yb_host_alloc.restype = C.c_void_p
    lib.yb_host_free.argtypes = [C.c_void_p, C.c_void_p]
    lib.yb_host_free.restype = None
    lib.yb_set_scores.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    lib.yb_set_scores.restype = C.c_int
    lib.yb_run_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, P(yb_stats)]
    lib.yb_run_batch.restype = C.c_int
    lib.yb_resident_load.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    lib.yb_resident_load.restype = C.c_int
    lib.yb_resident_step.argtypes = [C.c_void_p, P(yb_stats)]
    lib.yb_resident_step.restype = C.c_int
    lib.yb_resident_fetch.argtypes = [C.c_void_p, C.c_void_p]
    lib.yb_resident_fetch.restype = C.c_int
    lib.yb_submit.argtypes = [C.c_void_p, P(yb_job)]
    lib.yb_submit.restype = C.c_int64
    lib.yb_flush.argtypes = [C.c_void_p, P(yb_stats)]
    lib.yb_flush.restype = C.c_int
    lib.yb_fetch.argtypes = [C.c_void_p, C.c_int64, P(yb_result)]
    lib.yb_fetch.restype = C.c_int
    lib.yb_clear.argtypes = [C.c_void_p]
    lib.yb_clear.restype = None
    lib.yb_assemble.argtypes = [P(yb_job), P(yb_result), C.c_void_p]
    lib.yb_assemble.restype = C.c_int
    lib.yb_check_band.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
    lib.yb_check_band.restype = C.c_int64
    lib.yb_script_unpack.argtypes = [P(yb_result), C.c_void_p]
    lib.yb_script_unpack.restype = C.c_int
    lib.yb_pair_facts.argtypes = [P(yb_job), P(C.c_int64), P(C.c_int32), P(C.c_int32), C.c_char_p, C.c_int]
    lib.yb_pair_facts.restype = C.c_int
    lib.yb_score_blocks.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, P(yb_stats)]
    lib.yb_score_blocks.restype = C.c_int
    lib.yb_plan_split.argtypes = [C.c_int64, C.c_void_p, C.c_int, C.c_void_p]
    lib.yb_plan_split.restype = C.c_int
    _LIB = lib
    return lib


def plan_split(cells, nparts: int) -> np.ndarray:
    """Contiguous, cell-balanced cut of a reference-ordered job list into `nparts` ranges (yb_plan_split).
    Host-only: usable without a GPU (e.g. by every rank of a multi-process run to find its own range)."""
    cells = np.ascontiguousarray(cells, dtype=np.int64)
    cuts = np.zeros(nparts + 1, dtype=np.int64)
    rc = load_library().yb_plan_split(len(cells), cells.ctypes.data, int(nparts), cuts.ctypes.data)
    if rc != 0:
        raise YamaError(rc, "yb_plan_split: bad arguments")
    return cuts


def hox70_tables(which: int = 70):
    """The tables init_scores70/85 build (mz_scores.c:9-27,34-81): (ss[128,128], gop[16], gap_extend)."""
    m70 = [[91, -114, -31, -123], [-114, 100, -125, -31], [-31, -125, 100, -114], [-123, -31, -114, 91]]
    m85 = [[86, -135, -68, -157], [-135, 100, -148, -68], [-68, -148, 100, -135], [-157, -68, -135, 86]]
    mat, gop_open, gap_ext = (m85, 600, 50) if which == 85 else (m70, 400, 30)
    ss = np.full((128, 128), -100, dtype=np.int32)
    for i, a in enumerate("ACGT"):
        for j, b in enumerate("ACGT"):
            for x in (a, a.lower()):
                for y in (b, b.lower()):
                    ss[ord(x), ord(y)] = mat[i][j]
    ss[ord("-"), :] = -gap_ext
    ss[:, ord("-")] = -gap_ext
    ss[ord("-"), ord("-")] = 0
    gop = np.zeros(16, dtype=np.int32)
    gop[[1, 2, 6, 9, 13, 14]] = gop_open
    return ss, gop, gap_ext


def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint8)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


class YamaB200:
    """Context over one or more B200s.  `yama()` mirrors the reference call one pair at a time;
    `run_batch()` is the batched product path (thousands of pairs per launch)."""

    def __init__(self, devices: Sequence[int] | None = None, scores: int | tuple = 70):
        self.lib = load_library()
        h = C.c_void_p()
        if devices is None:
            rc = self.lib.yb_create(None, 0, C.byref(h))
        else:
            arr = (C.c_int * len(devices))(*devices)
            rc = self.lib.yb_create(arr, len(devices), C.byref(h))
        if rc != 0:
            raise YamaError(rc, "yb_create failed: no usable CUDA device (this library has no CPU path)")
        self.h = h
        self._keep = None
        if isinstance(scores, int):
            scores = hox70_tables(scores)
        self.set_scores(*scores)

    def close(self):
        if getattr(self, "h", None):
            self.lib.yb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self, rc):
        return YamaError(rc, self.lib.yb_last_error(self.h).decode())

    @property
    def n_devices(self) -> int:
        return self.lib.yb_device_count(self.h)

    def set_scores(self, ss, gop, gap_extend):
        ss = _i32(ss).reshape(128, 128)
        gop = _i32(gop).reshape(16)
        rc = self.lib.yb_set_scores(self.h, ss.ctypes.data, gop.ctypes.data, int(gap_extend))
        if rc != 0:
            raise self._err(rc)

    # ---- pinned host memory (yb_host_alloc) -------------------------------------------------
    def host_array(self, nbytes: int) -> np.ndarray:
        """A uint8 array over a pinned block of the context: inputs built inside it are copied to the device as they
        are, without a pass over them on the host.  Lives until close()."""
        p = self.lib.yb_host_alloc(self.h, max(1, int(nbytes)))
        if not p:
            raise YamaError(-1, "yb_host_alloc failed")
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(max(1, int(nbytes)),))

    def pin_pools(self, jobs: np.ndarray, pools) -> np.ndarray:
        """jobs whose A / B / LB / RB pointers point into the four contiguous arrays `pools` (as tools.synth.SynthBatch
        lays them out) -> the same jobs over pinned copies of those arrays."""
        out = jobs.copy()
        for name, arr in zip(("A", "B", "LB", "RB"), pools):
            raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
            dst = self.host_array(raw.nbytes)
            dst[:raw.nbytes] = raw
            out[name] = jobs[name] - np.uint64(arr.ctypes.data) + np.uint64(dst.ctypes.data)
        return out

    # ---- batch -----------------------------------------------------------------------------
    @staticmethod
    def make_jobs(problems) -> tuple[np.ndarray, list]:
        """problems: iterable of (A[M,K] u8, B[N,L] u8, LB[M+1], RB[M+1]) -> (job array, keep-alive)."""
        keep = []
        jobs = np.zeros(len(problems), dtype=JOB_DTYPE)
        for i, (A, B, LB, RB) in enumerate(problems):
            A, B, LB, RB = _u8(A), _u8(B), _i32(LB), _i32(RB)
            keep.append((A, B, LB, RB))
            jobs[i] = (A.shape[1], A.shape[0], B.shape[1], B.shape[0],
                       A.ctypes.data, B.ctypes.data, LB.ctypes.data, RB.ctypes.data)
        return jobs, keep

    def run_batch(self, jobs: np.ndarray, check: bool = True, out: np.ndarray | None = None):
        """jobs: array of JOB_DTYPE (pointers into live host memory).  Returns (results, stats).  `out` lets a caller
        reuse its result array across calls, as a C caller of yb_run_batch would."""
        assert jobs.dtype == JOB_DTYPE and jobs.flags.c_contiguous
        if out is None:
            res = np.zeros(len(jobs), dtype=RESULT_DTYPE)
        else:
            assert out.dtype == RESULT_DTYPE and out.flags.c_contiguous and len(out) == len(jobs)
            res = out
        st = yb_stats()
        rc = self.lib.yb_run_batch(self.h, len(jobs), jobs.ctypes.data, res.ctypes.data, C.byref(st))
        if rc != 0 and check:
            raise self._err(rc)
        return res, st

    def resident_load(self, jobs: np.ndarray):
        assert jobs.dtype == JOB_DTYPE and jobs.flags.c_contiguous
        self._keep = jobs
        rc = self.lib.yb_resident_load(self.h, len(jobs), jobs.ctypes.data)
        if rc != 0:
            raise self._err(rc)

    def resident_step(self) -> yb_stats:
        st = yb_stats()
        rc = self.lib.yb_resident_step(self.h, C.byref(st))
        if rc != 0:
            raise self._err(rc)
        return st

    def resident_fetch(self) -> np.ndarray:
        res = np.zeros(len(self._keep), dtype=RESULT_DTYPE)
        rc = self.lib.yb_resident_fetch(self.h, res.ctypes.data)
        if rc != 0:
            raise self._err(rc)
        return res

    @staticmethod
    def script_of(res_row) -> np.ndarray:
        """The edit script as one byte per op (the reference's script[], reversed order), from the packed result."""
        n = int(res_row["m_new"])
        if n == 0 or not int(res_row["script"]):
            return np.zeros(0, dtype=np.uint8)
        packed = np.ctypeslib.as_array(C.cast(int(res_row["script"]), C.POINTER(C.c_uint8)), shape=((n + 3) // 4,))
        ops = (packed[:, None] >> np.array([0, 2, 4, 6], dtype=np.uint8)[None, :]) & 3
        return ops.reshape(-1)[:n].astype(np.uint8)

    def assemble(self, job_row, res_row) -> np.ndarray:
        """Column assembly of mz_yama.c:293-313 -> array [m_new, K+L]."""
        j = yb_job(int(job_row["K"]), int(job_row["M"]), int(job_row["L"]), int(job_row["N"]),
                   int(job_row["A"]), int(job_row["B"]), int(job_row["LB"]), int(job_row["RB"]))
        r = yb_result(int(res_row["status"]), int(res_row["m_new"]), int(res_row["C"]), int(res_row["D"]),
                      int(res_row["I"]), 0, int(res_row["cells"]), int(res_row["script"]))
        out = np.empty((r.m_new, j.K + j.L), dtype=np.uint8)
        rc = self.lib.yb_assemble(C.byref(j), C.byref(r), out.ctypes.data)
        if rc != 0:
            raise YamaError(rc, "new_align: edit script does not consume both alignments")
        return out

    # ---- block scoring (mafScoreRange, mz_scores.c:124-152) -------------------------------------
    @staticmethod
    def make_blocks(blocks) -> tuple[np.ndarray, list]:
        """blocks: iterable of (text[rows, textSize] u8, start, size) -> (block array, keep-alive)."""
        keep = []
        arr = np.zeros(len(blocks), dtype=BLOCK_DTYPE)
        for i, (text, start, size) in enumerate(blocks):
            text = _u8(text)
            if text.ndim != 2:
                raise ValueError("a block is a [rows, columns] byte matrix")
            ptrs = np.array([text.ctypes.data + j * text.strides[0] for j in range(text.shape[0])], dtype=np.uint64)
            keep.append((text, ptrs))
            arr[i] = (text.shape[0], text.shape[1], start, size, ptrs.ctypes.data if len(ptrs) else 0)
        return arr, keep

    def score_blocks(self, blocks: np.ndarray, check: bool = True):
        """blocks: array of BLOCK_DTYPE -> (scores float64[n], stats)."""
        assert blocks.dtype == BLOCK_DTYPE and blocks.flags.c_contiguous
        scores = np.zeros(len(blocks), dtype=np.float64)
        st = yb_stats()
        rc = self.lib.yb_score_blocks(self.h, len(blocks), blocks.ctypes.data, scores.ctypes.data, C.byref(st))
        if rc != 0 and check:
            raise self._err(rc)
        return scores, st

    def mafScoreRange(self, text, start: int, size: int) -> float:
        """mafScoreRange(maf, start, size) of one block given as its [rows, textSize] text (mz_scores.c:124)."""
        arr, keep = self.make_blocks([(text, start, size)])
        sc, _ = self.score_blocks(arr)
        del keep
        return float(sc[0])

    # ---- the reference's own call shape -------------------------------------------------------
    def yama(self, A, K, M, B, L, N, LB, RB):
        """yama(A,K,M,B,L,N,LB,RB) -> (AL_new[m_new, K+L], m_new), as mz_yama.h:22.
        A is [M,K] (column i of the reference = A[i-1]), B is [N,L]."""
        A = _u8(A).reshape(M, K)
        B = _u8(B).reshape(N, L)
        jobs, keep = self.make_jobs([(A, B, LB, RB)])
        res, _ = self.run_batch(jobs)
        al = self.assemble(jobs[0], res[0])
        del keep
        return al, int(res[0]["m_new"])
