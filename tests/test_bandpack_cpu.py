"""CPU suite: the host half of the delta-coded bands (multiz_b200/csrc/band_scan.cpp:yb_band_pack, both builds) against a
numpy model, and a numpy restatement of what yb_band_expand (plan_kernels.cuh) does with the bytes -- the round trip must
give back the caller's ints for ANY int array (valid bands, steps of 254 / 255 / 256, negative steps, int32 wrap), because
K0 validates the restored rows in the reference's words.  Also the job dump the benchmark's cfg2real workload is made of."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from multiz_b200 import load_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pack(lib, X):
    X = np.ascontiguousarray(X, dtype=np.int32)
    out = np.zeros(len(X), dtype=np.uint8)
    fn = lib.yb_band_pack
    fn.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    fn.restype = C.c_int
    marked = fn(len(X) - 1, X.ctypes.data, out.ctypes.data)
    return out, marked


def expand_model(first, deltas, exceptions):
    """yb_band_expand: prefix sum of the bytes from X[0]; a row listed as an exception adds (step - 255) to itself and
    every later row.  uint32 arithmetic, like the kernel's."""
    v = (np.uint32(first) + np.cumsum(deltas.astype(np.uint64))).astype(np.uint64)
    for r, step in exceptions:
        v[r:] += np.uint64((int(step) - 255) & 0xffffffff)
    return (v & np.uint64(0xffffffff)).astype(np.uint32).view(np.int32)


CASES = {
    "diagonal": lambda rng: np.cumsum(rng.integers(0, 3, 700)).astype(np.int32),
    "flat": lambda rng: np.zeros(130, np.int32),
    "one_row": lambda rng: np.array([0, 5], np.int32),
    "steps_254_255_256": lambda rng: np.cumsum(np.array([0, 1, 254, 1, 255, 1, 256, 2, 1000, 0, 254, 255], np.int64)).astype(np.int32),
    "decreasing": lambda rng: np.array([0, 4, 3, 3, 10, 2, 2, 300, 299], np.int32),
    "wrap": lambda rng: np.array([0, 2**31 - 1, -2**31, -1, 0, 7], np.int64).astype(np.int32),
    "random": lambda rng: rng.integers(-2**31, 2**31 - 1, 257).astype(np.int32),
    "long": lambda rng: np.cumsum(rng.choice([0, 1, 1, 2, 300], 10_001)).astype(np.int32),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("build", ["dispatch", "generic"])
def test_band_pack_round_trip(name, build):
    if build == "generic":
        # the baseline (non-AVX2) build runs in a process of its own: the dispatcher reads YB_NO_AVX2 once
        code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import test_bandpack_cpu as t; "
                "t.test_band_pack_round_trip(%r, 'dispatch'); print('ok')") % (ROOT, os.path.join(ROOT, "tests"), name)
        p = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, YB_NO_AVX2="1"), capture_output=True, text=True, timeout=120)
        assert p.returncode == 0 and "ok" in p.stdout, p.stderr[-400:]
        return
    lib = load_library()
    X = CASES[name](np.random.default_rng(3))
    got, marked = pack(lib, X)
    step = (X[1:].astype(np.int64) - X[:-1].astype(np.int64)) & 0xffffffff       # uint32 difference
    want = np.concatenate([[0], np.where(step < 255, step, 255)]).astype(np.uint8)
    assert np.array_equal(got, want)
    assert marked == int((step >= 255).sum())
    exc = [(r + 1, int(np.int32(np.uint32(step[r])))) for r in range(len(step)) if step[r] >= 255]
    assert np.array_equal(expand_model(X[0], got, exc), X)


def test_job_dump_of_a_real_merge_round_trips(tmp_path, oracle):
    """YB_DUMP_JOBS (integration/yama_dropin.cpp) -> tools.synth.RecordedBatch: what bench.py --workload cfg2real replays.  The
    dump must hold exactly the jobs yama() received -- the oracle aligns a few of them and the cell count matches the
    drop-in's own statistics."""
    import re
    import shutil
    from dropin_util import GOLD_MAF, SHIM_MULTIZ
    from tools.synth import RecordedBatch
    if not os.path.exists(SHIM_MULTIZ):
        pytest.skip("integration/_ref/bin/multiz_shim not built")
    for f in ("ref.sp1.maf", "ref.sp2.maf"):
        shutil.copy(os.path.join(GOLD_MAF, f), tmp_path)
    dump = str(tmp_path / "jobs.bin")
    p = subprocess.run([SHIM_MULTIZ, "ref.sp1.maf", "ref.sp2.maf", "1", "o1", "o2"], cwd=tmp_path, capture_output=True,
                       env=dict(os.environ, YB_DUMP_JOBS=dump, YB_DROPIN_STATS="1"), timeout=300)
    assert p.returncode == 0, p.stderr[-300:]
    st = {k: int(v) for k, v in re.findall(rb"(jobs|cells)=(\d+)", p.stderr)}
    st = {k.decode(): v for k, v in st.items()}
    rb = RecordedBatch(dump)
    assert rb.n == st["jobs"] and rb.cells == st["cells"] and rb.n > 0
    for i in (0, rb.n // 2, rb.n - 1):
        A, B, LB, RB = rb.problem(i)
        o = oracle.yama(A, B, LB, RB, want_tback=False)
        assert o["cells"] == int((RB.astype(np.int64) - LB + 1).sum()) and o["m_new"] >= max(A.shape[0], B.shape[0])
