"""CPU suite, part 2: the C-ABI library loads, exports every symbol include/yama_b200.h declares, and its
host-only entry points (band validation mz_yama.c:58-71, column assembly mz_yama.c:293-313, the sharding
plan) behave like the reference.  No device compute is attempted here; yb_create must FAIL without a GPU
(there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import multiz_b200
from multiz_b200 import yama as ymod
from golden_util import GoldenYama

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "yama_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(yb_[a-z_0-9]+)\s*\(", src)))


def test_header_and_library_agree():
    lib = multiz_b200.load_library()
    names = _header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"libyama_b200.so does not export {n}"
    assert sorted(multiz_b200.ABI_SYMBOLS) == names


def test_struct_layouts_match_header():
    assert C.sizeof(ymod.yb_job) == 48 and ymod.JOB_DTYPE.itemsize == 48
    assert C.sizeof(ymod.yb_result) == 40 and ymod.RESULT_DTYPE.itemsize == 40


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(multiz_b200.YamaError) as e:
        multiz_b200.YamaB200(devices=[0])
    assert e.value.code == -1           # YB_ERR_CUDA


def test_check_band_matches_reference_wording(oracle):
    lib = multiz_b200.load_library()
    msg = C.create_string_buffer(256)

    def chk(M, N, LB, RB):
        LB, RB = np.asarray(LB, np.int32), np.asarray(RB, np.int32)
        rc = lib.yb_check_band(M, N, LB.ctypes.data, RB.ctypes.data, msg, 256)
        omsg = C.create_string_buffer(256)
        orc = oracle.lib.oracle_check_band(M, N, LB.ctypes.data, RB.ctypes.data, omsg, 256)
        assert (rc < 0) == (orc < 0)
        if rc < 0:
            assert rc == -2 and msg.value == omsg.value
        else:
            assert rc == orc
        return rc

    assert chk(3, 12, [0, 0, 0, 0], [12, 12, 12, 12]) == 52
    assert chk(3, 12, [1, 1, 1, 1], [12, 12, 12, 12]) == -2
    assert chk(3, 12, [0, 0, 0, 0], [12, 5, 12, 12]) == -2
    assert chk(3, 12, [0, 1, 0, 0], [12, 12, 12, 12]) == -2
    assert chk(3, 12, [0, 0, 0, 0], [12, 12, 11, 12]) == -2
    assert chk(2, 4, [0, 0, 0], [4, 4, 4]) == 15                   # N < 10: width needs only N
    rng = np.random.default_rng(0)
    from tools.synth import random_band
    for _ in range(200):
        M, N = int(rng.integers(1, 60)), int(rng.integers(1, 60))
        LB, RB = random_band(rng, M, N, ("smooth", "ragged", "full")[_ % 3])
        if rng.random() < 0.3:
            k = int(rng.integers(0, M + 1))
            RB = RB.copy(); RB[k] = max(0, RB[k] - int(rng.integers(1, 20)))
        chk(M, N, LB, RB)


def test_assemble_matches_reference_columns():
    """yb_assemble (mz_yama.c:293-313) fed with the REFERENCE's edit scripts reproduces its columns."""
    lib = multiz_b200.load_library()
    g = GoldenYama("yama_small.npz")
    for i in range(g.n):
        A, B, LB, RB = (np.ascontiguousarray(x) for x in g.problem(i))
        e = g.expected(i)
        ops = np.concatenate([e["script"], np.zeros((-len(e["script"])) % 4, np.uint8)]).reshape(-1, 4)
        script = np.ascontiguousarray((ops << np.array([0, 2, 4, 6], np.uint8)).sum(axis=1).astype(np.uint8))   # 2 bits/op
        job = ymod.yb_job(A.shape[1], A.shape[0], B.shape[1], B.shape[0], A.ctypes.data, B.ctypes.data,
                          LB.ctypes.data, RB.ctypes.data)
        res = ymod.yb_result(0, e["m_new"], 0, 0, 0, 0, e["cells"], script.ctypes.data)
        out = np.zeros((e["m_new"], A.shape[1] + B.shape[1]), np.uint8)
        assert lib.yb_assemble(C.byref(job), C.byref(res), out.ctypes.data) == 0
        assert np.array_equal(out, e["al"]), i
    # a script that does not consume both alignments is refused (mz_yama.c:310-312)
    unpacked = np.zeros(e["m_new"], np.uint8)
    res = ymod.yb_result(0, e["m_new"], 0, 0, 0, 0, e["cells"], script.ctypes.data)
    assert lib.yb_script_unpack(C.byref(res), unpacked.ctypes.data) == 0 and np.array_equal(unpacked, e["script"])
    bad = np.zeros(1, np.uint8)
    res = ymod.yb_result(0, 1, 0, 0, 0, 0, 0, bad.ctypes.data)
    if A.shape[0] + B.shape[0] > 2:
        assert lib.yb_assemble(C.byref(job), C.byref(res), out.ctypes.data) == -5


def test_plan_split_properties():
    rng = np.random.default_rng(3)
    for _ in range(100):
        n = int(rng.integers(0, 400))
        cells = rng.integers(1, 10 ** int(rng.integers(2, 7)), size=n)
        parts = int(rng.integers(1, 9))
        cuts = multiz_b200.plan_split(cells, parts)
        assert cuts[0] == 0 and cuts[-1] == n and np.all(np.diff(cuts) >= 0)
        if n >= 8 * parts:
            cost = np.add.reduceat(np.concatenate([cells + 2000, [0]]), np.minimum(cuts[:-1], n))[:parts]
            cost = np.where(np.diff(cuts) > 0, cost, 0)
            assert cost.max() <= cost.sum() / parts + (cells.max() + 2000)      # within one pair of ideal
    assert list(multiz_b200.plan_split([100, 200, 50000, 30, 40, 50, 60000], 3)) == [0, 3, 6, 7]


def test_pair_facts_simd_scan_equals_scalar(oracle):
    """yb_pair_facts cross-checks the vectorised host scan (band_scan.cpp) against the scalar restatement of
    mz_yama.c:58-71 and the schedule builder; cells must also equal the oracle's tback_size."""
    from tools.synth import random_band, SynthBatch
    lib = multiz_b200.load_library()
    rng = np.random.default_rng(11)
    msg = C.create_string_buffer(256)
    cells, wmax, nsteps = C.c_int64(), C.c_int32(), C.c_int32()

    def facts(M, N, LB, RB):
        LB, RB = np.ascontiguousarray(LB, np.int32), np.ascontiguousarray(RB, np.int32)
        job = ymod.yb_job(1, M, 1, N, None, None, LB.ctypes.data, RB.ctypes.data)
        return lib.yb_pair_facts(C.byref(job), C.byref(cells), C.byref(wmax), C.byref(nsteps), msg, 256)

    n_bad = 0
    for it in range(600):
        M, N = int(rng.integers(1, 400)), int(rng.integers(1, 400))
        LB, RB = random_band(rng, M, N, ("smooth", "ragged", "full")[it % 3])
        if it % 4 == 3:                       # corrupt: narrow row, non-monotone step or bad terminator
            RB, LB = RB.copy(), LB.copy()
            k = int(rng.integers(0, M + 1))
            which = int(rng.integers(0, 3))
            if which == 0: RB[k] = max(0, int(RB[k]) - int(rng.integers(1, 30)))
            elif which == 1: LB[k] = int(LB[k]) + int(rng.integers(1, 30))
            else: LB[0] = 1
        rc = facts(M, N, LB, RB)
        omsg = C.create_string_buffer(256)
        want = oracle.lib.oracle_check_band(M, N, np.ascontiguousarray(LB, np.int32).ctypes.data,
                                            np.ascontiguousarray(RB, np.int32).ctypes.data, omsg, 256)
        if want < 0:
            n_bad += 1
            assert rc == -2 and msg.value == omsg.value, (it, rc, msg.value, omsg.value)
        else:
            assert rc == 0 and cells.value == want, (it, rc)
            assert wmax.value == int((RB.astype(np.int64) - LB + 1).max()) and nsteps.value % 8 == 0
    assert n_bad > 20
    sb = SynthBatch(8, [3] * 6, [1] * 6, [5000, 4097, 4096, 8193, 33, 64], R=30)     # across the 4096-row flush
    for i in range(sb.n):
        A, B, LB, RB = sb.problem(i)
        assert facts(A.shape[0], B.shape[0], LB, RB) == 0 and cells.value == int((RB.astype(np.int64) - LB + 1).sum())


def test_pair_facts_baseline_build_too():
    """The same cross-check with the non-AVX2 build of the band scan (YB_NO_AVX2=1), in a fresh process."""
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import test_abi_cpu as t\nfrom oracle.oracle_py import Oracle\nt.test_pair_facts_simd_scan_equals_scalar(Oracle(70))\nprint('ok')"
            % (ROOT, os.path.join(ROOT, "tests")))
    p = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, YB_NO_AVX2="1"), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0 and p.stdout.strip().endswith(b"ok"), p.stderr.decode()[-800:]
