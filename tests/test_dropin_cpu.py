"""CPU suite, part 4: host logic of the drop-in (integration/yama_dropin.cpp).  The reference's own host code
(multiz.c, mz_preyama.c, maf.c, ...) runs unmodified around our `yama` symbol; the speculative record/replay
driver must reproduce the reference's output byte for byte whatever the aligner behind the C ABI is.  Here
that aligner is the oracle (integration/_ref/bin/multiz_shim -> oracle/libyama_shim.so): a test of the HOST
side only.  The same cases run against the CUDA library in tests/test_dropin_gpu.py."""
import os

import pytest

from dropin_util import (REF_MULTIZ, SHIM_MULTIZ, SHIM_SERVER, check_against_live_reference, check_golden_cases,
                         check_speculation, make_roast_dataset, run_roast, server_env, stop_server)

pytestmark = pytest.mark.skipif(not os.path.exists(SHIM_MULTIZ),
                                reason="integration/_ref/bin/multiz_shim not built (needs /root/reference at build time)")


@pytest.mark.parametrize("mode", ["defer", "batch", "direct", "stream"])
def test_golden_maf_cases(tmp_path, mode):
    check_golden_cases(SHIM_MULTIZ, tmp_path, env={"YB_DROPIN": mode})


def test_golden_maf_cases_with_block_scores_through_the_abi(tmp_path):
    """YB_SCORE=gpu routes the host's mafScoreRange (mz_scores.c:124) through yb_score_blocks in the real pass and
    skips it in the speculative passes; every `a score=` line must still match the reference's."""
    check_golden_cases(SHIM_MULTIZ, tmp_path, env={"YB_SCORE": "gpu"})


def test_golden_maf_cases_through_the_resident_server(tmp_path):
    """YB_SERVER: the drop-in ships its batches (and, with YB_SCORE=gpu, its block scores) to yama_b200d over a unix
    socket (integration/yb_wire.h) instead of owning a context; the server is started on demand, serves every
    invocation of the test, and the output bytes stay the reference's.  Host logic only: the server here is the same
    program linked against the oracle shim."""
    assert os.path.exists(SHIM_SERVER)
    env = server_env(tmp_path, SHIM_SERVER)
    try:
        check_golden_cases(SHIM_MULTIZ, tmp_path / "a", env=dict(env, YB_DROPIN="batch"))
        assert os.path.exists(env["YB_SERVER"])                     # one server, still there for the next invocation
        check_golden_cases(SHIM_MULTIZ, tmp_path / "b", env=dict(env, YB_SCORE="gpu"))
        check_golden_cases(SHIM_MULTIZ, tmp_path / "c", env=dict(env, YB_DROPIN="direct"))
        check_golden_cases(SHIM_MULTIZ, tmp_path / "d", env=env)        # the default behind a server: streamed replay
    finally:
        stop_server(env)


def test_resident_server_answers_concurrent_clients(tmp_path):
    """Several multiz invocations at once (parallel pipelines) share one yama_b200d: every client stays connected and
    the server answers their requests one at a time; each output must still be the reference's."""
    import shutil
    import subprocess
    from dropin_util import GOLD_MAF
    env = dict(os.environ, **server_env(tmp_path, SHIM_SERVER))
    try:
        procs = []
        for k in range(6):
            d = tmp_path / f"w{k}"
            d.mkdir()
            for f in ("ref.sp1.maf", "ref.sp2.maf"):
                shutil.copy(os.path.join(GOLD_MAF, f), d)
            v = str(k % 2)
            e = dict(env, YB_DROPIN="stream") if k >= 4 else env
            procs.append((v, subprocess.Popen([SHIM_MULTIZ, "ref.sp1.maf", "ref.sp2.maf", v, "o1", "o2"], cwd=d, env=e,
                                              stdout=subprocess.PIPE, stderr=subprocess.PIPE)))
        for v, p in procs:
            out, err = p.communicate(timeout=300)
            assert p.returncode == 0, err.decode()[-300:]
            want = open(os.path.join(GOLD_MAF, f"v{v}.stdout"), "rb").read()
            strip = lambda b: b"".join(l for l in b.splitlines(keepends=True) if not l.startswith(b"#"))
            assert strip(out) == strip(want)
    finally:
        stop_server(env)


def test_no_server_and_no_spawn_fails_loudly(tmp_path):
    import shutil
    from dropin_util import GOLD_MAF, run_tool
    for f in ("ref.sp1.maf", "ref.sp2.maf"):
        shutil.copy(os.path.join(GOLD_MAF, f), tmp_path)
    env = {"YB_SERVER": str(tmp_path / "nobody.sock"), "YB_SERVER_SPAWN": "0", "YB_SERVER_WAIT_S": "0.2"}
    rc, out, err = run_tool(SHIM_MULTIZ, ["ref.sp1.maf", "ref.sp2.maf", "1", "o1", "o2"], str(tmp_path), env)
    assert rc != 0 and b"no server answers on" in err


@pytest.mark.skipif(not os.path.exists(REF_MULTIZ), reason="oracle/_ref/bin/multiz not built")
def test_fresh_data_against_reference_binary(tmp_path):
    rep = check_against_live_reference(SHIM_MULTIZ, tmp_path / "defer", ref_len=60_000, n_species=4, seed=5,
                                       env={"YB_DROPIN_STATS": "1"})
    # deferred output: the host runs once for v=1, twice for v=0 (stage 2 consumes stage 1's output, mz_preyama.c:335)
    check_speculation(rep)
    rep = check_against_live_reference(SHIM_MULTIZ, tmp_path / "batch", ref_len=60_000, n_species=4, seed=5,
                                       env={"YB_DROPIN_STATS": "1", "YB_DROPIN": "batch"})
    check_speculation(rep, "batch")


@pytest.mark.skipif(not os.path.exists(REF_MULTIZ), reason="oracle/_ref/bin not built")
def test_roast_driver_runs_the_dropin(tmp_path):
    """The reference's roast (auto_mz.c) exec's `multiz` from PATH; with the drop-in there its progressive alignment
    of a 4-species tree (v=1 and v=0 merges, maf_project in between) is unchanged apart from the '#' provenance lines."""
    import shutil
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    make_roast_dataset(a, 40_000, 3, seed=8)
    shutil.copytree(a, b)
    tree = "((ref sp1) (sp2 sp3))"
    want = run_roast(REF_MULTIZ, a, tree)
    got = run_roast(SHIM_MULTIZ, b, tree)
    assert len(want) > 10_000 and got == want


def _drop_leading_blocks(src, dst, keep_from):
    """Copy a MAF, leaving out the blocks before block number `keep_from` (header and trailer kept)."""
    text = open(src, "rb").read()
    head, *blocks = text.split(b"\na ")
    tail = b""
    if blocks and b"##eof" in blocks[-1]:
        blocks[-1], tail = blocks[-1].split(b"##eof", 1)
        tail = b"##eof" + tail
    open(dst, "wb").write(head + b"".join(b"\na " + b for b in blocks[keep_from:]) + tail)
    return len(blocks)


@pytest.mark.skipif(not os.path.exists(REF_MULTIZ), reason="oracle/_ref/bin/multiz not built")
@pytest.mark.parametrize("v", [1, 0])
def test_speculative_passes_never_touch_the_output_files(tmp_path, v):
    """multiz opens out1/out2 with "w" (multiz.c:242-243).  A speculative pass that did the same would truncate what
    the real pass -- running beside it in streamed mode -- has already flushed: file1 here carries hundreds of KB of
    blocks before file2's first block, so out1 is far beyond one stdio buffer when the second speculative pass (v=0)
    starts.  Every mode must give the reference's bytes."""
    import shutil
    from tools.mafsynth import make_dataset
    from dropin_util import run_tool
    da = str(tmp_path / "ref")
    make_dataset(da, ref_len=400_000, n_species=2, seed=21)
    n = _drop_leading_blocks(os.path.join(da, "ref.sp2.maf"), os.path.join(da, "late.maf"), keep_from=250)
    assert n > 300
    argv = ["ref.sp1.maf", "late.maf", str(v), "o1", "o2"]
    rc, want, _ = run_tool(REF_MULTIZ, argv, da)
    assert rc == 0
    want1, want2 = (open(os.path.join(da, f), "rb").read() for f in ("o1", "o2"))
    assert len(want1) > 200_000
    for mode in ("defer", "stream", "batch", "direct"):
        db = str(tmp_path / mode)
        shutil.copytree(da, db)
        os.remove(os.path.join(db, "o1")); os.remove(os.path.join(db, "o2"))
        rc, out, err = run_tool(SHIM_MULTIZ, argv, db, {"YB_DROPIN": mode})
        assert rc == 0, err.decode()[-300:]
        assert out == want, mode
        assert open(os.path.join(db, "o1"), "rb").read() == want1, mode
        assert open(os.path.join(db, "o2"), "rb").read() == want2, mode


def test_inputs_that_cannot_be_read_twice(tmp_path):
    """The reference reads /dev/stdin (maf.c:343) and FIFOs; record/replay would run main() twice over them.  Such an
    invocation runs once, one pair per launch, and still gives the reference's bytes."""
    import shutil
    import subprocess
    from dropin_util import GOLD_MAF
    for f in ("ref.sp1.maf", "ref.sp2.maf"):
        shutil.copy(os.path.join(GOLD_MAF, f), tmp_path)
    want = open(os.path.join(GOLD_MAF, "v1.stdout"), "rb").read()
    strip = lambda b: b"".join(l for l in b.splitlines(keepends=True) if not l.startswith(b"#"))
    with open(tmp_path / "ref.sp1.maf", "rb") as f:
        p = subprocess.run([SHIM_MULTIZ, "/dev/stdin", "ref.sp2.maf", "1", "o1", "o2"], cwd=tmp_path, stdin=f,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert p.returncode == 0, p.stderr.decode()[-300:]
    assert strip(p.stdout) == strip(want)
    fifo = tmp_path / "in.fifo"
    os.mkfifo(fifo)
    feeder = subprocess.Popen(["sh", "-c", f"cat ref.sp2.maf > {fifo}"], cwd=tmp_path)
    p = subprocess.run([SHIM_MULTIZ, "R=30", "ref.sp1.maf", str(fifo), "1", "o1", "o2"], cwd=tmp_path,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    feeder.wait(timeout=60)
    assert p.returncode == 0, p.stderr.decode()[-300:]
    assert strip(p.stdout) == strip(want)


@pytest.mark.parametrize("mode", ["batch", "stream"])
def test_replay_hits_are_verified(tmp_path, mode):
    """A replay-table hit is trusted only if the job's dimensions and a second, independent digest match too.  With the
    key deliberately weakened to (K, L) every job collides with the first one of its depth: all but one become misses,
    aligned one at a time, and the output is still the reference's."""
    check_golden_cases(SHIM_MULTIZ, tmp_path, env={"YB_DROPIN": mode, "YB_DROPIN_WEAK_KEY": "1"})


@pytest.mark.skipif(not os.path.exists(REF_MULTIZ), reason="oracle/_ref/bin not built")
def test_tba_driver_runs_the_dropin(tmp_path):
    """BASELINE configs[3], host logic: the reference's tba (tba.c:114-276) over a 4-species tree -- maf_project,
    pair2tb, `multiz ... 1 out1 out2`, get_covered per cross-subtree pair -- with the drop-in as the multiz on PATH
    gives the reference's threaded blockset (apart from '#' provenance lines), also with E=ref."""
    import shutil
    from dropin_util import run_tba
    from tools.mafsynth import make_tba_dataset
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    files = make_tba_dataset(a, ["ref", "sp1", "sp2", "sp3"], ref_len=25_000, seed=12)
    shutil.copytree(a, b)
    tree = "((ref sp1) (sp2 sp3))"
    for extra in ((), ("E=ref",)):
        want = run_tba(REF_MULTIZ, a, tree, files, extra=extra)
        got = run_tba(SHIM_MULTIZ, b, tree, files, extra=extra)
        assert len(want) > 50_000 and got == want, extra
