"""GPU parity at the MAF level (BASELINE.json configs[0]): the reference's own multiz host linked against
libyama_b200.so (integration/_ref/bin/multiz) must write byte-identical MAF to the reference binary --
stdout, out1 and out2 -- in v=1 and v=0, batched (record/replay) and one pair per launch."""
import os

import pytest

from dropin_util import (GPU_MULTIC, GPU_MULTIZ, GPU_SERVER, REF_MULTIZ, check_against_live_reference, check_golden_cases,
                         check_speculation, make_roast_dataset, run_roast, run_tool, server_env, stop_server)

pytestmark = [pytest.mark.gpu]


def _need(path):
    assert os.path.exists(path), f"{path} missing: run __graft_entry__.build() where /root/reference exists"


@pytest.mark.parametrize("mode", ["defer", "batch", "direct", "stream"])
def test_golden_maf_cases_gpu(tmp_path, mode):
    _need(GPU_MULTIZ)
    check_golden_cases(GPU_MULTIZ, tmp_path, env={"YB_DROPIN": mode})


def test_block_scores_on_the_device(tmp_path):
    """YB_SCORE=gpu: the host's mafScoreRange calls (mz_preyama.c:79, multi_util.c:509..802) answered by
    yb_score_kernel -- the `a score=` lines of every golden case and of a fresh progressive merge stay identical."""
    _need(GPU_MULTIZ); _need(REF_MULTIZ)
    check_golden_cases(GPU_MULTIZ, tmp_path / "golden", env={"YB_SCORE": "gpu"})
    rep = check_against_live_reference(GPU_MULTIZ, tmp_path / "fresh", ref_len=100_000, n_species=4, seed=12,
                                       env={"YB_SCORE": "gpu", "YB_DROPIN_STATS": "1"})
    stats = [last[0] for _, _, last in rep if last]
    assert stats and all("score_calls=" in s and "score_calls=0 " not in s for s in stats), stats


def test_resident_server_backend(tmp_path):
    """YB_SERVER: yama_b200d owns the CUDA context; every multiz invocation of the test (golden cases, a progressive
    4-way merge with v=1 and v=0, block scores on the device) goes through it and stays byte-identical."""
    _need(GPU_MULTIZ); _need(GPU_SERVER); _need(REF_MULTIZ)
    env = server_env(tmp_path, GPU_SERVER, idle_s=60)
    try:
        check_golden_cases(GPU_MULTIZ, tmp_path / "golden", env=env)
        rep = check_against_live_reference(GPU_MULTIZ, tmp_path / "fresh", ref_len=150_000, n_species=4, seed=21,
                                           env=dict(env, YB_SCORE="gpu"))
        check_speculation(rep)
        assert all("create_ms=0 " in last[0] for _, _, last in rep if last)      # no CUDA start-up in the tool itself
        check_golden_cases(GPU_MULTIZ, tmp_path / "stream", env=dict(env, YB_DROPIN="stream"))   # streamed replay behind the server
    finally:
        stop_server(env)


def test_cfg1_one_megabase_merge(tmp_path):
    """configs[0]: multiz merge of two synthetic pairwise MAFs on a 1 Mb reference, R=30 M=1, v=1 and v=0."""
    _need(GPU_MULTIZ); _need(REF_MULTIZ)
    rep = check_against_live_reference(GPU_MULTIZ, tmp_path, ref_len=1_000_000, n_species=2, seed=1,
                                       env={"YB_DROPIN_STATS": "1"})
    check_speculation(rep)
    for v, _, last in rep:
        print("cfg1 v=%d:" % v, last[0])


def test_progressive_five_way_small(tmp_path):
    """configs[1] in miniature: progressive 5-way merge (K grows 2..5), 200 kb, R=30 and R=100."""
    _need(GPU_MULTIZ); _need(REF_MULTIZ)
    check_against_live_reference(GPU_MULTIZ, tmp_path / "r30", ref_len=200_000, n_species=5, seed=2, versions=(1,))
    check_against_live_reference(GPU_MULTIZ, tmp_path / "r100", ref_len=100_000, n_species=3, seed=3, versions=(1, 0),
                                 extra=("R=100",))


def test_multic_same_boundary(tmp_path):
    """multic reaches yama through the same pre_yama() (multic.c:72); its drop-in build must agree with the
    reference's multic."""
    _need(GPU_MULTIC)
    ref_multic = os.path.join(os.path.dirname(REF_MULTIZ), "multic")
    _need(ref_multic)
    from tools.mafsynth import make_dataset
    d = str(tmp_path / "d")
    make_dataset(d, ref_len=80_000, n_species=2, seed=9)
    argv = ["ref.sp1.maf", "ref.sp2.maf", "1"]
    rc_r, out_r, _ = run_tool(ref_multic, argv, d)
    rc_o, out_o, err = run_tool(GPU_MULTIC, argv, d)
    assert rc_r == rc_o, err.decode()[-300:]
    assert out_o == out_r


def test_roast_five_species_tree(tmp_path):
    """configs[3] in miniature: the reference's roast driver over a 5-species tree, 300 kb, with the GPU multiz on PATH
    (all visible GPUs); identical to the run with the reference's multiz apart from '#' lines."""
    import shutil
    _need(GPU_MULTIZ); _need(REF_MULTIZ)
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    make_roast_dataset(a, 300_000, 4, seed=12)
    shutil.copytree(a, b)
    tree = "((ref sp1) ((sp2 sp3) sp4))"
    want = run_roast(REF_MULTIZ, a, tree)
    got = run_roast(GPU_MULTIZ, b, tree)
    assert len(want) > 100_000 and got == want


def test_tba_eight_species_tree(tmp_path):
    """configs[3] (tba arm): the reference's tba driver (tba.c:114-276) over an 8-species tree on a 300 kb ancestor --
    28 pairwise files, FASTA per species, maf_project / pair2tb / get_covered from the reference and the GPU multiz on
    PATH (always v=1, with out1/out2 files).  The threaded blockset must equal the one the reference's multiz gives,
    apart from '#' lines (they carry the temp-file names, which embed getpid())."""
    import shutil
    from dropin_util import run_tba
    from tools.mafsynth import make_tba_dataset
    _need(GPU_MULTIZ); _need(REF_MULTIZ)
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    files = make_tba_dataset(a, ["ref"] + [f"sp{i}" for i in range(1, 8)], ref_len=300_000, seed=3)
    shutil.copytree(a, b)
    tree = "(((ref sp1) (sp2 sp3)) ((sp4 sp5) (sp6 sp7)))"
    want = run_tba(REF_MULTIZ, a, tree, files)
    got = run_tba(GPU_MULTIZ, b, tree, files)
    assert len(want) > 3_000_000 and got == want
    srv = server_env(tmp_path, GPU_SERVER, idle_s=60)
    try:
        got2 = run_tba(GPU_MULTIZ, b, tree, files, env=srv)       # the same pipeline behind the resident server
    finally:
        stop_server(srv)
    assert got2 == want


def test_kernel_limits_stop_the_tool_like_a_failing_yama(tmp_path):
    """Limits the reference does not have: more than 255 rows in a profile, a band row wider than the
    widest ring.  There is no CPU path to fall back to, so the tool must stop the way the reference stops when yama()
    fails -- a message on stderr, a non-zero exit code, everything BEFORE the failing pair written, nothing of it or
    after it -- instead of writing a wrong or partial merge silently."""
    import random
    _need(GPU_MULTIZ); _need(REF_MULTIZ)
    rng = random.Random(3)
    n1, n2 = 400, 300
    seq = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
    ref = seq(n1 + 50 + n2)
    def block(start, n, rows):
        t = ref[start:start + n]
        lines = ["a score=0.0", "s ref.chr1 %d %d + %d %s" % (start, n, len(ref), t)]
        for k in range(rows):
            lines.append("s sp%d.chr1 %d %d + %d %s" % (k, 10, n, 5000, t))
        return "\n".join(lines) + "\n\n"
    head = "##maf version=1 scoring=multiz\n"
    f1 = head + block(0, n1, 1) + block(n1 + 50, n2, 299)          # second block: K = 300 rows
    f2 = head + "a score=0.0\ns ref.chr1 0 %d + %d %s\ns other.chr1 0 %d + %d %s\n\n" % (len(ref), len(ref), ref, len(ref), len(ref), ref)
    outs = {}
    for name, tool in (("ref", REF_MULTIZ), ("gpu", GPU_MULTIZ)):
        d = tmp_path / name
        d.mkdir()
        (d / "a.maf").write_text(f1); (d / "b.maf").write_text(f2)
        outs[name] = run_tool(tool, ["a.maf", "b.maf", "1", "o1", "o2"], str(d))
    rc_r, out_r, _ = outs["ref"]
    rc_g, out_g, err_g = outs["gpu"]
    assert rc_r == 0                                              # the reference has no such limit
    assert rc_g != 0 and b"exceeds the kernel limit" in err_g, err_g[-300:]
    import re
    blocks_r = re.findall(rb"^a .*?\n\n", out_r, flags=re.S | re.M)
    blocks_g = re.findall(rb"^a .*?\n\n", out_g, flags=re.S | re.M)
    assert len(blocks_r) >= 2 and blocks_g == blocks_r[:len(blocks_g)] and len(blocks_g) >= 1      # a true prefix
    assert not any(b.count(b"\ns ") > 10 for b in blocks_g)      # nothing of the 300-row merge
