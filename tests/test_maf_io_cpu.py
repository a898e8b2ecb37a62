"""CPU suite: the drop-in's MAF reader / writer (integration/maf_dropin.c -- the `mafNext`, `mafReadAll`, `mafWrite` symbols,
SURVEY 8(f) rank 4) against the reference's own (maf.c:133-216, :251-294).  The reference binary and the drop-in host run
the same command on MAF files written the way real files are NOT always written: tabs and runs of blanks between fields,
comment lines between and inside blocks, `i`/`e`/`q` lines, signed and zero-padded numbers, amplifier / copy flags, CR-LF
line ends, text followed by blanks, a last line without a newline, source names with and without a contig.  stdout, out1
and out2 must agree byte for byte -- blocks that overlap nothing are copied to out1 / out2 by mafWrite, merged ones are
re-formatted -- and a malformed file must stop both tools with the same message and exit code."""
import os
import random
import re
import shutil

import pytest

from dropin_util import REF_MULTIZ, SHIM_MULTIZ, run_tool

pytestmark = pytest.mark.skipif(not (os.path.exists(SHIM_MULTIZ) and os.path.exists(REF_MULTIZ)),
                                reason="integration/_ref/bin/multiz_shim or the reference binary not built")


def blocks_of(path):
    """[(a-line, [s-lines])] of a MAF file written by tools/mafsynth.py."""
    out, cur = [], None
    for ln in open(path).read().splitlines():
        if ln.startswith("a"):
            cur = (ln, [])
            out.append(cur)
        elif ln.startswith("s") and cur is not None:
            cur[1].append(ln)
    return out


def restyle(blocks, rng, eol="\n", tail=True):
    """The same blocks, written oddly (every variation is one the reference's sscanf-based reader accepts)."""
    txt = ["##maf version=1 scoring=multiz" + eol, "# a comment the reader echoes" + eol, "#" + eol]
    for k, (a, rows) in enumerate(blocks):
        style = k % 9
        if style == 1:
            txt.append("# comment between blocks, mentions eof so it is not echoed" + eol)
        if style == 2:
            a = "a score=%d.5" % (k * 7)
        if style == 3 and len(rows) >= 2:
            a = a + " amplifier=1"
        if style == 4 and len(rows) >= 2:
            a = "a copy=1 score=12.0"
        txt.append(a + eol)
        for j, s in enumerate(rows):
            f = s.split()
            sep = ["\t", "  ", " \t ", "    "][(k + j) % 4] if style in (5, 6) else " "
            if style == 6:
                f[2] = "+" + f[2]                           # %d takes a sign
                f[5] = "0" + f[5]                           # ... and leading zeros
            if style == 7:
                f[1] = f[1].split(".")[0] if j == 0 else f[1] + "x.y"     # no contig / a second dot
            line = sep.join(f)
            if style == 8:
                line += "   "                               # %s stops at the blank
            txt.append(line + eol)
            if style == 5 and j == 0:
                txt.append("i %s N 0 C 0" % f[1] + eol)     # not an `s` line: skipped (maf.c:166-167)
            if style == 1 and j == 0:
                txt.append("# a comment inside a block ends it for the reader? no: it is skipped" + eol)
        txt.append(eol)
    s = "".join(txt)
    if not tail:
        s = s.rstrip("\r\n")
    return s


def norm_err(b):
    return re.sub(rb"^[^:\n]*multiz[a-z_]*: ", b"multiz: ", b, flags=re.M)


def both(tmp_path, name, files, argv):
    outs = []
    for tool, sub, env in ((REF_MULTIZ, "ref", None), (SHIM_MULTIZ, "our", None), (SHIM_MULTIZ, "our_direct", {"YB_DROPIN": "direct"})):
        d = tmp_path / f"{name}_{sub}"
        d.mkdir()
        for fn, content in files.items():
            (d / fn).write_text(content, newline="")
        rc, out, err = run_tool(tool, argv, str(d), env)
        extra = {fn: (d / fn).read_bytes() for fn in ("o1", "o2") if (d / fn).exists()}
        outs.append((rc, out, norm_err(err) if rc else b"", extra))
    assert outs[0][0] == outs[1][0] == outs[2][0], (name, [o[0] for o in outs], outs[1][2][-300:])
    assert outs[0][1] == outs[1][1] and outs[0][1] == outs[2][1], f"{name}: stdout differs"
    assert outs[0][3] == outs[1][3] and outs[0][3] == outs[2][3], f"{name}: out1/out2 differ"
    assert outs[0][2] == outs[1][2] == outs[2][2], (name, outs[0][2], outs[1][2])
    return outs[0]


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    from tools.mafsynth import make_dataset
    d = tmp_path_factory.mktemp("mafio")
    make_dataset(str(d), ref_len=60_000, n_species=2, seed=11, lower=0.02)
    return blocks_of(str(d / "ref.sp1.maf")), blocks_of(str(d / "ref.sp2.maf"))


@pytest.mark.parametrize("eol,tail", [("\n", True), ("\r\n", True), ("\n", False)])
def test_odd_but_valid_files_read_and_written_like_the_reference(tmp_path, dataset, eol, tail):
    b1, b2 = dataset
    rng = random.Random(5)
    # drop some blocks of each file so that pieces of the other overlap nothing and are copied through mafWrite
    f1 = restyle([b for i, b in enumerate(b1) if i % 5 != 2], rng, eol, tail)
    f2 = restyle([b for i, b in enumerate(b2) if i % 4 != 1], rng, eol, tail)
    rc, out, _, extra = both(tmp_path, "odd", {"a.maf": f1, "b.maf": f2}, ["a.maf", "b.maf", "1", "o1", "o2"])
    if eol == "\r\n":
        assert rc != 0          # the reference does not take CR-LF files ("\r\n" does not end a block): the same death
        return
    assert rc == 0 and out.count(b"\na ") > 10 and extra["o1"].count(b"\ns ") > 0
    rc, out, _, _ = both(tmp_path, "odd_v0_stdout", {"a.maf": f1, "b.maf": f2}, ["a.maf", "b.maf", "0"])
    assert rc == 0


def test_long_rows_cross_the_read_buffer(tmp_path):
    """Rows longer than the reader's 1 MiB piece and a file of many pieces."""
    rng = random.Random(9)
    n = 1_300_000
    seq = "".join(rng.choice("ACGT") for _ in range(n))
    other = "".join(c if rng.random() > 0.1 else rng.choice("ACGT-") for c in seq)
    size2 = sum(ch != "-" for ch in other)
    f1 = "##maf version=1 scoring=x\na score=1.0\ns ref.chr1 0 %d + %d %s\ns sp1.chr1 0 %d + %d %s\n\n" % (n, n, seq, size2, size2, other)
    f2 = "##maf version=1 scoring=x\na score=1.0\ns ref.chr1 5 40 + %d %s\ns sp2.chr1 0 40 + 40 %s\n\n" % (n, seq[5:45], seq[5:45])
    # (the overlap is 40 columns; everything else of the long block goes to out1 through mafWrite)
    rc, out, _, extra = both(tmp_path, "long", {"a.maf": f1, "b.maf": f2}, ["a.maf", "b.maf", "1", "o1", "o2"])
    assert rc == 0 and len(extra["o1"]) > 2 * n


BAD = {
    "missing_field": "a score=0\ns ref.chr1 0 4 + ACGT\n\n",
    "size_mismatch": "a score=0\ns ref.chr1 0 5 + 100 ACGT\n\n",
    "bad_coords": "a score=0\ns ref.chr1 98 4 + 100 ACGT\n\n",
    "bad_coords_second_row": "a score=0\ns ref.chr1 0 4 + 100 ACGT\ns sp.chr1 98 4 + 100 ACGT\n\n",
    "negative_start": "a score=0\ns ref.chr1 -3 4 + 100 ACGT\n\n",
    "zero_size": "a score=0\ns ref.chr1 0 0 + 100 ----\n\n",
    "ragged_rows": "a score=0\ns ref.chr1 0 4 + 100 ACGT\ns sp.chr1 0 4 + 100 ACGT-\n\n",
    "no_a_line": "s ref.chr1 0 4 + 100 ACGT\n\n",
    "huge_number": "a score=0\ns ref.chr1 0 4 + 99999999999 ACGT\n\n",
    "empty_s": "a score=0\ns\n\n",
}


@pytest.mark.parametrize("case", sorted(BAD))
def test_malformed_files_die_with_the_reference_message(tmp_path, case):
    good = "##maf version=1 scoring=x\na score=0\ns ref.chr1 0 4 + 100 ACGT\ns sp2.chr1 0 4 + 50 ACGT\n\n"
    bad = "##maf version=1 scoring=x\n" + BAD[case]
    rc, _, err, _ = both(tmp_path, case, {"a.maf": bad, "b.maf": good}, ["a.maf", "b.maf", "1", "o1", "o2"])
    if case not in ("huge_number",):          # (the reference accepts what %d makes of an overflowing number)
        assert rc != 0 and err


def test_reference_reader_stays_reachable(tmp_path, dataset):
    """YB_MAF=ref routes mafNext / mafWrite back to the reference's own definitions (kept in maf.o under other names)."""
    b1, b2 = dataset
    rng = random.Random(1)
    files = {"a.maf": restyle(b1, rng), "b.maf": restyle(b2, rng)}
    d = tmp_path / "refio"
    d.mkdir()
    for fn, content in files.items():
        (d / fn).write_text(content, newline="")
    rc_a, out_a, _ = run_tool(SHIM_MULTIZ, ["a.maf", "b.maf", "1"], str(d), {"YB_MAF": "ref"})
    rc_b, out_b, _ = run_tool(SHIM_MULTIZ, ["a.maf", "b.maf", "1"], str(d))
    assert rc_a == 0 and rc_b == 0 and out_a == out_b
    _ = shutil
