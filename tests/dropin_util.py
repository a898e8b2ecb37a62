"""Helpers for the MAF-level parity tests: run a multiz-compatible tool and compare every output byte with
the reference's (committed golden outputs, or the reference binary in oracle/_ref/bin run beside it)."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from tools.make_golden import MAF_CASES  # noqa: E402  (the argv of every golden case)

GOLD_MAF = os.path.join(ROOT, "tests", "golden", "maf")
REF_MULTIZ = os.path.join(ROOT, "oracle", "_ref", "bin", "multiz")
GPU_MULTIZ = os.path.join(ROOT, "integration", "_ref", "bin", "multiz")
GPU_MULTIC = os.path.join(ROOT, "integration", "_ref", "bin", "multic")
SHIM_MULTIZ = os.path.join(ROOT, "integration", "_ref", "bin", "multiz_shim")
GPU_SERVER = os.path.join(ROOT, "integration", "_ref", "bin", "yama_b200d")
SHIM_SERVER = os.path.join(ROOT, "integration", "_ref", "bin", "yama_b200d_shim")


def server_env(tmp_path, server_bin, idle_s=20):
    """Environment that sends the drop-in's batches to a resident server on a private socket (started on demand)."""
    return {"YB_SERVER": str(tmp_path / "yb.sock"), "YB_SERVER_BIN": server_bin, "YB_SERVER_IDLE_S": str(idle_s),
            "YB_DROPIN_STATS": "1"}


def stop_server(env):
    """Send SIGTERM to the server this test started (found by its unique --socket argument in /proc) instead of
    letting it idle out."""
    import signal
    sock = env["YB_SERVER"].encode()
    for pid in os.listdir("/proc"):
        if not pid.isdigit():
            continue
        try:
            argv = open(f"/proc/{pid}/cmdline", "rb").read().split(b"\0")
        except OSError:
            continue
        if len(argv) >= 3 and argv[0].endswith(b"yama_b200d") and sock in argv:
            try:
                os.kill(int(pid), signal.SIGTERM)
            except OSError:
                pass


def run_tool(tool, argv, cwd, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([tool] + list(argv), cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e, timeout=timeout)
    return p.returncode, p.stdout, p.stderr


def check_golden_cases(tool, tmp_path, env=None):
    """Every case of tools/make_golden.py:MAF_CASES, byte for byte (stdout + out1 + out2)."""
    d = str(tmp_path / "maf")
    os.makedirs(d, exist_ok=True)
    for f in ("ref.sp1.maf", "ref.sp2.maf", "ref.sp3.maf", "v1.stdout"):
        shutil.copy(os.path.join(GOLD_MAF, f), d)
    for name, argv in MAF_CASES:
        outs = [a for a in argv if a.startswith(name + ".out")]
        # the tool writes NAME.out1/2 into the scratch dir; the expected bytes are the golden files of that name
        rc, out, err = run_tool(tool, argv, d, env)
        assert rc == 0, (name, err.decode()[-500:])
        want = open(os.path.join(GOLD_MAF, name + ".stdout"), "rb").read()
        if name == "step2":
            pass    # its input v1.stdout was copied from the golden set above
        assert out == want, f"{name}: stdout differs from the reference's"
        for o in outs:
            assert open(os.path.join(d, o), "rb").read() == open(os.path.join(GOLD_MAF, o), "rb").read(), (name, o)


def check_against_live_reference(tool, tmp_path, ref_len, n_species, seed, env=None, versions=(1, 0), extra=()):
    """Fresh synthetic data; the reference binary and `tool` run with identical argv in sibling dirs."""
    from tools.mafsynth import make_dataset
    da, db = str(tmp_path / "ref"), str(tmp_path / "our")
    make_dataset(da, ref_len=ref_len, n_species=n_species, seed=seed, lower=0.01)
    shutil.copytree(da, db)
    report = []
    for v in versions:
        argv = list(extra) + ["ref.sp1.maf", "ref.sp2.maf", str(v), f"u1.v{v}", f"u2.v{v}"]
        rc_r, out_r, _ = run_tool(REF_MULTIZ, argv, da)
        rc_o, out_o, err_o = run_tool(tool, argv, db, env)
        assert rc_r == 0 and rc_o == 0, err_o.decode()[-500:]
        assert out_o == out_r, f"v={v}: stdout differs"
        for f in (f"u1.v{v}", f"u2.v{v}"):
            assert open(os.path.join(db, f), "rb").read() == open(os.path.join(da, f), "rb").read(), f
        report.append((v, len(out_r), err_o.decode().strip().splitlines()[-1:] if err_o else []))
        if v == 1 and n_species >= 3:          # progressive merge: acc = multiz(acc, ref.spI, 1)
            acc = "acc2.maf"
            for d, o in ((da, out_r), (db, out_o)):
                open(os.path.join(d, acc), "wb").write(o)
            for i in range(3, n_species + 1):
                argv = list(extra) + [acc, f"ref.sp{i}.maf", "1", "w1", "w2"]
                rc_r, out_r, _ = run_tool(REF_MULTIZ, argv, da)
                rc_o, out_o, err_o = run_tool(tool, argv, db, env)
                assert rc_r == 0 and rc_o == 0, err_o.decode()[-500:]
                assert out_o == out_r, f"progressive step {i}: stdout differs"
                acc = f"acc{i}.maf"
                for d, o in ((da, out_r), (db, out_o)):
                    open(os.path.join(d, acc), "wb").write(o)
    return report


def parse_stats(line):
    """'yama_b200: passes=2 batches=2 jobs=...' -> dict of numbers."""
    import re
    return {k: float(v) for k, v in re.findall(r"(\w+)=([0-9.]+)", line)}


def check_speculation(report, mode="defer"):
    """How many speculative passes of the host the drop-in needed (its stats line).  Deferred mode (the default): v=1
    none -- the host runs once, its merged blocks are written when the batch is back; v=0 one, for the first yama() of
    every overlap (stage 2 consumes stage 1's output, mz_preyama.c:335).  Batch mode: one and two.  A handful of v=0
    second-stage calls may miss in batch mode because the REFERENCE's band for them depends on an uninitialised heap
    byte (mz_preyama.c:296 passes rows 1..K of a K-row A to mapping()); the drop-in then aligns that pair
    synchronously, so the output stays exact; see DESIGN.md."""
    for v, _, last in report:
        assert last, "no stats line"
        st = parse_stats(last[0])
        want = {("defer", 1): 0, ("defer", 0): 1, ("batch", 1): 1, ("batch", 0): 2}[(mode, v)]
        assert st["passes"] == want, last
        assert st["failed"] == 0
        if v == 1 or mode == "defer":
            assert st["misses"] == 0, last
        else:
            assert st["misses"] <= max(2, 0.01 * st["calls"]), last


def run_roast(multiz_tool, workdir, tree, env=None):
    """The reference's own roast driver (oracle/_ref/bin/roast, which exec's `multiz` and `maf_project` from PATH,
    auto_mz.c:19-20) with `multiz_tool` first on PATH.  Returns the output MAF without '#' lines (they embed getpid())."""
    import shutil as _sh
    refbin = os.path.dirname(REF_MULTIZ)
    bindir = os.path.join(workdir, "_bin")
    os.makedirs(bindir, exist_ok=True)
    for name, src in (("multiz", multiz_tool), ("maf_project", os.path.join(refbin, "maf_project")),
                      ("roast", os.path.join(refbin, "roast"))):
        dst = os.path.join(bindir, name)
        if os.path.lexists(dst):
            os.remove(dst)
        os.symlink(src, dst)
    e = dict(os.environ)
    e.update(env or {})
    e["PATH"] = bindir + os.pathsep + e.get("PATH", "")
    files = sorted(f for f in os.listdir(workdir) if f.endswith(".sing.maf"))
    out = os.path.join(workdir, "roast_out.maf")
    if os.path.exists(out):
        os.remove(out)
    p = subprocess.run([os.path.join(bindir, "roast"), f"T={workdir}", "E=ref", tree] + files + [out], cwd=workdir, env=e,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
    assert p.returncode == 0, p.stderr.decode()[-500:]
    _ = _sh
    return b"".join(l for l in open(out, "rb").read().splitlines(keepends=True) if not l.startswith(b"#"))


def make_roast_dataset(workdir, ref_len, n_species, seed):
    from tools.mafsynth import make_dataset
    os.makedirs(workdir, exist_ok=True)
    for p in make_dataset(workdir, ref_len=ref_len, n_species=n_species, seed=seed):
        sp = os.path.basename(p).split(".")[1]
        os.replace(p, os.path.join(workdir, f"ref.{sp}.sing.maf"))


def run_tba(multiz_tool, workdir, tree, files, env=None, extra=()):
    """The reference's own tba driver (oracle/_ref/bin/tba, tba.c:114-276: per cross-subtree pair it shells out to
    maf_project, pair2tb, `multiz ... 1 ...`, get_covered, all found on PATH) with `multiz_tool` as the multiz on PATH.
    Returns the output MAF without '#' lines (they embed getpid() through the temp-file names, tba.c:299-302)."""
    refbin = os.path.dirname(REF_MULTIZ)
    bindir = os.path.join(workdir, "_bin")
    os.makedirs(bindir, exist_ok=True)
    for name, src in (("multiz", multiz_tool), ("maf_project", os.path.join(refbin, "maf_project")),
                      ("pair2tb", os.path.join(refbin, "pair2tb")), ("get_covered", os.path.join(refbin, "get_covered")),
                      ("tba", os.path.join(refbin, "tba"))):
        dst = os.path.join(bindir, name)
        if os.path.lexists(dst):
            os.remove(dst)
        os.symlink(src, dst)
    e = dict(os.environ)
    e.update(env or {})
    e["PATH"] = bindir + os.pathsep + e.get("PATH", "")
    out = os.path.join(workdir, "tba_out.maf")
    if os.path.exists(out):
        os.remove(out)
    p = subprocess.run([os.path.join(bindir, "tba")] + list(extra) + [tree] + list(files) + [out], cwd=workdir, env=e,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=1800)
    assert p.returncode == 0, p.stderr.decode()[-800:]
    return b"".join(l for l in open(out, "rb").read().splitlines(keepends=True) if not l.startswith(b"#"))
