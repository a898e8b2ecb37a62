"""CPU suite: the count-vector closed form that the CUDA kernels evaluate written
out in plain Python and compared with the oracle -- final C/D/I and EVERY traceback byte.  The oracle keeps the
reference's O(K*L) character loops (mz_yama.c:113-242); this model replaces each of them by the dot product of
per-row and per-column counts, with the reference's existence guards and its in-place dp[] semantics.  It documents,
runnable without a GPU, why the kernels' arithmetic is the reference's."""
import numpy as np
import pytest

from tools.synth import random_problem

MININT = -(2 ** 30)          # INT_MIN / 2, mz_yama.c:29
FLAG_C, FLAG_I, FLAG_D = 0, 1, 2


def _wrap(x):                 # int32 two's-complement, like the reference built with -fwrapv and the kernels
    return (int(x) + 2 ** 31) % 2 ** 32 - 2 ** 31


def _pick(x, y, z):           # mz_yama.c:138-154
    if x >= y and x >= z:
        return x, FLAG_C
    if y > z:
        return y, FLAG_D
    return z, FLAG_I


def _cls(col):                # six classes: A C G T other dash (mz_scores.c:39-54)
    n = [0] * 6
    for ch in col:
        ch = int(ch)
        if ch == 45:
            n[5] += 1
        else:
            n[{97: 0, 99: 1, 103: 2, 116: 3}.get(ch | 0x20, 4)] += 1
    return n


def _trans(prev, cur):        # counts of (previous is dash, this is dash); no previous column/row = all non-dash
    t = [[0, 0], [0, 0]]
    for k in range(len(cur)):
        p = 0 if prev is None else int(prev[k] == 45)
        t[p][int(cur[k] == 45)] += 1
    return t


def count_form_yama(A, B, LB, RB, S6, GO, GE, gates=True):
    """A [M,K], B [N,L] -> (C, D, I at (M,N), traceback bytes in the reference's row-major band order).
    gates=False drops the reference's existence guards on candidates (the fill kernels' GATED=false variant): a candidate
    from a node that does not exist -- exactly MININT -- is charged like any other."""
    M, K = A.shape
    N, L = B.shape
    cA = [None] + [_cls(A[r - 1]) for r in range(1, M + 1)]
    cB = [None] + [_cls(B[c - 1]) for c in range(1, N + 1)]
    a = [None] + [_trans(A[r - 2] if r > 1 else None, A[r - 1]) for r in range(1, M + 1)]      # a[r][s][u]
    b = [None] + [_trans(B[c - 2] if c > 1 else None, B[c - 1]) for c in range(1, N + 1)]      # b[c][t][v]
    w = [None] + [[sum(cA[r][k] * S6[k][l] for k in range(6)) for l in range(6)] for r in range(1, M + 1)]
    dpC, dpD, dpI = [MININT] * (N + 1), [MININT] * (N + 1), [MININT] * (N + 1)
    dpC[0] = dpD[0] = dpI[0] = 0                                                    # mz_yama.c:82
    tb = [0]
    for c in range(1, RB[0] + 1):                                                   # :84-91
        dpI[c] = _wrap(dpI[c - 1] - (L - cB[c][5]) * K * GE)
        tb.append(FLAG_I << 4)
    for r in range(1, M + 1):
        dA = cA[r][5]; ndA = K - dA
        a00, a01, a10, a11 = a[r][0][0], a[r][0][1], a[r][1][0], a[r][1][1]
        c = LB[r] - 1
        if LB[r - 1] <= c:                                                          # :100-106
            gc, gd, gi = dpC[c], dpD[c], dpI[c]
        else:
            gc = gd = gi = MININT
        C = D = I = MININT
        for c in range(LB[r], RB[r] + 1):
            # ---- I node (:113-166): from (r, c-1), the values of this row's previous column
            if c > LB[r]:
                dB = cB[c][5]; ndB = L - dB; b10 = b[c][1][0]
                x, y, z = C, D, I
                if r < M:
                    if not gates or c > LB[r - 1] + 1:
                        x -= GO * (ndA * ndB + dA * b10)
                    y -= GO * K * ndB
                    if not gates or c > LB[r] + 1:
                        z -= GO * K * b10
                nI, fi = _pick(_wrap(x), _wrap(y), _wrap(z))
                nI = _wrap(nI - ndB * K * GE)
            else:
                nI, fi = MININT, 0
            # ---- C node (:168-205): from the diagonal
            if c > LB[r - 1]:
                x, y, z = gc, gd, gi
                if c > 1:
                    dB = cB[c][5]; ndB = L - dB; b01, b10 = b[c][0][1], b[c][1][0]
                    if r > 1 and (not gates or c > LB[r - 2] + 1):
                        x -= GO * (a00 * b01 + a01 * ndB + a10 * dB + a11 * b10)
                    if r > 1:
                        y -= GO * (dA * ndB + a10 * dB)
                    if not gates or c > LB[r - 1] + 1:
                        z -= GO * (ndA * dB + dA * b10)
                nC, fc = _pick(_wrap(x), _wrap(y), _wrap(z))
                nC = _wrap(nC + sum(w[r][l] * cB[c][l] for l in range(6)))
            else:
                nC, fc = MININT, 0
            # ---- D node (:207-242): from (r-1, c), still in dp[]
            x, y, z = dpC[c], dpD[c], dpI[c]
            if 0 < c < N:
                dB = cB[c][5]; ndB = L - dB
                if r > 1 and (not gates or c > LB[r - 2]):
                    x -= GO * (ndA * ndB + a10 * dB)
                if r > 1:
                    y -= GO * L * a10
                if not gates or c > LB[r - 1]:
                    z -= GO * L * ndA
            nD, fd = _pick(_wrap(x), _wrap(y), _wrap(z))
            nD = _wrap(nD - ndA * L * GE)
            gc, gd, gi = dpC[c], dpD[c], dpI[c]                                     # :245-250
            C, D, I = nC, nD, nI
            dpC[c], dpD[c], dpI[c] = C, D, I
            tb.append(fc | (fd << 2) | (fi << 4))                                   # :253
    return (dpC[N], dpD[N], dpI[N]), np.asarray(tb, dtype=np.uint8)


@pytest.mark.parametrize("band", ["smooth", "full", "ragged"])
def test_count_vector_form_reproduces_every_traceback_byte(oracle, band):
    reps = b"ACGTN-"
    S6 = [[int(oracle.ss[x, y]) for y in reps] for x in reps]
    GO, GE = int(oracle.gop[1]), int(oracle.gap_ext)
    rng = np.random.default_rng({"smooth": 1, "full": 2, "ragged": 3}[band])
    for it in range(60):
        K, L = int(rng.integers(1, 7)), int(rng.integers(1, 7))
        M, N = int(rng.integers(1, 36)), int(rng.integers(1, 36))
        A, B, LB, RB = random_problem(rng, K, L, M, N, band=band, alphabet=("acgt", "mixed", "weird")[it % 3])
        want = oracle.yama(A, B, LB, RB, want_tback=True)
        cdi, tb = count_form_yama(np.asarray(A), np.asarray(B), [int(v) for v in LB], [int(v) for v in RB], S6, GO, GE)
        assert cdi == tuple(int(v) for v in want["cdi"]), (band, it, K, L, M, N)
        assert np.array_equal(tb, want["tback"]), (band, it, K, L, M, N)


def _script(tb, cdi, M, N, LB, RB):
    """mz_yama.c:257-291 on the row-major band-compact traceback bytes."""
    start = [0]
    for r in range(M + 1):
        start.append(start[-1] + RB[r] - LB[r] + 1)
    C, D, I = cdi
    node = FLAG_C if (C >= D and C >= I) else (FLAG_D if D >= I else FLAG_I)
    r, c, out = M, N, []
    while r > 0 or c > 0:
        assert r >= 0 and LB[r] <= c <= RB[r]
        st = int(tb[start[r] + c - LB[r]])
        out.append(node)
        if node == FLAG_I:
            c -= 1; node = st >> 4
        elif node == FLAG_D:
            r -= 1; node = (st >> 2) & 3
        else:
            r -= 1; c -= 1; node = st & 3
    return out


def test_dropping_the_existence_guards_changes_no_script(oracle):
    """The argument behind the fill kernels' GATED=false variant, executable: with connected bands
    and scores far from 2^28, charging candidates from non-existent nodes can only change traceback bytes of unreachable
    nodes -- the final scores and the edit script stay the reference's."""
    reps = b"ACGTN-"
    S6 = [[int(oracle.ss[x, y]) for y in reps] for x in reps]
    GO, GE = int(oracle.gop[1]), int(oracle.gap_ext)
    rng = np.random.default_rng(2718)
    changed_bytes = 0
    for it in range(240):
        band = ("smooth", "full", "ragged")[it % 3]
        K, L = int(rng.integers(1, 7)), int(rng.integers(1, 7))
        M, N = int(rng.integers(1, 36)), int(rng.integers(1, 36))
        A, B, LB, RB = random_problem(rng, K, L, M, N, band=band, alphabet=("acgt", "mixed", "weird")[it % 3])
        LB, RB = [int(v) for v in LB], [int(v) for v in RB]
        if any(LB[r] > RB[r - 1] + 1 for r in range(1, M + 1)):
            continue                                   # (the library keeps the guards for disconnected bands)
        want = oracle.yama(A, B, LB, RB, want_tback=True)
        cdi, tb = count_form_yama(np.asarray(A), np.asarray(B), LB, RB, S6, GO, GE, gates=False)
        assert cdi == tuple(int(v) for v in want["cdi"]), (band, it)
        assert _script(tb, cdi, M, N, LB, RB) == [int(v) for v in want["script"]], (band, it)
        changed_bytes += int((tb != want["tback"]).sum())
    # (on connected bands the guards rarely decide even an unreachable node's byte: `changed_bytes` is usually 0)
