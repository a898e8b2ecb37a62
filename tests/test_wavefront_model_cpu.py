"""CPU suite: the data flow of the bulk fill kernel (fill_body2, multiz_b200/csrc/yama_kernels.cuh) under K0's schedule
(yb_plan_kernel, plan_kernels.cuh), as a symbolic simulation.  No arithmetic: every grid point is a token (r, c); a lane that
is outside its row hands down the token STALE (the never-written dp[] entries of mz_yama.c:93-94, MININT in the kernel).  The
model executes the protocol step by step -- three shuffles from the lane above, lane 0 reading and lane 31 writing the ring,
the row switch at steps that are multiples of YB_F2_SW, the idle stores of the ring lane -- and asserts for EVERY cell of the
band that the lane computing it holds exactly the three neighbours the recurrence of mz_yama.c:113-242 reads:

    left  (r, c-1)   or nothing at the row's first cell
    up    (r-1, c)   if that grid point lies in row r-1's band, else STALE
    diag  (r-1, c-1) likewise

and that every cell of the band is computed exactly once.  This is the executable form of the schedule's argument: slack
3 + (YB_F2_SW - 1), a block of rows starting at least 32 steps after the one above, the ring of RING >= widest row + 32
entries.  Too little slack must FAIL the model (the A/B run that shipped YB_F2_SW = 4 found YB_SLACK=5 breaking real
pairs; the model must say so without a GPU).
"""
import numpy as np
import pytest

from tools.synth import SynthBatch, random_band

B = 32
STALE = ("stale",)


def schedule(M, LB, RB, slack):
    """yb_plan_kernel: one offset per block of 32 rows; rows 32b+1..32b+32 run on lanes 0..31 with column = step - (off_b + lane)."""
    nblk = (M + B - 1) // B
    off, out = 0, []
    for b in range(nblk):
        out.append(off)
        if b == nblk - 1:
            break
        nd = B
        for r in range(B * b + 1, min(M - B, B * b + B) + 1):
            nd = max(nd, int(RB[r + 1]) - int(LB[r + B]) + slack)
        off += nd
    last = out[-1] + (M - 1) % B + int(RB[M])
    return out, ((last + 2) + 7) & ~7


def simulate(M, N, LB, RB, sw, slack, ring_entries):
    sched, n_steps = schedule(M, LB, RB, slack)
    off = lambda r: sched[(r - 1) // B] + (r - 1) % B
    ring = {}
    for c in range(0, int(RB[1]) + 1):                       # row 0 (mz_yama.c:83-94): computed up to RB[0], stale up to RB[1]
        ring[c % ring_entries] = (0, c) if c <= RB[0] else STALE
    row = [l + 1 for l in range(B)]                          # the row each lane is on (> M: out of rows)
    col0 = [-(off(r)) if r <= M else None for r in row]      # column at step 0
    last = [STALE] * B                                       # what each lane computed in the previous step (its hand-down)
    prev_up = [STALE] * B                                    # the up value of the previous step = this step's diagonal
    seen = set()
    for t in range(n_steps):
        ups = [None] * B
        for l in range(B):                                   # top of the step: everybody reads
            if l == 0:
                r = row[0]
                c = t + col0[0] if r <= M else None
                ups[0] = ring.get(c % ring_entries, ("never",)) if c is not None else STALE
            else:
                ups[l] = last[l - 1]
        new_last = [STALE] * B
        for l in range(B):
            r = row[l]
            if r <= M:
                c = t + col0[l]
                if (sw == 1 or t % sw == 0) and c > RB[r]:   # the row switch
                    if l == B - 1:                           # the ring lane leaves the columns its reader still asks for
                        rbn = int(RB[r + 1]) if r < M else int(RB[r])
                        for cc in range(int(RB[r]) + 1 if sw == 1 else c, rbn + 1):
                            ring[cc % ring_entries] = STALE
                    r += B
                    row[l] = r
                    if r <= M:
                        col0[l] = -off(r)
                        c = t + col0[l]
            if r > M:
                new_last[l] = STALE
                continue
            inside = LB[r] <= c <= RB[r]
            if inside:
                assert (r, c) not in seen, f"cell {(r, c)} computed twice"
                seen.add((r, c))
                want_up = (r - 1, c) if LB[r - 1] <= c <= RB[r - 1] else STALE
                want_diag = (r - 1, c - 1) if LB[r - 1] <= c - 1 <= RB[r - 1] else STALE
                want_left = (r, c - 1) if c - 1 >= LB[r] else STALE
                if c > LB[r - 1]:                            # the C node exists (mz_yama.c:202-204): its diagonal is read
                    assert prev_up[l] == want_diag, f"cell {(r, c)} step {t}: diagonal {prev_up[l]} != {want_diag}"
                assert ups[l] == want_up, f"cell {(r, c)} step {t}: up {ups[l]} != {want_up}"
                assert last[l] == want_left, f"cell {(r, c)} step {t}: left {last[l]} != {want_left}"
                new_last[l] = (r, c)
            if l == B - 1 and c >= LB[r]:                    # the ring lane stores every step from its row's first column on
                ring[c % ring_entries] = (r, c) if inside else STALE
        prev_up = ups
        last = new_last
    want = sum(int(RB[r]) - int(LB[r]) + 1 for r in range(1, M + 1))
    assert len(seen) == want, f"{len(seen)} cells computed, the band has {want}"


def bands(seed, n):
    rng = np.random.default_rng(seed)
    out = []
    for it in range(n):
        M, N = int(rng.integers(1, 150)), int(rng.integers(1, 110))
        kind = ("smooth", "full", "ragged")[it % 3]
        if kind == "full" and N > 90:
            N = 90
        LB, RB = random_band(rng, M, N, kind)
        if all(LB[r] <= RB[r - 1] + 1 for r in range(1, M + 1)):      # the bulk class takes connected bands only
            out.append((M, N, LB.astype(np.int64), RB.astype(np.int64)))
    sb = SynthBatch(5, [2, 3, 2, 4], [1, 1, 2, 1], [400, 97, 33, 260], R=30)    # pre_yama-like diagonal bands
    for i in range(sb.n):
        _, _, LB, RB = sb.problem(i)
        out.append((int(sb.M[i]), int(sb.N[i]), LB.astype(np.int64), RB.astype(np.int64)))
    return out


@pytest.mark.parametrize("sw", [1, 2, 4, 8])
def test_every_cell_sees_its_three_neighbours(sw):
    for M, N, LB, RB in bands(11 + sw, 45):
        wmax = int((RB - LB).max()) + 1
        simulate(M, N, LB, RB, sw, slack=3 + sw - 1, ring_entries=128 if wmax + 32 <= 128 else 512)


def test_too_little_slack_is_caught():
    """YB_F2_SW = 4 with the slack of YB_F2_SW = 1: a lane switches rows too late for its new row's first cells."""
    broken = 0
    for M, N, LB, RB in bands(3, 60):
        try:
            simulate(M, N, LB, RB, 4, slack=3, ring_entries=512)
        except AssertionError:
            broken += 1
    assert broken > 0


def test_a_ring_smaller_than_the_widest_row_is_caught():
    broken = total = 0
    for M, N, LB, RB in bands(7, 80):
        wmax = int((RB - LB).max()) + 1
        if wmax < 40 or M < 40:
            continue
        total += 1
        try:
            simulate(M, N, LB, RB, 4, slack=6, ring_entries=wmax - 8)
        except AssertionError:
            broken += 1
    assert total > 5 and broken > 0
