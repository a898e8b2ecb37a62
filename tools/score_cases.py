"""Seeded alignment blocks for the block-scoring path (mafScoreRange, mz_scores.c:124-152).

score_cases() -> list of (text[rows, textSize] u8, start, size): gapped alignments the way multiz makes them (rows are
mutated copies of an ancestor with runs of dashes), plus the shapes that stress the kernel's decomposition: one and two
rows, single columns, ranges that start inside the text, widths around the 128-column warp unit, more than 255 rows
(byte-lane counter flush), lower case / N / other letters, all-dash columns.
"""
from __future__ import annotations

import numpy as np

ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)


def alignment_block(rng, rows: int, cols: int, sub=0.1, gap_open=0.03, gap_len=6, lower=0.05, other=0.01) -> np.ndarray:
    anc = ALPHA[rng.integers(0, 4, cols)]
    blk = np.tile(anc, (rows, 1))
    mut = rng.random((rows, cols)) < sub
    blk[mut] = ALPHA[rng.integers(0, 4, int(mut.sum()))]
    low = rng.random((rows, cols)) < lower
    blk[low] |= 0x20
    oth = rng.random((rows, cols)) < other
    blk[oth] = np.frombuffer(b"NnXRy", dtype=np.uint8)[rng.integers(0, 5, int(oth.sum()))]
    for r in range(rows):                                   # runs of dashes
        c = 0
        while c < cols:
            if rng.random() < gap_open:
                ln = int(rng.integers(1, 2 * gap_len))
                blk[r, c:c + ln] = ord("-")
                c += ln
            c += 1
    return blk


def score_cases(seed=20260118):
    rng = np.random.default_rng(seed)
    cases = []
    for rows, cols in ((1, 30), (2, 1), (2, 57), (3, 127), (3, 128), (3, 129), (5, 255), (5, 256), (5, 257), (8, 513),
                       (12, 1000), (33, 300), (64, 140), (100, 90), (256, 37), (300, 70), (600, 9)):
        blk = alignment_block(rng, rows, cols)
        cases.append((blk, 0, cols))
        if cols > 3:
            st = int(rng.integers(1, cols - 1))
            cases.append((blk, st, int(rng.integers(1, cols - st + 1))))
            cases.append((blk, cols - 1, 1))
    for _ in range(40):                                     # random shapes and ranges
        rows, cols = int(rng.integers(2, 20)), int(rng.integers(1, 400))
        blk = alignment_block(rng, rows, cols, sub=float(rng.random()) * 0.3, gap_open=float(rng.random()) * 0.1)
        st = int(rng.integers(0, cols))
        cases.append((blk, st, int(rng.integers(1, cols - st + 1))))
    blk = alignment_block(rng, 6, 200)
    blk[:, 50:60] = ord("-")                                # all-dash columns
    blk[2, :] = ord("-")                                    # an all-dash row
    cases.append((blk, 0, 200))
    cases.append((blk, 55, 100))
    cases.append((np.full((4, 64), ord("-"), dtype=np.uint8), 0, 64))
    return cases
