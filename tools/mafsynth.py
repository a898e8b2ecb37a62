#!/usr/bin/env python
"""Synthetic multi-species MAFs for multiz / roast (measurement + test infrastructure; no network, so no
real genomes).  SURVEY.md §8(d) cfg 1/2: a random ACGT ancestor, every species = ancestor + substitutions +
indel events of length U[1,8]; the TRUE pairwise alignment of the reference species against each other
species is cut into blocks of U[200,2000] columns separated by U[0,50] unaligned columns, block ends trimmed
to columns where both rows carry a base.  Files are what `multiz ref.sp1.maf ref.sp2.maf v` expects
(Appendix C): same top `src`, sorted by reference start, single coverage.

  python tools/mafsynth.py --out DIR --ref-len 1000000 --species 2 [--seed 1]
"""
from __future__ import annotations

import argparse
import os

import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
DASH = ord("-")


def _runs(rng, n, rate, maxlen=8):
    """Per-site run lengths: an event starts at a site with prob `rate`, length U[1,maxlen]."""
    ev = rng.random(n) < rate
    return np.where(ev, rng.integers(1, maxlen + 1, size=n), 0).astype(np.int64)


class Species:
    """One species in ancestor coordinates: base per ancestor site ('-' if deleted) + insertions after sites."""

    def __init__(self, rng, anc, sub, indel, lower=0.0):
        n = len(anc)
        b = anc.copy()
        mut = rng.random(n) < sub
        b[mut] = BASES[(np.searchsorted(BASES, anc[mut]) + rng.integers(1, 4, size=int(mut.sum()))) & 3]
        if lower > 0:
            lo = rng.random(n) < lower
            b[lo] |= 0x20
        dl = _runs(rng, n, indel / 2)
        deleted = np.zeros(n + 9, dtype=np.int64)
        idx = np.nonzero(dl)[0]
        np.add.at(deleted, idx, 1)
        np.add.at(deleted, idx + dl[idx], -1)
        deleted = np.cumsum(deleted)[:n] > 0
        b[deleted] = DASH
        self.base = b
        self.ins = _runs(rng, n, indel / 2)
        self.ins[-1] = 0
        self.rng_seed = int(rng.integers(1 << 62))


def pairwise_columns(ref: Species, sp: Species):
    """True alignment of two species: (top, bot) uint8 arrays, no all-dash columns."""
    n = len(ref.base)
    keep = ~((ref.base == DASH) & (sp.base == DASH))
    per = keep.astype(np.int64) + ref.ins + sp.ins
    off = np.concatenate([[0], np.cumsum(per)])
    T = int(off[-1])
    top = np.full(T, DASH, np.uint8)
    bot = np.full(T, DASH, np.uint8)
    k = np.nonzero(keep)[0]
    top[off[k]] = ref.base[k]
    bot[off[k]] = sp.base[k]
    rng = np.random.default_rng(ref.rng_seed ^ sp.rng_seed)
    for who, other_shift, dst in ((ref, keep.astype(np.int64), top), (sp, keep.astype(np.int64) + ref.ins, bot)):
        s = np.nonzero(who.ins)[0]
        if len(s) == 0:
            continue
        ln = who.ins[s]
        start = off[s] + other_shift[s]
        pos = np.repeat(start - np.concatenate([[0], np.cumsum(ln)[:-1]]), ln) + np.arange(int(ln.sum()))
        # inserted bases are a property of the species, not of the pair: seed from the species alone
        dst[pos] = BASES[np.random.default_rng(who.rng_seed).integers(0, 4, size=len(pos))]
    del rng, n
    return top, bot


def write_pairwise_maf(path, ref_name, sp_name, top, bot, seed, blk=(200, 2000), gap=(0, 50)):
    rng = np.random.default_rng(seed)
    T = len(top)
    tb, bb = top != DASH, bot != DASH
    both = tb & bb
    ctop = np.concatenate([[0], np.cumsum(tb)])
    cbot = np.concatenate([[0], np.cumsum(bb)])
    ref_size, sp_size = int(ctop[-1]), int(cbot[-1])
    nb = 0
    with open(path, "w") as f:
        f.write("##maf version=1 scoring=synthetic\n")
        c = int(rng.integers(gap[0], gap[1] + 1))
        while c < T:
            e = min(T, c + int(rng.integers(blk[0], blk[1] + 1)))
            idx = np.nonzero(both[c:e])[0]
            if len(idx) >= 2:
                s, t = c + int(idx[0]), c + int(idx[-1]) + 1
                rs, rn = int(ctop[s]), int(ctop[t] - ctop[s])
                ss_, sn = int(cbot[s]), int(cbot[t] - cbot[s])
                w = max(len(str(rs)), len(str(ss_)))
                w2 = max(len(str(rn)), len(str(sn)))
                f.write("a score=0.0\n")
                f.write(f"s {ref_name:<12} {rs:>{w}} {rn:>{w2}} + {ref_size} {top[s:t].tobytes().decode()}\n")
                f.write(f"s {sp_name:<12} {ss_:>{w}} {sn:>{w2}} + {sp_size} {bot[s:t].tobytes().decode()}\n\n")
                nb += 1
            c = e + int(rng.integers(gap[0], gap[1] + 1))
        f.write("##eof maf\n")
    return nb


def make_dataset(out, ref_len=100_000, n_species=2, seed=1, sub=0.10, indel=0.01, lower=0.0, blk=(200, 2000)):
    """Writes out/ref.spI.maf for I=1..n_species; returns the list of paths."""
    os.makedirs(out, exist_ok=True)
    rng = np.random.default_rng(seed)
    anc = BASES[rng.integers(0, 4, size=ref_len)]
    ref = Species(rng, anc, sub / 2, indel / 2)
    paths = []
    for i in range(1, n_species + 1):
        sp = Species(rng, anc, sub, indel, lower)
        top, bot = pairwise_columns(ref, sp)
        p = os.path.join(out, f"ref.sp{i}.maf")
        write_pairwise_maf(p, "ref.chr1", f"sp{i}.chr1", top, bot, seed * 1000 + i, blk=blk)
        paths.append(p)
    return paths


def species_sequence(sp: Species):
    """The species' own sequence (what its FASTA file holds): its bases in ancestor order, insertions after their site."""
    ins_bases = BASES[np.random.default_rng(sp.rng_seed).integers(0, 4, size=int(sp.ins.sum()))]
    n = len(sp.base)
    per = (sp.base != DASH).astype(np.int64) + sp.ins
    off = np.concatenate([[0], np.cumsum(per)])
    seq = np.zeros(int(off[-1]), np.uint8)
    k = np.nonzero(sp.base != DASH)[0]
    seq[off[k]] = sp.base[k]
    s = np.nonzero(sp.ins)[0]
    if len(s):
        ln = sp.ins[s]
        start = off[s] + (sp.base[s] != DASH)
        pos = np.repeat(start - np.concatenate([[0], np.cumsum(ln)[:-1]]), ln) + np.arange(int(ln.sum()))
        seq[pos] = ins_bases
    del n
    return seq


def make_tba_dataset(out, names, ref_len=50_000, seed=1, sub=0.10, indel=0.01, blk=(200, 2000)):
    """What the reference's `tba` wants in its working directory: for every ordered pair x < y of `names`
    the single-coverage pairwise file x.y.sing.maf (top row x), plus a FASTA file named exactly as each species with the
    header >name:chr:start:strand:srcSize (multi_util.c:311-322), which pair2tb reads to add the unaligned stretches."""
    os.makedirs(out, exist_ok=True)
    rng = np.random.default_rng(seed)
    anc = BASES[rng.integers(0, 4, size=ref_len)]
    sp = {nm: Species(rng, anc, sub if i else sub / 2, indel if i else indel / 2) for i, nm in enumerate(names)}
    for nm in names:
        seq = species_sequence(sp[nm])
        with open(os.path.join(out, nm), "w") as f:
            f.write(f">{nm}:chr1:1:+:{len(seq)}\n")
            txt = seq.tobytes().decode()
            for i in range(0, len(txt), 60):
                f.write(txt[i:i + 60] + "\n")
    files = []
    for i, x in enumerate(names):
        for j, y in enumerate(names):
            if i >= j:
                continue
            top, bot = pairwise_columns(sp[x], sp[y])
            fn = f"{x}.{y}.sing.maf"
            write_pairwise_maf(os.path.join(out, fn), f"{x}.chr1", f"{y}.chr1", top, bot, seed * 1000 + 37 * i + j, blk=blk)
            files.append(fn)
    return files


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--ref-len", type=int, default=1_000_000)
    ap.add_argument("--species", type=int, default=2)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--sub", type=float, default=0.10)
    ap.add_argument("--indel", type=float, default=0.01)
    a = ap.parse_args()
    for p in make_dataset(a.out, a.ref_len, a.species, a.seed, a.sub, a.indel):
        print(p)
