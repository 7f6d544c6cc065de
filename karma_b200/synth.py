"""Synthetic Trinity-like assemblies (S0-S3).

All generators are numpy-vectorised and seeded with ``numpy.random.default_rng``
(PCG64).  They return the packed form the C-ABI consumes --
``bases`` uint8[sum L] (ASCII, uppercase ACGT), ``offsets`` int64[N+1] and
``key_len`` int32[N] (= len(">name"), what kmer.py:213 divides by) -- plus a
lazy way to get the ``OrderedDict{">name": seq}`` karma.py:40-61 would build.
"""
from collections import OrderedDict

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _lengths(rng, n, lo=200, hi=15000):
    # clip(rint(lognormal(ln 700, 0.9)), 200, 15000)  -- SURVEY.md 8d / BASELINE.md 2
    return np.clip(np.rint(rng.lognormal(np.log(700.0), 0.9, n)), lo, hi).astype(np.int64)


def _key_len(gene, iso):
    # len(">TRINITY_DN{gene}_c0_g1_i{iso}")
    def ndig(x):
        return np.floor(np.log10(np.maximum(x, 1))).astype(np.int64) + 1
    return (len(">TRINITY_DN") + ndig(gene) + len("_c0_g1_i") + ndig(iso)).astype(np.int32)


class Assembly:
    """Packed synthetic assembly."""

    def __init__(self, bases, offsets, gene, iso):
        self.bases = bases
        self.offsets = offsets
        self.gene = gene
        self.iso = iso
        self.key_len = _key_len(gene, iso)

    @property
    def n(self):
        return len(self.offsets) - 1

    def names(self):
        return [f">TRINITY_DN{g}_c0_g1_i{i}" for g, i in zip(self.gene.tolist(), self.iso.tolist())]

    def as_dict(self):
        """OrderedDict{">name": sequence} exactly as read_fasta_file would hand it
        to KmerClustering (karma.py:40-61)."""
        raw = self.bases.tobytes().decode("ascii")
        off = self.offsets.tolist()
        d = OrderedDict()
        for r, name in enumerate(self.names()):
            d[name] = raw[off[r]:off[r + 1]]
        return d

    def slice(self, lo, hi):
        """Rows [lo, hi) as a new Assembly (row shard for one rank)."""
        o = self.offsets[lo:hi + 1]
        return Assembly(self.bases[o[0]:o[-1]].copy(), (o - o[0]).copy(),
                        self.gene[lo:hi], self.iso[lo:hi])


def s0_iid(n, seed=1234):
    """S0: i.i.d. uniform ACGT, log-normal lengths (the survey generator)."""
    rng = np.random.default_rng(seed)
    lens = _lengths(rng, n)
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    bases = _ACGT[rng.integers(0, 4, int(offsets[-1]), dtype=np.uint8)]
    return Assembly(bases, offsets, np.arange(n, dtype=np.int64), np.ones(n, dtype=np.int64))


def s1_families(n, seed=1235, r=4, sub_rate=0.01, dup_frac=0.0, long_frac=0.0,
                long_lo=50_000, long_hi=200_000):
    """S1 gene families: G=ceil(n/r) genes (log-normal length, per-gene GC ~
    N(0.45,0.05)); each contig is a random sub-interval covering >=60 % of its
    gene with ``sub_rate`` substitutions.  ``dup_frac`` of the contigs are exact
    duplicates of an earlier contig (S2); ``long_frac`` of the genes are long
    (S3)."""
    rng = np.random.default_rng(seed)
    g = max(1, -(-n // r))
    glen = _lengths(rng, g)
    if long_frac > 0:
        nl = max(1, int(round(g * long_frac)))
        which = rng.choice(g, nl, replace=False)
        glen[which] = rng.integers(long_lo, long_hi + 1, nl)
    goff = np.zeros(g + 1, dtype=np.int64)
    np.cumsum(glen, out=goff[1:])
    gc = np.clip(rng.normal(0.45, 0.05, g), 0.2, 0.8)
    tot = int(goff[-1])
    gc_pos = np.repeat(gc, glen)
    u = rng.random(tot)
    is_gc = u < gc_pos
    pick = rng.integers(0, 2, tot, dtype=np.uint8)
    # A=0 C=1 G=2 T=3 : GC -> {C,G}, AT -> {A,T}
    gene_codes = np.where(is_gc, 1 + pick, 3 * pick).astype(np.uint8)

    gene = np.sort(rng.integers(0, g, n)) if n > g else np.arange(n, dtype=np.int64)
    gene = gene.astype(np.int64)
    # isoform index within gene (1-based)
    first = np.r_[True, gene[1:] != gene[:-1]]
    grp_start = np.maximum.accumulate(np.where(first, np.arange(n), 0))
    iso = (np.arange(n) - grp_start + 1).astype(np.int64)

    L = glen[gene]
    frac = rng.uniform(0.6, 1.0, n)
    clen = np.maximum(np.minimum(L, 200), np.rint(L * frac).astype(np.int64))
    clen = np.minimum(clen, L)
    start = (rng.random(n) * (L - clen + 1)).astype(np.int64)

    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(clen, out=offsets[1:])
    total = int(offsets[-1])
    src0 = goff[gene] + start
    idx = np.repeat(src0 - offsets[:-1], clen) + np.arange(total, dtype=np.int64)
    codes = gene_codes[idx]
    if sub_rate > 0:
        m = rng.random(total) < sub_rate
        codes = np.where(m, (codes + rng.integers(1, 4, total, dtype=np.uint8)) & 3, codes).astype(np.uint8)
    bases = _ACGT[codes]
    asm = Assembly(bases, offsets, gene, iso)
    if dup_frac > 0 and n > 1:
        nd = int(n * dup_frac)
        dst = rng.choice(np.arange(1, n), nd, replace=False)
        src = (rng.random(nd) * dst).astype(np.int64)
        # rebuild with duplicated sequences (lengths change)
        lens = np.diff(offsets)
        src_of = np.arange(n)
        src_of[dst] = src_of[src]           # one level is enough for a synthetic
        nl = lens[src_of]
        noff = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(nl, out=noff[1:])
        idx = np.repeat(offsets[:-1][src_of] - noff[:-1], nl) + np.arange(int(noff[-1]), dtype=np.int64)
        asm = Assembly(bases[idx], noff, gene, iso)
    return asm


def s2_redundant(n, seed=1237):
    """S2: redundant multi-assembler merge -- r=16 isoforms/gene, 10 % exact duplicates."""
    return s1_families(n, seed=seed, r=16, dup_frac=0.10)


def s3_long(n, seed=1238):
    """S3: S1 plus 1 % long genes (50-200 kb) for the long-contig split path."""
    return s1_families(n, seed=seed, r=4, long_frac=0.01)


def make(kind, n, seed=None):
    f = {"S0": s0_iid, "S1": s1_families, "S2": s2_redundant, "S3": s3_long}[kind]
    return f(n) if seed is None else f(n, seed=seed)
