"""Seeded synthetic template databases and reads for the BASELINE.json configs.

Everything is generated from a numpy PRNG seeded by the caller; FASTA/FASTQ text is what the
reference CLI consumes, the numeric forms (uint8 codes 0..3, 4 = N) feed the 2-bit packers in
``kma_b200.records`` directly.
"""
from __future__ import annotations

import numpy as np

_BASES = np.frombuffer(b"ACGTN", dtype=np.uint8)
_COMP = np.array([3, 2, 1, 0, 4], dtype=np.uint8)


def decode(codes: np.ndarray) -> str:
    return _BASES[codes].tobytes().decode()


def revcomp(codes: np.ndarray) -> np.ndarray:
    return _COMP[codes[::-1]]


def mutate_subs(rng, seq, rate):
    """Substitutions only: every hit base is replaced by one of the 3 other bases."""
    if rate <= 0:
        return seq
    out = seq.copy()
    hit = np.flatnonzero(rng.random(len(seq)) < rate)
    out[hit] = (out[hit] + rng.integers(1, 4, size=len(hit)).astype(np.uint8)) & 3
    return out


def mutate_indel(rng, seq, sub, ins, dele):
    """Per-base substitution / insertion / deletion (Nanopore-like)."""
    r = rng.random(len(seq))
    keep = r >= dele
    s = seq.copy()
    hit = np.flatnonzero((r >= dele) & (r < dele + sub))
    s[hit] = (s[hit] + rng.integers(1, 4, size=len(hit)).astype(np.uint8)) & 3
    s = s[keep]
    n_ins = rng.binomial(len(s), ins)
    if n_ins:
        pos = np.sort(rng.integers(0, len(s) + 1, size=n_ins))
        s = np.insert(s, pos, rng.integers(0, 4, size=n_ins).astype(np.uint8))
    return s


def gene_db(seed: int, n_families: int = 300, n_variants: int = 10, len_lo: int = 500, len_hi: int = 3000,
            indel_frac: float = 0.1):
    """Redundant AMR-gene-style DB: families of near-identical variants (C1/C2/C3/C5 shape)."""
    rng = np.random.default_rng(seed)
    names, seqs = [], []
    for f in range(n_families):
        L = int(rng.integers(len_lo, len_hi + 1))
        base = rng.integers(0, 4, size=L).astype(np.uint8)
        rate = float(rng.choice([0.002, 0.005, 0.01, 0.02]))
        for v in range(n_variants):
            if v == 0:
                s = base
            else:
                s = mutate_indel(rng, base, rate, rate * indel_frac * 0.5, rate * indel_frac * 0.5)
            names.append(f"fam{f:05d}_v{v:02d}")
            seqs.append(s)
    return names, seqs


def genome_db(seed: int, length: int = 5_000_000):
    rng = np.random.default_rng(seed)
    return ["genome1"], [rng.integers(0, 4, size=length).astype(np.uint8)]


def write_fasta(path, names, seqs, width=0):
    with open(path, "w") as fh:
        for n, s in zip(names, seqs):
            fh.write(f">{n}\n{decode(s)}\n")


def _concat(seqs):
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    offs = np.r_[0, np.cumsum(lens)]
    return np.concatenate(seqs), lens, offs


def short_reads(seed: int, seqs, n: int, L: int = 150, sub: float = 0.005, n_rate: float = 0.0,
                junk_frac: float = 0.0):
    """Single-end reads: uniform template, uniform start, random strand, substitutions (+ optional N's,
    + optional fraction of random reads that map nowhere). Vectorised; uint8 [n, L]."""
    rng = np.random.default_rng(seed)
    cat, lens, offs = _concat(seqs)
    ok = np.flatnonzero(lens >= L)
    tsel = ok[rng.integers(0, len(ok), size=n)]
    start = offs[tsel] + (rng.random(n) * (lens[tsel] - L + 1)).astype(np.int64)
    out = cat[start[:, None] + np.arange(L)[None, :]]
    hit = rng.random((n, L)) < sub
    out[hit] = (out[hit] + rng.integers(1, 4, size=int(hit.sum())).astype(np.uint8)) & 3
    if junk_frac > 0:
        junk = np.flatnonzero(rng.random(n) < junk_frac)
        out[junk] = rng.integers(0, 4, size=(len(junk), L)).astype(np.uint8)
    rc = rng.random(n) < 0.5
    out[rc] = _COMP[out[rc][:, ::-1]]
    if n_rate > 0:
        out[rng.random((n, L)) < n_rate] = 4
    return out


def paired_reads(seed: int, seqs, n: int, L: int = 150, sub: float = 0.005, ins_lo: int = 200, ins_hi: int = 450):
    """FR pairs with insert U[ins_lo, ins_hi] clipped to the template. Vectorised; two uint8 [n, L]."""
    rng = np.random.default_rng(seed)
    cat, lens, offs = _concat(seqs)
    ok = np.flatnonzero(lens >= L)
    tsel = ok[rng.integers(0, len(ok), size=n)]
    ins = np.minimum(lens[tsel], np.maximum(L, rng.integers(ins_lo, ins_hi + 1, size=n)))
    start = offs[tsel] + (rng.random(n) * (lens[tsel] - ins + 1)).astype(np.int64)
    left = cat[start[:, None] + np.arange(L)[None, :]]                       # fragment 5' end, forward
    right = _COMP[cat[(start + ins - 1)[:, None] - np.arange(L)[None, :]]]   # fragment 3' end, reverse strand
    flip = rng.random(n) < 0.5
    r1 = np.where(flip[:, None], right, left)
    r2 = np.where(flip[:, None], left, right)
    for r in (r1, r2):
        hit = rng.random((n, L)) < sub
        r[hit] = (r[hit] + rng.integers(1, 4, size=int(hit.sum())).astype(np.uint8)) & 3
    return r1, r2


def long_reads(seed: int, seqs, n: int, len_lo: int = 5000, len_hi: int = 20000, err: float = 0.10):
    """Nanopore-like reads: concatenated random templates and random spacers, 10 % error (1/3 sub, del, ins)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        L = int(rng.integers(len_lo, len_hi + 1))
        parts, tot = [], 0
        while tot < L:
            t = seqs[int(rng.integers(0, len(seqs)))]
            if rng.random() < 0.5:
                t = revcomp(t)
            sp = rng.integers(0, 4, size=int(rng.integers(200, 2001))).astype(np.uint8)
            parts += [t, sp]
            tot += len(t) + len(sp)
        s = np.concatenate(parts)[:L]
        out.append(mutate_indel(rng, s, err / 3, err / 3, err / 3))
    return out


def write_fastq(path, reads, prefix="r", qual="I"):
    if isinstance(reads, np.ndarray) and reads.ndim == 2:      # fixed-length reads: one byte matrix per id width
        n, L = reads.shape
        head = np.frombuffer(f"@{prefix}".encode(), dtype=np.uint8)
        c = len(head)
        with open(path, "wb") as fh:
            lo, w = 0, 1
            while lo < n:
                hi = min(n, 10 ** w)
                m = hi - lo
                ids = np.arange(lo, hi).astype(f"S{w}").view(np.uint8).reshape(m, w)
                rec = np.empty((m, c + w + 1 + L + 3 + L + 1), dtype=np.uint8)
                rec[:, :c] = head
                rec[:, c:c + w] = ids
                rec[:, c + w] = 10
                rec[:, c + w + 1:c + w + 1 + L] = _BASES[reads[lo:hi]]
                rec[:, c + w + 1 + L:c + w + 4 + L] = np.frombuffer(b"\n+\n", dtype=np.uint8)
                rec[:, c + w + 4 + L:c + w + 4 + 2 * L] = ord(qual)
                rec[:, -1] = 10
                fh.write(rec.tobytes())
                lo, w = hi, w + 1
        return
    with open(path, "w") as fh:
        for i, r in enumerate(reads):
            s = decode(np.asarray(r))
            fh.write(f"@{prefix}{i}\n{s}\n+\n{qual * len(s)}\n")


def fastq_fixed(reads: np.ndarray, prefix: str = "r", first: int = 0, qual: int = 73) -> np.ndarray:
    """Vectorised 4-line FASTQ text (uint8 array) of equal-length reads (codes 0-4) with fixed-width names
    '@<prefix><9-digit index>' -- the text whose stage-1 stream is records.stage1_records_fast(reads, prefix, first)."""
    n, L = reads.shape
    pre = prefix.encode()
    hl = 1 + len(pre) + 9
    rec = np.empty((n, hl + 1 + L + 3 + L + 1), dtype=np.uint8)
    rec[:, 0] = ord("@")
    rec[:, 1:1 + len(pre)] = np.frombuffer(pre, dtype=np.uint8)
    idx = np.arange(first, first + n, dtype=np.int64)
    for d in range(9):
        rec[:, hl - 1 - d] = 48 + (idx // 10 ** d) % 10
    rec[:, hl] = 10
    rec[:, hl + 1:hl + 1 + L] = np.frombuffer(b"ACGTN", dtype=np.uint8)[reads]
    rec[:, hl + 1 + L:hl + 4 + L] = np.frombuffer(b"\n+\n", dtype=np.uint8)
    rec[:, hl + 4 + L:hl + 4 + 2 * L] = qual
    rec[:, -1] = 10
    return rec.reshape(-1)
