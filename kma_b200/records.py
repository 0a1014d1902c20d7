"""Stage-1 record stream (runinput.c:765-787 printFsa / printFsa_pair) built with numpy, and parsers for
the stage-1 / stage-2 (ankers.c:30-50) wire formats. Host-side glue only: no alignment logic here."""
from __future__ import annotations

import numpy as np


def pack_2bit(codes: np.ndarray):
    """codes: uint8[L] in 0..4 (4 = N) -> (uint64 words MSB-first, int32 N positions). compdna.c:99-127"""
    L = len(codes)
    words = (L + 31) >> 5
    c = np.zeros(words * 32, dtype=np.uint64)
    npos = np.flatnonzero(codes == 4).astype(np.int32)
    c[:L] = np.where(codes == 4, 0, codes)
    sh = (np.uint64(62) - np.uint64(2) * (np.arange(words * 32, dtype=np.uint64) & np.uint64(31)))
    w = np.bitwise_or.reduce((c << sh).reshape(words, 32), axis=1) if words else np.zeros(0, np.uint64)
    return w.astype(np.uint64), npos


def stage1_records_fixed(reads: np.ndarray, prefix: str = "r", pair_with: np.ndarray | None = None) -> np.ndarray:
    """Vectorised stage-1 stream for equal-length reads (uint8 [n, L] codes). Header i is b'<prefix><i>\\0'
    padded... no: real variable-length names, exactly what `kma -s1` would emit for '@<prefix><i>'."""
    n, L = reads.shape
    words = (L + 31) >> 5
    c = np.zeros((n, words * 32), dtype=np.uint64)
    c[:, :L] = np.where(reads == 4, 0, reads)
    sh = (np.uint64(62) - np.uint64(2) * (np.arange(words * 32, dtype=np.uint64) & np.uint64(31)))
    packed = np.bitwise_or.reduce((c << sh).reshape(n, words, 32), axis=2)  # [n, words] uint64
    out = bytearray()
    has_n = (reads == 4).any(axis=1)
    names = [f"{prefix}{i}".encode() + b"\0" for i in range(n)]
    for i in range(n):
        npos = np.flatnonzero(reads[i] == 4).astype(np.int32) if has_n[i] else np.zeros(0, np.int32)
        out += np.array([L, words, len(npos), len(names[i])], dtype=np.int32).tobytes()
        out += packed[i].tobytes()
        out += npos.tobytes()
        out += names[i]
    return np.frombuffer(bytes(out), dtype=np.uint8)


def stage1_records_fast(reads: np.ndarray, prefix: str = "r", first: int = 0) -> np.ndarray:
    """Fully vectorised stage-1 stream for N-free equal-length reads with fixed-width names
    '<prefix><9-digit index>\\0' (what `kma -s1` emits for a FASTQ with those names)."""
    n, L = reads.shape
    assert not (reads == 4).any(), "N-free reads only (use stage1_records_fixed)"
    words = (L + 31) >> 5
    c = np.zeros((n, words * 32), dtype=np.uint64)
    c[:, :L] = reads
    sh = (np.uint64(62) - np.uint64(2) * (np.arange(words * 32, dtype=np.uint64) & np.uint64(31)))
    packed = np.bitwise_or.reduce((c << sh).reshape(n, words, 32), axis=2)
    pre = prefix.encode()
    hl = len(pre) + 10
    rec = np.zeros((n, 16 + 8 * words + hl), dtype=np.uint8)
    rec[:, :16] = np.frombuffer(np.array([L, words, 0, hl], dtype=np.int32).tobytes(), dtype=np.uint8)
    rec[:, 16:16 + 8 * words] = packed.view(np.uint8).reshape(n, 8 * words)
    o = 16 + 8 * words
    rec[:, o:o + len(pre)] = np.frombuffer(pre, dtype=np.uint8)
    idx = np.arange(first, first + n, dtype=np.int64)
    for d in range(9):
        rec[:, o + len(pre) + 8 - d] = 48 + (idx // 10 ** d) % 10
    return rec.reshape(-1)


def stage1_pairs_fast(r1: np.ndarray, r2: np.ndarray, prefix: str = "r", first: int = 0) -> np.ndarray:
    """Fully vectorised stage-1 stream of N-free equal-length read pairs (printFsa_pair, runinput.c:789) with
    fixed-width names '<prefix><9-digit index>\\0' in both files: mate records interleaved, first mate with a negative
    header length."""
    n, L = r1.shape
    a = stage1_records_fast(r1, prefix, first).reshape(n, -1)
    b = stage1_records_fast(r2, prefix, first).reshape(n, -1)
    hl = len(prefix.encode()) + 10
    a[:, 12:16] = np.frombuffer(np.array([-hl], dtype=np.int32).tobytes(), dtype=np.uint8)
    return np.concatenate([a, b], axis=1).reshape(-1)


def stage1_pairs(r1, r2, prefix="r") -> np.ndarray:
    """Stage-1 stream of read pairs (printFsa_pair, runinput.c:789): mate records interleaved, the first mate written
    with a NEGATIVE header length. Names as `kma -ipe a.fq b.fq` emits them for '@<prefix><i>' in both files."""
    out = bytearray()
    for i, (a, b) in enumerate(zip(r1, r2)):
        name = f"{prefix}{i}".encode() + b"\0"
        for mate, r in enumerate((a, b)):
            r = np.asarray(r, dtype=np.uint8)
            w, npos = pack_2bit(r)
            out += np.array([len(r), len(w), len(npos), -len(name) if mate == 0 else len(name)], dtype=np.int32).tobytes()
            out += w.tobytes() + npos.tobytes() + name
    return np.frombuffer(bytes(out), dtype=np.uint8)


def stage1_records(reads, names=None, prefix="r") -> np.ndarray:
    """General (ragged) stage-1 stream."""
    out = bytearray()
    for i, r in enumerate(reads):
        r = np.asarray(r, dtype=np.uint8)
        w, npos = pack_2bit(r)
        name = (names[i] if names is not None else f"{prefix}{i}").encode() + b"\0"
        out += np.array([len(r), len(w), len(npos), len(name)], dtype=np.int32).tobytes()
        out += w.tobytes() + npos.tobytes() + name
    return np.frombuffer(bytes(out), dtype=np.uint8)


def parse_stage2(buf) -> list[dict]:
    """Split a stage-2 stream into records (dicts); stops at the terminator if present."""
    b = memoryview(np.ascontiguousarray(buf, dtype=np.uint8)).tobytes()
    recs, p = [], 0
    while p + 28 <= len(b):
        h = np.frombuffer(b, dtype=np.int32, count=7, offset=p)
        if h[0] < 0:
            break
        seqlen, words, nN, score, nt, hl, flag = (int(x) for x in h)
        p += 28
        seq = np.frombuffer(b, dtype=np.uint64, count=words, offset=p) if p % 8 == 0 else \
            np.frombuffer(b[p:p + 8 * words], dtype=np.uint64)
        p += 8 * words
        N = np.frombuffer(b[p:p + 4 * nN], dtype=np.int32); p += 4 * nN
        T = np.frombuffer(b[p:p + 4 * nt], dtype=np.int32); p += 4 * nt
        name = b[p:p + hl]; p += hl
        recs.append(dict(seqlen=seqlen, seq=seq, N=N, score=score, templates=T, name=name, flag=flag))
    return recs
