"""Template database writer in the reference's on-disk format (host-side tool, numpy).

Writes <prefix>.comp.b / .length.b / .seq.b / .name with exactly the layout `kma index` produces
(hashmapkma.c:722-775 hashMapKMA_dump, makeindex.c:263-272, updateindex.c:172), so that both the
unmodified reference binary and libkmagpu read it. Only the in-scope shape is produced: plain
k-mers (flag 0, no prefix), k <= 16, forward strand only (updateindex.c:58-73).

The byte content is not required to equal `kma index` output (slot order inside a bucket and the
order of the de-duplicated template lists are free) -- lookups, and therefore mapping results, are
identical; tests/test_dbbuild.py checks that with the reference binary.
"""
from __future__ import annotations

import numpy as np

from .records import pack_2bit


def _kmers_of(seq: np.ndarray, k: int) -> np.ndarray:
    """all forward k-mers (uint64) of an N-free code array; k-mers touching an N are dropped"""
    L = len(seq)
    if L < k:
        return np.zeros(0, dtype=np.uint64)
    s = np.where(seq == 4, 0, seq).astype(np.uint64)
    km = np.zeros(L - k + 1, dtype=np.uint64)
    for i in range(k):
        km = (km << np.uint64(2)) | s[i:L - k + 1 + i]
    if (seq == 4).any():
        bad = np.convolve((seq == 4).astype(np.int32), np.ones(k, dtype=np.int32), mode="valid") > 0
        km = km[~bad]
    return km


def build_db(prefix: str, names, seqs, k: int = 16, initial_size: int = 1 << 20) -> dict:
    assert 4 <= k <= 16, "k <= 16 (32-bit keys) only"
    ntempl = len(seqs)
    DB_size = ntempl + 1
    # (k-mer, template) pairs, unique and sorted by k-mer then template id (hashmap.c:120-162 keeps lists sorted)
    parts = []
    for t, s in enumerate(seqs, start=1):
        km = _kmers_of(np.asarray(s, dtype=np.uint8), k)
        parts.append((km << np.uint64(32)) | np.uint64(t))
    pairs = np.unique(np.concatenate(parts))
    del parts
    keys_all = (pairs >> np.uint64(32)).astype(np.uint32)
    tids = (pairs & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    del pairs
    starts = np.flatnonzero(np.r_[True, keys_all[1:] != keys_all[:-1]])
    keys = keys_all[starts]
    n = len(keys)
    lens = np.diff(np.r_[starts, len(tids)]).astype(np.int64)

    # de-duplicate identical template lists (compress.c valuesHash_add): hash, group, verify
    rng = np.random.default_rng(0x6b6d61)
    maxlen = int(lens.max())
    R1 = rng.integers(1, 1 << 63, size=maxlen, dtype=np.uint64) | np.uint64(1)
    R2 = rng.integers(1, 1 << 63, size=maxlen, dtype=np.uint64) | np.uint64(1)
    within = np.arange(len(tids), dtype=np.int64) - np.repeat(starts, lens)
    t64 = tids.astype(np.uint64) + np.uint64(1)
    with np.errstate(over="ignore"):
        h1 = np.add.reduceat(t64 * R1[within], starts)
        h2 = np.add.reduceat((t64 ^ np.uint64(0x9E3779B97F4A7C15)) * R2[within], starts)
    order = np.lexsort((h2, h1, lens))
    same = (lens[order][1:] == lens[order][:-1]) & (h1[order][1:] == h1[order][:-1]) & (h2[order][1:] == h2[order][:-1])
    cls_sorted = np.r_[0, np.cumsum(~same)]
    cls = np.empty(n, dtype=np.int64)
    cls[order] = cls_sorted
    rep = np.full(int(cls_sorted[-1]) + 1, -1, dtype=np.int64)
    rep[cls[::-1]] = np.arange(n - 1, -1, -1)  # first group of each class
    # verify: every group equals its representative element-wise
    rep_of = rep[cls]
    src = np.repeat(starts[rep_of], lens) + within
    assert np.array_equal(tids, tids[src]), "template-list hash collision"
    # values[]: [count, ids...] per class, classes laid out in order of their representative
    cls_order = np.argsort(rep, kind="stable")
    cls_len = lens[rep[cls_order]]
    cls_off = np.zeros(len(rep), dtype=np.int64)
    cls_off[cls_order] = np.r_[0, np.cumsum(cls_len + 1)[:-1]]
    v_index = int((cls_len + 1).sum())
    vdtype = np.uint16 if DB_size < 65535 else np.uint32
    values = np.zeros(v_index, dtype=vdtype)
    rs = starts[rep[cls_order]]
    o = cls_off[cls_order]
    values[o] = cls_len
    dst = np.repeat(o + 1, cls_len) + (np.arange(int(cls_len.sum())) - np.repeat(np.r_[0, np.cumsum(cls_len)[:-1]], cls_len))
    srci = np.repeat(rs, cls_len) + (np.arange(int(cls_len.sum())) - np.repeat(np.r_[0, np.cumsum(cls_len)[:-1]], cls_len))
    values[dst] = tids[srci]
    value_index = cls_off[cls].astype(np.uint32)

    # open hash: size = power of two >= n starting at initial_size (hashmap.c:188-240)
    size = initial_size
    while n > size:
        size <<= 1
    assert (1 << (2 * k)) > 2 * size, "megamap layout not produced by this writer"
    bucket = keys & np.uint32(size - 1)
    bo = np.argsort(bucket, kind="stable")
    keys_s, vidx_s, bucket_s = keys[bo], value_index[bo], bucket[bo]
    exist = np.full(size, n, dtype=np.uint32)  # null_index = n
    first = np.flatnonzero(np.r_[True, bucket_s[1:] != bucket_s[:-1]])
    exist[bucket_s[first]] = first.astype(np.uint32)
    # terminating key: first key whose bucket differs from the last key's bucket (compress.c:575-579)
    i = 0
    while i < n - 1 and bucket_s[i] == bucket_s[n - 1]:
        i += 1
    key_index = np.r_[keys_s, keys_s[i]].astype(np.uint32)

    with open(prefix + ".comp.b", "wb") as f:
        f.write(np.array([DB_size, k, 0], dtype=np.uint32).tobytes())
        f.write(np.array([0, size, n, v_index, n], dtype=np.uint64).tobytes())
        f.write(exist.tobytes())
        f.write(values.tobytes())
        f.write(key_index.tobytes())
        f.write(vidx_s.astype(np.uint32).tobytes())
        f.write(np.array([k, 0], dtype=np.uint32).tobytes())
    with open(prefix + ".length.b", "wb") as f:
        f.write(np.array([DB_size], dtype=np.int32).tobytes())
        f.write(np.array([k] + [len(s) for s in seqs], dtype=np.int32).tobytes())
    with open(prefix + ".seq.b", "wb") as f:
        for s in seqs:
            w, _ = pack_2bit(np.asarray(s, dtype=np.uint8))
            nw = (len(s) >> 5) + 1
            f.write(np.r_[w, np.zeros(nw - len(w), dtype=np.uint64)].astype(np.uint64).tobytes())
    with open(prefix + ".name", "w") as f:
        for nm in names:
            f.write(nm + "\n")
    return dict(DB_size=DB_size, n=n, size=size, v_index=v_index, lists=len(rep))
