"""Test helpers: oracle bindings (TEST INFRASTRUCTURE), golden fixtures, reference binary wrapper."""
from __future__ import annotations

import contextlib
import ctypes as C
import gzip
import os
import shutil
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_KMA = os.path.join(ROOT, "oracle", "_ref", "kma")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libkma_ref.so")

_orc = None


def orc():
    global _orc
    if _orc is None:
        L = C.CDLL(os.path.join(ROOT, "oracle", "liborc.so"))
        L.orc_db_open.restype = C.c_void_p
        L.orc_db_open.argtypes = [C.c_char_p]
        L.orc_db_close.argtypes = [C.c_void_p]
        L.orc_lookup.restype = C.c_int64
        L.orc_lookup.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_seed_stream.restype = C.c_int64
        L.orc_seed_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_chain_stream.restype = C.c_int64
        L.orc_chain_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_double,
                                       C.c_double, C.c_void_p, C.c_size_t, C.c_void_p]
        _orc = L
    return _orc


def oracle_params(exhaustive=0, apm=0):
    p = (C.c_int32 * 40)()
    orc().orc_default_params(p)
    p[32] = exhaustive  # M MM U W1 Wl Mn PE (7) + d[25] -> exhaustive at index 32
    p[33] = apm         # 0 = -apm p, 1 = -apm u (the default pairing)
    return p


def oracle_seed_stream(db_prefix: str, s1: np.ndarray, exhaustive=0, stats=None, apm=0, proxi=1.0) -> np.ndarray:
    L = orc()
    L.orc_set_proxi.argtypes = [C.c_double]
    L.orc_set_proxi(float(proxi))   # -proxi (kma.c:702-718); 1.0 = off
    db = L.orc_db_open(os.fsencode(db_prefix))
    assert db, f"oracle cannot open {db_prefix}"
    cap = 3 * len(s1) + 4096
    while True:
        out = np.zeros(cap, dtype=np.uint8)
        st = (C.c_int64 * 7)()
        n = L.orc_seed_stream(db, oracle_params(exhaustive, apm), s1.ctypes.data, len(s1), out.ctypes.data, len(out), st)
        if n >= 0 or cap > (1 << 32):
            break
        cap *= 8
    L.orc_db_close(db)
    assert n >= 0
    if stats is not None:
        stats.update(dict(zip(["reads", "mapped", "read_words", "lookups", "hits", "list_fetches", "list_ids"], list(st))))
    return out[:n].copy()


def oracle_chain_stream(db_prefix: str, s1: np.ndarray, exhaustive=0, minlen=16, mrs=0.5, coverT=0.1, mrc=0.0, stats=None, lc=0, proxi=1.0) -> np.ndarray:
    """Stage 2 in chain mode (save_kmers_chain, the default without -1t1); CLI defaults kma.c:309-320. lc = -lc (kma.c:694)."""
    L = orc()
    L.orc_set_proxi.argtypes = [C.c_double]
    L.orc_set_proxi(float(proxi))
    L.orc_chain_set_lc(int(lc))
    db = L.orc_db_open(os.fsencode(db_prefix))
    assert db, f"oracle cannot open {db_prefix}"
    cap = 8 * len(s1) + 4096
    while True:
        out = np.zeros(cap, dtype=np.uint8)
        st = (C.c_int64 * 7)()
        n = L.orc_chain_stream(db, oracle_params(exhaustive), s1.ctypes.data, len(s1), minlen, mrs, coverT, mrc,
                               out.ctypes.data, len(out), st)
        if n != -1 or cap > (1 << 32):
            break
        cap *= 8
    L.orc_db_close(db)
    assert n >= 0, f"oracle chain error {n}"
    if stats is not None:
        stats.update(dict(zip(["reads", "mapped", "read_words", "lookups", "hits", "list_fetches", "list_ids"], list(st))))
    return out[:n].copy()


class soft_proxi:
    """context: the oracle's stage 2 adds the soft proximity sums (kmers.c:133-153) into .sums [DB_size] while it is open"""
    def __init__(self, db_prefix):
        self.sums = np.zeros(int(np.fromfile(db_prefix + ".length.b", dtype=np.int32, count=1)[0]), dtype=np.uint64)

    def __enter__(self):
        orc().orc_set_soft_proxi.argtypes = [C.c_void_p]
        orc().orc_set_soft_proxi(self.sums.ctypes.data)
        return self

    def __exit__(self, *a):
        orc().orc_set_soft_proxi(None)

    def trailer(self) -> bytes:
        """what save_kmers_batch appends to the stream: the first 24 bytes of the sums, then all of them (kmers.c:151-153)"""
        b = self.sums.tobytes()
        return b[:24] + b


def have_ref() -> bool:
    return os.path.exists(REF_KMA)


def ref_kma(args, cwd=None, stdout=None):
    """Run the unmodified reference binary (oracle/_ref/kma)."""
    r = subprocess.run([REF_KMA] + list(args), cwd=cwd, stdout=stdout or subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    return r.stdout


def ref_index(fasta: str, prefix: str):
    ref_kma(["index", "-i", fasta, "-o", prefix])


@contextlib.contextmanager
def golden_dir():
    """tests/golden/*.gz unpacked into a temp dir (the .comp.b is 4 MiB of mostly null slots)."""
    d = tempfile.mkdtemp(prefix="kma_golden_")
    try:
        for f in os.listdir(GOLDEN):
            src = os.path.join(GOLDEN, f)
            if f.endswith(".gz"):
                with gzip.open(src, "rb") as fi, open(os.path.join(d, f[:-3]), "wb") as fo:
                    shutil.copyfileobj(fi, fo)
            elif not f.endswith(".py"):
                shutil.copy(src, os.path.join(d, f))
        yield d
    finally:
        shutil.rmtree(d, ignore_errors=True)


REF_ALN = os.path.join(ROOT, "oracle", "_ref", "ref_aln")


def ref_align(db_prefix: str, s2: bytes, tmp: str, one2one=True, cand=True, pe=False, min_frac=1.0):
    """Ground truth of the alignment pass from the unmodified reference (oracle/ref_harness.c):
    (frag_raw bytes, alignment_scores, uniq_alignment_scores, cand rows [n, 8] or None)."""
    p = os.path.join(tmp, "s2.bin")
    with open(p, "wb") as f:
        f.write(s2)
    args = [REF_ALN, db_prefix, p, os.path.join(tmp, "fr.out"), os.path.join(tmp, "sc.out")]
    if cand:
        args.append(os.path.join(tmp, "cand.out"))
    if one2one:
        args.append("-1t1")
    if pe:
        args.append("-apm-u" if pe == "u" else "-apm-p")
    if min_frac != 1.0:
        args += ["-mf", repr(float(min_frac))]
    r = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    frag = open(os.path.join(tmp, "fr.out"), "rb").read()
    sc = np.fromfile(os.path.join(tmp, "sc.out"), dtype=np.uint8)
    n = int(sc[:4].view(np.int32)[0])
    arr = sc[4:].view(np.uint64)
    c = np.fromfile(os.path.join(tmp, "cand.out"), dtype=np.int32).reshape(-1, 8) if cand else None
    return frag, arr[:n].copy(), arr[n:2 * n].copy(), c


def oracle_align_stream(db_prefix: str, s2: np.ndarray, one2one=True, want_cand=True, apm=0, mq=0, min_frac=1.0):
    """(frag_raw bytes, alignment_scores, uniq_alignment_scores, cand rows, nw cells) from the C oracle."""
    L = orc()
    L.orc_align_set_minfrac.argtypes = [C.c_double]
    L.orc_align_set_minfrac(float(min_frac))
    L.orc_align_stream.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_int,
                                   C.c_int, C.c_double, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_void_p, C.c_void_p,
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_int64)]
    L.orc_free.argtypes = [C.c_void_p]
    db = L.orc_db_open(os.fsencode(db_prefix))
    assert db
    DB = int(np.fromfile(db_prefix + ".length.b", dtype=np.int32, count=1)[0])
    a = np.zeros(DB, dtype=np.uint64)
    u = np.zeros(DB, dtype=np.uint64)
    fo, fb, co, cr, cells = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t(), C.c_int64()
    s2 = np.ascontiguousarray(s2, dtype=np.uint8)
    rc = L.orc_align_stream(db, os.fsencode(db_prefix), oracle_params(apm=apm), s2.ctypes.data, len(s2), int(one2one), 0.5, int(mq), 16, 0.0,
                            C.byref(fo), C.byref(fb), a.ctypes.data, u.ctypes.data,
                            C.byref(co) if want_cand else None, C.byref(cr), C.byref(cells))
    assert rc == 0
    frag = C.string_at(fo, fb.value) if fb.value else b""
    cand = None
    if want_cand:
        cand = np.frombuffer(C.string_at(co, cr.value * 32), dtype=np.int32).reshape(-1, 8).copy() if cr.value else np.zeros((0, 8), np.int32)
        L.orc_free(co)
    L.orc_free(fo)
    L.orc_db_close(db)
    return frag, a, u, cand, cells.value


def cand_equal(a, b):
    """candidate rows equal; the reference leaves `match` undefined where nothing aligned"""
    if a.shape != b.shape:
        return False
    ok = a == b
    ok[:, 5] |= (b[:, 2] == 0)
    return bool(ok.all())


def parse_frag_raw(buf: bytes):
    """frag_raw stream (updatescores.c:284-295, single-end records) -> list of dicts"""
    recs, p = [], 0
    while p + 20 <= len(buf):
        q_len, n, score, hl, flag = np.frombuffer(buf, dtype=np.int32, count=5, offset=p)
        p += 20
        n = abs(int(n))
        q = np.frombuffer(buf, dtype=np.uint8, count=int(q_len), offset=p); p += int(q_len)
        hdr = buf[p:p + int(hl)]; p += int(hl)
        arr = np.frombuffer(buf[p:p + 12 * n], dtype=np.int32).reshape(3, n); p += 12 * n
        if score < 0:   # a mate block follows (update_Scores_pe)
            q2, hl2, fl2 = np.frombuffer(buf, dtype=np.int32, count=3, offset=p)
            p += 12 + int(q2) + int(hl2)
        recs.append(dict(q=q, hdr=hdr, score=int(score), flag=int(flag), start=arr[0], end=arr[1], tmpl=arr[2]))
    return recs


def assembly_records(frag_raw: bytes, zero_every=5, max_hits=2) -> np.ndarray:
    """per-template fragment records as assemble_KMA reads them (frags.c:45-48): int32[8]{template, q_len, nHits,
    score, start, end, hdrlen, flag} + read bytes + header. Reads chosen on the reverse strand are reverse-complemented
    (what ConClave does before it files them); every `zero_every`-th record gets score 0 so that anker_rc decides."""
    comp = np.array([3, 2, 1, 0, 4, 5], dtype=np.uint8)
    out = bytearray()
    for i, r in enumerate(parse_frag_raw(frag_raw)):
        for j in range(min(max_hits, len(r["tmpl"]))):
            t = int(r["tmpl"][j])
            q = r["q"] if t > 0 else comp[r["q"][::-1]]
            score = 0 if (zero_every and i % zero_every == 0) else abs(r["score"])
            if score == 0 and i % (2 * zero_every) == 0:   # wrong way round: anker_rc has to turn the read
                q = comp[q[::-1]]
            out += np.array([abs(t), len(q), len(r["tmpl"]), score, int(r["start"][j]), int(r["end"][j]), len(r["hdr"]), r["flag"]],
                            dtype=np.int32).tobytes()
            out += q.tobytes() + r["hdr"]
    return np.frombuffer(bytes(out), dtype=np.uint8)


def ref_trace(db_prefix: str, frags: np.ndarray, tmp: str, one2one=True, matrix=None, ts=0):
    """ground truth of the traceback alignment (assemble_KMA's anker_rc + KMA) from the unmodified reference.
    matrix = "sparse" | "dense": also run the reference's alnToMat / alnToMatDense on every accepted alignment and
    return (trace bytes, {template: uint16 counts[t_len, 6] of the template nodes}, {template: total nodes})."""
    p = os.path.join(tmp, "frags.bin")
    frags.tofile(p)
    args = [REF_ALN, "-trace", db_prefix, p, os.path.join(tmp, "trace.out")] + (["-1t1"] if one2one else [])
    if ts:
        args += ["-ts", str(int(ts))]   # trimSeeds (chain.c:496), kma.c:571
    if matrix:
        args += ["-mat", os.path.join(tmp, "mat.out")] + (["-dense"] if matrix == "dense" else [])
    r = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    trace = open(os.path.join(tmp, "trace.out"), "rb").read()
    if not matrix:
        return trace
    buf = open(os.path.join(tmp, "mat.out"), "rb").read()
    mats, nodes, o = {}, {}, 0
    while o + 12 <= len(buf):
        t, tl, nn = np.frombuffer(buf, dtype=np.int32, count=3, offset=o)
        o += 12
        mats[int(t)] = np.frombuffer(buf, dtype=np.uint16, count=6 * int(tl), offset=o).reshape(int(tl), 6).copy()
        nodes[int(t)] = int(nn)
        o += 12 * int(tl)
    return trace, mats, nodes


def matrix_offsets(db_prefix: str) -> np.ndarray:
    """offset (in template positions) of template t in the all-template count matrix: sum(len[1..t-1]); entry DB_size = total"""
    lengths = np.fromfile(db_prefix + ".length.b", dtype=np.int32)[1:].astype(np.int64)
    off = np.zeros(len(lengths) + 1, dtype=np.int64)
    off[2:] = np.cumsum(lengths[1:])
    return off


def oracle_matrix(db_prefix: str, frags: np.ndarray, trace: bytes, dense=False, saturate=True) -> np.ndarray:
    """uint16 [total template positions, 6] from the oracle's restatement of alnToMat (template nodes) / alnToMatDense"""
    L = orc()
    raw = np.fromfile(db_prefix + ".length.b", dtype=np.int32)
    DB, lengths = int(raw[0]), np.ascontiguousarray(raw[1:])
    off = matrix_offsets(db_prefix)
    counts = np.zeros((int(off[DB]), 6), dtype=np.uint16)
    frags = np.ascontiguousarray(frags, dtype=np.uint8)
    tr = np.frombuffer(trace, dtype=np.uint8)
    L.orc_matrix_stream.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    n = L.orc_matrix_stream(lengths.ctypes.data, DB, frags.ctypes.data, len(frags), tr.ctypes.data, len(tr), int(dense), counts.ctypes.data)
    assert n >= 0
    if not saturate:
        assert counts.max() < 65535, "per-rank counts for the all-reduce must not have saturated"
    return counts


def oracle_trace(db_prefix: str, frags: np.ndarray, one2one=True, ts=0) -> bytes:
    L = orc()
    L.orc_trace_set_ts(int(ts))
    L.orc_trace_stream.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_int,
                                   C.c_int, C.c_double, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.orc_free.argtypes = [C.c_void_p]
    db = L.orc_db_open(os.fsencode(db_prefix))
    assert db
    frags = np.ascontiguousarray(frags, dtype=np.uint8)
    o, ob = C.c_void_p(), C.c_size_t()
    rc = L.orc_trace_stream(db, os.fsencode(db_prefix), oracle_params(), frags.ctypes.data, len(frags), int(one2one), 0.5, 0, 16, 0.0,
                            C.byref(o), C.byref(ob))
    assert rc == 0
    out = C.string_at(o, ob.value) if ob.value else b""
    L.orc_free(o)
    L.orc_db_close(db)
    return out


def parse_trace(buf: bytes):
    """trace output -> list of (int32[12] header, t, s, q rows)"""
    out, p = [], 0
    while p + 48 <= len(buf):
        h = np.frombuffer(buf, dtype=np.int32, count=12, offset=p).copy(); p += 48
        n = int(h[11])
        rows = [buf[p + i * n:p + (i + 1) * n] for i in range(3)]
        p += 3 * n
        out.append((h, rows))
    return out


def ref_conclave(db_prefix: str, frag_raw: bytes, a: np.ndarray, u: np.ndarray, tmp: str, max_frag=None, lc=False, c2=None, and_mode=False):
    """ground truth of ConClave's choice pass + printFrags from the unmodified reference (ref_harness -conclave):
    (list of per-file byte strings, w_scores u64[DB], fragmentCounts u32[DB], readCounts u32[DB])"""
    fp, sp, op = os.path.join(tmp, "cc_frag.bin"), os.path.join(tmp, "cc_sc.bin"), os.path.join(tmp, "cc_out.bin")
    open(fp, "wb").write(frag_raw)
    with open(sp, "wb") as f:
        f.write(np.array([len(a)], dtype=np.int32).tobytes() + a.astype(np.uint64).tobytes() + u.astype(np.uint64).tobytes())
    args = [REF_ALN, "-conclave", db_prefix, fp, sp, op] + ([str(max_frag)] if max_frag else [])
    if c2:   # (scoreT, evalue): runConClave2 (-ConClave 2); the unique scores it updates come back as a fifth value
        args += ["-c2", repr(float(c2[0])), repr(float(c2[1]))] + (["-and"] if and_mode else [])
    args += ["-lc"] if lc else []
    r = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    buf = open(op, "rb").read()
    nf = int(np.frombuffer(buf, dtype=np.int32, count=1)[0])
    o, files = 4, []
    for _ in range(nf):
        nb = int(np.frombuffer(buf, dtype=np.int64, count=1, offset=o)[0])
        files.append(buf[o + 8:o + 8 + nb])
        o += 8 + nb
    DB = len(a)
    w = np.frombuffer(buf, dtype=np.uint64, count=DB, offset=o).copy(); o += 8 * DB
    fc = np.frombuffer(buf, dtype=np.uint32, count=DB, offset=o).copy(); o += 4 * DB
    rc = np.frombuffer(buf, dtype=np.uint32, count=DB, offset=o).copy(); o += 4 * DB
    if c2:
        return files, w, fc, rc, np.frombuffer(buf, dtype=np.uint64, count=DB, offset=o).copy()
    return files, w, fc, rc


def oracle_conclave(db_prefix: str, frag_raw: bytes, a: np.ndarray, u: np.ndarray, lc=False):
    """(per-template fragment records of one file, w_scores, fragmentCounts, readCounts) from the C oracle"""
    L = orc()
    raw = np.fromfile(db_prefix + ".length.b", dtype=np.int32)
    DB, lengths = int(raw[0]), np.ascontiguousarray(raw[1:])
    fr = np.frombuffer(frag_raw, dtype=np.uint8)
    out = np.zeros(2 * len(fr) + 4096, dtype=np.uint8)
    w, fc, rc = np.zeros(DB, np.uint64), np.zeros(DB, np.uint32), np.zeros(DB, np.uint32)
    a, u = np.ascontiguousarray(a, dtype=np.uint64), np.ascontiguousarray(u, dtype=np.uint64)
    L.orc_conclave_stream.restype = C.c_int64
    L.orc_conclave_stream.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_conclave_set_lc(int(lc))
    n = L.orc_conclave_stream(lengths.ctypes.data, DB, fr.ctypes.data, len(fr), a.ctypes.data, u.ctypes.data, out.ctypes.data, len(out),
                              w.ctypes.data, fc.ctypes.data, rc.ctypes.data)
    L.orc_conclave_set_lc(0)
    assert n >= 0, f"oracle conclave error {n}"
    return out[:n].tobytes(), w, fc, rc


def ref_p_chisqr():
    """the reference's own p_chisqr (stdstat.c:136) as a C function pointer: double (*)(long double)"""
    ref = C.CDLL(REF_SO)
    return C.cast(ref.p_chisqr, C.c_void_p)


def oracle_conclave2(db_prefix: str, frag_raw: bytes, a: np.ndarray, u: np.ndarray, scoreT=0.5, evalue=0.05, lc=False, and_mode=False):
    """runConClave2 (-ConClave 2) from the C oracle over the reference's p_chisqr:
    (fragment records of the one file, w_scores, fragmentCounts, readCounts, updated unique scores)"""
    L = orc()
    raw = np.fromfile(db_prefix + ".length.b", dtype=np.int32)
    DB, lengths = int(raw[0]), np.ascontiguousarray(raw[1:])
    fr = np.frombuffer(frag_raw, dtype=np.uint8)
    out = np.zeros(2 * len(fr) + 4096, dtype=np.uint8)
    w, fc, rc = np.zeros(DB, np.uint64), np.zeros(DB, np.uint32), np.zeros(DB, np.uint32)
    a, u = np.ascontiguousarray(a, dtype=np.uint64), np.array(u, dtype=np.uint64)
    L.orc_conclave2_stream.restype = C.c_int64
    L.orc_conclave2_stream.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p]
    L.orc_conclave_set_lc(int(lc))
    n = L.orc_conclave2_stream(lengths.ctypes.data, DB, fr.ctypes.data, len(fr), a.ctypes.data, u.ctypes.data, out.ctypes.data, len(out),
                               w.ctypes.data, fc.ctypes.data, rc.ctypes.data, float(scoreT), float(evalue), int(and_mode), ref_p_chisqr())
    L.orc_conclave_set_lc(0)
    assert n >= 0, f"oracle conclave2 error {n}"
    return out[:n].tobytes(), w, fc, rc, u


def ref_memscore(db_prefix: str, s2: bytes, tmp: str):
    """-mem_mode: the k-mer score collection of runKMA_MEM from the unmodified reference (ref_harness -memscore):
    (frag_raw bytes, alignment_scores, uniq_alignment_scores)"""
    p = os.path.join(tmp, "ms_s2.bin")
    open(p, "wb").write(s2)
    args = [REF_ALN, "-memscore", db_prefix, p, os.path.join(tmp, "ms_fr.out"), os.path.join(tmp, "ms_sc.out")]
    r = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    frag = open(os.path.join(tmp, "ms_fr.out"), "rb").read()
    sc = np.fromfile(os.path.join(tmp, "ms_sc.out"), dtype=np.uint8)
    n = int(sc[:4].view(np.int32)[0])
    arr = sc[4:].view(np.uint64)
    return frag, arr[:n].copy(), arr[n:2 * n].copy()


def oracle_memscore(db_prefix: str, s2) -> tuple:
    L = orc()
    raw = np.fromfile(db_prefix + ".length.b", dtype=np.int32)
    DB, lengths = int(raw[0]), np.ascontiguousarray(raw[1:])
    s2 = np.ascontiguousarray(np.frombuffer(s2, dtype=np.uint8) if isinstance(s2, (bytes, bytearray)) else s2, dtype=np.uint8)
    a, u = np.zeros(DB, np.uint64), np.zeros(DB, np.uint64)
    fo, fb = C.c_void_p(), C.c_size_t()
    L.orc_memscore_stream.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_void_p, C.c_void_p]
    L.orc_free.argtypes = [C.c_void_p]
    rc = L.orc_memscore_stream(lengths.ctypes.data, DB, s2.ctypes.data, len(s2), C.byref(fo), C.byref(fb), a.ctypes.data, u.ctypes.data)
    assert rc == 0
    frag = C.string_at(fo, fb.value) if fb.value else b""
    L.orc_free(fo)
    return frag, a, u


# ---------------------------------------------------------------- consensus call of the assembly pass

def template_bases(db_prefix: str, t: int) -> np.ndarray:
    """bases 0-3 of template t from .seq.b / .length.b"""
    raw = np.fromfile(db_prefix + ".length.b", dtype=np.int32)
    lengths = raw[1:]
    off = 0
    for i in range(1, t):
        off += (int(lengths[i]) >> 5) + 1
    words = np.fromfile(db_prefix + ".seq.b", dtype=np.uint64, count=(int(lengths[t]) >> 5) + 1, offset=8 * off)
    pos = np.arange(int(lengths[t]))
    return ((words[pos >> 5] << ((pos & 31).astype(np.uint64) << np.uint64(1))) >> np.uint64(62)).astype(np.uint8)


def random_count_matrix(rng, tb: np.ndarray) -> np.ndarray:
    """uint16 [t_len, 6] base counts around template bases tb with every regime callConsensus distinguishes: no depth,
    depth below -bcd, ties, a best base short of half the depth, gap majorities, borderline significance, saturation."""
    n = len(tb)
    m = np.zeros((n, 6), dtype=np.int64)
    regime = rng.integers(0, 10, size=n)
    depth = np.select([regime == 0, regime <= 2, regime <= 6, regime <= 8], [0, rng.integers(1, 6, size=n), rng.integers(4, 80, size=n),
                      rng.integers(80, 6000, size=n)], default=rng.integers(60000, 70000, size=n))
    for i in range(n):
        d = int(depth[i])
        if d == 0:
            continue
        kind = rng.integers(0, 8)
        if kind <= 2:      # clean majority for the template base
            p = np.full(6, 0.01); p[tb[i]] = 1.0
        elif kind == 3:    # a variant base with errors
            p = np.full(6, 0.03); p[(tb[i] + 1 + rng.integers(0, 3)) % 4] = 1.0
        elif kind == 4:    # two competing bases
            p = np.full(6, 0.02); a, b = rng.choice(6, size=2, replace=False); p[a] = 1.0; p[b] = rng.choice([1.0, 0.8, 0.5])
        elif kind == 5:    # deletion majority
            p = np.full(6, 0.05); p[5] = 1.0; p[tb[i]] = rng.choice([0.0, 0.3, 0.9])
        elif kind == 6:    # flat
            p = np.ones(6)
        else:              # N-rich
            p = np.full(6, 0.1); p[4] = 1.0
        m[i] = rng.multinomial(d, p / p.sum())
        if kind == 4 and rng.integers(0, 2):
            m[i, b] = m[i, a]   # exact tie
    return np.minimum(m, 65535).astype(np.uint16)


def ref_consensus(db_prefix: str, mats: dict, tmp: str, bcd=1, evalue=0.05, caller=0, sig=0, support=0.0):
    """ground truth: the reference's own callConsensus (ref_harness -consensus). mats: {template: uint16 [t_len, 6]}.
    Returns {template: (t, s, q bytes, uint64 stats[5] = depth, depthVar, len, aln_len, cover)}."""
    p = os.path.join(tmp, "cmat.bin")
    with open(p, "wb") as f:
        for t, m in mats.items():
            f.write(np.array([t, len(m), len(m)], dtype=np.int32).tobytes())
            f.write(np.ascontiguousarray(m, dtype=np.uint16).tobytes())
    o = os.path.join(tmp, "cons.out")
    r = subprocess.run([REF_ALN, "-consensus", db_prefix, p, o, "-bcd", str(bcd), "-evalue", repr(evalue), "-caller", str(caller),
                        "-sig", str(sig), "-support", repr(support)], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    buf, out, q = open(o, "rb").read(), {}, 0
    while q + 48 <= len(buf):
        t, tl = (int(x) for x in np.frombuffer(buf, dtype=np.int32, count=2, offset=q))
        st = np.frombuffer(buf, dtype=np.uint64, count=5, offset=q + 8).copy()
        q += 48
        out[t] = (buf[q:q + tl], buf[q + tl:q + 2 * tl], buf[q + 2 * tl:q + 3 * tl], st)
        q += 3 * tl
    return out


def oracle_consensus(db_prefix: str, t: int, counts: np.ndarray, bcd=1, evalue=0.05, caller=0, sig=0, support=0.0):
    L = orc()
    L.orc_consensus.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p]
    raw = np.fromfile(db_prefix + ".length.b", dtype=np.int32)
    lengths = raw[1:]
    off = sum((int(lengths[i]) >> 5) + 1 for i in range(1, t))
    tl = int(lengths[t])
    assert counts.shape == (tl, 6)
    words = np.fromfile(db_prefix + ".seq.b", dtype=np.uint64, count=(tl >> 5) + 1, offset=8 * off)
    counts = np.ascontiguousarray(counts, dtype=np.uint16)
    ts, ss, qs = (np.zeros(tl, dtype=np.uint8) for _ in range(3))
    st = np.zeros(5, dtype=np.uint64)
    rc = L.orc_consensus(counts.ctypes.data, words.ctypes.data, tl, bcd, caller, sig, support, evalue, ts.ctypes.data, ss.ctypes.data,
                         qs.ctypes.data, st.ctypes.data)
    assert rc == 0
    return ts.tobytes(), ss.tobytes(), qs.tobytes(), st


def oracle_chi2_min(evalue: float) -> float:
    L = orc()
    L.orc_chi2_min.restype = C.c_double
    L.orc_chi2_min.argtypes = [C.c_double]
    return float(L.orc_chi2_min(evalue))


# ---------------------------------------------------------------- stage 1 (FASTQ / FASTA text -> stage-1 records)

def fastq_text(reads, quals=None, prefix="r", fasta=False, crlf=False, desc=True) -> bytes:
    """4-line FASTQ (or 2-line FASTA) text of reads given as codes 0-4; quals: list of uint8 phred+offset arrays"""
    nl = b"\r\n" if crlf else b"\n"
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    out = bytearray()
    for i, r in enumerate(reads):
        seq = lut[np.asarray(r, dtype=np.uint8)].tobytes()
        if fasta:
            out += b">" + f"{prefix}{i} some description".encode() + nl + seq + nl
        else:
            q = np.asarray(quals[i], dtype=np.uint8).tobytes() if quals is not None else b"I" * len(seq)
            out += b"@" + f"{prefix}{i}".encode() + (b" 1:N:0" if desc and i % 3 == 0 else b"") + nl + seq + nl + b"+" + nl + q + nl
    return bytes(out)


def random_quals(rng, reads, scale=33):
    """phred qualities with decaying ends, so that the -mp end trim cuts something off most reads"""
    out = []
    for r in reads:
        n = len(r)
        q = rng.integers(22, 41, size=n)
        a, b = int(rng.integers(0, 12)), int(rng.integers(0, 25))
        if a and a < n:
            q[:a] = rng.integers(2, 24, size=a)
        if b and b < n:
            q[n - b:] = rng.integers(2, 24, size=b)
        if rng.integers(0, 25) == 0:
            q[:] = rng.integers(2, 19, size=n)   # nothing survives
        if n and rng.integers(0, 6) == 0:
            q[0] = 31                            # '@' as the first quality character (phred 33)
        out.append((q + scale).astype(np.uint8))
    return out


def oracle_stage1(text1: bytes, text2: bytes | None = None, fastq=True, min_phred=20, phred_scale=33, minlen=16, maxlen=2147483647,
                  min_q=0, hardmask_q=0):
    L = orc()
    from kma_b200 import api as _api
    prob = _api.quality_prob()
    L.orc_stage1_set_quality.argtypes = [C.c_int, C.c_int, C.c_void_p]
    L.orc_stage1_set_quality(int(min_q), int(hardmask_q), prob.ctypes.data)   # -eq / -mi; prob must outlive the call below
    L.orc_stage1.restype = C.c_size_t
    L.orc_stage1.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                             C.c_size_t, C.POINTER(C.c_int64)]
    cap = len(text1) + (len(text2) if text2 else 0) + 4096
    out = np.zeros(cap, dtype=np.uint8)
    cnt = C.c_int64()
    n = L.orc_stage1(text1, len(text1), text2, len(text2) if text2 else 0, int(fastq), min_phred, phred_scale, minlen, maxlen,
                     out.ctypes.data, cap, C.byref(cnt))
    assert n <= cap
    return out[:n].tobytes(), cnt.value


def oracle_fasta_unwrap(text: bytes) -> bytes:
    L = orc()
    L.orc_fasta_unwrap.restype = C.c_size_t
    L.orc_fasta_unwrap.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
    out = np.empty(len(text) + 1, dtype=np.uint8)
    return out[: L.orc_fasta_unwrap(text, len(text), out.ctypes.data)].tobytes()


def build_db(tmp_path, names, seqs) -> str:
    """a database in the reference's format made by kma_b200.dbbuild (no reference binary needed)"""
    from kma_b200 import dbbuild
    prefix = os.path.join(str(tmp_path), "db")
    dbbuild.build_db(prefix, names, seqs)
    return prefix
