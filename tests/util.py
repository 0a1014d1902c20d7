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
        _orc = L
    return _orc


def oracle_params(exhaustive=0):
    p = (C.c_int32 * 40)()
    orc().orc_default_params(p)
    p[32] = exhaustive  # M MM U W1 Wl Mn PE (7) + d[25] -> exhaustive at index 32
    return p


def oracle_seed_stream(db_prefix: str, s1: np.ndarray, exhaustive=0, stats=None) -> np.ndarray:
    L = orc()
    db = L.orc_db_open(os.fsencode(db_prefix))
    assert db, f"oracle cannot open {db_prefix}"
    cap = 3 * len(s1) + 4096
    while True:
        out = np.zeros(cap, dtype=np.uint8)
        st = (C.c_int64 * 7)()
        n = L.orc_seed_stream(db, oracle_params(exhaustive), s1.ctypes.data, len(s1), out.ctypes.data, len(out), st)
        if n >= 0 or cap > (1 << 32):
            break
        cap *= 8
    L.orc_db_close(db)
    assert n >= 0
    if stats is not None:
        stats.update(dict(zip(["reads", "mapped", "read_words", "lookups", "hits", "list_fetches", "list_ids"], list(st))))
    return out[:n].copy()


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


def ref_align(db_prefix: str, s2: bytes, tmp: str, one2one=True, cand=True, pe=False):
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
        args.append("-apm-p")
    r = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    frag = open(os.path.join(tmp, "fr.out"), "rb").read()
    sc = np.fromfile(os.path.join(tmp, "sc.out"), dtype=np.uint8)
    n = int(sc[:4].view(np.int32)[0])
    arr = sc[4:].view(np.uint64)
    c = np.fromfile(os.path.join(tmp, "cand.out"), dtype=np.int32).reshape(-1, 8) if cand else None
    return frag, arr[:n].copy(), arr[n:2 * n].copy(), c


def oracle_align_stream(db_prefix: str, s2: np.ndarray, one2one=True, want_cand=True):
    """(frag_raw bytes, alignment_scores, uniq_alignment_scores, cand rows, nw cells) from the C oracle."""
    L = orc()
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
    rc = L.orc_align_stream(db, os.fsencode(db_prefix), oracle_params(), s2.ctypes.data, len(s2), int(one2one), 0.5, 0, 16, 0.0,
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
