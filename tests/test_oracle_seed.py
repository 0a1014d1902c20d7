"""CPU tests: the oracle restatement of stage 2 is pinned to the reference (golden + live binary)."""
import os
import re
import subprocess

import numpy as np
import pytest

from kma_b200 import synth, records
from tests import util


def test_oracle_matches_golden_stage2():
    with util.golden_dir() as g:
        s1 = np.fromfile(f"{g}/s1.bin", dtype=np.uint8)
        s2 = np.fromfile(f"{g}/s2.bin", dtype=np.uint8)
        got = util.oracle_seed_stream(f"{g}/db", s1)
    assert got.tobytes() == s2.tobytes()


def test_numpy_stage1_writer_matches_reference_s1():
    """kma_b200.records.stage1_records reproduces `kma -s1` for the golden FASTQ."""
    with util.golden_dir() as g:
        s1 = np.fromfile(f"{g}/s1.bin", dtype=np.uint8)
        reads, names = [], []
        lines = open(f"{g}/reads.fq").read().split("\n")
        tr = np.full(256, 4, dtype=np.uint8)
        for i, c in enumerate(b"ACGT"):
            tr[c] = i
        for i in range(0, len(lines) - 3, 4):
            names.append(lines[i][1:])
            reads.append(tr[np.frombuffer(lines[i + 1].encode(), dtype=np.uint8)])
        mine = records.stage1_records(reads, names=names)
    assert mine.tobytes() == s1.tobytes()


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,exhaustive", [(5, 0), (6, 1)])
def test_oracle_matches_live_reference(tmp_path, seed, exhaustive):
    """Fresh seeded data -> reference `kma -s2` vs oracle, byte for byte (ragged lengths, N's, junk, short reads)."""
    names, seqs = synth.gene_db(seed, n_families=25, n_variants=6, len_lo=200, len_hi=1200)
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    rng = np.random.default_rng(seed)
    reads = []
    for L in (8, 15, 16, 17, 31, 32, 33, 64, 100, 151, 250, 300, 700):
        n = 60
        rr = synth.short_reads(seed * 100 + L, [s for s in seqs if len(s) >= L] or seqs, n, L=L, sub=0.02,
                               n_rate=0.004 if L > 20 else 0.0, junk_frac=0.1)
        reads += list(rr)
    reads.append(np.full(40, 4, dtype=np.uint8))          # all N
    reads.append(np.zeros(50, dtype=np.uint8))            # poly-A
    order = rng.permutation(len(reads))
    reads = [reads[i] for i in order]
    synth.write_fastq(tmp_path / "r.fq", reads)
    util.ref_index("db.fsa", "db") if False else util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    extra = ["-ex_mode"] if exhaustive else []
    s1 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-1t1", "-s1"] + extra, cwd=tmp_path)
    s2 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-1t1", "-s2"] + extra, cwd=tmp_path)
    got = util.oracle_seed_stream(str(tmp_path / "db"), np.frombuffer(s1, dtype=np.uint8), exhaustive=exhaustive)
    assert got.tobytes() == s2


def test_abi_exports_every_declared_symbol():
    """libkmagpu.so loads without a GPU and exports everything include/kmagpu.h declares."""
    import ctypes
    from kma_b200 import api
    L = api.lib()
    hdr = open(os.path.join(util.ROOT, "include", "kmagpu.h")).read()
    names = set(re.findall(r"\b(kmagpu_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 8
    for n in names:
        assert hasattr(L, n), n
    # no device here -> loud failure, not a CPU fallback
    if L.kmagpu_device_count() == 0:
        h = ctypes.c_void_p()
        assert L.kmagpu_db_open(b"/nonexistent", 0, ctypes.byref(h)) != 0
        assert b"no CUDA device" in L.kmagpu_last_error()
