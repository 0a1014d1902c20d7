"""CPU tests: the oracle restatement of ConClave's choice pass + per-template bucketing (runConClave, conclave.c:43;
printFrags, frags.c:30) is pinned byte for byte to the reference's own functions run by oracle/ref_harness.c -conclave."""
import numpy as np
import pytest

from kma_b200 import synth
from tests import util

pytestmark = pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")


def se_case(tmp_path, seed, n=1500, L=150, chain=False):
    names, seqs = synth.gene_db(seed, n_families=10, n_variants=8, len_lo=300, len_hi=1200)
    # a few templates stored on the other strand so that reads are chosen reverse-complemented
    for i in range(0, len(seqs), 7):
        names.append(names[i] + "_rc")
        seqs.append(synth.revcomp(seqs[i]))
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    rng = np.random.default_rng(seed)
    reads = list(synth.short_reads(seed + 1, seqs, n, L=L, sub=0.01, junk_frac=0.02))
    if chain:
        reads += synth.long_reads(seed + 2, seqs, 60, len_lo=1000, len_hi=4000, err=0.06)
    synth.write_fastq(tmp_path / "r.fq", reads)
    flags = [] if chain else ["-1t1"]
    s2 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-s2"] + flags, cwd=tmp_path)
    frag, a, u, _ = util.ref_align(str(tmp_path / "db"), s2, str(tmp_path), one2one=not chain, cand=False)
    return str(tmp_path / "db"), frag, a, u


@pytest.mark.parametrize("seed,chain", [(71, False), (72, True)])
def test_conclave_single_end(tmp_path, seed, chain):
    prefix, frag, a, u = se_case(tmp_path, seed, chain=chain)
    files, w, fc, rc = util.ref_conclave(prefix, frag, a, u, str(tmp_path))
    got, ow, ofc, orc_ = util.oracle_conclave(prefix, frag, a, u)
    assert len(files) == 1 and len(files[0]) > 10000
    assert got == files[0]
    assert np.array_equal(ow, w) and np.array_equal(ofc, fc) and np.array_equal(orc_, rc)
    assert (w > 0).sum() > 5
    recs = np.frombuffer(got[:32], dtype=np.int32)
    assert recs[0] > 0


def test_conclave_paired_end(tmp_path):
    names, seqs = synth.gene_db(73, n_families=10, n_variants=8, len_lo=500, len_hi=1500)
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    r1, r2 = synth.paired_reads(74, seqs, 1200, sub=0.01)
    synth.write_fastq(tmp_path / "a.fq", r1)
    synth.write_fastq(tmp_path / "b.fq", r2)
    s2 = util.ref_kma(["-ipe", "a.fq", "b.fq", "-o", "o", "-t_db", "db", "-apm", "p", "-s2"], cwd=tmp_path)
    prefix = str(tmp_path / "db")
    frag, a, u, _ = util.ref_align(prefix, s2, str(tmp_path), one2one=False, cand=False, pe=True)
    files, w, fc, rc = util.ref_conclave(prefix, frag, a, u, str(tmp_path))
    got, ow, ofc, orc_ = util.oracle_conclave(prefix, frag, a, u)
    assert got == files[0]
    assert np.array_equal(ow, w) and np.array_equal(ofc, fc) and np.array_equal(orc_, rc)
    assert int(rc.sum()) > int(fc.sum())          # pairs count two reads per fragment


def test_conclave_chunks_of_maxfrag(tmp_path):
    """the reference cuts a new file every maxFrag fragments: each file is one oracle call on that slice of records"""
    from kma_b200 import api
    prefix, frag, a, u = se_case(tmp_path, 75, n=900)
    files, w, fc, rc = util.ref_conclave(prefix, frag, a, u, str(tmp_path), max_frag=400)
    assert len(files) == 3
    off = api.record_offsets(4, np.frombuffer(frag, dtype=np.uint8))
    tot_w = np.zeros_like(w)
    for i, f in enumerate(files):
        lo, hi = int(off[min(400 * i, len(off) - 1)]), int(off[min(400 * (i + 1), len(off) - 1)])
        got, ow, _, _ = util.oracle_conclave(prefix, frag[lo:hi], a, u)
        assert got == f
        tot_w += ow
    assert np.array_equal(tot_w, w)


@pytest.mark.parametrize("seed,chain", [(75, False), (76, True)])
def test_conclave_length_corrected(tmp_path, seed, chain):
    """-lc: runConClave_lc (conclave.c:215-384) orders the candidates by score per template base before the total"""
    prefix, frag, a, u = se_case(tmp_path, seed, chain=chain)
    files, w, fc, rc = util.ref_conclave(prefix, frag, a, u, str(tmp_path), lc=True)
    got, ow, ofc, orc_ = util.oracle_conclave(prefix, frag, a, u, lc=True)
    assert got == files[0] and np.array_equal(ow, w) and np.array_equal(ofc, fc) and np.array_equal(orc_, rc)
    # arbitrary global score arrays (any values are valid input) make the two key orders disagree
    rng = np.random.default_rng(seed)
    a2 = rng.integers(1, 50000, size=len(a)).astype(np.uint64)
    u2 = rng.integers(0, 3, size=len(u)).astype(np.uint64)
    files2, w2, fc2, rc2 = util.ref_conclave(prefix, frag, a2, u2, str(tmp_path), lc=True)
    got2, ow2, ofc2, orc2 = util.oracle_conclave(prefix, frag, a2, u2, lc=True)
    assert got2 == files2[0] and np.array_equal(ow2, w2) and np.array_equal(ofc2, fc2) and np.array_equal(orc2, rc2)
    # a hand-made record with two candidates whose total and per-base orders disagree: the short template wins under -lc
    lengths = np.fromfile(prefix + ".length.b", dtype=np.int32)[1:]
    t_short, t_long = int(np.argmin(lengths[1:])) + 1, int(np.argmax(lengths[1:])) + 1
    a3 = np.zeros(len(a), dtype=np.uint64)
    a3[t_short], a3[t_long] = 1000, 1001
    assert 1000 / lengths[t_short] > 1001 / lengths[t_long]
    read = np.arange(40, dtype=np.uint8) % 4
    rec = (np.array([40, 2, 30, 3, 0], dtype=np.int32).tobytes() + read.tobytes() + b"x1\0" +
           np.array([0, 0, 40, 40, t_short, t_long], dtype=np.int32).tobytes())
    u3 = np.zeros(len(a), dtype=np.uint64)
    for lc, want_t in ((False, t_long), (True, t_short)):
        files3, w3, _, _ = util.ref_conclave(prefix, rec, a3, u3, str(tmp_path), lc=lc)
        got3, ow3, _, _ = util.oracle_conclave(prefix, rec, a3, u3, lc=lc)
        assert got3 == files3[0] and np.array_equal(ow3, w3)
        assert int(np.frombuffer(got3[:4], dtype=np.int32)[0]) == want_t


def _cc2_case(tmp_path, seed, kind):
    if kind == "pe":
        names, seqs = synth.gene_db(seed, n_families=10, n_variants=8, len_lo=500, len_hi=1500)
        synth.write_fasta(tmp_path / "db.fsa", names, seqs)
        util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
        r1, r2 = synth.paired_reads(seed + 1, seqs, 1500, sub=0.01)
        synth.write_fastq(tmp_path / "a.fq", r1)
        synth.write_fastq(tmp_path / "b.fq", r2)
        s2 = util.ref_kma(["-ipe", "a.fq", "b.fq", "-o", "o", "-t_db", "db", "-apm", "p", "-s2"], cwd=tmp_path)
        prefix = str(tmp_path / "db")
        frag, a, u, _ = util.ref_align(prefix, s2, str(tmp_path), one2one=False, cand=False, pe=True)
        return prefix, frag, a, u
    return se_case(tmp_path, seed, n=2500, chain=kind == "chain")


@pytest.mark.parametrize("seed,kind,scoreT,evalue,lc,and_mode", [(81, "se", 0.5, 0.05, False, False), (82, "chain", 0.25, 0.05, False, False),
                                                                 (83, "pe", 0.5, 0.05, False, False), (84, "se", 0.5, 1e-6, True, False),
                                                                 (85, "se", 2.0, 0.05, False, True), (86, "chain", 0.5, 0.5, True, True)])
def test_conclave_version_2(tmp_path, seed, kind, scoreT, evalue, lc, and_mode):
    """-ConClave 2 (runConClave2 / _lc, conclave.c:386-1110): provisional sums, significance filter, unique-score update,
    weighted random draw with the 4-key fallback -- fragments, sums, counts and the updated unique scores vs the reference"""
    prefix, frag, a, u = _cc2_case(tmp_path, seed, kind)
    files, w, fc, rc, u2 = util.ref_conclave(prefix, frag, a, u, str(tmp_path), lc=lc, c2=(scoreT, evalue), and_mode=and_mode)
    got, ow, ofc, orc_, ou = util.oracle_conclave2(prefix, frag, a, u, scoreT=scoreT, evalue=evalue, lc=lc, and_mode=and_mode)
    one, _, _, _ = util.oracle_conclave(prefix, frag, a, u, lc=lc)
    assert len(files) == 1 and got == files[0]
    assert np.array_equal(ow, w) and np.array_equal(ofc, fc) and np.array_equal(orc_, rc) and np.array_equal(ou, u2)
    assert got != one, "the case is meant to make version 2 choose differently from version 1"
