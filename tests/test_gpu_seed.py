"""GPU parity tests (B200): the CUDA stage-2 path through the C ABI vs the oracle / golden stream."""
import numpy as np
import pytest

from kma_b200 import api, synth, records
from tests import util

pytestmark = pytest.mark.gpu


def _gpu_stream(prefix, s1, exhaustive=0):
    db = api.TemplateDB(prefix, device=0)
    p = api.default_params()
    p.exhaustive = exhaustive
    out, n, st = db.save_kmers_batch(s1, p)
    db.close()
    return out.tobytes() + api.stream_terminator(n), st


def test_golden_stage2_stream():
    with util.golden_dir() as g:
        s1 = np.fromfile(f"{g}/s1.bin", dtype=np.uint8)
        s2 = np.fromfile(f"{g}/s2.bin", dtype=np.uint8)
        got, st = _gpu_stream(f"{g}/db", s1)
    assert st.launches > 0
    assert got == s2.tobytes()


def test_lookup_matches_oracle():
    with util.golden_dir() as g:
        L = util.orc()
        odb = L.orc_db_open(f"{g}/db".encode())
        db = api.TemplateDB(f"{g}/db")
        rng = np.random.default_rng(3)
        names_seqs = open(f"{g}/db.fsa").read().split("\n")
        tr = np.full(256, 0, dtype=np.uint64)
        for i, c in enumerate(b"ACGT"):
            tr[c] = i
        s = tr[np.frombuffer(names_seqs[1].encode(), dtype=np.uint8)]
        kmers = [int(sum(int(s[j + i]) << (2 * (15 - i)) for i in range(16))) for j in range(0, len(s) - 16, 3)]
        kmers += [int(x) for x in rng.integers(0, 1 << 32, size=2000)]
        kmers = np.array(kmers, dtype=np.uint64)
        got = db.lookup(kmers)
        want = np.array([L.orc_lookup(odb, int(k)) for k in kmers], dtype=np.int64)
        L.orc_db_close(odb)
        db.close()
    assert (want >= 0).sum() > 10
    assert np.array_equal(got, want)


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,exhaustive", [(11, 0), (12, 1)])
def test_ragged_reads_vs_oracle(tmp_path, seed, exhaustive):
    names, seqs = synth.gene_db(seed, n_families=30, n_variants=8, len_lo=200, len_hi=1500)
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    reads = []
    for L in (0, 8, 15, 16, 17, 31, 32, 33, 64, 100, 151, 250, 271, 272, 300, 511, 700, 1200):
        if L == 0:
            continue
        pool = [s for s in seqs if len(s) >= L] or seqs
        L2 = min(L, min(len(s) for s in pool))
        reads += list(synth.short_reads(seed * 100 + L, pool, 80, L=L2, sub=0.02,
                                        n_rate=0.004 if L > 20 else 0.0, junk_frac=0.1))
    reads.append(np.full(40, 4, dtype=np.uint8))
    reads.append(np.zeros(50, dtype=np.uint8))
    reads += list(synth.long_reads(seed, seqs, 6, len_lo=2000, len_hi=6000))
    s1 = records.stage1_records(reads)
    want = util.oracle_seed_stream(str(tmp_path / "db"), s1, exhaustive=exhaustive)
    got, st = _gpu_stream(str(tmp_path / "db"), s1, exhaustive)
    assert got == want.tobytes()


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
def test_many_templates_takes_dense_path(tmp_path):
    """A k-mer rich family (300 near-identical variants) overflows the shared table -> dense scratch path."""
    names, seqs = synth.gene_db(21, n_families=2, n_variants=300, len_lo=400, len_hi=500)
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    reads = synth.short_reads(22, seqs, 400, L=150, sub=0.01)
    s1 = records.stage1_records_fixed(reads)
    want = util.oracle_seed_stream(str(tmp_path / "db"), s1)
    got, st = _gpu_stream(str(tmp_path / "db"), s1)
    assert st.overflow_reads > 0
    assert got == want.tobytes()


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
def test_full_size_properties(tmp_path):
    """C1-sized batch (100k x 150 bp): equality with the oracle plus size-independent properties."""
    names, seqs = synth.gene_db(42)
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    reads = synth.short_reads(7, seqs, 100000, n_rate=0.0005, junk_frac=0.02)
    s1 = records.stage1_records_fixed(reads)
    got, st = _gpu_stream(str(tmp_path / "db"), s1)
    want = util.oracle_seed_stream(str(tmp_path / "db"), s1)
    assert got == want.tobytes()
    recs = records.parse_stage2(np.frombuffer(got, dtype=np.uint8))
    assert len(recs) == st.mapped
    # idempotence: same batch again -> same bytes
    got2, _ = _gpu_stream(str(tmp_path / "db"), s1)
    assert got2 == got
    # strand symmetry: reverse-complementing every read leaves scores and |template sets| unchanged
    rc = np.where(reads == 4, 4, 3 - reads)[:, ::-1].copy()
    rc[reads[:, ::-1] == 4] = 4
    got_rc, _ = _gpu_stream(str(tmp_path / "db"), records.stage1_records_fixed(rc))
    recs_rc = records.parse_stage2(np.frombuffer(got_rc, dtype=np.uint8))
    assert [abs(r["score"]) for r in recs] == [abs(r["score"]) for r in recs_rc]
    assert [sorted(abs(t) for t in r["templates"]) for r in recs] == [sorted(abs(t) for t in r["templates"]) for r in recs_rc]


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", [51, 52])
def test_paired_end_stage2_vs_oracle(tmp_path, seed):
    """-ipe ... -apm p: get_kmers_for_pair + save_kmers_penaltyPair + printPair on the GPU, byte-exact"""
    from tests.test_oracle_pair import make_pairs
    prefix, s1, s2 = make_pairs(tmp_path, seed, n=3000)
    want = util.oracle_seed_stream(prefix, s1)
    assert want.tobytes() == s2
    got, st = _gpu_stream(prefix, s1)
    assert got == s2


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
def test_paired_end_many_templates_dense_path(tmp_path):
    """pairs from a 300-variant family: the per-strand template lists overflow the shared table -> dense path"""
    names, seqs = synth.gene_db(23, n_families=2, n_variants=300, len_lo=500, len_hi=600)
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    r1, r2 = synth.paired_reads(24, seqs, 300, sub=0.01, ins_lo=200, ins_hi=400)
    s1 = records.stage1_pairs(r1, r2)
    want = util.oracle_seed_stream(str(tmp_path / "db"), s1)
    got, st = _gpu_stream(str(tmp_path / "db"), s1)
    assert st.overflow_reads > 0
    assert got == want.tobytes()


def test_database_with_more_than_65535_templates(tmp_path):
    """DB_size >= 65535 switches the template lists of .comp.b from uint16 to uint32 (hashmapkma.c:339-348): stage 2,
    the alignment pass and chain mode against the oracle on a database of 66 000 short templates"""
    from kma_b200 import dbbuild
    rng = np.random.default_rng(77)
    base = [rng.integers(0, 4, size=120).astype(np.uint8) for _ in range(6600)]
    seqs, names = [], []
    for f, b in enumerate(base):
        for v in range(10):
            s = b.copy()
            if v:
                s[rng.integers(0, 120, size=2)] = rng.integers(0, 4, size=2)
            seqs.append(s)
            names.append(f"t{f}_{v}")
    prefix = str(tmp_path / "db")
    dbbuild.build_db(prefix, names, seqs)
    reads = synth.short_reads(78, seqs, 3000, L=100, sub=0.01)
    s1 = records.stage1_records_fast(reads)
    want = util.oracle_seed_stream(prefix, s1)
    db = api.TemplateDB(prefix, device=0)
    assert db.info.DB_size > 65535
    got, n, st = db.save_kmers_batch(s1)
    assert got.tobytes() + api.stream_terminator(n) == want.tobytes()
    ofrag, oa, ou, _, _ = util.oracle_align_stream(prefix, want, want_cand=False)
    frag, a, u, _, _ = db.alnFrags_batch(want)
    assert frag.tobytes() == ofrag and np.array_equal(a, oa) and np.array_equal(u, ou)
    p = api.default_params()
    p.kmerscan = 1
    cwant = util.oracle_chain_stream(prefix, s1)
    cgot, cn, _ = db.save_kmers_batch(s1, p)
    db.close()
    assert cgot.tobytes() + api.stream_terminator(cn) == cwant.tobytes()
