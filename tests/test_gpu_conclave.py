"""GPU parity tests (B200): ConClave's choice pass + per-template bucketing through the C ABI vs the oracle that
tests/test_oracle_conclave.py pins to the reference's own runConClave / printFrags; and the whole of stage 3 on the
device: alignment pass -> ConClave -> traceback alignment + base counts."""
import numpy as np
import pytest

from kma_b200 import api, synth
from tests import util
from tests.test_oracle_conclave import se_case

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")]


@pytest.mark.parametrize("seed,chain", [(71, False), (72, True)])
def test_conclave_single_end(tmp_path, seed, chain):
    prefix, frag, a, u = se_case(tmp_path, seed, chain=chain)
    want, ow, ofc, orc_ = util.oracle_conclave(prefix, frag, a, u)
    db = api.TemplateDB(prefix, device=0)
    got, w, fc, rc, n = db.conclave_batch(frag, a, u)
    db.close()
    assert n > 1000
    assert got.tobytes() == want
    assert np.array_equal(w, ow) and np.array_equal(fc, ofc) and np.array_equal(rc, orc_)


def test_conclave_length_corrected(tmp_path):
    """-lc (runConClave_lc): per-base score before the total, real and arbitrary score arrays"""
    prefix, frag, a, u = se_case(tmp_path, 77, chain=True)
    rng = np.random.default_rng(77)
    a2 = rng.integers(1, 50000, size=len(a)).astype(np.uint64)
    u2 = rng.integers(0, 3, size=len(u)).astype(np.uint64)
    db = api.TemplateDB(prefix, device=0)
    db.conclave_mode(True)
    for aa, uu in ((a, u), (a2, u2)):
        want, ow, ofc, orc_ = util.oracle_conclave(prefix, frag, aa, uu, lc=True)
        got, w, fc, rc, _ = db.conclave_batch(frag, aa, uu)
        assert got.tobytes() == want and np.array_equal(w, ow) and np.array_equal(fc, ofc) and np.array_equal(rc, orc_)
    db.conclave_mode(False)
    want, ow, _, _ = util.oracle_conclave(prefix, frag, a2, u2)
    got, w, _, _, _ = db.conclave_batch(frag, a2, u2)
    db.close()
    assert got.tobytes() == want and np.array_equal(w, ow)


def test_conclave_paired_end_and_empty(tmp_path):
    names, seqs = synth.gene_db(73, n_families=10, n_variants=8, len_lo=500, len_hi=1500)
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    r1, r2 = synth.paired_reads(74, seqs, 1200, sub=0.01)
    synth.write_fastq(tmp_path / "a.fq", r1)
    synth.write_fastq(tmp_path / "b.fq", r2)
    s2 = util.ref_kma(["-ipe", "a.fq", "b.fq", "-o", "o", "-t_db", "db", "-apm", "p", "-s2"], cwd=tmp_path)
    prefix = str(tmp_path / "db")
    frag, a, u, _ = util.ref_align(prefix, s2, str(tmp_path), one2one=False, cand=False, pe=True)
    want, ow, ofc, orc_ = util.oracle_conclave(prefix, frag, a, u)
    db = api.TemplateDB(prefix, device=0)
    got, w, fc, rc, n = db.conclave_batch(frag, a, u)
    assert got.tobytes() == want
    assert np.array_equal(w, ow) and np.array_equal(fc, ofc) and np.array_equal(rc, orc_)
    empty, w0, _, _, n0 = db.conclave_batch(b"", a, u)
    db.close()
    assert n0 == 0 and empty.tobytes() == np.array([-1], dtype=np.int32).tobytes() and not w0.any()


def test_resident_pair_stream_through_conclave(tmp_path):
    """paired reads: the alignment pass leaves pair records with mate blocks and pairs resolved as two single records
    in one slot; ConClave on that resident stream == ConClave on the downloaded one == the oracle"""
    names, seqs = synth.gene_db(75, n_families=10, n_variants=8, len_lo=500, len_hi=1500)
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    r1, r2 = synth.paired_reads(76, seqs, 1500, sub=0.02)
    r2 = [r if i % 5 else synth.revcomp(np.asarray(r)) for i, r in enumerate(r2)]   # some pairs that are not proper pairs
    synth.write_fastq(tmp_path / "a.fq", r1)
    synth.write_fastq(tmp_path / "b.fq", r2)
    prefix = str(tmp_path / "db")
    for apm in ("p", "u"):
        s2 = np.frombuffer(util.ref_kma(["-ipe", "a.fq", "b.fq", "-o", "o", "-t_db", "db", "-apm", apm, "-s2"], cwd=tmp_path), dtype=np.uint8)
        db = api.TemplateDB(prefix, device=0)
        p = api.default_params()
        p.apm = 1 if apm == "u" else 0
        frag, a, u, _, _ = db.alnFrags_batch(s2, p)
        want, ow, ofc, orc_ = util.oracle_conclave(prefix, frag.tobytes(), a, u)
        got, w, fc, rc, _ = db.conclave_batch(frag, a, u)
        assert got.tobytes() == want
        db.align_upload(s2); db.align_run(p)
        g2, w2, fc2, rc2, _ = db.conclave_resident(a, u, cap=len(want) + 4096, source="align")
        db.close()
        assert g2.tobytes() == want and np.array_equal(w2, ow) and np.array_equal(fc2, ofc) and np.array_equal(rc2, orc_), apm
        rec = api.record_offsets(4, np.frombuffer(frag.tobytes(), dtype=np.uint8))
        assert len(rec) - 1 > 1200


def test_stage3_on_the_device(tmp_path):
    """alignment pass -> (score arrays) -> ConClave -> traceback alignment + base counts, every step on the GPU, equal to
    the oracle chain (each link of which is pinned to the reference)"""
    names, seqs = synth.gene_db(81, n_families=10, n_variants=8, len_lo=300, len_hi=1200)
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    rng = np.random.default_rng(81)
    reads = [synth.mutate_indel(rng, r, 0.02, 0.01, 0.01) for r in synth.short_reads(82, seqs, 1500, L=150, sub=0.0, junk_frac=0.02)]
    synth.write_fastq(tmp_path / "r.fq", reads)
    s2 = np.frombuffer(util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-1t1", "-s2"], cwd=tmp_path), dtype=np.uint8)
    prefix = str(tmp_path / "db")
    ofrag, oa, ou, _, _ = util.oracle_align_stream(prefix, s2, want_cand=False)
    ofrags, ow, _, _ = util.oracle_conclave(prefix, ofrag, oa, ou)
    otrace = util.oracle_trace(prefix, np.frombuffer(ofrags, dtype=np.uint8))
    omat = util.oracle_matrix(prefix, np.frombuffer(ofrags, dtype=np.uint8), otrace)
    db = api.TemplateDB(prefix, device=0)
    p = api.default_params()
    p.one2one = 1
    p.matrix = 1
    frag, a, u, _, _ = db.alnFrags_batch(s2, p)
    frags, w, _, _, _ = db.conclave_batch(frag, a, u)
    db.matrix_reset()
    trace, n, _ = db.assemble_align_batch(frags, p)
    mat = db.matrix_download()
    # the same with the fragments staying in HBM between ConClave and the traceback pass, with and without row output
    none, w2, _, _, _ = db.conclave_batch(frag, a, u, download=False)
    db.matrix_reset()
    trace2, n2, _ = db.trace_from_conclave(p)
    mat2 = db.matrix_download()
    db.matrix_reset()
    none2, n3, st3 = db.trace_from_conclave(p, download=False)
    mat3 = db.matrix_download()
    db.matrix_reset()
    none3, n4, _ = db.assemble_align_batch(frags, p, download=False)
    mat4 = db.matrix_download()
    # ... and with the frag_raw stream staying in HBM between the alignment pass and ConClave as well
    db.align_upload(s2); db.align_run(p)
    g5, w5, fc5, rc5, _ = db.conclave_resident(a, u, cap=len(frags) + 4096, source="align")
    assert g5.tobytes() == ofrags and np.array_equal(w5, ow)
    db.close()
    assert none is None and none2 is None and none3 is None and n2 == n3 == n4 == n and np.array_equal(w2, w)
    assert trace2.tobytes() == trace.tobytes() and np.array_equal(mat2, mat) and np.array_equal(mat3, mat) and np.array_equal(mat4, mat)
    assert frags.tobytes() == ofrags and np.array_equal(w, ow)
    assert trace.tobytes() == otrace
    assert np.array_equal(mat, omat) and int(mat.sum()) > 100000


def test_memscore_vs_oracle(tmp_path):
    """-mem_mode score collection (update_Scores_MEM / _pe_MEM): single reads with N's and strand ties, then pairs"""
    from tests.test_oracle_memscore import _db
    prefix, seqs = _db(tmp_path, 91)
    reads = list(synth.short_reads(92, seqs, 1200, L=150, sub=0.01, junk_frac=0.03, n_rate=0.002))
    reads += [np.zeros(12, dtype=np.uint8)]
    for i in range(0, 40, 2):
        r = reads[i]
        reads[i] = np.concatenate([r[:75], synth.revcomp(r[:75])])
    synth.write_fastq(tmp_path / "r.fq", reads)
    s2 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-1t1", "-s2"], cwd=tmp_path)
    r1, r2 = synth.paired_reads(94, seqs, 800, sub=0.01)
    synth.write_fastq(tmp_path / "a.fq", r1)
    synth.write_fastq(tmp_path / "b.fq", r2)
    s2pe = util.ref_kma(["-ipe", "a.fq", "b.fq", "-o", "o", "-t_db", "db", "-apm", "p", "-s2"], cwd=tmp_path)
    db = api.TemplateDB(prefix, device=0)
    for stream in (s2, s2pe):
        ofrag, oa, ou = util.oracle_memscore(prefix, stream)
        frag, a, u, n = db.memscore_batch(stream)
        assert n > 500 and frag.tobytes() == ofrag
        assert np.array_equal(a, oa) and np.array_equal(u, ou)
    db.close()


def test_mem_mode_flow_on_one_genome(tmp_path):
    """C4-shaped: one genome as the only template, reads through stage 2 (-1t1), -mem_mode score collection, ConClave,
    traceback alignment and base counts -- every step on the GPU, equal to the oracle chain"""
    from kma_b200 import records
    rng = np.random.default_rng(5)
    genome = rng.integers(0, 4, size=200_000).astype(np.uint8)
    synth.write_fasta(tmp_path / "db.fsa", ["genome"], [genome])
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    prefix = str(tmp_path / "db")
    reads = [synth.mutate_indel(rng, r, 0.01, 0.003, 0.003) for r in synth.short_reads(6, [genome], 3000, L=150, sub=0.0)]
    s1 = records.stage1_records(reads)
    os2 = util.oracle_seed_stream(prefix, s1)
    ofrag, oa, ou = util.oracle_memscore(prefix, os2)
    ofrags, ow, _, _ = util.oracle_conclave(prefix, ofrag, oa, ou)
    otrace = util.oracle_trace(prefix, np.frombuffer(ofrags, dtype=np.uint8))
    omat = util.oracle_matrix(prefix, np.frombuffer(ofrags, dtype=np.uint8), otrace)
    db = api.TemplateDB(prefix, device=0)
    p = api.default_params()
    p.one2one = 1
    p.matrix = 1
    s2, nreads, _ = db.save_kmers_batch(s1, p)
    s2 = s2.tobytes() + api.stream_terminator(nreads)
    assert s2 == os2.tobytes()
    frag, a, u, _ = db.memscore_batch(s2)
    frags, w, _, _, _ = db.conclave_batch(frag, a, u)
    db.matrix_reset()
    trace, n, _ = db.assemble_align_batch(frags, p)
    mat = db.matrix_download(1)
    ct, cs, cq, cst, _ = db.consensus(1)   # the consensus of the genome from the device's own counts
    # the whole flow again without a record leaving HBM: FASTQ text -> stage 1 -> stage 2 -> score collection -> ConClave
    # -> traceback + base counts -> consensus
    text = util.fastq_text(reads, desc=False)
    _, cnt, _, _, _ = db.run_input_text(text, download=False)
    db.seed_run(p)
    f2, a2, u2, _ = db.memscore_from_seed(download=True, cap=len(frag) + 4096)
    g2, w2, _, _, _ = db.conclave_resident(a2, u2, download=True, cap=len(frags) + 4096)
    db.matrix_reset()
    none, n2, _ = db.trace_from_conclave(p, download=False)
    mat_r = db.matrix_download(1)
    ct2 = db.consensus(1)
    assert cnt == nreads and f2.tobytes() == ofrag and np.array_equal(a2, a) and g2.tobytes() == ofrags and np.array_equal(w2, w)
    assert none is None and n2 == n and np.array_equal(mat_r, mat) and ct2[2].tobytes() == cq.tobytes()
    db.close()
    wt, ws, wq, wst = util.oracle_consensus(prefix, 1, omat)
    assert ct.tobytes() == wt and cs.tobytes() == ws and cq.tobytes() == wq
    assert [int(cst[0][k]) for k in ("depth", "depthVar", "len", "aln_len", "cover")] == [int(x) for x in wst] and int(wst[4]) > 150_000
    assert frag.tobytes() == ofrag and frags.tobytes() == ofrags and int(w[1]) == int(ow[1]) > 0
    assert trace.tobytes() == otrace
    assert np.array_equal(mat, omat) and mat.shape == (200_000, 6)
    depth = mat[:, :4].sum(axis=1)
    assert depth.mean() > 1.5     # 3000 x 150 bp over 200 kb


@pytest.mark.parametrize("seed,kind,scoreT,evalue,lc,and_mode", [(81, "se", 0.5, 0.05, False, False), (82, "chain", 0.25, 0.05, False, False),
                                                                 (83, "pe", 0.5, 0.05, False, False), (84, "se", 0.5, 1e-6, True, False),
                                                                 (85, "se", 2.0, 0.05, False, True), (86, "chain", 0.5, 0.5, True, True)])
def test_conclave_version_2(tmp_path, seed, kind, scoreT, evalue, lc, and_mode):
    """-ConClave 2 (runConClave2 / _lc, conclave.c:386-1110) on the device: provisional sums, the host's significance filter over
    the reference's p_chisqr, unique-score update, weighted random draw with the 4-key fallback -- fragments, sums, counts and
    the updated unique scores vs the oracle that tests/test_oracle_conclave.py pins to the reference's own functions"""
    from tests.test_oracle_conclave import _cc2_case
    prefix, frag, a, u = _cc2_case(tmp_path, seed, kind)
    want, ow, ofc, orc_, ou = util.oracle_conclave2(prefix, frag, a, u, scoreT=scoreT, evalue=evalue, lc=lc, and_mode=and_mode)
    db = api.TemplateDB(prefix, device=0)
    db.conclave_mode(lc)
    db.conclave_version(2, util.ref_p_chisqr(), scoreT=scoreT, evalue=evalue, and_mode=and_mode)
    got, w, fc, rc, n = db.conclave_batch(frag, a, u)
    gu = db.conclave_uniq_scores()
    assert got.tobytes() == want
    assert np.array_equal(w, ow) and np.array_equal(fc, ofc) and np.array_equal(rc, orc_) and np.array_equal(gu, ou)
    db.conclave_version(1)   # back to runConClave
    want1, ow1, _, _ = util.oracle_conclave(prefix, frag, a, u, lc=lc)
    got1, w1, _, _, _ = db.conclave_batch(frag, a, u)
    db.close()
    assert got1.tobytes() == want1 and np.array_equal(w1, ow1) and got1.tobytes() != got.tobytes()
