"""GPU parity tests (B200): the CUDA alignment pass (MEM seeding, chaining, warp-wavefront NW, selection,
update_Scores, frag_raw writer) through the C ABI vs the oracle, which test_oracle_align.py pins to the unmodified
reference."""
import ctypes as C

import numpy as np
import pytest

from kma_b200 import api, synth, records
from tests import util
from tests.test_nw_emu import problem

pytestmark = pytest.mark.gpu


def _first_diff(a, b):
    n = min(len(a), len(b))
    d = np.flatnonzero(np.frombuffer(a[:n], np.uint8) != np.frombuffer(b[:n], np.uint8))
    return int(d[0]) if len(d) else n


def _check_align(prefix, s2, params=None):
    s2 = np.frombuffer(s2, dtype=np.uint8) if isinstance(s2, (bytes, bytearray)) else s2
    ofrag, oa, ou, ocand, cells = util.oracle_align_stream(prefix, s2)
    db = api.TemplateDB(prefix, device=0)
    frag, a, u, cand, st = db.alnFrags_batch(s2, params, want_cand=True)
    db.close()
    assert st.launches > 0 and st.tasks == len(ocand)
    if not util.cand_equal(cand, ocand):
        bad = np.flatnonzero(((cand != ocand) & ~((np.arange(8) == 5) & (ocand[:, 2:3] == 0))).any(axis=1))
        raise AssertionError(f"{len(bad)} of {len(cand)} candidate rows differ; first: got {cand[bad[0]]} want {ocand[bad[0]]}")
    assert np.array_equal(a, oa) and np.array_equal(u, ou)
    fb = frag.tobytes()
    assert fb == ofrag, f"frag_raw differs at byte {_first_diff(fb, ofrag)} of {len(ofrag)} (got {len(fb)})"
    assert st.nw_full_cells + st.nw_band_cells == cells
    return st, cand


@pytest.mark.parametrize("mq", [200, 300, -1])
def test_golden_alignment_pass_with_min_mapq(mq):
    """-mq: chainSeeds' mapQ (chain.c:256) is only evaluated when it is compared; the comparison is unsigned < int
    as in align.c:658, so a negative -mq rejects every chain"""
    with util.golden_dir() as g:
        s2 = np.fromfile(f"{g}/s2.bin", dtype=np.uint8)
        prefix = f"{g}/db"
        ofrag, oa, ou, ocand, _ = util.oracle_align_stream(prefix, s2, mq=mq)
        ofrag0, _, _, ocand0, _ = util.oracle_align_stream(prefix, s2)
        db = api.TemplateDB(prefix, device=0)
        p = api.default_params()
        p.mq = mq
        frag, a, u, cand, st = db.alnFrags_batch(s2, p, want_cand=True)
        db.close()
    assert st.launches > 0 and util.cand_equal(cand, ocand)
    assert np.array_equal(a, oa) and np.array_equal(u, ou) and frag.tobytes() == ofrag
    assert not np.array_equal(ocand, ocand0)   # the threshold does reject chains of the golden batch
    if mq != 200:
        assert len(ofrag) == 0
    else:
        assert 0 < len(ofrag) < len(ofrag0)


def test_golden_alignment_pass():
    with util.golden_dir() as g:
        s2 = np.fromfile(f"{g}/s2.bin", dtype=np.uint8)
        st, cand = _check_align(f"{g}/db", s2)
    assert st.frags > 1000 and st.mems > 0


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,L,sub,indel", [(31, 150, 0.01, 0.0), (32, 150, 0.03, 0.01), (33, 400, 0.05, 0.02),
                                               (34, 1000, 0.04, 0.03), (35, 3000, 0.08, 0.06)])
def test_fresh_data_vs_oracle(tmp_path, seed, L, sub, indel):
    """substitutions + indels + N's + strand ties; long reads exercise the banded NW, long tails and the chainer"""
    names, seqs = synth.gene_db(seed, n_families=12, n_variants=6, len_lo=max(300, L + 50), len_hi=max(1500, 2 * L))
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    rng = np.random.default_rng(seed)
    base = synth.short_reads(seed + 1, seqs, 500, L=L, sub=0.0, n_rate=0.0, junk_frac=0.03)
    reads = [synth.mutate_indel(rng, r, sub, indel / 2, indel / 2) for r in base]
    for r in reads[::9]:
        if len(r) > 40:
            r[rng.integers(0, len(r), size=2)] = 4
    for i in range(0, 60, 2):
        r = reads[i]
        reads[i] = np.concatenate([r[: len(r) // 2], synth.revcomp(r[: len(r) // 2])])
    synth.write_fastq(tmp_path / "r.fq", reads)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    s2 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-1t1", "-s2"], cwd=tmp_path)
    st, cand = _check_align(str(tmp_path / "db"), s2)
    assert len(cand) > 300
    if L >= 400:
        assert st.nw_band_calls > 0


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
def test_repeats_and_small_scratch_overflow(tmp_path):
    """templates with internal repeats (repeated k-mers -> one MEM per occurrence) and a read long enough that its
    tails overflow the per-warp traceback scratch -> large-scratch path"""
    rng = np.random.default_rng(5)
    seqs = []
    for f in range(6):
        unit = rng.integers(0, 4, size=int(rng.integers(40, 120))).astype(np.uint8)
        parts = []
        for _ in range(int(rng.integers(8, 30))):
            parts.append(synth.mutate_subs(rng, unit, 0.01) if rng.random() < 0.6 else rng.integers(0, 4, size=80).astype(np.uint8))
        seqs.append(np.concatenate(parts))
    names = [f"rep{i}" for i in range(len(seqs))]
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    reads = list(synth.short_reads(6, seqs, 300, L=200, sub=0.02, n_rate=0.0, junk_frac=0.0))
    reads += [synth.mutate_indel(rng, s[: min(len(s), 2500)].copy(), 0.05, 0.03, 0.03) for s in seqs]
    synth.write_fastq(tmp_path / "r.fq", reads)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    s2 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-1t1", "-s2"], cwd=tmp_path)
    st, cand = _check_align(str(tmp_path / "db"), s2)
    assert st.mems > st.tasks


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
def test_seed_then_align_resident(tmp_path):
    """stage 2 -> stage 3 without leaving HBM gives the same bytes as the two-call path"""
    names, seqs = synth.gene_db(42, n_families=40, n_variants=8)
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    reads = synth.short_reads(7, seqs, 20000, n_rate=0.0005, junk_frac=0.02)
    s1 = records.stage1_records_fixed(reads)
    db = api.TemplateDB(str(tmp_path / "db"))
    s2, n, _ = db.save_kmers_batch(s1)
    frag_a, a1, u1, _, st1 = db.alnFrags_batch(s2)
    db.seed_upload(s1); db.seed_run()
    assert db.align_from_seed() == n
    st2 = db.align_run()
    frag_b, a2, u2, _ = db.align_download()
    db.close()
    assert frag_a.tobytes() == frag_b.tobytes() and np.array_equal(a1, a2) and np.array_equal(u1, u2)
    ofrag, oa, ou, _, _ = util.oracle_align_stream(str(tmp_path / "db"), s2, want_cand=False)
    assert frag_a.tobytes() == ofrag and np.array_equal(a1, oa) and np.array_equal(u1, ou)
    # conservation: every kept read adds its score once per kept template
    assert int(a1.sum()) >= int(u1.sum()) > 0 and st1.frags == st2.frags


def test_nw_batch_vs_oracle():
    """NW_score / NW_band_score as stand-alone problems against golden templates: all k modes, ragged sizes"""
    L = util.orc()
    pen = util.oracle_params()
    rng = np.random.default_rng(11)
    with util.golden_dir() as g:
        db = api.TemplateDB(f"{g}/db")
        lens = np.fromfile(f"{g}/db.length.b", dtype=np.int32)[1:]
        seqb = np.fromfile(f"{g}/db.seq.b", dtype=np.uint64)
        off = np.concatenate([[0, 0], np.cumsum((lens[1:] >> 5) + 1)])
        probs, qs, qoff = [], [], 0
        for it in range(400):
            t = int(rng.integers(1, len(lens)))
            tl = int(lens[t])
            if it < 250:
                t_len, q_len, band = int(rng.integers(1, min(tl, 260))), int(rng.integers(1, 260)), 0
            else:
                t_len = int(rng.integers(70, min(tl, 900)))
                q_len = max(66, t_len + int(rng.integers(-50, 50)))
                band = abs(t_len - q_len) + 64
                if q_len <= band or t_len <= band:
                    band = 0
            t_s = int(rng.integers(0, tl - t_len + 1))
            # query: the template window with errors, or unrelated
            tw = seqb[off[t]:off[t] + (tl >> 5) + 1]
            tb = np.array([(int(tw[i >> 5]) >> (62 - 2 * (i & 31))) & 3 for i in range(t_s, t_s + t_len)], dtype=np.uint8)
            if rng.random() < 0.85:
                q = synth.mutate_indel(rng, np.resize(tb, q_len + 20), 0.05, 0.03, 0.03)[:q_len]
                if len(q) < q_len:
                    q = np.concatenate([q, rng.integers(0, 4, size=q_len - len(q)).astype(np.uint8)])
            else:
                q = rng.integers(0, 4, size=q_len).astype(np.uint8)
            if rng.random() < 0.2:
                q[rng.integers(0, q_len)] = 4
            k = int(rng.choice([0, -1, -2, 1, 2]))
            probs.append([t, t_s, t_s + t_len, qoff, 0, q_len, k, band])
            qs.append(q)
            qoff += q_len
        probs = np.array(probs, dtype=np.int32)
        qpool = np.concatenate(qs)
        out, status, cells, steps, ms = db.nw_batch(probs, qpool)
        db.close()
        assert (status == 0).all() and cells > 0 and steps > 0
        for i, p in enumerate(probs):
            want = (C.c_int * 6)()
            tw = np.ascontiguousarray(np.concatenate([seqb[off[p[0]]:off[p[0]] + (int(lens[p[0]]) >> 5) + 1], np.zeros(2, np.uint64)]))
            q = np.ascontiguousarray(qpool[p[3]:p[3] + p[5]])
            L.orc_nw(pen, tw.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p), int(p[6]), int(p[1]), int(p[2]), 0,
                     int(p[5]), int(p[7]), want)
            assert list(out[i]) == list(want), (i, list(p), list(out[i]), list(want))


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", [51, 53])
def test_paired_end_alignment_vs_oracle(tmp_path, seed):
    """-apm p pairs through alnFragsPenaltyPE + update_Scores_pe/_se on the GPU: frag_raw bytes and ConClave sums"""
    from tests.test_oracle_pair import make_pairs
    prefix, s1, s2 = make_pairs(tmp_path, seed, n=3000)
    s2 = np.frombuffer(s2, dtype=np.uint8)
    ofrag, oa, ou, ocand, cells = util.oracle_align_stream(prefix, s2, one2one=False)
    db = api.TemplateDB(prefix)
    frag, a, u, cand, st = db.alnFrags_batch(s2, want_cand=True)
    # stage 2 + stage 3 chained in HBM give the same stream
    db.seed_upload(s1); db.seed_run()
    db.align_from_seed(); db.align_run()
    frag2, a2, u2, _ = db.align_download()
    db.close()
    assert st.tasks == len(ocand)
    ok = (cand[:, 1:] == ocand[:, 1:])
    ok[:, 4] |= (ocand[:, 2] == 0)   # `match` is undefined where nothing aligned
    assert ok.all(), f"candidate rows differ, first: {cand[~ok.all(axis=1)][:1]} want {ocand[~ok.all(axis=1)][:1]}"
    assert np.array_equal(a, oa) and np.array_equal(u, ou)
    fb = frag.tobytes()
    assert fb == ofrag, f"frag_raw differs at byte {_first_diff(fb, ofrag)} of {len(ofrag)} (got {len(fb)})"
    assert frag2.tobytes() == ofrag and np.array_equal(a2, oa) and np.array_equal(u2, ou)


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,L,sub,indel", [(31, 150, 0.01, 0.0), (32, 150, 0.03, 0.01), (33, 400, 0.05, 0.02),
                                               (34, 1000, 0.04, 0.03), (35, 3000, 0.08, 0.06)])
def test_traceback_alignment_vs_oracle(tmp_path, seed, L, sub, indel):
    """assemble_KMA's anker_rc + KMA with aligned rows on the GPU: headers and t/s/q rows byte-exact"""
    from tests.test_oracle_trace import make_frags
    prefix, frags = make_frags(tmp_path, seed, L, sub, indel)
    want = util.oracle_trace(prefix, frags)
    db = api.TemplateDB(prefix)
    got, n, st = db.assemble_align_batch(frags)
    db.close()
    W, G = util.parse_trace(want), util.parse_trace(got.tobytes())
    assert n == len(W) == len(G)
    for i, ((hw, rw), (hg, rg)) in enumerate(zip(W, G)):
        assert np.array_equal(hw, hg), (i, hw, hg)
        assert rw == rg, (i, hw)
    assert got.tobytes() == want
    assert st.nw_full_cells + st.nw_band_cells > 0


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,L,sub,indel,ts", [(32, 150, 0.03, 0.01, 2), (33, 400, 0.05, 0.02, 2), (34, 1000, 0.04, 0.03, 3), (35, 3000, 0.08, 0.06, 2)])
def test_traceback_alignment_with_seed_trimming(tmp_path, seed, L, sub, indel, ts):
    """-ts (trimSeeds, chain.c:496-538; the -ont / -ill / -asm presets set 2) vs the reference's own KMA: a chain that starts
    at MEM 0 is trimmed like any other (only next == 0 ends the walk)"""
    from tests.test_oracle_trace import make_frags
    prefix, frags = make_frags(tmp_path, seed, L, sub, indel)
    want = util.ref_trace(prefix, frags, str(tmp_path), ts=ts)
    assert util.oracle_trace(prefix, frags, ts=ts) == want and want != util.ref_trace(prefix, frags, str(tmp_path))
    db = api.TemplateDB(prefix)
    p = api.default_params()
    p.ts = ts
    got, n, st = db.assemble_align_batch(frags, p)
    db.close()
    assert got.tobytes() == want


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
def test_chunked_pipeline_equals_single_call(tmp_path):
    """MapPipeline (host threads, one library handle + stream each, chunks in turn) == one resident call; pairs are
    never split across chunks"""
    from kma_b200 import pipeline
    from tests.test_oracle_pair import make_pairs
    prefix, s1, s2 = make_pairs(tmp_path, 52, n=3000)
    db = api.TemplateDB(prefix)
    db.seed_upload(s1); db.seed_run(); db.align_from_seed(); db.align_run()
    frag, a, u, _ = db.align_download()
    db.close()
    pipe = pipeline.MapPipeline(prefix, workers=2)
    bounds = pipe.chunk_bounds(s1, 5)
    assert len(bounds) == 5 and bounds[0][0] == 0
    outs = [np.empty(len(frag) + 4096, dtype=np.uint8) for _ in bounds]
    scores = (np.zeros_like(a), np.zeros_like(u))
    res = pipe.map(np.ascontiguousarray(s1), bounds, outs, scores)
    pipe.close()
    got = b"".join(f.tobytes() for f, _ in res)
    assert got == frag.tobytes()
    assert np.array_equal(scores[0], a) and np.array_equal(scores[1], u)
    assert sum(n for _, n in res) == 3000


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
def test_nanopore_like_long_reads(tmp_path):
    """C3-shaped input: 5-20 kb reads with 10 % errors (1/3 sub, del, ins) built from templates and spacers, mapped with
    -1t1 records: hundreds of MEMs per pair, wide banded NW, large traceback matrices (large-scratch path) -- alignment
    pass and traceback alignment byte-exact vs the oracle"""
    names, seqs = synth.gene_db(61, n_families=15, n_variants=6, len_lo=800, len_hi=3000)
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    reads = synth.long_reads(62, seqs, 120, len_lo=5000, len_hi=20000)
    synth.write_fastq(tmp_path / "r.fq", reads, qual="5")
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    s2 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-1t1", "-s2"], cwd=tmp_path)
    prefix = str(tmp_path / "db")
    st, cand = _check_align(prefix, s2)
    assert st.nw_band_calls > 0 and st.mems > 10 * st.tasks
    ofrag, _, _, _, _ = util.oracle_align_stream(prefix, np.frombuffer(s2, dtype=np.uint8), want_cand=False)
    frags = util.assembly_records(ofrag, zero_every=4, max_hits=1)
    want = util.oracle_trace(prefix, frags)
    db = api.TemplateDB(prefix)
    got, n, st2 = db.assemble_align_batch(frags)
    db.close()
    assert got.tobytes() == want
    assert st2.nw_band_cells > 0


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("k", [10, 12, 14])
def test_small_k_databases(tmp_path, k):
    """`kma index -k K`: K = 10 gives the direct-addressed table (megaMap_getGlobal), every K a different k for the
    per-template position index; stage 2, alignment pass and traceback alignment vs the oracle"""
    names, seqs = synth.gene_db(71, n_families=10, n_variants=5, len_lo=400, len_hi=1200)
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    reads = synth.short_reads(72, seqs, 600, L=150, sub=0.02, n_rate=0.002, junk_frac=0.05)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db", "-k", str(k)], cwd=tmp_path)
    prefix = str(tmp_path / "db")
    s1 = records.stage1_records_fixed(reads)
    want2 = util.oracle_seed_stream(prefix, s1)
    db = api.TemplateDB(prefix)
    assert db.info.kmersize == k and db.info.kmerindex == k and db.info.mega == (1 if k == 10 else 0)
    s2, n, _ = db.save_kmers_batch(s1)
    db.close()
    assert s2.tobytes() + api.stream_terminator(n) == want2.tobytes()
    _check_align(prefix, want2)
    ofrag, _, _, _, _ = util.oracle_align_stream(prefix, want2, want_cand=False)
    frags = util.assembly_records(ofrag)
    db = api.TemplateDB(prefix)
    got, _, _ = db.assemble_align_batch(frags)
    db.close()
    assert got.tobytes() == util.oracle_trace(prefix, frags)


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
def test_single_genome_template(tmp_path):
    """C4-shaped database: one 400 kb random genome (one big position-index table, long tails clipped to read + 64)"""
    rng = np.random.default_rng(81)
    genome = rng.integers(0, 4, size=400_000).astype(np.uint8)
    genome[100_000:100_300] = genome[250_000:250_300]          # a repeat: k-mers with two positions
    synth.write_fasta(tmp_path / "db.fsa", ["genome"], [genome])
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    reads = synth.short_reads(82, [genome], 3000, L=150, sub=0.01, n_rate=0.001, junk_frac=0.02)
    reads = list(reads) + [synth.mutate_subs(rng, genome[99_900:100_500].copy(), 0.01), synth.revcomp(genome[249_950:250_400])]
    prefix = str(tmp_path / "db")
    s1 = records.stage1_records(reads)
    want2 = util.oracle_seed_stream(prefix, s1)
    db = api.TemplateDB(prefix)
    s2, n, _ = db.save_kmers_batch(s1)
    db.close()
    assert s2.tobytes() + api.stream_terminator(n) == want2.tobytes()
    st, cand = _check_align(prefix, want2)
    assert st.frags > 2800


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,L,sub,indel,mode", [(41, 150, 0.02, 0.02, 1), (42, 150, 0.02, 0.02, 2), (43, 1000, 0.04, 0.04, 1),
                                                    (44, 1000, 0.04, 0.04, 2)])
def test_base_count_matrix_vs_oracle(tmp_path, seed, L, sub, indel, mode):
    """alnToMat (template nodes) / alnToMatDense counts accumulated on the device by the traceback pass, in two
    batches (the matrix is HBM resident between calls), against the oracle that tests/test_oracle_trace.py pins to
    the reference's own functions; per-template and whole-database downloads; the zero-copy device view for NCCL"""
    from tests.test_oracle_trace import make_frags
    prefix, frags = make_frags(tmp_path, seed, L, sub, indel, n=900)
    trace = util.oracle_trace(prefix, frags)
    want = util.oracle_matrix(prefix, frags, trace, dense=mode == 2)
    db = api.TemplateDB(prefix, device=0)
    p = api.default_params()
    p.one2one = 1
    p.matrix = mode
    db.matrix_reset()
    off = api.record_offsets(3, frags)
    cut = int(off[len(off) // 2])
    got1, n1, _ = db.assemble_align_batch(frags[:cut], p)
    got2, n2, _ = db.assemble_align_batch(frags[cut:], p)
    assert got1.tobytes() + got2.tobytes() == trace
    got = db.matrix_download()
    assert got.shape == want.shape and want.sum() > 1000
    assert np.array_equal(got, want)
    moff = util.matrix_offsets(prefix)
    t = next(t for t in range(1, len(moff) - 1) if want[moff[t]:moff[t + 1]].any())
    one = db.matrix_download(t)
    assert np.array_equal(one, want[moff[t]:moff[t + 1]])
    dev = db.matrix_tensor()
    assert int(dev.sum().item()) == int(want.astype(np.int64).sum())
    from kma_b200 import dist
    assert np.array_equal(dist.allreduce_matrix(dev), want)    # world size 1: clamp only
    db.matrix_reset()
    assert int(db.matrix_download().sum()) == 0
    db.close()


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", [61, 63])
def test_union_pairing_vs_reference(tmp_path, seed):
    """-apm u, the reference's default pairing: save_kmers_unionPair (stage 2, byte-exact vs `kma -ipe -apm u -s2`) and
    alnFragsUnionPE (alignment pass, vs the oracle pinned to alnFrags_threaded), from the stream and chained in HBM"""
    from tests.test_oracle_pair import make_pairs
    prefix, s1, s2 = make_pairs(tmp_path, seed, n=3000, apm="u")
    db = api.TemplateDB(prefix)
    p = api.default_params()
    p.apm = 1
    got, n, st = db.save_kmers_batch(s1, p)
    assert got.tobytes() + api.stream_terminator(n) == s2
    s2a = np.frombuffer(s2, dtype=np.uint8)
    ofrag, oa, ou, ocand, cells = util.oracle_align_stream(prefix, s2a, one2one=False, apm=1)
    frag, a, u, cand, sta = db.alnFrags_batch(s2a, p, want_cand=True)
    db.seed_upload(s1); db.seed_run(p)
    db.align_from_seed(); db.align_run(p)
    frag2, a2, u2, _ = db.align_download()
    db.close()
    assert sta.tasks == len(ocand)
    assert np.array_equal(a, oa) and np.array_equal(u, ou)
    fb = frag.tobytes()
    assert fb == ofrag, f"frag_raw differs at byte {_first_diff(fb, ofrag)} of {len(ofrag)} (got {len(fb)})"
    assert frag2.tobytes() == ofrag and np.array_equal(a2, oa) and np.array_equal(u2, ou)


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
def test_pair_kernel_without_counters_is_the_same(tmp_path):
    """params.counters = 0 runs the pair kernel built without its statistic counters (kmagpu_align_fast.cu): same
    frag_raw bytes, score arrays and candidate rows; only the counters of the statistics stay zero. Short reads (the
    10-CTA variant) and long reads (the 6-CTA variant)."""
    from tests.test_oracle_pair import make_pairs
    prefix, s1, s2 = make_pairs(tmp_path, 53, n=2500)
    names, seqs = synth.gene_db(61, n_families=15, n_variants=6, len_lo=800, len_hi=3000)
    db = api.TemplateDB(prefix)
    for stream in (np.frombuffer(s2, dtype=np.uint8) if isinstance(s2, (bytes, bytearray)) else np.ascontiguousarray(s2),):
        p = api.default_params()
        frag1, a1, u1, c1, st1 = db.alnFrags_batch(stream, p, want_cand=True)
        p.counters = 0
        frag0, a0, u0, c0, st0 = db.alnFrags_batch(stream, p, want_cand=True)
        assert len(frag1) > 100000 and frag0.tobytes() == frag1.tobytes() and np.array_equal(a0, a1) and np.array_equal(u0, u1) and np.array_equal(c0, c1)
        assert st1.mems > 0 and st1.index_probes > 0 and st0.tasks == st1.tasks
        assert st0.mems * 4 < st1.mems and st0.index_probes * 4 < st1.index_probes   # only the large-scratch retries still count
    db.close()
    # long reads
    lp = tmp_path / "long"
    lp.mkdir()
    synth.write_fasta(lp / "db.fsa", names, seqs)
    reads = synth.long_reads(62, seqs, 60, len_lo=3000, len_hi=9000)
    synth.write_fastq(lp / "r.fq", reads, qual="5")
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=lp)
    s2l = np.frombuffer(util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-1t1", "-s2"], cwd=lp), dtype=np.uint8)
    db = api.TemplateDB(str(lp / "db"))
    p = api.default_params()
    p.one2one = 1
    frag1, a1, u1, c1, st1 = db.alnFrags_batch(s2l, p, want_cand=True)
    p.counters = 0
    frag0, a0, u0, c0, st0 = db.alnFrags_batch(s2l, p, want_cand=True)
    db.close()
    # the NW queue kernels count their cells in either build; only the pair kernel's own counters (MEMs, index probes) go away
    assert frag0.tobytes() == frag1.tobytes() and np.array_equal(a0, a1) and np.array_equal(c0, c1)
    assert st1.nw_band_calls > 0 and st0.nw_band_calls == st1.nw_band_calls and st0.nw_band_cells == st1.nw_band_cells
    assert st1.mems > 0 and st0.mems * 4 < st1.mems
