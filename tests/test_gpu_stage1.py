"""Stage 1 through the C ABI. CPU part: the host record splitter (kmagpu_fastx_split). GPU part (B200): the device path
(kmagpu_stage1_batch: translation, end trim, filters, pairing rule, compDNA, printFsa) vs the oracle that
tests/test_oracle_stage1.py pins to `kma -s1`, and text -> stage 1 -> stage 2 chained in HBM."""
import numpy as np
import pytest

from kma_b200 import api, synth
from tests import util


def _reads(seed, n=700, L=150, n_rate=0.01):
    names, seqs = synth.gene_db(seed, n_families=4, n_variants=3, len_lo=400, len_hi=1200)
    rng = np.random.default_rng(seed)
    reads = [np.array(r) for r in synth.short_reads(seed + 1, seqs, n, L=L, sub=0.01, n_rate=n_rate)]
    for i in range(0, n, 7):
        reads[i] = reads[i][: int(rng.integers(5, L))]
    for i in range(3, n, 11):
        reads[i][: int(rng.integers(1, 6))] = 4
        reads[i][-int(rng.integers(1, 6)):] = 4
    reads[5] = reads[5][:0]           # an empty sequence line
    reads[9] = np.full(40, 4, dtype=np.uint8)   # nothing but N
    return rng, names, seqs, reads


def test_fastx_split_host():
    rng, _, _, reads = _reads(3, n=120)
    for crlf in (False, True):
        text = util.fastq_text(reads, util.random_quals(rng, reads), crlf=crlf)
        f, used = api.fastx_split(text)
        assert len(f) == len(reads) and used == len(text)
        lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
        for i, r in enumerate(reads):
            ho, hl, so, sl, qo = (int(x) for x in f[i])
            assert text[ho:ho + hl] == f"r{i}".encode() + (b" 1:N:0" if i % 3 == 0 else b"")
            assert text[so:so + sl] == lut[r].tobytes() and text[ho - 1:ho] == b"@"
            assert qo > so and (sl == 0 or 33 <= text[qo] < 127)
        # a chunk cut inside a record: the whole records before it, nothing more
        cut = int(f[50][2]) + 3
        f2, used2 = api.fastx_split(text[:cut])
        assert len(f2) == 50 and used2 == int(f[50][0]) - 1 and np.array_equal(f2, f[:50])
        # byte ranges split by several threads give the same table (quality lines that start with '@' included)
        for th in (2, 3, 7):
            assert np.array_equal(api.fastx_split_parallel(text, threads=th), f)
    fa = util.fastq_text(reads[:30], fasta=True)
    f3, used3 = api.fastx_split(fa, fastq=False)
    assert len(f3) == 30 and used3 == len(fa) and fa[int(f3[7][0]):int(f3[7][0]) + int(f3[7][1])] == b"r7 some description"
    with pytest.raises(api.KmaGpuError):
        api.fastx_split(b"ACGT\nACGT\n")


gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("seed,kw", [(11, {}), (12, {"min_phred": 30}), (13, {"minlen": 60}), (14, {"maxlen": 120}), (15, {"min_phred": 0}),
                                     (16, {"min_phred": 33, "phred_scale": 64})])
def test_single_end_vs_oracle(tmp_path, seed, kw):
    rng, names, seqs, reads = _reads(seed)
    scale = kw.get("phred_scale", 33)
    text = util.fastq_text(reads, util.random_quals(rng, reads, scale=scale), crlf=seed == 13)
    want, wcnt = util.oracle_stage1(text, **kw)
    prefix = util.build_db(tmp_path, names, seqs)
    db = api.TemplateDB(prefix, device=0)
    f, _ = api.fastx_split(text)
    got, cnt, ms = db.run_input_batch(text, f, **kw)
    db.close()
    assert got.tobytes() == want and cnt == wcnt and ms > 0


@gpu
def test_fasta_and_long_reads(tmp_path):
    rng, names, seqs, reads = _reads(17, n_rate=0.03)
    reads += [np.array(r) for r in synth.long_reads(18, seqs, 6, len_lo=800, len_hi=5000)]
    prefix = util.build_db(tmp_path, names, seqs)
    db = api.TemplateDB(prefix, device=0)
    text = util.fastq_text(reads, fasta=True)
    f, _ = api.fastx_split(text, fastq=False)
    got, cnt, _ = db.run_input_batch(text, f, fastq=False, minlen=40)
    want, wcnt = util.oracle_stage1(text, fastq=False, minlen=40)
    assert got.tobytes() == want and cnt == wcnt
    empty, c0, _ = db.run_input_batch(b"", np.zeros((0, 5), dtype=np.uint32))
    db.close()
    assert len(empty) == 0 and c0 == 0


@gpu
@pytest.mark.parametrize("seed,kw", [(21, {}), (22, {"min_phred": 28, "minlen": 50})])
def test_paired_end_vs_oracle(tmp_path, seed, kw):
    rng, names, seqs, r1 = _reads(seed, n=500)
    _, _, _, r2 = _reads(seed + 100, n=500)
    t1 = util.fastq_text(r1, util.random_quals(rng, r1))
    t2 = util.fastq_text(r2, util.random_quals(rng, r2))
    want, wcnt = util.oracle_stage1(t1, t2, **kw)
    f1, _ = api.fastx_split(t1)
    f2, _ = api.fastx_split(t2)
    f2 = f2.copy()
    f2[:, [0, 2, 4]] += len(t1)                      # both files in one buffer, mates interleaved
    fields = np.stack([f1, f2], axis=1).reshape(-1, 5)
    prefix = util.build_db(tmp_path, names, seqs)
    db = api.TemplateDB(prefix, device=0)
    got, cnt, _ = db.run_input_batch(t1 + t2, fields, paired=True, **kw)
    got2, cnt2, _ = db.run_input_batch(t1, fields, paired=True, text2=t2, **kw)   # the two files as separate host buffers
    db.close()
    assert got2.tobytes() == want and cnt2 == wcnt
    assert got.tobytes() == want and cnt == wcnt


@gpu
def test_text_to_stage2_in_hbm(tmp_path):
    """FASTQ text -> stage 1 on the device -> stage 2 without the records leaving HBM, single and paired"""
    rng, names, seqs, reads = _reads(31, n=900)
    prefix = util.build_db(tmp_path, names, seqs)
    db = api.TemplateDB(prefix, device=0)
    text = util.fastq_text(reads, util.random_quals(rng, reads))
    s1, _ = util.oracle_stage1(text)
    want = util.oracle_seed_stream(prefix, np.frombuffer(s1, dtype=np.uint8))
    f, _ = api.fastx_split(text)
    _, cnt, _ = db.run_input_batch(text, f, download=False)
    st = db.seed_run(api.default_params())
    out = np.empty(len(want) + 64, dtype=np.uint8)
    got = db.seed_download(out)
    assert got.tobytes() + api.stream_terminator(cnt) == want.tobytes() and st.mapped > 300
    r1, r2 = synth.paired_reads(33, seqs, 400, sub=0.01)
    t1, t2 = util.fastq_text(r1), util.fastq_text(r2)
    s1pe, _ = util.oracle_stage1(t1, t2)
    wantpe = util.oracle_seed_stream(prefix, np.frombuffer(s1pe, dtype=np.uint8))
    f1, _ = api.fastx_split(t1)
    f2, _ = api.fastx_split(t2)
    f2 = f2.copy(); f2[:, [0, 2, 4]] += len(t1)
    _, cntpe, _ = db.run_input_batch(t1 + t2, np.stack([f1, f2], axis=1).reshape(-1, 5), paired=True, download=False)
    db.seed_run(api.default_params())
    out = np.empty(len(wantpe) + 64, dtype=np.uint8)
    gotpe = db.seed_download(out)
    db.close()
    assert gotpe.tobytes() + api.stream_terminator(cntpe) == wantpe.tobytes()


def _first_record_diff(a: bytes, b: bytes) -> str:
    """first stage-1 record in which two streams differ (for assertion messages)"""
    pa = pb = i = 0
    while pa + 16 <= len(a) and pb + 16 <= len(b):
        ha, hb = np.frombuffer(a, np.int32, 4, pa), np.frombuffer(b, np.int32, 4, pb)
        sa, sb = 16 + 8 * int(ha[1]) + 4 * int(ha[2]) + abs(int(ha[3])), 16 + 8 * int(hb[1]) + 4 * int(hb[2]) + abs(int(hb[3]))
        if a[pa:pa + sa] != b[pb:pb + sb]:
            return f"record {i}: got {ha.tolist()} {a[pa + sa - abs(int(ha[3])):pa + sa]!r} want {hb.tolist()} {b[pb + sb - abs(int(hb[3])):pb + sb]!r}"
        pa += sa; pb += sb; i += 1
    return f"lengths {len(a)} vs {len(b)} after {i} equal records"


@gpu
def test_multi_line_fasta_on_the_device(tmp_path):
    """multi-line FASTA: the host unwrap (kmagpu_fasta_unwrap) + the device's 2-line path vs the oracle that
    tests/test_oracle_stage1.py pins to `kma -s1` on the wrapped file"""
    from tests.test_oracle_stage1 import _wrap_fasta
    rng, names, seqs, reads = _reads(51, n=400, L=400, n_rate=0.02)
    text = _wrap_fasta(rng, reads, crlf=True)
    flat, used = api.fasta_unwrap(text)
    assert used == len(text) and flat == util.oracle_fasta_unwrap(text)
    want, wcnt = util.oracle_stage1(flat, fastq=False, minlen=40)
    prefix = util.build_db(tmp_path, names, seqs)
    db = api.TemplateDB(prefix, device=0)
    got, cnt, _, _, _ = db.run_input_text(flat, fastq=False, minlen=40)
    db.close()
    assert got.tobytes() == want and cnt == wcnt and cnt > 300


@gpu
@pytest.mark.parametrize("seed,kw", [(31, {"min_q": 20}), (32, {"min_q": 25, "min_phred": 10}), (33, {"hardmask_q": 30, "min_phred": 30}),
                                     (34, {"min_q": 18, "hardmask_q": 28, "min_phred": 28, "minlen": 40}), (35, {"min_q": 30, "min_phred": 35}),
                                     (36, {"min_q": 12, "min_phred": 0})])
def test_quality_trim_and_hard_mask(tmp_path, seed, kw):
    """-eq / -mi: phredStat's bidirectional quality trim and hard mask (runinput.c:168-313) on the device, vs the oracle that
    tests/test_oracle_stage1.py pins to `kma -s1 -eq / -mi`; single reads through both entry points and read pairs"""
    rng, names, seqs, reads = _reads(seed, n=900)
    quals = util.random_quals(rng, reads)
    for i in range(0, len(reads), 3):
        L = len(reads[i])
        if L < 4:
            continue
        ramp = np.linspace(40, 2, L) if i % 2 else np.concatenate([np.linspace(3, 40, L // 2), np.linspace(40, 3, L - L // 2)])
        quals[i] = np.clip(ramp + rng.integers(-6, 7, size=L), 0, 41).astype(np.uint8) + 33
    text = util.fastq_text(reads, quals)
    want, wcnt = util.oracle_stage1(text, **kw)
    plain, _ = util.oracle_stage1(text, **{k: v for k, v in kw.items() if k not in ("min_q", "hardmask_q")})
    assert want != plain or "min_q" not in kw
    prefix = util.build_db(tmp_path, names, seqs)
    db = api.TemplateDB(prefix, device=0)
    f, _ = api.fastx_split(text)
    got, cnt, _ = db.run_input_batch(text, f, **kw)
    assert cnt == wcnt, (cnt, wcnt)
    assert got.tobytes() == want, _first_record_diff(got.tobytes(), want)
    got, cnt, _, _, _ = db.run_input_text(text, **kw)
    assert got.tobytes() == want and cnt == wcnt
    _, _, _, r2 = _reads(seed + 100, n=900)
    t2 = util.fastq_text(r2, util.random_quals(rng, r2))
    wpe, cpe = util.oracle_stage1(text, t2, **kw)
    gpe, gcpe, _, _, _ = db.run_input_text(text, text2=t2, **kw)
    db.close()
    assert gpe.tobytes() == wpe and gcpe == cpe


@gpu
@pytest.mark.parametrize("crlf", [False, True])
def test_device_splitter_vs_oracle(tmp_path, crlf):
    """kmagpu_stage1_text: the record splitter on the device too -- single end, FASTA, pairs, a chunk cut inside a record,
    a file without its last newline"""
    rng, names, seqs, reads = _reads(41, n=900)
    prefix = util.build_db(tmp_path, names, seqs)
    db = api.TemplateDB(prefix, device=0)
    text = util.fastq_text(reads, util.random_quals(rng, reads), crlf=crlf)
    want, wcnt = util.oracle_stage1(text)
    got, cnt, ms, used, _ = db.run_input_text(text)
    assert got.tobytes() == want and cnt == wcnt and used == len(text) and ms > 0
    got, cnt, _, used, _ = db.run_input_text(text[:-1])               # no newline at the end of the file
    assert got.tobytes() == want and cnt == wcnt and used == len(text) - 1
    f, _ = api.fastx_split(text)
    cut = int(f[500][2]) + 7                                          # a chunk that ends inside record 500
    part, pcnt, _, used, _ = db.run_input_text(text[:cut], eof=False)
    w2, c2 = util.oracle_stage1(text[:int(f[500][0]) - 1])
    assert part.tobytes() == w2 and pcnt == c2 and used == int(f[500][0]) - 1
    rest, rcnt, _, used2, _ = db.run_input_text(text[used:])          # the carried-over rest gives the other records
    assert part.tobytes() + rest.tobytes() == want and pcnt + rcnt == wcnt
    fa = util.fastq_text(reads, fasta=True, crlf=crlf)
    wfa, cfa = util.oracle_stage1(fa, fastq=False, minlen=40)
    gfa, gc, _, _, _ = db.run_input_text(fa, fastq=False, minlen=40)
    assert gfa.tobytes() == wfa and gc == cfa
    _, _, _, r2 = _reads(141, n=820)                                  # the second file is shorter: 820 pairs are taken
    t2 = util.fastq_text(r2, util.random_quals(rng, r2), crlf=crlf)
    f1, _ = api.fastx_split(text)
    wpe, cpe = util.oracle_stage1(text[:int(f1[820][0]) - 1], t2)
    gpe, gcpe, _, u1, u2 = db.run_input_text(text, text2=t2)
    assert gpe.tobytes() == wpe and gcpe == cpe and u1 == int(f1[820][0]) - 1 and u2 == len(t2)
    with pytest.raises(api.KmaGpuError):
        db.run_input_text(b"@r0\nACGTACGTACGTACGTACGT\n+\nIIIIIIIIIIIIIIIIIIII\nr1\nACGT\n+\nIIII\n")   # second record without '@'
    empty, c0, _, _, _ = db.run_input_text(b"")
    db.close()
    assert len(empty) == 0 and c0 == 0


@gpu
def test_text_pipelines_equal_record_pipeline(tmp_path):
    """MapPipeline.map_text (host splitter) and map_text_device_split == the resident call on the stage-1 records of the
    same reads; files whose records do not line up make the device-split path step aside"""
    from kma_b200 import pipeline, records
    names, seqs = synth.gene_db(51, n_families=10, n_variants=6, len_lo=500, len_hi=1500)
    prefix = util.build_db(tmp_path, names, seqs)
    r1, r2 = synth.paired_reads(52, seqs, 3000, sub=0.01)
    r1, r2 = np.asarray(r1), np.asarray(r2)
    s1 = records.stage1_pairs_fast(r1, r2)
    p = api.default_params()
    db = api.TemplateDB(prefix)
    db.seed_upload(s1); db.seed_run(p); db.align_from_seed(); db.align_run(p)
    frag, a, u, _ = db.align_download()
    db.close()
    t1, t2 = synth.fastq_fixed(r1), synth.fastq_fixed(r2)
    pipe = pipeline.MapPipeline(prefix, workers=2, params=p)
    outs = [np.empty(len(frag) + 4096, dtype=np.uint8) for _ in range(5)]
    f1, f2 = api.fastx_split_parallel(t1, threads=3), api.fastx_split_parallel(t2, threads=2)
    for mode in ("host", "device"):
        scores = (np.zeros_like(a), np.zeros_like(u))
        res = pipe.map_text(t1, f1, t2, f2, 5, outs, scores) if mode == "host" else pipe.map_text_device_split(t1, t2, 5, outs, scores)
        assert res is not None and b"".join(f.tobytes() for f, _ in res) == frag.tobytes(), mode
        assert np.array_equal(scores[0], a) and np.array_equal(scores[1], u) and sum(c for _, c in res) == 3000
    # second file with longer names: same records, different byte positions -> chunks do not line up
    t2b = np.frombuffer(util.fastq_text(list(r2), prefix="a_longer_name_"), dtype=np.uint8)
    scores = (np.zeros_like(a), np.zeros_like(u))
    assert pipe.map_text_device_split(t1, t2b, 5, outs, scores) is None
    pipe.close()
