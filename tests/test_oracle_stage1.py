"""CPU tests: the oracle restatement of stage 1 (record splitter, to2Bit, phredStat's end trim / fsastat's N trim, the
-ml / -xl filters, the pairing rule of run_input_PE, compDNA, printFsa / printFsa_pair) is pinned byte-exact to the
unmodified reference: `kma -i / -ipe ... -s1`."""
import numpy as np
import pytest

from kma_b200 import synth
from tests import util

pytestmark = pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")


def make(tmp_path, seed, n=600, L=150, n_rate=0.01):
    names, seqs = synth.gene_db(seed, n_families=4, n_variants=3, len_lo=400, len_hi=1200)
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    rng = np.random.default_rng(seed)
    reads = [np.array(r) for r in synth.short_reads(seed + 1, seqs, n, L=L, sub=0.01, n_rate=n_rate)]
    for i in range(0, n, 7):    # ragged lengths, some below -ml
        reads[i] = reads[i][: int(rng.integers(5, L))]
    for i in range(3, n, 11):   # N runs at the ends (fsastat trims them)
        reads[i][: int(rng.integers(1, 6))] = 4
        reads[i][-int(rng.integers(1, 6)):] = 4
    return rng, reads


@pytest.mark.parametrize("seed,extra,kw", [(11, [], {}), (12, ["-mp", "30"], {"min_phred": 30}), (13, ["-ml", "60"], {"minlen": 60}),
                                            (14, ["-xl", "120"], {"maxlen": 120}), (15, ["-mp", "0"], {"min_phred": 0})])
def test_single_end_fastq(tmp_path, seed, extra, kw):
    rng, reads = make(tmp_path, seed)
    text = util.fastq_text(reads, util.random_quals(rng, reads), crlf=seed == 13)
    (tmp_path / "r.fq").write_bytes(text)
    want = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-s1"] + extra, cwd=tmp_path)
    got, cnt = util.oracle_stage1(text, **kw)
    assert got == want
    assert 0 < cnt < len(reads) or kw.get("min_phred") == 0


def test_phred64(tmp_path):
    rng, reads = make(tmp_path, 16, n=200)
    text = util.fastq_text(reads, [np.maximum(q, 95) for q in util.random_quals(rng, reads, scale=64)])
    (tmp_path / "r.fq").write_bytes(text)
    want = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-s1", "-mp", "33"], cwd=tmp_path)
    got, _ = util.oracle_stage1(text, min_phred=33, phred_scale=64)
    assert got == want


def test_fasta(tmp_path):
    rng, reads = make(tmp_path, 17, n_rate=0.03)
    text = util.fastq_text(reads, fasta=True)
    (tmp_path / "r.fa").write_bytes(text)
    want = util.ref_kma(["-i", "r.fa", "-o", "o", "-t_db", "db", "-s1", "-ml", "40"], cwd=tmp_path)
    got, _ = util.oracle_stage1(text, fastq=False, minlen=40)
    assert got == want


@pytest.mark.parametrize("seed,extra,kw", [(21, [], {}), (22, ["-mp", "28", "-ml", "50"], {"min_phred": 28, "minlen": 50})])
def test_paired_end(tmp_path, seed, extra, kw):
    rng, r1 = make(tmp_path, seed, n=500)
    _, r2 = make(tmp_path, seed + 100, n=500)
    t1 = util.fastq_text(r1, util.random_quals(rng, r1))
    t2 = util.fastq_text(r2, util.random_quals(rng, r2))
    (tmp_path / "a.fq").write_bytes(t1)
    (tmp_path / "b.fq").write_bytes(t2)
    want = util.ref_kma(["-ipe", "a.fq", "b.fq", "-o", "o", "-t_db", "db", "-s1"] + extra, cwd=tmp_path)
    got, cnt = util.oracle_stage1(t1, t2, **kw)
    assert got == want
    h = np.frombuffer(want[:16], dtype=np.int32)
    assert cnt > 300 and (np.frombuffer(want, dtype=np.uint8).size > 0) and h[0] > 0


def test_quality_table_equals_the_reference_literals():
    """prob[q] = 10^(-q/10): kma.c:219 holds the table as 32-digit literals, kma_b200.api.quality_prob() computes it"""
    import os, re
    from kma_b200 import api
    src = "/root/reference/kma.c"
    if not os.path.exists(src):
        pytest.skip("reference sources not here")
    m = re.search(r"static const double prob\[256\] = \{([^}]*)\}", open(src).read())
    lit = np.array([float(x) for x in m.group(1).replace("\n", " ").split(",")], dtype=np.float64)
    assert len(lit) == 256 and np.array_equal(lit, api.quality_prob())


@pytest.mark.parametrize("seed,extra,kw", [(31, ["-eq", "20"], {"min_q": 20}), (32, ["-eq", "25", "-mp", "10"], {"min_q": 25, "min_phred": 10}),
                                            (33, ["-mi", "30"], {"hardmask_q": 30, "min_phred": 30}),
                                            (34, ["-eq", "18", "-mi", "28", "-ml", "40"], {"min_q": 18, "hardmask_q": 28, "min_phred": 28, "minlen": 40}),
                                            (35, ["-eq", "30", "-mp", "35"], {"min_q": 30, "min_phred": 35})])
def test_quality_trim_and_hard_mask(tmp_path, seed, extra, kw):
    """-eq (phredStat's bidirectional trim, runinput.c:196-296) and -mi (hard mask on the raw quality byte, :183; the CLI
    also raises -mp to it, kma.c:1554, which is all it ever does: a byte that survives the end trim is above it)"""
    rng, reads = make(tmp_path, seed, n=800)
    quals = util.random_quals(rng, reads)
    for i in range(0, len(reads), 3):   # reads whose quality decays towards one or both ends: the segment-wise trim has work to do
        L = len(reads[i])
        ramp = np.linspace(40, 2, L) if i % 2 else np.concatenate([np.linspace(3, 40, L // 2), np.linspace(40, 3, L - L // 2)])
        quals[i] = np.clip(ramp + rng.integers(-6, 7, size=L), 0, 41).astype(np.uint8) + 33
    text = util.fastq_text(reads, quals)
    (tmp_path / "r.fq").write_bytes(text)
    want = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-s1"] + extra, cwd=tmp_path)
    plain = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-s1"], cwd=tmp_path)
    got, cnt = util.oracle_stage1(text, **kw)
    assert want != plain
    assert got == want


def test_quality_trim_paired(tmp_path):
    rng, r1 = make(tmp_path, 41, n=400)
    _, r2 = make(tmp_path, 141, n=400)
    q1, q2 = util.random_quals(rng, r1), util.random_quals(rng, r2)
    t1, t2 = util.fastq_text(r1, q1), util.fastq_text(r2, q2)
    (tmp_path / "r1.fq").write_bytes(t1)
    (tmp_path / "r2.fq").write_bytes(t2)
    want = util.ref_kma(["-ipe", "r1.fq", "r2.fq", "-o", "o", "-t_db", "db", "-s1", "-eq", "22", "-mi", "25"], cwd=tmp_path)
    got, _ = util.oracle_stage1(t1, t2, min_q=22, hardmask_q=25, min_phred=25)
    assert got == want


def _wrap_fasta(rng, reads, crlf=False):
    """multi-line FASTA: lines of 60, a few 70, blank lines, lower case, blanks inside lines"""
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    nl = b"\r\n" if crlf else b"\n"
    out = []
    for i, r in enumerate(reads):
        s = lut[r].tobytes()
        if i % 5 == 0:
            s = s.lower()
        w = 70 if i % 4 == 0 else 60
        lines = [s[j:j + w] for j in range(0, len(s), w)] or [b""]
        if i % 7 == 0 and len(lines) > 1:
            lines.insert(1, b"")
        if i % 9 == 0:
            lines[0] = lines[0][:10] + b" " + lines[0][10:]
        out.append(b">r%d some description" % i + (b"  " if i % 3 == 0 else b"") + nl + nl.join(lines) + nl)
    return b"".join(out)


@pytest.mark.parametrize("crlf", [False, True])
def test_multi_line_fasta(tmp_path, crlf):
    """FileBuffgetFsa (seqparse.c:66-160) keeps every byte that translates below 8 between a header line and the next '>':
    the oracle's unwrap + its 2-line path == `kma -s1` on the wrapped file; the library's host function gives the same bytes,
    whole and in chunks"""
    from kma_b200 import api
    rng, reads = make(tmp_path, 51, n=300, L=400, n_rate=0.02)
    text = _wrap_fasta(rng, reads, crlf)
    (tmp_path / "r.fa").write_bytes(text)
    want = util.ref_kma(["-i", "r.fa", "-o", "o", "-t_db", "db", "-s1", "-ml", "40"], cwd=tmp_path)
    flat = util.oracle_fasta_unwrap(text)
    got, _ = util.oracle_stage1(flat, fastq=False, minlen=40)
    assert got == want and len(want) > 10000
    lib_flat, used = api.fasta_unwrap(text)
    assert lib_flat == flat and used == len(text)
    # chunks: a cut inside a record leaves that record (and the bytes after it) for the next call
    cut = len(text) // 2
    a, ua = api.fasta_unwrap(text[:cut], eof=False)
    assert 0 < ua <= cut and text[ua:ua + 1] == b">"
    b, ub = api.fasta_unwrap(text[ua:])
    assert a + b == flat and ua + ub == len(text)
    with pytest.raises(api.KmaGpuError):
        api.fasta_unwrap(b"ACGT\n>r1\nACGT\n")


def test_host_reader_gz_fasta_and_phred_scale(tmp_path):
    """kma_b200.pipeline.read_reads / phred_scale: gzip input (one and several members), multi-line FASTA, and the phred scale
    getPhredFileBuff (seqparse.c:551) reports -- the text they hand to stage 1 gives the reference's stage-1 stream"""
    import gzip, re
    from kma_b200 import pipeline
    rng, reads = make(tmp_path, 61, n=300)
    for scale, mp in ((33, "20"), (64, "20")):
        quals = util.random_quals(rng, reads, scale=scale)
        if scale == 64:
            quals = [np.maximum(q, 95) for q in quals]
        text = util.fastq_text(reads, quals)
        half = text.index(b"\n@r150") + 1
        (tmp_path / "r.fq.gz").write_bytes(gzip.compress(text[:half]) + gzip.compress(text[half:]))   # two members
        got, fastq, ps = pipeline.read_reads(str(tmp_path / "r.fq.gz"))
        r = __import__("subprocess").run([util.REF_KMA, "-i", "r.fq.gz", "-o", "o", "-t_db", "db", "-s1", "-mp", mp], cwd=tmp_path,
                                          stdout=__import__("subprocess").PIPE, stderr=__import__("subprocess").PIPE)
        ref_scale = int(re.search(rb"Phred scale:\s*(\d+)", r.stderr).group(1))
        assert got == text and fastq and ps == ref_scale == scale
        assert util.oracle_stage1(got, min_phred=int(mp), phred_scale=ps)[0] == r.stdout
    fa = _wrap_fasta(rng, reads)
    (tmp_path / "r.fa.gz").write_bytes(gzip.compress(fa))
    got, fastq, _ = pipeline.read_reads(str(tmp_path / "r.fa.gz"))
    want = util.ref_kma(["-i", "r.fa.gz", "-o", "o", "-t_db", "db", "-s1"], cwd=tmp_path)
    assert not fastq and util.oracle_stage1(got, fastq=False)[0] == want


def test_fasta_unwrap_fuzz():
    """kmagpu_fasta_unwrap == the oracle's restatement on random texts: odd bytes, blank lines, '>' right after a header, records
    without sequence, missing final newline, and every cut position of a small text in the chunked mode"""
    from kma_b200 import api
    rng = np.random.default_rng(5)
    alphabet = np.frombuffer(b"ACGTNacgtnRYKMxX*- \t\r\n\n\n", dtype=np.uint8)
    for it in range(200):
        recs = []
        for r in range(int(rng.integers(1, 6))):
            hdr = b">h%d" % r + (b" d e" if rng.random() < 0.5 else b"") + (b"\r" if rng.random() < 0.2 else b"")
            body = alphabet[rng.integers(0, len(alphabet), size=int(rng.integers(0, 80)))].tobytes().replace(b">", b"")
            recs.append(hdr + b"\n" + body + (b"\n" if rng.random() < 0.8 else b""))
        text = b"".join(recs)
        want = util.oracle_fasta_unwrap(text)
        got, used = api.fasta_unwrap(text)
        assert got == want and used == len(text), (it, text)
        if it % 10 == 0:   # a header line cut off by the end of the file is no record
            t2 = text + b">last one"
            assert api.fasta_unwrap(t2)[0] == util.oracle_fasta_unwrap(t2) == want
        if it < 20:   # chunked: any cut, the rest carried over
            for cut in range(1, len(text)):
                a, ua = api.fasta_unwrap(text[:cut], eof=False)
                b, ub = api.fasta_unwrap(text[ua:])
                assert a + b == want and ua + ub == len(text), (it, cut, text)
