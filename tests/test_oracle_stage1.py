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
