"""CPU tests: -mem_mode's k-mer score collection (runKMA_MEM runkma.c:1088-1140, update_Scores_MEM / _pe_MEM) restated
in the oracle, pinned to the reference's own functions run by oracle/ref_harness.c -memscore."""
import numpy as np
import pytest

from kma_b200 import synth
from tests import util

pytestmark = pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")


def _db(tmp_path, seed):
    names, seqs = synth.gene_db(seed, n_families=10, n_variants=6, len_lo=300, len_hi=1200)
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    return str(tmp_path / "db"), seqs


def test_memscore_single_end(tmp_path):
    prefix, seqs = _db(tmp_path, 91)
    rng = np.random.default_rng(91)
    reads = list(synth.short_reads(92, seqs, 1200, L=150, sub=0.01, junk_frac=0.03, n_rate=0.002))
    reads += [rng.integers(0, 4, size=12).astype(np.uint8)]
    for i in range(0, 40, 2):   # strand ties
        r = reads[i]
        reads[i] = np.concatenate([r[:75], synth.revcomp(r[:75])])
    synth.write_fastq(tmp_path / "r.fq", reads)
    s2 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-1t1", "-s2"], cwd=tmp_path)
    frag, a, u = util.ref_memscore(prefix, s2, str(tmp_path))
    ofrag, oa, ou = util.oracle_memscore(prefix, s2)
    assert len(frag) > 100000 and ofrag == frag
    assert np.array_equal(oa, a) and np.array_equal(ou, u) and int(u.sum()) > 0


def test_memscore_paired_end(tmp_path):
    prefix, seqs = _db(tmp_path, 93)
    r1, r2 = synth.paired_reads(94, seqs, 800, sub=0.01)
    synth.write_fastq(tmp_path / "a.fq", r1)
    synth.write_fastq(tmp_path / "b.fq", r2)
    s2 = util.ref_kma(["-ipe", "a.fq", "b.fq", "-o", "o", "-t_db", "db", "-apm", "p", "-s2"], cwd=tmp_path)
    frag, a, u = util.ref_memscore(prefix, s2, str(tmp_path))
    ofrag, oa, ou = util.oracle_memscore(prefix, s2)
    assert ofrag == frag
    assert np.array_equal(oa, a) and np.array_equal(ou, u)
