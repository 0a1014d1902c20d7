"""CPU tests: the oracle restatement of the alignment pass (MEM seeding, chaining, NW, alnFragsSE, update_Scores)
is pinned to the unmodified reference driven by oracle/ref_harness.c."""
import numpy as np
import pytest

from kma_b200 import synth
from tests import util

pytestmark = pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")


def _check(tmp_path, s2, prefix):
    frag, a, u, cand = util.ref_align(prefix, s2, str(tmp_path))
    ofrag, oa, ou, ocand, cells = util.oracle_align_stream(prefix, np.frombuffer(s2, dtype=np.uint8))
    assert util.cand_equal(ocand, cand)
    assert ofrag == frag
    assert np.array_equal(oa, a) and np.array_equal(ou, u)
    return cand, cells


def test_golden_stage2_through_alignment(tmp_path):
    with util.golden_dir() as g:
        s2 = open(f"{g}/s2.bin", "rb").read()
        cand, cells = _check(tmp_path, s2, f"{g}/db")
    assert len(cand) > 2000 and cells > 0


@pytest.mark.parametrize("seed,L,sub,indel", [(31, 150, 0.01, 0.0), (32, 150, 0.03, 0.01), (33, 400, 0.05, 0.02),
                                               (34, 1000, 0.04, 0.03), (35, 3000, 0.08, 0.06)])
def test_fresh_data_vs_reference(tmp_path, seed, L, sub, indel):
    """substitutions + indels, reads longer than 64 + band so that the banded NW and long tails are exercised"""
    names, seqs = synth.gene_db(seed, n_families=12, n_variants=6, len_lo=max(300, L + 50), len_hi=max(1500, 2 * L))
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    rng = np.random.default_rng(seed)
    base = synth.short_reads(seed + 1, seqs, 700, L=L, sub=0.0, n_rate=0.0, junk_frac=0.03)
    reads = [synth.mutate_indel(rng, r, sub, indel / 2, indel / 2) for r in base]
    for r in reads[::9]:
        if len(r) > 40:
            r[rng.integers(0, len(r), size=2)] = 4          # a few N's
    # palindromic-ish reads to force strand ties (read + its reverse complement overlapping the same template)
    for i in range(0, 60, 2):
        r = reads[i]
        reads[i] = np.concatenate([r[: len(r) // 2], synth.revcomp(r[: len(r) // 2])])
    synth.write_fastq(tmp_path / "r.fq", reads)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    s2 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-1t1", "-s2"], cwd=tmp_path)
    lib = util.orc()
    lib.orc_nw_band_calls.restype = __import__('ctypes').c_int64
    before = lib.orc_nw_band_calls()
    cand, cells = _check(tmp_path, s2, str(tmp_path / "db"))
    assert len(cand) > 500
    nband = lib.orc_nw_band_calls() - before
    print('banded NW calls:', nband)
    if L >= 400:
        assert nband > 0


@pytest.mark.parametrize("seed,err,n_rate", [(51, 0.10, 0.0), (52, 0.04, 0.002)])
def test_chain_mode_records_with_query_bounds(tmp_path, seed, err, n_rate):
    """stage-2 records of save_kmers_chain carry query bounds behind the name (qseqs.c:41); alnFragsSE hands them to
    anker_rc_comp / KMA_score (alnfrags.c:1091-1099), which restrict the seed scan of the first / last stretch."""
    from tests.test_oracle_chain import chain_case, tie_case
    prefix, s1, s2 = chain_case(tmp_path, seed, 80, 1000, 5000, err, n_rate)
    nb = sum(1 for r in __import__("kma_b200.records", fromlist=["x"]).parse_stage2(np.frombuffer(s2, np.uint8))
             if len(r["name"]) > 9 and r["name"][-9] == 0)
    assert nb > 50
    frag, a, u, cand = util.ref_align(prefix, s2, str(tmp_path), one2one=False)
    ofrag, oa, ou, ocand, cells = util.oracle_align_stream(prefix, np.frombuffer(s2, dtype=np.uint8), one2one=False)
    assert util.cand_equal(ocand, cand)
    assert ofrag == frag
    assert np.array_equal(oa, a) and np.array_equal(ou, u)
    d2 = tmp_path / "t"
    d2.mkdir()
    prefix, s1, s2 = tie_case(d2, seed)     # strand ties: anker_rc_comp with mirrored bounds
    frag, a, u, cand = util.ref_align(prefix, s2, str(d2), one2one=False)
    ofrag, oa, ou, ocand, cells = util.oracle_align_stream(prefix, np.frombuffer(s2, dtype=np.uint8), one2one=False)
    assert util.cand_equal(ocand, cand)
    assert ofrag == frag
    assert np.array_equal(oa, a) and np.array_equal(ou, u)
