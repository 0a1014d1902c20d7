"""CPU tests: the oracle restatement of the paired-end path (-ipe ... -apm p: get_kmers_for_pair,
save_kmers_penaltyPair, printPair, alnFragsPenaltyPE, update_Scores_pe/_se) is pinned to the unmodified reference:
stage-2 streams byte-exact vs `kma -s2`, frag_raw stream + ConClave arrays vs alnFrags_threaded (oracle/ref_harness.c)."""
import collections

import numpy as np
import pytest

from kma_b200 import synth, records
from tests import util

pytestmark = pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")


def make_pairs(tmp_path, seed, n=1500, apm="p"):
    """pairs with junk mates, N's, mates from different templates, same-strand mates"""
    names, seqs = synth.gene_db(seed, n_families=20, n_variants=6, len_lo=400, len_hi=1500)
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    r1, r2 = synth.paired_reads(seed + 1, seqs, n, sub=0.02)
    rng = np.random.default_rng(seed)
    r1, r2 = list(r1), list(r2)
    for i in range(0, n, 17):
        r2[i] = rng.integers(0, 4, size=150).astype(np.uint8)
    for i in range(5, n, 23):
        r1[i] = rng.integers(0, 4, size=150).astype(np.uint8)
    for i in range(3, n, 29):
        r1[i] = r1[i].copy()
        r1[i][rng.integers(0, 150, size=3)] = 4
    for i in range(11, n, 37):
        r1[i] = rng.integers(0, 4, size=150).astype(np.uint8)
        r2[i] = rng.integers(0, 4, size=150).astype(np.uint8)
    for i in range(13, n, 41):
        r2[i] = r1[(i * 7) % n]
    synth.write_fastq(tmp_path / "r1.fq", r1)
    synth.write_fastq(tmp_path / "r2.fq", r2)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    args = ["-ipe", "r1.fq", "r2.fq", "-o", "o", "-t_db", "db", "-apm", apm]
    s1 = util.ref_kma(args + ["-s1"], cwd=tmp_path)
    s2 = util.ref_kma(args + ["-s2"], cwd=tmp_path)
    return str(tmp_path / "db"), np.frombuffer(s1, dtype=np.uint8), s2


@pytest.mark.parametrize("seed", [51, 52, 53])
def test_pair_seeding_and_alignment_vs_reference(tmp_path, seed):
    prefix, s1, s2 = make_pairs(tmp_path, seed)
    got = util.oracle_seed_stream(prefix, s1)
    assert got.tobytes() == s2
    recs = records.parse_stage2(np.frombuffer(s2, dtype=np.uint8))
    kinds = collections.Counter((r["flag"], len(r["templates"]) == 0) for r in recs)
    assert sum(1 for (f, first) in kinds if first) >= 2 and len(kinds) >= 8   # proper pairs in both orders + single mates
    frag, a, u, _ = util.ref_align(prefix, s2, str(tmp_path), one2one=False, cand=False, pe=True)
    ofrag, oa, ou, ocand, cells = util.oracle_align_stream(prefix, np.frombuffer(s2, dtype=np.uint8), one2one=False)
    assert ofrag == frag
    assert np.array_equal(oa, a) and np.array_equal(ou, u)
    assert cells > 0 and len(ocand) > 1000


@pytest.mark.parametrize("seed", [61, 62, 63])
def test_union_pairing_stage2_vs_reference(tmp_path, seed):
    """-apm u, the reference's default pairing: save_kmers_unionPair with getF_Best / getR_Best (savekmers.c:3367, 1648, 1682)"""
    prefix, s1, s2 = make_pairs(tmp_path, seed, apm="u")
    got = util.oracle_seed_stream(prefix, s1, apm=1)
    assert got.tobytes() == s2
    recs = records.parse_stage2(np.frombuffer(s2, dtype=np.uint8))
    kinds = collections.Counter((r["flag"], len(r["templates"]) == 0) for r in recs)
    assert sum(1 for (f, first) in kinds if first) >= 2 and len(kinds) >= 8
    # stage 3: alnFragsUnionPE (alnfrags.c:1220) + update_Scores_pe / _se
    frag, a, u, _ = util.ref_align(prefix, s2, str(tmp_path), one2one=False, cand=False, pe="u")
    ofrag, oa, ou, ocand, cells = util.oracle_align_stream(prefix, np.frombuffer(s2, dtype=np.uint8), one2one=False, apm=1)
    assert ofrag == frag
    assert np.array_equal(oa, a) and np.array_equal(ou, u)
    pfrag, _, _, _, _ = util.oracle_align_stream(prefix, np.frombuffer(s2, dtype=np.uint8), one2one=False, apm=0)
    assert pfrag != ofrag, "the case is meant to tell the two pairings apart"
