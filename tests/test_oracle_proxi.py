"""CPU tests: the oracle restatement of stage 2 under -proxi (kma.c:702-718: getProxiMatch savekmers.c:296,
getSecondProxiPen :1514, getF_Proxi :1764, getR_Proxi :1825, getProxiChainTemplates kmeranker.c:235, chooseChain's
proximity test :524-532) is pinned byte for byte to `kma -s2 -proxi X` -- what the -ont / -ill / -asm presets bind
(kma.c:1129, 1188, 1213). Stage 2 sees |X| (kma.c:1605); the sign only matters to stage 3 and to -mem_mode."""
import numpy as np
import pytest

from kma_b200 import synth
from tests import util
from tests.test_oracle_pair import make_pairs
from tests.test_oracle_chain import chain_case, tie_case, recombinant_case, overlap_case, lc_case

pytestmark = pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")


def se_case(tmp_path, seed, proxi, n=3000):
    names, seqs = synth.gene_db(seed, n_families=20, n_variants=8, len_lo=400, len_hi=1500)
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    reads = list(synth.short_reads(seed + 1, seqs, n, L=150, sub=0.03, junk_frac=0.05, n_rate=0.002))
    reads += list(synth.short_reads(seed + 2, seqs, 200, L=60, sub=0.05))
    synth.write_fastq(tmp_path / "r.fq", reads)
    args = ["-i", "r.fq", "-o", "o", "-t_db", "db", "-1t1", "-proxi", str(proxi)]
    s1 = util.ref_kma(args + ["-s1"], cwd=tmp_path)
    s2 = util.ref_kma(args + ["-s2"], cwd=tmp_path)
    s2_plain = util.ref_kma(args[:-2] + ["-s2"], cwd=tmp_path)
    return str(tmp_path / "db"), np.frombuffer(s1, dtype=np.uint8), s2, s2_plain


@pytest.mark.parametrize("seed,proxi", [(71, 0.9), (72, -0.98), (73, 0.5), (74, 0.0)])
def test_proxi_single_reads(tmp_path, seed, proxi):
    prefix, s1, s2, s2_plain = se_case(tmp_path, seed, proxi)
    assert s2 != s2_plain
    assert util.oracle_seed_stream(prefix, s1, proxi=abs(proxi)).tobytes() == s2
    assert util.oracle_seed_stream(prefix, s1).tobytes() == s2_plain


def pairs_with(tmp_path, seed, apm, proxi):
    prefix, s1, s2_plain = make_pairs(tmp_path, seed, apm=apm)
    s2 = util.ref_kma(["-ipe", "r1.fq", "r2.fq", "-o", "o", "-t_db", "db", "-apm", apm, "-proxi", str(proxi), "-s2"], cwd=tmp_path)
    return prefix, s1, s2, s2_plain


@pytest.mark.parametrize("seed,apm,proxi", [(81, "p", 0.9), (82, "p", -0.7), (83, "u", 0.9), (84, "u", 0.6), (85, "p", 0.98), (86, "u", -0.98)])
def test_proxi_read_pairs(tmp_path, seed, apm, proxi):
    prefix, s1, s2, s2_plain = pairs_with(tmp_path, seed, apm, proxi)
    assert s2 != s2_plain
    assert util.oracle_seed_stream(prefix, s1, apm=1 if apm == "u" else 0, proxi=abs(proxi)).tobytes() == s2


def chain_with(tmp_path, shape, proxi, lc=0):
    kw, extra = {}, []
    if shape == "chain":
        prefix, s1, _ = chain_case(tmp_path, 13, 120, 1000, 6000, 0.10, 0.002)
    elif shape == "chain_clean":
        prefix, s1, _ = chain_case(tmp_path, 12, 120, 1000, 6000, 0.03, 0.0)
    elif shape == "recombinant":
        prefix, s1, _ = recombinant_case(tmp_path, 77, 3000)
    elif shape == "overlap":
        prefix, s1, _ = overlap_case(tmp_path, 42, 0.5)
        kw, extra = {"coverT": 0.5}, ["-mct", "0.5"]
    elif shape == "lc":
        prefix, s1, _, _ = lc_case(tmp_path, 52)
    else:
        prefix, s1, _ = tie_case(tmp_path, 31)
    args = ["-i", "r.fq", "-o", "o", "-t_db", "db", "-s2", "-proxi", str(proxi)] + extra + (["-lc"] if lc else [])
    s2 = util.ref_kma(args, cwd=tmp_path)
    return prefix, s1, s2, dict(kw, lc=lc, proxi=abs(proxi))


@pytest.mark.parametrize("shape,proxi,lc", [("chain", 0.9, 0), ("chain_clean", -0.9, 0), ("recombinant", 0.9, 0), ("overlap", 0.8, 0),
                                            ("tie", 0.95, 0), ("lc", -0.9, 1), ("chain", -0.98, 1), ("recombinant", 0.7, 1)])
def test_proxi_chain_mode(tmp_path, shape, proxi, lc):
    prefix, s1, s2, kw = chain_with(tmp_path, shape, proxi, lc)
    got = util.oracle_chain_stream(prefix, s1, **kw)
    assert got.tobytes() == s2


@pytest.mark.parametrize("seed,proxi", [(71, 0.9), (72, -0.98), (73, -0.5)])
def test_proxi_alignment_pass_single_reads(tmp_path, seed, proxi):
    """stage 3 under -proxi: update_Scores' minFrac branches (updatescores.c:217-268) on the candidate lists stage 2 left"""
    prefix, s1, s2, _ = se_case(tmp_path, seed, proxi, n=1500)
    frag, a, u, _ = util.ref_align(prefix, s2, str(tmp_path), one2one=True, cand=False, min_frac=proxi)
    ofrag, oa, ou, _, _ = util.oracle_align_stream(prefix, np.frombuffer(s2, dtype=np.uint8), one2one=True, want_cand=False, min_frac=proxi)
    plain, _, _, _, _ = util.oracle_align_stream(prefix, np.frombuffer(s2, dtype=np.uint8), one2one=True, want_cand=False)
    assert ofrag == frag and ofrag != plain
    assert np.array_equal(oa, a) and np.array_equal(ou, u)


@pytest.mark.parametrize("seed,apm,proxi", [(81, "p", 0.9), (82, "p", -0.7), (83, "u", 0.9), (86, "u", -0.98)])
def test_proxi_alignment_pass_read_pairs(tmp_path, seed, apm, proxi):
    prefix, s1, s2, _ = pairs_with(tmp_path, seed, apm, proxi)
    frag, a, u, _ = util.ref_align(prefix, s2, str(tmp_path), one2one=False, cand=False, pe=apm, min_frac=proxi)
    ofrag, oa, ou, _, _ = util.oracle_align_stream(prefix, np.frombuffer(s2, dtype=np.uint8), one2one=False, want_cand=False,
                                                   apm=1 if apm == "u" else 0, min_frac=proxi)
    assert ofrag == frag
    assert np.array_equal(oa, a) and np.array_equal(ou, u)


@pytest.mark.parametrize("kind", ["se", "pe_p", "pe_u", "chain", "chain_lc"])
def test_soft_proximity_sums_of_mem_mode(tmp_path, kind):
    """-proxi < 0 with -mem_mode: stage 2 gets the negative value (kma.c:1605), every template a get*Proxi* function keeps
    adds its score to softProxi[], and the sums travel behind the stream (kmers.c:133-153) to become runKMA_MEM's
    alignment_scores (runkma.c:1153)"""
    extra = ["-mem_mode", "-proxi", "-0.9", "-s2"]
    if kind == "se":
        prefix, s1, _, _ = se_case(tmp_path, 71, 0.9, n=1500)
        s2 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-1t1"] + extra, cwd=tmp_path)
        with util.soft_proxi(prefix) as sp:
            got = util.oracle_seed_stream(prefix, s1, proxi=0.9).tobytes() + sp.trailer()
    elif kind in ("pe_p", "pe_u"):
        apm = kind[-1]
        prefix, s1, _ = make_pairs(tmp_path, 82, apm=apm)
        s2 = util.ref_kma(["-ipe", "r1.fq", "r2.fq", "-o", "o", "-t_db", "db", "-apm", apm] + extra, cwd=tmp_path)
        with util.soft_proxi(prefix) as sp:
            got = util.oracle_seed_stream(prefix, s1, apm=1 if apm == "u" else 0, proxi=0.9).tobytes() + sp.trailer()
    else:
        lc = kind == "chain_lc"
        prefix, s1, _ = chain_case(tmp_path, 12, 120, 1000, 6000, 0.03, 0.0)
        s2 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db"] + extra + (["-lc"] if lc else []), cwd=tmp_path)
        with util.soft_proxi(prefix) as sp:
            got = util.oracle_chain_stream(prefix, s1, proxi=0.9, lc=int(lc)).tobytes() + sp.trailer()
    assert sp.sums.sum() > 0
    assert got == s2
