"""CPU tests: the oracle restatement of the consensus call of the assembly pass (callConsensus assembly.c:1499-1631,
the base callers :162-271, the significance tests :141-160) is pinned byte-exact to the unmodified reference's own
callConsensus driven by oracle/ref_harness.c -consensus."""
import numpy as np
import pytest

from kma_b200 import synth
from tests import util

pytestmark = pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")

# (caller, significance, support, bcd, evalue): the CLI's combinations -- default, -bc90, -bcg, -bcNano (-> and90),
# -bcNano -bc 0.7 -bcd 10 (-ont), -bcg -bc 0.9 -bcd 10 (-mint2), the two -ref_fsa callers, strict and lax evalues
CASES = [(0, 0, 0.0, 1, 0.05), (0, 1, 0.0, 1, 0.05), (1, 0, 0.0, 1, 0.05), (3, 1, 0.0, 1, 0.05), (3, 2, 0.7, 10, 0.05),
         (1, 2, 0.9, 10, 0.05), (2, 0, 0.0, 1, 0.05), (4, 1, 0.0, 5, 0.05), (0, 0, 0.0, 3, 1e-6), (3, 2, 0.5, 1, 0.9),
         (0, 0, 0.0, 1, 1.0), (2, 2, 0.7, 20, 1e-9)]


def make_db(tmp_path, seed):
    names, seqs = synth.gene_db(seed, n_families=4, n_variants=3, len_lo=300, len_hi=2500)
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    return str(tmp_path / "db"), len(seqs)


@pytest.mark.parametrize("case", range(len(CASES)))
def test_consensus_vs_reference(tmp_path, case):
    caller, sig, support, bcd, evalue = CASES[case]
    prefix, n = make_db(tmp_path, 70 + case)
    rng = np.random.default_rng(500 + case)
    mats = {t: util.random_count_matrix(rng, util.template_bases(prefix, t)) for t in range(1, n + 1)}
    want = util.ref_consensus(prefix, mats, str(tmp_path), bcd=bcd, evalue=evalue, caller=caller, sig=sig, support=support)
    assert sorted(want) == sorted(mats)
    seen = set()
    for t, m in mats.items():
        got = util.oracle_consensus(prefix, t, m, bcd=bcd, evalue=evalue, caller=caller, sig=sig, support=support)
        wt, ws, wq, wst = want[t]
        assert got[0] == wt and got[2] == wq and got[1] == ws, f"template {t}"
        assert np.array_equal(got[3], wst), (t, got[3], wst)
        seen |= set(wq)
    assert {ord("A"), ord("a")} <= seen and (ord("n") in seen or caller == 1) and (ord("-") in seen or caller in (2, 4))


def test_chi2_threshold_is_the_decision():
    """the threshold the device compares against splits p_chisqr(x) <= evalue exactly (monotone in x)"""
    import ctypes as C
    ref = C.CDLL(util.REF_SO)
    ref.p_chisqr.restype = C.c_double
    ref.p_chisqr.argtypes = [C.c_longdouble]
    for ev in (0.05, 0.01, 1e-6, 0.9, 1e-9):
        x0 = util.oracle_chi2_min(ev)
        assert 0 < x0 < 49
        assert ref.p_chisqr(x0) <= ev < ref.p_chisqr(np.nextafter(x0, 0.0))
        rng = np.random.default_rng(7)
        for x in np.concatenate([rng.uniform(0, 49, 2000), x0 + rng.uniform(-1e-9, 1e-9, 2000)]):
            assert (ref.p_chisqr(float(x)) <= ev) == (x >= x0)
    assert util.oracle_chi2_min(1.0) == 0.0
