"""GPU parity tests (B200): stage 2 in chain mode (save_kmers_chain, the long-read default without -1t1) through the
C ABI vs the unmodified reference (`kma -s2`) and the oracle restatement, byte for byte."""
import numpy as np
import pytest

from kma_b200 import api, synth, records
from tests import util
from tests.test_oracle_chain import chain_case, tie_case, recombinant_case, overlap_case, lc_case

pytestmark = pytest.mark.gpu


def _gpu_chain(prefix, s1, exhaustive=0, minlen=16, mrs=0.5, coverT=0.1, mrc=0.0, lc=0):
    db = api.TemplateDB(prefix, device=0)
    p = api.default_params()
    p.kmerscan = 1
    p.lc = lc
    p.exhaustive = exhaustive
    p.minlen = minlen
    p.scoreT = mrs
    p.coverT = coverT
    p.mrc = mrc
    out, n, st = db.save_kmers_batch(s1, p)
    db.close()
    return out.tobytes() + api.stream_terminator(n), st


def _first_diff(a: bytes, b: bytes) -> str:
    """which record differs (for the assertion message)"""
    ra, rb = records.parse_stage2(np.frombuffer(a, np.uint8)), records.parse_stage2(np.frombuffer(b, np.uint8))
    brief = lambda x: (x["name"][:12], x["seqlen"], x["score"], x["templates"].tolist(), x["name"][-8:].hex(), x["flag"],
                       hash(x["seq"].tobytes()), x["N"].tolist()[:4])
    for i, (x, y) in enumerate(zip(ra, rb)):
        if brief(x) != brief(y):
            return f"record {i}: got {brief(x)} want {brief(y)}"
    return f"{len(ra)} vs {len(rb)} records"


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,err,n_rate", [(11, 0.10, 0.0), (12, 0.03, 0.0), (13, 0.10, 0.002), (14, 0.0, 0.001)])
def test_chain_matches_reference(tmp_path, seed, err, n_rate):
    prefix, s1, s2 = chain_case(tmp_path, seed, 120, 1000, 6000, err, n_rate)
    stats = {}
    want = util.oracle_chain_stream(prefix, s1, stats=stats).tobytes()
    assert want == s2
    got, st = _gpu_chain(prefix, s1)
    assert st.launches > 0
    assert got == s2, _first_diff(got, s2)
    for key in ("reads", "mapped", "lookups", "hits", "list_fetches", "list_ids"):
        assert getattr(st, key) == stats[key], key


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
def test_chain_exhaustive_minlen_mrc(tmp_path):
    prefix, s1, s2 = chain_case(tmp_path, 21, 60, 500, 3000, 0.08, extra=["-ex_mode", "-ml", "100"])
    got, _ = _gpu_chain(prefix, s1, exhaustive=1, minlen=100)
    assert got == s2, _first_diff(got, s2)
    d2 = tmp_path / "b"
    d2.mkdir()
    prefix, s1, s2 = chain_case(d2, 22, 60, 300, 2500, 0.05, extra=["-mrc", "0.7"])
    got, _ = _gpu_chain(prefix, s1, mrc=0.7)
    assert got == s2, _first_diff(got, s2)


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", [31, 32])
def test_chain_ties_and_both_strands(tmp_path, seed):
    prefix, s1, s2 = tie_case(tmp_path, seed)
    got, _ = _gpu_chain(prefix, s1)
    assert got == s2, _first_diff(got, s2)


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
def test_chain_recombinant_reads(tmp_path):
    prefix, s1, s2 = recombinant_case(tmp_path, 77)
    got, _ = _gpu_chain(prefix, s1)
    assert got == s2, _first_diff(got, s2)


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", [51, 52, 53])
def test_chain_length_corrected(tmp_path, seed):
    """-lc (kma.c:694-700): ankerScoreLen / testExtensionScoreLen / proxiTestBestScoreLen / getBestAnkerScoreLen /
    getTieAnkerScoreLen and the swap of savekmers.c:5657 vs `kma -s2 -lc`"""
    prefix, s1, s2, s2_plain = lc_case(tmp_path, seed)
    got, _ = _gpu_chain(prefix, s1, lc=1)
    assert got == s2, _first_diff(got, s2)
    got, _ = _gpu_chain(prefix, s1)
    assert got == s2_plain, _first_diff(got, s2_plain)


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("shape", ["chain", "recombinant", "overlap", "tie"])
def test_chain_length_corrected_other_shapes(tmp_path, shape):
    if shape == "chain":
        prefix, s1, _ = chain_case(tmp_path, 13, 120, 1000, 6000, 0.10, 0.002)
        kw, extra = {}, []
    elif shape == "recombinant":
        prefix, s1, _ = recombinant_case(tmp_path, 77, 3000)
        kw, extra = {}, []
    elif shape == "overlap":
        prefix, s1, _ = overlap_case(tmp_path, 42, 0.5)
        kw, extra = {"coverT": 0.5}, ["-mct", "0.5"]
    else:
        prefix, s1, _ = tie_case(tmp_path, 31)
        kw, extra = {}, []
    s2 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-s2", "-lc"] + extra, cwd=tmp_path)
    assert util.oracle_chain_stream(prefix, s1, lc=1, **kw).tobytes() == s2
    got, _ = _gpu_chain(prefix, s1, lc=1, **kw)
    assert got == s2, _first_diff(got, s2)


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,mct", [(41, 0.1), (42, 0.5), (43, 0.9)])
def test_chain_overlapping_regions(tmp_path, seed, mct):
    prefix, s1, s2 = overlap_case(tmp_path, seed, mct)
    got, _ = _gpu_chain(prefix, s1, coverT=mct)
    assert got == s2, _first_diff(got, s2)


def test_chain_long_reads_vs_oracle(tmp_path):
    """C3-shaped reads (5-20 kb, 10 % errors, several genes per read) against the oracle alone (no reference binary
    needed): many ankers per strand, several regions per read, multi-chunk anker scans."""
    names, seqs = synth.gene_db(5, n_families=40, n_variants=8, len_lo=500, len_hi=3000)
    prefix = str(tmp_path / "db")
    from kma_b200 import dbbuild
    dbbuild.build_db(prefix, names, seqs, k=16)
    reads = synth.long_reads(6, seqs, 300, len_lo=5000, len_hi=20000, err=0.10)
    s1 = records.stage1_records(reads)
    want = util.oracle_chain_stream(prefix, s1).tobytes()
    got, st = _gpu_chain(prefix, s1)
    assert st.mapped > 200
    assert got == want, _first_diff(got, want)


def _chain_params():
    p = api.default_params()
    p.kmerscan = 1
    return p


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case,seed,err,n_rate", [("chain", 51, 0.10, 0.0), ("chain", 52, 0.04, 0.002), ("tie", 53, 0, 0)])
def test_chain_records_through_alignment(tmp_path, case, seed, err, n_rate):
    """Stage-2 records of chain mode carry query bounds (qseqs.c:41); the alignment pass restricts its seed scans with
    them (alnfrags.c:1091-1099). CUDA vs oracle (pinned to alnFrags_threaded in tests/test_oracle_align.py), from the
    reference's stream and chained in HBM behind the CUDA chain kernel."""
    if case == "chain":
        prefix, s1, s2 = chain_case(tmp_path, seed, 80, 1000, 5000, err, n_rate)
    else:
        prefix, s1, s2 = tie_case(tmp_path, seed)
    s2 = np.frombuffer(s2, dtype=np.uint8)
    ofrag, oa, ou, ocand, cells = util.oracle_align_stream(prefix, s2, one2one=False)
    db = api.TemplateDB(prefix, device=0)
    p = _chain_params()
    frag, a, u, cand, st = db.alnFrags_batch(s2, p, want_cand=True)
    assert st.tasks == len(ocand)
    if not util.cand_equal(cand, ocand):
        bad = np.flatnonzero(((cand != ocand) & ~((np.arange(8) == 5) & (ocand[:, 2:3] == 0))).any(axis=1))
        raise AssertionError(f"{len(bad)} of {len(cand)} candidate rows differ; first: got {cand[bad[0]]} want {ocand[bad[0]]}")
    assert np.array_equal(a, oa) and np.array_equal(u, ou)
    assert frag.tobytes() == ofrag
    # the same, stage 2 and the alignment pass chained in HBM
    db.seed_upload(s1)
    sst = db.seed_run(p)
    assert sst.mapped > 0
    db.align_from_seed()
    st2 = db.align_run(p, want_cand=True)
    frag2, a2, u2, cand2 = db.align_download(want_cand=True)
    db.close()
    assert st2.tasks == len(ocand)
    assert util.cand_equal(cand2, ocand)
    assert np.array_equal(a2, oa) and np.array_equal(u2, ou)
    assert frag2.tobytes() == ofrag


def test_chain_golden_reads_with_early_n():
    """the golden 150 bp reads through chain mode; 98 of them carry an N in their first k bases, where the reference
    reads past its reverse-complement buffer (undefined): the CUDA path follows the oracle's zero-bits convention"""
    with util.golden_dir() as g:
        s1 = np.fromfile(f"{g}/s1.bin", dtype=np.uint8)
        want = util.oracle_chain_stream(f"{g}/db", s1).tobytes()
        got, st = _gpu_chain(f"{g}/db", s1)
    assert st.mapped > 2000
    assert got == want, _first_diff(got, want)
