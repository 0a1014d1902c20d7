"""CPU tests: the oracle restatement of the traceback alignment of the assembly pass (assemble_KMA's inner loop:
anker_rc + KMA with aligned rows, NW / NW_band with strings, lead/trail trimming, acceptance) is pinned byte-exact to
the unmodified reference driven by oracle/ref_harness.c -trace."""
import numpy as np
import pytest

from kma_b200 import synth
from tests import util

pytestmark = pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")


def make_frags(tmp_path, seed, L, sub, indel, n=400):
    names, seqs = synth.gene_db(seed, n_families=12, n_variants=6, len_lo=max(300, L + 50), len_hi=max(1500, 2 * L))
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    rng = np.random.default_rng(seed)
    base = synth.short_reads(seed + 1, seqs, n, L=L, sub=0.0, n_rate=0.0, junk_frac=0.03)
    reads = [synth.mutate_indel(rng, r, sub, indel / 2, indel / 2) for r in base]
    for r in reads[::9]:
        if len(r) > 40:
            r[rng.integers(0, len(r), size=2)] = 4
    synth.write_fastq(tmp_path / "r.fq", reads)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    s2 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-1t1", "-s2"], cwd=tmp_path)
    frag, _, _, _ = util.ref_align(str(tmp_path / "db"), s2, str(tmp_path), cand=False)
    return str(tmp_path / "db"), util.assembly_records(frag)


@pytest.mark.parametrize("seed,L,sub,indel", [(31, 150, 0.01, 0.0), (32, 150, 0.03, 0.01), (33, 400, 0.05, 0.02),
                                               (34, 1000, 0.04, 0.03), (35, 3000, 0.08, 0.06)])
def test_trace_vs_reference(tmp_path, seed, L, sub, indel):
    prefix, frags = make_frags(tmp_path, seed, L, sub, indel)
    want = util.ref_trace(prefix, frags, str(tmp_path))
    got = util.oracle_trace(prefix, frags)
    assert got == want
    recs = util.parse_trace(want)
    assert len(recs) > 300 and sum(int(h[0]) for h, _ in recs) > 300     # accepted alignments
    assert sum(int(h[10]) for h, _ in recs) > 10                          # reads anker_rc had to turn around
    if indel:
        assert any(b"_" in rows[1] for _, rows in recs)


@pytest.mark.parametrize("seed,L,sub,indel,ts", [(32, 150, 0.03, 0.01, 2), (33, 400, 0.05, 0.02, 2), (34, 1000, 0.04, 0.03, 3), (35, 3000, 0.08, 0.06, 2),
                                                  (36, 600, 0.06, 0.04, 40)])
def test_trace_with_seed_trimming_vs_reference(tmp_path, seed, L, sub, indel, ts):
    """-ts: trimSeeds (chain.c:496-538) through the reference's own KMA (ref_harness -trace -ts); seeds shorter than ts keep one base"""
    prefix, frags = make_frags(tmp_path, seed, L, sub, indel)
    want = util.ref_trace(prefix, frags, str(tmp_path), ts=ts)
    assert want != util.ref_trace(prefix, frags, str(tmp_path))
    assert util.oracle_trace(prefix, frags, ts=ts) == want


@pytest.mark.parametrize("seed,L,sub,indel,mode", [(41, 150, 0.02, 0.02, "sparse"), (42, 150, 0.02, 0.02, "dense"),
                                                    (43, 1000, 0.04, 0.04, "sparse"), (44, 1000, 0.04, 0.04, "dense")])
def test_matrix_counts_vs_reference(tmp_path, seed, L, sub, indel, mode):
    """the per-position base counts alnToMat (template nodes) / alnToMatDense add for every accepted alignment: oracle
    restatement vs the reference's own functions run by ref_harness -trace -mat"""
    prefix, frags = make_frags(tmp_path, seed, L, sub, indel, n=900)
    trace, mats, nodes = util.ref_trace(prefix, frags, str(tmp_path), matrix=mode)
    assert util.oracle_trace(prefix, frags) == trace
    got = util.oracle_matrix(prefix, frags, trace, dense=mode == "dense")
    off = util.matrix_offsets(prefix)
    assert len(mats) > 20
    seen = np.zeros(len(got), dtype=bool)
    for t, m in mats.items():
        assert np.array_equal(got[off[t]:off[t] + len(m)], m), f"template {t}"
        seen[off[t]:off[t] + len(m)] = True
    assert not got[~seen].any()
    assert int(got[:, 5].sum()) > 0                     # deletions were counted
    if mode == "sparse":
        assert any(nodes[t] > len(m) for t, m in mats.items())   # reads with insertions grew the reference's node list
