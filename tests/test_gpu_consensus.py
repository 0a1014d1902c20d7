"""GPU parity tests (B200): the consensus call of the assembly pass (kmagpu_consensus = callConsensus over the template
nodes of the device matrix) through the C ABI vs the oracle that tests/test_oracle_consensus.py pins to the reference's
own callConsensus; every caller / significance combination, single templates and the whole database, and the consensus
of a matrix the device itself accumulated (stage 3 end to end)."""
import ctypes as C

import numpy as np
import pytest

from kma_b200 import api, synth
from tests import util
from tests.test_oracle_consensus import CASES

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")]


def _load(db, mats, off):
    """write {template: uint16 counts} into the device matrix (zero-copy torch view of the library's buffer)"""
    import torch
    db.matrix_reset()
    dev = db.matrix_tensor()
    host = np.zeros(dev.numel(), dtype=np.int32).reshape(-1, 6)
    for t, m in mats.items():
        host[off[t]:off[t] + len(m)] = m
    dev.copy_(torch.from_numpy(host.reshape(-1)))
    torch.cuda.synchronize()


def _db(tmp_path, seed, **kw):
    names, seqs = synth.gene_db(seed, **kw)
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    return str(tmp_path / "db"), seqs


@pytest.mark.parametrize("case", range(len(CASES)))
def test_consensus_vs_oracle(tmp_path, case):
    caller, sig, support, bcd, evalue = CASES[case]
    prefix, seqs = _db(tmp_path, 170 + case, n_families=5, n_variants=4, len_lo=200, len_hi=3000)
    rng = np.random.default_rng(900 + case)
    n = len(seqs)
    mats = {t: util.random_count_matrix(rng, util.template_bases(prefix, t)) for t in range(1, n + 1) if t % 5 != 3}
    off = util.matrix_offsets(prefix)
    db = api.TemplateDB(prefix, device=0)
    _load(db, mats, off)
    kw = dict(bcd=bcd, evalue=evalue, caller=caller, significance=sig, support=support)
    t_all, s_all, q_all, st_all, ms = db.consensus(0, **kw)
    assert ms > 0 and len(t_all) == off[n + 1]
    for t in range(1, n + 1):
        m = mats.get(t, np.zeros((len(seqs[t - 1]), 6), dtype=np.uint16))
        wt, ws, wq, wst = util.oracle_consensus(prefix, t, m, bcd=bcd, evalue=evalue, caller=caller, sig=sig, support=support)
        a, b = int(off[t]), int(off[t + 1])
        assert t_all[a:b].tobytes() == wt and q_all[a:b].tobytes() == wq and s_all[a:b].tobytes() == ws, f"template {t}"
        got = st_all[t]
        assert [int(got[k]) for k in ("depth", "depthVar", "len", "aln_len", "cover")] == [int(x) for x in wst], f"template {t}"
        if t % 4 == 1:   # the single-template entry point (rows start inside a tile)
            t1, s1, q1, st1, _ = db.consensus(t, **kw)
            assert t1.tobytes() == wt and q1.tobytes() == wq and s1.tobytes() == ws
            assert [int(st1[0][k]) for k in ("depth", "depthVar", "len", "aln_len", "cover")] == [int(x) for x in wst]
    db.close()


def test_consensus_with_the_references_p_chisqr(tmp_path):
    """the threshold located over the reference's own p_chisqr (function pointer from oracle/_ref) gives the same rows"""
    prefix, seqs = _db(tmp_path, 190, n_families=3, n_variants=3, len_lo=300, len_hi=900)
    rng = np.random.default_rng(190)
    mats = {t: util.random_count_matrix(rng, util.template_bases(prefix, t)) for t in range(1, len(seqs) + 1)}
    ref = C.CDLL(util.REF_SO)
    fn = C.cast(ref.p_chisqr, C.c_void_p)
    db = api.TemplateDB(prefix, device=0)
    _load(db, mats, util.matrix_offsets(prefix))
    for ev in (0.05, 1e-4, 1e-13):   # the last one lies in the reference's p-value table (q > 49)
        a = db.consensus(0, evalue=ev, p_chisqr=fn)
        want = util.ref_consensus(prefix, mats, str(tmp_path), evalue=ev)
        o = util.matrix_offsets(prefix)
        for t, (wt, ws, wq, wst) in want.items():
            assert a[2][o[t]:o[t + 1]].tobytes() == wq and a[1][o[t]:o[t + 1]].tobytes() == ws
            assert int(a[3][t]["cover"]) == int(wst[4]) and int(a[3][t]["depthVar"]) == int(wst[1])
        if ev >= 1e-11:
            b = db.consensus(0, evalue=ev)
            assert all(np.array_equal(x, y) for x, y in zip(a[:4], b[:4]))
    with pytest.raises(api.KmaGpuError):
        db.consensus(0, evalue=1e-13)
    db.close()


def test_consensus_of_the_device_matrix(tmp_path):
    """reads -> alignment pass -> ConClave -> traceback + base counts -> consensus, all on the device; the consensus
    equals the oracle's call on the oracle chain's matrix"""
    prefix, seqs = _db(tmp_path, 181, n_families=8, n_variants=6, len_lo=300, len_hi=1200)
    rng = np.random.default_rng(181)
    reads = [synth.mutate_indel(rng, r, 0.03, 0.01, 0.01) for r in synth.short_reads(182, seqs, 4000, L=150, sub=0.0, junk_frac=0.02)]
    synth.write_fastq(tmp_path / "r.fq", reads)
    s2 = np.frombuffer(util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-1t1", "-s2"], cwd=tmp_path), dtype=np.uint8)
    ofrag, oa, ou, _, _ = util.oracle_align_stream(prefix, s2, want_cand=False)
    ofrags, _, _, _ = util.oracle_conclave(prefix, ofrag, oa, ou)
    otrace = util.oracle_trace(prefix, np.frombuffer(ofrags, dtype=np.uint8))
    omat = util.oracle_matrix(prefix, np.frombuffer(ofrags, dtype=np.uint8), otrace)
    db = api.TemplateDB(prefix, device=0)
    p = api.default_params()
    p.one2one = 1
    p.matrix = 1
    frag, a, u, _, _ = db.alnFrags_batch(s2, p)
    frags, _, _, _, _ = db.conclave_batch(frag, a, u)
    db.matrix_reset()
    db.assemble_align_batch(frags, p)
    t_all, s_all, q_all, st, _ = db.consensus(0)
    db.close()
    off = util.matrix_offsets(prefix)
    called = 0
    for t in range(1, len(seqs) + 1):
        wt, ws, wq, wst = util.oracle_consensus(prefix, t, omat[off[t]:off[t + 1]])
        assert q_all[off[t]:off[t + 1]].tobytes() == wq and s_all[off[t]:off[t + 1]].tobytes() == ws
        assert int(st[t]["depth"]) == int(wst[0]) and int(st[t]["aln_len"]) == int(wst[3])
        called += int(wst[3])
    assert called > 5000
