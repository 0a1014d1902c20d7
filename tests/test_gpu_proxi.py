"""GPU parity tests (B200): -proxi (kma.c:702-718; bound by the -ont / -ill / -asm presets) through the C ABI --
stage 2 (getProxiMatch, getSecondProxiPen, getF_Proxi / getR_Proxi, getProxiChainTemplates, chooseChain's proximity
test) byte for byte vs `kma -s2 -proxi X`, and the minFrac branches of update_Scores / _se / _pe in the alignment pass vs
the reference's own alnFrags_threaded."""
import numpy as np
import pytest

from kma_b200 import api
from tests import util
from tests.test_oracle_proxi import se_case, pairs_with, chain_with

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")]


def _stage2(prefix, s1, **kw):
    db = api.TemplateDB(prefix, device=0)
    p = api.default_params()
    for k, v in kw.items():
        setattr(p, k, v)
    out, n, st = db.save_kmers_batch(s1, p)
    db.close()
    return out.tobytes() + api.stream_terminator(n)


def _stage3(prefix, s2, **kw):
    db = api.TemplateDB(prefix, device=0)
    p = api.default_params()
    for k, v in kw.items():
        setattr(p, k, v)
    frag, a, u, _, _ = db.alnFrags_batch(np.frombuffer(s2, dtype=np.uint8), p)
    db.close()
    return frag.tobytes(), a, u


@pytest.mark.parametrize("seed,proxi", [(71, 0.9), (72, -0.98), (73, 0.5), (74, 0.0)])
def test_proxi_single_reads(tmp_path, seed, proxi):
    prefix, s1, s2, s2_plain = se_case(tmp_path, seed, proxi)
    assert _stage2(prefix, s1, one2one=1, minFrac=abs(proxi)) == s2
    assert _stage2(prefix, s1, one2one=1) == s2_plain
    frag, a, u, _ = util.ref_align(prefix, s2, str(tmp_path), one2one=True, cand=False, min_frac=proxi)
    gfrag, ga, gu = _stage3(prefix, s2, one2one=1, minFrac=proxi)
    assert gfrag == frag and np.array_equal(ga, a) and np.array_equal(gu, u)


@pytest.mark.parametrize("seed,apm,proxi", [(81, "p", 0.9), (82, "p", -0.7), (83, "u", 0.9), (84, "u", 0.6), (85, "p", 0.98), (86, "u", -0.98)])
def test_proxi_read_pairs(tmp_path, seed, apm, proxi):
    prefix, s1, s2, s2_plain = pairs_with(tmp_path, seed, apm, proxi)
    a_ = 1 if apm == "u" else 0
    assert _stage2(prefix, s1, apm=a_, minFrac=abs(proxi)) == s2
    frag, a, u, _ = util.ref_align(prefix, s2, str(tmp_path), one2one=False, cand=False, pe=apm, min_frac=proxi)
    gfrag, ga, gu = _stage3(prefix, s2, apm=a_, minFrac=proxi)
    assert gfrag == frag and np.array_equal(ga, a) and np.array_equal(gu, u)


@pytest.mark.parametrize("shape,proxi,lc", [("chain", 0.9, 0), ("chain_clean", -0.9, 0), ("recombinant", 0.9, 0), ("overlap", 0.8, 0),
                                            ("tie", 0.95, 0), ("lc", -0.9, 1), ("chain", -0.98, 1), ("recombinant", 0.7, 1)])
def test_proxi_chain_mode(tmp_path, shape, proxi, lc):
    prefix, s1, s2, kw = chain_with(tmp_path, shape, proxi, lc)
    got = _stage2(prefix, s1, kmerscan=1, lc=lc, minFrac=abs(proxi), coverT=kw.get("coverT", 0.1))
    assert got == s2


@pytest.mark.parametrize("kind", ["se", "pe_p", "pe_u", "chain", "chain_lc"])
def test_soft_proximity_sums(tmp_path, kind):
    """-proxi < 0 with -mem_mode: every template a get*Proxi* function keeps adds its score to softProxi[] (kmers.c:133-153).
    Stream and sums vs `kma -mem_mode -proxi -0.9 -s2` (the sums travel behind the stream); two batches add up"""
    from tests.test_oracle_pair import make_pairs
    from tests.test_oracle_chain import chain_case
    extra = ["-mem_mode", "-proxi", "-0.9", "-s2"]
    kw = {}
    if kind == "se":
        prefix, s1, _, _ = se_case(tmp_path, 71, 0.9, n=1500)
        s2 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-1t1"] + extra, cwd=tmp_path)
        kw = dict(one2one=1)
    elif kind in ("pe_p", "pe_u"):
        apm = kind[-1]
        prefix, s1, _ = make_pairs(tmp_path, 82, apm=apm)
        s2 = util.ref_kma(["-ipe", "r1.fq", "r2.fq", "-o", "o", "-t_db", "db", "-apm", apm] + extra, cwd=tmp_path)
        kw = dict(apm=1 if apm == "u" else 0)
    else:
        lc = kind == "chain_lc"
        prefix, s1, _ = chain_case(tmp_path, 12, 120, 1000, 6000, 0.03, 0.0)
        s2 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db"] + extra + (["-lc"] if lc else []), cwd=tmp_path)
        kw = dict(kmerscan=1, lc=int(lc))
    db = api.TemplateDB(prefix, device=0)
    DB = db.info.DB_size
    p = api.default_params()
    for k, v in kw.items():
        setattr(p, k, v)
    p.minFrac = -0.9
    with pytest.raises(api.KmaGpuError):
        db.save_kmers_batch(s1, p)          # the sums have to be started first
    db.softproxi_reset()
    out, n, _ = db.save_kmers_batch(s1, p)
    sums = db.softproxi_download()
    stream = out.tobytes() + api.stream_terminator(n)
    trailer = sums.tobytes()[:24] + sums.tobytes()
    assert len(s2) == len(stream) + 24 + 8 * DB and sums.sum() > 0
    assert stream + trailer == s2
    db.save_kmers_batch(s1, p)              # a second batch joins the same sums
    assert np.array_equal(db.softproxi_download(), 2 * sums)
    db.close()
