"""The numpy database writer produces a DB the unmodified reference maps identically with."""
import numpy as np
import pytest

from kma_b200 import synth, dbbuild
from tests import util


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
def test_reference_maps_identically_with_our_db(tmp_path):
    names, seqs = synth.gene_db(77, n_families=20, n_variants=6, len_lo=250, len_hi=900)
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    reads = synth.short_reads(78, seqs, 1500, L=120, sub=0.01, n_rate=0.001, junk_frac=0.05)
    synth.write_fastq(tmp_path / "r.fq", reads)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "ref"], cwd=tmp_path)
    info = dbbuild.build_db(str(tmp_path / "mine"), names, seqs)
    assert info["DB_size"] == len(seqs) + 1
    for f in (".length.b", ".seq.b", ".name"):
        assert (tmp_path / ("mine" + f)).read_bytes() == (tmp_path / ("ref" + f)).read_bytes(), f
    # same stage-2 candidates (template lists are equal as lists; offsets differ but are not on the wire)
    a = util.ref_kma(["-i", "r.fq", "-o", "a", "-t_db", "ref", "-1t1", "-s2"], cwd=tmp_path)
    b = util.ref_kma(["-i", "r.fq", "-o", "b", "-t_db", "mine", "-1t1", "-s2"], cwd=tmp_path)
    assert a == b
    # and the same final results through the whole reference pipeline
    util.ref_kma(["-i", "r.fq", "-o", "a", "-t_db", "ref", "-1t1", "-t", "1"], cwd=tmp_path)
    util.ref_kma(["-i", "r.fq", "-o", "b", "-t_db", "mine", "-1t1", "-t", "1"], cwd=tmp_path)
    for ext in (".res", ".fsa", ".aln"):
        assert (tmp_path / ("a" + ext)).read_bytes() == (tmp_path / ("b" + ext)).read_bytes(), ext
    # the oracle reads it too
    s1 = np.frombuffer(util.ref_kma(["-i", "r.fq", "-o", "a", "-t_db", "ref", "-1t1", "-s1"], cwd=tmp_path), dtype=np.uint8)
    assert util.oracle_seed_stream(str(tmp_path / "mine"), s1).tobytes() == a


def test_presets_mirror_the_cli():
    """api.preset(): the option bundles of kma.c:1100-1240 (behaviour checked at file level by tests/test_host_shim.py's preset cases)"""
    from kma_b200 import api
    ont, ill, asm = api.preset("ont"), api.preset("ill"), api.preset("asm")
    assert (ont["params"].kmerscan, ont["params"].lc, ont["params"].ts, ont["params"].minFrac, ont["ingest"]["min_q"]) == (1, 1, 2, -0.9, 10)
    assert (ont["consensus"]["caller"], ont["consensus"]["support"], ont["params"].scoreT, ont["params"].mrc) == (3, 0.7, 0.25, 0.7)
    assert (ill["params"].kmerscan, ill["params"].one2one, ill["params"].minFrac, ill["params"].mrc, ill["consensus"]["support"]) == (0, 1, -0.98, 0.1, 0.9)
    assert (asm["params"].ts, asm["consensus"]["evalue"], asm["consensus"]["bcd"], asm["params"].coverT) == (2, 0.5, 1, 0.1)
    import pytest
    with pytest.raises(ValueError):
        api.preset("x")
