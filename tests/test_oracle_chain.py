"""CPU tests: the oracle restatement of stage 2 in chain mode (save_kmers_chain, the long-read default) is pinned
byte for byte to the unmodified reference (`kma -s2` without -1t1)."""
import numpy as np
import pytest

from kma_b200 import synth
from tests import util


def chain_case(tmp_path, seed, n_reads, len_lo, len_hi, err, n_rate=0.0, short=True, extra=()):
    names, seqs = synth.gene_db(seed, n_families=20, n_variants=6, len_lo=300, len_hi=1500)
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    rng = np.random.default_rng(seed + 1)
    reads = synth.long_reads(seed + 2, seqs, n_reads, len_lo=len_lo, len_hi=len_hi, err=err)
    if short:    # short and clean reads go through the same code path
        reads += list(synth.short_reads(seed + 3, seqs, 150, L=150, sub=0.01, junk_frac=0.1))
        reads += list(synth.short_reads(seed + 4, seqs, 40, L=40, sub=0.0))
        reads += [np.zeros(60, dtype=np.uint8), rng.integers(0, 4, size=15).astype(np.uint8)]
    if n_rate:   # N's anywhere but the first k bases (the reference reads past the read there, see orc_chain.c)
        for r in reads:
            hit = np.flatnonzero(rng.random(len(r)) < n_rate)
            r[hit[hit >= 16]] = 4
    order = rng.permutation(len(reads))
    reads = [reads[i] for i in order]
    synth.write_fastq(tmp_path / "r.fq", reads, qual="5")
    s1 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-s1"] + list(extra), cwd=tmp_path)
    s2 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-s2"] + list(extra), cwd=tmp_path)
    return str(tmp_path / "db"), np.frombuffer(s1, dtype=np.uint8), s2


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,err,n_rate", [(11, 0.10, 0.0), (12, 0.03, 0.0), (13, 0.10, 0.002), (14, 0.0, 0.001)])
def test_chain_oracle_matches_live_reference(tmp_path, seed, err, n_rate):
    prefix, s1, s2 = chain_case(tmp_path, seed, 120, 1000, 6000, err, n_rate)
    got = util.oracle_chain_stream(prefix, s1)
    assert len(s2) > 1000
    assert got.tobytes() == s2


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
def test_chain_oracle_exhaustive_and_minlen(tmp_path):
    prefix, s1, s2 = chain_case(tmp_path, 21, 60, 500, 3000, 0.08, extra=["-ex_mode", "-ml", "100"])
    got = util.oracle_chain_stream(prefix, s1, exhaustive=1, minlen=100)
    assert got.tobytes() == s2


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
def test_chain_oracle_mrc(tmp_path):
    """-mrc reaches mrchain (kmeranker.c:57); its test `q_len < mrc * maplen` cannot hold for mrc <= 1 (maplen <= q_len)."""
    prefix, s1, s2 = chain_case(tmp_path, 22, 60, 300, 2500, 0.05, extra=["-mrc", "0.7"])
    got = util.oracle_chain_stream(prefix, s1, mrc=0.7)
    assert got.tobytes() == s2


def tie_case(tmp_path, seed):
    """Templates present on both strands (strand ties, rc == 3) and reads that carry the same gene two or three times
    without errors (equal-scoring ankers: the `ties` path of save_kmers_chain, savekmers.c:5701-5781)."""
    names, seqs = synth.gene_db(seed, n_families=12, n_variants=4, len_lo=300, len_hi=900)
    for i in range(0, len(seqs), 5):
        names.append(names[i] + "_rc")
        seqs.append(synth.revcomp(seqs[i]))
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    rng = np.random.default_rng(seed)
    reads = []
    for _ in range(150):
        t = seqs[int(rng.integers(0, len(seqs)))]
        reps = int(rng.integers(2, 4))
        parts = []
        for r in range(reps):
            parts.append(t if rng.random() < 0.7 else synth.mutate_subs(rng, t, 0.01))
            parts.append(rng.integers(0, 4, size=int(rng.integers(0, 120))).astype(np.uint8))
        s = np.concatenate(parts)
        reads.append(synth.revcomp(s) if rng.random() < 0.5 else s)
    reads += list(synth.short_reads(seed + 3, seqs, 200, L=200, sub=0.0))
    synth.write_fastq(tmp_path / "r.fq", reads, qual="5")
    s1 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-s1"], cwd=tmp_path)
    s2 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-s2"], cwd=tmp_path)
    return str(tmp_path / "db"), np.frombuffer(s1, dtype=np.uint8), s2


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", [31, 32])
def test_chain_oracle_ties_and_both_strands(tmp_path, seed):
    prefix, s1, s2 = tie_case(tmp_path, seed)
    scores = []
    ip = 0
    while ip + 28 <= len(s2):
        h = np.frombuffer(s2, dtype=np.int32, count=7, offset=ip)
        if h[0] < 0:
            break
        scores.append(int(h[3]))
        ip += 28 + 8 * int(h[1]) + 4 * int(h[2]) + 4 * int(h[4]) + int(h[5])
    assert min(scores) < 0, "the case is meant to produce strand ties"
    got = util.oracle_chain_stream(prefix, s1)
    assert got.tobytes() == s2


def recombinant_case(tmp_path, seed, n_reads=6000):
    """Reads spliced from two close variants of a family: alternating template lists inside one chain, which is where
    equal-scoring ankers (getTieAnkerScore returning an anker) turn up."""
    names, seqs = synth.gene_db(seed, n_families=6, n_variants=12, len_lo=400, len_hi=900)
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    rng = np.random.default_rng(seed)
    reads = []
    for _ in range(n_reads):
        f = int(rng.integers(0, 6))
        a, b = rng.integers(0, 12, size=2)
        A, B = seqs[f * 12 + a], seqs[f * 12 + b]
        L = min(len(A), len(B))
        x = int(rng.integers(50, L - 50))
        s = np.concatenate([A[:x], B[x:L]])
        lo = int(rng.integers(0, L - 120))
        hi = int(rng.integers(lo + 100, L + 1))
        s = synth.mutate_subs(rng, s[lo:hi], 0.01)
        reads.append(synth.revcomp(s) if rng.random() < 0.5 else s)
    synth.write_fastq(tmp_path / "r.fq", reads, qual="5")
    s1 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-s1"], cwd=tmp_path)
    s2 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-s2"], cwd=tmp_path)
    return str(tmp_path / "db"), np.frombuffer(s1, dtype=np.uint8), s2


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
def test_chain_oracle_recombinant_reads(tmp_path):
    prefix, s1, s2 = recombinant_case(tmp_path, 77)
    got = util.oracle_chain_stream(prefix, s1)
    assert got.tobytes() == s2


def overlap_case(tmp_path, seed, mct):
    """Regions that overlap on the read: templates that share their ends with their neighbours (segment-tree extend /
    split / containment) and templates stored on the other strand with an offset (chooseChain's partial overlaps)."""
    rng = np.random.default_rng(seed)
    rnd = lambda n: rng.integers(0, 4, size=int(n)).astype(np.uint8)
    names, seqs, reads = [], [], []
    for g in range(40):
        A, B, Cc = rnd(rng.integers(300, 700)), rnd(rng.integers(120, 400)), rnd(rng.integers(300, 700))
        ov = int(rng.integers(5, 60))
        names += [f"g{g}_A", f"g{g}_B", f"g{g}_C"]
        seqs += [A, np.concatenate([A[-ov:], B, Cc[:ov]]), Cc]
        X, Y = rnd(rng.integers(400, 800)), rnd(rng.integers(200, 500))
        cut = int(rng.integers(100, len(X) - 100))
        names += [f"g{g}_X", f"g{g}_Yrc"]
        seqs += [X, synth.revcomp(np.concatenate([X[cut:], Y]))]
        for _ in range(6):
            r = synth.mutate_subs(rng, np.concatenate([rnd(rng.integers(0, 30)), A, B, Cc, rnd(rng.integers(0, 30))]), float(rng.choice([0, 0.01, 0.03])))
            reads.append(synth.revcomp(r) if rng.random() < 0.5 else r)
            r = synth.mutate_subs(rng, np.concatenate([rnd(rng.integers(1, 30)), X, Y[:int(rng.integers(50, len(Y) + 1))]]), float(rng.choice([0, 0, 0.01])))
            reads.append(synth.revcomp(r) if rng.random() < 0.5 else r)
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    synth.write_fastq(tmp_path / "r.fq", reads, qual="5")
    extra = ["-mct", str(mct)]
    s1 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-s1"] + extra, cwd=tmp_path)
    s2 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-s2"] + extra, cwd=tmp_path)
    return str(tmp_path / "db"), np.frombuffer(s1, dtype=np.uint8), s2


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,mct", [(41, 0.1), (42, 0.5), (43, 0.9)])
def test_chain_oracle_overlapping_regions(tmp_path, seed, mct):
    prefix, s1, s2 = overlap_case(tmp_path, seed, mct)
    got = util.oracle_chain_stream(prefix, s1, coverT=mct)
    assert got.tobytes() == s2


def lc_case(tmp_path, seed):
    """-lc (kma.c:694-700): families whose variants differ in length (truncated and extended copies) so that the
    length-corrected anker score picks other ankers / templates than the plain one, plus the tie and recombinant shapes."""
    rng = np.random.default_rng(seed)
    names, seqs = synth.gene_db(seed, n_families=10, n_variants=5, len_lo=300, len_hi=1200)
    for i in range(len(seqs)):
        s = seqs[i]
        r = rng.random()
        if r < 0.3:
            names.append(names[i] + "_cut"); seqs.append(s[int(rng.integers(0, 80)): len(s) - int(rng.integers(0, 200))].copy())
        elif r < 0.6:
            names.append(names[i] + "_ext")
            seqs.append(np.concatenate([rng.integers(0, 4, size=int(rng.integers(20, 300))).astype(np.uint8), s,
                                        rng.integers(0, 4, size=int(rng.integers(20, 300))).astype(np.uint8)]))
    synth.write_fasta(tmp_path / "db.fsa", names, seqs)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp_path)
    reads = synth.long_reads(seed + 2, seqs, 150, len_lo=200, len_hi=3000, err=0.05)
    reads += list(synth.short_reads(seed + 3, seqs, 300, L=150, sub=0.01, junk_frac=0.05))
    for _ in range(150):
        t = seqs[int(rng.integers(0, len(seqs)))]
        parts = []
        for r in range(int(rng.integers(1, 4))):
            parts.append(t if rng.random() < 0.7 else synth.mutate_subs(rng, t, 0.01))
            parts.append(rng.integers(0, 4, size=int(rng.integers(0, 120))).astype(np.uint8))
        s = np.concatenate(parts)
        reads.append(synth.revcomp(s) if rng.random() < 0.5 else s)
    synth.write_fastq(tmp_path / "r.fq", reads, qual="5")
    s1 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-s1", "-lc"], cwd=tmp_path)
    s2 = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-s2", "-lc"], cwd=tmp_path)
    s2_plain = util.ref_kma(["-i", "r.fq", "-o", "o", "-t_db", "db", "-s2"], cwd=tmp_path)
    return str(tmp_path / "db"), np.frombuffer(s1, dtype=np.uint8), s2, s2_plain


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", [51, 52, 53])
def test_chain_oracle_length_corrected(tmp_path, seed):
    prefix, s1, s2, s2_plain = lc_case(tmp_path, seed)
    assert s2 != s2_plain, "the case is meant to make -lc choose differently"
    got = util.oracle_chain_stream(prefix, s1, lc=1)
    assert got.tobytes() == s2
    assert util.oracle_chain_stream(prefix, s1).tobytes() == s2_plain
