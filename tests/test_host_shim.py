"""File-level drop-in (north star): the reference's own host with the mapping core swapped for libkmagpu.so
(host/kmagpu_shim.c linked over the UNMODIFIED reference objects with --wrap, oracle/Makefile.host) must write the same
.res / .fsa / .aln / .frag.gz / .mat.gz as `kma -t 1` on reduced shapes of BASELINE.json's configs.

  * `-m gpu`: oracle/_ref/kma_gpu (links kma_b200/libkmagpu.so, built here where /root/reference exists; travels with gpurun);
  * `-m "not gpu"`: oracle/_ref/kma_gpu_mock -- the same shim over tests/mock/mock_kmagpu.c (the C ABI answered by the CPU
    oracle), which checks the shim's host logic (chunking, pipes, the KMA / anker_rc result table, hand-over to the
    reference's writers) in a container without a GPU.
.res: integers equal, floats within 1e-9 relative; every other file byte-equal after gunzip."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from kma_b200 import synth
from tests import util

REF = os.path.join(util.ROOT, "oracle", "_ref")


def _build_host():
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-f", "Makefile.host"], cwd=os.path.join(util.ROOT, "oracle"),
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def _run(binary, args, cwd, env=None):
    r = subprocess.run([os.path.join(REF, binary)] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    assert r.returncode == 0, f"{binary} {' '.join(args)} -> {r.returncode}\n{r.stderr.decode()[-3000:]}"


def _read(path):
    if path.endswith(".gz"):
        with gzip.open(path, "rb") as f:
            return f.read()
    with open(path, "rb") as f:
        return f.read()


def _same_res(a: bytes, b: bytes):
    la, lb = a.decode().splitlines(), b.decode().splitlines()
    assert len(la) == len(lb), f".res: {len(la)} vs {len(lb)} lines"
    for x, y in zip(la, lb):
        fx, fy = x.split("\t"), y.split("\t")
        assert len(fx) == len(fy), (x, y)
        for u, v in zip(fx, fy):
            u, v = u.strip(), v.strip()
            if u == v:
                continue
            try:
                iu, iv = int(u), int(v)
                assert iu == iv, (x, y)
            except ValueError:
                du, dv = float(u), float(v)
                assert abs(du - dv) <= 1e-9 * max(abs(du), abs(dv)), (x, y)


def _compare(cwd, want="ref", got="gpu", exts=("res", "fsa", "aln", "frag.gz", "mat.gz"), sort_frag=False, at_least=3):
    n = 0
    for e in exts:
        pw, pg = os.path.join(cwd, f"{want}.{e}"), os.path.join(cwd, f"{got}.{e}")
        assert os.path.exists(pw) == os.path.exists(pg), e
        if not os.path.exists(pw):
            continue
        a, b = _read(pw), _read(pg)
        if e == "res":
            _same_res(a, b)
        elif e == "frag.gz" and sort_frag:
            assert sorted(a.splitlines()) == sorted(b.splitlines()), e
        else:
            assert a == b, f".{e} differs ({len(a)} vs {len(b)} bytes)"
        n += len(a) > 0
    assert n >= at_least, "nothing to compare"


def _gene_case(tmp, seed=5, fam=12, var=5):
    names, seqs = synth.gene_db(seed, n_families=fam, n_variants=var, len_lo=400, len_hi=1500)
    synth.write_fasta(tmp / "db.fsa", names, seqs)
    util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp)
    return names, seqs


CASES = {
    # C1: short single-end reads, -1t1
    "c1_se_1t1": dict(reads="se", n=1500, args=["-1t1", "-matrix"]),
    # C1 shape through the default k-mer scan (save_kmers_chain) with reads carrying N's and unmappable reads
    "c1_se_chain": dict(reads="se", n=1200, args=["-matrix"], n_rate=0.01),
    # C2: read pairs, penalty pairing
    "c2_pe_apm_p": dict(reads="pe", n=900, args=["-apm", "p", "-matrix"]),
    # C2 with the default pairing (-apm u) and -1t1
    "c2_pe_apm_u": dict(reads="pe", n=700, args=["-1t1"]),
    # C3: Nanopore-like long reads, chain mode, -bcNano -bc 0.7 (insertion nodes all over the matrix)
    "c3_long_nano": dict(reads="long", n=40, args=["-bcNano", "-bc", "0.7", "-matrix"]),
    # C4: one genome, -mem_mode -1t1, consensus + base counts (the alignment happens in the assembly only)
    "c4_genome_mem": dict(reads="genome", n=3000, args=["-mem_mode", "-1t1", "-matrix"]),
    # seed trimming of the traceback alignment (-ts, bound by the -ont / -ill presets)
    "c1_se_ts2": dict(reads="se", n=800, args=["-1t1", "-ts", "2", "-matrix"]),
    # -lc with -1t1: runConClave_lc alone (the anker selection of -lc only exists in the chain scan)
    "c1_se_lc_1t1": dict(reads="se", n=800, args=["-1t1", "-lc", "-matrix"]),
    # -lc with the chain scan: length-corrected anker selection (kma.c:694-700) + runConClave_lc, short and long reads
    "c1_se_lc_chain": dict(reads="se", n=800, args=["-lc", "-matrix"], n_rate=0.005),
    "c3_long_lc": dict(reads="long", n=40, args=["-lc", "-bcNano", "-bc", "0.7"]),
    # -proxi: proximity scoring in stage 2 (getProxiMatch / getSecondProxiPen / getF_Proxi / getR_Proxi /
    # getProxiChainTemplates) and the minFrac branches of update_Scores* in stage 3
    "c1_se_proxi": dict(reads="se", n=800, args=["-1t1", "-proxi", "0.9", "-matrix"]),
    "c2_pe_proxi_p": dict(reads="pe", n=600, args=["-apm", "p", "-proxi", "-0.9"]),
    "c2_pe_proxi_u": dict(reads="pe", n=600, args=["-proxi", "0.8", "-1t1"]),
    "c3_long_proxi": dict(reads="long", n=40, args=["-proxi", "0.9"]),
    # the reference's presets (kma.c:1100-1240): -ont = chain scan, -lc, -proxi -0.9, -ts 2, -eq 10, -mrs 0.25, -mrc 0.7, -bcNano -bc 0.7;
    # -ill = -1t1, -lc, -proxi -0.98, -mrc 0.1, -bc 0.9 -bcd 10; -asm = -lc, -proxi -0.9, -ts 2, -bc 0.5, -mrs 0.25, -mrc 0.7
    "preset_ont": dict(reads="long", n=40, args=["-ont", "-matrix"]),
    "preset_ill_se": dict(reads="se", n=800, args=["-ill", "-matrix"]),
    "preset_ill_pe": dict(reads="pe", n=600, args=["-ill"]),
    "preset_asm": dict(reads="long", n=30, args=["-asm"]),
    # options that stay in the reference's own host code around the device calls: ConClave version 2 (runConClave2,
    # conclave.c:386), dense base counts, reference-guided consensus, stage-1 quality filters (-eq / -mi / -mp / -5p)
    # soft proximity together with -mem_mode: stage 2 collects the softProxi sums (kmers.c:133-153), they travel behind the
    # stream and become runKMA_MEM's alignment_scores (runkma.c:1153) -- the -ill and -ont presets on a genome in -mem_mode
    "c4_genome_mem_ill": dict(reads="genome", n=3000, args=["-mem_mode", "-ill", "-matrix"]),
    "c4_genome_mem_soft_chain": dict(reads="genome", n=2000, args=["-mem_mode", "-proxi", "-0.9"]),
    "c1_se_mem_soft": dict(reads="se", n=800, args=["-mem_mode", "-1t1", "-proxi", "-0.9", "-matrix"]),
    "c2_pe_mem_soft_u": dict(reads="pe", n=600, args=["-mem_mode", "-proxi", "-0.8"]),
    "c2_pe_mem_soft_p": dict(reads="pe", n=600, args=["-mem_mode", "-apm", "p", "-proxi", "-0.9"]),
    "c3_long_mem_ont": dict(reads="long", n=40, args=["-mem_mode", "-ont"]),
    "c1_se_conclave2": dict(reads="se", n=800, args=["-1t1", "-ConClave", "2", "-matrix"]),
    "c3_long_conclave2_lc": dict(reads="long", n=40, args=["-ConClave", "2", "-lc"]),
    "c1_se_dense_reffsa": dict(reads="se", n=800, args=["-1t1", "-dense", "-ref_fsa", "-matrix"]),
    "c1_se_quality_filters": dict(reads="se", n=800, args=["-1t1", "-eq", "15", "-mp", "25", "-mi", "40", "-5p", "3"], n_rate=0.01),
}


# more of the reference's options around the device calls, over the mock ABI only (CPU suite): other penalties (-cge sets
# MM -3, W1 -5, PE 17), base callers and significance tests, -and, output switches
CASES_HOST = {
    "c1_se_cge": dict(reads="se", n=600, args=["-1t1", "-cge", "-matrix"]),
    "c2_pe_cge": dict(reads="pe", n=500, args=["-cge", "-apm", "p"]),
    "c1_se_bc90_and": dict(reads="se", n=600, args=["-1t1", "-bc90", "-and", "-mrs", "0.3"]),
    "c1_se_bcg_nc_nf": dict(reads="se", n=600, args=["-1t1", "-bcg", "-nc", "-nf"]),
    "c3_long_mrc_mct": dict(reads="long", n=30, args=["-mrc", "0.5", "-mct", "0.3", "-ml", "60"]),
    "c1_se_exhaustive": dict(reads="se", n=600, args=["-1t1", "-ex_mode", "-matrix"], n_rate=0.01),
}


def _make_case(tmp, name):
    c = CASES.get(name) or CASES_HOST[name]
    if c["reads"] == "genome":
        names, seqs = synth.genome_db(4, length=60000)
        synth.write_fasta(tmp / "db.fsa", names, seqs)
        util.ref_kma(["index", "-i", "db.fsa", "-o", "db"], cwd=tmp)
        r = synth.short_reads(14, seqs, c["n"], L=150, sub=0.01, n_rate=0.002)
        synth.write_fastq(tmp / "r.fq", list(r))
        return ["-i", "r.fq", "-t_db", "db"] + c["args"]
    names, seqs = _gene_case(tmp)
    if c["reads"] == "se":
        r = synth.short_reads(11, seqs, c["n"], L=150, sub=0.01, n_rate=c.get("n_rate", 0.0))
        synth.write_fastq(tmp / "r.fq", list(r))
        inp = ["-i", "r.fq"]
    elif c["reads"] == "pe":
        r1, r2 = synth.paired_reads(12, seqs, c["n"], L=150, sub=0.01)
        synth.write_fastq(tmp / "r1.fq", list(r1))
        synth.write_fastq(tmp / "r2.fq", list(r2))
        inp = ["-ipe", "r1.fq", "r2.fq"]
    else:
        r = synth.long_reads(13, seqs, c["n"], len_lo=1500, len_hi=5000)
        synth.write_fastq(tmp / "r.fq", r, qual="5")
        inp = ["-i", "r.fq"]
    return inp + ["-t_db", "db"] + c["args"]


@pytest.mark.parametrize("name", list(CASES))
def test_shim_host_logic_with_the_oracle_behind_the_abi(tmp_path, name):
    if not os.path.isdir("/root/reference") and not os.path.exists(os.path.join(REF, "kma_gpu_mock")):
        pytest.skip("reference host not built in this environment")
    _build_host()
    args = _make_case(tmp_path, name)
    _run("kma", args + ["-o", "ref", "-t", "1"], tmp_path)
    _run("kma_gpu_mock", args + ["-o", "gpu", "-t", "1"], tmp_path)
    _compare(tmp_path)


@pytest.mark.parametrize("name", list(CASES_HOST))
def test_shim_host_logic_more_options(tmp_path, name):
    if not os.path.isdir("/root/reference") and not os.path.exists(os.path.join(REF, "kma_gpu_mock")):
        pytest.skip("reference host not built in this environment")
    _build_host()
    args = _make_case(tmp_path, name)
    _run("kma", args + ["-o", "ref", "-t", "1"], tmp_path)
    _run("kma_gpu_mock", args + ["-o", "gpu", "-t", "1"], tmp_path)
    exts = ("res", "fsa", "aln", "frag.gz", "mat.gz")
    if "-nc" in args or "-nf" in args:
        exts = tuple(e for e in exts if not (e in ("fsa", "aln") and "-nc" in args) and not (e == "frag.gz" and "-nf" in args))
    _compare(tmp_path, exts=exts, at_least=min(3, len(exts) - 1))


def test_shim_refuses_what_the_gpu_path_does_not_cover(tmp_path):
    if not os.path.exists(os.path.join(REF, "kma_gpu_mock")) and not os.path.isdir("/root/reference"):
        pytest.skip("reference host not built in this environment")
    _build_host()
    args = _make_case(tmp_path, "c1_se_1t1")
    r = subprocess.run([os.path.join(REF, "kma_gpu_mock")] + args + ["-o", "x", "-sam"], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0 and b"no CPU fallback" in r.stderr
    r = subprocess.run([os.path.join(REF, "kma_gpu_mock")] + args + ["-o", "y", "-ca"], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0 and b"no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_output_files_equal_the_reference(tmp_path, name):
    """the five output files of the reference host running on libkmagpu.so vs `kma -t 1`"""
    assert os.path.exists(os.path.join(REF, "kma_gpu")), "oracle/_ref/kma_gpu missing: run __graft_entry__.build() where /root/reference exists"
    args = _make_case(tmp_path, name)
    _run("kma", args + ["-o", "ref", "-t", "1"], tmp_path)
    _run("kma_gpu", args + ["-o", "gpu", "-t", "1"], tmp_path)
    _compare(tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES_HOST))
def test_output_files_more_options(tmp_path, name):
    """other penalties (-cge), base callers / significance tests, -and, -nc / -nf, -ex_mode through the device (checked by
    tools/host_options_gpu.py as well)"""
    assert os.path.exists(os.path.join(REF, "kma_gpu")), "oracle/_ref/kma_gpu missing: run __graft_entry__.build() where /root/reference exists"
    args = _make_case(tmp_path, name)
    _run("kma", args + ["-o", "ref", "-t", "1"], tmp_path)
    _run("kma_gpu", args + ["-o", "gpu", "-t", "1"], tmp_path)
    exts = ("res", "fsa", "aln", "frag.gz", "mat.gz")
    if "-nc" in args or "-nf" in args:
        exts = tuple(e for e in exts if not (e in ("fsa", "aln") and "-nc" in args) and not (e == "frag.gz" and "-nf" in args))
    _compare(tmp_path, exts=exts, at_least=min(3, len(exts) - 1))


@pytest.mark.gpu
def test_output_files_with_host_threads(tmp_path):
    """-t 3: the reference's assembly threads pull the device results out of order; only the order of .frag.gz lines may differ"""
    args = _make_case(tmp_path, "c1_se_1t1")
    _run("kma", args + ["-o", "ref", "-t", "1"], tmp_path)
    _run("kma_gpu", args + ["-o", "gpu", "-t", "3"], tmp_path)
    _compare(tmp_path, sort_frag=True)
