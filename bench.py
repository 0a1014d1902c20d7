#!/usr/bin/env python
"""bench.py -- KMA mapping-core throughput on B200 (mapped reads/s + NW GCUPS), one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU program

A "step" is one pass of the hot path over one batch of synthetic reads per GPU (BASELINE.json configs[1], C2: read pairs
against the redundant gene DB, -ipe -apm p). Weak scaling: every rank maps its own batch against its own replica of the
database; the exchanges are the all-reduce of the two ConClave score arrays and of the base-count matrix (NCCL inside the
library).
  value  stage 2 (k-mer seeding + template scoring + pair selection) + stage 3a (MEM chaining + NW + alnFragsPE +
         update_Scores) with the batch resident in HBM (device events inside libkmagpu).
  e2e    the WHOLE program span from host buffers: FASTQ text of the two files in pinned host memory -> record splitter +
         stage 1 -> stage 2 -> alignment pass -> ConClave (global sums) -> traceback alignment + base counts -> consensus;
         the per-template fragment stream and the consensus rows come back to the host (what the reference's writers turn
         into .frag.gz / .res / .fsa / .aln). Every copy is inside the timed region. The reference arm runs plain
         `kma -ipe ... -o out -t <cores>` on the same reads: the same span.
  parity the numbers are only worth something if the results are the reference's: the ConClave arrays and the frag_raw
         multiset of the step's own reads against the unmodified reference, and the output files of the reference host
         running on libkmagpu.so (oracle/_ref/kma_gpu) against `kma` on a sample.
`roofline` describes the kernel with the largest share of the step; `nw` the NW kernels of the mapping path on banded
C3-shaped problems; c3 / c4 / c5 the other BASELINE configs beside the headline (c5: the k-mer table no longer fits L2).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from kma_b200 import synth, records, dbbuild  # noqa: E402

DB_SEED, READ_SEED = 42, 7
WORKLOAD = ("C2: redundant gene DB (300 families x 10 variants, 0.5-3 kb, k=16) + 2M synthetic 2x150 bp read pairs per GPU, "
            "-ipe -apm p: stage 2 (k-mer seeding + template scoring + pair selection) + stage 3 alignment pass "
            "(MEM chaining + NW + alnFragsPenaltyPE + update_Scores)")
METRIC = "mapped reads/sec (seeding + chaining + NW alignment pass)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def ncu_traffic(kernel, pairs):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json, written by
    tools/ncu_summary.py), scaled by the batch size when the capture ran another one. None when there is no capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p)).get(kernel)
    if not t or not t.get("pairs"):
        return None
    return int(t["dram_bytes"] * pairs / t["pairs"])


def make_db(workdir):
    prefix = os.path.join(workdir, "db")
    names, seqs = synth.gene_db(DB_SEED)
    if not os.path.exists(prefix + ".comp.b"):
        dbbuild.build_db(prefix + ".tmp", names, seqs)
        for ext in (".comp.b", ".length.b", ".seq.b", ".name"):
            os.replace(prefix + ".tmp" + ext, prefix + ext)
    return prefix, names, seqs


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(st, values_width):
    """exist probe per lookup, key + value offset per hit (>= 1 probe), template list per list
    fetch, packed read words. Probe chains beyond the first key are not counted (conservative)."""
    return (4 * st.lookups + 8 * st.hits + values_width * (st.list_fetches + st.list_ids) + 8 * st.read_words)


REF = os.path.join(ROOT, "oracle", "_ref")


def pair_fastq(tmp, r1, r2, tag, first=0):
    """the two FASTQ files of a set of read pairs, written once (outside every timed region)"""
    f1, f2 = os.path.join(tmp, f"{tag}_{len(r1)}_1.fq"), os.path.join(tmp, f"{tag}_{len(r1)}_2.fq")
    if not (os.path.exists(f1) and os.path.exists(f2)):
        for f, r in ((f1, r1), (f2, r2)):
            open(f + ".tmp", "wb").write(synth.fastq_fixed(np.asarray(r), first=first).tobytes())
            os.replace(f + ".tmp", f)
    return f1, f2


def ref_hotpath(prefix, inputs, cores, tmp, tag, flags=(), aln_flags=()):
    """The unmodified reference over the hot path alone: `kma ... -s2` (FASTQ parse + stage 2) piped into alnFrags_threaded
    on `cores` pthreads (oracle/ref_harness.c drives the reference's own stage-3a entry point the way runKMA does). Leaves
    the frag_raw stream and the two ConClave arrays in <tmp>/fr_<tag>.out / sc_<tag>.out. Returns seconds."""
    kma, aln = os.path.join(REF, "kma"), os.path.join(REF, "ref_aln")
    fr, sc = os.path.join(tmp, f"fr_{tag}.out"), os.path.join(tmp, f"sc_{tag}.out")
    t0 = time.perf_counter()
    with open(os.devnull, "wb") as dn:
        p1 = subprocess.Popen([kma] + list(inputs) + ["-o", os.path.join(tmp, "o_" + tag), "-t_db", prefix, "-s2", "-t", str(cores)] + list(flags),
                              stdout=subprocess.PIPE, stderr=dn)
        p2 = subprocess.Popen([aln, prefix, "-", fr, sc, "-t", str(cores)] + list(aln_flags), stdin=p1.stdout, stdout=dn, stderr=dn)
        p1.stdout.close()
        rc2 = p2.wait()
        rc1 = p1.wait()
    dt = time.perf_counter() - t0
    if rc1 or rc2:
        raise RuntimeError(f"reference run failed: kma rc={rc1}, ref_aln rc={rc2}")
    return dt, fr, sc


def ref_program(binary, prefix, inputs, cores, cwd, out, flags=(), env=None):
    """the whole reference program (or the same host on libkmagpu.so): input files -> .res / .fsa / .aln / .frag.gz. Seconds."""
    t0 = time.perf_counter()
    r = subprocess.run([os.path.join(REF, binary)] + list(inputs) + ["-o", out, "-t_db", prefix, "-t", str(cores)] + list(flags),
                       cwd=cwd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env)
    if r.returncode:
        raise RuntimeError(f"{binary} failed ({r.returncode}): {r.stderr.decode()[-500:]}")
    return time.perf_counter() - t0


def cpu_reference(prefix, r1, r2, cores, tmp):
    """Reference arm: the unmodified program on the read pairs, `kma -ipe f1 f2 -apm p -t cores -o out` -- FASTQ parse,
    stage 2, alignment pass, ConClave, assembly, consensus, writers: the span of our `e2e`. Returns reads/s (2 per pair)."""
    f1, f2 = pair_fastq(tmp, r1, r2, "sample")
    dt = ref_program("kma", prefix, ["-ipe", f1, f2], cores, tmp, os.path.join(tmp, "ref_out"), ["-apm", "p"])
    return 2 * len(r1) / dt, dt


def read_scores(path, DB):
    raw = np.fromfile(path, dtype=np.uint8)
    n = int(np.frombuffer(raw[:4].tobytes(), dtype=np.int32)[0])
    assert n == DB, "score file does not match the database"
    v = np.frombuffer(raw[4:4 + 16 * DB].tobytes(), dtype=np.uint64)
    return v[:DB].copy(), v[DB:].copy()


def frag_multiset(api, buf):
    """the frag_raw records of a stream (updatescores.c:284-295; a pair's record includes its mate block), sorted: thread
    scheduling permutes the reference's output order, the multiset is what must be equal"""
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    off = api.record_offsets(4, buf)
    mv = memoryview(buf)
    recs = [bytes(mv[int(off[i]):int(off[i + 1])]) for i in range(len(off) - 1)]
    recs.sort()
    return recs


def compare_hotpath(api, got_frag, got_a, got_u, fr_path, sc_path, DB):
    ref_a, ref_u = read_scores(sc_path, DB)
    ref_frag = np.fromfile(fr_path, dtype=np.uint8)
    t0 = time.perf_counter()
    mine, theirs = frag_multiset(api, got_frag), frag_multiset(api, ref_frag)
    return {"scores_equal": bool(np.array_equal(got_a, ref_a) and np.array_equal(got_u, ref_u)),
            "frag_sorted_equal": mine == theirs, "frag_records": len(mine), "frag_records_reference": len(theirs),
            "frag_bytes": int(len(got_frag)), "score_sum": int(ref_a.sum()), "compare_s": round(time.perf_counter() - t0, 1)}


def compare_files(cwd, want, got, exts=("res", "fsa", "aln", "frag.gz", "mat.gz")):
    """the output files of two runs: .res numbers within 1e-9 relative, everything else byte-equal after gunzip
    (.frag.gz as a sorted set of lines: the reference's threads permute them)"""
    import gzip
    out = {}
    for e in exts:
        pw, pg = os.path.join(cwd, f"{want}.{e}"), os.path.join(cwd, f"{got}.{e}")
        if not os.path.exists(pw) and not os.path.exists(pg):
            continue
        if not (os.path.exists(pw) and os.path.exists(pg)):
            out[e] = False
            continue
        rd = (lambda p: gzip.open(p, "rb").read()) if e.endswith(".gz") else (lambda p: open(p, "rb").read())
        a, b = rd(pw), rd(pg)
        if e == "res":
            ok = True
            la, lb = a.decode().splitlines(), b.decode().splitlines()
            ok = len(la) == len(lb)
            for x, y in zip(la, lb):
                fx, fy = x.split("\t"), y.split("\t")
                ok = ok and len(fx) == len(fy)
                for u, v in zip(fx, fy):
                    u, v = u.strip(), v.strip()
                    if u == v:
                        continue
                    try:
                        ok = ok and abs(float(u) - float(v)) <= 1e-9 * max(abs(float(u)), abs(float(v)))
                    except ValueError:
                        ok = False
            out[e] = ok
        elif e == "frag.gz":
            out[e] = sorted(a.splitlines()) == sorted(b.splitlines())
        else:
            out[e] = a == b
        out[e + "_bytes"] = len(a)
    return out


def parity_c2(api, prefix, r1, r2, cores, tmp, device, file_pairs):
    """Parity where the numbers are taken. (1) The step's own read pairs through the unmodified reference's hot path
    (`kma -ipe ... -apm p -s2 -t cores | alnFrags_threaded`): both ConClave arrays equal, frag_raw multiset equal.
    (2) The first `file_pairs` pairs through the whole program twice -- `kma` and the same host linked over libkmagpu.so
    (oracle/_ref/kma_gpu) -- and the output files compared."""
    from kma_b200 import records
    n = len(r1)
    out = {"config": "C2", "n": n}
    f1, f2 = pair_fastq(tmp, r1, r2, "step")
    dt, fr, sc = ref_hotpath(prefix, ["-ipe", f1, f2], cores, tmp, "c2", ["-apm", "p"], ["-apm-p"])
    db = api.TemplateDB(prefix, device=device)
    p = api.default_params()
    p.counters = 0
    db.seed_upload(records.stage1_pairs_fast(r1, r2))
    db.seed_run(p)
    db.align_from_seed()
    db.align_run(p)
    frag, a, u, _ = db.align_download()
    out.update(compare_hotpath(api, frag, a, u, fr, sc, db.info.DB_size))
    out["reference_s"] = round(dt, 1)
    db.close()
    if os.path.exists(os.path.join(REF, "kma_gpu")) and file_pairs:
        m = min(file_pairs, n)
        g1, g2 = pair_fastq(tmp, r1[:m], r2[:m], "files")
        env = dict(os.environ, KMAGPU_DEVICE=str(device))
        t_ref = ref_program("kma", prefix, ["-ipe", g1, g2], cores, tmp, "pf_ref", ["-apm", "p", "-matrix"])
        t_gpu = ref_program("kma_gpu", prefix, ["-ipe", g1, g2], 1, tmp, "pf_gpu", ["-apm", "p", "-matrix"], env=env)
        out["files"] = dict(compare_files(tmp, "pf_ref", "pf_gpu"), pairs=m, reference_s=round(t_ref, 1), gpu_host_s=round(t_gpu, 1),
                            what="kma -ipe -apm p -matrix vs the same host on libkmagpu.so (oracle/_ref/kma_gpu)")
    return out


def cpu_port(prefix, s1, nreads):
    import ctypes as C
    L = C.CDLL(os.path.join(ROOT, "oracle", "liborc.so"))
    L.orc_db_open.restype = C.c_void_p
    L.orc_db_open.argtypes = [C.c_char_p]
    L.orc_seed_stream.restype = C.c_int64
    L.orc_seed_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    db = L.orc_db_open(prefix.encode())
    p = (C.c_int32 * 40)()
    L.orc_default_params(p)
    out = np.zeros(3 * len(s1) + 4096, dtype=np.uint8)
    L.orc_align_stream.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_int,
                                   C.c_int, C.c_double, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]
    DB = int(np.fromfile(prefix + ".length.b", dtype=np.int32, count=1)[0])
    a, u = np.zeros(DB, np.uint64), np.zeros(DB, np.uint64)
    fo, fb = C.c_void_p(), C.c_size_t()
    t0 = time.perf_counter()
    n2 = L.orc_seed_stream(db, p, s1.ctypes.data, len(s1), out.ctypes.data, len(out), None)
    L.orc_align_stream(db, prefix.encode(), p, out.ctypes.data, n2, 0, 0.5, 0, 16, 0.0, C.byref(fo), C.byref(fb),
                       a.ctypes.data, u.ctypes.data, None, None, None)
    dt = time.perf_counter() - t0
    return nreads / dt, dt


def nw_gcups(db, seqs, peak_iops, n=24000, seed=3):
    """Banded-NW GCUPS on C3-shaped problems (template windows of 1-3 kb vs a 9 %-error copy, band = |dl| + 64 as
    KMA_score chooses it) through the NW queue kernels the alignment pass runs, timed alone with CUDA events inside the
    library (burst)."""
    rng = np.random.default_rng(seed)
    probs, qs, qoff = [], [], 0
    while len(probs) < n:
        t = int(rng.integers(0, len(seqs)))
        tl = len(seqs[t])
        if tl < 1000:
            continue
        t_len = int(rng.integers(1000, tl + 1))
        t_s = int(rng.integers(0, tl - t_len + 1))
        q = synth.mutate_indel(rng, seqs[t][t_s:t_s + t_len], 0.03, 0.03, 0.03)
        band = abs(t_len - len(q)) + 64
        if len(q) <= band or t_len <= band:
            continue
        probs.append([t + 1, t_s, t_s + t_len, qoff, 0, len(q), 0, band])
        qs.append(q)
        qoff += len(q)
    probs = np.array(probs, dtype=np.int32)
    qpool = np.concatenate(qs)
    best = None
    for _ in range(4):
        out, status, cells, steps, ms = db.nw_batch(probs, qpool)
        if best is None or ms < best:
            best = ms
    gc = cells / best / 1e6
    return {"kernel": "nw_warp_kernel / nw_thread_kernel: the NW queue kernels of the alignment pass (kmagpu_nw_batch feeds the same queue)", "problems": int(n), "cells": int(cells), "ms": best,
            "gcups": gc, "lane_utilisation": cells / (32.0 * steps), "int_ops_per_cell": 12,
            "int_roofline": {"achieved_tiops": gc * 12 / 1e3, "peak_tiops": peak_iops / 1e12, "frac": gc * 12e9 / peak_iops,
                             "peak_kind": "148 SMs x 128 int32 lanes x max SM clock"},
            "not_ok": int((status != 0).sum())}


COUNTER_FIELDS = ("mems", "nw_full_calls", "nw_band_calls", "nw_full_cells", "nw_band_cells", "nw_steps", "index_probes", "mem_bases",
                  "read_bytes")


def copy_counters(src, dst):
    """the in-kernel statistic counters of an alignment pass run with params.counters = 1 onto the stats of a timed run"""
    for f in COUNTER_FIELDS:
        setattr(dst, f, getattr(src, f))


def c4_flow(api, workdir, device, genome_bases=5_000_000, n=1_000_000):
    """BASELINE.json configs[3] (C4) beside the headline: one synthetic genome as the only template, 150 bp reads,
    -mem_mode -1t1, base counts and consensus, every stream resident in HBM between the stages: FASTQ text (pinned) ->
    record splitter + stage 1 -> stage 2 -> k-mer score collection -> ConClave -> traceback alignment + base counts ->
    consensus; wall clock per call, the text upload and the consensus download included. The stage outputs of this
    chain are compared with the oracle chain in tests/test_gpu_conclave.py and tools/c4_perf.py."""
    import torch
    from kma_b200 import dbbuild
    wd = os.path.join(workdir, f"c4_{genome_bases}")
    os.makedirs(wd, exist_ok=True)
    prefix = os.path.join(wd, "db")
    genome = np.random.default_rng(4).integers(0, 4, size=genome_bases).astype(np.uint8)
    if not os.path.exists(prefix + ".comp.b"):
        dbbuild.build_db(prefix, ["genome"], [genome])
    db = api.TemplateDB(prefix, device=device)
    p = api.default_params()
    p.one2one = 1
    p.matrix = 1
    a = synth.fastq_fixed(np.asarray(synth.short_reads(6, [genome], n, L=150, sub=0.01)))
    text = torch.empty(len(a), dtype=torch.uint8, pin_memory=True)
    text.numpy()[:] = a
    best = None
    for _ in range(4):
        t = {}
        t0 = time.perf_counter(); _, cnt, ms1, _, _ = db.run_input_text(text, download=False); t["split+stage1"] = time.perf_counter() - t0
        t0 = time.perf_counter(); st = db.seed_run(p); t["stage2"] = time.perf_counter() - t0
        t0 = time.perf_counter(); _, sa_, su_, _ = db.memscore_from_seed(download=False); t["score_collection"] = time.perf_counter() - t0
        t0 = time.perf_counter(); db.conclave_resident(sa_, su_, download=False); t["conclave"] = time.perf_counter() - t0
        db.matrix_reset()
        t0 = time.perf_counter(); _, nrec, sa = db.trace_from_conclave(p, download=False); t["traceback+counts"] = time.perf_counter() - t0
        t0 = time.perf_counter(); ct, cs, cq, cst, msc = db.consensus(1); t["consensus"] = time.perf_counter() - t0
        tot = sum(t.values())
        if best is None or tot < best[0]:
            best = (tot, t, cnt, nrec, sa, cst, st, ms1, msc)
    tot, t, cnt, nrec, sa, cst, st, ms1, msc = best
    db.close()
    return {"workload": f"C4: {n} synthetic 150 bp reads vs one {genome_bases / 1e6:.0f} Mb genome, -mem_mode -1t1, base counts + consensus, resident in HBM",
            "reads_per_s": n / tot, "ms": tot * 1e3, "stage_wall_ms": {k: round(v * 1e3, 2) for k, v in t.items()},
            "kernel_ms": {"stage1": ms1, "stage2": st.ms_total, "traceback": sa.ms_align, "consensus": msc},
            "reads_kept": int(cnt), "fragments": int(nrec), "h2d_bytes": int(len(a)), "d2h_bytes": 3 * genome_bases,
            "mean_depth": float(cst[0]["depth"]) / genome_bases, "consensus_matches_template": int(cst[0]["cover"])}


def c3_chain(db, api, seqs, prefix, workdir, cores, peak_gbs, n=20000, ref_n=2000, seed=22):
    """BASELINE.json configs[2] (C3) beside the headline: Nanopore-like reads (5-20 kb, 10 % errors) through stage 2 in
    chain mode (save_kmers_chain, the reference's default without -1t1: `chain_kernel`) and the alignment pass with
    the records' query bounds, chained in HBM; kernels timed by CUDA events inside the library, `e2e` from pinned-free
    host buffers (H2D of the stage-1 records, D2H of frag_raw + score arrays inside the timed region). The unmodified
    reference runs `kma -s2 -t cores | alnFrags_threaded` on the first `ref_n` reads."""
    reads = synth.long_reads(seed, seqs, n)
    s1 = records.stage1_records(reads)
    bases = int(sum(len(r) for r in reads))
    p = api.default_params()
    p.kmerscan = 1
    db.seed_upload(s1)
    best = None
    p.counters = 0   # timed without the alignment kernel's statistic counters (measurement instrumentation)
    for _ in range(4):
        st = db.seed_run(p)
        nrec = db.align_from_seed()
        sa = db.align_run(p)
        if best is None or st.ms_total + sa.ms_total < best[0]:
            best = (st.ms_total + sa.ms_total, st, sa, nrec)
    ms, st, sa, nrec = best
    p.counters = 1   # one more pass for the counts the figures below are made of
    db.seed_run(p); db.align_from_seed()
    copy_counters(db.align_run(p), sa)
    p.counters = 0
    # end to end through the chunked multi-stream pipeline (pinned host in / out, every copy inside the timed region)
    import torch
    from kma_b200 import pipeline
    s1p = torch.empty(len(s1), dtype=torch.uint8, pin_memory=True)
    s1p.numpy()[:] = s1
    pipe = pipeline.MapPipeline(prefix, device=db.device, workers=4, params=p)
    bounds = pipe.chunk_bounds(s1p.numpy(), 8)
    per_chunk = (db.align_out_bytes() // max(1, len(bounds))) * 3 // 2 + (1 << 20)
    outs = [torch.empty(per_chunk, dtype=torch.uint8, pin_memory=True) for _ in bounds]
    scores = (np.zeros(db.info.DB_size, np.uint64), np.zeros(db.info.DB_size, np.uint64))
    pipe.map(s1p, bounds, outs, scores)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r_e2e = pipe.map(s1p, bounds, outs, scores)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    d2h = sum(int(f.numel() if hasattr(f, "numel") else len(f)) for f, _ in r_e2e) + 16 * db.info.DB_size * len(bounds)
    pipe.close()
    cells = sa.nw_full_cells + sa.nw_band_cells
    out = {"workload": f"C3: {n} synthetic Nanopore-like reads (5-20 kb, 10 % errors, {bases / 1e6:.0f} Mb) vs the redundant gene DB, chain mode (no -1t1)",
           "reads_per_s": n / (ms * 1e-3), "bases_per_s": bases / (ms * 1e-3), "ms": ms,
           "e2e_reads_per_s": n / t_e2e, "h2d_bytes": int(len(s1)), "d2h_bytes": int(d2h), "e2e_pipeline": {"workers": 4, "chunks": len(bounds)},
           "stage2_records": int(nrec), "alignments": int(sa.tasks), "frag_records": int(sa.frags),
           "chain_kernel_ms": st.ms_seed, "lookups_per_read": st.lookups / n, "ankers_per_read": st.list_fetches / n,
           "aln_pair_kernel_ms": sa.ms_align, "nw_cells": int(cells), "nw_cells_banded_fraction": sa.nw_band_cells / max(1, cells),
           "align_gcups": cells / max(sa.ms_align, 1e-9) / 1e6}
    # chain_kernel, algorithmic bytes: the seeding figure of SURVEY 8d + 40 B per anker (written once,
    # read once) + 32 B per (anker, template) visit of the chaining DP (one 16-byte row read and written)
    vw = 2 if db.info.DB_size < 65535 else 4
    alg = algorithmic_bytes(st, vw) + 40 * st.list_fetches + 32 * st.list_ids
    ach = alg / (st.ms_seed * 1e-3) / 1e9
    out["chain_roofline"] = {"kernel": "chain_kernel", "bound": "hbm", "achieved": ach, "peak": peak_gbs, "unit": "GB/s",
                             "frac": ach / peak_gbs, "algorithmic_bytes_per_launch": int(alg), "bytes_per_read": alg / n,
                             "kernel_ms": st.ms_seed, "traffic": None,
                             "note": "table and per-warp rows are L2 resident; bound by dependent L2 round trips (ncu: long_scoreboard)"}
    # ncu DRAM traffic of chain_kernel per launch (profiles/traffic.json, scaled by reads)
    tr = ncu_traffic("chain_kernel", n)
    out["chain_roofline"]["traffic"] = tr
    if os.path.exists(os.path.join(REF, "kma")) and os.path.exists(os.path.join(REF, "ref_aln")):
        fq = os.path.join(workdir, f"c3_{ref_n}.fq")
        synth.write_fastq(fq, reads[:ref_n], qual="5")
        dt, fr, sc = ref_hotpath(prefix, ["-i", fq], cores, workdir, "c3")
        out["cpu_reference"] = {"reads_per_s": ref_n / dt, "cores": cores, "seconds": dt,
                                "sample": f"first {ref_n} reads; unmodified kma -s2 -t {cores} | alnFrags_threaded on {cores} pthreads"}
        # parity on the reference's sample: the same reads through the device path
        db.seed_upload(records.stage1_records(reads[:ref_n]))
        db.seed_run(p)
        db.align_from_seed()
        db.align_run(p)
        frag, a, u, _ = db.align_download()
        out["parity"] = dict(compare_hotpath(api, frag, a, u, fr, sc, db.info.DB_size), config="C3", n=ref_n)
    return out


def c4_files(api, workdir, device, cores, genome_bases, n):
    """C4 at file level: `kma -mem_mode -1t1 -matrix` and the same host on libkmagpu.so on a sample of the reads"""
    if not os.path.exists(os.path.join(REF, "kma_gpu")):
        return None
    wd = os.path.join(workdir, f"c4_{genome_bases}")
    genome = np.random.default_rng(4).integers(0, 4, size=genome_bases).astype(np.uint8)
    fq = os.path.join(wd, f"c4_{n}.fq")
    if not os.path.exists(fq):
        open(fq + ".tmp", "wb").write(synth.fastq_fixed(np.asarray(synth.short_reads(6, [genome], n, L=150, sub=0.01))).tobytes())
        os.replace(fq + ".tmp", fq)
    env = dict(os.environ, KMAGPU_DEVICE=str(device))
    flags = ["-mem_mode", "-1t1", "-matrix"]
    t_ref = ref_program("kma", "db", ["-i", fq], cores, wd, "pf_ref", flags)
    t_gpu = ref_program("kma_gpu", "db", ["-i", fq], 1, wd, "pf_gpu", flags, env=env)
    return dict(compare_files(wd, "pf_ref", "pf_gpu"), config="C4", n=n, reference_s=round(t_ref, 1), gpu_host_s=round(t_gpu, 1),
                what="kma -mem_mode -1t1 -matrix vs the same host on libkmagpu.so")


C5_FAMILIES, C5_TLEN = 5000, 10000


def c5_db(workdir):
    """BASELINE.json configs[4] at its full size (50k templates in 5000 families of 10 variants, ~500 Mb, ~113 M distinct
    16-mers): a k-mer table of 1.5 GB and a position index of ~13 GB, two orders of magnitude beyond the 126 MB L2 -- the
    HBM-resident gather. Built once per box with the reference's own indexer (about two minutes), cached in the work
    directory; --c5-families scales it down."""
    wd = os.path.join(workdir, f"c5_{C5_FAMILIES}_{C5_TLEN}")
    os.makedirs(wd, exist_ok=True)
    prefix = os.path.join(wd, "db")
    names, seqs = synth.gene_db(55, n_families=C5_FAMILIES, n_variants=10, len_lo=C5_TLEN * 3 // 4, len_hi=C5_TLEN * 5 // 4)
    how = "cached"
    if not os.path.exists(prefix + ".comp.b"):
        t0 = time.perf_counter()
        if os.path.exists(os.path.join(REF, "kma")):
            synth.write_fasta(os.path.join(wd, "db.fsa"), names, seqs)
            r = subprocess.run([os.path.join(REF, "kma"), "index", "-i", "db.fsa", "-o", "dbtmp"], cwd=wd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
            if r.returncode:
                raise RuntimeError("kma index failed: " + r.stderr.decode()[-300:])
            os.remove(os.path.join(wd, "db.fsa"))
            how = "kma index"
        else:
            dbbuild.build_db(os.path.join(wd, "dbtmp"), names, seqs)
            how = "kma_b200.dbbuild"
        for ext in (".length.b", ".seq.b", ".name", ".comp.b"):
            os.replace(os.path.join(wd, "dbtmp" + ext), prefix + ext)
        how += f" {time.perf_counter() - t0:.0f} s"
    return prefix, seqs, how


def c5_leg(api, workdir, rank, world, device, cores, pk, dist, n=4_000_000, ref_n=2000):
    """C5 beside the headline: short single-end reads against the large redundant database, -1t1, stage 2 + alignment pass
    resident in HBM, every rank its own reads against its own replica (the scaling-sweep config). Seeding here is the
    random-sector gather of SURVEY 8d: `roofline` is seed_se_kernel against the measured HBM peak, in algorithmic bytes
    and in 32-byte sectors."""
    import torch
    if rank == 0:
        prefix, seqs, how = c5_db(workdir)
    if world > 1:
        dist.barrier()
    if rank != 0:
        prefix, seqs, how = c5_db(workdir)
    t0 = time.perf_counter()
    db = api.TemplateDB(prefix, device=device)
    t_open = time.perf_counter() - t0
    reads = synth.short_reads(56 + 1000 * rank, seqs, n)
    s1 = records.stage1_records_fast(reads, first=rank * n)
    p = api.default_params()
    p.one2one = 1
    p.counters = 0
    db.seed_upload(s1)
    best = None
    for _ in range(4):
        st = db.seed_run(p)
        db.align_from_seed()
        sa = db.align_run(p)
        if best is None or st.ms_total + sa.ms_total < best[0]:
            best = (st.ms_total + sa.ms_total, st, sa)
    ms, st, sa = best
    tt = torch.tensor([ms, st.ms_seed, sa.ms_align], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_max, seed_max, aln_max = (float(x) for x in tt.cpu())
    info = db.info
    vw = 2 if info.DB_size < 65535 else 4
    alg = algorithmic_bytes(st, vw)
    # 32-byte sectors the gather needs at the least: one per bucket probe (the bucket entry carries its first key and value), the lists
    sectors = st.lookups + st.list_fetches + (vw * st.list_ids + 31) // 32 + (8 * st.read_words + 31) // 32
    ach = alg / (st.ms_seed * 1e-3) / 1e9
    out = {"workload": f"C5 ({'full' if C5_FAMILIES >= 5000 else 'reduced'} scale): {info.DB_size - 1} templates / {info.seq_bases / 1e6:.0f} Mb redundant DB ({C5_FAMILIES} families x 10), "
                       f"{n} synthetic 150 bp single-end reads per GPU, -1t1: stage 2 + alignment pass resident in HBM",
           "db": {"templates": info.DB_size - 1, "bases": int(info.seq_bases), "kmers": int(info.n), "hash_slots": int(info.size),
                  "device_bytes": int(info.device_bytes), "built": how, "open_s": round(t_open, 1)},
           "reads_per_s": world * n / (ms_max * 1e-3), "ms": ms_max, "seed_kernel_ms": seed_max, "align_ms": aln_max,
           "lookups_per_read": st.lookups / n, "lookups_per_s": st.lookups / (st.ms_seed * 1e-3), "alignments_per_read": sa.tasks / max(1, sa.reads),
           "overflow_reads": int(st.overflow_reads),
           "roofline": {"kernel": "seed_se_kernel", "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                        "algorithmic_bytes_per_launch": int(alg), "kernel_ms": st.ms_seed, "traffic": ncu_traffic("c5:seed_se_kernel", n),
                        "sector_bytes_per_launch": int(32 * sectors), "sector_GBs": 32 * sectors / (st.ms_seed * 1e-3) / 1e9,
                        "sector_frac": 32 * sectors / (st.ms_seed * 1e-3) / 1e9 / pk["hbm_gbs"],
                        "note": "k-mer table out of L2: every bucket probe is a 32-byte HBM sector (16-byte bucket entries {key0, value0, pos, cnt}: "
                                "one sector per lookup unless the first key of a multi-key bucket differs)"}}
    tr = out["roofline"]["traffic"]
    if tr:   # DRAM bytes of the ncu capture over this run's kernel time: what the random gather costs the memory system
        out["roofline"]["dram_GBs"] = tr / (st.ms_seed * 1e-3) / 1e9
        out["roofline"]["dram_frac"] = out["roofline"]["dram_GBs"] / pk["hbm_gbs"]
    if rank == 0 and os.path.exists(os.path.join(REF, "kma")) and os.path.exists(os.path.join(REF, "ref_aln")):
        fq = os.path.join(workdir, f"c5_{ref_n}.fq")
        open(fq, "wb").write(synth.fastq_fixed(np.asarray(reads[:ref_n])).tobytes())
        dt, fr, sc = ref_hotpath(prefix, ["-i", fq], cores, workdir, "c5", ["-1t1"], ["-1t1"])
        db.seed_upload(records.stage1_records_fast(reads[:ref_n]))
        db.seed_run(p)
        db.align_from_seed()
        db.align_run(p)
        frag, a, u, _ = db.align_download()
        out["parity"] = dict(compare_hotpath(api, frag, a, u, fr, sc, info.DB_size), config="C5", n=ref_n)
        out["cpu_reference"] = {"reads_per_s": ref_n / dt, "cores": cores, "seconds": round(dt, 2),
                                "sample": f"first {ref_n} reads; unmodified kma -1t1 -s2 -t {cores} | alnFrags_threaded (incl. loading the {info.device_bytes / 1e9:.1f} GB-class DB)"}
    db.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=2_000_000, help="read pairs per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=500_000, help="read pairs per step of the reference program (a step of the reference arm / the cpu_baseline leg)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity legs (reference hot path on the step's own reads, output files on a sample)")
    ap.add_argument("--file-pairs", type=int, default=100_000, help="read pairs of the file-level parity sample")
    ap.add_argument("--no-c3", action="store_true", help="skip the C3 (long reads, chain mode) side measurement")
    ap.add_argument("--no-c4", action="store_true", help="skip the C4 (one genome, -mem_mode, consensus; resident flow) side measurement")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5 (large redundant DB: k-mer table out of L2) side measurement")
    ap.add_argument("--c5-reads", type=int, default=4_000_000)
    ap.add_argument("--c5-families", type=int, default=C5_FAMILIES, help="families of 10 templates of ~10 kb in the C5 database (5000 = BASELINE.json configs[4])")
    ap.add_argument("--e2e-workers", type=int, default=4, help="host threads / library handles (clones of one database image) of the end-to-end pipelines")
    ap.add_argument("--e2e-chunks", type=int, default=16, help="chunks the batch is cut into for the hot-path-only end-to-end pipeline")
    args = ap.parse_args()
    globals()["C5_FAMILIES"] = args.c5_families

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    workdir = os.path.join(tempfile.gettempdir(), "kma_b200_bench")
    os.makedirs(workdir, exist_ok=True)
    have_ref = os.path.exists(os.path.join(REF, "kma"))

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        prefix, names, seqs = make_db(workdir)
        sample = min(args.cpu_sample, args.pairs)
        r1, r2 = synth.paired_reads(READ_SEED, seqs, sample)
        vals = []
        for i in range(args.warmup + args.steps):
            if have_ref:
                v, dt = cpu_reference(prefix, r1, r2, cores, workdir)
            else:
                v, dt = cpu_port(prefix, records.stage1_pairs_fast(r1, r2), 2 * sample)
            if i >= args.warmup:
                vals.append((v, dt))
        v = sum(2 * sample for _ in vals) / sum(dt for _, dt in vals)
        what = (f"{sample} read pairs of the same workload per step; the unmodified program, kma -ipe f1 f2 -apm p -t {cores} -o out: FASTQ parse, stage 2, "
                "alignment pass, ConClave, assembly + consensus, .res/.fsa/.aln/.frag.gz writers") if have_ref else \
               f"{sample} read pairs per step; oracle/liborc.so stage 2 + alignment pass (the reference is not built here)"
        line = {"impl": "reference", "metric": METRIC,
                "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * statistics.mean(dt for _, dt in vals), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                # the named workload, under the keys of the GPU arm's line; each step of this arm is a bounded sample of it
                "config": {"workload": WORKLOAD, "pairs_per_gpu_per_step": args.pairs, "reads_per_gpu_per_step": 2 * args.pairs, "read_len": 150},
                "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores if have_ref else 1, "kind": "reference" if have_ref else "port", "sample": what,
                                 "sample_pairs_per_step": sample},
                "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    import torch.distributed as dist
    from kma_b200 import api, pipeline

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own version / debug lines go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if rank == 0:
        prefix, names, seqs = make_db(workdir)
    if world > 1:
        dist.barrier()
    if rank != 0:
        prefix, names, seqs = make_db(workdir)

    r1, r2 = synth.paired_reads(READ_SEED + 1000 * rank, seqs, args.pairs)
    r1, r2 = np.asarray(r1), np.asarray(r2)
    s1_np = records.stage1_pairs_fast(r1, r2, first=rank * args.pairs)
    s1 = torch.empty(len(s1_np), dtype=torch.uint8, pin_memory=True)
    s1.numpy()[:] = s1_np

    db = api.TemplateDB(prefix, device=local_rank)
    params = api.default_params()
    params.counters = 0
    vw = 2 if db.info.DB_size < 65535 else 4
    DBn = db.info.DB_size
    total_bases = int(db.info.seq_bases)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_resident():
        st = db.seed_run(params)
        db.align_from_seed()
        sa = db.align_run(params)
        return st, sa

    # ---- resident-input timing (value): kernels only, device events inside the library
    db.seed_upload(s1)
    for _ in range(args.warmup):
        st, sa = step_resident()
    sampler = ClockSampler(local_rank)
    sync_all()
    sampler.start()
    t_dev, t_seed, t_pair, launches = 0.0, 0.0, 0.0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st, sa = step_resident()
        t_dev += st.ms_total + sa.ms_total
        t_seed += st.ms_seed
        t_pair += sa.ms_align
        launches += st.launches + sa.launches
    sync_all()
    t_wall = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    out_bytes = db.align_out_bytes()
    # the timed steps run without the pair kernel's statistic counters (measurement instrumentation); one untimed step with
    # them gives the algorithmic-byte inputs
    params.counters = 1
    _, sa_counted = step_resident()
    params.counters = 0
    copy_counters(sa_counted, sa)

    # ---- hot path only, end to end (extra): stage-1 records in pinned host memory -> frag_raw + score arrays on the host
    pipe = pipeline.MapPipeline(prefix, device=local_rank, workers=args.e2e_workers, params=params)
    scores = (np.zeros(DBn, np.uint64), np.zeros(DBn, np.uint64))
    bounds = pipe.chunk_bounds(s1.numpy(), args.e2e_chunks)
    per_chunk = (out_bytes // max(1, len(bounds))) * 5 // 4 + (1 << 20)
    outs = [torch.empty(per_chunk, dtype=torch.uint8, pin_memory=True) for _ in bounds]
    for _ in range(max(1, args.warmup // 2)):
        pipe.map(s1, bounds, outs, scores)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r_hot = pipe.map(s1, bounds, outs, scores)
    sync_all()
    t_hot = (time.perf_counter() - t0) * 1e3
    hot_bytes = sum(int(f.numel() if hasattr(f, "numel") else len(f)) for f, _ in r_hot)
    assert hot_bytes == out_bytes, "chunked end-to-end run produced a different frag_raw size"
    del outs

    # ---- e2e: the whole program span. FASTQ text of both files in pinned host memory -> ... -> consensus rows + the
    # per-template fragment stream on the host; both exchanges (ConClave sums, base-count matrix) through NCCL inside the
    # library. Every byte crosses PCIe inside the timed region.
    if world > 1:
        pipe.dbs[0].comm_init_torch()
    txt, text_bytes = [], 0
    for r in (r1, r2):
        a = synth.fastq_fixed(r, first=rank * args.pairs)
        t = torch.empty(len(a), dtype=torch.uint8, pin_memory=True)
        t.numpy()[:] = a
        txt.append(t)
        text_bytes += len(a)
    del a
    W = args.e2e_workers
    frag_cap = (2 * args.pairs * 260) // W * 5 // 4 + (1 << 20)
    frag_outs = [torch.empty(frag_cap, dtype=torch.uint8, pin_memory=True) for _ in range(W)]
    cons_out = [torch.empty(total_bases, dtype=torch.uint8, pin_memory=True) for _ in range(3)] + \
               [torch.empty(DBn * api.CONSENSUS_STATS.itemsize, dtype=torch.uint8, pin_memory=True)]
    for _ in range(max(2, args.warmup // 2)):
        full = pipe.map_to_consensus(txt[0], txt[1], frag_outs, params, cons_out=cons_out)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        full = pipe.map_to_consensus(txt[0], txt[1], frag_outs, params, cons_out=cons_out)
    sync_all()
    t_e2e = (time.perf_counter() - t0) * 1e3
    assert full["reads"] == args.pairs, "the text path kept a different number of pairs"
    cons_stats = full["consensus"][3]
    e2e_d2h = int(full["frag_bytes"]) + 3 * total_bases + int(cons_stats.nbytes) + 16 * DBn
    # warm, repeated timing of the two exchanges alone (device events inside the library)
    ar_scores = min(pipe.dbs[0].allreduce_scores(download=False)[2] for _ in range(20))
    ar_matrix = min(pipe.dbs[0].allreduce_matrix() for _ in range(5))
    pipe.close()
    del frag_outs, txt, cons_out

    tt = torch.tensor([t_dev, t_e2e, t_wall, t_hot, ar_scores, ar_matrix], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(st.reads), float(st.mapped), float(sa.frags), float(args.pairs), float(full["fragments"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    t_dev_max, t_e2e_max, t_wall_max, t_hot_max, ar_scores_max, ar_matrix_max = (float(x) for x in tt.cpu())
    total_reads = float(cnt[0]) * args.steps

    pk, pk_kind = peaks()
    peak_iops = 148 * 128 * pk.get("sm_max_mhz", 1965.0) * 1e6   # int32 lanes x clock
    alg_seed = algorithmic_bytes(st, vw)
    cells = sa.nw_full_cells + sa.nw_band_cells
    ms_seed = t_seed / args.steps
    ms_pair = t_pair / args.steps
    # alignment pass (pair kernel + NW queue kernels), algorithmic bytes per launch: per pair its read (packed
    # words + 0-4 bytes + N list) and the 32-byte result row; 8 B per position-index probe the reference's seed scan makes;
    # 2 x 2 bit per base compared by MEM extension; per NW cell 1 B traceback written + 1 B query base + 2 bit template base
    alg_pair = (sa.read_bytes + 32 * sa.tasks + 8 * sa.index_probes + sa.mem_bases // 2 + (9 * cells) // 4)
    ach_pair = alg_pair / (ms_pair * 1e-3) / 1e9
    ach_seed = alg_seed / (ms_seed * 1e-3) / 1e9
    rf_pair = {"kernel": "aln_pair_kernel + nw_thread_kernel / nw_warp_kernel", "bound": "hbm", "achieved": ach_pair, "peak": pk["hbm_gbs"], "unit": "GB/s",
               "frac": ach_pair / pk["hbm_gbs"], "peak_kind": pk_kind, "traffic": ncu_traffic("aln_pair_kernel", args.pairs),
               "algorithmic_bytes_per_launch": alg_pair, "kernel_ms": ms_pair,
               "note": "latency/issue bound (dependent index probes, short DP); see nw for the integer roofline",
               "per_read": {"alignments": sa.tasks / sa.reads, "index_probes": sa.index_probes / sa.reads, "mems": sa.mems / sa.reads, "nw_cells": cells / sa.reads, "bytes": alg_pair / sa.reads}}
    rf_seed = {"kernel": "seed_se_kernel<hash, common shape>", "bound": "hbm", "achieved": ach_seed, "peak": pk["hbm_gbs"], "unit": "GB/s",
               "frac": ach_seed / pk["hbm_gbs"], "peak_kind": pk_kind, "traffic": ncu_traffic("seed_se_kernel", args.pairs), "algorithmic_bytes_per_launch": alg_seed, "kernel_ms": ms_seed,
               "note": "the C2 k-mer table (32 MB of bucket entries + lists) is L2 resident: the gather never reaches HBM here; c5.roofline is the same kernel with the table out of L2",
               "per_read": {"lookups": st.lookups / st.reads, "hits": st.hits / st.reads,
                            "list_fetches": st.list_fetches / st.reads, "bytes": alg_seed / st.reads}}
    line = {
        "metric": METRIC,
        "value": total_reads / (t_dev_max * 1e-3), "unit": "reads/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_dev_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_gpu_per_step": args.pairs, "reads_per_gpu_per_step": 2 * args.pairs, "read_len": 150,
                   "db_templates": db.info.DB_size - 1, "db_kmers": int(db.info.n),
                   "db_device_bytes": int(db.info.device_bytes), "mapped_pair_fraction": float(cnt[1]) / float(cnt[3]),
                   "frag_records_per_pair": float(cnt[2]) / float(cnt[3]),
                   "alignments_per_read": sa.tasks / max(1, sa.reads),
                   "cache": f"stage-1 batch {len(s1_np) / 1e6:.0f} MB + stage-2 stream + read slab + frag_raw {out_bytes / 1e6:.0f} MB per step exceed the 126 MB L2; "
                            "the 150 MB database image (hash table + per-template position index) is mostly L2-resident by nature of this config",
                   "sharding": "reads sharded by rank, database replicated per GPU; per step one NCCL all-reduce of the ConClave score arrays and one of the base-count matrix, inside libkmagpu",
                   "allreduce_scores_ms": ar_scores_max, "allreduce_matrix_ms": ar_matrix_max},
        "e2e": {"value": total_reads / (t_e2e_max * 1e-3), "unit": "reads/s", "h2d_bytes_per_step": int(text_bytes),
                "d2h_bytes_per_step": e2e_d2h, "ms_per_step": t_e2e_max / args.steps,
                "span": "FASTQ text of both files (pinned host) -> record splitter + stage 1 -> stage 2 -> alignment pass -> [all-reduce of the ConClave sums] -> "
                        "ConClave -> traceback alignment + base counts -> [all-reduce of the matrix] -> consensus; down: per-template fragment stream + consensus "
                        "rows + per-template sums (what .frag.gz / .res / .fsa / .aln are written from). The reference arm runs the same span: kma -ipe ... -o out",
                "pipeline": {"workers": W, "slices": W, "handles": "clones of one database image per GPU"},
                "fragments_per_step": float(cnt[4]), "templates_with_consensus": int((cons_stats["cover"] > 0).sum()),
                "host_link_GBs": (text_bytes + e2e_d2h) / (t_e2e_max / args.steps * 1e-3) / 1e9},
        "e2e_hotpath": {"what": "stage 2 + alignment pass only, from stage-1 records in pinned host memory to frag_raw + score arrays on the host (the boundary of "
                                "kmagpu_seed_batch / kmagpu_align_batch), chunked over the worker handles",
                        "value": total_reads / (t_hot_max * 1e-3), "unit": "reads/s", "ms_per_step": t_hot_max / args.steps,
                        "h2d_bytes_per_step": int(len(s1_np)), "d2h_bytes_per_step": out_bytes + 16 * DBn * len(bounds), "chunks": len(bounds)},
        "gpu_launches": launches,
        "wall_ms_per_step_resident": t_wall_max / args.steps,
        "stage_ms": {"seed_total": st.ms_total, "seed_kernel": st.ms_seed, "align_pairs_and_nw": sa.ms_align,
                     "align_sizes_prep_select_emit": sa.ms_reduce},
        "clocks": clocks,
    }
    if ms_seed > ms_pair:
        line["roofline"], line["roofline_pair"] = rf_seed, rf_pair
    else:
        line["roofline"], line["roofline_seed"] = rf_pair, rf_seed
    if rank == 0:
        line["nw"] = nw_gcups(db, seqs, peak_iops)
    side_err = {}
    if not args.no_c3 and rank == 0:
        try:
            line["c3"] = c3_chain(db, api, seqs, prefix, workdir, cores, pk["hbm_gbs"])
        except Exception as e:   # a side measurement must not cost the headline line
            side_err["c3"] = repr(e)[:300]
    if not args.no_c4 and rank == 0:
        try:
            line["c4"] = c4_flow(api, workdir, local_rank)
            if have_ref and not args.no_parity:
                line["c4"]["parity"] = c4_files(api, workdir, local_rank, cores, 5_000_000, 100_000)
        except Exception as e:
            side_err["c4"] = repr(e)[:300]
    if not args.no_c5:   # every rank: the scaling-sweep config
        try:
            c5 = c5_leg(api, workdir, rank, world, local_rank, cores, pk, dist, n=args.c5_reads)
            if rank == 0:
                line["c5"] = c5
        except Exception as e:
            side_err["c5"] = repr(e)[:300]
            if world > 1:
                raise
    if side_err:
        line["side_errors"] = side_err

    if rank == 0 and world == 1 and have_ref and not args.no_parity:
        try:
            line["parity"] = parity_c2(api, prefix, r1, r2, cores, workdir, local_rank, args.file_pairs)
        except Exception as e:
            line["parity"] = {"config": "C2", "error": repr(e)[:300]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = min(args.cpu_sample, args.pairs)
        if have_ref:
            v, dt = cpu_reference(prefix, r1[:sample], r2[:sample], cores, workdir)
            line["cpu_baseline"] = {"value": v, "unit": "reads/s", "cores": cores, "kind": "reference",
                                    "sample": f"first {sample} read pairs of the step; the unmodified program, kma -ipe f1 f2 -apm p -t {cores} -o out (FASTQ parse ... "
                                              f"consensus + writers: the span of e2e), {dt:.1f} s"}
        else:
            v, dt = cpu_port(prefix, records.stage1_pairs_fast(r1[:sample], r2[:sample]), 2 * sample)
            line["cpu_baseline"] = {"value": v, "unit": "reads/s", "cores": 1, "kind": "port",
                                    "sample": f"first {sample} read pairs of the step; oracle/liborc.so stage 2 + alignment pass, {dt:.1f} s"}
    db.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
