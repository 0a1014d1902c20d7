"""Exploratory timing of the resident pipeline (not the bench contract): stage 2 + alignment pass on C1-shaped data,
and the NW batch kernel on long banded problems."""
import os, sys, time, tempfile, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kma_b200 import api, synth, records, dbbuild

def main():
    nreads = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    wd = os.path.join(tempfile.gettempdir(), "kma_b200_bench"); os.makedirs(wd, exist_ok=True)
    prefix = os.path.join(wd, "db")
    names, seqs = synth.gene_db(42)
    if not os.path.exists(prefix + ".comp.b"):
        dbbuild.build_db(prefix, names, seqs)
    t0 = time.time(); db = api.TemplateDB(prefix); print("db_open s", time.time() - t0, "device MB", db.info.device_bytes / 1e6)
    reads = synth.short_reads(7, seqs, nreads)
    s1 = records.stage1_records_fast(reads)
    db.seed_upload(s1)
    for it in range(3):
        st = db.seed_run()
        n = db.align_from_seed()
        t0 = time.time(); sa = db.align_run(); t1 = time.time()
        print(json.dumps({"seed_ms": st.ms_total, "align": sa.as_dict(), "wall_align_ms": (t1 - t0) * 1e3}))
    # NW batch: banded problems of C3 shape (t ~ 2000, band 64..100) and small full ones
    rng = np.random.default_rng(1)
    lens = np.array([len(s) for s in seqs])
    for shape in ("band", "full_small", "full_mid"):
        probs, qs, qoff = [], [], 0
        for i in range(20000 if shape != "band" else 4000):
            t = int(rng.integers(0, len(seqs)))
            tl = len(seqs[t])
            if shape == "band":
                t_len = min(tl, int(rng.integers(1000, 3000)))
                t_s = int(rng.integers(0, tl - t_len + 1))
                q = synth.mutate_indel(rng, seqs[t][t_s:t_s + t_len], 0.03, 0.03, 0.03)
                band = abs(t_len - len(q)) + 64
                if len(q) <= band or t_len <= band: continue
            else:
                t_len = int(rng.integers(1, 40)) if shape == "full_small" else int(rng.integers(60, 200))
                t_len = min(t_len, tl)
                t_s = int(rng.integers(0, tl - t_len + 1))
                q = synth.mutate_indel(rng, seqs[t][t_s:t_s + t_len], 0.05, 0.03, 0.03)
                if len(q) == 0: continue
                band = 0
                if not (len(q) <= abs(t_len - len(q)) + 64 or t_len <= abs(t_len - len(q)) + 64): continue
            probs.append([t + 1, t_s, t_s + t_len, qoff, 0, len(q), 0, band]); qs.append(q); qoff += len(q)
        probs = np.array(probs, dtype=np.int32); qpool = np.concatenate(qs)
        for it in range(3):
            out, status, cells, steps, ms = db.nw_batch(probs, qpool)
        print(shape, "n", len(probs), "cells", cells, "steps", steps, "ms", ms, "GCUPS", cells / ms / 1e6, "lane util", cells / (32.0 * steps), "bad", int((status != 0).sum()))
    db.close()
main()
