"""Exploratory timing of the C3 pipeline (not the bench contract): stage 2 in chain mode (save_kmers_chain) + the
alignment pass on Nanopore-like reads (5-20 kb, 10 % errors) against the redundant gene DB, resident in HBM; the
unmodified reference (`kma -s2` without -1t1 piped into alnFrags_threaded) on a sample of the same reads beside it."""
import os, sys, time, tempfile, json, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kma_b200 import api, synth, records, dbbuild


def main():
    nreads = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    ref_n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    wd = os.path.join(tempfile.gettempdir(), "kma_b200_bench"); os.makedirs(wd, exist_ok=True)
    prefix = os.path.join(wd, "db")
    names, seqs = synth.gene_db(42)
    if not os.path.exists(prefix + ".comp.b"):
        dbbuild.build_db(prefix, names, seqs)
    db = api.TemplateDB(prefix)
    t0 = time.time()
    reads = synth.long_reads(22, seqs, nreads)
    s1 = records.stage1_records(reads)
    bases = sum(len(r) for r in reads)
    print("reads", nreads, "bases", bases, "gen s", round(time.time() - t0, 1), flush=True)
    p = api.default_params(); p.kmerscan = 1; p.counters = int(os.environ.get("KG_COUNTERS", "1"))
    for mode in ("chain", "1t1"):
        p.kmerscan = 1 if mode == "chain" else 0
        p.one2one = 0 if mode == "chain" else 1
        db.seed_upload(s1)
        for it in range(3):
            t0 = time.time(); st = db.seed_run(p); t1 = time.time()
            n = db.align_from_seed()
            sa = db.align_run(p); t2 = time.time()
        d = sa.as_dict()
        cells = d["nw_full_cells"] + d["nw_band_cells"]
        print(json.dumps({"mode": mode, "seed": st.as_dict(), "seed_wall_ms": (t1 - t0) * 1e3, "records": n, "align": d,
                          "align_wall_ms": (t2 - t1) * 1e3, "reads_per_s": nreads / (t2 - t0),
                          "align_gcups": cells / max(d["ms_align"], 1e-9) / 1e6}), flush=True)
    db.close()
    if ref_n:
        ref = os.path.join(ROOT, "oracle", "_ref")
        fq = os.path.join(wd, "c3.fq")
        synth.write_fastq(fq, reads[:ref_n], qual="5")
        nthr = os.cpu_count()
        for flags, tag in (([], "chain"), (["-1t1"], "1t1")):
            t0 = time.time()
            s2 = subprocess.run([os.path.join(ref, "kma"), "-i", fq, "-o", os.path.join(wd, "o"), "-t_db", prefix, "-s2", "-t", str(nthr)] + flags,
                                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
            t1 = time.time()
            open(os.path.join(wd, "c3.s2"), "wb").write(s2)
            subprocess.run([os.path.join(ref, "ref_aln"), prefix, os.path.join(wd, "c3.s2"), os.path.join(wd, "fr.out"), os.path.join(wd, "sc.out"),
                            "-t", str(nthr)] + flags, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            t2 = time.time()
            print(json.dumps({"reference": tag, "reads": ref_n, "threads": nthr, "stage12_s": t1 - t0, "align_s": t2 - t1,
                              "reads_per_s": ref_n / (t2 - t0)}), flush=True)


main()
