"""Exploratory run of a scaled C5 (BASELINE.json configs[4]: large highly redundant DB, short reads, -1t1): a database
whose k-mer table no longer fits the 126 MB L2, so that seeding is the HBM random-sector gather SURVEY 8d describes.
usage: c5_perf.py [families=1250] [template_len=10000] [reads=4000000] [check=2000]
Uses the database of bench.py's C5 leg (built with the reference's `kma index`, cached per box), maps `reads` 150 bp single-end reads (stage 2 + alignment
pass, resident), checks the first `check` reads against the oracle and prints the stage timings and the seeding
roofline fraction."""
import os, sys, time, tempfile, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kma_b200 import api, synth, records, dbbuild
from tests import util


def main():
    fam = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
    tl = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    nreads = int(sys.argv[3]) if len(sys.argv) > 3 else 4_000_000
    check = int(sys.argv[4]) if len(sys.argv) > 4 else 2000
    # the database of bench.py's C5 leg (same cache directory: built once per box, by the reference's indexer when it is there)
    import bench
    bench.C5_FAMILIES, bench.C5_TLEN = fam, tl
    workdir = os.path.join(tempfile.gettempdir(), "kma_b200_bench"); os.makedirs(workdir, exist_ok=True)
    t0 = time.time()
    prefix, seqs, how = bench.c5_db(workdir)
    t1 = time.time()
    db = api.TemplateDB(prefix)
    t2 = time.time()
    info = db.info
    print(json.dumps({"templates": info.DB_size - 1, "bases": int(info.seq_bases), "kmers": int(info.n), "hash_slots": int(info.size),
                      "device_MB": info.device_bytes / 1e6, "build_s": round(t1 - t0, 1), "open_s": round(t2 - t1, 1)}), flush=True)
    reads = synth.short_reads(56, seqs, nreads)
    s1 = records.stage1_records_fast(reads)
    p = api.default_params(); p.one2one = 1
    if check:
        c1 = records.stage1_records_fast(reads[:check])
        want2 = util.oracle_seed_stream(prefix, c1)
        got2, n, _ = db.save_kmers_batch(c1, p)
        assert got2.tobytes() + api.stream_terminator(n) == want2.tobytes(), "stage 2 differs from the oracle"
        ofrag, oa, ou, _, _ = util.oracle_align_stream(prefix, want2, want_cand=False)
        frag, a, u, _, _ = db.alnFrags_batch(want2, p)
        assert frag.tobytes() == ofrag and np.array_equal(a, oa) and np.array_equal(u, ou), "alignment pass differs from the oracle"
        print("parity ok on", check, "reads", flush=True)
    db.seed_upload(s1)
    for it in range(3):
        st = db.seed_run(p)
        db.align_from_seed()
        sa = db.align_run(p)
    vw = 2 if info.DB_size < 65535 else 4
    alg = 4 * st.lookups + 8 * st.hits + vw * (st.list_fetches + st.list_ids) + 8 * st.read_words
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    print(json.dumps({"reads": nreads, "seed_ms": st.ms_seed, "seed_total_ms": st.ms_total, "align_ms": sa.ms_align, "align_total_ms": sa.ms_total,
                      "reads_per_s": nreads / ((st.ms_total + sa.ms_total) * 1e-3), "lookups_per_read": st.lookups / nreads,
                      "seed_alg_GBs": alg / (st.ms_seed * 1e-3) / 1e9, "seed_frac_of_hbm": alg / (st.ms_seed * 1e-3) / 1e9 / peaks["hbm_gbs"],
                      "lookups_per_s": st.lookups / (st.ms_seed * 1e-3), "alignments_per_read": sa.tasks / max(1, sa.reads),
                      "overflow_reads": st.overflow_reads}), flush=True)
    db.close()


main()
