"""Exploratory: resident PE pipeline timing per step at several batch sizes (chunking / working-set effects)."""
import os, sys, tempfile, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kma_b200 import api, synth, records
import bench

def main():
    sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [250000, 1000000, 2000000]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    wd = os.path.join(tempfile.gettempdir(), "kma_b200_bench"); os.makedirs(wd, exist_ok=True)
    prefix, names, seqs = bench.make_db(wd)
    db = api.TemplateDB(prefix)
    p = api.default_params()
    p.counters = int(os.environ.get("KG_COUNTERS", "0"))   # 0: the production pair kernel (no statistic counters)
    for n in sizes:
        r1, r2 = synth.paired_reads(7, seqs, n)
        s1 = records.stage1_pairs_fast(r1, r2)
        db.seed_upload(s1)
        rows = []
        for it in range(steps):
            st = db.seed_run(p); db.align_from_seed(); sa = db.align_run(p)
            rows.append((round(st.ms_seed, 2), round(sa.ms_prep, 2), round(sa.ms_align, 2), round(sa.ms_reduce, 2)))
        per = [round(1e3 * (a + b + c + d) / (2 * n), 4) for a, b, c, d in rows]
        print(json.dumps({"pairs": n, "seed/prep/pairs/reduce ms": rows, "us_per_read": per, "tasks": sa.tasks}))
    db.close()
main()
