"""NW batch kernel alone on C3-shaped banded problems (profiling target for ncu)."""
import os, sys, tempfile, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kma_b200 import api, synth, dbbuild
import bench

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
    wd = os.path.join(tempfile.gettempdir(), "kma_b200_bench"); os.makedirs(wd, exist_ok=True)
    prefix, names, seqs = bench.make_db(wd)
    db = api.TemplateDB(prefix)
    pk, _ = bench.peaks()
    print(json.dumps(bench.nw_gcups(db, seqs, 148 * 128 * pk.get("sm_max_mhz", 1965.0) * 1e6, n=n)))
    db.close()
main()
