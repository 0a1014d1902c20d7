"""consensus_kernel alone on a database whose base-count matrix is far larger than the L2 (scaled C5: 12.5 k templates,
125 Mb -> 3 GB of counts): kernel ms and algorithmic GB/s (24 B counts in + 3 B rows out + 0.25 B template per position).
usage: consensus_perf.py [families=1250] [template_len=10000] [reps=5]"""
import os, sys, json, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from kma_b200 import api, synth, dbbuild
from tests import util


def main():
    fam = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
    tl = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    wd = os.path.join(tempfile.gettempdir(), f"kma_b200_c5_{fam}_{tl}"); os.makedirs(wd, exist_ok=True)
    prefix = os.path.join(wd, "db")
    names, seqs = synth.gene_db(55, n_families=fam, n_variants=10, len_lo=tl * 3 // 4, len_hi=tl * 5 // 4)
    if not os.path.exists(prefix + ".comp.b"):
        dbbuild.build_db(prefix, names, seqs)
    db = api.TemplateDB(prefix)
    db.matrix_reset()
    dev = db.matrix_tensor().view(-1, 6)
    n = dev.shape[0]
    # depth ~ 40 with the template base in the majority, some noise, some uncovered positions
    g = torch.Generator(device="cuda").manual_seed(1)
    dev.copy_(torch.randint(0, 3, dev.shape, generator=g, device="cuda", dtype=torch.int32))
    tb = torch.randint(0, 6, (n,), generator=g, device="cuda")
    dev[torch.arange(n, device="cuda"), tb] += torch.randint(0, 60, (n,), generator=g, device="cuda", dtype=torch.int32)
    dev[::17] = 0
    torch.cuda.synchronize()
    # parity of one template against the oracle on this very matrix
    t = 7
    off = util.matrix_offsets(prefix)
    m = np.minimum(dev[off[t]:off[t + 1]].cpu().numpy(), 65535).astype(np.uint16)
    wt, ws, wq, wst = util.oracle_consensus(prefix, t, m)
    got = db.consensus(t)
    assert got[0].tobytes() == wt and got[1].tobytes() == ws and got[2].tobytes() == wq and int(got[3][0]["depth"]) == int(wst[0])
    best = None
    for _ in range(reps):
        ms = db.consensus(0)[4]
        best = ms if best is None or ms < best else best
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = peaks.get("hbm_gbs", 6650.0)
    alg = 27.25 * n
    print(json.dumps({"kernel": "consensus_kernel", "positions": int(n), "templates": len(seqs), "ms": best, "alg_bytes": alg,
                      "alg_GBs": alg / best / 1e6, "peak_GBs": peak, "frac_of_hbm": alg / best / 1e6 / peak,
                      "Gpositions_per_s": n / best / 1e6}))
    db.close()


main()
