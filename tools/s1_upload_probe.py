"""How long the text upload + stage 1 of one slice takes on an idle GPU (wall clock around run_input_text), against the raw
pinned copy of the same bytes: is the 30 GB/s of the end-to-end step's first phase the link or the path?"""
import json, os, sys, time, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from kma_b200 import api, synth
wd = os.path.join(tempfile.gettempdir(), "kma_b200_bench"); os.makedirs(wd, exist_ok=True)
prefix, names, seqs = bench.make_db(wd)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
r1, r2 = synth.paired_reads(7, seqs, n)
t1, t2 = synth.fastq_fixed(np.asarray(r1)), synth.fastq_fixed(np.asarray(r2))
p1 = torch.empty(len(t1), dtype=torch.uint8, pin_memory=True); p1.numpy()[:] = t1
p2 = torch.empty(len(t2), dtype=torch.uint8, pin_memory=True); p2.numpy()[:] = t2
d = torch.empty(len(t1) + len(t2), dtype=torch.uint8, device="cuda")
db = api.TemplateDB(prefix)
out = {"bytes": len(t1) + len(t2)}
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _, cnt, ms, u1, u2 = db.run_input_text(p1.numpy(), text2=p2.numpy(), download=False)
    out["run_input_text_ms"] = round((time.perf_counter() - t0) * 1e3, 2); out["stage1_kernels_ms"] = round(ms, 2)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d[:len(t1)].copy_(p1, non_blocking=True); d[len(t1):].copy_(p2, non_blocking=True); torch.cuda.synchronize()
    out["raw_copy_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
out["raw_GBs"] = round(out["bytes"] / out["raw_copy_ms"] / 1e6, 1); out["path_GBs"] = round(out["bytes"] / out["run_input_text_ms"] / 1e6, 1)
print(json.dumps(out))
