"""Merge the per-capture DRAM traffic files of tools/r02_final.sh (gpurun_out/traffic_<tag>_{c2,c3,c5,nw}.json, written by
tools/ncu_summary.py) into gpurun_out/traffic_<tag>.json with the keys bench.py's ncu_traffic() reads: the C2 kernels as they
are (the pair entry = aln_pair_kernel + the NW queue kernels of the same step, the label of `roofline_pair`), `chain_kernel`
from the C3 capture, `c5:seed_se_kernel` from the C5 capture, the stand-alone NW kernel from the NW capture.
Copy the result to profiles/traffic.json to publish it. usage: merge_traffic.py <tag>"""
import json, os, sys

tag = sys.argv[1]
def load(x):
    p = f"gpurun_out/traffic_{tag}_{x}.json"
    return json.load(open(p)) if os.path.exists(p) else {}

out = {}
c2 = load("c2")
for k, v in c2.items():
    out[k] = v
if "aln_pair_kernel" in c2:
    extra = sum(c2[k]["dram_bytes"] for k in ("nw_thread_kernel", "nw_warp_kernel") if k in c2)
    out["aln_pair_kernel"] = dict(c2["aln_pair_kernel"], dram_bytes=c2["aln_pair_kernel"]["dram_bytes"] + extra,
                                  note="aln_pair_kernel + the longest nw_thread_kernel / nw_warp_kernel launch of the step")
c3 = load("c3")
if "chain_kernel" in c3:
    out["chain_kernel"] = c3["chain_kernel"]
for k, v in c3.items():
    out["c3:" + k] = v
for k, v in load("c5").items():
    out["c5:" + k] = v
for k, v in load("nw").items():
    out["nw:" + k] = v
json.dump(out, open(f"gpurun_out/traffic_{tag}.json", "w"), indent=1)
print("traffic keys:", sorted(out))
