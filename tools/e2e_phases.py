"""Where the whole-span end-to-end step (bench.py's `e2e`: FASTQ text -> ... -> consensus, pipeline.map_to_consensus) spends
its wall clock, per worker and phase, at 1..N ranks:
  python tools/e2e_phases.py [pairs] [workers]                                  (one GPU)
  python -m torch.distributed.run --nproc-per-node N ... tools/e2e_phases.py    (N GPUs; rank 0 prints)
Prints one JSON line per timed step: for every worker the seconds since the step began at which it left each phase."""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from kma_b200 import api, pipeline, synth  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    wts = [float(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else None
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    workdir = os.path.join(tempfile.gettempdir(), "kma_b200_bench")
    os.makedirs(workdir, exist_ok=True)
    if rank == 0:
        prefix, names, seqs = bench.make_db(workdir)
    if world > 1:
        dist.barrier()
    if rank != 0:
        prefix, names, seqs = bench.make_db(workdir)
    r1, r2 = synth.paired_reads(bench.READ_SEED + 1000 * rank, seqs, pairs)
    params = api.default_params()
    params.counters = 0
    pipe = pipeline.MapPipeline(prefix, device=local_rank, workers=W, params=params)
    if world > 1:
        pipe.dbs[0].comm_init_torch()
    txt = []
    for r in (np.asarray(r1), np.asarray(r2)):
        a = synth.fastq_fixed(r, first=rank * pairs)
        t = torch.empty(len(a), dtype=torch.uint8, pin_memory=True)
        t.numpy()[:] = a
        txt.append(t)
    fr = [w / sum(wts) for w in wts] if wts else [1.0 / W] * W
    frag_outs = [torch.empty(int(2 * pairs * 260 * f * 5 / 4) + (1 << 20), dtype=torch.uint8, pin_memory=True) for f in fr]
    info = pipe.dbs[0].info
    cons_out = [torch.empty(int(info.seq_bases), dtype=torch.uint8, pin_memory=True) for _ in range(3)] + \
               [torch.empty(info.DB_size * api.CONSENSUS_STATS.itemsize, dtype=torch.uint8, pin_memory=True)]

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    for step in range(6):
        sync()
        tr = []
        t0 = time.perf_counter()
        pipe.map_to_consensus(txt[0], txt[1], frag_outs, params, trace=tr, cons_out=cons_out, slice_weights=wts)
        t1 = time.perf_counter()
        if step >= 3 and rank == 0:
            rows = {}
            for w, what, t in tr:
                rows.setdefault("all" if w < 0 else f"w{w}", []).append((what, round((t - t0) * 1e3, 1)))
            print(json.dumps({"world": world, "pairs": pairs, "workers": W, "step_ms": round((t1 - t0) * 1e3, 1), "phases_ms": rows}))
    pipe.close()
    if world > 1:
        dist.destroy_process_group()


main()
