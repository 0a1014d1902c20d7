"""Multi-GPU check under torchrun (one rank per GPU): the two exchanges of the path on real devices, THROUGH THE C ENTRY
POINTS of libkmagpu (kmagpu_comm_init / kmagpu_allreduce_scores / kmagpu_allreduce_matrix: ncclAllReduce inside the
library, in place in HBM). torch.distributed only carries the 128-byte NCCL id and gathers the streams for the check.
  1. stage 2 + alignment pass on a rank's shard of the reads (single-end and paired: shards never split a pair), ConClave
     score arrays all-reduced on the device -> equal to the single-process oracle; frag_raw streams concatenated in rank
     order -> the single-process stream; ConClave then reads the device-resident sums (scores = NULL);
  2. traceback alignment + base counts on a rank's shard of the fragment records, count matrix all-reduced in place ->
     equal to the single-process oracle matrix.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dist_check.py"""
import os, sys, tempfile, pathlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as td
from kma_b200 import api, dist
from tests import util


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    wd = pathlib.Path(tempfile.gettempdir()) / "kma_b200_distcheck"
    if rank == 0:
        wd.mkdir(exist_ok=True)
        from tests.test_oracle_trace import make_frags
        prefix, frags = make_frags(wd, 91, 150, 0.02, 0.02, n=4000)
        frags.tofile(wd / "frags.bin")
    td.barrier()
    prefix = str(wd / "db")
    # 1. stages 2 + 3a
    with util.golden_dir() as g:
        s1 = np.fromfile(f"{g}/s1.bin", dtype=np.uint8)
        db = api.TemplateDB(f"{g}/db", device=local)
        db.comm_init_torch()
        frag, a, u, n, ms_ar = dist.map_sharded_device(db, s1, rank, world)
        whole = dist.gather_streams(frag, dst=0)
        # ConClave on the device-resident global sums == ConClave on the downloaded ones
        f_dev, w_dev, _, _, _ = db.conclave_from_align(None, None, out=np.empty(4 * len(frag) + 4096, np.uint8))
        f_host, w_host, _, _, _ = db.conclave_from_align(a, u, out=np.empty(4 * len(frag) + 4096, np.uint8))
        assert f_dev.tobytes() == f_host.tobytes() and np.array_equal(w_dev, w_host), "ConClave on device-resident sums differs"
        times = [db.allreduce_scores(download=False)[2] for _ in range(20)]   # warm, repeated (sums grow; only timing)
        db.close()
        if rank == 0:
            s2 = util.oracle_seed_stream(f"{g}/db", s1)
            ofrag, oa, ou, _, _ = util.oracle_align_stream(f"{g}/db", s2, want_cand=False)
            assert whole == ofrag, "rank-ordered frag_raw differs from the single-process stream"
            assert np.array_equal(a, oa) and np.array_equal(u, ou), "all-reduced ConClave sums differ"
            print(f"scores all-reduce ({2 * len(a)} u64): first {ms_ar:.3f} ms, warm min {min(times):.4f} ms / median {sorted(times)[10]:.4f} ms", flush=True)
        # paired reads: the shards must not split a pair
        from kma_b200 import synth, records
        tb = [util.template_bases(f"{g}/db", t) for t in range(1, 9)]
        r1, r2 = synth.paired_reads(5, tb, 3001)
        sp = records.stage1_pairs_fast(np.asarray(r1), np.asarray(r2))
        db = api.TemplateDB(f"{g}/db", device=local)
        db.comm_init_torch()
        p = api.default_params()
        fragp, ap, up, npairs, _ = dist.map_sharded_device(db, sp, rank, world, params=p)
        wholep = dist.gather_streams(fragp, dst=0)
        db.close()
        if rank == 0:
            s2 = util.oracle_seed_stream(f"{g}/db", sp)
            ofrag, oa, ou, _, _ = util.oracle_align_stream(f"{g}/db", s2, want_cand=False, one2one=False)
            assert wholep == ofrag and np.array_equal(ap, oa) and np.array_equal(up, ou), "paired shards differ from the single-process run"
    # 2. assembly pass: base counts
    frags = np.fromfile(wd / "frags.bin", dtype=np.uint8)
    off = api.record_offsets(3, frags)
    lo, hi = dist.shard_records(off, rank, world)
    db = api.TemplateDB(prefix, device=local)
    p = api.default_params()
    p.one2one = 1
    p.matrix = 1
    db.matrix_reset()
    db.assemble_align_batch(frags[lo:hi], p)
    db.comm_init_torch()
    ms = db.allreduce_matrix()       # ncclAllReduce inside the library, in place
    total = db.matrix_download()     # after the in-place reduce every rank holds the sum (clamped to 65535 on the way out)
    nent = total.size
    db.close()
    if rank == 0:
        want = util.oracle_matrix(prefix, frags, util.oracle_trace(prefix, frags))
        assert np.array_equal(total, want), "all-reduced base counts differ from the single-process matrix"
        print(f"dist_check ok: world {world}, reads {n} on rank 0, matrix {nent} uint32 all-reduced in {ms:.3f} ms, counts {int(want.sum())}", flush=True)
    td.barrier()
    td.destroy_process_group()


main()
