"""Multi-GPU check under torchrun (one rank per GPU, NCCL): the two exchanges of the path on real devices.
  1. stage 2 + alignment pass on a rank's shard of the reads, ConClave score arrays all-reduced -> equal to the
     single-process oracle; frag_raw streams concatenated in rank order -> the single-process stream;
  2. traceback alignment + base counts on a rank's shard of the fragment records, count matrix all-reduced in place over
     NCCL through the zero-copy view of the library's device buffer -> equal to the single-process oracle matrix.
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
        frag, a, u, n = dist.map_sharded(dist.gpu_pipeline(db), s1, rank, world, device=torch.device("cuda", local))
        whole = dist.gather_streams(frag, dst=0)
        db.close()
        if rank == 0:
            s2 = util.oracle_seed_stream(f"{g}/db", s1)
            ofrag, oa, ou, _, _ = util.oracle_align_stream(f"{g}/db", s2, want_cand=False)
            assert whole == ofrag, "rank-ordered frag_raw differs from the single-process stream"
            assert np.array_equal(a, oa) and np.array_equal(u, ou), "all-reduced ConClave sums differ"
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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = db.matrix_tensor()
    torch.cuda.synchronize()
    e0.record()
    total = dist.allreduce_matrix(t)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    mine = db.matrix_download()      # after the in-place reduce every rank holds the sum
    db.close()
    assert np.array_equal(mine, total)
    if rank == 0:
        want = util.oracle_matrix(prefix, frags, util.oracle_trace(prefix, frags))
        assert np.array_equal(total, want), "all-reduced base counts differ from the single-process matrix"
        print(f"dist_check ok: world {world}, reads {n} on rank 0, matrix {t.numel()} int32 all-reduced in {ms:.3f} ms "
              f"(incl. clamp + D2H), counts {int(want.sum())}", flush=True)
    td.barrier()
    td.destroy_process_group()


main()
