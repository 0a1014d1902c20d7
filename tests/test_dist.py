"""CPU tests of the N>1 path (world size 2, gloo): read sharding, the ConClave score all-reduce and the rank-ordered
frag_raw concatenation of kma_b200/dist.py. The per-rank compute is injected; here it is the oracle (the checker),
on the GPU box it is dist.gpu_pipeline(TemplateDB)."""
import os
import socket

import numpy as np
import pytest

from kma_b200 import api, dist
from tests import util


def _oracle_compute(prefix):
    def run(shard):
        shard = np.ascontiguousarray(shard)
        s2 = util.oracle_seed_stream(prefix, shard) if len(shard) else np.zeros(0, np.uint8)
        n = len(api.record_offsets(1, shard)) - 1
        frag, a, u, _, _ = util.oracle_align_stream(prefix, s2, want_cand=False)
        return frag, a, u, n
    return run


def _worker(rank, world, port, prefix, s1_path, out_dir):
    import torch.distributed as td
    td.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    s1 = np.fromfile(s1_path, dtype=np.uint8)
    frag, a, u, n = dist.map_sharded(_oracle_compute(prefix), s1, rank, world)
    whole = dist.gather_streams(frag, dst=0)
    if rank == 0:
        open(os.path.join(out_dir, "frag.bin"), "wb").write(whole)
        np.save(os.path.join(out_dir, "a.npy"), a)
        np.save(os.path.join(out_dir, "u.npy"), u)
    np.save(os.path.join(out_dir, f"n{rank}.npy"), np.array([n]))
    td.barrier()
    td.destroy_process_group()


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 100, 101):
        for w in (1, 2, 3, 8):
            spans = [dist.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_record_walk_matches_parser():
    with util.golden_dir() as g:
        s1 = np.fromfile(f"{g}/s1.bin", dtype=np.uint8)
        s2 = np.fromfile(f"{g}/s2.bin", dtype=np.uint8)
        o1, o2 = api.record_offsets(1, s1), api.record_offsets(2, s2)
        from kma_b200 import records
        assert len(o2) - 1 == len(records.parse_stage2(s2))
        assert int(o1[-1]) == len(s1) and int(o2[-1]) == len(s2) - 4   # stage 2 ends with the int32 terminator
        # a truncated tail is not a whole record
        assert len(api.record_offsets(1, s1[:-3])) == len(o1) - 1
        # shards are contiguous and cover the stream
        parts = [dist.shard_stream(1, s1, r, 3) for r in range(3)]
        assert b"".join(p.tobytes() for p in parts) == s1.tobytes()


def test_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    with util.golden_dir() as g:
        prefix, s1_path = f"{g}/db", f"{g}/s1.bin"
        s1 = np.fromfile(s1_path, dtype=np.uint8)
        want_frag, want_a, want_u, n_all = _oracle_compute(prefix)(s1)
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        mp.spawn(_worker, args=(2, port, prefix, s1_path, str(tmp_path)), nprocs=2, join=True)
    got = open(tmp_path / "frag.bin", "rb").read()
    assert got == want_frag
    assert np.array_equal(np.load(tmp_path / "a.npy"), want_a) and np.array_equal(np.load(tmp_path / "u.npy"), want_u)
    assert int(np.load(tmp_path / "n0.npy")[0]) + int(np.load(tmp_path / "n1.npy")[0]) == n_all


def _matrix_worker(rank, world, port, prefix, frags_path, out_dir):
    import torch.distributed as td
    td.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    frags = np.fromfile(frags_path, dtype=np.uint8)
    off = api.record_offsets(3, frags)
    lo, hi = dist.shard_records(off, rank, world)
    mine = frags[lo:hi]
    trace = util.oracle_trace(prefix, mine)
    counts = util.oracle_matrix(prefix, mine, trace, dense=True, saturate=False)
    total = dist.allreduce_matrix(counts)
    if rank == 0:
        np.save(os.path.join(out_dir, "mat.npy"), total)
    td.barrier()
    td.destroy_process_group()


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
def test_matrix_allreduce_world_size_2_gloo(tmp_path):
    """fragment records sharded over two ranks, per-rank base counts summed and clamped = the single-process matrix"""
    import torch.multiprocessing as mp
    from tests.test_oracle_trace import make_frags
    prefix, frags = make_frags(tmp_path, 61, 150, 0.02, 0.02, n=600)
    frags_path = str(tmp_path / "frags.bin")
    frags.tofile(frags_path)
    want = util.oracle_matrix(prefix, frags, util.oracle_trace(prefix, frags), dense=True)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_matrix_worker, args=(2, port, prefix, frags_path, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "mat.npy")
    assert want.sum() > 1000 and np.array_equal(got, want)
