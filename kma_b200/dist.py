"""Multi-GPU plumbing of the mapping core: one process per GPU, reads sharded by rank, database
replicated, no data-path collective. The only exchange of stages 2 + 3a is the sum over ranks of the two ConClave
accumulators `alignment_scores[DB_size]` / `uniq_alignment_scores[DB_size]` (runkma.c:98-99, updatescores.c:228/276)
that ConClave's choice pass reads globally (conclave.c:80-123) -- one all-reduce over `torch.distributed` (NCCL over
NVLink on GPUs, gloo in the CPU tests). frag_raw streams stay per rank; concatenated in rank order they are the
single-process stream because shards are contiguous slices of the record stream."""
from __future__ import annotations

import numpy as np

from . import api


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous, balanced split of n records: rank r gets [lo, hi)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def group_starts(stage: int, buf, off: np.ndarray) -> np.ndarray:
    """indices of the records a shard may start at: never the second mate of a pair. Stage 1 marks the first mate with a
    negative header length (runinput.c:789), stage 2 writes it without templates (printPair, ankers.c:150)."""
    n = len(off) - 1
    if n == 0 or stage not in (1, 2):
        return np.arange(n + 1)
    field = 12 if stage == 1 else 16
    v = np.ndarray(n, dtype="<i4", buffer=np.ascontiguousarray(
        np.stack([buf[off[:-1].astype(np.int64) + field + b] for b in range(4)], axis=1)).tobytes())
    first_mate = v < 0 if stage == 1 else v == 0
    second = np.zeros(n + 1, dtype=bool)
    second[1:n] = first_mate[: n - 1]       # the record after a first mate is its mate
    return np.flatnonzero(~second)


def shard_stream(stage: int, buf, rank: int, world: int) -> np.ndarray:
    """the slice of whole records of a stage-1 / stage-2 stream that `rank` maps; pairs are never split"""
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    off = api.record_offsets(stage, buf)
    starts = group_starts(stage, buf, off)   # groups = single reads and pairs; the last entry is the end of the stream
    lo, hi = shard_bounds(len(starts) - 1, rank, world)
    return buf[int(off[starts[lo]]):int(off[starts[hi]])]


def allreduce_scores(alignment_scores: np.ndarray, uniq_alignment_scores: np.ndarray, device=None):
    """sum the ConClave accumulators over all ranks (in place on copies; returns the two reduced uint64 arrays)"""
    import torch
    import torch.distributed as dist
    both = np.concatenate([alignment_scores, uniq_alignment_scores]).astype(np.uint64)
    t = torch.from_numpy(both.view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    out = t.cpu().numpy().view(np.uint64)
    n = len(alignment_scores)
    return out[:n].copy(), out[n:].copy()


def allreduce_matrix(counts):
    """Second exchange of the path: the per-position base counts of the assembly pass summed over ranks.
    `counts` is the UNSATURATED matrix -- TemplateDB.matrix_tensor() (device, int32, reduced in place over NCCL) or a
    numpy integer array (CPU tests, gloo). +1 increments commute and the reference saturates at 65535
    (assembly.c:1436), so min(sum over ranks, 65535) is what one process would have produced. Exact for the template
    nodes of alnToMat and for all of alnToMatDense; alnToMat's insertion nodes depend on the read order and are not
    part of the matrix. Returns the clamped uint16 counts as a numpy array [positions, 6]."""
    import torch
    import torch.distributed as dist
    t = counts if isinstance(counts, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(counts).astype(np.int32).reshape(-1))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.clamp(max=65535).cpu().numpy().astype(np.uint16).reshape(-1, 6)


def shard_records(off: np.ndarray, rank: int, world: int) -> tuple[int, int]:
    """byte range of the whole records rank `rank` takes, given record offsets with the end offset appended"""
    lo, hi = shard_bounds(len(off) - 1, rank, world)
    return int(off[lo]), int(off[hi])


def gather_streams(local: bytes, dst: int = 0):
    """rank-ordered concatenation of the per-rank byte streams on rank `dst` (None elsewhere)"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    parts = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(local, parts, dst=dst)
    return b"".join(parts) if parts is not None else None


def map_sharded(compute, stage1, rank: int, world: int, device=None):
    """Map one batch across `world` ranks. `compute(stage1_shard) -> (frag_raw bytes, alignment_scores,
    uniq_alignment_scores, nreads)` is the per-rank pipeline (TemplateDB stage 2 + alignment pass on the rank's GPU).
    Returns (this rank's frag_raw bytes, globally reduced alignment_scores, uniq_alignment_scores, local read count)."""
    shard = shard_stream(1, stage1, rank, world)
    frag, a, u, n = compute(shard)
    a, u = allreduce_scores(a, u, device)
    return frag, a, u, n


def gpu_pipeline(db: "api.TemplateDB", params=None):
    """the per-rank compute of map_sharded on a TemplateDB: stage 2 and the alignment pass chained in HBM"""
    def run(stage1_shard):
        n = db.seed_upload(np.ascontiguousarray(stage1_shard))
        db.seed_run(params)
        db.align_from_seed()
        db.align_run(params)
        frag, a, u, _ = db.align_download()
        return frag.tobytes(), a, u, n
    return run


def map_sharded_device(db: "api.TemplateDB", stage1, rank: int, world: int, params=None):
    """map_sharded with the exchange inside the library: the ConClave sums stay in HBM, are all-reduced in place by NCCL on
    the library's stream (kmagpu_allreduce_scores) and are what a following db.conclave_from_align(None, None) reads.
    The handle needs db.comm_init(...) first. Returns (frag_raw bytes, alignment_scores, uniq_alignment_scores, reads, all-reduce ms)."""
    shard = shard_stream(1, stage1, rank, world)
    db.scores_reset()
    n = db.seed_upload(np.ascontiguousarray(shard))
    db.seed_run(params)
    db.align_from_seed()
    db.align_run(params)
    frag, _, _, _ = db.align_download()
    a, u, ms = db.allreduce_scores()
    return frag.tobytes(), a, u, n, ms
