"""ctypes binding of libkmagpu.so (include/kmagpu.h) -- the host-side mirror of the reference's
stage drivers. Names follow the reference: `save_kmers_batch` (kmers.c:51) is stage 2.

There is no CPU fallback: if the CUDA library is missing or no device is present every compute
call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkmagpu.so")


class KmaGpuError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("M", C.c_int32), ("MM", C.c_int32), ("U", C.c_int32), ("W1", C.c_int32), ("Wl", C.c_int32),
                ("Mn", C.c_int32), ("PE", C.c_int32), ("d", C.c_int32 * 25), ("exhaustive", C.c_int32),
                ("mq", C.c_int32), ("one2one", C.c_int32), ("minlen", C.c_int32), ("kmerscan", C.c_int32), ("matrix", C.c_int32), ("apm", C.c_int32), ("counters", C.c_int32),
                ("ts", C.c_int32), ("lc", C.c_int32), ("scoreT", C.c_double), ("minFrac", C.c_double), ("mrc", C.c_double), ("coverT", C.c_double)]


class DbInfo(C.Structure):
    _fields_ = [("DB_size", C.c_int32), ("kmersize", C.c_int32), ("kmerindex", C.c_int32), ("mega", C.c_int32),
                ("size", C.c_uint64), ("n", C.c_uint64), ("v_index", C.c_uint64), ("device_bytes", C.c_uint64),
                ("seq_bases", C.c_uint64)]


class SeedStats(C.Structure):
    _fields_ = [("reads", C.c_int64), ("mapped", C.c_int64), ("read_words", C.c_int64), ("lookups", C.c_int64),
                ("hits", C.c_int64), ("list_fetches", C.c_int64), ("list_ids", C.c_int64),
                ("overflow_reads", C.c_int64), ("ms_seed", C.c_float), ("ms_emit", C.c_float),
                ("ms_h2d", C.c_float), ("ms_total", C.c_float), ("launches", C.c_int32), ("reserved", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class AlignStats(C.Structure):
    _fields_ = [("reads", C.c_int64), ("tasks", C.c_int64), ("frags", C.c_int64), ("mems", C.c_int64),
                ("nw_full_calls", C.c_int64), ("nw_band_calls", C.c_int64), ("nw_full_cells", C.c_int64),
                ("nw_band_cells", C.c_int64), ("nw_steps", C.c_int64), ("overflow_tasks", C.c_int64),
                ("index_probes", C.c_int64), ("mem_bases", C.c_int64), ("read_bytes", C.c_int64),
                ("ms_prep", C.c_float), ("ms_align", C.c_float), ("ms_reduce", C.c_float), ("ms_h2d", C.c_float),
                ("ms_total", C.c_float), ("launches", C.c_int32), ("reserved", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class IngestParams(C.Structure):
    _fields_ = [("fastq", C.c_int32), ("paired", C.c_int32), ("min_phred", C.c_int32), ("phred_scale", C.c_int32),
                ("minlen", C.c_int32), ("maxlen", C.c_int32), ("min_q", C.c_int32), ("hardmask_q", C.c_int32), ("trans", C.c_uint8 * 256),
                ("prob", C.c_double * 256)]


def quality_prob() -> np.ndarray:
    """prob[q], the error probability of phred score q, as the CLI's table holds it (kma.c:219): pow(10, -0.1 * q) written
    with 32 decimal places and read back (entries below 1e-15 keep fewer than 17 significant digits, and -0.1 * q is not
    -q / 10: a third of the entries sit one ulp off the correctly rounded value). tests/test_oracle_stage1.py compares
    the result with the reference's literals. Host glue like to2bit(): the reference passes its own table."""
    import math
    return np.array([float("%.32f" % math.pow(10, -0.1 * q)) for q in range(256)], dtype=np.float64)


def to2bit() -> np.ndarray:
    """the byte -> code table the CLI builds (kma.c:1439-1482): 0-3 bases (IUPAC codes fold onto one of their bases),
    4 = N / X, 16 = newline, 8 = everything else. Host glue: the reference passes its own table."""
    t = np.full(256, 8, dtype=np.uint8)
    t[ord("\n")] = 16
    for v, chars in enumerate(("AaRrMmDd", "CcYyBb", "GgSsKkVv", "TtWwHhUu", "NnXx")):
        for ch in chars:
            t[ord(ch)] = v
    return t


def fasta_unwrap(text, trans: np.ndarray | None = None, eof: bool = True):
    """host only: multi-line FASTA -> 2-line FASTA as FileBuffgetFsa reads it (seqparse.c:66-160) -> (bytes, input bytes used)"""
    trans = to2bit() if trans is None else trans
    buf = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else text
    out = np.empty(len(buf) + 1, dtype=np.uint8)
    used = C.c_size_t()
    L = lib()
    L.kmagpu_fasta_unwrap.restype = C.c_int64
    L.kmagpu_fasta_unwrap.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    n = L.kmagpu_fasta_unwrap(buf.ctypes.data, len(buf), trans.ctypes.data, int(eof), out.ctypes.data, len(out), C.byref(used))
    if n < 0:
        raise KmaGpuError(L.kmagpu_last_error().decode())
    return out[:n].tobytes(), used.value


def fastx_split(text, fastq: bool = True, trans: np.ndarray | None = None):
    """host only: line structure of a FASTQ / FASTA chunk -> (uint32 fields[n, 5], bytes used)"""
    trans = to2bit() if trans is None else trans
    buf = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else text
    used = C.c_size_t()
    n = lib().kmagpu_fastx_split(buf.ctypes.data, len(buf), int(fastq), trans.ctypes.data, None, 0, C.byref(used))
    if n < 0:
        raise KmaGpuError(lib().kmagpu_last_error().decode())
    fields = np.zeros((n, 5), dtype=np.uint32)
    lib().kmagpu_fastx_split(buf.ctypes.data, len(buf), int(fastq), trans.ctypes.data, fields.ctypes.data, n, C.byref(used))
    return fields, used.value


def fastx_split_parallel(text, threads: int = 8, fastq: bool = True, trans: np.ndarray | None = None) -> np.ndarray:
    """fastx_split by `threads` host threads over byte ranges of the chunk (cut at record starts found by
    kmagpu_fastx_sync; the C calls run without the GIL). -> uint32 fields[n, 5] with offsets into the whole chunk"""
    from concurrent.futures import ThreadPoolExecutor
    trans = to2bit() if trans is None else trans
    buf = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else text
    buf = buf.numpy() if hasattr(buf, "numpy") else buf
    nb = len(buf)
    L = lib()
    cuts = sorted({L.kmagpu_fastx_sync(buf.ctypes.data, nb, int(fastq), (nb * i) // threads) for i in range(threads)} | {nb})

    def part(i):
        lo, hi = cuts[i], cuts[i + 1]
        f, used = fastx_split(buf[lo:hi], fastq, trans)
        if used != hi - lo:
            raise KmaGpuError(f"text range {lo}..{hi} does not end at a record boundary")
        f[:, [0, 2, 4]] += np.uint32(lo)
        return f

    with ThreadPoolExecutor(max_workers=threads) as ex:
        parts = list(ex.map(part, range(len(cuts) - 1)))
    return np.concatenate(parts) if parts else np.zeros((0, 5), dtype=np.uint32)


class ConsensusParams(C.Structure):
    _fields_ = [("bcd", C.c_int32), ("caller", C.c_int32), ("significance", C.c_int32), ("reserved", C.c_int32),
                ("support", C.c_double), ("chi2_min", C.c_double)]


CONSENSUS_STATS = np.dtype([("depth", "<u8"), ("depthVar", "<u8"), ("len", "<u4"), ("aln_len", "<u4"), ("cover", "<u4"),
                            ("reserved", "<u4")])

_lib = None


def lib():
    """Load libkmagpu.so (built in-tree by __graft_entry__.build()). Fails loudly when absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KmaGpuError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        L.kmagpu_last_error.restype = C.c_char_p
        L.kmagpu_db_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        L.kmagpu_db_close.argtypes = [C.c_void_p]
        L.kmagpu_db_close.restype = None
        L.kmagpu_db_clone.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.kmagpu_comm_unique_id.argtypes = [C.c_void_p, C.c_size_t]
        L.kmagpu_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.kmagpu_comm_destroy.argtypes = [C.c_void_p]
        L.kmagpu_comm_destroy.restype = None
        L.kmagpu_scores_reset.argtypes = [C.c_void_p]
        L.kmagpu_allreduce_scores.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        L.kmagpu_allreduce_matrix.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.kmagpu_allreduce_u64.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.kmagpu_db_get_info.argtypes = [C.c_void_p, C.POINTER(DbInfo)]
        L.kmagpu_default_params.argtypes = [C.POINTER(Params)]
        L.kmagpu_default_params.restype = None
        L.kmagpu_seed_batch.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                        C.POINTER(C.c_size_t), C.POINTER(C.c_int64), C.POINTER(SeedStats)]
        L.kmagpu_seed_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int64)]
        L.kmagpu_seed_run.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(SeedStats)]
        L.kmagpu_seed_download.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.kmagpu_memscore_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p,
                                            C.c_void_p, C.POINTER(C.c_int64)]
        L.kmagpu_conclave_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                            C.POINTER(C.c_size_t), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
        L.kmagpu_matrix_reset.argtypes = [C.c_void_p]
        L.kmagpu_matrix_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.kmagpu_matrix_download.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.kmagpu_lookup_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.kmagpu_fastx_split.restype = C.c_int64
        L.kmagpu_fastx_split.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.kmagpu_fastx_sync.restype = C.c_size_t
        L.kmagpu_fastx_sync.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_size_t]
        L.kmagpu_stage1_batch.argtypes = [C.c_void_p, C.POINTER(IngestParams), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                          C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int64), C.POINTER(C.c_float)]
        L.kmagpu_stage1_text.argtypes = [C.c_void_p, C.POINTER(IngestParams), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int,
                                         C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                         C.POINTER(C.c_int64), C.POINTER(C.c_float)]
        L.kmagpu_chi2_threshold.restype = C.c_double
        L.kmagpu_chi2_threshold.argtypes = [C.c_double, C.c_void_p]
        L.kmagpu_consensus.argtypes = [C.c_void_p, C.c_int32, C.POINTER(ConsensusParams), C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_size_t, C.c_void_p, C.POINTER(C.c_float)]
        L.kmagpu_align_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int64)]
        L.kmagpu_align_from_seed.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.kmagpu_align_run.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int, C.POINTER(AlignStats)]
        L.kmagpu_align_download.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.kmagpu_align_batch.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                         C.POINTER(C.c_size_t), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                         C.POINTER(C.c_size_t), C.POINTER(AlignStats)]
        L.kmagpu_nw_batch.argtypes = [C.c_void_p, C.POINTER(Params), C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                      C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_float)]
        L.kmagpu_trace_batch.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                         C.POINTER(C.c_size_t), C.POINTER(C.c_int64), C.POINTER(AlignStats)]
        L.kmagpu_memscore_from_seed.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
        L.kmagpu_conclave_resident.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
        L.kmagpu_conclave_mode.argtypes = [C.c_void_p, C.c_int]
        L.kmagpu_conclave_from_align.argtypes = L.kmagpu_conclave_resident.argtypes
        L.kmagpu_trace_from_conclave.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int64),
                                                 C.POINTER(AlignStats)]
        L.kmagpu_record_walk.restype = C.c_int64
        L.kmagpu_record_walk.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise KmaGpuError(lib().kmagpu_last_error().decode(errors="replace"))


def default_params() -> Params:
    p = Params()
    lib().kmagpu_default_params(C.byref(p))
    return p


def preset(name: str) -> dict:
    """What the reference's presets bind (kma.c:1100-1240), as arguments for this package's calls. Host glue like
    default_params(): `params` for the stage-2 / stage-3 / traceback calls, `ingest` for run_input_text / run_input_batch,
    `consensus` for TemplateDB.consensus, `conclave_lc` for conclave_mode.
      ont: chain scan, -lc, -proxi -0.9, -ts 2, -eq 10, -mrs 0.25, -mrc 0.7, -mct 0.1, -bcNano, -bc 0.7, -bcd 10
      ill: -1t1, -lc, -proxi -0.98, -mrc 0.1, -bc 0.9, -bcd 10
      asm: chain scan, -lc, -proxi -0.9, -ts 2, -mrs 0.25, -mrc 0.7, -mct 0.1, -bc 0.5, -p 0.5, -bcd 1"""
    p = default_params()
    p.kmerscan, p.one2one = 1, 0   # the CLI's default scan is save_kmers_chain (savekmers.c:40)
    p.apm = 1                      # -apm u, the default pairing
    ingest, cons = {}, dict(bcd=1, evalue=0.05, caller=0, significance=0, support=0.0)
    if name == "ont":
        p.lc, p.minFrac, p.ts, p.scoreT, p.mrc, p.coverT = 1, -0.9, 2, 0.25, 0.7, 0.1
        ingest = dict(min_q=10)
        cons.update(bcd=10, caller=3, significance=2, support=0.7)
    elif name == "ill":
        p.kmerscan, p.one2one = 0, 1
        p.lc, p.minFrac, p.mrc = 1, -0.98, 0.1
        cons.update(bcd=10, significance=2, support=0.9)
    elif name == "asm":
        p.lc, p.minFrac, p.ts, p.scoreT, p.mrc, p.coverT = 1, -0.9, 2, 0.25, 0.7, 0.1
        cons.update(bcd=1, evalue=0.5, significance=2, support=0.5)
    else:
        raise ValueError(f"preset {name!r}: ont, ill or asm")
    return dict(params=p, ingest=ingest, consensus=cons, conclave_lc=True)


def _ptr(a):
    """address of a numpy array or torch tensor's storage"""
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.ctypes.data


class TemplateDB:
    """HBM-resident template database (hashMapKMA_load + .length.b/.seq.b of runKMA)."""

    def __init__(self, prefix: str, device: int = 0, _clone_of: "TemplateDB | None" = None):
        self._h = C.c_void_p()
        if _clone_of is not None:
            _check(lib().kmagpu_db_clone(_clone_of._h, C.byref(self._h)))
        else:
            _check(lib().kmagpu_db_open(os.fsencode(prefix), device, C.byref(self._h)))
        self.info = DbInfo()
        _check(lib().kmagpu_db_get_info(self._h, C.byref(self.info)))
        self.device = device
        self.prefix = prefix
        self._lengths = None

    @property
    def lengths(self) -> np.ndarray:
        """.length.b: lengths[t] of template t >= 1 (lengths[0] = k of the alignment index)"""
        if self._lengths is None:
            self._lengths = np.fromfile(self.prefix + ".length.b", dtype=np.int32)[1:]
        return self._lengths

    # ---- multi-GPU exchange inside the library (kmagpu_comm.cu): NCCL all-reduces in place in HBM
    @staticmethod
    def comm_unique_id() -> bytes:
        """rank 0: the 128-byte NCCL id every rank passes to comm_init"""
        buf = C.create_string_buffer(128)
        _check(lib().kmagpu_comm_unique_id(buf, 128))
        return buf.raw

    def comm_init(self, uid: bytes, rank: int, world: int):
        _check(lib().kmagpu_comm_init(self._h, uid, rank, world))

    def comm_init_torch(self):
        """comm_init with the id broadcast over an initialised torch.distributed process group (host plumbing only)"""
        import torch.distributed as td
        rank, world = td.get_rank(), td.get_world_size()
        box = [self.comm_unique_id() if rank == 0 else None]
        td.broadcast_object_list(box, src=0)
        self.comm_init(box[0], rank, world)

    def scores_reset(self):
        """start the run-wide ConClave accumulators of this handle on the device"""
        _check(lib().kmagpu_scores_reset(self._h))

    def allreduce_scores(self, download=True):
        """sum the device-resident accumulators over ranks, in place -> (alignment_scores, uniq_alignment_scores, all-reduce ms)"""
        ms = C.c_float()
        a = np.zeros(self.info.DB_size, np.uint64) if download else None
        u = np.zeros(self.info.DB_size, np.uint64) if download else None
        _check(lib().kmagpu_allreduce_scores(self._h, a.ctypes.data if download else None, u.ctypes.data if download else None, C.byref(ms)))
        return a, u, ms.value

    def allreduce_matrix(self) -> float:
        ms = C.c_float()
        _check(lib().kmagpu_allreduce_matrix(self._h, C.byref(ms)))
        return ms.value

    def allreduce_u64(self, arr: np.ndarray):
        assert arr.dtype == np.uint64 and arr.flags.c_contiguous
        _check(lib().kmagpu_allreduce_u64(self._h, arr.ctypes.data, arr.size))
        return arr

    def clone(self) -> "TemplateDB":
        """a further handle on the same HBM image (kmagpu_db_clone): own stream and batch buffers, for another host thread"""
        return TemplateDB(self.prefix, self.device, _clone_of=self)

    def close(self):
        if self._h:
            lib().kmagpu_db_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- stage 1 -----------------------------------------------------------------------------
    def run_input_batch(self, text, fields: np.ndarray, fastq=True, paired=False, min_phred=20, phred_scale=33, minlen=16,
                        maxlen=2147483647, trans: np.ndarray | None = None, download=True, text2=None, min_q=0, hardmask_q=0):
        """FASTQ / FASTA text + its line structure (fastx_split) -> stage-1 records (run_input / run_input_PE per read:
        translation, end trim, -ml / -xl, pairing rule, compDNA, printFsa). The stream stays on the device as the input
        of seed_run(); download=False skips the copy back. text2: the second file's chunk of a pair of files (its fields
        count their offsets from len(text) on). -> (stage-1 bytes | None, count, kernel ms)"""
        ip = IngestParams()
        ip.fastq, ip.paired, ip.min_phred, ip.phred_scale, ip.minlen, ip.maxlen = int(fastq), int(paired), min_phred, phred_scale, minlen, maxlen
        tab = to2bit() if trans is None else np.ascontiguousarray(trans, dtype=np.uint8)   # held in a name while memmove reads it
        C.memmove(ip.trans, tab.ctypes.data, 256)
        ip.min_q, ip.hardmask_q = int(min_q), int(hardmask_q)   # -eq / -mi (phredStat, runinput.c:168-313)
        if min_q or hardmask_q:
            prob = quality_prob()   # held in a name: the temporary would be gone before memmove reads it
            C.memmove(ip.prob, prob.ctypes.data, 2048)
        buf = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else text
        fields = np.ascontiguousarray(fields, dtype=np.uint32)
        nbytes = int(buf.numel() if hasattr(buf, "numel") else buf.size)
        buf2 = np.frombuffer(text2, dtype=np.uint8) if isinstance(text2, (bytes, bytearray)) else text2
        nbytes2 = 0 if buf2 is None else int(buf2.numel() if hasattr(buf2, "numel") else buf2.size)
        out = np.empty(nbytes + nbytes2 + 64 if download else 0, dtype=np.uint8)
        ob, cnt, ms = C.c_size_t(), C.c_int64(), C.c_float()
        _check(lib().kmagpu_stage1_batch(self._h, C.byref(ip), _ptr(buf), nbytes, _ptr(buf2) if nbytes2 else None, nbytes2,
                                         fields.ctypes.data, len(fields),
                                         out.ctypes.data if download else None, len(out), C.byref(ob), C.byref(cnt), C.byref(ms)))
        return (out[: ob.value] if download else None), cnt.value, ms.value

    def run_input_text(self, text, text2=None, fastq=True, min_phred=20, phred_scale=33, minlen=16, maxlen=2147483647,
                       trans: np.ndarray | None = None, download=True, eof=True, min_q=0, hardmask_q=0):
        """run_input / run_input_PE on chunks of file text with the record splitter on the device as well
        (kmagpu_stage1_text). text2: the second file's chunk (pairs by record index). -> (stage-1 bytes | None, count,
        kernel ms, bytes used of text, bytes used of text2)"""
        ip = IngestParams()
        ip.fastq, ip.paired, ip.min_phred, ip.phred_scale, ip.minlen, ip.maxlen = int(fastq), int(text2 is not None), min_phred, phred_scale, minlen, maxlen
        tab = to2bit() if trans is None else np.ascontiguousarray(trans, dtype=np.uint8)   # held in a name while memmove reads it
        C.memmove(ip.trans, tab.ctypes.data, 256)
        ip.min_q, ip.hardmask_q = int(min_q), int(hardmask_q)   # -eq / -mi (phredStat, runinput.c:168-313)
        if min_q or hardmask_q:
            prob = quality_prob()   # held in a name: the temporary would be gone before memmove reads it
            C.memmove(ip.prob, prob.ctypes.data, 2048)
        buf = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else text
        buf2 = np.frombuffer(text2, dtype=np.uint8) if isinstance(text2, (bytes, bytearray)) else text2
        nb = int(buf.numel() if hasattr(buf, "numel") else buf.size)
        nb2 = 0 if buf2 is None else int(buf2.numel() if hasattr(buf2, "numel") else buf2.size)
        out = np.empty(nb + nb2 + 64 if download else 0, dtype=np.uint8)
        ob, cnt, ms, u1, u2 = C.c_size_t(), C.c_int64(), C.c_float(), C.c_size_t(), C.c_size_t()
        _check(lib().kmagpu_stage1_text(self._h, C.byref(ip), _ptr(buf) if nb else None, nb, _ptr(buf2) if buf2 is not None else None, nb2,
                                        int(eof), C.byref(u1), C.byref(u2), out.ctypes.data if download else None, len(out), C.byref(ob),
                                        C.byref(cnt), C.byref(ms)))
        return (out[: ob.value] if download else None), cnt.value, ms.value, u1.value, u2.value

    # --- stage 2 -----------------------------------------------------------------------------
    def save_kmers_batch(self, stage1, params: Params | None = None, out=None):
        """stage-1 records (uint8 array / pinned tensor) -> (stage-2 bytes incl. no terminator, nreads, stats)"""
        p = params or default_params()
        nbytes = int(stage1.numel() if hasattr(stage1, "numel") else stage1.size)
        if out is not None:
            cap = int(out.numel() if hasattr(out, "numel") else out.size)
            ob, nr, st = C.c_size_t(), C.c_int64(), SeedStats()
            _check(lib().kmagpu_seed_batch(self._h, C.byref(p), _ptr(stage1), nbytes, _ptr(out), cap,
                                           C.byref(ob), C.byref(nr), C.byref(st)))
            return out[: ob.value], nr.value, st
        # caller gave no buffer: run, then size the output exactly
        nr = self.seed_upload(stage1)
        st = self.seed_run(p)
        ob = C.c_size_t()
        lib().kmagpu_seed_download(self._h, None, 0, C.byref(ob))
        out = np.empty(ob.value + 8, dtype=np.uint8)
        return self.seed_download(out), nr, st

    def seed_upload(self, stage1):
        nbytes = int(stage1.numel() if hasattr(stage1, "numel") else stage1.size)
        nr = C.c_int64()
        _check(lib().kmagpu_seed_upload(self._h, _ptr(stage1), nbytes, C.byref(nr)))
        return nr.value

    def seed_run(self, params: Params | None = None):
        p = params or default_params()
        st = SeedStats()
        _check(lib().kmagpu_seed_run(self._h, C.byref(p), C.byref(st)))
        return st

    def seed_download(self, out):
        cap = int(out.numel() if hasattr(out, "numel") else out.size)
        ob = C.c_size_t()
        _check(lib().kmagpu_seed_download(self._h, _ptr(out), cap, C.byref(ob)))
        return out[: ob.value]

    # --- stage 3, alignment pass ---------------------------------------------------------------
    def align_upload(self, stage2):
        nbytes = int(stage2.numel() if hasattr(stage2, "numel") else stage2.size)
        nr = C.c_int64()
        _check(lib().kmagpu_align_upload(self._h, _ptr(stage2), nbytes, C.byref(nr)))
        return nr.value

    def align_from_seed(self):
        """align the stage-2 stream the last seed_run left in HBM (no host round trip)"""
        nr = C.c_int64()
        _check(lib().kmagpu_align_from_seed(self._h, C.byref(nr)))
        return nr.value

    def align_run(self, params: Params | None = None, want_cand=False):
        p = params or default_params()
        st = AlignStats()
        _check(lib().kmagpu_align_run(self._h, C.byref(p), int(want_cand), C.byref(st)))
        return st

    def align_out_bytes(self) -> int:
        """size of the frag_raw stream of the last align_run"""
        ob, cr = C.c_size_t(), C.c_size_t()
        lib().kmagpu_align_download(self._h, None, 0, C.byref(ob), None, None, None, 0, C.byref(cr))
        return ob.value

    def align_download(self, out=None, scores=None, want_cand=False):
        """-> (frag_raw bytes, alignment_scores, uniq_alignment_scores, cand rows or None); `scores` = a pair of
        uint64[DB_size] arrays to ADD into (the ConClave accumulators of runkma.c:98-99)"""
        ob, cr = C.c_size_t(), C.c_size_t()
        lib().kmagpu_align_download(self._h, None, 0, C.byref(ob), None, None, None, 0, C.byref(cr))
        if out is None:
            out = np.empty(ob.value + 8, dtype=np.uint8)
        cap = int(out.numel() if hasattr(out, "numel") else out.size)
        a, u = scores if scores is not None else (np.zeros(self.info.DB_size, np.uint64), np.zeros(self.info.DB_size, np.uint64))
        cand = np.empty((cr.value, 8), dtype=np.int32) if want_cand else None
        _check(lib().kmagpu_align_download(self._h, _ptr(out), cap, C.byref(ob), a.ctypes.data, u.ctypes.data,
                                           cand.ctypes.data if want_cand else None, cr.value, C.byref(cr)))
        return out[: ob.value], a, u, cand

    def align_scores(self, scores=None):
        """only the two ConClave score arrays of the last align_run; the frag_raw stream stays in HBM
        (conclave_resident(..., source="align"))"""
        a, u = scores if scores is not None else (np.zeros(self.info.DB_size, np.uint64), np.zeros(self.info.DB_size, np.uint64))
        ob, cr = C.c_size_t(), C.c_size_t()
        _check(lib().kmagpu_align_download(self._h, None, 0, C.byref(ob), a.ctypes.data, u.ctypes.data, None, 0, C.byref(cr)))
        return a, u

    def alnFrags_batch(self, stage2, params: Params | None = None, want_cand=False, scores=None):
        """alnFrags_threaded (alnfrags.c:2150) over a batch of stage-2 records ->
        (frag_raw bytes, alignment_scores, uniq_alignment_scores, cand rows, stats)"""
        self.align_upload(stage2)
        st = self.align_run(params, want_cand)
        frag, a, u, cand = self.align_download(scores=scores, want_cand=want_cand)
        return frag, a, u, cand, st

    def assemble_align_batch(self, frags, params: Params | None = None, download=True):
        """the alignment part of assemble_KMA's inner loop (assembly.c:1868-1961: anker_rc + KMA with traceback +
        acceptance) over per-template fragment records (frags.c:45-48) -> (output bytes, nrecords, stats); per record
        int32[12]{accepted, read_score, start, end, score, len, pos, match, tGaps, qGaps, turned, ncol} + t/s/q rows"""
        p = params or default_params()
        frags = np.ascontiguousarray(frags, dtype=np.uint8)
        cap = 64 + 60 * (len(frags) // 32 + 1) + 16 * len(frags) if download else 0
        out = np.empty(cap, dtype=np.uint8)
        ob, nr, st = C.c_size_t(), C.c_int64(), AlignStats()
        _check(lib().kmagpu_trace_batch(self._h, C.byref(p), frags.ctypes.data, len(frags), out.ctypes.data if download else None, cap,
                                        C.byref(ob), C.byref(nr), C.byref(st)))
        return (out[: ob.value] if download else None), nr.value, st

    # --- -mem_mode: k-mer score collection of runKMA_MEM ------------------------------------------
    def memscore_batch(self, stage2, scores=None):
        """update_Scores_MEM / _pe_MEM over a batch of stage-2 records -> (frag_raw bytes, alignment_scores, uniq_alignment_scores, nrecords)"""
        s2 = np.ascontiguousarray(np.frombuffer(stage2, dtype=np.uint8) if isinstance(stage2, (bytes, bytearray)) else stage2, dtype=np.uint8)
        DB = self.info.DB_size
        a, u = scores if scores is not None else (np.zeros(DB, np.uint64), np.zeros(DB, np.uint64))
        out = np.empty(8 * len(s2) + 4096, dtype=np.uint8)
        ob, nr = C.c_size_t(), C.c_int64()
        _check(lib().kmagpu_memscore_batch(self._h, s2.ctypes.data, len(s2), out.ctypes.data, len(out), C.byref(ob), a.ctypes.data,
                                           u.ctypes.data, C.byref(nr)))
        return out[: ob.value], a, u, nr.value

    def conclave_mode(self, length_corrected: bool):
        """ConClavePtr: False = runConClave, True = runConClave_lc (-lc)"""
        _check(lib().kmagpu_conclave_mode(self._h, int(length_corrected)))

    def softproxi_reset(self):
        """start the soft proximity sums of this database image (kmers.c:133-153); every seed_run with params.minFrac < 0 adds to them"""
        _check(lib().kmagpu_softproxi_reset(self._h))

    def softproxi_download(self) -> np.ndarray:
        s = np.zeros(self.info.DB_size, dtype=np.uint64)
        L = lib()
        L.kmagpu_softproxi_download.argtypes = [C.c_void_p, C.c_void_p]
        _check(L.kmagpu_softproxi_download(self._h, s.ctypes.data))
        return s

    def conclave_version(self, version: int, p_chisqr=None, scoreT: float = 0.5, evalue: float = 0.05, and_mode: bool = False):
        """-ConClave 2: the ConClave calls run runConClave2 / runConClave2_lc (conclave.c:386 / 749) over the batch they get (the
        whole run). p_chisqr: the caller's chi-square tail function as a C pointer double (*)(long double) (stdstat.c:136).
        version 1 restores runConClave."""
        L = lib()
        L.kmagpu_conclave_version.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p]
        _check(L.kmagpu_conclave_version(self._h, int(version), float(scoreT), float(evalue), int(and_mode), p_chisqr))

    def conclave_uniq_scores(self) -> np.ndarray:
        """the unique scores as the last ConClave call left them (runConClave2 adds to them, conclave.c:519)"""
        u = np.zeros(self.info.DB_size, dtype=np.uint64)
        L = lib()
        L.kmagpu_conclave_uniq_scores.argtypes = [C.c_void_p, C.c_void_p]
        _check(L.kmagpu_conclave_uniq_scores(self._h, u.ctypes.data))
        return u

    def memscore_from_seed(self, scores=None, download=True, cap=None):
        """memscore_batch on the stage-2 stream the last seed_run left in HBM; the frag_raw stream stays resident for
        conclave_resident. -> (frag_raw bytes | None, alignment_scores, uniq_alignment_scores, nrecords)"""
        DB = self.info.DB_size
        a, u = scores if scores is not None else (np.zeros(DB, np.uint64), np.zeros(DB, np.uint64))
        ob, nr = C.c_size_t(), C.c_int64()
        out = np.empty(int(cap) if (download and cap) else 0, dtype=np.uint8)
        if download and not cap:   # size it: run once without output, then fetch
            raise KmaGpuError("memscore_from_seed(download=True) needs cap (bytes of the output buffer)")
        _check(lib().kmagpu_memscore_from_seed(self._h, out.ctypes.data if download else None, len(out), C.byref(ob), a.ctypes.data,
                                               u.ctypes.data, C.byref(nr)))
        return (out[: ob.value] if download else None), a, u, nr.value

    def conclave_resident(self, alignment_scores, uniq_alignment_scores, totals=None, download=True, cap=None, source="memscore"):
        """conclave_batch on the frag_raw stream the last score collection (source="memscore") or the last align_run
        (source="align") left in HBM"""
        a = np.ascontiguousarray(alignment_scores, dtype=np.uint64)
        u = np.ascontiguousarray(uniq_alignment_scores, dtype=np.uint64)
        DB = self.info.DB_size
        w, fc, rc = totals if totals is not None else (np.zeros(DB, np.uint64), np.zeros(DB, np.uint32), np.zeros(DB, np.uint32))
        if download and not cap:
            raise KmaGpuError("conclave_resident(download=True) needs cap (bytes of the output buffer)")
        out = np.empty(int(cap) if download else 0, dtype=np.uint8)
        ob, nr = C.c_size_t(), C.c_int64()
        fn = lib().kmagpu_conclave_from_align if source == "align" else lib().kmagpu_conclave_resident
        _check(fn(self._h, a.ctypes.data, u.ctypes.data, out.ctypes.data if download else None, len(out),
                                              C.byref(ob), w.ctypes.data, fc.ctypes.data, rc.ctypes.data, C.byref(nr)))
        self._frag_bytes = ob.value
        return (out[: ob.value] if download else None), w, fc, rc, nr.value

    def conclave_from_align(self, alignment_scores=None, uniq_alignment_scores=None, out=None, totals=None):
        """ConClave on the frag_raw stream the last align_run of this handle left in HBM. Scores None: the run-wide sums the
        database image holds on the device (scores_reset ... allreduce_scores) -- they never left HBM. out: a (pinned)
        uint8 buffer for the per-template fragment stream, None: the fragments stay in HBM for trace_from_conclave.
        -> (fragment bytes | None, w_scores, fragmentCounts, readCounts, nrecords)"""
        DB = self.info.DB_size
        a = None if alignment_scores is None else np.ascontiguousarray(alignment_scores, dtype=np.uint64)
        u = None if uniq_alignment_scores is None else np.ascontiguousarray(uniq_alignment_scores, dtype=np.uint64)
        w, fc, rc = totals if totals is not None else (np.zeros(DB, np.uint64), np.zeros(DB, np.uint32), np.zeros(DB, np.uint32))
        cap = 0 if out is None else int(out.numel() if hasattr(out, "numel") else out.size)
        ob, nr = C.c_size_t(), C.c_int64()
        _check(lib().kmagpu_conclave_from_align(self._h, None if a is None else a.ctypes.data, None if u is None else u.ctypes.data,
                                                None if out is None else _ptr(out), cap, C.byref(ob), w.ctypes.data, fc.ctypes.data,
                                                rc.ctypes.data, C.byref(nr)))
        self._frag_bytes = ob.value
        return (None if out is None else out[: ob.value]), w, fc, rc, nr.value

    # --- ConClave choice pass + per-template bucketing --------------------------------------------
    def conclave_batch(self, frag_raw, alignment_scores, uniq_alignment_scores, totals=None, download=True):
        """runConClave (conclave.c:43) + printFrags (frags.c:30) over one chunk of frag_raw records with the GLOBAL score
        arrays -> (per-template fragment records incl. the -1 terminator, w_scores, fragmentCounts, readCounts, nrecords);
        `totals` = (w_scores u64, fragmentCounts u32, readCounts u32) arrays to add into"""
        fr = np.ascontiguousarray(np.frombuffer(frag_raw, dtype=np.uint8) if isinstance(frag_raw, (bytes, bytearray)) else frag_raw, dtype=np.uint8)
        a = np.ascontiguousarray(alignment_scores, dtype=np.uint64)
        u = np.ascontiguousarray(uniq_alignment_scores, dtype=np.uint64)
        DB = self.info.DB_size
        w, fc, rc = totals if totals is not None else (np.zeros(DB, np.uint64), np.zeros(DB, np.uint32), np.zeros(DB, np.uint32))
        out = np.empty(2 * len(fr) + 64 if download else 0, dtype=np.uint8)   # download=False: fragments stay in HBM for trace_from_conclave
        ob, nr = C.c_size_t(), C.c_int64()
        _check(lib().kmagpu_conclave_batch(self._h, fr.ctypes.data, len(fr), a.ctypes.data, u.ctypes.data,
                                           out.ctypes.data if download else None, len(out),
                                           C.byref(ob), w.ctypes.data, fc.ctypes.data, rc.ctypes.data, C.byref(nr)))
        self._frag_bytes = ob.value
        return (out[: ob.value] if download else None), w, fc, rc, nr.value

    def trace_from_conclave(self, params: Params | None = None, download=True):
        """assemble_align_batch on the fragment stream the last conclave_batch of this handle left in HBM; download=False:
        no row output, only the base counts (params.matrix) and the statistics -> (output bytes | None, nrecords, stats)"""
        p = params or default_params()
        cap = 0
        if download:
            cap = 64 + 16 * int(self._frag_bytes) + 60 * (int(self._frag_bytes) // 32 + 1)
        out = np.empty(cap, dtype=np.uint8)
        ob, nr, st = C.c_size_t(), C.c_int64(), AlignStats()
        _check(lib().kmagpu_trace_from_conclave(self._h, C.byref(p), out.ctypes.data if download else None, cap, C.byref(ob), C.byref(nr),
                                                C.byref(st)))
        return (out[: ob.value] if download else None), nr.value, st

    # --- base-count matrix of the assembly pass (alnToMat / alnToMatDense) ----------------------
    def matrix_reset(self):
        _check(lib().kmagpu_matrix_reset(self._h))

    def matrix_download(self, template: int = 0) -> np.ndarray:
        """uint16 [positions, 6] counts {A, C, G, T, N, gap} of one template (template > 0) or of the whole database"""
        n = C.c_size_t()
        _check(lib().kmagpu_matrix_download(self._h, int(template), None, 0, C.byref(n)))
        out = np.empty(n.value, dtype=np.uint16)
        _check(lib().kmagpu_matrix_download(self._h, int(template), out.ctypes.data, n.value, C.byref(n)))
        return out.reshape(-1, 6)

    def consensus(self, template: int = 0, bcd: int = 1, evalue: float = 0.05, caller: int = 0, significance: int = 0,
                  support: float = 0.0, p_chisqr=None, out=None):
        """callConsensus (assembly.c:1499) over the template nodes of the device matrix. caller: 0 baseCaller, 1 orgBaseCaller
        (-bcg), 2 refCaller, 3 nanoCaller (-bcNano), 4 refNanoCaller; significance: 0 significantNuc, 1 significantAnd90Nuc
        (-bc90), 2 significantAndSupport (-bc support). p_chisqr: C function pointer of the reference's p_chisqr (None: the
        closed form with the host libm). -> (t, s, q uint8 rows, stats structured array, kernel ms); template = 0: all
        templates concatenated, stats indexed by template id. out: (t, s, q, stats) buffers to fill instead of fresh arrays
        (pinned uint8 tensors / arrays of at least the row length; stats of DB_size * CONSENSUS_STATS.itemsize bytes): a
        pipeline that calls this every batch downloads at the link's speed instead of through pageable staging."""
        x0 = lib().kmagpu_chi2_threshold(float(evalue), p_chisqr)
        if x0 < 0:
            raise KmaGpuError(lib().kmagpu_last_error().decode())
        cp = ConsensusParams(int(bcd), int(caller), int(significance), 0, float(support), x0)
        info = self.info
        n = int(self.lengths[template]) if template else int(np.sum(self.lengths[1:], dtype=np.int64))
        nst = 1 if template else info.DB_size
        if out is None:
            t, s, q = (np.empty(n, dtype=np.uint8) for _ in range(3))
            st = np.zeros(nst, dtype=CONSENSUS_STATS)
        else:
            rows = [o.numpy() if hasattr(o, "numpy") else o for o in out]
            if any(len(r) < n for r in rows[:3]) or rows[3].nbytes < nst * CONSENSUS_STATS.itemsize:
                raise KmaGpuError("consensus: the out buffers are shorter than the rows")
            t, s, q = (r[:n] for r in rows[:3])
            st = rows[3].view(np.uint8)[:nst * CONSENSUS_STATS.itemsize].view(CONSENSUS_STATS)
        ms = C.c_float()
        _check(lib().kmagpu_consensus(self._h, int(template), C.byref(cp), t.ctypes.data, s.ctypes.data, q.ctypes.data, n,
                                      st.ctypes.data, C.byref(ms)))
        return t, s, q, st, ms.value

    def matrix_tensor(self):
        """the unsaturated device matrix as a torch int32 tensor (zero copy) for the NCCL all-reduce over ranks"""
        import torch
        ptr, n = C.c_void_p(), C.c_uint64()
        _check(lib().kmagpu_matrix_device(self._h, C.byref(ptr), C.byref(n)))

        class _View:
            __cuda_array_interface__ = {"shape": (int(n.value),), "typestr": "<i4", "data": (int(ptr.value), False), "version": 2}
        return torch.as_tensor(_View(), device=torch.device("cuda", self.device))

    def nw_batch(self, prob: np.ndarray, qpool: np.ndarray, params: Params | None = None):
        """NW_score / NW_band_score over independent problems; prob[n, 8] int32 =
        {template, t_s, t_e, q_off, q_s, q_e, k, band}. -> (out[n, 6], status[n], cells, steps, ms)"""
        p = params or default_params()
        prob = np.ascontiguousarray(prob, dtype=np.int32)
        qpool = np.ascontiguousarray(qpool, dtype=np.uint8)
        n = len(prob)
        out = np.zeros((n, 6), dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        cells, steps, ms = C.c_int64(), C.c_int64(), C.c_float()
        _check(lib().kmagpu_nw_batch(self._h, C.byref(p), n, prob.ctypes.data, qpool.ctypes.data, len(qpool), out.ctypes.data,
                                     status.ctypes.data, C.byref(cells), C.byref(steps), C.byref(ms)))
        return out, status, cells.value, steps.value, ms.value

    def lookup(self, kmers: np.ndarray) -> np.ndarray:
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        out = np.empty(len(kmers), dtype=np.int64)
        _check(lib().kmagpu_lookup_batch(self._h, kmers.ctypes.data, len(kmers), out.ctypes.data))
        return out


def record_offsets(stage: int, buf) -> np.ndarray:
    """byte offsets of the whole records at the head of a stage-1 / stage-2 stream, plus the end offset (n + 1 values).
    Host-only: works without a device."""
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    used = C.c_size_t()
    n = lib().kmagpu_record_walk(stage, buf.ctypes.data, len(buf), None, 0, C.byref(used))
    if n < 0:
        raise KmaGpuError(lib().kmagpu_last_error().decode(errors="replace"))
    off = np.empty(n + 1, dtype=np.uint64)
    lib().kmagpu_record_walk(stage, buf.ctypes.data, len(buf), off.ctypes.data, n, C.byref(used))
    off[n] = used.value
    return off


def stream_terminator(nreads: int) -> bytes:
    """kmers.c:257 -- int32 -(number of reads consumed) ends the stage-2 stream"""
    return np.array([-nreads], dtype=np.int32).tobytes()
