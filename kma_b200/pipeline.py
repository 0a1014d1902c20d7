"""Host-side driver of the mapping core for one GPU: what the reference's stage drivers (save_kmers_batch kmers.c:51,
runKMA's alignment loop runkma.c:291-445) do with T pthreads over a shared FILE*, done here with a few host threads
that each own one TemplateDB handle (own CUDA stream and batch buffers; the handles are clones of one another and share
ONE read-only database image in HBM, kmagpu_db_clone) and take the chunks of a record stream in turn. While one chunk's frag_raw stream is on its way
back over PCIe, the next chunk's records are being walked / uploaded and a third is in the kernels -- the copies hide
behind the compute instead of adding to it. Output order is chunk order = input order."""
from __future__ import annotations

import os
import threading

import numpy as np

from . import api


def phred_scale(text, chunk: int = 1048576) -> int:
    """getPhredFileBuff (seqparse.c:551-589) on the first buffer of a FASTQ file (CHUNK = 1 MiB, filebuff.h:36): 33 as soon as a
    quality byte lies in 54..58, 64 if one above 94 was seen, 0 for a byte below 33; reads longer than 301 count as 33"""
    buf = memoryview(text)[:chunk].tobytes()
    scale, maxlen, pos, n = 33, 0, 0, len(buf)
    while pos < n:
        for _ in range(3):                       # the three lines before the quality line (the scan starts behind its first byte)
            nl = buf.find(b"\n", pos + 1)
            if nl < 0:
                return scale if maxlen <= 301 else 33
            pos = nl
        end = buf.find(b"\n", pos + 1)
        line = buf[pos + 1:end if end >= 0 else n]
        for c in line:                           # in file order: the first decisive byte wins (a '\r' counts as below 33)
            if c < 33:
                return 0
            if 53 < c < 59:
                return 33
            if c > 94:
                scale = 64
        maxlen = max(maxlen, len(line))
        if end < 0:
            break
        pos = end
    return scale if maxlen <= 301 else 33


def read_reads(path: str):
    """Host I/O in front of stage 1, what openAndDetermine + the FileBuff readers do (filebuff.c, seqparse.c): the file as plain
    text -- gzip members inflated with zlib like the reference's BuffgzFileBuff (filebuff.c:29-74) --, FASTA unwrapped to
    one sequence line per record (kmagpu_fasta_unwrap = FileBuffgetFsa's view of it). -> (text bytes, fastq?, phred scale).
    The decompression is sequential per file (that is DEFLATE); pair files can be read by two threads."""
    import zlib
    raw = open(path, "rb").read()
    if raw[:2] == b"\x1f\x8b":
        parts, data = [], raw
        while data:                              # concatenated members (bgzip, cat a.gz b.gz)
            d = zlib.decompressobj(31)
            parts.append(d.decompress(data))
            data = d.unused_data
        raw = b"".join(parts)
    if not raw:
        return raw, True, 33
    if raw[:1] == b">":
        flat, used = api.fasta_unwrap(raw)
        return flat, False, 33
    if raw[:1] != b"@":
        raise api.KmaGpuError(f"{path}: neither FASTQ nor FASTA")
    return raw, True, phred_scale(raw)


class MapPipeline:
    def __init__(self, prefix: str, device: int = 0, workers: int = 2, params=None):
        first = api.TemplateDB(prefix, device)
        self.dbs = [first] + [first.clone() for _ in range(workers - 1)]   # one image per GPU however many workers
        self.params = params or api.default_params()
        self.info = self.dbs[0].info

    def close(self):
        for d in self.dbs:
            d.close()

    def chunk_bounds(self, stage1, n_chunks: int):
        """byte ranges of n_chunks contiguous groups of whole records; pairs are never split"""
        off = api.record_offsets(1, stage1)
        n = len(off) - 1
        cuts = [0]
        for c in range(1, n_chunks):
            i = (n * c) // n_chunks
            # a second mate (positive header length right after a negative one) must stay with its first mate
            if 0 < i < n and int(np.frombuffer(stage1[int(off[i - 1]) + 12:int(off[i - 1]) + 16].tobytes(), dtype=np.int32)[0]) < 0:
                i += 1
            cuts.append(int(off[min(i, n)]))
        cuts.append(int(off[n]))
        return [(cuts[i], cuts[i + 1]) for i in range(n_chunks) if cuts[i + 1] > cuts[i]]

    def map_text(self, text1, fields1, text2, fields2, n_chunks, outs, scores, **ingest):
        """The same from FASTQ text (pinned uint8 arrays / tensors): stage 1 runs on the device too (run_input_batch), so a
        chunk's reads go text -> HBM -> stage-1 records -> stage 2 -> alignment pass without coming back. fields1 / fields2:
        api.fastx_split of the two files (fields2 = None: single end). Chunks are ranges of reads (pairs)."""
        n = len(fields1)
        cuts = [(n * c) // n_chunks for c in range(n_chunks + 1)]
        chunks = [(cuts[i], cuts[i + 1]) for i in range(n_chunks) if cuts[i + 1] > cuts[i]]
        res = [None] * len(chunks)
        part = [(np.zeros_like(scores[0]), np.zeros_like(scores[1])) for _ in self.dbs]
        err = []
        end1 = int(text1.numel() if hasattr(text1, "numel") else text1.size)
        end2 = 0 if text2 is None else int(text2.numel() if hasattr(text2, "numel") else text2.size)

        def span(fields, a, b, end):   # bytes of reads a..b-1: from the '@' of a to the '@' of b
            return int(fields[a][0]) - 1, (int(fields[b][0]) - 1 if b < len(fields) else end)

        def work(w):
            db = self.dbs[w]
            try:
                for i in range(w, len(chunks), len(self.dbs)):
                    a, b = chunks[i]
                    lo1, hi1 = span(fields1, a, b, end1)
                    f = fields1[a:b].copy()
                    f[:, [0, 2, 4]] -= np.uint32(lo1)
                    t2 = None
                    if fields2 is not None:
                        lo2, hi2 = span(fields2, a, b, end2)
                        g = fields2[a:b].copy()
                        g[:, [0, 2, 4]] = (g[:, [0, 2, 4]].astype(np.int64) + (hi1 - lo1 - lo2)).astype(np.uint32)
                        f = np.stack([f, g], axis=1).reshape(-1, 5)
                        t2 = text2[lo2:hi2]
                    _, cnt, _ = db.run_input_batch(text1[lo1:hi1], f, paired=fields2 is not None, download=False, text2=t2, **ingest)
                    db.seed_run(self.params)
                    db.align_from_seed()
                    db.align_run(self.params)
                    frag, _, _, _ = db.align_download(out=outs[i], scores=part[w])
                    res[i] = (frag, cnt)
            except Exception as e:
                err.append(e)

        ts = [threading.Thread(target=work, args=(w,)) for w in range(len(self.dbs))]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if err:
            raise err[0]
        for a, u in part:
            scores[0][:] += a
            scores[1][:] += u
        return res

    def map_text_device_split(self, text1, text2, n_chunks, outs, scores, fastq=True, **ingest):
        """map_text with the record splitter on the device as well (run_input_text): the host only looks for n_chunks
        record starts at even byte fractions of each file (kmagpu_fastx_sync). That pairs the right mates only when the
        two files' records line up chunk by chunk (equal-length reads and names, the usual shape of an Illumina pair of
        files); every chunk checks that both of its texts were consumed to the last byte, and the caller falls back to
        map_text (host splitter, pairs by index) when one was not. Returns None in that case."""
        L = api.lib()
        def cuts(t):
            if t is None:
                return None
            a = t.numpy() if hasattr(t, "numpy") else t
            nb = len(a)
            return sorted({L.kmagpu_fastx_sync(a.ctypes.data, nb, int(fastq), (nb * i) // n_chunks) for i in range(n_chunks)} | {nb})
        c1, c2 = cuts(text1), cuts(text2)
        if c2 is not None and len(c2) != len(c1):
            return None
        nch = len(c1) - 1
        res = [None] * nch
        part = [(np.zeros_like(scores[0]), np.zeros_like(scores[1])) for _ in self.dbs]
        err, bad = [], []

        def work(w):
            db = self.dbs[w]
            try:
                for i in range(w, nch, len(self.dbs)):
                    t2 = None if c2 is None else text2[c2[i]:c2[i + 1]]
                    _, cnt, _, u1, u2 = db.run_input_text(text1[c1[i]:c1[i + 1]], text2=t2, fastq=fastq, download=False, **ingest)
                    if u1 != c1[i + 1] - c1[i] or (c2 is not None and u2 != c2[i + 1] - c2[i]):
                        bad.append(i)
                        return
                    db.seed_run(self.params)
                    db.align_from_seed()
                    db.align_run(self.params)
                    frag, _, _, _ = db.align_download(out=outs[i], scores=part[w])
                    res[i] = (frag, cnt)
            except Exception as e:
                err.append(e)

        ts = [threading.Thread(target=work, args=(w,)) for w in range(len(self.dbs))]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if err:
            raise err[0]
        if bad:
            return None
        for a, u in part:
            scores[0][:] += a
            scores[1][:] += u
        return res

    def map(self, stage1, bounds, outs, scores):
        """stage 2 + alignment pass over the chunks `bounds` of the stage-1 stream (a pinned uint8 tensor / array).
        outs[i]: buffer for chunk i's frag_raw bytes; scores: (alignment_scores, uniq_alignment_scores) uint64 arrays
        that receive the sums. Returns per-chunk (frag bytes, groups read)."""
        res = [None] * len(bounds)
        part = [(np.zeros_like(scores[0]), np.zeros_like(scores[1])) for _ in self.dbs]
        err = []

        def work(w):
            db = self.dbs[w]
            try:
                for i in range(w, len(bounds), len(self.dbs)):
                    lo, hi = bounds[i]
                    n = db.seed_upload(stage1[lo:hi])
                    db.seed_run(self.params)
                    db.align_from_seed()
                    db.align_run(self.params)
                    frag, _, _, _ = db.align_download(out=outs[i], scores=part[w])
                    res[i] = (frag, n)
            except Exception as e:   # surfaced to the caller below
                err.append(e)

        ts = [threading.Thread(target=work, args=(w,)) for w in range(len(self.dbs))]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if err:
            raise err[0]
        for a, u in part:
            scores[0][:] += a
            scores[1][:] += u
        return res

    def map_to_consensus(self, text1, text2, frag_outs, params, fastq=True, consensus_args=None, trace=None, cons_out=None, slice_weights=None, **ingest):
        """The whole mapping core on one batch of FASTQ text (pinned uint8 tensors; text2 = the second file of a pair of
        files or None), every stream resident in HBM, what `kma -i / -ipe ... -o out` computes between its input files and
        its writers: record splitter + stage 1 -> stage 2 -> alignment pass | ConClave sums all-reduced over ranks (NCCL
        inside the library, in place) | ConClave choice -> traceback alignment + base counts | matrix all-reduced over
        ranks | consensus of every template. Each worker handle (clones of one database image: one set of sums) takes one
        slice of the batch, so one slice's PCIe copies hide behind another's kernels; the two exchanges are the only
        points where the workers meet. Down come: per worker its per-template fragment stream into frag_outs[w] (what the
        .frag.gz writer needs), then the consensus rows and per-template sums (what .res / .fsa / .aln are written from).
        Paired files must line up slice by slice (equal-length records), as map_text_device_split requires.
        slice_weights: relative sizes of the workers' slices (default equal). A small first slice lets the kernels start
        while the rest of the text is still crossing PCIe.
        cons_out: (t, s, q, stats) pinned buffers for the consensus rows (api.TemplateDB.consensus's out).
        trace: optional list; (worker, phase, perf_counter seconds) is appended as each worker leaves a phase.
        Returns dict(reads, fragments, consensus=(t, s, q, stats), totals=(w_scores, fragmentCounts, readCounts), frag_bytes)."""
        import time
        import torch  # noqa: F401  (pinned tensors come from the caller)

        def mark(w, what):
            if trace is not None:
                trace.append((w, what, time.perf_counter()))
        L = api.lib()
        W = len(self.dbs)

        wts = list(slice_weights) if slice_weights else [1.0] * W
        if len(wts) != W or min(wts) <= 0:
            raise api.KmaGpuError("slice_weights: one positive weight per worker")
        fracs = [sum(wts[:i]) / sum(wts) for i in range(W)]

        def cuts(t):
            a = t.numpy() if hasattr(t, "numpy") else t
            nb = len(a)
            return sorted({L.kmagpu_fastx_sync(a.ctypes.data, nb, int(fastq), int(nb * f)) for f in fracs} | {nb})
        c1 = cuts(text1)
        c2 = cuts(text2) if text2 is not None else None
        if c2 is not None and len(c2) != len(c1):
            raise api.KmaGpuError("the two files do not split into the same number of slices")
        nsl = len(c1) - 1
        DB = self.info.DB_size
        totals = [(np.zeros(DB, np.uint64), np.zeros(DB, np.uint32), np.zeros(DB, np.uint32)) for _ in range(W)]
        res = {"reads": [0] * W, "fragments": [0] * W, "frag": [None] * W}
        err = []
        lead = self.dbs[0]
        p_trace = api.Params.from_buffer_copy(bytes(params))
        p_trace.matrix = 1
        lead.scores_reset()
        lead.matrix_reset()

        def exchange_scores():
            try:
                mark(-1, "all workers at the score exchange")
                lead.allreduce_scores(download=False)
                mark(-1, "scores all-reduced")
            except Exception as e:
                err.append(e)

        def exchange_matrix():
            try:
                mark(-1, "all workers at the matrix exchange")
                lead.allreduce_matrix()
                mark(-1, "matrix all-reduced")
            except Exception as e:
                err.append(e)
        b1 = threading.Barrier(W, action=exchange_scores)
        b2 = threading.Barrier(W, action=exchange_matrix)
        # the slices go up one after the other, in order: four uploads issued at once share the link, and the first slice -- the
        # one the kernels wait for -- arrived after 17 ms instead of the 8 ms it takes alone (profiles/r02_e2e_phases_v5.log)
        up_done = [threading.Event() for _ in range(W)]
        in_order = os.environ.get("KMA_B200_E2E_UPLOAD_ORDER", "1") != "0"

        def work(w):
            db = self.dbs[w]
            try:
                if w < nsl:
                    t2 = None if c2 is None else text2[c2[w]:c2[w + 1]]
                    if in_order and w:
                        up_done[w - 1].wait()
                    try:
                        _, cnt, _, u1, u2 = db.run_input_text(text1[c1[w]:c1[w + 1]], text2=t2, fastq=fastq, download=False, **ingest)
                    finally:
                        up_done[w].set()
                    if u1 != c1[w + 1] - c1[w] or (c2 is not None and u2 != c2[w + 1] - c2[w]):
                        raise api.KmaGpuError("the two files' records do not line up slice by slice")
                    res["reads"][w] = cnt
                    mark(w, "text up + stage 1")
                    db.seed_run(params)
                    mark(w, "stage 2")
                    db.align_from_seed()
                    db.align_run(params)
                    mark(w, "alignment pass")
            except Exception as e:
                err.append(e)
            up_done[w].set()   # a worker without a slice (or one that failed earlier) must not hold up the next one
            b1.wait()
            try:
                if w < nsl and not err:
                    frag, _, _, _, _ = db.conclave_from_align(None, None, out=frag_outs[w], totals=totals[w])
                    res["frag"][w] = frag
                    mark(w, "ConClave + fragments down")
                    _, nfr, _ = db.trace_from_conclave(p_trace, download=False)
                    res["fragments"][w] = nfr
                    mark(w, "traceback + base counts")
            except Exception as e:
                err.append(e)
            b2.wait()

        ts = [threading.Thread(target=work, args=(w,)) for w in range(W)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if err:
            raise err[0]
        mark(-1, "workers joined")
        wsc = sum(t[0] for t in totals)
        lead.allreduce_u64(wsc)
        mark(-1, "w_scores all-reduced")
        cons = lead.consensus(0, out=cons_out, **(consensus_args or {}))
        mark(-1, "consensus down")
        return {"reads": sum(res["reads"]), "fragments": sum(res["fragments"]), "consensus": cons[:4],
                "totals": (wsc, sum(t[1] for t in totals), sum(t[2] for t in totals)),
                "frag_bytes": sum(int(f.numel() if hasattr(f, "numel") else len(f)) for f in res["frag"] if f is not None)}
