/*
 * kmagpu.h -- C ABI of the B200-native KMA mapping core (libkmagpu.so).
 *
 * Drop-in boundary: KMA 1.5.1 has no FFI; its "plugin API" is a set of global C
 * function pointers invoked once per read from T pthreads. A GPU needs batches, so the seams
 * this library binds to are the reference's three record streams. Every entry point below cites
 * the reference interface it replaces; INTEGRATION.md shows the patch a KMA maintainer applies.
 *
 * Conventions (mirroring pherror.h / the reference's ownership rules):
 *   - plain pointers and sizes only; the caller owns every buffer, the library never frees or
 *     reallocs caller memory;
 *   - every function returns 0 on success, non-zero on failure; kmagpu_last_error() returns a
 *     thread-local message (the host shim turns that into the reference's print + exit(errno));
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef KMAGPU_H
#define KMAGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kmagpu_db kmagpu_db;

/* POD copy of `Penalties` (penalties.h:22-33) + the CLI scalars the hot path reads. */
typedef struct kmagpu_params {
	int32_t M, MM, U, W1, Wl, Mn, PE; /* kma.c:327-336, 1308-1328 */
	int32_t d[25];                    /* rewards->d[t][q], row-major 5x5 */
	int32_t exhaustive;               /* -ex_mode (kma.c:551) */
	int32_t mq;                       /* -mq  minimum mapQ (align.c:658) */
	int32_t one2one;                  /* -1t1 */
	int32_t minlen;                   /* -ml  minimum alignment length (alnfrags.c:1156), default 16 */
	int32_t kmerscan;                 /* which kmerScan (savekmers.h:50): 0 = save_kmers (-1t1, savekmers.c:2442),
	                                     1 = save_kmers_chain (the default without -1t1, savekmers.c:5127) */
	int32_t matrix;                   /* kmagpu_trace_batch: add every accepted alignment to the base-count matrix as alnToMatPtr does
	                                     (assembly.c:1968): 0 = no, 1 = alnToMat (template nodes, assembly.c:1317), 2 = alnToMatDense (-dense, :1446) */
	int32_t apm;                      /* pairing of read pairs (-apm): 0 = p (save_kmers_penaltyPair savekmers.c:3572 / alnFragsPenaltyPE
	                                     alnfrags.c:1596), 1 = u, the reference's default (save_kmers_unionPair :3367 / alnFragsUnionPE :1220) */
	int32_t counters;                 /* alignment pass: collect the in-kernel statistic counters (kmagpu_align_stats mems, index_probes,
	                                     mem_bases, read_bytes, nw_*): measurement only, ~10 % of the pair kernel; kmagpu_default_params sets 1 */
	int32_t ts;                       /* -ts  seed trim of the traceback alignment (trimSeeds chain.c:496, called by KMA align.c:413); default 0 */
	int32_t lc;                       /* -lc (kma.c:694-700): save_kmers_chain selects its ankers by the length-corrected score (ankerScoreLen,
	                                     testExtensionScoreLen, proxiTestBestScoreLen, getBestAnkerScoreLen kmeranker.c:432, getTieAnkerScoreLen :496);
	                                     the ConClave side of -lc is the `lc` argument of kmagpu_conclave_* */
	double scoreT;                    /* -mrs (alnfrags.c:1168; also `mrs` of save_kmers_chain, kmers.c:51) */
	double minFrac;                   /* -mf  (updatescores.c:217-268) */
	double mrc;                       /* -mrc (alnfrags.h:38 mrcheck) */
	double coverT;                    /* -cov  maximum overlap of two regions of one read (save_kmers_chain), default 0.1 */
} kmagpu_params;

typedef struct kmagpu_db_info {
	int32_t DB_size;     /* templates + 1 (template ids are 1-based) */
	int32_t kmersize;    /* k of the .comp.b hash */
	int32_t kmerindex;   /* k of the per-template alignment index (.length.b[0]) */
	int32_t mega;        /* direct-addressed table (megaMap_getGlobal) */
	uint64_t size, n, v_index;
	uint64_t device_bytes; /* HBM held by this database */
	uint64_t seq_bases;    /* total template bases */
} kmagpu_db_info;

/* counters of the last seeding call -- the algorithmic-bytes inputs of SURVEY.md §8d */
typedef struct kmagpu_seed_stats {
	int64_t reads, mapped, read_words;
	int64_t lookups, hits, list_fetches, list_ids;
	int64_t overflow_reads;   /* reads that took the dense-scratch path */
	float ms_seed, ms_emit;   /* CUDA-event time of the scoring kernels / record writer */
	float ms_h2d, ms_total;   /* H2D copy; whole device-side step (first kernel start -> writer end) */
	int32_t launches;         /* kernels launched by the call */
	int32_t reserved;
} kmagpu_seed_stats;

void kmagpu_default_params(kmagpu_params *p);
const char *kmagpu_last_error(void);
int kmagpu_device_count(void);

/* Replaces hashMapKMA_load (hashmapkma.c:275) + the .length.b/.seq.b loads of runKMA
 * (runkma.c:161-220): reads <prefix>.comp.b/.length.b/.seq.b and makes them HBM resident on
 * `device`. */
int kmagpu_db_open(const char *prefix, int device, kmagpu_db **out);
void kmagpu_db_close(kmagpu_db *db);
/* A further handle on the SAME HBM image (hash table, template sequences, position index: read-only, shared, freed by
 * the last handle to close) with its own CUDA stream and batch buffers: what the reference's T worker threads share when
 * they all read one HashMapKMA / HashMapCCI (savekmers.c:94, alnfrags.c:2150). One handle per host thread; handles may
 * run concurrently. The base-count matrix (kmagpu_matrix_*), the run-wide score accumulators (kmagpu_scores_reset) and the NCCL
 * communicator belong to the image: the handles of one GPU add to ONE set of sums (atomics). */
int kmagpu_db_clone(kmagpu_db *db, kmagpu_db **out);
int kmagpu_db_get_info(const kmagpu_db *db, kmagpu_db_info *info);

/* Replaces the per-read loop of save_kmers_threaded (savekmers.c:94-271) with kmerScan =
 * save_kmers (-1t1, savekmers.c:2442): consumes `nbytes` of whole stage-1 records
 * (runinput.c:765-787 printFsa layout) and writes the stage-2 records (ankers.c:30-50
 * print_ankers layout) of the reads that map, in input order (= the reference's `-t 1` order).
 * The stream terminator (kmers.c:257) is NOT written; the caller appends -(sum of *nreads). */
int kmagpu_seed_batch(kmagpu_db *db, const kmagpu_params *p, const void *stage1, size_t nbytes,
                      void *stage2_out, size_t out_cap, size_t *out_bytes, int64_t *nreads,
                      kmagpu_seed_stats *stats);

/* The same call split in three so a caller (bench.py) can time the kernels with the batch
 * already resident in HBM: upload = parse + H2D, run = kernels only, download = D2H. */
int kmagpu_seed_upload(kmagpu_db *db, const void *stage1, size_t nbytes, int64_t *nreads);
int kmagpu_seed_run(kmagpu_db *db, const kmagpu_params *p, kmagpu_seed_stats *stats);
int kmagpu_seed_download(kmagpu_db *db, void *stage2_out, size_t out_cap, size_t *out_bytes);

/* ------------------------------------------------------------------ alignment pass (stage 3, first half) */

/* one row per (read, candidate template), in stream order: what KMA_score (align.c:509) returned for it */
typedef struct kmagpu_cand {
	int32_t read;      /* index of the stage-2 record in the batch */
	int32_t tmpl;      /* template id, signed as aligned (negative = reverse strand) */
	int32_t score, len, pos, match, tGaps, qGaps;   /* AlnScore (nw.h:37-44) */
} kmagpu_cand;

typedef struct kmagpu_align_stats {
	int64_t reads, tasks, frags;        /* records in, (read, template) pairs aligned, frag_raw records out */
	int64_t mems;                       /* maximal exact matches found */
	int64_t nw_full_calls, nw_band_calls;
	int64_t nw_full_cells, nw_band_cells;  /* DP cells as the reference counts them: t_len*q_len, t_len*(band+1) */
	int64_t nw_steps;                   /* warp wavefront steps (32 cell slots each) */
	int64_t overflow_tasks;             /* pairs re-run on the large-scratch path */
	int64_t index_probes;               /* hashMapCCI_get calls the reference's sequential seed scan makes */
	int64_t mem_bases;                  /* sum of MEM lengths (bases compared by seed extension) */
	int64_t read_bytes;                 /* packed words + 0-4 bytes + N list of the reads, once per pair */
	float ms_prep, ms_align, ms_reduce, ms_h2d, ms_total;
	int32_t launches, reserved;
} kmagpu_align_stats;

/* Replaces alnFrags_threaded (alnfrags.c:2150-2294) with alnFragsPE = alnFragsSE (alnfrags.c:1052-1218):
 * consumes `nbytes` of whole stage-2 records (ankers.c:30-50; a trailing terminator is ignored) and, per read,
 * runs anker_rc_comp / KMA_score against every candidate template, keeps the best hits (update_Scores,
 * updatescores.c:203-298), ADDS the ConClave sums into alignment_scores[DB_size] / uniq_alignment_scores[DB_size]
 * (runkma.c:98-99) and writes the frag_raw records (updatescores.c:284-295) in input order (= `-t 1` order).
 * cand_out (optional, cand_cap rows) receives the per-candidate AlnScore rows. */
int kmagpu_align_batch(kmagpu_db *db, const kmagpu_params *p, const void *stage2, size_t nbytes,
                       void *frag_out, size_t out_cap, size_t *out_bytes,
                       uint64_t *alignment_scores, uint64_t *uniq_alignment_scores,
                       kmagpu_cand *cand_out, size_t cand_cap, size_t *cand_rows, kmagpu_align_stats *stats);

/* The same call split so that a caller can keep the batch resident in HBM between the stages and time the kernels
 * alone: upload = record walk + H2D; from_seed = take the stage-2 stream kmagpu_seed_run left on the device. */
int kmagpu_align_upload(kmagpu_db *db, const void *stage2, size_t nbytes, int64_t *nreads);
int kmagpu_align_from_seed(kmagpu_db *db, int64_t *nreads);
int kmagpu_align_run(kmagpu_db *db, const kmagpu_params *p, int want_cand, kmagpu_align_stats *stats);
int kmagpu_align_download(kmagpu_db *db, void *frag_out, size_t out_cap, size_t *out_bytes,
                          uint64_t *alignment_scores, uint64_t *uniq_alignment_scores,
                          kmagpu_cand *cand_out, size_t cand_cap, size_t *cand_rows);

/* Replaces the alignment part of assemble_KMA's inner loop (assembly.c:1868-1961): per fragment record of the
 * per-template files ConClave writes (frags.c:45-48: int32[8]{template, q_len, nHits, score, start, end, hdrlen, flag}
 * + read bytes 0-4 + header; a trailing int32 -1 is ignored) runs anker_rc (align.c:780, when score == 0) and KMA
 * (align.c:214) with traceback, then the acceptance test. Output per record, input order:
 * int32[12]{accepted, read_score, start, end, score, len, pos, match, tGaps, qGaps, turned, ncol} + the aligned rows
 * t[ncol] s[ncol] q[ncol] (Aln, nw.h:46-56: bases 0-3, 4 = N, 5 = gap; '|' match, '_' otherwise) -- what alnToMat
 * (assembly.c:1317) and updateFrags consume next. `turned` = the read was reverse-complemented by anker_rc. */
int kmagpu_trace_batch(kmagpu_db *db, const kmagpu_params *p, const void *frags, size_t nbytes,
                       void *out, size_t out_cap, size_t *out_bytes, int64_t *nrecords, kmagpu_align_stats *stats);

/* -mem_mode: replaces the "Collecting k-mer scores" loop of runKMA_MEM (runkma.c:1088-1140) with update_Scores_MEM
 * (updatescores.c:26) / update_Scores_pe_MEM (:64): every stage-2 record (pairs included) becomes a frag_raw record
 * whose hits are its candidate templates over their whole length, scored with stage 2's k-mer score, written in input
 * order; the scores are ADDED into alignment_scores / uniq_alignment_scores [DB_size] (NULL = not wanted). */
int kmagpu_memscore_batch(kmagpu_db *db, const void *stage2, size_t nbytes, void *frag_out, size_t out_cap, size_t *out_bytes,
                          uint64_t *alignment_scores, uint64_t *uniq_alignment_scores, int64_t *nrecords);

/* kmagpu_memscore_batch on the stage-2 stream the last kmagpu_seed_run of this handle left in HBM. Either call keeps its
 * frag_raw stream resident for kmagpu_conclave_resident; frag_out = NULL skips the download. */
int kmagpu_memscore_from_seed(kmagpu_db *db, void *frag_out, size_t out_cap, size_t *out_bytes, uint64_t *alignment_scores,
                              uint64_t *uniq_alignment_scores, int64_t *nrecords);

/* Replaces runConClave (conclave.c:43-213, -ConClave 1) + printFrags (frags.c:30-61) for one chunk of frag_raw records
 * (the reference cuts a new chunk every maxFrag fragments, conclave.c:196-207): per record the template with the
 * largest GLOBAL alignment score wins (ties: score per template base, unique score, smaller id) -- so
 * alignment_scores / uniq_alignment_scores [DB_size] must be the sums over the whole run and over all GPUs -- reads
 * chosen on the reverse strand are reverse-complemented and their query bounds mirrored; w_scores[DB_size],
 * fragmentCounts[DB_size], readCounts[DB_size] are ADDED into (NULL = not wanted). frags_out receives the
 * per-template fragment stream of frags.c:45-48 (template order, reverse arrival order inside a template, int32 -1
 * at the end): the input of kmagpu_trace_batch. */
int kmagpu_conclave_batch(kmagpu_db *db, const void *frag_raw, size_t nbytes, const uint64_t *alignment_scores,
                          const uint64_t *uniq_alignment_scores, void *frags_out, size_t out_cap, size_t *out_bytes,
                          uint64_t *w_scores, uint32_t *fragmentCounts, uint32_t *readCounts, int64_t *nrecords);

/* Which ConClavePtr (conclave.c:30) kmagpu_conclave_batch / _resident / _from_align stand for: 0 = runConClave (the
 * default), 1 = runConClave_lc (conclave.c:215-384, bound by -lc: score per template base before the total). */
int kmagpu_conclave_mode(kmagpu_db *db, int length_corrected);

/* -ConClave 2 (runkma.c:591: ConClave2Ptr = runConClave2 / runConClave2_lc, conclave.c:386 / 749): the ConClave entry points
 * then make a provisional choice, drop the templates whose provisional sum is not significant (the chi-square test of
 * conclave.c:467-491 in the reference's long double arithmetic over the caller's p_chisqr, stdstat.c:136; or / and the
 * depth test `w >= scoreT * t_len` with cmp_or / cmp_and, kma.c:916), let reads with exactly one significant candidate add
 * to its unique score, and draw the final template with probability proportional to the unique scores (4-key order as the
 * fallback). The batch of such a call has to be the whole run. scoreT / evalue as runKMA passes them. version 1 = runConClave.
 * kmagpu_conclave_uniq_scores: the unique scores as the last ConClave call left them (runConClave2 updates them in place). */
int kmagpu_conclave_version(kmagpu_db *db, int version, double scoreT, double evalue, int and_mode, double (*p_chisqr)(long double));
int kmagpu_conclave_uniq_scores(kmagpu_db *db, uint64_t *uniq_alignment_scores);

/* kmagpu_conclave_batch on the frag_raw stream the last score collection (kmagpu_memscore_batch / _from_seed) of this
 * handle left in HBM; with kmagpu_trace_from_conclave the whole -mem_mode flow (stage 1 text -> stage 2 -> score
 * collection -> ConClave -> traceback + base counts -> consensus) runs without a record leaving the device. */
int kmagpu_conclave_resident(kmagpu_db *db, const uint64_t *alignment_scores, const uint64_t *uniq_alignment_scores, void *frags_out,
                             size_t out_cap, size_t *out_bytes, uint64_t *w_scores, uint32_t *fragmentCounts, uint32_t *readCounts,
                             int64_t *nrecords);

/* ... and on the frag_raw stream the last kmagpu_align_run of this handle left in HBM. (ConClave needs the score sums
 * of the WHOLE run: a multi-batch run keeps one handle per resident batch, or takes the host path for the others.) */
int kmagpu_conclave_from_align(kmagpu_db *db, const uint64_t *alignment_scores, const uint64_t *uniq_alignment_scores, void *frags_out,
                               size_t out_cap, size_t *out_bytes, uint64_t *w_scores, uint32_t *fragmentCounts, uint32_t *readCounts,
                               int64_t *nrecords);

/* kmagpu_trace_batch on the fragment stream the last kmagpu_conclave_batch / _resident of this handle left in HBM (that call may
 * pass frags_out = NULL when the host does not need the fragments). out = NULL in either trace call: no row output,
 * only the base counts (params->matrix) and the statistics -- what -dense / -matrix runs need. */
int kmagpu_trace_from_conclave(kmagpu_db *db, const kmagpu_params *params, void *out, size_t out_cap, size_t *out_bytes,
                               int64_t *nrecords, kmagpu_align_stats *stats);

/* The per-position base counts of the assembly pass (Assembly.counts[6] = {A, C, G, T, N, gap}, assembly.h:55-58) for
 * the template nodes of every template, HBM resident: uint32 [sum of template lengths][6], template t starting at
 * position sum(len[1..t-1]). kmagpu_trace_batch with params->matrix != 0 adds the accepted alignments of a batch
 * (alnToMat, assembly.c:1317-1444, restricted to the template nodes -- insertion nodes are order dependent and stay
 * on the host -- or alnToMatDense, assembly.c:1446-1497). Counts are kept unsaturated on the device so that ranks can
 * be summed (kmagpu_matrix_device gives the device pointer for the NCCL all-reduce); kmagpu_matrix_download clamps
 * to the reference's uint16 saturation (assembly.c:1436). tmpl = 0 downloads every template, else one template's
 * len * 6 entries; counts = NULL only reports the entry count. */
int kmagpu_matrix_reset(kmagpu_db *db);
int kmagpu_matrix_device(kmagpu_db *db, void **device_ptr, uint64_t *entries);
int kmagpu_matrix_download(kmagpu_db *db, int32_t tmpl, uint16_t *counts, size_t cap_entries, size_t *entries);

/* Replaces callConsensus (assembly.c:1499-1631; called from assemble_KMA, assembly.c:2065) over the template nodes of
 * the HBM-resident base-count matrix: per template position the template base, the called base and the match mark
 * (the t / q / s rows of `Assem`, assembly.h:34-53) and per template the sums callConsensus leaves in aligned_assem.
 * Insertion nodes (order dependent) stay with the host, which splices their calls between the rows. */
typedef struct kmagpu_consensus_params {
	int32_t bcd;          /* -bcd: minimum depth of an upper-case call (kma.c:318, default 1) */
	int32_t caller;       /* baseCall (assembly.c:46): 0 baseCaller (:162), 1 orgBaseCaller (-bcg, :181), 2 refCaller (:193),
	                         3 nanoCaller (-bcNano, :205), 4 refNanoCaller (:238) */
	int32_t significance; /* significantBase (assembly.c:45): 0 significantNuc (:141), 1 significantAnd90Nuc (-bc90, :145),
	                         2 significantAndSupport (-bc x, :149) */
	int32_t reserved;
	double support;       /* the x of -bc x */
	double chi2_min;      /* kmagpu_chi2_threshold(evalue, p_chisqr): a call is significant from this statistic on */
} kmagpu_consensus_params;

typedef struct kmagpu_consensus_stats {   /* what callConsensus adds to aligned_assem (assembly.c:1621-1627) */
	uint64_t depth, depthVar;
	uint32_t len, aln_len, cover, reserved;
} kmagpu_consensus_stats;

/* Host only: the smallest statistic x with p_chisqr(x) <= evalue, found by bisection over p_chisqr (stdstat.c:136; the
 * reference passes its own function, NULL uses the closed form of stdstat.c:146 with the host libm, which covers
 * evalue >= 1e-11). Negative on error. */
double kmagpu_chi2_threshold(double evalue, double (*p_chisqr)(long double));

/* tmpl != 0: rows of that template (len bytes each, cap = capacity of each row buffer), stats[1]. tmpl = 0: every
 * template, rows concatenated in template order, stats[DB_size] indexed by template id. Row pointers may be NULL. */
int kmagpu_consensus(kmagpu_db *db, int32_t tmpl, const kmagpu_consensus_params *cp, uint8_t *t, uint8_t *s, uint8_t *q,
                     size_t cap, kmagpu_consensus_stats *stats, float *ms);

/* Stage 1 on the device: replaces, per read, what run_input / run_input_PE (runinput.c:370-560) do between the record
 * splitter (FileBuffgetFq seqparse.c:241 / FileBuffgetFsa) and the stage-1 pipe: base translation through `trans`
 * (the reference passes its to2Bit, kma.c:1439-1482), phredStat's end trim (runinput.c:127-167; the default branch:
 * -mp only -- -eq, the hard mask and the QC report are not built) or fsastat's N trim (runinput.c:315-368), the -ml /
 * -xl filters, the pairing rule of run_input_PE (runinput.c:528-539), compDNA (compdna.c:99-127) and the records of
 * printFsa / printFsa_pair (runinput.c:765-825). */
typedef struct kmagpu_ingest_params {
	int32_t fastq;        /* 1: qualities present (phredStat), 0: FASTA (fsastat) */
	int32_t paired;       /* reads 2i and 2i + 1 are mates (run_input_PE / run_input_INT) */
	int32_t min_phred;    /* -mp (kma.c:293, default 20) */
	int32_t phred_scale;  /* what getPhredFileBuff (seqparse.c:551) found: 33 or 64 */
	int32_t minlen;       /* -ml (kma.c:309, default 16) */
	int32_t maxlen;       /* -xl (kma.c:310, default 2147483647) */
	int32_t min_q;        /* -eq (kma.c:617): phredStat's bidirectional trim until the mean error probability is below 10^(-eq/10)
	                         (runinput.c:196-296); also raises min_phred to it (runinput.c:380). 0 = off */
	int32_t hardmask_q;   /* -mi (kma.c:608): bases whose RAW quality byte is below it become N (runinput.c:183). 0 = off */
	uint8_t trans[256];   /* byte -> 0-3 base, 4 N, 8 other, 16 newline */
	double prob[256];     /* prob[q] = 10^(-q/10), the table the CLI hands to run_input (kma.c:219); read only when min_q or
	                         hardmask_q is set */
} kmagpu_ingest_params;

/* Host only: the line structure of a chunk of 4-line FASTQ (fastq != 0) or 2-line FASTA text. fields[i][5] = {header
 * offset (past '@' / '>'), header length without trailing white space, sequence offset, sequence length without
 * trailing bytes `trans` maps to 8, quality offset}. Returns the number of whole records, *used = the bytes they span
 * (fields = NULL only counts); -1 on malformed input. */
int64_t kmagpu_fastx_split(const void *text, size_t nbytes, int fastq, const uint8_t *trans, uint32_t *fields, size_t cap, size_t *used);

/* Host only: multi-line FASTA -> the 2-line form kmagpu_fastx_split / kmagpu_stage1_text take. What FileBuffgetFsa
 * (seqparse.c:66-160) keeps of a record: the header line, and between it and the next '>' every byte `trans` maps below 8
 * (line ends, '\r' and blanks drop out wherever they stand) as ONE sequence line. Whole records only: *used = the input
 * bytes consumed (without eof the last record stays, its end is not known yet). Returns the bytes written (out_cap >=
 * nbytes + 1), -1 on error. */
int64_t kmagpu_fasta_unwrap(const void *text, size_t nbytes, const uint8_t *trans, int eof, void *out, size_t out_cap, size_t *used);

/* Host only: the start of the first record at or after byte `from` (FASTQ: a line starting with '@' whose second-next
 * line starts with '+'), so that several host threads can run kmagpu_fastx_split on byte ranges of one chunk.
 * Returns nbytes when there is none. */
size_t kmagpu_fastx_sync(const void *text, size_t nbytes, int fastq, size_t from);

/* text (host) -> stage-1 records in input order. The stream always stays in HBM as the input of the next
 * kmagpu_seed_run (as if kmagpu_seed_upload had been called with it); stage1_out != NULL also downloads it.
 * count = what run_input returns (one per printed read or pair), ms = the three kernels and their scans. Paired
 * input: interleave the two files' fields (mates at 2i, 2i + 1); text2 (may be NULL) is the second file's chunk, its
 * fields count their offsets from text_bytes on (the two chunks sit back to back in one device buffer). */
int kmagpu_stage1_batch(kmagpu_db *db, const kmagpu_ingest_params *ip, const void *text, size_t text_bytes, const void *text2,
                        size_t text2_bytes, const uint32_t *fields, size_t nreads, void *stage1_out, size_t cap, size_t *out_bytes,
                        int64_t *count, float *ms);

/* The same with the record splitter on the device too: the chunk(s) of file text go to HBM as they are, newline
 * positions are counted / scanned / scattered by 64-byte blocks and a thread per record builds the field row
 * kmagpu_fastx_split would have produced. text2 != NULL with ip->paired: the chunk of the second file; records are
 * paired by index, min(records1, records2) pairs are taken. eof != 0: a last line without its newline counts.
 * used1 / used2 = the bytes the taken records span (the caller carries the rest over to its next chunk). */
int kmagpu_stage1_text(kmagpu_db *db, const kmagpu_ingest_params *ip, const void *text1, size_t bytes1, const void *text2, size_t bytes2,
                       int eof, size_t *used1, size_t *used2, void *stage1_out, size_t cap, size_t *out_bytes, int64_t *count, float *ms);

/* NW_score (nw.c:642) / NW_band_score (nw.c:892) over a batch of independent problems, one warp each.
 * prob[i] = {template id, t_s, t_e, q_off, q_s, q_e, k, band (0 = full matrix)}; the query bytes (0-4) of problem i
 * start at qpool + q_off. out[i] = {score, len, pos, match, tGaps, qGaps}; status[i] != 0: not computed
 * (1 scratch too small, 2 band narrower than the length difference). cells = sum of DP cells, ms = kernel time. */
int kmagpu_nw_batch(kmagpu_db *db, const kmagpu_params *p, size_t n, const int32_t *prob, const uint8_t *qpool,
                    size_t qbytes, int32_t *out, int32_t *status, int64_t *cells, int64_t *steps, float *ms);

/* Host-only helper (no device needed): walk the whole records at the head of a stage-1 (stage = 1, 16-byte headers,
 * runinput.c:765-787 / loadFsa savekmers.c:50-92), stage-2 (stage = 2, 28-byte headers, ankers.c:163-220) or assembly
 * fragment (stage = 3, 32-byte headers, frags.c:45-48) or frag_raw (stage = 4, 20-byte headers, updatescores.c:284-295;
 * a record with a negative score includes the mate block that follows it) stream.
 * Returns the number of whole records (stopping at a terminator or a partial record), stores their byte offsets in
 * offsets[0..min(count, cap)) when offsets != NULL and the bytes they span in *used. -1 on a corrupt header. */
int64_t kmagpu_record_walk(int stage, const void *buf, size_t nbytes, uint64_t *offsets, size_t cap, size_t *used);

/* ------------------------------------------------------------------ multi-GPU exchange
 * One process per GPU, reads sharded by rank, database replicated. The only exchanges of the path are sums over ranks,
 * done with ncclAllReduce INSIDE the library, in place in HBM, on the handle's stream (NCCL is bound at run time with
 * dlopen; a single-GPU host never loads it):
 *   - alignment_scores / uniq_alignment_scores [DB_size] (runkma.c:98-99), which update_Scores adds to per read
 *     (updatescores.c:228/276) and ConClave's choice pass reads as GLOBAL sums (conclave.c:80-123);
 *   - the base-count matrix of the assembly pass (assembly.c:1436: +1 per aligned base; unsaturated uint32 sums commute).
 * kmagpu_comm_unique_id: rank 0 makes the 128-byte NCCL id, the host hands it to the other ranks (file, MPI, TCP store ...);
 * kmagpu_comm_init: every rank, once per handle (world = 1: no communicator, the reductions below are no-ops).
 * kmagpu_scores_reset: start the run-wide device accumulators of this handle; every kmagpu_align_run / kmagpu_memscore_*
 *   then also adds its batch sums to them. kmagpu_allreduce_scores sums them over ranks (ms = device time of the
 *   all-reduce) and optionally copies them out; kmagpu_conclave_* with alignment_scores = uniq_alignment_scores = NULL
 *   reads them on the device -- the score arrays never leave HBM between the alignment pass and ConClave.
 * kmagpu_allreduce_matrix: the handle's base-count matrix, in place. kmagpu_allreduce_u64: any host array of counters
 *   (w_scores ...), the generic export of SURVEY 8b(5). */
int kmagpu_comm_unique_id(void *id128, size_t cap);
int kmagpu_comm_init(kmagpu_db *db, const void *id128, int rank, int world);
void kmagpu_comm_destroy(kmagpu_db *db);
int kmagpu_scores_reset(kmagpu_db *db);
/* Soft proximity (-proxi < 0 together with -mem_mode: the only case in which stage 2 is handed a negative minFrac, kma.c:1605):
 * every template a get*Proxi* function keeps adds its score to softProxi[] (savekmers.c:330, 1577, 1636, 1803, 1868;
 * kmeranker.c:357), save_kmers_batch appends the sums to the stage-2 stream (6 ints = their first 24 bytes, then DB_size
 * unsigned longs, kmers.c:151-153) and runKMA_MEM takes them for alignment_scores (runkma.c:1153). kmagpu_softproxi_reset starts
 * the sums of this handle's database image; every kmagpu_seed_run with params.minFrac < 0 then adds its batch;
 * kmagpu_softproxi_download copies them out ([DB_size]; sum them over ranks with kmagpu_allreduce_u64). */
int kmagpu_softproxi_reset(kmagpu_db *db);
int kmagpu_softproxi_download(kmagpu_db *db, uint64_t *sums);
int kmagpu_allreduce_scores(kmagpu_db *db, uint64_t *alignment_scores, uint64_t *uniq_alignment_scores, float *ms);
int kmagpu_allreduce_matrix(kmagpu_db *db, float *ms);
int kmagpu_allreduce_u64(kmagpu_db *db, uint64_t *buf, size_t n);

/* hashMap_get (hashmapkma.h:58; hashMap_getGlobal hashmapkma.c:149 / megaMap_getGlobal :264) over
 * a batch of k-mers: out[i] = offset of the template list inside values[], or -1. Test hook. */
int kmagpu_lookup_batch(kmagpu_db *db, const uint64_t *kmers, size_t n, int64_t *out);

#ifdef __cplusplus
}
#endif
#endif
