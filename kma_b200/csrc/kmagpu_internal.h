// Internal declarations shared by the translation units of libkmagpu.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>
#include <vector>
#include "../../include/kmagpu.h"
#include "kmagpu_tindex.cuh"

void kmagpu_set_error(const char *fmt, ...);

#define KG_CUDA(call)                                                                              \
	do {                                                                                           \
		cudaError_t e__ = (call);                                                                  \
		if (e__ != cudaSuccess) {                                                                  \
			kmagpu_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
			return -1;                                                                             \
		}                                                                                          \
	} while (0)

// Device view of the template k-mer hash (.comp.b re-laid for one-sector lookups).
struct KgHashView {
	const uint32_t *exist;   // mega (direct-addressed) tables only: [size] value offset, 1 = null
	const uint4 *bk;         // hashed tables: [size] bucket -> {key0, value0, pos, cnt}: the first {key, value offset} of the bucket's
	                         // run in kv inline, where the run starts and how many entries a failing scan examines (0: empty).
	                         // One 16-byte load answers every miss in an empty or single-key bucket and every hit on a first key
	const uint2 *kv;         // [n + 1]  {key, value offset}: key_index and value_index fused; read only behind a first-key mismatch
	const uint16_t *values_s;  // template lists, u16 when DB_size < 65535 ...
	const uint32_t *values_w;  // ... else u32
	uint64_t hmask;          // size - 1
	uint32_t null_index;
	uint32_t n;
	int32_t kmersize;
	int32_t DB_size;
	int32_t mega;
};

// growable device / pinned buffers
struct KgBuf {
	void *p = nullptr;
	size_t cap = 0;
	bool pinned = false;
	int reserve(size_t bytes);
	void release();
};

struct SeedBatch {
	KgBuf d_in, d_off, d_res, d_pool, d_recoff, d_out, d_ctr, d_partial, d_dense;
	KgBuf d_soft;   // soft proximity sums of the batch at hand ([DB_size] u64; joins KgImageRef::d_soft when the batch is through)
	KgBuf d_kinds, d_mates, d_pool2;   // paired end: record kinds (0 single, 1/2 mates), per-mate strand lists
	KgBuf h_kinds;                     // pinned staging of the record kinds
	KgBuf d_chain, d_regpool, d_regs, d_rsize, d_partial2;   // chain mode: per-warp scratch, regions (as found / in read order), sizes
	size_t chain_stride = 0, reg_cap = 0;
	int chain_grid = 0;
	int32_t max_seqlen = 0;            // longest read of the batch
	int64_t out_nrec = 0;              // records in the device output stream (chain mode: regions, else reads)
	const uint32_t *out_recoff = nullptr;   // their offsets (device)
	int64_t npairs = 0;
	size_t pool2_cap = 0;
	KgBuf h_off;                 // pinned staging of record offsets
	KgBuf h_in, h_out;           // pinned staging of the streams
	int64_t nreads = 0;
	size_t in_bytes = 0;
	size_t out_bytes = 0;
	size_t pool_cap = 0;         // ints
	bool ran = false;
};

// working buffers of stage 1 (kmagpu_stage1.cu): kept across calls, a pipeline calls it once per chunk
struct Stage1Batch {
	KgBuf d_text, d_fields, d_win, d_u32, d_kind, d_partial, d_ctr, h_ctr;
	KgBuf d_cnt1, d_cnt2, d_lines1, d_lines2;   // device record splitter: newline counts + offsets per 64-byte block, line ends
	KgBuf d_prob, h_prob;                       // -eq / -mi: prob[] shifted by the phred scale + 10^(-eq/10) (device, pinned staging)
};

// working buffers of the traceback pass (kmagpu_trace_batch / kmagpu_trace_from_conclave), kept across calls
struct TraceBatch {
	KgBuf d_in, d_off, d_recs, d_sz, d_partial, d_ctr, d_slab, d_rows, d_outs, d_ovf, d_out;
};

// the per-template fragment stream the last kmagpu_conclave_batch wrote, resident for kmagpu_trace_from_conclave
struct FragBatch {
	KgBuf d_out, d_sz;
	KgBuf d_sc, d_items, d_keys, d_vals, d_partial, d_ctr, d_acc, d_tmp;   // ConClave's working buffers, kept across calls
	const uint32_t *off = nullptr;   // record offsets (inside d_sz)
	int64_t n = 0;
	size_t bytes = 0;
	bool valid = false;
};

// the frag_raw stream of the last -mem_mode score collection, resident for kmagpu_conclave_resident: one slot per
// stage-2 record (empty slots: mates consumed with their first read, reads shorter than k)
struct RawBatch {
	KgBuf d_in, d_off, d_recs, d_sz, d_partial, d_ctr, d_acc, d_out;
	const uint32_t *off = nullptr;   // slot offsets (inside d_sz), n + 1 entries
	int64_t n = 0;
	size_t bytes = 0;
	bool valid = false;
};

// one batch of the alignment pass (kmagpu_align.cu)
struct AlignBatch {
	KgBuf d_in, d_off, d_reads, d_slab, d_sz, d_partial, d_taskread, d_cand, d_recsize, d_out, d_ctr, d_scores,
	      d_scratch, d_ovf, d_res;
	KgBuf d_sorttmp;
	KgBuf d_probs, d_order;        // NW problem queue of the phase-split alignment pass + its per-class order lists
	size_t prob_cap = 0;           // queue capacity (grows to what a batch asked for, kept across calls)
	KgBuf h_off;
	const uint8_t *in = nullptr;   // device pointer of the stage-2 stream (d_in or the seeding output)
	int64_t nreads = 0, ntasks = 0;
	size_t in_bytes = 0, out_bytes = 0;
	bool ran = false, want_cand = false;
	std::vector<int32_t> h_cand;
	std::vector<uint64_t> h_scores;
};

// The read-only database image in HBM is shared by every handle made from it with kmagpu_db_clone: the handles differ
// in their stream, events and batch buffers only. The last handle to close frees the image.
struct KgImageRef {
	int refs = 1;
	// state every handle of the image adds to with atomics, so that worker handles of one GPU feed ONE set of sums:
	// the base-count matrix of the assembly pass, the run-wide ConClave accumulators, and the GPU's NCCL communicator
	unsigned int *d_mat = nullptr;     // uint32 [sum of template lengths][6], template t at d_mat_off[t] positions
	int64_t *d_mat_off = nullptr;
	size_t mat_entries = 0;
	unsigned long long *d_run_scores = nullptr;   // [alignment_scores[DB_size], uniq_alignment_scores[DB_size]]
	unsigned long long *d_soft = nullptr;         // soft proximity sums of the run (kmers.c:133-153), [DB_size]; NULL until kmagpu_softproxi_reset
	void *comm = nullptr;
	int comm_rank = 0, comm_world = 1;
};

struct kmagpu_db {
	int device = 0;
	KgImageRef *image = nullptr;
	kmagpu_db_info info{};
	KgHashView hv{};
	void *d_exist = nullptr, *d_kv = nullptr, *d_values = nullptr;
	// alignment side (.length.b / .seq.b)
	std::vector<int32_t> lengths;      // [DB_size], lengths[0] = kmerindex
	std::vector<int64_t> seq_off;      // [DB_size] word offset of template t
	uint64_t *d_seq = nullptr;         // all of .seq.b
	int32_t *d_lengths = nullptr;
	int64_t *d_seq_off = nullptr;
	size_t seq_words = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev[8]{};
	int sm_count = 148;
	SeedBatch seed;
	Stage1Batch s1;
	TraceBatch trc;
	FragBatch frg;
	RawBatch raw;
	int conclave_lc = 0;   // 1: runConClave_lc
	struct { int version = 1; double scoreT = 0.5, evalue = 0.05; bool and_mode = false; double (*p_chisqr)(long double) = nullptr; } cc2;   // -ConClave 2
	// per-template alignment index (kmagpu_tindex.cu)
	void *d_tmeta = nullptr, *d_tslots = nullptr;
	int32_t *d_tdups = nullptr;
	KgTIndexView tix{};
	AlignBatch aln;
	KgBuf d_cons_rows, d_cons_stat;   // consensus rows / per-template sums of the last kmagpu_consensus call (kept: no allocation per call)
};
int kg_scores_accumulate(kmagpu_db *db, const unsigned long long *batch_scores);
int kg_softproxi_accumulate(kmagpu_db *db, const unsigned long long *batch_sums);

int kg_tindex_build(kmagpu_db *db);
int kg_check_record(const uint8_t *rec, int stage, int DB_size, size_t at);
int kg_align_free(kmagpu_db *db);

int kg_seed_free(kmagpu_db *db);
int kg_stage1_free(kmagpu_db *db);
int kg_memscore_free(kmagpu_db *db);
int kg_seed_device_output(kmagpu_db *db, const uint8_t **out, const uint32_t **rec_off, int64_t *nreads, size_t *bytes);
int kg_chain_run(kmagpu_db *db, const kmagpu_params *prm, kmagpu_seed_stats *stats);
