// -mem_mode: the "Collecting k-mer scores" pass of runKMA_MEM (runkma.c:1088-1140) with update_Scores_MEM
// (updatescores.c:26-62) and update_Scores_pe_MEM (:64-113) on the GPU. In this mode the reference skips the alignment
// pass: every stage-2 record becomes a frag_raw record whose hits are its candidate templates over their whole length
// (start 0, end = template length), scored with the k-mer score stage 2 gave the read (both mates' scores for a
// pair); a record with a single candidate also adds to the unique scores. One thread per record sizes it and adds
// the ConClave sums atomically, a scan turns sizes into offsets, one warp per record unpacks the 2-bit read
// (unCompDNA, compdna.c:178) and writes the record. Output order = input order.
#include "kmagpu_internal.h"
#include "kmagpu_dev.cuh"
#include <string.h>
#include <vector>

struct MsRec { uint32_t off, toff; int32_t two, score, hits; };   // toff: offset of the record that carries the templates

__global__ void __launch_bounds__(256) ms_sizes_kernel(const uint8_t *__restrict__ in, const uint32_t *__restrict__ off, int n, int k,
		int DB_size, MsRec *recs, uint32_t *size, unsigned long long *as, unsigned long long *uas, unsigned long long *ctr) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n) return;
	const uint8_t *rec = in + off[r];
	MsRec R = {off[r], off[r], 0, 0, 0};
	uint32_t sz = 0;
	const bool empty = off[r + 1] - off[r] < 28;   // a slot of the resident stage-2 stream that holds no record
	const int q_len = empty ? 0 : (int)ld_u32u(rec), sc1 = empty ? 0 : (int)ld_u32u(rec + 12), nt1 = empty ? 1 : (int)ld_u32u(rec + 16);
	// the mate of a pair is consumed together with the record before it (printPair, ankers.c:150)
	const bool is_mate = r > 0 && off[r] - off[r - 1] >= 28 && (int)ld_u32u(in + off[r - 1] + 16) == 0;
	if (!is_mate && !empty) {
		const uint8_t *recT = rec;
		int read_score = 0, q2 = 0, hl2 = 0;
		bool pe = false, ok = true;
		if (nt1 == 0) {
			if (r + 1 >= n) { atomicAdd(&ctr[1], 1ull); ok = false; }
			else { recT = in + off[r + 1]; read_score = abs((int)ld_u32u(recT + 12)); q2 = (int)ld_u32u(recT); hl2 = (int)ld_u32u(recT + 20); pe = true; }
		}
		if (ok && q_len >= k) {
			const int nt = (int)ld_u32u(recT + 16);
			const uint8_t *T = recT + 28 + 8 * (size_t)ld_u32u(recT + 4) + 4 * (size_t)ld_u32u(recT + 8);
			const int last = nt ? (int)ld_u32u(T + 4 * (size_t)(nt - 1)) : 0;
			const bool two = pe && read_score && k <= q2;
			const int score = abs(sc1) + (two ? read_score : 0);
			R.toff = (uint32_t)(recT - in); R.two = two; R.score = score;
			R.hits = (sc1 < 0 && 0 < last) ? -nt : nt;
			sz = 20u + (uint32_t)q_len + ld_u32u(rec + 20) + 12u * (uint32_t)nt + (two ? 12u + (uint32_t)q2 + (uint32_t)hl2 : 0u);
			for (int i = 0; i < nt; ++i) {
				const int t = abs((int)ld_u32u(T + 4 * (size_t)i));
				if (t <= 0 || t >= DB_size) { atomicAdd(&ctr[1], 1ull); continue; }
				atomicAdd(&as[t], (unsigned long long)score);
				if (nt == 1) atomicAdd(&uas[t], (unsigned long long)score);
			}
		}
	}
	recs[r] = R;
	size[r] = sz;
}

// bytes 0-3 of a packed read with its N positions set to 4 (unCompDNA, compdna.c:178), written by one warp
__device__ __forceinline__ void ms_unpack(const uint8_t *rec, uint8_t *o, unsigned lane) {
	const int q_len = (int)ld_u32u(rec), words = (int)ld_u32u(rec + 4), nN = (int)ld_u32u(rec + 8);
	const uint8_t *seq = rec + 28, *N = seq + 8 * (size_t)words;
	for (int i = lane; i < q_len; i += 32) o[i] = (uint8_t)((ld_u64u(seq + 8 * (size_t)(i >> 5)) << ((i & 31) << 1)) >> 62);
	__syncwarp();
	for (int i = lane; i < nN; i += 32) o[ld_u32u(N + 4 * (size_t)i)] = 4;
	__syncwarp();
}

__global__ void __launch_bounds__(256) ms_emit_kernel(const uint8_t *__restrict__ in, const MsRec *__restrict__ recs, int n,
		const uint32_t *__restrict__ size, const uint32_t *__restrict__ out_off, const int32_t *__restrict__ lengths, uint8_t *__restrict__ out) {
	const unsigned lane = threadIdx.x & 31;
	const int warps = (gridDim.x * blockDim.x) >> 5;
	for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
		if (!size[r]) continue;
		const MsRec R = recs[r];
		const uint8_t *rec = in + R.off, *recT = in + R.toff;
		const int q_len = (int)ld_u32u(rec), hl = (int)ld_u32u(rec + 20), flag = (int)ld_u32u(rec + 24);
		const int nt = abs(R.hits);
		uint8_t *o = out + out_off[r];
		if (lane < 5) {
			const int32_t h = lane == 0 ? q_len : lane == 1 ? R.hits : lane == 2 ? (R.two ? -R.score : R.score) : lane == 3 ? hl : flag;
			st_u32b(o + 4 * lane, (uint32_t)h);
		}
		o += 20;
		ms_unpack(rec, o, lane);
		o += q_len;
		const uint8_t *hdr = rec + 28 + 8 * (size_t)ld_u32u(rec + 4) + 4 * (size_t)ld_u32u(rec + 8) + 4 * (size_t)ld_u32u(rec + 16);
		for (int i = lane; i < hl; i += 32) o[i] = hdr[i];
		o += hl;
		const uint8_t *T = recT + 28 + 8 * (size_t)ld_u32u(recT + 4) + 4 * (size_t)ld_u32u(recT + 8);
		for (int i = lane; i < nt; i += 32) {
			const int t = (int)ld_u32u(T + 4 * (size_t)i);
			st_u32b(o + 4 * (size_t)i, 0u);                                              // best_start_pos: zeros (calloc, runkma.c:1079)
			st_u32b(o + 4 * (size_t)(nt + i), (uint32_t)__ldg(lengths + abs(t)));        // best_end_pos = template length
			st_u32b(o + 4 * (size_t)(2 * nt + i), (uint32_t)t);
		}
		o += 12 * (size_t)nt;
		if (R.two) {
			const int q2 = (int)ld_u32u(recT), hl2 = (int)ld_u32u(recT + 20);
			if (lane < 3) st_u32b(o + 4 * lane, lane == 0 ? (uint32_t)q2 : lane == 1 ? (uint32_t)hl2 : ld_u32u(recT + 24));
			o += 12;
			ms_unpack(recT, o, lane);
			o += q2;
			const uint8_t *hdr2 = T + 4 * (size_t)nt;
			for (int i = lane; i < hl2; i += 32) o[i] = hdr2[i];
		}
	}
}

int kg_memscore_free(kmagpu_db *db) {
	RawBatch &w = db->raw;
	KgBuf *all[] = {&w.d_in, &w.d_off, &w.d_recs, &w.d_sz, &w.d_partial, &w.d_ctr, &w.d_acc, &w.d_out};
	for (KgBuf *x : all) x->release();
	w.valid = false;
	return 0;
}

// n stage-2 slots in HBM (din, offsets doff[n + 1]) -> frag_raw stream in db->raw (+ optional download) and the score sums
static int memscore_core(kmagpu_db *db, const uint8_t *din, const uint32_t *doff, int n, void *frag_out, size_t out_cap, size_t *out_bytes,
                         uint64_t *alignment_scores, uint64_t *uniq_alignment_scores) {
	const int DB = db->info.DB_size, ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
	cudaStream_t st = db->stream;
	RawBatch &w = db->raw;
	w.valid = false;
	if (w.d_recs.reserve(sizeof(MsRec) * (size_t)n) || w.d_sz.reserve(4 * (size_t)(2 * n + 4)) || w.d_partial.reserve(4 * (size_t)(ntiles + 2)) ||
	    w.d_ctr.reserve(64) || w.d_acc.reserve(16 * (size_t)DB)) return -1;
	uint32_t *size = (uint32_t *)w.d_sz.p, *ooff = size + n + 1;
	unsigned long long *ctr = (unsigned long long *)w.d_ctr.p, *as = (unsigned long long *)w.d_acc.p, *uas = as + DB;
	KG_CUDA(cudaMemsetAsync(ctr, 0, 64, st));
	KG_CUDA(cudaMemsetAsync(w.d_acc.p, 0, 16 * (size_t)DB, st));
	ms_sizes_kernel<<<(n + 255) / 256, 256, 0, st>>>(din, doff, n, db->info.kmerindex, DB, (MsRec *)w.d_recs.p, size, as, uas, ctr);
	kg_exscan(size, n, ooff, (uint32_t *)w.d_partial.p, ctr + 2, st);
	unsigned long long h[8];
	KG_CUDA(cudaMemcpyAsync(h, ctr, 64, cudaMemcpyDeviceToHost, st));
	KG_CUDA(cudaStreamSynchronize(st));
	KG_CUDA(cudaGetLastError());
	if (h[1]) { kmagpu_set_error("%llu stage-2 records are truncated pairs or name a template outside the database", h[1]); return -1; }
	KG_SCAN_FITS(h[2], "the frag_raw stream");
	const size_t ob = (size_t)h[2];
	if (out_bytes) *out_bytes = ob;
	if (frag_out && ob > out_cap) { kmagpu_set_error("frag_raw output needs %zu bytes, caller gave %zu", ob, out_cap); return -1; }
	if (w.d_out.reserve(ob + 64)) return -1;
	ms_emit_kernel<<<kg_wave_grid(ms_emit_kernel, 256, db->sm_count), 256, 0, st>>>(din, (const MsRec *)w.d_recs.p, n, size, ooff, db->d_lengths, (uint8_t *)w.d_out.p);
	const uint32_t total = (uint32_t)ob;
	KG_CUDA(cudaMemcpyAsync(ooff + n, &total, 4, cudaMemcpyHostToDevice, st));   // closing offset for the next stage
	KG_CUDA(cudaMemsetAsync((uint8_t *)w.d_out.p + ob, 0, 64, st));
	if (frag_out && ob) KG_CUDA(cudaMemcpyAsync(frag_out, w.d_out.p, ob, cudaMemcpyDeviceToHost, st));
	std::vector<uint64_t> acc(2 * (size_t)DB);
	kg_scores_accumulate(db, as);   // the run-wide sums on the device (kmagpu_scores_reset)
	KG_CUDA(cudaMemcpyAsync(acc.data(), w.d_acc.p, 16 * (size_t)DB, cudaMemcpyDeviceToHost, st));
	KG_CUDA(cudaStreamSynchronize(st));
	KG_CUDA(cudaGetLastError());
	for (int t = 0; t < DB; ++t) {
		if (alignment_scores) alignment_scores[t] += acc[t];
		if (uniq_alignment_scores) uniq_alignment_scores[t] += acc[(size_t)DB + t];
	}
	w.off = ooff; w.n = n; w.bytes = ob; w.valid = true;
	return 0;
}

extern "C" int kmagpu_memscore_batch(kmagpu_db *db, const void *stage2, size_t nbytes, void *frag_out, size_t out_cap, size_t *out_bytes,
                                     uint64_t *alignment_scores, uint64_t *uniq_alignment_scores, int64_t *nrecords) {
	if (!db || (!stage2 && nbytes)) { kmagpu_set_error("null argument"); return -1; }
	if (!db->d_lengths) { kmagpu_set_error("database has no template lengths (.length.b missing)"); return -1; }
	if (nbytes >= (1ull << 32) - 64) { kmagpu_set_error("stage-2 batch of %zu bytes exceeds the 4 GiB per-call limit; split it", nbytes); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	if (out_bytes) *out_bytes = 0;
	if (nrecords) *nrecords = 0;
	db->raw.valid = false;
	size_t used = 0;
	const int64_t n64 = kmagpu_record_walk(2, stage2, nbytes, nullptr, 0, &used);
	if (n64 < 0) return -1;
	const int n = (int)n64;
	if (nrecords) *nrecords = n64;
	if (n == 0) return 0;
	std::vector<uint64_t> off64((size_t)n);
	kmagpu_record_walk(2, stage2, nbytes, off64.data(), (size_t)n, &used);
	std::vector<uint32_t> off((size_t)n + 1);
	for (int i = 0; i < n; ++i) {
		off[i] = (uint32_t)off64[i];
		if (kg_check_record((const uint8_t *)stage2 + off64[i], 2, db->info.DB_size, (size_t)off64[i])) return -1;
	}
	off[n] = (uint32_t)used;
	RawBatch &w = db->raw;
	if (w.d_in.reserve(used + 64) || w.d_off.reserve(4 * ((size_t)n + 2))) return -1;
	cudaStream_t st = db->stream;
	KG_CUDA(cudaMemcpyAsync(w.d_in.p, stage2, used, cudaMemcpyHostToDevice, st));
	KG_CUDA(cudaMemsetAsync((uint8_t *)w.d_in.p + used, 0, 64, st));
	KG_CUDA(cudaMemcpyAsync(w.d_off.p, off.data(), 4 * ((size_t)n + 1), cudaMemcpyHostToDevice, st));
	KG_CUDA(cudaStreamSynchronize(st));
	return memscore_core(db, (const uint8_t *)w.d_in.p, (const uint32_t *)w.d_off.p, n, frag_out, out_cap, out_bytes, alignment_scores,
	                     uniq_alignment_scores);
}

// The same on the stage-2 stream the last kmagpu_seed_run left in HBM.
extern "C" int kmagpu_memscore_from_seed(kmagpu_db *db, void *frag_out, size_t out_cap, size_t *out_bytes, uint64_t *alignment_scores,
                                         uint64_t *uniq_alignment_scores, int64_t *nrecords) {
	if (!db) { kmagpu_set_error("null argument"); return -1; }
	if (!db->d_lengths) { kmagpu_set_error("database has no template lengths (.length.b missing)"); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	if (out_bytes) *out_bytes = 0;
	if (nrecords) *nrecords = 0;
	db->raw.valid = false;
	const uint8_t *out; const uint32_t *roff; int64_t n; size_t bytes;
	if (kg_seed_device_output(db, &out, &roff, &n, &bytes)) return -1;
	if (nrecords) *nrecords = n;
	if (n == 0) return 0;
	RawBatch &w = db->raw;   // the slot offsets with their closing entry
	if (w.d_off.reserve(4 * ((size_t)n + 2))) return -1;
	const uint32_t total = (uint32_t)bytes;
	KG_CUDA(cudaMemcpyAsync(w.d_off.p, roff, 4 * (size_t)n, cudaMemcpyDeviceToDevice, db->stream));
	KG_CUDA(cudaMemcpyAsync((uint32_t *)w.d_off.p + n, &total, 4, cudaMemcpyHostToDevice, db->stream));
	KG_CUDA(cudaStreamSynchronize(db->stream));
	return memscore_core(db, out, (const uint32_t *)w.d_off.p, (int)n, frag_out, out_cap, out_bytes, alignment_scores, uniq_alignment_scores);
}
