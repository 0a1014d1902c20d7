// ConClave's choice pass and the per-template bucketing that follows it, on the GPU.
//
// What is computed is runConClave (conclave.c:43-213, the default -ConClave 1) + printFrags (frags.c:30-61) for one
// chunk of frag_raw records (updatescores.c:284-295): per read the template with the largest GLOBAL alignment score
// wins (ties: score per template base as a double, then unique score, then the smaller template id), reads chosen on
// the reverse strand are reverse-complemented (strrc, stdnuc.c:450) with their query bounds mirrored, the weighted
// scores and read / fragment counts are summed per template, and the fragments leave grouped by template -- inside a
// template in REVERSE arrival order (the reference prepends to a linked list), the mate of a pair before its first
// read. The output is the per-template fragment stream kmagpu_trace_batch consumes (frags.c:45-48).
// How it is computed is not the reference's serial re-read with a malloc per read:
//   * one thread per record walks its candidate list against the score arrays (HBM/L2 resident) and makes the choice
//     with the reference's int truncations (best_read_score / bestNum are ints compared with the 64-bit sums);
//   * the order "template ascending, arrival descending" is one 64-bit radix sort (CUB) of (template, ~arrival) keys;
//   * a scan turns record sizes into offsets and one warp per fragment writes its record.
#include "kmagpu_internal.h"
#include "kmagpu_dev.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <string.h>
#include <vector>

struct CcItem {
	uint32_t q_off, hdr_off;     // byte offsets of the read bytes / name in the frag_raw stream
	int32_t q_len, hl, tmpl, bestHits, score, start, end, flag, rc, has_bound, b0, b1;
};

// the 4-key order over a record's candidates (conclave.c:66-113; runConClave_lc :238-285 with lc): returns the winning
// candidate's index (-1: none) with the reference's int truncations (best_read_score / bestNum are ints compared with the
// 64-bit sums). first = the value bestTemplate starts from: -1 in runConClave, 0 in runConClave2's fallback (conclave.c:598),
// which only matters to the last tie rule (the smaller template id wins against abs(first)).
__device__ int cc_four_keys(const uint8_t *T, int bestHits, const unsigned long long *__restrict__ as, const unsigned long long *__restrict__ uas,
		const int32_t *__restrict__ lengths, int DB_size, int lc, int first, unsigned long long *ctr) {
	double bestScore = 0;
	int best_read_score = 0, bestNum = 0, bestTemplate = first, besti = -1;
	for (int i = 0; i < bestHits; ++i) {
		const int tt = (int)ld_u32u(T + 4 * (size_t)i);
		const int t = tt < 0 ? -tt : tt;
		if (t <= 0 || t >= DB_size) { atomicAdd(&ctr[1], 1ull); continue; }
		const unsigned long long a = as[t], u = uas[t];
		const double tmp_score = __ddiv_rn(1.0 * (double)a, (double)__ldg(lengths + t));
		// the reference compares the 64-bit sums with ints: the ints are converted (sign-extended) to unsigned long
		const unsigned long long brs = (unsigned long long)(long long)best_read_score, bn = (unsigned long long)(long long)bestNum;
		bool take = false;
		if (lc) {   // runConClave_lc (conclave.c:215-384, -lc): score per template base first, then the total
			if (tmp_score > bestScore) take = true;
			else if (tmp_score == bestScore) {
				if (a > brs) take = true;
				else if (a == brs) {
					if (u > bn) take = true;
					else if (u == bn && t < abs(bestTemplate)) take = true;
				}
			}
		} else if (a > brs) take = true;
		else if (a == brs) {
			if (tmp_score > bestScore) take = true;
			else if (tmp_score == bestScore) {
				if (u > bn) take = true;
				else if (u == bn && t < abs(bestTemplate)) take = true;
			}
		}
		if (take) { bestTemplate = tt; best_read_score = (int)a; bestScore = tmp_score; bestNum = (int)u; besti = i; }
	}
	return besti;
}

// the final choice of runConClave2 for a record with bestHits != 1 (conclave.c:547-655): a candidate drawn with probability
// proportional to the (updated) unique scores -- Lehmer generator 16807 seeded from the read's first and last 7 bases --
// else the 4-key order; -1: no candidate (the record is skipped)
__device__ int cc2_choice(const uint8_t *q, int q_len, const uint8_t *T, int bestHits, const unsigned long long *__restrict__ as,
		const unsigned long long *__restrict__ uas, const int32_t *__restrict__ lengths, int DB_size, int lc, unsigned long long *ctr) {
	int tot = 0;
	for (int i = bestHits; i--;) {
		const int t = abs((int)ld_u32u(T + 4 * (size_t)i));
		if (t > 0 && t < DB_size) tot = (int)((long long)tot + (long long)uas[t]);   // int += unsigned long: the low 32 bits
	}
	if (tot && 16 <= q_len) {
		int rnd = q[0], i = -1, j = q_len;
		while (++i < 7) rnd = (((rnd << 2) | q[i]) << 2) | q[--j];
		rnd = 16807 * (rnd % 127773) - 2836 * (rnd / 127773);   // minimal standard
		if (rnd <= 0) rnd += 0x7fffffff;
		const double tmp_score = __ddiv_rn((double)rnd, 2147483647.0);
		const unsigned randScore = __double2uint_rz(__dmul_rn(tmp_score, (double)tot));
		unsigned long long score = 0;
		for (i = 0; i != bestHits; ++i) {
			const int tt = (int)ld_u32u(T + 4 * (size_t)i), t = abs(tt);
			if (t <= 0 || t >= DB_size) continue;
			score += uas[t];
			if ((unsigned long long)randScore < score) return tt ? i : -1;
		}
	}
	return cc_four_keys(T, bestHits, as, uas, lengths, DB_size, lc, 0, ctr);
}

// one frag_raw record at byte `pos`: the choice among its candidates, the item(s) it contributes (idx, and idx + 1 for
// the mate block of a pair record); returns the bytes the record spans
__device__ uint32_t cc_record(const uint8_t *__restrict__ in, uint32_t pos, int idx, const unsigned long long *__restrict__ as,
		const unsigned long long *__restrict__ uas, const int32_t *__restrict__ lengths, int DB_size, CcItem *items, unsigned long long *keys,
		unsigned long long *w, unsigned int *fc, unsigned int *rcn, unsigned long long *ctr, bool *has_mate, int lc, int version) {
	const uint8_t *rec = in + pos;
	const int q_len = (int)ld_u32u(rec), sparse = (int)ld_u32u(rec + 4), sc = (int)ld_u32u(rec + 8), hl = (int)ld_u32u(rec + 12);
	int flag = (int)ld_u32u(rec + 16);
	const int bestHits = abs(sparse), read_score = abs(sc);
	const uint8_t *S = rec + 20 + (size_t)q_len + (size_t)hl, *E = S + 4 * (size_t)bestHits, *T = E + 4 * (size_t)bestHits;
	int bestTemplate = 0, start = 0, end = 0;
	if (version == 2 ? bestHits != 1 : bestHits > 1) {
		const int bi = version == 2 ? cc2_choice(rec + 20, q_len, T, bestHits, as, uas, lengths, DB_size, lc, ctr)
		                            : cc_four_keys(T, bestHits, as, uas, lengths, DB_size, lc, -1, ctr);
		if (bi >= 0) { bestTemplate = (int)ld_u32u(T + 4 * (size_t)bi); start = (int)ld_u32u(S + 4 * (size_t)bi); end = (int)ld_u32u(E + 4 * (size_t)bi); }
		else bestTemplate = version == 2 ? 0 : -1;
	} else { bestTemplate = (int)ld_u32u(T); start = (int)ld_u32u(S); end = (int)ld_u32u(E); }
	const bool skipped = version == 2 && bestTemplate == 0;   // runConClave2 without a candidate: the record leaves no fragment (conclave.c:722)
	CcItem a;
	a.q_off = pos + 20u; a.hdr_off = a.q_off + (uint32_t)q_len; a.q_len = q_len; a.hl = hl;
	a.rc = 0; a.has_bound = 0; a.b0 = a.b1 = 0;
	if (bestTemplate < 0) {
		bestTemplate = -bestTemplate; a.rc = 1; flag |= 16;
		const uint8_t *hdr = rec + 20 + q_len;
		if (9 < hl && hdr[hl - 9] == 0) {   // mirrored query bounds (conclave.c:131-140)
			a.has_bound = 1;
			a.b0 = q_len - (int)ld_u32u(hdr + hl - 4);
			a.b1 = q_len - (int)ld_u32u(hdr + hl - 8);
		}
	}
	const bool ok = bestTemplate > 0 && bestTemplate < DB_size;
	if (!ok && !skipped) atomicAdd(&ctr[1], 1ull);
	a.tmpl = bestTemplate; a.bestHits = bestHits; a.score = sparse < 0 ? 0 : read_score; a.start = start; a.end = end; a.flag = flag;
	const bool mate = sc < 0;
	if (ok) {
		atomicAdd(&w[bestTemplate], (unsigned long long)read_score);
		atomicAdd(&fc[bestTemplate], 1u);
		atomicAdd(&rcn[bestTemplate], mate ? 2u : 1u);
	}
	items[idx] = a;
	keys[idx] = ok ? ((unsigned long long)(unsigned)bestTemplate << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)idx) : ~0ull;
	uint32_t bytes = 20u + (uint32_t)q_len + (uint32_t)hl + 12u * (uint32_t)bestHits;
	if (mate) {   // the mate block: int32[3]{q_len, hdrlen, flag} + bytes; same template and span, never turned
		const uint8_t *m = T + 4 * (size_t)bestHits;
		CcItem b = a;
		b.q_len = (int)ld_u32u(m); b.hl = (int)ld_u32u(m + 4); b.flag = (int)ld_u32u(m + 8);
		b.q_off = (uint32_t)(m + 12 - in); b.hdr_off = b.q_off + (uint32_t)b.q_len;
		b.rc = 0; b.has_bound = 0;
		items[idx + 1] = b;
		if (ok) keys[idx + 1] = ((unsigned long long)(unsigned)bestTemplate << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)(idx + 1));
		bytes += 12u + (uint32_t)b.q_len + (uint32_t)b.hl;
	}
	atomicAdd(&ctr[0], (ok ? 1ull : 0ull) + (ok && mate ? 1ull : 0ull));
	*has_mate = mate;
	return bytes;
}


// one thread per slot of the frag_raw stream: a slot holds one record (host streams: every record is its own slot), none
// (resident -mem_mode stream) or up to two (resident alignment-pass stream: a pair resolved as two single reads)
__global__ void __launch_bounds__(256) cc_choose_kernel(const uint8_t *__restrict__ in, const uint32_t *__restrict__ off, int n,
		const unsigned long long *__restrict__ as, const unsigned long long *__restrict__ uas, const int32_t *__restrict__ lengths,
		int DB_size, CcItem *items, unsigned long long *keys, uint32_t *vals, unsigned long long *w, unsigned int *fc, unsigned int *rcn,
		unsigned long long *ctr, int lc, int version) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n) return;
	keys[2 * r] = ~0ull; keys[2 * r + 1] = ~0ull; vals[2 * r] = 2u * (unsigned)r; vals[2 * r + 1] = 2u * (unsigned)r + 1u;
	uint32_t pos = off[r];
	const uint32_t end = off[r + 1];
	for (int idx = 2 * r; idx < 2 * r + 2 && end - pos >= 20u && pos < end;) {
		bool mate = false;
		pos += cc_record(in, pos, idx, as, uas, lengths, DB_size, items, keys, w, fc, rcn, ctr, &mate, lc, version);
		idx += mate ? 2 : 1;
	}
	if (pos != end && end - off[r] >= 20u) atomicAdd(&ctr[3], 1ull);   // the slot's bytes are not a whole number of records
}

// runConClave2's two passes before the final choice, one thread per slot like cc_choose_kernel:
// pass 0 (conclave.c:405-465): the provisional choice (4-key order) adds the read score to w[template];
// pass 1 (:493-530): a read with several candidates of which exactly ONE kept its (significant) w adds its score to that
// template's unique score.
__global__ void __launch_bounds__(256) cc2_pre_kernel(const uint8_t *__restrict__ in, const uint32_t *__restrict__ off, int n,
		const unsigned long long *__restrict__ as, unsigned long long *uas, const int32_t *__restrict__ lengths, int DB_size,
		unsigned long long *w, unsigned long long *ctr, int lc, int pass) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n) return;
	uint32_t pos = off[r];
	const uint32_t end = off[r + 1];
	for (int k = 0; k < 2 && end - pos >= 20u && pos < end; ++k) {
		const uint8_t *rec = in + pos;
		const int q_len = (int)ld_u32u(rec), bestHits = abs((int)ld_u32u(rec + 4)), sc = (int)ld_u32u(rec + 8), hl = (int)ld_u32u(rec + 12);
		const int read_score = abs(sc);
		const uint8_t *T = rec + 20 + (size_t)q_len + (size_t)hl + 8 * (size_t)bestHits;
		if (pass == 0) {
			int best = 0;
			if (bestHits > 1) { const int bi = cc_four_keys(T, bestHits, as, uas, lengths, DB_size, lc, -1, ctr); best = bi >= 0 ? (int)ld_u32u(T + 4 * (size_t)bi) : -1; }
			else best = (int)ld_u32u(T);
			const int t = abs(best);
			if (t > 0 && t < DB_size) atomicAdd(&w[t], (unsigned long long)read_score);
			else atomicAdd(&ctr[1], 1ull);
		} else if (bestHits != 1) {
			int best = 0;
			for (int i = bestHits; i--;) {
				const int t = abs((int)ld_u32u(T + 4 * (size_t)i));
				if (t > 0 && t < DB_size && w[t]) { if (best) { best = 0; break; } else best = t; }
			}
			if (best) atomicAdd(&uas[best], (unsigned long long)read_score);
		}
		pos += 20u + (uint32_t)q_len + (uint32_t)hl + 12u * (uint32_t)bestHits;
		if (sc < 0) { const uint8_t *m = in + pos; pos += 12u + ld_u32u(m) + ld_u32u(m + 4); k = 2; }   // the mate block closes the slot
	}
}

__global__ void __launch_bounds__(256) cc_sizes_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ vals,
		const CcItem *__restrict__ items, int nitems, uint32_t *size) {
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= nitems) return;
	uint32_t s = 0;
	if (keys[k] != ~0ull) { const CcItem it = items[vals[k]]; s = 32u + (uint32_t)it.q_len + (uint32_t)it.hl; }
	size[k] = s;
}

// one warp per fragment, in output order: int32 template, int32[7] {q_len, bestHits, score, start, end, hdrlen, flag},
// read bytes (0-5 codes), name (frags.c:45-48)
__global__ void __launch_bounds__(256) cc_emit_kernel(const uint8_t *__restrict__ in, const unsigned long long *__restrict__ keys,
		const uint32_t *__restrict__ vals, const CcItem *__restrict__ items, int nitems, const uint32_t *__restrict__ out_off,
		uint8_t *__restrict__ out) {
	const unsigned lane = threadIdx.x & 31;
	const int warps = (gridDim.x * blockDim.x) >> 5;
	for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < nitems; k += warps) {
		if (keys[k] == ~0ull) continue;
		const CcItem it = items[vals[k]];
		uint8_t *o = out + out_off[k];
		if (lane < 8) {
			const int32_t h = lane == 0 ? it.tmpl : lane == 1 ? it.q_len : lane == 2 ? it.bestHits : lane == 3 ? it.score
			                : lane == 4 ? it.start : lane == 5 ? it.end : lane == 6 ? it.hl : it.flag;
			st_u32b(o + 4 * lane, (uint32_t)h);
		}
		o += 32;
		const uint8_t *q = in + it.q_off, *hdr = in + it.hdr_off;
		if (it.rc) {
			for (int i = lane; i < it.q_len; i += 32) { const uint8_t c = q[it.q_len - 1 - i]; o[i] = c < 4 ? (uint8_t)(3 - c) : c; }
		} else for (int i = lane; i < it.q_len; i += 32) o[i] = q[i];
		o += it.q_len;
		for (int i = lane; i < it.hl; i += 32) o[i] = hdr[i];
		__syncwarp();
		if (it.has_bound && lane == 0) { st_u32b(o + it.hl - 8, (uint32_t)it.b0); st_u32b(o + it.hl - 4, (uint32_t)it.b1); }
	}
}

__global__ void cc_tail_kernel(uint8_t *out, const unsigned long long *total) { st_u32b(out + *total, 0xFFFFFFFFu); }

// ConClave over n frag_raw slots in HBM (din, offsets doff[n + 1]; empty slots allowed)
static int conclave_core(kmagpu_db *db, const uint8_t *din, const uint32_t *doff, int n, const uint64_t *alignment_scores,
                         const uint64_t *uniq_alignment_scores, void *frags_out, size_t out_cap, size_t *out_bytes, uint64_t *w_scores,
                         uint32_t *fragmentCounts, uint32_t *readCounts) {
	const int DB = db->info.DB_size;
	cudaStream_t st = db->stream;
	const int ni = 2 * n, ntiles = (ni + SCAN_TILE - 1) / SCAN_TILE;
	FragBatch &fb = db->frg;   // working buffers persist; the fragment stream and its offsets stay for kmagpu_trace_from_conclave
	KgBuf &d_sc = fb.d_sc, &d_items = fb.d_items, &d_keys = fb.d_keys, &d_vals = fb.d_vals, &d_partial = fb.d_partial, &d_ctr = fb.d_ctr,
	      &d_acc = fb.d_acc, &d_tmp = fb.d_tmp;
	KgBuf &d_sz = fb.d_sz, &d_out = fb.d_out;
	if (d_sc.reserve(16 * (size_t)DB) ||
	    d_items.reserve(sizeof(CcItem) * (size_t)ni) || d_keys.reserve(16 * (size_t)ni) || d_vals.reserve(8 * (size_t)ni) ||
	    d_sz.reserve(4 * (size_t)(2 * ni + 4)) || d_partial.reserve(4 * (size_t)(ntiles + 2)) || d_ctr.reserve(64) ||
	    d_acc.reserve(16 * (size_t)DB)) return -1;
	unsigned long long *as = (unsigned long long *)d_sc.p, *uas = as + DB;
	unsigned long long *keys = (unsigned long long *)d_keys.p, *keys2 = keys + ni;
	uint32_t *vals = (uint32_t *)d_vals.p, *vals2 = vals + ni;
	uint32_t *size = (uint32_t *)d_sz.p, *ooff = size + ni + 1;
	unsigned long long *ctr = (unsigned long long *)d_ctr.p;
	unsigned long long *w = (unsigned long long *)d_acc.p;
	unsigned int *fc = (unsigned int *)(w + DB), *rcn = fc + DB;
	if (alignment_scores) {
		KG_CUDA(cudaMemcpyAsync(as, alignment_scores, 8 * (size_t)DB, cudaMemcpyHostToDevice, st));
		KG_CUDA(cudaMemcpyAsync(uas, uniq_alignment_scores, 8 * (size_t)DB, cudaMemcpyHostToDevice, st));
	} else {   // the run-wide sums this handle holds on the device (kmagpu_scores_reset ... kmagpu_allreduce_scores): they never left HBM
		if (!db->image->d_run_scores) { kmagpu_set_error("ConClave without score arrays needs the device-resident sums (kmagpu_scores_reset)"); return -1; }
		KG_CUDA(cudaMemcpyAsync(as, db->image->d_run_scores, 16 * (size_t)DB, cudaMemcpyDeviceToDevice, st));
	}
	KG_CUDA(cudaMemsetAsync(ctr, 0, 64, st));
	KG_CUDA(cudaMemsetAsync(d_acc.p, 0, 16 * (size_t)DB, st));
	const int version = db->cc2.version == 2 ? 2 : 1;
	if (version == 2) {
		// runConClave2 (conclave.c:386-747): provisional sums -> significance on the host (long double arithmetic and the caller's
		// p_chisqr, exactly as conclave.c:467-491) -> unique-score update -> the sums start again for the final choice
		if (!db->cc2.p_chisqr) { kmagpu_set_error("ConClave 2 needs the caller's p_chisqr (kmagpu_conclave_version)"); return -1; }
		cc2_pre_kernel<<<(n + 255) / 256, 256, 0, st>>>(din, doff, n, as, uas, db->d_lengths, DB, w, ctr, db->conclave_lc, 0);
		std::vector<unsigned long> hw((size_t)DB);
		KG_CUDA(cudaMemcpyAsync(hw.data(), w, 8 * (size_t)DB, cudaMemcpyDeviceToHost, st));
		KG_CUDA(cudaStreamSynchronize(st));
		KG_CUDA(cudaGetLastError());
		unsigned long Nhits = 0, template_tot_ulen = 0;
		for (int t = 1; t < DB; ++t) { Nhits += hw[t]; template_tot_ulen += (unsigned long)db->lengths[t]; }
		for (int t = DB; --t;) {
			int read_score;
			if ((read_score = (int)hw[t])) {
				const int t_len = db->lengths[t];
				long double expected = t_len, q_value;
				expected /= (1 < (template_tot_ulen - t_len) ? (template_tot_ulen - t_len) : 1);
				expected *= (Nhits - read_score);
				q_value = read_score - expected;
				q_value /= (expected + read_score);
				q_value *= read_score - expected;
				const double p_value = db->cc2.p_chisqr(q_value);
				const int a = (p_value <= db->cc2.evalue && read_score > expected), b = (read_score >= db->cc2.scoreT * t_len);
				if ((db->cc2.and_mode ? (a && b) : (a || b)) == 0) hw[t] = 0;
			}
		}
		KG_CUDA(cudaMemcpyAsync(w, hw.data(), 8 * (size_t)DB, cudaMemcpyHostToDevice, st));
		cc2_pre_kernel<<<(n + 255) / 256, 256, 0, st>>>(din, doff, n, as, uas, db->d_lengths, DB, w, ctr, db->conclave_lc, 1);
		KG_CUDA(cudaMemsetAsync(d_acc.p, 0, 16 * (size_t)DB, st));
		KG_CUDA(cudaStreamSynchronize(st));   // hw is a host vector: the upload must be done before it goes
	}
	cc_choose_kernel<<<(n + 255) / 256, 256, 0, st>>>(din, doff, n, as, uas, db->d_lengths, DB,
		(CcItem *)d_items.p, keys, vals, w, fc, rcn, ctr, db->conclave_lc, version);
	size_t tmp_bytes = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, vals, vals2, ni, 0, 64, st);
	if (d_tmp.reserve(tmp_bytes + 64)) return -1;
	cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, keys, keys2, vals, vals2, ni, 0, 64, st);
	cc_sizes_kernel<<<(ni + 255) / 256, 256, 0, st>>>(keys2, vals2, (const CcItem *)d_items.p, ni, size);
	kg_exscan(size, ni, ooff, (uint32_t *)d_partial.p, ctr + 2, st);
	unsigned long long h[8];
	KG_CUDA(cudaMemcpyAsync(h, ctr, 64, cudaMemcpyDeviceToHost, st));
	KG_CUDA(cudaStreamSynchronize(st));
	KG_CUDA(cudaGetLastError());
	if (h[1]) { kmagpu_set_error("%llu frag_raw candidates name a template outside the database", h[1]); return -1; }
	if (h[3]) { kmagpu_set_error("%llu slots of the frag_raw stream do not hold whole records", h[3]); return -1; }
	KG_SCAN_FITS(h[2] + 4, "the fragment stream");
	const size_t ob = (size_t)h[2] + 4;
	if (out_bytes) *out_bytes = ob;
	if (frags_out && ob > out_cap) { kmagpu_set_error("fragment output needs %zu bytes, caller gave %zu", ob, out_cap); return -1; }
	if (d_out.reserve(ob + 64)) return -1;
	cc_emit_kernel<<<kg_wave_grid(cc_emit_kernel, 256, db->sm_count), 256, 0, st>>>(din, keys2, vals2, (const CcItem *)d_items.p, ni, ooff,
		(uint8_t *)d_out.p);
	cc_tail_kernel<<<1, 1, 0, st>>>((uint8_t *)d_out.p, ctr + 2);
	KG_CUDA(cudaMemsetAsync((uint8_t *)d_out.p + ob, 0, 64, st));
	if (frags_out) KG_CUDA(cudaMemcpyAsync(frags_out, d_out.p, ob, cudaMemcpyDeviceToHost, st));
	std::vector<uint8_t> acc(16 * (size_t)DB);
	KG_CUDA(cudaMemcpyAsync(acc.data(), d_acc.p, 16 * (size_t)DB, cudaMemcpyDeviceToHost, st));
	KG_CUDA(cudaStreamSynchronize(st));
	KG_CUDA(cudaGetLastError());
	db->frg.off = ooff; db->frg.n = (int64_t)h[0]; db->frg.bytes = ob; db->frg.valid = true;   // valid items sort first
	const uint64_t *hw = (const uint64_t *)acc.data();
	const uint32_t *hfc = (const uint32_t *)(hw + DB), *hrc = hfc + DB;
	for (int t = 0; t < DB; ++t) {
		if (w_scores) w_scores[t] += hw[t];
		if (fragmentCounts) fragmentCounts[t] += hfc[t];
		if (readCounts) readCounts[t] += hrc[t];
	}
	return 0;
}
// printFrags of a chunk without fragments: the terminator alone (what kmagpu_conclave_batch does for n == 0)
static int conclave_empty(kmagpu_db *db, void *frags_out, size_t out_cap, size_t *out_bytes) {
	if (out_bytes) *out_bytes = 4;
	db->frg.valid = true; db->frg.n = 0; db->frg.bytes = 4;
	if (!frags_out) return 0;
	if (out_cap < 4) { kmagpu_set_error("fragment output needs 4 bytes"); return -1; }
	const int32_t m1 = -1;
	memcpy(frags_out, &m1, 4);
	return 0;
}

extern "C" int kmagpu_conclave_batch(kmagpu_db *db, const void *frag_raw, size_t nbytes, const uint64_t *alignment_scores,
                                     const uint64_t *uniq_alignment_scores, void *frags_out, size_t out_cap, size_t *out_bytes,
                                     uint64_t *w_scores, uint32_t *fragmentCounts, uint32_t *readCounts, int64_t *nrecords) {
	if (!db || (!frag_raw && nbytes) || (!alignment_scores != !uniq_alignment_scores)) { kmagpu_set_error("null argument"); return -1; }
	db->frg.valid = false; db->frg.n = 0; db->frg.bytes = 0;
	if (!db->d_lengths) { kmagpu_set_error("database has no template lengths (.length.b missing)"); return -1; }
	if (nbytes >= (1ull << 32) - 64) { kmagpu_set_error("frag_raw batch of %zu bytes exceeds the 4 GiB per-call limit; split it", nbytes); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	if (out_bytes) *out_bytes = 0;
	if (nrecords) *nrecords = 0;
	size_t used = 0;
	const int64_t n64 = kmagpu_record_walk(4, frag_raw, nbytes, nullptr, 0, &used);
	if (n64 < 0) return -1;
	const int n = (int)n64;
	if (nrecords) *nrecords = n64;
	cudaStream_t st = db->stream;
	if (n == 0) {   // printFrags of an empty chunk: the terminator alone
		if (out_bytes) *out_bytes = 4;
		db->frg.valid = true;
		if (!frags_out) return 0;
		if (out_cap < 4) { kmagpu_set_error("fragment output needs 4 bytes"); return -1; }
		const int32_t m1 = -1;
		memcpy(frags_out, &m1, 4);
		return 0;
	}
	std::vector<uint64_t> off64((size_t)n);
	kmagpu_record_walk(4, frag_raw, nbytes, off64.data(), (size_t)n, &used);
	std::vector<uint32_t> off((size_t)n + 1);
	for (int i = 0; i < n; ++i) off[i] = (uint32_t)off64[i];
	off[n] = (uint32_t)used;
	KgBuf d_in, d_off;
	struct G2 { KgBuf *a, *b; ~G2() { a->release(); b->release(); } } g2 = {&d_in, &d_off};
	if (d_in.reserve(used + 64) || d_off.reserve(4 * ((size_t)n + 2))) return -1;
	KG_CUDA(cudaMemcpyAsync(d_in.p, frag_raw, used, cudaMemcpyHostToDevice, st));
	KG_CUDA(cudaMemsetAsync((uint8_t *)d_in.p + used, 0, 64, st));
	KG_CUDA(cudaMemcpyAsync(d_off.p, off.data(), 4 * ((size_t)n + 1), cudaMemcpyHostToDevice, st));
	KG_CUDA(cudaStreamSynchronize(st));
	return conclave_core(db, (const uint8_t *)d_in.p, (const uint32_t *)d_off.p, n, alignment_scores, uniq_alignment_scores, frags_out, out_cap,
	                     out_bytes, w_scores, fragmentCounts, readCounts);
}

// The same on the frag_raw stream the last -mem_mode score collection (kmagpu_memscore_batch / _from_seed) left in HBM.
extern "C" int kmagpu_conclave_resident(kmagpu_db *db, const uint64_t *alignment_scores, const uint64_t *uniq_alignment_scores, void *frags_out,
                                        size_t out_cap, size_t *out_bytes, uint64_t *w_scores, uint32_t *fragmentCounts, uint32_t *readCounts,
                                        int64_t *nrecords) {
	if (!db || (!alignment_scores != !uniq_alignment_scores)) { kmagpu_set_error("null argument"); return -1; }
	if (!db->d_lengths) { kmagpu_set_error("database has no template lengths (.length.b missing)"); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	if (out_bytes) *out_bytes = 0;
	db->frg.valid = false; db->frg.n = 0; db->frg.bytes = 0;
	const RawBatch &r = db->raw;
	if (!r.valid) { kmagpu_set_error("kmagpu_conclave_resident without a preceding score collection on this handle"); return -1; }
	if (nrecords) *nrecords = r.n;
	if (r.n == 0) return conclave_empty(db, frags_out, out_cap, out_bytes);
	return conclave_core(db, (const uint8_t *)r.d_out.p, r.off, (int)r.n, alignment_scores, uniq_alignment_scores, frags_out, out_cap, out_bytes,
	                     w_scores, fragmentCounts, readCounts);
}

// The same on the frag_raw stream the last kmagpu_align_run left in HBM (one slot per stage-2 record: empty, one record,
// or two for a pair that was resolved as two single reads).
extern "C" int kmagpu_conclave_from_align(kmagpu_db *db, const uint64_t *alignment_scores, const uint64_t *uniq_alignment_scores, void *frags_out,
                                          size_t out_cap, size_t *out_bytes, uint64_t *w_scores, uint32_t *fragmentCounts, uint32_t *readCounts,
                                          int64_t *nrecords) {
	if (!db || (!alignment_scores != !uniq_alignment_scores)) { kmagpu_set_error("null argument"); return -1; }
	if (!db->d_lengths) { kmagpu_set_error("database has no template lengths (.length.b missing)"); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	if (out_bytes) *out_bytes = 0;
	db->frg.valid = false; db->frg.n = 0; db->frg.bytes = 0;
	AlignBatch &b = db->aln;
	if (!b.ran) { kmagpu_set_error("kmagpu_conclave_from_align without a preceding kmagpu_align_run on this handle"); return -1; }
	const int n = (int)b.nreads;
	if (nrecords) *nrecords = n;
	if (n == 0 || b.out_bytes == 0) return conclave_empty(db, frags_out, out_cap, out_bytes);
	uint32_t *recoff = (uint32_t *)b.d_recsize.p + n + 1;
	const uint32_t total = (uint32_t)b.out_bytes;
	KG_CUDA(cudaMemcpyAsync(recoff + n, &total, 4, cudaMemcpyHostToDevice, db->stream));
	KG_CUDA(cudaStreamSynchronize(db->stream));
	return conclave_core(db, (const uint8_t *)b.d_out.p, recoff, n, alignment_scores, uniq_alignment_scores, frags_out, out_cap, out_bytes,
	                     w_scores, fragmentCounts, readCounts);
}

// -ConClave 2 (runkma.c:591): the ConClave entry points run runConClave2 / runConClave2_lc (conclave.c:386 / 749) over the
// batch they are given, which then has to be the whole run (its significance filter sums over every read). scoreT / evalue as
// runKMA passes them, and_mode = cmp_and (kma.c:916), p_chisqr = the caller's (stdstat.c:136). version 1 restores runConClave.
extern "C" int kmagpu_conclave_version(kmagpu_db *db, int version, double scoreT, double evalue, int and_mode, double (*p_chisqr)(long double)) {
	if (!db || (version != 1 && version != 2)) { kmagpu_set_error("ConClave version %d: 1 or 2", version); return -1; }
	db->cc2.version = version; db->cc2.scoreT = scoreT; db->cc2.evalue = evalue; db->cc2.and_mode = and_mode != 0; db->cc2.p_chisqr = p_chisqr;
	return 0;
}

// the unique scores as the last ConClave call of this handle left them (runConClave2 adds to them, conclave.c:519)
extern "C" int kmagpu_conclave_uniq_scores(kmagpu_db *db, uint64_t *uniq_alignment_scores) {
	if (!db || !uniq_alignment_scores) { kmagpu_set_error("null argument"); return -1; }
	if (!db->frg.d_sc.p) { kmagpu_set_error("kmagpu_conclave_uniq_scores before a ConClave call on this handle"); return -1; }
	const size_t DB = (size_t)db->info.DB_size;
	KG_CUDA(cudaSetDevice(db->device));
	KG_CUDA(cudaMemcpyAsync(uniq_alignment_scores, (const unsigned long long *)db->frg.d_sc.p + DB, 8 * DB, cudaMemcpyDeviceToHost, db->stream));
	KG_CUDA(cudaStreamSynchronize(db->stream));
	return 0;
}

// which ConClavePtr the three ConClave entry points stand for: 0 = runConClave (conclave.c:43), 1 = runConClave_lc (:215, -lc)
extern "C" int kmagpu_conclave_mode(kmagpu_db *db, int length_corrected) {
	if (!db) { kmagpu_set_error("null argument"); return -1; }
	db->conclave_lc = length_corrected != 0;
	return 0;
}
