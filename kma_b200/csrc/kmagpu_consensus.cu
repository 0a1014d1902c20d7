// Consensus call of the assembly pass on the device: callConsensus (assembly.c:1499-1631) over the template nodes of
// the HBM-resident base-count matrix, with the reference's five base callers (assembly.c:162-271) and its three
// significance tests (assembly.c:141-160).
//
// Every template position is independent (without insertion nodes assembly[pos].next == pos + 1), so the pass is one
// stream over the matrix: 24 bytes of counts in, 3 bytes (t, s, q rows) out per position -- HBM bound. A block takes a
// contiguous run of 256-position tiles; the tiles arrive as 6 KB TMA bulk copies (cp.async.bulk + mbarrier) in a ring of
// four shared-memory stages, three of them in flight while one is being called; one thread calls one position, the rows leave as byte stores of consecutive lanes. depth / depthVar / aln_len / cover of
// a template are summed per warp by shuffles while the warp stays inside one template (lane 0 keeps the running sums
// and flushes them with four atomics when the template changes), so a 5 Mb template costs a few thousand atomics.
//
// The significance test p_chisqr((X-Y)^2 / (X+Y)) <= evalue depends on the counts only through the double
// (X-Y)^2 / (X+Y), which IEEE division reproduces bit for bit on the device; p_chisqr falls with its argument, so the
// decision is `statistic >= chi2_min` with chi2_min located on the HOST by bisection over the caller's own p_chisqr
// (kmagpu_chi2_threshold: host libm, the reference's function when the caller passes it) -- no erf on the device.
#include "kmagpu_internal.h"
#include "kmagpu_dev.cuh"
#include <math.h>
#include <string.h>

#define CS_TILE 256

struct CsParams { int bcd, caller, sig; double support, chi2_min; };
struct CsStat { unsigned long long depth, depthVar; unsigned int len, aln_len, cover, reserved; };

__device__ __forceinline__ int cs_base(int i) { return (int)(__byte_perm(0x54474341u, 0x00002d4eu, (unsigned)i) & 0xffu); }   // "ACGTN-"[i]
__device__ __forceinline__ int cs_lower(int c) { return (c >= 'A' && c <= 'Z') ? c + 32 : c; }

// significantNuc / significantAnd90Nuc / significantAndSupport (assembly.c:141-160). The statistic (X-Y)^2 / (X+Y)
// is compared with chi2_min as the reference's double quotient; the division itself only runs when the product form
// cannot decide with a margin (1e-13 relative) far above the rounding of either side.
template <int SIG> __device__ __forceinline__ bool cs_significant(const CsParams &P, int X, int Y) {
	if (!(Y < X)) return false;
	if (SIG == 1 && !(9 * (X + Y) <= 10 * X)) return false;
	if (SIG == 2 && !(P.support * (double)(X + Y) <= (double)X)) return false;
	const long long d = (long long)X - Y;
	const double dd = (double)(d * d), n = (double)(X + Y), lim = P.chi2_min * n;
	if (dd > lim * (1.0 + 1e-13)) return true;
	if (dd < lim * (1.0 - 1e-13)) return false;
	return dd / n >= P.chi2_min;
}

__device__ __forceinline__ int cs_best_base(const unsigned *c, int *best) {   // first maximum over A C G T N, 0 when all are 0
	int bb = 0, bn = 0;
#pragma unroll
	for (int j = 0; j < 5; ++j) if (bb < (int)c[j]) { bb = (int)c[j]; bn = j; }
	*best = bn;
	return bb;
}

// baseCaller / orgBaseCaller / refCaller / nanoCaller / refNanoCaller (assembly.c:162-271)
template <int CALLER, int SIG>
__device__ __forceinline__ int cs_base_call(const CsParams &P, int bestNuc, int bestScore, int depthUpdate, const unsigned *c) {
	const bool sigf = depthUpdate != 0 && cs_significant<SIG>(P, bestScore, depthUpdate - bestScore);
	int bn;
	switch (CALLER) {
	case 0:
		if (depthUpdate == 0) return '-';
		if (!sigf) return (bestNuc == '-' && bestScore != depthUpdate) ? 'n' : cs_lower(bestNuc);   // tNuc is never '-' on a template node
		return bestNuc;
	case 1:
		if (depthUpdate == 0 || bestNuc == '-') return '-';
		return sigf ? bestNuc : cs_lower(bestNuc);
	case 2:
		if (depthUpdate == 0 || bestNuc == '-') return 'n';
		return sigf ? bestNuc : cs_lower(bestNuc);
	case 3:
		if (depthUpdate == 0) return '-';
		if (!sigf) {
			if (bestNuc == '-' && bestScore != depthUpdate) return cs_best_base(c, &bn) == 0 ? '-' : cs_lower(cs_base(bn));
			return cs_lower(bestNuc);
		}
		return bestNuc;
	default:
		if (depthUpdate == 0) return 'n';
		if (!sigf) {
			if (bestNuc == '-') return cs_best_base(c, &bn) == 0 ? 'n' : cs_lower(cs_base(bn));
			return cs_lower(bestNuc);
		}
		return bestNuc == '-' ? 'n' : bestNuc;
	}
}

__device__ __forceinline__ unsigned long long cs_warp_sum64(unsigned long long v) {
#pragma unroll
	for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

__device__ __forceinline__ void cs_flush(CsStat *stats, int t, unsigned long long depth, unsigned long long var, unsigned aln, unsigned cover) {
	if (t <= 0 || !(depth | var | aln | cover)) return;
	atomicAdd(&stats[t].depth, depth); atomicAdd(&stats[t].depthVar, var);
	atomicAdd(&stats[t].aln_len, aln); atomicAdd(&stats[t].cover, cover);
}

// the warp's running sums into the per-template totals: one reduction when every lane ran inside the same template
__device__ __forceinline__ void cs_fold(CsStat *stats, int t, unsigned long long depth, unsigned long long var, unsigned aln, unsigned cover,
                                        unsigned lane) {
	const int t0 = __shfl_sync(0xffffffffu, t, 0);
	if (__all_sync(0xffffffffu, t == t0)) {
		depth = cs_warp_sum64(depth); var = cs_warp_sum64(var);
		aln = __reduce_add_sync(0xffffffffu, aln); cover = __reduce_add_sync(0xffffffffu, cover);
		if (lane == 0) cs_flush(stats, t0, depth, var, aln, cover);
	} else cs_flush(stats, t, depth, var, aln, cover);
}

// ---- TMA bulk copies (cp.async.bulk, 1-D) completing on an mbarrier
__device__ __forceinline__ unsigned cs_saddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cs_mbar_init(unsigned long long *bar, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cs_saddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void cs_bulk_load(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cs_saddr(bar)), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(cs_saddr(dst)), "l"(src),
	             "r"(bytes), "r"(cs_saddr(bar)) : "memory");
}
__device__ __forceinline__ void cs_mbar_wait(unsigned long long *bar, unsigned parity) {
	asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"(cs_saddr(bar)),
	             "r"(parity) : "memory");
}

#define CS_STAGES 4
#define CS_TILE_BYTES (CS_TILE * 24)

// positions [p0, p1) of the all-template matrix (mat_bytes: size of the whole matrix); tiles are aligned to CS_TILE
// positions of that space so that every tile starts on a 16-byte boundary. A block owns tiles_per_block consecutive
// tiles and keeps CS_STAGES - 1 of them in flight: thread 0 posts one 6 KB bulk copy per tile into a ring of shared-
// memory stages, everybody waits on the stage's mbarrier. rows: t at out_t[pos - p0], likewise s and q.
template <int CALLER, int SIG>
__global__ void __launch_bounds__(CS_TILE) consensus_kernel(CsParams P, const unsigned int *__restrict__ mat, unsigned long long mat_bytes,
		const int64_t *__restrict__ mat_off, int DB_size, const KgTMeta *__restrict__ meta, const uint64_t *__restrict__ seq, long long p0,
		long long p1, long long tiles_per_block, uint8_t *__restrict__ out_t, uint8_t *__restrict__ out_s, uint8_t *__restrict__ out_q,
		CsStat *stats) {
	__shared__ __align__(128) unsigned stage[CS_STAGES][CS_TILE * 6];
	__shared__ __align__(8) unsigned long long bar[CS_STAGES];
	const unsigned lane = threadIdx.x & 31;
	const long long tile0 = p0 / CS_TILE + (long long)blockIdx.x * tiles_per_block;
	long long tile_end = (p1 + CS_TILE - 1) / CS_TILE;
	if (tile_end > tile0 + tiles_per_block) tile_end = tile0 + tiles_per_block;
	const int ntiles = tile_end > tile0 ? (int)(tile_end - tile0) : 0;
	if (threadIdx.x == 0) {
		for (int i = 0; i < CS_STAGES; ++i) cs_mbar_init(&bar[i], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	auto post = [&](int k) {   // thread 0: tile k of this block into stage k % CS_STAGES
		const unsigned long long at = (unsigned long long)(tile0 + k) * CS_TILE_BYTES;
		unsigned long long left = mat_bytes - at;
		const unsigned bytes = left >= CS_TILE_BYTES ? CS_TILE_BYTES : (unsigned)((left + 15) & ~15ull);   // the allocation is padded
		cs_bulk_load(stage[k % CS_STAGES], (const uint8_t *)mat + at, bytes, &bar[k % CS_STAGES]);
	};
	if (threadIdx.x == 0) for (int k = 0; k < CS_STAGES - 1 && k < ntiles; ++k) post(k);
	int run_t = 0;                                            // the template this thread's running sums belong to
	unsigned long long run_depth = 0, run_var = 0;
	unsigned run_aln = 0, run_cover = 0;
	int cur = 0;                                              // template of this thread's previous position and its range
	long long cur_lo = 0, cur_hi = 0, cur_seq = 0;
	for (int k = 0; k < ntiles; ++k) {
		if (threadIdx.x == 0 && k + CS_STAGES - 1 < ntiles) post(k + CS_STAGES - 1);   // its stage was released by the barrier below
		const long long pos = (tile0 + k) * CS_TILE + threadIdx.x;
		const bool in = pos >= p0 && pos < p1;
		int t = 0, tn = 0;
		if (in) {   // template of the position and its base, while the counts are still on their way
			if (pos < cur_lo || pos >= cur_hi) {
				int lo = 1, hi = DB_size - 1;
				while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (mat_off[mid] <= pos) lo = mid; else hi = mid - 1; }
				cur = lo; cur_lo = mat_off[lo]; cur_hi = mat_off[lo + 1]; cur_seq = meta[lo].seq_off;
			}
			t = cur;
			const int tp = (int)(pos - cur_lo);
			tn = (int)((__ldg(seq + cur_seq + (tp >> 5)) << ((tp & 31) << 1)) >> 62);
		}
		cs_mbar_wait(&bar[k % CS_STAGES], (unsigned)(k / CS_STAGES) & 1u);
		// per-template sums run in registers; a warp folds them (shuffles, four atomics) when it leaves a template
		if (__any_sync(0xffffffffu, in && t != run_t)) {
			cs_fold(stats, run_t, run_depth, run_var, run_aln, run_cover, lane);
			run_t = t; run_depth = run_var = 0; run_aln = run_cover = 0;
		}
		if (in) {
			const unsigned *tile = stage[k % CS_STAGES];
			unsigned c[6];
			{
				const uint2 a = *(const uint2 *)(tile + 6 * threadIdx.x), b = *(const uint2 *)(tile + 6 * threadIdx.x + 2),
				            d = *(const uint2 *)(tile + 6 * threadIdx.x + 4);
				c[0] = min(a.x, 65535u); c[1] = min(a.y, 65535u); c[2] = min(b.x, 65535u);   // the reference's uint16 saturation (assembly.c:1436)
				c[3] = min(b.y, 65535u); c[4] = min(d.x, 65535u); c[5] = min(d.y, 65535u);
			}
			const int ct = tn == 0 ? (int)c[0] : tn == 1 ? (int)c[1] : tn == 2 ? (int)c[2] : (int)c[3];
			int bestNuc = tn, bestScore = ct, depthUpdate = 0;
#pragma unroll
			for (int j = 0; j < 6; ++j) {
				if (bestScore < (int)c[j]) { bestScore = (int)c[j]; bestNuc = j; }
				depthUpdate += (int)c[j];
			}
			int call = cs_base(bestNuc);
			if (!depthUpdate) call = '-';
			else if ((bestScore << 1) < depthUpdate) {            // minor base call (assembly.c:1563-1579)
				if (call == '-') {
					int bb = (int)c[4], bn = 4;
#pragma unroll
					for (int j = 0; j < 4; ++j) if (bb < (int)c[j]) { bb = (int)c[j]; bn = j; }
					call = cs_lower(cs_base(bn));
				} else call = cs_lower(call);
				bestScore = depthUpdate - (int)c[5];
			} else if (depthUpdate < P.bcd) call = cs_lower(call);
			call = cs_base_call<CALLER, SIG>(P, call, bestScore, depthUpdate, c);
			int sc = '_';
			if (call != '-') {
				run_depth += (unsigned long long)depthUpdate; run_var += (unsigned long long)depthUpdate * (unsigned long long)depthUpdate; ++run_aln;
				if (cs_base(tn) == (call >= 'a' ? call - 32 : call)) { ++run_cover; sc = '|'; }
			}
			out_t[pos - p0] = (uint8_t)cs_base(tn); out_s[pos - p0] = (uint8_t)sc; out_q[pos - p0] = (uint8_t)call;
		}
		__syncthreads();   // every thread has read its counts: the stage may be refilled
	}
	cs_fold(stats, run_t, run_depth, run_var, run_aln, run_cover, lane);
}

// ---------------------------------------------------------------- host side

static double cs_builtin_p(long double q) {   // stdstat.c:136-147 up to 49; above it the reference reads a table of values <= 1e-11
	if (q < 0) return 1e-26;
	if (q > 49) return 1e-11;
	return 1 - 1.772453850 * erf(sqrt((double)(0.5L * q))) / tgamma(0.5);
}

extern "C" double kmagpu_chi2_threshold(double evalue, double (*p_chisqr)(long double)) {
	const bool own = p_chisqr == nullptr;
	if (own) {
		if (evalue < 1e-11) { kmagpu_set_error("evalue %g lies in the range of the reference's p-value table: pass its p_chisqr", evalue); return -1.0; }
		p_chisqr = &cs_builtin_p;
	}
	if (p_chisqr(0.0L) <= evalue) return 0.0;
	union { double d; uint64_t u; } lo, hi, mid;
	lo.d = 0.0; hi.d = own ? 49.0 : 256.0;
	if (!(p_chisqr(hi.d) <= evalue)) { kmagpu_set_error("no statistic up to %g reaches evalue %g", hi.d, evalue); return -1.0; }
	while (hi.u - lo.u > 1) {   // doubles of one sign order like their bit patterns
		mid.u = lo.u + ((hi.u - lo.u) >> 1);
		if (p_chisqr(mid.d) <= evalue) hi = mid; else lo = mid;
	}
	return hi.d;
}

extern "C" int kmagpu_consensus(kmagpu_db *db, int32_t tmpl, const kmagpu_consensus_params *cp, uint8_t *t, uint8_t *s, uint8_t *q,
                                size_t cap, kmagpu_consensus_stats *stats, float *ms) {
	if (!db || !cp) { kmagpu_set_error("null argument"); return -1; }
	if (!db->image->d_mat) { kmagpu_set_error("kmagpu_consensus before any alignment was added to the matrix"); return -1; }
	if (cp->caller < 0 || cp->caller > 4 || cp->significance < 0 || cp->significance > 2) { kmagpu_set_error("unknown base caller / significance test"); return -1; }
	if (!(cp->chi2_min >= 0.0)) { kmagpu_set_error("chi2_min must come from kmagpu_chi2_threshold"); return -1; }
	static_assert(sizeof(CsStat) == sizeof(kmagpu_consensus_stats), "stats layout");
	KG_CUDA(cudaSetDevice(db->device));
	const int DB = db->info.DB_size;
	if (tmpl < 0 || tmpl >= DB) { kmagpu_set_error("template %d outside the database", tmpl); return -1; }
	long long p0 = 0, p1 = (long long)(db->image->mat_entries / 6);
	if (tmpl) {
		for (int i = 1; i < tmpl; ++i) p0 += db->lengths[i];
		p1 = p0 + db->lengths[tmpl];
	}
	const size_t n = (size_t)(p1 - p0);
	if ((t || s || q) && n > cap) { kmagpu_set_error("consensus rows need %zu bytes each, caller gave %zu", n, cap); return -1; }
	KgBuf &rows = db->d_cons_rows, &dstat = db->d_cons_stat;
	if (rows.reserve(3 * n + 64) || dstat.reserve(sizeof(CsStat) * (size_t)DB)) return -1;
	cudaStream_t st = db->stream;
	KG_CUDA(cudaMemsetAsync(dstat.p, 0, sizeof(CsStat) * (size_t)DB, st));
	CsParams P = {cp->bcd, cp->caller, cp->significance, cp->support, cp->chi2_min};
	const long long tiles = (p1 + CS_TILE - 1) / CS_TILE - p0 / CS_TILE;
	typedef void (*kern_t)(CsParams, const unsigned int *, unsigned long long, const int64_t *, int, const KgTMeta *, const uint64_t *, long long,
	                       long long, long long, uint8_t *, uint8_t *, uint8_t *, CsStat *);
	static const kern_t kerns[5][3] = {
		{consensus_kernel<0, 0>, consensus_kernel<0, 1>, consensus_kernel<0, 2>}, {consensus_kernel<1, 0>, consensus_kernel<1, 1>, consensus_kernel<1, 2>},
		{consensus_kernel<2, 0>, consensus_kernel<2, 1>, consensus_kernel<2, 2>}, {consensus_kernel<3, 0>, consensus_kernel<3, 1>, consensus_kernel<3, 2>},
		{consensus_kernel<4, 0>, consensus_kernel<4, 1>, consensus_kernel<4, 2>}};
	const kern_t kern = kerns[cp->caller][cp->significance];
	int per_sm = 0;   // one wave of resident blocks, each with a contiguous run of tiles
	KG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CS_TILE, 0));
	long long grid = (long long)db->sm_count * (per_sm > 0 ? per_sm : 1);
	if (grid > tiles) grid = tiles > 0 ? tiles : 1;
	const long long per = (tiles + grid - 1) / grid;
	grid = per ? (tiles + per - 1) / per : 1;
	if (grid < 1) grid = 1;
	KG_CUDA(cudaEventRecord(db->ev[2], st));
	if (n) kern<<<(unsigned)grid, CS_TILE, 0, st>>>(P, db->image->d_mat, 4ull * db->image->mat_entries, db->image->d_mat_off, DB, db->tix.meta, db->tix.seq, p0, p1, per,
		(uint8_t *)rows.p, (uint8_t *)rows.p + n, (uint8_t *)rows.p + 2 * n, (CsStat *)dstat.p);
	KG_CUDA(cudaEventRecord(db->ev[3], st));
	if (t) KG_CUDA(cudaMemcpyAsync(t, rows.p, n, cudaMemcpyDeviceToHost, st));
	if (s) KG_CUDA(cudaMemcpyAsync(s, (uint8_t *)rows.p + n, n, cudaMemcpyDeviceToHost, st));
	if (q) KG_CUDA(cudaMemcpyAsync(q, (uint8_t *)rows.p + 2 * n, n, cudaMemcpyDeviceToHost, st));
	if (stats) {
		if (tmpl) KG_CUDA(cudaMemcpyAsync(stats, (CsStat *)dstat.p + tmpl, sizeof(CsStat), cudaMemcpyDeviceToHost, st));
		else KG_CUDA(cudaMemcpyAsync(stats, dstat.p, sizeof(CsStat) * (size_t)DB, cudaMemcpyDeviceToHost, st));
	}
	cudaError_t e = cudaStreamSynchronize(st);
	if (e == cudaSuccess) e = cudaGetLastError();
	if (e != cudaSuccess) { kmagpu_set_error("consensus: %s", cudaGetErrorString(e)); return -1; }
	if (stats) {   // aligned_assem->len = asm_len (assembly.c:1625)
		if (tmpl) stats[0].len = (uint32_t)db->lengths[tmpl];
		else for (int i = 1; i < DB; ++i) stats[i].len = (uint32_t)db->lengths[i];
	}
	if (ms) cudaEventElapsedTime(ms, db->ev[2], db->ev[3]);
	return 0;
}
