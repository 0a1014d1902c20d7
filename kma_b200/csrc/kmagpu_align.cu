// The alignment pass of KMA on the GPU: stage-2 records in, frag_raw records + ConClave score sums out.
//
// WHAT: alnFrags_threaded / alnFragsSE (alnfrags.c:1052-1218, 2150-2294): for every read and every candidate
// template, (anker_rc_comp when the strand is undecided, align.c:993) -> KMA_score (align.c:509: MEM seeds against
// the template's position index, chainSeeds chain.c:79, lead/trail tail and gap Needleman-Wunsch nw.c:642/892),
// then the per-read selection and update_Scores (updatescores.c:203-298).
//
// HOW (ours):
//   * the batch stays in HBM; a prep kernel unpacks every read once into an aligned slab (2-bit words, 0-4 bytes,
//     N list; reverse complement only for strand-tie reads);
//   * the unit of work is a (read, template) pair, one warp each, pulled from an atomic counter by a persistent
//     grid -- candidate counts and NW sizes are irregular, pairs are not;
//   * MEM discovery: the 32 lanes probe 32 consecutive query positions at once, the first hit (ballot) is extended
//     on the packed words with XOR + clz/ffs, 32 bases per step; repeated k-mers extend one occurrence per lane;
//   * chaining: the 128-MEM look-ahead of chainSeeds is evaluated by the lanes in parallel and folded with the
//     reference's <= / < tie rules (first maximum, last fully-compatible maximum);
//   * NW: the continuous warp wavefront of kmagpu_nw.cuh;
//   * pairs whose MEM list or traceback matrix exceeds the per-warp scratch are re-run by the same code on a small
//     grid with scratch sized from what they asked for -- results never depend on the scratch size;
//   * selection, the u64 ConClave sums (atomics) and the frag_raw byte stream are produced on the device in input
//     order (size scan + writer kernel), so the host does no per-read work.
#include "kmagpu_internal.h"
#include "kmagpu_dev.cuh"
#include "kmagpu_nw.cuh"
#include <string.h>
#include <algorithm>
#ifndef KG_PAIR_VARIANT_ONLY
#include <cub/device/device_radix_sort.cuh>
#endif

#define AL_WARPS 4            // warps per CTA of the pair kernel
#ifndef KG_STAGE_INL
#define KG_STAGE_INL                // inlined per call site: 50 ms vs 71 ms out of line (arguments by reference go through the stack)
#endif
#define AL_BANDW 64           // align.c:511
#ifndef AL_CLAIM
#define AL_CLAIM 4            // tasks a warp claims from the work counter at a time
#endif
#ifndef AL_MINB
#define AL_MINB 6             // resident CTAs per SM the pair kernel is compiled for (register cap 65536 / (128 * AL_MINB))
#endif
// The pair kernel exists twice: short reads (little state per task) run faster with more, register-poorer warps, long
// reads (C3) with fewer, register-richer ones (profiles/r01_ab_minb.log: 10 CTAs/SM -8.6 % on C2, +39 % on C3).
#ifndef AL_MINB_SHORT
#define AL_MINB_SHORT 10
#endif
#define AL_SHORT_MAXQ 512     // batches whose longest read is at most this use the short-read variant
#define ST_OK 0
#define ST_OVERFLOW 1
#define ST_GIVEUP 4           // KMA_score gave the chain up (align.c:715): queued gap / trail problems of the task are void

struct AlnRead {
	uint32_t rec_off;
	int32_t q_len, words, nN, rc_flag, nt, hl, flag;
	uint32_t slab_off;   // 8-byte units
	uint32_t task0;
	int32_t kind;        // 0 single read, 1 first record of a pair (no templates, ankers.c:150), 2 its mate (carries the templates)
	int32_t fneg;        // kind 2: index of the first negative template (nt when none): both reads flip there (alnfrags.c:1630)
	int32_t two;         // the slab also holds the reverse complement (strand-tie reads, pairs with a negative template)
	int32_t q_start, q_end;   // query bounds of a chain-mode record (qseqs.c:41, alnfrags.c:1091-1099); else 0, q_len
};

struct AlnCand { int32_t tmpl, score, len, pos, match, tGaps, qGaps, status; };

struct AlnParams {
	NwPen pen;
	int32_t k, mq, one2one, exhaustive, minlen, Wl, PE, apm, ts, pad0;
	double scoreT, mrc, minFrac;
};

enum { A_WORK = 0, A_OVF = 1, A_NEED_E = 2, A_NEED_MEM = 3, A_NEED_Q = 4, A_MEMS = 5, A_FULL_CALLS = 6, A_BAND_CALLS = 7,
       A_FULL_CELLS = 8, A_BAND_CELLS = 9, A_STEPS = 10, A_SLAB = 11, A_TASKS = 12, A_OUT = 13, A_FRAGS = 14, A_BAD = 15,
       A_MAXQ = 16, A_LOOKUPS = 17, A_MEMBASES = 18, A_READBYTES = 19, A_NPROB = 20, A_PCLS = 21 /* .. 24 */, A_N = 32 };

// ---------------------------------------------------------------- slab layout

struct QView { const uint64_t *w; const uint8_t *b; const int32_t *N; };

__host__ __device__ __forceinline__ uint32_t slab_W(int words) { return (uint32_t)words + 2; }
__host__ __device__ __forceinline__ uint32_t slab_B(int q_len) { return ((uint32_t)q_len + 8) >> 3; }
__host__ __device__ __forceinline__ uint32_t slab_N(int nN) { return ((uint32_t)nN + 2) >> 1; }
__host__ __device__ __forceinline__ uint32_t slab_stride(const AlnRead &R) { return slab_W(R.words) + slab_B(R.q_len) + slab_N(R.nN); }

__device__ __forceinline__ QView read_view(const uint64_t *slab, const AlnRead &R, int strand) {
	const uint64_t *base = slab + R.slab_off + (strand ? slab_stride(R) : 0);
	QView v;
	v.w = base;
	v.b = (const uint8_t *)(base + slab_W(R.words));
	v.N = (const int32_t *)(base + slab_W(R.words) + slab_B(R.q_len));
	return v;
}

// ---------------------------------------------------------------- pass 1: sizes

static __global__ void aln_sizes_kernel(const uint8_t *__restrict__ in, const uint32_t *__restrict__ off, int n, int k,
                                 AlnRead *reads, uint32_t *slab_sz, uint32_t *task_sz, unsigned long long *ctr) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n) return;
	AlnRead R;
	memset(&R, 0, sizeof(R));
	R.rec_off = off[r];
	uint32_t ss = 0, ts = 0;
	if (off[r + 1] - off[r] >= 28) {
		const uint8_t *rec = in + R.rec_off;
		R.q_len = (int)ld_u32u(rec); R.words = (int)ld_u32u(rec + 4); R.nN = (int)ld_u32u(rec + 8);
		R.rc_flag = (int)ld_u32u(rec + 12); R.nt = (int)ld_u32u(rec + 16); R.hl = (int)ld_u32u(rec + 20);
		R.flag = (int)ld_u32u(rec + 24);
		int prev_nt = 1, prev_len = 0, prev_rc = 0;
		if (r > 0 && off[r] - off[r - 1] >= 28) {
			prev_nt = (int)ld_u32u(in + off[r - 1] + 16); prev_len = (int)ld_u32u(in + off[r - 1]); prev_rc = (int)ld_u32u(in + off[r - 1] + 12);
		}
		R.kind = R.nt == 0 ? 1 : (prev_nt == 0 ? 2 : 0);
		R.fneg = R.nt;
		R.q_start = 0; R.q_end = R.q_len;
		if (9 < R.hl) {
			const uint8_t *he = rec + 28 + 8 * (size_t)R.words + 4 * (size_t)R.nN + 4 * (size_t)R.nt + (size_t)R.hl;
			if (he[-9] == 0) { R.q_start = (int)ld_u32u(he - 8); R.q_end = (int)ld_u32u(he - 4); }
		}
		R.two = R.rc_flag < 0;
		if (R.kind) {   // a pair is reverse-complemented from its first negative template on: look for one
			const bool mate_ok = R.kind == 2 || (r + 1 < n && off[r + 2] - off[r + 1] >= 28);
			const uint8_t *trec = R.kind == 2 ? rec : in + off[mate_ok ? r + 1 : r];
			int nt2 = R.nt;
			if (R.kind == 1) nt2 = mate_ok ? (int)ld_u32u(trec + 16) : 0;
			const uint8_t *T = trec + 28 + 8 * (size_t)ld_u32u(trec + 4) + 4 * (size_t)ld_u32u(trec + 8);
			int f = nt2;
			for (int i = 0; i < nt2; ++i) if ((int)ld_u32u(T + 4 * (size_t)i) < 0) { f = i; break; }
			if (R.kind == 2) R.fneg = f;
			R.two |= f < nt2;
		}
		ss = slab_stride(R) * (R.two ? 2u : 1u);
		if (R.kind == 2) {
			// alnFrags_threaded (alnfrags.c:2250): both mates reach k or the pair degenerates to forms stage 2 never writes
			if (prev_len < k || R.q_len < k || prev_rc < 0) atomicAdd(&ctr[A_BAD], 1ull);
			else ts = 2u * (uint32_t)R.nt;
		} else if (R.kind == 0) ts = R.q_len >= k ? (uint32_t)R.nt : 0u;
		warp_max_u64(&ctr[A_MAXQ], (unsigned)R.q_len);
	}
	reads[r] = R;
	slab_sz[r] = ss; task_sz[r] = ts;
}

// ---------------------------------------------------------------- pass 2: unpack (one warp per read)

__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
	for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}

#ifndef ST_MINB
#define ST_MINB 8   // resident CTAs per SM the record-streaming kernels (prep / emit) are compiled for: latency-bound, 32 registers and all 64 warps (C2: 9.5 -> 8.45 ms, profiles/r02_stream_kernels.log)
#endif
static __global__ void __launch_bounds__(256, ST_MINB) aln_prep_kernel(const uint8_t *__restrict__ in, int n, AlnRead *reads,
		const uint32_t *__restrict__ slab_off, const uint32_t *__restrict__ task_off, uint64_t *slab, int32_t *task_read, int k) {
	const unsigned lane = threadIdx.x & 31;
	const int warps = (gridDim.x * blockDim.x) >> 5;
	for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
		AlnRead R = reads[r];
		R.slab_off = slab_off[r]; R.task0 = task_off[r];
		if (lane == 0) { reads[r].slab_off = R.slab_off; reads[r].task0 = R.task0; }
		if (R.q_len == 0 && R.words == 0) continue;
		const uint8_t *rec = in + R.rec_off, *seq = rec + 28, *Ns = seq + 8 * (size_t)R.words;
		const int L = R.q_len, words = R.words, nN = R.nN;
		for (int strand = 0; strand < (R.two ? 2 : 1); ++strand) {
			uint64_t *base = slab + R.slab_off + (strand ? slab_stride(R) : 0);
			uint8_t *b = (uint8_t *)(base + slab_W(words));
			int32_t *N = (int32_t *)(base + slab_W(words) + slab_B(L));
			// one round per packed word: every lane reads the word (one broadcast load from the read-only record), lane 0
			// stores it, lane i takes base 32 w + i of it as a byte 0-3 (unCompDNA, compdna.c:178). The rounds are
			// independent: unrolled, their loads are in flight together
			const int nbytes = (int)(slab_B(L) << 3);
#pragma unroll 4
			for (int w = 0; w < words + 2; ++w) {
				uint64_t x = 0;
				if (w < words) {
					if (!strand) x = ld_u64u(seq + 8 * (size_t)w);
					else {   // rc_comp (compdna.c:228)
						x = rev2(~fwd32(seq, words, L - 32 * (w + 1)));
						const int c = L - 32 * w;
						if (c < 32) x = c > 0 ? x & (~0ull << (64 - 2 * c)) : 0ull;
					}
				}
				if (lane == 0) base[w] = x;
				const int i = 32 * w + (int)lane;
				if (i < nbytes) b[i] = i < L ? (uint8_t)((x << (lane << 1)) >> 62) : (uint8_t)0;
			}
			__syncwarp();
#pragma unroll 1
			for (int i = lane; i <= nN; i += 32) {
				int v = L;   // sentinel N[nN] = q_len (savekmers.c:2483, alnfrags.c:1071)
				if (i < nN) v = strand ? L - 1 - (int)ld_u32u(Ns + 4 * (size_t)(nN - 1 - i)) : (int)ld_u32u(Ns + 4 * (size_t)i);
				N[i] = v;
				if (i < nN) b[v] = 4;
			}
		}
		const int ntask = (int)(task_off[r + 1] - task_off[r]);
#pragma unroll 1
		for (int i = lane; i < ntask; i += 32) task_read[R.task0 + i] = r;
		__syncwarp();
	}
}

// ---------------------------------------------------------------- MEMs

// One 32-byte record per MEM (tS, tE, qS, qE, W, sc, nx, pad): a MEM is one sector, and the table is one pointer in
// registers instead of seven (the pair kernel runs under a 48-register cap).
struct Mems {
	int *base;
	int cap;
	__device__ __forceinline__ int &tS(int i) const { return base[8 * i]; }
	__device__ __forceinline__ int &tE(int i) const { return base[8 * i + 1]; }
	__device__ __forceinline__ int &qS(int i) const { return base[8 * i + 2]; }
	__device__ __forceinline__ int &qE(int i) const { return base[8 * i + 3]; }
	__device__ __forceinline__ int &W(int i) const { return base[8 * i + 4]; }
	__device__ __forceinline__ int &sc(int i) const { return base[8 * i + 5]; }
	__device__ __forceinline__ int &nx(int i) const { return base[8 * i + 6]; }
	__device__ __forceinline__ int4 pos(int i) const { return *(const int4 *)(base + 8 * i); }   // tS, tE, qS, qE
	__device__ __forceinline__ void set(int i, int ts, int te, int qs, int qe) const {
		*(int4 *)(base + 8 * i) = make_int4(ts, te, qs, qe);
		base[8 * i + 4] = qe - qs;
	}
	__device__ __forceinline__ void shift(int o) { base += 8 * o; cap -= o; }
};

// KG_STAT: the statistic counters of the alignment kernels (algorithmic-byte inputs of the bench); -DKG_NO_STATS compiles them out
#ifdef KG_NO_STATS
#define KG_STAT(x)
#else
#define KG_STAT(x) x
#endif
struct WarpCtr { unsigned long long full_calls, band_calls, full_cells, band_cells, steps, mems, lookups, mem_bases, read_bytes; unsigned need_e, need_mem, need_q; };

__device__ __noinline__ int warp_max(int v) {
#pragma unroll
	for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}
__device__ __noinline__ int warp_min(int v) {
#pragma unroll
	for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}
__device__ __noinline__ int warp_sum(int v) {
#pragma unroll
	for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// one exact seed at (query p, template 1-based v) extended to a maximal exact match on the packed words
__device__ __forceinline__ void mem_from_seed(const uint64_t *qw, const uint64_t *tseq, int t_len, int k, int p, int v, int lo,
                                              int fwd_lim, int *qs, int *ts, int *qe, int *te) {
	const int t0 = v - 1;
	const int bw = ext_bwd(qw, p, tseq, t0, min(p - lo, t0));
	int fmax = min(fwd_lim - (p + k), t_len - (t0 + k));
	const int fw = fmax > 0 ? ext_fwd(qw, p + k, tseq, t0 + k, fmax) : 0;
	*qs = p - bw; *ts = v - bw; *qe = p + k + fw; *te = v + k + fw;
}

// MODE 0: the seed scan of KMA_score (align.c:534-640); MODE 1: one strand of anker_rc_comp (align.c:1044-1143).
// BYTES: the byte-read flavour of KMA (align.c:246-377) and anker_rc (align.c:823-957): a stretch between N's is only
// entered, and re-entered after a MEM, while MORE than k bases remain before its end; MEMs always extend to the
// stretch end. Appends to M starting at index n; *sscore = strand score of MODE 1. Returns ST_OVERFLOW when M is full.
template <int MODE, bool BYTES>
// Query bounds [start, q_end) (chain-mode records): KMA_score starts its first stretch at `start` and clips only the
// last stretch to q_end (align.c:535-541); anker_rc_comp stops entering stretches at q_end (align.c:1044).
__device__ KG_STAGE_INL int scan_mems(const KgTIndexView &ix, const KgTMeta &m, const uint64_t *tseq, const QView &q, int nN1, int q_len,
                         int start, int q_end, Mems &M, int &n, int &sscore, WarpCtr &wc) {
	const int lane = threadIdx.x & 31;
	const int k = ix.k, t_len = m.len;
	int j = start, s = 0;
	for (int seg = 0; seg < nN1 && j < ((MODE == 1 || BYTES) ? q_end : q_len); ++seg) {
		const int realN = q.N[seg];
		// byte flavour (align.c:250-254, 823-827): a stretch runs from the cursor to the next N at or after it (N's before
		// the cursor are never looked at), or to q_end when there is none
		if (BYTES && realN < j) continue;
		const int segN = ((MODE == 0 || BYTES) && seg == nN1 - 1) ? q_end : realN, end = segN - k + 1, lo = seg ? q.N[seg - 1] + 1 : 0;
		const int fwd_lim = (BYTES || MODE == 0) ? segN : end;
		if (BYTES && !(j < segN - k)) { j = segN + 1; continue; }
		while (j < end) {
			const int p0 = j + lane;
			int val = 0;
			if (p0 < end) val = tix_get(ix, m, kmer_at(q.w, p0, k));
			const unsigned hits = __ballot_sync(0xffffffffu, val != 0);
			if (!hits) { KG_STAT(wc.lookups += (unsigned long long)min(32, end - j);) j += 32; continue; }
			const int f = __ffs(hits) - 1, p = j + f;
			KG_STAT(wc.lookups += (unsigned long long)(f + 1);)   // the probes the reference's sequential scan makes
			const int v = __shfl_sync(0xffffffffu, val, f);
			if (v > 0) {
				if (n >= M.cap) { wc.need_mem = max(wc.need_mem, (unsigned)n + 1024u); return ST_OVERFLOW; }
				int qs, ts, qe, te;
				mem_from_seed(q.w, tseq, t_len, k, p, v, lo, fwd_lim, &qs, &ts, &qe, &te);
				if (lane == 0) M.set(n, ts, te, qs, qe);
				++n;
				s += qe - qs;
				KG_STAT(wc.mem_bases += (unsigned long long)(qe - qs);)
				j = (BYTES || MODE == 0) ? qe : qe + 1;
			} else {
				int cnt;
				const int32_t *d = tix_dups(ix, v, &cnt);
				if (n + cnt > M.cap) { wc.need_mem = max(wc.need_mem, (unsigned)(n + cnt) + 1024u); return ST_OVERFLOW; }
				int bias = p;
#pragma unroll 1
				for (int c = lane; c < cnt; c += 32) {   // every occurrence, ascending template position
					int qs, ts, qe, te;
					mem_from_seed(q.w, tseq, t_len, k, p, __ldg(d + c), lo, fwd_lim, &qs, &ts, &qe, &te);
					M.set(n + c, ts, te, qs, qe);
					bias = max(bias, qe);
				}
				bias = warp_max(bias);
				n += cnt;
				KG_STAT(wc.mem_bases += (unsigned long long)cnt * (unsigned long long)k;)
				s += k + (bias - p);
				j = bias + 1;
			}
			if (BYTES && !(j < segN - k)) break;   // "update position" (align.c:309-315): the stretch is left
		}
		j = realN + 1;
	}
	sscore = s;
	__syncwarp();
	return ST_OK;
}

// ---------------------------------------------------------------- chaining (chainSeeds, chain.c:79-260)

// cold paths kept out of line so that the chaining loop stays short in the instruction cache
__device__ __noinline__ int cold_div(int a, int b) { return a / b; }
__device__ __noinline__ unsigned chain_mapq(int bestScore, int secondScore, int w) {   // chain.c:256
	const double wgt = w / 10.0;
	return (unsigned)ceil(40 * (1 - 1.0 * secondScore / bestScore) * (wgt < 1 ? wgt : 1.0) * log((double)bestScore));
}

// kr = 0xFFFFFFFF / k + 1: __umulhi(Ms, kr) == Ms / k exactly for Ms < 2^32 / k (gaps are bounded by the read length)
__device__ __forceinline__ int tail_mm(int Ms, int k, unsigned kr, int Mv, int MMv) {   // mismatch estimate of Ms unaligned bases
	int MMs;
	if (Ms == 2) { MMs = 2; Ms = 0; }
	else {
		const int q = (unsigned)Ms < (1u << 26) ? (int)__umulhi((unsigned)Ms, kr) : cold_div(Ms, k);
		MMs = q + (Ms != q * k ? 1 : 0); MMs = max(2, MMs);
		Ms = min(Ms - MMs, k); Ms = min(Ms, MMs);
	}
	return Ms * Mv + MMs * MMv;
}

#define NOCAND (-0x7fffffff - 1)

// mapQ is only computed when the caller compares it (-mq != 0): its double-precision log is ~120 warp instructions per chain
__device__ int chain_warp(const NwPen &pen, Mems &M, int n, int q_len, int t_len, int k, unsigned *mapQ, bool want_mapq) {
	const int lane = threadIdx.x & 31;
	const unsigned kr = 0xFFFFFFFFu / (unsigned)k + 1u;
	const int W1 = pen.W1, U = pen.U, Mv = pen.M, MMv = pen.MM;
	int bestPos = n - 1, bestScore = 0, secondScore = 0;
	if (lane == 0) { M.sc(n) = 0; M.nx(n) = 0; }
	__syncwarp();
	for (int i = n - 1; i >= 0; --i) {
		const int weight = M.W(i) * Mv, tEnd = M.tE(i), qEnd = M.qE(i);
		int gap = min(t_len - tEnd, q_len - qEnd), Ms = gap;
		if (--gap) gap = gap * U + W1; else gap = W1;
		Ms = tail_mm(Ms, k, kr, Mv, MMv);
		const int score0 = weight + (Ms < gap ? gap : Ms);
		const int lim = min(n, i + 128);
		// lane-local fold over j = i+1+lane, +32, ...: best value, first j reaching it, last "<=" j reaching it
		int lb = NOCAND, lfirst = 0x7fffffff, llast = -1;
#pragma unroll 1
		for (int j = i + 1 + lane; j < lim; j += 32) {
			const int4 pj = M.pos(j);
			const int qSj = pj.z, tSj = pj.x;
			int g = 0, type = -1;   // 0: fully compatible (<=), 1: overlap cut (<)
			if (qEnd < qSj) {
				if (tEnd < tSj) {
					const int tGap = tSj - tEnd, qGap = qSj - qEnd;
					if ((g = abs(tGap - qGap))) g = (g - 1) * U + W1;
					g += weight + M.sc(j) + tail_mm(min(tGap, qGap), k, kr, Mv, MMv);
					type = 0;
				} else if (k <= pj.y - tEnd) {
					if ((g = qSj - qEnd)) g = (g - 1) * U + W1;
					g += weight + M.sc(j) - (tSj - tEnd) * Mv;
					type = 1;
				}
			} else if (k <= pj.w - qEnd) {
				const int tStart = tSj + qEnd - qSj;
				if (tEnd < tStart) {
					if ((g = tStart - tEnd)) g = (g - 1) * U + W1;
					g += weight + M.sc(j) - (tStart - tEnd) * Mv;
					type = 1;
				}
			}
			if (type >= 0) {
				if (lb == NOCAND || g > lb) { lb = g; lfirst = j; llast = type == 0 ? j : -1; }
				else if (g == lb && type == 0) llast = j;
			}
		}
		const int mx = warp_max(lb);
		int score = score0, next = 0;
		if (mx != NOCAND && mx >= score0) {
			const int first = warp_min(lb == mx ? lfirst : 0x7fffffff);
			const int last = warp_max(lb == mx ? llast : -1);
			if (mx > score0) { score = mx; next = max(first, last); }
			else if (last >= 0) next = last;
		}
		int w = M.W(i);
		if (next) w += M.W(next) - k + 1; else w -= k - 1;
		gap = min(M.tS(i), M.qS(i)); Ms = gap;
		if (0 < --gap) gap = gap * U + W1; else if (gap == 0) gap = W1; else gap = 0;
		Ms = tail_mm(Ms, k, kr, Mv, MMv);
		__syncwarp();
		if (lane == 0) { M.W(i) = w; M.sc(i) = score; M.nx(i) = next; }
		__syncwarp();
		score += Ms < gap ? gap : Ms;
		if (bestScore <= score) {
			if (next != bestPos) secondScore = bestScore;
			bestScore = score; bestPos = i;
		} else if (secondScore <= score && next != bestPos) secondScore = bestScore;
	}
	if (want_mapq && 0 < bestScore) *mapQ = chain_mapq(bestScore, secondScore, M.W(bestPos));   // only compared with -mq, default 0
	else *mapQ = 0;
	__syncwarp();
	if (lane == 0) M.sc(bestPos) = bestScore;
	__syncwarp();
	return bestPos;
}

// ---------------------------------------------------------------- stitching (KMA_score, align.c:641-748)

// ---- NW problem queue. The alignment pass is split in phases: the pair kernel finds MEMs, chains them and, instead of
// running Needleman-Wunsch itself, appends every tail / gap problem KMA_score would hand to NW_score / NW_band_score
// (align.c:92-98, 186-192, 478-484) to a queue; queue kernels then solve the problems -- one THREAD per problem for the
// small ones (classes 0-3 by size, so that a warp's 32 problems are alike), one WARP per problem for the rest -- and add
// their AlnScore fields into the task's candidate row (all of them are sums; the lead tail also moves `pos`).
#define NWQ_CLASSES 4
#define NWQ_NARROW 4   // cells per lane of the widest row the narrow warp kernel sweeps
struct NwProb { int32_t task, tmpl, t_s, t_e, q_s, q_e, kband; uint32_t qoff; };   // kband = (k + 2) | band << 8; query bytes at qbase + (qoff << qshift)
// order / cells: per class `cap` queue slots and their cell counts in arrival order; the thread classes are then sorted by
// cells so that the 32 problems of a warp are alike
struct NwQueue { NwProb *probs; uint32_t *order, *cells; unsigned cap; unsigned long long *ctr; int d8; };

// 0, 1: one thread per problem (small full matrices); 2: everything else -- bands, wide matrices -- one warp per problem
// (nw_warp: row sweep for rows of up to 256 cells, the continuous wavefront beyond). A third scheme for the bands --
// column blocks in registers, 8 lanes per problem, three shuffles per step -- was built, verified byte for byte and
// measured at 143 GCUPS against the row sweep's 213 (profiles/r02_ab_column_blocks.log: ~77 instructions per cell as
// compiled, 168 registers), so the row sweep stays.
__host__ __device__ __forceinline__ int nwq_class(int t_l, int q_l, int band, int d8) {
	if (band || q_l > 64 || t_l > 128 || t_l <= 0 || q_l <= 0) {
		// warp per problem; rows of up to NWQ_NARROW * 32 cells (every default band) go to the build of the kernel that
		// sweeps nothing wider: 166 registers for the 256-cell sweep were 3 resident CTAs per SM for all of them
		const int W = band ? ((band + 1) | 1) : q_l;   // nw_geo_init: an odd band is widened by one, a row is band + 1 cells
		return (d8 && W <= 32 * NWQ_NARROW) ? 2 : 3;
	}
	return (q_l > 32 || t_l > 64) ? 1 : 0;
}
static const int nwq_cells[2] = {2048, 8192};   // traceback bytes per problem

struct TaskCtx {
	const NwPen *pen;
	const uint64_t *tseq;
	const uint8_t *qb;
	NwScratch nw;
	WarpCtr *wc;
	const NwQueue *queue;   // pair kernel: where the NW problems go
	const uint64_t *slab;
	int task, tmpl;
};

// The NW calls that need no matrix, solved where they arise: one side empty (nw.c:663-684) and the single mismatch
// between two MEMs (73 % of the calls of short reads, SURVEY 6: one cell of nw.c:166-212 in closed form).
__device__ __forceinline__ bool nw_closed_form(const NwPen &pen, const uint64_t *tseq, const uint8_t *query, int k, int t_s, int t_e,
                                               int q_s, int q_e, NwStat &s) {
	const int t_len = t_e - t_s, q_len = q_e - q_s;
	if (nw_trivial(pen, t_len, q_len, s)) return true;
	if (t_len == 1 && q_len == 1 && k == 0) {
		const int W1 = pen.W1, U = pen.U, NEG = 2 * (pen.MM + U + W1);
		const int sub = pen.d[nw_nuc(tseq, t_s) * 5 + query[q_s]];
		// Dleft = D(0,-1) = W1, aD = D(-1,0) = W1, Qleft = aP = NEG, Ddiag = D(-1,-1) = 0
		int Q = W1 + W1, P = W1 + W1, D, e, fl = 0, x;
		if (Q < P) { D = P; e = 4; } else { D = Q; e = 2; }
		x = NEG + U;
		if (Q < x) { Q = x; if (D <= x) { D = x; e = 3; } } else fl |= 16;
		if (P < x) { P = x; if (D <= x) { D = x; e = 5; } } else fl |= 32;
		x = sub;
		if (D <= x) { D = x; e = 1; }
		if (e == 1 || fl) {   // diagonal: one column; else the run closes here and the walk crosses one gap of each kind (codes 36 / 18)
			s.score = D; s.pos = 0;
			if (e == 1) { s.len = 1; s.match = 1; s.tGaps = 0; s.qGaps = 0; }
			else { s.len = 2; s.match = 0; s.tGaps = 1; s.qGaps = 1; }
			return true;
		}
	}
	return false;
}

// append one NW problem of the task to the queue (all lanes call with the same arguments)
__device__ __forceinline__ void nw_enqueue(const TaskCtx &c, int k, int t_s, int t_e, int q_s, int q_e) {
	const int t_l = t_e - t_s, q_l = q_e - q_s;
	int band = abs(t_l - q_l) + AL_BANDW;
	if (q_l <= band || t_l <= band) band = 0;
	const int cls = nwq_class(t_l, q_l, band, c.queue->d8);
	if ((threadIdx.x & 31) == 0) {
		const NwQueue &Q = *c.queue;
		const unsigned long long slot = atomicAdd(&Q.ctr[A_NPROB], 1ull);
		const unsigned long long ci = atomicAdd(&Q.ctr[A_PCLS + cls], 1ull);
		if (slot < Q.cap && ci < Q.cap) {
			NwProb p;
			p.task = c.task; p.tmpl = c.tmpl; p.t_s = t_s; p.t_e = t_e; p.q_s = q_s; p.q_e = q_e;
			p.kband = (k + 2) | (band << 8);
			p.qoff = (uint32_t)((c.qb - (const uint8_t *)c.slab) >> 3);
			*(int4 *)&Q.probs[slot] = *(const int4 *)&p;
			*((int4 *)&Q.probs[slot] + 1) = *((const int4 *)&p + 1);
			Q.order[(size_t)cls * Q.cap + ci] = (uint32_t)slot;
			if (cls < 2) Q.cells[(size_t)cls * Q.cap + ci] = (uint32_t)(t_l * q_l);
		}
	}
	if (cls >= 2) {   // the warp-per-problem kernels size their scratch from the largest problem
		NwGeo g;
		if (nw_geo_init(g, *c.pen, t_l, q_l, k, band, NW_RS_MAXC)) {
			c.wc->need_e = max(c.wc->need_e, (unsigned)min((size_t)0xF0000000u, g.ebytes() + 4096));
			c.wc->need_q = max(c.wc->need_q, (unsigned)q_l + 64u);
		}
	}
}

__device__ KG_STAGE_INL int kma_score_warp(const AlnParams &P, const TaskCtx &c, const KgTIndexView &ix, const KgTMeta &m, const QView &q,
                              int nN1, int q_len, int q_start, int q_end, Mems &M, int n, NwStat *out) {
	const int lane = threadIdx.x & 31;
	const int k = ix.k, t_len = m.len, U = P.pen.U, Mv = P.pen.M;
	NwStat s = {0, 1, 0, 0, 0, 0};
	if (!n) {
		int dummy;
		if (scan_mems<0, false>(ix, m, c.tseq, q, nN1, q_len, q_start, q_end, M, n, dummy, *c.wc)) return ST_OVERFLOW;
	}
	KG_STAT(c.wc->mems += (unsigned long long)n;)
	if (!n) { *out = s; return ST_OK; }
	unsigned mapQ = 0;
	int start = chain_warp(*c.pen, M, n, q_len, t_len, k, &mapQ, P.mq != 0);
	if ((P.mq != 0 && mapQ < (unsigned)P.mq) || M.sc(start) < k) { *out = s; return ST_OK; }

	// leading tail (leadTailAln, align.c:53-138)
	{
		const int t_e = M.tS(start) - 1, q_e = M.qS(start);
		s.score = 0; s.len = 0; s.pos = t_e; s.match = 0; s.tGaps = 0; s.qGaps = 0;
		if (q_e) {
			int t_s = 0, q_s = 0;
			if ((q_e << 1) < t_e || (q_e + AL_BANDW) < t_e) t_s = t_e - (q_e + min(q_e, AL_BANDW));
			else if ((t_e << 1) < q_e || (t_e + AL_BANDW) < q_e) q_s = q_e - (t_e + min(t_e, AL_BANDW));
			if (t_e - t_s > 0 && q_e - q_s > 0) nw_enqueue(c, -1 - (t_s == 0), t_s, t_e, q_s, q_e);   // adds its fields and moves pos by len - tGaps
		}
	}
	for (;;) {
		const int qS = M.qS(start), qE = M.qE(start);
		const int len = qE - qS;
		s.len += len; s.match += len;
		int sc = 0;
#pragma unroll 1
		for (int i = qS + lane; i < qE; i += 32) { const int b = c.qb[i]; sc += c.pen->d[b * 5 + b]; }
		s.score += warp_sum(sc);
		const int nxt = M.nx(start);
		if (!nxt) break;
		const int q_s = qE, t_s = M.tE(start) - 1;
		int t_e, t_l, q_e;
		start = nxt;
		int qSn = M.qS(start), tSn = M.tS(start);
		if (qSn < q_s) { tSn += q_s - qSn; qSn = q_s; }
		t_e = tSn - 1;
		if (t_e < t_s) {
			if (t_s <= M.tE(start)) { qSn += t_s - t_e; t_e = t_s; t_l = 0; }
			else t_l = t_len - t_s + t_e;
		} else t_l = t_e - t_s;
		__syncwarp();
		if (lane == 0) { M.qS(start) = qSn; M.tS(start) = tSn; }
		__syncwarp();
		q_e = qSn;
		if (abs(t_l - q_e + q_s) * U > q_len * Mv || t_l > q_len || q_e - q_s > (q_len >> 1)) {   // align.c:715
			const int keep = s.pos;
			s.score = 0; s.len = 1; s.pos = keep; s.match = 0; s.tGaps = 0; s.qGaps = 0;
			*out = s;
			return ST_GIVEUP;   // queued problems of the task are void, but for the lead tail's move of pos
		}
		if (t_l > 0 || q_e - q_s > 0) {
			NwStat a;
			if (nw_closed_form(*c.pen, c.tseq, c.qb, 0, t_s, t_e, q_s, q_e, a)) {
				s.score += a.score; s.len += a.len; s.match += a.match; s.tGaps += a.tGaps; s.qGaps += a.qGaps;
				KG_STAT(if (t_l == 1 && q_e - q_s == 1) { ++c.wc->full_calls; ++c.wc->full_cells; })   // the one-cell matrix, as the reference counts it
			} else nw_enqueue(c, 0, t_s, t_e, q_s, q_e);
		}
	}
	// trailing tail (trailTailAln, align.c:140-212)
	{
		const int t_s = M.tE(start) - 1, q_s = M.qE(start);
		int q_e = q_len, t_e = t_len;
		if (((q_len - q_s) << 1) < (t_len - t_s) || (q_len - q_s + AL_BANDW) < (t_len - t_s)) {
			t_e = q_len - q_s; t_e = t_s + (t_e + min(t_e, AL_BANDW));
		} else if (((t_len - t_s) << 1) < (q_len - q_s) || (t_len - t_s + AL_BANDW) < (q_len - q_s)) {
			q_e = t_len - t_s; q_e = q_s + (q_e + min(q_e, AL_BANDW));
		}
		if (t_e - t_s > 0 && q_e - q_s > 0) nw_enqueue(c, 1 + (t_e == t_len), t_s, t_e, q_s, q_e);
	}
	*out = s;
	return ST_OK;
}

// preseed (align.c:750-770): does any k-spaced k-mer of the byte read occur in the template? The key is built from
// bytes exactly as makeKmer (stdnuc.c:424) does, so an N (4) spills into the neighbouring base; bytes past the end
// of the read count as 0.
__device__ bool preseed_hit(const KgTIndexView &ix, const KgTMeta &m, const uint8_t *qb, int q_len, int lim) {
	const int lane = threadIdx.x & 31, k = ix.k;
	for (int i0 = 0; i0 < lim; i0 += 32 * k) {
		const int i = i0 + lane * k;
		int hit = 0;
		if (i < lim) {
			uint64_t key = 0;
			for (int b = 0; b < k; ++b) key = (b ? key << 2 : 0) | (uint64_t)(i + b < q_len ? qb[i + b] : 0);
			hit = tix_get(ix, m, key) != 0;
		}
		if (__any_sync(0xffffffffu, hit)) return true;
	}
	return false;
}

// ---------------------------------------------------------------- the pair kernel

__device__ int align_pair(const AlnParams &P, const NwPen *pen, const KgTIndexView &ix, const uint64_t *slab, const AlnRead &R,
                          int tmpl, Mems M, const NwQueue *queue, int task, WarpCtr &wc, AlnCand *out) {
	const int at = abs(tmpl), q_len = R.q_len, nN1 = R.nN + 1, k = ix.k;
	const KgTMeta m = ix.meta[at];
	TaskCtx c;
	c.pen = pen; c.tseq = ix.seq + m.seq_off; c.wc = &wc; c.queue = queue; c.slab = slab; c.task = task; c.tmpl = at;
	NwStat a = {0, 0, 0, 0, 0, 0};
	int n = 0, strand = 0, st = ST_OK;
	if (R.rc_flag < 0) {   // strand undecided: anker_rc_comp (align.c:993-1176)
		const QView qf = read_view(slab, R, 0), qr = read_view(slab, R, 1);
		int sf = 0, sr = 0, nf = 0, ntot;
		// query bounds are mirrored for the reverse strand; preseed only runs without a lower bound (align.c:1031-1041)
		const bool pre = R.q_start || P.exhaustive || preseed_hit(ix, m, qf.b, q_len, R.q_end);
		if (pre && scan_mems<1, false>(ix, m, c.tseq, qf, nN1, q_len, R.q_start, R.q_end, M, nf, sf, wc)) return ST_OVERFLOW;
		ntot = nf;
		if (scan_mems<1, false>(ix, m, c.tseq, qr, nN1, q_len, q_len - R.q_end, q_len - R.q_start, M, ntot, sr, wc)) return ST_OVERFLOW;
		const int best = max(sf, sr);
		if (P.one2one && best < k && best * k < (q_len - k - best)) { n = 0; strand = -1; }
		else if (best == sf) {   // forward wins ties; a zero score means nothing seeded on either strand
			if (best) { n = nf; strand = 0; tmpl = at; } else strand = -1;
		} else { M.shift(nf); n = ntot - nf; strand = 1; tmpl = -at; }
		if (strand >= 0) {
			const QView q = strand ? qr : qf;
			c.qb = q.b;
			if ((st = kma_score_warp(P, c, ix, m, q, nN1, q_len, 0, q_len, M, n, &a)) == ST_OVERFLOW) return ST_OVERFLOW;   // MEMs are in place: no scan
		}
	} else {
		const QView q = read_view(slab, R, 0);   // SE records carry the strand stage 2 chose (ankers.c:30-50)
		c.qb = q.b;
		if ((st = kma_score_warp(P, c, ix, m, q, nN1, q_len, R.q_start, R.q_end, M, 0, &a)) == ST_OVERFLOW) return ST_OVERFLOW;
	}
	out->tmpl = tmpl; out->score = a.score; out->len = a.len; out->pos = a.pos; out->match = a.match;
	out->tGaps = a.tGaps; out->qGaps = a.qGaps; out->status = st;
	return ST_OK;
}

// KMA_score of one read in a given orientation against one template
__device__ int align_fixed(const AlnParams &P, const NwPen *pen, const KgTIndexView &ix, const uint64_t *slab, const AlnRead &R,
                           int strand, int at, Mems M, const NwQueue *queue, int task, WarpCtr &wc, AlnCand *out) {
	const KgTMeta m = ix.meta[at];
	TaskCtx c;
	c.pen = pen; c.tseq = ix.seq + m.seq_off; c.wc = &wc; c.queue = queue; c.slab = slab; c.task = task; c.tmpl = at;
	const QView q = read_view(slab, R, strand);
	c.qb = q.b;
	NwStat a = {0, 0, 0, 0, 0, 0};
	const int st = kma_score_warp(P, c, ix, m, q, R.nN + 1, R.q_len, 0, R.q_len, M, 0, &a);
	if (st == ST_OVERFLOW) return ST_OVERFLOW;
	out->tmpl = at; out->score = a.score; out->len = a.len; out->pos = a.pos; out->match = a.match;
	out->tGaps = a.tGaps; out->qGaps = a.qGaps; out->status = st;
	return ST_OK;
}

struct ScratchLayout { size_t stride; int mem_cap, q_cap; size_t e_cap; };

template <int MINB>
__global__ void __launch_bounds__(AL_WARPS * 32, MINB) aln_pair_kernel(const AlnParams P_, const KgTIndexView ix_, const uint8_t *__restrict__ in,
		const AlnRead *__restrict__ reads, const uint64_t *slab, const int32_t *__restrict__ task_read, int ntasks,
		const int32_t *__restrict__ task_list, AlnCand *cand, uint8_t *scratch, ScratchLayout lay,
		unsigned long long *ctr, int32_t *ovf_list, const NwQueue queue_) {
	// the out-of-line stages take the parameters, the index view and the queue by reference: one copy per CTA in shared
	// memory instead of one per thread in local memory
	__shared__ AlnParams sP;
	__shared__ KgTIndexView six;
	__shared__ NwQueue squeue;
	for (int i = threadIdx.x; i < (int)(sizeof(AlnParams) / 4); i += blockDim.x) ((int *)&sP)[i] = ((const int *)&P_)[i];
	for (int i = threadIdx.x; i < (int)(sizeof(KgTIndexView) / 4); i += blockDim.x) ((int *)&six)[i] = ((const int *)&ix_)[i];
	for (int i = threadIdx.x; i < (int)(sizeof(NwQueue) / 4); i += blockDim.x) ((int *)&squeue)[i] = ((const int *)&queue_)[i];
	__syncthreads();
	const AlnParams &P = sP;
	const KgTIndexView &ix = six;
	NwPen &spen = sP.pen;
	const int lane = threadIdx.x & 31;
	const size_t wid = (size_t)blockIdx.x * AL_WARPS + (threadIdx.x >> 5);
	Mems M;
	M.base = (int *)(scratch + wid * lay.stride);
	M.cap = lay.mem_cap;
	WarpCtr wc;
	memset(&wc, 0, sizeof(wc));
	unsigned long long tnext = 0;
	int tleft = 0;   // tasks are claimed AL_CLAIM at a time: one contended atomic per four tasks, and a read's tasks stay on one warp;
	                 // one at a time when the batch has few (long) tasks per warp: there the balance is what counts
	const int claim = (long long)ntasks >= 64ll * gridDim.x * AL_WARPS ? AL_CLAIM : 1;
	for (;;) {
		if (!tleft) {
			if (lane == 0) tnext = atomicAdd(&ctr[A_WORK], (unsigned long long)claim);
			tnext = __shfl_sync(0xffffffffu, tnext, 0);
			tleft = claim;
		}
		const unsigned long long t = tnext++;
		--tleft;
		if (t >= (unsigned long long)ntasks) break;
		const int task = task_list ? task_list[t] : (int)t;
		const int r = task_read[task];
		const AlnRead R = reads[r];
		const uint8_t *rec = in + R.rec_off;
		int ti = task - (int)R.task0;
		const int mate = R.kind == 2 ? ti & 1 : 0;
		if (R.kind == 2) ti >>= 1;
		const int tmpl = (int)ld_u32u(rec + 28 + 8 * (size_t)R.words + 4 * (size_t)R.nN + 4 * (size_t)ti);
		AlnCand res;
		int st;
		if (R.kind == 2) {   // a mate of a pair against one template, strands decided by stage 2 (alnfrags.c:1645-1661, 1712-1731)
			const AlnRead Rq = mate ? R : reads[r - 1];
			KG_STAT(wc.read_bytes += 8ull * (unsigned long long)Rq.words + (unsigned long long)Rq.q_len + 4ull * (unsigned long long)Rq.nN;)
			st = align_fixed(P, &spen, ix, slab, Rq, ti >= R.fneg, abs(tmpl), M, &squeue, task, wc, &res);
			res.tmpl = tmpl;
		} else {
			KG_STAT(wc.read_bytes += 8ull * (unsigned long long)R.words + (unsigned long long)R.q_len + 4ull * (unsigned long long)R.nN;)
			st = align_pair(P, &spen, ix, slab, R, tmpl, M, &squeue, task, wc, &res);
		}
		__syncwarp();
		if (st != ST_OK) {
			res.tmpl = tmpl; res.score = res.len = res.pos = res.match = res.tGaps = res.qGaps = 0; res.status = ST_OVERFLOW;
			if (lane == 0) { const unsigned long long o = atomicAdd(&ctr[A_OVF], 1ull); ovf_list[o] = task; }
		}
		if (lane == 0) cand[task] = res;
	}
	if (lane == 0) {
		if (wc.mems) atomicAdd(&ctr[A_MEMS], wc.mems);
		if (wc.lookups) atomicAdd(&ctr[A_LOOKUPS], wc.lookups);
		if (wc.mem_bases) atomicAdd(&ctr[A_MEMBASES], wc.mem_bases);
		if (wc.read_bytes) atomicAdd(&ctr[A_READBYTES], wc.read_bytes);
		if (wc.full_calls) { atomicAdd(&ctr[A_FULL_CALLS], wc.full_calls); atomicAdd(&ctr[A_FULL_CELLS], wc.full_cells); }
		if (wc.need_e) atomicMax(&ctr[A_NEED_E], (unsigned long long)wc.need_e);
		if (wc.need_mem) atomicMax(&ctr[A_NEED_MEM], (unsigned long long)wc.need_mem);
		if (wc.need_q) atomicMax(&ctr[A_NEED_Q], (unsigned long long)wc.need_q);
	}
}

// The pair kernel without the statistic counters is compiled in translation units of its own (kmagpu_align_fast.cu for
// long reads, kmagpu_align_fast_short.cu for short ones include this file inside a namespace with KG_NO_STATS and
// KG_PAIR_VARIANT_ONLY): the counters cost it registers and time. The launchers have C linkage and take the structs by
// address because the translation units define them in different namespaces (same source, same layout).
#ifdef KG_PAIR_VARIANT_ONLY
extern "C" void KG_VARIANT_LAUNCHER(int grid, cudaStream_t st, const void *P, const void *ix, const uint8_t *in,
                                    const void *reads, const uint64_t *slab, const int32_t *task_read, int ntasks, const int32_t *task_list,
                                    void *cand, uint8_t *scratch, const void *lay, unsigned long long *ctr, int32_t *ovf_list, const void *queue) {
	aln_pair_kernel<KG_VARIANT_MINB><<<grid, AL_WARPS * 32, 0, st>>>(*(const AlnParams *)P, *(const KgTIndexView *)ix, in, (const AlnRead *)reads, slab,
		task_read, ntasks, task_list, (AlnCand *)cand, scratch, *(const ScratchLayout *)lay, ctr, ovf_list, *(const NwQueue *)queue);
}
#else
#define KG_DECL_LAUNCHER(name)                                                                                                              \
	extern "C" void name(int grid, cudaStream_t st, const void *P, const void *ix, const uint8_t *in, const void *reads, const uint64_t *slab, \
	                     const int32_t *task_read, int ntasks, const int32_t *task_list, void *cand, uint8_t *scratch, const void *lay,      \
	                     unsigned long long *ctr, int32_t *ovf_list, const void *queue)
KG_DECL_LAUNCHER(kg_launch_pair_fast_long);
KG_DECL_LAUNCHER(kg_launch_pair_fast_short);

// ---------------------------------------------------------------- selection + ConClave sums (one thread per read)

// what the writer needs per stage-2 record. Single read: one frag_raw record (kept, best). Pair (stored with the mate
// that carries the templates): up to two records; rec[x] = {kept, score field, flag, mate (0 = first record of the
// pair), orientation (1 = reverse complement bytes), offset of its start[]/end[]/template[] arrays in the pair's ints}
struct AlnRes {
	int32_t kept, best;
	int32_t form;        // 0 single read, 1 proper pair (update_Scores_pe), 2 one or two update_Scores_se records
	int32_t nrec;
	int32_t rkept[2], rscore[2], rflag[2], rmate[2], rorient[2], roff[2];
};

struct PeEnt { int32_t mt, bT, bTr, bS, bE; };   // the five parallel arrays of alnFragsPenaltyPE, index = template index

// update_Scores_se / update_Scores_pe (updatescores.c:300-488): keep the best templates, add the ConClave sums.
// T/Sc/S/E are strided views into the PeEnt array (stride 5 ints); kept triples are copied to out[3 * cap].
__device__ int pe_keep(double minFrac, int pe, int n, int best, const int32_t *S, const int32_t *E, const int32_t *T, const int32_t *Sc,
                       int32_t *out, int cap, unsigned long long *as, unsigned long long *uas) {
	int kept = 0;
	const int mode = minFrac == 1.0 ? 0 : (minFrac < 0 ? 1 : 2);
	const double thr = fabs(minFrac) * best;
	for (int i = 0; i < n; ++i) {
		const int sc = Sc[5 * i];
		const bool keep = mode == 0 ? sc == best : thr <= sc;
		if (keep) {
			out[kept] = S[5 * i]; out[cap + kept] = E[5 * i]; out[2 * cap + kept] = T[5 * i];
			const int add = pe ? (mode == 2 ? best : sc) : (mode == 1 ? sc : best);
			atomicAdd(&as[abs(T[5 * i])], (unsigned long long)add);
			++kept;
		}
	}
	if (kept == 1) atomicAdd(&uas[abs(out[2 * cap])], (unsigned long long)best);
	return kept;
}

// alnFragsPenaltyPE (alnfrags.c:1596-1972) for one pair, after its 2 * nt KMA_score results are in cand[task0 ..).
__device__ void reduce_pair(const AlnParams &P, const AlnRead &RA, const AlnRead &RB, const uint8_t *recB, AlnCand *cand,
                            const KgTMeta *meta, unsigned long long *as, unsigned long long *uas, AlnRes &o, uint32_t &size) {
	const int nt = RB.nt, k = P.k, Wl = -P.Wl, PE = P.PE;
	int32_t *base = (int32_t *)(cand + RB.task0);
	PeEnt *ent = (PeEnt *)base;
	int32_t *outA = base + 5 * (nt + 1), *outB = outA + 3 * nt;
	const uint8_t *T = recB + 28 + 8 * (size_t)RB.words + 4 * (size_t)RB.nN;
	int best1 = 0, best2 = 0, comp = 0, start = 0, end = 0, hits = 0;
	double score = 0;
	for (int ti = 1; ti <= nt; ++ti) {
		const AlnCand a1 = cand[RB.task0 + 2 * (ti - 1)], a2 = cand[RB.task0 + 2 * (ti - 1) + 1];   // read before the slots are reused
		const int tmpl = (int)ld_u32u(T + 4 * (size_t)(ti - 1));
		const int t_len = meta[abs(tmpl)].len;
		PeEnt e;
		e.mt = tmpl;
		int rs = a1.score;
		if (P.minlen <= a1.len && 0 < rs && ((P.mrc * RA.q_len <= a1.len - a1.qGaps) || (P.mrc * t_len <= a1.len - a1.tGaps))) {
			start = a1.pos; end = a1.pos + a1.len - a1.tGaps;
			if (start == 0) rs += Wl;
			if (end == t_len) rs += Wl;
			score = 1.0 * rs / a1.len;
		} else rs = 0;
		if (rs > k && score >= P.scoreT) { e.bT = rs; e.bS = start; e.bE = end; if (best1 < rs) best1 = rs; }
		else { e.bT = 0; e.bS = -1; e.bE = -1; }
		rs = a2.score;
		if (P.minlen <= a2.len && 0 < rs && ((P.mrc * RB.q_len <= a2.len - a2.qGaps) || (P.mrc * t_len <= a2.len - a2.tGaps))) {
			start = a2.pos; end = a2.pos + a2.len - a2.tGaps;
			if (start == 0) rs += Wl;
			if (end == t_len) rs += Wl;
			score = 1.0 * rs / a2.len;
		} else rs = 0;
		if (rs > k && score >= P.scoreT) {
			e.bTr = rs;
			if (e.bT) { if (start < e.bS) e.bS = start; else e.bE = end; }
			else { e.bS = start; e.bE = end; }
			if (best2 < rs) best2 = rs;
		} else e.bTr = 0;
		rs += e.bT;
		if (comp < rs) comp = rs;
		if (ti == 1) { PeEnt z = {nt, 0, 0, 0, 0}; ent[0] = z; }
		ent[ti] = e;
	}
	o.form = 2; o.nrec = 0; size = 0;
	if (!best1 && !best2) return;
	const int flipped = RB.fneg < nt, rc = !flipped;
	const double af = P.minFrac < 0 ? -P.minFrac : P.minFrac;
	int flag = RA.flag, flag_r = RB.flag, o1 = flipped, o2 = flipped;
	int32_t *e0 = (int32_t *)ent;
#define F_MT 0
#define F_BT 1
#define F_BTR 2
#define F_BS 3
#define F_BE 4
	auto put = [&](int x, int kept, int sc, int fl, int mate, int orient, int32_t *arr) {
		o.rkept[x] = kept; o.rscore[x] = sc; o.rflag[x] = fl; o.rmate[x] = mate; o.rorient[x] = orient; o.roff[x] = (int32_t)(arr - base);
	};
	const uint32_t szA = (uint32_t)RA.q_len + (uint32_t)RA.hl, szB = (uint32_t)RB.q_len + (uint32_t)RB.hl;
	bool proper;
	int best = 0;
	if (P.apm == 1) {   // alnFragsUnionPE (alnfrags.c:1408-1422): templates both mates reach within minFrac of their own best
		if (best1 && best2) {
			const double sc = af * best1, sc_r = af * best2;
			for (int ti = 1; ti <= nt; ++ti)
				if (sc <= ent[ti].bT && sc_r <= ent[ti].bTr) {
					const PeEnt e = ent[ti];
					ent[hits].bTr = e.bT + e.bTr; ent[hits].bT = e.mt; ent[hits].bS = e.bS; ent[hits].bE = e.bE; ++hits;
				}
		}
		proper = hits != 0;
		best = best1 + best2;
	} else {            // alnFragsPenaltyPE (alnfrags.c:1787-1808)
		proper = comp && af * (best1 + best2) <= (comp + PE);
		if (proper) {
			best = comp + PE;
			for (int ti = 1; ti <= nt; ++ti)
				if (ent[ti].bT && ent[ti].bTr) {
					const PeEnt e = ent[ti];
					ent[hits].bTr = e.bT + e.bTr + PE; ent[hits].bT = e.mt; ent[hits].bS = e.bS; ent[hits].bE = e.bE; ++hits;
				}
		}
	}
	if (proper) {   // proper pair
		o.form = 1; o.nrec = 2;
		if (ent[0].bT < 0) {
			for (int i = 0; i < hits; ++i) ent[i].bT = -ent[i].bT;
			const int kept = pe_keep(P.minFrac, 1, hits, best, e0 + F_BS, e0 + F_BE, e0 + F_BT, e0 + F_BTR, outA, nt, as, uas);
			put(0, kept, -best, flag_r, 1, o2, outA); put(1, 0, 0, flag, 0, o1, outA);
			size = 20u + szB + 12u * kept + 12u + szA;
		} else {
			if (!rc) { o1 = o2 = 0; flag ^= 48; flag_r ^= 48; }
			const int kept = pe_keep(P.minFrac, 1, hits, best, e0 + F_BS, e0 + F_BE, e0 + F_BT, e0 + F_BTR, outA, nt, as, uas);
			put(0, kept, -best, flag, 0, o1, outA); put(1, 0, 0, flag_r, 1, o2, outA);
			size = 20u + szA + 12u * kept + 12u + szB;
		}
	} else if (best1 && best2) {                         // both map, not as a pair
		int hits_r = 0, ti = 1, last = nt, tmp;
		const double sc = af * best1, sc_r = af * best2;
		while (ti <= last) {
			if (sc <= ent[ti].bT) { ent[hits].mt = ent[ti].mt; ent[hits].bT = ent[ti].bT; ent[hits].bS = ent[ti].bS; ent[hits].bE = ent[ti].bE; ++hits; ++ti; }
			else if (sc_r <= ent[ti].bTr) {
				tmp = ent[ti].mt; ent[ti].mt = ent[last].mt; ent[last].mt = tmp;
				tmp = ent[ti].bTr; ent[ti].bTr = ent[last].bTr; ent[last].bTr = tmp;
				tmp = ent[ti].bS; ent[ti].bS = ent[last].bS; ent[last].bS = tmp;
				tmp = ent[ti].bE; ent[ti].bE = ent[last].bE; ent[last].bE = tmp;
				++hits_r; --last;
			} else ++ti;
		}
		if (ent[0].bT < 0) { for (int i = 0; i < hits; ++i) ent[i].bT = -ent[i].bT; }
		else if (!rc) { o1 = 0; flag ^= 16; flag_r ^= 32; }
		if (ent[last].bTr < 0) { for (int i = 0; i < hits_r; ++i) ent[last + i].bTr = -ent[last + i].bTr; }
		else if (!rc) { o2 = 0; flag ^= 32; flag_r ^= 16; }
		if (flag & 2) { flag ^= 2; flag_r ^= 2; }
		const int k1 = pe_keep(P.minFrac, 0, hits, best1, e0 + F_BS, e0 + F_BE, e0 + F_MT, e0 + F_BT, outA, nt, as, uas);
		ent[0].mt = k1;   // the reference stores the count in slot 0 between the two calls (alnfrags.c:1884)
		int32_t *eL = e0 + 5 * last;
		const int k2 = pe_keep(P.minFrac, 0, hits_r, best2, eL + F_BS, eL + F_BE, eL + F_MT, eL + F_BTR, outB, nt, as, uas);
		o.nrec = 2;
		put(0, k1, best1, flag, 0, o1, outA); put(1, k2, best2, flag_r, 1, o2, outB);
		size = 20u + szA + 12u * k1 + 20u + szB + 12u * k2;
	} else if (best1) {                                  // first mate only
		for (int ti = 1; ti <= nt; ++ti)
			if (ent[ti].bT) { const PeEnt e = ent[ti]; ent[hits].bTr = e.bT; ent[hits].bT = e.mt; ent[hits].bS = e.bS; ent[hits].bE = e.bE; ++hits; }
		if (ent[0].bT < 0) { for (int i = 0; i < hits; ++i) ent[i].bT = -ent[i].bT; }
		else if (!rc) { o1 = 0; flag ^= 16; flag_r ^= 32; }
		flag |= 8; flag_r ^= 4;
		if (flag & 2) { flag ^= 2; flag_r ^= 2; }
		const int k1 = pe_keep(P.minFrac, 0, hits, best1, e0 + F_BS, e0 + F_BE, e0 + F_BT, e0 + F_BTR, outA, nt, as, uas);
		o.nrec = 1;
		put(0, k1, best1, flag, 0, o1, outA);
		size = 20u + szA + 12u * k1;
	} else {                                             // second mate only
		for (int ti = 1; ti <= nt; ++ti)
			if (ent[ti].bTr) { const PeEnt e = ent[ti]; ent[hits].bTr = e.bTr; ent[hits].bT = e.mt; ent[hits].bS = e.bS; ent[hits].bE = e.bE; ++hits; }
		if (ent[0].bTr < 0) { for (int i = 0; i < hits; ++i) ent[i].bTr = -ent[i].bTr; }
		else if (!rc) { o2 = 0; flag ^= 32; flag_r ^= 16; }
		flag_r |= 8; flag ^= 4;
		if (flag_r & 2) { flag ^= 2; flag_r ^= 2; }
		const int k2 = pe_keep(P.minFrac, 0, hits, best2, e0 + F_BS, e0 + F_BE, e0 + F_BT, e0 + F_BTR, outA, nt, as, uas);
		o.nrec = 1;
		put(0, k2, best2, flag_r, 1, o2, outA);
		size = 20u + szB + 12u * k2;
	}
}

__global__ void aln_reduce_kernel(AlnParams P, const uint8_t *__restrict__ in, const AlnRead *__restrict__ reads, int n, AlnCand *cand,
                                  const KgTMeta *__restrict__ meta, unsigned long long *as, unsigned long long *uas,
                                  uint32_t *recsize, AlnRes *res, unsigned long long *ctr) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n) return;
	const AlnRead R = reads[r];
	if (R.kind) {
		AlnRes o;
		memset(&o, 0, sizeof(o));
		uint32_t size = 0;
		if (R.kind == 2 && R.nt > 0 && reads[r - 1].q_len >= P.k && R.q_len >= P.k) {
			reduce_pair(P, reads[r - 1], R, in + R.rec_off, cand, meta, as, uas, o, size);
			if (o.nrec) warp_add_u64(&ctr[A_FRAGS], (unsigned)(o.form == 1 ? 1 : o.nrec));
		}
		recsize[r] = size;
		res[r] = o;
		return;
	}
	const int k = P.k, q_len = R.q_len;
	const int nt = q_len >= k ? R.nt : 0;
	AlnCand *c = cand + R.task0;
	double bestScore = 0;
	int best_read = 0, hits = 0;
	for (int ti = 0; ti < nt; ++ti) {   // alnfrags.c:1131-1198
		const AlnCand a = c[ti];
		const int t_len = meta[abs(a.tmpl)].len, aln_len = a.len, start = a.pos;
		int end = start + aln_len - a.tGaps, read_score = a.score;
		double score;
		if (t_len < end) end -= t_len;
		if (q_len <= aln_len || t_len <= aln_len) score = aln_len; else score = q_len < t_len ? q_len : t_len;
		if (P.minlen <= aln_len && ((P.mrc * q_len <= a.len - a.qGaps) || (P.mrc * t_len <= a.len - a.tGaps))) score = read_score / score;
		else { read_score = 0; score = 0; }
		if (k < read_score && P.scoreT <= score) {
			AlnCand h;
			h.tmpl = a.tmpl; h.pos = start; h.match = end; h.score = read_score; h.len = aln_len; h.tGaps = h.qGaps = h.status = 0;
			c[hits++] = h;
			if (bestScore < score) bestScore = score;
			if (best_read < read_score) best_read = read_score;
		}
	}
	uint32_t size = 0;
	AlnRes o;
	memset(&o, 0, sizeof(o));
	if (best_read > k) {   // update_Scores (updatescores.c:203-298)
		int kept = 0;
		double minScore = 0, minFrac = P.minFrac;
		const int mode = P.minFrac == 1.0 ? 0 : (P.minFrac < 0 ? 1 : 2);
		if (mode) { minScore = fabs(P.minFrac) * bestScore; minFrac = fabs(P.minFrac) * best_read; }
		for (int i = 0; i < hits; ++i) {
			const AlnCand h = c[i];
			bool keep;
			if (mode == 0) { const double ms = h.score / h.len; keep = ms == bestScore || h.score == best_read; }
			else keep = (h.len * minScore <= h.score) || minFrac <= h.score;
			if (keep) {
				c[kept++] = h;
				atomicAdd(&as[abs(h.tmpl)], (unsigned long long)(mode == 2 ? best_read : h.score));
			}
		}
		if (kept == 1) atomicAdd(&uas[abs(c[0].tmpl)], (unsigned long long)best_read);
		o.kept = kept; o.best = best_read;
		size = 20u + (uint32_t)q_len + (uint32_t)R.hl + 12u * (uint32_t)kept;
		warp_add_u64(&ctr[A_FRAGS], 1u);
	}
	recsize[r] = size;
	res[r] = o;
}

// frag_raw record (updatescores.c:284-295): int32[5]{q_len, hits, score, hdrlen, flag} read(0-4) header start[] end[] template[]
__global__ void __launch_bounds__(256, ST_MINB) aln_emit_kernel(const uint8_t *__restrict__ in, const AlnRead *__restrict__ reads, int n,
		const uint64_t *__restrict__ slab, const AlnCand *__restrict__ cand, const AlnRes *__restrict__ res,
		const uint32_t *__restrict__ out_off, uint8_t *__restrict__ out) {
	const unsigned lane = threadIdx.x & 31;
	const int warps = (gridDim.x * blockDim.x) >> 5;
	for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
		const AlnRes rs = res[r];
		if (rs.nrec) {   // pair: update_Scores_pe (one record + mate block) or update_Scores_se records (updatescores.c:300-488)
			const int32_t *pbase = (const int32_t *)(cand + reads[r].task0);
			const int cap = reads[r].nt;
			uint8_t *o = out + out_off[r];
#pragma unroll
			for (int x = 0; x < 2; ++x) {   // unrolled: rs stays in registers (indexed by x it lived in local memory)
				if (x >= rs.nrec) break;
				const AlnRead Rm = reads[r - 1 + rs.rmate[x]];
				const QView q = read_view(slab, Rm, rs.rorient[x]);
				const uint8_t *hdr = in + Rm.rec_off + 28 + 8 * (size_t)Rm.words + 4 * (size_t)Rm.nN + 4 * (size_t)Rm.nt;
				const bool mate_block = rs.form == 1 && x == 1;   // int32[3]{q_len, hdrlen, flag} read header
				if (mate_block) {
					if (lane < 3) st_u32b(o + 4 * lane, (uint32_t)(lane == 0 ? Rm.q_len : lane == 1 ? Rm.hl : rs.rflag[x]));
					o += 12;
				} else {
					if (lane < 5) {
						const int32_t h = lane == 0 ? Rm.q_len : lane == 1 ? rs.rkept[x] : lane == 2 ? rs.rscore[x] : lane == 3 ? Rm.hl : rs.rflag[x];
						st_u32b(o + 4 * lane, (uint32_t)h);
					}
					o += 20;
				}
				warp_copy(o, q.b, Rm.q_len, lane);
				o += Rm.q_len;
				warp_copy(o, hdr, Rm.hl, lane);
				o += Rm.hl;
				if (!mate_block) {
					const int32_t *arr = pbase + rs.roff[x];
					const int kept = rs.rkept[x];
					warp_store_u32(o, 3 * kept, lane, [&](int i) { const int a = i / kept; return arr[a * cap + (i - a * kept)]; });
					o += 12 * (size_t)kept;
				}
				__syncwarp();
			}
			continue;
		}
		if (rs.best == 0) continue;
		const AlnRead R = reads[r];
		uint8_t *o = out + out_off[r];
		if (lane < 5) {
			const int32_t h = lane == 0 ? R.q_len : lane == 1 ? rs.kept : lane == 2 ? rs.best : lane == 3 ? R.hl : R.flag;
			st_u32b(o + 4 * lane, (uint32_t)h);
		}
		o += 20;
		const QView q = read_view(slab, R, 0);
		warp_copy(o, q.b, R.q_len, lane);
		o += R.q_len;
		const uint8_t *hdr = in + R.rec_off + 28 + 8 * (size_t)R.words + 4 * (size_t)R.nN + 4 * (size_t)R.nt;
		warp_copy(o, hdr, R.hl, lane);
		o += R.hl;
		const AlnCand *c = cand + R.task0;
		const int kept = rs.kept;
		warp_store_u32(o, 3 * kept, lane, [&](int i) { const int a = i / kept, x = i - a * kept; return a == 0 ? c[x].pos : (a == 1 ? c[x].match : c[x].tmpl); });
	}
}


// ---------------------------------------------------------------- traceback alignment (assemble_KMA's inner loop)

#define ST_ROWS 3   // aligned rows longer than the per-record capacity

struct TrRec {
	uint32_t rec_off;
	int32_t tmpl, q_len, score, hl, nN, words;
	uint32_t slab_off;      // 8-byte units
	uint32_t row_cap;       // columns reserved per row
	unsigned long long row_off;   // byte offset of the record's three rows in the row pool
	int32_t q_start, q_end; // query bounds of a chain-mode fragment (name ends in \0, start, end: assembly.c:1916-1923); else 0, q_len
};

__host__ __device__ __forceinline__ uint32_t tr_stride(const TrRec &R) { return slab_W(R.words) + slab_B(R.q_len) + slab_N(R.nN); }

__device__ __forceinline__ QView tr_view(const uint64_t *slab, const TrRec &R, int strand) {
	const uint64_t *base = slab + R.slab_off + (strand ? tr_stride(R) : 0);
	QView v;
	v.w = base;
	v.b = (const uint8_t *)(base + slab_W(R.words));
	v.N = (const int32_t *)(base + slab_W(R.words) + slab_B(R.q_len));
	return v;
}

// one warp per record: header fields, number of N's, slab and row-pool sizes
__global__ void __launch_bounds__(256) tr_sizes_kernel(const uint8_t *__restrict__ in, const uint32_t *__restrict__ off, int n, int DB_size,
		TrRec *recs, uint32_t *slab_sz, uint32_t *row_sz, unsigned long long *ctr) {
	const unsigned lane = threadIdx.x & 31;
	const int warps = (gridDim.x * blockDim.x) >> 5;
	int maxq = 0;   // one atomic per warp at the end, not one per record
	for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
		const uint8_t *rec = in + off[r];
		TrRec R;
		R.rec_off = off[r];
		R.tmpl = (int)ld_u32u(rec); R.q_len = (int)ld_u32u(rec + 4); R.score = (int)ld_u32u(rec + 12); R.hl = (int)ld_u32u(rec + 24);
		const uint8_t *q = rec + 32;
		int cnt = 0;
#pragma unroll 1
		for (int i = lane; i < R.q_len; i += 32) cnt += q[i] == 4;
		R.nN = warp_sum(cnt);
		R.words = (R.q_len + 31) >> 5;
		R.row_cap = 3u * (uint32_t)R.q_len + 256u;
		R.slab_off = 0; R.row_off = 0;
		R.q_start = 0; R.q_end = R.q_len;
		if (9 < R.hl) {
			const uint8_t *he = q + R.q_len + R.hl;
			if (he[-9] == 0) { R.q_start = (int)ld_u32u(he - 8); R.q_end = (int)ld_u32u(he - 4); }
		}
		if (lane == 0) {
			if (R.tmpl <= 0 || R.tmpl >= DB_size) atomicAdd(&ctr[A_BAD], 1ull);
			recs[r] = R;
			slab_sz[r] = tr_stride(R) * (R.score == 0 ? 2u : 1u);
			row_sz[r] = (3u * R.row_cap + 7u) >> 3;   // 8-byte units
			maxq = max(maxq, R.q_len);
		}
	}
	if (lane == 0 && maxq > 0) atomicMax(&ctr[A_MAXQ], (unsigned long long)maxq);
}

// bytes -> slab (packed words, bytes, N list + sentinel), reverse complement too for reads without a strand
__global__ void __launch_bounds__(256, ST_MINB) tr_prep_kernel(const uint8_t *__restrict__ in, int n, TrRec *recs, const uint32_t *__restrict__ slab_off,
		const uint32_t *__restrict__ row_off, uint64_t *slab) {
	const unsigned lane = threadIdx.x & 31;
	const int warps = (gridDim.x * blockDim.x) >> 5;
	for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
		TrRec R = recs[r];
		R.slab_off = slab_off[r]; R.row_off = 8ull * row_off[r];
		if (lane == 0) { recs[r].slab_off = R.slab_off; recs[r].row_off = R.row_off; }
		const uint8_t *src = in + R.rec_off + 32;
		const int L = R.q_len;
		for (int strand = 0; strand < (R.score == 0 ? 2 : 1); ++strand) {
			uint64_t *base = slab + R.slab_off + (strand ? tr_stride(R) : 0);
			uint8_t *b = (uint8_t *)(base + slab_W(R.words));
			int32_t *N = (int32_t *)(base + slab_W(R.words) + slab_B(L));
			// one round per 32 bases: the byte (complemented and mirrored for the reverse strand), its 2-bit code OR-reduced
			// over the warp into the packed word (compDNA's `(word << 2) | base` is the OR of `base << (62 - 2 lane)`), and
			// the N positions in ascending order by ballot rank
			const int nbytes = (int)(slab_B(L) << 3);
			int cnt = 0;
#pragma unroll 2
			for (int w = 0; w < R.words + 2; ++w) {
				const int i = 32 * w + (int)lane;
				uint8_t v = 0;
				if (i < L) { v = strand ? __ldg(src + L - 1 - i) : __ldg(src + i); if (strand && v < 4) v = 3 - v; }
				if (i < nbytes) b[i] = v;
				const uint32_t c = v < 4 ? v : 0u;
				const uint32_t hi = __reduce_or_sync(0xffffffffu, lane < 16 ? c << (30 - 2 * lane) : 0u);
				const uint32_t lo = __reduce_or_sync(0xffffffffu, lane >= 16 ? c << (62 - 2 * lane) : 0u);
				if (lane == 0) base[w] = ((uint64_t)hi << 32) | lo;
				const bool isn = i < L && v == 4;
				const unsigned mk = __ballot_sync(0xffffffffu, isn);
				if (isn) N[cnt + __popc(mk & ((1u << lane) - 1))] = i;
				cnt += __popc(mk);
			}
			if (lane == 0) N[cnt] = L;   // sentinel
			__syncwarp();
		}
	}
}

// KMA (align.c:214-507): MEMs (found here with the byte-read scan unless n != 0), chain, stitch with NW -- writing the
// aligned rows. *ncol = columns written (0 when nothing aligned).
__device__ int kma_trace_warp(const AlnParams &P, const TaskCtx &c, const KgTIndexView &ix, const KgTMeta &m, const QView &q,
                              int nN1, int q_len, int q_start, int q_end, Mems &M, int n, NwStat *out, const NwRows &rows, int row_cap, int *ncol) {
	const int lane = threadIdx.x & 31;
	const int k = ix.k, t_len = m.len, U = P.pen.U, Mv = P.pen.M;
	NwStat s = {0, 1, 0, 0, 0, 0};
	*ncol = 0;
	if (!n) {
		int dummy;
		if (scan_mems<0, true>(ix, m, c.tseq, q, nN1, q_len, q_start, q_end, M, n, dummy, *c.wc)) return ST_OVERFLOW;
	}
	KG_STAT(c.wc->mems += (unsigned long long)n;)
	if (!n) { *out = s; return ST_OK; }
	unsigned mapQ = 0;
	int start = chain_warp(*c.pen, M, n, q_len, t_len, k, &mapQ, P.mq != 0);
	if ((P.mq != 0 && mapQ < (unsigned)P.mq) || M.sc(start) < k) { *out = s; return ST_OK; }
	if (P.ts) {   // trimSeeds (chain.c:496-538): the first ts bases of every seed of the chain go back to the DP
		__syncwarp();
		if (lane == 0) {
			// MEM 0 is a valid chain start: only next == 0 ends the walk (the reference's do ... while, chain.c:509-524)
			int cidx = start;
			bool go = true;
			if (!M.qS(cidx)) { cidx = M.nx(cidx); go = cidx != 0; }
			while (go) {
				const int len = M.qE(cidx) - M.qS(cidx), cut = len < P.ts ? len - 1 : P.ts;
				M.tS(cidx) += cut; M.qS(cidx) += cut;
				cidx = M.nx(cidx);
				go = cidx != 0;
			}
		}
		__syncwarp();
	}
	auto nw_rows = [&](int kk, int t_s, int t_e, int q_s, int q_e, int at, NwStat *a) -> int {
		if (at + (t_e - t_s) + (q_e - q_s) + 8 > row_cap) return ST_ROWS;
		const int t_l = t_e - t_s, q_l = q_e - q_s;
		if (t_l == 1 && q_l == 1 && kk == 0) {   // the single mismatch between two MEMs: one cell in closed form, one diagonal column
			NwStat a1;
			if (nw_closed_form(*c.pen, c.tseq, c.qb, 0, t_s, t_e, q_s, q_e, a1) && a1.len == 1) {
				if (lane == 0) {
					const uint8_t tb = (uint8_t)nw_nuc(c.tseq, t_s), qb = c.qb[q_s];
					rows.t[at] = tb; rows.s[at] = tb == qb ? '|' : '_'; rows.q[at] = qb;
				}
				++c.wc->full_calls; ++c.wc->full_cells;
				*a = a1;
				return ST_OK;
			}
		}
		int band = abs(t_l - q_l) + AL_BANDW;
		if (q_l <= band || t_l <= band) band = 0;
		NwRows r = {rows.t + at, rows.s + at, rows.q + at};
		unsigned long long cells = 0;
		const int st = nw_warp<0>(*c.pen, c.tseq, c.qb, kk, t_s, t_e, q_s, q_e, band, c.nw, a, &cells, &r);
		if (st != NW_OK) {
			NwGeo g;
			nw_geo_init(g, *c.pen, t_l, q_l, kk, band, 0);
			c.wc->need_e = max(c.wc->need_e, (unsigned)min((size_t)0xF0000000u, g.ebytes() + 4096));
			c.wc->need_q = max(c.wc->need_q, (unsigned)q_l + 64u);
			return ST_OVERFLOW;
		}
		if (cells) { if (band) { ++c.wc->band_calls; c.wc->band_cells += cells; } else { ++c.wc->full_calls; c.wc->full_cells += cells; } }
		return ST_OK;
	};
	// leading tail (leadTailAln with Frag_align, align.c:53-138)
	{
		const int t_e = M.tS(start) - 1, q_e = M.qS(start);
		s.score = 0; s.len = 0; s.pos = t_e; s.match = 0; s.tGaps = 0; s.qGaps = 0;
		if (q_e) {
			int t_s = 0, q_s = 0;
			if ((q_e << 1) < t_e || (q_e + AL_BANDW) < t_e) t_s = t_e - (q_e + min(q_e, AL_BANDW));
			else if ((t_e << 1) < q_e || (t_e + AL_BANDW) < q_e) q_s = q_e - (t_e + min(t_e, AL_BANDW));
			if (t_e - t_s > 0 && q_e - q_s > 0) {
				NwStat a;
				const int st = nw_rows(-1 - (t_s == 0), t_s, t_e, q_s, q_e, 0, &a);
				if (st) return st;
				if (t_s == 0) {   // trim leading gap columns (align.c:99-113)
					int bias = 0;
					while (bias < a.len) {
						const int i = bias + lane;
						const int gt = i < a.len && rows.t[i] == 5, gq = i < a.len && rows.q[i] == 5;
						const unsigned gm = __ballot_sync(0xffffffffu, gt || gq), tm = __ballot_sync(0xffffffffu, gt);
						const int run = ~gm ? __ffs(~gm) - 1 : 32;
						const unsigned low = run == 32 ? 0xffffffffu : ((1u << run) - 1);
						a.tGaps -= __popc(tm & low); a.qGaps -= __popc(gm & ~tm & low);
						bias += run;
						if (run < 32) break;
					}
					if (bias > a.len) bias = a.len;
					if (bias) {   // shift the rows left by `bias`, 32 columns at a time
						for (int i0 = 0; i0 < a.len - bias; i0 += 32) {
							const int i = i0 + lane;
							uint8_t vt = 0, vs = 0, vq = 0;
							if (i < a.len - bias) { vt = rows.t[i + bias]; vs = rows.s[i + bias]; vq = rows.q[i + bias]; }
							__syncwarp();
							if (i < a.len - bias) { rows.t[i] = vt; rows.s[i] = vs; rows.q[i] = vq; }
							__syncwarp();
						}
						a.len -= bias;
					}
				}
				s.pos -= a.len - a.tGaps;
				s.score = a.score; s.len = a.len; s.match = a.match; s.tGaps = a.tGaps; s.qGaps = a.qGaps;
			}
		}
	}
	for (;;) {
		const int qS = M.qS(start), qE = M.qE(start);
		const int len = qE - qS;
		if (s.len + len + 8 > row_cap) return ST_ROWS;
		int sc = 0;
#pragma unroll 1
		for (int i = qS + lane; i < qE; i += 32) {
			const int b = c.qb[i];
			rows.t[s.len + i - qS] = (uint8_t)b; rows.s[s.len + i - qS] = '|'; rows.q[s.len + i - qS] = (uint8_t)b;
			sc += P.pen.d[b * 5 + b];
		}
		s.len += len; s.match += len;
		s.score += warp_sum(sc);
		const int nxt = M.nx(start);
		if (!nxt) break;
		const int q_s = qE, t_s = M.tE(start) - 1;
		int t_e, t_l, q_e;
		start = nxt;
		int qSn = M.qS(start), tSn = M.tS(start);
		if (qSn < q_s) { tSn += q_s - qSn; qSn = q_s; }
		t_e = tSn - 1;
		if (t_e < t_s) {
			if (t_s <= M.tE(start)) { qSn += t_s - t_e; t_e = t_s; t_l = 0; }
			else t_l = t_len - t_s + t_e;
		} else t_l = t_e - t_s;
		__syncwarp();
		if (lane == 0) { M.qS(start) = qSn; M.tS(start) = tSn; }
		__syncwarp();
		q_e = qSn;
		if (abs(t_l - q_e + q_s) * U > q_len * Mv || t_l > q_len || q_e - q_s > (q_len >> 1)) {   // align.c:465
			const int keep = s.pos;
			s.score = 0; s.len = 1; s.pos = keep; s.match = 0; s.tGaps = 0; s.qGaps = 0;
			*out = s;
			return ST_OK;
		}
		if (t_l > 0 || q_e - q_s > 0) {
			NwStat a;
			const int st = nw_rows(0, t_s, t_e, q_s, q_e, s.len, &a);
			if (st) return st;
			s.score += a.score; s.len += a.len; s.match += a.match; s.tGaps += a.tGaps; s.qGaps += a.qGaps;
		}
	}
	// trailing tail (trailTailAln with Frag_align, align.c:147-212)
	{
		const int t_s = M.tE(start) - 1, q_s = M.qE(start);
		int q_e = q_len, t_e = t_len;
		if (((q_len - q_s) << 1) < (t_len - t_s) || (q_len - q_s + AL_BANDW) < (t_len - t_s)) {
			t_e = q_len - q_s; t_e = t_s + (t_e + min(t_e, AL_BANDW));
		} else if (((t_len - t_s) << 1) < (q_len - q_s) || (t_len - t_s + AL_BANDW) < (q_len - q_s)) {
			q_e = t_len - t_s; q_e = q_s + (q_e + min(q_e, AL_BANDW));
		}
		if (t_e - t_s > 0 && q_e - q_s > 0) {
			NwStat a;
			const int st = nw_rows(1 + (t_e == t_len), t_s, t_e, q_s, q_e, s.len, &a);
			if (st) return st;
			if (t_e == t_len) {   // trim trailing gap columns (align.c:183-199); column 0 is never trimmed
				int bias = a.len - 1;
				const uint8_t *rt = rows.t + s.len, *rq = rows.q + s.len;
				while (bias > 0) {
					const int i = bias - lane;
					const int gt = i > 0 && rt[i] == 5, gq = i > 0 && rq[i] == 5;
					const unsigned gm = __ballot_sync(0xffffffffu, gt || gq), tm = __ballot_sync(0xffffffffu, gt);
					const int run = ~gm ? __ffs(~gm) - 1 : 32;
					const unsigned low = run == 32 ? 0xffffffffu : ((1u << run) - 1);
					a.tGaps -= __popc(tm & low); a.qGaps -= __popc(gm & ~tm & low);
					bias -= run;
					if (run < 32) break;
				}
				if (bias < 0) bias = 0;
				a.len = bias + 1;
			}
			s.score += a.score; s.len += a.len; s.match += a.match; s.tGaps += a.tGaps; s.qGaps += a.qGaps;
		}
	}
	*ncol = s.len;
	*out = s;
	return ST_OK;
}

struct TrOut { int32_t h[12]; int32_t status; };

// one warp per fragment record: (anker_rc when the strand is open) -> KMA -> acceptance (assembly.c:1925-1961)
// anker_rc (align.c:780-991) for a fragment whose strand is open: both strands' MEMs, the better one stays in M.
// Out of line: fragments that come from the alignment pass carry their strand, so this is off their path.
struct TrAnker { int strand, nmem, go, st, oriented; };
__device__ __noinline__ void tr_anker_rc(const AlnParams &P, const KgTIndexView &ix, const KgTMeta &m, const uint64_t *tseq, const uint64_t *slab,
                                         const TrRec &R, Mems &M, WarpCtr &wc, TrAnker &A) {
	const int lane = threadIdx.x & 31, k = ix.k, q_len = R.q_len, nN1 = R.nN + 1;
	const QView qf = tr_view(slab, R, 0), qr = tr_view(slab, R, 1);
	int sf = 0, sr = 0, nf = 0, ntot, st = ST_OK;
	A.strand = 0; A.nmem = 0; A.go = 0; A.oriented = 0;
	// query bounds (chain-mode fragments): a lower bound skips preseed, the reverse strand sees them mirrored (align.c:806-817)
	const bool pre = R.q_start || P.exhaustive || preseed_hit(ix, m, qf.b, q_len, R.q_end - R.q_start);
	if (pre) st = scan_mems<1, true>(ix, m, tseq, qf, nN1, q_len, R.q_start, R.q_end, M, nf, sf, wc);
	ntot = nf;
	if (!st) st = scan_mems<1, true>(ix, m, tseq, qr, nN1, q_len, q_len - R.q_end, q_len - R.q_start, M, ntot, sr, wc);
	const int best = max(sf, sr);
	if (!st) {
		int turned = 0;
		if (P.one2one && best < k && best * k < (q_len - k - best)) turned = 1;   // rejected; the read stays turned
		else if (best == sf) { A.nmem = nf; A.go = best != 0; }
		else { A.strand = 1; M.shift(nf); A.nmem = ntot - nf; A.go = 1; turned = 1; }
		if (turned) {   // "oriented": do the bytes differ from what came in? (a palindrome does not)
			int diff = 0;
#pragma unroll 1
			for (int i = lane; i < q_len; i += 32) diff |= qf.b[i] != qr.b[i];
			A.oriented = __any_sync(0xffffffffu, diff) ? 1 : 0;
		}
	}
	A.st = st;
}

#ifndef TR_MINB
#define TR_MINB 8             // resident CTAs per SM the traceback kernel is compiled for (64 registers). C2 traceback stage: 4 -> 32.0 ms, 5 -> 30.1, 6 -> 29.6, 8 -> 28.5 (profiles/r02_trace_kernel.log)
#endif
// one warp per fragment record: (anker_rc when the strand is open) -> KMA -> acceptance (assembly.c:1925-1961)
__global__ void __launch_bounds__(AL_WARPS * 32, TR_MINB) tr_task_kernel(const AlnParams P_, const KgTIndexView ix_, const TrRec *__restrict__ recs,
		const uint64_t *slab, int n, const int32_t *__restrict__ task_list, TrOut *outs, uint8_t *rowpool, uint8_t *scratch,
		ScratchLayout lay, unsigned long long *ctr, int32_t *ovf_list) {
	// the parameter blocks once per CTA in shared memory: the out-of-line stages take them by reference
	__shared__ AlnParams sP;
	__shared__ KgTIndexView six;
	__shared__ NwRow sring[AL_WARPS][NW_RING];
	for (int i = threadIdx.x; i < (int)(sizeof(AlnParams) / 4); i += blockDim.x) ((int *)&sP)[i] = ((const int *)&P_)[i];
	for (int i = threadIdx.x; i < (int)(sizeof(KgTIndexView) / 4); i += blockDim.x) ((int *)&six)[i] = ((const int *)&ix_)[i];
	__syncthreads();
	const AlnParams &P = sP;
	const KgTIndexView &ix = six;
	NwPen &spen = sP.pen;
	const int lane = threadIdx.x & 31;
	const size_t wid = (size_t)blockIdx.x * AL_WARPS + (threadIdx.x >> 5);
	uint8_t *sp = scratch + wid * lay.stride;
	Mems M0;
	{
		int *p = (int *)sp;
		const int c1 = lay.mem_cap + 1;
		M0.base = p;
		M0.cap = lay.mem_cap;
		sp += (size_t)c1 * 32;
	}
	NwScratch nws;
	nws.ring = sring[threadIdx.x >> 5];
	nws.rowbuf = (NwRow *)sp; nws.e_cap = lay.e_cap; nws.q_cap = lay.q_cap; nws.finish();
	WarpCtr wc;
	memset(&wc, 0, sizeof(wc));
	unsigned long long tnext = 0;
	int tleft = 0;
	const int claim = (long long)n >= 64ll * gridDim.x * AL_WARPS ? AL_CLAIM : 1;   // as in the pair kernel
	for (;;) {
		if (!tleft) {
			if (lane == 0) tnext = atomicAdd(&ctr[A_WORK], (unsigned long long)claim);
			tnext = __shfl_sync(0xffffffffu, tnext, 0);
			tleft = claim;
		}
		const unsigned long long t = tnext++;
		--tleft;
		if (t >= (unsigned long long)n) break;
		const int r = task_list ? task_list[t] : (int)t;
		const TrRec R = recs[r];
		const KgTMeta m = ix.meta[R.tmpl];
		const int q_len = R.q_len, nN1 = R.nN + 1, t_len = m.len;
		Mems M = M0;
		TaskCtx c;
		c.pen = &spen; c.tseq = ix.seq + m.seq_off; c.nw = nws; c.wc = &wc;
		int strand = 0, nmem = 0, st = ST_OK, oriented = 0;
		bool go = R.score != 0;
		if (!go) {
			TrAnker A;
			tr_anker_rc(P, ix, m, c.tseq, slab, R, M, wc, A);
			strand = A.strand; nmem = A.nmem; go = A.go != 0; st = A.st; oriented = A.oriented;
		}
		NwStat a = {0, 0, 0, 0, 0, 0};
		int ncol = 0;
		bool done = false;
		if (go && !st) {
			const QView q = tr_view(slab, R, strand);
			c.qb = q.b;
			NwRows rows;
			rows.t = rowpool + R.row_off; rows.s = rows.t + R.row_cap; rows.q = rows.s + R.row_cap;
			st = kma_trace_warp(P, c, ix, m, q, nN1, q_len, R.q_start, R.q_end, M, nmem, &a, rows, (int)R.row_cap, &ncol);
			done = !st;
		}
		__syncwarp();
		if (st == ST_OVERFLOW && lane == 0) { const unsigned long long x = atomicAdd(&ctr[A_OVF], 1ull); ovf_list[x] = r; }
		if (st == ST_ROWS && lane == 0) atomicAdd(&ctr[A_BAD], 1ull);
		if (lane == 0) {   // the acceptance test of assemble_KMA (assembly.c:1925-1961) and the header the host needs
			TrOut o;
#pragma unroll
			for (int i = 0; i < 12; ++i) o.h[i] = 0;
			o.h[10] = oriented;
			if (done) {
				const int aln_len = a.len, start = a.pos;
				int end = start + aln_len - a.tGaps, read_score = a.score;
				double score;
				if (t_len < end) end -= t_len;
				if (start == 0) read_score += -P.Wl;
				if (end == t_len) read_score += -P.Wl;
				if (P.minlen <= aln_len && ((P.mrc * q_len <= a.len - a.qGaps) || (P.mrc * t_len <= a.len - a.tGaps))) score = 1.0 * read_score / aln_len;
				else { read_score = 0; score = 0; }
				o.h[0] = 0 < read_score && P.scoreT <= score;
				o.h[1] = read_score; o.h[2] = start; o.h[3] = end;
				o.h[4] = a.score; o.h[5] = a.len; o.h[6] = a.pos; o.h[7] = a.match; o.h[8] = a.tGaps; o.h[9] = a.qGaps;
				o.h[11] = ncol;
			}
			o.status = st;
			outs[r] = o;
		}
	}
	if (lane == 0) {
		if (wc.mems) atomicAdd(&ctr[A_MEMS], wc.mems);
		if (wc.full_calls) { atomicAdd(&ctr[A_FULL_CALLS], wc.full_calls); atomicAdd(&ctr[A_FULL_CELLS], wc.full_cells); }
		if (wc.band_calls) { atomicAdd(&ctr[A_BAND_CALLS], wc.band_calls); atomicAdd(&ctr[A_BAND_CELLS], wc.band_cells); }
		if (wc.need_e) atomicMax(&ctr[A_NEED_E], (unsigned long long)wc.need_e);
		if (wc.need_mem) atomicMax(&ctr[A_NEED_MEM], (unsigned long long)wc.need_mem);
		if (wc.need_q) atomicMax(&ctr[A_NEED_Q], (unsigned long long)wc.need_q);
	}
}

__global__ void tr_outsize_kernel(const TrOut *__restrict__ outs, int n, uint32_t *size) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r < n) size[r] = 48u + 3u * (uint32_t)outs[r].h[11];
}

// per record: int32[12] header + t, s, q rows (ncol bytes each), input order
__global__ void __launch_bounds__(256) tr_emit_kernel(const TrRec *__restrict__ recs, const TrOut *__restrict__ outs, int n,
		const uint8_t *__restrict__ rowpool, const uint32_t *__restrict__ out_off, uint8_t *__restrict__ out) {
	const unsigned lane = threadIdx.x & 31;
	const int warps = (gridDim.x * blockDim.x) >> 5;
	for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
		const TrOut o = outs[r];
		const TrRec R = recs[r];
		uint8_t *dst = out + out_off[r];
		if (lane < 12) st_u32b(dst + 4 * lane, (uint32_t)o.h[lane]);
		dst += 48;
		const int ncol = o.h[11];
		const uint8_t *src = rowpool + R.row_off;
		for (int row = 0; row < 3; ++row)
#pragma unroll 1
			for (int i = lane; i < ncol; i += 32) dst[(size_t)row * ncol + i] = src[(size_t)row * R.row_cap + i];
	}
}


// ---------------------------------------------------------------- base-count matrix (alnToMat / alnToMatDense)

// One warp per accepted alignment: +1 on counts[template position][query code] for every aligned column that has a
// template base (assembly.c:1317-1444 restricted to the template nodes; assembly.c:1446-1497 when dense). The
// reference serialises these updates under a lock and saturates its uint16 counters; +1 increments commute, so the
// device adds atomically into uint32 and the read-out clamps to 65535 (= the saturated sum). Gap-column trimming
// follows the reference: alnToMat drops gap columns at both ends (its trailing loop stops at column 0) and moves
// the start past leading deletions, alnToMatDense only drops trailing ones.
__global__ void __launch_bounds__(256) tr_matrix_kernel(const TrRec *__restrict__ recs, const TrOut *__restrict__ outs, int n,
		const uint8_t *__restrict__ rowpool, const KgTMeta *__restrict__ meta, const int64_t *__restrict__ mat_off, int dense,
		unsigned int *mat, unsigned long long *ctr) {
	const unsigned lane = threadIdx.x & 31;
	const unsigned lt = (1u << lane) - 1;
	const int warps = (gridDim.x * blockDim.x) >> 5;
	unsigned long long added = 0;
	for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
		const TrOut o = outs[r];
		if (!o.h[0]) continue;
		const TrRec R = recs[r];
		const uint8_t *t = rowpool + R.row_off, *q = t + 2 * (size_t)R.row_cap;
		const int t_len = meta[R.tmpl].len;
		unsigned int *C = mat + 6 * (size_t)mat_off[R.tmpl];
		int aln_len = o.h[5], start = o.h[6], i0 = 0;
		// trailing gap columns
		{
			int i = aln_len - 1;
			for (;;) {
				const int c = i - (int)lane;
				const bool gap = c >= (dense ? 0 : 1) && (t[c] == 5 || q[c] == 5);
				const unsigned stop = __ballot_sync(0xffffffffu, !gap);
				if (stop) { i -= __ffs(stop) - 1; break; }
				i -= 32;
			}
			aln_len = i + 1;
		}
		if (!dense) {   // leading gap columns; deletions move the start
			for (;;) {
				const int c = i0 + (int)lane;
				const bool gap = c < aln_len && (t[c] == 5 || q[c] == 5);
				const unsigned stop = __ballot_sync(0xffffffffu, !gap);
				const int take = stop ? __ffs(stop) - 1 : 32;
				start += __popc(__ballot_sync(0xffffffffu, gap && q[c] == 5) & (take == 32 ? 0xffffffffu : (1u << take) - 1));
				i0 += take;
				if (stop) break;
			}
		}
		int pos = start % t_len;
		for (int base = i0; base < aln_len; base += 32) {
			const int c = base + (int)lane;
			const bool has = c < aln_len && t[c] != 5;
			const unsigned m = __ballot_sync(0xffffffffu, has);
			if (has) {
				int p = pos + __popc(m & lt);
				if (p >= t_len) p %= t_len;
				atomicAdd(&C[6 * (size_t)p + q[c]], 1u);
			}
			pos += __popc(m);
			if (pos >= t_len) pos %= t_len;
			added += lane == 0 ? __popc(m) : 0;
		}
	}
	if (lane == 0 && added) atomicAdd(&ctr[A_MEMBASES], added);
}

__global__ void __launch_bounds__(256) mat_clamp_kernel(const unsigned int *__restrict__ mat, size_t n, uint16_t *__restrict__ out) {
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
		out[i] = (uint16_t)min(mat[i], 65535u);
}

// ---------------------------------------------------------------- host side

int kg_align_free(kmagpu_db *db) {
	{
		TraceBatch &t = db->trc;
		KgBuf *tr[] = {&t.d_in, &t.d_off, &t.d_recs, &t.d_sz, &t.d_partial, &t.d_ctr, &t.d_slab, &t.d_rows, &t.d_outs, &t.d_ovf, &t.d_out,
		               &db->frg.d_out, &db->frg.d_sz, &db->frg.d_sc, &db->frg.d_items, &db->frg.d_keys, &db->frg.d_vals, &db->frg.d_partial,
		               &db->frg.d_ctr, &db->frg.d_acc, &db->frg.d_tmp};
		for (KgBuf *x : tr) x->release();
		db->frg.valid = false;
	}
	AlignBatch &b = db->aln;
	KgBuf *all[] = {&b.d_in, &b.d_off, &b.d_reads, &b.d_slab, &b.d_sz, &b.d_partial, &b.d_taskread, &b.d_cand, &b.d_recsize,
	                &b.d_out, &b.d_ctr, &b.d_scores, &b.d_scratch, &b.d_ovf, &b.d_res, &b.h_off, &b.d_probs, &b.d_order, &b.d_sorttmp};
	for (KgBuf *x : all) x->release();
	return 0;
}

static AlnParams make_params(const kmagpu_db *db, const kmagpu_params *p) {
	AlnParams P;
	memset(&P, 0, sizeof(P));
	P.pen.W1 = p->W1; P.pen.U = p->U; P.pen.MM = p->MM; P.pen.M = p->M;
	memcpy(P.pen.d, p->d, sizeof(P.pen.d));
	P.pen.d8 = 1;
	for (int i = 0; i < 25; ++i) if (p->d[i] < -128 || p->d[i] > 127) P.pen.d8 = 0;
	P.k = db->info.kmerindex; P.mq = p->mq; P.one2one = p->one2one; P.exhaustive = p->exhaustive; P.minlen = p->minlen; P.Wl = p->Wl; P.PE = p->PE; P.apm = p->apm; P.ts = p->ts;
	P.scoreT = p->scoreT; P.mrc = p->mrc; P.minFrac = p->minFrac;
	return P;
}

extern "C" int kmagpu_align_upload(kmagpu_db *db, const void *stage2, size_t nbytes, int64_t *nreads_out) {
	if (!db || (!stage2 && nbytes)) { kmagpu_set_error("null argument"); return -1; }
	if (nbytes >= (1ull << 32) - 64) { kmagpu_set_error("stage-2 batch of %zu bytes exceeds the 4 GiB per-call limit; split it", nbytes); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	AlignBatch &b = db->aln;
	b.h_off.pinned = true;
	const uint8_t *in = (const uint8_t *)stage2;
	size_t guess = nbytes / 64 + 16;
	if (b.h_off.reserve(4 * (guess + 1))) return -1;
	uint32_t *off = (uint32_t *)b.h_off.p;
	size_t cap = b.h_off.cap / 4 - 1, n = 0, ip = 0;
	while (ip + 28 <= nbytes) {   // record walk (get_ankers, ankers.c:163-220)
		int32_t h[7];
		memcpy(h, in + ip, 28);
		if (h[0] < 0) break;   // stream terminator -(number of reads)
		size_t len = 28 + 8 * (size_t)(uint32_t)h[1] + 4 * (size_t)(uint32_t)h[2] + 4 * (size_t)(uint32_t)h[4] + (size_t)(uint32_t)h[5];
		if (h[1] < 0 || h[2] < 0 || h[4] < 0 || h[5] < 0 || ip + len > nbytes) { kmagpu_set_error("stage-2 stream is truncated or corrupt at byte %zu", ip); return -1; }
		if (kg_check_record(in + ip, 2, db->info.DB_size, ip)) return -1;
		if (n == cap) {
			KgBuf bigger; bigger.pinned = true;
			if (bigger.reserve(8 * (cap + 1))) return -1;
			memcpy(bigger.p, off, 4 * n);
			b.h_off.release();
			b.h_off = bigger;
			off = (uint32_t *)b.h_off.p; cap = b.h_off.cap / 4 - 1;
		}
		off[n++] = (uint32_t)ip;
		ip += len;
	}
	off[n] = (uint32_t)ip;
	b.nreads = (int64_t)n; b.in_bytes = ip; b.ran = false;
	if (nreads_out) *nreads_out = (int64_t)n;
	if (b.d_in.reserve(ip + 64) || b.d_off.reserve(4 * (n + 2))) return -1;
	KG_CUDA(cudaEventRecord(db->ev[5], db->stream));
	KG_CUDA(cudaMemcpyAsync(b.d_in.p, in, ip, cudaMemcpyHostToDevice, db->stream));
	KG_CUDA(cudaMemsetAsync((uint8_t *)b.d_in.p + ip, 0, 64, db->stream));
	KG_CUDA(cudaMemcpyAsync(b.d_off.p, off, 4 * (n + 1), cudaMemcpyHostToDevice, db->stream));
	KG_CUDA(cudaEventRecord(db->ev[6], db->stream));
	KG_CUDA(cudaStreamSynchronize(db->stream));
	b.in = (const uint8_t *)b.d_in.p;
	return 0;
}

// defined in kmagpu_seed.cu: device view of the last stage-2 stream (all reads; unmapped ones have empty records)
int kg_seed_device_output(kmagpu_db *db, const uint8_t **out, const uint32_t **rec_off, int64_t *nreads, size_t *bytes);

extern "C" int kmagpu_align_from_seed(kmagpu_db *db, int64_t *nreads_out) {
	if (!db) { kmagpu_set_error("null argument"); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	AlignBatch &b = db->aln;
	const uint8_t *out; const uint32_t *roff; int64_t n; size_t bytes;
	if (kg_seed_device_output(db, &out, &roff, &n, &bytes)) return -1;
	if (b.d_off.reserve(4 * ((size_t)n + 2))) return -1;
	const uint32_t total = (uint32_t)bytes;
	KG_CUDA(cudaMemcpyAsync(b.d_off.p, roff, 4 * (size_t)n, cudaMemcpyDeviceToDevice, db->stream));
	KG_CUDA(cudaMemcpyAsync((uint32_t *)b.d_off.p + n, &total, 4, cudaMemcpyHostToDevice, db->stream));
	KG_CUDA(cudaStreamSynchronize(db->stream));
	b.in = out; b.nreads = n; b.in_bytes = bytes; b.ran = false;
	if (nreads_out) *nreads_out = n;
	return 0;
}

static ScratchLayout make_layout(int mem_cap, int q_cap, size_t e_cap) {
	ScratchLayout l;
	l.mem_cap = mem_cap; l.q_cap = q_cap; l.e_cap = (e_cap + 255) & ~(size_t)255;
	size_t s = (size_t)(mem_cap + 1) * 32;
	s += (size_t)q_cap * 12 + l.e_cap;
	l.stride = (s + 255) & ~(size_t)255;
	return l;
}

// ---------------------------------------------------------------- NW queue kernels (phase 2 of the alignment pass)

// add a problem's AlnScore into the candidate row of its task (AlnCand as int32[8]: tmpl score len pos match tGaps qGaps status)
__device__ __forceinline__ void nwq_apply(int32_t *row, const NwStat &a, int k, int status) {
	if (k < 0) atomicSub(&row[3], a.len - a.tGaps);   // leadTailAln: pos -= len - tGaps (align.c:119)
	if (status == ST_GIVEUP) return;
	atomicAdd(&row[1], a.score); atomicAdd(&row[2], a.len); atomicAdd(&row[4], a.match); atomicAdd(&row[5], a.tGaps); atomicAdd(&row[6], a.qGaps);
}

// classes 0-1: one thread per problem (nw_thread); the previous DP row and the query bases of the CTA's 128 problems in
// shared memory [column][thread], traceback bytes in a per-warp scratch [cell][lane]
template <int QMAX, bool D8>
__global__ void __launch_bounds__(128) nw_thread_kernel(const NwPen pen, const KgTIndexView ix, const NwProb *__restrict__ probs,
		const uint32_t *__restrict__ order, int n, const uint8_t *qbase, int qshift, int32_t *res, uint8_t *escratch, int ecells,
		unsigned long long *ctr) {
	extern __shared__ NwRow srows[];   // [QMAX][128] rows, then [QMAX][128] query bytes
	__shared__ NwPen spen;
	__shared__ unsigned long long stab[5];
	if (threadIdx.x < sizeof(NwPen) / 4) ((int *)&spen)[threadIdx.x] = ((const int *)&pen)[threadIdx.x];
	__syncthreads();
	if (threadIdx.x < 5) stab[threadIdx.x] = nw_pack_row(spen, threadIdx.x);
	__syncthreads();
	const int tid = blockIdx.x * 128 + threadIdx.x, nthreads = gridDim.x * 128;
	uint8_t *E = escratch + (size_t)(tid >> 5) * (size_t)ecells * 32 + (tid & 31);
	NwRow *rows = srows + threadIdx.x;
	uint8_t *qs = (uint8_t *)(srows + QMAX * 128) + threadIdx.x;
	unsigned long long cells = 0;
	unsigned calls = 0;
	for (int idx = tid; idx < n; idx += nthreads) {
		const NwProb p = probs[order[idx]];
		int32_t *row = res + 8 * (size_t)p.task;
		const int status = row[7], k = (p.kband & 255) - 2;
		if (status == ST_GIVEUP && k >= 0) continue;
		const KgTMeta m = ix.meta[p.tmpl];
		const int t_len = p.t_e - p.t_s, q_len = p.q_e - p.q_s;
		const uint8_t *qlast = qbase + ((size_t)p.qoff << qshift) + p.q_e - 1;
		for (int j = 0; j < q_len; ++j) qs[j * 128] = (uint8_t)(qlast[-j] << 3);
		NwStat a;
		nw_thread<D8>(spen, stab, ix.seq + m.seq_off, p.t_s, t_len, qs, 128, q_len, k, rows, 128, E, 32, &a);
		nwq_apply(row, a, k, status);
		cells += (unsigned long long)(t_len * q_len); ++calls;
	}
#pragma unroll
	for (int o = 16; o; o >>= 1) { cells += __shfl_xor_sync(0xffffffffu, cells, o); calls += __shfl_xor_sync(0xffffffffu, calls, o); }
	if ((threadIdx.x & 31) == 0 && calls) { atomicAdd(&ctr[A_FULL_CELLS], cells); atomicAdd(&ctr[A_FULL_CALLS], (unsigned long long)calls); }
}

// classes 2-3: one warp per problem (nw_warp: row sweep for rows of up to 32 * RSMAX cells, the continuous wavefront beyond)
#ifndef NW_NARROW_MINB
#define NW_NARROW_MINB 6   // resident CTAs per SM the narrow build is compiled for
#endif
template <int RSMAX>
__global__ void __launch_bounds__(AL_WARPS * 32, RSMAX <= NWQ_NARROW ? NW_NARROW_MINB : 1) nw_warp_kernel(const NwPen pen, const KgTIndexView ix, const NwProb *__restrict__ probs,
		const uint32_t *__restrict__ order, int n, const uint8_t *qbase, int qshift, int32_t *res, int32_t *status_out,
		uint8_t *scratch, ScratchLayout lay, unsigned long long *ctr) {
	__shared__ NwPen spen;
	__shared__ NwRow sring[AL_WARPS][NW_RING];
	if (threadIdx.x < sizeof(NwPen) / 4) ((int *)&spen)[threadIdx.x] = ((const int *)&pen)[threadIdx.x];
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const size_t wid = (size_t)blockIdx.x * AL_WARPS + (threadIdx.x >> 5);
	NwScratch nws;
	nws.ring = sring[threadIdx.x >> 5];
	nws.rowbuf = (NwRow *)(scratch + wid * lay.stride); nws.e_cap = lay.e_cap; nws.q_cap = lay.q_cap; nws.finish();
	unsigned long long fcells = 0, bcells = 0, fcalls = 0, bcalls = 0, steps = 0;
	for (;;) {
		unsigned long long t = 0;
		if (lane == 0) t = atomicAdd(&ctr[A_WORK], 1ull);
		t = __shfl_sync(0xffffffffu, t, 0);
		if (t >= (unsigned long long)n) break;
		const NwProb p = probs[order[t]];
		int32_t *row = res + 8 * (size_t)p.task;
		const int status = row[7], k = (p.kband & 255) - 2, band = p.kband >> 8;
		if (status == ST_GIVEUP && k >= 0) continue;
		const KgTMeta m = ix.meta[p.tmpl];
		NwStat a = {0, 0, 0, 0, 0, 0};
		unsigned long long cells = 0;
		const int st = nw_warp<RSMAX>(spen, ix.seq + m.seq_off, qbase + ((size_t)p.qoff << qshift), k, p.t_s, p.t_e, p.q_s, p.q_e, band, nws, &a, &cells);
		if (st == NW_OK) {
			if (lane == 0) nwq_apply(row, a, k, status);
			if (cells) {
				NwGeo g;
				nw_geo_init(g, spen, p.t_e - p.t_s, p.q_e - p.q_s, k, band, RSMAX);
				steps += g.C ? (unsigned long long)g.t_len * g.C : (unsigned long long)g.Tmax;
				if (band) { bcells += cells; ++bcalls; } else { fcells += cells; ++fcalls; }
			}
		} else if (lane == 0 && !status_out) atomicAdd(&ctr[A_BAD], 1ull);
		if (lane == 0 && status_out) status_out[p.task] = st;
		__syncwarp();
	}
	if (lane == 0) {
		if (fcalls) { atomicAdd(&ctr[A_FULL_CALLS], fcalls); atomicAdd(&ctr[A_FULL_CELLS], fcells); }
		if (bcalls) { atomicAdd(&ctr[A_BAND_CALLS], bcalls); atomicAdd(&ctr[A_BAND_CELLS], bcells); }
		if (steps) atomicAdd(&ctr[A_STEPS], steps);
	}
}

// Solve the queued problems. counts[c] = problems of class c (their queue slots in order[c * cap ..), their cell counts
// in cells[c * cap ..) for the thread classes; cells == NULL: the caller's order is sorted already). need_e / need_q:
// scratch the largest class-2 problem asked for. sorted: 4 * cap uint32 of working space for the sort. Uses (and may
// grow) the batch's per-warp scratch buffer.
template <int QMAX, bool D8>
static int nw_thread_launch(kmagpu_db *db, const AlnParams &P, const NwProb *probs, const uint32_t *order, int n, const uint8_t *qbase, int qshift,
                            int32_t *res, int ecells, unsigned long long *ctr) {
	const size_t smem = (size_t)QMAX * 128 * (sizeof(NwRow) + 1);
	KG_CUDA(cudaFuncSetAttribute(nw_thread_kernel<QMAX, D8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int per_sm = 0;
	KG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nw_thread_kernel<QMAX, D8>, 128, smem));
	const int grid = std::max(1, std::min((n + 127) / 128, db->sm_count * std::max(per_sm, 1)));
	KgBuf &scr = db->aln.d_scratch;
	if (scr.reserve((size_t)grid * 4 * (size_t)ecells * 32)) return -1;
	nw_thread_kernel<QMAX, D8><<<grid, 128, smem, db->stream>>>(P.pen, db->tix, probs, order, n, qbase, qshift, res, (uint8_t *)scr.p, ecells, ctr);
	return 0;
}

static int nw_queue_run(kmagpu_db *db, const AlnParams &P, const NwProb *probs, const uint32_t *order, const uint32_t *cells, uint32_t *sorted,
                        size_t cap, const unsigned long long *counts, size_t need_e, int need_q, const uint8_t *qbase, int qshift,
                        int32_t *res, int32_t *status_out, unsigned long long *ctr, int *launches) {
	cudaStream_t st = db->stream;
	KgBuf &scr = db->aln.d_scratch;
	unsigned long long cnt[NWQ_CLASSES];
	for (int c = 0; c < NWQ_CLASSES; ++c) cnt[c] = counts[c];
	// a thread class with too few problems to fill the machine (C2: 2400 problems of up to 64 x 128 cells = 19 CTAs, 0.75 ms
	// of one long thread each) goes to the narrow warp kernel instead: a warp sweeps such a problem in ~20 us
	for (int c = 0; c < 2; ++c)
		if (cnt[c] && cnt[c] < 32ull * (unsigned long long)db->sm_count && cnt[2] + cnt[c] <= cap) {
			KG_CUDA(cudaMemcpyAsync(const_cast<uint32_t *>(order) + 2 * cap + cnt[2], order + (size_t)c * cap, 4 * (size_t)cnt[c], cudaMemcpyDeviceToDevice, st));
			cnt[2] += cnt[c]; cnt[c] = 0;
		}
	for (int c = 0; c < NWQ_CLASSES; ++c) {
		const int n = (int)cnt[c];
		if (!n) continue;
		const uint32_t *ord = order + (size_t)c * cap;
		if (cells && c < 2) {   // largest problems first, neighbours alike: what a warp works on together finishes together
			uint32_t *ks = sorted, *vs = ks + cap;
			size_t tmp_bytes = 0;
			const int bits = 14;   // at most 128 x 64 cells
			cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, cells + (size_t)c * cap, ks, ord, vs, n, 0, bits, st);
			if (db->aln.d_sorttmp.reserve(tmp_bytes + 256)) return -1;
			cub::DeviceRadixSort::SortPairsDescending(db->aln.d_sorttmp.p, tmp_bytes, cells + (size_t)c * cap, ks, ord, vs, n, 0, bits, st);
			ord = vs;
			*launches += 3;
		}
		int rc = 0;
		if (c == 0) rc = P.pen.d8 ? nw_thread_launch<32, true>(db, P, probs, ord, n, qbase, qshift, res, nwq_cells[0], ctr)
		                          : nw_thread_launch<32, false>(db, P, probs, ord, n, qbase, qshift, res, nwq_cells[0], ctr);
		else if (c == 1) rc = P.pen.d8 ? nw_thread_launch<64, true>(db, P, probs, ord, n, qbase, qshift, res, nwq_cells[1], ctr)
		                               : nw_thread_launch<64, false>(db, P, probs, ord, n, qbase, qshift, res, nwq_cells[1], ctr);
		else {
			ScratchLayout lay;
			lay.mem_cap = 0; lay.q_cap = std::max(need_q, 256); lay.e_cap = (std::max<size_t>(need_e, 65536) + 255) & ~(size_t)255;
			lay.stride = ((size_t)lay.q_cap * 12 + lay.e_cap + 255) & ~(size_t)255;
			int per_sm = 0;   // what the build's registers allow (narrow: NW_NARROW_MINB, wide: 3)
			if (c == 2) KG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nw_warp_kernel<NWQ_NARROW>, AL_WARPS * 32, 0));
			else KG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nw_warp_kernel<NW_RS_MAXC>, AL_WARPS * 32, 0));
			int grid = (int)std::min<size_t>((size_t)db->sm_count * (size_t)std::max(per_sm, 1), ((size_t)n + AL_WARPS - 1) / AL_WARPS);
			if (lay.stride * (size_t)grid * AL_WARPS > scr.cap) {
				size_t freeb = 0, totalb = 0;
				cudaMemGetInfo(&freeb, &totalb);
				while (grid > 1 && lay.stride * (size_t)grid * AL_WARPS > freeb / 2 + scr.cap) grid = (grid + 1) / 2;
			}
			if (scr.reserve(lay.stride * (size_t)grid * AL_WARPS)) return -1;
			KG_CUDA(cudaMemsetAsync(ctr + A_WORK, 0, 8, st));
			if (c == 2) nw_warp_kernel<NWQ_NARROW><<<grid, AL_WARPS * 32, 0, st>>>(P.pen, db->tix, probs, ord, n, qbase, qshift, res, status_out, (uint8_t *)scr.p, lay, ctr);
			else nw_warp_kernel<NW_RS_MAXC><<<grid, AL_WARPS * 32, 0, st>>>(P.pen, db->tix, probs, ord, n, qbase, qshift, res, status_out, (uint8_t *)scr.p, lay, ctr);
		}
		if (rc) return -1;
		++*launches;
	}
	return 0;
}

extern "C" int kmagpu_align_run(kmagpu_db *db, const kmagpu_params *prm, int want_cand, kmagpu_align_stats *stats) {
	if (!db || !prm) { kmagpu_set_error("null argument"); return -1; }
	if (!db->d_tslots) { kmagpu_set_error("database has no alignment index (.seq.b / .length.b missing)"); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	AlignBatch &b = db->aln;
	const int n = (int)b.nreads;
	const int DB = db->info.DB_size;
	if (stats) memset(stats, 0, sizeof(*stats));
	b.out_bytes = 0; b.ntasks = 0; b.ran = true; b.want_cand = want_cand != 0;
	b.h_cand.clear();
	b.h_scores.assign(2 * (size_t)DB, 0);
	if (n == 0) return 0;
	const AlnParams P = make_params(db, prm);
	const int ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
	if (b.d_reads.reserve(sizeof(AlnRead) * (size_t)n) || b.d_sz.reserve(4 * (size_t)(4 * n + 8)) ||
	    b.d_partial.reserve(4 * (size_t)(ntiles + 2)) || b.d_ctr.reserve(8 * A_N) || b.d_scores.reserve(16 * (size_t)DB) ||
	    b.d_recsize.reserve(4 * (size_t)(2 * n + 4)) || b.d_res.reserve(sizeof(AlnRes) * (size_t)n)) return -1;
	uint32_t *slab_sz = (uint32_t *)b.d_sz.p, *slab_off = slab_sz + n + 1, *task_sz = slab_off + n + 1, *task_off = task_sz + n + 1;
	uint32_t *partial = (uint32_t *)b.d_partial.p;
	unsigned long long *ctr = (unsigned long long *)b.d_ctr.p;
	unsigned long long *as = (unsigned long long *)b.d_scores.p, *uas = as + DB;
	AlnRead *reads = (AlnRead *)b.d_reads.p;
	int launches = 0;
	unsigned long long h[A_N];
	cudaStream_t st = db->stream;

	KG_CUDA(cudaMemsetAsync(ctr, 0, 8 * A_N, st));
	KG_CUDA(cudaMemsetAsync(as, 0, 16 * (size_t)DB, st));
	KG_CUDA(cudaEventRecord(db->ev[2], st));
	aln_sizes_kernel<<<(n + 255) / 256, 256, 0, st>>>(b.in, (const uint32_t *)b.d_off.p, n, P.k, reads, slab_sz, task_sz, ctr);
	kg_exscan(slab_sz, n, slab_off, partial, ctr + A_SLAB, st);
	kg_exscan(task_sz, n, task_off, partial, ctr + A_TASKS, st);
	launches += 7;
	KG_CUDA(cudaMemcpyAsync(h, ctr, 8 * A_N, cudaMemcpyDeviceToHost, st));
	KG_CUDA(cudaStreamSynchronize(st));
	KG_CUDA(cudaGetLastError());
	if (h[A_BAD]) { kmagpu_set_error("%llu pair records in the stage-2 stream in a form -apm p never writes (mate shorter than k, or strand-undecided pair)", h[A_BAD]); return -1; }
	KG_SCAN_FITS(h[A_SLAB], "the unpacked reads");
	if (h[A_TASKS] >= (1ull << 31)) { kmagpu_set_error("%llu (read, template) pairs in one batch: split it", h[A_TASKS]); return -1; }
	const size_t slab_units = (size_t)h[A_SLAB];
	const int ntasks = (int)h[A_TASKS];
	const int maxq = (int)h[A_MAXQ];
	b.ntasks = ntasks;
	if (b.d_slab.reserve(8 * (slab_units + 4)) || b.d_taskread.reserve(4 * ((size_t)ntasks + 1)) ||
	    b.d_cand.reserve(sizeof(AlnCand) * ((size_t)ntasks + 1)) || b.d_ovf.reserve(4 * ((size_t)ntasks + 1))) return -1;
	KG_CUDA(cudaMemcpyAsync(task_off + n, &ntasks, 4, cudaMemcpyHostToDevice, st));   // the prep kernel reads off[r + 1]
	aln_prep_kernel<<<kg_wave_grid(aln_prep_kernel, 256, db->sm_count), 256, 0, st>>>(b.in, n, reads, slab_off, task_off, (uint64_t *)b.d_slab.p, (int32_t *)b.d_taskread.p, P.k);
	++launches;
	KG_CUDA(cudaEventRecord(db->ev[3], st));
	KG_CUDA(cudaEventRecord(db->ev[4], st));

	if (ntasks) {
		// phase 1: MEMs + chaining per (read, template) pair; per-warp scratch = the MEM table. NW problems go to the queue.
		const ScratchLayout lay = make_layout(2048, 0, 0);
		const bool short_reads = maxq <= AL_SHORT_MAXQ;
		int grid = db->sm_count * (short_reads ? AL_MINB_SHORT : AL_MINB);
		size_t freeb = 0, totalb = 0;
		if (lay.stride * (size_t)grid * AL_WARPS > b.d_scratch.cap) {   // only when the scratch has to grow
			cudaMemGetInfo(&freeb, &totalb);
			while (grid > db->sm_count && lay.stride * (size_t)grid * AL_WARPS > freeb / 2 + b.d_scratch.cap) grid -= db->sm_count;
		}
		if (b.d_scratch.reserve(lay.stride * (size_t)grid * AL_WARPS)) return -1;
		if (b.prob_cap < (size_t)ntasks + 65536) b.prob_cap = (size_t)ntasks + 65536;
		unsigned long long first_ovf = 0;
		KG_CUDA(cudaEventRecord(db->ev[3], st));   // ms_align = the pair kernel(s) + the NW queue kernels; host-side sizing above is in ms_total
		for (int attempt = 0;; ++attempt) {
			if (b.prob_cap >= (1ull << 32)) { kmagpu_set_error("NW problem queue exceeds 2^32 entries; split the batch"); return -1; }
			// order[classes][cap], cells[classes][cap], 2 * cap of sorted keys / slots
			if (b.d_probs.reserve(sizeof(NwProb) * b.prob_cap) || b.d_order.reserve(4 * (2 * NWQ_CLASSES + 2) * b.prob_cap)) return -1;
			NwQueue queue;
			queue.probs = (NwProb *)b.d_probs.p; queue.order = (uint32_t *)b.d_order.p; queue.cells = queue.order + NWQ_CLASSES * b.prob_cap;
			queue.cap = (unsigned)b.prob_cap; queue.ctr = ctr; queue.d8 = P.pen.d8;
			if (!prm->counters)   // production: no statistic counters (stats->mems, index_probes, mem_bases, read_bytes stay 0)
				(short_reads ? kg_launch_pair_fast_short : kg_launch_pair_fast_long)(grid, st, &P, &db->tix, b.in, reads, (const uint64_t *)b.d_slab.p,
					(const int32_t *)b.d_taskread.p, ntasks, nullptr, b.d_cand.p, (uint8_t *)b.d_scratch.p, &lay, ctr, (int32_t *)b.d_ovf.p, &queue);
			else if (short_reads)
				aln_pair_kernel<AL_MINB_SHORT><<<grid, AL_WARPS * 32, 0, st>>>(P, db->tix, b.in, reads, (const uint64_t *)b.d_slab.p,
					(const int32_t *)b.d_taskread.p, ntasks, nullptr, (AlnCand *)b.d_cand.p, (uint8_t *)b.d_scratch.p, lay, ctr,
					(int32_t *)b.d_ovf.p, queue);
			else
				aln_pair_kernel<AL_MINB><<<grid, AL_WARPS * 32, 0, st>>>(P, db->tix, b.in, reads, (const uint64_t *)b.d_slab.p,
					(const int32_t *)b.d_taskread.p, ntasks, nullptr, (AlnCand *)b.d_cand.p, (uint8_t *)b.d_scratch.p, lay, ctr,
					(int32_t *)b.d_ovf.p, queue);
			++launches;
			KG_CUDA(cudaMemcpyAsync(h, ctr, 8 * A_N, cudaMemcpyDeviceToHost, st));
			KG_CUDA(cudaStreamSynchronize(st));
			KG_CUDA(cudaGetLastError());
			int novf = (int)h[A_OVF];
			first_ovf = h[A_OVF];
			// large MEM tables: the same code on few warps, table sized from what the pairs asked for
			for (int round = 0; novf; ++round) {
				if (round == 8) { kmagpu_set_error("%d read/template pairs do not fit the alignment scratch", novf); return -1; }
				const ScratchLayout big = make_layout(std::max<int>(2048, (int)h[A_NEED_MEM]), 0, 0);
				cudaMemGetInfo(&freeb, &totalb);
				int g2 = std::min(db->sm_count, (novf + AL_WARPS - 1) / AL_WARPS);
				while (g2 > 1 && big.stride * (size_t)g2 * AL_WARPS > (freeb + b.d_scratch.cap) / 2) g2 = (g2 + 1) / 2;
				if (b.d_scratch.reserve(big.stride * (size_t)g2 * AL_WARPS)) return -1;
				// the overflow list becomes the task list; d_ovf collects what still does not fit
				KgBuf list2;
				if (list2.reserve(4 * ((size_t)novf + 1))) return -1;
				KG_CUDA(cudaMemcpyAsync(list2.p, b.d_ovf.p, 4 * (size_t)novf, cudaMemcpyDeviceToDevice, st));
				KG_CUDA(cudaMemsetAsync(ctr + A_WORK, 0, 8 * 2, st));   // A_WORK, A_OVF
				KG_CUDA(cudaMemsetAsync(ctr + A_NEED_MEM, 0, 8, st));
				aln_pair_kernel<AL_MINB><<<g2, AL_WARPS * 32, 0, st>>>(P, db->tix, b.in, reads, (const uint64_t *)b.d_slab.p,
					(const int32_t *)b.d_taskread.p, novf, (const int32_t *)list2.p, (AlnCand *)b.d_cand.p,
					(uint8_t *)b.d_scratch.p, big, ctr, (int32_t *)b.d_ovf.p, queue);
				++launches;
				KG_CUDA(cudaMemcpyAsync(h, ctr, 8 * A_N, cudaMemcpyDeviceToHost, st));
				KG_CUDA(cudaStreamSynchronize(st));
				KG_CUDA(cudaGetLastError());
				list2.release();
				novf = (int)h[A_OVF];
			}
			if (h[A_NPROB] <= b.prob_cap) break;
			// the queue was too small: the count is exact now, redo phase 1 (the capacity persists with the handle)
			if (attempt) { kmagpu_set_error("NW problem queue overflow persists (%llu problems)", h[A_NPROB]); return -1; }
			b.prob_cap = (size_t)h[A_NPROB] + (size_t)h[A_NPROB] / 8 + 65536;
			KG_CUDA(cudaMemsetAsync(ctr, 0, 8 * A_N, st));
		}
		h[A_OVF] = first_ovf;
		// phase 2: the queued NW problems add their scores into the candidate rows
		if (h[A_NPROB]) {
			uint32_t *ord = (uint32_t *)b.d_order.p;
			if (nw_queue_run(db, P, (const NwProb *)b.d_probs.p, ord, ord + NWQ_CLASSES * b.prob_cap, ord + 2 * NWQ_CLASSES * b.prob_cap, b.prob_cap,
			                 &h[A_PCLS], (size_t)h[A_NEED_E], (int)h[A_NEED_Q], (const uint8_t *)b.d_slab.p, 3,
			                 (int32_t *)b.d_cand.p, nullptr, ctr, &launches)) return -1;
		}
		KG_CUDA(cudaEventRecord(db->ev[4], st));
		unsigned long long h2[A_N];
		KG_CUDA(cudaMemcpyAsync(h2, ctr, 8 * A_N, cudaMemcpyDeviceToHost, st));
		KG_CUDA(cudaStreamSynchronize(st));
		KG_CUDA(cudaGetLastError());
		if (h2[A_BAD]) { kmagpu_set_error("%llu NW problems did not fit the scratch sized for them", h2[A_BAD]); return -1; }
		for (int i : {A_FULL_CALLS, A_BAND_CALLS, A_FULL_CELLS, A_BAND_CELLS, A_STEPS}) h[i] = h2[i];
	}
	if (b.want_cand && ntasks) {   // per-candidate rows, before the selection compacts them in place
		std::vector<AlnCand> hc((size_t)ntasks);
		std::vector<int32_t> tr((size_t)ntasks);
		KG_CUDA(cudaMemcpyAsync(hc.data(), b.d_cand.p, sizeof(AlnCand) * (size_t)ntasks, cudaMemcpyDeviceToHost, st));
		KG_CUDA(cudaMemcpyAsync(tr.data(), b.d_taskread.p, 4 * (size_t)ntasks, cudaMemcpyDeviceToHost, st));
		KG_CUDA(cudaStreamSynchronize(st));
		b.h_cand.resize(8 * (size_t)ntasks);
		for (int i = 0; i < ntasks; ++i) {
			int32_t *o = &b.h_cand[8 * (size_t)i];
			o[0] = tr[i]; o[1] = hc[i].tmpl; o[2] = hc[i].score; o[3] = hc[i].len; o[4] = hc[i].pos; o[5] = hc[i].match;
			o[6] = hc[i].tGaps; o[7] = hc[i].qGaps;
		}
	}
	uint32_t *recsize = (uint32_t *)b.d_recsize.p, *recoff = recsize + n + 1;
	aln_reduce_kernel<<<(n + 127) / 128, 128, 0, st>>>(P, b.in, reads, n, (AlnCand *)b.d_cand.p, db->tix.meta, as, uas, recsize,
		(AlnRes *)b.d_res.p, ctr);
	kg_exscan(recsize, n, recoff, partial, ctr + A_OUT, st);
	launches += 4;
	unsigned long long h3[A_N];
	KG_CUDA(cudaMemcpyAsync(h3, ctr, 8 * A_N, cudaMemcpyDeviceToHost, st));
	KG_CUDA(cudaStreamSynchronize(st));
	KG_CUDA(cudaGetLastError());
	KG_SCAN_FITS(h3[A_OUT], "the frag_raw stream");   // reads travel at 1 byte per base here, 4x their stage-2 size
	b.out_bytes = (size_t)h3[A_OUT];
	if (b.d_out.reserve(b.out_bytes + 64)) return -1;
	aln_emit_kernel<<<kg_wave_grid(aln_emit_kernel, 256, db->sm_count), 256, 0, st>>>(b.in, reads, n, (const uint64_t *)b.d_slab.p, (const AlnCand *)b.d_cand.p,
		(const AlnRes *)b.d_res.p, recoff, (uint8_t *)b.d_out.p);
	++launches;
	KG_CUDA(cudaEventRecord(db->ev[7], st));
	kg_scores_accumulate(db, as);   // the run-wide sums on the device, for kmagpu_allreduce_scores / ConClave (kmagpu_scores_reset)
	KG_CUDA(cudaMemcpyAsync(b.h_scores.data(), as, 16 * (size_t)DB, cudaMemcpyDeviceToHost, st));
	KG_CUDA(cudaStreamSynchronize(st));
	KG_CUDA(cudaGetLastError());
	if (stats) {
		stats->reads = n; stats->tasks = ntasks; stats->frags = (int64_t)h3[A_FRAGS]; stats->mems = (int64_t)h[A_MEMS];
		stats->nw_full_calls = (int64_t)h[A_FULL_CALLS]; stats->nw_band_calls = (int64_t)h[A_BAND_CALLS];
		stats->nw_full_cells = (int64_t)h[A_FULL_CELLS]; stats->nw_band_cells = (int64_t)h[A_BAND_CELLS];
		stats->nw_steps = (int64_t)h[A_STEPS];
		stats->index_probes = (int64_t)h[A_LOOKUPS]; stats->mem_bases = (int64_t)h[A_MEMBASES]; stats->read_bytes = (int64_t)h[A_READBYTES];
		stats->overflow_tasks = ntasks ? (int64_t)h[A_OVF] : 0;
		cudaEventElapsedTime(&stats->ms_align, db->ev[3], db->ev[4]);
		cudaEventElapsedTime(&stats->ms_total, db->ev[2], db->ev[7]);
		stats->ms_prep = 0; stats->ms_reduce = stats->ms_total - stats->ms_align;   // everything around the pair kernel
		cudaEventElapsedTime(&stats->ms_total, db->ev[2], db->ev[7]);
		stats->launches = launches;
	}
	return 0;
}

extern "C" int kmagpu_align_download(kmagpu_db *db, void *frag_out, size_t out_cap, size_t *out_bytes,
                                     uint64_t *alignment_scores, uint64_t *uniq_alignment_scores,
                                     kmagpu_cand *cand_out, size_t cand_cap, size_t *cand_rows) {
	if (!db) { kmagpu_set_error("null argument"); return -1; }
	AlignBatch &b = db->aln;
	if (!b.ran) { kmagpu_set_error("kmagpu_align_download before kmagpu_align_run"); return -1; }
	if (out_bytes) *out_bytes = b.out_bytes;
	if (cand_rows) *cand_rows = b.h_cand.size() / 8;
	if (frag_out && b.out_bytes > out_cap) { kmagpu_set_error("frag_raw output needs %zu bytes, caller gave %zu", b.out_bytes, out_cap); return -1; }
	if (cand_out && b.h_cand.size() / 8 > cand_cap) { kmagpu_set_error("candidate rows need %zu entries, caller gave %zu", b.h_cand.size() / 8, cand_cap); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	if (b.out_bytes && frag_out) {   // frag_out = NULL: the stream stays in HBM (kmagpu_conclave_from_align), only the score arrays come back
		KG_CUDA(cudaMemcpyAsync(frag_out, b.d_out.p, b.out_bytes, cudaMemcpyDeviceToHost, db->stream));
		KG_CUDA(cudaStreamSynchronize(db->stream));
	}
	const size_t DB = (size_t)db->info.DB_size;
	if (alignment_scores) for (size_t i = 0; i < DB; ++i) alignment_scores[i] += b.h_scores[i];
	if (uniq_alignment_scores) for (size_t i = 0; i < DB; ++i) uniq_alignment_scores[i] += b.h_scores[DB + i];
	if (cand_out && !b.h_cand.empty()) memcpy(cand_out, b.h_cand.data(), 4 * b.h_cand.size());
	return 0;
}

extern "C" int kmagpu_align_batch(kmagpu_db *db, const kmagpu_params *p, const void *stage2, size_t nbytes,
                                  void *frag_out, size_t out_cap, size_t *out_bytes,
                                  uint64_t *alignment_scores, uint64_t *uniq_alignment_scores,
                                  kmagpu_cand *cand_out, size_t cand_cap, size_t *cand_rows, kmagpu_align_stats *stats) {
	int64_t n;
	if (kmagpu_align_upload(db, stage2, nbytes, &n)) return -1;
	if (kmagpu_align_run(db, p, cand_out != nullptr, stats)) return -1;
	if (kmagpu_align_download(db, frag_out, out_cap, out_bytes, alignment_scores, uniq_alignment_scores, cand_out, cand_cap, cand_rows)) return -1;
	if (stats) cudaEventElapsedTime(&stats->ms_h2d, db->ev[5], db->ev[6]);
	return 0;
}


// ---------------------------------------------------------------- traceback batch (host)

// the alignment part of assemble_KMA's inner loop over n fragment records that sit in HBM (din, record offsets doff[n + 1]);
// out == NULL: no row output (base counts / statistics only)
static int trace_core(kmagpu_db *db, const kmagpu_params *prm, const uint8_t *din, const uint32_t *doff, int n, void *out, size_t out_cap,
                      size_t *out_bytes, kmagpu_align_stats *stats) {
	const AlnParams P = make_params(db, prm);
	cudaStream_t st = db->stream;
	const int ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
	TraceBatch &tb = db->trc;   // working buffers persist across calls
	KgBuf &d_recs = tb.d_recs, &d_sz = tb.d_sz, &d_partial = tb.d_partial, &d_ctr = tb.d_ctr, &d_slab = tb.d_slab, &d_rows = tb.d_rows,
	      &d_outs = tb.d_outs, &d_ovf = tb.d_ovf, &d_out = tb.d_out;
	if (d_recs.reserve(sizeof(TrRec) * (size_t)n) ||
	    d_sz.reserve(4 * (size_t)(6 * n + 12)) || d_partial.reserve(4 * (size_t)(ntiles + 2)) || d_ctr.reserve(8 * A_N) ||
	    d_outs.reserve(sizeof(TrOut) * (size_t)n) || d_ovf.reserve(4 * ((size_t)n + 1))) return -1;
	uint32_t *slab_sz = (uint32_t *)d_sz.p, *slab_off = slab_sz + n + 1, *row_sz = slab_off + n + 1, *row_off = row_sz + n + 1,
	         *osz = row_off + n + 1, *ooff = osz + n + 1;
	unsigned long long *ctr = (unsigned long long *)d_ctr.p;
	unsigned long long h[A_N];
	int launches = 0;
	KG_CUDA(cudaMemsetAsync(ctr, 0, 8 * A_N, st));
	KG_CUDA(cudaEventRecord(db->ev[2], st));
	tr_sizes_kernel<<<kg_wave_grid(tr_sizes_kernel, 256, db->sm_count), 256, 0, st>>>(din, doff, n, db->info.DB_size,
		(TrRec *)d_recs.p, slab_sz, row_sz, ctr);
	kg_exscan(slab_sz, n, slab_off, (uint32_t *)d_partial.p, ctr + A_SLAB, st);
	kg_exscan(row_sz, n, row_off, (uint32_t *)d_partial.p, ctr + A_TASKS, st);
	launches += 7;
	KG_CUDA(cudaMemcpyAsync(h, ctr, 8 * A_N, cudaMemcpyDeviceToHost, st));
	KG_CUDA(cudaStreamSynchronize(st));
	KG_CUDA(cudaGetLastError());
	if (h[A_BAD]) { kmagpu_set_error("%llu fragment records name a template outside the database", h[A_BAD]); return -1; }
	const int maxq = (int)h[A_MAXQ];
	KG_SCAN_FITS(h[A_SLAB], "the unpacked fragments");
	KG_SCAN_FITS(h[A_TASKS], "the alignment row pool");
	if (d_slab.reserve(8 * ((size_t)h[A_SLAB] + 4)) || d_rows.reserve(8 * ((size_t)h[A_TASKS] + 4))) return -1;
	tr_prep_kernel<<<kg_wave_grid(tr_prep_kernel, 256, db->sm_count), 256, 0, st>>>(din, n, (TrRec *)d_recs.p, slab_off, row_off, (uint64_t *)d_slab.p);
	++launches;
	const int q_cap = std::min(std::max(maxq + 64, 256), 1 << 20);
	const size_t e_cap = std::min<size_t>(std::max<size_t>(2 * (size_t)maxq * (size_t)maxq + 65536, 65536), 4u << 20);
	ScratchLayout lay = make_layout(2048, q_cap, e_cap);
	AlignBatch &b = db->aln;   // the per-warp scratch is shared with the alignment pass
	int grid = db->sm_count * TR_MINB;
	size_t freeb = 0, totalb = 0;
	if (lay.stride * (size_t)grid * AL_WARPS > b.d_scratch.cap) {
		cudaMemGetInfo(&freeb, &totalb);
		while (grid > db->sm_count && lay.stride * (size_t)grid * AL_WARPS > freeb / 2 + b.d_scratch.cap) grid -= db->sm_count;
	}
	if (b.d_scratch.reserve(lay.stride * (size_t)grid * AL_WARPS)) return -1;
	KG_CUDA(cudaEventRecord(db->ev[3], st));
	tr_task_kernel<<<grid, AL_WARPS * 32, 0, st>>>(P, db->tix, (const TrRec *)d_recs.p, (const uint64_t *)d_slab.p, n, nullptr,
		(TrOut *)d_outs.p, (uint8_t *)d_rows.p, (uint8_t *)b.d_scratch.p, lay, ctr, (int32_t *)d_ovf.p);
	KG_CUDA(cudaEventRecord(db->ev[4], st));
	++launches;
	KG_CUDA(cudaMemcpyAsync(h, ctr, 8 * A_N, cudaMemcpyDeviceToHost, st));
	KG_CUDA(cudaStreamSynchronize(st));
	KG_CUDA(cudaGetLastError());
	int novf = (int)h[A_OVF];
	const unsigned long long first_ovf = h[A_OVF];
	for (int round = 0; novf; ++round) {   // large-scratch path, as in the alignment pass
		if (round == 8) { kmagpu_set_error("%d fragment records do not fit the alignment scratch", novf); return -1; }
		const ScratchLayout big = make_layout(std::max<int>(2048, (int)h[A_NEED_MEM]), std::max<int>(q_cap, (int)h[A_NEED_Q]),
		                                      std::max<size_t>(e_cap, (size_t)h[A_NEED_E]));
		cudaMemGetInfo(&freeb, &totalb);
		int g2 = std::min(db->sm_count, (novf + AL_WARPS - 1) / AL_WARPS);
		while (g2 > 1 && big.stride * (size_t)g2 * AL_WARPS > (freeb + b.d_scratch.cap) / 2) g2 = (g2 + 1) / 2;
		if (b.d_scratch.reserve(big.stride * (size_t)g2 * AL_WARPS)) return -1;
		KgBuf list2;
		if (list2.reserve(4 * ((size_t)novf + 1))) return -1;
		KG_CUDA(cudaMemcpyAsync(list2.p, d_ovf.p, 4 * (size_t)novf, cudaMemcpyDeviceToDevice, st));
		KG_CUDA(cudaMemsetAsync(ctr + A_WORK, 0, 8 * 5, st));
		tr_task_kernel<<<g2, AL_WARPS * 32, 0, st>>>(P, db->tix, (const TrRec *)d_recs.p, (const uint64_t *)d_slab.p, novf,
			(const int32_t *)list2.p, (TrOut *)d_outs.p, (uint8_t *)d_rows.p, (uint8_t *)b.d_scratch.p, big, ctr, (int32_t *)d_ovf.p);
		++launches;
		KG_CUDA(cudaMemcpyAsync(h, ctr, 8 * A_N, cudaMemcpyDeviceToHost, st));
		KG_CUDA(cudaStreamSynchronize(st));
		KG_CUDA(cudaGetLastError());
		list2.release();
		novf = (int)h[A_OVF];
	}
	if (h[A_BAD]) { kmagpu_set_error("%llu alignments are longer than 3 * read length + 256 columns", h[A_BAD]); return -1; }
	if (prm->matrix) {   // alnToMatPtr (assembly.c:1968) on every accepted alignment
		if (prm->matrix != 1 && prm->matrix != 2) { kmagpu_set_error("matrix mode %d: 1 = alnToMat (template nodes), 2 = alnToMatDense", prm->matrix); return -1; }
		if (!db->image->d_mat && kmagpu_matrix_reset(db)) return -1;
		tr_matrix_kernel<<<kg_wave_grid(tr_matrix_kernel, 256, db->sm_count), 256, 0, st>>>((const TrRec *)d_recs.p, (const TrOut *)d_outs.p, n, (const uint8_t *)d_rows.p,
			db->tix.meta, db->image->d_mat_off, prm->matrix == 2, db->image->d_mat, ctr);
		++launches;
	}
	if (out) {
		tr_outsize_kernel<<<(n + 255) / 256, 256, 0, st>>>((const TrOut *)d_outs.p, n, osz);
		kg_exscan(osz, n, ooff, (uint32_t *)d_partial.p, ctr + A_OUT, st);
		launches += 4;
		unsigned long long h2[A_N];
		KG_CUDA(cudaMemcpyAsync(h2, ctr, 8 * A_N, cudaMemcpyDeviceToHost, st));
		KG_CUDA(cudaStreamSynchronize(st));
		KG_SCAN_FITS(h2[A_OUT], "the traceback output");
		const size_t ob = (size_t)h2[A_OUT];
		if (out_bytes) *out_bytes = ob;
		if (ob > out_cap) { kmagpu_set_error("trace output needs %zu bytes, caller gave %zu", ob, out_cap); return -1; }
		if (d_out.reserve(ob + 64)) return -1;
		tr_emit_kernel<<<kg_wave_grid(tr_emit_kernel, 256, db->sm_count), 256, 0, st>>>((const TrRec *)d_recs.p, (const TrOut *)d_outs.p, n, (const uint8_t *)d_rows.p, ooff,
			(uint8_t *)d_out.p);
		++launches;
		KG_CUDA(cudaEventRecord(db->ev[7], st));
		KG_CUDA(cudaMemcpyAsync(out, d_out.p, ob, cudaMemcpyDeviceToHost, st));
	} else KG_CUDA(cudaEventRecord(db->ev[7], st));
	KG_CUDA(cudaMemcpyAsync(h, ctr, 8 * A_N, cudaMemcpyDeviceToHost, st));   // counters incl. the matrix kernel's
	KG_CUDA(cudaStreamSynchronize(st));
	KG_CUDA(cudaGetLastError());
	if (stats) {
		stats->reads = n; stats->tasks = n; stats->mems = (int64_t)h[A_MEMS];
		stats->nw_full_calls = (int64_t)h[A_FULL_CALLS]; stats->nw_band_calls = (int64_t)h[A_BAND_CALLS];
		stats->nw_full_cells = (int64_t)h[A_FULL_CELLS]; stats->nw_band_cells = (int64_t)h[A_BAND_CELLS];
		stats->overflow_tasks = (int64_t)first_ovf;
		cudaEventElapsedTime(&stats->ms_align, db->ev[3], db->ev[4]);
		cudaEventElapsedTime(&stats->ms_total, db->ev[2], db->ev[7]);
		stats->launches = launches;
	}
	return 0;
}

extern "C" int kmagpu_trace_batch(kmagpu_db *db, const kmagpu_params *prm, const void *frags, size_t nbytes,
                                  void *out, size_t out_cap, size_t *out_bytes, int64_t *nrecords, kmagpu_align_stats *stats) {
	if (!db || !prm || (!frags && nbytes)) { kmagpu_set_error("null argument"); return -1; }
	if (!db->d_tslots) { kmagpu_set_error("database has no alignment index (.seq.b / .length.b missing)"); return -1; }
	if (nbytes >= (1ull << 32) - 64) { kmagpu_set_error("fragment batch of %zu bytes exceeds the 4 GiB per-call limit; split it", nbytes); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	if (stats) memset(stats, 0, sizeof(*stats));
	if (out_bytes) *out_bytes = 0;
	if (nrecords) *nrecords = 0;
	size_t used = 0;
	const int64_t n64 = kmagpu_record_walk(3, frags, nbytes, nullptr, 0, &used);
	if (n64 < 0) return -1;
	const int n = (int)n64;
	if (nrecords) *nrecords = n64;
	if (n == 0) return 0;
	std::vector<uint64_t> off64((size_t)n);
	kmagpu_record_walk(3, frags, nbytes, off64.data(), (size_t)n, &used);
	std::vector<uint32_t> off((size_t)n + 1);
	for (int i = 0; i < n; ++i) off[i] = (uint32_t)off64[i];
	off[n] = (uint32_t)used;
	TraceBatch &t = db->trc;
	if (t.d_in.reserve(used + 64) || t.d_off.reserve(4 * ((size_t)n + 2))) return -1;
	cudaStream_t st = db->stream;
	KG_CUDA(cudaMemcpyAsync(t.d_in.p, frags, used, cudaMemcpyHostToDevice, st));
	KG_CUDA(cudaMemsetAsync((uint8_t *)t.d_in.p + used, 0, 64, st));
	KG_CUDA(cudaMemcpyAsync(t.d_off.p, off.data(), 4 * ((size_t)n + 1), cudaMemcpyHostToDevice, st));
	KG_CUDA(cudaStreamSynchronize(st));   // `off` lives on this stack frame
	return trace_core(db, prm, (const uint8_t *)t.d_in.p, (const uint32_t *)t.d_off.p, n, out, out_cap, out_bytes, stats);
}

// The same on the fragment stream the last kmagpu_conclave_batch left in HBM (no host round trip of the fragments).
extern "C" int kmagpu_trace_from_conclave(kmagpu_db *db, const kmagpu_params *prm, void *out, size_t out_cap, size_t *out_bytes,
                                          int64_t *nrecords, kmagpu_align_stats *stats) {
	if (!db || !prm) { kmagpu_set_error("null argument"); return -1; }
	if (!db->d_tslots) { kmagpu_set_error("database has no alignment index (.seq.b / .length.b missing)"); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	if (stats) memset(stats, 0, sizeof(*stats));
	if (out_bytes) *out_bytes = 0;
	const FragBatch &f = db->frg;
	if (!f.valid) { kmagpu_set_error("kmagpu_trace_from_conclave without a preceding kmagpu_conclave_batch on this handle"); return -1; }
	if (nrecords) *nrecords = f.n;
	if (f.n == 0) return 0;
	return trace_core(db, prm, (const uint8_t *)f.d_out.p, f.off, (int)f.n, out, out_cap, out_bytes, stats);
}

// ---------------------------------------------------------------- base-count matrix: host side

// The matrix belongs to the database image: every handle (kmagpu_db_clone) adds to the same counts with atomics.
extern "C" int kmagpu_matrix_reset(kmagpu_db *db) {
	if (!db) { kmagpu_set_error("null argument"); return -1; }
	if (!db->d_tmeta) { kmagpu_set_error("database has no template sequences"); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	const int DB = db->info.DB_size;
	KgImageRef *img = db->image;
	if (!img->d_mat) {
		std::vector<int64_t> off((size_t)DB + 1, 0);
		for (int t = 2; t <= DB; ++t) off[t] = off[t - 1] + db->lengths[t - 1];
		const size_t entries = 6 * (size_t)off[DB];
		int64_t *d_off = nullptr;
		unsigned int *d_mat = nullptr;
		KG_CUDA(cudaMalloc(&d_off, 8 * ((size_t)DB + 1)));
		KG_CUDA(cudaMemcpy(d_off, off.data(), 8 * ((size_t)DB + 1), cudaMemcpyHostToDevice));
		KG_CUDA(cudaMalloc(&d_mat, 4 * entries + 64));
		img->d_mat_off = d_off; img->mat_entries = entries; img->d_mat = d_mat;
		db->info.device_bytes += 4 * entries;
	}
	KG_CUDA(cudaMemsetAsync(img->d_mat, 0, 4 * img->mat_entries, db->stream));
	KG_CUDA(cudaStreamSynchronize(db->stream));
	return 0;
}

extern "C" int kmagpu_matrix_device(kmagpu_db *db, void **ptr, uint64_t *entries) {
	if (!db || !ptr || !entries) { kmagpu_set_error("null argument"); return -1; }
	if (!db->image->d_mat && kmagpu_matrix_reset(db)) return -1;
	*ptr = db->image->d_mat; *entries = db->image->mat_entries;
	return 0;
}

extern "C" int kmagpu_matrix_download(kmagpu_db *db, int32_t tmpl, uint16_t *counts, size_t cap_entries, size_t *entries) {
	if (!db) { kmagpu_set_error("null argument"); return -1; }
	if (!db->image->d_mat) { kmagpu_set_error("kmagpu_matrix_download before any alignment was added"); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	const int DB = db->info.DB_size;
	if (tmpl < 0 || tmpl >= DB) { kmagpu_set_error("template %d outside the database", tmpl); return -1; }
	size_t first = 0, cnt = db->image->mat_entries;   // template 0 = the whole database
	if (tmpl) {
		for (int t = 1; t < tmpl; ++t) first += 6 * (size_t)db->lengths[t];
		cnt = 6 * (size_t)db->lengths[tmpl];
	}
	if (entries) *entries = cnt;
	if (!counts) return 0;
	if (cnt > cap_entries) { kmagpu_set_error("matrix needs %zu entries, caller gave %zu", cnt, cap_entries); return -1; }
	KgBuf tmp;
	if (tmp.reserve(2 * cnt + 64)) return -1;
	mat_clamp_kernel<<<db->sm_count * 4, 256, 0, db->stream>>>(db->image->d_mat + first, cnt, (uint16_t *)tmp.p);
	cudaError_t e = cudaMemcpyAsync(counts, tmp.p, 2 * cnt, cudaMemcpyDeviceToHost, db->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(db->stream);
	tmp.release();
	if (e != cudaSuccess) { kmagpu_set_error("matrix download: %s", cudaGetErrorString(e)); return -1; }
	return 0;
}

// ---------------------------------------------------------------- stand-alone NW batch

// The problems go through the same queue kernels as the alignment pass's (nw_queue_run): what this entry point times is
// the NW the mapping path runs.
extern "C" int kmagpu_nw_batch(kmagpu_db *db, const kmagpu_params *p, size_t n, const int32_t *prob, const uint8_t *qpool,
                               size_t qbytes, int32_t *out, int32_t *status, int64_t *cells, int64_t *steps, float *ms) {
	if (!db || !p || (n && (!prob || !qpool || !out || !status))) { kmagpu_set_error("null argument"); return -1; }
	if (!db->d_tmeta) { kmagpu_set_error("database has no template sequences"); return -1; }
	if (n >= (1ull << 31)) { kmagpu_set_error("too many NW problems in one call"); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	if (cells) *cells = 0;
	if (steps) *steps = 0;
	if (ms) *ms = 0;
	if (!n) return 0;
	size_t need_e = 65536;
	int need_q = 256;
	const AlnParams P = make_params(db, p);
	std::vector<NwProb> hp(n);
	std::vector<uint32_t> horder(NWQ_CLASSES * n);
	unsigned long long counts[NWQ_CLASSES] = {0, 0, 0, 0};
	for (size_t i = 0; i < n; ++i) {
		const int32_t *pr = prob + 8 * i;
		if (pr[0] <= 0 || pr[0] >= db->info.DB_size || pr[1] < 0 || pr[2] < pr[1] || pr[2] > db->lengths[pr[0]] || pr[4] < 0 ||
		    pr[5] < pr[4] || pr[3] < 0 || (size_t)pr[3] + (size_t)pr[5] > qbytes || pr[6] < -2 || pr[6] > 2 || pr[7] < 0 || pr[7] >= (1 << 23)) {
			kmagpu_set_error("NW problem %zu is out of range", i);
			return -1;
		}
		const int t_l = pr[2] - pr[1], q_l = pr[5] - pr[4];
		NwGeo g;
		const int cls = nwq_class(t_l, q_l, pr[7], P.pen.d8);
		if (cls >= 2 && t_l > 0 && q_l > 0 && nw_geo_init(g, P.pen, t_l, q_l, pr[6], pr[7], NW_RS_MAXC)) {
			need_e = std::max(need_e, g.ebytes() + 256);
			need_q = std::max(need_q, q_l + 64);
		}
		NwProb &q = hp[i];
		q.task = (int32_t)i; q.tmpl = pr[0]; q.t_s = pr[1]; q.t_e = pr[2]; q.q_s = pr[4]; q.q_e = pr[5]; q.kband = (pr[6] + 2) | (pr[7] << 8);
		q.qoff = (uint32_t)pr[3];
		horder[(size_t)cls * n + counts[cls]++] = (uint32_t)i;
	}
	for (int c = 0; c < NWQ_CLASSES; ++c)   // sorted by cells, largest first (the alignment pass sorts on the device)
		if (c < 2) std::stable_sort(horder.begin() + (size_t)c * n, horder.begin() + (size_t)c * n + counts[c], [&](uint32_t x, uint32_t y) {
			const long long bx = hp[x].kband >> 8, by = hp[y].kband >> 8;
			return (long long)(hp[x].t_e - hp[x].t_s) * (bx ? bx + 2 : hp[x].q_e - hp[x].q_s) > (long long)(hp[y].t_e - hp[y].t_s) * (by ? by + 2 : hp[y].q_e - hp[y].q_s); });
	uint8_t *dq = nullptr;
	NwProb *dprob = nullptr;
	uint32_t *dorder = nullptr;
	int32_t *dres = nullptr, *dstat = nullptr;
	unsigned long long *ctr = nullptr;
	KG_CUDA(cudaMalloc(&dq, qbytes + 64));
	KG_CUDA(cudaMalloc(&dprob, sizeof(NwProb) * n));
	KG_CUDA(cudaMalloc(&dorder, 4 * NWQ_CLASSES * n));
	KG_CUDA(cudaMalloc(&dres, 32 * n));
	KG_CUDA(cudaMalloc(&dstat, 4 * n));
	KG_CUDA(cudaMalloc(&ctr, 8 * A_N));
	cudaStream_t st = db->stream;
	KG_CUDA(cudaMemcpyAsync(dq, qpool, qbytes, cudaMemcpyHostToDevice, st));
	KG_CUDA(cudaMemcpyAsync(dprob, hp.data(), sizeof(NwProb) * n, cudaMemcpyHostToDevice, st));
	KG_CUDA(cudaMemcpyAsync(dorder, horder.data(), 4 * NWQ_CLASSES * n, cudaMemcpyHostToDevice, st));
	KG_CUDA(cudaMemsetAsync(dres, 0, 32 * n, st));
	KG_CUDA(cudaMemsetAsync(dstat, 0, 4 * n, st));
	KG_CUDA(cudaMemsetAsync(ctr, 0, 8 * A_N, st));
	KG_CUDA(cudaEventRecord(db->ev[2], st));
	int launches = 0;
	const int rc = nw_queue_run(db, P, dprob, dorder, nullptr, nullptr, n, counts, need_e, need_q, dq, 0, dres, dstat, ctr, &launches);
	KG_CUDA(cudaEventRecord(db->ev[3], st));
	std::vector<int32_t> hres(8 * n);
	unsigned long long hc[A_N];
	if (!rc) {
		KG_CUDA(cudaMemcpyAsync(hres.data(), dres, 32 * n, cudaMemcpyDeviceToHost, st));
		KG_CUDA(cudaMemcpyAsync(status, dstat, 4 * n, cudaMemcpyDeviceToHost, st));
		KG_CUDA(cudaMemcpyAsync(hc, ctr, 8 * A_N, cudaMemcpyDeviceToHost, st));
	}
	KG_CUDA(cudaStreamSynchronize(st));
	cudaFree(dq); cudaFree(dprob); cudaFree(dorder); cudaFree(dres); cudaFree(dstat); cudaFree(ctr);
	if (rc) return -1;
	KG_CUDA(cudaGetLastError());
	for (size_t i = 0; i < n; ++i) {
		const int32_t *r = &hres[8 * i];
		int32_t *o = out + 6 * i;
		o[0] = r[1]; o[1] = r[2]; o[2] = 0; o[3] = r[4]; o[4] = r[5]; o[5] = r[6];   // AlnScore.pos of NW_score itself is 0
	}
	if (cells) *cells = (int64_t)(hc[A_FULL_CELLS] + hc[A_BAND_CELLS]);
	if (steps) *steps = (int64_t)hc[A_STEPS];
	if (ms) cudaEventElapsedTime(ms, db->ev[2], db->ev[3]);
	return 0;
}
#endif   // KG_PAIR_VARIANT_ONLY
