// Stage 2 of KMA in chain mode on the GPU: the long-read default of the reference (no -1t1).
//
// What is computed is save_kmers_chain (savekmers.c:5127-5944) with the default selection functions of kmeranker.c
// (getBestChainTemplates :83, pruneAnkers :372, getBestAnkerScore :398, getTieAnkerScore :480, chooseChain :512,
// mrchain :57), the query segment tree (seqmenttree.c:107-232) and insertKmerBound (qseqs.c:41) for every read of a
// batch of stage-1 records. How it is computed is not the reference's:
//   * one warp per read, reads pulled from an atomic work counter by a persistent grid;
//   * ankers: the 32 lanes look up 256 consecutive k-mer positions at once (8 independent probes per lane in flight);
//     whether a hit opens a new anker only depends on the hit before it (same list, 0 or k missed positions), so
//     anker indices are ballot ranks, anker weights differences of a running warp prefix sum, and no lane ever
//     waits for a sequential scan; the reverse strand is never materialised (its k-mers are the bit-reversed
//     complements of the forward ones, including the reference's k-base shift after an N, savekmers.c:5443);
//   * chaining DP: ankers in order, one lane per template of the anker's list against a per-warp dense
//     {score, last end, seen} row in global scratch (L2 resident); the length-corrected tie score is folded in list
//     order with the reference's double arithmetic only when two template lengths actually differ;
//   * the anker "linked list" of the reference is only ever walked in array order, so pruning and best-anker
//     selection are array reductions (max / last arg-max / count) instead of pointer chases;
//   * back-walks (getBestChainTemplates) and tie ankers run warp-wide over the template lists; the segment tree of
//     emitted query intervals (a few dozen nodes) lives in shared memory and is updated by lane 0 with an explicit
//     stack in place of the reference's recursion;
//   * every read leaves its regions in a pool; a scan orders them by read, a second scan turns record sizes into
//     offsets and a writer kernel emits the stage-2 byte stream (ankers.c:30-50 + 9 bound bytes) in input order.
// Two situations are undefined behaviour in the reference (oracle/orc_chain.c header): a back-walk that leaves the
// anker array and more regions on one read than the segment tree holds. Both are counted and fail the call.
#include "kmagpu_internal.h"
#include "kmagpu_dev.cuh"
#include "kmagpu_seed.cuh"
#include <string.h>
#include <algorithm>

#define KC_WARPS 4
#define KC_CHUNK 256
#define KC_PER_LANE (KC_CHUNK / 32)
#define KC_WORDS 14           // staged u64 words per chunk: (256 + 2 * 32 + 31) / 32 + 2
#define KC_ST 64              // segment tree capacity (ST_CAP of the oracle)
#define FULL 0xffffffffu

enum { C_REGS = 12, C_EWALK = 13, C_ETREE = 14, C_BYTES = 15 };

struct ChainParams {
	int32_t M, MM, U, W1, Wl, exhaustive, minlen;
	int32_t lc;   // -lc (kma.c:694-700): length-corrected anker selection (ankerScoreLen, testExtensionScoreLen, ...)
	int32_t use_proxi, pad;   // -proxi (kma.c:702-718): getChainTemplates = getProxiChainTemplates, chooseChain's proximity test
	double proxi;             // |minFrac|
	unsigned long long *soft; // soft proximity sums of the batch (kmers.c:133-153), NULL = off
	double mrs, coverT, mrc;
};
struct ChainRes { uint32_t reg_off; int32_t nreg; };
struct Region { int32_t read, score, ntmpl, rev, b0, b1; uint32_t pool_off, size; };   // 32 bytes

struct Ank { int *start, *end, *weight, *score; uint32_t *vals; int *slen, *llen; };   // ankers of one strand (SoA, per-warp scratch); score_len / len_len kept under -lc only

struct STree { unsigned start[KC_ST + 2], end[KC_ST + 2], cov[KC_ST + 2]; int b0[KC_ST + 2], b1[KC_ST + 2]; int n, err; };
struct SFrame { int root, state; unsigned pos, right; };

struct ChainScratch {   // per-warp layout inside the scratch allocation
	size_t cap;         // ankers per strand
	size_t D;           // DB_size + 1
	size_t regcap;
	size_t stride;
};

// score of chaining an anker to the previous anker of the same template across `gaps` bases
// (savekmers.c:5524-5552 = kmeranker.c:154-187, mlen == kmersize)
__device__ __forceinline__ int link_score(const ChainParams &p, int k, int gaps, int weight) {
	if (gaps == -k) return weight - (k - 1) * p.M;
	if (gaps == 0) return weight + p.MM;
	if (0 < gaps) {
		int mm, m;
		if (gaps <= 2) { mm = gaps; m = 0; }
		else {
			mm = gaps / k + (gaps % k ? 1 : 0); if (mm < 2) mm = 2;
			m = gaps - mm; if (k < m) m = k; if (mm < m) m = mm;
		}
		const int a = p.W1 + (gaps - 1) * p.U, b = mm * p.MM + m * p.M;
		return weight + (a <= b ? b : a);
	}
	return weight + gaps * p.M - (gaps + 1) * p.U + p.W1;
}

// ---------------------------------------------------------------- segment tree (lane 0 only)

// addSeqmentTrees (seqmenttree.c:107-181), the recursion unrolled over an explicit stack
__device__ __noinline__ unsigned st_add(STree &T, SFrame *fs, int root0, int node) {
	int sp = 0;
	unsigned ret = 0;
	fs[0].root = root0; fs[0].state = 0;
	while (sp >= 0) {
		SFrame &f = fs[sp];
		const int root = f.root;
		int child = -1;
		switch (f.state) {
		case 0:
			if (T.b0[root] >= 0) {
				if (T.start[node] < T.start[root] && T.end[root] < T.end[node]) {
					T.start[root] = T.start[node]; T.end[root] = T.end[node]; T.cov[root] = T.cov[node];
					T.cov[node] = 0; T.b0[root] = -1;
					ret = T.cov[root]; --sp;
					break;
				} else if (T.end[root] < T.end[node]) T.end[root] = T.end[node];
				else if (T.start[node] < T.start[root]) T.start[root] = T.start[node];
				{
					const unsigned pos = T.start[T.b1[root]];
					if (T.end[node] < pos) { f.state = 1; child = T.b0[root]; }
					else if (pos <= T.start[node]) { f.state = 2; child = T.b1[root]; }
					else {   // split: the same node serves both halves
						f.pos = T.start[node];
						T.start[node] = T.end[T.b0[root]] + 1;
						T.cov[node] = T.end[node] - T.start[node];
						f.state = 3; child = T.b1[root];
					}
				}
			} else if (T.end[node] < T.start[root] || T.end[root] < T.start[node]) {   // disjoint leaf: bud
				const int bud = node + 1;
				T.start[bud] = T.start[root]; T.end[bud] = T.end[root]; T.cov[bud] = T.cov[root]; T.b0[bud] = -1;
				if (T.end[node] < T.start[root]) { T.start[root] = T.start[node]; T.b0[root] = node; T.b1[root] = bud; }
				else { T.end[root] = T.end[node]; T.b0[root] = bud; T.b1[root] = node; }
				T.cov[root] += T.cov[node];
				ret = T.cov[root]; --sp;
			} else {   // overlapping leaf: extend
				if (T.start[node] < T.start[root]) T.start[root] = T.start[node];
				if (T.end[root] < T.end[node]) T.end[root] = T.end[node];
				T.cov[node] = 0;
				T.cov[root] = T.end[root] - T.start[root];
				ret = T.cov[root]; --sp;
			}
			break;
		case 1: T.cov[root] = T.cov[T.b1[root]] + ret; ret = T.cov[root]; --sp; break;
		case 2: T.cov[root] = T.cov[T.b0[root]] + ret; ret = T.cov[root]; --sp; break;
		case 3:
			f.right = ret;
			T.start[node] = f.pos;
			T.end[node] = T.end[T.b0[root]];
			T.cov[node] = T.end[node] - T.start[node];
			f.state = 4; child = T.b0[root];
			break;
		default: T.cov[root] = f.right + ret; ret = T.cov[root]; --sp; break;
		}
		if (child >= 0) {
			if (sp + 1 > KC_ST) { T.err = 1; return ret; }
			++sp;
			fs[sp].root = child; fs[sp].state = 0;
		}
	}
	return ret;
}

// growSeqmentTree (seqmenttree.c:183); the reference's resize is undefined behaviour: refuse (returns -1)
__device__ __noinline__ int st_grow(STree &T, SFrame *fs, unsigned start, unsigned end) {
	if (KC_ST <= T.n + 2) return -1;
	if (T.n == 0) {
		T.n = 1; T.start[0] = start; T.end[0] = end; T.cov[0] = end - start; T.b0[0] = T.b1[0] = -1;
		return 0;
	}
	const int node = T.n;
	T.start[node] = start; T.end[node] = end; T.cov[node] = end - start; T.b0[node] = -1;
	T.cov[0] = st_add(T, fs, 0, node);
	if (T.cov[node]) T.n += 2;
	return T.err ? -1 : 0;
}

// queSeqmentTree (seqmenttree.c:211)
__device__ __noinline__ unsigned st_query(const STree &T, SFrame *fs, unsigned start, unsigned end) {
	int sp = 0;
	unsigned sum = 0;
	fs[0].root = 0;
	while (sp >= 0) {
		const int s = fs[sp--].root;
		if (end < T.start[s] || T.end[s] < start) continue;
		if (start <= T.start[s] && T.end[s] <= end) { sum += T.cov[s]; continue; }
		if (T.b0[s] >= 0) {
			if (sp + 2 > KC_ST) return sum;
			fs[++sp].root = T.b0[s];
			fs[++sp].root = T.b1[s];
			continue;
		}
		if (T.start[s] <= start && end <= T.end[s]) sum += end - start;
		else if (T.start[s] <= start && start < T.end[s]) sum += T.end[s] - start;
		else if (T.start[s] < end && end <= T.end[s]) sum += end - T.start[s];
	}
	return sum;
}

// ---------------------------------------------------------------- ankers of one strand (savekmers.c:5227-5448)

struct ChainStats { unsigned lookups, hits, lists, listids; };

__device__ int find_ankers(const KgHashView &hv, const ChainParams &p, const ReadCtx &rc, const int strand, const Ank &A,
                           uint64_t *sw, uint32_t *hits, ChainStats &ws) {
	const unsigned lane = threadIdx.x & 31;
	const unsigned lt = (1u << lane) - 1;
	const int k = hv.kmersize, L = rc.seqlen, npos = L - k + 1, sh = 64 - 2 * k;

	// quick check in this strand's own coordinates: every k-th k-mer of every N-free stretch until the first hit
	bool hit = p.exhaustive != 0;
	for (int seg = 0, s = 0; seg <= rc.nN && !hit; ++seg) {
		const int e = seg < rc.nN ? n_at(rc, seg, strand) : L;
		for (int qb = s; qb < e - k + 1 && !hit; qb += 32 * k) {
			const int q = qb + (int)lane * k;
			const bool act = q < e - k + 1;
			bool hp = false;
			if (act) {
				uint64_t km = fwd32(rc.seq, rc.words, strand ? L - k - q : q) >> sh;
				if (strand) km = rev2(~km) >> sh;
				hp = hash_lookup(hv, km) != KG_MISS;
			}
			const unsigned am = __ballot_sync(FULL, act), hm = __ballot_sync(FULL, hp);
			if (lane == 0) ws.lookups += hm ? __ffs(hm) : __popc(am);
			hit = hm != 0;
		}
		s = e + 1;
	}
	if (!hit) return 0;

	int prevPos = -1, nank = 0, Wcarry = 0, startW = 0;
	uint32_t prevOff = KG_MISS;
	for (int c0 = 0; c0 < npos; c0 += KC_CHUNK) {
		// stage the forward words the chunk needs (k-mers at j, or at j - k on the reverse strand behind an N)
		const int flo = max(0, c0 - k), fhi = min(L, c0 + KC_CHUNK + k - 1) - 1;
		const int w0 = flo >> 5, w1 = fhi >> 5;
		__syncwarp();
#pragma unroll 1
		for (int w = w0 + (int)lane; w <= w1 + 1; w += 32) sw[w - w0] = w < rc.words ? ld_u64u(rc.seq + 8 * (size_t)w) : 0ull;
		__syncwarp();
		// phase 1: gather, in three rounds so that 8 independent probes per lane are in flight (exist -> kv -> chain)
		uint32_t e1[KC_PER_LANE];
		if (rc.nN) {   // reads with N's (rare): validity per position, one probe at a time
#pragma unroll 1
			for (int u = 0; u < KC_PER_LANE; ++u) {
				const int j = c0 + u * 32 + (int)lane;
				int ss = 0;
				uint32_t v = KG_MISS;
				if (j < npos && pos_valid(rc, j, k, 0, &ss)) {
					uint64_t km;
					if (strand == 0) km = kmer_from(sw, w0, j, k);
					else if (ss && j < k) {
						// the reference's shifted reverse cursor (savekmers.c:5443) starts past the end of the read when an
						// N sits in the first k bases: undefined there, zero bits here (oracle/orc_chain.c header), i.e.
						// the complement of T in front of the read
						const uint64_t head = kmer_from(sw, w0, 0, k) >> (2 * (k - j));
						km = rev2(~(head | (~0ull << (2 * j)))) >> sh;
					} else km = rev2(~kmer_from(sw, w0, ss ? j - k : j, k)) >> sh;
					v = hash_lookup(hv, km);
					ws.lookups++;
				}
				hits[u * 32 + lane] = v;
			}
		} else {
		if (hv.mega) {
#pragma unroll
			for (int u = 0; u < KC_PER_LANE; ++u) {
				const int j = c0 + u * 32 + (int)lane;
				e1[u] = KG_MISS;
				if (j < npos) {
					uint64_t km = kmer_from(sw, w0, j, k);
					if (strand) km = rev2(~km) >> sh;
					const uint32_t v = __ldg(hv.exist + km);
					e1[u] = v != 1u ? v : KG_MISS;
					ws.lookups++;
				}
			}
		} else {
			// four independent 16-byte bucket loads in flight per lane, twice (see hash_resolve: the entry answers the probe unless
			// its first key differs and the bucket holds more); eight at once cost the kernel more in spilled registers than the
			// deeper queue gained (19.4 vs 18.2 ms on C3)
#pragma unroll
			for (int h = 0; h < KC_PER_LANE; h += 4) {
				uint32_t key[4];
				uint4 b4[4];
#pragma unroll
				for (int u = 0; u < 4; ++u) {
					const int j = c0 + (h + u) * 32 + (int)lane;
					key[u] = 0; b4[u] = make_uint4(0, 0, 0, 0);
					if (j < npos) {
						uint64_t km = kmer_from(sw, w0, j, k);
						if (strand) km = rev2(~km) >> sh;
						key[u] = (uint32_t)km;
						b4[u] = __ldg(hv.bk + (uint32_t)(km & hv.hmask));
						ws.lookups++;
					}
				}
#pragma unroll
				for (int u = 0; u < 4; ++u) e1[h + u] = hash_resolve(hv, b4[u], key[u]);
			}
		}
#pragma unroll
		for (int u = 0; u < KC_PER_LANE; ++u) hits[u * 32 + lane] = e1[u];
		}
		__syncwarp();
		// phase 2: ankers of the chunk, 32 positions at a time
#pragma unroll 1
		for (int u = 0; u < KC_PER_LANE; ++u) {
			const uint32_t myoff = hits[u * 32 + lane];
			const unsigned hm = __ballot_sync(FULL, myoff != KG_MISS);
			if (!hm) continue;
			const int j = c0 + u * 32 + (int)lane;
			const bool h = myoff != KG_MISS;
			const unsigned below = hm & lt;
			const int pl = below ? 31 - __clz(below) : -1;
			uint32_t pOff = __shfl_sync(FULL, myoff, pl < 0 ? 0 : pl);
			int pPos = j - ((int)lane - pl);
			if (pl < 0) { pOff = prevOff; pPos = prevPos; }
			const int gap = j - pPos - 1;
			const bool cont = h && pPos >= 0 && myoff == pOff && (gap == 0 || gap == k);
			const bool isstart = h && !cont;
			int w = cont ? (gap == 0 ? p.M : k * p.M + p.MM) : 0;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, w, o); if ((int)lane >= o) w += y; }
			w += Wcarry;   // running weight over all continuation hits up to and including this lane
			const unsigned smask = __ballot_sync(FULL, isstart);
			const unsigned sbelow = smask & lt;
			const int sl = sbelow ? 31 - __clz(sbelow) : 0;
			int Wst = __shfl_sync(FULL, w, sl);
			if (!sbelow) Wst = startW;
			if (isstart) {
				const int idx = nank + __popc(sbelow);
				A.start[idx] = j; A.vals[idx] = myoff;
				if (idx > 0) { A.weight[idx - 1] = k * p.M + w - Wst; A.end[idx - 1] = pPos + 1 + k; }
				if (pPos < 0 || myoff != pOff) { ws.lists++; ws.listids += (unsigned)list_len(hv, myoff); }
			}
			ws.hits += h ? 1u : 0u;
			const int hl = 31 - __clz(hm);
			prevPos = c0 + u * 32 + hl;
			prevOff = __shfl_sync(FULL, myoff, hl);
			Wcarry = __shfl_sync(FULL, w, 31);
			if (smask) { startW = __shfl_sync(FULL, w, 31 - __clz(smask)); nank += __popc(smask); }
		}
	}
	if (nank && lane == 0) {
		// the last anker ends at seqlen - gaps (savekmers.c:5329): the scan stops at the first N it cannot pass
		int eL = L;
		for (int i = 0; i < rc.nN; ++i) { const int n = n_at(rc, i, 0); if (n >= L - k) { eL = n; break; } }
		A.weight[nank - 1] = k * p.M + Wcarry - startW;
		A.end[nank - 1] = L - (eL - prevPos);
	}
	__syncwarp();
	return nank;
}

// ---------------------------------------------------------------- getBestChainTemplates (kmeranker.c:83-233)

struct WarpCtx {
	Ank V[2];
	int4 *st;          // per template {score, extendScore, include, -}
	int *bt[2];
	int cnt[2];
	int k;
	int seqlen;               // of the read at hand
	const int32_t *lengths;   // template lengths
};

__device__ __noinline__ int chain_templates_proxi(const KgHashView &hv, const ChainParams &p, WarpCtx &W, const int s, const int src, int *dst,
                               int *count, int *err);

// Walk back from anker `src` of strand s, re-scoring its templates anker by anker until one of them reproduces src's
// score at a chain start. dst[1 .. *count] receives the templates that reach it. Marks walked ankers as used.
// Returns the anker the chain starts at, -1 if no template is left. *err is set when the walk leaves the array.
__device__ __noinline__ int chain_templates(const KgHashView &hv, const ChainParams &p, WarpCtx &W, const int s, const int src, int *dst,
                               int *count, int *err) {
	if (p.use_proxi) return chain_templates_proxi(hv, p, W, s, src, dst, count, err);
	const unsigned lane = threadIdx.x & 31;
	const unsigned lt = (1u << lane) - 1;
	const Ank &V = W.V[s];
	const int k = W.k;
	__syncwarp();
	const uint32_t soff = V.vals[src];
	const int nl0 = list_len(hv, soff);
	bool more = false;
#pragma unroll 1
	for (int i = lane; i < nl0; i += 32) {
		const int t = list_id(hv, soff, i);
		dst[1 + i] = t;
		int4 x = W.st[t];
		x.z = (x.z + 1) & 255;
		W.st[t] = x;
		more |= x.z == 1;
	}
	more = __any_sync(FULL, more);
	__syncwarp();
	// kmerAnkerScore: the anker's score, under -lc the score of its length-corrected best template (ankerScoreLen), which then
	// has to be reproduced by a template of that corrected length (testExtensionScoreLen, kmeranker.c:45)
	const bool lc = p.lc != 0;
	const int bestScore = lc ? V.slen[src] : V.score[src];
	const int target = lc ? V.llen[src] : 1;
	int prev = src;
	for (int node = src; more; --node) {
		if (node < 0) { *err = 1; break; }
		const uint32_t off = V.vals[node];
		const int nl = list_len(hv, off);
		const int start = V.start[node], end = V.end[node], weight = V.weight[node];
		bool used = false, done = false;
#pragma unroll 1
		for (int i = lane; i < nl; i += 32) {
			const int t = list_id(hv, off, i);
			int4 x = W.st[t];
			if (!x.z) continue;
			int score = x.x;
			if (x.y == 0) score = weight;
			else { score += link_score(p, k, x.y - end, weight); used = true; }
			if (bestScore <= score) {
				int open = score;
				if (start) { const int g = p.W1 + (start - 1) * p.U; open = score + (p.Wl < g ? g : p.Wl); }
				if (open == bestScore && (!lc || min(W.seqlen, __ldg(W.lengths + t)) == target)) { score = bestScore; done = true; }
			}
			x.x = score; x.y = start;
			W.st[t] = x;
		}
		used = __any_sync(FULL, used);
		done = __any_sync(FULL, done);
		if (used && lane == 0) V.score[node] = 0;
		if (done) { more = false; prev = node; }
		__syncwarp();
	}
	int j = 0;
	for (int base = 0; base < nl0; base += 32) {
		const int i = base + (int)lane;
		bool keep = false;
		int t = 0;
		if (i < nl0) {
			t = dst[1 + i];
			const int4 x = W.st[t];
			keep = x.z == 1 && bestScore <= x.x;
			if (lc && x.z == 1 && !keep)   // proxiTestBestScoreLen (kmeranker.c:53), proxi == 1.0
				keep = __dmul_rn(__ddiv_rn((double)bestScore, (double)target), (double)min(W.seqlen, __ldg(W.lengths + t))) <= (double)x.x;
			W.st[t] = make_int4(0, 0, 0, 0);
		}
		const unsigned m = __ballot_sync(FULL, keep);
		__syncwarp();
		if (keep) dst[1 + j + __popc(m & lt)] = t;
		j += __popc(m);
		__syncwarp();
	}
	*count = j;
	return j ? prev : -1;
}

// getProxiChainTemplates (kmeranker.c:235-370, bound by -proxi): the walk back scores EVERY template it meets (not only
// those of src's list), in the reference's order -- anker by anker, each list from its end --, stops where one of them
// reproduces src's score at a chain start, and keeps the templates within minFrac of that score that carry no mark of
// the tie path (include[]). Always returns the anker the walk stopped at.
__device__ __noinline__ int chain_templates_proxi(const KgHashView &hv, const ChainParams &p, WarpCtx &W, const int s, const int src, int *dst,
                               int *count, int *err) {
	const unsigned lane = threadIdx.x & 31;
	const unsigned lt = (1u << lane) - 1;
	const Ank &V = W.V[s];
	const int k = W.k;
	__syncwarp();
	const bool lc = p.lc != 0;
	const int bestScore = lc ? V.slen[src] : V.score[src];
	const int target = lc ? V.llen[src] : 1;
	const double proxiScore = __dmul_rn(p.proxi, (double)bestScore);
	int prev = src, n = 0;
	bool more = true;
	for (int node = src; more; --node) {
		if (node < 0) { *err = 1; break; }
		const uint32_t off = V.vals[node];
		const int nl = list_len(hv, off);
		const int start = V.start[node], end = V.end[node], weight = V.weight[node];
		bool used = false, done = false;
#pragma unroll 1
		for (int base = nl - 1; base >= 0; base -= 32) {   // lane l takes list entry base - l: the reference's order
			const int i = base - (int)lane;
			bool isnew = false;
			int t = 0;
			if (i >= 0) {
				t = list_id(hv, off, i);
				int4 x = W.st[t];
				int score = x.x;
				if (x.y == 0) { score = weight; isnew = true; }
				else { score += link_score(p, k, x.y - end, weight); used = true; }
				if (bestScore <= score) {
					int open = score;
					if (start) { const int g = p.W1 + (start - 1) * p.U; open = score + (p.Wl < g ? g : p.Wl); }
					if (open == bestScore && (!lc || min(W.seqlen, __ldg(W.lengths + t)) == target)) { score = bestScore; done = true; }
				}
				x.x = score; x.y = start;
				W.st[t] = x;
			}
			const unsigned nm = __ballot_sync(FULL, isnew);
			if (isnew) dst[1 + n + __popc(nm & lt)] = t;
			n += __popc(nm);
		}
		used = __any_sync(FULL, used);
		done = __any_sync(FULL, done);
		if (used && lane == 0) V.score[node] = 0;
		if (done) { more = false; prev = node; }
		__syncwarp();
	}
	int j = 0;
	for (int base = 0; base < n; base += 32) {
		const int i = base + (int)lane;
		bool keep = false;
		int t = 0;
		if (i < n) {
			t = dst[1 + i];
			const int4 x = W.st[t];
			bool ok = proxiScore <= (double)x.x;   // proxiTestBestScore / ...ScoreLen (kmeranker.c:49-55)
			if (lc && !ok) ok = __dmul_rn(__ddiv_rn(proxiScore, (double)target), (double)min(W.seqlen, __ldg(W.lengths + t))) <= (double)x.x;
			keep = x.z == 0 && ok;
			if (keep && p.soft) atomicAdd(p.soft + t, (unsigned long long)x.x);   // kmeranker.c:357-359
			W.st[t] = make_int4(0, 0, 0, 0);
		}
		const unsigned m = __ballot_sync(FULL, keep);
		__syncwarp();
		if (keep) dst[1 + j + __popc(m & lt)] = t;
		j += __popc(m);
		__syncwarp();
	}
	*count = j;
	return prev;
}

// getBestAnkerScore (kmeranker.c:398) as an array reduction: the LAST anker with the largest non-zero score,
// ties = how many others share it. Returns -1 when every anker is used up.
__device__ __noinline__ int best_anker(const Ank &V, int cnt, unsigned *ties) {
	const unsigned lane = threadIdx.x & 31;
	int best = 0, idx = -1, n = 0;
#pragma unroll 1
	for (int a = lane; a < cnt; a += 32) {
		const int sc = V.score[a];
		if (sc == 0) continue;
		if (idx < 0 || best < sc) { best = sc; idx = a; n = 1; }
		else if (best == sc) { idx = a; ++n; }
	}
#pragma unroll
	for (int o = 16; o; o >>= 1) {
		const int ob = __shfl_xor_sync(FULL, best, o), oi = __shfl_xor_sync(FULL, idx, o), on = __shfl_xor_sync(FULL, n, o);
		if (oi >= 0) {
			if (idx < 0 || best < ob) { best = ob; idx = oi; n = on; }
			else if (best == ob) { idx = max(idx, oi); n += on; }
		}
	}
	*ties = idx >= 0 ? (unsigned)(n - 1) : 0u;
	return idx;
}

// getBestAnkerScoreLen (kmeranker.c:432, -lc): the fold rescales every anker to the length of the best one so far, so it
// is evaluated in array order over the ankers still in play (a ballot finds them, every lane runs the same fold)
__device__ __noinline__ int best_anker_len(const Ank &V, int cnt, unsigned *ties) {
	const unsigned lane = threadIdx.x & 31;
	int best = -1, bS = 0, bL = 1;
	unsigned t = 0;
#pragma unroll 1
	for (int base = 0; base < cnt; base += 32) {
		const int a = base + (int)lane;
		const int sc = a < cnt ? V.score[a] : 0;
		const int sl = sc ? V.slen[a] : 0, ll = sc ? V.llen[a] : 1;
		unsigned m = __ballot_sync(FULL, sc != 0);
		while (m) {
			const int l = __ffs(m) - 1;
			m &= m - 1;
			const int nsl = __shfl_sync(FULL, sl, l), nll = __shfl_sync(FULL, ll, l);
			bool take = false;
			if (best < 0) take = true;
			else {
				double x = (double)nsl;
				if (nll != bL) x = __dmul_rn(__ddiv_rn(x, (double)nll), (double)bL);
				if ((double)bS < x) { take = true; t = 0; }
				else if ((double)bS == x) {
					if (bS < nsl) { take = true; t = 0; }
					else if (bS == nsl) { take = true; ++t; }
				}
			}
			if (take) { best = base + l; bS = nsl; bL = nll; }
		}
	}
	*ties = t;
	return best;
}

// getTieAnkerScore (kmeranker.c:480): nearest anker before src that starts behind `stop` and scores like best;
// getTieAnkerScoreLen (:496, lc): ... whose corrected score and length equal the best one's
__device__ __noinline__ int tie_anker(const Ank &V, int stop, int src, int bestScore, bool lc, int bestLen) {
	const unsigned lane = threadIdx.x & 31;
	if (src < 0 || V.start[src] <= stop) return -1;
	for (int hi = src - 1; hi >= 0; hi -= 32) {
		const int a = hi - (int)lane;   // lane 0 = nearest
		const bool in = a >= 0;
		const int st = in ? V.start[a] : 0;
		const bool out = !in || st <= stop;
		const bool match = !out && (lc ? (V.slen[a] == bestScore && V.llen[a] == bestLen) : V.score[a] == bestScore);
		const unsigned om = __ballot_sync(FULL, out), mm = __ballot_sync(FULL, match);
		const int fo = om ? __ffs(om) - 1 : 32, fm = mm ? __ffs(mm) - 1 : 32;
		if (fm < fo) return hi - fm;
		if (om) return -1;
	}
	return -1;
}

// chooseChain (kmeranker.c:512-592); proxi == 1.0: off
__device__ __noinline__ int choose_chain(int fscore, int fend, int rscore, int rend, int cs, int cs_r, double coverT, double proxi, int *Start, int *Len) {
	int rc, start, end;
	if (proxi == 1.0) rc = rscore < fscore ? 1 : fscore < rscore ? 2 : 3;
	else if (rscore <= fscore) rc = (__dmul_rn(proxi, (double)fscore) <= (double)rscore) ? 3 : 1;   // the other strand within the proximity
	else rc = (__dmul_rn(proxi, (double)rscore) <= (double)fscore) ? 3 : 2;
	if (rc == 1) { start = cs; end = fend; }
	else if (rc == 2) { start = cs_r; end = rend; }
	else if (fend < cs_r) { start = cs; end = fend; rc = 1; }
	else if (rend < cs) { start = cs_r; end = rend; rc = 2; }
	else if (cs <= cs_r && rend <= fend) { start = cs; end = fend; }
	else if (cs_r <= cs && fend <= rend) { start = cs_r; end = rend; }
	else if (rend < fend) {
		const int a = fend - cs, b = rend - cs_r, m = a < b ? a : b;
		start = cs_r;
		if (__dmul_rn(coverT, (double)m) <= (double)((unsigned)rend - (unsigned)cs)) end = fend;
		else { end = rend; rc = 2; }
	} else {
		const int a = fend - cs, b = rend - cs_r, m = a < b ? a : b;
		start = cs;
		if (__dmul_rn(coverT, (double)m) <= (double)((unsigned)fend - (unsigned)cs_r)) end = rend;
		else { end = fend; rc = 1; }
	}
	*Start = start; *Len = end - start;
	return rc;
}

// ---------------------------------------------------------------- the chain kernel

__global__ void __launch_bounds__(KC_WARPS * 32)
chain_kernel(KgHashView hv, ChainParams p, const int32_t *__restrict__ lengths, const uint8_t *__restrict__ in,
             const uint32_t *__restrict__ rec_off, int nreads, ChainRes *__restrict__ res, uint32_t *__restrict__ nregs,
             int32_t *__restrict__ pool, unsigned long long pool_cap, Region *__restrict__ regpool, unsigned long long reg_cap,
             unsigned long long *ctr, uint8_t *scratch, ChainScratch lay) {
	__shared__ uint64_t s_words[KC_WARPS][KC_WORDS];
	__shared__ uint32_t s_hits[KC_WARPS][KC_CHUNK];
	__shared__ STree s_tree[KC_WARPS];
	__shared__ SFrame s_frames[KC_WARPS][KC_ST + 2];

	const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const unsigned lt = (1u << lane) - 1;
	uint64_t *sw = s_words[wid];
	STree &T = s_tree[wid];
	SFrame *fs = s_frames[wid];
	const int k = hv.kmersize;

	WarpCtx W;
	Region *regs;
	{
		uint8_t *base = scratch + lay.stride * ((size_t)blockIdx.x * KC_WARPS + wid);
		W.st = (int4 *)base; base += 16 * lay.D;
		for (int s = 0; s < 2; ++s) { W.bt[s] = (int *)base; base += 4 * (2 * lay.D + 4); }
		regs = (Region *)base; base += sizeof(Region) * lay.regcap;
		for (int s = 0; s < 2; ++s) {
			int *a = (int *)base; base += 28 * lay.cap;
			W.V[s].start = a; W.V[s].end = a + lay.cap; W.V[s].weight = a + 2 * lay.cap; W.V[s].score = a + 3 * lay.cap;
			W.V[s].vals = (uint32_t *)(a + 4 * lay.cap);
			W.V[s].slen = a + 5 * lay.cap; W.V[s].llen = a + 6 * lay.cap;
		}
		W.k = k;
		W.lengths = lengths;
	}
	const bool lc = p.lc != 0;
	ChainStats ws = {0, 0, 0, 0};
	unsigned mapped = 0, words_seen = 0, e_walk = 0, e_tree = 0;

	for (;;) {
		unsigned long long wk = 0;
		if (lane == 0) wk = atomicAdd(&ctr[C_WORK], 1ull);
		wk = __shfl_sync(FULL, wk, 0);
		if (wk >= (unsigned long long)nreads) break;
		const int r = (int)wk;

		ReadCtx rc;
		rc.rec = in + rec_off[r];
		rc.seqlen = (int)ld_u32u(rc.rec);
		rc.words = (int)ld_u32u(rc.rec + 4);
		rc.nN = (int)ld_u32u(rc.rec + 8);
		rc.hdrlen = abs((int)ld_u32u(rc.rec + 12));
		rc.seq = rc.rec + 16;
		rc.N = rc.seq + 8 * (size_t)rc.words;
		words_seen += rc.words;
		const int seqlen = rc.seqlen;
		W.seqlen = seqlen;
		const uint32_t base_size = 28u + 8u * rc.words + 4u * rc.nN + (uint32_t)rc.hdrlen + 9u;

		int nreg = 0;
		int err = 0;
		do {
			if (seqlen < k) break;
			// ---- ankers
#pragma unroll 1
			for (int s = 0; s < 2; ++s) {
				if (lane == 0) { W.V[s].start[0] = 0; W.V[s].end[0] = 0; W.V[s].score[0] = 0; W.V[s].vals[0] = KG_MISS; }
				__syncwarp();
				W.cnt[s] = find_ankers(hv, p, rc, s, W.V[s], sw, s_hits[wid], ws);
			}
			if (!W.cnt[0] && !W.cnt[1]) break;

			// ---- chaining DP over the ankers of each strand (savekmers.c:5457-5640)
			unsigned ties = 0, ties_len = 0;
			int bIdx[2] = {0, 0}, blIdx[2] = {0, 0};
			int btN[2] = {0, 0};
#pragma unroll 1
			for (int s = 0; s < 2; ++s) {
				const Ank &V = W.V[s];
				int *bests = W.bt[s];
				int nb = 0, bi = 0, bScore = 0, bSL = 0;
				int bl = 0, blS = 0, blL = 1;   // -lc: the last best length-corrected anker, its score_len / len_len
				for (int a = 0; a < W.cnt[s]; ++a) {
					const int start = V.start[a], end = V.end[a], weight = V.weight[a];
					const uint32_t off = V.vals[a];
					const int nl = list_len(hv, off);
					int nscore = 0, nsl = 0, nll = 1;
					for (int base = nl - 1; base >= 0; base -= 32) {   // lane l takes list entry base - l: the reference's order
						const int i = base - (int)lane;
						const bool act = i >= 0;
						int t = 0, score = 0, ll = 0;
						bool isnew = false;
						if (act) {
							t = list_id(hv, off, i);
							const int4 x = W.st[t];
							ll = min(seqlen, __ldg(lengths + t));
							if (!x.z) {
								isnew = true;
								if (start) { const int g = p.W1 + (start - 1) * p.U; score = weight + (p.Wl < g ? g : p.Wl); }
								else score = weight;
							} else {
								score = x.x + link_score(p, k, start - x.y, weight);
								if (score < 0) {   // restarting the chain here may be better
									int test = start ? p.W1 + (start - 1) * p.U : 0;
									if (test < p.Wl) test = p.Wl;
									if (score < test + weight) score = test + weight;
								}
							}
							W.st[t] = make_int4(score, end, 1, 0);
						}
						const unsigned nm = __ballot_sync(FULL, isnew);
						if (isnew) bests[1 + nb + __popc(nm & lt)] = t;
						nb += __popc(nm);
						const int cntl = min(32, base + 1);
						// node->score: plain maximum; node->score_len: fold in list order (savekmers.c:5573-5608)
						int mx = act ? score : INT_MIN;
#pragma unroll
						for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, o));
						if (nscore < mx) nscore = mx;
						// every template of the chunk as long as the current pick (or nothing picked yet): the fold is a maximum
						const int ll0 = __shfl_sync(FULL, ll, 0);
						if (__all_sync(FULL, !act || ll == ll0) && (nll == ll0 || nll == 1)) {
							if (nsl < mx) { nsl = mx; nll = ll0; }
						} else
						for (int l = 0; l < cntl; ++l) {
							const int sc = __shfl_sync(FULL, score, l), tl = __shfl_sync(FULL, ll, l);
							bool upd;
							if (nll == tl) upd = nsl < sc;
							else if (nll == 1) upd = 0 < sc;
							else {
								const double sl = __dmul_rn(__ddiv_rn((double)sc, (double)tl), (double)nll);
								upd = (double)nsl < sl || ((double)nsl == sl && nsl < sc);
							}
							if (upd) { nsl = sc; nll = tl; }
						}
					}
					if (lane == 0) V.score[a] = nscore;
					if (lc) {   // savekmers.c:5590-5609; the first anker of a strand is compared with itself
						if (lane == 0) { V.slen[a] = nsl; V.llen[a] = nll; }
						if (a == 0) { blS = nsl; blL = nll; }
						double x = (double)nscore;
						if (nll != blL) x = __dmul_rn(__ddiv_rn(x, (double)nll), (double)blL);
						bool take = false;
						if ((double)blS < x) { take = true; ties_len = 0; }
						else if ((double)blS == x) {
							if (blS < nsl) { take = true; ties_len = 0; }
							else if (blS == nsl) { take = true; ++ties_len; }
						}
						if (take) { bl = a; blS = nsl; blL = nll; }
					}
					if (a == 0) { ++ties; bi = 0; bScore = nscore; bSL = nsl; }   // the first anker ties with itself
					else if (bScore < nscore) { bi = a; bScore = nscore; bSL = nsl; ties = 0; }
					else if (bScore == nscore) {
						if (bSL < nsl) ties = 0; else ++ties;
						bi = a; bSL = nsl;
					}
					__syncwarp();
				}
#pragma unroll 1
				for (int i = lane; i < nb; i += 32) W.st[bests[1 + i]] = make_int4(0, 0, 0, 0);
				__syncwarp();
				bIdx[s] = bi;
				blIdx[s] = bl;
			}
			int scF = W.V[0].score[bIdx[0]], scR = W.V[1].score[bIdx[1]];
			if (scF < k && scR < k) break;
			const int lcF = W.V[0].score[blIdx[0]], lcR = W.V[1].score[blIdx[1]];

			const int V_start[2] = {W.V[0].start[0], W.V[1].start[0]};
			// pruneAnkers (kmeranker.c:372): ankers scoring below k leave the list for good
			__syncwarp();
			for (int s = 0; s < 2; ++s)
#pragma unroll 1
				for (int a = lane; a < max(W.cnt[s], 1); a += 32) if (W.V[s].score[a] < k) W.V[s].score[a] = 0;
			__syncwarp();
			if (scF < k) scF = 0;
			if (scR < k) scR = 0;
			if (lc) {
				// savekmers.c:5657-5664: the length-corrected bests take over. pruneAnkers only unlinks an anker, so the new
				// best keeps its score even below k -- unless it is the plain best of a strand without any anker >= k, whose
				// score was zeroed (:5645-5650)
				const int nF = (!scF && blIdx[0] == bIdx[0]) ? 0 : lcF, nR = (!scR && blIdx[1] == bIdx[1]) ? 0 : lcR;
				__syncwarp();
				if (lane == 0) { W.V[0].score[blIdx[0]] = nF; W.V[1].score[blIdx[1]] = nR; }
				__syncwarp();
				scF = nF; scR = nR;
				bIdx[0] = blIdx[0]; bIdx[1] = blIdx[1];
				ties = ties_len;
			}

			int cs[2] = {-1, -1}, start = 0, len = 0, rcm = 0, tmp;
			if (!scF || !scR) {
				const int s = scF ? 0 : 1;
				tmp = chain_templates(hv, p, W, s, bIdx[s], W.bt[s], &btN[s], &err);
				if (err) break;
				cs[s] = W.V[s].start[max(tmp, 0)];   // the DP guarantees a chain start
				start = cs[s]; len = W.V[s].end[bIdx[s]] - start; rcm = s + 1;
			} else {
				tmp = chain_templates(hv, p, W, 0, bIdx[0], W.bt[0], &btN[0], &err); if (err) break;
				cs[0] = W.V[0].start[max(tmp, 0)];
				tmp = chain_templates(hv, p, W, 1, bIdx[1], W.bt[1], &btN[1], &err); if (err) break;
				cs[1] = W.V[1].start[max(tmp, 0)];
				rcm = choose_chain(scF, W.V[0].end[bIdx[0]], scR, W.V[1].end[bIdx[1]], cs[0], cs[1], p.coverT, p.proxi, &start, &len);
			}
			if (len < p.minlen || max(scF, scR) < k) break;

			if (lane == 0) { T.n = 0; T.err = 0; }
			__syncwarp();
			while (bIdx[0] >= 0 || bIdx[1] >= 0) {
				if (ties) {   // equal ankers further up the read join when they overlap enough (savekmers.c:5701-5781)
#pragma unroll 1
					for (int s = 0; s < 2; ++s) {
						if (!(rcm & (s + 1))) continue;
						int *bl = W.bt[s];
						const Ank &V = W.V[s];
						const int bsScore = lc ? V.slen[bIdx[s]] : V.score[bIdx[s]], bsLen = lc ? V.llen[bIdx[s]] : 0;
						const int stop = start < V_start[s] ? V_start[s] : start;
						int v = bIdx[s];
						while ((v = tie_anker(V, stop, v, bsScore, lc, bsLen)) >= 0) {
							if ((double)((unsigned)V.end[v] - (unsigned)start) < __dmul_rn(p.coverT, (double)len)) break;
#pragma unroll 1
							for (int i = lane; i < btN[s]; i += 32) W.st[bl[1 + i]] = make_int4(0, 0, 1, 0);
							__syncwarp();
							int add = 0;
							chain_templates(hv, p, W, s, v, bl + btN[s], &add, &err);
							if (err) break;
							btN[s] += add;
						}
						if (err) break;
#pragma unroll 1
						for (int i = lane; i < btN[s]; i += 32) W.st[bl[1 + i]] = make_int4(0, 0, 0, 0);
						__syncwarp();
					}
					if (err) break;
				}
				if (p.mrc != 0.0) {   // mrchain (kmeranker.c:57)
#pragma unroll 1
					for (int s = 0; s < 2; ++s) {
						if (!(rcm & (s + 1))) continue;
						const double thr = __dmul_rn(p.mrc, (double)len);
						if (!((double)seqlen < thr)) continue;
						int *bl = W.bt[s];
						int j = 0;
						for (int base = 0; base < btN[s]; base += 32) {
							const int i = base + (int)lane;
							const int t = i < btN[s] ? bl[1 + i] : 0;
							const bool keep = i < btN[s] && thr <= (double)__ldg(lengths + t);
							const unsigned m = __ballot_sync(FULL, keep);
							__syncwarp();
							if (keep) bl[1 + j + __popc(m & lt)] = t;
							j += __popc(m);
							__syncwarp();
						}
						btN[s] = j;
						if (!j) rcm ^= s + 1;
					}
				}

				if (rcm) {
					int grown = 0;
					if (lane == 0) grown = st_grow(T, fs, (unsigned)start, (unsigned)(start + len));
					grown = __shfl_sync(FULL, grown, 0);
					if (grown) { err = 2; break; }
					if ((size_t)nreg >= lay.regcap) { err = 2; break; }
					const int side = (rcm & 1) ? 0 : 1;
					const int b0 = (rcm & 1) ? start : seqlen - W.V[1].end[bIdx[1]], b1 = (rcm & 1) ? start + len : seqlen - start;
					const int nt = (rcm == 3) ? btN[0] + btN[1] : btN[side];
					int score = W.V[side].score[bIdx[side]];
					if (rcm == 3) score = -score;
					unsigned long long po = 0;
					if (lane == 0) po = atomicAdd(&ctr[C_POOL], (unsigned long long)nt);
					po = __shfl_sync(FULL, po, 0);
					if (po + nt <= pool_cap) {
						int32_t *dstp = pool + po;
#pragma unroll 1
						for (int i = lane; i < btN[side]; i += 32) dstp[i] = W.bt[side][1 + i];
						if (rcm == 3) for (int i = lane; i < btN[1]; i += 32) dstp[btN[0] + i] = -W.bt[1][1 + i];
					} else if (lane == 0) atomicAdd(&ctr[C_POOLFAIL], 1ull);
					if (lane == 0) {
						Region g;
						g.read = r; g.score = score; g.ntmpl = nt; g.rev = side; g.b0 = b0; g.b1 = b1;
						g.pool_off = (uint32_t)po; g.size = base_size + 4u * (uint32_t)nt;
						regs[nreg] = g;
						W.V[side].score[bIdx[side]] = 0;
						if (rcm == 3) W.V[1].score[bIdx[1]] = 0;
					}
					++nreg;
					btN[side] = 0;
					if (rcm == 3) btN[1] = 0;
					__syncwarp();
				}

				// next chain of either strand (savekmers.c:5838-5924)
				ties = 0; rcm = 0;
#pragma unroll 1
				for (int s = 0; s < 2 && !err; ++s) {
					if (bIdx[s] < 0) continue;
					const Ank &V = W.V[s];
					bool first = true;
					for (;;) {
						int b = bIdx[s];
						if (!first) {
							if (!(b >= 0 && V.score[b] == 0)) break;
							bIdx[s] = b = lc ? best_anker_len(V, W.cnt[s], &ties) : best_anker(V, W.cnt[s], &ties);
							if (b < 0) break;
						}
						const int bsc = V.score[b];
						const bool ok_score = first ? bsc != 0 : k < bsc;
						first = false;
						tmp = -1;
						if (ok_score) tmp = chain_templates(hv, p, W, s, b, W.bt[s], &btN[s], &err);
						if (err) break;
						if (tmp >= 0) {
							cs[s] = V.start[tmp];
							const int bend = V.end[b];
							unsigned cover = 0;
							if (lane == 0 && T.n) cover = st_query(T, fs, (unsigned)cs[s], (unsigned)bend);
							cover = __shfl_sync(FULL, cover, 0);
							len = bend - cs[s];
							if (p.minlen <= len && (double)cover <= __dmul_rn(p.coverT, (double)len) &&
							    __dmul_rn(p.mrs, (double)len) <= (double)bsc) rcm |= s + 1;
							else { if (lane == 0) V.score[b] = 0; __syncwarp(); }
						} else { if (lane == 0) V.score[b] = 0; __syncwarp(); }
					}
				}
				if (err) break;
				if (bIdx[0] < 0 && bIdx[1] < 0) break;
				if (bIdx[0] >= 0 && bIdx[1] >= 0)
					rcm = choose_chain(W.V[0].score[bIdx[0]], W.V[0].end[bIdx[0]], W.V[1].score[bIdx[1]], W.V[1].end[bIdx[1]],
					                   cs[0], cs[1], p.coverT, p.proxi, &start, &len);
				else if (bIdx[0] >= 0) { rcm = 1; start = cs[0]; len = W.V[0].end[bIdx[0]] - start; }
				else { rcm = 2; start = cs[1]; len = W.V[1].end[bIdx[1]] - start; }
			}
		} while (0);

		if (err) {
			// the per-template rows may be dirty: wipe them before the next read
#pragma unroll 1
			for (size_t i = lane; i < lay.D; i += 32) W.st[i] = make_int4(0, 0, 0, 0);
			if (err == 1) ++e_walk; else ++e_tree;
			nreg = 0;
		}
		// hand the regions over
		unsigned long long ro = 0;
		if (lane == 0 && nreg) ro = atomicAdd(&ctr[C_REGS], (unsigned long long)nreg);
		ro = __shfl_sync(FULL, ro, 0);
		__syncwarp();
		if (nreg) {
			if (ro + nreg <= reg_cap) for (int i = lane; i < nreg; i += 32) regpool[ro + i] = regs[i];
			else if (lane == 0) atomicAdd(&ctr[C_POOLFAIL], 1ull);
			++mapped;
		}
		if (lane == 0) { res[r].reg_off = (uint32_t)ro; res[r].nreg = nreg; nregs[r] = (uint32_t)nreg; }
		__syncwarp();
	}
	for (int o = 16; o; o >>= 1) {
		ws.lookups += __shfl_xor_sync(FULL, ws.lookups, o);
		ws.hits += __shfl_xor_sync(FULL, ws.hits, o);
		ws.lists += __shfl_xor_sync(FULL, ws.lists, o);
		ws.listids += __shfl_xor_sync(FULL, ws.listids, o);
	}
	if (lane == 0) {
		atomicAdd(&ctr[C_LOOKUPS], (unsigned long long)ws.lookups);
		atomicAdd(&ctr[C_HITS], (unsigned long long)ws.hits);
		atomicAdd(&ctr[C_LISTS], (unsigned long long)ws.lists);
		atomicAdd(&ctr[C_LISTIDS], (unsigned long long)ws.listids);
		atomicAdd(&ctr[C_MAPPED], (unsigned long long)mapped);
		atomicAdd(&ctr[C_WORDS], (unsigned long long)words_seen);
		if (e_walk) atomicAdd(&ctr[C_EWALK], (unsigned long long)e_walk);
		if (e_tree) atomicAdd(&ctr[C_ETREE], (unsigned long long)e_tree);
	}
}

// regions in read order: ordered[regbase[r] + i] = regpool[res[r].reg_off + i], sizes for the second scan
__global__ void __launch_bounds__(256) chain_order_kernel(const ChainRes *__restrict__ res, const uint32_t *__restrict__ regbase,
		int nreads, const Region *__restrict__ regpool, Region *__restrict__ ordered, uint32_t *__restrict__ rsize) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= nreads) return;
	const ChainRes c = res[r];
	for (int i = 0; i < c.nreg; ++i) {
		const Region g = regpool[c.reg_off + i];
		ordered[regbase[r] + i] = g;
		rsize[regbase[r] + i] = g.size;
	}
}

// one warp per region: the stage-2 record of print_ankers (ankers.c:30-50) whose name carries the query bounds
// (insertKmerBound, qseqs.c:41: a zero byte and two ints)
__global__ void __launch_bounds__(256) chain_emit_kernel(const uint8_t *__restrict__ in, const uint32_t *__restrict__ rec_off,
		const Region *__restrict__ regs, int nregs, const uint32_t *__restrict__ out_off, const int32_t *__restrict__ pool,
		uint8_t *__restrict__ out) {
	const unsigned lane = threadIdx.x & 31;
	const int warps = (gridDim.x * blockDim.x) >> 5;
	for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < nregs; g += warps) {
		const Region rg = regs[g];
		const uint8_t *rec = in + rec_off[rg.read];
		const int seqlen = (int)ld_u32u(rec), words = (int)ld_u32u(rec + 4), nN = (int)ld_u32u(rec + 8);
		const int hdrlen = abs((int)ld_u32u(rec + 12));
		const uint8_t *seq = rec + 16, *N = seq + 8 * (size_t)words, *hdr = N + 4 * (size_t)nN;
		uint8_t *o = out + out_off[g];
		const bool rev = rg.rev != 0;
		if (lane < 7) {
			const int32_t h = lane == 0 ? seqlen : lane == 1 ? words : lane == 2 ? nN : lane == 3 ? rg.score
			                : lane == 4 ? rg.ntmpl : lane == 5 ? hdrlen + 9 : 0;
			st_u32b(o + 4 * lane, (uint32_t)h);
		}
		o += 28;
#pragma unroll 1
		for (int w = lane; w < words; w += 32) {
			uint64_t x;
			if (!rev) x = ld_u64u(seq + 8 * (size_t)w);
			else {
				x = rev2(~fwd32(seq, words, seqlen - 32 * (w + 1)));
				const int c = seqlen - 32 * w;
				if (c < 32) x &= ~0ull << (64 - 2 * c);
			}
			st_u32b(o + 8 * (size_t)w, (uint32_t)x);
			st_u32b(o + 8 * (size_t)w + 4, (uint32_t)(x >> 32));
		}
		o += 8 * (size_t)words;
#pragma unroll 1
		for (int i = lane; i < nN; i += 32) {
			const uint32_t v = rev ? (uint32_t)(seqlen - 1 - (int)ld_u32u(N + 4 * (size_t)(nN - 1 - i))) : ld_u32u(N + 4 * (size_t)i);
			st_u32b(o + 4 * (size_t)i, v);
		}
		o += 4 * (size_t)nN;
#pragma unroll 1
		for (int i = lane; i < rg.ntmpl; i += 32) st_u32b(o + 4 * (size_t)i, (uint32_t)pool[rg.pool_off + i]);
		o += 4 * (size_t)rg.ntmpl;
#pragma unroll 1
		for (int i = lane; i < hdrlen; i += 32) o[i] = hdr[i];
		o += hdrlen;
		if (lane == 0) { o[0] = 0; st_u32b(o + 1, (uint32_t)rg.b0); st_u32b(o + 5, (uint32_t)rg.b1); }
	}
}

// ---------------------------------------------------------------- host side

// kmerScan = save_kmers_chain: called by kmagpu_seed_run when params->kmerscan == 1
int kg_chain_run(kmagpu_db *db, const kmagpu_params *prm, kmagpu_seed_stats *stats) {
	SeedBatch &b = db->seed;
	const int n = (int)b.nreads;
	if (!db->d_lengths) { kmagpu_set_error("chain mode needs template lengths (.length.b missing)"); return -1; }
	if (b.npairs) { kmagpu_set_error("chain mode takes single reads (the reference maps pairs with the pair functions)"); return -1; }
	ChainParams cp;
	memset(&cp, 0, sizeof(cp));
	cp.M = prm->M; cp.MM = prm->MM; cp.U = prm->U; cp.W1 = prm->W1; cp.Wl = prm->Wl; cp.exhaustive = prm->exhaustive;
	cp.minlen = prm->minlen; cp.mrs = prm->scoreT; cp.coverT = prm->coverT; cp.mrc = prm->mrc;
	cp.lc = prm->lc != 0;
	cp.proxi = fabs(prm->minFrac);   // stage 2 sees |minFrac| (kma.c:1605)
	cp.use_proxi = cp.proxi != 1.0;
	const bool soft = prm->minFrac < 0 && cp.use_proxi && db->image->d_soft;
	if (prm->minFrac < 0 && cp.use_proxi && !soft) { kmagpu_set_error("soft proximity (minFrac < 0 in stage 2) needs kmagpu_softproxi_reset first"); return -1; }
	if (soft) {
		if (b.d_soft.reserve(8 * (size_t)db->info.DB_size)) return -1;
		cp.soft = (unsigned long long *)b.d_soft.p;
	}

	const int ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
	if (b.d_res.reserve(sizeof(ChainRes) * (size_t)n) || b.d_recoff.reserve(4 * (size_t)(2 * n + 2)) ||
	    b.d_ctr.reserve(8 * C_N) || b.d_partial.reserve(4 * (size_t)(ntiles + 2))) return -1;

	// per-warp scratch: per-template rows, the two template lists, region staging, the anker arrays of both strands
	ChainScratch lay;
	lay.D = (size_t)db->info.DB_size + 1;
	lay.cap = (((size_t)std::max(b.max_seqlen, 64) + 8) + 3) & ~(size_t)3;
	lay.regcap = lay.cap / 8 + 64;
	lay.stride = (16 * lay.D + 2 * 4 * (2 * lay.D + 4) + sizeof(Region) * lay.regcap + 2 * 28 * lay.cap + 255) & ~(size_t)255;
	int grid = db->sm_count * 5;
	const size_t budget = (size_t)12 << 30;
	while (grid > db->sm_count && lay.stride * (size_t)grid * KC_WARPS > budget) grid -= db->sm_count;
	grid = (int)std::min<size_t>((size_t)grid, ((size_t)n + KC_WARPS - 1) / KC_WARPS);
	const size_t sbytes = lay.stride * (size_t)grid * KC_WARPS;
	if (b.d_chain.cap < sbytes || !b.d_chain.p) {
		if (b.d_chain.reserve(sbytes)) return -1;
	}
	// the per-template rows must start clean; the layout moves with the batch, so clear every time it is (re)laid
	if (b.chain_stride != lay.stride || b.chain_grid < grid) {
		KG_CUDA(cudaMemsetAsync(b.d_chain.p, 0, sbytes, db->stream));
		b.chain_stride = lay.stride; b.chain_grid = grid;
	}

	if (b.pool_cap < (size_t)n * 32 + 1024) b.pool_cap = (size_t)n * 32 + 1024;
	if (b.reg_cap < (size_t)n * 2 + 1024) b.reg_cap = (size_t)n * 2 + 1024;
	uint32_t *nregs = (uint32_t *)b.d_recoff.p, *regbase = nregs + n + 1;
	uint32_t *partial = (uint32_t *)b.d_partial.p;
	unsigned long long *ctr = (unsigned long long *)b.d_ctr.p;
	int launches = 0;
	for (int attempt = 0;; ++attempt) {
		if (b.d_pool.reserve(4 * b.pool_cap) || b.d_regpool.reserve(sizeof(Region) * b.reg_cap)) return -1;
		KG_CUDA(cudaMemsetAsync(ctr, 0, 8 * C_N, db->stream));
		if (soft) KG_CUDA(cudaMemsetAsync(b.d_soft.p, 0, 8 * (size_t)db->info.DB_size, db->stream));   // per attempt: a pool overflow redoes the batch
		KG_CUDA(cudaEventRecord(db->ev[2], db->stream));
		chain_kernel<<<grid, KC_WARPS * 32, 0, db->stream>>>(db->hv, cp, db->d_lengths, (const uint8_t *)b.d_in.p,
			(const uint32_t *)b.d_off.p, n, (ChainRes *)b.d_res.p, nregs, (int32_t *)b.d_pool.p, (unsigned long long)b.pool_cap,
			(Region *)b.d_regpool.p, (unsigned long long)b.reg_cap, ctr, (uint8_t *)b.d_chain.p, lay);
		KG_CUDA(cudaEventRecord(db->ev[3], db->stream));
		kg_exscan(nregs, n, regbase, partial, ctr + C_TOTAL, db->stream);
		launches += 4;
		unsigned long long h[C_N];
		KG_CUDA(cudaMemcpyAsync(h, ctr, 8 * C_N, cudaMemcpyDeviceToHost, db->stream));
		KG_CUDA(cudaStreamSynchronize(db->stream));
		KG_CUDA(cudaGetLastError());
		if (h[C_EWALK] || h[C_ETREE]) {
			kmagpu_set_error("chain mode: %llu read(s) walk below their first anker and %llu exceed the segment tree: "
			                 "undefined behaviour in the reference (kmeranker.c:83, seqmenttree.c:53)", h[C_EWALK], h[C_ETREE]);
			b.chain_stride = 0;
			return -1;
		}
		if (h[C_POOLFAIL]) {
			if (attempt > 4) { kmagpu_set_error("chain pools overflow persists"); return -1; }
			b.pool_cap = std::max(b.pool_cap, (size_t)h[C_POOL] + 1024);
			b.reg_cap = std::max(b.reg_cap, (size_t)h[C_REGS] + 1024);
			continue;
		}
		if (soft) kg_softproxi_accumulate(db, (const unsigned long long *)b.d_soft.p);
		const size_t NR = (size_t)h[C_TOTAL];
		const int rtiles = (int)((NR + SCAN_TILE - 1) / SCAN_TILE);
		b.out_nrec = (int64_t)NR;
		b.out_bytes = 0;
		if (NR) {
			if (b.d_regs.reserve(sizeof(Region) * NR) || b.d_rsize.reserve(4 * (2 * NR + 2)) ||
			    b.d_partial2.reserve(4 * (size_t)(rtiles + 2))) return -1;
			uint32_t *rsize = (uint32_t *)b.d_rsize.p, *roff = rsize + NR + 1;
			chain_order_kernel<<<(n + 255) / 256, 256, 0, db->stream>>>((const ChainRes *)b.d_res.p, regbase, n,
				(const Region *)b.d_regpool.p, (Region *)b.d_regs.p, rsize);
			kg_exscan(rsize, (int)NR, roff, (uint32_t *)b.d_partial2.p, ctr + C_BYTES, db->stream);
			unsigned long long total = 0;
			KG_CUDA(cudaMemcpyAsync(&total, ctr + C_BYTES, 8, cudaMemcpyDeviceToHost, db->stream));
			KG_CUDA(cudaStreamSynchronize(db->stream));
			if (total >= (1ull << 32)) { kmagpu_set_error("chain mode output of %llu bytes exceeds 4 GiB per call; split the batch", total); return -1; }
			b.out_bytes = (size_t)total;
			if (b.d_out.reserve(b.out_bytes + 64)) return -1;
			chain_emit_kernel<<<kg_wave_grid(chain_emit_kernel, 256, db->sm_count), 256, 0, db->stream>>>((const uint8_t *)b.d_in.p, (const uint32_t *)b.d_off.p,
				(const Region *)b.d_regs.p, (int)NR, roff, (const int32_t *)b.d_pool.p, (uint8_t *)b.d_out.p);
			launches += 5;
			b.out_recoff = roff;
		} else b.out_recoff = nullptr;
		KG_CUDA(cudaEventRecord(db->ev[4], db->stream));
		KG_CUDA(cudaStreamSynchronize(db->stream));
		KG_CUDA(cudaGetLastError());
		if (stats) {
			stats->reads = n; stats->mapped = (int64_t)h[C_MAPPED]; stats->read_words = (int64_t)h[C_WORDS];
			stats->lookups = (int64_t)h[C_LOOKUPS]; stats->hits = (int64_t)h[C_HITS];
			stats->list_fetches = (int64_t)h[C_LISTS]; stats->list_ids = (int64_t)h[C_LISTIDS];
			stats->overflow_reads = 0;
			cudaEventElapsedTime(&stats->ms_seed, db->ev[2], db->ev[3]);
			cudaEventElapsedTime(&stats->ms_emit, db->ev[3], db->ev[4]);
			cudaEventElapsedTime(&stats->ms_total, db->ev[2], db->ev[4]);
			stats->launches = launches;
			stats->reserved = (int32_t)NR;
		}
		return 0;
	}
}
