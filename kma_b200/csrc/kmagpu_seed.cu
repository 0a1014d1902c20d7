// Stage-2 of KMA on the GPU: k-mer seeding + template candidate scoring (-1t1, single end).
//
// What is computed is save_kmers (savekmers.c:2442-3065) for every read of a batch of stage-1
// records; how it is computed is not the reference's:
//   * one warp per read, reads pulled from an atomic work counter by a persistent grid;
//   * phase 1 (gather): the 32 lanes look up 256 consecutive k-mer positions of a strand at once
//     (8 independent probes in flight per lane) against the HBM/L2-resident fused hash
//     {exist -> (key, value offset)}; the reverse strand is never materialised -- its k-mers are
//     the bit-reversed complements of the forward ones;
//   * phase 2 (score): the warp walks the hit positions found by ballot; hits on the same
//     template list collapse into a run score, a list change is handled by all lanes at once
//     (one lane per template of the list) against a per-warp shared-memory hash of the
//     templates seen by this read (insertion order kept with ballot/popc ranks);
//   * reads that see more distinct templates than the shared table holds are re-run by the same
//     code over a dense per-warp scratch in global memory (exactly the reference's Score[] /
//     extendScore[] / include[] arrays), so the result never depends on the table size;
//   * record sizes are prefix-summed on the device and a writer kernel emits the stage-2 byte
//     stream (ankers.c:30-50) in input order, so the host does no per-read work.
#include "kmagpu_internal.h"
#include "kmagpu_dev.cuh"
#include "kmagpu_seed.cuh"
#include <string.h>
#include <algorithm>

#define KG_CHUNK 256          // k-mer positions gathered per phase-1 round (8 per lane)
#define KG_PER_LANE (KG_CHUNK / 32)
#ifndef KG_CAP_LOG
#define KG_CAP_LOG 8
#endif
#define KG_CAP (1 << KG_CAP_LOG)          // shared hash slots per warp
#define KG_FILL (3 * KG_CAP / 4)          // distinct templates a read may see before it goes to the dense path
#ifndef KG_CLAIM
#define KG_CLAIM 4            // reads a warp claims from the work counter at a time
#endif
#define KG_WARPS 4            // warps per CTA
#ifndef KG_MINB
#define KG_MINB 8             // CTAs per SM the seeding grid is sized for (24.9 KB of shared memory each: at most 9). The kernel carries no
#endif                        // minimum-blocks bound: forcing 6 / 8 / 9 measured 41.8 / 38.8 / 41.9 ms against 37.7 ms without (profiles/r01_ab_seed.log)
#define KG_QCAP 48            // queued segments per warp: at most 16 stay behind a round, a round files at most 32
#define KG_WORDS 12           // staged u64 words per chunk: (256 + 31 + 31) / 32 + 2

struct SeedRes { int32_t score, ntmpl, flag; uint32_t pool_off; int32_t src, rev; };   // src: record that supplies read + name, rev: emit its reverse complement

// paired end: what get_kmers_for_pair (savekmers.c:427) leaves behind for one mate -- per strand the templates seen
// (first-seen order) with their clamped scores, and the larger of the two strands' hit counts
struct MateRes { uint32_t off_f, off_r; int32_t n_f, n_r, hits, scanned; };

struct SeedParams {
	int32_t M, MM, U, W1, exhaustive;
	int32_t use_proxi;   // -proxi (kma.c:702-718): getMatch = getProxiMatch (savekmers.c:296) instead of getBestMatch
	double proxi;        // |minFrac| as save_kmers_batch hands it on (kmers.c:133-141)
	unsigned long long *soft;   // soft proximity (-proxi < 0 with -mem_mode, kmers.c:133-153): softProxi[template] += score for every
	                            // template a get*Proxi* function keeps; NULL = off. Only the dense pass runs then (no second try per read)
};

// proxiScore = minFrac * bestScore, truncated into an int as the reference's assignment does
__device__ __forceinline__ int proxi_of(double f, int best) { return __double2int_rz(__dmul_rn(f, (double)best)); }

// score of a hit that resumes template bookkeeping after `gaps` missed k-mer positions.
// run == true : contribution to the run score of an unchanged template list (savekmers.c:2529-2569)
// run == false: direct per-template score after a list change            (savekmers.c:2592-2625)
// more than k missed positions (an indel or several mismatches): rare, and its division is long -- kept out of the scan loops
__device__ __noinline__ int gap_score_far(const SeedParams &p, int k, int gaps) {
	int g = gaps - (k - 1), mm, m;
	if (g <= 2) { mm = g; m = 0; }
	else {
		mm = g / k + (g % k ? 1 : 0); mm = max(mm, 2);
		m = min(min(g - mm, k), mm);
	}
	const int a = p.W1 + (g - 1) * p.U, b = mm * p.MM + m * p.M;
	return k * p.M + (a <= b ? b : a);
}

__device__ __forceinline__ int gap_score(const SeedParams &p, int k, int gaps, bool run) {
	if (gaps == 0) return p.M;
	if (gaps == k) return k * p.M + p.MM;
	if (k < gaps) return gap_score_far(p, k, gaps);
	return gaps * p.M + (k - gaps) * p.U + p.W1;
}

// ---------------------------------------------------------------- template bookkeeping

template <bool DENSE>
struct Store {
	// hash mode: shared memory; dense mode: per-warp global scratch indexed by template id
	int *keys, *score, *ext;   // hash: [KG_CAP]; dense: keys unused, score/ext [DB_size + 1]
	uint8_t *incl;             // dense only
	int *cand;                 // first-seen order: hash -> slot, dense -> template id
	int ncand;

	__device__ __forceinline__ int find(int t) const {
		if (DENSE) return t;
		unsigned h = ((unsigned)t * 0x9E3779B1u) >> (32 - KG_CAP_LOG);
		while (keys[h] != t) h = (h + 1) & (KG_CAP - 1);
		return (int)h;
	}
	__device__ __forceinline__ int find_or_insert(int t, bool *isnew) {
		if (DENSE) { *isnew = !incl[t]; incl[t] = 1; return t; }
		unsigned h = ((unsigned)t * 0x9E3779B1u) >> (32 - KG_CAP_LOG);
		for (;;) {
			int cur = keys[h];
			if (cur == t) { *isnew = false; return (int)h; }
			if (cur == 0) {
				int old = atomicCAS(&keys[h], 0, t);
				if (old == 0) { *isnew = true; return (int)h; }
				if (old == t) { *isnew = false; return (int)h; }
			}
			h = (h + 1) & (KG_CAP - 1);
		}
	}
	__device__ __forceinline__ int tmpl_of(int slot) const { return DENSE ? slot : keys[slot]; }
};

struct WarpStats { unsigned lookups, hits, lists, listids; };

// One strand of one read. Returns best score (>= 0), leaves the arg-max template ids in
// st.cand[0 .. *nbest) (first-seen order) and the store clean. Returns -1 on table overflow
// (hash mode only; store is left clean).
#ifndef KG_SCAN_INL
#define KG_SCAN_INL                // inlined per call site: measured 40 ms vs 61 ms out of line (by-reference arguments go through
#endif                             // the stack), although the four copies make 200 KB of code
// GENERIC = false: the kernel for the common shape -- hashed table (not direct addressed), 16-bit template lists, read
// without N's -- compiled without the other branches, which keeps its loops short in the instruction cache.
template <bool DENSE, bool GENERIC>
__device__ KG_SCAN_INL int scan_strand(const KgHashView &hv, const SeedParams &p, const ReadCtx &rc, int strand,
                           Store<DENSE> &st, uint32_t *hits, uint64_t *sw, int *sq, int *nbest, WarpStats &ws,
                           int2 *pool2 = nullptr, unsigned long long pool2_cap = 0, unsigned long long *ctr = nullptr,
                           uint32_t *list_off = nullptr, bool pre = false) {
	// pre: the kernel staged the whole read (one chunk) and ran the quick check for both strands already
	const unsigned lane = threadIdx.x & 31;
	const int k = hv.kmersize;
	const int L = rc.seqlen;
	const int npos = L - k + 1;   // k-mer positions on the strand
	st.ncand = 0;
	*nbest = 0;

	// stage the forward words a strand chunk [c0, c0 + KG_CHUNK) needs; returns first staged word
	auto stage = [&](int c0) -> int {
		int flo, fhi;   // forward base range touched
		if (strand == 0) { flo = c0; fhi = min(L, c0 + KG_CHUNK + k - 1) - 1; }
		else { fhi = L - 1 - c0; flo = max(0, L - k - (c0 + KG_CHUNK - 1)); }
		int w0 = flo >> 5, w1 = fhi >> 5;
		__syncwarp();
#pragma unroll 1
		for (int w = w0 + (int)lane; w <= w1 + 1; w += 32)
			sw[w - w0] = w < rc.words ? ld_u64u(rc.seq + 8 * (size_t)w) : 0ull;
		__syncwarp();
		return w0;
	};
	auto kmer_of = [&](int w0, int j) -> uint64_t {   // k-mer at strand position j
		if (strand == 0) return kmer_from(sw, w0, j, k);
		uint64_t f = kmer_from(sw, w0, L - k - j, k);
		return rev2(~f) >> (64 - 2 * k);
	};

	// ---- quick check (savekmers.c:2485-2495): every k-th k-mer of each N-free stretch
	bool any = p.exhaustive != 0 || pre;
	if (!any) {
		for (int c0 = 0; c0 < npos && !any; c0 += KG_CHUNK) {
			int w0 = stage(c0);
			bool h = false;
			if (!GENERIC || rc.nN == 0) {
				// probes c0' = multiples of k inside the chunk
				int first = ((c0 + k - 1) / k) * k;
				const int lim = min(npos, c0 + KG_CHUNK);
				for (int jb = first; jb < lim && !h; jb += 32 * k) {
					const int j = jb + (int)lane * k;
					const bool act = j < lim;
					const bool hp = act && hash_lookup(hv, kmer_of(w0, j)) != KG_MISS;
					const unsigned am = __ballot_sync(0xffffffffu, act), hm = __ballot_sync(0xffffffffu, hp);
					// algorithmic probe count: the reference stops at the first hit
					if (lane == 0) ws.lookups += hm ? __ffs(hm) : __popc(am);
					h = hm != 0;
				}
			} else {
#pragma unroll 1
				for (int j = c0 + (int)lane; j < min(npos, c0 + KG_CHUNK); j += 32) {
					int ss;
					if (pos_valid(rc, j, k, strand, &ss) && (j - ss) % k == 0) {
						h |= hash_lookup(hv, kmer_of(w0, j)) != KG_MISS;
						ws.lookups++;
					}
				}
			}
			any = __any_sync(0xffffffffu, h);
		}
		if (!any) return 0;
	}

	// ---- exhaustive scan (savekmers.c:2511-2706)
	// The reference walks the hits in order: a hit on the list of the hit before it adds a gap-class score to the run of
	// that list (gap 0: +M), a hit on another list flushes the run into the templates of the old list and scores the
	// templates of the new one by their own gap since they were last seen. Restated per SEGMENT = maximal series of hits
	// on one list (gaps included): every hit's contribution to its segment's run score depends only on the hit before
	// it, so the lanes compute them in parallel and a prefix sum folds them; a segment then touches every template of
	// its list ONCE (first-seen score or gap score, + the whole run score, + the position it was last seen at) -- the
	// sums and the last-seen positions are the reference's, without its flush pass.
	const unsigned full = 0xffffffffu, lt = (1u << lane) - 1u;
	int nhits = 0;
	bool overflow = false;
	int prev_pos = -1;                 // position / list of the last hit so far
	uint32_t prev_off = KG_MISS;
	// Closed segments wait in a per-warp queue {list, first and last hit position, run score}; the last entry is the
	// open segment, which later hits on its list still extend. The queue is applied in PACKED rounds: the (segment,
	// template) items of consecutive segments fill the 32 lanes (a list of this kind of database names ~7 templates, so
	// four segments share a round). Items of one round that name the same template are found with match.any: the
	// lowest of them does the table work (find or insert, first-seen score or gap score from the table's last-seen
	// position), the others resume from the segment before them in the round -- whose last position is what the table
	// would hold by then -- and add their part atomically; the highest one leaves its segment's last position behind.
	uint32_t *q_off = (uint32_t *)sq;
	int *q_first = sq + KG_QCAP, *q_last = sq + 2 * KG_QCAP, *q_run = sq + 3 * KG_QCAP;
	int nseg = 0;

	auto drain = [&](int cnt) -> bool {   // applies queue entries [0, cnt), cnt <= 32
		uint32_t off = 0;
		int nl = 0;
		if ((int)lane < cnt) { off = q_off[lane]; nl = GENERIC ? list_len(hv, off) : (int)__ldg(hv.values_s + off); }
		int incl = nl;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(full, incl, o); if ((int)lane >= o) incl += y; }
		const int excl = incl - nl, total = __shfl_sync(full, incl, 31);
		ws.lists += cnt; ws.listids += total;
#pragma unroll 1
		for (int g0 = 0; g0 < total; g0 += 32) {
			if (!DENSE && st.ncand + min(32, total - g0) > KG_FILL) return false;
			const bool act = g0 + (int)lane < total;
			const int g = act ? g0 + (int)lane : total - 1;
			int sg = 0;   // the segment of item g: how many segments end at or before it
#pragma unroll
			for (int step = 16; step; step >>= 1) { const int v = __shfl_sync(full, incl, sg + step - 1); if (v <= g) sg += step; }
			const uint32_t soff = __shfl_sync(full, off, sg);
			const int idx = g - __shfl_sync(full, excl, sg);
			const int first = q_first[sg], last = q_last[sg], run = q_run[sg];
			const int T = GENERIC ? list_id(hv, soff, idx) : (int)__ldg(hv.values_s + soff + 1 + idx);
			const unsigned mm = __match_any_sync(full, act ? T : -1 - (int)lane);
			const unsigned below = mm & lt;
			const bool lead = act && !below;
			bool isnew = false;
			int sl = 0;
			if (lead) sl = st.find_or_insert(T, &isnew);
			sl = __shfl_sync(full, sl, __ffs(mm) - 1);
			const int plast = __shfl_sync(full, last, below ? 31 - __clz(below) : 0);
			int add = run;
			if (lead && isnew) add += k * p.M;                                                       // savekmers.c:2682-2688
			else add += gap_score(p, k, (first - 1) - (lead ? st.ext[sl] : plast), false);            // savekmers.c:2583-2655, 2575-2582
			if (lead && isnew) st.score[sl] = add;
			__syncwarp();
			if (act && !(lead && isnew)) atomicAdd(&st.score[sl], add);
			if (act && (mm >> lane) == 1u) st.ext[sl] = last;
			const unsigned nm = __ballot_sync(full, isnew);
			if (isnew) st.cand[st.ncand + __popc(nm & lt)] = sl;
			st.ncand += __popc(nm);
			__syncwarp();
		}
		return true;
	};

	for (int c0 = 0; c0 < npos && !overflow; c0 += KG_CHUNK) {
		const int w0 = pre ? 0 : stage(c0);
		const int lim = min(npos, c0 + KG_CHUNK), nround = (lim - c0 + 31) >> 5;
		// phase 1: gather, four rounds of 32 positions at a time so that 4 independent probes per lane are in flight
		if (GENERIC && rc.nN) {   // reads with N's (rare): validity per position, one probe at a time; kept off the hot path
#pragma unroll 1
			for (int u = 0; u < nround; ++u) {
				const int j = c0 + u * 32 + (int)lane;
				int ss;
				uint32_t v = KG_MISS;
				if (j < npos && pos_valid(rc, j, k, strand, &ss)) { v = hash_lookup(hv, kmer_of(w0, j)); ws.lookups++; }
				hits[u * 32 + lane] = v;
			}
		} else {
#pragma unroll 1
			for (int ug = 0; ug < nround; ug += 4) {
				uint32_t e1[4];
				if (GENERIC && hv.mega) {
#pragma unroll
					for (int u = 0; u < 4; ++u) {
						const int j = c0 + (ug + u) * 32 + (int)lane;
						e1[u] = KG_MISS;
						if (j < lim) { const uint32_t v = __ldg(hv.exist + kmer_of(w0, j)); e1[u] = v != 1u ? v : KG_MISS; ws.lookups++; }
					}
				} else {
					// four independent 16-byte bucket loads in flight per lane; a bucket entry answers the probe by itself unless its
					// first key differs and the bucket holds more (then the run in kv is read: one more dependent sector)
					uint32_t key[4];
					uint4 b4[4];
#pragma unroll
					for (int u = 0; u < 4; ++u) {
						const int j = c0 + (ug + u) * 32 + (int)lane;
						const bool ok = j < lim;
						const uint64_t km = ok ? kmer_of(w0, j) : 0ull;
						key[u] = (uint32_t)km;
						b4[u] = make_uint4(0, 0, 0, 0);
						if (ok) { b4[u] = __ldg(hv.bk + (uint32_t)(km & hv.hmask)); ws.lookups++; }
					}
#pragma unroll
					for (int u = 0; u < 4; ++u) e1[u] = hash_resolve(hv, b4[u], key[u]);
				}
#pragma unroll
				for (int u = 0; u < 4; ++u) hits[(ug + u) * 32 + lane] = e1[u];
			}
		}
		__syncwarp();

		// phase 2: the segments of this chunk, 32 positions per round, no loop over the segments: every lane that starts
		// one files it
#pragma unroll 1
		for (int u = 0; u < nround && !overflow; ++u) {
			const uint32_t myoff = hits[u * 32 + lane];
			const bool hit = myoff != KG_MISS;
			const unsigned hm = __ballot_sync(full, hit);
			if (hm) {
				nhits += __popc(hm);
				const int base = c0 + u * 32;
				// the hit before this position: in this round, or the one carried over
				const unsigned below = hm & lt;
				const int pl = below ? 31 - __clz(below) : -1;
				uint32_t poff = __shfl_sync(full, myoff, pl < 0 ? 0 : pl);
				int ppos = base + pl;
				if (pl < 0) { poff = prev_off; ppos = prev_pos; }
				const bool same = hit && poff == myoff;   // continues (gap 0) or resumes the list of the hit before it
				int ps = same ? gap_score(p, k, base + (int)lane - ppos - 1, true) : 0;   // what this hit adds to its segment's run score
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(full, ps, o); if ((int)lane >= o) ps += y; }
				const unsigned sm = __ballot_sync(full, hit && !same);   // list changes: a new segment starts
				const int fs = sm ? __ffs(sm) - 1 : 32;
				const unsigned head = hm & (fs == 32 ? full : ((1u << fs) - 1u));   // hits that still belong to the open segment
				if (head) {
					const int hl = 31 - __clz(head);
					const int addrun = __shfl_sync(full, ps, hl);
					if (lane == 0) { q_run[nseg - 1] += addrun; q_last[nseg - 1] = base + hl; }
				}
				if (sm) {
					const bool start = (sm >> lane) & 1u;
					const unsigned above = sm & ~lt & ~(1u << lane);   // the starts after this lane
					const unsigned upto = above ? ((1u << (__ffs(above) - 1)) - 1u) : full;
					const unsigned body = hm & upto & ~lt;              // this segment's hits in the round
					const int ll = start ? 31 - __clz(body) : (int)lane;
					const int runv = __shfl_sync(full, ps, ll) - ps;
					const int qi = nseg + __popc(sm & lt);
					if (start) { q_off[qi] = myoff; q_first[qi] = base + (int)lane; q_last[qi] = base + ll; q_run[qi] = runv; }
					nseg += __popc(sm);
				}
				const int lh = 31 - __clz(hm);
				prev_pos = base + lh;
				prev_off = __shfl_sync(full, myoff, lh);
				__syncwarp();
			}
			// make room for the next round (all but the open segment, 32 at most per pass); after the strand's last
			// round everything goes (savekmers.c:2707-2722)
			const bool fin = u == nround - 1 && c0 + KG_CHUNK >= npos;
			while (nseg > (fin ? 0 : KG_QCAP - 32)) {
				const int cnt = min(fin ? nseg : nseg - 1, 32);
				if (!drain(cnt)) { overflow = true; break; }
				uint32_t mo = 0; int mf = 0, ml = 0, mr = 0;
				const bool mv = (int)lane < nseg - cnt;
				if (mv) { mo = q_off[cnt + lane]; mf = q_first[cnt + lane]; ml = q_last[cnt + lane]; mr = q_run[cnt + lane]; }
				__syncwarp();
				if (mv) { q_off[lane] = mo; q_first[lane] = mf; q_last[lane] = ml; q_run[lane] = mr; }
				nseg -= cnt;
				__syncwarp();
			}
		}
		__syncwarp();
	}
	ws.hits += lane == 0 ? nhits : 0;

	if (overflow) {   // hash mode only: wipe and report
#pragma unroll 1
		for (int i = lane; i < KG_CAP; i += 32) st.keys[i] = 0;
		__syncwarp();
		return -1;
	}

	if (pool2) {   // paired end: every template seen keeps its clamped score (savekmers.c:654-686); returns the hit count
		unsigned long long po = 0;
		if (lane == 0) po = atomicAdd(&ctr[C_POOL2], (unsigned long long)st.ncand);
		po = __shfl_sync(0xffffffffu, po, 0);
		const bool fits = po + st.ncand <= pool2_cap;
		if (!fits && lane == 0) atomicAdd(&ctr[C_POOLFAIL], 1ull);
#pragma unroll 1
		for (int i = lane; i < st.ncand; i += 32) {
			const int s = st.cand[i];
			if (fits) pool2[po + i] = make_int2(st.tmpl_of(s), max(st.score[s], 0));
			if (DENSE) { st.score[s] = 0; st.ext[s] = 0; st.incl[s] = 0; }
		}
		__syncwarp();
		if (!DENSE) {
#pragma unroll 1
			for (int i = lane; i < KG_CAP; i += 32) st.keys[i] = 0;
			__syncwarp();
		}
		*nbest = st.ncand;
		*list_off = (uint32_t)po;
		return nhits;
	}
	// arg-max set in first-seen order (getBestMatch, savekmers.c:273-294), negatives clamp to 0; under -proxi every
	// template whose raw score reaches minFrac * best (getProxiMatch, savekmers.c:296-340)
	int best = 0;
#pragma unroll 1
	for (int i = lane; i < st.ncand; i += 32) best = max(best, st.score[st.cand[i]]);
#pragma unroll
	for (int o = 16; o; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
	const int thr = p.use_proxi ? proxi_of(p.proxi, best) : 0;
	int nb = 0;
#pragma unroll 1
	for (int base = 0; base < st.ncand; base += 32) {
		int i = base + (int)lane;
		bool ok = false;
		int t = 0;
		if (i < st.ncand) {
			int s = st.cand[i];
			ok = p.use_proxi ? thr <= st.score[s] : max(st.score[s], 0) == best;
			t = st.tmpl_of(s);
			if (ok && p.soft) atomicAdd(p.soft + t, (unsigned long long)st.score[s]);   // getProxiMatch (savekmers.c:330-332)
			if (DENSE) { st.score[s] = 0; st.ext[s] = 0; st.incl[s] = 0; }
		}
		unsigned m = __ballot_sync(0xffffffffu, ok);
		__syncwarp();
		if (ok) st.cand[nb + __popc(m & ((1u << lane) - 1))] = t;
		nb += __popc(m);
		__syncwarp();
	}
	if (!DENSE) {
#pragma unroll 1
		for (int i = lane; i < KG_CAP; i += 32) st.keys[i] = 0;
		__syncwarp();
	}
	*nbest = nhits ? nb : 0;
	return nhits ? best : 0;
}

// ---------------------------------------------------------------- the seeding kernel

#ifndef KG_LB
#define KG_LB 8               // resident CTAs per SM the kernels are compiled for (register cap 65536 / (128 * KG_LB) = 64). Measured on C2
#endif                        // (profiles/r02_seed_occupancy.log): 8 -> 28.9 ms, 10 (48 registers, spills) 29.3, 12 (40) 31.1, no cap (6 resident) 35.4
template <bool DENSE, bool GENERIC>
__global__ void __launch_bounds__(KG_WARPS * 32, KG_LB)
seed_se_kernel(KgHashView hv, SeedParams p, const uint8_t *__restrict__ in, const uint32_t *__restrict__ rec_off,
               int nreads, SeedRes *__restrict__ res, uint32_t *__restrict__ recsize, int32_t *__restrict__ pool,
               unsigned long long pool_cap, unsigned long long *ctr, uint32_t *__restrict__ ovf_list,
               uint8_t *dense_scratch, size_t dense_stride, const uint8_t *__restrict__ kinds, MateRes *__restrict__ mates,
               int2 *pool2, unsigned long long pool2_cap, uint32_t *__restrict__ nlist, int from_nlist) {
	__shared__ uint32_t s_hits[KG_WARPS][KG_CHUNK];
	__shared__ uint64_t s_words[KG_WARPS][KG_WORDS];
	__shared__ int s_queue[KG_WARPS][4 * KG_QCAP];
	__shared__ int s_tab[DENSE ? 1 : KG_WARPS][DENSE ? 1 : 3 * KG_CAP];
	__shared__ int s_cand[DENSE ? 1 : KG_WARPS][DENSE ? 1 : 2 * KG_CAP];

	const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	uint32_t *hits = s_hits[wid];
	uint64_t *sw = s_words[wid];
	int *sq = s_queue[wid];
	Store<DENSE> st;
	int *candF, *candR;
	if (DENSE) {
		uint8_t *base = dense_scratch + dense_stride * ((size_t)blockIdx.x * KG_WARPS + wid);
		const size_t D = (size_t)hv.DB_size + 1;
		st.score = (int *)base; st.ext = st.score + D; candF = st.ext + D; candR = candF + D;
		st.incl = (uint8_t *)(candR + D); st.keys = nullptr;
	} else {
		st.keys = s_tab[wid]; st.score = st.keys + KG_CAP; st.ext = st.score + KG_CAP; st.incl = nullptr;
		candF = s_cand[wid]; candR = candF + KG_CAP;
#pragma unroll 1
		for (int i = lane; i < KG_CAP; i += 32) st.keys[i] = 0;
		__syncwarp();
	}
	WarpStats ws = {0, 0, 0, 0};
	unsigned mapped = 0, words_seen = 0;
	const int k = hv.kmersize;
	// work: every read (the first pass), the reads with N's the first pass set aside (GENERIC, from_nlist), or the reads
	// whose template table overflowed (DENSE)
	const int total = DENSE ? (int)ctr[C_OVF] : (from_nlist ? (int)ctr[C_NLIST] : nreads);

	unsigned long long wnext = 0;
	int wleft = 0;   // reads are claimed KG_CLAIM at a time: one contended atomic per four reads. A batch of few (long) reads is
	                 // claimed read by read: there the balance over the warps is what counts
	const int claim = (long long)total >= 64ll * gridDim.x * KG_WARPS ? KG_CLAIM : 1;
	for (;;) {
		if (!wleft) {
			if (lane == 0) wnext = atomicAdd(&ctr[DENSE ? C_WORK2 : (from_nlist ? C_WORK3 : C_WORK)], (unsigned long long)claim);
			wnext = __shfl_sync(0xffffffffu, wnext, 0);
			wleft = claim;
		}
		const unsigned long long w = wnext++;
		--wleft;
		if (w >= (unsigned long long)total) break;
		const int r = DENSE ? (int)ovf_list[w] : (from_nlist ? (int)nlist[w] : (int)w);

		ReadCtx rc;
		rc.rec = in + rec_off[r];
		rc.seqlen = (int)ld_u32u(rc.rec);
		rc.words = (int)ld_u32u(rc.rec + 4);
		rc.nN = (int)ld_u32u(rc.rec + 8);
		rc.hdrlen = abs((int)ld_u32u(rc.rec + 12));
		rc.seq = rc.rec + 16;
		rc.N = rc.seq + 8 * (size_t)rc.words;
		if (!GENERIC && rc.nN) {   // not this kernel's shape: the generic pass takes it
			if (lane == 0) nlist[atomicAdd(&ctr[C_NLIST], 1ull)] = (uint32_t)r;
			continue;
		}
		words_seen += rc.words;

		// both strands through ONE call site (one inlined copy of the scan: four copies made 200 KB of code). A mate of
		// a pair keeps every template's score for pair_select_kernel, a single read only its arg-max sets.
		const bool mate = kinds && kinds[r];
		int sres0 = 0, sres1 = 0, scnt0 = 0, scnt1 = 0;   // per strand: score / hit count, templates kept, their pool offset
		uint32_t soff0 = 0, soff1 = 0;
		bool ovf = false;
		if (rc.seqlen >= k) {
			// a read of one chunk (every short read): its words are staged once for both strands, and the quick check
			// (savekmers.c:2485-2495: every k-th k-mer until one is known) probes both strands at once, 16 lanes each
			const int npos = rc.seqlen - k + 1;
			const bool pre = npos <= KG_CHUNK && !(GENERIC && rc.nN);
			unsigned want = 3u;
			if (pre) {
				__syncwarp();
				if ((int)lane <= ((rc.seqlen - 1) >> 5) + 1) sw[lane] = (int)lane < rc.words ? ld_u64u(rc.seq + 8 * (size_t)lane) : 0ull;
				__syncwarp();
				if (!p.exhaustive) {
					want = 0u;
					unsigned open = 3u;   // strands still probing
					const int sd = (int)(lane >> 4);
#pragma unroll 1
					for (int jb = 0; jb < npos && open; jb += 16 * k) {
						const int j = jb + (int)(lane & 15u) * k;
						const bool act = j < npos && ((open >> sd) & 1u);
						bool hp = false;
						if (act) {
							uint64_t km = kmer_from(sw, 0, sd ? rc.seqlen - k - j : j, k);
							if (sd) km = rev2(~km) >> (64 - 2 * k);
							hp = hash_lookup(hv, km) != KG_MISS;
						}
						const unsigned am = __ballot_sync(0xffffffffu, act), hm = __ballot_sync(0xffffffffu, hp);
#pragma unroll
						for (int sx = 0; sx < 2; ++sx) {
							const unsigned a = (am >> (16 * sx)) & 0xffffu, h = (hm >> (16 * sx)) & 0xffffu;
							if (!((open >> sx) & 1u)) continue;
							if (lane == 0) ws.lookups += h ? __ffs(h) : __popc(a);   // algorithmic probe count: the reference stops at the first hit
							if (h) { want |= 1u << sx; open &= ~(1u << sx); }
						}
					}
				}
			}
#pragma unroll 1
			for (int strand = 0; strand < 2 && !ovf; ++strand) {
				if (!((want >> strand) & 1u)) continue;
				st.cand = strand ? candR : candF;
				int cnt = 0;
				uint32_t off = 0;
				const int sr = scan_strand<DENSE, GENERIC>(hv, p, rc, strand, st, hits, sw, sq, &cnt, ws, mate ? pool2 : nullptr, pool2_cap, ctr, &off, pre);
				if (strand) { sres1 = sr; scnt1 = cnt; soff1 = off; } else { sres0 = sr; scnt0 = cnt; soff0 = off; }
				ovf = sr < 0;
			}
			if (ovf && lane == 0) ovf_list[atomicAdd(&ctr[C_OVF], 1ull)] = (uint32_t)r;   // table overflow: dense pass
		}
		if (mate) {
			MateRes m = {0, 0, 0, 0, 0, 0};
			if (rc.seqlen >= k && !ovf) {
				m.off_f = soff0; m.off_r = soff1; m.n_f = scnt0; m.n_r = scnt1; m.hits = max(sres0, sres1); m.scanned = 1;
			}
			if (lane == 0 && !ovf) mates[r] = m;
			__syncwarp();
			continue;
		}
		SeedRes out = {0, 0, 0, 0, r, 0};
		uint32_t size = 0;
		if (rc.seqlen >= k) {
			const int nf = scnt0, nr = scnt1, bf = sres0, br = sres1;
			if (ovf) out.flag = -1;
			else if ((bf > 0 || br > 0) && (k <= bf || k <= br)) {   // savekmers.c:3039-3061
				int nt = bf > br ? nf : (bf < br ? nr : nf + nr);
				unsigned long long po = 0;
				if (lane == 0) po = atomicAdd(&ctr[C_POOL], (unsigned long long)nt);
				po = __shfl_sync(0xffffffffu, po, 0);
				if (po + nt > pool_cap) {
					if (lane == 0) atomicAdd(&ctr[C_POOLFAIL], 1ull);
				} else {
					int32_t *dst = pool + po;
#pragma unroll 1
					for (int i = lane; i < nt; i += 32) {   // forward set, reverse set, or both with the reverse ids negated
						int v;
						if (bf > br) v = candF[i];
						else if (bf < br) v = candR[i];
						else v = i < nf ? candF[i] : -candR[i - nf];
						dst[i] = v;
					}
				}
				out.score = bf > br ? bf : (bf < br ? br : -bf);
				out.ntmpl = nt;
				out.flag = bf < br ? 16 : 0;
				out.rev = bf < br ? 1 : 0;
				out.pool_off = (uint32_t)po;
				size = 28u + 8u * rc.words + 4u * rc.nN + 4u * nt + rc.hdrlen;
				++mapped;
			}
		}
		if (lane == 0) { res[r] = out; recsize[r] = size; }
		__syncwarp();
	}
	// per-warp statistics -> global
	for (int o = 16; o; o >>= 1) ws.lookups += __shfl_xor_sync(0xffffffffu, ws.lookups, o);
	if (lane == 0) {
		atomicAdd(&ctr[C_LOOKUPS], (unsigned long long)ws.lookups);
		atomicAdd(&ctr[C_HITS], (unsigned long long)ws.hits);
		atomicAdd(&ctr[C_LISTS], (unsigned long long)ws.lists);
		atomicAdd(&ctr[C_LISTIDS], (unsigned long long)ws.listids);
		atomicAdd(&ctr[C_MAPPED], (unsigned long long)mapped);
		if (!DENSE) atomicAdd(&ctr[C_WORDS], (unsigned long long)words_seen);
	}
}


// ---------------------------------------------------------------- paired-end selection (-apm p)

__device__ __forceinline__ int list_score(const int2 *L, int n, int t) {   // Score[t] of a strand (0 when not seen)
	for (int i = 0; i < n; ++i) if (L[i].x == t) return L[i].y;
	return 0;
}

// save_kmers_penaltyPair (savekmers.c:3572-3777) over the per-strand score lists of the two mates, one thread per
// pair: getFirstPen (:1383), getSecondBestPen (:1415) / getF_Best (:1648), the proper-pair test and flag logic, and
// printPair's record order (ankers.c:150). Slot r / r+1 of res + recsize describe the first / second record emitted.
__global__ void pair_select_kernel(const uint8_t *__restrict__ in, const uint32_t *__restrict__ rec_off, int nrec,
		const uint8_t *__restrict__ kinds, const MateRes *__restrict__ mates, const int2 *__restrict__ pool2, int32_t *pool,
		unsigned long long pool_cap, unsigned long long *ctr, SeedRes *res, uint32_t *recsize, int k, int PE, int apm,
		int use_proxi, double pf, unsigned long long *soft) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= nrec || kinds[r] != 1) return;
	// a strand list that did not fit the pool was never written and its offset points past the allocation: the host
	// grows the pools and redoes the batch, nothing of this attempt is used
	if (ctr[C_POOLFAIL]) { recsize[r] = recsize[r + 1] = 0; return; }
	const MateRes A = mates[r], B = mates[r + 1];
	const uint8_t *recA = in + rec_off[r], *recB = in + rec_off[r + 1];
	const int lenA = (int)ld_u32u(recA), lenB = (int)ld_u32u(recB);
	const int2 *F1 = pool2 + A.off_f, *R1 = pool2 + A.off_r, *F2 = pool2 + B.off_f, *R2 = pool2 + B.off_r;
	const int n1 = A.n_f + A.n_r, n2 = B.n_f + B.n_r;
	SeedRes o1 = {0, 0, 0, 0, r, 0}, o2 = {0, 0, 0, 0, r + 1, 0};
	uint32_t sz1 = 0, sz2 = 0;
	unsigned long long po = atomicAdd(&ctr[C_POOL], (unsigned long long)(n1 + n2));
	if (po + n1 + n2 > pool_cap) { atomicAdd(&ctr[C_POOLFAIL], 1ull); res[r] = o1; res[r + 1] = o2; recsize[r] = recsize[r + 1] = 0; return; }
	int32_t *rt = pool + po, *bt = rt + n1;
	int nrt = 0, nbt = 0, best = 0, best_r = 0;
	const int hc = A.hits, hc_r = B.hits;
	bool proper = false;
	if (apm == 1) {
		// save_kmers_unionPair (savekmers.c:3367-3570, the default pairing) with getF_Best (:1648) / getR_Best (:1682):
		// each mate keeps its arg-max set; the pair is proper when a template of the first mate's set is also among
		// the second mate's best on the opposite strand
		// getF_Best / getR_Best keep the arg-max set; getF_Proxi (:1764) / getR_Proxi (:1825) every template within minFrac of it
		int thr2 = 0;   // the second mate's proximity score (the union test needs it)
		auto argmax = [&](const int2 *F, int nf, const int2 *R, int nr, int32_t *dst, int *cnt) {
			int b = 0, h = 0;
			if (use_proxi) {
				for (int i = 0; i < nf + nr; ++i) b = max(b, i < nf ? F[i].y : R[i - nf].y);
				const int thr = proxi_of(pf, b);
				for (int i = 0; i < nf + nr; ++i) {
					const int t = i < nf ? F[i].x : -R[i - nf].x, sc = i < nf ? F[i].y : R[i - nf].y;
					if (thr <= sc) { dst[h++] = t; if (soft) atomicAdd(soft + abs(t), (unsigned long long)sc); }   // getF_Proxi / getR_Proxi
				}
				thr2 = thr;
				*cnt = h;
				return b;
			}
			for (int i = 0; i < nf + nr; ++i) {
				const int t = i < nf ? F[i].x : -R[i - nf].x, sc = i < nf ? F[i].y : R[i - nf].y;
				if (b < sc) { b = sc; h = 1; dst[0] = t; }
				else if (b == sc) dst[h++] = t;
			}
			*cnt = h;
			return b;
		};
		if (hc) {
			best = argmax(F1, A.n_f, R1, A.n_r, rt, &nrt);
			if (k < best && best * k < (lenA - best)) best = 0;
		}
		if (hc_r) {
			if (best) {
				best_r = argmax(F2, B.n_f, R2, B.n_r, bt, &nbt);
				int hits = 0;
				if (use_proxi) {   // getR_Proxi's union: templates of the first set the second mate keeps (with a score) on the opposite strand
					for (int i = 0; i < nrt; ++i) {
						const int t = rt[i];
						const int sc = 0 < t ? list_score(R2, B.n_r, t) : list_score(F2, B.n_f, -t);
						if (sc && thr2 <= sc) {
							const int tmp = rt[hits]; rt[hits] = rt[i]; rt[i] = tmp;
							++hits;
						}
					}
				} else if (best_r) {
					for (int i = 0; i < nrt; ++i) {
						const int t = rt[i];
						if ((0 < t ? list_score(R2, B.n_r, t) : list_score(F2, B.n_f, -t)) == best_r) {
							const int tmp = rt[hits]; rt[hits] = rt[i]; rt[i] = tmp;
							++hits;
						}
					}
				}
				if (hits) { proper = true; nrt = hits; }
			} else best_r = argmax(F2, B.n_f, R2, B.n_r, rt, &nrt);
			if (k < best_r && best_r * k < (lenB - best_r)) { best_r = 0; proper = false; }
		}
		int curA = A.scanned, curB = B.scanned;
		int flag = 65, flag_r = 129;
		const uint32_t baseA = 28u + 8u * ld_u32u(recA + 4) + 4u * ld_u32u(recA + 8) + (uint32_t)abs((int)ld_u32u(recA + 12));
		const uint32_t baseB = 28u + 8u * ld_u32u(recB + 4) + 4u * ld_u32u(recB + 8) + (uint32_t)abs((int)ld_u32u(recB + 12));
		const uint32_t rt_off = (uint32_t)po, bt_off = (uint32_t)po + (uint32_t)n1;
		if (0 < best && 0 < best_r) {
			if (proper) {
				flag |= 2; flag_r |= 2;
				if (0 < rt[0]) {
					flag |= 32; flag_r |= 16; curA ^= 1;
					o1.score = best; o1.ntmpl = 0; o1.flag = flag; o1.src = r; o1.rev = curA; o1.pool_off = rt_off; sz1 = baseA;
					o2.score = best_r; o2.ntmpl = nrt; o2.flag = flag_r; o2.src = r + 1; o2.rev = curB; o2.pool_off = rt_off; sz2 = baseB + 4u * nrt;
				} else {
					flag |= 16; flag_r |= 32; curB ^= 1;
					for (int i = 0; i < nrt; ++i) rt[i] = -rt[i];
					o1.score = best_r; o1.ntmpl = 0; o1.flag = flag_r; o1.src = r + 1; o1.rev = curB; o1.pool_off = rt_off; sz1 = baseB;
					o2.score = best; o2.ntmpl = nrt; o2.flag = flag; o2.src = r; o2.rev = curA; o2.pool_off = rt_off; sz2 = baseA + 4u * nrt;
				}
			} else {
				int sA = best, sB = best_r;
				if (0 < rt[0]) { curA ^= 1; if (rt[nrt - 1] < 0) sA = -sA; }
				else { flag |= 16; flag_r |= 32; for (int i = 0; i < nrt; ++i) rt[i] = -rt[i]; }
				if (0 < bt[0]) { curB ^= 1; if (bt[nbt - 1] < 0) sB = -sB; }
				else { flag |= 32; flag_r |= 16; for (int i = 0; i < nbt; ++i) bt[i] = -bt[i]; }
				o1.score = sA; o1.ntmpl = nrt; o1.flag = flag; o1.src = r; o1.rev = curA; o1.pool_off = rt_off; sz1 = baseA + 4u * nrt;
				o2.score = sB; o2.ntmpl = nbt; o2.flag = flag_r; o2.src = r + 1; o2.rev = curB; o2.pool_off = bt_off; sz2 = baseB + 4u * nbt;
			}
		} else if (0 < best) {
			int sA = best;
			flag |= 8 | 32;
			if (0 < rt[0]) { curA ^= 1; if (rt[nrt - 1] < 0) sA = -sA; }
			else { flag |= 16; for (int i = 0; i < nrt; ++i) rt[i] = -rt[i]; }
			o1.score = sA; o1.ntmpl = nrt; o1.flag = flag; o1.src = r; o1.rev = curA; o1.pool_off = rt_off; sz1 = baseA + 4u * nrt;
		} else if (0 < best_r) {
			int sB = best_r;
			flag_r |= 8 | 32;
			if (0 < rt[0]) { curB ^= 1; if (rt[nrt - 1] < 0) sB = -sB; }
			else { flag_r |= 16; for (int i = 0; i < nrt; ++i) rt[i] = -rt[i]; }
			o2.score = sB; o2.ntmpl = nrt; o2.flag = flag_r; o2.src = r + 1; o2.rev = curB; o2.pool_off = rt_off; sz2 = baseB + 4u * nrt;
		}
		res[r] = o1; res[r + 1] = o2;
		recsize[r] = sz1; recsize[r + 1] = sz2;
		if (sz1 || sz2) atomicAdd(&ctr[C_MAPPED], 1ull);
		return;
	}
	if (hc) {   // getFirstPen
		for (int i = 0; i < A.n_f; ++i) { rt[nrt++] = F1[i].x; best = max(best, F1[i].y); }
		for (int i = 0; i < A.n_r; ++i) { rt[nrt++] = -R1[i].x; best = max(best, R1[i].y); }
	}
	if (hc_r) {
		if (0 < best) {   // getSecondBestPen
			for (int i = 0; i < B.n_f; ++i) { bt[nbt++] = F2[i].x; best_r = max(best_r, F2[i].y); }
			for (int i = 0; i < B.n_r; ++i) { bt[nbt++] = -R2[i].x; best_r = max(best_r, R2[i].y); }
			int hits = 0;
			if (use_proxi) {   // getSecondProxiPen (savekmers.c:1514-1646)
				if (best_r) {
					int comp = 0;
					for (int i = 0; i < nrt; ++i) {
						const int t = rt[i];
						int sc = 0 < t ? list_score(R2, B.n_r, t) : list_score(F2, B.n_f, -t);
						if (0 < sc) { sc += i < A.n_f ? F1[i].y : R1[i - A.n_f].y; comp = max(comp, sc); }
					}
					if (best + best_r - PE <= comp) {
						const int thr = proxi_of(pf, comp);
						for (int i = 0; i < nrt; ++i) {
							const int t = rt[i];
							int sc = 0 < t ? list_score(R2, B.n_r, t) : list_score(F2, B.n_f, -t);
							if (0 < sc) {
								sc += i < A.n_f ? F1[i].y : R1[i - A.n_f].y;
								if (thr <= sc) { rt[hits++] = t; if (soft) atomicAdd(soft + abs(t), (unsigned long long)sc); }
							}
						}
					}
				}
				if (hits) { proper = true; nrt = hits; }
				else {
					int thr = proxi_of(pf, best);
					for (int i = 0; i < nrt; ++i) if (thr <= (i < A.n_f ? F1[i].y : R1[i - A.n_f].y)) rt[hits++] = rt[i];
					nrt = hits;
					hits = 0;
					thr = proxi_of(pf, best_r);
					for (int i = 0; i < nbt; ++i) {
						const int sc = i < B.n_f ? F2[i].y : R2[i - B.n_f].y;
						if (thr <= sc) { if (soft) atomicAdd(soft + abs(bt[i]), (unsigned long long)sc); bt[hits++] = bt[i]; }
					}
					nbt = hits;
				}
			} else {
			if (best_r) {
				int comp = max(0, best + best_r - PE);
				for (int i = 0; i < nrt; ++i) {
					const int t = rt[i];
					int sc = 0 < t ? list_score(R2, B.n_r, t) : list_score(F2, B.n_f, -t);   // the mate on the opposite strand
					if (0 < sc) {
						sc += i < A.n_f ? F1[i].y : R1[i - A.n_f].y;
						if (comp < sc) { comp = sc; hits = 1; rt[0] = t; }
						else if (comp == sc) rt[hits++] = t;
					}
				}
			}
			if (hits) { proper = true; nrt = hits; }
			else {
				for (int i = 0; i < nrt; ++i) if (best == (i < A.n_f ? F1[i].y : R1[i - A.n_f].y)) rt[hits++] = rt[i];
				nrt = hits;
				hits = 0;
				for (int i = 0; i < nbt; ++i) {
					const int t = bt[i];
					if (0 < t) { if (best_r == F2[i].y) bt[hits++] = t; }
					else if (best_r <= R2[i - B.n_f].y) bt[hits++] = t;
				}
				nbt = hits;
			}
			}
		} else if (use_proxi) {   // getF_Proxi on the second mate alone
			nrt = 0;
			for (int i = 0; i < n2; ++i) best_r = max(best_r, i < B.n_f ? F2[i].y : R2[i - B.n_f].y);
			const int thr = proxi_of(pf, best_r);
			for (int i = 0; i < n2; ++i) {
				const int t = i < B.n_f ? F2[i].x : -R2[i - B.n_f].x, sc = i < B.n_f ? F2[i].y : R2[i - B.n_f].y;
				if (thr <= sc) { rt[nrt++] = t; if (soft) atomicAdd(soft + abs(t), (unsigned long long)sc); }
			}
		} else {          // getF_Best: arg-max set of the second mate alone (written where regionTemplates is expected)
			nrt = 0;
			for (int i = 0; i < n2; ++i) {
				const int t = i < B.n_f ? F2[i].x : -R2[i - B.n_f].x, sc = i < B.n_f ? F2[i].y : R2[i - B.n_f].y;
				if (best_r < sc) { best_r = sc; nrt = 1; rt[0] = t; }
				else if (best_r == sc) rt[nrt++] = t;
			}
		}
	}
	// the reads are left reverse-complemented by the scan (savekmers.c:471); cur = 1 means "emit the reverse complement"
	int curA = A.scanned, curB = B.scanned;
	int flag = 65, flag_r = 129;
	const uint32_t baseA = 28u + 8u * ld_u32u(recA + 4) + 4u * ld_u32u(recA + 8) + (uint32_t)abs((int)ld_u32u(recA + 12));
	const uint32_t baseB = 28u + 8u * ld_u32u(recB + 4) + 4u * ld_u32u(recB + 8) + (uint32_t)abs((int)ld_u32u(recB + 12));
	const uint32_t rt_off = (uint32_t)po, bt_off = (uint32_t)po + (uint32_t)n1;
	if (0 < best && 0 < best_r) {
		if (proper) {
			flag |= 2; flag_r |= 2;
			const int comp = min(hc + hc_r, best + best_r);
			if (k <= comp || (lenA + lenB - comp - (k << 1)) < comp * k) {
				if (0 < rt[0]) {
					flag |= 32; flag_r |= 16; curA ^= 1;
					o1.score = best; o1.ntmpl = 0; o1.flag = flag; o1.src = r; o1.rev = curA; o1.pool_off = rt_off; sz1 = baseA;
					o2.score = best_r; o2.ntmpl = nrt; o2.flag = flag_r; o2.src = r + 1; o2.rev = curB; o2.pool_off = rt_off; sz2 = baseB + 4u * nrt;
				} else {
					flag |= 16; flag_r |= 32; curB ^= 1;
					for (int i = 0; i < nrt; ++i) rt[i] = -rt[i];
					o1.score = best_r; o1.ntmpl = 0; o1.flag = flag_r; o1.src = r + 1; o1.rev = curB; o1.pool_off = rt_off; sz1 = baseB;
					o2.score = best; o2.ntmpl = nrt; o2.flag = flag; o2.src = r; o2.rev = curA; o2.pool_off = rt_off; sz2 = baseA + 4u * nrt;
				}
			}
		} else {
			int h = min(hc, best), h_r = min(hc_r, best_r), sA = best, sB = best_r;
			h = k <= h || (lenA - h - k) < h * k;
			if (h) {
				if (0 < rt[0]) { curA ^= 1; if (rt[nrt - 1] < 0) sA = -sA; }
				else { flag |= 16; flag_r |= 32; for (int i = 0; i < nrt; ++i) rt[i] = -rt[i]; }
			}
			h_r = k <= h_r || (lenB - h_r - k) < h_r * k;
			if (h_r) {
				if (0 < bt[0]) { curB ^= 1; if (bt[nbt - 1] < 0) sB = -sB; }
				else { flag |= 32; flag_r |= 16; for (int i = 0; i < nbt; ++i) bt[i] = -bt[i]; }
			}
			if (h) { o1.score = sA; o1.ntmpl = nrt; o1.flag = flag; o1.src = r; o1.rev = curA; o1.pool_off = rt_off; sz1 = baseA + 4u * nrt; }
			if (h_r) { o2.score = sB; o2.ntmpl = nbt; o2.flag = flag_r; o2.src = r + 1; o2.rev = curB; o2.pool_off = bt_off; sz2 = baseB + 4u * nbt; }
		}
	} else if (0 < best) {
		const int h = min(hc, best);
		if (k <= h || (lenA - h - k) < h * k) {
			int sA = best;
			flag |= 8 | 32;
			if (0 < rt[0]) { curA ^= 1; if (rt[nrt - 1] < 0) sA = -sA; }
			else { flag |= 16; for (int i = 0; i < nrt; ++i) rt[i] = -rt[i]; }
			o1.score = sA; o1.ntmpl = nrt; o1.flag = flag; o1.src = r; o1.rev = curA; o1.pool_off = rt_off; sz1 = baseA + 4u * nrt;
		}
	} else if (0 < best_r) {
		const int h = min(hc_r, best_r);
		if (k <= h || (lenB - h - k) < h * k) {
			int sB = best_r;
			flag_r |= 8 | 32;
			if (0 < rt[0]) { curB ^= 1; if (rt[nrt - 1] < 0) sB = -sB; }
			else { flag_r |= 16; for (int i = 0; i < nrt; ++i) rt[i] = -rt[i]; }
			o2.score = sB; o2.ntmpl = nrt; o2.flag = flag_r; o2.src = r + 1; o2.rev = curB; o2.pool_off = rt_off; sz2 = baseB + 4u * nrt;
		}
	}
	res[r] = o1; res[r + 1] = o2;
	recsize[r] = sz1; recsize[r + 1] = sz2;
	if (sz1 || sz2) atomicAdd(&ctr[C_MAPPED], 1ull);
}

// ---------------------------------------------------------------- stage-2 record writer

// one warp per mapped read: header, sequence (forward copy or reverse complement, compdna.c:228),
// N list, template list, name bytes (ankers.c:30-50)
__global__ void __launch_bounds__(256) emit_records_kernel(const uint8_t *__restrict__ in, const uint32_t *__restrict__ rec_off,
		int nreads, const SeedRes *__restrict__ res, const uint32_t *__restrict__ out_off, const int32_t *__restrict__ pool,
		uint8_t *__restrict__ out) {
	const unsigned lane = threadIdx.x & 31;
	const int warps = (gridDim.x * blockDim.x) >> 5;
	for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nreads; r += warps) {
		const SeedRes rs = res[r];
		if (rs.score == 0) continue;
		const uint8_t *rec = in + rec_off[rs.src];
		const int seqlen = (int)ld_u32u(rec), words = (int)ld_u32u(rec + 4), nN = (int)ld_u32u(rec + 8);
		const int hdrlen = abs((int)ld_u32u(rec + 12));
		const uint8_t *seq = rec + 16, *N = seq + 8 * (size_t)words, *hdr = N + 4 * (size_t)nN;
		uint8_t *o = out + out_off[r];
		const bool rev = rs.rev != 0;
		if (lane < 7) {
			int32_t h = lane == 0 ? seqlen : lane == 1 ? words : lane == 2 ? nN : lane == 3 ? rs.score
			          : lane == 4 ? rs.ntmpl : lane == 5 ? hdrlen : rs.flag;
			st_u32b(o + 4 * lane, (uint32_t)h);
		}
		o += 28;
#pragma unroll 1
		for (int w = lane; w < words; w += 32) {
			uint64_t x;
			if (!rev) x = ld_u64u(seq + 8 * (size_t)w);
			else {
				x = rev2(~fwd32(seq, words, seqlen - 32 * (w + 1)));
				int c = seqlen - 32 * w;   // valid bases in this word
				if (c < 32) x &= ~0ull << (64 - 2 * c);
			}
			st_u32b(o + 8 * (size_t)w, (uint32_t)x);
			st_u32b(o + 8 * (size_t)w + 4, (uint32_t)(x >> 32));
		}
		o += 8 * (size_t)words;
#pragma unroll 1
		for (int i = lane; i < nN; i += 32) {
			uint32_t v = rev ? (uint32_t)(seqlen - 1 - (int)ld_u32u(N + 4 * (size_t)(nN - 1 - i))) : ld_u32u(N + 4 * (size_t)i);
			st_u32b(o + 4 * (size_t)i, v);
		}
		o += 4 * (size_t)nN;
#pragma unroll 1
		for (int i = lane; i < rs.ntmpl; i += 32) st_u32b(o + 4 * (size_t)i, (uint32_t)pool[rs.pool_off + i]);
		o += 4 * (size_t)rs.ntmpl;
#pragma unroll 1
		for (int i = lane; i < hdrlen; i += 32) o[i] = hdr[i];
	}
}

__global__ void lookup_kernel(KgHashView hv, const uint64_t *kmers, size_t n, int64_t *out) {
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t v = hash_lookup(hv, kmers[i]);
	out[i] = v == KG_MISS ? -1 : (int64_t)v;
}

// ---------------------------------------------------------------- host side

int kg_seed_free(kmagpu_db *db) {
	SeedBatch &b = db->seed;
	KgBuf *all[] = {&b.d_in, &b.d_off, &b.d_res, &b.d_pool, &b.d_recoff, &b.d_out, &b.d_ctr, &b.d_partial, &b.d_dense, &b.d_soft, &b.d_kinds, &b.d_mates, &b.d_pool2,
	                &b.h_off, &b.h_in, &b.h_out, &b.h_kinds, &b.d_chain, &b.d_regpool, &b.d_regs, &b.d_rsize, &b.d_partial2};
	for (KgBuf *x : all) x->release();
	return 0;
}

extern "C" int kmagpu_seed_upload(kmagpu_db *db, const void *stage1, size_t nbytes, int64_t *nreads_out) {
	if (!db || (!stage1 && nbytes)) { kmagpu_set_error("null argument"); return -1; }
	if (nbytes >= (1ull << 31)) { kmagpu_set_error("stage-1 batch of %zu bytes exceeds the 2 GiB per-call limit; split it", nbytes); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	SeedBatch &b = db->seed;
	b.h_off.pinned = true; b.h_kinds.pinned = true;
	const uint8_t *in = (const uint8_t *)stage1;
	// the bytes go first: the copy engine moves them while the host walks the record headers below
	if (b.d_in.reserve(nbytes + 64)) return -1;
	KG_CUDA(cudaEventRecord(db->ev[0], db->stream));
	if (nbytes) KG_CUDA(cudaMemcpyAsync(b.d_in.p, in, nbytes, cudaMemcpyHostToDevice, db->stream));
	KG_CUDA(cudaMemsetAsync((uint8_t *)b.d_in.p + nbytes, 0, 64, db->stream));
	// record boundaries: the only sequential step (16-byte header walk, loadFsa savekmers.c:50-92). Pairs: the first
	// mate is written with a negative header length (runinput.c:789), its mate follows.
	size_t guess = nbytes / 48 + 16;
	if (b.h_off.reserve(4 * (guess + 1)) || b.h_kinds.reserve(guess + 2)) return -1;
	uint32_t *off = (uint32_t *)b.h_off.p;
	uint8_t *kinds = (uint8_t *)b.h_kinds.p;
	size_t cap = std::min(b.h_off.cap / 4 - 1, b.h_kinds.cap - 2), n = 0, ip = 0, npairs = 0;
	bool mate = false;
	int32_t maxlen = 0;
	while (ip + 16 <= nbytes) {
		int32_t h[4];
		memcpy(h, in + ip, 16);
		if (h[0] < 0) break;   // a terminator was included
		size_t len = 16 + 8 * (size_t)(uint32_t)h[1] + 4 * (size_t)(uint32_t)h[2] + (size_t)abs(h[3]);
		if (h[1] < 0 || h[2] < 0 || ip + len > nbytes) { kmagpu_set_error("stage-1 stream is truncated or corrupt at byte %zu", ip); return -1; }
		if (kg_check_record(in + ip, 1, db->info.DB_size, ip)) return -1;
		if (n == cap) {
			KgBuf bigger, bigk; bigger.pinned = true; bigk.pinned = true;
			if (bigger.reserve(8 * (cap + 1)) || bigk.reserve(2 * (cap + 2))) return -1;
			memcpy(bigger.p, off, 4 * n); memcpy(bigk.p, kinds, n);
			b.h_off.release(); b.h_kinds.release();
			b.h_off = bigger; b.h_kinds = bigk;
			off = (uint32_t *)b.h_off.p; kinds = (uint8_t *)b.h_kinds.p;
			cap = std::min(b.h_off.cap / 4 - 1, b.h_kinds.cap - 2);
		}
		if (mate) { kinds[n] = 2; mate = false; }
		else if (h[3] < 0) { kinds[n] = 1; mate = true; ++npairs; }
		else kinds[n] = 0;
		off[n++] = (uint32_t)ip;
		maxlen = std::max(maxlen, h[0]);
		ip += len;
	}
	if (mate) { kmagpu_set_error("stage-1 stream ends inside a pair"); return -1; }
	off[n] = (uint32_t)ip;
	kinds[n] = 0;
	b.npairs = (int64_t)npairs;
	b.max_seqlen = maxlen;
	b.nreads = (int64_t)n;
	b.in_bytes = ip;
	b.ran = false;
	// what the reference counts (savekmers.c:183): one per single read, one per pair
	if (nreads_out) *nreads_out = (int64_t)(n - npairs);
	if (b.d_off.reserve(4 * (n + 1))) return -1;
	if (npairs) {
		if (b.d_kinds.reserve(n + 1)) return -1;
		KG_CUDA(cudaMemcpyAsync(b.d_kinds.p, kinds, n + 1, cudaMemcpyHostToDevice, db->stream));
	}
	KG_CUDA(cudaMemcpyAsync(b.d_off.p, off, 4 * (n + 1), cudaMemcpyHostToDevice, db->stream));
	KG_CUDA(cudaEventRecord(db->ev[1], db->stream));
	KG_CUDA(cudaStreamSynchronize(db->stream));
	return 0;
}

// soft proximity: the overflow list = every read, so that the dense pass alone maps the batch
static __global__ void seed_all_dense_kernel(uint32_t *ovf_list, int n, unsigned long long *ctr) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) ovf_list[i] = (uint32_t)i;
	if (i == 0) ctr[C_OVF] = (unsigned long long)n;
}

extern "C" int kmagpu_seed_run(kmagpu_db *db, const kmagpu_params *prm, kmagpu_seed_stats *stats) {
	if (!db || !prm) { kmagpu_set_error("null argument"); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	SeedBatch &b = db->seed;
	const int n = (int)b.nreads;
	if (stats) memset(stats, 0, sizeof(*stats));
	b.out_bytes = 0;
	b.ran = true;
	b.out_nrec = 0; b.out_recoff = nullptr;
	if (n == 0) return 0;
	// kmerScan (save_kmers_chain) only sees single reads; read pairs always go through save_kmers_pair (savekmers.c:196-199)
	const bool all_pairs = b.npairs > 0 && 2 * b.npairs == b.nreads;
	if (prm->kmerscan == 1 && b.npairs > 0 && !all_pairs) {
		kmagpu_set_error("chain mode: a batch mixes single reads and read pairs; hand them over in runs of one kind");
		return -1;
	}
	if (prm->kmerscan == 1 && !all_pairs) return kg_chain_run(db, prm, stats);
	if (prm->kmerscan != 0 && prm->kmerscan != 1) { kmagpu_set_error("kmerscan %d: only save_kmers (0) and save_kmers_chain (1) are built", prm->kmerscan); return -1; }
	SeedParams sp = {prm->M, prm->MM, prm->U, prm->W1, prm->exhaustive, 0, 1.0, nullptr};
	sp.proxi = fabs(prm->minFrac);       // stage 2 sees |minFrac| (kma.c:1605, kmers.c:133-141); the sign is stage 3's (soft proximity)
	sp.use_proxi = sp.proxi != 1.0;
	// soft proximity: a negative minFrac reaches stage 2 only in -mem_mode (kma.c:1605); the sums of this batch are collected in a
	// scratch array per attempt (a pool overflow redoes the batch) and join the image's sums once the batch is through
	const bool soft = prm->minFrac < 0 && sp.use_proxi && db->image->d_soft;
	if (prm->minFrac < 0 && sp.use_proxi && !soft) { kmagpu_set_error("soft proximity (minFrac < 0 in stage 2) needs kmagpu_softproxi_reset first"); return -1; }
	if (soft) {
		if (b.d_soft.reserve(8 * (size_t)db->info.DB_size)) return -1;
		sp.soft = (unsigned long long *)b.d_soft.p;
	}
	const int ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
	if (b.pool_cap < (size_t)n * 16 + 1024) b.pool_cap = (size_t)n * 16 + 1024;
	if (b.d_res.reserve(sizeof(SeedRes) * (size_t)n) || b.d_recoff.reserve(4 * (size_t)(2 * n + 2)) ||
	    b.d_ctr.reserve(8 * C_N) || b.d_partial.reserve(4 * (size_t)(ntiles + 1) + 8 * (size_t)n)) return -1;
	// persistent grid: as many CTAs as are resident at once (registers and shared memory decide)
	int per_sm = KG_MINB;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, seed_se_kernel<false, false>, KG_WARPS * 32, 0) != cudaSuccess || per_sm < 1) per_sm = KG_MINB;
	const int grid = db->sm_count * per_sm;
	// dense fallback scratch: (score, ext, candF, candR: int) + incl (byte) per template, per warp
	const int dense_grid = db->sm_count * 2;
	const size_t D = (size_t)db->info.DB_size + 1;
	const size_t dense_stride = ((16 * D + D) + 255) & ~(size_t)255;
	const size_t dense_bytes = dense_stride * (size_t)dense_grid * KG_WARPS;
	if (!b.d_dense.p) {
		if (b.d_dense.reserve(dense_bytes)) return -1;
		KG_CUDA(cudaMemsetAsync(b.d_dense.p, 0, b.d_dense.cap, db->stream));
	}
	const bool pe = b.npairs > 0;
	const uint8_t *kinds = pe ? (const uint8_t *)b.d_kinds.p : nullptr;
	if (pe) {
		if (b.pool2_cap < (size_t)n * 24 + 1024) b.pool2_cap = (size_t)n * 24 + 1024;
		if (b.pool_cap < (size_t)n * 24 + 1024) b.pool_cap = (size_t)n * 24 + 1024;
		if (b.d_mates.reserve(sizeof(MateRes) * ((size_t)n + 1))) return -1;
	}
	uint32_t *recsize = (uint32_t *)b.d_recoff.p, *recoff = recsize + n + 1;
	uint32_t *partial = (uint32_t *)b.d_partial.p, *ovf = partial + ntiles + 1, *nlist = ovf + n;
	unsigned long long *ctr = (unsigned long long *)b.d_ctr.p;
	int launches = 0;
	for (int attempt = 0;; ++attempt) {
		if (b.d_pool.reserve(4 * b.pool_cap)) return -1;
		if (pe && b.d_pool2.reserve(8 * b.pool2_cap)) return -1;
		KG_CUDA(cudaMemsetAsync(ctr, 0, 8 * C_N, db->stream));
		if (soft) KG_CUDA(cudaMemsetAsync(b.d_soft.p, 0, 8 * (size_t)db->info.DB_size, db->stream));
		KG_CUDA(cudaEventRecord(db->ev[2], db->stream));
		// the common shape (hashed table, 16-bit lists) runs the specialised kernel and sets reads with N's aside for the
		// generic one; any other database runs the generic kernel on every read
		const bool common = !db->hv.mega && db->hv.values_s;
		if (soft) {
			// every read takes the dense pass: it cannot overflow, so no read is scanned twice and its kept templates add once
			seed_all_dense_kernel<<<(n + 255) / 256, 256, 0, db->stream>>>(ovf, n, ctr);
		} else {
		if (common)
			seed_se_kernel<false, false><<<grid, KG_WARPS * 32, 0, db->stream>>>(db->hv, sp, (const uint8_t *)b.d_in.p,
				(const uint32_t *)b.d_off.p, n, (SeedRes *)b.d_res.p, recsize, (int32_t *)b.d_pool.p,
				(unsigned long long)b.pool_cap, ctr, ovf, nullptr, 0, kinds, (MateRes *)b.d_mates.p, (int2 *)b.d_pool2.p,
				(unsigned long long)b.pool2_cap, nlist, 0);
		seed_se_kernel<false, true><<<common ? db->sm_count * 2 : grid, KG_WARPS * 32, 0, db->stream>>>(db->hv, sp, (const uint8_t *)b.d_in.p,
			(const uint32_t *)b.d_off.p, n, (SeedRes *)b.d_res.p, recsize, (int32_t *)b.d_pool.p,
			(unsigned long long)b.pool_cap, ctr, ovf, nullptr, 0, kinds, (MateRes *)b.d_mates.p, (int2 *)b.d_pool2.p,
			(unsigned long long)b.pool2_cap, nlist, common ? 1 : 0);
		}
		seed_se_kernel<true, true><<<dense_grid, KG_WARPS * 32, 0, db->stream>>>(db->hv, sp, (const uint8_t *)b.d_in.p,
			(const uint32_t *)b.d_off.p, n, (SeedRes *)b.d_res.p, recsize, (int32_t *)b.d_pool.p,
			(unsigned long long)b.pool_cap, ctr, ovf, (uint8_t *)b.d_dense.p, dense_stride, kinds, (MateRes *)b.d_mates.p,
			(int2 *)b.d_pool2.p, (unsigned long long)b.pool2_cap, nlist, 0);
		if (pe) {
			pair_select_kernel<<<(n + 127) / 128, 128, 0, db->stream>>>((const uint8_t *)b.d_in.p, (const uint32_t *)b.d_off.p, n, kinds,
				(const MateRes *)b.d_mates.p, (const int2 *)b.d_pool2.p, (int32_t *)b.d_pool.p, (unsigned long long)b.pool_cap, ctr,
				(SeedRes *)b.d_res.p, recsize, db->hv.kmersize, prm->PE, prm->apm, sp.use_proxi, sp.proxi, sp.soft);
			++launches;
		}
		KG_CUDA(cudaEventRecord(db->ev[3], db->stream));
		kg_exscan(recsize, n, recoff, partial, ctr + C_TOTAL, db->stream);
		launches += 6;
		unsigned long long h[C_N];
		KG_CUDA(cudaMemcpyAsync(h, ctr, 8 * C_N, cudaMemcpyDeviceToHost, db->stream));
		KG_CUDA(cudaStreamSynchronize(db->stream));
		KG_CUDA(cudaGetLastError());
		if (h[C_POOLFAIL]) {   // template pool too small: grow and redo (rare)
			if (attempt > 4) { kmagpu_set_error("template pool overflow persists"); return -1; }
			b.pool_cap = std::max(b.pool_cap, (size_t)h[C_POOL] + 1024);
			b.pool2_cap = std::max(b.pool2_cap, (size_t)h[C_POOL2] + 1024);
			continue;
		}
		KG_SCAN_FITS(h[C_TOTAL], "the stage-2 stream");
		if (soft) kg_softproxi_accumulate(db, (const unsigned long long *)b.d_soft.p);
		b.out_bytes = (size_t)h[C_TOTAL];
		b.out_nrec = n; b.out_recoff = recoff;
		if (b.d_out.reserve(b.out_bytes + 64)) return -1;
		emit_records_kernel<<<kg_wave_grid(emit_records_kernel, 256, db->sm_count), 256, 0, db->stream>>>((const uint8_t *)b.d_in.p, (const uint32_t *)b.d_off.p,
			n, (const SeedRes *)b.d_res.p, recoff, (const int32_t *)b.d_pool.p, (uint8_t *)b.d_out.p);
		KG_CUDA(cudaEventRecord(db->ev[4], db->stream));
		KG_CUDA(cudaStreamSynchronize(db->stream));
		KG_CUDA(cudaGetLastError());
		++launches;
		if (stats) {
			stats->reads = n; stats->mapped = (int64_t)h[C_MAPPED]; stats->read_words = (int64_t)h[C_WORDS];
			stats->lookups = (int64_t)h[C_LOOKUPS]; stats->hits = (int64_t)h[C_HITS];
			stats->list_fetches = (int64_t)h[C_LISTS]; stats->list_ids = (int64_t)h[C_LISTIDS];
			stats->overflow_reads = (int64_t)h[C_OVF];
			cudaEventElapsedTime(&stats->ms_seed, db->ev[2], db->ev[3]);
			cudaEventElapsedTime(&stats->ms_emit, db->ev[3], db->ev[4]);
			cudaEventElapsedTime(&stats->ms_total, db->ev[2], db->ev[4]);
			stats->launches = launches;
		}
		return 0;
	}
}

// device view of the stage-2 stream of the last run, for kmagpu_align_from_seed (records of unmapped reads are empty)
int kg_seed_device_output(kmagpu_db *db, const uint8_t **out, const uint32_t **rec_off, int64_t *nreads, size_t *bytes) {
	SeedBatch &b = db->seed;
	if (!b.ran) { kmagpu_set_error("kmagpu_align_from_seed before kmagpu_seed_run"); return -1; }
	*out = (const uint8_t *)b.d_out.p;
	*rec_off = b.out_recoff;
	*nreads = b.out_nrec;
	*bytes = b.out_bytes;
	return 0;
}

extern "C" int kmagpu_seed_download(kmagpu_db *db, void *stage2_out, size_t out_cap, size_t *out_bytes) {
	if (!db) { kmagpu_set_error("null argument"); return -1; }
	SeedBatch &b = db->seed;
	if (!b.ran) { kmagpu_set_error("kmagpu_seed_download before kmagpu_seed_run"); return -1; }
	if (out_bytes) *out_bytes = b.out_bytes;
	if (b.out_bytes > out_cap) { kmagpu_set_error("stage-2 output needs %zu bytes, caller gave %zu", b.out_bytes, out_cap); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	if (b.out_bytes) {
		KG_CUDA(cudaMemcpyAsync(stage2_out, b.d_out.p, b.out_bytes, cudaMemcpyDeviceToHost, db->stream));
		KG_CUDA(cudaStreamSynchronize(db->stream));
	}
	return 0;
}

extern "C" int kmagpu_seed_batch(kmagpu_db *db, const kmagpu_params *p, const void *stage1, size_t nbytes,
                                 void *stage2_out, size_t out_cap, size_t *out_bytes, int64_t *nreads,
                                 kmagpu_seed_stats *stats) {
	if (kmagpu_seed_upload(db, stage1, nbytes, nreads)) return -1;
	if (kmagpu_seed_run(db, p, stats)) return -1;
	if (kmagpu_seed_download(db, stage2_out, out_cap, out_bytes)) return -1;
	if (stats) {
		cudaEventElapsedTime(&stats->ms_h2d, db->ev[0], db->ev[1]);
	}
	return 0;
}

extern "C" int kmagpu_lookup_batch(kmagpu_db *db, const uint64_t *kmers, size_t n, int64_t *out) {
	if (!db || !kmers || !out) { kmagpu_set_error("null argument"); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	uint64_t *dk = nullptr; int64_t *dout = nullptr;
	KG_CUDA(cudaMalloc(&dk, 8 * n + 8));
	KG_CUDA(cudaMalloc(&dout, 8 * n + 8));
	KG_CUDA(cudaMemcpy(dk, kmers, 8 * n, cudaMemcpyHostToDevice));
	if (n) lookup_kernel<<<(unsigned)((n + 255) / 256), 256, 0, db->stream>>>(db->hv, dk, n, dout);
	KG_CUDA(cudaStreamSynchronize(db->stream));
	KG_CUDA(cudaMemcpy(out, dout, 8 * n, cudaMemcpyDeviceToHost));
	cudaFree(dk); cudaFree(dout);
	KG_CUDA(cudaGetLastError());
	return 0;
}
