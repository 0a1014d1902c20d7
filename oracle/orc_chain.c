/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the long-read default of KMA stage 2: save_kmers_chain.
 *
 * A plain-C restatement of what the reference computes per read when no -1t1 is given
 * (reference: savekmers.c:5127-5944 `save_kmers_chain`; kmeranker.c:83 `getBestChainTemplates`, :372 `pruneAnkers`,
 * :398 `getBestAnkerScore`, :480 `getTieAnkerScore`, :512 `chooseChain`, :57 `mrchain`; seqmenttree.c:107-232;
 * qseqs.c:41 `insertKmerBound`). Default selection functions (kmeranker.c:25-30) and, after orc_chain_set_lc(1), the
 * length-corrected ones -lc binds (kma.c:694-700: ankerScoreLen, testExtensionScoreLen, proxiTestBestScoreLen,
 * getBestAnkerScoreLen kmeranker.c:432, getTieAnkerScoreLen :496, and the swap of savekmers.c:5657-5664); no -proxi.
 *
 * It exists only so that tests can compare the CUDA chain kernel with something that is pinned byte for byte to the
 * unmodified reference (`kma -s2` without -1t1, tests/test_oracle_chain.py). Nothing outside tests/, smoke() and
 * bench.py's CPU legs may call it.
 *
 * Reference behaviours kept on purpose (the test data exercises them):
 *  - the last anker of a strand ends at `seqlen - gaps` (position of its last k-mer when the scan runs to the end,
 *    savekmers.c:5329) while every other anker ends at last k-mer + k + 1 (:5287);
 *  - the reverse strand is scanned "in forward notation"; after an N the reverse cursor restarts at `seqlen - j`
 *    (savekmers.c:5443) instead of `seqlen - k - j`, so the k-mers of every later stretch are taken k bases off;
 *  - the first anker of a strand ties with itself (`++ties`, savekmers.c:5617-5626), `ties` is not reset between
 *    strands;
 *  - the segment tree's split insertion re-uses one node for both halves (seqmenttree.c:141-153).
 * Not reproducible (undefined behaviour in the reference, excluded from the parity surface): an N inside the first k
 * bases makes the shifted reverse cursor read past the packed read; more than 31 emitted regions on one read reach
 * the segment tree's resize, which copies child links into freed memory (seqmenttree.c:53-66). The first is computed
 * with zero bits past the end of the read, the second is reported through the return value instead of guessed.
 */
#include "orc.h"
#include <stdlib.h>
#include <string.h>

typedef struct { int score, weight, score_len, len_len, start, end; int64_t vals; int next; } ank_t;

static int g_lc = 0;
void orc_chain_set_lc(int lc) { g_lc = lc; }   /* -lc (kma.c:694): length-corrected anker selection */

typedef struct {
	const orc_db *db; const orc_params *p;
	int *score, *ext; unsigned char *incl;   /* Score[], extendScore[], include[] (savekmers.c:134-147) */
	int k, seqlen;
} cctx;

static inline uint64_t kmer_at(const uint64_t *seq, int pos, int k) {
	int w = pos >> 5, b = (pos & 31) << 1, sh = 64 - 2 * k;
	uint64_t x = seq[w] << b;
	if (b > sh) x |= seq[w + 1] >> (64 - b);
	return x >> sh;
}

static inline int list_at(const orc_db *db, int64_t off, int i) {   /* i-th template id of the list at `off` */
	return db->values_short ? ((const uint16_t *)db->values)[off + 1 + i] : (int)((const uint32_t *)db->values)[off + 1 + i];
}
static inline int list_n(const orc_db *db, int64_t off) {
	return db->values_short ? ((const uint16_t *)db->values)[off] : (int)((const uint32_t *)db->values)[off];
}

/* ---------------------------------------------------------------- ankers of one strand (savekmers.c:5227-5448) */
static int find_ankers(const cctx *c, const uint64_t *seq, const int32_t *sN, const int32_t *fN, int nN, int reverse,
                       int exhaustive, ank_t *A, orc_stats *st) {
	const orc_db *db = c->db;
	const int k = c->k, seqlen = c->seqlen, seqend = seqlen - k + 1, M = c->p->M, MM = c->p->MM;
	int hit = exhaustive, n = 0, cur = 0, Ms = 0, MMs = 0, gaps = 0, j = 0;
	int64_t last = -1;

	A[0].start = 0; A[0].end = 0; A[0].vals = -1; A[0].next = -1;
	/* quick check in this strand's own coordinates: every k-th k-mer of every N-free stretch until the first hit */
	for (int seg = 0, s = 0; seg <= nN && !hit; ++seg) {
		int e = seg < nN ? sN[seg] : seqlen;
		for (int q = s; q < e - k + 1 && !hit; q += k) {
			if (st) st->lookups++;
			hit = orc_lookup(db, kmer_at(seq, q, k)) >= 0;
		}
		s = e + 1;
	}
	if (!hit) return 0;

	int rcpos = seqlen - k;   /* reverse strand: cursor into the reverse-complemented words */
	for (int seg = 0; seg <= nN && j < seqend; ++seg) {
		const int e = seg < nN ? fN[seg] : seqlen;
		for (; j + k <= e; ++j, --rcpos) {
			const int64_t off = orc_lookup(db, kmer_at(seq, reverse ? rcpos : j, k));
			if (st) st->lookups++;
			if (off < 0) { ++gaps; continue; }
			if (st) st->hits++;
			if (off == last && gaps == 0) ++Ms;
			else if (off == last && gaps == k) { Ms += k; ++MMs; }
			else {
				if (last >= 0) {   /* close the running anker and link it to the next one */
					A[cur].weight = Ms * M + MMs * MM;
					A[cur].end = j - gaps + k;
					A[cur].next = cur + 1;
					++cur;
				}
				A[cur].start = j; A[cur].vals = off; A[cur].next = -1;
				if (st && off != last) st->list_fetches++, st->list_ids += list_n(db, off);
				last = off; Ms = k; MMs = 0; ++n;
			}
			gaps = 0;
		}
		gaps += e + 1 - j;   /* gap over the N */
		j = e + 1;
		rcpos = seqlen - j;  /* (sic) savekmers.c:5443 */
	}
	if (last >= 0) { A[cur].weight = Ms * M + MMs * MM; A[cur].end = seqlen - gaps; }
	return n;
}

/* score of chaining an anker to the previous anker of the same template across `gaps` bases
 * (savekmers.c:5524-5552 = kmeranker.c:154-187; mlen == kmersize) */
static int link_score(const orc_params *p, int k, int gaps, int weight) {
	if (gaps == -k) return weight - (k - 1) * p->M;
	if (gaps == 0) return weight + p->MM;
	if (0 < gaps) {
		int mm, m;
		if (gaps <= 2) { mm = gaps; m = 0; }
		else {
			mm = gaps / k + (gaps % k ? 1 : 0); if (mm < 2) mm = 2;
			m = gaps - mm; if (k < m) m = k; if (mm < m) m = mm;
		}
		const int a = p->W1 + (gaps - 1) * p->U, b = mm * p->MM + m * p->M;
		return weight + (a <= b ? b : a);
	}
	return weight + gaps * p->M - (gaps + 1) * p->U + p->W1;
}

/* getBestChainTemplates (kmeranker.c:83-233): walk back from `src`, re-scoring its templates anker by anker until one
 * of them reproduces src's score at a chain start; bests[0] = count, bests[1..] = the templates that reach it.
 * Marks the walked ankers as used (score = 0), returns the anker the chain starts at (NULL if no template is left).
 * `lo` guards the walk (the reference has no bound; running below the array is reported as -2 through *err). */
static ank_t *chain_templates(cctx *c, ank_t *src, ank_t *lo, int *bests, int *err) {
	const orc_db *db = c->db; const orc_params *p = c->p;
	const int k = c->k;
	if (!src) return 0;
	int more = 0, nl = list_n(db, src->vals);
	bests[0] = nl;
	for (int i = nl; i >= 1; --i) {
		const int t = list_at(db, src->vals, i - 1);
		bests[i] = t;
		if (++c->incl[t] == 1) more = 1;
	}
	const int bestScore = g_lc ? src->score_len : src->score;   /* kmerAnkerScore */
	const int target_len = src->len_len, q_len = c->seqlen;
	const int32_t *lengths = db->lengths;
	ank_t *prev = src;
	for (ank_t *node = src; more; --node) {
		if (node < lo) { *err = -2; break; }
		nl = list_n(db, node->vals);
		const int start = node->start, end = node->end;
		for (int i = nl - 1; i >= 0; --i) {
			const int t = list_at(db, node->vals, i);
			if (!c->incl[t]) continue;
			int score = c->score[t];
			const int pos = c->ext[t];
			if (pos == 0) score = node->weight;
			else {
				score += link_score(p, k, pos - end, node->weight);
				node->score = 0;   /* used */
			}
			if (bestScore <= score) {   /* does this anker open the chain? */
				int open = score;
				if (node->start) {
					int g = p->W1 + (node->start - 1) * p->U;
					open = score + (p->Wl < g ? g : p->Wl);
				}
				/* testExtension: under -lc only a template of the anker's own corrected length closes the chain */
				if (open == bestScore && (!g_lc || (q_len < lengths[t] ? q_len : lengths[t]) == target_len)) { score = bestScore; more = 0; prev = node; }
			}
			c->ext[t] = start;
			c->score[t] = score;
		}
	}
	int j = 0;
	for (int i = 1; i <= bests[0]; ++i) {
		const int t = bests[i];
		int ok = bestScore <= c->score[t];   /* proxiTestBest, proxi == 1.0 */
		if (g_lc && !ok) ok = (double)bestScore / target_len * (q_len < lengths[t] ? q_len : lengths[t]) <= c->score[t];
		if (c->incl[t] == 1 && ok) bests[++j] = t;
		c->score[t] = 0; c->incl[t] = 0; c->ext[t] = 0;
	}
	bests[0] = j;
	return j ? prev : 0;
}

/* getProxiChainTemplates (kmeranker.c:235-370, bound by -proxi): the walk back scores EVERY template it meets on the way
 * (not only those of src's own list), stops where one of them reproduces src's score at a chain start, and keeps the
 * templates within minFrac of that score that are not marked in include[] (the marks of the tie path). A template met
 * on an anker that starts at query position 0 keeps extendScore == 0 and is listed again when it is met again. */
static ank_t *chain_templates_proxi(cctx *c, ank_t *src, ank_t *lo, int *bests, int *err) {
	const orc_db *db = c->db; const orc_params *p = c->p;
	const int k = c->k;
	if (!src) return 0;
	bests[0] = 0;
	const int bestScore = g_lc ? src->score_len : src->score;
	const double proxiScore = orc_get_proxi() * bestScore;
	const int target_len = src->len_len, q_len = c->seqlen;
	const int32_t *lengths = db->lengths;
	ank_t *prev = src;
	int more = 1;
	for (ank_t *node = src; more; --node) {
		if (node < lo) { *err = -2; break; }
		const int nl = list_n(db, node->vals);
		const int start = node->start, end = node->end;
		for (int i = nl - 1; i >= 0; --i) {
			const int t = list_at(db, node->vals, i);
			int score = c->score[t];
			const int pos = c->ext[t];
			if (pos == 0) { score = node->weight; bests[++bests[0]] = t; }
			else {
				score += link_score(p, k, pos - end, node->weight);
				node->score = 0;
			}
			if (bestScore <= score) {
				int open = score;
				if (node->start) {
					int g = p->W1 + (node->start - 1) * p->U;
					open = score + (p->Wl < g ? g : p->Wl);
				}
				if (open == bestScore && (!g_lc || (q_len < lengths[t] ? q_len : lengths[t]) == target_len)) { score = bestScore; more = 0; prev = node; }
			}
			c->ext[t] = start;
			c->score[t] = score;
		}
	}
	int j = 0;
	for (int i = 1; i <= bests[0]; ++i) {
		const int t = bests[i];
		int ok = proxiScore <= c->score[t];   /* proxiTestBest */
		if (g_lc && !ok) ok = proxiScore / target_len * (q_len < lengths[t] ? q_len : lengths[t]) <= c->score[t];
		if (!c->incl[t] && ok) { bests[++j] = t; if (orc_get_soft_proxi()) orc_get_soft_proxi()[t] += c->score[t]; }
		c->score[t] = 0; c->ext[t] = 0; c->incl[t] = 0;
	}
	bests[0] = j;
	return prev;
}

static ank_t *get_chain_templates(cctx *c, ank_t *src, ank_t *lo, int *bests, int *err) {   /* getChainTemplates */
	return orc_get_proxi() != 1.0 ? chain_templates_proxi(c, src, lo, bests, err) : chain_templates(c, src, lo, bests, err);
}

static int prune(ank_t *A, int k) {   /* pruneAnkers (kmeranker.c:372): unlink ankers scoring below k; head or -1 */
	int i = 0;
	while (A[i].score < k && (i = A[i].next) >= 0);
	if (i < 0) return -1;
	int prev = i, node = i;
	while ((node = A[node].next) >= 0) if (k <= A[node].score) { A[prev].next = node; prev = node; }
	A[prev].next = -1;
	return i;
}

static ank_t *best_anker(ank_t *A, int *head, unsigned *ties) {   /* getBestAnkerScore (kmeranker.c:398) */
	*ties = 0;
	int prev = *head;
	while (prev >= 0 && A[prev].score == 0) prev = A[prev].next;
	*head = prev;
	if (prev < 0) return 0;
	int best = prev, node = A[prev].next;
	while (node >= 0) {
		if (A[node].score) {
			if (A[best].score < A[node].score) { best = node; *ties = 0; }
			else if (A[best].score == A[node].score) { best = node; ++*ties; }
			A[prev].next = node; prev = node;
		}
		node = A[node].next;
	}
	A[prev].next = -1;
	return A + best;
}

static ank_t *best_anker_len(ank_t *A, int *head, unsigned *ties) {   /* getBestAnkerScoreLen (kmeranker.c:432) */
	*ties = 0;
	int prev = *head;
	while (prev >= 0 && A[prev].score == 0) prev = A[prev].next;
	*head = prev;
	if (prev < 0) return 0;
	int best = prev, node = A[prev].next;
	while (node >= 0) {
		if (A[node].score) {
			double sl = A[node].score_len;
			if (A[node].len_len != A[best].len_len) { sl /= A[node].len_len; sl *= A[best].len_len; }
			if (A[best].score_len < sl) { best = node; *ties = 0; }
			else if (A[best].score_len == sl) {
				if (A[best].score_len < A[node].score_len) { best = node; *ties = 0; }
				else if (A[best].score_len == A[node].score_len) { best = node; ++*ties; }
			}
			A[prev].next = node; prev = node;
		}
		node = A[node].next;
	}
	A[prev].next = -1;
	return A + best;
}

static ank_t *tie_anker(int stop, ank_t *src, const ank_t *best) {   /* getTieAnkerScore (kmeranker.c:480), ...ScoreLen (:496) */
	if (!src || src->start <= stop) return 0;
	while (stop < (--src)->start)
		if (g_lc ? (src->score_len == best->score_len && src->len_len == best->len_len) : src->score == best->score) return src;
	return 0;
}

/* chooseChain (kmeranker.c:512-592) */
static int choose_chain(const ank_t *f, const ank_t *r, int cs, int cs_r, double coverT, int *Start, int *Len) {
	const double proxi = orc_get_proxi();
	int rc, start, end;
	if (proxi == 1.0) rc = r->score < f->score ? 1 : f->score < r->score ? 2 : 3;
	else if (r->score <= f->score) rc = (proxi * f->score <= r->score) ? 3 : 1;   /* the other strand within the proximity */
	else rc = (proxi * r->score <= f->score) ? 3 : 2;
	if (rc == 1) { start = cs; end = f->end; }
	else if (rc == 2) { start = cs_r; end = r->end; }
	else if (f->end < cs_r) { start = cs; end = f->end; rc = 1; }
	else if (r->end < cs) { start = cs_r; end = r->end; rc = 2; }
	else if (cs <= cs_r && r->end <= f->end) { start = cs; end = f->end; }
	else if (cs_r <= cs && f->end <= r->end) { start = cs_r; end = r->end; }
	else if (r->end < f->end) {
		int a = f->end - cs, b = r->end - cs_r, m = a < b ? a : b;
		start = cs_r;
		if (coverT * m <= (double)((unsigned)r->end - (unsigned)cs)) end = f->end;
		else { end = r->end; rc = 2; }
	} else {
		int a = f->end - cs, b = r->end - cs_r, m = a < b ? a : b;
		start = cs;
		if (coverT * m <= (double)((unsigned)f->end - (unsigned)cs_r)) end = r->end;
		else { end = f->end; rc = 1; }
	}
	*Start = start; *Len = end - start;
	return rc;
}

/* mrchain (kmeranker.c:57): with -mrc, templates shorter than mrc * maplen are dropped when the read is shorter too */
static int mr_chain(int *bests, const int32_t *lengths, int q_len, int maplen, double mrc) {
	if (mrc) {
		if (q_len < mrc * maplen) {
			int n = 0;
			for (int i = 1; i <= bests[0]; ++i) if (mrc * maplen <= lengths[bests[i]]) bests[++n] = bests[i];
			return (bests[0] = n);
		}
	}
	return 1;
}

/* ---------------------------------------------------------------- segment tree of used query intervals */
#define ST_CAP 64
typedef struct { unsigned start[ST_CAP + 2], end[ST_CAP + 2], cov[ST_CAP + 2]; int b0[ST_CAP + 2], b1[ST_CAP + 2]; int n; } stree;

static unsigned st_add(stree *T, int root, int node) {   /* addSeqmentTrees (seqmenttree.c:107-181) */
	if (T->b0[root] >= 0) {
		if (T->start[node] < T->start[root] && T->end[root] < T->end[node]) {
			T->start[root] = T->start[node]; T->end[root] = T->end[node]; T->cov[root] = T->cov[node];
			T->cov[node] = 0; T->b0[root] = -1;
			return T->cov[root];
		} else if (T->end[root] < T->end[node]) T->end[root] = T->end[node];
		else if (T->start[node] < T->start[root]) T->start[root] = T->start[node];
		unsigned pos = T->start[T->b1[root]];
		if (T->end[node] < pos) T->cov[root] = T->cov[T->b1[root]] + st_add(T, T->b0[root], node);
		else if (pos <= T->start[node]) T->cov[root] = T->cov[T->b0[root]] + st_add(T, T->b1[root], node);
		else {   /* split: the same node serves both halves */
			pos = T->start[node];
			T->start[node] = T->end[T->b0[root]] + 1;
			T->cov[node] = T->end[node] - T->start[node];
			unsigned right = st_add(T, T->b1[root], node);
			T->start[node] = pos;
			T->end[node] = T->end[T->b0[root]];
			T->cov[node] = T->end[node] - T->start[node];
			T->cov[root] = right + st_add(T, T->b0[root], node);
		}
	} else if (T->end[node] < T->start[root] || T->end[root] < T->start[node]) {   /* disjoint leaf: bud */
		const int bud = node + 1;
		T->start[bud] = T->start[root]; T->end[bud] = T->end[root]; T->cov[bud] = T->cov[root]; T->b0[bud] = -1;
		if (T->end[node] < T->start[root]) { T->start[root] = T->start[node]; T->b0[root] = node; T->b1[root] = bud; }
		else { T->end[root] = T->end[node]; T->b0[root] = bud; T->b1[root] = node; }
		T->cov[root] += T->cov[node];
	} else {   /* overlapping leaf: extend */
		if (T->start[node] < T->start[root]) T->start[root] = T->start[node];
		if (T->end[root] < T->end[node]) T->end[root] = T->end[node];
		T->cov[node] = 0;
		T->cov[root] = T->end[root] - T->start[root];
	}
	return T->cov[root];
}

static int st_grow(stree *T, unsigned start, unsigned end) {   /* growSeqmentTree (seqmenttree.c:183) */
	if (ST_CAP <= T->n + 2) return -1;   /* the reference's resize is undefined behaviour: refuse */
	if (T->n == 0) {
		T->n = 1; T->start[0] = start; T->end[0] = end; T->cov[0] = end - start; T->b0[0] = T->b1[0] = -1;
		return 0;
	}
	const int node = T->n;
	T->start[node] = start; T->end[node] = end; T->cov[node] = end - start; T->b0[node] = -1;
	T->cov[0] = st_add(T, 0, node);
	if (T->cov[node]) T->n += 2;
	return 0;
}

static unsigned st_query(const stree *T, int s, unsigned start, unsigned end) {   /* queSeqmentTree (seqmenttree.c:211) */
	if (end < T->start[s] || T->end[s] < start) return 0;
	if (start <= T->start[s] && T->end[s] <= end) return T->cov[s];
	if (T->b0[s] >= 0) return st_query(T, T->b0[s], start, end) + st_query(T, T->b1[s], start, end);
	if (T->start[s] <= start && end <= T->end[s]) return end - start;
	if (T->start[s] <= start && start < T->end[s]) return T->end[s] - start;
	if (T->start[s] < end && end <= T->end[s]) return end - T->start[s];
	return 0;
}

/* ---------------------------------------------------------------- records */
static size_t emit_chain_record(uint8_t *out, const uint64_t *seq, int seqlen, const int32_t *N, int nN, int score,
                                const int *tmpl, int ntmpl, const uint8_t *hdr, int hdrlen, int b_start, int b_end) {
	int32_t h[7] = {seqlen, (seqlen + 31) >> 5, nN, score, ntmpl, hdrlen + 9, 0};
	int32_t bound[2] = {b_start, b_end};
	uint8_t *o = out;
	memcpy(o, h, 28); o += 28;
	memcpy(o, seq, 8 * (size_t)h[1]); o += 8 * (size_t)h[1];
	memcpy(o, N, 4 * (size_t)nN); o += 4 * (size_t)nN;
	memcpy(o, tmpl, 4 * (size_t)ntmpl); o += 4 * (size_t)ntmpl;
	memcpy(o, hdr, hdrlen); o += hdrlen;
	*o++ = 0; memcpy(o, bound, 8); o += 8;   /* insertKmerBound (qseqs.c:41) */
	return o - out;
}

typedef struct {
	cctx c;
	ank_t *VF, *VR; size_t acap;
	int *bt, *bt_r;
	uint64_t *w[2]; int32_t *N[2]; size_t wcap, ncap;
	int minlen; double mrs, coverT, mrc;
} chain_ws;

/* One read. Returns bytes appended to `out` (>= 0) or a negative error: -1 output full, -2 back-walk left the anker
 * array, -3 segment tree resize. */
static int64_t chain_read(chain_ws *W, int seqlen, int nN, const uint8_t *hdr, int hdrlen, uint8_t *out, size_t cap, orc_stats *st) {
	cctx *c = &W->c;
	const orc_params *p = c->p;
	const int k = c->k;
	const int32_t *lengths = c->db->lengths;
	const uint64_t *fw = W->w[0], *rw = W->w[1];
	const int32_t *fN = W->N[0], *rN = W->N[1];
	ank_t *VF = W->VF, *VR = W->VR;
	int *bt = W->bt, *bt_r = W->bt_r, err = 0;
	size_t op = 0;
	stree T; T.n = 0;
	c->seqlen = seqlen;

	const unsigned nF = (unsigned)find_ankers(c, fw, fN, fN, nN, 0, p->exhaustive, VF, st);
	const unsigned nR = (unsigned)find_ankers(c, rw, rN, fN, nN, 1, p->exhaustive, VR, st);
	if (!nF && !nR) return 0;

	/* chaining DP over the ankers of each strand (savekmers.c:5457-5640) */
	ank_t *best = 0, *best_r = VF, *best_len = 0, *best_len_r = VF;
	unsigned ties = 0, ties_len = 0;
	VF[0].score = 0;
	bt[0] = 0; bt_r[0] = 0;
	for (int pass = 0; pass < 2; ++pass) {
		ank_t *V = pass ? VR : VF;
		int *bests = pass ? bt_r : bt;
		const unsigned cnt = pass ? nR : nF;
		if (pass) { V[0].score = 0; V[0].score_len = 0; V[0].len_len = 1; best = best_r; best_r = V; best_len = best_len_r; best_len_r = V; }
		bests[0] = 0;
		for (unsigned a = 0; a < cnt; ++a) {
			ank_t *node = V + a;
			const int start = node->start, end = node->end;
			node->score = 0; node->score_len = 0; node->len_len = 1;
			for (int i = list_n(c->db, node->vals) - 1; i >= 0; --i) {
				const int t = list_at(c->db, node->vals, i);
				int score;
				if (!c->incl[t]) {
					c->incl[t] = 1;
					bests[++bests[0]] = t;
					if (start) {
						int g = p->W1 + (start - 1) * p->U;
						score = node->weight + (p->Wl < g ? g : p->Wl);
					} else score = node->weight;
				} else {
					score = c->score[t] + link_score(p, k, start - c->ext[t], node->weight);
					if (score < 0) {   /* restarting the chain here may be better */
						int test = start ? p->W1 + (start - 1) * p->U : 0;
						if (test < p->Wl) test = p->Wl;
						if (score < test + node->weight) score = test + node->weight;
					}
				}
				if (node->score < score) node->score = score;
				int len_len = lengths[t];
				if (seqlen < len_len) len_len = seqlen;
				double sl = score;
				if (node->len_len != len_len) { sl /= len_len; sl *= node->len_len; }
				if (node->score_len < sl || (node->score_len == sl && node->score_len < score)) {
					node->score_len = score; node->len_len = len_len;
				}
				c->score[t] = score;
				c->ext[t] = end;
			}
			{   /* last best length-corrected anker (savekmers.c:5590-5609) */
				double sl = node->score;
				if (node->len_len != best_len_r->len_len) { sl /= node->len_len; sl *= best_len_r->len_len; }
				if (best_len_r->score_len < sl) { best_len_r = node; ties_len = 0; }
				else if (best_len_r->score_len == sl) {
					if (best_len_r->score_len < node->score_len) { best_len_r = node; ties_len = 0; }
					else if (best_len_r->score_len == node->score_len) { best_len_r = node; ++ties_len; }
				}
			}
			if (best_r->score < node->score) { best_r = node; ties = 0; }
			else if (best_r->score == node->score) {
				if (best_r->score_len < node->score_len) { best_r = node; ties = 0; }
				else { best_r = node; ++ties; }
			}
		}
		for (int i = 1; i <= bests[0]; ++i) { const int t = bests[i]; c->score[t] = 0; c->ext[t] = 0; c->incl[t] = 0; }
	}
	if (best->score < k && best_r->score < k) return 0;

	const int VF_start = VF[0].start, VR_start = VR[0].start;
	int headF = prune(VF, k), headR = prune(VR, k);
	if (headF < 0) best->score = 0;
	if (headR < 0) best_r->score = 0;
	bt[0] = 0; bt_r[0] = 0;
	if (g_lc) { ties = ties_len; best = best_len; best_r = best_len_r; }   /* savekmers.c:5657-5664 */

	int cs = -1, cs_r = -1, start, len, rc;
	ank_t *tmp;
	if (!best->score || !best_r->score) {
		if (best->score) {
			tmp = get_chain_templates(c, best, VF, bt, &err);
			if (err) return err;
			cs = tmp->start; start = cs; len = best->end - start; rc = 1;
		} else {
			tmp = get_chain_templates(c, best_r, VR, bt_r, &err);
			if (err) return err;
			cs_r = tmp->start; start = cs_r; len = best_r->end - start; rc = 2;
		}
	} else {
		tmp = get_chain_templates(c, best, VF, bt, &err); if (err) return err;
		cs = tmp->start;
		tmp = get_chain_templates(c, best_r, VR, bt_r, &err); if (err) return err;
		cs_r = tmp->start;
		rc = choose_chain(best, best_r, cs, cs_r, W->coverT, &start, &len);
	}
	{
		const int score = best->score < best_r->score ? best_r->score : best->score;
		if (len < W->minlen || score < k) return 0;
	}

	while (best || best_r) {
		if (ties) {   /* ankers with the same score further up the read join when they overlap enough (savekmers.c:5701-5781) */
			for (int side = 1; side <= 2; ++side) {
				if (!(rc & side)) continue;
				int *bl = side == 1 ? bt : bt_r;
				ank_t *bs = side == 1 ? best : best_r, *V = bs, *lo = side == 1 ? VF : VR;
				const int vstart = side == 1 ? VF_start : VR_start;
				while ((V = tie_anker(start < vstart ? vstart : start, V, bs))) {
					if ((double)((unsigned)V->end - (unsigned)start) < W->coverT * len) V = 0;
					else {
						for (int i = 1; i <= bl[0]; ++i) { const int t = bl[i]; c->incl[t] = 1; c->score[t] = 0; c->ext[t] = 0; }
						int *tail = bl + bl[0];   /* the new set is collected behind the current one */
						const int keep = *tail;
						*tail = 0;
						get_chain_templates(c, V, lo, tail, &err);
						if (err) return err;
						bl[0] += *tail;
						*tail = keep;
					}
				}
				for (int i = 1; i <= bl[0]; ++i) { const int t = bl[i]; c->incl[t] = 0; c->score[t] = 0; c->ext[t] = 0; }
			}
		}
		if ((rc & 1) && !mr_chain(bt, lengths, seqlen, len, W->mrc)) rc ^= 1;
		if ((rc & 2) && !mr_chain(bt_r, lengths, seqlen, len, W->mrc)) rc ^= 2;

		if (rc) {
			if (st_grow(&T, (unsigned)start, (unsigned)(start + len))) return -3;
			const int b0 = (rc & 1) ? start : seqlen - best_r->end, b1 = (rc & 1) ? start + len : seqlen - start;
			const size_t need = 28 + 8 * (size_t)((seqlen + 31) >> 5) + 4 * (size_t)nN + 4 * (size_t)(bt[0] + bt_r[0]) + hdrlen + 9;
			if (op + need > cap) return -1;
			if (rc & 1) {
				if (rc & 2) {
					for (int i = 1; i <= bt_r[0]; ++i) bt[bt[0] + i] = -bt_r[i];
					bt[0] += bt_r[0];
					best->score = -best->score;
					best_r->score = 0;
					bt_r[0] = 0;
				}
				op += emit_chain_record(out + op, fw, seqlen, fN, nN, best->score, bt + 1, bt[0], hdr, hdrlen, b0, b1);
				best->score = 0; bt[0] = 0;
			} else {
				op += emit_chain_record(out + op, rw, seqlen, rN, nN, best_r->score, bt_r + 1, bt_r[0], hdr, hdrlen, b0, b1);
				best_r->score = 0; bt_r[0] = 0;
			}
		}

		/* next chain of either strand (savekmers.c:5838-5924) */
		ties = 0; rc = 0;
		for (int side = 1; side <= 2; ++side) {
			ank_t **bp = side == 1 ? &best : &best_r, *V = side == 1 ? VF : VR;
			int *bl = side == 1 ? bt : bt_r, *head = side == 1 ? &headF : &headR, *csp = side == 1 ? &cs : &cs_r;
			if (!*bp) continue;
			int first = 1;
			for (;;) {
				ank_t *b = *bp;
				if (!first) {
					if (!(b && b->score == 0)) break;
					*bp = b = g_lc ? best_anker_len(V, head, &ties) : best_anker(V, head, &ties);
					if (!b) break;
				}
				const int ok_score = first ? b->score != 0 : k < b->score;
				first = 0;
				if (ok_score && (tmp = get_chain_templates(c, b, V, bl, &err))) {
					if (err) return err;
					*csp = tmp->start;
					const unsigned cover = T.n ? st_query(&T, 0, (unsigned)*csp, (unsigned)b->end) : 0;
					len = b->end - *csp;
					if (W->minlen <= len && cover <= W->coverT * len && W->mrs * len <= b->score) rc |= side;
					else b->score = 0;
				} else {
					if (err) return err;
					b->score = 0;
				}
			}
		}
		if (!best && !best_r) break;
		if (best && best_r) rc = choose_chain(best, best_r, cs, cs_r, W->coverT, &start, &len);
		else if (best) { rc = 1; start = cs; len = best->end - start; }
		else { rc = 2; start = cs_r; len = best_r->end - start; }
	}
	return (int64_t)op;
}

/* Whole stage 2 in chain mode: stage-1 stream in, stage-2 stream out (including the terminator).
 * Returns bytes written, or a negative error (see chain_read). */
int64_t orc_chain_stream(const orc_db *db, const orc_params *p, const uint8_t *in, size_t in_bytes, int minlen,
                         double mrs, double coverT, double mrc, uint8_t *out, size_t cap, orc_stats *st) {
	const size_t D = (size_t)db->DB_size + 1;
	chain_ws W; memset(&W, 0, sizeof(W));
	if (!db->lengths) return -5;
	W.c.db = db; W.c.p = p; W.c.k = (int)db->kmersize;
	W.c.score = calloc(D, sizeof(int)); W.c.ext = calloc(D, sizeof(int)); W.c.incl = calloc(D, 1);
	W.bt = malloc(sizeof(int) * (2 * D + 4)); W.bt_r = malloc(sizeof(int) * (2 * D + 4));
	W.minlen = minlen; W.mrs = mrs; W.coverT = coverT; W.mrc = mrc;
	size_t ip = 0, op = 0;
	int32_t nreads = 0;
	int64_t ret = 0;
	while (ip + 16 <= in_bytes) {
		int32_t h[4]; memcpy(h, in + ip, 16);
		if (h[0] < 0) break;
		ip += 16;
		const int seqlen = h[0], words = h[1], nN = h[2], hdrlen = abs(h[3]);
		if ((size_t)words + 4 > W.wcap) {
			W.wcap = 2 * (size_t)words + 4;
			for (int i = 0; i < 2; ++i) W.w[i] = realloc(W.w[i], 8 * W.wcap);
		}
		if ((size_t)nN + 2 > W.ncap) {
			W.ncap = 2 * (size_t)nN + 2;
			for (int i = 0; i < 2; ++i) W.N[i] = realloc(W.N[i], 4 * W.ncap);
		}
		if ((size_t)seqlen + 4 > W.acap) {
			W.acap = 2 * (size_t)seqlen + 4;
			W.VF = realloc(W.VF, sizeof(ank_t) * W.acap); W.VR = realloc(W.VR, sizeof(ank_t) * W.acap);
			memset(W.VF, 0, sizeof(ank_t) * W.acap); memset(W.VR, 0, sizeof(ank_t) * W.acap);
		}
		memcpy(W.w[0], in + ip, 8 * (size_t)words); ip += 8 * (size_t)words;
		memcpy(W.N[0], in + ip, 4 * (size_t)nN); ip += 4 * (size_t)nN;
		const uint8_t *hdr = in + ip; ip += hdrlen;
		for (int i = 0; i < 4; ++i) W.w[0][words + i < (int)W.wcap ? words + i : words] = 0;
		orc_revcomp(W.w[0], seqlen, W.N[0], nN, W.w[1], W.N[1]);
		for (int i = 0; i < 4; ++i) W.w[1][words + i < (int)W.wcap ? words + i : words] = 0;
		++nreads;
		if (st) st->reads++, st->read_words += words;
		if (seqlen < W.c.k) continue;
		int64_t w = chain_read(&W, seqlen, nN, hdr, hdrlen, out + op, cap - op - 4, st);
		if (w < 0) { ret = w; goto done; }
		if (w && st) st->mapped++;
		op += (size_t)w;
	}
	if (op + 4 > cap) { ret = -1; goto done; }
	nreads = -nreads; memcpy(out + op, &nreads, 4); op += 4;
	ret = (int64_t)op;
done:
	free(W.c.score); free(W.c.ext); free(W.c.incl); free(W.bt); free(W.bt_r); free(W.VF); free(W.VR);
	for (int i = 0; i < 2; ++i) { free(W.w[i]); free(W.N[i]); }
	return ret;
}
