/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of KMA 1.5.1's alignment pass (stage 3, first half):
 * stage-2 records in, frag_raw records + ConClave score arrays out.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Parity pin: tests/test_oracle_align.py compares (a) every per-candidate AlnScore and (b) the whole frag_raw
 * stream + alignment_scores/uniq_alignment_scores with the UNMODIFIED reference driven by oracle/ref_harness.c
 * (oracle/_ref/ref_aln), on fresh seeded data and on the committed golden streams.
 *
 * Restated (not copied) from:
 *   hashmapcci.c:95-199, 409-505  HashMapCCI get/get_bound/getDubPos/getNextDubPos/load  -> tindex_* (sorted
 *        (k-mer, position) arrays: same answers and the same ascending enumeration order)
 *   align.c:509-748   KMA_score (seed scan, stitching)          -> kma_score
 *   align.c:53-212    leadTailAln / trailTailAln                -> lead_tail / trail_tail
 *   align.c:750-770   preseed                                   -> preseed_hit
 *   align.c:993-1176  anker_rc_comp                             -> pick_strand
 *   chain.c:79-260    chainSeeds                                -> chain_mems
 *   nw.c:642-890      NW_score                                  -> nw_full
 *   nw.c:892-1188     NW_band_score                             -> nw_band
 *   alnfrags.c:1052-1218 alnFragsSE, :2150-2294 alnFrags_threaded (SE) -> orc_align_stream
 *   updatescores.c:203-298 update_Scores (frag_raw record)      -> reduce_and_emit
 *   alnfrags.c:1596-1972 alnFragsPenaltyPE                       -> align_pe
 *   updatescores.c:300-488 update_Scores_se / update_Scores_pe    -> emit_se / emit_pe
 * Out of scope here (asserted): circular templates (t_len < 0), chain-mode q-bounds, strand-undecided pairs.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "orc.h"

#define IMIN(a, b) ((a) < (b) ? (a) : (b))
#define IMAX(a, b) ((a) < (b) ? (b) : (a))

/* ------------------------------------------------------------------ template side ---- */

int orc_db_load_seq(orc_db *db, const char *prefix) {
	char path[4096];
	if (db->seq) return 0;
	if (!db->lengths) return -1;
	db->seq_off = calloc(db->DB_size + 1, sizeof(int64_t));
	for (int t = 1; t < db->DB_size; ++t) db->seq_off[t + 1] = db->seq_off[t] + ((db->lengths[t] >> 5) + 1);
	size_t words = (size_t)db->seq_off[db->DB_size];
	db->seq = calloc(words + 2, 8);
	snprintf(path, sizeof(path), "%s.seq.b", prefix);
	FILE *f = fopen(path, "rb");
	if (!f) return -1;
	size_t got = fread(db->seq, 8, words, f);
	fclose(f);
	return got == words ? 0 : -1;
}

static inline int nuc_at(const uint64_t *seq, int pos) { return (int)((seq[pos >> 5] << ((pos & 31) << 1)) >> 62); }

static inline uint64_t kmer_at(const uint64_t *seq, int pos, int k) {
	int w = pos >> 5, b = (pos & 31) << 1, sh = 64 - 2 * k;
	uint64_t x = seq[w] << b;
	if (b > sh) x |= seq[w + 1] >> (64 - b);
	return x >> sh;
}

/* per-template position index: (k-mer, 1-based position) sorted by k-mer then position */
typedef struct { uint64_t *kmer; int32_t *pos; int n, len, k; const uint64_t *seq; } tindex;

typedef struct { uint64_t k; int32_t p; } kp_t;
static int kp_cmp(const void *a, const void *b) {
	const kp_t *x = a, *y = b;
	if (x->k != y->k) return x->k < y->k ? -1 : 1;
	return x->p - y->p;
}

static tindex *tindex_build(const uint64_t *seq, int len, int k) {
	tindex *ix = calloc(1, sizeof(tindex));
	int n = len - k + 1;
	if (n < 0) n = 0;
	kp_t *v = malloc(sizeof(kp_t) * (n + 1));
	int m = 0;
	for (int i = 0; i < n; ++i) {
		uint64_t key = kmer_at(seq, i, k);
		if (key == 0) continue; /* hashMapCCI_add skips poly-A (hashmapcci.c:414) */
		v[m].k = key; v[m].p = i + 1; ++m;
	}
	qsort(v, m, sizeof(kp_t), kp_cmp);
	ix->kmer = malloc(8 * (m + 1)); ix->pos = malloc(4 * (m + 1));
	for (int i = 0; i < m; ++i) { ix->kmer[i] = v[i].k; ix->pos[i] = v[i].p; }
	free(v);
	ix->n = m; ix->len = len; ix->k = k; ix->seq = seq;
	return ix;
}

static void tindex_free(tindex *ix) { if (ix) { free(ix->kmer); free(ix->pos); free(ix); } }

/* first slot holding key, or -1; *cnt = occurrences */
static int tindex_find(const tindex *ix, uint64_t key, int *cnt) {
	int lo = 0, hi = ix->n;
	while (lo < hi) { int mid = (lo + hi) >> 1; if (ix->kmer[mid] < key) lo = mid + 1; else hi = mid; }
	if (lo == ix->n || ix->kmer[lo] != key) { *cnt = 0; return -1; }
	int e = lo;
	while (e < ix->n && ix->kmer[e] == key) ++e;
	*cnt = e - lo;
	return lo;
}

/* hashMapCCI_get: 0 = absent, +pos = unique k-mer, -pos(first) = repeated k-mer */
static int tindex_get(const tindex *ix, uint64_t key, int *slot, int *cnt) {
	*slot = tindex_find(ix, key, cnt);
	if (*slot < 0) return 0;
	return *cnt == 1 ? ix->pos[*slot] : -ix->pos[*slot];
}

/* hashMapCCI_get_bound used as a boolean by preseed: any occurrence with min < pos < max */
static int tindex_any_bound(const tindex *ix, uint64_t key, int min, int max) {
	int cnt, s = tindex_find(ix, key, &cnt);
	for (int i = 0; i < cnt; ++i) if (min < ix->pos[s + i] && ix->pos[s + i] < max) return 1;
	return 0;
}

/* ------------------------------------------------------------------ MEM list --------- */

typedef struct {
	int cap, len;
	int *tStart, *tEnd, *qStart, *qEnd, *weight, *score, *next;
} mems_t;

static void mems_reserve(mems_t *p, int need) {
	if (need < p->cap) return;
	int c = p->cap ? p->cap : 1024;
	while (c <= need) c <<= 1;
	p->tStart = realloc(p->tStart, 4 * (size_t)c); p->tEnd = realloc(p->tEnd, 4 * (size_t)c);
	p->qStart = realloc(p->qStart, 4 * (size_t)c); p->qEnd = realloc(p->qEnd, 4 * (size_t)c);
	p->weight = realloc(p->weight, 4 * (size_t)c); p->score = realloc(p->score, 4 * (size_t)c);
	p->next = realloc(p->next, 4 * (size_t)c);
	p->cap = c;
}

/* extend an exact seed at (query i, template 1-based value) to a maximal exact match.
 * fwd_lim = exclusive query limit of the forward extension. Returns query end; *tEnd1 = value + 1. */
static void mem_from_seed(const tindex *ix, const uint8_t *q, int i, int value, int k, int fwd_lim,
                          int *qs, int *ts, int *qe, int *te) {
	int prev = value - 2, j;
	for (j = i - 1; 0 <= j && 0 <= prev && q[j] == nuc_at(ix->seq, prev); --j) --prev;
	*qs = j + 1; *ts = prev + 2;
	value += k - 1;
	int l = i + k;
	while (l < fwd_lim && value < ix->len && q[l] == nuc_at(ix->seq, value)) { ++l; ++value; }
	*qe = l; *te = value + 1;
}

/* ------------------------------------------------------------------ chaining --------- */

static int tail_mm(const orc_params *p, int Ms, int k) { /* mismatch estimate of a gap of Ms bases */
	int MMs;
	if (Ms == 2) { MMs = 2; Ms = 0; }
	else {
		MMs = Ms / k + (Ms % k ? 1 : 0); MMs = IMAX(2, MMs);
		Ms = IMIN(Ms - MMs, k); Ms = IMIN(Ms, MMs);
	}
	return Ms * p->M + MMs * p->MM;
}

static int chain_mems(const orc_params *p, mems_t *pt, int q_len, int t_len, int k, unsigned *mapQ) {
	const int W1 = p->W1, U = p->U, M = p->M;
	int n = pt->len, bestPos = n - 1, bestScore = 0, secondScore = 0;
	mems_reserve(pt, n + 1);
	pt->score[n] = 0; pt->next[n] = 0;
	for (int i = n - 1; i >= 0; --i) {
		int weight = pt->weight[i] * M, tEnd = pt->tEnd[i], qEnd = pt->qEnd[i], gap, Ms, score;
		pt->next[i] = 0;
		gap = IMIN(t_len - tEnd, q_len - qEnd);
		Ms = gap;
		if (--gap) gap = gap * U + W1; else gap = W1;       /* (the reference's third branch is unreachable) */
		Ms = tail_mm(p, Ms, k);
		score = weight + (Ms < gap ? gap : Ms);
		int lim = IMIN(n, i + 128);
		for (int j = i + 1; j < lim; ++j) {
			if (qEnd < pt->qStart[j]) {
				if (tEnd < pt->tStart[j]) {
					int tGap = pt->tStart[j] - tEnd, qGap = pt->qStart[j] - qEnd;
					if ((gap = abs(tGap - qGap))) gap = (gap - 1) * U + W1;
					gap += weight + pt->score[j] + tail_mm(p, IMIN(tGap, qGap), k);
					if (score <= gap) { score = gap; pt->next[i] = j; }
				} else if (k <= pt->tEnd[j] - tEnd) {
					if ((gap = pt->qStart[j] - qEnd)) gap = (gap - 1) * U + W1;
					gap += weight + pt->score[j] - (pt->tStart[j] - tEnd) * M;
					if (score < gap) { score = gap; pt->next[i] = j; }
				}
			} else if (k <= pt->qEnd[j] - qEnd) {
				int tStart = pt->tStart[j] + qEnd - pt->qStart[j];
				if (tEnd < tStart) {
					if ((gap = tStart - tEnd)) gap = (gap - 1) * U + W1;
					gap += weight + pt->score[j] - (tStart - tEnd) * M;
					if (score < gap) { score = gap; pt->next[i] = j; }
				}
			}
		}
		if (pt->next[i]) pt->weight[i] += pt->weight[pt->next[i]] - k + 1;
		else pt->weight[i] -= k - 1;
		pt->score[i] = score;
		gap = IMIN(pt->tStart[i], pt->qStart[i]);
		Ms = gap;
		if (0 < --gap) gap = gap * U + W1; else if (gap == 0) gap = W1; else gap = 0;
		Ms = tail_mm(p, Ms, k);
		score += Ms < gap ? gap : Ms;
		if (bestScore <= score) {
			if (pt->next[i] != bestPos) secondScore = bestScore;
			bestScore = score; bestPos = i;
		} else if (secondScore <= score && pt->next[i] != bestPos) secondScore = bestScore;
	}
	*mapQ = 0 < bestScore ? (unsigned)ceil(40 * (1 - 1.0 * secondScore / bestScore) * (pt->weight[bestPos] / 10.0 < 1 ? pt->weight[bestPos] / 10.0 : 1) * log(bestScore)) : 0;
	pt->score[bestPos] = bestScore;
	return bestPos;
}

/* ------------------------------------------------------------------ Needleman-Wunsch - */

typedef struct { int score, len, pos, match, tGaps, qGaps; } aln_t;
/* aligned strings of one NW call: template / match / query rows (0-3 bases, 4 = N, 5 = gap; '|' or '_') */
typedef struct { uint8_t *t, *s, *q; } astr;

static int64_t g_band_calls = 0, g_full_calls = 0;
int64_t orc_nw_band_calls(void) { return g_band_calls; }
int64_t orc_nw_full_calls(void) { return g_full_calls; }

typedef struct { int *D[2], *P[2]; size_t cols; uint8_t *E; size_t esize; int64_t cells; } nw_ws;

static void ws_reserve(nw_ws *w, size_t cols, size_t ebytes) {
	if (w->cols < cols) {
		w->cols = cols * 2;
		for (int i = 0; i < 2; ++i) { free(w->D[i]); free(w->P[i]); w->D[i] = calloc(w->cols, 4); w->P[i] = calloc(w->cols, 4); }
	}
	if (w->esize < ebytes) { w->esize = ebytes * 2; free(w->E); w->E = calloc(w->esize, 1); }
}

/* one DP cell; returns the traceback byte (nw.c:166-212): 1 diag, 2/3 Q open/extend, 4/5 P open/extend,
 * +16 the Q run may open here, +32 the P run may open here */
static inline uint8_t nw_cell(int Dright, int Qright, int Ddown, int Pdown, int Ddiag, int sub, int W1, int U,
                              int *Dout, int *Qout, int *Pout) {
	int Q = Dright + W1, P = Ddown + W1, D, x;
	uint8_t e, fl = 0;
	if (Q < P) { D = P; e = 4; } else { D = Q; e = 2; }
	x = Qright + U;
	if (Q < x) { Q = x; if (D <= x) { D = x; e = 3; } } else fl |= 16;
	x = Pdown + U;
	if (P < x) { P = x; if (D <= x) { D = x; e = 5; } } else fl |= 32;
	x = Ddiag + sub;
	if (D <= x) { D = x; e = 1; }
	*Dout = D; *Qout = Q; *Pout = P;
	return fl | e;
}

/* one side empty (nw.c:49-85); with `so` the all-gap rows are written too */
static aln_t nw_trivial(const orc_params *p, int t_len, int q_len, astr *so, const uint64_t *tseq, int t_s, const uint8_t *q) {
	aln_t s = {0, 0, 0, 0, 0, 0};
	if (t_len == q_len) return s;
	if (t_len == 0) {
		s.len = q_len; s.tGaps = q_len; s.score = p->W1 + (q_len - 1) * p->U;
		if (so) { memset(so->s, '_', q_len); memset(so->t, 5, q_len); memcpy(so->q, q, q_len); }
	} else {
		s.len = t_len; s.qGaps = t_len; s.score = p->W1 + (t_len - 1) * p->U;
		if (so) { memset(so->s, '_', t_len); memset(so->q, 5, t_len); for (int m = 0; m < t_len; ++m) so->t[m] = (uint8_t)nuc_at(tseq, t_s + m); }
	}
	return s;
}

/* walk the traceback bytes from (m, n); `stride` = bytes per row, `dn` = column step of a vertical move
 * (0 for the full matrix, -1 for the band whose rows are skewed). With `so` the aligned rows are written as
 * NW / NW_band do (nw.c:250-305, 575-637): tseq/t0 = template and the position of row 0, q/qp = query and the
 * query position of the start cell. */
static void nw_walk(const uint8_t *E, size_t stride, int m, int n, int dn, aln_t *s, astr *so, const uint64_t *tseq, int t0,
                    const uint8_t *q, int qp) {
	const uint8_t *row = E + (size_t)m * stride;
	int tp = t0 + m;
	s->len = s->match = s->tGaps = s->qGaps = 0;
	while (row[n] != 0) {
		int e = row[n] & 7;
		if (e == 1) {
			if (so) { so->t[s->len] = (uint8_t)nuc_at(tseq, tp); so->q[s->len] = q[qp]; so->s[s->len] = so->t[s->len] == so->q[s->len] ? '|' : '_'; }
			++s->match; row += stride; n += 1 + dn; ++tp; ++qp;
		} else if (e >= 4) {
			while (!(row[n] >> 4)) {
				if (so) { so->t[s->len] = (uint8_t)nuc_at(tseq, tp); so->q[s->len] = 5; so->s[s->len] = '_'; }
				row += stride; n += dn; ++s->len; ++s->qGaps; ++tp;
			}
			if (so) { so->t[s->len] = (uint8_t)nuc_at(tseq, tp); so->q[s->len] = 5; so->s[s->len] = '_'; }
			++s->qGaps; row += stride; n += dn; ++tp;
		} else {
			while (!(row[n] >> 3)) {
				if (so) { so->t[s->len] = 5; so->q[s->len] = q[qp]; so->s[s->len] = '_'; }
				++n; ++s->len; ++s->tGaps; ++qp;
			}
			if (so) { so->t[s->len] = 5; so->q[s->len] = q[qp]; so->s[s->len] = '_'; }
			++s->tGaps; ++n; ++qp;
		}
		++s->len;
	}
}

static aln_t nw_full(const orc_params *p, nw_ws *w, const uint64_t *tseq, const uint8_t *query, int k,
                     int t_s, int t_e, int q_s, int q_e, astr *so) {
	const int W1 = p->W1, U = p->U, t_len = t_e - t_s, q_len = q_e - q_s;
	const uint8_t *q = query + q_s;
	if (t_len == 0 || q_len == 0) return nw_trivial(p, t_len, q_len, so, tseq, t_s, q);
	const size_t stride = (size_t)q_len + 1;
	++g_full_calls;
	ws_reserve(w, (size_t)q_len + 2, ((size_t)q_len + 2) * ((size_t)t_len + 2));
	w->cells += (int64_t)t_len * q_len;
	const int NEG = (t_len + q_len) * (p->MM + U + W1);
	int *Dp = w->D[1], *Pp = w->P[1], *Dc = w->D[0], *Pc = w->P[0];
	uint8_t *E = w->E, *last = E + (size_t)t_len * stride;
	aln_t s = {NEG, 0, 0, 0, 0, 0};
	int best_m = 0, best_n = 0;
	/* boundary row (all template consumed) and boundary column (all query consumed) */
	if (k == 2) { for (int n = 0; n <= q_len; ++n) { Dp[n] = 0; Pp[n] = NEG; last[n] = 0; } }
	else {
		for (int n = 0; n < q_len; ++n) { Dp[n] = W1 + (q_len - 1 - n) * U; Pp[n] = NEG; last[n] = 3; }
		last[q_len - 1] = 18; last[q_len] = 0; Dp[q_len] = 0; Pp[q_len] = 0;
	}
	for (int m = 0; m < t_len; ++m) E[(size_t)m * stride + q_len] = 0 < k ? 0 : 5;
	if (!(0 < k)) E[(size_t)(t_len - 1) * stride + q_len] = 36;

	for (int m = t_len - 1; m >= 0; --m) {
		uint8_t *row = E + (size_t)m * stride;
		const int tn = nuc_at(tseq, t_s + m);
		int Qr = NEG;
		Dc[q_len] = 0 < k ? 0 : W1 + (t_len - 1 - m) * U;
		for (int n = q_len - 1; n >= 0; --n)
			row[n] = nw_cell(Dc[n + 1], Qr, Dp[n], Pp[n], Dp[n + 1], p->d[tn * 5 + q[n]], W1, U, &Dc[n], &Qr, &Pc[n]);
		if (k < 0 && s.score < Dc[0]) { s.score = Dc[0]; best_m = m; }
		int *t1 = Dc; Dc = Dp; Dp = t1; t1 = Pc; Pc = Pp; Pp = t1;
	}
	if (k < 0) {
		best_n = 0;
		if (k == -2) for (int n = 0; n < q_len; ++n) if (s.score <= Dp[n]) { s.score = Dp[n]; best_m = 0; best_n = n; }
	} else { s.score = Dp[0]; best_m = 0; best_n = 0; }
	int sc = s.score;
	nw_walk(E, stride, best_m, best_n, 0, &s, so, tseq, t_s, q, best_n);
	s.score = sc; s.pos = 0;
	return s;
}

static aln_t nw_band(const orc_params *p, nw_ws *w, const uint64_t *tseq, const uint8_t *query, int k,
                     int t_s, int t_e, int q_s, int q_e, int band, astr *so) {
	const int W1 = p->W1, U = p->U, t_len = t_e - t_s, q_len = q_e - q_s;
	const uint8_t *q = query + q_s;
	if (t_len == 0 || q_len == 0) return nw_trivial(p, t_len, q_len, so, tseq, t_s, q);
	if (band & 1) ++band;
	++g_band_calls;
	const int half = band >> 1, bq = band + 1;
	const size_t stride = (size_t)bq + 1;
	ws_reserve(w, 2 * (size_t)band + 4, ((size_t)band + 3) * ((size_t)t_len + 2));
	w->cells += (int64_t)t_len * bq;
	const int NEG = (t_len + q_len) * (p->MM + U + W1);
	int *Dp = w->D[1], *Pp = w->P[1], *Dc = w->D[0], *Pc = w->P[0];
	uint8_t *E = w->E, *last = E + (size_t)t_len * stride;
	aln_t s = {NEG, 0, 0, 0, 0, 0};
	int c = (t_len + q_len) >> 1, sn = q_len - 1 - (c - half), en = 0, best_m = 0, best_n = 0, n;
	if (k != 2) {
		for (n = sn - 1; n >= 0; --n) { Dp[n] = W1 + (sn - n - 1) * U; Pp[n] = NEG; last[n] = 3; }
		last[sn - 1] = 18; last[sn] = 0; Dp[sn] = 0; Pp[sn] = 0;
	} else for (n = sn; n >= 0; --n) { Dp[n] = 0; Pp[n] = NEG; last[n] = 0; }

	for (int m = t_len - 1; m >= 0; --m, --c) {
		uint8_t *row = E + (size_t)m * stride;
		int sq = c + half, eq = c - half, qp, Qr = NEG, Dn, Qn, Pn;
		if (eq < 0) { eq = 0; ++en; } else en = 0;
		if (sq < q_len - 1) { sn = bq - 1; Dc[bq] = NEG; row[bq] = 37; }
		else { sq = q_len - 1; sn = en + (q_len - eq); Dc[sn] = 0 < k ? 0 : W1 + (t_len - 1 - m) * U; row[sn] = 0 < k ? 0 : 37; --sn; }
		const int tn = nuc_at(tseq, t_s + m);
		for (n = sn, qp = sq; n > en; --qp, --n)
			row[n] = nw_cell(Dc[n + 1], Qr, Dp[n - 1], Pp[n - 1], Dp[n], p->d[tn * 5 + q[qp]], W1, U, &Dc[n], &Qr, &Pc[n]);
		/* left edge of the band: no vertical (P) move into this cell (nw.c:1076-1102) */
		{
			uint8_t e, fl = 0;
			Qn = Dc[n + 1] + W1;
			if (Qn < Qr + U) { Qn = Qr + U; e = 3; } else { e = 2; fl = 16; }
			Pc[n] = NEG;
			Dn = Dp[n] + p->d[tn * 5 + q[qp]];
			if (Qn <= Dn) row[n] = fl | 1; else { Dn = Qn; row[n] = fl | e; }
			Dc[n] = Dn; (void)Pn;
		}
		if (eq == 0 && k < 0 && s.score < Dc[n]) { s.score = Dc[n]; best_m = m; best_n = n; }
		int *t1 = Dc; Dc = Dp; Dp = t1; t1 = Pc; Pc = Pp; Pp = t1;
	}
	if (best_m == 0) { best_n = en; s.score = Dp[en]; }
	int qstart = 0;   /* the reference's q_pos: 0 unless the free-query-prefix scan moves it (nw.c:561-574) */
	if (k == -2) for (n = en; n < bq; ++n) if (s.score <= Dp[n]) { s.score = Dp[n]; best_m = 0; best_n = n; qstart = n - en; }
	int sc = s.score;
	nw_walk(E, stride, best_m, best_n, -1, &s, so, tseq, t_s, q, qstart);
	s.score = sc; s.pos = 0;
	return s;
}

/* stand-alone NW entry for the per-call parity tests (band == 0: NW_score, else NW_band_score) */
void orc_nw(const orc_params *p, const uint64_t *tseq, const uint8_t *query, int k, int t_s, int t_e, int q_s, int q_e,
            int band, int *out6) {
	nw_ws w; memset(&w, 0, sizeof(w));
	aln_t a = band ? nw_band(p, &w, tseq, query, k, t_s, t_e, q_s, q_e, band, 0) : nw_full(p, &w, tseq, query, k, t_s, t_e, q_s, q_e, 0);
	out6[0] = a.score; out6[1] = a.len; out6[2] = a.pos; out6[3] = a.match; out6[4] = a.tGaps; out6[5] = a.qGaps;
	for (int i = 0; i < 2; ++i) { free(w.D[i]); free(w.P[i]); } free(w.E);
}

/* ------------------------------------------------------------------ seed-and-extend -- */

#define BANDW 64

static aln_t nw_auto(const orc_params *p, nw_ws *w, const uint64_t *tseq, const uint8_t *q, int k,
                     int t_s, int t_e, int q_s, int q_e) {
	int band = abs((t_e - t_s) - (q_e - q_s)) + BANDW;
	if (q_e - q_s <= band || t_e - t_s <= band) return nw_full(p, w, tseq, q, k, t_s, t_e, q_s, q_e, 0);
	return nw_band(p, w, tseq, q, k, t_s, t_e, q_s, q_e, band, 0);
}

static aln_t lead_tail(const orc_params *p, nw_ws *w, const uint64_t *tseq, const uint8_t *q, int t_e, int q_e) {
	aln_t s = {0, 0, t_e, 0, 0, 0};
	if (!q_e) return s;
	int t_s = 0, q_s = 0;
	if ((q_e << 1) < t_e || (q_e + BANDW) < t_e) t_s = t_e - (q_e + IMIN(q_e, BANDW));
	else if ((t_e << 1) < q_e || (t_e + BANDW) < q_e) q_s = q_e - (t_e + IMIN(t_e, BANDW));
	if (t_e - t_s > 0 && q_e - q_s > 0) {
		aln_t a = nw_auto(p, w, tseq, q, -1 - (t_s == 0), t_s, t_e, q_s, q_e);
		s.pos -= a.len - a.tGaps;
		s.score = a.score; s.len = a.len; s.match = a.match; s.tGaps = a.tGaps; s.qGaps = a.qGaps;
	}
	return s;
}

static void trail_tail(const orc_params *p, nw_ws *w, aln_t *s, const uint64_t *tseq, const uint8_t *q,
                       int t_s, int t_len, int q_s, int q_len) {
	int q_e = q_len, t_e = t_len;
	if (((q_len - q_s) << 1) < (t_len - t_s) || (q_len - q_s + BANDW) < (t_len - t_s)) {
		t_e = q_len - q_s; t_e = t_s + (t_e + IMIN(t_e, BANDW));
	} else if (((t_len - t_s) << 1) < (q_len - q_s) || (t_len - t_s + BANDW) < (q_len - q_s)) {
		q_e = t_len - t_s; q_e = q_s + (q_e + IMIN(q_e, BANDW));
	}
	if (t_e - t_s > 0 && q_e - q_s > 0) {
		aln_t a = nw_auto(p, w, tseq, q, 1 + (t_e == t_len), t_s, t_e, q_s, q_e);
		s->score += a.score; s->len += a.len; s->match += a.match; s->tGaps += a.tGaps; s->qGaps += a.qGaps;
	}
}

static aln_t aln_zero(void) { aln_t s = {0, 1, 0, 0, 0, 0}; return s; }

/* KMA_score: MEMs (found here unless pt->len != 0 on entry), chain, stitch with NW. qN[] = N positions with a
 * q_len sentinel at qN[nN] (the caller's N[0]++ convention). */
static aln_t kma_score(const orc_params *p, nw_ws *w, const tindex *ix, const uint8_t *q, int q_len, int q_start, int q_end,
                       const uint64_t *qcomp, const int32_t *qN, int nN1, int mq, mems_t *pt) {
	const int k = ix->k, t_len = ix->len, U = p->U, M = p->M;
	int n = pt->len;
	if (!n) {
		/* q_start / q_end: the query bounds a chain-mode record carries (alnfrags.c:1091-1099). Only the first stretch
		 * starts at q_start and only the last one stops at q_end (align.c:535-541, 638). */
		int j = q_start;
		for (int seg = 0; seg < nN1; ++seg) {
			int end = (seg != nN1 - 1 ? qN[seg] : q_end) - k + 1;
			while (j < end) {
				int slot, cnt, value = tindex_get(ix, kmer_at(qcomp, j, k), &slot, &cnt);
				if (value == 0) { ++j; continue; }
				mems_reserve(pt, n + cnt + 1);
				if (0 < value) {
					mem_from_seed(ix, q, j, value, k, end + k - 1, &pt->qStart[n], &pt->tStart[n], &pt->qEnd[n], &pt->tEnd[n]);
					pt->weight[n] = pt->qEnd[n] - pt->qStart[n];
					j = pt->qEnd[n];
					++n;
				} else {
					int bias = j;
					for (int c = 0; c < cnt; ++c) {   /* every occurrence, ascending template position */
						mem_from_seed(ix, q, j, ix->pos[slot + c], k, end + k - 1, &pt->qStart[n], &pt->tStart[n], &pt->qEnd[n], &pt->tEnd[n]);
						pt->weight[n] = pt->qEnd[n] - pt->qStart[n];
						if (bias < pt->qEnd[n]) bias = pt->qEnd[n];
						++n;
					}
					j = bias + 1;
				}
			}
			j = qN[seg] + 1;
		}
	}
	pt->len = n;
	if (!n) return aln_zero();
	unsigned mapQ = 0;
	int start = chain_mems(p, pt, q_len, t_len, k, &mapQ);
	if (mapQ < (unsigned)mq || pt->score[start] < k) { pt->len = 0; return aln_zero(); }

	aln_t s = lead_tail(p, w, ix->seq, q, pt->tStart[start] - 1, pt->qStart[start]);
	for (;;) {
		int len = pt->qEnd[start] - pt->qStart[start];
		s.len += len; s.match += len;
		for (int i = pt->qStart[start]; i < pt->qEnd[start]; ++i) s.score += p->d[q[i] * 5 + q[i]];
		if (!pt->next[start]) break;
		int q_s = pt->qEnd[start], t_s = pt->tEnd[start] - 1, t_e, t_l, q_e;
		start = pt->next[start];
		if (pt->qStart[start] < q_s) { pt->tStart[start] += q_s - pt->qStart[start]; pt->qStart[start] = q_s; }
		t_e = pt->tStart[start] - 1;
		if (t_e < t_s) {
			if (t_s <= pt->tEnd[start]) { pt->qStart[start] += t_s - t_e; t_e = t_s; t_l = 0; }
			else t_l = t_len - t_s + t_e;   /* circular joining: never produced by the linear chainer */
		} else t_l = t_e - t_s;
		q_e = pt->qStart[start];
		if (abs(t_l - q_e + q_s) * U > q_len * M || t_l > q_len || q_e - q_s > (q_len >> 1)) { int keep = s.pos; pt->len = 0; s = aln_zero(); s.pos = keep; return s; }
		if (t_l > 0 || q_e - q_s > 0) {
			aln_t a;
			int band = abs(t_l - q_e + q_s) + BANDW;
			if (q_e - q_s <= band || t_l <= band) a = nw_full(p, w, ix->seq, q, 0, t_s, t_e, q_s, q_e, 0);
			else a = nw_band(p, w, ix->seq, q, 0, t_s, t_e, q_s, q_e, band, 0);
			s.score += a.score; s.len += a.len; s.match += a.match; s.tGaps += a.tGaps; s.qGaps += a.qGaps;
		}
	}
	trail_tail(p, w, &s, ix->seq, q, pt->tEnd[start] - 1, t_len, pt->qEnd[start], q_len);
	pt->len = 0;
	return s;
}

/* preseed (align.c:750): does any k-spaced k-mer of the byte read occur in the template? Bytes past the end of
 * the read are read as 0 (the reference reads whatever its buffer holds there). */
static int preseed_hit(const tindex *ix, const uint8_t *q, int q_len, int lim) {
	for (int i = 0; i < lim; i += ix->k) {
		uint64_t key = 0;
		for (int b = 0; b < ix->k; ++b) key = (b ? key << 2 : 0) | (i + b < q_len ? q[i + b] : 0);
		if (tindex_any_bound(ix, key, 0, ix->len)) return 1;
	}
	return 0;
}

/* anker_rc_comp: MEMs of both strands; the strand with the larger MEM mass wins (ties: forward) and its MEMs are
 * left in pt. Returns +score (forward), -score (reverse) or 0. */
static int pick_strand(const orc_params *p, const tindex *ix, const uint8_t *qf, const uint8_t *qr,
                       const uint64_t *cf, const uint64_t *cr, const int32_t *Nf, const int32_t *Nr, int nN1,
                       int q_len, int q_start, int q_end, int one2one, mems_t *pt) {
	const int k = ix->k, t_len = ix->len;
	int score_f = 0, best = 0, tot = 0, cnt_strand[2] = {0, 0}, sc[2] = {0, 0};
	pt->len = 0;
	for (int rc = 0; rc < 2; ++rc) {
		const uint8_t *q = rc ? qr : qf;
		const uint64_t *comp = rc ? cr : cf;
		const int32_t *Ns = rc ? Nr : Nf;
		/* query bounds (align.c:1031-1041): mirrored for the reverse strand; preseed only without a lower bound */
		const int qs = rc ? q_len - q_end : q_start, qe = rc ? q_len - q_start : q_end;
		int i = (rc || qs) ? qs : (preseed_hit(ix, q, q_len, qe) ? 0 : q_len), seg = 0, s = 0, mc = 0;
		while (i < qe) {
			int end = Ns[seg++] - k + 1;
			while (i < end) {
				int slot, cnt, value = tindex_get(ix, kmer_at(comp, i, k), &slot, &cnt);
				if (value == 0) { ++i; continue; }
				mems_reserve(pt, tot + cnt + 1);
				if (0 < value) {
					mem_from_seed(ix, q, i, value, k, end, &pt->qStart[tot], &pt->tStart[tot], &pt->qEnd[tot], &pt->tEnd[tot]);
					pt->weight[tot] = pt->tEnd[tot] - pt->tStart[tot];
					s += pt->qEnd[tot] - pt->qStart[tot];
					i = pt->qEnd[tot] + 1;
					++tot; ++mc;
				} else {
					int bias = i;
					s += k;
					for (int c = 0; c < cnt; ++c) {
						mem_from_seed(ix, q, i, ix->pos[slot + c], k, end, &pt->qStart[tot], &pt->tStart[tot], &pt->qEnd[tot], &pt->tEnd[tot]);
						pt->weight[tot] = pt->qEnd[tot] - pt->qStart[tot];
						if (bias < pt->qEnd[tot]) bias = pt->qEnd[tot];
						++tot; ++mc;
					}
					s += bias - i;
					i = bias + 1;
				}
			}
			i = end + k;
		}
		sc[rc] = s; cnt_strand[rc] = mc;
		if (best < s) best = s;
	}
	score_f = sc[0];
	if (one2one && best < k && best * k < (q_len - k - best)) { pt->len = 0; return 0; }
	if (best == score_f) { pt->len = cnt_strand[0]; return best; }
	int off = cnt_strand[0], mc = cnt_strand[1];
	if (off) {
		memmove(pt->tStart, pt->tStart + off, 4 * (size_t)mc); memmove(pt->tEnd, pt->tEnd + off, 4 * (size_t)mc);
		memmove(pt->qStart, pt->qStart + off, 4 * (size_t)mc); memmove(pt->qEnd, pt->qEnd + off, 4 * (size_t)mc);
		memmove(pt->weight, pt->weight + off, 4 * (size_t)mc);
	}
	pt->len = mc;
	(void)t_len; (void)p;
	return -best;
}


/* ------------------------------------------------------------------ output buffer ---- */

typedef struct { uint8_t *p; size_t len, cap; } obuf;
static void ob_put(obuf *o, const void *src, size_t n) {
	if (o->len + n > o->cap) { o->cap = (o->len + n) * 2 + 4096; o->p = realloc(o->p, o->cap); }
	memcpy(o->p + o->len, src, n); o->len += n;
}

static void unpack(const uint64_t *seq, int len, const int32_t *N, int nN, uint8_t *out) {
	for (int i = 0; i < len; ++i) out[i] = (uint8_t)nuc_at(seq, i);
	for (int i = 0; i < nN; ++i) out[N[i]] = 4;
	out[len] = 0;
}

/* ------------------------------------------------------------------ paired end ------- */

typedef struct {
	int q_len, words, nN, hl, flag;
	uint64_t *w[2]; int32_t *N[2]; uint8_t *b[2];   /* [0] as recorded, [1] reverse complement */
	const uint8_t *hdr;
} pe_mate;

/* update_Scores_se (updatescores.c:300-388): T/S/E/Sc are the candidate arrays (0-based), n candidates */
static int emit_se(obuf *frag, uint64_t *as, uint64_t *uas, const uint8_t *q, int q_len, const uint8_t *hdr, int hl, int flag,
                   double minFrac, int n, int best, int *S, int *E, int *T, const int *Sc) {
	int kept = 0;
	const int mode = minFrac == 1.0 ? 0 : (minFrac < 0 ? 1 : 2);
	const double thr = fabs(minFrac) * best;
	for (int i = 0; i < n; ++i) {
		int keep = mode == 0 ? Sc[i] == best : thr <= Sc[i];
		if (keep) {
			T[kept] = T[i]; S[kept] = S[i]; E[kept] = E[i]; ++kept;
			as[abs(T[kept - 1])] += mode == 1 ? (uint64_t)Sc[i] : (uint64_t)best;
		}
	}
	if (kept == 1) uas[abs(T[0])] += best;
	int32_t h[5] = {q_len, kept, best, hl, flag};
	ob_put(frag, h, 20); ob_put(frag, q, q_len); ob_put(frag, hdr, hl);
	ob_put(frag, S, 4 * (size_t)kept); ob_put(frag, E, 4 * (size_t)kept); ob_put(frag, T, 4 * (size_t)kept);
	return kept;
}

/* update_Scores_pe (updatescores.c:390-488) */
static void emit_pe(obuf *frag, uint64_t *as, uint64_t *uas, const uint8_t *q, int q_len, const uint8_t *hdr, int hl, int flag,
                    const uint8_t *q2, int q2_len, const uint8_t *hdr2, int hl2, int flag2,
                    double minFrac, int n, int best, int *S, int *E, int *T, const int *Sc) {
	int kept = 0;
	const int mode = minFrac == 1.0 ? 0 : (minFrac < 0 ? 1 : 2);
	const double thr = fabs(minFrac) * best;
	for (int i = 0; i < n; ++i) {
		int keep = mode == 0 ? Sc[i] == best : thr <= Sc[i];
		if (keep) {
			T[kept] = T[i]; S[kept] = S[i]; E[kept] = E[i]; ++kept;
			as[abs(T[kept - 1])] += mode == 2 ? (uint64_t)best : (uint64_t)Sc[i];
		}
	}
	if (kept == 1) uas[abs(T[0])] += best;
	int32_t h[5] = {q_len, kept, -best, hl, flag};
	ob_put(frag, h, 20); ob_put(frag, q, q_len); ob_put(frag, hdr, hl);
	ob_put(frag, S, 4 * (size_t)kept); ob_put(frag, E, 4 * (size_t)kept); ob_put(frag, T, 4 * (size_t)kept);
	int32_t h2[3] = {q2_len, hl2, flag2};
	ob_put(frag, h2, 12); ob_put(frag, q2, q2_len); ob_put(frag, hdr2, hl2);
}

/* score of one mate against one template the way alnFragsPenaltyPE rates it (alnfrags.c:1683-1718) */
static int pe_rate(const orc_params *p, const aln_t *a, int q_len, int t_len, int minlen, double mrc, int *start, int *end, double *score) {
	int read_score = a->score;
	if (minlen <= a->len && 0 < read_score && ((mrc * q_len <= a->len - a->qGaps) || (mrc * t_len <= a->len - a->tGaps))) {
		*start = a->pos; *end = a->pos + a->len - a->tGaps;
		if (*start == 0) read_score += -p->Wl;
		if (*end == t_len) read_score += -p->Wl;
		*score = 1.0 * read_score / a->len;
	} else read_score = 0;
	return read_score;
}

/* alnFragsPenaltyPE (alnfrags.c:1596-1972), strands decided by stage 2 (points->len == 0). mt[1..nt] = templates,
 * arrays bT/bTr/bS/bE have nt + 2 entries. cand rows (optional): two per template (mate 1, mate 2). */
static void align_pe(const orc_params *p, nw_ws *ws, mems_t *pt, tindex **tix, orc_db *db, int k, pe_mate *m1, pe_mate *m2,
                     int *mt, int nt, double scoreT, int mq, int minlen, double mrc, double minFrac,
                     int *bT, int *bTr, int *bS, int *bE, obuf *frag, uint64_t *as, uint64_t *uas, obuf *cand, int ridx) {
	int flipped = 0, best1 = 0, best2 = 0, comp = 0, start = 0, end = 0, hits = 0;
	double score = 0;
	int flag = m1->flag, flag_r = m2->flag;
	for (int ti = 1; ti <= nt; ++ti) {
		int at = abs(mt[ti]);
		if (mt[ti] < 0) flipped = 1;   /* both reads are reverse-complemented once, at the first negative template */
		if (!tix[at]) tix[at] = tindex_build(db->seq + db->seq_off[at], db->lengths[at], k);
		const tindex *ix = tix[at];
		const int t_len = db->lengths[at], o = flipped;
		pt->len = 0;
		aln_t a = kma_score(p, ws, ix, m1->b[o], m1->q_len, 0, m1->q_len, m1->w[o], m1->N[o], m1->nN + 1, mq, pt);
		if (cand) { int32_t row[8] = {ridx, mt[ti], a.score, a.len, a.pos, a.match, a.tGaps, a.qGaps}; ob_put(cand, row, 32); }
		int rs = pe_rate(p, &a, m1->q_len, t_len, minlen, mrc, &start, &end, &score);
		if (rs > k && score >= scoreT) { bT[ti] = rs; bS[ti] = start; bE[ti] = end; if (best1 < rs) best1 = rs; }
		else { bT[ti] = 0; bS[ti] = -1; bE[ti] = -1; }
		pt->len = 0;
		a = kma_score(p, ws, ix, m2->b[o], m2->q_len, 0, m2->q_len, m2->w[o], m2->N[o], m2->nN + 1, mq, pt);
		if (cand) { int32_t row[8] = {ridx + 1, mt[ti], a.score, a.len, a.pos, a.match, a.tGaps, a.qGaps}; ob_put(cand, row, 32); }
		rs = pe_rate(p, &a, m2->q_len, t_len, minlen, mrc, &start, &end, &score);
		if (rs > k && score >= scoreT) {
			bTr[ti] = rs;
			if (bT[ti]) { if (start < bS[ti]) bS[ti] = start; else bE[ti] = end; }
			else { bS[ti] = start; bE[ti] = end; }
			if (best2 < rs) best2 = rs;
		} else bTr[ti] = 0;
		rs += bT[ti];
		if (comp < rs) comp = rs;
	}
	if (!best1 && !best2) return;
	const int rc = !flipped;
	const double af = minFrac < 0 ? -minFrac : minFrac;
	const uint8_t *q1 = m1->b[flipped], *q2 = m2->b[flipped];
	int proper, best = 0;
	if (p->apm == 1) {   /* alnFragsUnionPE (alnfrags.c:1408-1422): templates both mates reach within minFrac of their own best */
		if (best1 && best2) {
			const double sc = af * best1, sc_r = af * best2;
			for (int ti = 1; ti <= nt; ++ti)
				if (sc <= bT[ti] && sc_r <= bTr[ti]) { bTr[hits] = bT[ti] + bTr[ti]; bT[hits] = mt[ti]; bS[hits] = bS[ti]; bE[hits] = bE[ti]; ++hits; }
		}
		proper = hits != 0;
		best = best1 + best2;
	} else {             /* alnFragsPenaltyPE (alnfrags.c:1787-1808) */
		proper = comp && af * (best1 + best2) <= (comp + p->PE);
		if (proper) {
			best = comp + p->PE;
			for (int ti = 1; ti <= nt; ++ti)
				if (bT[ti] && bTr[ti]) { bTr[hits] = bT[ti] + bTr[ti] + p->PE; bT[hits] = mt[ti]; bS[hits] = bS[ti]; bE[hits] = bE[ti]; ++hits; }
		}
	}
	if (proper) {   /* proper pair */
		if (bT[0] < 0) {
			for (int i = 0; i < hits; ++i) bT[i] = -bT[i];
			emit_pe(frag, as, uas, q2, m2->q_len, m2->hdr, m2->hl, flag_r, q1, m1->q_len, m1->hdr, m1->hl, flag, minFrac, hits, best, bS, bE, bT, bTr);
		} else {
			if (!rc) { q1 = m1->b[0]; q2 = m2->b[0]; flag ^= 48; flag_r ^= 48; }
			emit_pe(frag, as, uas, q1, m1->q_len, m1->hdr, m1->hl, flag, q2, m2->q_len, m2->hdr, m2->hl, flag_r, minFrac, hits, best, bS, bE, bT, bTr);
		}
	} else if (best1 && best2) {                             /* both map, not as a pair */
		int hits_r = 0, ti = 1, last = nt, tmp;
		const double sc = af * best1, sc_r = af * best2;
		while (ti <= last) {
			if (sc <= bT[ti]) { mt[hits] = mt[ti]; bT[hits] = bT[ti]; bS[hits] = bS[ti]; bE[hits] = bE[ti]; ++hits; ++ti; }
			else if (sc_r <= bTr[ti]) {
				tmp = mt[ti]; mt[ti] = mt[last]; mt[last] = tmp;
				tmp = bTr[ti]; bTr[ti] = bTr[last]; bTr[last] = tmp;
				tmp = bS[ti]; bS[ti] = bS[last]; bS[last] = tmp;
				tmp = bE[ti]; bE[ti] = bE[last]; bE[last] = tmp;
				++hits_r; --last;
			} else ++ti;
		}
		int *bTr2 = bTr + last;
		if (bT[0] < 0) { for (int i = 0; i < hits; ++i) bT[i] = -bT[i]; }
		else if (!rc) { q1 = m1->b[0]; flag ^= 16; flag_r ^= 32; }
		if (bTr2[0] < 0) { for (int i = 0; i < hits_r; ++i) bTr2[i] = -bTr2[i]; }
		else if (!rc) { q2 = m2->b[0]; flag ^= 32; flag_r ^= 16; }
		if (flag & 2) { flag ^= 2; flag_r ^= 2; }
		mt[0] = emit_se(frag, as, uas, q1, m1->q_len, m1->hdr, m1->hl, flag, minFrac, hits, best1, bS, bE, mt, bT);   /* the count lands in slot 0 */
		emit_se(frag, as, uas, q2, m2->q_len, m2->hdr, m2->hl, flag_r, minFrac, hits_r, best2, bS + last, bE + last, mt + last, bTr2);
	} else if (best1) {                                      /* first mate only */
		for (int ti = 1; ti <= nt; ++ti)
			if (bT[ti]) { bTr[hits] = bT[ti]; bT[hits] = mt[ti]; bS[hits] = bS[ti]; bE[hits] = bE[ti]; ++hits; }
		if (bT[0] < 0) { for (int i = 0; i < hits; ++i) bT[i] = -bT[i]; }
		else if (!rc) { q1 = m1->b[0]; flag ^= 16; flag_r ^= 32; }
		flag |= 8; flag_r ^= 4;
		if (flag & 2) { flag ^= 2; flag_r ^= 2; }
		emit_se(frag, as, uas, q1, m1->q_len, m1->hdr, m1->hl, flag, minFrac, hits, best1, bS, bE, bT, bTr);
	} else {                                                 /* second mate only */
		for (int ti = 1; ti <= nt; ++ti)
			if (bTr[ti]) { bTr[hits] = bTr[ti]; bT[hits] = mt[ti]; bS[hits] = bS[ti]; bE[hits] = bE[ti]; ++hits; }
		if (bTr[0] < 0) { for (int i = 0; i < hits; ++i) bTr[i] = -bTr[i]; }
		else if (!rc) { q2 = m2->b[0]; flag ^= 32; flag_r ^= 16; }
		flag_r |= 8; flag ^= 4;
		if (flag_r & 2) { flag ^= 2; flag_r ^= 2; }
		emit_se(frag, as, uas, q2, m2->q_len, m2->hdr, m2->hl, flag_r, minFrac, hits, best2, bS, bE, bT, bTr);
	}
}


/* ------------------------------------------------------------------ traceback alignment (assembly pass) */

/* does byte b occur in q[from, len)? position or -1 (charpos, stdnuc.c:436) */
static int next_n(const uint8_t *q, int from, int len) {
	for (int i = from; i < len; ++i) if (q[i] == 4) return i;
	return -1;
}

static void bytes_rc(uint8_t *q, int len) {   /* strrc, stdnuc.c:450 */
	static const uint8_t comp[6] = {3, 2, 1, 0, 4, 5};
	for (int i = 0, j = len - 1; i < j; ++i, --j) { uint8_t c = comp[q[i]]; q[i] = comp[q[j]]; q[j] = c; }
	if (len & 1) q[len >> 1] = comp[q[len >> 1]];
}

static uint64_t kmer_bytes(const uint8_t *q, int pos, int k) { uint64_t key = 0; for (int b = 0; b < k; ++b) key = (key << 2) | q[pos + b]; return key; }

/* The byte-read seed scan shared by KMA (align.c:246-377) and anker_rc (align.c:823-957): N-free stretches are found
 * with charpos, a stretch is only entered (and re-entered after a MEM) while more than k bases remain before its end,
 * k-mers slide one base at a time. mode 0 = KMA (MEM jumps to its end), mode 1 = anker_rc (also scores the strand).
 * Appends MEMs at pt[n..]; returns the new count, *score the strand score. */
static int scan_bytes(const tindex *ix, const uint8_t *q, int q_len, int q_start, int q_end, int mode, mems_t *pt, int n, int first_i, int *score) {
	const int k = ix->k;
	int i = first_i, sc = 0;
	(void)q_start;
	while (i < q_end) {
		int end = next_n(q, i, q_len);
		if (end == -1) end = q_end;
		int p = i;   /* start of the k-mer under the cursor; the reference's i is p + k - 1 */
		if (!(p < end - k)) { i = end + 1; continue; }
		while (p + k - 1 < end) {
			int slot, cnt, value = tindex_get(ix, kmer_bytes(q, p, k), &slot, &cnt);
			if (value == 0) { ++p; continue; }
			mems_reserve(pt, n + cnt + 1);
			int nextp;
			if (0 < value) {
				mem_from_seed(ix, q, p, value, k, end, &pt->qStart[n], &pt->tStart[n], &pt->qEnd[n], &pt->tEnd[n]);
				pt->weight[n] = mode ? pt->tEnd[n] - pt->tStart[n] : pt->qEnd[n] - pt->qStart[n];
				sc += pt->qEnd[n] - pt->qStart[n];
				nextp = pt->qEnd[n];
				++n;
			} else {
				int bias = p;
				sc += k;
				for (int c = 0; c < cnt; ++c) {
					mem_from_seed(ix, q, p, ix->pos[slot + c], k, end, &pt->qStart[n], &pt->tStart[n], &pt->qEnd[n], &pt->tEnd[n]);
					pt->weight[n] = pt->qEnd[n] - pt->qStart[n];
					if (bias < pt->qEnd[n]) bias = pt->qEnd[n];
					++n;
				}
				sc += bias - p;
				nextp = bias + 1;
			}
			if (nextp < end - k) p = nextp; else { p = end + 1; break; }   /* "update position" (align.c:309-315) */
		}
		i = end + 1;
	}
	if (score) *score = sc;
	return n;
}

/* anker_rc (align.c:780-991): MEMs of the byte read on both strands, the better strand's MEMs stay in pt and the read
 * is left in that orientation. Returns the winning strand score (0: nothing). */
static int pick_strand_bytes(const tindex *ix, uint8_t *q, int q_len, int q_start, int q_end, int one2one, int exhaustive, mems_t *pt) {
	const int k = ix->k;
	int sf = 0, sr = 0;
	pt->len = 0;
	/* align.c:813-817: a lower query bound skips preseed; preseed returns 0 on a hit, >= q_len otherwise */
	int first = q_start ? q_start : (exhaustive || preseed_hit(ix, q, q_len, q_end - q_start) ? 0 : q_len);
	int nf = scan_bytes(ix, q, q_len, q_start, q_end, 1, pt, 0, first, &sf);
	bytes_rc(q, q_len);
	/* the bounds mirrored (align.c:808-812) */
	int ntot = scan_bytes(ix, q, q_len, q_len - q_end, q_len - q_start, 1, pt, nf, q_len - q_end, &sr);
	int best = sf < sr ? sr : sf;
	if (one2one && best < k && best * k < (q_len - k - best)) { pt->len = 0; return 0; }   /* read stays reverse-complemented */
	if (best == sf) { bytes_rc(q, q_len); pt->len = nf; return best; }
	int mc = ntot - nf;
	if (nf) {
		memmove(pt->tStart, pt->tStart + nf, 4 * (size_t)mc); memmove(pt->tEnd, pt->tEnd + nf, 4 * (size_t)mc);
		memmove(pt->qStart, pt->qStart + nf, 4 * (size_t)mc); memmove(pt->qEnd, pt->qEnd + nf, 4 * (size_t)mc);
		memmove(pt->weight, pt->weight + nf, 4 * (size_t)mc);
	}
	pt->len = mc;
	return best;
}

typedef struct { astr a; int len; astr frag; int cap; } trace_buf;

static void tb_reserve(trace_buf *b, int need) {
	if (need <= b->cap) return;
	b->cap = 2 * need + 64;
	b->a.t = realloc(b->a.t, b->cap); b->a.s = realloc(b->a.s, b->cap); b->a.q = realloc(b->a.q, b->cap);
	b->frag.t = realloc(b->frag.t, b->cap); b->frag.s = realloc(b->frag.s, b->cap); b->frag.q = realloc(b->frag.q, b->cap);
}

static aln_t nw_auto_str(const orc_params *p, nw_ws *w, const uint64_t *tseq, const uint8_t *q, int k,
                         int t_s, int t_e, int q_s, int q_e, astr *so) {
	int band = abs((t_e - t_s) - (q_e - q_s)) + BANDW;
	if (q_e - q_s <= band || t_e - t_s <= band) return nw_full(p, w, tseq, q, k, t_s, t_e, q_s, q_e, so);
	return nw_band(p, w, tseq, q, k, t_s, t_e, q_s, q_e, band, so);
}

/* KMA (align.c:214-507): seed, chain, stitch -- with the aligned rows. tb->a receives t/s/q, tb->len columns. */
/* -ts (trimSeeds, chain.c:496-538; called by KMA only, align.c:413) */
static double g_min_frac = 1.0;
void orc_align_set_minfrac(double f) { g_min_frac = f; }   /* minFrac of -proxi as runKMA hands it to the alignment threads (kma.c:1622, runkma.c:351) */

static int g_trim_seeds = 0;
void orc_trace_set_ts(int ts) { g_trim_seeds = ts; }

static aln_t kma_trace(const orc_params *p, nw_ws *w, const tindex *ix, const uint8_t *q, int q_len, int q_start, int q_end, int mq, mems_t *pt, trace_buf *tb) {
	const int k = ix->k, t_len = ix->len, U = p->U, M = p->M;
	tb_reserve(tb, 2 * q_len + 2 * BANDW + 64);
	tb->len = 0;
	int n = pt->len;
	if (!n) n = scan_bytes(ix, q, q_len, q_start, q_end, 0, pt, 0, q_start, 0);
	pt->len = n;
	if (!n) return aln_zero();
	unsigned mapQ = 0;
	int start = chain_mems(p, pt, q_len, t_len, k, &mapQ);
	if (mapQ < (unsigned)mq || pt->score[start] < k) { pt->len = 0; return aln_zero(); }

	/* trimSeeds (chain.c:496-538): the first ts bases of every seed of the chain go back to the DP (all but one base of a
	 * seed shorter than ts); the first seed keeps its start when it begins at the query start */
	if (g_trim_seeds) {
		int c = start, go = 1;   /* MEM 0 is a valid chain start: only next[] == 0 ends the walk (do ... while, chain.c:509-524) */
		if (!pt->qStart[c]) { c = pt->next[c]; go = c != 0; }
		while (go) {
			int len = pt->qEnd[c] - pt->qStart[c];
			const int cut = len < g_trim_seeds ? len - 1 : g_trim_seeds;
			pt->tStart[c] += cut; pt->qStart[c] += cut;
			c = pt->next[c];
			go = c != 0;
		}
	}

	/* leading tail (leadTailAln with Frag_align, align.c:53-138): leading gap columns are trimmed when the window
	 * starts at the template start */
	aln_t s = {0, 0, pt->tStart[start] - 1, 0, 0, 0};
	{
		const int t_e = pt->tStart[start] - 1, q_e = pt->qStart[start];
		if (q_e) {
			int t_s = 0, q_s = 0;
			if ((q_e << 1) < t_e || (q_e + BANDW) < t_e) t_s = t_e - (q_e + IMIN(q_e, BANDW));
			else if ((t_e << 1) < q_e || (t_e + BANDW) < q_e) q_s = q_e - (t_e + IMIN(t_e, BANDW));
			if (t_e - t_s > 0 && q_e - q_s > 0) {
				tb_reserve(tb, tb->len + (t_e - t_s) + (q_e - q_s) + 8);
				aln_t a = nw_auto_str(p, w, ix->seq, q, -1 - (t_s == 0), t_s, t_e, q_s, q_e, &tb->frag);
				int bias = 0;
				if (t_s == 0) {
					while (bias < a.len && (tb->frag.t[bias] == 5 || tb->frag.q[bias] == 5)) {
						if (tb->frag.t[bias] == 5) --a.tGaps; else --a.qGaps;
						++bias;
					}
					a.len -= bias;
				}
				memcpy(tb->a.t, tb->frag.t + bias, a.len); memcpy(tb->a.s, tb->frag.s + bias, a.len); memcpy(tb->a.q, tb->frag.q + bias, a.len);
				s.pos -= a.len - a.tGaps;
				s.score = a.score; s.len = a.len; s.match = a.match; s.tGaps = a.tGaps; s.qGaps = a.qGaps;
			}
		}
	}
	for (;;) {
		const int q_s0 = pt->qStart[start], len = pt->qEnd[start] - q_s0;
		tb_reserve(tb, s.len + len + 8);
		memcpy(tb->a.t + s.len, q + q_s0, len); memset(tb->a.s + s.len, '|', len); memcpy(tb->a.q + s.len, q + q_s0, len);
		s.len += len; s.match += len;
		for (int i = q_s0; i < pt->qEnd[start]; ++i) s.score += p->d[q[i] * 5 + q[i]];
		if (!pt->next[start]) break;
		int q_s = pt->qEnd[start], t_s = pt->tEnd[start] - 1, t_e, t_l, q_e;
		start = pt->next[start];
		if (pt->qStart[start] < q_s) { pt->tStart[start] += q_s - pt->qStart[start]; pt->qStart[start] = q_s; }
		t_e = pt->tStart[start] - 1;
		if (t_e < t_s) {
			if (t_s <= pt->tEnd[start]) { pt->qStart[start] += t_s - t_e; t_e = t_s; t_l = 0; }
			else t_l = t_len - t_s + t_e;
		} else t_l = t_e - t_s;
		q_e = pt->qStart[start];
		if (abs(t_l - q_e + q_s) * U > q_len * M || t_l > q_len || q_e - q_s > (q_len >> 1)) {
			int keep = s.pos; pt->len = 0; tb->len = 0; s = aln_zero(); s.pos = keep; return s;
		}
		if (t_l > 0 || q_e - q_s > 0) {
			tb_reserve(tb, s.len + t_l + (q_e - q_s) + 8);
			aln_t a = nw_auto_str(p, w, ix->seq, q, 0, t_s, t_e, q_s, q_e, &tb->frag);
			memcpy(tb->a.t + s.len, tb->frag.t, a.len); memcpy(tb->a.s + s.len, tb->frag.s, a.len); memcpy(tb->a.q + s.len, tb->frag.q, a.len);
			s.score += a.score; s.len += a.len; s.match += a.match; s.tGaps += a.tGaps; s.qGaps += a.qGaps;
		}
	}
	/* trailing tail (trailTailAln with Frag_align, align.c:147-212): trailing gap columns trimmed at the template end */
	{
		const int t_s = pt->tEnd[start] - 1, q_s = pt->qEnd[start];
		int q_e = q_len, t_e = t_len;
		if (((q_len - q_s) << 1) < (t_len - t_s) || (q_len - q_s + BANDW) < (t_len - t_s)) { t_e = q_len - q_s; t_e = t_s + (t_e + IMIN(t_e, BANDW)); }
		else if (((t_len - t_s) << 1) < (q_len - q_s) || (t_len - t_s + BANDW) < (q_len - q_s)) { q_e = t_len - t_s; q_e = q_s + (q_e + IMIN(q_e, BANDW)); }
		if (t_e - t_s > 0 && q_e - q_s > 0) {
			tb_reserve(tb, s.len + (t_e - t_s) + (q_e - q_s) + 8);
			aln_t a = nw_auto_str(p, w, ix->seq, q, 1 + (t_e == t_len), t_s, t_e, q_s, q_e, &tb->frag);
			if (t_e == t_len) {
				int bias = a.len - 1;
				while (bias && (tb->frag.t[bias] == 5 || tb->frag.q[bias] == 5)) {
					if (tb->frag.t[bias] == 5) --a.tGaps; else --a.qGaps;
					--bias;
				}
				++bias;
				if (bias != a.len) a.len = bias;
			}
			memcpy(tb->a.t + s.len, tb->frag.t, a.len); memcpy(tb->a.s + s.len, tb->frag.s, a.len); memcpy(tb->a.q + s.len, tb->frag.q, a.len);
			s.score += a.score; s.len += a.len; s.match += a.match; s.tGaps += a.tGaps; s.qGaps += a.qGaps;
		}
	}
	tb->len = s.len;
	pt->len = 0;
	return s;
}

/* The alignment part of assemble_KMA's inner loop (assembly.c:1868-1961) over a stream of per-template fragment
 * records (frags.c:45-48): int32[8]{template, q_len, nHits, score, start, end, hdrlen, flag} + read bytes (0-4) +
 * header. Per record the output holds int32[12]{accepted, read_score, start, end, score, len, pos, match, tGaps,
 * qGaps, oriented (1: the read was reverse-complemented by anker_rc), ncol} + t[ncol] s[ncol] q[ncol]. */
int orc_trace_stream(orc_db *db, const char *prefix, const orc_params *p, const uint8_t *in, size_t in_bytes, int one2one,
                     double scoreT, int mq, int minlen, double mrc, uint8_t **out, size_t *out_bytes) {
	if (orc_db_load_seq(db, prefix)) return -1;
	int k = db->lengths[0];
	if (k < 4 || 31 < k) k = 16;
	const int Wl = -p->Wl;
	tindex **tix = calloc(db->DB_size, sizeof(tindex *));
	nw_ws ws; memset(&ws, 0, sizeof(ws));
	mems_t pt; memset(&pt, 0, sizeof(pt));
	trace_buf tb; memset(&tb, 0, sizeof(tb));
	obuf o = {0, 0, 0};
	size_t ip = 0;
	uint8_t *q = 0; size_t qcap = 0;
	while (ip + 32 <= in_bytes) {
		int32_t h[8]; memcpy(h, in + ip, 32);
		if (h[0] < 0) break;
		ip += 32;
		const int tmpl = h[0], q_len = h[1], hl = h[6];
		int read_score = h[3];
		if ((size_t)q_len + 64 > qcap) { qcap = 2 * (size_t)q_len + 64; q = realloc(q, qcap); }
		memcpy(q, in + ip, q_len); memset(q + q_len, 0, 32); ip += (size_t)q_len + hl;
		if (!tix[tmpl]) tix[tmpl] = tindex_build(db->seq + db->seq_off[tmpl], db->lengths[tmpl], k);
		const tindex *ix = tix[tmpl];
		const int t_len = ix->len;
		int32_t r[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
		tb.len = 0;
		pt.len = 0;
		int go = read_score != 0;
		/* q-bound of chain-mode records (assembly.c:1916-1923) */
		int q_start = 0, q_end = q_len;
		if (9 < hl && in[ip - 9] == 0) { memcpy(&q_start, in + ip - 8, 4); memcpy(&q_end, in + ip - 4, 4); }
		if (!go) {
			uint8_t first = q_len ? q[0] : 0, last = q_len ? q[q_len - 1] : 0;
			go = pick_strand_bytes(ix, q, q_len, q_start, q_end, one2one, p->exhaustive, &pt) != 0;
			/* did the read end up reverse-complemented? compare with a fresh copy */
			r[10] = q_len && memcmp(q, in + ip - hl - q_len, q_len) != 0;
			(void)first; (void)last;
		}
		if (go) {
			aln_t a = kma_trace(p, &ws, ix, q, q_len, q_start, q_end, mq, &pt, &tb);
			int aln_len = a.len, start = a.pos, end = start + aln_len - a.tGaps;
			if (t_len < end) end -= t_len;
			read_score = a.score;
			if (start == 0) read_score += Wl;
			if (end == t_len) read_score += Wl;
			double score;
			if (minlen <= aln_len && ((mrc * q_len <= a.len - a.qGaps) || (mrc * t_len <= a.len - a.tGaps))) score = 1.0 * read_score / aln_len;
			else { read_score = 0; score = 0; }
			r[0] = 0 < read_score && scoreT <= score;
			r[1] = read_score; r[2] = start; r[3] = end;
			r[4] = a.score; r[5] = a.len; r[6] = a.pos; r[7] = a.match; r[8] = a.tGaps; r[9] = a.qGaps;
		}
		pt.len = 0;
		r[11] = tb.len;   /* columns of the rows that follow (0 when nothing aligned) */
		ob_put(&o, r, 48);
		ob_put(&o, tb.a.t, tb.len); ob_put(&o, tb.a.s, tb.len); ob_put(&o, tb.a.q, tb.len);
	}
	for (int t = 0; t < db->DB_size; ++t) tindex_free(tix[t]);
	free(tix); free(q);
	free(pt.tStart); free(pt.tEnd); free(pt.qStart); free(pt.qEnd); free(pt.weight); free(pt.score); free(pt.next);
	for (int i = 0; i < 2; ++i) { free(ws.D[i]); free(ws.P[i]); } free(ws.E);
	free(tb.a.t); free(tb.a.s); free(tb.a.q); free(tb.frag.t); free(tb.frag.s); free(tb.frag.q);
	*out = o.p; *out_bytes = o.len;
	return 0;
}

/* ------------------------------------------------------------------ the stream ------- */

/* stage-2 stream (single-end records and -apm p pairs) -> frag_raw stream (without the final int32 0 of runkma.c:444) and the two
 * ConClave accumulators. cand (optional): 8 x int32 per (read, candidate), same rows as ref_harness.c writes.
 * Returned buffers are malloc'ed; free with orc_free. */
int orc_align_stream(orc_db *db, const char *prefix, const orc_params *p, const uint8_t *in, size_t in_bytes,
                     int one2one, double scoreT, int mq, int minlen, double mrc,
                     uint8_t **frag_out, size_t *frag_bytes, uint64_t *as, uint64_t *uas,
                     int32_t **cand_out, size_t *cand_rows, int64_t *nw_cells) {
	if (orc_db_load_seq(db, prefix)) return -1;
	int k = db->lengths[0];
	if (k < 4 || 31 < k) k = 16;
	const int Wl = -p->Wl;
	(void)Wl;
	tindex **tix = calloc(db->DB_size, sizeof(tindex *));
	nw_ws ws; memset(&ws, 0, sizeof(ws));
	mems_t pt; memset(&pt, 0, sizeof(pt));
	obuf frag = {0, 0, 0}, cand = {0, 0, 0};
	size_t ip = 0, cap = 0;
	uint64_t *seq = 0, *rseq = 0; int32_t *N = 0, *rN = 0; uint8_t *qf = 0, *qr = 0;
	int *bT = 0, *bS = 0, *bE = 0, *Sc = 0, *Ln = 0; int bcap = 0;
	int ridx = 0;
	memset(as, 0, 8 * (size_t)db->DB_size); memset(uas, 0, 8 * (size_t)db->DB_size);

	while (ip + 28 <= in_bytes) {
		int32_t h[7]; memcpy(h, in + ip, 28);
		if (h[0] < 0) break;
		ip += 28;
		const int q_len = h[0], words = h[1], nN = h[2], rc_flag = h[3], nt = h[4], hl = h[5];
		int flag = h[6];
		if ((size_t)words + 2 > cap) {
			cap = 2 * (size_t)words + 2;
			seq = realloc(seq, 8 * cap); rseq = realloc(rseq, 8 * cap);
			qf = realloc(qf, 64 * cap + 64); qr = realloc(qr, 64 * cap + 64);
		}
		N = realloc(N, 4 * (size_t)(nN + 2)); rN = realloc(rN, 4 * (size_t)(nN + 2));
		memcpy(seq, in + ip, 8 * (size_t)words); seq[words] = 0; ip += 8 * (size_t)words;
		memcpy(N, in + ip, 4 * (size_t)nN); ip += 4 * (size_t)nN;
		const int32_t *T = (const int32_t *)0; int32_t *Tbuf = malloc(4 * (size_t)(nt + 1));
		memcpy(Tbuf, in + ip, 4 * (size_t)nt); ip += 4 * (size_t)nt; T = Tbuf;
		const uint8_t *hdr = in + ip; ip += hl;
		if (nt == 0) {   /* first record of a pair (printPair, ankers.c:150): the mate follows and carries the templates */
			free(Tbuf);
			if (ip + 28 > in_bytes) break;
			int32_t g[7]; memcpy(g, in + ip, 28); ip += 28;
			pe_mate m1, m2;
			memset(&m1, 0, sizeof(m1)); memset(&m2, 0, sizeof(m2));
			m1.q_len = q_len; m1.words = words; m1.nN = nN; m1.hl = hl; m1.flag = flag; m1.hdr = hdr;
			m2.q_len = g[0]; m2.words = g[1]; m2.nN = g[2]; m2.hl = g[5]; m2.flag = g[6];
			const int nt2 = g[4];
			pe_mate *mm[2] = {&m1, &m2};
			const uint64_t *src[2] = {seq, (const uint64_t *)0};
			for (int x = 0; x < 2; ++x) {
				pe_mate *m = mm[x];
				for (int o = 0; o < 2; ++o) {
					m->w[o] = calloc((size_t)m->words + 2, 8); m->N[o] = calloc((size_t)m->nN + 2, 4); m->b[o] = calloc((size_t)m->q_len + 64, 1);
				}
				if (x == 0) { memcpy(m->w[0], src[0], 8 * (size_t)m->words); memcpy(m->N[0], N, 4 * (size_t)m->nN); }
				else {
					memcpy(m->w[0], in + ip, 8 * (size_t)m->words); ip += 8 * (size_t)m->words;
					memcpy(m->N[0], in + ip, 4 * (size_t)m->nN); ip += 4 * (size_t)m->nN;
				}
				orc_revcomp(m->w[0], m->q_len, m->N[0], m->nN, m->w[1], m->N[1]);
				for (int o = 0; o < 2; ++o) { unpack(m->w[o], m->q_len, m->N[o], m->nN, m->b[o]); m->N[o][m->nN] = m->q_len; }
			}
			int *mt = malloc(4 * (size_t)(nt2 + 2));
			mt[0] = nt2;
			memcpy(mt + 1, in + ip, 4 * (size_t)nt2); ip += 4 * (size_t)nt2;
			m2.hdr = in + ip; ip += m2.hl;
			if (rc_flag < 0) { fprintf(stderr, "orc_align_stream: strand-undecided pairs are not produced by -apm p\n"); return -2; }
			if (k <= m1.q_len && k <= m2.q_len) {
				int *bT = calloc((size_t)nt2 + 2, 4), *bTr = calloc((size_t)nt2 + 2, 4), *bS2 = calloc((size_t)nt2 + 2, 4), *bE2 = calloc((size_t)nt2 + 2, 4);
				align_pe(p, &ws, &pt, tix, db, k, &m1, &m2, mt, nt2, scoreT, mq, minlen, mrc, g_min_frac, bT, bTr, bS2, bE2, &frag, as, uas,
				         cand_out ? &cand : 0, ridx);
				free(bT); free(bTr); free(bS2); free(bE2);
			}
			for (int x = 0; x < 2; ++x) for (int o = 0; o < 2; ++o) { free(mm[x]->w[o]); free(mm[x]->N[o]); free(mm[x]->b[o]); }
			free(mt);
			ridx += 2;
			continue;
		}
		if (q_len < k) { free(Tbuf); ++ridx; continue; }

		if (rc_flag < 0) { orc_revcomp(seq, q_len, N, nN, rseq, rN); rseq[words] = 0; unpack(rseq, q_len, rN, nN, qr); rN[nN] = q_len; }
		unpack(seq, q_len, N, nN, qf); N[nN] = q_len;
		if (nt > bcap) { bcap = 2 * nt; bT = realloc(bT, 4 * bcap); bS = realloc(bS, 4 * bcap); bE = realloc(bE, 4 * bcap); Sc = realloc(Sc, 4 * bcap); Ln = realloc(Ln, 4 * bcap); }

		double bestScore = 0; int best_read = 0, hits = 0;
		const int arc = rc_flag < 0;
		pt.len = 0;
		int q_start = 0, q_end = q_len;   /* q-bound of a chain-mode record (alnfrags.c:1091-1099, qseqs.c:41) */
		if (9 < hl && hdr[hl - 9] == 0) { memcpy(&q_start, hdr + hl - 8, 4); memcpy(&q_end, hdr + hl - 4, 4); }
		for (int ti = 0; ti < nt; ++ti) {
			int tmpl = T[ti], at = abs(tmpl);
			if (!tix[at]) tix[at] = tindex_build(db->seq + db->seq_off[at], db->lengths[at], k);
			const tindex *ix = tix[at];
			aln_t a;
			if (arc) {
				int rc = pick_strand(p, ix, qf, qr, seq, rseq, N, rN, nN + 1, q_len, q_start, q_end, one2one, &pt);
				if (rc < 0) { tmpl = -at; a = kma_score(p, &ws, ix, qr, q_len, q_len - q_end, q_len - q_start, rseq, rN, nN + 1, mq, &pt); }
				else if (rc) { tmpl = at; a = kma_score(p, &ws, ix, qf, q_len, q_start, q_end, seq, N, nN + 1, mq, &pt); }
				else { memset(&a, 0, sizeof(a)); pt.len = 0; }
			} else if (tmpl < 0) a = kma_score(p, &ws, ix, qr, q_len, q_len - q_end, q_len - q_start, rseq, rN, nN + 1, mq, &pt);
			else a = kma_score(p, &ws, ix, qf, q_len, q_start, q_end, seq, N, nN + 1, mq, &pt);
			if (cand_out) { int32_t row[8] = {ridx, tmpl, a.score, a.len, a.pos, a.match, a.tGaps, a.qGaps}; ob_put(&cand, row, 32); }

			int aln_len = a.len, start = a.pos, end = start + aln_len - a.tGaps, t_len = db->lengths[at], read_score;
			double score;
			if (t_len < end) end -= t_len;
			if (q_len <= aln_len || t_len <= aln_len) score = aln_len; else score = q_len < t_len ? q_len : t_len;
			read_score = a.score;
			if (minlen <= aln_len && ((mrc * q_len <= a.len - a.qGaps) || (mrc * t_len <= a.len - a.tGaps))) score = read_score / score;
			else { read_score = 0; score = 0; }
			if (k < read_score && scoreT <= score) {
				bT[hits] = tmpl; bS[hits] = start; bE[hits] = end; Sc[hits] = read_score; Ln[hits] = aln_len; ++hits;
				if (bestScore < score) bestScore = score;
				if (best_read < read_score) best_read = read_score;
			}
		}
		/* note: a reverse-strand record (rc_flag < 0 path absent) aligns the strand stage 2 wrote: qf */
		if (best_read > k) {   /* update_Scores (updatescores.c:203-298), its three minFrac branches */
			int kept = 0;
			const double mf = g_min_frac < 0 ? -g_min_frac : g_min_frac;
			const double minScoreP = mf * bestScore, minFracP = mf * best_read;
			for (int i = 0; i < hits; ++i) {
				int keep;
				if (g_min_frac == 1.0) { double minScore = Sc[i] / Ln[i]; keep = minScore == bestScore || Sc[i] == best_read; }
				else keep = (Ln[i] * minScoreP <= Sc[i]) || minFracP <= Sc[i];
				if (keep) {
					bT[kept] = bT[i]; bS[kept] = bS[i]; bE[kept] = bE[i]; ++kept;
					as[abs(bT[kept - 1])] += (g_min_frac == 1.0 || g_min_frac < 0) ? Sc[i] : best_read;
				}
			}
			if (kept == 1) uas[abs(bT[0])] += best_read;
			int32_t b[5] = {q_len, kept, best_read, hl, flag};
			ob_put(&frag, b, 20); ob_put(&frag, qf, q_len); ob_put(&frag, hdr, hl);
			ob_put(&frag, bS, 4 * (size_t)kept); ob_put(&frag, bE, 4 * (size_t)kept); ob_put(&frag, bT, 4 * (size_t)kept);
		}
		free(Tbuf);
		++ridx;
	}
	for (int t = 0; t < db->DB_size; ++t) tindex_free(tix[t]);
	free(tix); free(seq); free(rseq); free(N); free(rN); free(qf); free(qr);
	free(bT); free(bS); free(bE); free(Sc); free(Ln);
	free(pt.tStart); free(pt.tEnd); free(pt.qStart); free(pt.qEnd); free(pt.weight); free(pt.score); free(pt.next);
	for (int i = 0; i < 2; ++i) { free(ws.D[i]); free(ws.P[i]); } free(ws.E);
	*frag_out = frag.p; *frag_bytes = frag.len;
	if (cand_out) { *cand_out = (int32_t *)cand.p; *cand_rows = cand.len / 32; }
	if (nw_cells) *nw_cells = ws.cells;
	return 0;
}


/* ------------------------------------------------------------------ -mem_mode -------- */

/* The "Collecting k-mer scores" loop of runKMA_MEM (runkma.c:1088-1140) with update_Scores_MEM (updatescores.c:26) and
 * update_Scores_pe_MEM (:64): in -mem_mode the alignment pass is skipped, every stage-2 record becomes a frag_raw record
 * whose hits are its candidate templates spanning the whole template (start 0, end = template length), scored with
 * the k-mer score of stage 2; a record with one candidate also adds to the unique scores. lengths[0] = k of the
 * alignment index. Returns 0; *frag_out is malloc'ed. */
int orc_memscore_stream(const int32_t *lengths, int DB_size, const uint8_t *in, size_t in_bytes, uint8_t **frag_out, size_t *frag_bytes,
                        uint64_t *as, uint64_t *uas) {
	int k = lengths[0];
	if (k < 4 || 31 < k) k = 16;
	obuf o = {0, 0, 0};
	size_t ip = 0;
	uint8_t *q = 0, *q2 = 0; size_t qcap = 0;
	memset(as, 0, 8 * (size_t)DB_size); memset(uas, 0, 8 * (size_t)DB_size);
	while (ip + 28 <= in_bytes) {
		int32_t h[7], g[7];
		memcpy(h, in + ip, 28);
		if (h[0] < 0) break;
		const uint8_t *rec1 = in + ip;
		ip += 28 + 8 * (size_t)h[1] + 4 * (size_t)h[2] + 4 * (size_t)h[4] + (size_t)h[5];
		const uint8_t *recT = rec1;   /* the record that carries the templates */
		int read_score = 0, pe = 0;
		if (h[4] == 0) {              /* first record of a pair: the mate follows (ankers.c:150) */
			if (ip + 28 > in_bytes) break;
			memcpy(g, in + ip, 28);
			recT = in + ip;
			ip += 28 + 8 * (size_t)g[1] + 4 * (size_t)g[2] + 4 * (size_t)g[4] + (size_t)g[5];
			read_score = abs(g[3]);
			pe = 1;
		} else memcpy(g, h, 28);
		const int q_len = h[0];
		if (q_len < k) continue;
		if ((size_t)(q_len > g[0] ? q_len : g[0]) + 64 > qcap) { qcap = 2 * (size_t)(q_len > g[0] ? q_len : g[0]) + 64; q = realloc(q, qcap); q2 = realloc(q2, qcap); }
		{   /* unCompDNA of the first record */
			uint64_t *w = malloc(8 * ((size_t)h[1] + 1)); int32_t *N = malloc(4 * ((size_t)h[2] + 1));
			memcpy(w, rec1 + 28, 8 * (size_t)h[1]); memcpy(N, rec1 + 28 + 8 * (size_t)h[1], 4 * (size_t)h[2]);
			unpack(w, q_len, N, h[2], q);
			free(w); free(N);
		}
		const int nt = g[4];
		const uint8_t *T = recT + 28 + 8 * (size_t)g[1] + 4 * (size_t)g[2];
		int32_t last;
		memcpy(&last, T + 4 * (size_t)(nt - 1), 4);
		int bestHits = nt;
		if (h[3] < 0 && 0 < last) bestHits = -bestHits;
		const int two = pe && read_score && k <= g[0];
		const int score = abs(h[3]) + (two ? read_score : 0);
		int32_t b[5] = {q_len, bestHits, two ? -score : score, h[5], h[6]};
		const uint8_t *hdr1 = rec1 + 28 + 8 * (size_t)h[1] + 4 * (size_t)h[2] + 4 * (size_t)h[4];
		ob_put(&o, b, 20); ob_put(&o, q, q_len); ob_put(&o, hdr1, h[5]);
		for (int i = 0; i < nt; ++i) { int32_t z = 0; ob_put(&o, &z, 4); }
		for (int i = 0; i < nt; ++i) { int32_t t; memcpy(&t, T + 4 * (size_t)i, 4); int32_t e = lengths[abs(t)]; ob_put(&o, &e, 4); }
		ob_put(&o, T, 4 * (size_t)nt);
		if (two) {
			uint64_t *w = malloc(8 * ((size_t)g[1] + 1)); int32_t *N = malloc(4 * ((size_t)g[2] + 1));
			memcpy(w, recT + 28, 8 * (size_t)g[1]); memcpy(N, recT + 28 + 8 * (size_t)g[1], 4 * (size_t)g[2]);
			unpack(w, g[0], N, g[2], q2);
			free(w); free(N);
			int32_t m[3] = {g[0], g[5], g[6]};
			ob_put(&o, m, 12); ob_put(&o, q2, g[0]); ob_put(&o, T + 4 * (size_t)nt, g[5]);
		}
		if (nt == 1) { as[abs(last)] += score; uas[abs(last)] += score; }
		else for (int i = 0; i < nt; ++i) { int32_t t; memcpy(&t, T + 4 * (size_t)i, 4); as[abs(t)] += score; }
	}
	free(q); free(q2);
	*frag_out = o.p; *frag_bytes = o.len;
	return 0;
}

/* ------------------------------------------------------------------ base-count matrix - */

/* alnToMat (assembly.c:1317-1444) restricted to the template nodes, and alnToMatDense (assembly.c:1446-1497): the
 * per-position counts counts[6] = {A, C, G, T, N, gap} every accepted alignment of the traceback pass adds to its
 * template. Insertion nodes (alnToMat's linked nodes behind position t_len; order dependent) are not
 * restated: an insertion column leaves the template position where it is. Inputs are the fragment records and the
 * trace output they produced (orc_trace_stream layout). counts = uint16 [sum of template lengths][6], template t at
 * offset sum(len[1..t-1]); increments saturate at 65535 (assembly.c:1436). The reference's quirks are kept: alnToMat
 * trims gap columns at both ends (its trailing loop stops at column 0), alnToMatDense only at the end. */
int orc_matrix_stream(const int32_t *lengths, int DB_size, const uint8_t *frags, size_t fb, const uint8_t *trace, size_t tb,
                      int dense, uint16_t *counts) {
	int64_t *off = malloc(8 * (size_t)(DB_size + 1));
	off[0] = off[1] = 0;
	for (int t = 2; t <= DB_size; ++t) off[t] = off[t - 1] + lengths[t - 1];
	size_t ip = 0, tp = 0;
	int n = 0;
	while (ip + 32 <= fb && tp + 48 <= tb) {
		int32_t h[8], r[12];
		memcpy(h, frags + ip, 32);
		if (h[0] < 0) break;
		ip += 32 + (size_t)h[1] + (size_t)h[6];
		memcpy(r, trace + tp, 48); tp += 48;
		const int ncol = r[11];
		const uint8_t *t = trace + tp, *q = t + 2 * (size_t)ncol;
		tp += 3 * (size_t)ncol;
		if (!r[0]) continue;
		const int tmpl = h[0], t_len = lengths[tmpl];
		uint16_t *C = counts + 6 * off[tmpl];
		int aln_len = r[5], start = r[6], i;
		if (dense) {
			i = aln_len - 1;
			while (i >= 0 && (t[i] == 5 || q[i] == 5)) --i;   /* the reference has no lower bound here */
			aln_len = i + 1;
			i = 0;
		} else {
			i = aln_len - 1;
			while (i && (t[i] == 5 || q[i] == 5)) --i;
			aln_len = i + 1;
			i = 0;
			while (i < aln_len && (t[i] == 5 || q[i] == 5)) { if (q[i] == 5) ++start; ++i; }
		}
		int pos = start;
		for (; i < aln_len; ++i) {
			if (t[i] == 5) continue;
			if (pos >= t_len) pos -= t_len;   /* next of the last template node is node 0 */
			uint16_t *c = C + 6 * (size_t)pos + q[i];
			if (!++*c) *c = 65535;
			++pos;
		}
		++n;
	}
	free(off);
	return n;
}

void orc_free(void *p) { free(p); }
