/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of KMA 1.5.1's stage-2 seeding path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this. The product (kma_b200/csrc) never links or calls it.
 *
 * Parity pin: checked byte-for-byte against the stage-2 stream of the unmodified reference
 * (`oracle/_ref/kma ... -s2`, built by oracle/Makefile.ref from /root/reference) in
 * tests/test_oracle_seed.py, and against the committed streams in tests/golden/.
 *
 * Restated (not copied) from:
 *   hashmapkma.c:149-178  hashMap_getGlobal   (flag == 0 path)      -> orc_lookup
 *   hashmapkma.c:264-273  megaMap_getGlobal                          -> orc_lookup (mega)
 *   hashmapkma.c:275-455  hashMapKMA_load     (.comp.b layout)      -> orc_db_open
 *   compdna.c:228-256     rc_comp                                    -> orc_revcomp
 *   savekmers.c:2442-3065 save_kmers (-1t1)                          -> orc_seed_read
 *   savekmers.c:273-294   getBestMatch                               -> best_set
 *   ankers.c:30-50        print_ankers (stage-2 record)              -> emit_record
 *   savekmers.c:50-92     loadFsa (stage-1 record)                   -> load_mate / orc_seed_stream
 *   savekmers.c:427-688   get_kmers_for_pair                         -> pair_kmers
 *   savekmers.c:1383-1511, 1648-1680 getFirstPen / getSecondBestPen / getF_Best -> first_pen / second_pen / f_best
 *   savekmers.c:3572-3777 save_kmers_penaltyPair (-apm p)            -> seed_pair
 *   ankers.c:150-161      printPair                                  -> seed_pair
 *   kmers.c:257           stream terminator  int32 -(#reads)
 *
 * Structure differs from the reference on purpose: the k-mer scan is expressed as a stream of
 * (position, value-list offset) hits feeding one scoring state machine that is shared by both
 * strands, instead of two unrolled copies.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "orc.h"

/* ------------------------------------------------------------------ .comp.b ---------- */

/* -proxi (kma.c:702-718): |minFrac| as save_kmers_batch hands it to the get*Proxi* functions (kmers.c:133-150); 1.0 = off */
static double g_proxi = 1.0;
void orc_set_proxi(double f) { g_proxi = f < 0 ? -f : f; }
double orc_get_proxi(void) { return g_proxi; }
/* soft proximity (-proxi < 0 together with -mem_mode, kmers.c:133-153): every template a get*Proxi* function keeps adds
 * its score to softProxi[]; save_kmers_batch appends the sums to the stream (6 ints = the first 24 bytes, then DB_size
 * unsigned longs) and runKMA_MEM takes them for alignment_scores (runkma.c:1153-1156). NULL = off. */
static uint64_t *g_soft = 0;
void orc_set_soft_proxi(uint64_t *sums) { g_soft = sums; }
uint64_t *orc_get_soft_proxi(void) { return g_soft; }

static int rd(FILE *f, void *dst, size_t n) { return fread(dst, 1, n, f) == n ? 0 : -1; }

orc_db *orc_db_open(const char *prefix) {
	char path[4096];
	orc_db *db = calloc(1, sizeof(orc_db));
	FILE *f;
	uint32_t u32[3];
	uint64_t u64[5];

	snprintf(path, sizeof(path), "%s.comp.b", prefix);
	if (!(f = fopen(path, "rb"))) { free(db); return 0; }
	if (rd(f, u32, 12) || rd(f, u64, 40)) { fclose(f); free(db); return 0; }
	db->DB_size = u32[0]; db->mlen = u32[1]; db->prefix_len = u32[2];
	db->prefix = u64[0]; db->size = u64[1]; db->n = u64[2]; db->v_index = u64[3]; db->null_index = u64[4];
	db->kmask = db->mlen >= 32 ? ~0ull : ((1ull << (2 * db->mlen)) - 1);
	db->mega = (db->size - 1) == db->kmask;
	db->exist_wide = db->mega ? (db->v_index > 0xFFFFFFFFull) : (db->n > 0xFFFFFFFFull);
	db->values_short = db->DB_size < 65535;
	db->key_wide = db->mlen > 16;
	db->vidx_wide = !(db->v_index < 0xFFFFFFFFull);

	size_t nb = db->size * (db->exist_wide ? 8 : 4);
	db->exist = malloc(nb);
	if (rd(f, db->exist, nb)) goto fail;
	nb = db->v_index * (db->values_short ? 2 : 4);
	db->values = malloc(nb ? nb : 1);
	if (rd(f, db->values, nb)) goto fail;
	if (!db->mega) {
		nb = (db->n + 1) * (db->key_wide ? 8 : 4);
		db->key_index = malloc(nb);
		if (rd(f, db->key_index, nb)) goto fail;
		nb = db->n * (db->vidx_wide ? 8 : 4);
		db->value_index = malloc(nb ? nb : 1);
		if (rd(f, db->value_index, nb)) goto fail;
	}
	if (rd(f, u32, 8) == 0) { db->kmersize = u32[0]; db->flag = u32[1]; }
	else { db->kmersize = db->mlen; db->flag = 0; }
	fclose(f);
	db->hmask = db->size - 1; /* "make indexing a masking problem" */

	/* template lengths (optional for -1t1, needed for alignment) */
	snprintf(path, sizeof(path), "%s.length.b", prefix);
	if ((f = fopen(path, "rb"))) {
		int32_t n;
		if (rd(f, &n, 4) == 0 && n == db->DB_size) {
			db->lengths = malloc(sizeof(int32_t) * n);
			if (rd(f, db->lengths, sizeof(int32_t) * n)) { free(db->lengths); db->lengths = 0; }
		}
		fclose(f);
	}
	return db;
fail:
	fclose(f);
	orc_db_close(db);
	return 0;
}

void orc_db_close(orc_db *db) {
	if (!db) return;
	free(db->exist); free(db->values); free(db->key_index); free(db->value_index); free(db->lengths);
	free(db->seq); free(db->seq_off);
	free(db);
}

static inline uint64_t get_exist(const orc_db *db, uint64_t i) {
	return db->exist_wide ? ((uint64_t *)db->exist)[i] : ((uint32_t *)db->exist)[i];
}
static inline uint64_t get_key(const orc_db *db, uint64_t i) {
	return db->key_wide ? ((uint64_t *)db->key_index)[i] : ((uint32_t *)db->key_index)[i];
}
static inline uint64_t get_vidx(const orc_db *db, uint64_t i) {
	return db->vidx_wide ? ((uint64_t *)db->value_index)[i] : ((uint32_t *)db->value_index)[i];
}

/* k-mer -> offset of its template list inside values[], or -1. Offsets identify lists: the
 * indexer de-duplicates identical lists (compress.c valuesHash_add), so "same list as the
 * previous hit" (savekmers.c:2522 pointer compare) == "same offset". */
int64_t orc_lookup(const orc_db *db, uint64_t key) {
	if (db->mega) {
		uint64_t v = get_exist(db, key & db->kmask);
		return v != 1 ? (int64_t)v : -1;
	}
	uint64_t bucket = key & db->hmask;
	uint64_t pos = get_exist(db, bucket);
	if (pos == db->null_index) return -1;
	for (uint64_t k = get_key(db, pos); k != key; k = get_key(db, ++pos)) {
		if ((k & db->hmask) != bucket) return -1;
	}
	return (int64_t)get_vidx(db, pos);
}

int orc_list(const orc_db *db, int64_t off, int *n_out, const void **ids) {
	if (db->values_short) {
		const uint16_t *p = (const uint16_t *)db->values + off;
		*n_out = p[0]; *ids = p + 1;
	} else {
		const uint32_t *p = (const uint32_t *)db->values + off;
		*n_out = (int)p[0]; *ids = p + 1;
	}
	return 0;
}
static inline int list_id(const orc_db *db, const void *ids, int i) {
	return db->values_short ? ((const uint16_t *)ids)[i] : (int)((const uint32_t *)ids)[i];
}

/* ------------------------------------------------------------------ 2-bit codec ------ */

static uint64_t rev2(uint64_t w) {
	w = ((w >> 2) & 0x3333333333333333ull) | ((w & 0x3333333333333333ull) << 2);
	w = ((w >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((w & 0x0F0F0F0F0F0F0F0Full) << 4);
	return __builtin_bswap64(w);
}

/* reverse complement of a packed read; N positions mirrored. N bases are stored as A, so they
 * come out as T in the reverse strand words (compdna.c:236-237 complements blindly). */
void orc_revcomp(const uint64_t *seq, int seqlen, const int32_t *N, int nN, uint64_t *rseq, int32_t *rN) {
	int words = (seqlen + 31) >> 5;
	for (int i = 0; i < words; ++i) rseq[words - 1 - i] = rev2(~seq[i]);
	int pad = 2 * (32 * words - seqlen);
	if (pad) {
		for (int i = 0; i + 1 < words; ++i) rseq[i] = (rseq[i] << pad) | (rseq[i + 1] >> (64 - pad));
		rseq[words - 1] <<= pad;
	}
	for (int i = 0; i < nN; ++i) rN[i] = seqlen - 1 - N[nN - 1 - i];
}

static inline uint64_t kmer_at(const uint64_t *seq, int pos, int k) {
	int w = pos >> 5, b = (pos & 31) << 1, sh = 64 - 2 * k;
	uint64_t x = seq[w] << b;
	if (b > sh) x |= seq[w + 1] >> (64 - b);
	return x >> sh;
}

/* ------------------------------------------------------------------ scoring ---------- */

/* score contribution of one hit that continues a template's run after `gaps` missed k-mers.
 * `same_list` selects the accumulate-in-run bookkeeping of savekmers.c:2529-2569 (returns the
 * deltas through acc[4] = {Ms, MMs, Us, W1s}) versus the direct per-template score of
 * savekmers.c:2592-2625. Only mlen == kmersize is in scope (flag == 0 databases). */
static void run_deltas(const orc_params *p, int k, int gaps, int acc[4]) {
	if (gaps == 0) { acc[0] += 1; }
	else if (gaps == k) { acc[0] += k; acc[1] += 1; }
	else if (k < gaps) {
		int g = gaps - (k - 1), mm, m;
		acc[0] += k;
		if (g <= 2) { mm = g; m = 0; }
		else {
			mm = g / k + (g % k ? 1 : 0); if (mm < 2) mm = 2;
			m = g - mm; if (k < m) m = k; if (mm < m) m = mm;
		}
		if (p->W1 + (g - 1) * p->U <= mm * p->MM + m * p->M) { acc[1] += mm; acc[0] += m; }
		else { acc[3] += 1; acc[2] += g - 1; }
	} else { acc[0] += gaps; acc[3] += 1; acc[2] += k - gaps; }
}

static int resume_score(const orc_params *p, int k, int gaps) {
	if (gaps == 0) return p->M;
	if (gaps == k) return k * p->M + p->MM; /* fwd gaps*M+MM == rc k*M+MM when mlen == k */
	if (k < gaps) {
		int g = gaps - (k - 1), mm, m, a, b;
		if (g <= 2) { mm = g; m = 0; }
		else {
			mm = g / k + (g % k ? 1 : 0); if (mm < 2) mm = 2;
			m = g - mm; if (k < m) m = k; if (mm < m) m = mm;
		}
		a = p->W1 + (g - 1) * p->U; b = mm * p->MM + m * p->M;
		return k * p->M + (a <= b ? b : a);
	}
	return gaps * p->M + (k - gaps) * p->U + p->W1;
}

typedef struct {
	int *score;  /* DB_size, zero between reads */
	int *ext;    /* DB_size */
	char *incl;  /* DB_size */
} scratch_t;

/* One strand. cand[0] = count, cand[1..] = templates in first-seen order. Returns best score and
 * leaves the arg-max set (first-seen order) in cand. `*lookups` counts hash probes issued. */
static int scan_strand(const orc_db *db, const orc_params *p, const uint64_t *seq, int seqlen,
                       const int32_t *N, int nN, scratch_t *S, int *cand, orc_stats *st, int *all_scores) {
	const int k = db->kmersize;
	int hit = p->exhaustive;

	/* quick check: every k-th k-mer of every N-free stretch, until the first hit */
	for (int seg = 0, s = 0; seg <= nN && !hit; ++seg) {
		int e = seg < nN ? N[seg] : seqlen;
		for (int j = s; j < e - k + 1 && !hit; j += k) {
			if (st) st->lookups++;
			hit = orc_lookup(db, kmer_at(seq, j, k)) >= 0;
		}
		s = e + 1;
	}
	cand[0] = 0;
	if (!hit) return 0;

	int64_t last = -1;
	int last_pos = 0, nhits = 0;
	int acc[4] = {0, 0, 0, 0};
	for (int seg = 0, s = 0; seg <= nN && s < seqlen - k + 1; ++seg) {
		int e = seg < nN ? N[seg] : seqlen;
		for (int j = s; j + k <= e; ++j) {
			int64_t off = orc_lookup(db, kmer_at(seq, j, k));
			if (st) st->lookups++;
			if (off < 0) continue;
			if (st) st->hits++;
			int nl; const void *ids;
			if (off == last) {
				run_deltas(p, k, j - last_pos - 1, acc);
			} else if (last >= 0) {
				int sc = acc[0] * p->M + acc[1] * p->MM + acc[2] * p->U + acc[3] * p->W1;
				orc_list(db, last, &nl, &ids);
				for (int i = 0; i < nl; ++i) { int t = list_id(db, ids, i); S->score[t] += sc; S->ext[t] = last_pos; }
				orc_list(db, off, &nl, &ids);
				if (st) st->list_fetches++, st->list_ids += nl;
				for (int i = 0; i < nl; ++i) {
					int t = list_id(db, ids, i);
					if (S->incl[t]) S->score[t] += resume_score(p, k, (j - 1) - S->ext[t]);
					else { S->score[t] = k * p->M; S->incl[t] = 1; cand[++cand[0]] = t; }
				}
				acc[0] = acc[1] = acc[2] = acc[3] = 0;
			} else {
				orc_list(db, off, &nl, &ids);
				if (st) st->list_fetches++, st->list_ids += nl;
				for (int i = 0; i < nl; ++i) {
					int t = list_id(db, ids, i);
					S->score[t] = k * p->M; S->incl[t] = 1; cand[i + 1] = t;
				}
				cand[0] = nl;
			}
			last = off; last_pos = j; ++nhits;
		}
		s = e + 1;
	}
	if (last >= 0) {
		int nl; const void *ids;
		int sc = acc[0] * p->M + acc[1] * p->MM + acc[2] * p->U + acc[3] * p->W1;
		orc_list(db, last, &nl, &ids);
		for (int i = 0; i < nl; ++i) S->score[list_id(db, ids, i)] += sc;
	}
	if (all_scores) {   /* get_kmers_for_pair (savekmers.c:654-686): every template seen keeps its clamped score */
		for (int i = 1; i <= cand[0]; ++i) {
			int t = cand[i];
			all_scores[t] = S->score[t] < 0 ? 0 : S->score[t];
			S->score[t] = 0; S->ext[t] = 0; S->incl[t] = 0;
		}
		return nhits;
	}
	if (g_proxi != 1.0) {   /* getProxiMatch (savekmers.c:296-340): every template within minFrac of the best score, raw scores */
		int best = 0, nb = 0;
		for (int i = 1; i <= cand[0]; ++i) if (best < S->score[cand[i]]) best = S->score[cand[i]];
		const int proxi = (int)(g_proxi * best);
		for (int i = 1; i <= cand[0]; ++i) {
			int t = cand[i];
			if (proxi <= S->score[t]) { cand[++nb] = t; if (g_soft) g_soft[t] += S->score[t]; }
			S->score[t] = 0; S->ext[t] = 0; S->incl[t] = 0;
		}
		cand[0] = nhits ? nb : 0;
		return nhits ? best : 0;
	}
	/* arg-max set, first-seen order; scratch returned to zero (getBestMatch, savekmers.c:273) */
	int best = 0, nb = 0;
	for (int i = 1; i <= cand[0]; ++i) {
		int t = cand[i], sc = S->score[t] < 0 ? 0 : S->score[t];
		if (sc > best) { best = sc; nb = 1; cand[1] = t; }
		else if (sc == best) cand[++nb] = t;
		S->score[t] = 0; S->ext[t] = 0; S->incl[t] = 0;
	}
	cand[0] = nhits ? nb : 0;
	return nhits ? best : 0;
}

/* ------------------------------------------------------------------ records ---------- */

static size_t emit_record(uint8_t *out, const uint64_t *seq, int seqlen, const int32_t *N, int nN,
                          int score, const int *tmpl, int ntmpl, const uint8_t *hdr, int hdrlen, int flag) {
	int32_t h[7] = {seqlen, (seqlen + 31) >> 5, nN, score, ntmpl, hdrlen, flag};
	uint8_t *o = out;
	memcpy(o, h, 28); o += 28;
	memcpy(o, seq, 8 * (size_t)h[1]); o += 8 * (size_t)h[1];
	memcpy(o, N, 4 * (size_t)nN); o += 4 * (size_t)nN;
	memcpy(o, tmpl, 4 * (size_t)ntmpl); o += 4 * (size_t)ntmpl;
	memcpy(o, hdr, hdrlen); o += hdrlen;
	return o - out;
}


/* ------------------------------------------------------------------ paired end (-apm p) */

/* One mate as save_kmers_penaltyPair sees it: both strand forms plus which one the in-place comp_rc calls of the
 * reference currently leave in the buffer (get_kmers_for_pair reverse-complements its read, savekmers.c:471). */
typedef struct {
	uint64_t *w[2]; int32_t *N[2];
	int seqlen, words, nN, cur;
	const uint8_t *hdr; int hdrlen;
} mate_t;

typedef struct {
	scratch_t S;
	int *Score, *Score_r;        /* per-template scores of the mate scanned last (forward / reverse strand) */
	int *bt, *bt_r;              /* bestTemplates / bestTemplates_r: templates seen per strand, [0] = count */
	int *rt, *rs;                /* regionTemplates / regionScores of the first mate */
} pair_ws;

/* get_kmers_for_pair (savekmers.c:427-688) */
static int pair_kmers(const orc_db *db, const orc_params *p, mate_t *m, pair_ws *ws, orc_stats *st) {
	ws->bt[0] = 0; ws->bt_r[0] = 0;
	if (m->seqlen < (int)db->kmersize) return 0;
	int hf = scan_strand(db, p, m->w[0], m->seqlen, m->N[0], m->nN, &ws->S, ws->bt, st, ws->Score);
	int hr = scan_strand(db, p, m->w[1], m->seqlen, m->N[1], m->nN, &ws->S, ws->bt_r, st, ws->Score_r);
	m->cur = 1;
	return hf < hr ? hr : hf;
}

static size_t emit_mate(uint8_t *out, const mate_t *m, int score, const int *tmpl, int ntmpl, int flag) {
	return emit_record(out, m->w[m->cur], m->seqlen, m->N[m->cur], m->nN, score, tmpl, ntmpl, m->hdr, m->hdrlen, flag);
}

/* getFirstPen (savekmers.c:1383) */
static int first_pen(pair_ws *ws) {
	int best = 0, n = 0;
	for (int i = 1; i <= ws->bt[0]; ++i) {
		int t = ws->bt[i], sc = ws->Score[t];
		if (best < sc) best = sc;
		++n; ws->rt[n] = t; ws->rs[n] = sc; ws->Score[t] = 0;
	}
	for (int i = 1; i <= ws->bt_r[0]; ++i) {
		int t = ws->bt_r[i], sc = ws->Score_r[t];
		if (best < sc) best = sc;
		++n; ws->rt[n] = -t; ws->rs[n] = sc; ws->Score_r[t] = 0;
	}
	ws->rt[0] = n;
	return best;
}

/* getSecondBestPen (savekmers.c:1415) */
static int second_pen(pair_ws *ws, int bestScore, int PE) {
	int best_r = 0, n;
	for (int i = 1; i <= ws->bt[0]; ++i) if (best_r < ws->Score[ws->bt[i]]) best_r = ws->Score[ws->bt[i]];
	n = ws->bt[0];
	for (int i = 1; i <= ws->bt_r[0]; ++i) {
		if (best_r < ws->Score_r[ws->bt_r[i]]) best_r = ws->Score_r[ws->bt_r[i]];
		ws->bt[++n] = -ws->bt_r[i];
	}
	ws->bt[0] = n;
	int hits = 0;
	if (best_r) {
		int comp = bestScore + best_r - PE;
		if (comp < 0) comp = 0;
		for (int i = 1; i <= ws->rt[0]; ++i) {
			int t = ws->rt[i], sc = 0 < t ? ws->Score_r[t] : ws->Score[-t];   /* the mate must hit the opposite strand */
			if (0 < sc) {
				sc += ws->rs[i];
				if (comp < sc) { comp = sc; hits = 1; ws->rt[hits] = t; }
				else if (comp == sc) ws->rt[++hits] = t;
			}
		}
	}
	if (hits) {   /* proper pair */
		ws->rt[0] = -hits;
		for (int i = ws->bt[0]; i != 0; --i) { if (0 < ws->bt[i]) ws->Score[ws->bt[i]] = 0; else ws->Score_r[-ws->bt[i]] = 0; }
	} else {      /* best hits of each mate on its own */
		for (int i = 1; i <= ws->rt[0]; ++i) if (bestScore == ws->rs[i]) ws->rt[++hits] = ws->rt[i];
		ws->rt[0] = hits;
		hits = 0;
		for (int i = 1; i <= ws->bt[0]; ++i) {
			int t = ws->bt[i];
			if (0 < t) { if (best_r == ws->Score[t]) ws->bt[++hits] = t; ws->Score[t] = 0; }
			else { if (best_r <= ws->Score_r[-t]) ws->bt[++hits] = t; ws->Score_r[-t] = 0; }
		}
		ws->bt[0] = hits;
	}
	return best_r;
}

/* getF_Best (savekmers.c:1648) */
static int f_best(pair_ws *ws) {
	int best = 0, hits = 0;
	for (int i = 1; i <= ws->bt[0]; ++i) {
		int t = ws->bt[i], sc = ws->Score[t];
		if (best < sc) { best = sc; hits = 1; ws->rt[hits] = t; }
		else if (best == sc) ws->rt[++hits] = t;
		ws->Score[t] = 0;
	}
	for (int i = 1; i <= ws->bt_r[0]; ++i) {
		int t = ws->bt_r[i], sc = ws->Score_r[t];
		if (best < sc) { best = sc; hits = 1; ws->rt[hits] = -t; }
		else if (best == sc) ws->rt[++hits] = -t;
		ws->Score_r[t] = 0;
	}
	ws->rt[0] = hits;
	return best;
}

/* getR_Best (savekmers.c:1682-1762): the second mate's arg-max set goes to bt (reverse-strand ids negated); the
 * templates of the first mate's set that the second mate also has among ITS best on the opposite strand are swapped to
 * the front of rt, and rt[0] goes negative to mark the pair (union). Only the best templates still carry a score when
 * the union is checked: the others were zeroed while the maximum was searched. */
static int r_best(pair_ws *ws) {
	int best_r = 0, hits = 0, sc;
	for (int i = 1; i <= ws->bt[0]; ++i) {
		if (best_r < (sc = ws->Score[ws->bt[i]])) {
			for (int j = hits; j != 0; --j) ws->Score[ws->bt[j]] = 0;
			best_r = sc; hits = 1; ws->bt[hits] = ws->bt[i];
		} else if (best_r == sc) ws->bt[++hits] = ws->bt[i];
		else ws->Score[ws->bt[i]] = 0;
	}
	for (int i = 1; i <= ws->bt_r[0]; ++i) {
		if (best_r < (sc = ws->Score_r[ws->bt_r[i]])) {
			for (int j = hits; j != 0; --j) { if (0 < ws->bt[j]) ws->Score[ws->bt[j]] = 0; else ws->Score_r[-ws->bt[j]] = 0; }
			best_r = sc; hits = 1; ws->bt[hits] = -ws->bt_r[i];
		} else if (best_r == sc) ws->bt[++hits] = -ws->bt_r[i];
		else ws->Score_r[ws->bt_r[i]] = 0;
	}
	ws->bt[0] = hits;
	hits = 0;
	for (int i = 1; i <= ws->rt[0]; ++i) {
		const int t = ws->rt[i];
		if (0 < t ? ws->Score_r[t] : ws->Score[-t]) {
			++hits;
			const int tmp = ws->rt[hits]; ws->rt[hits] = ws->rt[i]; ws->rt[i] = tmp;
		}
	}
	if (hits) ws->rt[0] = -hits;
	for (int i = ws->bt[0]; i != 0; --i) { if (0 < ws->bt[i]) ws->Score[ws->bt[i]] = 0; else ws->Score_r[-ws->bt[i]] = 0; }
	return best_r;
}

/* getSecondProxiPen (savekmers.c:1514-1646): as getSecondBestPen, but a union within minFrac of the best union survives,
 * and without a union every template within minFrac of a mate's own best */
static int second_proxi_pen(pair_ws *ws, int bestScore, int PE) {
	int best_r = 0, n;
	for (int i = 1; i <= ws->bt[0]; ++i) if (best_r < ws->Score[ws->bt[i]]) best_r = ws->Score[ws->bt[i]];
	n = ws->bt[0];
	for (int i = 1; i <= ws->bt_r[0]; ++i) {
		if (best_r < ws->Score_r[ws->bt_r[i]]) best_r = ws->Score_r[ws->bt_r[i]];
		ws->bt[++n] = -ws->bt_r[i];
	}
	ws->bt[0] = n;
	int hits = 0;
	if (best_r) {
		int comp = 0;
		for (int i = 1; i <= ws->rt[0]; ++i) {
			int t = ws->rt[i], sc = 0 < t ? ws->Score_r[t] : ws->Score[-t];
			if (0 < sc) { sc += ws->rs[i]; if (comp < sc) comp = sc; }
		}
		if ((bestScore + best_r - PE) <= comp) {
			const int proxi = (int)(g_proxi * comp);
			for (int i = 1; i <= ws->rt[0]; ++i) {
				int t = ws->rt[i], sc = 0 < t ? ws->Score_r[t] : ws->Score[-t];
				if (0 < sc) { sc += ws->rs[i]; if (proxi <= sc) { ws->rt[++hits] = t; if (g_soft) g_soft[abs(t)] += sc; } }
			}
		}
	}
	if (hits) {
		ws->rt[0] = -hits;
		for (int i = ws->bt[0]; i != 0; --i) { if (0 < ws->bt[i]) ws->Score[ws->bt[i]] = 0; else ws->Score_r[-ws->bt[i]] = 0; }
	} else {
		int proxi = (int)(g_proxi * bestScore);
		for (int i = 1; i <= ws->rt[0]; ++i) if (proxi <= ws->rs[i]) ws->rt[++hits] = ws->rt[i];
		ws->rt[0] = hits;
		hits = 0;
		proxi = (int)(g_proxi * best_r);
		for (int i = 1; i <= ws->bt[0]; ++i) {
			int t = ws->bt[i];
			if (0 < t) { if (proxi <= ws->Score[t]) { ws->bt[++hits] = t; if (g_soft) g_soft[t] += ws->Score[t]; } ws->Score[t] = 0; }
			else { if (proxi <= ws->Score_r[-t]) { ws->bt[++hits] = t; if (g_soft) g_soft[-t] += ws->Score_r[-t]; } ws->Score_r[-t] = 0; }
		}
		ws->bt[0] = hits;
	}
	return best_r;
}

/* getF_Proxi (savekmers.c:1764) */
static int f_proxi(pair_ws *ws) {
	int best = 0, hits = 0;
	for (int i = 1; i <= ws->bt[0]; ++i) if (best < ws->Score[ws->bt[i]]) best = ws->Score[ws->bt[i]];
	for (int i = 1; i <= ws->bt_r[0]; ++i) if (best < ws->Score_r[ws->bt_r[i]]) best = ws->Score_r[ws->bt_r[i]];
	const int proxi = (int)(g_proxi * best);
	for (int i = 1; i <= ws->bt[0]; ++i) {
		int t = ws->bt[i];
		if (proxi <= ws->Score[t]) { ws->rt[++hits] = t; if (g_soft) g_soft[t] += ws->Score[t]; }
		ws->Score[t] = 0;
	}
	for (int i = 1; i <= ws->bt_r[0]; ++i) {
		int t = ws->bt_r[i];
		if (proxi <= ws->Score_r[t]) { ws->rt[++hits] = -t; if (g_soft) g_soft[t] += ws->Score_r[t]; }
		ws->Score_r[t] = 0;
	}
	ws->rt[0] = hits;
	return best;
}

/* getR_Proxi (savekmers.c:1825): the templates of the first mate's set that are within the second mate's proximity on
 * the opposite strand (they alone still carry a score) are swapped to the front of rt */
static int r_proxi(pair_ws *ws) {
	int best = 0, hits = 0;
	const int nf = ws->bt[0];
	for (int i = 1; i <= nf; ++i) if (best < ws->Score[ws->bt[i]]) best = ws->Score[ws->bt[i]];
	for (int i = 1; i <= ws->bt_r[0]; ++i) if (best < ws->Score_r[ws->bt_r[i]]) best = ws->Score_r[ws->bt_r[i]];
	const int proxi = (int)(g_proxi * best);
	for (int i = 1; i <= nf; ++i) {
		int t = ws->bt[i];
		if (proxi <= ws->Score[t]) { ws->bt[++hits] = t; if (g_soft) g_soft[t] += ws->Score[t]; } else ws->Score[t] = 0;
	}
	for (int i = 1; i <= ws->bt_r[0]; ++i) {
		int t = ws->bt_r[i];
		if (proxi <= ws->Score_r[t]) { ws->bt[++hits] = -t; if (g_soft) g_soft[t] += ws->Score_r[t]; } else ws->Score_r[t] = 0;
	}
	ws->bt[0] = hits;
	hits = 0;
	for (int i = 1; i <= ws->rt[0]; ++i) {
		const int t = ws->rt[i];
		if (0 < t ? ws->Score_r[t] : ws->Score[-t]) {
			++hits;
			const int tmp = ws->rt[hits]; ws->rt[hits] = ws->rt[i]; ws->rt[i] = tmp;
		}
	}
	if (hits) ws->rt[0] = -hits;
	for (int i = ws->bt[0]; i != 0; --i) { if (0 < ws->bt[i]) ws->Score[ws->bt[i]] = 0; else ws->Score_r[-ws->bt[i]] = 0; }
	return best;
}

/* save_kmers_unionPair (savekmers.c:3367-3570, the default pairing) with getF = getF_Best, getR = getR_Best, rev = 1 */
static size_t seed_pair_union(const orc_db *db, const orc_params *p, mate_t *m1, mate_t *m2, pair_ws *ws, uint8_t *out, orc_stats *st) {
	const int k = db->kmersize;
	size_t op = 0;
	int best = 0, best_r = 0, flag = 65, flag_r = 129;
	int *rt = ws->rt, *bt = ws->bt;
	if (pair_kmers(db, p, m1, ws, st)) {
		best = g_proxi != 1.0 ? f_proxi(ws) : f_best(ws);
		if (k < best && best * k < (m1->seqlen - best)) best = 0;
	}
	if (pair_kmers(db, p, m2, ws, st)) {
		best_r = g_proxi != 1.0 ? (best ? r_proxi(ws) : f_proxi(ws)) : (best ? r_best(ws) : f_best(ws));
		if (k < best_r && best_r * k < (m2->seqlen - best_r)) { best_r = 0; rt[0] = abs(rt[0]); }
	} else {   /* the lists are emptied: a first mate that kept its score still has its set in rt */
		bt[0] = 0; ws->bt_r[0] = 0;
	}
	if (0 < best && 0 < best_r) {
		if (rt[0] < 0) {   /* union found: a pair */
			flag |= 2; flag_r |= 2;
			rt[0] = -rt[0];
			if (0 < rt[1]) {
				flag |= 32; flag_r |= 16;
				m1->cur ^= 1;
				op += emit_mate(out + op, m1, best, rt + 1, 0, flag);
				op += emit_mate(out + op, m2, best_r, rt + 1, rt[0], flag_r);
			} else {
				flag |= 16; flag_r |= 32;
				m2->cur ^= 1;
				for (int i = rt[0]; i != 0; --i) rt[i] = -rt[i];
				op += emit_mate(out + op, m2, best_r, rt + 1, 0, flag_r);
				op += emit_mate(out + op, m1, best, rt + 1, rt[0], flag);
			}
		} else {           /* two single mates */
			if (0 < rt[1]) { m1->cur ^= 1; if (rt[rt[0]] < 0) best = -best; }
			else { flag |= 16; flag_r |= 32; for (int i = 1; i <= rt[0]; ++i) rt[i] = -rt[i]; }
			if (0 < bt[1]) { m2->cur ^= 1; if (bt[bt[0]] < 0) best_r = -best_r; }
			else { flag |= 32; flag_r |= 16; for (int i = 1; i <= bt[0]; ++i) bt[i] = -bt[i]; }
			op += emit_mate(out + op, m1, best, rt + 1, rt[0], flag);
			op += emit_mate(out + op, m2, best_r, bt + 1, bt[0], flag_r);
		}
	} else if (best) {
		flag |= 8 | 32;
		if (0 < rt[1]) { m1->cur ^= 1; if (rt[rt[0]] < 0) best = -best; }
		else { flag |= 16; for (int i = 1; i <= rt[0]; ++i) rt[i] = -rt[i]; }
		op += emit_mate(out + op, m1, best, rt + 1, rt[0], flag);
	} else if (best_r) {
		flag_r |= 8 | 32;
		if (0 < rt[1]) { m2->cur ^= 1; if (rt[rt[0]] < 0) best_r = -best_r; }
		else { flag_r |= 16; for (int i = 1; i <= rt[0]; ++i) rt[i] = -rt[i]; }
		op += emit_mate(out + op, m2, best_r, rt + 1, rt[0], flag_r);
	}
	return op;
}

static int imin(int a, int b) { return a < b ? a : b; }

/* save_kmers_penaltyPair (savekmers.c:3572-3777) with printPtr = print_ankers, printPairPtr = printPair,
 * deConPrintPtr = printPtr, rev = 1 (no prefix). Appends 0, 1 or 2 stage-2 records to out; returns bytes written. */
static size_t seed_pair(const orc_db *db, const orc_params *p, mate_t *m1, mate_t *m2, pair_ws *ws, uint8_t *out, orc_stats *st) {
	const int k = db->kmersize;
	size_t op = 0;
	int hc, hc_r, best = 0, best_r = 0, flag = 65, flag_r = 129;
	if ((hc = pair_kmers(db, p, m1, ws, st))) best = first_pen(ws);
	if ((hc_r = pair_kmers(db, p, m2, ws, st))) {
		if (g_proxi != 1.0) best_r = 0 < best ? second_proxi_pen(ws, best, p->PE) : f_proxi(ws);
		else best_r = 0 < best ? second_pen(ws, best, p->PE) : f_best(ws);
	}
	int *rt = ws->rt, *bt = ws->bt;
	if (0 < best && 0 < best_r) {
		if (rt[0] < 0) {   /* proper pair */
			flag |= 2; flag_r |= 2;
			int comp = imin(hc + hc_r, best + best_r);
			if (k <= comp || (m1->seqlen + m2->seqlen - comp - (k << 1)) < comp * k) {
				rt[0] = -rt[0];
				if (0 < rt[1]) {
					flag |= 32; flag_r |= 16;
					m1->cur ^= 1;
					op += emit_mate(out + op, m1, best, rt + 1, 0, flag);          /* printPair: first record has no templates */
					op += emit_mate(out + op, m2, best_r, rt + 1, rt[0], flag_r);
				} else {
					flag |= 16; flag_r |= 32;
					m2->cur ^= 1;
					for (int i = rt[0]; i != 0; --i) rt[i] = -rt[i];
					op += emit_mate(out + op, m2, best_r, rt + 1, 0, flag_r);
					op += emit_mate(out + op, m1, best, rt + 1, rt[0], flag);
				}
			}
		} else {           /* two single mates */
			int h = imin(hc, best), h_r = imin(hc_r, best_r);
			h = k <= h || (m1->seqlen - h - k) < h * k;
			if (h) {
				if (0 < rt[1]) { m1->cur ^= 1; if (rt[rt[0]] < 0) best = -best; }
				else { flag |= 16; flag_r |= 32; for (int i = rt[0]; i != 0; --i) rt[i] = -rt[i]; }
			}
			h_r = k <= h_r || (m2->seqlen - h_r - k) < h_r * k;
			if (h_r) {
				if (0 < bt[1]) { m2->cur ^= 1; if (bt[bt[0]] < 0) best_r = -best_r; }
				else { flag |= 32; flag_r |= 16; for (int i = bt[0]; i != 0; --i) bt[i] = -bt[i]; }
			}
			if (h) op += emit_mate(out + op, m1, best, rt + 1, rt[0], flag);
			if (h_r) op += emit_mate(out + op, m2, best_r, bt + 1, bt[0], flag_r);
		}
	} else if (0 < best) {
		int h = imin(hc, best);
		if (k <= h || (m1->seqlen - h - k) < h * k) {
			flag |= 8 | 32;
			if (0 < rt[1]) { m1->cur ^= 1; if (rt[rt[0]] < 0) best = -best; }
			else { flag |= 16; for (int i = rt[0]; i != 0; --i) rt[i] = -rt[i]; }
			op += emit_mate(out + op, m1, best, rt + 1, rt[0], flag);
		}
	} else if (0 < best_r) {
		int h = imin(hc_r, best_r);
		if (k <= h || (m2->seqlen - h - k) < h * k) {
			flag_r |= 8 | 32;
			if (0 < rt[1]) { m2->cur ^= 1; if (rt[rt[0]] < 0) best_r = -best_r; }
			else { flag_r |= 16; for (int i = 1; i <= rt[0]; ++i) rt[i] = -rt[i]; }
			op += emit_mate(out + op, m2, best_r, rt + 1, rt[0], flag_r);
		}
	}
	return op;
}

/* one stage-1 record (loadFsa, savekmers.c:50-92) -> both strand forms of the read. Returns the signed header
 * length field (< 0: first mate of a pair), 0 at the end of the stream. */
typedef struct { uint64_t *w[2]; int32_t *N[2]; size_t wcap, ncap; } mate_buf;

static int load_mate(const uint8_t *in, size_t in_bytes, size_t *ip, mate_buf *b, mate_t *m) {
	if (*ip + 16 > in_bytes) return 0;
	int32_t h[4]; memcpy(h, in + *ip, 16);
	if (h[0] < 0) return 0;
	*ip += 16;
	const int seqlen = h[0], words = h[1], nN = h[2];
	if ((size_t)words + 2 > b->wcap) {
		b->wcap = 2 * (size_t)words + 2;
		for (int i = 0; i < 2; ++i) b->w[i] = realloc(b->w[i], 8 * b->wcap);
	}
	if ((size_t)nN + 2 > b->ncap) {
		b->ncap = 2 * (size_t)nN + 2;
		for (int i = 0; i < 2; ++i) b->N[i] = realloc(b->N[i], 4 * b->ncap);
	}
	memcpy(b->w[0], in + *ip, 8 * (size_t)words); b->w[0][words] = 0; *ip += 8 * (size_t)words;
	memcpy(b->N[0], in + *ip, 4 * (size_t)nN); *ip += 4 * (size_t)nN;
	m->hdr = in + *ip; m->hdrlen = abs(h[3]); *ip += m->hdrlen;
	m->seqlen = seqlen; m->words = words; m->nN = nN; m->cur = 0;
	for (int i = 0; i < 2; ++i) { m->w[i] = b->w[i]; m->N[i] = b->N[i]; }
	orc_revcomp(b->w[0], seqlen, b->N[0], nN, b->w[1], b->N[1]); b->w[1][words] = 0;
	return h[3] ? h[3] : 1;
}

/* Whole stage 2: stage-1 stream in, stage-2 stream out (including the terminator). Single-end records go through
 * save_kmers (-1t1), pairs (first mate written with a negative header length, runinput.c:789) through
 * save_kmers_penaltyPair (-apm p). Returns bytes written, or -1 if `cap` is too small. */
int64_t orc_seed_stream(const orc_db *db, const orc_params *p, const uint8_t *in, size_t in_bytes,
                        uint8_t *out, size_t cap, orc_stats *st) {
	const int k = db->kmersize;
	const size_t D = (size_t)db->DB_size + 1;
	pair_ws ws;
	ws.S.score = calloc(D, sizeof(int)); ws.S.ext = calloc(D, sizeof(int)); ws.S.incl = calloc(D, 1);
	ws.Score = calloc(D, sizeof(int)); ws.Score_r = calloc(D, sizeof(int));
	ws.bt = malloc(sizeof(int) * (2 * D + 4)); ws.bt_r = malloc(sizeof(int) * (2 * D + 4));
	ws.rt = malloc(sizeof(int) * (2 * D + 4)); ws.rs = malloc(sizeof(int) * (2 * D + 4));
	int *cf = ws.bt, *cr = ws.bt_r;
	mate_buf b1, b2; memset(&b1, 0, sizeof(b1)); memset(&b2, 0, sizeof(b2));
	mate_t m1, m2;
	size_t ip = 0, op = 0;
	int32_t nreads = 0;
	int64_t ret = 0;
	int go;

	while ((go = load_mate(in, in_bytes, &ip, &b1, &m1)) != 0) {
		++nreads;
		if (st) st->reads++, st->read_words += m1.words;
		if (go < 0) {   /* paired end */
			if (!load_mate(in, in_bytes, &ip, &b2, &m2)) break;
			if (st) st->reads++, st->read_words += m2.words;
			size_t need = 2 * 28 + 8 * (size_t)(m1.words + m2.words) + 4 * (size_t)(m1.nN + m2.nN) + 8 * D + m1.hdrlen + m2.hdrlen;
			if (op + need + 4 > cap) { ret = -1; goto done; }
			size_t w = p->apm == 1 ? seed_pair_union(db, p, &m1, &m2, &ws, out + op, st) : seed_pair(db, p, &m1, &m2, &ws, out + op, st);
			if (w && st) st->mapped++;
			op += w;
			continue;
		}
		const int seqlen = m1.seqlen, words = m1.words, nN = m1.nN, hdrlen = m1.hdrlen;
		if (seqlen < k) continue;
		int bf = scan_strand(db, p, m1.w[0], seqlen, m1.N[0], nN, &ws.S, cf, st, 0);
		int br = scan_strand(db, p, m1.w[1], seqlen, m1.N[1], nN, &ws.S, cr, st, 0);
		if ((bf > 0 || br > 0) && (k <= bf || k <= br)) {
			size_t need = 28 + 8 * (size_t)words + 4 * (size_t)nN + 4 * (size_t)(cf[0] + cr[0]) + hdrlen;
			if (op + need + 4 > cap) { ret = -1; goto done; }
			if (bf > br) op += emit_record(out + op, m1.w[0], seqlen, m1.N[0], nN, bf, cf + 1, cf[0], m1.hdr, hdrlen, 0);
			else if (bf < br) op += emit_record(out + op, m1.w[1], seqlen, m1.N[1], nN, br, cr + 1, cr[0], m1.hdr, hdrlen, 16);
			else {
				for (int i = 1; i <= cr[0]; ++i) cf[++cf[0]] = -cr[i];
				op += emit_record(out + op, m1.w[0], seqlen, m1.N[0], nN, -bf, cf + 1, cf[0], m1.hdr, hdrlen, 0);
			}
			if (st) st->mapped++;
		}
	}
	if (op + 4 > cap) { ret = -1; goto done; }
	nreads = -nreads; memcpy(out + op, &nreads, 4); op += 4;
	ret = (int64_t)op;
done:
	free(ws.S.score); free(ws.S.ext); free(ws.S.incl); free(ws.Score); free(ws.Score_r);
	free(ws.bt); free(ws.bt_r); free(ws.rt); free(ws.rs);
	for (int i = 0; i < 2; ++i) { free(b1.w[i]); free(b1.N[i]); free(b2.w[i]); free(b2.N[i]); }
	return ret;
}

/* default CLI scoring (kma.c:327-336, 1308-1328): M=1, MM=-2, U=-1, W1=-3, Wl=-6, Mn=0, PE=7 */
void orc_default_params(orc_params *p) {
	memset(p, 0, sizeof(*p));
	p->M = 1; p->MM = -2; p->U = -1; p->W1 = -3; p->Wl = -6; p->Mn = 0; p->PE = 7;
	for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) p->d[i * 5 + j] = i == j ? 1 : -2;
	for (int i = 0; i < 5; ++i) p->d[4 * 5 + i] = p->d[i * 5 + 4] = 0;
}
