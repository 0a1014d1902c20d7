/* TEST INFRASTRUCTURE ONLY -- CPU oracle for ConClave's choice pass and the per-template bucketing that follows it.
 *
 * A plain-C restatement of runConClave (conclave.c:43-213, the default -ConClave 1) and printFrags (frags.c:30-61):
 * per frag_raw record (updatescores.c:284-295) the template with the largest global alignment score wins, ties by
 * score per template base (double), then by unique score, then by the smaller template id; reads chosen on the
 * reverse strand are reverse-complemented (strrc, stdnuc.c:450) and their query bounds mirrored; the weighted scores
 * and read / fragment counts are summed per template; the records leave in template order, inside a template in
 * REVERSE arrival order (the reference prepends to a linked list), the mate of a pair before its first read.
 * The int truncations of the reference (best_read_score and bestNum are ints compared with the 64-bit sums,
 * conclave.c:45-46, 92-94) are kept by using the same C types.
 * One call = one "file" of the reference (it cuts a new one every maxFrag records, conclave.c:196-207).
 * orc_conclave2_stream restates runConClave2 / runConClave2_lc (conclave.c:386-747, 749-1110; -ConClave 2): a first choice pass
 * sums provisional w_scores, templates whose sum is not significant (chi-square of observed vs expected hits, and / or
 * the -mrs depth test, stdstat.c:23-31) lose it, reads with exactly one significant candidate add their score to that
 * template's unique score, and the final choice draws a candidate with probability proportional to the unique scores
 * (a Lehmer generator seeded from the read's first and last 7 bases) before falling back to the 4-key order.
 * Pinned to the reference's own functions run by ref_harness -conclave / -conclave2 (tests/test_oracle_conclave.py). */
#include "orc.h"
#include <stdlib.h>
#include <string.h>
#include <limits.h>

typedef struct { int tmpl, buf[7]; const uint8_t *q, *hdr; int rc; int b0, b1, has_bound; } cfrag;

static int orc_lc = 0;
void orc_conclave_set_lc(int lc) { orc_lc = lc; }   /* 1: runConClave_lc */

static int orc_cc_version = 1;   /* 2: the final choice of runConClave2 (set by orc_conclave2_stream around its last pass) */

/* the 4-key order over a record's candidates (conclave.c:66-113 / 414-447); first = value bestTemplate starts from */
static int cc_four_keys(const int32_t *template_lengths, const uint64_t *alignment_scores, const uint64_t *uniq_alignment_scores,
                        const uint8_t *S, const uint8_t *E, const uint8_t *T, int bestHits, int first, int *start, int *end) {
	int32_t v;
#define AT(arr, i) (memcpy(&v, (const uint8_t *)(arr) + 4 * (size_t)(i), 4), v)
	double bestScore = 0, tmp_score;
	int best_read_score = 0, bestNum = 0, bestTemplate = first;
	for (int i = 0; i != bestHits; ++i) {
		const int tt = AT(T, i), tmp_start = AT(S, i), tmp_end = AT(E, i);
		const int t = tt < 0 ? -tt : tt;
		tmp_score = 1.0 * alignment_scores[t] / template_lengths[t];
		int take = 0;
		if (orc_lc) {
			if (tmp_score > bestScore) take = 1;
			else if (tmp_score == bestScore) {
				if (alignment_scores[t] > best_read_score) take = 1;
				else if (alignment_scores[t] == best_read_score) {
					if (uniq_alignment_scores[t] > bestNum) take = 1;
					else if (uniq_alignment_scores[t] == bestNum && t < abs(bestTemplate)) take = 1;
				}
			}
		} else if (alignment_scores[t] > best_read_score) take = 1;
		else if (alignment_scores[t] == best_read_score) {
			if (tmp_score > bestScore) take = 1;
			else if (tmp_score == bestScore) {
				if (uniq_alignment_scores[t] > bestNum) take = 1;
				else if (uniq_alignment_scores[t] == bestNum && t < abs(bestTemplate)) take = 1;
			}
		}
		if (take) {
			bestTemplate = tt; best_read_score = alignment_scores[t]; bestScore = tmp_score;
			bestNum = uniq_alignment_scores[t]; *start = tmp_start; *end = tmp_end;
		}
	}
	return bestTemplate;
#undef AT
}

/* the final choice of runConClave2 for a record with bestHits != 1 (conclave.c:547-655) */
static int cc2_choice(const int32_t *template_lengths, const uint64_t *alignment_scores, const uint64_t *uniq_alignment_scores,
                      const uint8_t *q, int q_len, const uint8_t *S, const uint8_t *E, const uint8_t *T, int bestHits, int *start, int *end) {
	int32_t v;
#define AT(arr, i) (memcpy(&v, (const uint8_t *)(arr) + 4 * (size_t)(i), 4), v)
	int bestTemplate = 0, tot = 0, i, j, rnd;
	*start = 0; *end = 0;
	for (i = bestHits; i--;) tot += uniq_alignment_scores[abs(AT(T, i))];
	if (tot && 16 <= q_len) {
		rnd = q[0]; i = -1; j = q_len;
		while (++i < 7) rnd = (((rnd << 2) | q[i]) << 2) | q[--j];
		rnd = 16807 * (rnd % 127773) - 2836 * (rnd / 127773);   /* minimal standard */
		if (rnd <= 0) rnd += 0x7fffffff;
		double tmp_score = rnd;
		tmp_score /= INT_MAX;
		const unsigned randScore = tmp_score * tot;
		uint64_t score = 0;
		for (i = 0; i != bestHits; ++i) {
			score += uniq_alignment_scores[abs(AT(T, i))];
			if (randScore < score) { bestTemplate = AT(T, i); *start = AT(S, i); *end = AT(E, i); break; }
		}
		if (bestTemplate == 0) tot = 0;
	} else tot = 0;
	if (tot == 0) bestTemplate = cc_four_keys(template_lengths, alignment_scores, uniq_alignment_scores, S, E, T, bestHits, 0, start, end);
	return bestTemplate;
#undef AT
}

int64_t orc_conclave_stream(const int32_t *template_lengths, int DB_size, const uint8_t *frag, size_t fb,
                            const uint64_t *alignment_scores, const uint64_t *uniq_alignment_scores,
                            uint8_t *out, size_t cap, uint64_t *w_scores, uint32_t *fragmentCounts, uint32_t *readCounts) {
	static const unsigned char comp[6] = {3, 2, 1, 0, 4, 5};
	size_t ip = 0, n = 0, acap = 1024;
	cfrag *F = malloc(acap * sizeof(cfrag));
	while (ip + 20 <= fb) {
		int stats[5];
		memcpy(stats, frag + ip, 20);
		if (stats[0] == 0) break;
		ip += 20;
		const int q_len = stats[0], sparse = stats[1], bestHits = abs(sparse), read_score = abs(stats[2]), hl = stats[3];
		int flag = stats[4];
		const uint8_t *q = frag + ip, *hdr = q + q_len;
		const int32_t *bs = (const int32_t *)(hdr + hl);   /* unaligned: read through memcpy below */
		ip += (size_t)q_len + hl + 12 * (size_t)bestHits;
		int bestTemplate, start, end;
		int32_t v;
#define AT(arr, i) (memcpy(&v, (const uint8_t *)(arr) + 4 * (size_t)(i), 4), v)
		const uint8_t *S = (const uint8_t *)bs, *E = S + 4 * (size_t)bestHits, *T = E + 4 * (size_t)bestHits;
		if (orc_cc_version == 2 && bestHits != 1) {
			bestTemplate = cc2_choice(template_lengths, alignment_scores, uniq_alignment_scores, q, q_len, S, E, T, bestHits, &start, &end);
			if (bestTemplate == 0) {   /* no candidate left: the record (and its mate block) is skipped (conclave.c:722-725) */
				if (stats[2] < 0) {
					int m[3];
					if (ip + 12 > fb) { free(F); return -3; }
					memcpy(m, frag + ip, 12); ip += 12 + (size_t)m[0] + m[1];
				}
				continue;
			}
		} else if (bestHits > 1) {
			double bestScore = 0, tmp_score;
			int best_read_score = 0, bestNum = 0;
			bestTemplate = -1; start = 0; end = 0;
			for (int i = 0; i != bestHits; ++i) {
				const int tt = AT(T, i), tmp_start = AT(S, i), tmp_end = AT(E, i);
				const int t = tt < 0 ? -tt : tt;
				tmp_score = 1.0 * alignment_scores[t] / template_lengths[t];
				int take = 0;
				if (orc_lc) {   /* runConClave_lc (conclave.c:215-384, -lc): score per template base first, then the total */
					if (tmp_score > bestScore) take = 1;
					else if (tmp_score == bestScore) {
						if (alignment_scores[t] > best_read_score) take = 1;
						else if (alignment_scores[t] == best_read_score) {
							if (uniq_alignment_scores[t] > bestNum) take = 1;
							else if (uniq_alignment_scores[t] == bestNum && t < abs(bestTemplate)) take = 1;
						}
					}
				} else if (alignment_scores[t] > best_read_score) take = 1;
				else if (alignment_scores[t] == best_read_score) {
					if (tmp_score > bestScore) take = 1;
					else if (tmp_score == bestScore) {
						if (uniq_alignment_scores[t] > bestNum) take = 1;
						else if (uniq_alignment_scores[t] == bestNum && t < abs(bestTemplate)) take = 1;
					}
				}
				if (take) {
					bestTemplate = tt; best_read_score = alignment_scores[t]; bestScore = tmp_score;
					bestNum = uniq_alignment_scores[t]; start = tmp_start; end = tmp_end;
				}
			}
		} else { bestTemplate = AT(T, 0); start = AT(S, 0); end = AT(E, 0); }
		if (n + 2 > acap) { acap *= 2; F = realloc(F, acap * sizeof(cfrag)); }
		cfrag *f = F + n++;
		memset(f, 0, sizeof(*f));
		f->q = q; f->hdr = hdr;
		if (bestTemplate < 0) {
			bestTemplate = -bestTemplate; f->rc = 1; flag |= 16;
			if (9 < hl && hdr[hl - 9] == 0) {
				int a, b; memcpy(&a, hdr + hl - 8, 4); memcpy(&b, hdr + hl - 4, 4);
				f->has_bound = 1; f->b0 = q_len - b; f->b1 = q_len - a;
			}
		}
		if (bestTemplate < 0 || DB_size <= bestTemplate) { free(F); return -2; }
		w_scores[bestTemplate] += read_score;
		fragmentCounts[bestTemplate]++; readCounts[bestTemplate]++;
		f->tmpl = bestTemplate;
		f->buf[0] = q_len; f->buf[1] = bestHits; f->buf[2] = sparse < 0 ? 0 : read_score; f->buf[3] = start; f->buf[4] = end;
		f->buf[5] = hl; f->buf[6] = flag;
		if (stats[2] < 0) {   /* the mate of the pair follows: same template and span, its own bytes, never turned */
			int m[3];
			if (ip + 12 > fb) { free(F); return -3; }
			memcpy(m, frag + ip, 12); ip += 12;
			readCounts[bestTemplate]++;
			cfrag *g = F + n++;
			memset(g, 0, sizeof(*g));
			g->tmpl = bestTemplate; g->q = frag + ip; g->hdr = g->q + m[0];
			ip += (size_t)m[0] + m[1];
			g->buf[0] = m[0]; g->buf[1] = bestHits; g->buf[2] = sparse < 0 ? 0 : read_score; g->buf[3] = start; g->buf[4] = end;
			g->buf[5] = m[1]; g->buf[6] = m[2];
		}
	}
	/* printFrags: template order; inside a template the list was built by prepending */
	size_t *cnt = calloc((size_t)DB_size + 1, sizeof(size_t));
	for (size_t i = 0; i < n; ++i) cnt[F[i].tmpl + 1]++;
	for (int t = 0; t < DB_size; ++t) cnt[t + 1] += cnt[t];
	size_t *order = malloc((n + 1) * sizeof(size_t));
	for (size_t i = n; i-- > 0;) order[cnt[F[i].tmpl]++] = i;   /* descending arrival inside each template */
	size_t op = 0;
	for (size_t k = 0; k < n; ++k) {
		const cfrag *f = F + order[k];
		const size_t need = 32 + (size_t)f->buf[0] + (size_t)f->buf[5];
		if (op + need + 4 > cap) { free(F); free(cnt); free(order); return -1; }
		memcpy(out + op, &f->tmpl, 4); memcpy(out + op + 4, f->buf, 28); op += 32;
		if (f->rc) for (int i = 0; i < f->buf[0]; ++i) out[op + i] = comp[f->q[f->buf[0] - 1 - i]];
		else memcpy(out + op, f->q, f->buf[0]);
		op += f->buf[0];
		memcpy(out + op, f->hdr, f->buf[5]);
		if (f->has_bound) { memcpy(out + op + f->buf[5] - 8, &f->b0, 4); memcpy(out + op + f->buf[5] - 4, &f->b1, 4); }
		op += f->buf[5];
	}
	if (op + 4 > cap) { free(F); free(cnt); free(order); return -1; }
	const int minus1 = -1;
	memcpy(out + op, &minus1, 4); op += 4;
	free(F); free(cnt); free(order);
	return (int64_t)op;
}


/* runConClave2 / runConClave2_lc (conclave.c:386-747 / 749-1110) over one frag_raw stream = the whole run.
 * uniq_alignment_scores is updated in place as the reference does (conclave.c:519); and_mode = cmp_and (kma.c:916) instead of
 * cmp_or; p_chisqr = the caller's (stdstat.c:136). Output as orc_conclave_stream. */
int64_t orc_conclave2_stream(const int32_t *template_lengths, int DB_size, const uint8_t *frag, size_t fb,
                             const uint64_t *alignment_scores, uint64_t *uniq_alignment_scores,
                             uint8_t *out, size_t cap, uint64_t *w_scores, uint32_t *fragmentCounts, uint32_t *readCounts,
                             double scoreT, double evalue, int and_mode, double (*p_chisqr)(long double)) {
	int32_t v;
#define AT(arr, i) (memcpy(&v, (const uint8_t *)(arr) + 4 * (size_t)(i), 4), v)
	/* pass 1: provisional choice -> w_scores */
	for (int pass = 0; pass < 2; ++pass) {
		size_t ip = 0;
		if (pass == 1) {   /* between the passes: discard insignificant templates (conclave.c:467-491) */
			unsigned long Nhits = 0, template_tot_ulen = 0;
			for (int t = 1; t < DB_size; ++t) { Nhits += w_scores[t]; template_tot_ulen += template_lengths[t]; }
			for (int t = DB_size; --t;) {
				int read_score;
				if ((read_score = w_scores[t])) {
					const int t_len = template_lengths[t];
					long double expected = t_len, q_value;
					expected /= (1 < (template_tot_ulen - t_len) ? (template_tot_ulen - t_len) : 1);
					expected *= (Nhits - read_score);
					q_value = read_score - expected;
					q_value /= (expected + read_score);
					q_value *= read_score - expected;
					const double p_value = p_chisqr(q_value);
					const int a = (p_value <= evalue && read_score > expected), b = (read_score >= scoreT * t_len);
					if ((and_mode ? (a && b) : (a || b)) == 0) w_scores[t] = 0;
				}
			}
		}
		while (ip + 20 <= fb) {
			int stats[5];
			memcpy(stats, frag + ip, 20);
			if (stats[0] == 0) break;
			ip += 20;
			const int q_len = stats[0], bestHits = abs(stats[1]), read_score = abs(stats[2]), hl = stats[3];
			const uint8_t *S = frag + ip + q_len + hl, *E = S + 4 * (size_t)bestHits, *T = E + 4 * (size_t)bestHits;
			ip += (size_t)q_len + hl + 12 * (size_t)bestHits;
			if (pass == 0) {
				int start, end;
				const int best = bestHits > 1 ? cc_four_keys(template_lengths, alignment_scores, uniq_alignment_scores, S, E, T, bestHits, -1, &start, &end)
				                              : AT(T, 0);
				w_scores[abs(best)] += read_score;
			} else if (bestHits != 1) {   /* pass 2: the sorting keys (conclave.c:493-530) */
				int best = 0;
				for (int i = bestHits; i--;) {
					const int t = abs(AT(T, i));
					if (w_scores[t]) { if (best) { best = 0; break; } else best = t; }
				}
				if (best) uniq_alignment_scores[best] += read_score;
			}
			if (stats[2] < 0) {
				int m[3];
				if (ip + 12 > fb) return -3;
				memcpy(m, frag + ip, 12); ip += 12 + (size_t)m[0] + m[1];
			}
		}
	}
#undef AT
	/* pass 3: the final choice and the fragments */
	memset(w_scores, 0, (size_t)DB_size * sizeof(uint64_t));
	orc_cc_version = 2;
	const int64_t r = orc_conclave_stream(template_lengths, DB_size, frag, fb, alignment_scores, uniq_alignment_scores, out, cap, w_scores,
	                                      fragmentCounts, readCounts);
	orc_cc_version = 1;
	return r;
}
