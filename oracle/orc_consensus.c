/* TEST INFRASTRUCTURE ONLY -- CPU oracle (restatement) of the consensus call of the assembly pass:
 * callConsensus (assembly.c:1499-1631) over the template nodes of one template, with the reference's five base
 * callers (assembly.c:162-271) and three significance tests (assembly.c:141-160). Pinned to the reference's own
 * callConsensus through oracle/ref_harness.c -consensus (tests/test_oracle_consensus.py).
 *
 * p_chisqr (stdstat.c:136-147): for statistics up to 49 the closed form 1 - 1.772453850 * erf(sqrt(q / 2)) / tgamma(1/2)
 * evaluated with the host libm; above 49 the reference reads a step table (fastp, stdstat.c:37-134) whose values there
 * are all <= 1e-11. The oracle returns 1e-11 for q > 49, which decides `p <= evalue` identically for every
 * evalue >= 1e-11; orc_consensus refuses smaller evalues instead of guessing the table. */
#include <math.h>
#include <string.h>
#include "orc.h"

static double orc_p_chisqr(long double q) {
	if (q < 0) return 1e-26;
	if (q > 49) return 1e-11;
	return 1 - 1.772453850 * erf(sqrt(0.5 * q)) / tgamma(0.5);
}

/* assembly.c:141-160 */
static int orc_significant(int sig, double support, int X, int Y, double evalue) {
	if (!(Y < X)) return 0;
	if (sig == 1 && !(9 * (X + Y) <= 10 * X)) return 0;
	if (sig == 2 && !(support * (X + Y) <= X)) return 0;
	return orc_p_chisqr(pow(X - Y, 2) / (X + Y)) <= evalue;
}

static int lower(int c) { return (c >= 'A' && c <= 'Z') ? c + 32 : c; }
static int upper(int c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }

/* assembly.c:162-271; caller: 0 baseCaller, 1 orgBaseCaller, 2 refCaller, 3 nanoCaller, 4 refNanoCaller */
static int orc_base_call(int caller, int sig, double support, int bestNuc, int tNuc, int bestScore, int depthUpdate, double evalue,
                         const uint16_t *counts) {
	static const char bases[6] = "ACGTN-";
	int j, bb, bn;
	switch (caller) {
	case 0:
		if (depthUpdate == 0) return '-';
		if (!orc_significant(sig, support, bestScore, depthUpdate - bestScore, evalue)) {
			if (bestNuc == '-' && tNuc != '-' && bestScore != depthUpdate) return 'n';
			return lower(bestNuc);
		}
		return bestNuc;
	case 1:
		if (depthUpdate == 0 || bestNuc == '-') return '-';
		if (!orc_significant(sig, support, bestScore, depthUpdate - bestScore, evalue)) return lower(bestNuc);
		return bestNuc;
	case 2:
		if (depthUpdate == 0 || (bestNuc == '-' && tNuc != '-')) return 'n';
		if (!orc_significant(sig, support, bestScore, depthUpdate - bestScore, evalue)) return lower(bestNuc);
		return bestNuc;
	case 3:
		if (depthUpdate == 0) return '-';
		if (!orc_significant(sig, support, bestScore, depthUpdate - bestScore, evalue)) {
			if (bestNuc == '-' && tNuc != '-' && bestScore != depthUpdate) {
				for (j = 0, bb = 0, bn = -1; j < 5; ++j) if (bb < counts[j]) { bb = counts[j]; bn = j; }
				return bb == 0 ? '-' : lower(bases[bn]);
			}
			return lower(bestNuc);
		}
		return bestNuc;
	default:
		if (depthUpdate == 0) return 'n';
		if (!orc_significant(sig, support, bestScore, depthUpdate - bestScore, evalue)) {
			if (bestNuc == '-') {
				for (j = 0, bb = 0, bn = -1; j < 5; ++j) if (bb < counts[j]) { bb = counts[j]; bn = j; }
				return bb == 0 ? 'n' : lower(bases[bn]);
			}
			return lower(bestNuc);
		}
		return bestNuc == '-' ? 'n' : bestNuc;
	}
}

/* callConsensus over the t_len template nodes (pos < t_len, visited in position order: without insertion nodes
 * assembly[pos].next == pos + 1). counts: uint16 [t_len][6]; seq: the template's packed words (stdnuc.h:20 getNuc).
 * stats: {depth, depthVar, len, aln_len, cover}. Returns 0, or -1 for an evalue the oracle cannot decide. */
int orc_consensus(const uint16_t *counts, const uint64_t *seq, int t_len, int bcd, int caller, int sig, double support, double evalue,
                  uint8_t *t, uint8_t *s, uint8_t *q, uint64_t *stats) {
	static const char bases[6] = "ACGTN-";
	uint64_t depth = 0, depthVar = 0, aln_len = 0, cover = 0;
	if (evalue < 1e-11) return -1;
	for (int pos = 0; pos < t_len; ++pos) {
		const uint16_t *c = counts + 6 * (size_t)pos;
		int bestNuc = (int)((seq[pos >> 5] << ((pos & 31) << 1)) >> 62), j, bestScore, bestBaseScore;
		unsigned long depthUpdate = 0;
		t[pos] = bases[bestNuc];
		bestScore = c[bestNuc];
		for (j = 0; j < 6; ++j) {
			if (bestScore < c[j]) { bestScore = c[j]; bestNuc = j; }
			depthUpdate += c[j];
		}
		bestNuc = bases[bestNuc];
		if (!depthUpdate) bestNuc = '-';
		else if ((unsigned long)(bestScore << 1) < depthUpdate) {   /* minor base call, assembly.c:1563-1579 */
			if (bestNuc == '-') {
				bestBaseScore = c[4]; bestNuc = 4;
				for (j = 0; j < 4; ++j) if (bestBaseScore < c[j]) { bestBaseScore = c[j]; bestNuc = j; }
				bestNuc = lower(bases[bestNuc]);
			} else bestNuc = lower(bestNuc);
			bestScore = (int)(depthUpdate - c[5]);
		} else if (depthUpdate < (unsigned long)bcd) bestNuc = lower(bestNuc);
		bestNuc = orc_base_call(caller, sig, support, bestNuc, t[pos], bestScore, (int)depthUpdate, evalue, c);
		q[pos] = (uint8_t)bestNuc;
		if (bestNuc != '-') {
			depth += depthUpdate; depthVar += depthUpdate * depthUpdate; ++aln_len;
			if (t[pos] == upper(bestNuc)) { ++cover; s[pos] = '|'; } else s[pos] = '_';
		} else s[pos] = '_';
	}
	stats[0] = depth; stats[1] = depthVar; stats[2] = (uint64_t)t_len; stats[3] = aln_len; stats[4] = cover;
	return 0;
}

/* the statistic from which on a call is significant: smallest double x in [0, 49] with p_chisqr(x) <= evalue
 * (p_chisqr falls with x); what the tests hand to the device path as kmagpu_consensus_params.chi2_min */
double orc_chi2_min(double evalue) {
	union { double d; uint64_t u; } lo, hi, mid;
	if (orc_p_chisqr(0.0L) <= evalue) return 0.0;
	lo.d = 0.0; hi.d = 49.0;
	if (!(orc_p_chisqr(hi.d) <= evalue)) return -1.0;
	while (hi.u - lo.u > 1) {
		mid.u = lo.u + ((hi.u - lo.u) >> 1);
		if (orc_p_chisqr(mid.d) <= evalue) hi = mid; else lo = mid;
	}
	return hi.d;
}
