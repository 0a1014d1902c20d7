/* TEST INFRASTRUCTURE ONLY -- CPU oracle (restatement) of stage 1 for well-formed single-line FASTQ / FASTA text:
 * the record splitter (FileBuffgetFq seqparse.c:241-400, FileBuffgetFsa), the base translation (to2Bit,
 * kma.c:1439-1482), phredStat (runinput.c:127-313: the -mp end trim, and after orc_stage1_set_quality the -mi hard mask
 * and the bidirectional -eq trim over the caller's prob[] table; no QC report), fsastat's N trim (runinput.c:315-368), the -ml / -xl filters, the pairing rule of run_input_PE
 * (runinput.c:516-539), compDNA (compdna.c:99-127) and the records of printFsa / printFsa_pair (runinput.c:765-825).
 * Pinned byte-exact to `kma -i / -ipe ... -s1` (tests/test_oracle_stage1.py). */
#include <string.h>
#include <math.h>
#include "orc.h"

void orc_to2bit(uint8_t *trans) {   /* kma.c:1439-1482: everything else 8, newline 16 */
	static const char *cls[5] = {"AaRrMmDd", "CcYyBb", "GgSsKkVv", "TtWwHhUu", "NnXx"};
	memset(trans, 8, 256);
	trans['\n'] = 16;
	for (int v = 0; v < 5; ++v) for (const char *c = cls[v]; *c; ++c) trans[(uint8_t)*c] = (uint8_t)v;
}

typedef struct { const uint8_t *hdr; int hdr_len; const uint8_t *seq; int seq_len; const uint8_t *qual; } s1_rec;

/* next record of a 4-line FASTQ / 2-line FASTA text; header without '@' / '>' and without trailing white space,
 * sequence without trailing bytes that translate to 8 (a '\r'). Returns 0 at the end of the text. */
static int s1_next(const uint8_t *text, size_t n, size_t *pos, int fastq, const uint8_t *trans, s1_rec *r) {
	size_t p = *pos;
	if (p >= n) return 0;
	const uint8_t *e = memchr(text + p, '\n', n - p);
	if (!e) return 0;
	r->hdr = text + p + 1;
	int hl = (int)(e - (text + p + 1));
	while (hl > 0 && (r->hdr[hl - 1] == ' ' || (r->hdr[hl - 1] >= 9 && r->hdr[hl - 1] <= 13))) --hl;
	r->hdr_len = hl;
	p = (size_t)(e - text) + 1;
	e = memchr(text + p, '\n', n - p);
	if (!e) e = text + n;
	r->seq = text + p;
	int sl = (int)(e - (text + p));
	while (sl > 0 && trans[r->seq[sl - 1]] == 8) --sl;
	r->seq_len = sl;
	p = (size_t)(e - text) + 1;
	r->qual = 0;
	if (fastq) {
		e = p < n ? memchr(text + p, '\n', n - p) : 0;   /* the '+' line */
		if (!e) return 0;
		p = (size_t)(e - text) + 1;
		if (p + (size_t)sl > n) return 0;
		r->qual = text + p;
		e = memchr(text + p + sl, '\n', n - p - sl);
		p = e ? (size_t)(e - text) + 1 : n;
	}
	*pos = p;
	return 1;
}

/* -eq (minQ), -mi (hardmaskQ) and the table kma.c:219 hands to run_input: prob[q] = 10^(-q / 10) */
static int g_minq = 0, g_maskq = 0;
static const double *g_prob = 0;
void orc_stage1_set_quality(int minQ, int hardmaskQ, const double *prob) { g_minq = minQ; g_maskq = hardmaskQ; g_prob = prob; }
int orc_stage1_maskq(void) { return g_maskq; }

/* the -mi / -eq part of phredStat (runinput.c:168-313) on the end-trimmed window [*start, *end): returns len - ns.
 * masked[] (may be NULL) is not kept: a base is an N afterwards iff it was one or its quality is below hardmaskQ (the raw
 * byte is compared, without the phred offset: runinput.c:183). */
static int s1_quality(const s1_rec *r, const uint8_t *trans, int scale, int thr, int minlen, int *start, int *end) {
	const uint8_t *q = r->qual, *sq = r->seq;
	const double *prob = g_prob - scale;
	int s = *start, e = *end, len = e - s, ns = 0;
	double sp = 0;
#define ISN(i) (trans[sq[i]] == 4 || q[i] < g_maskq)
	for (int i = s; i < e; ++i) { sp += prob[q[i]]; ns += ISN(i); }
	const double minP = pow(10, (-0.1) * g_minq);
	if (minlen <= (len - ns) && (minP * len) < sp) {
		int ns5 = 0, ns3 = 0, l5 = 0, l3 = 0, p5 = s, p3 = e - 1;
		double sp5 = 0, sp3 = 0;
		while (l3 < len && thr <= q[p3]) { sp3 += prob[q[p3]]; ++l3; ns3 += ISN(p3); --p3; }
		while (l3 < len && q[p3] < thr) { sp3 += prob[q[p3]]; ++l3; ns3 += ISN(p3); --p3; }
		while (minlen <= (len - ns) && (minP * len) < sp) {
			if ((sp5 * l3) < (sp3 * l5)) {
				e -= l3; ns -= ns3; len -= l3; sp -= sp3;
				ns3 = 0; l3 = 0; sp3 = 0;
				while (l3 < len && thr <= q[p3]) { sp3 += prob[q[p3]]; ++l3; ns3 += ISN(p3); --p3; }
				while (l3 < len && q[p3] < thr) { sp3 += prob[q[p3]]; ++l3; ns3 += ISN(p3); --p3; }
			} else {
				s += l5; len -= l5; ns -= ns5; sp -= sp5;
				ns5 = 0; l5 = 0; sp5 = 0;
				while (l5 < len && thr <= q[p5]) { sp5 += prob[q[p5]]; ++l5; ns5 += ISN(p5); ++p5; }
				while (l5 < len && q[p5] < thr) { sp5 += prob[q[p5]]; ++l5; ns5 += ISN(p5); ++p5; }
			}
		}
	}
#undef ISN
	*start = s; *end = e;
	return len - ns;
}

/* phredStat / fsastat: the kept window and the length the -ml filter sees */
static int s1_window(const s1_rec *r, const uint8_t *trans, int fastq, int scale, int thr, int minlen, int maxlen, int *start, int *end) {
	int s = 0, e = r->seq_len;
	if (maxlen < r->seq_len) { *start = *end = 0; return 0; }
	if (fastq) {
		while (s < e && r->qual[s] < thr) ++s;
		while (s < e && r->qual[e - 1] < thr) --e;
		*start = s; *end = e;
		if (!g_minq && !g_maskq) return e - s;
		return s1_quality(r, trans, scale, thr, minlen, start, end);
	}
	while (s < e && trans[r->seq[e - 1]] == 4) --e;
	while (s < e && trans[r->seq[s]] == 4) ++s;
	int ns = 0;
	for (int i = s; i < e; ++i) ns += trans[r->seq[i]] == 4;
	*start = s; *end = e;
	return e - s - ns;
}

/* compDNA + printFsa: one stage-1 record */
static size_t s1_emit(const s1_rec *r, const uint8_t *trans, int start, int end, int neg, uint8_t *out, size_t cap, size_t at) {
	const int L = end - start, words = (L + 31) >> 5;
	int nN = 0;
#define ISN(i) (trans[r->seq[i]] == 4 || (r->qual && r->qual[i] < g_maskq))
	for (int i = start; i < end; ++i) nN += ISN(i);
	const int hl = r->hdr_len + 1;
	const size_t need = 16 + 8 * (size_t)words + 4 * (size_t)nN + (size_t)hl;
	if (at + need > cap) return at + need;
	int32_t h[4] = {L, words, nN, neg ? -hl : hl};
	memcpy(out + at, h, 16);
	uint64_t *w = (uint64_t *)(out + at + 16);
	int32_t *N = (int32_t *)(out + at + 16 + 8 * (size_t)words);
	int k = 0;
	for (int i = 0; i < words; ++i) {
		uint64_t v = 0;
		for (int j = 0; j < 32; ++j) {
			const int p = 32 * i + j;
			const int c = p < L ? (ISN(start + p) ? 4 : trans[r->seq[start + p]]) : 0;
			v <<= 2;
			if (c == 4) { int32_t pp = p; memcpy(&N[k++], &pp, 4); } else v |= (uint64_t)(c & 3);
		}
		memcpy(&w[i], &v, 8);
	}
	memcpy(out + at + need - hl, r->hdr, (size_t)r->hdr_len);
	out[at + need - 1] = 0;
	return at + need;
#undef ISN
}

/* FileBuffgetFsa (seqparse.c:66-160) on multi-line FASTA, restated as a rewrite into 2-line FASTA: the header line as it
 * is, then every byte up to the next '>' (or the end of the text) that trans maps below 8, as one line. Returns the bytes
 * written (out holds at least n + 1). */
size_t orc_fasta_unwrap(const uint8_t *text, size_t n, uint8_t *out) {
	uint8_t trans[256];
	orc_to2bit(trans);
	size_t p = 0, o = 0;
	while (p < n && text[p] == '>') {
		const size_t o0 = o;
		while (p < n && text[p] != '\n') out[o++] = text[p++];
		if (p >= n) { o = o0; break; }   /* a header line cut off by the end of the file is no record (seqparse.c:85-91) */
		out[o++] = text[p++];
		while (p < n && text[p] != '>') { if (trans[text[p]] < 8) out[o++] = text[p]; ++p; }
		out[o++] = '\n';
	}
	return o;
}

/* text2 != NULL: run_input_PE over the two files in lockstep. Returns the bytes of the stream (if > cap: needed). */
size_t orc_stage1(const uint8_t *text1, size_t n1, const uint8_t *text2, size_t n2, int fastq, int min_phred, int phred_scale,
                  int minlen, int maxlen, uint8_t *out, size_t cap, int64_t *count) {
	uint8_t trans[256];
	orc_to2bit(trans);
	if (min_phred < g_minq) min_phred = g_minq;   /* runinput.c:380 */
	const int thr = phred_scale + min_phred;
	size_t p1 = 0, p2 = 0, at = 0;
	int64_t cnt = 0;
	s1_rec a, b;
	for (;;) {
		const int g1 = s1_next(text1, n1, &p1, fastq, trans, &a);
		const int g2 = text2 ? s1_next(text2, n2, &p2, fastq, trans, &b) : 0;
		if (!g1 && !g2) break;
		int s1 = 0, e1 = 0, s2 = 0, e2 = 0;
		const int l1 = g1 ? s1_window(&a, trans, fastq, phred_scale, thr, minlen, maxlen, &s1, &e1) : -1;
		const int l2 = g2 ? s1_window(&b, trans, fastq, phred_scale, thr, minlen, maxlen, &s2, &e2) : -1;
		const int k1 = g1 && minlen <= l1, k2 = g2 && minlen <= l2;
		if (k1 && k2) { at = s1_emit(&a, trans, s1, e1, 1, out, cap, at); at = s1_emit(&b, trans, s2, e2, 0, out, cap, at); ++cnt; }
		else if (k1) { at = s1_emit(&a, trans, s1, e1, 0, out, cap, at); ++cnt; }
		else if (k2) { at = s1_emit(&b, trans, s2, e2, 0, out, cap, at); ++cnt; }
	}
	if (count) *count = cnt;
	return at;
}
