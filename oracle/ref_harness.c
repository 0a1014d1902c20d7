/*
 * TEST INFRASTRUCTURE ONLY -- driver around the UNMODIFIED reference (linked from oracle/_ref/libkma.a).
 * Nothing here restates reference logic: it sets the globals the way kma.c / runkma.c do and calls the
 * reference's own stage-3 entry point, so tests get ground truth for the alignment pass.
 *
 *   ref_aln <db_prefix> <stage2.bin | -> <frag_raw.out> <scores.out> [cand.out] [-1t1] [-apm-p] [-t N]
 *
 * "-" reads the stage-2 stream from stdin (`kma ... -s2 | ref_aln db - ...`); -t N runs alnFrags_threaded on N pthreads
 * the way runKMA does (runkma.c:300-440: one Aln_thread per thread, shared input/output/score arrays, the reference's
 * own spin locks) -- this is the CPU baseline of the alignment pass in bench.py.
 *
 * frag_raw.out : what alnFrags_threaded (alnfrags.c:2150) writes to frag_out_raw for the stream
 * scores.out   : int32 DB_size, uint64 alignment_scores[DB_size], uint64 uniq_alignment_scores[DB_size]
 * cand.out     : (optional) one 8 x int32 row per (read, candidate template) from a second pass that calls the
 *                reference's anker_rc_comp / KMA_score exactly as alnFragsSE (alnfrags.c:1080-1128) does:
 *                {read index, template (signed as aligned), score, len, pos, match, tGaps, qGaps}
 * Built by Makefile.ref; sources are compiled where they lie under /root/reference.
 */
#define _GNU_SOURCE
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <pthread.h>
#include "align.h"
#include "assembly.h"
#include "conclave.h"
#include "updatescores.h"
#include "frags.h"
#include "alnfrags.h"
#include "ankers.h"
#include "chain.h"
#include "compdna.h"
#include "hashmapcci.h"
#include "nw.h"
#include "penalties.h"
#include "qseqs.h"
#include "runkma.h"
#include "stdstat.h"
#ifndef MAX
#define MAX(X, Y) ((X) < (Y) ? (Y) : (X))
#endif

static Penalties *make_rewards(void) {
	Penalties *r = calloc(1, sizeof(Penalties));
	int Ts = -2, Tv = -2, i, j;
	r->M = 1; r->MM = (Ts + Tv - 1) / 2; r->U = -1; r->W1 = -3; r->Wl = -6; r->Mn = 0; r->PE = 7;
	int **d = malloc(5 * sizeof(int *) + 25 * sizeof(int));
	d[0] = (int *)(d + 5);
	for (i = 0; i < 5; ++i) d[i] = d[0] + 5 * i;
	for (i = 0; i < 4; ++i) {
		for (j = 0; j < 4; ++j) d[i][j] = Tv;
		d[i][4] = r->Mn;
		d[i][(i - 2) < 0 ? (i + 2) : (i - 2)] = Ts;
		d[i][i] = r->M;
	}
	for (j = 0; j < 5; ++j) d[4][j] = r->Mn;
	d[4][4] = 0;
	r->d = d;
	return r;
}

/* ref_aln -trace <db> <frags.bin> <out.bin> [-1t1]: the alignment part of assemble_KMA's inner loop
 * (assembly.c:1868-1961) -- the reference's own anker_rc + KMA per fragment record (frags.c:45-48); per record
 * int32[12]{accepted, read_score, start, end, score, len, pos, match, tGaps, qGaps, oriented, 0} + t s q rows. */
static int trace_main(int argc, char **argv) {
	int one2one = 0, exhaustive = 0, ts = 0, dense = 0;
	const char *mat_path = 0;   /* -mat <file> [-dense]: the reference's alnToMat / alnToMatDense on every accepted alignment */
	for (int a = 5; a < argc; ++a) {
		if (!strcmp(argv[a], "-1t1")) one2one = 1;
		else if (!strcmp(argv[a], "-dense")) dense = 1;
		else if (!strcmp(argv[a], "-mat") && a + 1 < argc) mat_path = argv[++a];
		else if (!strcmp(argv[a], "-ts") && a + 1 < argc) ts = atoi(argv[++a]);   /* kma.c:571 */
	}
	char path[4096];
	int *template_lengths; long unsigned *as, *uas;
	char *p2 = malloc(strlen(argv[2]) + 64); strcpy(p2, argv[2]);
	int DB_size = load_DBs_KMA(p2, &as, &uas, &template_lengths, 0);
	int kmersize = template_lengths[0];
	if (kmersize < 4 || 31 < kmersize) kmersize = 16;
	snprintf(path, sizeof(path), "%s.seq.b", argv[2]);
	int seq_in = open(path, O_RDONLY);
	if (seq_in < 0) { perror(path); return 1; }
	long *seq_indexes = malloc((DB_size + 1) * sizeof(long));
	seq_indexes[0] = 0; seq_indexes[1] = 0;
	for (int i = 2; i < DB_size; ++i) seq_indexes[i] = seq_indexes[i - 1] + ((template_lengths[i - 1] >> 5) + 1) * sizeof(long unsigned);
	Penalties *rewards = make_rewards();
	preseed(0, 0, exhaustive);
	trimSeedsPtr(0, ts);
	anker_rc(0, 0, one2one, 0, 0, 0);
	anker_rc_comp(0, 0, (unsigned char *)(&one2one), 0, 0, 0, 0, 0);
	alignLoadPtr = &alignLoad_fly;
	HashMapCCI **templates_index = calloc(DB_size, sizeof(HashMapCCI *));
	NWmat *NWm = malloc(sizeof(NWmat));
	NWm->NW_s = 1024 * 1024; NWm->NW_q = 1024; NWm->E = malloc(NWm->NW_s);
	NWm->D[0] = malloc((NWm->NW_q << 1) * sizeof(int)); NWm->P[0] = malloc((NWm->NW_q << 1) * sizeof(int));
	NWm->D[1] = NWm->D[0] + NWm->NW_q; NWm->P[1] = NWm->P[0] + NWm->NW_q; NWm->rewards = rewards;
	AlnPoints *points = seedPoint_init(1024, rewards);
	FILE *in = fopen(argv[3], "rb"), *out = fopen(argv[4], "wb");
	if (!in || !out) { perror("open"); return 1; }
	AssemInfo **mats = calloc(DB_size, sizeof(AssemInfo *));
	Assem *aa = calloc(1, sizeof(Assem));
	Aln *aligned = calloc(1, sizeof(Aln)), *gap_align = calloc(1, sizeof(Aln));
	int delta = 0;
	unsigned char *qseq = 0, *orig = 0; int qsize = 0;
	int h[8];
	const int Wl = -rewards->Wl, minlen = 16, mq = 0;
	const double scoreT = 0.5, mrc = 0.0;
	while (fread(h, 4, 8, in) == 8 && h[0] >= 0) {
		int template = h[0], q_len = h[1], read_score = h[3], st2 = h[4], st3 = h[5], hl = h[6];
		if (qsize < q_len + 64) { qsize = 2 * q_len + 64; qseq = realloc(qseq, qsize); orig = realloc(orig, qsize); }
		if (fread(qseq, 1, q_len, in) != (size_t)q_len) break;
		memcpy(orig, qseq, q_len);
		/* q-bound of chain-mode records, as assemble_KMA reads it (assembly.c:1916-1923) */
		int q_start = 0, q_end = q_len;
		{
			unsigned char *hb = malloc(hl + 1);
			if (fread(hb, 1, hl, in) != (size_t)hl) break;
			if (2 * sizeof(int) + 1 < (size_t)hl && hb[hl - 2 * sizeof(int) - 1] == 0) { memcpy(&q_start, hb + hl - 8, 4); memcpy(&q_end, hb + hl - 4, 4); }
			free(hb);
		}
		if (delta < q_len) {
			delta = q_len << 1;
			aligned->t = realloc(aligned->t, (delta + 1) << 1); aligned->s = realloc(aligned->s, (delta + 1) << 1); aligned->q = realloc(aligned->q, (delta + 1) << 1);
			gap_align->t = realloc(gap_align->t, (delta + 1) << 1); gap_align->s = realloc(gap_align->s, (delta + 1) << 1); gap_align->q = realloc(gap_align->q, (delta + 1) << 1);
		}
		if (!templates_index[template]) templates_index[template] = alignLoadPtr(0, seq_in, template_lengths[template], kmersize, seq_indexes[template]);
		HashMapCCI *ti = templates_index[template];
		int t_len = template_lengths[template];
		int r[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, len = 0;
		points->len = 0;
		int go = read_score || anker_rc(ti, qseq, q_len, q_start, q_end, points);
		r[10] = memcmp(orig, qseq, q_len) != 0;
		if (go) {
			if (st3 <= st2) { st2 = 0; st3 = t_len; }
			AlnScore a = KMA(ti, qseq, q_len, q_start, q_end, aligned, gap_align, st2, t_len < st3 ? t_len : st3, mq, scoreT, points, NWm);
			int aln_len = a.len, start = a.pos, end = start + aln_len - a.tGaps;
			double score;
			if (t_len < end) end -= t_len;
			read_score = a.score;
			if (start == 0) read_score += Wl;
			if (end == t_len) read_score += Wl;
			if (minlen <= aln_len && ((mrc * q_len <= a.len - a.qGaps) || (mrc * t_len <= a.len - a.tGaps))) score = 1.0 * read_score / aln_len;
			else { read_score = 0; score = 0; }
			r[0] = 0 < read_score && scoreT <= score;
			r[1] = read_score; r[2] = start; r[3] = end;
			r[4] = a.score; r[5] = a.len; r[6] = a.pos; r[7] = a.match; r[8] = a.tGaps; r[9] = a.qGaps;
			len = aligned->len;
			if (mat_path && r[0]) {   /* assembly.c:1968: the matrix as assemble_KMA initialises it (assembly.c:1811-1856) */
				if (!mats[template]) {
					AssemInfo *m = mats[template] = malloc(sizeof(AssemInfo));
					m->len = t_len; m->size = t_len << 1;
					m->assmb = malloc(m->size * sizeof(Assembly));
					for (int i = 0; i < t_len; ++i) { memset(m->assmb[i].counts, 0, 12); m->assmb[i].next = i + 1; }
					m->assmb[t_len - 1].next = 0;
				}
				if (dense) alnToMatDense(mats[template], aa, aligned, a, t_len, h[7]);
				else alnToMat(mats[template], aa, aligned, a, t_len, h[7]);
			}
		}
		points->len = 0;
		r[11] = len;
		fwrite(r, 4, 12, out);
		fwrite(aligned->t, 1, len, out); fwrite(aligned->s, 1, len, out); fwrite(aligned->q, 1, len, out);
	}
	fclose(in); fclose(out);
	if (mat_path) {   /* per template that saw an alignment: int32 {template, t_len, nodes}, then counts[6] of the t_len template nodes */
		FILE *mo = fopen(mat_path, "wb");
		for (int t = 1; t < DB_size; ++t) if (mats[t]) {
			int hd[3] = {t, template_lengths[t], mats[t]->len};
			fwrite(hd, 4, 3, mo);
			for (int i = 0; i < template_lengths[t]; ++i) fwrite(mats[t]->assmb[i].counts, 2, 6, mo);
		}
		fclose(mo);
	}
	return 0;
}

/* -conclave db frag_raw.bin scores.bin out.bin: the reference's runConClave (conclave.c:43-213) + printFrags
 * (frags.c:30-61) on a frag_raw stream with the given ConClave accumulators (scores.bin: int32 DB_size, u64
 * alignment_scores[DB_size], u64 uniq_alignment_scores[DB_size] -- the file the alignment pass of this harness writes).
 * out.bin: int32 nfiles, then per file int64 bytes + the bytes printFrags wrote; then u64 w_scores[DB_size],
 * u32 fragmentCounts[DB_size], u32 readCounts[DB_size]. maxFrag as argv[6] (default 1048576, kma.c:340). */
static int conclave_main(int argc, char **argv) {
	if (argc < 6) { fprintf(stderr, "usage: ref_aln -conclave db frag_raw.bin scores.bin out.bin [maxFrag]\n"); return 2; }
	int *template_lengths; long unsigned *as, *uas;
	char *p2 = malloc(strlen(argv[2]) + 64); strcpy(p2, argv[2]);
	int DB_size = load_DBs_KMA(p2, &as, &uas, &template_lengths, 0);
	int maxFrag = argc > 6 && argv[6][0] != '-' ? atoi(argv[6]) : 1048576;
	const int lc = !strcmp(argv[argc - 1], "-lc");   /* runConClave_lc (kma.c:700: -lc) */
	int version = 1; double c2_scoreT = 0.5, c2_evalue = 0.05;   /* -c2 scoreT evalue [-and]: runConClave2 (-ConClave 2, runkma.c:591) */
	for (int a = 6; a < argc; ++a) {
		if (!strcmp(argv[a], "-c2") && a + 2 < argc) { version = 2; c2_scoreT = strtod(argv[a + 1], 0); c2_evalue = strtod(argv[a + 2], 0); }
		else if (!strcmp(argv[a], "-and")) cmp = &cmp_and;   /* kma.c:916 */
	}
	FILE *sc = fopen(argv[4], "rb");
	int n = 0;
	if (!sc || fread(&n, 4, 1, sc) != 1 || n != DB_size) { fprintf(stderr, "scores file does not match the database\n"); return 1; }
	if (fread(as, 8, DB_size, sc) != (size_t)DB_size || fread(uas, 8, DB_size, sc) != (size_t)DB_size) return 1;
	fclose(sc);
	/* frag_raw as runKMA leaves it: the records followed by an int 0 (runkma.c:444) */
	FILE *in = fopen(argv[3], "rb");
	if (!in) { perror(argv[3]); return 1; }
	FILE *tmp = tmpfile();
	char buf[1 << 16]; size_t got; long maxq = 1024;
	while ((got = fread(buf, 1, sizeof(buf), in))) fwrite(buf, 1, got, tmp);
	int zero = 0; fwrite(&zero, 4, 1, tmp);
	fseek(in, 0, SEEK_END); maxq = ftell(in) + 64; fclose(in);
	rewind(tmp);
	Qseqs *header = setQseqs(maxq), *qseq = setQseqs(maxq);
	int *bestTemplates = malloc(((DB_size + 1) << 1) * sizeof(int)), *bs = malloc(((DB_size + 1) << 1) * sizeof(int)), *be = malloc(((DB_size + 1) << 1) * sizeof(int));
	FILE **template_fragments = calloc(DB_size + 1, sizeof(FILE *));
	Frag **alignFrags = calloc(DB_size, sizeof(Frag *));
	long unsigned *w_scores = calloc(DB_size, sizeof(long unsigned));
	unsigned *fragmentCounts = calloc(DB_size, sizeof(unsigned)), *readCounts = calloc(DB_size, sizeof(unsigned));
	int files;
	if (version == 2) {
		long unsigned template_tot_ulen = 0;
		for (int i = 1; i < DB_size; ++i) template_tot_ulen += template_lengths[i];   /* runkma.c:581-585 */
		files = (lc ? runConClave2_lc : runConClave2)(tmp, &template_fragments, DB_size, maxFrag, w_scores, fragmentCounts, readCounts, as, uas,
		                        template_lengths, header, qseq, bestTemplates, bs, be, alignFrags, template_tot_ulen, c2_scoreT, c2_evalue);
	} else files = (lc ? runConClave_lc : runConClave)(tmp, &template_fragments, DB_size, maxFrag, w_scores, fragmentCounts, readCounts, as, uas, template_lengths,
	                        header, qseq, bestTemplates, bs, be, alignFrags);
	FILE *out = fopen(argv[5], "wb");
	fwrite(&files, 4, 1, out);
	for (int f = 0; f < files; ++f) {
		FILE *tf = template_fragments[f];
		fseek(tf, 0, SEEK_END);
		long long bytes = ftell(tf);
		rewind(tf);
		fwrite(&bytes, 8, 1, out);
		while ((got = fread(buf, 1, sizeof(buf), tf))) fwrite(buf, 1, got, out);
	}
	fwrite(w_scores, 8, DB_size, out); fwrite(fragmentCounts, 4, DB_size, out); fwrite(readCounts, 4, DB_size, out);
	if (version == 2) fwrite(uas, 8, DB_size, out);   /* runConClave2 updates the unique scores (conclave.c:519) */
	fclose(out);
	return 0;
}

/* -memscore db s2.bin frag_raw.out scores.out: the "Collecting k-mer scores" loop of runKMA_MEM (runkma.c:1088-1140,
 * -mem_mode) around the reference's own get_ankers, unCompDNA, update_Scores_MEM and update_Scores_pe_MEM. */
static int memscore_main(int argc, char **argv) {
	if (argc < 6) { fprintf(stderr, "usage: ref_aln -memscore db s2.bin frag_raw.out scores.out\n"); return 2; }
	int *template_lengths; long unsigned *alignment_scores, *uniq_alignment_scores;
	char *p2 = malloc(strlen(argv[2]) + 64); strcpy(p2, argv[2]);
	int DB_size = load_DBs_KMA(p2, &alignment_scores, &uniq_alignment_scores, &template_lengths, 0);
	int kmersize = template_lengths[0];
	if (kmersize < 4 || 31 < kmersize) kmersize = 16;
	FILE *inputfile = fopen(argv[3], "rb"), *frag_out_raw = fopen(argv[4], "wb");
	if (!inputfile || !frag_out_raw) { perror("open"); return 1; }
	CompDNA *qseq_comp = malloc(sizeof(CompDNA)), *qseq_r_comp = malloc(sizeof(CompDNA));
	allocComp(qseq_comp, 1024); allocComp(qseq_r_comp, 1024);
	Qseqs *qseq = setQseqs(1024), *qseq_r = setQseqs(1024), *header = setQseqs(256), *header_r = setQseqs(256);
	int *matched_templates = malloc(((DB_size + 1) << 1) * sizeof(int));
	int *best_start_pos = calloc((DB_size << 1), sizeof(int)), *best_end_pos = malloc((DB_size << 1) * sizeof(int));
	int *bestTemplates = matched_templates + 1;
	int rc_flag, flag, flag_r, read_score = 0, best_read_score, bestHits, i, delta = 1024;
	qseq_r->len = 0;
	while ((rc_flag = get_ankers(matched_templates, qseq_comp, header, &flag, inputfile)) != 0) {
		if (*matched_templates) read_score = 0;
		else {
			read_score = get_ankers(matched_templates, qseq_r_comp, header_r, &flag_r, inputfile);
			read_score = labs(read_score);
			qseq_r->len = qseq_r_comp->seqlen;
		}
		qseq->len = qseq_comp->seqlen;
		if (kmersize <= qseq->len) {
			if (delta <= MAX(qseq->len, qseq_r->len)) {
				delta = MAX(qseq->len, qseq_r->len); delta <<= 1;
				qseq->size = delta; qseq_r->size = delta;
				free(qseq->seq); free(qseq_r->seq);
				qseq->seq = malloc(delta); qseq_r->seq = malloc(delta);
			}
			unCompDNA(qseq_comp, qseq->seq);
			best_read_score = abs(rc_flag);
			for (i = 1, bestHits = 0; i <= *matched_templates; ++i, ++bestHits) best_end_pos[bestHits] = template_lengths[abs(matched_templates[i])];
			if (rc_flag < 0 && 0 < matched_templates[*matched_templates]) bestHits = -bestHits;
			if (read_score && kmersize <= qseq_r->len) {
				unCompDNA(qseq_r_comp, qseq_r->seq);
				update_Scores_pe_MEM(qseq->seq, qseq->len, qseq_r->seq, qseq_r->len, bestHits, best_read_score + read_score, best_start_pos, best_end_pos,
				                     bestTemplates, header, header_r, flag, flag_r, alignment_scores, uniq_alignment_scores, frag_out_raw);
			} else update_Scores_MEM(qseq->seq, qseq->len, bestHits, best_read_score, best_start_pos, best_end_pos, bestTemplates, header, flag,
			                         alignment_scores, uniq_alignment_scores, frag_out_raw);
		}
	}
	fclose(frag_out_raw);
	FILE *so = fopen(argv[5], "wb");
	fwrite(&DB_size, 4, 1, so); fwrite(alignment_scores, 8, DB_size, so); fwrite(uniq_alignment_scores, 8, DB_size, so);
	fclose(so);
	return 0;
}

/* -consensus db mat.bin out.bin [-bcd N] [-evalue X] [-caller 0..4] [-sig 0..2] [-support X]: the reference's own
 * callConsensus (assembly.c:1499-1631) on base-count matrices holding template nodes only. mat.bin: per template
 * int32 {template, t_len, nodes} + uint16 counts[t_len][6] (what -trace -mat writes); out.bin: per template
 * int32 {template, t_len}, u64 {depth, depthVar, len, aln_len, cover}, then the t, s, q rows (t_len bytes each). */
static int consensus_main(int argc, char **argv) {
	if (argc < 5) { fprintf(stderr, "usage: ref_aln -consensus db mat.bin out.bin [options]\n"); return 2; }
	int bcd = 1, caller = 0, sig = 0;
	double evalue = 0.05, support = 0.0;
	for (int a = 5; a + 1 < argc; a += 2) {
		if (!strcmp(argv[a], "-bcd")) bcd = atoi(argv[a + 1]);
		else if (!strcmp(argv[a], "-evalue")) evalue = strtod(argv[a + 1], 0);
		else if (!strcmp(argv[a], "-caller")) caller = atoi(argv[a + 1]);
		else if (!strcmp(argv[a], "-sig")) sig = atoi(argv[a + 1]);
		else if (!strcmp(argv[a], "-support")) support = strtod(argv[a + 1], 0);
	}
	/* the bindings kma.c:743-766 makes for -bc / -bc90 / -bcg / -bcNano (and the two ref callers) */
	baseCall = caller == 0 ? &baseCaller : caller == 1 ? &orgBaseCaller : caller == 2 ? &refCaller : caller == 3 ? &nanoCaller : &refNanoCaller;
	significantBase = sig == 0 ? &significantNuc : sig == 1 ? &significantAnd90Nuc : &significantAndSupport;
	if (sig == 2) significantAndSupport(0, 0, support);
	int *template_lengths; long unsigned *as, *uas;
	char *p2 = malloc(strlen(argv[2]) + 64); strcpy(p2, argv[2]);
	int DB_size = load_DBs_KMA(p2, &as, &uas, &template_lengths, 0);
	char path[4096];
	snprintf(path, sizeof(path), "%s.seq.b", argv[2]);
	FILE *sf = fopen(path, "rb");
	if (!sf) { perror(path); return 1; }
	long *seq_indexes = malloc((DB_size + 1) * sizeof(long));
	seq_indexes[0] = 0; seq_indexes[1] = 0;
	for (int i = 2; i < DB_size; ++i) seq_indexes[i] = seq_indexes[i - 1] + ((template_lengths[i - 1] >> 5) + 1) * sizeof(long unsigned);
	FILE *in = fopen(argv[3], "rb"), *out = fopen(argv[4], "wb");
	if (!in || !out) { perror("open"); return 1; }
	int hd[3];
	while (fread(hd, 4, 3, in) == 3) {
		const int t = hd[0], t_len = hd[1];
		if (t <= 0 || t >= DB_size || t_len != template_lengths[t]) { fprintf(stderr, "matrix does not match the database\n"); return 1; }
		AssemInfo m;
		m.len = t_len; m.size = t_len << 1;
		m.assmb = malloc(m.size * sizeof(Assembly));
		for (int i = 0; i < t_len; ++i) {
			if (fread(m.assmb[i].counts, 2, 6, in) != 6) return 1;
			m.assmb[i].next = i + 1;
		}
		m.assmb[t_len - 1].next = 0;
		const int words = (t_len >> 5) + 1;
		long unsigned *seq = calloc(words + 1, sizeof(long unsigned));
		fseek(sf, seq_indexes[t], SEEK_SET);
		if (fread(seq, sizeof(long unsigned), words, sf) != (size_t)words) { fprintf(stderr, "short read of %s\n", path); return 1; }
		Assem aa;
		memset(&aa, 0, sizeof(aa));
		aa.size = (t_len + 1) << 1;
		aa.t = malloc(aa.size); aa.s = malloc(aa.size); aa.q = malloc(aa.size);
		callConsensus(&m, &aa, seq, t_len, bcd, evalue, 1);
		int oh[2] = {t, t_len};
		unsigned long long st[5] = {aa.depth, aa.depthVar, aa.len, aa.aln_len, aa.cover};
		fwrite(oh, 4, 2, out); fwrite(st, 8, 5, out);
		fwrite(aa.t, 1, t_len, out); fwrite(aa.s, 1, t_len, out); fwrite(aa.q, 1, t_len, out);
		free(aa.t); free(aa.s); free(aa.q); free(seq); free(m.assmb);
	}
	fclose(in); fclose(out);
	return 0;
}

int main(int argc, char **argv) {
	if (argc >= 5 && !strcmp(argv[1], "-memscore")) return memscore_main(argc, argv);
	if (argc >= 5 && !strcmp(argv[1], "-conclave")) return conclave_main(argc, argv);
	if (argc >= 5 && !strcmp(argv[1], "-trace")) return trace_main(argc, argv);
	if (argc >= 5 && !strcmp(argv[1], "-consensus")) return consensus_main(argc, argv);
	if (argc < 5) { fprintf(stderr, "usage: ref_aln db s2.bin frag_raw.out scores.out [cand.out] [-1t1]\n"); return 2; }
	int one2one = 0, exhaustive = 0, ts = 0;
	const char *cand_path = 0;
	int nthreads = 1;
	double min_frac = 1.0;
	for (int a = 5; a < argc; ++a) {
		if (!strcmp(argv[a], "-1t1")) one2one = 1;
		else if (!strcmp(argv[a], "-apm-p")) alnFragsPE = &alnFragsPenaltyPE;   /* kma.c:458: -apm p */
		else if (!strcmp(argv[a], "-apm-u")) alnFragsPE = &alnFragsUnionPE;     /* kma.c:460: -apm u (the default) */
		else if (!strcmp(argv[a], "-t") && a + 1 < argc) nthreads = atoi(argv[++a]);
		else if (!strcmp(argv[a], "-mf") && a + 1 < argc) min_frac = strtod(argv[++a], 0);   /* minFrac of -proxi as runKMA gets it (kma.c:1622) */
		else cand_path = argv[a];
	}
	if (nthreads < 1) nthreads = 1;
	char path[4096];
	int *template_lengths; long unsigned *as, *uas;
	snprintf(path, sizeof(path), "%s", argv[1]);
	char *p2 = malloc(strlen(path) + 64); strcpy(p2, path);
	int DB_size = load_DBs_KMA(p2, &as, &uas, &template_lengths, 0);
	int kmersize = template_lengths[0];
	if (kmersize < 4 || 31 < kmersize) kmersize = 16;
	snprintf(path, sizeof(path), "%s.seq.b", argv[1]);
	int seq_in = open(path, O_RDONLY);
	if (seq_in < 0) { perror(path); return 1; }
	long *seq_indexes = malloc((DB_size + 1) * sizeof(long));
	seq_indexes[0] = 0; seq_indexes[1] = 0;
	for (int i = 2; i < DB_size; ++i) seq_indexes[i] = seq_indexes[i - 1] + ((template_lengths[i - 1] >> 5) + 1) * sizeof(long unsigned);

	Penalties *rewards = make_rewards();
	/* the NULL-argument "setter" calls of kma.c:1249-1252, 1428-1429 */
	preseed(0, 0, exhaustive);
	trimSeedsPtr(0, ts);
	anker_rc(0, 0, one2one, 0, 0, 0);
	anker_rc_comp(0, 0, (unsigned char *)(&one2one), 0, 0, 0, 0, 0);
	alignLoadPtr = &alignLoad_fly;

	for (int pass = 0; pass < (cand_path ? 2 : 1); ++pass) {
		FILE *in = strcmp(argv[2], "-") ? fopen(argv[2], "rb") : stdin;
		if (!in) { perror(argv[2]); return 1; }
		HashMapCCI **templates_index = calloc(DB_size, sizeof(HashMapCCI *));
		CompDNA *qc = malloc(sizeof(CompDNA)), *qrc = malloc(sizeof(CompDNA));
		allocComp(qc, 1024); allocComp(qrc, 1024);
		NWmat *NWm = malloc(sizeof(NWmat));
		NWm->NW_s = 1024 * 1024; NWm->NW_q = 1024; NWm->E = malloc(NWm->NW_s);
		NWm->D[0] = malloc((NWm->NW_q << 1) * sizeof(int)); NWm->P[0] = malloc((NWm->NW_q << 1) * sizeof(int));
		NWm->D[1] = NWm->D[0] + NWm->NW_q; NWm->P[1] = NWm->P[0] + NWm->NW_q; NWm->rewards = rewards;
		AlnPoints *points = seedPoint_init(1024, rewards);
		int *matched = malloc(((DB_size + 1) << 1) * sizeof(int));
		if (pass == 0) {
			FILE *fout = fopen(argv[3], "wb");
			pthread_t *tid = calloc(nthreads, sizeof(pthread_t));
			Aln_thread *first = 0;
			for (int ti = 0; ti < nthreads; ++ti) {
				Aln_thread *t = calloc(1, sizeof(Aln_thread));
				t->matched_templates = ti ? malloc(((DB_size + 1) << 1) * sizeof(int)) : matched;
				t->bestTemplates = malloc(((DB_size + 1) << 1) * sizeof(int));
				t->bestTemplates_r = malloc(((DB_size + 1) << 1) * sizeof(int));
				t->best_start_pos = malloc((DB_size << 1) * sizeof(int));
				t->best_end_pos = malloc((DB_size << 1) * sizeof(int));
				t->Lengths = malloc((DB_size << 1) * sizeof(int));
				t->alignment_scores = as; t->uniq_alignment_scores = uas;
				t->seq_indexes = seq_indexes; t->inputfile = in;
				t->frag_out_raw = fout; t->frag_out_all = 0; t->seq_in = seq_in;
				if (ti) {
					t->qseq_comp = malloc(sizeof(CompDNA)); t->qseq_r_comp = malloc(sizeof(CompDNA));
					allocComp(t->qseq_comp, 1024); allocComp(t->qseq_r_comp, 1024);
					NWmat *m = malloc(sizeof(NWmat));
					m->NW_s = 1024 * 1024; m->NW_q = 1024; m->E = malloc(m->NW_s);
					m->D[0] = malloc((m->NW_q << 1) * sizeof(int)); m->P[0] = malloc((m->NW_q << 1) * sizeof(int));
					m->D[1] = m->D[0] + m->NW_q; m->P[1] = m->P[0] + m->NW_q; m->rewards = rewards;
					t->NWmatrices = m;
					t->points = seedPoint_init(1024, rewards);
				} else { t->qseq_comp = qc; t->qseq_r_comp = qrc; t->NWmatrices = NWm; t->points = points; }
				t->qseq = setQseqs(1024); t->qseq_r = setQseqs(1024); t->header = setQseqs(256); t->header_r = setQseqs(256);
				t->kmersize = kmersize; t->minlen = 16; t->mq = 0; t->sam = 0;
				t->scoreT = 0.5; t->mrc = 0.0; t->minFrac = min_frac;
				t->template_lengths = template_lengths; t->templates_index = templates_index;
				if (ti) pthread_create(&tid[ti], 0, &alnFrags_threaded, t); else first = t;
			}
			alnFrags_threaded(first);
			for (int ti = 1; ti < nthreads; ++ti) pthread_join(tid[ti], 0);
			Aln_thread *t = first;
			fclose(t->frag_out_raw);
			FILE *so = fopen(argv[4], "wb");
			fwrite(&DB_size, 4, 1, so); fwrite(as, 8, DB_size, so); fwrite(uas, 8, DB_size, so);
			fclose(so);
		} else {
			/* per-candidate ground truth: the calls alnFragsSE makes, results recorded instead of reduced */
			FILE *co = fopen(cand_path, "wb");
			Qseqs *header = setQseqs(256);
			unsigned char *qseq = malloc(1 << 24), *qseq_r = malloc(1 << 24);
			int hdr[7], ridx = 0;
			while (fread(hdr, 4, 7, in) == 7 && hdr[0] >= 0) {
				qc->seqlen = hdr[0]; qc->complen = hdr[1];
				if (qc->size <= (unsigned)hdr[0]) { freeComp(qc); allocComp(qc, hdr[0] << 1); freeComp(qrc); allocComp(qrc, hdr[0] << 1); qc->seqlen = hdr[0]; qc->complen = hdr[1]; }
				qc->N[0] = hdr[2]; matched[0] = hdr[4]; header->len = hdr[5];
				if (header->size <= header->len) { header->size = header->len << 1; header->seq = realloc(header->seq, header->size); }
				if (fread(qc->seq, 8, qc->complen, in) != qc->complen) break;
				if (fread(qc->N + 1, 4, qc->N[0], in) != (size_t)qc->N[0]) break;
				if (fread(matched + 1, 4, matched[0], in) != (size_t)matched[0]) break;
				if (fread(header->seq, 1, header->len, in) != (size_t)header->len) break;
				int rc_flag = hdr[3], q_len = qc->seqlen;
				if (matched[0] == 0) { fprintf(stderr, "cand pass: paired records not supported\n"); return 3; }
				if (q_len >= kmersize) {
					if (rc_flag < 0) { rc_comp(qc, qrc); unCompDNA(qrc, qseq_r); qrc->N[0]++; qrc->N[qrc->N[0]] = q_len; }
					unCompDNA(qc, qseq); qc->N[0]++; qc->N[qc->N[0]] = q_len;
					int arc = rc_flag < 0;
					points->len = 0;
					for (int t_i = 1; t_i <= matched[0]; ++t_i) {
						int template = matched[t_i], at = abs(template), rc;
						AlnScore st;
						if (!templates_index[at]) templates_index[at] = alignLoadPtr(0, seq_in, template_lengths[at], kmersize, seq_indexes[at]);
						int q_start = 0, q_end = q_len;   /* q-bound exactly as alnFragsSE takes it (alnfrags.c:1091-1099) */
						if (2 * sizeof(int) + 1 < (size_t)header->len && header->seq[header->len - 2 * sizeof(int) - 1] == 0) {
							int *qb = (int *)(header->seq + (header->len - 2 * sizeof(int)));
							q_start = qb[0]; q_end = qb[1];
						}
						if (arc) {
							rc = anker_rc_comp(templates_index[at], qseq, qseq_r, qc, qrc, q_start, q_end, points);
							if (rc < 0) { template = -at; st = KMA_score(templates_index[at], qseq_r, q_len, q_len - q_end, q_len - q_start, qrc, 0, 0.5, points, NWm); }
							else if (rc) { template = at; st = KMA_score(templates_index[at], qseq, q_len, q_start, q_end, qc, 0, 0.5, points, NWm); }
							else { memset(&st, 0, sizeof(st)); points->len = 0; }
						} else if (template < 0) st = KMA_score(templates_index[at], qseq_r, q_len, q_len - q_end, q_len - q_start, qrc, 0, 0.5, points, NWm);
						else st = KMA_score(templates_index[at], qseq, q_len, q_start, q_end, qc, 0, 0.5, points, NWm);
						int row[8] = {ridx, template, st.score, st.len, st.pos, st.match, st.tGaps, st.qGaps};
						fwrite(row, 4, 8, co);
					}
				}
				++ridx;
			}
			fclose(co);
		}
		if (in != stdin) fclose(in);
	}
	return 0;
}
