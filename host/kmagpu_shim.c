/*
 * kmagpu_shim.c -- the KMA 1.5.1 host with its mapping core on the GPU.
 *
 * This file is linked with the UNMODIFIED object files of the reference (oracle/Makefile.host builds them from
 * /root/reference; nothing of the reference is patched or copied) and with libkmagpu.so. The reference has no plugin
 * API: its hot path is reached through plain function symbols. The GNU linker's --wrap rebinds exactly those symbols:
 *
 *   save_kmers_batch   (kmers.c:51)        stage 2: the per-read loop of save_kmers_threaded (savekmers.c:94-271)
 *                                          becomes kmagpu_seed_batch over chunks of the stage-1 pipe;
 *   alnFrags_threaded  (alnfrags.c:2150)   stage 3a: get_ankers + alnFragsSE / alnFragsPE + update_Scores become
 *                                          kmagpu_align_batch over chunks of the stage-2 pipe;
 *   runKMA, runKMA_MEM (runkma.c:130/909)  only to re-point assembly_KMA_Ptr at shim_assemble below;
 *   KMA, anker_rc      (align.c:214/780)   the traceback alignment of assemble_KMA (assembly.c:1925-1934): answered
 *                                          from the results kmagpu_trace_batch produced for the template's fragments.
 *
 * Everything else -- option parsing, stage 1, ConClave, alnToMat with its insertion nodes, callConsensus, every writer
 * (.res .fsa .aln .frag.gz .mat.gz ...) -- is the reference's own code, so the files come out of the reference's writers.
 * Configurations the GPU path does not cover end with an error message and exit(1); nothing falls back to the CPU path.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "align.h"
#include "alnfrags.h"
#include "ankers.h"
#include "assembly.h"
#include "chain.h"
#include "hashmapcci.h"
#include "kmapipe.h"
#include "kmeranker.h"
#include "kmers.h"
#include "penalties.h"
#include "pherror.h"
#include "runkma.h"
#include "savekmers.h"

#include "kmagpu.h"

/* ------------------------------------------------------------------ shared state */

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static kmagpu_db *g_db_seed, *g_db_aln;   /* two handles on ONE HBM image: stage 2 and stage 3 run in different threads */
static int g_one2one;                     /* -1t1, as the CLI hands it to anker_rc (kma.c:1428) */
static double g_coverT = 0.1;

#include <execinfo.h>
static void shim_atexit(void) { void *bt[32]; int n = backtrace(bt, 32); fprintf(stderr, "[shim] exit called from:\n"); backtrace_symbols_fd(bt, n, 2); }
static int shim_debug(void) { static int d = -1; if (d < 0) { d = getenv("KMAGPU_DEBUG") != 0; if (d) atexit(shim_atexit); } return d; }
#define SHIM_TRACE(...) do { if (shim_debug()) { fprintf(stderr, "[shim] " __VA_ARGS__); fputc('\n', stderr); } } while (0)

/* _exit, not exit: the stages are threads of one process (kmaPipeThread), and exit() would wait for the stdio locks of the
   pipes another stage is blocked on */
static void shim_die(const char *what) {
	fprintf(stderr, "kma (GPU host): %s: %s\n", what, kmagpu_last_error());
	fflush(stderr);
	_exit(1);
}

static void shim_unsupported(const char *what) {
	fprintf(stderr, "kma (GPU host): %s is not covered by the GPU mapping core; there is no CPU fallback.\n", what);
	fflush(stderr);
	_exit(1);
}

static int shim_device(void) {
	const char *e = getenv("KMAGPU_DEVICE");
	return e ? atoi(e) : 0;
}

/* the first caller loads the database image, the other one gets a clone of the handle */
static kmagpu_db *shim_db(const char *prefix, int stage) {
	kmagpu_db **slot = stage == 2 ? &g_db_seed : &g_db_aln, **other = stage == 2 ? &g_db_aln : &g_db_seed;
	pthread_mutex_lock(&g_lock);
	if (!*slot) {
		if (*other) { if (kmagpu_db_clone(*other, slot)) shim_die("kmagpu_db_clone"); }
		else if (kmagpu_db_open(prefix, shim_device(), slot)) shim_die("kmagpu_db_open");
	}
	pthread_mutex_unlock(&g_lock);
	return *slot;
}

static void shim_params(kmagpu_params *p, const Penalties *rewards) {
	int i, j;
	kmagpu_default_params(p);
	p->M = rewards->M; p->MM = rewards->MM; p->U = rewards->U; p->W1 = rewards->W1; p->Wl = rewards->Wl;
	p->Mn = rewards->Mn; p->PE = rewards->PE;
	for (i = 0; i < 5; ++i) for (j = 0; j < 5; ++j) p->d[i * 5 + j] = rewards->d[i][j];
	p->one2one = g_one2one;
	p->counters = 0;
	p->coverT = g_coverT;
	if (chainSeedsPtr != &chainSeeds) shim_unsupported("circular chaining (-ca)");
	if (leadTailAlnPtr != &leadTailAln || trailTailAlnPtr != &trailTailAln) shim_unsupported("-ssa (skipped tail alignments)");
}

/* the -ts value lives in a static of trimSeeds (chain.c:496): read it back with a one-seed probe chain */
static int shim_trim(void) {
	int tS = 1, tE = 65, qS = 1, qE = 65, nx = 0;
	AlnPoints pts;
	memset(&pts, 0, sizeof(pts));
	pts.tStart = &tS; pts.tEnd = &tE; pts.qStart = &qS; pts.qEnd = &qE; pts.next = &nx;
	pts.len = 1;
	if (trimSeedsPtr != &trimSeeds) shim_unsupported("-ssa (trimSeedsNoLead)");
	trimSeedsPtr(&pts, 0);
	return qS - 1;
}

/* read exactly n bytes; 0 at a clean end of stream before the first byte */
static int shim_read(void *dst, size_t n, FILE *f) {
	size_t got = fread(dst, 1, n, f);
	if (got == n) return 1;
	if (got == 0) return 0;
	fprintf(stderr, "kma (GPU host): record stream ends inside a record\n");
	exit(1);
}

static void *shim_grow(void *p, size_t *cap, size_t need) {
	if (need <= *cap) return p;
	*cap = need + need / 2 + 4096;
	p = realloc(p, *cap);
	if (!p) { ERROR(); }
	return p;
}

/* ------------------------------------------------------------------ stage 2 */

#define SHIM_CHUNK (64u << 20)   /* bytes of records handed to the device per call */

int __wrap_save_kmers_batch(char *templatefilename, char *exePrev, unsigned shm, int thread_num, const int exhaustive, Penalties *rewards,
                            FILE *out, int sam, int minlen, double mrs, double coverT, double minFrac) {
	kmagpu_params prm;
	kmagpu_db *db;
	FILE *in;
	unsigned char *buf = 0, *obuf = 0;
	size_t cap = 0, ocap = 0, fill = 0, obytes;
	int64_t n, total = 0;
	int status = 0, pending_mate = 0, eof = 0, hdr[4], have_hdr = 0, chunk_pairs = 0;
	(void)shm; (void)thread_num;

	SHIM_TRACE("save_kmers_batch(%s)", templatefilename);
	if (!(in = kmaPipe(exePrev, "rb", 0, 0))) { ERROR(); }
	SHIM_TRACE("pipe open");
	if (deConPrintPtr == &deConPrint) shim_unsupported("-decon");
	if (printPtr != &print_ankers) shim_unsupported("sparse / split databases");
	if (sam == 1 && out != stdout) shim_unsupported("SAM output of unmapped reads");
	g_coverT = coverT;
	SHIM_TRACE("checks done");
	shim_params(&prm, rewards);
	SHIM_TRACE("params done");
	prm.exhaustive = exhaustive; prm.minlen = minlen; prm.scoreT = mrs; prm.coverT = coverT;
	prm.minFrac = minFrac;   /* -proxi (kma.c:702-718): stage 2 is handed |minFrac| (kma.c:1605) */
	if (kmerScan == &save_kmers) prm.kmerscan = 0;
	else if (kmerScan == &save_kmers_chain) prm.kmerscan = 1;
	else shim_unsupported("this k-mer scan (-hmm / -Sparse / count modes)");
	prm.lc = kmerAnkerScore != &ankerScore;   /* -lc (kma.c:694-700) */
	if (save_kmers_pair == &save_kmers_unionPair) prm.apm = 1;
	else if (save_kmers_pair == &save_kmers_penaltyPair) prm.apm = 0;
	else shim_unsupported("-apm f");
	if (get_kmers_for_pair_ptr != &get_kmers_for_pair) shim_unsupported("this pair scan");
	SHIM_TRACE("stage 2: kmerscan %d apm %d", prm.kmerscan, prm.apm);
	db = shim_db(templatefilename, 2);
	/* soft proximity (kmers.c:133-153): a negative minFrac reaches this function only in -mem_mode (kma.c:1605) */
	if (minFrac < 0 && minFrac != -1.0 && kmagpu_softproxi_reset(db)) shim_die("kmagpu_softproxi_reset");
	fprintf(stderr, "# Finding k-mer ankers (GPU)\n");

	while (!eof || fill || have_hdr) {
		/* fill the chunk with whole records (loadFsa, savekmers.c:50-92); a first mate carries a negative header length
		   and its mate follows (runinput.c:789): never cut between the two */
		while (!eof && (fill < SHIM_CHUNK || pending_mate)) {
			size_t len;
			if (have_hdr) have_hdr = 0;
			else if (!shim_read(hdr, sizeof(hdr), in)) { eof = 1; break; }
			/* kmerScan = save_kmers_chain only sees single reads, pairs go through save_kmers_pair (savekmers.c:196-199):
			   in chain mode a chunk holds one kind of record */
			if (prm.kmerscan && fill && !pending_mate && (hdr[3] < 0) != chunk_pairs) { have_hdr = 1; break; }
			if (!fill) chunk_pairs = hdr[3] < 0;
			len = 8 * (size_t)hdr[1] + 4 * (size_t)hdr[2] + (size_t)abs(hdr[3]);
			buf = shim_grow(buf, &cap, fill + 16 + len);
			memcpy(buf + fill, hdr, 16);
			if (len && !shim_read(buf + fill + 16, len, in)) { fprintf(stderr, "kma (GPU host): truncated stage-1 stream\n"); _exit(1); }
			fill += 16 + len;
			pending_mate = !pending_mate && hdr[3] < 0;
		}
		if (!fill) break;
		/* a stage-2 record is its stage-1 record + 12 header bytes + 4 bytes per template; chain mode may emit several
		   per read */
		obuf = shim_grow(obuf, &ocap, 4 * fill + (64u << 20));
		SHIM_TRACE("stage 2: %zu bytes of records -> device", fill);
		if (kmagpu_seed_batch(db, &prm, buf, fill, obuf, ocap, &obytes, &n, 0)) shim_die("kmagpu_seed_batch");
		SHIM_TRACE("stage 2: %lld reads, %zu bytes of stage-2 records", (long long)n, obytes);
		sfwrite(obuf, 1, obytes, out);
		total += n;
		fill = 0;
	}
	/* number of fragments, negated: the terminating record (kmers.c:257) */
	sfwrite(&(int){-(int)total}, sizeof(int), 1, out);
	if (minFrac < 0 && minFrac != -1.0) {   /* the sums travel behind the stream: their first 6 ints, then all of them (kmers.c:151-153) */
		kmagpu_db_info info;
		long unsigned *soft;
		if (kmagpu_db_get_info(db, &info)) shim_die("kmagpu_db_get_info");
		soft = calloc((size_t)info.DB_size + 3, sizeof(long unsigned));
		if (!soft) { ERROR(); }
		if (kmagpu_softproxi_download(db, (uint64_t *)soft)) shim_die("kmagpu_softproxi_download");
		sfwrite(soft, sizeof(int), 6, out);
		sfwrite(soft, sizeof(long unsigned), info.DB_size, out);
		free(soft);
	}
	kmaPipe(0, 0, in, &status);
	fprintf(stderr, "# Query ankered\n#\n");
	free(buf); free(obuf);
	return status;
}

/* ------------------------------------------------------------------ stage 3a */

void *__wrap_alnFrags_threaded(void *arg) {
	static int taken = 0;
	Aln_thread *thr = arg;
	kmagpu_params prm;
	kmagpu_db *db;
	kmagpu_db_info info;
	unsigned char *buf = 0, *obuf = 0;
	size_t cap = 0, ocap = 0, fill = 0, obytes;
	int hdr[7], eof = 0, pending_mate = 0, nfrags = 0, t, maxq = 0, maxh = 0;
	FILE *in = thr->inputfile;

	/* the reference runs this function on T threads that pull records one by one; a batch needs one consumer */
	pthread_mutex_lock(&g_lock);
	if (taken) { pthread_mutex_unlock(&g_lock); return NULL; }
	taken = 1;
	pthread_mutex_unlock(&g_lock);

	SHIM_TRACE("alnFrags_threaded");
	if (thr->sam) shim_unsupported("-sam");
	if (thr->frag_out_all) shim_unsupported("-a (all fragments)");
	shim_params(&prm, thr->NWmatrices->rewards);
	prm.minlen = thr->minlen; prm.mq = thr->mq; prm.scoreT = thr->scoreT; prm.mrc = thr->mrc; prm.minFrac = thr->minFrac;
	if (alnFragsPE == &alnFragsUnionPE) prm.apm = 1;
	else if (alnFragsPE == &alnFragsPenaltyPE) prm.apm = 0;
	else shim_unsupported("-apm f");
	/* the database prefix: stage 2 opened it already (same process, kmaPipeThread); if not, it is not known here */
	db = g_db_aln ? g_db_aln : (g_db_seed ? shim_db(0, 3) : 0);
	if (!db) { fprintf(stderr, "kma (GPU host): the alignment pass needs the database stage 2 opened (forked stages are not supported)\n"); _exit(1); }
	kmagpu_db_get_info(db, &info);
	if (info.kmerindex != thr->kmersize) shim_unsupported("-k different from the index' k");

	while (!eof || fill) {
		while (!eof && (fill < SHIM_CHUNK || pending_mate)) {   /* get_ankers (ankers.c:163-220) */
			size_t len;
			if (!shim_read(hdr, 4, in)) { eof = 1; break; }
			if (hdr[0] < 0) { nfrags = -hdr[0]; eof = 1; break; }
			if (!shim_read(hdr + 1, 24, in)) { fprintf(stderr, "kma (GPU host): truncated stage-2 stream\n"); _exit(1); }
			len = 8 * (size_t)hdr[1] + 4 * (size_t)hdr[2] + 4 * (size_t)hdr[4] + (size_t)hdr[5];
			buf = shim_grow(buf, &cap, fill + 28 + len);
			memcpy(buf + fill, hdr, 28);
			if (len && !shim_read(buf + fill + 28, len, in)) { fprintf(stderr, "kma (GPU host): truncated stage-2 stream\n"); _exit(1); }
			fill += 28 + len;
			pending_mate = hdr[4] == 0;   /* the first record of a pair has no templates (ankers.c:150) */
			if (maxq < hdr[0]) maxq = hdr[0];
			if (maxh < hdr[5]) maxh = hdr[5];
		}
		if (!fill) break;
		obuf = shim_grow(obuf, &ocap, 5 * fill + (16u << 20));
		SHIM_TRACE("alignment pass: %zu bytes of stage-2 records -> device", fill);
		if (kmagpu_align_batch(db, &prm, buf, fill, obuf, ocap, &obytes, (uint64_t *)thr->alignment_scores, (uint64_t *)thr->uniq_alignment_scores,
		                       0, 0, 0, 0)) shim_die("kmagpu_align_batch");
		SHIM_TRACE("alignment pass: %zu bytes of frag_raw", obytes);
		sfwrite(obuf, 1, obytes, thr->frag_out_raw);
		fill = 0;
	}
	/* ConClave and the assembly read the longest read / name into these buffers without a size check: the reference grows
	   them while it parses the records (alnfrags.c:2226-2235, ankers.c:199-206) */
	if (thr->qseq->size <= maxq) {
		free(thr->qseq->seq); free(thr->qseq_r->seq);
		thr->qseq->size = thr->qseq_r->size = maxq << 1;
		thr->qseq->seq = smalloc(thr->qseq->size); thr->qseq_r->seq = smalloc(thr->qseq_r->size);
	}
	if (thr->header->size <= maxh) {
		free(thr->header->seq); free(thr->header_r->seq);
		thr->header->size = thr->header_r->size = maxh << 1;
		thr->header->seq = smalloc(thr->header->size); thr->header_r->seq = smalloc(thr->header_r->size);
	}
	SHIM_TRACE("alnFrags_threaded: %d fragments", nfrags);
	*thr->matched_templates = nfrags;   /* what get_ankers leaves behind at the end of the stream (ankers.c:168) */
	/* the assembly reads seq / len / kmerindex of every template that took part (runkma.c:784, 813; assembly.c:2065):
	   alnFragsSE would have loaded them with alignLoadPtr (alnfrags.c:1083-1089). The position index itself stays on the GPU. */
	for (t = 1; t < info.DB_size; ++t) {
		if (thr->alignment_scores[t] && !thr->templates_index[t]) {
			HashMapCCI *ix = calloc(1, sizeof(HashMapCCI));
			const size_t bytes = ((size_t)(thr->template_lengths[t] >> 5) + 1) * sizeof(long unsigned);
			if (!ix || !(ix->seq = malloc(bytes))) { ERROR(); }
			ix->len = thr->template_lengths[t];
			ix->kmerindex = thr->kmersize;
			if (pread(thr->seq_in, ix->seq, bytes, thr->seq_indexes[t]) != (ssize_t)bytes) { fprintf(stderr, "Corrupted *.seq.b\n"); _exit(1); }
			thr->templates_index[t] = ix;
		}
	}
	free(buf); free(obuf);
	return NULL;
}

/* ------------------------------------------------------------------ stage 3b: traceback alignment inside assemble_KMA */

/* results of kmagpu_trace_batch for the next fragments of the current template, in file order */
typedef struct {
	const unsigned char *read;   /* the fragment's bytes as stored in the file */
	const int32_t *h;            /* int32[12]: accepted read_score start end | score len pos match tGaps qGaps | turned ncol */
	const unsigned char *rows;   /* t, s, q: ncol bytes each */
	int q_len, used;
} ShimFrag;

static struct {
	void *(*real)(void *);       /* the reference's assemble_KMA / assemble_KMA_dense ... */
	int template;
	FILE **files;                /* the per-template fragment files (frags.c:30-61), each sorted by template */
	int file_count, file_i;
	off_t *pos;                  /* our read cursor in each file (pread: the reference's own FILE position is untouched) */
	char *done;
	unsigned char *in[2], *out[2];   /* two chunks stay alive: with T threads a record of the previous chunk may still be in flight */
	size_t in_cap[2], out_cap[2];
	ShimFrag *frag[2];
	size_t nfrag[2], frag_cap[2], next[2];
	int cur;
	kmagpu_params prm;
	int ready;
} g_tr;

static __thread const ShimFrag *tl_pending;   /* anker_rc found this fragment; the KMA call that follows uses it */

/* load and align the next chunk of the current template's fragments; 0 when there are none left */
static int shim_next_chunk(void) {
	const int c = g_tr.cur ^ 1;
	size_t fill = 0, n = 0, obytes, off, i;
	int64_t nrec;
	int32_t h[8];
	while (g_tr.file_i < g_tr.file_count && fill < SHIM_CHUNK / 2) {
		const int f = g_tr.file_i;
		int fd;
		if (!g_tr.files[f] || g_tr.done[f]) { ++g_tr.file_i; continue; }
		fd = fileno(g_tr.files[f]);
		if (pread(fd, h, 32, g_tr.pos[f]) != 32 || h[0] == -1 || h[0] > g_tr.template) { g_tr.done[f] = 1; ++g_tr.file_i; continue; }
		if (h[0] == g_tr.template) {
			const size_t len = 32 + (size_t)h[1] + (size_t)h[6];
			g_tr.in[c] = shim_grow(g_tr.in[c], &g_tr.in_cap[c], fill + len);
			if (pread(fd, g_tr.in[c] + fill, len, g_tr.pos[f]) != (ssize_t)len) { fprintf(stderr, "kma (GPU host): truncated fragment file\n"); _exit(1); }
			fill += len; ++n;
		}
		g_tr.pos[f] += 32 + (off_t)h[1] + (off_t)h[6];
	}
	if (!n) return 0;
	/* per fragment 48 bytes + three rows of at most 3 * q_len + 256 columns (the library's row capacity) */
	g_tr.out[c] = shim_grow(g_tr.out[c], &g_tr.out_cap[c], 9 * fill + 1024 * n + 4096);
	SHIM_TRACE("assembly: template %d, %zu fragments (%zu bytes) -> device", g_tr.template, n, fill);
	if (kmagpu_trace_batch(g_db_aln, &g_tr.prm, g_tr.in[c], fill, g_tr.out[c], g_tr.out_cap[c], &obytes, &nrec, 0)) shim_die("kmagpu_trace_batch");
	if ((size_t)nrec != n) { fprintf(stderr, "kma (GPU host): fragment count mismatch\n"); _exit(1); }
	if (g_tr.frag_cap[c] < n) {
		g_tr.frag_cap[c] = n + n / 2 + 64;
		g_tr.frag[c] = realloc(g_tr.frag[c], g_tr.frag_cap[c] * sizeof(ShimFrag));
		if (!g_tr.frag[c]) { ERROR(); }
	}
	for (i = 0, fill = 0, off = 0; i < n; ++i) {
		ShimFrag *fr = g_tr.frag[c] + i;
		const int32_t *rh = (const int32_t *)(g_tr.in[c] + fill);
		fr->read = g_tr.in[c] + fill + 32; fr->q_len = rh[1]; fr->used = 0;
		fr->h = (const int32_t *)(g_tr.out[c] + off);
		fr->rows = g_tr.out[c] + off + 48;
		fill += 32 + (size_t)rh[1] + (size_t)rh[6];
		off += 48 + 3 * (size_t)fr->h[11];
	}
	g_tr.nfrag[c] = n; g_tr.next[c] = 0;
	g_tr.cur = c;
	return 1;
}

/* the result for this read: the next unused fragment of the template with these bytes (file order = call order with one
   thread; with several threads the calls arrive slightly out of order, hence the search) */
static const ShimFrag *shim_lookup(const unsigned char *qseq, int q_len) {
	int pass, c;
	size_t i;
	const ShimFrag *hit = 0;
	pthread_mutex_lock(&g_lock);
	for (pass = 0; pass < 64 && !hit; ++pass) {
		for (c = g_tr.cur ^ 1; !hit; c ^= 1) {   /* older chunk first */
			ShimFrag *fr = g_tr.frag[c];
			while (g_tr.next[c] < g_tr.nfrag[c] && fr[g_tr.next[c]].used) ++g_tr.next[c];
			for (i = g_tr.next[c]; i < g_tr.nfrag[c]; ++i)
				if (!fr[i].used && fr[i].q_len == q_len && !memcmp(fr[i].read, qseq, (size_t)q_len)) { fr[i].used = 1; hit = fr + i; break; }
			if (c == g_tr.cur) break;
		}
		if (!hit && !shim_next_chunk()) break;
	}
	pthread_mutex_unlock(&g_lock);
	if (!hit) { fprintf(stderr, "kma (GPU host): a fragment of template %d reached KMA without a device result\n", g_tr.template); _exit(1); }
	return hit;
}

int __real_anker_rc(const HashMapCCI *template_index, unsigned char *qseq, int q_len, int q_start, int q_end, AlnPoints *points);

int __wrap_anker_rc(const HashMapCCI *template_index, unsigned char *qseq, int q_len, int q_start, int q_end, AlnPoints *points) {
	const ShimFrag *fr;
	int i;
	if (!template_index) {   /* the CLI's setter call (kma.c:1428) */
		g_one2one = q_len;
		return __real_anker_rc(template_index, qseq, q_len, q_start, q_end, points);
	}
	if (!g_tr.ready) return __real_anker_rc(template_index, qseq, q_len, q_start, q_end, points);
	fr = shim_lookup(qseq, q_len);
	if (fr->h[10]) {   /* the read ends up reverse-complemented (strrc, align.c:799 / 973) */
		for (i = 0; i < q_len / 2; ++i) {
			const unsigned char a = qseq[i], b = qseq[q_len - 1 - i];
			qseq[i] = b < 4 ? 3 - b : b; qseq[q_len - 1 - i] = a < 4 ? 3 - a : a;
		}
		if (q_len & 1) { const unsigned char a = qseq[q_len / 2]; qseq[q_len / 2] = a < 4 ? 3 - a : a; }
	}
	points->len = 0;
	if (!fr->h[5]) return 0;   /* no strand seeded: KMA is not called (a KMA result always has len >= 1) */
	tl_pending = fr;
	return 1;
}

AlnScore __real_KMA(const HashMapCCI *template_index, const unsigned char *qseq, int q_len, int q_start, int q_end, Aln *aligned,
                    Aln *Frag_align, int min, int max, int mq, double scoreT, AlnPoints *points, NWmat *matrices);

AlnScore __wrap_KMA(const HashMapCCI *template_index, const unsigned char *qseq, int q_len, int q_start, int q_end, Aln *aligned,
                    Aln *Frag_align, int min, int max, int mq, double scoreT, AlnPoints *points, NWmat *matrices) {
	const ShimFrag *fr = tl_pending;
	AlnScore st;
	int ncol;
	if (!g_tr.ready) return __real_KMA(template_index, qseq, q_len, q_start, q_end, aligned, Frag_align, min, max, mq, scoreT, points, matrices);
	tl_pending = 0;
	if (!fr) fr = shim_lookup(qseq, q_len);
	st.score = fr->h[4]; st.len = fr->h[5]; st.pos = fr->h[6]; st.match = fr->h[7]; st.tGaps = fr->h[8]; st.qGaps = fr->h[9];
	ncol = fr->h[11];
	aligned->start = 0; aligned->end = 0; aligned->mapQ = 0;
	if (ncol) {
		memcpy(aligned->t, fr->rows, (size_t)ncol);
		memcpy(aligned->s, fr->rows + ncol, (size_t)ncol);
		memcpy(aligned->q, fr->rows + 2 * (size_t)ncol, (size_t)ncol);
	}
	aligned->s[ncol] = 0;
	aligned->len = ncol;
	points->len = 0;
	return st;
}

/* assembly_KMA_Ptr points here: prepare the template's fragments, then run the reference's assembly function */
static void *shim_assemble(void *arg) {
	Assemble_thread *thr = arg;
	int f;
	if (thr->num == 0 && thr->template >= 0) {   /* the main thread enters once per template (runkma.c:793) */
		if (thr->sam) shim_unsupported("-sam");
		if (thr->xml_out) shim_unsupported("-xml");
		pthread_mutex_lock(&g_lock);
		g_tr.template = thr->template;
		g_tr.files = thr->files; g_tr.file_i = 0;
		if (g_tr.file_count < thr->file_count) {
			g_tr.pos = realloc(g_tr.pos, thr->file_count * sizeof(off_t));
			g_tr.done = realloc(g_tr.done, thr->file_count);
			if (!g_tr.pos || !g_tr.done) { ERROR(); }
		}
		g_tr.file_count = thr->file_count;
		for (f = 0; f < thr->file_count; ++f) {
			g_tr.done[f] = thr->files[f] == 0;
			if (thr->files[f]) g_tr.pos[f] = ftello(thr->files[f]);
		}
		g_tr.nfrag[0] = g_tr.nfrag[1] = 0; g_tr.next[0] = g_tr.next[1] = 0;
		shim_params(&g_tr.prm, thr->NWmatrices->rewards);
		g_tr.prm.minlen = thr->minlen; g_tr.prm.mq = thr->mq; g_tr.prm.scoreT = thr->scoreT; g_tr.prm.mrc = thr->mrc;
		g_tr.prm.ts = shim_trim();
		if (!g_db_aln) { fprintf(stderr, "kma (GPU host): the assembly needs the database of the alignment pass\n"); _exit(1); }
		g_tr.ready = 1;
		pthread_mutex_unlock(&g_lock);
	}
	return g_tr.real(arg);
}

static void shim_hook_assembly(void) {
	if (assembly_KMA_Ptr != &shim_assemble && assembly_KMA_Ptr != &skip_assemble_KMA) {
		g_tr.real = assembly_KMA_Ptr;
		assembly_KMA_Ptr = &shim_assemble;
	}
}

int __real_runKMA(char *templatefilename, char *outputfilename, char *exePrev, int ConClave, int kmersize, int minlen, Penalties *rewards,
                  int extendedFeatures, double ID_t, double Depth_t, int mq, double scoreT, double mrc, double minFrac, double evalue,
                  double support, int bcd, int ref_fsa, int print_matrix, int print_all, long unsigned tsv, int vcf, int xml, int sam, int nc,
                  int nf, unsigned shm, int thread_num, int maxFrag, int verbose);

int __wrap_runKMA(char *templatefilename, char *outputfilename, char *exePrev, int ConClave, int kmersize, int minlen, Penalties *rewards,
                  int extendedFeatures, double ID_t, double Depth_t, int mq, double scoreT, double mrc, double minFrac, double evalue,
                  double support, int bcd, int ref_fsa, int print_matrix, int print_all, long unsigned tsv, int vcf, int xml, int sam, int nc,
                  int nf, unsigned shm, int thread_num, int maxFrag, int verbose) {
	SHIM_TRACE("runKMA(%s, %s)", templatefilename, outputfilename);
	if (kmaPipe != &kmaPipeThread) shim_unsupported("-status (forked stages)");
	shim_db(templatefilename, 3);
	shim_hook_assembly();
	return __real_runKMA(templatefilename, outputfilename, exePrev, ConClave, kmersize, minlen, rewards, extendedFeatures, ID_t, Depth_t, mq, scoreT,
	                     mrc, minFrac, evalue, support, bcd, ref_fsa, print_matrix, print_all, tsv, vcf, xml, sam, nc, nf, shm, thread_num, maxFrag,
	                     verbose);
}

int __real_runKMA_MEM(char *templatefilename, char *outputfilename, char *exePrev, int ConClave, int kmersize, int minlen, Penalties *rewards,
                      int extendedFeatures, double ID_t, double Depth_t, int mq, double scoreT, double mrc, double minFrac, double evalue,
                      double support, int bcd, int ref_fsa, int print_matrix, int print_all, long unsigned tsv, int vcf, int xml, int sam,
                      int nc, int nf, unsigned shm, int thread_num, int maxFrag, int verbose);

/* -mem_mode: stage 2 is the wrapped save_kmers_batch; the k-mer score collection (runkma.c:1088-1140) is a per-record sum
   and stays the reference's; the assembly's alignments come from the device like runKMA's */
int __wrap_runKMA_MEM(char *templatefilename, char *outputfilename, char *exePrev, int ConClave, int kmersize, int minlen, Penalties *rewards,
                      int extendedFeatures, double ID_t, double Depth_t, int mq, double scoreT, double mrc, double minFrac, double evalue,
                      double support, int bcd, int ref_fsa, int print_matrix, int print_all, long unsigned tsv, int vcf, int xml, int sam,
                      int nc, int nf, unsigned shm, int thread_num, int maxFrag, int verbose) {
	SHIM_TRACE("runKMA_MEM(%s, %s)", templatefilename, outputfilename);
	if (kmaPipe != &kmaPipeThread) shim_unsupported("-status (forked stages)");
	shim_db(templatefilename, 3);
	shim_hook_assembly();
	return __real_runKMA_MEM(templatefilename, outputfilename, exePrev, ConClave, kmersize, minlen, rewards, extendedFeatures, ID_t, Depth_t, mq,
	                         scoreT, mrc, minFrac, evalue, support, bcd, ref_fsa, print_matrix, print_all, tsv, vcf, xml, sam, nc, nf, shm, thread_num,
	                         maxFrag, verbose);
}

/* -mem_mode loads a template when its assembly starts (assembly.c:1806): the sequence is all the host still needs, the
   k-mer position index (hashMapCCI_add over every position, the slow part of hashmapcci.c:512-605) lives on the device */
HashMapCCI *__wrap_hashMapCCI_load_thread(HashMapCCI *src, int seq, int len, int kmersize, int thread_num) {
	(void)thread_num;
	pthread_mutex_lock(&g_lock);
	if (src->len == 0) {
		const long check = (((long)len >> 5) + 1) * (long)sizeof(long unsigned);
		hashMapCCI_initialize(src, len, kmersize);
		if (read(seq, src->seq, check) != check) { fprintf(stderr, "Corrupted *.seq.b\n"); _exit(1); }
	}
	pthread_mutex_unlock(&g_lock);
	return src;
}
