/* TEST INFRASTRUCTURE ONLY -- never shipped, never linked by the product.
 * The C ABI of include/kmagpu.h (the entry points host/kmagpu_shim.c calls) answered by the CPU oracle (oracle/liborc.so),
 * so that the host shim's own logic -- record chunking, pipe handling, the KMA / anker_rc result table, the hand-over to the
 * reference's writers -- can be exercised and debugged in a container without a GPU (tests/test_host_shim.py, `-m "not gpu"`).
 * The product library is kma_b200/libkmagpu.so; the GPU test of the shim links that one. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/kmagpu.h"
#include "../../oracle/orc.h"

struct kmagpu_db { orc_db *o; char prefix[4096]; int *refs; };
static char g_err[512] = "";
const char *kmagpu_last_error(void) { return g_err; }
int kmagpu_device_count(void) { return 1; }

void kmagpu_default_params(kmagpu_params *p) {
	int i, j;
	memset(p, 0, sizeof(*p));
	p->M = 1; p->MM = -2; p->U = -1; p->W1 = -3; p->Wl = -6; p->Mn = 0; p->PE = 7;
	for (i = 0; i < 4; ++i) for (j = 0; j < 4; ++j) p->d[i * 5 + j] = i == j ? 1 : -2;
	p->scoreT = 0.5; p->minFrac = 1.0; p->minlen = 16; p->coverT = 0.1; p->counters = 1;
}

int kmagpu_db_open(const char *prefix, int device, kmagpu_db **out) {
	kmagpu_db *db = calloc(1, sizeof(*db));
	(void)device;
	snprintf(db->prefix, sizeof(db->prefix), "%s", prefix);
	if (!(db->o = orc_db_open(prefix)) || orc_db_load_seq(db->o, prefix)) { snprintf(g_err, sizeof(g_err), "cannot open %s", prefix); free(db); return -1; }
	db->refs = malloc(sizeof(int)); *db->refs = 1;
	*out = db;
	return 0;
}
int kmagpu_db_clone(kmagpu_db *src, kmagpu_db **out) {
	kmagpu_db *db = malloc(sizeof(*db));
	*db = *src; ++*db->refs; *out = db;
	return 0;
}
void kmagpu_db_close(kmagpu_db *db) { if (db && --*db->refs == 0) { orc_db_close(db->o); free(db->refs); } free(db); }
int kmagpu_db_get_info(const kmagpu_db *db, kmagpu_db_info *info) {
	memset(info, 0, sizeof(*info));
	info->DB_size = db->o->DB_size; info->kmersize = (int)db->o->kmersize; info->kmerindex = db->o->lengths ? db->o->lengths[0] : 0;
	return 0;
}

/* soft proximity sums of the run (one mock process = one run) */
static uint64_t *g_soft = 0;
int kmagpu_softproxi_reset(kmagpu_db *db) {
	free(g_soft);
	g_soft = calloc((size_t)db->o->DB_size + 3, sizeof(uint64_t));
	return g_soft ? 0 : -1;
}
int kmagpu_softproxi_download(kmagpu_db *db, uint64_t *sums) {
	if (!g_soft) { snprintf(g_err, sizeof(g_err), "kmagpu_softproxi_download before kmagpu_softproxi_reset"); return -1; }
	memcpy(sums, g_soft, sizeof(uint64_t) * (size_t)db->o->DB_size);
	return 0;
}

static void to_orc(const kmagpu_params *p, orc_params *o) {
	orc_default_params(o);
	o->M = p->M; o->MM = p->MM; o->U = p->U; o->W1 = p->W1; o->Wl = p->Wl; o->Mn = p->Mn; o->PE = p->PE;
	memcpy(o->d, p->d, sizeof(o->d));
	o->exhaustive = p->exhaustive; o->apm = p->apm;
}

int kmagpu_seed_batch(kmagpu_db *db, const kmagpu_params *p, const void *stage1, size_t nbytes, void *out, size_t cap, size_t *out_bytes,
                      int64_t *nreads, kmagpu_seed_stats *stats) {
	orc_params o;
	orc_stats st;
	int64_t n;
	(void)stats;
	to_orc(p, &o);
	memset(&st, 0, sizeof(st));
	orc_chain_set_lc(p->lc);
	orc_set_proxi(p->minFrac);
	orc_set_soft_proxi(p->minFrac < 0 && p->minFrac != -1.0 ? g_soft : 0);
	/* save_kmers_chain only sees single reads, pairs always go through save_kmers_pair (savekmers.c:196-199) */
	n = (p->kmerscan && nbytes >= 16 && ((const int *)stage1)[3] >= 0) ? orc_chain_stream(db->o, &o, stage1, nbytes, p->minlen, p->scoreT, p->coverT, p->mrc, out, cap, &st)
	                : orc_seed_stream(db->o, &o, stage1, nbytes, out, cap, &st);
	if (n < 0) { snprintf(g_err, sizeof(g_err), "oracle stage 2 failed (%lld)", (long long)n); return -1; }
	*out_bytes = (size_t)n - 4;   /* the oracle's streams end with the terminator int; kmagpu_seed_batch leaves it to the caller (kmers.c:257) */
	if (nreads) {   /* one per single read, one per pair (savekmers.c:183) */
		const unsigned char *b = stage1;
		size_t ip = 0; int64_t cnt = 0; int mate = 0;
		while (ip + 16 <= nbytes) {
			int h[4]; memcpy(h, b + ip, 16);
			ip += 16 + 8 * (size_t)h[1] + 4 * (size_t)h[2] + (size_t)abs(h[3]);
			if (mate) mate = 0; else { ++cnt; mate = h[3] < 0; }
		}
		*nreads = cnt;
	}
	return 0;
}

int kmagpu_align_batch(kmagpu_db *db, const kmagpu_params *p, const void *stage2, size_t nbytes, void *frag_out, size_t out_cap, size_t *out_bytes,
                       uint64_t *as, uint64_t *uas, kmagpu_cand *cand_out, size_t cand_cap, size_t *cand_rows, kmagpu_align_stats *stats) {
	orc_params o;
	uint8_t *fo = 0; size_t fb = 0;
	uint64_t *a = calloc(db->o->DB_size, 8), *u = calloc(db->o->DB_size, 8);
	int i;
	(void)cand_out; (void)cand_cap; (void)cand_rows; (void)stats;
	to_orc(p, &o);
	orc_align_set_minfrac(p->minFrac);
	if (orc_align_stream(db->o, db->prefix, &o, stage2, nbytes, p->one2one, p->scoreT, p->mq, p->minlen, p->mrc, &fo, &fb, a, u, 0, 0, 0)) {
		snprintf(g_err, sizeof(g_err), "oracle alignment pass failed"); return -1;
	}
	if (fb > out_cap) { snprintf(g_err, sizeof(g_err), "frag_raw needs %zu bytes", fb); return -1; }
	memcpy(frag_out, fo, fb); *out_bytes = fb;
	for (i = 0; i < db->o->DB_size; ++i) { if (as) as[i] += a[i]; if (uas) uas[i] += u[i]; }
	orc_free(fo); free(a); free(u);
	return 0;
}

int kmagpu_trace_batch(kmagpu_db *db, const kmagpu_params *p, const void *frags, size_t nbytes, void *out, size_t out_cap, size_t *out_bytes,
                       int64_t *nrecords, kmagpu_align_stats *stats) {
	orc_params o;
	uint8_t *to = 0; size_t tb = 0;
	const unsigned char *b = frags;
	size_t ip = 0; int64_t n = 0;
	(void)stats;
	to_orc(p, &o);
	orc_trace_set_ts(p->ts);
	if (orc_trace_stream(db->o, db->prefix, &o, frags, nbytes, p->one2one, p->scoreT, p->mq, p->minlen, p->mrc, &to, &tb)) {
		snprintf(g_err, sizeof(g_err), "oracle traceback pass failed"); return -1;
	}
	if (tb > out_cap) { snprintf(g_err, sizeof(g_err), "trace output needs %zu bytes", tb); return -1; }
	memcpy(out, to, tb); *out_bytes = tb;
	while (ip + 32 <= nbytes) { int h[8]; memcpy(h, b + ip, 32); if (h[0] < 0) break; ip += 32 + (size_t)h[1] + (size_t)h[6]; ++n; }
	if (nrecords) *nrecords = n;
	orc_free(to);
	return 0;
}
