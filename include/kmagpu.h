/*
 * kmagpu.h -- C ABI of the B200-native KMA mapping core (libkmagpu.so).
 *
 * Drop-in boundary: KMA 1.5.1 has no FFI; its "plugin API" is a set of global C
 * function pointers invoked once per read from T pthreads. A GPU needs batches, so the seams
 * this library binds to are the reference's three record streams. Every entry point below cites
 * the reference interface it replaces; INTEGRATION.md shows the patch a KMA maintainer applies.
 *
 * Conventions (mirroring pherror.h / the reference's ownership rules):
 *   - plain pointers and sizes only; the caller owns every buffer, the library never frees or
 *     reallocs caller memory;
 *   - every function returns 0 on success, non-zero on failure; kmagpu_last_error() returns a
 *     thread-local message (the host shim turns that into the reference's print + exit(errno));
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef KMAGPU_H
#define KMAGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kmagpu_db kmagpu_db;

/* POD copy of `Penalties` (penalties.h:22-33) + the CLI scalars the hot path reads. */
typedef struct kmagpu_params {
	int32_t M, MM, U, W1, Wl, Mn, PE; /* kma.c:327-336, 1308-1328 */
	int32_t d[25];                    /* rewards->d[t][q], row-major 5x5 */
	int32_t exhaustive;               /* -ex_mode (kma.c:551) */
	int32_t mq;                       /* -mq  minimum mapQ (align.c:658) */
	int32_t one2one;                  /* -1t1 */
	int32_t reserved[5];
	double scoreT;                    /* -mrs (alnfrags.c:1168) */
	double minFrac;
} kmagpu_params;

typedef struct kmagpu_db_info {
	int32_t DB_size;     /* templates + 1 (template ids are 1-based) */
	int32_t kmersize;    /* k of the .comp.b hash */
	int32_t kmerindex;   /* k of the per-template alignment index (.length.b[0]) */
	int32_t mega;        /* direct-addressed table (megaMap_getGlobal) */
	uint64_t size, n, v_index;
	uint64_t device_bytes; /* HBM held by this database */
	uint64_t seq_bases;    /* total template bases */
} kmagpu_db_info;

/* counters of the last seeding call -- the algorithmic-bytes inputs of SURVEY.md §8d */
typedef struct kmagpu_seed_stats {
	int64_t reads, mapped, read_words;
	int64_t lookups, hits, list_fetches, list_ids;
	int64_t overflow_reads;   /* reads that took the dense-scratch path */
	float ms_seed, ms_emit;   /* CUDA-event time of the scoring kernels / record writer */
	float ms_h2d, ms_total;   /* H2D copy; whole device-side step (first kernel start -> writer end) */
	int32_t launches;         /* kernels launched by the call */
	int32_t reserved;
} kmagpu_seed_stats;

void kmagpu_default_params(kmagpu_params *p);
const char *kmagpu_last_error(void);
int kmagpu_device_count(void);

/* Replaces hashMapKMA_load (hashmapkma.c:275) + the .length.b/.seq.b loads of runKMA
 * (runkma.c:161-220): reads <prefix>.comp.b/.length.b/.seq.b and makes them HBM resident on
 * `device`. */
int kmagpu_db_open(const char *prefix, int device, kmagpu_db **out);
void kmagpu_db_close(kmagpu_db *db);
int kmagpu_db_get_info(const kmagpu_db *db, kmagpu_db_info *info);

/* Replaces the per-read loop of save_kmers_threaded (savekmers.c:94-271) with kmerScan =
 * save_kmers (-1t1, savekmers.c:2442): consumes `nbytes` of whole stage-1 records
 * (runinput.c:765-787 printFsa layout) and writes the stage-2 records (ankers.c:30-50
 * print_ankers layout) of the reads that map, in input order (= the reference's `-t 1` order).
 * The stream terminator (kmers.c:257) is NOT written; the caller appends -(sum of *nreads). */
int kmagpu_seed_batch(kmagpu_db *db, const kmagpu_params *p, const void *stage1, size_t nbytes,
                      void *stage2_out, size_t out_cap, size_t *out_bytes, int64_t *nreads,
                      kmagpu_seed_stats *stats);

/* The same call split in three so a caller (bench.py) can time the kernels with the batch
 * already resident in HBM: upload = parse + H2D, run = kernels only, download = D2H. */
int kmagpu_seed_upload(kmagpu_db *db, const void *stage1, size_t nbytes, int64_t *nreads);
int kmagpu_seed_run(kmagpu_db *db, const kmagpu_params *p, kmagpu_seed_stats *stats);
int kmagpu_seed_download(kmagpu_db *db, void *stage2_out, size_t out_cap, size_t *out_bytes);

/* hashMap_get (hashmapkma.h:58; hashMap_getGlobal hashmapkma.c:149 / megaMap_getGlobal :264) over
 * a batch of k-mers: out[i] = offset of the template list inside values[], or -1. Test hook. */
int kmagpu_lookup_batch(kmagpu_db *db, const uint64_t *kmers, size_t n, int64_t *out);

#ifdef __cplusplus
}
#endif
#endif
