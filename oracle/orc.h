/* TEST INFRASTRUCTURE ONLY -- shared declarations of the CPU oracle (see orc_seed.c / orc_align.c). */
#ifndef ORC_H
#define ORC_H
#include <stddef.h>
#include <stdint.h>

typedef struct {
	/* header of .comp.b (hashmapkma.c:282-289) */
	int32_t DB_size; uint32_t mlen, prefix_len; uint64_t prefix, size, n, v_index, null_index;
	uint32_t kmersize, flag;
	uint64_t kmask, hmask;
	int mega, exist_wide, values_short, key_wide, vidx_wide;
	void *exist, *values, *key_index, *value_index;
	/* .length.b / .seq.b */
	int32_t *lengths;   /* [DB_size], lengths[0] = k of the alignment index */
	uint64_t *seq;      /* packed templates, all of .seq.b */
	int64_t *seq_off;   /* [DB_size] word offset of template t in seq */
} orc_db;

typedef struct {
	int32_t M, MM, U, W1, Wl, Mn, PE;
	int32_t d[25];      /* substitution matrix d[t][q], 5x5 */
	int32_t exhaustive; /* -ex_mode */
	int32_t apm;        /* pairing of read pairs: 0 = -apm p (save_kmers_penaltyPair / alnFragsPenaltyPE), 1 = -apm u, the default
	                       (save_kmers_unionPair / alnFragsUnionPE) */
} orc_params;

typedef struct {
	int64_t reads, mapped, read_words, lookups, hits, list_fetches, list_ids;
} orc_stats;

orc_db *orc_db_open(const char *prefix);
void orc_db_close(orc_db *db);
int64_t orc_lookup(const orc_db *db, uint64_t key);
int orc_list(const orc_db *db, int64_t off, int *n_out, const void **ids);
void orc_revcomp(const uint64_t *seq, int seqlen, const int32_t *N, int nN, uint64_t *rseq, int32_t *rN);
int64_t orc_seed_stream(const orc_db *db, const orc_params *p, const uint8_t *in, size_t in_bytes,
                        uint8_t *out, size_t cap, orc_stats *st);
void orc_default_params(orc_params *p);
int64_t orc_chain_stream(const orc_db *db, const orc_params *p, const uint8_t *in, size_t in_bytes, int minlen,
                         double mrs, double coverT, double mrc, uint8_t *out, size_t cap, orc_stats *st);
size_t orc_fasta_unwrap(const uint8_t *text, size_t n, uint8_t *out);   /* multi-line FASTA -> 2-line FASTA as FileBuffgetFsa reads it */
void orc_stage1_set_quality(int minQ, int hardmaskQ, const double *prob);   /* -eq, -mi and prob[256] = 10^(-q/10) (kma.c:219) of phredStat; 0, 0: off */
void orc_set_proxi(double minFrac); /* -proxi (kma.c:702-718) of stage 2: getProxiMatch, getSecondProxiPen, getF_Proxi / getR_Proxi, getProxiChainTemplates, chooseChain; 1.0 = off */
double orc_get_proxi(void);
void orc_set_soft_proxi(uint64_t *sums);   /* soft proximity sums (kmers.c:133-153, -proxi < 0 with -mem_mode), [DB_size]; NULL = off */
uint64_t *orc_get_soft_proxi(void);
void orc_chain_set_lc(int lc);    /* -lc (kma.c:694-700): length-corrected anker selection of save_kmers_chain, default 0 */
int orc_db_load_seq(orc_db *db, const char *prefix);
int orc_align_stream(orc_db *db, const char *prefix, const orc_params *p, const uint8_t *in, size_t in_bytes,
                     int one2one, double scoreT, int mq, int minlen, double mrc,
                     uint8_t **frag_out, size_t *frag_bytes, uint64_t *as, uint64_t *uas,
                     int32_t **cand_out, size_t *cand_rows, int64_t *nw_cells);
int orc_trace_stream(orc_db *db, const char *prefix, const orc_params *p, const uint8_t *in, size_t in_bytes, int one2one,
                     double scoreT, int mq, int minlen, double mrc, uint8_t **out, size_t *out_bytes);
void orc_align_set_minfrac(double minFrac);   /* -proxi as stage 3 sees it (update_Scores*, alnFrags*PE), default 1.0 */
void orc_trace_set_ts(int ts);   /* -ts of the traceback pass (trimSeeds, chain.c:496), default 0 */
int orc_matrix_stream(const int32_t *lengths, int DB_size, const uint8_t *frags, size_t fb, const uint8_t *trace, size_t tb,
                      int dense, uint16_t *counts);
int64_t orc_conclave_stream(const int32_t *template_lengths, int DB_size, const uint8_t *frag, size_t fb,
                            const uint64_t *alignment_scores, const uint64_t *uniq_alignment_scores,
                            uint8_t *out, size_t cap, uint64_t *w_scores, uint32_t *fragmentCounts, uint32_t *readCounts);
int64_t orc_conclave2_stream(const int32_t *template_lengths, int DB_size, const uint8_t *frag, size_t fb,
                             const uint64_t *alignment_scores, uint64_t *uniq_alignment_scores,
                             uint8_t *out, size_t cap, uint64_t *w_scores, uint32_t *fragmentCounts, uint32_t *readCounts,
                             double scoreT, double evalue, int and_mode, double (*p_chisqr)(long double));   /* runConClave2(_lc), -ConClave 2 */
int orc_memscore_stream(const int32_t *lengths, int DB_size, const uint8_t *in, size_t in_bytes, uint8_t **frag_out, size_t *frag_bytes,
                        uint64_t *as, uint64_t *uas);
int orc_consensus(const uint16_t *counts, const uint64_t *seq, int t_len, int bcd, int caller, int sig, double support, double evalue,
                  uint8_t *t, uint8_t *s, uint8_t *q, uint64_t *stats);
double orc_chi2_min(double evalue);
void orc_to2bit(uint8_t *trans);
size_t orc_stage1(const uint8_t *text1, size_t n1, const uint8_t *text2, size_t n2, int fastq, int min_phred, int phred_scale,
                  int minlen, int maxlen, uint8_t *out, size_t cap, int64_t *count);
void orc_conclave_set_lc(int lc);
void orc_free(void *p);
void orc_nw(const orc_params *p, const uint64_t *tseq, const uint8_t *query, int k, int t_s, int t_e, int q_s, int q_e,
            int band, int *out6);
#endif
