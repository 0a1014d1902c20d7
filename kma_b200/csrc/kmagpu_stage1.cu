// Stage 1 on the device: what run_input / run_input_PE (runinput.c:370-560) do per read between the record splitter
// (FileBuffgetFq seqparse.c:241 / FileBuffgetFsa) and the stage-1 pipe -- base translation through the caller's `trans`
// table (to2Bit, kma.c:1439-1482), phredStat (runinput.c:127-313: the -mp end trim; with -eq / -mi also the hard mask and
// the bidirectional quality trim, a sequential walk with double sums that one lane runs per read) or fsastat's N trim
// (runinput.c:315-368), the -ml / -xl filters, the pairing rule of run_input_PE (runinput.c:528-539), compDNA
// (compdna.c:99-127) and the records of printFsa / printFsa_pair (runinput.c:765-825).
//
// The host keeps the one sequential step, finding the line ends (kmagpu_fastx_split, memchr speed); the raw text goes
// to HBM once and three small kernels turn it into the stage-1 stream in input order: a warp per read finds the kept
// window and counts its N's, a thread per read (or pair) applies the filters and sizes the records, two scans place
// them, a warp per kept read packs 32 bases per step into a 2-bit word with two warp-wide OR reductions and writes the
// N positions by ballot rank. The stream and its record offsets stay in HBM as the input of kmagpu_seed_run, so reads
// never come back to the host between the file and stage 2.
#include "kmagpu_internal.h"
#include "kmagpu_dev.cuh"
#include <string.h>
#include <algorithm>
#include <vector>

struct S1Win { int32_t start, end, nN, klen; };   // kept window, its N count, the length the -ml filter sees

struct S1Tab { uint8_t t[256]; };

// fields[r] = {header offset, header length, sequence offset, sequence length, quality offset}
__global__ void __launch_bounds__(256) s1_window_kernel(const uint8_t *__restrict__ text, const uint32_t *__restrict__ fields, int n, S1Tab tab,
		int fastq, int thr, int maxlen, S1Win *win, int minQ, int maskQ, int minlen, const double *__restrict__ probtab) {
	__shared__ uint8_t tr[256];
	tr[threadIdx.x] = tab.t[threadIdx.x];
	__syncthreads();
	const unsigned lane = threadIdx.x & 31;
	const int warps = (gridDim.x * blockDim.x) >> 5;
	for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
		const uint32_t *f = fields + 5 * (size_t)r;
		const uint8_t *seq = text + f[2], *qual = text + f[4];
		const int len = (int)f[3];
		int start = 0, end = 0, nN = 0;
		if (len <= maxlen) {
			// first / last position that stays: quality >= threshold (phredStat) or not an N (fsastat)
			start = len;
			for (int base = 0; base < len; base += 32) {
				const int i = base + (int)lane;
				const bool ok = i < len && (fastq ? (int)qual[i] >= thr : tr[seq[i]] != 4);
				const unsigned m = __ballot_sync(0xffffffffu, ok);
				if (m) { start = base + __ffs(m) - 1; break; }
			}
			end = start;
			for (int base = len - 1; base >= start; base -= 32) {
				const int i = base - (int)lane;
				const bool ok = i >= start && (fastq ? (int)qual[i] >= thr : tr[seq[i]] != 4);
				const unsigned m = __ballot_sync(0xffffffffu, ok);
				if (m) { end = base - (__ffs(m) - 1) + 1; break; }
			}
			for (int base = start; base < end; base += 32) {
				const int i = base + (int)lane;
				nN += __popc(__ballot_sync(0xffffffffu, i < end && tr[seq[i]] == 4));
			}
		}
		int klen = fastq ? end - start : end - start - nN;
		if (fastq && (minQ | maskQ) && len <= maxlen) {
			// the -mi / -eq part of phredStat (runinput.c:168-313) on the end-trimmed window: sums of doubles in the reference's
			// order, so one lane walks it (a rare option; the warp's other lanes wait)
			if (lane == 0) {
				const double *prob = probtab;   // already shifted by the phred scale
#define S1_ISN(i) (tr[seq[i]] == 4 || (int)qual[i] < maskQ)
				int s = start, e = end, L = e - s, ns = 0;
				double sp = 0;
				for (int i = s; i < e; ++i) { sp = __dadd_rn(sp, prob[qual[i]]); ns += S1_ISN(i); }
				const double minP = prob[256];   // pow(10, -0.1 * minQ) from the host's libm
				if (minlen <= L - ns && __dmul_rn(minP, (double)L) < sp) {
					int ns5 = 0, ns3 = 0, l5 = 0, l3 = 0, p5 = s, p3 = e - 1;
					double sp5 = 0, sp3 = 0;
					while (l3 < L && thr <= (int)qual[p3]) { sp3 = __dadd_rn(sp3, prob[qual[p3]]); ++l3; ns3 += S1_ISN(p3); --p3; }
					while (l3 < L && (int)qual[p3] < thr) { sp3 = __dadd_rn(sp3, prob[qual[p3]]); ++l3; ns3 += S1_ISN(p3); --p3; }
					while (minlen <= L - ns && __dmul_rn(minP, (double)L) < sp) {
						if (__dmul_rn(sp5, (double)l3) < __dmul_rn(sp3, (double)l5)) {
							e -= l3; ns -= ns3; L -= l3; sp = __dsub_rn(sp, sp3);
							ns3 = 0; l3 = 0; sp3 = 0;
							while (l3 < L && thr <= (int)qual[p3]) { sp3 = __dadd_rn(sp3, prob[qual[p3]]); ++l3; ns3 += S1_ISN(p3); --p3; }
							while (l3 < L && (int)qual[p3] < thr) { sp3 = __dadd_rn(sp3, prob[qual[p3]]); ++l3; ns3 += S1_ISN(p3); --p3; }
						} else {
							s += l5; L -= l5; ns -= ns5; sp = __dsub_rn(sp, sp5);
							ns5 = 0; l5 = 0; sp5 = 0;
							while (l5 < L && thr <= (int)qual[p5]) { sp5 = __dadd_rn(sp5, prob[qual[p5]]); ++l5; ns5 += S1_ISN(p5); ++p5; }
							while (l5 < L && (int)qual[p5] < thr) { sp5 = __dadd_rn(sp5, prob[qual[p5]]); ++l5; ns5 += S1_ISN(p5); ++p5; }
						}
					}
				}
#undef S1_ISN
				start = s; end = e; nN = ns; klen = L - ns;
			}
			start = __shfl_sync(0xffffffffu, start, 0); end = __shfl_sync(0xffffffffu, end, 0);
			nN = __shfl_sync(0xffffffffu, nN, 0); klen = __shfl_sync(0xffffffffu, klen, 0);
		}
		if (lane == 0) { S1Win w = {start, end, nN, klen}; win[r] = w; }
	}
}

// one thread per read (single) or per pair: -ml filter, pairing rule, record sizes and kinds (0 single, 1 / 2 mates)
__global__ void __launch_bounds__(256) s1_decide_kernel(const uint32_t *__restrict__ fields, const S1Win *__restrict__ win, int n, int paired,
		int minlen, uint32_t *size, uint32_t *keep, uint8_t *kind, unsigned long long *ctr) {
	const int u = blockIdx.x * blockDim.x + threadIdx.x;
	const int units = paired ? n >> 1 : n;
	if (u >= units) return;
	const int r0 = paired ? 2 * u : u, cnt = paired ? 2 : 1;
	bool k[2] = {false, false};
	for (int j = 0; j < cnt; ++j) k[j] = minlen <= win[r0 + j].klen;
	const bool both = paired && k[0] && k[1];
	int maxL = 0;
	for (int j = 0; j < cnt; ++j) {
		const int r = r0 + j;
		uint32_t sz = 0;
		if (k[j]) {
			const S1Win w = win[r];
			const int L = w.end - w.start;
			sz = 16u + 8u * (uint32_t)((L + 31) >> 5) + 4u * (uint32_t)w.nN + fields[5 * (size_t)r + 1] + 1u;
			maxL = max(maxL, L);
		}
		size[r] = sz; keep[r] = k[j] ? 1u : 0u;
		kind[r] = both ? (uint8_t)(1 + j) : 0;
	}
	if (k[0] || k[1]) atomicAdd(&ctr[0], 1ull);           // what run_input counts: one per printed read or pair
	if (both) atomicAdd(&ctr[1], 1ull);
	if (maxL) atomicMax(&ctr[2], (unsigned long long)maxL);
}

__device__ __forceinline__ void s1_store_u64(uint8_t *p, unsigned long long v) {   // records are not aligned
	if (((uintptr_t)p & 3) == 0) { ((uint32_t *)p)[0] = (uint32_t)v; ((uint32_t *)p)[1] = (uint32_t)(v >> 32); }
	else { st_u32b(p, (uint32_t)v); st_u32b(p + 4, (uint32_t)(v >> 32)); }
}

// one warp per kept read: compDNA + printFsa
__global__ void __launch_bounds__(256) s1_emit_kernel(const uint8_t *__restrict__ text, const uint32_t *__restrict__ fields, const S1Win *__restrict__ win,
		int n, S1Tab tab, const uint32_t *__restrict__ size, const uint32_t *__restrict__ boff, const uint32_t *__restrict__ ridx,
		const uint8_t *__restrict__ kind, uint8_t *out, uint32_t *rec_off, uint8_t *rec_kind, int maskQ) {
	__shared__ uint8_t tr[256];
	tr[threadIdx.x] = tab.t[threadIdx.x];
	__syncthreads();
	const unsigned lane = threadIdx.x & 31, lt = (1u << lane) - 1;
	const int warps = (gridDim.x * blockDim.x) >> 5;
	for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
		if (!size[r]) continue;
		const uint32_t *f = fields + 5 * (size_t)r;
		const S1Win w = win[r];
		const uint8_t *seq = text + f[2] + w.start, *hdr = text + f[0];
		const uint8_t *qual = text + f[4] + w.start;   // read only under -mi (FASTQ)
		const int L = w.end - w.start, words = (L + 31) >> 5, hl = (int)f[1] + 1;
		uint8_t *o = out + boff[r];
		if (lane == 0) {
			rec_off[ridx[r]] = boff[r]; rec_kind[ridx[r]] = kind[r];
			st_u32b(o, (uint32_t)L); st_u32b(o + 4, (uint32_t)words); st_u32b(o + 8, (uint32_t)w.nN);
			st_u32b(o + 12, (uint32_t)(kind[r] == 1 ? -hl : hl));
		}
		uint8_t *ow = o + 16, *oN = ow + 8 * (size_t)words, *oh = oN + 4 * (size_t)w.nN;
		int nbase = 0;
		for (int wd = 0; wd < words; ++wd) {
			const int i = 32 * wd + (int)lane;
			const unsigned c = i < L ? tr[seq[i]] : 0u;
			const bool isN = c == 4u || (maskQ && i < L && (int)qual[i] < maskQ);   // hard mask: runinput.c:183
			// (word << 2) | base for 32 bases = the OR of base << (62 - 2 * lane); an N shifts in zero (compdna.c:113-121)
			const unsigned long long x = isN ? 0ull : (unsigned long long)c << (62 - 2 * (int)lane);
			const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(x >> 32)), lo = __reduce_or_sync(0xffffffffu, (unsigned)x);
			if (lane == 0) s1_store_u64(ow + 8 * (size_t)wd, ((unsigned long long)hi << 32) | lo);
			const unsigned m = __ballot_sync(0xffffffffu, isN);
			if (isN) st_u32b(oN + 4 * (size_t)(nbase + __popc(m & lt)), (uint32_t)i);
			nbase += __popc(m);
		}
		for (int i = lane; i < hl; i += 32) oh[i] = i < hl - 1 ? hdr[i] : (uint8_t)0;
	}
}

// ---------------------------------------------------------------- host side

static inline bool s1_space(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }

// Host only: the line structure of a chunk of 4-line FASTQ (fastq != 0) or 2-line FASTA text. fields[i] = {header
// offset (past '@' / '>'), header length without trailing white space, sequence offset, sequence length without
// trailing bytes that `trans` maps to 8 (a '\r'), quality offset}. Returns the number of whole records (stops at a
// partial one), *used = the bytes they span; -1 when a record does not start with '@' / '>'.
extern "C" int64_t kmagpu_fastx_split(const void *text_, size_t nbytes, int fastq, const uint8_t *trans, uint32_t *fields, size_t cap,
                                      size_t *used) {
	const uint8_t *text = (const uint8_t *)text_;
	if (!text || !trans) { kmagpu_set_error("null argument"); return -1; }
	if (nbytes >= (1ull << 32)) { kmagpu_set_error("text chunk of %zu bytes exceeds the 4 GiB per-call limit; split it", nbytes); return -1; }
	size_t p = 0, n = 0;
	while (p < nbytes) {
		if (text[p] != (fastq ? '@' : '>')) { kmagpu_set_error("malformed input at byte %zu", p); return -1; }
		const uint8_t *e = (const uint8_t *)memchr(text + p, '\n', nbytes - p);
		if (!e) break;
		size_t ho = p + 1, hl = (size_t)(e - text) - ho;
		while (hl > 0 && s1_space(text[ho + hl - 1])) --hl;
		size_t so = (size_t)(e - text) + 1;
		if (so >= nbytes) break;
		e = (const uint8_t *)memchr(text + so, '\n', nbytes - so);
		if (!e && fastq) break;
		size_t send = e ? (size_t)(e - text) : nbytes, sl = send - so, next = e ? send + 1 : nbytes, qo = 0;
		while (sl > 0 && trans[text[so + sl - 1]] == 8) --sl;
		if (fastq) {
			if (next >= nbytes) break;
			e = (const uint8_t *)memchr(text + next, '\n', nbytes - next);
			if (!e) break;
			qo = (size_t)(e - text) + 1;
			if (qo + sl > nbytes) break;
			e = (const uint8_t *)memchr(text + qo + sl, '\n', nbytes - qo - sl);
			next = e ? (size_t)(e - text) + 1 : nbytes;
		}
		if (fields && n < cap) {
			uint32_t *f = fields + 5 * n;
			f[0] = (uint32_t)ho; f[1] = (uint32_t)hl; f[2] = (uint32_t)so; f[3] = (uint32_t)sl; f[4] = (uint32_t)qo;
		}
		++n;
		p = next;
	}
	if (used) *used = p;
	return (int64_t)n;
}

// Host only: the start of the first record at or after byte `from` of a FASTQ / FASTA chunk, so that several threads
// can split byte ranges of one chunk independently. FASTQ: a line that starts with '@' whose second-next line starts
// with '+' (a quality line may start with '@', a sequence line cannot start with '+'). Returns nbytes when none.
extern "C" size_t kmagpu_fastx_sync(const void *text_, size_t nbytes, int fastq, size_t from) {
	const uint8_t *text = (const uint8_t *)text_;
	if (!text) return nbytes;
	size_t p = from;
	if (p > 0) {   // move to the start of the next line unless `from` already is one
		if (p >= nbytes) return nbytes;
		if (text[p - 1] != '\n') {
			const uint8_t *e = (const uint8_t *)memchr(text + p, '\n', nbytes - p);
			if (!e) return nbytes;
			p = (size_t)(e - text) + 1;
		}
	}
	while (p < nbytes) {
		const uint8_t *e1 = (const uint8_t *)memchr(text + p, '\n', nbytes - p);
		if (text[p] == (fastq ? '@' : '>')) {
			if (!fastq) return p;
			if (!e1) return nbytes;
			const size_t l2 = (size_t)(e1 - text) + 1;
			const uint8_t *e2 = l2 < nbytes ? (const uint8_t *)memchr(text + l2, '\n', nbytes - l2) : nullptr;
			if (!e2) return nbytes;
			const size_t l3 = (size_t)(e2 - text) + 1;
			if (l3 < nbytes && text[l3] == '+') return p;
		}
		if (!e1) return nbytes;
		p = (size_t)(e1 - text) + 1;
	}
	return nbytes;
}

// Host only: multi-line FASTA -> the 2-line form the splitters and kernels take. FileBuffgetFsa (seqparse.c:66-160) keeps,
// between a header line and the next '>', every byte `trans` maps below 8 -- line ends, '\r', blanks and anything else
// drop out wherever they stand -- so a record becomes its header line as it is plus ONE sequence line of the kept bytes.
// Whole records only: *used = the bytes consumed (up to the start of the last record unless eof: its end is only known
// when the next '>' or the end of the file is seen). Returns the bytes written to out (cap >= nbytes + 1), -1 on error.
extern "C" int64_t kmagpu_fasta_unwrap(const void *text_, size_t nbytes, const uint8_t *trans, int eof, void *out_, size_t cap, size_t *used) {
	const uint8_t *text = (const uint8_t *)text_;
	uint8_t *out = (uint8_t *)out_;
	if (!text || !trans || !out) { kmagpu_set_error("null argument"); return -1; }
	if (cap < nbytes + 1) { kmagpu_set_error("kmagpu_fasta_unwrap needs %zu output bytes, caller gave %zu", nbytes + 1, cap); return -1; }
	size_t p = 0, o = 0, done_in = 0, done_out = 0;
	while (p < nbytes) {
		if (text[p] != '>') { kmagpu_set_error("malformed FASTA at byte %zu", p); return -1; }
		const uint8_t *e = (const uint8_t *)memchr(text + p, '\n', nbytes - p);
		if (!e) break;   // header line not complete
		const size_t hend = (size_t)(e - text) + 1;
		memcpy(out + o, text + p, hend - p);
		o += hend - p;
		// sequence: up to the next '>' (anywhere, like the reference's byte loop) or the end of the chunk
		const uint8_t *nx = (const uint8_t *)memchr(text + hend, '>', nbytes - hend);
		const size_t send = nx ? (size_t)(nx - text) : nbytes;
		if (!nx && !eof) { o = done_out; break; }   // the record may continue in the next chunk
		for (size_t i = hend; i < send; ++i) { const uint8_t c = text[i]; if (trans[c] < 8) out[o++] = c; }
		out[o++] = '\n';
		p = send;
		done_in = p; done_out = o;
	}
	if (used) *used = done_in;
	return (int64_t)done_out;
}

int kg_stage1_free(kmagpu_db *db) {
	Stage1Batch &w = db->s1;
	KgBuf *all[] = {&w.d_text, &w.d_fields, &w.d_win, &w.d_u32, &w.d_kind, &w.d_partial, &w.d_ctr, &w.h_ctr, &w.d_cnt1, &w.d_cnt2, &w.d_lines1,
	                &w.d_lines2, &w.d_prob, &w.h_prob};
	for (KgBuf *x : all) x->release();
	return 0;
}

// the part both entry points share: d_text and d_fields hold the chunk and its n field rows; window scan -> filters ->
// scans -> records into the seed batch
static int s1_core(kmagpu_db *db, const kmagpu_ingest_params *ip, int n, void *stage1_out, size_t cap, size_t *out_bytes, int64_t *count) {
	SeedBatch &b = db->seed;
	Stage1Batch &w = db->s1;
	cudaStream_t st = db->stream;
	const int ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
	if (w.d_win.reserve(sizeof(S1Win) * (size_t)n) || w.d_u32.reserve(16 * ((size_t)n + 2)) || w.d_kind.reserve((size_t)n + 8) ||
	    w.d_partial.reserve(4 * (size_t)(ntiles + 2))) return -1;
	uint32_t *size = (uint32_t *)w.d_u32.p, *keep = size + n + 1, *boff = keep + n + 1, *ridx = boff + n + 1;
	unsigned long long *ctr = (unsigned long long *)w.d_ctr.p;
	S1Tab tab;
	memcpy(tab.t, ip->trans, 256);
	const int grid = db->sm_count * 8;
	// -eq / -mi: phredStat's quality trim needs prob[] shifted by the phred scale (runinput.c:410 passes prob - phredScale)
	// and 10^(-0.1 * minQ) from the host's libm
	const bool quality = ip->fastq && (ip->min_q || ip->hardmask_q);
	const int min_phred = quality && ip->min_phred < ip->min_q ? ip->min_q : ip->min_phred;   // runinput.c:380
	const double *d_prob = nullptr;
	if (quality) {
		if (ip->phred_scale < 0 || ip->phred_scale > 128) { kmagpu_set_error("phred scale %d", ip->phred_scale); return -1; }
		w.h_prob.pinned = true;
		if (w.d_prob.reserve(8 * 260) || w.h_prob.reserve(8 * 260)) return -1;
		double *hp = (double *)w.h_prob.p;
		for (int q = 0; q < 256; ++q) hp[q] = q >= ip->phred_scale ? ip->prob[q - ip->phred_scale] : 1.0;   // below the scale the reference reads before its table
		hp[256] = pow(10, (-0.1) * ip->min_q);
		KG_CUDA(cudaMemcpyAsync(w.d_prob.p, hp, 8 * 257, cudaMemcpyHostToDevice, st));
		d_prob = (const double *)w.d_prob.p;
	}
	s1_window_kernel<<<grid, 256, 0, st>>>((const uint8_t *)w.d_text.p, (const uint32_t *)w.d_fields.p, n, tab, ip->fastq, ip->phred_scale + min_phred,
		ip->maxlen, (S1Win *)w.d_win.p, quality ? ip->min_q : 0, quality ? ip->hardmask_q : 0, ip->minlen, d_prob);
	const int units = ip->paired ? n / 2 : n;
	s1_decide_kernel<<<(units + 255) / 256, 256, 0, st>>>((const uint32_t *)w.d_fields.p, (const S1Win *)w.d_win.p, n, ip->paired, ip->minlen, size, keep,
		(uint8_t *)w.d_kind.p, ctr);
	kg_exscan(size, n, boff, (uint32_t *)w.d_partial.p, ctr + 3, st);
	kg_exscan(keep, n, ridx, (uint32_t *)w.d_partial.p, ctr + 4, st);
	unsigned long long *h = (unsigned long long *)w.h_ctr.p;
	KG_CUDA(cudaMemcpyAsync(h, ctr, 64, cudaMemcpyDeviceToHost, st));
	KG_CUDA(cudaStreamSynchronize(st));
	KG_CUDA(cudaGetLastError());
	const size_t ob = (size_t)h[3], nrec = (size_t)h[4];
	if (ob >= (1ull << 31)) { kmagpu_set_error("stage-1 stream of %zu bytes exceeds the 2 GiB per-call limit of stage 2; split the text", ob); return -1; }
	if (b.d_in.reserve(ob + 64) || b.d_off.reserve(4 * (nrec + 1)) || b.d_kinds.reserve(nrec + 1)) return -1;
	s1_emit_kernel<<<grid, 256, 0, st>>>((const uint8_t *)w.d_text.p, (const uint32_t *)w.d_fields.p, (const S1Win *)w.d_win.p, n, tab, size, boff, ridx,
		(const uint8_t *)w.d_kind.p, (uint8_t *)b.d_in.p, (uint32_t *)b.d_off.p, (uint8_t *)b.d_kinds.p, quality ? ip->hardmask_q : 0);
	h[7] = (unsigned long long)ob;   // the closing offset, from pinned memory
	KG_CUDA(cudaMemcpyAsync((uint32_t *)b.d_off.p + nrec, &h[7], 4, cudaMemcpyHostToDevice, st));
	KG_CUDA(cudaMemsetAsync((uint8_t *)b.d_kinds.p + nrec, 0, 1, st));
	KG_CUDA(cudaMemsetAsync((uint8_t *)b.d_in.p + ob, 0, 64, st));
	KG_CUDA(cudaEventRecord(db->ev[1], st));
	if (stage1_out) {
		if (ob > cap) { kmagpu_set_error("stage-1 output needs %zu bytes, caller gave %zu", ob, cap); cudaStreamSynchronize(st); return -1; }
		if (ob) KG_CUDA(cudaMemcpyAsync(stage1_out, b.d_in.p, ob, cudaMemcpyDeviceToHost, st));
	}
	KG_CUDA(cudaStreamSynchronize(st));
	KG_CUDA(cudaGetLastError());
	// the batch now looks as kmagpu_seed_upload would have left it
	b.nreads = (int64_t)nrec; b.npairs = (int64_t)h[1]; b.max_seqlen = (int32_t)h[2]; b.in_bytes = ob;
	if (out_bytes) *out_bytes = ob;
	if (count) *count = (int64_t)h[0];
	return 0;
}

static int s1_begin(kmagpu_db *db, const void *text, size_t text1_bytes, const void *text2, size_t text2_bytes, size_t *out_bytes, int64_t *count, float *ms) {
	KG_CUDA(cudaSetDevice(db->device));
	if (out_bytes) *out_bytes = 0;
	if (count) *count = 0;
	if (ms) *ms = 0.f;
	SeedBatch &b = db->seed;
	b.nreads = 0; b.npairs = 0; b.in_bytes = 0; b.max_seqlen = 0; b.ran = false;
	Stage1Batch &w = db->s1;   // buffers persist: no allocation (and no implicit device synchronisation) in the steady state
	w.h_ctr.pinned = true;
	if (w.h_ctr.reserve(128) || w.d_ctr.reserve(64) || w.d_text.reserve(text1_bytes + text2_bytes + 128)) return -1;
	cudaStream_t st = db->stream;
	if (text1_bytes) KG_CUDA(cudaMemcpyAsync(w.d_text.p, text, text1_bytes, cudaMemcpyHostToDevice, st));
	if (text2_bytes) KG_CUDA(cudaMemcpyAsync((uint8_t *)w.d_text.p + text1_bytes, text2, text2_bytes, cudaMemcpyHostToDevice, st));
	KG_CUDA(cudaMemsetAsync((uint8_t *)w.d_text.p + text1_bytes + text2_bytes, 0, 64, st));
	KG_CUDA(cudaMemsetAsync(w.d_ctr.p, 0, 64, st));
	return 0;
}

extern "C" int kmagpu_stage1_batch(kmagpu_db *db, const kmagpu_ingest_params *ip, const void *text, size_t text1_bytes, const void *text2,
                                   size_t text2_bytes, const uint32_t *fields, size_t nreads, void *stage1_out, size_t cap, size_t *out_bytes,
                                   int64_t *count, float *ms) {
	if (!db || !ip || (!text && text1_bytes) || (!text2 && text2_bytes) || (!fields && nreads)) { kmagpu_set_error("null argument"); return -1; }
	const size_t text_bytes = text1_bytes + text2_bytes;   // the second file's text follows the first in one device buffer
	if (text_bytes >= (1ull << 32) - 128) { kmagpu_set_error("text chunk of %zu bytes exceeds the 4 GiB per-call limit; split it", text_bytes); return -1; }
	if (nreads >= (1ull << 31)) { kmagpu_set_error("too many reads in one call"); return -1; }
	if (ip->paired && (nreads & 1)) { kmagpu_set_error("paired input needs an even number of reads (mates at 2i, 2i + 1)"); return -1; }
	for (size_t i = 0; i < nreads; ++i) {
		const uint32_t *f = fields + 5 * i;
		if ((size_t)f[0] + f[1] > text_bytes || (size_t)f[2] + f[3] > text_bytes || (ip->fastq && (size_t)f[4] + f[3] > text_bytes)) {
			kmagpu_set_error("read %zu points outside the text", i); return -1;
		}
	}
	if (s1_begin(db, text, text1_bytes, text2, text2_bytes, out_bytes, count, ms)) return -1;
	const int n = (int)nreads;
	if (n == 0) return 0;
	Stage1Batch &w = db->s1;
	if (w.d_fields.reserve(20 * (size_t)n)) return -1;
	KG_CUDA(cudaMemcpyAsync(w.d_fields.p, fields, 20 * (size_t)n, cudaMemcpyHostToDevice, db->stream));
	KG_CUDA(cudaEventRecord(db->ev[0], db->stream));
	if (s1_core(db, ip, n, stage1_out, cap, out_bytes, count)) return -1;
	if (ms) cudaEventElapsedTime(ms, db->ev[0], db->ev[1]);
	return 0;
}

// ---------------------------------------------------------------- the record splitter on the device
// Line ends of a text: one thread per 64-byte block counts its newlines (byte-wise SIMD compare), a scan places them,
// the same threads write the positions; a thread per record then turns four (FASTQ) or two (FASTA) consecutive lines
// into the field row the host splitter would have produced.

#define S1_BLK 64

__device__ __forceinline__ unsigned s1_nl_mask(unsigned w) { return __vcmpeq4(w, 0x0a0a0a0au); }   // 0xff per newline byte

__global__ void __launch_bounds__(256) s1_nl_count_kernel(const uint4 *__restrict__ text, uint32_t nblk, uint32_t nbytes, int eof, uint32_t *cnt) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nblk) return;
	uint32_t c = 0;
#pragma unroll
	for (int j = 0; j < S1_BLK / 16; ++j) {   // the buffer is zero padded past nbytes: no newline there
		const uint4 v = __ldg(text + (size_t)i * (S1_BLK / 16) + j);
		c += (__popc(s1_nl_mask(v.x)) + __popc(s1_nl_mask(v.y)) + __popc(s1_nl_mask(v.z)) + __popc(s1_nl_mask(v.w))) >> 3;
	}
	// a last line without its newline counts as a line at the end of the file
	if (eof && i == nblk - 1 && nbytes && ((const uint8_t *)text)[nbytes - 1] != '\n') ++c;
	cnt[i] = c;
}

__global__ void __launch_bounds__(256) s1_nl_scatter_kernel(const uint8_t *__restrict__ text, uint32_t nblk, uint32_t nbytes, int eof,
		const uint32_t *__restrict__ off, uint32_t *line_end) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nblk) return;
	uint32_t o = off[i];
	const uint32_t base = i * S1_BLK;
#pragma unroll 4
	for (int j = 0; j < S1_BLK / 4; ++j) {
		unsigned m = s1_nl_mask(__ldg((const unsigned *)(text + base) + j)) & 0x01010101u;
		while (m) { const int b = (__ffs(m) - 1) >> 3; line_end[o++] = base + 4 * j + b; m &= m - 1; }
	}
	if (eof && i == nblk - 1 && nbytes && text[nbytes - 1] != '\n') line_end[o] = nbytes;
}

// record r of a file whose text starts at byte `base` of the chunk buffer: lines lpr*r .. lpr*r + lpr - 1
__global__ void __launch_bounds__(256) s1_fields_kernel(const uint8_t *__restrict__ text, const uint32_t *__restrict__ line_end, int nrec, uint32_t base,
		int fastq, S1Tab tab, int stride, int which, uint32_t *fields, unsigned long long *ctr) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= nrec) return;
	const int lpr = fastq ? 4 : 2, l0 = lpr * r;
	const uint32_t s0 = l0 ? line_end[l0 - 1] + 1 : 0, e0 = line_end[l0], e1 = line_end[l0 + 1];
	const uint8_t *t = text + base;
	if (t[s0] != (fastq ? '@' : '>')) atomicAdd(&ctr[5], 1ull);
	uint32_t ho = s0 + 1, hl = e0 > ho ? e0 - ho : 0;
	while (hl > 0) { const uint8_t c = t[ho + hl - 1]; if (c == ' ' || (c >= 9 && c <= 13)) --hl; else break; }
	uint32_t so = e0 + 1, sl = e1 > so ? e1 - so : 0, qo = 0;
	while (sl > 0 && tab.t[t[so + sl - 1]] == 8) --sl;
	if (fastq) {
		qo = line_end[l0 + 2] + 1;
		if (line_end[l0 + 3] < qo + sl) atomicAdd(&ctr[5], 1ull);   // quality line shorter than the sequence
	}
	uint32_t *f = fields + 5 * ((size_t)stride * r + which);
	f[0] = base + ho; f[1] = hl; f[2] = base + so; f[3] = sl; f[4] = base + qo;
}

static int s1_lines(kmagpu_db *db, const uint8_t *d_text, size_t nbytes, int eof, KgBuf &d_cnt, unsigned long long *total_ctr) {
	const uint32_t nblk = (uint32_t)((nbytes + S1_BLK - 1) / S1_BLK);
	Stage1Batch &w = db->s1;
	const int ntiles = (int)((nblk + SCAN_TILE - 1) / SCAN_TILE);
	if (d_cnt.reserve(8 * ((size_t)nblk + 2)) || w.d_partial.reserve(4 * (size_t)(ntiles + 2))) return -1;
	uint32_t *cnt = (uint32_t *)d_cnt.p, *off = cnt + nblk + 1;
	cudaStream_t st = db->stream;
	s1_nl_count_kernel<<<(nblk + 255) / 256, 256, 0, st>>>((const uint4 *)d_text, nblk, (uint32_t)nbytes, eof, cnt);
	kg_exscan(cnt, (int)nblk, off, (uint32_t *)w.d_partial.p, total_ctr, st);
	return 0;
}

extern "C" int kmagpu_stage1_text(kmagpu_db *db, const kmagpu_ingest_params *ip, const void *text1, size_t bytes1, const void *text2, size_t bytes2,
                                  int eof, size_t *used1, size_t *used2, void *stage1_out, size_t cap, size_t *out_bytes, int64_t *count, float *ms) {
	if (!db || !ip || (!text1 && bytes1) || (!text2 && bytes2)) { kmagpu_set_error("null argument"); return -1; }
	if (used1) *used1 = 0;
	if (used2) *used2 = 0;
	if (ip->paired && !text2) { kmagpu_set_error("paired text input needs the second file's chunk"); return -1; }
	if (!ip->paired && bytes2) { kmagpu_set_error("a second chunk needs paired = 1"); return -1; }
	// the second chunk starts on a 64-byte boundary of the device buffer so that both are scanned in aligned blocks
	const size_t base2 = (bytes1 + S1_BLK - 1) / S1_BLK * S1_BLK;
	if (base2 + bytes2 >= (1ull << 32) - 256) { kmagpu_set_error("text chunks of %zu bytes exceed the 4 GiB per-call limit; split them", bytes1 + bytes2); return -1; }
	if (s1_begin(db, nullptr, 0, nullptr, 0, out_bytes, count, ms)) return -1;
	if (bytes1 == 0) return 0;
	Stage1Batch &w = db->s1;
	cudaStream_t st = db->stream;
	if (w.d_text.reserve(base2 + bytes2 + 256)) return -1;
	uint8_t *dt = (uint8_t *)w.d_text.p;
	KG_CUDA(cudaMemsetAsync(w.d_ctr.p, 0, 64, st));
	KG_CUDA(cudaMemcpyAsync(dt, text1, bytes1, cudaMemcpyHostToDevice, st));
	KG_CUDA(cudaMemsetAsync(dt + bytes1, 0, base2 - bytes1 + 64, st));
	if (bytes2) {
		KG_CUDA(cudaMemcpyAsync(dt + base2, text2, bytes2, cudaMemcpyHostToDevice, st));
		KG_CUDA(cudaMemsetAsync(dt + base2 + bytes2, 0, 128, st));
	}
	KG_CUDA(cudaEventRecord(db->ev[0], st));
	unsigned long long *ctr = (unsigned long long *)w.d_ctr.p, *h = (unsigned long long *)w.h_ctr.p;
	// pass 1: newline counts per block + scan, for both chunks (ctr[6], ctr[7] = their line counts)
	if (s1_lines(db, dt, bytes1, eof, w.d_cnt1, ctr + 6)) return -1;
	if (bytes2 && s1_lines(db, dt + base2, bytes2, eof, w.d_cnt2, ctr + 7)) return -1;
	KG_CUDA(cudaMemcpyAsync(h, ctr, 64, cudaMemcpyDeviceToHost, st));
	KG_CUDA(cudaStreamSynchronize(st));
	KG_CUDA(cudaGetLastError());
	const int lpr = ip->fastq ? 4 : 2;
	const size_t lines1 = (size_t)h[6], lines2 = (size_t)h[7];
	size_t nrec = lines1 / lpr;
	if (ip->paired && lines2 / lpr < nrec) nrec = lines2 / lpr;
	if (nrec >= (1ull << 30)) { kmagpu_set_error("too many reads in one call"); return -1; }
	if (nrec == 0) return 0;
	// pass 2: line ends, then the field rows (mates interleaved)
	if (w.d_lines1.reserve(4 * (lines1 + 2)) || (bytes2 && w.d_lines2.reserve(4 * (lines2 + 2)))) return -1;
	const int stride = ip->paired ? 2 : 1, n = (int)nrec * stride;
	if (w.d_fields.reserve(20 * (size_t)n)) return -1;
	S1Tab tab;
	memcpy(tab.t, ip->trans, 256);
	{
		const uint32_t nblk = (uint32_t)((bytes1 + S1_BLK - 1) / S1_BLK);
		s1_nl_scatter_kernel<<<(nblk + 255) / 256, 256, 0, st>>>(dt, nblk, (uint32_t)bytes1, eof, (const uint32_t *)w.d_cnt1.p + nblk + 1, (uint32_t *)w.d_lines1.p);
		s1_fields_kernel<<<((int)nrec + 255) / 256, 256, 0, st>>>(dt, (const uint32_t *)w.d_lines1.p, (int)nrec, 0u, ip->fastq, tab, stride, 0,
			(uint32_t *)w.d_fields.p, ctr);
	}
	if (bytes2) {
		const uint32_t nblk = (uint32_t)((bytes2 + S1_BLK - 1) / S1_BLK);
		s1_nl_scatter_kernel<<<(nblk + 255) / 256, 256, 0, st>>>(dt + base2, nblk, (uint32_t)bytes2, eof, (const uint32_t *)w.d_cnt2.p + nblk + 1, (uint32_t *)w.d_lines2.p);
		s1_fields_kernel<<<((int)nrec + 255) / 256, 256, 0, st>>>(dt, (const uint32_t *)w.d_lines2.p, (int)nrec, (uint32_t)base2, ip->fastq, tab, stride, 1,
			(uint32_t *)w.d_fields.p, ctr);
	}
	// bytes the whole records span (what the host carries over is the rest)
	uint32_t *hu = (uint32_t *)(h + 8);   // past the eight counters s1_core reads back
	KG_CUDA(cudaMemcpyAsync(hu, (const uint32_t *)w.d_lines1.p + nrec * lpr - 1, 4, cudaMemcpyDeviceToHost, st));
	if (bytes2) KG_CUDA(cudaMemcpyAsync(hu + 2, (const uint32_t *)w.d_lines2.p + nrec * lpr - 1, 4, cudaMemcpyDeviceToHost, st));
	if (s1_core(db, ip, n, stage1_out, cap, out_bytes, count)) return -1;   // synchronises the stream
	if (h[5]) { kmagpu_set_error("%llu records are malformed (no '@' / '>' at the start, or a quality line shorter than its sequence)", h[5]); return -1; }
	if (used1) *used1 = std::min<size_t>((size_t)hu[0] + 1, bytes1);
	if (used2 && bytes2) *used2 = std::min<size_t>((size_t)hu[2] + 1, bytes2);
	if (ms) cudaEventElapsedTime(ms, db->ev[0], db->ev[1]);
	return 0;
}
