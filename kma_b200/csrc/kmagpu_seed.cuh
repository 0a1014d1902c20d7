// Device helpers shared by the two stage-2 kernels (kmagpu_seed.cu: save_kmers / pairs, kmagpu_chain.cu: save_kmers_chain):
// the fused template k-mer hash lookup, template lists, the per-read view of a stage-1 record and its k-mer windows.
#pragma once
#include "kmagpu_internal.h"
#include "kmagpu_dev.cuh"

#define KG_MISS 0xFFFFFFFFu

// counters living in d_ctr (uint64 slots)
enum { C_WORK = 0, C_POOL = 1, C_OVF = 2, C_WORK2 = 3, C_POOLFAIL = 4, C_LOOKUPS = 5, C_HITS = 6, C_LISTS = 7,
       C_LISTIDS = 8, C_MAPPED = 9, C_WORDS = 10, C_TOTAL = 11, C_POOL2 = 12, C_NLIST = 13, C_WORK3 = 14, C_N = 16 };

// ---------------------------------------------------------------- small device helpers

// the rest of a bucket's run after its first key did not match: entries pos + 1 .. pos + cnt - 1 are the ones the
// reference's scan (hashmapkma.c:149-178: same bucket, below n) would still examine
__device__ __forceinline__ uint32_t hash_chain(const KgHashView &hv, const uint4 b, uint32_t key) {
	for (uint32_t i = 1; i < b.w; ++i) {
		const uint2 e = __ldg(hv.kv + b.z + i);
		if (e.x == key) return e.y;
	}
	return KG_MISS;
}

// one bucket entry resolved against a key
__device__ __forceinline__ uint32_t hash_resolve(const KgHashView &hv, const uint4 b, uint32_t key) {
	if (b.w == 0) return KG_MISS;
	if (b.x == key) return b.y;
	return b.w == 1 ? KG_MISS : hash_chain(hv, b, key);
}

__device__ __forceinline__ uint32_t hash_lookup(const KgHashView &hv, uint64_t key) {
	if (hv.mega) {
		uint32_t v = __ldg(hv.exist + key);
		return v != 1u ? v : KG_MISS;
	}
	return hash_resolve(hv, __ldg(hv.bk + (uint32_t)(key & hv.hmask)), (uint32_t)key);
}

__device__ __forceinline__ int list_len(const KgHashView &hv, uint32_t off) {
	return hv.values_s ? (int)__ldg(hv.values_s + off) : (int)__ldg(hv.values_w + off);
}
__device__ __forceinline__ int list_id(const KgHashView &hv, uint32_t off, int i) {
	return hv.values_s ? (int)__ldg(hv.values_s + off + 1 + i) : (int)__ldg(hv.values_w + off + 1 + i);
}

// ---------------------------------------------------------------- per-read context

struct ReadCtx {
	const uint8_t *rec;   // stage-1 record
	const uint8_t *seq;   // packed words (unaligned)
	const uint8_t *N;     // int32 list (unaligned)
	int seqlen, words, nN, hdrlen;
};

// i-th N position in strand coordinates (reverse strand mirrors the list, compdna.c:249-254)
__device__ __forceinline__ int n_at(const ReadCtx &rc, int i, int strand) {
	return strand ? rc.seqlen - 1 - (int)ld_u32u(rc.N + 4 * (rc.nN - 1 - i)) : (int)ld_u32u(rc.N + 4 * i);
}

// validity of k-mer position j (strand coords) and start of its N-free stretch
__device__ __forceinline__ bool pos_valid(const ReadCtx &rc, int j, int k, int strand, int *segstart) {
	*segstart = 0;
	if (j + k > rc.seqlen) return false;
	if (rc.nN == 0) return true;
	int lo = 0, hi = rc.nN;   // first N >= j
	while (lo < hi) {
		int mid = (lo + hi) >> 1;
		if (n_at(rc, mid, strand) < j) lo = mid + 1; else hi = mid;
	}
	if (lo > 0) *segstart = n_at(rc, lo - 1, strand) + 1;
	return lo == rc.nN || n_at(rc, lo, strand) > j + k - 1;
}

// forward-strand k-mer at forward position pos from the staged words (window starts at word w0)
__device__ __forceinline__ uint64_t kmer_from(const uint64_t *sw, int w0, int pos, int k) {
	int w = (pos >> 5) - w0, b = (pos & 31) << 1, sh = 64 - 2 * k;
	uint64_t x = sw[w] << b;
	if (b > sh) x |= sw[w + 1] >> (64 - b);
	return x >> sh;
}

