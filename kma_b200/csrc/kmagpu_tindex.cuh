// Per-template position index on the device: what HashMapCCI (hashmapcci.c:95-199, 409-505) answers, laid out for
// one-sector lookups.
//
// Every template t owns an open-addressing table of 8-byte slots {k-mer, value} inside one big allocation:
//   value > 0 : the k-mer occurs once, at this 1-based template position   (hashMapCCI_get returns +pos)
//   value < 0 : the k-mer repeats; -(value) - 1 is the offset of {count, pos_0 < pos_1 < ...} in dups[]
//               (hashMapCCI_get returns -pos_0; getDubPos/getNextDubPos enumerate in ascending position, the
//               order the reference's chains are built in -- SURVEY appendix 14)
// key 0 (poly-A) is never indexed (hashmapcci.c:414), which makes 0 the empty-slot marker. k <= 16 (32-bit keys).
#pragma once
#include <stdint.h>

struct KgTMeta {
	int64_t slot_off;   // first slot of the template's table
	int64_t seq_off;    // word offset of the template in seq
	int32_t len;        // template length in bases
	int32_t shift;      // 32 - log2(table size)
};

struct KgTIndexView {
	const KgTMeta *meta;    // [DB_size]
	const uint2 *slots;
	const int32_t *dups;
	const uint64_t *seq;    // all of .seq.b (+ 2 zero words)
	int32_t k;              // k of the alignment index (.length.b[0])
};

#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t tix_hash(uint32_t key, int shift) { return (key * 0x9E3779B1u) >> shift; }

// hashMapCCI_get: 0 absent, +pos unique, negative = repeated (decode with tix_dups)
__device__ __forceinline__ int tix_get(const KgTIndexView &ix, const KgTMeta &m, uint64_t key) {
	if (key == 0 || (key >> 32)) return 0;
	const uint2 *tab = ix.slots + m.slot_off;
	const uint32_t mask = 0xFFFFFFFFu >> m.shift;
	uint32_t h = tix_hash((uint32_t)key, m.shift);
	for (;;) {
		const uint2 e = __ldg(tab + h);
		if (e.x == (uint32_t)key) return (int)e.y;
		if (e.x == 0) return 0;
		h = (h + 1) & mask;
	}
}

__device__ __forceinline__ const int32_t *tix_dups(const KgTIndexView &ix, int value, int *cnt) {
	const int32_t *d = ix.dups + (size_t)(-(int64_t)value - 1);
	*cnt = __ldg(d);
	return d + 1;
}

// 32 bases starting at base `pos` >= 0, left aligned (seq is padded so word w+1 is readable)
__device__ __forceinline__ uint64_t win32(const uint64_t *seq, int pos) {
	const int w = pos >> 5, b = (pos & 31) << 1;
	uint64_t x = seq[w] << b;
	if (b) x |= seq[w + 1] >> (64 - b);
	return x;
}
__device__ __forceinline__ uint64_t kmer_at(const uint64_t *seq, int pos, int k) { return win32(seq, pos) >> (64 - 2 * k); }

// number of equal bases going forward from q[qi], t[ti], at most maxn
__device__ __forceinline__ int ext_fwd(const uint64_t *q, int qi, const uint64_t *t, int ti, int maxn) {
	int n = 0;
	while (n < maxn) {
		const uint64_t x = win32(q, qi + n) ^ win32(t, ti + n);
		if (x) { n += __clzll((long long)x) >> 1; break; }
		n += 32;
	}
	return n < maxn ? n : maxn;
}
// number of equal bases going backward from q[qi-1], t[ti-1], at most maxn (<= qi, <= ti)
__device__ __forceinline__ int ext_bwd(const uint64_t *q, int qi, const uint64_t *t, int ti, int maxn) {
	int n = 0;
	while (n < maxn) {
		const int step = maxn - n < 32 ? maxn - n : 32;
		const uint64_t x = (win32(q, qi - n - step) ^ win32(t, ti - n - step)) >> (64 - 2 * step);
		if (x) { n += (__ffsll((long long)x) - 1) >> 1; break; }
		n += step;
	}
	return n;
}
#endif
