// Device helpers shared by the translation units of libkmagpu.so: unaligned little-endian access to the
// reference's record streams, 2-bit sequence windows, and the three-kernel exclusive scan that turns per-read
// record sizes into output offsets.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t ld_u32u(const uint8_t *p) {  // unaligned little-endian load
	uintptr_t a = (uintptr_t)p;
	const uint32_t *q = (const uint32_t *)(a & ~(uintptr_t)3);
	unsigned sh = (unsigned)(a & 3) * 8;
	uint32_t lo = __ldg(q);
	if (sh == 0) return lo;
	return __funnelshift_r(lo, __ldg(q + 1), sh);
}
__device__ __forceinline__ uint64_t ld_u64u(const uint8_t *p) {
	return (uint64_t)ld_u32u(p) | ((uint64_t)ld_u32u(p + 4) << 32);
}
__device__ __forceinline__ void st_u32b(uint8_t *p, uint32_t v) {   // unaligned store
	p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}

// n bytes from src to dst by one warp, any alignment on either side: the destination is written in aligned 4-byte
// words (one 128-byte store per warp instruction; each word = two aligned loads + a funnel shift), at most 3 bytes at
// either end go byte-wise. src must be read-only for the kernel (__ldg).
__device__ __forceinline__ void warp_copy(uint8_t *dst, const uint8_t *src, int n, unsigned lane) {
	int head = (int)((4u - (unsigned)((uintptr_t)dst & 3)) & 3u);
	if (head > n) head = n;
	if ((int)lane < head) dst[lane] = __ldg(src + lane);
	const int nw = (n - head) >> 2;
	uint32_t *d4 = (uint32_t *)(dst + head);
	const uint8_t *s = src + head;
#pragma unroll 1
	for (int w = (int)lane; w < nw; w += 32) d4[w] = ld_u32u(s + 4 * (size_t)w);
	const int done = head + 4 * nw;
	if ((int)lane < n - done) dst[done + lane] = __ldg(src + done + lane);
}
// n 32-bit values val(i) to dst (any alignment) by one warp: aligned destinations take whole-word stores
template <typename F> __device__ __forceinline__ void warp_store_u32(uint8_t *dst, int n, unsigned lane, F val) {
	if (((uintptr_t)dst & 3) == 0) {
#pragma unroll 1
		for (int i = (int)lane; i < n; i += 32) ((uint32_t *)dst)[i] = (uint32_t)val(i);
	} else {
#pragma unroll 1
		for (int i = (int)lane; i < n; i += 32) st_u32b(dst + 4 * (size_t)i, (uint32_t)val(i));
	}
}

// reverse the order of the 32 two-bit symbols of w
__device__ __forceinline__ uint64_t rev2(uint64_t w) {
	w = __brevll(w);
	return ((w >> 1) & 0x5555555555555555ull) | ((w & 0x5555555555555555ull) << 1);
}

// 32 bases of an (unaligned) packed read starting at base position pos (pos may be negative / past the end)
__device__ __forceinline__ uint64_t fwd32(const uint8_t *seq, int words, int pos) {
	if (pos <= -32) return 0;
	if (pos < 0) return (words > 0 ? ld_u64u(seq) : 0ull) >> (2 * -pos);
	int w = pos >> 5, b = (pos & 31) << 1;
	uint64_t x = w < words ? ld_u64u(seq + 8 * (size_t)w) << b : 0ull;
	if (b && w + 1 < words) x |= ld_u64u(seq + 8 * (size_t)(w + 1)) >> (64 - b);
	return x;
}

// one atomic per warp for the counters every thread would otherwise hit at one address (a same-address atomic per record
// serialises in L2: 4 M records were 2 ms of a 2.7 ms kernel)
__device__ __forceinline__ void warp_add_u64(unsigned long long *p, unsigned v) {
	const unsigned m = __activemask();
	const unsigned s = __reduce_add_sync(m, v);
	if ((threadIdx.x & 31) == (unsigned)(__ffs(m) - 1) && s) atomicAdd(p, (unsigned long long)s);
}
__device__ __forceinline__ void warp_max_u64(unsigned long long *p, unsigned v) {
	const unsigned m = __activemask();
	const unsigned s = __reduce_max_sync(m, v);
	if ((threadIdx.x & 31) == (unsigned)(__ffs(m) - 1) && s) atomicMax(p, (unsigned long long)s);
}

// grid of a persistent (grid-stride) kernel: exactly one wave of resident CTAs. A fixed "8 per SM" ran in two waves when
// the kernel's registers allowed 6 (aln_emit_kernel: 28 % of the warp slots active).
template <typename K> static inline int kg_wave_grid(K kernel, int block, int sm_count) {
	int per_sm = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
	return sm_count * per_sm;
}

// ---------------------------------------------------------------- exclusive scan (u32 sizes -> u32 offsets)

#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

static __device__ __forceinline__ uint32_t block_exscan(uint32_t v, uint32_t *total) {
	__shared__ uint32_t wsum[SCAN_THREADS / 32];
	const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	uint32_t x = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
	if (lane == 31) wsum[wid] = x;
	__syncthreads();
	if (wid == 0) {
		uint32_t s = lane < SCAN_THREADS / 32 ? wsum[lane] : 0;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
		if (lane < SCAN_THREADS / 32) wsum[lane] = s;
	}
	__syncthreads();
	uint32_t prev = wid ? wsum[wid - 1] : 0;
	*total = wsum[SCAN_THREADS / 32 - 1];
	__syncthreads();
	return prev + x - v;
}

static __global__ void __launch_bounds__(SCAN_THREADS) scan_tiles_kernel(const uint32_t *size, int n, uint32_t *off, uint32_t *partial) {
	const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
	uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i) { v[i] = base + i < n ? size[base + i] : 0; s += v[i]; }
	uint32_t tot, ex = block_exscan(s, &tot);
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i) { if (base + i < n) off[base + i] = ex; ex += v[i]; }
	if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

// *total receives the sum of all sizes as a 64-bit number: offsets are 32-bit, so a caller whose total reaches 2^32 must
// refuse the batch (KG_SCAN_FITS) before any kernel writes through the wrapped offsets
static __global__ void __launch_bounds__(SCAN_THREADS) scan_partials_kernel(uint32_t *partial, int nb, unsigned long long *total) {
	unsigned long long carry = 0;
	for (int base = 0; base < nb; base += SCAN_THREADS) {
		int i = base + threadIdx.x;
		uint32_t v = i < nb ? partial[i] : 0, tot;
		uint32_t ex = block_exscan(v, &tot);   // a group of 256 tiles of 2048 sizes: the callers' sizes keep this below 2^32
		if (i < nb) partial[i] = (uint32_t)(carry + ex);
		carry += tot;
	}
	if (threadIdx.x == 0) *total = carry;
}
#define KG_SCAN_FITS(total, what)                                                                                         \
	do {                                                                                                                  \
		if ((unsigned long long)(total) >= (1ull << 32)) {                                                                \
			kmagpu_set_error("%s of this batch needs %llu bytes / entries, more than 32-bit offsets address: split the batch", what, (unsigned long long)(total)); \
			return -1;                                                                                                    \
		}                                                                                                                 \
	} while (0)

static __global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(uint32_t *off, int n, const uint32_t *partial) {
	const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
	const uint32_t add = partial[blockIdx.x];
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i) if (base + i < n) off[base + i] += add;
}

// off[i] = sum(size[0..i)), *total = sum(size[0..n)); partial needs (n + SCAN_TILE - 1) / SCAN_TILE + 1 entries
static inline void kg_exscan(const uint32_t *size, int n, uint32_t *off, uint32_t *partial, unsigned long long *total,
                             cudaStream_t st) {
	const int ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
	scan_tiles_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(size, n, off, partial);
	scan_partials_kernel<<<1, SCAN_THREADS, 0, st>>>(partial, ntiles, total);
	scan_add_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(off, n, partial);
}
