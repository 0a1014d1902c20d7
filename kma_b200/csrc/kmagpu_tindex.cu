// Builds the per-template position index (kmagpu_tindex.cuh) on the device from the HBM-resident .seq.b.
//
// The reference builds a HashMapCCI per template, lazily, single-threaded (hashMapCCI_load, hashmapcci.c:470-505;
// alnfrags.c:1083-1089). Here all templates are indexed at once by four data-parallel passes over the packed
// database: (1) every position inserts its k-mer into its template's table with atomicCAS and counts occurrences,
// (2) repeated k-mers are given a slice of dups[], (3) every position writes itself into its slot or slice,
// (4) slices are sorted ascending -- the enumeration order of the reference's chains.
#include "kmagpu_internal.h"
#include <algorithm>

// template that owns packed word w: largest t in [1, DB_size) with seq_off[t] <= w
__device__ __forceinline__ int owner_of_word(const KgTMeta *meta, int DB_size, int64_t w) {
	int lo = 1, hi = DB_size - 1;
	while (lo < hi) {
		int mid = (lo + hi + 1) >> 1;
		if (meta[mid].seq_off <= w) lo = mid; else hi = mid - 1;
	}
	return lo;
}

__device__ __forceinline__ uint32_t find_slot(const uint2 *tab, const KgTMeta &m, uint32_t key) {
	const uint32_t mask = 0xFFFFFFFFu >> m.shift;
	uint32_t h = tix_hash(key, m.shift);
	while (tab[h].x != key) h = (h + 1) & mask;
	return h;
}

// pass 0 = insert + count, pass 2 = place
__global__ void tix_positions_kernel(KgTIndexView ix, uint2 *slots, int32_t *dups, int DB_size, int64_t total_bases, int pass) {
	const int k = ix.k;
	for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total_bases; g += (int64_t)gridDim.x * blockDim.x) {
		const int t = owner_of_word(ix.meta, DB_size, g >> 5);
		const KgTMeta m = ix.meta[t];
		const int64_t pos = g - (m.seq_off << 5);
		if (pos + k > m.len) continue;
		const uint32_t key = (uint32_t)kmer_at(ix.seq + m.seq_off, (int)pos, k);
		if (key == 0) continue;   // hashMapCCI_add skips poly-A (hashmapcci.c:414)
		uint2 *tab = slots + m.slot_off;
		if (pass == 0) {
			const uint32_t mask = 0xFFFFFFFFu >> m.shift;
			uint32_t h = tix_hash(key, m.shift);
			for (;;) {
				const uint32_t old = atomicCAS(&tab[h].x, 0u, key);
				if (old == 0u || old == key) { atomicAdd(&tab[h].y, 1u); break; }
				h = (h + 1) & mask;
			}
		} else {
			const uint32_t h = find_slot(tab, m, key);
			const int v = (int)tab[h].y;
			if (v < 0) {
				int32_t *d = dups + (size_t)(-(int64_t)v - 1);
				const int idx = atomicAdd(d, 1);
				d[1 + idx] = (int32_t)pos + 1;
			} else tab[h].y = (uint32_t)pos + 1;   // sole occurrence: no other writer
		}
	}
}

// pass 1: slots of repeated k-mers get a slice {cursor = 0, positions...} of dups[]
__global__ void tix_assign_kernel(uint2 *slots, int64_t nslots, int32_t *dups, unsigned long long *cursor, int count_only) {
	for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nslots; s += (int64_t)gridDim.x * blockDim.x) {
		const uint2 e = slots[s];
		if (e.x == 0 || e.y < 2) continue;
		const unsigned long long off = atomicAdd(cursor, (unsigned long long)e.y + 1);
		if (!count_only) { dups[off] = 0; slots[s].y = (uint32_t)(-(int64_t)off - 1); }
	}
}

// pass 3: ascending positions inside every slice (shell sort; slices are short except in low-complexity sequence)
__global__ void tix_sort_kernel(const uint2 *slots, int64_t nslots, int32_t *dups) {
	for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nslots; s += (int64_t)gridDim.x * blockDim.x) {
		const uint2 e = slots[s];
		if (e.x == 0 || (int)e.y >= 0) continue;
		int32_t *d = dups + (size_t)(-(int64_t)(int)e.y - 1);
		const int n = d[0];
		int32_t *a = d + 1;
		int gap = 1;
		while (gap < n / 3) gap = 3 * gap + 1;
		for (; gap > 0; gap /= 3)
			for (int i = gap; i < n; ++i) {
				const int32_t v = a[i];
				int j = i;
				for (; j >= gap && a[j - gap] > v; j -= gap) a[j] = a[j - gap];
				a[j] = v;
			}
	}
}

int kg_tindex_build(kmagpu_db *db) {
	const int DB = db->info.DB_size;
	const int k = db->info.kmerindex;
	if (db->lengths.empty() || !db->d_seq) { kmagpu_set_error("alignment index needs .length.b and .seq.b"); return -1; }
	if (k < 4 || k > 16) { kmagpu_set_error("alignment index k = %d: only 4 <= k <= 16 is supported", k); return -1; }
	std::vector<KgTMeta> meta(DB);
	int64_t nslots = 0;
	meta[0] = KgTMeta{0, 0, 0, 31};
	for (int t = 1; t < DB; ++t) {
		const int64_t nk = std::max<int64_t>(0, (int64_t)db->lengths[t] - k + 1);
		int lg = 1;
		while ((1ll << lg) < 2 * nk) ++lg;
		meta[t] = KgTMeta{nslots, db->seq_off[t], db->lengths[t], 32 - lg};
		nslots += 1ll << lg;
	}
	KG_CUDA(cudaMalloc(&db->d_tmeta, sizeof(KgTMeta) * (size_t)DB));
	KG_CUDA(cudaMemcpy(db->d_tmeta, meta.data(), sizeof(KgTMeta) * (size_t)DB, cudaMemcpyHostToDevice));
	KG_CUDA(cudaMalloc(&db->d_tslots, 8 * (size_t)nslots + 8));
	KG_CUDA(cudaMemset(db->d_tslots, 0, 8 * (size_t)nslots + 8));
	unsigned long long *cursor = nullptr;
	KG_CUDA(cudaMalloc(&cursor, 8));
	KG_CUDA(cudaMemset(cursor, 0, 8));
	KgTIndexView ix{(const KgTMeta *)db->d_tmeta, (const uint2 *)db->d_tslots, nullptr, db->d_seq, k};
	const int64_t total_bases = (int64_t)db->seq_words * 32;
	const int grid = db->sm_count * 16;
	uint2 *slots = (uint2 *)db->d_tslots;
	if (DB > 1 && total_bases > 0) {
		tix_positions_kernel<<<grid, 256, 0, db->stream>>>(ix, slots, nullptr, DB, total_bases, 0);
		tix_assign_kernel<<<grid, 256, 0, db->stream>>>(slots, nslots, nullptr, cursor, 1);
		unsigned long long ndup = 0;
		KG_CUDA(cudaMemcpyAsync(&ndup, cursor, 8, cudaMemcpyDeviceToHost, db->stream));
		KG_CUDA(cudaStreamSynchronize(db->stream));
		KG_CUDA(cudaMalloc(&db->d_tdups, 4 * (size_t)ndup + 16));
		KG_CUDA(cudaMemsetAsync(cursor, 0, 8, db->stream));
		tix_assign_kernel<<<grid, 256, 0, db->stream>>>(slots, nslots, db->d_tdups, cursor, 0);
		tix_positions_kernel<<<grid, 256, 0, db->stream>>>(ix, slots, db->d_tdups, DB, total_bases, 2);
		tix_sort_kernel<<<grid, 256, 0, db->stream>>>(slots, nslots, db->d_tdups);
		KG_CUDA(cudaStreamSynchronize(db->stream));
		KG_CUDA(cudaGetLastError());
		db->info.device_bytes += 8 * (uint64_t)nslots + 4 * (uint64_t)ndup + sizeof(KgTMeta) * (uint64_t)DB;
	} else {
		KG_CUDA(cudaMalloc(&db->d_tdups, 16));
	}
	cudaFree(cursor);
	db->tix = KgTIndexView{(const KgTMeta *)db->d_tmeta, (const uint2 *)db->d_tslots, db->d_tdups, db->d_seq, k};
	return 0;
}
