// Database side of libkmagpu.so: parse the reference's on-disk index and make it HBM resident.
//
// On-disk format stays the reference's (.comp.b hashmapkma.c:275-455, .length.b makeindex.c:263-272,
// .seq.b updateindex.c:172 / runkma.c:216-220). The DEVICE layout is ours: key_index[] and
// value_index[] are fused into one 8-byte {key, value offset} array so that a hit costs two
// dependent sectors (exist -> kv) instead of three.
#include "kmagpu_internal.h"
#include <errno.h>
#include <string.h>
#include <stdlib.h>
#include <mutex>
#include <thread>
#include <algorithm>

static std::mutex g_image_mutex;

static thread_local char g_err[1024] = "";

void kmagpu_set_error(const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

extern "C" const char *kmagpu_last_error(void) { return g_err; }

extern "C" int kmagpu_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

extern "C" void kmagpu_default_params(kmagpu_params *p) {
	memset(p, 0, sizeof(*p));
	p->M = 1; p->MM = -2; p->U = -1; p->W1 = -3; p->Wl = -6; p->Mn = 0; p->PE = 7;
	for (int i = 0; i < 4; ++i)
		for (int j = 0; j < 4; ++j) p->d[i * 5 + j] = i == j ? 1 : -2;
	p->scoreT = 0.5;
	p->minFrac = 1.0;
	p->mrc = 0.0;
	p->minlen = 16;
	p->coverT = 0.1;
	p->counters = 1;
}

int KgBuf::reserve(size_t bytes) {
	if (bytes <= cap && p) return 0;
	release();
	size_t want = bytes + bytes / 4 + 256;
	cudaError_t e = pinned ? cudaMallocHost(&p, want) : cudaMalloc(&p, want);
	if (e != cudaSuccess) {
		p = nullptr; cap = 0;
		kmagpu_set_error("allocation of %zu %s bytes failed: %s", want, pinned ? "pinned" : "device", cudaGetErrorString(e));
		return -1;
	}
	cap = want;
	return 0;
}

void KgBuf::release() {
	if (p) { if (pinned) cudaFreeHost(p); else cudaFree(p); }
	p = nullptr; cap = 0;
}

static int read_all(FILE *f, void *dst, size_t n) { return fread(dst, 1, n, f) == n ? 0 : -1; }

extern "C" int kmagpu_db_open(const char *prefix, int device, kmagpu_db **out) {
	*out = nullptr;
	int ndev = kmagpu_device_count();
	if (ndev <= 0) { kmagpu_set_error("no CUDA device: libkmagpu has no CPU fallback"); return -1; }
	if (device < 0 || device >= ndev) { kmagpu_set_error("device %d out of range (%d devices)", device, ndev); return -1; }
	KG_CUDA(cudaSetDevice(device));

	std::string path = std::string(prefix) + ".comp.b";
	FILE *f = fopen(path.c_str(), "rb");
	if (!f) { kmagpu_set_error("%s: %s", path.c_str(), strerror(errno)); return -1; }
	uint32_t h32[3]; uint64_t h64[5];
	if (read_all(f, h32, 12) || read_all(f, h64, 40)) { fclose(f); kmagpu_set_error("%s: truncated header", path.c_str()); return -1; }
	const uint32_t DB_size = h32[0], mlen = h32[1], prefix_len = h32[2];
	const uint64_t pfx = h64[0], size = h64[1], n = h64[2], v_index = h64[3], null_index = h64[4];
	if (size < n || size == 0 || (size & (size - 1))) { fclose(f); kmagpu_set_error("%s: wrong format of DB", path.c_str()); return -1; }
	if (prefix_len || pfx) { fclose(f); kmagpu_set_error("%s: sparse (prefix) databases are out of scope", path.c_str()); return -1; }
	if (mlen > 16) { fclose(f); kmagpu_set_error("%s: k > 16 (64-bit keys) not supported yet", path.c_str()); return -1; }
	const uint64_t kmask = (1ull << (2 * mlen)) - 1;
	const bool mega = (size - 1) == kmask;
	if ((mega ? v_index : n) > 0xFFFFFFFFull || v_index >= 0xFFFFFFFFull) { fclose(f); kmagpu_set_error("%s: 64-bit indexes not supported yet", path.c_str()); return -1; }
	const bool vshort = DB_size < 65535;

	kmagpu_db *db = new kmagpu_db();
	db->device = device;
	db->image = new KgImageRef();
	std::vector<uint32_t> exist(size), keys, vidx;
	std::vector<uint8_t> values(v_index * (vshort ? 2 : 4));
	int bad = read_all(f, exist.data(), size * 4) || read_all(f, values.data(), values.size());
	if (!bad && !mega) {
		keys.resize(n + 1); vidx.resize(n);
		bad = read_all(f, keys.data(), (n + 1) * 4) || read_all(f, vidx.data(), n * 4);
	}
	uint32_t tail[2] = {mlen, 0};
	if (!bad && read_all(f, tail, 8)) { tail[0] = mlen; tail[1] = 0; }
	fclose(f);
	if (bad) { delete db; kmagpu_set_error("%s: truncated", path.c_str()); return -1; }
	if (tail[1] != 0 || tail[0] != mlen) { delete db; kmagpu_set_error("%s: minimizer / homopolymer databases (flag != 0) are out of scope", path.c_str()); return -1; }

	db->info.DB_size = (int32_t)DB_size;
	db->info.kmersize = (int32_t)tail[0];
	db->info.mega = mega;
	db->info.size = size; db->info.n = n; db->info.v_index = v_index;

	auto fail = [&](const char *what) { kmagpu_db_close(db); if (what) kmagpu_set_error("%s", what); return -1; };
	size_t dev_bytes = 0;
	if (mega) {
		if (cudaMalloc(&db->d_exist, size * 4) != cudaSuccess) return fail("cudaMalloc exist");
		if (cudaMemcpy(db->d_exist, exist.data(), size * 4, cudaMemcpyHostToDevice) != cudaSuccess) return fail("H2D exist");
		dev_bytes += size * 4;
	}
	if (cudaMalloc(&db->d_values, values.size() + 64) != cudaSuccess) return fail("cudaMalloc values");
	if (cudaMemcpy(db->d_values, values.data(), values.size(), cudaMemcpyHostToDevice) != cudaSuccess) return fail("H2D values");
	dev_bytes += values.size();
	if (!mega) {
		std::vector<uint2> kv(n + 1);
		for (uint64_t i = 0; i < n; ++i) kv[i] = make_uint2(keys[i], vidx[i]);
		kv[n] = make_uint2(keys[n], 0xFFFFFFFFu);
		if (cudaMalloc(&db->d_kv, (n + 1) * 8) != cudaSuccess) return fail("cudaMalloc kv");
		if (cudaMemcpy(db->d_kv, kv.data(), (n + 1) * 8, cudaMemcpyHostToDevice) != cudaSuccess) return fail("H2D kv");
		dev_bytes += (n + 1) * 8;
		// bucket entries {key0, value0, pos, cnt}: the reference's lookup (hashmapkma.c:149-178) reads exist[bucket], then the
		// keys from there on while they mismatch, stay in the bucket and lie below n. cnt = how many entries such a scan examines
		// when every key mismatches, so a device lookup that examines entries pos .. pos + cnt - 1 in order returns what it returns.
		// The first entry sits in the bucket itself: a probe of an empty or single-key bucket and a hit on a first key cost one
		// 16-byte load instead of two dependent ones. (d_exist holds this array for hashed tables.)
		std::vector<uint4> bk(size);
		const uint32_t hm = (uint32_t)(size - 1);
		auto fill = [&](uint64_t b0, uint64_t b1) {
			for (uint64_t b = b0; b < b1; ++b) {
				const uint32_t pos = exist[b];
				if (pos == (uint32_t)null_index || pos > n) { bk[b] = make_uint4(0, 0, 0, 0); continue; }
				uint32_t p = pos, cnt = 1;
				while ((kv[p].x & hm) == (uint32_t)b && p < n) { ++p; ++cnt; }
				bk[b] = make_uint4(kv[pos].x, kv[pos].y, pos, cnt);
			}
		};
		{
			const unsigned nth = size >= (1u << 22) ? std::min(8u, std::max(1u, std::thread::hardware_concurrency())) : 1u;
			std::vector<std::thread> th;
			for (unsigned t = 1; t < nth; ++t) th.emplace_back(fill, size * t / nth, size * (t + 1) / nth);
			fill(0, size / nth);
			for (auto &x : th) x.join();
		}
		if (cudaMalloc(&db->d_exist, size * 16) != cudaSuccess) return fail("cudaMalloc buckets");
		if (cudaMemcpy(db->d_exist, bk.data(), size * 16, cudaMemcpyHostToDevice) != cudaSuccess) return fail("H2D buckets");
		dev_bytes += size * 16;
	}
	db->hv.exist = mega ? (const uint32_t *)db->d_exist : nullptr;
	db->hv.bk = mega ? nullptr : (const uint4 *)db->d_exist;
	db->hv.kv = (const uint2 *)db->d_kv;
	db->hv.values_s = vshort ? (const uint16_t *)db->d_values : nullptr;
	db->hv.values_w = vshort ? nullptr : (const uint32_t *)db->d_values;
	db->hv.hmask = mega ? kmask : size - 1;
	db->hv.null_index = (uint32_t)null_index;
	db->hv.n = (uint32_t)n;
	db->hv.kmersize = (int32_t)tail[0];
	db->hv.DB_size = (int32_t)DB_size;
	db->hv.mega = mega;

	// .length.b: int32 DB_size, int32 len[DB_size] (len[0] = k of the alignment index)
	path = std::string(prefix) + ".length.b";
	if ((f = fopen(path.c_str(), "rb"))) {
		int32_t cnt = 0;
		if (read_all(f, &cnt, 4) == 0 && cnt == (int32_t)DB_size) {
			db->lengths.resize(cnt);
			if (read_all(f, db->lengths.data(), 4 * (size_t)cnt)) db->lengths.clear();
		}
		fclose(f);
	}
	if (!db->lengths.empty()) {
		db->info.kmerindex = db->lengths[0];
		// word offsets exactly as runkma.c:216-220 derives seq_indexes
		db->seq_off.assign(DB_size + 1, 0);
		uint64_t bases = 0;
		for (uint32_t t = 1; t < DB_size; ++t) {
			db->seq_off[t + 1] = db->seq_off[t] + ((db->lengths[t] >> 5) + 1);
			bases += db->lengths[t];
		}
		db->info.seq_bases = bases;
		path = std::string(prefix) + ".seq.b";
		if ((f = fopen(path.c_str(), "rb"))) {
			size_t words = (size_t)db->seq_off[DB_size];
			std::vector<uint64_t> seq(words + 2, 0);
			if (read_all(f, seq.data(), words * 8) == 0) {
				db->seq_words = words;
				if (cudaMalloc(&db->d_seq, (words + 2) * 8) != cudaSuccess) { fclose(f); return fail("cudaMalloc seq"); }
				cudaMemcpy(db->d_seq, seq.data(), (words + 2) * 8, cudaMemcpyHostToDevice);
				cudaMalloc(&db->d_lengths, 4 * (size_t)DB_size);
				cudaMemcpy(db->d_lengths, db->lengths.data(), 4 * (size_t)DB_size, cudaMemcpyHostToDevice);
				cudaMalloc(&db->d_seq_off, 8 * (size_t)(DB_size + 1));
				cudaMemcpy(db->d_seq_off, db->seq_off.data(), 8 * (size_t)(DB_size + 1), cudaMemcpyHostToDevice);
				dev_bytes += (words + 2) * 8 + 12 * (size_t)DB_size;
			}
			fclose(f);
		}
	}
	db->info.device_bytes = dev_bytes;

	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) db->sm_count = prop.multiProcessorCount;
	if (cudaStreamCreateWithFlags(&db->stream, cudaStreamNonBlocking) != cudaSuccess) return fail("stream");
	for (auto &e : db->ev) if (cudaEventCreate(&e) != cudaSuccess) return fail("event");
	if (cudaGetLastError() != cudaSuccess) return fail("CUDA error while loading the database");
	if (db->d_seq && kg_tindex_build(db)) return fail(nullptr);
	*out = db;
	return 0;
}

// A second handle on the same HBM image: own stream, events and batch buffers (so that it can run concurrently with
// the first from another host thread), no second copy of the hash table, the sequences or the position index.
extern "C" int kmagpu_db_clone(kmagpu_db *src, kmagpu_db **out) {
	if (!src || !out) { kmagpu_set_error("null argument"); return -1; }
	*out = nullptr;
	KG_CUDA(cudaSetDevice(src->device));
	kmagpu_db *db = new kmagpu_db();
	db->device = src->device; db->image = src->image; db->info = src->info; db->hv = src->hv;
	db->d_exist = src->d_exist; db->d_kv = src->d_kv; db->d_values = src->d_values;
	db->lengths = src->lengths; db->seq_off = src->seq_off;
	db->d_seq = src->d_seq; db->d_lengths = src->d_lengths; db->d_seq_off = src->d_seq_off; db->seq_words = src->seq_words;
	db->sm_count = src->sm_count; db->conclave_lc = src->conclave_lc;
	db->d_tmeta = src->d_tmeta; db->d_tslots = src->d_tslots; db->d_tdups = src->d_tdups; db->tix = src->tix;
	{ std::lock_guard<std::mutex> g(g_image_mutex); ++db->image->refs; }
	bool ok = cudaStreamCreateWithFlags(&db->stream, cudaStreamNonBlocking) == cudaSuccess;
	for (auto &e : db->ev) ok = ok && cudaEventCreate(&e) == cudaSuccess;
	if (!ok) { kmagpu_db_close(db); kmagpu_set_error("stream / event creation failed"); return -1; }
	*out = db;
	return 0;
}

extern "C" void kmagpu_db_close(kmagpu_db *db) {
	if (!db) return;
	cudaSetDevice(db->device);
	kg_seed_free(db);
	kg_stage1_free(db);
	kg_memscore_free(db);
	kg_align_free(db);
	db->d_cons_rows.release(); db->d_cons_stat.release();
	bool last = true;
	KgImageRef *img = db->image;
	if (img) {
		std::lock_guard<std::mutex> g(g_image_mutex);
		last = --img->refs == 0;
	}
	if (last) {
		cudaFree(db->d_tmeta); cudaFree(db->d_tslots); cudaFree(db->d_tdups);
		cudaFree(db->d_exist); cudaFree(db->d_kv); cudaFree(db->d_values);
		cudaFree(db->d_seq); cudaFree(db->d_lengths); cudaFree(db->d_seq_off);
		if (img) {
			kmagpu_comm_destroy(db);
			cudaFree(img->d_mat); cudaFree(img->d_mat_off); cudaFree(img->d_run_scores); cudaFree(img->d_soft);
			delete img;
		}
	}
	for (auto &e : db->ev) if (e) cudaEventDestroy(e);
	if (db->stream) cudaStreamDestroy(db->stream);
	delete db;
}

// Host-side validation of one record of a caller-supplied stream, so that the kernels can trust its fields: the sequence
// fits its packed words, the N positions are ascending and inside the read, the template ids name templates of this
// database. (Streams that a previous stage left in HBM are the library's own output and are not re-checked.)
int kg_check_record(const uint8_t *rec, int stage, int DB_size, size_t at) {
	int32_t h[7];
	memcpy(h, rec, stage == 1 ? 16 : 28);
	const int64_t seqlen = h[0], words = h[1], nN = h[2];
	if (seqlen < 0 || seqlen > 32 * words) {
		kmagpu_set_error("record at byte %zu: %lld bases do not fit %lld packed words", at, (long long)seqlen, (long long)words);
		return -1;
	}
	const uint8_t *N = rec + (stage == 1 ? 16 : 28) + 8 * (size_t)words;
	int32_t prev = -1;
	for (int64_t i = 0; i < nN; ++i) {
		int32_t v;
		memcpy(&v, N + 4 * (size_t)i, 4);
		if (v <= prev || v >= seqlen) { kmagpu_set_error("record at byte %zu: N position %d outside the read or out of order", at, v); return -1; }
		prev = v;
	}
	if (stage == 2) {
		const uint8_t *T = N + 4 * (size_t)nN;
		for (int32_t i = 0; i < h[4]; ++i) {
			int32_t t;
			memcpy(&t, T + 4 * (size_t)i, 4);
			if (t == 0 || t <= -DB_size || t >= DB_size) { kmagpu_set_error("record at byte %zu names template %d outside the database", at, t); return -1; }
		}
	}
	return 0;
}

extern "C" int kmagpu_db_get_info(const kmagpu_db *db, kmagpu_db_info *info) {
	if (!db || !info) { kmagpu_set_error("null argument"); return -1; }
	*info = db->info;
	return 0;
}

extern "C" int64_t kmagpu_record_walk(int stage, const void *buf, size_t nbytes, uint64_t *offsets, size_t cap, size_t *used) {
	const uint8_t *in = (const uint8_t *)buf;
	const size_t hdr = stage == 1 ? 16 : (stage == 2 ? 28 : (stage == 3 ? 32 : 20));
	size_t ip = 0;
	int64_t n = 0;
	if (stage < 1 || stage > 4) { kmagpu_set_error("kmagpu_record_walk: stage must be 1, 2, 3 or 4"); return -1; }
	while (ip + hdr <= nbytes) {
		int32_t h[8];
		memcpy(h, in + ip, hdr);
		if (h[0] < 0 || (stage == 4 && h[0] == 0)) break;   // stream terminator
		size_t len;
		if (stage == 4) {   // frag_raw record (updatescores.c:284-295); a negative score means the mate block follows
			if (h[3] < 0) { kmagpu_set_error("corrupt frag_raw record at byte %zu", ip); return -1; }
			len = 20 + (size_t)h[0] + (size_t)h[3] + 12 * (size_t)abs(h[1]);
			if (h[2] < 0) {
				if (ip + len + 12 > nbytes) break;
				int32_t m[3];
				memcpy(m, in + ip + len, 12);
				if (m[0] < 0 || m[1] < 0) { kmagpu_set_error("corrupt frag_raw mate block at byte %zu", ip + len); return -1; }
				len += 12 + (size_t)m[0] + (size_t)m[1];
			}
		} else if (stage == 1) {
			if (h[1] < 0 || h[2] < 0) { kmagpu_set_error("corrupt stage-1 record at byte %zu", ip); return -1; }
			len = 16 + 8 * (size_t)h[1] + 4 * (size_t)h[2] + (size_t)abs(h[3]);
		} else if (stage == 2) {
			if (h[1] < 0 || h[2] < 0 || h[4] < 0 || h[5] < 0) { kmagpu_set_error("corrupt stage-2 record at byte %zu", ip); return -1; }
			len = 28 + 8 * (size_t)h[1] + 4 * (size_t)h[2] + 4 * (size_t)h[4] + (size_t)h[5];
		} else {   // per-template fragment record of the assembly pass (frags.c:45-48)
			if (h[1] < 0 || h[6] < 0) { kmagpu_set_error("corrupt fragment record at byte %zu", ip); return -1; }
			len = 32 + (size_t)h[1] + (size_t)h[6];
		}
		if (ip + len > nbytes) break;   // partial record: the caller refills
		if (offsets && (size_t)n < cap) offsets[n] = ip;
		++n;
		ip += len;
	}
	if (used) *used = ip;
	return n;
}
