// Multi-GPU exchange of the mapping core INSIDE the library: one process per GPU, reads sharded by rank,
// database replicated; the only exchanges are sums over ranks of
//   * the two ConClave accumulators alignment_scores / uniq_alignment_scores [DB_size] u64 (runkma.c:98-99; added to by
//     update_Scores, updatescores.c:228/276; read globally by ConClave's choice pass, conclave.c:80-123), and
//   * the base-count matrix of the assembly pass (assembly.c:1436: +1 per aligned base, so unsaturated sums commute).
// Both stay in HBM: ncclAllReduce runs in place on the handle's stream, behind the kernels that produced the numbers, and
// kmagpu_conclave_* reads the reduced sums from the device (alignment_scores = NULL). NCCL is bound at run time
// (dlopen of libnccl.so.2: the copy the process already has, e.g. PyTorch's, or the system's), so the library carries no
// link-time dependency and single-GPU hosts never load it.
#include "kmagpu_internal.h"
#include <dlfcn.h>
#include <string.h>
#include <mutex>

// the stable part of nccl.h this file needs
typedef struct { char internal[128]; } kgNcclUniqueId;
typedef void *kgNcclComm;
enum { KG_NCCL_SUM = 0, KG_NCCL_UINT32 = 3, KG_NCCL_UINT64 = 5 };

static struct {
	void *lib;
	int (*GetUniqueId)(kgNcclUniqueId *);
	int (*CommInitRank)(kgNcclComm *, int, kgNcclUniqueId, int);
	int (*AllReduce)(const void *, void *, size_t, int, int, kgNcclComm, cudaStream_t);
	int (*CommDestroy)(kgNcclComm);
	const char *(*GetErrorString)(int);
} g_nccl;
static std::mutex g_nccl_mutex;

static int nccl_load() {
	std::lock_guard<std::mutex> g(g_nccl_mutex);
	if (g_nccl.lib) return 0;
	void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
	if (!h) { kmagpu_set_error("NCCL not found: %s", dlerror()); return -1; }
	*(void **)&g_nccl.GetUniqueId = dlsym(h, "ncclGetUniqueId");
	*(void **)&g_nccl.CommInitRank = dlsym(h, "ncclCommInitRank");
	*(void **)&g_nccl.AllReduce = dlsym(h, "ncclAllReduce");
	*(void **)&g_nccl.CommDestroy = dlsym(h, "ncclCommDestroy");
	*(void **)&g_nccl.GetErrorString = dlsym(h, "ncclGetErrorString");
	if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy || !g_nccl.GetErrorString) {
		kmagpu_set_error("libnccl lacks an entry point this library needs");
		return -1;
	}
	g_nccl.lib = h;
	return 0;
}

#define KG_NCCL(call)                                                                                   \
	do {                                                                                                \
		const int r__ = (call);                                                                         \
		if (r__ != 0) { kmagpu_set_error("%s -> %s", #call, g_nccl.GetErrorString(r__)); return -1; }  \
	} while (0)

extern "C" int kmagpu_comm_unique_id(void *id, size_t cap) {
	if (!id || cap < sizeof(kgNcclUniqueId)) { kmagpu_set_error("kmagpu_comm_unique_id needs a 128-byte buffer"); return -1; }
	if (nccl_load()) return -1;
	KG_NCCL(g_nccl.GetUniqueId((kgNcclUniqueId *)id));
	return 0;
}

extern "C" int kmagpu_comm_init(kmagpu_db *db, const void *id, int rank, int world) {
	if (!db || !id || world < 1 || rank < 0 || rank >= world) { kmagpu_set_error("bad argument"); return -1; }
	KgImageRef *img = db->image;
	if (img->comm) { kmagpu_set_error("this database image already has a communicator"); return -1; }
	img->comm_rank = rank; img->comm_world = world;
	if (world == 1) return 0;
	if (nccl_load()) return -1;
	KG_CUDA(cudaSetDevice(db->device));
	kgNcclUniqueId uid;
	memcpy(&uid, id, sizeof(uid));
	KG_NCCL(g_nccl.CommInitRank((kgNcclComm *)&img->comm, world, uid, rank));
	return 0;
}

extern "C" void kmagpu_comm_destroy(kmagpu_db *db) {
	if (!db || !db->image) return;
	KgImageRef *img = db->image;
	if (img->comm && g_nccl.CommDestroy) { cudaSetDevice(db->device); g_nccl.CommDestroy((kgNcclComm)img->comm); }
	img->comm = nullptr; img->comm_world = 1; img->comm_rank = 0;
}

// the run-wide ConClave accumulators on the device: [alignment_scores[DB], uniq_alignment_scores[DB]], one set per database
// image -- every handle of the image (kmagpu_db_clone) adds its batches to them
extern "C" int kmagpu_scores_reset(kmagpu_db *db) {
	if (!db) { kmagpu_set_error("null argument"); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	KgImageRef *img = db->image;
	const size_t bytes = 16 * (size_t)db->info.DB_size;
	if (!img->d_run_scores) {
		unsigned long long *p = nullptr;
		KG_CUDA(cudaMalloc(&p, bytes));
		img->d_run_scores = p;
	}
	KG_CUDA(cudaMemsetAsync(img->d_run_scores, 0, bytes, db->stream));
	KG_CUDA(cudaStreamSynchronize(db->stream));
	return 0;
}

static __global__ void add_u64_kernel(unsigned long long *dst, const unsigned long long *src, size_t n) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && src[i]) atomicAdd(&dst[i], src[i]);   // handles of one image add from their own streams
}

// called by kmagpu_align_run / the -mem_mode score collection: this batch's sums join the run's (stream-ordered)
int kg_scores_accumulate(kmagpu_db *db, const unsigned long long *batch_scores) {
	if (!db->image->d_run_scores) return 0;
	const size_t n = 2 * (size_t)db->info.DB_size;
	add_u64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, db->stream>>>(db->image->d_run_scores, batch_scores, n);
	return 0;
}

// soft proximity sums of the run (kmers.c:133-153): one array per database image, every handle's batches join it
extern "C" int kmagpu_softproxi_reset(kmagpu_db *db) {
	if (!db) { kmagpu_set_error("null argument"); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	KgImageRef *img = db->image;
	const size_t bytes = 8 * (size_t)db->info.DB_size;
	if (!img->d_soft) {
		unsigned long long *p = nullptr;
		KG_CUDA(cudaMalloc(&p, bytes));
		img->d_soft = p;
	}
	KG_CUDA(cudaMemsetAsync(img->d_soft, 0, bytes, db->stream));
	KG_CUDA(cudaStreamSynchronize(db->stream));
	return 0;
}

int kg_softproxi_accumulate(kmagpu_db *db, const unsigned long long *batch_sums) {
	if (!db->image->d_soft) return 0;
	const size_t n = (size_t)db->info.DB_size;
	add_u64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, db->stream>>>(db->image->d_soft, batch_sums, n);
	return 0;
}

extern "C" int kmagpu_softproxi_download(kmagpu_db *db, uint64_t *sums) {
	if (!db || !sums) { kmagpu_set_error("null argument"); return -1; }
	if (!db->image->d_soft) { kmagpu_set_error("kmagpu_softproxi_download before kmagpu_softproxi_reset"); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	KG_CUDA(cudaMemcpyAsync(sums, db->image->d_soft, 8 * (size_t)db->info.DB_size, cudaMemcpyDeviceToHost, db->stream));
	KG_CUDA(cudaStreamSynchronize(db->stream));
	return 0;
}

extern "C" int kmagpu_allreduce_scores(kmagpu_db *db, uint64_t *alignment_scores, uint64_t *uniq_alignment_scores, float *ms) {
	if (!db) { kmagpu_set_error("null argument"); return -1; }
	KgImageRef *img = db->image;
	if (!img->d_run_scores) { kmagpu_set_error("kmagpu_allreduce_scores before kmagpu_scores_reset"); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	const size_t DB = (size_t)db->info.DB_size;
	if (ms) KG_CUDA(cudaEventRecord(db->ev[5], db->stream));
	if (img->comm) KG_NCCL(g_nccl.AllReduce(img->d_run_scores, img->d_run_scores, 2 * DB, KG_NCCL_UINT64, KG_NCCL_SUM, (kgNcclComm)img->comm, db->stream));
	if (ms) KG_CUDA(cudaEventRecord(db->ev[6], db->stream));
	if (alignment_scores) KG_CUDA(cudaMemcpyAsync(alignment_scores, img->d_run_scores, 8 * DB, cudaMemcpyDeviceToHost, db->stream));
	if (uniq_alignment_scores) KG_CUDA(cudaMemcpyAsync(uniq_alignment_scores, img->d_run_scores + DB, 8 * DB, cudaMemcpyDeviceToHost, db->stream));
	KG_CUDA(cudaStreamSynchronize(db->stream));
	if (ms) cudaEventElapsedTime(ms, db->ev[5], db->ev[6]);
	return 0;
}

extern "C" int kmagpu_allreduce_matrix(kmagpu_db *db, float *ms) {
	if (!db) { kmagpu_set_error("null argument"); return -1; }
	KgImageRef *img = db->image;
	if (!img->d_mat) { kmagpu_set_error("kmagpu_allreduce_matrix before any alignment was added to the matrix"); return -1; }
	KG_CUDA(cudaSetDevice(db->device));
	if (ms) KG_CUDA(cudaEventRecord(db->ev[5], db->stream));
	if (img->comm) KG_NCCL(g_nccl.AllReduce(img->d_mat, img->d_mat, img->mat_entries, KG_NCCL_UINT32, KG_NCCL_SUM, (kgNcclComm)img->comm, db->stream));
	if (ms) KG_CUDA(cudaEventRecord(db->ev[6], db->stream));
	KG_CUDA(cudaStreamSynchronize(db->stream));
	if (ms) cudaEventElapsedTime(ms, db->ev[5], db->ev[6]);
	return 0;
}

// any host array of u64 counters (w_scores, read counts widened by the caller ...): up, summed over ranks, down
extern "C" int kmagpu_allreduce_u64(kmagpu_db *db, uint64_t *buf, size_t n) {
	if (!db || (!buf && n)) { kmagpu_set_error("null argument"); return -1; }
	if (!db->image->comm || !n) return 0;
	KG_CUDA(cudaSetDevice(db->device));
	KgBuf tmp;
	if (tmp.reserve(8 * n)) return -1;
	cudaError_t e = cudaMemcpyAsync(tmp.p, buf, 8 * n, cudaMemcpyHostToDevice, db->stream);
	int r = 0;
	if (e == cudaSuccess) r = g_nccl.AllReduce(tmp.p, tmp.p, n, KG_NCCL_UINT64, KG_NCCL_SUM, (kgNcclComm)db->image->comm, db->stream);
	if (e == cudaSuccess && !r) e = cudaMemcpyAsync(buf, tmp.p, 8 * n, cudaMemcpyDeviceToHost, db->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(db->stream);
	tmp.release();
	if (r) { kmagpu_set_error("ncclAllReduce -> %s", g_nccl.GetErrorString(r)); return -1; }
	if (e != cudaSuccess) { kmagpu_set_error("kmagpu_allreduce_u64: %s", cudaGetErrorString(e)); return -1; }
	return 0;
}
