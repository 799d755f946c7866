// comm.cu — multi-GPU entry points of the C ABI: one process per GPU, the graph replicated in every
// GPU's HBM, walks / queries split into contiguous ranges with GLOBAL ids feeding the counter-based RNG, and ONE
// exchange at the end: the result blocks are gathered over NCCL (NVLink / NVSwitch on a B200 box).  There is no
// data-path collective -- walks and queries are independent given the read-only graph (node2vec.py:53-57,
// SingleRandomWalk.java:39-45).
//
// NCCL is bound at run time (dlopen of libnccl.so.2) so that libgraphwalk.so keeps no link-time dependency: a
// single-GPU host (the reference's own usage) never touches it, and a process that already loaded a NCCL (e.g.
// torch's) shares that copy.  Only the stable C entry points are used; blocks are moved as bytes.
#include <dlfcn.h>

#include <cstring>
#include <vector>

#include "common.cuh"

namespace gw {

typedef struct { char internal[128]; } nccl_unique_id;      // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)
typedef void *nccl_comm_t;
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(nccl_unique_id *) = nullptr;
    int (*CommInitRank)(nccl_comm_t *, int, nccl_unique_id, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int /*ncclDataType_t*/, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.lib) return GW_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *nm : names)
        if ((lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!lib) return fail(GW_E_STATE, "NCCL is not available: %s", dlerror());
    NcclApi a;
    a.lib = lib;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(lib, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(lib, "ncclCommDestroy");
    a.Broadcast = (decltype(a.Broadcast))dlsym(lib, "ncclBroadcast");
    a.GroupStart = (decltype(a.GroupStart))dlsym(lib, "ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))dlsym(lib, "ncclGroupEnd");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.Broadcast || !a.GroupStart || !a.GroupEnd || !a.GetErrorString)
        return fail(GW_E_STATE, "libnccl lacks a required entry point");
    g_nccl = a;
    return GW_OK;
}

#define GW_NCCL(expr)                                                                                   \
    do {                                                                                                \
        int _r = (expr);                                                                                \
        if (_r != 0) return gw::fail(GW_E_CUDA, "%s failed: %s", #expr, gw::g_nccl.GetErrorString(_r)); \
    } while (0)

// contiguous slice [lo, hi) of n units for `rank`; sizes differ by at most one (dist.py shard_range)
static void shard_range(int64_t n, int rank, int world, int64_t *lo, int64_t *hi) {
    const int64_t base = n / world, rem = n % world;
    *lo = rank * base + (rank < rem ? rank : rem);
    *hi = *lo + base + (rank < rem ? 1 : 0);
}

}  // namespace gw

struct gw_comm {
    int rank = 0, nranks = 1, device = 0;
    gw::nccl_comm_t comm = nullptr;
    cudaStream_t stream = nullptr;
};

using namespace gw;

// every rank ends with all blocks: block r (count[r] bytes at offset off[r] of d_full) is broadcast from rank r
static int all_gather_blocks(gw_comm *c, void *d_full, const std::vector<size_t> &off, const std::vector<size_t> &cnt) {
    GW_NCCL(g_nccl.GroupStart());
    for (int r = 0; r < c->nranks; r++) {
        if (cnt[r] == 0) continue;
        char *p = (char *)d_full + off[r];
        GW_NCCL(g_nccl.Broadcast(p, p, cnt[r], 0 /* ncclInt8 */, r, c->comm, c->stream));
    }
    GW_NCCL(g_nccl.GroupEnd());
    GW_CUDA(cudaStreamSynchronize(c->stream));
    return GW_OK;
}

extern "C" {

int gw_comm_unique_id(void *id128) {
    if (!id128) return fail(GW_E_INVALID, "id buffer is NULL");
    GW_TRY(nccl_load());
    nccl_unique_id id;
    GW_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return GW_OK;
}

int gw_comm_init(int32_t rank, int32_t nranks, const void *id128, int32_t device, gw_comm **out) {
    if (!out || !id128) return fail(GW_E_INVALID, "bad arguments");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(GW_E_INVALID, "rank %d is outside [0, %d)", rank, nranks);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(GW_E_CUDA, "no CUDA device is visible (this library has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(GW_E_INVALID, "device %d is outside [0, %d)", device, ndev);
    GW_TRY(nccl_load());
    GW_CUDA(cudaSetDevice(device));
    gw_comm *c = new gw_comm;
    c->rank = rank; c->nranks = nranks; c->device = device;
    nccl_unique_id id;
    memcpy(&id, id128, sizeof(id));
    int r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
    if (r != 0) { delete c; return fail(GW_E_CUDA, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); }
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        g_nccl.CommDestroy(c->comm); delete c;
        return fail(GW_E_CUDA, "cudaStreamCreate failed");
    }
    *out = c;
    return GW_OK;
}

int gw_comm_info(const gw_comm *c, int32_t *rank, int32_t *nranks, int32_t *device) {
    if (!c) return fail(GW_E_INVALID, "communicator is NULL");
    if (rank) *rank = c->rank;
    if (nranks) *nranks = c->nranks;
    if (device) *device = c->device;
    return GW_OK;
}

int gw_comm_free(gw_comm *c) {
    if (!c) return GW_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    delete c;
    return GW_OK;
}

int gw_shard_range(int64_t n, int32_t rank, int32_t nranks, int64_t *lo, int64_t *hi) {
    if (n < 0 || nranks < 1 || rank < 0 || rank >= nranks || !lo || !hi) return fail(GW_E_INVALID, "bad arguments");
    shard_range(n, rank, nranks, lo, hi);
    return GW_OK;
}

int gw_node2vec_walks_sharded(gw_graph *g, gw_comm *c, double p, double q, int32_t walk_length, const int64_t *starts,
                              int64_t n_starts, uint64_t seed, int32_t gather, int32_t *out_walks, int32_t *out_lens) {
    if (!g || !c) return fail(GW_E_INVALID, "graph or communicator is NULL");
    if (g->device != c->device) return fail(GW_E_INVALID, "graph lives on device %d, communicator on %d", g->device, c->device);
    if (n_starts < 0 || walk_length < 1 || (n_starts > 0 && (!starts || !out_walks))) return fail(GW_E_INVALID, "bad arguments");
    if (n_starts == 0) return GW_OK;
    GW_CUDA(cudaSetDevice(c->device));
    int64_t lo, hi;
    shard_range(n_starts, c->rank, c->nranks, &lo, &hi);
    for (int64_t i = 0; i < n_starts; i++)                        // every rank checks ALL starts: ranks must fail together, never one inside NCCL
        if (starts[i] < 0 || starts[i] >= g->n) return fail(GW_E_KEY, "start node %lld is outside [0, %lld)", (long long)starts[i], (long long)g->n);
    const size_t L = (size_t)walk_length;
    DevBuf<int32_t> dw, dl;
    DevBuf<int64_t> ds;
    if (gather) {
        if (dw.alloc((size_t)n_starts * L) != cudaSuccess || dl.alloc((size_t)n_starts) != cudaSuccess) {
            cudaGetLastError();
            return fail(GW_E_TOO_LARGE, "the gathered corpus of %lld walks does not fit on the device; call with gather = 0", (long long)n_starts);
        }
    } else {
        GW_CUDA(dw.alloc((size_t)(hi - lo) * L)); GW_CUDA(dl.alloc((size_t)(hi - lo)));
    }
    int32_t *d_local = gather ? dw.p + (size_t)lo * L : dw.p, *d_local_lens = gather ? dl.p + lo : dl.p;
    if (hi > lo) {
        GW_CUDA(ds.alloc((size_t)(hi - lo)));
        GW_CUDA(cudaMemcpyAsync(ds.p, starts + lo, sizeof(int64_t) * (size_t)(hi - lo), cudaMemcpyHostToDevice, c->stream));
        GW_TRY(gw_node2vec_walks_dev(g, p, q, walk_length, ds.p, hi - lo, seed, (uint64_t)lo, d_local, d_local_lens, c->stream));
    }
    if (!gather) {                                               // this rank's block only, at its global offset
        GW_CUDA(cudaMemcpyAsync(out_walks + (size_t)lo * L, d_local, sizeof(int32_t) * (size_t)(hi - lo) * L, cudaMemcpyDeviceToHost, c->stream));
        if (out_lens) GW_CUDA(cudaMemcpyAsync(out_lens + lo, d_local_lens, sizeof(int32_t) * (size_t)(hi - lo), cudaMemcpyDeviceToHost, c->stream));
        GW_CUDA(cudaStreamSynchronize(c->stream));
        return GW_OK;
    }
    std::vector<size_t> off(c->nranks), cnt(c->nranks), off2(c->nranks), cnt2(c->nranks);
    for (int r = 0; r < c->nranks; r++) {
        int64_t a, b;
        shard_range(n_starts, r, c->nranks, &a, &b);
        off[r] = (size_t)a * L * sizeof(int32_t); cnt[r] = (size_t)(b - a) * L * sizeof(int32_t);
        off2[r] = (size_t)a * sizeof(int32_t); cnt2[r] = (size_t)(b - a) * sizeof(int32_t);
    }
    GW_TRY(all_gather_blocks(c, dw.p, off, cnt));
    GW_TRY(all_gather_blocks(c, dl.p, off2, cnt2));
    GW_CUDA(cudaMemcpy(out_walks, dw.p, sizeof(int32_t) * (size_t)n_starts * L, cudaMemcpyDeviceToHost));
    if (out_lens) GW_CUDA(cudaMemcpy(out_lens, dl.p, sizeof(int32_t) * (size_t)n_starts, cudaMemcpyDeviceToHost));
    return GW_OK;
}

int gw_simrank_topk_sharded(gw_graph *g, gw_comm *c, const int64_t *queries, int64_t nq, double cdecay, int32_t step,
                            int32_t sample, int32_t k, int32_t mode, uint64_t seed, int32_t *out_ids, double *out_scores) {
    if (!g || !c) return fail(GW_E_INVALID, "graph or communicator is NULL");
    if (g->device != c->device) return fail(GW_E_INVALID, "graph lives on device %d, communicator on %d", g->device, c->device);
    if (nq < 0 || k < 1 || (nq > 0 && (!queries || !out_ids || !out_scores))) return fail(GW_E_INVALID, "bad arguments");
    for (int64_t i = 0; i < nq; i++)
        if (queries[i] < 0 || queries[i] >= g->n) return fail(GW_E_KEY, "query vertex %lld is outside [0, %lld)", (long long)queries[i], (long long)g->n);
    if (nq == 0) return GW_OK;
    GW_CUDA(cudaSetDevice(c->device));
    int64_t lo, hi;
    shard_range(nq, c->rank, c->nranks, &lo, &hi);
    DevBuf<int32_t> di;
    DevBuf<double> dsc;
    DevBuf<int64_t> dq;
    GW_CUDA(di.alloc((size_t)nq * k)); GW_CUDA(dsc.alloc((size_t)nq * k));
    if (hi > lo) {
        GW_CUDA(dq.alloc((size_t)(hi - lo)));
        GW_CUDA(cudaMemcpyAsync(dq.p, queries + lo, sizeof(int64_t) * (size_t)(hi - lo), cudaMemcpyHostToDevice, c->stream));
        GW_TRY(gw_simrank_topk_dev(g, dq.p, hi - lo, cdecay, step, sample, k, mode, seed, (uint64_t)lo, di.p + (size_t)lo * k,
                                   dsc.p + (size_t)lo * k, c->stream));
    }
    std::vector<size_t> off(c->nranks), cnt(c->nranks), off2(c->nranks), cnt2(c->nranks);
    for (int r = 0; r < c->nranks; r++) {
        int64_t a, b;
        shard_range(nq, r, c->nranks, &a, &b);
        off[r] = (size_t)a * k * sizeof(int32_t); cnt[r] = (size_t)(b - a) * k * sizeof(int32_t);
        off2[r] = (size_t)a * k * sizeof(double); cnt2[r] = (size_t)(b - a) * k * sizeof(double);
    }
    GW_TRY(all_gather_blocks(c, di.p, off, cnt));
    GW_TRY(all_gather_blocks(c, dsc.p, off2, cnt2));
    GW_CUDA(cudaMemcpy(out_ids, di.p, sizeof(int32_t) * (size_t)nq * k, cudaMemcpyDeviceToHost));
    GW_CUDA(cudaMemcpy(out_scores, dsc.p, sizeof(double) * (size_t)nq * k, cudaMemcpyDeviceToHost));
    return GW_OK;
}

}  // extern "C"
