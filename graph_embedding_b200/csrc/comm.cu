// comm.cu — multi-GPU entry points of the C ABI: one process per GPU, the graph replicated in every
// GPU's HBM, walks / queries split into contiguous ranges with GLOBAL ids feeding the counter-based RNG, and ONE
// exchange at the end: the result blocks are gathered over NCCL (NVLink / NVSwitch on a B200 box).  There is no
// data-path collective -- walks and queries are independent given the read-only graph (node2vec.py:53-57,
// SingleRandomWalk.java:39-45).
//
// Collective-call discipline: arguments that every rank passes alike are checked first, identically everywhere; what
// can only be known per rank (the range check of its own slice, its allocations, its launch status, an accumulator
// overflow) is AGREED on with a one-int all-reduce before any rank enters the data exchange, so a failure on one
// rank is returned by all of them and nobody is left waiting inside NCCL.
//
// NCCL is bound at run time (dlopen of libnccl.so.2) so that libgraphwalk.so keeps no link-time dependency: a
// single-GPU host (the reference's own usage) never touches it, and a process that already loaded a NCCL (e.g.
// torch's) shares that copy.  Only the stable C entry points are used; blocks are moved as bytes.
#include <dlfcn.h>

#include <cstring>
#include <string>
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
    int (*AllReduce)(const void *, void *, size_t, int /*ncclDataType_t*/, int /*ncclRedOp_t*/, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
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
    a.AllReduce = (decltype(a.AllReduce))dlsym(lib, "ncclAllReduce");
    a.Send = (decltype(a.Send))dlsym(lib, "ncclSend");
    a.Recv = (decltype(a.Recv))dlsym(lib, "ncclRecv");
    a.GroupStart = (decltype(a.GroupStart))dlsym(lib, "ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))dlsym(lib, "ncclGroupEnd");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.Broadcast || !a.AllReduce || !a.Send || !a.Recv || !a.GroupStart ||
        !a.GroupEnd || !a.GetErrorString)
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
    // grow-only device workspace: repeated sharded calls allocate nothing
    void *ws_corpus = nullptr; size_t ws_corpus_bytes = 0;     // gathered corpus / top-k tiles
    void *ws_lens = nullptr; size_t ws_lens_bytes = 0;
    void *ws_in = nullptr; size_t ws_in_bytes = 0;             // this rank's start nodes / queries + {bad count, bad value}
    int *d_status = nullptr;                                   // one int: the status every rank agrees on
    double last_compute_ms = 0, last_gather_ms = 0;            // device time of the last sharded call: own slice | exchange
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};      // slice start | slice end | exchange start | exchange end
};

using namespace gw;

static int comm_grow(void **p, size_t *have, size_t need) {
    if (*have >= need) return GW_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *have = 0;
    if (cudaMalloc(p, need) != cudaSuccess) { cudaGetLastError(); *p = nullptr; return GW_E_TOO_LARGE; }
    *have = need;
    return GW_OK;
}

// Every rank contributes its local status (GW_OK or a negative GW_E_* code); all ranks receive the smallest one.  A rank
// that failed locally still takes part, so nobody is left waiting inside a later collective.
static int agree(gw_comm *c, int local, int *agreed) {
    *agreed = local;
    if (c->nranks == 1) return GW_OK;
    GW_CUDA(cudaMemcpyAsync(c->d_status, &local, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    GW_NCCL(g_nccl.AllReduce(c->d_status, c->d_status, 1, 2 /* ncclInt32 */, 3 /* ncclMin */, c->comm, c->stream));
    GW_CUDA(cudaMemcpyAsync(agreed, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    GW_CUDA(cudaStreamSynchronize(c->stream));
    return GW_OK;
}

// agree() + the common epilogue: the local failure wins (with its own message), else a peer's failure is reported
static int agree_or_fail(gw_comm *c, int local, const char *what) {
    const std::string msg = last_error();
    int agreed = GW_OK;
    GW_TRY(agree(c, local, &agreed));
    if (local != GW_OK) { last_error() = msg; return local; }
    if (agreed == GW_OK) return GW_OK;
    if (agreed == GW_E_KEY) return fail(GW_E_KEY, "another rank found a %s outside the graph", what);
    if (agreed == GW_E_TOO_LARGE) return fail(GW_E_TOO_LARGE, "another rank could not allocate its buffers");
    return fail(agreed, "another rank failed with status %d", agreed);
}

// block r (cnt[r] bytes at offset off[r] of d_full) travels from rank r to every rank (root < 0) or to `root` only
static int gather_blocks(gw_comm *c, void *d_full, const std::vector<size_t> &off, const std::vector<size_t> &cnt, int root) {
    if (c->nranks == 1) return GW_OK;
    GW_NCCL(g_nccl.GroupStart());
    for (int r = 0; r < c->nranks; r++) {
        if (cnt[r] == 0) continue;
        char *p = (char *)d_full + off[r];
        if (root < 0) {
            GW_NCCL(g_nccl.Broadcast(p, p, cnt[r], 0 /* ncclInt8 */, r, c->comm, c->stream));
        } else if (r != root) {
            if (c->rank == r) GW_NCCL(g_nccl.Send(p, cnt[r], 0, root, c->comm, c->stream));
            if (c->rank == root) GW_NCCL(g_nccl.Recv(p, cnt[r], 0, r, c->comm, c->stream));
        }
    }
    GW_NCCL(g_nccl.GroupEnd());
    return GW_OK;
}

__global__ void k_check_range(const int64_t *__restrict__ v, int64_t count, int64_t n, unsigned long long *bad) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int64_t x = v[i];
    if (x < 0 || x >= n) { atomicAdd(bad, 1ull); bad[1] = (unsigned long long)x; }
}

// this rank's slice [lo, hi) of a host id array onto the device, range-checked there; returns the local status
static int stage_and_check(gw_comm *c, const gw_graph *g, const int64_t *host, int64_t lo, int64_t hi, const char *what) {
    const size_t cnt = (size_t)(hi - lo);
    if (comm_grow(&c->ws_in, &c->ws_in_bytes, sizeof(int64_t) * cnt + 16) != GW_OK)
        return fail(GW_E_TOO_LARGE, "no device memory for %zu ids", cnt);
    unsigned long long *d_bad = (unsigned long long *)((int64_t *)c->ws_in + cnt), h_bad[2] = {0, 0};
    GW_CUDA(cudaMemsetAsync(d_bad, 0, 16, c->stream));
    if (cnt) {
        GW_CUDA(cudaMemcpyAsync(c->ws_in, host + lo, sizeof(int64_t) * cnt, cudaMemcpyHostToDevice, c->stream));
        k_check_range<<<(unsigned)((cnt + 255) / 256), 256, 0, c->stream>>>((const int64_t *)c->ws_in, (int64_t)cnt, g->n, d_bad);
        GW_LAUNCHED();
    }
    GW_CUDA(cudaMemcpyAsync(h_bad, d_bad, 16, cudaMemcpyDeviceToHost, c->stream));
    GW_CUDA(cudaStreamSynchronize(c->stream));
    if (h_bad[0]) return fail(GW_E_KEY, "%s %lld is outside [0, %lld)", what, (long long)h_bad[1], (long long)g->n);
    return GW_OK;
}

static void record_times(gw_comm *c) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]) == cudaSuccess) c->last_compute_ms = ms; else cudaGetLastError();
    if (cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]) == cudaSuccess) c->last_gather_ms = ms; else cudaGetLastError();
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
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc((void **)&c->d_status, sizeof(int)) != cudaSuccess) {
        if (c->stream) cudaStreamDestroy(c->stream);
        g_nccl.CommDestroy(c->comm); delete c;
        return fail(GW_E_CUDA, "communicator stream / status word could not be created");
    }
    for (int i = 0; i < 4; i++) cudaEventCreate(&c->ev[i]);
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
    cudaFree(c->ws_corpus); cudaFree(c->ws_lens); cudaFree(c->ws_in); cudaFree(c->d_status);
    for (int i = 0; i < 4; i++)
        if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    delete c;
    return GW_OK;
}

int gw_comm_last_times(const gw_comm *c, double *compute_ms, double *gather_ms) {
    if (!c) return fail(GW_E_INVALID, "communicator is NULL");
    if (compute_ms) *compute_ms = c->last_compute_ms;
    if (gather_ms) *gather_ms = c->last_gather_ms;
    return GW_OK;
}

int gw_shard_range(int64_t n, int32_t rank, int32_t nranks, int64_t *lo, int64_t *hi) {
    if (n < 0 || nranks < 1 || rank < 0 || rank >= nranks || !lo || !hi) return fail(GW_E_INVALID, "bad arguments");
    shard_range(n, rank, nranks, lo, hi);
    return GW_OK;
}

int gw_node2vec_walks_sharded(gw_graph *g, gw_comm *c, double p, double q, int32_t walk_length, const int64_t *starts,
                              int64_t n_starts, uint64_t seed, int32_t gather, int32_t *out_walks, int32_t *out_lens) {
    // arguments every rank passes alike: checked first, identically everywhere, before anything rank-dependent
    if (!g || !c) return fail(GW_E_INVALID, "graph or communicator is NULL");
    if (g->device != c->device) return fail(GW_E_INVALID, "graph lives on device %d, communicator on %d", g->device, c->device);
    if (n_starts < 0 || walk_length < 1 || (n_starts > 0 && !starts)) return fail(GW_E_INVALID, "bad arguments");
    if (gather < 0 || gather > 2) return fail(GW_E_INVALID, "gather must be 0 (keep sharded), 1 (every rank) or 2 (rank 0 only)");
    if (!(p > 0) || !(q > 0)) return fail(GW_E_INVALID, "p and q must be positive");
    if (g->flags & GW_F_MULTI) return fail(GW_E_STATE, "node2vec walks need a SIMPLE-mode (sorted) graph");
    if (n_starts == 0) return GW_OK;
    GW_CUDA(cudaSetDevice(c->device));
    const bool receives = gather == 1 || (gather == 2 && c->rank == 0);
    int64_t lo, hi;
    shard_range(n_starts, c->rank, c->nranks, &lo, &hi);
    int local = GW_OK;
    if ((gather == 0 && hi > lo && !out_walks) || (receives && !out_walks)) local = fail(GW_E_INVALID, "out_walks is NULL");
    const size_t L = (size_t)walk_length;
    c->last_compute_ms = c->last_gather_ms = 0;

    if (gather == 0) {
        // corpora beyond one GPU / host (R-MAT-26: 215 GB): every rank walks its slice through the host hand-off pipeline
        // (chunked, pinned ring, copy threads) and writes only its own rows; no data moves between ranks
        if (local == GW_OK && hi > lo)
            local = gw_node2vec_walks(g, p, q, walk_length, starts + lo, hi - lo, seed, (uint64_t)lo, out_walks + (size_t)lo * L,
                                      out_lens ? out_lens + lo : nullptr);
        return agree_or_fail(c, local, "start node");
    }

    // gathered corpus: the full [n_starts, L] block lives in device memory on every rank (the NCCL exchange is in place)
    if (local == GW_OK) local = stage_and_check(c, g, starts, lo, hi, "start node");
    if (local == GW_OK && (comm_grow(&c->ws_corpus, &c->ws_corpus_bytes, sizeof(int32_t) * (size_t)n_starts * L) != GW_OK ||
                           comm_grow(&c->ws_lens, &c->ws_lens_bytes, sizeof(int32_t) * (size_t)n_starts) != GW_OK))
        local = fail(GW_E_TOO_LARGE, "the gathered corpus of %lld walks does not fit on the device; call with gather = 0", (long long)n_starts);
    GW_TRY(agree_or_fail(c, local, "start node"));
    int32_t *d_walks = (int32_t *)c->ws_corpus, *d_lens = (int32_t *)c->ws_lens;
    cudaEventRecord(c->ev[0], c->stream);
    if (hi > lo)
        local = gw_node2vec_walks_dev(g, p, q, walk_length, (const int64_t *)c->ws_in, hi - lo, seed, (uint64_t)lo, d_walks + (size_t)lo * L,
                                      d_lens + lo, c->stream);
    cudaEventRecord(c->ev[1], c->stream);
    GW_TRY(agree_or_fail(c, local, "start node"));         // a launch failure on one rank must not strand the others
    std::vector<size_t> off(c->nranks), cnt(c->nranks), off2(c->nranks), cnt2(c->nranks);
    for (int r = 0; r < c->nranks; r++) {
        int64_t a, b;
        shard_range(n_starts, r, c->nranks, &a, &b);
        off[r] = (size_t)a * L * sizeof(int32_t); cnt[r] = (size_t)(b - a) * L * sizeof(int32_t);
        off2[r] = (size_t)a * sizeof(int32_t); cnt2[r] = (size_t)(b - a) * sizeof(int32_t);
    }
    cudaEventRecord(c->ev[2], c->stream);
    GW_TRY(gather_blocks(c, d_walks, off, cnt, gather == 2 ? 0 : -1));
    GW_TRY(gather_blocks(c, d_lens, off2, cnt2, gather == 2 ? 0 : -1));
    cudaEventRecord(c->ev[3], c->stream);
    GW_CUDA(cudaStreamSynchronize(c->stream));
    record_times(c);
    if (receives)     // device corpus -> caller's (pinned or pageable) buffer through the hand-off pipeline
        GW_TRY(corpus_to_host(g, p, q, walk_length, nullptr, d_walks, d_lens, n_starts, seed, 0, out_walks, out_lens));
    return GW_OK;
}

int gw_simrank_topk_sharded(gw_graph *g, gw_comm *c, const int64_t *queries, int64_t nq, double cdecay, int32_t step,
                            int32_t sample, int32_t k, int32_t mode, uint64_t seed, int32_t *out_ids, double *out_scores) {
    if (!g || !c) return fail(GW_E_INVALID, "graph or communicator is NULL");
    if (g->device != c->device) return fail(GW_E_INVALID, "graph lives on device %d, communicator on %d", g->device, c->device);
    if (nq < 0 || k < 1 || (nq > 0 && (!queries || !out_ids || !out_scores))) return fail(GW_E_INVALID, "bad arguments");
    // the estimator's own argument checks (step, sample, decay, k, mode, graph kind) on every rank alike, whatever its
    // slice holds: gw_simrank_check_args is what gw_simrank_topk_dev runs before it looks at the queries
    GW_TRY(gw_simrank_check_args(g, cdecay, step, sample, k, mode));
    if (nq == 0) return GW_OK;
    GW_CUDA(cudaSetDevice(c->device));
    c->last_compute_ms = c->last_gather_ms = 0;
    int64_t lo, hi;
    shard_range(nq, c->rank, c->nranks, &lo, &hi);
    int local = stage_and_check(c, g, queries, lo, hi, "query vertex");
    const size_t bs = sizeof(double) * (size_t)nq * k, bi = sizeof(int32_t) * (size_t)nq * k;
    if (local == GW_OK && comm_grow(&c->ws_corpus, &c->ws_corpus_bytes, bs + bi) != GW_OK)
        local = fail(GW_E_TOO_LARGE, "no device memory for %lld x %d result tiles", (long long)nq, k);
    GW_TRY(agree_or_fail(c, local, "query vertex"));
    double *d_sc = (double *)c->ws_corpus;
    int32_t *d_ids = (int32_t *)((char *)c->ws_corpus + bs);
    cudaEventRecord(c->ev[0], c->stream);
    if (hi > lo)
        local = gw_simrank_topk_dev(g, (const int64_t *)c->ws_in, hi - lo, cdecay, step, sample, k, mode, seed, (uint64_t)lo,
                                    d_ids + (size_t)lo * k, d_sc + (size_t)lo * k, c->stream);
    cudaEventRecord(c->ev[1], c->stream);
    if (local == GW_OK && hi > lo) {               // an accumulator overflow in this slice: truncated tiles must not be gathered as good ones
        int32_t code = 0;
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) local = fail(GW_E_CUDA, "SimRank kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
        else if (gw_simrank_last_error(g, &code) != GW_OK) local = GW_E_CUDA;
        else if (code) local = fail(GW_E_STATE, "SimRank accumulator overflow (code %d)", code);
    }
    GW_TRY(agree_or_fail(c, local, "query vertex"));
    std::vector<size_t> off(c->nranks), cnt(c->nranks), off2(c->nranks), cnt2(c->nranks);
    for (int r = 0; r < c->nranks; r++) {
        int64_t a, b;
        shard_range(nq, r, c->nranks, &a, &b);
        off[r] = (size_t)a * k * sizeof(double); cnt[r] = (size_t)(b - a) * k * sizeof(double);
        off2[r] = (size_t)a * k * sizeof(int32_t); cnt2[r] = (size_t)(b - a) * k * sizeof(int32_t);
    }
    cudaEventRecord(c->ev[2], c->stream);
    GW_TRY(gather_blocks(c, d_sc, off, cnt, -1));
    GW_TRY(gather_blocks(c, d_ids, off2, cnt2, -1));
    cudaEventRecord(c->ev[3], c->stream);
    GW_CUDA(cudaMemcpyAsync(out_scores, d_sc, bs, cudaMemcpyDeviceToHost, c->stream));
    GW_CUDA(cudaMemcpyAsync(out_ids, d_ids, bi, cudaMemcpyDeviceToHost, c->stream));
    GW_CUDA(cudaStreamSynchronize(c->stream));
    record_times(c);
    return GW_OK;
}

}  // extern "C"
