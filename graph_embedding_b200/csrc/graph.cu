// graph.cu — graph loader: edge list -> device CSR with sorted adjacency, synthetic generators.
//
// Replaces  node2vec/src/main.py:76-89 read_graph (networkx DiGraph -> to_undirected)      [SIMPLE]
//           DeepSim/TopSimAll/src/structures/Graph.java:28-57 (List<Integer>[] adjacency)  [MULTI]
//           DeepSim/TopSimAll/src/utils/graphTools/RMATGraphGenerator.java:101-150 (input shape)
// Heavy lifting (ranking ids, symmetrising, sorting, de-duplicating) is done on the device with
// CUB radix sorts; only weight conflict resolution of weighted graphs (a networkx iteration-order
// rule, small inputs) stays on the host.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <fstream>
#include <iterator>
#include <thread>
#include <numeric>
#include <unordered_map>

#include "common.cuh"

namespace gw {

std::string &last_error() {
    static thread_local std::string e;
    return e;
}
int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}
std::atomic<int64_t> g_launches{0};
static thread_local int t_device = -1;

int device_info(int *sm_count, size_t *free_bytes) {
    int dev = 0;
    GW_CUDA(cudaGetDevice(&dev));
    if (sm_count) GW_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    if (free_bytes) {
        size_t tot = 0;
        GW_CUDA(cudaMemGetInfo(free_bytes, &tot));
    }
    return GW_OK;
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_map_ids(const int64_t *__restrict__ ids, int64_t m, const int64_t *__restrict__ uniq,
                          int64_t n, int32_t *__restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    int64_t x = ids[i];
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (uniq[mid] < x) lo = mid + 1; else hi = mid;
    }
    out[i] = (int32_t)lo;
}

__global__ void k_narrow_ids(const int64_t *__restrict__ ids, int64_t m, int64_t n, int32_t *__restrict__ out,
                             int *__restrict__ bad) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    int64_t x = ids[i];
    if (x < 0 || x >= n) { *bad = 1; x = 0; }
    out[i] = (int32_t)x;
}

// SIMPLE: keys (s<<32|d) [+ (d<<32|s) when undirected]
__global__ void k_make_keys(const int32_t *__restrict__ s, const int32_t *__restrict__ d, int64_t m,
                            int undirected, uint64_t *__restrict__ keys) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint64_t a = (uint32_t)s[i], b = (uint32_t)d[i];
    keys[i] = (a << 32) | b;
    if (undirected) keys[m + i] = (b << 32) | a;
}

// MULTI: entry 2i = (a->b), 2i+1 = (b->a), in line order (Graph.addEdge, Graph.java:53-57)
__global__ void k_make_multi(const int32_t *__restrict__ s, const int32_t *__restrict__ d, int64_t m,
                             int32_t *__restrict__ ksrc, int32_t *__restrict__ vdst) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    ksrc[2 * i] = s[i]; vdst[2 * i] = d[i];
    ksrc[2 * i + 1] = d[i]; vdst[2 * i + 1] = s[i];
}

__global__ void k_split_keys(const uint64_t *__restrict__ keys, int64_t nnz, int32_t *__restrict__ col) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < nnz) col[i] = (int32_t)(uint32_t)keys[i];
}

// row_ptr[v] = first position whose source is >= v  (keys sorted by source)
__global__ void k_row_ptr_from_keys(const uint64_t *__restrict__ keys, int64_t nnz, int64_t n,
                                    int64_t *__restrict__ row_ptr) {
    int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (v > n) return;
    uint64_t target = (uint64_t)v << 32;
    int64_t lo = 0, hi = nnz;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < target) lo = mid + 1; else hi = mid;
    }
    row_ptr[v] = lo;
}
__global__ void k_row_ptr_from_src(const int32_t *__restrict__ src, int64_t nnz, int64_t n,
                                   int64_t *__restrict__ row_ptr) {
    int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (v > n) return;
    int64_t lo = 0, hi = nnz;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)src[mid] < v) lo = mid + 1; else hi = mid;
    }
    row_ptr[v] = lo;
}

__global__ void k_meta(const int64_t *__restrict__ row_ptr, int64_t n, uint2 *__restrict__ meta,
                       int32_t *__restrict__ max_deg) {
    int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    uint32_t d = 0;
    if (v < n) {
        int64_t a = row_ptr[v], b = row_ptr[v + 1];
        d = (uint32_t)(b - a);
        meta[v] = make_uint2((uint32_t)a, d);
    }
    // block max -> one atomic
    __shared__ uint32_t smax;
    if (threadIdx.x == 0) smax = 0;
    __syncthreads();
    uint32_t wmax = d;
    for (int o = 16; o; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if ((threadIdx.x & 31) == 0) atomicMax(&smax, wmax);
    __syncthreads();
    if (threadIdx.x == 0 && smax) atomicMax(max_deg, (int32_t)smax);
}

// R-MAT: per-bit quadrant descent (RMATGraphGenerator.java:119-145) with Philox instead of
// java.util.Random; tuple i -> keys[2i], keys[2i+1] (both directions); self loops -> ~0 sentinel.
__global__ void k_rmat(int scale, int64_t n_tuples, uint32_t ta, uint32_t tb, uint32_t tc, uint2 key,
                       uint64_t *__restrict__ keys) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_tuples) return;
    uint32_t from = 0, to = 0;
    for (int lvl = 0; lvl < scale; lvl += 4) {
        uint4 r = Philox::gen(make_uint4((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)lvl, 0x524D4154u), key);
        uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (lvl + j < scale) {
                uint32_t x = rr[j];
                uint32_t fb, tbit;
                if (x < ta) { fb = 0; tbit = 0; }        // top-left
                else if (x < tb) { fb = 1; tbit = 0; }   // top-right  (col high)
                else if (x < tc) { fb = 0; tbit = 1; }   // bottom-left (row high)
                else { fb = 1; tbit = 1; }
                from = (from << 1) | fb;
                to = (to << 1) | tbit;
            }
        }
    }
    if (from == to) {
        keys[2 * i] = ~0ull; keys[2 * i + 1] = ~0ull;
    } else {
        keys[2 * i] = ((uint64_t)from << 32) | to;
        keys[2 * i + 1] = ((uint64_t)to << 32) | from;
    }
}

__global__ void k_nonisolated_flags(const uint2 *__restrict__ meta, int64_t n, uint8_t *__restrict__ f) {
    int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (v < n) f[v] = meta[v].y != 0;
}

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

// ---------------------------------------------------------------------------------------------
// device pipelines
// ---------------------------------------------------------------------------------------------
// Finishes a graph whose d_row_ptr / d_col (/ d_w) are set.
static int finalize_graph(gw_graph *g) {
    if (g->nnz >= (int64_t)0xFFFFFFFFll)
        return fail(GW_E_TOO_LARGE, "graph has %lld directed entries; this build packs row offsets in 32 bits",
                    (long long)g->nnz);
    GW_CUDA(cudaMalloc((void **)&g->d_meta, sizeof(uint2) * (size_t)std::max<int64_t>(g->n, 1)));
    DevBuf<int32_t> mx;
    GW_CUDA(mx.alloc(1));
    GW_CUDA(cudaMemset(mx.p, 0, sizeof(int32_t)));
    if (g->n > 0) {
        k_meta<<<grid_for(g->n, 256), 256>>>(g->d_row_ptr, g->n, g->d_meta, mx.p);
        GW_LAUNCHED();
    }
    GW_CUDA(cudaMemcpy(&g->max_degree, mx.p, sizeof(int32_t), cudaMemcpyDeviceToHost));
    return GW_OK;
}

// sorted unique u64 keys (device, may contain a trailing ~0 sentinel run) -> CSR.  Takes
// ownership of nothing; keys buffer can be freed by the caller afterwards.
static int csr_from_sorted_keys(gw_graph *g, const uint64_t *d_keys, int64_t nkeys) {
    g->nnz = nkeys;
    GW_CUDA(cudaMalloc((void **)&g->d_col, sizeof(int32_t) * (size_t)std::max<int64_t>(nkeys, 1)));
    GW_CUDA(cudaMalloc((void **)&g->d_row_ptr, sizeof(int64_t) * (size_t)(g->n + 1)));
    if (nkeys > 0) {
        k_split_keys<<<grid_for(nkeys, 256), 256>>>(d_keys, nkeys, g->d_col);
        GW_LAUNCHED();
    }
    k_row_ptr_from_keys<<<grid_for(g->n + 1, 256), 256>>>(d_keys, nkeys, g->n, g->d_row_ptr);
    GW_LAUNCHED();
    return finalize_graph(g);
}

// sort + unique u64 keys in place-ish; returns device buffer with the unique keys
static int sort_unique_keys(DevBuf<uint64_t> &keys, int64_t count, int end_bit, DevBuf<uint64_t> &out,
                            int64_t *n_unique) {
    DevBuf<uint64_t> alt;
    GW_CUDA(alt.alloc((size_t)count));
    cub::DoubleBuffer<uint64_t> db(keys.p, alt.p);
    size_t tb = 0;
    GW_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, db, count, 0, end_bit));
    DevBuf<uint8_t> tmp;
    GW_CUDA(tmp.alloc(tb));
    GW_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tb, db, count, 0, end_bit));
    g_launches.fetch_add(1);
    uint64_t *sorted = db.Current();
    uint64_t *other = db.Alternate();
    DevBuf<int64_t> nsel;
    GW_CUDA(nsel.alloc(1));
    size_t tb2 = 0;
    GW_CUDA(cub::DeviceSelect::Unique(nullptr, tb2, sorted, other, nsel.p, count));
    if (tb2 > tb) { GW_CUDA(tmp.alloc(tb2)); }
    GW_CUDA(cub::DeviceSelect::Unique(tmp.p, tb2, sorted, other, nsel.p, count));
    g_launches.fetch_add(1);
    GW_CUDA(cudaMemcpy(n_unique, nsel.p, sizeof(int64_t), cudaMemcpyDeviceToHost));
    // hand back the buffer holding the unique keys, free the other
    if (other == keys.p) { out.p = keys.take(); alt.release(); }
    else { out.p = alt.take(); keys.release(); }
    out.n = (size_t)count;
    return GW_OK;
}

static int bits_for(int64_t n) {
    int b = 1;
    while (((int64_t)1 << b) < n) b++;
    return b;
}

// unweighted build from dense int32 endpoints on the device
static int build_unweighted_dev(gw_graph *g, const int32_t *d_s, const int32_t *d_d, int64_t m, int directed,
                                int mode) {
    if (mode == GW_MODE_MULTI) {
        int64_t cnt = 2 * m;
        DevBuf<int32_t> ks, vd, ks2, vd2;
        GW_CUDA(ks.alloc((size_t)std::max<int64_t>(cnt, 1)));
        GW_CUDA(vd.alloc((size_t)std::max<int64_t>(cnt, 1)));
        GW_CUDA(ks2.alloc((size_t)std::max<int64_t>(cnt, 1)));
        GW_CUDA(vd2.alloc((size_t)std::max<int64_t>(cnt, 1)));
        if (m > 0) {
            k_make_multi<<<grid_for(m, 256), 256>>>(d_s, d_d, m, ks.p, vd.p);
            GW_LAUNCHED();
            size_t tb = 0;   // radix sort is stable: line order survives inside each row
            GW_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, ks.p, ks2.p, vd.p, vd2.p, cnt, 0, bits_for(g->n)));
            DevBuf<uint8_t> tmp;
            GW_CUDA(tmp.alloc(tb));
            GW_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, ks.p, ks2.p, vd.p, vd2.p, cnt, 0, bits_for(g->n)));
            g_launches.fetch_add(1);
        }
        g->nnz = cnt;
        GW_CUDA(cudaMalloc((void **)&g->d_row_ptr, sizeof(int64_t) * (size_t)(g->n + 1)));
        k_row_ptr_from_src<<<grid_for(g->n + 1, 256), 256>>>(ks2.p, cnt, g->n, g->d_row_ptr);
        GW_LAUNCHED();
        g->d_col = vd2.take();
        return finalize_graph(g);
    }
    int64_t cnt = directed ? m : 2 * m;
    DevBuf<uint64_t> keys, uniq;
    GW_CUDA(keys.alloc((size_t)std::max<int64_t>(cnt, 1)));
    int64_t nu = 0;
    if (m > 0) {
        k_make_keys<<<grid_for(m, 256), 256>>>(d_s, d_d, m, directed ? 0 : 1, keys.p);
        GW_LAUNCHED();
        GW_TRY(sort_unique_keys(keys, cnt, 32 + bits_for(g->n), uniq, &nu));
    }
    return csr_from_sorted_keys(g, uniq.p, nu);
}

}  // namespace gw

using namespace gw;

// ---------------------------------------------------------------------------------------------
// host-side pieces
// ---------------------------------------------------------------------------------------------
namespace {

struct DEdge { int32_t s, d; int64_t first_line; double w; };

// Weighted graphs: networkx semantics resolved on the host (inputs of this kind are small).
// DiGraph: duplicate (u,v) keep the first adjacency position and the LAST weight; undirected:
// to_undirected() re-adds edges walking nodes in insertion order and successors in insertion
// order, the last directed edge met fixes the pair's weight (node2vec/src/main.py:81,86-87).
int build_weighted_host(gw_graph *g, const std::vector<int32_t> &s, const std::vector<int32_t> &d,
                        const double *w, int directed, const std::vector<int64_t> &node_order) {
    int64_t m = (int64_t)s.size();
    std::vector<int64_t> idx(m);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) {
        if (s[a] != s[b]) return s[a] < s[b];
        return d[a] < d[b];
    });
    std::vector<DEdge> D;
    for (int64_t i = 0; i < m;) {
        int64_t j = i;
        while (j + 1 < m && s[idx[j + 1]] == s[idx[i]] && d[idx[j + 1]] == d[idx[i]]) j++;
        D.push_back({s[idx[i]], d[idx[i]], idx[i], w[idx[j]]});
        i = j + 1;
    }
    std::vector<DEdge> E;  // final directed entries
    if (directed) {
        E = D;
    } else {
        struct U { int32_t a, b; int64_t k0, k1; double w; };
        std::vector<U> us;
        us.reserve(D.size());
        for (auto &e : D) us.push_back({std::min(e.s, e.d), std::max(e.s, e.d), node_order[e.s], e.first_line, e.w});
        std::sort(us.begin(), us.end(), [](const U &x, const U &y) {
            if (x.a != y.a) return x.a < y.a;
            if (x.b != y.b) return x.b < y.b;
            if (x.k0 != y.k0) return x.k0 < y.k0;
            return x.k1 < y.k1;
        });
        for (size_t i = 0; i < us.size();) {
            size_t j = i;
            while (j + 1 < us.size() && us[j + 1].a == us[i].a && us[j + 1].b == us[i].b) j++;
            E.push_back({us[i].a, us[i].b, 0, us[j].w});
            if (us[i].a != us[i].b) E.push_back({us[i].b, us[i].a, 0, us[j].w});
            i = j + 1;
        }
        std::sort(E.begin(), E.end(), [](const DEdge &x, const DEdge &y) {
            if (x.s != y.s) return x.s < y.s;
            return x.d < y.d;
        });
    }
    g->nnz = (int64_t)E.size();
    std::vector<int64_t> rp(g->n + 1, 0);
    std::vector<int32_t> col(E.size());
    std::vector<double> ww(E.size());
    for (size_t i = 0; i < E.size(); i++) { rp[E[i].s + 1]++; col[i] = E[i].d; ww[i] = E[i].w; }
    for (int64_t v = 0; v < g->n; v++) rp[v + 1] += rp[v];
    GW_CUDA(cudaMalloc((void **)&g->d_row_ptr, sizeof(int64_t) * rp.size()));
    GW_CUDA(cudaMalloc((void **)&g->d_col, sizeof(int32_t) * std::max<size_t>(col.size(), 1)));
    GW_CUDA(cudaMalloc((void **)&g->d_w, sizeof(double) * std::max<size_t>(ww.size(), 1)));
    GW_CUDA(cudaMemcpy(g->d_row_ptr, rp.data(), sizeof(int64_t) * rp.size(), cudaMemcpyHostToDevice));
    if (!col.empty()) {
        GW_CUDA(cudaMemcpy(g->d_col, col.data(), sizeof(int32_t) * col.size(), cudaMemcpyHostToDevice));
        GW_CUDA(cudaMemcpy(g->d_w, ww.data(), sizeof(double) * ww.size(), cudaMemcpyHostToDevice));
    }
    return finalize_graph(g);
}

int ensure_device() {
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt == 0)
        return fail(GW_E_CUDA, "no CUDA device available (%s); libgraphwalk has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (t_device >= 0) GW_CUDA(cudaSetDevice(t_device));
    return GW_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int gw_version(void) { return 100; }
const char *gw_last_error(void) { return last_error().c_str(); }
int64_t gw_kernel_launches(void) { return g_launches.load(); }

int gw_device_count(int *count) {
    if (!count) return fail(GW_E_INVALID, "count is NULL");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; return fail(GW_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    return GW_OK;
}

int gw_set_device(int device) {
    GW_CUDA(cudaSetDevice(device));
    t_device = device;
    return GW_OK;
}

int gw_graph_free(gw_graph *g) {
    if (!g) return GW_OK;
    cudaSetDevice(g->device);
    cudaFree(g->d_meta); cudaFree(g->d_col); cudaFree(g->d_w); cudaFree(g->d_row_ptr);
    cudaFree(g->d_anJ); cudaFree(g->d_anq); cudaFree(g->d_aeoff); cudaFree(g->d_aeJ); cudaFree(g->d_aeq);
    cudaFree(g->d_simrank_scratch); cudaFree(g->d_nbr4); cudaFree(g->d_bloom); cudaFree(g->d_rowhash); cudaFree(g->d_hybrid_scratch);
    cudaFree(g->ws_starts); cudaFree(g->ws_sr_dev);
    if (g->ws_sr_pin) cudaFreeHost(g->ws_sr_pin);
    for (int i = 0; i < 2; i++) { cudaFree(g->ws_out[i]); cudaFree(g->ws_lens[i]); cudaFree(g->ws_pack[i]); if (g->ws_stream[i]) cudaStreamDestroy(g->ws_stream[i]); }
    for (int i = 0; i < gw_graph::WS_SLOTS; i++) { if (g->ws_pin[i]) cudaFreeHost(g->ws_pin[i]); if (g->ws_pin_event[i]) cudaEventDestroy(g->ws_pin_event[i]); }
    delete g->ws_pool;
    if (g->ws_event) cudaEventDestroy(g->ws_event);
    delete g;
    return GW_OK;
}

int gw_graph_from_edges(const int64_t *src, const int64_t *dst, const double *w, int64_t m, int directed,
                        int mode, int64_t n_slots, gw_graph **out) {
    if (!out) return fail(GW_E_INVALID, "out is NULL");
    *out = nullptr;
    if (m < 0 || (m > 0 && (!src || !dst))) return fail(GW_E_INVALID, "bad edge arrays");
    if (mode != GW_MODE_SIMPLE && mode != GW_MODE_MULTI) return fail(GW_E_INVALID, "unknown mode %d", mode);
    if (mode == GW_MODE_MULTI && n_slots <= 0) return fail(GW_E_INVALID, "MULTI mode needs the vertex count V");
    if (mode == GW_MODE_MULTI && (w || directed))
        return fail(GW_E_INVALID, "MULTI mode is the undirected unweighted structures/Graph.java");
    if (m >= ((int64_t)1 << 30)) return fail(GW_E_TOO_LARGE, "host edge arrays above 2^30 entries are not supported");
    GW_TRY(ensure_device());
    gw_graph *g = new gw_graph();
    struct Guard { gw_graph *g; ~Guard() { if (g) gw_graph_free(g); } } guard{g};
    GW_CUDA(cudaGetDevice(&g->device));
    g->flags = (directed ? GW_F_DIRECTED : 0) | (w ? GW_F_WEIGHTED : 0) | (mode == GW_MODE_MULTI ? GW_F_MULTI : 0);

    DevBuf<int64_t> ds, dd;
    GW_CUDA(ds.alloc((size_t)std::max<int64_t>(m, 1)));
    GW_CUDA(dd.alloc((size_t)std::max<int64_t>(m, 1)));
    if (m > 0) {
        GW_CUDA(cudaMemcpy(ds.p, src, sizeof(int64_t) * (size_t)m, cudaMemcpyHostToDevice));
        GW_CUDA(cudaMemcpy(dd.p, dst, sizeof(int64_t) * (size_t)m, cudaMemcpyHostToDevice));
    }
    DevBuf<int32_t> s32, d32;
    GW_CUDA(s32.alloc((size_t)std::max<int64_t>(m, 1)));
    GW_CUDA(d32.alloc((size_t)std::max<int64_t>(m, 1)));

    if (n_slots > 0) {   // ids are already dense indices
        if (n_slots >= ((int64_t)1 << 31)) return fail(GW_E_TOO_LARGE, "more than 2^31-1 vertices");
        g->n = n_slots;
        DevBuf<int> bad;
        GW_CUDA(bad.alloc(1));
        GW_CUDA(cudaMemset(bad.p, 0, sizeof(int)));
        if (m > 0) {
            k_narrow_ids<<<grid_for(m, 256), 256>>>(ds.p, m, g->n, s32.p, bad.p); GW_LAUNCHED();
            k_narrow_ids<<<grid_for(m, 256), 256>>>(dd.p, m, g->n, d32.p, bad.p); GW_LAUNCHED();
        }
        int hb = 0;
        GW_CUDA(cudaMemcpy(&hb, bad.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (hb) return fail(GW_E_KEY, "edge endpoint outside [0, %lld)", (long long)n_slots);
    } else {             // rank the ids that occur (ascending original id)
        DevBuf<int64_t> all, all2, uniq;
        int64_t cnt = 2 * m;
        GW_CUDA(all.alloc((size_t)std::max<int64_t>(cnt, 1)));
        GW_CUDA(all2.alloc((size_t)std::max<int64_t>(cnt, 1)));
        GW_CUDA(uniq.alloc((size_t)std::max<int64_t>(cnt, 1)));
        int64_t n = 0;
        if (m > 0) {
            GW_CUDA(cudaMemcpy(all.p, ds.p, sizeof(int64_t) * (size_t)m, cudaMemcpyDeviceToDevice));
            GW_CUDA(cudaMemcpy(all.p + m, dd.p, sizeof(int64_t) * (size_t)m, cudaMemcpyDeviceToDevice));
            size_t tb = 0, tb2 = 0;
            GW_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, all.p, all2.p, cnt));
            DevBuf<int64_t> nsel;
            GW_CUDA(nsel.alloc(1));
            GW_CUDA(cub::DeviceSelect::Unique(nullptr, tb2, all2.p, uniq.p, nsel.p, cnt));
            DevBuf<uint8_t> tmp;
            GW_CUDA(tmp.alloc(std::max(tb, tb2)));
            GW_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tb, all.p, all2.p, cnt));
            GW_CUDA(cub::DeviceSelect::Unique(tmp.p, tb2, all2.p, uniq.p, nsel.p, cnt));
            g_launches.fetch_add(2);
            GW_CUDA(cudaMemcpy(&n, nsel.p, sizeof(int64_t), cudaMemcpyDeviceToHost));
            k_map_ids<<<grid_for(m, 256), 256>>>(ds.p, m, uniq.p, n, s32.p); GW_LAUNCHED();
            k_map_ids<<<grid_for(m, 256), 256>>>(dd.p, m, uniq.p, n, d32.p); GW_LAUNCHED();
        }
        g->n = n;
        g->node_ids.resize((size_t)n);
        if (n > 0) GW_CUDA(cudaMemcpy(g->node_ids.data(), uniq.p, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost));
    }
    ds.release(); dd.release();

    // list(G.nodes()) order = first appearance in the file, u before v on each line (node2vec.py:47)
    std::vector<int32_t> hs, hd;
    std::vector<int64_t> node_order;
    if (n_slots <= 0 || w) {
        hs.resize((size_t)m); hd.resize((size_t)m);
        if (m > 0) {
            GW_CUDA(cudaMemcpy(hs.data(), s32.p, sizeof(int32_t) * (size_t)m, cudaMemcpyDeviceToHost));
            GW_CUDA(cudaMemcpy(hd.data(), d32.p, sizeof(int32_t) * (size_t)m, cudaMemcpyDeviceToHost));
        }
        node_order.assign((size_t)g->n, -1);
        g->first_seen.clear();
        g->first_seen.reserve((size_t)g->n);
        for (int64_t i = 0; i < m; i++) {
            int32_t e[2] = {hs[i], hd[i]};
            for (int t = 0; t < 2; t++)
                if (node_order[e[t]] < 0) { node_order[e[t]] = (int64_t)g->first_seen.size(); g->first_seen.push_back(e[t]); }
        }
        if (n_slots > 0) {   // slots never mentioned come last, ascending (they are isolated)
            for (int64_t v = 0; v < g->n; v++)
                if (node_order[v] < 0) { node_order[v] = (int64_t)g->first_seen.size(); g->first_seen.push_back(v); }
        }
    }

    if (w) {
        GW_TRY(build_weighted_host(g, hs, hd, w, directed, node_order));
    } else {
        GW_TRY(build_unweighted_dev(g, s32.p, d32.p, m, directed, mode));
    }
    GW_CUDA(cudaDeviceSynchronize());
    guard.g = nullptr;
    *out = g;
    return GW_OK;
}

// ---- text parsing ----------------------------------------------------------------------------
static void split_fields(const std::string &line, const std::string &delim, std::vector<std::string> &out) {
    out.clear();
    if (delim.empty()) {  // any whitespace, runs collapse (python str.split())
        size_t i = 0, n = line.size();
        while (i < n) {
            while (i < n && isspace((unsigned char)line[i])) i++;
            size_t j = i;
            while (j < n && !isspace((unsigned char)line[j])) j++;
            if (j > i) out.emplace_back(line.substr(i, j - i));
            i = j;
        }
        return;
    }
    size_t pos = 0;
    for (;;) {
        size_t q = line.find(delim, pos);
        if (q == std::string::npos) { out.emplace_back(line.substr(pos)); break; }
        out.emplace_back(line.substr(pos, q - pos));
        pos = q + delim.size();
    }
}
static std::string strip(const std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) a++;
    while (b > a && isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}
static bool parse_i64(const std::string &f, int64_t *v) {
    std::string t = strip(f);
    if (t.empty()) return false;
    errno = 0;
    char *end = nullptr;
    long long x = strtoll(t.c_str(), &end, 10);
    if (errno || *end) return false;
    *v = x;
    return true;
}
static bool parse_f64(const std::string &f, double *v) {
    std::string t = strip(f);
    if (t.empty()) return false;
    errno = 0;
    char *end = nullptr;
    double x = strtod(t.c_str(), &end);
    if (*end) return false;
    *v = x;
    return true;
}

// One line of an edge list, with the reference loaders' semantics.  Returns 0 (edge appended), 1 (line skipped)
// or -1 (error, message in *err).
static int parse_edge_line(std::string &line, const std::string &delim, int weighted, int mode, const char *path,
                           int64_t lineno, std::vector<std::string> &fld, std::vector<int64_t> &src,
                           std::vector<int64_t> &dst, std::vector<double> &w, std::string *err) {
    char buf[512];
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (mode == GW_MODE_SIMPLE) {   // networkx parse_edgelist
        size_t h = line.find('#');
        if (h != std::string::npos) line.resize(h);
        line = strip(line);
        if (line.empty()) return 1;
        split_fields(line, delim, fld);
        if (fld.size() < 2) return 1;
        int64_t u, v;
        if (!parse_i64(fld[0], &u) || !parse_i64(fld[1], &v)) {
            snprintf(buf, sizeof(buf), "%s:%lld: failed to convert nodes %s,%s to type int", path, (long long)lineno,
                     fld[0].c_str(), fld[1].c_str());
            *err = buf;
            return -1;
        }
        if (weighted) {
            double x;
            if (fld.size() != 3) {
                snprintf(buf, sizeof(buf), "%s:%lld: edge data and data_keys ('weight',) are not the same length", path,
                         (long long)lineno);
                *err = buf;
                return -1;
            }
            if (!parse_f64(fld[2], &x)) {
                snprintf(buf, sizeof(buf), "%s:%lld: failed to convert weight data %s to type float", path, (long long)lineno,
                         fld[2].c_str());
                *err = buf;
                return -1;
            }
            w.push_back(x);
        } else if (fld.size() > 2 && strip(fld[2]).size() && strip(fld[2])[0] != '{') {
            snprintf(buf, sizeof(buf), "%s:%lld: failed to convert edge data to dictionary", path, (long long)lineno);
            *err = buf;
            return -1;
        }
        src.push_back(u); dst.push_back(v);
        return 0;
    }
    // Graph(String, int): line.split(SEPARATOR), ids[0], ids[1]
    if (line.empty()) return 1;
    split_fields(line, delim.empty() ? std::string(",") : delim, fld);
    int64_t u, v;
    if (fld.size() < 2 || !parse_i64(fld[0], &u) || !parse_i64(fld[1], &v)) {
        snprintf(buf, sizeof(buf), "%s:%lld: NumberFormatException for input line \"%.200s\"", path, (long long)lineno, line.c_str());
        *err = buf;
        return -1;
    }
    src.push_back(u); dst.push_back(v);
    return 0;
}

// The file is read whole and cut at line boundaries into one chunk per host thread; every chunk is parsed with
// the same per-line routine and the per-chunk edge vectors are concatenated in file order (first-appearance
// order of the nodes and the multigraph's adjacency order depend on it).  The first error in file order wins.
int gw_graph_load_edgelist(const char *path, const char *delimiter, int weighted, int directed, int mode,
                           int64_t n_slots, gw_graph **out) {
    if (!path || !out) return fail(GW_E_INVALID, "path/out is NULL");
    std::string data;
    {
        FILE *fp = fopen(path, "rb");
        if (!fp) return fail(GW_E_IO, "cannot open %s", path);
        std::vector<char> blk(1 << 22);
        size_t got;
        if (fseek(fp, 0, SEEK_END) == 0) { long sz = ftell(fp); if (sz > 0) data.reserve((size_t)sz); }
        rewind(fp);
        while ((got = fread(blk.data(), 1, blk.size(), fp)) > 0) data.append(blk.data(), got);
        fclose(fp);
    }
    const std::string delim = delimiter ? delimiter : "";
    const size_t size = data.size();
    const bool timing = getenv("GW_TIMING") != nullptr;
    const auto t_read = std::chrono::steady_clock::now();
    int nt = (int)std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), size / (1 << 20) + 1);
    nt = std::min(nt, 64);
    std::vector<size_t> cut(nt + 1, size);
    cut[0] = 0;
    for (int t = 1; t < nt; t++) {                 // chunk t starts right after the first newline at or after size*t/nt
        size_t pos = size * (size_t)t / (size_t)nt;
        if (pos < cut[t - 1]) pos = cut[t - 1];
        const void *nl = pos < size ? memchr(data.data() + pos, '\n', size - pos) : nullptr;
        cut[t] = nl ? (size_t)((const char *)nl - data.data()) + 1 : size;
    }
    struct Chunk { std::vector<int64_t> src, dst; std::vector<double> w; int64_t lines = 0, err_line = -1; std::string err; };
    std::vector<Chunk> ch(nt);
    auto work = [&](int t) {
        Chunk &c = ch[t];
        std::vector<std::string> fld;
        std::string line;
        const char *p = data.data() + cut[t], *end = data.data() + cut[t + 1];
        while (p < end) {
            const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
            const char *le = nl ? nl : end;
            line.assign(p, (size_t)(le - p));
            c.lines++;
            if (c.err_line < 0) {
                std::string e;
                if (parse_edge_line(line, delim, weighted, mode, path, c.lines, fld, c.src, c.dst, c.w, &e) < 0) { c.err_line = c.lines; c.err = e; }
            }
            p = nl ? nl + 1 : end;
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    int64_t base = 0, total = 0;
    for (int t = 0; t < nt; t++) {
        if (ch[t].err_line >= 0) {
            // the message was formatted with the chunk-local line number: rewrite it with the file's
            std::string e = ch[t].err;
            const std::string tag = std::string(path) + ":" + std::to_string(ch[t].err_line) + ":";
            const size_t at = e.find(tag);
            if (at != std::string::npos) e.replace(at, tag.size(), std::string(path) + ":" + std::to_string(base + ch[t].err_line) + ":");
            return fail(GW_E_IO, "%s", e.c_str());
        }
        base += ch[t].lines;
        total += (int64_t)ch[t].src.size();
    }
    std::vector<int64_t> src, dst;
    std::vector<double> w;
    if (nt == 1) { src.swap(ch[0].src); dst.swap(ch[0].dst); w.swap(ch[0].w); }
    else {
        src.reserve((size_t)total); dst.reserve((size_t)total);
        if (weighted) w.reserve((size_t)total);
        for (int t = 0; t < nt; t++) {
            src.insert(src.end(), ch[t].src.begin(), ch[t].src.end());
            dst.insert(dst.end(), ch[t].dst.begin(), ch[t].dst.end());
            if (weighted) w.insert(w.end(), ch[t].w.begin(), ch[t].w.end());
            std::vector<int64_t>().swap(ch[t].src); std::vector<int64_t>().swap(ch[t].dst);
        }
    }
    const auto t_parse = std::chrono::steady_clock::now();
    int rc = gw_graph_from_edges(src.data(), dst.data(), weighted ? w.data() : nullptr, (int64_t)src.size(), directed,
                                 mode, n_slots, out);
    if (timing) {
        const auto t_end = std::chrono::steady_clock::now();
        fprintf(stderr, "gw_graph_load_edgelist: %d threads, parse+concat %.1f ms, device build %.1f ms, %zu edges\n", nt,
                std::chrono::duration<double, std::milli>(t_parse - t_read).count(),
                std::chrono::duration<double, std::milli>(t_end - t_parse).count(), src.size());
    }
    return rc;
}

// ---- generators --------------------------------------------------------------------------------
int gw_graph_rmat(int scale, int64_t n_tuples, double a, double b, double c, uint64_t seed, gw_graph **out) {
    if (!out) return fail(GW_E_INVALID, "out is NULL");
    *out = nullptr;
    if (scale < 1 || scale > 30 || n_tuples < 0) return fail(GW_E_INVALID, "bad R-MAT shape");
    if (a <= 0 || b < 0 || c < 0 || a + b + c >= 1.0) return fail(GW_E_INVALID, "bad R-MAT probabilities");
    GW_TRY(ensure_device());
    gw_graph *g = new gw_graph();
    struct Guard { gw_graph *g; ~Guard() { if (g) gw_graph_free(g); } } guard{g};
    GW_CUDA(cudaGetDevice(&g->device));
    g->n = (int64_t)1 << scale;
    int64_t cnt = 2 * n_tuples;
    DevBuf<uint64_t> keys, uniq;
    GW_CUDA(keys.alloc((size_t)std::max<int64_t>(cnt, 1)));
    int64_t nu = 0;
    if (n_tuples > 0) {
        double s = 4294967296.0;
        uint32_t ta = (uint32_t)std::min(a * s, 4294967295.0), tb = (uint32_t)std::min((a + b) * s, 4294967295.0),
                 tc = (uint32_t)std::min((a + b + c) * s, 4294967295.0);
        k_rmat<<<grid_for(n_tuples, 256), 256>>>(scale, n_tuples, ta, tb, tc,
                                                   make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)), keys.p);
        GW_LAUNCHED();
        GW_TRY(sort_unique_keys(keys, cnt, 64, uniq, &nu));
        // self loops were written as ~0: after the sort they are the last unique key
        uint64_t lastk = 0;
        if (nu > 0) GW_CUDA(cudaMemcpy(&lastk, uniq.p + (nu - 1), sizeof(uint64_t), cudaMemcpyDeviceToHost));
        if (nu > 0 && lastk == ~0ull) nu--;
    }
    GW_TRY(csr_from_sorted_keys(g, uniq.p, nu));
    GW_CUDA(cudaDeviceSynchronize());
    guard.g = nullptr;
    *out = g;
    return GW_OK;
}

int gw_graph_barabasi_albert(int64_t n, int m, uint64_t seed, gw_graph **out) {
    if (!out) return fail(GW_E_INVALID, "out is NULL");
    *out = nullptr;
    if (m < 1 || n <= m || n >= ((int64_t)1 << 31)) return fail(GW_E_INVALID, "bad Barabasi-Albert shape");
    // Batagelj-Brandes repeated-endpoint list; the preferential attachment chain is inherently
    // sequential, so the edge list is produced on the host and sorted into CSR on the device.
    int64_t n_edges = (int64_t)m * (m - 1) / 2 + (n - m) * (int64_t)m;
    std::vector<int64_t> src, dst;
    src.reserve((size_t)n_edges); dst.reserve((size_t)n_edges);
    std::vector<int32_t> rep;
    rep.reserve((size_t)(2 * n_edges));
    uint64_t st = seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
    auto next = [&]() {   // splitmix64
        uint64_t z = (st += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    };
    for (int i = 0; i < m; i++)
        for (int j = i + 1; j < m; j++) { src.push_back(i); dst.push_back(j); rep.push_back(i); rep.push_back(j); }
    if (m == 1) rep.push_back(0);
    std::vector<int32_t> tg((size_t)m);
    for (int64_t v = m; v < n; v++) {
        int got = 0;
        while (got < m) {
            int32_t t = rep[(size_t)(((unsigned __int128)next() * rep.size()) >> 64)];
            bool dup = false;
            for (int j = 0; j < got; j++) dup |= (tg[j] == t);
            if (!dup) tg[got++] = t;
        }
        for (int j = 0; j < m; j++) { src.push_back(v); dst.push_back(tg[j]); rep.push_back((int32_t)v); rep.push_back(tg[j]); }
    }
    rep.clear(); rep.shrink_to_fit();
    return gw_graph_from_edges(src.data(), dst.data(), nullptr, (int64_t)src.size(), 0, GW_MODE_SIMPLE, n, out);
}

// ---- export ------------------------------------------------------------------------------------
int gw_graph_info(const gw_graph *g, int64_t *n_nodes, int64_t *n_entries, int32_t *flags, int32_t *max_degree,
                  int32_t *device) {
    if (!g) return fail(GW_E_INVALID, "graph is NULL");
    if (n_nodes) *n_nodes = g->n;
    if (n_entries) *n_entries = g->nnz;
    if (flags) *flags = g->flags;
    if (max_degree) *max_degree = g->max_degree;
    if (device) *device = g->device;
    return GW_OK;
}

int gw_graph_csr(const gw_graph *g, int64_t *row_ptr, int32_t *col_idx, double *weights, int64_t *node_ids,
                 int64_t *first_seen) {
    if (!g) return fail(GW_E_INVALID, "graph is NULL");
    GW_CUDA(cudaSetDevice(g->device));
    if (row_ptr) GW_CUDA(cudaMemcpy(row_ptr, g->d_row_ptr, sizeof(int64_t) * (size_t)(g->n + 1), cudaMemcpyDeviceToHost));
    if (col_idx && g->nnz) GW_CUDA(cudaMemcpy(col_idx, g->d_col, sizeof(int32_t) * (size_t)g->nnz, cudaMemcpyDeviceToHost));
    if (weights) {
        if (g->d_w) { if (g->nnz) GW_CUDA(cudaMemcpy(weights, g->d_w, sizeof(double) * (size_t)g->nnz, cudaMemcpyDeviceToHost)); }
        else for (int64_t i = 0; i < g->nnz; i++) weights[i] = 1.0;
    }
    if (node_ids) {
        if (!g->node_ids.empty()) memcpy(node_ids, g->node_ids.data(), sizeof(int64_t) * (size_t)g->n);
        else for (int64_t i = 0; i < g->n; i++) node_ids[i] = i;
    }
    if (first_seen) {
        if (!g->first_seen.empty()) memcpy(first_seen, g->first_seen.data(), sizeof(int64_t) * (size_t)g->n);
        else for (int64_t i = 0; i < g->n; i++) first_seen[i] = i;
    }
    return GW_OK;
}

int gw_graph_device_views(const gw_graph *g, const void **meta, const int32_t **col) {
    if (!g) return fail(GW_E_INVALID, "graph is NULL");
    if (meta) *meta = g->d_meta;
    if (col) *col = g->d_col;
    return GW_OK;
}

int gw_graph_nonisolated(const gw_graph *g, int64_t *out, int64_t *count) {
    if (!g || !count) return fail(GW_E_INVALID, "graph/count is NULL");
    GW_CUDA(cudaSetDevice(g->device));
    DevBuf<uint8_t> flags;
    DevBuf<int64_t> sel, nsel;
    GW_CUDA(flags.alloc((size_t)std::max<int64_t>(g->n, 1)));
    GW_CUDA(sel.alloc((size_t)std::max<int64_t>(g->n, 1)));
    GW_CUDA(nsel.alloc(1));
    *count = 0;
    if (g->n == 0) return GW_OK;
    k_nonisolated_flags<<<grid_for(g->n, 256), 256>>>(g->d_meta, g->n, flags.p);
    GW_LAUNCHED();
    thrust::counting_iterator<int64_t> it(0);
    size_t tb = 0;
    GW_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, it, flags.p, sel.p, nsel.p, g->n));
    DevBuf<uint8_t> tmp;
    GW_CUDA(tmp.alloc(tb));
    GW_CUDA(cub::DeviceSelect::Flagged(tmp.p, tb, it, flags.p, sel.p, nsel.p, g->n));
    g_launches.fetch_add(1);
    GW_CUDA(cudaMemcpy(count, nsel.p, sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (out && *count) GW_CUDA(cudaMemcpy(out, sel.p, sizeof(int64_t) * (size_t)*count, cudaMemcpyDeviceToHost));
    return GW_OK;
}

}  // extern "C"
