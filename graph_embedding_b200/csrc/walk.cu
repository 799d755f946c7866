// walk.cu — node2vec second-order biased walks on the device.
//
// Replaces node2vec/src/node2vec.py:13-39 node2vec_walk, :41-59 simulate_walks and :150-160
// alias_draw.  Two walkers:
//
//  * replay  — consumes the reference's own uniform stream and materialised alias tables and
//              reproduces its walks bit for bit (fp64 floor(U1*K), fp64 U2 < q[kk]).
//  * free    — the production walker.  alias_edges (sum deg^2 entries) cannot exist at scale, so
//              the second-order law of get_alias_edge (:61-81) is sampled ON THE FLY by
//              rejection: propose x from the static first-order law of N(cur) (uniform for
//              unweighted graphs, alias_nodes otherwise), accept with w(x)/ub where
//              w = 1/p (x == prev), 1 (x adjacent to prev), 1/q (otherwise).  For undirected
//              unweighted graphs the return edge's excess mass (1/p - ub) is folded into one
//              extra proposal slot, and proposals whose acceptance variate is below
//              min(1, 1/q) are accepted without touching N(prev) at all.  Adjacency of x and
//              prev is a binary search in the sorted row of prev (whose {offset,degree} stay in
//              registers) — or of x when the graph is directed.  RNG = Philox4x32-10 keyed by
//              (seed, global walk id), counter (step, attempt block): the corpus is independent
//              of launch geometry, batch split and GPU count.
//
// HBM traffic per accepted step: one 8-byte meta load (L2 resident), one 32-byte
// sector of col[] per proposal, S(d_prev) sectors per membership search, 4 bytes stored.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace gw {

// ---------------------------------------------------------------------------------------------
// replay walker
// ---------------------------------------------------------------------------------------------
__global__ void k_walk_replay(const uint2 *__restrict__ meta, const int32_t *__restrict__ col,
                              const int32_t *__restrict__ anJ, const double *__restrict__ anq,
                              const int64_t *__restrict__ aeoff, const int32_t *__restrict__ aeJ,
                              const double *__restrict__ aeq, int32_t L, const int64_t *__restrict__ starts,
                              int64_t n_walks, const double *__restrict__ uni, int64_t n_uni,
                              const int64_t *__restrict__ draw_off, int32_t *__restrict__ out,
                              int32_t *__restrict__ lens, int *__restrict__ err) {
    int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (w >= n_walks) return;
    int64_t pos = draw_off ? draw_off[w] : (int64_t)2 * (L - 1) * w;
    int32_t *o = out + w * L;
    int32_t cur = (int32_t)starts[w];
    o[0] = cur;
    int64_t e_prev = -1;
    int32_t len = 1;
    for (; len < L; len++) {
        uint2 m = meta[cur];
        int32_t K = (int32_t)m.y;
        if (K == 0) break;                                      // node2vec.py:36-37
        if (pos + 2 > n_uni) { atomicExch(err, 1); break; }
        double u1 = uni[pos], u2 = uni[pos + 1];
        pos += 2;
        const int32_t *J;
        const double *q;
        if (len == 1) { J = anJ + m.x; q = anq + m.x; }         // :28-29
        else { J = aeJ + aeoff[e_prev]; q = aeq + aeoff[e_prev]; }   // :32-34
        int32_t kk = (int32_t)floor(__dmul_rn(u1, (double)K));  // :156
        int32_t k = (u2 < q[kk]) ? kk : J[kk];                  // :157-160
        e_prev = (int64_t)m.x + k;
        cur = col[e_prev];
        o[len] = cur;
    }
    if (lens) lens[w] = len;
    for (int32_t i = len; i < L; i++) o[i] = -1;
}

// ---------------------------------------------------------------------------------------------
// free-running walker
// ---------------------------------------------------------------------------------------------
struct WalkParams {
    const uint2 *meta;
    const int32_t *col;
    const double *w;        // weighted only
    const int32_t *anJ;     // weighted only
    const double *anq;      // weighted only
    const int64_t *starts;
    int64_t n_walks;
    int32_t L;
    float inv_p, inv_q;     // 1/p, 1/q
    float ub;               // envelope height for non-folded slots
    float lb;               // weight every non-return candidate is guaranteed to reach
    float ret_w;            // acceptance height of prev when drawn from a regular slot
    uint32_t fold16;        // width of the folded return slot in 16.16 slot units, 0 = no folding
    int32_t first_order;    // p == q == 1
    uint2 key;
    uint64_t walk_id_base;
    int32_t *out;
    int32_t *lens;
};

__device__ __forceinline__ float u32_to_unit(uint32_t r) {   // [0,1) with 24 bits
    return (float)(r >> 8) * (1.0f / 16777216.0f);
}

// lower_bound of x in the sorted row [row, row+d); true when present
__device__ __forceinline__ bool row_contains(const int32_t *__restrict__ row, uint32_t d, int32_t x) {
    uint32_t lo = 0, hi = d;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        int32_t v = __ldg(row + mid);
        if (v < x) lo = mid + 1; else hi = mid;
    }
    return lo < d && __ldg(row + lo) == x;
}

template <bool WEIGHTED, bool DIRECTED>
__global__ void __launch_bounds__(256) k_walk_free(WalkParams P) {
    int64_t wi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (wi >= P.n_walks) return;
    const uint64_t wid = P.walk_id_base + (uint64_t)wi;
    int32_t *o = P.out + wi * P.L;
    int32_t cur = (int32_t)P.starts[wi];
    int32_t prev = -1;
    uint2 mprev = make_uint2(0, 0);
    o[0] = cur;
    int32_t len = 1;
    for (; len < P.L; len++) {
        const uint2 m = __ldg(P.meta + cur);
        const uint32_t d = m.y;
        if (d == 0) break;
        int32_t nxt = -1;
        uint32_t attempt = 0;
        if (!WEIGHTED) {
            // every Philox call yields two (slot, variate) proposals
            for (;;) {
                uint4 r = Philox::gen(make_uint4((uint32_t)wid, (uint32_t)(wid >> 32), (uint32_t)len, attempt), P.key);
                attempt++;
                uint32_t rs[2] = {r.x, r.z}, ry[2] = {r.y, r.w};
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    if (nxt >= 0) break;
                    if (prev < 0 || P.first_order) {       // first step / p=q=1: plain uniform draw
                        nxt = __ldg(P.col + m.x + scale_u32(rs[t], d));
                        break;
                    }
                    // slots [0,d) regular, [d, d+fold) = folded return mass; 16.16 fixed point
                    uint64_t s16 = __umul64hi((uint64_t)rs[t] << 32, ((uint64_t)d << 16) + P.fold16);
                    uint32_t k = (uint32_t)(s16 >> 16);
                    if (k >= d) { nxt = prev; break; }
                    int32_t x = __ldg(P.col + m.x + k);
                    float y = u32_to_unit(ry[t]) * P.ub;
                    if (x == prev) { if (y < P.ret_w) nxt = x; continue; }
                    if (y < P.lb) { nxt = x; continue; }   // accepted whatever the adjacency is
                    bool adj;
                    if (DIRECTED) {   // G.has_edge(x, prev): prev in N_out(x)   (node2vec.py:73)
                        uint2 mx = __ldg(P.meta + x);
                        adj = row_contains(P.col + mx.x, mx.y, prev);
                    } else {          // undirected: same as x in N(prev); prev's row bounds are in registers
                        adj = row_contains(P.col + mprev.x, mprev.y, x);
                    }
                    if (y < (adj ? 1.0f : P.inv_q)) nxt = x;
                }
                if (nxt >= 0) break;
            }
        } else {
            // one proposal per Philox call: alias_draw on alias_nodes[cur] (static first-order
            // weights, node2vec.py:150-160), then the p/q acceptance test
            for (;;) {
                uint4 r = Philox::gen(make_uint4((uint32_t)wid, (uint32_t)(wid >> 32), (uint32_t)len, attempt), P.key);
                attempt++;
                uint32_t kk = scale_u32(r.x, d);
                double qk = __ldg(P.anq + m.x + kk);
                uint32_t k = ((double)u32_to_unit(r.y) < qk) ? kk : (uint32_t)__ldg(P.anJ + m.x + kk);
                int32_t x = __ldg(P.col + m.x + k);
                if (prev < 0 || P.first_order) { nxt = x; break; }
                float y = u32_to_unit(r.z) * P.ub;
                if (x == prev) { if (y < P.ret_w) { nxt = x; break; } continue; }
                if (y < P.lb) { nxt = x; break; }
                bool adj;
                if (DIRECTED) {
                    uint2 mx = __ldg(P.meta + x);
                    adj = row_contains(P.col + mx.x, mx.y, prev);
                } else {
                    adj = row_contains(P.col + mprev.x, mprev.y, x);
                }
                if (y < (adj ? 1.0f : P.inv_q)) { nxt = x; break; }
            }
        }
        o[len] = nxt;
        prev = cur;
        mprev = m;
        cur = nxt;
    }
    if (P.lens) P.lens[wi] = len;
    for (int32_t i = len; i < P.L; i++) o[i] = -1;
}

// ---------------------------------------------------------------------------------------------
// byte model of SURVEY.md §8(d): steps and sum of S(d_prev) = max(1, ceil(log2(d+1)) - 2)
// ---------------------------------------------------------------------------------------------
__global__ void k_byte_model(const uint2 *__restrict__ meta, const int32_t *__restrict__ walks, int64_t n_walks,
                             int32_t L, int second_order, unsigned long long *__restrict__ acc) {
    int64_t wi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    unsigned long long steps = 0, sec = 0;
    if (wi < n_walks) {
        const int32_t *o = walks + wi * L;
        for (int32_t i = 1; i < L; i++) {
            if (o[i] < 0) break;
            steps++;
            if (second_order && i >= 2) {
                uint32_t d = meta[o[i - 2]].y;
                int lg = 32 - __clz(d);                  // ceil(log2(d+1)) for d >= 1
                sec += (unsigned long long)max(1, lg - 2);
            }
        }
    }
    for (int ofs = 16; ofs; ofs >>= 1) {
        steps += __shfl_xor_sync(0xffffffffu, steps, ofs);
        sec += __shfl_xor_sync(0xffffffffu, sec, ofs);
    }
    if ((threadIdx.x & 31) == 0 && steps) { atomicAdd(acc, steps); atomicAdd(acc + 1, sec); }
}

}  // namespace gw

using namespace gw;

// start-node validation on the device (the host loop costs 4 ms for 4 M starts, 15 % of a PCIe-bound call):
// bad[0] = number of out-of-range entries, bad[1] = one offending value
__global__ void k_check_starts(const int64_t *__restrict__ starts, int64_t n_starts, int64_t n, unsigned long long *bad) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_starts) return;
    const int64_t v = starts[i];
    if (v < 0 || v >= n) { atomicAdd(bad, 1ull); bad[1] = (unsigned long long)v; }
}

static int check_starts_host(const gw_graph *g, const int64_t *starts, int64_t n) {
    for (int64_t i = 0; i < n; i++)
        if (starts[i] < 0 || starts[i] >= g->n)
            return fail(GW_E_KEY, "start node index %lld is not a vertex of the graph", (long long)starts[i]);
    return GW_OK;
}

extern "C" {

int gw_node2vec_walks_dev(gw_graph *g, double p, double q, int32_t walk_length, const int64_t *d_starts,
                          int64_t n_starts, uint64_t seed, uint64_t walk_id_base, int32_t *d_out_walks,
                          int32_t *d_out_lens, void *stream) {
    if (!g) return fail(GW_E_INVALID, "graph is NULL");
    if (!(p > 0) || !(q > 0)) return fail(GW_E_INVALID, "p and q must be positive");
    if (walk_length < 1 || n_starts < 0) return fail(GW_E_INVALID, "bad walk_length / n_starts");
    if (g->flags & GW_F_MULTI) return fail(GW_E_STATE, "node2vec walks need a SIMPLE-mode (sorted) graph");
    if (n_starts == 0) return GW_OK;
    GW_CUDA(cudaSetDevice(g->device));
    const bool weighted = (g->flags & GW_F_WEIGHTED) != 0, directed = (g->flags & GW_F_DIRECTED) != 0;
    if (weighted && !g->d_anJ) GW_TRY(gw_alias_nodes(g, nullptr, nullptr));
    // production path: exact mixture sampling on common-neighbour counts (walk_cn.cu)
    const char *force = getenv("GW_WALKER");
    if (!weighted && !directed && !(force && !strcmp(force, "rejection"))) {
        int rc = ensure_common_counts(g, (cudaStream_t)stream, !(p == 1.0 && q == 1.0));   // first order: no counts
        if (rc == GW_OK)
            return launch_walk_cn(g, p, q, walk_length, d_starts, n_starts, seed, walk_id_base, d_out_walks, d_out_lens,
                                  (cudaStream_t)stream);
        if (rc != GW_E_STATE) return rc;      // GW_E_STATE = graph has self loops: generic walker below
    }
    WalkParams P;
    P.meta = g->d_meta; P.col = g->d_col; P.w = g->d_w; P.anJ = g->d_anJ; P.anq = g->d_anq;
    P.starts = d_starts; P.n_walks = n_starts; P.L = walk_length;
    P.inv_p = (float)(1.0 / p); P.inv_q = (float)(1.0 / q);
    P.first_order = (p == 1.0 && q == 1.0);
    float ub_nr = std::max(1.0f, P.inv_q);          // bound over non-return candidates
    if (!weighted && !directed && P.inv_p > ub_nr) {   // fold the return edge's excess into one slot
        P.ub = ub_nr; P.ret_w = ub_nr;
        P.fold16 = (uint32_t)std::min(4.0e9, std::floor((double)(P.inv_p - ub_nr) / ub_nr * 65536.0 + 0.5));
    } else {
        P.ub = std::max(ub_nr, P.inv_p); P.ret_w = P.inv_p; P.fold16 = 0;
    }
    P.lb = std::min(1.0f, P.inv_q);
    P.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    P.walk_id_base = walk_id_base;
    P.out = d_out_walks; P.lens = d_out_lens;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned grid = (unsigned)((n_starts + 255) / 256);
    if (weighted && directed) k_walk_free<true, true><<<grid, 256, 0, st>>>(P);
    else if (weighted) k_walk_free<true, false><<<grid, 256, 0, st>>>(P);
    else if (directed) k_walk_free<false, true><<<grid, 256, 0, st>>>(P);
    else k_walk_free<false, false><<<grid, 256, 0, st>>>(P);
    GW_LAUNCHED();
    return GW_OK;
}

// Host-buffer entry point: chunked pipeline, three ways to hand the corpus over.  The corpus is
// 4*L bytes per walk (1.3 GB per pass at R-MAT scale-22), so the call is PCIe-bound; chunk c+1 is walked on one stream
// while chunk c leaves on the other.
//   direct  the caller's buffer is page-locked: cudaMemcpyAsync straight into it (no host thread touches the data);
//   ring    the caller's buffer is pageable (a numpy array, a JVM heap array): chunks land in a library-owned pinned
//           ring and the copy threads move them on, so the DMA never falls back to the driver's staged pageable path;
//   packed  graphs of <= 2^24 vertices: ids cross PCIe as 3 bytes (k_pack24) and the copy threads widen them while
//           draining the ring -- 25 % fewer bytes over the link that bounds the call.
// Staging buffers, streams, ring and threads live in the graph handle (grow-only): repeated calls allocate nothing.
static int grow(void **p, size_t *have, size_t need) {
    if (*have >= need) return GW_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *have = 0;
    GW_CUDA(cudaMalloc(p, need));
    *have = need;
    return GW_OK;
}
static int grow_pinned(void **p, size_t *have, size_t need) {
    if (*have >= need) return GW_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr; *have = 0;
    GW_CUDA(cudaMallocHost(p, need));
    *have = need;
    return GW_OK;
}

// 4 ids -> 12 bytes; -1 (padding after a dead end) becomes 0xFFFFFF, the host restores it from the walk's length
__global__ void k_pack24(const int32_t *__restrict__ in, uint32_t *__restrict__ out, int64_t count) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t i = 4 * t;
    if (i >= count) return;
    uint32_t a, b = 0, c = 0, d = 0;
    if (i + 3 < count) {
        const int4 v = *reinterpret_cast<const int4 *>(in + i);
        a = (uint32_t)v.x; b = (uint32_t)v.y; c = (uint32_t)v.z; d = (uint32_t)v.w;
    } else {
        a = (uint32_t)in[i];
        if (i + 1 < count) b = (uint32_t)in[i + 1];
        if (i + 2 < count) c = (uint32_t)in[i + 2];
    }
    a &= 0xFFFFFFu; b &= 0xFFFFFFu; c &= 0xFFFFFFu; d &= 0xFFFFFFu;
    out[3 * t] = a | (b << 24);
    out[3 * t + 1] = (b >> 8) | (c << 16);
    out[3 * t + 2] = (c >> 16) | (d << 8);
}

enum { GW_HANDOFF_DIRECT = 1, GW_HANDOFF_RING = 2, GW_HANDOFF_PACKED = 3 };

static int pick_handoff(const gw_graph *g, const void *out_walks, int threads) {
    const bool can_pack = g->n <= ((int64_t)1 << 24);
    const char *e = getenv("GW_E2E");                         // experiment knob: direct | ring | packed
    if (e && !strcmp(e, "direct")) return GW_HANDOFF_DIRECT;
    if (e && !strcmp(e, "ring")) return GW_HANDOFF_RING;
    if (e && !strcmp(e, "packed")) return can_pack ? GW_HANDOFF_PACKED : GW_HANDOFF_RING;
    cudaPointerAttributes at;
    bool pinned = false;
    if (cudaPointerGetAttributes(&at, out_walks) == cudaSuccess) pinned = at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
    else cudaGetLastError();
    // measured (profiles/r2_handoff_sweep.txt, README R2-3): 3-byte ids + copy threads beat plain DMA from ~6 threads up
    // (12 threads: 14.4-15.7 G steps/s into pinned OR pageable memory against 12.9-13.5 G for direct DMA: the link carries
    // 25 % less); with few threads (N ranks sharing a box) direct DMA into page-locked memory wins and costs no core
    // ... but the packed ring costs 2.5 bytes of HOST memory traffic per corpus byte against 1 for direct DMA, and the host is
    // what N ranks share: at N = 4 packed gave 16.0 G steps/s for the whole box (profiles/r2_bench_n4.json), direct DMA 31.7 G
    // (round 1).  A pinned caller therefore gets the packed ring only when it has the box to itself.
    const char *lw = getenv("LOCAL_WORLD_SIZE");
    const bool alone = !(lw && atoi(lw) > 1);
    if (can_pack && threads >= 6 && (alone || !pinned)) return GW_HANDOFF_PACKED;
    if (pinned) return GW_HANDOFF_DIRECT;
    return can_pack ? GW_HANDOFF_PACKED : GW_HANDOFF_RING;
}

}  // extern "C"

// The hand-off pipeline.  Source of chunk i: the walker (d_corpus == NULL: walks of d_starts[lo, lo+cnt) produced into
// the handle's chunk buffers) or rows [lo, lo+cnt) of a corpus that already sits in device memory (gathered over NCCL).
int gw::corpus_to_host(gw_graph *g, double p, double q, int32_t walk_length, const int64_t *d_starts, const int32_t *d_corpus,
                       const int32_t *d_corpus_lens, int64_t n_starts, uint64_t seed, uint64_t walk_id_base, int32_t *out_walks,
                       int32_t *out_lens) {
    if (!g->ws_pool) g->ws_pool = new CopyPool(default_copy_threads());
    const int mode = pick_handoff(g, out_walks, g->ws_pool->threads());
    const bool ring = mode != GW_HANDOFF_DIRECT, packed = mode == GW_HANDOFF_PACKED;
    const bool produce = d_corpus == nullptr;
    const size_t L = (size_t)walk_length;
    const int64_t chunk_bytes = ring ? ((int64_t)32 << 20) : ((int64_t)48 << 20);
    const int64_t chunk = std::max<int64_t>(1024, std::min<int64_t>(n_starts, chunk_bytes / ((int64_t)L * 4))) & ~(int64_t)7;
    const bool have_lens = produce || d_corpus_lens != nullptr;
    const bool want_lens = have_lens && (out_lens != nullptr || packed);
    for (int i = 0; i < 2; i++)
        if (!g->ws_stream[i]) GW_CUDA(cudaStreamCreateWithFlags(&g->ws_stream[i], cudaStreamNonBlocking));
    const size_t chunk_ids = (size_t)chunk * L;
    const size_t pack_bytes = (chunk_ids + 3) / 4 * 12 + 16;
    for (int i = 0; i < 2; i++) {
        if (produce) {
            GW_TRY(grow(&g->ws_out[i], &g->ws_out_bytes[i], sizeof(int32_t) * chunk_ids));
            GW_TRY(grow(&g->ws_lens[i], &g->ws_lens_bytes[i], sizeof(int32_t) * (size_t)chunk));
        }
        if (packed) GW_TRY(grow(&g->ws_pack[i], &g->ws_pack_bytes[i], pack_bytes));
    }
    const size_t payload = packed ? pack_bytes : sizeof(int32_t) * chunk_ids;
    const size_t lens_off = (payload + 63) & ~(size_t)63;
    if (ring)
        for (int i = 0; i < gw_graph::WS_SLOTS; i++) {
            GW_TRY(grow_pinned(&g->ws_pin[i], &g->ws_pin_bytes[i], lens_off + sizeof(int32_t) * (size_t)chunk + 64));
            if (!g->ws_pin_event[i]) GW_CUDA(cudaEventCreateWithFlags(&g->ws_pin_event[i], cudaEventDisableTiming));
        }
    g->last_handoff = mode;
    const int64_t nchunks = (n_starts + chunk - 1) / chunk;
    constexpr int K = gw_graph::WS_SLOTS, LAG = K - 1;
    for (int64_t i = 0; i < nchunks + (ring ? LAG : 0); i++) {
        if (i < nchunks) {
            const int c = (int)(i & 1), r = (int)(i % K);
            const int64_t lo = i * chunk, cnt = std::min(chunk, n_starts - lo);
            cudaStream_t st = g->ws_stream[c];
            const int32_t *src = produce ? (const int32_t *)g->ws_out[c] : d_corpus + (size_t)lo * L;
            const int32_t *src_lens = produce ? (const int32_t *)g->ws_lens[c] : (d_corpus_lens ? d_corpus_lens + lo : nullptr);
            if (produce)
                GW_TRY(gw_node2vec_walks_dev(g, p, q, walk_length, d_starts + lo, cnt, seed, walk_id_base + (uint64_t)lo,
                                             (int32_t *)g->ws_out[c], want_lens ? (int32_t *)g->ws_lens[c] : nullptr, st));
            if (!ring) {
                GW_CUDA(cudaMemcpyAsync(out_walks + (size_t)lo * L, src, sizeof(int32_t) * (size_t)cnt * L, cudaMemcpyDeviceToHost, st));
                if (out_lens && have_lens)
                    GW_CUDA(cudaMemcpyAsync(out_lens + lo, src_lens, sizeof(int32_t) * (size_t)cnt, cudaMemcpyDeviceToHost, st));
            } else {
                const size_t ids = (size_t)cnt * L;
                if (packed) {
                    k_pack24<<<(unsigned)(((ids + 3) / 4 + 255) / 256), 256, 0, st>>>(src, (uint32_t *)g->ws_pack[c], (int64_t)ids);
                    GW_LAUNCHED();
                    GW_CUDA(cudaMemcpyAsync(g->ws_pin[r], g->ws_pack[c], (ids + 3) / 4 * 12, cudaMemcpyDeviceToHost, st));
                } else {
                    GW_CUDA(cudaMemcpyAsync(g->ws_pin[r], src, sizeof(int32_t) * ids, cudaMemcpyDeviceToHost, st));
                }
                if (want_lens)
                    GW_CUDA(cudaMemcpyAsync((char *)g->ws_pin[r] + lens_off, src_lens, sizeof(int32_t) * (size_t)cnt, cudaMemcpyDeviceToHost, st));
                GW_CUDA(cudaEventRecord(g->ws_pin_event[r], st));
            }
        }
        const int64_t j = i - LAG;
        if (ring && j >= 0) {                     // drain chunk j while the device works on j+1 .. j+LAG
            const int r = (int)(j % K);
            const int64_t lo = j * chunk, cnt = std::min(chunk, n_starts - lo);
            GW_CUDA(cudaEventSynchronize(g->ws_pin_event[r]));
            drain_chunk(g->ws_pool, g->ws_pin[r], packed ? 1 : 0, want_lens ? (const int32_t *)((char *)g->ws_pin[r] + lens_off) : nullptr,
                        cnt, walk_length, out_walks + (size_t)lo * L, out_lens ? out_lens + lo : nullptr);
        }
    }
    GW_CUDA(cudaStreamSynchronize(g->ws_stream[0]));
    GW_CUDA(cudaStreamSynchronize(g->ws_stream[1]));
    return GW_OK;
}

extern "C" {

int gw_node2vec_walks(gw_graph *g, double p, double q, int32_t walk_length, const int64_t *starts, int64_t n_starts,
                      uint64_t seed, uint64_t walk_id_base, int32_t *out_walks, int32_t *out_lens) {
    if (!g) return fail(GW_E_INVALID, "graph is NULL");
    if (n_starts < 0 || (n_starts > 0 && (!starts || !out_walks))) return fail(GW_E_INVALID, "bad arguments");
    if (walk_length < 1) return fail(GW_E_INVALID, "bad walk_length");
    if (!(p > 0) || !(q > 0)) return fail(GW_E_INVALID, "p and q must be positive");
    if (g->flags & GW_F_MULTI) return fail(GW_E_STATE, "node2vec walks need a SIMPLE-mode (sorted) graph");
    if (n_starts == 0) return GW_OK;
    GW_CUDA(cudaSetDevice(g->device));
    for (int i = 0; i < 2; i++)
        if (!g->ws_stream[i]) GW_CUDA(cudaStreamCreateWithFlags(&g->ws_stream[i], cudaStreamNonBlocking));
    if (!g->ws_event) GW_CUDA(cudaEventCreateWithFlags(&g->ws_event, cudaEventDisableTiming));
    GW_TRY(grow(&g->ws_starts, &g->ws_starts_bytes, sizeof(int64_t) * (size_t)n_starts + 16));   // + {bad count, bad value}
    // one-off preprocessing (common-neighbour counts) must not race with the two streams
    if (!(g->flags & (GW_F_DIRECTED | GW_F_WEIGHTED))) {
        int rc = ensure_common_counts(g, g->ws_stream[0], !(p == 1.0 && q == 1.0));
        if (rc != GW_OK && rc != GW_E_STATE) return rc;
    }
    if ((g->flags & GW_F_WEIGHTED) && !g->d_anJ) GW_TRY(gw_alias_nodes(g, nullptr, nullptr));
    GW_CUDA(cudaMemcpyAsync(g->ws_starts, starts, sizeof(int64_t) * (size_t)n_starts, cudaMemcpyHostToDevice, g->ws_stream[0]));
    {
        unsigned long long *d_bad = (unsigned long long *)((int64_t *)g->ws_starts + n_starts), h_bad[2] = {0, 0};
        GW_CUDA(cudaMemsetAsync(d_bad, 0, 16, g->ws_stream[0]));
        k_check_starts<<<(unsigned)((n_starts + 255) / 256), 256, 0, g->ws_stream[0]>>>((const int64_t *)g->ws_starts, n_starts, g->n, d_bad);
        GW_LAUNCHED();
        GW_CUDA(cudaMemcpyAsync(h_bad, d_bad, 16, cudaMemcpyDeviceToHost, g->ws_stream[0]));
        GW_CUDA(cudaStreamSynchronize(g->ws_stream[0]));
        if (h_bad[0]) return fail(GW_E_KEY, "start node index %lld is not a vertex of the graph", (long long)h_bad[1]);
    }
    GW_CUDA(cudaEventRecord(g->ws_event, g->ws_stream[0]));
    GW_CUDA(cudaStreamWaitEvent(g->ws_stream[1], g->ws_event, 0));
    return corpus_to_host(g, p, q, walk_length, (const int64_t *)g->ws_starts, nullptr, nullptr, n_starts, seed, walk_id_base,
                          out_walks, out_lens);
}

int gw_graph_last_handoff(const gw_graph *g, int32_t *mode, int32_t *copy_threads) {
    if (!g) return fail(GW_E_INVALID, "graph is NULL");
    if (mode) *mode = g->last_handoff;
    if (copy_threads) *copy_threads = g->ws_pool ? g->ws_pool->threads() : 0;
    return GW_OK;
}

int gw_node2vec_walks_replay(gw_graph *g, int32_t walk_length, const int64_t *starts, int64_t n_starts,
                             const double *uniforms, int64_t n_uniforms, const int64_t *draw_offset,
                             int32_t *out_walks, int32_t *out_lens) {
    if (!g) return fail(GW_E_INVALID, "graph is NULL");
    if (n_starts < 0 || (n_starts > 0 && (!starts || !out_walks)) || n_uniforms < 0 || (n_uniforms > 0 && !uniforms))
        return fail(GW_E_INVALID, "bad arguments");
    if (walk_length < 1) return fail(GW_E_INVALID, "bad walk_length");
    if (!g->d_anJ || !g->d_aeoff)
        return fail(GW_E_STATE, "replay needs gw_alias_nodes and gw_alias_edges (preprocess_transition_probs) first");
    GW_TRY(check_starts_host(g, starts, n_starts));
    if (n_starts == 0) return GW_OK;
    GW_CUDA(cudaSetDevice(g->device));
    DevBuf<int64_t> ds, doff;
    DevBuf<double> du;
    DevBuf<int32_t> dw, dl;
    DevBuf<int> derr;
    GW_CUDA(ds.alloc((size_t)n_starts));
    GW_CUDA(du.alloc((size_t)std::max<int64_t>(n_uniforms, 1)));
    GW_CUDA(dw.alloc((size_t)n_starts * walk_length));
    GW_CUDA(dl.alloc((size_t)n_starts));
    GW_CUDA(derr.alloc(1));
    GW_CUDA(cudaMemset(derr.p, 0, sizeof(int)));
    GW_CUDA(cudaMemcpy(ds.p, starts, sizeof(int64_t) * (size_t)n_starts, cudaMemcpyHostToDevice));
    if (n_uniforms) GW_CUDA(cudaMemcpy(du.p, uniforms, sizeof(double) * (size_t)n_uniforms, cudaMemcpyHostToDevice));
    if (draw_offset) {
        GW_CUDA(doff.alloc((size_t)n_starts + 1));
        GW_CUDA(cudaMemcpy(doff.p, draw_offset, sizeof(int64_t) * (size_t)(n_starts + 1), cudaMemcpyHostToDevice));
    }
    k_walk_replay<<<(unsigned)((n_starts + 127) / 128), 128>>>(g->d_meta, g->d_col, g->d_anJ, g->d_anq, g->d_aeoff,
                                                               g->d_aeJ, g->d_aeq, walk_length, ds.p, n_starts, du.p,
                                                               n_uniforms, doff.p, dw.p, dl.p, derr.p);
    GW_LAUNCHED();
    int herr = 0;
    GW_CUDA(cudaMemcpy(&herr, derr.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (herr) return fail(GW_E_INVALID, "uniform stream exhausted: %lld draws do not cover the walks", (long long)n_uniforms);
    GW_CUDA(cudaMemcpy(out_walks, dw.p, sizeof(int32_t) * (size_t)n_starts * walk_length, cudaMemcpyDeviceToHost));
    if (out_lens) GW_CUDA(cudaMemcpy(out_lens, dl.p, sizeof(int32_t) * (size_t)n_starts, cudaMemcpyDeviceToHost));
    return GW_OK;
}

int gw_graph_prepare_walks(gw_graph *g, double *build_ms) {
    if (!g) return fail(GW_E_INVALID, "graph is NULL");
    GW_CUDA(cudaSetDevice(g->device));
    if ((g->flags & (GW_F_DIRECTED | GW_F_WEIGHTED | GW_F_MULTI)) == 0) {
        int rc = ensure_common_counts(g, nullptr);
        if (rc != GW_OK && rc != GW_E_STATE) return rc;
    }
    if (build_ms) *build_ms = g->common_build_ms;
    return GW_OK;
}

__global__ void k_export_counts(const int4 *__restrict__ nbr4, int64_t nnz, int packed, int32_t *__restrict__ cnt, int32_t *__restrict__ ridx) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const uint32_t y = (uint32_t)nbr4[e].y;
    cnt[e] = packed ? (int32_t)(y & 0xFFFFu) : (int32_t)y;
    if (ridx) ridx[e] = packed ? (int32_t)(y >> 16) : -1;
}

int gw_graph_common_counts(gw_graph *g, int32_t *counts, int32_t *reverse_index) {
    if (!g || !counts) return fail(GW_E_INVALID, "bad arguments");
    if (g->flags & (GW_F_DIRECTED | GW_F_WEIGHTED | GW_F_MULTI))
        return fail(GW_E_STATE, "common-neighbour counts exist for undirected, unweighted SIMPLE graphs only");
    GW_CUDA(cudaSetDevice(g->device));
    int rc = ensure_common_counts(g, nullptr);
    if (rc != GW_OK) return rc == GW_E_STATE ? fail(GW_E_STATE, "graph has self loops: counts are not defined") : rc;
    if (g->nnz == 0) return GW_OK;
    DevBuf<int32_t> dc, dr;
    GW_CUDA(dc.alloc((size_t)g->nnz));
    if (reverse_index) GW_CUDA(dr.alloc((size_t)g->nnz));
    k_export_counts<<<(unsigned)((g->nnz + 255) / 256), 256>>>(g->d_nbr4, g->nnz, g->nbr4_packed, dc.p, dr.p);
    GW_LAUNCHED();
    GW_CUDA(cudaMemcpy(counts, dc.p, sizeof(int32_t) * (size_t)g->nnz, cudaMemcpyDeviceToHost));
    if (reverse_index) GW_CUDA(cudaMemcpy(reverse_index, dr.p, sizeof(int32_t) * (size_t)g->nnz, cudaMemcpyDeviceToHost));
    return GW_OK;
}

int gw_node2vec_walk_traffic_dev(gw_graph *g, double p, double q, int32_t walk_length, const int64_t *d_starts,
                                 int64_t n_starts, uint64_t seed, uint64_t walk_id_base, int64_t *stats5, void *stream) {
    if (!g || !stats5 || (n_starts > 0 && !d_starts)) return fail(GW_E_INVALID, "bad arguments");
    if (g->flags & (GW_F_DIRECTED | GW_F_WEIGHTED | GW_F_MULTI))
        return fail(GW_E_STATE, "traffic counting mode exists for the mixture walker only (undirected, unweighted)");
    GW_CUDA(cudaSetDevice(g->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ensure_common_counts(g, st, !(p == 1.0 && q == 1.0));
    if (rc != GW_OK) return rc == GW_E_STATE ? fail(GW_E_STATE, "graph has self loops: mixture walker not applicable") : rc;
    DevBuf<unsigned long long> acc;
    GW_CUDA(acc.alloc(5));
    GW_CUDA(cudaMemsetAsync(acc.p, 0, 5 * sizeof(unsigned long long), st));
    if (n_starts > 0) GW_TRY(count_walk_cn(g, p, q, walk_length, d_starts, n_starts, seed, walk_id_base, acc.p, st));
    unsigned long long h[5];
    GW_CUDA(cudaMemcpyAsync(h, acc.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    GW_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < 5; i++) stats5[i] = (int64_t)h[i];
    return GW_OK;
}

int gw_walks_byte_model_dev(const gw_graph *g, const int32_t *d_walks, int64_t n_walks, int32_t walk_length,
                            int second_order, int64_t *steps, int64_t *sum_search_sectors, void *stream) {
    if (!g || !d_walks || !steps || !sum_search_sectors) return fail(GW_E_INVALID, "bad arguments");
    GW_CUDA(cudaSetDevice(g->device));
    DevBuf<unsigned long long> acc;
    GW_CUDA(acc.alloc(2));
    cudaStream_t st = (cudaStream_t)stream;
    GW_CUDA(cudaMemsetAsync(acc.p, 0, 2 * sizeof(unsigned long long), st));
    if (n_walks > 0) {
        k_byte_model<<<(unsigned)((n_walks + 255) / 256), 256, 0, st>>>(g->d_meta, d_walks, n_walks, walk_length,
                                                                        second_order, acc.p);
        GW_LAUNCHED();
    }
    unsigned long long h[2];
    GW_CUDA(cudaMemcpyAsync(h, acc.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    GW_CUDA(cudaStreamSynchronize(st));
    *steps = (int64_t)h[0];
    *sum_search_sectors = (int64_t)h[1];
    return GW_OK;
}

}  // extern "C"
