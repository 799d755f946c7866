// common.cuh — shared declarations of libgraphwalk (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/graphwalk.h"
#include "hostpipe.h"

struct gw_graph;
namespace gw {

// ---- error plumbing -------------------------------------------------------------------------
std::string &last_error();
int fail(int code, const char *fmt, ...);
extern std::atomic<int64_t> g_launches;

#define GW_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            return gw::fail(GW_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                           \
    } while (0)

#define GW_TRY(expr)                 \
    do {                             \
        int _r = (expr);             \
        if (_r != GW_OK) return _r;  \
    } while (0)

// counts a kernel launch and checks the launch error
#define GW_LAUNCHED()                                                                         \
    do {                                                                                      \
        gw::g_launches.fetch_add(1, std::memory_order_relaxed);                               \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess)                                                                \
            return gw::fail(GW_E_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                              \
    } while (0)

template <typename T>
struct DevBuf {  // RAII device allocation
    T *p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc((void **)&p, count * sizeof(T));
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    T *take() {
        T *r = p;
        p = nullptr;
        n = 0;
        return r;
    }
};

int device_info(int *sm_count, size_t *free_bytes);
int ensure_common_counts(gw_graph *g, cudaStream_t st, bool need_counts = true);
int count_walk_cn(gw_graph *g, double p, double q, int32_t L, const int64_t *d_starts, int64_t n_starts, uint64_t seed,
                  uint64_t walk_id_base, unsigned long long *d_stats, cudaStream_t st);
// walk.cu: moves a corpus to host memory through the hand-off pipeline (direct / pinned ring / packed ring); the corpus
// is produced chunk by chunk by the walker (d_corpus == NULL) or already sits in device memory (d_corpus != NULL).
int corpus_to_host(gw_graph *g, double p, double q, int32_t walk_length, const int64_t *d_starts, const int32_t *d_corpus,
                   const int32_t *d_corpus_lens, int64_t n_starts, uint64_t seed, uint64_t walk_id_base, int32_t *out_walks,
                   int32_t *out_lens);
int launch_walk_cn(gw_graph *g, double p, double q, int32_t L, const int64_t *d_starts, int64_t n_starts, uint64_t seed,
                   uint64_t walk_id_base, int32_t *d_out, int32_t *d_lens, cudaStream_t st);

}  // namespace gw

// ---- the graph handle -------------------------------------------------------------------------
// HBM layout:
//   meta[n]  uint2 {offset, degree}: ONE 8-byte load gives both ends of a row; the whole array
//            is 33.5 MB at R-MAT scale-22 and stays L2 resident (126 MB L2).
//   col[nnz] int32 neighbour lists; SIMPLE mode rows ascending, MULTI mode rows in file order.
//   w[nnz]   fp64 weights, only for weighted graphs.
struct gw_graph {
    int device = 0;
    int64_t n = 0;
    int64_t nnz = 0;
    int32_t flags = 0;
    int32_t max_degree = 0;
    uint2 *d_meta = nullptr;
    int32_t *d_col = nullptr;
    double *d_w = nullptr;
    int64_t *d_row_ptr = nullptr;  // int64[n+1], API export + table offsets
    std::vector<int64_t> node_ids;    // empty = identity
    std::vector<int64_t> first_seen;  // empty = identity
    // alias_nodes (lazy)
    int32_t *d_anJ = nullptr;
    double *d_anq = nullptr;
    // alias_edges (lazy) + the p,q they were built for
    int64_t *d_aeoff = nullptr;
    int32_t *d_aeJ = nullptr;
    double *d_aeq = nullptr;
    int64_t ae_total = 0;
    double ae_p = 0, ae_q = 0;
    // per-entry {neighbour, |N(u) & N(v)|} pairs for the second-order walker (lazy; undirected,
    // unweighted, loop-free graphs only)
    int4 *d_nbr4 = nullptr;        // {neighbour, |N(u) & N(v)|, offset(v), degree(v)} per directed entry
    int nbr4_has_counts = 0;
    int nbr4_packed = 0;          // nbr4[].y = count | reverse index << 16 (every degree < 65536)
    // word-blocked Bloom filter over the undirected edge set (lazy; q < 1 walks): "x not adjacent to prev" in
    // ONE random 8-byte access instead of a binary search over N(prev); positives are verified exactly
    unsigned long long *d_bloom = nullptr;
    uint64_t bloom_words = 0;
    // per-row hash sets for LONG rows of heavy-tailed graphs (lazy): row v with degree > 256 owns the 2*deg
    // slots rowhash[2*off(v) ...]; membership in a 100k-entry row costs 1-2 accesses instead of 17
    int32_t *d_rowhash = nullptr;
    int has_self_loops = -1;       // -1 unknown
    double common_build_ms = 0;
    // host-API workspace (grow-only): staging buffers and two streams for the chunked pipeline
    void *ws_starts = nullptr; size_t ws_starts_bytes = 0;
    void *ws_out[2] = {nullptr, nullptr}; size_t ws_out_bytes[2] = {0, 0};
    void *ws_lens[2] = {nullptr, nullptr}; size_t ws_lens_bytes[2] = {0, 0};
    void *ws_pack[2] = {nullptr, nullptr}; size_t ws_pack_bytes[2] = {0, 0};      // 3-byte ids of a chunk (packed hand-off)
    cudaStream_t ws_stream[2] = {nullptr, nullptr};
    cudaEvent_t ws_event = nullptr;
    // pinned staging ring of the corpus hand-off + the copy threads that drain it into the caller's (pageable) buffer
    static constexpr int WS_SLOTS = 4;
    void *ws_pin[WS_SLOTS] = {nullptr, nullptr, nullptr, nullptr}; size_t ws_pin_bytes[WS_SLOTS] = {0, 0, 0, 0};
    cudaEvent_t ws_pin_event[WS_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    gw::CopyPool *ws_pool = nullptr;
    int last_handoff = 0;            // how the last gw_node2vec_walks call moved its corpus: 1 direct, 2 ring, 3 packed ring
    // SimRank host-API workspace (grow-only): device queries / ids / scores and a pinned staging block
    void *ws_sr_dev = nullptr; size_t ws_sr_dev_bytes = 0;
    void *ws_sr_pin = nullptr; size_t ws_sr_pin_bytes = 0;
    // SimRank bookkeeping
    int64_t simrank_last_steps = 0;
    void *d_simrank_scratch = nullptr;
    size_t simrank_scratch_bytes = 0;
    void *d_hybrid_scratch = nullptr;
    size_t hybrid_scratch_bytes = 0;
    uint64_t simrank_layout = 0;
    int simrank_dirty = 0;
    int64_t simrank_last_slow = 0;   // queries of the last call that went through the hash kernel
};

// ---- Philox4x32-10 (Salmon et al., SC'11), counter-based ------------------------------------------
namespace gw {
struct Philox {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    __host__ __device__ static inline uint4 gen(uint4 ctr, uint2 key) {
#pragma unroll
        for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
            uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
            uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
#else
            uint64_t p0 = (uint64_t)M0 * ctr.x, p1 = (uint64_t)M1 * ctr.z;
            uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
            uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
            ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
            key.x += W0;
            key.y += W1;
        }
        return ctr;
    }
};
// ---- L2 residency control (sm_80+ cache-policy operands) -----------------------------------------
// meta[] (8 B per vertex, 33.5 MB at scale-22) is re-read by every step and fits the 126 MB L2;
// the adjacency arrays are touched once per random sector.  Loads of the former carry an
// evict_last policy, loads of the latter evict_first, so the stream of one-shot sectors does not
// wash the row descriptors out of L2.
#ifdef __CUDACC__
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint2 ld_u2_policy(const uint2 *ptr, uint64_t pol) {
    uint2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(ptr), "l"(pol));
    return v;
}
__device__ __forceinline__ int2 ld_i2_policy(const int2 *ptr, uint64_t pol) {
    int2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.s32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(ptr), "l"(pol));
    return v;
}
// L2::64B: a random 16-byte entry pulls 64 bytes from HBM instead of the default 128 (ncu: 61 B vs 120 B of
// dram__bytes_read per access, same access rate -- profiles/README.md)
#ifndef GW_LD_PREFETCH
#define GW_LD_PREFETCH ".L2::64B"
#endif
__device__ __forceinline__ int4 ld_i4_policy(const int4 *ptr, uint64_t pol) {
    int4 v;
    asm volatile("ld.global.nc.L2::cache_hint" GW_LD_PREFETCH ".v4.s32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr), "l"(pol));
    return v;
}
__device__ __forceinline__ int32_t ld_i32_policy(const int32_t *ptr, uint64_t pol) {
    int32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(ptr), "l"(pol));
    return v;
}
#endif
#ifdef __CUDACC__
// java.util.Random on the device (replay kernels): next(bits) and nextInt(bound) with its rejection loop
__device__ __forceinline__ int32_t jr_next(uint64_t &seed, int bits) {
    seed = (seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
    return (int32_t)((int64_t)seed >> (48 - bits));
}
__device__ __forceinline__ int32_t jr_next_int(uint64_t &seed, int32_t bound) {
    int32_t v = jr_next(seed, 31);
    const int32_t m = bound - 1;
    if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)v) >> 31);
    int32_t u = v;
    for (;;) {
        v = u % bound;
        if ((int32_t)((uint32_t)u - (uint32_t)v + (uint32_t)m) >= 0) break;      // u - r + m < 0 in Java int arithmetic
        u = jr_next(seed, 31);
    }
    return v;
}
#endif
// uniform index in [0,d) from 32 random bits (bias <= d / 2^32)
__host__ __device__ static inline uint32_t scale_u32(uint32_t r, uint32_t d) {
#ifdef __CUDA_ARCH__
    return __umulhi(r, d);
#else
    return (uint32_t)(((uint64_t)r * d) >> 32);
#endif
}
}  // namespace gw
