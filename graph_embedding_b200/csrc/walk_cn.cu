// walk_cn.cu — the production second-order walker for undirected, unweighted, loop-free graphs
// (every BASELINE config): exact mixture sampling driven by per-edge common-neighbour counts.
//
// The law of get_alias_edge (node2vec/src/node2vec.py:61-81) over x in N(cur), given prev, is
//   w(x) = r = 1/p  (x == prev),   b = 1  (x adjacent to prev),   a = 1/q  (otherwise).
// With c = |N(cur) & N(prev)| known, d = deg(cur), lo = min(a,b), r0 = min(r,lo), it is the mixture
//   A: mass lo*(d-1) + r0      uniform over N(cur), prev thinned to r0/lo      -> NO adjacency test
//   R: mass r - r0             prev                                              -> no memory access
//   C: mass (b-lo)*c           uniform over N(cur) & N(prev)   (only when q > 1)
//   O: mass (a-lo)*(d-1-c)     uniform over N(cur) \ N(prev) \ {prev}   (only when q < 1)
// so the binary search over N(prev) that the rejection sampler pays for EVERY proposal is only
// needed in O, and C is resolved by a WARP-COOPERATIVE sorted-list intersection: the 32 lanes
// stream the shorter row in coalesced 128-byte lines, each lane binary-searches its element in the
// other row, a ballot/popc/fns picks the j-th match.  c(prev,cur) is symmetric and rides in the
// same 16-byte entry as the neighbour id AND the neighbour's row descriptor ({nbr, cnt, offset,
// degree} quads, `nbr4`), so reading the next vertex also reads the next step's count and the next
// row's bounds: ONE random 32-byte sector per step in component A, and no dependent meta[] load.
// The upper half of the count word holds the REVERSE index of the edge (where u sits in N(v)), so the
// walker always knows the position of prev inside the row it samples and never proposes it (RIDX).
//
// This replaces preprocess_transition_probs' sum(deg^2) alias_edges by a sum-over-edges
// intersection pass (k_common_counts), the only preprocessing that fits HBM at scale.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace gw {

__device__ __forceinline__ bool sorted_contains(const int32_t *__restrict__ row, uint32_t d, int32_t x) {
    uint32_t lo = 0, hi = d;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        int32_t v = __ldg(row + mid);
        if (v < x) lo = mid + 1; else hi = mid;
    }
    return lo < d && __ldg(row + lo) == x;
}

// ---- per-row hash sets of long rows (heavy-tailed graphs) ----------------------------------------
constexpr uint32_t ROWHASH_MIN_DEG = 256;      // rows longer than this own 2*deg open-addressing slots at rowhash[2*off]
__device__ __forceinline__ uint32_t rowhash_slot(int32_t x, uint32_t size) {
    uint32_t h = (uint32_t)x;
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    return __umulhi(h, size);
}
// x in N(row)?  Long rows with a hash set: linear probing from the hashed slot (load factor 1/2); others:
// binary search over the sorted row.
__device__ __forceinline__ bool row_contains(const int32_t *__restrict__ col, const int32_t *__restrict__ rowhash, uint2 mrow,
                                             int32_t x) {
    if (rowhash != nullptr && mrow.y > ROWHASH_MIN_DEG) {
        const uint32_t size = 2u * mrow.y;
        const int32_t *tab = rowhash + 2ull * mrow.x;
        uint32_t s = rowhash_slot(x, size);
        for (;;) {
            const int32_t v = __ldg(tab + s);
            if (v == x) return true;
            if (v < 0) return false;
            s = s + 1 == size ? 0 : s + 1;
        }
    }
    return sorted_contains(col + mrow.x, mrow.y, x);
}
__global__ void k_rowhash_build(const uint2 *__restrict__ meta, const int32_t *__restrict__ col, int64_t n,
                                int32_t *__restrict__ rowhash) {
    const int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t v = warp; v < n; v += nwarps) {
        const uint2 m = __ldg(meta + v);
        if (m.y <= ROWHASH_MIN_DEG) continue;
        const uint32_t size = 2u * m.y;
        int32_t *tab = rowhash + 2ull * m.x;
        for (uint32_t i = lane; i < size; i += 32) tab[i] = -1;
        __syncwarp();
        for (uint32_t i = lane; i < m.y; i += 32) {
            const int32_t x = __ldg(col + m.x + i);
            uint32_t s = rowhash_slot(x, size);
            while (atomicCAS(tab + s, -1, x) != -1) s = s + 1 == size ? 0 : s + 1;     // sorted SIMPLE rows hold no duplicates
        }
    }
}

// ---- edge Bloom filter: 16 bits per undirected edge, 4 bits of one 64-bit word per key ----------
__device__ __forceinline__ uint64_t edge_hash(int32_t a, int32_t b) {      // unordered pair
    uint64_t k = a < b ? ((uint64_t)(uint32_t)a << 32) | (uint32_t)b : ((uint64_t)(uint32_t)b << 32) | (uint32_t)a;
    k ^= k >> 30; k *= 0xBF58476D1CE4E5B9ull; k ^= k >> 27; k *= 0x94D049BB133111EBull; k ^= k >> 31;     // splitmix64 finaliser
    return k;
}
__device__ __forceinline__ unsigned long long bloom_mask(uint64_t h) {
    return (1ull << (h & 63)) | (1ull << ((h >> 6) & 63)) | (1ull << ((h >> 12) & 63)) | (1ull << ((h >> 18) & 63));
}
__device__ __forceinline__ uint64_t bloom_word(uint64_t h, uint64_t nwords) { return __umul64hi(h, nwords); }

__global__ void k_bloom_build(const uint2 *__restrict__ meta, const int32_t *__restrict__ col, int64_t n,
                              unsigned long long *__restrict__ bloom, uint64_t nwords) {
    // one warp per row: entries (u, v) with u < v set their 4 bits
    const int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warp; u < n; u += nwarps) {
        const uint2 m = __ldg(meta + u);
        for (uint32_t i = lane; i < m.y; i += 32) {
            const int32_t v = __ldg(col + m.x + i);
            if ((int64_t)v > u) {
                const uint64_t h = edge_hash((int32_t)u, v);
                atomicOr(bloom + bloom_word(h, nwords), bloom_mask(h));
            }
        }
    }
}

// first-order walks need no counts: nbr4[e] = {v, 0, offset(v), degree(v)}
__global__ void k_nbr4_nocount(const uint2 *__restrict__ meta, const int32_t *__restrict__ col, int64_t nnz,
                               int4 *__restrict__ nbr4) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    int32_t v = __ldg(col + e);
    uint2 mv = __ldg(meta + v);
    nbr4[e] = make_int4(v, 0, (int)mv.x, (int)mv.y);
}

// v1 (kept for A/B runs and the bit-identity test, GW_CN_BUILD=v1): one warp per directed entry e = (u -> v): nbr4[e] = {v, |N(u) & N(v)|, offset(v), degree(v)}.
// pack (every degree < 65536): .y = count | (position of u inside N(v)) << 16 -- the REVERSE index: a walker that
// moved u -> v knows where its previous vertex sits in the row it is about to sample from and can draw from
// N(v) \ {u} without rejection.
__global__ void __launch_bounds__(256) k_common_counts_v1(const uint2 *__restrict__ meta, const int32_t *__restrict__ col,
                                                        const int64_t *__restrict__ row_ptr, int64_t n, int64_t nnz,
                                                        int4 *__restrict__ nbr4, int *__restrict__ self_loops, int pack) {
    const int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e = warp; e < nnz; e += nwarps) {
        int64_t lo = 0, hi = n;                      // source row of entry e
        while (hi - lo > 1) {
            int64_t mid = (lo + hi) >> 1;
            if (row_ptr[mid] <= e) lo = mid; else hi = mid;
        }
        const int32_t u = (int32_t)lo, v = __ldg(col + e);
        if (u == v && lane == 0) atomicExch(self_loops, 1);
        if (u > v) continue;                         // |N(u) & N(v)| is symmetric: the warp of (v -> u) writes both entries
        const uint2 mv = __ldg(meta + v), mu = __ldg(meta + u);
        uint2 ms = mu, ml = mv;
        if (ms.y > ml.y) { uint2 t = ms; ms = ml; ml = t; }
        int cnt = 0;
        for (uint32_t i = lane; i < ms.y; i += 32)
            cnt += sorted_contains(col + ml.x, ml.y, __ldg(col + ms.x + i)) ? 1 : 0;
        for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) {
            uint32_t lo2 = 0, hi2 = mv.y;                       // position of u in the sorted row of v = the mirrored entry
            while (lo2 < hi2) {
                const uint32_t mid = (lo2 + hi2) >> 1;
                if (__ldg(col + mv.x + mid) < u) lo2 = mid + 1; else hi2 = mid;
            }
            const uint32_t k = (uint32_t)(e - (int64_t)mu.x);  // position of v in the row of u
            nbr4[e] = make_int4(v, (int)((uint32_t)cnt | (pack ? lo2 << 16 : 0u)), (int)mv.x, (int)mv.y);
            if (u != v && lo2 < mv.y)
                nbr4[(size_t)mv.x + lo2] = make_int4(u, (int)((uint32_t)cnt | (pack ? k << 16 : 0u)), (int)mu.x, (int)mu.y);
        }
    }
}

// ---- common-neighbour counts, v2: every undirected edge is intersected ONCE, by its BIGGER endpoint -------------
// Order the endpoints by (degree, id).  The task of vertex u stages N(u) as an open-addressing hash set in shared
// memory (2 slots per entry) and, for every neighbour v that is smaller in that order, streams the SHORT row N(v) in
// coalesced 128-byte lines against it: long rows are only ever probed (in shared memory), never streamed and never
// binary-searched.  The stream also meets u itself inside N(v) -- that is the reverse index of the edge, for free --
// and the position of v inside N(u) is the loop index, so both mirrored entries are written without a search.
// Work = sum over edges of min(deg) probes of ~2 shared-memory words, against sum of min(deg) * log2(max(deg))
// dependent L2/DRAM loads in v1.  Three task shapes:
//   small  (deg <= 64):       one warp per u, all smaller rows flattened into one lane-dense index space
//   block  (64 < deg <= cap): one CTA per (u, chunk of 1024 neighbours), one warp per smaller row
//   giant  (deg > cap):       same tasks, membership by binary search over N(u) in global memory (L2-resident hubs)
#ifndef GW_CC_SMALL
#define GW_CC_SMALL 64
#endif
constexpr uint32_t CC_SMALL = GW_CC_SMALL;   // rows up to this length: warp tasks (a multiple of 32)
constexpr uint32_t CC_CHUNK = 1024;          // neighbours per CTA task
constexpr uint32_t CC_HASH_MAX = 16384;      // shared-memory slots of a CTA task: rows up to 8192 entries

__device__ __forceinline__ uint32_t cc_slot(int32_t x, uint32_t mask) {
    uint32_t h = (uint32_t)x * 0x9E3779B1u;
    h ^= h >> 15;
    return h & mask;
}
__device__ __forceinline__ void cc_insert(int32_t *tab, uint32_t mask, int32_t x) {
    uint32_t s = cc_slot(x, mask);
    while (atomicCAS(tab + s, -1, x) != -1) s = (s + 1) & mask;       // SIMPLE rows hold no duplicates
}
__device__ __forceinline__ bool cc_contains(const int32_t *tab, uint32_t mask, int32_t x) {
    uint32_t s = cc_slot(x, mask);
    for (;;) {
        const int32_t v = tab[s];
        if (v == x) return true;
        if (v < 0) return false;
        s = (s + 1) & mask;
    }
}
__device__ __forceinline__ bool cc_smaller(uint32_t dv, int32_t v, uint32_t du, int32_t u) {   // (dv, v) < (du, u)
    return dv < du || (dv == du && v < u);
}
__device__ __forceinline__ void cc_write_pair(int4 *__restrict__ nbr4, int pack, int32_t u, uint2 mu, uint32_t k, int32_t v, uint2 mv,
                                              uint32_t pos_u_in_v, uint32_t cnt) {
    nbr4[(size_t)mu.x + k] = make_int4(v, (int)(cnt | (pack ? pos_u_in_v << 16 : 0u)), (int)mv.x, (int)mv.y);
    nbr4[(size_t)mv.x + pos_u_in_v] = make_int4(u, (int)(cnt | (pack ? k << 16 : 0u)), (int)mu.x, (int)mu.y);
}

// block / giant tasks are listed by this pass: tasks[i] = {u, first neighbour index}
__global__ void k_cc_list_tasks(const uint2 *__restrict__ meta, int64_t n, uint2 *__restrict__ tasks, unsigned int *__restrict__ ntasks,
                                unsigned int cap) {
    const int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (u >= n) return;
    const uint32_t d = __ldg(meta + u).y;
    if (d <= CC_SMALL) return;
    const uint32_t nt = (d + CC_CHUNK - 1) / CC_CHUNK;
    const unsigned int at = atomicAdd(ntasks, nt);
    for (uint32_t t = 0; t < nt && at + t < cap; t++) tasks[at + t] = make_uint2((uint32_t)u, t * CC_CHUNK);
}

// one warp per small vertex; per-warp shared memory: hash[128] | pre[65] | offv[64] | acc[64]
__global__ void __launch_bounds__(256) k_cc_small(const uint2 *__restrict__ meta, const int32_t *__restrict__ col, int64_t n,
                                                   int4 *__restrict__ nbr4, int *__restrict__ self_loops, int pack) {
    __shared__ int32_t s_hash[8][2 * CC_SMALL];
    __shared__ uint32_t s_pre[8][CC_SMALL + 1];
    __shared__ uint32_t s_off[8][CC_SMALL];
    __shared__ uint32_t s_acc[8][CC_SMALL];          // count | (position of u inside N(v)) << 16
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int32_t *hash = s_hash[wib];
    uint32_t *pre = s_pre[wib], *offv = s_off[wib], *acc = s_acc[wib];
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t uu = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; uu < n; uu += nwarps) {
        const int32_t u = (int32_t)uu;
        const uint2 mu = __ldg(meta + u);
        if (mu.y == 0 || mu.y > CC_SMALL) continue;
        uint32_t mask = 63;
        while (mask + 1 < 2 * mu.y) mask = 2 * mask + 1;             // 64 .. 2 * CC_SMALL slots
        for (uint32_t i = lane; i <= mask; i += 32) hash[i] = -1;
        __syncwarp();
        // the row of u (two entries per lane), descriptors of its neighbours, eligibility = smaller endpoint
        uint32_t run = 0;
#pragma unroll
        for (int h = 0; h < (int)(CC_SMALL / 32); h++) {
            const uint32_t k = h * 32 + lane;
            uint32_t dv = 0;
            if (k < mu.y) {
                const int32_t v = __ldg(col + mu.x + k);
                cc_insert(hash, mask, v);
                const uint2 mv = __ldg(meta + v);
                offv[k] = mv.x;
                acc[k] = 0;
                if (v == u) { atomicExch(self_loops, 1); nbr4[(size_t)mu.x + k] = make_int4(u, 0, (int)mu.x, (int)mu.y); }
                else if (cc_smaller(mv.y, v, mu.y, u)) dv = mv.y;
            }
            uint32_t inc = dv;                                        // inclusive warp scan of the eligible lengths
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            if (k < CC_SMALL) pre[k] = run + inc - dv;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) pre[CC_SMALL] = run;
        __syncwarp();
        const uint32_t total = run;
        for (uint32_t t = lane; t < total; t += 32) {
            uint32_t lo = 0, hi = CC_SMALL;                           // largest k with pre[k] <= t (empty rows repeat their prefix)
            while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (pre[mid] <= t) lo = mid; else hi = mid; }
            const uint32_t i = t - pre[lo];
            const int32_t x = __ldg(col + offv[lo] + i);
            if (x == u) atomicAdd(acc + lo, i << 16);
            else if (cc_contains(hash, mask, x)) atomicAdd(acc + lo, 1u);
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < (int)(CC_SMALL / 32); h++) {
            const uint32_t k = h * 32 + lane;
            if (k < mu.y && pre[k + 1] > pre[k]) {                    // eligible (an eligible row holds u: never empty)
                const int32_t v = __ldg(col + mu.x + k);
                const uint32_t a = acc[k];
                cc_write_pair(nbr4, pack, u, mu, k, v, make_uint2(offv[k], pre[k + 1] - pre[k]), a >> 16, a & 0xFFFFu);
            }
        }
        __syncwarp();
    }
}

// one CTA per task {u, k0}: neighbours k0 .. k0+CC_CHUNK of u, one warp per smaller row
__global__ void __launch_bounds__(256) k_cc_block(const uint2 *__restrict__ meta, const int32_t *__restrict__ col,
                                                   const uint2 *__restrict__ tasks, unsigned int ntasks, int4 *__restrict__ nbr4,
                                                   int *__restrict__ self_loops, int pack) {
    extern __shared__ int32_t s_tab[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (unsigned int ti = blockIdx.x; ti < ntasks; ti += gridDim.x) {
        const uint2 tk = tasks[ti];
        const int32_t u = (int32_t)tk.x;
        const uint2 mu = __ldg(meta + u);
        const bool in_smem = 2 * mu.y <= CC_HASH_MAX;
        uint32_t mask = 127;
        if (in_smem) {
            while (mask + 1 < 2 * mu.y) mask = 2 * mask + 1;
            for (uint32_t i = threadIdx.x; i <= mask; i += blockDim.x) s_tab[i] = -1;
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < mu.y; i += blockDim.x) cc_insert(s_tab, mask, __ldg(col + mu.x + i));
            __syncthreads();
        }
        const uint32_t kend = min(mu.y, tk.y + CC_CHUNK);
        // neighbour -> its row descriptor -> its row are three dependent loads (47 % of the stall samples, ncu source page):
        // the neighbour of iteration k + 16 and the descriptor of iteration k + 8 are requested while iteration k streams
        uint32_t k = tk.y + wib;
        int32_t v = k < kend ? __ldg(col + mu.x + k) : -1;
        int32_t v_next = k + 8 < kend ? __ldg(col + mu.x + k + 8) : -1;
        uint2 mv = v >= 0 ? __ldg(meta + v) : make_uint2(0, 0);
        for (; k < kend; k += 8) {
            const int32_t v_next2 = k + 16 < kend ? __ldg(col + mu.x + k + 16) : -1;
            const uint2 mv_next = v_next >= 0 ? __ldg(meta + v_next) : make_uint2(0, 0);
            const int32_t v_cur = v;
            const uint2 mv_cur = mv;
            v = v_next; mv = mv_next; v_next = v_next2;
            if (v_cur == u) {
                if (lane == 0) { atomicExch(self_loops, 1); nbr4[(size_t)mu.x + k] = make_int4(u, 0, (int)mu.x, (int)mu.y); }
                continue;
            }
            if (!cc_smaller(mv_cur.y, v_cur, mu.y, u)) continue;
            uint32_t cnt = 0, pos = 0;
            for (uint32_t b0 = 0; b0 < mv_cur.y; b0 += 32) {
                const uint32_t i = b0 + lane;
                if (i < mv_cur.y) {
                    const int32_t x = __ldg(col + mv_cur.x + i);
                    if (x == u) pos = i;
                    else if (in_smem ? cc_contains(s_tab, mask, x) : sorted_contains(col + mu.x, mu.y, x)) cnt++;
                }
            }
            for (int o = 16; o; o >>= 1) { cnt += __shfl_xor_sync(0xffffffffu, cnt, o); pos += __shfl_xor_sync(0xffffffffu, pos, o); }
            if (lane == 0) cc_write_pair(nbr4, pack, u, mu, k, v_cur, mv_cur, pos, cnt);
        }
        __syncthreads();                                             // the table is rebuilt by the next task
    }
}

struct CnParams {
    const uint2 *meta;
    const int32_t *col;
    const int4 *nbr4;
    const unsigned long long *bloom;   // edge Bloom filter (component O), may be NULL
    uint64_t bloom_words;
    const int32_t *rowhash;            // hash sets of long rows (HUB graphs), may be NULL
    const int64_t *starts;
    int64_t n_walks;
    int32_t L;
    float a, b, r;        // 1/q, 1, 1/p
    float lo, r0;
    uint2 key;
    uint64_t walk_id_base;
    int32_t *out;
    int32_t *lens;
    unsigned long long *stats;   // COUNT mode: [0] steps, [1] random accesses, [2] streamed row bytes, [3] intersections, [4] extra proposals
};

__device__ __forceinline__ float unit24(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }
__device__ __forceinline__ double unit32(uint32_t r) { return ((double)r + 0.5) * (1.0 / 4294967296.0); }

// Component choice of the mixture for rows longer than WIDE_MIN_DEG entries (HUB instantiations only): the same
// decision tree as the fp32 code in k_walk_cn, evaluated in fp64 with 32-bit uniforms.
#ifndef GW_WIDE_MIN_DEG
#define GW_WIDE_MIN_DEG 4096
#endif
constexpr uint32_t WIDE_MIN_DEG = GW_WIDE_MIN_DEG;
__device__ __noinline__ int pick_component_wide(const CnParams &P, uint32_t d, int32_t c, uint32_t r0bits, uint32_t r2bits, bool ridx) {
    const double a = (double)P.a, b = (double)P.b, r = (double)P.r, lo = (double)P.lo, r0 = (double)P.r0;
    const double dm1 = (double)(d - 1);
    const double MR = r - r0, MA = lo * dm1 + r0, MC = (b - lo) * (double)c, MO = (a - lo) * (dm1 - (double)c);
    const double u = unit32(r0bits) * (MR + MA + MC + MO);
    const bool haveC = MC > 0.0, haveO = MO > 0.0;
    int comp;
    if (d == 1 || u < MR) comp = 0;
    else if (u < MR + MA) comp = 1;
    else if (haveC && (u < MR + MA + MC || !haveO)) comp = 2;
    else if (haveO) comp = 3;
    else comp = 1;
    if (ridx && comp == 1 && unit32(r2bits) * MA < r0) comp = 0;
    return comp;
}

// Uniform draw from N(cur) & N(prev) by rejection: x uniform over the SHORTER row S, accepted iff it is in
// the other row T (uniform over S, conditioned on membership, is uniform over S & T).  The membership test
// is one Bloom word; only a positive pays for the binary search over T, which also yields x's position.
// Kept out of line: it is the rare path of heavy-tailed graphs and must not cost the common path registers.
// Returns x's nbr4 entry inside N(cur).
__device__ __noinline__ int4 common_by_rejection(const CnParams &P, uint2 m, uint2 mprev, bool cur_short, int32_t owner_t,
                                                 uint64_t wid, int32_t pos, uint32_t rk, unsigned long long *acc,
                                                 unsigned long long *prop, uint32_t *kout) {
    const uint32_t s_off = cur_short ? m.x : mprev.x, s_deg = cur_short ? m.y : mprev.y;
    const uint32_t t_off = cur_short ? mprev.x : m.x, t_deg = cur_short ? mprev.y : m.y;
    const uint32_t tsec = (uint32_t)max(1, (32 - __clz(t_deg)) - 2);
    uint32_t att = 0;                                        // owner_t: the vertex whose row is T (Bloom keys are edges)
    for (;;) {
        const uint32_t k = scale_u32(rk, s_deg);
        const int32_t x = __ldg(P.col + s_off + k);
        if (acc) (*acc)++;
        if (P.rowhash != nullptr && t_deg > ROWHASH_MIN_DEG) {           // exact membership in 1-2 accesses
            if (acc) (*acc)++;
            if (row_contains(P.col, P.rowhash, make_uint2(t_off, t_deg), x)) {
                uint32_t idx = k;
                if (!cur_short) {                                          // x's position inside N(cur) = T: one search per ACCEPTED draw
                    uint32_t lo2 = 0, hi2 = t_deg;
                    while (lo2 < hi2) {
                        const uint32_t mid = (lo2 + hi2) >> 1;
                        if (__ldg(P.col + t_off + mid) < x) lo2 = mid + 1; else hi2 = mid;
                    }
                    idx = lo2;
                    if (acc) (*acc) += tsec;
                }
                if (acc) (*acc)++;
                *kout = idx;
                return __ldg(P.nbr4 + m.x + idx);
            }
        } else {
            bool maybe = true;
            if (P.bloom) {
                const uint64_t h = edge_hash(owner_t, x);
                maybe = (__ldg(P.bloom + bloom_word(h, P.bloom_words)) & bloom_mask(h)) == bloom_mask(h);
                if (acc) (*acc)++;
            }
            if (maybe) {
                uint32_t lo2 = 0, hi2 = t_deg;                   // lower bound of x in T
                while (lo2 < hi2) {
                    const uint32_t mid = (lo2 + hi2) >> 1;
                    if (__ldg(P.col + t_off + mid) < x) lo2 = mid + 1; else hi2 = mid;
                }
                if (acc) (*acc) += tsec;
                if (lo2 < t_deg && __ldg(P.col + t_off + lo2) == x) {
                    if (acc) (*acc)++;
                    *kout = cur_short ? k : lo2;
                    return __ldg(P.nbr4 + m.x + (cur_short ? k : lo2));      // x's entry inside N(cur)
                }
            }
        }
        uint4 r2 = Philox::gen(make_uint4((uint32_t)wid, (uint32_t)(wid >> 32), (uint32_t)pos, ++att), P.key);
        rk = r2.x;
        if (prop) (*prop)++;
    }
}

// Whole second-order step by rejection, for contexts where BOTH rows are long (heavy-tailed graphs, q > 1):
// an exact common-neighbour draw would stream a long row, this costs a bounded number of accesses.
// Return with probability r / (r + W'), W' = a (d-1-c) + b c (exact, c is known); otherwise propose x uniform
// over N(cur), drop prev, accept with w(x) / max(a, b) where w = b for common neighbours and a otherwise.
// "u < min(a, b)" accepts without looking; else one Bloom word decides "not common" and only positives are
// verified by the search over N(prev).  Returns x's nbr4 entry, or .x == -2 for the return step.
__device__ __noinline__ int4 step_by_rejection(const CnParams &P, uint2 m, uint2 mprev, int32_t prev, int32_t c, uint64_t wid,
                                               int32_t pos, uint4 rnd, unsigned long long *acc, unsigned long long *prop,
                                               uint32_t *kout) {
    const uint32_t d = m.y;
    const double Wp = (double)P.a * ((double)(d - 1) - (double)c) + (double)P.b * (double)c;      // both rows are long: fp64 + 32-bit uniform
    if (d == 1 || unit32(rnd.x) * ((double)P.r + Wp) < (double)P.r) return make_int4(-2, 0, 0, 0);
    const float hi = fmaxf(P.a, P.b), lo = fminf(P.a, P.b);
    const uint32_t ssec = (uint32_t)max(1, (32 - __clz(mprev.y)) - 2);
    uint32_t rk = rnd.y, ra = rnd.z, att = 0;
    for (;;) {
        const uint32_t k = scale_u32(rk, d);
        const int4 e = __ldg(P.nbr4 + m.x + k);
        if (acc) (*acc)++;
        *kout = k;
        if (e.x != prev) {
            const float u = unit24(ra) * hi;
            bool take = u < lo;
            if (!take) {                                       // the class of x matters
                bool common;
                if (c == 0) {                                  // no common neighbour exists: x is in class "a"
                    common = false;
                } else if (P.rowhash) {                               // exact, 1-2 accesses (both rows are long here)
                    common = row_contains(P.col, P.rowhash, mprev, e.x);
                    if (acc) (*acc)++;
                } else {
                    const uint64_t h = edge_hash(prev, e.x);
                    common = (__ldg(P.bloom + bloom_word(h, P.bloom_words)) & bloom_mask(h)) == bloom_mask(h);
                    if (acc) (*acc)++;
                    if (common) {
                        common = sorted_contains(P.col + mprev.x, mprev.y, e.x);
                        if (acc) (*acc) += ssec;
                    }
                }
                take = u < (common ? P.b : P.a);
            }
            if (take) return e;
        }
        const uint4 r2 = Philox::gen(make_uint4((uint32_t)wid, (uint32_t)(wid >> 32), (uint32_t)pos, ++att), P.key);
        rk = r2.x; ra = r2.y;
        if (prop) (*prop)++;
    }
}

// COUNT = byte-model mode: same walks, no corpus stores, per-step algorithmic bytes summed.
// HUB = the graph has rows longer than 2048 entries: common-neighbour draws of long rows go through
// common_by_rejection (compiled out otherwise: the call costs the common path 8 % on flat-degree graphs).
// RIDX = nbr4[].y carries the reverse index (k_common_counts, pack): the walker tracks where prev sits in N(cur)
// (rprev) and where cur sits in N(prev) (kcur) and draws from N(cur) \ {prev} directly.  Instantiated only where the
// rejection of prev would otherwise loop (p thins the return edge below min(1, 1/q), or q < 1 excludes it): there a
// retry of ONE lane (probability ~ 1/deg) made its whole warp wait for a second memory latency in ~45 % of the
// warp-steps (ncu source page, R-MAT-22 p=4 q=0.5).
template <bool VEC8, bool COUNT, int MINB, bool HUB, bool RIDX>
__global__ void __launch_bounds__(256, MINB) k_walk_cn(CnParams P) {
    const int lane = threadIdx.x & 31;
    const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
    const int64_t wi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool valid = wi < P.n_walks;
    const uint64_t wid = P.walk_id_base + (uint64_t)wi;
    int32_t *o = P.out + (valid ? wi : 0) * P.L;
    int32_t cur = valid ? (int32_t)P.starts[wi] : 0;
    int32_t prev = -1;
    uint2 mprev = make_uint2(0, 0);
    uint2 m = valid ? ld_u2_policy(P.meta + cur, pol_keep) : make_uint2(0, 0);   // the only meta[] load of the walk
    int32_t c = 0;                       // |N(prev) & N(cur)|
    uint32_t rprev = 0, kcur = 0;        // RIDX: index of prev inside N(cur), index of cur inside N(prev)
    bool alive = valid;
    int32_t len = 1;
    int32_t buf[8];
#pragma unroll
    for (int i = 0; i < 8; i++) buf[i] = -1;
    buf[0] = cur;
    if (!VEC8 && !COUNT && valid) o[0] = cur;
    unsigned long long st_steps = 0, st_acc = 0, st_bytes = 0, st_isect = 0, st_prop = 0;

    // positions are produced in blocks of 8 so that the staging buffer is indexed statically
    for (int32_t base = 0; base < P.L; base += 8) {
#pragma unroll
        for (int s = 0; s < 8; s++) {
            const int32_t pos = base + s;
            if (pos == 0) continue;                   // the start node
            if (pos >= P.L) break;                    // uniform across the grid
            int32_t nxt = -1, cn = 0;
            uint32_t nk = 0, nr = 0;                     // RIDX: index of nxt inside N(cur), index of cur inside N(nxt)
            uint2 mn = make_uint2(0, 0);                 // row descriptor of nxt
            auto CNT = [](int y) { return RIDX ? (int32_t)((uint32_t)y & 0xFFFFu) : y; };
            auto RIX = [](int y) { return (uint32_t)y >> 16; };
            bool want_isect = false;
            uint32_t jsel = 0;
            if (alive) {
                const uint32_t d = m.y;
                if (d == 0) {
                    alive = false;
                } else {
                    uint4 rnd = Philox::gen(make_uint4((uint32_t)wid, (uint32_t)(wid >> 32), (uint32_t)pos, 0u), P.key);
                    if (COUNT) { st_steps++; if (prev < 0) st_acc++; }
                    if (prev < 0) {                                   // first step: alias_nodes law = uniform
                        nk = scale_u32(rnd.y, d);
                        int4 e = ld_i4_policy(P.nbr4 + m.x + nk, pol_stream);
                        nxt = e.x; cn = CNT(e.y); nr = RIX(e.y); mn = make_uint2((uint32_t)e.z, (uint32_t)e.w);
                    } else if (HUB && P.bloom != nullptr && P.b > P.a && min(d, mprev.y) > 256u) {
                        unsigned long long racc = 0, rprop = 0;
                        const int4 e = step_by_rejection(P, m, mprev, prev, c, wid, pos, rnd, COUNT ? &racc : nullptr,
                                                         COUNT ? &rprop : nullptr, &nk);
                        if (e.x == -2) { nxt = prev; cn = c; mn = mprev; nk = rprev; nr = kcur; }
                        else { nxt = e.x; cn = CNT(e.y); nr = RIX(e.y); mn = make_uint2((uint32_t)e.z, (uint32_t)e.w); }
                        if (COUNT) { st_acc += racc; st_prop += rprop; }
                    } else {
                        int comp;                                     // 0 = R, 1 = A, 2 = C, 3 = O
                        // Component choice in fp32.  Rows longer than WIDE_MIN_DEG entries (HUB instantiations only) must get
                        // the component that fp64 masses and a 32-bit uniform select: fp32 masses and 24-bit uniforms resolve
                        // a probability to ~6e-8 of the total, and the return mass of a 163 k-entry row is ~1e-4 of it (a
                        // 6e-4 relative error no test could see).  For them the SAME decision tree runs on the 32-bit draw;
                        // its fp32 evaluation differs from the fp64 one by < 2^-21 of the total mass, so it stands
                        // whenever the draw is further than 2^-20 of the total from every boundary (all but ~6e-6 of the
                        // steps) and the fp64 routine settles the rest.  Pure ALU, off the memory path.
                        const bool widep = HUB && d > WIDE_MIN_DEG;
                        const float dm1 = (float)(d - 1);
                        const float MR = P.r - P.r0;
                        const float MA = P.lo * dm1 + P.r0;
                        const float MC = (P.b - P.lo) * (float)c;
                        const float MO = (P.a - P.lo) * (dm1 - (float)c);
                        const float tot = MR + MA + MC + MO;
                        const float u = (widep ? (float)rnd.x * (1.0f / 4294967296.0f) : unit24(rnd.x)) * tot;
                        const bool haveC = MC > 0.0f, haveO = MO > 0.0f;
                        if (d == 1 || u < MR) comp = 0;               // d == 1: prev is the only neighbour
                        else if (u < MR + MA) comp = 1;
                        else if (haveC && (u < MR + MA + MC || !haveO)) comp = 2;
                        else if (haveO) comp = 3;
                        else comp = 1;
                        // RIDX: component A's share of prev (mass r0 of MA) is a return step decided up front; everything
                        // else draws from the d-1 entries that are not prev (index shifted past rprev)
                        float u2 = 0.0f;
                        if (RIDX && comp == 1) {
                            u2 = (widep ? (float)rnd.z * (1.0f / 4294967296.0f) : unit24(rnd.z)) * MA;
                            if (u2 < P.r0) comp = 0;
                        }
                        if (widep) {
                            const float eps = tot * (1.0f / 1048576.0f);
                            const bool near = fabsf(u - MR) < eps || fabsf(u - (MR + MA)) < eps || fabsf(u - (MR + MA + MC)) < eps ||
                                              (RIDX && u >= MR && u < MR + MA && fabsf(u2 - P.r0) < MA * (1.0f / 1048576.0f));
                            if (near) comp = pick_component_wide(P, d, c, rnd.x, rnd.z, RIDX);
                        }
                        if (COUNT && comp != 0) st_acc++;                  // one random {nbr,cnt,off,deg} access
                        // A and O both open with one proposal from N(cur): ONE load instruction for the lanes of either
                        // component.  Issued inside the two branches, a warp whose lanes split between A and O (q < 1:
                        // about half and half) waited for two memory latencies per step instead of one.
                        int4 e = make_int4(0, 0, 0, 0);
                        if (RIDX) { nk = scale_u32(rnd.y, d - 1); nk += (nk >= rprev) ? 1u : 0u; }
                        else nk = scale_u32(rnd.y, d);
                        if (comp == 1 || comp == 3) e = ld_i4_policy(P.nbr4 + m.x + nk, pol_stream);
                        if (comp == 0) {                              // R: return
                            nxt = prev; cn = c; mn = mprev; nk = rprev; nr = kcur;
                        } else if (comp == 1) {                       // A: uniform over N(cur), prev thinned
                            if (RIDX) {
                                nxt = e.x; cn = CNT(e.y); nr = RIX(e.y); mn = make_uint2((uint32_t)e.z, (uint32_t)e.w);
                            } else {
                                uint32_t ra = rnd.z, att = 0;
                                for (;;) {
                                    if (e.x != prev || unit24(ra) * P.lo < P.r0) { nxt = e.x; cn = e.y; mn = make_uint2((uint32_t)e.z, (uint32_t)e.w); break; }
                                    uint4 r2 = Philox::gen(make_uint4((uint32_t)wid, (uint32_t)(wid >> 32), (uint32_t)pos, ++att), P.key);
                                    ra = r2.y;
                                    e = ld_i4_policy(P.nbr4 + m.x + scale_u32(r2.x, d), pol_stream);
                                    if (COUNT) { st_acc++; st_prop++; }
                                }
                            }
                        } else if (comp == 2) {                       // C: uniform over N(cur) & N(prev)
                            const bool cur_short = d <= mprev.y;
                            const uint32_t s_deg = cur_short ? d : mprev.y, t_deg = cur_short ? mprev.y : d;
                            if (HUB && s_deg > 64u && c >= 16) {
                                // (per-lane rejection: ~2 dependent accesses per proposal, |S|/c proposals; the warp-
                                // cooperative stream: ~17 cached loads per 32 elements -> rejection wins from c ~ 13 up)
                                // long rows that share many neighbours (hub pairs of heavy-tailed graphs)
                                unsigned long long racc = 0, rprop = 0;
                                const int4 e = common_by_rejection(P, m, mprev, cur_short, cur_short ? prev : cur, wid, pos, rnd.y,
                                                                   COUNT ? &racc : nullptr, COUNT ? &rprop : nullptr, &nk);
                                nxt = e.x; cn = CNT(e.y); nr = RIX(e.y); mn = make_uint2((uint32_t)e.z, (uint32_t)e.w);
                                if (COUNT) { st_acc += racc; st_prop += rprop; }
                            } else {
                                want_isect = true;
                                jsel = scale_u32(rnd.y, (uint32_t)c);
                                if (COUNT) {   // both rows read once, in whole sectors
                                    st_bytes += 32ull * ((d * 4 + 31) / 32) + 32ull * ((mprev.y * 4 + 31) / 32);
                                    st_isect++;
                                }
                            }
                        } else {                                      // O: uniform over N(cur) \ N(prev) \ {prev}
                            uint32_t att = 0;
                            const uint32_t ssec = (uint32_t)max(1, (32 - __clz(mprev.y)) - 2);   // S(d_prev) random sectors per search
                            for (;;) {
                                bool take = RIDX || e.x != prev;     // RIDX: prev was never proposed
                                if (take && c != 0) {                // c == 0: N(cur) & N(prev) is empty, nothing to exclude, no test
                                    // "not adjacent to prev": one 8-byte Bloom word says so for ~99 % of the
                                    // non-neighbours; only positives pay for the exact search over N(prev)
                                    bool maybe = true;
                                    if (P.bloom) {
                                        const uint64_t h = edge_hash(prev, e.x);
                                        unsigned long long w;
                                        asm volatile("ld.global.nc" GW_LD_PREFETCH ".u64 %0, [%1];" : "=l"(w) : "l"(P.bloom + bloom_word(h, P.bloom_words)));
                                        const unsigned long long bm = bloom_mask(h);
                                        maybe = (w & bm) == bm;
                                        if (COUNT) st_acc++;
                                    }
                                    if (maybe) {
                                        if (COUNT) st_acc += ssec;
                                        take = !row_contains(P.col, HUB ? P.rowhash : nullptr, mprev, e.x);
                                    }
                                }
                                if (take) { nxt = e.x; cn = CNT(e.y); nr = RIX(e.y); mn = make_uint2((uint32_t)e.z, (uint32_t)e.w); break; }
                                uint4 r2 = Philox::gen(make_uint4((uint32_t)wid, (uint32_t)(wid >> 32), (uint32_t)pos, ++att), P.key);
                                if (RIDX) { nk = scale_u32(r2.x, d - 1); nk += (nk >= rprev) ? 1u : 0u; }
                                else nk = scale_u32(r2.x, d);
                                e = ld_i4_policy(P.nbr4 + m.x + nk, pol_stream);
                                if (COUNT) { st_acc++; st_prop++; }
                            }
                        }
                    }
                }
            }
            // ---- warp-cooperative intersections, one requesting lane at a time ----
            uint32_t req = __ballot_sync(0xffffffffu, want_isect);
            while (req) {
                const int src = __ffs(req) - 1;
                req &= req - 1;
                const uint32_t c_off = __shfl_sync(0xffffffffu, m.x, src), c_deg = __shfl_sync(0xffffffffu, m.y, src);
                const uint32_t p_off = __shfl_sync(0xffffffffu, mprev.x, src), p_deg = __shfl_sync(0xffffffffu, mprev.y, src);
                const uint32_t j = __shfl_sync(0xffffffffu, jsel, src);
                const bool scan_cur = c_deg <= p_deg;          // stream the shorter row
                const uint32_t s_off = scan_cur ? c_off : p_off, s_deg = scan_cur ? c_deg : p_deg;
                const uint32_t t_off = scan_cur ? p_off : c_off, t_deg = scan_cur ? p_deg : c_deg;
                // long target rows (HUB graphs): one Bloom word answers "not in T" for ~99 % of the streamed
                // elements; only positives pay for the binary search
                const bool use_bloom = HUB && P.bloom != nullptr && t_deg > 256u;
                const int32_t owner_t = __shfl_sync(0xffffffffu, scan_cur ? prev : cur, src);
                uint32_t seen = 0;
                int32_t xsel = -1;
                uint32_t isel = 0;
                for (uint32_t b0 = 0; b0 < s_deg; b0 += 32) {
                    const uint32_t i = b0 + lane;
                    int32_t x = (i < s_deg) ? __ldg(P.col + s_off + i) : -1;
                    bool f = i < s_deg;
                    if (f && use_bloom) {
                        const uint64_t h = edge_hash(owner_t, x);
                        f = (__ldg(P.bloom + bloom_word(h, P.bloom_words)) & bloom_mask(h)) == bloom_mask(h);
                    }
                    f = f && row_contains(P.col, HUB ? P.rowhash : nullptr, make_uint2(t_off, t_deg), x);
                    uint32_t bal = __ballot_sync(0xffffffffu, f);
                    uint32_t nb = __popc(bal);
                    if (seen + nb > j) {
                        int ln = __fns(bal, 0, (int)(j - seen) + 1);
                        xsel = __shfl_sync(0xffffffffu, x, ln);
                        isel = b0 + ln;
                        break;
                    }
                    seen += nb;
                }
                if (lane == src) {
                    if (xsel < 0) {                   // counts and rows disagree: cannot happen; stay exact-ish
                        nxt = prev; cn = c; mn = mprev; nk = rprev; nr = kcur;
                    } else {
                        uint32_t idx = isel;          // index of xsel inside N(cur)
                        if (!scan_cur) {
                            uint32_t lo2 = 0, hi2 = c_deg;
                            while (lo2 < hi2) {
                                uint32_t mid = (lo2 + hi2) >> 1;
                                if (__ldg(P.col + c_off + mid) < xsel) lo2 = mid + 1; else hi2 = mid;
                            }
                            idx = lo2;
                        }
                        nxt = xsel;
                        int4 e = __ldg(P.nbr4 + c_off + idx);
                        cn = CNT(e.y); nk = idx; nr = RIX(e.y); mn = make_uint2((uint32_t)e.z, (uint32_t)e.w);
                    }
                }
            }
            if (alive) {
                buf[s] = nxt;
                if (!VEC8 && !COUNT) o[pos] = nxt;
                prev = cur; mprev = m; cur = nxt; c = cn; m = mn;
                if (RIDX) { kcur = nk; rprev = nr; }
                len = pos + 1;
            }
        }
        if (VEC8 && !COUNT && valid) {      // one full 32-byte sector per 8 steps, no read-modify-write in L2/DRAM
            int4 *dst = reinterpret_cast<int4 *>(o + base);
            dst[0] = make_int4(buf[0], buf[1], buf[2], buf[3]);
            dst[1] = make_int4(buf[4], buf[5], buf[6], buf[7]);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) buf[i] = -1;
    }
    if (COUNT) {
        for (int o2 = 16; o2; o2 >>= 1) {
            st_steps += __shfl_xor_sync(0xffffffffu, st_steps, o2); st_bytes += __shfl_xor_sync(0xffffffffu, st_bytes, o2);
            st_acc += __shfl_xor_sync(0xffffffffu, st_acc, o2);
            st_isect += __shfl_xor_sync(0xffffffffu, st_isect, o2); st_prop += __shfl_xor_sync(0xffffffffu, st_prop, o2);
        }
        if (lane == 0) { atomicAdd(P.stats, st_steps); atomicAdd(P.stats + 1, st_acc); atomicAdd(P.stats + 2, st_bytes); atomicAdd(P.stats + 3, st_isect); atomicAdd(P.stats + 4, st_prop); }
        return;
    }
    if (valid) {
        if (P.lens) P.lens[wi] = len;
        if (!VEC8)
            for (int32_t i = len; i < P.L; i++) o[i] = -1;
    }
}

// Builds nbr4 once per graph.  need_counts = false (first-order walks) skips the intersection pass.
// Returns GW_E_STATE (usable by the caller as "not applicable") when counts are needed and the
// graph has self loops (c(prev,cur) is then not symmetric).
int ensure_common_counts(gw_graph *g, cudaStream_t st, bool need_counts) {
    if (g->d_nbr4 && (g->nbr4_has_counts || !need_counts)) return GW_OK;
    if (need_counts && g->has_self_loops == 1) return GW_E_STATE;
    if (!g->d_nbr4) GW_CUDA(cudaMalloc((void **)&g->d_nbr4, sizeof(int4) * (size_t)std::max<int64_t>(g->nnz, 1)));
    if (!need_counts) {
        if (g->nnz > 0) {
            k_nbr4_nocount<<<(unsigned)((g->nnz + 255) / 256), 256, 0, st>>>(g->d_meta, g->d_col, g->nnz, g->d_nbr4);
            GW_LAUNCHED();
        }
        GW_CUDA(cudaStreamSynchronize(st));
        return GW_OK;
    }
    DevBuf<int> flag;
    GW_CUDA(flag.alloc(1));
    GW_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
    // the task list lives outside the timed region on both sides: cudaMalloc / cudaFree of its ~nnz/4 bytes are host calls
    // that took 0.4-0.8 s on some boxes of the pool and were being charged to the kernels (profiles/README.md R2-1)
    const unsigned int cap = (unsigned int)std::min<int64_t>(g->nnz / 64 + g->nnz / CC_CHUNK + 16, 0x7FFFFFFF);
    DevBuf<uint2> tasks;
    DevBuf<unsigned int> nt;
    {
        const char *bv0 = getenv("GW_CN_BUILD");
        if (g->nnz > 0 && (uint32_t)g->max_degree > CC_SMALL && !(bv0 && !strcmp(bv0, "v1"))) { GW_CUDA(tasks.alloc(cap)); GW_CUDA(nt.alloc(1)); }
    }
    cudaEvent_t e0, e1;
    GW_CUDA(cudaEventCreate(&e0)); GW_CUDA(cudaEventCreate(&e1));
    GW_CUDA(cudaEventRecord(e0, st));
    if (g->nnz > 0) {
        int sms = 148;
        device_info(&sms, nullptr);
        const char *np = getenv("GW_CN_RIDX");                   // experiment knob: "0" keeps plain counts (rejection of prev)
        g->nbr4_packed = (g->max_degree < 65536 && !(np && !strcmp(np, "0"))) ? 1 : 0;
        // CTAs of k_cc_small in the launch per SM (8 of 256 threads are resident): 8 / 16 / 32 / 64 / 128 / 256 -> R-MAT-22 22.1 /
        // 21.6 / 20.8 / 20.6 / 20.3 / 20.3 ms, R-MAT-26 443 / 437 / 432 / 426 / 424 / 422 ms (tools/prep_waves_probe.py)
        int cc_ctas = 128;
        if (const char *cs = getenv("GW_CC_CTAS_PER_SM")) cc_ctas = std::max(1, atoi(cs));   // experiment knob
        const char *bv = getenv("GW_CN_BUILD");                  // experiment knob: "v1" = one warp per directed entry
        if (bv && !strcmp(bv, "v1")) {
            k_common_counts_v1<<<sms * 16, 256, 0, st>>>(g->d_meta, g->d_col, g->d_row_ptr, g->n, g->nnz, g->d_nbr4, flag.p, g->nbr4_packed);
            GW_LAUNCHED();
        } else {
            const bool have_tasks = (uint32_t)g->max_degree > CC_SMALL;
            if (!have_tasks) {
                k_cc_small<<<sms * cc_ctas, 256, 0, st>>>(g->d_meta, g->d_col, g->n, g->d_nbr4, flag.p, g->nbr4_packed);
                GW_LAUNCHED();
            } else {
                GW_CUDA(cudaMemsetAsync(nt.p, 0, sizeof(unsigned int), st));
                k_cc_list_tasks<<<(unsigned)((g->n + 255) / 256), 256, 0, st>>>(g->d_meta, g->n, tasks.p, nt.p, cap);
                GW_LAUNCHED();
                unsigned int hn = 0;
                cudaEvent_t t1, t2;
                const bool timing = getenv("GW_TIMING") != nullptr;
                if (timing) { cudaEventCreate(&t1); cudaEventCreate(&t2); }
                GW_CUDA(cudaMemcpyAsync(&hn, nt.p, sizeof(hn), cudaMemcpyDeviceToHost, st));
                k_cc_small<<<sms * cc_ctas, 256, 0, st>>>(g->d_meta, g->d_col, g->n, g->d_nbr4, flag.p, g->nbr4_packed);   // runs while the host waits for the task count
                GW_LAUNCHED();
                GW_CUDA(cudaStreamSynchronize(st));
                if (hn > cap) return fail(GW_E_STATE, "common-neighbour task list overflow (%u > %u)", hn, cap);
                uint32_t slots = 128;
                while (slots < 2u * (uint32_t)g->max_degree && slots < CC_HASH_MAX) slots <<= 1;
                const size_t smem = sizeof(int32_t) * slots;
                GW_CUDA(cudaFuncSetAttribute(k_cc_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                if (timing) cudaEventRecord(t1, st);
                if (hn > 0) {
                    k_cc_block<<<hn, 256, smem, st>>>(g->d_meta, g->d_col, tasks.p, hn, g->d_nbr4, flag.p, g->nbr4_packed);
                    GW_LAUNCHED();
                }
                if (timing) cudaEventRecord(t2, st);
                GW_CUDA(cudaStreamSynchronize(st));
                if (timing) {
                    float a = 0, b = 0;
                    cudaEventElapsedTime(&a, e0, t1); cudaEventElapsedTime(&b, t1, t2);
                    fprintf(stderr, "common counts: task list + warp tasks %.1f ms (%u CTA tasks, %zu B of shared memory each), CTA tasks %.1f ms\n", a, hn, smem, b);
                    cudaEventDestroy(t1); cudaEventDestroy(t2);
                }
            }
        }
    }
    GW_CUDA(cudaEventRecord(e1, st));
    int h = 0;
    GW_CUDA(cudaMemcpyAsync(&h, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    GW_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    g->common_build_ms = ms;
    g->has_self_loops = h;
    if (h) return GW_E_STATE;       // rows/descriptors are valid, counts unusable
    g->nbr4_has_counts = 1;
    return GW_OK;
}

// Lazy Bloom filter over the edge set (16 bits per undirected edge, power-of-two-free sizing).
static int ensure_bloom(gw_graph *g, cudaStream_t st) {
    if (g->d_bloom) return GW_OK;
    const char *off = getenv("GW_BLOOM");
    if (off && !strcmp(off, "0")) return GW_OK;             // experiment knob: exact searches only
    uint64_t nwords = std::max<uint64_t>(1024, (uint64_t)(g->nnz / 2) / 4 + 1);     // 64 bits per 4 edges
    if (cudaMalloc((void **)&g->d_bloom, nwords * 8) != cudaSuccess) { cudaGetLastError(); g->d_bloom = nullptr; return GW_OK; }   // optional
    g->bloom_words = nwords;
    GW_CUDA(cudaMemsetAsync(g->d_bloom, 0, nwords * 8, st));
    int sms = 148;
    device_info(&sms, nullptr);
    k_bloom_build<<<sms * 16, 256, 0, st>>>(g->d_meta, g->d_col, g->n, g->d_bloom, nwords);
    GW_LAUNCHED();
    GW_CUDA(cudaStreamSynchronize(st));
    return GW_OK;
}

// Lazy hash sets of the long rows (heavy-tailed graphs only): 2 slots per directed entry, 8 bytes per entry.
static int ensure_rowhash(gw_graph *g, cudaStream_t st) {
    if (g->d_rowhash || g->nnz == 0) return GW_OK;
    const char *off = getenv("GW_ROWHASH");
    if (off && !strcmp(off, "0")) return GW_OK;             // experiment knob: binary searches only
    if (cudaMalloc((void **)&g->d_rowhash, sizeof(int32_t) * 2 * (size_t)g->nnz) != cudaSuccess) {
        cudaGetLastError(); g->d_rowhash = nullptr; return GW_OK;                       // optional
    }
    int sms = 148;
    device_info(&sms, nullptr);
    k_rowhash_build<<<sms * 16, 256, 0, st>>>(g->d_meta, g->d_col, g->n, g->d_rowhash);
    GW_LAUNCHED();
    GW_CUDA(cudaStreamSynchronize(st));
    return GW_OK;
}
// A packed nbr4 MUST be read by a RIDX instantiation (the count shares .y with the reverse index); it pays where the
// return edge is thinned (r0 < lo) or excluded (q < 1), and costs nothing elsewhere (same loads, two more registers).
static bool use_ridx(const gw_graph *g, const CnParams &) { return g->nbr4_packed != 0; }
static bool is_hub_graph(const gw_graph *g) { return g->max_degree > 2048 || getenv("GW_CN_HUB") != nullptr; }   // env: test knob for small graphs

int launch_walk_cn(gw_graph *g, double p, double q, int32_t L, const int64_t *d_starts, int64_t n_starts, uint64_t seed,
                   uint64_t walk_id_base, int32_t *d_out, int32_t *d_lens, cudaStream_t st) {
    CnParams P;
    if (!(p == 1.0 && q == 1.0)) GW_TRY(ensure_bloom(g, st));
    if (!(p == 1.0 && q == 1.0) && is_hub_graph(g)) GW_TRY(ensure_rowhash(g, st));
    P.meta = g->d_meta; P.col = g->d_col; P.nbr4 = g->d_nbr4; P.starts = d_starts; P.n_walks = n_starts; P.L = L;
    P.bloom = g->d_bloom; P.bloom_words = g->bloom_words; P.rowhash = g->d_rowhash;
    P.a = (float)(1.0 / q); P.b = 1.0f; P.r = (float)(1.0 / p);
    P.lo = std::min(P.a, P.b); P.r0 = std::min(P.r, P.lo);
    P.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    P.walk_id_base = walk_id_base; P.out = d_out; P.lens = d_lens;
    P.stats = nullptr;
    unsigned grid = (unsigned)((n_starts + 255) / 256);
    bool vec = (L % 8 == 0) && ((reinterpret_cast<uintptr_t>(d_out) & 31) == 0);
    const char *occ = getenv("GW_CN_MINB");     // experiment knob: resident blocks per SM the kernel is compiled for
    int minb = occ ? atoi(occ) : 5;   // 5 is best for every p, q since O steps of contexts without common neighbours skip the adjacency test (R-MAT-22 p=4 q=0.5: 39.7 / 38.6 / 36.5 G steps/s at 5 / 6 / 8); q >= 1 loses 7 % at 6
    const bool hub = is_hub_graph(g);
    const bool ridx = use_ridx(g, P);
    if (ridx) {                                   // packed nbr4[].y: every reader of this graph's counts must unpack
        if (!vec) { if (hub) k_walk_cn<false, false, 5, true, true><<<grid, 256, 0, st>>>(P); else k_walk_cn<false, false, 5, false, true><<<grid, 256, 0, st>>>(P); }
        else if (hub) k_walk_cn<true, false, 5, true, true><<<grid, 256, 0, st>>>(P);
        else if (minb >= 6) k_walk_cn<true, false, 6, false, true><<<grid, 256, 0, st>>>(P);
        else k_walk_cn<true, false, 5, false, true><<<grid, 256, 0, st>>>(P);
    }
    else if (!vec) { if (hub) k_walk_cn<false, false, 5, true, false><<<grid, 256, 0, st>>>(P); else k_walk_cn<false, false, 5, false, false><<<grid, 256, 0, st>>>(P); }
    else if (hub) {
        // (the heavy-tailed instantiation carries the rejection paths and spills at 48 registers: GW_CN_MINB=3/4 trade
        // resident CTAs for registers)
        if (minb >= 6) k_walk_cn<true, false, 6, true, false><<<grid, 256, 0, st>>>(P);
        else if (occ && minb == 4) k_walk_cn<true, false, 4, true, false><<<grid, 256, 0, st>>>(P);
        else if (occ && minb == 3) k_walk_cn<true, false, 3, true, false><<<grid, 256, 0, st>>>(P);
        else k_walk_cn<true, false, 5, true, false><<<grid, 256, 0, st>>>(P);
    }
    else if (minb >= 8) k_walk_cn<true, false, 8, false, false><<<grid, 256, 0, st>>>(P);
    else if (minb >= 6) k_walk_cn<true, false, 6, false, false><<<grid, 256, 0, st>>>(P);
    else k_walk_cn<true, false, 5, false, false><<<grid, 256, 0, st>>>(P);
    GW_LAUNCHED();
    return GW_OK;
}

// Byte model of the mixture walker on the SAME walks (same seed / ids): no stores, counters only.
int count_walk_cn(gw_graph *g, double p, double q, int32_t L, const int64_t *d_starts, int64_t n_starts, uint64_t seed,
                  uint64_t walk_id_base, unsigned long long *d_stats, cudaStream_t st) {
    CnParams P;
    if (!(p == 1.0 && q == 1.0)) GW_TRY(ensure_bloom(g, st));
    if (!(p == 1.0 && q == 1.0) && is_hub_graph(g)) GW_TRY(ensure_rowhash(g, st));
    P.meta = g->d_meta; P.col = g->d_col; P.nbr4 = g->d_nbr4; P.starts = d_starts; P.n_walks = n_starts; P.L = L;
    P.bloom = g->d_bloom; P.bloom_words = g->bloom_words; P.rowhash = g->d_rowhash;
    P.a = (float)(1.0 / q); P.b = 1.0f; P.r = (float)(1.0 / p);
    P.lo = std::min(P.a, P.b); P.r0 = std::min(P.r, P.lo);
    P.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    P.walk_id_base = walk_id_base; P.out = nullptr; P.lens = nullptr;
    P.stats = d_stats;
    unsigned grid = (unsigned)((n_starts + 255) / 256);
    if (use_ridx(g, P)) {
        if (is_hub_graph(g)) k_walk_cn<false, true, 5, true, true><<<grid, 256, 0, st>>>(P);
        else k_walk_cn<false, true, 5, false, true><<<grid, 256, 0, st>>>(P);
    }
    else if (is_hub_graph(g)) k_walk_cn<false, true, 5, true, false><<<grid, 256, 0, st>>>(P);
    else k_walk_cn<false, true, 5, false, false><<<grid, 256, 0, st>>>(P);
    GW_LAUNCHED();
    return GW_OK;
}

}  // namespace gw
