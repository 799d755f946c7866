// skipgram.cu — the consumer of the walk corpus: skip-gram with negative sampling on the device.
//
// Replaces node2vec/src/main.py:92-101 learn_embeddings = gensim 0.13.3 (node2vec/requirements.txt:3)
//   Word2Vec(walks, size=dimensions, window=window_size, min_count=0, sg=1, workers=workers, iter=iter)
// with gensim's defaults alpha=0.025, min_alpha=0.0001, sample=1e-3, negative=5, hs=0.  gensim is a third-party
// dependency that is not under /root/reference; its published algorithm, restated:
//   vocabulary scan      scale_vocab: word_probability = (sqrt(cnt / (sample * total)) + 1) * (sample * total) / cnt,
//                        sample_int = round(min(1, probability) * 2^32); a word is dropped from a sentence when
//                        sample_int < a fresh 32-bit random number
//   negative sampling    make_cum_table: words drawn proportionally to cnt^0.75 (here: the cumulative table is
//                        materialised as an inverse lookup of 2^k slots, one random access per draw)
//   train_batch_sg       per sentence (after subsampling), per position i: b = random % window; for every j in
//                        [i - window + b, i + window - b], j != i: fast_sentence_sg_neg(center = sent[i], context = sent[j])
//   fast_sentence_sg_neg row1 = syn0[context]; targets = center (label 1) then `negative` draws (label 0, a draw equal to
//                        the center is skipped); f = row1 . syn1neg[target]; |f| >= 6 skips the target;
//                        g = (label - sigmoid(f)) * alpha (EXP_TABLE of 1000 entries over [-6, 6));
//                        work += g * syn1neg[target]; syn1neg[target] += g * row1; finally row1 += work
//   learning rate        alpha falls linearly from alpha to min_alpha with the raw words processed
//   initialisation       syn0 = (uniform[0,1) - 0.5) / size, syn1neg = 0
// One warp per sentence (a walk), the vector dimension spread over the lanes; sentences run concurrently and update
// shared rows without locks, as gensim's worker threads do (Hogwild).  The corpus stays in HBM -- or is never
// materialised at all: gw_node2vec_embeddings regenerates every pass of walks from its seed (5 ms per pass) instead of
// moving 13-215 GB through PCIe, which is what bounds the walk API end to end.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace gw {

constexpr int SG_EXP_TABLE = 1000;
constexpr float SG_MAX_EXP = 6.0f;
constexpr int SG_NB = 6;                     // targets (positive + negatives) whose rows are requested together
__constant__ float c_exp_table[SG_EXP_TABLE];

struct SgParams {
    const int32_t *walks;
    int64_t n_walks;
    int32_t L;
    float *syn0, *syn1;
    const uint32_t *keep;        // sample_int per word; NULL = no subsampling
    const int32_t *negtab;
    uint32_t negtab_mask;        // table size - 1 (power of two)
    int32_t window, negative;
    double alpha0, alpha_min;
    double words_before, total_words;
    uint2 key;
    uint64_t sentence_id_base;
    unsigned long long *pairs;   // trained (center, context) pairs, for the byte model
};

__device__ __forceinline__ uint32_t sg_next(uint64_t &s) {       // gensim / word2vec.c: next_random * 25214903917 + 11, bits 16..47
    s = s * 25214903917ULL + 11ULL;
    return (uint32_t)(s >> 16);
}

__global__ void k_sg_count(const int32_t *__restrict__ walks, int64_t count, unsigned long long *__restrict__ cnt) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int32_t w = walks[i];
    if (w >= 0) atomicAdd(cnt + w, 1ull);
}

// sample_int of scale_vocab and cnt^0.75 per word
__global__ void k_sg_vocab(const unsigned long long *__restrict__ cnt, int64_t n, double sample, double total, uint32_t *__restrict__ keep,
                           double *__restrict__ pw) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = (double)cnt[i];
    pw[i] = v > 0 ? pow(v, 0.75) : 0.0;
    double prob = 1.0;
    if (sample > 0 && v > 0) {
        const double thr = sample * total;
        prob = (sqrt(v / thr) + 1.0) * (thr / v);
        if (prob > 1.0) prob = 1.0;
    }
    const double si = rint(prob * 4294967296.0);
    keep[i] = si >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)si;
}

// inverse lookup of the cumulative cnt^0.75 table: word i owns the slots [round(cum(i-1) * T), round(cum(i) * T))
__global__ void k_sg_negtab(const double *__restrict__ cum, int64_t n, double total_pw, uint32_t T, int32_t *__restrict__ tab) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double lo = i ? cum[i - 1] : 0.0, hi = cum[i];
    uint32_t a = (uint32_t)llrint(lo / total_pw * (double)T), b = (uint32_t)llrint(hi / total_pw * (double)T);
    if (i == n - 1) b = T;
    for (uint32_t t = a; t < b; t++) tab[t] = (int32_t)i;
}

__global__ void k_sg_init(float *__restrict__ syn0, float *__restrict__ syn1, int64_t n, int32_t dim, uint2 key) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;       // one thread per 4 floats
    const int64_t total = n * (int64_t)dim;
    if (4 * i >= total) return;
    const uint4 r = Philox::gen(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 0x5347u, 0u), key);
    const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
    for (int k = 0; k < 4 && 4 * i + k < total; k++) {
        syn0[4 * i + k] = ((float)(rw[k] >> 8) * (1.0f / 16777216.0f) - 0.5f) / (float)dim;      // (random - 0.5) / size
        syn1[4 * i + k] = 0.0f;
    }
}

// PER consecutive floats of a row, as the widest aligned vector the lane's piece allows (rows start 4 * dim bytes apart from a
// 256-byte aligned base: a lane's piece of 4 / 8 floats is 16-byte aligned, of 2 floats 8-byte aligned)
template <int PER>
__device__ __forceinline__ void ld_row(float (&x)[PER], const float *p) {
    if constexpr (PER % 4 == 0) {
#pragma unroll
        for (int q = 0; q < PER / 4; q++) {
            const float4 v = *reinterpret_cast<const float4 *>(p + 4 * q);
            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
    } else if constexpr (PER == 2) {
        const float2 v = *reinterpret_cast<const float2 *>(p);
        x[0] = v.x; x[1] = v.y;
    } else {
#pragma unroll
        for (int k = 0; k < PER; k++) x[k] = p[k];
    }
}
template <int PER>
__device__ __forceinline__ void st_row(float *p, const float (&x)[PER]) {
    if constexpr (PER % 4 == 0) {
#pragma unroll
        for (int q = 0; q < PER / 4; q++) *reinterpret_cast<float4 *>(p + 4 * q) = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
    } else if constexpr (PER == 2) {
        *reinterpret_cast<float2 *>(p) = make_float2(x[0], x[1]);
    } else {
#pragma unroll
        for (int k = 0; k < PER; k++) p[k] = x[k];
    }
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// One group of up to SG_NB targets of a (word, word2) pair, in gensim's order: tg[u] < 0 = no target in this slot; slot 0
// of the group that starts at d0 == 0 is the positive target (its row lives in yc).  ALL syn1neg rows of the group are
// requested at once, then the group is processed target by target -- six 512-byte rows in flight per warp instead of
// one.  A negative drawn twice inside a group is re-read after the first update, so the arithmetic is that of the
// one-row-at-a-time loop, operation for operation.
template <int PER>
__device__ __forceinline__ void sg_group(const SgParams &P, const int32_t (&tg)[SG_NB], int d0, int lane, float alpha,
                                         const float (&x)[PER], float (&work)[PER], float (&yc)[PER]) {
    constexpr int dim = 32 * PER;
    float yv[SG_NB][PER];
#pragma unroll
    for (int u = 0; u < SG_NB; u++)
        if (tg[u] >= 0 && d0 + u != 0) ld_row<PER>(yv[u], P.syn1 + (size_t)tg[u] * dim + lane * PER);
#pragma unroll
    for (int u = 0; u < SG_NB; u++) {
        if (tg[u] < 0) continue;
        const bool pos = d0 + u == 0;
        float *r2 = P.syn1 + (size_t)tg[u] * dim + lane * PER;
        bool dup = false;
#pragma unroll
        for (int e = 0; e < u; e++) dup |= (e + d0 != 0) && tg[e] == tg[u];
        float y[PER], f = 0.0f;
        if (dup) ld_row<PER>(yv[u], r2);
#pragma unroll
        for (int k = 0; k < PER; k++) { y[k] = pos ? yc[k] : yv[u][k]; f = fmaf(x[k], y[k], f); }
        for (int o = 16; o; o >>= 1) f += __shfl_xor_sync(0xffffffffu, f, o);
        if (f <= -SG_MAX_EXP || f >= SG_MAX_EXP) continue;
        const float sig = c_exp_table[(int)((f + SG_MAX_EXP) * (SG_EXP_TABLE / SG_MAX_EXP / 2.0f))];
        const float g = ((pos ? 1.0f : 0.0f) - sig) * alpha;
        float upd[PER];
#pragma unroll
        for (int k = 0; k < PER; k++) {
            work[k] = fmaf(g, y[k], work[k]);
            upd[k] = fmaf(g, x[k], y[k]);
        }
        if (pos) {
#pragma unroll
            for (int k = 0; k < PER; k++) yc[k] = upd[k];
        } else {
            st_row<PER>(r2, upd);
        }
    }
}

// subsampling + compaction of one sentence into shared memory (train_batch_sg: words that fail the draw vanish); returns
// its length.  Called by all 32 lanes.
__device__ __forceinline__ int sg_compact(const SgParams &P, const int32_t *row, uint64_t sid, int lane, int32_t *sent) {
    int m = 0;
    for (int base = 0; base < P.L; base += 32) {
        const int pos = base + lane;
        int32_t w = pos < P.L ? row[pos] : -1;
        bool keep = w >= 0;
        if (keep && P.keep) {
            const uint4 r = Philox::gen(make_uint4((uint32_t)sid, (uint32_t)(sid >> 32), (uint32_t)(pos >> 2), 0x5342u), P.key);
            const uint32_t rw = (pos & 3) == 0 ? r.x : (pos & 3) == 1 ? r.y : (pos & 3) == 2 ? r.z : r.w;
            keep = !(P.keep[w] < rw);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (keep) sent[m + __popc(bal & ((1u << lane) - 1))] = w;
        m += __popc(bal);
    }
    __syncwarp();
    return m;
}

__device__ __forceinline__ float sg_alpha(const SgParams &P, int64_t s) {   // linear in the raw words seen before this sentence
    double prog = (P.words_before + (double)s * (double)P.L) / P.total_words;
    if (prog > 1.0) prog = 1.0;
    return (float)fmax(P.alpha_min, P.alpha0 - (P.alpha0 - P.alpha_min) * prog);
}

__device__ __forceinline__ uint64_t sg_stream(const SgParams &P, uint64_t sid) {   // one scalar random stream per sentence, replicated in every lane
    const uint4 r = Philox::gen(make_uint4((uint32_t)sid, (uint32_t)(sid >> 32), 0u, 0x5353u), P.key);
    return ((uint64_t)r.x << 32) | r.y;
}

// PER floats of a row per lane (dim = 32 * PER); one warp per sentence.  General kernel: any number of negatives, the
// targets of a pair in groups of SG_NB.
template <int PER>
__global__ void __launch_bounds__(256, PER <= 4 ? 3 : 2) k_sgns(SgParams P) {
    extern __shared__ int32_t s_sent[];                              // [warps per CTA][L]
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int32_t *sent = s_sent + (size_t)wib * P.L;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    constexpr int dim = 32 * PER;
    unsigned long long my_pairs = 0;
    for (int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; s < P.n_walks; s += nwarps) {
        const uint64_t sid = P.sentence_id_base + (uint64_t)s;
        const int m = sg_compact(P, P.walks + s * P.L, sid, lane, sent);
        if (m < 2) continue;
        const float alpha = sg_alpha(P, s);
        uint64_t rs = sg_stream(P, sid);
        for (int i = 0; i < m; i++) {
            const int32_t center = sent[i];
            const int b = (int)(sg_next(rs) % (uint32_t)P.window);               // reduced_windows[i]
            const int j0 = max(0, i - P.window + b), j1 = min(m, i + P.window + 1 - b);
            // the positive row syn1neg[center] is target 0 of EVERY pair of this center: it stays in registers across the
            // window (same operations in the same order as reading and writing it per pair; a negative draw equal to
            // the center is skipped by the algorithm, so nothing else touches the row in between)
            float *rc = P.syn1 + (size_t)center * dim + lane * PER;
            float yc[PER];
            ld_row<PER>(yc, rc);
            for (int j = j0; j < j1; j++) {
                if (j == i) continue;
                float *r1 = P.syn0 + (size_t)sent[j] * dim + lane * PER;
                float x[PER], work[PER];
                ld_row<PER>(x, r1);
#pragma unroll
                for (int k = 0; k < PER; k++) work[k] = 0.0f;
                for (int d0 = 0; d0 <= P.negative; d0 += SG_NB) {
                    int32_t tg[SG_NB];                                           // the draws depend on the random stream only: made first
#pragma unroll
                    for (int u = 0; u < SG_NB; u++) {
                        const int d = d0 + u;
                        tg[u] = -1;
                        if (d == 0) tg[u] = center;
                        else if (d <= P.negative) {
                            const int32_t t = P.negtab[sg_next(rs) & P.negtab_mask];
                            tg[u] = t == center ? -1 : t;                       // a draw equal to the centre word is skipped
                        }
                    }
                    sg_group<PER>(P, tg, d0, lane, alpha, x, work, yc);
                }
#pragma unroll
                for (int k = 0; k < PER; k++) x[k] += work[k];
                st_row<PER>(r1, x);
                my_pairs++;
            }
            st_row<PER>(rc, yc);
        }
        __syncwarp();
    }
    if (lane == 0 && my_pairs && P.pairs) atomicAdd(P.pairs, my_pairs);
}

// Production kernel for negative < SG_NB (gensim's default 5): the same loops, software-pipelined over the PAIRS of a
// sentence.  The random stream of a sentence does not depend on the data (one draw per centre, `negative` draws per pair),
// so a generator runs two pairs ahead of the arithmetic: the negative-table lookups of pair s + 2 are issued, the rows of
// pair s + 1 (context, negatives, a new centre) are prefetched into L2, then pair s is trained.  A pair no longer waits
// for two dependent DRAM latencies (table, then rows) but for L2 hits; values are still LOADED after the previous pair's
// stores, so the arithmetic is unchanged (one warp in order reproduces the restatement exactly as before).
constexpr int SG_PN = SG_NB - 1;
struct SgPair {
    int i, j;
    int32_t t[SG_PN];            // raw draws of the negative table (compared with the centre when used)
};

template <int PER>
__global__ void __launch_bounds__(256, PER <= 4 ? 3 : 2) k_sgns_pipe(SgParams P) {
    extern __shared__ int32_t s_sent[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int32_t *sent = s_sent + (size_t)wib * P.L;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    constexpr int dim = 32 * PER;
    unsigned long long my_pairs = 0;
    for (int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; s < P.n_walks; s += nwarps) {
        const uint64_t sid = P.sentence_id_base + (uint64_t)s;
        const int m = sg_compact(P, P.walks + s * P.L, sid, lane, sent);
        if (m < 2) continue;
        const float alpha = sg_alpha(P, s);
        uint64_t rs = sg_stream(P, sid);
        int gi = -1, gj = 0, gj1 = 0;                                  // generator: centre, next context position, end of the window
        auto gen = [&](SgPair &d) -> bool {
            for (;;) {
                if (gj < gj1) {
                    if (gj == gi) { gj++; continue; }
                    d.i = gi; d.j = gj++;
#pragma unroll
                    for (int u = 0; u < SG_PN; u++) d.t[u] = u < P.negative ? P.negtab[sg_next(rs) & P.negtab_mask] : -1;
                    return true;
                }
                if (++gi >= m) return false;
                const int b = (int)(sg_next(rs) % (uint32_t)P.window);           // reduced_windows[i]
                gj = max(0, gi - P.window + b); gj1 = min(m, gi + P.window + 1 - b);
            }
        };
        SgPair d0, d1, d2;
        bool h0 = gen(d0), h1 = h0 && gen(d1);
        int ci = -1;                                                   // centre whose syn1neg row is in yc
        float *rc = nullptr;
        float yc[PER];
        while (h0) {
            const bool h2 = h1 && gen(d2);                             // table lookups of pair s + 2 leave now
            if (h1) {                                                  // rows of pair s + 1 -> L2
                prefetch_l2(P.syn0 + (size_t)sent[d1.j] * dim + lane * PER);
#pragma unroll
                for (int u = 0; u < SG_PN; u++)
                    if (d1.t[u] >= 0) prefetch_l2(P.syn1 + (size_t)d1.t[u] * dim + lane * PER);
                if (d1.i != d0.i) prefetch_l2(P.syn1 + (size_t)sent[d1.i] * dim + lane * PER);
            }
            if (d0.i != ci) {                                          // new centre: its row stays in registers across the window
                if (ci >= 0) st_row<PER>(rc, yc);
                ci = d0.i;
                rc = P.syn1 + (size_t)sent[ci] * dim + lane * PER;
                ld_row<PER>(yc, rc);
            }
            const int32_t center = sent[ci];
            float *r1 = P.syn0 + (size_t)sent[d0.j] * dim + lane * PER;
            float x[PER], work[PER];
            ld_row<PER>(x, r1);
#pragma unroll
            for (int k = 0; k < PER; k++) work[k] = 0.0f;
            int32_t tg[SG_NB];
            tg[0] = center;
#pragma unroll
            for (int u = 0; u < SG_PN; u++) tg[u + 1] = d0.t[u] == center ? -1 : d0.t[u];   // a draw equal to the centre word is skipped
            sg_group<PER>(P, tg, 0, lane, alpha, x, work, yc);
#pragma unroll
            for (int k = 0; k < PER; k++) x[k] += work[k];
            st_row<PER>(r1, x);
            my_pairs++;
            d0 = d1; d1 = d2; h0 = h1; h1 = h2;
        }
        if (ci >= 0) st_row<PER>(rc, yc);
        __syncwarp();
    }
    if (lane == 0 && my_pairs && P.pairs) atomicAdd(P.pairs, my_pairs);
}

}  // namespace gw

struct gw_sgns {
    int device = 0;
    int64_t n = 0;
    int32_t dim = 0;
    uint64_t seed = 0;
    float *syn0 = nullptr, *syn1 = nullptr;
    unsigned long long *cnt = nullptr;       // raw counts per word
    uint32_t *keep = nullptr;
    int32_t *negtab = nullptr;
    uint32_t negtab_size = 0;
    unsigned long long *pairs = nullptr;
    double total_words = 0;                  // raw words of the scanned corpus
    int32_t negative = 0;
    bool vocab_ready = false, table_loaded = false;
};

using namespace gw;

static int sg_check(const gw_sgns *m) {
    if (!m) return fail(GW_E_INVALID, "model is NULL");
    return GW_OK;
}

extern "C" {

int gw_sgns_create(int64_t n_words, int32_t dimensions, uint64_t seed, int32_t device, gw_sgns **out) {
    if (n_words < 0 || n_words > 0x7FFFFFFF || !out) return fail(GW_E_INVALID, "bad arguments");
    if (dimensions != 32 && dimensions != 64 && dimensions != 128 && dimensions != 256)
        return fail(GW_E_INVALID, "dimensions must be 32, 64, 128 or 256 (one warp per vector, 1/2/4/8 floats per lane)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(GW_E_CUDA, "no CUDA device is visible (this library has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(GW_E_INVALID, "device %d is outside [0, %d)", device, ndev);
    GW_CUDA(cudaSetDevice(device));
    struct { int device; int64_t n; } gg{device, n_words}, *g = &gg;
    gw_sgns *m = new gw_sgns;
    m->device = g->device; m->n = g->n; m->dim = dimensions; m->seed = seed;
    const size_t cells = (size_t)std::max<int64_t>(g->n, 1) * dimensions;
    if (cudaMalloc((void **)&m->syn0, cells * 4) != cudaSuccess || cudaMalloc((void **)&m->syn1, cells * 4) != cudaSuccess ||
        cudaMalloc((void **)&m->cnt, sizeof(unsigned long long) * (size_t)std::max<int64_t>(g->n, 1)) != cudaSuccess ||
        cudaMalloc((void **)&m->keep, sizeof(uint32_t) * (size_t)std::max<int64_t>(g->n, 1)) != cudaSuccess ||
        cudaMalloc((void **)&m->pairs, sizeof(unsigned long long)) != cudaSuccess) {
        cudaGetLastError();
        gw_sgns_free(m);
        return fail(GW_E_TOO_LARGE, "two %lld x %d fp32 matrices do not fit on the device", (long long)g->n, dimensions);
    }
    GW_CUDA(cudaMemset(m->cnt, 0, sizeof(unsigned long long) * (size_t)std::max<int64_t>(g->n, 1)));
    GW_CUDA(cudaMemset(m->pairs, 0, sizeof(unsigned long long)));
    if (g->n > 0) {
        k_sg_init<<<(unsigned)((cells / 4 + 256) / 256), 256>>>(m->syn0, m->syn1, g->n, dimensions, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        GW_LAUNCHED();
    }
    float tab[SG_EXP_TABLE];
    for (int i = 0; i < SG_EXP_TABLE; i++) {                  // word2vec_inner.pyx: EXP_TABLE[i] = exp((i / 1000 * 2 - 1) * 6) / (exp(..) + 1)
        const float e = (float)exp((i / (double)SG_EXP_TABLE * 2.0 - 1.0) * SG_MAX_EXP);
        tab[i] = e / (e + 1.0f);
    }
    GW_CUDA(cudaMemcpyToSymbol(c_exp_table, tab, sizeof(tab)));
    GW_CUDA(cudaDeviceSynchronize());
    *out = m;
    return GW_OK;
}

int gw_sgns_free(gw_sgns *m) {
    if (!m) return GW_OK;
    cudaSetDevice(m->device);
    cudaFree(m->syn0); cudaFree(m->syn1); cudaFree(m->cnt); cudaFree(m->keep); cudaFree(m->negtab); cudaFree(m->pairs);
    delete m;
    return GW_OK;
}

int gw_sgns_count_dev(gw_sgns *m, const int32_t *d_walks, int64_t n_walks, int32_t walk_length, void *stream) {
    GW_TRY(sg_check(m));
    if (n_walks < 0 || walk_length < 1 || (n_walks > 0 && !d_walks)) return fail(GW_E_INVALID, "bad arguments");
    if (n_walks == 0) return GW_OK;
    GW_CUDA(cudaSetDevice(m->device));
    const int64_t count = n_walks * walk_length;
    k_sg_count<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_walks, count, m->cnt);
    GW_LAUNCHED();
    m->vocab_ready = false;
    return GW_OK;
}

int gw_sgns_finalize_vocab(gw_sgns *m, double sample, int32_t negative) {
    GW_TRY(sg_check(m));
    if (sample < 0 || negative < 1 || negative > 64) return fail(GW_E_INVALID, "sample must be >= 0 and negative in 1..64");
    GW_CUDA(cudaSetDevice(m->device));
    GW_CUDA(cudaDeviceSynchronize());
    const int64_t n = m->n;
    if (n == 0) { m->vocab_ready = true; return GW_OK; }
    std::vector<unsigned long long> h((size_t)n);
    GW_CUDA(cudaMemcpy(h.data(), m->cnt, sizeof(unsigned long long) * (size_t)n, cudaMemcpyDeviceToHost));
    double total = 0;
    for (int64_t i = 0; i < n; i++) total += (double)h[i];
    if (total <= 0) return fail(GW_E_STATE, "the vocabulary scan saw no word: call gw_sgns_count_dev on the corpus first");
    m->total_words = total;
    DevBuf<double> pw;
    GW_CUDA(pw.alloc((size_t)n));
    k_sg_vocab<<<(unsigned)((n + 255) / 256), 256>>>(m->cnt, n, sample, total, m->keep, pw.p);
    GW_LAUNCHED();
    // cumulative cnt^0.75 on the host in fp64, left to right (make_cum_table's order); 8 bytes per vertex, once
    std::vector<double> cum((size_t)n);
    GW_CUDA(cudaMemcpy(cum.data(), pw.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    double run = 0;
    for (int64_t i = 0; i < n; i++) { run += cum[i]; cum[i] = run; }
    GW_CUDA(cudaMemcpy(pw.p, cum.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
    uint32_t T = 1u << 20;
    while ((int64_t)T < 32 * n && T < (1u << 28)) T <<= 1;                         // >= 32 slots per word on average
    if (m->negtab_size != T) {
        cudaFree(m->negtab); m->negtab = nullptr; m->negtab_size = 0;
        GW_CUDA(cudaMalloc((void **)&m->negtab, sizeof(int32_t) * (size_t)T));
        m->negtab_size = T;
    }
    k_sg_negtab<<<(unsigned)((n + 255) / 256), 256>>>(pw.p, n, run, T, m->negtab);
    GW_LAUNCHED();
    GW_CUDA(cudaDeviceSynchronize());
    m->negative = negative;
    m->vocab_ready = true;
    return GW_OK;
}

int gw_sgns_train_dev(gw_sgns *m, const int32_t *d_walks, int64_t n_walks, int32_t walk_length, int32_t window, double alpha,
                      double min_alpha, double words_before, double total_words, uint64_t sentence_id_base, int32_t subsample,
                      int32_t sequential, void *stream) {
    GW_TRY(sg_check(m));
    if (!m->vocab_ready) return fail(GW_E_STATE, "gw_sgns_finalize_vocab has not been called since the last vocabulary scan");
    if (n_walks < 0 || walk_length < 1 || walk_length > 4096 || (n_walks > 0 && !d_walks)) return fail(GW_E_INVALID, "bad arguments (walk_length <= 4096)");
    if (window < 1 || !(alpha > 0) || min_alpha < 0 || min_alpha > alpha || !(total_words > 0)) return fail(GW_E_INVALID, "bad window / alpha / total_words");
    if (n_walks == 0 || m->n == 0) return GW_OK;
    GW_CUDA(cudaSetDevice(m->device));
    SgParams P;
    P.walks = d_walks; P.n_walks = n_walks; P.L = walk_length; P.syn0 = m->syn0; P.syn1 = m->syn1;
    P.keep = subsample ? m->keep : nullptr; P.negtab = m->negtab; P.negtab_mask = m->negtab_size - 1;
    P.window = window; P.negative = m->negative; P.alpha0 = alpha; P.alpha_min = min_alpha;
    P.words_before = words_before; P.total_words = total_words;
    P.key = make_uint2((uint32_t)m->seed, (uint32_t)(m->seed >> 32)); P.sentence_id_base = sentence_id_base; P.pairs = m->pairs;
    int sms = 148;
    device_info(&sms, nullptr);
    // Hogwild needs far more rows than writers: every warp holds ~7 rows between their load and their store, and an update
    // another warp makes to one of them in between is lost.  At most one concurrent warp per 16 words (karate: 2 warps,
    // R-MAT-22: no limit below the grid): tools/sg_hogwild_probe.py -- on 34 words 8 warps of the pipelined kernel lose
    // 0.05 of edge AUC against the sequential order, 2 warps 0.01; at 4 M words the cap is never reached
    int64_t max_warps = std::max<int64_t>(1, m->n / 16);
    if (const char *mw = getenv("GW_SG_WARPS")) max_warps = std::max<int64_t>(1, atoll(mw));     // experiment knob
    const int threads = sequential ? 32 : 32 * (int)std::min<int64_t>(8, max_warps);
    // CTAs in the launch per SM (3 of 256 threads are resident at dimensions <= 128): 8 -> 836 M pairs/s, 3 -> 870, 32 -> 873,
    // 128 -> 882 (2^20 walks of the R-MAT-22 corpus): many short CTAs leave no tail (the SMs do not all run at the same rate: profiles/r2_gather_waves.txt)
    int ctas_per_sm = 128;
    if (const char *cs = getenv("GW_SG_CTAS_PER_SM")) ctas_per_sm = std::max(1, atoi(cs));      // experiment knob
    const unsigned grid = sequential ? 1u : (unsigned)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>((n_walks + 7) / 8, (int64_t)sms * ctas_per_sm), max_warps / (threads / 32)));
    const size_t smem = sizeof(int32_t) * (size_t)(threads / 32) * walk_length;
    cudaStream_t st = (cudaStream_t)stream;
    const char *pe = getenv("GW_SG_PIPE");                  // "0": the unpipelined kernel (A/B measurements)
    const bool pipe = m->negative < SG_NB && !(pe && !strcmp(pe, "0"));
#define GW_SG(PER_) do { if (pipe) k_sgns_pipe<PER_><<<grid, threads, smem, st>>>(P); else k_sgns<PER_><<<grid, threads, smem, st>>>(P); } while (0)
    switch (m->dim) {
        case 32: GW_SG(1); break;
        case 64: GW_SG(2); break;
        case 128: GW_SG(4); break;
        default: GW_SG(8); break;
    }
#undef GW_SG
    GW_LAUNCHED();
    return GW_OK;
}

int gw_sgns_info(const gw_sgns *m, int64_t *n, int32_t *dimensions, double *total_words, int64_t *trained_pairs) {
    GW_TRY(sg_check(m));
    if (n) *n = m->n;
    if (dimensions) *dimensions = m->dim;
    if (total_words) *total_words = m->total_words;
    if (trained_pairs) {
        unsigned long long h = 0;
        GW_CUDA(cudaSetDevice(m->device));
        GW_CUDA(cudaDeviceSynchronize());
        GW_CUDA(cudaMemcpy(&h, m->pairs, sizeof(h), cudaMemcpyDeviceToHost));
        *trained_pairs = (int64_t)h;
    }
    return GW_OK;
}

int gw_sgns_vectors(const gw_sgns *m, float *out_syn0, float *out_syn1neg, int64_t *out_counts) {
    GW_TRY(sg_check(m));
    GW_CUDA(cudaSetDevice(m->device));
    GW_CUDA(cudaDeviceSynchronize());
    const size_t cells = (size_t)m->n * m->dim;
    if (out_syn0 && cells) GW_CUDA(cudaMemcpy(out_syn0, m->syn0, cells * 4, cudaMemcpyDeviceToHost));
    if (out_syn1neg && cells) GW_CUDA(cudaMemcpy(out_syn1neg, m->syn1, cells * 4, cudaMemcpyDeviceToHost));
    if (out_counts && m->n) {
        static_assert(sizeof(unsigned long long) == sizeof(int64_t), "count width");
        GW_CUDA(cudaMemcpy(out_counts, m->cnt, sizeof(int64_t) * (size_t)m->n, cudaMemcpyDeviceToHost));
    }
    return GW_OK;
}

int gw_sgns_set_vectors(gw_sgns *m, const float *syn0, const float *syn1neg) {
    GW_TRY(sg_check(m));
    GW_CUDA(cudaSetDevice(m->device));
    const size_t cells = (size_t)m->n * m->dim;
    if (syn0 && cells) GW_CUDA(cudaMemcpy(m->syn0, syn0, cells * 4, cudaMemcpyHostToDevice));
    if (syn1neg && cells) GW_CUDA(cudaMemcpy(m->syn1, syn1neg, cells * 4, cudaMemcpyHostToDevice));
    return GW_OK;
}

// main.py:104-114 from simulate_walks on, without the corpus ever leaving the device -- or existing as a whole: every
// pass of walks (one walk per entry of starts[w * n_starts ...], walk ids w * n_starts + i) is REGENERATED from the
// seed whenever it is needed: once for the vocabulary scan, once per training epoch.
int gw_node2vec_embeddings(gw_graph *g, double p, double q, int32_t walk_length, int32_t num_walks, const int64_t *starts,
                           int64_t n_starts, int32_t dimensions, int32_t window, int32_t iter, int32_t negative, double sample,
                           double alpha, double min_alpha, uint64_t seed, float *out_vectors, int64_t *out_counts,
                           double *out_seconds3) {
    if (!g) return fail(GW_E_INVALID, "graph is NULL");
    if (num_walks < 1 || iter < 1 || n_starts < 0 || (n_starts > 0 && !starts) || !out_vectors) return fail(GW_E_INVALID, "bad arguments");
    if (walk_length < 2) return fail(GW_E_INVALID, "walk_length must be at least 2");
    GW_CUDA(cudaSetDevice(g->device));
    gw_sgns *m = nullptr;
    GW_TRY(gw_sgns_create(g->n, dimensions, seed, g->device, &m));
    struct Guard { gw_sgns *m; ~Guard() { gw_sgns_free(m); } } guard{m};
    if (n_starts == 0) return gw_sgns_vectors(m, out_vectors, nullptr, out_counts);
    DevBuf<int64_t> ds;
    DevBuf<int32_t> dw;
    GW_CUDA(ds.alloc((size_t)n_starts * num_walks));
    if (dw.alloc((size_t)n_starts * walk_length) != cudaSuccess) { cudaGetLastError(); return fail(GW_E_TOO_LARGE, "one pass of walks does not fit on the device"); }
    GW_CUDA(cudaMemcpy(ds.p, starts, sizeof(int64_t) * (size_t)n_starts * num_walks, cudaMemcpyHostToDevice));
    for (int64_t i = 0; i < n_starts * num_walks; i++)
        if (starts[i] < 0 || starts[i] >= g->n) return fail(GW_E_KEY, "start node index %lld is not a vertex of the graph", (long long)starts[i]);
    cudaEvent_t ev[4];
    for (auto &e : ev) GW_CUDA(cudaEventCreate(&e));
    float t_walk = 0, t_scan = 0, t_train = 0;
    auto pass = [&](int w) -> int {
        return gw_node2vec_walks_dev(g, p, q, walk_length, ds.p + (size_t)w * n_starts, n_starts, seed, (uint64_t)w * (uint64_t)n_starts, dw.p, nullptr, nullptr);
    };
    // vocabulary scan (build_vocab): one pass over the corpus
    for (int w = 0; w < num_walks; w++) {
        cudaEventRecord(ev[0]);
        GW_TRY(pass(w));
        cudaEventRecord(ev[1]);
        GW_TRY(gw_sgns_count_dev(m, dw.p, n_starts, walk_length, nullptr));
        cudaEventRecord(ev[2]);
        GW_CUDA(cudaEventSynchronize(ev[2]));
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, ev[0], ev[1]); cudaEventElapsedTime(&b, ev[1], ev[2]);
        t_walk += a; t_scan += b;
    }
    GW_TRY(gw_sgns_finalize_vocab(m, sample, negative));
    const double total = m->total_words * iter;
    double before = 0;
    for (int e = 0; e < iter; e++)
        for (int w = 0; w < num_walks; w++) {
            cudaEventRecord(ev[0]);
            GW_TRY(pass(w));
            cudaEventRecord(ev[1]);
            GW_TRY(gw_sgns_train_dev(m, dw.p, n_starts, walk_length, window, alpha, min_alpha, before, total,
                                     ((uint64_t)e * num_walks + w) * (uint64_t)n_starts, sample > 0 ? 1 : 0, 0, nullptr));
            cudaEventRecord(ev[2]);
            GW_CUDA(cudaEventSynchronize(ev[2]));
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, ev[0], ev[1]); cudaEventElapsedTime(&b, ev[1], ev[2]);
            t_walk += a; t_train += b;
            before += (double)n_starts * walk_length;
        }
    for (auto &e : ev) cudaEventDestroy(e);
    if (out_seconds3) { out_seconds3[0] = t_walk * 1e-3; out_seconds3[1] = t_scan * 1e-3; out_seconds3[2] = t_train * 1e-3; }
    return gw_sgns_vectors(m, out_vectors, nullptr, out_counts);
}

}  // extern "C"
