// simrank.cu — DeepSim/TopSim Monte-Carlo SimRank: fused walk-and-meet kernel + per-query top-k.
//
// Replaces DeepSim/TopSimAll/src/simrank/SingleRandomWalk.java:39-106 (compute / walk /
// computePathSim / isFirstMeet), utils/Print.java:31-41 + lxctools/FixedMaxPQ.java (top-k) and
// simrank/SimRank.java:36-77 (exact iteration).
//
// One persistent CTA per in-flight query.  Each thread owns whole samples: it walks 2*STEP
// uniform steps from the query vertex with the path in registers, tests first meeting
// (path[j] != path[2i-j] for j < i) and adds C^i * deg(path[i]) / deg(path[2i]) / SAMPLE to the
// query's accumulator, which is a two-tier open-addressing hash table: tier 1 in shared memory
// (hot, many-hit targets), tier 2 in a per-CTA global scratch that stays in L2 (the long tail of
// single-hit targets).  Scores are accumulated as 32.32 fixed point with native 32-bit atomics
// (low word add + carry into the high word), so a query's result is bit-identical for any
// launch geometry, thread interleaving or GPU count.  Top-k = one histogram pass over tier 1
// (lower bound on the k-th score), one candidate-collection pass over the touched slots, rank
// by counting among the candidates; a k-round arg-max fallback covers mass ties.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "common.cuh"

namespace gw {

#ifndef SR_BLOCK_THREADS
#define SR_BLOCK_THREADS 512
#endif
constexpr int SR_BLOCK = SR_BLOCK_THREADS;
#ifndef SR_ILP
#define SR_ILP 1                     // samples walked in lock step per thread (independent loads in flight)
#endif
constexpr int SR_HS = 8192;          // tier-1 slots (shared memory)
constexpr int SR_T1_PROBES = 8;      // bounded probing in tier 1, then fall through to tier 2
constexpr int SR_BINS = 1024;
constexpr int SR_CAND = 512;
constexpr uint32_t SR_EMPTY = 0xFFFFFFFFu;
constexpr double SR_FIX = 4294967296.0;   // 2^32

struct SimrankParams {
    const uint2 *meta;
    const int32_t *col;
    const int4 *nbr4;                // {neighbour, -, offset(nbr), degree(nbr)}: ONE random access per walk step
    const int64_t *queries;
    int64_t nq;
    int64_t n;
    int32_t sample;
    int32_t k;
    float coef[16];                  // C^i / SAMPLE (i = 1..STEP)
    double cpow64[16];               // C^i in fp64 (arithmetic-reference kernel only)
    uint2 key;
    uint64_t query_id_base;
    // per-CTA global scratch
    uint32_t *gkeys, *olist;
    unsigned long long *gval;        // tier-2 scores, 32.32 fixed point, one native 64-bit RED per hit
    uint32_t gs_mask;                // tier-2 slots - 1 (power of two)
    uint32_t olist_cap;
    // outputs
    int32_t *out_ids;
    double *out_scores;
    double *out_dense;               // optional dense rows [nq * n]
    unsigned long long *steps;
    int *err;
    // log kernel: per-CTA append log; slow-path hand-over list
    uint2 *log;
    uint32_t log_cap;
    int32_t *qlist_out;              // written by the log kernel
    const int32_t *qlist;            // read by the hash kernel (NULL = all queries)
    uint32_t *qcount;
    unsigned long long *prof;        // SR_PROFILE builds: per-phase clock64 totals of CTA 0
    uint32_t *work;                  // path-tree kernel: [0] / [1] = next unclaimed query of the log / exact launch
    // Scratch SLOTS: the log and path-tree kernels are launched with many more CTAs than fit the machine (one 512/1024-thread
    // CTA per SM is resident); a CTA takes one of `nslots` per-CTA scratch areas when it starts and gives it back when it
    // ends.  nslots >= the CTAs that can be resident, so a starting CTA always finds one.
    uint32_t *slots;                 // [nslots] 0 = free
    uint32_t nslots;
    uint32_t quota;                  // path-tree kernel: queries a CTA claims before it ends (0 = until none is left)
    double out_scale;                // fixed point -> score: 2^-32 (Monte Carlo), SAMPLE * 2^-32 (path tree, x SAMPLE as the reference)
    double inv_sample;               // path tree: contributions are accumulated / SAMPLE so that they fit the 0.32 fixed-point table
};

struct SrShared {
    uint32_t keys[SR_HS];
    uint32_t lo[SR_HS];
    uint32_t hi[SR_HS];
    uint32_t hist[SR_BINS];
    unsigned long long cand_score[SR_CAND];
    uint32_t cand_id[SR_CAND];
    uint32_t ocount;                 // touched slots (both tiers)
    uint32_t ccount;                 // candidates
    uint32_t thr_bin;
    unsigned long long red_score[SR_BLOCK / 32];
    uint32_t red_id[SR_BLOCK / 32];
    unsigned long long sel_score;
    uint32_t sel_id;
};

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

__device__ __forceinline__ uint32_t score_bin(unsigned long long s) {   // monotone, 16 bins per octave
    if (s == 0) return 0;
    int p = 63 - __clzll((long long)s);
    uint32_t sub = p >= 4 ? (uint32_t)((s >> (p - 4)) & 15) : (uint32_t)((s << (4 - p)) & 15);
    return (uint32_t)p * 16 + sub;
}

// (score desc, id asc) strict order: is a better than b ?
__device__ __forceinline__ bool better(unsigned long long sa, uint32_t ia, unsigned long long sb, uint32_t ib) {
    return sa > sb || (sa == sb && ia < ib);
}

__device__ __forceinline__ void fixed_add(uint32_t *lo, uint32_t *hi, unsigned long long v) {
    uint32_t vl = (uint32_t)v;
    uint32_t old = atomicAdd(lo, vl);
    uint32_t carry = (uint32_t)(old + vl < old);
    uint32_t vh = (uint32_t)(v >> 32) + carry;
    if (vh) atomicAdd(hi, vh);
}

// Warp-lock-step insert: all 32 lanes probe together (one slot per pending lane per round) and
// leave together, so the five inserts of a sample never desynchronise the warp.  Tier 1 (shared
// memory) is probed a bounded number of rounds; what does not fit goes to tier 2 (global, L2):
// one CAS that either claims or matches the slot, then a fire-and-forget 64-bit RED.
__device__ __forceinline__ void acc_add_warp(SrShared &S, const SimrankParams &P, uint32_t *gkeys,
                                             unsigned long long *gval, uint32_t *olist, bool has, uint32_t key,
                                             unsigned long long v) {
    const uint32_t h = hash32(key);
    bool pending = has;
    uint32_t slot = h & (SR_HS - 1);
#pragma unroll 1
    for (int pr = 0; pr < SR_T1_PROBES; pr++) {
        if (!__any_sync(0xffffffffu, pending)) return;
        if (pending) {
            uint32_t k0 = ((volatile uint32_t *)S.keys)[slot];
            if (k0 == SR_EMPTY) {
                k0 = atomicCAS(&S.keys[slot], SR_EMPTY, key);
                if (k0 == SR_EMPTY) {
                    uint32_t o = atomicAdd(&S.ocount, 1u);
                    if (o < P.olist_cap) olist[o] = slot; else atomicExch(P.err, 1);
                    k0 = key;
                }
            }
            if (k0 == key) { fixed_add(&S.lo[slot], &S.hi[slot], v); pending = false; }
            else slot = (slot + 1) & (SR_HS - 1);
        }
    }
    uint32_t g = ((h >> 3) * 0x9E3779B1u) & P.gs_mask;
#pragma unroll 1
    for (uint32_t pr = 0; pr <= P.gs_mask; pr++) {
        if (!__any_sync(0xffffffffu, pending)) return;
        if (pending) {
            uint32_t k0 = atomicCAS(&gkeys[g], SR_EMPTY, key);
            if (k0 == SR_EMPTY) {
                uint32_t o = atomicAdd(&S.ocount, 1u);
                if (o < P.olist_cap) olist[o] = g | 0x80000000u; else atomicExch(P.err, 1);
                k0 = key;
            }
            if (k0 == key) { atomicAdd(&gval[g], v); pending = false; }
            else g = (g + 1) & P.gs_mask;
        }
    }
    if (pending) atomicExch(P.err, 2);
}

__device__ __forceinline__ void read_entry(const SrShared &S, const uint32_t *gkeys, const unsigned long long *gval,
                                           uint32_t o, uint32_t &id, unsigned long long &sc) {
    if (o & 0x80000000u) {       // tier-2 words are updated by atomics in L2: never trust L1
        uint32_t s = o & 0x7FFFFFFFu;
        id = __ldcg(gkeys + s);
        sc = __ldcg(gval + s);
    } else {
        id = S.keys[o];
        sc = ((unsigned long long)S.hi[o] << 32) | S.lo[o];
    }
}

// The 16-byte neighbour entry of one walk step.  asm volatile pins the ISSUE point: the loads of a
// step go out before the accumulator code that overlaps their latency.
__device__ __forceinline__ int4 ld_nbr4(const int4 *ptr) {
    int4 v;
    asm volatile("ld.global.nc" GW_LD_PREFETCH ".v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
    return v;
}

// ILP samples per thread (samples g*ILP .. g*ILP+ILP-1), walked in lock step: 2*STEP uniform steps
// from v with the whole paths in registers (SingleRandomWalk.java:53-72), ILP independent random
// loads in flight per thread.  The first-meet contribution of level i (computePathSim, :81-92)
// is emitted as soon as path[2i] is known, AFTER the loads of step 2i have been issued, so the
// accumulator work hides under the memory latency of the next step.  emit(ok, target, x) is
// called STEP*ILP times by every lane (lock-step), x = C^i * deg(path[i]) / deg(path[2i]) / SAMPLE.
// Sample s draws word (t & 3) of Philox(qid, s, t >> 2) at step t: paths do not depend on ILP,
// block size or grid.
template <int STEP, int ILP, bool F64 = false, typename Emit>
__device__ __forceinline__ int walk_group(const SimrankParams &P, int32_t v, uint2 mv, uint64_t qid, int32_t g,
                                          Emit &&emit) {
    constexpr int LEN = 2 * STEP;
    int32_t path[ILP][LEN + 1];
    uint32_t dmid[ILP][STEP + 1];           // deg(path[i]), i <= STEP
    uint2 m[ILP];                           // row descriptor of the current vertex
    bool alive[ILP];
    uint4 r[ILP];
    uint32_t dith[ILP];                     // the latest step's random word: its low bits dither the fixed-point rounding
    int steps = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) {
        path[k][0] = v;
        dmid[k][0] = 0;
        m[k] = mv;
        alive[k] = g * ILP + k < P.sample;
        dith[k] = 0;
    }
    auto emit_level = [&](int i, int k) {
        const int32_t target = path[k][2 * i];          // -1 when the walk ended before 2i steps (:66)
        bool ok = target >= 0 && target != v;
#pragma unroll
        for (int j = 0; j < i; j++) ok &= (path[k][j] != path[k][2 * i - j]);   // isFirstMeet :100-106
        // m[k] still describes path[2i]: its degree is the divisor
        if constexpr (F64) {
            // SingleRandomWalk.java:89 in its own type and operation order: cache[i] * deg(inter) / deg(target) / SAMPLE
            emit(ok, (uint32_t)target, P.cpow64[i] * (double)dmid[k][i] / (double)max(m[k].y, 1u) / (double)P.sample);
        } else {
            // the increment in 32.32 fixed-point units, rounded STOCHASTICALLY: floor(x * 2^32 + u), u from 16 spare bits
            // of the sample's own Philox stream.  A plain round-to-nearest repeats the same error for every hit of the same
            // (level, degree, degree) triple -- up to 2e-5 of an increment at SAMPLE = 1e5, 4.6e-6 absolute on a score;
            // dithered, the error of a sum of n hits is ~0.3 sqrt(n) units (2e-8 for 1e5 hits), deterministic in the seed.
            const float x = __fdividef(P.coef[i] * (float)dmid[k][i], (float)max(m[k].y, 1u));
            emit(ok, (uint32_t)target, floorf(fmaf(x, 4294967296.0f, (float)(dith[k] & 0xFFFFu) * (1.0f / 65536.0f))));
        }
    };
#pragma unroll
    for (int t = 0; t < LEN; t++) {
        int4 e[ILP];
#pragma unroll
        for (int k = 0; k < ILP; k++) {
            if ((t & 3) == 0)
                r[k] = Philox::gen(make_uint4((uint32_t)qid, (uint32_t)(qid >> 32), (uint32_t)(g * ILP + k), (uint32_t)(t >> 2)), P.key);
            const uint32_t rw = (t & 3) == 0 ? r[k].x : (t & 3) == 1 ? r[k].y : (t & 3) == 2 ? r[k].z : r[k].w;
            dith[k] = rw;
            alive[k] = alive[k] && m[k].y != 0;        // Graph.randNeighbor == -1 on a dead end (Graph.java:69-73)
            e[k] = make_int4(-1, 0, 0, 0);
            if (alive[k]) e[k] = ld_nbr4(P.nbr4 + m[k].x + scale_u32(rw, m[k].y));
        }
        if (t >= 2 && (t & 1) == 0) {
#pragma unroll
            for (int k = 0; k < ILP; k++) emit_level(t >> 1, k);
        }
#pragma unroll
        for (int k = 0; k < ILP; k++) {
            path[k][t + 1] = e[k].x;
            if (alive[k]) {
                m[k] = make_uint2((uint32_t)e[k].z, (uint32_t)e[k].w);   // next row descriptor rides in the same 16 bytes
                steps++;
            }
            if (t + 1 <= STEP) dmid[k][t + 1] = (uint32_t)e[k].w;
        }
    }
#pragma unroll
    for (int k = 0; k < ILP; k++) emit_level(STEP, k);
    return steps;
}

// walk_group emits increments already in fixed-point units (integer-valued floats, stochastically rounded)
__device__ __forceinline__ unsigned long long to_fixed(float x) { return __float2ull_rz(x); }

// largest bin b with count(bins >= b) >= K (0 when fewer than K entries); warp 0 only
__device__ __forceinline__ uint32_t threshold_bin(const uint32_t *hist, uint32_t K, int lane) {
    uint32_t base = (31 - lane) * 32;      // lane 0 owns the top 32 bins
    uint32_t cnt = 0;
    for (int j = 0; j < 32; j++) cnt += hist[base + j];
    uint32_t incl = cnt;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    uint32_t before = incl - cnt;          // entries in bins above this lane's range
    uint32_t found = 0xFFFFFFFFu;
    if (before < K && incl >= K) {
        uint32_t c = before;
        for (int j = 31; j >= 0; j--) { c += hist[base + j]; if (c >= K) { found = base + j; break; } }
    }
    uint32_t any = __ballot_sync(0xffffffffu, found != 0xFFFFFFFFu);
    uint32_t thr = 0;
    if (any) thr = __shfl_sync(0xffffffffu, found, __ffs(any) - 1);
    return thr;
}

// rank-by-counting among the candidates in shared memory; writes the K best (score desc, id asc)
template <typename Sh>
__device__ __forceinline__ void emit_ranked(const Sh &S, uint32_t C, uint32_t K, int32_t *oid, double *osc, int tid, double scale) {
    for (uint32_t a = tid; a < C; a += SR_BLOCK) {
        unsigned long long sa = S.cand_score[a];
        uint32_t ia = S.cand_id[a];
        uint32_t rank = 0;
        for (uint32_t b = 0; b < C; b++) rank += better(S.cand_score[b], S.cand_id[b], sa, ia) ? 1u : 0u;
        if (rank < K) { oid[rank] = (int32_t)ia; osc[rank] = (double)sa * scale; }
    }
    for (uint32_t r = C + tid; r < K; r += SR_BLOCK) { oid[r] = -1; osc[r] = 0.0; }
}

// Phases B (dense row / top-k) and C (clear the touched slots) of the hash accumulator, shared by
// the Monte-Carlo hash kernel and the hybrid path-tree kernel.
__device__ __noinline__ void finish_query(SrShared &S, const SimrankParams &P, uint32_t *gkeys, unsigned long long *gval,
                                          uint32_t *olist, int64_t qi, int tid) {
    __syncthreads();
    const uint32_t M = min(S.ocount, P.olist_cap);

    if (P.out_dense) {
        // ---------------- dense row (getResult()) ----------------
        double *row = P.out_dense + (size_t)qi * (size_t)P.n;
        for (uint32_t e = tid; e < M; e += SR_BLOCK) {
            uint32_t id; unsigned long long sc;
            read_entry(S, gkeys, gval, olist[e], id, sc);
            row[id] = (double)sc * P.out_scale;
        }
    }
    if (P.out_ids) {
        // ---------------- phase B: top-k ----------------
        const uint32_t K = (uint32_t)P.k;
        for (int i = tid; i < SR_BINS; i += SR_BLOCK) S.hist[i] = 0;
        __syncthreads();
        // histogram over tier-1 entries only: their k-th largest score is a lower bound of
        // the final k-th largest
        uint32_t n_t1 = 0;
        for (uint32_t e = tid; e < M; e += SR_BLOCK) {
            uint32_t o = olist[e];
            if (!(o & 0x80000000u)) {
                unsigned long long sc = ((unsigned long long)S.hi[o] << 32) | S.lo[o];
                atomicAdd(&S.hist[score_bin(sc)], 1u);
                n_t1++;
            }
        }
        __syncthreads();
        if (tid < 32) {
            uint32_t thr = threshold_bin(S.hist, K, tid);
            if (tid == 0) { S.thr_bin = thr; S.ccount = 0; }
        }
        __syncthreads();
        const uint32_t thr = S.thr_bin;
        for (uint32_t e = tid; e < M; e += SR_BLOCK) {
            uint32_t id; unsigned long long sc;
            read_entry(S, gkeys, gval, olist[e], id, sc);
            if (sc != 0 && score_bin(sc) >= thr) {
                uint32_t c = atomicAdd(&S.ccount, 1u);
                if (c < SR_CAND) { S.cand_score[c] = sc; S.cand_id[c] = id; }
            }
        }
        __syncthreads();
        const uint32_t C = S.ccount;
        int32_t *oid = P.out_ids + (size_t)qi * K;
        double *osc = P.out_scores + (size_t)qi * K;
        if (C <= SR_CAND) {
            emit_ranked(S, C, K, oid, osc, tid, P.out_scale);
        } else {
            // fallback (mass ties at the threshold): K rounds of block arg-max over all entries
            unsigned long long last_s = ~0ull;
            uint32_t last_i = 0;
            bool first = true;
            for (uint32_t r = 0; r < K; r++) {
                unsigned long long bs = 0; uint32_t bi = SR_EMPTY;
                for (uint32_t e = tid; e < M; e += SR_BLOCK) {
                    uint32_t id; unsigned long long sc;
                    read_entry(S, gkeys, gval, olist[e], id, sc);
                    if (sc == 0) continue;
                    if (!first && !better(last_s, last_i, sc, id)) continue;   // already emitted
                    if (bi == SR_EMPTY || better(sc, id, bs, bi)) { bs = sc; bi = id; }
                }
                for (int o = 16; o; o >>= 1) {
                    unsigned long long s2 = __shfl_xor_sync(0xffffffffu, bs, o);
                    uint32_t i2 = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (i2 != SR_EMPTY && (bi == SR_EMPTY || better(s2, i2, bs, bi))) { bs = s2; bi = i2; }
                }
                if ((tid & 31) == 0) { S.red_score[tid >> 5] = bs; S.red_id[tid >> 5] = bi; }
                __syncthreads();
                if (tid == 0) {
                    unsigned long long fs = 0; uint32_t fi = SR_EMPTY;
                    for (int w = 0; w < SR_BLOCK / 32; w++) {
                        uint32_t i2 = S.red_id[w];
                        if (i2 != SR_EMPTY && (fi == SR_EMPTY || better(S.red_score[w], i2, fs, fi))) { fs = S.red_score[w]; fi = i2; }
                    }
                    S.sel_score = fs; S.sel_id = fi;
                    if (fi != SR_EMPTY) { oid[r] = (int32_t)fi; osc[r] = (double)fs * P.out_scale; }
                    else { oid[r] = -1; osc[r] = 0.0; }
                }
                __syncthreads();
                last_s = S.sel_score; last_i = S.sel_id; first = false;
                if (last_i == SR_EMPTY) {   // exhausted: pad the rest
                    for (uint32_t r2 = r + 1 + tid; r2 < K; r2 += SR_BLOCK) { oid[r2] = -1; osc[r2] = 0.0; }
                    break;
                }
            }
        }
    }
    __syncthreads();
    // ---------------- phase C: clear only the touched slots ----------------
    for (uint32_t e = tid; e < M; e += SR_BLOCK) {
        uint32_t o = olist[e];
        if (o & 0x80000000u) { uint32_t s = o & 0x7FFFFFFFu; gkeys[s] = SR_EMPTY; gval[s] = 0ull; }
        else { S.keys[o] = SR_EMPTY; S.lo[o] = 0; S.hi[o] = 0; }
    }
    if (tid == 0) { S.ocount = 0; S.ccount = 0; }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// hash kernel: two-tier hash accumulator (tier 2 in global memory).  Exact for every input;
// used for dense rows and as the slow path of the log kernel below.
// ---------------------------------------------------------------------------------------------
template <int STEP>
__global__ void __launch_bounds__(SR_BLOCK, 2) k_simrank_mc(SimrankParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SrShared &S = *reinterpret_cast<SrShared *>(smem_raw);
    constexpr int LEN = 2 * STEP;
    const int tid = threadIdx.x;
    const size_t gs = (size_t)P.gs_mask + 1;
    uint32_t *gkeys = P.gkeys + blockIdx.x * gs;
    unsigned long long *gval = P.gval + blockIdx.x * gs;
    uint32_t *olist = P.olist + (size_t)blockIdx.x * P.olist_cap;

    for (int i = tid; i < SR_HS; i += SR_BLOCK) { S.keys[i] = SR_EMPTY; S.lo[i] = 0; S.hi[i] = 0; }
    if (tid == 0) { S.ocount = 0; S.ccount = 0; }
    __syncthreads();
    unsigned long long my_steps = 0;

    // slow-path launch: the queries the log kernel could not finish exactly (P.qlist / P.qcount)
    const int64_t n_work = P.qlist ? (int64_t)*P.qcount : P.nq;
    for (int64_t wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
        const int64_t qi = P.qlist ? (int64_t)P.qlist[wi] : wi;
        const int32_t v = (int32_t)P.queries[qi];
        const uint64_t qid = P.query_id_base + (uint64_t)qi;

        // ---------------- phase A: walk + first-meet accumulation ----------------
        // warp-uniform trip count: every lane runs every round, lanes past SAMPLE are masked
        const uint2 mv = __ldg(P.meta + v);
        const int32_t ngroups = (P.sample + SR_ILP - 1) / SR_ILP;
        for (int32_t g0 = tid - (tid & 31); g0 < ngroups; g0 += SR_BLOCK) {
            my_steps += (unsigned long long)walk_group<STEP, SR_ILP>(P, v, mv, qid, g0 + (tid & 31),
                [&](bool ok, uint32_t target, float x) {
                    acc_add_warp(S, P, gkeys, gval, olist, ok, target, to_fixed(x));
                });
        }
        finish_query(S, P, gkeys, gval, olist, qi, tid);
    }
    for (int o = 16; o; o >>= 1) my_steps += __shfl_xor_sync(0xffffffffu, my_steps, o);
    if ((tid & 31) == 0 && my_steps && !P.qlist) atomicAdd(P.steps, my_steps);   // handed-over queries were counted by the log kernel
}

// ---------------------------------------------------------------------------------------------
// arithmetic reference (GW_SIMRANK_MC_F64): the SAME Philox walks as the two production kernels, every
// increment evaluated and added in fp64 as SingleRandomWalk.java:89 does, straight into the dense row.
// Exists to bound what fp32 increments + 32.32 fixed-point accumulation cost (tests: <= 1e-6 absolute).
// ---------------------------------------------------------------------------------------------
template <int STEP>
__global__ void __launch_bounds__(256) k_simrank_f64(SimrankParams P) {
    unsigned long long my_steps = 0;
    for (int64_t qi = blockIdx.x; qi < P.nq; qi += gridDim.x) {
        const int32_t v = (int32_t)P.queries[qi];
        const uint64_t qid = P.query_id_base + (uint64_t)qi;
        const uint2 mv = __ldg(P.meta + v);
        double *row = P.out_dense + (size_t)qi * (size_t)P.n;
        for (int32_t g = threadIdx.x; g < P.sample; g += blockDim.x)
            my_steps += (unsigned long long)walk_group<STEP, 1, true>(P, v, mv, qid, g, [&](bool ok, uint32_t target, double x) {
                if (ok) atomicAdd(row + target, x);
            });
    }
    for (int o = 16; o; o >>= 1) my_steps += __shfl_xor_sync(0xffffffffu, my_steps, o);
    if ((threadIdx.x & 31) == 0 && my_steps) atomicAdd(P.steps, my_steps);
}

// ---------------------------------------------------------------------------------------------
// log kernel (production): no global atomics.  A contribution goes to the shared-memory table when
// its key fits there; otherwise it is APPENDED to a per-CTA log (coalesced 8-byte stores) and its
// value is added to a small shared-memory sketch cell (an upper bound of every logged key's total).
// Top-k: threshold bin from the table (a lower bound of the k-th score), table candidates, then ONE
// streaming pass over the log that keeps only entries whose sketch cell could reach the threshold
// (almost none: logged keys are the single-hit tail) and sums those exactly in a tiny third table.
// A query that cannot be finished exactly this way (threshold at the single-hit level, sketch
// saturation, too many survivors) is flagged and re-run by the hash kernel; both kernels add the
// same 32.32 fixed-point integers, so the result does not depend on which one produced it.
// ---------------------------------------------------------------------------------------------
#ifndef SR_SKETCH_CELLS
#define SR_SKETCH_CELLS 16384
#endif
constexpr int SR_SKETCH = SR_SKETCH_CELLS;
#ifndef SR_RING_STAGES
#define SR_RING_STAGES 2
#endif
#ifndef SR_WILP
#define SR_WILP 1                            // samples walked in lock step per walker thread of the log kernel
#endif
#ifndef SR_T3_SLOTS
#define SR_T3_SLOTS 512
#endif
constexpr int SR_T3 = SR_T3_SLOTS;
constexpr int SR_LCAND = 256;

struct SrLogShared {
    uint32_t keys[SR_HS];
    uint32_t lo[SR_HS];                  // 0.32 fixed point; a carry out of it sends the query to the hash kernel
    uint32_t sketch[SR_SKETCH];          // 2^-24 units, rounded up
    uint32_t hist[SR_BINS];              // threshold histogram, then tier-3 keys (SR_T3 <= SR_BINS)
    uint32_t t3lo[SR_T3], t3hi[SR_T3];
    unsigned long long cand_score[SR_LCAND];
    uint32_t cand_id[SR_LCAND];
    uint32_t lcount;                     // log entries
    uint32_t ccount;                     // candidates
    uint32_t t3count;
    uint32_t thr_bin;
    uint32_t slow;                       // this query needs the hash kernel
};

// walker -> accumulator hand-over: one single-producer single-consumer ring per warp pair, a stage =
// the WILP * STEP (key, x) contributions of WILP samples per lane; full/empty mbarriers per stage
constexpr int SR_PAIRS = 16;                 // walker warps = accumulator warps per CTA (1024 threads, one CTA per SM)
constexpr int SR_ABLOCK = SR_PAIRS * 32;     // accumulator threads (phases B and C run on them alone)
static_assert(SR_ABLOCK == SR_BLOCK, "phase B/C strides assume SR_BLOCK accumulator threads");
template <int STEP>
struct SrRing {
    static constexpr int WILP = STEP <= 5 ? SR_WILP : 1;     // samples in flight per walker thread
    static constexpr int STAGES = STEP <= 5 ? SR_RING_STAGES : SR_RING_STAGES / 2;   // walkers run this far ahead of phase B
    uint2 slot[SR_PAIRS][STAGES][WILP * STEP * 32];
    unsigned long long full[SR_PAIRS][STAGES];
    unsigned long long empty[SR_PAIRS][STAGES];
};
// Called by ONE thread of a starting CTA.  Why many CTAs per SM slot instead of one persistent CTA with a static share of the
// queries: the SMs are not equal for random access -- equal chains of dependent random loads take 1.54 ms on 12 SMs, 2.29 ms on 96
// and 3.06 ms on 40 (tools/gather_bench.cu mode -5, profiles/r2_gather_waves.txt) -- so equal shares end with the slowest class
// (39.6 G loads/s) while small CTAs handed out by the hardware keep every SM busy (51 G/s).  k_simrank_log: 464 k -> 548 k
// queries/s with 16 CTAs per SM slot, results bit-identical.
__device__ __forceinline__ uint32_t take_slot(const SimrankParams &P) {
    uint32_t sl = blockIdx.x % P.nslots;
    while (atomicCAS(P.slots + sl, 0u, 1u) != 0u) sl = sl + 1 == P.nslots ? 0 : sl + 1;
    __threadfence();
    return sl;
}
__device__ __forceinline__ void give_slot(const SimrankParams &P, uint32_t sl) {   // after the CTA's last access to the slot's scratch
    __threadfence();
    atomicExch(P.slots + sl, 0u);
}

__device__ __forceinline__ void acc_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(SR_ABLOCK) : "memory"); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {      // release.cta: my stores before it are visible to the waiter
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra LAB_WAIT;\nDONE:\n}"
                 ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

// The STEP contributions of one sample per lane into the query's accumulator, batched so that the
// shared-memory round trips of the levels overlap instead of chaining: all bucket loads first, then
// all adds (return values consumed at the very end), ONE log reservation per warp for every level.
// 4-key buckets: one 16-byte shared load sees every slot a key may live in.  Slots of a bucket fill
// in order and never empty during a query, so "first empty slot of my view + CAS" places a key
// exactly once (a stale view only makes the CAS return the occupant, after which the bucket is
// re-read).  What does not fit goes to the per-CTA log + sketch.  Called by all 32 lanes together.
template <int STEP>
__device__ __forceinline__ void log_insert_chunk(SrLogShared &S, const SimrankParams &P, uint2 *log, int lane,
                                                 const uint32_t *key, const uint32_t *v32) {
    uint4 kk[STEP];
    uint32_t b0[STEP];
    bool pending[STEP];
    uint32_t wrapped = 0;
#pragma unroll
    for (int i = 0; i < STEP; i++) {
        b0[i] = (hash32(key[i]) & (SR_HS / 4 - 1)) * 4;
        pending[i] = key[i] != SR_EMPTY;
        kk[i] = make_uint4(0, 0, 0, 0);
        if (pending[i])
            asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(kk[i].x), "=r"(kk[i].y), "=r"(kk[i].z), "=r"(kk[i].w)
                         : "r"((uint32_t)__cvta_generic_to_shared(&S.keys[b0[i]])));
    }
    uint32_t old[STEP];
#pragma unroll
    for (int i = 0; i < STEP; i++) {
        old[i] = 0;
        if (pending[i]) {
            uint4 v = kk[i];
            for (int attempt = 0; attempt < 5; attempt++) {
                int j = v.x == key[i] ? 0 : v.y == key[i] ? 1 : v.z == key[i] ? 2 : v.w == key[i] ? 3 : -1;
                if (j < 0) {
                    int e = v.x == SR_EMPTY ? 0 : v.y == SR_EMPTY ? 1 : v.z == SR_EMPTY ? 2 : v.w == SR_EMPTY ? 3 : -1;
                    if (e < 0) break;                           // bucket full of other keys: log
                    uint32_t k0 = atomicCAS(&S.keys[b0[i] + e], SR_EMPTY, key[i]);
                    if (k0 == SR_EMPTY || k0 == key[i]) j = e;
                }
                if (j >= 0) {
                    old[i] = atomicAdd(&S.lo[b0[i] + j], v32[i]);
                    pending[i] = false;
                    break;
                }
                asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                             : "r"((uint32_t)__cvta_generic_to_shared(&S.keys[b0[i]])));
            }
        }
    }
    // overflow: append to the log, one shared-memory atomic per warp for all levels
    uint32_t mask[STEP], total = 0;
#pragma unroll
    for (int i = 0; i < STEP; i++) { mask[i] = __ballot_sync(0xffffffffu, pending[i]); total += __popc(mask[i]); }
    if (total) {
        uint32_t basep = 0;
        if (lane == 0) basep = atomicAdd(&S.lcount, total);
        basep = __shfl_sync(0xffffffffu, basep, 0);
        uint32_t sold[STEP], sadd[STEP];
#pragma unroll
        for (int i = 0; i < STEP; i++) {
            sold[i] = 0; sadd[i] = 0;
            if (pending[i]) {
                const uint32_t pos = basep + __popc(mask[i] & ((1u << lane) - 1));
                if (pos < P.log_cap) log[pos] = make_uint2(key[i], v32[i]);
                sadd[i] = (v32[i] >> 8) + 1u;
                sold[i] = atomicAdd(&S.sketch[(hash32(key[i]) >> 13) & (SR_SKETCH - 1)], sadd[i]);
            }
            basep += __popc(mask[i]);
        }
#pragma unroll
        for (int i = 0; i < STEP; i++) if (sold[i] + sadd[i] < sold[i]) wrapped = 1;    // sketch cell wrapped
    }
#pragma unroll
    for (int i = 0; i < STEP; i++) if (old[i] + v32[i] < old[i]) wrapped = 1;   // score >= 1.0: exact path
    if (wrapped) S.slow = 1;
}

// key[i] == SR_EMPTY: no contribution at level i.  fx: 32.32 fixed point; a single contribution >= 1.0 sends the
// query to the exact kernel.
template <int STEP>
__device__ __forceinline__ void log_insert_batch(SrLogShared &S, const SimrankParams &P, uint2 *log, int lane,
                                                 const uint32_t (&key)[STEP], const unsigned long long (&fx)[STEP]) {
    uint32_t v32[STEP];
    bool big = false;
#pragma unroll
    for (int i = 0; i < STEP; i++) { v32[i] = (uint32_t)fx[i]; big |= (key[i] != SR_EMPTY) && (fx[i] >> 32) != 0; }
    if (big) S.slow = 1;
    constexpr int C0 = STEP <= 5 ? STEP : 5;            // at most 5 levels in flight (registers)
    log_insert_chunk<C0>(S, P, log, lane, key, v32);
    if (STEP > C0) log_insert_chunk<(STEP > C0 ? STEP - C0 : 1)>(S, P, log, lane, key + C0, v32 + C0);
}

// Phases B (top-k) and C (reset) of the log-structured accumulator, run by SR_BLOCK threads (atid = 0..SR_BLOCK-1)
// that synchronise through bar(): the accumulator warps of k_simrank_log (named barrier) or the whole CTA of the
// path-tree kernel (__syncthreads).  The caller has made every insert of the query visible (bar()) before.
template <typename Bar>
__device__ __forceinline__ void log_finish_query(SrLogShared &S, const SimrankParams &P, uint2 *log, int64_t qi, int atid, Bar bar) {
    uint32_t *t3keys = S.hist;
        const uint32_t Lc = min(S.lcount, P.log_cap);
        if (S.lcount > P.log_cap && atid == 0) S.slow = 1;

        // ---------------- phase B: top-k ----------------
        const uint32_t K = (uint32_t)P.k;
        {
        for (int i = atid; i < SR_BINS; i += SR_BLOCK) S.hist[i] = 0;
        bar();
        for (int i = atid; i < SR_HS; i += SR_BLOCK) {
            if (S.keys[i] != SR_EMPTY) {
                unsigned long long sc = S.lo[i];
                if (sc) atomicAdd(&S.hist[score_bin(sc)], 1u);
            }
        }
        bar();
        if (atid < 32) {
            uint32_t thr = threshold_bin(S.hist, K, atid);
            if (atid == 0) { S.thr_bin = thr; S.ccount = 0; S.t3count = 0; }
        }
        bar();
        const uint32_t thr = S.thr_bin;
        if (thr == 0 && Lc > 0 && atid == 0) S.slow = 1;   // no usable lower bound: every logged key could matter
        for (int i = atid; i < SR_T3; i += SR_BLOCK) { t3keys[i] = SR_EMPTY; S.t3lo[i] = 0; S.t3hi[i] = 0; }
        for (int i = atid; i < SR_HS; i += SR_BLOCK) {     // table candidates
            uint32_t id = S.keys[i];
            if (id != SR_EMPTY) {
                unsigned long long sc = S.lo[i];
                if (sc != 0 && score_bin(sc) >= thr) {
                    uint32_t c = atomicAdd(&S.ccount, 1u);
                    if (c < SR_LCAND) { S.cand_score[c] = sc; S.cand_id[c] = id; }
                }
            }
        }
        bar();
        if (thr != 0) {
            // one streaming pass over the log; survivors are summed exactly in tier 3
            // (8 independent loads in flight per thread: the pass is L2-latency bound otherwise)
            for (uint32_t e0 = atid; e0 < Lc; e0 += SR_BLOCK * 8) {
              uint2 ens[8];
#pragma unroll
              for (int u = 0; u < 8; u++) {
                  const uint32_t e = e0 + u * SR_BLOCK;
                  ens[u] = e < Lc ? __ldcg(log + e) : make_uint2(SR_EMPTY, 0u);
              }
#pragma unroll
              for (int u = 0; u < 8; u++) {
                const uint2 en = ens[u];
                if (en.x == SR_EMPTY) continue;
                uint32_t cell = S.sketch[(hash32(en.x) >> 13) & (SR_SKETCH - 1)];
                if (score_bin((unsigned long long)cell << 8) >= thr) {
                    uint32_t slot = (hash32(en.x) >> 4) & (SR_T3 - 1);
                    bool done = false;
                    for (int pr = 0; pr < SR_T3 && !done; pr++) {
                        uint32_t k0 = ((volatile uint32_t *)t3keys)[slot];
                        if (k0 == SR_EMPTY) {
                            k0 = atomicCAS(&t3keys[slot], SR_EMPTY, en.x);
                            if (k0 == SR_EMPTY) {
                                if (atomicAdd(&S.t3count, 1u) >= (uint32_t)(SR_T3 * 3 / 4)) S.slow = 1;
                                k0 = en.x;
                            }
                        }
                        if (k0 == en.x) { fixed_add(&S.t3lo[slot], &S.t3hi[slot], (unsigned long long)en.y); done = true; }
                        else slot = (slot + 1) & (SR_T3 - 1);
                    }
                    if (!done) S.slow = 1;
                }
              }
            }
            bar();
            for (int i = atid; i < SR_T3; i += SR_BLOCK) {
                uint32_t id = t3keys[i];
                if (id != SR_EMPTY) {
                    unsigned long long sc = ((unsigned long long)S.t3hi[i] << 32) | S.t3lo[i];
                    if (sc != 0 && score_bin(sc) >= thr) {
                        uint32_t c = atomicAdd(&S.ccount, 1u);
                        if (c < SR_LCAND) { S.cand_score[c] = sc; S.cand_id[c] = id; }
                    }
                }
            }
        }
        bar();
        const uint32_t C = S.ccount;
        int32_t *oid = P.out_ids + (size_t)qi * K;
        double *osc = P.out_scores + (size_t)qi * K;
        const bool slow = S.slow != 0 || C > SR_LCAND;
        if (!slow) {
            emit_ranked(S, C, K, oid, osc, atid, P.out_scale);
        } else if (atid == 0) {
            uint32_t w = atomicAdd(P.qcount, 1u);     // hand the query to the hash kernel
            P.qlist_out[w] = (int32_t)qi;
        }
        }
        bar();
        // ---------------- phase C: reset ----------------
        for (int i = atid; i < SR_HS; i += SR_BLOCK) { S.keys[i] = SR_EMPTY; S.lo[i] = 0; }
        for (int i = atid; i < SR_SKETCH; i += SR_BLOCK) S.sketch[i] = 0;
        if (atid == 0) { S.lcount = 0; S.ccount = 0; S.t3count = 0; S.slow = 0; }
        bar();
}

// One CTA of 1024 threads per SM, split by role.  Warps [0, SR_PAIRS) are WALKERS: they do nothing but
// walk (WILP samples in flight per thread) and hand the (key, x) contributions of every sample to
// their partner through a shared-memory ring.  Warps [SR_PAIRS, 2*SR_PAIRS) are ACCUMULATORS: they
// drain the rings into the query's hash table / log, and run top-k (phase B) and the reset (phase C)
// among themselves behind a named barrier.  The walkers never meet a CTA-wide barrier: while the
// accumulators rank query q the walkers are already walking query q+1, so the random loads keep the
// memory system at its ceiling all the time instead of alternating with the shared-memory work.
// (The unsplit kernel ran phase A as "10 dependent loads, then STEP inserts" in every warp, in step
// with every other warp: walk time and accumulate time added up, 6.1 ms where the walks alone take
// 3.9 ms -- profiles/README.md.)
#ifdef SR_PROFILE
#define SR_TICK(slot) do { if (prof_on) { long long t_ = clock64(); atomicAdd(P.prof + (slot), (unsigned long long)(t_ - t_last)); t_last = t_; } } while (0)
#else
#define SR_TICK(slot) do { } while (0)
#endif

template <int STEP>
__global__ void __launch_bounds__(2 * SR_ABLOCK, 1) k_simrank_log(SimrankParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SrLogShared &S = *reinterpret_cast<SrLogShared *>(smem_raw);
    using Ring = SrRing<STEP>;
    constexpr int WILP = Ring::WILP;
    Ring &R = *reinterpret_cast<Ring *>(smem_raw + ((sizeof(SrLogShared) + 15) & ~(size_t)15));
    const int tid = threadIdx.x, lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    __shared__ uint32_t s_slot;
    // Queries are CLAIMED one at a time from a global counter (P.work[2]) by walker warp 0 and published to the other 31 warps
    // through an 8-entry ring, ONE QUERY AHEAD of its own use (no warp ever waits for the claim): the SMs do not run at the
    // same rate (take_slot), and a static deal ends with the slowest.  Every warp stays within two queries of every other (a
    // walker is at most Ring::STAGES stages ahead of its accumulator, the accumulators meet at a barrier per query), so an
    // entry is never overwritten before its last reader has taken it.
    __shared__ int32_t s_qseq[8];
    __shared__ uint32_t s_qpub;
    if (tid == 0) { s_slot = take_slot(P); s_qpub = 0; }
    auto next_query = [&](uint32_t k) -> int64_t {              // the k-th query of this CTA, -1 = none left; called by whole warps
        int32_t q;
        if (wrp == 0) {
            q = 0;
            if (lane == 0) {
                for (uint32_t e = k == 0 ? 0u : k + 1; e <= k + 1; e++) {          // entries 0 and 1 at the start, then k + 1
                    const uint32_t c = atomicAdd(P.work + 2, 1u);
                    ((volatile int32_t *)s_qseq)[e & 7] = (int64_t)c < P.nq ? (int32_t)c : -1;
                    __threadfence_block();
                    *(volatile uint32_t *)&s_qpub = e + 1;
                }
                q = ((volatile int32_t *)s_qseq)[k & 7];
            }
            q = __shfl_sync(0xffffffffu, q, 0);
        } else {
            while (*(volatile uint32_t *)&s_qpub <= k) { }
            __threadfence_block();
            q = ((volatile int32_t *)s_qseq)[k & 7];
        }
        return (int64_t)q;
    };
    uint32_t *t3keys = S.hist;

    if (tid < SR_PAIRS * Ring::STAGES) {
        mbar_init(&R.full[tid / Ring::STAGES][tid % Ring::STAGES], 32);
        mbar_init(&R.empty[tid / Ring::STAGES][tid % Ring::STAGES], 32);
    }
    for (int i = tid; i < SR_HS; i += 2 * SR_ABLOCK) { S.keys[i] = SR_EMPTY; S.lo[i] = 0; }
    for (int i = tid; i < SR_SKETCH; i += 2 * SR_ABLOCK) S.sketch[i] = 0;
    if (tid == 0) { S.lcount = 0; S.ccount = 0; S.t3count = 0; S.slow = 0; }
    __syncthreads();
    uint32_t stage_no = 0;                   // stages handed over so far by this warp pair (same count on both sides)
    const int32_t ngroups = (P.sample + WILP - 1) / WILP;
#ifdef SR_PROFILE
    const bool prof_on = blockIdx.x == 0 && lane == 0 && (wrp == 0 || wrp == SR_PAIRS);
    long long t_last = clock64();
#endif

    if (wrp < SR_PAIRS) {
        // =============================== walkers ===============================
        unsigned long long my_steps = 0;
        for (uint32_t kq = 0;; kq++) {
            const int64_t qi = next_query(kq);
            if (qi < 0) break;
            const int32_t v = (int32_t)P.queries[qi];
            const uint64_t qid = P.query_id_base + (uint64_t)qi;
            const uint2 mv = __ldg(P.meta + v);
            for (int32_t g0 = wrp * 32; g0 < ngroups; g0 += SR_PAIRS * 32, stage_no++) {
                const uint32_t st = stage_no % Ring::STAGES, ph = (stage_no / Ring::STAGES) & 1;
                SR_TICK(0);                                    // walker: walking
                mbar_wait(&R.empty[wrp][st], ph ^ 1);          // the accumulator has taken the stage's previous content
                SR_TICK(1);                                    // walker: waiting for a free stage
                uint2 *slot = R.slot[wrp][st];
                int cnt = 0;                                   // emit order: level-major, sample-minor (walk_group)
                my_steps += (unsigned long long)walk_group<STEP, WILP>(P, v, mv, qid, g0 + lane,
                    [&](bool ok, uint32_t key, float x) {
                        slot[cnt * 32 + lane] = make_uint2(ok ? key : SR_EMPTY, __float_as_uint(x));
                        cnt++;
                    });
                mbar_arrive(&R.full[wrp][st]);
            }
        }
        for (int o = 16; o; o >>= 1) my_steps += __shfl_xor_sync(0xffffffffu, my_steps, o);
        if (lane == 0 && my_steps) atomicAdd(P.steps, my_steps);
        return;
    }

    // =============================== accumulators ===============================
    const int pw = wrp - SR_PAIRS, atid = tid - SR_ABLOCK;
    const uint32_t my_slot = s_slot;                           // only the accumulator warps touch the log
    uint2 *log = P.log + (size_t)my_slot * P.log_cap;
    for (uint32_t kq = 0;; kq++) {
        const int64_t qi = next_query(kq);
        if (qi < 0) break;
        // ---------------- phase A: drain my walker's ring ----------------
        for (int32_t g0 = pw * 32; g0 < ngroups; g0 += SR_PAIRS * 32, stage_no++) {
            const uint32_t st = stage_no % Ring::STAGES, ph = (stage_no / Ring::STAGES) & 1;
            SR_TICK(2);                                        // accumulator: inserting
            mbar_wait(&R.full[pw][st], ph);
            SR_TICK(3);                                        // accumulator: waiting for its walker
#pragma unroll
            for (int k = 0; k < WILP; k++) {
                uint32_t ek[STEP];
                unsigned long long ex[STEP];
#pragma unroll
                for (int i = 0; i < STEP; i++) {
                    const uint2 en = R.slot[pw][st][(i * WILP + k) * 32 + lane];
                    ek[i] = en.x; ex[i] = to_fixed(__uint_as_float(en.y));
                }
                if (k == WILP - 1) mbar_arrive(&R.empty[pw][st]);          // stage is in registers: give it back
                log_insert_batch<STEP>(S, P, log, lane, ek, ex);
            }
        }
        SR_TICK(2);
        acc_barrier();
        SR_TICK(4);                                            // waiting for the other accumulators' last stages
        log_finish_query(S, P, log, qi, atid, [] { acc_barrier(); });
        SR_TICK(10);                                           // reset
    }
    acc_barrier();                                             // only the accumulator warps touch the log
    if (atid == 0) give_slot(P, my_slot);
}

// ---------------------------------------------------------------------------------------------
// hybrid kernel: TopSim_singleSample.walk (simrank/TopSim_singleSample.java:62-203).  A weighted path
// tree per query: a path of weight w at a vertex of degree d is split into all d neighbours with
// weight w/d when w >= d (:99-125), otherwise into ceil(w) random neighbours with weight w/ceil(w)
// (:126-149); at every even level 2i each path adds w * C^i * deg(path[i]) / deg(path[2i]) to
// sim[source][path[2i]] when it is a first meeting (:167-203; scores stay x SAMPLE as in the reference).
//
// Once a path has been SAMPLED its children carry weight <= 1 and have exactly one child per level
// from then on: the tree is an enumerated prefix (few, heavy paths) with independent chains hanging
// off it.  Phase 1 expands the prefix level-synchronously (structure-of-arrays double buffer in global
// memory, children allocated by a block scan, one thread per CHILD so that a hub does not serialise a
// lane) and hands every path that must sample to a chain-parent list; phase 2 gives every chain to one
// thread that copies the parent's history into registers ONCE and walks to depth 2*STEP like the
// Monte-Carlo walker (one random 16-byte nbr4 entry per step).  Before the split every level copied
// every path's whole history through HBM (28 % of the stall samples, 67 GB of DRAM traffic per 2048
// queries) and paid a dozen CTA barriers per level for ~10 000 single-child paths.
// paths(l) <= 1 + l*SAMPLE (weights are conserved and a path has at most w+1 children) sizes the buffers.
// Chain RNG: Philox keyed by (query id, parent's index in its level buffer, level, child number): every
// quantity is produced by deterministic scans, so results do not depend on scheduling.
// ---------------------------------------------------------------------------------------------
struct HybridParams {
    int32_t *vbuf;        // [grid][2][LEN+1][cap]   level buffers of the enumerated prefix
    double *wbuf;         // [grid][2][cap]
    uint2 *dbuf;          // [grid][2][cap]          row descriptor {offset, degree} of each path's last vertex (it arrives with the nbr4 entry)
    int32_t *chist;       // [grid][LEN+1][cap]      chain parents: history up to their level
    double *cw;           // [grid][cap]             weight of each of the parent's chains = w / ceil(w)
    uint32_t *cnum;       // [grid][cap]             number of chains = ceil(w)
    uint32_t *cofs;       // [grid][cap + 1]         exclusive prefix of cnum
    uint32_t *eofs;       // [grid][cap]             expansion of a level: children per path, then their exclusive prefix
    uint32_t *eoff;       // [grid][cap]             expansion of a level: row offset of an enumerating path's last vertex
    double *ecw;          // [grid][cap]             expansion of a level: weight each child of an enumerating path receives
    uint2 *cpar;          // [grid][2*cap]           chain -> {parent, child number}
    uint4 *crec;          // [grid][cap]             parent record {key, offset, degree of its last vertex, -}: ONE 16-byte load
    uint32_t hrow;        // ints per history slot (2*STEP+1 rounded up to a multiple of 4: read as 16-byte pieces)
    uint32_t cap;
    double cpow[16];      // C^i
};

template <bool LOGACC>
struct HyShared {
    typename std::conditional<LOGACC, SrLogShared, SrShared>::type acc;
    uint32_t ofs[SR_BLOCK + 1];          // children before each thread's run of paths (expansion), + total
    uint32_t warp_tot[SR_BLOCK / 32];
    uint32_t n_in, n_cp, n_chain;
    uint32_t slot;                       // scratch slot of this CTA (take_slot)
    uint32_t next_q;                     // the query this CTA fetched from the work counter
};

// block exclusive scan of one value per thread; returns the exclusive prefix, *total = block sum
template <typename HY>
__device__ __forceinline__ uint32_t block_scan(HY &Y, uint32_t val, int tid, uint32_t *total) {
    const int lane = tid & 31, wrp = tid >> 5;
    uint32_t incl = val;
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) Y.warp_tot[wrp] = incl;
    __syncthreads();
    if (wrp == 0) {
        uint32_t t = lane < SR_BLOCK / 32 ? Y.warp_tot[lane] : 0, inc2 = t;
        for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, inc2, o); if (lane >= o) inc2 += u; }
        if (lane < SR_BLOCK / 32) Y.warp_tot[lane] = inc2 - t;
        if (lane == SR_BLOCK / 32 - 1) Y.ofs[SR_BLOCK] = inc2;
    }
    __syncthreads();
    const uint32_t excl = Y.warp_tot[wrp] + incl - val;
    *total = Y.ofs[SR_BLOCK];
    __syncthreads();
    return excl;
}

// LOGACC = true: contributions go to the log-structured accumulator of the Monte-Carlo kernel (shared-memory table +
// per-CTA log + sketch, no global atomics); a query it cannot finish exactly is handed to the LOGACC = false
// instantiation (exact two-tier hash, also the dense-rows path).  Both add the same integers
// round(val / SAMPLE * 2^32), so results do not depend on which one ran; scores leave x SAMPLE (P.out_scale).
// CILP chains per thread walked in lock step (chain ids t_first + k * t_stride), then computePathSim for the levels each
// chain added (TopSim_singleSample.java:167-203).  emit(k, i, ok, target, fx) is called for every k < CILP and i = 1..STEP by
// every lane (dead lanes emit ok = false), fx = round(weight * C^i * deg(path[i]) / deg(path[2i]) / SAMPLE * 2^32).
// Set-up is ONE dependent step after cpar[t]: key, offset, weight, row descriptor and the whole history slot of the
// parent are independent loads (the history is read unconditionally and masked by the level).
template <int STEP, int CILP, typename Emit>
__device__ __forceinline__ int hy_walk_chains(const SimrankParams &P, const HybridParams &H, const uint2 *cpar, const uint4 *crec,
                                              const double *cw, const int32_t *chist,
                                              uint64_t qid, int32_t v, uint32_t t_first, uint32_t t_stride, uint32_t n_chain, Emit &&emit) {
    constexpr int LEN = 2 * STEP, HROW = (LEN + 1 + 3) & ~3;
    int32_t path[CILP][LEN + 1];
    uint32_t dgs[CILP][LEN + 1];
    int lvl[CILP], len[CILP];
    double wq[CILP];
    uint32_t ctr_p[CILP], ctr_lj[CILP];
    uint2 m[CILP];
    bool live[CILP], alive[CILP];
    uint4 r[CILP];
    int steps = 0;
#pragma unroll
    for (int k = 0; k < CILP; k++) {
        const uint32_t t = t_first + (uint32_t)k * t_stride;
        live[k] = t < n_chain;
        lvl[k] = LEN; len[k] = 0; wq[k] = 0.0; ctr_p[k] = 0; ctr_lj[k] = 0; m[k] = make_uint2(0, 0); r[k] = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int pos = 0; pos <= LEN; pos++) { path[k][pos] = -1; dgs[k][pos] = 0; }
        if (live[k]) {
            const uint2 cp = cpar[t];                                   // {parent, level << 24 | child number}: 8 bytes, coalesced over t
            const int lv = (int)(cp.y >> 24);
            const uint4 rec = crec[cp.x];
            const int4 *hrow = reinterpret_cast<const int4 *>(chist + (size_t)cp.x * HROW);
            wq[k] = cw[cp.x];
            m[k] = make_uint2(rec.y, rec.z);
            int32_t hv[HROW];
#pragma unroll
            for (int c4 = 0; c4 < HROW / 4; c4++) { const int4 q4 = 4 * c4 <= lv ? hrow[c4] : make_int4(-1, -1, -1, -1); hv[4 * c4] = q4.x; hv[4 * c4 + 1] = q4.y; hv[4 * c4 + 2] = q4.z; hv[4 * c4 + 3] = q4.w; }
#pragma unroll
            for (int pos = 0; pos <= LEN; pos++) path[k][pos] = hv[pos];
            lvl[k] = lv;
            ctr_p[k] = rec.x >> 5;
            ctr_lj[k] = cp.y;                                          // (level << 24) | child number
#pragma unroll
            for (int pos = 0; pos <= LEN; pos++) if (pos > lvl[k]) path[k][pos] = -1;
            len[k] = lvl[k];
        }
        alive[k] = live[k];
    }
#pragma unroll
    for (int sidx = 0; sidx < LEN; sidx++) {
        int4 e[CILP];
        bool go[CILP];
#pragma unroll
        for (int k = 0; k < CILP; k++) {
            go[k] = false;
            if (alive[k] && sidx >= lvl[k]) {
                const int off = sidx - lvl[k];
                if ((off & 3) == 0)
                    r[k] = Philox::gen(make_uint4((uint32_t)qid ^ (0x9E3779B9u * (uint32_t)(off >> 2)), (uint32_t)(qid >> 32), ctr_p[k], ctr_lj[k]), P.key);
                const uint32_t rw = (off & 3) == 0 ? r[k].x : (off & 3) == 1 ? r[k].y : (off & 3) == 2 ? r[k].z : r[k].w;
                if (m[k].y == 0) alive[k] = false;                        // randNeighbor == -1: the chain ends
                else { e[k] = ld_nbr4(P.nbr4 + m[k].x + scale_u32(rw, m[k].y)); go[k] = true; }
            }
        }
#pragma unroll
        for (int k = 0; k < CILP; k++)
            if (go[k]) {
                path[k][sidx + 1] = e[k].x;
                dgs[k][sidx + 1] = (uint32_t)e[k].w;
                m[k] = make_uint2((uint32_t)e[k].z, (uint32_t)e[k].w);
                len[k] = sidx + 1;
                steps++;
            }
    }
#pragma unroll
    for (int i = 1; i <= STEP; i++) {
#pragma unroll
        for (int k = 0; k < CILP; k++) {
            const int32_t target = path[k][2 * i];
            bool ok = live[k] && 2 * i > lvl[k] && 2 * i <= len[k] && target != v;
#pragma unroll
            for (int j = 0; j < i; j++) ok &= (path[k][j] != path[k][2 * i - j]);
            unsigned long long fx = 0;
            if (ok) {
                const uint32_t dmid = i > lvl[k] ? dgs[k][i] : __ldg(P.meta + path[k][i]).y;
                const double val = wq[k] * H.cpow[i] * (double)dmid / (double)dgs[k][2 * i];
                fx = __double2ull_rn(val * P.inv_sample * SR_FIX);
            }
            emit(k, i, ok, (uint32_t)target, fx);
        }
    }
    return steps;
}

// walker -> accumulator rings of the chain phase (log-structured instantiation): warps [0, 8) walk chains, warps [8, 16)
// insert their contributions -- the split k_simrank_log uses, for the same reason (profiles/README.md R2-6: 81 % of the
// kernel is the chain phase, and inside it every warp alternated between ~7 memory latencies and five insertions)
#ifndef HY_CILP
#define HY_CILP 2                    // chains walked in lock step per thread of the path-tree kernel's chain phase
#endif
template <int STEP>
struct HyRing {
    static constexpr int CILP = STEP <= 5 ? HY_CILP : 1;
    static constexpr int PAIRS = SR_BLOCK / 64, STAGES = 2;
    uint2 slot[PAIRS][STAGES][CILP * STEP * 32];
    unsigned long long full[PAIRS][STAGES];
    unsigned long long empty[PAIRS][STAGES];
};

#ifdef HY_PROFILE
#define HY_TICK(slot) do { if (blockIdx.x == 0 && threadIdx.x == 0) { long long t_ = clock64(); atomicAdd(P.prof + (slot), (unsigned long long)(t_ - hy_last)); hy_last = t_; } } while (0)
#else
#define HY_TICK(slot) do { } while (0)
#endif
#ifdef HY_PROFILE
#define HY_WTICK(slot) do { if (blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == SR_BLOCK / 2)) { long long t_ = clock64(); atomicAdd(P.prof + (slot), (unsigned long long)(t_ - hy_wlast)); hy_wlast = t_; } } while (0)
#else
#define HY_WTICK(slot) do { } while (0)
#endif
template <int STEP, bool LOGACC, bool SPLIT = false>
__global__ void __launch_bounds__(SR_BLOCK, 1) k_topsim_hybrid(SimrankParams P, HybridParams H) {
#ifdef HY_PROFILE
    long long hy_last = clock64(), hy_wlast = clock64();
#endif
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HyShared<LOGACC> &Y = *reinterpret_cast<HyShared<LOGACC> *>(smem_raw);
    auto &S = Y.acc;
    HyRing<STEP> &R = *reinterpret_cast<HyRing<STEP> *>(smem_raw + ((sizeof(HyShared<LOGACC>) + 15) & ~(size_t)15));   // LOGACC only
    uint32_t stage_no = 0;                                     // ring stages produced / consumed by this warp so far
    if constexpr (LOGACC && SPLIT) {
        if (threadIdx.x < HyRing<STEP>::PAIRS * HyRing<STEP>::STAGES) {
            mbar_init(&R.full[threadIdx.x / HyRing<STEP>::STAGES][threadIdx.x % HyRing<STEP>::STAGES], 32);
            mbar_init(&R.empty[threadIdx.x / HyRing<STEP>::STAGES][threadIdx.x % HyRing<STEP>::STAGES], 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x == 0) Y.slot = take_slot(P);
    __syncthreads();
    const size_t sl = Y.slot;
    uint2 *log = P.log + (size_t)sl * P.log_cap;
    constexpr int LEN = 2 * STEP;
    const int tid = threadIdx.x, lane = tid & 31;
    const size_t gs = (size_t)P.gs_mask + 1;
    uint32_t *gkeys = P.gkeys + sl * gs;
    unsigned long long *gval = P.gval + sl * gs;
    uint32_t *olist = P.olist + (size_t)sl * P.olist_cap;
    const size_t cap = H.cap;
    int32_t *vb = H.vbuf + (size_t)sl * 2 * (LEN + 1) * cap;
    double *wb = H.wbuf + (size_t)sl * 2 * cap;
    uint2 *db = H.dbuf + (size_t)sl * 2 * cap;
    int32_t *chist = H.chist + (size_t)sl * (size_t)((LEN + 1 + 3) & ~3) * cap;     // history slots of (2*STEP+1 rounded up to 4) ints
    double *cw = H.cw + (size_t)sl * cap;
    uint32_t *cnum = H.cnum + (size_t)sl * cap;
    uint32_t *cofs = H.cofs + (size_t)sl * (cap + 1);
    uint32_t *eofs = H.eofs + (size_t)sl * cap;
    uint32_t *eoff = H.eoff + (size_t)sl * cap;
    double *ecw = H.ecw + (size_t)sl * cap;
    uint2 *cpar = H.cpar + (size_t)sl * 2 * cap;
    uint4 *crec = H.crec + (size_t)sl * cap;
    constexpr int HROW = (LEN + 1 + 3) & ~3;

    if constexpr (LOGACC) {
        for (int i = tid; i < SR_HS; i += SR_BLOCK) { S.keys[i] = SR_EMPTY; S.lo[i] = 0; }
        for (int i = tid; i < SR_SKETCH; i += SR_BLOCK) S.sketch[i] = 0;
        if (tid == 0) { S.lcount = 0; S.ccount = 0; S.t3count = 0; S.slow = 0; }
    } else {
        for (int i = tid; i < SR_HS; i += SR_BLOCK) { S.keys[i] = SR_EMPTY; S.lo[i] = 0; S.hi[i] = 0; }
        if (tid == 0) { S.ocount = 0; S.ccount = 0; }
    }
    __syncthreads();
    unsigned long long my_steps = 0;
    // one contribution per lane, all 32 lanes together
    auto add = [&](bool ok, uint32_t key, unsigned long long fx) {
        if constexpr (LOGACC) {
            const uint32_t k1[1] = {ok ? key : SR_EMPTY};
            const unsigned long long f1[1] = {fx};
            log_insert_batch<1>(S, P, log, lane, k1, f1);
        } else {
            acc_add_warp(S, P, gkeys, gval, olist, ok, key, fx);
        }
    };

    // exact instantiation as the slow path: only the queries the log instantiation handed over (P.qlist / P.qcount)
    const int64_t n_work = P.qlist ? (int64_t)*P.qcount : P.nq;
    // Queries are CLAIMED from a global counter, not dealt out by stride: a path tree costs between a few thousand and a
    // few hundred thousand steps depending on the hubs near its root, and with a static deal the kernel ends when the
    // unluckiest of 148 CTAs does (measured: +10 % on BA-10M).  Results are keyed by the query index, not by the CTA.
    for (uint32_t done = 0; P.quota == 0 || done < P.quota; done++) {
        if (tid == 0) Y.next_q = atomicAdd(P.work + (LOGACC ? 0 : 1), 1u);
        __syncthreads();
        const int64_t wi = (int64_t)Y.next_q;
        if (wi >= n_work) break;
        const int64_t qi = P.qlist ? (int64_t)P.qlist[wi] : wi;
        const int32_t v = (int32_t)P.queries[qi];
        const uint64_t qid = P.query_id_base + (uint64_t)qi;
        if (tid == 0) { vb[0] = v; wb[0] = (double)P.sample; db[0] = __ldg(P.meta + v); Y.n_in = 1; Y.n_cp = 0; Y.n_chain = 0; }
        __syncthreads();
        HY_TICK(3);                                                // (previous query's top-k / reset ends here)
        // ======================= phase 1: the enumerated prefix, level by level =======================
        for (int l = 0; l <= LEN; l++) {
            const int b = l & 1;
            const int32_t *vin = vb + (size_t)b * (LEN + 1) * cap;
            const double *win = wb + (size_t)b * cap;
            int32_t *vout = vb + (size_t)(b ^ 1) * (LEN + 1) * cap;
            double *wout = wb + (size_t)(b ^ 1) * cap;
            const uint2 *din = db + (size_t)b * cap;
            uint2 *dout = db + (size_t)(b ^ 1) * cap;
            const uint32_t n_in = Y.n_in;
            if (n_in == 0) break;                                  // every path has been handed to the chains (uniform)
            // ---- computePathSim at even levels (i = l/2), :80-83 and :157 ----
            if (l >= 2 && (l & 1) == 0) {
                const int i = l >> 1;
                for (uint32_t base = 0; base < n_in; base += SR_BLOCK) {
                    const uint32_t p = base + tid;
                    bool ok = p < n_in;
                    int32_t target = -1;
                    unsigned long long fx = 0;
                    if (ok) {
                        target = vin[(size_t)l * cap + p];
                        ok = target != v && target >= 0;                          // :184-185
                        for (int j = 0; j < i && ok; j++)                         // isFirstMeet :211-218
                            ok = vin[(size_t)j * cap + p] != vin[(size_t)(l - j) * cap + p];
                        if (ok) {
                            const int32_t inter = vin[(size_t)i * cap + p];
                            const double val = win[p] * H.cpow[i] * (double)__ldg(P.meta + inter).y /
                                               (double)__ldg(P.meta + target).y;   // :189
                            fx = __double2ull_rn(val * P.inv_sample * SR_FIX);
                        }
                    }
                    add(ok, (uint32_t)target, fx);
                }
            }
            HY_TICK(8);                                            // prefix: contributions of an even level
            if (l == LEN) break;
            // ---- expand level l -> l+1: enumerating paths stay in the level buffers, sampling paths become chain parents ----
            // Three passes over the WHOLE level with three barriers, whatever its size (a level of 4 000 paths used to take
            // eight rounds of 512 with two dependent DRAM latencies and seven barriers each).
            // pass 1: classify every path
            for (uint32_t base = 0; base < n_in; base += SR_BLOCK * 4) {
                double w[4];
                uint2 m[4];                                                       // no DRAM access here: the descriptors came with the entries
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t p = base + (uint32_t)u * SR_BLOCK + tid;
                    w[u] = p < n_in ? win[p] : 0.0;
                    m[u] = p < n_in ? din[p] : make_uint2(0u, 0u);
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t p = base + (uint32_t)u * SR_BLOCK + tid;
                    if (p >= n_in) continue;
                    uint32_t nchild = 0;
                    if (m[u].y != 0 && w[u] >= (double)m[u].y) {                  // :99-125 enumerate
                        nchild = m[u].y;
                        eoff[p] = m[u].x;
                        ecw[p] = w[u] / (double)m[u].y;
                    } else if (m[u].y != 0) {                                     // :126-149 sample ceil(w): a chain parent
                        const int number = ((double)(int)w[u] == w[u]) ? (int)w[u] : (int)w[u] + 1;
                        if (number > 0) {
                            const uint32_t k = atomicAdd(&Y.n_cp, 1u);
                            const uint32_t o = atomicAdd(&Y.n_chain, (uint32_t)number);   // its chains' ids (any order: draws are keyed by path, not by id)
                            if (k < cap) {
                                // parent-major history: a chain reads its parent's 2*STEP+1 slots as one contiguous piece
                                for (int pos = 0; pos <= l; pos++) chist[(size_t)k * HROW + pos] = vin[(size_t)pos * cap + p];
                                crec[k] = make_uint4((p << 5) | (uint32_t)l, m[u].x, m[u].y, (uint32_t)number);
                                cw[k] = w[u] / (double)number;
                                cnum[k] = (uint32_t)number;
                                cofs[k] = o;
                            }
                        }
                    }                                                             // degree 0: randNeighbor == -1, no child
                    eofs[p] = nchild;
                }
            }
            __syncthreads();
            HY_TICK(9);                                            // prefix: pass 1
            // pass 2: exclusive prefix of the child counts in path order -- every thread owns a contiguous run of paths, ONE block scan
            const uint32_t run_len = (n_in + SR_BLOCK - 1) / SR_BLOCK;
            const uint32_t p0 = min(n_in, (uint32_t)tid * run_len), p1 = min(n_in, p0 + run_len);
            uint32_t mine = 0;
            for (uint32_t p = p0; p < p1; p++) mine += eofs[p];
            uint32_t T;
            uint32_t run = block_scan(Y, mine, tid, &T);
            Y.ofs[tid] = run;
            for (uint32_t p = p0; p < p1; p++) { const uint32_t c = eofs[p]; eofs[p] = run; run += c; }
            __syncthreads();
            if ((uint64_t)T > cap) { if (tid == 0) { atomicExch(P.err, 3); Y.n_in = 0; } __syncthreads(); break; }
            HY_TICK(10);                                           // prefix: pass 2
            // pass 3: one thread per child; its path = the last one whose prefix is <= the child's index (threads' runs by
            // bisection in shared memory, then along the run)
            for (uint32_t c = tid; c < T; c += SR_BLOCK) {
                uint32_t lo2 = 0, hi2 = SR_BLOCK;                      // last thread t with ofs[t] <= c
                while (hi2 - lo2 > 1) { const uint32_t mid = (lo2 + hi2) >> 1; if (Y.ofs[mid] <= c) lo2 = mid; else hi2 = mid; }
                uint32_t parent = lo2 * run_len;
                const uint32_t pend = min(n_in, parent + run_len);
                while (parent + 1 < pend && eofs[parent + 1] <= c) parent++;
                const uint32_t j = c - eofs[parent];
                for (int pos = 0; pos <= l; pos++) vout[(size_t)pos * cap + c] = vin[(size_t)pos * cap + parent];
                const int4 e = ld_nbr4(P.nbr4 + eoff[parent] + j);                  // the level's ONE round trip to DRAM
                vout[(size_t)(l + 1) * cap + c] = e.x;
                dout[c] = make_uint2((uint32_t)e.z, (uint32_t)e.w);
                wout[c] = ecw[parent];
                my_steps++;
            }
            __syncthreads();
            HY_TICK(11);                                           // prefix: pass 3
            if (tid == 0) Y.n_in = T;
            __syncthreads();
        }
        __syncthreads();
        HY_TICK(0);                                                // prefix
        // ======================= phase 2: the chains =======================
        const uint32_t n_cp = Y.n_cp;
        if (n_cp > cap) { if (tid == 0) atomicExch(P.err, 3); }
        else if (n_cp > 0) {
            // chain -> parent map (the chains' ids were handed out when their parents were created)
            const uint32_t n_chain = Y.n_chain;
            if (n_chain > 2 * cap) { if (tid == 0) atomicExch(P.err, 3); }
            else {
                for (uint32_t k = tid; k < n_cp; k += SR_BLOCK) {
                    const uint32_t o = cofs[k], c = cnum[k];
                    const uint32_t lv24 = (crec[k].x & 31u) << 24;            // the level rides in the chain map: a chain then knows which
                    for (uint32_t j = 0; j < c; j++) cpar[o + j] = make_uint2(k, lv24 | j);   // pieces of its parent's history exist
                }
                __syncthreads();
                HY_TICK(1);                                        // chain set-up (scan, chain -> parent map)
                if constexpr (LOGACC && SPLIT) {
                    using Ring = HyRing<STEP>;
                    constexpr int CILP = Ring::CILP;
                    const int wrp = tid >> 5;
                    if (wrp < Ring::PAIRS) {                   // ---- walkers ----
                        bool overflow = false;
                        for (uint32_t g0 = (uint32_t)wrp * 32 * CILP; g0 < n_chain; g0 += Ring::PAIRS * 32 * CILP, stage_no++) {
                            const uint32_t st = stage_no % Ring::STAGES, ph = (stage_no / Ring::STAGES) & 1;
                            HY_WTICK(5);                                        // walker: walking
                            mbar_wait(&R.empty[wrp][st], ph ^ 1);
                            HY_WTICK(4);                                        // walker: waiting for a free stage
                            uint2 *slot = R.slot[wrp][st];
                            my_steps += (unsigned long long)hy_walk_chains<STEP, CILP>(P, H, cpar, crec, cw, chist, qid, v,
                                g0 + lane, 32u, n_chain, [&](int k, int i, bool ok, uint32_t key, unsigned long long fx) {
                                    if (ok && fx > 0xFFFFFFFFull) { overflow = true; fx = 0xFFFFFFFFull; }     // the exact instantiation redoes the query
                                    slot[((i - 1) * CILP + k) * 32 + lane] = make_uint2(ok ? key : SR_EMPTY, (uint32_t)fx);
                                });
                            mbar_arrive(&R.full[wrp][st]);
                        }
                        if (overflow) S.slow = 1;
                    } else {                                   // ---- accumulators ----
                        const int pw = wrp - Ring::PAIRS;
                        for (uint32_t g0 = (uint32_t)pw * 32 * CILP; g0 < n_chain; g0 += Ring::PAIRS * 32 * CILP, stage_no++) {
                            const uint32_t st = stage_no % Ring::STAGES, ph = (stage_no / Ring::STAGES) & 1;
                            HY_WTICK(7);                                        // accumulator: inserting
                            mbar_wait(&R.full[pw][st], ph);
                            HY_WTICK(6);                                        // accumulator: waiting for its walker
                            // all CILP * STEP contributions of the lane in ONE batch: the insert is a chain of dependent
                            // shared-memory round trips per warp, so its throughput is the number of entries in flight
                            uint32_t ek[CILP * STEP], ev[CILP * STEP];
#pragma unroll
                            for (int i = 0; i < CILP * STEP; i++) {
                                const uint2 en = R.slot[pw][st][i * 32 + lane];
                                ek[i] = en.x; ev[i] = en.y;
                            }
                            mbar_arrive(&R.empty[pw][st]);                                 // stage is in registers: give it back
                            log_insert_chunk<CILP * STEP>(S, P, log, lane, ek, ev);
                        }
                    }
                } else if constexpr (LOGACC) {
                    // every warp walks TWO chains per lane in lock step (1024 loads in flight per SM: the chain phase is a
                    // stream of true DRAM misses and 512 in flight do not reach the access ceiling), then inserts the
                    // 2 * STEP contributions of the lane as one batch
                    constexpr int CILP = HyRing<STEP>::CILP;
                    bool overflow = false;
                    for (uint32_t g0 = (uint32_t)(tid >> 5) * 32 * CILP; g0 < n_chain; g0 += SR_BLOCK * CILP) {
                        uint32_t ek[CILP][STEP], ev[CILP][STEP];
                        my_steps += (unsigned long long)hy_walk_chains<STEP, CILP>(P, H, cpar, crec, cw, chist, qid, v, g0 + lane, 32u,
                            n_chain, [&](int k, int i, bool ok, uint32_t key, unsigned long long fx) {
                                if (ok && fx > 0xFFFFFFFFull) { overflow = true; fx = 0xFFFFFFFFull; }
                                ek[k][i - 1] = ok ? key : SR_EMPTY;
                                ev[k][i - 1] = (uint32_t)fx;
                            });
                        if constexpr (CILP <= 2) {
                            log_insert_chunk<CILP * STEP>(S, P, log, lane, &ek[0][0], &ev[0][0]);
                        } else {
#pragma unroll
                            for (int k = 0; k < CILP; k++) log_insert_chunk<STEP>(S, P, log, lane, ek[k], ev[k]);
                        }
                    }
                    if (overflow) S.slow = 1;
                } else {
                    for (uint32_t t0 = (uint32_t)(tid - lane); t0 < n_chain; t0 += SR_BLOCK)
                        my_steps += (unsigned long long)hy_walk_chains<STEP, 1>(P, H, cpar, crec, cw, chist, qid, v, t0 + lane, 0u,
                            n_chain, [&](int, int, bool ok, uint32_t key, unsigned long long fx) { add(ok, key, fx); });
                }
            }
        }
        __syncthreads();
        HY_TICK(2);                                                // chains
        if constexpr (LOGACC) {
            __syncthreads();
            log_finish_query(S, P, log, qi, tid, [] { __syncthreads(); });
        } else {
            finish_query(S, P, gkeys, gval, olist, qi, tid);
        }
    }
    for (int o = 16; o; o >>= 1) my_steps += __shfl_xor_sync(0xffffffffu, my_steps, o);
    if (lane == 0 && my_steps && !P.qlist) atomicAdd(P.steps, my_steps);   // handed-over queries were counted by the log instantiation
    __syncthreads();
    if (tid == 0) give_slot(P, (uint32_t)sl);
}

// ---------------- replay mode: java.util.Random on the device ----------------
// java.util.Random (JDK): seed = (seed * 0x5DEECE66D + 0xB) mod 2^48, next(bits) = (int)(seed >>> (48 - bits));
// nextInt(bound): power of two -> (bound * next(31)) >> 31, else u % bound with the int-overflow rejection test.
// (jr_next / jr_next_int live in common.cuh: doublewalk.cu replays with them too)

// lxctools/FixedCacheMap.java:26-100 on the device, one instance per replayed query: 1-based binary min-heap on float
// values (keys/vals [nmax+1]) and the key -> heap slot map as a dense pos[n] array (0 = absent; the reference keeps a
// HashMap<Integer, Short>, same contents for nmax <= 32767, which the entry point enforces).
struct DevCacheMap {
    int32_t *keys; float *vals; int32_t *pos; int nmax, n;
    __device__ __forceinline__ bool greater(int a, int b) const { return vals[a] > vals[b]; }
    __device__ __forceinline__ void exch(int a, int b) {
        pos[keys[a]] = b; pos[keys[b]] = a;
        const int32_t tk = keys[a]; keys[a] = keys[b]; keys[b] = tk;
        const float tv = vals[a]; vals[a] = vals[b]; vals[b] = tv;
    }
    __device__ void sink(int i) {
        while (2 * i <= n) {
            int j = 2 * i;
            if (j < n && greater(j, j + 1)) j++;
            if (!greater(i, j)) break;
            exch(i, j);
            i = j;
        }
    }
    __device__ void swim(int i) {
        while (i > 1 && greater(i / 2, i)) { exch(i, i / 2); i = i / 2; }
    }
    __device__ void put(int32_t key, float value) {               // :32-50
        const int idx = pos[key];
        if (idx != 0) { vals[idx] = __fadd_rn(vals[idx], value); sink(idx); }
        else if (n < nmax) { n++; keys[n] = key; vals[n] = value; pos[key] = n; swim(n); }
        else if (value > vals[1]) { pos[keys[1]] = 0; keys[1] = key; vals[1] = value; pos[key] = 1; sink(1); }
    }
};
struct ReplaySink {          // where a replayed path's contribution goes: dense double rows or one DevCacheMap per query
    double *dense; int32_t *ckeys; float *cvals; int32_t *cpos; int32_t *csize; int32_t cap;
};

// One thread per query, samples in order, fp64 accumulation in the reference's operation order
// (SingleRandomWalk.java:89: cache[i] * deg(inter) / deg(target) / SAMPLE, left to right).
// CACHE: SingleRandomWalk_M.java:81-94 -- the same walks, the increment cast to float and put() into the query's cache.
template <bool CACHE>
__global__ void k_simrank_javarng(const uint2 *__restrict__ meta, const int32_t *__restrict__ col,
                                  const int64_t *__restrict__ queries, int64_t nq, int64_t n, int32_t sample, int32_t step,
                                  const double *__restrict__ cache, uint64_t *__restrict__ states, ReplaySink sk,
                                  unsigned long long *__restrict__ steps_out) {
    const int64_t qi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    const int32_t v = (int32_t)queries[qi];
    double *row = CACHE ? nullptr : sk.dense + (size_t)qi * (size_t)n;
    DevCacheMap cm{nullptr, nullptr, nullptr, 0, 0};
    if constexpr (CACHE)
        cm = DevCacheMap{sk.ckeys + (size_t)qi * (size_t)(sk.cap + 1), sk.cvals + (size_t)qi * (size_t)(sk.cap + 1),
                         sk.cpos + (size_t)qi * (size_t)n, sk.cap, 0};
    uint64_t seed = states[qi];
    const int max_step = 2 * step;
    int32_t path[21];
    unsigned long long steps = 0;
    for (int32_t s = 0; s < sample; s++) {
        int len = 0;
        path[0] = v;
        int32_t cur = v;
        while (len < max_step) {                                   // :64-68
            const uint2 m = meta[cur];
            if (m.y == 0) break;                                   // randNeighbor == -1
            cur = col[m.x + (uint32_t)jr_next_int(seed, (int32_t)m.y)];
            path[++len] = cur;
            steps++;
        }
        if (len == 0) continue;                                    // :82
        for (int i = 1; i <= step && 2 * i <= len; i++) {          // :84-91
            const int32_t inter = path[i], target = path[2 * i];
            if (target == v) continue;
            bool first = true;
            for (int j = 0; j < i; j++) first &= (path[j] != path[2 * i - j]);
            if (first) {
                const double x = __ddiv_rn(__ddiv_rn(__dmul_rn(cache[i], (double)meta[inter].y), (double)meta[target].y), (double)sample);
                if constexpr (CACHE) cm.put(target, __double2float_rn(x));
                else row[target] = __dadd_rn(row[target], x);
            }
        }
    }
    if constexpr (CACHE) sk.csize[qi] = cm.n;
    else row[v] = 0.0;                                             // :43
    states[qi] = seed;
    atomicAdd(steps_out, steps);
}

// Replay of the path-tree estimators: one thread per query walks the reference's FIFO queue level by level
// (TopSim_singleSample.java:62-158 / TopSim_Enumerate.java:61-130) in two global-memory level buffers:
// vb[level parity][position 0..2*STEP][cap] (structure of arrays), wb[level parity][cap].
// CACHE: TopSim_singleSample_M.java:224-225 -- increment / SAMPLE cast to float and put() into the query's cache.
template <bool CACHE>
__global__ void k_topsim_javarng(const uint2 *__restrict__ meta, const int32_t *__restrict__ col,
                                 const int64_t *__restrict__ queries, int64_t nq, int64_t n, int32_t sample, int32_t step,
                                 int32_t mode, const double *__restrict__ cache, int64_t cap, int32_t *__restrict__ vbuf,
                                 double *__restrict__ wbuf, uint64_t *__restrict__ states, ReplaySink sk,
                                 int *__restrict__ err) {
    const int64_t qi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    const int32_t v = (int32_t)queries[qi];
    const int max_step = 2 * step, LEN1 = max_step + 1;
    double *row = CACHE ? nullptr : sk.dense + (size_t)qi * (size_t)n;
    DevCacheMap cm{nullptr, nullptr, nullptr, 0, 0};
    if constexpr (CACHE)
        cm = DevCacheMap{sk.ckeys + (size_t)qi * (size_t)(sk.cap + 1), sk.cvals + (size_t)qi * (size_t)(sk.cap + 1),
                         sk.cpos + (size_t)qi * (size_t)n, sk.cap, 0};
    int32_t *vb = vbuf + (size_t)qi * 2 * LEN1 * cap;
    double *wb = wbuf + (size_t)qi * 2 * cap;
    uint64_t seed = states[qi];
    int64_t n0 = 1;
    for (int i = 0; i <= max_step; i++) vb[(size_t)i * cap] = -1;
    vb[0] = v; wb[0] = (double)sample;
    int path_len = 0, topsim = 1;
    bool overflow = false;
    for (;;) {
        const int b = path_len & 1;
        const int32_t *vin = vb + (size_t)b * LEN1 * cap;
        const double *win = wb + (size_t)b * cap;
        int32_t *vout = vb + (size_t)(b ^ 1) * LEN1 * cap;
        double *wout = wb + (size_t)(b ^ 1) * cap;
        const bool last = path_len >= max_step;
        if (last || path_len / 2 == topsim) {                       // computePathSim, :80-83 and :157
            const int start = topsim;
            if (path_len != 0) {
                for (int64_t k = 0; k < n0; k++) {
                    for (int i = start; i <= step && 2 * i <= path_len; i++) {          // :180-192
                        const int32_t inter = vin[(size_t)i * cap + k], target = vin[(size_t)(2 * i) * cap + k];
                        if (target == v || target == -1) continue;
                        bool first = true;
                        for (int j = 0; j < i; j++) first &= (vin[(size_t)j * cap + k] != vin[(size_t)(2 * i - j) * cap + k]);
                        if (first) {
                            const double x = __ddiv_rn(__dmul_rn(__dmul_rn(win[k], cache[i]), (double)meta[inter].y), (double)meta[target].y);
                            if constexpr (CACHE) cm.put(target, __double2float_rn(__ddiv_rn(x, (double)sample)));
                            else row[target] = __dadd_rn(row[target], x);                                // :189
                        }
                    }
                }
            }
            if (!last) topsim++;
        }
        if (last) break;
        int64_t n1 = 0;
        for (int64_t k = 0; k < n0 && !overflow; k++) {
            const int32_t c = vin[(size_t)path_len * cap + k];
            const double wt = win[k];
            const uint2 m = meta[c];
            const int d = (int)m.y;
            if (d != 0 && (mode == 1 || wt >= (double)d)) {          // :99-125 enumerate
                const double ns = __ddiv_rn(wt, (double)d);
                if (n1 + d > cap) { overflow = true; break; }
                for (int j = 0; j < d; j++) {
                    for (int pos = 0; pos <= path_len; pos++) vout[(size_t)pos * cap + n1] = vin[(size_t)pos * cap + k];
                    for (int pos = path_len + 2; pos <= max_step; pos++) vout[(size_t)pos * cap + n1] = -1;
                    vout[(size_t)(path_len + 1) * cap + n1] = col[m.x + j];
                    wout[n1] = ns;
                    n1++;
                }
            } else if (mode == 0) {                                  // :126-149 ceil(weight) random children
                const int number = ((double)(int)wt == wt) ? (int)wt : (int)wt + 1;
                for (int j = 0; j < number; j++) {
                    if (d == 0) break;                               // randNeighbor == -1
                    const int32_t nb = col[m.x + (uint32_t)jr_next_int(seed, d)];
                    if (n1 + 1 > cap) { overflow = true; break; }
                    for (int pos = 0; pos <= path_len; pos++) vout[(size_t)pos * cap + n1] = vin[(size_t)pos * cap + k];
                    for (int pos = path_len + 2; pos <= max_step; pos++) vout[(size_t)pos * cap + n1] = -1;
                    vout[(size_t)(path_len + 1) * cap + n1] = nb;
                    wout[n1] = __ddiv_rn(wt, (double)number);
                    n1++;
                }
            }
        }
        if (overflow) break;
        n0 = n1;
        path_len++;
    }
    if (overflow) atomicExch(err, 1);
    if constexpr (CACHE) sk.csize[qi] = cm.n;
    else row[v] = 0.0;
    states[qi] = seed;
}

// ---------------- exact SimRank (SimRank.java:36-77) as dense sweeps ----------------
// T[i][:] = mean over a in N(i) of S[a][:]   (rows of degree 0 -> 0)
__global__ void k_row_average(const uint2 *__restrict__ meta, const int32_t *__restrict__ col, int64_t n,
                              const double *__restrict__ Sm, double *__restrict__ T, double scale,
                              int pin_diag) {
    int64_t i = blockIdx.y;
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint2 m = meta[i];
    double acc = 0.0;
    for (uint32_t a = 0; a < m.y; a++) acc += Sm[(int64_t)col[m.x + a] * n + j];
    double r = m.y ? scale * acc / (double)m.y : 0.0;
    if (pin_diag && i == j) r = 1.0;
    T[i * n + j] = r;
}
__global__ void k_transpose(const double *__restrict__ A, double *__restrict__ B, int64_t n) {
    __shared__ double tile[32][33];
    int64_t x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y)
        if (x < n && y0 + r < n) tile[r][threadIdx.x] = A[(y0 + r) * n + x];
    __syncthreads();
    int64_t tx = blockIdx.y * 32 + threadIdx.x, ty0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y)
        if (tx < n && ty0 + r < n) B[(ty0 + r) * n + tx] = tile[threadIdx.x][r];
}
__global__ void k_identity(double *__restrict__ A, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) A[i * n + i] = 1.0;
}
__global__ void k_gather_rows_zero_diag(const double *__restrict__ A, int64_t n, const int64_t *__restrict__ rows,
                                        int64_t nrows, double *__restrict__ out) {
    int64_t r = blockIdx.y;
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n || r >= nrows) return;
    int64_t i = rows[r];
    out[r * n + j] = (i == j) ? 0.0 : A[i * n + j];
}

// A _dev call whose accumulators overflowed (err != 0) may leave entries in the hash tables, and nobody has to read
// the error before the next call is enqueued: the next call re-initialises the tables ON THE DEVICE when it finds the
// previous call's error word set (it runs before the header is cleared; costs one empty launch otherwise).
__global__ void k_sr_clean_if_err(const int *__restrict__ err, unsigned long long *__restrict__ gval, uint32_t *__restrict__ gkeys,
                                  size_t count) {
    if (*err == 0) return;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        gval[i] = 0ull;
        gkeys[i] = 0xFFFFFFFFu;
    }
}

}  // namespace gw

using namespace gw;

template <int STEP>
static int launch_kernels(const SimrankParams &P, int grid, int log_grid, bool use_log, cudaStream_t st) {
    if (use_log) {
        size_t smem = ((sizeof(SrLogShared) + 15) & ~(size_t)15) + sizeof(SrRing<STEP>);
        GW_CUDA(cudaFuncSetAttribute(k_simrank_log<STEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_simrank_log<STEP><<<log_grid, 2 * SR_ABLOCK, smem, st>>>(P);
        GW_LAUNCHED();
    }
    // hash kernel: everything (dense rows / forced) or only the queries the log kernel handed over
    SimrankParams Q = P;
    if (use_log) Q.qlist = P.qlist_out;
    size_t smem = sizeof(SrShared);
    GW_CUDA(cudaFuncSetAttribute(k_simrank_mc<STEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_simrank_mc<STEP><<<grid, SR_BLOCK, smem, st>>>(Q);
    GW_LAUNCHED();
    return GW_OK;
}

extern "C" int gw_simrank_check_args(const gw_graph *g, double c, int32_t step, int32_t sample, int32_t k, int32_t mode) {
    if (!g) return fail(GW_E_INVALID, "graph is NULL");
    if (g->flags & GW_F_DIRECTED) return fail(GW_E_INVALID, "SimRank path is defined on undirected graphs (structures/Graph.java)");
    if (step < 1 || step > 10) return fail(GW_E_INVALID, "step must be in 1..10");
    if (sample < 1) return fail(GW_E_INVALID, "sample must be positive");
    if (!(c > 0) || !(c < 1)) return fail(GW_E_INVALID, "decay c must be in (0,1)");
    if (k < 1 || k > SR_LCAND / 2) return fail(GW_E_INVALID, "k must be in 1..%d", SR_LCAND / 2);
    if (mode != GW_SIMRANK_MC && mode != GW_SIMRANK_HYBRID && mode != GW_SIMRANK_MC_F64) return fail(GW_E_INVALID, "unknown estimator mode %d", mode);
    return GW_OK;
}

static int simrank_run(gw_graph *g, const int64_t *d_queries, int64_t nq, double c, int32_t step, int32_t sample,
                       int32_t k, int32_t mode, uint64_t seed, uint64_t query_id_base, int32_t *d_out_ids,
                       double *d_out_scores, double *d_out_dense, cudaStream_t st, bool sync_steps) {
    GW_TRY(gw_simrank_check_args(g, c, step, sample, d_out_ids ? k : 1, mode));
    if (nq == 0) return GW_OK;
    if (nq >= ((int64_t)1 << 31)) return fail(GW_E_TOO_LARGE, "more than 2^31-1 queries in one call");
    GW_CUDA(cudaSetDevice(g->device));
    int sms = 0;
    GW_TRY(device_info(&sms, nullptr));
    const bool hybrid = mode == GW_SIMRANK_HYBRID;
    if (mode == GW_SIMRANK_MC_F64 && !d_out_dense) return fail(GW_E_INVALID, "GW_SIMRANK_MC_F64 produces dense rows only (gw_simrank_rows)");
    const int grid = (int)std::min<int64_t>(nq, (int64_t)sms * (hybrid ? 1 : 2));
    // log and path-tree kernels: one 1024- / 512-thread CTA is resident per SM, but the launch holds SR_WAVES CTAs per SM
    // (each with 1/SR_WAVES of the queries) -- see take_slot for the measurement behind it.  GW_SR_WAVES overrides (1 = persistent).
    int sr_waves = 16, log_waves = 1;                                               // the log kernel claims its queries one by one: persistent
    if (const char *wv = getenv("GW_SR_WAVES")) sr_waves = log_waves = std::max(1, atoi(wv));
    const int nslots = (int)std::min<int64_t>(nq, (int64_t)sms);                    // scratch areas = CTAs that can be resident
    const int log_grid = (int)std::min<int64_t>(nq, (int64_t)sms * log_waves);
    const char *force = getenv("GW_SIMRANK");
    const bool use_log = !hybrid && d_out_ids && !d_out_dense && !(force && !strcmp(force, "hash"));
    // hash-kernel tier-2 table: >= 2x the distinct targets one query can produce
    // (hybrid: every path of every even level may add a target, paths(l) <= 1 + l*SAMPLE)
    const int64_t distinct = std::min<int64_t>(hybrid ? (int64_t)sample * step * (step + 1) + step : (int64_t)sample * step, g->n);
    uint32_t gs = 1024;
    while ((int64_t)gs < 2 * distinct) gs <<= 1;
    const uint32_t ocap = (uint32_t)distinct + 1;
    // log capacity = the most contributions one query can make (path tree: every path of every even level)
    const uint32_t log_cap = (uint32_t)std::min<int64_t>(hybrid ? (int64_t)sample * step * (step + 1) + step : (int64_t)sample * step,
                                                         (int64_t)0x7FFFFFFF);
    // layout: [4 KB header: counters at 0..255, scratch-slot flags from 1024][gval u64 grid*gs][gkeys u32 grid*gs][olist u32 grid*ocap][log uint2 grid*log_cap][qlist i32 nq]
    size_t off_gval = 4096;
    size_t off_gkeys = off_gval + (size_t)grid * gs * 8;
    size_t off_olist = off_gkeys + (size_t)grid * gs * 4;
    size_t off_log = (off_olist + (size_t)grid * ocap * 4 + 15) & ~(size_t)15;
    size_t off_qlist = off_log + (size_t)std::max(grid, nslots) * log_cap * 8;
    size_t need = off_qlist + (size_t)nq * 4 + 16;
    const bool fresh = g->simrank_scratch_bytes < need || g->simrank_layout != (uint64_t)gs * 1000003u + (uint64_t)grid;
    if (g->simrank_scratch_bytes < need) {
        cudaFree(g->d_simrank_scratch);
        g->d_simrank_scratch = nullptr;
        g->simrank_scratch_bytes = 0;
        GW_CUDA(cudaMalloc(&g->d_simrank_scratch, need));
        g->simrank_scratch_bytes = need;
    }
    unsigned char *base = (unsigned char *)g->d_simrank_scratch;
    SimrankParams P;
    P.steps = (unsigned long long *)base;
    P.err = (int *)(base + 16);
    P.qcount = (uint32_t *)(base + 32);
    P.gval = (unsigned long long *)(base + off_gval);
    P.gkeys = (uint32_t *)(base + off_gkeys);
    P.olist = (uint32_t *)(base + off_olist);
    P.log = (uint2 *)(base + off_log);
    P.qlist_out = (int32_t *)(base + off_qlist);
    P.qlist = nullptr;
    P.log_cap = log_cap;
    if (!fresh && !g->simrank_dirty) {
        k_sr_clean_if_err<<<sms * 4, 256, 0, st>>>(P.err, P.gval, P.gkeys, (size_t)grid * gs);
        GW_LAUNCHED();
    }
    GW_CUDA(cudaMemsetAsync(base, 0, 4096, st));
    if (nslots > 768) return fail(GW_E_STATE, "%d scratch slots do not fit the header", nslots);
    P.slots = (uint32_t *)(base + 1024); P.nslots = (uint32_t)nslots; P.quota = 0;
    P.prof = (unsigned long long *)(base + 64);
    P.work = (uint32_t *)(base + 48);
    if (fresh || g->simrank_dirty) {   // the hash kernel leaves its tables clean; only (re)initialise when the layout changes
        GW_CUDA(cudaMemsetAsync(P.gval, 0, (size_t)grid * gs * 8, st));
        GW_CUDA(cudaMemsetAsync(P.gkeys, 0xFF, (size_t)grid * gs * 4, st));
        g->simrank_layout = (uint64_t)gs * 1000003u + (uint64_t)grid;
        g->simrank_dirty = 0;
    }
    GW_TRY(ensure_common_counts(g, st, false));               // nbr4 rows (no counts needed)
    P.meta = g->d_meta; P.col = g->d_col; P.nbr4 = g->d_nbr4; P.queries = d_queries; P.nq = nq; P.n = g->n;
    P.sample = sample; P.k = k;
    P.out_scale = hybrid ? (double)sample / SR_FIX : 1.0 / SR_FIX;
    P.inv_sample = 1.0 / (double)sample;
    for (int i = 0; i < 16; i++) P.coef[i] = 0;
    for (int i = 1; i <= step; i++) P.coef[i] = (float)(pow(c, i) / (double)sample);   // cache[i] = Math.pow(C, i) (:34-36), / SAMPLE (:89)
    for (int i = 0; i < 16; i++) P.cpow64[i] = (i >= 1 && i <= step) ? pow(c, i) : 0.0;
    P.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    P.query_id_base = query_id_base;
    P.gs_mask = gs - 1; P.olist_cap = ocap;
    P.out_ids = d_out_ids; P.out_scores = d_out_scores; P.out_dense = d_out_dense;
    if (hybrid) {
        HybridParams H;
        H.cap = (uint32_t)std::min<int64_t>((int64_t)2 * step * sample + 1, (int64_t)0x07FFFFFF);   // index << 5 must fit 32 bits
        const size_t LEN1 = 2 * (size_t)step + 1, capz = H.cap;
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
        const size_t o_w = take((size_t)grid * 2 * capz * sizeof(double));
        const size_t o_d = take((size_t)grid * 2 * capz * sizeof(uint2));
        const size_t o_cw = take((size_t)grid * capz * sizeof(double));
        const size_t o_v = take((size_t)grid * 2 * LEN1 * capz * sizeof(int32_t));
        const size_t HR = (LEN1 + 3) & ~(size_t)3;
        const size_t o_ch = take((size_t)grid * HR * capz * sizeof(int32_t));
        const size_t o_cn = take((size_t)grid * capz * sizeof(uint32_t));
        const size_t o_co = take((size_t)grid * (capz + 1) * sizeof(uint32_t));
        const size_t o_ck = take((size_t)grid * capz * sizeof(uint32_t));
        const size_t o_eo = take((size_t)grid * capz * sizeof(uint32_t));
        const size_t o_ew = take((size_t)grid * capz * sizeof(double));
        const size_t o_cp = take((size_t)grid * 2 * capz * sizeof(uint2));
        const size_t o_cm = take((size_t)grid * capz * sizeof(uint4));
        if (g->hybrid_scratch_bytes < off + 16) {
            cudaFree(g->d_hybrid_scratch);
            g->d_hybrid_scratch = nullptr; g->hybrid_scratch_bytes = 0;
            if (cudaMalloc(&g->d_hybrid_scratch, off + 16) != cudaSuccess) {
                cudaGetLastError();
                return fail(GW_E_TOO_LARGE, "hybrid estimator needs %zu bytes of path buffers", off);
            }
            g->hybrid_scratch_bytes = off + 16;
        }
        unsigned char *hb = (unsigned char *)g->d_hybrid_scratch;
        H.wbuf = (double *)(hb + o_w); H.dbuf = (uint2 *)(hb + o_d); H.cw = (double *)(hb + o_cw);
        H.vbuf = (int32_t *)(hb + o_v); H.chist = (int32_t *)(hb + o_ch);
        H.cnum = (uint32_t *)(hb + o_cn); H.cofs = (uint32_t *)(hb + o_co);
        H.eofs = (uint32_t *)(hb + o_ck); H.eoff = (uint32_t *)(hb + o_eo); H.ecw = (double *)(hb + o_ew); H.cpar = (uint2 *)(hb + o_cp); H.crec = (uint4 *)(hb + o_cm); H.hrow = (uint32_t)HR;
        for (int i = 0; i < 16; i++) H.cpow[i] = i <= step ? pow(c, i) : 0.0;
        // top-k: log-structured instantiation first, then the exact one over the queries it handed over;
        // dense rows: the exact instantiation alone
        const bool hy_log = d_out_ids && !d_out_dense && !(force && !strcmp(force, "hash"));
        SimrankParams Q = P;
        if (hy_log) Q.qlist = P.qlist_out;
        const char *hsp = getenv("GW_HY_SPLIT");               // experiment knob: "1" = walker / accumulator warps in the chain phase
        const bool hy_split = hsp && !strcmp(hsp, "1");
        // log instantiation: sr_waves CTAs per SM in the launch, each ends after its share of the queries (claimed dynamically)
        const int hy_grid = (int)std::min<int64_t>(nq, (int64_t)sms * sr_waves);
        SimrankParams PL = P;
        PL.quota = (uint32_t)((nq + hy_grid - 1) / hy_grid);
#define GW_HY(N) case N: \
            if (hy_log && hy_split) { \
                const size_t hsm = ((sizeof(HyShared<true>) + 15) & ~(size_t)15) + sizeof(HyRing<N>); \
                GW_CUDA(cudaFuncSetAttribute(k_topsim_hybrid<N, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsm)); \
                k_topsim_hybrid<N, true, true><<<hy_grid, SR_BLOCK, hsm, st>>>(PL, H); \
                GW_LAUNCHED(); \
            } else if (hy_log) { \
                GW_CUDA(cudaFuncSetAttribute(k_topsim_hybrid<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HyShared<true>))); \
                k_topsim_hybrid<N, true><<<hy_grid, SR_BLOCK, sizeof(HyShared<true>), st>>>(PL, H); \
                GW_LAUNCHED(); \
            } \
            GW_CUDA(cudaFuncSetAttribute(k_topsim_hybrid<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HyShared<false>))); \
            k_topsim_hybrid<N, false><<<grid, SR_BLOCK, sizeof(HyShared<false>), st>>>(Q, H); break;
        switch (step) { GW_HY(1) GW_HY(2) GW_HY(3) GW_HY(4) GW_HY(5) GW_HY(6) GW_HY(7) GW_HY(8) GW_HY(9) default: GW_HY(10) }
#undef GW_HY
        GW_LAUNCHED();
    } else if (mode == GW_SIMRANK_MC_F64) {
        const int fgrid = (int)std::min<int64_t>(nq, (int64_t)sms * 8);
#define GW_F64(N) case N: k_simrank_f64<N><<<fgrid, 256, 0, st>>>(P); break;
        switch (step) { GW_F64(1) GW_F64(2) GW_F64(3) GW_F64(4) GW_F64(5) GW_F64(6) GW_F64(7) GW_F64(8) GW_F64(9) default: GW_F64(10) }
#undef GW_F64
        GW_LAUNCHED();
    } else
    switch (step) {
        case 1: GW_TRY(launch_kernels<1>(P, grid, log_grid, use_log, st)); break;
        case 2: GW_TRY(launch_kernels<2>(P, grid, log_grid, use_log, st)); break;
        case 3: GW_TRY(launch_kernels<3>(P, grid, log_grid, use_log, st)); break;
        case 4: GW_TRY(launch_kernels<4>(P, grid, log_grid, use_log, st)); break;
        case 5: GW_TRY(launch_kernels<5>(P, grid, log_grid, use_log, st)); break;
        case 6: GW_TRY(launch_kernels<6>(P, grid, log_grid, use_log, st)); break;
        case 7: GW_TRY(launch_kernels<7>(P, grid, log_grid, use_log, st)); break;
        case 8: GW_TRY(launch_kernels<8>(P, grid, log_grid, use_log, st)); break;
        case 9: GW_TRY(launch_kernels<9>(P, grid, log_grid, use_log, st)); break;
        default: GW_TRY(launch_kernels<10>(P, grid, log_grid, use_log, st)); break;
    }
    if (sync_steps) {
        unsigned long long hs = 0;
        int herr = 0;
        uint32_t nslow = 0;
        GW_CUDA(cudaMemcpyAsync(&hs, P.steps, sizeof(hs), cudaMemcpyDeviceToHost, st));
        GW_CUDA(cudaMemcpyAsync(&herr, P.err, sizeof(herr), cudaMemcpyDeviceToHost, st));
        GW_CUDA(cudaMemcpyAsync(&nslow, P.qcount, sizeof(nslow), cudaMemcpyDeviceToHost, st));
        GW_CUDA(cudaStreamSynchronize(st));
        g->simrank_last_steps = (int64_t)hs;
        g->simrank_last_slow = (int64_t)nslow;
        if (herr) { g->simrank_dirty = 1; return fail(GW_E_STATE, "SimRank accumulator overflow (code %d)", herr); }
    }
    return GW_OK;
}

static int check_queries_host(const gw_graph *g, const int64_t *q, int64_t nq) {
    for (int64_t i = 0; i < nq; i++)
        if (q[i] < 0 || q[i] >= g->n) return fail(GW_E_KEY, "query vertex %lld is outside [0, %lld)", (long long)q[i], (long long)g->n);
    return GW_OK;
}

extern "C" {

int gw_simrank_topk_dev(gw_graph *g, const int64_t *d_queries, int64_t nq, double c, int32_t step, int32_t sample,
                        int32_t k, int32_t mode, uint64_t seed, uint64_t query_id_base, int32_t *d_out_ids,
                        double *d_out_scores, void *stream) {
    if (nq < 0 || (nq > 0 && (!d_queries || !d_out_ids || !d_out_scores))) return fail(GW_E_INVALID, "bad arguments");
    return simrank_run(g, d_queries, nq, c, step, sample, k, mode, seed, query_id_base, d_out_ids, d_out_scores, nullptr,
                       (cudaStream_t)stream, false);
}

int gw_simrank_topk(gw_graph *g, const int64_t *queries, int64_t nq, double c, int32_t step, int32_t sample, int32_t k,
                    int32_t mode, uint64_t seed, uint64_t query_id_base, int32_t *out_ids, double *out_scores) {
    if (!g) return fail(GW_E_INVALID, "graph is NULL");
    if (nq < 0 || (nq > 0 && (!queries || !out_ids || !out_scores))) return fail(GW_E_INVALID, "bad arguments");
    if (k < 1) return fail(GW_E_INVALID, "k must be positive");
    GW_TRY(check_queries_host(g, queries, nq));
    if (nq == 0) return GW_OK;
    GW_CUDA(cudaSetDevice(g->device));
    // one device block {queries | ids | scores} and one pinned block of the same layout, both kept in the
    // handle: repeated calls touch neither cudaMalloc nor the driver's pageable-copy staging
    const size_t bq = sizeof(int64_t) * (size_t)nq, bs = sizeof(double) * (size_t)nq * k, bi = sizeof(int32_t) * (size_t)nq * k;
    const size_t off_s = (bq + 255) & ~(size_t)255, off_i = (off_s + bs + 255) & ~(size_t)255, total = off_i + bi;
    if (g->ws_sr_dev_bytes < total) {
        cudaFree(g->ws_sr_dev); g->ws_sr_dev = nullptr; g->ws_sr_dev_bytes = 0;
        GW_CUDA(cudaMalloc(&g->ws_sr_dev, total));
        g->ws_sr_dev_bytes = total;
    }
    if (g->ws_sr_pin_bytes < total) {
        if (g->ws_sr_pin) cudaFreeHost(g->ws_sr_pin);
        g->ws_sr_pin = nullptr; g->ws_sr_pin_bytes = 0;
        GW_CUDA(cudaMallocHost(&g->ws_sr_pin, total));
        g->ws_sr_pin_bytes = total;
    }
    unsigned char *dv = (unsigned char *)g->ws_sr_dev, *pin = (unsigned char *)g->ws_sr_pin;
    memcpy(pin, queries, bq);
    GW_CUDA(cudaMemcpyAsync(dv, pin, bq, cudaMemcpyHostToDevice, nullptr));
    GW_TRY(simrank_run(g, (const int64_t *)dv, nq, c, step, sample, k, mode, seed, query_id_base, (int32_t *)(dv + off_i),
                       (double *)(dv + off_s), nullptr, nullptr, false));
    GW_CUDA(cudaMemcpyAsync(pin + off_s, dv + off_s, total - off_s, cudaMemcpyDeviceToHost, nullptr));
    unsigned long long hs = 0;
    int herr = 0;
    uint32_t nslow = 0;
    unsigned char *hdr = (unsigned char *)g->d_simrank_scratch;
    GW_CUDA(cudaMemcpyAsync(&hs, hdr, sizeof(hs), cudaMemcpyDeviceToHost, nullptr));
    GW_CUDA(cudaMemcpyAsync(&herr, hdr + 16, sizeof(herr), cudaMemcpyDeviceToHost, nullptr));
    GW_CUDA(cudaMemcpyAsync(&nslow, hdr + 32, sizeof(nslow), cudaMemcpyDeviceToHost, nullptr));
    GW_CUDA(cudaStreamSynchronize(nullptr));
    g->simrank_last_steps = (int64_t)hs;
    g->simrank_last_slow = (int64_t)nslow;
    if (herr) { g->simrank_dirty = 1; return fail(GW_E_STATE, "SimRank accumulator overflow (code %d)", herr); }
    memcpy(out_scores, pin + off_s, bs);
    memcpy(out_ids, pin + off_i, bi);
    return GW_OK;
}

int gw_simrank_rows(gw_graph *g, const int64_t *queries, int64_t nq, double c, int32_t step, int32_t sample,
                    int32_t mode, uint64_t seed, uint64_t query_id_base, double *out_dense) {
    if (!g) return fail(GW_E_INVALID, "graph is NULL");
    if (nq < 0 || (nq > 0 && (!queries || !out_dense))) return fail(GW_E_INVALID, "bad arguments");
    GW_TRY(check_queries_host(g, queries, nq));
    if (nq == 0) return GW_OK;
    GW_CUDA(cudaSetDevice(g->device));
    DevBuf<int64_t> dq;
    DevBuf<double> dd;
    GW_CUDA(dq.alloc((size_t)nq)); GW_CUDA(dd.alloc((size_t)nq * (size_t)g->n));
    GW_CUDA(cudaMemcpy(dq.p, queries, sizeof(int64_t) * (size_t)nq, cudaMemcpyHostToDevice));
    GW_CUDA(cudaMemset(dd.p, 0, sizeof(double) * (size_t)nq * (size_t)g->n));
    GW_TRY(simrank_run(g, dq.p, nq, c, step, sample, 0, mode, seed, query_id_base, nullptr, nullptr, dd.p, nullptr, true));
    GW_CUDA(cudaMemcpy(out_dense, dd.p, sizeof(double) * (size_t)nq * (size_t)g->n, cudaMemcpyDeviceToHost));
    return GW_OK;
}

int gw_simrank_rows_javarng(gw_graph *g, const int64_t *queries, int64_t nq, double c, int32_t step, int32_t sample,
                            uint64_t *rng_state, double *out_dense) {
    if (!g) return fail(GW_E_INVALID, "graph is NULL");
    if (g->flags & GW_F_DIRECTED) return fail(GW_E_INVALID, "SimRank path is defined on undirected graphs (structures/Graph.java)");
    if (nq < 0 || (nq > 0 && (!queries || !rng_state || !out_dense))) return fail(GW_E_INVALID, "bad arguments");
    if (step < 1 || step > 10) return fail(GW_E_INVALID, "step must be in 1..10");
    if (sample < 1) return fail(GW_E_INVALID, "sample must be positive");
    GW_TRY(check_queries_host(g, queries, nq));
    if (nq == 0) return GW_OK;
    GW_CUDA(cudaSetDevice(g->device));
    double cache[16] = {0};
    for (int i = 1; i <= step; i++) cache[i] = pow(c, i);          // Math.pow(C, i), :34-36
    DevBuf<int64_t> dq;
    DevBuf<double> dd, dc;
    DevBuf<unsigned long long> ds, dsteps;
    GW_CUDA(dq.alloc((size_t)nq)); GW_CUDA(dd.alloc((size_t)nq * (size_t)g->n)); GW_CUDA(dc.alloc(16));
    GW_CUDA(ds.alloc((size_t)nq)); GW_CUDA(dsteps.alloc(1));
    GW_CUDA(cudaMemcpy(dq.p, queries, sizeof(int64_t) * (size_t)nq, cudaMemcpyHostToDevice));
    GW_CUDA(cudaMemcpy(dc.p, cache, sizeof(cache), cudaMemcpyHostToDevice));
    GW_CUDA(cudaMemcpy(ds.p, rng_state, sizeof(uint64_t) * (size_t)nq, cudaMemcpyHostToDevice));
    GW_CUDA(cudaMemset(dd.p, 0, sizeof(double) * (size_t)nq * (size_t)g->n));
    GW_CUDA(cudaMemset(dsteps.p, 0, sizeof(unsigned long long)));
    k_simrank_javarng<false><<<(unsigned)((nq + 31) / 32), 32>>>(g->d_meta, g->d_col, dq.p, nq, g->n, sample, step, dc.p,
                                                               (uint64_t *)ds.p, ReplaySink{dd.p, nullptr, nullptr, nullptr, nullptr, 0}, dsteps.p);
    GW_LAUNCHED();
    unsigned long long hs = 0;
    GW_CUDA(cudaMemcpy(out_dense, dd.p, sizeof(double) * (size_t)nq * (size_t)g->n, cudaMemcpyDeviceToHost));
    GW_CUDA(cudaMemcpy(rng_state, ds.p, sizeof(uint64_t) * (size_t)nq, cudaMemcpyDeviceToHost));
    GW_CUDA(cudaMemcpy(&hs, dsteps.p, sizeof(hs), cudaMemcpyDeviceToHost));
    g->simrank_last_steps = (int64_t)hs;
    if (g->d_simrank_scratch) GW_CUDA(cudaMemcpy(g->d_simrank_scratch, &hs, sizeof(hs), cudaMemcpyHostToDevice));   // gw_simrank_last_steps reads it there
    return GW_OK;
}

int gw_topsim_rows_javarng(gw_graph *g, const int64_t *queries, int64_t nq, double c, int32_t step, int32_t sample,
                           int32_t mode, int64_t max_paths, uint64_t *rng_state, double *out_dense) {
    if (!g) return fail(GW_E_INVALID, "graph is NULL");
    if (g->flags & GW_F_DIRECTED) return fail(GW_E_INVALID, "SimRank path is defined on undirected graphs (structures/Graph.java)");
    if (nq < 0 || (nq > 0 && (!queries || !rng_state || !out_dense))) return fail(GW_E_INVALID, "bad arguments");
    if (step < 1 || step > 10) return fail(GW_E_INVALID, "step must be in 1..10");
    if (sample < 1 || max_paths < 1) return fail(GW_E_INVALID, "sample and max_paths must be positive");
    if (mode != 0 && mode != 1) return fail(GW_E_INVALID, "mode must be 0 (TopSim_singleSample) or 1 (TopSim_Enumerate)");
    GW_TRY(check_queries_host(g, queries, nq));
    if (nq == 0) return GW_OK;
    GW_CUDA(cudaSetDevice(g->device));
    double cache[16] = {0};
    for (int i = 1; i <= step; i++) cache[i] = pow(c, i);          // Math.pow(C, i)
    const size_t LEN1 = 2 * (size_t)step + 1;
    DevBuf<int64_t> dq;
    DevBuf<double> dd, dc, dw;
    DevBuf<int32_t> dv;
    DevBuf<unsigned long long> ds;
    DevBuf<int> derr;
    if (dv.alloc((size_t)nq * 2 * LEN1 * (size_t)max_paths) != cudaSuccess || dw.alloc((size_t)nq * 2 * (size_t)max_paths) != cudaSuccess) {
        cudaGetLastError();
        return fail(GW_E_TOO_LARGE, "path buffers for %lld queries x %lld paths do not fit", (long long)nq, (long long)max_paths);
    }
    GW_CUDA(dq.alloc((size_t)nq)); GW_CUDA(dd.alloc((size_t)nq * (size_t)g->n)); GW_CUDA(dc.alloc(16));
    GW_CUDA(ds.alloc((size_t)nq)); GW_CUDA(derr.alloc(1));
    GW_CUDA(cudaMemcpy(dq.p, queries, sizeof(int64_t) * (size_t)nq, cudaMemcpyHostToDevice));
    GW_CUDA(cudaMemcpy(dc.p, cache, sizeof(cache), cudaMemcpyHostToDevice));
    GW_CUDA(cudaMemcpy(ds.p, rng_state, sizeof(uint64_t) * (size_t)nq, cudaMemcpyHostToDevice));
    GW_CUDA(cudaMemset(dd.p, 0, sizeof(double) * (size_t)nq * (size_t)g->n));
    GW_CUDA(cudaMemset(derr.p, 0, sizeof(int)));
    k_topsim_javarng<false><<<(unsigned)((nq + 31) / 32), 32>>>(g->d_meta, g->d_col, dq.p, nq, g->n, sample, step, mode, dc.p, max_paths, dv.p,
                                                              dw.p, (uint64_t *)ds.p, ReplaySink{dd.p, nullptr, nullptr, nullptr, nullptr, 0}, derr.p);
    GW_LAUNCHED();
    int herr = 0;
    GW_CUDA(cudaMemcpy(&herr, derr.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (herr) return fail(GW_E_TOO_LARGE, "a level of the path tree exceeded max_paths = %lld", (long long)max_paths);
    GW_CUDA(cudaMemcpy(out_dense, dd.p, sizeof(double) * (size_t)nq * (size_t)g->n, cudaMemcpyDeviceToHost));
    GW_CUDA(cudaMemcpy(rng_state, ds.p, sizeof(uint64_t) * (size_t)nq, cudaMemcpyDeviceToHost));
    return GW_OK;
}

int gw_simrank_cache_javarng(gw_graph *g, const int64_t *queries, int64_t nq, double c, int32_t step, int32_t sample,
                             int32_t mode, int32_t capacity, int64_t max_paths, uint64_t *rng_state, int32_t *out_keys,
                             float *out_vals, int32_t *out_sizes) {
    if (!g) return fail(GW_E_INVALID, "graph is NULL");
    if (g->flags & GW_F_DIRECTED) return fail(GW_E_INVALID, "SimRank path is defined on undirected graphs (structures/Graph.java)");
    if (nq < 0 || (nq > 0 && (!queries || !rng_state || !out_keys || !out_vals || !out_sizes))) return fail(GW_E_INVALID, "bad arguments");
    if (step < 1 || step > 10) return fail(GW_E_INVALID, "step must be in 1..10");
    if (sample < 1) return fail(GW_E_INVALID, "sample must be positive");
    if (mode != 0 && mode != 1) return fail(GW_E_INVALID, "mode must be 0 (SingleRandomWalk_M) or 1 (TopSim_singleSample_M)");
    if (capacity < 1 || capacity > 32767)
        return fail(GW_E_INVALID, "capacity must be in 1..32767 (FixedCacheMap.java:17 stores heap slots as Short)");
    if (mode == 1 && max_paths < 1) return fail(GW_E_INVALID, "max_paths must be positive");
    GW_TRY(check_queries_host(g, queries, nq));
    if (nq == 0) return GW_OK;
    GW_CUDA(cudaSetDevice(g->device));
    double cache[16] = {0};
    for (int i = 1; i <= step; i++) cache[i] = pow(c, i);          // Math.pow(C, i)
    const size_t LEN1 = 2 * (size_t)step + 1, slots = (size_t)capacity + 1;
    DevBuf<int64_t> dq;
    DevBuf<double> dc, dw;
    DevBuf<int32_t> dv, dk, dpos, dsz;
    DevBuf<float> dval;
    DevBuf<unsigned long long> ds, dsteps;
    DevBuf<int> derr;
    if (dk.alloc((size_t)nq * slots) != cudaSuccess || dval.alloc((size_t)nq * slots) != cudaSuccess ||
        dpos.alloc((size_t)nq * (size_t)g->n) != cudaSuccess ||
        (mode == 1 && (dv.alloc((size_t)nq * 2 * LEN1 * (size_t)max_paths) != cudaSuccess || dw.alloc((size_t)nq * 2 * (size_t)max_paths) != cudaSuccess))) {
        cudaGetLastError();
        return fail(GW_E_TOO_LARGE, "cache / path buffers for %lld queries do not fit", (long long)nq);
    }
    GW_CUDA(dq.alloc((size_t)nq)); GW_CUDA(dc.alloc(16)); GW_CUDA(dsz.alloc((size_t)nq));
    GW_CUDA(ds.alloc((size_t)nq)); GW_CUDA(derr.alloc(1)); GW_CUDA(dsteps.alloc(1));
    GW_CUDA(cudaMemcpy(dq.p, queries, sizeof(int64_t) * (size_t)nq, cudaMemcpyHostToDevice));
    GW_CUDA(cudaMemcpy(dc.p, cache, sizeof(cache), cudaMemcpyHostToDevice));
    GW_CUDA(cudaMemcpy(ds.p, rng_state, sizeof(uint64_t) * (size_t)nq, cudaMemcpyHostToDevice));
    GW_CUDA(cudaMemset(dk.p, 0, sizeof(int32_t) * (size_t)nq * slots));
    GW_CUDA(cudaMemset(dval.p, 0, sizeof(float) * (size_t)nq * slots));
    GW_CUDA(cudaMemset(dpos.p, 0, sizeof(int32_t) * (size_t)nq * (size_t)g->n));
    GW_CUDA(cudaMemset(derr.p, 0, sizeof(int)));
    GW_CUDA(cudaMemset(dsteps.p, 0, sizeof(unsigned long long)));
    const ReplaySink sk{nullptr, dk.p, dval.p, dpos.p, dsz.p, capacity};
    if (mode == 0)
        k_simrank_javarng<true><<<(unsigned)((nq + 31) / 32), 32>>>(g->d_meta, g->d_col, dq.p, nq, g->n, sample, step, dc.p,
                                                                  (uint64_t *)ds.p, sk, dsteps.p);
    else
        k_topsim_javarng<true><<<(unsigned)((nq + 31) / 32), 32>>>(g->d_meta, g->d_col, dq.p, nq, g->n, sample, step, 0, dc.p, max_paths,
                                                                 dv.p, dw.p, (uint64_t *)ds.p, sk, derr.p);
    GW_LAUNCHED();
    int herr = 0;
    GW_CUDA(cudaMemcpy(&herr, derr.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (herr) return fail(GW_E_TOO_LARGE, "a level of the path tree exceeded max_paths = %lld", (long long)max_paths);
    // heap slot 0 is unused (1-based arrays): the caller receives slots 1..capacity of every query
    GW_CUDA(cudaMemcpy2D(out_keys, sizeof(int32_t) * (size_t)capacity, dk.p + 1, sizeof(int32_t) * slots,
                         sizeof(int32_t) * (size_t)capacity, (size_t)nq, cudaMemcpyDeviceToHost));
    GW_CUDA(cudaMemcpy2D(out_vals, sizeof(float) * (size_t)capacity, dval.p + 1, sizeof(float) * slots,
                         sizeof(float) * (size_t)capacity, (size_t)nq, cudaMemcpyDeviceToHost));
    GW_CUDA(cudaMemcpy(out_sizes, dsz.p, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost));
    GW_CUDA(cudaMemcpy(rng_state, ds.p, sizeof(uint64_t) * (size_t)nq, cudaMemcpyDeviceToHost));
    return GW_OK;
}

int gw_simrank_last_steps(const gw_graph *g, int64_t *steps) {
    if (!g || !steps) return fail(GW_E_INVALID, "bad arguments");
    if (g->d_simrank_scratch) {   // refresh from the device counter (covers the _dev entry point)
        unsigned long long hs = 0;
        GW_CUDA(cudaSetDevice(g->device));
        GW_CUDA(cudaDeviceSynchronize());
        GW_CUDA(cudaMemcpy(&hs, g->d_simrank_scratch, sizeof(hs), cudaMemcpyDeviceToHost));
        *steps = (int64_t)hs;
        return GW_OK;
    }
    *steps = g->simrank_last_steps;
    return GW_OK;
}

int gw_simrank_last_error(const gw_graph *g, int32_t *code) {
    if (!g || !code) return fail(GW_E_INVALID, "bad arguments");
    *code = 0;
    if (g->d_simrank_scratch) {
        int h = 0;
        GW_CUDA(cudaSetDevice(g->device));
        GW_CUDA(cudaDeviceSynchronize());
        GW_CUDA(cudaMemcpy(&h, (unsigned char *)g->d_simrank_scratch + 16, sizeof(h), cudaMemcpyDeviceToHost));
        *code = h;
    }
    return GW_OK;
}

int gw_simrank_last_slow_queries(const gw_graph *g, int64_t *count) {
    if (!g || !count) return fail(GW_E_INVALID, "bad arguments");
    *count = g->simrank_last_slow;
    if (g->d_simrank_scratch) {
        uint32_t h = 0;
        GW_CUDA(cudaSetDevice(g->device));
        GW_CUDA(cudaDeviceSynchronize());
        GW_CUDA(cudaMemcpy(&h, (unsigned char *)g->d_simrank_scratch + 32, sizeof(h), cudaMemcpyDeviceToHost));
        *count = (int64_t)h;
#ifdef HY_PROFILE
        {
            unsigned long long pr[16];
            GW_CUDA(cudaMemcpy(pr, (unsigned char *)g->d_simrank_scratch + 64, sizeof(pr), cudaMemcpyDeviceToHost));
            const char *nm[12] = {"prefix: rest (level switches)", "chain set-up", "chains", "top-k + reset + query switch", "walker 0: waits for a stage",
                                  "walker 0: walks (+ other phases)", "accumulator 0: waits for walker", "accumulator 0: inserts (+ other phases)",
                                  "prefix: contributions of even levels", "prefix: pass 1 (classify)", "prefix: pass 2 (scan)", "prefix: pass 3 (children)"};
            for (int i = 0; i < 12; i++) fprintf(stderr, "HY_PROFILE %-40s %12llu cycles\n", nm[i], pr[i]);
        }
#endif
#ifdef SR_PROFILE
        unsigned long long pr[16];
        GW_CUDA(cudaMemcpy(pr, (unsigned char *)g->d_simrank_scratch + 64, sizeof(pr), cudaMemcpyDeviceToHost));
        const char *nm[11] = {"walker walking", "walker waiting for a stage", "acc inserting", "acc waiting for walker", "acc barrier after A",
                              "B histogram", "B threshold+table candidates", "B log pass", "B tier-3 scan", "B ranking", "C reset"};
        for (int i = 0; i < 11; i++) fprintf(stderr, "SR_PROFILE %-30s %12llu cycles\n", nm[i], pr[i]);
#endif
    }
    return GW_OK;
}

int gw_simrank_exact(gw_graph *g, double c, int32_t iters, const int64_t *rows, int64_t nrows, double *out_dense) {
    if (!g) return fail(GW_E_INVALID, "graph is NULL");
    if (g->flags & GW_F_DIRECTED) return fail(GW_E_INVALID, "SimRank path is defined on undirected graphs");
    if (iters < 0 || nrows < 0 || (nrows > 0 && (!rows || !out_dense))) return fail(GW_E_INVALID, "bad arguments");
    GW_TRY(check_queries_host(g, rows, nrows));
    if (nrows == 0) return GW_OK;
    GW_CUDA(cudaSetDevice(g->device));
    const int64_t n = g->n;
    size_t freeb = 0;
    GW_TRY(device_info(nullptr, &freeb));
    size_t mat = sizeof(double) * (size_t)n * (size_t)n;
    if (mat * 2 + sizeof(double) * (size_t)nrows * n > freeb / 10 * 9)
        return fail(GW_E_TOO_LARGE, "exact SimRank needs two dense %lld x %lld fp64 matrices", (long long)n, (long long)n);
    DevBuf<double> A, B, out;
    DevBuf<int64_t> dr;
    GW_CUDA(A.alloc((size_t)n * n)); GW_CUDA(B.alloc((size_t)n * n));
    GW_CUDA(out.alloc((size_t)nrows * n)); GW_CUDA(dr.alloc((size_t)nrows));
    GW_CUDA(cudaMemcpy(dr.p, rows, sizeof(int64_t) * (size_t)nrows, cudaMemcpyHostToDevice));
    GW_CUDA(cudaMemset(A.p, 0, mat));
    k_identity<<<(unsigned)((n + 255) / 256), 256>>>(A.p, n); GW_LAUNCHED();
    dim3 ga((unsigned)((n + 255) / 256), (unsigned)n), gt((unsigned)((n + 31) / 32), (unsigned)((n + 31) / 32));
    for (int it = 0; it < iters; it++) {
        // S <- c * P S P^T with the diagonal pinned to 1:   T = P S ; S' = c * P T^T
        k_row_average<<<ga, 256>>>(g->d_meta, g->d_col, n, A.p, B.p, 1.0, 0); GW_LAUNCHED();
        k_transpose<<<gt, dim3(32, 8)>>>(B.p, A.p, n); GW_LAUNCHED();
        k_row_average<<<ga, 256>>>(g->d_meta, g->d_col, n, A.p, B.p, c, 1); GW_LAUNCHED();
        std::swap(A.p, B.p);
    }
    dim3 gg((unsigned)((n + 255) / 256), (unsigned)nrows);
    k_gather_rows_zero_diag<<<gg, 256>>>(A.p, n, dr.p, nrows, out.p); GW_LAUNCHED();
    GW_CUDA(cudaMemcpy(out_dense, out.p, sizeof(double) * (size_t)nrows * n, cudaMemcpyDeviceToHost));
    return GW_OK;
}

}  // extern "C"
