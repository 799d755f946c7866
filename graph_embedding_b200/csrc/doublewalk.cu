// doublewalk.cu — DoubleRandomWalk (DeepSim/TopSimAll/src/simrank/DoubleRandomWalk.java) on the device.
//
// The reference samples, for EVERY vertex, SAMPLE independent walks of STEP steps (samplePaths :50-65) and scores a
// pair (v, w) by looking at all SAMPLE^2 path pairs: the first position where the two paths hold the same vertex
// contributes C^(position+1) (getSim :77-91).  O(V^2 SAMPLE^2 STEP) compares -- a research variant the reference
// only runs on small graphs -- but embarrassingly parallel over (v, w).
//
//  * k_dw_sample         one thread per (vertex, sample): Philox4x32-10 keyed by (seed, vertex), counter (sample, block)
//  * k_dw_sample_javarng one thread per vertex walks its SAMPLE paths with java.util.Random from a given 48-bit state
//                        (replay; the host chains the states, simrank.py)
//  * k_dw_sims           one thread per (row r, column w).  Paths are kept vertex-minor on the device,
//                        P[(i * STEP + s) * nv + v], so the 32 columns of a warp read one 128-byte line per (i, s) and
//                        the row's own entries are warp-uniform broadcasts.  EXACT instantiation: fp64 adds in the
//                        reference's (i, j) order for the pair (min, max) as computeSims :67-75 calls it -- bit-exact;
//                        counting instantiation: integer first-meeting counts per position, combined once at the end
//                        (production; differs from the above only in fp64 rounding order).
#include <cmath>

#include "common.cuh"

namespace gw {

__global__ void k_dw_sample(const uint2 *__restrict__ meta, const int32_t *__restrict__ col, const int64_t *__restrict__ verts,
                            int64_t nv, int32_t sample, int32_t step, uint64_t seed, int32_t *__restrict__ P) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nv * sample) return;
    const int64_t vi = t % nv;                                   // vertex-minor: neighbouring threads write neighbouring words
    const int32_t i = (int32_t)(t / nv);
    const int32_t v = (int32_t)verts[vi];
    const uint2 key = make_uint2((uint32_t)seed ^ (uint32_t)v, (uint32_t)(seed >> 32) ^ 0x44524157u);
    int32_t cur = v;
    uint4 r = make_uint4(0, 0, 0, 0);
    bool dead = false;
    for (int s = 0; s < step; s++) {
        int32_t w = 0;                                           // Java's int[] default for slots after a dead end
        if (!dead) {
            if ((s & 3) == 0) r = Philox::gen(make_uint4((uint32_t)i, (uint32_t)(s >> 2), (uint32_t)((uint64_t)v >> 32), 0x4457u), key);
            const uint32_t bits = (s & 3) == 0 ? r.x : (s & 3) == 1 ? r.y : (s & 3) == 2 ? r.z : r.w;
            const uint2 m = meta[cur];
            if (m.y == 0) { w = -1; dead = true; }
            else { cur = col[m.x + scale_u32(bits, m.y)]; w = cur; }
        }
        P[((size_t)i * step + s) * (size_t)nv + vi] = w;
    }
}

__global__ void k_dw_sample_javarng(const uint2 *__restrict__ meta, const int32_t *__restrict__ col,
                                    const int64_t *__restrict__ verts, int64_t nv, int32_t sample, int32_t step,
                                    uint64_t *__restrict__ states, int32_t *__restrict__ P) {
    const int64_t vi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (vi >= nv) return;
    const int32_t v = (int32_t)verts[vi];
    uint64_t seed = states[vi];
    for (int32_t i = 0; i < sample; i++) {                       // sample(src) :56-65
        int32_t cur = v;
        bool dead = false;
        for (int s = 0; s < step; s++) {
            int32_t w = 0;
            if (!dead) {
                const uint2 m = meta[cur];
                if (m.y == 0) { w = -1; dead = true; }           // randNeighbor == -1 is stored, then break
                else { cur = col[m.x + (uint32_t)jr_next_int(seed, (int32_t)m.y)]; w = cur; }
            }
            P[((size_t)i * step + s) * (size_t)nv + vi] = w;
        }
    }
    states[vi] = seed;
}

// host layout [nv][sample][step]  <->  device layout [sample][step][nv]
__global__ void k_dw_to_vertex_minor(const int32_t *__restrict__ H, int64_t nv, int64_t per, int32_t *__restrict__ P) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nv * per) return;
    const int64_t vi = t % nv, k = t / nv;
    P[k * nv + vi] = H[vi * per + k];
}
__global__ void k_dw_to_vertex_major(const int32_t *__restrict__ P, int64_t nv, int64_t per, int32_t *__restrict__ H) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nv * per) return;
    const int64_t k = t % per, vi = t / per;
    H[vi * per + k] = P[k * nv + vi];
}

template <bool EXACT>
__global__ void k_dw_sims(const int32_t *__restrict__ P, int64_t nv, int32_t sample, int32_t step,
                          const double *__restrict__ cache, const int64_t *__restrict__ rows, double *__restrict__ out) {
    const int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (w >= nv) return;
    const int64_t r = rows[blockIdx.y];
    double *o = out + (size_t)blockIdx.y * (size_t)nv;
    if (w == r) { o[w] = 0.0; return; }                          // the diagonal is never written (:69)
    const int64_t va = r < w ? r : w, vb = r < w ? w : r;        // getSim(i, j) is only ever called with i < j
    double result = 0.0;
    unsigned long long cnt[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int32_t i = 0; i < sample; i++) {
        const int32_t *pa = P + (size_t)i * step * (size_t)nv + va;
        for (int32_t j = 0; j < sample; j++) {
            const int32_t *pb = P + (size_t)j * step * (size_t)nv + vb;
            for (int s = 0; s < step; s++) {                     // :81-87
                const int32_t x = __ldg(pa + (size_t)s * nv);
                if (x == -1) break;
                const int32_t y = __ldg(pb + (size_t)s * nv);
                if (y == -1) break;
                if (x == y) {
                    if constexpr (EXACT) result = __dadd_rn(result, cache[s + 1]);
                    else cnt[s]++;
                    break;
                }
            }
        }
    }
    if constexpr (!EXACT)
        for (int s = 0; s < step; s++) result += (double)cnt[s] * cache[s + 1];
    o[w] = __ddiv_rn(result, (double)(sample * sample));         // :90, int product
}

static int dw_check(const gw_graph *g, int64_t nv, int32_t sample, int32_t step) {
    if (!g) return fail(GW_E_INVALID, "graph is NULL");
    if (g->flags & GW_F_DIRECTED) return fail(GW_E_INVALID, "SimRank path is defined on undirected graphs (structures/Graph.java)");
    if (nv < 0) return fail(GW_E_INVALID, "bad arguments");
    if (step < 1 || step > 9) return fail(GW_E_INVALID, "step must be in 1..9");
    if (sample < 1 || sample > 46340) return fail(GW_E_INVALID, "sample must be in 1..46340 (SAMPLE * SAMPLE is an int product, DoubleRandomWalk.java:90)");
    return GW_OK;
}
static int dw_check_vertices(const gw_graph *g, const int64_t *v, int64_t nv) {
    for (int64_t i = 0; i < nv; i++)
        if (v[i] < 0 || v[i] >= g->n) return fail(GW_E_KEY, "vertex %lld is outside [0, %lld)", (long long)v[i], (long long)g->n);
    return GW_OK;
}

}  // namespace gw

using namespace gw;

extern "C" {

int gw_double_walk_paths(gw_graph *g, const int64_t *vertices, int64_t nv, int32_t sample, int32_t step, uint64_t seed,
                         uint64_t *rng_state, int32_t *out_paths) {
    GW_TRY(dw_check(g, nv, sample, step));
    if (nv > 0 && (!vertices || !out_paths)) return fail(GW_E_INVALID, "bad arguments");
    GW_TRY(dw_check_vertices(g, vertices, nv));
    if (nv == 0) return GW_OK;
    GW_CUDA(cudaSetDevice(g->device));
    const size_t per = (size_t)sample * step, total = per * (size_t)nv;
    DevBuf<int64_t> dv;
    DevBuf<int32_t> dP, dH;
    DevBuf<unsigned long long> ds;
    GW_CUDA(dv.alloc((size_t)nv)); GW_CUDA(dP.alloc(total)); GW_CUDA(dH.alloc(total));
    GW_CUDA(cudaMemcpy(dv.p, vertices, sizeof(int64_t) * (size_t)nv, cudaMemcpyHostToDevice));
    if (rng_state) {
        GW_CUDA(ds.alloc((size_t)nv));
        GW_CUDA(cudaMemcpy(ds.p, rng_state, sizeof(uint64_t) * (size_t)nv, cudaMemcpyHostToDevice));
        k_dw_sample_javarng<<<(unsigned)((nv + 63) / 64), 64>>>(g->d_meta, g->d_col, dv.p, nv, sample, step, (uint64_t *)ds.p, dP.p);
        GW_LAUNCHED();
        GW_CUDA(cudaMemcpy(rng_state, ds.p, sizeof(uint64_t) * (size_t)nv, cudaMemcpyDeviceToHost));
    } else {
        const int64_t threads = nv * sample;
        k_dw_sample<<<(unsigned)((threads + 255) / 256), 256>>>(g->d_meta, g->d_col, dv.p, nv, sample, step, seed, dP.p);
        GW_LAUNCHED();
    }
    k_dw_to_vertex_major<<<(unsigned)((total + 255) / 256), 256>>>(dP.p, nv, (int64_t)per, dH.p);
    GW_LAUNCHED();
    GW_CUDA(cudaMemcpy(out_paths, dH.p, sizeof(int32_t) * total, cudaMemcpyDeviceToHost));
    return GW_OK;
}

int gw_double_walk_sims(gw_graph *g, const int32_t *paths, int64_t nv, int32_t sample, int32_t step, double c,
                        const int64_t *rows, int64_t nrows, int32_t exact_order, double *out_dense) {
    GW_TRY(dw_check(g, nv, sample, step));
    if (nrows < 0 || (nrows > 0 && (!rows || !out_dense)) || (nv > 0 && !paths)) return fail(GW_E_INVALID, "bad arguments");
    for (int64_t i = 0; i < nrows; i++)
        if (rows[i] < 0 || rows[i] >= nv) return fail(GW_E_KEY, "row %lld is outside [0, %lld)", (long long)rows[i], (long long)nv);
    if (nrows == 0 || nv == 0) return GW_OK;
    GW_CUDA(cudaSetDevice(g->device));
    const size_t per = (size_t)sample * step, total = per * (size_t)nv;
    double cache[16] = {0};
    for (int i = 0; i <= step; i++) cache[i] = pow(c, i);          // :33-35
    DevBuf<int32_t> dP, dH;
    DevBuf<int64_t> dr;
    DevBuf<double> dc, dout;
    GW_CUDA(dP.alloc(total)); GW_CUDA(dH.alloc(total)); GW_CUDA(dr.alloc((size_t)nrows)); GW_CUDA(dc.alloc(16));
    if (dout.alloc((size_t)nrows * (size_t)nv) != cudaSuccess) {
        cudaGetLastError();
        return fail(GW_E_TOO_LARGE, "%lld x %lld result rows do not fit", (long long)nrows, (long long)nv);
    }
    GW_CUDA(cudaMemcpy(dH.p, paths, sizeof(int32_t) * total, cudaMemcpyHostToDevice));
    GW_CUDA(cudaMemcpy(dr.p, rows, sizeof(int64_t) * (size_t)nrows, cudaMemcpyHostToDevice));
    GW_CUDA(cudaMemcpy(dc.p, cache, sizeof(cache), cudaMemcpyHostToDevice));
    k_dw_to_vertex_minor<<<(unsigned)((total + 255) / 256), 256>>>(dH.p, nv, (int64_t)per, dP.p);
    GW_LAUNCHED();
    for (int64_t r0 = 0; r0 < nrows; r0 += 32768) {                // gridDim.y limit
        const int64_t nr = nrows - r0 < 32768 ? nrows - r0 : 32768;
        dim3 grid((unsigned)((nv + 127) / 128), (unsigned)nr);
        if (exact_order) k_dw_sims<true><<<grid, 128>>>(dP.p, nv, sample, step, dc.p, dr.p + r0, dout.p + (size_t)r0 * nv);
        else k_dw_sims<false><<<grid, 128>>>(dP.p, nv, sample, step, dc.p, dr.p + r0, dout.p + (size_t)r0 * nv);
        GW_LAUNCHED();
    }
    GW_CUDA(cudaMemcpy(out_dense, dout.p, sizeof(double) * (size_t)nrows * (size_t)nv, cudaMemcpyDeviceToHost));
    return GW_OK;
}

}  // extern "C"
