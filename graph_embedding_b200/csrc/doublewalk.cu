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

// ---------------- TopSim_doubleSample / TopSim_Dev: path-mass trees and their products ----------------
// sample(src) of both classes (TopSim_doubleSample.java:66-151 = TopSim_Dev.java:104-199) is the enumerate-or-sample
// tree of TopSim_singleSample, STEP levels deep; computePath (:153-178 / :200-226) keeps, per level, the weight of the
// LAST path of the queue standing on each target (an overwrite, not a sum).  Only (last vertex, weight) of a path is
// ever read, so a level is two arrays.  One thread per tree: the queue order IS the result (overwrites) and, in replay
// mode, the java.util.Random draw order.
template <bool JAVA>
__global__ void k_mass_tree(const uint2 *__restrict__ meta, const int32_t *__restrict__ col, const int64_t *__restrict__ sources,
                            int64_t ns, int64_t n, double weight0, int32_t step, int64_t cap, int32_t *__restrict__ cbuf,
                            double *__restrict__ wbuf, uint64_t seed, uint64_t call_base, uint64_t *__restrict__ states,
                            double *__restrict__ mass, int *__restrict__ err) {
    const int64_t ci = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (ci >= ns) return;
    const int32_t src = (int32_t)sources[ci];
    int32_t *c0 = cbuf + (size_t)ci * 2 * cap, *c1 = c0 + cap;
    double *w0 = wbuf + (size_t)ci * 2 * cap, *w1 = w0 + cap;
    double *m = mass + (size_t)ci * (size_t)n * (size_t)(step + 1);
    uint64_t jstate = JAVA ? states[ci] : 0;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x4d415353u);
    const uint64_t call = call_base + (uint64_t)ci;
    uint32_t draws = 0;
    uint4 r = make_uint4(0, 0, 0, 0);
    int64_t n0 = 1;
    bool overflow = false;
    c0[0] = src; w0[0] = weight0;
    for (int path_len = 0; path_len < step && !overflow; path_len++) {
        int64_t n1 = 0;
        for (int64_t k = 0; k < n0 && !overflow; k++) {
            const int32_t c = c0[k];
            const double wt = w0[k];
            const uint2 mt = meta[c];
            const int d = (int)mt.y;
            if (d != 0 && wt >= (double)d) {
                const double nsw = __ddiv_rn(wt, (double)d);
                if (n1 + d > cap) { overflow = true; break; }
                for (int j = 0; j < d; j++) { c1[n1] = col[mt.x + j]; w1[n1] = nsw; n1++; }
            } else {
                const int number = ((double)(int)wt == wt) ? (int)wt : (int)wt + 1;
                for (int j = 0; j < number; j++) {
                    if (d == 0) break;                               // randNeighbor == -1
                    uint32_t idx;
                    if constexpr (JAVA) idx = (uint32_t)jr_next_int(jstate, d);
                    else {
                        if ((draws & 3) == 0) r = Philox::gen(make_uint4(draws >> 2, (uint32_t)call, (uint32_t)(call >> 32), 0x4d54u), key);
                        const uint32_t bits = (draws & 3) == 0 ? r.x : (draws & 3) == 1 ? r.y : (draws & 3) == 2 ? r.z : r.w;
                        draws++;
                        idx = scale_u32(bits, (uint32_t)d);
                    }
                    if (n1 + 1 > cap) { overflow = true; break; }
                    c1[n1] = col[mt.x + idx]; w1[n1] = __ddiv_rn(wt, (double)number); n1++;
                }
            }
        }
        if (overflow) break;
        int32_t *tc = c0; c0 = c1; c1 = tc;
        double *tw = w0; w0 = w1; w1 = tw;
        n0 = n1;
        const int level = path_len + 1;                              // computePath(queue, level, level)
        for (int64_t k = 0; k < n0; k++)
            if (c0[k] != src) m[(size_t)c0[k] * (step + 1) + level] = w0[k];
    }
    if (overflow) atomicExch(err, 1);
    if constexpr (JAVA) states[ci] = jstate;
}
__global__ void k_fill_f64(double *__restrict__ p, size_t n, double v) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// getSim (TopSim_doubleSample.java:189-199 / TopSim_Dev.java:233-244) for a list of (a, b) tree pairs.
// EXACT: one thread per pair, the reference's (target, level) order, (cache * a) * b.  Otherwise one warp per pair:
// lanes stride over the flattened [n][step+1] rows (coalesced on both), shuffle reduction.
template <bool EXACT>
__global__ void k_mass_sims(const double *__restrict__ mass, int64_t n, int32_t step, const double *__restrict__ cache,
                            const int64_t *__restrict__ pa, const int64_t *__restrict__ pb, int64_t np, double *__restrict__ out) {
    const size_t row = (size_t)n * (size_t)(step + 1);
    if constexpr (EXACT) {
        const int64_t pi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
        if (pi >= np) return;
        const double *ma = mass + (size_t)pa[pi] * row, *mb = mass + (size_t)pb[pi] * row;
        double result = 0.0;
        for (int64_t i = 0; i < n; i++)
            for (int s = 1; s <= step; s++) {
                const double a = ma[(size_t)i * (step + 1) + s], b = mb[(size_t)i * (step + 1) + s];
                if (a >= 0 && b >= 0) result = __dadd_rn(result, __dmul_rn(__dmul_rn(cache[s], a), b));
            }
        out[pi] = result;
    } else {
        const int64_t pi = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
        const int lane = threadIdx.x & 31;
        if (pi >= np) return;
        const double *ma = mass + (size_t)pa[pi] * row, *mb = mass + (size_t)pb[pi] * row;
        double result = 0.0;
        for (size_t e = lane; e < row; e += 32) {
            const int s = (int)(e % (size_t)(step + 1));
            const double a = ma[e], b = mb[e];
            if (s >= 1 && a >= 0 && b >= 0) result += cache[s] * a * b;
        }
        for (int o = 16; o; o >>= 1) result += __shfl_xor_sync(0xffffffffu, result, o);
        if (lane == 0) out[pi] = result;
    }
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

int gw_topsim_mass(gw_graph *g, const int64_t *sources, int64_t ns, double weight, int32_t step, int64_t max_paths,
                   uint64_t seed, uint64_t call_id_base, uint64_t *rng_state, double *out_mass) {
    if (!g) return fail(GW_E_INVALID, "graph is NULL");
    if (g->flags & GW_F_DIRECTED) return fail(GW_E_INVALID, "SimRank path is defined on undirected graphs (structures/Graph.java)");
    if (ns < 0 || (ns > 0 && (!sources || !out_mass))) return fail(GW_E_INVALID, "bad arguments");
    if (step < 1 || step > 15) return fail(GW_E_INVALID, "step must be in 1..15");
    if (!(weight >= 0) || weight > 2e9 || max_paths < 1) return fail(GW_E_INVALID, "weight must be in [0, 2e9] and max_paths positive");
    GW_TRY(dw_check_vertices(g, sources, ns));
    if (ns == 0) return GW_OK;
    GW_CUDA(cudaSetDevice(g->device));
    const size_t total = (size_t)ns * (size_t)g->n * (size_t)(step + 1);
    DevBuf<int64_t> dsrc;
    DevBuf<int32_t> dc;
    DevBuf<double> dw, dm;
    DevBuf<unsigned long long> dst;
    DevBuf<int> derr;
    if (dc.alloc((size_t)ns * 2 * (size_t)max_paths) != cudaSuccess || dw.alloc((size_t)ns * 2 * (size_t)max_paths) != cudaSuccess ||
        dm.alloc(total) != cudaSuccess) {
        cudaGetLastError();
        return fail(GW_E_TOO_LARGE, "level buffers / mass rows for %lld trees do not fit", (long long)ns);
    }
    GW_CUDA(dsrc.alloc((size_t)ns)); GW_CUDA(derr.alloc(1));
    GW_CUDA(cudaMemcpy(dsrc.p, sources, sizeof(int64_t) * (size_t)ns, cudaMemcpyHostToDevice));
    GW_CUDA(cudaMemset(derr.p, 0, sizeof(int)));
    k_fill_f64<<<(unsigned)((total + 255) / 256), 256>>>(dm.p, total, -1.0);
    GW_LAUNCHED();
    if (rng_state) {
        GW_CUDA(dst.alloc((size_t)ns));
        GW_CUDA(cudaMemcpy(dst.p, rng_state, sizeof(uint64_t) * (size_t)ns, cudaMemcpyHostToDevice));
        k_mass_tree<true><<<(unsigned)((ns + 31) / 32), 32>>>(g->d_meta, g->d_col, dsrc.p, ns, g->n, weight, step, max_paths, dc.p, dw.p,
                                                            seed, call_id_base, (uint64_t *)dst.p, dm.p, derr.p);
    } else {
        k_mass_tree<false><<<(unsigned)((ns + 31) / 32), 32>>>(g->d_meta, g->d_col, dsrc.p, ns, g->n, weight, step, max_paths, dc.p, dw.p,
                                                             seed, call_id_base, nullptr, dm.p, derr.p);
    }
    GW_LAUNCHED();
    int herr = 0;
    GW_CUDA(cudaMemcpy(&herr, derr.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (herr) return fail(GW_E_TOO_LARGE, "a level of the path tree exceeded max_paths = %lld", (long long)max_paths);
    GW_CUDA(cudaMemcpy(out_mass, dm.p, sizeof(double) * total, cudaMemcpyDeviceToHost));
    if (rng_state) GW_CUDA(cudaMemcpy(rng_state, dst.p, sizeof(uint64_t) * (size_t)ns, cudaMemcpyDeviceToHost));
    return GW_OK;
}

int gw_topsim_mass_sims(gw_graph *g, const double *mass, int64_t ns, int32_t step, double c, const int64_t *pair_a,
                        const int64_t *pair_b, int64_t npairs, int32_t exact_order, double *out) {
    if (!g) return fail(GW_E_INVALID, "graph is NULL");
    if (ns < 0 || npairs < 0 || (npairs > 0 && (!mass || !pair_a || !pair_b || !out))) return fail(GW_E_INVALID, "bad arguments");
    if (step < 1 || step > 15) return fail(GW_E_INVALID, "step must be in 1..15");
    for (int64_t i = 0; i < npairs; i++)
        if (pair_a[i] < 0 || pair_a[i] >= ns || pair_b[i] < 0 || pair_b[i] >= ns)
            return fail(GW_E_KEY, "pair %lld refers to a tree outside [0, %lld)", (long long)i, (long long)ns);
    if (npairs == 0) return GW_OK;
    GW_CUDA(cudaSetDevice(g->device));
    const size_t total = (size_t)ns * (size_t)g->n * (size_t)(step + 1);
    double cache[16] = {0};
    for (int i = 0; i <= step; i++) cache[i] = pow(c, i);
    DevBuf<double> dm, dc, dout;
    DevBuf<int64_t> da, db;
    if (dm.alloc(total) != cudaSuccess) { cudaGetLastError(); return fail(GW_E_TOO_LARGE, "mass rows of %lld trees do not fit", (long long)ns); }
    GW_CUDA(dc.alloc(16)); GW_CUDA(dout.alloc((size_t)npairs)); GW_CUDA(da.alloc((size_t)npairs)); GW_CUDA(db.alloc((size_t)npairs));
    GW_CUDA(cudaMemcpy(dm.p, mass, sizeof(double) * total, cudaMemcpyHostToDevice));
    GW_CUDA(cudaMemcpy(dc.p, cache, sizeof(cache), cudaMemcpyHostToDevice));
    GW_CUDA(cudaMemcpy(da.p, pair_a, sizeof(int64_t) * (size_t)npairs, cudaMemcpyHostToDevice));
    GW_CUDA(cudaMemcpy(db.p, pair_b, sizeof(int64_t) * (size_t)npairs, cudaMemcpyHostToDevice));
    if (exact_order) k_mass_sims<true><<<(unsigned)((npairs + 127) / 128), 128>>>(dm.p, g->n, step, dc.p, da.p, db.p, npairs, dout.p);
    else k_mass_sims<false><<<(unsigned)((npairs * 32 + 127) / 128), 128>>>(dm.p, g->n, step, dc.p, da.p, db.p, npairs, dout.p);
    GW_LAUNCHED();
    GW_CUDA(cudaMemcpy(out, dout.p, sizeof(double) * (size_t)npairs, cudaMemcpyDeviceToHost));
    return GW_OK;
}

}  // extern "C"
