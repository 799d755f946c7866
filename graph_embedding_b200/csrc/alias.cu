// alias.cu — bit-exact alias-table construction on the device.
//
// Replaces node2vec/src/node2vec.py:116-147 alias_setup, :61-81 get_alias_edge and :83-113
// preprocess_transition_probs.  The table algorithm (two LIFO stacks, fp64 update
// q[large] = (q[large] + q[small]) - 1.0) is sequential per table and its result depends on the
// fp64 operation order, so parallelism comes from ACROSS tables: one thread per table, all
// arithmetic through __d*_rn intrinsics (no FMA contraction), normalisation as
// fl(K * fl(w / sum)) with the sum accumulated left to right over the sorted neighbour list.
// The two stacks live inside the output J array as intrusive linked lists (an index is on at
// most one stack at a time and its J entry is only final once it has been popped as `small`),
// so no scratch memory is needed.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include <cub/cub.cuh>

#include "common.cuh"

namespace gw {

// Builds one table in place.  On entry q[0..K) holds K*prob; J is uninitialised.
__device__ __forceinline__ void alias_build(int32_t *__restrict__ J, double *__restrict__ q, int32_t K) {
    int32_t small_top = -1, large_top = -1;   // stack heads; J[k] = next-below link
    for (int32_t kk = 0; kk < K; kk++) {      // node2vec.py:129-134 (append in index order)
        if (q[kk] < 1.0) { J[kk] = small_top; small_top = kk; }
        else { J[kk] = large_top; large_top = kk; }
    }
    while (small_top >= 0 && large_top >= 0) {                // :136
        int32_t small = small_top; small_top = J[small];      // smaller.pop()
        int32_t large = large_top; large_top = J[large];      // larger.pop()
        J[small] = large;                                     // :140 (final)
        double ql = __dadd_rn(__dadd_rn(q[large], q[small]), -1.0);   // :141
        q[large] = ql;
        if (ql < 1.0) { J[large] = small_top; small_top = large; }   // :142-145
        else { J[large] = large_top; large_top = large; }
    }
    // whatever is still stacked never received an alias: J stays 0 (np.zeros, :125)
    while (small_top >= 0) { int32_t k = small_top; small_top = J[k]; J[k] = 0; }
    while (large_top >= 0) { int32_t k = large_top; large_top = J[k]; J[k] = 0; }
}

__global__ void k_alias_single(const double *__restrict__ probs, int32_t K, int32_t *J, double *q) {
    if (blockIdx.x || threadIdx.x) return;
    for (int32_t k = 0; k < K; k++) q[k] = __dmul_rn((double)K, probs[k]);   // :130
    alias_build(J, q, K);
}

// alias_nodes (node2vec.py:91-97): thread per vertex
__global__ void k_alias_nodes(const uint2 *__restrict__ meta, const double *__restrict__ w, int64_t n,
                              int32_t *__restrict__ J, double *__restrict__ q) {
    int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (v >= n) return;
    uint2 m = meta[v];
    int32_t K = (int32_t)m.y;
    if (K == 0) return;
    int32_t *Jv = J + m.x;
    double *qv = q + m.x;
    if (w) {
        const double *wv = w + m.x;
        double norm = 0.0;
        for (int32_t k = 0; k < K; k++) norm = __dadd_rn(norm, wv[k]);          // sum(), :94
        for (int32_t k = 0; k < K; k++) qv[k] = __dmul_rn((double)K, __ddiv_rn(wv[k], norm));
    } else {
        double pr = __ddiv_rn(1.0, (double)K);                                   // float(1)/K
        double qq = __dmul_rn((double)K, pr);
        for (int32_t k = 0; k < K; k++) qv[k] = qq;
    }
    alias_build(Jv, qv, K);
}

__device__ __forceinline__ bool has_edge(const uint2 *__restrict__ meta, const int32_t *__restrict__ col,
                                         int32_t a, int32_t b) {   // b in sorted N_out(a)
    uint2 m = meta[a];
    uint32_t lo = 0, hi = m.y;
    const int32_t *row = col + m.x;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (row[mid] < b) lo = mid + 1; else hi = mid;
    }
    return lo < m.y && row[lo] == b;
}

__global__ void k_entry_degree(const uint2 *__restrict__ meta, const int32_t *__restrict__ col, int64_t nnz,
                               int64_t *__restrict__ out) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e < nnz) out[e] = meta[col[e]].y;
    if (e == nnz) out[e] = 0;
}
__global__ void k_entry_keys(const uint2 *__restrict__ meta, const int32_t *__restrict__ col, int64_t nnz,
                             uint32_t *__restrict__ key, uint32_t *__restrict__ val) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e < nnz) { key[e] = meta[col[e]].y; val[e] = (uint32_t)e; }
}

// alias_edges (node2vec.py:61-81, :99-108): thread per directed CSR entry e = (u -> v)
// `order` lists the entries by DEcreasing table length: the 32 tables of a warp have (almost) the same
// length, so the lanes leave the sequential Vose loop together, and the longest tables start first.
// (In CSR order a warp's time is its longest table's: hubs are the targets of many edges, and the
// blog graph ran 10x slower that way.)
__global__ void k_alias_edges(const uint2 *__restrict__ meta, const int32_t *__restrict__ col,
                              const double *__restrict__ w, const int64_t *__restrict__ row_ptr, int64_t n,
                              int64_t nnz, double p, double qparam, const int64_t *__restrict__ aeoff,
                              const uint32_t *__restrict__ order, int32_t *__restrict__ J, double *__restrict__ q) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nnz) return;
    const int64_t e = order ? (int64_t)order[t] : t;
    // source u of entry e: last row with row_ptr[u] <= e
    int64_t lo = 0, hi = n;
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (row_ptr[mid] <= e) lo = mid; else hi = mid;
    }
    int32_t u = (int32_t)lo, v = col[e];
    uint2 mv = meta[v];
    int32_t K = (int32_t)mv.y;
    if (K == 0) return;
    int32_t *Je = J + aeoff[e];
    double *qe = q + aeoff[e];
    double norm = 0.0;
    for (int32_t k = 0; k < K; k++) {
        int32_t nbr = col[mv.x + k];
        double wt = w ? w[mv.x + k] : 1.0;
        double un;
        if (nbr == u) un = __ddiv_rn(wt, p);                        // :71-72
        else if (has_edge(meta, col, nbr, u)) un = wt;              // :73-74
        else un = __ddiv_rn(wt, qparam);                            // :75-76
        qe[k] = un;
        norm = __dadd_rn(norm, un);                                 // :78 sum()
    }
    for (int32_t k = 0; k < K; k++) qe[k] = __dmul_rn((double)K, __ddiv_rn(qe[k], norm));   // :79, :130
    alias_build(Je, qe, K);
}

// ---- warp-per-table builder ------------------------------------------------------------------------
// A table of length K is staged in shared memory (q as fp64, J as int32: 12 bytes per entry): the
// unnormalised weights are evaluated by the 32 lanes (coalesced reads of N(dst), one binary search
// over N(src) per entry), lane 0 runs the two inherently sequential passes (left-to-right sum and
// the LIFO pairing loop, bit-exact operation order) against shared memory instead of HBM, and the
// finished table leaves in coalesced stores.  With one thread per table every step of the pairing
// loop is a random DRAM access into a private table (blog.txt: 3.7e8 entries took 0.26-1.6 s).
// One table per CTA (32 threads for short tables, 128 for long ones: the parallel passes scale, the
// sequential ones run on thread 0), dynamic shared memory = 12 * kcap bytes; the launcher picks kcap
// and the CTA width per size class.
__global__ void __launch_bounds__(128) k_alias_edges_warp(const uint2 *__restrict__ meta, const int32_t *__restrict__ col,
                                                         const double *__restrict__ w, const int64_t *__restrict__ row_ptr,
                                                         int64_t n, double p, double qparam,
                                                         const int64_t *__restrict__ aeoff, const uint32_t *__restrict__ order,
                                                         int64_t first, int64_t count, int32_t kcap,
                                                         int32_t *__restrict__ Jg, double *__restrict__ qg) {
    extern __shared__ __align__(16) unsigned char alias_smem[];
    double *q = reinterpret_cast<double *>(alias_smem);
    int32_t *J = reinterpret_cast<int32_t *>(alias_smem + sizeof(double) * ((size_t)kcap + 2));
    const int lane = threadIdx.x, nth = blockDim.x;
    for (int64_t t = blockIdx.x; t < count; t += gridDim.x) {
        const int64_t e = (int64_t)order[first + t];
        int64_t lo = 0, hi = n;                      // source u of entry e: last row with row_ptr[u] <= e
        while (hi - lo > 1) {
            int64_t mid = (lo + hi) >> 1;
            if (row_ptr[mid] <= e) lo = mid; else hi = mid;
        }
        const int32_t u = (int32_t)lo, v = col[e];
        const uint2 mv = meta[v];
        const int32_t K = (int32_t)mv.y;
        if (K == 0 || K > kcap) continue;                  // uniform across the CTA
        for (int32_t k = lane; k < K; k += nth) {                             // node2vec.py:69-76
            const int32_t nbr = col[mv.x + k];
            const double wt = w ? w[mv.x + k] : 1.0;
            double un;
            if (nbr == u) un = __ddiv_rn(wt, p);
            else if (has_edge(meta, col, nbr, u)) un = wt;
            else un = __ddiv_rn(wt, qparam);
            q[k] = un;
        }
        __syncthreads();
        if (lane == 0) {
            double norm = 0.0;                                                // :78 sum(), left to right
            int32_t k = 0;
            for (; k + 4 <= K; k += 4) {                                      // four loads in flight, adds in order
                const double a0 = q[k], a1 = q[k + 1], a2 = q[k + 2], a3 = q[k + 3];
                norm = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(norm, a0), a1), a2), a3);
            }
            for (; k < K; k++) norm = __dadd_rn(norm, q[k]);
            q[K] = norm;                                                      // spare slot (kcap + 1 doubles are allocated)
        }
        __syncthreads();
        const double norm = q[K];
        for (int32_t k = lane; k < K; k += nth) q[k] = __dmul_rn((double)K, __ddiv_rn(q[k], norm));   // :79, :130
        __syncthreads();
        if (lane == 0) alias_build(J, q, K);                                  // :125-147 on shared memory
        __syncthreads();
        int32_t *Je = Jg + aeoff[e];
        double *qe = qg + aeoff[e];
        for (int32_t k = lane; k < K; k += nth) { Je[k] = J[k]; qe[k] = q[k]; }
        __syncthreads();
    }
}

}  // namespace gw

using namespace gw;

extern "C" {

int gw_alias_setup(const double *probs, int64_t K, int32_t *J, double *q) {
    if (K < 0 || (K > 0 && (!probs || !J || !q))) return fail(GW_E_INVALID, "bad arguments");
    if (K == 0) return GW_OK;
    if (K >= ((int64_t)1 << 31)) return fail(GW_E_TOO_LARGE, "table too large");
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0)
        return fail(GW_E_CUDA, "no CUDA device available; libgraphwalk has no CPU fallback");
    DevBuf<double> dp, dq;
    DevBuf<int32_t> dJ;
    GW_CUDA(dp.alloc((size_t)K)); GW_CUDA(dq.alloc((size_t)K)); GW_CUDA(dJ.alloc((size_t)K));
    GW_CUDA(cudaMemcpy(dp.p, probs, sizeof(double) * (size_t)K, cudaMemcpyHostToDevice));
    k_alias_single<<<1, 32>>>(dp.p, (int32_t)K, dJ.p, dq.p);
    GW_LAUNCHED();
    GW_CUDA(cudaMemcpy(J, dJ.p, sizeof(int32_t) * (size_t)K, cudaMemcpyDeviceToHost));
    GW_CUDA(cudaMemcpy(q, dq.p, sizeof(double) * (size_t)K, cudaMemcpyDeviceToHost));
    return GW_OK;
}

int gw_alias_nodes(gw_graph *g, int32_t *J, double *q) {
    if (!g) return fail(GW_E_INVALID, "graph is NULL");
    if (g->flags & GW_F_MULTI) return fail(GW_E_STATE, "alias tables need a SIMPLE-mode (sorted) graph");
    GW_CUDA(cudaSetDevice(g->device));
    if (!g->d_anJ) {
        size_t cnt = (size_t)std::max<int64_t>(g->nnz, 1);
        GW_CUDA(cudaMalloc((void **)&g->d_anJ, sizeof(int32_t) * cnt));
        GW_CUDA(cudaMalloc((void **)&g->d_anq, sizeof(double) * cnt));
        if (g->n > 0) {
            k_alias_nodes<<<(unsigned)((g->n + 127) / 128), 128>>>(g->d_meta, g->d_w, g->n, g->d_anJ, g->d_anq);
            GW_LAUNCHED();
        }
        GW_CUDA(cudaDeviceSynchronize());
    }
    if (J && g->nnz) GW_CUDA(cudaMemcpy(J, g->d_anJ, sizeof(int32_t) * (size_t)g->nnz, cudaMemcpyDeviceToHost));
    if (q && g->nnz) GW_CUDA(cudaMemcpy(q, g->d_anq, sizeof(double) * (size_t)g->nnz, cudaMemcpyDeviceToHost));
    return GW_OK;
}

static int build_ae_offsets(gw_graph *g, DevBuf<int64_t> &off, int64_t *total) {
    GW_CUDA(off.alloc((size_t)(g->nnz + 1)));
    k_entry_degree<<<(unsigned)((g->nnz + 1 + 255) / 256), 256>>>(g->d_meta, g->d_col, g->nnz, off.p);
    GW_LAUNCHED();
    size_t tb = 0;
    GW_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, off.p, off.p, g->nnz + 1));
    DevBuf<uint8_t> tmp;
    GW_CUDA(tmp.alloc(tb));
    GW_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, off.p, off.p, g->nnz + 1));
    g_launches.fetch_add(1);
    GW_CUDA(cudaMemcpy(total, off.p + g->nnz, sizeof(int64_t), cudaMemcpyDeviceToHost));
    return GW_OK;
}

int gw_alias_edges_size(const gw_graph *g, int64_t *total) {
    if (!g || !total) return fail(GW_E_INVALID, "graph/total is NULL");
    GW_CUDA(cudaSetDevice(g->device));
    if (g->d_aeoff) { *total = g->ae_total; return GW_OK; }
    DevBuf<int64_t> off;
    return build_ae_offsets(const_cast<gw_graph *>(g), off, total);
}

int gw_alias_edges(gw_graph *g, double p, double q, int64_t budget_bytes, int64_t *off, int32_t *J, double *qv) {
    if (!g) return fail(GW_E_INVALID, "graph is NULL");
    if (g->flags & GW_F_MULTI) return fail(GW_E_STATE, "alias tables need a SIMPLE-mode (sorted) graph");
    if (!(p > 0) || !(q > 0)) return fail(GW_E_INVALID, "p and q must be positive");
    GW_CUDA(cudaSetDevice(g->device));
    if (!g->d_aeoff || g->ae_p != p || g->ae_q != q) {
        const bool keep = g->d_aeoff && g->d_aeJ && g->d_aeq;      // same graph: same sizes, only p/q changed
        if (!keep) {
            cudaFree(g->d_aeoff); cudaFree(g->d_aeJ); cudaFree(g->d_aeq);
            g->d_aeoff = nullptr; g->d_aeJ = nullptr; g->d_aeq = nullptr;
        }
        DevBuf<int64_t> o;
        int64_t total = 0;
        if (keep) { o.p = g->d_aeoff; g->d_aeoff = nullptr; total = g->ae_total; }
        else GW_TRY(build_ae_offsets(g, o, &total));
        size_t freeb = 0;
        GW_TRY(device_info(nullptr, &freeb));
        int64_t need = total * 12;
        int64_t budget = budget_bytes > 0 ? budget_bytes : (int64_t)(freeb / 4 * 3);
        if (!keep && need > budget)
            return fail(GW_E_TOO_LARGE, "alias_edges needs %lld entries (%lld bytes) > budget %lld bytes; use the "
                        "free-running walker, which evaluates the p/q bias on the fly",
                        (long long)total, (long long)need, (long long)budget);
        if (!keep) {
            GW_CUDA(cudaMalloc((void **)&g->d_aeJ, sizeof(int32_t) * (size_t)std::max<int64_t>(total, 1)));
            GW_CUDA(cudaMalloc((void **)&g->d_aeq, sizeof(double) * (size_t)std::max<int64_t>(total, 1)));
        }
        if (g->nnz > 0) {
            cudaEvent_t ev0, ev1, ev2;
            cudaEventCreate(&ev0); cudaEventCreate(&ev1); cudaEventCreate(&ev2);
            cudaEventRecord(ev0);
            // entries sorted by decreasing table length (nnz < 2^32: 32-bit entry ids)
            DevBuf<uint32_t> k0, k1, v0, v1;
            DevBuf<uint8_t> tmp;
            const size_t cnt = (size_t)g->nnz;
            GW_CUDA(k0.alloc(cnt)); GW_CUDA(k1.alloc(cnt)); GW_CUDA(v0.alloc(cnt)); GW_CUDA(v1.alloc(cnt));
            k_entry_keys<<<(unsigned)((g->nnz + 255) / 256), 256>>>(g->d_meta, g->d_col, g->nnz, k0.p, v0.p);
            GW_LAUNCHED();
            size_t tb = 0;
            GW_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, k0.p, k1.p, v0.p, v1.p, (int64_t)cnt));
            GW_CUDA(tmp.alloc(tb));
            GW_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp.p, tb, k0.p, k1.p, v0.p, v1.p, (int64_t)cnt));
            g_launches.fetch_add(1);
            // size classes over the sorted order: [0,c0) longer than any shared-memory stage -> one thread per table;
            // (8192,18000] / (4096,8192] / ... / (32,64] -> one CTA per table, 12*kcap bytes of shared memory;
            // <= 32 -> one thread per table (tiny private tables stay in L1)
            std::vector<uint32_t> hk(cnt);
            GW_CUDA(cudaMemcpy(hk.data(), k1.p, sizeof(uint32_t) * cnt, cudaMemcpyDeviceToHost));
            cudaEventRecord(ev1);
            auto first_le = [&](uint32_t bound) {                 // first position whose length is <= bound (descending keys)
                return (int64_t)(std::partition_point(hk.begin(), hk.end(), [&](uint32_t x) { return x > bound; }) - hk.begin());
            };
            constexpr int NCLS = 9;
            const int32_t caps[NCLS] = {18000, 8192, 4096, 2048, 1024, 512, 256, 128, 64};   // finer classes = more tables resident per SM
            int64_t pos = first_le((uint32_t)caps[0]);
            if (pos > 0) {                                          // hubs beyond the largest stage
                k_alias_edges<<<(unsigned)((pos + 127) / 128), 128>>>(g->d_meta, g->d_col, g->d_w, g->d_row_ptr, g->n, pos, p, q, o.p,
                                                                      v1.p, g->d_aeJ, g->d_aeq);
                GW_LAUNCHED();
            }
            int sms = 148;
            device_info(&sms, nullptr);
            for (int c = 0; c < NCLS; c++) {
                const int64_t end = first_le(c + 1 < NCLS ? (uint32_t)caps[c + 1] : 32u);
                if (end > pos) {
                    const size_t smem = 12 * (size_t)caps[c] + 16;
                    const int threads = caps[c] >= 1024 ? 128 : (caps[c] >= 256 ? 64 : 32);
                    GW_CUDA(cudaFuncSetAttribute(k_alias_edges_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    const int per_sm = (int)std::min<size_t>(32, (size_t)(220 * 1024) / (smem + 1024));
                    const unsigned grid = (unsigned)std::min<int64_t>(end - pos, (int64_t)sms * std::max(per_sm, 1));
                    k_alias_edges_warp<<<grid, threads, smem>>>(g->d_meta, g->d_col, g->d_w, g->d_row_ptr, g->n, p, q, o.p, v1.p, pos,
                                                           end - pos, caps[c], g->d_aeJ, g->d_aeq);
                    GW_LAUNCHED();
                }
                pos = std::max(pos, end);
            }
            if ((int64_t)cnt > pos) {                               // tiny tables
                k_alias_edges<<<(unsigned)(((int64_t)cnt - pos + 127) / 128), 128>>>(g->d_meta, g->d_col, g->d_w, g->d_row_ptr, g->n,
                                                                                   (int64_t)cnt - pos, p, q, o.p, v1.p + pos, g->d_aeJ,
                                                                                   g->d_aeq);
                GW_LAUNCHED();
            }
            cudaEventRecord(ev2);
            GW_CUDA(cudaDeviceSynchronize());      // the sort buffers die with this scope
            if (getenv("GW_TIMING")) {
                float a = 0, b = 0;
                cudaEventElapsedTime(&a, ev0, ev1); cudaEventElapsedTime(&b, ev1, ev2);
                fprintf(stderr, "gw_alias_edges: sort+classes %.2f ms, table kernels %.2f ms, %lld entries\n", a, b, (long long)total);
            }
            cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev2);
        }
        GW_CUDA(cudaDeviceSynchronize());
        g->d_aeoff = o.take();
        g->ae_total = total; g->ae_p = p; g->ae_q = q;
    }
    if (off) GW_CUDA(cudaMemcpy(off, g->d_aeoff, sizeof(int64_t) * (size_t)(g->nnz + 1), cudaMemcpyDeviceToHost));
    if (J && g->ae_total) GW_CUDA(cudaMemcpy(J, g->d_aeJ, sizeof(int32_t) * (size_t)g->ae_total, cudaMemcpyDeviceToHost));
    if (qv && g->ae_total) GW_CUDA(cudaMemcpy(qv, g->d_aeq, sizeof(double) * (size_t)g->ae_total, cudaMemcpyDeviceToHost));
    return GW_OK;
}

}  // extern "C"
