// slab_probe.cu — go / no-go experiment for a step-synchronous, slab-binned walker.
//
// Question: a dependent random 16-byte gather over a 2 GiB array runs at ~49 G accesses/s whatever the load flavour
// (profiles/README.md section 2).  If that ceiling is address translation (TLB reach ~256 MiB), a walker that first
// BINS its live walkers by the 64..512 MiB slab of nbr4 they will touch, and then gathers slab by slab, should run the
// gather phase several times faster -- at the price of one radix pass over the walker records per step.
//
// What is measured (W walkers = one step of the R-MAT-22 workload, E entries of 16 bytes):
//   A. gather of W INDEPENDENT random entries, index list in random order          (the per-step gather as it is today)
//   B. the same gather with the index list sorted by slab (random inside a slab)   (what binning buys the gather phase)
//   C. the same gather fully sorted by address                                     (upper bound: DRAM page locality too)
//   D. the binning pass itself: histogram + scan + scatter of W 8-byte {walker, index} records into B bins
//   E. a whole emulated step loop: T steps of [gather in bin order -> next index from the loaded entry -> re-bin],
//      against T steps of the plain dependent chain, same number of accesses.
// Net verdict = E_binned vs E_chain.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/slab_probe tools/slab_probe.cu
#include <cuda_runtime.h>
#include <cassert>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __host__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

__device__ __forceinline__ int4 ld64(const int4 *p) {
    int4 v;
    asm volatile("ld.global.nc.L2::64B.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__global__ void k_fill(int4 *a, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = make_int4((int)mix(i * 2654435761u + 1), (int)i, 0, 0);
}

// A/B/C: one independent load per thread, index list given
__global__ void k_gather(const int4 *__restrict__ a, const uint32_t *__restrict__ idx, uint32_t n, uint32_t salt, uint32_t mask,
                         uint32_t *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int4 v = ld64(a + ((idx[i] + salt) & mask));              // salt rotates the address range by whole 256 MiB: every rep is cold
    out[i] = (uint32_t)v.x + (uint32_t)v.y;
}

// plain dependent chain (today's walker): T steps per thread
__global__ void k_chain(const int4 *__restrict__ a, uint32_t mask, int steps, uint32_t n, uint32_t *out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    uint32_t i = mix(t) & mask, acc = 0;
    for (int s = 0; s < steps; s++) {
        int4 v = ld64(a + i);
        acc += v.y;
        i = mix((uint32_t)v.x + s) & mask;
    }
    out[t] = acc;
}

// ---- binned step: records {walker, index}; gather in record order, produce the next index, count its bin ----
constexpr int MAXB = 64;
constexpr int ITEMS = 4;                     // records per thread: 4096-record tiles keep the scan input small
__global__ void __launch_bounds__(1024) k_step_gather(const int4 *__restrict__ a, uint2 *__restrict__ rec, uint32_t n, uint32_t mask, int s,
                                                       int shift, int nbins, uint32_t *__restrict__ hist /*[grid][nbins]*/,
                                                       uint32_t *__restrict__ outT /*[n] this step's column*/) {
    __shared__ uint32_t sh[MAXB];
    if (threadIdx.x < MAXB) sh[threadIdx.x] = 0;
    __syncthreads();
    uint2 r[ITEMS];
    int4 v[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        uint32_t i = (blockIdx.x * ITEMS + k) * 1024 + threadIdx.x;
        if (i < n) { r[k] = rec[i]; assert(r[k].y <= mask && r[k].x < n); v[k] = ld64(a + r[k].y); }
    }
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        uint32_t i = (blockIdx.x * ITEMS + k) * 1024 + threadIdx.x;
        if (i < n) {
            uint32_t nxt = mix((uint32_t)v[k].x + s) & mask;
            outT[r[k].x] = (uint32_t)v[k].y;                    // corpus column of this step (transposed layout), scattered 4-byte store
            rec[i] = make_uint2(r[k].x, nxt);
            assert((nxt >> shift) < (uint32_t)nbins);
            atomicAdd(&sh[nxt >> shift], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < nbins) hist[blockIdx.x * nbins + threadIdx.x] = sh[threadIdx.x];
}
// exclusive scan over hist laid out [grid][nbins] in bin-major order (bin b of all CTAs, then bin b+1): one CTA
__global__ void __launch_bounds__(1024) k_scan(uint32_t *__restrict__ hist, int grid, int nbins) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int total = grid * nbins;
    for (int base = 0; base < total; base += 1024) {
        int j = base + threadIdx.x;                              // j enumerates (bin, cta) bin-major
        uint32_t v = 0;
        if (j < total) { int b = j / grid, c = j % grid; v = hist[c * nbins + b]; }
        uint32_t inc = v;
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if ((threadIdx.x & 31) >= o) inc += t; }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = warp_tot[threadIdx.x], wi = w;
            for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, wi, o); if (threadIdx.x >= o) wi += t; }
            warp_tot[threadIdx.x] = wi - w;
        }
        __syncthreads();
        uint32_t excl = carry + warp_tot[threadIdx.x >> 5] + inc - v;
        if (j < total) { int b = j / grid, c = j % grid; hist[c * nbins + b] = excl; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(1024) k_scatter(const uint2 *__restrict__ rec, uint2 *__restrict__ dst, uint32_t n, int shift, int nbins,
                                                   const uint32_t *__restrict__ hist) {
    __shared__ uint32_t cur[MAXB];
    if (threadIdx.x < nbins) cur[threadIdx.x] = hist[blockIdx.x * nbins + threadIdx.x];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        uint32_t i = (blockIdx.x * ITEMS + k) * 1024 + threadIdx.x;
        if (i < n) {
            uint2 r = rec[i];
            assert((r.y >> shift) < (uint32_t)nbins);
            uint32_t pos = atomicAdd(&cur[r.y >> shift], 1u);
            assert(pos < n);
            dst[pos] = r;
        }
    }
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main(int argc, char **argv) {
    const uint32_t E = 1u << 27;                 // 2 GiB of int4
    const uint32_t W = argc > 1 ? (uint32_t)atoi(argv[1]) : 4178039u;   // walkers of the R-MAT-22 pass
    const int T = argc > 2 ? atoi(argv[2]) : 79;
    int4 *a;
    CK(cudaMalloc(&a, (size_t)E * 16));
    k_fill<<<(E + 255) / 256, 256>>>(a, E);
    uint32_t *d_idx, *d_out;
    CK(cudaMalloc(&d_idx, (size_t)W * 4));
    CK(cudaMalloc(&d_out, (size_t)W * 4 * 2));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<uint32_t> h(W);
    printf("E = %u entries (%.0f MiB), W = %u accesses per step, T = %d steps\n", E, E * 16.0 / (1 << 20), W, T);

    // ---- A / B / C ----
    auto run_gather = [&](const char *name) {
        CK(cudaMemcpy(d_idx, h.data(), (size_t)W * 4, cudaMemcpyHostToDevice));
        float best = 1e9;
        for (int rep = 0; rep < 5; rep++) {
            CK(cudaEventRecord(e0));
            k_gather<<<(W + 255) / 256, 256>>>(a, d_idx, W, (uint32_t)rep * (E / 8) + 4099u * 16u * (uint32_t)rep, E - 1, d_out);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            best = std::min(best, time_ms(e0, e1));
        }
        printf("  %-44s %8.1f us  %7.1f G loads/s\n", name, best * 1e3, W / (best * 1e-3) / 1e9);
    };
    for (uint32_t i = 0; i < W; i++) h[i] = mix(i * 7919u + 13) & (E - 1);
    printf("independent gather, one load per thread (each rep touches cold entries: 3 %% of the array per step)\n");
    run_gather("A random order");
    for (int mib : {1024, 512, 256, 128, 64, 32}) {
        const uint32_t per = (uint32_t)((size_t)mib << 20) / 16;           // entries per slab
        std::vector<uint32_t> s = h;
        std::stable_sort(s.begin(), s.end(), [&](uint32_t x, uint32_t y) { return x / per < y / per; });
        std::swap(s, h);
        char nm[64];
        snprintf(nm, sizeof nm, "B binned by %4d MiB slab (%3u bins)", mib, (E + per - 1) / per);
        run_gather(nm);
        std::swap(s, h);
    }
    {
        std::vector<uint32_t> s = h;
        std::sort(h.begin(), h.end());
        run_gather("C fully sorted by address");
        h = s;
    }

    // ---- E: whole step loops ----
    printf("step loops, %d steps x %u walkers (dependent: the next index comes out of the loaded entry)\n", T, W);
    {
        float best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(e0));
            k_chain<<<(W + 255) / 256, 256>>>(a, E - 1, T, W, d_out);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            best = std::min(best, time_ms(e0, e1));
        }
        printf("  %-44s %8.1f us/step  %7.1f G steps/s\n", "chain (one thread per walker, today)", best * 1e3 / T, (double)W * T / (best * 1e-3) / 1e9);
    }
    uint2 *rec[2];
    CK(cudaMalloc(&rec[0], (size_t)W * 8)); CK(cudaMalloc(&rec[1], (size_t)W * 8));
    uint32_t *d_hist, *d_outT;
    const int grid = (int)((W + 1024 * ITEMS - 1) / (1024 * ITEMS));
    CK(cudaMalloc(&d_hist, (size_t)grid * MAXB * 4));
    CK(cudaMalloc(&d_outT, (size_t)W * 4 * 8));                  // 8 columns of the transposed corpus, reused round-robin
    std::vector<uint2> hr(W);
    for (int mib : {512, 256, 128, 64}) {
        const uint32_t per = (uint32_t)((size_t)mib << 20) / 16;
        const int nbins = (int)((E + per - 1) / per);
        int shift = 0;
        while ((1u << shift) < per) shift++;
        if (nbins > MAXB) continue;
        for (uint32_t i = 0; i < W; i++) hr[i] = make_uint2(i, mix(i) & (E - 1));
        std::stable_sort(hr.begin(), hr.end(), [&](uint2 x, uint2 y) { return (x.y >> shift) < (y.y >> shift); });
        CK(cudaMemcpy(rec[0], hr.data(), (size_t)W * 8, cudaMemcpyHostToDevice));
        cudaStream_t st;
        CK(cudaStreamCreate(&st));
        {   // one step outside the graph first: a failing assert then names its line
            k_step_gather<<<grid, 1024, 0, st>>>(a, rec[0], W, E - 1, 0, shift, nbins, d_hist, d_outT);
            CK(cudaStreamSynchronize(st));
            k_scan<<<1, 1024, 0, st>>>(d_hist, grid, nbins);
            CK(cudaStreamSynchronize(st));
            k_scatter<<<grid, 1024, 0, st>>>(rec[0], rec[1], W, shift, nbins, d_hist);
            CK(cudaStreamSynchronize(st));
            printf("  (one un-captured step ran, %d bins)\n", nbins);
        }
        cudaGraph_t graph;
        cudaGraphExec_t exec;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
        int cur = 0;
        for (int s = 0; s < T; s++) {
            k_step_gather<<<grid, 1024, 0, st>>>(a, rec[cur], W, E - 1, s, shift, nbins, d_hist, d_outT + (size_t)(s & 7) * W);
            k_scan<<<1, 1024, 0, st>>>(d_hist, grid, nbins);
            k_scatter<<<grid, 1024, 0, st>>>(rec[cur], rec[cur ^ 1], W, shift, nbins, d_hist);
            cur ^= 1;
        }
        CK(cudaStreamEndCapture(st, &graph));
        CK(cudaGraphInstantiate(&exec, graph, 0));
        float best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaMemcpy(rec[0], hr.data(), (size_t)W * 8, cudaMemcpyHostToDevice));
            CK(cudaEventRecord(e0, st));
            CK(cudaGraphLaunch(exec, st));
            CK(cudaEventRecord(e1, st));
            CK(cudaEventSynchronize(e1));
            best = std::min(best, time_ms(e0, e1));
        }
        // phase split: the gather + histogram kernel alone (10 steps back to back, records re-binned by nobody: after the
        // first of them the list is in random order, so this is the UNBINNED gather kernel); binning = step - binned gather
        CK(cudaMemcpy(rec[0], hr.data(), (size_t)W * 8, cudaMemcpyHostToDevice));
        CK(cudaEventRecord(e0, st));
        k_step_gather<<<grid, 1024, 0, st>>>(a, rec[0], W, E - 1, 0, shift, nbins, d_hist, d_outT);
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        const float g_us = time_ms(e0, e1) * 1e3;               // ONE gather over a binned list (cold)
        const float b_us = best * 1e3 / T - g_us;               // what is left of a step: scan + scatter (+ launch gaps inside the graph)
        char nm[96];
        snprintf(nm, sizeof nm, "binned, %4d MiB slabs (%2d bins), CUDA graph", mib, nbins);
        printf("  %-44s %8.1f us/step  %7.1f G steps/s   [binned gather+hist kernel %.1f us; scan + scatter = the rest: %.1f us]\n", nm,
               best * 1e3 / T, (double)W * T / (best * 1e-3) / 1e9, g_us, b_us);
        cudaGraphExecDestroy(exec); cudaGraphDestroy(graph); cudaStreamDestroy(st);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
