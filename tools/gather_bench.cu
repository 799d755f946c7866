// Random-gather microbenchmark: how many independent random 16-byte (or 8-byte) loads per second
// does one B200 sustain, per load flavour and L2 fetch granularity?  Sets the practical ceiling of
// the walkers (one dependent random sector per step).   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int MODE>
__device__ __forceinline__ int4 load16(const int4 *p, uint64_t pol) {
    int4 v;
    if (MODE == 0) v = *p;                                   // ld.global
    else if (MODE == 1) v = __ldg(p);                        // ld.global.nc
    else if (MODE == 2) v = __ldcg(p);                       // ld.global.cg (L2 only)
    else if (MODE == 3) v = __ldcs(p);                       // ld.global.cs (streaming)
    else if (MODE == 4) asm volatile("ld.global.nc.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    else if (MODE == 5) asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (MODE == 6) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    else if (MODE == 7) asm volatile("ld.global.nc.L2::64B.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (MODE == 8) asm volatile("ld.global.nc.L2::128B.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (MODE == 9) asm volatile("ld.global.nc.L2::256B.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (MODE == 10) asm volatile("ld.global.L1::no_allocate.L2::64B.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else asm volatile("ld.volatile.global.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// dependent chain: next index comes from the loaded value (like a walk)
template <int MODE>
__global__ void k_chain(const int4 *__restrict__ a, uint32_t mask, int steps, uint32_t *out) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    uint32_t i = mix(blockIdx.x * blockDim.x + threadIdx.x) & mask;
    uint32_t acc = 0;
    for (int s = 0; s < steps; s++) {
        int4 v = load16<MODE>(a + i, pol);
        acc += v.y;
        i = mix((uint32_t)v.x + s) & mask;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// dependent chain whose trajectory is the thread's own: two threads that meet on an entry part again (k_chain's merge for good,
// which turns long runs into same-address hot spots)
__global__ void k_chain_own(const int4 *__restrict__ a, uint32_t mask, int steps, uint32_t salt, uint32_t *out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t i = mix(t * 2654435761u + salt) & mask;
    uint32_t acc = 0;
    for (int s = 0; s < steps; s++) {
        const int4 v = __ldg(a + i);
        acc += v.y;
        i = mix((uint32_t)v.x + s + t * 0x9E3779B9u) & mask;
    }
    out[t] = acc;
}

// single-wave experiments: thread-id offset (is it the ids?) and a per-block start delay (is it the lock step?)
__global__ void k_chain_var(const int4 *__restrict__ a, uint32_t mask, int steps, uint32_t salt, uint32_t toff, int delay_us, uint32_t *out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x + toff;
    if (delay_us) {
        const long long until = clock64() + (long long)(mix(blockIdx.x * 7919u + salt) % (uint32_t)delay_us) * 1900;
        while (clock64() < until) { }
    }
    uint32_t i = mix(t * 2654435761u + salt) & mask;
    uint32_t acc = 0;
    for (int s = 0; s < steps; s++) {
        const int4 v = __ldg(a + i);
        acc += v.y;
        i = mix((uint32_t)v.x + s + t * 0x9E3779B9u) & mask;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// do CTAs that retire at once count as "turnover"?  The first `ndummy` blocks of the launch return immediately
__global__ void k_chain_dummy(const int4 *__restrict__ a, uint32_t mask, int steps, uint32_t salt, uint32_t ndummy, uint32_t *out) {
    if (blockIdx.x < ndummy) return;
    const uint32_t t = (blockIdx.x - ndummy) * blockDim.x + threadIdx.x;
    uint32_t i = mix(t * 2654435761u + salt) & mask;
    uint32_t acc = 0;
    for (int s = 0; s < steps; s++) {
        const int4 v = __ldg(a + i);
        acc += v.y;
        i = mix((uint32_t)v.x + s + t * 0x9E3779B9u) & mask;
    }
    out[t] = acc;
}

// do SHORT working waves count?  The first `nprime` blocks walk `psteps` loads, the others `steps`
__global__ void k_chain_prime(const int4 *__restrict__ a, uint32_t mask, int steps, int psteps, uint32_t salt, uint32_t nprime, uint32_t *out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const int my = blockIdx.x < nprime ? psteps : steps;
    uint32_t i = mix(t * 2654435761u + salt) & mask;
    uint32_t acc = 0;
    for (int s = 0; s < my; s++) {
        const int4 v = __ldg(a + i);
        acc += v.y;
        i = mix((uint32_t)v.x + s + t * 0x9E3779B9u) & mask;
    }
    out[t] = acc;
}

// ... or is it PENDING CTAs that matter?  The LAST `ntail` blocks of the launch return at once: they stay undispatched while the
// working wave holds every CTA slot
__global__ void k_chain_tail(const int4 *__restrict__ a, uint32_t mask, int steps, uint32_t salt, uint32_t nwork, uint32_t *out) {
    if (blockIdx.x >= nwork) return;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t i = mix(t * 2654435761u + salt) & mask;
    uint32_t acc = 0;
    for (int s = 0; s < steps; s++) {
        const int4 v = __ldg(a + i);
        acc += v.y;
        i = mix((uint32_t)v.x + s + t * 0x9E3779B9u) & mask;
    }
    out[t] = acc;
}

// when, inside a many-wave launch, is the rate high?  Every block records its start and end on the global timer
__global__ void k_chain_times(const int4 *__restrict__ a, uint32_t mask, int steps, uint32_t salt, uint32_t *out, unsigned long long *t0, unsigned long long *t1,
                              uint32_t *sm) {
    unsigned long long now;
    if (threadIdx.x == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now)); t0[blockIdx.x] = now; uint32_t id; asm volatile("mov.u32 %0, %%smid;" : "=r"(id)); sm[blockIdx.x] = id; }
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t i = mix(t * 2654435761u + salt) & mask;
    uint32_t acc = 0;
    for (int s = 0; s < steps; s++) {
        const int4 v = __ldg(a + i);
        acc += v.y;
        i = mix((uint32_t)v.x + s + t * 0x9E3779B9u) & mask;
    }
    out[t] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now)); t1[blockIdx.x] = now; }
}

// pair mode: every step loads the random entry AND its partner at byte distance `dist` (same naturally
// aligned 2*dist block).  If pairs run at the single-load rate the ceiling is a DRAM activate / L2-miss
// REQUEST rate that locality can amortise, not bytes.
__global__ void k_pair(const int4 *__restrict__ a, uint32_t mask, int steps, uint32_t dist16, uint32_t *out) {
    uint32_t i = mix(blockIdx.x * blockDim.x + threadIdx.x) & mask;
    uint32_t acc = 0;
    for (int s = 0; s < steps; s++) {
        int4 v = __ldg(a + i);
        int4 w = __ldg(a + (i ^ dist16));
        acc += v.y + w.y;
        i = mix((uint32_t)v.x + s) & mask;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// strided-footprint mode: random entries, but only one 128-byte line out of every `stride` lines is ever
// touched: footprint = n/stride lines (L2 resident when small) spread over the whole address range (every
// 2 MB page is touched).  Separates "L2 capacity" from "TLB reach".
__global__ void k_strided(const int4 *__restrict__ a, uint32_t mask_lines, uint32_t stride_lines, int steps, uint32_t *out) {
    uint32_t i = mix(blockIdx.x * blockDim.x + threadIdx.x) & mask_lines;
    uint32_t acc = 0;
    for (int s = 0; s < steps; s++) {
        int4 v = __ldg(a + (size_t)i * stride_lines * 8);
        acc += v.y;
        i = mix((uint32_t)v.x + s + acc) & mask_lines;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
__global__ void k_chain_mod(const int4 *__restrict__ a, uint32_t n, int steps, uint32_t *out) {
    uint32_t i = mix(blockIdx.x * blockDim.x + threadIdx.x) % n;
    uint32_t acc = 0;
    for (int s = 0; s < steps; s++) {
        int4 v = load16<MODE>(a + i, 0);
        acc += v.y;
        i = __umulhi(mix((uint32_t)v.x + s), n);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void k_fill(int4 *a, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) a[i] = make_int4((int)mix((uint32_t)i), (int)i, 0, 0);
}

template <int MODE>
float run(const int4 *a, uint32_t mask, int nthreads, int steps, uint32_t *out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_chain<MODE><<<nthreads / 256, 256>>>(a, mask, steps, out);
    cudaEventRecord(e0);
    k_chain<MODE><<<nthreads / 256, 256>>>(a, mask, steps, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

// 4-byte loads, ILP independent chains per thread
template <int ILP>
__global__ void k_chain4(const uint32_t *__restrict__ a, uint32_t mask, int steps, uint32_t *out) {
    uint32_t i[ILP], acc = 0;
    for (int j = 0; j < ILP; j++) i[j] = mix((blockIdx.x * blockDim.x + threadIdx.x) * ILP + j) & mask;
    for (int s = 0; s < steps; s++) {
        uint32_t v[ILP];
#pragma unroll
        for (int j = 0; j < ILP; j++) v[j] = __ldg(a + i[j]);
#pragma unroll
        for (int j = 0; j < ILP; j++) { acc += v[j]; i[j] = mix(v[j] + s) & mask; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_fill4(uint32_t *a, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) a[i] = mix((uint32_t)i);
}
template <int ILP>
float run4(const uint32_t *a, uint32_t mask, int nthreads, int steps, uint32_t *out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_chain4<ILP><<<nthreads / 256, 256>>>(a, mask, steps, out);
    cudaEventRecord(e0);
    k_chain4<ILP><<<nthreads / 256, 256>>>(a, mask, steps, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main(int argc, char **argv) {
    if (argc > 2) {   // mode 2: 4-byte random loads over arrays of several sizes, ILP 1 and 4
        uint32_t *b, *out4; int nthreads = 1 << 22, steps = 80;
        cudaMalloc(&out4, nthreads * 4);
        for (int lg = 24; lg <= 30; lg += 2) {           // 64 MiB .. 4 GiB of uint32
            size_t n4 = (size_t)1 << lg;
            cudaMalloc(&b, n4 * 4);
            k_fill4<<<(unsigned)((n4 + 255) / 256), 256>>>(b, n4);
            float m1 = run4<1>(b, (uint32_t)(n4 - 1), nthreads, steps, out4);
            float m4 = run4<4>(b, (uint32_t)(n4 - 1), nthreads / 4, steps, out4);
            printf("4-byte random loads, array %6.0f MiB: ILP1 %7.3f ms %6.1f G loads/s | ILP4 %7.3f ms %6.1f G loads/s\n",
                   n4 * 4.0 / (1 << 20), m1, (double)nthreads * steps / m1 / 1e6, m4, (double)nthreads * steps / m4 / 1e6);
            cudaFree(b);
        }
        printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        return 0;
    }
    if (argc > 1 && atoi(argv[1]) == -4) {   // mode 6: dense arrays on a fine size grid, default vs L2::64B loads
        uint32_t *out; int nthreads = 1 << 22, steps = 80;
        cudaMalloc(&out, nthreads * 4);
        int sizes[] = {64, 96, 128, 192, 256, 320, 384, 512, 768, 1024, 1536, 2048, 4096};
        for (int si = 0; si < 13; si++) {
            size_t n = ((size_t)sizes[si] << 20) / 16; int4 *a;
            cudaMalloc(&a, n * sizeof(int4));
            k_fill<<<(unsigned)((n + 255) / 256), 256>>>(a, n);
            // non power-of-two sizes: the chain kernels mask with the next power of two minus one, so use a
            // modulo variant here
            float m0, m1;
            {
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                k_chain_mod<1><<<nthreads / 256, 256>>>(a, (uint32_t)n, steps, out);
                cudaEventRecord(e0);
                k_chain_mod<1><<<nthreads / 256, 256>>>(a, (uint32_t)n, steps, out);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                cudaEventElapsedTime(&m0, e0, e1);
                k_chain_mod<7><<<nthreads / 256, 256>>>(a, (uint32_t)n, steps, out);
                cudaEventRecord(e0);
                k_chain_mod<7><<<nthreads / 256, 256>>>(a, (uint32_t)n, steps, out);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                cudaEventElapsedTime(&m1, e0, e1);
            }
            printf("dense %5d MiB: default %7.3f ms %6.1f G loads/s | L2::64B %7.3f ms %6.1f G loads/s\n", sizes[si], m0,
                   (double)nthreads * steps / m0 / 1e6, m1, (double)nthreads * steps / m1 / 1e6);
            cudaFree(a);
        }
        printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        return 0;
    }
    if (argc > 1 && atoi(argv[1]) == -2) {   // mode 4: fixed 64 MiB footprint spread over growing address ranges
        uint32_t *out; int nthreads = 1 << 22, steps = 80;
        cudaMalloc(&out, nthreads * 4);
        for (int lg = 22; lg <= 29; lg++) {          // array of 2^lg int4 = 64 MiB .. 8 GiB
            size_t n = (size_t)1 << lg; int4 *a;
            if (cudaMalloc(&a, n * sizeof(int4)) != cudaSuccess) break;
            k_fill<<<(unsigned)((n + 255) / 256), 256>>>(a, n);
            uint32_t lines = (uint32_t)(n / 8), foot_lines = 1u << 19;    // 2^19 lines x 128 B = 64 MiB touched
            uint32_t stride = lines / foot_lines;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            k_strided<<<nthreads / 256, 256>>>(a, foot_lines - 1, stride, steps, out);
            cudaEventRecord(e0);
            k_strided<<<nthreads / 256, 256>>>(a, foot_lines - 1, stride, steps, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("64 MiB footprint over %6.0f MiB (stride %4u lines): %7.3f ms  %6.1f G loads/s\n", n * 16.0 / (1 << 20), stride, ms, (double)nthreads * steps / ms / 1e6);
            cudaFree(a);
        }
        printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        return 0;
    }
    if (argc > 1 && atoi(argv[1]) == -5) {   // mode 7: chain length x occupancy, chains that cannot merge (the thread id enters every index)
        size_t n = (size_t)1 << 27; int4 *a; uint32_t *out;
        cudaMalloc(&a, n * sizeof(int4)); cudaMalloc(&out, 148 * 2048 * 4);
        uint32_t *out2; cudaMalloc(&out2, (size_t)4 << 24);
        k_fill<<<(unsigned)((n + 255) / 256), 256>>>(a, n);
        const int lens[] = {10, 80, 400, 2000};
        for (int tpsm = 512; tpsm <= 2048; tpsm *= 2)
            for (int li = 0; li < 4; li++) {
                const int steps = lens[li], reps = 4000 / steps > 0 ? 4000 / steps : 1;
                const int grid = 148 * (tpsm / 256);
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                k_chain_own<<<grid, 256>>>(a, (uint32_t)(n - 1), steps, 0, out);
                cudaEventRecord(e0);
                for (int r = 0; r < reps; r++) k_chain_own<<<grid, 256>>>(a, (uint32_t)(n - 1), steps, 1 + r, out);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                printf("%4d threads/SM, %4d loads per chain, %3d launches: %6.1f G loads/s\n", tpsm, steps, reps, (double)grid * 256 * steps * reps / ms / 1e6);
            }
        // the same chains as MANY waves of blocks in one launch (blocks retire and start all the time: the warps of an SM
        // are at different steps), and with 2.5 GiB (the BA n = 1e7 m = 8 nbr4 array)
        for (int li = 0; li < 3; li++) {
            const int steps = lens[li + 1];
            const int grid = (1 << 22) / 256 * (li == 0 ? 4 : 1);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            k_chain_own<<<grid, 256>>>(a, (uint32_t)(n - 1), steps, 0, out2);
            cudaEventRecord(e0);
            k_chain_own<<<grid, 256>>>(a, (uint32_t)(n - 1), steps, 1, out2);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("%d blocks of 256 in one launch, %4d loads per chain: %6.1f G loads/s\n", grid, steps, (double)grid * 256 * steps / ms / 1e6);
        }
        {
            struct { uint32_t toff; int delay; int steps; const char *what; } v[] = {
                {0, 0, 2000, "one wave, ids from 0"}, {10000000u, 0, 2000, "one wave, ids from 1e7"},
                {0, 200, 2000, "one wave, blocks start 0-200 us apart"}, {0, 2000, 2000, "one wave, blocks start 0-2 ms apart"}};
            for (int vi = 0; vi < 4; vi++) {
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                k_chain_var<<<1184, 256>>>(a, (uint32_t)(n - 1), v[vi].steps, 0, v[vi].toff, v[vi].delay, out2);
                cudaEventRecord(e0);
                k_chain_var<<<1184, 256>>>(a, (uint32_t)(n - 1), v[vi].steps, 1, v[vi].toff, v[vi].delay, out2);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                printf("%-40s %4d loads per chain: %7.3f ms %6.1f G loads/s (delays included)\n", v[vi].what, v[vi].steps, ms, 1184.0 * 256 * v[vi].steps / ms / 1e6);
            }
        }
        {
            const uint32_t nd[] = {0, 1184, 2368, 4736, 18944, 148000};
            for (int vi = 0; vi < 6; vi++) {
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                k_chain_dummy<<<1184 + nd[vi], 256>>>(a, (uint32_t)(n - 1), 2000, 0, nd[vi], out2);
                cudaEventRecord(e0);
                k_chain_dummy<<<1184 + nd[vi], 256>>>(a, (uint32_t)(n - 1), 2000, 1, nd[vi], out2);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                printf("one working wave after %6u blocks that return at once: %7.3f ms %6.1f G loads/s\n", nd[vi], ms, 1184.0 * 256 * 2000 / ms / 1e6);
            }
        }
        {
            const int ps[] = {1, 10, 50, 200};
            for (int vi = 0; vi < 4; vi++)
                for (uint32_t np = 2368; np <= 4736; np += 2368) {
                    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                    k_chain_prime<<<1184 + np, 256>>>(a, (uint32_t)(n - 1), 2000, ps[vi], 0, np, out2);
                    cudaEventRecord(e0);
                    k_chain_prime<<<1184 + np, 256>>>(a, (uint32_t)(n - 1), 2000, ps[vi], 1, np, out2);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1);
                    const double loads = 1184.0 * 256 * 2000 + (double)np * 256 * ps[vi];
                    printf("one working wave after %u blocks of %3d loads per chain: %7.3f ms %6.1f G loads/s overall\n", np, ps[vi], ms, loads / ms / 1e6);
                }
        }
        {
            const uint32_t nt[] = {0, 148, 1184, 2368, 4736, 18944};
            for (int vi = 0; vi < 6; vi++) {
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                k_chain_tail<<<1184 + nt[vi], 256>>>(a, (uint32_t)(n - 1), 2000, 0, 1184, out2);
                cudaEventRecord(e0);
                k_chain_tail<<<1184 + nt[vi], 256>>>(a, (uint32_t)(n - 1), 2000, 1, 1184, out2);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                printf("one working wave with %6u blocks that return at once PENDING behind it: %7.3f ms %6.1f G loads/s\n", nt[vi], ms, 1184.0 * 256 * 2000 / ms / 1e6);
            }
        }
        const int tgrids[] = {1184, 2368, 4736, 18944};
        for (int tg = 0; tg < 4; tg++) {   // per-block lifetimes: loads/s by millisecond (a block's loads spread evenly over its lifetime)
            const int grid = tgrids[tg], steps = 400;
            unsigned long long *t0, *t1, *h0 = new unsigned long long[grid], *h1 = new unsigned long long[grid];
            uint32_t *dsm, *hsm = new uint32_t[grid];
            cudaMalloc(&t0, grid * 8); cudaMalloc(&t1, grid * 8); cudaMalloc(&dsm, grid * 4);
            k_chain_times<<<grid, 256>>>(a, (uint32_t)(n - 1), steps, 0, out2, t0, t1, dsm);
            k_chain_times<<<grid, 256>>>(a, (uint32_t)(n - 1), steps, 1, out2, t0, t1, dsm);
            cudaDeviceSynchronize();
            cudaMemcpy(hsm, dsm, grid * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(h0, t0, grid * 8, cudaMemcpyDeviceToHost); cudaMemcpy(h1, t1, grid * 8, cudaMemcpyDeviceToHost);
            unsigned long long lo = ~0ull, hi = 0;
            for (int b = 0; b < grid; b++) { if (h0[b] < lo) lo = h0[b]; if (h1[b] > hi) hi = h1[b]; }
            const int nb = (int)((hi - lo) / 1000000) + 1;
            double *bins = new double[nb]();
            double life_first = 0, life_mid = 0, life_last = 0;
            for (int b = 0; b < grid; b++) {
                const double s0 = (double)(h0[b] - lo) / 1e6, s1 = (double)(h1[b] - lo) / 1e6, per = 256.0 * steps / (s1 - s0);
                for (int k = (int)s0; k <= (int)s1 && k < nb; k++) { const double a0 = s0 > k ? s0 : k, a1 = s1 < k + 1 ? s1 : k + 1; if (a1 > a0) bins[k] += per * (a1 - a0); }
                if (b < 1184) life_first += s1 - s0; else if (b >= grid - 1184) life_last += s1 - s0; else if (b >= 8 * 1184 && b < 9 * 1184) life_mid += s1 - s0;
            }
            printf("%d-block launch, %d loads per chain: total %.2f ms; mean block lifetime: first 1184 blocks %.3f ms, blocks 9472..10655 %.3f ms, last 1184 blocks %.3f ms\n",
                   grid, steps, (double)(hi - lo) / 1e6, life_first / 1184, life_mid / 1184, life_last / 1184);
            printf("G loads/s by millisecond:");
            for (int k = 0; k < nb; k++) printf(" %.1f", bins[k] / 1e6);
            printf("\n");
            if (grid == 1184) {   // which SMs are slow?  mean lifetime of the 8 blocks of every SM, sorted
                double sum[256] = {0}; int cnt[256] = {0}; double v[256]; int nv = 0, maxid = 0;
                for (int b = 0; b < grid; b++) { const int id = hsm[b] & 255; sum[id] += (double)(h1[b] - h0[b]) / 1e6; cnt[id]++; if (id > maxid) maxid = id; }
                for (int id = 0; id <= maxid; id++) if (cnt[id]) v[nv++] = sum[id] / cnt[id];
                for (int i = 0; i < nv; i++) for (int j = i + 1; j < nv; j++) if (v[j] < v[i]) { double x = v[i]; v[i] = v[j]; v[j] = x; }
                printf("one wave: %d SMs hold blocks (%d..%d blocks each); mean block lifetime per SM, sorted, every 10th: ", nv, 8, 8);
                for (int i = 0; i < nv; i += 10) printf("%.2f ", v[i]);
                printf("... max %.2f ms\n", v[nv - 1]);
                printf("per SM id (mean lifetime ms):");
                for (int id = 0; id <= maxid; id++) if (cnt[id]) printf(" %d:%.2f", id, sum[id] / cnt[id]);
                printf("\n");
            }
        }
        // how many blocks does it take?  (1184 = one resident wave of 8 per SM)
        const int grids[] = {1184, 1332, 2368, 4736, 9472, 18944, 1184};
        for (int gi = 0; gi < 7; gi++) {
            const int steps = 400, grid = grids[gi];
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            k_chain_own<<<grid, 256>>>(a, (uint32_t)(n - 1), steps, 0, out2);
            cudaEventRecord(e0);
            k_chain_own<<<grid, 256>>>(a, (uint32_t)(n - 1), steps, 1, out2);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("%5d blocks of 256, %4d loads per chain: %7.3f ms %6.1f G loads/s\n", grid, steps, ms, (double)grid * 256 * steps / ms / 1e6);
        }
        printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        return 0;
    }
    if (argc > 1 && atoi(argv[1]) == -3) {   // mode 5: occupancy sweep (resident threads per SM), one wave, long chains
        size_t n = (size_t)1 << 27; int4 *a; uint32_t *out; int steps = 4000;
        cudaMalloc(&a, n * sizeof(int4)); cudaMalloc(&out, 148 * 2048 * 4);
        k_fill<<<(unsigned)((n + 255) / 256), 256>>>(a, n);
        for (int tpsm = 64; tpsm <= 2048; tpsm *= 2) {
            int bs = tpsm < 256 ? tpsm : 256, grid = 148 * (tpsm / bs);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            k_chain<1><<<grid, bs>>>(a, (uint32_t)(n - 1), 100, out);
            cudaEventRecord(e0);
            k_chain<1><<<grid, bs>>>(a, (uint32_t)(n - 1), steps, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("%4d threads/SM: %8.3f ms  %6.1f G loads/s  latency/load %.0f ns\n", tpsm, ms, (double)grid * bs * steps / ms / 1e6, ms * 1e6 / steps);
        }
        printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        return 0;
    }
    if (argc > 1 && atoi(argv[1]) < 0) {   // mode 3: paired loads at growing distance
        size_t n = (size_t)1 << 27; int4 *a; uint32_t *out; int nthreads = 1 << 22, steps = 80;
        cudaMalloc(&a, n * sizeof(int4)); cudaMalloc(&out, nthreads * 4);
        k_fill<<<(unsigned)((n + 255) / 256), 256>>>(a, n);
        for (uint32_t dist = 16; dist <= (1u << 22); dist <<= 1) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            k_pair<<<nthreads / 256, 256>>>(a, (uint32_t)(n - 1), steps, dist / 16, out);
            cudaEventRecord(e0);
            k_pair<<<nthreads / 256, 256>>>(a, (uint32_t)(n - 1), steps, dist / 16, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("pair distance %8u B: %7.3f ms  %6.1f G pairs/s  %6.1f G loads/s\n", dist, ms, (double)nthreads * steps / ms / 1e6, 2.0 * nthreads * steps / ms / 1e6);
        }
        printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        return 0;
    }
    int logn = argc > 1 ? atoi(argv[1]) : 27;            // 2^27 x 16 B = 2 GiB
    size_t n = (size_t)1 << logn;
    int4 *a; uint32_t *out;
    int nthreads = 1 << 22, steps = 80;
    cudaMalloc(&a, n * sizeof(int4)); cudaMalloc(&out, nthreads * 4);
    k_fill<<<(unsigned)((n + 255) / 256), 256>>>(a, n);
    const char *names[] = {"ld.global", "ld.global.nc", "ld.global.cg", "ld.global.cs", "nc+L2 evict_first hint", "nc+L1::no_allocate", "nc+L1::no_allocate+L2 hint", "nc.L2::64B", "nc.L2::128B", "nc.L2::256B", "L1::no_allocate.L2::64B", "ld.volatile"};
    for (int gran = 0; gran < 3; gran++) {
        size_t g = gran == 0 ? 0 : (gran == 1 ? 32 : 128);
        if (g) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g); printf("set L2 fetch granularity %zu: %s\n", g, cudaGetErrorString(e)); }
        size_t cur = 0; cudaDeviceGetLimit(&cur, cudaLimitMaxL2FetchGranularity);
        printf("-- L2 fetch granularity limit = %zu B, array %.1f GiB, %d threads x %d dependent loads\n", cur, n * 16.0 / (1 << 30), nthreads, steps);
        float ms[12];
        ms[0] = run<0>(a, (uint32_t)(n - 1), nthreads, steps, out); ms[1] = run<1>(a, (uint32_t)(n - 1), nthreads, steps, out);
        ms[2] = run<2>(a, (uint32_t)(n - 1), nthreads, steps, out); ms[3] = run<3>(a, (uint32_t)(n - 1), nthreads, steps, out);
        ms[4] = run<4>(a, (uint32_t)(n - 1), nthreads, steps, out); ms[5] = run<5>(a, (uint32_t)(n - 1), nthreads, steps, out);
        ms[6] = run<6>(a, (uint32_t)(n - 1), nthreads, steps, out);
        ms[7] = run<7>(a, (uint32_t)(n - 1), nthreads, steps, out); ms[8] = run<8>(a, (uint32_t)(n - 1), nthreads, steps, out);
        ms[9] = run<9>(a, (uint32_t)(n - 1), nthreads, steps, out); ms[10] = run<10>(a, (uint32_t)(n - 1), nthreads, steps, out);
        ms[11] = run<11>(a, (uint32_t)(n - 1), nthreads, steps, out);
        for (int m = 0; m < 12; m++)
            printf("   %-30s %7.3f ms  %6.1f G loads/s  (%.2f TB/s of 32 B sectors)\n", names[m], ms[m], (double)nthreads * steps / ms[m] / 1e6, (double)nthreads * steps * 32 / ms[m] / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
