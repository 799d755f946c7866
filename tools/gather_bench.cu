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
    else asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
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
    int logn = argc > 1 ? atoi(argv[1]) : 27;            // 2^27 x 16 B = 2 GiB
    size_t n = (size_t)1 << logn;
    int4 *a; uint32_t *out;
    int nthreads = 1 << 22, steps = 80;
    cudaMalloc(&a, n * sizeof(int4)); cudaMalloc(&out, nthreads * 4);
    k_fill<<<(unsigned)((n + 255) / 256), 256>>>(a, n);
    const char *names[] = {"ld.global", "ld.global.nc", "ld.global.cg", "ld.global.cs", "nc+L2 evict_first hint", "nc+L1::no_allocate", "nc+L1::no_allocate+L2 hint"};
    for (int gran = 0; gran < 3; gran++) {
        size_t g = gran == 0 ? 0 : (gran == 1 ? 32 : 128);
        if (g) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g); printf("set L2 fetch granularity %zu: %s\n", g, cudaGetErrorString(e)); }
        size_t cur = 0; cudaDeviceGetLimit(&cur, cudaLimitMaxL2FetchGranularity);
        printf("-- L2 fetch granularity limit = %zu B, array %.1f GiB, %d threads x %d dependent loads\n", cur, n * 16.0 / (1 << 30), nthreads, steps);
        float ms[7];
        ms[0] = run<0>(a, (uint32_t)(n - 1), nthreads, steps, out); ms[1] = run<1>(a, (uint32_t)(n - 1), nthreads, steps, out);
        ms[2] = run<2>(a, (uint32_t)(n - 1), nthreads, steps, out); ms[3] = run<3>(a, (uint32_t)(n - 1), nthreads, steps, out);
        ms[4] = run<4>(a, (uint32_t)(n - 1), nthreads, steps, out); ms[5] = run<5>(a, (uint32_t)(n - 1), nthreads, steps, out);
        ms[6] = run<6>(a, (uint32_t)(n - 1), nthreads, steps, out);
        for (int m = 0; m < 7; m++)
            printf("   %-30s %7.3f ms  %6.1f G loads/s  (%.2f TB/s of 32 B sectors)\n", names[m], ms[m], (double)nthreads * steps / ms[m] / 1e6, (double)nthreads * steps * 32 / ms[m] / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
