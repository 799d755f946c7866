// TLB probe: does the way a large array is allocated change the random-access ceiling?
// gather_bench showed that a 64 MiB (L2-resident) footprint spread over > 256 MiB of address space
// runs at the same ~48 G loads/s as a 2 GiB random gather: the ceiling is address translation, not
// DRAM.  This probe times the same dependent random 16-byte loads over 2 GiB allocated by
//   (a) cudaMalloc,
//   (b) cuMemCreate in ONE physical handle + a virtual range aligned to 512 MiB / 1 GiB,
//   (c) cuMemCreate with the RECOMMENDED granularity in 512 MiB handles,
// and prints the allocation granularities the driver reports.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/tlb_probe tools/tlb_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

__global__ void k_chain(const int4 *__restrict__ a, uint32_t mask, int steps, uint32_t *out) {
    uint32_t i = mix(blockIdx.x * blockDim.x + threadIdx.x) & mask;
    uint32_t acc = 0;
    for (int s = 0; s < steps; s++) {
        int4 v = __ldg(a + i);
        acc += v.y;
        i = mix((uint32_t)v.x + s) & mask;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_fill(int4 *a, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) a[i] = make_int4((int)mix((uint32_t)i), (int)i, 0, 0);
}

static void bench(const char *name, int4 *a, size_t n, uint32_t *out) {
    int nthreads = 1 << 22, steps = 80;
    k_fill<<<(unsigned)((n + 255) / 256), 256>>>(a, n);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_chain<<<nthreads / 256, 256>>>(a, (uint32_t)(n - 1), steps, out);
    cudaEventRecord(e0);
    k_chain<<<nthreads / 256, 256>>>(a, (uint32_t)(n - 1), steps, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-58s ptr %p  %7.3f ms  %6.1f G loads/s  (%s)\n", name, (void *)a, ms, (double)nthreads * steps / ms / 1e6,
           cudaGetErrorString(cudaGetLastError()));
}

#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *s_; cuGetErrorString(r_, &s_); printf("%s failed: %s\n", #x, s_); return 1; } } while (0)

static int vmm_alloc(size_t bytes, size_t chunk, size_t va_align, int4 **out_ptr, const char *tag, unsigned char compression) {
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = 0;
    prop.allocFlags.compressionType = compression;
    CUdeviceptr va = 0;
    CK(cuMemAddressReserve(&va, bytes, va_align, 0, 0));
    for (size_t off = 0; off < bytes; off += chunk) {
        CUmemGenericAllocationHandle h;
        CK(cuMemCreate(&h, chunk, &prop, 0));
        CK(cuMemMap(va + off, chunk, 0, h, 0));
        CK(cuMemRelease(h));
    }
    CUmemAccessDesc acc = {};
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CK(cuMemSetAccess(va, bytes, &acc, 1));
    *out_ptr = (int4 *)va;
    (void)tag;
    return 0;
}

int main(int argc, char **argv) {
    int logn = argc > 1 ? atoi(argv[1]) : 27;
    size_t n = (size_t)1 << logn, bytes = n * sizeof(int4);
    cudaSetDevice(0);
    cudaFree(0);
    uint32_t *out; cudaMalloc(&out, (1 << 22) * 4);
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = 0;
    size_t gmin = 0, grec = 0;
    cuMemGetAllocationGranularity(&gmin, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
    cuMemGetAllocationGranularity(&grec, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
    printf("allocation granularity: minimum %zu B, recommended %zu B; array %.1f GiB\n", gmin, grec, bytes / 1073741824.0);

    int4 *a = nullptr;
    cudaMalloc(&a, bytes);
    bench("cudaMalloc", a, n, out);
    cudaFree(a);

    cudaMallocAsync(&a, bytes, 0);
    cudaStreamSynchronize(0);
    bench("cudaMallocAsync (default pool)", a, n, out);
    cudaFreeAsync(a, 0);
    cudaStreamSynchronize(0);

    const size_t M512 = (size_t)512 << 20, G1 = (size_t)1 << 30;
    if (!vmm_alloc(bytes, bytes, G1, &a, "one", 0)) bench("cuMemCreate one handle, VA aligned 1 GiB", a, n, out);
    if (!vmm_alloc(bytes, M512, M512, &a, "512", 0)) bench("cuMemCreate 512 MiB handles, VA aligned 512 MiB", a, n, out);
    if (!vmm_alloc(bytes, (size_t)2 << 20, (size_t)2 << 20, &a, "2m", 0)) bench("cuMemCreate 2 MiB handles, VA aligned 2 MiB", a, n, out);
    if (!vmm_alloc(bytes, grec ? ((M512 + grec - 1) / grec) * grec : M512, M512, &a, "rec", 0)) bench("cuMemCreate recommended-granularity handles", a, n, out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
