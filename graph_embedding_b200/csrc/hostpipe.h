// hostpipe.h — host side of the corpus hand-off: a persistent pool of copy threads and the
// routines that move a corpus chunk from a pinned staging slot into the caller's buffer, which may be pageable
// (a numpy array, a JVM heap array) and may need its vertex ids unpacked from 3 to 4 bytes.
//
// Plain C++ (no CUDA): compiled by g++, linked into libgraphwalk.so.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace gw {

// A fixed set of worker threads that runs fn(part, nparts, arg) for part = 0..nparts-1 and returns when all
// parts are done; the calling thread works too.  Workers spin briefly between jobs (a corpus is drained chunk by
// chunk, one job every few hundred microseconds) and sleep on a condition variable when the pipeline is idle.
class CopyPool {
   public:
    typedef void (*job_fn)(int part, int nparts, void *arg);
    explicit CopyPool(int threads);
    ~CopyPool();
    int threads() const { return nthreads_; }
    void run(job_fn fn, void *arg);

   private:
    struct Impl;
    Impl *impl_;
    int nthreads_;
};

// Number of copy threads for this process: GW_HOST_THREADS, else hardware threads / LOCAL_WORLD_SIZE (torchrun
// exports it: one rank per GPU shares the box), clamped to [1, 12].
int default_copy_threads();

// dst[i] = little-endian 24-bit src[3i .. 3i+2], zero-extended, for i in [0, count).  src must be readable up to
// 3*count + 4 bytes (staging slots carry slack).  Picks AVX2 at run time when the CPU has it.
void unpack24(const uint8_t *src, int32_t *dst, size_t count);
// the portable loop (also the tail of the vector routine)
void unpack24_scalar(const uint8_t *src, int32_t *dst, size_t count);
#if defined(__x86_64__)
void unpack24_avx2(const uint8_t *src, int32_t *dst, size_t count);     // unpack_avx2.cpp, built with -mavx2
void copy_stream_avx2(const void *src, void *dst, size_t bytes);        // memcpy with non-temporal stores (no read-for-ownership)
#endif
// plain chunk copy of the ring (ids stay 4 bytes): non-temporal stores when the CPU has AVX2, memcpy otherwise
void copy_stream(const void *src, void *dst, size_t bytes);

// One chunk of walks, staged in pinned memory, into rows [0, n_walks) of dst (row length L ids).
//   packed != 0: src holds 3-byte ids; else 4-byte ids (plain copy).
//   lens (may be NULL): per-walk lengths; positions >= lens[w] are written as -1 (a packed pad carries no sign).
// Splits the rows over the pool's threads.
void drain_chunk(CopyPool *pool, const void *src, int packed, const int32_t *lens, int64_t n_walks, int32_t L, int32_t *dst,
                 int32_t *dst_lens);

}  // namespace gw
