// hostpipe.cpp — copy-thread pool and the unpack / copy routines of the corpus hand-off (see hostpipe.h).
#include "hostpipe.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace gw {

struct CopyPool::Impl {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<uint64_t> generation{0};      // bumped once per job
    std::atomic<int> pending{0};              // workers still inside the current job
    std::atomic<bool> stop{false};
    job_fn fn = nullptr;
    void *arg = nullptr;
    int nparts = 1;

    void worker(int part) {
        uint64_t seen = 0;
        for (;;) {
            // spin for the next job (chunks arrive every few hundred microseconds), then sleep
            const auto t0 = std::chrono::steady_clock::now();
            bool got = false;
            for (int it = 0;; it++) {
                if (stop.load(std::memory_order_acquire)) return;
                if (generation.load(std::memory_order_acquire) != seen) { got = true; break; }
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
                if ((it & 255) == 255) {
                    if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(400)) break;
                    std::this_thread::yield();
                }
            }
            if (!got) {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop.load() || generation.load() != seen; });
                if (stop.load()) return;
            }
            seen = generation.load(std::memory_order_acquire);
            fn(part, nparts, arg);
            pending.fetch_sub(1, std::memory_order_acq_rel);
        }
    }
};

CopyPool::CopyPool(int threads) : impl_(new Impl), nthreads_(threads < 1 ? 1 : threads) {
    impl_->nparts = nthreads_;
    for (int i = 1; i < nthreads_; i++) impl_->workers.emplace_back([this, i] { impl_->worker(i); });
}

CopyPool::~CopyPool() {
    {
        std::lock_guard<std::mutex> lk(impl_->mu);
        impl_->stop.store(true);
    }
    impl_->cv.notify_all();
    for (auto &t : impl_->workers) t.join();
    delete impl_;
}

void CopyPool::run(job_fn fn, void *arg) {
    Impl &m = *impl_;
    m.fn = fn;
    m.arg = arg;
    m.pending.store(nthreads_ - 1, std::memory_order_release);
    {
        std::lock_guard<std::mutex> lk(m.mu);                 // pairs with the sleepers' predicate check
        m.generation.fetch_add(1, std::memory_order_acq_rel);
    }
    m.cv.notify_all();
    fn(0, nthreads_, arg);
    while (m.pending.load(std::memory_order_acquire) != 0) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
}

int default_copy_threads() {
    const char *e = getenv("GW_HOST_THREADS");
    if (e && atoi(e) > 0) return atoi(e) > 64 ? 64 : atoi(e);
    int hw = (int)std::thread::hardware_concurrency();
    if (hw < 1) hw = 1;
    const char *lw = getenv("LOCAL_WORLD_SIZE");
    const int ranks = (lw && atoi(lw) > 0) ? atoi(lw) : 1;
    int t = hw / ranks;
    if (t < 1) t = 1;
    if (t > 12) t = 12;                 // profiles/r2_handoff_sweep.txt: the ring saturates the host memory system at 12
    return t;
}

void unpack24_scalar(const uint8_t *src, int32_t *dst, size_t count) {
    for (size_t i = 0; i < count; i++) {
        const uint8_t *p = src + 3 * i;
        dst[i] = (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16));
    }
}

void unpack24(const uint8_t *src, int32_t *dst, size_t count) {
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) { unpack24_avx2(src, dst, count); return; }
#endif
    unpack24_scalar(src, dst, count);
}

void copy_stream(const void *src, void *dst, size_t bytes) {
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) { copy_stream_avx2(src, dst, bytes); return; }
#endif
    memcpy(dst, src, bytes);
}

namespace {
struct DrainJob {
    const uint8_t *src;
    int packed;
    const int32_t *lens;
    int64_t n_walks;
    int32_t L;
    int32_t *dst;
    int32_t *dst_lens;
};

void drain_part(int part, int nparts, void *arg) {
    const DrainJob &J = *(const DrainJob *)arg;
    // whole rows per part, cut on multiples of 8 rows so that every part starts 32-byte aligned when the base is
    const int64_t per = ((J.n_walks + nparts - 1) / nparts + 7) & ~(int64_t)7;
    const int64_t w0 = (int64_t)part * per, w1 = w0 + per < J.n_walks ? w0 + per : J.n_walks;
    if (w0 >= w1) return;
    const size_t first = (size_t)w0 * (size_t)J.L, count = (size_t)(w1 - w0) * (size_t)J.L;
    if (J.packed) unpack24(J.src + 3 * first, J.dst + first, count);
    else copy_stream(J.src + 4 * first, J.dst + first, 4 * count);
    if (J.lens) {
        if (J.packed)
            for (int64_t w = w0; w < w1; w++) {
                const int32_t len = J.lens[w];
                if (len < J.L) {
                    int32_t *row = J.dst + (size_t)w * (size_t)J.L;
                    for (int32_t i = len < 0 ? 0 : len; i < J.L; i++) row[i] = -1;
                }
            }
        if (J.dst_lens) memcpy(J.dst_lens + w0, J.lens + w0, sizeof(int32_t) * (size_t)(w1 - w0));
    }
}
}  // namespace

void drain_chunk(CopyPool *pool, const void *src, int packed, const int32_t *lens, int64_t n_walks, int32_t L, int32_t *dst,
                 int32_t *dst_lens) {
    DrainJob J{(const uint8_t *)src, packed, lens, n_walks, L, dst, dst_lens};
    if (pool && pool->threads() > 1) pool->run(drain_part, &J);
    else drain_part(0, 1, &J);
}

}  // namespace gw

// ---- C ABI: the host half of the packed hand-off on its own (a corpus kept packed, e.g. in a file or a message) ----
extern "C" int gw_corpus_unpack24(const void *packed, const int32_t *lens, int64_t n_walks, int32_t walk_length, int32_t threads,
                                  int32_t *out_walks) {
    if (n_walks < 0 || walk_length < 1 || (n_walks > 0 && (!packed || !out_walks))) return -1;      // GW_E_INVALID
    if (n_walks == 0) return 0;
    const int t = threads > 0 ? (threads > 64 ? 64 : threads) : gw::default_copy_threads();
    // the vector routine may read up to 4 bytes past the last id: widen the tail from a padded copy
    const int64_t tail = n_walks < 8 ? n_walks : 8, head = n_walks - tail;
    if (head > 0) {
        gw::CopyPool pool(t);
        gw::drain_chunk(&pool, packed, 1, lens, head, walk_length, out_walks, nullptr);
    }
    const size_t tb = 3 * (size_t)tail * (size_t)walk_length;
    std::vector<uint8_t> pad(tb + 16, 0);
    memcpy(pad.data(), (const uint8_t *)packed + 3 * (size_t)head * (size_t)walk_length, tb);
    gw::drain_chunk(nullptr, pad.data(), 1, lens ? lens + head : nullptr, tail, walk_length, out_walks + (size_t)head * (size_t)walk_length, nullptr);
    return 0;
}

// ---- random.shuffle(nodes) of simulate_walks (node2vec.py:51) at native speed, on the interpreter's own generator ----
// CPython's random.shuffle (3.2+, the reference's pinned 3.5 included): for i = n-1 .. 1: j = _randbelow(i + 1); swap(x[i], x[j]),
// with _randbelow(m) = getrandbits(m.bit_length()) redrawn while >= m, and getrandbits(k <= 32) = genrand_uint32() >> (32 - k)
// of MT19937.  The caller passes random.getstate()'s 624 words + index and puts them back afterwards: the permutation AND
// the state left behind are those of the pure-Python loop, which costs 0.7 us per element (3 s for the 4.2 M nodes of
// R-MAT-22, per pass) against 10 ns here.
extern "C" int gw_py_random_shuffle(uint32_t *mt624, int32_t *mt_index, int64_t *items, int64_t n) {
    if (!mt624 || !mt_index || n < 0 || (n > 0 && !items) || *mt_index < 0 || *mt_index > 624 || n > 0xFFFFFFFFLL) return -1;
    uint32_t *mt = mt624;
    int idx = *mt_index;
    auto next32 = [&]() -> uint32_t {
        if (idx >= 624) {                                      // genrand_uint32's block regeneration (mt19937ar.c)
            const uint32_t UP = 0x80000000u, LO = 0x7fffffffu, A = 0x9908b0dfu;
            int kk = 0;
            for (; kk < 624 - 397; kk++) { const uint32_t y = (mt[kk] & UP) | (mt[kk + 1] & LO); mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? A : 0u); }
            for (; kk < 623; kk++) { const uint32_t y = (mt[kk] & UP) | (mt[kk + 1] & LO); mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? A : 0u); }
            const uint32_t y = (mt[623] & UP) | (mt[0] & LO);
            mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
        return y;
    };
    for (int64_t i = n - 1; i >= 1; i--) {
        const uint32_t m = (uint32_t)(i + 1);
        const int k = 32 - __builtin_clz(m);                   // m.bit_length(), m >= 2
        uint32_t r = next32() >> (32 - k);
        while (r >= m) r = next32() >> (32 - k);
        const int64_t t = items[i]; items[i] = items[r]; items[r] = t;
    }
    *mt_index = idx;
    return 0;
}
