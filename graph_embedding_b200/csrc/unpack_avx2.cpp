// unpack_avx2.cpp — 3-byte -> 4-byte vertex ids with AVX2 (built with -mavx2; chosen at run time by unpack24()).
// Eight ids per iteration: two 16-byte loads 12 bytes apart, one byte shuffle per 128-bit lane, one 32-byte store
// (non-temporal when the destination is 32-byte aligned: the corpus is written once and read by someone else).
#include "hostpipe.h"

#if defined(__x86_64__)
#include <immintrin.h>

namespace gw {

void unpack24_avx2(const uint8_t *src, int32_t *dst, size_t count) {
    const __m256i shuf = _mm256_setr_epi8(0, 1, 2, -1, 3, 4, 5, -1, 6, 7, 8, -1, 9, 10, 11, -1,
                                          0, 1, 2, -1, 3, 4, 5, -1, 6, 7, 8, -1, 9, 10, 11, -1);
    size_t i = 0;
    // head: scalar until dst is 32-byte aligned (rows of 80 ids keep every part aligned when the base is)
    while (i < count && ((uintptr_t)(dst + i) & 31)) {
        const uint8_t *p = src + 3 * i;
        dst[i] = (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16));
        i++;
    }
    // the second load of an iteration reads bytes [12, 28) of its 24: stay 8 ids short of the end
    const size_t vec_end = count >= 8 ? count - 8 : 0;
    for (; i + 32 <= vec_end; i += 32) {
#pragma GCC unroll 4
        for (int j = 0; j < 4; j++) {
            const uint8_t *p = src + 3 * (i + 8 * j);
            const __m128i lo = _mm_loadu_si128((const __m128i *)p);
            const __m128i hi = _mm_loadu_si128((const __m128i *)(p + 12));
            const __m256i v = _mm256_shuffle_epi8(_mm256_set_m128i(hi, lo), shuf);
            _mm256_stream_si256((__m256i *)(dst + i + 8 * j), v);
        }
    }
    for (; i < vec_end; i += 8) {
        const uint8_t *p = src + 3 * i;
        const __m128i lo = _mm_loadu_si128((const __m128i *)p);
        const __m128i hi = _mm_loadu_si128((const __m128i *)(p + 12));
        const __m256i v = _mm256_shuffle_epi8(_mm256_set_m128i(hi, lo), shuf);
        _mm256_stream_si256((__m256i *)(dst + i), v);
    }
    _mm_sfence();
    unpack24_scalar(src + 3 * i, dst + i, count - i);
}

void copy_stream_avx2(const void *src, void *dst, size_t bytes) {
    const uint8_t *s = (const uint8_t *)src;
    uint8_t *d = (uint8_t *)dst;
    size_t i = 0;
    while (i < bytes && ((uintptr_t)(d + i) & 31)) { d[i] = s[i]; i++; }
    for (; i + 128 <= bytes; i += 128) {
        const __m256i a = _mm256_loadu_si256((const __m256i *)(s + i)), b = _mm256_loadu_si256((const __m256i *)(s + i + 32));
        const __m256i c = _mm256_loadu_si256((const __m256i *)(s + i + 64)), e = _mm256_loadu_si256((const __m256i *)(s + i + 96));
        _mm256_stream_si256((__m256i *)(d + i), a); _mm256_stream_si256((__m256i *)(d + i + 32), b);
        _mm256_stream_si256((__m256i *)(d + i + 64), c); _mm256_stream_si256((__m256i *)(d + i + 96), e);
    }
    _mm_sfence();
    for (; i < bytes; i++) d[i] = s[i];
}

}  // namespace gw
#endif
