// hostcopy.c — the copy the pool threads run between a pinned staging buffer and pageable caller memory.
// Non-temporal stores: the destination (a fresh malloc() block of the output mesh, or the pinned ring on the way in) is
// far larger than the caches and is not read again soon, so a regular store would first pull every line in
// (read-for-ownership) — a third of the memory traffic of the copy.  Plain C for gcc; no CUDA here.
#include <emmintrin.h>
#include <immintrin.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

__attribute__((target("avx2"))) static void copy_nt_avx2(char *d, const char *s, size_t n) {
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    const __m256i a = _mm256_loadu_si256((const __m256i *)(s + i));
    const __m256i b = _mm256_loadu_si256((const __m256i *)(s + i + 32));
    const __m256i c = _mm256_loadu_si256((const __m256i *)(s + i + 64));
    const __m256i e = _mm256_loadu_si256((const __m256i *)(s + i + 96));
    _mm256_stream_si256((__m256i *)(d + i), a);
    _mm256_stream_si256((__m256i *)(d + i + 32), b);
    _mm256_stream_si256((__m256i *)(d + i + 64), c);
    _mm256_stream_si256((__m256i *)(d + i + 96), e);
  }
  if (i < n) memcpy(d + i, s + i, n - i);
}
static void copy_nt_sse2(char *d, const char *s, size_t n) {
  size_t i = 0;
  for (; i + 64 <= n; i += 64) {
    const __m128i a = _mm_loadu_si128((const __m128i *)(s + i));
    const __m128i b = _mm_loadu_si128((const __m128i *)(s + i + 16));
    const __m128i c = _mm_loadu_si128((const __m128i *)(s + i + 32));
    const __m128i e = _mm_loadu_si128((const __m128i *)(s + i + 48));
    _mm_stream_si128((__m128i *)(d + i), a);
    _mm_stream_si128((__m128i *)(d + i + 16), b);
    _mm_stream_si128((__m128i *)(d + i + 32), c);
    _mm_stream_si128((__m128i *)(d + i + 48), e);
  }
  if (i < n) memcpy(d + i, s + i, n - i);
}

// dst and src must not overlap
void b2m_stream_copy(void *dst, const void *src, size_t n) {
  static int have_avx2 = -1;
  if (have_avx2 < 0) have_avx2 = __builtin_cpu_supports("avx2") ? 1 : 0;
  char *d = (char *)dst;
  const char *s = (const char *)src;
  if (n < 4096) { memcpy(d, s, n); return; }
  const size_t head = (size_t)(-(uintptr_t)d & 31);  // non-temporal stores need an aligned destination
  if (head) { memcpy(d, s, head); d += head; s += head; n -= head; }
  if (have_avx2) copy_nt_avx2(d, s, n);
  else copy_nt_sse2(d, s, n);
  _mm_sfence();
}

// dst[i] = (double)src[i]: the widening half of a device-to-host copy of f32-valued doubles
__attribute__((target("avx2"))) static void widen_avx2(double *d, const float *s, size_t n) {
  size_t i = 0;
  for (; i + 16 <= n; i += 16) {
    const __m128 a = _mm_loadu_ps(s + i), b = _mm_loadu_ps(s + i + 4), c = _mm_loadu_ps(s + i + 8), e = _mm_loadu_ps(s + i + 12);
    _mm256_stream_pd(d + i, _mm256_cvtps_pd(a));
    _mm256_stream_pd(d + i + 4, _mm256_cvtps_pd(b));
    _mm256_stream_pd(d + i + 8, _mm256_cvtps_pd(c));
    _mm256_stream_pd(d + i + 12, _mm256_cvtps_pd(e));
  }
  for (; i < n; i++) d[i] = (double)s[i];
}
static void widen_sse2(double *d, const float *s, size_t n) {
  size_t i = 0;
  for (; i + 4 <= n; i += 4) {
    const __m128 a = _mm_loadu_ps(s + i);
    _mm_stream_pd(d + i, _mm_cvtps_pd(a));
    _mm_stream_pd(d + i + 2, _mm_cvtps_pd(_mm_movehl_ps(a, a)));
  }
  for (; i < n; i++) d[i] = (double)s[i];
}
void b2m_stream_widen(double *dst, const float *src, size_t n) {
  while (n && ((uintptr_t)dst & 31)) { *dst++ = (double)*src++; n--; }  // non-temporal stores need an aligned destination
  if (__builtin_cpu_supports("avx2")) widen_avx2(dst, src, n);
  else widen_sse2(dst, src, n);
  _mm_sfence();
}
