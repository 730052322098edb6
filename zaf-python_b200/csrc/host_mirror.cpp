// Host half of the half-spectrum D2H path (zafb_stft_host_f32, zafb_host_mirror_fill): given frames whose bins 0 .. N/2
// have arrived over PCIe, write bins N/2+1 .. N-1 as the conjugates of bins N/2-1 .. 1 -- the Hermitian mirror of a real
// signal's spectrum, the same bits the device kernel stores.  Pure data movement with a sign flip; it has to keep up with
// the DMA engine (7.7 MB per 140 us at cfg 2), so the loops read ASCENDING addresses (hardware prefetchers follow them),
// write with non-temporal stores (the mirrored half is not read again here), and use 32-byte AVX2 operations where the CPU
// has them (checked at run time; x86-64 baseline SSE2 otherwise, plain loops on other architectures).
#include <cstddef>
#include <cstdint>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define ZAFB_HOST_X86 1
#else
#define ZAFB_HOST_X86 0
#endif

namespace zafb {

struct cfloat {
    float re, im;
};

namespace {

void fill_scalar(cfloat* o, int64_t n) {
    for (int64_t k = 1; k < n / 2; ++k) {
        o[n - k].re = o[k].re;
        o[n - k].im = -o[k].im;
    }
}

#if ZAFB_HOST_X86
// sources k = 1 .. n/2 - 1 ascending, two per step: destinations (n-k-1, n-k) <- (conj o[k+1], conj o[k]); n a multiple of 4,
// so n - k - 1 is even (16-byte aligned relative to the frame) for odd k
void fill_sse2(cfloat* o, int64_t n, bool aligned) {
    const __m128 sign = _mm_castsi128_ps(_mm_set_epi32(int(0x80000000u), 0, int(0x80000000u), 0));
    const int64_t h = n / 2;
    int64_t k = 1;
    for (; k + 1 < h; k += 2) {
        const __m128 v = _mm_loadu_ps(reinterpret_cast<const float*>(o + k));                 // (o[k], o[k+1])
        const __m128 r = _mm_xor_ps(_mm_shuffle_ps(v, v, _MM_SHUFFLE(1, 0, 3, 2)), sign);     // (conj o[k+1], conj o[k])
        if (aligned) _mm_stream_ps(reinterpret_cast<float*>(o + n - k - 1), r);
        else _mm_storeu_ps(reinterpret_cast<float*>(o + n - k - 1), r);
    }
    for (; k < h; ++k) {
        o[n - k].re = o[k].re;
        o[n - k].im = -o[k].im;
    }
}

// four per step: destinations n-k-3 .. n-k <- conj of sources k+3 .. k; for k = 1 mod 4 the destination index is a multiple of 4
__attribute__((target("avx2"))) void fill_avx2(cfloat* o, int64_t n, bool aligned) {
    const __m256 sign = _mm256_castsi256_ps(_mm256_set_epi32(int(0x80000000u), 0, int(0x80000000u), 0, int(0x80000000u), 0,
                                                              int(0x80000000u), 0));
    const int64_t h = n / 2;
    int64_t k = 1;
    for (; k + 3 < h; k += 4) {
        const __m256d v = _mm256_loadu_pd(reinterpret_cast<const double*>(o + k));              // one complex64 per 64-bit lane
        const __m256 r = _mm256_xor_ps(_mm256_castpd_ps(_mm256_permute4x64_pd(v, 0x1B)), sign);  // lanes reversed, conjugated
        if (aligned) _mm256_stream_ps(reinterpret_cast<float*>(o + n - k - 3), r);
        else _mm256_storeu_ps(reinterpret_cast<float*>(o + n - k - 3), r);
    }
    for (; k < h; ++k) {
        o[n - k].re = o[k].re;
        o[n - k].im = -o[k].im;
    }
}

__attribute__((target("avx2"))) void conj_row_avx2(const cfloat* src, cfloat* dst, int64_t nt) {
    const __m256 sign = _mm256_castsi256_ps(_mm256_set_epi32(int(0x80000000u), 0, int(0x80000000u), 0, int(0x80000000u), 0,
                                                              int(0x80000000u), 0));
    int64_t j = 0;
    while (j < nt && (reinterpret_cast<uintptr_t>(dst + j) & 31) != 0) {  // peel to a 32-byte aligned destination
        dst[j].re = src[j].re;
        dst[j].im = -src[j].im;
        ++j;
    }
    for (; j + 4 <= nt; j += 4)
        _mm256_stream_ps(reinterpret_cast<float*>(dst + j), _mm256_xor_ps(_mm256_loadu_ps(reinterpret_cast<const float*>(src + j)), sign));
    for (; j < nt; ++j) {
        dst[j].re = src[j].re;
        dst[j].im = -src[j].im;
    }
}

void conj_row_sse2(const cfloat* src, cfloat* dst, int64_t nt) {
    const __m128 sign = _mm_castsi128_ps(_mm_set_epi32(int(0x80000000u), 0, int(0x80000000u), 0));
    int64_t j = 0;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) != 0 && nt > 0) {
        dst[0].re = src[0].re;
        dst[0].im = -src[0].im;
        j = 1;
    }
    for (; j + 2 <= nt; j += 2)
        _mm_stream_ps(reinterpret_cast<float*>(dst + j), _mm_xor_ps(_mm_loadu_ps(reinterpret_cast<const float*>(src + j)), sign));
    for (; j < nt; ++j) {
        dst[j].re = src[j].re;
        dst[j].im = -src[j].im;
    }
}

bool have_avx2() {
    static const bool yes = __builtin_cpu_supports("avx2");
    return yes;
}
#endif

}  // namespace

// out[N - k] = conj(out[k]), k = 1 .. N/2 - 1, for `frames` consecutive frame-major frames of n bins (n a multiple of 4)
void host_mirror_fill_frames(void* out, int64_t frames, int64_t n) {
    cfloat* o = static_cast<cfloat*>(out);
#if ZAFB_HOST_X86
    if (have_avx2()) {
        const bool aligned = (reinterpret_cast<uintptr_t>(o) & 31) == 0;  // frames are n * 8 bytes apart, n a multiple of 4
        for (int64_t f = 0; f < frames; ++f) fill_avx2(o + f * n, n, aligned);
    } else {
        const bool aligned = (reinterpret_cast<uintptr_t>(o) & 15) == 0;
        for (int64_t f = 0; f < frames; ++f) fill_sse2(o + f * n, n, aligned);
    }
    _mm_sfence();
#else
    for (int64_t f = 0; f < frames; ++f) fill_scalar(o + f * n, n);
#endif
}

// BIN_MAJOR twin: rows [r_lo, r_hi) of the flattened (clip, k) index, k = 1 .. N/2 - 1: row N - k of the clip = conj(row k)
void host_mirror_fill_rows(void* out, int64_t nt, int64_t n, int64_t r_lo, int64_t r_hi) {
    cfloat* o = static_cast<cfloat*>(out);
    const int64_t per_clip = n / 2 - 1;
    for (int64_t r = r_lo; r < r_hi; ++r) {
        const int64_t clip = r / per_clip, k = 1 + (r - clip * per_clip);
        const cfloat* src = o + (clip * n + k) * nt;
        cfloat* dst = o + (clip * n + (n - k)) * nt;
#if ZAFB_HOST_X86
        if (have_avx2()) conj_row_avx2(src, dst, nt);
        else conj_row_sse2(src, dst, nt);
#else
        for (int64_t j = 0; j < nt; ++j) {
            dst[j].re = src[j].re;
            dst[j].im = -src[j].im;
        }
#endif
    }
#if ZAFB_HOST_X86
    _mm_sfence();
#endif
}

}  // namespace zafb
