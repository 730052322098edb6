// Shared helpers for libzafb200: error plumbing, complex arithmetic, compile-time twiddles.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/zafb200.h"

namespace zafb {

// ---------------------------------------------------------------- errors
std::string& last_error_ref();
int fail(int code, const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
extern std::atomic<int64_t> g_h2d_bytes, g_d2h_bytes;  // bytes the host pipelines have moved over the link

#define ZAFB_CUDA(expr)                                                                   \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess)                                                            \
            return ::zafb::fail(_e == cudaErrorMemoryAllocation ? ZAFB_E_NOMEM : ZAFB_E_CUDA, \
                                "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                                __FILE__, __LINE__);                                      \
    } while (0)

#define ZAFB_LAUNCH_CHECK()                                                               \
    do {                                                                                  \
        ::zafb::g_launches.fetch_add(1, std::memory_order_relaxed);                       \
        ZAFB_CUDA(cudaGetLastError());                                                    \
    } while (0)

#define ZAFB_REQUIRE(cond, ...)                                                           \
    do {                                                                                  \
        if (!(cond)) return ::zafb::fail(ZAFB_E_BADARG, __VA_ARGS__);                     \
    } while (0)

inline bool is_pow2(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }
inline int ilog2(int64_t v) {
    int l = 0;
    while ((int64_t(1) << (l + 1)) <= v) ++l;
    return l;
}
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Upload a host vector of doubles as fp32 / a table of complex doubles as float2.
int upload_f32(float** dev, const double* host, size_t n);
int upload_c32(float2** dev, const double* host_ri, size_t n);
// W_n^t = exp(-2 pi i t / n), t in [0, count), generated in float64.
int upload_twiddles(float2** dev, int64_t n, int64_t count);
int sm_count();
// tuning switch read from the environment (experiments only; the default is the shipped path)
int env_flag(const char* name, int dflt);

// ---------------------------------------------------------------- complex helpers
#define ZAFB_HD __host__ __device__ __forceinline__

ZAFB_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
ZAFB_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
ZAFB_HD float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
ZAFB_HD float2 cmul_conj(float2 a, float2 b) {  // a * conj(b)
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
ZAFB_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
ZAFB_HD float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }

#ifdef __CUDACC__
// pull one 128-byte line towards L2 (no register, no scoreboard entry): used to fetch a warp's NEXT frame while it
// transforms the current one
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif

// float64 twins (the optional float64 route of melspectrogram / mfcc)
ZAFB_HD double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
ZAFB_HD double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
ZAFB_HD double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// ---------------------------------------------------------------- compile-time trigonometry
// constexpr sin/cos in double (argument reduced to [0, pi/4], Taylor to < 1e-17) so that every
// in-register twiddle is an immediate operand rounded once from float64.
namespace ct {
constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double sin_small(double x) {  // |x| <= pi/4
    double x2 = x * x, term = x, sum = x;
    for (int i = 1; i < 14; ++i) {
        term *= -x2 / double((2 * i) * (2 * i + 1));
        sum += term;
    }
    return sum;
}
constexpr double cos_small(double x) {
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int i = 1; i < 14; ++i) {
        term *= -x2 / double((2 * i - 1) * (2 * i));
        sum += term;
    }
    return sum;
}
// cos(2 pi num/den), sin(2 pi num/den) with exact octant reduction on the integer fraction.
constexpr double cos2pi(long long num, long long den) {
    num %= den;
    if (num < 0) num += den;
    // reduce to first octant using symmetries on the rational angle num/den in [0,1)
    if (2 * num > den) return cos2pi(den - num, den);           // cos(2pi - a) = cos a
    if (4 * num > den) return -cos2pi(den - 2 * num, 2 * den);  // a in (pi/2, pi]: -cos(pi - a)
    if (8 * num > den) {                                        // a in (pi/4, pi/2]: sin(pi/2 - a)
        return sin_small(2.0 * kPi * double(den - 4 * num) / double(4 * den));
    }
    return cos_small(2.0 * kPi * double(num) / double(den));
}
constexpr double sin2pi(long long num, long long den) {
    num %= den;
    if (num < 0) num += den;
    if (2 * num > den) return -sin2pi(den - num, den);          // sin(2pi - a) = -sin a
    if (4 * num > den) return sin2pi(den - 2 * num, 2 * den);   // sin(pi - a)
    if (8 * num > den) {                                        // cos(pi/2 - a)
        return cos_small(2.0 * kPi * double(den - 4 * num) / double(4 * den));
    }
    return sin_small(2.0 * kPi * double(num) / double(den));
}
}  // namespace ct

// W_DEN^NUM = exp(-2 pi i NUM / DEN) as fp32 immediates.
template <int NUM, int DEN>
struct Tw {
    static constexpr float re = float(ct::cos2pi(NUM, DEN));
    static constexpr float im = float(-ct::sin2pi(NUM, DEN));
};

// a * W_DEN^NUM with the trivial rotations folded away at compile time.
template <int NUM_, int DEN>
ZAFB_HD float2 mul_tw(float2 a) {
    constexpr int NUM = ((NUM_ % DEN) + DEN) % DEN;
    if constexpr (NUM == 0) {
        return a;
    } else if constexpr (4 * NUM == DEN) {  // -i
        return make_float2(a.y, -a.x);
    } else if constexpr (2 * NUM == DEN) {  // -1
        return make_float2(-a.x, -a.y);
    } else if constexpr (4 * NUM == 3 * DEN) {  // +i
        return make_float2(-a.y, a.x);
    } else if constexpr (8 * NUM == DEN) {  // (1 - i)/sqrt2
        constexpr float c = Tw<1, 8>::re;
        return make_float2(c * (a.x + a.y), c * (a.y - a.x));
    } else if constexpr (8 * NUM == 3 * DEN) {  // (-1 - i)/sqrt2
        constexpr float c = Tw<1, 8>::re;
        return make_float2(c * (a.y - a.x), -c * (a.x + a.y));
    } else if constexpr (8 * NUM == 5 * DEN) {  // (-1 + i)/sqrt2
        constexpr float c = Tw<1, 8>::re;
        return make_float2(-c * (a.x + a.y), c * (a.x - a.y));
    } else if constexpr (8 * NUM == 7 * DEN) {  // (1 + i)/sqrt2
        constexpr float c = Tw<1, 8>::re;
        return make_float2(c * (a.x - a.y), c * (a.x + a.y));
    } else {
        constexpr float wr = Tw<NUM, DEN>::re, wi = Tw<NUM, DEN>::im;
        return make_float2(a.x * wr - a.y * wi, a.x * wi + a.y * wr);
    }
}

// the same rotation for a float64 value (constants in full double precision)
template <int NUM_, int DEN>
ZAFB_HD double2 mul_tw(double2 a) {
    constexpr int NUM = ((NUM_ % DEN) + DEN) % DEN;
    if constexpr (NUM == 0) {
        return a;
    } else if constexpr (4 * NUM == DEN) {
        return make_double2(a.y, -a.x);
    } else if constexpr (2 * NUM == DEN) {
        return make_double2(-a.x, -a.y);
    } else if constexpr (4 * NUM == 3 * DEN) {
        return make_double2(-a.y, a.x);
    } else {
        constexpr double wr = ct::cos2pi(NUM, DEN), wi = -ct::sin2pi(NUM, DEN);
        return make_double2(a.x * wr - a.y * wi, a.x * wi + a.y * wr);
    }
}

__host__ __device__ constexpr int bitrev(int v, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1) << (bits - 1 - i);
    return r;
}
__host__ __device__ constexpr int clog2(int v) { return v <= 1 ? 0 : 1 + clog2(v / 2); }

}  // namespace zafb
