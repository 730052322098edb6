// Definitions shared by mdct.cu (plans, frame-major kernels, entry points) and mdct_binmajor.cu (the kernels that read and
// write the reference's C-order memory directly): the two translation units compile in parallel.
#pragma once

#include "fft_core.cuh"

struct zafb_mdct_plan {
    int64_t n = 0, m = 0;
    int log2m = -1;               // power-of-two M only
    float* d_window = nullptr;    // n floats
    float2* d_tw_fft = nullptr;   // W_{M/2}^t, t < M/2
    float2* d_pre = nullptr;      // e^{-i pi m / M}, m < M/2
    float2* d_post = nullptr;     // e^{-i pi (m + 1/4) / M}, m < M/2
    float* d_cos = nullptr;       // direct path: cos(2 pi t / (8M)), t < 8M
    float2* d_tw_4step = nullptr; // n == 2048 / 1024: W_H^{k1*n2} at [k1*32 + n2], H = n/4 (warp kernels)
    int force_kernel = 0;         // 0 auto, 1 generic, 2 warp (tests)
};

namespace zafb {

// N = 4096 (M = 2048, 1024-point FFT, warp_fft1024), N = 2048 (M = 1024, 512-point FFT, warp_fft512),
// N = 1024 (M = 512, 256-point FFT, warp_fft256) or N = 512 (M = 256, 128-point FFT, warp_fft128)
template <int N>
struct MdctGeom {
    static_assert(N == 512 || N == 1024 || N == 2048 || N == 4096, "mdct warp kernels exist for window lengths 512 ... 4096");
    static constexpr int NTQ = N == 1024 ? 8 : N == 512 ? 12 : 1;  // per-lane twiddles of the warp FFT
    static constexpr int M = N / 2;          // coefficients per frame
    static constexpr int H = M / 2;          // complex FFT length
    static constexpr int REGS = H / 32;      // float2 per lane
    static constexpr int LOGR = clog2(REGS);
    static constexpr int Q = M / 4;          // quarter of the frame, in sample pairs
    static constexpr int TWDEN = M / 16;     // pre[lane + 32 r] = pre[lane] W_TWDEN^r (e^{-i pi 32 r / M})
    static constexpr int TABLES = M + H;     // float2: window pairs, W_H four-step table
    static constexpr int TILE = REGS * kFft1024Pitch;
};

template <int N>
__device__ __forceinline__ void mdct_warp_fft(float2 (&v)[MdctGeom<N>::REGS], const float2* __restrict__ tw, float2* buf, int lane,
                                              const float2 (&tq)[MdctGeom<N>::NTQ]) {
    if constexpr (N == 4096) warp_fft1024<false>(v, tw, buf, lane);
    else if constexpr (N == 2048) warp_fft512(v, tw, buf, lane);
    else if constexpr (N == 1024) warp_fft256(v, tw, buf, lane, tq);
    else warp_fft128(v, tw, buf, lane, tq);
}


}  // namespace zafb
