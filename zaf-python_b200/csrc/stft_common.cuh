// Definitions shared by stft.cu (plans, forward kernels, entry points) and istft.cu (the inverse warp kernels): the two
// translation units compile in parallel -- together they were a 6-minute ptxas run.
#pragma once

#include "fft_core.cuh"

struct zafb_stft_plan {
    int64_t n = 0, hop = 0;
    int log2n = -1;          // -1 if n is not a power of two
    float* d_window = nullptr;      // n floats: fp32(window)
    float2* d_window_half = nullptr;  // n/2 float2: 0.5 * window pairs (power-of-two n only)
    float2* d_tw_half = nullptr;    // W_{n/2}^t, t < n/2
    float2* d_tw_full = nullptr;    // W_n^t, t < n
    float2* d_tw_4step = nullptr;   // n == 2048 / 1024: W_{n/2}^{k1*n2} at [k1*32 + n2]
    double gain = 1.0;              // sum(w[0:N:hop]) accumulated like Python's builtin sum (zaf.py:241)
    int force_kernel = 0;           // 0 auto, 1 generic, 2 warp (tests)
};

namespace zafb {

#ifndef ZAFB_STFT256_CTAS
#define ZAFB_STFT256_CTAS 4  // CTAs per SM of stft_warp_kernel<256>: 3 -> 2.56 ms, 4 -> 2.44, 5 -> 2.50, 6 -> 2.78 (profiles/r01t_stft256_occupancy.txt)
#endif
#ifndef ZAFB_ISTFT_SMALL_CTAS
#define ZAFB_ISTFT_SMALL_CTAS 3  // CTAs per SM of istft_warp_kernel<256> (2 -> 2.94 ms, 3 -> 2.86; N = 512 is slower with 3)
#endif
#ifndef ZAFB_STFT4096_WARPS
#define ZAFB_STFT4096_WARPS 8    // warps per CTA of stft_warp_kernel<4096> (one CTA per SM; 10 warps spill: 4.50 -> 5.31 ms)
#endif
constexpr int kMaxDynSmem = 227 * 1024;  // the sm_100 opt-in maximum per CTA

// One warp transforms M = N / 2 complex points held in its registers (REGS = M / 32 per lane): N = 2048 -> warp_fft1024,
// N = 1024 -> warp_fft512.  In: v[r] = z[lane + 32 r].  Out: Z[lane + 32 k2] = v[bitrev(k2, log2 REGS)].
template <int N>
struct WarpGeom {
    static_assert(N == 256 || N == 512 || N == 1024 || N == 2048 || N == 4096, "warp kernels exist for window lengths 256 ... 4096");
    static constexpr int CTAS_PER_SM = N == 4096 ? 1 : N == 2048 ? 2 : N == 256 ? ZAFB_STFT256_CTAS : 3;  // registers: 2 * REGS of frame state per lane
    static constexpr int M = N / 2;
    static constexpr int REGS = M / 32;
    static constexpr int LOGR = clog2(REGS);
    static constexpr int TILE = REGS * kFft1024Pitch;  // float2 per warp: the transpose tile of the four-step FFT
};

// per-lane constants of the N = 512 transform (warp_fft256); empty for the other sizes
template <int N>
struct LaneTw {
    float2 tq[N == 512 ? 8 : N == 256 ? 12 : 1];
    __device__ __forceinline__ void init(int lane) {
        if constexpr (N == 512) warp_fft256_lane_twiddles(tq, lane);
        if constexpr (N == 256) warp_fft128_lane_twiddles(tq, lane);
    }
};

template <int N>
__device__ __forceinline__ void warp_fft_half(float2 (&v)[N / 64], const float2* __restrict__ tw4, float2* buf, int lane,
                                              const LaneTw<N>& lt) {
    if constexpr (N == 4096) warp_fft2048(v, tw4, buf, lane);
    else if constexpr (N == 2048) warp_fft1024<false>(v, tw4, buf, lane);
    else if constexpr (N == 1024) warp_fft512(v, tw4, buf, lane);
    else if constexpr (N == 512) warp_fft256(v, tw4, buf, lane, lt.tq);
    else warp_fft128(v, tw4, buf, lane, lt.tq);
}


// istft.cu: the warp-per-run ISTFT kernels (frame-major spectra; one-sided and masked variants) and the direct C-order kernel
int istft_warp_dispatch(const zafb_stft_plan* p, const float2* spec, int64_t n_clips, int64_t nt, float* y, int64_t y_stride,
                        cudaStream_t st, int64_t spec_pitch, int onesided, const float* mask, int64_t mask_pitch);
int istft_binmajor_dispatch(const zafb_stft_plan* p, const float2* spec, int64_t n_clips, int64_t nt, float* y, int64_t y_stride,
                            cudaStream_t st);

}  // namespace zafb
