// melspectrogram / mfcc (zaf.py:324-375, 378-454): STFT -> |X| or |X|^2 of bins 1..N/2 -> mel filterbank
// -> [ln(. + eps) -> orthonormal DCT-II over the mel axis, rows 1..n_coef], fused in ONE kernel per call:
// the spectrum never goes to HBM (the reference materialises a complex128 (N, nt) STFT and multiplies by the
// densified, 98.7 %-zero filterbank with dgemm).
//
// The filterbank arrives dense (the wrapper's .toarray(), exactly like zaf.py:373) and is packed at plan
// creation into one contiguous band per row [first non-zero column, last non-zero column]: 882 weights
// instead of 65 536 at BASELINE cfg 3.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "fft_core.cuh"
#include "gemm_tc.cuh"
#include "host_pipe.cuh"
#include "transpose.cuh"

using namespace zafb;

struct zafb_mel_plan {
    int64_t n = 0, hop = 0, n_mels = 0, n_coef = 0;  // n_coef = rows actually produced by mfcc (<= n_mels - 1)
    int log2n = 0;
    float* d_window = nullptr;
    float2* d_tw_half = nullptr;   // W_{N/2}^t
    float2* d_tw_full = nullptr;   // W_N^t, t < N/2
    int* d_band_lo = nullptr;      // per mel row: first column
    int* d_band_len = nullptr;     // number of columns
    int* d_band_off = nullptr;     // offset into d_weights
    float* d_weights = nullptr;
    float* d_dct = nullptr;        // n_coef x n_mels, rows 1..n_coef of the orthonormal DCT-II matrix
    int64_t nnz_packed = 0;
    // N = 1024 warp kernel: mel rows r = lane + 32 g (g < 4) belong to lane `lane`
    bool warp_ok = false;
    int grp_len[4] = {0, 0, 0, 0};  // longest band of each row group
    int grp_off[4] = {0, 0, 0, 0};  // offset of the group's [c][lane] weight block in d_wt
    float* d_wt = nullptr;          // zero-padded band weights, [g][c][lane]
    float2* d_tw_4step = nullptr;   // W_{n/2}^{k1 n2}, [k1][n2]
    float2* d_tw_n = nullptr;       // W_n^t, t < 32
    int* d_lo = nullptr;            // first column of row lane + 32 g, at [g * 32 + lane]
    float* d_dh = nullptr;          // DCT-II half table for lane (c8 = lane & 7, g = lane >> 3): float4 at [(t * qm4 + mq) * 32 + lane]
                                    //   = D[c8 + 8 t + 1][g * 4 qm4 + 4 mq .. +3]  (zero beyond n_coef / half_mels)
    int coef_pad = 0;               // t groups: ceil(n_coef / 8), at most 8
    int qm4 = 0;                    // float4 per quarter of the folded mel axis: ceil(half_mels / 16)
    int half_mels = 0;              // ceil(n_mels / 2)
    int force_kernel = 0;           // 0 auto, 1 generic, 2 warp (tests)
    // float64 route: window and twiddles in double, any power-of-two window length.  precision 0 = automatic (mfcc: fp32,
    // then the frames whose quietest mel band lies below the fp32 floor of the FFT are recomputed in float64), 32, 64.
    int precision = 0;
    double* d_window64 = nullptr;
    double2* d_tw_half64 = nullptr;  // W_{N/2}^t
    double2* d_tw_full64 = nullptr;  // W_N^t, t < N/2
    double2* d_tw_4step64 = nullptr; // N = 1024: W_512^{k1 n2}, [k1][n2]
    // tensor-core route (ZAFB_MEL_ROUTE_TENSOR): the filterbank and the MFCC DCT rows as dense TF32 hi/lo operands
    int route = 0;
    float* d_fb_hi = nullptr;       // n_mels x (n/2)
    float* d_fb_lo = nullptr;
    float* d_dct_hi = nullptr;      // n_coef x ld_mel
    float* d_dct_lo = nullptr;
    int64_t ld_mel = 0;             // n_mels rounded up to 4
    // window lengths that are not powers of two (the reference accepts any, zaf.py:369 -> stft): the any-length STFT kernels
    // write the spectrum to scratch and mel_from_spectrum_kernel applies |.|, the filterbank and the logarithm / DCT
    zafb_stft_plan* stft_any = nullptr;
};

namespace {

constexpr int kMaxDynSmem = 200 * 1024;

// mode 0: melspectrogram (magnitude); mode 1: mfcc (power -> log -> DCT)
__global__ void mel_frame_kernel(const float* __restrict__ x, int64_t ns, int64_t clip_stride, int64_t nt, int64_t hop,
                                 int log2n, const float* __restrict__ window, const float2* __restrict__ tw_half,
                                 const float2* __restrict__ tw_full, const int* __restrict__ band_lo,
                                 const int* __restrict__ band_len, const int* __restrict__ band_off,
                                 const float* __restrict__ weights, const float* __restrict__ dct, int n_mels, int n_coef,
                                 int mode, float* __restrict__ out, int layout, int64_t total_frames,
                                 int* __restrict__ risk_count, int* __restrict__ risk_list, float risk_ratio) {
    extern __shared__ float2 smem2[];
    const int n = 1 << log2n, m = n >> 1;
    float2* a = smem2;
    float2* b = smem2 + m;
    float* spec = reinterpret_cast<float*>(smem2 + 2 * m);  // m floats: column c <-> FFT bin c + 1
    float* mel = spec + m;                                  // n_mels floats
    __shared__ unsigned s_pmax, s_emin;                     // fp32 bit patterns of non-negative values order like integers
    const int tid = threadIdx.x, nth = blockDim.x;
    const int rows = mode == 0 ? n_mels : n_coef;
    for (int64_t f = blockIdx.x; f < total_frames; f += gridDim.x) {
        if (tid == 0) {
            s_pmax = 0u;
            s_emin = 0x7f800000u;
        }
        const int64_t clip = f / nt, j = f - clip * nt;
        const int64_t start = j * hop - m;
        const float* xc = x + clip * clip_stride;
        for (int i = tid; i < m; i += nth) {
            const int64_t s = start + 2 * i;
            const float x0 = (s >= 0 && s < ns) ? xc[s] : 0.f;
            const float x1 = (s + 1 >= 0 && s + 1 < ns) ? xc[s + 1] : 0.f;
            a[i] = make_float2(x0 * window[2 * i], x1 * window[2 * i + 1]);
        }
        __syncthreads();
        const float2* z = block_fft(a, b, tw_half, log2n - 1, tid, nth);
        // bins k = 1 .. N/2 (zaf.py:370: no DC, with Nyquist)
        for (int k = 1 + tid; k <= m; k += nth) {
            float re, im;
            if (k == m) {
                re = z[0].x - z[0].y;
                im = 0.f;
            } else {
                const float2 zk = z[k], zp = z[m - k];
                const float2 e = make_float2(0.5f * (zk.x + zp.x), 0.5f * (zk.y - zp.y));
                const float2 od = make_float2(0.5f * (zk.y + zp.y), 0.5f * (zp.x - zk.x));
                const float2 t = cmul(tw_full[k], od);
                re = e.x + t.x;
                im = e.y + t.y;
            }
            const float p = re * re + im * im;
            spec[k - 1] = mode == 0 ? sqrtf(p) : p;
            if (risk_count != nullptr) atomicMax(&s_pmax, __float_as_uint(p));
        }
        __syncthreads();
        for (int r = tid; r < n_mels; r += nth) {
            const int lo = band_lo[r], len = band_len[r];
            const float* w = weights + band_off[r];
            float acc = 0.f;
            for (int c = 0; c < len; ++c) acc = fmaf(w[c], spec[lo + c], acc);
            if (mode == 0) {
                if (layout == ZAFB_LAYOUT_FRAME_MAJOR) out[f * n_mels + r] = acc;
                else out[(clip * n_mels + r) * nt + j] = acc;
            } else {
                mel[r] = acc + 2.220446049250313e-16f;  // np.finfo(float).eps, zaf.py:445
                if (risk_count != nullptr && acc > 0.f) atomicMin(&s_emin, __float_as_uint(acc));  // exact zeros are ln(eps) in either precision
            }
        }
        if (mode == 1) {
            __syncthreads();
            // a mel band below risk_ratio x the strongest bin sits at the fp32 floor of this FFT: the frame is queued for the
            // float64 kernel that runs next on the stream and overwrites this frame's coefficients
            if (risk_count != nullptr && tid == 0 && __uint_as_float(s_emin) < risk_ratio * __uint_as_float(s_pmax))
                risk_list[atomicAdd(risk_count, 1)] = int(f);
            // ln(mel_r) - ln(mel_0) instead of ln(mel_r): rows k >= 1 of the DCT-II matrix sum to zero, so the
            // constant drops out exactly, and the log of a ratio keeps ~1e-7 absolute accuracy where the log
            // itself (values ~10) only has ~1e-6 in fp32.
            const float ref0 = mel[0];
            __syncthreads();
            for (int r = tid; r < n_mels; r += nth) mel[r] = logf(mel[r] / ref0);
            __syncthreads();
            for (int i = tid; i < n_coef; i += nth) {
                const float* d = dct + int64_t(i) * n_mels;
                float acc0 = 0.f, acc1 = 0.f;
                int r = 0;
                for (; r + 1 < n_mels; r += 2) {
                    acc0 = fmaf(d[r], mel[r], acc0);
                    acc1 = fmaf(d[r + 1], mel[r + 1], acc1);
                }
                if (r < n_mels) acc0 = fmaf(d[r], mel[r], acc0);
                const float v = acc0 + acc1;
                if (layout == ZAFB_LAYOUT_FRAME_MAJOR) out[f * rows + i] = v;
                else out[(clip * rows + i) * nt + j] = v;
            }
        }
        __syncthreads();
    }
}


// ------------------------------------------------------------------------------------------
// float64 route: the same pipeline as mel_frame_kernel with the window multiply, the FFT, |X|^2, the filterbank
// sums and the logarithm in double precision (B200 runs FP64 at half the FP32 rate).  An fp32 FFT resolves a bin
// only to ~1e-8 of the spectral peak, which the LOG in mfcc turns into 1e-4..1e-3 errors on the mel bands of
// purely tonal signals that lie 100+ dB below the peak; this route keeps them at 1e-7.  Inputs and outputs stay fp32.
// ------------------------------------------------------------------------------------------
__global__ void mel_frame_kernel_f64(const float* __restrict__ x, int64_t ns, int64_t clip_stride, int64_t nt, int64_t hop,
                                     int log2n, const double* __restrict__ window, const double2* __restrict__ tw_half,
                                     const double2* __restrict__ tw_full, const int* __restrict__ band_lo,
                                     const int* __restrict__ band_len, const int* __restrict__ band_off,
                                     const float* __restrict__ weights, const float* __restrict__ dct, int n_mels, int n_coef,
                                     int mode, float* __restrict__ out, int layout, int64_t total_frames,
                                     const int* __restrict__ list, const int* __restrict__ list_count) {
    extern __shared__ double2 smem_d[];
    const int n = 1 << log2n, m = n >> 1;
    double2* a = smem_d;
    double2* b = smem_d + m;
    double* spec = reinterpret_cast<double*>(smem_d + 2 * m);  // m doubles: column c <-> FFT bin c + 1
    double* mel = spec + m;                                    // n_mels doubles
    const int tid = threadIdx.x, nth = blockDim.x;
    const int rows = mode == 0 ? n_mels : n_coef;
    const int64_t n_items = list != nullptr ? int64_t(*list_count) : total_frames;  // all frames, or the queued ones
    for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int64_t f = list != nullptr ? int64_t(list[it]) : it;
        const int64_t clip = f / nt, j = f - clip * nt;
        const int64_t start = j * hop - m;
        const float* xc = x + clip * clip_stride;
        for (int i = tid; i < m; i += nth) {
            const int64_t s = start + 2 * i;
            const double x0 = (s >= 0 && s < ns) ? double(xc[s]) : 0.0;
            const double x1 = (s + 1 >= 0 && s + 1 < ns) ? double(xc[s + 1]) : 0.0;
            a[i] = make_double2(x0 * window[2 * i], x1 * window[2 * i + 1]);
        }
        __syncthreads();
        const double2* z = block_fft(a, b, tw_half, log2n - 1, tid, nth);
        for (int k = 1 + tid; k <= m; k += nth) {  // bins 1 .. N/2 (zaf.py:370: no DC, with Nyquist)
            double re, im;
            if (k == m) {
                re = z[0].x - z[0].y;
                im = 0.0;
            } else {
                const double2 zk = z[k], zp = z[m - k];
                const double2 e = make_double2(0.5 * (zk.x + zp.x), 0.5 * (zk.y - zp.y));
                const double2 od = make_double2(0.5 * (zk.y + zp.y), 0.5 * (zp.x - zk.x));
                const double2 t = cmul(tw_full[k], od);
                re = e.x + t.x;
                im = e.y + t.y;
            }
            const double p = re * re + im * im;
            spec[k - 1] = mode == 0 ? sqrt(p) : p;
        }
        __syncthreads();
        for (int r = tid; r < n_mels; r += nth) {
            const int lo = band_lo[r], len = band_len[r];
            const float* w = weights + band_off[r];
            double acc = 0.0;
            for (int c = 0; c < len; ++c) acc = fma(double(w[c]), spec[lo + c], acc);
            if (mode == 0) {
                if (layout == ZAFB_LAYOUT_FRAME_MAJOR) out[f * n_mels + r] = float(acc);
                else out[(clip * n_mels + r) * nt + j] = float(acc);
            } else {
                mel[r] = log(acc + 2.220446049250313e-16);  // np.finfo(float).eps, zaf.py:445
            }
        }
        if (mode == 1) {
            __syncthreads();
            for (int i = tid; i < n_coef; i += nth) {
                const float* d = dct + int64_t(i) * n_mels;
                double acc = 0.0;
                for (int r = 0; r < n_mels; ++r) acc = fma(double(d[r]), mel[r], acc);
                if (layout == ZAFB_LAYOUT_FRAME_MAJOR) out[f * rows + i] = float(acc);
                else out[(clip * rows + i) * nt + j] = float(acc);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// N = 1024 (BASELINE cfg 3): one warp per frame, no block-level synchronisation.
//   load 512 sample pairs (coalesced 8-byte loads) * window pairs held in REGISTERS (they are the
//   same for every frame a lane processes) -> warp_fft512 -> real-input unpack with one xor-style
//   shuffle per bin -> |X|^2 (or |X|) of bins 1..512 into the warp's shared-memory tile ->
//   banded filterbank, lane l owning mel rows l, l+32, l+64, l+96 (weights zero-padded per row
//   group, [g][c][lane] so the weight loads are conflict-free) -> melspectrogram rows, or
//   ln ratio -> DCT-II through its even/odd symmetry  C[k] = sum_{m < n/2} D[k][m] (L[m] +- L[n-1-m])
//   with lane l owning coefficients l and l + 32.
// ------------------------------------------------------------------------------------------
#ifndef ZAFB_MEL_OCC1024
#define ZAFB_MEL_OCC1024 3
#endif
constexpr int kMelOcc1024 = ZAFB_MEL_OCC1024;   // CTAs per SM at N = 1024 (3 needs the window in shared memory)
constexpr int kWarps = 8;        // N = 1024
constexpr int kWarps2048 = 6;    // N = 2048: the tiles are twice as large; 6 warps keep two CTAs per SM
constexpr int kMelWarpTile = 16 * kFft1024Pitch;  // float2 per warp at N = 1024: FFT transpose tile, then spectrum / log-mel scratch

__device__ __forceinline__ void split_tf32_dev(float v, float& hi, float& lo) {
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    hi = __uint_as_float(h);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(v - hi));
    lo = __uint_as_float(l);
}

// N = 1024 (BASELINE cfg 3) or 2048 (the reference's own example, zaf.py:347-357: 44.1 kHz, 0.04 s window).
// SPLIT = true (N = 1024): the kernel stops after the spectrum and writes |X| (MODE 0) or |X|^2 (MODE 1) of bins 1..512 as
// TF32 hi/lo halves to out / out_lo ([frame][512]) -- the A operand of the tensor-core filterbank product.
template <int N, int MODE, bool SPLIT = false>  // MODE 0 melspectrogram, 1 mfcc
__global__ void __launch_bounds__((N == 2048 ? kWarps2048 : kWarps) * 32, N == 1024 ? kMelOcc1024 : 2)
mel_warp_kernel(const float* __restrict__ x, int64_t ns, int64_t clip_stride, int64_t nt, int hop,
                const float2* __restrict__ win_pairs, const float2* __restrict__ tw4,
                const float2* __restrict__ tw_full, const float* __restrict__ wt, const int* __restrict__ lo_tab,
                int4 grp_len, int4 grp_off, int wt_total, const float4* __restrict__ dh, int n_mels, int half_mels,
                int n_coef, int dct_shape, float* __restrict__ out, int64_t total_frames,
                float* __restrict__ out_lo, int* __restrict__ risk_count, int* __restrict__ risk_list, float risk_ratio) {
    const int tgroups = dct_shape >> 8, qm4 = dct_shape & 255;  // MFCC DCT: coefficient groups of 8, float4 per mel quarter
    constexpr int M = N / 2, REGS = M / 32, LOGR = clog2(REGS);
    static_assert(N == 512 || N == 1024 || N == 2048, "warp kernels exist for window lengths 512, 1024 and 2048");
    constexpr int WARPS = N == 2048 ? kWarps2048 : kWarps;
    // float2 per warp: the FFT transpose tile, then the spectrum (M floats) / log-mel scratch (256 floats at least)
    constexpr int TILE = (REGS < 16 ? 16 : REGS) * kFft1024Pitch;
    // window pairs in registers unless they do not fit (N = 2048: 32 more float2; N = 1024 at three CTAs per SM)
    constexpr bool WIN_REGS = N == 512 || (N == 1024 && kMelOcc1024 == 2);
    static_assert(!SPLIT || N == 1024, "the tensor-core front end is written for N = 1024");
    extern __shared__ float2 smem2[];
    float2* s_tw = smem2;                                    // M: W_M^{k1 n2}
    float2* s_win = smem2 + M;                               // M (N = 2048 only): 0.5 * window pairs
    float* s_wt = reinterpret_cast<float*>(smem2 + (WIN_REGS ? M : 2 * M));  // wt_total floats
    float4* s_dh = reinterpret_cast<float4*>(s_wt + ((wt_total + 3) & ~3));  // tgroups * qm4 * 32 float4
    const int dh_count = MODE == 1 ? tgroups * qm4 * 32 : 0;
    float2* s_warp = reinterpret_cast<float2*>(s_dh + dh_count);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2* s_buf = s_warp + warp * TILE;
    float* s_spec = reinterpret_cast<float*>(s_buf);         // M floats (+ slack) once the FFT is done
    for (int i = tid; i < M; i += WARPS * 32) s_tw[i] = tw4[i];
    if constexpr (!WIN_REGS)
        for (int i = tid; i < M; i += WARPS * 32) {
            const float2 w = win_pairs[i];
            s_win[i] = make_float2(0.5f * w.x, 0.5f * w.y);
        }
    for (int i = tid; i < wt_total; i += WARPS * 32) s_wt[i] = wt[i];
    for (int i = tid; i < dh_count; i += WARPS * 32) s_dh[i] = dh[i];
    float2 win[WIN_REGS ? REGS : 1];
    if constexpr (WIN_REGS) {
#pragma unroll
        for (int r = 0; r < REGS; ++r) {
            const float2 w = win_pairs[lane + 32 * r];
            win[r] = make_float2(0.5f * w.x, 0.5f * w.y);  // the 1/2 of the real-input split, exact in fp32
        }
    }
    const float2 c_lane = tw_full[lane];  // W_N^lane
    float2 tq[N == 512 ? 8 : 1];          // per-lane twiddles of warp_fft256
    if constexpr (N == 512) warp_fft256_lane_twiddles(tq, lane);
    int lo[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) lo[g] = lo_tab[g * 32 + lane];
    const int glen[4] = {grp_len.x, grp_len.y, grp_len.z, grp_len.w};
    const int goff[4] = {grp_off.x, grp_off.y, grp_off.z, grp_off.w};
    __syncthreads();

    for (int64_t f = int64_t(blockIdx.x) * WARPS + warp; f < total_frames; f += int64_t(gridDim.x) * WARPS) {
        const int64_t clip = f / nt, j = f - clip * nt;
        const int64_t start = j * hop - M;
        const float* xc = x + clip * clip_stride;
        float2 v[REGS];
        if (start >= 0 && start + N <= ns) {
            const float2* fp = reinterpret_cast<const float2*>(xc + start) + lane;
#pragma unroll
            for (int r = 0; r < REGS; ++r) v[r] = __ldg(fp + 32 * r);
        } else {
#pragma unroll
            for (int r = 0; r < REGS; ++r) {
                const int64_t s0 = start + 2 * (lane + 32 * r);
                v[r].x = (s0 >= 0 && s0 < ns) ? __ldg(xc + s0) : 0.f;
                v[r].y = (s0 + 1 >= 0 && s0 + 1 < ns) ? __ldg(xc + s0 + 1) : 0.f;
            }
        }
#pragma unroll
        for (int r = 0; r < REGS; ++r) {
            const float2 w = WIN_REGS ? win[WIN_REGS ? r : 0] : s_win[lane + 32 * r];
            v[r].x *= w.x;
            v[r].y *= w.y;
        }
        // Z[lane + 32 k] = v[bitrev(k, LOGR)]
        if constexpr (N == 512) warp_fft256(v, s_tw, s_buf, lane, tq);
        else if constexpr (N == 1024) warp_fft512(v, s_tw, s_buf, lane);
        else warp_fft1024<false>(v, s_tw, s_buf, lane);

        // X[k] = E + W_N^k O,  E = Z[k] + conj(Z[M-k]),  O = -i (Z[k] - conj(Z[M-k])),  k = lane + 32 kap;
        // column c = k - 1 (zaf.py:370 drops DC, keeps Nyquist); lane 0 / kap 0 produces the Nyquist bin instead of DC.
        const int src = (32 - lane) & 31;
        float pmax = 0.f;  // strongest bin of the frame (MODE 1: the reference level of the fp32-floor test below)
        static_for<0, REGS>([&](auto kc) {
            constexpr int kap = decltype(kc)::value;
            const float2 z = v[bitrev(kap, LOGR)];
            const float2 mine = v[bitrev(REGS - 1 - kap, LOGR)];
            float2 pz;
            pz.x = __shfl_sync(0xffffffffu, mine.x, src);
            pz.y = __shfl_sync(0xffffffffu, mine.y, src);
            if (lane == 0) pz = v[bitrev((REGS - kap) & (REGS - 1), LOGR)];
            const float2 e = make_float2(z.x + pz.x, z.y - pz.y);
            const float2 od = make_float2(z.y + pz.y, pz.x - z.x);
            const float2 t = cmul(mul_tw<kap, N / 32>(c_lane), od);
            float2 xk = cadd(e, t);
            if (kap == 0 && lane == 0) xk = csub(e, t);  // X[M] = E[0] - O[0]
            const float p2 = xk.x * xk.x + xk.y * xk.y;
            if constexpr (MODE == 1 && !SPLIT) pmax = fmaxf(pmax, p2);
            // (the spectrum overwrites the transpose tile: every lane finished reading it inside the FFT)
            const int col = (kap == 0 && lane == 0) ? M - 1 : lane + 32 * kap - 1;
            s_spec[col] = MODE == 0 ? sqrtf(p2) : p2;
        });
        __syncwarp();

        if constexpr (SPLIT) {
#pragma unroll
            for (int i = 0; i < REGS; ++i) {
                float hi, lo_part;
                split_tf32_dev(s_spec[lane + 32 * i], hi, lo_part);
                out[f * M + lane + 32 * i] = hi;
                out_lo[f * M + lane + 32 * i] = lo_part;
            }
            __syncwarp();
            continue;
        }

        float mel[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const float* w = s_wt + goff[g] + lane;
            const float* sp = s_spec + lo[g];
            const int len = glen[g];
            float acc = 0.f;
            for (int c = 0; c < len; ++c) acc = fmaf(w[32 * c], sp[c], acc);
            mel[g] = acc;
        }
        __syncwarp();  // every lane is done reading the spectrum

        if constexpr (MODE == 0) {
            float* o = out + f * n_mels;
#pragma unroll
            for (int g = 0; g < 4; ++g)
                if (lane + 32 * g < n_mels) o[lane + 32 * g] = mel[g];
        } else {
            if (risk_count != nullptr) {
                // A mel band below risk_ratio x the strongest bin sits at the fp32 floor of this FFT (a bin is resolved to
                // ~1e-8 of the peak amplitude) and the logarithm would turn that into an absolute error: the frame is queued
                // for the float64 kernel that runs next on the stream and overwrites this frame's coefficients.
                float emin = 3.4e38f;
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    if (lane + 32 * g < n_mels && mel[g] > 0.f) emin = fminf(emin, mel[g]);  // an exact zero (silence, an empty row) is ln(eps) in either precision
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    emin = fminf(emin, __shfl_xor_sync(0xffffffffu, emin, o));
                    pmax = fmaxf(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
                }
                if (lane == 0 && emin < risk_ratio * pmax) risk_list[atomicAdd(risk_count, 1)] = int(f);
            }
            // ln(mel_r + eps) - ln(mel_0 + eps): rows k >= 1 of the DCT-II matrix sum to zero, so the constant drops
            // out exactly, and the log of a ratio keeps fp32 absolute accuracy where the log itself does not.
            const float ref0 = __shfl_sync(0xffffffffu, mel[0], 0) + 2.220446049250313e-16f;
            float* s_log = s_spec;            // n_mels floats
            float* s_sym = s_spec + 128;      // S[m] at [m], A[m] at [64 + m], m < half_mels <= 64
#pragma unroll
            for (int g = 0; g < 4; ++g)
                if (lane + 32 * g < n_mels) s_log[lane + 32 * g] = logf((mel[g] + 2.220446049250313e-16f) / ref0);
            __syncwarp();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = lane + 32 * h;
                float sv = 0.f, av = 0.f;
                if (m < half_mels) {
                    const int mm = n_mels - 1 - m;
                    const float a = s_log[m], b = (mm != m) ? s_log[mm] : 0.f;
                    sv = a + b;
                    av = (mm != m) ? a - b : 0.f;
                }
                s_sym[m] = sv;
                s_sym[64 + m] = av;
            }
            __syncwarp();
            // C[k] = sum_{m < half} D[k][m] (L[m] +- L[n-1-m]).  Lane (c8 = lane & 7, g = lane >> 3) accumulates the
            // coefficients k = c8 + 8 t + 1 (t < tgroups) over quarter g of the folded mel axis -- every lane busy, every
            // table entry read once -- then the four quarters are added with two xor shuffles and lane c8 + 8 g ends up
            // holding coefficient index lane (t = g) and 32 + lane (t = 4 + g).
            const int c8 = lane & 7, g = lane >> 3;
            const float4* sa = reinterpret_cast<const float4*>(s_sym + ((c8 & 1) ? 0 : 64)) + g * qm4;  // k even -> S, k odd -> A
            const float4* d = s_dh + lane;
            float acc[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) acc[t] = 0.f;
            for (int mq = 0; mq < qm4; ++mq) {
                const float4 s4 = sa[mq];
#pragma unroll
                for (int t = 0; t < 8; ++t)
                    if (t < tgroups) {
                        const float4 d4 = d[(t * qm4 + mq) * 32];
                        acc[t] = fmaf(d4.x, s4.x, fmaf(d4.y, s4.y, fmaf(d4.z, s4.z, fmaf(d4.w, s4.w, acc[t]))));
                    }
            }
            float lo_val = 0.f, hi_val = 0.f;
#pragma unroll
            for (int t = 0; t < 8; ++t)
                if (t < tgroups) {
                    float a = acc[t];
                    a += __shfl_xor_sync(0xffffffffu, a, 8);
                    a += __shfl_xor_sync(0xffffffffu, a, 16);
                    if (t < 4) lo_val = (g == t) ? a : lo_val;
                    else hi_val = (g == t - 4) ? a : hi_val;
                }
            if (lane < n_coef) out[f * n_coef + lane] = lo_val;
            if (32 + lane < n_coef) out[f * n_coef + 32 + lane] = hi_val;
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------
// Any window length (not a power of two): the two-sided spectrum comes from the any-length STFT kernels (frame-major
// scratch); one CTA per frame applies |X| or |X|^2 of bins 1 .. floor(N/2), the banded filterbank and, for mfcc, the
// logarithm and the DCT-II rows -- the tail of mel_frame_kernel on a spectrum read from memory.
// ------------------------------------------------------------------------------------------
__global__ void mel_from_spectrum_kernel(const float2* __restrict__ spec, int n, int64_t nt, int64_t frame0,
                                         const int* __restrict__ band_lo, const int* __restrict__ band_len,
                                         const int* __restrict__ band_off, const float* __restrict__ weights,
                                         const float* __restrict__ dct, int n_mels, int n_coef, int mode, float* __restrict__ out,
                                         int layout, int64_t frames) {
    extern __shared__ float smem_f[];
    const int m = n / 2;
    float* sp = smem_f;      // m floats: column c <-> bin c + 1
    float* mel = sp + m;     // n_mels floats
    const int tid = threadIdx.x, nth = blockDim.x;
    const int rows = mode == 0 ? n_mels : n_coef;
    for (int64_t fl = blockIdx.x; fl < frames; fl += gridDim.x) {
        const float2* X = spec + fl * n;
        const int64_t f = frame0 + fl, clip = f / nt, j = f - clip * nt;
        for (int c = tid; c < m; c += nth) {
            const float2 v = X[c + 1];
            const float p = v.x * v.x + v.y * v.y;
            sp[c] = mode == 0 ? sqrtf(p) : p;
        }
        __syncthreads();
        for (int r = tid; r < n_mels; r += nth) {
            const int lo = band_lo[r], len = band_len[r];
            const float* w = weights + band_off[r];
            float acc = 0.f;
            for (int c = 0; c < len; ++c) acc = fmaf(w[c], sp[lo + c], acc);
            if (mode == 0) {
                if (layout == ZAFB_LAYOUT_FRAME_MAJOR) out[f * n_mels + r] = acc;
                else out[(clip * n_mels + r) * nt + j] = acc;
            } else {
                mel[r] = acc + 2.220446049250313e-16f;  // np.finfo(float).eps, zaf.py:445
            }
        }
        if (mode == 1) {
            __syncthreads();
            const float ref0 = mel[0];  // log of a ratio: see mel_frame_kernel
            __syncthreads();
            for (int r = tid; r < n_mels; r += nth) mel[r] = logf(mel[r] / ref0);
            __syncthreads();
            for (int i = tid; i < n_coef; i += nth) {
                const float* d = dct + int64_t(i) * n_mels;
                float acc0 = 0.f, acc1 = 0.f;
                int r = 0;
                for (; r + 1 < n_mels; r += 2) {
                    acc0 = fmaf(d[r], mel[r], acc0);
                    acc1 = fmaf(d[r + 1], mel[r + 1], acc1);
                }
                if (r < n_mels) acc0 = fmaf(d[r], mel[r], acc0);
                if (layout == ZAFB_LAYOUT_FRAME_MAJOR) out[f * rows + i] = acc0 + acc1;
                else out[(clip * rows + i) * nt + j] = acc0 + acc1;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// N = 1024 in double precision, one warp per frame: the mfcc frames the fp32 kernel queued (list != nullptr), or every
// frame (precision = 64).  The pipeline of mel_warp_kernel with window, FFT (warp_fft512_f64), unpack, |X|^2 (or |X|),
// filterbank sums, logarithm and DCT-II sums in FP64 (B200: half the FP32 rate); inputs, weights and outputs stay fp32.
// ------------------------------------------------------------------------------------------
constexpr int kWarps64 = 8;
constexpr int kMelTile64 = 16 * kFft1024Pitch;  // double2 per warp

__global__ void __launch_bounds__(kWarps64 * 32, 1)
mel_warp_kernel_f64(const float* __restrict__ x, int64_t ns, int64_t clip_stride, int64_t nt, int hop,
                    const double* __restrict__ window, const double2* __restrict__ tw4, const double2* __restrict__ tw_full,
                    const int* __restrict__ band_lo, const int* __restrict__ band_len, const int* __restrict__ band_off,
                    const float* __restrict__ weights, const float* __restrict__ dct, int n_mels, int n_coef, int mode,
                    float* __restrict__ out, int64_t total_frames, const int* __restrict__ list,
                    const int* __restrict__ list_count) {
    constexpr int N = 1024, M = 512, REGS = 16, LOGR = 4;
    extern __shared__ double2 smem_d[];
    double2* s_tw = smem_d;                                   // 512: W_512^{k1 n2}
    double2* s_win = smem_d + M;                              // 512: 0.5 * window pairs
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double2* s_buf = smem_d + 2 * M + warp * kMelTile64;
    double* s_spec = reinterpret_cast<double*>(s_buf);        // 512 doubles once the FFT is done, then 128 log-mel values
    for (int i = tid; i < M; i += kWarps64 * 32) {
        s_tw[i] = tw4[i];
        s_win[i] = make_double2(0.5 * window[2 * i], 0.5 * window[2 * i + 1]);
    }
    const double2 c_lane = tw_full[lane];  // W_N^lane
    __syncthreads();
    const int rows = mode == 0 ? n_mels : n_coef;
    const int64_t n_items = list != nullptr ? int64_t(*list_count) : total_frames;
    for (int64_t it = int64_t(blockIdx.x) * kWarps64 + warp; it < n_items; it += int64_t(gridDim.x) * kWarps64) {
        const int64_t f = list != nullptr ? int64_t(list[it]) : it;
        const int64_t clip = f / nt, j = f - clip * nt;
        const int64_t start = j * hop - M;
        const float* xc = x + clip * clip_stride;
        double2 v[REGS];
#pragma unroll
        for (int r = 0; r < REGS; ++r) {
            const int64_t s0 = start + 2 * (lane + 32 * r);
            const double x0 = (s0 >= 0 && s0 < ns) ? double(__ldg(xc + s0)) : 0.0;
            const double x1 = (s0 + 1 >= 0 && s0 + 1 < ns) ? double(__ldg(xc + s0 + 1)) : 0.0;
            const double2 w = s_win[lane + 32 * r];
            v[r] = make_double2(x0 * w.x, x1 * w.y);
        }
        warp_fft512_f64(v, s_tw, s_buf, lane);  // Z[lane + 32 k] = v[bitrev(k, 4)]
        const int src = (32 - lane) & 31;
        static_for<0, REGS>([&](auto kc) {
            constexpr int kap = decltype(kc)::value;
            const double2 z = v[bitrev(kap, LOGR)];
            const double2 mine = v[bitrev(REGS - 1 - kap, LOGR)];
            double2 pz;
            pz.x = __shfl_sync(0xffffffffu, mine.x, src);
            pz.y = __shfl_sync(0xffffffffu, mine.y, src);
            if (lane == 0) pz = v[bitrev((REGS - kap) & (REGS - 1), LOGR)];
            const double2 e = make_double2(z.x + pz.x, z.y - pz.y);
            const double2 od = make_double2(z.y + pz.y, pz.x - z.x);
            const double2 t = cmul(mul_tw<kap, N / 32>(c_lane), od);
            double2 xk = cadd(e, t);
            if (kap == 0 && lane == 0) xk = csub(e, t);  // X[M] = E[0] - O[0]
            const double p2 = xk.x * xk.x + xk.y * xk.y;
            const int col = (kap == 0 && lane == 0) ? M - 1 : lane + 32 * kap - 1;
            s_spec[col] = mode == 0 ? sqrt(p2) : p2;
        });
        __syncwarp();
        double mel[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int r = lane + 32 * g;
            double acc = 0.0;
            if (r < n_mels) {
                const int lo = band_lo[r], len = band_len[r];
                const float* w = weights + band_off[r];
                for (int c = 0; c < len; ++c) acc = fma(double(__ldg(w + c)), s_spec[lo + c], acc);
            }
            mel[g] = acc;
        }
        __syncwarp();  // every lane is done reading the spectrum
        if (mode == 0) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
                if (lane + 32 * g < n_mels) out[f * n_mels + lane + 32 * g] = float(mel[g]);
        } else {
#pragma unroll
            for (int g = 0; g < 4; ++g)
                if (lane + 32 * g < n_mels) s_spec[lane + 32 * g] = log(mel[g] + 2.220446049250313e-16);  // zaf.py:445
            __syncwarp();
            for (int i = lane; i < n_coef; i += 32) {
                const float* d = dct + int64_t(i) * n_mels;
                double a0 = 0.0, a1 = 0.0;
                int r = 0;
                for (; r + 1 < n_mels; r += 2) {
                    a0 = fma(double(__ldg(d + r)), s_spec[r], a0);
                    a1 = fma(double(__ldg(d + r + 1)), s_spec[r + 1], a1);
                }
                if (r < n_mels) a0 = fma(double(__ldg(d + r)), s_spec[r], a0);
                out[f * rows + i] = float(a0 + a1);
            }
            __syncwarp();
        }
    }
}

// tensor route, MFCC: ln((mel_r + eps) / (mel_0 + eps)) of a [frames][n_mels] matrix as TF32 hi/lo halves (pitch ld);
// the common term ln(mel_0 + eps) drops out of DCT rows >= 1 exactly (their weights sum to zero).
__global__ void mel_log_split_kernel(const float* __restrict__ mel, int64_t frames, int n_mels, int64_t ld,
                                     float* __restrict__ hi, float* __restrict__ lo) {
    const int64_t total = frames * ld;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int64_t f = i / ld;
        const int r = int(i - f * ld);
        float h = 0.f, l = 0.f;
        if (r < n_mels) {
            const float ref0 = mel[f * n_mels] + 2.220446049250313e-16f;
            split_tf32_dev(logf((mel[f * n_mels + r] + 2.220446049250313e-16f) / ref0), h, l);
        }
        hi[i] = h;
        lo[i] = l;
    }
}

bool g_attr_done = false;
int set_kernel_attrs() {
    if (g_attr_done) return ZAFB_OK;
    ZAFB_CUDA(cudaFuncSetAttribute(mel_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    ZAFB_CUDA(cudaFuncSetAttribute(mel_frame_kernel_f64, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    ZAFB_CUDA(cudaFuncSetAttribute(mel_warp_kernel_f64, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    ZAFB_CUDA((cudaFuncSetAttribute(mel_warp_kernel<1024, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(mel_warp_kernel<1024, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(mel_warp_kernel<1024, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(mel_warp_kernel<1024, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(mel_warp_kernel<2048, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(mel_warp_kernel<2048, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(mel_warp_kernel<512, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(mel_warp_kernel<512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    g_attr_done = true;
    return ZAFB_OK;
}

template <class T>
int upload_vec(T** dev, const std::vector<T>& v) {
    ZAFB_CUDA(cudaMalloc(reinterpret_cast<void**>(dev), (v.size() ? v.size() : 1) * sizeof(T)));
    ZAFB_CUDA(cudaMemcpy(*dev, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return ZAFB_OK;
}

// ------------------------------------------------------------------------------------------
// Tensor-core route (north star: "the mel filterbank matmul on tensor cores ... a dense contraction"):
//   (1) mel_warp_kernel<1024, MODE, true>  frames -> |X| or |X|^2 of bins 1..512, split into TF32 hi/lo halves
//   (2) gemm3xtf32 (tcgen05 / TMEM / TMA): [frames x 512] . filterbank[n_mels x 512]^T -> mel [frames x n_mels]
//   (3) MFCC only: log-ratio + split, then gemm3xtf32 with the DCT-II rows 1..n_coef
// in chunks of 16 384 frames so that the 64 MB of split spectra stay in the 126 MB L2 between (1) and (2).
// The fused banded-FMA kernel above does 74x fewer multiply-adds and never writes the spectrum; it is the
// default and the faster one (DESIGN.md section 4.3 has both timings) -- this route is the dense form.
// ------------------------------------------------------------------------------------------
int launch_tensor(const zafb_mel_plan* p, int mode, const float* x, int64_t n_clips, int64_t ns, int64_t clip_stride,
                  int64_t nt, float* out, cudaStream_t st) {
    static bool pool_ready = false;
    if (!pool_ready) {  // keep the stream-ordered scratch in the pool between calls
        int dev = 0;
        cudaMemPool_t pool;
        uint64_t keep = UINT64_MAX;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        pool_ready = true;
    }
    const int64_t total = n_clips * nt;
    const int64_t clips_per_chunk = std::max<int64_t>(1, 16384 / (nt > 0 ? nt : 1));
    const int64_t chunk_frames = std::min(total, clips_per_chunk * nt);
    const int64_t n_mels = p->n_mels, ld = p->ld_mel, n_coef = p->n_coef;
    // scratch: spectra hi/lo, and for MFCC the mel matrix and its log hi/lo
    auto round64 = [](size_t v) { return (v + 63) & ~size_t(63); };  // keep every segment 256-byte aligned (TMA: 16)
    const size_t spec_f = round64(size_t(chunk_frames) * 512);
    const size_t mel_f = mode == 1 ? round64(size_t(chunk_frames) * size_t(n_mels)) : 0;
    const size_t log_f = mode == 1 ? round64(size_t(chunk_frames) * size_t(ld)) : 0;
    float* ws = nullptr;
    ZAFB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ws), (2 * spec_f + mel_f + 2 * log_f) * sizeof(float), st));
    float *spec_hi = ws, *spec_lo = ws + spec_f, *mel_buf = ws + 2 * spec_f, *log_hi = mel_buf + mel_f, *log_lo = log_hi + log_f;
    const size_t smem = (kMelOcc1024 == 2 ? 512 : 1024) * sizeof(float2) + 16 + size_t(kWarps) * kMelWarpTile * sizeof(float2);
    const int4 zero4 = make_int4(0, 0, 0, 0);
    int rc = ZAFB_OK;
    for (int64_t c0 = 0; c0 < n_clips && rc == ZAFB_OK; c0 += clips_per_chunk) {
        const int64_t nc = std::min(clips_per_chunk, n_clips - c0);
        const int64_t frames = nc * nt;
        int64_t ctas = ceil_div(frames, kWarps);
        if (ctas > int64_t(sm_count()) * 2) ctas = int64_t(sm_count()) * 2;
        auto kern = mode == 0 ? mel_warp_kernel<1024, 0, true> : mel_warp_kernel<1024, 1, true>;
        kern<<<unsigned(ctas), kWarps * 32, smem, st>>>(x + c0 * clip_stride, ns, clip_stride, nt, int(p->hop),
                                                         reinterpret_cast<const float2*>(p->d_window), p->d_tw_4step, p->d_tw_n,
                                                         p->d_wt, p->d_lo, zero4, zero4, 0, nullptr, int(n_mels), 0, 0, 0, spec_hi, frames,
                                                         spec_lo, nullptr, nullptr, 0.f);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        float* mel_out = mode == 0 ? out + c0 * nt * n_mels : mel_buf;
        rc = gemm3xtf32(spec_hi, spec_lo, 512, p->d_fb_hi, p->d_fb_lo, 512, mel_out, n_mels, frames, n_mels, 512, st);
        if (rc == ZAFB_OK && mode == 1) {
            int64_t blocks = ceil_div(frames * ld, 256);
            if (blocks > int64_t(sm_count()) * 16) blocks = int64_t(sm_count()) * 16;
            mel_log_split_kernel<<<unsigned(blocks), 256, 0, st>>>(mel_buf, frames, int(n_mels), ld, log_hi, log_lo);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            rc = gemm3xtf32(log_hi, log_lo, ld, p->d_dct_hi, p->d_dct_lo, ld, out + c0 * nt * n_coef, n_coef, frames, n_coef,
                            n_mels, st);
        }
    }
    cudaFreeAsync(ws, st);
    if (rc == ZAFB_OK) ZAFB_CUDA(cudaGetLastError());
    return rc;
}

int launch(const zafb_mel_plan* p, int mode, const float* x, int64_t n_clips, int64_t ns, int64_t clip_stride, float* out,
           int layout, void* stream) {
    ZAFB_REQUIRE(p != nullptr, "plan is NULL");
    ZAFB_REQUIRE(n_clips >= 0 && ns >= 0 && clip_stride >= ns, "bad batch geometry");
    ZAFB_REQUIRE(layout == ZAFB_LAYOUT_FRAME_MAJOR || layout == ZAFB_LAYOUT_BIN_MAJOR, "bad layout %d", layout);
    int rc = set_kernel_attrs();
    if (rc != ZAFB_OK) return rc;
    int64_t nt = 0;
    zafb_stft_geometry(ns, p->n, p->hop, nullptr, &nt, nullptr);
    const int64_t total = n_clips * nt;
    const int64_t rows = mode == 0 ? p->n_mels : p->n_coef;
    if (total == 0 || rows == 0) return ZAFB_OK;
    ZAFB_REQUIRE(out != nullptr && (x != nullptr || ns == 0), "x/out is NULL");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (p->stft_any != nullptr) {  // window length not a power of two: spectrum through scratch, at most ~512 MB at a time
        const size_t clip_bytes = size_t(nt) * p->n * sizeof(float2);
        int64_t per = int64_t((size_t(512) << 20) / (clip_bytes ? clip_bytes : 1));
        if (per < 1) per = 1;
        if (per > n_clips) per = n_clips;
        float* scratch = nullptr;
        ZAFB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&scratch), size_t(per) * clip_bytes, st));
        const size_t smem = size_t(p->n / 2 + p->n_mels) * sizeof(float) + 16;
        rc = ZAFB_OK;
        for (int64_t c0 = 0; c0 < n_clips && rc == ZAFB_OK; c0 += per) {
            const int64_t nc = std::min(per, n_clips - c0), frames = nc * nt;
            rc = zafb_stft_f32(p->stft_any, x + c0 * clip_stride, nc, ns, clip_stride, scratch, ZAFB_LAYOUT_FRAME_MAJOR, stream);
            if (rc != ZAFB_OK) break;
            const int64_t grid = frames < int64_t(sm_count()) * 16 ? frames : int64_t(sm_count()) * 16;
            mel_from_spectrum_kernel<<<unsigned(grid), 128, smem, st>>>(
                reinterpret_cast<const float2*>(scratch), int(p->n), nt, c0 * nt, p->d_band_lo, p->d_band_len, p->d_band_off,
                p->d_weights, p->d_dct, int(p->n_mels), int(p->n_coef), mode, out, layout, frames);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            if (cudaGetLastError() != cudaSuccess) rc = fail(ZAFB_E_CUDA, "mel_from_spectrum_kernel launch failed");
        }
        cudaFreeAsync(scratch, st);
        return rc;
    }
    {
        const bool aligned = reinterpret_cast<uintptr_t>(x) % 8 == 0 && (n_clips <= 1 || clip_stride % 2 == 0) && p->hop % 2 == 0;
        if (layout == ZAFB_LAYOUT_BIN_MAJOR && p->warp_ok && aligned && p->force_kernel != 1) {
            // the reference's C-order memory: the frame-major kernels into scratch, then a tiled transpose
            return bin_major_from_frame_major(out, n_clips, nt, rows, st, [&](int64_t c0, int64_t n, float* scratch) {
                return launch(p, mode, x + c0 * clip_stride, n, ns, clip_stride, scratch, ZAFB_LAYOUT_FRAME_MAJOR, stream);
            });
        }
    }
    // float64 kernels: every frame (list == nullptr) or the frames queued in `list` (count read on the device)
    auto launch_f64 = [&](float* o, int lay, const int* list, const int* list_count) -> int {
        if (p->n == 1024 && p->n_mels <= 128 && lay == ZAFB_LAYOUT_FRAME_MAJOR && p->d_tw_4step64 != nullptr) {
            const size_t smem64 = size_t(1024 + kWarps64 * kMelTile64) * sizeof(double2);
            int64_t ctas = list != nullptr ? int64_t(sm_count()) : ceil_div(total, kWarps64);
            if (ctas > int64_t(sm_count())) ctas = int64_t(sm_count());
            mel_warp_kernel_f64<<<unsigned(ctas), kWarps64 * 32, smem64, st>>>(
                x, ns, clip_stride, nt, int(p->hop), p->d_window64, p->d_tw_4step64, p->d_tw_full64, p->d_band_lo, p->d_band_len,
                p->d_band_off, p->d_weights, p->d_dct, int(p->n_mels), int(p->n_coef), mode, o, total, list, list_count);
            ZAFB_LAUNCH_CHECK();
            return ZAFB_OK;
        }
        const int mh = int(p->n / 2);
        const size_t smem64 = size_t(p->n) * sizeof(double2) + size_t(mh + p->n_mels) * sizeof(double) + 16;
        if (smem64 > size_t(kMaxDynSmem)) return fail(ZAFB_E_UNSUPPORTED, "mel float64 route: window_length %lld too large", (long long)p->n);
        int th = mh / 4;
        if (th < 64) th = 64;
        if (th > 256) th = 256;
        const int64_t grid = total < int64_t(sm_count()) * 16 ? total : int64_t(sm_count()) * 16;
        mel_frame_kernel_f64<<<unsigned(grid), th, smem64, st>>>(
            x, ns, clip_stride, nt, p->hop, p->log2n, p->d_window64, p->d_tw_half64, p->d_tw_full64, p->d_band_lo, p->d_band_len,
            p->d_band_off, p->d_weights, p->d_dct, int(p->n_mels), int(p->n_coef), mode, o, lay, total, list, list_count);
        ZAFB_LAUNCH_CHECK();
        return ZAFB_OK;
    };
    if (p->precision == 64) return launch_f64(out, layout, nullptr, nullptr);
    // precision 0 (automatic), mfcc only: the fp32 kernels queue the frames whose quietest mel band lies below the fp32
    // floor of their FFT (risk_ratio x the strongest bin's power); the float64 kernel then recomputes exactly those
    // frames, stream-ordered, no host round trip.  Scratch: a counter and one int per frame from the stream's pool.
    const bool auto_f64 = mode == 1 && p->precision == 0 && p->route != ZAFB_MEL_ROUTE_TENSOR && total < (int64_t(1) << 31) &&
                          p->d_window64 != nullptr;
    int* risk = nullptr;
    const float risk_ratio = 1e-8f;  // -80 dB: a band that far below the strongest bin has ~1e-4 relative error in fp32 (the bin floor is ~1e-8 of the peak amplitude)
    if (auto_f64) {
        static bool pool_ready = false;
        if (!pool_ready) {  // keep the stream-ordered scratch in the pool between calls
            int dev = 0;
            cudaMemPool_t pool;
            uint64_t keep = UINT64_MAX;
            if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            pool_ready = true;
        }
        ZAFB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&risk), size_t(total + 32) * sizeof(int), st));
        ZAFB_CUDA(cudaMemsetAsync(risk, 0, sizeof(int), st));
    }
    int* const risk_count = risk;
    int* const risk_list = risk ? risk + 32 : nullptr;
    auto finish = [&](int rc_fp32, float* o, int lay) -> int {  // the float64 pass over the queued frames, then the scratch goes back
        int rc2 = rc_fp32;
        if (risk != nullptr) {
            if (rc2 == ZAFB_OK) rc2 = launch_f64(o, lay, risk_list, risk_count);
            cudaFreeAsync(risk, st);
        }
        return rc2;
    };
    {
        const bool aligned = reinterpret_cast<uintptr_t>(x) % 8 == 0 && (n_clips <= 1 || clip_stride % 2 == 0) && p->hop % 2 == 0;
        const bool ok = p->warp_ok && layout == ZAFB_LAYOUT_FRAME_MAJOR && aligned;
        if (p->force_kernel == 2 && !ok) {
            if (risk) cudaFreeAsync(risk, st);
            return fail(ZAFB_E_UNSUPPORTED, "mel warp kernel needs N = 512, 1024 or 2048, <=128 mels, <=64 coefficients, frame-major layout, even hop/stride");
        }
        if (p->route == ZAFB_MEL_ROUTE_TENSOR) {
            if (!ok || p->d_fb_hi == nullptr)
                return fail(ZAFB_E_UNSUPPORTED, "mel tensor-core route needs N=1024, <=128 mels, frame-major layout, even hop/stride");
            return launch_tensor(p, mode, x, n_clips, ns, clip_stride, nt, out, static_cast<cudaStream_t>(stream));
        }
        if (ok && p->force_kernel != 1) {
            const int wt_total = p->grp_off[3] + 32 * p->grp_len[3];
            const int dh_count = mode == 1 ? p->coef_pad * p->qm4 * 32 : 0;
            const bool big = p->n == 2048;
            const int warps = big ? kWarps2048 : kWarps;
            const int64_t half = p->n / 2;
            const int64_t tile_rows = half / 32 < 16 ? 16 : half / 32;  // the tile also holds the spectrum / log-mel scratch
            const bool win_smem = big || (p->n == 1024 && kMelOcc1024 != 2);
            const size_t smem = size_t(win_smem ? 2 * half : half) * sizeof(float2) + size_t((wt_total + 3) & ~3) * sizeof(float) +
                                size_t(dh_count) * sizeof(float4) + size_t(warps) * tile_rows * kFft1024Pitch * sizeof(float2);
            if (smem <= size_t(kMaxDynSmem) / 2 + 8 * 1024) {  // two CTAs per SM fit (227 KB per SM)
                int64_t ctas = ceil_div(total, warps);
                const int occ = p->n == 1024 ? kMelOcc1024 : 2;
                if (ctas > int64_t(sm_count()) * occ) ctas = int64_t(sm_count()) * occ;
                const int4 gl = make_int4(p->grp_len[0], p->grp_len[1], p->grp_len[2], p->grp_len[3]);
                const int4 go = make_int4(p->grp_off[0], p->grp_off[1], p->grp_off[2], p->grp_off[3]);
                auto kern = big ? (mode == 0 ? mel_warp_kernel<2048, 0> : mel_warp_kernel<2048, 1>)
                            : p->n == 512 ? (mode == 0 ? mel_warp_kernel<512, 0> : mel_warp_kernel<512, 1>)
                                          : (mode == 0 ? mel_warp_kernel<1024, 0> : mel_warp_kernel<1024, 1>);
                kern<<<unsigned(ctas), warps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
                    x, ns, clip_stride, nt, int(p->hop), reinterpret_cast<const float2*>(p->d_window), p->d_tw_4step, p->d_tw_n,
                    p->d_wt, p->d_lo, gl, go, wt_total, reinterpret_cast<const float4*>(p->d_dh), int(p->n_mels), p->half_mels,
                    int(p->n_coef), (p->coef_pad << 8) | p->qm4, out, total, nullptr, risk_count, risk_list, risk_ratio);
                g_launches.fetch_add(1, std::memory_order_relaxed);
                return finish(cudaGetLastError() == cudaSuccess ? ZAFB_OK : fail(ZAFB_E_CUDA, "mel_warp_kernel launch failed"), out, layout);
            }
        }
    }
    const int m = int(p->n / 2);
    const size_t smem = size_t(p->n) * sizeof(float2) + size_t(m + p->n_mels) * sizeof(float) + 16;
    if (smem > size_t(kMaxDynSmem)) {
        if (risk) cudaFreeAsync(risk, st);
        return fail(ZAFB_E_UNSUPPORTED, "mel: window_length %lld too large", (long long)p->n);
    }
    int th = m / 4;
    if (th < 64) th = 64;
    if (th > 256) th = 256;
    const int64_t grid = total < int64_t(sm_count()) * 32 ? total : int64_t(sm_count()) * 32;
    mel_frame_kernel<<<unsigned(grid), th, smem, st>>>(
        x, ns, clip_stride, nt, p->hop, p->log2n, p->d_window, p->d_tw_half, p->d_tw_full, p->d_band_lo, p->d_band_len,
        p->d_band_off, p->d_weights, p->d_dct, int(p->n_mels), int(p->n_coef), mode, out, layout, total, risk_count, risk_list,
        risk_ratio);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return finish(cudaGetLastError() == cudaSuccess ? ZAFB_OK : fail(ZAFB_E_CUDA, "mel_frame_kernel launch failed"), out, layout);
}

}  // namespace

extern "C" {

int zafb_mel_plan_create(zafb_mel_plan** out, const double* window, int64_t n, int64_t hop, const double* fb,
                         int64_t n_mels, int64_t n_coef) {
    ZAFB_REQUIRE(out != nullptr && window != nullptr && fb != nullptr, "plan/window/filterbank is NULL");
    ZAFB_REQUIRE(n >= 2 && hop >= 1 && n_mels >= 1 && n_coef >= 0, "bad mel plan parameters");
    if (n_mels > 4096) return fail(ZAFB_E_UNSUPPORTED, "too many mel filters (%lld)", (long long)n_mels);
    const bool any_length = !is_pow2(n) || n < 4;  // not a power of two: the any-length STFT + mel_from_spectrum_kernel
    if (any_length && size_t(n / 2 + n_mels) * sizeof(float) + 16 > size_t(kMaxDynSmem))
        return fail(ZAFB_E_UNSUPPORTED, "melspectrogram/mfcc: window_length %lld too large", (long long)n);
    zafb_mel_plan* p = new zafb_mel_plan();
    p->n = n;
    p->hop = hop;
    p->n_mels = n_mels;
    p->n_coef = n_coef < n_mels - 1 ? n_coef : n_mels - 1;  // audio_mfcc[1 : n_coef + 1] of n_mels rows (zaf.py:452)
    if (p->n_coef < 0) p->n_coef = 0;
    p->log2n = ilog2(n);
    const int64_t cols = n / 2;
    std::vector<int> lo(n_mels, 0), len(n_mels, 0), off(n_mels, 0);
    std::vector<float> w;
    for (int64_t r = 0; r < n_mels; ++r) {
        int64_t first = -1, last = -1;
        for (int64_t c = 0; c < cols; ++c)
            if (fb[r * cols + c] != 0.0) {
                if (first < 0) first = c;
                last = c;
            }
        off[r] = int(w.size());
        if (first >= 0) {
            lo[r] = int(first);
            len[r] = int(last - first + 1);
            for (int64_t c = first; c <= last; ++c) w.push_back(static_cast<float>(fb[r * cols + c]));
        }
    }
    p->nnz_packed = int64_t(w.size());
    // rows 1..n_coef of the orthonormal DCT-II matrix: sqrt(2/n_mels) cos(pi (2m+1) k / (2 n_mels)), k >= 1
    std::vector<float> d(size_t(p->n_coef) * n_mels);
    const double pi = 3.14159265358979323846264338327950288;
    for (int64_t i = 0; i < p->n_coef; ++i)
        for (int64_t mm = 0; mm < n_mels; ++mm)
            d[i * n_mels + mm] =
                static_cast<float>(std::sqrt(2.0 / double(n_mels)) * std::cos(pi * double(2 * mm + 1) * double(i + 1) / double(2 * n_mels)));
    int rc = upload_f32(&p->d_window, window, n);
    if (any_length) {
        if (rc == ZAFB_OK) rc = zafb_stft_plan_create(&p->stft_any, window, n, hop);
        if (rc == ZAFB_OK) rc = upload_vec(&p->d_band_lo, lo);
        if (rc == ZAFB_OK) rc = upload_vec(&p->d_band_len, len);
        if (rc == ZAFB_OK) rc = upload_vec(&p->d_band_off, off);
        if (rc == ZAFB_OK) rc = upload_vec(&p->d_weights, w);
        if (rc == ZAFB_OK) rc = upload_vec(&p->d_dct, d);
        if (rc != ZAFB_OK) {
            zafb_mel_plan_destroy(p);
            return rc;
        }
        p->precision = 32;  // this route is fp32 throughout
        *out = p;
        return ZAFB_OK;
    }
    if (rc == ZAFB_OK) rc = upload_twiddles(&p->d_tw_half, n / 2, n / 2);
    if (rc == ZAFB_OK) rc = upload_twiddles(&p->d_tw_full, n, n / 2);
    if (rc == ZAFB_OK) rc = upload_vec(&p->d_band_lo, lo);
    if (rc == ZAFB_OK) rc = upload_vec(&p->d_band_len, len);
    if (rc == ZAFB_OK) rc = upload_vec(&p->d_band_off, off);
    if (rc == ZAFB_OK) rc = upload_vec(&p->d_weights, w);
    if (rc == ZAFB_OK) rc = upload_vec(&p->d_dct, d);
    if (rc == ZAFB_OK) {  // the float64 tables (mfcc recomputes at-risk frames in double; precision = 64 runs everything there)
        const int64_t h = n / 2;
        std::vector<double2> th(h), tf(h);
        for (int64_t t = 0; t < h; ++t) {
            th[t] = make_double2(std::cos(-2.0 * pi * double(t) / double(h)), std::sin(-2.0 * pi * double(t) / double(h)));
            tf[t] = make_double2(std::cos(-2.0 * pi * double(t) / double(n)), std::sin(-2.0 * pi * double(t) / double(n)));
        }
        std::vector<double> w64(window, window + n);
        rc = upload_vec(&p->d_window64, w64);
        if (rc == ZAFB_OK) rc = upload_vec(&p->d_tw_half64, th);
        if (rc == ZAFB_OK) rc = upload_vec(&p->d_tw_full64, tf);
        if (rc == ZAFB_OK && n == 1024) {
            std::vector<double2> t4(h);
            for (int64_t k1 = 0; k1 < h / 32; ++k1)
                for (int64_t n2 = 0; n2 < 32; ++n2) {
                    const double a = -2.0 * pi * double((k1 * n2) % h) / double(h);
                    t4[k1 * 32 + n2] = make_double2(std::cos(a), std::sin(a));
                }
            rc = upload_vec(&p->d_tw_4step64, t4);
        }
    }
    if (rc == ZAFB_OK && (n == 512 || n == 1024 || n == 2048) && n_mels <= 128 && p->n_coef <= 64) {
        // row groups of 32 rows, zero-padded to the longest band of the group; the band start is clamped so that
        // lo + grp_len never leaves the 512-column spectrum (the padding weights are zero)
        std::vector<int> lo4(128, 0);
        std::vector<float> wt;
        for (int g = 0; g < 4; ++g) {
            int longest = 0;
            for (int l = 0; l < 32; ++l) {
                const int64_t r = l + 32 * g;
                if (r < n_mels && len[r] > longest) longest = len[r];
            }
            p->grp_len[g] = longest;
            p->grp_off[g] = int(wt.size());
            wt.resize(wt.size() + size_t(32) * longest, 0.f);
            for (int l = 0; l < 32; ++l) {
                const int64_t r = l + 32 * g;
                if (r >= n_mels || len[r] == 0) continue;
                int start = lo[r];
                if (start + longest > int(cols)) start = int(cols) - longest;  // shift left, pad in front
                lo4[g * 32 + l] = start;
                for (int c = 0; c < len[r]; ++c) wt[p->grp_off[g] + 32 * (lo[r] - start + c) + l] = w[off[r] + c];
            }
        }
        p->half_mels = int((n_mels + 1) / 2);
        p->coef_pad = int((p->n_coef + 7) / 8);
        p->qm4 = (p->half_mels + 15) / 16;
        std::vector<float> dh(size_t(p->coef_pad ? p->coef_pad : 1) * (p->qm4 ? p->qm4 : 1) * 32 * 4, 0.f);
        for (int t = 0; t < p->coef_pad; ++t)
            for (int mq = 0; mq < p->qm4; ++mq)
                for (int l = 0; l < 32; ++l) {
                    const int64_t i = (l & 7) + 8 * t;  // coefficient index, row k = i + 1
                    for (int e = 0; e < 4; ++e) {
                        const int64_t mm = int64_t(l >> 3) * 4 * p->qm4 + 4 * mq + e;
                        if (i < p->n_coef && mm < p->half_mels)
                            dh[(size_t(t * p->qm4 + mq) * 32 + l) * 4 + e] = static_cast<float>(
                                std::sqrt(2.0 / double(n_mels)) * std::cos(pi * double(2 * mm + 1) * double(i + 1) / double(2 * n_mels)));
                    }
                }
        const int64_t half = n / 2;  // four-step twiddles W_half^{k1 n2}, [k1][n2], half = (half / 32) x 32
        std::vector<double> t4(2 * half);
        for (int64_t k1 = 0; k1 < half / 32; ++k1)
            for (int64_t n2 = 0; n2 < 32; ++n2) {
                const double a = -2.0 * pi * double((k1 * n2) % half) / double(half);
                t4[2 * (k1 * 32 + n2)] = std::cos(a);
                t4[2 * (k1 * 32 + n2) + 1] = std::sin(a);
            }
        rc = upload_c32(&p->d_tw_4step, t4.data(), half);
        if (rc == ZAFB_OK) rc = upload_twiddles(&p->d_tw_n, n, 32);
        if (rc == ZAFB_OK) rc = upload_vec(&p->d_wt, wt);
        if (rc == ZAFB_OK) rc = upload_vec(&p->d_lo, lo4);
        if (rc == ZAFB_OK) rc = upload_vec(&p->d_dh, dh);
        p->warp_ok = rc == ZAFB_OK;
        if (rc == ZAFB_OK && n == 1024) {  // dense TF32 hi/lo operands of the tensor-core route
            p->ld_mel = (n_mels + 3) & ~int64_t(3);
            std::vector<float> hi(size_t(n_mels) * 512), lo(hi.size());
            split_tf32_host(fb, hi.size(), hi.data(), lo.data());
            rc = upload_vec(&p->d_fb_hi, hi);
            if (rc == ZAFB_OK) rc = upload_vec(&p->d_fb_lo, lo);
            if (rc == ZAFB_OK && p->n_coef > 0) {
                std::vector<double> d(size_t(p->n_coef) * p->ld_mel, 0.0);
                for (int64_t i = 0; i < p->n_coef; ++i)
                    for (int64_t mm = 0; mm < n_mels; ++mm)
                        d[i * p->ld_mel + mm] = std::sqrt(2.0 / double(n_mels)) *
                                                std::cos(pi * double(2 * mm + 1) * double(i + 1) / double(2 * n_mels));
                std::vector<float> dhi(d.size()), dlo(d.size());
                split_tf32_host(d.data(), d.size(), dhi.data(), dlo.data());
                rc = upload_vec(&p->d_dct_hi, dhi);
                if (rc == ZAFB_OK) rc = upload_vec(&p->d_dct_lo, dlo);
            }
        }
    }
    if (rc != ZAFB_OK) {
        zafb_mel_plan_destroy(p);
        return rc;
    }
    *out = p;
    return ZAFB_OK;
}

// test hook: 0 = auto, 1 = generic kernel only, 2 = require the warp kernel
int zafb_mel_plan_force_kernel(zafb_mel_plan* p, int which) {
    ZAFB_REQUIRE(p != nullptr && which >= 0 && which <= 2, "bad plan / kernel id");
    p->force_kernel = which;
    return ZAFB_OK;
}

int zafb_mel_plan_set_precision(zafb_mel_plan* p, int bits, const double* window) {
    ZAFB_REQUIRE(p != nullptr, "plan is NULL");
    ZAFB_REQUIRE(bits == 0 || bits == 32 || bits == 64, "precision must be 0 (automatic), 32 or 64 (got %d)", bits);
    (void)window;  // the float64 tables are built from the window handed to zafb_mel_plan_create
    if (p->stft_any != nullptr) {  // window length not a power of two: fp32 only
        if (bits == 64) return fail(ZAFB_E_UNSUPPORTED, "the float64 route needs a power-of-two window length");
        return ZAFB_OK;
    }
    p->precision = bits;
    return ZAFB_OK;
}

int zafb_mel_plan_set_route(zafb_mel_plan* p, int route) {
    ZAFB_REQUIRE(p != nullptr, "plan is NULL");
    ZAFB_REQUIRE(route == ZAFB_MEL_ROUTE_FUSED || route == ZAFB_MEL_ROUTE_TENSOR, "bad route %d", route);
    if (route == ZAFB_MEL_ROUTE_TENSOR && p->d_fb_hi == nullptr)
        return fail(ZAFB_E_UNSUPPORTED, "mel tensor-core route needs window_length 1024 and <= 128 mel rows");
    p->route = route;
    return ZAFB_OK;
}

int zafb_mel_plan_destroy(zafb_mel_plan* p) {
    if (!p) return ZAFB_OK;
    zafb_stft_plan_destroy(p->stft_any);
    cudaFree(p->d_window64);
    cudaFree(p->d_tw_half64);
    cudaFree(p->d_tw_full64);
    cudaFree(p->d_tw_4step64);
    cudaFree(p->d_fb_hi);
    cudaFree(p->d_fb_lo);
    cudaFree(p->d_dct_hi);
    cudaFree(p->d_dct_lo);
    cudaFree(p->d_window);
    cudaFree(p->d_tw_half);
    cudaFree(p->d_tw_full);
    cudaFree(p->d_band_lo);
    cudaFree(p->d_band_len);
    cudaFree(p->d_band_off);
    cudaFree(p->d_weights);
    cudaFree(p->d_dct);
    cudaFree(p->d_tw_4step);
    cudaFree(p->d_tw_n);
    cudaFree(p->d_wt);
    cudaFree(p->d_lo);
    cudaFree(p->d_dh);
    delete p;
    return ZAFB_OK;
}

int zafb_melspectrogram_f32(const zafb_mel_plan* p, const float* x, int64_t n_clips, int64_t ns, int64_t clip_stride,
                            float* out, int layout, void* stream) {
    return launch(p, 0, x, n_clips, ns, clip_stride, out, layout, stream);
}

int zafb_mfcc_f32(const zafb_mel_plan* p, const float* x, int64_t n_clips, int64_t ns, int64_t clip_stride, float* out,
                  int layout, void* stream) {
    return launch(p, 1, x, n_clips, ns, clip_stride, out, layout, stream);
}

static int mel_host(const zafb_mel_plan* p, int mode, const float* x, int64_t n_clips, int64_t ns, int64_t clip_stride,
                    float* out, int layout) {
    ZAFB_REQUIRE(p != nullptr, "plan is NULL");
    ZAFB_REQUIRE(n_clips >= 0 && ns >= 0 && clip_stride >= ns, "bad batch geometry");
    int64_t nt = 0;
    zafb_stft_geometry(ns, p->n, p->hop, nullptr, &nt, nullptr);
    if (n_clips == 0) return ZAFB_OK;
    ZAFB_REQUIRE(out != nullptr && (x != nullptr || ns == 0), "x/out is NULL");
    const int64_t rows = mode == 0 ? p->n_mels : p->n_coef;
    const size_t out_clip = size_t(nt) * size_t(rows) * sizeof(float);
    const int64_t dpitch = (ns + 1) & ~int64_t(1);
    return run_host_pipeline(x, size_t(clip_stride) * sizeof(float), size_t(ns) * sizeof(float), size_t(dpitch) * sizeof(float),
                             out, out_clip, out_clip, out_clip, n_clips,
                             [&](void* d_in, void* d_out, int64_t, int64_t nc, cudaStream_t st) {
                                 return launch(p, mode, static_cast<const float*>(d_in), nc, ns, dpitch,
                                               static_cast<float*>(d_out), layout, st);
                             });
}

int zafb_melspectrogram_host_f32(const zafb_mel_plan* p, const float* x, int64_t n_clips, int64_t ns, int64_t clip_stride,
                                 float* out, int layout) {
    return mel_host(p, 0, x, n_clips, ns, clip_stride, out, layout);
}

int zafb_mfcc_host_f32(const zafb_mel_plan* p, const float* x, int64_t n_clips, int64_t ns, int64_t clip_stride, float* out,
                       int layout) {
    return mel_host(p, 1, x, n_clips, ns, clip_stride, out, layout);
}

}  // extern "C"
