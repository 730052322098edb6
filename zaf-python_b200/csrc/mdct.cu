// MDCT / IMDCT kernels (zaf.py:984-1075, 1078-1184).
//
// Both directions are a DCT-IV of size M = N/2 computed through an M/2-point complex FFT:
//     t[m] = (v[2m] + i v[M-1-2m]) e^{-i pi m / M},   y = FFT_{M/2}(t) . e^{-i pi (m + 1/4) / M},
//     DCT4(v)[2m] = Re y[m],   DCT4(v)[M-1-2m] = -Im y[m]
// MDCT :  X[:, j] = DCT4(fold(w . frame_j)),  fold(u) = [ -c_r - d , a - b_r ]   (u = [a b c d], _r = reversed)
// IMDCT:  (p, q) = halves of DCT4(X[:, j]);  frame_j = (2/M) w . [ q, -q_r, -p_r, -p ];  y = TDAC overlap-add.
// (Identities checked against the reference's pre/post-twiddle formulation in tests/ and SURVEY.md 8a.)
//
//   mdct_generic_kernel   one CTA per frame, Stockham FFT in shared memory, power-of-two M >= 2.
//   mdct_direct_kernel    any even N: direct O(M N) cosine sum (table lookup), the reference accepts any even N.
//   imdct_tile_kernel     one CTA per run of consecutive hop-blocks; every output sample is the sum of exactly
//                         two frames (h-1 and h), computed in that order and written once -- no atomics.
#include <climits>
#include <cmath>
#include <cstdint>
#include <vector>

#include "fft_core.cuh"
#include "host_pipe.cuh"
#include "transpose.cuh"
#include "mdct_common.cuh"

using namespace zafb;

namespace zafb {  // mdct_binmajor.cu: the kernels that write / read the reference's C-order memory directly
bool mdct_binmajor_supported(const zafb_mdct_plan* p, int64_t nt);
int mdct_binmajor_launch(const zafb_mdct_plan* p, const float* x, int64_t n_clips, int64_t ns, int64_t clip_stride, int64_t nt,
                         float* out, cudaStream_t st);
int imdct_binmajor_launch(const zafb_mdct_plan* p, const float* spec, int64_t n_clips, int64_t nt, int64_t out_len, float* y,
                          int64_t y_stride, cudaStream_t st);
}  // namespace zafb


namespace {

constexpr int kMaxDynSmem = 200 * 1024;

// DCT-IV of the M values in v (shared) -> out (shared, M floats).  a/b: M/2 float2 scratch each.
__device__ __forceinline__ void dct4_block(const float* v, float* out, float2* a, float2* b,
                                           const float2* __restrict__ tw_fft, const float2* __restrict__ pre,
                                           const float2* __restrict__ post, int log2m, int tid, int nth) {
    const int m_len = 1 << log2m;
    const int h = m_len >> 1;
    for (int m = tid; m < h; m += nth) a[m] = cmul(make_float2(v[2 * m], v[m_len - 1 - 2 * m]), pre[m]);
    __syncthreads();
    const float2* y = block_fft(a, b, tw_fft, log2m - 1, tid, nth);
    for (int m = tid; m < h; m += nth) {
        const float2 r = cmul(y[m], post[m]);
        out[2 * m] = r.x;
        out[m_len - 1 - 2 * m] = -r.y;
    }
    __syncthreads();
}

__global__ void mdct_generic_kernel(const float* __restrict__ x, int64_t ns, int64_t clip_stride, int64_t nt,
                                    int log2m, const float* __restrict__ window, const float2* __restrict__ tw_fft,
                                    const float2* __restrict__ pre, const float2* __restrict__ post,
                                    float* __restrict__ out, int layout, int64_t total_frames) {
    extern __shared__ float2 smem2[];
    const int m_len = 1 << log2m, h = m_len >> 1;
    float2* a = smem2;
    float2* b = smem2 + h;
    float* v = reinterpret_cast<float*>(smem2 + 2 * h);
    float* res = v + m_len;
    const int tid = threadIdx.x, nth = blockDim.x;
    for (int64_t f = blockIdx.x; f < total_frames; f += gridDim.x) {
        const int64_t clip = f / nt, j = f - clip * nt;
        const int64_t start = (j - 1) * m_len;  // frame j covers original samples [(j-1)M, (j+1)M)
        const float* xc = x + clip * clip_stride;
        auto u = [&](int i) -> float {           // windowed sample i of the 2M-long frame
            const int64_t s = start + i;
            return (s >= 0 && s < ns) ? xc[s] * window[i] : 0.f;
        };
        for (int i = tid; i < h; i += nth) {
            v[i] = -u(3 * h - 1 - i) - u(3 * h + i);      // -c_r - d
            v[h + i] = u(i) - u(m_len - 1 - i);           //  a - b_r
        }
        __syncthreads();
        dct4_block(v, res, a, b, tw_fft, pre, post, log2m, tid, nth);
        for (int k = tid; k < m_len; k += nth) {
            if (layout == ZAFB_LAYOUT_FRAME_MAJOR) out[f * m_len + k] = res[k];
            else out[(clip * m_len + k) * nt + j] = res[k];
        }
        __syncthreads();
    }
}

// any even N: X[k] = sum_n u[n] cos(pi/M (n + 1/2 + M/2)(k + 1/2)) = cos(2 pi (2n+1+M)(2k+1) / (8M))
__global__ void mdct_direct_kernel(const float* __restrict__ x, int64_t ns, int64_t clip_stride, int64_t nt, int m_len,
                                   const float* __restrict__ window, const float* __restrict__ costab,
                                   float* __restrict__ out, int layout, int64_t total_frames) {
    extern __shared__ float2 smem2[];
    float* u = reinterpret_cast<float*>(smem2);
    const int tid = threadIdx.x, nth = blockDim.x;
    const int n = 2 * m_len, period = 8 * m_len;
    for (int64_t f = blockIdx.x; f < total_frames; f += gridDim.x) {
        const int64_t clip = f / nt, j = f - clip * nt;
        const int64_t start = (j - 1) * m_len;
        const float* xc = x + clip * clip_stride;
        for (int i = tid; i < n; i += nth) {
            const int64_t s = start + i;
            u[i] = (s >= 0 && s < ns) ? xc[s] * window[i] : 0.f;
        }
        __syncthreads();
        for (int k = tid; k < m_len; k += nth) {
            const int kk = 2 * k + 1;
            int idx = int((int64_t(1 + m_len) * kk) % period);
            const int step = (2 * kk) % period;
            float acc = 0.f;
            for (int i = 0; i < n; ++i) {
                acc = fmaf(u[i], costab[idx], acc);
                idx += step;
                if (idx >= period) idx -= period;
            }
            if (layout == ZAFB_LAYOUT_FRAME_MAJOR) out[f * m_len + k] = acc;
            else out[(clip * m_len + k) * nt + j] = acc;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// N = 2048 (M = 1024): one warp per frame, everything in registers except one transpose.
//
// With the windowed frame u viewed as 1024 pairs P[p] = (u[2p], u[2p+1]) the DCT-IV input
// t[m] = (v[2m] + i v[M-1-2m]) pre[m] of the folded block v = [-c_r - d, a - b_r] needs, for m < 256,
//     P[768+m], P[767-m], P[255-m], P[256+m]
// and t[511-m] needs exactly the same four pairs.  A lane therefore loads 4 x 8 coalesced pairs,
// forms t[m] for its own m = lane + 32 r (r < 8) and hands the raw value of t[511-m] to lane
// 31 - lane (register 15 - r) with one xor-31 shuffle.  After the 512-point FFT the outputs
// X[2k] = Re Y[k], X[1023-2k] = -Im Y[k] are regrouped into pairs (X[2P], X[2P+1]) with the same
// shuffle, so loads and stores are all full 256-byte warp transactions.
// (Index maps validated in float64 against the closed form of zaf.py:1047-1073.)
// ------------------------------------------------------------------------------------------
constexpr int kWarps = 8;
#ifndef ZAFB_IMDCT_SMALL_OCC
#define ZAFB_IMDCT_SMALL_OCC 3  // CTAs per SM of imdct_warp_kernel<512> (2 -> 5.39 ms, 3 -> 5.21; N = 1024 spills with 3: 4.54 -> 4.88)
#endif

template <int N>
__device__ __forceinline__ void load_tables(float2* smem, const float2* __restrict__ win_pairs,
                                            const float2* __restrict__ tw4, int tid, float win_scale = 1.0f) {
    using G = MdctGeom<N>;
    for (int i = tid; i < G::M; i += kWarps * 32) {
        const float2 w = win_pairs[i];
        smem[i] = make_float2(w.x * win_scale, w.y * win_scale);
    }
    for (int i = tid; i < G::H; i += kWarps * 32) smem[G::M + i] = tw4[i];
}

template <int N, int OCC>
__global__ void __launch_bounds__(kWarps * 32, OCC)
mdct_warp_kernel(const float* __restrict__ x, int64_t ns, int64_t clip_stride, int64_t nt,
                 const float2* __restrict__ win_pairs, const float2* __restrict__ tw4,
                 const float2* __restrict__ pre, const float2* __restrict__ post, float* __restrict__ out,
                 int64_t total_frames) {
    using G = MdctGeom<N>;
    constexpr int M = G::M, REGS = G::REGS, LOGR = G::LOGR, Q = G::Q, QR = REGS / 2;
    extern __shared__ float2 smem2[];
    const float2* s_win = smem2;          // M pairs of the window
    const float2* s_tw = smem2 + M;       // H: W_H^{k1 n2}
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2* s_buf = smem2 + G::TABLES + warp * G::TILE;
    load_tables<N>(smem2, win_pairs, tw4, tid);
    const float2 c_lane = pre[lane];   // e^{-i pi lane / M}; pre[lane + 32 r] = c_lane W_TWDEN^r
    const float2 p_lane = post[lane];  // e^{-i pi (lane + 1/4) / M}; post[lane + 32 k] = p_lane W_TWDEN^k
    float2 tq[G::NTQ];
    if constexpr (N == 1024) warp_fft256_lane_twiddles(tq, lane);
    if constexpr (N == 512) warp_fft128_lane_twiddles(tq, lane);
    __syncthreads();

    for (int64_t f = int64_t(blockIdx.x) * kWarps + warp; f < total_frames; f += int64_t(gridDim.x) * kWarps) {
        const int64_t clip = f / nt, j = f - clip * nt;
        const int64_t start = (j - 1) * M;  // frame j covers original samples [(j-1)M, (j+1)M)
        const float* xc = x + clip * clip_stride;

        // pairs 3Q+m, 3Q-1-m, Q-1-m, Q+m for m = lane + 32 r, r < QR
        float2 pr[QR][4];
        if (start >= 0 && start + N <= ns) {
            const float2* fp = reinterpret_cast<const float2*>(xc + start);
#pragma unroll
            for (int r = 0; r < QR; ++r) {
                const int m = lane + 32 * r;
                pr[r][0] = __ldg(fp + 3 * Q + m);
                pr[r][1] = __ldg(fp + 3 * Q - 1 - m);
                pr[r][2] = __ldg(fp + Q - 1 - m);
                pr[r][3] = __ldg(fp + Q + m);
            }
        } else {
            auto ld = [&](int p) {
                const int64_t s0 = start + 2 * p;
                return make_float2((s0 >= 0 && s0 < ns) ? __ldg(xc + s0) : 0.f,
                                   (s0 + 1 >= 0 && s0 + 1 < ns) ? __ldg(xc + s0 + 1) : 0.f);
            };
#pragma unroll
            for (int r = 0; r < QR; ++r) {
                const int m = lane + 32 * r;
                pr[r][0] = ld(3 * Q + m);
                pr[r][1] = ld(3 * Q - 1 - m);
                pr[r][2] = ld(Q - 1 - m);
                pr[r][3] = ld(Q + m);
            }
        }
        float2 v[REGS];
        static_for<0, QR>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            const int m = lane + 32 * r;
            const float2 w1 = s_win[3 * Q + m], w2 = s_win[3 * Q - 1 - m], w3 = s_win[Q - 1 - m], w4 = s_win[Q + m];
            const float2 p1 = make_float2(pr[r][0].x * w1.x, pr[r][0].y * w1.y);
            const float2 p2 = make_float2(pr[r][1].x * w2.x, pr[r][1].y * w2.y);
            const float2 p3 = make_float2(pr[r][2].x * w3.x, pr[r][2].y * w3.y);
            const float2 p4 = make_float2(pr[r][3].x * w4.x, pr[r][3].y * w4.y);
            v[r] = make_float2(-p2.y - p1.x, p3.y - p4.x);               // raw t[m]
            const float2 other = make_float2(p3.x - p4.y, -p2.x - p1.y);  // raw t[H - 1 - m], belongs to lane 31 - lane
            v[REGS - 1 - r].x = __shfl_xor_sync(0xffffffffu, other.x, 31);
            v[REGS - 1 - r].y = __shfl_xor_sync(0xffffffffu, other.y, 31);
        });
        static_for<0, REGS>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            v[r] = cmul(v[r], mul_tw<r, G::TWDEN>(c_lane));
        });

        mdct_warp_fft<N>(v, s_tw, s_buf, lane, tq);  // Y[lane + 32 k] = v[bitrev(k)]

        static_for<0, REGS>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            v[bitrev(k, LOGR)] = cmul(v[bitrev(k, LOGR)], mul_tw<k, G::TWDEN>(p_lane));
        });
        float2* o = reinterpret_cast<float2*>(out + f * M) + lane;
        static_for<0, REGS>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            const float im = __shfl_xor_sync(0xffffffffu, v[bitrev(REGS - 1 - k, LOGR)].y, 31);
            __stcs(o + 32 * k, make_float2(v[bitrev(k, LOGR)].x, -im));
        });
    }
}

// IMDCT, N = 2048 or 1024: one warp per run of consecutive hop-blocks of one clip; the second half of the
// previous frame (already scaled and windowed) is carried in registers, so every output sample is
// carry (frame h-1) + first half (frame h), written once.  A run re-reads the one frame before it.
template <int N, int OCC>
__global__ void __launch_bounds__(kWarps * 32, OCC)
imdct_warp_kernel(const float* __restrict__ spec, int64_t nt, const float2* __restrict__ win_pairs,
                  const float2* __restrict__ tw4, const float2* __restrict__ pre,
                  const float2* __restrict__ post, int64_t runs_per_clip, int run_len, int64_t total_runs,
                  int64_t out_len, float* __restrict__ y, int64_t y_stride, int y_aligned, int prefetch) {
    using G = MdctGeom<N>;
    constexpr int M = G::M, H = G::H, REGS = G::REGS, LOGR = G::LOGR, HR = REGS / 2;
    extern __shared__ float2 smem2[];
    const float2* s_win = smem2;
    const float2* s_tw = smem2 + M;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2* s_buf = smem2 + G::TABLES + warp * G::TILE;
    constexpr float kScale = 2.0f / float(M);  // a power of two: folding it into the window table changes no bit
    load_tables<N>(smem2, win_pairs, tw4, tid, kScale);

    const float2 c_lane = pre[lane];
    const float2 p_lane = post[lane];
    float2 tq[G::NTQ];
    if constexpr (N == 1024) warp_fft256_lane_twiddles(tq, lane);
    if constexpr (N == 512) warp_fft128_lane_twiddles(tq, lane);
    __syncthreads();

    for (int64_t task = int64_t(blockIdx.x) * kWarps + warp; task < total_runs; task += int64_t(gridDim.x) * kWarps) {
        const int64_t clip = task / runs_per_clip;
        const int64_t run = task - clip * runs_per_clip;
        const int64_t h0 = 1 + run * run_len;
        int64_t h1 = h0 + run_len;
        if (h1 > nt) h1 = nt;
        float* yc = y + clip * y_stride;
        float2 carry[REGS];
#pragma unroll
        for (int i = 0; i < REGS; ++i) carry[i] = make_float2(0.f, 0.f);

        for (int64_t j = h0 - 1; j < h1; ++j) {
            const float2* X = reinterpret_cast<const float2*>(spec + (clip * nt + j) * M) + lane;
            // next frame of the run towards L2: M * 4 / 128 lines (one per lane at M = 1024)
            if (prefetch && j + 1 < h1) {
#pragma unroll
                for (int i = lane; i < M / 32; i += 32) prefetch_l2(reinterpret_cast<const char*>(X - lane + H) + i * 128);
            }
            float2 xp[REGS], v[REGS];
#pragma unroll
            for (int r = 0; r < REGS; ++r) xp[r] = __ldg(X + 32 * r);
            static_for<0, REGS>([&](auto rc) {
                constexpr int r = decltype(rc)::value;
                const float im = __shfl_xor_sync(0xffffffffu, xp[REGS - 1 - r].y, 31);  // X[M - 1 - 2m]
                v[r] = cmul(make_float2(xp[r].x, im), mul_tw<r, G::TWDEN>(c_lane));
            });

            mdct_warp_fft<N>(v, s_tw, s_buf, lane, tq);

            // A[k] = res[2k] = Re(Y post), B[k] = res[M - 1 - 2k] = -Im(Y post), k = lane + 32 kap
            static_for<0, REGS>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                const float2 t = cmul(v[bitrev(k, LOGR)], mul_tw<k, G::TWDEN>(p_lane));
                v[bitrev(k, LOGR)] = make_float2(t.x, -t.y);
            });
            const bool emit = j >= h0;
            const int64_t base = (j - 1) * M;  // output index of OLA sample j*M (trim = M)
            const bool whole = y_aligned && base + M <= out_len;
            static_for<0, REGS>([&](auto rc) {
                constexpr int rho = decltype(rc)::value;
                constexpr int own = rho < HR ? rho + HR : rho - HR;                  // register (kap) of this lane's value
                constexpr int oth = rho < HR ? HR - 1 - rho : 3 * HR - 1 - rho;      // register of lane 31 - lane's value
                const float2 mine = v[bitrev(own, LOGR)];
                float2 part;
                part.x = __shfl_xor_sync(0xffffffffu, v[bitrev(oth, LOGR)].x, 31);
                part.y = __shfl_xor_sync(0xffffffffu, v[bitrev(oth, LOGR)].y, 31);
                float2 first, second;
                if constexpr (rho < HR) {
                    first = make_float2(mine.x, part.y);     // ( A[P + M/4],  B[M/4 - 1 - P])
                    second = make_float2(-mine.y, -part.x);  // (-B[P + M/4], -A[M/4 - 1 - P])
                } else {
                    first = make_float2(-mine.y, -part.x);   // (-B[P - M/4], -A[3M/4 - 1 - P])
                    second = make_float2(-mine.x, -part.y);  // (-A[P - M/4], -B[3M/4 - 1 - P])
                }
                const int P = lane + 32 * rho;
                const float2 w1 = s_win[P], w2 = s_win[H + P];
                if (emit) {
                    const float2 o = make_float2(fmaf(w1.x, first.x, carry[rho].x), fmaf(w1.y, first.y, carry[rho].y));
                    if (whole) {  // the hop-block lies inside the output and rows are 8-byte aligned: no per-store checks
                        __stcs(reinterpret_cast<float2*>(yc + base) + P, o);
                    } else {
                        const int64_t idx = base + 2 * P;
                        if (y_aligned && idx + 1 < out_len) {
                            __stcs(reinterpret_cast<float2*>(yc + idx), o);
                        } else {
                            if (idx < out_len) yc[idx] = o.x;
                            if (idx + 1 < out_len) yc[idx + 1] = o.y;
                        }
                    }
                }
                carry[rho] = make_float2(w2.x * second.x, w2.y * second.y);
            });
        }
    }
}

// IMDCT.  Output sample p (OLA coordinates, hop-block h = p / M) = second half of frame h-1 + first half of frame h.
// The reference keeps OLA[M : M*nt - 1]: hop-blocks 1 .. nt-1, last sample dropped (zaf.py:1182).
// One CTA owns hop-blocks [h0, h1) of one clip and walks frames h0-1 .. h1-1, carrying the second half.
__global__ void imdct_tile_kernel(const float* __restrict__ spec, int64_t nt, int m_len, int log2m, int layout,
                                  const float* __restrict__ window, const float2* __restrict__ tw_fft,
                                  const float2* __restrict__ pre, const float2* __restrict__ post,
                                  const float* __restrict__ costab, int64_t blocks_per_tile, int64_t tiles_per_clip,
                                  int64_t out_len, float* __restrict__ y, int64_t y_stride) {
    extern __shared__ float2 smem2[];
    const int h = m_len >> 1;
    float2* a = smem2;
    float2* b = smem2 + (h > 0 ? h : 1);
    float* v = reinterpret_cast<float*>(smem2 + 2 * (h > 0 ? h : 1));
    float* res = v + m_len;           // DCT-IV result (p | q), or the whole 2M block on the direct path
    float* carry = res + 2 * m_len;   // second half of the previous frame, already scaled and windowed
    const int tid = threadIdx.x, nth = blockDim.x;
    const int64_t clip = blockIdx.x / tiles_per_clip;
    const int64_t tile = blockIdx.x - clip * tiles_per_clip;
    const int64_t h0 = 1 + tile * blocks_per_tile;
    int64_t h1 = h0 + blocks_per_tile;
    if (h1 > nt) h1 = nt;
    const float scale = 2.0f / float(m_len);
    float* yc = y + clip * y_stride;
    for (int64_t j = h0 - 1; j < h1; ++j) {
        for (int k = tid; k < m_len; k += nth)
            v[k] = (layout == ZAFB_LAYOUT_FRAME_MAJOR) ? spec[(clip * nt + j) * m_len + k] : spec[(clip * m_len + k) * nt + j];
        __syncthreads();
        if (log2m >= 1) {
            dct4_block(v, res, a, b, tw_fft, pre, post, log2m, tid, nth);
        } else {  // any M: the 2M-long block directly, cos(pi/M (n+1/2+M/2)(k+1/2)) = costab[((2n+1+M)(2k+1)) mod 8M]
            const int period = 8 * m_len;
            for (int i = tid; i < 2 * m_len; i += nth) {
                const int nn = 2 * i + 1 + m_len;
                int idx = nn % period;
                const int step = (2 * nn) % period;
                float acc = 0.f;
                for (int k = 0; k < m_len; ++k) {
                    acc = fmaf(v[k], costab[idx], acc);
                    idx += step;
                    if (idx >= period) idx -= period;
                }
                res[i] = acc;
            }
            __syncthreads();
        }
        // power-of-two path: block = [ q, -q_r, -p_r, -p ] with p = res[0:h], q = res[h:M]:
        //   first half  (n < M):        n < h ? q[n] : -q[M-1-(n-h)]
        //   second half (n' = n-M < M): n' < h ? -p[h-1-n'] : -p[n'-h]
        const bool folded = log2m >= 1;
        if (j >= h0) {
            const int64_t base = j * m_len - m_len;  // output index of OLA sample j*M (trim = M)
            for (int i = tid; i < m_len; i += nth) {
                const float first = folded ? ((i < h) ? res[h + i] : -res[m_len - 1 - (i - h)]) : res[i];
                const int64_t o = base + i;
                if (o < out_len) yc[o] = carry[i] + scale * window[i] * first;
            }
        }
        __syncthreads();
        for (int i = tid; i < m_len; i += nth) {
            const float second = folded ? ((i < h) ? -res[h - 1 - i] : -res[i - h]) : res[m_len + i];
            carry[i] = scale * window[m_len + i] * second;
        }
        __syncthreads();
    }
}

bool g_attr_done = false;
int set_kernel_attrs() {
    if (g_attr_done) return ZAFB_OK;
    ZAFB_CUDA((cudaFuncSetAttribute(mdct_warp_kernel<2048, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(imdct_warp_kernel<2048, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(mdct_warp_kernel<4096, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(imdct_warp_kernel<4096, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(mdct_warp_kernel<1024, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(imdct_warp_kernel<1024, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(imdct_warp_kernel<512, ZAFB_IMDCT_SMALL_OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA(cudaFuncSetAttribute(mdct_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    ZAFB_CUDA(cudaFuncSetAttribute(mdct_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    ZAFB_CUDA(cudaFuncSetAttribute(imdct_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    g_attr_done = true;
    return ZAFB_OK;
}

int threads_for(int points) {
    int t = points / 4;
    if (t < 32) t = 32;
    if (t > 256) t = 256;
    return t;
}

}  // namespace

extern "C" {

int zafb_mdct_plan_create(zafb_mdct_plan** out, const double* window, int64_t n) {
    ZAFB_REQUIRE(out != nullptr && window != nullptr, "plan/window is NULL");
    ZAFB_REQUIRE(n >= 2 && n % 2 == 0, "mdct: window_length must be even and >= 2 (got %lld)", (long long)n);
    if (n > (1 << 16)) return fail(ZAFB_E_UNSUPPORTED, "mdct: window_length %lld too large", (long long)n);
    zafb_mdct_plan* p = new zafb_mdct_plan();
    p->n = n;
    p->m = n / 2;
    p->log2m = (is_pow2(p->m) && p->m >= 2) ? ilog2(p->m) : -1;
    const double pi = 3.14159265358979323846264338327950288;
    int rc = upload_f32(&p->d_window, window, n);
    if (rc == ZAFB_OK && p->log2m >= 1) {
        const int64_t h = p->m / 2;
        std::vector<double> pre(2 * h), post(2 * h);
        for (int64_t m = 0; m < h; ++m) {
            pre[2 * m] = std::cos(-pi * double(m) / double(p->m));
            pre[2 * m + 1] = std::sin(-pi * double(m) / double(p->m));
            post[2 * m] = std::cos(-pi * (double(m) + 0.25) / double(p->m));
            post[2 * m + 1] = std::sin(-pi * (double(m) + 0.25) / double(p->m));
        }
        rc = upload_c32(&p->d_pre, pre.data(), h);
        if (rc == ZAFB_OK) rc = upload_c32(&p->d_post, post.data(), h);
        if (rc == ZAFB_OK) rc = upload_twiddles(&p->d_tw_fft, h, h);
        if (rc == ZAFB_OK && (n == 4096 || n == 2048 || n == 1024 || n == 512)) {  // W_H^{k1*n2} laid out [k1][n2] for the warp kernels, H = n/4
            const int64_t hh = n / 4;
            std::vector<double> t(2 * hh);
            for (int64_t k1 = 0; k1 < hh / 32; ++k1)
                for (int64_t n2 = 0; n2 < 32; ++n2) {
                    const double a = -2.0 * pi * double((k1 * n2) % hh) / double(hh);
                    t[2 * (k1 * 32 + n2)] = std::cos(a);
                    t[2 * (k1 * 32 + n2) + 1] = std::sin(a);
                }
            rc = upload_c32(&p->d_tw_4step, t.data(), hh);
        }
    } else if (rc == ZAFB_OK) {
        std::vector<double> c(8 * p->m);
        for (int64_t t = 0; t < 8 * p->m; ++t) c[t] = std::cos(2.0 * pi * double(t) / double(8 * p->m));
        rc = upload_f32(&p->d_cos, c.data(), c.size());
    }
    if (rc != ZAFB_OK) {
        zafb_mdct_plan_destroy(p);
        return rc;
    }
    *out = p;
    return ZAFB_OK;
}

int zafb_mdct_plan_destroy(zafb_mdct_plan* p) {
    if (!p) return ZAFB_OK;
    cudaFree(p->d_window);
    cudaFree(p->d_tw_fft);
    cudaFree(p->d_pre);
    cudaFree(p->d_post);
    cudaFree(p->d_cos);
    cudaFree(p->d_tw_4step);
    delete p;
    return ZAFB_OK;
}

// test hook: 0 = auto, 1 = generic kernels only, 2 = require the warp kernels
int zafb_mdct_plan_force_kernel(zafb_mdct_plan* p, int which) {
    ZAFB_REQUIRE(p != nullptr && which >= 0 && which <= 2, "bad plan / kernel id");
    p->force_kernel = which;
    return ZAFB_OK;
}

int zafb_mdct_f32(const zafb_mdct_plan* p, const float* x, int64_t n_clips, int64_t ns, int64_t clip_stride,
                  float* out, int layout, void* stream) {
    ZAFB_REQUIRE(p != nullptr, "plan is NULL");
    ZAFB_REQUIRE(n_clips >= 0 && ns >= 0 && clip_stride >= ns, "bad batch geometry");
    ZAFB_REQUIRE(layout == ZAFB_LAYOUT_FRAME_MAJOR || layout == ZAFB_LAYOUT_BIN_MAJOR, "bad layout %d", layout);
    int rc = set_kernel_attrs();
    if (rc != ZAFB_OK) return rc;
    int64_t nt = 0;
    zafb_mdct_geometry(ns, p->n, nullptr, &nt, nullptr);
    const int64_t total = n_clips * nt;
    if (total == 0) return ZAFB_OK;
    ZAFB_REQUIRE(out != nullptr && (x != nullptr || ns == 0), "x/out is NULL");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int m = int(p->m);
    {
        const bool x_aligned = reinterpret_cast<uintptr_t>(x) % 8 == 0 && (n_clips <= 1 || clip_stride % 2 == 0);
        // the reference's C-order memory, written directly at any 4-byte phase of the result (ZAFB_MDCT_BM_DIRECT=0:
        // frame-major scratch + tiled transpose)
        if (layout == ZAFB_LAYOUT_BIN_MAJOR && x_aligned && p->force_kernel != 1 && mdct_binmajor_supported(p, nt) &&
            env_flag("ZAFB_MDCT_BM_DIRECT", 1))
            return mdct_binmajor_launch(p, x, n_clips, ns, clip_stride, nt, out, st);
        const bool aligned = x_aligned && reinterpret_cast<uintptr_t>(out) % 8 == 0;
        const bool warp_ok = (p->n == 4096 || p->n == 2048 || p->n == 1024 || p->n == 512) && aligned;
        if (p->force_kernel == 2 && !warp_ok)
            return fail(ZAFB_E_UNSUPPORTED, "mdct warp kernel needs N = 512, 1024, 2048 or 4096, even clip_stride, 8-byte aligned x/out");
        if (warp_ok && p->force_kernel != 1) {
            auto run = [&](const float* xs, int64_t clips, float* dst) -> int {
                const int64_t frames = clips * nt;
                const bool big = p->n == 2048, huge = p->n == 4096, small = p->n == 512;
                const size_t smem = huge    ? (MdctGeom<4096>::TABLES + kWarps * MdctGeom<4096>::TILE) * sizeof(float2)
                                    : big   ? (MdctGeom<2048>::TABLES + kWarps * MdctGeom<2048>::TILE) * sizeof(float2)
                                    : small ? (MdctGeom<512>::TABLES + kWarps * MdctGeom<512>::TILE) * sizeof(float2)
                                            : (MdctGeom<1024>::TABLES + kWarps * MdctGeom<1024>::TILE) * sizeof(float2);
                // 80 registers, no spills: 3 CTAs (24 warps) per SM measured 4.5 % faster than 2 on cfg 4
                // (N = 4096: 64 points and 4 x 16 sample pairs per lane -- one CTA per SM)
                constexpr int occ = 3;
                const int occ_n = huge ? 1 : occ;
                int64_t ctas = ceil_div(frames, kWarps);
                if (ctas > int64_t(sm_count()) * occ_n) ctas = int64_t(sm_count()) * occ_n;
                auto kern = huge ? mdct_warp_kernel<4096, 1> : big ? mdct_warp_kernel<2048, occ>
                            : small ? mdct_warp_kernel<512, occ> : mdct_warp_kernel<1024, occ>;
                kern<<<unsigned(ctas), kWarps * 32, smem, st>>>(
                    xs, ns, clip_stride, nt, reinterpret_cast<const float2*>(p->d_window), p->d_tw_4step, p->d_pre, p->d_post,
                    dst, frames);
                ZAFB_LAUNCH_CHECK();
                return ZAFB_OK;
            };
            if (layout == ZAFB_LAYOUT_FRAME_MAJOR) return run(x, n_clips, out);
            return bin_major_from_frame_major(out, n_clips, nt, p->m, st, [&](int64_t c0, int64_t n, float* scratch) {
                return run(x + c0 * clip_stride, n, scratch);
            });
        }
    }
    // generic kernels (one CTA per frame) can store either layout, but BIN_MAJOR means one 4-byte element per row: that
    // layout goes through frame-major scratch and the tiled transpose (ZAFB_GENERIC_BM_DIRECT=1 keeps the strided stores)
    auto run_generic = [&](const float* xs, int64_t clips, float* dst, int lay) -> int {
        const int64_t frames = clips * nt;
        const int64_t g = frames < int64_t(sm_count()) * 32 ? frames : int64_t(sm_count()) * 32;
        if (p->log2m >= 1) {
            const size_t smem = size_t(m) * sizeof(float2) + 2 * size_t(m) * sizeof(float);
            if (smem > size_t(kMaxDynSmem)) return fail(ZAFB_E_UNSUPPORTED, "mdct: window too large for shared memory");
            mdct_generic_kernel<<<unsigned(g), threads_for(m / 2), smem, st>>>(xs, ns, clip_stride, nt, p->log2m, p->d_window,
                                                                               p->d_tw_fft, p->d_pre, p->d_post, dst, lay, frames);
        } else {
            const size_t smem = size_t(2 * m) * sizeof(float) + 16;
            if (smem > size_t(kMaxDynSmem)) return fail(ZAFB_E_UNSUPPORTED, "mdct: window too large for shared memory");
            int th = m < 256 ? ((m + 31) / 32) * 32 : 256;
            mdct_direct_kernel<<<unsigned(g), th, smem, st>>>(xs, ns, clip_stride, nt, m, p->d_window, p->d_cos, dst, lay, frames);
        }
        ZAFB_LAUNCH_CHECK();
        return ZAFB_OK;
    };
    if (layout == ZAFB_LAYOUT_FRAME_MAJOR || nt == 1 || m == 1 || env_flag("ZAFB_GENERIC_BM_DIRECT", 0))
        return run_generic(x, n_clips, out, layout);
    return bin_major_from_frame_major(out, n_clips, nt, p->m, st, [&](int64_t c0, int64_t nc, float* scratch) {
        return run_generic(x + c0 * clip_stride, nc, scratch, ZAFB_LAYOUT_FRAME_MAJOR);
    });
}

int zafb_imdct_f32(const zafb_mdct_plan* p, const float* spec, int64_t n_clips, int64_t nt, int layout, float* y,
                   int64_t y_stride, void* stream) {
    ZAFB_REQUIRE(p != nullptr, "plan is NULL");
    ZAFB_REQUIRE(n_clips >= 0 && nt >= 0, "bad batch geometry");
    ZAFB_REQUIRE(layout == ZAFB_LAYOUT_FRAME_MAJOR || layout == ZAFB_LAYOUT_BIN_MAJOR, "bad layout %d", layout);
    int rc = set_kernel_attrs();
    if (rc != ZAFB_OK) return rc;
    int64_t len = 0;
    zafb_imdct_geometry(p->m, nt, nullptr, &len);
    ZAFB_REQUIRE(y_stride >= len, "y_stride %lld < output length %lld", (long long)y_stride, (long long)len);
    if (n_clips == 0 || len == 0) return ZAFB_OK;
    ZAFB_REQUIRE(spec != nullptr && y != nullptr, "spec/y is NULL");
    const int m = int(p->m);
    // C-order input read directly (any 4-byte phase): one CTA per clip, so only when the batch fills the SMs
    // (ZAFB_IMDCT_BM_MIN_CLIPS, default half the SM count); ZAFB_IMDCT_BM_DIRECT=0: tiled transpose into scratch first
    if (layout == ZAFB_LAYOUT_BIN_MAJOR && p->force_kernel != 1 && mdct_binmajor_supported(p, nt) &&
        env_flag("ZAFB_IMDCT_BM_DIRECT", 1) && n_clips >= env_flag("ZAFB_IMDCT_BM_MIN_CLIPS", sm_count() / 2))
        return imdct_binmajor_launch(p, spec, n_clips, nt, len, y, y_stride, static_cast<cudaStream_t>(stream));
    {
        const bool warp_ok = (p->n == 4096 || p->n == 2048 || p->n == 1024 || p->n == 512) && reinterpret_cast<uintptr_t>(spec) % 8 == 0;
        if (p->force_kernel == 2 && !warp_ok)
            return fail(ZAFB_E_UNSUPPORTED, "imdct warp kernel needs N = 512, 1024, 2048 or 4096, 8-byte aligned spectra");
        if (warp_ok && p->force_kernel != 1) {
            cudaStream_t st = static_cast<cudaStream_t>(stream);
            auto run = [&](const float* sp, int64_t clips, float* yy) -> int {
                const int64_t nblocks = nt - 1;
                constexpr int occ = 2;  // the carried half-frame needs 32 more registers; 3 CTAs/SM would spill
                const bool big = p->n == 2048, huge = p->n == 4096, small = p->n == 512;
                const int occ_n = huge ? 1 : (small ? ZAFB_IMDCT_SMALL_OCC : occ);
                const int64_t resident_warps = int64_t(sm_count()) * occ_n * kWarps;
                int64_t best_len = nblocks, best_cost = INT64_MAX;
                for (int64_t l = nblocks < 8 ? nblocks : 8; l <= nblocks && l <= 2048; ++l) {
                    const int64_t runs = clips * ceil_div(nblocks, l);
                    const int64_t cost = ceil_div(runs, resident_warps) * (l + 1);
                    if (cost < best_cost || (cost == best_cost && l > best_len)) {
                        best_cost = cost;
                        best_len = l;
                    }
                }
                const int64_t runs_per_clip = ceil_div(nblocks, best_len);
                const int64_t total = clips * runs_per_clip;
                int64_t ctas = ceil_div(total, kWarps);
                if (ctas > int64_t(sm_count()) * occ_n) ctas = int64_t(sm_count()) * occ_n;
                const size_t smem = huge    ? (MdctGeom<4096>::TABLES + kWarps * MdctGeom<4096>::TILE) * sizeof(float2)
                                    : big   ? (MdctGeom<2048>::TABLES + kWarps * MdctGeom<2048>::TILE) * sizeof(float2)
                                    : small ? (MdctGeom<512>::TABLES + kWarps * MdctGeom<512>::TILE) * sizeof(float2)
                                            : (MdctGeom<1024>::TABLES + kWarps * MdctGeom<1024>::TILE) * sizeof(float2);
                const int y_aligned = (reinterpret_cast<uintptr_t>(yy) % 8 == 0 && (clips <= 1 || y_stride % 2 == 0)) ? 1 : 0;
                auto kern = huge ? imdct_warp_kernel<4096, 1> : big ? imdct_warp_kernel<2048, occ>
                            : small ? imdct_warp_kernel<512, ZAFB_IMDCT_SMALL_OCC> : imdct_warp_kernel<1024, occ>;
                kern<<<unsigned(ctas), kWarps * 32, smem, st>>>(
                    sp, nt, reinterpret_cast<const float2*>(p->d_window), p->d_tw_4step, p->d_pre, p->d_post, runs_per_clip,
                    int(best_len), total, len, yy, y_stride, y_aligned, env_flag("ZAFB_IMDCT_PREFETCH", 1));
                ZAFB_LAUNCH_CHECK();
                return ZAFB_OK;
            };
            if (layout == ZAFB_LAYOUT_FRAME_MAJOR) return run(spec, n_clips, y);
            return frame_major_from_bin_major(spec, n_clips, nt, p->m, st, [&](int64_t c0, int64_t nc, const float* scratch) {
                return run(scratch, nc, y + c0 * y_stride);
            });
        }
    }
    const int64_t hop_blocks = nt - 1;  // hop-blocks 1 .. nt-1 are written
    int64_t per_tile = 16;
    if (per_tile > hop_blocks) per_tile = hop_blocks;
    const int64_t tiles = ceil_div(hop_blocks, per_tile);
    const int h = m / 2 > 0 ? m / 2 : 1;
    const size_t smem = size_t(2 * h) * sizeof(float2) + 4 * size_t(m) * sizeof(float) + 16;
    if (smem > size_t(kMaxDynSmem)) return fail(ZAFB_E_UNSUPPORTED, "imdct: window too large for shared memory");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto run_generic = [&](const float* sp, int64_t clips, float* yy, int lay) -> int {
        const int64_t blocks = clips * tiles;
        if (blocks > 0x7fffffffLL) return fail(ZAFB_E_UNSUPPORTED, "imdct: too many tiles");
        imdct_tile_kernel<<<unsigned(blocks), threads_for(m / 2), smem, st>>>(
            sp, nt, m, p->log2m, lay, p->d_window, p->d_tw_fft, p->d_pre, p->d_post, p->d_cos, per_tile, tiles, len, yy,
            y_stride);
        ZAFB_LAUNCH_CHECK();
        return ZAFB_OK;
    };
    // BIN_MAJOR input would be read one 4-byte element per row: transpose it into frame-major scratch first
    if (layout == ZAFB_LAYOUT_FRAME_MAJOR || nt == 1 || m == 1 || env_flag("ZAFB_GENERIC_BM_DIRECT", 0))
        return run_generic(spec, n_clips, y, layout);
    return frame_major_from_bin_major(spec, n_clips, nt, p->m, st, [&](int64_t c0, int64_t nc, const float* scratch) {
        return run_generic(scratch, nc, y + c0 * y_stride, ZAFB_LAYOUT_FRAME_MAJOR);
    });
}

// ------------------------------------------------------------------ host-buffer pipelines
int zafb_mdct_host_f32(const zafb_mdct_plan* p, const float* x, int64_t n_clips, int64_t ns, int64_t clip_stride,
                       float* out, int layout) {
    ZAFB_REQUIRE(p != nullptr, "plan is NULL");
    ZAFB_REQUIRE(n_clips >= 0 && ns >= 0 && clip_stride >= ns, "bad batch geometry");
    int64_t nt = 0;
    zafb_mdct_geometry(ns, p->n, nullptr, &nt, nullptr);
    if (n_clips == 0) return ZAFB_OK;
    ZAFB_REQUIRE(out != nullptr && (x != nullptr || ns == 0), "x/out is NULL");
    const size_t out_clip = size_t(nt) * p->m * sizeof(float);
    const int64_t dpitch = (ns + 1) & ~int64_t(1);
    return run_host_pipeline(x, size_t(clip_stride) * sizeof(float), size_t(ns) * sizeof(float), size_t(dpitch) * sizeof(float),
                             out, out_clip, out_clip, out_clip, n_clips,
                             [&](void* d_in, void* d_out, int64_t, int64_t nc, cudaStream_t st) {
                                 return zafb_mdct_f32(p, static_cast<const float*>(d_in), nc, ns, dpitch,
                                                      static_cast<float*>(d_out), layout, st);
                             });
}

int zafb_imdct_host_f32(const zafb_mdct_plan* p, const float* spec, int64_t n_clips, int64_t nt, int layout, float* y,
                        int64_t y_stride) {
    ZAFB_REQUIRE(p != nullptr, "plan is NULL");
    ZAFB_REQUIRE(n_clips >= 0 && nt >= 0, "bad batch geometry");
    int64_t len = 0;
    zafb_imdct_geometry(p->m, nt, nullptr, &len);
    ZAFB_REQUIRE(y_stride >= len, "y_stride too small");
    if (n_clips == 0 || len == 0) return ZAFB_OK;
    ZAFB_REQUIRE(spec != nullptr && y != nullptr, "spec/y is NULL");
    const size_t in_clip = size_t(nt) * p->m * sizeof(float);
    // the reference's odd output length M(nt-1)-1 would misalign every other row: even device pitch
    const int64_t dpitch = (len + 1) & ~int64_t(1);
    return run_host_pipeline(spec, in_clip, in_clip, in_clip, y, size_t(y_stride) * sizeof(float), size_t(len) * sizeof(float),
                             size_t(dpitch) * sizeof(float), n_clips,
                             [&](void* d_in, void* d_out, int64_t, int64_t nc, cudaStream_t st) {
                                 return zafb_imdct_f32(p, static_cast<const float*>(d_in), nc, nt, layout,
                                                       static_cast<float*>(d_out), dpitch, st);
                             });
}

}  // extern "C"
