// DCT / DST types I..IV, orthonormal (zaf.py:703-839, 842-981), batched over vectors.
//
// Every one of the eight transforms is  out[k] = s_out[k] * sum_n s_in[n] x[n] T[(a(n) b(k)) mod P]
// with T = cos or sin of 2 pi t / P and affine index maps a(n) = a0 + da n, b(k) = b0 + db k:
//
//   DCT-I   P = 2(N-1)  a = n     b = k      s_in = e_n            s_out = sqrt(2/(N-1)) e_k   (e_0 = e_{N-1} = 1/sqrt2)
//   DCT-II  P = 4N      a = 2n+1  b = k      s_in = 1              s_out = sqrt(2/N) c_k       (c_0 = 1/sqrt2)
//   DCT-III P = 4N      a = n     b = 2k+1   s_in = c_n            s_out = sqrt(2/N)
//   DCT-IV  P = 8N      a = 2n+1  b = 2k+1   s_in = 1              s_out = sqrt(2/N)
//   DST-I   P = 2(N+1)  a = n+1   b = k+1    s_in = 1              s_out = sqrt(2/(N+1))
//   DST-II  P = 4N      a = 2n+1  b = k+1    s_in = 1              s_out = sqrt(2/N) d_k       (d_{N-1} = 1/sqrt2)
//   DST-III P = 4N      a = n+1   b = 2k+1   s_in = d_n            s_out = sqrt(2/N)
//   DST-IV  P = 8N      a = 2n+1  b = 2k+1   s_in = 1              s_out = sqrt(2/N)
//
// (closed forms verified against the reference's mirrored/zero-stuffed FFT constructions).
// The reference spends a 2(N-1)- to 8N-point complex FFT per vector on this; here:
//   dct_fft_kernel     power-of-two N, types II/III/IV (and the DST twins through reversal / sign maps):
//                      one N/2-point complex FFT in shared memory per vector.
//   dct_direct_kernel  everything else (types I, whose natural FFT lengths 2(N-1) / 2(N+1) are not powers of two,
//                      and non-power-of-two N): the trigonometric sum itself, table-driven, exact index arithmetic.
#include <cmath>
#include <cstdint>
#include <vector>

#include "fft_core.cuh"
#include "host_pipe.cuh"
#include "gemm_tc.cuh"

using namespace zafb;

struct zafb_dct_plan {
    int kind = 0, type = 0;
    int64_t n = 0;
    int period = 0, a0 = 0, da = 0, b0 = 0, db = 0;
    float* d_tab = nullptr;     // period floats
    float* d_sin = nullptr;     // n floats
    float* d_sout = nullptr;    // n floats
    // FFT path
    int log2n = -1;
    float2* d_tw_fft = nullptr;   // W_{N/2}^t
    float2* d_tw_a = nullptr;     // per-type twiddles (see kernels)
    float2* d_tw_b = nullptr;
    float2* d_tw_4step = nullptr; // n == 1024: W_512^{k1*n2} at [k1*32 + n2], k1 < 16 (warp kernel)
    int force_direct = 0;         // test hook: 1 = direct kernel, 2 = require the tensor-core matrix path,
                                  //            3 = block FFT kernel (no warp kernel), 4 = require the warp kernel
    // matrix path (types I and non-power-of-two N): out = x . Mat^T on the tensor cores, Mat[k][n] in hi/lo TF32 halves
    float* d_mat_hi = nullptr;
    float* d_mat_lo = nullptr;
    int64_t ldk = 0;              // row pitch of the matrix and of the split input (n rounded up to 4)
    // even / odd form of the matrix path (types I and II): Mat[k][n-1-m] = (-1)^k Mat[k][m], so the even outputs are a
    // product with s[m] = x[m] + x[n-1-m] and the odd ones with d[m] = x[m] - x[n-1-m] -- two half-size products, run as
    // ONE tiled launch over the stacked operand [Mat_even ; Mat_odd] and the folded input [s | d]
    float* d_eo_hi = nullptr;
    float* d_eo_lo = nullptr;
    GemmTile* d_eo_tiles = nullptr;
    int eo_tiles = 0;
    int64_t eo_ldb = 0, eo_lda = 0, eo_dcol = 0;  // operand pitches, first column of d in the folded input
    // the same operands cut into 256-row blocks for the CTA-pair kernel (gemm3xtf32_pair_tiled); dense form likewise
    float* d_eo2_hi = nullptr;
    float* d_eo2_lo = nullptr;
    GemmTile* d_eo2_tiles = nullptr;
    int eo2_tiles = 0;
    GemmTile* d_mat2_tiles = nullptr;
    int mat2_tiles = 0;
};

namespace {

constexpr int kMaxDynSmem = 200 * 1024;

__global__ void dct_direct_kernel(const float* __restrict__ x, int64_t batch, int64_t stride, int n, int period, int a0,
                                  int da, int b0, int db, const float* __restrict__ tab, const float* __restrict__ s_in,
                                  const float* __restrict__ s_out, float* __restrict__ out, int64_t out_stride) {
    extern __shared__ float2 smem2[];
    float* xs = reinterpret_cast<float*>(smem2);
    const int tid = threadIdx.x, nth = blockDim.x;
    for (int64_t v = blockIdx.x; v < batch; v += gridDim.x) {
        for (int i = tid; i < n; i += nth) xs[i] = x[v * stride + i] * s_in[i];
        __syncthreads();
        for (int k = tid; k < n; k += nth) {
            const int64_t b = b0 + int64_t(db) * k;
            int idx = int((int64_t(a0) * b) % period);
            const int step = int((int64_t(da) * b) % period);
            float acc0 = 0.f, acc1 = 0.f;
            int i = 0;
            for (; i + 1 < n; i += 2) {
                acc0 = fmaf(xs[i], tab[idx], acc0);
                idx += step;
                if (idx >= period) idx -= period;
                acc1 = fmaf(xs[i + 1], tab[idx], acc1);
                idx += step;
                if (idx >= period) idx -= period;
            }
            if (i < n) acc0 = fmaf(xs[i], tab[idx], acc0);
            out[v * out_stride + k] = (acc0 + acc1) * s_out[k];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// FFT path, power-of-two N >= 4.  mode: 2 = DCT-II core, 3 = DCT-III core, 4 = DCT-IV core.
// DST twins: flags bit0 = reverse the input, bit1 = multiply input n by (-1)^n, bit2 = reverse the output,
// bit3 = multiply output k by (-1)^k:
//   DST-II(x)  = reverse(DCT-II((-1)^n x))      DST-III(x) = (-1)^k DCT-III(reverse(x))
//   DST-IV(x)  = (-1)^k DCT-IV(reverse(x))
//
// DCT-II (Makhoul): v = [x0, x2, ..., x_{N-2}, x_{N-1}, ..., x3, x1]; V = FFT_N(v) (real input, via the N/2-point
//   complex FFT of v[2m] + i v[2m+1]); X_k = sqrt(2/N) c_k Re(V_k e^{-i pi k / 2N}).
// DCT-III: the exact inverse: V_k = (C_k - i C_{N-k}) e^{+i pi k / 2N} with C = input / (sqrt(2/N) c_k) ... folded into
//   tables; v = Re(IFFT_N(V)) computed with the half-size complex FFT; un-permute.
// DCT-IV: t[m] = (v[2m] + i v[N-1-2m]) e^{-i pi m / N}; y = FFT_{N/2}(t) e^{-i pi (m + 1/4)/N}; out[2m] = Re y, out[N-1-2m] = -Im y.
__device__ __forceinline__ float load_in(const float* __restrict__ xv, int n, int i, int flags) {
    const int src = (flags & 1) ? (n - 1 - i) : i;
    float val = xv[src];
    if ((flags & 2) && (src & 1)) val = -val;
    return val;
}
__device__ __forceinline__ void store_out(float* __restrict__ ov, int n, int k, float val, int flags) {
    const int dst = (flags & 4) ? (n - 1 - k) : k;
    if ((flags & 8) && (dst & 1)) val = -val;
    ov[dst] = val;
}

__global__ void dct_fft_kernel(const float* __restrict__ x, int64_t batch, int64_t stride, int log2n, int mode, int flags,
                               const float2* __restrict__ tw_fft, const float2* __restrict__ tw_a,
                               const float2* __restrict__ tw_b, float* __restrict__ out, int64_t out_stride) {
    extern __shared__ float2 smem2[];
    const int n = 1 << log2n, h = n >> 1;
    float2* a = smem2;
    float2* b = smem2 + h;
    float* s = reinterpret_cast<float*>(smem2 + 2 * h);  // n floats of staging
    const int tid = threadIdx.x, nth = blockDim.x;
    const float norm = sqrtf(2.0f / float(n));
    for (int64_t v = blockIdx.x; v < batch; v += gridDim.x) {
        const float* xv = x + v * stride;
        float* ov = out + v * out_stride;
        if (mode == 4) {
            for (int i = tid; i < n; i += nth) s[i] = load_in(xv, n, i, flags);
            __syncthreads();
            for (int m = tid; m < h; m += nth) a[m] = cmul(make_float2(s[2 * m], s[n - 1 - 2 * m]), tw_a[m]);
            __syncthreads();
            const float2* y = block_fft(a, b, tw_fft, log2n - 1, tid, nth);
            for (int m = tid; m < h; m += nth) {
                const float2 r = cmul(y[m], tw_b[m]);
                store_out(ov, n, 2 * m, norm * r.x, flags);
                store_out(ov, n, n - 1 - 2 * m, -norm * r.y, flags);
            }
        } else if (mode == 2) {
            // Makhoul permutation: v[i] = x[2i] (i < N/2), v[N-1-i] = x[2i+1]
            for (int i = tid; i < h; i += nth) {
                s[i] = load_in(xv, n, 2 * i, flags);
                s[n - 1 - i] = load_in(xv, n, 2 * i + 1, flags);
            }
            __syncthreads();
            for (int m = tid; m < h; m += nth) a[m] = make_float2(s[2 * m], s[2 * m + 1]);
            __syncthreads();
            const float2* z = block_fft(a, b, tw_fft, log2n - 1, tid, nth);
            // real-input unpack: V_k = E + w_k O, k = 0..N/2 ; tw_a[k] = W_N^k (k < N/2); tw_b[k] = sqrt(2/N) c_k e^{-i pi k/2N}, k < N
            for (int k = tid; k <= h; k += nth) {
                const float2 zk = z[k & (h - 1)];
                const float2 zp = z[(h - k) & (h - 1)];
                const float2 e = make_float2(0.5f * (zk.x + zp.x), 0.5f * (zk.y - zp.y));
                const float2 od = make_float2(0.5f * (zk.y + zp.y), 0.5f * (zp.x - zk.x));
                float2 vk;
                if (k == h) vk = make_float2(e.x - od.x, 0.f);   // V_{N/2} = E_0 - O_0 (k & (h-1) == 0)
                else vk = cadd(e, cmul(tw_a[k], od));
                // X_k = Re(V_k t_k);  X_{N-k} = Re(conj(V_k) t_{N-k})
                store_out(ov, n, k, vk.x * tw_b[k].x - vk.y * tw_b[k].y, flags);
                if (k > 0 && k < h) store_out(ov, n, n - k, vk.x * tw_b[n - k].x + vk.y * tw_b[n - k].y, flags);
            }
        } else {  // mode 3
            // C_k = x_k / c_k (c_0 = 1/sqrt2):  V_k = (C_k - i C_{N-k}) u_k, u_k = e^{+i pi k/2N} (tw_b), C_N = 0
            for (int i = tid; i < n; i += nth) s[i] = load_in(xv, n, i, flags);
            __syncthreads();
            // half-size inverse: v = Re IFFT_N(V) with V Hermitian.  Z[k] = E_k + i O_k, E_k = (V_k + conj(V_{N/2-k}))/2 ... use
            // E_k = (V_k + V_{k+N/2})/2, O_k = conj(w_k) (V_k - V_{k+N/2})/2 with V_{k+N/2} = conj(V_{N/2-k}).
            for (int k = tid; k < h; k += nth) {
                auto V = [&](int q) -> float2 {  // q in [0, N/2]
                    const float cq = s[q] * (q == 0 ? 1.41421356237309505f : 1.0f);  // C_k = x_k / c_k
                    const float cn = (q == 0) ? 0.f : s[n - q];
                    return cmul(make_float2(cq, -cn), tw_b[q]);
                };
                const float2 vk = V(k);
                const float2 vc = cconj(V(h - k));          // V_{k+N/2}
                const float2 e = make_float2(0.5f * (vk.x + vc.x), 0.5f * (vk.y + vc.y));
                const float2 d = make_float2(0.5f * (vk.x - vc.x), 0.5f * (vk.y - vc.y));
                const float2 o = cmul_conj(d, tw_a[k]);     // conj(w_k) * d
                // inverse FFT through the forward one: IFFT(Z) = conj(FFT(conj(Z))) / (N/2)
                const float2 zz = make_float2(e.x - o.y, e.y + o.x);  // E + i O
                a[k] = cconj(zz);
            }
            __syncthreads();
            const float2* y = block_fft(a, b, tw_fft, log2n - 1, tid, nth);
            // v[2m] = Re z[m], v[2m+1] = Im z[m] with z = conj(y)/(N/2);  then x[2i] = v[i], x[2i+1] = v[N-1-i]
            // overall factor sqrt(N/2) (from C_k) * 1/(N/2) (inverse FFT) = sqrt(2/N)
            for (int m = tid; m < h; m += nth) {
                const float v0 = norm * y[m].x;
                const float v1 = -norm * y[m].y;
                // position p = 2m holds v0, p = 2m+1 holds v1;  v[i] -> x[2i] for i < N/2, v[N-1-i] -> x[2i+1]
                const int p0 = 2 * m, p1 = 2 * m + 1;
                const int d0 = (p0 < h) ? 2 * p0 : 2 * (n - 1 - p0) + 1;
                const int d1 = (p1 < h) ? 2 * p1 : 2 * (n - 1 - p1) + 1;
                store_out(ov, n, d0, v0, flags);
                store_out(ov, n, d1, v1, flags);
            }
        }
        __syncthreads();
    }
}


// ---------------------------------------------------------------------------------------------
// N = 1024 (the reference's example length, zaf.py:726-753): one warp per vector, types II / III / IV
// and their DST twins.  The 1024 real values are packed into ONE 512-point complex FFT held in the
// registers of the warp (warp_fft512: two in-register radix-16 passes around one shared-memory
// transpose); permutations, real-input split and twiddles are register renaming, xor-31 / mirror
// shuffles and immediates.  Loads and stores are 8- or 16-byte coalesced, one pass over HBM.
//   MODE 4  t[m] = (P[m].x + i P[511-m].y) e^{-i pi m/N},  P[p] = (x[2p], x[2p+1])   (same core as the MDCT kernel)
//   MODE 2  Makhoul: z[q] = (x[4q], x[4q+2]), z[511-q] = (x[4q+3], x[4q+1]), q < 256;  V = real-input split of FFT(z);
//           U_k = V_k e^{-i pi k/2N}:  X_k = s c_k Re U_k,  X_{N-k} = -s Im U_k,  X_{N/2} = s (Re Z_0 - Im Z_0)/sqrt2
//   MODE 3  the exact inverse of MODE 2 (zaf.py:799-817 builds it from a 4N-point FFT instead)
// DST twins: DST-II(x) = reverse(DCT-II((-1)^n x)), DST-III(x) = (-1)^k DCT-III(reverse x),
// DST-IV(x) = (-1)^k DCT-IV(reverse x).   (All index maps validated in float64 against scipy's orthonormal transforms.)
// ---------------------------------------------------------------------------------------------
constexpr int kDctWarps = 8;

// N = 512, 1024, 2048, 4096 (r02: the same code, REGS = N / 64 points per lane; warp_fft256 / warp_fft512 / warp_fft1024 /
// warp_fft2048)
template <int N, int MODE, bool DST>
__global__ void __launch_bounds__(kDctWarps * 32, N >= 2048 ? 1 : (MODE == 3 ? 2 : 3))   // type III holds the N inputs AND N/2 products
dct_warp_kernel(const float* __restrict__ x, int64_t batch, int64_t stride, const float2* __restrict__ tw4,
                    const float2* __restrict__ tw_a, const float2* __restrict__ tw_b, float* __restrict__ out,
                    int64_t out_stride) {
    constexpr int H = N / 2, REGS = H / 32, LOGR = clog2(REGS);
    static_assert(N == 512 || N == 1024 || N == 2048 || N == 4096, "dct warp kernel: N = 512, 1024, 2048 or 4096");
    extern __shared__ float2 smem2[];
    float2* s_tw = smem2;  // H: W_H^{k1 n2}
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2* s_buf = smem2 + H + warp * (REGS * kFft1024Pitch);
    for (int i = tid; i < H; i += kDctWarps * 32) s_tw[i] = tw4[i];
    constexpr float kNorm = N == 512 ? 0.0625f : N == 1024 ? 0.04419417382415922f : N == 2048 ? 0.03125f : 0.022097086912079608f;   // sqrt(2 / N)
    float2 tq[N == 512 ? 8 : 1];
    if constexpr (N == 512) warp_fft256_lane_twiddles(tq, lane);
    auto warp_fft = [&](float2 (&a)[REGS]) {
        if constexpr (N == 512) warp_fft256(a, s_tw, s_buf, lane, tq);
        else if constexpr (N == 1024) warp_fft512(a, s_tw, s_buf, lane);
        else if constexpr (N == 2048) warp_fft1024<false>(a, s_tw, s_buf, lane);
        else warp_fft2048(a, s_tw, s_buf, lane);
    };
    constexpr float kR2 = 0.70710678118654752f;
    const float2 ta = tw_a[lane];
    float2 tb = tw_b[lane];
    if constexpr (MODE == 4) tb = cscale(tb, kNorm);                       // post twiddle carries sqrt(2/N)
    if constexpr (MODE == 2) tb = lane == 0 ? make_float2(0.5f * kNorm, 0.f) : cscale(tb, 0.5f);  // s/2 e^{-i pi lane/2N}, no c_k
    if constexpr (MODE == 3) tb = cscale(tb, 0.5f * kNorm);                // s/2 e^{+i pi lane/2N}
    __syncthreads();
    const int mirror = (32 - lane) & 31;

    for (int64_t vec = int64_t(blockIdx.x) * kDctWarps + warp; vec < batch; vec += int64_t(gridDim.x) * kDctWarps) {
        const float* xv = x + vec * stride;
        float* ov = out + vec * out_stride;
        float2 v[REGS];

        if constexpr (MODE == 4) {
            float2 xp[REGS];
            const float2* P = reinterpret_cast<const float2*>(xv) + lane;
#pragma unroll
            for (int r = 0; r < REGS; ++r) xp[r] = __ldg(P + 32 * r);
            static_for<0, REGS>([&](auto rc) {
                constexpr int r = decltype(rc)::value;
                const float other = __shfl_xor_sync(0xffffffffu, xp[REGS - 1 - r].y, 31);  // x[N-1-2m]
                const float2 t = DST ? make_float2(other, xp[r].x) : make_float2(xp[r].x, other);
                v[r] = cmul(t, mul_tw<r, N / 16>(ta));
            });
            warp_fft(v);
            static_for<0, REGS>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                v[bitrev(k, LOGR)] = cmul(v[bitrev(k, LOGR)], mul_tw<k, N / 16>(tb));
            });
            float2* o = reinterpret_cast<float2*>(ov) + lane;
            static_for<0, REGS>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                const float im = __shfl_xor_sync(0xffffffffu, v[bitrev(REGS - 1 - k, LOGR)].y, 31);
                __stcs(o + 32 * k, make_float2(v[bitrev(k, LOGR)].x, DST ? im : -im));
            });
        } else if constexpr (MODE == 2) {
            const float4* Q = reinterpret_cast<const float4*>(xv) + lane;
            static_for<0, REGS / 2>([&](auto rc) {
                constexpr int r = decltype(rc)::value;
                float4 q = __ldg(Q + 32 * r);
                if constexpr (DST) {  // (-1)^n x[n]
                    q.y = -q.y;
                    q.w = -q.w;
                }
                v[r] = make_float2(q.x, q.z);                                   // z[q]
                v[REGS - 1 - r].x = __shfl_xor_sync(0xffffffffu, q.w, 31);     // z[H - 1 - q'] of lane 31 - lane
                v[REGS - 1 - r].y = __shfl_xor_sync(0xffffffffu, q.y, 31);
            });
            warp_fft(v);  // Z[lane + 32 k2] = v[bitrev(k2, LOGR)]
            static_for<0, REGS>([&](auto kc) {
                constexpr int k2 = decltype(kc)::value;
                const float2 z = v[bitrev(k2, LOGR)];
                const float2 mine = v[bitrev(REGS - 1 - k2, LOGR)];
                float2 p;
                p.x = __shfl_sync(0xffffffffu, mine.x, mirror);
                p.y = __shfl_sync(0xffffffffu, mine.y, mirror);
                if (lane == 0) p = v[bitrev((REGS - k2) & (REGS - 1), LOGR)];
                const float2 e = make_float2(z.x + p.x, z.y - p.y);
                const float2 od = make_float2(z.y + p.y, p.x - z.x);
                const float2 V = cadd(e, cmul(mul_tw<k2, N / 32>(ta), od));
                const float2 U = cmul(V, mul_tw<k2, N / 8>(tb));
                float lo = U.x, hi = -U.y;
                const int k = lane + 32 * k2;
                int ihi = N - k;
                if (k2 == 0 && lane == 0) {  // k = 0: X_0 carries c_0 = 1/sqrt2; the mirror slot holds X_{N/2}
                    lo = tb.x * (e.x + od.x) * kR2;
                    hi = tb.x * (e.x - od.x) * kR2;
                    ihi = H;
                }
                if constexpr (DST) {  // reversed output
                    ov[N - 1 - k] = lo;
                    ov[N - 1 - ihi] = hi;
                } else {
                    ov[k] = lo;
                    ov[ihi] = hi;
                }
            });
        } else {  // MODE 3
            float xr[2 * REGS];
#pragma unroll
            for (int i = 0; i < 2 * REGS; ++i) xr[i] = DST ? __ldg(xv + N - 1 - lane - 32 * i) : __ldg(xv + lane + 32 * i);
            // V_k = (C_k - i C_{N-k}) u_k (scaled by s/2), k = lane + 32 r
            static_for<0, REGS>([&](auto rc) {
                constexpr int r = decltype(rc)::value;
                float ck = xr[r];
                float cn = __shfl_sync(0xffffffffu, xr[2 * REGS - 1 - r], mirror);
                if (lane == 0) {
                    cn = r == 0 ? 0.f : xr[(2 * REGS - r) & (2 * REGS - 1)];
                    if (r == 0) ck *= 1.41421356237309505f;  // C_0 = x_0 / c_0
                }
                v[r] = cmul(make_float2(ck, -cn), mul_tw<N / 8 - r, N / 8>(tb));
            });
            float2 zin[REGS];
            static_for<0, REGS>([&](auto rc) {
                constexpr int r = decltype(rc)::value;
                float2 pv;
                pv.x = __shfl_sync(0xffffffffu, v[REGS - 1 - r].x, mirror);
                pv.y = __shfl_sync(0xffffffffu, v[REGS - 1 - r].y, mirror);
                if (lane == 0) {
                    if constexpr (r == 0) pv = make_float2(1.41421356237309505f * xr[REGS] * tb.x, 0.f);  // V_{N/2} is real
                    else pv = v[REGS - r];
                }
                const float2 e = make_float2(v[r].x + pv.x, v[r].y - pv.y);   // V_k + conj(V_{N/2-k})
                const float2 d = make_float2(v[r].x - pv.x, v[r].y + pv.y);
                const float2 o = cmul_conj(d, mul_tw<r, N / 32>(ta));
                // Z = E + i O; the inverse FFT runs as conj(FFT(conj Z))
                zin[r] = make_float2(e.x - o.y, -(e.y + o.x));
            });
            warp_fft(zin);  // conj(z[lane + 32 k2]) = zin[bitrev(k2, LOGR)]
            float4* o4 = reinterpret_cast<float4*>(ov) + lane;
            static_for<0, REGS / 2>([&](auto rc) {
                constexpr int r = decltype(rc)::value;
                const float2 own = zin[bitrev(r, LOGR)];
                float2 oth;
                oth.x = __shfl_xor_sync(0xffffffffu, zin[bitrev(REGS - 1 - r, LOGR)].x, 31);
                oth.y = __shfl_xor_sync(0xffffffffu, zin[bitrev(REGS - 1 - r, LOGR)].y, 31);
                // x[4q] = Re z[q], x[4q+1] = Im z[511-q], x[4q+2] = Im z[q], x[4q+3] = Re z[511-q]
                float4 q = make_float4(own.x, -oth.y, -own.y, oth.x);
                if constexpr (DST) {
                    q.y = -q.y;
                    q.w = -q.w;
                }
                __stcs(o4 + 32 * r, q);
            });
        }
    }
}

// Input of the even / odd matrix path: s[m] = x[m] + x[n-1-m] (the middle sample of an odd n once) in columns [0, ceil(n/2)),
// d[m] = x[m] - x[n-1-m] in columns [dcol, dcol + n/2), zeros elsewhere, split into TF32 hi / lo halves (row pitch ld).
__global__ void fold_split_tf32_kernel(const float* __restrict__ x, int64_t rows, int n, int64_t ldx, float* __restrict__ hi,
                                       float* __restrict__ lo, int64_t ld, int64_t dcol) {
    const int nh = (n + 1) / 2, nd = n / 2;
    const int64_t total = rows * ld;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int64_t r = i / ld;
        const int c = int(i - r * ld);
        const float* xr = x + r * ldx;
        float v = 0.f;
        if (c < nh) v = (c == n - 1 - c) ? xr[c] : xr[c] + xr[n - 1 - c];
        else if (c >= dcol && c < dcol + nd) v = xr[c - dcol] - xr[n - 1 - (c - dcol)];
        uint32_t h, l;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
        const float hv = __uint_as_float(h);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(v - hv));
        hi[i] = hv;
        lo[i] = __uint_as_float(l);
    }
}

// The same for n a multiple of 8 (no padding columns: ld = n, s in columns [0, n/2), d in [n/2, n)): one thread per
// group of four columns of s AND d -- two 16-byte loads (forward and mirrored), four 16-byte stores, no index division.
__device__ __forceinline__ void split4(const float4 v, float4& h, float4& l) {
    auto one = [](float x, float& hv, float& lv) {
        uint32_t a, b;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(a) : "f"(x));
        hv = __uint_as_float(a);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(x - hv));
        lv = __uint_as_float(b);
    };
    one(v.x, h.x, l.x);
    one(v.y, h.y, l.y);
    one(v.z, h.z, l.z);
    one(v.w, h.w, l.w);
}
__global__ void fold_split_tf32_vec_kernel(const float* __restrict__ x, int64_t rows, int n, int64_t ldx, float* __restrict__ hi,
                                           float* __restrict__ lo) {
    const int groups = n / 8;  // float4 groups per half row
    const int rows_per_block = blockDim.x / groups > 0 ? blockDim.x / groups : 1;
    const int g0 = threadIdx.x % groups, rr = threadIdx.x / groups;
    if (rr >= rows_per_block && blockDim.x >= groups) return;
    for (int64_t r = int64_t(blockIdx.x) * rows_per_block + rr; r < rows; r += int64_t(gridDim.x) * rows_per_block) {
        const float* xr = x + r * ldx;
        for (int g = g0; g < groups; g += blockDim.x) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(xr + 4 * g));
            const float4 b = __ldg(reinterpret_cast<const float4*>(xr + n - 4 - 4 * g));
            const float4 s = make_float4(a.x + b.w, a.y + b.z, a.z + b.y, a.w + b.x);
            const float4 d = make_float4(a.x - b.w, a.y - b.z, a.z - b.y, a.w - b.x);
            float4 h, l;
            split4(s, h, l);
            *reinterpret_cast<float4*>(hi + r * n + 4 * g) = h;
            *reinterpret_cast<float4*>(lo + r * n + 4 * g) = l;
            split4(d, h, l);
            *reinterpret_cast<float4*>(hi + r * n + n / 2 + 4 * g) = h;
            *reinterpret_cast<float4*>(lo + r * n + n / 2 + 4 * g) = l;
        }
    }
}

bool g_attr_done = false;
int set_kernel_attrs() {
    if (g_attr_done) return ZAFB_OK;
    ZAFB_CUDA(cudaFuncSetAttribute(dct_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    ZAFB_CUDA(cudaFuncSetAttribute(dct_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<512, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<1024, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<2048, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<4096, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<512, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<1024, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<2048, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<4096, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<512, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<1024, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<2048, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<4096, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<512, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<1024, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<2048, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<4096, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<512, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<1024, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<2048, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<4096, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<512, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<1024, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<2048, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(dct_warp_kernel<4096, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    g_attr_done = true;
    return ZAFB_OK;
}

}  // namespace

extern "C" {

int zafb_dct_plan_create(zafb_dct_plan** out, int kind, int type, int64_t n) {
    ZAFB_REQUIRE(out != nullptr, "plan is NULL");
    ZAFB_REQUIRE(kind == 0 || kind == 1, "kind must be 0 (DCT) or 1 (DST)");
    ZAFB_REQUIRE(type >= 1 && type <= 4, "type must be 1..4");
    ZAFB_REQUIRE(n >= 1, "vector length must be >= 1");
    ZAFB_REQUIRE(!(kind == 0 && type == 1 && n < 2), "DCT-I needs at least 2 samples");
    if (n > 16384) return fail(ZAFB_E_UNSUPPORTED, "dct/dst: vector length %lld too large", (long long)n);
    zafb_dct_plan* p = new zafb_dct_plan();
    p->kind = kind;
    p->type = type;
    p->n = n;
    const double pi = 3.14159265358979323846264338327950288;
    const double r2 = std::sqrt(0.5);
    std::vector<double> sin_(n, 1.0), sout(n, 1.0);
    double norm = std::sqrt(2.0 / double(n));
    if (type == 1) {
        if (kind == 0) {
            p->period = int(2 * (n - 1)); p->a0 = 0; p->da = 1; p->b0 = 0; p->db = 1;
            norm = std::sqrt(2.0 / double(n - 1));
            sin_[0] = sin_[n - 1] = r2;
            for (auto& v : sout) v = norm;
            sout[0] *= r2;
            sout[n - 1] *= r2;
        } else {
            p->period = int(2 * (n + 1)); p->a0 = 1; p->da = 1; p->b0 = 1; p->db = 1;
            norm = std::sqrt(2.0 / double(n + 1));
            for (auto& v : sout) v = norm;
        }
    } else if (type == 2) {
        p->period = int(4 * n); p->a0 = 1; p->da = 2; p->b0 = kind; p->db = 1;
        for (auto& v : sout) v = norm;
        sout[kind == 0 ? 0 : n - 1] *= r2;
    } else if (type == 3) {
        p->period = int(4 * n); p->a0 = kind; p->da = 1; p->b0 = 1; p->db = 2;
        for (auto& v : sout) v = norm;
        sin_[kind == 0 ? 0 : n - 1] = r2;
    } else {
        p->period = int(8 * n); p->a0 = 1; p->da = 2; p->b0 = 1; p->db = 2;
        for (auto& v : sout) v = norm;
    }
    std::vector<double> tab(p->period);
    for (int t = 0; t < p->period; ++t) {
        const double a = 2.0 * pi * double(t) / double(p->period);
        tab[t] = kind == 0 ? std::cos(a) : std::sin(a);
    }
    int rc = upload_f32(&p->d_tab, tab.data(), tab.size());
    if (rc == ZAFB_OK) rc = upload_f32(&p->d_sin, sin_.data(), n);
    if (rc == ZAFB_OK) rc = upload_f32(&p->d_sout, sout.data(), n);

    // FFT path tables (power-of-two N >= 4, types II..IV)
    if (rc == ZAFB_OK && type >= 2 && is_pow2(n) && n >= 4) {
        p->log2n = ilog2(n);
        const int64_t h = n / 2;
        rc = upload_twiddles(&p->d_tw_fft, h, h);
        std::vector<double> ta, tb;
        if (type == 4) {
            ta.resize(2 * h);
            tb.resize(2 * h);
            for (int64_t m = 0; m < h; ++m) {
                ta[2 * m] = std::cos(-pi * double(m) / double(n));
                ta[2 * m + 1] = std::sin(-pi * double(m) / double(n));
                tb[2 * m] = std::cos(-pi * (double(m) + 0.25) / double(n));
                tb[2 * m + 1] = std::sin(-pi * (double(m) + 0.25) / double(n));
            }
        } else {
            // tw_a[k] = W_N^k, k < N/2
            ta.resize(2 * h);
            for (int64_t k = 0; k < h; ++k) {
                ta[2 * k] = std::cos(-2.0 * pi * double(k) / double(n));
                ta[2 * k + 1] = std::sin(-2.0 * pi * double(k) / double(n));
            }
            tb.resize(2 * n);
            for (int64_t k = 0; k < n; ++k) {
                if (type == 2) {  // sqrt(2/N) c_k e^{-i pi k / 2N}
                    const double c = norm * (k == 0 ? r2 : 1.0);
                    tb[2 * k] = c * std::cos(-pi * double(k) / double(2 * n));
                    tb[2 * k + 1] = c * std::sin(-pi * double(k) / double(2 * n));
                } else {  // e^{+i pi k / 2N}
                    tb[2 * k] = std::cos(pi * double(k) / double(2 * n));
                    tb[2 * k + 1] = std::sin(pi * double(k) / double(2 * n));
                }
            }
        }
        if (rc == ZAFB_OK) rc = upload_c32(&p->d_tw_a, ta.data(), ta.size() / 2);
        if (rc == ZAFB_OK) rc = upload_c32(&p->d_tw_b, tb.data(), tb.size() / 2);
    }
    if (rc == ZAFB_OK && p->log2n >= 9 && p->log2n <= 12) {  // W_H^{k1*n2} laid out [k1][n2] for the warp kernels (N = 512 ... 4096)
        const int64_t hh = n / 2;
        std::vector<double> t(2 * hh);
        for (int64_t k1 = 0; k1 < hh / 32; ++k1)
            for (int64_t n2 = 0; n2 < 32; ++n2) {
                const double a = -2.0 * pi * double((k1 * n2) % hh) / double(hh);
                t[2 * (k1 * 32 + n2)] = std::cos(a);
                t[2 * (k1 * 32 + n2) + 1] = std::sin(a);
            }
        rc = upload_c32(&p->d_tw_4step, t.data(), hh);
    }
    // matrix path for everything the FFT path does not cover: Mat[k][m] = s_out[k] s_in[m] T[(a(m) b(k)) mod P]
    if (rc == ZAFB_OK && p->log2n < 2 && n >= 16 && n <= 8192) {
        p->ldk = (n + 3) & ~int64_t(3);
        std::vector<double> mat(size_t(n) * p->ldk, 0.0);
        for (int64_t k = 0; k < n; ++k) {
            const int64_t b = p->b0 + int64_t(p->db) * k;
            for (int64_t m = 0; m < n; ++m) {
                const int64_t a = p->a0 + int64_t(p->da) * m;
                mat[k * p->ldk + m] = sout[k] * sin_[m] * tab[(a * b) % p->period];
            }
        }
        std::vector<float> hi(mat.size()), lo(mat.size());
        split_tf32_host(mat.data(), mat.size(), hi.data(), lo.data());
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p->d_mat_hi), hi.size() * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&p->d_mat_lo), lo.size() * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpy(p->d_mat_hi, hi.data(), hi.size() * sizeof(float), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(p->d_mat_lo, lo.data(), lo.size() * sizeof(float), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) rc = fail(ZAFB_E_CUDA, "dct: uploading the transform matrix failed: %s", cudaGetErrorString(e));
        // even / odd form: only where the symmetry holds to float64 rounding (it does for types I and II of both kinds)
        bool sym = type <= 2;
        double worst = 0.0;
        for (int64_t k = 0; k < n && sym; ++k)
            for (int64_t m = 0; m < n / 2; ++m) {
                const double a = mat[k * p->ldk + m], b = mat[k * p->ldk + (n - 1 - m)];
                worst = std::max(worst, std::fabs(b - ((k & 1) ? -a : a)));
            }
        if (worst > 1e-12) sym = false;
        if (rc == ZAFB_OK) {  // dense form on the CTA-pair kernel: 256-row blocks of Mat
            std::vector<GemmTile> tiles;
            for (int64_t t = 0; t * 256 < n; ++t) tiles.push_back(GemmTile{0, int(n), int(t * 256), 1, int(std::min<int64_t>(256, n - t * 256))});
            p->mat2_tiles = int(tiles.size());
            e = cudaMalloc(reinterpret_cast<void**>(&p->d_mat2_tiles), tiles.size() * sizeof(GemmTile));
            if (e == cudaSuccess) e = cudaMemcpy(p->d_mat2_tiles, tiles.data(), tiles.size() * sizeof(GemmTile), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) rc = fail(ZAFB_E_CUDA, "dct: uploading the tile table failed: %s", cudaGetErrorString(e));
        }
        if (rc == ZAFB_OK && sym) {
            const int64_t nh = (n + 1) / 2, nd = n / 2;        // columns of s (with the middle sample of an odd n) and of d
            const int64_t ne = (n + 1) / 2, no = n / 2;        // even and odd outputs
            p->eo_ldb = (nh + 3) & ~int64_t(3);
            p->eo_dcol = p->eo_ldb;
            p->eo_lda = p->eo_ldb + ((nd + 3) & ~int64_t(3));
            // stacked operand [Mat_even ; Mat_odd] in blocks of `blk` rows + one tile descriptor per block
            auto build = [&](int64_t blk, float** d_hi, float** d_lo, GemmTile** d_tiles, int* n_tiles) -> cudaError_t {
                const int64_t te = (ne + blk - 1) / blk, to = (no + blk - 1) / blk;
                *n_tiles = int(te + to);
                std::vector<double> st(size_t(te + to) * blk * p->eo_ldb, 0.0);
                for (int64_t k = 0; k < n; ++k) {
                    const int64_t row = (k & 1) ? te * blk + k / 2 : k / 2;
                    const int64_t cols = (k & 1) ? nd : nh;
                    for (int64_t m = 0; m < cols; ++m) st[row * p->eo_ldb + m] = mat[k * p->ldk + m];
                }
                std::vector<GemmTile> tiles;
                for (int64_t t = 0; t < te; ++t)
                    tiles.push_back(GemmTile{0, int(nh), int(2 * t * blk), 2, int(std::min<int64_t>(blk, ne - t * blk)), 1});
                for (int64_t t = 0; t < to; ++t)
                    tiles.push_back(GemmTile{int(p->eo_dcol), int(nd), int(2 * t * blk + 1), 2, int(std::min<int64_t>(blk, no - t * blk)), 2});
                std::vector<float> ehi(st.size()), elo(st.size());
                split_tf32_host(st.data(), st.size(), ehi.data(), elo.data());
                cudaError_t e2 = cudaMalloc(reinterpret_cast<void**>(d_hi), ehi.size() * sizeof(float));
                if (e2 == cudaSuccess) e2 = cudaMalloc(reinterpret_cast<void**>(d_lo), elo.size() * sizeof(float));
                if (e2 == cudaSuccess) e2 = cudaMalloc(reinterpret_cast<void**>(d_tiles), tiles.size() * sizeof(GemmTile));
                if (e2 == cudaSuccess) e2 = cudaMemcpy(*d_hi, ehi.data(), ehi.size() * sizeof(float), cudaMemcpyHostToDevice);
                if (e2 == cudaSuccess) e2 = cudaMemcpy(*d_lo, elo.data(), elo.size() * sizeof(float), cudaMemcpyHostToDevice);
                if (e2 == cudaSuccess) e2 = cudaMemcpy(*d_tiles, tiles.data(), tiles.size() * sizeof(GemmTile), cudaMemcpyHostToDevice);
                return e2;
            };
            e = build(128, &p->d_eo_hi, &p->d_eo_lo, &p->d_eo_tiles, &p->eo_tiles);
            if (e == cudaSuccess) e = build(256, &p->d_eo2_hi, &p->d_eo2_lo, &p->d_eo2_tiles, &p->eo2_tiles);
            if (e != cudaSuccess) rc = fail(ZAFB_E_CUDA, "dct: uploading the even/odd transform matrix failed: %s", cudaGetErrorString(e));
        }
    }
    if (rc != ZAFB_OK) {
        zafb_dct_plan_destroy(p);
        return rc;
    }
    *out = p;
    return ZAFB_OK;
}

int zafb_dct_plan_destroy(zafb_dct_plan* p) {
    if (!p) return ZAFB_OK;
    cudaFree(p->d_tab);
    cudaFree(p->d_sin);
    cudaFree(p->d_sout);
    cudaFree(p->d_tw_fft);
    cudaFree(p->d_tw_a);
    cudaFree(p->d_tw_b);
    cudaFree(p->d_mat_hi);
    cudaFree(p->d_mat_lo);
    cudaFree(p->d_eo_hi);
    cudaFree(p->d_eo_lo);
    cudaFree(p->d_eo_tiles);
    cudaFree(p->d_eo2_hi);
    cudaFree(p->d_eo2_lo);
    cudaFree(p->d_eo2_tiles);
    cudaFree(p->d_mat2_tiles);
    cudaFree(p->d_tw_4step);
    delete p;
    return ZAFB_OK;
}

// test hook: 1 = always use the direct kernel, 2 = require the tensor-core matrix path
int zafb_dct_plan_force_direct(zafb_dct_plan* p, int on) {
    ZAFB_REQUIRE(p != nullptr, "plan is NULL");
    p->force_direct = on;
    return ZAFB_OK;
}

int zafb_dct_f32(const zafb_dct_plan* p, const float* x, int64_t batch, int64_t stride, float* out, int64_t out_stride,
                 void* stream) {
    ZAFB_REQUIRE(p != nullptr, "plan is NULL");
    ZAFB_REQUIRE(batch >= 0 && stride >= p->n && out_stride >= p->n, "bad batch geometry");
    if (batch == 0) return ZAFB_OK;
    ZAFB_REQUIRE(x != nullptr && out != nullptr, "x/out is NULL");
    int rc = set_kernel_attrs();
    if (rc != ZAFB_OK) return rc;
    const int n = int(p->n);
    const int64_t grid = batch < int64_t(sm_count()) * 16 ? batch : int64_t(sm_count()) * 16;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool mat_ok = p->d_mat_hi != nullptr && reinterpret_cast<uintptr_t>(out) % 4 == 0;
    if (p->force_direct == 2 && !mat_ok)
        return fail(ZAFB_E_UNSUPPORTED, "dct: no tensor-core matrix path for this plan (power-of-two types II-IV use the FFT)");
    if (mat_ok && (p->force_direct == 2 || (p->force_direct == 0 && batch >= 8))) {
        // split the input into TF32 halves in a stream-ordered scratch buffer, then one 3xTF32 GEMM
        static bool pool_ready = false;
        if (!pool_ready) {  // keep the stream-ordered scratch in the pool between calls instead of returning it to the OS
            int dev = 0;
            cudaMemPool_t pool;
            uint64_t keep = UINT64_MAX;
            if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            pool_ready = true;
        }
        float* ws = nullptr;
        const bool pair = env_flag("ZAFB_GEMM_PAIR", 1) != 0 && batch >= 256;  // CTA-pair kernel (256 x 256 tiles)
        if (p->d_eo_tiles != nullptr && !env_flag("ZAFB_DCT_DENSE", 0) && pair && n % 8 == 0 && stride % 4 == 0 &&
            reinterpret_cast<uintptr_t>(x) % 16 == 0 && env_flag("ZAFB_GEMM_FUSED_FOLD", 1)) {
            // even / odd form with the fold and the TF32 split inside the GEMM kernel: the input is read as it is
            return gemm3xtf32_pair_fold(x, stride, n, p->d_eo2_hi, p->d_eo2_lo, p->eo_ldb, int64_t(p->eo2_tiles) * 256, p->eo2_tiles,
                                        p->d_eo2_tiles, out, out_stride, batch, st);
        }
        if (p->d_eo_tiles != nullptr && !env_flag("ZAFB_DCT_DENSE", 0)) {
            // even / odd form: fold the input into [s | d] while splitting it, then one tiled product with half the work
            const size_t half = size_t(batch) * size_t(p->eo_lda);
            ZAFB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ws), 2 * half * sizeof(float), st));
            int64_t blocks = ceil_div(int64_t(half), 256);
            if (blocks > int64_t(sm_count()) * 16) blocks = int64_t(sm_count()) * 16;
            if (n % 8 == 0 && stride % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 && p->eo_lda == n && p->eo_dcol == n / 2) {
                const int groups = n / 8, threads = groups >= 256 ? 256 : (256 / groups) * groups;
                int64_t vb = ceil_div(batch, int64_t(threads / groups > 0 ? threads / groups : 1));
                if (vb > int64_t(sm_count()) * 16) vb = int64_t(sm_count()) * 16;
                fold_split_tf32_vec_kernel<<<unsigned(vb), threads, 0, st>>>(x, batch, n, stride, ws, ws + half);
            } else {
                fold_split_tf32_kernel<<<unsigned(blocks), 256, 0, st>>>(x, batch, n, stride, ws, ws + half, p->eo_lda, p->eo_dcol);
            }
            ZAFB_LAUNCH_CHECK();
            if (pair)
                rc = gemm3xtf32_pair_tiled(ws, ws + half, p->eo_lda, p->eo_lda, p->d_eo2_hi, p->d_eo2_lo, p->eo_ldb,
                                           int64_t(p->eo2_tiles) * 256, p->eo2_tiles, p->d_eo2_tiles, out, out_stride, batch, st);
            else
                rc = gemm3xtf32_tiled(128, ws, ws + half, p->eo_lda, p->eo_lda, p->d_eo_hi, p->d_eo_lo, p->eo_ldb, p->eo_tiles,
                                      p->d_eo_tiles, out, out_stride, batch, st);
            cudaFreeAsync(ws, st);
            return rc;
        }
        const size_t half = size_t(batch) * size_t(p->ldk);
        ZAFB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ws), 2 * half * sizeof(float), st));
        rc = split_tf32(x, batch, n, stride, ws, ws + half, p->ldk, st);
        if (rc == ZAFB_OK && pair)
            rc = gemm3xtf32_pair_tiled(ws, ws + half, p->ldk, n, p->d_mat_hi, p->d_mat_lo, p->ldk, n, p->mat2_tiles, p->d_mat2_tiles,
                                       out, out_stride, batch, st);
        else if (rc == ZAFB_OK)
            rc = gemm3xtf32(ws, ws + half, p->ldk, p->d_mat_hi, p->d_mat_lo, p->ldk, out, out_stride, batch, n, n, st);
        cudaFreeAsync(ws, st);
        return rc;
    }
    {
        const bool aligned = reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0 &&
                             stride % 4 == 0 && out_stride % 4 == 0;
        const bool warp_ok = p->d_tw_4step != nullptr && p->type >= 2 && aligned;
        if (p->force_direct == 4 && !warp_ok)
            return fail(ZAFB_E_UNSUPPORTED, "dct warp kernel needs N = 512, 1024, 2048 or 4096, type 2..4, 16-byte aligned rows");
        if (warp_ok && (p->force_direct == 0 || p->force_direct == 4)) {
            const size_t smem = size_t(n / 2 + kDctWarps * (n / 64) * kFft1024Pitch) * sizeof(float2);
            int64_t ctas = ceil_div(batch, kDctWarps);
            const int occ = n >= 2048 ? 1 : (p->type == 3 ? 2 : 3);
            if (ctas > int64_t(sm_count()) * occ) ctas = int64_t(sm_count()) * occ;
            const unsigned g = unsigned(ctas), b = kDctWarps * 32;
#define ZAFB_DCT_WARP_N(NN, MODE, DST)                                                                            \
    dct_warp_kernel<NN, MODE, DST><<<g, b, smem, st>>>(x, batch, stride, p->d_tw_4step, p->d_tw_a, p->d_tw_b, out, out_stride)
#define ZAFB_DCT_WARP(MODE, DST)                                                                                  \
    do {                                                                                                          \
        if (n == 512) ZAFB_DCT_WARP_N(512, MODE, DST);                                                            \
        else if (n == 1024) ZAFB_DCT_WARP_N(1024, MODE, DST);                                                     \
        else if (n == 2048) ZAFB_DCT_WARP_N(2048, MODE, DST);                                                     \
        else ZAFB_DCT_WARP_N(4096, MODE, DST);                                                                    \
    } while (0)
            if (p->type == 2) { if (p->kind) ZAFB_DCT_WARP(2, true); else ZAFB_DCT_WARP(2, false); }
            else if (p->type == 3) { if (p->kind) ZAFB_DCT_WARP(3, true); else ZAFB_DCT_WARP(3, false); }
            else { if (p->kind) ZAFB_DCT_WARP(4, true); else ZAFB_DCT_WARP(4, false); }
#undef ZAFB_DCT_WARP_N
#undef ZAFB_DCT_WARP
            ZAFB_LAUNCH_CHECK();
            return ZAFB_OK;
        }
    }
    if (p->log2n >= 2 && (p->force_direct == 0 || p->force_direct == 3)) {
        int flags = 0;
        if (p->kind == 1) flags = (p->type == 2) ? (2 | 4) : (1 | 8);
        const size_t smem = size_t(n) * sizeof(float2) + size_t(n) * sizeof(float);
        int th = n / 8;
        if (th < 32) th = 32;
        if (th > 256) th = 256;
        dct_fft_kernel<<<unsigned(grid), th, smem, st>>>(x, batch, stride, p->log2n, p->type, flags, p->d_tw_fft, p->d_tw_a,
                                                         p->d_tw_b, out, out_stride);
    } else {
        const size_t smem = size_t(n) * sizeof(float) + 16;
        int th = n < 256 ? ((n + 31) / 32) * 32 : 256;
        dct_direct_kernel<<<unsigned(grid), th, smem, st>>>(x, batch, stride, n, p->period, p->a0, p->da, p->b0, p->db, p->d_tab,
                                                            p->d_sin, p->d_sout, out, out_stride);
    }
    ZAFB_LAUNCH_CHECK();
    return ZAFB_OK;
}

int zafb_dct_host_f32(const zafb_dct_plan* p, const float* x, int64_t batch, int64_t stride, float* out,
                      int64_t out_stride) {
    ZAFB_REQUIRE(p != nullptr, "plan is NULL");
    ZAFB_REQUIRE(batch >= 0 && stride >= p->n && out_stride >= p->n, "bad batch geometry");
    if (batch == 0) return ZAFB_OK;
    ZAFB_REQUIRE(x != nullptr && out != nullptr, "x/out is NULL");
    const int64_t dpitch = (p->n + 3) & ~int64_t(3);  // 16-byte aligned rows on the device
    return run_host_pipeline(x, size_t(stride) * sizeof(float), size_t(p->n) * sizeof(float), size_t(dpitch) * sizeof(float), out,
                             size_t(out_stride) * sizeof(float), size_t(p->n) * sizeof(float), size_t(dpitch) * sizeof(float),
                             batch, [&](void* d_in, void* d_out, int64_t, int64_t nb, cudaStream_t st) {
                                 return zafb_dct_f32(p, static_cast<const float*>(d_in), nb, dpitch,
                                                     static_cast<float*>(d_out), dpitch, st);
                             });
}

}  // extern "C"
