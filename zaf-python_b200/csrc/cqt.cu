// cqtspectrogram / cqtchromagram (zaf.py:562-635, 638-700).
//
// Per frame the reference computes abs(K * fft(x[i : i + L])) with K the sparse (n_freqs, L) CQT kernel.
// One CTA per frame:
//   1. the L real samples (zero-padded by predication, zaf.py:612-620) are loaded as L/2 complex points into
//      shared memory (128 KB at L = 32 768);
//   2. in-place radix-4 (+ one radix-2) decimation-in-frequency FFT -- in place because two 128 KB ping-pong
//      buffers do not fit; the result is left in digit-reversed order and read through pos();
//   3. each kernel row is one contiguous band (true for every kernel zaf.cqtkernel builds; gaps inside a band
//      are packed as zeros): one warp per row walks its band, unpacking the real-input spectrum bin on the fly
//      (X[c] = E + W_L^c O) and accumulating K[r, c] * X[c]; columns above L/2 use X[c] = conj(X[L - c]);
//   4. magnitude, optional chroma fold (rows i :: octave_resolution summed), store.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "fft_core.cuh"
#include "gemm_tc.cuh"
#include "host_pipe.cuh"

using namespace zafb;

struct zafb_cqt_plan {
    int64_t n_freqs = 0, fft_length = 0, step = 0;
    int log2m = 0;                 // M = fft_length / 2 complex points
    float2* d_tw_fft = nullptr;    // W_M^t, t < M
    float2* d_tw_full = nullptr;   // W_L^t, t <= M
    int* d_band_lo = nullptr;
    int* d_band_len = nullptr;
    int* d_band_off = nullptr;
    float2* d_weights = nullptr;   // packed complex bands
    int64_t packed = 0;
    // L = 32768 register-FFT kernel
    float2* d_t1 = nullptr;        // W_16384^{il q} at [q * 32 + il], il < 32, q < 32
    float2* d_t2 = nullptr;        // W_512^{ih q}  at [q * 16 + ih], ih < 16, q < 32
    int pair_lo = 0, pair_hi = -1; // range of min(k, M - k) over every column any band touches
    bool real_weights = false;     // every imaginary part is below fp32 resolution of the row's real parts
    float* d_weights_re = nullptr; // real parts only (same packing as d_weights)
    int* d_sched = nullptr;        // rows sorted by decreasing band length, dealt to the 16 warps longest-first
    int* d_sched_cnt = nullptr;    // rows per warp
    int sched_stride = 0;
    int force_kernel = 0;          // 0 auto, 1 generic, 2 register-FFT kernel, 3 even/odd kernel (tests)
    // even/odd kernel (L = 32768): W_8192^{tl k1} at [k1 * 32 + tl], W_256^{th k1} at [k1 * 8 + th], W_256^{t0 k2} at [k2 * 16 + t0]
    float2* d_eo_t1 = nullptr;
    float2* d_eo_t2 = nullptr;
    float2* d_eo_t3 = nullptr;
    int* d_eo_seg_xoff = nullptr;  // per thread (256): offset of its band segment in the unpacked-bin array
    float* d_eo_seg_w = nullptr;   // segment weights transposed, [i][thread], i < eo_seg rounded up to 8, zero-padded
    int2* d_eo_row_seg = nullptr;  // per row: (first segment, number of segments)
    int eo_seg = 0;                // segment length (odd), 0 = the kernel cannot serve this operator
    // tensor-core route: the kernel as a dense real (n_freqs x kp) operand over the columns [col_lo, col_hi] the bands
    // touch, TF32 hi/lo halves; the spectrum's real and imaginary parts are two rows of the other operand
    int route = 0;
    int col_lo = 0, col_hi = -1;
    int64_t kp = 0;                // col_hi - col_lo + 1 rounded up to 4
    float* d_kern_hi = nullptr;
    float* d_kern_lo = nullptr;
    // the same operand PACKED: groups of 16 consecutive rows, each stored dense over the union of its rows' bands only
    // (62 080 elements instead of 269 976 at cfg 5); group g multiplies spectrum columns
    // [pk_col0[g], pk_col0[g] + pk_k[g]) and produces operator rows [16 g, 16 g + 16)
    std::vector<int> pk_col0, pk_k;
    int pk_ld = 0;                 // common row pitch of the packed groups (the longest union span, rounded up to 4)
    float* d_pk_hi = nullptr;      // [groups * 16][pk_ld]
    float* d_pk_lo = nullptr;
    GemmTile* d_pk_tiles = nullptr;
    int64_t pk_elems = 0;          // elements inside the union spans (the packed operand proper)
};

namespace {

constexpr int kMaxDynSmem = 220 * 1024;
constexpr int kThreads = 1024;

// position of frequency k after the in-place DIF passes (radix 4 ... 4 [2])
__device__ __forceinline__ int dif_pos(int k, int log2m) {
    int pos = 0;
    int rem = log2m;
    while (rem >= 2) {
        rem -= 2;
        pos += (k & 3) << rem;
        k >>= 2;
    }
    if (rem == 1) pos += (k & 1);
    return pos;
}

__device__ __forceinline__ void fft_inplace_dif(float2* z, const float2* __restrict__ tw, int log2m, int tid, int nth) {
    const int M = 1 << log2m;
    int span_log = log2m;
    while (span_log >= 2) {
        const int quarter = 1 << (span_log - 2);
        const int tshift = log2m - span_log;  // W_S^t = W_M^{t << tshift}
        for (int t = tid; t < (M >> 2); t += nth) {
            const int i = t & (quarter - 1);
            const int base = ((t - i) << 2) + i;
            float2 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = z[base + q * quarter];
            fft_reg<4>(v);  // bit-reversed: X[q] = v[bitrev(q, 2)]
            z[base] = v[0];
            if (i == 0) {
                z[base + quarter] = v[2];
                z[base + 2 * quarter] = v[1];
                z[base + 3 * quarter] = v[3];
            } else {
                z[base + quarter] = cmul(v[2], tw[(i) << tshift]);
                z[base + 2 * quarter] = cmul(v[1], tw[(2 * i) << tshift]);
                z[base + 3 * quarter] = cmul(v[3], tw[(3 * i) << tshift]);
            }
        }
        __syncthreads();
        span_log -= 2;
    }
    if (span_log == 1) {
        for (int t = tid; t < (M >> 1); t += nth) {
            const float2 p = z[2 * t], q = z[2 * t + 1];
            z[2 * t] = cadd(p, q);
            z[2 * t + 1] = csub(p, q);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads, 1)
cqt_frame_kernel(const float* __restrict__ x, int64_t ns, int64_t clip_stride, int64_t nt, int64_t step, int64_t front,
                 int log2m, const float2* __restrict__ tw_fft, const float2* __restrict__ tw_full,
                 const int* __restrict__ band_lo, const int* __restrict__ band_len, const int* __restrict__ band_off,
                 const float2* __restrict__ weights, int n_freqs, int octave, float* __restrict__ out, int layout,
                 int64_t total_frames) {
    extern __shared__ float2 smem2[];
    const int M = 1 << log2m;
    const int L = M << 1;
    float2* z = smem2;
    float* q = reinterpret_cast<float*>(smem2 + M);  // n_freqs magnitudes
    const int tid = threadIdx.x, nth = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nth >> 5;
    for (int64_t f = blockIdx.x; f < total_frames; f += gridDim.x) {
        const int64_t clip = f / nt, j = f - clip * nt;
        const int64_t start = j * step - front;
        const float* xc = x + clip * clip_stride;
        for (int i = tid; i < M; i += nth) {
            const int64_t s = start + 2 * i;
            const float x0 = (s >= 0 && s < ns) ? __ldg(xc + s) : 0.f;
            const float x1 = (s + 1 >= 0 && s + 1 < ns) ? __ldg(xc + s + 1) : 0.f;
            z[i] = make_float2(x0, x1);
        }
        __syncthreads();
        fft_inplace_dif(z, tw_fft, log2m, tid, nth);
        const float2 z0 = z[0];
        for (int r = warp; r < n_freqs; r += nwarps) {
            const int lo = band_lo[r], len = band_len[r];
            const float2* w = weights + band_off[r];
            float ar = 0.f, ai = 0.f;
            for (int c = lane; c < len; c += 32) {
                int col = lo + c;
                const bool mirror = col > M;
                const int k = mirror ? L - col : col;
                float2 X;
                if (k == M) {
                    X = make_float2(z0.x - z0.y, 0.f);
                } else if (k == 0) {
                    X = make_float2(z0.x + z0.y, 0.f);
                } else {
                    const float2 zk = z[dif_pos(k, log2m)];
                    const float2 zp = z[dif_pos(M - k, log2m)];
                    const float2 e = make_float2(0.5f * (zk.x + zp.x), 0.5f * (zk.y - zp.y));
                    const float2 od = make_float2(0.5f * (zk.y + zp.y), 0.5f * (zp.x - zk.x));
                    X = cadd(e, cmul(__ldg(tw_full + k), od));
                }
                if (mirror) X.y = -X.y;
                const float2 kv = __ldg(w + c);
                ar += kv.x * X.x - kv.y * X.y;
                ai += kv.x * X.y + kv.y * X.x;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ar += __shfl_xor_sync(0xffffffffu, ar, o);
                ai += __shfl_xor_sync(0xffffffffu, ai, o);
            }
            if (lane == 0) q[r] = sqrtf(ar * ar + ai * ai);
        }
        __syncthreads();
        if (octave > 0) {
            for (int i = tid; i < octave; i += nth) {
                float acc = 0.f;
                for (int r = i; r < n_freqs; r += octave) acc += q[r];  // zaf.py:693-698
                if (layout == ZAFB_LAYOUT_FRAME_MAJOR) out[f * octave + i] = acc;
                else out[(clip * octave + i) * nt + j] = acc;
            }
        } else {
            for (int r = tid; r < n_freqs; r += nth) {
                if (layout == ZAFB_LAYOUT_FRAME_MAJOR) out[f * n_freqs + r] = q[r];
                else out[(clip * int64_t(n_freqs) + r) * nt + j] = q[r];
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// L = 32768 (M = 16384 complex points; BASELINE cfg 5 and the reference's example kernel):
// one CTA of 512 threads per frame, the FFT in THREE in-place decimation-in-time passes
// 16384 = 16 x 32 x 32 with the radix-16 / radix-32 butterflies held in registers (fft_reg),
// instead of seven radix-4 passes through shared memory.  With n = a + 32 b + 1024 c and
// k = kc + 16 kb + 512 ka:
//   pass 1  thread (a, b) loads x[a + 32 b + 1024 c], c < 16, straight from global memory
//           (coalesced 8-byte loads), FFT-16 over c -> S1[a][b][kc]           at 512 a + 16 b  + kc
//   pass 2  thread (a, kc): S1[a][b][kc] W_512^{b kc}, FFT-32 over b -> S2     at 512 a + 16 kb + kc
//   pass 3  thread j = 16 kb + kc: S2[a][j] W_M^{a j}, FFT-32 over a -> X[k]   at 512 ka + j = k
// so the spectrum ends in NATURAL order.  Every address has its low four bits XOR-ed with bits
// 4..7 and 9..12 (swz), which makes all three passes and the consumers bank-conflict free.
// Then the real-input split is applied IN PLACE to the bins some band needs (pairs k, M - k),
// and one warp per kernel row accumulates its band.
// ------------------------------------------------------------------------------------------
constexpr int kRegThreads = 512;
constexpr int kRegM = 16384;

__device__ __forceinline__ int swz(int a) { return a ^ (((a >> 4) ^ (a >> 9)) & 15); }

// EXPORT = true (tensor-core route): the kernel stops after the real-input split and writes Re X[k] and Im X[k],
// k in [col_lo, col_hi], as rows 2 f and 2 f + 1 of the TF32 hi / lo matrices a_hi / a_lo (row pitch kp).
template <bool EXPORT>
__global__ void __launch_bounds__(kRegThreads, 1)
cqt32768_kernel(const float* __restrict__ x, int64_t ns, int64_t clip_stride, int64_t nt, int64_t step, int64_t front,
                const float2* __restrict__ t1, const float2* __restrict__ t2, const float2* __restrict__ tw_full,
                const int* __restrict__ band_lo, const int* __restrict__ band_len, const int* __restrict__ band_off,
                const float2* __restrict__ weights, const float* __restrict__ weights_re, int packed, int smem_weights,
                int smem_split_tw, const int* __restrict__ sched, const int* __restrict__ sched_cnt, int sched_stride,
                int n_freqs, int octave, int pair_lo, int pair_hi, float* __restrict__ out, int layout,
                int64_t total_frames, float* __restrict__ a_hi, float* __restrict__ a_lo, int col_lo, int col_hi, int64_t kp) {
    extern __shared__ float2 smem2[];
    constexpr int M = kRegM, L = 2 * kRegM;
    float2* z = smem2;
    float2* s_t1 = smem2 + M;        // 1024: W_M^{il r} at [r * 32 + il]
    float2* s_t2 = s_t1 + 1024;      // 512:  W_512^{ih r} at [r * 16 + ih]
    float* q = reinterpret_cast<float*>(s_t2 + 512);  // n_freqs magnitudes (rounded up to an even count)
    float2* s_split = reinterpret_cast<float2*>(q + ((n_freqs + 1) & ~1));  // W_L^k, k in [pair_lo, pair_hi]
    float* s_w = reinterpret_cast<float*>(s_split + (smem_split_tw ? pair_hi - pair_lo + 1 : 0));  // real band weights
    __shared__ float s_nyq;          // X[M]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 1024; i += kRegThreads) s_t1[i] = t1[i];
    for (int i = tid; i < 512; i += kRegThreads) s_t2[i] = t2[i];
    if (smem_split_tw)
        for (int i = tid; i <= pair_hi - pair_lo; i += kRegThreads) s_split[i] = tw_full[pair_lo + i];
    if (smem_weights)
        for (int i = tid; i < packed; i += kRegThreads) s_w[i] = weights_re[i];
    const int my_rows = sched_cnt[warp];
    __syncthreads();
    const int a2 = tid >> 4, c2 = tid & 15;  // pass 2: a, kc

    for (int64_t f = blockIdx.x; f < total_frames; f += gridDim.x) {
        const int64_t clip = f / nt, j = f - clip * nt;
        const int64_t start = j * step - front;
        const float* xc = x + clip * clip_stride;
        const bool fast = start >= 0 && start + L <= ns && ((reinterpret_cast<uintptr_t>(xc + start) & 7) == 0);
        // ---- pass 1: two (a, b) items per thread, t' = a + 32 b = tid + 512 h
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int tp = tid + 512 * h;
            float2 u[16];
            if (fast) {
                const float2* fp = reinterpret_cast<const float2*>(xc + start) + tp;
#pragma unroll
                for (int c = 0; c < 16; ++c) u[c] = __ldg(fp + 1024 * c);
            } else {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const int64_t s0 = start + 2 * (tp + 1024 * c);
                    u[c].x = (s0 >= 0 && s0 < ns) ? __ldg(xc + s0) : 0.f;
                    u[c].y = (s0 + 1 >= 0 && s0 + 1 < ns) ? __ldg(xc + s0 + 1) : 0.f;
                }
            }
            fft_reg<16>(u);
            const int a = tp & 31, b = tp >> 5;
            float2* zb = z + 512 * a + 16 * b;
            const int sw = (a ^ b) & 15;
            static_for<0, 16>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                zb[k ^ sw] = u[bitrev(k, 4)];
            });
        }
        __syncthreads();
        float2 v[32];
        // ---- pass 2: thread (a, kc), FFT-32 over b
        {
            float2* zb = z + 512 * a2;
            const int sw = c2 ^ (a2 & 15);
            static_for<0, 32>([&](auto rc) {
                constexpr int r = decltype(rc)::value;
                const float2 y = zb[16 * r + (sw ^ (r & 15))];
                if constexpr (r > 0) v[r] = cmul(y, s_t2[r * 16 + c2]);
                else v[r] = y;
            });
            fft_reg<32>(v);
            static_for<0, 32>([&](auto kc) {
                constexpr int kb = decltype(kc)::value;
                zb[16 * kb + (sw ^ (kb & 15))] = v[bitrev(kb, 5)];
            });
        }
        __syncthreads();
        // ---- pass 3: thread j = tid, FFT-32 over a; twiddle W_M^{a j} = W_M^{a lane} W_512^{a warp}
        {
            const int sw = (tid ^ (tid >> 4)) & 15;   // kc ^ (kb & 15)
            const int hi = tid & ~15;
            static_for<0, 32>([&](auto rc) {
                constexpr int r = decltype(rc)::value;
                const float2 y = z[512 * r + hi + (sw ^ (r & 15))];
                if constexpr (r > 0) v[r] = cmul(cmul(y, s_t1[r * 32 + lane]), s_t2[r * 16 + warp]);
                else v[r] = y;
            });
            fft_reg<32>(v);
            static_for<0, 32>([&](auto kc) {
                constexpr int ka = decltype(kc)::value;
                z[512 * ka + hi + (sw ^ (ka & 15))] = v[bitrev(ka, 5)];
            });
        }
        __syncthreads();
        // ---- real-input split, in place, for the pairs (k, M - k) some band needs
        for (int k = pair_lo + tid; k <= pair_hi; k += kRegThreads) {
            if (k == 0) {
                const float2 z0 = z[0];
                z[0] = make_float2(z0.x + z0.y, 0.f);
                s_nyq = z0.x - z0.y;
            } else {
                const int pk = swz(k), pm = swz(M - k);
                const float2 zk = z[pk], zp = z[pm];
                const float2 e = make_float2(0.5f * (zk.x + zp.x), 0.5f * (zk.y - zp.y));
                const float2 od = make_float2(0.5f * (zk.y + zp.y), 0.5f * (zp.x - zk.x));
                const float2 t = cmul(smem_split_tw ? s_split[k - pair_lo] : __ldg(tw_full + k), od);
                z[pk] = cadd(e, t);                                       // X[k]
                if (pm != pk) z[pm] = make_float2(e.x - t.x, t.y - e.y);  // X[M - k] = conj(E - W O)
            }
        }
        __syncthreads();
        if constexpr (EXPORT) {
            float* rh = a_hi + 2 * f * kp;
            float* rl = a_lo + 2 * f * kp;
            for (int c = tid; c < int(kp); c += kRegThreads) {
                const int k = col_lo + c;
                float2 X = make_float2(0.f, 0.f);  // padding columns beyond col_hi stay zero
                if (k <= col_hi) X = (k == M) ? make_float2(s_nyq, 0.f) : z[swz(k)];
                uint32_t h, l;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(X.x));
                float hv = __uint_as_float(h);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(X.x - hv));
                rh[c] = hv;
                rl[c] = __uint_as_float(l);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(X.y));
                hv = __uint_as_float(h);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(X.y - hv));
                rh[kp + c] = hv;
                rl[kp + c] = __uint_as_float(l);
            }
            __syncthreads();  // z is rewritten by the next frame's pass 1
            continue;
        }
        // ---- banded kernel rows, one warp per row; rows are dealt to the warps longest-first at plan time
        for (int it = 0; it < my_rows; ++it) {
            const int r = sched[warp * sched_stride + it];
            const int lo = band_lo[r], len = band_len[r];
            float ar = 0.f, ai = 0.f;
            if (smem_weights && lo + len <= M) {  // real weights in shared memory, no mirrored columns
                const float* w = s_w + band_off[r];
                float br = 0.f, bi = 0.f;
                int c = lane;
                for (; c + 32 < len; c += 64) {
                    const float2 X0 = z[swz(lo + c)], X1 = z[swz(lo + c + 32)];
                    const float w0 = w[c], w1 = w[c + 32];
                    ar = fmaf(w0, X0.x, ar);
                    ai = fmaf(w0, X0.y, ai);
                    br = fmaf(w1, X1.x, br);
                    bi = fmaf(w1, X1.y, bi);
                }
                if (c < len) {
                    const float2 X0 = z[swz(lo + c)];
                    const float w0 = w[c];
                    ar = fmaf(w0, X0.x, ar);
                    ai = fmaf(w0, X0.y, ai);
                }
                ar += br;
                ai += bi;
            } else {
                const float2* w = weights + band_off[r];
                for (int c = lane; c < len; c += 32) {
                    const int col = lo + c;
                    const bool mirror = col > M;
                    const int k = mirror ? L - col : col;
                    float2 X = (k == M) ? make_float2(s_nyq, 0.f) : z[swz(k)];
                    if (mirror) X.y = -X.y;
                    const float2 kv = __ldg(w + c);
                    ar += kv.x * X.x - kv.y * X.y;
                    ai += kv.x * X.y + kv.y * X.x;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ar += __shfl_xor_sync(0xffffffffu, ar, o);
                ai += __shfl_xor_sync(0xffffffffu, ai, o);
            }
            if (lane == 0) q[r] = sqrtf(ar * ar + ai * ai);
        }
        __syncthreads();
        if (octave > 0) {
            for (int i = tid; i < octave; i += kRegThreads) {
                float acc = 0.f;
                for (int r = i; r < n_freqs; r += octave) acc += q[r];  // zaf.py:693-698
                if (layout == ZAFB_LAYOUT_FRAME_MAJOR) out[f * octave + i] = acc;
                else out[(clip * octave + i) * nt + j] = acc;
            }
        } else {
            for (int r = tid; r < n_freqs; r += kRegThreads) {
                if (layout == ZAFB_LAYOUT_FRAME_MAJOR) out[f * n_freqs + r] = q[r];
                else out[(clip * int64_t(n_freqs) + r) * nt + j] = q[r];
            }
        }
        // the next frame's pass 1 only writes z, which nobody reads any more; q[] is rewritten after four more barriers
    }
}

// ------------------------------------------------------------------------------------------
// L = 32768, real kernels whose bands stay below L/2 (every kernel zaf.cqtkernel builds): the even/odd kernel.
//
// Only the bins the bands touch are needed (22 ... 3235 of 32768 at BASELINE cfg 5), and they are all even or odd bins of
// the frame's spectrum X.  One decimation-in-frequency step on the REAL frame x[0 .. L) with H = L/2, Q = L/4 splits it
// into two complex transforms of Q = 8192 points -- half the shared memory of the 16384-point form, so TWO CTAs of 256
// threads fit on an SM and their barrier phases overlap:
//   even bins  X[2k'] = RFFT_H(s)[k'],  s[n] = x[n] + x[n+H]: the usual packed form z[m] = s[2m] + i s[2m+1],
//              Z = FFT_Q(z), X[2k'] = E + W_H^{k'} O with E, O from Z[k'] and conj(Z[Q-k']);
//   odd bins   X[2k'+1] = Y[k'] = sum_n d[n] W_L^{n(2k'+1)},  d[n] = x[n] - x[n+H]: an odd-frequency DFT of a real
//              sequence, which is ONE complex FFT_Q with no unpacking: v[n] = (d[n] - i d[n+Q]) W_L^n, V = FFT_Q(v),
//              Y[2q] = V[q], Y[2q+1] = conj(V[Q-1-q])          (validated in float64 against numpy.fft).
// FFT_Q, 8192 = 32 x 16 x 16, 32 points per thread and pass, butterflies in registers (fft_reg):
//   n = t + 256 c, t = t0 + 16 t1,  k = k1 + 32 k2 + 512 k3
//   pass 1  thread t: FFT-32 over c -> k1, times W_8192^{t k1}                         -> buf[k1][t]
//   pass 2  item (k1, t0): FFT-16 over t1 -> k2, times W_256^{t0 k2}                   -> buf[k1][t0 + 16 k2]   (in place)
//   pass 3  item (k1, k2): FFT-16 over t0 -> k3                                        -> buf[k1][16 k2 + k3]   (in place)
// Rows of buf have a pitch of 257 complex values: every access of every pass is base + constant and conflict-free (half
// warps run along t, t0 or k1).  The needed bins are unpacked into a small natural-order staging array that the band
// phase (one warp per kernel row) reads with unit stride; the band weights come from global memory (L1 / L2).
// ------------------------------------------------------------------------------------------
constexpr int kEoThreads = 256;
constexpr int kEoQ = 8192;
constexpr int kEoPitch = 257;
constexpr int kEoBuf = 32 * kEoPitch;  // float2
constexpr int kEoPad = 512;            // zeroed entries behind the unpacked bins (>= the longest segment, rounded up to 8)

__device__ __forceinline__ int eo_addr(int k) { return (k & 31) * kEoPitch + ((k >> 5) & 15) * 16 + (k >> 9); }

// passes 1 (store side), 2 and 3 of FFT_Q on v[c] = point t + 256 c
__device__ __forceinline__ void eo_fft8192(float2 (&v)[32], float2* __restrict__ buf, const float2* __restrict__ s_t1,
                                           const float2* __restrict__ s_t2, const float2* __restrict__ s_t3, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    fft_reg<32>(v);
    {
        float2* b = buf + tid;
        static_for<0, 32>([&](auto kc) {
            constexpr int k1 = decltype(kc)::value;
            float2 y = v[bitrev(k1, 5)];
            if constexpr (k1 > 0) y = cmul(cmul(y, s_t1[k1 * 32 + lane]), s_t2[k1 * 8 + warp]);
            b[k1 * kEoPitch] = y;
        });
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // pass 2: item (k1, t0); both items unrolled so that the second one's loads overlap the first FFT
        const int i = tid + kEoThreads * h;
        const int k1 = i >> 4, t0 = i & 15;
        float2* b = buf + k1 * kEoPitch + t0;
        float2 u[16];
#pragma unroll
        for (int t1 = 0; t1 < 16; ++t1) u[t1] = b[16 * t1];
        fft_reg<16>(u);
        const float2* tw = s_t3 + t0;
        static_for<0, 16>([&](auto kc) {
            constexpr int k2 = decltype(kc)::value;
            float2 y = u[bitrev(k2, 4)];
            if constexpr (k2 > 0) y = cmul(y, tw[k2 * 16]);
            b[16 * k2] = y;
        });
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // pass 3: item (k1, k2)
        const int j = tid + kEoThreads * h;
        float2* b = buf + (j & 31) * kEoPitch + 16 * (j >> 5);
        float2 u[16];
#pragma unroll
        for (int t0 = 0; t0 < 16; ++t0) u[t0] = b[t0];
        fft_reg<16>(u);
        static_for<0, 16>([&](auto kc) {
            constexpr int k3 = decltype(kc)::value;
            b[k3] = u[bitrev(k3, 4)];
        });
    }
    __syncthreads();
}

// EXPORT = true (tensor-core route): stops after the unpack and writes Re X[k] and Im X[k], k in [col_lo, col_hi], as rows
// 2 f and 2 f + 1 of the TF32 hi / lo matrices a_hi / a_lo (row pitch kp), like cqt32768_kernel<true>.
template <bool EXPORT>
__global__ void __launch_bounds__(kEoThreads, 2)
cqt_eo_kernel(const float* __restrict__ x, int64_t ns, int64_t clip_stride, int64_t nt, int64_t step, int64_t front,
              const float2* __restrict__ t1, const float2* __restrict__ t2, const float2* __restrict__ t3,
              const float2* __restrict__ tw_full, const int* __restrict__ band_lo, const int* __restrict__ band_len,
              const int* __restrict__ band_off, const float* __restrict__ weights_re, const int* __restrict__ seg_xoff,
              const float* __restrict__ seg_w, const int2* __restrict__ row_seg, int seg, int n_freqs, int octave,
              int col_lo, int col_hi,
              float* __restrict__ out, int layout, int64_t total_frames, float* __restrict__ a_hi, float* __restrict__ a_lo,
              int64_t kp) {
    extern __shared__ float2 smem2[];
    constexpr int L = 32768, H = L / 2, Q = kEoQ;
    float2* buf = smem2;                    // kEoBuf
    float2* s_t1 = buf + kEoBuf;            // 1024
    float2* s_t2 = s_t1 + 1024;             // 256
    float2* s_t3 = s_t2 + 256;              // 256
    float2* xs = s_t3 + 256;                // col_hi - col_lo + 1 unpacked bins, natural order
    // (kEoPad more entries behind the last bin: the zero-weight tail of a row's last segment reads them; they stay zero)
    float2* part = xs + (col_hi - col_lo + 1) + kEoPad;                // one partial sum per thread
    float* q = reinterpret_cast<float*>(part + kEoThreads);            // n_freqs magnitudes (chromagram fold)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kEoPad; i += kEoThreads) xs[(col_hi - col_lo + 1) + i] = make_float2(0.f, 0.f);
    for (int i = tid; i < 1024; i += kEoThreads) s_t1[i] = t1[i];
    if (tid < 256) {
        s_t2[tid] = t2[tid];
        s_t3[tid] = t3[tid];
    }
    const float2 w_t = tw_full[tid];  // W_L^t
    // bins k = 2 k' (even) and 2 k' + 1 (odd) inside [col_lo, col_hi]
    const int e_lo = (col_lo + 1) >> 1, e_hi = col_hi >> 1;
    const int o_lo = col_lo >> 1, o_hi = (col_hi - 1) >> 1;
    __syncthreads();

    for (int64_t f = blockIdx.x; f < total_frames; f += gridDim.x) {
        const int64_t clip = f / nt, j = f - clip * nt;
        const int64_t start = j * step - front;
        const float* xc = x + clip * clip_stride;
        const bool inside = start >= 0 && start + L <= ns;
        float2 v[32];
        // ---- even bins: z[m] = s[2m] + i s[2m+1], s[n] = x[n] + x[n + H]
        if (inside && ((reinterpret_cast<uintptr_t>(xc + start) & 7) == 0)) {
            const float2* p = reinterpret_cast<const float2*>(xc + start) + tid;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const float2 a = __ldg(p + 256 * c), b = __ldg(p + 256 * c + H / 2);
                v[c] = make_float2(a.x + b.x, a.y + b.y);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const int64_t s0 = start + 2 * (tid + 256 * c);
                auto ld = [&](int64_t s) { return (s >= 0 && s < ns) ? __ldg(xc + s) : 0.f; };
                v[c] = make_float2(ld(s0) + ld(s0 + H), ld(s0 + 1) + ld(s0 + 1 + H));
            }
        }
        eo_fft8192(v, buf, s_t1, s_t2, s_t3, tid);
        for (int k = e_lo + tid; k <= e_hi; k += kEoThreads) {  // X[2k] = E + W_H^k O
            float2 r;
            if (k == 0) {
                const float2 z0 = buf[0];
                r = make_float2(z0.x + z0.y, 0.f);
            } else {
                const float2 zk = buf[eo_addr(k)], zp = buf[eo_addr(Q - k)];
                const float2 e = make_float2(0.5f * (zk.x + zp.x), 0.5f * (zk.y - zp.y));
                const float2 od = make_float2(0.5f * (zk.y + zp.y), 0.5f * (zp.x - zk.x));
                r = cadd(e, cmul(__ldg(tw_full + 2 * k), od));
            }
            xs[2 * k - col_lo] = r;
        }
        // ---- odd bins: v[n] = (d[n] - i d[n + Q]) W_L^n, d[n] = x[n] - x[n + H]; W_L^n = W_L^t W_128^c
        if (inside) {
            const float* p = xc + start + tid;
            static_for<0, 32>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                const float d1 = __ldg(p + 256 * c) - __ldg(p + 256 * c + H);
                const float d2 = __ldg(p + 256 * c + Q) - __ldg(p + 256 * c + Q + H);
                const float2 u = make_float2(fmaf(d2, w_t.y, d1 * w_t.x), fmaf(-d2, w_t.x, d1 * w_t.y));
                v[c] = mul_tw<c, 128>(u);
            });
        } else {
            static_for<0, 32>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                const int64_t s0 = start + tid + 256 * c;
                auto ld = [&](int64_t s) { return (s >= 0 && s < ns) ? __ldg(xc + s) : 0.f; };
                const float d1 = ld(s0) - ld(s0 + H);
                const float d2 = ld(s0 + Q) - ld(s0 + Q + H);
                const float2 u = make_float2(fmaf(d2, w_t.y, d1 * w_t.x), fmaf(-d2, w_t.x, d1 * w_t.y));
                v[c] = mul_tw<c, 128>(u);
            });
        }
        __syncthreads();  // the even unpack has read buf
        eo_fft8192(v, buf, s_t1, s_t2, s_t3, tid);
        for (int k = o_lo + tid; k <= o_hi; k += kEoThreads) {  // X[2k+1] = Y[k]: V[k/2] or conj(V[Q-1-(k-1)/2])
            float2 r;
            if (k & 1) {
                r = buf[eo_addr(Q - 1 - (k >> 1))];
                r.y = -r.y;
            } else {
                r = buf[eo_addr(k >> 1)];
            }
            xs[2 * k + 1 - col_lo] = r;
        }
        __syncthreads();
        if constexpr (EXPORT) {
            float* rh = a_hi + 2 * f * kp;
            float* rl = a_lo + 2 * f * kp;
            for (int c = tid; c < int(kp); c += kEoThreads) {
                float2 X = make_float2(0.f, 0.f);  // padding columns beyond col_hi stay zero
                if (col_lo + c <= col_hi) X = xs[c];
                uint32_t h, l;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(X.x));
                float hv = __uint_as_float(h);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(X.x - hv));
                rh[c] = hv;
                rl[c] = __uint_as_float(l);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(X.y));
                hv = __uint_as_float(h);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(X.y - hv));
                rh[kp + c] = hv;
                rl[kp + c] = __uint_as_float(l);
            }
            __syncthreads();  // xs is rewritten by the next frame's even unpack
            continue;
        }
        // ---- banded kernel rows, one SEGMENT per thread: a row is cut into segments of `seg` consecutive columns (seg odd:
        // the lanes of a warp then read xs from distinct banks), at most one segment per thread, so the whole contraction is
        // `seg` multiply-adds per thread with no cross-lane traffic.  The weights are stored transposed, wT[i][thread]
        // (zero-padded to a multiple of 8 rows): coalesced loads from L1 / L2, eight in flight per thread.
        {
            const float2* X = xs + seg_xoff[tid];
            const float* wp = seg_w + tid;
            float ar = 0.f, ai = 0.f;
            for (int i0 = 0; i0 < seg; i0 += 8) {
                float wv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) wv[i] = __ldg(wp + (i0 + i) * kEoThreads);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 xv = X[i0 + i];
                    ar = fmaf(wv[i], xv.x, ar);
                    ai = fmaf(wv[i], xv.y, ai);
                }
            }
            part[tid] = make_float2(ar, ai);
        }
        __syncthreads();
        if (tid < n_freqs) {  // row sums over the row's segments, in segment order
            const int2 rs = __ldg(row_seg + tid);  // (first segment, number of segments)
            float ar = 0.f, ai = 0.f;
            for (int i = 0; i < rs.y; ++i) {
                const float2 pv = part[rs.x + i];
                ar += pv.x;
                ai += pv.y;
            }
            const float mag = sqrtf(ar * ar + ai * ai);
            if (octave > 0) q[tid] = mag;
            else if (layout == ZAFB_LAYOUT_FRAME_MAJOR) out[f * n_freqs + tid] = mag;
            else out[(clip * int64_t(n_freqs) + tid) * nt + j] = mag;
        }
        for (int r = kEoThreads + tid; r < n_freqs; r += kEoThreads) {  // more rows than threads (not the case for zaf.cqtkernel's grids)
            const int2 rs = __ldg(row_seg + r);
            float ar = 0.f, ai = 0.f;
            for (int i = 0; i < rs.y; ++i) {
                const float2 pv = part[rs.x + i];
                ar += pv.x;
                ai += pv.y;
            }
            const float mag = sqrtf(ar * ar + ai * ai);
            if (octave > 0) q[r] = mag;
            else if (layout == ZAFB_LAYOUT_FRAME_MAJOR) out[f * n_freqs + r] = mag;
            else out[(clip * int64_t(n_freqs) + r) * nt + j] = mag;
        }
        if (octave > 0) {
            __syncthreads();
            for (int i = tid; i < octave; i += kEoThreads) {
                float acc = 0.f;
                for (int r = i; r < n_freqs; r += octave) acc += q[r];  // zaf.py:693-698
                if (layout == ZAFB_LAYOUT_FRAME_MAJOR) out[f * octave + i] = acc;
                else out[(clip * octave + i) * nt + j] = acc;
            }
        }
        // q[] and xs are rewritten only after the next frame's barriers
    }
}

// tensor-core route: magnitude (and optional chroma fold, zaf.py:693-698) of the product rows (Re, Im) per frame
__global__ void cqt_magnitude_kernel(const float* __restrict__ c, int64_t frames, int n_freqs, int octave, int64_t nt,
                                     int64_t frame0, float* __restrict__ out, int layout) {
    const int rows = octave > 0 ? octave : n_freqs;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < frames * rows; i += int64_t(gridDim.x) * blockDim.x) {
        const int64_t f = i / rows;
        const int r = int(i - f * rows);
        const float* re = c + 2 * f * n_freqs;
        const float* im = re + n_freqs;
        float v = 0.f;
        if (octave > 0) {
            for (int k = r; k < n_freqs; k += octave) v += sqrtf(re[k] * re[k] + im[k] * im[k]);
        } else {
            v = sqrtf(re[r] * re[r] + im[r] * im[r]);
        }
        const int64_t fg = frame0 + f, clip = fg / nt, j = fg - clip * nt;
        if (layout == ZAFB_LAYOUT_FRAME_MAJOR) out[fg * rows + r] = v;
        else out[(clip * rows + r) * nt + j] = v;
    }
}

bool g_attr_done = false;
int set_kernel_attrs() {
    if (g_attr_done) return ZAFB_OK;
    ZAFB_CUDA(cudaFuncSetAttribute(cqt_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    ZAFB_CUDA(cudaFuncSetAttribute(cqt32768_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    ZAFB_CUDA(cudaFuncSetAttribute(cqt32768_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    ZAFB_CUDA(cudaFuncSetAttribute(cqt_eo_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    ZAFB_CUDA(cudaFuncSetAttribute(cqt_eo_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    g_attr_done = true;
    return ZAFB_OK;
}

template <class T>
int upload_vec(T** dev, const std::vector<T>& v) {
    ZAFB_CUDA(cudaMalloc(reinterpret_cast<void**>(dev), (v.size() ? v.size() : 1) * sizeof(T)));
    ZAFB_CUDA(cudaMemcpy(*dev, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return ZAFB_OK;
}

}  // namespace

extern "C" {

int zafb_cqt_plan_create(zafb_cqt_plan** out, int64_t n_freqs, int64_t fft_length, const int32_t* indptr,
                         const int32_t* indices, const double* data_ri, int64_t step) {
    ZAFB_REQUIRE(out != nullptr && indptr != nullptr, "plan/indptr is NULL");
    ZAFB_REQUIRE(n_freqs >= 1 && fft_length >= 2 && step >= 1, "bad cqt plan parameters");
    if (!is_pow2(fft_length) || fft_length < 4)
        return fail(ZAFB_E_UNSUPPORTED, "cqt: fft_length %lld is not a power of two >= 4", (long long)fft_length);
    const int64_t m = fft_length / 2;
    if (size_t(m) * sizeof(float2) + size_t(n_freqs) * sizeof(float) + 64 > size_t(kMaxDynSmem))
        return fail(ZAFB_E_UNSUPPORTED, "cqt: fft_length %lld does not fit in shared memory", (long long)fft_length);
    ZAFB_REQUIRE(fft_length >= step, "cqt: step_length %lld exceeds fft_length %lld", (long long)step, (long long)fft_length);
    zafb_cqt_plan* p = new zafb_cqt_plan();
    p->n_freqs = n_freqs;
    p->fft_length = fft_length;
    p->step = step;
    p->log2m = ilog2(m);
    std::vector<int> lo(n_freqs, 0), len(n_freqs, 0), off(n_freqs, 0);
    std::vector<float2> w;
    for (int64_t r = 0; r < n_freqs; ++r) {
        const int b = indptr[r], e = indptr[r + 1];
        off[r] = int(w.size());
        if (e <= b) continue;
        int cmin = indices[b], cmax = indices[b];
        for (int i = b; i < e; ++i) {
            if (indices[i] < 0 || indices[i] >= fft_length) {
                delete p;
                return fail(ZAFB_E_BADARG, "cqt: column index %d out of range", indices[i]);
            }
            cmin = indices[i] < cmin ? indices[i] : cmin;
            cmax = indices[i] > cmax ? indices[i] : cmax;
        }
        lo[r] = cmin;
        len[r] = cmax - cmin + 1;
        const size_t base = w.size();
        w.resize(base + size_t(len[r]), make_float2(0.f, 0.f));
        for (int i = b; i < e; ++i)  // duplicates are summed, like scipy's CSR mat-vec
        {
            float2& dst = w[base + size_t(indices[i] - cmin)];
            dst.x += static_cast<float>(data_ri[2 * i]);
            dst.y += static_cast<float>(data_ri[2 * i + 1]);
        }
    }
    p->packed = int64_t(w.size());
    {   // which pairs (k, M - k) of the half spectrum the bands touch
        int plo = int(m), phi = -1;
        for (int64_t r = 0; r < n_freqs; ++r)
            for (int c = 0; c < len[r]; ++c) {
                const int64_t col = lo[r] + c;
                const int64_t k = col > m ? fft_length - col : col;
                const int pr = int(k < m - k ? k : m - k);
                plo = pr < plo ? pr : plo;
                phi = pr > phi ? pr : phi;
            }
        p->pair_lo = plo;
        p->pair_hi = phi;
    }
    int rc = upload_twiddles(&p->d_tw_fft, m, m);
    if (rc == ZAFB_OK && fft_length == 32768) {
        const double pi = 3.14159265358979323846264338327950288;
        std::vector<double> a(2 * 1024), b(2 * 512);
        for (int q = 0; q < 32; ++q) {
            for (int il = 0; il < 32; ++il) {
                const double ang = -2.0 * pi * double(il * q) / 16384.0;
                a[2 * (q * 32 + il)] = std::cos(ang);
                a[2 * (q * 32 + il) + 1] = std::sin(ang);
            }
            for (int ih = 0; ih < 16; ++ih) {
                const double ang = -2.0 * pi * double((ih * q) % 512) / 512.0;
                b[2 * (q * 16 + ih)] = std::cos(ang);
                b[2 * (q * 16 + ih) + 1] = std::sin(ang);
            }
        }
        rc = upload_c32(&p->d_t1, a.data(), 1024);
        if (rc == ZAFB_OK) rc = upload_c32(&p->d_t2, b.data(), 512);
        // real kernel?  (zaf.cqtkernel's rows are real to 1e-16: centred temporal kernels, zaf.py:515-557.)  An imaginary
        // part below 2^-40 of the row's largest real part cannot change an fp32 accumulation, so it is dropped.
        bool real = true;
        std::vector<float> wre(w.size());
        for (int64_t r = 0; r < n_freqs && real; ++r) {
            float mx = 0.f;
            for (int c = 0; c < len[r]; ++c) mx = std::fmax(mx, std::fabs(w[off[r] + c].x));
            for (int c = 0; c < len[r]; ++c)
                if (std::fabs(w[off[r] + c].y) > mx * 9.1e-13f) real = false;
        }
        for (size_t i = 0; i < w.size(); ++i) wre[i] = w[i].x;
        p->real_weights = real;
        if (rc == ZAFB_OK && real) rc = upload_vec(&p->d_weights_re, wre);
        {   // dense operand of the tensor-core route (real kernels whose bands stay in the half spectrum)
            int clo = int(fft_length), chi = -1;
            bool half = true;
            for (int64_t r = 0; r < n_freqs; ++r) {
                if (len[r] == 0) continue;
                clo = lo[r] < clo ? lo[r] : clo;
                chi = lo[r] + len[r] - 1 > chi ? lo[r] + len[r] - 1 : chi;
                if (lo[r] + len[r] - 1 > m) half = false;
            }
            if (rc == ZAFB_OK && real && half && chi >= clo) {
                p->col_lo = clo;
                p->col_hi = chi;
                p->kp = (int64_t(chi - clo + 1) + 3) & ~int64_t(3);
                std::vector<double> dense(size_t(n_freqs) * p->kp, 0.0);
                for (int64_t r = 0; r < n_freqs; ++r)
                    for (int c = 0; c < len[r]; ++c) dense[r * p->kp + (lo[r] - clo + c)] = double(w[off[r] + c].x);
                std::vector<float> hi(dense.size()), lo_(dense.size());
                split_tf32_host(dense.data(), dense.size(), hi.data(), lo_.data());
                rc = upload_vec(&p->d_kern_hi, hi);
                if (rc == ZAFB_OK) rc = upload_vec(&p->d_kern_lo, lo_);
                // packed form: 16-row groups over the union of their bands, all groups stored with one row pitch
                std::vector<int> glo_v, ghi_v;
                for (int64_t r0 = 0; r0 < n_freqs; r0 += 16) {
                    const int64_t r1 = std::min<int64_t>(n_freqs, r0 + 16);
                    int glo = int(fft_length), ghi = -1;
                    for (int64_t r = r0; r < r1; ++r) {
                        if (len[r] == 0) continue;
                        glo = std::min(glo, lo[r]);
                        ghi = std::max(ghi, lo[r] + len[r] - 1);
                    }
                    if (ghi < glo) glo = ghi = clo;  // an all-zero group still needs a (1-column, zero) operand
                    const int c0 = ((glo - clo) / 4) * 4;  // TMA box origins must be 16-byte aligned: 4 columns
                    const int kg = ghi - clo - c0 + 1;
                    p->pk_col0.push_back(c0);
                    p->pk_k.push_back(kg);
                    p->pk_elems += int64_t(16) * kg;
                    p->pk_ld = std::max(p->pk_ld, (kg + 3) & ~3);
                }
                const size_t groups = p->pk_k.size();
                std::vector<double> packed(groups * 16 * size_t(p->pk_ld), 0.0);
                for (int64_t r = 0; r < n_freqs; ++r)
                    for (int c = 0; c < len[r]; ++c)
                        packed[size_t(r) * p->pk_ld + (lo[r] - clo - p->pk_col0[r / 16] + c)] = double(w[off[r] + c].x);
                std::vector<float> phi(packed.size()), plo(packed.size());
                split_tf32_host(packed.data(), packed.size(), phi.data(), plo.data());
                if (rc == ZAFB_OK) rc = upload_vec(&p->d_pk_hi, phi);
                if (rc == ZAFB_OK) rc = upload_vec(&p->d_pk_lo, plo);
                std::vector<GemmTile> tiles(groups);
                for (size_t g = 0; g < groups; ++g)
                    tiles[g] = GemmTile{p->pk_col0[g], p->pk_k[g], int(16 * g), 1, int(std::min<int64_t>(16, n_freqs - int64_t(16 * g)))};
                if (rc == ZAFB_OK) rc = upload_vec(&p->d_pk_tiles, tiles);
            }
        }
        // longest-processing-time-first deal of the rows to the warps of a CTA (16 for the register-FFT kernel, 8 for the
        // even/odd kernel)
        auto deal = [&](int n_warps, int** d_sched, int** d_cnt, int* stride) -> int {
            std::vector<int> order(n_freqs);
            for (int64_t r = 0; r < n_freqs; ++r) order[r] = int(r);
            std::stable_sort(order.begin(), order.end(), [&](int x1, int x2) { return len[x1] > len[x2]; });
            std::vector<std::vector<int>> lists(n_warps);
            std::vector<int64_t> load(n_warps, 0);
            for (int r : order) {
                int best = 0;
                for (int wv = 1; wv < n_warps; ++wv)
                    if (load[wv] < load[best]) best = wv;
                lists[best].push_back(r);
                load[best] += (len[r] + 31) / 32 + 1;  // iterations + the reduction
            }
            size_t longest = 1;
            for (auto& l : lists) longest = l.size() > longest ? l.size() : longest;
            std::vector<int> sched(n_warps * longest, 0), cnt(n_warps, 0);
            for (int wv = 0; wv < n_warps; ++wv) {
                cnt[wv] = int(lists[wv].size());
                for (size_t i = 0; i < lists[wv].size(); ++i) sched[wv * longest + i] = lists[wv][i];
            }
            *stride = int(longest);
            int r2 = upload_vec(d_sched, sched);
            if (r2 == ZAFB_OK) r2 = upload_vec(d_cnt, cnt);
            return r2;
        };
        if (rc == ZAFB_OK) rc = deal(kRegThreads / 32, &p->d_sched, &p->d_sched_cnt, &p->sched_stride);
        if (rc == ZAFB_OK && real && p->d_kern_hi != nullptr) {
            // even/odd kernel: every row cut into segments of `seg` columns, at most kEoThreads segments in all; seg is the
            // smallest odd length that fits (odd: consecutive segments of a row start in different shared-memory banks)
            int seg = 0;
            for (int cand = 1; cand <= kEoPad - 8; cand += 2) {
                int64_t count = 0;
                for (int64_t r = 0; r < n_freqs; ++r) count += len[r] > 0 ? (len[r] + cand - 1) / cand : 0;
                if (count <= kEoThreads) {
                    seg = cand;
                    break;
                }
            }
            if (seg > 0) {
                const int seg8 = (seg + 7) & ~7;
                std::vector<int> xoff(kEoThreads, 0);
                std::vector<float> wt(size_t(seg8) * kEoThreads, 0.f);
                std::vector<int2> rs(n_freqs, make_int2(0, 0));
                int next = 0;
                for (int64_t r = 0; r < n_freqs; ++r) {
                    const int count = len[r] > 0 ? (len[r] + seg - 1) / seg : 0;
                    rs[r] = make_int2(next, count);
                    for (int sg = 0; sg < count; ++sg, ++next) {
                        xoff[next] = lo[r] - p->col_lo + sg * seg;
                        for (int e = 0; e < seg && sg * seg + e < len[r]; ++e)
                            wt[size_t(e) * kEoThreads + next] = w[off[r] + sg * seg + e].x;
                    }
                }
                p->eo_seg = seg;
                rc = upload_vec(&p->d_eo_seg_xoff, xoff);
                if (rc == ZAFB_OK) rc = upload_vec(&p->d_eo_seg_w, wt);
                if (rc == ZAFB_OK) rc = upload_vec(&p->d_eo_row_seg, rs);
            }
        }
        if (rc == ZAFB_OK) {  // twiddles of the even/odd kernel's 8192-point transforms
            std::vector<double> e1(2 * 1024), e2(2 * 256), e3(2 * 256);
            for (int k1 = 0; k1 < 32; ++k1) {
                for (int tl = 0; tl < 32; ++tl) {
                    const double ang = -2.0 * pi * double(tl * k1) / 8192.0;
                    e1[2 * (k1 * 32 + tl)] = std::cos(ang);
                    e1[2 * (k1 * 32 + tl) + 1] = std::sin(ang);
                }
                for (int th = 0; th < 8; ++th) {
                    const double ang = -2.0 * pi * double((th * k1) % 256) / 256.0;
                    e2[2 * (k1 * 8 + th)] = std::cos(ang);
                    e2[2 * (k1 * 8 + th) + 1] = std::sin(ang);
                }
            }
            for (int k2 = 0; k2 < 16; ++k2)
                for (int t0 = 0; t0 < 16; ++t0) {
                    const double ang = -2.0 * pi * double(t0 * k2) / 256.0;
                    e3[2 * (k2 * 16 + t0)] = std::cos(ang);
                    e3[2 * (k2 * 16 + t0) + 1] = std::sin(ang);
                }
            rc = upload_c32(&p->d_eo_t1, e1.data(), 1024);
            if (rc == ZAFB_OK) rc = upload_c32(&p->d_eo_t2, e2.data(), 256);
            if (rc == ZAFB_OK) rc = upload_c32(&p->d_eo_t3, e3.data(), 256);
        }
    }
    if (rc == ZAFB_OK) rc = upload_twiddles(&p->d_tw_full, fft_length, m + 1);
    if (rc == ZAFB_OK) rc = upload_vec(&p->d_band_lo, lo);
    if (rc == ZAFB_OK) rc = upload_vec(&p->d_band_len, len);
    if (rc == ZAFB_OK) rc = upload_vec(&p->d_band_off, off);
    if (rc == ZAFB_OK) rc = upload_vec(&p->d_weights, w);
    if (rc != ZAFB_OK) {
        zafb_cqt_plan_destroy(p);
        return rc;
    }
    *out = p;
    return ZAFB_OK;
}

int zafb_cqt_plan_destroy(zafb_cqt_plan* p) {
    if (!p) return ZAFB_OK;
    cudaFree(p->d_tw_fft);
    cudaFree(p->d_tw_full);
    cudaFree(p->d_band_lo);
    cudaFree(p->d_band_len);
    cudaFree(p->d_band_off);
    cudaFree(p->d_weights);
    cudaFree(p->d_t1);
    cudaFree(p->d_t2);
    cudaFree(p->d_weights_re);
    cudaFree(p->d_sched);
    cudaFree(p->d_sched_cnt);
    cudaFree(p->d_kern_hi);
    cudaFree(p->d_kern_lo);
    cudaFree(p->d_pk_hi);
    cudaFree(p->d_pk_lo);
    cudaFree(p->d_pk_tiles);
    cudaFree(p->d_eo_t1);
    cudaFree(p->d_eo_t2);
    cudaFree(p->d_eo_t3);
    cudaFree(p->d_eo_seg_xoff);
    cudaFree(p->d_eo_seg_w);
    cudaFree(p->d_eo_row_seg);
    delete p;
    return ZAFB_OK;
}

int zafb_cqt_plan_set_route(zafb_cqt_plan* p, int route) {
    ZAFB_REQUIRE(p != nullptr, "plan is NULL");
    ZAFB_REQUIRE(route == ZAFB_CQT_ROUTE_FUSED || route == ZAFB_CQT_ROUTE_TENSOR, "bad route %d", route);
    if (route == ZAFB_CQT_ROUTE_TENSOR && p->d_kern_hi == nullptr)
        return fail(ZAFB_E_UNSUPPORTED, "cqt tensor-core route needs fft_length 32768 and a real kernel whose bands stay below fft_length/2");
    p->route = route;
    return ZAFB_OK;
}

// test hook: 0 = auto, 1 = generic kernel only, 2 = require the register-FFT kernel, 3 = require the even/odd kernel
int zafb_cqt_plan_force_kernel(zafb_cqt_plan* p, int which) {
    ZAFB_REQUIRE(p != nullptr && which >= 0 && which <= 3, "bad plan / kernel id");
    p->force_kernel = which;
    return ZAFB_OK;
}

int zafb_cqt_f32(const zafb_cqt_plan* p, const float* x, int64_t n_clips, int64_t ns, int64_t clip_stride,
                 int64_t octave_resolution, float* out, int layout, void* stream) {
    ZAFB_REQUIRE(p != nullptr, "plan is NULL");
    ZAFB_REQUIRE(n_clips >= 0 && ns >= 0 && clip_stride >= ns, "bad batch geometry");
    ZAFB_REQUIRE(layout == ZAFB_LAYOUT_FRAME_MAJOR || layout == ZAFB_LAYOUT_BIN_MAJOR, "bad layout %d", layout);
    ZAFB_REQUIRE(octave_resolution >= 0 && octave_resolution <= p->n_freqs, "bad octave_resolution");
    int rc = set_kernel_attrs();
    if (rc != ZAFB_OK) return rc;
    int64_t nt = 0, front = 0;
    rc = zafb_cqt_geometry(ns, p->step, p->fft_length, &nt, &front, nullptr);
    if (rc != ZAFB_OK) return rc;
    const int64_t total = n_clips * nt;
    if (total == 0) return ZAFB_OK;
    ZAFB_REQUIRE(out != nullptr && x != nullptr, "x/out is NULL");
    const int64_t m = p->fft_length / 2;
    {
        size_t smem_reg = size_t(kRegM + 1536) * sizeof(float2) + size_t((p->n_freqs + 1) & ~int64_t(1)) * sizeof(float) + 64;
        const bool ok = p->fft_length == 32768 && p->d_t1 != nullptr && p->d_sched != nullptr && smem_reg <= size_t(kMaxDynSmem);
        // optional shared-memory residents, in order of benefit: the real band weights, the split twiddles
        int smem_weights = 0, smem_split = 0;
        const size_t split_bytes = p->pair_hi >= p->pair_lo ? size_t(p->pair_hi - p->pair_lo + 1) * sizeof(float2) : 0;
        if (ok && split_bytes && smem_reg + split_bytes <= size_t(kMaxDynSmem)) {
            smem_split = 1;
            smem_reg += split_bytes;
        }
        if (ok && p->real_weights && smem_reg + size_t(p->packed) * sizeof(float) <= size_t(kMaxDynSmem)) {
            smem_weights = 1;
            smem_reg += size_t(p->packed) * sizeof(float);
        }
        if (p->force_kernel == 2 && !ok) return fail(ZAFB_E_UNSUPPORTED, "cqt register-FFT kernel needs fft_length 32768");
        // even/odd kernel: real weights, every band inside [0, L/2), the unpacked bins fit next to the 64 KB transform buffer
        const int64_t nb = int64_t(p->col_hi) - p->col_lo + 1;
        const size_t smem_eo = size_t(kEoBuf + 1536 + (nb > 0 ? nb : 0) + kEoPad + kEoThreads) * sizeof(float2) +
                               size_t((p->n_freqs + 1) & ~int64_t(1)) * sizeof(float) + 16;
        const bool eo_ok = ok && p->real_weights && p->d_eo_t1 != nullptr && p->eo_seg > 0 && p->d_kern_hi != nullptr &&
                           p->col_hi < p->fft_length / 2 && p->col_lo >= 0 && smem_eo <= size_t(kMaxDynSmem);
        if (p->force_kernel == 3 && !eo_ok)
            return fail(ZAFB_E_UNSUPPORTED, "cqt even/odd kernel needs fft_length 32768 and a real kernel with bands below fft_length/2");
        const bool use_eo = eo_ok && (p->force_kernel == 3 || (p->force_kernel == 0 && env_flag("ZAFB_CQT_EO", 1)));
        auto launch_eo = [&](auto kern, const float* xs, int64_t frames, float* o, float* ah, float* al, cudaStream_t st) {
            // two CTAs per SM when the shared memory allows it (it does for every kernel zaf.cqtkernel builds up to ~4700 bins)
            const int64_t per_sm = 2 * (smem_eo + 1024) <= size_t(232448) ? 2 : 1;
            const int64_t cap = int64_t(sm_count()) * per_sm;
            const int64_t grid = frames < cap ? frames : cap;
            kern<<<unsigned(grid), kEoThreads, smem_eo, st>>>(
                xs, ns, clip_stride, nt, p->step, front, p->d_eo_t1, p->d_eo_t2, p->d_eo_t3, p->d_tw_full, p->d_band_lo,
                p->d_band_len, p->d_band_off, p->d_weights_re, p->d_eo_seg_xoff, p->d_eo_seg_w, p->d_eo_row_seg, (p->eo_seg + 7) & ~7,
                int(p->n_freqs), int(octave_resolution), p->col_lo, p->col_hi, o, layout, frames, ah, al, p->kp);
            g_launches.fetch_add(1, std::memory_order_relaxed);
        };
        if (p->route == ZAFB_CQT_ROUTE_TENSOR) {
            // The kernel application as a dense contraction on the tensor cores (BASELINE cfg 5: "sparse CQT kernel as packed
            // tensor-core GEMM"): per chunk of clips (1) cqt32768_kernel<true>: FFT + real-input split, Re / Im of the bins
            // [col_lo, col_hi] written as two TF32 hi/lo rows per frame, (2) gemm3xtf32: (2 frames x kp) . (n_freqs x kp)^T,
            // (3) magnitude (+ chroma fold).  The fused banded form above does 7x fewer multiply-adds and never writes the
            // spectrum; this route is the dense-operand form and the measured evidence.
            if (!ok || p->d_kern_hi == nullptr)
                return fail(ZAFB_E_UNSUPPORTED, "cqt tensor-core route needs fft_length 32768 and a real half-spectrum kernel");
            cudaStream_t st = static_cast<cudaStream_t>(stream);
            static bool pool_ready = false;
            if (!pool_ready) {
                int dev = 0;
                cudaMemPool_t pool;
                uint64_t keep = UINT64_MAX;
                if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
                    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
                pool_ready = true;
            }
            const int64_t kp = p->kp, nf = p->n_freqs;
            // chunk: one wave of 128-row GEMM tiles over all SMs (2 rows per frame), at most 512 MB of split spectra
            const int64_t frames_cap = std::max<int64_t>(1, std::min<int64_t>(int64_t(sm_count()) * 64, (int64_t(512) << 20) / (2 * kp * 4 * 2)));
            const int64_t clips_per = std::max<int64_t>(1, std::min<int64_t>(n_clips, frames_cap / nt));
            auto round64 = [](size_t v) { return (v + 63) & ~size_t(63); };
            const size_t a_f = round64(size_t(clips_per) * nt * 2 * kp), c_f = round64(size_t(clips_per) * nt * 2 * nf);
            float* ws = nullptr;
            ZAFB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ws), (2 * a_f + c_f) * sizeof(float), st));
            float *a_hi = ws, *a_lo = ws + a_f, *cbuf = ws + 2 * a_f;
            const size_t smem_x = size_t(kRegM + 1536) * sizeof(float2) + size_t((nf + 1) & ~int64_t(1)) * sizeof(float) + 64 +
                                  (smem_split ? split_bytes : 0);
            for (int64_t c0 = 0; c0 < n_clips && rc == ZAFB_OK; c0 += clips_per) {
                const int64_t nc = std::min(clips_per, n_clips - c0), frames = nc * nt;
                if (use_eo) {
                    launch_eo(cqt_eo_kernel<true>, x + c0 * clip_stride, frames, nullptr, a_hi, a_lo, st);
                } else {
                    const int64_t grid = frames < int64_t(sm_count()) ? frames : int64_t(sm_count());
                    cqt32768_kernel<true><<<unsigned(grid), kRegThreads, smem_x, st>>>(
                        x + c0 * clip_stride, ns, clip_stride, nt, p->step, front, p->d_t1, p->d_t2, p->d_tw_full, p->d_band_lo,
                        p->d_band_len, p->d_band_off, p->d_weights, p->d_weights_re, int(p->packed), 0, smem_split, p->d_sched,
                        p->d_sched_cnt, p->sched_stride, int(nf), 0, p->pair_lo, p->pair_hi, nullptr, layout, frames, a_hi, a_lo,
                        p->col_lo, p->col_hi, kp);
                    g_launches.fetch_add(1, std::memory_order_relaxed);
                }
                if (p->d_pk_hi != nullptr && !env_flag("ZAFB_CQT_TENSOR_DENSE", 0)) {
                    // the PACKED banded contraction: every 16-row group multiplies only the union of its rows' bands
                    rc = gemm3xtf32_tiled(16, a_hi, a_lo, kp, kp, p->d_pk_hi, p->d_pk_lo, p->pk_ld, int(p->pk_k.size()), p->d_pk_tiles, cbuf,
                                          nf, 2 * frames, st);
                } else {  // the dense (n_freqs x kp) block
                    rc = gemm3xtf32(a_hi, a_lo, kp, p->d_kern_hi, p->d_kern_lo, kp, cbuf, nf, 2 * frames, nf, kp, st);
                }
                if (rc == ZAFB_OK) {
                    const int rows = octave_resolution > 0 ? int(octave_resolution) : int(nf);
                    int64_t blocks = ceil_div(frames * rows, 256);
                    if (blocks > int64_t(sm_count()) * 16) blocks = int64_t(sm_count()) * 16;
                    cqt_magnitude_kernel<<<unsigned(blocks), 256, 0, st>>>(cbuf, frames, int(nf), int(octave_resolution), nt, c0 * nt,
                                                                           out, layout);
                    g_launches.fetch_add(1, std::memory_order_relaxed);
                }
            }
            cudaFreeAsync(ws, st);
            if (rc == ZAFB_OK) ZAFB_CUDA(cudaGetLastError());
            return rc;
        }
        if (use_eo) {
            launch_eo(cqt_eo_kernel<false>, x, total, out, nullptr, nullptr, static_cast<cudaStream_t>(stream));
            ZAFB_CUDA(cudaGetLastError());
            return ZAFB_OK;
        }
        if (ok && p->force_kernel != 1) {
            const int64_t grid = total < int64_t(sm_count()) ? total : int64_t(sm_count());
            cqt32768_kernel<false><<<unsigned(grid), kRegThreads, smem_reg, static_cast<cudaStream_t>(stream)>>>(
                x, ns, clip_stride, nt, p->step, front, p->d_t1, p->d_t2, p->d_tw_full, p->d_band_lo, p->d_band_len,
                p->d_band_off, p->d_weights, p->d_weights_re, int(p->packed), smem_weights, smem_split, p->d_sched,
                p->d_sched_cnt, p->sched_stride, int(p->n_freqs), int(octave_resolution), p->pair_lo, p->pair_hi, out, layout,
                total, nullptr, nullptr, 0, -1, 0);
            ZAFB_LAUNCH_CHECK();
            return ZAFB_OK;
        }
    }
    const size_t smem = size_t(m) * sizeof(float2) + size_t(p->n_freqs) * sizeof(float) + 64;
    int th = int(m / 4);
    if (th < 64) th = 64;
    if (th > kThreads) th = kThreads;
    const int64_t resident = int64_t(sm_count()) * (smem > 100 * 1024 ? 1 : 2);
    const int64_t grid = total < resident ? total : resident;
    cqt_frame_kernel<<<unsigned(grid), th, smem, static_cast<cudaStream_t>(stream)>>>(
        x, ns, clip_stride, nt, p->step, front, p->log2m, p->d_tw_fft, p->d_tw_full, p->d_band_lo, p->d_band_len,
        p->d_band_off, p->d_weights, int(p->n_freqs), int(octave_resolution), out, layout, total);
    ZAFB_LAUNCH_CHECK();
    return ZAFB_OK;
}

int zafb_cqt_host_f32(const zafb_cqt_plan* p, const float* x, int64_t n_clips, int64_t ns, int64_t clip_stride,
                      int64_t octave_resolution, float* out, int layout) {
    ZAFB_REQUIRE(p != nullptr, "plan is NULL");
    ZAFB_REQUIRE(n_clips >= 0 && ns >= 0 && clip_stride >= ns, "bad batch geometry");
    ZAFB_REQUIRE(octave_resolution >= 0 && octave_resolution <= p->n_freqs, "bad octave_resolution");
    int64_t nt = 0;
    int rc = zafb_cqt_geometry(ns, p->step, p->fft_length, &nt, nullptr, nullptr);
    if (rc != ZAFB_OK) return rc;
    if (n_clips == 0 || nt == 0) return ZAFB_OK;
    ZAFB_REQUIRE(out != nullptr && x != nullptr, "x/out is NULL");
    const int64_t rows = octave_resolution > 0 ? octave_resolution : p->n_freqs;
    const size_t out_clip = size_t(nt) * size_t(rows) * sizeof(float);
    const int64_t dpitch = (ns + 1) & ~int64_t(1);
    return run_host_pipeline(x, size_t(clip_stride) * sizeof(float), size_t(ns) * sizeof(float), size_t(dpitch) * sizeof(float),
                             out, out_clip, out_clip, out_clip, n_clips,
                             [&](void* d_in, void* d_out, int64_t, int64_t nc, cudaStream_t st) {
                                 return zafb_cqt_f32(p, static_cast<const float*>(d_in), nc, ns, dpitch, octave_resolution,
                                                     static_cast<float*>(d_out), layout, st);
                             });
}

}  // extern "C"
