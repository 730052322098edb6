// FFT building blocks shared by every transform kernel.
//
//  (1) fft_reg<N>: radix-2 decimation-in-frequency FFT of N complex points held in the registers
//      of ONE thread (fully unrolled, twiddles are immediates).  The result is left in
//      bit-reversed register order: X[k] is v[bitrev(k, log2 N)], which costs nothing because
//      every index is a compile-time constant.
//  (2) stockham_pass<R>: one radix-R butterfly of an autosort (Stockham) FFT whose data lives in
//      shared memory; the generic block FFT used for sizes without a specialised kernel.
//
// All functions are __host__ __device__ so tests/host/fft_core_test.cu can check them on the CPU.
#pragma once

#include <type_traits>

#include "common.cuh"

namespace zafb {

// ---------------------------------------------------------------- in-register FFT
template <int N, int J, int STRIDE>
struct DifButterflies {
    template <class C>  // C = float2 or double2
    static ZAFB_HD void run(C* v) {
        // span N/2 butterflies of one radix-2 DIF stage on v[0], v[STRIDE], ...
        C a = v[J * STRIDE];
        C b = v[(J + N / 2) * STRIDE];
        v[J * STRIDE] = cadd(a, b);
        v[(J + N / 2) * STRIDE] = mul_tw<J, N>(csub(a, b));
        if constexpr (J + 1 < N / 2) DifButterflies<N, J + 1, STRIDE>::run(v);
    }
};

// In-place DIF FFT of v[0], v[STRIDE], ..., v[(N-1)*STRIDE]; output bit-reversed.
template <int N, int STRIDE = 1, class C>
ZAFB_HD void fft_reg(C* v) {
    if constexpr (N >= 2) {
        DifButterflies<N, 0, STRIDE>::run(v);
        fft_reg<N / 2, STRIDE>(v);
        fft_reg<N / 2, STRIDE>(v + (N / 2) * STRIDE);
    }
}

// ---------------------------------------------------------------- Stockham pass (shared memory)
// One radix-R butterfly j (0 <= j < M/R) of the pass with sub-transform length Ns (product of the
// radices of the previous passes).  tw[t] = exp(-2 pi i t / M), t < M.  Natural order in and out.
template <int R, class C>
ZAFB_HD void stockham_pass(const C* __restrict__ in, C* __restrict__ out,
                           const C* __restrict__ tw, int M, int Ns, int j) {
    const int k = j & (Ns - 1);
    const int tstride = M / (Ns * R);
    C v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        v[r] = in[j + r * (M / R)];
        if (r > 0 && Ns > 1) v[r] = cmul(v[r], tw[r * k * tstride]);
    }
    fft_reg<R>(v);
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) out[j0 + r * Ns] = v[bitrev(r, clog2(R))];
}

// Radix schedule for an M-point Stockham FFT: radix-4 passes plus one radix-2 or radix-8 pass so
// that the number of passes is minimal.  Returns the number of passes, radices in `radix`.
ZAFB_HD int stockham_schedule(int log2m, int* radix) {
    int n = 0;
    int rem = log2m;
    while (rem > 0) {
        int r;
        if (rem == 1) r = 1;
        else if (rem == 3) r = 3;
        else r = 2;
        radix[n++] = 1 << r;
        rem -= r;
    }
    return n;
}

#ifdef __CUDACC__
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

// ---------------------------------------------------------------- one warp, 1024 complex points
// Four-step FFT 1024 = 32 x 32 by ONE warp.  In: v[r] = x[lane + 32 r].  Out: X[lane + 32 k2] is
// v[bitrev(k2, 5)] (same element-to-thread map as the input, register order bit-reversed).
//   tw4[k1 * 32 + n2] = W_1024^{k1 n2} (shared memory).
// SPLIT = false: `buf` is a 32 x kFft1024Pitch float2 tile, one 64-bit transpose.
// SPLIT = true : `buf` is a 32 x kFft1024Pitch float tile, the real and imaginary parts are
//                transposed one after the other (half the shared memory, twice the instructions).
constexpr int kFft1024Pitch = 33;  // row pitch of the transpose tile (conflict-free)

template <bool SPLIT>
__device__ __forceinline__ void warp_fft1024(float2 (&v)[32], const float2* __restrict__ tw4, void* buf, int lane) {
    fft_reg<32>(v);  // over n1 (register index); thread = n2
    if constexpr (!SPLIT) {
        float2* s = static_cast<float2*>(buf);
        static_for<0, 32>([&](auto k1c) {
            constexpr int k1 = decltype(k1c)::value;
            float2 y = v[bitrev(k1, 5)];
            if constexpr (k1 > 0) y = cmul(y, tw4[k1 * 32 + lane]);
            s[k1 * kFft1024Pitch + lane] = y;
        });
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 32; ++n2) v[n2] = s[lane * kFft1024Pitch + n2];
        __syncwarp();
    } else {
        float* s = static_cast<float*>(buf);
        float im[32];
        static_for<0, 32>([&](auto k1c) {
            constexpr int k1 = decltype(k1c)::value;
            float2 y = v[bitrev(k1, 5)];
            if constexpr (k1 > 0) y = cmul(y, tw4[k1 * 32 + lane]);
            s[k1 * kFft1024Pitch + lane] = y.x;
            im[k1] = y.y;
        });
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 32; ++n2) v[n2].x = s[lane * kFft1024Pitch + n2];
        __syncwarp();
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) s[k1 * kFft1024Pitch + lane] = im[k1];
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 32; ++n2) v[n2].y = s[lane * kFft1024Pitch + n2];
        __syncwarp();
    }
    fft_reg<32>(v);  // over n2; thread = k1
}

// ---------------------------------------------------------------- one warp, 2048 complex points
// Four-step FFT 2048 = 64 x 32 by ONE warp (window length 4096).  In: v[r] = x[lane + 32 r], r < 64.  Out: X[lane + 32 k]
// is v[bitrev(k, 6)], the convention of the smaller transforms.  Step 1: FFT-64 over the register index (n1), twiddle
// W_2048^{k1 n2} from tw4[k1 * 32 + n2] (64 x 32 table).  Step 2: 64 FFT-32 over n2, two per lane: rows k1 = lane and
// lane + 32 of the 64 x kFft1024Pitch transpose tile, held in v[0..31] and v[32..63]; their outputs k2 are
// X[k1 + 64 k2] = X[lane + 32 (h + 2 k2)] with h the row half, and bitrev(h + 2 k2, 6) = 32 h + bitrev(k2, 5).
__device__ __forceinline__ void warp_fft2048(float2 (&v)[64], const float2* __restrict__ tw4, float2* buf, int lane) {
    fft_reg<64>(v);
    static_for<0, 64>([&](auto k1c) {
        constexpr int k1 = decltype(k1c)::value;
        float2 y = v[bitrev(k1, 6)];
        if constexpr (k1 > 0) y = cmul(y, tw4[k1 * 32 + lane]);
        buf[k1 * kFft1024Pitch + lane] = y;
    });
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) {
        v[n2] = buf[lane * kFft1024Pitch + n2];
        v[32 + n2] = buf[(lane + 32) * kFft1024Pitch + n2];
    }
    __syncwarp();
    fft_reg<32>(&v[0]);
    fft_reg<32>(&v[32]);
}

// ---------------------------------------------------------------- one warp, 512 complex points
// Four-step FFT 512 = 16 x 32 by ONE warp.  In: v[r] = x[lane + 32 r], r < 16.  Out: X[lane + 32 k]
// is v[bitrev(k, 4)].  tw[k1 * 32 + n2] = W_512^{k1 n2} (shared memory, 16 x 32); buf: 16 x
// kFft1024Pitch float2.  Step 1: FFT-16 over the register index.  Step 2: the 16 FFT-32 over n2
// are shared by lane pairs (k1, parity): each lane does the radix-2 butterfly for its output
// parity while reading its row back from the transpose tile, then an FFT-16 in registers.
__device__ __forceinline__ void warp_fft512(float2 (&v)[16], const float2* __restrict__ tw, float2* buf, int lane) {
    fft_reg<16>(v);
    static_for<0, 16>([&](auto k1c) {
        constexpr int k1 = decltype(k1c)::value;
        float2 y = v[bitrev(k1, 4)];
        if constexpr (k1 > 0) y = cmul(y, tw[k1 * 32 + lane]);
        buf[k1 * kFft1024Pitch + lane] = y;
    });
    __syncwarp();
    const float2* row = buf + (lane & 15) * kFft1024Pitch;
    const bool odd = lane >= 16;
    const float sgn = odd ? -1.f : 1.f;
    static_for<0, 16>([&](auto nc) {
        constexpr int n2 = decltype(nc)::value;
        const float2 a = row[n2];
        const float2 b = row[n2 + 16];
        const float2 d = make_float2(fmaf(sgn, b.x, a.x), fmaf(sgn, b.y, a.y));
        if constexpr (n2 == 0) {
            v[n2] = d;
        } else {
            const float wr = odd ? Tw<n2, 32>::re : 1.f;
            const float wi = odd ? Tw<n2, 32>::im : 0.f;
            v[n2] = make_float2(d.x * wr - d.y * wi, d.x * wi + d.y * wr);
        }
    });
    __syncwarp();
    fft_reg<16>(v);
}

// The same four-step FFT in double precision (the float64 route of mfcc).  buf: 16 x kFft1024Pitch double2,
// tw[k1 * 32 + n2] = W_512^{k1 n2} in double.  In / out conventions as warp_fft512.
__device__ __forceinline__ void warp_fft512_f64(double2 (&v)[16], const double2* __restrict__ tw, double2* buf, int lane) {
    fft_reg<16>(v);
    static_for<0, 16>([&](auto k1c) {
        constexpr int k1 = decltype(k1c)::value;
        double2 y = v[bitrev(k1, 4)];
        if constexpr (k1 > 0) y = cmul(y, tw[k1 * 32 + lane]);
        buf[k1 * kFft1024Pitch + lane] = y;
    });
    __syncwarp();
    const double2* row = buf + (lane & 15) * kFft1024Pitch;
    const bool odd = lane >= 16;
    const double sgn = odd ? -1.0 : 1.0;
    static_for<0, 16>([&](auto nc) {
        constexpr int n2 = decltype(nc)::value;
        const double2 a = row[n2];
        const double2 b = row[n2 + 16];
        const double2 d = make_double2(fma(sgn, b.x, a.x), fma(sgn, b.y, a.y));
        if constexpr (n2 == 0) {
            v[n2] = d;
        } else {
            constexpr double cr = ct::cos2pi(n2, 32), ci = -ct::sin2pi(n2, 32);
            const double wr = odd ? cr : 1.0;
            const double wi = odd ? ci : 0.0;
            v[n2] = make_double2(d.x * wr - d.y * wi, d.x * wi + d.y * wr);
        }
    });
    __syncwarp();
    fft_reg<16>(v);
}

// ---------------------------------------------------------------- one warp, 256 complex points
// Four-step FFT 256 = 8 x 32 by ONE warp.  In: v[r] = x[lane + 32 r], r < 8.  Out: X[lane + 32 k] is v[bitrev(k, 3)].
// tw[k1 * 32 + n2] = W_256^{k1 n2} (shared memory, 8 x 32); buf: 8 x kFft1024Pitch float2.
// Step 1: FFT-8 over the register index.  Step 2: the 8 FFT-32 over n2 are shared by lane quads (k1 = lane & 7,
// q = lane >> 3): with n2 = a + 8 b and k2 = q + 4 j,
//     X[k1 + 8 k2] = sum_a W_8^{a j} W_32^{a q} ( (y[a] + (-1)^q y[a+16]) + (-i)^q (y[a+8] + (-1)^q y[a+24]) ),
// i.e. each lane forms output q of the radix-4 butterflies of its row while reading it back from the transpose
// tile, applies its own twiddles tq[a] = W_32^{a q} (8 per-lane constants, see warp_fft256_lane_twiddles) and
// finishes with an FFT-8 in registers.  (Validated in float64 against numpy.fft.)
__device__ __forceinline__ void warp_fft256_lane_twiddles(float2 (&tq)[8], int lane) {
    const int q = lane >> 3;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        float sn, cs;
        sincospif(-float(a * q) / 16.0f, &sn, &cs);  // W_32^{a q} = exp(-2 pi i a q / 32); exact for the multiples of pi/2
        tq[a] = make_float2(cs, sn);
    }
}

__device__ __forceinline__ void warp_fft256(float2 (&v)[8], const float2* __restrict__ tw, float2* buf, int lane,
                                            const float2 (&tq)[8]) {
    fft_reg<8>(v);
    static_for<0, 8>([&](auto k1c) {
        constexpr int k1 = decltype(k1c)::value;
        float2 y = v[bitrev(k1, 3)];
        if constexpr (k1 > 0) y = cmul(y, tw[k1 * 32 + lane]);
        buf[k1 * kFft1024Pitch + lane] = y;
    });
    __syncwarp();
    const float2* row = buf + (lane & 7) * kFft1024Pitch;
    const int q = lane >> 3;
    const float sgn = (q & 1) ? -1.f : 1.f;
    const float cq = q == 0 ? 1.f : (q == 2 ? -1.f : 0.f);  // (-i)^q = cq - i sq
    const float sq = q == 1 ? 1.f : (q == 3 ? -1.f : 0.f);
    static_for<0, 8>([&](auto ac) {
        constexpr int a = decltype(ac)::value;
        const float2 y0 = row[a], y1 = row[a + 8], y2 = row[a + 16], y3 = row[a + 24];
        const float2 e = make_float2(fmaf(sgn, y2.x, y0.x), fmaf(sgn, y2.y, y0.y));
        const float2 o = make_float2(fmaf(sgn, y3.x, y1.x), fmaf(sgn, y3.y, y1.y));
        const float2 t = make_float2(e.x + cq * o.x + sq * o.y, e.y + cq * o.y - sq * o.x);
        if constexpr (a == 0) v[a] = t;
        else v[a] = cmul(t, tq[a]);
    });
    __syncwarp();
    fft_reg<8>(v);
}

// ---------------------------------------------------------------- one warp, 128 complex points
// Four-step FFT 128 = 4 x 32 by ONE warp (window length 256; MDCT window 512).  In: v[r] = x[lane + 32 r], r < 4.  Out:
// X[lane + 32 k] is v[bitrev(k, 2)].  tw[k1 * 32 + n2] = W_128^{k1 n2} (shared memory, 4 x 32); buf: 4 x kFft1024Pitch
// float2.  Step 1: FFT-4 over the register index.  Step 2: the 4 FFT-32 over n2 are shared by lane octets (k1 = lane & 3,
// o = lane >> 2): with n2 = a + 4 b and k2 = o + 8 j,
//     X[k1 + 4 k2] = sum_a W_4^{a j} W_32^{a o} ( sum_b W_8^{b o} y[a + 4 b] ),
// i.e. each lane forms output o of the four radix-8 butterflies of its row while reading it back from the transpose
// tile (per-lane constants tq[b] = W_8^{b o}, b < 8), applies tq[8 + a] = W_32^{a o} and finishes with an FFT-4 in registers.
__device__ __forceinline__ void warp_fft128_lane_twiddles(float2 (&tq)[12], int lane) {
    const int o = lane >> 2;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        float sn, cs;
        sincospif(-float(b * o) / 4.0f, &sn, &cs);  // W_8^{b o}
        tq[b] = make_float2(cs, sn);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        float sn, cs;
        sincospif(-float(a * o) / 16.0f, &sn, &cs);  // W_32^{a o}
        tq[8 + a] = make_float2(cs, sn);
    }
}

__device__ __forceinline__ void warp_fft128(float2 (&v)[4], const float2* __restrict__ tw, float2* buf, int lane,
                                            const float2 (&tq)[12]) {
    fft_reg<4>(v);
    static_for<0, 4>([&](auto k1c) {
        constexpr int k1 = decltype(k1c)::value;
        float2 y = v[bitrev(k1, 2)];
        if constexpr (k1 > 0) y = cmul(y, tw[k1 * 32 + lane]);
        buf[k1 * kFft1024Pitch + lane] = y;
    });
    __syncwarp();
    const float2* row = buf + (lane & 3) * kFft1024Pitch;
    static_for<0, 4>([&](auto ac) {
        constexpr int a = decltype(ac)::value;
        float2 u = row[a];
#pragma unroll
        for (int b = 1; b < 8; ++b) {
            const float2 y = row[a + 4 * b];
            u.x = fmaf(y.x, tq[b].x, fmaf(-y.y, tq[b].y, u.x));
            u.y = fmaf(y.x, tq[b].y, fmaf(y.y, tq[b].x, u.y));
        }
        if constexpr (a == 0) v[a] = u;
        else v[a] = cmul(u, tq[8 + a]);
    });
    __syncwarp();
    fft_reg<4>(v);
}

// Whole-block M-point FFT in shared memory (ping-pong between a and b).  Every thread of the
// group [0, nthreads) calls it with its tid; a leading __syncthreads() is the caller's job
// (data in `a` must be visible).  Returns the buffer that holds the result; ends with a barrier.
template <class C>
__device__ __forceinline__ C* block_fft(C* a, C* b, const C* __restrict__ tw, int log2m, int tid, int nthreads) {
    const int M = 1 << log2m;
    int radix[16];
    const int npass = stockham_schedule(log2m, radix);
    int Ns = 1;
    for (int p = 0; p < npass; ++p) {
        const int R = radix[p];
        for (int j = tid; j < M / R; j += nthreads) {
            if (R == 4) stockham_pass<4>(a, b, tw, M, Ns, j);
            else if (R == 8) stockham_pass<8>(a, b, tw, M, Ns, j);
            else stockham_pass<2>(a, b, tw, M, Ns, j);
        }
        __syncthreads();
        C* t = a;
        a = b;
        b = t;
        Ns *= R;
    }
    return a;
}
#endif

}  // namespace zafb
