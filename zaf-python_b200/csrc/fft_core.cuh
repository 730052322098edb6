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

#include "common.cuh"

namespace zafb {

// ---------------------------------------------------------------- in-register FFT
template <int N, int J, int STRIDE>
struct DifButterflies {
    static ZAFB_HD void run(float2* v) {
        // span N/2 butterflies of one radix-2 DIF stage on v[0], v[STRIDE], ...
        float2 a = v[J * STRIDE];
        float2 b = v[(J + N / 2) * STRIDE];
        v[J * STRIDE] = cadd(a, b);
        v[(J + N / 2) * STRIDE] = mul_tw<J, N>(csub(a, b));
        if constexpr (J + 1 < N / 2) DifButterflies<N, J + 1, STRIDE>::run(v);
    }
};

// In-place DIF FFT of v[0], v[STRIDE], ..., v[(N-1)*STRIDE]; output bit-reversed.
template <int N, int STRIDE = 1>
ZAFB_HD void fft_reg(float2* v) {
    if constexpr (N >= 2) {
        DifButterflies<N, 0, STRIDE>::run(v);
        fft_reg<N / 2, STRIDE>(v);
        fft_reg<N / 2, STRIDE>(v + (N / 2) * STRIDE);
    }
}

// ---------------------------------------------------------------- Stockham pass (shared memory)
// One radix-R butterfly j (0 <= j < M/R) of the pass with sub-transform length Ns (product of the
// radices of the previous passes).  tw[t] = exp(-2 pi i t / M), t < M.  Natural order in and out.
template <int R>
ZAFB_HD void stockham_pass(const float2* __restrict__ in, float2* __restrict__ out,
                           const float2* __restrict__ tw, int M, int Ns, int j) {
    const int k = j & (Ns - 1);
    const int tstride = M / (Ns * R);
    float2 v[R];
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
// Whole-block M-point FFT in shared memory (ping-pong between a and b).  Every thread of the
// group [0, nthreads) calls it with its tid; a leading __syncthreads() is the caller's job
// (data in `a` must be visible).  Returns the buffer that holds the result; ends with a barrier.
__device__ __forceinline__ float2* block_fft(float2* a, float2* b, const float2* __restrict__ tw,
                                             int log2m, int tid, int nthreads) {
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
        float2* t = a;
        a = b;
        b = t;
        Ns *= R;
    }
    return a;
}
#endif

}  // namespace zafb
