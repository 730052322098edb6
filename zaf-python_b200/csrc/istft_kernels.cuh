// The ISTFT warp kernels and their launchers (templates): included by istft.cu (window lengths 256 ... 1024) and
// istft_large.cu (2048, 4096), which instantiate disjoint sets so that the two files compile in parallel.
#pragma once

#include <algorithm>
#include <climits>
#include <cstdint>

#include "stft_common.cuh"

namespace zafb {
namespace {

// ------------------------------------------------------------------------------------------
// ISTFT, N = 2048 or 1024, hop = N / R (R = 2, 4, 8), frame-major spectra: one warp per RUN of
// consecutive output hop-blocks of one clip.
//
// Frame j adds its part q (samples [q hop, (q+1) hop)) into hop-block j + q of the overlap-add
// signal (zaf.py:227-233); block h is complete once frame h has been added.  A warp walks the
// frames of its run in increasing order, keeps the R - 1 unfinished blocks in a private
// shared-memory ring, and writes every finished block once: no atomics, no inter-warp
// synchronisation, and the summation order of every output sample is the reference's (frames
// h-R+1, ..., h).  A run re-reads the R - 1 frames before its first block (warm-up).
//
// Per frame: Re(ifft(X)) for an arbitrary (not necessarily Hermitian) X is the c2r transform of
// H[k] = (X[k] + conj(X[N-k])) / 2; it is packed into ONE 1024-point complex FFT,
//   Z[k] = E[k] + i O[k],  E = H[k] + H[k+1024],  O = (H[k] - H[k+1024]) conj(W_2048^k),
//   y[2n] + i y[2n+1] = conj(FFT_1024(conj(Z)))[n] / (2 N)        (validated in float64).
// ------------------------------------------------------------------------------------------
__host__ __device__ constexpr int istft_ctas_per_sm(int n) { return n == 4096 ? 1 : n <= 256 ? ZAFB_ISTFT_SMALL_CTAS : 2; }
// floats of transpose tile per warp
__host__ __device__ constexpr int istft_tile_floats(int n) { return n == 2048 ? 32 * kFft1024Pitch : 2 * (n / 64) * kFft1024Pitch; }

// MASKED (non-reference extension, instantiated for N = 2048, hop = N/4 only): a real time-frequency mask given for bins
// 0 .. N/2 of every frame (mask_pitch floats apart) and mirrored onto the upper bins like np.concatenate((m, m[-2:0:-1]))
// (zaf.py:185-186) is applied while loading -- m[k] X[k] + conj(m[k] X[N-k]) = m[k] (X[k] + conj X[N-k]) -- so the
// stft -> mask -> istft chain needs no pass that rewrites the spectrum.
template <int N, int R, int WARPS, bool ONESIDED = false, bool MASKED = false>
__global__ void __launch_bounds__(WARPS * 32, istft_ctas_per_sm(N))
istft_warp_kernel(const float2* __restrict__ spec, int64_t nt, const float2* __restrict__ tw4,
                  const float2* __restrict__ tw_full, float scale, int64_t runs_per_clip, int run_len,
                  int64_t total_runs, float* __restrict__ y, int64_t y_stride, int prefetch, int64_t spec_pitch_arg,
                  const float* __restrict__ mask, int64_t mask_pitch) {
    // ONESIDED (non-reference extension): only bins 0 .. N/2 are given, spec_pitch_arg complex elements apart, the rest is
    // their Hermitian mirror -- half the reads.  The two-sided spectrum has a compile-time frame pitch of N.
    const int64_t spec_pitch = ONESIDED ? spec_pitch_arg : int64_t(N);
    constexpr bool onesided = ONESIDED;
    using G = WarpGeom<N>;
    constexpr int M = G::M, REGS = G::REGS, LOGR = G::LOGR;
    constexpr int HOP = N / R;
    constexpr int K = REGS / R;          // registers (float2) per part
    static_assert(K >= 1, "hop too small for this window length");
    constexpr int SLOTS = R - 1;
    constexpr int RING = SLOTS * (HOP / 2);  // float2 per warp
    // transpose tile per warp: N = 2048 uses the split (float) tile of warp_fft1024<true>, N = 1024 the float2 tile of
    // warp_fft512 -- REGS * pitch * 4 bytes * (N == 2048 ? 1 : 2) = the same 4224 bytes either way
    constexpr int TILE_FLOATS = istft_tile_floats(N);
    extern __shared__ float2 smem[];
    float2* s_tw = smem;  // M
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    float2* s_ring = smem + M + warp * RING;
    float* s_buf = reinterpret_cast<float*>(smem + M + WARPS * RING) + warp * TILE_FLOATS;

    for (int i = tid; i < M; i += WARPS * 32) s_tw[i] = tw4[i];
    const float2 c_lane = tw_full[lane];  // W_N^lane
    LaneTw<N> lt;
    lt.init(lane);
    __syncthreads();

    for (int64_t task = int64_t(blockIdx.x) * WARPS + warp; task < total_runs;
         task += int64_t(gridDim.x) * WARPS) {
        const int64_t clip = task / runs_per_clip;
        const int64_t run = task - clip * runs_per_clip;
        const int64_t h_begin = (R - 1) + run * run_len;  // first finished block of the run (OLA coordinates)
        int64_t h_end = h_begin + run_len;
        if (h_end > nt) h_end = nt;
        for (int i = lane; i < RING; i += 32) s_ring[i] = make_float2(0.f, 0.f);
        __syncwarp();
        float* yc = y + clip * y_stride;
        int slot0 = int((h_begin - (R - 1)) % SLOTS);  // ring slot of block j

        for (int64_t j = h_begin - (R - 1); j < h_end; ++j) {
            const float2* X = spec + (clip * nt + j) * spec_pitch;
            const float* mrow = MASKED ? mask + (clip * nt + j) * mask_pitch : nullptr;
            if (prefetch && j + 1 < h_end) {  // the next frame of this run towards L2: N * 8 / 128 lines, N / 512 per lane
#pragma unroll
                for (int i = 0; i < N / 512; ++i)  // (N = 512: one line per lane; one-sided frames are half as long)
                    if (!onesided || (lane + 32 * i) * 16 <= M) prefetch_l2(X + spec_pitch + (lane + 32 * i) * 16);
            }
            float2 v[REGS];
            // r and REGS - 1 - r back to back: the mirrored loads (c, d) of one hit the lines the direct
            // loads (a, b) of the other have just brought into L1
            static_for<0, REGS>([&](auto tc) {
                constexpr int t = decltype(tc)::value;
                constexpr int r = (t % 2 == 0) ? t / 2 : REGS - 1 - t / 2;
                const int k = lane + 32 * r;
                float2 h0, h1;
                if constexpr (onesided) {  // X[N - k] = conj(X[k]), X[M + k] = conj(X[M - k]); k = 0 pairs X[0] and X[M] with themselves
                    const float2 a = __ldg(X + k);
                    const float2 c = __ldg(X + M - k);
                    h0 = make_float2(2.f * a.x, k == 0 ? 0.f : 2.f * a.y);
                    h1 = make_float2(2.f * c.x, k == 0 ? 0.f : -2.f * c.y);
                } else {
                    const float2 a = __ldg(X + k);
                    const float2 b = __ldg(X + M + k);
                    const float2 c = __ldg(X + M - k);
                    const float2 d = __ldg(X + ((N - k) & (N - 1)));
                    h0 = make_float2(a.x + d.x, a.y - d.y);  // 2 H[k]
                    h1 = make_float2(b.x + c.x, b.y - c.y);  // 2 H[k + M]
                }
                if constexpr (MASKED) {  // bins k and N - k carry m[k]; bins M + k and M - k carry m[M - k]
                    const float m0 = __ldg(mrow + k), m1 = __ldg(mrow + M - k);
                    h0 = make_float2(h0.x * m0, h0.y * m0);
                    h1 = make_float2(h1.x * m1, h1.y * m1);
                }
                const float2 e = cadd(h0, h1);
                const float2 o = cmul_conj(csub(h0, h1), mul_tw<r, N / 32>(c_lane));
                // conj(Z) = conj(e + i o)
                v[r] = make_float2(e.x - o.y, -(e.y + o.x));
            });

            // conj(z[lane + 32 k2]) = v[bitrev(k2)]
            if constexpr (N == 4096) warp_fft2048(v, s_tw, reinterpret_cast<float2*>(s_buf), lane);
            else if constexpr (N == 2048) warp_fft1024<true>(v, s_tw, s_buf, lane);
            else if constexpr (N == 1024) warp_fft512(v, s_tw, reinterpret_cast<float2*>(s_buf), lane);
            else if constexpr (N == 512) warp_fft256(v, s_tw, reinterpret_cast<float2*>(s_buf), lane, lt.tq);
            else warp_fft128(v, s_tw, reinterpret_cast<float2*>(s_buf), lane, lt.tq);

            static_for<0, REGS>([&](auto k2c) {
                constexpr int k2 = decltype(k2c)::value;
                constexpr int q = k2 / K;   // part of the frame
                constexpr int i = k2 % K;
                const float2 z = make_float2(v[bitrev(k2, LOGR)].x, -v[bitrev(k2, LOGR)].y);
                if constexpr (q == 0) {
                    float2 acc = z;
                    if constexpr (R > 1) {
                        const float2 prev = s_ring[slot0 * (HOP / 2) + lane + 32 * i];
                        acc = make_float2(prev.x + z.x, prev.y + z.y);
                    }
                    if (j >= h_begin) {
                        float2* dst = reinterpret_cast<float2*>(yc + (j - (R - 1)) * HOP) + lane + 32 * i;
                        *dst = make_float2(acc.x * scale, acc.y * scale);
                    }
                } else if constexpr (q == R - 1) {
                    s_ring[slot0 * (HOP / 2) + lane + 32 * i] = z;  // block j + R - 1 starts in the slot block j left
                } else {
                    int s = slot0 + q;
                    if (s >= SLOTS) s -= SLOTS;
                    float2* cell = s_ring + s * (HOP / 2) + lane + 32 * i;
                    const float2 prev = *cell;
                    *cell = make_float2(prev.x + z.x, prev.y + z.y);
                }
            });
            slot0 = (slot0 + 1 == SLOTS) ? 0 : slot0 + 1;
        }
        __syncwarp();
    }
}


template <int N, int R, int WARPS>
int launch_istft_warp_w(const zafb_stft_plan* p, const float2* spec, int64_t n_clips, int64_t nt, float* y,
                        int64_t y_stride, cudaStream_t st, int64_t spec_pitch, int onesided, const float* mask = nullptr,
                        int64_t mask_pitch = 0) {
    constexpr bool kHasMasked = N == 2048 && R == 4 && WARPS == 8;  // the masked kernels exist for the BASELINE geometry
    if (mask != nullptr && !kHasMasked)
        return fail(ZAFB_E_UNSUPPORTED, "masked istft: fused kernel exists for window_length 2048, hop 512 only");
    static bool attr = false;
    if (!attr) {
        ZAFB_CUDA((cudaFuncSetAttribute(istft_warp_kernel<N, R, WARPS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
        ZAFB_CUDA((cudaFuncSetAttribute(istft_warp_kernel<N, R, WARPS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
        if constexpr (kHasMasked) {
            ZAFB_CUDA((cudaFuncSetAttribute(istft_warp_kernel<N, R, WARPS, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
            ZAFB_CUDA((cudaFuncSetAttribute(istft_warp_kernel<N, R, WARPS, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
        }
        attr = true;
    }
    const int64_t nblocks = nt - (R - 1);  // finished hop-blocks per clip == output length / hop
    constexpr int kCtasPerSm = istft_ctas_per_sm(N);
    const int64_t resident_warps = int64_t(sm_count()) * kCtasPerSm * WARPS;
    // run length: minimise (runs per warp) x (frames per run, warm-up included)
    int64_t best_len = nblocks, best_cost = INT64_MAX;
    for (int64_t len = nblocks < 8 ? nblocks : 8; len <= nblocks && len <= 1024; ++len) {
        const int64_t runs = n_clips * ceil_div(nblocks, len);
        const int64_t cost = ceil_div(runs, resident_warps) * (len + R - 1);
        if (cost < best_cost || (cost == best_cost && len > best_len)) {
            best_cost = cost;
            best_len = len;
        }
    }
    const int64_t runs_per_clip = ceil_div(nblocks, best_len);
    const int64_t total = n_clips * runs_per_clip;
    int64_t ctas = ceil_div(total, WARPS);
    if (ctas > int64_t(sm_count()) * kCtasPerSm) ctas = int64_t(sm_count()) * kCtasPerSm;
    constexpr int HOP = N / R;
    constexpr size_t smem = (N / 2) * sizeof(float2) + size_t(WARPS) * ((R - 1) * (HOP / 2) * sizeof(float2) +
                                                                               istft_tile_floats(N) * sizeof(float));
    static_assert(smem <= size_t(kMaxDynSmem), "istft warp kernel: shared memory");
    const float scale = static_cast<float>(1.0 / (2.0 * double(N) * p->gain));
    auto kern = onesided ? istft_warp_kernel<N, R, WARPS, true> : istft_warp_kernel<N, R, WARPS, false>;
    if constexpr (kHasMasked) {
        if (mask != nullptr) kern = onesided ? istft_warp_kernel<N, R, WARPS, true, true> : istft_warp_kernel<N, R, WARPS, false, true>;
    }
    kern<<<static_cast<unsigned>(ctas), WARPS * 32, smem, st>>>(
        spec, nt, p->d_tw_4step, p->d_tw_full, scale, runs_per_clip, int(best_len), total, y, y_stride,
        env_flag("ZAFB_ISTFT_PREFETCH", 1), spec_pitch, mask, mask_pitch);
    ZAFB_LAUNCH_CHECK();
    return ZAFB_OK;
}

template <int N, int R>
int launch_istft_warp(const zafb_stft_plan* p, const float2* spec, int64_t n_clips, int64_t nt, float* y,
                      int64_t y_stride, cudaStream_t st, int64_t spec_pitch, int onesided, const float* mask = nullptr,
                      int64_t mask_pitch = 0) {
    // N = 4096: 6 warps (a 17 KB transpose tile plus up to 14 KB of overlap-add ring per warp)
    if constexpr (N == 4096) return launch_istft_warp_w<N, R, 6>(p, spec, n_clips, nt, y, y_stride, st, spec_pitch, onesided, mask, mask_pitch);
    else {
        if (env_flag("ZAFB_ISTFT_WARPS", 8) == 6 && mask == nullptr)
            return launch_istft_warp_w<N, R, 6>(p, spec, n_clips, nt, y, y_stride, st, spec_pitch, onesided);
        return launch_istft_warp_w<N, R, 8>(p, spec, n_clips, nt, y, y_stride, st, spec_pitch, onesided, mask, mask_pitch);
    }
}

// ------------------------------------------------------------------------------------------
// ISTFT reading BIN_MAJOR memory [clip][bin][frame] -- the reference's C order (zaf.py:214: the (N, nt) array zaf.stft
// returns) -- directly, instead of a tiled transpose into frame-major scratch (three passes over the spectrum).
//
// A CTA of 16 warps walks along one clip in tiles of F = 16 consecutive frames.  Load phase: thread (u, w) reads frame
// w of bin rows k = u, u + 32, ... and of their mirror rows N - k -- 16 frames of a row are one 128-byte run -- and parks
// Hs[k] = X[k] + conj(X[N-k]), k = 0 .. N/2, in the frame's slot of a ring of F + R - 1 shared-memory regions (the only
// combination of the two-sided spectrum that Re(ifft) depends on, zaf.py:223).  Transform phase: warp w runs the packed
// c2r transform of istft_warp_kernel on its frame (the slot doubles as the FFT's transpose tile, then receives the N
// time samples).  Overlap-add phase: every finished hop-block is the sum of R frame parts, added in increasing frame
// order like the reference (zaf.py:227-233) and like istft_warp_kernel -- the results are bit-identical to the
// frame-major path -- and leaves as contiguous 16-byte stores.  The R - 1 last frames of a tile stay in the ring.
// ------------------------------------------------------------------------------------------
#ifndef ZAFB_ISTFT_BM_LOAD_BATCH
#define ZAFB_ISTFT_BM_LOAD_BATCH 8  // cfg 2: 4 -> 6.2 ms, 8 -> 5.38 ms, 16 -> 5.98 ms (spills)
#endif
template <int N>
struct IstftBinMajorGeom {
    static constexpr int F = 16;
    static constexpr int M = N / 2;
    static constexpr int TILE = N == 2048 ? (32 * kFft1024Pitch + 1) / 2 : (N / 64) * kFft1024Pitch;  // float2 units
    static constexpr int NEED = (M + 1) > TILE ? (M + 1) : TILE;
    static constexpr int PITCH = NEED | 1;  // odd: the 16 frames of one bin land in distinct bank pairs
};

template <int N, int R>
__global__ void __launch_bounds__(512, 1)
istft_binmajor_kernel(const float2* __restrict__ spec, int nt, const float2* __restrict__ tw4,
                      const float2* __restrict__ tw_full, float scale, float* __restrict__ y, int64_t y_stride,
                      int runs_per_clip, int tiles_per_run, int64_t total_runs, int prefetch) {
    using G = WarpGeom<N>;
    using B = IstftBinMajorGeom<N>;
    constexpr int M = G::M, REGS = G::REGS, LOGR = G::LOGR;
    constexpr int F = B::F, SLOTS = F + R - 1, PITCH = B::PITCH, HOP = N / R;
    extern __shared__ __align__(16) float2 smem[];
    float2* s_tw = smem;          // M: W_M^{k1 n2}
    float2* s_reg = smem + M;     // SLOTS regions of PITCH float2
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < M; i += 512) s_tw[i] = tw4[i];
    const float2 c_lane = tw_full[lane];  // W_N^lane
    LaneTw<N> lt;
    lt.init(lane);
    const int lw = tid & (F - 1), lu = tid / F;  // load phase: frame within the tile, bin index mod 32
    const int tiles_per_clip = (nt + F - 1) / F;
    __syncthreads();

    for (int64_t run = blockIdx.x; run < total_runs; run += gridDim.x) {
        const int64_t clip = run / runs_per_clip;
        const int t0 = int(run - clip * runs_per_clip) * tiles_per_run;
        const int t1 = min(t0 + tiles_per_run, tiles_per_clip);
        const float2* sc = spec + clip * int64_t(N) * nt;
        float* yc = y + clip * y_stride;
        // a run that starts inside a clip first transforms the R - 1 frames before it (no output: they belong to the
        // previous run's blocks, but their later parts overlap into this run's first blocks)
        for (int t = t0 > 0 ? t0 - 1 : t0; t < t1; ++t) {
            const bool warmup = t < t0;
            const int j0 = warmup ? t0 * F - (R - 1) : t * F;
            const int cnt = warmup ? R - 1 : F;
            // ---- load: Hs[k] = X[k] + conj(X[N - k]) for the tile's frames
            {
                const int j = j0 + lw;
                if (lw < cnt && j < nt) {
                    float2* slot = s_reg + (j % SLOTS) * PITCH;
                    const float2* col = sc + j;
                    // kLoadBatch bins = 2 kLoadBatch independent 8-byte loads in flight per thread before the first use
                    // (the phase waits on HBM latency: 4 bins at a time 6.2 ms on cfg 2)
                    constexpr int KB = ZAFB_ISTFT_BM_LOAD_BATCH;
                    int k = lu;
#pragma unroll 1
                    for (; k + 32 * (KB - 1) <= M; k += 32 * KB) {
                        float2 a[KB], d[KB];
#pragma unroll
                        for (int b = 0; b < KB; ++b) {
                            a[b] = __ldg(col + int64_t(k + 32 * b) * nt);
                            d[b] = __ldg(col + int64_t((N - k - 32 * b) & (N - 1)) * nt);
                        }
#pragma unroll
                        for (int b = 0; b < KB; ++b) slot[k + 32 * b] = make_float2(a[b].x + d[b].x, a[b].y - d[b].y);
                    }
                    for (; k <= M; k += 32) {
                        const float2 a = __ldg(col + int64_t(k) * nt);
                        const float2 d = __ldg(col + int64_t((N - k) & (N - 1)) * nt);
                        slot[k] = make_float2(a.x + d.x, a.y - d.y);
                    }
                }
            }
            __syncthreads();
            // the next tile's 128-byte runs (one per bin row and mirror row, up to two L2 lines each) start their way
            // to L2 now, so that its load phase finds them there instead of waiting on HBM after the transform
            if (prefetch && t + 1 < t1) {
                const int jn = (t + 1) * F;
                const int jl = jn + F - 1 < nt ? jn + F - 1 : nt - 1;
                for (int idx = tid; idx < 2 * (M + 1); idx += 512) {
                    const int k = idx >> 1;
                    const int64_t row = (idx & 1) ? int64_t((N - k) & (N - 1)) : int64_t(k);
                    prefetch_l2(sc + row * nt + jn);
                    prefetch_l2(sc + row * nt + jl);
                }
            }
            // ---- transform: warp w, frame j0 + w
            {
                const int j = j0 + warp;
                if (warp < cnt && j < nt) {  // warp-uniform
                    float2* slot = s_reg + (j % SLOTS) * PITCH;
                    float2 v[REGS];
                    static_for<0, REGS>([&](auto tc) {
                        constexpr int r = decltype(tc)::value;
                        const int k = lane + 32 * r;
                        const float2 h0 = slot[k];                    // 2 H[k]
                        const float2 hm = slot[M - k];
                        const float2 h1 = make_float2(hm.x, -hm.y);   // 2 H[k + M] = conj(Hs[M - k])
                        const float2 e = cadd(h0, h1);
                        const float2 o = cmul_conj(csub(h0, h1), mul_tw<r, N / 32>(c_lane));
                        v[r] = make_float2(e.x - o.y, -(e.y + o.x));  // conj(Z), Z = E + i O
                    });
                    __syncwarp();  // every lane has read the spectrum: the slot becomes the transpose tile
                    if constexpr (N == 2048) warp_fft1024<true>(v, s_tw, reinterpret_cast<float*>(slot), lane);
                    else if constexpr (N == 1024) warp_fft512(v, s_tw, slot, lane);
                    else if constexpr (N == 512) warp_fft256(v, s_tw, slot, lane, lt.tq);
                    else warp_fft128(v, s_tw, slot, lane, lt.tq);
                    __syncwarp();
                    static_for<0, REGS>([&](auto k2c) {   // y[2n] + i y[2n+1] = conj(v[bitrev(k2)]), n = lane + 32 k2
                        constexpr int k2 = decltype(k2c)::value;
                        slot[lane + 32 * k2] = make_float2(v[bitrev(k2, LOGR)].x, -v[bitrev(k2, LOGR)].y);
                    });
                }
            }
            __syncthreads();
            // ---- overlap-add: hop-blocks h in [j0, j0 + F) that exist (h >= R - 1, h < nt), two samples per thread
            if (!warmup) {
                const int h_lo = j0 > R - 1 ? j0 : R - 1;
                const int h_hi = (j0 + F < nt ? j0 + F : nt);
                const int pairs = (h_hi - h_lo) * (HOP / 2);
                for (int pr = tid; pr < pairs; pr += 512) {
                    const int h = h_lo + pr / (HOP / 2);
                    const int i2 = pr % (HOP / 2);
                    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                    for (int q = R - 1; q >= 0; --q) {  // frames h - R + 1 ... h, in the reference's order
                        const float2 part = s_reg[((h - q) % SLOTS) * PITCH + q * (HOP / 2) + i2];
                        if (q == R - 1) acc = part;
                        else acc = make_float2(acc.x + part.x, acc.y + part.y);
                    }
                    reinterpret_cast<float2*>(yc + int64_t(h - (R - 1)) * HOP)[i2] = make_float2(acc.x * scale, acc.y * scale);
                }
                __syncthreads();
            }
        }
    }
}

template <int N, int R>
int launch_istft_binmajor(const zafb_stft_plan* p, const float2* spec, int64_t n_clips, int64_t nt, float* y, int64_t y_stride,
                          cudaStream_t st) {
    using B = IstftBinMajorGeom<N>;
    constexpr size_t smem = (size_t(N / 2) + size_t(B::F + R - 1) * B::PITCH) * sizeof(float2);
    static_assert(smem <= size_t(kMaxDynSmem), "istft bin-major kernel: shared memory");
    static bool attr = false;
    if (!attr) {
        ZAFB_CUDA((cudaFuncSetAttribute(istft_binmajor_kernel<N, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
        attr = true;
    }
    // runs of consecutive tiles of one clip: whole clips when there are enough of them, else about two runs per SM
    const int64_t sms = sm_count();
    const int64_t tiles_per_clip = ceil_div(nt, B::F);
    int64_t runs_per_clip = n_clips >= 2 * sms ? 1 : std::min<int64_t>(tiles_per_clip, ceil_div(2 * sms, n_clips));
    if (const int forced = env_flag("ZAFB_ISTFT_BM_RUNS_PER_CLIP", 0); forced > 0) runs_per_clip = std::min<int64_t>(tiles_per_clip, forced);  // tests
    const int64_t tiles_per_run = ceil_div(tiles_per_clip, runs_per_clip);
    runs_per_clip = ceil_div(tiles_per_clip, tiles_per_run);
    const int64_t runs = n_clips * runs_per_clip;
    const float scale = static_cast<float>(1.0 / (2.0 * double(N) * p->gain));
    const int64_t ctas = std::min<int64_t>(runs, sms);
    istft_binmajor_kernel<N, R><<<unsigned(ctas), 512, smem, st>>>(spec, int(nt), p->d_tw_4step, p->d_tw_full, scale, y, y_stride,
                                                                    int(runs_per_clip), int(tiles_per_run), runs,
                                                                    env_flag("ZAFB_ISTFT_BM_PREFETCH", 0));
    ZAFB_LAUNCH_CHECK();
    return ZAFB_OK;
}

}  // namespace
}  // namespace zafb
