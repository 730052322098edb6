// STFT / ISTFT kernels (zaf.py:45-141, 144-243).
//
//   stft_warp_kernel<N>          the north-star path (N = 2048; also 256, 512, 1024, 4096): one warp per frame.  The
//                                frame's N real samples are read as N/2 complex points straight into registers
//                                (coalesced 8-byte loads), windowed, transformed by a four-step FFT (in-register
//                                radix-2 FFTs around one shared-memory transpose), unpacked to the real-input spectrum
//                                with warp shuffles and stored as the full two-sided spectrum, frame-major, with
//                                coalesced streaming stores.
//   stft_warp_binmajor_kernel<N> the same transform written directly into the reference's C-order memory
//                                [clip][bin][frame]: 16-frame tiles in a shared-memory ring, per-row sector-aligned
//                                store windows.
//   istft_warp_kernel<N, R>      hop = N / R: one warp per run of consecutive output hop-blocks, overlap-add in a
//                                private shared-memory ring, every sample written once in the reference's order.
//   stft_generic_kernel          any power-of-two N >= 2: one CTA per frame, Stockham FFT of N/2 complex points in
//                                shared memory.
//   stft_dft_kernel              any other N (the reference accepts every N): direct O(N^2) DFT.
//   istft_tile_kernel            overlap-add in gather form: a CTA owns a tile of output samples, inverse transforms
//                                every frame that touches it in increasing frame order (the reference's summation
//                                order, zaf.py:227-233) and writes each sample once: no atomics, bit-reproducible.
//   stft_host_mirrored           host-buffer pipeline that copies bins 0 .. N/2 only and lets host threads write the
//                                Hermitian mirror.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <atomic>
#include <thread>
#include <type_traits>
#include <vector>

#include <sched.h>

#include "fft_core.cuh"
#include "host_pipe.cuh"
#include "transpose.cuh"
#include "stft_common.cuh"

namespace zafb {
void host_mirror_fill_frames(void* out, int64_t frames, int64_t n);                                  // host_mirror.cpp
void host_mirror_fill_rows(void* out, int64_t nt, int64_t n, int64_t r_lo, int64_t r_hi);
}  // namespace zafb
using namespace zafb;

namespace {

// ------------------------------------------------------------------------------------------
// N = 256 ... 4096: one warp per frame
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ void st_stream(float2* p, float2 v) { __stcs(p, v); }

// shared -> global bulk copy (TMA engine), tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(static_cast<uint32_t>(__cvta_generic_to_shared(ssrc))), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int PENDING>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PENDING) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// N = 2048 or 1024, one warp per frame.
// BULK = true (N = 2048 only): the spectrum leaves through shared memory and four 4 KB cp.async.bulk stores per frame
// (issued by one lane, executed by the TMA engine) instead of 64 st.global per lane.
template <int N, bool BULK, int WARPS, bool ONESIDED = false>
__global__ void __launch_bounds__(WARPS * 32, WarpGeom<N>::CTAS_PER_SM)
stft_warp_kernel(const float* __restrict__ x, int64_t ns, int64_t clip_stride, int64_t nt, int hop,
                 const float2* __restrict__ win_half, const float2* __restrict__ tw4,
                 const float2* __restrict__ tw_full, float2* __restrict__ out, int64_t total_frames, int sequential,
                 int64_t out_pitch_arg) {
    // ONESIDED (non-reference extension, runs the `sequential` store order): only bins 0 .. N/2 are stored, out_pitch_arg
    // complex elements apart; the reference's two-sided spectrum has a compile-time frame pitch of N.
    const int64_t out_pitch = ONESIDED ? out_pitch_arg : int64_t(N);
    constexpr bool onesided = ONESIDED;
    using G = WarpGeom<N>;
    constexpr int M = G::M, REGS = G::REGS, LOGR = G::LOGR;
    static_assert(!BULK || N == 2048, "the bulk-store variant is written for N = 2048");
    extern __shared__ __align__(128) float2 smem[];
    float2* s_win = smem;       // M: (0.5 w[2n], 0.5 w[2n+1])
    float2* s_tw = smem + M;    // M: W_M^{k1*n2} at [k1*32+n2]
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    float2* s_buf = smem + 2 * M + warp * G::TILE;

    for (int i = tid; i < M; i += WARPS * 32) {
        s_win[i] = win_half[i];
        s_tw[i] = tw4[i];
    }
    const float2 c_lane = tw_full[lane];  // W_N^lane
    LaneTw<N> lt;
    lt.init(lane);
    __syncthreads();

    for (int64_t f = int64_t(blockIdx.x) * WARPS + warp; f < total_frames; f += int64_t(gridDim.x) * WARPS) {
        const int64_t clip = f / nt;
        const int64_t j = f - clip * nt;
        const int64_t start = j * hop - M;  // first sample of the frame (may be < 0): centre padding floor(N/2), zaf.py:99
        const float* xc = x + clip * clip_stride;

        float2 v[REGS];
        if (start >= 0 && start + N <= ns) {
            const float2* p = reinterpret_cast<const float2*>(xc + start) + lane;
#pragma unroll
            for (int r = 0; r < REGS; ++r) v[r] = __ldg(p + 32 * r);
        } else {
#pragma unroll
            for (int r = 0; r < REGS; ++r) {
                const int64_t s = start + 2 * (lane + 32 * r);
                v[r].x = (s >= 0 && s < ns) ? __ldg(xc + s) : 0.f;
                v[r].y = (s + 1 >= 0 && s + 1 < ns) ? __ldg(xc + s + 1) : 0.f;
            }
        }
#pragma unroll
        for (int r = 0; r < REGS; ++r) {
            const float2 w = s_win[lane + 32 * r];
            v[r].x *= w.x;
            v[r].y *= w.y;
        }

        warp_fft_half<N>(v, s_tw, s_buf, lane, lt);  // Z[lane + 32 k2] = v[bitrev(k2)]

        // real-input unpack: X[k] = E + w_k O, X[k+M] = E - w_k O with
        //   E = Z[k] + conj(Z[M-k]),  O = -i (Z[k] - conj(Z[M-k])),  w_k = W_N^k   (the 1/2 is in the window)
        // computed for k < M/2 only; the other half of the two-sided spectrum is its conjugate mirror,
        // X[M-k] = conj(X[k+M]) and X[N-k] = conj(X[k]), stored by the same lane (a warp still writes
        // 32 consecutive bins per instruction, in descending lane order).  k = M/2 is its own mirror (lane 0).
        const int src = (32 - lane) & 31;
        if constexpr (BULK) {
            const float2 z512 = v[bitrev(16, 5)];
            // descending k2: the results X[k] and X[k+1024] replace the two registers iteration k2 has just consumed
            // (lane 0 reads register 32 - k2, which the earlier iterations have not touched yet)
            static_for<0, 16>([&](auto ic) {
                constexpr int k2 = 15 - decltype(ic)::value;
                const float2 z = v[bitrev(k2, 5)];
                const float2 mine = v[bitrev(31 - k2, 5)];
                float2 p;
                p.x = __shfl_sync(0xffffffffu, mine.x, src);
                p.y = __shfl_sync(0xffffffffu, mine.y, src);
                if (lane == 0) p = v[bitrev((32 - k2) & 31, 5)];
                const float2 e = make_float2(z.x + p.x, z.y - p.y);
                const float2 od = make_float2(z.y + p.y, p.x - z.x);
                const float2 t = cmul(mul_tw<k2, 64>(c_lane), od);
                v[bitrev(k2, 5)] = cadd(e, t);       // X[k]
                v[bitrev(31 - k2, 5)] = csub(e, t);  // X[k + 1024]
            });
            // quarters of the frame: [0,512) = X[k]; [512,1024) = conj(X[k+1024]) mirrored; [1024,1536) = X[k+1024];
            // [1536,2048) = conj(X[k]) mirrored.  Two 4 KB halves of the warp's transpose tile ping-pong.
            float2* q0 = s_buf;
            float2* q1 = s_buf + 512;
            float2* g = out + f * out_pitch;
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) q0[lane + 32 * k2] = v[bitrev(k2, 5)];
            fence_async_smem();
            __syncwarp();
            if (lane == 0) bulk_store(g, q0, 4096);
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2)
                if (k2 > 0 || lane != 0) q1[512 - lane - 32 * k2] = cconj(v[bitrev(31 - k2, 5)]);
            if (lane == 0) q1[0] = make_float2(2.f * z512.x, -2.f * z512.y);  // X[512]: Z[512] pairs with itself, w = -i
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                bulk_store(g + 512, q1, 4096);
                bulk_wait_read<1>();  // q0 has been read
            }
            __syncwarp();
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) q0[lane + 32 * k2] = v[bitrev(31 - k2, 5)];
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                bulk_store(g + 1024, q0, 4096);
                bulk_wait_read<1>();  // q1 has been read
            }
            __syncwarp();
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2)
                if (k2 > 0 || lane != 0) q1[512 - lane - 32 * k2] = cconj(v[bitrev(k2, 5)]);
            if (lane == 0) q1[0] = make_float2(2.f * z512.x, 2.f * z512.y);   // X[1536]
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                bulk_store(g + 1536, q1, 4096);
                bulk_wait_read<0>();  // the tile is free again before the next frame's transpose
            }
            __syncwarp();
            continue;
        }
        float2* const of = out + f * out_pitch;
        float2* o = of + lane;
        float2* om = of + M - lane;
        if (ONESIDED || sequential) {
            // the same unpack, but the results first replace the registers they were computed from (descending k2,
            // see the bulk variant), then the frame is written in ascending address order, quarter by quarter
            const float2 zmid = v[bitrev(REGS / 2, LOGR)];
            static_for<0, REGS / 2>([&](auto ic) {
                constexpr int k2 = REGS / 2 - 1 - decltype(ic)::value;
                const float2 z = v[bitrev(k2, LOGR)];
                const float2 mine = v[bitrev(REGS - 1 - k2, LOGR)];
                float2 p;
                p.x = __shfl_sync(0xffffffffu, mine.x, src);
                p.y = __shfl_sync(0xffffffffu, mine.y, src);
                if (lane == 0) p = v[bitrev((REGS - k2) & (REGS - 1), LOGR)];
                const float2 e = make_float2(z.x + p.x, z.y - p.y);
                const float2 od = make_float2(z.y + p.y, p.x - z.x);
                const float2 t = cmul(mul_tw<k2, N / 32>(c_lane), od);
                v[bitrev(k2, LOGR)] = cadd(e, t);             // X[k]
                v[bitrev(REGS - 1 - k2, LOGR)] = csub(e, t);  // X[k + M]
            });
            static_for<0, REGS / 2>([&](auto kc) {            // [0, M/2): X[k]
                constexpr int k2 = decltype(kc)::value;
                st_stream(o + 32 * k2, v[bitrev(k2, LOGR)]);
            });
            if (lane == 0) st_stream(of + M / 2, make_float2(2.f * zmid.x, -2.f * zmid.y));
            static_for<0, REGS / 2>([&](auto ic) {            // (M/2, M): conj(X[k + M]) mirrored, ascending addresses
                constexpr int k2 = REGS / 2 - 1 - decltype(ic)::value;
                if (k2 > 0 || lane != 0) st_stream(om - 32 * k2, cconj(v[bitrev(REGS - 1 - k2, LOGR)]));
            });
            if constexpr (onesided) {                         // bins 0 .. M only: the Nyquist bin X[M] closes the frame
                if (lane == 0) st_stream(of + M, v[bitrev(REGS - 1, LOGR)]);
                continue;
            }
            static_for<0, REGS / 2>([&](auto kc) {            // [M, 3M/2): X[k + M]
                constexpr int k2 = decltype(kc)::value;
                st_stream(o + M + 32 * k2, v[bitrev(REGS - 1 - k2, LOGR)]);
            });
            if (lane == 0) st_stream(of + M + M / 2, make_float2(2.f * zmid.x, 2.f * zmid.y));
            static_for<0, REGS / 2>([&](auto ic) {            // (3M/2, N): conj(X[k]) mirrored
                constexpr int k2 = REGS / 2 - 1 - decltype(ic)::value;
                if (k2 > 0 || lane != 0) st_stream(om + M - 32 * k2, cconj(v[bitrev(k2, LOGR)]));
            });
            continue;
        }
        static_for<0, REGS / 2>([&](auto k2c) {
            constexpr int k2 = decltype(k2c)::value;
            const float2 z = v[bitrev(k2, LOGR)];
            const float2 mine = v[bitrev(REGS - 1 - k2, LOGR)];
            float2 p;
            p.x = __shfl_sync(0xffffffffu, mine.x, src);
            p.y = __shfl_sync(0xffffffffu, mine.y, src);
            if (lane == 0) p = v[bitrev((REGS - k2) & (REGS - 1), LOGR)];
            const float2 e = make_float2(z.x + p.x, z.y - p.y);
            const float2 od = make_float2(z.y + p.y, p.x - z.x);
            const float2 w = mul_tw<k2, N / 32>(c_lane);
            const float2 t = cmul(w, od);
            const float2 lo = cadd(e, t), hi = csub(e, t);
            st_stream(o + 32 * k2, lo);
            st_stream(o + M + 32 * k2, hi);
            if (k2 > 0 || lane != 0) {
                st_stream(om - 32 * k2, cconj(hi));
                st_stream(om + M - 32 * k2, cconj(lo));
            }
        });
        if (lane == 0) {  // k = M/2: Z[M/2] pairs with itself, w = -i
            const float2 z = v[bitrev(REGS / 2, LOGR)];
            st_stream(of + M / 2, make_float2(2.f * z.x, -2.f * z.y));
            st_stream(of + M + M / 2, make_float2(2.f * z.x, 2.f * z.y));
        }
    }
}

// ------------------------------------------------------------------------------------------
// The same transform writing BIN_MAJOR memory [clip][bin][frame] -- the reference's C order (zaf.py:128) -- directly.
//
// A CTA of 16 warps walks along one clip in tiles of F = 16 consecutive frames.  Each warp transforms one frame exactly
// like stft_warp_kernel (same arithmetic, bit-identical values) and parks its M + 1 distinct results
//     R[k] = X[k],  R[M/2 + k] = X[M + k]  (k < M/2),  R[M] = X[M/2]
// in the frame's slot of a ring of F + 3 shared-memory regions (the slot doubles as the FFT's transpose tile).  After a
// barrier the CTA stores the tile bin by bin: thread (u, w) writes frame w of row b1(u) and of the mirror row b2(u)
// (conjugated), so 16 frames of a row leave as one 128-byte run.
//
// Why the ring: rows are nt * 8 bytes long, and for odd nt (939 at cfg 2) a run that starts at a tile boundary starts
// in the middle of a 32-byte sector of most rows.  Partial-sector writes are what makes a transposed store slow on
// this memory system (measured on cfg 2: 16.0 ms with 64-byte runs, 9.1 ms with 128-byte runs; 6.6 / 4.5 ms for the
// same kernels when nt is a multiple of 4, i.e. every run sector-aligned; profiles/r01n_bm_probe.log).  So every row
// gets its own window: row b writes frames [j0 - s, j0 + 16 - s) with s = (address of element (b, j0) / 8) mod 4, which
// makes every run start on a sector boundary; the up to 3 frames before j0 are still in the ring from the previous
// tile.  s depends only on b mod 4 (N, M and M/2 are multiples of 4), i.e. it is one constant per thread for the
// rows it stores directly and one for their mirrors.  A run of tiles ends with a flush step for the last s frames.
// ------------------------------------------------------------------------------------------
template <int N>
struct BinMajorGeom {
    static constexpr int F = 16;          // frames per tile == warps per CTA
    static constexpr int SLOTS = F + 3;   // ring of frame regions: the tile plus the three frames before it
    static constexpr int M = N / 2;
    static constexpr int NEED = WarpGeom<N>::TILE > M + 1 ? WarpGeom<N>::TILE : M + 1;
    // region pitch (float2): >= NEED, == 1 mod 16 -> the 16 frames of one bin (consecutive slots) land in 16 distinct
    // bank pairs (except for the slots on either side of the ring's wrap-around)
    static constexpr int PITCH = ((NEED + 14) / 16) * 16 + 1;
    static_assert(PITCH >= NEED && PITCH % 16 == 1, "bad region pitch");
    static constexpr size_t SMEM = (size_t(N) + size_t(SLOTS) * PITCH) * sizeof(float2);
};

template <int N, bool STREAMING>
__global__ void __launch_bounds__(512, 1)
stft_warp_binmajor_kernel(const float* __restrict__ x, int64_t ns, int64_t clip_stride, int nt, int hop,
                          const float2* __restrict__ win_half, const float2* __restrict__ tw4,
                          const float2* __restrict__ tw_full, float2* __restrict__ out, int phase0, int runs_per_clip,
                          int tiles_per_run, int64_t total_runs) {
    using G = WarpGeom<N>;
    using B = BinMajorGeom<N>;
    constexpr int M = G::M, REGS = G::REGS, LOGR = G::LOGR;
    constexpr int F = B::F, SLOTS = B::SLOTS, PITCH = B::PITCH;
    extern __shared__ __align__(128) float2 smem[];
    float2* s_win = smem;       // M: (0.5 w[2n], 0.5 w[2n+1])
    float2* s_tw = smem + M;    // M: W_M^{k1*n2} at [k1*32+n2]
    float2* s_reg = smem + 2 * M;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    for (int i = tid; i < M; i += F * 32) {
        s_win[i] = win_half[i];
        s_tw[i] = tw4[i];
    }
    const float2 c_lane = tw_full[lane];  // W_N^lane
    LaneTw<N> lt;
    lt.init(lane);
    __syncthreads();

    auto store = [](float2* p, float2 v) {
        if constexpr (STREAMING) __stcs(p, v);
        else *p = v;
    };
    const int sw = tid & (F - 1);  // store phase: frame within the row's window
    const int su = tid / F;        // store phase: bin index mod 32
    // sector phase of element (b, j0): (phase0 + b * nt + j0) mod 4 with j0 = 0 mod 4; b = su mod 4 for the direct rows
    // (u, M + k), -su mod 4 for the mirrors (N - u, M - k), 0 for the rows M/2 and 3M/2 (handled by su == 0: same as sa)
    const int sa = (phase0 + (su & 3) * (nt & 3)) & 3;
    const int sb = (phase0 + ((4 - su) & 3) * (nt & 3)) & 3;
    const int tiles_per_clip = (nt + F - 1) / F;

    for (int64_t run = blockIdx.x; run < total_runs; run += gridDim.x) {
        const int64_t clip = run / runs_per_clip;
        const int t0 = int(run - clip * runs_per_clip) * tiles_per_run;
        const int t1 = min(t0 + tiles_per_run, tiles_per_clip);
        const int jlo = t0 * F, jhi = min(nt, t1 * F);  // the frames this run owns
        const float* xc = x + clip * clip_stride;
        float2* oc = out + clip * int64_t(N) * nt;

        for (int t = t0; t <= t1; ++t) {  // t == t1: flush of the frames still waiting for their row's window
            const int j0 = t * F;
            const int j = j0 + warp;
            if (t < t1 && j < nt) {  // warp-uniform
                // the frame's N samples as N/2 complex points, straight into registers (zeros outside [0, ns): zaf.py:99-125)
                const int64_t start = int64_t(j) * hop - M;
                float2 v[REGS];
                if (start >= 0 && start + N <= ns) {
                    const float2* p = reinterpret_cast<const float2*>(xc + start) + lane;
#pragma unroll
                    for (int r = 0; r < REGS; ++r) v[r] = __ldg(p + 32 * r);
                } else {
#pragma unroll
                    for (int r = 0; r < REGS; ++r) {
                        const int64_t s = start + 2 * (lane + 32 * r);
                        v[r].x = (s >= 0 && s < ns) ? __ldg(xc + s) : 0.f;
                        v[r].y = (s + 1 >= 0 && s + 1 < ns) ? __ldg(xc + s + 1) : 0.f;
                    }
                }
                // the samples the NEXT tile adds (F hops further on) towards L2 while this tile is transformed and stored
                if (t + 1 < t1) {
                    const int64_t nx = start + int64_t(F) * hop + N - hop;  // first sample frame j + F has and frame j + F - 1 has not
                    for (int64_t s = nx + 32 * lane; s < nx + hop && s < ns; s += 32 * 32)
                        if (s >= 0) prefetch_l2(xc + s);
                }
                float2* s_buf = s_reg + (j % SLOTS) * PITCH;
#pragma unroll
                for (int r = 0; r < REGS; ++r) {
                    const float2 w = s_win[lane + 32 * r];
                    v[r].x *= w.x;
                    v[r].y *= w.y;
                }
                warp_fft_half<N>(v, s_tw, s_buf, lane, lt);  // Z[lane + 32 k2] = v[bitrev(k2)]

                // the real-input unpack of stft_warp_kernel (see there), results parked in the frame's region
                const int src = (32 - lane) & 31;
                const float2 zmid = v[bitrev(REGS / 2, LOGR)];
                static_for<0, REGS / 2>([&](auto ic) {
                    constexpr int k2 = REGS / 2 - 1 - decltype(ic)::value;
                    const float2 z = v[bitrev(k2, LOGR)];
                    const float2 mine = v[bitrev(REGS - 1 - k2, LOGR)];
                    float2 p;
                    p.x = __shfl_sync(0xffffffffu, mine.x, src);
                    p.y = __shfl_sync(0xffffffffu, mine.y, src);
                    if (lane == 0) p = v[bitrev((REGS - k2) & (REGS - 1), LOGR)];
                    const float2 e = make_float2(z.x + p.x, z.y - p.y);
                    const float2 od = make_float2(z.y + p.y, p.x - z.x);
                    const float2 tt = cmul(mul_tw<k2, N / 32>(c_lane), od);
                    v[bitrev(k2, LOGR)] = cadd(e, tt);             // X[k]
                    v[bitrev(REGS - 1 - k2, LOGR)] = csub(e, tt);  // X[k + M]
                });
                __syncwarp();  // every lane is done with the transpose tile
                static_for<0, REGS / 2>([&](auto kc) {
                    constexpr int k2 = decltype(kc)::value;
                    s_buf[lane + 32 * k2] = v[bitrev(k2, LOGR)];
                    s_buf[M / 2 + lane + 32 * k2] = v[bitrev(REGS - 1 - k2, LOGR)];
                });
                if (lane == 0) s_buf[M] = make_float2(2.f * zmid.x, -2.f * zmid.y);  // X[M/2]: Z[M/2] pairs with itself, w = -i
            }
            __syncthreads();
            // Rows leave in batches of 8 (4 values of u x the two families): all shared-memory reads of a batch are issued
            // before its stores and every store has its own address registers, so nothing serialises on a register.
            constexpr int kBatch = M / 64 < 4 ? M / 64 : 4;
            static_assert((M / 64) % kBatch == 0, "batches of four bins");
            const int64_t step = int64_t(32) * nt;   // 32 rows
            const int64_t half = int64_t(M) * nt;    // M rows
            const int ja = j0 - sa + sw;  // the frame this thread stores for its direct rows
            if (ja >= jlo && ja < jhi) {
                const float2* r = s_reg + (ja % SLOTS) * PITCH + su;
                float2* o = oc + ja + int64_t(su) * nt;
#pragma unroll 1
                for (int i = 0; i < M / 64; i += kBatch, o += kBatch * step) {  // rows u = su + 32 i and M + u
                    float2 va[kBatch], vb[kBatch];
#pragma unroll
                    for (int b = 0; b < kBatch; ++b) {
                        va[b] = r[32 * (i + b)];
                        vb[b] = r[M / 2 + 32 * (i + b)];
                    }
#pragma unroll
                    for (int b = 0; b < kBatch; ++b) {
                        store(o + b * step, va[b]);
                        store(o + b * step + half, vb[b]);
                    }
                }
                if (su == 0) {  // rows M/2 and 3M/2
                    const float2 val = r[M];
                    store(oc + ja + int64_t(M / 2) * nt, val);
                    store(oc + ja + int64_t(M + M / 2) * nt, cconj(val));
                }
            }
            const int jb = j0 - sb + sw;  // ... and for their conjugate mirrors, rows N - u and M - u
            if (jb >= jlo && jb < jhi) {
                const float2* r = s_reg + (jb % SLOTS) * PITCH + su;
                float2* o = oc + jb + int64_t(N - su) * nt;
#pragma unroll 1
                for (int i = 0; i < M / 64; i += kBatch, o -= kBatch * step) {
                    float2 va[kBatch], vb[kBatch];
#pragma unroll
                    for (int b = 0; b < kBatch; ++b) {
                        va[b] = cconj(r[32 * (i + b)]);
                        vb[b] = cconj(r[M / 2 + 32 * (i + b)]);
                    }
#pragma unroll
                    for (int b = 0; b < kBatch; ++b) {
                        if (b > 0 || i > 0 || su > 0) {  // rows 0 and M have no mirror
                            store(o - b * step, va[b]);
                            store(o - b * step - half, vb[b]);
                        }
                    }
                }
            }
            __syncthreads();  // the ring slots of the next tile become transpose tiles again
        }
    }
}

// ------------------------------------------------------------------------------------------
// generic power-of-two N: one CTA per frame (grid-stride), Stockham FFT in shared memory
// ------------------------------------------------------------------------------------------
__global__ void stft_generic_kernel(const float* __restrict__ x, int64_t ns, int64_t clip_stride, int64_t nt,
                                    int64_t hop, int log2n, const float* __restrict__ window,
                                    const float2* __restrict__ tw_half, const float2* __restrict__ tw_full,
                                    float2* __restrict__ out, int layout, int64_t total_frames) {
    extern __shared__ float2 smem[];
    const int n = 1 << log2n;
    const int m = n >> 1;
    float2* a = smem;
    float2* b = smem + m;
    const int tid = threadIdx.x, nth = blockDim.x;
    for (int64_t f = blockIdx.x; f < total_frames; f += gridDim.x) {
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
        for (int k = tid; k < m; k += nth) {
            const float2 zk = z[k];
            const float2 zp = z[(m - k) & (m - 1)];
            const float2 e = make_float2(0.5f * (zk.x + zp.x), 0.5f * (zk.y - zp.y));
            const float2 od = make_float2(0.5f * (zk.y + zp.y), 0.5f * (zp.x - zk.x));
            const float2 t = cmul(tw_full[k], od);
            const float2 lo = cadd(e, t), hi = csub(e, t);
            if (layout == ZAFB_LAYOUT_FRAME_MAJOR) {
                out[f * n + k] = lo;
                out[f * n + m + k] = hi;
            } else {
                out[(clip * n + k) * nt + j] = lo;
                out[(clip * n + m + k) * nt + j] = hi;
            }
        }
        __syncthreads();
    }
}

// any N: direct DFT, one CTA per frame
__global__ void stft_dft_kernel(const float* __restrict__ x, int64_t ns, int64_t clip_stride, int64_t nt,
                                int64_t hop, int n, const float* __restrict__ window,
                                const float2* __restrict__ tw_full, float2* __restrict__ out, int layout,
                                int64_t total_frames) {
    extern __shared__ float2 smem[];
    float* fr = reinterpret_cast<float*>(smem);
    const int tid = threadIdx.x, nth = blockDim.x;
    const int pad = n / 2;
    for (int64_t f = blockIdx.x; f < total_frames; f += gridDim.x) {
        const int64_t clip = f / nt, j = f - clip * nt;
        const int64_t start = j * hop - pad;
        const float* xc = x + clip * clip_stride;
        for (int i = tid; i < n; i += nth) {
            const int64_t s = start + i;
            fr[i] = (s >= 0 && s < ns) ? xc[s] * window[i] : 0.f;
        }
        __syncthreads();
        for (int k = tid; k < n; k += nth) {
            float re = 0.f, im = 0.f;
            int idx = 0;
            for (int i = 0; i < n; ++i) {
                const float2 w = tw_full[idx];
                re = fmaf(fr[i], w.x, re);
                im = fmaf(fr[i], w.y, im);
                idx += k;
                if (idx >= n) idx -= n;
            }
            const float2 v = make_float2(re, im);
            if (layout == ZAFB_LAYOUT_FRAME_MAJOR) out[f * n + k] = v;
            else out[(clip * n + k) * nt + j] = v;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// ISTFT: gather-form overlap-add, one CTA per (clip, tile of output samples)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t floor_div(int64_t a, int64_t b) {
    int64_t q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
    return q;
}

// log2n >= 0: power-of-two path (full-size complex FFT of conj(X)); log2n < 0: direct inverse DFT.
__global__ void istft_tile_kernel(const float2* __restrict__ spec, int64_t nt, int64_t hop, int n, int log2n,
                                  int layout, const float2* __restrict__ tw_full, float scale,
                                  int64_t tile, int64_t tiles_per_clip, int64_t ola_len, int64_t out_start,
                                  int64_t out_len, float* __restrict__ y, int64_t y_stride) {
    extern __shared__ float2 smem[];
    float2* a = smem;
    float2* b = smem + n;
    float* acc = reinterpret_cast<float*>(smem + 2 * n);
    const int tid = threadIdx.x, nth = blockDim.x;
    const int64_t clip = blockIdx.x / tiles_per_clip;
    const int64_t t = blockIdx.x - clip * tiles_per_clip;
    const int64_t p0 = out_start + t * tile;  // OLA coordinates of the tile
    int64_t p1 = p0 + tile;
    if (p1 > out_start + out_len) p1 = out_start + out_len;
    for (int64_t i = tid; i < tile; i += nth) acc[i] = 0.f;
    int64_t j_lo = floor_div(p0 - n, hop) + 1;
    if (j_lo < 0) j_lo = 0;
    int64_t j_hi = floor_div(p1 - 1, hop);
    if (j_hi > nt - 1) j_hi = nt - 1;
    __syncthreads();
    for (int64_t j = j_lo; j <= j_hi; ++j) {
        for (int k = tid; k < n; k += nth) {
            const float2 v = (layout == ZAFB_LAYOUT_FRAME_MAJOR) ? spec[(clip * nt + j) * n + k]
                                                                 : spec[(clip * n + k) * nt + j];
            a[k] = cconj(v);
        }
        __syncthreads();
        const float2* res;
        if (log2n >= 0) {
            res = block_fft(a, b, tw_full, log2n, tid, nth);
        } else {
            for (int i = tid; i < n; i += nth) {
                float re = 0.f;
                int idx = 0;
                for (int k = 0; k < n; ++k) {
                    const float2 w = tw_full[idx];
                    re += a[k].x * w.x - a[k].y * w.y;
                    idx += i;
                    if (idx >= n) idx -= n;
                }
                b[i] = make_float2(re, 0.f);
            }
            __syncthreads();
            res = b;
        }
        const int64_t base = j * hop;
        for (int i = tid; i < n; i += nth) {
            const int64_t p = base + i;
            if (p >= p0 && p < p1) acc[p - p0] += res[i].x;
        }
        __syncthreads();
    }
    float* yc = y + clip * y_stride;
    for (int64_t p = p0 + tid; p < p1; p += nth) yc[p - out_start] = acc[p - p0] * scale;
    (void)ola_len;
}

bool g_attr_done = false;
int set_kernel_attrs() {
    if (g_attr_done) return ZAFB_OK;
    ZAFB_CUDA((cudaFuncSetAttribute(stft_warp_kernel<2048, false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(stft_warp_kernel<2048, true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(stft_warp_kernel<2048, false, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(stft_warp_kernel<1024, false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(stft_warp_kernel<512, false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(stft_warp_kernel<4096, false, ZAFB_STFT4096_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(stft_warp_kernel<256, false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(stft_warp_kernel<2048, false, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(stft_warp_kernel<2048, false, 6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(stft_warp_kernel<1024, false, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(stft_warp_kernel<512, false, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(stft_warp_kernel<4096, false, ZAFB_STFT4096_WARPS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA((cudaFuncSetAttribute(stft_warp_kernel<256, false, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)));
    ZAFB_CUDA(cudaFuncSetAttribute(stft_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    ZAFB_CUDA(cudaFuncSetAttribute(stft_dft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    ZAFB_CUDA(cudaFuncSetAttribute(istft_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    g_attr_done = true;
    return ZAFB_OK;
}

template <int N, bool STREAMING>
int launch_stft_binmajor_t(const zafb_stft_plan* p, const float* x, int64_t n_clips, int64_t ns, int64_t clip_stride,
                           int64_t nt, float2* out, cudaStream_t st) {
    using B = BinMajorGeom<N>;
    static_assert(B::SMEM <= size_t(kMaxDynSmem), "ring of frame regions does not fit");
    static bool attr = false;
    if (!attr) {
        ZAFB_CUDA((cudaFuncSetAttribute(stft_warp_binmajor_kernel<N, STREAMING>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kMaxDynSmem)));
        attr = true;
    }
    // runs of consecutive tiles of one clip: whole clips when there are enough of them, else about four runs per SM
    const int64_t sms = sm_count();
    const int64_t tiles_per_clip = ceil_div(nt, B::F);
    int64_t runs_per_clip = n_clips >= 4 * sms ? 1 : std::min<int64_t>(tiles_per_clip, ceil_div(4 * sms, n_clips));
    if (const int forced = env_flag("ZAFB_STFT_BM_RUNS_PER_CLIP", 0); forced > 0)  // tests
        runs_per_clip = std::min<int64_t>(tiles_per_clip, forced);
    const int64_t tiles_per_run = ceil_div(tiles_per_clip, runs_per_clip);
    runs_per_clip = ceil_div(tiles_per_clip, tiles_per_run);
    const int64_t runs = n_clips * runs_per_clip;
    const int64_t ctas = std::min<int64_t>(sms, runs);
    const int phase0 = int((reinterpret_cast<uintptr_t>(out) >> 3) & 3);
    stft_warp_binmajor_kernel<N, STREAMING><<<static_cast<unsigned>(ctas), B::F * 32, B::SMEM, st>>>(
        x, ns, clip_stride, int(nt), static_cast<int>(p->hop), p->d_window_half, p->d_tw_4step, p->d_tw_full, out, phase0,
        int(runs_per_clip), int(tiles_per_run), runs);
    ZAFB_LAUNCH_CHECK();
    return ZAFB_OK;
}

int launch_stft_binmajor(const zafb_stft_plan* p, const float* x, int64_t n_clips, int64_t ns, int64_t clip_stride,
                         int64_t nt, float2* out, cudaStream_t st) {
    const bool cs = env_flag("ZAFB_STFT_BM_CS", 0) != 0;  // 1: streaming (evict-first) stores
    if (p->n == 2048)
        return cs ? launch_stft_binmajor_t<2048, true>(p, x, n_clips, ns, clip_stride, nt, out, st)
                  : launch_stft_binmajor_t<2048, false>(p, x, n_clips, ns, clip_stride, nt, out, st);
    if (p->n == 1024)
        return cs ? launch_stft_binmajor_t<1024, true>(p, x, n_clips, ns, clip_stride, nt, out, st)
                  : launch_stft_binmajor_t<1024, false>(p, x, n_clips, ns, clip_stride, nt, out, st);
    if (p->n == 512)
        return cs ? launch_stft_binmajor_t<512, true>(p, x, n_clips, ns, clip_stride, nt, out, st)
                  : launch_stft_binmajor_t<512, false>(p, x, n_clips, ns, clip_stride, nt, out, st);
    return cs ? launch_stft_binmajor_t<256, true>(p, x, n_clips, ns, clip_stride, nt, out, st)
              : launch_stft_binmajor_t<256, false>(p, x, n_clips, ns, clip_stride, nt, out, st);
}

int fft_threads(int points) {  // threads for a block FFT of `points` complex points
    int t = points / 4;
    if (t < 32) t = 32;
    if (t > 256) t = 256;
    return t;
}

}  // namespace

extern "C" {

int zafb_stft_plan_create(zafb_stft_plan** out, const double* window, int64_t n, int64_t hop) {
    ZAFB_REQUIRE(out != nullptr && window != nullptr, "plan/window is NULL");
    ZAFB_REQUIRE(n >= 1 && hop >= 1, "window_length and step_length must be >= 1");
    if (n > (1 << 20)) return fail(ZAFB_E_UNSUPPORTED, "window_length %lld too large", (long long)n);
    zafb_stft_plan* p = new zafb_stft_plan();
    p->n = n;
    p->hop = hop;
    p->log2n = (is_pow2(n) && n >= 2) ? ilog2(n) : -1;
    int rc = upload_f32(&p->d_window, window, n);
    if (rc == ZAFB_OK) rc = upload_twiddles(&p->d_tw_full, n, n);
    if (rc == ZAFB_OK && p->log2n >= 1) {
        std::vector<double> half(n);
        for (int64_t i = 0; i < n; ++i) half[i] = 0.5 * window[i];
        float* tmp = nullptr;
        rc = upload_f32(&tmp, half.data(), n);
        p->d_window_half = reinterpret_cast<float2*>(tmp);
        if (rc == ZAFB_OK) rc = upload_twiddles(&p->d_tw_half, n / 2, n / 2);
    }
    if (rc == ZAFB_OK && (n == 4096 || n == 2048 || n == 1024 || n == 512 || n == 256)) {
        // four-step twiddles of the warp kernels: W_M^{k1*n2} laid out [k1][n2], M = n/2 = (M/32) x 32
        const int64_t m = n / 2;
        std::vector<double> t(2 * m);
        const double pi = 3.14159265358979323846264338327950288;
        for (int64_t k1 = 0; k1 < m / 32; ++k1)
            for (int64_t n2 = 0; n2 < 32; ++n2) {
                const double a = -2.0 * pi * double((k1 * n2) % m) / double(m);
                t[2 * (k1 * 32 + n2)] = std::cos(a);
                t[2 * (k1 * 32 + n2) + 1] = std::sin(a);
            }
        rc = upload_c32(&p->d_tw_4step, t.data(), m);
    }
    // COLA gain exactly as the reference: Python's builtin sum over float64 (zaf.py:241)
    double g = 0.0;
    for (int64_t i = 0; i < n; i += hop) g += window[i];
    p->gain = g;
    if (rc != ZAFB_OK) {
        zafb_stft_plan_destroy(p);
        return rc;
    }
    *out = p;
    return ZAFB_OK;
}

int zafb_stft_plan_destroy(zafb_stft_plan* p) {
    if (!p) return ZAFB_OK;
    cudaFree(p->d_window);
    cudaFree(p->d_window_half);
    cudaFree(p->d_tw_half);
    cudaFree(p->d_tw_full);
    cudaFree(p->d_tw_4step);
    delete p;
    return ZAFB_OK;
}

// test hook: 0 = auto, 1 = generic kernels only, 2 = require the warp kernel
int zafb_stft_plan_force_kernel(zafb_stft_plan* p, int which) {
    ZAFB_REQUIRE(p != nullptr && which >= 0 && which <= 2, "bad plan / kernel id");
    p->force_kernel = which;
    return ZAFB_OK;
}

// out_pitch / onesided: see stft_warp_kernel (warp kernels, FRAME_MAJOR only).  The public two-sided entry point passes (N, 0).
static int stft_impl(const zafb_stft_plan* p, const float* x, int64_t n_clips, int64_t ns, int64_t clip_stride,
                     float* out, int layout, void* stream, int64_t out_pitch, int onesided) {
    ZAFB_REQUIRE(p != nullptr, "plan is NULL");
    ZAFB_REQUIRE(n_clips >= 0 && ns >= 0, "n_clips and ns must be >= 0");
    ZAFB_REQUIRE(clip_stride >= ns, "clip_stride %lld < ns %lld", (long long)clip_stride, (long long)ns);
    ZAFB_REQUIRE(layout == ZAFB_LAYOUT_FRAME_MAJOR || layout == ZAFB_LAYOUT_BIN_MAJOR, "bad layout %d", layout);
    ZAFB_REQUIRE(out != nullptr && (x != nullptr || ns == 0 || n_clips == 0), "x/out is NULL");
    int rc = set_kernel_attrs();
    if (rc != ZAFB_OK) return rc;
    int64_t nt = 0;
    zafb_stft_geometry(ns, p->n, p->hop, nullptr, &nt, nullptr);
    const int64_t total = n_clips * nt;
    if (total == 0) return ZAFB_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int sms = sm_count();
    float2* o = reinterpret_cast<float2*>(out);

    const bool aligned = (reinterpret_cast<uintptr_t>(x) % 8 == 0) && (n_clips <= 1 || clip_stride % 2 == 0) && (p->hop % 2 == 0);
    const bool warp_ok = (p->n == 4096 || p->n == 2048 || p->n == 1024 || p->n == 512 || p->n == 256) && aligned;
    if (p->force_kernel == 2 && !warp_ok)
        return fail(ZAFB_E_UNSUPPORTED, "warp kernel needs N = 256 ... 4096 (a power of two), even hop/stride, 8-byte aligned x");
    if (warp_ok && p->force_kernel != 1) {
        // measured on cfg 2 (B200), profiles/r01_stft_experiments.txt: direct streaming stores 3.04-3.11 ms, TMA bulk stores
        // 3.13-3.20 ms; 6 warps per CTA 3.04, 8 -> 3.11, 10 -> 3.36, 4 -> 3.40.  The defaults are the fastest combination.
        const int bulk = (p->n == 2048 && !onesided && out_pitch == p->n) ? env_flag("ZAFB_STFT_BULK", 0) : 0;
        const int n = int(p->n);
        auto run = [&](const float* xs, int64_t clips, float2* dst, int64_t pitch = -1) -> int {
            if (pitch < 0) pitch = n;
            const int64_t frames = clips * nt;
            const int warps = n == 4096 ? ZAFB_STFT4096_WARPS : (n == 2048 && !bulk && env_flag("ZAFB_STFT_WARPS", 6) == 6) ? 6 : 8;
            const size_t smem = (size_t(n) + size_t(warps) * (n / 64) * kFft1024Pitch) * sizeof(float2);
            int64_t ctas = ceil_div(frames, warps);
            const int64_t resident = int64_t(sms) * (n == 4096 ? 1 : n == 2048 ? 2 : n == 256 ? ZAFB_STFT256_CTAS : 3);
            if (ctas > resident) ctas = resident;
            auto kern = onesided ? (n == 4096 ? stft_warp_kernel<4096, false, ZAFB_STFT4096_WARPS, true>
                                    : n == 256 ? stft_warp_kernel<256, false, 8, true>
                                    : n == 512 ? stft_warp_kernel<512, false, 8, true>
                                    : n == 1024 ? stft_warp_kernel<1024, false, 8, true>
                                    : warps == 6 ? stft_warp_kernel<2048, false, 6, true>
                                                 : stft_warp_kernel<2048, false, 8, true>)
                        : n == 4096 ? stft_warp_kernel<4096, false, ZAFB_STFT4096_WARPS>
                        : n == 256 ? stft_warp_kernel<256, false, 8>
                        : n == 512 ? stft_warp_kernel<512, false, 8>
                        : n == 1024 ? stft_warp_kernel<1024, false, 8>
                        : warps == 6 ? stft_warp_kernel<2048, false, 6>
                        : (bulk && reinterpret_cast<uintptr_t>(dst) % 16 == 0 ? stft_warp_kernel<2048, true, 8>
                                                                             : stft_warp_kernel<2048, false, 8>);
            kern<<<static_cast<unsigned>(ctas), warps * 32, smem, st>>>(
                xs, ns, clip_stride, nt, static_cast<int>(p->hop), p->d_window_half, p->d_tw_4step, p->d_tw_full, dst, frames,
                env_flag("ZAFB_STFT_SEQ", 1), pitch);
            ZAFB_LAUNCH_CHECK();
            return ZAFB_OK;
        };
        if (onesided) return run(x, n_clips, o, out_pitch);
        if (layout == ZAFB_LAYOUT_FRAME_MAJOR) return run(x, n_clips, o);
        // the reference's C-order memory, written directly by stft_warp_binmajor_kernel (ZAFB_STFT_BM_DIRECT=0: the
        // older route, frame-major into scratch + tiled transpose)
        if (n <= 2048 && env_flag("ZAFB_STFT_BM_DIRECT", 1) && nt < (int64_t(1) << 27) && reinterpret_cast<uintptr_t>(out) % 8 == 0)
            return launch_stft_binmajor(p, x, n_clips, ns, clip_stride, nt, o, st);
        return bin_major_from_frame_major(o, n_clips, nt, p->n, st, [&](int64_t c0, int64_t nc, float2* scratch) {
            return run(x + c0 * clip_stride, nc, scratch);
        });
    }
    if (onesided || out_pitch != p->n)
        return fail(ZAFB_E_UNSUPPORTED, "one-sided stft: internal error (the caller compacts the spectrum for this geometry)");
    // generic kernels: one CTA per frame.  They can store either layout, but a BIN_MAJOR store is one 8-byte element per
    // row (measured 0.035 of the HBM peak at N = 256), so that layout goes through frame-major scratch and the tiled
    // transpose as well (ZAFB_GENERIC_BM_DIRECT=1 keeps the strided stores).
    auto run_generic = [&](const float* xs, int64_t clips, float2* dst, int lay) -> int {
        const int64_t frames = clips * nt;
        const int64_t grid = frames < int64_t(sms) * 32 ? frames : int64_t(sms) * 32;
        if (p->log2n >= 1) {
            const int m = int(p->n / 2);
            const size_t smem = size_t(p->n) * sizeof(float2);
            if (smem > size_t(kMaxDynSmem))
                return fail(ZAFB_E_UNSUPPORTED, "window_length %lld needs %zu B of shared memory", (long long)p->n, smem);
            stft_generic_kernel<<<static_cast<unsigned>(grid), fft_threads(m), smem, st>>>(
                xs, ns, clip_stride, nt, p->hop, p->log2n, p->d_window, p->d_tw_half, p->d_tw_full, dst, lay, frames);
        } else {
            const size_t smem = size_t(p->n) * sizeof(float) + 16;
            if (smem > size_t(kMaxDynSmem))
                return fail(ZAFB_E_UNSUPPORTED, "window_length %lld needs %zu B of shared memory", (long long)p->n, smem);
            int th = int(p->n) < 256 ? ((int(p->n) + 31) / 32) * 32 : 256;
            stft_dft_kernel<<<static_cast<unsigned>(grid), th, smem, st>>>(xs, ns, clip_stride, nt, p->hop, int(p->n),
                                                                            p->d_window, p->d_tw_full, dst, lay, frames);
        }
        ZAFB_LAUNCH_CHECK();
        return ZAFB_OK;
    };
    if (layout == ZAFB_LAYOUT_FRAME_MAJOR || nt == 1 || p->n == 1 || env_flag("ZAFB_GENERIC_BM_DIRECT", 0))
        return run_generic(x, n_clips, o, layout);
    return bin_major_from_frame_major(o, n_clips, nt, p->n, st, [&](int64_t c0, int64_t nc, float2* scratch) {
        return run_generic(x + c0 * clip_stride, nc, scratch, ZAFB_LAYOUT_FRAME_MAJOR);
    });
}

int zafb_stft_f32(const zafb_stft_plan* p, const float* x, int64_t n_clips, int64_t ns, int64_t clip_stride,
                  float* out, int layout, void* stream) {
    return stft_impl(p, x, n_clips, ns, clip_stride, out, layout, stream, p ? p->n : 0, 0);
}

// spec_pitch / onesided: see istft_warp_kernel.  The public two-sided entry point passes (N, 0).
static int istft_impl(const zafb_stft_plan* p, const float* spec, int64_t n_clips, int64_t nt, int layout, float* y,
                      int64_t y_stride, void* stream, int64_t spec_pitch, int onesided) {
    ZAFB_REQUIRE(p != nullptr, "plan is NULL");
    ZAFB_REQUIRE(n_clips >= 0 && nt >= 0, "n_clips and nt must be >= 0");
    ZAFB_REQUIRE(layout == ZAFB_LAYOUT_FRAME_MAJOR || layout == ZAFB_LAYOUT_BIN_MAJOR, "bad layout %d", layout);
    int rc = set_kernel_attrs();
    if (rc != ZAFB_OK) return rc;
    int64_t ola = 0, start = 0, len = 0;
    zafb_istft_geometry(p->n, nt, p->hop, &ola, &start, &len);
    ZAFB_REQUIRE(y_stride >= len, "y_stride %lld < output length %lld", (long long)y_stride, (long long)len);
    if (n_clips == 0 || len == 0) return ZAFB_OK;
    ZAFB_REQUIRE(spec != nullptr && y != nullptr, "spec/y is NULL");
    const int n = int(p->n);
    {
        const bool aligned = reinterpret_cast<uintptr_t>(spec) % 8 == 0 && reinterpret_cast<uintptr_t>(y) % 8 == 0 &&
                             (n_clips <= 1 || y_stride % 2 == 0);
        const int64_t ratio = (p->hop > 0 && n % p->hop == 0) ? n / p->hop : 0;
        const bool warp_ok = (n == 4096 || n == 2048 || n == 1024 || n == 512 || n == 256) && aligned &&
                             (ratio == 2 || ratio == 4 || (ratio == 8 && n != 256));  // N = 256: 4 points per lane, at most 4 parts
        if (p->force_kernel == 2 && !warp_ok)
            return fail(ZAFB_E_UNSUPPORTED, "istft warp kernel needs N = 256 ... 4096, hop = N/2, N/4 or N/8 (N/8: N >= 512), even y_stride");
        if (warp_ok && p->force_kernel != 1) {
            cudaStream_t st = static_cast<cudaStream_t>(stream);
            auto run = [&](const float2* s2, int64_t clips, float* yy, int64_t pitch) -> int {
                return istft_warp_dispatch(p, s2, clips, nt, yy, y_stride, st, pitch, onesided, nullptr, 0);  // istft.cu
            };
            const float2* s2 = reinterpret_cast<const float2*>(spec);
            if (layout == ZAFB_LAYOUT_FRAME_MAJOR) return run(s2, n_clips, y, spec_pitch);
            // the reference's C order read directly, for every batch size (a clip's result must not depend on the batch
            // around it); ZAFB_ISTFT_BM_DIRECT=0: the older route through frame-major scratch
            if ((n == 2048 || n == 1024) && (ratio == 2 || ratio == 4) && nt < (int64_t(1) << 26) && env_flag("ZAFB_ISTFT_BM_DIRECT", 1)) {
                return istft_binmajor_dispatch(p, s2, n_clips, nt, y, y_stride, st);
            }
            return frame_major_from_bin_major(s2, n_clips, nt, int64_t(n), st, [&](int64_t c0, int64_t nc, const float2* scratch) {
                return run(scratch, nc, y + c0 * y_stride, int64_t(n));
            });
        }
    }
    if (onesided) return fail(ZAFB_E_UNSUPPORTED, "one-sided istft: internal error (the caller expands the spectrum for this geometry)");
    // tile: about 4 windows of output, bounded by shared memory (2 n float2 + tile floats)
    const size_t fft_bytes = size_t(2) * n * sizeof(float2);
    if (fft_bytes + 4096 > size_t(kMaxDynSmem))
        return fail(ZAFB_E_UNSUPPORTED, "istft: window_length %d needs %zu B of shared memory", n, fft_bytes);
    int64_t tile = 4 * int64_t(n);
    const int64_t max_tile = int64_t((kMaxDynSmem - fft_bytes) / sizeof(float));
    if (tile > max_tile) tile = max_tile;
    if (tile > len) tile = len;
    const int64_t tiles = ceil_div(len, tile);
    const size_t smem = fft_bytes + size_t(tile) * sizeof(float);
    const float scale = static_cast<float>(1.0 / (double(n) * p->gain));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto run_generic = [&](const float2* s2, int64_t clips, float* yy, int lay) -> int {
        const int64_t blocks = clips * tiles;
        if (blocks > 0x7fffffffLL) return fail(ZAFB_E_UNSUPPORTED, "istft: too many tiles (%lld)", (long long)blocks);
        istft_tile_kernel<<<static_cast<unsigned>(blocks), fft_threads(n), smem, st>>>(
            s2, nt, p->hop, n, p->log2n >= 1 ? p->log2n : (n == 1 ? 0 : -1), lay, p->d_tw_full, scale, tile, tiles, ola, start,
            len, yy, y_stride);
        ZAFB_LAUNCH_CHECK();
        return ZAFB_OK;
    };
    const float2* s2 = reinterpret_cast<const float2*>(spec);
    // BIN_MAJOR input would be read one 8-byte element per row: transpose it into frame-major scratch first
    if (layout == ZAFB_LAYOUT_FRAME_MAJOR || nt == 1 || n == 1 || env_flag("ZAFB_GENERIC_BM_DIRECT", 0))
        return run_generic(s2, n_clips, y, layout);
    return frame_major_from_bin_major(s2, n_clips, nt, int64_t(n), st, [&](int64_t c0, int64_t nc, const float2* scratch) {
        return run_generic(scratch, nc, y + c0 * y_stride, ZAFB_LAYOUT_FRAME_MAJOR);
    });
}

int zafb_istft_f32(const zafb_stft_plan* p, const float* spec, int64_t n_clips, int64_t nt, int layout, float* y,
                   int64_t y_stride, void* stream) {
    return istft_impl(p, spec, n_clips, nt, layout, y, y_stride, stream, p ? p->n : 0, 0);
}

// ------------------------------------------------------------------ one-sided spectra (non-reference extension)
// Bins 0 .. floor(N/2) of every frame, FRAME_MAJOR, `pitch` complex elements between frames: the rest of zaf.stft's
// two-sided spectrum is the Hermitian mirror (zafb_spec_mirror_f32 rebuilds it).  Warp kernels store / load the half
// directly (half the HBM traffic of the spectrum); other geometries go through two-sided scratch.
static bool onesided_warp_ok(const zafb_stft_plan* p, const void* sig, int64_t n_clips, int64_t stride, const void* spec) {
    return (p->n == 4096 || p->n == 2048 || p->n == 1024 || p->n == 512 || p->n == 256) && p->hop % 2 == 0 &&
           reinterpret_cast<uintptr_t>(sig) % 8 == 0 && (n_clips <= 1 || stride % 2 == 0) && reinterpret_cast<uintptr_t>(spec) % 8 == 0 &&
           p->force_kernel != 1;
}

int zafb_stft_onesided_f32(const zafb_stft_plan* p, const float* x, int64_t n_clips, int64_t ns, int64_t clip_stride, float* out,
                           int64_t out_pitch, void* stream) {
    ZAFB_REQUIRE(p != nullptr, "plan is NULL");
    ZAFB_REQUIRE(n_clips >= 0 && ns >= 0 && clip_stride >= ns, "bad batch geometry");
    const int64_t n = p->n, bins = n / 2 + 1;
    ZAFB_REQUIRE(out_pitch >= bins, "out_pitch %lld < %lld one-sided bins", (long long)out_pitch, (long long)bins);
    int64_t nt = 0;
    zafb_stft_geometry(ns, n, p->hop, nullptr, &nt, nullptr);
    if (n_clips * nt == 0) return ZAFB_OK;
    ZAFB_REQUIRE(out != nullptr && (x != nullptr || ns == 0), "x/out is NULL");
    if (onesided_warp_ok(p, x, n_clips, clip_stride, out))
        return stft_impl(p, x, n_clips, ns, clip_stride, out, ZAFB_LAYOUT_FRAME_MAJOR, stream, out_pitch, 1);
    // two-sided into scratch (at most ~1 GB at a time), then a strided device copy of the lower half
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t clip_bytes = size_t(nt) * n * sizeof(float2);
    int64_t per = int64_t((size_t(1) << 30) / (clip_bytes ? clip_bytes : 1));
    if (per < 1) per = 1;
    if (per > n_clips) per = n_clips;
    float* scratch = nullptr;
    ZAFB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&scratch), size_t(per) * clip_bytes, st));
    int rc = ZAFB_OK;
    for (int64_t c0 = 0; c0 < n_clips && rc == ZAFB_OK; c0 += per) {
        const int64_t nc = std::min(per, n_clips - c0);
        rc = stft_impl(p, x + c0 * clip_stride, nc, ns, clip_stride, scratch, ZAFB_LAYOUT_FRAME_MAJOR, stream, n, 0);
        if (rc == ZAFB_OK) {
            const cudaError_t e = cudaMemcpy2DAsync(out + size_t(c0) * nt * out_pitch * 2, size_t(out_pitch) * sizeof(float2), scratch,
                                                    size_t(n) * sizeof(float2), size_t(bins) * sizeof(float2), size_t(nc * nt),
                                                    cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) rc = fail(ZAFB_E_CUDA, "one-sided stft: %s", cudaGetErrorString(e));
        }
    }
    cudaFreeAsync(scratch, st);
    return rc;
}

int zafb_istft_onesided_f32(const zafb_stft_plan* p, const float* spec, int64_t n_clips, int64_t nt, int64_t spec_pitch, float* y,
                            int64_t y_stride, void* stream) {
    ZAFB_REQUIRE(p != nullptr, "plan is NULL");
    ZAFB_REQUIRE(n_clips >= 0 && nt >= 0, "n_clips and nt must be >= 0");
    const int64_t n = p->n, bins = n / 2 + 1;
    ZAFB_REQUIRE(spec_pitch >= bins, "spec_pitch %lld < %lld one-sided bins", (long long)spec_pitch, (long long)bins);
    int64_t len = 0;
    zafb_istft_geometry(n, nt, p->hop, nullptr, nullptr, &len);
    ZAFB_REQUIRE(y_stride >= len, "y_stride %lld < output length %lld", (long long)y_stride, (long long)len);
    if (n_clips == 0 || len == 0) return ZAFB_OK;
    ZAFB_REQUIRE(spec != nullptr && y != nullptr, "spec/y is NULL");
    const int64_t ratio = (p->hop > 0 && n % p->hop == 0) ? n / p->hop : 0;
    if (onesided_warp_ok(p, y, n_clips, y_stride, spec) && (ratio == 2 || ratio == 4 || (ratio == 8 && n != 256)))
        return istft_impl(p, spec, n_clips, nt, ZAFB_LAYOUT_FRAME_MAJOR, y, y_stride, stream, spec_pitch, 1);
    // other geometries: rebuild the two-sided spectrum in scratch (at most ~1 GB at a time) and run the two-sided transform
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t clip_bytes = size_t(nt) * n * sizeof(float2);
    int64_t per = int64_t((size_t(1) << 30) / (clip_bytes ? clip_bytes : 1));
    if (per < 1) per = 1;
    if (per > n_clips) per = n_clips;
    float* scratch = nullptr;
    ZAFB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&scratch), size_t(per) * clip_bytes, st));
    int rc = ZAFB_OK;
    for (int64_t c0 = 0; c0 < n_clips && rc == ZAFB_OK; c0 += per) {
        const int64_t nc = std::min(per, n_clips - c0);
        rc = zafb_spec_mirror_f32(spec + size_t(c0) * nt * spec_pitch * 2, spec_pitch, nc * nt, n, scratch, stream);
        if (rc == ZAFB_OK)
            rc = istft_impl(p, scratch, nc, nt, ZAFB_LAYOUT_FRAME_MAJOR, y + c0 * y_stride, y_stride, stream, n, 0);
    }
    cudaFreeAsync(scratch, st);
    return rc;
}

// ISTFT of mask * X with the mask multiply fused into the loads (see istft_warp_kernel, MASKED).  FRAME_MAJOR spectra of
// the BASELINE geometry (window_length 2048, hop 512), two-sided (onesided = 0, frame pitch N) or one-sided (bins 0 .. N/2,
// spec_pitch complex elements per frame); every other geometry returns ZAFB_E_UNSUPPORTED and the caller multiplies
// first (zafb_spec_mask_f32).
int zafb_istft_masked_f32(const zafb_stft_plan* p, const float* spec, int64_t n_clips, int64_t nt, int64_t spec_pitch, int onesided,
                          const float* mask, int64_t mask_pitch, float* y, int64_t y_stride, void* stream) {
    ZAFB_REQUIRE(p != nullptr, "plan is NULL");
    ZAFB_REQUIRE(n_clips >= 0 && nt >= 0, "n_clips and nt must be >= 0");
    const int64_t n = p->n, bins = n / 2 + 1;
    ZAFB_REQUIRE(mask_pitch >= bins, "mask_pitch %lld < %lld mask bins", (long long)mask_pitch, (long long)bins);
    ZAFB_REQUIRE(onesided ? spec_pitch >= bins : spec_pitch == n, "bad spectrum pitch %lld", (long long)spec_pitch);
    int64_t len = 0;
    zafb_istft_geometry(n, nt, p->hop, nullptr, nullptr, &len);
    ZAFB_REQUIRE(y_stride >= len, "y_stride %lld < output length %lld", (long long)y_stride, (long long)len);
    if (n_clips == 0 || len == 0) return ZAFB_OK;
    ZAFB_REQUIRE(spec != nullptr && y != nullptr && mask != nullptr, "spec/mask/y is NULL");
    const bool aligned = reinterpret_cast<uintptr_t>(spec) % 8 == 0 && reinterpret_cast<uintptr_t>(y) % 8 == 0 &&
                         (n_clips <= 1 || y_stride % 2 == 0);
    if (n != 2048 || p->hop != 512 || !aligned)
        return fail(ZAFB_E_UNSUPPORTED, "masked istft: fused kernel exists for window_length 2048, hop 512, 8-byte aligned buffers only");
    int rc = set_kernel_attrs();
    if (rc != ZAFB_OK) return rc;
    return istft_warp_dispatch(p, reinterpret_cast<const float2*>(spec), n_clips, nt, y, y_stride, static_cast<cudaStream_t>(stream),
                               spec_pitch, onesided ? 1 : 0, mask, mask_pitch);
}

// ------------------------------------------------------------------ host-buffer pipelines
}  // extern "C"

namespace {

// The fills themselves live in host_mirror.cpp (plain C++ for the host compiler: AVX2 / SSE2 / portable loops).
void mirror_fill(float2* out, int64_t frames, int64_t n) { host_mirror_fill_frames(out, frames, n); }
void mirror_fill_rows(float2* out, int64_t nt, int64_t n, int64_t r_lo, int64_t r_hi) { host_mirror_fill_rows(out, nt, n, r_lo, r_hi); }

// Fill threads: ZAFB_HOST_MIRROR_THREADS, else min(16, host cores / processes sharing the host), where the process count
// is LOCAL_WORLD_SIZE (set by torchrun: one rank per GPU) or 1.  Measured on a 16-core B200 host (cfg 2, 15.75 GB
// result): 4 threads 276 ms (no better than the full copy, 282 ms), 8 -> 193 ms, 12-16 -> 172-177 ms; fewer than 6
// threads leave the path off.
int host_mirror_threads() {
    int t = env_flag("ZAFB_HOST_MIRROR_THREADS", -1);
    if (t >= 0) return t;
    int procs = env_flag("LOCAL_WORLD_SIZE", 1);
    if (procs < 1) procs = 1;
    int cores = int(std::thread::hardware_concurrency());
    cpu_set_t set;  // the cores this process may run on (a container's affinity mask can be smaller than the machine)
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) cores = CPU_COUNT(&set);
    // More than two ranks sharing one host: the aggregate D2H rate is bounded by the host's memory system, not by a
    // rank's own PCIe link (SCALE_r01: 88-99 GB/s over 4-8 ranks), and the fill's extra read + write of the mirrored half
    // makes that bound worse (N = 4 with the fill 4.84e6 frames/s < N = 8 without it 5.36e6): the full copy is used.
    if (procs > 2) return 0;
    t = cores / procs;
    if (t > 16) t = 16;
    return t >= 6 ? t : 0;
}

// STFT into HOST memory with half the PCIe traffic: the spectrum of a real signal is Hermitian, so only bins 0 .. N/2
// (of every frame, or -- C order -- the first N/2 + 1 rows of every clip) cross the link, one strided D2H copy per chunk,
// and host threads write the mirrored half, X[N - k] = conj(X[k]) -- an exact copy with a sign flip, bit-identical to
// what the kernel stores on the device.
// Chunk c's fill overlaps the copies of the chunks after it.
int stft_host_mirrored(const zafb_stft_plan* p, const float* x, int64_t n_clips, int64_t ns, int64_t clip_stride, float* out,
                       int64_t nt, int layout, int threads) {
    const bool frame_major = layout == ZAFB_LAYOUT_FRAME_MAJOR;
    HostPipe& hp = host_pipe();
    std::lock_guard<std::mutex> lock(hp.mu);
    const int64_t n = p->n;
    const size_t out_clip = size_t(nt) * n * sizeof(float2);
    const int64_t dpitch = (ns + 1) & ~int64_t(1);
    const size_t in_dev = size_t(dpitch) * sizeof(float);
    // 16 MB stages unless ZAFB_PIPE_CHUNK_MB says otherwise: the fill threads start sooner and read a chunk soon after
    // the DMA wrote it (cfg 2: 207 ms against 214 ms with 64 MB stages, profiles/r01q_e2e_mirror.log)
    const size_t stage = getenv("ZAFB_PIPE_CHUNK_MB") ? host_pipe_chunk_bytes() : (size_t(16) << 20);
    int64_t per = int64_t(stage / out_clip);
    if (per < 1) per = 1;
    if (per * HostPipe::kStages > n_clips) per = (n_clips + HostPipe::kStages - 1) / HostPipe::kStages;
    if (per < 1) per = 1;
    int rc = hp.ensure(size_t(per) * in_dev, size_t(per) * out_clip);
    if (rc != ZAFB_OK) return rc;
    const int64_t n_chunks = ceil_div(n_clips, per);
    int dev = 0;
    ZAFB_CUDA(cudaGetDevice(&dev));
    static std::vector<cudaEvent_t> events;  // guarded by hp.mu; they belong to events_dev
    static int events_dev = -1;
    if (events_dev != dev) {  // the process switched devices: events cannot be recorded on another device's stream
        for (cudaEvent_t e : events) cudaEventDestroy(e);
        events.clear();
        events_dev = dev;
    }
    while (int64_t(events.size()) < n_chunks) {
        cudaEvent_t e;
        ZAFB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        events.push_back(e);
    }
    std::atomic<int64_t> recorded{0};
    std::atomic<int> stop{0};
    const bool nofill = env_flag("ZAFB_HOST_MIRROR_NOFILL", 0) != 0;
    float2* o2 = reinterpret_cast<float2*>(out);
    std::vector<std::thread> pool;
    pool.reserve(threads);
    auto spawn = [&](int w) {
        pool.emplace_back([&, w]() {
            cudaSetDevice(dev);
            for (int64_t c = 0; c < n_chunks; ++c) {
                while (recorded.load(std::memory_order_acquire) <= c) {
                    if (stop.load(std::memory_order_relaxed)) return;
                    std::this_thread::yield();
                }
                if (cudaEventSynchronize(events[c]) != cudaSuccess) return;
                const int64_t c0 = c * per;
                const int64_t nc = (c0 + per <= n_clips) ? per : n_clips - c0;
                if (nofill) continue;  // measurement only (ZAFB_HOST_MIRROR_NOFILL=1): the copy pipeline without the fill
                if (frame_major) {
                    const int64_t frames = nc * nt;
                    const int64_t f_lo = frames * w / threads, f_hi = frames * (w + 1) / threads;
                    mirror_fill(o2 + (c0 * nt + f_lo) * n, f_hi - f_lo, n);
                } else {  // C order: whole rows, N - k <- conj(k)
                    const int64_t rows = nc * (n / 2 - 1);
                    mirror_fill_rows(o2 + c0 * nt * n, nt, n, rows * w / threads, rows * (w + 1) / threads);
                }
            }
        });
    };
    try {
        for (int w = 0; w < threads; ++w) spawn(w);
    } catch (...) {  // no exception crosses the C ABI: the threads that did start are told to stop, the call fails
        stop.store(1);  // nothing has been recorded yet: every started worker is in its wait loop and sees this
        for (auto& t : pool) t.join();
        return fail(ZAFB_E_NOMEM, "host pipeline (mirrored): could not start %d fill threads", threads);
    }
    int s = 0;
    for (int64_t c = 0; c < n_chunks && rc == ZAFB_OK; ++c, s = (s + 1) % HostPipe::kStages) {
        const int64_t c0 = c * per;
        const int64_t nc = (c0 + per <= n_clips) ? per : n_clips - c0;
        cudaError_t e = cudaSuccess;
        if (ns > 0)
            e = cudaMemcpy2DAsync(hp.d_in[s], in_dev, x + c0 * clip_stride, size_t(clip_stride) * sizeof(float),
                                  size_t(ns) * sizeof(float), size_t(nc), cudaMemcpyHostToDevice, hp.st[s]);
        if (e == cudaSuccess) {
            rc = zafb_stft_f32(p, static_cast<const float*>(hp.d_in[s]), nc, ns, dpitch, static_cast<float*>(hp.d_out[s]),
                               layout, hp.st[s]);
            if (rc != ZAFB_OK) break;
            if (frame_major)  // bins 0 .. N/2 of every frame
                e = cudaMemcpy2DAsync(o2 + c0 * nt * n, size_t(n) * sizeof(float2), hp.d_out[s], size_t(n) * sizeof(float2),
                                      size_t(n / 2 + 1) * sizeof(float2), size_t(nc * nt), cudaMemcpyDeviceToHost, hp.st[s]);
            else              // rows 0 .. N/2 of every clip
                e = cudaMemcpy2DAsync(o2 + c0 * nt * n, out_clip, hp.d_out[s], out_clip, size_t(n / 2 + 1) * nt * sizeof(float2),
                                      size_t(nc), cudaMemcpyDeviceToHost, hp.st[s]);
        }
        g_h2d_bytes.fetch_add(nc * ns * int64_t(sizeof(float)), std::memory_order_relaxed);
        g_d2h_bytes.fetch_add(nc * nt * (n / 2 + 1) * int64_t(sizeof(float2)), std::memory_order_relaxed);
        if (e == cudaSuccess) e = cudaEventRecord(events[c], hp.st[s]);
        if (e != cudaSuccess) {
            rc = fail(ZAFB_E_CUDA, "host pipeline (mirrored): %s", cudaGetErrorString(e));
            break;
        }
        recorded.store(c + 1, std::memory_order_release);
    }
    if (rc != ZAFB_OK) stop.store(1);
    for (auto& t : pool) t.join();
    for (int i = 0; i < HostPipe::kStages; ++i) {
        cudaError_t e = cudaStreamSynchronize(hp.st[i]);
        if (e != cudaSuccess && rc == ZAFB_OK) rc = fail(ZAFB_E_CUDA, "host pipeline: %s", cudaGetErrorString(e));
    }
    return rc;
}

}  // namespace

extern "C" {

int zafb_host_mirror_fill(float* spectrum, int64_t frames, int64_t n) {
    ZAFB_REQUIRE(frames >= 0 && n >= 4 && n % 4 == 0, "window_length must be a positive multiple of 4");
    if (frames == 0) return ZAFB_OK;
    ZAFB_REQUIRE(spectrum != nullptr, "spectrum is NULL");
    mirror_fill(reinterpret_cast<float2*>(spectrum), frames, n);
    return ZAFB_OK;
}

int zafb_stft_host_f32(const zafb_stft_plan* p, const float* x, int64_t n_clips, int64_t ns, int64_t clip_stride,
                       float* out, int layout) {
    ZAFB_REQUIRE(p != nullptr, "plan is NULL");
    ZAFB_REQUIRE(n_clips >= 0 && ns >= 0 && clip_stride >= ns, "bad batch geometry");
    int64_t nt = 0;
    zafb_stft_geometry(ns, p->n, p->hop, nullptr, &nt, nullptr);
    if (n_clips == 0) return ZAFB_OK;
    ZAFB_REQUIRE(out != nullptr && (x != nullptr || ns == 0), "x/out is NULL");
    const size_t out_clip = size_t(nt) * p->n * sizeof(float2);
    // device-side clip pitch: even number of samples so the float2 fast path stays aligned
    const int64_t dpitch = (ns + 1) & ~int64_t(1);
    // large results: half-spectrum copy + mirror fill on host threads (ZAFB_HOST_MIRROR=0 turns it off)
    // (cudaMemcpy2D pitches are limited to 2 GB: a C-order clip longer than that takes the full-copy pipeline)
    const bool pitch_ok = (layout == ZAFB_LAYOUT_FRAME_MAJOR || out_clip < (size_t(1) << 31)) &&
                          size_t(clip_stride) * sizeof(float) < (size_t(1) << 31);
    if (pitch_ok && p->n % 4 == 0 && p->n >= 64 &&
        size_t(n_clips) * out_clip >= (size_t(env_flag("ZAFB_HOST_MIRROR_MIN_MB", 256)) << 20) && env_flag("ZAFB_HOST_MIRROR", 1)) {
        const int threads = host_mirror_threads();
        if (threads >= 2) return stft_host_mirrored(p, x, n_clips, ns, clip_stride, out, nt, layout, threads);
    }
    return run_host_pipeline(x, size_t(clip_stride) * sizeof(float), size_t(ns) * sizeof(float), size_t(dpitch) * sizeof(float),
                             out, out_clip, out_clip, out_clip, n_clips,
                             [&](void* d_in, void* d_out, int64_t, int64_t nc, cudaStream_t st) {
                                 return zafb_stft_f32(p, static_cast<const float*>(d_in), nc, ns, dpitch,
                                                      static_cast<float*>(d_out), layout, st);
                             });
}

int zafb_istft_host_f32(const zafb_stft_plan* p, const float* spec, int64_t n_clips, int64_t nt, int layout, float* y,
                        int64_t y_stride) {
    ZAFB_REQUIRE(p != nullptr, "plan is NULL");
    ZAFB_REQUIRE(n_clips >= 0 && nt >= 0, "bad batch geometry");
    int64_t len = 0;
    zafb_istft_geometry(p->n, nt, p->hop, nullptr, nullptr, &len);
    ZAFB_REQUIRE(y_stride >= len, "y_stride too small");
    if (n_clips == 0 || len == 0) return ZAFB_OK;
    ZAFB_REQUIRE(spec != nullptr && y != nullptr, "spec/y is NULL");
    const size_t in_clip = size_t(nt) * p->n * sizeof(float2);
    const int64_t dpitch = (len + 1) & ~int64_t(1);
    return run_host_pipeline(spec, in_clip, in_clip, in_clip, y, size_t(y_stride) * sizeof(float), size_t(len) * sizeof(float),
                             size_t(dpitch) * sizeof(float), n_clips,
                             [&](void* d_in, void* d_out, int64_t, int64_t nc, cudaStream_t st) {
                                 return zafb_istft_f32(p, static_cast<const float*>(d_in), nc, nt, layout,
                                                       static_cast<float*>(d_out), dpitch, st);
                             });
}

}  // extern "C"
