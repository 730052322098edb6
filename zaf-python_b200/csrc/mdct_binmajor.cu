// MDCT / IMDCT straight to and from BIN_MAJOR memory [clip][bin][frame] -- the reference's C order (zaf.py:1073, 1159) --
// for window lengths 2048 (BASELINE cfg 4) and 1024.  Same arithmetic as mdct_warp_kernel / imdct_warp_kernel (mdct.cu):
// one warp transforms one frame through an M/2-point complex FFT in its registers.  What differs is the memory side:
//
//   mdct_binmajor_kernel<N>    a CTA of 16 warps walks along one clip in tiles of 32 consecutive frames (two per warp) and
//                              parks every frame's M coefficients in the frame's slot of a ring of 32 + 7 shared-memory
//                              regions; after a barrier warp u stores rows u, u + 16, ...: the 32 frames of a row leave
//                              as one 128-byte run.  Rows are nt * 4 bytes long, so a run that starts on a tile boundary
//                              starts in the middle of a 32-byte sector in seven rows out of eight when nt is odd (1293
//                              at cfg 4); as in stft_warp_binmajor_kernel every row therefore gets its own window
//                              [j0 - s, j0 + 32 - s), s = (element address of (row, j0)) mod 8, which starts every run on
//                              a sector boundary -- the up to 7 frames before j0 are still in the ring.
//   imdct_binmajor_kernel<N>   the mirror image: thread (u, w) reads frame j0 + w of rows u, u + 16, ... (128-byte runs)
//                              into the ring, warp w transforms frames j0 + 2w and j0 + 2w + 1 with the TDAC carry
//                              between them in registers, parks the halves that meet a frame of another warp, and after a
//                              barrier adds them in the reference's order (frame h-1, then frame h) -- no atomics.
//
// A slot is M + 1 floats wide at least and the slot pitch is 1 mod 32, so the 32 frames of one bin sit in 32 distinct
// banks; the slot doubles as the warp FFT's transpose tile (float2: it starts at the slot's first 8-byte boundary).
#include "mdct_common.cuh"

namespace zafb {
namespace {

constexpr int kBmWarps = 16;
constexpr int kBmMaxDynSmem = 200 * 1024;

template <int N, int AHEAD>
struct MdctBmGeom {
    using G = MdctGeom<N>;
    static constexpr int F = 32;               // frames per tile: 32 floats of a row are one 128-byte run
    static constexpr int SLOTS = F + AHEAD;    // ring: the tile plus the frames before it that are still needed
    static constexpr int NEED = (2 * G::TILE > G::M ? 2 * G::TILE : G::M) + 1;  // floats; + 1: the tile's 8-byte alignment
    static constexpr int PITCH = ((NEED + 30) / 32) * 32 + 1;
    static_assert(PITCH >= NEED && PITCH % 32 == 1, "bad slot pitch");
    static constexpr size_t SMEM = size_t(G::TABLES) * sizeof(float2) + size_t(SLOTS) * PITCH * sizeof(float);
    static_assert(SMEM <= size_t(kBmMaxDynSmem), "ring does not fit");
};

template <int N>
__device__ __forceinline__ void bm_load_tables(float2* smem, const float2* __restrict__ win_pairs,
                                               const float2* __restrict__ tw4, int tid, float win_scale) {
    using G = MdctGeom<N>;
    for (int i = tid; i < G::M; i += kBmWarps * 32) {
        const float2 w = win_pairs[i];
        smem[i] = make_float2(w.x * win_scale, w.y * win_scale);
    }
    for (int i = tid; i < G::H; i += kBmWarps * 32) smem[G::M + i] = tw4[i];
}

// the slot's transpose tile / float2 view: the first 8-byte boundary of the slot (slots start at odd float offsets
// every other time because the pitch is odd)
__device__ __forceinline__ float* slot_aligned(float* slot) {
    return slot + ((__cvta_generic_to_shared(slot) >> 2) & 1);
}

// A pair (a[2P], a[2P+1]) of a slot whose base may sit at an odd float offset: two 4-byte accesses.  Lanes 0-15 touch
// the even element first and lanes 16-31 the odd one, so every access of the warp covers 32 distinct banks.
__device__ __forceinline__ void put_pair(float* s, int P, int lane, float2 v) {
    const int hi = lane >> 4;
    s[2 * P + hi] = hi ? v.y : v.x;
    s[2 * P + (hi ^ 1)] = hi ? v.x : v.y;
}
__device__ __forceinline__ float2 get_pair(const float* s, int P, int lane) {
    const int hi = lane >> 4;
    const float a = s[2 * P + hi];
    const float b = s[2 * P + (hi ^ 1)];
    return hi ? make_float2(b, a) : make_float2(a, b);
}

template <int N>
__global__ void __launch_bounds__(kBmWarps * 32, 1)
mdct_binmajor_kernel(const float* __restrict__ x, int64_t ns, int64_t clip_stride, int nt,
                     const float2* __restrict__ win_pairs, const float2* __restrict__ tw4,
                     const float2* __restrict__ pre, const float2* __restrict__ post, float* __restrict__ out,
                     int phase0, int runs_per_clip, int tiles_per_run, int64_t total_runs) {
    using G = MdctGeom<N>;
    using B = MdctBmGeom<N, 7>;
    constexpr int M = G::M, REGS = G::REGS, LOGR = G::LOGR, Q = G::Q, QR = REGS / 2;
    constexpr int F = B::F, SLOTS = B::SLOTS, PITCH = B::PITCH;
    extern __shared__ __align__(16) float2 smem2[];
    const float2* s_win = smem2;
    const float2* s_tw = smem2 + M;
    float* s_ring = reinterpret_cast<float*>(smem2 + G::TABLES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    bm_load_tables<N>(smem2, win_pairs, tw4, tid, 1.0f);
    const float2 c_lane = pre[lane];
    const float2 p_lane = post[lane];
    float2 tq[G::NTQ];
    if constexpr (N == 1024) warp_fft256_lane_twiddles(tq, lane);
    __syncthreads();

    // store phase: warp u owns rows u + 16 i, lane w the frame j0 - s + w of the row's window.  Element (b, j0) sits at
    // float index phase0 + b * nt + j0 (mod 8) of its sector; j0 and 16 i are multiples of 8, so s is one constant per warp.
    const int s_row = (phase0 + (warp & 7) * (nt & 7)) & 7;
    const int tiles_per_clip = (nt + F - 1) / F;

    for (int64_t run = blockIdx.x; run < total_runs; run += gridDim.x) {
        const int64_t clip = run / runs_per_clip;
        const int t0 = int(run - clip * runs_per_clip) * tiles_per_run;
        const int t1 = min(t0 + tiles_per_run, tiles_per_clip);
        const int jlo = t0 * F, jhi = min(nt, t1 * F);
        const float* xc = x + clip * clip_stride;
        float* oc = out + clip * int64_t(M) * nt;

        for (int t = t0; t <= t1; ++t) {  // t == t1: flush of the frames still waiting for their row's window
            const int j0 = t * F;
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const int j = j0 + 2 * warp + h;
                if (t >= t1 || j >= nt) break;  // warp-uniform
                const int64_t start = int64_t(j - 1) * M;  // frame j covers original samples [(j-1)M, (j+1)M)
                // (issuing the second frame's loads right after the first frame's windowing step keeps 32 more registers
                // live through the FFT: 208 bytes of spills, 5.86 -> 7.84 ms on cfg 4 -- measured, not kept)
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
                // the half frame the NEXT tile's frame j + F adds, towards L2 while this tile is transformed and stored
                if (h == 1 && t + 1 < t1) {
                    const int64_t nx = start + int64_t(F) * M;  // two frames' worth of new samples: [nx, nx + 2M) for this warp
                    for (int64_t s = nx + 32 * lane; s < nx + 2 * M && s < ns; s += 32 * 32) prefetch_l2(xc + s);
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
                    v[r] = make_float2(-p2.y - p1.x, p3.y - p4.x);
                    const float2 other = make_float2(p3.x - p4.y, -p2.x - p1.y);
                    v[REGS - 1 - r].x = __shfl_xor_sync(0xffffffffu, other.x, 31);
                    v[REGS - 1 - r].y = __shfl_xor_sync(0xffffffffu, other.y, 31);
                });
                static_for<0, REGS>([&](auto rc) {
                    constexpr int r = decltype(rc)::value;
                    v[r] = cmul(v[r], mul_tw<r, G::TWDEN>(c_lane));
                });
                float* slot = s_ring + (j % SLOTS) * PITCH;
                mdct_warp_fft<N>(v, s_tw, reinterpret_cast<float2*>(slot_aligned(slot)), lane, tq);
                static_for<0, REGS>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    v[bitrev(k, LOGR)] = cmul(v[bitrev(k, LOGR)], mul_tw<k, G::TWDEN>(p_lane));
                });
                __syncwarp();  // every lane is done with the transpose tile
                static_for<0, REGS>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    const float im = __shfl_xor_sync(0xffffffffu, v[bitrev(REGS - 1 - k, LOGR)].y, 31);
                    put_pair(slot, lane + 32 * k, lane, make_float2(v[bitrev(k, LOGR)].x, -im));  // (X[2P], X[2P+1])
                });
            }
            __syncthreads();
            // (16-byte stores -- lane (rr, q) four frames of rows 4 u + rr + 64 i -- were measured: 5.86 -> 6.51 ms on cfg 4)
            const int ja = j0 - s_row + lane;
            if (ja >= jlo && ja < jhi) {
                const float* r = s_ring + (ja % SLOTS) * PITCH + warp;
                float* o = oc + int64_t(warp) * nt + ja;
                const int64_t step = int64_t(kBmWarps) * nt;
                constexpr int kBatch = 8;
#pragma unroll 1
                for (int i = 0; i < M / kBmWarps; i += kBatch, o += kBatch * step) {
                    float val[kBatch];
#pragma unroll
                    for (int b = 0; b < kBatch; ++b) val[b] = r[kBmWarps * (i + b)];
#pragma unroll
                    for (int b = 0; b < kBatch; ++b) o[b * step] = val[b];
                }
            }
            __syncthreads();  // the ring slots of the next tile become transpose tiles again
        }
    }
}

// IMDCT from BIN_MAJOR memory.  One CTA walks a whole clip (hop-block h = second half of frame h-1 + first half of frame h).
template <int N>
__global__ void __launch_bounds__(kBmWarps * 32, 1)
imdct_binmajor_kernel(const float* __restrict__ spec, int nt, const float2* __restrict__ win_pairs,
                      const float2* __restrict__ tw4, const float2* __restrict__ pre, const float2* __restrict__ post,
                      int64_t n_clips, int64_t out_len, float* __restrict__ y, int64_t y_stride, int y_aligned) {
    using G = MdctGeom<N>;
    using B = MdctBmGeom<N, 1>;
    constexpr int M = G::M, H = G::H, REGS = G::REGS, LOGR = G::LOGR, HR = REGS / 2;
    constexpr int F = B::F, SLOTS = B::SLOTS, PITCH = B::PITCH;
    extern __shared__ __align__(16) float2 smem2[];
    const float2* s_win = smem2;
    const float2* s_tw = smem2 + M;
    float* s_ring = reinterpret_cast<float*>(smem2 + G::TABLES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr float kScale = 2.0f / float(M);  // a power of two: folding it into the window table changes no bit
    bm_load_tables<N>(smem2, win_pairs, tw4, tid, kScale);
    const float2 c_lane = pre[lane];
    const float2 p_lane = post[lane];
    float2 tq[G::NTQ];
    if constexpr (N == 1024) warp_fft256_lane_twiddles(tq, lane);
    __syncthreads();
    const int tiles = (nt + F - 1) / F;

    // store of one hop-block pair position: o = (samples 2P, 2P+1 of hop-block hb), hb >= 1
    auto emit = [&](float* yc, int hb, int P, float2 o) {
        const int64_t idx = int64_t(hb - 1) * M + 2 * P;  // the reference trims the first M samples (zaf.py:1182)
        if (y_aligned && idx + 1 < out_len) {
            __stcs(reinterpret_cast<float2*>(yc + idx), o);
        } else {
            if (idx < out_len) yc[idx] = o.x;
            if (idx + 1 < out_len) yc[idx + 1] = o.y;
        }
    };

    for (int64_t clip = blockIdx.x; clip < n_clips; clip += gridDim.x) {
        const float* sc = spec + clip * int64_t(M) * nt;
        float* yc = y + clip * y_stride;
        for (int t = 0; t < tiles; ++t) {
            const int j0 = t * F;
            // load phase: lane w reads frame j0 + w of rows warp + 16 i (one 128-byte run per row and warp).  Measured and
            // not kept: requesting the next tile's values into registers before the combine phase (9.1 -> 9.5 ms on cfg 4,
            // the 64 live registers spill), an L2 prefetch of the tile after it (10.6 ms), and a variant with 16-frame
            // tiles whose next tile arrives by cp.async (4-byte LDGSTS) during the transform (8.26 ms against 8.13;
            // ncu: MIO throttle is the top stall, 1 666 warp instructions per frame, IPC 1.65 -- DESIGN.md 4.1d).
            {
                const int j = j0 + lane;
                if (j < nt) {
                    float* dst = s_ring + (j % SLOTS) * PITCH + warp;
                    const float* src = sc + int64_t(warp) * nt + j;
                    const int64_t step = int64_t(kBmWarps) * nt;
                    constexpr int kBatch = M / kBmWarps;  // every row of the tile in flight at once (16 at a time: 9.1 ms on cfg 4, 32: 8.4)
#pragma unroll 1
                    for (int i = 0; i < M / kBmWarps; i += kBatch, src += kBatch * step) {
                        float val[kBatch];
#pragma unroll
                        for (int b = 0; b < kBatch; ++b) val[b] = __ldg(src + b * step);
#pragma unroll
                        for (int b = 0; b < kBatch; ++b) dst[kBmWarps * (i + b)] = val[b];
                    }
                }
            }
            __syncthreads();
            const int ja = j0 + 2 * warp;  // this warp's frames: ja and ja + 1
            float2 carry[REGS];
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const int j = ja + h;
                if (j >= nt) break;  // warp-uniform
                float* slot = s_ring + (j % SLOTS) * PITCH;
                float* al = slot_aligned(slot);
                float2 xp[REGS], v[REGS];
#pragma unroll
                for (int r = 0; r < REGS; ++r) xp[r] = get_pair(slot, lane + 32 * r, lane);
                __syncwarp();  // the slot becomes the transpose tile
                static_for<0, REGS>([&](auto rc) {
                    constexpr int r = decltype(rc)::value;
                    const float im = __shfl_xor_sync(0xffffffffu, xp[REGS - 1 - r].y, 31);  // X[M - 1 - 2m]
                    v[r] = cmul(make_float2(xp[r].x, im), mul_tw<r, G::TWDEN>(c_lane));
                });
                mdct_warp_fft<N>(v, s_tw, reinterpret_cast<float2*>(al), lane, tq);
                static_for<0, REGS>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    const float2 tt = cmul(v[bitrev(k, LOGR)], mul_tw<k, G::TWDEN>(p_lane));
                    v[bitrev(k, LOGR)] = make_float2(tt.x, -tt.y);
                });
                __syncwarp();  // every lane is done with the transpose tile
                float2* park = reinterpret_cast<float2*>(al);
                static_for<0, REGS>([&](auto rc) {
                    constexpr int rho = decltype(rc)::value;
                    constexpr int own = rho < HR ? rho + HR : rho - HR;
                    constexpr int oth = rho < HR ? HR - 1 - rho : 3 * HR - 1 - rho;
                    const float2 mine = v[bitrev(own, LOGR)];
                    float2 part;
                    part.x = __shfl_xor_sync(0xffffffffu, v[bitrev(oth, LOGR)].x, 31);
                    part.y = __shfl_xor_sync(0xffffffffu, v[bitrev(oth, LOGR)].y, 31);
                    float2 first, second;
                    if constexpr (rho < HR) {
                        first = make_float2(mine.x, part.y);
                        second = make_float2(-mine.y, -part.x);
                    } else {
                        first = make_float2(-mine.y, -part.x);
                        second = make_float2(-mine.x, -part.y);
                    }
                    const int P = lane + 32 * rho;
                    const float2 w2 = s_win[H + P];
                    const float2 sec = make_float2(w2.x * second.x, w2.y * second.y);
                    if (h == 0) {
                        park[P] = first;  // unwindowed: the combine phase applies fma(w, first, carry) like imdct_warp_kernel
                        carry[rho] = sec;
                    } else {
                        const float2 w1 = s_win[P];
                        emit(yc, j, P, make_float2(fmaf(w1.x, first.x, carry[rho].x), fmaf(w1.y, first.y, carry[rho].y)));
                        park[P] = sec;    // meets the first half of frame j + 1 (the next warp's, or the next tile's)
                    }
                });
            }
            __syncthreads();
            // combine: hop-block ja = second half of frame ja - 1 (parked by the previous warp, or by warp 15 of the
            // previous tile) + first half of frame ja (parked by this warp)
            if (ja >= 1 && ja < nt) {
                const float2* fa = reinterpret_cast<const float2*>(slot_aligned(s_ring + (ja % SLOTS) * PITCH));
                const float2* sb = reinterpret_cast<const float2*>(slot_aligned(s_ring + ((ja - 1) % SLOTS) * PITCH));
#pragma unroll
                for (int rho = 0; rho < REGS; ++rho) {
                    const int P = lane + 32 * rho;
                    const float2 first = fa[P], cr = sb[P], w1 = s_win[P];
                    emit(yc, ja, P, make_float2(fmaf(w1.x, first.x, cr.x), fmaf(w1.y, first.y, cr.y)));
                }
            }
            __syncthreads();  // the next tile's load phase overwrites the slots just read
        }
    }
}

template <int N>
int launch_mdct_bm(const zafb_mdct_plan* p, const float* x, int64_t n_clips, int64_t ns, int64_t clip_stride, int64_t nt,
                   float* out, cudaStream_t st) {
    using B = MdctBmGeom<N, 7>;
    static bool attr = false;
    if (!attr) {
        ZAFB_CUDA((cudaFuncSetAttribute(mdct_binmajor_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBmMaxDynSmem)));
        attr = true;
    }
    // runs of consecutive tiles of one clip: whole clips when there are enough of them, else about four runs per SM
    const int64_t sms = sm_count();
    const int64_t tiles_per_clip = ceil_div(nt, B::F);
    int64_t runs_per_clip = n_clips >= 4 * sms ? 1 : std::min<int64_t>(tiles_per_clip, ceil_div(4 * sms, n_clips));
    if (const int forced = env_flag("ZAFB_MDCT_BM_RUNS_PER_CLIP", 0); forced > 0)  // tests
        runs_per_clip = std::min<int64_t>(tiles_per_clip, forced);
    const int64_t tiles_per_run = ceil_div(tiles_per_clip, runs_per_clip);
    runs_per_clip = ceil_div(tiles_per_clip, tiles_per_run);
    const int64_t runs = n_clips * runs_per_clip;
    const int64_t ctas = std::min<int64_t>(sms, runs);
    const int phase0 = int((reinterpret_cast<uintptr_t>(out) >> 2) & 7);
    mdct_binmajor_kernel<N><<<unsigned(ctas), kBmWarps * 32, B::SMEM, st>>>(
        x, ns, clip_stride, int(nt), reinterpret_cast<const float2*>(p->d_window), p->d_tw_4step, p->d_pre, p->d_post, out,
        phase0, int(runs_per_clip), int(tiles_per_run), runs);
    ZAFB_LAUNCH_CHECK();
    return ZAFB_OK;
}

template <int N>
int launch_imdct_bm(const zafb_mdct_plan* p, const float* spec, int64_t n_clips, int64_t nt, int64_t out_len, float* y,
                    int64_t y_stride, cudaStream_t st) {
    using B = MdctBmGeom<N, 1>;
    static bool attr = false;
    if (!attr) {
        ZAFB_CUDA((cudaFuncSetAttribute(imdct_binmajor_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBmMaxDynSmem)));
        attr = true;
    }
    const int64_t ctas = std::min<int64_t>(sm_count(), n_clips);
    const int y_aligned = (reinterpret_cast<uintptr_t>(y) % 8 == 0 && (n_clips <= 1 || y_stride % 2 == 0)) ? 1 : 0;
    imdct_binmajor_kernel<N><<<unsigned(ctas), kBmWarps * 32, B::SMEM, st>>>(
        spec, int(nt), reinterpret_cast<const float2*>(p->d_window), p->d_tw_4step, p->d_pre, p->d_post, n_clips, out_len, y,
        y_stride, y_aligned);
    ZAFB_LAUNCH_CHECK();
    return ZAFB_OK;
}

}  // namespace

// x: 8-byte aligned rows (even clip_stride); out: any 4-byte aligned address.  Window lengths 2048 and 1024 only.
bool mdct_binmajor_supported(const zafb_mdct_plan* p, int64_t nt) {
    return (p->n == 2048 || p->n == 1024) && nt >= 1 && nt < (int64_t(1) << 30);
}

int mdct_binmajor_launch(const zafb_mdct_plan* p, const float* x, int64_t n_clips, int64_t ns, int64_t clip_stride, int64_t nt,
                         float* out, cudaStream_t st) {
    return p->n == 2048 ? launch_mdct_bm<2048>(p, x, n_clips, ns, clip_stride, nt, out, st)
                        : launch_mdct_bm<1024>(p, x, n_clips, ns, clip_stride, nt, out, st);
}

int imdct_binmajor_launch(const zafb_mdct_plan* p, const float* spec, int64_t n_clips, int64_t nt, int64_t out_len, float* y,
                          int64_t y_stride, cudaStream_t st) {
    return p->n == 2048 ? launch_imdct_bm<2048>(p, spec, n_clips, nt, out_len, y, y_stride, st)
                        : launch_imdct_bm<1024>(p, spec, n_clips, nt, out_len, y, y_stride, st);
}

}  // namespace zafb
