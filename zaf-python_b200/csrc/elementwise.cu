// Device-side elementwise stages between the transforms: the reference's own demos are chains --
// stft -> time-frequency mask -> istft (zaf.py:162-198) and mdct -> (quantise) -> imdct (zaf.py:1098-1105) -- and with
// these kernels such a chain never leaves HBM.  All of them are streaming, HBM-bound passes (one read per operand, one
// write), grid-stride over 148 x 8 CTAs, 16-byte accesses where the geometry allows.
//
//   zafb_spec_abs_f32     |X| of the first `keep` bins of a complex spectrum            (zaf.py:176-177: abs(X[0:N/2+1, :]))
//   zafb_spec_mask_f32    X * mask, the mask given for all bins or for bins 0 .. N/2 and mirrored onto the rest
//                         (zaf.py:185-186: np.concatenate((m, m[-2:0:-1, :])) * X)
//   zafb_ratio_min_f32    min(a, b) / a                                                 (zaf.py:181-182)
//   zafb_mul_f32          a * b (real masks on MDCT coefficients)
//   zafb_quantize_f32     step * rint(x / step): uniform scalar quantise-dequantise (round half to even, like np.round)
//   zafb_count_mismatch_u32  number of differing 32-bit words of two buffers (bitwise comparison of sharded against
//                         unsharded results without a host round trip)
#include <cstdint>

#include "common.cuh"

using namespace zafb;

namespace {

constexpr int kThreads = 256;

int64_t grid_for(int64_t items) {
    int64_t blocks = ceil_div(items, kThreads);
    const int64_t cap = int64_t(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    return blocks < 1 ? 1 : blocks;
}

// x: blocks over the `items` of one clip, y: clips -- about sm_count() * 8 blocks in all
dim3 grid_2d(int64_t items, int64_t clips) {
    const int64_t cap = int64_t(sm_count()) * 8;
    int64_t gy = clips < cap ? clips : cap;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    int64_t gx = ceil_div(items, kThreads);
    const int64_t per = ceil_div(cap, gy);
    if (gx > per) gx = per;
    if (gx < 1) gx = 1;
    return dim3(unsigned(gx), unsigned(gy));
}

// The four spectrum kernels walk rows (frames, or bin rows of a clip) with one warp per row and no index division: a
// 64-bit division per element made them instruction-bound (the same finding as for the GEMM pre-pass in dctdst.cu).
//
// FRAME_MAJOR: spec[row][bins] (row = clip * frames + frame) -> out[row][keep]
__global__ void abs_frame_major_kernel(const float2* __restrict__ spec, int64_t rows, int bins, int keep, float* __restrict__ out,
                                       int vec) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = int64_t(gridDim.x) * (blockDim.x >> 5);
    for (int64_t r = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += nwarps) {
        const float2* in = spec + r * bins;
        float* o = out + r * keep;
        int k0 = 0;
        if (vec) {  // two bins per access (16-byte loads, 8-byte stores when the rows allow), four accesses in flight per lane
            const float4* in4 = reinterpret_cast<const float4*>(in);
            float2* o2 = reinterpret_cast<float2*>(o);
            const int pairs = keep >> 1;
            for (int p = lane; p < pairs; p += 128) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (p + 32 * u < pairs) v[u] = __ldg(in4 + p + 32 * u);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (p + 32 * u < pairs) {
                        const float m0 = sqrtf(v[u].x * v[u].x + v[u].y * v[u].y), m1 = sqrtf(v[u].z * v[u].z + v[u].w * v[u].w);
                        if (vec == 2) {  // odd `keep`: the output rows are only 4-byte aligned
                            o[2 * (p + 32 * u)] = m0;
                            o[2 * (p + 32 * u) + 1] = m1;
                        } else {
                            o2[p + 32 * u] = make_float2(m0, m1);
                        }
                    }
            }
            k0 = pairs * 2;
        }
        for (int k = k0 + lane; k < keep; k += 32) {
            const float2 v = __ldg(in + k);
            o[k] = sqrtf(v.x * v.x + v.y * v.y);
        }
    }
}

// BIN_MAJOR: spec[clip][bins][frames] -> out[clip][keep][frames]; the kept rows of a clip are one contiguous run
__global__ void abs_bin_major_kernel(const float2* __restrict__ spec, int64_t clips, int64_t bins_frames, int64_t keep_frames,
                                     float* __restrict__ out) {
    const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x, nth = int64_t(gridDim.x) * blockDim.x;
    for (int64_t c = blockIdx.y; c < clips; c += gridDim.y) {
        const float2* in = spec + c * bins_frames;
        float* o = out + c * keep_frames;
        for (int64_t i = tid; i < keep_frames; i += nth) {
            const float2 v = __ldg(in + i);
            o[i] = sqrtf(v.x * v.x + v.y * v.y);
        }
    }
}

// FRAME_MAJOR: out[row][k] = spec[row][k] * mask[row][k <= bins/2 or full ? k : bins - k]
__global__ void mask_frame_major_kernel(const float2* spec, int64_t rows, int bins, const float* __restrict__ mask, int mask_bins,
                                        float2* out, int vec) {
    // spec and out may be the same buffer (in place): every lane reads its own elements before it writes them
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = int64_t(gridDim.x) * (blockDim.x >> 5);
    for (int64_t r = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += nwarps) {
        const float2* in = spec + r * bins;
        float2* o = out + r * bins;
        const float* mr = mask + r * mask_bins;
        int k0 = 0;
        if (vec) {  // two bins per access, four accesses in flight per lane
            const float4* in4 = reinterpret_cast<const float4*>(in);
            float4* o4 = reinterpret_cast<float4*>(o);
            const int pairs = bins >> 1;
            for (int p = lane; p < pairs; p += 128) {
                float4 v[4];
                float m0[4], m1[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int q = p + 32 * u;
                    if (q < pairs) {
                        v[u] = in4[q];
                        const int ka = 2 * q, kb = 2 * q + 1;
                        m0[u] = __ldg(mr + (ka < mask_bins ? ka : bins - ka));
                        m1[u] = __ldg(mr + (kb < mask_bins ? kb : bins - kb));
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (p + 32 * u < pairs) o4[p + 32 * u] = make_float4(v[u].x * m0[u], v[u].y * m0[u], v[u].z * m1[u], v[u].w * m1[u]);
            }
            k0 = pairs * 2;
        }
        for (int k = k0 + lane; k < bins; k += 32) {
            const float m = __ldg(mr + (k < mask_bins ? k : bins - k));
            const float2 v = in[k];
            o[k] = make_float2(v.x * m, v.y * m);
        }
    }
}

// BIN_MAJOR: out[clip][k][j] = spec[clip][k][j] * mask[clip][k < mask_bins ? k : bins - k][j]; one warp per bin row
__global__ void mask_bin_major_kernel(const float2* __restrict__ spec, int64_t clips, int bins, int64_t frames,
                                      const float* __restrict__ mask, int mask_bins, float2* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = int64_t(gridDim.x) * (blockDim.x >> 5);
    for (int64_t c = blockIdx.y; c < clips; c += gridDim.y) {
        for (int64_t k = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); k < bins; k += nwarps) {
            const int km = k < mask_bins ? int(k) : bins - int(k);
            const float2* in = spec + (c * bins + k) * frames;
            float2* o = out + (c * bins + k) * frames;
            const float* mr = mask + (c * mask_bins + km) * frames;
            for (int64_t j = lane; j < frames; j += 32) {
                const float m = __ldg(mr + j);
                const float2 v = in[j];
                o[j] = make_float2(v.x * m, v.y * m);
            }
        }
    }
}

enum { kOpRatioMin = 0, kOpMul = 1, kOpQuantize = 2 };

template <int OP>
__device__ __forceinline__ float apply(float a, float b, float step) {
    if constexpr (OP == kOpRatioMin) return fminf(a, b) / a;
    else if constexpr (OP == kOpMul) return a * b;
    else return step * rintf(a / step);
}

template <int OP>
__global__ void binary_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float step,
                              float* __restrict__ out, int vec4) {
    const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x, nth = int64_t(gridDim.x) * blockDim.x;
    int64_t done = 0;
    if (vec4) {  // 16-byte accesses over the aligned bulk
        const int64_t n4 = n / 4;
        const float4* a4 = reinterpret_cast<const float4*>(a);
        const float4* b4 = reinterpret_cast<const float4*>(OP == kOpQuantize ? a : b);
        float4* o4 = reinterpret_cast<float4*>(out);
        for (int64_t i = tid; i < n4; i += nth) {
            const float4 x = a4[i];
            const float4 y = OP == kOpQuantize ? x : b4[i];
            o4[i] = make_float4(apply<OP>(x.x, y.x, step), apply<OP>(x.y, y.y, step), apply<OP>(x.z, y.z, step),
                                apply<OP>(x.w, y.w, step));
        }
        done = n4 * 4;
    }
    for (int64_t i = done + tid; i < n; i += nth) out[i] = apply<OP>(a[i], OP == kOpQuantize ? 0.f : b[i], step);
}

// two-sided frames from one-sided ones: dst[f][k] = src[f][k] (k <= n/2, skipped when src == dst) and
// dst[f][n - k] = conj(src[f][k]) (1 <= k <= (n - 1) / 2); reads ascend, mirrored writes descend, both coalesced
__global__ void mirror_kernel(const float2* __restrict__ src, int64_t src_pitch, int64_t frames, int n, float2* __restrict__ dst,
                              int copy_low) {
    const int half = n / 2 + 1;
    const int last_mirrored = (n - 1) / 2;
    const int64_t total = frames * half;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int64_t f = i / half;
        const int k = int(i - f * half);
        const float2 v = src[f * src_pitch + k];
        if (copy_low) dst[f * n + k] = v;
        if (k >= 1 && k <= last_mirrored) dst[f * n + (n - k)] = make_float2(v.x, -v.y);
    }
}

__global__ void mismatch_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, int64_t n,
                                unsigned long long* __restrict__ count) {
    unsigned long long local = 0;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
        local += a[i] != b[i];
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

template <int OP>
int launch_binary(const float* a, const float* b, int64_t n, float step, float* out, void* stream) {
    ZAFB_REQUIRE(n >= 0, "element count must be >= 0");
    if (n == 0) return ZAFB_OK;
    ZAFB_REQUIRE(a != nullptr && out != nullptr && (OP == kOpQuantize || b != nullptr), "operand is NULL");
    const bool vec4 = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(out) |
                        (OP == kOpQuantize ? 0 : reinterpret_cast<uintptr_t>(b))) & 15) == 0;
    binary_kernel<OP><<<unsigned(grid_for(vec4 ? n / 4 + 1 : n)), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        a, b, n, step, out, vec4 ? 1 : 0);
    ZAFB_LAUNCH_CHECK();
    return ZAFB_OK;
}

}  // namespace

extern "C" {

int zafb_spec_abs_f32(const float* spec, int64_t n_clips, int64_t bins, int64_t frames, int layout, int64_t keep_bins,
                      float* out, void* stream) {
    ZAFB_REQUIRE(n_clips >= 0 && bins >= 1 && frames >= 0, "bad spectrum geometry");
    ZAFB_REQUIRE(keep_bins >= 1 && keep_bins <= bins, "keep_bins %lld outside [1, %lld]", (long long)keep_bins, (long long)bins);
    ZAFB_REQUIRE(layout == ZAFB_LAYOUT_FRAME_MAJOR || layout == ZAFB_LAYOUT_BIN_MAJOR, "bad layout %d", layout);
    ZAFB_REQUIRE(bins < (int64_t(1) << 30), "too many bins");
    const int64_t total = n_clips * frames * keep_bins;
    if (total == 0) return ZAFB_OK;
    ZAFB_REQUIRE(spec != nullptr && out != nullptr, "spec/out is NULL");
    const float2* s2 = reinterpret_cast<const float2*>(spec);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (layout == ZAFB_LAYOUT_FRAME_MAJOR)
        abs_frame_major_kernel<<<unsigned(grid_for(n_clips * frames * 32)), kThreads, 0, st>>>(
            s2, n_clips * frames, int(bins), int(keep_bins), out,
            (bins % 2 == 0 && reinterpret_cast<uintptr_t>(spec) % 16 == 0)
                ? ((keep_bins % 2 == 0 && reinterpret_cast<uintptr_t>(out) % 8 == 0) ? 1 : 2) : 0);
    else
        abs_bin_major_kernel<<<grid_2d(keep_bins * frames, n_clips), kThreads, 0, st>>>(s2, n_clips, bins * frames, keep_bins * frames, out);
    ZAFB_LAUNCH_CHECK();
    return ZAFB_OK;
}

int zafb_spec_mask_f32(const float* spec, int64_t n_clips, int64_t bins, int64_t frames, int layout, const float* mask,
                       int64_t mask_bins, float* out, void* stream) {
    ZAFB_REQUIRE(n_clips >= 0 && bins >= 1 && frames >= 0, "bad spectrum geometry");
    ZAFB_REQUIRE(layout == ZAFB_LAYOUT_FRAME_MAJOR || layout == ZAFB_LAYOUT_BIN_MAJOR, "bad layout %d", layout);
    ZAFB_REQUIRE(bins < (int64_t(1) << 30), "too many bins");
    // a mask for every bin, or for bins 0 .. floor(N/2) mirrored onto N-k like np.concatenate((m, m[-2:0:-1])) (zaf.py:185)
    ZAFB_REQUIRE(mask_bins == bins || mask_bins == bins / 2 + 1, "mask must have %lld or %lld bins, not %lld", (long long)bins,
                 (long long)(bins / 2 + 1), (long long)mask_bins);
    const int64_t total = n_clips * frames * bins;
    if (total == 0) return ZAFB_OK;
    ZAFB_REQUIRE(spec != nullptr && mask != nullptr && out != nullptr, "spec/mask/out is NULL");
    const float2* s2 = reinterpret_cast<const float2*>(spec);
    float2* o2 = reinterpret_cast<float2*>(out);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (layout == ZAFB_LAYOUT_FRAME_MAJOR)
        mask_frame_major_kernel<<<unsigned(grid_for(n_clips * frames * 32)), kThreads, 0, st>>>(
            s2, n_clips * frames, int(bins), mask, int(mask_bins), o2,
            int(bins % 2 == 0 && reinterpret_cast<uintptr_t>(spec) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0));
    else
        mask_bin_major_kernel<<<grid_2d(bins * kThreads / 8, n_clips), kThreads, 0, st>>>(s2, n_clips, int(bins), frames, mask, int(mask_bins), o2);
    ZAFB_LAUNCH_CHECK();
    return ZAFB_OK;
}

int zafb_ratio_min_f32(const float* a, const float* b, int64_t n, float* out, void* stream) {
    return launch_binary<kOpRatioMin>(a, b, n, 0.f, out, stream);
}

int zafb_mul_f32(const float* a, const float* b, int64_t n, float* out, void* stream) {
    return launch_binary<kOpMul>(a, b, n, 0.f, out, stream);
}

int zafb_quantize_f32(const float* x, int64_t n, float step, float* out, void* stream) {
    ZAFB_REQUIRE(step > 0.f, "quantiser step must be > 0");
    return launch_binary<kOpQuantize>(x, nullptr, n, step, out, stream);
}

int zafb_spec_mirror_f32(const float* src, int64_t src_pitch, int64_t frames, int64_t n, float* dst, void* stream) {
    ZAFB_REQUIRE(frames >= 0 && n >= 1 && n < (int64_t(1) << 30), "bad frame geometry");
    ZAFB_REQUIRE(src_pitch >= n / 2 + 1, "src_pitch %lld < %lld one-sided bins", (long long)src_pitch, (long long)(n / 2 + 1));
    if (frames == 0) return ZAFB_OK;
    ZAFB_REQUIRE(src != nullptr && dst != nullptr, "src/dst is NULL");
    ZAFB_REQUIRE(src != dst || src_pitch == n, "in place needs src_pitch == window_length");
    mirror_kernel<<<unsigned(grid_for(frames * (n / 2 + 1))), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float2*>(src), src_pitch, frames, int(n), reinterpret_cast<float2*>(dst), src != dst ? 1 : 0);
    ZAFB_LAUNCH_CHECK();
    return ZAFB_OK;
}

int zafb_count_mismatch_u32(const void* a, const void* b, int64_t n_words, int64_t* count, void* stream) {
    ZAFB_REQUIRE(n_words >= 0 && count != nullptr, "bad word count / count is NULL");
    *count = 0;
    if (n_words == 0) return ZAFB_OK;
    ZAFB_REQUIRE(a != nullptr && b != nullptr, "buffer is NULL");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long* d = nullptr;
    ZAFB_CUDA(cudaMalloc(reinterpret_cast<void**>(&d), sizeof(*d)));
    cudaError_t e = cudaMemsetAsync(d, 0, sizeof(*d), st);
    if (e == cudaSuccess) {
        mismatch_kernel<<<unsigned(grid_for(n_words)), kThreads, 0, st>>>(static_cast<const uint32_t*>(a),
                                                                           static_cast<const uint32_t*>(b), n_words, d);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        e = cudaGetLastError();
    }
    unsigned long long h = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ZAFB_E_CUDA, "count_mismatch: %s", cudaGetErrorString(e));
    *count = int64_t(h);
    return ZAFB_OK;
}

}  // extern "C"
