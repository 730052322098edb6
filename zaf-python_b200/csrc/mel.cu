// melspectrogram / mfcc (zaf.py:324-375, 378-454): STFT -> |X| or |X|^2 of bins 1..N/2 -> mel filterbank
// -> [ln(. + eps) -> orthonormal DCT-II over the mel axis, rows 1..n_coef], fused in ONE kernel per call:
// the spectrum never goes to HBM (the reference materialises a complex128 (N, nt) STFT and multiplies by the
// densified, 98.7 %-zero filterbank with dgemm).
//
// The filterbank arrives dense (the wrapper's .toarray(), exactly like zaf.py:373) and is packed at plan
// creation into one contiguous band per row [first non-zero column, last non-zero column]: 882 weights
// instead of 65 536 at BASELINE cfg 3.
#include <cmath>
#include <vector>

#include "fft_core.cuh"

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
};

namespace {

constexpr int kMaxDynSmem = 200 * 1024;

// mode 0: melspectrogram (magnitude); mode 1: mfcc (power -> log -> DCT)
__global__ void mel_frame_kernel(const float* __restrict__ x, int64_t ns, int64_t clip_stride, int64_t nt, int64_t hop,
                                 int log2n, const float* __restrict__ window, const float2* __restrict__ tw_half,
                                 const float2* __restrict__ tw_full, const int* __restrict__ band_lo,
                                 const int* __restrict__ band_len, const int* __restrict__ band_off,
                                 const float* __restrict__ weights, const float* __restrict__ dct, int n_mels, int n_coef,
                                 int mode, float* __restrict__ out, int layout, int64_t total_frames) {
    extern __shared__ float2 smem2[];
    const int n = 1 << log2n, m = n >> 1;
    float2* a = smem2;
    float2* b = smem2 + m;
    float* spec = reinterpret_cast<float*>(smem2 + 2 * m);  // m floats: column c <-> FFT bin c + 1
    float* mel = spec + m;                                  // n_mels floats
    const int tid = threadIdx.x, nth = blockDim.x;
    const int rows = mode == 0 ? n_mels : n_coef;
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
            }
        }
        if (mode == 1) {
            __syncthreads();
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

bool g_attr_done = false;
int set_kernel_attrs() {
    if (g_attr_done) return ZAFB_OK;
    ZAFB_CUDA(cudaFuncSetAttribute(mel_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    g_attr_done = true;
    return ZAFB_OK;
}

template <class T>
int upload_vec(T** dev, const std::vector<T>& v) {
    ZAFB_CUDA(cudaMalloc(reinterpret_cast<void**>(dev), (v.size() ? v.size() : 1) * sizeof(T)));
    ZAFB_CUDA(cudaMemcpy(*dev, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return ZAFB_OK;
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
    const int m = int(p->n / 2);
    const size_t smem = size_t(p->n) * sizeof(float2) + size_t(m + p->n_mels) * sizeof(float) + 16;
    if (smem > size_t(kMaxDynSmem)) return fail(ZAFB_E_UNSUPPORTED, "mel: window_length %lld too large", (long long)p->n);
    int th = m / 4;
    if (th < 64) th = 64;
    if (th > 256) th = 256;
    const int64_t grid = total < int64_t(sm_count()) * 32 ? total : int64_t(sm_count()) * 32;
    mel_frame_kernel<<<unsigned(grid), th, smem, static_cast<cudaStream_t>(stream)>>>(
        x, ns, clip_stride, nt, p->hop, p->log2n, p->d_window, p->d_tw_half, p->d_tw_full, p->d_band_lo, p->d_band_len,
        p->d_band_off, p->d_weights, p->d_dct, int(p->n_mels), int(p->n_coef), mode, out, layout, total);
    ZAFB_LAUNCH_CHECK();
    return ZAFB_OK;
}

}  // namespace

extern "C" {

int zafb_mel_plan_create(zafb_mel_plan** out, const double* window, int64_t n, int64_t hop, const double* fb,
                         int64_t n_mels, int64_t n_coef) {
    ZAFB_REQUIRE(out != nullptr && window != nullptr && fb != nullptr, "plan/window/filterbank is NULL");
    ZAFB_REQUIRE(n >= 2 && hop >= 1 && n_mels >= 1 && n_coef >= 0, "bad mel plan parameters");
    if (!is_pow2(n) || n < 4)
        return fail(ZAFB_E_UNSUPPORTED, "melspectrogram/mfcc: window_length %lld is not a power of two >= 4", (long long)n);
    if (n_mels > 4096) return fail(ZAFB_E_UNSUPPORTED, "too many mel filters (%lld)", (long long)n_mels);
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
    if (rc == ZAFB_OK) rc = upload_twiddles(&p->d_tw_half, n / 2, n / 2);
    if (rc == ZAFB_OK) rc = upload_twiddles(&p->d_tw_full, n, n / 2);
    if (rc == ZAFB_OK) rc = upload_vec(&p->d_band_lo, lo);
    if (rc == ZAFB_OK) rc = upload_vec(&p->d_band_len, len);
    if (rc == ZAFB_OK) rc = upload_vec(&p->d_band_off, off);
    if (rc == ZAFB_OK) rc = upload_vec(&p->d_weights, w);
    if (rc == ZAFB_OK) rc = upload_vec(&p->d_dct, d);
    if (rc != ZAFB_OK) {
        zafb_mel_plan_destroy(p);
        return rc;
    }
    *out = p;
    return ZAFB_OK;
}

int zafb_mel_plan_destroy(zafb_mel_plan* p) {
    if (!p) return ZAFB_OK;
    cudaFree(p->d_window);
    cudaFree(p->d_tw_half);
    cudaFree(p->d_tw_full);
    cudaFree(p->d_band_lo);
    cudaFree(p->d_band_len);
    cudaFree(p->d_band_off);
    cudaFree(p->d_weights);
    cudaFree(p->d_dct);
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

}  // extern "C"
