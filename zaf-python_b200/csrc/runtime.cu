// Runtime plumbing of libzafb200: device, memory, streams, events, error strings, and the
// bit-exact integer bookkeeping of the reference's framing (no device needed for the latter).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "host_pipe.cuh"

namespace zafb {

std::string& last_error_ref() {
    static thread_local std::string msg;
    return msg;
}

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return code;
}

std::atomic<int64_t> g_launches{0};
std::atomic<int64_t> g_h2d_bytes{0}, g_d2h_bytes{0};

int upload_f32(float** dev, const double* host, size_t n) {
    std::vector<float> tmp(n ? n : 1);
    for (size_t i = 0; i < n; ++i) tmp[i] = static_cast<float>(host[i]);
    ZAFB_CUDA(cudaMalloc(reinterpret_cast<void**>(dev), (n ? n : 1) * sizeof(float)));
    ZAFB_CUDA(cudaMemcpy(*dev, tmp.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    return ZAFB_OK;
}

int upload_c32(float2** dev, const double* host_ri, size_t n) {
    std::vector<float2> tmp(n ? n : 1);
    for (size_t i = 0; i < n; ++i)
        tmp[i] = make_float2(static_cast<float>(host_ri[2 * i]), static_cast<float>(host_ri[2 * i + 1]));
    ZAFB_CUDA(cudaMalloc(reinterpret_cast<void**>(dev), (n ? n : 1) * sizeof(float2)));
    ZAFB_CUDA(cudaMemcpy(*dev, tmp.data(), n * sizeof(float2), cudaMemcpyHostToDevice));
    return ZAFB_OK;
}

int upload_twiddles(float2** dev, int64_t n, int64_t count) {
    std::vector<double> t(2 * static_cast<size_t>(count > 0 ? count : 1));
    const double pi = 3.14159265358979323846264338327950288;
    for (int64_t i = 0; i < count; ++i) {
        const double a = -2.0 * pi * static_cast<double>(i % n) / static_cast<double>(n);
        t[2 * i] = std::cos(a);
        t[2 * i + 1] = std::sin(a);
    }
    return upload_c32(dev, t.data(), static_cast<size_t>(count));
}

int env_flag(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

int sm_count() {
    static int cached = 0;
    if (cached) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    cached = n;
    return n;
}

// ------------------------------------------------------------------ PCM input (zaf.py:1199-1202)
// wavread normalises integer PCM by 2^(8 itemsize - 1); here that happens on the GPU so that only the 16-bit samples
// cross PCIe: interleaved (frame, channel) int16 -> planar [channel][frame] fp32 (x / 32768 is exact in fp32), or the
// channel mean (the reference examples' np.mean(audio_signal, 1)) when `mono` is set.
__global__ void pcm16_to_f32_kernel(const int16_t* __restrict__ pcm, int64_t frames, int channels, int mono,
                                    float* __restrict__ out, int64_t out_stride) {
    const float scale = 1.0f / 32768.0f;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < frames; i += int64_t(gridDim.x) * blockDim.x) {
        const int16_t* p = pcm + i * channels;
        if (mono) {
            float acc = 0.f;
            for (int c = 0; c < channels; ++c) acc += float(p[c]) * scale;
            out[i] = acc / float(channels);
        } else {
            for (int c = 0; c < channels; ++c) out[int64_t(c) * out_stride + i] = float(p[c]) * scale;
        }
    }
}

// ------------------------------------------------------------------ host-buffer pipeline state
HostPipe& host_pipe() {
    static HostPipe hp;
    return hp;
}

size_t host_pipe_chunk_bytes() {
    int mb = env_flag("ZAFB_PIPE_CHUNK_MB", 64);
    if (mb < 1) mb = 1;
    return size_t(mb) << 20;
}

void HostPipe::release() {
    for (int i = 0; i < kStages; ++i) {
        if (st[i]) cudaStreamDestroy(st[i]);
        cudaFree(d_in[i]);
        cudaFree(d_out[i]);
        st[i] = nullptr;
        d_in[i] = d_out[i] = nullptr;
    }
    in_cap = out_cap = 0;
    device = -1;
}

int HostPipe::ensure(size_t in_bytes, size_t out_bytes) {
    int dev = 0;
    ZAFB_CUDA(cudaGetDevice(&dev));
    if (dev != device) {  // first use, or the process switched devices
        if (device >= 0) {
            cudaSetDevice(device);
            release();
            cudaSetDevice(dev);
        }
        for (int i = 0; i < kStages; ++i) ZAFB_CUDA(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
        device = dev;
    }
    if (in_bytes > in_cap) {
        for (int i = 0; i < kStages; ++i) {
            cudaFree(d_in[i]);
            d_in[i] = nullptr;
        }
        in_cap = 0;
        for (int i = 0; i < kStages; ++i) ZAFB_CUDA(cudaMalloc(&d_in[i], in_bytes));
        in_cap = in_bytes;
    }
    if (out_bytes > out_cap) {
        for (int i = 0; i < kStages; ++i) {
            cudaFree(d_out[i]);
            d_out[i] = nullptr;
        }
        out_cap = 0;
        for (int i = 0; i < kStages; ++i) ZAFB_CUDA(cudaMalloc(&d_out[i], out_bytes));
        out_cap = out_bytes;
    }
    return ZAFB_OK;
}

}  // namespace zafb

using namespace zafb;

extern "C" {

const char* zafb_last_error(void) { return last_error_ref().c_str(); }
const char* zafb_version(void) { return "zafb200 0.1 (sm_100a)"; }

int zafb_device_count(int* count) {
    ZAFB_REQUIRE(count != nullptr, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        cudaGetLastError();
        return fail(ZAFB_E_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    }
    *count = n;
    return ZAFB_OK;
}

int zafb_init(int device) {
    // One process drives one GPU (one process per GPU is the multi-GPU model): plans, the host pipeline, the
    // per-kernel shared-memory attributes and the SM count are per-device state cached for the life of the process,
    // so a second zafb_init on ANOTHER device is refused instead of silently reusing device-0 tables.
    static std::atomic<int> bound{-1};
    int expected = -1;
    if (!bound.compare_exchange_strong(expected, device) && expected != device)
        return fail(ZAFB_E_UNSUPPORTED, "this process is bound to device %d; zafb_init(%d) refused (one process per GPU)",
                    expected, device);
    ZAFB_CUDA(cudaSetDevice(device));
    ZAFB_CUDA(cudaFree(nullptr));
    int major = 0;
    ZAFB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10)
        return fail(ZAFB_E_UNSUPPORTED, "device %d has compute capability %d.x; this library is sm_100a only",
                    device, major);
    return ZAFB_OK;
}

int zafb_shutdown(void) {
    HostPipe& hp = host_pipe();
    std::lock_guard<std::mutex> lock(hp.mu);
    hp.release();
    return ZAFB_OK;
}

int zafb_pcm16_to_f32(const int16_t* pcm_dev, int64_t frames, int channels, int mono, float* out_dev, int64_t out_stride,
                      void* stream) {
    ZAFB_REQUIRE(frames >= 0 && channels >= 1 && channels <= 64, "pcm: need frames >= 0 and 1..64 channels");
    ZAFB_REQUIRE(mono || out_stride >= frames, "pcm: out_stride %lld < frames %lld", (long long)out_stride, (long long)frames);
    if (frames == 0) return ZAFB_OK;
    ZAFB_REQUIRE(pcm_dev != nullptr && out_dev != nullptr, "pcm/out is NULL");
    int64_t blocks = ceil_div(frames, 256);
    if (blocks > int64_t(sm_count()) * 16) blocks = int64_t(sm_count()) * 16;
    pcm16_to_f32_kernel<<<unsigned(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(pcm_dev, frames, channels, mono, out_dev,
                                                                                       out_stride);
    ZAFB_LAUNCH_CHECK();
    return ZAFB_OK;
}

int zafb_device_info(int device, int* sm, int* cc_major, int* cc_minor, size_t* total_mem, char* name,
                     size_t name_len) {
    cudaDeviceProp p;
    ZAFB_CUDA(cudaGetDeviceProperties(&p, device));
    if (sm) *sm = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = p.totalGlobalMem;
    if (name && name_len) {
        strncpy(name, p.name, name_len - 1);
        name[name_len - 1] = 0;
    }
    return ZAFB_OK;
}

int zafb_malloc(void** p, size_t bytes) {
    ZAFB_REQUIRE(p != nullptr, "dev_ptr is NULL");
    ZAFB_CUDA(cudaMalloc(p, bytes ? bytes : 1));
    return ZAFB_OK;
}
int zafb_free(void* p) {
    ZAFB_CUDA(cudaFree(p));
    return ZAFB_OK;
}
int zafb_host_alloc(void** p, size_t bytes) {
    ZAFB_REQUIRE(p != nullptr, "host_ptr is NULL");
    ZAFB_CUDA(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault));
    return ZAFB_OK;
}
int zafb_host_free(void* p) {
    ZAFB_CUDA(cudaFreeHost(p));
    return ZAFB_OK;
}
int zafb_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
    ZAFB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
    return ZAFB_OK;
}
int zafb_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
    ZAFB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    return ZAFB_OK;
}
int zafb_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream) {
    ZAFB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    return ZAFB_OK;
}
int zafb_memcpy2d(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes, size_t rows, int kind,
                  void* stream) {
    ZAFB_REQUIRE(kind >= 0 && kind <= 2, "memcpy2d: kind must be 0 (host to device), 1 (device to host) or 2 (device to device)");
    ZAFB_REQUIRE(dst_pitch >= width_bytes && src_pitch >= width_bytes, "memcpy2d: pitch smaller than the row width");
    if (width_bytes == 0 || rows == 0) return ZAFB_OK;
    const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    ZAFB_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, rows, k, static_cast<cudaStream_t>(stream)));
    return ZAFB_OK;
}
int zafb_memset(void* p, int value, size_t bytes, void* stream) {
    ZAFB_CUDA(cudaMemsetAsync(p, value, bytes, static_cast<cudaStream_t>(stream)));
    return ZAFB_OK;
}

int zafb_stream_create(void** s) {
    ZAFB_REQUIRE(s != nullptr, "stream is NULL");
    cudaStream_t st;
    ZAFB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    *s = st;
    return ZAFB_OK;
}
int zafb_stream_destroy(void* s) {
    ZAFB_CUDA(cudaStreamDestroy(static_cast<cudaStream_t>(s)));
    return ZAFB_OK;
}
int zafb_stream_sync(void* s) {
    ZAFB_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(s)));
    return ZAFB_OK;
}
int zafb_device_sync(void) {
    ZAFB_CUDA(cudaDeviceSynchronize());
    return ZAFB_OK;
}
int zafb_event_create(void** e) {
    ZAFB_REQUIRE(e != nullptr, "event is NULL");
    cudaEvent_t ev;
    ZAFB_CUDA(cudaEventCreate(&ev));
    *e = ev;
    return ZAFB_OK;
}
int zafb_event_destroy(void* e) {
    ZAFB_CUDA(cudaEventDestroy(static_cast<cudaEvent_t>(e)));
    return ZAFB_OK;
}
int zafb_event_record(void* e, void* s) {
    ZAFB_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(e), static_cast<cudaStream_t>(s)));
    return ZAFB_OK;
}
int zafb_event_sync(void* e) {
    ZAFB_CUDA(cudaEventSynchronize(static_cast<cudaEvent_t>(e)));
    return ZAFB_OK;
}
int zafb_stream_wait_event(void* s, void* e) {
    ZAFB_REQUIRE(e != nullptr, "event is NULL");
    ZAFB_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(s), static_cast<cudaEvent_t>(e), 0));
    return ZAFB_OK;
}
int zafb_event_elapsed_ms(void* a, void* b, float* ms) {
    ZAFB_REQUIRE(ms != nullptr, "ms is NULL");
    ZAFB_CUDA(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(a), static_cast<cudaEvent_t>(b)));
    return ZAFB_OK;
}
int64_t zafb_launch_count(void) { return g_launches.load(); }
int zafb_host_copy_bytes(int64_t* h2d, int64_t* d2h) {
    if (h2d) *h2d = g_h2d_bytes.load();
    if (d2h) *d2h = g_d2h_bytes.load();
    return ZAFB_OK;
}

// ------------------------------------------------------------------ integer bookkeeping
// The reference computes these with Python floats (true division + ceil/floor); IEEE double
// division followed by ceil()/floor() is the same operation, so the results are bit-identical.
int zafb_stft_geometry(int64_t ns, int64_t n, int64_t hop, int64_t* pad, int64_t* nt, int64_t* tail) {
    ZAFB_REQUIRE(ns >= 0 && n >= 1 && hop >= 1, "stft geometry: need ns >= 0, window_length >= 1, step_length >= 1");
    const int64_t p = n / 2;                                                       // zaf.py:99
    const int64_t t =
        static_cast<int64_t>(std::ceil(static_cast<double>((ns + 2 * p) - n) / static_cast<double>(hop))) + 1;  // :102-109
    if (pad) *pad = p;
    if (nt) *nt = t;
    if (tail) *tail = (t * hop + (n - hop) - p) - ns;                              // :116-121
    return ZAFB_OK;
}

int zafb_istft_geometry(int64_t n, int64_t nt, int64_t hop, int64_t* ola, int64_t* trim, int64_t* out_len) {
    ZAFB_REQUIRE(n >= 1 && nt >= 0 && hop >= 1, "istft geometry: need window_length >= 1, nt >= 0, step_length >= 1");
    const int64_t total = nt * hop + (n - hop);                                    // zaf.py:217
    const int64_t tr = n - hop;                                                    // :236-238
    // Python slice semantics of [tr : total - tr] (negative stop counts from the end, empty if reversed)
    int64_t stop = total - tr;
    int64_t start = tr;
    if (start < 0) start = start + total < 0 ? 0 : start + total;
    if (stop < 0) stop = stop + total < 0 ? 0 : stop + total;
    if (start > total) start = total;
    if (stop > total) stop = total;
    if (ola) *ola = total;
    if (trim) *trim = start;  // == N - hop whenever hop <= N
    if (out_len) *out_len = stop > start ? stop - start : 0;
    return ZAFB_OK;
}

int zafb_mdct_geometry(int64_t ns, int64_t n, int64_t* half, int64_t* nt, int64_t* tail) {
    ZAFB_REQUIRE(ns >= 0 && n >= 2, "mdct geometry: need ns >= 0 and window_length >= 2");
    const int64_t m = n / 2;                                                       // zaf.py:1029-1030
    const int64_t t = static_cast<int64_t>(std::ceil(static_cast<double>(ns) / static_cast<double>(m))) + 1;  // :1033
    if (half) *half = m;
    if (nt) *nt = t;
    if (tail) *tail = (t + 1) * m - ns;                                            // :1038
    return ZAFB_OK;
}

int zafb_imdct_geometry(int64_t m, int64_t nt, int64_t* ola, int64_t* out_len) {
    ZAFB_REQUIRE(m >= 1 && nt >= 0, "imdct geometry: need number_frequencies >= 1, nt >= 0");
    const int64_t total = m * (nt + 1);                                            // zaf.py:1132
    const int64_t len = total - 2 * m - 1;                                         // [M : -M-1], :1182
    if (ola) *ola = total;
    if (out_len) *out_len = len > 0 ? len : 0;
    return ZAFB_OK;
}

int zafb_cqt_geometry(int64_t ns, int64_t step, int64_t fft_length, int64_t* nt, int64_t* front, int64_t* back) {
    ZAFB_REQUIRE(ns >= 0 && step >= 1 && fft_length >= 1, "cqt geometry: need ns >= 0, step >= 1, fft_length >= 1");
    ZAFB_REQUIRE(fft_length >= step, "cqt geometry: step_length %lld exceeds fft_length %lld (np.pad would reject the negative pad, zaf.py:612)",
                 (long long)step, (long long)fft_length);
    if (nt) *nt = static_cast<int64_t>(std::floor(static_cast<double>(ns) / static_cast<double>(step)));   // zaf.py:606
    if (front) *front = static_cast<int64_t>(std::ceil(static_cast<double>(fft_length - step) / 2.0));     // :615
    if (back) *back = static_cast<int64_t>(std::floor(static_cast<double>(fft_length - step) / 2.0));      // :616
    return ZAFB_OK;
}

}  // extern "C"
