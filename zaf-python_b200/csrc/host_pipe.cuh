// Host-buffer pipeline shared by every "_host_f32" entry point: the batch is cut into chunks of
// rows (clips), and chunk i's H2D copy, kernels and D2H copy run on stream i mod 3, so the two
// copy engines and the SMs overlap.  Streams and device staging buffers are created once per
// process and kept (a cudaMalloc / cudaFree pair per call costs more than a small transform).
#pragma once

#include <chrono>
#include <mutex>

#include "common.cuh"

namespace zafb {

struct HostPipe {
    static constexpr int kStages = 3;
    std::mutex mu;
    cudaStream_t st[kStages] = {};
    void* d_in[kStages] = {};
    void* d_out[kStages] = {};
    size_t in_cap = 0, out_cap = 0;
    int device = -1;

    int ensure(size_t in_bytes, size_t out_bytes);
    void release();
};

HostPipe& host_pipe();
size_t host_pipe_chunk_bytes();  // device bytes per pipeline stage (ZAFB_PIPE_CHUNK_MB, default 64)

// Rows of `in_width` bytes at pitch `in_host_pitch` on the host are staged to rows of pitch
// `in_dev_pitch` on the device; launch(d_in, d_out, first_row, n_rows, stream) enqueues the kernels
// of one chunk; rows of `out_width` bytes at device pitch `out_dev_pitch` are copied back to host
// rows of pitch `out_host_pitch`.  Returns when every chunk has completed.
template <class Launch>
int run_host_pipeline(const void* in_host, size_t in_host_pitch, size_t in_width, size_t in_dev_pitch, void* out_host,
                      size_t out_host_pitch, size_t out_width, size_t out_dev_pitch, int64_t n_rows, Launch&& launch) {
    if (n_rows <= 0) return ZAFB_OK;
    HostPipe& hp = host_pipe();
    std::lock_guard<std::mutex> lock(hp.mu);
    const bool trace = env_flag("ZAFB_TRACE", 0) != 0;
    const auto t0 = std::chrono::steady_clock::now();
    const size_t row_max = in_dev_pitch > out_dev_pitch ? in_dev_pitch : out_dev_pitch;
    int64_t per = int64_t(host_pipe_chunk_bytes() / (row_max ? row_max : 1));
    if (per < 1) per = 1;
    if (size_t(n_rows) * row_max <= (size_t(4) << 20)) {
        per = n_rows;  // a small batch: one chunk (splitting it would only add launches)
    } else if (per * HostPipe::kStages > n_rows) {
        per = (n_rows + HostPipe::kStages - 1) / HostPipe::kStages;  // at least kStages chunks in flight
    }
    if (per < 1) per = 1;
    int rc = hp.ensure(size_t(per) * in_dev_pitch, size_t(per) * out_dev_pitch);
    if (rc != ZAFB_OK) return rc;
    const auto t1 = std::chrono::steady_clock::now();
    int s = 0;
    for (int64_t r0 = 0; r0 < n_rows; r0 += per, s = (s + 1) % HostPipe::kStages) {
        const int64_t nr = (r0 + per <= n_rows) ? per : n_rows - r0;
        if (in_width > 0) {
            const char* src = static_cast<const char*>(in_host) + size_t(r0) * in_host_pitch;
            // a failed enqueue must not return before the streams are drained: earlier chunks' D2H copies may still be
            // writing into the caller's result buffer
            const cudaError_t e =
                (in_host_pitch == in_width && in_dev_pitch == in_width)
                    ? cudaMemcpyAsync(hp.d_in[s], src, size_t(nr) * in_width, cudaMemcpyHostToDevice, hp.st[s])
                    : cudaMemcpy2DAsync(hp.d_in[s], in_dev_pitch, src, in_host_pitch, in_width, size_t(nr),
                                        cudaMemcpyHostToDevice, hp.st[s]);
            if (e != cudaSuccess) {
                rc = fail(ZAFB_E_CUDA, "host pipeline: H2D copy failed: %s", cudaGetErrorString(e));
                break;
            }
        }
        g_h2d_bytes.fetch_add(int64_t(nr) * int64_t(in_width), std::memory_order_relaxed);
        g_d2h_bytes.fetch_add(int64_t(nr) * int64_t(out_width), std::memory_order_relaxed);
        rc = launch(hp.d_in[s], hp.d_out[s], r0, nr, hp.st[s]);
        if (rc != ZAFB_OK) break;
        if (out_width > 0) {
            char* dst = static_cast<char*>(out_host) + size_t(r0) * out_host_pitch;
            const cudaError_t e =
                (out_host_pitch == out_width && out_dev_pitch == out_width)
                    ? cudaMemcpyAsync(dst, hp.d_out[s], size_t(nr) * out_width, cudaMemcpyDeviceToHost, hp.st[s])
                    : cudaMemcpy2DAsync(dst, out_host_pitch, hp.d_out[s], out_dev_pitch, out_width, size_t(nr),
                                        cudaMemcpyDeviceToHost, hp.st[s]);
            if (e != cudaSuccess) {
                rc = fail(ZAFB_E_CUDA, "host pipeline: D2H copy failed: %s", cudaGetErrorString(e));
                break;
            }
        }
    }
    const auto t2 = std::chrono::steady_clock::now();
    for (int i = 0; i < HostPipe::kStages; ++i) {
        cudaError_t e = cudaStreamSynchronize(hp.st[i]);
        if (e != cudaSuccess && rc == ZAFB_OK) rc = fail(ZAFB_E_CUDA, "host pipeline: %s", cudaGetErrorString(e));
    }
    if (trace) {
        const auto t3 = std::chrono::steady_clock::now();
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "[zafb pipe] rows %lld x (in %zu, out %zu) B, %lld rows/chunk: ensure %.2f ms, enqueue %.2f ms, drain %.2f ms\n",
                (long long)n_rows, in_width, out_width, (long long)per, ms(t0, t1), ms(t1, t2), ms(t2, t3));
    }
    return rc;
}

}  // namespace zafb
