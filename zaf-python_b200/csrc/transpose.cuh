// Layout conversion between the two (bins, frames) matrix layouts of the C ABI:
//   FRAME_MAJOR [clip][frame][bin]  <->  BIN_MAJOR [clip][bin][frame]  (the reference's C order, zaf.py:128).
// The warp-level transform kernels read and write frame-major memory (a frame is one contiguous run); the
// BIN_MAJOR contract is met by a tiled transpose pass through a stream-ordered scratch buffer, in chunks of clips.
#pragma once

#include <algorithm>

#include "common.cuh"

namespace zafb {

// out[b][c][r] = in[b][r][c]   (in: batch x rows x cols, out: batch x cols x rows), T = float or float2
int transpose_batched_f32(const float* in, float* out, int64_t batch, int64_t rows, int64_t cols, cudaStream_t st);
int transpose_batched_c32(const float2* in, float2* out, int64_t batch, int64_t rows, int64_t cols, cudaStream_t st);

inline int transpose_batched(const float* in, float* out, int64_t b, int64_t r, int64_t c, cudaStream_t st) {
    return transpose_batched_f32(in, out, b, r, c, st);
}
inline int transpose_batched(const float2* in, float2* out, int64_t b, int64_t r, int64_t c, cudaStream_t st) {
    return transpose_batched_c32(in, out, b, r, c, st);
}

size_t transpose_chunk_bytes();  // scratch per chunk (ZAFB_TRANSPOSE_CHUNK_MB, default 1024)
void keep_stream_pool();         // stream-ordered scratch stays in the pool between calls

// Produce a BIN_MAJOR result with a kernel that only writes FRAME_MAJOR: per chunk of clips,
// produce(clip0, n, scratch) fills scratch[n][frames][bins], then it is transposed into out[clip0 ..][bins][frames].
template <class T, class Produce>
int bin_major_from_frame_major(T* out, int64_t n_clips, int64_t frames, int64_t bins, cudaStream_t st, Produce&& produce) {
    if (n_clips * frames * bins == 0) return ZAFB_OK;
    keep_stream_pool();
    const size_t clip_bytes = size_t(frames) * size_t(bins) * sizeof(T);
    const int64_t per = std::max<int64_t>(1, std::min<int64_t>(n_clips, int64_t(transpose_chunk_bytes() / clip_bytes)));
    T* scratch = nullptr;
    ZAFB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&scratch), size_t(per) * clip_bytes, st));
    int rc = ZAFB_OK;
    for (int64_t c0 = 0; c0 < n_clips && rc == ZAFB_OK; c0 += per) {
        const int64_t n = std::min(per, n_clips - c0);
        rc = produce(c0, n, scratch);
        if (rc == ZAFB_OK) rc = transpose_batched(scratch, out + c0 * frames * bins, n, frames, bins, st);
    }
    cudaFreeAsync(scratch, st);
    return rc;
}

// Feed a kernel that only reads FRAME_MAJOR from a BIN_MAJOR input: per chunk, in[clip0 ..][bins][frames] is transposed
// into scratch[n][frames][bins], then consume(clip0, n, scratch) runs.
template <class T, class Consume>
int frame_major_from_bin_major(const T* in, int64_t n_clips, int64_t frames, int64_t bins, cudaStream_t st, Consume&& consume) {
    if (n_clips * frames * bins == 0) return ZAFB_OK;
    keep_stream_pool();
    const size_t clip_bytes = size_t(frames) * size_t(bins) * sizeof(T);
    const int64_t per = std::max<int64_t>(1, std::min<int64_t>(n_clips, int64_t(transpose_chunk_bytes() / clip_bytes)));
    T* scratch = nullptr;
    ZAFB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&scratch), size_t(per) * clip_bytes, st));
    int rc = ZAFB_OK;
    for (int64_t c0 = 0; c0 < n_clips && rc == ZAFB_OK; c0 += per) {
        const int64_t n = std::min(per, n_clips - c0);
        rc = transpose_batched(in + c0 * frames * bins, scratch, n, bins, frames, st);
        if (rc == ZAFB_OK) rc = consume(c0, n, scratch);
    }
    cudaFreeAsync(scratch, st);
    return rc;
}

}  // namespace zafb
