// Batched tiled transpose (see transpose.cuh): 32 x 32 tiles through padded shared memory, 256-byte (float2) or
// 128-byte (float) coalesced rows on both sides, one CTA of 32 x 8 threads per tile, grid-stride over the batch.
#include "transpose.cuh"

namespace zafb {
namespace {

template <class T>
__global__ void __launch_bounds__(256)
transpose_tile_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t batch, int rows, int cols) {
    __shared__ T tile[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int64_t b = blockIdx.z; b < batch; b += gridDim.z) {
        const T* src = in + b * int64_t(rows) * cols;
        T* dst = out + b * int64_t(rows) * cols;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
            const int r = r0 + ty + i, c = c0 + tx;
            if (r < rows && c < cols) tile[ty + i][tx] = __ldcs(src + int64_t(r) * cols + c);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
            const int c = c0 + ty + i, r = r0 + tx;
            if (r < rows && c < cols) __stcs(dst + int64_t(c) * rows + r, tile[tx][ty + i]);
        }
        __syncthreads();
    }
}

template <class T>
int launch(const T* in, T* out, int64_t batch, int64_t rows, int64_t cols, cudaStream_t st) {
    if (batch * rows * cols == 0) return ZAFB_OK;
    ZAFB_REQUIRE(rows < (int64_t(1) << 30) && cols < (int64_t(1) << 30), "transpose: matrix too large");
    const int64_t gx = ceil_div(cols, 32), gy = ceil_div(rows, 32);
    ZAFB_REQUIRE(gy <= 65535, "transpose: too many row tiles");
    const dim3 grid(unsigned(gx), unsigned(gy), unsigned(std::min<int64_t>(batch, 65535)));
    transpose_tile_kernel<T><<<grid, dim3(32, 8), 0, st>>>(in, out, batch, int(rows), int(cols));
    ZAFB_LAUNCH_CHECK();
    return ZAFB_OK;
}

}  // namespace

int transpose_batched_f32(const float* in, float* out, int64_t batch, int64_t rows, int64_t cols, cudaStream_t st) {
    return launch<float>(in, out, batch, rows, cols, st);
}
int transpose_batched_c32(const float2* in, float2* out, int64_t batch, int64_t rows, int64_t cols, cudaStream_t st) {
    return launch<float2>(in, out, batch, rows, cols, st);
}

size_t transpose_chunk_bytes() {
    int mb = env_flag("ZAFB_TRANSPOSE_CHUNK_MB", 1024);
    if (mb < 1) mb = 1;
    return size_t(mb) << 20;
}

void keep_stream_pool() {
    static bool done = false;
    if (done) return;
    int dev = 0;
    cudaMemPool_t pool;
    uint64_t keep = UINT64_MAX;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    done = true;
}

}  // namespace zafb
