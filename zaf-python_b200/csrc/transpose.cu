// Batched tiled transpose (see transpose.cuh): 64 x 64 tiles through padded shared memory.
#include "transpose.cuh"

namespace zafb {
namespace {

// 64 x 64 tile per CTA of 256 threads: every thread has 16 loads in flight before the barrier (the 32 x 32 / 4-load
// version measured 3.4 TB/s, latency-bound: long-scoreboard stalls 23 cycles per issue), rows are 512-byte (float2) or
// 256-byte (float) coalesced on both sides, the tile pitch of 65 keeps the transposed reads conflict-free.
constexpr int kTile = 64;

template <class T>
__global__ void __launch_bounds__(256)
transpose_tile_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t batch, int rows, int cols) {
    extern __shared__ unsigned char tile_raw[];
    T(*tile)[kTile + 1] = reinterpret_cast<T(*)[kTile + 1]>(tile_raw);
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
    const int c0 = blockIdx.x * kTile, r0 = blockIdx.y * kTile;
    for (int64_t b = blockIdx.z; b < batch; b += gridDim.z) {
        const T* src = in + b * int64_t(rows) * cols;
        T* dst = out + b * int64_t(rows) * cols;
        T v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int r = r0 + ty + 4 * i, c = c0 + tx;
            if (r < rows && c < cols) v[i] = __ldcs(src + int64_t(r) * cols + c);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) tile[ty + 4 * i][tx] = v[i];
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int c = c0 + ty + 4 * i, r = r0 + tx;
            if (r < rows && c < cols) __stcs(dst + int64_t(c) * rows + r, tile[tx][ty + 4 * i]);
        }
        __syncthreads();
    }
}

template <class T>
int launch(const T* in, T* out, int64_t batch, int64_t rows, int64_t cols, cudaStream_t st) {
    if (batch * rows * cols == 0) return ZAFB_OK;
    ZAFB_REQUIRE(rows < (int64_t(1) << 30) && cols < (int64_t(1) << 30), "transpose: matrix too large");
    const int64_t gx = ceil_div(cols, kTile), gy = ceil_div(rows, kTile);
    ZAFB_REQUIRE(gy <= 65535, "transpose: too many row tiles");
    const dim3 grid(unsigned(gx), unsigned(gy), unsigned(std::min<int64_t>(batch, 65535)));
    const size_t smem = sizeof(T) * kTile * (kTile + 1);
    transpose_tile_kernel<T><<<grid, 256, smem, st>>>(in, out, batch, int(rows), int(cols));
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
