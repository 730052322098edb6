// Batched tiled transpose (see transpose.cuh): 64 x 64 tiles through padded shared memory.
#include "transpose.cuh"

namespace zafb {
namespace {

// TR x TC tile (TR rows of the input = the contiguous axis of the OUTPUT) per CTA of 256 threads, 16 loads in flight
// per thread before the barrier.  Reads are TC-element runs that neighbouring CTAs (x-fastest launch order) extend to
// whole input rows; writes are TR-element runs, so TR is the long side: 128 x 32 writes 1 KB (float2) runs.
// Tile pitch TC + 1 keeps the transposed reads conflict-free.
template <class T, int TR, int TC>
__global__ void __launch_bounds__(256, 4)
transpose_tile_kernel(const T* __restrict__ in, T* __restrict__ out, int rows, int cols) {
    static_assert(TR * TC == 4096 && 256 % TC == 0 && 256 % TR == 0, "256 threads x 16 elements");
    extern __shared__ unsigned char tile_raw[];
    T(*tile)[TC + 1] = reinterpret_cast<T(*)[TC + 1]>(tile_raw);
    const int c0 = blockIdx.x * TC, r0 = blockIdx.y * TR;
    const int lx = threadIdx.x % TC, ly = threadIdx.x / TC;   // load:  column lx, rows ly + (256 / TC) i
    const int sx = threadIdx.x % TR, sy = threadIdx.x / TR;   // store: row sx,    columns sy + (256 / TR) i
    // one matrix per blockIdx.z (no loop over the batch in here: the compiler would hoist all 32 addresses out of it,
    // 160 registers, one CTA per SM -- measured 3.5 TB/s); element offsets inside a matrix fit 32 bits
    const T* src = in + int64_t(blockIdx.z) * rows * cols;
    T* dst = out + int64_t(blockIdx.z) * rows * cols;
    T v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int r = r0 + ly + (256 / TC) * i, c = c0 + lx;
        if (r < rows && c < cols) v[i] = __ldcs(src + (r * cols + c));
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) tile[ly + (256 / TC) * i][lx] = v[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int c = c0 + sy + (256 / TR) * i, r = r0 + sx;
        if (r < rows && c < cols) __stcs(dst + (c * rows + r), tile[sx][sy + (256 / TR) * i]);
    }
}

template <class T, int TR, int TC>
int launch_tiles(const T* in, T* out, int64_t batch, int64_t rows, int64_t cols, cudaStream_t st) {
    const int64_t gx = ceil_div(cols, TC), gy = ceil_div(rows, TR);
    ZAFB_REQUIRE(gy <= 65535, "transpose: too many row tiles");
    const size_t smem = sizeof(T) * TR * (TC + 1);
    static bool attr = false;
    if (!attr) {  // ask for the large shared-memory carve-out so that four CTAs (4 x 34 KB) fit on an SM
        ZAFB_CUDA((cudaFuncSetAttribute(transpose_tile_kernel<T, TR, TC>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared)));
        attr = true;
    }
    for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
        const int64_t nb = std::min<int64_t>(65535, batch - b0);
        const dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(gy), static_cast<unsigned>(nb));
        transpose_tile_kernel<T, TR, TC><<<grid, 256, smem, st>>>(in + b0 * rows * cols, out + b0 * rows * cols, int(rows), int(cols));
        ZAFB_LAUNCH_CHECK();
    }
    return ZAFB_OK;
}

template <class T>
int launch(const T* in, T* out, int64_t batch, int64_t rows, int64_t cols, cudaStream_t st) {
    if (batch * rows * cols == 0) return ZAFB_OK;
    ZAFB_REQUIRE(rows * cols < (int64_t(1) << 31), "transpose: matrix too large (rows x cols must be below 2^31)");
    const int shape = env_flag("ZAFB_TRANSPOSE_TILE", 128);
    if (shape == 64) return launch_tiles<T, 64, 64>(in, out, batch, rows, cols, st);
    if (shape == 256) return launch_tiles<T, 256, 16>(in, out, batch, rows, cols, st);
    return launch_tiles<T, 128, 32>(in, out, batch, rows, cols, st);
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
