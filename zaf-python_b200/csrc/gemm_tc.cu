// Dense contraction on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), fp32-accurate.
//
//     C[M x N] = A[M x K] . B[N x K]^T          (A, B K-major, C row-major)
//
// used where the hot path really is a dense matrix product:
//   * dct / dst of types whose natural FFT length is not a power of two (types I at N = 2^k are
//     real FFTs of length 2(N-1) = 2.3.11.31 / 2(N+1) = 2.5^2.41 at N = 1024, zaf.py:770-771,
//     907-910): batch x closed-form orthonormal matrix (hard part 8b);
//   * the mel filterbank / CQT-kernel application in isolation.
//
// Precision: a single TF32 product (10-bit mantissa) misses the 1e-5 parity bar by 20x
// (SURVEY.md section 7.3), so every operand is split x = hi + lo with hi = tf32(x),
// lo = tf32(x - hi) and the kernel accumulates  A_hi B_hi + A_hi B_lo + A_lo B_hi  in the fp32
// TMEM accumulator ("3xTF32"; the dropped lo.lo term is 2^-22 relative).  The tensor core adds
// into its accumulator with truncation, an error that grows linearly with the length of the
// accumulation chain (measured: 1.0e-5 of max|C| at K = 1000 in one chain), so K is cut into chunks
// of 128: each chunk is accumulated in one of two TMEM buffers and the epilogue warps add the
// chunks in registers with round-to-nearest fp32 while the next chunk is being multiplied.
//
// Kernel anatomy (one 128 x BN output tile per CTA, 192 threads):
//   warp 4   TMA producer: per 32-column K block, four cp.async.bulk.tensor loads (A_hi, A_lo,
//            B_hi, B_lo tiles, 128-byte swizzle) into a 3-stage shared-memory ring, mbarrier
//            complete_tx signalling;
//   warp 5   allocates TMEM (2 x BN columns), then ONE thread issues
//            tcgen05.mma.cta_group::1.kind::tf32 (M = 128, N = BN, K = 8): 4 K-steps x 3 products
//            per stage; tcgen05.commit frees the stage and, every 4 stages, publishes the chunk;
//   warps 0-3 epilogue: per chunk tcgen05.ld (32 lanes x 32 columns per warp) -> running sums in
//            registers -> release the TMEM buffer; after the last chunk, registers -> global.
#include <cuda.h>

#include <cstdint>
#include <mutex>
#include <vector>

#include "gemm_tc.cuh"

namespace zafb {
namespace {

constexpr int kBM = 128;       // UMMA M
constexpr int kBK = 32;        // fp32 elements per K block = one 128-byte swizzle row
constexpr int kUmmaK = 8;      // tf32: 32 bytes per instruction
constexpr int kStages = 3;
constexpr int kThreads = 192;
constexpr int kChunkKb = 4;    // K blocks per accumulation chain (128 elements of K)

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major operand tile, 128-byte swizzle, rows of 128 bytes, 8-row groups 1024 bytes apart
// (cute::UMMA::SmemDescriptor: start >> 4 | LBO 1 << 16 | SBO 64 << 32 | version 1 << 46 | SWIZZLE_128B 2 << 61)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    return uint64_t((smem_addr & 0x3FFFF) >> 4) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) |
           (uint64_t(2) << 61);
}

template <int BN>
struct Smem {
    static constexpr int kATile = kBM * kBK * 4;  // 16 KB
    static constexpr int kBTile = BN * kBK * 4;
    static constexpr int kStage = 2 * kATile + 2 * kBTile;
    static constexpr int kBars = kStages * kStage;          // full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]
    static constexpr int kTmemPtr = kBars + 8 * (2 * kStages + 4);
    static constexpr int kTotal = kTmemPtr + 16 + 1024;     // + slack to align the base to 1024 bytes
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm3xtf32_kernel(const __grid_constant__ CUtensorMap tm_ahi, const __grid_constant__ CUtensorMap tm_alo,
                  const __grid_constant__ CUtensorMap tm_bhi, const __grid_constant__ CUtensorMap tm_blo,
                  float* __restrict__ C, int64_t M, int64_t N, int64_t K, int64_t ldc, const GemmTile* __restrict__ tiles) {
    // tiles (structured operators; nullptr = dense): output-column block n_blk multiplies A columns [a_col0, a_col0 + k_len)
    // with columns [0, k_len) of B rows [n_blk BN, n_blk BN + BN) and writes its n_valid columns to C columns
    // c_col0 + i * c_stride.
    using S = Smem<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar_full = base + S::kBars, bar_empty = bar_full + 8 * kStages;
    const uint32_t bar_tfull = bar_empty + 8 * kStages, bar_tempty = bar_tfull + 16;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + S::kTmemPtr);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_blk = blockIdx.x, m_blk = blockIdx.y;
    GemmTile tile;
    tile.a_col0 = 0;
    tile.k_len = int(K);
    tile.c_col0 = n_blk * BN;
    tile.c_stride = 1;
    tile.n_valid = int(N - int64_t(n_blk) * BN < BN ? N - int64_t(n_blk) * BN : BN);
    if (tiles != nullptr) tile = tiles[n_blk];
    const int a_col0 = tile.a_col0;
    const int num_kb = (tile.k_len + kBK - 1) / kBK;
    const int num_chunks = (num_kb + kChunkKb - 1) / kChunkKb;
    constexpr uint32_t kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;  // two accumulator buffers (a power of two >= 32)

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_tfull + 8 * b, 1);
            mbar_init(bar_tempty + 8 * b, 4);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), kTmemCols);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {  // ===== TMA producer
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages, it = kb / kStages;
                mbar_wait(bar_empty + 8 * s, (it & 1) ^ 1);
                const uint32_t st = base + s * S::kStage;
                mbar_expect_tx(bar_full + 8 * s, S::kStage);
                tma_load_2d(st, &tm_ahi, bar_full + 8 * s, a_col0 + kb * kBK, m_blk * kBM);
                tma_load_2d(st + S::kATile, &tm_alo, bar_full + 8 * s, a_col0 + kb * kBK, m_blk * kBM);
                tma_load_2d(st + 2 * S::kATile, &tm_bhi, bar_full + 8 * s, kb * kBK, n_blk * BN);
                tma_load_2d(st + 2 * S::kATile + S::kBTile, &tm_blo, bar_full + 8 * s, kb * kBK, n_blk * BN);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {  // ===== MMA issuer
            // instruction descriptor: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), K-major both, N >> 3 at 17, M >> 4 at 24
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(BN >> 3) << 17) | (uint32_t(kBM >> 4) << 24);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages, it = kb / kStages;
                const int chunk = kb / kChunkKb, buf = chunk & 1;
                if (kb % kChunkKb == 0) {  // a new accumulation chain: its TMEM buffer must have been drained
                    mbar_wait(bar_tempty + 8 * buf, ((chunk >> 1) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                mbar_wait(bar_full + 8 * s, it & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = base + s * S::kStage;
                const uint32_t acc = tmem_acc + uint32_t(buf * BN);
                const uint64_t a_hi = umma_desc_k_sw128(st), a_lo = umma_desc_k_sw128(st + S::kATile);
                const uint64_t b_hi = umma_desc_k_sw128(st + 2 * S::kATile), b_lo = umma_desc_k_sw128(st + 2 * S::kATile + S::kBTile);
#pragma unroll
                for (int k = 0; k < kBK / kUmmaK; ++k) {
                    const uint64_t adv = uint64_t((k * kUmmaK * 4) >> 4);  // +32 bytes inside the swizzle row
                    umma_tf32(acc, a_lo + adv, b_hi + adv, idesc, ((kb % kChunkKb) | k) != 0);
                    umma_tf32(acc, a_hi + adv, b_lo + adv, idesc, 1);
                    umma_tf32(acc, a_hi + adv, b_hi + adv, idesc, 1);
                }
                umma_commit(bar_empty + 8 * s);  // the stage is free once these MMAs have read it
                if (kb % kChunkKb == kChunkKb - 1 || kb == num_kb - 1) umma_commit(bar_tfull + 8 * buf);  // chunk complete
            }
        }
    } else {  // ===== epilogue: warp w owns TMEM lanes [32 w, 32 w + 32) = rows of the tile
        float sum[BN];
#pragma unroll
        for (int i = 0; i < BN; ++i) sum[i] = 0.f;
        for (int chunk = 0; chunk < num_chunks; ++chunk) {
            const int buf = chunk & 1;
            mbar_wait(bar_tfull + 8 * buf, (chunk >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if constexpr (BN == 16) {  // the packed banded contraction: 16 operator rows per tile
                float v[16];
                tmem_ld16(tmem_acc + (uint32_t(32 * warp) << 16) + uint32_t(buf * BN), v);
#pragma unroll
                for (int i = 0; i < 16; ++i) sum[i] += v[i];
            } else {
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    float v[32];
                    tmem_ld32(tmem_acc + (uint32_t(32 * warp) << 16) + uint32_t(buf * BN + c0), v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) sum[c0 + i] += v[i];
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_tempty + 8 * buf) : "memory");
        }
        const int64_t row = int64_t(m_blk) * kBM + 32 * warp + lane;
        if (row < M) {
            float* crow = C + row * ldc + tile.c_col0;
            const int ncol = tile.n_valid;
            if (tile.c_stride == 1 && (ldc % 4 == 0) && (tile.c_col0 % 4 == 0) && (reinterpret_cast<uintptr_t>(C) % 16 == 0) && ncol >= BN) {
#pragma unroll
                for (int i = 0; i < BN; i += 4)
                    *reinterpret_cast<float4*>(crow + i) = make_float4(sum[i], sum[i + 1], sum[i + 2], sum[i + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < BN; ++i)
                    if (i < ncol) crow[i * tile.c_stride] = sum[i];
            }
        }
    }
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_dealloc(tmem_acc, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// CTA-pair form (tcgen05 cta_group::2): two CTAs of one cluster (two SMs of one TPC) compute a 256 x 256 output tile.
// CTA r of the pair stages rows [128 r, 128 r + 128) of the A tile AND rows [128 r, 128 r + 128) of the B tile; the
// tensor cores of both SMs read both halves of B, so each SM loads 64 KB per 32-column K block for 128 x 256 outputs --
// half the operand bytes per output of the single-CTA 128 x 128 kernel, whose time is the L2 -> SM operand stream
// (profiles/r02_dct1_ncu_summary.txt: tensor pipe 32 % active, 148 SMs x 1 MB per tile through L2).  The kernel is
// persistent: a pair walks over the work items (M block, column tile) pair, pair + G, ..., so the TMA producer runs
// ahead across tiles and the epilogue of one tile (registers -> shared-memory transpose -> coalesced
// stores) overlaps the first accumulation chains of the next.
//   warps 0-7  epilogue (warp w: TMEM lanes 32 (w & 3) .. + 32, columns 128 (w >> 2) .. + 128 of each chain buffer)
//   warp 8     TMA producer (both CTAs; complete_tx on the LEADER's full barrier)
//   warp 9     TMEM allocation (both CTAs), MMA issue (leader CTA only), commits multicast to both CTAs
constexpr int kPairBN = 256;
constexpr int kPairThreads = 320;
constexpr int kPairThreadsFold = 384;
constexpr int kPairTile = kBM * kBK * 4;               // 16 KB: A_hi, A_lo, B_hi, B_lo tiles of one CTA
constexpr int kPairStage = 4 * kPairTile;              // 64 KB
constexpr int kPairBars = kStages * kPairStage;        // full[3], empty[3], tfull[2], tempty[2], raw[3]
constexpr int kPairTmemPtr = kPairBars + 8 * (3 * kStages + 4);
constexpr int kPairXpose = kPairTmemPtr + 16;          // 8 warps x 32 x 33 floats
constexpr int kPairSmem = kPairXpose + 8 * 32 * 33 * 4 + 1024;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    // default semantics (release at CTA scope), as in CUTLASS' ClusterBarrier::arrive(cta_id): a cluster-scope release
    // compiles to MEMBAR.ALL.GPU + ERRBAR on every arrival (measured: the hottest instructions of the converter warps)
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is signalled on a barrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {  // arrives on `bar` of BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(uint16_t(3)) : "memory");
}

// FOLD = true: the A operand is never staged in HBM.  TMA loads the raw fp32 blocks x[row][32 kb ..) and the mirrored
// x[row][n - 32 (kb + 1) ..) into the stage's two A slots; two converter warps per CTA (warps 10, 11) form
// s = x[m] + x[n-1-m] or d = x[m] - x[n-1-m] (GemmTile::fold = 1 | 2), split it into TF32 hi / lo halves and overwrite
// the slots in place (generic-proxy stores + fence.proxy.async + an arrival on the leader's full barrier).
template <bool FOLD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FOLD ? kPairThreadsFold : kPairThreads, 1)
gemm3xtf32_pair_kernel(const __grid_constant__ CUtensorMap tm_ahi, const __grid_constant__ CUtensorMap tm_alo,
                       const __grid_constant__ CUtensorMap tm_bhi, const __grid_constant__ CUtensorMap tm_blo,
                       float* __restrict__ C, int64_t M, int64_t ldc, int m_pairs, int n_tiles,
                       const GemmTile* __restrict__ tiles, const float* __restrict__ x, int64_t ldx, int n_fold) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar_full = base + kPairBars, bar_empty = bar_full + 8 * kStages;
    const uint32_t bar_tfull = bar_empty + 8 * kStages, bar_tempty = bar_tfull + 16, bar_raw = bar_tempty + 16;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + kPairTmemPtr);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    // work item w = (M block, column tile), column tile fastest: pairs that run side by side share an M block, so its A
    // columns are fetched from HBM once and the interleaved even / odd output columns of a row meet in L2
    const int64_t work = int64_t(m_pairs) * n_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full + 8 * s, FOLD ? 5 : 1);  // the leader's expect_tx arrival (bytes of both CTAs complete on it)
                                                        // + FOLD: two converter warps of each CTA
            mbar_init(bar_raw + 8 * s, 1);              // FOLD: this CTA's raw x tiles have landed
            mbar_init(bar_empty + 8 * s, 1);  // one multicast commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_tfull + 8 * b, 1);
            mbar_init(bar_tempty + 8 * b, 16);  // 8 epilogue warps of each CTA arrive on the leader's barrier
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(const_cast<uint32_t*>(tmem_slot))), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();  // the peer's barriers are initialised before anything signals them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 8) {
        if (lane == 0) {  // ===== TMA producer (both CTAs)
            const uint32_t full_leader = mapa_rank(bar_full, 0);
            uint32_t g = 0;
            for (int64_t w = pair; w < work; w += num_pairs) {
                const int mp = int(w / n_tiles), t = int(w - int64_t(mp) * n_tiles);
                const int a_row = (2 * mp + int(rank)) * kBM;
                {
                    const GemmTile tile = tiles[t];
                    const int num_kb = (tile.k_len + kBK - 1) / kBK;
                    const int b_row = t * kPairBN + int(rank) * kBM;
                    for (int kb = 0; kb < num_kb; ++kb, ++g) {
                        const uint32_t s = g % kStages, it = g / kStages;
                        mbar_wait(bar_empty + 8 * s, (it & 1) ^ 1);
                        const uint32_t st = base + s * kPairStage;
                        if (rank == 0) mbar_expect_tx(bar_full + 8 * s, FOLD ? 4 * kPairTile : 2 * kPairStage);
                        const uint32_t fb = full_leader + 8 * s;
                        if constexpr (!FOLD) {
                            tma_load_2d_pair(st, &tm_ahi, fb, tile.a_col0 + kb * kBK, a_row);
                            tma_load_2d_pair(st + kPairTile, &tm_alo, fb, tile.a_col0 + kb * kBK, a_row);
                        } else {  // raw rows (tm_ahi = the map of x): forward block, mirrored block
                            mbar_expect_tx(bar_raw + 8 * s, 2 * kPairTile);
                            tma_load_2d(st, &tm_ahi, bar_raw + 8 * s, kb * kBK, a_row);
                            tma_load_2d(st + kPairTile, &tm_ahi, bar_raw + 8 * s, n_fold - kBK * (kb + 1), a_row);
                        }
                        tma_load_2d_pair(st + 2 * kPairTile, &tm_bhi, fb, kb * kBK, b_row);
                        tma_load_2d_pair(st + 3 * kPairTile, &tm_blo, fb, kb * kBK, b_row);
                    }
                }
            }
            // tail: every multicast commit aimed at this CTA's empty barriers has landed before the CTA may exit
            for (int i = 0; i < kStages; ++i, ++g) mbar_wait(bar_empty + 8 * (g % kStages), ((g / kStages) & 1) ^ 1);
        }
    } else if (warp == 9) {
        if (lane == 0 && rank == 0) {  // ===== MMA issuer (leader CTA)
            // D = F32, A = B = TF32, K-major both, N = 256, M = 256 (128 rows in each CTA's TMEM)
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(kPairBN >> 3) << 17) | (uint32_t(256 >> 4) << 24);
            uint32_t g = 0, gc = 0;
            for (int64_t w = pair; w < work; w += num_pairs) {
                {
                    const int num_kb = (tiles[w % n_tiles].k_len + kBK - 1) / kBK;
                    for (int kb = 0; kb < num_kb; ++kb, ++g) {
                        const uint32_t s = g % kStages, it = g / kStages;
                        const uint32_t buf = gc & 1;
                        if (kb % kChunkKb == 0) {
                            mbar_wait(bar_tempty + 8 * buf, ((gc >> 1) & 1) ^ 1);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        }
                        mbar_wait(bar_full + 8 * s, it & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t st = base + s * kPairStage;
                        const uint32_t acc = tmem_acc + buf * kPairBN;
                        const uint64_t a_hi = umma_desc_k_sw128(st), a_lo = umma_desc_k_sw128(st + kPairTile);
                        const uint64_t b_hi = umma_desc_k_sw128(st + 2 * kPairTile), b_lo = umma_desc_k_sw128(st + 3 * kPairTile);
#pragma unroll
                        for (int k = 0; k < kBK / kUmmaK; ++k) {
                            const uint64_t adv = uint64_t((k * kUmmaK * 4) >> 4);
                            umma_tf32_pair(acc, a_lo + adv, b_hi + adv, idesc, ((kb % kChunkKb) | k) != 0);
                            umma_tf32_pair(acc, a_hi + adv, b_lo + adv, idesc, 1);
                            umma_tf32_pair(acc, a_hi + adv, b_hi + adv, idesc, 1);
                        }
                        umma_commit_pair(bar_empty + 8 * s);
                        if (kb % kChunkKb == kChunkKb - 1 || kb == num_kb - 1) {
                            umma_commit_pair(bar_tfull + 8 * buf);
                            ++gc;
                        }
                    }
                }
            }
            // tail: the last arrivals of both CTAs' epilogue warps have landed on this CTA's barriers
            for (int i = 0; i < 2; ++i, ++gc) mbar_wait(bar_tempty + 8 * (gc & 1), ((gc >> 1) & 1) ^ 1);
        }
    } else if (FOLD && warp >= 10) {  // ===== converter warps (both CTAs): raw tiles -> folded, split A tiles, in place
        // The producer's TMA put x[row][32 kb .. + 32) into the A_hi slot and the mirrored block x[row][n - 32 (kb + 1) .. + 32)
        // into the A_lo slot.  Fold index m = 32 kb + i pairs with element 31 - i of the mirrored block, i.e. 16-byte chunk j
        // with chunk 7 - j reversed; one thread owns chunks j and 7 - j of a row in both slots, reads all four, then
        // overwrites them with hi / lo of s = x[m] + x[n-1-m] (fold 1) or d = x[m] - x[n-1-m] (fold 2).
        const int cw = warp - 10, rl = lane & 7, jp = lane >> 3;
        const uint32_t full_leader = mapa_rank(bar_full, 0);
        // round to nearest, ties away, on the 13 dropped mantissa bits: what cvt.rna.tf32.f32 returns for finite values
        // (and split_tf32_host computes), in two integer instructions instead of the five the cvt expands to
        auto tf32_split = [](float v, float& hv, float& lv) {
            hv = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
            lv = __uint_as_float((__float_as_uint(v - hv) + 0x1000u) & 0xFFFFE000u);
        };
        auto fold4 = [&](const float4 f, const float4 mrev, float sg, float4& hv, float4& lv) {
            tf32_split(fmaf(sg, mrev.w, f.x), hv.x, lv.x);
            tf32_split(fmaf(sg, mrev.z, f.y), hv.y, lv.y);
            tf32_split(fmaf(sg, mrev.y, f.z), hv.z, lv.z);
            tf32_split(fmaf(sg, mrev.x, f.w), hv.w, lv.w);
        };
        uint32_t g = 0;
        for (int64_t w = pair; w < work; w += num_pairs) {
            const GemmTile tile = tiles[w % n_tiles];
            const int num_kb = (tile.k_len + kBK - 1) / kBK;
            const float sg = tile.fold == 2 ? -1.f : 1.f;
            for (int kb = 0; kb < num_kb; ++kb, ++g) {
                const uint32_t s = g % kStages;
                mbar_wait(bar_raw + 8 * s, (g / kStages) & 1);
                uint8_t* hi_slot = base_ptr + s * kPairStage;
                uint8_t* lo_slot = hi_slot + kPairTile;
#pragma unroll 4
                for (int it = 0; it < 8; ++it) {
                    const int row = 64 * cw + 8 * it + rl;
                    const int oa = row * 128 + ((jp ^ (row & 7)) << 4), ob = row * 128 + (((7 - jp) ^ (row & 7)) << 4);
                    const float4 f0 = *reinterpret_cast<const float4*>(hi_slot + oa);
                    const float4 f1 = *reinterpret_cast<const float4*>(hi_slot + ob);
                    const float4 m0 = *reinterpret_cast<const float4*>(lo_slot + oa);
                    const float4 m1 = *reinterpret_cast<const float4*>(lo_slot + ob);
                    float4 h0, l0, h1, l1;
                    fold4(f0, m1, sg, h0, l0);
                    fold4(f1, m0, sg, h1, l1);
                    *reinterpret_cast<float4*>(hi_slot + oa) = h0;
                    *reinterpret_cast<float4*>(hi_slot + ob) = h1;
                    *reinterpret_cast<float4*>(lo_slot + oa) = l0;
                    *reinterpret_cast<float4*>(lo_slot + ob) = l1;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(full_leader + 8 * s);
            }
        }
    } else {  // ===== epilogue (both CTAs): TMEM lanes = the CTA's own 128 rows
        const int q = warp & 3, ch = warp >> 2;
        const uint32_t tempty_leader = mapa_rank(bar_tempty, 0);
        float* xp = reinterpret_cast<float*>(base_ptr + kPairXpose) + warp * (32 * 33);
        uint32_t gc = 0;
        for (int64_t w = pair; w < work; w += num_pairs) {
            const int mp = int(w / n_tiles), t = int(w - int64_t(mp) * n_tiles);
            const int64_t row0 = int64_t(2 * mp + int(rank)) * kBM + 32 * q;
            {
                const GemmTile tile = tiles[t];
                const int num_kb = (tile.k_len + kBK - 1) / kBK;
                const int num_chunks = (num_kb + kChunkKb - 1) / kChunkKb;
                float sum[128];
#pragma unroll
                for (int i = 0; i < 128; ++i) sum[i] = 0.f;
                for (int chunk = 0; chunk < num_chunks; ++chunk, ++gc) {
                    const uint32_t buf = gc & 1;
                    mbar_wait(bar_tfull + 8 * buf, (gc >> 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                    for (int c0 = 0; c0 < 128; c0 += 32) {
                        float v[32];
                        tmem_ld32(tmem_acc + (uint32_t(32 * q) << 16) + buf * kPairBN + uint32_t(128 * ch + c0), v);
#pragma unroll
                        for (int i = 0; i < 32; ++i) sum[c0 + i] += v[i];
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(tempty_leader + 8 * buf);
                }
                // registers (lane = row) -> 32 x 33 tile -> lane = column: every store instruction writes one row segment
                const int col_base = 128 * ch;
#pragma unroll
                for (int c0 = 0; c0 < 128; c0 += 32) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) xp[lane * 33 + i] = sum[c0 + i];
                    __syncwarp();
                    const int col = col_base + c0 + lane;
                    if (col < tile.n_valid) {
                        float* cp = C + row0 * ldc + tile.c_col0 + int64_t(col) * tile.c_stride;
#pragma unroll 8
                        for (int j = 0; j < 32; ++j)
                            if (row0 + j < M) cp[int64_t(j) * ldc] = xp[j * 33 + lane];
                    }
                    __syncwarp();
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();  // both CTAs are done with the pair's TMEM and with each other's barriers
    if (warp == 9) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(512u) : "memory");
    }
}

// hi = tf32(x) (round to nearest), lo = tf32(x - hi)
__global__ void split_tf32_kernel(const float* __restrict__ x, int64_t rows, int64_t cols, int64_t ldx, float* __restrict__ hi,
                                  float* __restrict__ lo, int64_t ld_out) {
    // 32 x 8 threads: eight rows per block step, a warp walks along one row (no index division, coalesced)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int64_t r = int64_t(blockIdx.x) * 8 + ty; r < rows; r += int64_t(gridDim.x) * 8) {
        const float* xr = x + r * ldx;
        for (int64_t c = tx; c < ld_out; c += 32) {
            const float v = c < cols ? xr[c] : 0.f;
            uint32_t h, l;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
            const float hv = __uint_as_float(h);
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(v - hv));
            hi[r * ld_out + c] = hv;
            lo[r * ld_out + c] = __uint_as_float(l);
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// rows x cols fp32 matrix, row pitch ld elements, box = 32 columns x box_rows rows, 128-byte swizzle, zero fill
int make_map(CUtensorMap* map, const float* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(ZAFB_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t gdim[2] = {cuuint64_t(cols), cuuint64_t(rows)};
    const cuuint64_t gstride[1] = {cuuint64_t(ld) * sizeof(float)};
    const cuuint32_t box[2] = {cuuint32_t(kBK), cuuint32_t(box_rows)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ZAFB_E_CUDA, "cuTensorMapEncodeTiled failed (%d) for a %lld x %lld matrix, pitch %lld",
                                       int(r), (long long)rows, (long long)cols, (long long)ld);
    return ZAFB_OK;
}

template <int BN>
int launch(const float* a_hi, const float* a_lo, int64_t lda, const float* b_hi, const float* b_lo, int64_t ldb, float* c,
           int64_t ldc, int64_t M, int64_t N, int64_t K, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        ZAFB_CUDA(cudaFuncSetAttribute(gemm3xtf32_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BN>::kTotal));
        attr = true;
    }
    CUtensorMap ma, mal, mb, mbl;
    int rc = make_map(&ma, a_hi, M, K, lda, kBM);
    if (rc == ZAFB_OK) rc = make_map(&mal, a_lo, M, K, lda, kBM);
    if (rc == ZAFB_OK) rc = make_map(&mb, b_hi, N, K, ldb, BN);
    if (rc == ZAFB_OK) rc = make_map(&mbl, b_lo, N, K, ldb, BN);
    if (rc != ZAFB_OK) return rc;
    const dim3 grid(unsigned((N + BN - 1) / BN), unsigned((M + kBM - 1) / kBM));
    gemm3xtf32_kernel<BN><<<grid, kThreads, Smem<BN>::kTotal, st>>>(ma, mal, mb, mbl, c, M, N, K, ldc, nullptr);
    ZAFB_LAUNCH_CHECK();
    return ZAFB_OK;
}

}  // namespace

// Structured operators in ONE launch: output-column block t (BN columns of B rows [t BN, t BN + BN)) has its own A column
// range, K length and output columns (GemmTile, a DEVICE array of n_tiles entries).  Used for
//   * the PACKED banded contraction (BN = 16: 16-row groups of a banded operator, each over the union of its bands only);
//   * the even / odd split of a transform matrix with input symmetry (BN = 128: DCT / DST types I and II, half the work).
template <int BN>
static int launch_tiled(const float* a_hi, const float* a_lo, int64_t lda, int64_t a_cols, const float* b_hi, const float* b_lo,
                        int64_t ldb, int n_tiles, const GemmTile* d_tiles, float* c, int64_t ldc, int64_t M, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        ZAFB_CUDA(cudaFuncSetAttribute(gemm3xtf32_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BN>::kTotal));
        attr = true;
    }
    CUtensorMap ma, mal, mb, mbl;
    int rc = make_map(&ma, a_hi, M, a_cols, lda, kBM);
    if (rc == ZAFB_OK) rc = make_map(&mal, a_lo, M, a_cols, lda, kBM);
    if (rc == ZAFB_OK) rc = make_map(&mb, b_hi, int64_t(n_tiles) * BN, ldb, ldb, BN);
    if (rc == ZAFB_OK) rc = make_map(&mbl, b_lo, int64_t(n_tiles) * BN, ldb, ldb, BN);
    if (rc != ZAFB_OK) return rc;
    const dim3 grid(unsigned(n_tiles), unsigned((M + kBM - 1) / kBM));
    gemm3xtf32_kernel<BN><<<grid, kThreads, Smem<BN>::kTotal, st>>>(ma, mal, mb, mbl, c, M, int64_t(n_tiles) * BN, a_cols, ldc, d_tiles);
    ZAFB_LAUNCH_CHECK();
    return ZAFB_OK;
}

int gemm3xtf32_tiled(int bn, const float* a_hi, const float* a_lo, int64_t lda, int64_t a_cols, const float* b_hi, const float* b_lo,
                     int64_t ldb, int n_tiles, const GemmTile* d_tiles, float* c, int64_t ldc, int64_t M, cudaStream_t st) {
    ZAFB_REQUIRE(M >= 0 && n_tiles >= 1 && (bn == 16 || bn == 128), "tiled gemm: bad shape");
    if (M == 0) return ZAFB_OK;
    ZAFB_REQUIRE(lda % 4 == 0 && ldb % 4 == 0, "tiled gemm: operand row pitches must be multiples of 4 elements");
    if (bn == 16) return launch_tiled<16>(a_hi, a_lo, lda, a_cols, b_hi, b_lo, ldb, n_tiles, d_tiles, c, ldc, M, st);
    return launch_tiled<128>(a_hi, a_lo, lda, a_cols, b_hi, b_lo, ldb, n_tiles, d_tiles, c, ldc, M, st);
}

// Number of CTA pairs that are resident at once (the kernels are persistent: a pair that had to wait for another one to
// finish would double the run time).  cudaOccupancyMaxActiveClusters knows the GPC layout; half the SM count is the fallback.
template <class K>
int resident_pairs(K kernel, int threads, size_t smem) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(sm_count()) & ~1u);
    cfg.blockDim = dim3(unsigned(threads));
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&clusters, kernel, &cfg) != cudaSuccess || clusters < 1) {
        cudaGetLastError();
        clusters = sm_count() / 2;
    }
    return clusters;
}

// CTA-pair kernel over 256-row blocks of B (block t = B rows [256 t, 256 t + 256), its own A column range / output columns)
int gemm3xtf32_pair_tiled(const float* a_hi, const float* a_lo, int64_t lda, int64_t a_cols, const float* b_hi, const float* b_lo,
                          int64_t ldb, int64_t b_rows, int n_tiles, const GemmTile* d_tiles, float* c, int64_t ldc, int64_t M,
                          cudaStream_t st) {
    ZAFB_REQUIRE(M >= 0 && n_tiles >= 1 && d_tiles != nullptr && b_rows >= 1, "pair gemm: bad shape");
    if (M == 0) return ZAFB_OK;
    ZAFB_REQUIRE(lda % 4 == 0 && ldb % 4 == 0, "pair gemm: operand row pitches must be multiples of 4 elements");
    ZAFB_REQUIRE(M < (int64_t(1) << 31), "pair gemm: dimension too large");
    static bool attr = false;
    if (!attr) {
        ZAFB_CUDA(cudaFuncSetAttribute(gemm3xtf32_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem));
        attr = true;
    }
    CUtensorMap ma, mal, mb, mbl;
    int rc = make_map(&ma, a_hi, M, a_cols, lda, kBM);
    if (rc == ZAFB_OK) rc = make_map(&mal, a_lo, M, a_cols, lda, kBM);
    if (rc == ZAFB_OK) rc = make_map(&mb, b_hi, b_rows, ldb, ldb, kBM);  // rows past b_rows read as zeros
    if (rc == ZAFB_OK) rc = make_map(&mbl, b_lo, b_rows, ldb, ldb, kBM);
    if (rc != ZAFB_OK) return rc;
    const int64_t m_pairs = (M + 2 * kBM - 1) / (2 * kBM);
    static const int resident = resident_pairs(gemm3xtf32_pair_kernel<false>, kPairThreads, kPairSmem);
    int64_t pairs = resident;
    if (pairs > m_pairs * n_tiles) pairs = m_pairs * n_tiles;
    gemm3xtf32_pair_kernel<false><<<unsigned(2 * pairs), kPairThreads, kPairSmem, st>>>(ma, mal, mb, mbl, c, M, ldc, int(m_pairs),
                                                                                        n_tiles, d_tiles, nullptr, 0, 0);
    ZAFB_LAUNCH_CHECK();
    return ZAFB_OK;
}

int gemm3xtf32_pair_fold(const float* x, int64_t ldx, int n, const float* b_hi, const float* b_lo, int64_t ldb, int64_t b_rows,
                         int n_tiles, const GemmTile* d_tiles, float* c, int64_t ldc, int64_t M, cudaStream_t st) {
    ZAFB_REQUIRE(M >= 0 && n_tiles >= 1 && d_tiles != nullptr && b_rows >= 1, "pair gemm: bad shape");
    if (M == 0) return ZAFB_OK;
    ZAFB_REQUIRE(n % 8 == 0 && ldx % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 && ldb % 4 == 0,
                 "fused fold: n must be a multiple of 8 and the rows 16-byte aligned");
    ZAFB_REQUIRE(M < (int64_t(1) << 31), "pair gemm: dimension too large");
    static bool attr = false;
    if (!attr) {
        ZAFB_CUDA(cudaFuncSetAttribute(gemm3xtf32_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem));
        attr = true;
    }
    CUtensorMap mx, mb, mbl;
    int rc = make_map(&mx, x, M, n, ldx, kBM);
    if (rc == ZAFB_OK) rc = make_map(&mb, b_hi, b_rows, ldb, ldb, kBM);
    if (rc == ZAFB_OK) rc = make_map(&mbl, b_lo, b_rows, ldb, ldb, kBM);
    if (rc != ZAFB_OK) return rc;
    const int64_t m_pairs = (M + 2 * kBM - 1) / (2 * kBM);
    static const int resident = resident_pairs(gemm3xtf32_pair_kernel<true>, kPairThreadsFold, kPairSmem);
    int64_t pairs = resident;
    if (pairs > m_pairs * n_tiles) pairs = m_pairs * n_tiles;
    gemm3xtf32_pair_kernel<true><<<unsigned(2 * pairs), kPairThreadsFold, kPairSmem, st>>>(mx, mx, mb, mbl, c, M, ldc, int(m_pairs),
                                                                                           n_tiles, d_tiles, x, ldx, n);
    ZAFB_LAUNCH_CHECK();
    return ZAFB_OK;
}

int split_tf32(const float* x, int64_t rows, int64_t cols, int64_t ldx, float* hi, float* lo, int64_t ld_out, cudaStream_t st) {
    if (rows * ld_out == 0) return ZAFB_OK;
    int64_t blocks = (rows + 7) / 8;
    const int64_t cap = int64_t(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    split_tf32_kernel<<<unsigned(blocks), 256, 0, st>>>(x, rows, cols, ldx, hi, lo, ld_out);
    ZAFB_LAUNCH_CHECK();
    return ZAFB_OK;
}

void split_tf32_host(const double* x, size_t n, float* hi, float* lo) {
    auto rna = [](float v) {
        uint32_t u;
        memcpy(&u, &v, 4);
        u = (u + 0x1000u) & 0xFFFFE000u;  // round half away on the 13 dropped bits, like cvt.rna.tf32.f32
        float r;
        memcpy(&r, &u, 4);
        return r;
    };
    for (size_t i = 0; i < n; ++i) {
        const float h = rna(static_cast<float>(x[i]));
        hi[i] = h;
        lo[i] = rna(static_cast<float>(x[i] - double(h)));
    }
}

int gemm3xtf32(const float* a_hi, const float* a_lo, int64_t lda, const float* b_hi, const float* b_lo, int64_t ldb, float* c,
               int64_t ldc, int64_t M, int64_t N, int64_t K, cudaStream_t st) {
    ZAFB_REQUIRE(M >= 0 && N >= 0 && K >= 1, "gemm: bad shape");
    if (M == 0 || N == 0) return ZAFB_OK;
    ZAFB_REQUIRE(lda % 4 == 0 && ldb % 4 == 0, "gemm: operand row pitches must be multiples of 4 elements (TMA: 16-byte strides)");
    ZAFB_REQUIRE(reinterpret_cast<uintptr_t>(a_hi) % 16 == 0 && reinterpret_cast<uintptr_t>(a_lo) % 16 == 0 &&
                     reinterpret_cast<uintptr_t>(b_hi) % 16 == 0 && reinterpret_cast<uintptr_t>(b_lo) % 16 == 0,
                 "gemm: operands must be 16-byte aligned");
    ZAFB_REQUIRE(M < (int64_t(1) << 31) && N < (int64_t(1) << 31) && K < (int64_t(1) << 31), "gemm: dimension too large");
    if (N > 64) return launch<128>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c, ldc, M, N, K, st);
    if (N > 16) return launch<64>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c, ldc, M, N, K, st);
    return launch<16>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c, ldc, M, N, K, st);  // one 16-row group of a packed banded operator
}

}  // namespace zafb
