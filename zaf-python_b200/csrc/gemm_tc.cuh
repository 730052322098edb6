// fp32-accurate dense contraction on the tcgen05 tensor cores (3xTF32); see gemm_tc.cu.
#pragma once

#include <cstring>

#include "common.cuh"

namespace zafb {

// hi = tf32(x), lo = tf32(x - hi) of a rows x cols matrix (row pitch ldx) into two matrices of pitch
// ld_out >= cols (a multiple of 4; the padding columns are written as zeros).
int split_tf32(const float* x, int64_t rows, int64_t cols, int64_t ldx, float* hi, float* lo, int64_t ld_out, cudaStream_t st);
// the same split for a host-side float64 table (constant operators: rounded once from float64)
void split_tf32_host(const double* x, size_t n, float* hi, float* lo);

// C[M x N] (row pitch ldc) = A[M x K] . B[N x K]^T with A = a_hi + a_lo, B = b_hi + b_lo (K-major, pitches lda / ldb
// multiples of 4 elements, 16-byte aligned).  Asynchronous on `st`.
int gemm3xtf32(const float* a_hi, const float* a_lo, int64_t lda, const float* b_hi, const float* b_lo, int64_t ldb, float* c,
               int64_t ldc, int64_t M, int64_t N, int64_t K, cudaStream_t st);

// One output-column block of a structured product (see gemm3xtf32_tiled)
struct GemmTile {
    int a_col0;    // first A column this block multiplies
    int k_len;     // number of A / B columns
    int c_col0;    // first output column
    int c_stride;  // distance between its output columns
    int n_valid;   // output columns actually written (<= the block width)
    int fold;      // gemm3xtf32_pair_fold only: 1 = the block multiplies s[m] = x[m] + x[n-1-m], 2 = d[m] = x[m] - x[n-1-m]
};

// Structured operators in one launch: block t = B rows [t bn, t bn + bn) (bn = 16 or 128, all blocks share the row pitch
// ldb and store their k_len columns from column 0), multiplied with A columns [a_col0, a_col0 + k_len) and written to
// C columns c_col0 + i c_stride.  d_tiles is a DEVICE array of n_tiles entries.
int gemm3xtf32_tiled(int bn, const float* a_hi, const float* a_lo, int64_t lda, int64_t a_cols, const float* b_hi, const float* b_lo,
                     int64_t ldb, int n_tiles, const GemmTile* d_tiles, float* c, int64_t ldc, int64_t M, cudaStream_t st);

// The same structured product on CTA pairs (tcgen05 cta_group::2, 256 x 256 tiles, persistent): block t = B rows
// [256 t, 256 t + 256) of the b_rows x ldb operand (rows past b_rows read as zeros).  Half the operand bytes per output
// of the single-CTA kernel.
int gemm3xtf32_pair_tiled(const float* a_hi, const float* a_lo, int64_t lda, int64_t a_cols, const float* b_hi, const float* b_lo,
                          int64_t ldb, int64_t b_rows, int n_tiles, const GemmTile* d_tiles, float* c, int64_t ldc, int64_t M,
                          cudaStream_t st);

// The CTA-pair kernel with the even / odd fold of a transform with input symmetry fused into the operand path: A is never
// staged -- converter warps read the raw rows x (M x n, pitch ldx, n a multiple of 8, 16-byte aligned rows), form
// s / d per GemmTile::fold for m < n/2 and write the TF32 hi / lo operand tiles into shared memory themselves.
int gemm3xtf32_pair_fold(const float* x, int64_t ldx, int n, const float* b_hi, const float* b_lo, int64_t ldb, int64_t b_rows,
                         int n_tiles, const GemmTile* d_tiles, float* c, int64_t ldc, int64_t M, cudaStream_t st);

}  // namespace zafb
