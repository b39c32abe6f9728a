// tile_gemm.cuh — CTA-level fp32 SIMT GEMM on shared-memory tiles, the building block of the
// small per-row MLPs on the hot path (DeepCrossing residual units, AFM attention net, DIN
// activation unit, BST projections/FFN).
//
//   C[m][n] = sum_k A[k][m] * B[k][n]        m < M_TILE, n < N, k < K
//
// Both operands are "k-major": A is stored [K][lda] with the M_TILE rows (samples, (b,t) pairs,
// field pairs ...) contiguous, B is stored [K][ldb] with the N outputs contiguous.  Each thread
// owns 4 consecutive m and TN consecutive n: per k it issues one 16-byte load of A (contiguous
// across the lanes of a warp -> conflict free) and TN/4 16-byte loads of B (same address in all
// lanes of a warp when M_TILE >= 128 -> broadcast) for 4*TN FMAs.  Activations therefore live
// feature-major in shared memory, and a layer's output tile, written as [n][m], is directly the
// next layer's A operand.
#pragma once
#include "common.cuh"

namespace rk {

// epi(m0, n0, acc) receives the finished 4 x TN micro-tile, acc[i][j] = C[m0+i][n0+j].
// m_used (multiple of 4, <= lda) = rows actually present in the tile; N a multiple of TN.
template <int TN, int THREADS, class Epi>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ As, int lda,
                                          const float* __restrict__ Bs, int ldb, int K, int N,
                                          int m_used, Epi epi) {
    static_assert(TN % 4 == 0, "tile shape");
    const int MT = m_used >> 2;
    const int n_tiles = MT * (N / TN);
    for (int tile = threadIdx.x; tile < n_tiles; tile += THREADS) {
        const int m0 = (tile % MT) * 4;
        const int n0 = (tile / MT) * TN;
        float acc[4][TN];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
        const float* a = As + m0;
        const float* b = Bs + n0;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(a + k * lda);
            float bv[TN];
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                const float4 t = *reinterpret_cast<const float4*>(b + k * ldb + j);
                bv[j] = t.x; bv[j + 1] = t.y; bv[j + 2] = t.z; bv[j + 3] = t.w;
            }
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                acc[0][j] = fmaf(av.x, bv[j], acc[0][j]);
                acc[1][j] = fmaf(av.y, bv[j], acc[1][j]);
                acc[2][j] = fmaf(av.z, bv[j], acc[2][j]);
                acc[3][j] = fmaf(av.w, bv[j], acc[3][j]);
            }
        }
        epi(m0, n0, acc);
    }
}

// Store a 4 x TN micro-tile feature-major: dst[(n0+j)*ld + m0 .. m0+3] = acc[0..3][j].
template <int TN>
__device__ __forceinline__ void store_tile_kmajor(float* __restrict__ dst, int ld, int m0, int n0,
                                                  const float (&acc)[4][TN]) {
#pragma unroll
    for (int j = 0; j < TN; ++j)
        *reinterpret_cast<float4*>(dst + (n0 + j) * ld + m0) =
            make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
}

// Cooperative copy of a [rows][cols] row-major global matrix into shared memory, optionally
// transposed and zero-padded to [rows_p][ld] / [cols_p][ld].
template <int THREADS>
__device__ __forceinline__ void load_matrix(float* __restrict__ dst, int ld, int dst_rows,
                                            const float* __restrict__ src, int rows, int cols,
                                            bool transpose) {
    if (!transpose) {  // dst[r][c] = src[r][c]
        for (int i = threadIdx.x; i < dst_rows * ld; i += THREADS) {
            const int r = i / ld, c = i - r * ld;
            dst[i] = (r < rows && c < cols) ? __ldg(src + r * cols + c) : 0.f;
        }
    } else {           // dst[c][r] = src[r][c]; dst has dst_rows >= cols rows of ld >= rows
        for (int i = threadIdx.x; i < dst_rows * ld; i += THREADS) {
            const int c = i / ld, r = i - c * ld;
            dst[i] = (r < rows && c < cols) ? __ldg(src + r * cols + c) : 0.f;
        }
    }
}

}  // namespace rk
