// bst.cuh — parameters shared by the fp32 SIMT BST block (bst.cu) and the tcgen05 one (bst_tc.cu).
#pragma once
#include "common.cuh"

namespace rk {

constexpr int kBstRows = 128;   // (sample, position) rows per CTA tile
constexpr int kBstD    = 16;
constexpr float kLnEps = 1e-5f;

struct BstParams {
    const float* w[6];      // wq wk wv wo w1 w2, each [16][16] as registered (out, in)
    const float* vec[10];   // bq bk bv bo ln1_g ln1_b b1 b2 ln2_g ln2_b
    const float* pos;       // [max_len][16]
    const float* table;  const int64_t* idx;  int64_t table_rows;   // x = table[idx]  (first block)
    const float* x_in;                                              // or x = x_in[B,T,16]
    const int64_t* seq_len;
    int64_t B, n_tiles;
    int32_t T, S, pool_mean;
    const unsigned long long* rng;      // device: [seed, offset] of the dropout masks (drop_thr > 0)
    uint32_t drop_thr;                  // round(p * 65536); 0 = no dropout
    float drop_scale;                   // 1 / (1 - p)
};
enum { VBQ = 0, VBK, VBV, VBO, VG1, VBE1, VB1, VB2, VG2, VBE2 };
enum { MQ = 0, MK, MV, MO, M1, M2 };

// Tensor-core variant (bst_tc.cu): same inputs, outputs and partial layout as the SIMT kernels.
int bst_tc_bwd_ctas(int64_t B, int T);
int bst_tc_fwd(const BstParams& p, int nhead, float* y_out, float* pool_out, int pool_ld, int32_t* err_flag,
               cudaStream_t s);
int bst_tc_bwd(const BstParams& p, int nhead, const float* g_y, const float* g_pool, int g_pool_ld, float* g_x,
               float* g_params, float* partials, int n_ctas, int32_t* err_flag, cudaStream_t s);

}  // namespace rk
