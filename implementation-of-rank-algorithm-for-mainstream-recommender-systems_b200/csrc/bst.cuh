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

constexpr int kBstLd = 20;    // padded row stride of the row-major shared arrays

// out[n] += sum_i in[i] * M[i*16 + n]   (M k-major -> W.in ; M as registered -> W^T.in)
__device__ __forceinline__ void matvec16(const float* __restrict__ M, const float (&in)[16], float (&out)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
#pragma unroll
        for (int n4 = 0; n4 < 4; ++n4) {
            const float4 w = *reinterpret_cast<const float4*>(M + i * 16 + n4 * 4);
            out[4 * n4 + 0] = fmaf(in[i], w.x, out[4 * n4 + 0]);
            out[4 * n4 + 1] = fmaf(in[i], w.y, out[4 * n4 + 1]);
            out[4 * n4 + 2] = fmaf(in[i], w.z, out[4 * n4 + 2]);
            out[4 * n4 + 3] = fmaf(in[i], w.w, out[4 * n4 + 3]);
        }
    }
}
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&v)[16]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float4 t = *reinterpret_cast<const float4*>(p + 4 * c);
        v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
    }
}
__device__ __forceinline__ void store_row(float* __restrict__ p, const float (&v)[16]) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
        *reinterpret_cast<float4*>(p + 4 * c) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}
__device__ __forceinline__ void set_vec(float (&v)[16], const float* __restrict__ src) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = src[i];
}

// LayerNorm over 16 values: zh = (z - mean) * rstd, y = zh * g + b.
__device__ __forceinline__ float layer_norm16(const float (&z)[16], const float* g, const float* b,
                                              float (&zh)[16], float (&y)[16]) {
    float mean = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) mean += z[i];
    mean *= (1.0f / 16.0f);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float d = z[i] - mean; var = fmaf(d, d, var); }
    const float rstd = 1.0f / sqrtf(var * (1.0f / 16.0f) + kLnEps);
#pragma unroll
    for (int i = 0; i < 16; ++i) { zh[i] = (z[i] - mean) * rstd; y[i] = fmaf(zh[i], g[i], b[i]); }
    return rstd;
}
// dz from dy: gh = dy*g ; dz = rstd * (gh - mean(gh) - zh * mean(gh*zh))
__device__ __forceinline__ void layer_norm16_bwd(const float (&dy)[16], const float* g, const float (&zh)[16],
                                                 float rstd, float (&dz)[16]) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float gh = dy[i] * g[i]; a += gh; b = fmaf(gh, zh[i], b); }
    a *= (1.0f / 16.0f); b *= (1.0f / 16.0f);
#pragma unroll
    for (int i = 0; i < 16; ++i) dz[i] = rstd * (dy[i] * g[i] - a - zh[i] * b);
}

// The three dropout sites of the block (BST/bst.py:86 w_o output, :62 inside the FFN, :90 FFN output):
// 16 keep-bits each for this thread's row, regenerated identically by the backward.
struct BstDrop {
    uint32_t k[3];
    float scale;
    bool active;
};
__device__ __forceinline__ BstDrop bst_drop_masks(const BstParams& p, int64_t row) {
    BstDrop d;
    d.active = p.drop_thr != 0;
    d.scale = p.drop_scale;
    d.k[0] = d.k[1] = d.k[2] = 0xffffu;
    if (d.active) {
        const uint64_t seed = p.rng[0], offset = p.rng[1];
#pragma unroll
        for (int s = 0; s < 3; ++s) d.k[s] = dropout_keep16(seed, offset, (uint64_t)row, s, p.drop_thr);
    }
    return d;
}
__device__ __forceinline__ void bst_drop(const BstDrop& d, int site, float (&v)[16]) {
    if (!d.active) return;
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = ((d.k[site] >> i) & 1u) ? v[i] * d.scale : 0.f;
}

__device__ __forceinline__ void bst_load_x(const BstParams& p, int64_t b, int t, float (&x)[16], int32_t* err_flag) {
    const float* src;
    if (p.x_in) {
        src = p.x_in + (b * p.T + t) * 16;
    } else {
        const int64_t row = checked_row(__ldg(p.idx + b * p.T + t), p.table_rows, err_flag);
        src = p.table + row * 16;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + c);
        x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
    }
}

// One tile ahead: this thread's gather index of `tile` (0 for rows past its end, a dense x input, or no such
// tile) is loaded while the current tile is computed, and half-way through the tile the row it names, the
// sample's length and the upstream-gradient rows are pulled into L2 — the row gather at the top of a tile then
// pays one L2 latency instead of an index -> row chain to DRAM.
__device__ __forceinline__ int64_t bst_tile_index(const BstParams& p, int64_t tile, int s, int t, int tid) {
    if (p.x_in || tile >= p.n_tiles) return 0;
    const int64_t b0 = tile * p.S;
    const int ns = (int)((p.B - b0) < p.S ? (p.B - b0) : p.S);
    return tid < ns * p.T ? __ldg(p.idx + (b0 + s) * p.T + t) : 0;
}
__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
__device__ __forceinline__ void bst_load_x_at(const BstParams& p, int64_t b, int t, int64_t idx, float (&x)[16], int32_t* err_flag) {
    const float* src = p.x_in ? p.x_in + (b * p.T + t) * 16 : p.table + checked_row(idx, p.table_rows, err_flag) * 16;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + c);
        x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
    }
}
// rows of the next tile into L2; true when this thread owns a row there
__device__ __forceinline__ bool bst_prefetch_next(const BstParams& p, int64_t tile_n, int s, int t, int tid, int64_t idx_next,
                                                 int64_t* b_next) {
    if (tile_n >= p.n_tiles) return false;
    const int64_t b0 = tile_n * p.S;
    const int ns = (int)((p.B - b0) < p.S ? (p.B - b0) : p.S);
    if (tid >= ns * p.T) return false;
    const int64_t b = b0 + s;
    if (p.x_in) prefetch_l2(p.x_in + (b * p.T + t) * 16);
    else prefetch_l2(p.table + ((uint64_t)idx_next < (uint64_t)p.table_rows ? idx_next : 0) * 16);
    if (t == 0) prefetch_l2(p.seq_len + b);
    *b_next = b;
    return true;
}

// Scores, masked softmax and context of one query row against the L live keys of its sample.
// exp of the attention kernels: expf on the fp32 path (1e-5 parity), the hardware ex2-based __expf on the tensor-core
// path (2 instructions instead of ~15; ~1e-6 relative, far inside that path's 2e-2 bar)
template <bool FAST>
__device__ __forceinline__ float bst_exp(float x) { return FAST ? __expf(x) : expf(x); }

template <int H, bool FAST = false>
__device__ __forceinline__ void bst_attend(const float (&q)[16], const float* __restrict__ ks,
                                           const float* __restrict__ vs, int row0, int L, float (&ctx)[16],
                                           float (&m_out)[H], float (&l_out)[H]) {
    constexpr int DH = 16 / H;
    const float scale = 1.0f / sqrtf((float)DH);
    // two passes over the live keys, the H heads side by side (independent chains, one row load per key)
    float m[H], l[H];
#pragma unroll
    for (int h = 0; h < H; ++h) { m[h] = -INFINITY; l[h] = 0.f; }
#pragma unroll
    for (int i = 0; i < 16; ++i) ctx[i] = 0.f;
    for (int u = 0; u < L; ++u) {
        float kr[16];
        load_row(ks + (row0 + u) * kBstLd, kr);
#pragma unroll
        for (int h = 0; h < H; ++h) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < DH; ++j) s = fmaf(q[h * DH + j], kr[h * DH + j], s);
            m[h] = fmaxf(m[h], s * scale);
        }
    }
    for (int u = 0; u < L; ++u) {
        float kr[16], vr[16];
        load_row(ks + (row0 + u) * kBstLd, kr);
        load_row(vs + (row0 + u) * kBstLd, vr);
#pragma unroll
        for (int h = 0; h < H; ++h) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < DH; ++j) s = fmaf(q[h * DH + j], kr[h * DH + j], s);
            const float pr = bst_exp<FAST>(s * scale - m[h]);
            l[h] += pr;
#pragma unroll
            for (int j = 0; j < DH; ++j) ctx[h * DH + j] = fmaf(pr, vr[h * DH + j], ctx[h * DH + j]);
        }
    }
    // L == 0: every key masked -> 0/0 = NaN, exactly as softmax over all -inf in the reference
#pragma unroll
    for (int h = 0; h < H; ++h) {
        if (FAST) {                      // one reciprocal per head (inf * 0 = NaN keeps the L == 0 case)
            const float inv = 1.0f / l[h];
#pragma unroll
            for (int j = 0; j < DH; ++j) ctx[h * DH + j] *= inv;
        } else {
#pragma unroll
            for (int j = 0; j < DH; ++j) ctx[h * DH + j] = ctx[h * DH + j] / l[h];
        }
        m_out[h] = m[h];
        l_out[h] = l[h];
    }
}

__device__ __forceinline__ int bst_len(const BstParams& p, int64_t b) {
    const int64_t l = __ldg(p.seq_len + b);
    return l < 0 ? 0 : (l > p.T ? p.T : (int)l);
}

// Tensor-core variant (bst_tc.cu): same inputs, outputs and partial layout as the SIMT kernels.
int bst_tc_bwd_ctas(int64_t B, int T);
int bst_tc_fwd(const BstParams& p, int nhead, float* y_out, float* pool_out, int pool_ld, int32_t* err_flag,
               cudaStream_t s);

int bst_tc_bwd(const BstParams& p, int nhead, const float* g_y, const float* g_pool, int g_pool_ld, float* g_x,
               float* partials, int n_ctas, int32_t* err_flag, cudaStream_t s);

}  // namespace rk
