// bst_tc.cu — BST transformer block (BSTTransformer.forward, BST/bst.py:66-91, with the gather :224 and
// the pooling :238-241) with its dense parts on tcgen05 tensor cores: the Q/K/V projections, the output
// projection and both FFN layers (north star: "tcgen05 tensor cores only for the dense parts (... BST
// QKV/FFN projections)").  rk_bst_block_t.precision = RK_BST_BF16_TENSOR selects it; parity bar 2e-2.
//
// One thread = one (sample, position) row = one TMEM lane; 128 rows per tile, persistent CTAs.
//   A tile   [128 rows x 128 B]: four 32-byte slots of 16 bf16 per row, 128-byte swizzle, K-major.  A row's
//            activation is written by its own thread as split bf16: hi tile + lo tile (x = hi + lo).
//   B tiles  the six 16x16 weights as registered ([out][in] = K-major), four per 2 KB tile (one slot each),
//            hi + lo, converted once per CTA.
//   products A.W^T = A_hi.W_hi + A_lo.W_hi + A_hi.W_lo: three tcgen05.mma M128 N16 K16 per projection,
//            fp32 accumulators in TMEM (q | k | v in columns 0..47, the later layers in 48..63), read back
//            with tcgen05.ld by the row's thread, which adds the bias and carries on in registers.
// Scores, the masked softmax and the context (d_head = 16/nhead: 20x20x4 per sample) stay on the FMA pipe,
// keys/values exchanged through shared memory exactly as in the fp32 kernel (bst.cu); LayerNorm, LeakyReLU,
// dropout and pooling are per-row register work.  A projection costs the thread ~60 instructions (split,
// swizzled stores, one TMEM load) instead of 256 FMAs + 64 shared-memory loads.
#include <string.h>
#include "bst.cuh"
#include "umma.cuh"
#include "prof.cuh"

namespace rk {
namespace tc {

constexpr int kBtThreads  = 128;
constexpr int kBtTmemCols = 64;      // q | k | v | one 16-column output slot

struct BtFwdSmem {
    uint8_t *a_hi, *a_lo;            // [128][128 B]
    uint8_t *wa_hi, *wa_lo;          // [16][128 B]: slots Wq | Wk | Wv | Wo
    uint8_t *wb_hi, *wb_lo;          // [16][128 B]: slots W1 | W2
    float *ks, *vs;                  // [128][kBstLd]; ks doubles as the block-output rows for the pooling
    float *pos, *vec;                // [T][16], [10][16]
    uint64_t* bar;
    uint32_t* tmem_slot;
    __device__ BtFwdSmem(uint8_t* base, int T) {
        uint8_t* p = base;
        a_hi = p;  p += 128 * 128;
        a_lo = p;  p += 128 * 128;
        wa_hi = p; p += 16 * 128;
        wa_lo = p; p += 16 * 128;
        wb_hi = p; p += 16 * 128;
        wb_lo = p; p += 16 * 128;
        ks = (float*)p;  p += sizeof(float) * kBstRows * kBstLd;
        vs = (float*)p;  p += sizeof(float) * kBstRows * kBstLd;
        pos = (float*)p; p += sizeof(float) * T * 16;
        vec = (float*)p; p += sizeof(float) * 160;
        bar = (uint64_t*)p;       p += 8;
        tmem_slot = (uint32_t*)p;
    }
    static size_t bytes(int T) {
        return 1024 /* alignment slack */ + 2 * 128 * 128 + 4 * 16 * 128 + 2 * sizeof(float) * kBstRows * kBstLd +
               sizeof(float) * ((size_t)T * 16 + 160) + 16;
    }
};

// slot j of a tile line = bytes [32 j, 32 j + 32) = chunks 2j, 2j + 1
__device__ __forceinline__ void store_slot_split(uint8_t* hi, uint8_t* lo, int row, int slot, const float (&v)[16]) {
    float h8[8];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int j = 0; j < 8; ++j) h8[j] = v[8 * half + j];
        store_chunk_split(hi, lo, row, 2 * slot + half, h8);
    }
}

// the six weights -> split bf16 B tiles (one thread per 8 consecutive inputs of one output row)
__device__ __forceinline__ void stage_weight_tiles(const BstParams& p, uint8_t* wa_hi, uint8_t* wa_lo, uint8_t* wb_hi,
                                                   uint8_t* wb_lo, int tid, int n_threads) {
    for (int item = tid; item < 6 * 32; item += n_threads) {
        const int m = item >> 5, n = (item & 31) >> 1, half = item & 1;
        float v[8];
        ld8(p.w[m] + n * 16 + half * 8, v);
        uint8_t* hi = m < 4 ? wa_hi : wb_hi;
        uint8_t* lo = m < 4 ? wa_lo : wb_lo;
        store_chunk_split(hi, lo, n, 2 * (m & 3) + half, v);
    }
}

// D[:, col .. col+16) = A[slot_a] . W[slot_w]^T with split-bf16 operands (three MMAs)
__device__ __forceinline__ void mma_proj(uint32_t d_tmem, const uint64_t (&a_desc)[2], int slot_a,
                                         const uint64_t (&w_desc)[2], int slot_w) {
#pragma unroll
    for (int term = 0; term < 3; ++term)     // hi.hi + lo.hi + hi.lo
        umma_bf16(d_tmem, a_desc[term == 1] + 2 * slot_a, w_desc[term == 2] + 2 * slot_w, umma_idesc(16), term > 0);
}

template <int H>
__global__ void __launch_bounds__(kBtThreads)
bst_fwd_tc_kernel(const __grid_constant__ BstParams p, float* __restrict__ y_out, float* __restrict__ pool_out,
                  int pool_ld, int32_t* err_flag) {
    extern __shared__ uint8_t smem_raw_bt[];
    uint8_t* base = smem_raw_bt + ((1024u - (smem_u32(smem_raw_bt) & 1023u)) & 1023u);   // swizzle atoms need 1024 B
    BtFwdSmem sm(base, p.T);
    const int tid = threadIdx.x, warp = tid >> 5, T = p.T;

    if (tid == 0) mbar_init(sm.bar, 1);
    if (warp == 0) tmem_alloc(sm.tmem_slot, kBtTmemCols);
    stage_weight_tiles(p, sm.wa_hi, sm.wa_lo, sm.wb_hi, sm.wb_lo, tid, kBtThreads);
    for (int i = tid; i < 160; i += kBtThreads) sm.vec[i] = __ldg(p.vec[i >> 4] + (i & 15));
    for (int i = tid; i < T * 16; i += kBtThreads) sm.pos[i] = __ldg(p.pos + i);
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *sm.tmem_slot;
    const uint32_t my_tmem = tmem + ((uint32_t)(warp * 32) << 16);
    const uint64_t a_desc[2]  = {umma_desc(smem_u32(sm.a_hi)), umma_desc(smem_u32(sm.a_lo))};
    const uint64_t wa_desc[2] = {umma_desc(smem_u32(sm.wa_hi)), umma_desc(smem_u32(sm.wa_lo))};
    const uint64_t wb_desc[2] = {umma_desc(smem_u32(sm.wb_hi)), umma_desc(smem_u32(sm.wb_lo))};
    uint32_t ph = 0;

    // one tensor-core round trip: the rows' operands are in the A tile -> issue -> wait for the accumulators
    auto round_trip = [&](auto issue) {
        fence_async_smem();
        fence_before();
        __syncthreads();
        if (tid == 0) {
            fence_after();
            issue();
            umma_commit(sm.bar);
        }
        mbar_wait(sm.bar, ph);
        ph ^= 1;
        fence_after();
    };

    const int s_row = tid / T, t_row = tid - s_row * T;          // this thread's (sample, position) in every tile it is live in
    int64_t idx_cur = bst_tile_index(p, blockIdx.x, s_row, t_row, tid);
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int64_t b0 = tile * p.S;
        const int ns = (int)((p.B - b0) < p.S ? (p.B - b0) : p.S);
        const int rows = ns * T;
        const bool on = tid < rows;
        const int s = on ? s_row : 0, t = on ? t_row : 0;
        const int64_t b = b0 + s;
        const int64_t idx_next = bst_tile_index(p, tile + gridDim.x, s_row, t_row, tid);   // in flight during this tile
        float x[16], qk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { x[i] = 0.f; qk[i] = 0.f; }
        int L = 0;
        if (on) {
            L = bst_len(p, b);
            bst_load_x_at(p, b, t, idx_cur, x, err_flag);
#pragma unroll
            for (int i = 0; i < 16; ++i) qk[i] = x[i] + sm.pos[t * 16 + i];
        }
        store_slot_split(sm.a_hi, sm.a_lo, tid, 0, qk);
        store_slot_split(sm.a_hi, sm.a_lo, tid, 1, x);
        round_trip([&]() {
            mma_proj(tmem + 0, a_desc, 0, wa_desc, MQ);       // q = qk Wq^T
            mma_proj(tmem + 16, a_desc, 0, wa_desc, MK);      // k = qk Wk^T
            mma_proj(tmem + 32, a_desc, 1, wa_desc, MV);      // v = x  Wv^T
        });
        float q[16];
        {
            float qkv[32], vv[16];
            tmem_ld32(my_tmem + 0, qkv);
            tmem_ld16(my_tmem + 32, vv);
            float k[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                q[i] = qkv[i] + sm.vec[VBQ * 16 + i];
                k[i] = qkv[16 + i] + sm.vec[VBK * 16 + i];
                vv[i] += sm.vec[VBV * 16 + i];
            }
            if (on) {
                store_row(sm.ks + tid * kBstLd, k);
                store_row(sm.vs + tid * kBstLd, vv);
            }
        }
        fence_before();
        __syncthreads();
        float ctx[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) ctx[i] = 0.f;
        if (on) {
            float mh[H], lh[H];
            bst_attend<H, true>(q, sm.ks, sm.vs, s * T, L, ctx, mh, lh);
        }
        {
            int64_t b_next;
            bst_prefetch_next(p, tile + gridDim.x, s_row, t_row, tid, idx_next, &b_next);
            idx_cur = idx_next;
        }
        const BstDrop drop = bst_drop_masks(p, b * T + t);
        store_slot_split(sm.a_hi, sm.a_lo, tid, 0, ctx);
        round_trip([&]() { mma_proj(tmem + 48, a_desc, 0, wa_desc, MO); });
        float z[16], zh[16], o1[16];
        tmem_ld16(my_tmem + 48, z);
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] += sm.vec[VBO * 16 + i];
        bst_drop(drop, 0, z);
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] += qk[i];
        layer_norm16(z, sm.vec + VG1 * 16, sm.vec + VBE1 * 16, zh, o1);
        store_slot_split(sm.a_hi, sm.a_lo, tid, 0, o1);
        round_trip([&]() { mma_proj(tmem + 48, a_desc, 0, wb_desc, 0); });      // ffn.0
        float hp[16];
        tmem_ld16(my_tmem + 48, hp);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            hp[i] += sm.vec[VB1 * 16 + i];
            hp[i] = hp[i] > 0.f ? hp[i] : 0.01f * hp[i];
        }
        bst_drop(drop, 1, hp);
        store_slot_split(sm.a_hi, sm.a_lo, tid, 0, hp);
        round_trip([&]() { mma_proj(tmem + 48, a_desc, 0, wb_desc, 1); });      // ffn.3
        float f[16], y[16];
        tmem_ld16(my_tmem + 48, f);
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] += sm.vec[VB2 * 16 + i];
        bst_drop(drop, 2, f);
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = o1[i] + f[i];
        layer_norm16(z, sm.vec + VG2 * 16, sm.vec + VBE2 * 16, zh, y);
        if (on) {
            if (y_out) store_row(y_out + (b * T + t) * 16, y);
            store_row(sm.ks + tid * kBstLd, y);          // keys are no longer needed: the rows for the pooling
        }
        fence_before();
        __syncthreads();
        if (pool_out) {
            for (int item = tid; item < ns * 16; item += kBtThreads) {
                const int ss = item >> 4, n = item & 15;
                float a = 0.f;
                for (int tt = 0; tt < T; ++tt) a += sm.ks[(ss * T + tt) * kBstLd + n];
                if (p.pool_mean) a /= (float)__ldg(p.seq_len + b0 + ss);
                pool_out[(b0 + ss) * pool_ld + n] = a;
            }
        }
        __syncthreads();
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, kBtTmemCols);
}

template <int H>
static int launch_fwd(const BstParams& p, float* y_out, float* pool_out, int pool_ld, int32_t* err_flag, cudaStream_t s) {
    const size_t smem = BtFwdSmem::bytes(p.T);
    RK_CUDA(cudaFuncSetAttribute(bst_fwd_tc_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = p.n_tiles;
    const int64_t cap = (int64_t)sm_count() * 3;
    if (grid > cap) grid = cap;
    bst_fwd_tc_kernel<H><<<(int)grid, kBtThreads, smem, s>>>(p, y_out, pool_out, pool_ld, err_flag);
    RK_LAUNCH_CHECK();
    return 0;
}


// ================================================================================================
// Backward on the tensor core.  Same tiles, same thread = row = TMEM lane mapping.  The forward is
// recomputed (four projection round trips as above), then
//   * the transposed projections (d act = W2^T d f, d o1 += W1^T d h, d ctx = Wo^T d wo,
//     d qk = Wq^T dq + Wk^T dk, d x = d qk + Wv^T dv) are the same M128 N16 K16 split-bf16 MMAs against
//     transposed weight tiles, the row's gradient written into a 32-byte slot of the G tile;
//   * every weight gradient dW[n][k] = sum_rows g[row][n] a[row][k] AND every bias / LayerNorm column sum
//     is ONE family of MMAs per stage: G^T . X with both operands read MN-major from the very tiles the
//     projections use K-major (lines = rows = the K of this product; umma_desc_mn), X carrying a column of
//     ones.  The accumulators stay in TMEM for the whole life of the CTA (three 64-column regions, one per
//     stage, because the slots change meaning) and are read out once at the end: the batch-wide reductions
//     that were 25 % of the fp32 kernel (six outer products + ten column sums through shared memory, a dozen
//     barriers) cost 16 asynchronous MMAs per stage, issued together with the stage's projection.
//     G is used as hi + lo (the column sums are then exact to 2^-17), X as hi only (2^-9 per term on the
//     weight gradients: the bf16 bar of this path).
// Attention backward, LayerNorm backward, dropout masks and the position-table gradient are the fp32
// kernel's register / shared-memory code.
// ================================================================================================
constexpr int kBtBwdTmemCols = 256;       // [0,48) q|k|v or transposed products, [48,64) one output, 3 x 64 weight-gradient regions
constexpr uint32_t kWgA = 64, kWgB = 128, kWgC = 192;

struct BtBwdSmem {
    uint8_t *x_hi, *g_hi, *lo;       // [128][128 B] each: activations (hi), gradients (hi), the lo parts of whichever is being multiplied
    uint8_t *wa_hi, *wa_lo, *wb_hi, *wb_lo;         // Wq|Wk|Wv|Wo,  W1|W2|W1^T|W2^T
    uint8_t *wat_hi, *wat_lo;                       // Wq^T|Wk^T|Wv^T|Wo^T
    float *qs, *ks, *vs, *dc;        // [128][kBstLd]
    float *mrow, *lrow, *delta;      // [128][H]
    float *pos, *vec, *pacc;         // pacc [T][16]: this CTA's position-table gradient, element e owned by thread e % 128
    uint64_t* bar;
    uint32_t* tmem_slot;
    __device__ BtBwdSmem(uint8_t* base, int T, int H) {
        uint8_t* p = base;
        x_hi = p; p += 128 * 128;
        g_hi = p; p += 128 * 128;
        lo = p;   p += 128 * 128;
        wa_hi = p;  p += 16 * 128;  wa_lo = p;  p += 16 * 128;
        wb_hi = p;  p += 16 * 128;  wb_lo = p;  p += 16 * 128;
        wat_hi = p; p += 16 * 128;  wat_lo = p; p += 16 * 128;
        qs = (float*)p; p += sizeof(float) * kBstRows * kBstLd;
        ks = (float*)p; p += sizeof(float) * kBstRows * kBstLd;
        vs = (float*)p; p += sizeof(float) * kBstRows * kBstLd;
        dc = (float*)p; p += sizeof(float) * kBstRows * kBstLd;
        mrow = (float*)p;  p += sizeof(float) * kBstRows * H;
        lrow = (float*)p;  p += sizeof(float) * kBstRows * H;
        delta = (float*)p; p += sizeof(float) * kBstRows * H;
        pos = (float*)p; p += sizeof(float) * T * 16;
        vec = (float*)p; p += sizeof(float) * 160;
        pacc = (float*)p; p += sizeof(float) * T * 16;
        bar = (uint64_t*)p;       p += 8;
        tmem_slot = (uint32_t*)p;
    }
    static size_t bytes(int T, int H) {
        return 1024 + 3 * 128 * 128 + 6 * 16 * 128 + sizeof(float) * (4 * kBstRows * kBstLd + 3 * kBstRows * H) +
               sizeof(float) * ((size_t)T * 32 + 160) + 16;
    }
};

// transposed weights: line k of slot m holds W_m[.][k]
__device__ __forceinline__ void stage_weight_tiles_t(const BstParams& p, uint8_t* wat_hi, uint8_t* wat_lo, uint8_t* wb_hi,
                                                     uint8_t* wb_lo, int tid, int n_threads) {
    for (int item = tid; item < 6 * 32; item += n_threads) {
        const int m = item >> 5, k = (item & 31) >> 1, half = item & 1;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(p.w[m] + (8 * half + j) * 16 + k);
        uint8_t* hi = m < 4 ? wat_hi : wb_hi;
        uint8_t* lo = m < 4 ? wat_lo : wb_lo;
        store_chunk_split(hi, lo, k, 2 * (m < 4 ? m : m - 2) + half, v);        // W1^T, W2^T: slots 2, 3 of the W1|W2 tile
    }
}

// D[:, col..col+16) (+)= A[slot_a] . W[slot_w]^T, split bf16; `first` = overwrite instead of accumulate
__device__ __forceinline__ void mma_proj_acc(uint32_t d_tmem, const uint64_t (&a_desc)[2], int slot_a,
                                             const uint64_t (&w_desc)[2], int slot_w, bool first) {
#pragma unroll
    for (int term = 0; term < 3; ++term)
        umma_bf16(d_tmem, a_desc[term == 1] + 2 * slot_a, w_desc[term == 2] + 2 * slot_w, umma_idesc(16),
                  (!first || term > 0) ? 1u : 0u);
}

// region (+)= (G_hi + G_lo)^T . X_hi over the 128 rows of the tile: both operands MN-major (lines = rows = K)
// (descriptors of line 0 of the three tiles, made once; 16 lines further = +2048 bytes = +128 in the address field)
__device__ __forceinline__ void mma_wgrad(uint32_t d_tmem, uint64_t g_hi_mn, uint64_t g_lo_mn, uint64_t x_hi_mn, bool first_tile) {
#pragma unroll
    for (int part = 0; part < 2; ++part)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
            umma_bf16(d_tmem, (part ? g_lo_mn : g_hi_mn) + 128u * kk, x_hi_mn + 128u * kk,
                      umma_idesc(64) | kUmmaAMn | kUmmaBMn, (first_tile && part == 0 && kk == 0) ? 0u : 1u);
}

template <int H>
__global__ void __launch_bounds__(kBtThreads)
bst_bwd_tc_kernel(const __grid_constant__ BstParams p, const float* __restrict__ g_y, const float* __restrict__ g_pool,
                  int g_pool_ld, float* __restrict__ g_x, float* __restrict__ partials, int32_t* err_flag) {
    constexpr int DH = 16 / H;
    extern __shared__ uint8_t smem_raw_bt[];
    uint8_t* base = smem_raw_bt + ((1024u - (smem_u32(smem_raw_bt) & 1023u)) & 1023u);
    BtBwdSmem sm(base, p.T, H);
    PROF_DECL
    const int tid = threadIdx.x, warp = tid >> 5, T = p.T;
    const float scale = 1.0f / sqrtf((float)DH);

    if (tid == 0) mbar_init(sm.bar, 1);
    if (warp == 0) tmem_alloc(sm.tmem_slot, kBtBwdTmemCols);
    stage_weight_tiles(p, sm.wa_hi, sm.wa_lo, sm.wb_hi, sm.wb_lo, tid, kBtThreads);
    stage_weight_tiles_t(p, sm.wat_hi, sm.wat_lo, sm.wb_hi, sm.wb_lo, tid, kBtThreads);
    for (int i = tid; i < 160; i += kBtThreads) sm.vec[i] = __ldg(p.vec[i >> 4] + (i & 15));
    for (int i = tid; i < T * 16; i += kBtThreads) { sm.pos[i] = __ldg(p.pos + i); sm.pacc[i] = 0.f; }
    {   // slot 3 of every X line: a one in its first element (the column of ones of the weight-gradient products)
        float one[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        store_chunk(sm.x_hi, tid, 6, one);
        store_chunk(sm.x_hi, tid, 7, zero);
        store_chunk(sm.g_hi, tid, 6, zero);        // G slot 3 is used by stage A only; zero until then
        store_chunk(sm.g_hi, tid, 7, zero);
        store_chunk(sm.lo, tid, 6, zero);
        store_chunk(sm.lo, tid, 7, zero);
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *sm.tmem_slot;
    const uint32_t my_tmem = tmem + ((uint32_t)(warp * 32) << 16);
    const uint64_t x_desc[2]   = {umma_desc(smem_u32(sm.x_hi)), umma_desc(smem_u32(sm.lo))};
    const uint64_t g_desc[2]   = {umma_desc(smem_u32(sm.g_hi)), umma_desc(smem_u32(sm.lo))};
    const uint64_t wa_desc[2]  = {umma_desc(smem_u32(sm.wa_hi)), umma_desc(smem_u32(sm.wa_lo))};
    const uint64_t wb_desc[2]  = {umma_desc(smem_u32(sm.wb_hi)), umma_desc(smem_u32(sm.wb_lo))};
    const uint64_t wat_desc[2] = {umma_desc(smem_u32(sm.wat_hi)), umma_desc(smem_u32(sm.wat_lo))};
    const uint64_t xh = umma_desc_mn(smem_u32(sm.x_hi), 128 * 128), gh = umma_desc_mn(smem_u32(sm.g_hi), 128 * 128),
                   lo_addr = umma_desc_mn(smem_u32(sm.lo), 128 * 128);
    uint32_t ph = 0;
    bool first_tile = true;

    auto round_trip = [&](auto issue) {
        fence_async_smem();
        fence_before();
        __syncthreads();
        if (tid == 0) {
            fence_after();
            issue();
            umma_commit(sm.bar);
        }
        mbar_wait(sm.bar, ph);
        ph ^= 1;
        fence_after();
    };
    float zero16[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) zero16[i] = 0.f;

    const int s_row = tid / T, t_row = tid - s_row * T;
    int64_t idx_cur = bst_tile_index(p, blockIdx.x, s_row, t_row, tid);
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int64_t b0 = tile * p.S;
        const int ns = (int)((p.B - b0) < p.S ? (p.B - b0) : p.S);
        const int rows = ns * T;
        const bool on = tid < rows;
        const int s = on ? s_row : 0, t = on ? t_row : 0;
        const int64_t b = b0 + s;
        const int row0 = s * T;
        const int64_t idx_next = bst_tile_index(p, tile + gridDim.x, s_row, t_row, tid);
        PROF(0); PROF_COUNT(12);
        // ---- A. recompute the forward
        float qk[16];
        int L = 0;
        {
            float x[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { x[i] = 0.f; qk[i] = 0.f; }
            if (on) {
                L = bst_len(p, b);
                bst_load_x_at(p, b, t, idx_cur, x, err_flag);
#pragma unroll
                for (int i = 0; i < 16; ++i) qk[i] = x[i] + sm.pos[t * 16 + i];
            }
            store_slot_split(sm.x_hi, sm.lo, tid, 0, qk);
            store_slot_split(sm.x_hi, sm.lo, tid, 1, x);
        }
        round_trip([&]() {
            mma_proj(tmem + 0, x_desc, 0, wa_desc, MQ);
            mma_proj(tmem + 16, x_desc, 0, wa_desc, MK);
            mma_proj(tmem + 32, x_desc, 1, wa_desc, MV);
        });
        {
            float qkv[32], vv[16], q[16], k[16];
            tmem_ld32(my_tmem + 0, qkv);
            tmem_ld16(my_tmem + 32, vv);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                q[i] = qkv[i] + sm.vec[VBQ * 16 + i];
                k[i] = qkv[16 + i] + sm.vec[VBK * 16 + i];
                vv[i] += sm.vec[VBV * 16 + i];
            }
            store_row(sm.qs + tid * kBstLd, on ? q : zero16);
            store_row(sm.ks + tid * kBstLd, on ? k : zero16);
            store_row(sm.vs + tid * kBstLd, on ? vv : zero16);
        }
        fence_before();
        __syncthreads();
        PROF(1);
        float ctx[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) ctx[i] = 0.f;
        if (on) {
            float q[16], mh[H], lh[H];
            load_row(sm.qs + tid * kBstLd, q);
            bst_attend<H, true>(q, sm.ks, sm.vs, row0, L, ctx, mh, lh);
#pragma unroll
            for (int h = 0; h < H; ++h) { sm.mrow[tid * H + h] = mh[h]; sm.lrow[tid * H + h] = 1.0f / lh[h]; }
        }
        PROF(2);
        {
            int64_t b_next;
            if (bst_prefetch_next(p, tile + gridDim.x, s_row, t_row, tid, idx_next, &b_next)) {
                if (g_y) prefetch_l2(g_y + (b_next * T + t_row) * 16);
                if (g_pool && t_row == 0) prefetch_l2(g_pool + b_next * g_pool_ld);
            }
            idx_cur = idx_next;
        }
        const BstDrop drop = bst_drop_masks(p, b * T + t);
        store_slot_split(sm.x_hi, sm.lo, tid, 2, ctx);                    // slot 2: ctx, kept for stage B
        round_trip([&]() { mma_proj(tmem + 48, x_desc, 2, wa_desc, MO); });
        float zh1[16], zh2[16], rstd1 = 0.f, rstd2 = 0.f;
        unsigned hmask = 0;
        {
            float z[16], o1[16], hp[16], f[16], y[16];
            tmem_ld16(my_tmem + 48, z);
#pragma unroll
            for (int i = 0; i < 16; ++i) z[i] += sm.vec[VBO * 16 + i];
            bst_drop(drop, 0, z);
#pragma unroll
            for (int i = 0; i < 16; ++i) z[i] += qk[i];
            rstd1 = layer_norm16(z, sm.vec + VG1 * 16, sm.vec + VBE1 * 16, zh1, o1);
            store_slot_split(sm.x_hi, sm.lo, tid, 0, on ? o1 : zero16);   // slot 0: o1 (qk comes back for stage C)
            round_trip([&]() { mma_proj(tmem + 48, x_desc, 0, wb_desc, 0); });
            tmem_ld16(my_tmem + 48, hp);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                hp[i] += sm.vec[VB1 * 16 + i];
                if (hp[i] > 0.f) hmask |= 1u << i;
                hp[i] = hp[i] > 0.f ? hp[i] : 0.01f * hp[i];
            }
            bst_drop(drop, 1, hp);
            store_slot_split(sm.x_hi, sm.lo, tid, 1, on ? hp : zero16);   // slot 1: the (dropped) FFN activation
            round_trip([&]() { mma_proj(tmem + 48, x_desc, 1, wb_desc, 1); });
            tmem_ld16(my_tmem + 48, f);
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] += sm.vec[VB2 * 16 + i];
            bst_drop(drop, 2, f);
#pragma unroll
            for (int i = 0; i < 16; ++i) z[i] = o1[i] + f[i];
            rstd2 = layer_norm16(z, sm.vec + VG2 * 16, sm.vec + VBE2 * 16, zh2, y);
        }
        PROF(3);
        // ---- B. upstream gradient, LayerNorm 2 backward, FFN backward
        float dz[16];
        {
            float dy[16], t0[16], df[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) dy[i] = 0.f;
            if (on) {
                if (g_y) load_row(g_y + (b * T + t) * 16, dy);
                if (g_pool) {
                    const float inv = p.pool_mean ? 1.0f / (float)__ldg(p.seq_len + b) : 1.0f;
#pragma unroll
                    for (int i = 0; i < 16; ++i) dy[i] = fmaf(g_pool[b * g_pool_ld + i], inv, dy[i]);
                }
                layer_norm16_bwd(dy, sm.vec + VG2 * 16, zh2, rstd2, dz);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) dz[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) { t0[i] = on ? dy[i] * zh2[i] : 0.f; df[i] = dz[i]; }
            bst_drop(drop, 2, df);
            store_slot_split(sm.g_hi, sm.lo, tid, 0, df);                 // G: df | dh | dy.zh2 | dy
            store_slot_split(sm.g_hi, sm.lo, tid, 2, t0);
            store_slot_split(sm.g_hi, sm.lo, tid, 3, dy);
        }
        round_trip([&]() { mma_proj(tmem + 48, g_desc, 0, wb_desc, 3); });             // d act = W2^T df
        float do1[16];
        {
            float dh[16];
            tmem_ld16(my_tmem + 48, dh);
            bst_drop(drop, 1, dh);
#pragma unroll
            for (int i = 0; i < 16; ++i) dh[i] = on ? dh[i] * (((hmask >> i) & 1u) ? 1.0f : 0.01f) : 0.f;
            store_slot_split(sm.g_hi, sm.lo, tid, 1, dh);
        }
        round_trip([&]() {
            mma_proj(tmem + 48, g_desc, 1, wb_desc, 2);                                // W1^T dh
            mma_wgrad(tmem + kWgA, gh, lo_addr, xh, first_tile);                       // stage A: dW2, dW1, db2, db1, d ln2
        });
        tmem_ld16(my_tmem + 48, do1);
#pragma unroll
        for (int i = 0; i < 16; ++i) do1[i] = on ? do1[i] + dz[i] : 0.f;
        PROF(4);
        // ---- C. LayerNorm 1 backward, output projection backward
        float dz1[16];
        {
            float t0[16], dwo[16];
            if (on) layer_norm16_bwd(do1, sm.vec + VG1 * 16, zh1, rstd1, dz1);
            else {
#pragma unroll
                for (int i = 0; i < 16; ++i) dz1[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) { t0[i] = on ? do1[i] * zh1[i] : 0.f; dwo[i] = dz1[i]; }
            bst_drop(drop, 0, dwo);
            store_slot_split(sm.g_hi, sm.lo, tid, 0, dwo);                // G: dwo | do1.zh1 | do1 | 0
            store_slot_split(sm.g_hi, sm.lo, tid, 1, t0);
            store_slot_split(sm.g_hi, sm.lo, tid, 2, do1);
            store_slot_split(sm.g_hi, sm.lo, tid, 3, zero16);
        }
        round_trip([&]() {
            mma_proj(tmem + 48, g_desc, 0, wat_desc, MO);                              // d ctx = Wo^T dwo
            mma_wgrad(tmem + kWgB, gh, lo_addr, xh, first_tile);                       // stage B: dWo, dbo, d ln1
        });
        {
            float dctx[16];
            tmem_ld16(my_tmem + 48, dctx);
            if (on) {
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    float d = 0.f;
#pragma unroll
                    for (int j = 0; j < DH; ++j) d = fmaf(dctx[h * DH + j], ctx[h * DH + j], d);
                    sm.delta[tid * H + h] = d;
                }
            }
            store_row(sm.dc + tid * kBstLd, on ? dctx : zero16);
        }
        __syncthreads();
        PROF(5);
        // ---- D. attention backward: as a query (dq) and as a key (dk, dv)
        float dq[16], dk[16], dv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { dq[i] = 0.f; dk[i] = 0.f; dv[i] = 0.f; }
        if (on) {
            float q[16], kme[16], vme[16], dctx[16];
            load_row(sm.qs + tid * kBstLd, q);
            load_row(sm.ks + tid * kBstLd, kme);
            load_row(sm.vs + tid * kBstLd, vme);
            load_row(sm.dc + tid * kBstLd, dctx);
            float mq[H], ilq[H], dlq[H];
#pragma unroll
            for (int h = 0; h < H; ++h) { mq[h] = sm.mrow[tid * H + h]; ilq[h] = sm.lrow[tid * H + h]; dlq[h] = sm.delta[tid * H + h]; }
            for (int u = 0; u < L; ++u) {
                float kr[16], vr[16];
                load_row(sm.ks + (row0 + u) * kBstLd, kr);
                load_row(sm.vs + (row0 + u) * kBstLd, vr);
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    float sc = 0.f, dA = 0.f;
#pragma unroll
                    for (int j = 0; j < DH; ++j) {
                        sc = fmaf(q[h * DH + j], kr[h * DH + j], sc);
                        dA = fmaf(dctx[h * DH + j], vr[h * DH + j], dA);
                    }
                    const float a  = __expf(sc * scale - mq[h]) * ilq[h];
                    const float dS = a * (dA - dlq[h]) * scale;
#pragma unroll
                    for (int j = 0; j < DH; ++j) dq[h * DH + j] = fmaf(dS, kr[h * DH + j], dq[h * DH + j]);
                }
            }
            if (t < L) {   // this row is a live key: every position of the sample queries it
                for (int tq = 0; tq < T; ++tq) {
                    const int rq = row0 + tq;
                    float qr[16], dr[16];
                    load_row(sm.qs + rq * kBstLd, qr);
                    load_row(sm.dc + rq * kBstLd, dr);
#pragma unroll
                    for (int h = 0; h < H; ++h) {
                        float sc = 0.f, dA = 0.f;
#pragma unroll
                        for (int j = 0; j < DH; ++j) {
                            sc = fmaf(qr[h * DH + j], kme[h * DH + j], sc);
                            dA = fmaf(dr[h * DH + j], vme[h * DH + j], dA);
                        }
                        const float a  = __expf(sc * scale - sm.mrow[rq * H + h]) * sm.lrow[rq * H + h];
                        const float dS = a * (dA - sm.delta[rq * H + h]) * scale;
#pragma unroll
                        for (int j = 0; j < DH; ++j) {
                            dv[h * DH + j] = fmaf(a, dr[h * DH + j], dv[h * DH + j]);
                            dk[h * DH + j] = fmaf(dS, qr[h * DH + j], dk[h * DH + j]);
                        }
                    }
                }
            }
        }
        PROF(6);
        // ---- E. projection backward, input gradient, position-table gradient
        {
            float x[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = on ? qk[i] - sm.pos[t * 16 + i] : 0.f;     // the gathered row up to one rounding (it feeds a bf16 operand)
            store_slot_split(sm.g_hi, sm.lo, tid, 0, dq);                 // G: dq | dk | dv | 0
            store_slot_split(sm.g_hi, sm.lo, tid, 1, dk);
            store_slot_split(sm.g_hi, sm.lo, tid, 2, dv);
            float hq[8], hx[8];                                            // X: qk | x back in slots 0, 1 (hi is all stage C reads)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { hq[j] = qk[8 * half + j]; hx[j] = x[8 * half + j]; }
                store_chunk(sm.x_hi, tid, half, hq);
                store_chunk(sm.x_hi, tid, 2 + half, hx);
            }
        }
        round_trip([&]() {
            mma_proj_acc(tmem + 0, g_desc, 0, wat_desc, MQ, true);                     // Wq^T dq
            mma_proj_acc(tmem + 0, g_desc, 1, wat_desc, MK, false);                    //  + Wk^T dk
            mma_proj_acc(tmem + 16, g_desc, 2, wat_desc, MV, true);                    // Wv^T dv
            mma_wgrad(tmem + kWgC, gh, lo_addr, xh, first_tile);                       // stage C: dWq, dWk, dWv, dbq, dbk, dbv
        });
        {
            float d2[32], dqk[16];
            tmem_ld32(my_tmem + 0, d2);
#pragma unroll
            for (int i = 0; i < 16; ++i) dqk[i] = on ? d2[i] + dz1[i] : 0.f;             // + the residual path of LayerNorm 1
            if (on) {
                float dx[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) dx[i] = dqk[i] + d2[16 + i];
                store_row(g_x + (b * T + t) * 16, dx);
            }
            __syncthreads();                                               // everybody is done with dc as d ctx
            store_row(sm.dc + tid * kBstLd, dqk);
        }
        __syncthreads();
        for (int e = tid; e < T * 16; e += kBtThreads) {          // same thread, same order every tile: no barrier on pacc
            const int tt = e >> 4, n = e & 15;
            float a = sm.pacc[e];
#pragma unroll 2
            for (int ss = 0; ss < ns; ++ss) a += sm.dc[(ss * T + tt) * kBstLd + n];
            sm.pacc[e] = a;
        }
        first_tile = false;
        __syncthreads();
        PROF(7);
    }
    PROF(8);
    PROF_END;

    // ---- per-CTA partials: the position rows from registers, everything else out of TMEM (lanes 0..63)
    float* out = partials + (int64_t)blockIdx.x * (T * 16 + 6 * 256 + 160);
    for (int e = tid; e < T * 16; e += kBtThreads) out[e] = sm.pacc[e];
    float* o = out + T * 16;
    fence_after();
    if (first_tile) {                              // a CTA that owned no tile: nothing was accumulated
        for (int i = tid; i < 6 * 256 + 160; i += kBtThreads) o[i] = 0.f;
    } else if (warp < 2) {
        // lanes 0..63 = (gradient slot g, output n); tcgen05.ld is warp-collective, so every lane of the two warps
        // issues the same loads and keeps what its slot owns.  Offsets inside the partial:
        //   wq 0 bq 256 | wk 272 bk 528 | wv 544 bv 800 | wo 816 bo 1072 | g1 1088 be1 1104 | w1 1120 b1 1376 |
        //   w2 1392 b2 1648 | g2 1664 be2 1680;   -1 = this slot holds nothing there
        const int g = tid >> 4, n = tid & 15;
        const int16_t blocks[5][5] = {   // TMEM column, then the destination of slot 0..3
            {(int16_t)(kWgA + 16), 1392, -1, -1, -1},      // stage A x act : dW2 (df)
            {(int16_t)(kWgA + 0), -1, 1120, -1, -1},       // stage A x o1  : dW1 (dh)
            {(int16_t)(kWgB + 32), 816, -1, -1, -1},       // stage B x ctx : dWo (dwo)
            {(int16_t)(kWgC + 0), 0, 272, -1, -1},         // stage C x qk  : dWq (dq), dWk (dk)
            {(int16_t)(kWgC + 16), -1, -1, 544, -1}};      // stage C x x   : dWv (dv)
        const int16_t sums[3][5] = {     // the column of ones: bias and LayerNorm gradients
            {(int16_t)(kWgA + 48), 1648, 1376, 1664, 1680},     // db2 (df), db1 (dh), d ln2_g (dy.zh2), d ln2_b (dy)
            {(int16_t)(kWgB + 48), 1072, 1088, 1104, -1},       // dbo (dwo), d ln1_g (do1.zh1), d ln1_b (do1)
            {(int16_t)(kWgC + 48), 256, 528, 800, -1}};         // dbq, dbk, dbv
#pragma unroll
        for (int e = 0; e < 5; ++e) {
            float blk[16];
            tmem_ld16(my_tmem + (uint32_t)blocks[e][0], blk);
            const int off = blocks[e][1 + g];
            if (off >= 0) store_row(o + off + n * 16, blk);
        }
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const float v = tmem_ld1(my_tmem + (uint32_t)sums[e][0]);
            const int off = sums[e][1 + g];
            if (off >= 0) o[off + n] = v;
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, kBtBwdTmemCols);
}

template <int H>
static int launch_bwd(const BstParams& p, const float* g_y, const float* g_pool, int g_pool_ld, float* g_x, float* partials,
                      int n_ctas, int32_t* err_flag, cudaStream_t s) {
    const size_t smem = BtBwdSmem::bytes(p.T, H);
    RK_CHECK_ARG(smem <= 227 * 1024, "bst_bwd (tensor): %zu bytes of shared memory", smem);
    RK_CUDA(cudaFuncSetAttribute(bst_bwd_tc_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bst_bwd_tc_kernel<H><<<n_ctas, kBtThreads, smem, s>>>(p, g_y, g_pool, g_pool_ld, g_x, partials, err_flag);
    RK_LAUNCH_CHECK();
    return 0;
}

}  // namespace tc

int bst_tc_bwd_ctas(int64_t B, int T) {
    // two CTAs per SM: 2 x 256 TMEM columns, ~110 KB of shared memory each
    const int64_t tiles = ceil_div(B, kBstRows / T);
    const int64_t cap = (int64_t)sm_count() * 2;
    return (int)(tiles < cap ? tiles : cap);
}

int bst_tc_bwd(const BstParams& p, int nhead, const float* g_y, const float* g_pool, int g_pool_ld, float* g_x,
               float* partials, int n_ctas, int32_t* err_flag, cudaStream_t s) {
    for (int m = 0; m < 6; ++m)
        RK_CHECK_ARG(((uintptr_t)p.w[m] % 16) == 0, "bst (tensor): weight matrix %d must be 16-byte aligned", m);
    switch (nhead) {
        case 1:  return tc::launch_bwd<1>(p, g_y, g_pool, g_pool_ld, g_x, partials, n_ctas, err_flag, s);
        case 2:  return tc::launch_bwd<2>(p, g_y, g_pool, g_pool_ld, g_x, partials, n_ctas, err_flag, s);
        case 4:  return tc::launch_bwd<4>(p, g_y, g_pool, g_pool_ld, g_x, partials, n_ctas, err_flag, s);
        case 8:  return tc::launch_bwd<8>(p, g_y, g_pool, g_pool_ld, g_x, partials, n_ctas, err_flag, s);
        case 16: return tc::launch_bwd<16>(p, g_y, g_pool, g_pool_ld, g_x, partials, n_ctas, err_flag, s);
    }
    RK_CHECK_ARG(false, "bst: nhead %d does not divide d_model 16", nhead);
    return -1;
}

int bst_tc_fwd(const BstParams& p, int nhead, float* y_out, float* pool_out, int pool_ld, int32_t* err_flag,
               cudaStream_t s) {
    for (int m = 0; m < 6; ++m)
        RK_CHECK_ARG(((uintptr_t)p.w[m] % 16) == 0, "bst (tensor): weight matrix %d must be 16-byte aligned", m);
    switch (nhead) {
        case 1:  return tc::launch_fwd<1>(p, y_out, pool_out, pool_ld, err_flag, s);
        case 2:  return tc::launch_fwd<2>(p, y_out, pool_out, pool_ld, err_flag, s);
        case 4:  return tc::launch_fwd<4>(p, y_out, pool_out, pool_ld, err_flag, s);
        case 8:  return tc::launch_fwd<8>(p, y_out, pool_out, pool_ld, err_flag, s);
        case 16: return tc::launch_fwd<16>(p, y_out, pool_out, pool_ld, err_flag, s);
    }
    RK_CHECK_ARG(false, "bst: nhead %d does not divide d_model 16 (the reference's view() fails too)", nhead);
    return -1;
}

}  // namespace rk
