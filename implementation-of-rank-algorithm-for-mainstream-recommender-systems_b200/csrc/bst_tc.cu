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

    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int64_t b0 = tile * p.S;
        const int ns = (int)((p.B - b0) < p.S ? (p.B - b0) : p.S);
        const int rows = ns * T;
        const bool on = tid < rows;
        const int s = on ? tid / T : 0, t = on ? tid - s * T : 0;
        const int64_t b = b0 + s;
        float x[16], qk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { x[i] = 0.f; qk[i] = 0.f; }
        int L = 0;
        if (on) {
            L = bst_len(p, b);
            bst_load_x(p, b, t, x, err_flag);
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
            bst_attend<H>(q, sm.ks, sm.vs, s * T, L, ctx, mh, lh);
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

}  // namespace tc

int bst_tc_bwd_ctas(int64_t B, int T) {
    // the backward of the tensor-core block is the fp32 kernel for now: same partial layout and CTA count
    const int64_t tiles = ceil_div(B, kBstRows / T);
    const int64_t cap = (int64_t)sm_count() * 2;
    return (int)(tiles < cap ? tiles : cap);
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
