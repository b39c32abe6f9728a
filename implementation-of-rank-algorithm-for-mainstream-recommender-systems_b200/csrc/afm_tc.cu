// afm_tc.cu — AFM's attention pooling (AFM/afm.py:92-115) with the attention MLP on tcgen05.
// Same contract as afm.cu (rk_afm_fwd / rk_afm_bwd); selected per module (AFM.attention_precision).
//
// Rows of a tile are (sample, pair) couples: S = 128 / P whole samples per 128-row tile.
//   forward : pre = V W1^T on the tensor core (V = e_i * e_j, split-bf16: hi.hi + lo.hi + hi.lo,
//             fp32 accumulation in TMEM), then per row s = w2 . relu(pre + b1) + b2 straight from
//             TMEM, softmax over the sample's rows, out = sum_p a_p v_p.
// Operand tiles: one 128-byte line per row holding [hi(Kp) | lo(Kp)] bf16 (Kp = D rounded up to
// 16, <= 32), so the three split terms are three (A k-slice, B k-slice) pairings of the same two
// tiles.  HBM traffic is the same as the SIMT kernel's (idx + rows in, out written once).
#include <string.h>
#include "common.cuh"
#include "umma.cuh"

namespace rk {
namespace tc {

constexpr int kAfmTcThreads = 128;
constexpr int kAfmTcRows    = 128;
constexpr int kAfmTcMaxS    = 16;

struct AfmTcParams {
    FieldSet     fs;
    const float* w1;   // [A][D]
    const float* b1;   // [A]
    const float* w2;   // [A]
    const float* b2;   // [1]
    int32_t      D, Kp, A, Ap, P, S, estride;
    int64_t      B, n_tiles;
};

struct AfmTcFwdSmem {
    uint8_t *a1, *b1t;              // [128][128 B], [Ap][128 B]  (b1t = W1 as the B operand)
    float   *e[2];                  // gathered rows of the tile, [S][F][estride], double buffered
    float   *bias, *w2, *score, *attn;
    int     *pi, *pj;
    uint64_t* bar;
    uint32_t* tmem_slot;
    __device__ AfmTcFwdSmem(uint8_t* base, const AfmTcParams& p) {
        uint8_t* q = base;
        a1 = q;   q += 128 * 128;
        b1t = q;  q += 128 * 128;
        const size_t eb = sizeof(float) * p.S * p.fs.F * p.estride;
        e[0] = (float*)q;  q += eb;
        e[1] = (float*)q;  q += eb;
        bias = (float*)q;  q += sizeof(float) * p.Ap;
        w2 = (float*)q;    q += sizeof(float) * p.Ap;
        score = (float*)q; q += sizeof(float) * kAfmTcRows;
        attn = (float*)q;  q += sizeof(float) * kAfmTcRows;
        pi = (int*)q;      q += sizeof(int) * 128;
        pj = (int*)q;      q += sizeof(int) * 128;
        bar = (uint64_t*)q; q += 8;
        tmem_slot = (uint32_t*)q;
    }
    static size_t bytes(const AfmTcParams& p) {
        return 1024 + 2 * 128 * 128 + 2 * sizeof(float) * p.S * p.fs.F * p.estride + sizeof(float) * (2 * p.Ap + 2 * kAfmTcRows) +
               sizeof(int) * 256 + 16;
    }
};

// fp32 x[n] (n <= 32, zero padded to Kp) -> line `r` of a [hi | lo] operand tile
template <int KP>
__device__ __forceinline__ void store_line_split(uint8_t* tile, int r, const float (&x)[KP]) {
    constexpr int CH = KP / 8;                       // 16-byte chunks per half
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        float hi8[8], lo8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            hi8[j] = x[8 * c + j];
            lo8[j] = x[8 * c + j] - __bfloat162float(__float2bfloat16_rn(x[8 * c + j]));
        }
        store_chunk(tile, r, c, hi8);
        store_chunk(tile, r, CH + c, lo8);
    }
}

// Stage the embedding rows of tile `tile` (S samples x F fields) with 16-byte cp.async copies.
__device__ __forceinline__ void afm_issue_rows(const AfmTcParams& p, float* e, int64_t tile, int tid, int32_t* err_flag) {
    const int F = p.fs.F, c4 = p.D >> 2;
    const int64_t b0 = tile * p.S;
    const int n_s = (int)((p.B - b0) < p.S ? (p.B - b0) : p.S);
    for (int item = tid; item < n_s * F * c4; item += kAfmTcThreads) {
        const int c = item % c4, sf = item / c4, f = sf % F, s = sf / F;
        const int64_t row = checked_row(__ldg(p.fs.idx[f] + b0 + s), p.fs.rows[f], err_flag);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                     ::"r"(smem_u32(e + (s * F + f) * p.estride + 4 * c)), "l"(p.fs.weight[f] + row * p.D + 4 * c) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int KP>
__global__ void __launch_bounds__(kAfmTcThreads)
afm_fwd_tc_kernel(const __grid_constant__ AfmTcParams p, float* __restrict__ out, int32_t* err_flag) {
    extern __shared__ uint8_t afm_tc_raw[];
    uint8_t* base = afm_tc_raw + ((1024u - (smem_u32(afm_tc_raw) & 1023u)) & 1023u);
    AfmTcFwdSmem sm(base, p);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int F = p.fs.F, D = p.D, P = p.P;
    constexpr int LO = KP / 8;                       // 16-byte units from the hi half to the lo half of a line

    int64_t tile = blockIdx.x;
    if (tile < p.n_tiles) afm_issue_rows(p, sm.e[0], tile, tid, err_flag);

    if (tid == 0) mbar_init(sm.bar, 1);
    if (warp == 0) tmem_alloc(sm.tmem_slot, 128);
    for (int n = tid; n < p.Ap; n += kAfmTcThreads) {            // W1[n][:] -> B operand line n
        float x[KP];
#pragma unroll
        for (int d = 0; d < KP; ++d) x[d] = (n < p.A && d < D) ? __ldg(p.w1 + n * D + d) : 0.f;
        store_line_split<KP>(sm.b1t, n, x);
        sm.bias[n] = n < p.A ? __ldg(p.b1 + n) : 0.f;
        sm.w2[n]   = n < p.A ? __ldg(p.w2 + n) : 0.f;
    }
    if (tid == 0) {
        int q = 0;
        for (int i = 0; i < F; ++i)
            for (int j = i + 1; j < F; ++j) { sm.pi[q] = i; sm.pj[q] = j; ++q; }
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *sm.tmem_slot;
    const uint32_t my_tmem = tmem + ((uint32_t)(warp * 32) << 16);
    const uint64_t a_desc = umma_desc(smem_u32(sm.a1)), b_desc = umma_desc(smem_u32(sm.b1t));
    const uint32_t idesc = umma_idesc(p.Ap);
    const float b2 = __ldg(p.b2);
    float* pool = reinterpret_cast<float*>(sm.a1);   // [128][32] floats, chunk-swizzled; A1 is free after the MMA
    uint32_t phase = 0;
    int buf = 0;

    for (; tile < p.n_tiles; tile += gridDim.x, buf ^= 1) {
        const int64_t b0 = tile * p.S;
        const int n_s = (int)((p.B - b0) < p.S ? (p.B - b0) : p.S);
        const int n_rows = n_s * P;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                             // rows copied by every thread are visible
        // ---- this thread's pair row v = e_i * e_j, split into the A operand line
        const bool on = tid < n_rows;
        const int s = on ? tid / P : 0, pr = on ? tid - s * P : 0;
        float v[KP];
#pragma unroll
        for (int d = 0; d < KP; ++d) v[d] = 0.f;
        if (on) {
            const float* ei = sm.e[buf] + (s * F + sm.pi[pr]) * p.estride;
            const float* ej = sm.e[buf] + (s * F + sm.pj[pr]) * p.estride;
#pragma unroll
            for (int c = 0; c < KP / 4; ++c)
                if (4 * c < D) {
                    const float4 x = *reinterpret_cast<const float4*>(ei + 4 * c);
                    const float4 y = *reinterpret_cast<const float4*>(ej + 4 * c);
                    v[4 * c] = x.x * y.x; v[4 * c + 1] = x.y * y.y; v[4 * c + 2] = x.z * y.z; v[4 * c + 3] = x.w * y.w;
                }
        }
        store_line_split<KP>(sm.a1, tid, v);
        fence_async_smem();
        fence_before();
        __syncthreads();
        if (tid == 0) {
            fence_after();
#pragma unroll
            for (int term = 0; term < 3; ++term)         // hi.hi + lo.hi + hi.lo
#pragma unroll
                for (int kk = 0; kk < KP / 16; ++kk)
                    umma_bf16(tmem, a_desc + (term == 1 ? LO : 0) + 2 * kk, b_desc + (term == 2 ? LO : 0) + 2 * kk, idesc,
                              (term | kk) > 0);
            umma_commit(sm.bar);
        }
        // while the tensor core works: the next tile's rows (the other e buffer was last read two barriers ago)
        if (tile + gridDim.x < p.n_tiles) afm_issue_rows(p, sm.e[buf ^ 1], tile + gridDim.x, tid, err_flag);
        mbar_wait(sm.bar, phase);
        phase ^= 1;
        fence_after();
        float sc = b2;
        for (int ch = 0; ch < p.Ap / 32; ++ch) {
            float h[32];
            tmem_ld32(my_tmem + 32 * ch, h);
#pragma unroll
            for (int j = 0; j < 32; ++j) sc = fmaf(fmaxf(h[j] + sm.bias[32 * ch + j], 0.f), sm.w2[32 * ch + j], sc);
        }
        sm.score[tid] = sc;
        fence_before();
        __syncthreads();
        // ---- softmax over the sample's P rows: one warp per sample
        for (int ss = warp; ss < n_s; ss += kAfmTcThreads / 32) {
            const int r0 = ss * P;
            float mx = -INFINITY;
            for (int q = lane; q < P; q += 32) mx = fmaxf(mx, sm.score[r0 + q]);
            mx = warp_max(mx);
            float sum = 0.f;
            for (int q = lane; q < P; q += 32) sum += expf(sm.score[r0 + q] - mx);
            const float inv = 1.0f / warp_sum(sum);
            for (int q = lane; q < P; q += 32) sm.attn[r0 + q] = expf(sm.score[r0 + q] - mx) * inv;
        }
        __syncthreads();
        // ---- out[s][:] = sum_p a_p v_p through the (now free) A tile
        {
            const float a = on ? sm.attn[tid] : 0.f;
#pragma unroll
            for (int c = 0; c < KP / 4; ++c)
                *reinterpret_cast<float4*>(pool + tid * 32 + 4 * ((c ^ tid) & 7)) =
                    make_float4(a * v[4 * c], a * v[4 * c + 1], a * v[4 * c + 2], a * v[4 * c + 3]);
        }
        __syncthreads();
        for (int item = tid; item < n_s * D; item += kAfmTcThreads) {
            const int ss = item / D, d = item - ss * D;
            float acc = 0.f;
            for (int q = 0; q < P; ++q) {
                const int r = ss * P + q;
                acc += pool[r * 32 + 4 * (((d >> 2) ^ r) & 7) + (d & 3)];
            }
            out[(b0 + ss) * D + d] = acc;
        }
        __syncthreads();                             // pool / score / attn are rewritten by the next tile
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

}  // namespace tc

static int afm_tc_fill(const rk_field_t* fields, int F, const float* w1, const float* b1, const float* w2,
                       const float* b2, int A, int64_t B, tc::AfmTcParams* p) {
    if (int rc = pack_fields(fields, F, &p->fs)) return rc;
    RK_CHECK_ARG(F >= 2 && F <= 16, "afm: %d fields (supported: 2..16, i.e. <= 120 pairs)", F);
    const int D = fields[0].dim;
    for (int f = 0; f < F; ++f) {
        RK_CHECK_ARG(fields[f].dim == D, "afm: field %d dim %d != %d", f, fields[f].dim, D);
        RK_CHECK_ARG(((uintptr_t)fields[f].weight % 16) == 0, "afm: table %d not 16-byte aligned", f);
    }
    RK_CHECK_ARG(D % 4 == 0 && D >= 4 && D <= 32, "afm: embedding_dim %d not in {4,8,..,32}", D);
    RK_CHECK_ARG(A >= 1 && A <= 128, "afm: attention_factor %d outside [1,128]", A);
    RK_CHECK_ARG(w1 && b1 && w2 && b2, "afm: NULL attention weight");
    p->w1 = w1; p->b1 = b1; p->w2 = w2; p->b2 = b2;
    p->D = D; p->Kp = (D + 15) / 16 * 16; p->A = A; p->Ap = (A + 31) / 32 * 32;
    p->P = F * (F - 1) / 2;
    int S = tc::kAfmTcRows / p->P;
    p->S = S > tc::kAfmTcMaxS ? tc::kAfmTcMaxS : S;
    p->estride = D + 4;                  // float4 reads of different rows land in different bank groups
    p->B = B;
    p->n_tiles = ceil_div(B, p->S);
    return 0;
}

}  // namespace rk

extern "C" {

int rk_afm_tc_fwd(const rk_field_t* fields, int F, const float* w1, const float* b1, const float* w2,
                  const float* b2, int A, int64_t B, float* out, int32_t* err_flag, rk_stream_t stream_) {
    using namespace rk;
    tc::AfmTcParams p;
    if (int rc = afm_tc_fill(fields, F, w1, b1, w2, b2, A, B, &p)) return rc;
    RK_CHECK_ARG(out, "afm_tc_fwd: out is NULL");
    if (B == 0) return 0;
    const size_t smem = tc::AfmTcFwdSmem::bytes(p);
    RK_CHECK_ARG(smem <= 227 * 1024, "afm_tc_fwd: %zu bytes of shared memory", smem);
    auto kernel = p.Kp == 16 ? tc::afm_fwd_tc_kernel<16> : tc::afm_fwd_tc_kernel<32>;
    RK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = p.n_tiles;
    const int64_t cap = (int64_t)sm_count() * 4;      // 4 x 128 TMEM columns per SM
    if (grid > cap) grid = cap;
    kernel<<<(int)grid, tc::kAfmTcThreads, smem, (cudaStream_t)stream_>>>(p, out, err_flag);
    RK_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
