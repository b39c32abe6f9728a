// fwfm.cu — Field-weighted FM (FwFM/fwfm.py:87-139), fused with its 2F embedding gathers:
//   z = sum_f w_f[idx_f] + sum_{i<j} r_p <e_i, e_j> + bias,  y = sigmoid(z)        (:118-139)
// (p enumerates the pairs in the reference's loop order, i outer, j inner, :126-135), and the
// backward: per-occurrence row gradients g_e_i = g_z sum_{j != i} r_ij e_j, g_z itself for the
// first-order tables, and the batch-wide sums g_r_p = sum_b g_z <e_i, e_j>, g_bias = sum_b g_z.
//
// Layout (as fm.cu): LPS = next_pow2(D/VEC) lanes per sample, lane c owns columns
// [c*VEC, c*VEC+VEC) of every field and keeps them in registers; dot products over d are shuffle
// reductions inside the lane group.  HBM-bound: idx + rows read once, emb written once (saved
// for the backward), y written once.  The batch-wide sums are deterministic: fixed-order sums
// inside the CTA, one partial row per CTA, then launch_reduce_partials.
#include "common.cuh"

namespace rk {

constexpr int kFwThreads = 256;

struct FwArgs {
    FieldSet     fs;                          // embedding tables (off unused)
    const float* first[RK_MAX_FIELDS];        // first-order tables [rows, 1]
    int32_t      D;
};

__host__ __device__ inline int fw_pair(int i, int j, int F) { return i * (2 * F - i - 1) / 2 + (j - i - 1); }

template <int VEC, int MAXF>
__global__ void __launch_bounds__(kFwThreads)
fwfm_fwd_kernel(const __grid_constant__ FwArgs a, const float* __restrict__ field_weight,
                const float* __restrict__ bias, int64_t B, int lps_log2, float* __restrict__ emb,
                float* __restrict__ y, int32_t* err_flag) {
    __shared__ float r[RK_MAX_FIELDS * (RK_MAX_FIELDS - 1) / 2];
    const int F = a.fs.F, D = a.D, P = F * (F - 1) / 2;
    for (int p = threadIdx.x; p < P; p += kFwThreads) r[p] = __ldg(field_weight + p);
    __syncthreads();
    const int     lps = 1 << lps_log2;
    const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t b   = gid >> lps_log2;
    const int     c   = (int)(gid & (lps - 1));
    const bool    live   = b < B;
    const bool    active = live && c * VEC < D;
    Vec<VEC> e[MAXF];
    float first = 0.f;
#pragma unroll
    for (int f = 0; f < MAXF; ++f) {
        vec_zero(e[f]);
        if (f < F && live) {
            const int64_t row = checked_row(__ldg(a.fs.idx[f] + b), a.fs.rows[f], err_flag);
            if (active) {
                e[f].load(a.fs.weight[f] + row * D + c * VEC);
                e[f].store(emb + (b * F + f) * D + c * VEC);
            }
            if (c == 0) first += __ldg(a.first[f] + row);
        }
    }
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < MAXF; ++i)
#pragma unroll
        for (int j = i + 1; j < MAXF; ++j)
            if (j < F) {
                float dot = 0.f;
#pragma unroll
                for (int k = 0; k < VEC; ++k) dot = fmaf(e[i].v[k], e[j].v[k], dot);
                part = fmaf(r[fw_pair(i, j, F)], dot, part);
            }
    for (int o = lps >> 1; o > 0; o >>= 1) part += __shfl_xor_sync(kFull, part, o);
    if (live && c == 0) {
        const float z = first + part + __ldg(bias);
        y[b] = 1.0f / (1.0f + expf(-z));
    }
}

// One pass over the batch per CTA slice (grid-stride over "rounds" of kFwThreads/LPS samples).
// Per round the CTA stages g_z <e_i,e_j> of its samples in shared memory [sample][pair]; thread
// (pair, part) then adds the samples of its part in order, and the parts are folded in order, so
// every partial sum has a fixed association.
template <int VEC, int MAXF>
__global__ void __launch_bounds__(kFwThreads)
fwfm_bwd_kernel(const float* __restrict__ emb, const float* __restrict__ y, const float* __restrict__ g_y,
                const float* __restrict__ field_weight, int F, int D, int64_t B, int lps_log2,
                float* __restrict__ g_rows, float* __restrict__ g_z_out, float* __restrict__ partials) {
    extern __shared__ float fw_smem[];
    const int P = F * (F - 1) / 2, P1 = P + 1;           // + the bias column
    const int lps = 1 << lps_log2;
    const int spr = kFwThreads >> lps_log2;              // samples per round
    float* r     = fw_smem;                              // [P]
    float* stage = r + ((P + 3) & ~3);                   // [spr][P1]
    float* fold  = stage + spr * P1;                     // [8][P1]
    for (int p = threadIdx.x; p < P; p += kFwThreads) r[p] = __ldg(field_weight + p);
    const int s_loc = threadIdx.x >> lps_log2;
    const int c     = threadIdx.x & (lps - 1);
    // this thread's running sums: columns col = threadIdx.x % 32 + 32 q of part threadIdx.x / 32
    constexpr int kMaxQ = (RK_MAX_FIELDS * (RK_MAX_FIELDS - 1) / 2 + 1 + 31) / 32;
    float run[kMaxQ];
#pragma unroll
    for (int q = 0; q < kMaxQ; ++q) run[q] = 0.f;
    __syncthreads();

    const int64_t n_rounds = (B + spr - 1) / spr;
    for (int64_t round = blockIdx.x; round < n_rounds; round += gridDim.x) {
        const int64_t b = round * spr + s_loc;
        const bool live = b < B, active = live && c * VEC < D;
        Vec<VEC> e[MAXF];
#pragma unroll
        for (int f = 0; f < MAXF; ++f) {
            vec_zero(e[f]);
            if (f < F && active) e[f].load(emb + (b * F + f) * D + c * VEC);
        }
        float gz = 0.f;
        if (live) {
            const float yy = __ldg(y + b);
            gz = __ldg(g_y + b) * yy * (1.0f - yy);
        }
        Vec<VEC> g[MAXF];
#pragma unroll
        for (int f = 0; f < MAXF; ++f) vec_zero(g[f]);
#pragma unroll
        for (int i = 0; i < MAXF; ++i)
#pragma unroll
            for (int j = i + 1; j < MAXF; ++j)
                if (j < F) {
                    const float rp = r[fw_pair(i, j, F)];
                    float dot = 0.f;
#pragma unroll
                    for (int k = 0; k < VEC; ++k) {
                        dot = fmaf(e[i].v[k], e[j].v[k], dot);
                        g[i].v[k] = fmaf(rp, e[j].v[k], g[i].v[k]);
                        g[j].v[k] = fmaf(rp, e[i].v[k], g[j].v[k]);
                    }
                    for (int o = lps >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(kFull, dot, o);
                    if (c == 0) stage[s_loc * P1 + fw_pair(i, j, F)] = gz * dot;
                }
        if (c == 0) stage[s_loc * P1 + P] = gz;
#pragma unroll
        for (int f = 0; f < MAXF; ++f)
            if (f < F && active) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) g[f].v[k] *= gz;
                g[f].store(g_rows + (b * F + f) * D + c * VEC);
            }
        if (live && c == 0) g_z_out[b] = gz;
        __syncthreads();
        {
            const int part = threadIdx.x >> 5, col0 = threadIdx.x & 31;
#pragma unroll
            for (int q = 0; q < kMaxQ; ++q) {
                const int col = col0 + 32 * q;
                if (col < P1) {
                    float acc = run[q];
                    for (int s = part; s < spr; s += kFwThreads / 32) acc += stage[s * P1 + col];
                    run[q] = acc;
                }
            }
        }
        __syncthreads();
    }
    // fold the 8 parts in order; one partial row per CTA
    {
        const int part = threadIdx.x >> 5, col0 = threadIdx.x & 31;
#pragma unroll
        for (int q = 0; q < kMaxQ; ++q) {
            const int col = col0 + 32 * q;
            if (col < P1) fold[part * P1 + col] = run[q];
        }
    }
    __syncthreads();
    for (int col = threadIdx.x; col < P1; col += kFwThreads) {
        float acc = 0.f;
#pragma unroll
        for (int part = 0; part < kFwThreads / 32; ++part) acc += fold[part * P1 + col];
        partials[(int64_t)blockIdx.x * P1 + col] = acc;
    }
}

static int fw_ilog2_ceil(int x) {
    int l = 0;
    while ((1 << l) < x) ++l;
    return l;
}

static int fw_bwd_grid(int64_t B, int lps_log2) {
    const int64_t rounds = ceil_div(B, kFwThreads >> lps_log2);
    const int64_t cap = (int64_t)sm_count() * 2;
    return (int)(rounds < cap ? (rounds > 0 ? rounds : 1) : cap);
}

}  // namespace rk

extern "C" {

int rk_fwfm_fwd(const rk_field_t* second, const float* const* first_weight, const float* field_weight,
                const float* bias, int F, int64_t B, float* emb, float* y, int32_t* err_flag,
                rk_stream_t stream_) {
    using namespace rk;
    cudaStream_t s = (cudaStream_t)stream_;
    FwArgs a;
    RK_CHECK_ARG(F >= 2, "fwfm_fwd: F=%d (needs at least one pair)", F);
    if (int rc = pack_fields(second, F, &a.fs)) return rc;
    RK_CHECK_ARG(first_weight && field_weight && bias && emb && y, "fwfm_fwd: NULL pointer");
    const int D = second[0].dim;
    for (int f = 0; f < F; ++f) {
        RK_CHECK_ARG(second[f].dim == D, "fwfm_fwd: field %d dim %d != %d", f, second[f].dim, D);
        RK_CHECK_ARG(first_weight[f], "fwfm_fwd: first-order table %d is NULL", f);
        a.first[f] = first_weight[f];
    }
    a.D = D;
    if (B == 0) return 0;
    int vec = D % 4 == 0 ? 4 : (D % 2 == 0 ? 2 : 1);
    auto mis = [&](const void* p) { return ((uintptr_t)p % (4 * vec)) != 0; };
    while (vec > 1) {
        bool bad = mis(emb);
        for (int f = 0; f < F; ++f) bad = bad || mis(second[f].weight);
        if (!bad) break;
        vec >>= 1;
    }
    const int lanes = D / vec;
    RK_CHECK_ARG(lanes <= 32, "fwfm_fwd: embed_dim %d too wide (max %d)", D, 32 * vec);
    const int     lg      = fw_ilog2_ceil(lanes);
    const int64_t threads = B << lg;
    const int     grid    = (int)ceil_div(threads, kFwThreads);
#define RK_FW_FWD(V, M) fwfm_fwd_kernel<V, M><<<grid, kFwThreads, 0, s>>>(a, field_weight, bias, B, lg, emb, y, err_flag)
    if (F <= 8) {
        if (vec == 4) RK_FW_FWD(4, 8); else if (vec == 2) RK_FW_FWD(2, 8); else RK_FW_FWD(1, 8);
    } else {
        if (vec == 4) RK_FW_FWD(4, RK_MAX_FIELDS); else if (vec == 2) RK_FW_FWD(2, RK_MAX_FIELDS); else RK_FW_FWD(1, RK_MAX_FIELDS);
    }
#undef RK_FW_FWD
    RK_LAUNCH_CHECK();
    return 0;
}

int rk_fwfm_bwd_ctas(void) { return rk::sm_count() * 2; }

int rk_fwfm_bwd(const float* emb, const float* y, const float* g_y, const float* field_weight, int F,
                int D, int64_t B, float* g_rows, float* g_z, float* partials, float* g_pair,
                rk_stream_t stream_) {
    using namespace rk;
    cudaStream_t s = (cudaStream_t)stream_;
    RK_CHECK_ARG(F >= 2 && F <= RK_MAX_FIELDS && D >= 1 && B >= 0, "fwfm_bwd: F=%d D=%d B=%lld", F, D, (long long)B);
    RK_CHECK_ARG(emb && y && g_y && field_weight && g_rows && g_z && partials && g_pair, "fwfm_bwd: NULL pointer");
    const int P1 = F * (F - 1) / 2 + 1;
    if (B == 0) {
        RK_CUDA(cudaMemsetAsync(g_pair, 0, sizeof(float) * P1, s));
        return 0;
    }
    int vec = D % 4 == 0 ? 4 : (D % 2 == 0 ? 2 : 1);
    auto mis = [&](const void* p) { return ((uintptr_t)p % (4 * vec)) != 0; };
    while (vec > 1 && (mis(emb) || mis(g_rows))) vec >>= 1;
    const int lanes = D / vec;
    RK_CHECK_ARG(lanes <= 32, "fwfm_bwd: embed_dim %d too wide", D);
    const int lg   = fw_ilog2_ceil(lanes);
    const int spr  = kFwThreads >> lg;
    const int grid = fw_bwd_grid(B, lg);
    const size_t smem = sizeof(float) * (((P1 - 1 + 3) & ~3) + (size_t)spr * P1 + 8 * P1);
    RK_CHECK_ARG(smem <= 200 * 1024, "fwfm_bwd: %zu bytes of shared memory (F=%d, D=%d)", smem, F, D);
#define RK_FW_BWD(V, M)                                                                                       \
    do {                                                                                                      \
        RK_CUDA(cudaFuncSetAttribute(fwfm_bwd_kernel<V, M>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                     (int)smem));                                                             \
        fwfm_bwd_kernel<V, M><<<grid, kFwThreads, smem, s>>>(emb, y, g_y, field_weight, F, D, B, lg, g_rows,  \
                                                             g_z, partials);                                  \
    } while (0)
    if (F <= 8) {
        if (vec == 4) RK_FW_BWD(4, 8); else if (vec == 2) RK_FW_BWD(2, 8); else RK_FW_BWD(1, 8);
    } else {
        if (vec == 4) RK_FW_BWD(4, RK_MAX_FIELDS); else if (vec == 2) RK_FW_BWD(2, RK_MAX_FIELDS); else RK_FW_BWD(1, RK_MAX_FIELDS);
    }
#undef RK_FW_BWD
    RK_LAUNCH_CHECK();
    return launch_reduce_partials(partials, grid, P1, g_pair, s);
}

}  // extern "C"
