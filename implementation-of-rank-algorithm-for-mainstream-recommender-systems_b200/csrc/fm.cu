// fm.cu — DeepFM's FM part (DeepFM/deepfm.py:121-142), fused with the 2F embedding gathers:
//   deep_input = cat_f e_f            (:142)
//   fm_first   = sum_f w_f[idx_f]     (:123-127)
//   fm_second  = 0.5 * sum_d[(sum_f e_f)^2 - sum_f e_f^2]   (:134-140, the sum-square trick)
// and the per-occurrence gradient of the second-order rows for the backward.
//
// Layout: LPS = next_pow2(D/VEC) lanes per sample, lane c owns columns [c*VEC, c*VEC+VEC) of
// every field, so a table row and its slot in deep_input are read/written by adjacent lanes
// (64 contiguous bytes at D=16); the sum over d is a shuffle reduction inside the LPS-lane
// group.  HBM-bound: idx + rows read once, deep_input written once.
#include "common.cuh"

namespace rk {

struct FmArgs {
    FieldSet     fs;                          // second-order tables (off unused)
    const float* first[RK_MAX_FIELDS];        // first-order tables [rows, 1]
    int32_t      D;
};

template <int VEC>
__global__ void __launch_bounds__(256)
deepfm_fwd_kernel(const __grid_constant__ FmArgs a, int64_t B, int lps_log2,
                  float* __restrict__ deep_input, float* __restrict__ fm_first,
                  float* __restrict__ fm_second, int32_t* err_flag) {
    const int     lps = 1 << lps_log2;
    const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t b   = gid >> lps_log2;
    const int     c   = (int)(gid & (lps - 1));
    const int     F = a.fs.F, D = a.D;
    const bool    live   = b < B;
    const bool    active = live && c * VEC < D;
    Vec<VEC> S, Q;
    vec_zero(S);
    vec_zero(Q);
    float first = 0.f;
    if (live) {
#pragma unroll 2
        for (int f = 0; f < F; ++f) {
            const int64_t row = checked_row(__ldg(a.fs.idx[f] + b), a.fs.rows[f], err_flag);
            if (active) {
                Vec<VEC> e;
                e.load(a.fs.weight[f] + row * D + c * VEC);
                e.store(deep_input + (b * F + f) * D + c * VEC);
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    S.v[k] += e.v[k];
                    Q.v[k] += e.v[k] * e.v[k];
                }
            }
            if (c == 0) first += __ldg(a.first[f] + row);
        }
    }
    float part = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) part += S.v[k] * S.v[k] - Q.v[k];
    for (int o = lps >> 1; o > 0; o >>= 1) part += __shfl_xor_sync(kFull, part, o);
    if (live && c == 0) {
        fm_first[b]  = first;
        fm_second[b] = 0.5f * part;
    }
}

// g_rows[b, f, :] = g_deep[b, f, :] + g_second[b] * (S_b - e_{b,f}),  S_b = sum_f e_{b,f}
template <int VEC>
__global__ void __launch_bounds__(256)
deepfm_bwd_kernel(const float* __restrict__ deep_input, const float* __restrict__ g_deep,
                  const float* __restrict__ g_second, int F, int D, int64_t B, int lps_log2,
                  float* __restrict__ g_rows) {
    const int     lps = 1 << lps_log2;
    const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t b   = gid >> lps_log2;
    const int     c   = (int)(gid & (lps - 1));
    if (b >= B || c * VEC >= D) return;
    const float gs = g_second ? __ldg(g_second + b) : 0.f;
    Vec<VEC> S;
    vec_zero(S);
    const float* in = deep_input + b * F * D + c * VEC;
    for (int f = 0; f < F; ++f) {
        Vec<VEC> e;
        e.load(in + f * D);
#pragma unroll
        for (int k = 0; k < VEC; ++k) S.v[k] += e.v[k];
    }
    for (int f = 0; f < F; ++f) {
        Vec<VEC> e, g;
        e.load(in + f * D);
        if (g_deep) g.load(g_deep + (b * F + f) * D + c * VEC);
        else        vec_zero(g);
#pragma unroll
        for (int k = 0; k < VEC; ++k) g.v[k] += gs * (S.v[k] - e.v[k]);
        g.store(g_rows + (b * F + f) * D + c * VEC);
    }
}

static int ilog2_ceil(int x) {
    int l = 0;
    while ((1 << l) < x) ++l;
    return l;
}

}  // namespace rk

extern "C" {

int rk_deepfm_fwd(const rk_field_t* second, const float* const* first_weight, int F, int64_t B,
                  float* deep_input, float* fm_first, float* fm_second, int32_t* err_flag,
                  rk_stream_t stream_) {
    using namespace rk;
    cudaStream_t s = (cudaStream_t)stream_;
    FmArgs a;
    RK_CHECK_ARG(F >= 1, "deepfm_fwd: F=%d", F);
    if (int rc = pack_fields(second, F, &a.fs)) return rc;
    RK_CHECK_ARG(first_weight && deep_input && fm_first && fm_second, "deepfm_fwd: NULL pointer");
    const int D = second[0].dim;
    for (int f = 0; f < F; ++f) {
        RK_CHECK_ARG(second[f].dim == D, "deepfm_fwd: field %d dim %d != %d", f, second[f].dim, D);
        RK_CHECK_ARG(first_weight[f], "deepfm_fwd: first-order table %d is NULL", f);
        a.first[f] = first_weight[f];
    }
    a.D = D;
    if (B == 0) return 0;
    int vec = D % 4 == 0 ? 4 : (D % 2 == 0 ? 2 : 1);
    auto mis = [&](const void* p) { return ((uintptr_t)p % (4 * vec)) != 0; };
    while (vec > 1) {
        bool bad = mis(deep_input);
        for (int f = 0; f < F; ++f) bad = bad || mis(second[f].weight);
        if (!bad) break;
        vec >>= 1;
    }
    const int lanes = D / vec;
    RK_CHECK_ARG(lanes <= 32, "deepfm_fwd: embedding_dim %d too wide (max %d)", D, 32 * vec);
    const int     lg      = ilog2_ceil(lanes);
    const int64_t threads = B << lg;
    const int     grid    = (int)ceil_div(threads, 256);
    if (vec == 4)
        deepfm_fwd_kernel<4><<<grid, 256, 0, s>>>(a, B, lg, deep_input, fm_first, fm_second, err_flag);
    else if (vec == 2)
        deepfm_fwd_kernel<2><<<grid, 256, 0, s>>>(a, B, lg, deep_input, fm_first, fm_second, err_flag);
    else
        deepfm_fwd_kernel<1><<<grid, 256, 0, s>>>(a, B, lg, deep_input, fm_first, fm_second, err_flag);
    RK_LAUNCH_CHECK();
    return 0;
}

int rk_deepfm_bwd(const float* deep_input, const float* g_deep, const float* g_second, int F,
                  int D, int64_t B, float* g_rows, rk_stream_t stream_) {
    using namespace rk;
    cudaStream_t s = (cudaStream_t)stream_;
    RK_CHECK_ARG(F >= 1 && D >= 1 && B >= 0 && deep_input && g_rows, "deepfm_bwd: bad argument");
    if (B == 0) return 0;
    int vec = D % 4 == 0 ? 4 : (D % 2 == 0 ? 2 : 1);
    auto mis = [&](const void* p) { return p && ((uintptr_t)p % (4 * vec)) != 0; };
    while (vec > 1 && (mis(deep_input) || mis(g_deep) || mis(g_rows))) vec >>= 1;
    const int lanes = D / vec;
    RK_CHECK_ARG(lanes <= 32, "deepfm_bwd: embedding_dim %d too wide", D);
    const int     lg      = ilog2_ceil(lanes);
    const int64_t threads = B << lg;
    const int     grid    = (int)ceil_div(threads, 256);
    if (vec == 4)
        deepfm_bwd_kernel<4><<<grid, 256, 0, s>>>(deep_input, g_deep, g_second, F, D, B, lg, g_rows);
    else if (vec == 2)
        deepfm_bwd_kernel<2><<<grid, 256, 0, s>>>(deep_input, g_deep, g_second, F, D, B, lg, g_rows);
    else
        deepfm_bwd_kernel<1><<<grid, 256, 0, s>>>(deep_input, g_deep, g_second, F, D, B, lg, g_rows);
    RK_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
