// concat_cross.cu — fused gather+concat, and DCN CrossNet forward/backward on top of it.
//   rk_gather_concat_fwd : the per-field lookup loop + torch.cat of DCN/dcn.py:163-169 and
//                          DeepCrossing/deepcrossing.py:148-155 in one launch.
//   rk_crossnet_fwd/bwd  : cross_layer (DCN/dcn.py:25-50) applied L times (DCN/dcn.py:171-173),
//                          fused with that gather; backward w.r.t. x0 only (w_l, b_l are drawn
//                          fresh inside every call and never registered, DCN/dcn.py:37-41).
//
// Layout: one warp per sample; the concatenated row of d floats is cut into units of VEC floats
// (VEC = widest of 4/2/1 dividing every field dim and offset); lane l owns units l, l+32, ...
// Reads of a table row and writes of the output row are contiguous across lanes; x.w dot
// products are warp-shuffle reductions.  HBM-bound: every input byte read once, every output
// byte written once, w/b staged in shared memory once per CTA.
#include <string.h>
#include "common.cuh"

namespace rk {

struct ConcatArgs {
    FieldSet     fs;
    const float* dense;
    int32_t      n_dense;
    int32_t      d;        // total row width (floats)
    int32_t      n_units;  // d / VEC
};

// Where a lane's unit comes from: fld < 0 -> dense input, else field fld at column sub.
struct UnitSrc {
    int fld;
    int sub;
    int col;
    bool on;
};

template <int VEC>
__device__ __forceinline__ UnitSrc locate_unit(const ConcatArgs& a, int u) {
    UnitSrc s;
    s.on  = u < a.n_units;
    s.col = u * VEC;
    s.fld = -1;
    s.sub = s.col;
    if (s.on && s.col >= a.n_dense) {
        for (int f = 0; f < a.fs.F; ++f)
            if (s.col >= a.fs.off[f] && s.col < a.fs.off[f] + a.fs.dim[f]) {
                s.fld = f;
                s.sub = s.col - a.fs.off[f];
            }
    }
    return s;
}

template <int VEC>
__device__ __forceinline__ void load_unit(const ConcatArgs& a, const UnitSrc& s, int64_t b,
                                          Vec<VEC>& x, int32_t* err_flag) {
    if (!s.on) { vec_zero(x); return; }
    if (s.fld < 0) {
        x.load(a.dense + b * a.n_dense + s.sub);
    } else {
        const int64_t row = checked_row(__ldg(a.fs.idx[s.fld] + b), a.fs.rows[s.fld], err_flag);
        x.load(a.fs.weight[s.fld] + row * a.fs.dim[s.fld] + s.sub);
    }
}

template <int VEC, int R>
__global__ void __launch_bounds__(256)
gather_concat_kernel(const __grid_constant__ ConcatArgs a, int64_t B, float* __restrict__ out,
                     int ld_out, int32_t* err_flag) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    UnitSrc src[R];
#pragma unroll
    for (int r = 0; r < R; ++r) src[r] = locate_unit<VEC>(a, lane + 32 * r);
    for (int64_t b = warp0; b < B; b += nwarp) {
        Vec<VEC> x[R];
#pragma unroll
        for (int r = 0; r < R; ++r) load_unit<VEC>(a, src[r], b, x[r], err_flag);
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (src[r].on) x[r].store(out + b * ld_out + src[r].col);
    }
}

template <int VEC, int R>
__global__ void __launch_bounds__(256)
crossnet_fwd_kernel(const __grid_constant__ ConcatArgs a, const float* __restrict__ w,
                    const float* __restrict__ bias, int L, int64_t B,
                    float* __restrict__ concat_all, float* __restrict__ cross_vec,
                    int32_t* err_flag) {
    extern __shared__ float sm[];
    float* sw = sm;              // [L, d]
    float* sb = sm + L * a.d;    // [L, d]
    for (int i = threadIdx.x; i < L * a.d; i += blockDim.x) {
        sw[i] = w[i];
        sb[i] = bias[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    UnitSrc src[R];
#pragma unroll
    for (int r = 0; r < R; ++r) src[r] = locate_unit<VEC>(a, lane + 32 * r);

    for (int64_t b = warp0; b < B; b += nwarp) {
        Vec<VEC> x0[R], xl[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            load_unit<VEC>(a, src[r], b, x0[r], err_flag);
            xl[r] = x0[r];
            if (src[r].on) x0[r].store(concat_all + b * a.d + src[r].col);
        }
        for (int l = 0; l < L; ++l) {
            float dot = 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (src[r].on) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) dot += xl[r].v[e] * sw[l * a.d + src[r].col + e];
                }
            const float s = warp_sum(dot);
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (src[r].on) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e)
                        xl[r].v[e] = x0[r].v[e] * s + sb[l * a.d + src[r].col + e] + xl[r].v[e];
                }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (src[r].on) xl[r].store(cross_vec + b * a.d + src[r].col);
    }
}

// Backward through the L cross layers w.r.t. x0 (which is both the anchor and layer 0's input).
//   forward : x_{l+1} = x0 * s_l + b_l + x_l,  s_l = x_l . w_l
//   backward: g_x0 += g_{l+1} * s_l ;  g_l = g_{l+1} + w_l * (g_{l+1} . x0)
template <int VEC, int R>
__global__ void __launch_bounds__(256)
crossnet_bwd_kernel(const float* __restrict__ concat_all, const float* __restrict__ w,
                    const float* __restrict__ bias, int L, int d, int64_t B,
                    const float* __restrict__ g_concat, const float* __restrict__ g_cross,
                    float* __restrict__ g_x0) {
    extern __shared__ float sm[];
    float* sw = sm;
    float* sb = sm + L * d;
    float* ss = sm + 2 * L * d;  // [warps, RK_MAX_LAYERS] forward scalars s_l
    for (int i = threadIdx.x; i < L * d; i += blockDim.x) {
        sw[i] = w[i];
        sb[i] = bias[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* my_s = ss + wid * RK_MAX_LAYERS;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int n_units = d / VEC;
    bool on[R];
    int  col[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        on[r]  = lane + 32 * r < n_units;
        col[r] = (lane + 32 * r) * VEC;
    }
    for (int64_t b = warp0; b < B; b += nwarp) {
        Vec<VEC> x0[R], xl[R], g[R], gx[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            vec_zero(x0[r]); vec_zero(g[r]); vec_zero(gx[r]);
            if (on[r]) {
                x0[r].load(concat_all + b * d + col[r]);
                if (g_cross)  g[r].load(g_cross + b * d + col[r]);
                if (g_concat) gx[r].load(g_concat + b * d + col[r]);
            }
            xl[r] = x0[r];
        }
        // recompute the forward scalars
        for (int l = 0; l < L; ++l) {
            float dot = 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (on[r]) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) dot += xl[r].v[e] * sw[l * d + col[r] + e];
                }
            const float s = warp_sum(dot);
            if (lane == 0) my_s[l] = s;
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (on[r]) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e)
                        xl[r].v[e] = x0[r].v[e] * s + sb[l * d + col[r] + e] + xl[r].v[e];
                }
        }
        __syncwarp();
        for (int l = L - 1; l >= 0; --l) {
            const float s = my_s[l];
            float dot = 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (on[r]) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        gx[r].v[e] += g[r].v[e] * s;
                        dot += g[r].v[e] * x0[r].v[e];
                    }
                }
            const float t = warp_sum(dot);
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (on[r]) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) g[r].v[e] += sw[l * d + col[r] + e] * t;
                }
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (on[r]) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) gx[r].v[e] += g[r].v[e];
                gx[r].store(g_x0 + b * d + col[r]);
            }
    }
}

// Stand-alone cross_layer(x0, xl) (DCN/dcn.py:25-50) for callers that pass both tensors.
__global__ void __launch_bounds__(256)
cross_layer_fwd_kernel(const float* __restrict__ x0, const float* __restrict__ xl,
                       const float* __restrict__ w, const float* __restrict__ bias, int d,
                       int64_t B, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp0; b < B; b += nwarp) {
        float dot = 0.f;
        for (int c = lane; c < d; c += 32) dot += xl[b * d + c] * __ldg(w + c);
        const float s = warp_sum(dot);
        for (int c = lane; c < d; c += 32)
            out[b * d + c] = x0[b * d + c] * s + __ldg(bias + c) + xl[b * d + c];
    }
}

__global__ void __launch_bounds__(256)
cross_layer_bwd_kernel(const float* __restrict__ x0, const float* __restrict__ xl,
                       const float* __restrict__ w, int d, int64_t B,
                       const float* __restrict__ g, float* __restrict__ g_x0,
                       float* __restrict__ g_xl) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp0; b < B; b += nwarp) {
        float ds = 0.f, dt = 0.f;
        for (int c = lane; c < d; c += 32) {
            ds += xl[b * d + c] * __ldg(w + c);
            dt += g[b * d + c] * x0[b * d + c];
        }
        const float s = warp_sum(ds), t = warp_sum(dt);
        for (int c = lane; c < d; c += 32) {
            const float gv = g[b * d + c];
            g_x0[b * d + c] = gv * s;
            g_xl[b * d + c] = gv + __ldg(w + c) * t;
        }
    }
}

static int fill_concat_args(const rk_field_t* fields, int F, const float* dense, int n_dense,
                            ConcatArgs* a, int* vec) {
    if (int rc = pack_fields(fields, F, &a->fs)) return rc;
    RK_CHECK_ARG(n_dense >= 0 && (n_dense == 0 || dense), "concat: n_dense=%d dense=%p", n_dense,
                 (const void*)dense);
    int d = n_dense;
    for (int f = 0; f < F; ++f) {
        RK_CHECK_ARG(fields[f].out_off >= n_dense, "concat: field %d overlaps the dense block", f);
        const int end = fields[f].out_off + fields[f].dim;
        if (end > d) d = end;
    }
    {   // the fields must tile [n_dense, d) exactly: no gaps, no overlaps
        RK_CHECK_ARG(d <= 512, "concat: row of %d floats is wider than 512", d);
        unsigned char seen[512];
        memset(seen, 0, sizeof(seen));
        for (int f = 0; f < F; ++f)
            for (int c = fields[f].out_off; c < fields[f].out_off + fields[f].dim; ++c) {
                RK_CHECK_ARG(!seen[c], "concat: column %d is written by two fields", c);
                seen[c] = 1;
            }
        for (int c = n_dense; c < d; ++c)
            RK_CHECK_ARG(seen[c], "concat: column %d is covered by no field", c);
    }
    int v = pick_vec(fields, F, n_dense, 0);
    // vector loads also need 4*v-byte aligned bases
    auto aligned = [&](const void* p) { return ((uintptr_t)p % (4 * v)) == 0; };
    while (v > 1) {
        bool ok = aligned(dense);
        for (int f = 0; f < F; ++f) ok = ok && aligned(fields[f].weight);
        if (ok) break;
        v >>= 1;
    }
    a->dense   = dense;
    a->n_dense = n_dense;
    a->d       = d;
    a->n_units = d / v;
    *vec       = v;
    RK_CHECK_ARG(d > 0 && a->n_units <= 128, "concat: row of %d floats (vec %d) exceeds 128 units",
                 d, v);
    return 0;
}

static int warp_grid(int64_t B) {
    int64_t g = ceil_div(B, 8);  // 8 warps per CTA, one sample per warp per trip
    const int64_t cap = (int64_t)sm_count() * 8;
    if (g > cap) g = cap;
    return g > 0 ? (int)g : 1;
}

#define RK_DISPATCH_VEC_R(vec, rounds, CALL)                                        \
    do {                                                                            \
        if (vec == 4) {                                                             \
            if (rounds <= 1) { CALL(4, 1); } else if (rounds <= 2) { CALL(4, 2); }  \
            else { CALL(4, 4); }                                                    \
        } else if (vec == 2) {                                                      \
            if (rounds <= 1) { CALL(2, 1); } else if (rounds <= 2) { CALL(2, 2); }  \
            else { CALL(2, 4); }                                                    \
        } else {                                                                    \
            if (rounds <= 1) { CALL(1, 1); } else if (rounds <= 2) { CALL(1, 2); }  \
            else { CALL(1, 4); }                                                    \
        }                                                                           \
    } while (0)

}  // namespace rk

extern "C" {

int rk_gather_concat_fwd(const rk_field_t* fields, int F, const float* dense, int n_dense,
                         int64_t B, float* out, int ld_out, int32_t* err_flag,
                         rk_stream_t stream_) {
    using namespace rk;
    cudaStream_t s = (cudaStream_t)stream_;
    ConcatArgs a;
    int vec = 1;
    if (int rc = fill_concat_args(fields, F, dense, n_dense, &a, &vec)) return rc;
    RK_CHECK_ARG(B >= 0 && out && ld_out >= a.d, "gather_concat: B=%lld ld_out=%d d=%d",
                 (long long)B, ld_out, a.d);
    if (B == 0) return 0;
    while (vec > 1 && ((ld_out % vec) != 0 || ((uintptr_t)out % (4 * vec)) != 0)) {
        vec >>= 1;
        a.n_units = a.d / vec;
    }
    RK_CHECK_ARG(a.n_units <= 128, "gather_concat: row too wide");
    const int rounds = (a.n_units + 31) / 32;
    const int grid   = warp_grid(B);
#define CALL(V, R) gather_concat_kernel<V, R><<<grid, 256, 0, s>>>(a, B, out, ld_out, err_flag)
    RK_DISPATCH_VEC_R(vec, rounds, CALL);
#undef CALL
    RK_LAUNCH_CHECK();
    return 0;
}

int rk_crossnet_fwd(const rk_field_t* fields, int F, const float* dense, int n_dense,
                    const float* w, const float* b, int L, int64_t B, float* concat_all,
                    float* cross_vec, int32_t* err_flag, rk_stream_t stream_) {
    using namespace rk;
    cudaStream_t s = (cudaStream_t)stream_;
    ConcatArgs a;
    int vec = 1;
    if (int rc = fill_concat_args(fields, F, dense, n_dense, &a, &vec)) return rc;
    RK_CHECK_ARG(L >= 0 && L <= RK_MAX_LAYERS, "crossnet: L=%d outside [0,%d]", L, RK_MAX_LAYERS);
    RK_CHECK_ARG(B >= 0 && concat_all && cross_vec && (L == 0 || (w && b)),
                 "crossnet: NULL pointer");
    if (B == 0) return 0;
    while (vec > 1 && (((uintptr_t)concat_all % (4 * vec)) || ((uintptr_t)cross_vec % (4 * vec)))) {
        vec >>= 1;
        a.n_units = a.d / vec;
    }
    RK_CHECK_ARG(a.n_units <= 128, "crossnet: row too wide");
    const int    rounds = (a.n_units + 31) / 32;
    const int    grid   = warp_grid(B);
    const size_t smem   = (size_t)2 * L * a.d * sizeof(float);
    RK_CHECK_ARG(smem <= 48 * 1024, "crossnet: L*d too large for shared memory");
#define CALL(V, R) \
    crossnet_fwd_kernel<V, R><<<grid, 256, smem, s>>>(a, w, b, L, B, concat_all, cross_vec, err_flag)
    RK_DISPATCH_VEC_R(vec, rounds, CALL);
#undef CALL
    RK_LAUNCH_CHECK();
    return 0;
}

int rk_crossnet_bwd(const float* concat_all, const float* w, const float* b, int L, int d,
                    int64_t B, const float* g_concat_all, const float* g_cross_vec, float* g_x0,
                    rk_stream_t stream_) {
    using namespace rk;
    cudaStream_t s = (cudaStream_t)stream_;
    RK_CHECK_ARG(L >= 0 && L <= RK_MAX_LAYERS, "crossnet_bwd: L=%d", L);
    RK_CHECK_ARG(d > 0 && B >= 0 && concat_all && g_x0 && (L == 0 || (w && b)),
                 "crossnet_bwd: bad argument");
    if (B == 0) return 0;
    int vec = d % 4 == 0 ? 4 : (d % 2 == 0 ? 2 : 1);
    auto mis = [&](const void* p) { return p && ((uintptr_t)p % (4 * vec)) != 0; };
    while (vec > 1 && (mis(concat_all) || mis(g_concat_all) || mis(g_cross_vec) || mis(g_x0)))
        vec >>= 1;
    const int n_units = d / vec;
    RK_CHECK_ARG(n_units <= 128, "crossnet_bwd: row too wide");
    const int    rounds = (n_units + 31) / 32;
    const int    grid   = warp_grid(B);
    const size_t smem   = ((size_t)2 * L * d + 8 * RK_MAX_LAYERS) * sizeof(float);
    RK_CHECK_ARG(smem <= 48 * 1024, "crossnet_bwd: L*d too large for shared memory");
#define CALL(V, R)                                                                          \
    crossnet_bwd_kernel<V, R><<<grid, 256, smem, s>>>(concat_all, w, b, L, d, B, g_concat_all, \
                                                      g_cross_vec, g_x0)
    RK_DISPATCH_VEC_R(vec, rounds, CALL);
#undef CALL
    RK_LAUNCH_CHECK();
    return 0;
}

int rk_cross_layer_fwd(const float* x0, const float* xl, const float* w, const float* b, int d,
                       int64_t B, float* out, rk_stream_t stream_) {
    using namespace rk;
    RK_CHECK_ARG(x0 && xl && w && b && out && d > 0 && B >= 0, "cross_layer_fwd: bad argument");
    if (B == 0) return 0;
    cross_layer_fwd_kernel<<<warp_grid(B), 256, 0, (cudaStream_t)stream_>>>(x0, xl, w, b, d, B, out);
    RK_LAUNCH_CHECK();
    return 0;
}

int rk_cross_layer_bwd(const float* x0, const float* xl, const float* w, int d, int64_t B,
                       const float* g_out, float* g_x0, float* g_xl, rk_stream_t stream_) {
    using namespace rk;
    RK_CHECK_ARG(x0 && xl && w && g_out && g_x0 && g_xl && d > 0 && B >= 0,
                 "cross_layer_bwd: bad argument");
    if (B == 0) return 0;
    cross_layer_bwd_kernel<<<warp_grid(B), 256, 0, (cudaStream_t)stream_>>>(x0, xl, w, d, B, g_out,
                                                                            g_x0, g_xl);
    RK_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
