// afm.cu — AFM's pairwise-interaction attention pooling (AFM/afm.py:92-115, attention net
// :84-88), fused with the F embedding gathers, forward and backward, fp32 SIMT.
//   v_p = e_i * e_j (i<j, row-major pair order)      s_p = w2 . relu(W1 v_p + b1) + b2
//   a = softmax_p(s)                                  out = sum_p a_p v_p          [B, D]
// The reference materialises [B,P,D] and [B,P,A] in HBM (45 mul launches + 2 GEMMs); here one
// CTA keeps a tile of S samples x P pairs (<= 128 rows) in shared memory, feature-major, runs
// the attention MLP with tile_gemm.cuh and never stores the hidden layer.  The attention
// weights are REGISTERED parameters (attention.0/2.{weight,bias}), so the backward also
// produces dW1, db1, dw2, db2: each persistent CTA accumulates its share in registers over its
// tiles and a second kernel adds the per-CTA partials in a fixed order (no atomics).
#include <string.h>
#include "common.cuh"
#include "tile_gemm.cuh"

namespace rk {

constexpr int kAfmThreads = 256;
constexpr int kAfmRows    = 128;
constexpr int kAfmMaxS    = 16;
// Row stride of the feature-major tiles (AfmParams::ld): the smallest value >= the tile's rows
// that is a multiple of 4 and = 4 mod 32, so that both the m-contiguous float4 loads and the
// loads of consecutive features at a fixed m are conflict free.

struct AfmParams {
    FieldSet     fs;
    const float* w1;   // [A][D]  attention.0.weight
    const float* b1;   // [A]     attention.0.bias
    const float* w2;   // [A]     attention.2.weight ([1][A])
    const float* b2;   // [1]     attention.2.bias
    int32_t      D, A, Ap, P, S, ld;
    int64_t      B, n_tiles;
};

struct AfmSmem {
    float *w1t, *w1, *b1, *w2;     // [D][Ap], [Ap][D], [Ap], [Ap]
    float *e, *x, *h, *part;       // [S][F][D], [D][ld], [Ap][ld] (bwd), [Ap/8][ld]
    float *score, *attn, *ds, *gout;
    int   *pi, *pj, *pidx;         // pair -> (i, j); (f, g) -> pair
    __device__ AfmSmem(float* base, const AfmParams& p, bool bwd) {
        float* q = base;
        w1t = q;   q += p.D * p.Ap;
        w1 = q;    q += bwd ? p.Ap * p.D : 0;
        b1 = q;    q += p.Ap;
        w2 = q;    q += p.Ap;
        e = q;     q += p.S * p.fs.F * p.D;
        x = q;     q += p.D * p.ld;
        h = q;     q += bwd ? p.Ap * p.ld : 0;
        part = q;  q += (p.Ap / 8) * p.ld;
        score = q; q += p.ld;
        attn = q;  q += p.ld;
        ds = q;    q += p.ld;
        gout = q;  q += p.S * p.D;
        pi = (int*)q;   q += p.P;
        pj = (int*)q;   q += p.P;
        pidx = (int*)q; q += p.fs.F * p.fs.F;
    }
    static size_t bytes(const AfmParams& p, bool bwd) {
        size_t n = (size_t)p.D * p.Ap + (bwd ? (size_t)p.Ap * p.D : 0) + 2 * p.Ap + (size_t)p.S * p.fs.F * p.D +
                   (size_t)p.D * p.ld + (bwd ? (size_t)p.Ap * p.ld : 0) + (size_t)(p.Ap / 8) * p.ld +
                   3 * p.ld + (size_t)p.S * p.D + 2 * p.P + (size_t)p.fs.F * p.fs.F;
        return n * sizeof(float);
    }
};

__device__ __forceinline__ void afm_stage_weights(const AfmParams& p, const AfmSmem& sm, bool bwd) {
    const int tid = threadIdx.x;
    for (int i = tid; i < p.D * p.Ap; i += kAfmThreads) {        // w1t[d][n] = W1[n][d]
        const int d = i / p.Ap, n = i - d * p.Ap;
        sm.w1t[i] = n < p.A ? __ldg(p.w1 + n * p.D + d) : 0.f;
    }
    if (bwd)
        for (int i = tid; i < p.Ap * p.D; i += kAfmThreads) sm.w1[i] = i < p.A * p.D ? __ldg(p.w1 + i) : 0.f;
    for (int i = tid; i < p.Ap; i += kAfmThreads) {
        sm.b1[i] = i < p.A ? __ldg(p.b1 + i) : 0.f;
        sm.w2[i] = i < p.A ? __ldg(p.w2 + i) : 0.f;
    }
    const int F = p.fs.F;
    if (tid == 0) {
        int q = 0;
        for (int i = 0; i < F; ++i)
            for (int j = i + 1; j < F; ++j) {
                sm.pi[q] = i; sm.pj[q] = j;
                sm.pidx[i * F + j] = q; sm.pidx[j * F + i] = q;
                ++q;
            }
    }
}

// Gather the tile's embedding rows and build the pair products, feature-major.
__device__ __forceinline__ void afm_build_tile(const AfmParams& p, const AfmSmem& sm, int64_t b0, int ns,
                                               int m_used, int32_t* err_flag) {
    const int tid = threadIdx.x, F = p.fs.F, D = p.D, D4 = D / 4;
    for (int item = tid; item < ns * F * D4; item += kAfmThreads) {
        const int c4 = item % D4, sf = item / D4, f = sf % F, s = sf / F;
        const int64_t row = checked_row(__ldg(p.fs.idx[f] + b0 + s), p.fs.rows[f], err_flag);
        *reinterpret_cast<float4*>(sm.e + sf * D + c4 * 4) =
            __ldg(reinterpret_cast<const float4*>(p.fs.weight[f] + row * D) + c4);
    }
    __syncthreads();
    const int rows = ns * p.P;
    for (int item = tid; item < D * m_used; item += kAfmThreads) {
        const int d = item / m_used, m = item - d * m_used;
        float v = 0.f;
        if (m < rows) {
            const int s = m / p.P, q = m - s * p.P;
            v = sm.e[(s * F + sm.pi[q]) * D + d] * sm.e[(s * F + sm.pj[q]) * D + d];
        }
        sm.x[d * p.ld + m] = v;
    }
    __syncthreads();
}

// Hidden layer + scores.  KEEP_H also leaves relu(W1 v + b1) in sm.h (feature-major).
template <bool KEEP_H>
__device__ __forceinline__ void afm_scores(const AfmParams& p, const AfmSmem& sm, int m_used, int rows) {
    tile_gemm<8, kAfmThreads>(sm.x, p.ld, sm.w1t, p.Ap, p.D, p.Ap, m_used,
                              [&](int m0, int n0, float (&acc)[4][8]) {
        float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float bb = sm.b1[n0 + j], ww = sm.w2[n0 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i][j] = fmaxf(acc[i][j] + bb, 0.f);
                part[i] = fmaf(acc[i][j], ww, part[i]);
            }
        }
        if (KEEP_H) store_tile_kmajor<8>(sm.h, p.ld, m0, n0, acc);
        *reinterpret_cast<float4*>(sm.part + (n0 >> 3) * p.ld + m0) = make_float4(part[0], part[1], part[2], part[3]);
    });
    __syncthreads();
    const float b2 = __ldg(p.b2);
    for (int m = threadIdx.x; m < m_used; m += kAfmThreads) {
        float s = b2;
        for (int t = 0; t < p.Ap / 8; ++t) s += sm.part[t * p.ld + m];
        sm.score[m] = s;
        if (m >= rows) sm.attn[m] = 0.f;
    }
    __syncthreads();
    // softmax over each sample's P pairs: one warp per sample
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int s = warp; s * p.P < rows; s += kAfmThreads / 32) {
        const int r0 = s * p.P;
        float mx = -INFINITY;
        for (int q = lane; q < p.P; q += 32) mx = fmaxf(mx, sm.score[r0 + q]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int q = lane; q < p.P; q += 32) sum += expf(sm.score[r0 + q] - mx);
        const float inv = 1.0f / warp_sum(sum);
        for (int q = lane; q < p.P; q += 32) sm.attn[r0 + q] = expf(sm.score[r0 + q] - mx) * inv;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kAfmThreads, 4)
afm_fwd_kernel(const __grid_constant__ AfmParams p, float* __restrict__ out, int32_t* err_flag) {
    extern __shared__ __align__(16) float smem_raw[];
    AfmSmem sm(smem_raw, p, false);
    afm_stage_weights(p, sm, false);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, D = p.D;
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int64_t b0 = tile * p.S;
        const int ns = (int)((p.B - b0) < p.S ? (p.B - b0) : p.S);
        const int rows = ns * p.P, m_used = (rows + 3) / 4 * 4;
        afm_build_tile(p, sm, b0, ns, m_used, err_flag);
        afm_scores<false>(p, sm, m_used, rows);
        // weighted sum over the pairs: every warp takes (sample, 4 columns), lanes over the pairs
        for (int item = warp; item < ns * (D / 4); item += kAfmThreads / 32) {
            const int s = item / (D / 4), d0 = (item - s * (D / 4)) * 4;
            const int r0 = s * p.P;
            float a[4] = {0.f, 0.f, 0.f, 0.f};
            for (int q = lane; q < p.P; q += 32) {
                const float w = sm.attn[r0 + q];
#pragma unroll
                for (int j = 0; j < 4; ++j) a[j] = fmaf(w, sm.x[(d0 + j) * p.ld + r0 + q], a[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) a[j] = warp_sum(a[j]);
            if (lane == 0)
                *reinterpret_cast<float4*>(out + (b0 + s) * D + d0) = make_float4(a[0], a[1], a[2], a[3]);
        }
        __syncthreads();
    }
}

// partial layout per CTA: [dW1 A*D][db1 A][dw2 A][db2 1]
__global__ void __launch_bounds__(kAfmThreads, 2)
afm_bwd_kernel(const __grid_constant__ AfmParams p, const float* __restrict__ g_out,
               float* __restrict__ g_rows, float* __restrict__ partials, int32_t* err_flag) {
    extern __shared__ __align__(16) float smem_raw[];
    AfmSmem sm(smem_raw, p, true);
    afm_stage_weights(p, sm, true);
    __syncthreads();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = p.D, F = p.fs.F, Ap = p.Ap, NI = Ap / 32;
    // persistent accumulators: dW1[n = lane + 32 i][d = 4 warp + j], db1/dw2[n = tid]
    float accW[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) accW[i][j] = 0.f;
    float acc_b1 = 0.f, acc_w2 = 0.f, acc_b2 = 0.f;
    const bool w_on = warp < D / 4;

    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int64_t b0 = tile * p.S;
        const int ns = (int)((p.B - b0) < p.S ? (p.B - b0) : p.S);
        const int rows = ns * p.P, m_used = (rows + 3) / 4 * 4;
        for (int i = tid; i < ns * D; i += kAfmThreads) sm.gout[i] = g_out[b0 * D + i];
        afm_build_tile(p, sm, b0, ns, m_used, err_flag);
        afm_scores<true>(p, sm, m_used, rows);
        // d a_p = g_out . v_p ; d s_p = a_p (d a_p - sum_q a_q d a_q)
        for (int m = tid; m < m_used; m += kAfmThreads) {
            float a = 0.f;
            if (m < rows) {
                const int s = m / p.P;
                for (int d = 0; d < D; ++d) a = fmaf(sm.gout[s * D + d], sm.x[d * p.ld + m], a);
            }
            sm.ds[m] = a;
        }
        __syncthreads();
        for (int s = warp; s < ns; s += kAfmThreads / 32) {
            const int r0 = s * p.P;
            float dot = 0.f;
            for (int q = lane; q < p.P; q += 32) dot = fmaf(sm.attn[r0 + q], sm.ds[r0 + q], dot);
            dot = warp_sum(dot);
            for (int q = lane; q < p.P; q += 32) sm.ds[r0 + q] = sm.attn[r0 + q] * (sm.ds[r0 + q] - dot);
        }
        __syncthreads();
        // dw2[n] += sum_m ds[m] h[n][m];  db2 += sum_m ds[m];  then h <- dz = ds * w2[n] * (h > 0)
        if (tid < Ap) {
            const float* hrow = sm.h + tid * p.ld;
            float a = 0.f;
            for (int m = 0; m < m_used; m += 4) {
                const float4 hv = *reinterpret_cast<const float4*>(hrow + m);
                const float4 dv = *reinterpret_cast<const float4*>(sm.ds + m);
                a = fmaf(hv.x, dv.x, a); a = fmaf(hv.y, dv.y, a); a = fmaf(hv.z, dv.z, a); a = fmaf(hv.w, dv.w, a);
            }
            acc_w2 += a;
        } else if (warp == kAfmThreads / 32 - 1) {
            float a = 0.f;
            for (int m = lane; m < m_used; m += 32) a += sm.ds[m];
            a = warp_sum(a);
            acc_b2 += a;
        }
        __syncthreads();
        for (int item = tid; item < Ap * (m_used / 4); item += kAfmThreads) {
            const int n = item / (m_used / 4), m = (item - n * (m_used / 4)) * 4;
            float4 hv = *reinterpret_cast<float4*>(sm.h + n * p.ld + m);
            const float4 dv = *reinterpret_cast<const float4*>(sm.ds + m);
            const float ww = sm.w2[n];
            hv.x = hv.x > 0.f ? dv.x * ww : 0.f;
            hv.y = hv.y > 0.f ? dv.y * ww : 0.f;
            hv.z = hv.z > 0.f ? dv.z * ww : 0.f;
            hv.w = hv.w > 0.f ? dv.w * ww : 0.f;
            *reinterpret_cast<float4*>(sm.h + n * p.ld + m) = hv;
        }
        __syncthreads();
        // db1[n] += sum_m dz[n][m]
        if (tid < Ap) {
            const float* zrow = sm.h + tid * p.ld;
            float a = 0.f;
            for (int m = 0; m < m_used; m += 4) {
                const float4 z = *reinterpret_cast<const float4*>(zrow + m);
                a += (z.x + z.y) + (z.z + z.w);
            }
            acc_b1 += a;
        }
        // dW1[n][d] += sum_m dz[n][m] v[d][m]: lanes own consecutive n (stride p.ld = 4 mod 32 ->
        // conflict-free float4 loads along m), the warp's four v rows are broadcast
        if (w_on) {
            const float* xr = sm.x + (4 * warp) * p.ld;
            for (int m = 0; m < m_used; m += 4) {
                float4 xv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) xv[j] = *reinterpret_cast<const float4*>(xr + j * p.ld + m);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (i < NI) {
                        const float4 z = *reinterpret_cast<const float4*>(sm.h + (lane + 32 * i) * p.ld + m);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            accW[i][j] = fmaf(z.x, xv[j].x, accW[i][j]);
                            accW[i][j] = fmaf(z.y, xv[j].y, accW[i][j]);
                            accW[i][j] = fmaf(z.z, xv[j].z, accW[i][j]);
                            accW[i][j] = fmaf(z.w, xv[j].w, accW[i][j]);
                        }
                    }
                }
            }
        }
        __syncthreads();
        // dv[m][d] = a_m g_out[s][d] + sum_n dz[n][m] W1[n][d]  -> overwrite x (feature-major)
        tile_gemm<4, kAfmThreads>(sm.h, p.ld, sm.w1, D, Ap, D, m_used, [&](int m0, int n0, float (&acc)[4][4]) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = m0 + i;
                const int s = m < rows ? m / p.P : 0;
                const float a = sm.attn[m];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a, sm.gout[s * D + n0 + j], acc[i][j]);
            }
            store_tile_kmajor<4>(sm.x, p.ld, m0, n0, acc);
        });
        __syncthreads();
        // d e_f = sum_{g != f} dv_{pair(f,g)} * e_g
        for (int item = tid; item < ns * F * D; item += kAfmThreads) {
            const int d = item % D, sf = item / D, f = sf % F, s = sf / F;
            float a = 0.f;
            for (int g = 0; g < F; ++g) {
                if (g == f) continue;
                a = fmaf(sm.x[d * p.ld + s * p.P + sm.pidx[f * F + g]], sm.e[(s * F + g) * D + d], a);
            }
            g_rows[(b0 + s) * F * D + f * D + d] = a;
        }
        __syncthreads();
    }
    // per-CTA partial gradients of the registered attention weights
    float* out = partials + (int64_t)blockIdx.x * (p.A * D + 2 * p.A + 1);
    if (w_on) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int n = lane + 32 * i;
            if (i < NI && n < p.A) {
#pragma unroll
                for (int j = 0; j < 4; ++j) out[n * D + 4 * warp + j] = accW[i][j];
            }
        }
    }
    if (tid < p.A) {
        out[p.A * D + tid]       = acc_b1;
        out[p.A * D + p.A + tid] = acc_w2;
    }
    if (tid == kAfmThreads - 32) out[p.A * D + 2 * p.A] = acc_b2;
}

static int afm_fill(const rk_field_t* fields, int F, const float* w1, const float* b1, const float* w2,
                    const float* b2, int A, int64_t B, AfmParams* p) {
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
    p->D = D; p->A = A; p->Ap = (A + 31) / 32 * 32;
    p->P = F * (F - 1) / 2;
    int S = kAfmRows / p->P;
    p->S = S > kAfmMaxS ? kAfmMaxS : S;
    {
        const int rows = (p->S * p->P + 3) / 4 * 4;
        int ld = 4;
        while (ld < rows) ld += 32;      // 4, 36, 68, 100, 132: multiple of 4 and = 4 mod 32
        p->ld = ld;
    }
    p->B = B;
    p->n_tiles = ceil_div(B, p->S);
    return 0;
}

}  // namespace rk

extern "C" {

int rk_afm_bwd_ctas(int64_t B, int F) {
    if (F < 2 || F > 16 || B <= 0) return 1;
    int S = rk::kAfmRows / (F * (F - 1) / 2);
    if (S > rk::kAfmMaxS) S = rk::kAfmMaxS;
    int64_t tiles = rk::ceil_div(B, S);
    int64_t cap = 2 * (int64_t)rk::sm_count();   // two CTAs fit per SM (shared memory and registers)
    return (int)(tiles < cap ? tiles : cap);
}

int rk_afm_fwd(const rk_field_t* fields, int F, const float* w1, const float* b1, const float* w2,
               const float* b2, int A, int64_t B, float* out, int32_t* err_flag, rk_stream_t stream_) {
    using namespace rk;
    AfmParams p;
    if (int rc = afm_fill(fields, F, w1, b1, w2, b2, A, B, &p)) return rc;
    RK_CHECK_ARG(out, "afm_fwd: out is NULL");
    if (B == 0) return 0;
    const size_t smem = AfmSmem::bytes(p, false);
    RK_CHECK_ARG(smem <= 227 * 1024, "afm_fwd: %zu bytes of shared memory", smem);
    RK_CUDA(cudaFuncSetAttribute(afm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = p.n_tiles;
    const int64_t cap = (int64_t)sm_count() * 4;
    if (grid > cap) grid = cap;
    afm_fwd_kernel<<<(int)grid, kAfmThreads, smem, (cudaStream_t)stream_>>>(p, out, err_flag);
    RK_LAUNCH_CHECK();
    return 0;
}

int rk_afm_bwd(const rk_field_t* fields, int F, const float* w1, const float* b1, const float* w2,
               const float* b2, int A, int64_t B, const float* g_out, float* g_rows, float* g_w1,
               float* g_b1, float* g_w2, float* g_b2, float* partials, int n_ctas, int32_t* err_flag,
               rk_stream_t stream_) {
    using namespace rk;
    AfmParams p;
    if (int rc = afm_fill(fields, F, w1, b1, w2, b2, A, B, &p)) return rc;
    RK_CHECK_ARG(g_out && g_rows && g_w1 && g_b1 && g_w2 && g_b2 && partials, "afm_bwd: NULL pointer");
    RK_CHECK_ARG(n_ctas == rk_afm_bwd_ctas(B, F), "afm_bwd: n_ctas %d != rk_afm_bwd_ctas", n_ctas);
    RK_CHECK_ARG(g_b1 == g_w1 + (size_t)A * p.D && g_w2 == g_b1 + A && g_b2 == g_w2 + A,
                 "afm_bwd: g_w1|g_b1|g_w2|g_b2 must be one contiguous buffer in that order");
    if (B == 0) return 0;
    const size_t smem = AfmSmem::bytes(p, true);
    RK_CHECK_ARG(smem <= 227 * 1024, "afm_bwd: %zu bytes of shared memory", smem);
    RK_CUDA(cudaFuncSetAttribute(afm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaStream_t s = (cudaStream_t)stream_;
    afm_bwd_kernel<<<n_ctas, kAfmThreads, smem, s>>>(p, g_out, g_rows, partials, err_flag);
    RK_LAUNCH_CHECK();
    const int count = A * p.D + 2 * A + 1;
    return launch_reduce_partials(partials, n_ctas, count, g_w1, s);
}

}  // extern "C"
