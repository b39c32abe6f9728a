// tower.cu — fused activation + normalisation layers of the DNN towers.
//
//   Dice (DIN/din.py:26-36) followed by the tower's BatchNorm1d (DIN/din.py:272-285), training mode:
//     xh = (x - mean_B x) / sqrt(var_B x + eps1)           nn.BatchNorm1d(units, affine=False)
//     p  = sigmoid(xh);   y = alpha * (1 - p) * x + p * x
//     z  = gamma * (y - mean_B y) / sqrt(var_B y + eps2) + beta        (optional second BatchNorm1d)
//   torch runs this as ~25 launches per layer (two batch-norm stat passes, a dozen elementwise
//   kernels, single-CTA sums for d alpha); here it is one kernel per direction.
//
// A CTA owns COLS adjacent feature columns and the whole batch: thread t keeps rows t, t + 1024, ...
// of its columns in registers (<= 16 rows per thread, i.e. B <= 16384), so x is read once and z
// written once; every batch statistic is a block reduction with a fixed order (warp shuffle tree,
// then the 32 warp sums in order): results do not depend on scheduling.  Running statistics are
// updated in place as nn.BatchNorm1d does (momentum, unbiased variance, num_batches_tracked).
#include "common.cuh"

namespace rk {

constexpr int kTowerThreads = 1024;
constexpr int kTowerMaxRows = 16;

// block-wide sums of N values per thread; every thread gets the totals.  The tree and the cross-warp
// pass run in double: gradients of biases / BN shifts are sums with heavy cancellation, and a float
// tree would add ~1e-7 of sum|terms| to a result that can be 100x smaller.
template <int N>
__device__ __forceinline__ void block_sum(float (&v)[N], double* scratch /* [33][N] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double d[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        d[i] = (double)v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d[i] += __shfl_xor_sync(kFull, d[i], o);
    }
    __syncthreads();                                   // scratch may still be read from the previous call
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) scratch[warp * N + i] = d[i];
    }
    __syncthreads();
    if (threadIdx.x < N) {                             // one thread per value adds the 32 warp sums in order
        double t = 0.0;
#pragma unroll 8
        for (int w = 0; w < kTowerThreads / 32; ++w) t += scratch[w * N + threadIdx.x];
        scratch[32 * N + threadIdx.x] = t;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = (float)scratch[32 * N + i];
}

struct DiceBnArgs {
    const float* x;          // [B, units]
    const float* alpha;      // [units]
    const float* gamma;      // [units] or NULL (no second batch norm)
    const float* beta;       // [units] or NULL
    float*   rm1;  float* rv1;  int64_t* nbt1;     // running stats of Dice.bn (may be NULL)
    float*   rm2;  float* rv2;  int64_t* nbt2;     // running stats of the second BN (may be NULL)
    float    eps1, eps2, mom1, mom2;
    int64_t  B;
    int32_t  units;
};

template <int COLS, int ROWS>
__global__ void __launch_bounds__(kTowerThreads, 1)
dice_bn_fwd_kernel(const DiceBnArgs a, float* __restrict__ z, float* __restrict__ stats) {
    __shared__ double scratch[33 * COLS];
    const int c0 = blockIdx.x * COLS, tid = threadIdx.x;
    const int64_t B = a.B;
    const int U = a.units;
    const float invB = 1.0f / (float)B;
    Vec<COLS> x[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int64_t row = tid + (int64_t)r * kTowerThreads;
        if (row < B) x[r].load(a.x + row * U + c0); else vec_zero(x[r]);
    }
    float mu1[COLS], rstd1[COLS], mu2[COLS], rstd2[COLS], al[COLS];
    {   // batch statistics of x: mean, then centred second moment (two-pass, as ATen)
        float s[COLS];
#pragma unroll
        for (int c = 0; c < COLS; ++c) s[c] = 0.f;
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
            for (int c = 0; c < COLS; ++c) s[c] += x[r].v[c];
        block_sum<COLS>(s, scratch);
#pragma unroll
        for (int c = 0; c < COLS; ++c) { mu1[c] = s[c] * invB; s[c] = 0.f; }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const bool live = tid + (int64_t)r * kTowerThreads < B;
#pragma unroll
            for (int c = 0; c < COLS; ++c) {
                const float d = x[r].v[c] - mu1[c];
                s[c] += live ? d * d : 0.f;
            }
        }
        block_sum<COLS>(s, scratch);
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
            const float var = s[c] * invB;
            rstd1[c] = 1.0f / sqrtf(var + a.eps1);
            al[c] = __ldg(a.alpha + c0 + c);
            if (tid == 0 && a.rm1) {
                a.rm1[c0 + c] = (1.f - a.mom1) * a.rm1[c0 + c] + a.mom1 * mu1[c];
                a.rv1[c0 + c] = (1.f - a.mom1) * a.rv1[c0 + c] + a.mom1 * (B > 1 ? var * (float)B / (float)(B - 1) : var);
            }
        }
    }
    // y = x * (alpha + (1 - alpha) p), in place
    float s[COLS];
#pragma unroll
    for (int c = 0; c < COLS; ++c) s[c] = 0.f;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const bool live = tid + (int64_t)r * kTowerThreads < B;
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
            const float xv = x[r].v[c];
            const float p = 1.0f / (1.0f + expf(-(xv - mu1[c]) * rstd1[c]));
            const float y = al[c] * (1.0f - p) * xv + p * xv;
            x[r].v[c] = live ? y : 0.f;
            s[c] += x[r].v[c];
        }
    }
    if (a.gamma) {
        block_sum<COLS>(s, scratch);
#pragma unroll
        for (int c = 0; c < COLS; ++c) { mu2[c] = s[c] * invB; s[c] = 0.f; }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const bool live = tid + (int64_t)r * kTowerThreads < B;
#pragma unroll
            for (int c = 0; c < COLS; ++c) {
                const float d = x[r].v[c] - mu2[c];
                s[c] += live ? d * d : 0.f;
            }
        }
        block_sum<COLS>(s, scratch);
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
            const float var = s[c] * invB;
            rstd2[c] = 1.0f / sqrtf(var + a.eps2);
            if (tid == 0 && a.rm2) {
                a.rm2[c0 + c] = (1.f - a.mom2) * a.rm2[c0 + c] + a.mom2 * mu2[c];
                a.rv2[c0 + c] = (1.f - a.mom2) * a.rv2[c0 + c] + a.mom2 * (B > 1 ? var * (float)B / (float)(B - 1) : var);
            }
        }
    } else {
#pragma unroll
        for (int c = 0; c < COLS; ++c) { mu2[c] = 0.f; rstd2[c] = 1.f; }
    }
    if (tid == 0) {
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
            stats[0 * U + c0 + c] = mu1[c];  stats[1 * U + c0 + c] = rstd1[c];
            stats[2 * U + c0 + c] = mu2[c];  stats[3 * U + c0 + c] = rstd2[c];
        }
        if (blockIdx.x == 0) {
            if (a.nbt1) *a.nbt1 += 1;
            if (a.nbt2 && a.gamma) *a.nbt2 += 1;
        }
    }
    float g[COLS], b[COLS];
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
        g[c] = a.gamma ? __ldg(a.gamma + c0 + c) * rstd2[c] : 1.f;
        b[c] = a.gamma ? __ldg(a.beta + c0 + c) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int64_t row = tid + (int64_t)r * kTowerThreads;
        if (row < B) {
            Vec<COLS> o;
#pragma unroll
            for (int c = 0; c < COLS; ++c) o.v[c] = fmaf(x[r].v[c] - mu2[c], g[c], b[c]);
            o.store(z + row * U + c0);
        }
    }
}

// g_x from g_z; d alpha, d gamma, d beta are complete per column (the CTA sees the whole batch)
template <int COLS, int ROWS>
__global__ void __launch_bounds__(kTowerThreads, 1)
dice_bn_bwd_kernel(const DiceBnArgs a, const float* __restrict__ g_z, const float* __restrict__ stats,
                   float* __restrict__ g_x, float* __restrict__ g_alpha, float* __restrict__ g_gamma,
                   float* __restrict__ g_beta) {
    __shared__ double scratch[33 * 3 * COLS];
    const int c0 = blockIdx.x * COLS, tid = threadIdx.x;
    const int64_t B = a.B;
    const int U = a.units;
    const float invB = 1.0f / (float)B;
    const bool bn2 = a.gamma != nullptr;
    float mu1[COLS], rstd1[COLS], mu2[COLS], rstd2[COLS], al[COLS], gam[COLS];
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
        mu1[c] = stats[0 * U + c0 + c];  rstd1[c] = stats[1 * U + c0 + c];
        mu2[c] = stats[2 * U + c0 + c];  rstd2[c] = stats[3 * U + c0 + c];
        al[c] = __ldg(a.alpha + c0 + c);
        gam[c] = bn2 ? __ldg(a.gamma + c0 + c) : 1.f;
    }
    Vec<COLS> x[ROWS], g[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int64_t row = tid + (int64_t)r * kTowerThreads;
        if (row < B) { x[r].load(a.x + row * U + c0); g[r].load(g_z + row * U + c0); }
        else { vec_zero(x[r]); vec_zero(g[r]); }
    }
    // second batch norm: g_y = gamma rstd2 (g_z - mean g_z - yh mean(g_z yh))
    if (bn2) {
        float s[2 * COLS];
#pragma unroll
        for (int i = 0; i < 2 * COLS; ++i) s[i] = 0.f;
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
            for (int c = 0; c < COLS; ++c) {
                const float xv = x[r].v[c];
                const float p = 1.0f / (1.0f + expf(-(xv - mu1[c]) * rstd1[c]));
                const float yh = (al[c] * (1.0f - p) * xv + p * xv - mu2[c]) * rstd2[c];
                s[c] += g[r].v[c];                       // rows >= B hold g = 0
                s[COLS + c] += g[r].v[c] * yh;
            }
        block_sum<2 * COLS>(s, scratch);
        if (tid == 0) {
#pragma unroll
            for (int c = 0; c < COLS; ++c) { g_beta[c0 + c] = s[c]; g_gamma[c0 + c] = s[COLS + c]; }
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
            for (int c = 0; c < COLS; ++c) {
                const float xv = x[r].v[c];
                const float p = 1.0f / (1.0f + expf(-(xv - mu1[c]) * rstd1[c]));
                const float yh = (al[c] * (1.0f - p) * xv + p * xv - mu2[c]) * rstd2[c];
                g[r].v[c] = gam[c] * rstd2[c] * (g[r].v[c] - s[c] * invB - yh * s[COLS + c] * invB);
            }
    }
    // Dice: y = x (alpha + (1 - alpha) p), p = sigmoid(xh), xh = (x - mu1) rstd1 (batch norm without affine)
    float s[3 * COLS];
#pragma unroll
    for (int i = 0; i < 3 * COLS; ++i) s[i] = 0.f;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const bool live = tid + (int64_t)r * kTowerThreads < B;
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
            const float xv = x[r].v[c], gy = live ? g[r].v[c] : 0.f;
            const float xh = (xv - mu1[c]) * rstd1[c];
            const float p = 1.0f / (1.0f + expf(-xh));
            const float gxh = gy * xv * (1.0f - al[c]) * p * (1.0f - p);
            s[c] += gy * xv * (1.0f - p);                // d alpha
            s[COLS + c] += gxh;
            s[2 * COLS + c] += gxh * xh;
        }
    }
    block_sum<3 * COLS>(s, scratch);
    if (tid == 0) {
#pragma unroll
        for (int c = 0; c < COLS; ++c) g_alpha[c0 + c] = s[c];
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int64_t row = tid + (int64_t)r * kTowerThreads;
        if (row < B) {
            Vec<COLS> o;
#pragma unroll
            for (int c = 0; c < COLS; ++c) {
                const float xv = x[r].v[c], gy = g[r].v[c];
                const float xh = (xv - mu1[c]) * rstd1[c];
                const float p = 1.0f / (1.0f + expf(-xh));
                const float gxh = gy * xv * (1.0f - al[c]) * p * (1.0f - p);
                o.v[c] = gy * (al[c] + (1.0f - al[c]) * p) +
                         rstd1[c] * (gxh - s[COLS + c] * invB - xh * s[2 * COLS + c] * invB);
            }
            o.store(g_x + row * U + c0);
        }
    }
}

// ---- BatchNorm1d (affine, training) + ReLU / LeakyReLU: the DeepFM and BST towers
//      (DeepFM/deepfm.py:100-110, BST/bst.py:203-214):  z = act(gamma * xh + beta),
//      act(u) = u > 0 ? u : slope * u  (slope = 0: ReLU; 1: no activation).
struct BnActArgs {
    const float* x;
    const float* gamma;
    const float* beta;
    float*   rm;  float* rv;  int64_t* nbt;
    float    eps, mom, slope;
    int64_t  B;
    int32_t  units;
};

template <int COLS, int ROWS>
__global__ void __launch_bounds__(kTowerThreads, 1)
bn_act_fwd_kernel(const BnActArgs a, float* __restrict__ z, float* __restrict__ stats) {
    __shared__ double scratch[33 * COLS];
    const int c0 = blockIdx.x * COLS, tid = threadIdx.x;
    const int64_t B = a.B;
    const int U = a.units;
    const float invB = 1.0f / (float)B;
    Vec<COLS> x[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int64_t row = tid + (int64_t)r * kTowerThreads;
        if (row < B) x[r].load(a.x + row * U + c0); else vec_zero(x[r]);
    }
    float mu[COLS], rstd[COLS], s[COLS];
#pragma unroll
    for (int c = 0; c < COLS; ++c) s[c] = 0.f;
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int c = 0; c < COLS; ++c) s[c] += x[r].v[c];
    block_sum<COLS>(s, scratch);
#pragma unroll
    for (int c = 0; c < COLS; ++c) { mu[c] = s[c] * invB; s[c] = 0.f; }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const bool live = tid + (int64_t)r * kTowerThreads < B;
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
            const float d = x[r].v[c] - mu[c];
            s[c] += live ? d * d : 0.f;
        }
    }
    block_sum<COLS>(s, scratch);
    float g[COLS], b[COLS];
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
        const float var = s[c] * invB;
        rstd[c] = 1.0f / sqrtf(var + a.eps);
        g[c] = __ldg(a.gamma + c0 + c) * rstd[c];
        b[c] = __ldg(a.beta + c0 + c);
        if (tid == 0) {
            stats[0 * U + c0 + c] = mu[c];
            stats[1 * U + c0 + c] = rstd[c];
            if (a.rm) {
                a.rm[c0 + c] = (1.f - a.mom) * a.rm[c0 + c] + a.mom * mu[c];
                a.rv[c0 + c] = (1.f - a.mom) * a.rv[c0 + c] + a.mom * (B > 1 ? var * (float)B / (float)(B - 1) : var);
            }
        }
    }
    if (tid == 0 && blockIdx.x == 0 && a.nbt) *a.nbt += 1;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int64_t row = tid + (int64_t)r * kTowerThreads;
        if (row < B) {
            Vec<COLS> o;
#pragma unroll
            for (int c = 0; c < COLS; ++c) {
                const float u = fmaf(x[r].v[c] - mu[c], g[c], b[c]);
                o.v[c] = u > 0.f ? u : a.slope * u;
            }
            o.store(z + row * U + c0);
        }
    }
}

template <int COLS, int ROWS>
__global__ void __launch_bounds__(kTowerThreads, 1)
bn_act_bwd_kernel(const BnActArgs a, const float* __restrict__ g_z, const float* __restrict__ stats,
                  float* __restrict__ g_x, float* __restrict__ g_gamma, float* __restrict__ g_beta) {
    __shared__ double scratch[33 * 2 * COLS];
    const int c0 = blockIdx.x * COLS, tid = threadIdx.x;
    const int64_t B = a.B;
    const int U = a.units;
    const float invB = 1.0f / (float)B;
    float mu[COLS], rstd[COLS], gam[COLS], bet[COLS];
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
        mu[c] = stats[0 * U + c0 + c];  rstd[c] = stats[1 * U + c0 + c];
        gam[c] = __ldg(a.gamma + c0 + c);  bet[c] = __ldg(a.beta + c0 + c);
    }
    Vec<COLS> xh[ROWS], g[ROWS];
    float s[2 * COLS];
#pragma unroll
    for (int i = 0; i < 2 * COLS; ++i) s[i] = 0.f;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int64_t row = tid + (int64_t)r * kTowerThreads;
        if (row < B) { xh[r].load(a.x + row * U + c0); g[r].load(g_z + row * U + c0); }
        else { vec_zero(xh[r]); vec_zero(g[r]); }
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
            const float h = (xh[r].v[c] - mu[c]) * rstd[c];
            const float u = fmaf(h, gam[c], bet[c]);
            const float gu = row < B ? g[r].v[c] * (u > 0.f ? 1.0f : a.slope) : 0.f;
            xh[r].v[c] = h;
            g[r].v[c] = gu;
            s[c] += gu;
            s[COLS + c] += gu * h;
        }
    }
    block_sum<2 * COLS>(s, scratch);
    if (tid == 0) {
#pragma unroll
        for (int c = 0; c < COLS; ++c) { g_beta[c0 + c] = s[c]; g_gamma[c0 + c] = s[COLS + c]; }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int64_t row = tid + (int64_t)r * kTowerThreads;
        if (row < B) {
            Vec<COLS> o;
#pragma unroll
            for (int c = 0; c < COLS; ++c)
                o.v[c] = gam[c] * rstd[c] * (g[r].v[c] - s[c] * invB - xh[r].v[c] * s[COLS + c] * invB);
            o.store(g_x + row * U + c0);
        }
    }
}

// Columns per CTA: the widest vector that divides the width, keeps the per-thread tile (rows x cols
// values, twice that in the backward) within the register budget of a 1024-thread CTA, and still
// leaves about one wave of CTAs.
static int tower_cols(int units, int rows, int budget, const void* p0, const void* p1, const void* p2) {
    int cols = 4;
    auto bad = [&](int c) {
        return units % c != 0 || rows * c > budget || units / c < 96 || ((uintptr_t)p0 % (4 * c)) != 0 ||
               ((uintptr_t)p1 % (4 * c)) != 0 || ((uintptr_t)p2 % (4 * c)) != 0;
    };
    while (cols > 1 && bad(cols)) cols >>= 1;
    return cols;
}

}  // namespace rk

extern "C" {

int rk_dice_bn_max_batch(void) { return rk::kTowerThreads * rk::kTowerMaxRows; }

int rk_dice_bn_fwd(const float* x, int64_t B, int units, const float* alpha, float eps1, const float* gamma,
                   const float* beta, float eps2, float momentum1, float* running_mean1, float* running_var1,
                   int64_t* num_batches1, float momentum2, float* running_mean2, float* running_var2,
                   int64_t* num_batches2, float* z, float* stats, rk_stream_t stream_) {
    using namespace rk;
    RK_CHECK_ARG(x && alpha && z && stats, "dice_bn_fwd: NULL pointer");
    RK_CHECK_ARG((gamma == nullptr) == (beta == nullptr), "dice_bn_fwd: gamma and beta go together");
    RK_CHECK_ARG(units >= 1 && B >= 1 && B <= rk_dice_bn_max_batch(), "dice_bn_fwd: B=%lld units=%d (B <= %d)",
                 (long long)B, units, rk_dice_bn_max_batch());
    RK_CHECK_ARG((running_mean1 == nullptr) == (running_var1 == nullptr) &&
                 (running_mean2 == nullptr) == (running_var2 == nullptr), "dice_bn_fwd: running mean/var go together");
    DiceBnArgs a{x, alpha, gamma, beta, running_mean1, running_var1, num_batches1, running_mean2, running_var2,
                 num_batches2, eps1, eps2, momentum1, momentum2, B, units};
    const int rows = (int)ceil_div(B, kTowerThreads);
    const int cols = tower_cols(units, rows, 32, x, z, z);
    const int grid = units / cols;
    cudaStream_t s = (cudaStream_t)stream_;
#define RK_TOWER_FWD(C, R) dice_bn_fwd_kernel<C, R><<<grid, kTowerThreads, 0, s>>>(a, z, stats)
#define RK_TOWER_ROWS(M, C)                                            \
    do {                                                               \
        if (rows <= 1) M(C, 1); else if (rows <= 2) M(C, 2); else if (rows <= 4) M(C, 4);   \
        else if (rows <= 8) M(C, 8); else M(C, 16);                    \
    } while (0)
    if (cols == 4) RK_TOWER_ROWS(RK_TOWER_FWD, 4); else if (cols == 2) RK_TOWER_ROWS(RK_TOWER_FWD, 2);
    else RK_TOWER_ROWS(RK_TOWER_FWD, 1);
#undef RK_TOWER_FWD
    RK_LAUNCH_CHECK();
    return 0;
}

int rk_dice_bn_bwd(const float* x, const float* g_z, int64_t B, int units, const float* alpha, const float* gamma,
                   const float* stats, float* g_x, float* g_alpha, float* g_gamma, float* g_beta,
                   rk_stream_t stream_) {
    using namespace rk;
    RK_CHECK_ARG(x && g_z && alpha && stats && g_x && g_alpha, "dice_bn_bwd: NULL pointer");
    RK_CHECK_ARG(gamma == nullptr || (g_gamma && g_beta), "dice_bn_bwd: g_gamma / g_beta missing");
    RK_CHECK_ARG(units >= 1 && B >= 1 && B <= rk_dice_bn_max_batch(), "dice_bn_bwd: B=%lld units=%d", (long long)B, units);
    DiceBnArgs a{x, alpha, gamma, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0.f, 0.f, 0.f, 0.f, B, units};
    const int rows = (int)ceil_div(B, kTowerThreads);
    const int cols = tower_cols(units, rows, 16, x, g_z, g_x);
    const int grid = units / cols;
    cudaStream_t s = (cudaStream_t)stream_;
#define RK_TOWER_BWD(C, R) dice_bn_bwd_kernel<C, R><<<grid, kTowerThreads, 0, s>>>(a, g_z, stats, g_x, g_alpha, g_gamma, g_beta)
    if (cols == 4) RK_TOWER_ROWS(RK_TOWER_BWD, 4); else if (cols == 2) RK_TOWER_ROWS(RK_TOWER_BWD, 2);
    else RK_TOWER_ROWS(RK_TOWER_BWD, 1);
#undef RK_TOWER_BWD
#undef RK_TOWER_ROWS
    RK_LAUNCH_CHECK();
    return 0;
}

int rk_bn_act_fwd(const float* x, int64_t B, int units, const float* gamma, const float* beta, float eps,
                  float momentum, float* running_mean, float* running_var, int64_t* num_batches, float slope,
                  float* z, float* stats, rk_stream_t stream_) {
    using namespace rk;
    RK_CHECK_ARG(x && gamma && beta && z && stats, "bn_act_fwd: NULL pointer");
    RK_CHECK_ARG(units >= 1 && B >= 1 && B <= rk_dice_bn_max_batch(), "bn_act_fwd: B=%lld units=%d (B <= %d)",
                 (long long)B, units, rk_dice_bn_max_batch());
    RK_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "bn_act_fwd: running mean/var go together");
    BnActArgs a{x, gamma, beta, running_mean, running_var, num_batches, eps, momentum, slope, B, units};
    const int rows = (int)ceil_div(B, kTowerThreads);
    const int cols = tower_cols(units, rows, 32, x, z, z);
    const int grid = units / cols;
    cudaStream_t s = (cudaStream_t)stream_;
#define RK_TOWER_ROWS(M, C)                                            \
    do {                                                               \
        if (rows <= 1) M(C, 1); else if (rows <= 2) M(C, 2); else if (rows <= 4) M(C, 4);   \
        else if (rows <= 8) M(C, 8); else M(C, 16);                    \
    } while (0)
#define RK_BN_FWD(C, R) bn_act_fwd_kernel<C, R><<<grid, kTowerThreads, 0, s>>>(a, z, stats)
    if (cols == 4) RK_TOWER_ROWS(RK_BN_FWD, 4); else if (cols == 2) RK_TOWER_ROWS(RK_BN_FWD, 2);
    else RK_TOWER_ROWS(RK_BN_FWD, 1);
#undef RK_BN_FWD
    RK_LAUNCH_CHECK();
    return 0;
}

int rk_bn_act_bwd(const float* x, const float* g_z, int64_t B, int units, const float* gamma, const float* beta,
                  float slope, const float* stats, float* g_x, float* g_gamma, float* g_beta, rk_stream_t stream_) {
    using namespace rk;
    RK_CHECK_ARG(x && g_z && gamma && beta && stats && g_x && g_gamma && g_beta, "bn_act_bwd: NULL pointer");
    RK_CHECK_ARG(units >= 1 && B >= 1 && B <= rk_dice_bn_max_batch(), "bn_act_bwd: B=%lld units=%d", (long long)B, units);
    BnActArgs a{x, gamma, beta, nullptr, nullptr, nullptr, 0.f, 0.f, slope, B, units};
    const int rows = (int)ceil_div(B, kTowerThreads);
    const int cols = tower_cols(units, rows, 16, x, g_z, g_x);
    const int grid = units / cols;
    cudaStream_t s = (cudaStream_t)stream_;
#define RK_BN_BWD(C, R) bn_act_bwd_kernel<C, R><<<grid, kTowerThreads, 0, s>>>(a, g_z, stats, g_x, g_gamma, g_beta)
    if (cols == 4) RK_TOWER_ROWS(RK_BN_BWD, 4); else if (cols == 2) RK_TOWER_ROWS(RK_BN_BWD, 2);
    else RK_TOWER_ROWS(RK_BN_BWD, 1);
#undef RK_BN_BWD
#undef RK_TOWER_ROWS
    RK_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
