// optim.cu — row-wise (lazy) Adam on the touched rows of an embedding table.
//
// The reference trains with optim.Adam over DENSE [V, D] gradients: every step zero-fills and sweeps
// whole tables although a batch touches a few thousand rows (38 % of the DeepFM CPU step).  Given the
// compact gradient the occurrence plan already produces — the unique touched rows and one summed
// gradient row each — this kernel updates only those rows of weight / exp_avg / exp_avg_sq.  The
// arithmetic is torch.optim.SparseAdam's (moments of untouched rows are not decayed, the bias
// correction uses the global step), which is a different optimizer from dense Adam: opt-in.
#include "common.cuh"

namespace rk {

template <int VEC>
__global__ void __launch_bounds__(256)
rowwise_adam_kernel(float* __restrict__ weight, float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                    const int64_t* __restrict__ rows, const float* __restrict__ grads, int64_t n,
                    const int64_t* __restrict__ n_dev, int D, int64_t V,
                    float beta1, float beta2, float eps, float step_size, int32_t* err_flag) {
    const int units = D / VEC;
    const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t i = gid / units;
    const int u = (int)(gid - i * units);
    if (n_dev) {                                      // the live count sits on the device (no host sync upstream)
        const int64_t live = __ldg(n_dev);
        if (live < n) n = live;
    }
    if (i >= n) return;
    const int64_t raw = __ldg(rows + i);
    if ((uint64_t)raw >= (uint64_t)V) {               // never write outside the table
        if (err_flag) atomicOr(err_flag, 1);
        return;
    }
    const int64_t off = raw * D + u * VEC;
    Vec<VEC> g, m, v, w;
    g.load(grads + i * D + u * VEC);
    m.load_plain(exp_avg + off);
    v.load_plain(exp_avg_sq + off);
    w.load_plain(weight + off);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        // torch.optim.SparseAdam: m += (g - m)(1 - b1); v += (g^2 - v)(1 - b2); w -= step_size * m / (sqrt(v) + eps)
        m.v[k] = m.v[k] + (g.v[k] - m.v[k]) * (1.0f - beta1);
        v.v[k] = v.v[k] + (g.v[k] * g.v[k] - v.v[k]) * (1.0f - beta2);
        w.v[k] = w.v[k] - step_size * (m.v[k] / (sqrtf(v.v[k]) + eps));
    }
    m.store(exp_avg + off);
    v.store(exp_avg_sq + off);
    w.store(weight + off);
}

}  // namespace rk

static int rowwise_adam_launch(float* weight, float* exp_avg, float* exp_avg_sq, const int64_t* rows, const float* grads,
                               int64_t n, const int64_t* n_dev, int D, int64_t V, float lr, float beta1, float beta2,
                               float eps, int64_t step, int32_t* err_flag, rk_stream_t stream_) {
    using namespace rk;
    RK_CHECK_ARG(weight && exp_avg && exp_avg_sq && (n == 0 || (rows && grads)), "rowwise_adam: NULL pointer");
    RK_CHECK_ARG(n >= 0 && D >= 1 && V >= 1 && step >= 1, "rowwise_adam: n=%lld D=%d V=%lld step=%lld", (long long)n, D,
                 (long long)V, (long long)step);
    if (n == 0) return 0;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr * sqrt(bc2) / bc1);
    int vec = D % 4 == 0 ? 4 : (D % 2 == 0 ? 2 : 1);
    auto mis = [&](const void* p) { return ((uintptr_t)p % (4 * vec)) != 0; };
    while (vec > 1 && (mis(weight) || mis(exp_avg) || mis(exp_avg_sq) || mis(grads))) vec >>= 1;
    const int64_t threads = n * (D / vec);
    const int grid = (int)ceil_div(threads, 256);
    cudaStream_t s = (cudaStream_t)stream_;
    if (vec == 4) rowwise_adam_kernel<4><<<grid, 256, 0, s>>>(weight, exp_avg, exp_avg_sq, rows, grads, n, n_dev, D, V, beta1, beta2, eps, step_size, err_flag);
    else if (vec == 2) rowwise_adam_kernel<2><<<grid, 256, 0, s>>>(weight, exp_avg, exp_avg_sq, rows, grads, n, n_dev, D, V, beta1, beta2, eps, step_size, err_flag);
    else rowwise_adam_kernel<1><<<grid, 256, 0, s>>>(weight, exp_avg, exp_avg_sq, rows, grads, n, n_dev, D, V, beta1, beta2, eps, step_size, err_flag);
    RK_LAUNCH_CHECK();
    return 0;
}

extern "C" int rk_rowwise_adam(float* weight, float* exp_avg, float* exp_avg_sq, const int64_t* rows,
                               const float* grads, int64_t n, int D, int64_t V, float lr, float beta1, float beta2,
                               float eps, int64_t step, int32_t* err_flag, rk_stream_t stream_) {
    return rowwise_adam_launch(weight, exp_avg, exp_avg_sq, rows, grads, n, nullptr, D, V, lr, beta1, beta2, eps, step,
                               err_flag, stream_);
}

extern "C" int rk_rowwise_adam_touched(float* weight, float* exp_avg, float* exp_avg_sq, const int64_t* rows,
                                       const float* grads, const int64_t* count, int64_t capacity, int D, int64_t V,
                                       float lr, float beta1, float beta2, float eps, int64_t step, int32_t* err_flag,
                                       rk_stream_t stream_) {
    RK_CHECK_ARG(count, "rowwise_adam_touched: count is NULL");
    return rowwise_adam_launch(weight, exp_avg, exp_avg_sq, rows, grads, capacity, count, D, V, lr, beta1, beta2, eps,
                               step, err_flag, stream_);
}
