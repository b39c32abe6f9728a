// abi.cu — ABI bookkeeping: version, thread-local error text, field packing.
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include "common.cuh"
#include "prof.cuh"

namespace rk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

int sm_count() {
    static int cached[64] = {0};            // per device ordinal: a process may drive several GPUs
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < 64 && cached[dev]) return cached[dev];
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
        return 148;
    if (dev >= 0 && dev < 64) cached[dev] = n;
    return n;
}

int pick_vec(const rk_field_t* f, int F, int n_dense, int extra) {
    int g = 4;
    auto fold = [&](int x) {
        while (g > 1 && (x % g) != 0) g >>= 1;
    };
    fold(n_dense);
    fold(extra);
    for (int i = 0; i < F; ++i) {
        fold(f[i].dim);
        fold(f[i].out_off);
    }
    return g;
}

int pack_fields(const rk_field_t* f, int F, FieldSet* out) {
    RK_CHECK_ARG(F >= 0 && F <= RK_MAX_FIELDS, "F=%d outside [0,%d]", F, RK_MAX_FIELDS);
    RK_CHECK_ARG(F == 0 || f != nullptr, "fields is NULL");
    memset(out, 0, sizeof(*out));
    out->F = F;
    for (int i = 0; i < F; ++i) {
        RK_CHECK_ARG(f[i].weight && f[i].idx, "field %d: NULL weight or idx", i);
        RK_CHECK_ARG(f[i].rows > 0 && f[i].dim > 0, "field %d: rows=%lld dim=%d", i,
                     (long long)f[i].rows, f[i].dim);
        out->weight[i] = f[i].weight;
        out->idx[i]    = f[i].idx;
        out->rows[i]   = f[i].rows;
        out->dim[i]    = f[i].dim;
        out->off[i]    = f[i].out_off;
    }
    return 0;
}

// One warp per output element: lane l adds the partials of CTAs l, l+32, ... in double, then the
// 32 lane sums are folded by a fixed shuffle tree — the order never depends on scheduling.
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partials, int n_cta, int count, float* __restrict__ out) {
    const int i = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= count) return;
    double a = 0.0;
    for (int c = lane; c < n_cta; c += 32) a += (double)partials[(int64_t)c * count + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(kFull, a, o);
    if (lane == 0) out[i] = (float)a;
}

int launch_reduce_partials(const float* partials, int n_cta, int count, float* out, cudaStream_t s) {
    const int64_t threads = (int64_t)count * 32;
    reduce_partials_kernel<<<(int)ceil_div(threads, 256), 256, 0, s>>>(partials, n_cta, count, out);
    RK_LAUNCH_CHECK();
    return 0;
}

// Measurement aid: keeps the stream busy for `us` microseconds so that the host can enqueue a
// whole step behind it; the kernels then run back to back and CUDA events see device time only.
__global__ void spin_kernel(unsigned long long ns) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        __nanosleep(1000);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    } while (t1 - t0 < ns);
}

}  // namespace rk

extern "C" {

int rk_version(void) { return RK_ABI_VERSION; }

const char* rk_last_error(void) { return rk::g_err; }

int rk_device_sm_count(void) { return rk::sm_count(); }

long long rk_launch_count(void) { return rk::launches(); }

int rk_debug_spin(int us, rk_stream_t stream) {
    RK_CHECK_ARG(us >= 0 && us <= 100000, "rk_debug_spin: %d us outside [0, 100000]", us);
    rk::spin_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long)us * 1000ull);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) { rk::set_error("spin launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

}  // extern "C"

#ifdef RK_PROFILE
namespace rk { __device__ unsigned long long g_rk_prof[16]; }
extern "C" int rk_debug_profile(unsigned long long* out16, int reset) {
    RK_CUDA(cudaDeviceSynchronize());
    RK_CUDA(cudaMemcpyFromSymbol(out16, rk::g_rk_prof, sizeof(unsigned long long) * 16));
    if (reset) {
        unsigned long long z[16] = {0};
        RK_CUDA(cudaMemcpyToSymbol(rk::g_rk_prof, z, sizeof(z)));
    }
    return 0;
}
#endif
