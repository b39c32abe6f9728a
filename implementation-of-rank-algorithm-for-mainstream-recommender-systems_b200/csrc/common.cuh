// common.cuh — shared device/host helpers for librank_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "rank_b200.h"

#ifndef __CUDA_ARCH_FEAT_SM100_ALL
#if defined(__CUDA_ARCH__)
#error "librank_b200 is written for sm_100a only: compile with -gencode arch=compute_100a,code=sm_100a"
#endif
#endif

namespace rk {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ---- host-side error plumbing (thread-local message, see rk_last_error) -------------------
void set_error(const char* fmt, ...);
int  sm_count();
void note_launch();  // counts kernels launched through the ABI (rk_launch_count)
// out[i] = sum over c < n_cta (fixed order, double accumulation) of partials[c*count + i]:
// the deterministic last step of the batch-wide weight-gradient reductions (AFM, BST).
int  launch_reduce_partials(const float* partials, int n_cta, int count, float* out, cudaStream_t s);

#define RK_CHECK_ARG(cond, ...)              \
    do {                                     \
        if (!(cond)) {                       \
            rk::set_error(__VA_ARGS__);      \
            return -1;                       \
        }                                    \
    } while (0)

#define RK_CUDA(call)                                                                  \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) {                                                       \
            rk::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),      \
                          __FILE__, __LINE__);                                         \
            return (int)e_;                                                            \
        }                                                                              \
    } while (0)

#define RK_LAUNCH_CHECK()                                                              \
    do {                                                                               \
        rk::note_launch();                                                             \
        cudaError_t e_ = cudaPeekAtLastError();                                        \
        if (e_ != cudaSuccess) {                                                       \
            rk::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_),  \
                          __FILE__, __LINE__);                                         \
            return (int)e_;                                                            \
        }                                                                              \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers -----------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
template <int W>
__device__ __forceinline__ float group_sum(float v) {  // sum inside aligned groups of W lanes
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// Vector of N floats with matching-width global/shared access.
template <int N> struct Vec;
template <> struct Vec<1> {
    float v[1];
    __device__ __forceinline__ void load(const float* p) { v[0] = __ldg(p); }
    __device__ __forceinline__ void load_plain(const float* p) { v[0] = *p; }
    __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
};
template <> struct Vec<2> {
    float v[2];
    __device__ __forceinline__ void load(const float* p) {
        float2 t = __ldg(reinterpret_cast<const float2*>(p)); v[0] = t.x; v[1] = t.y;
    }
    __device__ __forceinline__ void load_plain(const float* p) {
        float2 t = *reinterpret_cast<const float2*>(p); v[0] = t.x; v[1] = t.y;
    }
    __device__ __forceinline__ void store(float* p) const {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    }
};
template <> struct Vec<4> {
    float v[4];
    __device__ __forceinline__ void load(const float* p) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ void load_plain(const float* p) {
        float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ void store(float* p) const {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <int N>
__device__ __forceinline__ void vec_zero(Vec<N>& a) {
#pragma unroll
    for (int i = 0; i < N; ++i) a.v[i] = 0.f;
}

// ---- counter-based random bits (Philox4x32-10) for in-kernel dropout masks ------------------
// The reference draws dropout masks from torch's CPU/CUDA generators (BST/bst.py:57,62,86,90);
// those draws cannot be replayed in a fused kernel, so masks are a pure function of
// (seed, offset, row): the backward regenerates exactly the bits the forward used and nothing
// is stored.  rng[0] = seed, rng[1] = offset (advanced once per forward by the caller).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
// 16 keep-bits (bit i set = element i kept) for dropout site `site` of `row`: two Philox calls,
// eight 16-bit uniforms each; an element is dropped when its uniform is below thr = round(p * 65536).
__device__ __forceinline__ uint32_t dropout_keep16(uint64_t seed, uint64_t offset, uint64_t row, int site,
                                                   uint32_t thr) {
    uint32_t bits = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)row, (uint32_t)(row >> 32), (uint32_t)(2 * site + half),
                                                 (uint32_t)offset),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(offset >> 32)));
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            bits |= ((w[j] & 0xffffu) >= thr ? 1u : 0u) << (8 * half + 2 * j);
            bits |= ((w[j] >> 16) >= thr ? 1u : 0u) << (8 * half + 2 * j + 1);
        }
    }
    return bits;
}

// Clamp an index into [0, rows); flag the batch as bad if it was outside (reference: IndexError).
__device__ __forceinline__ int64_t checked_row(int64_t i, int64_t rows, int32_t* err_flag) {
    if ((uint64_t)i >= (uint64_t)rows) {
        if (err_flag) atomicOr(err_flag, 1);
        return 0;
    }
    return i;
}

// Fields packed for pass-by-value as a __grid_constant__ kernel parameter.
struct FieldSet {
    const float*   weight[RK_MAX_FIELDS];
    const int64_t* idx[RK_MAX_FIELDS];
    int64_t        rows[RK_MAX_FIELDS];
    int32_t        dim[RK_MAX_FIELDS];
    int32_t        off[RK_MAX_FIELDS];
    int32_t        F;
};

// Widest vector width (4, 2 or 1 floats) that divides every dim and every offset.
int  pick_vec(const rk_field_t* f, int F, int n_dense, int extra);
int  pack_fields(const rk_field_t* f, int F, FieldSet* out);  // validates, returns 0 or <0

}  // namespace rk
