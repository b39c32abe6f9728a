// segment_reduce.cu — rk_embgrad_segment_reduce: dense [rows, dim] embedding gradients from
// per-occurrence gradient rows, using the (field,row)-sorted order made by rk_plan_build.
// Replaces ATen embedding_dense_backward (autograd of nn.Embedding, e.g. DCN/dcn.py:166 via
// loss.backward()).  No atomics: every row's occurrences are summed in sorted (= occurrence)
// order; rows hotter than one 16-occurrence chunk are finished by a second, equally
// deterministic pass that adds the chunk partials left to right.
//
// HBM/L2-bound: each gradient row is read once (gathered through perm), each touched table row
// written once.  One table with `dim` columns uses ceil(dim/vec) lanes per chunk, every lane
// owning `vec` columns, so no cross-lane traffic is needed.
#include <string.h>
#include "common.cuh"

namespace rk {

constexpr int kChunk = 16;

struct ReduceJob {
    const float* g;
    float*       dw;
    float*       lead;        // [n_chunks, dim] partial of a segment begun in an earlier chunk
    int64_t      ld;
    int64_t      seg_start;   // first sorted position of the field
    int64_t      n;           // occurrences of the field
    int64_t      thread_start;
    int64_t      warp_start;  // combine pass: first warp of this table
    uint32_t     key_base;
    uint32_t     dead_key;    // key_base + rows: occurrences that carry no gradient
    int32_t      dim;
    int32_t      lanes;
    int32_t      vec;
};
struct ReduceParams {
    ReduceJob job[RK_MAX_TABLES];
    int64_t   total_threads;
    int64_t   total_warps;
    int32_t   n_tables;
};

template <int V>
__device__ __forceinline__ void reduce_chunk(const ReduceJob& jb, const uint32_t* __restrict__ keys,
                                             const uint32_t* __restrict__ perm, int64_t c,
                                             int lane) {
    const int     col  = lane * V;
    const int64_t pos0 = jb.seg_start + c * kChunk;
    const int     cnt  = (int)((jb.n - c * kChunk) < kChunk ? (jb.n - c * kChunk) : kChunk);
    const uint32_t* K  = keys + pos0;
    const uint32_t* P  = perm + pos0;
    uint32_t cur  = K[0];
    if (cur == jb.dead_key) return;   // sorted: the whole chunk is padding
    bool     lead = (c > 0) && (keys[pos0 - 1] == cur);
    Vec<V>   acc;
    vec_zero(acc);
    float* lead_out = jb.lead + c * (int64_t)jb.dim + col;

#pragma unroll
    for (int j0 = 0; j0 < kChunk; j0 += 8) {
        if (j0 >= cnt) break;
        uint32_t kk[8];
        Vec<V>   vv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = j0 + u < cnt ? j0 + u : cnt - 1;
            kk[u] = K[j];
            vv[u].load_plain(jb.g + (int64_t)P[j] * jb.ld + col);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (j0 + u < cnt) {
                if (kk[u] != cur) {
                    if (lead) acc.store(lead_out);
                    else      acc.store(jb.dw + (int64_t)(cur - jb.key_base) * jb.dim + col);
                    lead = false;
                    vec_zero(acc);
                    cur = kk[u];
                    if (cur == jb.dead_key) return;   // rest of the chunk is padding
                }
#pragma unroll
                for (int e = 0; e < V; ++e) acc.v[e] += vv[u].v[e];
            }
        }
    }
    if (lead) acc.store(lead_out);
    else      acc.store(jb.dw + (int64_t)(cur - jb.key_base) * jb.dim + col);
}

__device__ __forceinline__ int find_job(const ReduceParams& p, int64_t t) {
    int j = 0;
#pragma unroll 1
    while (j + 1 < p.n_tables && t >= p.job[j + 1].thread_start) ++j;
    return j;
}

__global__ void __launch_bounds__(256)
segment_chunk_kernel(const __grid_constant__ ReduceParams p, const uint32_t* __restrict__ keys,
                     const uint32_t* __restrict__ perm) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= p.total_threads) return;
    const ReduceJob& jb = p.job[find_job(p, t)];
    const int64_t local = t - jb.thread_start;
    const int64_t c     = local / jb.lanes;
    const int     lane  = (int)(local - c * jb.lanes);
    if (c * kChunk >= jb.n) return;
    if (jb.vec == 4)      reduce_chunk<4>(jb, keys, perm, c, lane);
    else if (jb.vec == 2) reduce_chunk<2>(jb, keys, perm, c, lane);
    else                  reduce_chunk<1>(jb, keys, perm, c, lane);
}

// Second pass, one warp per chunk: the chunk that holds the head of a row spilling into later
// chunks adds their lead partials.  The end of the run is found by binary search on the sorted
// keys; the partials are summed by row-lanes in a fixed interleaved order and folded with a fixed
// shuffle tree, so the result does not depend on scheduling.  Runs longer than kLongRun chunks
// (a padding row hit by tens of thousands of positions) are summed by all warps of the CTA, whose
// per-warp totals are then added in warp order.
constexpr int kLongRun = 32;

// Is chunk c the head of a row that spills into later chunks?  If so: its key and last chunk.
__device__ __forceinline__ bool find_run(const ReduceJob& jb, const uint32_t* __restrict__ keys, int64_t c,
                                         uint32_t* key_out, int64_t* c_last_out) {
    const int64_t n_chunks = (jb.n + kChunk - 1) / kChunk;
    const int64_t pos0     = jb.seg_start + c * kChunk;
    const int     cnt      = (int)((jb.n - c * kChunk) < kChunk ? (jb.n - c * kChunk) : kChunk);
    const int64_t last     = pos0 + cnt - 1;
    if (c + 1 >= n_chunks) return false;
    const uint32_t k = keys[last];
    if (k == jb.dead_key || keys[last + 1] != k) return false;   // nothing spills out of this chunk
    const bool head_here = (keys[pos0] != k) || c == 0 || keys[pos0 - 1] != k;
    if (!head_here) return false;
    // chunks c+1 .. c_last start with key k (predicate is monotone over the sorted keys): gallop
    // from c+1 (most runs end within a chunk or two), then bisect the last doubling step
    int64_t lo = c + 1, step = 1;
    while (lo + step <= n_chunks - 1 && keys[jb.seg_start + (lo + step) * kChunk] == k) {
        lo += step;
        step <<= 1;
    }
    int64_t hi = lo + step - 1 < n_chunks - 1 ? lo + step - 1 : n_chunks - 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (keys[jb.seg_start + mid * kChunk] == k) lo = mid; else hi = mid - 1;
    }
    *key_out = k;
    *c_last_out = lo;
    return true;
}

// Sum of lead[first + slot + i*stride], i >= 0, up to c_last, over this warp's row-lanes; the
// warp total ends up in the lanes with rl == 0 (one per column lane).
template <int V>
__device__ __forceinline__ Vec<V> sum_run(const ReduceJob& jb, int64_t first, int64_t c_last, int lane,
                                          int slot0, int stride, int RL, bool* owner) {
    const int CL = jb.lanes;
    const int rl = lane / CL, cl = lane - rl * CL;
    const bool on = rl < RL;
    Vec<V> acc;
    vec_zero(acc);
    if (on) {
        const float* src = jb.lead + cl * V;
        int64_t cc = first + slot0 + rl;
        for (; cc + 3 * (int64_t)stride <= c_last; cc += 4 * (int64_t)stride) {   // four loads in flight
            Vec<V> p0, p1, p2, p3;
            p0.load_plain(src + cc * (int64_t)jb.dim);
            p1.load_plain(src + (cc + stride) * (int64_t)jb.dim);
            p2.load_plain(src + (cc + 2 * stride) * (int64_t)jb.dim);
            p3.load_plain(src + (cc + 3 * stride) * (int64_t)jb.dim);
#pragma unroll
            for (int e = 0; e < V; ++e) acc.v[e] += (p0.v[e] + p1.v[e]) + (p2.v[e] + p3.v[e]);
        }
        for (; cc <= c_last; cc += stride) {
            Vec<V> part;
            part.load_plain(src + cc * (int64_t)jb.dim);
#pragma unroll
            for (int e = 0; e < V; ++e) acc.v[e] += part.v[e];
        }
    }
    for (int o = RL >> 1; o > 0; o >>= 1) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const float other = __shfl_down_sync(kFull, acc.v[e], o * CL);
            if (on && rl < o) acc.v[e] += other;
        }
    }
    *owner = on && rl == 0;
    return acc;
}

struct LongRun {
    int64_t  c, c_last;
    uint32_t key;
    int32_t  table;      // index into ReduceParams::job, -1 = none
};

template <int V>
__device__ __forceinline__ void combine_dispatch(const ReduceJob& jb, bool long_phase, uint32_t k, int64_t c,
                                                 int64_t c_last, int lane, int warp, float* red) {
    const int CL = jb.lanes;
    int RL = 1;
    while (RL * 2 * CL <= 32) RL *= 2;             // row lanes, power of two
    const int cl = lane % CL;
    bool owner;
    if (!long_phase) {                             // the head's own warp sums the whole run
        Vec<V> acc = sum_run<V>(jb, c + 1, c_last, lane, 0, RL, RL, &owner);
        if (owner) {
            float* out = jb.dw + (int64_t)(k - jb.key_base) * jb.dim + cl * V;
            Vec<V> own;
            own.load_plain(out);
#pragma unroll
            for (int e = 0; e < V; ++e) own.v[e] += acc.v[e];
            own.store(out);
        }
        return;
    }
    // all 8 warps of the CTA: warp w takes the row-lane slots [w*RL, (w+1)*RL) of 8*RL
    Vec<V> acc = sum_run<V>(jb, c + 1, c_last, lane, warp * RL, 8 * RL, RL, &owner);
    if (owner) {
#pragma unroll
        for (int e = 0; e < V; ++e) red[warp * 128 + cl * V + e] = acc.v[e];
    }
    __syncthreads();
    if (warp == 0 && lane < CL) {
        float* out = jb.dw + (int64_t)(k - jb.key_base) * jb.dim + lane * V;
        Vec<V> own;
        own.load_plain(out);
        for (int ww = 0; ww < 8; ++ww) {
#pragma unroll
            for (int e = 0; e < V; ++e) own.v[e] += red[ww * 128 + lane * V + e];
        }
        own.store(out);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256)
segment_combine_kernel(const __grid_constant__ ReduceParams p, const uint32_t* __restrict__ keys) {
    __shared__ LongRun jobs[8];
    __shared__ float red[8 * 128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (lane == 0) jobs[warp].table = -1;
    if (w < p.total_warps) {
        int j = 0;
#pragma unroll 1
        while (j + 1 < p.n_tables && w >= p.job[j + 1].warp_start) ++j;
        const ReduceJob& jb = p.job[j];
        const int64_t c = w - jb.warp_start;
        uint32_t k;
        int64_t c_last;
        if (c * kChunk < jb.n && find_run(jb, keys, c, &k, &c_last)) {
            if (c_last - c > kLongRun) {
                if (lane == 0) { jobs[warp].c = c; jobs[warp].c_last = c_last; jobs[warp].key = k; jobs[warp].table = j; }
            } else if (jb.vec == 4) combine_dispatch<4>(jb, false, k, c, c_last, lane, warp, red);
            else if (jb.vec == 2)   combine_dispatch<2>(jb, false, k, c, c_last, lane, warp, red);
            else                    combine_dispatch<1>(jb, false, k, c, c_last, lane, warp, red);
        }
    }
    __syncthreads();
    for (int q = 0; q < 8; ++q) {                  // long runs found by this CTA's warps, one by one
        const LongRun r = jobs[q];
        if (r.table < 0) continue;                 // uniform over the CTA
        const ReduceJob& jb = p.job[r.table];
        if (jb.vec == 4)      combine_dispatch<4>(jb, true, r.key, r.c, r.c_last, lane, warp, red);
        else if (jb.vec == 2) combine_dispatch<2>(jb, true, r.key, r.c, r.c_last, lane, warp, red);
        else                  combine_dispatch<1>(jb, true, r.key, r.c, r.c_last, lane, warp, red);
    }
}

static int vec_of(int dim) { return dim % 4 == 0 ? 4 : (dim % 2 == 0 ? 2 : 1); }

}  // namespace rk

extern "C" {

size_t rk_reduce_workspace_bytes(const int64_t* n, int F, const rk_grad_table_t* tables,
                                 int n_tables) {
    size_t bytes = 0;
    for (int t = 0; t < n_tables; ++t) {
        const int f = tables[t].field;
        if (f < 0 || f >= F) continue;
        const size_t chunks = (size_t)rk::ceil_div(n[f], rk::kChunk);
        bytes += ((chunks * (size_t)tables[t].dim * 4) + 255) & ~(size_t)255;
    }
    return bytes ? bytes : 256;
}

int rk_embgrad_segment_reduce(const uint32_t* sorted_keys, const uint32_t* perm,
                              const int64_t* n, const int64_t* rows, int F,
                              const rk_grad_table_t* tables, int n_tables, void* ws,
                              size_t ws_bytes, rk_stream_t stream_) {
    using namespace rk;
    cudaStream_t s = (cudaStream_t)stream_;
    RK_CHECK_ARG(F >= 1 && F <= RK_MAX_FIELDS, "segment_reduce: F=%d", F);
    RK_CHECK_ARG(n_tables >= 1 && n_tables <= RK_MAX_TABLES, "segment_reduce: n_tables=%d",
                 n_tables);
    RK_CHECK_ARG(n && rows && tables, "segment_reduce: NULL host array");
    RK_CHECK_ARG(ws_bytes >= rk_reduce_workspace_bytes(n, F, tables, n_tables),
                 "segment_reduce: workspace too small");
    int64_t  start[RK_MAX_FIELDS + 1];
    uint32_t base[RK_MAX_FIELDS];
    int64_t  tot = 0, space = 0;
    for (int f = 0; f < F; ++f) {
        start[f] = tot;
        base[f]  = (uint32_t)space;
        tot += n[f];
        space += rows[f] + 1;   // same key space as rk_plan_build (one sentinel row per field)
    }
    if (tot == 0) return 0;
    RK_CHECK_ARG(sorted_keys && perm && ws, "segment_reduce: NULL device pointer");

    ReduceParams p;
    memset(&p, 0, sizeof(p));
    p.n_tables = 0;
    int64_t threads = 0, warps = 0;
    size_t  ws_off  = 0;
    for (int t = 0; t < n_tables; ++t) {
        const rk_grad_table_t& tb = tables[t];
        RK_CHECK_ARG(tb.field >= 0 && tb.field < F, "segment_reduce: table %d field %d", t,
                     tb.field);
        RK_CHECK_ARG(tb.dim > 0 && tb.g && tb.dw, "segment_reduce: table %d dim/pointers", t);
        const int f = tb.field;
        if (n[f] == 0) continue;
        int v = vec_of(tb.dim);
        // vector access needs matching alignment of the gradient rows and of the table
        while (v > 1 && ((tb.ld % v) != 0 || ((uintptr_t)tb.g % (4 * v)) != 0 ||
                         ((uintptr_t)tb.dw % (4 * v)) != 0))
            v >>= 1;
        ReduceJob& jb   = p.job[p.n_tables++];
        jb.g            = tb.g;
        jb.dw           = tb.dw;
        jb.ld           = tb.ld;
        jb.dim          = tb.dim;
        jb.vec          = v;
        jb.lanes        = tb.dim / v;
        RK_CHECK_ARG(jb.lanes <= 32, "segment_reduce: table %d dim %d too wide", t, tb.dim);
        jb.seg_start    = start[f];
        jb.n            = n[f];
        jb.key_base     = base[f];
        jb.dead_key     = base[f] + (uint32_t)rows[f];
        jb.warp_start   = warps;
        jb.lead         = (float*)((char*)ws + ws_off);
        jb.thread_start = threads;
        const int64_t chunks = ceil_div(n[f], kChunk);
        ws_off += ((size_t)chunks * tb.dim * 4 + 255) & ~(size_t)255;
        threads += ceil_div(chunks * jb.lanes, 32) * 32;  // keep warps inside one table
        warps += chunks;
    }
    p.total_threads = threads;
    p.total_warps   = warps;
    if (threads == 0) return 0;
    segment_chunk_kernel<<<(int)ceil_div(threads, 256), 256, 0, s>>>(p, sorted_keys, perm);
    RK_LAUNCH_CHECK();
    segment_combine_kernel<<<(int)ceil_div(warps * 32, 256), 256, 0, s>>>(p, sorted_keys);
    RK_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
