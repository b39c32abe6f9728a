// direct_reduce.cu — rk_embgrad_direct_reduce: dense [rows, dim] embedding gradients of tables read
// through a per-sample index column (n <= 8192 occurrences), in ONE launch, with no prior
// rk_plan_build and no pre-zeroed output.  Replaces ATen embedding_dense_backward (autograd of
// nn.Embedding, reached from loss.backward(), e.g. DCN/dcn.py:166, DeepFM/deepfm.py:170).
//
// Output-partitioned: every CTA owns 256 consecutive rows of one table.  It
//   1. writes zeros over its rows (the caller's gradient buffer needs no memset),
//   2. scans the WHOLE index column (<= 64 KB, L2-resident; every CTA of the table reads it) and
//      keeps the occurrences whose row falls into its range,
//   3. orders them by (row, occurrence) with one stable 8-bit counting pass in shared memory
//      (warp-match ranking in occurrence order, as rk_plan_build's sort does),
//   4. sums every row's gradient rows in that order: short runs one after the other by a lane
//      group, runs of more than 32 occurrences by a warp / the whole CTA over fixed interleaved
//      slots folded by a fixed shuffle tree.
// No atomics on floats, no dependence on scheduling: the result is a function of the inputs only.
// Tables that share an index column (DeepFM / FwFM: first-order [rows,1] and second-order
// [rows,D] tables of a field) are one job with several outputs: the column is scanned and ordered once.
//
// HBM/L2-bound integer + float work: per table, n*8 B of indices per CTA (L2), n*dim*4 B of gradient
// rows once, rows*dim*4 B of output once.
#include <string.h>
#include "common.cuh"

namespace rk {

constexpr int kDirThreads = 256;
constexpr int kDirWarps   = kDirThreads / 32;
constexpr int kDirMaxN    = RK_DIRECT_MAX_N;             // 8192
constexpr int kDirRounds  = kDirMaxN / kDirThreads;      // 32 occurrences per thread
constexpr int kDirRows    = 256;                         // rows per CTA = bins of the counting pass
constexpr int kDirOccBits = 24;                          // packed entry: local row << 24 | occurrence
constexpr int kDirMaxOut  = 3;
constexpr int kDirMaxJobs = RK_MAX_FIELDS;
constexpr int kSeqRun     = 32;                          // runs up to this length: one lane group, in order
constexpr int kWarpRun    = 1024;                        // up to this: one warp; longer: the whole CTA

struct DirectOut {
    const float* g;
    float*       dw;
    int32_t      ld;
    int16_t      dim;
    int8_t       vec;
    int8_t       lanes;
};
struct DirectJob {
    const int64_t* idx;
    int32_t        n;
    int32_t        rows;
    int32_t        cta_start;
    int32_t        n_out;
    DirectOut      out[kDirMaxOut];
};
struct DirectParams {
    DirectJob job[kDirMaxJobs];
    int32_t   n_jobs;
};

// Sum of the gradient rows of entries start+slot, start+slot+nslots, ... (< start+len), in that order.
template <int V>
__device__ __forceinline__ Vec<V> strided_sum(const uint32_t* __restrict__ ent, int start, int len, int slot,
                                              int nslots, const float* __restrict__ g, int64_t ld, int col) {
    Vec<V> acc;
    vec_zero(acc);
    int j = slot;
    for (; j + 3 * nslots < len; j += 4 * nslots) {                 // four loads in flight
        Vec<V> p0, p1, p2, p3;
        const uint32_t o0 = ent[start + j] & ((1u << kDirOccBits) - 1u);
        const uint32_t o1 = ent[start + j + nslots] & ((1u << kDirOccBits) - 1u);
        const uint32_t o2 = ent[start + j + 2 * nslots] & ((1u << kDirOccBits) - 1u);
        const uint32_t o3 = ent[start + j + 3 * nslots] & ((1u << kDirOccBits) - 1u);
        p0.load_plain(g + (int64_t)o0 * ld + col);
        p1.load_plain(g + (int64_t)o1 * ld + col);
        p2.load_plain(g + (int64_t)o2 * ld + col);
        p3.load_plain(g + (int64_t)o3 * ld + col);
#pragma unroll
        for (int e = 0; e < V; ++e) acc.v[e] = (((acc.v[e] + p0.v[e]) + p1.v[e]) + p2.v[e]) + p3.v[e];
    }
    for (; j < len; j += nslots) {
        Vec<V> q;
        const uint32_t o = ent[start + j] & ((1u << kDirOccBits) - 1u);
        q.load_plain(g + (int64_t)o * ld + col);
#pragma unroll
        for (int e = 0; e < V; ++e) acc.v[e] += q.v[e];
    }
    return acc;
}

// All rows of this CTA's range for one output table.  run_start[d] / run_len[d]: where local row d's
// occurrences sit in ent[].  red: [kDirWarps][32] floats of scratch.
template <int V>
__device__ __forceinline__ void reduce_rows(const DirectOut& o, const uint32_t* __restrict__ ent,
                                            const uint32_t* __restrict__ run_start,
                                            const uint32_t* __restrict__ run_len, int row0, int n_rows,
                                            float* red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int CL = o.lanes;                          // column lanes per row
    // ---- short runs: one lane group per row, occurrences added one after the other
    for (int item = tid; item < n_rows * CL; item += kDirThreads) {
        const int d = item / CL, cl = item - d * CL;
        const int len = (int)run_len[d];
        if (len == 0 || len > kSeqRun) continue;
        Vec<V> acc = strided_sum<V>(ent, (int)run_start[d], len, 0, 1, o.g, o.ld, cl * V);
        acc.store(o.dw + (int64_t)(row0 + d) * o.dim + cl * V);
    }
    // ---- medium runs: one warp per row, RL interleaved slots per column lane, fixed shuffle tree
    int RL = 1;
    while (RL * 2 * CL <= 32) RL *= 2;
    const int rl = lane / CL, cl = lane - rl * CL;
    const bool on = rl < RL;
    for (int d = warp; d < n_rows; d += kDirWarps) {
        const int len = (int)run_len[d];
        if (len <= kSeqRun || len > kWarpRun) continue;        // uniform over the warp
        Vec<V> acc;
        vec_zero(acc);
        if (on) acc = strided_sum<V>(ent, (int)run_start[d], len, rl, RL, o.g, o.ld, cl * V);
        for (int s = RL >> 1; s > 0; s >>= 1) {
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const float other = __shfl_down_sync(kFull, acc.v[e], s * CL);
                if (on && rl < s) acc.v[e] += other;
            }
        }
        if (on && rl == 0) acc.store(o.dw + (int64_t)(row0 + d) * o.dim + cl * V);
    }
    // ---- long runs (a handful per CTA at most): all warps, warp w takes slots [w*RL, (w+1)*RL) of 8*RL
    for (int d = 0; d < n_rows; ++d) {
        const int len = (int)run_len[d];
        if (len <= kWarpRun) continue;                         // uniform over the CTA
        Vec<V> acc;
        vec_zero(acc);
        if (on) acc = strided_sum<V>(ent, (int)run_start[d], len, warp * RL + rl, kDirWarps * RL, o.g, o.ld, cl * V);
        for (int s = RL >> 1; s > 0; s >>= 1) {
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const float other = __shfl_down_sync(kFull, acc.v[e], s * CL);
                if (on && rl < s) acc.v[e] += other;
            }
        }
        if (on && rl == 0) {
#pragma unroll
            for (int e = 0; e < V; ++e) red[warp * 32 + cl * V + e] = acc.v[e];
        }
        __syncthreads();
        if (warp == 0 && lane < CL) {
            Vec<V> tot;
            vec_zero(tot);
            for (int ww = 0; ww < kDirWarps; ++ww) {
#pragma unroll
                for (int e = 0; e < V; ++e) tot.v[e] += red[ww * 32 + lane * V + e];
            }
            tot.store(o.dw + (int64_t)(row0 + d) * o.dim + lane * V);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kDirThreads, 3)
direct_reduce_kernel(const __grid_constant__ DirectParams p, int32_t* err_flag) {
    __shared__ uint32_t ent[kDirMaxN];                    // 32 KB: entries ordered by (row, occurrence)
    __shared__ uint16_t cnt[kDirWarps][kDirRows];         // 4 KB
    __shared__ uint32_t run_start[kDirRows];
    __shared__ uint32_t run_len[kDirRows];
    __shared__ uint32_t warp_tot[kDirWarps];
    __shared__ float    red[kDirWarps * 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int j = 0;
#pragma unroll 1
    while (j + 1 < p.n_jobs && (int)blockIdx.x >= p.job[j + 1].cta_start) ++j;
    const DirectJob& jb = p.job[j];
    const int row0   = ((int)blockIdx.x - jb.cta_start) * kDirRows;
    const int n_rows = jb.rows - row0 < kDirRows ? jb.rows - row0 : kDirRows;
    const int n      = jb.n;

    // 1. zeros over this CTA's rows of every output (fire and forget; the sums land after a barrier)
    for (int q = 0; q < jb.n_out; ++q) {
        const DirectOut& o = jb.out[q];
        float* base = o.dw + (int64_t)row0 * o.dim;
        const int total = n_rows * o.dim;
        if (o.vec == 4) {
            for (int i = tid * 4; i < total; i += kDirThreads * 4)
                *reinterpret_cast<float4*>(base + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            for (int i = tid; i < total; i += kDirThreads) base[i] = 0.f;
        }
    }

    // 2. the whole index column, warp-blocked so that (warp, round, lane) order = occurrence order
    const int per_warp = ((n + kDirWarps * 32 - 1) / (kDirWarps * 32)) * 32;     // multiple of 32
    const int rounds   = per_warp / 32;
    const int wbase    = warp * per_warp;
    // per round one word: bits 0..8 = local row (kNone: not ours), bits 9.. = rank among equal rows so far
    constexpr uint32_t kNone = 0x1ffu;
    uint32_t st[kDirRounds];
#pragma unroll
    for (int h = 0; h < kDirRounds; h += 16) {          // 16 index loads in flight per thread
        int64_t raw[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int i = wbase + (h + r) * 32 + lane;
            raw[r] = (h + r < rounds && i < n) ? __ldg(jb.idx + i) : -1;
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int i = wbase + (h + r) * 32 + lane;
            uint32_t l = kNone;
            if (h + r < rounds && i < n) {
                const int64_t row = checked_row(raw[r], jb.rows, err_flag);
                const int64_t d = row - row0;
                if (d >= 0 && d < n_rows) l = (uint32_t)d;
            }
            st[h + r] = l;
        }
    }
    for (int i = tid; i < kDirWarps * kDirRows / 2; i += kDirThreads) reinterpret_cast<uint32_t*>(&cnt[0][0])[i] = 0;
    __syncthreads();

    // 3. stable rank inside the warp's block of occurrences
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < kDirRounds; ++r) {
        const bool valid = st[r] != kNone;
        if (__ballot_sync(kFull, valid) == 0u) continue;          // uniform: nobody of this round is ours
        const uint32_t tag = valid ? st[r] : (0x10000u | lane);   // the others match nobody
        const unsigned m   = __match_any_sync(kFull, tag);
        const uint32_t old = valid ? cnt[warp][st[r]] : 0u;
        __syncwarp();
        if (valid && (m & lt) == 0) cnt[warp][st[r]] = (uint16_t)(old + __popc(m));
        __syncwarp();
        if (valid) st[r] |= (old + __popc(m & lt)) << 9;
    }
    __syncthreads();

    // 4. per row: exclusive prefix over the warps, then exclusive scan over the rows
    {
        uint32_t tot = 0;
#pragma unroll
        for (int ww = 0; ww < kDirWarps; ++ww) {
            const uint32_t c = cnt[ww][tid];
            cnt[ww][tid] = (uint16_t)tot;
            tot += c;
        }
        uint32_t inc = tot;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const uint32_t v = __shfl_up_sync(kFull, inc, s);
            if (lane >= s) inc += v;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        uint32_t before = 0;
        for (int ww = 0; ww < warp; ++ww) before += warp_tot[ww];
        run_start[tid] = before + inc - tot;
        run_len[tid]   = tot;
    }
    __syncthreads();

    // 5. scatter into (row, occurrence) order
#pragma unroll
    for (int r = 0; r < kDirRounds; ++r) {
        if ((st[r] & kNone) != kNone) {
            const uint32_t d = st[r] & kNone;
            ent[run_start[d] + cnt[warp][d] + (st[r] >> 9)] = (d << kDirOccBits) | (uint32_t)(wbase + r * 32 + lane);
        }
    }
    __syncthreads();          // also orders the zero stores of step 1 before the sums below

    // 6. the sums
    for (int q = 0; q < jb.n_out; ++q) {
        const DirectOut& o = jb.out[q];
        if (o.vec == 4)      reduce_rows<4>(o, ent, run_start, run_len, row0, n_rows, red);
        else if (o.vec == 2) reduce_rows<2>(o, ent, run_start, run_len, row0, n_rows, red);
        else                 reduce_rows<1>(o, ent, run_start, run_len, row0, n_rows, red);
    }
}

}  // namespace rk

extern "C" {

int rk_embgrad_direct_reduce(const rk_direct_table_t* tables, int n_tables, int32_t* err_flag,
                             rk_stream_t stream_) {
    using namespace rk;
    cudaStream_t s = (cudaStream_t)stream_;
    RK_CHECK_ARG(tables && n_tables >= 1 && n_tables <= RK_MAX_TABLES, "direct_reduce: n_tables=%d", n_tables);
    DirectParams p;
    memset(&p, 0, sizeof(p));
    int ctas = 0;
    auto flush = [&]() -> int {
        if (p.n_jobs == 0) return 0;
        direct_reduce_kernel<<<ctas, kDirThreads, 0, s>>>(p, err_flag);
        RK_LAUNCH_CHECK();
        memset(&p, 0, sizeof(p));
        ctas = 0;
        return 0;
    };
    for (int t = 0; t < n_tables; ++t) {
        const rk_direct_table_t& tb = tables[t];
        RK_CHECK_ARG(tb.dw && tb.rows > 0 && tb.dim > 0 && tb.dim <= 128 && tb.n >= 0, "direct_reduce: table %d rows=%lld dim=%d n=%lld",
                     t, (long long)tb.rows, tb.dim, (long long)tb.n);
        RK_CHECK_ARG(tb.n <= kDirMaxN, "direct_reduce: table %d has %lld occurrences (max %d: use rk_plan_build + "
                     "rk_embgrad_segment_reduce)", t, (long long)tb.n, kDirMaxN);
        RK_CHECK_ARG(tb.rows < (1ll << 31) && tb.ld < (1ll << 31), "direct_reduce: table %d too large", t);
        RK_CHECK_ARG(tb.n == 0 || (tb.idx && tb.g), "direct_reduce: table %d NULL idx / g", t);
        int v = tb.dim % 4 == 0 ? 4 : (tb.dim % 2 == 0 ? 2 : 1);
        while (v > 1 && ((tb.ld % v) != 0 || ((uintptr_t)tb.g % (4 * v)) != 0 || ((uintptr_t)tb.dw % (4 * v)) != 0))
            v >>= 1;
        RK_CHECK_ARG(tb.dim / v <= 32, "direct_reduce: table %d dim %d too wide", t, tb.dim);
        DirectOut o;
        o.g = tb.g; o.dw = tb.dw; o.ld = (int32_t)tb.ld; o.dim = (int16_t)tb.dim; o.vec = (int8_t)v;
        o.lanes = (int8_t)(tb.dim / v);
        // same index column, height and length as an earlier table of this launch: one more output of that job
        int j = -1;
        for (int k = 0; k < p.n_jobs; ++k)
            if (p.job[k].idx == tb.idx && p.job[k].rows == (int32_t)tb.rows && p.job[k].n == (int32_t)tb.n &&
                p.job[k].n_out < kDirMaxOut) { j = k; break; }
        if (j < 0) {
            if (p.n_jobs == kDirMaxJobs) { int rc = flush(); if (rc) return rc; }
            j = p.n_jobs++;
            p.job[j].idx = tb.idx;
            p.job[j].n = (int32_t)tb.n;
            p.job[j].rows = (int32_t)tb.rows;
            p.job[j].cta_start = ctas;
            ctas += (int)ceil_div(tb.rows, kDirRows);
        }
        p.job[j].out[p.job[j].n_out++] = o;
    }
    return flush();
}

}  // extern "C"
