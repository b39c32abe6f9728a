// plan_sort.cu — rk_plan_build: stable LSD radix sort of every index occurrence of a batch
// by (field, row).  Replaces the sort inside ATen's embedding_dense_backward (reached from
// loss.backward(), e.g. DeepFM/deepfm.py:170, DIN/din.py:346) with one composite-key sort for
// all tables of the model; the order it produces makes the later segment reduction sum each
// row's gradients in occurrence order (deterministic, atomics-free).
//
// HBM-bound integer work: 8-bit digits, 3 kernels per pass (tile histogram, scan, stable
// scatter).  Keys are uint32 = key_base[field] + row; payload = occurrence number in its field.
#include <string.h>
#include "common.cuh"

namespace rk {

constexpr int kSortThreads = 256;
constexpr int kSortWarps   = kSortThreads / 32;
constexpr int kSortRounds  = 16;
constexpr int kSortTile    = kSortThreads * kSortRounds;  // 4096 keys per CTA
constexpr int kBins        = 256;

struct KeyBuild {
    const int64_t* idx[RK_MAX_FIELDS];
    const int64_t* len[RK_MAX_FIELDS];        // sequence fields: per-sample length, else NULL
    int32_t        T[RK_MAX_FIELDS];          // positions per sample of a sequence field
    int32_t        mode[RK_MAX_FIELDS];       // RK_LIVE_*
    int64_t        start[RK_MAX_FIELDS + 1];  // first occurrence of field f in the flat order
    int64_t        rows[RK_MAX_FIELDS];
    uint32_t       key_base[RK_MAX_FIELDS];
    int32_t        F;
};

__global__ void __launch_bounds__(256)
build_keys_kernel(const __grid_constant__ KeyBuild kb, uint32_t* __restrict__ keys,
                  uint32_t* __restrict__ vals, int32_t* err_flag) {
    const int64_t n = kb.start[kb.F];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        int f = 0;
#pragma unroll 1
        while (f + 1 < kb.F && i >= kb.start[f + 1]) ++f;
        const int64_t o   = i - kb.start[f];
        int64_t row = checked_row(kb.idx[f][o], kb.rows[f], err_flag);
        if (kb.mode[f] != RK_LIVE_ALL) {
            // padded history positions carry no gradient: park them on the field's sentinel row
            const int64_t b = o / kb.T[f], t = o - b * kb.T[f];
            const int64_t l = kb.len[f][b];
            const bool dead = t >= l && !(kb.mode[f] == RK_LIVE_PREFIX_OR_EMPTY && l <= 0);
            if (dead) row = kb.rows[f];
        }
        keys[i] = kb.key_base[f] + (uint32_t)row;
        vals[i] = (uint32_t)o;
    }
}

__global__ void __launch_bounds__(kSortThreads)
tile_hist_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift, int n_tiles,
                 uint32_t* __restrict__ hist) {
    __shared__ uint32_t h[kBins];
    for (int i = threadIdx.x; i < kBins; i += kSortThreads) h[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll 4
    for (int r = 0; r < kSortRounds; ++r) {
        const int64_t i = base + r * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & (kBins - 1)], 1u);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < kBins; d += kSortThreads)
        hist[(int64_t)d * n_tiles + blockIdx.x] = h[d];
}

// Exclusive scan of m counters in place, one CTA.
__global__ void __launch_bounds__(1024)
scan_kernel(uint32_t* __restrict__ data, int64_t m) {
    __shared__ uint32_t warp_tot[32];
    const int     t    = threadIdx.x;
    const int64_t per  = (m + 1023) / 1024;
    const int64_t lo   = t * per;
    const int64_t hi   = lo + per < m ? lo + per : m;
    uint32_t      sum  = 0;
    for (int64_t i = lo; i < hi; ++i) sum += data[i];
    // block-wide exclusive scan of `sum`
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(kFull, inc, o);
        if ((t & 31) >= o) inc += v;
    }
    if ((t & 31) == 31) warp_tot[t >> 5] = inc;
    __syncthreads();
    if (t < 32) {
        uint32_t w = warp_tot[t], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(kFull, winc, o);
            if (t >= o) winc += v;
        }
        warp_tot[t] = winc - w;
    }
    __syncthreads();
    uint32_t run = warp_tot[t >> 5] + inc - sum;
    for (int64_t i = lo; i < hi; ++i) {
        uint32_t c = data[i];
        data[i]    = run;
        run += c;
    }
}

// Stable scatter of one 4096-key tile: ranks are taken in (warp, round, lane) = input order.
__global__ void __launch_bounds__(kSortThreads)
scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
               uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n,
               int shift, int n_tiles, const uint32_t* __restrict__ hist_scanned) {
    __shared__ uint32_t cnt[kSortWarps][kBins];
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    for (int i = t; i < kSortWarps * kBins; i += kSortThreads) (&cnt[0][0])[i] = 0;
    __syncthreads();

    const int64_t  wbase = (int64_t)blockIdx.x * kSortTile + (int64_t)w * (kSortRounds * 32);
    const unsigned lt    = (1u << lane) - 1u;
    uint32_t key[kSortRounds];
    uint16_t rank[kSortRounds];
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
        const int64_t i = wbase + r * 32 + lane;
        key[r] = i < n ? keys_in[i] : 0u;
    }
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
        const int64_t  i     = wbase + r * 32 + lane;
        const bool     valid = i < n;
        const uint32_t d     = (key[r] >> shift) & (kBins - 1);
        const uint32_t tag   = valid ? d : (kBins | lane);  // invalid lanes match nobody
        const unsigned m     = __match_any_sync(kFull, tag);
        const uint32_t old   = valid ? cnt[w][d] : 0u;
        __syncwarp();
        if (valid && (m & lt) == 0) cnt[w][d] = old + __popc(m);
        __syncwarp();
        rank[r] = (uint16_t)(old + __popc(m & lt));
    }
    __syncthreads();
    // per digit: turn the per-warp counts into global bases (exclusive over warps)
    for (int d = t; d < kBins; d += kSortThreads) {
        uint32_t run = hist_scanned[(int64_t)d * n_tiles + blockIdx.x];
#pragma unroll
        for (int ww = 0; ww < kSortWarps; ++ww) {
            uint32_t c = cnt[ww][d];
            cnt[ww][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
        const int64_t i = wbase + r * 32 + lane;
        if (i < n) {
            const uint32_t d   = (key[r] >> shift) & (kBins - 1);
            const uint32_t pos = cnt[w][d] + rank[r];
            keys_out[pos] = key[r];
            vals_out[pos] = vals_in[i];
        }
    }
}

static int n_tiles_of(int64_t n) { return (int)ceil_div(n > 0 ? n : 1, kSortTile); }

}  // namespace rk

extern "C" {

size_t rk_plan_workspace_bytes(int64_t n_total) {
    if (n_total < 0) n_total = 0;
    const size_t n     = (size_t)n_total;
    const size_t tiles = (size_t)rk::n_tiles_of(n_total);
    // ping-pong keys + vals, then the tile histograms; each region 256-byte aligned
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    return al(n * 4) * 2 + al(tiles * rk::kBins * 4);
}

int rk_plan_build(const int64_t* const* idx, const int64_t* n, const int64_t* rows, int F,
                  const int64_t* const* seq_len, const int32_t* seq_T, const int32_t* live_mode,
                  uint32_t* sorted_keys, uint32_t* perm, void* ws, size_t ws_bytes,
                  int32_t* err_flag, rk_stream_t stream_) {
    using namespace rk;
    cudaStream_t s = (cudaStream_t)stream_;
    RK_CHECK_ARG(F >= 1 && F <= RK_MAX_FIELDS, "rk_plan_build: F=%d outside [1,%d]", F,
                 RK_MAX_FIELDS);
    RK_CHECK_ARG(idx && n && rows, "rk_plan_build: NULL host array");
    KeyBuild kb;
    memset(&kb, 0, sizeof(kb));
    kb.F = F;
    int64_t total = 0, space = 0;
    for (int f = 0; f < F; ++f) {
        RK_CHECK_ARG(n[f] >= 0 && rows[f] > 0, "rk_plan_build: field %d n=%lld rows=%lld", f,
                     (long long)n[f], (long long)rows[f]);
        RK_CHECK_ARG(idx[f] != nullptr || n[f] == 0, "rk_plan_build: field %d idx is NULL", f);
        kb.idx[f]      = idx[f];
        kb.start[f]    = total;
        kb.rows[f]     = rows[f];
        kb.key_base[f] = (uint32_t)space;
        kb.mode[f]     = live_mode ? live_mode[f] : RK_LIVE_ALL;
        if (kb.mode[f] != RK_LIVE_ALL) {
            RK_CHECK_ARG(kb.mode[f] == RK_LIVE_PREFIX || kb.mode[f] == RK_LIVE_PREFIX_OR_EMPTY,
                         "rk_plan_build: field %d live_mode %d", f, kb.mode[f]);
            RK_CHECK_ARG(seq_len && seq_len[f] && seq_T && seq_T[f] > 0 && n[f] % seq_T[f] == 0,
                         "rk_plan_build: field %d needs lengths and T dividing n", f);
            kb.len[f] = seq_len[f];
            kb.T[f]   = seq_T[f];
        }
        total += n[f];
        space += rows[f] + 1;   // +1: the sentinel row of dead occurrences sorts last in the field
    }
    kb.start[F] = total;
    RK_CHECK_ARG(total < (1ll << 31) && space < (1ll << 32),
                 "rk_plan_build: %lld occurrences / %lld rows exceed the 32-bit key space",
                 (long long)total, (long long)space);
    if (total == 0) return 0;
    RK_CHECK_ARG(sorted_keys && perm && ws, "rk_plan_build: NULL output or workspace");
    RK_CHECK_ARG(ws_bytes >= rk_plan_workspace_bytes(total),
                 "rk_plan_build: workspace %zu < %zu bytes", ws_bytes,
                 rk_plan_workspace_bytes(total));

    int bits = 1;
    while (bits < 32 && (1ull << bits) < (unsigned long long)space) ++bits;
    const int passes = (bits + 7) / 8;

    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    char*     base  = (char*)ws;
    uint32_t* keysY = (uint32_t*)base;
    uint32_t* valsY = (uint32_t*)(base + al((size_t)total * 4));
    uint32_t* hist  = (uint32_t*)(base + 2 * al((size_t)total * 4));
    uint32_t *kin, *vin, *kout, *vout;
    if (passes & 1) { kin = keysY; vin = valsY; kout = sorted_keys; vout = perm; }
    else            { kin = sorted_keys; vin = perm; kout = keysY; vout = valsY; }

    const int tiles = n_tiles_of(total);
    {
        int grid = (int)ceil_div(total, 256);
        const int cap = sm_count() * 8;
        if (grid > cap) grid = cap;
        build_keys_kernel<<<grid, 256, 0, s>>>(kb, kin, vin, err_flag);
        RK_LAUNCH_CHECK();
    }
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        tile_hist_kernel<<<tiles, kSortThreads, 0, s>>>(kin, total, shift, tiles, hist);
        RK_LAUNCH_CHECK();
        scan_kernel<<<1, 1024, 0, s>>>(hist, (int64_t)tiles * kBins);
        RK_LAUNCH_CHECK();
        scatter_kernel<<<tiles, kSortThreads, 0, s>>>(kin, vin, kout, vout, total, shift, tiles,
                                                      hist);
        RK_LAUNCH_CHECK();
        uint32_t* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    return 0;
}

}  // extern "C"
