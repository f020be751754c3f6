// sfm_bucket.cu -- transposition + deterministic reduce-by-feature of the SGD step, bucket form
// (DESIGN.md 3.3).  Replaces the two-pass global radix sort + fm_pull_chunks + fm_pull_finalize
// on the hot path; those stay for the row-sharded model and as the SFM_SORT=cub / SFM_BUCKET=0
// cross-check.
//
// The feature id is split into a BUCKET (top HB <= 11 bits) and a LOCAL id (low LB bits, 2^LB
// features whose gradient accumulators fit in shared memory).  One global pass groups the batch's
// entries by bucket; the second half of the sort never touches HBM: the reduce kernel ranks every
// tile of a bucket by local id inside shared memory and walks the runs.
//
//   bkt_count_kernel    per contiguous row range c (one per scatter CTA): bucket histogram
//                       -> counts[c][bucket]
//   bkt_offsets_plan_kernel  column sums: counts[c][b] <- entries of bucket b in ranges < c;
//                       totals[b]; the CTA that finishes last builds the plan: bucket starts,
//                       work items (a bucket is cut into items of <= 32768 entries), the item
//                       table and the largest-first item order
//   bkt_scatter_kernel  persistent, one CTA per range, tiles of 8192 entries in order: stable
//                       rank of every entry inside the tile (peer masks from ballots over the
//                       digit bits, warp-private counters, prefix over warps and digits), entries
//                       re-ordered in shared memory, written as ONE packed word
//                       (local id << RB | batch row) (+ the value for non-binary data) to
//                       start(range, bucket) + running offset.  Ranges and tiles are visited in
//                       order, so rows ascend inside a bucket: the pass is a stable partition.
//   bkt_pull_kernel     persistent over work items (atomic ticket): per tile of 4096 entries
//                       stable rank by local id (same scheme), re-order into shared memory, walk
//                       the runs (level 0: a group of LPR lanes adds 8*LPR consecutive entries in
//                       order, gathering the rows' factor sums S_r; level 1: a run spanning
//                       several groups is summed by the group where it starts), add the run sums
//                       to the item's shared-memory accumulators (one owner per (tile, feature):
//                       no atomics).  Item end: a single-item bucket is finalised in place
//                       (g = A - v D, then the SGD update or the dense gradient); the items of a
//                       multi-item bucket leave their touched accumulator rows in scratch and
//                       the LAST item to finish adds them in item order.
//
// Every sum has a fixed shape that depends on entry positions only: bitwise reproducible, no float
// atomics (the only atomics are the integer ticket / arrival counters).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <type_traits>

#include "sfm_common.h"

namespace sfm {

#define FULL 0xffffffffu

constexpr int BK_MAX_HB = 11;
constexpr int BK_MAX_LB = 10;
#ifndef BK_SC_IPT
#define BK_SC_IPT 32
#endif
#ifndef BK_SC_CTAS
#define BK_SC_CTAS 2
#endif
#ifndef BK_SC_RANK
#define BK_SC_RANK 2      // scatter ranking: 0 leader byte table + ballots, 1 MATCH.ANY, 2 leader id packed
#endif                    // into the top bits of the warp's digit counter (measured 0.440 / 0.53 / 0.389 ms)
constexpr int SC_THREADS = 256;              // scatter: 8 warps x SC_IPT entries per thread
constexpr int SC_WARPS = SC_THREADS / 32;
constexpr int SC_IPT = BK_SC_IPT;
constexpr int SC_TILE = SC_THREADS * SC_IPT;   // 8192
constexpr int SC_RANK = BK_SC_RANK;
static_assert(SC_RANK != 2 || 32 * SC_IPT < 2048, "packed leader: strip counts must fit 11 bits");
#ifndef BK_PL_THREADS
#define BK_PL_THREADS 512
#endif
#ifndef BK_PL_IPT
#define BK_PL_IPT 8
#endif
#ifndef BK_ITEM_TILES
#define BK_ITEM_TILES 8
#endif
constexpr int PL_THREADS = BK_PL_THREADS;    // pull: 16 warps x 8 entries per thread
constexpr int PL_WARPS = PL_THREADS / 32;
constexpr int PL_IPT = BK_PL_IPT;
#ifndef BK_PL_CTAS
#define BK_PL_CTAS (1024 / BK_PL_THREADS)
#endif
constexpr int PL_CTAS_PER_SM = BK_PL_CTAS;   // launch bound (register budget); the grid uses the real occupancy
constexpr int PL_TILE = PL_THREADS * PL_IPT;   // 4096
static_assert(PL_IPT % 2 == 0 && 32 * PL_IPT <= 65536, "strip ranks are packed two per word");
constexpr int PL_ITEM_TILES = BK_ITEM_TILES;
constexpr int PL_ITEM = PL_TILE * PL_ITEM_TILES;   // 32768 entries per work item
constexpr int PL_ORDER_CLASSES = 64;               // size classes of the largest-first item order
#ifndef BK_PL_U
#define BK_PL_U 4
#endif
#ifndef BK_SPREAD
#define BK_SPREAD 1       // partial tiles are spread over all warps / lane groups (A/B knob)
#endif
constexpr int PL_U = BK_PL_U;                // S-row gathers in flight per lane
constexpr uint32_t NO_DIGIT = 0xFFFFFFFFu;
#ifndef BK_STAGE_MULT
#define BK_STAGE_MULT 0   // all-ones data: mult_r staged in shared memory with the entries (else gathered in the walk)
#endif
#ifndef BK_LDCG
#define BK_LDCG 1         // S-row gathers bypass L1 (ld.global.cg): no L1 line is allocated per gather
#endif
constexpr bool STAGE_MULT = BK_STAGE_MULT != 0;

__device__ __forceinline__ float4 ld_gather4(const float4* p) {
#if BK_LDCG
    return __ldcg(p);
#else
    return __ldg(p);
#endif
}

__device__ __forceinline__ int64_t ent_of(const int64_t* __restrict__ out_ptr, int64_t out_base,
                                          int m, int64_t row) {
    return out_ptr ? __ldg(out_ptr + row) - out_base : row * (int64_t)m;
}

// peers = lanes of the warp (among `valid`) that hold the same digit.  Every lane stores its lane
// id into the warp's byte table at its digit (one of the lanes sharing a digit wins), reads the
// winner back, and the lanes with the same winner find each other with 5 ballots over the winner's
// bits -- instead of one ballot per digit bit (up to 11).  The table is never cleared: every
// reader has just written its own slot.
#ifndef BK_MATCH
#define BK_MATCH (BK_SC_RANK == 1)
#endif
#ifndef BK_PL_RANK
#define BK_PL_RANK 1      // reduce kernel's tile ranking: 0 table, 1 MATCH.ANY (0.669 -> 0.638 ms), 2 packed leader
#endif
#define BK_MATCH_PULL (BK_PL_RANK == 1)
static_assert(BK_PL_RANK != 2 || 32 * PL_IPT < 2048, "packed leader: strip counts must fit 11 bits");
template <bool MATCH>
__device__ __forceinline__ uint32_t peer_mask(uint8_t* __restrict__ tab, uint32_t d, bool valid,
                                              int lane) {
    if (MATCH) {
        const uint32_t mm = __match_any_sync(FULL, valid ? d : 0xFFFFFFFFu);
        return valid ? mm : 0u;
    }
    if (valid) tab[d] = (uint8_t)lane;
    __syncwarp();
    const uint32_t leader = valid ? (uint32_t)tab[d] : 0u;
    uint32_t m = __ballot_sync(FULL, valid);
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        const uint32_t bit = (leader >> b) & 1u;
        const uint32_t bal = __ballot_sync(FULL, bit != 0u);
        m &= bal ^ (bit - 1u);   // bit ? bal : ~bal
    }
    __syncwarp();   // every read is done before the next round's writes
    return m;
}

// block-wide exclusive scan of one value per thread (THREADS a multiple of 32, <= 1024)
template <int THREADS>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* wsum /* [THREADS/32] */,
                                                    uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t excl = incl - v, tot = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) {
        const uint32_t s = wsum[w];
        if (w < warp) excl += s;
        tot += s;
    }
    if (total) *total = tot;
    return excl;
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
bkt_count_kernel(const uint32_t* __restrict__ keys, int n_rows, int m,
                 const int64_t* __restrict__ out_ptr, int64_t out_base, int LB, int NB,
                 uint32_t* __restrict__ counts) {
    extern __shared__ uint32_t bk_hist[];
    for (int d = threadIdx.x; d < NB; d += 512) bk_hist[d] = 0;
    __syncthreads();
    const int c = blockIdx.x, G = gridDim.x;
    const int64_t r0 = (int64_t)c * n_rows / G, r1 = (int64_t)(c + 1) * n_rows / G;
    const int64_t lo = ent_of(out_ptr, out_base, m, r0), hi = ent_of(out_ptr, out_base, m, r1);
    for (int64_t i0 = lo; i0 < hi; i0 += 512 * 8) {
        uint32_t k[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int64_t i = i0 + u * 512 + threadIdx.x;
            k[u] = i < hi ? __ldg(keys + i) : NO_DIGIT;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (k[u] != NO_DIGIT) atomicAdd(&bk_hist[k[u] >> LB], 1u);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < NB; d += 512) counts[(size_t)c * NB + d] = bk_hist[d];
}

// The same histogram taken ONE STEP AHEAD, straight from the resident data set (all-ones rows of m
// entries, sampled row numbers in row_ids): the feature ids of a batch do not depend on the model,
// so the counts and the plan of batch t + 1 are built on the low-priority copy stream while step t
// computes.  The batch size is only known on the device at launch time.  One warp
// per row (two rows in flight), same row ranges as bkt_scatter_kernel.
__global__ void __launch_bounds__(512)
bkt_count_rows_kernel(const int32_t* __restrict__ idx, const int32_t* __restrict__ row_ids,
                      const int32_t* __restrict__ n_rows_dev, int n_rows_cap, int m, int LB,
                      int NB, uint32_t* __restrict__ counts) {
    extern __shared__ uint32_t bk_hist[];
    for (int d = threadIdx.x; d < NB; d += 512) bk_hist[d] = 0;
    __syncthreads();
    // a draw larger than the plan's buffers were carved for counts nothing (the plan then holds
    // one empty item per bucket, inside its tables); the host sees the same and plans in the step
    const int n_dev = __ldg(n_rows_dev);
    const int n_rows = n_dev <= n_rows_cap ? n_dev : 0;
    const int c = blockIdx.x, G = gridDim.x;
    const int64_t r0 = (int64_t)c * n_rows / G, r1 = (int64_t)(c + 1) * n_rows / G;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t r = r0 + warp; r < r1; r += 32) {
        const int64_t rb = r + 16;
        const int32_t* pa = idx + (int64_t)__ldg(row_ids + r) * m;
        const int32_t* pb = rb < r1 ? idx + (int64_t)__ldg(row_ids + rb) * m : nullptr;
        for (int j = lane; j < m; j += 32) {
            const uint32_t ka = (uint32_t)__ldg(pa + j);
            const uint32_t kb = pb ? (uint32_t)__ldg(pb + j) : 0u;
            atomicAdd(&bk_hist[ka >> LB], 1u);
            if (pb) atomicAdd(&bk_hist[kb >> LB], 1u);
        }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < NB; d += 512) counts[(size_t)c * NB + d] = bk_hist[d];
}

// counts[c][d] <- sum over ranges c' < c; totals[d] = sum over all ranges.  Grid: NB/32 (>= 1)
// CTAs of 1024 threads = 32 bucket lanes x 32 blocks of ranges.  The CTA that finishes last
// (integer ticket, reset for the next launch) then builds the plan: bucket_off[b] (exclusive scan
// of totals), item_start[b] (exclusive scan of the items per bucket, >= 1 each so that untouched
// buckets still get their L2 decay); [NB] = grand totals; item_bucket[item] = its bucket;
// item_order = the items largest first; and it zeroes the reduce kernel's ticket / arrival words.
__global__ void __launch_bounds__(1024)
bkt_offsets_plan_kernel(uint32_t* __restrict__ counts, int G, int NB, uint32_t* __restrict__ totals,
                        unsigned int* __restrict__ ticket, uint32_t* __restrict__ bucket_off,
                        uint32_t* __restrict__ item_start, uint32_t* __restrict__ item_bucket,
                        uint32_t* __restrict__ item_order, uint32_t* __restrict__ pull_work) {
    __shared__ uint32_t part[32][33];
    __shared__ bool last;
    {
        const int dl = threadIdx.x & 31, rb = threadIdx.x >> 5;
        const int d = blockIdx.x * 32 + dl;
        const int per = (G + 31) / 32;
        const int c0 = min(G, rb * per), c1 = min(G, c0 + per);
        uint32_t sum = 0;
        if (d < NB)
            for (int c = c0; c < c1; ++c) sum += counts[(size_t)c * NB + d];
        part[rb][dl] = sum;
        __syncthreads();
        uint32_t run = 0;
        for (int r = 0; r < rb; ++r) run += part[r][dl];
        if (d < NB) {
            for (int c = c0; c < c1; ++c) {
                const uint32_t t = counts[(size_t)c * NB + d];
                counts[(size_t)c * NB + d] = run;
                run += t;
            }
            if (rb == 31) totals[d] = run;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();

    __shared__ uint32_t wsum[32];
    __shared__ uint32_t cls_cnt[PL_ORDER_CLASSES];
    for (int i = threadIdx.x; i <= NB; i += 1024) pull_work[i] = 0u;
    const int b0 = threadIdx.x * 2;
    const uint32_t t0 = b0 < NB ? __ldcg(totals + b0) : 0u, t1 = b0 + 1 < NB ? __ldcg(totals + b0 + 1) : 0u;
    const uint32_t n0 = b0 < NB ? max(1u, (t0 + PL_ITEM - 1) / PL_ITEM) : 0u;
    const uint32_t n1 = b0 + 1 < NB ? max(1u, (t1 + PL_ITEM - 1) / PL_ITEM) : 0u;
    uint32_t tot_e = 0, tot_i = 0;
    const uint32_t ex_e = block_excl_scan<1024>(t0 + t1, wsum, &tot_e);
    __syncthreads();
    const uint32_t ex_i = block_excl_scan<1024>(n0 + n1, wsum, &tot_i);
    if (b0 < NB) {
        bucket_off[b0] = ex_e;
        item_start[b0] = ex_i;
        for (uint32_t j = 0; j < n0; ++j) item_bucket[ex_i + j] = (uint32_t)b0;
    }
    if (b0 + 1 < NB) {
        bucket_off[b0 + 1] = ex_e + t0;
        item_start[b0 + 1] = ex_i + n0;
        for (uint32_t j = 0; j < n1; ++j) item_bucket[ex_i + n0 + j] = (uint32_t)(b0 + 1);
    }
    if (threadIdx.x == 0) {
        bucket_off[NB] = tot_e;
        item_start[NB] = tot_i;
        *ticket = 0u;
    }
    // item_order: the work items largest first (counting sort by size class), so that the
    // persistent CTAs of the reduce end together instead of one of them starting a 32 K-entry item
    // last.  The order inside a class comes from integer atomics and is free to vary: it only
    // schedules work, every sum keeps its fixed shape.
    for (int c = threadIdx.x; c < PL_ORDER_CLASSES; c += 1024) cls_cnt[c] = 0;
    __syncthreads();
    auto cls_of = [](uint32_t entries) -> int {   // class 0 = largest
        const int c = (int)((PL_ITEM - min(entries, (uint32_t)PL_ITEM)) / (PL_ITEM / PL_ORDER_CLASSES));
        return min(c, PL_ORDER_CLASSES - 1);
    };
    auto for_my_items = [&](auto fn) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int b = b0 + h;
            if (b >= NB) continue;
            const uint32_t t = h ? t1 : t0, n = h ? n1 : n0, first = h ? ex_i + n0 : ex_i;
            for (uint32_t j = 0; j < n; ++j) {
                const uint32_t lo = j * (uint32_t)PL_ITEM;
                const uint32_t sz = t > lo ? min((uint32_t)PL_ITEM, t - lo) : 0u;
                fn(first + j, sz);
            }
        }
    };
    for_my_items([&](uint32_t, uint32_t sz) { atomicAdd(&cls_cnt[cls_of(sz)], 1u); });
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int c = 0; c < PL_ORDER_CLASSES; ++c) {
            const uint32_t t = cls_cnt[c];
            cls_cnt[c] = run;
            run += t;
        }
    }
    __syncthreads();
    for_my_items([&](uint32_t item, uint32_t sz) { item_order[atomicAdd(&cls_cnt[cls_of(sz)], 1u)] = item; });
}

// ------------------------------------------------------------------------------------------
// Stable partition of the entries by bucket.  PAYMODE 0: all-ones rows of m entries in row
// order, batch row of input position i = i / m (magic division, exact for i < 2^31, m < 512);
// 1: batch rows in rows_in (uint32); 2: {batch row, x bits} in pay2 (uint2).
// ------------------------------------------------------------------------------------------
template <int PAYMODE>
__global__ void __launch_bounds__(SC_THREADS, BK_SC_CTAS)
bkt_scatter_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ rows_in,
                   const uint2* __restrict__ pay2, uint32_t* __restrict__ packed_out,
                   uint32_t* __restrict__ vals_out, int n_rows, int m,
                   const int64_t* __restrict__ out_ptr, int64_t out_base, int LB, int HB,
                   const uint32_t* __restrict__ colpre, const uint32_t* __restrict__ bucket_off,
                   unsigned long long magic) {
    constexpr bool HAS_VAL = PAYMODE == 2;
    extern __shared__ __align__(16) unsigned char sc_smem[];
    const int NB = 1 << HB;
    const int RB = 32 - LB;
    const uint32_t lowmask = (1u << LB) - 1u;
    uint32_t* ent_s = reinterpret_cast<uint32_t*>(sc_smem);                      // [TILE]
    uint32_t* val_s = ent_s + SC_TILE;                                            // [TILE] (HAS_VAL)
    uint16_t* dig_s = reinterpret_cast<uint16_t*>(val_s + (HAS_VAL ? SC_TILE : 0));   // [TILE]
    uint16_t* whist = dig_s + SC_TILE;                                            // [WARPS][NB]
    uint32_t* goff = reinterpret_cast<uint32_t*>(whist + SC_WARPS * NB);          // [NB]
    uint16_t* lstart = reinterpret_cast<uint16_t*>(goff + NB);                    // [NB]
    uint8_t* ltab = reinterpret_cast<uint8_t*>(lstart + NB) + (threadIdx.x >> 5) * NB;   // [WARPS][NB] (SC_RANK 0)
    __shared__ uint32_t wsum[SC_WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x, G = gridDim.x;
    const int64_t r0 = (int64_t)c * n_rows / G, r1 = (int64_t)(c + 1) * n_rows / G;
    const int lo = (int)ent_of(out_ptr, out_base, m, r0), hi = (int)ent_of(out_ptr, out_base, m, r1);

    // digits owned by this thread in the prefix phase: words [w0, w0 + cw) of row_words
    const int row_words = NB >= 2 ? NB / 2 : 1;
    const int cw = row_words >= SC_THREADS ? row_words / SC_THREADS : 1;   // <= 4
    const int w0 = tid * cw;
    const bool has = w0 < row_words;
    uint32_t R[8];   // running global offset of my digits (start of this range's run in the bucket)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        R[2 * j] = R[2 * j + 1] = 0;
        if (has && j < cw) {
            const int b = 2 * (w0 + j);
            if (b < NB) R[2 * j] = __ldg(bucket_off + b) + __ldg(colpre + (size_t)c * NB + b);
            if (b + 1 < NB) R[2 * j + 1] = __ldg(bucket_off + b + 1) + __ldg(colpre + (size_t)c * NB + b + 1);
        }
    }
    const unsigned lt = (1u << lane) - 1u;
    uint16_t* wh = whist + warp * NB;
    const int strip = warp * (32 * SC_IPT);

    for (int tile_base = lo; tile_base < hi; tile_base += SC_TILE) {
        const int n_valid = min(SC_TILE, hi - tile_base);
        {   // zero the warp histograms
            uint32_t* z = reinterpret_cast<uint32_t*>(whist);
            const int words = SC_WARPS * NB / 2;
            for (int i = tid; i < words; i += SC_THREADS) z[i] = 0;
            if (NB == 1 && tid < SC_WARPS) whist[tid] = 0;
        }
        uint32_t key[SC_IPT];
#pragma unroll
        for (int r = 0; r < SC_IPT; ++r) {
            const int p = strip + r * 32 + lane;
            key[r] = p < n_valid ? __ldg(keys + tile_base + p) : 0u;
        }
        __syncthreads();
        // ---- stable rank inside the warp's strip, in entry order
        uint32_t rk2[SC_IPT / 2];
#pragma unroll
        for (int r = 0; r < SC_IPT; ++r) {
            if (strip + r * 32 < n_valid) {   // warp-uniform
                const bool valid = strip + r * 32 + lane < n_valid;
                const uint32_t d = key[r] >> LB;
                uint32_t rank;
                if (SC_RANK == 2) {
                    // the warp's counter of digit d carries the count so far in its low 11 bits;
                    // every lane holding d stamps its lane id into the top 5 bits (one wins), the
                    // lanes that read the same winner back are peers (5 ballots over its bits)
                    const uint32_t cnt = valid ? ((uint32_t)wh[d] & 0x7ffu) : 0u;
                    if (valid) wh[d] = (uint16_t)(cnt | ((uint32_t)lane << 11));
                    __syncwarp();
                    const uint32_t leader = valid ? ((uint32_t)wh[d] >> 11) : 0u;
                    uint32_t pm = __ballot_sync(FULL, valid);
#pragma unroll
                    for (int b = 0; b < 5; ++b) {
                        const uint32_t bit = (leader >> b) & 1u;
                        const uint32_t bal = __ballot_sync(FULL, bit != 0u);
                        pm &= bal ^ (bit - 1u);
                    }
                    rank = cnt + __popc(pm & lt);
                    if (valid && (uint32_t)lane == leader) wh[d] = (uint16_t)((cnt + __popc(pm)) | (leader << 11));
                    __syncwarp();
                } else {
                    const uint32_t pm = peer_mask<SC_RANK == 1>(ltab, d, valid, lane);
                    uint32_t old = 0;
                    if (valid && (pm & lt) == 0u) {   // first lane of the group owns the counter
                        old = wh[d];
                        wh[d] = (uint16_t)(old + __popc(pm));
                    }
                    old = __shfl_sync(FULL, old, valid ? __ffs(pm) - 1 : lane);
                    rank = old + __popc(pm & lt);
                }
                if (r & 1) rk2[r / 2] |= rank << 16; else rk2[r / 2] = rank;
            } else {
                if (!(r & 1)) rk2[r / 2] = 0;
            }
        }
        __syncthreads();
        // ---- per digit: exclusive prefix over the warps (two u16 counters per word; a tile has
        // 8192 entries, so the halves never carry), then over the digits
        {
            uint32_t* wrows = reinterpret_cast<uint32_t*>(whist);
            uint32_t cnt2[4] = {0, 0, 0, 0};
            uint32_t local = 0;
            if (NB >= 2) {
                if (has) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (j < cw) {
                            uint32_t run = 0;
#pragma unroll
                            for (int w = 0; w < SC_WARPS; ++w) {
                                uint32_t t = wrows[w * row_words + w0 + j];
                                if (SC_RANK == 2) t &= 0x07ff07ffu;   // drop the leader stamps
                                wrows[w * row_words + w0 + j] = run;
                                run += t;
                            }
                            cnt2[j] = run;
                            local += (run & 0xffffu) + (run >> 16);
                        }
                    }
                }
            } else if (tid == 0) {   // single bucket
                uint32_t run = 0;
                for (int w = 0; w < SC_WARPS; ++w) {
                    const uint32_t t = SC_RANK == 2 ? (whist[w] & 0x7ffu) : whist[w];
                    whist[w] = (uint16_t)run;
                    run += t;
                }
                cnt2[0] = run;
                local = run;
            }
            const uint32_t excl0 = block_excl_scan<SC_THREADS>(local, wsum, nullptr);
            uint32_t excl = excl0;
            if (has) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (j < cw) {
                        const int b = 2 * (w0 + j);
                        const uint32_t c_lo = cnt2[j] & 0xffffu, c_hi = cnt2[j] >> 16;
                        if (b < NB) {
                            lstart[b] = (uint16_t)excl;
                            goff[b] = R[2 * j] - excl;
                            R[2 * j] += c_lo;
                        }
                        if (b + 1 < NB) {
                            lstart[b + 1] = (uint16_t)(excl + c_lo);
                            goff[b + 1] = R[2 * j + 1] - (excl + c_lo);
                            R[2 * j + 1] += c_hi;
                        }
                        excl += c_lo + c_hi;
                    }
                }
            }
        }
        __syncthreads();
        // ---- re-order inside the tile
#pragma unroll
        for (int r0i = 0; r0i < SC_IPT; r0i += 8) {
            uint32_t rowv[8], xv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int p = strip + (r0i + u) * 32 + lane;
                rowv[u] = 0;
                xv[u] = 0;
                if (p < n_valid) {
                    if (PAYMODE == 0)
                        rowv[u] = (uint32_t)(((unsigned long long)(uint32_t)(tile_base + p) * magic) >> 40);
                    else if (PAYMODE == 1)
                        rowv[u] = __ldg(rows_in + tile_base + p);
                    else {
                        const uint2 pl = __ldg(pay2 + tile_base + p);
                        rowv[u] = pl.x;
                        xv[u] = pl.y;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = r0i + u;
                const int p = strip + r * 32 + lane;
                if (p < n_valid) {
                    const uint32_t d = key[r] >> LB;
                    const uint32_t rank = (r & 1) ? (rk2[r / 2] >> 16) : (rk2[r / 2] & 0xffffu);
                    const int q = (int)lstart[d] + (int)wh[d] + (int)rank;
                    ent_s[q] = (LB ? (key[r] & lowmask) << RB : 0u) | rowv[u];
                    dig_s[q] = (uint16_t)d;
                    if (HAS_VAL) val_s[q] = xv[u];
                }
            }
        }
        __syncthreads();
        // ---- contiguous runs per bucket go out
#pragma unroll 8
        for (int j = 0; j < SC_IPT; ++j) {
            const int q = j * SC_THREADS + tid;
            if (q < n_valid) {
                const uint32_t g = goff[dig_s[q]] + (uint32_t)q;
                packed_out[g] = ent_s[q];
                if (HAS_VAL) vals_out[g] = val_s[q];
            }
        }
        // the next tile's prefix phase (which rewrites goff / lstart) and re-order (ent_s) come
        // after its own barriers, which every thread reaches only after this write-out
    }
}

// ------------------------------------------------------------------------------------------
// Reduce + update.
// ------------------------------------------------------------------------------------------
template <int LPR>
struct Pl {
    static constexpr int G = PL_THREADS / LPR;      // lane groups per CTA
    static constexpr int SUB = PL_TILE / G;         // consecutive sorted entries per group (8*LPR)
    static constexpr int SUBSHIFT = SUB == 4 ? 2 : SUB == 8 ? 3 : SUB == 16 ? 4 : SUB == 32 ? 5 : SUB == 64 ? 6 : SUB == 128 ? 7 : SUB == 256 ? 8 : 9;
    static_assert((1 << SUBSHIFT) == SUB, "SUB must be a power of two in [4, 512]");
    static constexpr int TILE_PAD = PL_TILE + G;    // phys(q) = q + (q >> SUBSHIFT)
    static constexpr int REC = LPR * 4 + 4;         // [A (kp) | D | C | 0 | 0]
};

template <int LPR>
__host__ __device__ constexpr size_t pull_union_bytes(int lb) {
    const size_t nbl = (size_t)1 << lb;
    const size_t heads = (size_t)Pl<LPR>::G * LPR * 16 + (size_t)Pl<LPR>::G * 8 + (size_t)Pl<LPR>::G * 4;
    const size_t rank = (size_t)PL_WARPS * nbl + (nbl + (nbl & 1)) * 2;
    return ((heads > rank ? heads : rank) + 15) / 16 * 16;
}

struct PullArgs {
    const uint32_t* packed;
    const uint32_t* vals;        // x bits per entry (non-binary) or nullptr
    const uint32_t* bucket_off;  // [NB+1]
    const uint32_t* item_start;  // [NB+1]
    const uint32_t* item_bucket; // [items]
    const uint32_t* item_order;  // [items] ticket -> item, largest items first
    uint32_t* work;              // [0] ticket, [1 + b] arrivals of bucket b (zeroed per step)
    float* part;                 // [items][2^LB][REC] partial accumulators of multi-item buckets
    uint32_t* part_bits;         // [items][2^LB / 32 (>= 1)] touched bitmaps
    const float4* S4;
    const float* mult;
    float4* V4;
    float* W;
    float* W0;
    float4* G4;                  // dense gradient out (not FUSED)
    float* Gw;
    float* Gw0;
    uint32_t* touch_bits;        // [ceil(n_slots / 32)] (MODE 2)
    const double* d_scal;
    const int32_t* err;
    int64_t n_slots;
    int NB, LB, k0, k1;
    UpdateParams up;
};

// MODE 0: SGD update in place (one GPU); 1: dense gradient [gV | gw | gw0]; 2: sparse gradient --
// rows are written only for the features this rank's batch touched, plus one bit per feature in
// touch_bits (the peer-memory exchange then moves touched rows only, sfm_p2p.cu).
template <int LPR, bool BINARY, int MODE>
__global__ void __launch_bounds__(PL_THREADS, PL_CTAS_PER_SM)
bkt_pull_kernel(const PullArgs a) {
    constexpr bool FUSED = MODE == 0;
    using C = Pl<LPR>;
    constexpr int G = C::G, SUB = C::SUB, SS = C::SUBSHIFT, REC = C::REC;
    extern __shared__ __align__(16) unsigned char pl_smem[];
    const int LB = a.LB, RB = 32 - LB, NBL = 1 << LB;
    const uint32_t rowmask = RB >= 32 ? 0xffffffffu : (1u << RB) - 1u;
    // shared-memory carve-up (pull_smem() mirrors it)
    float4* accA = reinterpret_cast<float4*>(pl_smem);                   // [NBL][LPR]
    // union U: {headA, headDC, nxt} (walk phase) / {ltab, lstart} (ranking phase; both are first
    // written after the tile's first barrier, i.e. after every owner has read the heads)
    unsigned char* U = reinterpret_cast<unsigned char*>(accA + (size_t)NBL * LPR);
    float4* headA = reinterpret_cast<float4*>(U);                         // [G][LPR]
    float2* headDC = reinterpret_cast<float2*>(headA + G * LPR);          // [G]
    int* nxt = reinterpret_cast<int*>(headDC + G);                        // [G] level-1 chain links
    uint8_t* ltab0 = U;                                                   // [WARPS][NBL]
    uint16_t* lstart = reinterpret_cast<uint16_t*>(U + (size_t)PL_WARPS * NBL);   // [NBL]
    float* accC = reinterpret_cast<float*>(U + pull_union_bytes<LPR>(LB));   // [NBL]
    float* accD = accC + NBL;                                             // [NBL]
    uint32_t* ent_s = reinterpret_cast<uint32_t*>(accD + NBL);            // [TILE_PAD]
    float* val_s = reinterpret_cast<float*>(ent_s + C::TILE_PAD);         // [TILE_PAD] mult_r or x
    uint32_t* tch = reinterpret_cast<uint32_t*>(val_s + ((BINARY && !STAGE_MULT) ? 0 : C::TILE_PAD));   // [NBL]
    uint16_t* whist = reinterpret_cast<uint16_t*>(tch + NBL);             // [WARPS][NBL]
    // one bit per sorted position of the tile: set where a run (a new local id) begins
    uint32_t* bflag = reinterpret_cast<uint32_t*>(whist + (size_t)PL_WARPS * NBL);   // [PL_TILE / 32]
    __shared__ uint32_t wsum[PL_WARPS];
    __shared__ uint32_t s_item, s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = tid / LPR, fq = tid % LPR;
    const unsigned lt = (1u << lane) - 1u;
    uint16_t* wh = whist + warp * NBL;
    uint8_t* ltab = ltab0 + warp * NBL;
    const uint32_t total_items = __ldg(a.item_start + a.NB);
    const double count = a.d_scal[SC_COUNT];
    const float inv = count > 0.0 ? (float)(1.0 / count) : 0.f;
    const bool active = count > 0.0 && !(FUSED && *a.err);

    for (;;) {
        __syncthreads();   // previous item fully finished (shared scalars, accumulators)
        if (tid == 0) s_item = atomicAdd(a.work, 1u);
        __syncthreads();
        if (s_item >= total_items) break;
        const uint32_t item = __ldg(a.item_order + s_item);
        const int b = (int)__ldg(a.item_bucket + item);
        const uint32_t it0 = __ldg(a.item_start + b), it1 = __ldg(a.item_start + b + 1);
        const uint32_t boff = __ldg(a.bucket_off + b), bend = __ldg(a.bucket_off + b + 1);
        const uint32_t e_lo = boff + (item - it0) * (uint32_t)PL_ITEM;
        const uint32_t e_hi = min(bend, e_lo + (uint32_t)PL_ITEM);

        uint32_t ent[PL_IPT];
        float xv[PL_IPT];
        // the bucket's parameter rows are needed at the item's end: ask for them now (L2), so
        // that the finalisation of a small item is not one DRAM round trip per feature
        if (MODE != 2) {   // (sparse gradient: only the touched rows are read at all)
            const int64_t f0 = (int64_t)b * NBL;
            const int64_t nf = min((int64_t)NBL, a.n_slots - f0);
            const char* vb = reinterpret_cast<const char*>(a.V4 + f0 * LPR);
            for (int64_t off = (int64_t)tid * 128; off < nf * (LPR * 16); off += (int64_t)PL_THREADS * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(vb + off));
        }
        // (the tile's entries are only live from here to the re-order: the NEXT tile is pulled
        // into L2 by prefetch instructions during the walk, not held in registers)
        for (int i = tid; i < NBL * LPR; i += PL_THREADS) accA[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = tid; i < NBL; i += PL_THREADS) {
            accC[i] = 0.f;
            accD[i] = 0.f;
            tch[i] = 0u;
        }

        for (uint32_t tb = e_lo; tb < e_hi; tb += PL_TILE) {
            const int n_valid = (int)min((uint32_t)PL_TILE, e_hi - tb);
            const bool full = n_valid == PL_TILE;
            // A partial tile (the tail of an item; every tile of a small mini-batch) is spread
            // over ALL warps and lane groups: shorter strips to rank, fewer sequential chunks to
            // walk.  strip_len: entries per warp (multiple of 32); ssd: log2(entries per group).
            int strip_len = 32 * PL_IPT, ssd = SS;
            if (!full && BK_SPREAD) {
                strip_len = ((n_valid + PL_WARPS - 1) / PL_WARPS + 31) & ~31;
                const int per_g = (n_valid + G - 1) / G;
                ssd = 2;
                while ((1 << ssd) < per_g) ++ssd;
            }
            const int strip = warp * strip_len;
            auto phys = [ssd](int q) { return q + (q >> ssd); };
#pragma unroll
            for (int r = 0; r < PL_IPT; ++r) {
                const int p = strip + r * 32 + lane;
                const bool in = r * 32 < strip_len && p < n_valid;
                ent[r] = in ? __ldg(a.packed + tb + p) : 0u;
                if (!BINARY) xv[r] = in ? __uint_as_float(__ldg(a.vals + tb + p)) : 0.f;
            }
            // all-ones data: the rows' multipliers (4 MB array, L2 resident) are fetched while
            // the tile is ranked and ride along into shared memory
            float second[PL_IPT];
#pragma unroll
            for (int r = 0; r < PL_IPT; ++r) {
                const int p = strip + r * 32 + lane;
                second[r] = BINARY ? ((STAGE_MULT && r * 32 < strip_len && p < n_valid)
                                          ? __ldg(a.mult + (ent[r] & rowmask)) : 0.f)
                                   : xv[r];
            }
            {   // zero the warp histograms
                uint32_t* z = reinterpret_cast<uint32_t*>(whist);
                const int words = PL_WARPS * NBL / 2;
                for (int i = tid; i < words; i += PL_THREADS) z[i] = 0;
                if (NBL == 1 && tid < PL_WARPS) whist[tid] = 0;
            }
            __syncthreads();   // histograms zero; the previous tile's walk is over
            if (tid < PL_TILE / 32) bflag[tid] = 0u;   // read next after two more barriers
            uint32_t rk2[PL_IPT / 2];   // ranks inside the warp's strip (< 32 * PL_IPT), two per word
#pragma unroll
            for (int r = 0; r < PL_IPT; ++r) {
                if (!(r & 1)) rk2[r / 2] = 0;
                if (r * 32 < strip_len && strip + r * 32 < n_valid) {   // warp-uniform
                    const bool valid = strip + r * 32 + lane < n_valid;
                    const uint32_t d = LB ? ent[r] >> RB : 0u;
                    uint32_t rank;
                    if (BK_PL_RANK == 2) {   // leader id in the top 5 bits of the counter (see bkt_scatter_kernel)
                        const uint32_t cnt = valid ? ((uint32_t)wh[d] & 0x7ffu) : 0u;
                        if (valid) wh[d] = (uint16_t)(cnt | ((uint32_t)lane << 11));
                        __syncwarp();
                        const uint32_t leader = valid ? ((uint32_t)wh[d] >> 11) : 0u;
                        uint32_t pm = __ballot_sync(FULL, valid);
#pragma unroll
                        for (int bb = 0; bb < 5; ++bb) {
                            const uint32_t bit = (leader >> bb) & 1u;
                            const uint32_t bal = __ballot_sync(FULL, bit != 0u);
                            pm &= bal ^ (bit - 1u);
                        }
                        rank = cnt + __popc(pm & lt);
                        if (valid && (uint32_t)lane == leader) wh[d] = (uint16_t)((cnt + __popc(pm)) | (leader << 11));
                        __syncwarp();
                    } else {
                        const uint32_t pm = peer_mask<BK_PL_RANK == 1>(ltab, d, valid, lane);
                        uint32_t old = 0;
                        if (valid && (pm & lt) == 0u) {   // first lane of the group owns the counter
                            old = wh[d];
                            wh[d] = (uint16_t)(old + __popc(pm));
                        }
                        old = __shfl_sync(FULL, old, valid ? __ffs(pm) - 1 : lane);
                        rank = old + __popc(pm & lt);
                    }
                    rk2[r / 2] |= rank << ((r & 1) * 16);
                }
            }
            __syncthreads();
            {   // exclusive prefix over the warps per digit, then over the digits
                const int row_words = NBL >= 2 ? NBL / 2 : 1;
                uint32_t* wrows = reinterpret_cast<uint32_t*>(whist);
                uint32_t cnt2 = 0, local = 0;
                const bool has = tid < row_words;   // row_words <= 512 = PL_THREADS
                if (NBL >= 2) {
                    if (has) {
                        uint32_t run = 0;
#pragma unroll
                        for (int w = 0; w < PL_WARPS; ++w) {
                            uint32_t t = wrows[w * row_words + tid];
                            if (BK_PL_RANK == 2) t &= 0x07ff07ffu;   // drop the leader stamps
                            wrows[w * row_words + tid] = run;
                            run += t;
                        }
                        cnt2 = run;
                        local = (run & 0xffffu) + (run >> 16);
                    }
                } else if (tid == 0) {
                    uint32_t run = 0;
                    for (int w = 0; w < PL_WARPS; ++w) {
                        const uint32_t t = BK_PL_RANK == 2 ? (whist[w] & 0x7ffu) : whist[w];
                        whist[w] = (uint16_t)run;
                        run += t;
                    }
                    cnt2 = run;
                    local = run;
                }
                const uint32_t excl = block_excl_scan<PL_THREADS>(local, wsum, nullptr);
                if (has) {
                    const int d0 = 2 * tid;
                    if (d0 < NBL) lstart[d0] = (uint16_t)excl;
                    if (d0 + 1 < NBL) lstart[d0 + 1] = (uint16_t)(excl + (cnt2 & 0xffffu));
                }
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < PL_IPT; ++r) {
                const int p = strip + r * 32 + lane;
                if (r * 32 < strip_len && p < n_valid) {
                    const uint32_t d = LB ? ent[r] >> RB : 0u;
                    const uint32_t within = (uint32_t)wh[d] + ((rk2[r / 2] >> ((r & 1) * 16)) & 0xffffu);   // rank inside the local id's run
                    const int ql = (int)lstart[d] + (int)within;
                    if (within == 0u) atomicOr(&bflag[ql >> 5], 1u << (ql & 31));   // <= 2^LB per tile
                    const int q = phys(ql);
                    ent_s[q] = ent[r];
                    if (!BINARY || STAGE_MULT) val_s[q] = second[r];
                }
            }
            __syncthreads();
            // the next tile's entries travel to L2 while this one is walked
            for (uint32_t i = tb + PL_TILE + (uint32_t)tid * 32u; i < min(e_hi, tb + 2u * PL_TILE);
                 i += (uint32_t)PL_THREADS * 32u) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.packed + i));
                if (!BINARY) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.vals + i));
            }

            // ---- level 0: walk my SUB sorted entries (group g owns sorted positions
            // [g * SUB, (g + 1) * SUB): physical words g * (SUB + 1) + j)
            const int subd = 1 << ssd;             // == SUB on full tiles
            const int q0 = g << ssd;
            const uint32_t* my_e = ent_s + g * (subd + 1);
            const float* my_v = val_s + g * (subd + 1);
            const char* Sq = reinterpret_cast<const char*>(a.S4 + fq);   // this lane's quarter of a row
            auto flush = [&](uint32_t d, const float4& A, float D, float Cc) {
                float4 t = accA[d * LPR + fq];
                t.x += A.x; t.y += A.y; t.z += A.z; t.w += A.w;
                accA[d * LPR + fq] = t;
                if (fq == 0) {
                    accC[d] += Cc;
                    if (!BINARY) accD[d] += D;
                    tch[d] = 1u;
                }
            };
            auto dig = [&](uint32_t e) -> uint32_t { return LB ? e >> RB : 0u; };
            uint32_t cur = (full || q0 < n_valid) ? dig(my_e[0]) : NO_DIGIT;
            bool cur_is_head = cur != NO_DIGIT && g > 0 && dig(my_e[-2]) == cur;
            float4 A = make_float4(0.f, 0.f, 0.f, 0.f);
            float D = 0.f, Cc = 0.f;
            auto walk = [&](auto full_tag) {
                constexpr bool FULLT = decltype(full_tag)::value;
                const int subn = subd;   // SUB on full tiles
#pragma unroll 1
                for (int base = 0; base < subn; base += PL_U) {
                    uint32_t ee[PL_U];
                    float cc[PL_U], xx[PL_U];
                    float4 sv[PL_U];
#pragma unroll
                    for (int u = 0; u < PL_U; ++u) {
                        const bool ok = FULLT || q0 + base + u < n_valid;
                        const uint32_t e = ok ? my_e[base + u] : 0u;
                        const uint32_t row = e & rowmask;
                        const float sec = (BINARY && !STAGE_MULT) ? (ok ? __ldg(a.mult + row) : 0.f)
                                                                   : (ok ? my_v[base + u] : 0.f);
                        ee[u] = e;
                        if (BINARY) {
                            cc[u] = sec;
                            xx[u] = 1.f;
                        } else {
                            xx[u] = sec;
                            cc[u] = ok ? __ldg(a.mult + row) * sec : 0.f;
                        }
                        // row * LPR + fq < 2^32: checked by the caller (train_core)
                        sv[u] = ok ? ld_gather4(reinterpret_cast<const float4*>(
                                         Sq + (size_t)row * (size_t)(LPR * 16)))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    // no run starts inside this chunk (the common case on skewed ids, where runs
                    // are long): plain accumulation, no digit compares
                    const int qb = q0 + base;
                    const uint32_t starts = FULLT ? ((bflag[qb >> 5] >> (qb & 31)) & ((1u << PL_U) - 1u)) : 1u;
                    if (starts == 0u) {
#pragma unroll
                        for (int u = 0; u < PL_U; ++u) {
                            const float c = cc[u];
                            A.x = fmaf(c, sv[u].x, A.x);
                            A.y = fmaf(c, sv[u].y, A.y);
                            A.z = fmaf(c, sv[u].z, A.z);
                            A.w = fmaf(c, sv[u].w, A.w);
                            D = BINARY ? D : fmaf(c, xx[u], D);
                            Cc += c;
                        }
                        continue;
                    }
#pragma unroll
                    for (int u = 0; u < PL_U; ++u) {
                        const bool ok = FULLT || q0 + base + u < n_valid;
                        const uint32_t du = ok ? dig(ee[u]) : NO_DIGIT;
                        if (du != cur) {
                            if (cur != NO_DIGIT) {
                                if (cur_is_head) {
                                    headA[g * LPR + fq] = A;
                                    if (fq == 0) headDC[g] = make_float2(D, Cc);
                                } else {
                                    flush(cur, A, D, Cc);
                                }
                            }
                            cur = du;
                            cur_is_head = false;
                            A = make_float4(0.f, 0.f, 0.f, 0.f);
                            D = 0.f;
                            Cc = 0.f;
                        }
                        const float c = cc[u];   // 0 for slots past the end of a partial tile
                        A.x = fmaf(c, sv[u].x, A.x);
                        A.y = fmaf(c, sv[u].y, A.y);
                        A.z = fmaf(c, sv[u].z, A.z);
                        A.w = fmaf(c, sv[u].w, A.w);
                        D = BINARY ? D : fmaf(c, xx[u], D);
                        Cc += c;
                    }
                }
            };
            // a group whose entries all exist takes the unpredicated walk (with the compare-free
            // chunks), also on a partial tile; only the group that straddles the tile's end checks
            if (full || q0 + subd <= n_valid) walk(std::true_type{}); else walk(std::false_type{});
            // ---- level 1: runs that span several groups.  A group whose FIRST run continues
            // the previous group's last one holds a "head" partial; a head that fills its whole
            // group and continues links to the next group's head.  The chains are summed by
            // pointer jumping (fixed tree shape: depends on the run's position only), then the
            // group where the run starts adds the chain that follows it.
            const bool open = cur != NO_DIGIT;
            const bool continues = open && g < G - 1 && q0 + subd < n_valid && dig(my_e[subd + 1]) == cur;
            const bool owner = open && !cur_is_head;
            if (open && cur_is_head) {
                headA[g * LPR + fq] = A;   // the whole group is a middle / final piece of a run
                if (fq == 0) headDC[g] = make_float2(D, Cc);
            }
            if (fq == 0) nxt[g] = (open && cur_is_head && continues) ? g + 1 : -1;
            __syncthreads();
            for (;;) {
                const int n = nxt[g];
                float4 addA = make_float4(0.f, 0.f, 0.f, 0.f);
                float2 addDC = make_float2(0.f, 0.f);
                int n2 = -1;
                if (n >= 0) {
                    addA = headA[n * LPR + fq];
                    addDC = headDC[n];
                    n2 = nxt[n];
                }
                if (!__syncthreads_or(n >= 0)) break;   // also: every read is done
                if (n >= 0) {
                    float4 t = headA[g * LPR + fq];
                    t.x += addA.x; t.y += addA.y; t.z += addA.z; t.w += addA.w;
                    headA[g * LPR + fq] = t;
                    if (fq == 0) {
                        const float2 dc = headDC[g];
                        headDC[g] = make_float2(dc.x + addDC.x, dc.y + addDC.y);
                        nxt[g] = n2;
                    }
                }
                __syncthreads();
            }
            if (owner) {
                if (continues) {
                    const float4 h4 = headA[(g + 1) * LPR + fq];
                    const float2 dc = headDC[g + 1];
                    A.x += h4.x; A.y += h4.y; A.z += h4.z; A.w += h4.w;
                    D += dc.x;
                    Cc += dc.y;
                }
                flush(cur, A, D, Cc);
            }
            // next tile: its first barrier orders these shared-memory updates
        }
        __syncthreads();

        // ---- item end
        const bool single = it1 - it0 == 1u;
        const int bw = NBL >= 32 ? NBL / 32 : 1;
        if (!single) {
            float* P = a.part + (size_t)item * NBL * REC;
            for (int d = g; d < NBL; d += G) {
                if (tch[d]) {
                    reinterpret_cast<float4*>(P + (size_t)d * REC)[fq] = accA[d * LPR + fq];
                    if (fq == 0)
                        reinterpret_cast<float4*>(P + (size_t)d * REC)[LPR] =
                            make_float4(accD[d], accC[d], 0.f, 0.f);
                }
            }
            for (int wd = tid; wd < bw; wd += PL_THREADS) {
                uint32_t bits = 0;
                for (int j = 0; j < 32; ++j)
                    if (wd * 32 + j < NBL && tch[wd * 32 + j]) bits |= 1u << j;
                a.part_bits[(size_t)item * bw + wd] = bits;
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                const uint32_t prev = atomicAdd(a.work + 1 + b, 1u);
                s_last = prev == it1 - it0 - 1u;
            }
            __syncthreads();
            if (!s_last) continue;
            __threadfence();
        }
        // finalize the bucket's features: one lane group per feature
        if (b == 0 && tid == 0) {
            const float g0 = (float)a.d_scal[SC_GW0];
            if (FUSED) {
                if (a.k0 && active) *a.W0 = sgd_step(*a.W0, g0, inv, a.up.eta, a.up.reg0);
            } else {
                *a.Gw0 = a.k0 ? g0 : 0.f;
            }
        }
        // four features per lane group at a time: their parameter rows are requested together,
        // before the first of them is updated (the stores would otherwise order the loads)
        for (int d0 = g; d0 < NBL; d0 += 4 * G) {
          float4 vv[4];
          float ww[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
              const int dj = d0 + j * G;
              const int64_t fj = (int64_t)b * NBL + dj;
              bool in = dj < NBL && fj < a.n_slots;
              // sparse gradient: a feature this rank did not touch writes nothing -- skip its row
              if (MODE == 2 && single && in) in = tch[dj] != 0u;
              vv[j] = in ? a.V4[fj * LPR + fq] : make_float4(0.f, 0.f, 0.f, 0.f);
              ww[j] = (in && fq == 0 && a.k1) ? a.W[fj] : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int d = d0 + j * G;
            const int64_t f = (int64_t)b * NBL + d;
            if (d >= NBL || f >= a.n_slots) continue;
            float4 v = vv[j];
            const float wi = ww[j];
            float4 A;
            float D, Cc;
            bool touched;
            if (single) {
                A = accA[d * LPR + fq];
                D = accD[d];
                Cc = accC[d];
                touched = tch[d] != 0u;
            } else {
                A = make_float4(0.f, 0.f, 0.f, 0.f);
                D = 0.f;
                Cc = 0.f;
                touched = false;
                for (uint32_t it = it0; it < it1; ++it) {   // item order: fixed summation order
                    const uint32_t bits = __ldcg(a.part_bits + (size_t)it * bw + (d >> 5));
                    if ((bits >> (d & 31)) & 1u) {
                        touched = true;
                        const float* P = a.part + ((size_t)it * NBL + d) * REC;
                        const float4 pa = __ldcg(reinterpret_cast<const float4*>(P) + fq);
                        const float4 dc = __ldcg(reinterpret_cast<const float4*>(P) + LPR);
                        A.x += pa.x; A.y += pa.y; A.z += pa.z; A.w += pa.w;
                        D += dc.x;
                        Cc += dc.y;
                    }
                }
            }
            if (BINARY) D = Cc;
            float4 gr;
            gr.x = __fmaf_rn(-v.x, D, A.x);
            gr.y = __fmaf_rn(-v.y, D, A.y);
            gr.z = __fmaf_rn(-v.z, D, A.z);
            gr.w = __fmaf_rn(-v.w, D, A.w);
            if (FUSED) {
                if (active) {
                    v.x = sgd_step(v.x, gr.x, inv, a.up.eta, a.up.regv);
                    v.y = sgd_step(v.y, gr.y, inv, a.up.eta, a.up.regv);
                    v.z = sgd_step(v.z, gr.z, inv, a.up.eta, a.up.regv);
                    v.w = sgd_step(v.w, gr.w, inv, a.up.eta, a.up.regv);
                    a.V4[f * LPR + fq] = v;
                    if (fq == 0 && a.k1) a.W[f] = sgd_step(wi, Cc, inv, a.up.eta, a.up.regw);
                }
            } else if (MODE == 1 || touched) {
                a.G4[f * LPR + fq] = gr;
                if (fq == 0) a.Gw[f] = a.k1 ? Cc : 0.f;
            }
            if (MODE == 2 && !single && fq == 0) tch[d] = touched ? 1u : 0u;
          }
        }
        if (MODE == 2) {   // the bucket's touched bitmap (buckets are multiples of 32 features)
            __syncthreads();
            for (int wd = tid; wd < bw; wd += PL_THREADS) {
                const int64_t f0 = (int64_t)b * NBL + wd * 32;
                if (f0 >= a.n_slots) break;
                uint32_t bits = 0;
                for (int j = 0; j < 32; ++j)
                    if (wd * 32 + j < NBL && f0 + j < a.n_slots && tch[wd * 32 + j]) bits |= 1u << j;
                a.touch_bits[f0 >> 5] = bits;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Host side.
// ------------------------------------------------------------------------------------------
static inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }

static bool model_geometry(const ModelView& m, int key_bits, int* lb_out, int* hb_out) {
    if (!knobs().bucket) return false;
    if (key_bits < 1) key_bits = 1;
    int lb = 0;
    while (lb < BK_MAX_LB && ((size_t)2 << lb) * (size_t)(m.kp + 2) * 4 <= 37000) ++lb;
    if (lb > key_bits - 1) lb = key_bits - 1;   // at least one bucket bit
    if (lb < 0) lb = 0;
    const int hb = key_bits - lb;
    if (hb < 1 || hb > BK_MAX_HB) return false;
    *lb_out = lb;
    *hb_out = hb;
    return true;
}

bool bucket_sparse_capable(const ModelView& m, int key_bits) {
    int lb, hb;
    return model_geometry(m, key_bits, &lb, &hb) && lb >= 5;
}

bool bucket_geometry(const ModelView& m, int key_bits, int64_t n_rows, int64_t nnz, BucketGeom* g) {
    int lb, hb;
    if (!model_geometry(m, key_bits, &lb, &hb)) return false;
    if (n_rows <= 0 || nnz <= 0) return false;
    if (n_rows >= ((int64_t)1 << (32 - lb)) || nnz >= 2147483647LL - 2 * SC_TILE) return false;
    g->LB = lb;
    g->HB = hb;
    g->NB = 1 << hb;
    g->max_items = (int64_t)g->NB + nnz / PL_ITEM + 1;
    return true;
}

size_t bucket_tables_bytes(const BucketGeom& g) {   // bucket_off | item_start | item_bucket | item_order
    return al256(sizeof(uint32_t) * (size_t)(g.NB + 1)) * 2 + al256(sizeof(uint32_t) * (size_t)g.max_items) * 2;
}

struct TablePtrs {
    uint32_t *bucket_off, *item_start, *item_bucket, *item_order;
};

static TablePtrs carve_tables(const BucketGeom& g, const void* base) {
    unsigned char* p = static_cast<unsigned char*>(const_cast<void*>(base));
    TablePtrs t;
    t.bucket_off = reinterpret_cast<uint32_t*>(p);
    p += al256(sizeof(uint32_t) * (size_t)(g.NB + 1));
    t.item_start = reinterpret_cast<uint32_t*>(p);
    p += al256(sizeof(uint32_t) * (size_t)(g.NB + 1));
    t.item_bucket = reinterpret_cast<uint32_t*>(p);
    p += al256(sizeof(uint32_t) * (size_t)g.max_items);
    t.item_order = reinterpret_cast<uint32_t*>(p);
    return t;
}

size_t bucket_work_bytes(const ModelView& m, const BucketGeom& g, int sm_count) {
    const int G = 2 * sm_count;
    const size_t nbl = (size_t)1 << g.LB;
    size_t b = al256(sizeof(uint32_t) * (size_t)G * g.NB);          // counts / column prefixes
    b += al256(sizeof(uint32_t) * (size_t)g.NB);                      // totals
    b += al256(sizeof(uint32_t) * (size_t)(g.NB + 1));                // ticket + arrivals
    b += al256(sizeof(float) * (size_t)g.max_items * nbl * (m.kp + 4));
    b += al256(sizeof(uint32_t) * (size_t)g.max_items * (nbl >= 32 ? nbl / 32 : 1));
    return b;
}

struct WorkPtrs {
    uint32_t *counts, *totals, *work, *part_bits;
    float* part;
};

static WorkPtrs carve_work(const ModelView& m, const BucketGeom& g, int sm_count, void* base) {
    const int G = 2 * sm_count;
    const size_t nbl = (size_t)1 << g.LB;
    unsigned char* p = static_cast<unsigned char*>(base);
    WorkPtrs w;
    w.counts = reinterpret_cast<uint32_t*>(p);
    p += al256(sizeof(uint32_t) * (size_t)G * g.NB);
    w.totals = reinterpret_cast<uint32_t*>(p);
    p += al256(sizeof(uint32_t) * (size_t)g.NB);
    w.work = reinterpret_cast<uint32_t*>(p);
    p += al256(sizeof(uint32_t) * (size_t)(g.NB + 1));
    w.part = reinterpret_cast<float*>(p);
    p += al256(sizeof(float) * (size_t)g.max_items * nbl * (m.kp + 4));
    w.part_bits = reinterpret_cast<uint32_t*>(p);
    return w;
}

static size_t scatter_smem(int hb, bool has_val) {
    const size_t nb = (size_t)1 << hb;
    return (size_t)SC_TILE * 4 * (has_val ? 2 : 1) + (size_t)SC_TILE * 2 + (size_t)SC_WARPS * nb * 2 +
           nb * 4 + nb * 2 + (SC_RANK == 0 ? (size_t)SC_WARPS * nb : 0) + 64;
}

// counts + plan of a batch of resident all-ones rows, from the row numbers (any stream; the batch
// size is read on the device)
cudaError_t bucket_count_plan_rows(const ModelView& m, const BucketGeom& g, const int32_t* idx,
                                   const int32_t* row_ids, const int32_t* n_rows_dev, int n_rows_cap,
                                   int uniform_m, void* work, void* tables, unsigned int* ticket, int sm_count,
                                   cudaStream_t st, int64_t* launches) {
    if (!idx || !row_ids || !n_rows_dev || uniform_m < 1) return cudaErrorInvalidValue;
    const WorkPtrs w = carve_work(m, g, sm_count, work);
    const TablePtrs tp = carve_tables(g, tables);
    const int G = 2 * sm_count;
    bkt_count_rows_kernel<<<G, 512, sizeof(uint32_t) * g.NB, st>>>(idx, row_ids, n_rows_dev, n_rows_cap,
                                                                    uniform_m, g.LB, g.NB, w.counts);
    bkt_offsets_plan_kernel<<<(g.NB + 31) / 32, 1024, 0, st>>>(w.counts, G, g.NB, w.totals, ticket,
                                                               tp.bucket_off, tp.item_start, tp.item_bucket,
                                                               tp.item_order, w.work);
    *launches += 2;
    return cudaGetLastError();
}

cudaError_t bucket_transpose(const ModelView& m, const BatchView& b, const BucketGeom& g,
                             const uint32_t* keys, const uint2* pay, int implicit_div, void* work,
                             void* tables, uint32_t* packed, uint32_t* vals, unsigned int* ticket,
                             int sm_count, cudaStream_t st, int64_t* launches) {
    const WorkPtrs w = carve_work(m, g, sm_count, work);
    const TablePtrs tp = carve_tables(g, tables);
    const int G = 2 * sm_count;
    const int mm = b.uniform_m >= 0 ? b.uniform_m : 0;
    const int64_t* optr = b.uniform_m >= 0 && !b.out_ptr ? nullptr : b.out_ptr;
    if (!optr && b.uniform_m < 0) return cudaErrorInvalidValue;
    bkt_count_kernel<<<G, 512, sizeof(uint32_t) * g.NB, st>>>(keys, (int)b.n_rows, mm, optr,
                                                               b.out_base, g.LB, g.NB, w.counts);
    bkt_offsets_plan_kernel<<<(g.NB + 31) / 32, 1024, 0, st>>>(w.counts, G, g.NB, w.totals, ticket,
                                                               tp.bucket_off, tp.item_start, tp.item_bucket,
                                                               tp.item_order, w.work);
    *launches += 2;
    return bucket_scatter(m, b, g, keys, pay, implicit_div, work, tables, packed, vals, sm_count, st,
                          launches);
}

// the stable partition alone: counts / plan already sit in `work` / `tables` (built inline by
// bucket_transpose or one step ahead by bucket_count_plan_rows)
cudaError_t bucket_scatter(const ModelView& m, const BatchView& b, const BucketGeom& g,
                           const uint32_t* keys, const uint2* pay, int implicit_div, void* work,
                           const void* tables, uint32_t* packed, uint32_t* vals, int sm_count,
                           cudaStream_t st, int64_t* launches) {
    const WorkPtrs w = carve_work(m, g, sm_count, work);
    const TablePtrs tp = carve_tables(g, tables);
    const uint32_t* bucket_off = tp.bucket_off;
    const int G = 2 * sm_count;
    const bool has_val = b.val != nullptr;
    const int mm = b.uniform_m >= 0 ? b.uniform_m : 0;
    const int64_t* optr = b.uniform_m >= 0 && !b.out_ptr ? nullptr : b.out_ptr;
    if (!optr && b.uniform_m < 0) return cudaErrorInvalidValue;
    const size_t smem = scatter_smem(g.HB, has_val);
    const unsigned long long magic =
        implicit_div ? ((1ULL << 40) + (unsigned long long)implicit_div - 1) / (unsigned long long)implicit_div : 0ULL;
    cudaError_t e;
#define SC_LAUNCH(PM)                                                                           \
    {                                                                                           \
        static size_t set_smem = 0;                                                             \
        static int set_dev = -1;                                                                \
        int dev_now = 0;                                                                        \
        cudaGetDevice(&dev_now);                                                                \
        if (set_smem != smem || set_dev != dev_now) {                                           \
            set_dev = dev_now;                                                                  \
            e = cudaFuncSetAttribute(bkt_scatter_kernel<PM>,                                    \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
            if (e != cudaSuccess) return e;                                                     \
            set_smem = smem;                                                                    \
        }                                                                                       \
    }                                                                                           \
    bkt_scatter_kernel<PM><<<G, SC_THREADS, smem, st>>>(                                        \
        keys, reinterpret_cast<const uint32_t*>(pay), pay, packed, vals, (int)b.n_rows, mm, optr, \
        b.out_base, g.LB, g.HB, w.counts, bucket_off, magic)
    if (has_val) {
        SC_LAUNCH(2);
    } else if (implicit_div) {
        SC_LAUNCH(0);
    } else {
        SC_LAUNCH(1);
    }
#undef SC_LAUNCH
    *launches += 1;
    return cudaGetLastError();
}

template <int LPR>
static size_t pull_smem(int lb, bool binary) {
    using C = Pl<LPR>;
    const size_t nbl = (size_t)1 << lb;
    return nbl * LPR * 16 + pull_union_bytes<LPR>(lb) + nbl * 4 * 2 +
           (size_t)C::TILE_PAD * 4 * ((binary && !STAGE_MULT) ? 1 : 2) + nbl * 4 +
           (size_t)PL_WARPS * nbl * 2 + PL_TILE / 8 + 64;
}

template <int LPR>
static cudaError_t pull_dispatch2(const ModelView& m, const BucketGeom& g, const PullArgs& a,
                                  bool binary, int mode, int sm_count, cudaStream_t st) {
    const size_t smem = pull_smem<LPR>(g.LB, binary);
    cudaError_t e;
    int occ = 0;   // resident CTAs per SM: the grid of the persistent kernel
#define PL_LAUNCH(B, M)                                                                          \
    {                                                                                            \
        static size_t set_smem = 0;   /* attribute + occupancy: queried once per kernel, size, device */ \
        static int set_occ = 0, set_dev = -1;                                                    \
        int dev_now = 0;                                                                         \
        cudaGetDevice(&dev_now);                                                                 \
        if (set_smem != smem || set_occ < 1 || set_dev != dev_now) {                             \
            set_dev = dev_now;                                                                   \
            e = cudaFuncSetAttribute(bkt_pull_kernel<LPR, B, M>,                                 \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
            if (e != cudaSuccess) return e;                                                      \
            occ = 0;                                                                             \
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bkt_pull_kernel<LPR, B, M>,  \
                                                              PL_THREADS, smem);                 \
            if (e != cudaSuccess) return e;                                                      \
            set_occ = occ < 1 ? 1 : occ;                                                         \
            set_smem = smem;                                                                     \
        }                                                                                        \
        occ = set_occ;                                                                           \
    }                                                                                            \
    bkt_pull_kernel<LPR, B, M><<<occ * sm_count, PL_THREADS, smem, st>>>(a)
    if (binary) {
        if (mode == 0) { PL_LAUNCH(true, 0); } else if (mode == 1) { PL_LAUNCH(true, 1); } else { PL_LAUNCH(true, 2); }
    } else {
        if (mode == 0) { PL_LAUNCH(false, 0); } else if (mode == 1) { PL_LAUNCH(false, 1); } else { PL_LAUNCH(false, 2); }
    }
#undef PL_LAUNCH
    return cudaGetLastError();
}

cudaError_t bucket_pull(const ModelView& m, const BucketGeom& g, const uint32_t* packed,
                        const uint32_t* vals, const void* tables, void* work, const float* S,
                        const float* mult, const double* d_scal, const int32_t* d_err,
                        UpdateParams up, bool fused, float* grad, uint32_t* touch_bits,
                        bool work_zeroed, int sm_count, cudaStream_t st, int64_t* launches) {
    const WorkPtrs w = carve_work(m, g, sm_count, work);
    if (touch_bits && g.LB < 5) return cudaErrorInvalidValue;   // bitmap words must not straddle buckets
    const int mode = fused ? 0 : (touch_bits ? 2 : 1);
    if (!work_zeroed) {   // (a transposition built in this step left the ticket / arrival words zero)
        cudaError_t e = cudaMemsetAsync(w.work, 0, sizeof(uint32_t) * (size_t)(g.NB + 1), st);
        if (e != cudaSuccess) return e;
    }
    PullArgs a;
    a.packed = packed;
    a.vals = vals;
    const TablePtrs tp = carve_tables(g, tables);
    a.bucket_off = tp.bucket_off;
    a.item_start = tp.item_start;
    a.item_bucket = tp.item_bucket;
    a.item_order = tp.item_order;
    a.work = w.work;
    a.part = w.part;
    a.part_bits = w.part_bits;
    a.S4 = reinterpret_cast<const float4*>(S);
    a.mult = mult;
    a.V4 = reinterpret_cast<float4*>(m.v);
    a.W = m.w;
    a.W0 = m.w0;
    a.G4 = reinterpret_cast<float4*>(grad);
    a.Gw = grad ? grad + m.n_slots * m.kp : nullptr;
    a.Gw0 = grad ? a.Gw + m.n_slots : nullptr;
    a.touch_bits = touch_bits;
    a.d_scal = d_scal;
    a.err = d_err;
    a.n_slots = m.n_slots;
    a.NB = g.NB;
    a.LB = g.LB;
    a.k0 = m.k0;
    a.k1 = m.k1;
    a.up = up;
    const bool binary = vals == nullptr;
    *launches += 1;
    switch (m.lpr) {
        case 1: return pull_dispatch2<1>(m, g, a, binary, mode, sm_count, st);
        case 2: return pull_dispatch2<2>(m, g, a, binary, mode, sm_count, st);
        case 4: return pull_dispatch2<4>(m, g, a, binary, mode, sm_count, st);
        case 8: return pull_dispatch2<8>(m, g, a, binary, mode, sm_count, st);
        case 16: return pull_dispatch2<16>(m, g, a, binary, mode, sm_count, st);
        case 32: return pull_dispatch2<32>(m, g, a, binary, mode, sm_count, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sfm
