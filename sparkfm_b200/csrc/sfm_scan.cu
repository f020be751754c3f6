// sfm_scan.cu -- own ordered stream compaction (the mini-batch samplers) and exclusive prefix sums.
//
// Round 1 used cub::DeviceSelect::If / cub::DeviceScan here; these kernels replace the library
// calls on every training path.  Everything is a fixed sequence of plain launches
// without spin-waits, tickets or look-back: per-tile partial results, ONE CTA that scans the
// tiles' totals, and a write-out pass -- integer arithmetic only, so the result does not depend on
// scheduling.
//
//   select (Bernoulli sampler of DESIGN.md 2.5 / PARTITION sampler): rows keep ASCENDING order.
//     select_mask_kernel   thread w decides the 64 consecutive GLOBAL rows of block
//                          (off >> 6) + w and keeps the hits as one 64-bit mask word (nothing is
//                          read: the Bernoulli sampler is bit-sliced over the block, the PARTITION
//                          sampler hashes each row number) -> masks[w]; hits of the tile -> counts[c]
//     scan_tiles_kernel    one CTA: counts[] <- exclusive prefix, grand total -> *d_count
//     select_write_kernel  re-reads the mask words, ranks the set bits inside the tile and
//                          writes the local row numbers to out[base(c) + rank]
//   exclusive scan (row lengths -> output offsets of a ragged batch; the radix sort's counts):
//     scan_sum_kernel      sum of each tile of 2048 items -> sums[tile]
//     scan_tiles_kernel    one CTA: exclusive prefix of the sums
//     scan_write_kernel    exclusive prefix inside the tile + the tile's base (in place is fine)
#include <cuda_runtime.h>
#include <stdint.h>

#include "sfm_common.h"

namespace sfm {

#define FULL 0xffffffffu

constexpr int SEL_THREADS = 256;
constexpr int SEL_ROWS = SEL_THREADS * 64;       // rows per tile (16384): one 64-row block per thread
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_IPT = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_IPT;   // items per tile (2048)
constexpr int TILES_THREADS = 1024;

__host__ __device__ __forceinline__ uint64_t mix64_s(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// Bernoulli sampler of DESIGN.md 2.5, bit-sliced: the 64 rows of global block q are decided
// together.  Word i (i = 1, 2, ...) = output number 64 q + i - 1 of SplitMix64 seeded with the
// iteration key holds, for each of the 64 rows, binary digit i of that row's uniform variate u;
// u < p = thr / 2^53 is settled at the first digit where u and p differ, so the digits of p are
// walked MSB first while rows are still undecided (about 8 words per block, not 128 hashes).
struct BernoulliBlock {
    uint64_t key, thr;
    int last;   // 1-based position of the lowest set digit of thr (no row can be selected after it)
    __device__ __forceinline__ uint64_t operator()(uint64_t q) const {
        uint64_t und = ~0ULL, hit = 0ULL;
        uint64_t ctr = key + (q << 6) * 0x9E3779B97F4A7C15ULL;
        for (int i = 1; i <= last && und; ++i) {
            const uint64_t w = mix64_s(ctr);
            ctr += 0x9E3779B97F4A7C15ULL;
            if ((thr >> (53 - i)) & 1ULL) {
                hit |= und & ~w;
                und &= w;
            } else {
                und &= ~w;
            }
        }
        return hit;
    }
};

// PARTITION sampler: row r (global) belongs to part (mix64(key ^ mix64(r)) >> 11) % n_parts.
struct PartBlock {
    uint64_t key, n_parts, part;
    __device__ __forceinline__ uint64_t operator()(uint64_t q) const {
        uint64_t m = 0ULL;
#pragma unroll 4
        for (int j = 0; j < 64; ++j)
            m |= (uint64_t)(((mix64_s(key ^ mix64_s((q << 6) + (uint64_t)j)) >> 11) % n_parts) == part) << j;
        return m;
    }
};

// block-wide exclusive scan of one value per thread; *total = sum over the block (every thread)
template <int THREADS, typename T>
__device__ __forceinline__ T block_excl_scan_t(T v, T* wsum /* [THREADS / 32] */, T* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    T excl = incl - v, tot = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) {
        const T s = wsum[w];
        if (w < warp) excl += s;
        tot += s;
    }
    if (total) *total = tot;
    __syncthreads();   // wsum may be reused by the caller's next scan
    return excl;
}

// ------------------------------------------------------------------------------------------
// Thread w of the grid decides global block q = (off >> 6) + w: bit j of its mask <-> global row
// 64 q + j <-> local row 64 q + j - off; bits outside the local range [0, n) are cleared.
template <typename BlockFn>
__global__ void __launch_bounds__(SEL_THREADS)
select_mask_kernel(int64_t n, int64_t off, BlockFn fn, uint64_t* __restrict__ masks,
                   uint32_t* __restrict__ counts) {
    __shared__ uint32_t wsum[SEL_THREADS / 32];
    const int64_t word = (int64_t)blockIdx.x * SEL_THREADS + threadIdx.x;
    const int64_t q = (off >> 6) + word;
    const int64_t g0 = q << 6;           // global row of bit 0
    const int64_t avail = off + n - g0;  // rows of the shard from g0 on
    uint64_t mask = 0ULL;
    if (avail > 0) {
        mask = fn((uint64_t)q);
        if (g0 < off) mask &= ~0ULL << (off - g0);         // 1 .. 63 leading rows belong to the previous shard
        if (avail < 64) mask &= ~0ULL >> (64 - avail);
    }
    masks[word] = mask;   // the scratch holds whole tiles: no bound check
    uint32_t c = (uint32_t)__popcll(mask);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < SEL_THREADS / 32; ++w) t += wsum[w];
        counts[blockIdx.x] = t;
    }
}

// One CTA: a[0 .. n) <- exclusive prefix (in place); grand total -> *total_a / *total_b (either may
// be null).  Thread t owns the contiguous piece [t * per, (t + 1) * per).
template <typename T, typename TotA>
__global__ void __launch_bounds__(TILES_THREADS)
scan_tiles_kernel(T* __restrict__ a, int64_t n, TotA* __restrict__ total_a) {
    __shared__ T wsum[TILES_THREADS / 32];
    const int64_t per = (n + TILES_THREADS - 1) / TILES_THREADS;
    const int64_t lo = min(n, (int64_t)threadIdx.x * per), hi = min(n, lo + per);
    T sum = 0;
    for (int64_t i = lo; i < hi; ++i) sum += a[i];
    T tot = 0;
    T run = block_excl_scan_t<TILES_THREADS, T>(sum, wsum, &tot);
    for (int64_t i = lo; i < hi; ++i) {
        const T t = a[i];
        a[i] = run;
        run += t;
    }
    if (threadIdx.x == 0 && total_a) *total_a = (TotA)tot;
}

__global__ void __launch_bounds__(SEL_THREADS)
select_write_kernel(int64_t off, const uint64_t* __restrict__ masks,
                    const uint32_t* __restrict__ bases, int32_t* __restrict__ out) {
    __shared__ uint32_t wsum[SEL_THREADS / 32];
    const int64_t word = (int64_t)blockIdx.x * SEL_THREADS + threadIdx.x;
    uint64_t mask = masks[word];
    const uint32_t c = (uint32_t)__popcll(mask);
    uint32_t pos = bases[blockIdx.x] + block_excl_scan_t<SEL_THREADS, uint32_t>(c, wsum, nullptr);
    const int64_t r0 = (((off >> 6) + word) << 6) - off;   // local row of bit 0 (set bits are >= 0)
    while (mask) {
        const int j = __ffsll((long long)mask) - 1;
        mask &= mask - 1ULL;
        out[pos++] = (int32_t)(r0 + j);
    }
}

static inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }

// 64-row blocks a shard of n rows can touch at any global offset, in whole tiles
static inline int64_t select_tiles(int64_t n) {
    const int64_t words = (n > 0 ? (n + 63) / 64 : 0) + 1;
    return (words + SEL_THREADS - 1) / SEL_THREADS;
}

size_t select_temp_bytes(int64_t n) {
    const int64_t tiles = select_tiles(n);
    return al256(sizeof(uint64_t) * (size_t)tiles * SEL_THREADS) + al256(sizeof(uint32_t) * (size_t)tiles);
}

template <typename BlockFn>
static cudaError_t select_rows(void* tmp, size_t tmp_bytes, int64_t n, int64_t global_off, BlockFn fn,
                               int32_t* out_rows, int32_t* d_count, cudaStream_t st,
                               int64_t* launches) {
    if (n < 0 || n >= 2147483647LL - SEL_ROWS || global_off < 0) return cudaErrorInvalidValue;
    if (n == 0) return cudaMemsetAsync(d_count, 0, sizeof(int32_t), st);
    if (!tmp || tmp_bytes < select_temp_bytes(n)) return cudaErrorInvalidValue;
    const int64_t words = ((global_off + n + 63) >> 6) - (global_off >> 6);
    const int64_t tiles = (words + SEL_THREADS - 1) / SEL_THREADS;   // <= select_tiles(n)
    uint64_t* masks = static_cast<uint64_t*>(tmp);
    uint32_t* counts = reinterpret_cast<uint32_t*>(
        static_cast<unsigned char*>(tmp) + al256(sizeof(uint64_t) * (size_t)select_tiles(n) * SEL_THREADS));
    *launches += 3;
    select_mask_kernel<BlockFn><<<(unsigned)tiles, SEL_THREADS, 0, st>>>(n, global_off, fn, masks, counts);
    scan_tiles_kernel<uint32_t, int32_t><<<1, TILES_THREADS, 0, st>>>(counts, tiles, d_count);
    select_write_kernel<<<(unsigned)tiles, SEL_THREADS, 0, st>>>(global_off, masks, counts, out_rows);
    return cudaGetLastError();
}

cudaError_t sample_rows_device(void* tmp, size_t tmp_bytes, int64_t n, int64_t global_off,
                               uint64_t key, uint64_t thr, int32_t* out_rows, int32_t* d_count,
                               cudaStream_t st, int64_t* launches) {
    if (thr == 0 || n == 0) return cudaMemsetAsync(d_count, 0, sizeof(int32_t), st);
    if (thr >= (1ULL << 53)) return cudaErrorInvalidValue;   // fraction >= 1 never comes here
    int tz = 0;
    while (!((thr >> tz) & 1ULL)) ++tz;
    return select_rows(tmp, tmp_bytes, n, global_off, BernoulliBlock{key, thr, 53 - tz}, out_rows, d_count,
                       st, launches);
}

cudaError_t partition_rows_device(void* tmp, size_t tmp_bytes, int64_t n, int64_t global_off,
                                  uint64_t key, int64_t n_parts, int64_t part, int32_t* out_rows,
                                  int32_t* d_count, cudaStream_t st, int64_t* launches) {
    if (n_parts < 1 || part < 0 || part >= n_parts) return cudaErrorInvalidValue;
    return select_rows(tmp, tmp_bytes, n, global_off, PartBlock{key, (uint64_t)n_parts, (uint64_t)part},
                       out_rows, d_count, st, launches);
}

// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_sum_kernel(const T* __restrict__ in, int64_t n, T* __restrict__ sums) {
    __shared__ T wsum[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    T s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_IPT; ++j) {
        const int64_t i = base + j * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        T t = 0;
#pragma unroll
        for (int w = 0; w < SCAN_THREADS / 32; ++w) t += wsum[w];
        sums[blockIdx.x] = t;
    }
}

// thread t owns the SCAN_IPT consecutive items [tile * SCAN_TILE + t * SCAN_IPT, +SCAN_IPT)
template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_write_kernel(const T* in, int64_t n, const T* __restrict__ bases, T* out) {
    __shared__ T wsum[SCAN_THREADS / 32];
    const int64_t i0 = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_IPT;
    T v[SCAN_IPT];
    T s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_IPT; ++j) {
        v[j] = i0 + j < n ? in[i0 + j] : (T)0;
        s += v[j];
    }
    // (in == out is allowed: every item of the tile is in registers before the barrier inside the
    // block scan, and tiles are disjoint)
    T run = bases[blockIdx.x] + block_excl_scan_t<SCAN_THREADS, T>(s, wsum, nullptr);
#pragma unroll
    for (int j = 0; j < SCAN_IPT; ++j) {
        if (i0 + j < n) out[i0 + j] = run;
        run += v[j];
    }
}

template <typename T>
static size_t scan_bytes_t(int64_t n) {
    const int64_t tiles = n > 0 ? (n + SCAN_TILE - 1) / SCAN_TILE : 1;
    return al256(sizeof(T) * (size_t)tiles);
}

template <typename T>
static cudaError_t exclusive_scan_t(void* tmp, size_t tmp_bytes, const T* in, T* out, int64_t n,
                                    cudaStream_t st, int64_t* launches) {
    if (n <= 0) return cudaSuccess;
    if (!tmp || tmp_bytes < scan_bytes_t<T>(n)) return cudaErrorInvalidValue;
    const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles > 2147483647LL) return cudaErrorInvalidValue;
    T* sums = static_cast<T*>(tmp);
    *launches += 3;
    scan_sum_kernel<T><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, sums);
    scan_tiles_kernel<T, T><<<1, TILES_THREADS, 0, st>>>(sums, tiles, (T*)nullptr);
    scan_write_kernel<T><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, sums, out);
    return cudaGetLastError();
}

size_t scan_temp_bytes(int64_t n) { return scan_bytes_t<int64_t>(n); }

cudaError_t exclusive_scan_i64(void* tmp, size_t tmp_bytes, const int64_t* in, int64_t* out,
                               int64_t n, cudaStream_t st, int64_t* launches) {
    return exclusive_scan_t<int64_t>(tmp, tmp_bytes, in, out, n, st, launches);
}

size_t scan_u32_temp_bytes(int64_t n) { return scan_bytes_t<uint32_t>(n); }

cudaError_t exclusive_scan_u32(void* tmp, size_t tmp_bytes, const uint32_t* in, uint32_t* out,
                               int64_t n, cudaStream_t st, int64_t* launches) {
    return exclusive_scan_t<uint32_t>(tmp, tmp_bytes, in, out, n, st, launches);
}

}  // namespace sfm
