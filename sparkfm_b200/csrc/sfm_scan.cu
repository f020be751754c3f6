// sfm_scan.cu -- own ordered stream compaction (the mini-batch samplers) and exclusive prefix sums.
//
// Round 1 used cub::DeviceSelect::If / cub::DeviceScan here; these kernels replace the library
// calls on every training path.  Everything is a fixed sequence of plain launches
// without spin-waits, tickets or look-back: per-tile partial results, ONE CTA that scans the
// tiles' totals, and a write-out pass -- integer arithmetic only, so the result does not depend on
// scheduling.
//
//   select (Bernoulli sampler of DESIGN.md 2.5 / PARTITION sampler): rows keep ASCENDING order.
//     select_mask_kernel   thread t of tile c tests the 32 consecutive rows
//                          [(c * 256 + t) * 32, +32) (the predicate is a hash of the row number:
//                          nothing is read), keeps the hits as one 32-bit mask word
//                          -> masks[c * 256 + t]; hits of the tile -> counts[c]
//     scan_tiles_kernel    one CTA: counts[] <- exclusive prefix, grand total -> *d_count
//     select_write_kernel  re-reads the mask words (the hash is NOT evaluated again), ranks the
//                          set bits inside the tile and writes the row numbers to
//                          out[base(c) + rank]
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
constexpr int SEL_ROWS = SEL_THREADS * 32;       // rows per tile (8192)
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

// row r (global) is in the batch iff (mix64(key ^ mix64(r)) >> 11) < thr.
struct InBatch {
    uint64_t key, thr;
    __device__ __forceinline__ bool operator()(uint64_t global_row) const {
        return (mix64_s(key ^ mix64_s(global_row)) >> 11) < thr;
    }
};

// PARTITION sampler: row r (global) belongs to part (mix64(key ^ mix64(r)) >> 11) % n_parts.
struct InPart {
    uint64_t key, n_parts, part;
    __device__ __forceinline__ bool operator()(uint64_t global_row) const {
        return ((mix64_s(key ^ mix64_s(global_row)) >> 11) % n_parts) == part;
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
template <typename Pred>
__global__ void __launch_bounds__(SEL_THREADS)
select_mask_kernel(int64_t n, int64_t global_off, Pred pred, uint32_t* __restrict__ masks,
                   uint32_t* __restrict__ counts) {
    __shared__ uint32_t wsum[SEL_THREADS / 32];
    const int64_t word = (int64_t)blockIdx.x * SEL_THREADS + threadIdx.x;
    const int64_t r0 = word * 32;
    uint32_t mask = 0;
    if (r0 + 32 <= n) {
#pragma unroll 8
        for (int j = 0; j < 32; ++j)
            mask |= (pred((uint64_t)(global_off + r0 + j)) ? 1u : 0u) << j;
    } else {
        for (int j = 0; j < 32 && r0 + j < n; ++j)
            mask |= (pred((uint64_t)(global_off + r0 + j)) ? 1u : 0u) << j;
    }
    masks[word] = mask;   // the scratch holds whole tiles: no bound check
    uint32_t c = (uint32_t)__popc(mask);
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
select_write_kernel(int64_t n, const uint32_t* __restrict__ masks,
                    const uint32_t* __restrict__ bases, int32_t* __restrict__ out) {
    __shared__ uint32_t wsum[SEL_THREADS / 32];
    const int64_t word = (int64_t)blockIdx.x * SEL_THREADS + threadIdx.x;
    uint32_t mask = masks[word];
    const uint32_t c = (uint32_t)__popc(mask);
    uint32_t pos = bases[blockIdx.x] + block_excl_scan_t<SEL_THREADS, uint32_t>(c, wsum, nullptr);
    const int32_t r0 = (int32_t)(word * 32);   // n < 2^31 (checked by the launcher)
    while (mask) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1u;
        out[pos++] = r0 + j;
    }
}

static inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }

size_t select_temp_bytes(int64_t n) {
    const int64_t tiles = n > 0 ? (n + SEL_ROWS - 1) / SEL_ROWS : 1;
    return al256(sizeof(uint32_t) * (size_t)tiles * SEL_THREADS) + al256(sizeof(uint32_t) * (size_t)tiles);
}

template <typename Pred>
static cudaError_t select_rows(void* tmp, size_t tmp_bytes, int64_t n, int64_t global_off, Pred pred,
                               int32_t* out_rows, int32_t* d_count, cudaStream_t st,
                               int64_t* launches) {
    if (n < 0 || n >= 2147483647LL - SEL_ROWS) return cudaErrorInvalidValue;
    if (n == 0) return cudaMemsetAsync(d_count, 0, sizeof(int32_t), st);
    if (!tmp || tmp_bytes < select_temp_bytes(n)) return cudaErrorInvalidValue;
    const int64_t tiles = (n + SEL_ROWS - 1) / SEL_ROWS;
    uint32_t* masks = static_cast<uint32_t*>(tmp);
    uint32_t* counts = reinterpret_cast<uint32_t*>(static_cast<unsigned char*>(tmp) +
                                                   al256(sizeof(uint32_t) * (size_t)tiles * SEL_THREADS));
    *launches += 3;
    select_mask_kernel<Pred><<<(unsigned)tiles, SEL_THREADS, 0, st>>>(n, global_off, pred, masks, counts);
    scan_tiles_kernel<uint32_t, int32_t><<<1, TILES_THREADS, 0, st>>>(counts, tiles, d_count);
    select_write_kernel<<<(unsigned)tiles, SEL_THREADS, 0, st>>>(n, masks, counts, out_rows);
    return cudaGetLastError();
}

cudaError_t sample_rows_device(void* tmp, size_t tmp_bytes, int64_t n, int64_t global_off,
                               uint64_t key, uint64_t thr, int32_t* out_rows, int32_t* d_count,
                               cudaStream_t st, int64_t* launches) {
    return select_rows(tmp, tmp_bytes, n, global_off, InBatch{key, thr}, out_rows, d_count, st, launches);
}

cudaError_t partition_rows_device(void* tmp, size_t tmp_bytes, int64_t n, int64_t global_off,
                                  uint64_t key, int64_t n_parts, int64_t part, int32_t* out_rows,
                                  int32_t* d_count, cudaStream_t st, int64_t* launches) {
    if (n_parts < 1 || part < 0 || part >= n_parts) return cudaErrorInvalidValue;
    return select_rows(tmp, tmp_bytes, n, global_off, InPart{key, (uint64_t)n_parts, (uint64_t)part},
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
