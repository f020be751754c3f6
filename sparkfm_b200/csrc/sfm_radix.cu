// sfm_radix.cu -- stable LSD radix sort of the batch's (feature id, payload) entries in wide
// digits: 20 key bits are two passes of 10 bits instead of the library's three passes of 8.
//
// The transposition (entries in row order -> entries grouped by feature, rows ascending inside a
// feature) is what the deterministic reduce-by-feature consumes.  It is pure data
// movement, HBM-bound: every pass reads and writes n * (4 + sizeof payload) bytes, so the pass
// count is the cost.  One pass =
//   radix_count_kernel    per tile: digit histogram in shared memory -> counts[digit][tile]
//   exclusive scan        over the digit-major counts = global start of (digit, tile)
//   radix_scatter_kernel  per tile: stable rank of every entry (warp-private histograms filled
//                         in entry order with match.any, prefix over the warps), entries
//                         re-ordered in shared memory so that the global stores of one digit are
//                         contiguous, then written to start(digit, tile) + rank.
// Tiles are contiguous and ranked in order, so the sort is stable and its output is bit-identical
// to the library sort it replaces (SFM_SORT=cub selects the library; tests compare the two).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "sfm_common.h"

namespace sfm {

#ifndef SFM_RX_THREADS
#define SFM_RX_THREADS 512
#endif
constexpr int RX_THREADS = SFM_RX_THREADS;   // 512: two CTAs per SM at 10-bit digits
constexpr int RX_WARPS = RX_THREADS / 32;
constexpr int RX_MAX_BITS = 10;

// bytes of shared-memory region A: max(re-order buffers, per-warp peer masks at the widest digit)
__host__ __device__ constexpr size_t rx_region_a(size_t pay_bytes) {
    return (size_t)RX_THREADS * (pay_bytes == 4 ? 16 : 8) * (4 + pay_bytes) >
                   (size_t)4 * RX_WARPS * (1 << RX_MAX_BITS)
               ? (size_t)RX_THREADS * (pay_bytes == 4 ? 16 : 8) * (4 + pay_bytes)
               : (size_t)4 * RX_WARPS * (1 << RX_MAX_BITS);
}

static_assert(RX_THREADS % 32 == 0 && (1 << RX_MAX_BITS) / 2 <= 2 * RX_THREADS, "prefix phase: <= 2 words per thread");

template <typename PayT>
struct RxCfg {
    static constexpr int IPT = sizeof(PayT) == 4 ? 16 : 8;   // entries per thread
    static constexpr int TILE = RX_THREADS * IPT;
};

__global__ void __launch_bounds__(RX_THREADS)
radix_count_kernel(const uint32_t* __restrict__ keys, int n, int shift, int bits, int tile_items,
                   uint32_t* __restrict__ counts, int n_tiles) {
    extern __shared__ uint32_t rx_hist[];
    const int nb = 1 << bits;
    const uint32_t mask = (uint32_t)nb - 1u;
    for (int b = threadIdx.x; b < nb; b += RX_THREADS) rx_hist[b] = 0;
    __syncthreads();
    const int base = blockIdx.x * tile_items;   // tile_items is a multiple of 4 * RX_THREADS
    const int end = min(base + tile_items, n);
    if (end - base == tile_items) {             // full tile: 16-byte loads, all issued up front
        const uint4* k4 = reinterpret_cast<const uint4*>(keys + base);
        const int n4 = tile_items / 4;
        for (int i0 = 0; i0 < n4; i0 += 4 * RX_THREADS) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * RX_THREADS + threadIdx.x;
                v[u] = i < n4 ? __ldg(k4 + i) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i0 + u * RX_THREADS + threadIdx.x < n4) {
                    atomicAdd(&rx_hist[(v[u].x >> shift) & mask], 1u);
                    atomicAdd(&rx_hist[(v[u].y >> shift) & mask], 1u);
                    atomicAdd(&rx_hist[(v[u].z >> shift) & mask], 1u);
                    atomicAdd(&rx_hist[(v[u].w >> shift) & mask], 1u);
                }
            }
        }
    } else {
        for (int i = base + threadIdx.x; i < end; i += RX_THREADS)
            atomicAdd(&rx_hist[(__ldg(keys + i) >> shift) & mask], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nb; b += RX_THREADS)
        counts[(size_t)b * n_tiles + blockIdx.x] = rx_hist[b];
}

// IMPLICIT: the payload of input position i is i / m (all-ones rows of m entries each, in row
// order: the forward kernel then writes no payload at all); magic = ceil(2^40 / m), exact for
// i < 2^31 and m < 512.
template <typename PayT, bool IMPLICIT>
__global__ void __launch_bounds__(RX_THREADS, 1024 / RX_THREADS)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const PayT* __restrict__ pay_in,
                     uint32_t* __restrict__ keys_out, PayT* __restrict__ pay_out, int n, int shift,
                     int bits, const uint32_t* __restrict__ offsets, int n_tiles,
                     unsigned long long magic) {
    constexpr int IPT = RxCfg<PayT>::IPT;
    constexpr int TILE = RxCfg<PayT>::TILE;
    extern __shared__ __align__(16) unsigned char rx_smem[];
    const int nb = 1 << bits;
    const uint32_t mask = (uint32_t)nb - 1u;
    // region A: the re-order buffers, and before them (ranking phase) the per-warp peer masks
    constexpr size_t REGION_A = rx_region_a(sizeof(PayT));
    PayT* pay_s = reinterpret_cast<PayT*>(rx_smem);                       // [TILE]
    uint32_t* keys_s = reinterpret_cast<uint32_t*>(pay_s + TILE);          // [TILE]
    uint32_t* wmask = reinterpret_cast<uint32_t*>(rx_smem);                // [RX_WARPS][nb], aliases A
    uint32_t* binoff = reinterpret_cast<uint32_t*>(rx_smem + REGION_A);    // [nb] global start - local start
    uint16_t* lstart = reinterpret_cast<uint16_t*>(binoff + nb);           // [nb] local start of the digit
    uint16_t* whist = lstart + nb;                                         // [RX_WARPS][nb]
    __shared__ uint32_t wsum[RX_WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int tile_base = tile * TILE;
    const int n_valid = min(TILE, n - tile_base);
    const int row_words = nb / 2;   // one warp histogram = row_words 32-bit words (nb >= 2)

    {   // zero the warp histograms (u16 pairs as words)
        uint32_t* z = reinterpret_cast<uint32_t*>(whist);
        const int words = RX_WARPS * row_words;
        for (int i = tid; i < words; i += RX_THREADS) z[i] = 0;
        for (int i = tid; i < RX_WARPS * nb; i += RX_THREADS) wmask[i] = 0;
    }

    // ---- stable rank inside the warp's strip, in entry order.  The lanes holding the same digit
    // find each other through a per-warp mask word per digit (OR in the own lane bit, read the
    // word back): shared-memory traffic only.  match.any would do the same in one instruction but
    // costs ~45 cycles of the SM's single ADU pipe per warp on sm_100 and bounds the kernel.
    uint32_t key[IPT];
    uint32_t rk2[IPT / 2];   // two 16-bit ranks per register
    const int strip = warp * (32 * IPT);
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
        const int p = strip + r * 32 + lane;
        key[r] = p < n_valid ? __ldg(keys_in + tile_base + p) : 0u;
    }
    __syncthreads();   // histograms and masks are zero
    uint16_t* wh = whist + warp * nb;
    uint32_t* wm = wmask + warp * nb;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned lbit = 1u << lane;
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
        const bool valid = strip + r * 32 + lane < n_valid;
        const uint32_t d = (key[r] >> shift) & mask;
        if (valid) atomicOr(&wm[d], lbit);
        __syncwarp();
        const unsigned m = valid ? wm[d] : lbit;
        const uint32_t old = valid ? (uint32_t)wh[d] : 0u;
        __syncwarp();
        if (valid && (m & lt) == 0u) {   // first lane of the group: reset the mask, bump the count
            wm[d] = 0u;
            wh[d] = (uint16_t)(old + __popc(m));
        }
        __syncwarp();
        const uint32_t rank = old + __popc(m & lt);
        if (r & 1) rk2[r / 2] |= rank << 16; else rk2[r / 2] = rank;
    }
    __syncthreads();

    // ---- per digit: exclusive prefix over the warps (two digits per 32-bit word: counts stay
    // below 2^16, so the halves never carry), then over the digits of the tile
    {
        const int cw = row_words >= RX_THREADS ? row_words / RX_THREADS : 1;   // words per thread (<= 2)
        const int w0 = tid * cw;
        const bool has = w0 < row_words;
        uint32_t* wrows = reinterpret_cast<uint32_t*>(whist);
        uint32_t cnt2[2] = {0, 0};
        uint32_t local = 0;
        if (has) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (j < cw) {
                    uint32_t run = 0;
#pragma unroll
                    for (int w = 0; w < RX_WARPS; ++w) {
                        const uint32_t t = wrows[w * row_words + w0 + j];
                        wrows[w * row_words + w0 + j] = run;
                        run += t;
                    }
                    cnt2[j] = run;
                    local += (run & 0xffffu) + (run >> 16);
                }
            }
        }
        uint32_t incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint32_t excl = incl - local;
        for (int w = 0; w < warp; ++w) excl += wsum[w];
        if (has) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (j < cw) {
                    const int b = 2 * (w0 + j);
                    const uint32_t c_lo = cnt2[j] & 0xffffu, c_hi = cnt2[j] >> 16;
                    lstart[b] = (uint16_t)excl;
                    binoff[b] = __ldg(offsets + (size_t)b * n_tiles + tile) - excl;
                    if (b + 1 < nb) {
                        lstart[b + 1] = (uint16_t)(excl + c_lo);
                        binoff[b + 1] = __ldg(offsets + (size_t)(b + 1) * n_tiles + tile) - (excl + c_lo);
                    }
                    excl += c_lo + c_hi;
                }
            }
        }
    }
    __syncthreads();

    // ---- re-order inside the tile
#pragma unroll
    for (int r0 = 0; r0 < IPT; r0 += 8) {   // 8 payload loads in flight (register budget)
        PayT pv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int p = strip + (r0 + u) * 32 + lane;
            if (p < n_valid) {
                if (IMPLICIT)
                    pv[u] = (PayT)(((unsigned long long)(uint32_t)(tile_base + p) * magic) >> 40);
                else
                    pv[u] = pay_in[tile_base + p];
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = r0 + u;
            const int p = strip + r * 32 + lane;
            if (p < n_valid) {
                const uint32_t d = (key[r] >> shift) & mask;
                const uint32_t rank = (r & 1) ? (rk2[r / 2] >> 16) : (rk2[r / 2] & 0xffffu);
                const int q = (int)lstart[d] + (int)wh[d] + (int)rank;
                keys_s[q] = key[r];
                pay_s[q] = pv[u];
            }
        }
        asm volatile("" ::: "memory");
    }
    __syncthreads();

    // ---- contiguous runs per digit go out
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        const int q = j * RX_THREADS + tid;
        if (q < n_valid) {
            const uint32_t k = keys_s[q];
            const uint32_t g = binoff[(k >> shift) & mask] + (uint32_t)q;
            keys_out[g] = k;
            pay_out[g] = pay_s[q];
        }
    }
}

static inline size_t rx_align(size_t x) { return (x + 255) / 256 * 256; }

static void rx_plan(int end_bit, int* passes, int bits[4]) {
    int P = (end_bit + RX_MAX_BITS - 1) / RX_MAX_BITS;
    if (P < 1) P = 1;
    const int base = end_bit / P, extra = end_bit % P;
    for (int j = 0; j < P; ++j) bits[j] = base + (j < extra ? 1 : 0);
    if (bits[0] < 1) bits[0] = 1;
    *passes = P;
}

bool radix_usable(int64_t n, int end_bit) {
    if (knobs().sort_cub) return false;
    return n > 0 && n < 2000000000LL && end_bit >= 1 && end_bit <= 3 * RX_MAX_BITS;
}

template <typename PayT>
static size_t radix_temp_bytes_t(int64_t n, int end_bit) {
    int P, bits[4];
    rx_plan(end_bit, &P, bits);
    const int64_t tiles = (n + RxCfg<PayT>::TILE - 1) / RxCfg<PayT>::TILE;
    size_t b = rx_align(sizeof(uint32_t) * ((size_t)tiles << bits[0]));   // bits[0] is the widest
    b += rx_align(scan_u32_temp_bytes(tiles << bits[0]));
    if (P > 1) b += rx_align(sizeof(uint32_t) * (size_t)n) + rx_align(sizeof(PayT) * (size_t)n);
    return b;
}

size_t radix_temp_bytes(int64_t n, int end_bit, int pay_bytes) {
    return pay_bytes == 4 ? radix_temp_bytes_t<uint32_t>(n, end_bit)
                          : radix_temp_bytes_t<unsigned long long>(n, end_bit);
}

template <typename PayT>
static cudaError_t radix_sort_t(void* tmp, size_t tmp_bytes, const uint32_t* keys_in,
                                uint32_t* keys_out, const PayT* pay_in, PayT* pay_out, int64_t n64,
                                int end_bit, int implicit_div, cudaStream_t st, int64_t* launches) {
    constexpr int TILE = RxCfg<PayT>::TILE;
    int P, bits[4];
    rx_plan(end_bit, &P, bits);
    const int n = (int)n64;
    const int tiles = (n + TILE - 1) / TILE;
    if (radix_temp_bytes_t<PayT>(n64, end_bit) > tmp_bytes) return cudaErrorInvalidValue;
    unsigned char* p = static_cast<unsigned char*>(tmp);
    uint32_t* counts = reinterpret_cast<uint32_t*>(p);
    p += rx_align(sizeof(uint32_t) * ((size_t)tiles << bits[0]));
    void* scan_tmp = p;
    const size_t scan_bytes = scan_u32_temp_bytes((int64_t)tiles << bits[0]);
    p += rx_align(scan_bytes);
    uint32_t* keys_tmp = nullptr;
    PayT* pay_tmp = nullptr;
    if (P > 1) {
        keys_tmp = reinterpret_cast<uint32_t*>(p);
        p += rx_align(sizeof(uint32_t) * (size_t)n);
        pay_tmp = reinterpret_cast<PayT*>(p);
    }
    {   // per device and cheap: set on every call
        const size_t smem_max = rx_region_a(sizeof(PayT)) +
                                ((size_t)6 + 2 * RX_WARPS) * ((size_t)1 << RX_MAX_BITS);
        cudaError_t e = cudaFuncSetAttribute(radix_scatter_kernel<PayT, false>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem_max);
        if (e == cudaSuccess && implicit_div)
            e = cudaFuncSetAttribute(radix_scatter_kernel<PayT, true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        if (e != cudaSuccess) return e;
    }
    const uint32_t* src_k = keys_in;
    const PayT* src_p = pay_in;
    int shift = 0;
    for (int j = 0; j < P; ++j) {
        const bool to_out = ((P - 1 - j) & 1) == 0;
        uint32_t* dst_k = to_out ? keys_out : keys_tmp;
        PayT* dst_p = to_out ? pay_out : pay_tmp;
        const int nb = 1 << bits[j];
        radix_count_kernel<<<tiles, RX_THREADS, sizeof(uint32_t) * nb, st>>>(
            src_k, n, shift, bits[j], TILE, counts, tiles);
        cudaError_t e = exclusive_scan_u32(scan_tmp, scan_bytes, counts, counts,
                                           (int64_t)tiles * nb, st, launches);
        if (e != cudaSuccess) return e;
        const size_t smem = rx_region_a(sizeof(PayT)) + ((size_t)6 + 2 * RX_WARPS) * nb;
        if (j == 0 && implicit_div) {
            const unsigned long long magic =
                ((1ULL << 40) + (unsigned long long)implicit_div - 1) / (unsigned long long)implicit_div;
            radix_scatter_kernel<PayT, true><<<tiles, RX_THREADS, smem, st>>>(
                src_k, src_p, dst_k, dst_p, n, shift, bits[j], counts, tiles, magic);
        } else {
            radix_scatter_kernel<PayT, false><<<tiles, RX_THREADS, smem, st>>>(
                src_k, src_p, dst_k, dst_p, n, shift, bits[j], counts, tiles, 0ULL);
        }
        *launches += 2;
        src_k = dst_k;
        src_p = dst_p;
        shift += bits[j];
    }
    return cudaGetLastError();
}

cudaError_t radix_sort_pairs32(void* tmp, size_t tmp_bytes, const uint32_t* keys_in,
                               uint32_t* keys_out, const uint32_t* pay_in, uint32_t* pay_out,
                               int64_t n, int end_bit, int implicit_div, cudaStream_t st,
                               int64_t* launches) {
    if (implicit_div < 0 || implicit_div >= 512 || (!implicit_div && !pay_in))
        return cudaErrorInvalidValue;
    return radix_sort_t<uint32_t>(tmp, tmp_bytes, keys_in, keys_out, pay_in, pay_out, n, end_bit,
                                  implicit_div, st, launches);
}

cudaError_t radix_sort_pairs64(void* tmp, size_t tmp_bytes, const uint32_t* keys_in,
                               uint32_t* keys_out, const uint2* pay_in, uint2* pay_out, int64_t n,
                               int end_bit, cudaStream_t st, int64_t* launches) {
    return radix_sort_t<unsigned long long>(
        tmp, tmp_bytes, keys_in, keys_out, reinterpret_cast<const unsigned long long*>(pay_in),
        reinterpret_cast<unsigned long long*>(pay_out), n, end_bit, 0, st, launches);
}

}  // namespace sfm
