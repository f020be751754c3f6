// sfm_als.cu -- ALS.learn (fm/lib/ALS.scala:15-75) on the GPU: one call = one sweep of the
// closed-form coordinate updates over w0, every w_i and every v_if, for squared loss.
//
// The reference runs the sweep serially on the Spark driver over a hash map of the transposed
// data set; the coordinate order (ids ascending) matters because every accepted update changes
// the cached residuals e_r = yhat_r - y_r that the next coordinate reads.  Two coordinates
// interact only through rows they share, so the sweep is EXACTLY the sequential one if features
// are processed level by level, where level(i) = 1 + max level of the smaller ids that share a
// row with i: inside a level all columns are row-disjoint and run in parallel, one CTA per
// column.  Layout: the transposed input (DataSet.scala:31-38) is the batch's entry list sorted
// by (feature, row) -- the same emit + radix sort the SGD path uses -- kept resident together
// with the level schedule (the analogue of the reference's cached `transposeInput`, :34), the
// residuals e[n_rows] and the per-factor cache q[n_rows] (fp64, like the reference).
//   e = predict - y            als_residual_kernel       (:17, :142-144), fp64 accumulation
//   w0                         als_sum / als_w0_kernel   (:19-28, :152-154)
//   w_i, v_if per level        als_level_kernel<IS_V>    (:36-43, :45-70, :156-198)
//   q_r = sum_i v_if x_ri      als_q_kernel              (:146-150)
// The model is the handle's fp32 model: an accepted theta* is computed in fp64, rounded to fp32,
// and the residual update uses the rounded value (oracle flag FMO_ALS_STORE_F32).  Column sums
// are fixed-shape trees (deterministic); they differ from the reference's left folds by fp64
// rounding only.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "sfm_common.h"

namespace sfm {

#define CU(call)                                                                        \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess)                                                          \
            return set_err(h, e_ == cudaErrorMemoryAllocation ? SFM_ERR_OOM : SFM_ERR_CUDA, \
                           std::string(#call) + ": " + cudaGetErrorString(e_));        \
    } while (0)
#define RC(call)                       \
    do {                               \
        int rc_ = (call);              \
        if (rc_ != SFM_OK) return rc_; \
    } while (0)

struct AlsState {
    bool built = false;
    bool binary = false;
    int64_t n_rows = 0, nnz = 0;
    Buf pay;      // sorted entries: uint32 row (all-ones data) or uint2 {row, x bits}
    Buf colptr;   // int32 [n_slots + 1]
    Buf order;    // uint32 [n_slots]: feature ids by (level, id); level 0 = empty columns
    Buf lvl, lvl_sorted, rowlvl;
    Buf e, q;     // double [n_rows]
    Buf part;     // double [ALS_PARTS + 8]: reduction partials, then scalars
    std::vector<int32_t> lvl_off;   // host: start of level l in `order`, size max_level + 2
    cudaGraphExec_t exec = nullptr; // the sweep's launches, captured once
    float graph_key[4] = {-1.f, 0.f, 0.f, 0.f};   // flags, reg0, regw, regv the graph was built for
    int64_t graph_launches = 0;
    bool graph_failed = false;
};

constexpr int ALS_PARTS = 512;
constexpr int ALS_CTA = 128;

// ---------------------------------------------------------------------------------------------
__global__ void als_colptr_kernel(const uint32_t* __restrict__ keys, int nnz, int n_slots,
                                  int32_t* __restrict__ colptr) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p > nnz) return;
    const int64_t prev = p == 0 ? -1 : (int64_t)keys[p - 1];
    const int64_t cur = p == nnz ? (int64_t)n_slots : (int64_t)keys[p];
    for (int64_t f = prev + 1; f <= cur; ++f) colptr[f] = (int32_t)p;   // [prev+1, cur] start here
}

template <bool BINARY>
__device__ __forceinline__ void als_entry(const void* pay, int p, int& row, double& x) {
    if (BINARY) {
        row = (int)reinterpret_cast<const uint32_t*>(pay)[p];
        x = 1.0;
    } else {
        const uint2 t = reinterpret_cast<const uint2*>(pay)[p];
        row = (int)t.x;
        x = (double)__uint_as_float(t.y);
    }
}

template <bool BINARY>
__global__ void als_dup_kernel(const uint32_t* __restrict__ keys, const void* __restrict__ pay,
                               int nnz, int32_t* __restrict__ flag) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < 1 || p >= nnz) return;
    int r0, r1;
    double x;
    als_entry<BINARY>(pay, p - 1, r0, x);
    als_entry<BINARY>(pay, p, r1, x);
    if (keys[p] == keys[p - 1] && r0 == r1) atomicExch(flag, 1);
}

// level(i) = 1 + max over the column's rows of the level of the previous (smaller) id in that row;
// one CTA walks the features in id order (a one-off per data set, like the transposition).
template <bool BINARY>
__global__ void __launch_bounds__(1024)
als_levels_kernel(const int32_t* __restrict__ colptr, const void* __restrict__ pay, int n_slots,
                  int32_t* __restrict__ rowlvl, uint32_t* __restrict__ lvl,
                  int32_t* __restrict__ max_out) {
    __shared__ int wmax[32];
    __shared__ int cur;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int maxlvl = 0;
    for (int i = 0; i < n_slots; ++i) {
        const int a = colptr[i], b = colptr[i + 1];
        if (a == b) {   // uniform: empty column
            if (tid == 0) lvl[i] = 0;
            continue;
        }
        int m = 0;
        for (int p = a + tid; p < b; p += 1024) {
            int row;
            double x;
            als_entry<BINARY>(pay, p, row, x);
            m = max(m, rowlvl[row]);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (b - a > 32) {
            if (lane == 0) wmax[warp] = m;
            __syncthreads();
            if (tid == 0) {
                int t = 0;
                for (int w = 0; w < 32; ++w) t = max(t, wmax[w]);
                cur = t + 1;
            }
            __syncthreads();
            m = cur;
        } else {
            m = __shfl_sync(0xffffffffu, m, 0) + 1;   // only warp 0 saw entries
            if (tid == 0) cur = m;
            __syncthreads();
            m = cur;
        }
        for (int p = a + tid; p < b; p += 1024) {
            int row;
            double x;
            als_entry<BINARY>(pay, p, row, x);
            rowlvl[row] = m;
        }
        if (tid == 0) lvl[i] = (uint32_t)m;
        maxlvl = max(maxlvl, m);
        __syncthreads();   // rowlvl writes before the next column reads them; wmax / cur reuse
    }
    if (tid == 0) *max_out = maxlvl;
}

// e_r = predict(x_r) - y_r with fp64 accumulation (FMModel.scala:34-63 from fp32 parameters).
__global__ void __launch_bounds__(256)
als_residual_kernel(const float* __restrict__ V, const float* __restrict__ W,
                    const float* __restrict__ W0, int k, int kp, int k0, int k1,
                    const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ idx,
                    const float* __restrict__ val, const float* __restrict__ label, int uniform_m,
                    int n_rows, double* __restrict__ e) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = warp0; r < n_rows; r += nwarps) {
        const int64_t beg = uniform_m >= 0 ? r * uniform_m : row_ptr[r];
        const int64_t end = uniform_m >= 0 ? beg + uniform_m : row_ptr[r + 1];
        double s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0}, lin = 0.0;
        for (int64_t j = beg; j < end; ++j) {
            const int i = __ldg(idx + j);
            const double x = val ? (double)__ldg(val + j) : 1.0;
            if (lane == 0 && k1) lin += (double)__ldg(W + i) * x;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int f = lane + 32 * c;
                if (f < k) {
                    const double t = (double)__ldg(V + (int64_t)i * kp + f) * x;
                    s[c] += t;
                    ss[c] += t * t;
                }
            }
        }
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) acc += 0.5 * (s[c] * s[c] - ss[c]);
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) e[r] = (k0 ? (double)*W0 : 0.0) + lin + acc - (double)__ldg(label + r);
    }
}

// fixed-shape sum of e (SQUARE = false) or e^2: ALS_PARTS partials, folded by the consumer
template <bool SQUARE>
__global__ void __launch_bounds__(256)
als_sum_kernel(const double* __restrict__ e, int n, double* __restrict__ part) {
    __shared__ double sh[256];
    double a = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double t = e[i];
        a += SQUARE ? t * t : t;
    }
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}

__device__ __forceinline__ bool als_updatable(double nv, double v) {   // ALS.scala:178-180
    return !isnan(nv) && !isinf(nv) && nv != v;
}

// computeTheta (:167-176) + fp32 storage; returns the stored value, *delta = stored - theta
__device__ __forceinline__ float als_theta(float theta_f, double reg, double sum_e_h,
                                           double sum_h_sqr, double* delta) {
    const double theta = (double)theta_f;
    const double nv = -(sum_e_h - theta * sum_h_sqr) / (reg + sum_h_sqr);
    const double acc = als_updatable(nv, theta) ? nv : theta;
    const float st = (float)acc;
    *delta = als_updatable((double)st, theta) ? (double)st - theta : 0.0;
    return st;
}

// scal[0] = sum of the partials (sum e, or sum e^2 when w0 == nullptr); w0 step (:19-28)
__global__ void als_w0_kernel(const double* __restrict__ part, float* __restrict__ w0, double reg0,
                              int n_rows, int correct_e, double* __restrict__ scal) {
    if (threadIdx.x != 0) return;
    double s = 0.0;
    for (int i = 0; i < ALS_PARTS; ++i) s += part[i];
    scal[0] = s;
    if (w0) {
        double delta;
        *w0 = als_theta(*w0, reg0, s, (double)n_rows, &delta);
        // The reference adds (w0* - fm.w0) inside a LAZY RDD that is first materialised after
        // `fm.w0 = w0` (ALS.scala:17,24,27,31,142-144): the term is 0 there, but the same
        // re-evaluation runs fm.predict with the NEW w0, so its residuals do carry the shift.
        scal[1] = correct_e ? delta : 0.0;
    }
}

__global__ void als_shift_kernel(double* __restrict__ e, int n, const double* __restrict__ scal) {
    const double d = scal[1];
    if (d == 0.0) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        e[i] += d;
}

__global__ void __launch_bounds__(256)
als_q_kernel(const float* __restrict__ V, int kp, int f, const int64_t* __restrict__ row_ptr,
             const int32_t* __restrict__ idx, const float* __restrict__ val, int uniform_m,
             int n_rows, double* __restrict__ q) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
         r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t beg = uniform_m >= 0 ? r * uniform_m : row_ptr[r];
        const int64_t end = uniform_m >= 0 ? beg + uniform_m : row_ptr[r + 1];
        double s = 0.0;
        for (int64_t j = beg; j < end; ++j)
            s += (double)__ldg(V + (int64_t)__ldg(idx + j) * kp + f) * (val ? (double)__ldg(val + j) : 1.0);
        q[r] = s;
    }
}

// One level: CTA b owns column order[first + b]; columns of a level share no row.
template <bool IS_V, bool BINARY>
__global__ void __launch_bounds__(ALS_CTA)
als_level_kernel(const uint32_t* __restrict__ order, int first, const int32_t* __restrict__ colptr,
                 const void* __restrict__ pay, float* __restrict__ theta_base, int stride, int off,
                 double reg, int skip_id, double* __restrict__ e, double* __restrict__ q) {
    __shared__ double red[2][ALS_CTA / 32];
    __shared__ double s_delta;
    const int i = (int)order[first + blockIdx.x];
    if (i == skip_id) return;   // reference quirk: `0 until num_attribute` leaves the last slot alone
    const int a = colptr[i], b = colptr[i + 1];
    float* tp = theta_base + (int64_t)i * stride + off;
    const float theta_f = *tp;
    const double theta = (double)theta_f;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double sh = 0.0, se = 0.0;
    for (int p = a + tid; p < b; p += ALS_CTA) {
        int row;
        double x;
        als_entry<BINARY>(pay, p, row, x);
        const double hh = IS_V ? x * q[row] - x * x * theta : x;   // :56-58
        sh += hh * hh;
        se += e[row] * hh;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        sh += __shfl_xor_sync(0xffffffffu, sh, o);
        se += __shfl_xor_sync(0xffffffffu, se, o);
    }
    if (lane == 0) {
        red[0][warp] = sh;
        red[1][warp] = se;
    }
    __syncthreads();
    if (tid == 0) {
        double th = 0.0, te = 0.0;
        for (int w = 0; w < ALS_CTA / 32; ++w) {
            th += red[0][w];
            te += red[1][w];
        }
        double delta;
        *tp = als_theta(theta_f, reg, te, th, &delta);
        s_delta = delta;
    }
    __syncthreads();
    const double delta = s_delta;
    if (delta == 0.0) return;
    for (int p = a + tid; p < b; p += ALS_CTA) {
        int row;
        double x;
        als_entry<BINARY>(pay, p, row, x);
        if (IS_V) {
            const double qr = q[row];
            e[row] += (x * qr - x * x * theta) * delta;   // updateError :194-198
            q[row] = qr + x * delta;                      // :60-62
        } else {
            e[row] += x * delta;
        }
    }
}

// ---------------------------------------------------------------------------------------------
void als_free(sfm_handle* h) {
    AlsState* s = h->als;
    if (!s) return;
    if (s->exec) cudaGraphExecDestroy(s->exec);
    Buf* all[] = {&s->pay, &s->colptr, &s->order, &s->lvl, &s->lvl_sorted, &s->rowlvl, &s->e, &s->q,
                  &s->part};
    for (Buf* b : all)
        if (b->p) cudaFree(b->p);
    delete s;
    h->als = nullptr;
}

static int als_bits(int64_t n) {
    int b = 1;
    while (b < 32 && ((int64_t)1 << b) < n) ++b;
    return b;
}

static int als_build(sfm_handle* h, const BatchView& b) {
    als_free(h);
    AlsState* s = new (std::nothrow) AlsState;
    if (!s) return set_err(h, SFM_ERR_OOM, "host allocation failed");
    h->als = s;
    const ModelView& m = h->m;
    int64_t* L = &h->stats.kernel_launches;
    // rows below 2^30: the entry emitter ORs (row >> 30) above the feature bits of the sort key
    if (b.nnz >= 2147483647LL || b.n_rows >= (1LL << 30) || m.n_slots >= 2147483647LL)
        return set_err(h, SFM_ERR_ARG, "ALS: entries and slots must stay below 2^31, rows below 2^30");
    s->n_rows = b.n_rows;
    s->nnz = b.nnz;
    s->binary = b.val == nullptr;
    const int nnz = (int)b.nnz, n_slots = (int)m.n_slots;
    const size_t cnt = (size_t)(nnz > 0 ? nnz : 1);
    const size_t pay_sz = s->binary ? sizeof(uint32_t) : sizeof(uint2);
    RC(ensure(h, h->b_keys[0], sizeof(uint32_t) * cnt));
    RC(ensure(h, h->b_keys[1], sizeof(uint32_t) * cnt));
    RC(ensure(h, h->b_pay[0], pay_sz * cnt));
    RC(ensure(h, s->pay, pay_sz * cnt));
    RC(ensure(h, s->colptr, sizeof(int32_t) * ((size_t)n_slots + 1)));
    RC(ensure(h, s->order, sizeof(uint32_t) * (size_t)n_slots));
    RC(ensure(h, s->lvl, sizeof(uint32_t) * (size_t)n_slots));
    RC(ensure(h, s->lvl_sorted, sizeof(uint32_t) * (size_t)n_slots));
    RC(ensure(h, s->rowlvl, sizeof(int32_t) * (size_t)(b.n_rows > 0 ? b.n_rows : 1)));
    RC(ensure(h, s->e, sizeof(double) * (size_t)(b.n_rows > 0 ? b.n_rows : 1)));
    RC(ensure(h, s->q, sizeof(double) * (size_t)(b.n_rows > 0 ? b.n_rows : 1)));
    RC(ensure(h, s->part, sizeof(double) * (ALS_PARTS + 8)));
    const uint32_t* keys = (const uint32_t*)h->b_keys[1].p;
    // transposed input: entries sorted by (feature, row)
    if (nnz > 0) {
        const int key_bits = als_bits(m.n_slots);
        CU(launch_emit(b, key_bits, 30, m.n_slots, (uint32_t*)h->b_keys[0].p, (uint2*)h->b_pay[0].p,
                       h->sm_count, h->stream, L));
        const size_t sb = s->binary ? sort_pairs32_temp_bytes(nnz, key_bits)
                                    : sort_pairs_temp_bytes(nnz, key_bits);
        RC(ensure(h, h->b_sort_tmp, sb));
        if (s->binary)
            CU(sort_pairs32(h->b_sort_tmp.p, sb, (const uint32_t*)h->b_keys[0].p,
                            (uint32_t*)h->b_keys[1].p, (const uint32_t*)h->b_pay[0].p,
                            (uint32_t*)s->pay.p, nnz, key_bits, h->stream, L));
        else
            CU(sort_pairs(h->b_sort_tmp.p, sb, (const uint32_t*)h->b_keys[0].p,
                          (uint32_t*)h->b_keys[1].p, (const uint2*)h->b_pay[0].p, (uint2*)s->pay.p,
                          nnz, key_bits, h->stream, L));
    }
    als_colptr_kernel<<<(unsigned)((nnz + 1 + 255) / 256), 256, 0, h->stream>>>(
        keys, nnz, n_slots, (int32_t*)s->colptr.p);
    CU(cudaMemsetAsync(h->d_count, 0, sizeof(int32_t), h->stream));
    CU(cudaMemsetAsync(s->rowlvl.p, 0, sizeof(int32_t) * (size_t)(b.n_rows > 0 ? b.n_rows : 1),
                       h->stream));
    int32_t* d_max = h->d_slice;   // device scratch ints
    if (nnz > 1) {
        if (s->binary)
            als_dup_kernel<true><<<(unsigned)((nnz + 255) / 256), 256, 0, h->stream>>>(
                keys, s->pay.p, nnz, h->d_count);
        else
            als_dup_kernel<false><<<(unsigned)((nnz + 255) / 256), 256, 0, h->stream>>>(
                keys, s->pay.p, nnz, h->d_count);
    }
    if (s->binary)
        als_levels_kernel<true><<<1, 1024, 0, h->stream>>>((const int32_t*)s->colptr.p, s->pay.p,
                                                           n_slots, (int32_t*)s->rowlvl.p,
                                                           (uint32_t*)s->lvl.p, d_max);
    else
        als_levels_kernel<false><<<1, 1024, 0, h->stream>>>((const int32_t*)s->colptr.p, s->pay.p,
                                                            n_slots, (int32_t*)s->rowlvl.p,
                                                            (uint32_t*)s->lvl.p, d_max);
    *L += 3;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h->h_flags, h->d_count, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(h->h_flags + 1, d_max, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (h->h_flags[0])
        return set_err(h, SFM_ERR_ARG, "ALS: a row stores the same feature index twice");
    const int max_level = h->h_flags[1];
    // features by (level, id): stable sort of the levels with the id as payload
    {
        RC(ensure(h, h->b_keys[0], sizeof(uint32_t) * (size_t)(n_slots > nnz ? n_slots : cnt)));
        CU(launch_iota_u32((uint32_t*)h->b_keys[0].p, n_slots, h->stream, L));
        const int lbits = als_bits((int64_t)max_level + 1);
        const size_t sb = sort_pairs32_temp_bytes(n_slots, lbits);
        RC(ensure(h, h->b_sort_tmp, sb));
        CU(sort_pairs32(h->b_sort_tmp.p, sb, (const uint32_t*)s->lvl.p, (uint32_t*)s->lvl_sorted.p,
                        (const uint32_t*)h->b_keys[0].p, (uint32_t*)s->order.p, n_slots, lbits,
                        h->stream, L));
    }
    // start of every level in `order` (same boundary kernel: "slots" = levels)
    Buf lo;
    RC(ensure(h, lo, sizeof(int32_t) * ((size_t)max_level + 2)));
    als_colptr_kernel<<<(unsigned)((n_slots + 1 + 255) / 256), 256, 0, h->stream>>>(
        (const uint32_t*)s->lvl_sorted.p, n_slots, max_level + 1, (int32_t*)lo.p);
    ++*L;
    s->lvl_off.resize((size_t)max_level + 2);
    cudaError_t ce = cudaMemcpyAsync(s->lvl_off.data(), lo.p, sizeof(int32_t) * ((size_t)max_level + 2),
                                     cudaMemcpyDeviceToHost, h->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
    cudaFree(lo.p);
    CU(ce);
    s->built = true;
    return SFM_OK;
}

// Queues the kernels of one sweep on the compute stream (directly, or into a stream capture).
static cudaError_t als_enqueue(sfm_handle* h, AlsState* s, const BatchView& b, int32_t flags,
                               int64_t* n_launches) {
    const ModelView& m = h->m;
    const bool quirks = (flags & SFM_ALS_REF_QUIRKS) != 0;
    const int n_rows = (int)b.n_rows;
    const int skip_id = quirks ? (int)m.n_slots - 1 : -1;
    double* e = (double*)s->e.p;
    double* q = (double*)s->q.p;
    double* part = (double*)s->part.p;
    double* scal = part + ALS_PARTS;
    const int32_t* colptr = (const int32_t*)s->colptr.p;
    const uint32_t* order = (const uint32_t*)s->order.p;
    const int n_levels = (int)s->lvl_off.size() - 2;   // levels 1 .. n_levels hold columns
    int64_t rb64 = ((int64_t)n_rows + 7) / 8;
    if (rb64 > (int64_t)h->sm_count * 16) rb64 = (int64_t)h->sm_count * 16;
    const unsigned rb = (unsigned)(rb64 > 0 ? rb64 : 1);
    int64_t nl = 0;
    als_residual_kernel<<<rb, 256, 0, h->stream>>>(m.v, m.w, m.w0, m.k, m.kp, m.k0, m.k1, b.row_ptr,
                                                   b.idx, b.val, b.label, b.uniform_m, n_rows, e);
    ++nl;
    if (m.k0) {
        als_sum_kernel<false><<<ALS_PARTS, 256, 0, h->stream>>>(e, n_rows, part);
        als_w0_kernel<<<1, 32, 0, h->stream>>>(part, m.w0, (double)h->cfg.reg0, n_rows, 1, scal);
        als_shift_kernel<<<(unsigned)(h->sm_count * 4), 256, 0, h->stream>>>(e, n_rows, scal);
        nl += 3;
    }
#define ALS_LEVELS(ISV, BASE, STRIDE, OFF, REG)                                                   \
    for (int l = 1; l <= n_levels; ++l) {                                                         \
        const int first = s->lvl_off[(size_t)l], cntl = s->lvl_off[(size_t)l + 1] - first;        \
        if (cntl <= 0) continue;                                                                  \
        if (s->binary)                                                                            \
            als_level_kernel<ISV, true><<<(unsigned)cntl, ALS_CTA, 0, h->stream>>>(              \
                order, first, colptr, s->pay.p, BASE, STRIDE, OFF, REG, skip_id, e, q);           \
        else                                                                                      \
            als_level_kernel<ISV, false><<<(unsigned)cntl, ALS_CTA, 0, h->stream>>>(             \
                order, first, colptr, s->pay.p, BASE, STRIDE, OFF, REG, skip_id, e, q);           \
        ++nl;                                                                                     \
    }
    if (m.k1) ALS_LEVELS(false, m.w, 1, 0, (double)h->cfg.regw)
    for (int f = 0; f < m.k; ++f) {
        als_q_kernel<<<(unsigned)(h->sm_count * 8), 256, 0, h->stream>>>(
            m.v, m.kp, f, b.row_ptr, b.idx, b.val, b.uniform_m, n_rows, q);
        ++nl;
        ALS_LEVELS(true, m.v, m.kp, f, (double)h->cfg.regv)
    }
#undef ALS_LEVELS
    als_sum_kernel<true><<<ALS_PARTS, 256, 0, h->stream>>>(e, n_rows, part);
    als_w0_kernel<<<1, 32, 0, h->stream>>>(part, nullptr, 0.0, n_rows, 0, scal);
    nl += 2;
    *n_launches = nl;
    return cudaGetLastError();
}

// One ALS sweep over the resident data set `b` (every row, identity order).  The sweep is
// (k + 1) * levels tiny dependent kernels: they are captured once into a CUDA graph per
// (data set, flags, regularisation) and replayed (SFM_ALS_GRAPH=0: plain stream launches).
int als_sweep(sfm_handle* h, const BatchView& b, int32_t flags, double* rmse_out) {
    const ModelView& m = h->m;
    if (m.k > 128) return set_err(h, SFM_ERR_ARG, "ALS: at most 128 factors");
    if (!h->als || !h->als->built || h->als->n_rows != b.n_rows || h->als->nnz != b.nnz)
        RC(als_build(h, b));
    AlsState* s = h->als;
    const int n_rows = (int)b.n_rows;
    const char* genv = getenv("SFM_ALS_GRAPH");
    const bool want_graph = !(genv && genv[0] == '0') && !s->graph_failed;
    const float key[4] = {(float)flags, h->cfg.reg0, h->cfg.regw, h->cfg.regv};
    int64_t nl = 0;
    if (want_graph) {
        if (!s->exec || memcmp(key, s->graph_key, sizeof key) != 0) {
            if (s->exec) cudaGraphExecDestroy(s->exec);
            s->exec = nullptr;
            cudaGraph_t g = nullptr;
            cudaError_t ce = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
            if (ce == cudaSuccess) {
                const cudaError_t ke = als_enqueue(h, s, b, flags, &nl);
                ce = cudaStreamEndCapture(h->stream, &g);
                if (ce == cudaSuccess) ce = ke;
            }
            if (ce == cudaSuccess) ce = cudaGraphInstantiate(&s->exec, g, 0);
            if (g) cudaGraphDestroy(g);
            if (ce != cudaSuccess) {   // fall back to stream launches for this data set
                cudaGetLastError();
                s->exec = nullptr;
                s->graph_failed = true;
            } else {
                memcpy(s->graph_key, key, sizeof key);
                s->graph_launches = nl;
            }
        }
    }
    if (want_graph && s->exec) {
        CU(cudaGraphLaunch(s->exec, h->stream));
        nl = s->graph_launches;
    } else {
        CU(als_enqueue(h, s, b, flags, &nl));
    }
    h->stats.kernel_launches += nl;
    double* scal = (double*)s->part.p + ALS_PARTS;
    CU(cudaMemcpyAsync(h->h_scal, scal, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (rmse_out) *rmse_out = n_rows > 0 ? sqrt(h->h_scal[0] / (double)n_rows) : 0.0;
    return SFM_OK;
}

// residuals of the last sweep (tests, hosts that keep the reference's `e` map)
int als_residuals(sfm_handle* h, double* out, int64_t n) {
    if (!h->als || !h->als->built || h->als->n_rows != n)
        return set_err(h, SFM_ERR_STATE, "no ALS sweep has run on this data set");
    CU(cudaMemcpyAsync(out, h->als->e.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost,
                       h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return SFM_OK;
}

}  // namespace sfm
