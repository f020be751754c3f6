// sfm_shard.cu -- row-sharded V (BASELINE config 4 / north_star: "when the feature count exceeds
// one GPU, V is row-sharded with an all-to-all gather of the touched rows").
//
// Ownership: rank g owns the features [g*n_per, (g+1)*n_per), n_per = ceil(n_slots / G), and holds
// only those rows of V and w (plus the always-zero row).  w0 is replicated.  Rows of the data set
// stay data-parallel as in the replicated mode.  One iteration on every rank:
//
//   1. emit the batch's (feature, row) entries, radix sort by feature        (as in replicated mode)
//   2. run starts of the sorted keys -> the batch's unique features `uniq` (ascending) and the
//      compact id of every entry; owner boundaries by binary search (ranges are contiguous)
//   3. all-gather the G x G request counts; all-to-all the requested ids; owners gather the rows
//      and all-to-all them back  -> compact tables T_v[U][kp], T_w[U]
//   4. the batch is re-indexed to compact ids and the ordinary forward + pull kernels run against
//      the compact tables (nothing in them knows about sharding)
//   5. the compact gradient rows travel back to their owners (all-to-all); each owner adds the
//      contributions in RANK ORDER (one kernel per source: deterministic) into its dense shard
//      gradient and applies the dense regularised update to its own rows.
//
// The exchange is NCCL send/recv over NVLink; the dense update and the all-reduce volume drop by G.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <cub/cub.cuh>

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

static inline int64_t grid_for(int64_t n, int threads = 256) {
    int64_t b = (n + threads - 1) / threads;
    if (b > (int64_t)device_sm_count() * 32) b = (int64_t)device_sm_count() * 32;
    return b < 1 ? 1 : b;
}

// flags[p] = 1 at the first entry of every run of equal keys
__global__ void run_flags_kernel(const uint32_t* __restrict__ keys, int64_t n,
                                 uint32_t* __restrict__ flags) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride)
        flags[p] = (p == 0 || keys[p] != keys[p - 1]) ? 1u : 0u;
}

// crank = inclusive scan of flags; compact id of entry p = crank[p] - 1; uniq[id] = key at run start
__global__ void uniq_scatter_kernel(const uint32_t* __restrict__ keys,
                                    const uint32_t* __restrict__ flags, uint32_t* __restrict__ crank,
                                    int64_t n, int32_t* __restrict__ uniq,
                                    int32_t* __restrict__ n_uniq) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        const uint32_t c = crank[p] - 1u;
        crank[p] = c;
        if (flags[p]) uniq[c] = (int32_t)keys[p];
        if (p == n - 1) *n_uniq = (int32_t)(c + 1u);
    }
}

// bounds[o] = first index of uniq with value >= o * n_per  (o = 0..G); cnt[o] = bounds[o+1]-bounds[o]
__global__ void owner_bounds_kernel(const int32_t* __restrict__ uniq, const int32_t* __restrict__ n_uniq,
                                    int64_t n_per, int world, int32_t* __restrict__ bounds,
                                    int32_t* __restrict__ cnt) {
    __shared__ int32_t sb[1026];
    const int U = *n_uniq;
    for (int o = threadIdx.x; o <= world; o += blockDim.x) {
        const int64_t target = (int64_t)o * n_per;
        int lo = 0, hi = U;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((int64_t)uniq[mid] < target) lo = mid + 1; else hi = mid;
        }
        sb[o] = o == world ? U : lo;
    }
    __syncthreads();
    for (int o = threadIdx.x; o <= world; o += blockDim.x) {
        bounds[o] = sb[o];
        if (o < world) cnt[o] = sb[o + 1] - sb[o];
    }
}

// owner side: rows requested by the peers -> contiguous send buffers
__global__ void gather_rows_kernel(const float4* __restrict__ V4, const float* __restrict__ W,
                                   const int32_t* __restrict__ req, int64_t n_req, int64_t own_lo,
                                   int lpr, float4* __restrict__ out_v, float* __restrict__ out_w) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t total = n_req * lpr;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t j = e / lpr;
        const int q = (int)(e % lpr);
        const int64_t loc = (int64_t)req[j] - own_lo;
        out_v[e] = V4[loc * lpr + q];
        if (q == 0) out_w[j] = W[loc];
    }
}

__global__ void lut_scatter_kernel(const int32_t* __restrict__ uniq, int64_t U,
                                   int32_t* __restrict__ lut) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < U; c += stride)
        lut[uniq[c]] = (int32_t)c;
}

// compact copy of the batch: bidx = lut[idx], bval, blabel in batch order (warp per row)
__global__ void __launch_bounds__(256)
remap_batch_kernel(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ idx,
                   const float* __restrict__ val, const float* __restrict__ label,
                   const int32_t* __restrict__ row_ids, int64_t row_lo, int64_t n_rows,
                   const int64_t* __restrict__ out_ptr, int64_t out_base, int uniform_m,
                   const int32_t* __restrict__ lut, int32_t* __restrict__ bidx,
                   float* __restrict__ bval, float* __restrict__ blabel) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t pos = warp0; pos < n_rows; pos += nwarps) {
        const int64_t r = row_ids ? (int64_t)__ldg(row_ids + pos) : row_lo + pos;
        int64_t beg, end;
        if (uniform_m >= 0) {
            beg = r * uniform_m;
            end = beg + uniform_m;
        } else {
            beg = __ldg(row_ptr + r);
            end = __ldg(row_ptr + r + 1);
        }
        const int64_t obase = out_ptr ? __ldg(out_ptr + pos) - out_base : pos * (int64_t)uniform_m;
        for (int64_t j = beg + lane; j < end; j += 32) {
            bidx[obase + (j - beg)] = __ldg(lut + __ldg(idx + j));
            if (bval) bval[obase + (j - beg)] = __ldg(val + j);
        }
        if (lane == 0 && blabel && label) blabel[pos] = __ldg(label + r);
    }
}

// owner side: add one source rank's gradient rows into the dense shard gradient (ids unique per source)
__global__ void scatter_add_rows_kernel(const float4* __restrict__ gv, const float* __restrict__ gw,
                                        const int32_t* __restrict__ req, int64_t n_req,
                                        int64_t own_lo, int lpr, float4* __restrict__ acc_v,
                                        float* __restrict__ acc_w) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t total = n_req * lpr;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t j = e / lpr;
        const int q = (int)(e % lpr);
        const int64_t loc = (int64_t)req[j] - own_lo;
        float4 a = acc_v[loc * lpr + q];
        const float4 g = gv[e];
        a.x += g.x; a.y += g.y; a.z += g.z; a.w += g.w;
        acc_v[loc * lpr + q] = a;
        if (q == 0) acc_w[loc] += gw[j];
    }
}

__global__ void set_tail_kernel(float* __restrict__ dst, const double* __restrict__ d_scal, int k0) {
    *dst = k0 ? (float)d_scal[SC_GW0] : 0.f;
}

static int bits_for64(int64_t n) {
    int b = 1;
    while (b < 32 && ((int64_t)1 << b) < n) ++b;
    return b;
}

static void free_b(Buf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

static void free_shard_batch(ShardBatch& sb) {
    Buf* bs[] = {&sb.crank, &sb.pay, &sb.uniq, &sb.req, &sb.bidx, &sb.bval, &sb.blabel, &sb.optr};
    for (Buf* b : bs) free_b(*b);
    sb.built = false;
}

void shard_clear_cache(sfm_handle* h) {
    if (!h->shard) return;
    for (ShardBatch& sb : h->shard->cached) free_shard_batch(sb);
    h->shard->cached.clear();
}

// what the compute kernels need for one iteration
struct ShardPlan {
    ModelView mc;                          // compact model view (T_v, T_w)
    BatchView bc;                          // compact batch view
};

// ------------------------------------------------------------------------------------------
// Steps 1, 2, 3a, 4a: everything that depends on the batch's rows only.
// ------------------------------------------------------------------------------------------
static int shard_build(sfm_handle* h, const BatchView& b, bool need_label, ShardBatch& sb) {
    ShardState& s = *h->shard;
    const ModelView& m = h->m;
    int64_t* L = &h->stats.kernel_launches;
    const int G = h->world;
    const int64_t n = b.n_rows, E = b.nnz;
    const bool binary = b.val == nullptr;
    const size_t pay_sz = binary ? sizeof(uint32_t) : sizeof(uint2);
    const size_t cntE = (size_t)(E > 0 ? E : 1);
    RC(ensure(h, h->b_keys[0], sizeof(uint32_t) * cntE));
    RC(ensure(h, h->b_pay[0], pay_sz * cntE));
    RC(ensure(h, s.keys_tmp, sizeof(uint32_t) * cntE));
    RC(ensure(h, s.flags, sizeof(uint32_t) * cntE));
    RC(ensure(h, sb.pay, pay_sz * cntE));
    RC(ensure(h, sb.crank, sizeof(uint32_t) * cntE));
    RC(ensure(h, sb.uniq, sizeof(int32_t) * cntE));
    RC(ensure(h, s.small, sizeof(int32_t) * (size_t)(4 + 2 * (G + 1) + G * G)));
    RC(ensure(h, s.lut, sizeof(int32_t) * (size_t)m.n_slots));
    int32_t* d_nu = (int32_t*)s.small.p;           // [0]   n_uniq
    int32_t* d_bounds = d_nu + 4;                  // [G+1]
    int32_t* d_cnt = d_bounds + (G + 1);           // [G]
    int32_t* d_all = d_cnt + (G + 1);              // [G*G]
    const int key_bits = bits_for64(m.n_slots);

    // entry offsets of the batch rows
    BatchView bb = b;
    if (b.uniform_m < 0 && b.row_ids) {   // ragged sampled rows: own copy of the offsets
        RC(ensure(h, sb.optr, sizeof(int64_t) * (size_t)(n + 1)));
        RC(ensure(h, h->b_lens, sizeof(int64_t) * (size_t)(n + 1)));
        CU(cudaMemsetAsync((int64_t*)h->b_lens.p + n, 0, sizeof(int64_t), h->stream));
        CU(launch_row_lens(b.row_ptr, b.row_ids, n, (int64_t*)h->b_lens.p, h->stream, L));
        const size_t tb = scan_temp_bytes(n + 1);
        RC(ensure(h, h->b_sel_tmp, tb));
        CU(exclusive_scan_i64(h->b_sel_tmp.p, tb, (const int64_t*)h->b_lens.p, (int64_t*)sb.optr.p,
                              n + 1, h->stream, L));
        bb.out_ptr = (const int64_t*)sb.optr.p;
        bb.out_base = 0;
    }
    // 1. entries sorted by feature
    CU(cudaMemsetAsync(d_nu, 0, sizeof(int32_t) * 4, h->stream));
    if (E > 0) {
        CU(launch_emit(bb, key_bits, 30, m.n_slots, (uint32_t*)h->b_keys[0].p, (uint2*)h->b_pay[0].p,
                       h->sm_count, h->stream, L));
        const size_t tb = binary ? sort_pairs32_temp_bytes(E, key_bits) : sort_pairs_temp_bytes(E, key_bits);
        RC(ensure(h, h->b_sort_tmp, tb));
        if (binary)
            CU(sort_pairs32(h->b_sort_tmp.p, tb, (const uint32_t*)h->b_keys[0].p,
                            (uint32_t*)s.keys_tmp.p, (const uint32_t*)h->b_pay[0].p,
                            (uint32_t*)sb.pay.p, E, key_bits, h->stream, L));
        else
            CU(sort_pairs(h->b_sort_tmp.p, tb, (const uint32_t*)h->b_keys[0].p,
                          (uint32_t*)s.keys_tmp.p, (const uint2*)h->b_pay[0].p, (uint2*)sb.pay.p, E,
                          key_bits, h->stream, L));
        // 2. unique features + compact ids
        const uint32_t* keys1 = (const uint32_t*)s.keys_tmp.p;
        run_flags_kernel<<<(unsigned)grid_for(E), 256, 0, h->stream>>>(keys1, E, (uint32_t*)s.flags.p);
        size_t tb2 = 0;
        cub::DeviceScan::InclusiveSum(nullptr, tb2, (const uint32_t*)nullptr, (uint32_t*)nullptr, E);
        RC(ensure(h, h->b_sel_tmp, tb2));
        CU(cub::DeviceScan::InclusiveSum(h->b_sel_tmp.p, tb2, (const uint32_t*)s.flags.p,
                                         (uint32_t*)sb.crank.p, E, h->stream));
        uniq_scatter_kernel<<<(unsigned)grid_for(E), 256, 0, h->stream>>>(
            keys1, (const uint32_t*)s.flags.p, (uint32_t*)sb.crank.p, E, (int32_t*)sb.uniq.p, d_nu);
        *L += 4;
    }
    owner_bounds_kernel<<<1, 256, 0, h->stream>>>((const int32_t*)sb.uniq.p, d_nu, s.n_per, G,
                                                  d_bounds, d_cnt);
    ++*L;
    CU(cudaGetLastError());
    // 3a. everybody learns everybody's request counts (the one host sync of a build)
    RC(nccl_allgather_i32(h->nccl, h->comm, d_cnt, d_all, (size_t)G, h->stream, &h->err));
    const size_t hs = sizeof(int32_t) * (size_t)(4 + 2 * (G + 1) + G * G);
    if (s.h_small_cap < hs) {
        if (s.h_small) cudaFreeHost(s.h_small);
        CU(cudaMallocHost(&s.h_small, hs));
        s.h_small_cap = hs;
    }
    CU(cudaMemcpyAsync(s.h_small, s.small.p, hs, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    const int32_t* hv = s.h_small;
    const int32_t* h_bounds = hv + 4;
    const int32_t* h_all = hv + 4 + 2 * (G + 1);
    sb.U = hv[0];
    sb.send_off.assign(G, 0); sb.send_cnt.assign(G, 0);
    sb.recv_off.assign(G, 0); sb.recv_cnt.assign(G, 0);
    int64_t R = 0;
    for (int p = 0; p < G; ++p) {
        sb.send_off[p] = h_bounds[p];
        sb.send_cnt[p] = h_bounds[p + 1] - h_bounds[p];
        sb.recv_off[p] = R;
        sb.recv_cnt[p] = h_all[p * G + h->rank];
        R += sb.recv_cnt[p];
    }
    sb.R = R;
    const int64_t U = sb.U;
    RC(ensure(h, sb.req, sizeof(int32_t) * (size_t)(R > 0 ? R : 1)));
    RC(nccl_alltoallv_4b(h->nccl, h->comm, h->rank, G, sb.uniq.p, sb.send_off.data(),
                         sb.send_cnt.data(), sb.req.p, sb.recv_off.data(), sb.recv_cnt.data(), 1,
                         h->stream, &h->err));
    // 4a. compact copy of the batch
    RC(ensure(h, sb.bidx, sizeof(int32_t) * cntE));
    if (!binary) RC(ensure(h, sb.bval, sizeof(float) * cntE));
    RC(ensure(h, sb.blabel, sizeof(float) * (size_t)(n > 0 ? n : 1)));
    if (U > 0) {
        lut_scatter_kernel<<<(unsigned)grid_for(U), 256, 0, h->stream>>>((const int32_t*)sb.uniq.p, U,
                                                                         (int32_t*)s.lut.p);
        ++*L;
    }
    if (n > 0) {
        int64_t blocks = (n + 7) / 8;
        if (blocks > (int64_t)h->sm_count * 16) blocks = (int64_t)h->sm_count * 16;
        remap_batch_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(
            bb.row_ptr, bb.idx, bb.val, need_label ? bb.label : nullptr, bb.row_ids, bb.row_lo, n,
            bb.out_ptr, bb.out_base, bb.uniform_m, (const int32_t*)s.lut.p, (int32_t*)sb.bidx.p,
            binary ? nullptr : (float*)sb.bval.p, need_label ? (float*)sb.blabel.p : nullptr);
        ++*L;
    }
    CU(cudaGetLastError());
    sb.built = true;
    return SFM_OK;
}

// ------------------------------------------------------------------------------------------
// Step 3b: owners gather the requested rows, all-to-all -> compact tables; fills the plan.
// ------------------------------------------------------------------------------------------
static int shard_fetch(sfm_handle* h, const BatchView& b, const ShardBatch& sb, ShardPlan* sp) {
    ShardState& s = *h->shard;
    const ModelView& m = h->m;
    int64_t* L = &h->stats.kernel_launches;
    const int G = h->world;
    const int64_t U = sb.U, R = sb.R, E = b.nnz;
    const bool binary = b.val == nullptr;
    RC(ensure(h, s.out_v, sizeof(float) * (size_t)(R > 0 ? R : 1) * m.kp));
    RC(ensure(h, s.out_w, sizeof(float) * (size_t)(R > 0 ? R : 1)));
    RC(ensure(h, s.t_v, sizeof(float) * (size_t)(U + 1) * m.kp));
    RC(ensure(h, s.t_w, sizeof(float) * (size_t)(U + 1)));
    if (R > 0) {
        gather_rows_kernel<<<(unsigned)grid_for(R * m.lpr), 256, 0, h->stream>>>(
            (const float4*)s.v, s.w, (const int32_t*)sb.req.p, R, s.own_lo, m.lpr, (float4*)s.out_v.p,
            (float*)s.out_w.p);
        ++*L;
    }
    RC(nccl_alltoallv_4b(h->nccl, h->comm, h->rank, G, s.out_v.p, sb.recv_off.data(),
                         sb.recv_cnt.data(), s.t_v.p, sb.send_off.data(), sb.send_cnt.data(), m.kp,
                         h->stream, &h->err));
    RC(nccl_alltoallv_4b(h->nccl, h->comm, h->rank, G, s.out_w.p, sb.recv_off.data(),
                         sb.recv_cnt.data(), s.t_w.p, sb.send_off.data(), sb.send_cnt.data(), 1,
                         h->stream, &h->err));
    CU(cudaMemsetAsync((float*)s.t_v.p + (size_t)U * m.kp, 0, sizeof(float) * m.kp, h->stream));
    CU(cudaMemsetAsync((float*)s.t_w.p + U, 0, sizeof(float), h->stream));
    sp->mc = m;
    sp->mc.v = (float*)s.t_v.p;
    sp->mc.w = (float*)s.t_w.p;
    sp->mc.n_slots = U;
    BatchView& bc = sp->bc;
    bc = b;
    bc.idx = (const int32_t*)sb.bidx.p;
    bc.val = binary ? nullptr : (const float*)sb.bval.p;
    bc.label = (const float*)sb.blabel.p;
    bc.row_ids = nullptr;
    bc.row_lo = 0;
    bc.idx_len = E;
    bc.out_base = 0;
    bc.validated = true;
    if (b.uniform_m >= 0) {
        bc.row_ptr = nullptr;      // rows are pos*m .. pos*m+m in the compact copy
        bc.out_ptr = nullptr;
    } else if (b.row_ids) {
        bc.row_ptr = (const int64_t*)sb.optr.p;   // offsets computed by shard_build
        bc.out_ptr = bc.row_ptr;
    } else {
        // contiguous rows: the data set's own row pointers double as the compact CSR pointers;
        // they carry a base offset, so the compact arrays are addressed relative to it
        bc.row_ptr = b.out_ptr;
        bc.out_ptr = b.out_ptr;
        bc.out_base = b.out_base;
        bc.idx = (const int32_t*)sb.bidx.p - b.out_base;
        if (!binary) bc.val = (const float*)sb.bval.p - b.out_base;
        bc.idx_len = E + b.out_base;
    }
    return SFM_OK;
}

// Forward only (predict / evaluate) on a sharded model: yhat -> h->b_yhat.
int shard_forward(sfm_handle* h, const BatchView& b) {
    ShardBatch& sb = h->shard->scratch;
    ShardPlan sp;
    RC(shard_build(h, b, false, sb));
    RC(shard_fetch(h, b, sb, &sp));
    RC(ensure(h, h->b_yhat, sizeof(float) * (size_t)(b.n_rows > 0 ? b.n_rows : 1)));
    FwdOut o;
    memset(&o, 0, sizeof o);
    o.yhat = (float*)h->b_yhat.p;
    CU(launch_forward(sp.mc, sp.bc, o, false, h->d_err, h->sm_count, h->stream,
                      &h->stats.kernel_launches));
    return SFM_OK;
}

// One SGD iteration on a sharded model (asynchronous after its internal count sync).
int shard_train(sfm_handle* h, const BatchView& b, int64_t iter, int cache_slot) {
    ShardState& s = *h->shard;
    const ModelView& m = h->m;
    int64_t* L = &h->stats.kernel_launches;
    const int G = h->world;
    const int64_t n = b.n_rows, E = b.nnz;
    if ((n + 1) * (int64_t)m.lpr >= 4294967296LL)
        return set_err(h, SFM_ERR_ARG, "batch too large: rows * kp/4 must stay below 2^32");
    CU(cudaMemsetAsync(h->d_err, 0, sizeof(int32_t), h->stream));
    // fixed mini-batches (PARTITION sampler) keep their row-dependent plan; sampled ones rebuild it
    ShardBatch* sbp = &s.scratch;
    if (cache_slot >= 0) {
        if ((size_t)cache_slot >= s.cached.size()) s.cached.resize((size_t)cache_slot + 1);
        sbp = &s.cached[(size_t)cache_slot];
    }
    ShardBatch& sb = *sbp;
    if (cache_slot < 0 || !sb.built) RC(shard_build(h, b, true, sb));
    ShardPlan sp;
    RC(shard_fetch(h, b, sb, &sp));
    const int64_t U = sb.U, R = sb.R;
    const bool binary = b.val == nullptr;
    RC(ensure(h, h->b_S, sizeof(float) * (size_t)(n > 0 ? n : 1) * m.kp));
    RC(ensure(h, h->b_mult, sizeof(float) * (size_t)(n > 0 ? n : 1)));
    RC(ensure(h, h->b_loss, sizeof(float) * (size_t)(n > 0 ? n : 1)));
    RC(ensure(h, h->b_partials, sizeof(double) * 4 * 512));
    // 4b. forward + loss on the compact model
    FwdOut o;
    memset(&o, 0, sizeof o);
    o.S = (float*)h->b_S.p;
    o.mult = (float*)h->b_mult.p;
    o.loss = (float*)h->b_loss.p;
    CU(launch_forward(sp.mc, sp.bc, o, true, h->d_err, h->sm_count, h->stream, L));
    CU(launch_scalar_reduce(o.loss, o.mult, n, (double*)h->b_partials.p,
                            reinterpret_cast<unsigned int*>(h->d_count + 3), h->d_scal, h->d_err,
                            h->stream, L));
    RC(nccl_allreduce_f64(h->nccl, h->comm, h->d_scal, SC_N, h->stream, &h->err));
    // 4c. compact gradient [gV U*kp | gw U | gw0]
    const int64_t Uc = U > 0 ? U : 1;
    RC(ensure(h, h->b_seg, sizeof(int32_t) * 2 * (size_t)Uc));
    ModelView mc = sp.mc;
    mc.n_slots = Uc;   // U == 0: one dummy slot (the zero row), nothing is sent anyway
    RC(ensure(h, h->b_pull, pull_scratch_bytes(mc, E, 1)));
    RC(ensure(h, h->b_grad, sizeof(float) * ((size_t)Uc * (m.kp + 1) + 1)));
    UpdateParams up;
    up.eta = (float)((double)h->cfg.step_size / sqrt((double)(iter < 1 ? 1 : iter)));
    up.reg0 = h->cfg.reg0;
    up.regw = h->cfg.regw;
    up.regv = h->cfg.regv;
    CU(launch_pull(mc, (int32_t*)h->b_seg.p, 31, 1, (const uint32_t*)sb.crank.p,
                   (const uint2*)sb.pay.p, E, binary, o.S, o.mult,
                   (float*)h->b_pull.p, h->d_scal, h->d_err, up, false, (float*)h->b_grad.p,
                   h->sm_count, h->stream, L));
    // 5. gradient rows back to their owners, summed in rank order, dense update of the own shard
    RC(ensure(h, s.gr_v, sizeof(float) * (size_t)(R > 0 ? R : 1) * m.kp));
    RC(ensure(h, s.gr_w, sizeof(float) * (size_t)(R > 0 ? R : 1)));
    const size_t acc_len = (size_t)s.n_own * (m.kp + 1) + 1;
    RC(ensure(h, s.acc, sizeof(float) * acc_len));
    float* gcv = (float*)h->b_grad.p;
    float* gcw = gcv + (size_t)Uc * m.kp;
    RC(nccl_alltoallv_4b(h->nccl, h->comm, h->rank, G, gcv, sb.send_off.data(), sb.send_cnt.data(),
                         s.gr_v.p, sb.recv_off.data(), sb.recv_cnt.data(), m.kp, h->stream, &h->err));
    RC(nccl_alltoallv_4b(h->nccl, h->comm, h->rank, G, gcw, sb.send_off.data(), sb.send_cnt.data(),
                         s.gr_w.p, sb.recv_off.data(), sb.recv_cnt.data(), 1, h->stream, &h->err));
    CU(cudaMemsetAsync(s.acc.p, 0, sizeof(float) * acc_len, h->stream));
    float* acc_v = (float*)s.acc.p;
    float* acc_w = acc_v + (size_t)s.n_own * m.kp;
    for (int src = 0; src < G; ++src) {   // rank order: the summation order is fixed
        const int64_t c = sb.recv_cnt[src];
        if (c <= 0) continue;
        scatter_add_rows_kernel<<<(unsigned)grid_for(c * m.lpr), 256, 0, h->stream>>>(
            (const float4*)s.gr_v.p + sb.recv_off[src] * m.lpr, (const float*)s.gr_w.p + sb.recv_off[src],
            (const int32_t*)sb.req.p + sb.recv_off[src], c, s.own_lo, m.lpr, (float4*)acc_v, acc_w);
        ++*L;
    }
    set_tail_kernel<<<1, 1, 0, h->stream>>>(acc_w + s.n_own, h->d_scal, m.k0);
    ++*L;
    CU(cudaGetLastError());
    ModelView own = m;
    own.v = s.v;
    own.w = s.w;
    own.n_slots = s.n_own;
    if (s.n_own > 0 || m.k0)
        CU(launch_update(own, (const float*)s.acc.p, h->d_scal, h->d_err, up, h->stream, L));
    h->stats.train_steps += 1;
    h->stats.train_rows += n;
    h->stats.train_nnz += E;
    return SFM_OK;
}

}  // namespace sfm
