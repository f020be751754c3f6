// sfm_api.cu -- the C ABI of include/sparkfm_b200.h: handle management, host<->device plumbing
// and the per-iteration launch sequence.  No arithmetic of the hot path happens on the host.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <string>
#include <vector>

#include "sfm_common.h"

using namespace sfm;

namespace sfm {

static Knobs g_knobs;
const Knobs& knobs() { return g_knobs; }
void knobs_refresh() {
    Knobs k;
    const char* e;
    if ((e = getenv("SFM_SORT")) && e[0] == 'c') k.sort_cub = true;
    if (((e = getenv("SFM_BUCKET")) && e[0] == '0') || k.sort_cub) k.bucket = false;
    if (getenv("SFM_NO_FASTPATH")) k.no_fastpath = true;
    if ((e = getenv("SFM_PULL_BLOCK_MB"))) k.pull_block_mb = atoi(e);
    if ((e = getenv("SFM_AR_SLICES"))) k.ar_slices = atoi(e) < 1 ? 1 : (atoi(e) > 8 ? 8 : atoi(e));
    if ((e = getenv("SFM_SORT_AHEAD"))) k.sort_ahead = atoi(e);
    if ((e = getenv("SFM_BUCKET_CACHE")) && e[0] == '1') k.bucket_cache = true;
    if ((e = getenv("SFM_P2P_SPARSE")) && e[0] == '0') k.p2p_sparse = false;
    if ((e = getenv("SFM_GRAPH")) && e[0] == '0') k.step_graph = false;
    if ((e = getenv("SFM_STREAM_PRIO"))) k.stream_prio = e[0] != '0';
    if ((e = getenv("SFM_PLAN_AHEAD"))) k.plan_ahead = e[0] != '0';
    g_knobs = k;
}

int set_err(sfm_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}

static int cuda_fail(sfm_handle* h, cudaError_t e, const char* what) {
    std::string m = what;
    m += ": ";
    m += cudaGetErrorString(e);
    return set_err(h, e == cudaErrorMemoryAllocation ? SFM_ERR_OOM : SFM_ERR_CUDA, m);
}

#define CU(call)                                                       \
    do {                                                               \
        cudaError_t e_ = (call);                                       \
        if (e_ != cudaSuccess) return cuda_fail(h, e_, #call);         \
    } while (0)

#define RC(call)                     \
    do {                             \
        int rc_ = (call);            \
        if (rc_ != SFM_OK) return rc_; \
    } while (0)

int ensure(sfm_handle* h, Buf& b, size_t bytes) {
    if (bytes <= b.cap) return SFM_OK;
    if (h->capturing) {   // no allocation inside a stream capture: the step is re-run uncaptured
        h->capture_abort = true;
        return SFM_ERR_STATE;
    }
    if (b.p) {
        cudaStreamSynchronize(h->stream);
        cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;  // a little slack so slowly growing batches settle
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        want = bytes;
        e = cudaMalloc(&b.p, want);
    }
    if (e != cudaSuccess) {
        b.p = nullptr;
        return cuda_fail(h, e, "cudaMalloc(scratch)");
    }
    b.cap = want;
    return SFM_OK;
}

static void free_buf(Buf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

static int ensure_pinned(sfm_handle* h, size_t bytes) {
    if (bytes <= h->h_pinned_cap) return SFM_OK;
    if (h->h_pinned) {
        cudaStreamSynchronize(h->stream);
        cudaFreeHost(h->h_pinned);
        h->h_pinned = nullptr;
        h->h_pinned_cap = 0;
    }
    const size_t want = bytes + bytes / 4 + 4096;
    CU(cudaMallocHost(&h->h_pinned, want));
    h->h_pinned_cap = want;
    return SFM_OK;
}

static int kp_for(int k) {  // pad k to 4 * 2^j so a V row is a power-of-two number of float4
    int q = (k + 3) / 4;
    if (q < 1) q = 1;
    int p = 1;
    while (p < q) p <<= 1;
    return 4 * p;
}

static int bits_for(int64_t n_slots) {
    int b = 1;
    while (b < 32 && ((int64_t)1 << b) < n_slots) ++b;
    return b;
}

struct PhaseTimer {
    sfm_handle* h;
    explicit PhaseTimer(sfm_handle* hh) : h(hh) {
        if (h->phase_timing) cudaEventRecord(h->ev_a, h->stream);
    }
    void lap(double* acc) {
        if (!h->phase_timing) return;
        cudaEventRecord(h->ev_b, h->stream);
        cudaEventSynchronize(h->ev_b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev_a, h->ev_b);
        *acc += ms;
        if (acc != &h->stats.ms_predict) h->stats.ms_total_train += ms;
        cudaEventRecord(h->ev_a, h->stream);
    }
};

static UpdateParams update_params(const sfm_handle* h, int64_t iter) {
    UpdateParams up;
    up.eta = (float)((double)h->cfg.step_size / sqrt((double)(iter < 1 ? 1 : iter)));
    up.reg0 = h->cfg.reg0;
    up.regw = h->cfg.regw;
    up.regv = h->cfg.regv;
    return up;
}

static size_t grad_len(const sfm_handle* h) {
    return (size_t)h->m.n_slots * (size_t)(h->m.kp + 1) + 1;
}

// Upload an unpadded [n_slots][k] fp32 host matrix into the padded device V.
// Uploads `rows` rows of an unpadded [rows][k] fp32 host matrix into a padded device matrix.
static int upload_v_rows(sfm_handle* h, float* dst, const float* v, int64_t rows) {
    const ModelView& m = h->m;
    if (rows <= 0) return SFM_OK;
    if (m.k == 0 || !v) {
        CU(cudaMemsetAsync(dst, 0, sizeof(float) * (size_t)rows * m.kp, h->stream));
        return SFM_OK;
    }
    const size_t bytes = sizeof(float) * (size_t)rows * m.k;
    if (m.k == m.kp) {
        CU(cudaMemcpyAsync(dst, v, bytes, cudaMemcpyHostToDevice, h->stream));
    } else {
        RC(ensure(h, h->b_grad, bytes));
        CU(cudaMemcpyAsync(h->b_grad.p, v, bytes, cudaMemcpyHostToDevice, h->stream));
        CU(launch_pad_v((const float*)h->b_grad.p, dst, rows, m.k, m.kp, false, h->stream,
                        &h->stats.kernel_launches));
    }
    h->stats.h2d_bytes += (int64_t)bytes;
    return SFM_OK;
}

static int upload_v(sfm_handle* h, const float* v) { return upload_v_rows(h, h->m.v, v, h->m.n_slots); }

static int download_v(sfm_handle* h, const float* dev_padded, float* v) {
    const ModelView& m = h->m;
    if (m.k == 0 || !v) return SFM_OK;
    const size_t bytes = sizeof(float) * (size_t)m.n_slots * m.k;
    if (m.k == m.kp) {
        CU(cudaMemcpyAsync(v, dev_padded, bytes, cudaMemcpyDeviceToHost, h->stream));
    } else {
        RC(ensure(h, h->b_yhat, bytes));
        CU(launch_pad_v(dev_padded, (float*)h->b_yhat.p, m.n_slots, m.k, m.kp, true, h->stream,
                        &h->stats.kernel_launches));
        CU(cudaMemcpyAsync(v, h->b_yhat.p, bytes, cudaMemcpyDeviceToHost, h->stream));
    }
    h->stats.d2h_bytes += (int64_t)bytes;
    return SFM_OK;
}

static bool is_sharded(const sfm_handle* h) { return h->shard != nullptr; }

#define NEED_MODEL(h)                                                                          \
    do {                                                                                       \
        if ((h)->shard_requested && !(h)->shard)                                               \
            return set_err((h), SFM_ERR_STATE, "SFM_FLAG_SHARD_V: call sfm_comm_init first");  \
    } while (0)

// Queues the copy of a host CSR batch into a staging slot on stream `st`.
static int stage_csr(sfm_handle* h, Stage& sg, cudaStream_t st, const int64_t* row_ptr,
                     const int32_t* idx, const float* val, const float* label, int64_t n_rows) {
    sg.valid = false;
    if (n_rows < 0 || (n_rows > 0 && !row_ptr)) return set_err(h, SFM_ERR_ARG, "bad CSR arguments");
    const int64_t nnz = n_rows > 0 ? row_ptr[n_rows] - row_ptr[0] : 0;
    if (nnz < 0 || nnz >= 2147483647LL) return set_err(h, SFM_ERR_ARG, "batch nnz out of range [0, 2^31-1)");
    if (nnz > 0 && !idx) return set_err(h, SFM_ERR_ARG, "idx is NULL");
    if (n_rows > 0 && row_ptr[0] != 0) return set_err(h, SFM_ERR_INDEX, "row_ptr[0] must be 0");
    RC(ensure(h, sg.rowptr, sizeof(int64_t) * (size_t)(n_rows + 1)));
    RC(ensure(h, sg.idx, sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1)));
    if (val) RC(ensure(h, sg.val, sizeof(float) * (size_t)(nnz > 0 ? nnz : 1)));
    if (label) RC(ensure(h, sg.label, sizeof(float) * (size_t)(n_rows > 0 ? n_rows : 1)));
    if (n_rows > 0) {
        CU(cudaMemcpyAsync(sg.rowptr.p, row_ptr, sizeof(int64_t) * (size_t)(n_rows + 1),
                           cudaMemcpyHostToDevice, st));
        h->stats.h2d_bytes += (int64_t)sizeof(int64_t) * (n_rows + 1);
        if (label) {
            CU(cudaMemcpyAsync(sg.label.p, label, sizeof(float) * (size_t)n_rows,
                               cudaMemcpyHostToDevice, st));
            h->stats.h2d_bytes += (int64_t)sizeof(float) * n_rows;
        }
    } else {
        CU(cudaMemsetAsync(sg.rowptr.p, 0, sizeof(int64_t), st));
    }
    if (nnz > 0) {
        CU(cudaMemcpyAsync(sg.idx.p, idx, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, st));
        h->stats.h2d_bytes += (int64_t)sizeof(int32_t) * nnz;
        if (val) {
            CU(cudaMemcpyAsync(sg.val.p, val, sizeof(float) * (size_t)nnz, cudaMemcpyHostToDevice, st));
            h->stats.h2d_bytes += (int64_t)sizeof(float) * nnz;
        }
    }
    // row_ptr monotonicity and bounds are checked by the forward kernel (error flag)
    sg.n_rows = n_rows;
    sg.nnz = nnz;
    sg.has_val = val != nullptr;
    sg.has_label = label != nullptr;
    sg.uniform_m = -1;
    sg.valid = true;
    return SFM_OK;
}

// Compact form (sfm_stage_onehot): copy the bit-packed ids and the labels, unpack on `st`.
static int stage_onehot(sfm_handle* h, Stage& sg, cudaStream_t st, const uint32_t* packed_idx,
                        const uint32_t* label_bits, const float* label_f32, int64_t n_rows, int m,
                        int id_bits) {
    sg.valid = false;
    if (n_rows < 0 || m < 1 || m > 64 || id_bits < 1 || id_bits > 32)
        return set_err(h, SFM_ERR_ARG, "sfm_stage_onehot: need n_rows >= 0, 1 <= m <= 64, 1 <= id_bits <= 32");
    if (n_rows > 0 && (!packed_idx || (label_bits == nullptr) == (label_f32 == nullptr)))
        return set_err(h, SFM_ERR_ARG, "sfm_stage_onehot: packed_idx and exactly one label array are required");
    const int64_t nnz = n_rows * m;
    if (nnz >= 2147483647LL) return set_err(h, SFM_ERR_ARG, "batch nnz out of range [0, 2^31-1)");
    const size_t words = (size_t)(((uint64_t)nnz * (uint64_t)id_bits + 31) / 32) + 1;
    const size_t lwords = (size_t)((n_rows + 31) / 32);
    RC(ensure(h, sg.packed, sizeof(uint32_t) * words));
    RC(ensure(h, sg.idx, sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1)));
    RC(ensure(h, sg.label, sizeof(float) * (size_t)(n_rows > 0 ? n_rows : 1)));
    if (label_bits) RC(ensure(h, sg.lbits, sizeof(uint32_t) * (lwords > 0 ? lwords : 1)));
    if (!sg.d_bad) CU(cudaMalloc(&sg.d_bad, sizeof(int32_t)));
    CU(cudaMemsetAsync(sg.d_bad, 0, sizeof(int32_t), st));
    if (n_rows > 0) {
        CU(cudaMemcpyAsync(sg.packed.p, packed_idx, sizeof(uint32_t) * words, cudaMemcpyHostToDevice, st));
        h->stats.h2d_bytes += (int64_t)(sizeof(uint32_t) * words);
        if (label_bits) {
            CU(cudaMemcpyAsync(sg.lbits.p, label_bits, sizeof(uint32_t) * lwords, cudaMemcpyHostToDevice, st));
            h->stats.h2d_bytes += (int64_t)(sizeof(uint32_t) * lwords);
        } else {
            CU(cudaMemcpyAsync(sg.label.p, label_f32, sizeof(float) * (size_t)n_rows, cudaMemcpyHostToDevice, st));
            h->stats.h2d_bytes += (int64_t)sizeof(float) * n_rows;
        }
        CU(launch_unpack_onehot((const uint32_t*)sg.packed.p,
                                label_bits ? (const uint32_t*)sg.lbits.p : nullptr, nnz, n_rows, id_bits,
                                h->m.n_slots, (int32_t*)sg.idx.p, (float*)sg.label.p, sg.d_bad, st,
                                &h->stats.kernel_launches));
    }
    sg.n_rows = n_rows;
    sg.nnz = nnz;
    sg.has_val = false;
    sg.has_label = true;
    sg.uniform_m = m;
    sg.valid = true;
    return SFM_OK;
}

static void stage_view(const Stage& sg, BatchView* out) {
    if (sg.uniform_m >= 0) {   // unpacked one-hot batch: uniform rows, ids already range-checked
        out->row_ptr = nullptr;
        out->idx = (const int32_t*)sg.idx.p;
        out->val = nullptr;
        out->label = (const float*)sg.label.p;
        out->row_ids = nullptr;
        out->row_lo = 0;
        out->n_rows = sg.n_rows;
        out->nnz = sg.nnz;
        out->idx_len = sg.nnz;
        out->out_ptr = nullptr;
        out->out_base = 0;
        out->uniform_m = sg.uniform_m;
        out->validated = true;
        out->pre_err = sg.d_bad;
        return;
    }
    out->pre_err = nullptr;
    out->row_ptr = (const int64_t*)sg.rowptr.p;
    out->idx = (const int32_t*)sg.idx.p;
    out->val = sg.has_val ? (const float*)sg.val.p : nullptr;
    out->label = sg.has_label ? (const float*)sg.label.p : nullptr;
    out->row_ids = nullptr;
    out->row_lo = 0;
    out->n_rows = sg.n_rows;
    out->nnz = sg.nnz;
    out->idx_len = sg.nnz;
    out->out_ptr = (const int64_t*)sg.rowptr.p;
    out->out_base = 0;
    out->uniform_m = -1;
    out->validated = false;
}

static int read_err_flag(sfm_handle* h) {
    CU(cudaMemcpyAsync(h->h_flags, h->d_err, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    h->h_flags[1] = 0;
    if (h->p2p)
        CU(cudaMemcpyAsync(h->h_flags + 1, p2p_timeout_flag(h), sizeof(int32_t),
                           cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (h->comm) {   // a dead peer / failed transport shows up here, not in the enqueue calls
        std::string msg;
        if (nccl_async_error(h->nccl, h->comm, &msg) != SFM_OK) {
            nccl_abort(h->nccl, h->comm);
            h->comm = nullptr;   // world stays > 1: train_core refuses to run without the communicator
            return set_err(h, SFM_ERR_NCCL, msg + " (communicator aborted)");
        }
    }
    if (h->h_flags[1])
        return set_err(h, SFM_ERR_NCCL, "peer-memory gradient exchange timed out waiting for a rank");
    if (h->h_flags[0])
        return set_err(h, SFM_ERR_INDEX, "feature index outside [0, n_slots) in the batch");
    return SFM_OK;
}

// Row-sharded models index a lookup table with the raw feature ids, so host-supplied batches are
// range-checked first (resident data sets were checked when they were loaded).
static int validate_view(sfm_handle* h, const BatchView& b) {
    if (b.validated || b.nnz <= 0) return SFM_OK;
    int32_t* mm = h->d_count + 1;
    const int32_t init[2] = {INT32_MAX, INT32_MIN};
    RC(ensure_pinned(h, 64));
    memcpy(h->h_pinned, init, sizeof init);
    CU(cudaMemcpyAsync(mm, h->h_pinned, sizeof init, cudaMemcpyHostToDevice, h->stream));
    CU(launch_idx_range(b.idx, b.nnz, mm, h->stream, &h->stats.kernel_launches));
    CU(cudaMemcpyAsync(h->h_flags + 4, mm, sizeof init, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (h->h_flags[4] < 0 || (int64_t)h->h_flags[5] >= h->m.n_slots)
        return set_err(h, SFM_ERR_INDEX, "feature index outside [0, n_slots) in the batch");
    return SFM_OK;
}

// The launch sequence of one SGD iteration on one rank.  Asynchronous: the caller
// synchronises.  If grad_keep, the (all-reduced) dense gradient stays in h->b_grad and no update
// is applied.
static int train_core(sfm_handle* h, const BatchView& b, int64_t iter, bool grad_keep,
                      const PartCache* pc = nullptr, const AheadPlan* ap = nullptr) {
    NEED_MODEL(h);
    if (h->world > 1 && !h->comm)
        return set_err(h, SFM_ERR_NCCL, "the communicator was aborted after an asynchronous NCCL error");
    if (is_sharded(h)) {
        if (grad_keep) return set_err(h, SFM_ERR_STATE, "sfm_gradient is not available with SFM_FLAG_SHARD_V");
        RC(validate_view(h, b));
        const int slot = pc ? (int)(pc - h->parts.data()) : -1;
        return shard_train(h, b, iter, slot);
    }
    ModelView& m = h->m;
    const int64_t n = b.n_rows, nnz = b.nnz;
    int64_t* L = &h->stats.kernel_launches;
    RC(ensure(h, h->b_S, sizeof(float) * (size_t)(n > 0 ? n : 1) * m.kp));
    RC(ensure(h, h->b_mult, sizeof(float) * (size_t)(n > 0 ? n : 1)));
    RC(ensure(h, h->b_loss, sizeof(float) * (size_t)(n > 0 ? n : 1)));
    const bool binary = b.val == nullptr;  // all-ones data: 4-byte payload (the row)
    if (!pc)
        for (int i = 0; i < 2; ++i) {
            RC(ensure(h, h->b_keys[i], sizeof(uint32_t) * (size_t)(nnz > 0 ? nnz : 1)));
            RC(ensure(h, h->b_pay[i], (binary ? sizeof(uint32_t) : sizeof(uint2)) * (size_t)(nnz > 0 ? nnz : 1)));
        }
    int blk_shift = 30, n_blocks = 1;
    if (pc) {
        blk_shift = pc->blk_shift;
        n_blocks = pc->n_blocks;
    } else {
        pull_plan(m, n, &blk_shift, &n_blocks);
    }
    const int key_bits = bits_for(m.n_slots);
    int blk_bits = 0;
    while (((int64_t)1 << blk_bits) < n_blocks) ++blk_bits;
    if (key_bits + blk_bits > 32) return set_err(h, SFM_ERR_ARG, "sort key does not fit 32 bits");
    // bucket form of the transposition + reduce (sfm_bucket.cu) whenever it applies; otherwise the
    // two-pass global sort + chunked reduce
    const bool multi = h->world > 1;
    const bool fused = !multi && !grad_keep;
    const bool p2p = multi && h->p2p && !grad_keep;   // sum + update in one kernel over NVLink
    BucketGeom bg;
    bool bucket = false;
    // Multi-GPU, reduce / all-reduce overlap (SFM_AR_SLICES): the feature range is
    // cut into Q slices; slice q's gradient is all-reduced and applied on the comm stream while the
    // reduce of slice q+1 runs on the compute stream.  The sliced path issues a different sequence
    // of collectives, so whether it is taken must come out the same on every rank: it is decided
    // from the knobs, the communicator state and the model geometry only -- never from this rank's
    // batch (under the Bernoulli sampler the ranks' n / nnz differ; an empty local batch walks the
    // sliced path with zero chunks).  It uses the fully sorted form of the reduce.
    int n_slices = 1;
    if (multi && !p2p && !grad_keep && !h->phase_timing && knobs().ar_slices > 1 &&
        knobs().pull_block_mb <= 0 && !knobs().bucket_cache && m.n_slots >= 64 * knobs().ar_slices)
        n_slices = knobs().ar_slices;
    const bool sliced = n_slices > 1;
    if (pc) {
        bucket = pc->bucket;
        bg = pc->geom;
    } else {
        bucket = !sliced && n_blocks == 1 && bucket_geometry(m, key_bits, n, nnz, &bg);
    }
    // the bucket counts + plan of this batch were built one step ahead (sfm_train): usable when the
    // step takes the bucket path and its work items fit the buffers the plan was carved with
    const bool planned = ap && !pc && bucket && n <= ap->cap_rows && bg.LB == ap->geom.LB &&
                         bg.HB == ap->geom.HB && bg.max_items <= ap->geom.max_items;
    if (planned) bg = ap->geom;
    RC(ensure(h, h->b_partials, sizeof(double) * 4 * 512));
    const int end_bit = key_bits + blk_bits;
    size_t sort_bytes = 0;
    if (bucket) {
        if (!planned) {
            RC(ensure(h, h->b_bkt_work, bucket_work_bytes(m, bg, h->sm_count)));
            if (!pc) RC(ensure(h, h->b_bkt_tables, bucket_tables_bytes(bg)));
        }
    } else {
        RC(ensure(h, h->b_seg, sizeof(int32_t) * 2 * (size_t)m.n_slots * n_blocks));  // seg_lo | seg_hi
        RC(ensure(h, h->b_pull, pull_scratch_bytes(m, nnz, n_blocks)));
        if (nnz > 0 && !pc) {
            sort_bytes = binary ? sort_pairs32_temp_bytes(nnz, end_bit) : sort_pairs_temp_bytes(nnz, end_bit);
            RC(ensure(h, h->b_sort_tmp, sort_bytes));
        }
    }
    if (!fused && !p2p) RC(ensure(h, h->b_grad, sizeof(float) * grad_len(h)));

    if ((n + 1) * (int64_t)m.lpr >= 4294967296LL || n >= (1LL << 30))
        return set_err(h, SFM_ERR_ARG, "batch too large: rows must stay below 2^30 and rows * kp/4 below 2^32");
    PhaseTimer pt(h);
    if (b.pre_err)   // the staging kernel already range-checked the ids: its flag is the step's
        CU(cudaMemcpyAsync(h->d_err, b.pre_err, sizeof(int32_t), cudaMemcpyDeviceToDevice, h->stream));
    else
        CU(cudaMemsetAsync(h->d_err, 0, sizeof(int32_t), h->stream));
    FwdOut o;
    o.S = (float*)h->b_S.p;
    o.mult = (float*)h->b_mult.p;
    o.loss = (float*)h->b_loss.p;
    o.yhat = nullptr;
    o.keys = pc ? nullptr : (uint32_t*)h->b_keys[0].p;   // cached transposition: nothing to emit
    o.pay = pc ? nullptr : (uint2*)h->b_pay[0].p;
    // all-ones rows of m entries in row order: entry i belongs to batch row i / m, so the forward
    // kernel writes no payload and the first sort pass derives it (saves 8 bytes per entry)
    const int implicit_div = (!pc && binary && b.uniform_m > 0 && b.uniform_m < 512 && !b.out_ptr &&
                              b.out_base == 0 && nnz > 0 && (bucket || radix_usable(nnz, end_bit)))
                                 ? b.uniform_m : 0;
    if (implicit_div) o.pay = nullptr;
    o.key_bits = key_bits;
    o.blk_shift = blk_shift;
    CU(launch_forward(m, b, o, true, h->d_err, h->sm_count, h->stream, L));
    CU(launch_scalar_reduce(o.loss, o.mult, n, (double*)h->b_partials.p,
                            reinterpret_cast<unsigned int*>(h->d_count + 3), h->d_scal, h->d_err,
                            h->stream, L));
    pt.lap(&h->stats.ms_forward);
    if (multi) {
        if (sliced) {
            CU(cudaEventRecord(h->ev_pool[0], h->stream));
            CU(cudaStreamWaitEvent(h->comm_stream, h->ev_pool[0], 0));
            RC(nccl_allreduce_f64(h->nccl, h->comm, h->d_scal, SC_N, h->comm_stream, &h->err));
            CU(cudaEventRecord(h->ev_pool[1], h->comm_stream));
        } else if (!p2p) {   // peer-memory path: the scalars travel with the gradient
            RC(nccl_allreduce_f64(h->nccl, h->comm, h->d_scal, SC_N, h->stream, &h->err));
            pt.lap(&h->stats.ms_allreduce);
        }
    }
    const uint32_t* keys_sorted = pc ? (const uint32_t*)pc->keys.p : (const uint32_t*)h->b_keys[1].p;
    const uint2* pay_sorted = pc ? (const uint2*)pc->pay.p : (const uint2*)h->b_pay[1].p;
    void* const bk_work = planned ? ap->work : h->b_bkt_work.p;
    const void* const bk_tables = planned ? ap->tables : (pc ? pc->tables.p : h->b_bkt_tables.p);
    if (nnz > 0 && !pc && bucket && planned) {
        CU(bucket_scatter(m, b, bg, o.keys, o.pay, implicit_div, bk_work, bk_tables,
                          (uint32_t*)h->b_keys[1].p, binary ? nullptr : (uint32_t*)h->b_pay[1].p,
                          h->sm_count, h->stream, L));
    } else if (nnz > 0 && !pc && bucket) {
        CU(bucket_transpose(m, b, bg, o.keys, o.pay, implicit_div, h->b_bkt_work.p,
                            h->b_bkt_tables.p, (uint32_t*)h->b_keys[1].p,
                            binary ? nullptr : (uint32_t*)h->b_pay[1].p,
                            reinterpret_cast<unsigned int*>(h->d_count + 4), h->sm_count, h->stream, L));
    } else if (nnz > 0 && !pc) {
        if (binary)
            CU(sort_pairs32(h->b_sort_tmp.p, sort_bytes, o.keys, (uint32_t*)h->b_keys[1].p,
                            (const uint32_t*)o.pay, (uint32_t*)h->b_pay[1].p, nnz, end_bit,
                            h->stream, L, implicit_div));
        else
            CU(sort_pairs(h->b_sort_tmp.p, sort_bytes, o.keys, (uint32_t*)h->b_keys[1].p, o.pay,
                          (uint2*)h->b_pay[1].p, nnz, end_bit, h->stream, L));
    }
    pt.lap(&h->stats.ms_sort);
    const UpdateParams up = update_params(h, iter);
    if (sliced) {
        // where each feature slice starts in the sorted entry list (cached with the transposition)
        int32_t pos[17];
        pos[0] = 0;
        pos[n_slices] = (int32_t)nnz;
        PartCache* pcm = const_cast<PartCache*>(pc);
        if (pcm && pcm->n_slices == n_slices) {
            for (int q = 1; q < n_slices; ++q) pos[q] = pcm->slice_pos[q];
        } else {
            CU(launch_slice_bounds(keys_sorted, nnz, n_slices, m.n_slots, h->d_slice, h->stream, L));
            CU(cudaMemcpyAsync(h->h_slice, h->d_slice, sizeof(int32_t) * 16, cudaMemcpyDeviceToHost,
                               h->stream));
            CU(cudaStreamSynchronize(h->stream));
            for (int q = 1; q < n_slices; ++q) pos[q] = h->h_slice[q];
            if (pcm) {
                pcm->n_slices = n_slices;
                for (int q = 1; q < n_slices; ++q) pcm->slice_pos[q] = pos[q];
            }
        }
        const int64_t chb = pull_chunk_entries(m);
        const int64_t nchunks = (nnz + chb - 1) / chb;
        // slice-major gradient buffer: slice q = [gV rows of the slice | gw of the slice], so one
        // all-reduce per slice; gw0 sits at the very end
        float* grad = (float*)h->b_grad.p;
        float* gw0 = grad + (size_t)m.n_slots * (m.kp + 1);
        CU(cudaStreamWaitEvent(h->stream, h->ev_pool[1], 0));   // all-reduced scalars
        for (int q = 0; q < n_slices; ++q) {
            PullSlice sl;
            sl.feat_lo = (int64_t)q * m.n_slots / n_slices;
            sl.feat_hi = (int64_t)(q + 1) * m.n_slots / n_slices;
            sl.chunk_lo = q == 0 ? 0 : (pos[q] + chb - 1) / chb;
            sl.chunk_hi = q == n_slices - 1 ? nchunks : (pos[q + 1] + chb - 1) / chb;
            sl.first = q == 0;
            const size_t nf = (size_t)(sl.feat_hi - sl.feat_lo);
            float* base = grad + (size_t)sl.feat_lo * (m.kp + 1);
            sl.gv = base - (size_t)sl.feat_lo * m.kp;          // kernels index with global feature ids
            sl.gw = base + nf * m.kp - sl.feat_lo;
            sl.gw0 = gw0;
            CU(launch_pull_slice(m, (int32_t*)h->b_seg.p, key_bits, n_blocks, keys_sorted, pay_sorted,
                                 nnz, binary, o.S, o.mult, (float*)h->b_pull.p, h->d_scal, h->d_err,
                                 up, false, grad, h->sm_count, h->stream, sl, L));
            CU(cudaEventRecord(h->ev_pool[2 + q], h->stream));
            CU(cudaStreamWaitEvent(h->comm_stream, h->ev_pool[2 + q], 0));
            RC(nccl_allreduce_f32(h->nccl, h->comm, base, nf * (m.kp + 1), h->comm_stream, &h->err));
            CU(launch_update_ptrs(m, sl.gv, sl.gw, sl.gw0, h->d_scal, h->d_err, up, sl.feat_lo,
                                  sl.feat_hi, h->comm_stream, L));
        }
        CU(cudaEventRecord(h->ev_pool[2 + n_slices], h->comm_stream));
        CU(cudaStreamWaitEvent(h->stream, h->ev_pool[2 + n_slices], 0));
        h->stats.train_steps += 1;
        h->stats.train_rows += n;
        h->stats.train_nnz += nnz;
        return SFM_OK;
    }
    float* grad_out = fused ? nullptr : (p2p ? p2p_grad_buffer(h) : (float*)h->b_grad.p);
    // Sparse peer-memory exchange: whether it is used depends on the model geometry only (the same
    // on every rank); a rank whose own batch cannot take the bucket path this step (empty batch,
    // cached sorted form) publishes a dense gradient with an all-ones touched bitmap instead.
    const bool sparse_x = p2p && knobs().p2p_sparse && bucket_sparse_capable(m, key_bits);
    uint32_t* touch = nullptr;
    if (sparse_x) {
        if (bucket)
            touch = p2p_touch_bits(h);
        else
            CU(cudaMemsetAsync(p2p_touch_bits(h), 0xff, sizeof(uint32_t) * (size_t)((m.n_slots + 31) / 32),
                               h->stream));
    }
    if (bucket)
        CU(bucket_pull(m, bg, keys_sorted, binary ? nullptr : (const uint32_t*)pay_sorted,
                       bk_tables, bk_work, o.S, o.mult, h->d_scal,
                       h->d_err, up, fused, grad_out, touch, nnz > 0 && !pc, h->sm_count, h->stream, L));
    else
        CU(launch_pull(m, (int32_t*)h->b_seg.p, key_bits, n_blocks, keys_sorted, pay_sorted, nnz,
                       binary, o.S, o.mult, (float*)h->b_pull.p, h->d_scal, h->d_err, up, fused,
                       grad_out, h->sm_count, h->stream, L));
    pt.lap(&h->stats.ms_reduce);
    if (p2p) {
        RC(p2p_reduce_update(h, up, sparse_x));
        pt.lap(&h->stats.ms_allreduce);
        h->stats.train_steps += 1;
        h->stats.train_rows += n;
        h->stats.train_nnz += nnz;
        return SFM_OK;
    }
    if (multi) {
        // [gV | gw] only: the trailing gw0 slot already holds the GLOBAL value (it comes from
        // the all-reduced scalar block above)
        RC(nccl_allreduce_f32(h->nccl, h->comm, (float*)h->b_grad.p, grad_len(h) - 1, h->stream,
                              &h->err));
        pt.lap(&h->stats.ms_allreduce);
    }
    if (!fused && !grad_keep) {
        CU(launch_update(m, (const float*)h->b_grad.p, h->d_scal, h->d_err, up, h->stream, L));
        pt.lap(&h->stats.ms_update);
    }
    h->stats.train_steps += grad_keep ? 0 : 1;
    h->stats.train_rows += n;
    h->stats.train_nnz += nnz;
    return SFM_OK;
}

// One SGD iteration as ONE graph launch: train_core's launch sequence is captured
// from the compute stream, the resident executable graph is updated in place with this step's
// arguments (cudaGraphExecUpdate: same topology, new batch size / step number / buffers) and
// launched -- the device runs the ~8 dependent kernels of a step without per-launch gaps, which is
// what small mini-batches are bound by.  Anything that cannot be captured (a scratch buffer has to
// grow, NCCL collectives, per-phase timing, the row-sharded model) runs as plain stream launches.
static int train_step_run(sfm_handle* h, const BatchView& b, int64_t iter, const PartCache* pc = nullptr,
                          const AheadPlan* ap = nullptr) {
    const bool multi = h->world > 1;
    if (!knobs().step_graph || h->phase_timing || is_sharded(h) || (multi && !h->p2p) ||
        (h->shard_requested && !h->shard))
        return train_core(h, b, iter, false, pc, ap);
    const sfm_stats saved = h->stats;
    if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        return train_core(h, b, iter, false, pc, ap);
    }
    h->capturing = true;
    h->capture_abort = false;
    int rc = train_core(h, b, iter, false, pc, ap);
    h->capturing = false;
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(h->stream, &g);
    if (rc != SFM_OK || ce != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        if (h->capture_abort || ce != cudaSuccess) {   // e.g. a scratch buffer must grow: run it uncaptured
            h->capture_abort = false;
            h->stats = saved;
            h->err.clear();
            return train_core(h, b, iter, false, pc, ap);
        }
        return rc;
    }
    bool ready = false;
    if (h->step_exec) {
        cudaGraphExecUpdateResultInfo info;
        if (cudaGraphExecUpdate(h->step_exec, g, &info) == cudaSuccess) {
            ready = true;
        } else {   // another launch sequence (other path / sampler / exchange): new executable
            cudaGetLastError();
            cudaGraphExecDestroy(h->step_exec);
            h->step_exec = nullptr;
        }
    }
    cudaError_t le = cudaSuccess;
    if (!ready) le = cudaGraphInstantiate(&h->step_exec, g, 0);
    if (le == cudaSuccess) le = cudaGraphLaunch(h->step_exec, h->stream);
    cudaGraphDestroy(g);
    if (le != cudaSuccess) {
        if (h->step_exec) cudaGraphExecDestroy(h->step_exec);
        h->step_exec = nullptr;
        return cuda_fail(h, le, "step graph");
    }
    return SFM_OK;
}

// Builds the view of a batch of resident rows.  ids_dev: device int32 row ids or nullptr (all).
static int resident_batch(sfm_handle* h, const int32_t* ids_dev, int64_t n_ids, BatchView* b) {
    const Dataset& ds = h->ds;
    b->row_ptr = ds.row_ptr;
    b->idx = ds.idx;
    b->val = ds.val;
    b->label = ds.label;
    b->row_ids = ids_dev;
    b->row_lo = 0;
    b->idx_len = ds.nnz;
    b->uniform_m = ds.uniform_m;
    b->validated = true;   // sfm_load_dataset / sfm_synth_ctr_dataset checked the index range
    b->out_base = 0;
    if (!ids_dev) {
        b->n_rows = ds.n_rows;
        b->nnz = ds.nnz;
        // identity batch: output slot = CSR position (= pos * m when the rows are uniform)
        b->out_ptr = ds.uniform_m >= 0 ? nullptr : ds.row_ptr;
        return SFM_OK;
    }
    b->n_rows = n_ids;
    if (ds.uniform_m >= 0) {
        b->nnz = n_ids * ds.uniform_m;
        b->out_ptr = nullptr;
        return SFM_OK;
    }
    // ragged rows: lengths -> exclusive scan -> output offsets; total read back
    RC(ensure(h, h->b_lens, sizeof(int64_t) * (size_t)(n_ids + 1)));
    RC(ensure(h, h->b_out_ptr, sizeof(int64_t) * (size_t)(n_ids + 1)));
    CU(cudaMemsetAsync((int64_t*)h->b_lens.p + n_ids, 0, sizeof(int64_t), h->stream));
    CU(launch_row_lens(ds.row_ptr, ids_dev, n_ids, (int64_t*)h->b_lens.p, h->stream,
                       &h->stats.kernel_launches));
    const size_t sb = scan_temp_bytes(n_ids + 1);
    RC(ensure(h, h->b_sel_tmp, sb));
    CU(exclusive_scan_i64(h->b_sel_tmp.p, sb, (const int64_t*)h->b_lens.p,
                          (int64_t*)h->b_out_ptr.p, n_ids + 1, h->stream,
                          &h->stats.kernel_launches));
    int64_t* total = (int64_t*)(h->h_flags + 2);
    CU(cudaMemcpyAsync(total, (int64_t*)h->b_out_ptr.p + n_ids, sizeof(int64_t),
                       cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    b->nnz = *total;
    b->out_ptr = (const int64_t*)h->b_out_ptr.p;
    b->uniform_m = -1;
    return SFM_OK;
}

// Built-in sampler on the device; returns the local batch size.
static int sample_device(sfm_handle* h, int64_t iter, const int32_t** ids_dev, int64_t* n_ids) {
    const Dataset& ds = h->ds;
    const double frac = (double)h->cfg.mini_batch_fraction;
    if (frac >= 1.0) {
        *ids_dev = nullptr;
        *n_ids = ds.n_rows;
        return SFM_OK;
    }
    RC(ensure(h, h->b_row_ids, sizeof(int32_t) * (size_t)(ds.n_rows > 0 ? ds.n_rows : 1)));
    if (ds.n_rows == 0 || !(frac > 0.0)) {
        *ids_dev = (const int32_t*)h->b_row_ids.p;
        *n_ids = 0;
        return SFM_OK;
    }
    const uint64_t thr = (uint64_t)floor(frac * 9007199254740992.0);
    const uint64_t key = mix64(h->cfg.sampler_seed + (uint64_t)iter);
    const size_t sb = select_temp_bytes(ds.n_rows);
    RC(ensure(h, h->b_sel_tmp, sb));
    CU(sample_rows_device(h->b_sel_tmp.p, sb, ds.n_rows, ds.global_offset, key, thr,
                          (int32_t*)h->b_row_ids.p, h->d_count, h->stream,
                          &h->stats.kernel_launches));
    CU(cudaMemcpyAsync(h->h_flags + 1, h->d_count, sizeof(int32_t), cudaMemcpyDeviceToHost,
                       h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->stats.d2h_bytes += 4;
    *ids_dev = (const int32_t*)h->b_row_ids.p;
    *n_ids = h->h_flags[1];
    return SFM_OK;
}

static void free_parts(sfm_handle* h) {
    shard_clear_cache(h);
    for (PartCache& pc : h->parts) {
        free_buf(pc.row_ids);
        free_buf(pc.keys);
        free_buf(pc.pay);
        free_buf(pc.tables);
    }
    h->parts.clear();
}

static int64_t n_parts_for(const sfm_handle* h) {
    const double f = (double)h->cfg.mini_batch_fraction;
    if (!(f < 1.0)) return 1;
    const int64_t p = (int64_t)floor(1.0 / f + 0.5);
    return p < 1 ? 1 : p;
}

// Fixed mini-batch of iteration `iter` under the PARTITION sampler (or the whole data set when
// mini_batch_fraction >= 1): builds its resident transposition at first use (select rows, emit
// the entry list, stable radix sort by feature) and returns the cached view.
static int partition_batch(sfm_handle* h, int64_t iter, BatchView* b, const PartCache** out) {
    const Dataset& ds = h->ds;
    const int64_t P = n_parts_for(h);
    if ((int64_t)h->parts.size() != P) {
        CU(cudaStreamSynchronize(h->stream));
        free_parts(h);
        h->parts.resize((size_t)P);
    }
    const int64_t part = (iter - 1) % P;
    PartCache& pc = h->parts[(size_t)part];
    int64_t* L = &h->stats.kernel_launches;
    if (!pc.built) {
        const int32_t* ids_dev = nullptr;
        int64_t n = ds.n_rows;
        if (P > 1) {
            RC(ensure(h, h->b_row_ids, sizeof(int32_t) * (size_t)(ds.n_rows > 0 ? ds.n_rows : 1)));
            n = 0;
            if (ds.n_rows > 0) {
                const size_t sb = select_temp_bytes(ds.n_rows);
                RC(ensure(h, h->b_sel_tmp, sb));
                CU(partition_rows_device(h->b_sel_tmp.p, sb, ds.n_rows, ds.global_offset,
                                         mix64(h->cfg.sampler_seed), P, part,
                                         (int32_t*)h->b_row_ids.p, h->d_count, h->stream, L));
                CU(cudaMemcpyAsync(h->h_flags + 1, h->d_count, sizeof(int32_t),
                                   cudaMemcpyDeviceToHost, h->stream));
                CU(cudaStreamSynchronize(h->stream));
                n = h->h_flags[1];
            }
            RC(ensure(h, pc.row_ids, sizeof(int32_t) * (size_t)(n > 0 ? n : 1)));
            if (n > 0)
                CU(cudaMemcpyAsync(pc.row_ids.p, h->b_row_ids.p, sizeof(int32_t) * (size_t)n,
                                   cudaMemcpyDeviceToDevice, h->stream));
            ids_dev = (const int32_t*)pc.row_ids.p;
        }
        BatchView v;
        RC(resident_batch(h, ids_dev, n, &v));   // ragged rows: output offsets via scan
        if (v.nnz >= 2147483647LL) return set_err(h, SFM_ERR_ARG, "batch nnz must be < 2^31-1");
        pull_plan(h->m, n, &pc.blk_shift, &pc.n_blocks);
        pc.key_bits = bits_for(h->m.n_slots);
        int blk_bits = 0;
        while (((int64_t)1 << blk_bits) < pc.n_blocks) ++blk_bits;
        if (pc.key_bits + blk_bits > 32) return set_err(h, SFM_ERR_ARG, "sort key does not fit 32 bits");
        const bool binary = v.val == nullptr;
        const size_t pay_sz = binary ? sizeof(uint32_t) : sizeof(uint2);
        const size_t cnt = (size_t)(v.nnz > 0 ? v.nnz : 1);
        RC(ensure(h, h->b_keys[0], sizeof(uint32_t) * cnt));
        RC(ensure(h, h->b_pay[0], pay_sz * cnt));
        if (!is_sharded(h)) {
            RC(ensure(h, pc.keys, sizeof(uint32_t) * cnt));
            RC(ensure(h, pc.pay, pay_sz * cnt));
        }
        // a cached batch is sorted ONCE, so the fully sorted form + the chunked reduce (no ranking
        // in the steady state) is the faster steady state; SFM_BUCKET_CACHE=1 caches the bucket form
        pc.bucket = knobs().bucket_cache && !is_sharded(h) && pc.n_blocks == 1 &&
                    bucket_geometry(h->m, pc.key_bits, n, v.nnz, &pc.geom);
        if (pc.bucket) {   // entries grouped by bucket once; the reduce ranks them per tile
            CU(launch_emit(v, pc.key_bits, pc.blk_shift, h->m.n_slots, (uint32_t*)h->b_keys[0].p,
                           (uint2*)h->b_pay[0].p, h->sm_count, h->stream, L));
            RC(ensure(h, h->b_bkt_work, bucket_work_bytes(h->m, pc.geom, h->sm_count)));
            RC(ensure(h, pc.tables, bucket_tables_bytes(pc.geom)));
            CU(bucket_transpose(h->m, v, pc.geom, (const uint32_t*)h->b_keys[0].p,
                                (const uint2*)h->b_pay[0].p, 0, h->b_bkt_work.p, pc.tables.p,
                                (uint32_t*)pc.keys.p, binary ? nullptr : (uint32_t*)pc.pay.p,
                                reinterpret_cast<unsigned int*>(h->d_count + 4), h->sm_count, h->stream, L));
        } else if (v.nnz > 0 && !is_sharded(h)) {   // row-sharded models cache their own plan (sfm_shard.cu)
            CU(launch_emit(v, pc.key_bits, pc.blk_shift, h->m.n_slots, (uint32_t*)h->b_keys[0].p,
                           (uint2*)h->b_pay[0].p, h->sm_count, h->stream, L));
            const int end_bit = pc.key_bits + blk_bits;
            const size_t sb = binary ? sort_pairs32_temp_bytes(v.nnz, end_bit)
                                     : sort_pairs_temp_bytes(v.nnz, end_bit);
            RC(ensure(h, h->b_sort_tmp, sb));
            if (binary)
                CU(sort_pairs32(h->b_sort_tmp.p, sb, (const uint32_t*)h->b_keys[0].p,
                                (uint32_t*)pc.keys.p, (const uint32_t*)h->b_pay[0].p,
                                (uint32_t*)pc.pay.p, v.nnz, end_bit, h->stream, L));
            else
                CU(sort_pairs(h->b_sort_tmp.p, sb, (const uint32_t*)h->b_keys[0].p,
                              (uint32_t*)pc.keys.p, (const uint2*)h->b_pay[0].p, (uint2*)pc.pay.p,
                              v.nnz, end_bit, h->stream, L));
        }
        pc.n_rows = n;
        pc.nnz = v.nnz;
        pc.built = true;
    }
    // steady state: no scan, no sort -- the forward only needs the row list
    b->row_ptr = ds.row_ptr;
    b->idx = ds.idx;
    b->val = ds.val;
    b->label = ds.label;
    b->row_ids = P > 1 ? (const int32_t*)pc.row_ids.p : nullptr;
    b->row_lo = 0;
    b->n_rows = pc.n_rows;
    b->nnz = pc.nnz;
    b->idx_len = ds.nnz;
    b->out_ptr = (P > 1 || ds.uniform_m >= 0) ? nullptr : ds.row_ptr;   // identity batch, ragged rows
    b->out_base = 0;
    b->uniform_m = ds.uniform_m;
    b->validated = true;
    *out = &pc;
    return SFM_OK;
}

// View of the resident rows [lo, hi) whose output offsets start at 0 (needed when the entry list
// of a sub-range is materialised: row-sharded predict / evaluate).
static int subrange_view(sfm_handle* h, int64_t lo, int64_t hi, BatchView* b) {
    const Dataset& ds = h->ds;
    b->row_lo = lo;
    b->n_rows = hi - lo;
    if (ds.uniform_m >= 0) {
        b->nnz = (hi - lo) * ds.uniform_m;
        b->out_ptr = nullptr;
        b->out_base = 0;
        return SFM_OK;
    }
    int64_t* ends = (int64_t*)(h->h_flags + 8);  // 2 x int64
    CU(cudaMemcpyAsync(ends, ds.row_ptr + lo, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(ends + 1, ds.row_ptr + hi, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    b->nnz = ends[1] - ends[0];
    b->out_ptr = ds.row_ptr + lo;
    b->out_base = ends[0];
    return SFM_OK;
}

static bool use_partitions(const sfm_handle* h) {
    return h->cfg.sampler_mode == SFM_SAMPLER_PARTITION || !((double)h->cfg.mini_batch_fraction < 1.0);
}

static int finish_step(sfm_handle* h, double* mean_loss_out, int64_t* batch_out) {
    CU(cudaMemcpyAsync(h->h_scal, h->d_scal, sizeof(double) * SC_N, cudaMemcpyDeviceToHost,
                       h->stream));
    h->stats.d2h_bytes += (int64_t)sizeof(double) * SC_N + 4;
    RC(read_err_flag(h));
    if (h->world > 1 && !is_sharded(h) && h->h_scal[SC_ERR] != 0.0)
        return set_err(h, SFM_ERR_INDEX,
                       "another rank saw a feature index outside [0, n_slots): every rank skipped this update");
    const double cnt = h->h_scal[SC_COUNT];
    if (mean_loss_out) *mean_loss_out = cnt > 0.0 ? h->h_scal[SC_LOSS] / cnt : 0.0;
    if (batch_out) *batch_out = (int64_t)cnt;
    return SFM_OK;
}

static int upload_row_ids(sfm_handle* h, const int64_t* row_ids, int64_t n_ids,
                          const int32_t** ids_dev) {
    RC(ensure_pinned(h, sizeof(int32_t) * (size_t)(n_ids > 0 ? n_ids : 1)));
    int32_t* st = (int32_t*)h->h_pinned;
    for (int64_t i = 0; i < n_ids; ++i) {
        if (row_ids[i] < 0 || row_ids[i] >= h->ds.n_rows)
            return set_err(h, SFM_ERR_ARG, "row id outside the resident data set");
        st[i] = (int32_t)row_ids[i];
    }
    RC(ensure(h, h->b_row_ids, sizeof(int32_t) * (size_t)(n_ids > 0 ? n_ids : 1)));
    if (n_ids > 0) {
        CU(cudaMemcpyAsync(h->b_row_ids.p, st, sizeof(int32_t) * (size_t)n_ids,
                           cudaMemcpyHostToDevice, h->stream));
        h->stats.h2d_bytes += (int64_t)sizeof(int32_t) * n_ids;
    }
    *ids_dev = (const int32_t*)h->b_row_ids.p;
    return SFM_OK;
}

}  // namespace sfm

// =============================================================================== exports ===
extern "C" {

int32_t sfm_abi_version(void) { return SFM_ABI_VERSION; }

const char* sfm_status_string(int32_t s) {
    switch (s) {
        case SFM_OK: return "ok";
        case SFM_ERR_ARG: return "bad argument";
        case SFM_ERR_CUDA: return "CUDA error";
        case SFM_ERR_NCCL: return "NCCL error";
        case SFM_ERR_OOM: return "out of memory";
        case SFM_ERR_INDEX: return "feature index out of range or malformed CSR";
        case SFM_ERR_IO: return "I/O or parse error";
        case SFM_ERR_STATE: return "invalid state for this call";
    }
    return "unknown status";
}

int32_t sfm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int32_t sfm_host_alloc(void** ptr, uint64_t bytes) {
    if (!ptr) return SFM_ERR_ARG;
    *ptr = nullptr;
    if (bytes == 0) return SFM_OK;
    return cudaMallocHost(ptr, bytes) == cudaSuccess ? SFM_OK : SFM_ERR_OOM;
}

int32_t sfm_host_free(void* ptr) {
    if (!ptr) return SFM_OK;
    return cudaFreeHost(ptr) == cudaSuccess ? SFM_OK : SFM_ERR_CUDA;
}

int32_t sfm_create(const sfm_config* cfg, sfm_handle** out) {
    if (!cfg || !out) return SFM_ERR_ARG;
    *out = nullptr;
    knobs_refresh();   // experiment knobs are (re)read here, never in a launch path
    if (cfg->abi_version != SFM_ABI_VERSION) return SFM_ERR_ARG;
    if (cfg->k < 0 || cfg->k > 128 || cfg->n_slots < 1 || cfg->n_slots >= 2147483647LL)
        return SFM_ERR_ARG;
    if ((cfg->n_slots + 1) * (int64_t)(kp_for(cfg->k) / 4) >= 4294967296LL)
        return SFM_ERR_ARG;  // V row offsets are 32-bit float4 indices in the kernels
    if (cfg->task != SFM_TASK_REGRESSION && cfg->task != SFM_TASK_CLASSIFICATION) return SFM_ERR_ARG;
    const int32_t sampler = cfg->sampler_mode & 0xFF;
    if ((sampler != SFM_SAMPLER_BERNOULLI && sampler != SFM_SAMPLER_PARTITION) ||
        (cfg->sampler_mode & ~(0xFF | SFM_FLAG_SHARD_V)))
        return SFM_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return SFM_ERR_CUDA;  // no CPU fallback, by design
    }
    if (cfg->device < 0 || cfg->device >= ndev) return SFM_ERR_ARG;
    sfm_handle* h = new (std::nothrow) sfm_handle;
    if (!h) return SFM_ERR_OOM;
    h->cfg = *cfg;
    h->cfg.sampler_mode = sampler;
    h->shard_requested = (cfg->sampler_mode & SFM_FLAG_SHARD_V) != 0;
    h->device = cfg->device;
    ModelView& m = h->m;
    m.n_slots = cfg->n_slots;
    m.k = cfg->k;
    m.kp = kp_for(cfg->k);
    m.lpr = m.kp / 4;
    m.k0 = cfg->k0 ? 1 : 0;
    m.k1 = cfg->k1 ? 1 : 0;
    m.task = cfg->task;
    m.v = m.w = m.w0 = nullptr;
#define CK(call)                      \
    if ((call) != cudaSuccess) {      \
        cudaGetLastError();           \
        sfm_destroy(h);               \
        return SFM_ERR_CUDA;          \
    }
    CK(cudaSetDevice(h->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, h->device));
    h->sm_count = prop.multiProcessorCount;
    if (knobs().stream_prio) {
        // compute stream at the greatest priority, copy stream (sampler one iteration ahead,
        // staging copies + unpack) at the least: the block scheduler hands SM slots to the
        // sampler's CTAs only where the step's kernels leave them free (kernel tails, beside the
        // persistent latency-bound scatter / reduce) instead of beside the issue-bound forward
        int least = 0, greatest = 0;
        CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CK(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, greatest));
        CK(cudaStreamCreateWithPriority(&h->copy_stream, cudaStreamNonBlocking, least));
    } else {
        CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    }
    CK(cudaEventCreate(&h->ev_a));
    CK(cudaEventCreate(&h->ev_b));
    CK(cudaEventCreate(&h->ev_t0));
    CK(cudaEventCreate(&h->ev_t1));
    CK(cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
    for (Stage& sg : h->stage) CK(cudaEventCreateWithFlags(&sg.ready, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        CK(cudaEventCreateWithFlags(&h->ev_samp[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_used[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_plan[i], cudaEventDisableTiming));
    }
    CK(cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
    for (cudaEvent_t& e : h->ev_pool) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CK(cudaMalloc(&h->d_slice, sizeof(int32_t) * 16));
    CK(cudaMallocHost(&h->h_slice, sizeof(int32_t) * 16));
    CK(cudaMalloc(&h->d_count2, sizeof(int32_t) * 2));
    CK(cudaMallocHost(&h->h_count2, sizeof(int32_t) * 2));
    // one extra, always-zero row at index n_slots: the target of padded / rejected entries
    // (row-sharded models allocate their shard in sfm_comm_init instead)
    if (!h->shard_requested) {
        CK(cudaMalloc(&m.v, sizeof(float) * (size_t)(m.n_slots + 1) * m.kp));
        CK(cudaMalloc(&m.w, sizeof(float) * (size_t)(m.n_slots + 1)));
    }
    CK(cudaMalloc(&m.w0, sizeof(float) * 4));
    CK(cudaMalloc(&h->d_scal, sizeof(double) * 8));
    CK(cudaMalloc(&h->d_err, sizeof(int32_t) * 4));
    CK(cudaMalloc(&h->d_count, sizeof(int32_t) * 8));
    CK(cudaMallocHost(&h->h_scal, sizeof(double) * 8));
    CK(cudaMallocHost(&h->h_flags, sizeof(int32_t) * 16));
    if (m.v) CK(cudaMemsetAsync(m.v, 0, sizeof(float) * (size_t)(m.n_slots + 1) * m.kp, h->stream));
    if (m.w) CK(cudaMemsetAsync(m.w, 0, sizeof(float) * (size_t)(m.n_slots + 1), h->stream));
    CK(cudaMemsetAsync(m.w0, 0, sizeof(float) * 4, h->stream));
    CK(cudaMemsetAsync(h->d_count, 0, sizeof(int32_t) * 8, h->stream));   // [3], [4]: tickets (scalar reduce, bucket plan)
    CK(cudaMemsetAsync(h->d_scal, 0, sizeof(double) * 8, h->stream));
    CK(cudaMemsetAsync(h->d_err, 0, sizeof(int32_t) * 4, h->stream));
    CK(cudaStreamSynchronize(h->stream));
#undef CK
    *out = h;
    return SFM_OK;
}

int32_t sfm_destroy(sfm_handle* h) {
    if (!h) return SFM_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->step_exec) cudaGraphExecDestroy(h->step_exec);
    h->step_exec = nullptr;
    p2p_teardown(h);
    if (h->comm) nccl_destroy(h->nccl, h->comm);
    sfm_unload_dataset(h);
    Buf* bufs[] = {&h->b_row_ids, &h->b_out_ptr, &h->b_S, &h->b_mult, &h->b_loss, &h->b_yhat,
                   &h->b_keys[0], &h->b_keys[1], &h->b_pay[0], &h->b_pay[1], &h->b_seg,
                   &h->b_sort_tmp, &h->b_grad, &h->b_partials, &h->b_sel_tmp, &h->b_lens,
                   &h->b_pull, &h->b_bkt_work, &h->b_bkt_tables};
    for (Buf* b : bufs) free_buf(*b);
    if (h->shard) {
        ShardState& ss = *h->shard;
        shard_clear_cache(h);
        Buf* sb[] = {&ss.flags, &ss.keys_tmp, &ss.small, &ss.lut, &ss.out_v, &ss.out_w, &ss.t_v, &ss.t_w,
                     &ss.gr_v, &ss.gr_w, &ss.acc, &ss.scratch.crank, &ss.scratch.pay, &ss.scratch.uniq,
                     &ss.scratch.req, &ss.scratch.bidx, &ss.scratch.bval, &ss.scratch.blabel,
                     &ss.scratch.optr};
        for (Buf* b : sb) free_buf(*b);
        if (ss.v) cudaFree(ss.v);
        if (ss.w) cudaFree(ss.w);
        if (ss.h_small) cudaFreeHost(ss.h_small);
        delete h->shard;
        h->shard = nullptr;
    }
    for (PartCache& pc : h->pre) {
        free_buf(pc.row_ids);
        free_buf(pc.keys);
        free_buf(pc.pay);
    }
    free_buf(h->b_pre_keys0);
    free_buf(h->b_pre_pay0);
    free_buf(h->b_sort_tmp2);
    free_buf(h->b_ids2[0]);
    free_buf(h->b_ids2[1]);
    free_buf(h->b_samp_tmp);
    if (h->comm_stream) {
        cudaStreamSynchronize(h->comm_stream);
        cudaStreamDestroy(h->comm_stream);
    }
    for (cudaEvent_t e : h->ev_pool)
        if (e) cudaEventDestroy(e);
    if (h->d_slice) cudaFree(h->d_slice);
    if (h->h_slice) cudaFreeHost(h->h_slice);
    if (h->d_count2) cudaFree(h->d_count2);
    if (h->h_count2) cudaFreeHost(h->h_count2);
    for (int i = 0; i < 2; ++i) {
        if (h->ev_samp[i]) cudaEventDestroy(h->ev_samp[i]);
        if (h->ev_used[i]) cudaEventDestroy(h->ev_used[i]);
        if (h->ev_plan[i]) cudaEventDestroy(h->ev_plan[i]);
        free_buf(h->b_bkt_work2[i]);
        free_buf(h->b_bkt_tables2[i]);
    }
    for (Stage& sg : h->stage) {
        free_buf(sg.rowptr);
        free_buf(sg.idx);
        free_buf(sg.val);
        free_buf(sg.label);
        free_buf(sg.packed);
        free_buf(sg.lbits);
        if (sg.d_bad) cudaFree(sg.d_bad);
        if (sg.ready) cudaEventDestroy(sg.ready);
    }
    if (h->m.v) cudaFree(h->m.v);
    if (h->m.w) cudaFree(h->m.w);
    if (h->m.w0) cudaFree(h->m.w0);
    if (h->d_scal) cudaFree(h->d_scal);
    if (h->d_err) cudaFree(h->d_err);
    if (h->d_count) cudaFree(h->d_count);
    if (h->h_scal) cudaFreeHost(h->h_scal);
    if (h->h_flags) cudaFreeHost(h->h_flags);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    cudaEvent_t evs[] = {h->ev_a, h->ev_b, h->ev_t0, h->ev_t1, h->ev_copy};
    for (cudaEvent_t e : evs)
        if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    delete h;
    return SFM_OK;
}

const char* sfm_last_error(const sfm_handle* h) { return h ? h->err.c_str() : "null handle"; }

int32_t sfm_get_config(const sfm_handle* h, sfm_config* out) {
    if (!h || !out) return SFM_ERR_ARG;
    *out = h->cfg;
    if (h->shard_requested) out->sampler_mode |= SFM_FLAG_SHARD_V;
    return SFM_OK;
}

int32_t sfm_set_hyper(sfm_handle* h, float reg0, float regw, float regv, float step_size,
                      float mini_batch_fraction) {
    if (!h) return SFM_ERR_ARG;
    h->cfg.reg0 = reg0;
    h->cfg.regw = regw;
    h->cfg.regv = regv;
    h->cfg.step_size = step_size;
    h->cfg.mini_batch_fraction = mini_batch_fraction;
    return SFM_OK;
}

// ------------------------------------------------------------------------------ model ----
int32_t sfm_set_model(sfm_handle* h, float w0, const float* w, const float* v) {
    if (!h) return SFM_ERR_ARG;
    NEED_MODEL(h);
    CU(cudaSetDevice(h->device));
    ModelView& m = h->m;
    if (m.k > 0 && !v) return set_err(h, SFM_ERR_ARG, "v is NULL");
    RC(ensure_pinned(h, 64));
    float* st = (float*)h->h_pinned;
    st[0] = w0;
    CU(cudaMemcpyAsync(m.w0, st, sizeof(float), cudaMemcpyHostToDevice, h->stream));
    if (is_sharded(h)) {  // every rank is handed the full arrays and keeps its own rows
        ShardState& ss = *h->shard;
        if (ss.n_own > 0) {
            if (w)
                CU(cudaMemcpyAsync(ss.w, w + ss.own_lo, sizeof(float) * (size_t)ss.n_own,
                                   cudaMemcpyHostToDevice, h->stream));
            else
                CU(cudaMemsetAsync(ss.w, 0, sizeof(float) * (size_t)ss.n_own, h->stream));
            RC(upload_v_rows(h, ss.v, v ? v + (size_t)ss.own_lo * m.k : nullptr, ss.n_own));
        }
        CU(cudaStreamSynchronize(h->stream));
        return SFM_OK;
    }
    if (w) {
        CU(cudaMemcpyAsync(m.w, w, sizeof(float) * (size_t)m.n_slots, cudaMemcpyHostToDevice,
                           h->stream));
        h->stats.h2d_bytes += (int64_t)sizeof(float) * m.n_slots;
    } else {
        CU(cudaMemsetAsync(m.w, 0, sizeof(float) * (size_t)m.n_slots, h->stream));
    }
    RC(upload_v(h, v));
    CU(cudaStreamSynchronize(h->stream));
    return SFM_OK;
}

int32_t sfm_get_model(sfm_handle* h, float* w0, float* w, float* v) {
    if (!h) return SFM_ERR_ARG;
    NEED_MODEL(h);
    CU(cudaSetDevice(h->device));
    ModelView& m = h->m;
    RC(ensure_pinned(h, 64));
    float* st = (float*)h->h_pinned;
    CU(cudaMemcpyAsync(st, m.w0, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (is_sharded(h)) {  // collective: all-gather the shards, then copy out like a replica
        ShardState& ss = *h->shard;
        const size_t rows = (size_t)ss.n_per * h->world;
        RC(ensure(h, h->b_pull, sizeof(float) * rows * (m.kp + 1)));
        float* gv = (float*)h->b_pull.p;
        float* gw = gv + rows * m.kp;
        RC(nccl_allgather_f32(h->nccl, h->comm, ss.v, gv, (size_t)ss.n_per * m.kp, h->stream, &h->err));
        RC(nccl_allgather_f32(h->nccl, h->comm, ss.w, gw, (size_t)ss.n_per, h->stream, &h->err));
        if (w)
            CU(cudaMemcpyAsync(w, gw, sizeof(float) * (size_t)m.n_slots, cudaMemcpyDeviceToHost,
                               h->stream));
        RC(download_v(h, gv, v));
        CU(cudaStreamSynchronize(h->stream));
        if (w0) *w0 = st[0];
        return SFM_OK;
    }
    if (w) {
        CU(cudaMemcpyAsync(w, m.w, sizeof(float) * (size_t)m.n_slots, cudaMemcpyDeviceToHost,
                           h->stream));
        h->stats.d2h_bytes += (int64_t)sizeof(float) * m.n_slots;
    }
    RC(download_v(h, m.v, v));
    CU(cudaStreamSynchronize(h->stream));
    if (w0) *w0 = st[0];
    return SFM_OK;
}

int32_t sfm_set_model_f64(sfm_handle* h, double w0, const double* w, const double* v) {
    if (!h) return SFM_ERR_ARG;
    const ModelView& m = h->m;
    std::vector<float> wf, vf;
    try {
        if (w) {
            wf.resize((size_t)m.n_slots);
            for (int64_t i = 0; i < m.n_slots; ++i) wf[(size_t)i] = (float)w[i];
        }
        if (v && m.k > 0) {
            vf.resize((size_t)m.n_slots * m.k);
            for (size_t i = 0; i < vf.size(); ++i) vf[i] = (float)v[i];
        }
    } catch (const std::bad_alloc&) {
        return set_err(h, SFM_ERR_OOM, "host allocation failed");
    }
    return sfm_set_model(h, (float)w0, w ? wf.data() : nullptr, vf.empty() ? nullptr : vf.data());
}

int32_t sfm_get_model_f64(sfm_handle* h, double* w0, double* w, double* v) {
    if (!h) return SFM_ERR_ARG;
    const ModelView& m = h->m;
    std::vector<float> wf, vf;
    try {
        if (w) wf.resize((size_t)m.n_slots);
        if (v && m.k > 0) vf.resize((size_t)m.n_slots * m.k);
    } catch (const std::bad_alloc&) {
        return set_err(h, SFM_ERR_OOM, "host allocation failed");
    }
    float w0f = 0.f;
    RC(sfm_get_model(h, &w0f, w ? wf.data() : nullptr, vf.empty() ? nullptr : vf.data()));
    if (w0) *w0 = w0f;
    for (size_t i = 0; i < wf.size(); ++i) w[i] = wf[i];
    for (size_t i = 0; i < vf.size(); ++i) v[i] = vf[i];
    return SFM_OK;
}

int32_t sfm_init_model(sfm_handle* h, double mean, double stdev, uint64_t seed) {
    if (!h) return SFM_ERR_ARG;
    NEED_MODEL(h);
    const ModelView& m = h->m;
    if (is_sharded(h)) {  // each rank draws only the elements of its own rows (same stream of numbers)
        ShardState& ss = *h->shard;
        CU(cudaSetDevice(h->device));
        std::vector<float> vf;
        try {
            vf.resize((size_t)ss.n_own * (size_t)m.k);
        } catch (const std::bad_alloc&) {
            return set_err(h, SFM_ERR_OOM, "host allocation failed");
        }
        if (!vf.empty())
            init_gaussian_f32(vf.data(), ss.own_lo * m.k, (int64_t)vf.size(), mean, stdev, seed);
        CU(cudaMemsetAsync(m.w0, 0, sizeof(float), h->stream));
        if (ss.n_own > 0) {
            CU(cudaMemsetAsync(ss.w, 0, sizeof(float) * (size_t)ss.n_own, h->stream));
            RC(upload_v_rows(h, ss.v, vf.empty() ? nullptr : vf.data(), ss.n_own));
        }
        CU(cudaStreamSynchronize(h->stream));
        return SFM_OK;
    }
    std::vector<float> vf;
    try {
        vf.resize((size_t)m.n_slots * (size_t)m.k);
    } catch (const std::bad_alloc&) {
        return set_err(h, SFM_ERR_OOM, "host allocation failed");
    }
    if (!vf.empty()) init_gaussian_f32(vf.data(), 0, (int64_t)vf.size(), mean, stdev, seed);
    return sfm_set_model(h, 0.f, nullptr, vf.empty() ? nullptr : vf.data());
}

struct SfmFileHeader {
    char magic[8];
    uint32_t version;
    int32_t task, k, k0, k1;
    int64_t n_slots;
    float reg0, regw, regv, step_size, mini_batch_fraction;
    uint32_t pad;
    uint64_t sampler_seed;
};

int32_t sfm_save(sfm_handle* h, const char* path) {
    if (!h || !path) return SFM_ERR_ARG;
    const ModelView& m = h->m;
    std::vector<float> wf, vf;
    try {   // no exception may cross the C ABI (the host is a JVM / Python process)
        wf.resize((size_t)m.n_slots);
        vf.resize((size_t)m.n_slots * m.k);
    } catch (const std::exception&) {
        return set_err(h, SFM_ERR_OOM, "sfm_save: host buffer for the model");
    }
    float w0 = 0.f;
    RC(sfm_get_model(h, &w0, wf.data(), vf.empty() ? nullptr : vf.data()));
    SfmFileHeader hd;
    memset(&hd, 0, sizeof hd);
    memcpy(hd.magic, "SFMB200", 8);
    hd.version = 1;
    hd.task = h->cfg.task; hd.k = m.k; hd.k0 = m.k0; hd.k1 = m.k1; hd.n_slots = m.n_slots;
    hd.reg0 = h->cfg.reg0; hd.regw = h->cfg.regw; hd.regv = h->cfg.regv;
    hd.step_size = h->cfg.step_size; hd.mini_batch_fraction = h->cfg.mini_batch_fraction;
    hd.sampler_seed = h->cfg.sampler_seed;
    hd.pad = (uint32_t)(h->cfg.sampler_mode & ~SFM_FLAG_SHARD_V);   // sharding is a property of the run, not of the model
    FILE* f = fopen(path, "wb");
    if (!f) return set_err(h, SFM_ERR_IO, std::string("cannot open for writing: ") + path);
    bool ok = fwrite(&hd, sizeof hd, 1, f) == 1 && fwrite(&w0, sizeof w0, 1, f) == 1 &&
              fwrite(wf.data(), sizeof(float), wf.size(), f) == wf.size() &&
              fwrite(vf.data(), sizeof(float), vf.size(), f) == vf.size();
    ok = (fclose(f) == 0) && ok;
    return ok ? SFM_OK : set_err(h, SFM_ERR_IO, std::string("short write: ") + path);
}

int32_t sfm_load(const char* path, int32_t device, sfm_handle** out) {
    if (!path || !out) return SFM_ERR_ARG;
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return SFM_ERR_IO;
    SfmFileHeader hd;
    // the header is untrusted input: the same bounds sfm_create applies, and the payload size it
    // implies must be the size of the rest of the file, before anything is allocated
    bool good = fread(&hd, sizeof hd, 1, f) == 1 && memcmp(hd.magic, "SFMB200", 8) == 0 &&
                hd.version == 1 && hd.n_slots >= 1 && hd.n_slots < 2147483647LL && hd.k >= 0 &&
                hd.k <= 128 && (hd.task == SFM_TASK_REGRESSION || hd.task == SFM_TASK_CLASSIFICATION) &&
                (hd.k0 == 0 || hd.k0 == 1) && (hd.k1 == 0 || hd.k1 == 1);
    if (good) {
        const long pos = ftell(f);
        good = pos >= 0 && fseek(f, 0, SEEK_END) == 0;
        const long end = good ? ftell(f) : -1;
        const unsigned long long want = 4ull * (1ull + (unsigned long long)hd.n_slots * (1ull + (unsigned long long)hd.k));
        good = good && end >= pos && (unsigned long long)(end - pos) == want && fseek(f, pos, SEEK_SET) == 0;
    }
    if (!good) {
        fclose(f);
        return SFM_ERR_IO;
    }
    std::vector<float> wf, vf;
    try {
        wf.resize((size_t)hd.n_slots);
        vf.resize((size_t)hd.n_slots * hd.k);
    } catch (const std::exception&) {
        fclose(f);
        return SFM_ERR_OOM;
    }
    float w0 = 0.f;
    const bool ok = fread(&w0, sizeof w0, 1, f) == 1 &&
                    fread(wf.data(), sizeof(float), wf.size(), f) == wf.size() &&
                    fread(vf.data(), sizeof(float), vf.size(), f) == vf.size();
    fclose(f);
    if (!ok) return SFM_ERR_IO;
    sfm_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = SFM_ABI_VERSION;
    cfg.task = hd.task; cfg.k = hd.k; cfg.k0 = hd.k0; cfg.k1 = hd.k1; cfg.device = device;
    cfg.n_slots = hd.n_slots; cfg.reg0 = hd.reg0; cfg.regw = hd.regw; cfg.regv = hd.regv;
    cfg.step_size = hd.step_size; cfg.mini_batch_fraction = hd.mini_batch_fraction;
    cfg.sampler_seed = hd.sampler_seed;
    cfg.sampler_mode = (int32_t)hd.pad & ~SFM_FLAG_SHARD_V;   // a loaded handle is a plain replicated model
    sfm_handle* h = nullptr;
    RC(sfm_create(&cfg, &h));
    const int rc = sfm_set_model(h, w0, wf.data(), vf.empty() ? nullptr : vf.data());
    if (rc != SFM_OK) {
        sfm_destroy(h);
        return rc;
    }
    *out = h;
    return SFM_OK;
}

// ------------------------------------------------------------------------------ scorer ---
static int predict_view(sfm_handle* h, const BatchView& b, float* out_host) {
    NEED_MODEL(h);
    RC(ensure(h, h->b_yhat, sizeof(float) * (size_t)(b.n_rows > 0 ? b.n_rows : 1)));
    PhaseTimer pt(h);
    CU(cudaMemsetAsync(h->d_err, 0, sizeof(int32_t), h->stream));
    FwdOut o;
    memset(&o, 0, sizeof o);
    o.yhat = (float*)h->b_yhat.p;
    if (is_sharded(h)) {
        RC(validate_view(h, b));
        RC(shard_forward(h, b));
    } else {
        CU(launch_forward(h->m, b, o, false, h->d_err, h->sm_count, h->stream,
                          &h->stats.kernel_launches));
    }
    pt.lap(&h->stats.ms_predict);
    if (b.n_rows > 0 && out_host) {
        CU(cudaMemcpyAsync(out_host, o.yhat, sizeof(float) * (size_t)b.n_rows,
                           cudaMemcpyDeviceToHost, h->stream));
        h->stats.d2h_bytes += (int64_t)sizeof(float) * b.n_rows;
    }
    h->stats.predict_rows += b.n_rows;
    h->stats.predict_nnz += b.nnz;
    return read_err_flag(h);
}

int32_t sfm_predict(sfm_handle* h, const int64_t* row_ptr, const int32_t* idx, const float* val,
                    int64_t n_rows, float* out) {
    if (!h) return SFM_ERR_ARG;
    if (n_rows > 0 && !out) return set_err(h, SFM_ERR_ARG, "out is NULL");
    CU(cudaSetDevice(h->device));
    BatchView b;
    RC(stage_csr(h, h->stage[2], h->stream, row_ptr, idx, val, nullptr, n_rows));
    stage_view(h->stage[2], &b);
    return predict_view(h, b, out);
}

// ------------------------------------------------------------------------------ data set --
int32_t sfm_unload_dataset(sfm_handle* h) {
    if (!h) return SFM_ERR_ARG;
    Dataset& ds = h->ds;
    if (h->stream) cudaStreamSynchronize(h->stream);
    free_parts(h);
    als_free(h);
    if (ds.row_ptr) cudaFree(ds.row_ptr);
    if (ds.idx) cudaFree(ds.idx);
    if (ds.val) cudaFree(ds.val);
    if (ds.label) cudaFree(ds.label);
    ds = Dataset();
    return SFM_OK;
}

static int alloc_dataset(sfm_handle* h, int64_t n_rows, int64_t nnz, bool with_val) {
    Dataset& ds = h->ds;
    sfm_unload_dataset(h);
    CU(cudaMalloc(&ds.row_ptr, sizeof(int64_t) * (size_t)(n_rows + 1)));
    CU(cudaMalloc(&ds.idx, sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1)));
    if (with_val) CU(cudaMalloc(&ds.val, sizeof(float) * (size_t)(nnz > 0 ? nnz : 1)));
    CU(cudaMalloc(&ds.label, sizeof(float) * (size_t)(n_rows > 0 ? n_rows : 1)));
    ds.n_rows = n_rows;
    ds.nnz = nnz;
    return SFM_OK;
}

static int check_dataset_indices(sfm_handle* h) {
    Dataset& ds = h->ds;
    ds.max_index = -1;
    if (ds.nnz == 0) return SFM_OK;
    int32_t* mm = h->d_count + 1;  // 2 ints
    const int32_t init[2] = {INT32_MAX, INT32_MIN};
    RC(ensure_pinned(h, 64));
    memcpy(h->h_pinned, init, sizeof init);
    CU(cudaMemcpyAsync(mm, h->h_pinned, sizeof init, cudaMemcpyHostToDevice, h->stream));
    CU(launch_idx_range(ds.idx, ds.nnz, mm, h->stream, &h->stats.kernel_launches));
    CU(cudaMemcpyAsync(h->h_flags + 4, mm, sizeof init, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    const int32_t lo = h->h_flags[4], hi = h->h_flags[5];
    ds.max_index = hi;
    if (lo < 0 || (int64_t)hi >= h->m.n_slots) {
        char msg[160];
        snprintf(msg, sizeof msg, "data set feature indices span [%d, %d], model has n_slots = %lld",
                 lo, hi, (long long)h->m.n_slots);
        sfm_unload_dataset(h);
        return set_err(h, SFM_ERR_INDEX, msg);
    }
    return SFM_OK;
}

int32_t sfm_load_dataset(sfm_handle* h, const int64_t* row_ptr, const int32_t* idx,
                         const float* val, const float* label, int64_t n_rows,
                         int64_t global_row_offset) {
    if (!h) return SFM_ERR_ARG;
    if (n_rows < 0 || n_rows >= 2147483647LL || (n_rows > 0 && (!row_ptr || !label)))
        return set_err(h, SFM_ERR_ARG, "bad data set arguments");
    CU(cudaSetDevice(h->device));
    const int64_t nnz = n_rows > 0 ? row_ptr[n_rows] : 0;
    if (n_rows > 0 && row_ptr[0] != 0) return set_err(h, SFM_ERR_INDEX, "row_ptr[0] must be 0");
    if (nnz > 0 && !idx) return set_err(h, SFM_ERR_ARG, "idx is NULL");
    int32_t um = n_rows > 0 ? (int32_t)(row_ptr[1] - row_ptr[0]) : -1;
    for (int64_t r = 0; r < n_rows; ++r) {
        const int64_t len = row_ptr[r + 1] - row_ptr[r];
        if (len < 0) return set_err(h, SFM_ERR_INDEX, "row_ptr is not non-decreasing");
        if (len != um) um = -1;
    }
    // a value array that is all ones is dropped: the data set then takes the all-ones kernels
    // (same arithmetic: x = 1), which move 4 bytes less per entry and sort a 4-byte payload
    if (val) {
        bool ones = true;
        for (int64_t j = 0; j < nnz && ones; ++j) ones = val[j] == 1.0f;
        if (ones) val = nullptr;
    }
    RC(alloc_dataset(h, n_rows, nnz, val != nullptr));
    Dataset& ds = h->ds;
    if (n_rows > 0) {
        CU(cudaMemcpyAsync(ds.row_ptr, row_ptr, sizeof(int64_t) * (size_t)(n_rows + 1),
                           cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(ds.label, label, sizeof(float) * (size_t)n_rows, cudaMemcpyHostToDevice,
                           h->stream));
    } else {
        CU(cudaMemsetAsync(ds.row_ptr, 0, sizeof(int64_t), h->stream));
    }
    if (nnz > 0) {
        CU(cudaMemcpyAsync(ds.idx, idx, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice,
                           h->stream));
        if (val)
            CU(cudaMemcpyAsync(ds.val, val, sizeof(float) * (size_t)nnz, cudaMemcpyHostToDevice,
                               h->stream));
    }
    h->stats.h2d_bytes += (int64_t)(sizeof(int64_t) * (n_rows + 1) + sizeof(float) * n_rows +
                                    (sizeof(int32_t) + (val ? sizeof(float) : 0)) * nnz);
    ds.global_offset = global_row_offset;
    ds.uniform_m = um;
    RC(check_dataset_indices(h));
    ds.loaded = true;
    return SFM_OK;
}

int32_t sfm_synth_ctr_dataset(sfm_handle* h, int64_t n_rows, int64_t global_row_offset,
                              int32_t n_fields, const int32_t* field_log2_card,
                              const uint32_t* zipf_cdf, const int64_t* zipf_cdf_off,
                              uint64_t seed) {
    if (!h) return SFM_ERR_ARG;
    if (n_rows < 0 || n_fields < 1 || n_fields > 1024 || !field_log2_card || !zipf_cdf ||
        !zipf_cdf_off)
        return set_err(h, SFM_ERR_ARG, "bad synthetic data set arguments");
    const int64_t nnz = n_rows * n_fields;
    if (n_rows >= 2147483647LL / n_fields) return set_err(h, SFM_ERR_ARG, "synthetic shard too large");
    CU(cudaSetDevice(h->device));
    int64_t cdf_len = 0;
    for (int f = 0; f < n_fields; ++f) {
        if (field_log2_card[f] < 0 || field_log2_card[f] > 24 || zipf_cdf_off[f] < 0)
            return set_err(h, SFM_ERR_ARG, "bad field cardinality");
        const int64_t e = zipf_cdf_off[f] + ((int64_t)1 << field_log2_card[f]);
        if (e > cdf_len) cdf_len = e;
    }
    RC(alloc_dataset(h, n_rows, nnz, false));
    Dataset& ds = h->ds;
    Buf tab_card, tab_cdf, tab_off;
    int rc = ensure(h, tab_card, sizeof(int32_t) * n_fields);
    if (rc == SFM_OK) rc = ensure(h, tab_cdf, sizeof(uint32_t) * (size_t)cdf_len);
    if (rc == SFM_OK) rc = ensure(h, tab_off, sizeof(int64_t) * n_fields);
    cudaError_t e = cudaSuccess;
    if (rc == SFM_OK) {
        e = cudaMemcpyAsync(tab_card.p, field_log2_card, sizeof(int32_t) * n_fields,
                            cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(tab_cdf.p, zipf_cdf, sizeof(uint32_t) * (size_t)cdf_len,
                                cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(tab_off.p, zipf_cdf_off, sizeof(int64_t) * n_fields,
                                cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess && n_rows == 0) e = cudaMemsetAsync(ds.row_ptr, 0, sizeof(int64_t), h->stream);
        if (e == cudaSuccess)
            e = launch_synth_ctr(n_rows, global_row_offset, n_fields, (const int32_t*)tab_card.p,
                                 (const uint32_t*)tab_cdf.p, (const int64_t*)tab_off.p, seed,
                                 h->m.n_slots, ds.idx, ds.label, ds.row_ptr, h->stream,
                                 &h->stats.kernel_launches);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    }
    free_buf(tab_card);
    free_buf(tab_cdf);
    free_buf(tab_off);
    if (rc != SFM_OK) return rc;
    if (e != cudaSuccess) return cuda_fail(h, e, "synthetic data set generation");
    ds.global_offset = global_row_offset;
    ds.uniform_m = n_fields;
    RC(check_dataset_indices(h));
    ds.loaded = true;
    return SFM_OK;
}

int32_t sfm_get_dataset_rows(sfm_handle* h, int64_t row_lo, int64_t row_hi, int64_t* row_ptr,
                             int32_t* idx, float* val, float* label) {
    if (!h) return SFM_ERR_ARG;
    const Dataset& ds = h->ds;
    if (!ds.loaded) return set_err(h, SFM_ERR_STATE, "no resident data set");
    if (row_lo < 0 || row_hi < row_lo || row_hi > ds.n_rows || !row_ptr)
        return set_err(h, SFM_ERR_ARG, "bad row range");
    CU(cudaSetDevice(h->device));
    const int64_t n = row_hi - row_lo;
    CU(cudaMemcpyAsync(row_ptr, ds.row_ptr + row_lo, sizeof(int64_t) * (size_t)(n + 1),
                       cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    const int64_t base = row_ptr[0], cnt = row_ptr[n] - base;
    for (int64_t r = 0; r <= n; ++r) row_ptr[r] -= base;
    if (idx && cnt > 0)
        CU(cudaMemcpyAsync(idx, ds.idx + base, sizeof(int32_t) * (size_t)cnt,
                           cudaMemcpyDeviceToHost, h->stream));
    if (val && cnt > 0) {
        if (ds.val)
            CU(cudaMemcpyAsync(val, ds.val + base, sizeof(float) * (size_t)cnt,
                               cudaMemcpyDeviceToHost, h->stream));
        else
            for (int64_t j = 0; j < cnt; ++j) val[j] = 1.f;
    }
    if (label && n > 0)
        CU(cudaMemcpyAsync(label, ds.label + row_lo, sizeof(float) * (size_t)n,
                           cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return SFM_OK;
}

int32_t sfm_dataset_info(sfm_handle* h, int64_t* n_rows, int64_t* nnz, int32_t* max_index) {
    if (!h) return SFM_ERR_ARG;
    if (!h->ds.loaded) return set_err(h, SFM_ERR_STATE, "no resident data set");
    if (n_rows) *n_rows = h->ds.n_rows;
    if (nnz) *nnz = h->ds.nnz;
    if (max_index) *max_index = h->ds.max_index;
    return SFM_OK;
}

int32_t sfm_predict_resident(sfm_handle* h, int64_t row_lo, int64_t row_hi, float* out) {
    if (!h) return SFM_ERR_ARG;
    const Dataset& ds = h->ds;
    if (!ds.loaded) return set_err(h, SFM_ERR_STATE, "no resident data set");
    if (row_lo < 0 || row_hi < row_lo || row_hi > ds.n_rows) return set_err(h, SFM_ERR_ARG, "bad row range");
    CU(cudaSetDevice(h->device));
    BatchView b;
    RC(resident_batch(h, nullptr, 0, &b));
    b.row_lo = row_lo;
    b.n_rows = row_hi - row_lo;
    b.nnz = 0;  // statistics only; not read back for a sub-range
    if (is_sharded(h)) RC(subrange_view(h, row_lo, row_hi, &b));
    return predict_view(h, b, out);
}

int32_t sfm_evaluate(sfm_handle* h, double metrics[5]) {
    if (!h || !metrics) return SFM_ERR_ARG;
    NEED_MODEL(h);
    const Dataset& ds = h->ds;
    if (!ds.loaded) return set_err(h, SFM_ERR_STATE, "no resident data set");
    CU(cudaSetDevice(h->device));
    const int64_t tile = 1 << 22;
    RC(ensure(h, h->b_yhat, sizeof(float) * (size_t)(ds.n_rows < tile ? (ds.n_rows > 0 ? ds.n_rows : 1) : tile)));
    RC(ensure(h, h->b_partials, sizeof(double) * 4 * 512));
    double* acc = h->d_scal + 4;  // 4 doubles
    CU(cudaMemsetAsync(acc, 0, sizeof(double) * 4, h->stream));
    CU(cudaMemsetAsync(h->d_err, 0, sizeof(int32_t), h->stream));
    for (int64_t lo = 0; lo < ds.n_rows; lo += tile) {
        BatchView b;
        RC(resident_batch(h, nullptr, 0, &b));
        b.row_lo = lo;
        b.n_rows = (ds.n_rows - lo) < tile ? (ds.n_rows - lo) : tile;
        FwdOut o;
        memset(&o, 0, sizeof o);
        o.yhat = (float*)h->b_yhat.p;
        if (is_sharded(h)) {
            RC(subrange_view(h, lo, lo + b.n_rows, &b));
            RC(shard_forward(h, b));
        } else {
            CU(launch_forward(h->m, b, o, false, h->d_err, h->sm_count, h->stream,
                              &h->stats.kernel_launches));
        }
        CU(launch_metrics(o.yhat, ds.label + lo, b.n_rows, (double*)h->b_partials.p, acc,
                          h->stream, &h->stats.kernel_launches));
        h->stats.predict_rows += b.n_rows;
    }
    // [4..7] = sums, reuse slot layout: sums[0..3], then N
    double* h5 = h->h_scal;
    if (h->world > 1) {
        // fold N into the all-reduce: stash it in d_scal[3]... keep simple: reduce 4 sums, N apart
        RC(nccl_allreduce_f64(h->nccl, h->comm, acc, 4, h->stream, &h->err));
    }
    CU(cudaMemcpyAsync(h5, acc, sizeof(double) * 4, cudaMemcpyDeviceToHost, h->stream));
    RC(read_err_flag(h));
    double n_total = (double)ds.n_rows;
    if (h->world > 1) {
        double* dn = h->d_scal + 3;
        h->h_scal[4] = n_total;
        CU(cudaMemcpyAsync(dn, h->h_scal + 4, sizeof(double), cudaMemcpyHostToDevice, h->stream));
        RC(nccl_allreduce_f64(h->nccl, h->comm, dn, 1, h->stream, &h->err));
        CU(cudaMemcpyAsync(h->h_scal + 4, dn, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        n_total = h->h_scal[4];
    }
    if (n_total > 0) {
        metrics[0] = sqrt(h5[0] / n_total);
        metrics[1] = h5[1] / n_total;
        metrics[2] = h5[2] / n_total;
        metrics[3] = h5[3] / n_total;
    } else {
        metrics[0] = metrics[1] = metrics[2] = metrics[3] = 0.0;
    }
    metrics[4] = n_total;
    return SFM_OK;
}

int32_t sfm_evaluate_auc(sfm_handle* h, double out[3]) {
    if (!h || !out) return SFM_ERR_ARG;
    NEED_MODEL(h);
    const Dataset& ds = h->ds;
    if (!ds.loaded) return set_err(h, SFM_ERR_STATE, "no resident data set");
    CU(cudaSetDevice(h->device));
    const int64_t n = ds.n_rows;
    out[0] = NAN;
    out[1] = out[2] = 0.0;
    if (n == 0) return SFM_OK;
    // scores of every resident row, tile by tile, into one buffer
    Buf& sc = h->b_S;      // reused as scratch: [scores | sorted scores | rows | sorted rows]
    RC(ensure(h, sc, sizeof(float) * 4 * (size_t)n));
    float* score = (float*)sc.p;
    float* score_s = score + n;
    uint32_t* rows = (uint32_t*)(score_s + n);
    uint32_t* rows_s = rows + n;
    RC(ensure(h, h->b_partials, sizeof(double) * 4 * 512));
    CU(cudaMemsetAsync(h->d_err, 0, sizeof(int32_t), h->stream));
    const int64_t tile = 1 << 22;
    for (int64_t lo = 0; lo < n; lo += tile) {
        BatchView b;
        RC(resident_batch(h, nullptr, 0, &b));
        b.row_lo = lo;
        b.n_rows = (n - lo) < tile ? (n - lo) : tile;
        if (is_sharded(h)) {
            RC(subrange_view(h, lo, lo + b.n_rows, &b));
            RC(shard_forward(h, b));
            CU(cudaMemcpyAsync(score + lo, h->b_yhat.p, sizeof(float) * (size_t)b.n_rows,
                               cudaMemcpyDeviceToDevice, h->stream));
        } else {
            FwdOut o;
            memset(&o, 0, sizeof o);
            o.yhat = score + lo;
            CU(launch_forward(h->m, b, o, false, h->d_err, h->sm_count, h->stream,
                              &h->stats.kernel_launches));
        }
        h->stats.predict_rows += b.n_rows;
    }
    CU(launch_iota_u32(rows, n, h->stream, &h->stats.kernel_launches));
    const size_t tb = sort_f32_u32_temp_bytes(n);
    RC(ensure(h, h->b_sort_tmp, tb));
    CU(sort_f32_u32(h->b_sort_tmp.p, tb, score, score_s, rows, rows_s, n, h->stream,
                    &h->stats.kernel_launches));
    double* d2 = h->d_scal + 4;
    CU(launch_auc_ranks(score_s, rows_s, ds.label, n, (double*)h->b_partials.p, d2, h->stream,
                        &h->stats.kernel_launches));
    CU(cudaMemcpyAsync(h->h_scal + 4, d2, sizeof(double) * 2, cudaMemcpyDeviceToHost, h->stream));
    RC(read_err_flag(h));
    const double rank_sum = h->h_scal[4], n_pos = h->h_scal[5], n_neg = (double)n - n_pos;
    out[1] = n_pos;
    out[2] = n_neg;
    if (n_pos > 0 && n_neg > 0) out[0] = (rank_sum - n_pos * (n_pos + 1.0) / 2.0) / (n_pos * n_neg);
    return SFM_OK;
}

// ------------------------------------------------------------------------------ learner ---
int32_t sfm_train_step(sfm_handle* h, const int64_t* row_ids, int64_t n_ids, int64_t iter,
                       double* mean_loss_out, int64_t* batch_out) {
    if (!h) return SFM_ERR_ARG;
    if (!h->ds.loaded) return set_err(h, SFM_ERR_STATE, "no resident data set");
    if (iter < 1) return set_err(h, SFM_ERR_ARG, "iter is 1-based");
    CU(cudaSetDevice(h->device));
    const int32_t* ids_dev = nullptr;
    int64_t n = 0;
    if (row_ids || n_ids >= 0) {
        if (n_ids < 0 || (n_ids > 0 && !row_ids)) return set_err(h, SFM_ERR_ARG, "bad row id list");
        RC(upload_row_ids(h, row_ids, n_ids, &ids_dev));
        n = n_ids;
    } else if (use_partitions(h)) {
        BatchView pb;
        const PartCache* pc = nullptr;
        RC(partition_batch(h, iter, &pb, &pc));
        RC(train_step_run(h, pb, iter, pc));
        return finish_step(h, mean_loss_out, batch_out);
    } else {
        RC(sample_device(h, iter, &ids_dev, &n));
    }
    BatchView b;
    RC(resident_batch(h, ids_dev, n, &b));
    if (b.nnz >= 2147483647LL) return set_err(h, SFM_ERR_ARG, "batch nnz must be < 2^31-1");
    RC(train_step_run(h, b, iter));
    return finish_step(h, mean_loss_out, batch_out);
}

int32_t sfm_train_step_csr(sfm_handle* h, const int64_t* row_ptr, const int32_t* idx,
                           const float* val, const float* label, int64_t n_rows, int64_t iter,
                           double* mean_loss_out, int64_t* batch_out) {
    if (!h) return SFM_ERR_ARG;
    if (iter < 1) return set_err(h, SFM_ERR_ARG, "iter is 1-based");
    if (n_rows > 0 && !label) return set_err(h, SFM_ERR_ARG, "label is NULL");
    CU(cudaSetDevice(h->device));
    BatchView b;
    RC(stage_csr(h, h->stage[2], h->stream, row_ptr, idx, val, label, n_rows));
    stage_view(h->stage[2], &b);
    RC(train_step_run(h, b, iter));
    return finish_step(h, mean_loss_out, batch_out);
}

int32_t sfm_stage_csr(sfm_handle* h, int32_t slot, const int64_t* row_ptr, const int32_t* idx,
                      const float* val, const float* label, int64_t n_rows) {
    if (!h) return SFM_ERR_ARG;
    if (slot < 0 || slot > 1) return set_err(h, SFM_ERR_ARG, "slot must be 0 or 1");
    if (n_rows > 0 && !label) return set_err(h, SFM_ERR_ARG, "label is NULL");
    CU(cudaSetDevice(h->device));
    Stage& sg = h->stage[slot];
    RC(stage_csr(h, sg, h->copy_stream, row_ptr, idx, val, label, n_rows));
    CU(cudaEventRecord(sg.ready, h->copy_stream));
    return SFM_OK;
}

int32_t sfm_stage_onehot(sfm_handle* h, int32_t slot, const uint32_t* packed_idx,
                         const uint32_t* label_bits, const float* label_f32, int64_t n_rows,
                         int32_t m, int32_t id_bits) {
    if (!h) return SFM_ERR_ARG;
    if (slot < 0 || slot > 1) return set_err(h, SFM_ERR_ARG, "slot must be 0 or 1");
    if (is_sharded(h) || h->shard_requested)
        return set_err(h, SFM_ERR_STATE, "sfm_stage_onehot is not available with SFM_FLAG_SHARD_V");
    CU(cudaSetDevice(h->device));
    Stage& sg = h->stage[slot];
    RC(stage_onehot(h, sg, h->copy_stream, packed_idx, label_bits, label_f32, n_rows, m, id_bits));
    CU(cudaEventRecord(sg.ready, h->copy_stream));
    return SFM_OK;
}

int32_t sfm_train_step_staged(sfm_handle* h, int32_t slot, int64_t iter, double* mean_loss_out,
                              int64_t* batch_out) {
    if (!h) return SFM_ERR_ARG;
    if (slot < 0 || slot > 1) return set_err(h, SFM_ERR_ARG, "slot must be 0 or 1");
    if (iter < 1) return set_err(h, SFM_ERR_ARG, "iter is 1-based");
    Stage& sg = h->stage[slot];
    if (!sg.valid) return set_err(h, SFM_ERR_STATE, "nothing staged in this slot");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamWaitEvent(h->stream, sg.ready, 0));
    BatchView b;
    stage_view(sg, &b);
    RC(train_step_run(h, b, iter));
    sg.valid = false;
    return finish_step(h, mean_loss_out, batch_out);
}

// Queues the sampling of iteration `iter` into slot `slot` on the copy stream (it depends on the
// seed only, so it runs while the previous iteration computes).
static int sample_prefetch(sfm_handle* h, int64_t iter, int slot, bool plan = false) {
    const Dataset& ds = h->ds;
    const double frac = (double)h->cfg.mini_batch_fraction;
    const uint64_t thr = (uint64_t)floor(frac * 9007199254740992.0);
    const uint64_t key = mix64(h->cfg.sampler_seed + (uint64_t)iter);
    CU(cudaStreamWaitEvent(h->copy_stream, h->ev_used[slot], 0));  // slot's previous consumer
    CU(sample_rows_device(h->b_samp_tmp.p, h->b_samp_tmp.cap, ds.n_rows, ds.global_offset, key, thr,
                          (int32_t*)h->b_ids2[slot].p, h->d_count2 + slot, h->copy_stream,
                          &h->stats.kernel_launches));
    CU(cudaMemcpyAsync(h->h_count2 + slot, h->d_count2 + slot, sizeof(int32_t),
                       cudaMemcpyDeviceToHost, h->copy_stream));
    CU(cudaEventRecord(h->ev_samp[slot], h->copy_stream));
    h->stats.d2h_bytes += 4;
    if (plan) {
        // the batch's feature ids do not depend on the model: its bucket counts and work-item plan
        // are built here, behind the sampler, and the step starts at the scatter
        CU(bucket_count_plan_rows(h->m, h->ahead_geom, ds.idx, (const int32_t*)h->b_ids2[slot].p,
                                  h->d_count2 + slot, h->ahead_cap, ds.uniform_m, h->b_bkt_work2[slot].p,
                                  h->b_bkt_tables2[slot].p, reinterpret_cast<unsigned int*>(h->d_count + 5),
                                  h->sm_count, h->copy_stream, &h->stats.kernel_launches));
        CU(cudaEventRecord(h->ev_plan[slot], h->copy_stream));
    }
    return SFM_OK;
}

// Grows a buffer that the copy stream works on (ensure() only drains the compute stream).
static int ensure_side(sfm_handle* h, Buf& b, size_t bytes) {
    if (bytes <= b.cap) return SFM_OK;
    CU(cudaStreamSynchronize(h->copy_stream));
    return ensure(h, b, bytes);
}

// Bernoulli sampler, uniform rows: sample the batch of iteration `iter`, emit its entry list and
// radix-sort it by feature -- all on the copy stream, one iteration ahead, so that the sort
// (DRAM-bandwidth bound) runs concurrently with the previous iteration's forward (L1 bound) and
// reduce (latency bound) instead of after them.  The host only waits for the
// 4-byte batch size.
static int transpose_prefetch(sfm_handle* h, int64_t iter, int slot) {
    const Dataset& ds = h->ds;
    const ModelView& m = h->m;
    int64_t* L = &h->stats.kernel_launches;
    const double frac = (double)h->cfg.mini_batch_fraction;
    const uint64_t thr = (uint64_t)floor(frac * 9007199254740992.0);
    const uint64_t key = mix64(h->cfg.sampler_seed + (uint64_t)iter);
    PartCache& pc = h->pre[slot];
    cudaStream_t cs = h->copy_stream;
    CU(cudaStreamWaitEvent(cs, h->ev_used[slot], 0));   // the slot's previous consumer is done
    CU(sample_rows_device(h->b_samp_tmp.p, h->b_samp_tmp.cap, ds.n_rows, ds.global_offset, key, thr,
                          (int32_t*)h->b_ids2[slot].p, h->d_count2 + slot, cs, L));
    CU(cudaMemcpyAsync(h->h_count2 + slot, h->d_count2 + slot, sizeof(int32_t),
                       cudaMemcpyDeviceToHost, cs));
    CU(cudaEventRecord(h->ev_pool[18], cs));
    CU(cudaEventSynchronize(h->ev_pool[18]));
    h->stats.d2h_bytes += 4;
    const int64_t n = h->h_count2[slot];
    const int64_t nnz = n * ds.uniform_m;
    if (nnz >= 2147483647LL) return set_err(h, SFM_ERR_ARG, "batch nnz must be < 2^31-1");
    pc.n_rows = n;
    pc.nnz = nnz;
    pc.n_slices = 0;
    pc.key_bits = bits_for(m.n_slots);
    pull_plan(m, n, &pc.blk_shift, &pc.n_blocks);
    int blk_bits = 0;
    while (((int64_t)1 << blk_bits) < pc.n_blocks) ++blk_bits;
    if (pc.key_bits + blk_bits > 32) return set_err(h, SFM_ERR_ARG, "sort key does not fit 32 bits");
    const bool binary = ds.val == nullptr;
    const size_t pay_sz = binary ? sizeof(uint32_t) : sizeof(uint2);
    const size_t cnt = (size_t)(nnz > 0 ? nnz : 1);
    RC(ensure_side(h, h->b_pre_keys0, sizeof(uint32_t) * cnt));
    RC(ensure_side(h, h->b_pre_pay0, pay_sz * cnt));
    RC(ensure_side(h, pc.keys, sizeof(uint32_t) * cnt));
    RC(ensure_side(h, pc.pay, pay_sz * cnt));
    if (nnz > 0) {
        BatchView v;
        v.row_ptr = ds.row_ptr;
        v.idx = ds.idx;
        v.val = ds.val;
        v.label = ds.label;
        v.row_ids = (const int32_t*)h->b_ids2[slot].p;
        v.row_lo = 0;
        v.n_rows = n;
        v.nnz = nnz;
        v.idx_len = ds.nnz;
        v.out_ptr = nullptr;
        v.out_base = 0;
        v.uniform_m = ds.uniform_m;
        v.validated = true;
        CU(launch_emit(v, pc.key_bits, pc.blk_shift, m.n_slots, (uint32_t*)h->b_pre_keys0.p,
                       (uint2*)h->b_pre_pay0.p, h->sm_count, cs, L));
        const int end_bit = pc.key_bits + blk_bits;
        const size_t sb = binary ? sort_pairs32_temp_bytes(nnz, end_bit) : sort_pairs_temp_bytes(nnz, end_bit);
        RC(ensure_side(h, h->b_sort_tmp2, sb));
        if (binary)
            CU(sort_pairs32(h->b_sort_tmp2.p, sb, (const uint32_t*)h->b_pre_keys0.p,
                            (uint32_t*)pc.keys.p, (const uint32_t*)h->b_pre_pay0.p,
                            (uint32_t*)pc.pay.p, nnz, end_bit, cs, L));
        else
            CU(sort_pairs(h->b_sort_tmp2.p, sb, (const uint32_t*)h->b_pre_keys0.p,
                          (uint32_t*)pc.keys.p, (const uint2*)h->b_pre_pay0.p, (uint2*)pc.pay.p, nnz,
                          end_bit, cs, L));
    }
    pc.built = true;
    CU(cudaEventRecord(h->ev_samp[slot], cs));
    return SFM_OK;
}

int32_t sfm_train(sfm_handle* h, int64_t first_iter, int64_t n_iters, double* loss_history) {
    if (!h) return SFM_ERR_ARG;
    if (!h->ds.loaded) return set_err(h, SFM_ERR_STATE, "no resident data set");
    if (first_iter < 1 || n_iters < 0) return set_err(h, SFM_ERR_ARG, "bad iteration range");
    CU(cudaSetDevice(h->device));
    const Dataset& ds = h->ds;
    const double frac = (double)h->cfg.mini_batch_fraction;
    const bool parts = use_partitions(h);
    const bool sampled = !parts && frac < 1.0 && frac > 0.0 && ds.n_rows > 0;
    // uniform rows: the whole transposition of the next batch is built ahead on the copy stream
    const int ahead_env = knobs().sort_ahead;   // measured: co-scheduling the sort with the gather kernels does not pay
    const bool ahead = sampled && ds.uniform_m >= 0 && !is_sharded(h) && !h->phase_timing &&
                       ahead_env != 0;
    // plan-ahead: all-ones / valued rows of one length, bucket form -- the next batch's bucket counts
    // and plan ride behind its sampler on the (low-priority) copy stream; the work buffers are
    // carved for a batch of up to ahead_cap rows (a larger draw simply plans inside the step)
    bool plan_ahead = false;
    if (sampled && !ahead && knobs().plan_ahead && knobs().stream_prio && ds.uniform_m > 0 &&
        !is_sharded(h) && !h->phase_timing && knobs().pull_block_mb <= 0) {
        int64_t cap = (int64_t)(frac * (double)ds.n_rows * 1.25) + 65536;
        if (cap > ds.n_rows) cap = ds.n_rows;
        if (cap < 2147483647LL && bucket_geometry(h->m, bits_for(h->m.n_slots), cap, cap * ds.uniform_m, &h->ahead_geom)) {
            h->ahead_cap = (int32_t)cap;
            plan_ahead = true;
        }
    }
    if (sampled) {
        for (int i = 0; i < 2; ++i) RC(ensure(h, h->b_ids2[i], sizeof(int32_t) * (size_t)ds.n_rows));
        RC(ensure(h, h->b_samp_tmp, select_temp_bytes(ds.n_rows)));
        if (plan_ahead)
            for (int i = 0; i < 2; ++i) {
                RC(ensure(h, h->b_bkt_work2[i], bucket_work_bytes(h->m, h->ahead_geom, h->sm_count)));
                RC(ensure(h, h->b_bkt_tables2[i], bucket_tables_bytes(h->ahead_geom)));
            }
        CU(cudaEventRecord(h->ev_used[0], h->stream));
        CU(cudaEventRecord(h->ev_used[1], h->stream));
        if (n_iters > 0) RC(ahead ? transpose_prefetch(h, first_iter, 0) : sample_prefetch(h, first_iter, 0, plan_ahead));
    }
    double* hist = nullptr;
    if (n_iters > 0) CU(cudaMallocHost(&hist, sizeof(double) * SC_N * (size_t)n_iters));
    int rc = SFM_OK;
    for (int64_t t = 0; t < n_iters && rc == SFM_OK; ++t) {
        const int32_t* ids_dev = nullptr;
        int64_t n = 0;
        const int slot = (int)(t & 1);
        if (ahead) {
            // the batch (rows + sorted entries) was prepared on the copy stream; only the device waits
            const PartCache& pre = h->pre[slot];
            BatchView b;
            b.row_ptr = ds.row_ptr;
            b.idx = ds.idx;
            b.val = ds.val;
            b.label = ds.label;
            b.row_ids = (const int32_t*)h->b_ids2[slot].p;
            b.row_lo = 0;
            b.n_rows = pre.n_rows;
            b.nnz = pre.nnz;
            b.idx_len = ds.nnz;
            b.out_ptr = nullptr;
            b.out_base = 0;
            b.uniform_m = ds.uniform_m;
            b.validated = true;
            if (cudaStreamWaitEvent(h->stream, h->ev_samp[slot], 0) != cudaSuccess)
                rc = set_err(h, SFM_ERR_CUDA, "cudaStreamWaitEvent failed");
            if (rc == SFM_OK) rc = train_step_run(h, b, first_iter + t, &pre);
            if (rc == SFM_OK && cudaEventRecord(h->ev_used[slot], h->stream) != cudaSuccess)
                rc = set_err(h, SFM_ERR_CUDA, "cudaEventRecord failed");
            if (rc == SFM_OK &&
                cudaMemcpyAsync(hist + SC_N * t, h->d_scal, sizeof(double) * SC_N,
                                cudaMemcpyDeviceToHost, h->stream) != cudaSuccess)
                rc = set_err(h, SFM_ERR_CUDA, "loss history copy failed");
            if (rc == SFM_OK && t + 1 < n_iters) rc = transpose_prefetch(h, first_iter + t + 1, slot ^ 1);
            continue;
        }
        if (sampled) {
            if (cudaEventSynchronize(h->ev_samp[slot]) != cudaSuccess) {
                rc = set_err(h, SFM_ERR_CUDA, "sampler prefetch failed");
                break;
            }
            n = h->h_count2[slot];
            ids_dev = (const int32_t*)h->b_ids2[slot].p;
            if (cudaStreamWaitEvent(h->stream, h->ev_samp[slot], 0) != cudaSuccess ||
                (plan_ahead && cudaStreamWaitEvent(h->stream, h->ev_plan[slot], 0) != cudaSuccess))
                rc = set_err(h, SFM_ERR_CUDA, "cudaStreamWaitEvent failed");
            // next iteration's batch is drawn (and planned) while this one computes
            if (rc == SFM_OK && t + 1 < n_iters) rc = sample_prefetch(h, first_iter + t + 1, slot ^ 1, plan_ahead);
        } else if (!parts) {
            rc = sample_device(h, first_iter + t, &ids_dev, &n);
        }
        BatchView b;
        const PartCache* pc = nullptr;
        if (parts) {
            rc = partition_batch(h, first_iter + t, &b, &pc);
        } else {
            if (rc == SFM_OK) rc = resident_batch(h, ids_dev, n, &b);
            if (rc == SFM_OK && b.nnz >= 2147483647LL) rc = set_err(h, SFM_ERR_ARG, "batch nnz must be < 2^31-1");
        }
        AheadPlan apv;
        if (plan_ahead) {
            apv.geom = h->ahead_geom;
            apv.cap_rows = h->ahead_cap;
            apv.work = h->b_bkt_work2[slot].p;
            apv.tables = h->b_bkt_tables2[slot].p;
        }
        if (rc == SFM_OK) rc = train_step_run(h, b, first_iter + t, pc, plan_ahead ? &apv : nullptr);
        if (rc == SFM_OK && sampled && cudaEventRecord(h->ev_used[slot], h->stream) != cudaSuccess)
            rc = set_err(h, SFM_ERR_CUDA, "cudaEventRecord failed");
        if (rc == SFM_OK &&
            cudaMemcpyAsync(hist + SC_N * t, h->d_scal, sizeof(double) * SC_N,
                            cudaMemcpyDeviceToHost, h->stream) != cudaSuccess)
            rc = set_err(h, SFM_ERR_CUDA, "loss history copy failed");
    }
    if (rc == SFM_OK) rc = read_err_flag(h);
    else cudaStreamSynchronize(h->stream);
    cudaStreamSynchronize(h->copy_stream);
    if (rc == SFM_OK && h->world > 1 && !is_sharded(h))
        for (int64_t t = 0; t < n_iters; ++t)
            if (hist[SC_N * t + SC_ERR] != 0.0) {
                rc = set_err(h, SFM_ERR_INDEX,
                             "a rank saw a feature index outside [0, n_slots): every rank skipped that update");
                break;
            }
    if (rc == SFM_OK && loss_history)
        for (int64_t t = 0; t < n_iters; ++t) {
            const double c = hist[SC_N * t + SC_COUNT];
            loss_history[t] = c > 0.0 ? hist[SC_N * t + SC_LOSS] / c : 0.0;
        }
    h->stats.d2h_bytes += (int64_t)sizeof(double) * SC_N * n_iters;
    if (hist) cudaFreeHost(hist);
    return rc;
}

// ------------------------------------------------------------------------------ ALS -------
int32_t sfm_als_sweep(sfm_handle* h, int32_t flags, double* rmse_out) {
    if (!h) return SFM_ERR_ARG;
    NEED_MODEL(h);
    if (h->world > 1 || is_sharded(h))
        return set_err(h, SFM_ERR_STATE, "ALS runs on one GPU with a replicated model");
    const Dataset& ds = h->ds;
    if (!ds.loaded) return set_err(h, SFM_ERR_STATE, "no resident data set");
    CU(cudaSetDevice(h->device));
    BatchView b;
    RC(resident_batch(h, nullptr, ds.n_rows, &b));
    RC(als_sweep(h, b, flags, rmse_out));
    h->stats.train_steps += 1;
    h->stats.train_rows += ds.n_rows;
    h->stats.train_nnz += ds.nnz * (int64_t)(h->m.k + 1);
    return SFM_OK;
}

int32_t sfm_als_residuals(sfm_handle* h, double* out, int64_t n) {
    if (!h || !out || n < 0) return SFM_ERR_ARG;
    CU(cudaSetDevice(h->device));
    return als_residuals(h, out, n);
}

int32_t sfm_gradient(sfm_handle* h, const int64_t* row_ids, int64_t n_ids, float* grad_v,
                     float* grad_w, float* grad_w0, double* loss_sum, int64_t* batch_out) {
    if (!h) return SFM_ERR_ARG;
    if (!h->ds.loaded) return set_err(h, SFM_ERR_STATE, "no resident data set");
    CU(cudaSetDevice(h->device));
    const int32_t* ids_dev = nullptr;
    int64_t n = h->ds.n_rows;
    if (row_ids || n_ids >= 0) {
        if (n_ids < 0 || (n_ids > 0 && !row_ids)) return set_err(h, SFM_ERR_ARG, "bad row id list");
        RC(upload_row_ids(h, row_ids, n_ids, &ids_dev));
        n = n_ids;
    }
    BatchView b;
    RC(resident_batch(h, ids_dev, n, &b));
    RC(train_core(h, b, 1, true));
    const ModelView& m = h->m;
    const float* g = (const float*)h->b_grad.p;
    RC(download_v(h, g, grad_v));
    if (grad_w)
        CU(cudaMemcpyAsync(grad_w, g + m.n_slots * m.kp, sizeof(float) * (size_t)m.n_slots,
                           cudaMemcpyDeviceToHost, h->stream));
    RC(ensure_pinned(h, 64));
    CU(cudaMemcpyAsync(h->h_pinned, g + m.n_slots * m.kp + m.n_slots, sizeof(float),
                       cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(h->h_scal, h->d_scal, sizeof(double) * SC_N, cudaMemcpyDeviceToHost,
                       h->stream));
    RC(read_err_flag(h));
    if (grad_w0) *grad_w0 = *(float*)h->h_pinned;
    if (loss_sum) *loss_sum = h->h_scal[SC_LOSS];
    if (batch_out) *batch_out = (int64_t)h->h_scal[SC_COUNT];
    return SFM_OK;
}

// ------------------------------------------------------------------------------ multi-GPU -
int32_t sfm_comm_unique_id(uint8_t id[SFM_UNIQUE_ID_BYTES]) {
    if (!id) return SFM_ERR_ARG;
    std::string err;
    Nccl* n = nccl_load(&err);
    if (!n) return SFM_ERR_NCCL;
    return nccl_unique_id(n, id, &err);
}

int32_t sfm_comm_init(sfm_handle* h, const uint8_t id[SFM_UNIQUE_ID_BYTES], int32_t rank,
                      int32_t world_size) {
    if (!h || !id) return SFM_ERR_ARG;
    if (world_size < 1 || rank < 0 || rank >= world_size) return set_err(h, SFM_ERR_ARG, "bad rank / world size");
    if (h->comm) return set_err(h, SFM_ERR_STATE, "communicator already initialised");
    CU(cudaSetDevice(h->device));
    h->nccl = nccl_load(&h->err);
    if (!h->nccl) return SFM_ERR_NCCL;
    RC(nccl_init(h->nccl, &h->comm, id, rank, world_size, &h->err));
    h->rank = rank;
    h->world = world_size;
    if (h->shard_requested) {
        ShardState* ss = new (std::nothrow) ShardState;
        if (!ss) return set_err(h, SFM_ERR_OOM, "host allocation failed");
        const ModelView& m = h->m;
        ss->n_per = (m.n_slots + world_size - 1) / world_size;
        ss->own_lo = (int64_t)rank * ss->n_per;
        ss->n_own = m.n_slots - ss->own_lo;
        if (ss->n_own > ss->n_per) ss->n_own = ss->n_per;
        if (ss->n_own < 0) ss->n_own = 0;
        const size_t rows = (size_t)(ss->n_per > 0 ? ss->n_per : 1);
        if (cudaMalloc(&ss->v, sizeof(float) * rows * m.kp) != cudaSuccess ||
            cudaMalloc(&ss->w, sizeof(float) * rows) != cudaSuccess) {
            cudaGetLastError();
            if (ss->v) cudaFree(ss->v);
            delete ss;
            return set_err(h, SFM_ERR_OOM, "cannot allocate the model shard");
        }
        cudaMemsetAsync(ss->v, 0, sizeof(float) * rows * m.kp, h->stream);
        cudaMemsetAsync(ss->w, 0, sizeof(float) * rows, h->stream);
        CU(cudaStreamSynchronize(h->stream));
        h->shard = ss;
    } else {
        RC(p2p_setup(h));
    }
    return SFM_OK;
}

int32_t sfm_comm_info(const sfm_handle* h, int32_t* rank, int32_t* world_size) {
    if (!h) return SFM_ERR_ARG;
    if (rank) *rank = h->rank;
    if (world_size) *world_size = h->world;
    return SFM_OK;
}

int32_t sfm_comm_mode(const sfm_handle* h, int32_t* mode) {
    if (!h || !mode) return SFM_ERR_ARG;
    *mode = h->world <= 1 ? SFM_COMM_NONE
            : h->shard    ? SFM_COMM_SHARDED
            : h->p2p      ? SFM_COMM_PEER
                          : SFM_COMM_NCCL;
    return SFM_OK;
}

int32_t sfm_comm_broadcast_model(sfm_handle* h) {
    if (!h) return SFM_ERR_ARG;
    if (h->world <= 1 || is_sharded(h)) return SFM_OK;   // shards are disjoint: nothing to copy
    NEED_MODEL(h);
    CU(cudaSetDevice(h->device));
    const ModelView& m = h->m;
    RC(nccl_bcast_f32(h->nccl, h->comm, m.v, (size_t)m.n_slots * m.kp, 0, h->stream, &h->err));
    RC(nccl_bcast_f32(h->nccl, h->comm, m.w, (size_t)m.n_slots, 0, h->stream, &h->err));
    RC(nccl_bcast_f32(h->nccl, h->comm, m.w0, 1, 0, h->stream, &h->err));
    CU(cudaStreamSynchronize(h->stream));
    return SFM_OK;
}

// ------------------------------------------------------------------------------ stats -----
int32_t sfm_stats_get(sfm_handle* h, sfm_stats* out) {
    if (!h || !out) return SFM_ERR_ARG;
    *out = h->stats;
    return SFM_OK;
}

int32_t sfm_stats_reset(sfm_handle* h) {
    if (!h) return SFM_ERR_ARG;
    memset(&h->stats, 0, sizeof h->stats);
    return SFM_OK;
}

int32_t sfm_set_phase_timing(sfm_handle* h, int32_t enabled) {
    if (!h) return SFM_ERR_ARG;
    h->phase_timing = enabled != 0;
    return SFM_OK;
}

int32_t sfm_synchronize(sfm_handle* h) {
    if (!h) return SFM_ERR_ARG;
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    return SFM_OK;
}

int32_t sfm_timer_start(sfm_handle* h) {
    if (!h) return SFM_ERR_ARG;
    CU(cudaSetDevice(h->device));
    CU(cudaEventRecord(h->ev_t0, h->stream));
    return SFM_OK;
}

int32_t sfm_timer_stop(sfm_handle* h, float* ms) {
    if (!h || !ms) return SFM_ERR_ARG;
    CU(cudaSetDevice(h->device));
    CU(cudaEventRecord(h->ev_t1, h->stream));
    CU(cudaEventSynchronize(h->ev_t1));
    CU(cudaEventElapsedTime(ms, h->ev_t0, h->ev_t1));
    return SFM_OK;
}

}  // extern "C"
