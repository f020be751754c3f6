// sfm_common.h -- internal declarations shared by the translation units of libsparkfm_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/sparkfm_b200.h"

namespace sfm {

// SMs of the CURRENT device (cached per device): grid caps of the grid-stride helper kernels are
// multiples of it instead of a literal 148.
inline int device_sm_count() {
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cache[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
        cache[dev] = n;
    }
    return cache[dev];
}

// ---- device scalar block (doubles): written by the forward reduction, all-reduced across
// ranks, read by the update kernels.
enum { SC_LOSS = 0, SC_COUNT = 1, SC_GW0 = 2, SC_ERR = 3, SC_N = 4 };   // SC_ERR: ranks that saw a bad index

struct Buf {  // growable device allocation
    void* p = nullptr;
    size_t cap = 0;
};

struct Dataset {
    int64_t n_rows = 0, nnz = 0, global_offset = 0;
    int64_t* row_ptr = nullptr;  // [n_rows+1]
    int32_t* idx = nullptr;      // [nnz]
    float* val = nullptr;        // [nnz] or nullptr (all ones)
    float* label = nullptr;      // [n_rows]
    int32_t uniform_m = -1;      // every row has exactly this many entries, or -1
    int32_t max_index = -1;
    bool loaded = false;
};

// Device staging area of one host CSR mini-batch (slots 0/1: sfm_stage_csr, slot 2: the
// synchronous entry points).
struct Stage {
    Buf rowptr, idx, val, label;
    Buf packed, lbits;           // sfm_stage_onehot: bit-packed ids / label bits as copied
    int32_t* d_bad = nullptr;    // sfm_stage_onehot: set by the unpack kernel on an id >= n_slots
    int32_t uniform_m = -1;      // >= 0: uniform all-ones rows (no row_ptr), ids validated
    int64_t n_rows = 0, nnz = 0;
    bool has_val = false, has_label = false, valid = false;
    cudaEvent_t ready = nullptr;
};

// Resident transposition of one fixed mini-batch (PARTITION sampler / full batch): the batch's
// entries sorted by (feature, batch position), built at first use -- the analogue of the
// reference's cached `transposeInput` (DataSet.scala:48).
// Bucket form of the transposition (sfm_bucket.cu): feature id = bucket (top HB bits) | local id
// (low LB bits); NB = 2^HB buckets; a bucket is reduced in work items of <= 32768 entries.
struct BucketGeom {
    int LB = 0, HB = 0, NB = 0;
    int64_t max_items = 0;
};

// counts + plan of a batch built one step ahead (bucket_count_plan_rows): geometry they were
// carved with, and the per-slot work / table buffers
struct AheadPlan {
    BucketGeom geom;
    int64_t cap_rows = 0;   // the count kernel ignores a larger draw: so must the step
    void* work = nullptr;
    void* tables = nullptr;
};

struct PartCache {
    bool built = false;
    int64_t n_rows = 0, nnz = 0;
    int key_bits = 0, blk_shift = 30, n_blocks = 1;
    Buf row_ids, keys, pay;
    bool bucket = false;        // keys = packed entries grouped by bucket, pay = x bits (or unused)
    BucketGeom geom;
    Buf tables;                 // bucket_off | item_start
    int n_slices = 0;           // multi-GPU overlap: sorted position where each feature slice starts
    int32_t slice_pos[16] = {0};
};

// Row-sharded model (sfm_shard.cu).  ShardBatch = everything about a batch that depends on its
// ROW SET only (not on the model): sorted entries with compact feature ids, the unique features,
// the request lists exchanged with the owners, the compact copy of the batch.  Built per step for
// sampled batches; built once and kept for the fixed batches of the PARTITION sampler.
struct ShardBatch {
    bool built = false;
    int64_t U = 0, R = 0;   // unique features of this batch; rows requested from this rank
    std::vector<int64_t> send_off, send_cnt, recv_off, recv_cnt;
    Buf crank, pay, uniq, req, bidx, bval, blabel, optr;
};

// this rank owns features [own_lo, own_lo + n_own)
struct ShardState {
    int64_t n_per = 0, own_lo = 0, n_own = 0;
    float* v = nullptr;   // [n_per][kp]
    float* w = nullptr;   // [n_per]
    Buf flags, keys_tmp, small, lut, out_v, out_w, t_v, t_w, gr_v, gr_w, acc;
    ShardBatch scratch;
    std::vector<ShardBatch> cached;   // one per PARTITION mini-batch
    int32_t* h_small = nullptr;
    size_t h_small_cap = 0;
};

// One mini-batch as the kernels see it.
struct BatchView {
    const int64_t* row_ptr;  // CSR row pointers of the array the rows live in
    const int32_t* idx;
    const float* val;        // may be nullptr
    const float* label;
    const int32_t* row_ids;  // batch position -> row of that CSR, or nullptr = row_lo + pos
    int64_t row_lo;
    int64_t n_rows;          // batch rows on this rank
    int64_t nnz;             // batch entries on this rank
    int64_t idx_len;         // length of idx / val (row_ptr values outside [0, idx_len] are errors)
    const int64_t* out_ptr;  // batch position -> first output slot (+out_base), nullptr = pos*m
    int64_t out_base;
    int32_t uniform_m;
    bool validated = false;  // idx already checked against n_slots (resident data set)
    const int32_t* pre_err = nullptr;   // device flag raised while the batch was staged (or null)
};

struct ModelView {
    float* v;   // [n_slots][kp]
    float* w;   // [n_slots]
    float* w0;  // [1]
    int64_t n_slots;
    int32_t k, kp, lpr;
    int32_t k0, k1, task;
};

struct Nccl;      // sfm_nccl.cpp
struct P2PState;  // sfm_p2p.cu
struct AlsState;  // sfm_als.cu

}  // namespace sfm

struct sfm_handle {
    sfm_config cfg;
    sfm::ModelView m;
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_t0 = nullptr, ev_t1 = nullptr, ev_copy = nullptr;
    sfm::Dataset ds;
    // batch scratch (all growable)
    sfm::Buf b_row_ids, b_out_ptr, b_S, b_mult, b_loss, b_yhat, b_keys[2], b_pay[2], b_seg,
        b_sort_tmp, b_grad, b_partials, b_sel_tmp, b_lens, b_pull, b_bkt_work, b_bkt_tables;
    sfm::Stage stage[3];
    std::vector<sfm::PartCache> parts;  // PARTITION sampler caches (size P)
    sfm::PartCache pre[2];              // Bernoulli sampler: transposition built one step ahead
    sfm::Buf b_pre_keys0, b_pre_pay0, b_sort_tmp2;
    bool shard_requested = false;       // SFM_FLAG_SHARD_V at create; active once comm is up
    sfm::ShardState* shard = nullptr;
    sfm::P2PState* p2p = nullptr;       // NVLink peer-memory reduce+update (replicated multi-GPU)
    sfm::AlsState* als = nullptr;       // ALS: resident transposed input, level schedule, e / q caches
    // sampler prefetch (sfm_train): ids / count of the NEXT iteration are produced on copy_stream
    sfm::Buf b_ids2[2], b_samp_tmp;
    int32_t* d_count2 = nullptr;   // [2] device
    int32_t* h_count2 = nullptr;   // [2] pinned
    cudaEvent_t ev_samp[2] = {nullptr, nullptr}, ev_used[2] = {nullptr, nullptr};
    // plan-ahead (sfm_train, Bernoulli sampler, all-ones uniform rows, bucket form): the bucket
    // counts + work-item plan of the NEXT batch are built on copy_stream right behind its sampler
    sfm::Buf b_bkt_work2[2], b_bkt_tables2[2];
    sfm::BucketGeom ahead_geom;
    int32_t ahead_cap = 0;         // rows the ahead buffers were carved for
    cudaEvent_t ev_plan[2] = {nullptr, nullptr};
    cudaEvent_t ev_pool[20] = {nullptr};   // reduce / all-reduce overlap (multi-GPU)
    cudaStream_t comm_stream = nullptr;
    int32_t* d_slice = nullptr;            // [16] device
    int32_t* h_slice = nullptr;            // [16] pinned
    double* d_scal = nullptr;  // [SC_N] device
    int32_t* d_err = nullptr;  // device error flag
    int32_t* d_count = nullptr;  // device int (sampler count)
    double* h_scal = nullptr;  // pinned [SC_N]
    int32_t* h_flags = nullptr;  // pinned [16]: err, count, int64 total, min/max, 2 x int64
    void* h_pinned = nullptr;  // pinned staging for small host copies
    size_t h_pinned_cap = 0;
    // comm
    sfm::Nccl* nccl = nullptr;
    void* comm = nullptr;
    int rank = 0, world = 1;
    bool phase_timing = false;
    // one SGD step replayed as a CUDA graph: the step's launches are captured every step, the
    // resident executable graph is updated in place (same topology, new arguments) and launched
    cudaGraphExec_t step_exec = nullptr;
    bool capturing = false, capture_abort = false;
    sfm_stats stats{};
    std::string err;
};

namespace sfm {

// ---- experiment knobs: environment variables, read when a handle is created
// (never inside a launch path)
struct Knobs {
    bool bucket = true;        // SFM_BUCKET=0: two-pass global sort + chunked reduce instead of sfm_bucket.cu
    bool sort_cub = false;     // SFM_SORT=cub: library radix sort (implies !bucket)
    bool no_fastpath = false;  // SFM_NO_FASTPATH: generic forward kernel only
    int pull_block_mb = 0;     // SFM_PULL_BLOCK_MB
    int ar_slices = 1;         // SFM_AR_SLICES
    int sort_ahead = 0;        // SFM_SORT_AHEAD
    bool bucket_cache = false; // SFM_BUCKET_CACHE=1: PARTITION caches keep the bucket form
    bool stream_prio = false;  // SFM_STREAM_PRIO=1: compute stream at the greatest, copy stream at the least priority
    bool plan_ahead = false;   // SFM_PLAN_AHEAD=1: bucket counts + plan one step ahead on the copy stream (measured slower)
    bool p2p_sparse = true;    // SFM_P2P_SPARSE=0: peer-memory exchange moves the dense gradient
    bool step_graph = true;    // SFM_GRAPH=0: plain stream launches instead of the per-step CUDA graph
};
const Knobs& knobs();
void knobs_refresh();

// ---- helpers (sfm_api.cu)
int set_err(sfm_handle* h, int code, const std::string& msg);
int ensure(sfm_handle* h, Buf& b, size_t bytes);

// ---- kernel launchers (sfm_kernels.cu); all asynchronous on `st`, return cudaError_t
struct FwdOut {
    float* S;        // [n_rows][kp]   per-row factor sums (train)
    float* mult;     // [n_rows]       dLoss/dyhat (train)
    float* loss;     // [n_rows]       per-row loss (train)
    float* yhat;     // [n_rows]       predictions (predict) or nullptr
    uint32_t* keys;  // [nnz]          feature id of every batch entry (train)
    uint2* pay;      // [nnz]          {batch row, x bits | mult bits} (train)
    int key_bits;    // bits of the feature id inside a sort key
    int blk_shift;   // batch row >> blk_shift = row block, the key's prefix
};
cudaError_t launch_forward(const ModelView& m, const BatchView& b, const FwdOut& o, bool train,
                           int32_t* d_err, int sm_count, cudaStream_t st, int64_t* launches);
// loss / mult -> d_scal[SC_LOSS], [SC_GW0], [SC_COUNT] (fixed-shape fp64 tree, deterministic)
cudaError_t launch_scalar_reduce(const float* loss, const float* mult, int64_t n, double* partials,
                                 unsigned int* ticket, double* d_scal, const int32_t* d_err,
                                 cudaStream_t st, int64_t* launches);
struct UpdateParams {
    float eta, reg0, regw, regv;
};
#ifdef __CUDACC__
// theta - eta * (g / B + lambda * theta)   with the roundings pinned, so every kernel
// that applies the update (fused finalize, dense update, peer-memory update) gives the same bits.
__device__ __forceinline__ float sgd_step(float theta, float g, float inv, float eta, float reg) {
    return __fmaf_rn(-eta, __fmaf_rn(g, inv, __fmul_rn(reg, theta)), theta);
}
#endif
// reduce-by-feature over the sorted entries.  fused: apply the SGD update in place (one GPU);
// else write the dense gradient grad = [gV n_slots*kp | gw n_slots | gw0].
cudaError_t launch_pull(const ModelView& m, int32_t* seg, int key_bits, int n_blocks,
                        const uint32_t* keys, const uint2* pay, int64_t nnz, bool binary,
                        const float* S,
                        const float* mult, float* scratch, const double* d_scal,
                        const int32_t* d_err, UpdateParams up, bool fused, float* grad,
                        int sm_count, cudaStream_t st, int64_t* launches);
// a slice of the reduce: chunk range of the sorted entries + feature range of the finalize
struct PullSlice {
    int64_t chunk_lo = -1, chunk_hi = -1, feat_lo = -1, feat_hi = -1;   // -1 = everything
    bool first = true;                                                   // zero the run bounds
    float *gv = nullptr, *gw = nullptr, *gw0 = nullptr;   // optional pre-biased gradient pointers
};
cudaError_t launch_update_ptrs(const ModelView& m, const float* gv, const float* gw,
                               const float* gw0, const double* d_scal, const int32_t* d_err,
                               UpdateParams up, int64_t feat_lo, int64_t feat_hi, cudaStream_t st,
                               int64_t* launches);
cudaError_t launch_pull_slice(const ModelView& m, int32_t* seg, int key_bits, int n_blocks,
                              const uint32_t* keys, const uint2* pay, int64_t nnz, bool binary,
                              const float* S, const float* mult, float* scratch,
                              const double* d_scal, const int32_t* d_err, UpdateParams up,
                              bool fused, float* grad, int sm_count, cudaStream_t st,
                              const PullSlice& sl, int64_t* launches);
int64_t pull_chunk_entries(const ModelView& m);
cudaError_t launch_slice_bounds(const uint32_t* keys, int64_t nnz, int n_slices, int64_t n_slots,
                                int32_t* d_pos, cudaStream_t st, int64_t* launches);
cudaError_t launch_update_range(const ModelView& m, const float* grad, const double* d_scal,
                                const int32_t* d_err, UpdateParams up, int64_t feat_lo,
                                int64_t feat_hi, cudaStream_t st, int64_t* launches);
size_t pull_scratch_bytes(const ModelView& m, int64_t nnz, int n_blocks);
void pull_plan(const ModelView& m, int64_t n_rows, int* blk_shift, int* n_blocks);
// dense update from an (all-reduced) gradient buffer
cudaError_t launch_update(const ModelView& m, const float* grad, const double* d_scal,
                          const int32_t* d_err, UpdateParams up, cudaStream_t st,
                          int64_t* launches);
// per-row evaluation sums {sum (y-yhat)^2, sum (y-yhat), #sign agree, sum logloss} -> out[4]
cudaError_t launch_metrics(const float* yhat, const float* label, int64_t n, double* partials,
                           double* out4, cudaStream_t st, int64_t* launches);
// writes the (key, payload) entry list of a batch without running the model (partition caches)
cudaError_t launch_emit(const BatchView& b, int key_bits, int blk_shift, int64_t n_slots,
                        uint32_t* keys, uint2* pay, int sm_count, cudaStream_t st,
                        int64_t* launches);
cudaError_t launch_row_lens(const int64_t* row_ptr, const int32_t* row_ids, int64_t n,
                            int64_t* lens, cudaStream_t st, int64_t* launches);
cudaError_t launch_idx_range(const int32_t* idx, int64_t nnz, int32_t* d_minmax, cudaStream_t st,
                             int64_t* launches);
// bit-packed one-hot batch -> idx[n_entries] (ids >= n_slots: *bad = 1, id 0), label[n_rows]
cudaError_t launch_unpack_onehot(const uint32_t* packed, const uint32_t* label_bits,
                                 int64_t n_entries, int64_t n_rows, int id_bits, int64_t n_slots,
                                 int32_t* idx, float* label, int32_t* bad, cudaStream_t st,
                                 int64_t* launches);
cudaError_t launch_pad_v(const float* src, float* dst, int64_t n_slots, int k, int kp, bool unpad,
                         cudaStream_t st, int64_t* launches);
cudaError_t launch_synth_ctr(int64_t n_rows, int64_t row_off, int n_fields,
                             const int32_t* d_log2card, const uint32_t* d_cdf,
                             const int64_t* d_cdf_off, uint64_t seed, int64_t n_slots,
                             int32_t* idx, float* label, int64_t* row_ptr, cudaStream_t st,
                             int64_t* launches);

// ---- bucket-form transposition + reduce (sfm_bucket.cu)
// false: not applicable (SFM_BUCKET=0 / SFM_SORT=cub, more than 2^11 buckets needed, batch too large)
bool bucket_geometry(const ModelView& m, int key_bits, int64_t n_rows, int64_t nnz, BucketGeom* g);
// the bucket geometry of this model gives buckets of whole 32-feature bitmap words (depends on
// the model only, so every rank answers the same): the sparse gradient exchange can be used
bool bucket_sparse_capable(const ModelView& m, int key_bits);
size_t bucket_tables_bytes(const BucketGeom& g);
size_t bucket_work_bytes(const ModelView& m, const BucketGeom& g, int sm_count);
// entries (keys[i], payload) in row order -> packed (local id << (32-LB) | batch row) grouped by
// bucket, rows ascending inside a bucket (+ vals: x bits, non-binary data); fills `tables`.
// pay: nullptr with implicit_div = m (batch row of entry i = i / m), uint32 rows (binary data) or
// uint2 {row, x bits}.
cudaError_t bucket_transpose(const ModelView& m, const BatchView& b, const BucketGeom& g,
                             const uint32_t* keys, const uint2* pay, int implicit_div, void* work,
                             void* tables, uint32_t* packed, uint32_t* vals, unsigned int* ticket,
                             int sm_count, cudaStream_t st, int64_t* launches);
// the two halves of bucket_transpose: counts + plan of a batch of resident all-ones rows taken from
// the row numbers (model-independent: runs one step ahead on the copy stream; the batch size is
// read on the device), and the stable partition alone
cudaError_t bucket_count_plan_rows(const ModelView& m, const BucketGeom& g, const int32_t* idx,
                                   const int32_t* row_ids, const int32_t* n_rows_dev, int n_rows_cap,
                                   int uniform_m, void* work, void* tables, unsigned int* ticket, int sm_count,
                                   cudaStream_t st, int64_t* launches);
cudaError_t bucket_scatter(const ModelView& m, const BatchView& b, const BucketGeom& g,
                           const uint32_t* keys, const uint2* pay, int implicit_div, void* work,
                           const void* tables, uint32_t* packed, uint32_t* vals, int sm_count,
                           cudaStream_t st, int64_t* launches);
// reduce-by-feature over the bucketed entries + SGD update (fused) or dense gradient
cudaError_t bucket_pull(const ModelView& m, const BucketGeom& g, const uint32_t* packed,
                        const uint32_t* vals, const void* tables, void* work, const float* S,
                        const float* mult, const double* d_scal, const int32_t* d_err,
                        UpdateParams up, bool fused, float* grad, uint32_t* touch_bits,
                        bool work_zeroed, int sm_count, cudaStream_t st, int64_t* launches);

// ---- CUB wrappers (sfm_sort.cu)
size_t sort_pairs_temp_bytes(int64_t n, int end_bit);
// sorts (keys_in, pay_in) -> (keys_out, pay_out) by bits [0, end_bit), stable
cudaError_t sort_pairs(void* tmp, size_t tmp_bytes, const uint32_t* keys_in, uint32_t* keys_out,
                       const uint2* pay_in, uint2* pay_out, int64_t n, int end_bit,
                       cudaStream_t st, int64_t* launches);
size_t sort_pairs32_temp_bytes(int64_t n, int end_bit);
// implicit_div = m > 0 (only when radix_usable): val_in may be null, value of position i is i / m
cudaError_t sort_pairs32(void* tmp, size_t tmp_bytes, const uint32_t* keys_in, uint32_t* keys_out,
                         const uint32_t* val_in, uint32_t* val_out, int64_t n, int end_bit,
                         cudaStream_t st, int64_t* launches, int implicit_div = 0);
// wide-digit radix sort written for this path (sfm_radix.cu); same contract as sort_pairs*
bool radix_usable(int64_t n, int end_bit);   // false: SFM_SORT=cub or outside its limits
size_t radix_temp_bytes(int64_t n, int end_bit, int pay_bytes);
// implicit_div = m > 0: pay_in is not read, the payload of input position i is i / m (m < 512)
cudaError_t radix_sort_pairs32(void* tmp, size_t tmp_bytes, const uint32_t* keys_in,
                               uint32_t* keys_out, const uint32_t* pay_in, uint32_t* pay_out,
                               int64_t n, int end_bit, int implicit_div, cudaStream_t st,
                               int64_t* launches);
cudaError_t radix_sort_pairs64(void* tmp, size_t tmp_bytes, const uint32_t* keys_in,
                               uint32_t* keys_out, const uint2* pay_in, uint2* pay_out, int64_t n,
                               int end_bit, cudaStream_t st, int64_t* launches);
size_t scan_u32_temp_bytes(int64_t n);
cudaError_t exclusive_scan_u32(void* tmp, size_t tmp_bytes, const uint32_t* in, uint32_t* out,
                               int64_t n, cudaStream_t st, int64_t* launches);
size_t sort_f32_u32_temp_bytes(int64_t n);
cudaError_t sort_f32_u32(void* tmp, size_t tmp_bytes, const float* keys_in, float* keys_out,
                         const uint32_t* val_in, uint32_t* val_out, int64_t n, cudaStream_t st,
                         int64_t* launches);
// sum of (tie-averaged, 1-based) ranks of the positive rows and their count -> out2[0..1]
cudaError_t launch_auc_ranks(const float* sorted_scores, const uint32_t* sorted_rows,
                             const float* label, int64_t n, double* partials, double* out2,
                             cudaStream_t st, int64_t* launches);
cudaError_t launch_iota_u32(uint32_t* p, int64_t n, cudaStream_t st, int64_t* launches);
size_t scan_temp_bytes(int64_t n);
cudaError_t exclusive_scan_i64(void* tmp, size_t tmp_bytes, const int64_t* in, int64_t* out,
                               int64_t n, cudaStream_t st, int64_t* launches);
size_t select_temp_bytes(int64_t n);
cudaError_t partition_rows_device(void* tmp, size_t tmp_bytes, int64_t n, int64_t global_off,
                                  uint64_t key, int64_t n_parts, int64_t part, int32_t* out_rows,
                                  int32_t* d_count, cudaStream_t st, int64_t* launches);
// Bernoulli row sampler of DESIGN.md section 2.5 over local rows [0, n): global id = off + r
cudaError_t sample_rows_device(void* tmp, size_t tmp_bytes, int64_t n, int64_t global_off,
                               uint64_t key, uint64_t thr, int32_t* out_rows, int32_t* d_count,
                               cudaStream_t st, int64_t* launches);

// ---- row-sharded model (sfm_shard.cu)
int shard_forward(sfm_handle* h, const BatchView& b);                 // yhat -> h->b_yhat
int shard_train(sfm_handle* h, const BatchView& b, int64_t iter, int cache_slot);  // -1: no cache
void shard_clear_cache(sfm_handle* h);

// ---- gradient sum + update over NVLink peer memory (sfm_p2p.cu)
int p2p_setup(sfm_handle* h);      // collective; leaves h->p2p null when any rank cannot take part
void p2p_teardown(sfm_handle* h);
float* p2p_grad_buffer(sfm_handle* h);
uint32_t* p2p_touch_bits(sfm_handle* h);   // this rank's touched-feature bitmap (sparse exchange)
const int32_t* p2p_timeout_flag(sfm_handle* h);
int p2p_reduce_update(sfm_handle* h, UpdateParams up, bool sparse);

// ---- ALS sweep of the reference's own trainer (sfm_als.cu)
int als_sweep(sfm_handle* h, const BatchView& all_rows, int32_t flags, double* rmse_out);
int als_residuals(sfm_handle* h, double* out, int64_t n);
void als_free(sfm_handle* h);

// ---- host side (sfm_host.cpp)
uint64_t mix64(uint64_t x);
// elements [first, first + count) of the seeded Gaussian stream
void init_gaussian_f32(float* v, int64_t first, int64_t count, double mean, double stdev,
                       uint64_t seed);

// ---- NCCL via dlopen (sfm_nccl.cpp)
Nccl* nccl_load(std::string* err);
int nccl_unique_id(Nccl* n, uint8_t* id128, std::string* err);
int nccl_init(Nccl* n, void** comm, const uint8_t* id128, int rank, int world, std::string* err);
int nccl_destroy(Nccl* n, void* comm);
int nccl_async_error(Nccl* n, void* comm, std::string* err);   // ncclCommGetAsyncError
int nccl_abort(Nccl* n, void* comm);
int nccl_allreduce_f32(Nccl* n, void* comm, float* buf, size_t count, cudaStream_t st,
                       std::string* err);
int nccl_allreduce_f64(Nccl* n, void* comm, double* buf, size_t count, cudaStream_t st,
                       std::string* err);
int nccl_bcast_f32(Nccl* n, void* comm, float* buf, size_t count, int root, cudaStream_t st,
                   std::string* err);
int nccl_allgather_i32(Nccl* n, void* comm, const int32_t* send, int32_t* recv, size_t count,
                       cudaStream_t st, std::string* err);
int nccl_allgather_f32(Nccl* n, void* comm, const float* send, float* recv, size_t count,
                       cudaStream_t st, std::string* err);
int nccl_alltoallv_4b(Nccl* n, void* comm, int rank, int world, const void* send,
                      const int64_t* send_off, const int64_t* send_cnt, void* recv,
                      const int64_t* recv_off, const int64_t* recv_cnt, int64_t width,
                      cudaStream_t st, std::string* err);
int nccl_group_start(Nccl* n);
int nccl_group_end(Nccl* n);

}  // namespace sfm
