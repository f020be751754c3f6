/*
 * sparkfm_b200.h -- C ABI of libsparkfm_b200.so: the B200-native (sm_100a) replacement for
 * SparkFM's data-parallel hot path (FM predict + mini-batch SGD training).
 *
 * This is the drop-in boundary: what a JVM host (Scala `SGD extends FMLearn`, `FMWithSGD.train`,
 * `FMModel.predict`) binds through Panama / JNI.  Plain pointers and sizes only.  The reference
 * (edmundhung/SparkFM) has no native interface of its own, so every entry point cites the Scala
 * member it stands in for; paths are relative to src/main/scala/io/edstud/spark/ in the
 * reference tree.  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - every function returns SFM_OK (0) or a negative sfm_status; sfm_last_error(h) gives the
 *     message of the last failure on that handle (library-owned, valid until the next call).
 *   - one handle drives ONE GPU (one process per GPU under multi-GPU; ranks are joined with
 *     sfm_comm_init).  Calls on a handle must be serialised by the caller.
 *   - all host buffers are caller-owned and borrowed for the duration of the call only.
 *   - the model is fp32 on the device: w0, w[n_slots], V[n_slots][k] feature-major -- the memory
 *     order of the reference's column-major DenseMatrix(k, n+1) (fm/FMModel.scala:19).
 *     n_slots = num_attribute + 1 = max feature index + 1 (fm/FMModel.scala:18, DataSet.scala:27-29).
 *   - CSR: row_ptr int64[n_rows+1], idx int32[nnz], val float[nnz] (NULL = every value is 1.0f),
 *     label float[n_rows].  Entries keep their stored order; duplicate indices and explicit
 *     zeros are legal (Breeze activeIterator semantics, fm/FMModel.scala:45,58).  An index
 *     outside [0, n_slots) is a reported error (SFM_ERR_INDEX), never undefined behaviour.
 *   - there is NO CPU fallback: without a CUDA device sfm_create fails with SFM_ERR_CUDA.
 */
#ifndef SPARKFM_B200_H
#define SPARKFM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFM_ABI_VERSION 1

typedef enum {
    SFM_OK = 0,
    SFM_ERR_ARG = -1,    /* bad argument / bad state                                   */
    SFM_ERR_CUDA = -2,   /* CUDA runtime failure (incl. no device)                     */
    SFM_ERR_NCCL = -3,   /* NCCL failure                                               */
    SFM_ERR_OOM = -4,    /* host or device allocation failed                           */
    SFM_ERR_INDEX = -5,  /* feature index outside [0, n_slots) or malformed CSR        */
    SFM_ERR_IO = -6,     /* file / text parse error                                    */
    SFM_ERR_STATE = -7   /* call not valid now (no dataset loaded, comm not initialised) */
} sfm_status;

/* Mini-batch samplers (neither exists in the reference).
 * BERNOULLI: every iteration draws its own Bernoulli(mini_batch_fraction) sample (MLlib's
 *   `sample(false, fraction, 42 + i)` semantics); the batch is transposed (sorted by feature)
 *   inside every iteration.
 * PARTITION: the rows are split ONCE into P = round(1 / mini_batch_fraction) disjoint random
 *   mini-batches, iteration t uses batch (t - 1) mod P (epoch-wise sampling without
 *   replacement).  The transposition of each batch is built at its first use and kept resident,
 *   like the reference's cached `transposeInput` (DataSet.scala:48), so steady-state iterations
 *   do no sorting.  mini_batch_fraction >= 1 is full-batch gradient descent either way. */
#define SFM_SAMPLER_BERNOULLI 0
#define SFM_SAMPLER_PARTITION 1
/* OR-ed into sfm_config.sampler_mode: row-shard V and w over the ranks of the communicator
 * (BASELINE config 4: "V row-sharded across GPUs with all-to-all row gather").  Rank g owns the
 * features [g*ceil(n_slots/G), ...); the parameters are allocated by sfm_comm_init, so the model
 * calls (init / set / get / save) and every predict / train call come after it and are
 * COLLECTIVE: every rank makes them, each with its own rows.  sfm_gradient is not available. */
#define SFM_FLAG_SHARD_V 0x100

#define SFM_TASK_REGRESSION 0     /* Task.Regression      (Task.scala:5) */
#define SFM_TASK_CLASSIFICATION 1 /* Task.Classification  (Task.scala:5) */

/* Model + learner configuration.  Replaces the constructor arguments and vars of
 * fm/FMModel.scala:9-31 (num_attribute, num_factor, k0, k1, reg0/regw/regv) and the `task`
 * of fm/FM.scala:25-30 (ignored by the reference, fm/impl/FactorizationMachines.scala:12;
 * here it selects the loss). */
typedef struct sfm_config {
    int32_t abi_version;       /* must be SFM_ABI_VERSION                                   */
    int32_t task;              /* SFM_TASK_*                                                */
    int32_t k;                 /* num_factor, 0..128                    FMModel.scala:11    */
    int32_t k0;                /* use global bias                       FMModel.scala:25    */
    int32_t k1;                /* use linear term                       FMModel.scala:26    */
    int32_t device;            /* CUDA device ordinal                                       */
    int64_t n_slots;           /* num_attribute + 1                     FMModel.scala:18    */
    float reg0;                /* L2 on w0                              FMModel.scala:29    */
    float regw;                /* L2 on w                               FMModel.scala:30    */
    float regv;                /* L2 on V                               FMModel.scala:31    */
    float step_size;           /* SGD step; eta_t = step_size / sqrt(t)                     */
    float mini_batch_fraction; /* row-sampling rate per iteration, (0, 1]                   */
    int32_t sampler_mode;      /* SFM_SAMPLER_*                                             */
    uint64_t sampler_seed;     /* iteration t samples with seed sampler_seed + t (MLlib: 42) */
} sfm_config;

typedef struct sfm_handle sfm_handle;

/* Counters and phase timers accumulated since sfm_create / sfm_stats_reset. */
typedef struct sfm_stats {
    int64_t train_steps;
    int64_t train_rows;       /* samples that went through the gradient                     */
    int64_t train_nnz;
    int64_t predict_rows;
    int64_t predict_nnz;
    int64_t kernel_launches;  /* launches of THIS library's kernels (incl. its sort passes)  */
    int64_t h2d_bytes;
    int64_t d2h_bytes;
    double ms_forward;        /* CUDA-event time per phase, only while timing is enabled     */
    double ms_sort;
    double ms_reduce;         /* reduce-by-feature (+ fused update on one GPU)               */
    double ms_allreduce;
    double ms_update;
    double ms_predict;
    double ms_total_train;
} sfm_stats;

/* ---------------------------------------------------------------- library ------------- */
int32_t sfm_abi_version(void);
const char* sfm_status_string(int32_t status);
/* Number of visible CUDA devices (0 when there is none or no driver). */
int32_t sfm_device_count(void);

/* Pinned host memory for batch buffers (a JVM wraps it as a direct ByteBuffer / MemorySegment). */
int32_t sfm_host_alloc(void** ptr, uint64_t bytes);
int32_t sfm_host_free(void* ptr);

/* ---------------------------------------------------------------- handle -------------- */
/* new FMModel(num_attribute, num_factor)  (fm/FMModel.scala:9-31): allocates w0 = 0, w = 0,
 * V = 0 on cfg->device.  Call sfm_init_model or sfm_set_model next. */
int32_t sfm_create(const sfm_config* cfg, sfm_handle** out);
int32_t sfm_destroy(sfm_handle* h);
const char* sfm_last_error(const sfm_handle* h);
int32_t sfm_get_config(const sfm_handle* h, sfm_config* out);
/* Mutable learner hyper-parameters (reg0/regw/regv are `var`s at fm/FMModel.scala:29-31). */
int32_t sfm_set_hyper(sfm_handle* h, float reg0, float regw, float regv, float step_size,
                      float mini_batch_fraction);

/* ---------------------------------------------------------------- model --------------- */
/* V ~ N(mean, stdev^2) i.i.d., w = 0, w0 = 0 -- the initial state of fm/FMModel.scala:17-22
 * (defaults mean 0, stdev 0.01 at :12-13).  The reference ignores its `seed` (:14); here the
 * draw is a documented counter-based Box-Muller, reproducible. */
int32_t sfm_init_model(sfm_handle* h, double mean, double stdev, uint64_t seed);
/* Import / export the parameters (w may be NULL when k1 == 0, v when k == 0). */
int32_t sfm_set_model(sfm_handle* h, float w0, const float* w, const float* v);
int32_t sfm_get_model(sfm_handle* h, float* w0, float* w, float* v);
/* Same with the reference's element type (Double, fm/FMModel.scala:17-19). */
int32_t sfm_set_model_f64(sfm_handle* h, double w0, const double* w, const double* v);
int32_t sfm_get_model_f64(sfm_handle* h, double* w0, double* w, double* v);
/* Model persistence (absent from the reference; format in DESIGN.md section 6). */
int32_t sfm_save(sfm_handle* h, const char* path);
int32_t sfm_load(const char* path, int32_t device, sfm_handle** out);

/* ---------------------------------------------------------------- scorer -------------- */
/* Batched FMModel.predict (fm/FMModel.scala:34-63) over host CSR rows -> out[n_rows].
 * Replaces `dataset.rdd.mapValues(predict)` (Model.scala:14,22,29; fm/lib/ALS.scala:143). */
int32_t sfm_predict(sfm_handle* h, const int64_t* row_ptr, const int32_t* idx, const float* val,
                    int64_t n_rows, float* out);

/* ---------------------------------------------------------------- resident data set ---- */
/* DataSet.cache() (DataSet.scala:50-54, called at fm/impl/FactorizationMachines.scala:36):
 * copies this rank's rows to the device once.  global_row_offset is the index of row 0 of this
 * shard in the whole data set (0 on one GPU); the sampler hashes GLOBAL row numbers so that the
 * union of all ranks' batches equals the single-GPU batch. */
int32_t sfm_load_dataset(sfm_handle* h, const int64_t* row_ptr, const int32_t* idx,
                         const float* val, const float* label, int64_t n_rows,
                         int64_t global_row_offset);
int32_t sfm_unload_dataset(sfm_handle* h); /* DataSet.unpersist, DataSet.scala:56-60 */
/* Fills the resident data set ON THE DEVICE with the Criteo/Avazu-shaped synthetic CTR rows of
 * DESIGN.md section 5 (n_fields one-hot ids per row, Zipf within field, hashed into n_slots;
 * rows [global_row_offset, +n_rows)).  Bit-identical to sparkfm_b200.synth.ctr_rows. */
int32_t sfm_synth_ctr_dataset(sfm_handle* h, int64_t n_rows, int64_t global_row_offset,
                              int32_t n_fields, const int32_t* field_log2_card,
                              const uint32_t* zipf_cdf, const int64_t* zipf_cdf_off,
                              uint64_t seed);
/* Copies resident rows [row_lo, row_hi) back as CSR (val may be NULL). */
int32_t sfm_get_dataset_rows(sfm_handle* h, int64_t row_lo, int64_t row_hi, int64_t* row_ptr,
                             int32_t* idx, float* val, float* label);
int32_t sfm_dataset_info(sfm_handle* h, int64_t* n_rows, int64_t* nnz, int32_t* max_index);
/* Scores resident rows [row_lo, row_hi) -> out (host). */
int32_t sfm_predict_resident(sfm_handle* h, int64_t row_lo, int64_t row_hi, float* out);
/* Model.computeRMSE / computeMAE / computeAccuracy (Model.scala:13-30) over the resident rows of
 * ALL ranks, fused into one scoring pass: metrics[0] = sqrt(sum (y-yhat)^2 / N) (:14-15),
 * [1] = sum (y-yhat) / N (the reference's "MAE" has no abs, :22), [2] = fraction with
 * sign(y) == sign(yhat), >= 0 counted positive (:29, without its integer division),
 * [3] = mean logistic loss, [4] = N. */
int32_t sfm_evaluate(sfm_handle* h, double metrics[5]);
/* Area under the ROC curve of the model on THIS rank's resident rows (label > 0 = positive): the
 * predictions are radix-sorted on the device and the rank-sum (Mann-Whitney) statistic is taken
 * with average ranks for tied scores.  out[0] = AUC (NaN if one class is empty), out[1] = number
 * of positives, out[2] = number of negatives.  (The reference has no AUC; listed under "next" in
 * SURVEY.md section 8f.)  Not collective: with several ranks each one reports its own shard. */
int32_t sfm_evaluate_auc(sfm_handle* h, double out[3]);

/* ---------------------------------------------------------------- learner -------------- */
/* One call of the learner plugin, `FMLearn.learn(fm, dataset)` (fm/FMLearn.scala:12), for an
 * SGD learner (FMGradient + FMUpdater + gradient sum of BASELINE.json north_star; none of them
 * exists in the reference -- semantics in DESIGN.md section 2): iteration `iter` (1-based) on
 * the given rows of the RESIDENT data set.  row_ids are shard-local, int64, any order (the
 * order fixes the summation order); NULL + n_ids < 0 means "sample with the built-in sampler
 * at mini_batch_fraction".  mean_loss_out receives the mean per-sample loss over the GLOBAL
 * batch before the update (the analogue of the per-iteration RMSE log,
 * fm/impl/FactorizationMachines.scala:43); batch_out (may be NULL) the global batch size. */
int32_t sfm_train_step(sfm_handle* h, const int64_t* row_ids, int64_t n_ids, int64_t iter,
                       double* mean_loss_out, int64_t* batch_out);
/* Same, on a mini-batch the host packed into CSR buffers (the north_star boundary: "the Scala
 * host packs each mini-batch into CSR buffers (indices, values, labels)").  Copies host->device
 * inside the call. */
int32_t sfm_train_step_csr(sfm_handle* h, const int64_t* row_ptr, const int32_t* idx,
                           const float* val, const float* label, int64_t n_rows, int64_t iter,
                           double* mean_loss_out, int64_t* batch_out);
/* Pipelined form of sfm_train_step_csr for streaming hosts: sfm_stage_csr queues the host->device
 * copy of a mini-batch into staging slot 0 or 1 on the library's copy stream and returns at once
 * (the host buffers must stay unchanged until the sfm_train_step_staged that consumes the slot
 * returns; use sfm_host_alloc'd pinned memory for the copy to be asynchronous);
 * sfm_train_step_staged runs the iteration on a staged slot.  Staging batch t+1 before training on
 * batch t overlaps PCIe with compute. */
int32_t sfm_stage_csr(sfm_handle* h, int32_t slot, const int64_t* row_ptr, const int32_t* idx,
                      const float* val, const float* label, int64_t n_rows);
int32_t sfm_train_step_staged(sfm_handle* h, int32_t slot, int64_t iter, double* mean_loss_out,
                              int64_t* batch_out);
/* Compact staging for uniform all-ones mini-batches (one-hot CTR rows packed by the host from
 * RDD[LabeledPoint] whose rows all hold exactly m entries of value 1): no row_ptr, no values,
 * the n_rows*m feature ids bit-packed at id_bits (1..32) bits each -- entry e occupies bits
 * [e*id_bits, (e+1)*id_bits) of the little-endian uint32 stream packed_idx, which must hold
 * (n_rows*m*id_bits + 31)/32 + 1 words (one word of slack) -- and labels either 1 bit per row in
 * label_bits (bit r%32 of word r/32; set = positive, stored as 1.0f / 0.0f) or fp32 in label_f32
 * (exactly one of the two non-NULL).  A device kernel unpacks into the same staging slot
 * sfm_stage_csr fills, checking every id against n_slots (SFM_ERR_INDEX from the step).  Cuts the
 * host->device bytes of a Criteo-shaped batch (m = 39, 20-bit ids) from 168 to 98 per row.
 * Consumed by sfm_train_step_staged like any other staged slot. */
int32_t sfm_stage_onehot(sfm_handle* h, int32_t slot, const uint32_t* packed_idx,
                         const uint32_t* label_bits, const float* label_f32, int64_t n_rows,
                         int32_t m, int32_t id_bits);
/* Host-side packer for sfm_stage_onehot (multi-threaded; no device needed): idx[n_rows*m] ->
 * packed_idx, label[n_rows] -> label_bits (NULL: skip).  SFM_ERR_INDEX if an id needs more than
 * id_bits bits or is negative. */
int32_t sfm_pack_onehot(const int32_t* idx, const float* label, int64_t n_rows, int32_t m,
                        int32_t id_bits, uint32_t* packed_idx, uint32_t* label_bits);
/* The loop of FM.learnWith (fm/impl/FactorizationMachines.scala:42-46) with the built-in
 * sampler: iterations first_iter .. first_iter + n_iters - 1 on the resident data set, no host
 * round trip in between.  loss_history[n_iters] (may be NULL) gets each iteration's mean loss. */
int32_t sfm_train(sfm_handle* h, int64_t first_iter, int64_t n_iters, double* loss_history);
/* The built-in sampler, host side: global row ids of
 * [row_lo, row_hi) selected in iteration iter -- i.i.d. Bernoulli(floor(fraction * 2^53) / 2^53)
 * per row, counter-based and bit-sliced over aligned blocks of 64 global rows, bit-identical to
 * the device sampler.  out has capacity row_hi - row_lo. */
int32_t sfm_sample_rows(uint64_t seed, int64_t iter, double fraction, int64_t row_lo,
                        int64_t row_hi, int64_t* out, int64_t* n_out);
/* Host twin of the PARTITION sampler: global row ids of [row_lo, row_hi) that belong to
 * mini-batch `part` of `n_parts` (row r is in part (mix64(mix64(seed) ^ mix64(r)) >> 11) % n_parts). */
int32_t sfm_partition_rows(uint64_t seed, int64_t n_parts, int64_t part, int64_t row_lo,
                           int64_t row_hi, int64_t* out, int64_t* n_out);
/* Gradient of the current model on resident rows, no update: grad_v[n_slots][k], grad_w[n_slots],
 * grad_w0 (any may be NULL), loss_sum.  After sfm_comm_init the result is the sum over ranks.
 * For tests and for hosts that run their own updater. */
int32_t sfm_gradient(sfm_handle* h, const int64_t* row_ids, int64_t n_ids, float* grad_v,
                     float* grad_w, float* grad_w0, double* loss_sum, int64_t* batch_out);

/* ---------------------------------------------------------------- ALS ------------------ */
/* ALS.learn(fm, dataset) (fm/lib/ALS.scala:15-75), the trainer the reference ships: ONE sweep of the
 * closed-form coordinate updates over w0, every w_i and every v_if (squared loss on the labels
 * as stored, reg0/regw/regv of the handle) on the resident data set, in the reference's
 * coordinate order, residual cache e and per-factor cache q in fp64 on the device.  The
 * transposed input and the level schedule that makes the sequential sweep parallel are built at
 * the first call and kept (the reference's cached `transposeInput`, DataSet.scala:48).
 * flags: SFM_ALS_REF_QUIRKS reproduces the reference's one bug (`0 until num_attribute`,
 * ALS.scala:38,52, never trains the last slot) and is what a drop-in for ALS.run() passes; 0 trains
 * every slot.  (The residuals after the w0 step are yhat_new - y in both modes: the reference's
 * lazy RDD is re-evaluated with the new w0, ALS.scala:27,31,142-144.)  rmse_out: sqrt(mean e^2) of the residuals after the sweep (the per-iteration
 * train RMSE FactorizationMachines.scala:43 computes with an extra pass).  One GPU, replicated
 * model; rows must not store a feature index twice (SFM_ERR_ARG). */
#define SFM_ALS_REF_QUIRKS 1
int32_t sfm_als_sweep(sfm_handle* h, int32_t flags, double* rmse_out);
/* Residuals e_r = yhat_r - y_r after the last sweep, n = resident rows. */
int32_t sfm_als_residuals(sfm_handle* h, double* out, int64_t n);

/* ---------------------------------------------------------------- multi-GPU ------------ */
/* Replaces Spark's driver-side combination of partition results (`.sum()` Model.scala:14,
 * `.reduce(_+_)` fm/lib/ALS.scala:153; MLlib treeAggregate in north_star) by an NCCL all-reduce
 * of the (V, w, w0) gradient over NVLink.  Rank 0 makes the id, every rank (one process per
 * GPU) passes the same 128 bytes. */
#define SFM_UNIQUE_ID_BYTES 128
int32_t sfm_comm_unique_id(uint8_t id[SFM_UNIQUE_ID_BYTES]);
int32_t sfm_comm_init(sfm_handle* h, const uint8_t id[SFM_UNIQUE_ID_BYTES], int32_t rank,
                      int32_t world_size);
int32_t sfm_comm_info(const sfm_handle* h, int32_t* rank, int32_t* world_size);
/* How gradients are combined after sfm_comm_init: SFM_COMM_NONE (one GPU), SFM_COMM_NCCL (dense
 * ncclAllReduce + update kernel), SFM_COMM_PEER (ONE kernel that sums the ranks' gradients over
 * NVLink peer memory, updates and writes every replica; chosen when every rank can map the
 * others' buffers, SFM_P2P=0 in the environment forces NCCL), SFM_COMM_SHARDED (SFM_FLAG_SHARD_V). */
#define SFM_COMM_NONE 0
#define SFM_COMM_NCCL 1
#define SFM_COMM_PEER 2
#define SFM_COMM_SHARDED 3
int32_t sfm_comm_mode(const sfm_handle* h, int32_t* mode);
/* Copies rank 0's model to every rank. */
int32_t sfm_comm_broadcast_model(sfm_handle* h);

/* ---------------------------------------------------------------- text ingest ---------- */
/* FMUtils.loadLibFMFile (fm/FMUtils.scala:23-53) on an in-memory buffer: trims each line,
 * skips empty and '#' lines, splits on ' ', label = first token, `index:value` tokens with NO
 * index shift (:32), stored order and duplicates preserved.  Two-call protocol: call with
 * idx == NULL to get n_rows / nnz / max_index, allocate, call again.  num_features <= 0 means
 * "infer": then a row without features is an error like the Scala's `indices.max` (:45).
 * dimension_out = num_features > 0 ? num_features : max index  (vector length is +1, :50).
 * err_line (may be NULL) receives the 1-based line of a parse error. */
int32_t sfm_parse_libfm(const char* text, uint64_t len, int32_t num_features, int64_t* n_rows,
                        int64_t* nnz, int32_t* dimension_out, double* label, int64_t* row_ptr,
                        int32_t* idx, double* val, int64_t* err_line);
/* FMUtils.saveAsLibFMFile row formatter (fm/FMUtils.scala:58-74): label and `i+1:value` tokens
 * (the +1 asymmetry with the loader is the reference's, :63), numbers through
 * DecimalFormat("#.###") / "#" (HALF_EVEN).  Returns the needed size in *needed; writes when
 * cap is large enough. */
int32_t sfm_format_libfm(const double* label, const int64_t* row_ptr, const int32_t* idx,
                         const double* val, int64_t n_rows, char* out, uint64_t cap,
                         uint64_t* needed);

/* ---------------------------------------------------------------- stats ---------------- */
int32_t sfm_stats_get(sfm_handle* h, sfm_stats* out);
int32_t sfm_stats_reset(sfm_handle* h);
/* Per-phase CUDA-event timing costs a synchronisation per phase: off by default. */
int32_t sfm_set_phase_timing(sfm_handle* h, int32_t enabled);
/* Blocks until all work queued on the handle has finished. */
int32_t sfm_synchronize(sfm_handle* h);
/* Device time of kernels on the handle's compute stream between the two marks, in ms
 * (CUDA events on the launching stream; torch.cuda.Event cannot see this stream). */
int32_t sfm_timer_start(sfm_handle* h);
int32_t sfm_timer_stop(sfm_handle* h, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* SPARKFM_B200_H */
