/*
 * fm_oracle.h -- CPU oracle for the SparkFM data-parallel hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker / the CPU baseline, never as the thing shipped or measured as
 * "ours".  The product path (sparkfm_b200/) has no CPU fallback.
 *
 * PARITY UNPINNED.  The reference (edmundhung/SparkFM, Scala on Spark 1.2 + Breeze 0.10)
 * cannot be compiled or run in this image (no JVM/Scala/sbt/Spark), ships no tests, no golden
 * vectors and no fixtures.  So:
 *   - fmo_predict* restates src/main/scala/io/edstud/spark/fm/FMModel.scala:34-63 from the
 *     SOURCE (30 lines) and is checked against builder-authored hand-computed vectors
 *     (tests/golden/predict_kat.json) and an independent numpy restatement
 *     (oracle/fm_numpy.py).  It is pinned by no reference test and no reference output.
 *   - the SGD gradient / updater / sampler / gradient sum do not exist in the reference at
 *     all; the functions below implement the written spec in DESIGN.md
 *     section 2 (derived from BASELINE.json north_star + the forward in FMModel.scala).
 * All arithmetic is IEEE fp64, like the reference's `Double`.
 *
 * Layout convention shared with the product: V is feature-major, v[i*k + f] is factor f of
 * feature i -- the memory order of Breeze's column-major DenseMatrix(k rows, n+1 cols)
 * built at FMModel.scala:19.  n_slots = num_attribute + 1 (FMModel.scala:18-19).
 */
#ifndef FM_ORACLE_H
#define FM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FMO_TASK_REGRESSION 0     /* Task.Regression     (Task.scala:5) */
#define FMO_TASK_CLASSIFICATION 1 /* Task.Classification (Task.scala:5) */

typedef struct {
    int32_t task;    /* FMO_TASK_* */
    int32_t k;       /* num_factor            FMModel.scala:11 */
    int32_t k0;      /* use global bias       FMModel.scala:25 */
    int32_t k1;      /* use linear term       FMModel.scala:26 */
    int64_t n_slots; /* num_attribute + 1     FMModel.scala:18 */
    double reg0;     /* L2 on w0              FMModel.scala:29 */
    double regw;     /* L2 on w               FMModel.scala:30 */
    double regv;     /* L2 on V               FMModel.scala:31 */
} fmo_params;

/* splitmix64 finaliser: the one integer hash every seeded choice in this repo is built on. */
uint64_t fmo_mix64(uint64_t x);

/* FMModel.predict, faithful structure: k passes over the row's stored entries, left folds in
 * stored order, explicit zeros and duplicate indices included (FMModel.scala:34-63). */
double fmo_predict_row(const fmo_params* p, double w0, const double* w, const double* v,
                       const int32_t* idx, const double* val, int64_t nnz);

/* Batched faithful predict over CSR rows; out[r] for r in [0, n_rows). */
void fmo_predict(const fmo_params* p, double w0, const double* w, const double* v,
                 const int64_t* row_ptr, const int32_t* idx, const double* val,
                 int64_t n_rows, double* out);

/* Same numbers as fmo_predict up to fp64 rounding, one pass over the row (the "optimised CPU"
 * variant of BASELINE.md section 4); OpenMP over rows when n_threads > 1. */
void fmo_predict_fast(const fmo_params* p, double w0, const double* w, const double* v,
                      const int64_t* row_ptr, const int32_t* idx, const double* val,
                      int64_t n_rows, double* out, int n_threads);

/* Per-sample loss and gradient multiplier:
 *   regression:      mult = yhat - y,                     loss = (yhat - y)^2
 *   classification:  y = label > 0 ? +1 : -1, m = y*yhat, mult = -y * (1 - 1/(1+exp(-m))),
 *                    loss = log(1 + exp(-m)) (evaluated stably). */
void fmo_loss_mult(int32_t task, double yhat, double label, double* loss, double* mult);

/* One SGD step over the rows row_ids[0..n_ids) of a CSR dataset,
 * in list order.  grad is caller scratch of n_slots*(k+1)+1 doubles laid out [V | w | w0].
 * The update uses eta = step_size / sqrt(iter) (iter is 1-based) and divides the gradient by
 * batch_count (the GLOBAL batch size; pass n_ids on one node).  If batch_count == 0 nothing
 * is updated.  Returns the SUM of per-sample losses (not the mean). */
double fmo_train_step(const fmo_params* p, double* w0, double* w, double* v,
                      const int64_t* row_ptr, const int32_t* idx, const double* val,
                      const double* label, const int64_t* row_ids, int64_t n_ids,
                      int64_t iter, double step_size, int64_t batch_count, double* grad);

/* Gradient only: fills grad (zeroed here) and returns the loss sum; no update. */
double fmo_gradient(const fmo_params* p, double w0, const double* w, const double* v,
                    const int64_t* row_ptr, const int32_t* idx, const double* val,
                    const double* label, const int64_t* row_ids, int64_t n_ids, double* grad);

/* Updater only: theta <- theta - eta*(g/batch_count + lambda*theta) on every slot. */
void fmo_update(const fmo_params* p, double* w0, double* w, double* v, const double* grad,
                int64_t iter, double step_size, int64_t batch_count);

/* Multi-threaded SGD step for the CPU baseline: rows are split into n_threads contiguous
 * chunks, each thread sums into a private dense gradient, the partials are added in thread
 * order (the treeAggregate analogue), then the update runs over all slots.  scratch holds
 * n_threads * (n_slots*(k+1)+1) doubles.  Same result as fmo_train_step up to fp64 summation
 * order. */
double fmo_train_step_mt(const fmo_params* p, double* w0, double* w, double* v,
                         const int64_t* row_ptr, const int32_t* idx, const double* val,
                         const double* label, const int64_t* row_ids, int64_t n_ids,
                         int64_t iter, double step_size, int64_t batch_count,
                         double* scratch, int n_threads);

/* CPU baseline variants timed by bench.py; never used as checkers.
 * (i) faithful structure: fp64, k passes over the row per forward (FMModel.scala:48-51). */
double fmo_train_step_faithful_mt(const fmo_params* p, double* w0, double* w, double* v,
                                  const int64_t* row_ptr, const int32_t* idx, const double* val,
                                  const double* label, const int64_t* row_ids, int64_t n_ids,
                                  int64_t iter, double step_size, int64_t batch_count,
                                  double* scratch, int n_threads);
/* (ii) tuned: single pass, fp32 model and gradients, touched-only zeroing / reduction. */
typedef struct fmo_fast32 fmo_fast32;
fmo_fast32* fmo_fast32_create(const fmo_params* p, int n_threads);
void fmo_fast32_destroy(fmo_fast32* f);
void fmo_fast32_set_model(fmo_fast32* f, double w0, const double* w, const double* v);
void fmo_fast32_get_model(const fmo_fast32* f, double* w0, double* w, double* v);
double fmo_fast32_train_step(fmo_fast32* f, const int64_t* row_ptr, const int32_t* idx,
                             const float* val, const float* label, const int64_t* row_ids,
                             int64_t n_ids, int64_t iter, double step_size, int64_t batch_count);

/* Mini-batch sampler: i.i.d. Bernoulli(p) per row of the GLOBAL data set,
 * p = floor(fraction * 2^53) / 2^53, decided 64 rows (one aligned block q = r >> 6) at a time:
 * word i = mix64(key + (64 q + i - 1) * 0x9E3779B97F4A7C15) with key = mix64(seed + iter), i.e.
 * output number 64 q + i - 1 of SplitMix64 seeded with key, holds binary digit i of the uniform
 * variate of each of the 64 rows; a row is selected iff its variate is below p (first differing
 * digit).  fraction >= 1 selects every row.  Writes the selected global row ids of
 * [row_lo, row_hi) in ascending order to out (capacity row_hi-row_lo) and returns how many. */
int64_t fmo_sample_rows(uint64_t seed, int64_t iter, double fraction, int64_t row_lo,
                        int64_t row_hi, int64_t* out);

/* PARTITION sampler: row r belongs to mini-batch
 * (mix64(mix64(seed) ^ mix64(r)) >> 11) % n_parts; iteration t uses mini-batch (t-1) % n_parts. */
int64_t fmo_partition_rows(uint64_t seed, int64_t n_parts, int64_t part, int64_t row_lo,
                           int64_t row_hi, int64_t* out);

/* Seeded N(mean, stdev^2) initialisation of V (the reference's init at
 * FMModel.scala:19-22 ignores its seed, so this is a documented replacement): element e gets
 * Box-Muller of u1 = ((mix64(s+2e) >> 11) + 1) * 2^-53, u2 = (mix64(s+2e+1) >> 11) * 2^-53
 * with s = mix64(seed); rounded to fp32 then widened, so host fp32 copies are identical. */
void fmo_init_v(double* v, int64_t count, double mean, double stdev, uint64_t seed);

/* ALS.learn (fm/lib/ALS.scala:15-75), one call = one sweep over w0, w and every factor of V --
 * the only trainer the reference ships.  Pinned by the reference SOURCE (no reference test or
 * output exists).  Restated literally, fp64, driver-serial like the original:
 *   e_r = predict(x_r) - y_r                                        (:17, :142-144)
 *   w0*  = -(sum e - w0*N) / (reg0 + N); e += w0* - w0              (:19-28, :152-154, :167-176)
 *   for id ascending, column h = {x_ri}:  theta* = -(sum e*h - theta*sum h^2)/(reg + sum h^2);
 *       accepted iff finite and != theta (:178-180); then e_r += h_r (theta* - theta)   (:36-43)
 *   for f, q_r = sum_i v_if x_ri (:146-150); for id ascending: h_r = x_ri q_r - x_ri^2 v_if,
 *       same theta*, e update, then q_r += x_ri (v* - v)                                (:45-70)
 * Columns are the transposed input (DataSet.scala:31-38): rows ascending; a column without
 * stored entries is skipped (`features.contains(id)`).  A row must not store the same feature
 * twice (the transposition would carry a duplicate index; the CUDA path rejects it).
 * flags: FMO_ALS_REF_QUIRKS reproduces the one behaviour of the reference that is a bug:
 *   `for (id <- 0 until fm.num_attribute)` (:38, :52) never trains the last slot id = n_slots - 1.
 *   (The residual correction after the w0 step, :24, is NOT a second quirk: it is a lazy RDD
 *   closure over the mutable model that evaluates to e + (w0* - w0*) = e + 0, but the same lazy
 *   re-evaluation, first materialised at :31 after `fm.w0 = w0` at :27, runs fm.predict with
 *   the new w0 (:142-144) -- the residuals come out shifted, exactly as the plain algorithm has
 *   them.  Round 1 modelled this wrongly as "e stays stale".)
 *   FMO_ALS_STORE_F32: every accepted parameter is rounded to fp32 before it is stored and
 *   before the residual update uses it (what the fp32 device model does), so the CUDA path can
 *   be compared tightly.
 * e_out (n_rows doubles, may be NULL) receives the residuals after the sweep.
 * Returns sqrt(mean e^2) after the sweep, or -1 on allocation failure / duplicate entries. */
#define FMO_ALS_REF_QUIRKS 1
#define FMO_ALS_STORE_F32 2
double fmo_als_sweep(const fmo_params* p, double* w0, double* w, double* v,
                     const int64_t* row_ptr, const int32_t* idx, const double* val,
                     const double* label, int64_t n_rows, int32_t flags, double* e_out);

int fmo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
