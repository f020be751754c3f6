/*
 * fm_oracle.c -- CPU oracle (fp64) for the SparkFM hot path.  See fm_oracle.h: TEST
 * INFRASTRUCTURE ONLY, PARITY UNPINNED (no runnable reference, no reference golden vectors).
 *
 * Reference citations are relative to /root/reference/src/main/scala/io/edstud/spark/.
 */
#include "fm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

uint64_t fmo_mix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

int fmo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* fm/FMModel.scala:57-63 computeFactorComponents: materialise f_j = v(i, idx_j) * x_j over the
 * stored entries, then sum_f = f.reduce(_+_), sum_sqr_f = f.map(x*x).reduce(_+_) -- two left
 * folds in stored order. */
static void factor_components(const double* v, int32_t k, int32_t f, const int32_t* idx,
                              const double* val, int64_t nnz, double* sum_f, double* sum_sqr_f) {
    double s = 0.0, q = 0.0;
    for (int64_t j = 0; j < nnz; ++j) {
        double t = v[(int64_t)idx[j] * k + f] * val[j];
        if (j == 0) { s = t; q = t * t; }      /* reduce starts from the first element */
        else        { s = s + t; q = q + t * t; }
    }
    *sum_f = s;
    *sum_sqr_f = q;
}

/* fm/FMModel.scala:34-55 */
double fmo_predict_row(const fmo_params* p, double w0, const double* w, const double* v,
                       const int32_t* idx, const double* val, int64_t nnz) {
    double result = 0.0;                       /* :36 */
    if (p->k0) result += w0;                   /* :38-40 */
    if (nnz > 0) {                             /* :42 features.used > 0 */
        if (p->k1) {                           /* :44-46 map(w(i)*x).reduce(_+_) */
            double lin = w[idx[0]] * val[0];
            for (int64_t j = 1; j < nnz; ++j) lin = lin + w[idx[j]] * val[j];
            result += lin;
        }
        for (int32_t f = 0; f < p->k; ++f) {   /* :48-51 */
            double s, q;
            factor_components(v, p->k, f, idx, val, nnz, &s, &q);
            result += 0.5 * (s * s - q);
        }
    }
    return result;                             /* :54 */
}

void fmo_predict(const fmo_params* p, double w0, const double* w, const double* v,
                 const int64_t* row_ptr, const int32_t* idx, const double* val,
                 int64_t n_rows, double* out) {
    for (int64_t r = 0; r < n_rows; ++r) {
        int64_t b = row_ptr[r];
        out[r] = fmo_predict_row(p, w0, w, v, idx + b, val + b, row_ptr[r + 1] - b);
    }
}

/* One pass over the row: s_f and q_f for all factors at once.  Returns yhat and leaves the
 * per-factor sums in s[0..k) (needed by the gradient). */
static double forward_row(const fmo_params* p, double w0, const double* w, const double* v,
                          const int32_t* idx, const double* val, int64_t nnz, double* s,
                          double* q) {
    const int32_t k = p->k;
    double result = p->k0 ? w0 : 0.0;
    for (int32_t f = 0; f < k; ++f) { s[f] = 0.0; q[f] = 0.0; }
    if (nnz <= 0) return result;
    double lin = 0.0;
    for (int64_t j = 0; j < nnz; ++j) {
        const double x = val[j];
        const double* vi = v + (int64_t)idx[j] * k;
        lin += w[idx[j]] * x;
        for (int32_t f = 0; f < k; ++f) {
            double t = vi[f] * x;
            s[f] += t;
            q[f] += t * t;
        }
    }
    if (p->k1) result += lin;
    for (int32_t f = 0; f < k; ++f) result += 0.5 * (s[f] * s[f] - q[f]);
    return result;
}

void fmo_predict_fast(const fmo_params* p, double w0, const double* w, const double* v,
                      const int64_t* row_ptr, const int32_t* idx, const double* val,
                      int64_t n_rows, double* out, int n_threads) {
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel num_threads(n_threads)
    {
        double* s = (double*)malloc(sizeof(double) * 2 * (size_t)(p->k > 0 ? p->k : 1));
        double* q = s + (p->k > 0 ? p->k : 1);
#pragma omp for schedule(static)
        for (int64_t r = 0; r < n_rows; ++r) {
            int64_t b = row_ptr[r];
            out[r] = forward_row(p, w0, w, v, idx + b, val + b, row_ptr[r + 1] - b, s, q);
        }
        free(s);
    }
}

void fmo_loss_mult(int32_t task, double yhat, double label, double* loss, double* mult) {
    if (task == FMO_TASK_CLASSIFICATION) {
        const double y = label > 0.0 ? 1.0 : -1.0;
        const double m = y * yhat;
        /* log(1+exp(-m)) and sigmoid(-m) without overflow */
        const double e = exp(-fabs(m));
        *loss = (m > 0.0 ? 0.0 : -m) + log1p(e);
        const double sig_neg_m = m > 0.0 ? e / (1.0 + e) : 1.0 / (1.0 + e); /* 1 - 1/(1+exp(-m)) */
        *mult = -y * sig_neg_m;
    } else {
        const double d = yhat - label;
        *loss = d * d;
        *mult = d;
    }
}

/* Adds one sample's gradient into grad = [V (n_slots*k) | w (n_slots) | w0] and returns its
 * loss.  dV_if = (x_i*s_f - v_if*x_i^2)*mult, dw_i = x_i*mult, dw0 = mult. */
static double accumulate_row(const fmo_params* p, double w0, const double* w, const double* v,
                             const int32_t* idx, const double* val, int64_t nnz, double label,
                             double* grad, double* s, double* q) {
    const int32_t k = p->k;
    const int64_t n = p->n_slots;
    double loss, mult;
    const double yhat = forward_row(p, w0, w, v, idx, val, nnz, s, q);
    fmo_loss_mult(p->task, yhat, label, &loss, &mult);
    if (p->k0) grad[n * k + n] += mult;
    for (int64_t j = 0; j < nnz; ++j) {
        const int64_t i = idx[j];
        const double x = val[j];
        if (p->k1) grad[n * k + i] += x * mult;
        const double* vi = v + i * k;
        double* gi = grad + i * k;
        for (int32_t f = 0; f < k; ++f) gi[f] += (x * s[f] - vi[f] * x * x) * mult;
    }
    return loss;
}

double fmo_gradient(const fmo_params* p, double w0, const double* w, const double* v,
                    const int64_t* row_ptr, const int32_t* idx, const double* val,
                    const double* label, const int64_t* row_ids, int64_t n_ids, double* grad) {
    const int64_t len = p->n_slots * (p->k + 1) + 1;
    memset(grad, 0, sizeof(double) * (size_t)len);
    double* s = (double*)malloc(sizeof(double) * 2 * (size_t)(p->k > 0 ? p->k : 1));
    double* q = s + (p->k > 0 ? p->k : 1);
    double loss_sum = 0.0;
    for (int64_t t = 0; t < n_ids; ++t) {
        const int64_t r = row_ids ? row_ids[t] : t;
        const int64_t b = row_ptr[r];
        loss_sum += accumulate_row(p, w0, w, v, idx + b, val + b, row_ptr[r + 1] - b, label[r],
                                   grad, s, q);
    }
    free(s);
    return loss_sum;
}

void fmo_update(const fmo_params* p, double* w0, double* w, double* v, const double* grad,
                int64_t iter, double step_size, int64_t batch_count) {
    if (batch_count <= 0) return;
    const int32_t k = p->k;
    const int64_t n = p->n_slots;
    const double eta = step_size / sqrt((double)iter);
    const double inv = 1.0 / (double)batch_count;
    for (int64_t e = 0; e < n * k; ++e) v[e] -= eta * (grad[e] * inv + p->regv * v[e]);
    if (p->k1)
        for (int64_t i = 0; i < n; ++i) w[i] -= eta * (grad[n * k + i] * inv + p->regw * w[i]);
    if (p->k0) *w0 -= eta * (grad[n * k + n] * inv + p->reg0 * *w0);
}

double fmo_train_step(const fmo_params* p, double* w0, double* w, double* v,
                      const int64_t* row_ptr, const int32_t* idx, const double* val,
                      const double* label, const int64_t* row_ids, int64_t n_ids,
                      int64_t iter, double step_size, int64_t batch_count, double* grad) {
    const double loss_sum =
        fmo_gradient(p, *w0, w, v, row_ptr, idx, val, label, row_ids, n_ids, grad);
    fmo_update(p, w0, w, v, grad, iter, step_size, batch_count);
    return loss_sum;
}

double fmo_train_step_mt(const fmo_params* p, double* w0, double* w, double* v,
                         const int64_t* row_ptr, const int32_t* idx, const double* val,
                         const double* label, const int64_t* row_ids, int64_t n_ids,
                         int64_t iter, double step_size, int64_t batch_count,
                         double* scratch, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    const int32_t k = p->k;
    const int64_t n = p->n_slots;
    const int64_t len = n * (k + 1) + 1;
    double* losses = (double*)calloc((size_t)n_threads, sizeof(double));
    const double w0v = *w0;
#pragma omp parallel num_threads(n_threads)
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num();
#else
        const int t = 0;
#endif
        double* g = scratch + (int64_t)t * len;
        memset(g, 0, sizeof(double) * (size_t)len);
        double* s = (double*)malloc(sizeof(double) * 2 * (size_t)(k > 0 ? k : 1));
        double* q = s + (k > 0 ? k : 1);
        const int64_t lo = n_ids * t / n_threads, hi = n_ids * (t + 1) / n_threads;
        double ls = 0.0;
        for (int64_t u = lo; u < hi; ++u) {
            const int64_t r = row_ids ? row_ids[u] : u;
            const int64_t b = row_ptr[r];
            ls += accumulate_row(p, w0v, w, v, idx + b, val + b, row_ptr[r + 1] - b, label[r], g,
                                 s, q);
        }
        losses[t] = ls;
        free(s);
#pragma omp barrier
        /* partial sums added in thread order, then the update, both split by slot */
        if (batch_count > 0) {
            const double eta = step_size / sqrt((double)iter);
            const double inv = 1.0 / (double)batch_count;
#pragma omp for schedule(static)
            for (int64_t e = 0; e < len; ++e) {
                double acc = scratch[e];
                for (int u = 1; u < n_threads; ++u) acc += scratch[(int64_t)u * len + e];
                if (e < n * k) v[e] -= eta * (acc * inv + p->regv * v[e]);
                else if (e < n * k + n) {
                    if (p->k1) w[e - n * k] -= eta * (acc * inv + p->regw * w[e - n * k]);
                } else if (p->k0) *w0 -= eta * (acc * inv + p->reg0 * *w0);
            }
        }
    }
    double loss_sum = 0.0;
    for (int t = 0; t < n_threads; ++t) loss_sum += losses[t];
    free(losses);
    return loss_sum;
}

/* ---------------------------------------------------------------------------------------------
 * CPU baseline variants.  Both are timed by bench.py; neither is used as
 * a checker.
 *   (i)  fmo_train_step_faithful_mt: fp64, the forward of every row makes k passes over the row's
 *        stored entries exactly as FMModel.scala:48-51 does (computeFactorComponents per factor);
 *        gradient / reduction / update as fmo_train_step_mt.
 *   (ii) fmo_fast32_*: single pass per row, fp32 parameters and gradients, per-thread gradient
 *        rows zeroed and reduced only where touched (iteration stamps), dense decay pass only when
 *        a regulariser is non-zero: what a tuned CPU implementation of the same spec looks like.
 * ------------------------------------------------------------------------------------------- */
double fmo_train_step_faithful_mt(const fmo_params* p, double* w0, double* w, double* v,
                                  const int64_t* row_ptr, const int32_t* idx, const double* val,
                                  const double* label, const int64_t* row_ids, int64_t n_ids,
                                  int64_t iter, double step_size, int64_t batch_count,
                                  double* scratch, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    const int32_t k = p->k;
    const int64_t n = p->n_slots;
    const int64_t len = n * (k + 1) + 1;
    double* losses = (double*)calloc((size_t)n_threads, sizeof(double));
    const double w0v = *w0;
#pragma omp parallel num_threads(n_threads)
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num();
#else
        const int t = 0;
#endif
        double* g = scratch + (int64_t)t * len;
        memset(g, 0, sizeof(double) * (size_t)len);
        double* s = (double*)malloc(sizeof(double) * (size_t)(k > 0 ? k : 1));
        const int64_t lo = n_ids * t / n_threads, hi = n_ids * (t + 1) / n_threads;
        double ls = 0.0;
        for (int64_t u = lo; u < hi; ++u) {
            const int64_t r = row_ids ? row_ids[u] : u;
            const int64_t b = row_ptr[r], nnz = row_ptr[r + 1] - b;
            const int32_t* ri = idx + b;
            const double* rv = val + b;
            /* FMModel.predict, k passes (FMModel.scala:34-55) */
            double yhat = p->k0 ? w0v : 0.0;
            if (nnz > 0) {
                if (p->k1) {
                    double lin = w[ri[0]] * rv[0];
                    for (int64_t j = 1; j < nnz; ++j) lin = lin + w[ri[j]] * rv[j];
                    yhat += lin;
                }
                for (int32_t f = 0; f < k; ++f) {
                    double sf, qf;
                    factor_components(v, k, f, ri, rv, nnz, &sf, &qf);
                    s[f] = sf;
                    yhat += 0.5 * (sf * sf - qf);
                }
            } else {
                for (int32_t f = 0; f < k; ++f) s[f] = 0.0;
            }
            double loss, mult;
            fmo_loss_mult(p->task, yhat, label[r], &loss, &mult);
            ls += loss;
            if (p->k0) g[n * k + n] += mult;
            for (int64_t j = 0; j < nnz; ++j) {
                const int64_t i = ri[j];
                const double x = rv[j];
                if (p->k1) g[n * k + i] += x * mult;
                const double* vi = v + i * k;
                double* gi = g + i * k;
                for (int32_t f = 0; f < k; ++f) gi[f] += (x * s[f] - vi[f] * x * x) * mult;
            }
        }
        losses[t] = ls;
        free(s);
#pragma omp barrier
        if (batch_count > 0) {
            const double eta = step_size / sqrt((double)iter);
            const double inv = 1.0 / (double)batch_count;
#pragma omp for schedule(static)
            for (int64_t e = 0; e < len; ++e) {
                double acc = scratch[e];
                for (int u = 1; u < n_threads; ++u) acc += scratch[(int64_t)u * len + e];
                if (e < n * k) v[e] -= eta * (acc * inv + p->regv * v[e]);
                else if (e < n * k + n) {
                    if (p->k1) w[e - n * k] -= eta * (acc * inv + p->regw * w[e - n * k]);
                } else if (p->k0) *w0 -= eta * (acc * inv + p->reg0 * *w0);
            }
        }
    }
    double loss_sum = 0.0;
    for (int t = 0; t < n_threads; ++t) loss_sum += losses[t];
    free(losses);
    return loss_sum;
}

struct fmo_fast32 {
    fmo_params p;
    int n_threads;
    float w0;
    float* w;          /* [n_slots] */
    float* v;          /* [n_slots][k] */
    float** g;         /* per thread: [n_slots][k+1] (V row | w) */
    uint32_t** stamp;  /* per thread: [n_slots] iteration stamp of the last touch */
    double* loss;      /* per thread */
    double* gw0;       /* per thread */
    uint32_t tick;
};

fmo_fast32* fmo_fast32_create(const fmo_params* p, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    fmo_fast32* f = (fmo_fast32*)calloc(1, sizeof(fmo_fast32));
    if (!f) return NULL;
    f->p = *p;
    f->n_threads = n_threads;
    const size_t n = (size_t)p->n_slots, k = (size_t)p->k;
    f->w = (float*)calloc(n, sizeof(float));
    f->v = (float*)calloc(n * (k > 0 ? k : 1), sizeof(float));
    f->g = (float**)calloc((size_t)n_threads, sizeof(float*));
    f->stamp = (uint32_t**)calloc((size_t)n_threads, sizeof(uint32_t*));
    f->loss = (double*)calloc((size_t)n_threads, sizeof(double));
    f->gw0 = (double*)calloc((size_t)n_threads, sizeof(double));
    for (int t = 0; t < n_threads; ++t) {
        f->g[t] = (float*)malloc(sizeof(float) * n * (k + 1));
        f->stamp[t] = (uint32_t*)calloc(n, sizeof(uint32_t));
    }
    return f;
}

void fmo_fast32_destroy(fmo_fast32* f) {
    if (!f) return;
    for (int t = 0; t < f->n_threads; ++t) {
        free(f->g[t]);
        free(f->stamp[t]);
    }
    free(f->g); free(f->stamp); free(f->loss); free(f->gw0); free(f->w); free(f->v);
    free(f);
}

void fmo_fast32_set_model(fmo_fast32* f, double w0, const double* w, const double* v) {
    const int64_t n = f->p.n_slots, k = f->p.k;
    f->w0 = (float)w0;
    for (int64_t i = 0; i < n; ++i) f->w[i] = (float)w[i];
    for (int64_t e = 0; e < n * k; ++e) f->v[e] = (float)v[e];
}

void fmo_fast32_get_model(const fmo_fast32* f, double* w0, double* w, double* v) {
    const int64_t n = f->p.n_slots, k = f->p.k;
    *w0 = f->w0;
    for (int64_t i = 0; i < n; ++i) w[i] = f->w[i];
    for (int64_t e = 0; e < n * k; ++e) v[e] = f->v[e];
}

/* val may be NULL (all ones).  Returns the loss sum. */
double fmo_fast32_train_step(fmo_fast32* f, const int64_t* row_ptr, const int32_t* idx,
                             const float* val, const float* label, const int64_t* row_ids,
                             int64_t n_ids, int64_t iter, double step_size, int64_t batch_count) {
    const fmo_params* p = &f->p;
    const int32_t k = p->k;
    const int64_t n = p->n_slots;
    const int T = f->n_threads;
    const uint32_t tick = ++f->tick;
    const float w0v = f->w0;
    const float* w = f->w;
    const float* v = f->v;
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num();
#else
        const int t = 0;
#endif
        float* g = f->g[t];
        uint32_t* st = f->stamp[t];
        float s[128], q[128];
        const int64_t lo = n_ids * t / T, hi = n_ids * (t + 1) / T;
        double ls = 0.0, g0 = 0.0;
        for (int64_t u = lo; u < hi; ++u) {
            const int64_t r = row_ids ? row_ids[u] : u;
            const int64_t b = row_ptr[r], nnz = row_ptr[r + 1] - b;
            const int32_t* ri = idx + b;
            const float* rv = val ? val + b : NULL;
            for (int32_t c = 0; c < k; ++c) { s[c] = 0.f; q[c] = 0.f; }
            float lin = 0.f;
            for (int64_t j = 0; j < nnz; ++j) {
                const float x = rv ? rv[j] : 1.f;
                const float* vi = v + (int64_t)ri[j] * k;
                lin += w[ri[j]] * x;
                for (int32_t c = 0; c < k; ++c) {
                    const float a = vi[c] * x;
                    s[c] += a;
                    q[c] += a * a;
                }
            }
            float yhat = p->k0 ? w0v : 0.f;
            if (nnz > 0) {
                if (p->k1) yhat += lin;
                float pair = 0.f;
                for (int32_t c = 0; c < k; ++c) pair += s[c] * s[c] - q[c];
                yhat += 0.5f * pair;
            }
            double loss, mult;
            fmo_loss_mult(p->task, (double)yhat, (double)label[r], &loss, &mult);
            ls += loss;
            g0 += mult;
            const float mu = (float)mult;
            for (int64_t j = 0; j < nnz; ++j) {
                const int64_t i = ri[j];
                const float x = rv ? rv[j] : 1.f;
                float* gi = g + i * (k + 1);
                if (st[i] != tick) {   /* first touch in this iteration: clear the row */
                    st[i] = tick;
                    for (int32_t c = 0; c <= k; ++c) gi[c] = 0.f;
                }
                const float* vi = v + i * k;
                const float xm = x * mu, xxm = x * x * mu;
                for (int32_t c = 0; c < k; ++c) gi[c] += xm * s[c] - vi[c] * xxm;
                gi[k] += xm;
            }
        }
        f->loss[t] = ls;
        f->gw0[t] = g0;
#pragma omp barrier
        if (batch_count > 0) {
            const float eta = (float)(step_size / sqrt((double)iter));
            const float inv = (float)(1.0 / (double)batch_count);
            const float rv_ = (float)p->regv, rw_ = (float)p->regw;
            const int decay = rv_ != 0.f || rw_ != 0.f;
            float acc[129];
#pragma omp for schedule(static)
            for (int64_t i = 0; i < n; ++i) {
                int touched = 0;
                for (int u = 0; u < T; ++u) {
                    if (f->stamp[u][i] == tick) {
                        const float* gi = f->g[u] + i * (k + 1);
                        if (!touched) { for (int32_t c = 0; c <= k; ++c) acc[c] = gi[c]; touched = 1; }
                        else for (int32_t c = 0; c <= k; ++c) acc[c] += gi[c];
                    }
                }
                float* vi = f->v + i * k;
                if (touched) {
                    for (int32_t c = 0; c < k; ++c) vi[c] -= eta * (acc[c] * inv + rv_ * vi[c]);
                    if (p->k1) f->w[i] -= eta * (acc[k] * inv + rw_ * f->w[i]);
                } else if (decay) {
                    for (int32_t c = 0; c < k; ++c) vi[c] -= eta * rv_ * vi[c];
                    if (p->k1) f->w[i] -= eta * rw_ * f->w[i];
                }
            }
        }
    }
    double loss_sum = 0.0, g0 = 0.0;
    for (int t = 0; t < T; ++t) { loss_sum += f->loss[t]; g0 += f->gw0[t]; }
    if (batch_count > 0 && p->k0) {
        const float eta = (float)(step_size / sqrt((double)iter));
        f->w0 -= eta * ((float)(g0 / (double)batch_count) + (float)p->reg0 * f->w0);
    }
    return loss_sum;
}

/* Hits of global block q (rows 64q .. 64q+63) as a bit mask: DESIGN.md section 2.5. */
static uint64_t fmo_bernoulli_block(uint64_t key, uint64_t thr, int last, uint64_t q) {
    const uint64_t gamma = 0x9E3779B97F4A7C15ULL;
    uint64_t und = ~0ULL, hit = 0ULL, ctr = key + (q << 6) * gamma;
    for (int i = 1; i <= last && und; ++i) {
        const uint64_t w = fmo_mix64(ctr); /* SplitMix64(seed = key), output number 64q + i - 1 */
        ctr += gamma;
        if ((thr >> (53 - i)) & 1ULL) { /* digit i of p is 1: rows whose digit is 0 are below p */
            hit |= und & ~w;
            und &= w;
        } else { /* digit i of p is 0: rows whose digit is 1 are above p */
            und &= ~w;
        }
    }
    return hit;
}

int64_t fmo_sample_rows(uint64_t seed, int64_t iter, double fraction, int64_t row_lo,
                        int64_t row_hi, int64_t* out) {
    int64_t n = 0;
    if (fraction >= 1.0) {
        for (int64_t r = row_lo; r < row_hi; ++r) out[n++] = r;
        return n;
    }
    if (!(fraction > 0.0) || row_hi <= row_lo) return 0;
    const uint64_t thr = (uint64_t)floor(fraction * 9007199254740992.0); /* 2^53 */
    if (thr == 0) return 0;
    const uint64_t key = fmo_mix64(seed + (uint64_t)iter);
    int tz = 0;
    while (!((thr >> tz) & 1ULL)) ++tz;
    const int last = 53 - tz; /* no row can be selected after the lowest set digit of p */
    for (int64_t q = row_lo >> 6; q <= (row_hi - 1) >> 6; ++q) {
        uint64_t hit = fmo_bernoulli_block(key, thr, last, (uint64_t)q);
        while (hit) {
            const int j = __builtin_ctzll(hit);
            hit &= hit - 1ULL;
            const int64_t r = (q << 6) + j;
            if (r >= row_lo && r < row_hi) out[n++] = r;
        }
    }
    return n;
}

int64_t fmo_partition_rows(uint64_t seed, int64_t n_parts, int64_t part, int64_t row_lo,
                           int64_t row_hi, int64_t* out) {
    const uint64_t key = fmo_mix64(seed);
    int64_t n = 0;
    for (int64_t r = row_lo; r < row_hi; ++r)
        if ((int64_t)((fmo_mix64(key ^ fmo_mix64((uint64_t)r)) >> 11) % (uint64_t)n_parts) == part)
            out[n++] = r;
    return n;
}

void fmo_init_v(double* v, int64_t count, double mean, double stdev, uint64_t seed) {
    const uint64_t s = fmo_mix64(seed);
    const double two_pi = 6.283185307179586476925286766559;
    for (int64_t e = 0; e < count; ++e) {
        const double u1 =
            (double)((fmo_mix64(s + 2ULL * (uint64_t)e) >> 11) + 1ULL) * 0x1.0p-53;
        const double u2 = (double)(fmo_mix64(s + 2ULL * (uint64_t)e + 1ULL) >> 11) * 0x1.0p-53;
        const double z = sqrt(-2.0 * log(u1)) * cos(two_pi * u2);
        v[e] = (double)(float)(mean + stdev * z);
    }
}

/* ------------------------------------------------------------------------------------------
 * ALS.learn, fm/lib/ALS.scala:15-75 (see fm_oracle.h).
 * ---------------------------------------------------------------------------------------- */
static int als_updatable(double nv, double v) { /* :178-180 */
    return !isnan(nv) && !isinf(nv) && nv != v;
}

/* computeTheta (:167-176): returns the accepted value (theta itself when rejected). */
static double als_theta(double theta, double reg, double sum_e_h, double sum_h_sqr) {
    const double nv = -(sum_e_h - theta * sum_h_sqr) / (reg + sum_h_sqr);
    return als_updatable(nv, theta) ? nv : theta;
}

double fmo_als_sweep(const fmo_params* p, double* w0, double* w, double* v,
                     const int64_t* row_ptr, const int32_t* idx, const double* val,
                     const double* label, int64_t n_rows, int32_t flags, double* e_out) {
    const int32_t k = p->k;
    const int64_t n_slots = p->n_slots;
    const int64_t nnz = n_rows > 0 ? row_ptr[n_rows] : 0;
    const int quirks = (flags & FMO_ALS_REF_QUIRKS) != 0, f32 = (flags & FMO_ALS_STORE_F32) != 0;
    const int64_t id_end = quirks ? n_slots - 1 : n_slots;       /* `0 until num_attribute` */
    double* e = (double*)malloc(sizeof(double) * (size_t)(n_rows > 0 ? n_rows : 1));
    double* q = (double*)malloc(sizeof(double) * (size_t)(n_rows > 0 ? n_rows : 1));
    int64_t* colptr = (int64_t*)calloc((size_t)n_slots + 1, sizeof(int64_t));
    int64_t* crow = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nnz > 0 ? nnz : 1));
    double* cval = (double*)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * ((size_t)n_slots + 1));
    double rmse = -1.0;
    if (!e || !q || !colptr || !crow || !cval || !fill) goto done;

    /* transposeInput (DataSet.scala:31-38): column id -> (row, value), rows ascending */
    for (int64_t j = 0; j < nnz; ++j) colptr[idx[j] + 1]++;
    for (int64_t i = 0; i < n_slots; ++i) colptr[i + 1] += colptr[i];
    memcpy(fill, colptr, sizeof(int64_t) * ((size_t)n_slots + 1));
    for (int64_t r = 0; r < n_rows; ++r)
        for (int64_t j = row_ptr[r]; j < row_ptr[r + 1]; ++j) {
            const int64_t at = fill[idx[j]]++;
            if (at > colptr[idx[j]] && crow[at - 1] == r) goto done;   /* duplicate (row, id) */
            crow[at] = r;
            cval[at] = val[j];
        }

    /* e = predict - y (:142-144) */
    for (int64_t r = 0; r < n_rows; ++r) {
        const int64_t b = row_ptr[r];
        e[r] = fmo_predict_row(p, *w0, w, v, idx + b, val + b, row_ptr[r + 1] - b) - label[r];
    }

    if (p->k0) { /* :19-28 */
        double sum = 0.0;
        if (n_rows > 0) {
            sum = e[0];
            for (int64_t r = 1; r < n_rows; ++r) sum = sum + e[r];   /* error.reduce(_+_) */
        }
        double nw0 = als_theta(*w0, p->reg0, sum, (double)n_rows);   /* drawGlobalBias :152-154 */
        if (f32) nw0 = (double)(float)nw0;
        /* :24 adds (w0* - fm.w0) inside a lazy, uncached RDD (:142-144) that is first materialised
         * at :31, after `fm.w0 = w0` (:27): the added term is then 0, but the re-evaluation runs
         * fm.predict with the NEW w0 -- the residuals the sweep uses are yhat_new - y, i.e. the
         * old ones shifted by the w0 step, with or without FMO_ALS_REF_QUIRKS. */
        if (als_updatable(nw0, *w0))
            for (int64_t r = 0; r < n_rows; ++r) e[r] = e[r] + (nw0 - *w0);
        *w0 = nw0;
    }

    if (p->k1) { /* :36-43 */
        for (int64_t id = 0; id < id_end; ++id) {
            const int64_t a = colptr[id], b = colptr[id + 1];
            if (b == a) continue;                                    /* features.contains(id) */
            double sum_h_sqr = cval[a] * cval[a], sum_e_h = e[crow[a]] * cval[a]; /* :183-190 */
            for (int64_t j = a + 1; j < b; ++j) {
                sum_h_sqr = sum_h_sqr + cval[j] * cval[j];
                sum_e_h = sum_e_h + e[crow[j]] * cval[j];
            }
            double nt = als_theta(w[id], p->regw, sum_e_h, sum_h_sqr);
            if (f32) nt = (double)(float)nt;
            if (als_updatable(nt, w[id]))                            /* drawTheta :156-165 */
                for (int64_t j = a; j < b; ++j) e[crow[j]] += cval[j] * (nt - w[id]); /* :194-198 */
            w[id] = nt;
        }
    }

    for (int32_t f = 0; f < k; ++f) { /* :45-70 */
        for (int64_t r = 0; r < n_rows; ++r) {                       /* precomputeTermQ :146-150 */
            double s = 0.0;
            for (int64_t j = row_ptr[r]; j < row_ptr[r + 1]; ++j)
                s += v[(int64_t)idx[j] * k + f] * val[j];
            q[r] = s;
        }
        for (int64_t id = 0; id < id_end; ++id) {
            const int64_t a = colptr[id], b = colptr[id + 1];
            if (b == a) continue;
            const double vo = v[id * k + f];
            double sum_h_sqr = 0.0, sum_e_h = 0.0;
            for (int64_t j = a; j < b; ++j) {                        /* h = x q - x^2 v  (:56-58) */
                const double h = cval[j] * q[crow[j]] - cval[j] * cval[j] * vo;
                if (j == a) { sum_h_sqr = h * h; sum_e_h = e[crow[j]] * h; }
                else { sum_h_sqr = sum_h_sqr + h * h; sum_e_h = sum_e_h + e[crow[j]] * h; }
            }
            double nv = als_theta(vo, p->regv, sum_e_h, sum_h_sqr);
            if (f32) nv = (double)(float)nv;
            if (als_updatable(nv, vo))
                for (int64_t j = a; j < b; ++j) {
                    const double h = cval[j] * q[crow[j]] - cval[j] * cval[j] * vo;
                    e[crow[j]] += h * (nv - vo);
                }
            for (int64_t j = a; j < b; ++j) q[crow[j]] += cval[j] * (nv - vo);   /* :60-62 */
            v[id * k + f] = nv;                                      /* :64 */
        }
    }

    {
        double ss = 0.0;
        for (int64_t r = 0; r < n_rows; ++r) ss += e[r] * e[r];
        rmse = n_rows > 0 ? sqrt(ss / (double)n_rows) : 0.0;
        if (e_out) memcpy(e_out, e, sizeof(double) * (size_t)n_rows);
    }
done:
    free(e); free(q); free(colptr); free(crow); free(cval); free(fill);
    return rmse;
}
