// sfm_kernels.cu -- hand-written sm_100a kernels of the SparkFM hot path.
//
// Reference citations: /root/reference/src/main/scala/io/edstud/spark/ (the reference has no
// native code; each kernel names the Scala it replaces).  Spec sections: DESIGN.md.
//
//   fm_forward_kernel   FMModel.predict (fm/FMModel.scala:34-63) for a batch of CSR rows, one
//                       warp per row; in TRAIN mode also the per-sample loss / multiplier of the
//                       SGD gradient and the (feature, row) entry list that the
//                       reduce-by-feature consumes.
//   fm_pull_kernel      deterministic reduce-by-feature: for every feature the
//                       gradient is PULLED from the rows that contain it, in sorted (fixed)
//                       order -- no float atomics, no per-entry gradient is ever materialised.
//                       Fused with the SGD update on one GPU.
//   fm_update_kernel    dense regularised update from the all-reduced gradient (multi-GPU).
//
// All of it is HBM / L2 bound gather + reduction work: tensor cores do not apply.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "sfm_common.h"

namespace sfm {

#define FULL 0xffffffffu

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

__device__ __forceinline__ float4 shfl_xor4(float4 a, int off) {
    float4 r;
    r.x = __shfl_xor_sync(FULL, a.x, off);
    r.y = __shfl_xor_sync(FULL, a.y, off);
    r.z = __shfl_xor_sync(FULL, a.z, off);
    r.w = __shfl_xor_sync(FULL, a.w, off);
    return r;
}

// Per-sample loss and dLoss/dyhat, DESIGN.md 2.2 (oracle: fmo_loss_mult).
__device__ __forceinline__ void loss_mult(int task, float yhat, float label, float& loss,
                                          float& mult) {
    if (task == SFM_TASK_CLASSIFICATION) {
        const float y = label > 0.f ? 1.f : -1.f;
        const float m = y * yhat;
        const float e = expf(-fabsf(m));
        loss = fmaxf(-m, 0.f) + log1pf(e);
        const float sig = m > 0.f ? e / (1.f + e) : 1.f / (1.f + e);
        mult = -y * sig;
    } else {
        const float d = yhat - label;
        loss = d * d;
        mult = d;
    }
}

// ------------------------------------------------------------------------------------------
// Forward.  One warp per row.  A V row is kp floats = LPR float4; the warp's 32 lanes are
// NPP = 32/LPR "entry slots" x LPR "factor quads": lane = slot*LPR + fq.  Up to 64 CSR entries
// of the row (two tiles of 32) are read coalesced, one idx (+ one val) per lane per tile, and
// kept in registers; each pass broadcasts NPP of them with shuffles and every lane gathers ONE
// float4 of the entry's V row, so a pass issues NPP fully-used 16*LPR-byte row reads per warp.
//
// The pairwise term is accumulated as  P <- P + a*S; S <- S + a  (a = v_if * x_i), i.e.
// sum_{i<j} a_i a_j, which equals the reference's 0.5*(S^2 - sum a^2) (FMModel.scala:50) exactly
// in real arithmetic but has no cancellation in fp32; partial (S, P) pairs of different lanes
// merge with P = P1 + P2 + S1*S2.  The cross-slot merge halves the components a lane carries at
// each of the first two levels (4 -> 2 -> 1), so it costs 4+2+2.. shuffles instead of 8 per level.
//
// The kernel is instruction-issue bound (ncu: ~60 % issue-active, L2 ~21 %), hence the
// specialisations: HAS_VAL = false drops every multiply by x; UNIFORM rows need no row_ptr.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void acc_entry(float4& s, float4& p, const float4& a) {
    p.x = fmaf(a.x, s.x, p.x); s.x += a.x;
    p.y = fmaf(a.y, s.y, p.y); s.y += a.y;
    p.z = fmaf(a.z, s.z, p.z); s.z += a.z;
    p.w = fmaf(a.w, s.w, p.w); s.w += a.w;
}

// Per-sample loss and dLoss/dyhat, DESIGN.md 2.2 (oracle: fmo_loss_mult).  Fast intrinsics:
// relative error ~1e-6, far inside the 1e-4 loss tolerance.
__device__ __forceinline__ void loss_mult_fast(int task, float yhat, float label, float& loss,
                                               float& mult) {
    if (task == SFM_TASK_CLASSIFICATION) {
        const float y = label > 0.f ? 1.f : -1.f;
        const float m = y * yhat;
        const float e = __expf(-fabsf(m));
        const float r = __fdividef(1.f, 1.f + e);
        loss = fmaxf(-m, 0.f) + (e > 1e-4f ? __logf(1.f + e) : e * (1.f - 0.5f * e));
        mult = -y * (m > 0.f ? e * r : r);
    } else {
        const float d = yhat - label;
        loss = d * d;
        mult = d;
    }
}

// `id` is a valid row of V for every lane: entries past the end of the CSR row (and reported
// out-of-range indices) point at the all-zero row n_slots that sfm_create appends to V and w, so
// the pass loop needs no validity selects.  Row offsets are 32-bit (sfm_create bounds
// n_slots * LPR below 2^32).
template <int LPR, bool HAS_VAL>
__device__ __forceinline__ void forward_tile(const float4* __restrict__ V4,
                                             const float* __restrict__ W, int id, float x, int cnt,
                                             int slot, int fq, float4& s, float4& p, float& lin) {
    constexpr int NPP = 32 / LPR;
    constexpr int PCH = LPR < 8 ? LPR : 8;  // passes whose gathers are in flight together
    if (cnt == 32) {
#pragma unroll
        for (int t0 = 0; t0 < LPR; t0 += PCH) {
            float4 vv[PCH];
            float ww[PCH], xx[PCH];
#pragma unroll
            for (int t = 0; t < PCH; ++t) {
                const uint32_t row = (uint32_t)__shfl_sync(FULL, id, (t0 + t) * NPP + slot);
                if (HAS_VAL) xx[t] = __shfl_sync(FULL, x, (t0 + t) * NPP + slot);
                vv[t] = __ldg(V4 + (row * LPR + fq));
                ww[t] = (fq == 0) ? __ldg(W + row) : 0.f;
            }
#pragma unroll
            for (int t = 0; t < PCH; ++t) {
                float4 a = vv[t];
                if (HAS_VAL) {
                    const float px = xx[t];
                    a.x *= px; a.y *= px; a.z *= px; a.w *= px;
                    lin = fmaf(ww[t], px, lin);
                } else {
                    lin += ww[t];
                }
                acc_entry(s, p, a);
            }
        }
    } else {
        // partial tile: only the passes that hold entries (warp-uniform trip count)
        const int npass = (cnt + NPP - 1) / NPP;
        for (int t = 0; t < npass; ++t) {
            const uint32_t row = (uint32_t)__shfl_sync(FULL, id, t * NPP + slot);
            float4 a = __ldg(V4 + (row * LPR + fq));
            const float wv = (fq == 0) ? __ldg(W + row) : 0.f;
            if (HAS_VAL) {
                const float px = __shfl_sync(FULL, x, t * NPP + slot);
                a.x *= px; a.y *= px; a.z *= px; a.w *= px;
                lin = fmaf(wv, px, lin);
            } else {
                lin += wv;
            }
            acc_entry(s, p, a);
        }
    }
}

template <int LPR, bool TRAIN, bool HAS_VAL, bool UNIFORM>
__global__ void __launch_bounds__(256)
fm_forward_kernel(const float4* __restrict__ V4, const float* __restrict__ W,
                  const float* __restrict__ W0, int64_t n_slots, int k0, int k1, int task,
                  const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ idx,
                  const float* __restrict__ val, const float* __restrict__ label,
                  const int32_t* __restrict__ row_ids, int64_t row_lo, int64_t n_rows,
                  int64_t idx_len, const int64_t* __restrict__ out_ptr, int64_t out_base,
                  int uniform_m, int key_bits, int blk_shift, float* __restrict__ S,
                  float* __restrict__ mult_out,
                  float* __restrict__ loss_out, float* __restrict__ yhat_out,
                  uint32_t* __restrict__ keys, uint2* __restrict__ pay, int32_t* __restrict__ err) {
    constexpr int KP = LPR * 4;
    const int lane = threadIdx.x & 31;
    const int slot = lane / LPR;
    const int fq = lane % LPR;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const float w0 = k0 ? __ldg(W0) : 0.f;

    for (int64_t pos = warp0; pos < n_rows; pos += nwarps) {
        const int64_t r = row_ids ? (int64_t)__ldg(row_ids + pos) : row_lo + pos;
        int64_t beg, end;
        if (UNIFORM) {
            beg = r * uniform_m;
            end = beg + uniform_m;
        } else {
            beg = __ldg(row_ptr + r);
            end = __ldg(row_ptr + r + 1);
            if (end < beg || beg < 0 || end > idx_len) {  // malformed CSR: reported, row = empty
                if (lane == 0) atomicExch(err, 1);
                end = beg;
            }
        }
        int64_t obase = 0;
        if (TRAIN) obase = out_ptr ? __ldg(out_ptr + pos) - out_base : pos * (int64_t)uniform_m;
        // sort key = (row block << key_bits) | feature: the reduce then sweeps one L2-sized block
        // of S rows at a time
        const uint32_t kpre = TRAIN ? (uint32_t)(pos >> blk_shift) << key_bits : 0u;

        float4 s = f4_zero(), p = f4_zero();
        float lin = 0.f;
        const int zrow = (int)n_slots;  // the appended all-zero row
        for (int64_t tile = beg; tile < end; tile += 64) {
            const int64_t j0 = tile + lane, j1 = j0 + 32;
            int ia = zrow, ib = zrow;
            float xa = 0.f, xb = 0.f;
            if (j0 < end) {
                ia = __ldg(idx + j0);
                if (HAS_VAL) xa = __ldg(val + j0);
            }
            if (j1 < end) {
                ib = __ldg(idx + j1);
                if (HAS_VAL) xb = __ldg(val + j1);
            }
            if ((uint32_t)ia > (uint32_t)zrow || (uint32_t)ib > (uint32_t)zrow ||
                (j0 < end && ia == zrow) || (j1 < end && ib == zrow)) {
                atomicExch(err, 1);  // out-of-range index: reported, never dereferenced
                if ((uint32_t)ia > (uint32_t)zrow) ia = zrow;
                if ((uint32_t)ib > (uint32_t)zrow) ib = zrow;
            }
            if (TRAIN && keys != nullptr) {  // (nullptr: the transposition is cached)
                // entry list for the reduce-by-feature: key = block | feature, payload = {row, x}
                // (8 bytes) or, for all-ones data, just the row (4 bytes: less to sort)
                if (j0 < end) {
                    keys[obase + (j0 - beg)] = kpre | (ia == zrow ? 0u : (uint32_t)ia);
                    if (HAS_VAL) pay[obase + (j0 - beg)] = make_uint2((uint32_t)pos, ia == zrow ? 0u : __float_as_uint(xa));
                    else if (pay) reinterpret_cast<uint32_t*>(pay)[obase + (j0 - beg)] = (uint32_t)pos;
                }
                if (j1 < end) {
                    keys[obase + (j1 - beg)] = kpre | (ib == zrow ? 0u : (uint32_t)ib);
                    if (HAS_VAL) pay[obase + (j1 - beg)] = make_uint2((uint32_t)pos, ib == zrow ? 0u : __float_as_uint(xb));
                    else if (pay) reinterpret_cast<uint32_t*>(pay)[obase + (j1 - beg)] = (uint32_t)pos;
                }
            }
            const int cnt = (int)min((int64_t)64, end - tile);
            forward_tile<LPR, HAS_VAL>(V4, W, ia, xa, min(cnt, 32), slot, fq, s, p, lin);
            if (cnt > 32) forward_tile<LPR, HAS_VAL>(V4, W, ib, xb, cnt - 32, slot, fq, s, p, lin);
        }

        // ---- merge the NPP entry slots; components carried per lane: 4 -> 2 -> 1
        float sv[4] = {s.x, s.y, s.z, s.w};
        float pv[4] = {p.x, p.y, p.z, p.w};
        int ncomp = 4;
#pragma unroll
        for (int off = 16; off >= LPR; off >>= 1) {
            if (ncomp > 1) {
                const int half = ncomp >> 1;
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (c < half) {
                        const float ss = __shfl_xor_sync(FULL, up ? sv[c] : sv[c + half], off);
                        const float ps = __shfl_xor_sync(FULL, up ? pv[c] : pv[c + half], off);
                        const float sk = up ? sv[c + half] : sv[c];
                        const float pk = up ? pv[c + half] : pv[c];
                        pv[c] = (pk + ps) + sk * ss;
                        sv[c] = sk + ss;
                    }
                }
                ncomp = half;
            } else {
                const float ss = __shfl_xor_sync(FULL, sv[0], off);
                const float ps = __shfl_xor_sync(FULL, pv[0], off);
                pv[0] = (pv[0] + ps) + sv[0] * ss;
                sv[0] += ss;
            }
        }
        // which factors this lane now holds: quad fq, component offset from the two halvings
        constexpr int LV = (LPR == 32) ? 0 : (LPR == 16) ? 1 : (LPR == 8) ? 2 : (LPR == 4) ? 3
                         : (LPR == 2) ? 4 : 5;
        constexpr int NFIN = LV == 0 ? 4 : (LV == 1 ? 2 : 1);
        constexpr int DUPMASK = (LPR <= 4) ? (7 & ~(LPR - 1)) : 0;  // lanes holding duplicates
        int cbase = 0;
        if (LV >= 1) cbase += (lane & 16) ? 2 : 0;
        if (LV >= 2) cbase += (lane & 8) ? 1 : 0;
        const bool canon = (lane & DUPMASK) == 0;
        float tot = 0.f;
        if (canon) {
#pragma unroll
            for (int c = 0; c < NFIN; ++c) tot += pv[c];
        }
        if (k1) tot += lin;  // lin is non-zero on fq == 0 lanes only, one partial per slot
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) tot += __shfl_xor_sync(FULL, tot, off);
        const float yhat = w0 + tot;

        if (TRAIN) {
            if (canon) {
#pragma unroll
                for (int c = 0; c < NFIN; ++c) S[pos * KP + fq * 4 + cbase + c] = sv[c];
            }
            float ls = 0.f, mu = 0.f;
            loss_mult_fast(task, yhat, __ldg(label + r), ls, mu);  // uniform across the warp
            if (lane == 0) {
                loss_out[pos] = ls;
                mult_out[pos] = mu;
            }
        } else {
            if (lane == 0) yhat_out[pos] = yhat;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Fast path of the forward for the one-hot CTR layout (BASELINE configs 3/4 at k <= 16): resident,
// index-validated data set, every row has exactly m <= 64 entries, no value array, kp = 16.
// Same arithmetic and lane mapping as fm_forward_kernel<4, TRAIN, false, true>, minus everything
// generic: no row_ptr, no per-entry validity checks (sfm_load_dataset validated the indices), 32-bit
// address arithmetic, and the NEXT row's indices are loaded before the current row's gathers are
// consumed, which takes one memory round trip off the per-row dependency chain.
// ------------------------------------------------------------------------------------------
template <bool TRAIN>
__global__ void __launch_bounds__(256)
fm_forward_onehot16_kernel(const float4* __restrict__ V4, const float* __restrict__ W,
                           const float* __restrict__ W0, int zrow, int k0, int k1, int task,
                           const int32_t* __restrict__ idx, const float* __restrict__ label,
                           const int32_t* __restrict__ row_ids, int row_lo, int n_rows, int m,
                           int key_bits, int blk_shift, float* __restrict__ S,
                           float* __restrict__ mult_out, float* __restrict__ loss_out,
                           float* __restrict__ yhat_out, uint32_t* __restrict__ keys,
                           uint32_t* __restrict__ pay) {
    constexpr int LPR = 4, NPP = 8;
    const int lane = threadIdx.x & 31;
    const int slot = lane >> 2;
    const int fq = lane & 3;
    const int warp0 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    const float w0 = k0 ? __ldg(W0) : 0.f;
    const bool two = m > 32;
    const int npass_b = two ? (m - 32 + NPP - 1) / NPP : 0;   // passes over the second tile
    const bool has_a = lane < m, has_b = lane + 32 < m;

    int pos = warp0;
    int r = 0, ia = zrow, ib = zrow;
    if (pos < n_rows) {
        r = row_ids ? __ldg(row_ids + pos) : row_lo + pos;
        const int32_t* p = idx + (int64_t)r * m;
        if (has_a) ia = __ldg(p + lane);
        if (has_b) ib = __ldg(p + lane + 32);
    }
    while (pos < n_rows) {
        // ---- prefetch the next row's indices
        const int pos_n = pos + nwarps;
        int r_n = 0, ia_n = zrow, ib_n = zrow;
        if (pos_n < n_rows) {
            r_n = row_ids ? __ldg(row_ids + pos_n) : row_lo + pos_n;
            const int32_t* p = idx + (int64_t)r_n * m;
            if (has_a) ia_n = __ldg(p + lane);
            if (has_b) ib_n = __ldg(p + lane + 32);
        }
        const float lab = TRAIN ? __ldg(label + r) : 0.f;
        if (TRAIN && keys != nullptr) {
            const uint32_t kpre = (uint32_t)(pos >> blk_shift) << key_bits;
            const uint32_t o = (uint32_t)pos * (uint32_t)m + lane;
            // pay == nullptr: the sort derives the row from the entry position (o / m)
            if (has_a) { keys[o] = kpre | (uint32_t)ia; if (pay) pay[o] = (uint32_t)pos; }
            if (has_b) { keys[o + 32] = kpre | (uint32_t)ib; if (pay) pay[o + 32] = (uint32_t)pos; }
        }
        // ---- gathers: tile A (4 passes; lanes past m point at the zero row), then tile B
        float4 s = f4_zero(), p = f4_zero();
        float lin = 0.f;
        {
            float4 vv[4];
            float ww[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const uint32_t row = (uint32_t)__shfl_sync(FULL, ia, t * NPP + slot);
                vv[t] = __ldg(V4 + (row * LPR + fq));
                ww[t] = (fq == 0) ? __ldg(W + row) : 0.f;
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                lin += ww[t];
                acc_entry(s, p, vv[t]);
            }
        }
        for (int t = 0; t < npass_b; ++t) {
            const uint32_t row = (uint32_t)__shfl_sync(FULL, ib, t * NPP + slot);
            const float4 a = __ldg(V4 + (row * LPR + fq));
            lin += (fq == 0) ? __ldg(W + row) : 0.f;
            acc_entry(s, p, a);
        }
        // ---- merge the 8 entry slots: components per lane 4 -> 2 -> 1 (see fm_forward_kernel)
        float s0, s1, p0, p1;
        {
            const bool up = (lane & 16) != 0;
            const float sa = __shfl_xor_sync(FULL, up ? s.x : s.z, 16);
            const float pa = __shfl_xor_sync(FULL, up ? p.x : p.z, 16);
            const float sb = __shfl_xor_sync(FULL, up ? s.y : s.w, 16);
            const float pb = __shfl_xor_sync(FULL, up ? p.y : p.w, 16);
            const float ska = up ? s.z : s.x, pka = up ? p.z : p.x;
            const float skb = up ? s.w : s.y, pkb = up ? p.w : p.y;
            p0 = (pka + pa) + ska * sa; s0 = ska + sa;
            p1 = (pkb + pb) + skb * sb; s1 = skb + sb;
        }
        {
            const bool up = (lane & 8) != 0;
            const float ss = __shfl_xor_sync(FULL, up ? s0 : s1, 8);
            const float ps = __shfl_xor_sync(FULL, up ? p0 : p1, 8);
            const float sk = up ? s1 : s0, pk = up ? p1 : p0;
            p0 = (pk + ps) + sk * ss; s0 = sk + ss;
        }
        {
            const float ss = __shfl_xor_sync(FULL, s0, 4);
            const float ps = __shfl_xor_sync(FULL, p0, 4);
            p0 = (p0 + ps) + s0 * ss; s0 += ss;
        }
        const bool canon = (lane & 4) == 0;
        float tot = canon ? p0 : 0.f;
        if (k1) tot += lin;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) tot += __shfl_xor_sync(FULL, tot, off);
        const float yhat = w0 + tot;
        if (TRAIN) {
            const int cbase = ((lane & 16) ? 2 : 0) + ((lane & 8) ? 1 : 0);
            if (canon) S[(uint32_t)pos * 16u + fq * 4 + cbase] = s0;
            float ls, mu;
            loss_mult_fast(task, yhat, lab, ls, mu);
            if (lane == 0) {
                loss_out[pos] = ls;
                mult_out[pos] = mu;
            }
        } else if (lane == 0) {
            yhat_out[pos] = yhat;
        }
        pos = pos_n; r = r_n; ia = ia_n; ib = ib_n;
    }
}

template <int LPR>
static cudaError_t forward_dispatch(const ModelView& m, const BatchView& b, const FwdOut& o,
                                    bool train, int32_t* d_err, int sm_count, cudaStream_t st) {
    if (b.n_rows <= 0) return cudaSuccess;
    const int wpb = 8;
    int64_t blocks = (b.n_rows + wpb - 1) / wpb;
    const int64_t cap = (int64_t)sm_count * 16;  // persistent: <= 16 CTAs of 8 warps per SM queued
    if (blocks > cap) blocks = cap;
#define FWD_ARGS                                                                              \
    (const float4*)m.v, m.w, m.w0, m.n_slots, m.k0, m.k1, m.task, b.row_ptr, b.idx, b.val,    \
        b.label, b.row_ids, b.row_lo, b.n_rows, b.idx_len, b.out_ptr, b.out_base,             \
        b.uniform_m, o.key_bits, o.blk_shift, o.S, o.mult, o.loss, o.yhat, o.keys, o.pay, d_err
#define FWD_LAUNCH(T, HV, UN) fm_forward_kernel<LPR, T, HV, UN><<<g, t, 0, st>>>(FWD_ARGS)
    const dim3 g((unsigned)blocks), t(256);
    const bool un = b.uniform_m >= 0;
    if (LPR == 4 && un && !b.val && b.validated && b.uniform_m <= 64 && b.n_rows < (1 << 26) &&
        !knobs().no_fastpath) {
        if (train)
            fm_forward_onehot16_kernel<true><<<g, t, 0, st>>>(
                (const float4*)m.v, m.w, m.w0, (int)m.n_slots, m.k0, m.k1, m.task, b.idx, b.label,
                b.row_ids, (int)b.row_lo, (int)b.n_rows, b.uniform_m, o.key_bits, o.blk_shift, o.S,
                o.mult, o.loss, o.yhat, o.keys, (uint32_t*)o.pay);
        else
            fm_forward_onehot16_kernel<false><<<g, t, 0, st>>>(
                (const float4*)m.v, m.w, m.w0, (int)m.n_slots, m.k0, m.k1, m.task, b.idx, b.label,
                b.row_ids, (int)b.row_lo, (int)b.n_rows, b.uniform_m, o.key_bits, o.blk_shift, o.S,
                o.mult, o.loss, o.yhat, o.keys, (uint32_t*)o.pay);
        return cudaGetLastError();
    }
    if (train) {
        if (b.val) { if (un) FWD_LAUNCH(true, true, true); else FWD_LAUNCH(true, true, false); }
        else       { if (un) FWD_LAUNCH(true, false, true); else FWD_LAUNCH(true, false, false); }
    } else {
        if (b.val) { if (un) FWD_LAUNCH(false, true, true); else FWD_LAUNCH(false, true, false); }
        else       { if (un) FWD_LAUNCH(false, false, true); else FWD_LAUNCH(false, false, false); }
    }
#undef FWD_LAUNCH
#undef FWD_ARGS
    return cudaGetLastError();
}

cudaError_t launch_forward(const ModelView& m, const BatchView& b, const FwdOut& o, bool train,
                           int32_t* d_err, int sm_count, cudaStream_t st, int64_t* launches) {
    if (b.n_rows > 0) ++*launches;
    switch (m.lpr) {
        case 1: return forward_dispatch<1>(m, b, o, train, d_err, sm_count, st);
        case 2: return forward_dispatch<2>(m, b, o, train, d_err, sm_count, st);
        case 4: return forward_dispatch<4>(m, b, o, train, d_err, sm_count, st);
        case 8: return forward_dispatch<8>(m, b, o, train, d_err, sm_count, st);
        case 16: return forward_dispatch<16>(m, b, o, train, d_err, sm_count, st);
        case 32: return forward_dispatch<32>(m, b, o, train, d_err, sm_count, st);
    }
    return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------
// Fixed-shape fp64 reductions (deterministic: the mapping element -> thread -> tree slot depends
// on n only).  Stage 1: RED_BLOCKS blocks, each a contiguous span; stage 2: one block.
// ------------------------------------------------------------------------------------------
constexpr int RED_BLOCKS = 296;
constexpr int RED_THREADS = 256;

template <int NV>
__device__ __forceinline__ void block_tree(double (&acc)[NV], double* sm /* [NV][RED_THREADS] */) {
    const int t = threadIdx.x;
#pragma unroll
    for (int v = 0; v < NV; ++v) sm[v * RED_THREADS + t] = acc[v];
    __syncthreads();
    for (int off = RED_THREADS / 2; off > 0; off >>= 1) {
        if (t < off) {
#pragma unroll
            for (int v = 0; v < NV; ++v) sm[v * RED_THREADS + t] += sm[v * RED_THREADS + t + off];
        }
        __syncthreads();
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = sm[v * RED_THREADS];
}

// One launch: every block reduces its span to a partial; the block that arrives last (integer
// ticket, reset for the next launch) adds the partials in block order -- the same fixed fp64 tree
// whichever block that is.
__global__ void __launch_bounds__(RED_THREADS)
scalar_reduce_kernel(const float* __restrict__ loss, const float* __restrict__ mult, int64_t n,
                     double* __restrict__ partials, unsigned int* __restrict__ ticket,
                     double* __restrict__ d_scal, const int32_t* __restrict__ err) {
    __shared__ double sm[2 * RED_THREADS];
    __shared__ bool last;
    const int64_t span = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * span;
    const int64_t hi = min(n, lo + span);
    double acc[2] = {0.0, 0.0};
    for (int64_t i = lo + threadIdx.x; i < hi; i += RED_THREADS) {
        acc[0] += (double)loss[i];
        acc[1] += (double)mult[i];
    }
    block_tree<2>(acc, sm);
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x] = acc[0];
        partials[2 * blockIdx.x + 1] = acc[1];
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    acc[0] = acc[1] = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += RED_THREADS) {
        acc[0] += __ldcg(partials + 2 * i);
        acc[1] += __ldcg(partials + 2 * i + 1);
    }
    block_tree<2>(acc, sm);
    if (threadIdx.x == 0) {
        d_scal[SC_LOSS] = acc[0];
        d_scal[SC_GW0] = acc[1];
        d_scal[SC_COUNT] = (double)n;
        d_scal[SC_ERR] = (err && *err) ? 1.0 : 0.0;   // summed over the ranks: everybody skips the update
        *ticket = 0u;
    }
}

cudaError_t launch_scalar_reduce(const float* loss, const float* mult, int64_t n, double* partials,
                                 unsigned int* ticket, double* d_scal, const int32_t* d_err,
                                 cudaStream_t st, int64_t* launches) {
    *launches += 1;   // ticket: one device word, zero at handle creation, reset by the kernel itself
    scalar_reduce_kernel<<<RED_BLOCKS, RED_THREADS, 0, st>>>(loss, mult, n, partials, ticket, d_scal, d_err);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(RED_THREADS)
metrics1_kernel(const float* __restrict__ yhat, const float* __restrict__ label, int64_t n,
                double* __restrict__ partials) {
    __shared__ double sm[4 * RED_THREADS];
    const int64_t span = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * span;
    const int64_t hi = min(n, lo + span);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t i = lo + threadIdx.x; i < hi; i += RED_THREADS) {
        const float y = label[i], p = yhat[i];
        const double e = (double)y - (double)p;  // Model.scala:14 (y - yhat)
        acc[0] += e * e;
        acc[1] += e;
        acc[2] += ((y >= 0.f && p >= 0.f) || (y < 0.f && p < 0.f)) ? 1.0 : 0.0;  // Model.scala:29
        float ls, mu;
        loss_mult(SFM_TASK_CLASSIFICATION, p, y, ls, mu);
        acc[3] += (double)ls;
    }
    block_tree<4>(acc, sm);
    if (threadIdx.x == 0)
        for (int v = 0; v < 4; ++v) partials[4 * blockIdx.x + v] = acc[v];
}

__global__ void __launch_bounds__(RED_THREADS)
metrics2_kernel(const double* __restrict__ partials, int nblocks, double* __restrict__ out4) {
    __shared__ double sm[4 * RED_THREADS];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < nblocks; i += RED_THREADS)
        for (int v = 0; v < 4; ++v) acc[v] += partials[4 * i + v];
    block_tree<4>(acc, sm);
    if (threadIdx.x == 0)
        for (int v = 0; v < 4; ++v) out4[v] += acc[v];  // accumulates over tiles of rows
}

// AUC: rank-sum of the positives over the ascending-sorted scores, ties get their average rank
__global__ void __launch_bounds__(RED_THREADS)
auc_ranks1_kernel(const float* __restrict__ score, const uint32_t* __restrict__ row,
                  const float* __restrict__ label, int64_t n, double* __restrict__ partials) {
    __shared__ double sm[2 * RED_THREADS];
    const int64_t span = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * span;
    const int64_t hi = min(n, lo + span);
    double acc[2] = {0.0, 0.0};
    for (int64_t i = lo + threadIdx.x; i < hi; i += RED_THREADS) {
        if (!(label[row[i]] > 0.f)) continue;
        const float sc = score[i];
        int64_t first = i, last = i;
        if (i > 0 && score[i - 1] == sc) {           // tie group: find its ends by binary search
            int64_t a = 0, b = i;
            while (a < b) { const int64_t mid = (a + b) >> 1; if (score[mid] < sc) a = mid + 1; else b = mid; }
            first = a;
        }
        if (i + 1 < n && score[i + 1] == sc) {
            int64_t a = i + 1, b = n;
            while (a < b) { const int64_t mid = (a + b) >> 1; if (score[mid] <= sc) a = mid + 1; else b = mid; }
            last = a - 1;
        }
        acc[0] += 0.5 * (double)(first + last) + 1.0;
        acc[1] += 1.0;
    }
    block_tree<2>(acc, sm);
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x] = acc[0];
        partials[2 * blockIdx.x + 1] = acc[1];
    }
}

__global__ void __launch_bounds__(RED_THREADS)
auc_ranks2_kernel(const double* __restrict__ partials, int nblocks, double* __restrict__ out2) {
    __shared__ double sm[2 * RED_THREADS];
    double acc[2] = {0.0, 0.0};
    for (int i = threadIdx.x; i < nblocks; i += RED_THREADS) {
        acc[0] += partials[2 * i];
        acc[1] += partials[2 * i + 1];
    }
    block_tree<2>(acc, sm);
    if (threadIdx.x == 0) {
        out2[0] = acc[0];
        out2[1] = acc[1];
    }
}

cudaError_t launch_auc_ranks(const float* sorted_scores, const uint32_t* sorted_rows,
                             const float* label, int64_t n, double* partials, double* out2,
                             cudaStream_t st, int64_t* launches) {
    *launches += 2;
    auc_ranks1_kernel<<<RED_BLOCKS, RED_THREADS, 0, st>>>(sorted_scores, sorted_rows, label, n, partials);
    auc_ranks2_kernel<<<1, RED_THREADS, 0, st>>>(partials, RED_BLOCKS, out2);
    return cudaGetLastError();
}

__global__ void iota_u32_kernel(uint32_t* __restrict__ p, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = (uint32_t)i;
}

cudaError_t launch_iota_u32(uint32_t* p, int64_t n, cudaStream_t st, int64_t* launches) {
    if (n <= 0) return cudaSuccess;
    ++*launches;
    int64_t blocks = (n + 255) / 256;
    if (blocks > (int64_t)device_sm_count() * 16) blocks = (int64_t)device_sm_count() * 16;
    iota_u32_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, n);
    return cudaGetLastError();
}

cudaError_t launch_metrics(const float* yhat, const float* label, int64_t n, double* partials,
                           double* out4, cudaStream_t st, int64_t* launches) {
    *launches += 2;
    metrics1_kernel<<<RED_BLOCKS, RED_THREADS, 0, st>>>(yhat, label, n, partials);
    metrics2_kernel<<<1, RED_THREADS, 0, st>>>(partials, RED_BLOCKS, out4);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Reduce-by-feature, pull form.  With c = mult_r * x_ri:
//     gV_if = sum_r c * S_rf  -  v_if * sum_r c * x_ri        gw_i = sum_r c
// so a feature needs only the factor sums S_r (kp floats) and multipliers of the rows that
// contain it -- no per-entry gradient is ever materialised.  The entries arrive sorted by
// (feature, batch position); the reduction tree over them depends on POSITIONS only:
//
//   level 0  a group of LPR lanes walks SUB consecutive sorted entries, adding in order;
//   level 1  a CTA covers G*SUB consecutive entries; a run that spans several groups is summed
//            by the group where it starts, over the following groups' "head" partials, in order;
//   level 2  a run that spans several CTA chunks leaves one record per chunk (R2[chunk]); the
//            finalize kernel adds them in chunk order to the record of the run's start (R1[i]).
//
// Work is therefore balanced by construction (every CTA gets the same number of entries, however
// skewed the feature frequencies are) and the result is bitwise reproducible -- no float atomics.
// The CTA's tile of sorted (key, payload) entries is staged in shared memory with coalesced
// 128-bit loads, stored transposed ([entry][group], +1 padding) so that the walk is
// bank-conflict free; the only gathers left are the S rows.  The kernel also records where every
// feature's run starts and ends (seg_lo / seg_hi, zeroed per step) for the finalize pass.
// Records are REC = kp + 4 floats: [A (kp) | D | C | 0 | 0].
// BINARY (data set without a value array): the sorted payload is just the row (4 bytes), x = 1,
// D = C, mult_r is fetched while staging; otherwise payload = {row, x} and mult_r is gathered in
// the walk.
// ------------------------------------------------------------------------------------------
constexpr int PULL_SUB = 32;
#ifndef PULL_U
#define PULL_U 4
#endif
constexpr uint32_t PULL_SENTINEL = 0xFFFFFFFFu;

template <int LPR>
struct PullCfg {
    static constexpr int THREADS = LPR >= 4 ? 256 : 64 * LPR;
    static constexpr int G = THREADS / LPR;        // groups per CTA (<= 64)
    static constexpr int CHB = G * PULL_SUB;       // sorted entries per CTA
    static constexpr int REC = LPR * 4 + 4;
};

template <int LPR>
__device__ __forceinline__ void store_rec(float* __restrict__ rec, int fq, const float4& A,
                                          float D, float C) {
    reinterpret_cast<float4*>(rec)[fq] = A;
    if (fq == 0) reinterpret_cast<float4*>(rec)[LPR] = make_float4(D, C, 0.f, 0.f);
}

template <int LPR, bool BINARY>
__global__ void __launch_bounds__(PullCfg<LPR>::THREADS, 1280 / PullCfg<LPR>::THREADS)
fm_pull_chunks_kernel(const uint32_t* __restrict__ keys, const uint2* __restrict__ pay,
                      const float4* __restrict__ S4, const float* __restrict__ mult, int nnz,
                      float* __restrict__ R1, float* __restrict__ R2,
                      int32_t* __restrict__ seg_lo, int32_t* __restrict__ seg_hi, int key_bits,
                      int64_t n_slots, int chunk_base) {
    using Cfg = PullCfg<LPR>;
    const int chunk = blockIdx.x + chunk_base;   // slices of the chunk grid can be launched apart
    const uint32_t kmask = (1u << key_bits) - 1u;
    // key = (row block << key_bits) | feature  ->  slot block * n_slots + feature
    auto slot_of = [=](uint32_t key) -> int64_t {
        return (int64_t)(key >> key_bits) * n_slots + (int64_t)(key & kmask);
    };
    constexpr int G = Cfg::G, CHB = Cfg::CHB, REC = Cfg::REC, SUB = PULL_SUB, U = PULL_U;
    __shared__ uint32_t key_s[SUB][G + 1];
    __shared__ uint32_t row_s[SUB][G + 1];
    __shared__ float val_s[SUB][G + 1];   // BINARY: mult_r; else: x
    __shared__ float4 headA[G][LPR];
    __shared__ float2 headDC[G];
    __shared__ uint32_t edge[2];           // key before / after this chunk

    const int tid = threadIdx.x;
    const int g = tid / LPR;
    const int fq = tid % LPR;
    const int cstart = chunk * CHB;

    // ---- stage the tile: thread -> 4 consecutive entries (one 16-byte + two 16-byte loads)
    for (int q = tid; q < CHB / 4; q += Cfg::THREADS) {
        const int e0 = q * 4;
        const int p = cstart + e0;
        uint32_t k4[4], r4[4];
        float v4[4];
        if (p + 3 < nnz) {
            const uint4 kk = __ldg(reinterpret_cast<const uint4*>(keys + p));
            k4[0] = kk.x; k4[1] = kk.y; k4[2] = kk.z; k4[3] = kk.w;
            if (BINARY) {
                const uint4 rr = __ldg(reinterpret_cast<const uint4*>(
                    reinterpret_cast<const uint32_t*>(pay) + p));
                r4[0] = rr.x; r4[1] = rr.y; r4[2] = rr.z; r4[3] = rr.w;
            } else {
                const uint4 pa = __ldg(reinterpret_cast<const uint4*>(pay + p));
                const uint4 pb = __ldg(reinterpret_cast<const uint4*>(pay + p) + 1);
                r4[0] = pa.x; r4[1] = pa.z; r4[2] = pb.x; r4[3] = pb.z;
                v4[0] = __uint_as_float(pa.y); v4[1] = __uint_as_float(pa.w);
                v4[2] = __uint_as_float(pb.y); v4[3] = __uint_as_float(pb.w);
            }
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool ok = p + u < nnz;
                k4[u] = ok ? __ldg(keys + p + u) : PULL_SENTINEL;
                if (BINARY) {
                    r4[u] = ok ? __ldg(reinterpret_cast<const uint32_t*>(pay) + p + u) : 0u;
                } else {
                    const uint2 pl = ok ? __ldg(pay + p + u) : make_uint2(0u, 0u);
                    r4[u] = pl.x;
                    v4[u] = __uint_as_float(pl.y);
                }
            }
        }
        if (BINARY) {  // the row's multiplier rides along (4 MB array, L2 resident)
#pragma unroll
            for (int u = 0; u < 4; ++u) v4[u] = __ldg(mult + r4[u]);
        }
        const int gg = e0 / SUB, ee = e0 % SUB;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            key_s[ee + u][gg] = k4[u];
            row_s[ee + u][gg] = r4[u];
            val_s[ee + u][gg] = v4[u];
        }
    }
    if (tid == 0) {
        edge[0] = cstart > 0 ? __ldg(keys + cstart - 1) : PULL_SENTINEL;
        edge[1] = cstart + CHB < nnz ? __ldg(keys + cstart + CHB) : PULL_SENTINEL;
    }
    __syncthreads();

    // ---- level 0: walk my SUB entries
    const int p0 = cstart + g * SUB;
    uint32_t cur = key_s[0][g];
    const uint32_t before = g > 0 ? key_s[SUB - 1][g - 1] : edge[0];
    bool cur_is_head = cur != PULL_SENTINEL && before == cur;
    if (cur != PULL_SENTINEL && !cur_is_head && fq == 0) seg_lo[slot_of(cur)] = p0;
    float4 A = f4_zero();
    float D = 0.f, C = 0.f;
#pragma unroll 1
    for (int base = 0; base < SUB; base += U) {
        uint32_t kk[U];
        float cc[U], xx[U];
        float4 sv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            kk[u] = key_s[base + u][g];
            const uint32_t row = row_s[base + u][g];
            const float second = val_s[base + u][g];
            if (BINARY) {
                cc[u] = second;
                xx[u] = 1.f;
            } else {
                xx[u] = second;
                cc[u] = __ldg(mult + row) * second;
            }
            sv[u] = __ldg(S4 + (row * LPR + fq));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (kk[u] != cur) {
                // the run of `cur` ended at position p0 + base + u
                if (cur != PULL_SENTINEL) {
                    if (cur_is_head) {
                        if (g == 0) store_rec<LPR>(R2 + (int64_t)chunk * REC, fq, A, D, C);
                        else { headA[g][fq] = A; if (fq == 0) headDC[g] = make_float2(D, C); }
                    } else {
                        store_rec<LPR>(R1 + slot_of(cur) * REC, fq, A, D, C);
                    }
                    if (fq == 0) seg_hi[slot_of(cur)] = p0 + base + u;
                }
                cur = kk[u];
                cur_is_head = false;
                if (cur != PULL_SENTINEL && fq == 0) seg_lo[slot_of(cur)] = p0 + base + u;
                A = f4_zero();
                D = 0.f;
                C = 0.f;
            }
            const float c = kk[u] == PULL_SENTINEL ? 0.f : cc[u];
            A.x = fmaf(c, sv[u].x, A.x);
            A.y = fmaf(c, sv[u].y, A.y);
            A.z = fmaf(c, sv[u].z, A.z);
            A.w = fmaf(c, sv[u].w, A.w);
            D = BINARY ? D : fmaf(c, xx[u], D);
            C += c;
        }
    }
    // ---- level 1: the run still open at the end of my sub-chunk
    const uint32_t after = g < G - 1 ? key_s[0][g + 1] : edge[1];
    const bool open = cur != PULL_SENTINEL;
    const bool continues = open && after == cur;
    bool owner = false;
    if (open) {
        if (!continues && fq == 0) seg_hi[slot_of(cur)] = p0 + SUB < nnz ? p0 + SUB : nnz;
        if (cur_is_head && g != 0) {
            headA[g][fq] = A;               // a middle / final piece of somebody else's run
            if (fq == 0) headDC[g] = make_float2(D, C);
        } else {
            owner = true;                   // starts here (R1) or is the chunk's inherited run (R2)
        }
    }
    __syncthreads();
    if (owner) {
        if (continues) {
            for (int g2 = g + 1; g2 < G; ++g2) {
                if (key_s[0][g2] != cur) break;
                const float4 a = headA[g2][fq];
                const float2 dc = headDC[g2];
                A.x += a.x; A.y += a.y; A.z += a.z; A.w += a.w;
                D += dc.x;
                C += dc.y;
                if (key_s[SUB - 1][g2] != cur) break;   // the run ended inside g2
            }
        }
        float* dst = cur_is_head ? R2 + (int64_t)chunk * REC : R1 + slot_of(cur) * REC;
        store_rec<LPR>(dst, fq, A, D, C);
    }
}

// Level 2 + gradient / update.  One group of LPR lanes per feature, every feature visited (so
// untouched slots receive their L2 decay).
template <int LPR, bool FUSED, bool BINARY>
__global__ void __launch_bounds__(256)
fm_pull_finalize_kernel(float4* __restrict__ V4, float* __restrict__ W, float* __restrict__ W0,
                        int64_t n_slots, int k0, int k1, const int32_t* __restrict__ seg_lo,
                        const int32_t* __restrict__ seg_hi, const float* __restrict__ R1,
                        const float* __restrict__ R2, int n_blocks,
                        const double* __restrict__ d_scal, const int32_t* __restrict__ err,
                        UpdateParams up, float4* __restrict__ G4, float* __restrict__ Gw,
                        float* __restrict__ Gw0, int64_t feat_lo, int64_t feat_hi) {
    constexpr int CHB = PullCfg<LPR>::CHB;
    constexpr int REC = PullCfg<LPR>::REC;
    if (FUSED && *err) return;  // a bad index was seen: leave the model untouched
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t ngroups = (int64_t)gridDim.x * blockDim.x / LPR;
    const int fq = threadIdx.x % LPR;
    const double count = d_scal[SC_COUNT];
    const float inv = count > 0.0 ? (float)(1.0 / count) : 0.f;
    const bool active = count > 0.0;

    if (tid == 0 && feat_lo == 0) {
        const float g0 = (float)d_scal[SC_GW0];
        if (FUSED) {
            if (k0 && active) {
                const float w0 = *W0;
                *W0 = sgd_step(w0, g0, inv, up.eta, up.reg0);
            }
        } else {
            *Gw0 = k0 ? g0 : 0.f;
        }
    }
    for (int64_t i = feat_lo + tid / LPR; i < feat_hi; i += ngroups) {
        float4 v = V4[i * LPR + fq];   // issued before the dependent record loads
        const float wi = (fq == 0 && k1) ? W[i] : 0.f;
        float4 A = f4_zero();
        float D = 0.f, C = 0.f;
        for (int blk = 0; blk < n_blocks; ++blk) {  // row blocks in order: fixed summation order
            const int64_t slot = (int64_t)blk * n_slots + i;
            const int s = __ldg(seg_lo + slot), e = __ldg(seg_hi + slot);
            if (e > s) {
                const float* r = R1 + slot * REC;
                const float4 a0 = __ldg(reinterpret_cast<const float4*>(r) + fq);
                float4 dc = __ldg(reinterpret_cast<const float4*>(r) + LPR);
                A.x += a0.x; A.y += a0.y; A.z += a0.z; A.w += a0.w;
                D += dc.x;
                C += dc.y;
                const int c1 = (e - 1) / CHB;
                for (int c = s / CHB + 1; c <= c1; ++c) {
                    const float* r2 = R2 + (int64_t)c * REC;
                    const float4 a = __ldg(reinterpret_cast<const float4*>(r2) + fq);
                    dc = __ldg(reinterpret_cast<const float4*>(r2) + LPR);
                    A.x += a.x; A.y += a.y; A.z += a.z; A.w += a.w;
                    D += dc.x;
                    C += dc.y;
                }
            }
        }
        if (BINARY) D = C;
        float4 g;
        g.x = __fmaf_rn(-v.x, D, A.x);
        g.y = __fmaf_rn(-v.y, D, A.y);
        g.z = __fmaf_rn(-v.z, D, A.z);
        g.w = __fmaf_rn(-v.w, D, A.w);
        if (FUSED) {
            if (active) {
                v.x = sgd_step(v.x, g.x, inv, up.eta, up.regv);
                v.y = sgd_step(v.y, g.y, inv, up.eta, up.regv);
                v.z = sgd_step(v.z, g.z, inv, up.eta, up.regv);
                v.w = sgd_step(v.w, g.w, inv, up.eta, up.regv);
                V4[i * LPR + fq] = v;
                if (fq == 0 && k1) W[i] = sgd_step(wi, C, inv, up.eta, up.regw);
            }
        } else {
            G4[i * LPR + fq] = g;
            if (fq == 0) Gw[i] = k1 ? C : 0.f;
        }
    }
}

template <int LPR>
static int64_t pull_chunks_for(int64_t nnz) {
    return (nnz + PullCfg<LPR>::CHB - 1) / PullCfg<LPR>::CHB;
}

// Row blocking of the sort key: rows per block = the largest power of two whose
// S rows fit in ~16 MB (so a block stays L2 resident while the reduce sweeps it); the number of
// blocks is capped so that the per-(block, feature) records stay within ~2 GB.
void pull_plan(const ModelView& m, int64_t n_rows, int* blk_shift, int* n_blocks) {
    const int block_mb = knobs().pull_block_mb;  // measured on C3: blocking does not pay (profiles/)
    if (block_mb <= 0) {
        *blk_shift = 30;
        *n_blocks = 1;
        return;
    }
    int shift = 10;
    while (shift < 30 && ((int64_t)2 << shift) * m.kp * 4 <= ((int64_t)block_mb << 20)) ++shift;
    int64_t nb = n_rows > 0 ? ((n_rows - 1) >> shift) + 1 : 1;
    const int64_t rec_bytes = (int64_t)(m.kp + 4) * 4 + 8;
    int64_t cap = ((int64_t)2 << 30) / (m.n_slots * rec_bytes);
    if (cap > 64) cap = 64;
    if (cap < 1) cap = 1;
    while (nb > cap) {
        ++shift;
        nb = ((n_rows - 1) >> shift) + 1;
    }
    *blk_shift = shift;
    *n_blocks = (int)nb;
}

size_t pull_scratch_bytes(const ModelView& m, int64_t nnz, int n_blocks) {
    const int threads = m.lpr >= 4 ? 256 : 64 * m.lpr;
    const int64_t chb = (threads / m.lpr) * PULL_SUB;
    const int64_t rec = m.kp + 4;
    return sizeof(float) * (size_t)rec * (size_t)(m.n_slots * n_blocks + (nnz + chb - 1) / chb + 1);
}

// One slice of the reduce: chunks [chunk_lo, chunk_hi) of the sorted entries (level 0/1), then the
// finalize of features [feat_lo, feat_hi) (level 2 + gradient / update).  The whole reduce is the
// slice {0, nchunks, 0, n_slots} with `first` set (zeroes the run bounds).
template <int LPR>
static cudaError_t pull_dispatch(const ModelView& m, int32_t* seg, int key_bits, int n_blocks,
                                 const uint32_t* keys, const uint2* pay, int64_t nnz, bool binary,
                                 const float* S, const float* mult, float* scratch,
                                 const double* d_scal, const int32_t* d_err, UpdateParams up,
                                 bool fused, float* grad, int sm_count, cudaStream_t st,
                                 const PullSlice& sl) {
    using Cfg = PullCfg<LPR>;
    float* R1 = scratch;
    const size_t nslot = (size_t)m.n_slots * n_blocks;
    float* R2 = scratch + nslot * Cfg::REC;
    int32_t* seg_lo = seg;
    int32_t* seg_hi = seg + nslot;
    if (sl.first) {
        cudaError_t e = cudaMemsetAsync(seg, 0, sizeof(int32_t) * 2 * nslot, st);
        if (e != cudaSuccess) return e;
    }
    const int64_t nchunks_all = pull_chunks_for<LPR>(nnz);
    const int64_t c_lo = sl.chunk_lo < 0 ? 0 : sl.chunk_lo;
    const int64_t c_hi = sl.chunk_hi < 0 || sl.chunk_hi > nchunks_all ? nchunks_all : sl.chunk_hi;
    if (c_hi > c_lo) {
        if (binary)
            fm_pull_chunks_kernel<LPR, true><<<(unsigned)(c_hi - c_lo), Cfg::THREADS, 0, st>>>(
                keys, pay, (const float4*)S, mult, (int)nnz, R1, R2, seg_lo, seg_hi, key_bits,
                m.n_slots, (int)c_lo);
        else
            fm_pull_chunks_kernel<LPR, false><<<(unsigned)(c_hi - c_lo), Cfg::THREADS, 0, st>>>(
                keys, pay, (const float4*)S, mult, (int)nnz, R1, R2, seg_lo, seg_hi, key_bits,
                m.n_slots, (int)c_lo);
    }
    const int64_t f_lo = sl.feat_lo < 0 ? 0 : sl.feat_lo;
    const int64_t f_hi = sl.feat_hi < 0 || sl.feat_hi > m.n_slots ? m.n_slots : sl.feat_hi;
    if (f_hi <= f_lo) return cudaGetLastError();
    const int64_t threads = (f_hi - f_lo) * LPR;
    int64_t blocks = (threads + 255) / 256;
    const int64_t cap = (int64_t)sm_count * 128;   // short dependent chains: favour parallelism
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    float* gv = grad;
    float* gw = grad ? grad + m.n_slots * m.kp : nullptr;
    float* gw0 = grad ? gw + m.n_slots : nullptr;
    if (sl.gv) {  // slice-major gradient layout: pointers pre-biased so that global indices work
        gv = sl.gv;
        gw = sl.gw;
        gw0 = sl.gw0;
    }
#define FIN_ARGS                                                                              \
    (float4*)m.v, m.w, m.w0, m.n_slots, m.k0, m.k1, seg_lo, seg_hi, R1, R2, n_blocks, d_scal, \
        d_err, up, (float4*)gv, gw, gw0, f_lo, f_hi
    const dim3 gd((unsigned)blocks), bd(256);
    if (fused) {
        if (binary) fm_pull_finalize_kernel<LPR, true, true><<<gd, bd, 0, st>>>(FIN_ARGS);
        else        fm_pull_finalize_kernel<LPR, true, false><<<gd, bd, 0, st>>>(FIN_ARGS);
    } else {
        if (binary) fm_pull_finalize_kernel<LPR, false, true><<<gd, bd, 0, st>>>(FIN_ARGS);
        else        fm_pull_finalize_kernel<LPR, false, false><<<gd, bd, 0, st>>>(FIN_ARGS);
    }
#undef FIN_ARGS
    return cudaGetLastError();
}

cudaError_t launch_pull_slice(const ModelView& m, int32_t* seg, int key_bits, int n_blocks,
                              const uint32_t* keys, const uint2* pay, int64_t nnz, bool binary,
                              const float* S, const float* mult, float* scratch,
                              const double* d_scal, const int32_t* d_err, UpdateParams up,
                              bool fused, float* grad, int sm_count, cudaStream_t st,
                              const PullSlice& sl, int64_t* launches) {
    *launches += 2;
#define PD(L) pull_dispatch<L>(m, seg, key_bits, n_blocks, keys, pay, nnz, binary, S, mult, scratch, d_scal, d_err, up, fused, grad, sm_count, st, sl)
    switch (m.lpr) {
        case 1: return PD(1);
        case 2: return PD(2);
        case 4: return PD(4);
        case 8: return PD(8);
        case 16: return PD(16);
        case 32: return PD(32);
    }
#undef PD
    return cudaErrorInvalidValue;
}

cudaError_t launch_pull(const ModelView& m, int32_t* seg, int key_bits, int n_blocks,
                        const uint32_t* keys, const uint2* pay, int64_t nnz, bool binary,
                        const float* S, const float* mult, float* scratch, const double* d_scal,
                        const int32_t* d_err, UpdateParams up, bool fused, float* grad,
                        int sm_count, cudaStream_t st, int64_t* launches) {
    PullSlice all;
    return launch_pull_slice(m, seg, key_bits, n_blocks, keys, pay, nnz, binary, S, mult, scratch,
                             d_scal, d_err, up, fused, grad, sm_count, st, all, launches);
}

int64_t pull_chunk_entries(const ModelView& m) {
    const int threads = m.lpr >= 4 ? 256 : 64 * m.lpr;
    return (int64_t)(threads / m.lpr) * PULL_SUB;
}

// pos[q] = first sorted position whose key >= q * n_slots / n_slices   (q = 1 .. n_slices-1)
__global__ void slice_bounds_kernel(const uint32_t* __restrict__ keys, int nnz, int n_slices,
                                    int64_t n_slots, int32_t* __restrict__ pos) {
    const int q = threadIdx.x + 1;
    if (q >= n_slices) return;
    const uint32_t target = (uint32_t)((int64_t)q * n_slots / n_slices);
    int lo = 0, hi = nnz;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (keys[mid] < target) lo = mid + 1; else hi = mid;
    }
    pos[q] = lo;
}

cudaError_t launch_slice_bounds(const uint32_t* keys, int64_t nnz, int n_slices, int64_t n_slots,
                                int32_t* d_pos, cudaStream_t st, int64_t* launches) {
    ++*launches;
    slice_bounds_kernel<<<1, 64, 0, st>>>(keys, (int)nnz, n_slices, n_slots, d_pos);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Dense update from the all-reduced gradient [gV | gw | gw0]  (DESIGN.md 2.3).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fm_update_kernel(float4* __restrict__ V4, float* __restrict__ W, float* __restrict__ W0,
                 int lpr, int64_t feat_lo, int64_t feat_hi, int k0, int k1,
                 const float4* __restrict__ G4, const float* __restrict__ Gw,
                 const float* __restrict__ Gw0, const double* __restrict__ d_scal,
                 const int32_t* __restrict__ err, UpdateParams up) {
    if (*err || d_scal[SC_ERR] != 0.0) return;   // some rank saw a bad index: nobody updates
    const double count = d_scal[SC_COUNT];
    if (!(count > 0.0)) return;
    const float inv = (float)(1.0 / count);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = feat_lo * lpr + tid; i < feat_hi * lpr; i += stride) {
        float4 v = V4[i];
        const float4 g = __ldg(G4 + i);
        v.x = sgd_step(v.x, g.x, inv, up.eta, up.regv);
        v.y = sgd_step(v.y, g.y, inv, up.eta, up.regv);
        v.z = sgd_step(v.z, g.z, inv, up.eta, up.regv);
        v.w = sgd_step(v.w, g.w, inv, up.eta, up.regv);
        V4[i] = v;
    }
    if (k1)
        for (int64_t i = feat_lo + tid; i < feat_hi; i += stride) {
            const float w = W[i];
            W[i] = sgd_step(w, __ldg(Gw + i), inv, up.eta, up.regw);
        }
    if (k0 && tid == 0 && feat_lo == 0) {
        const float w0 = *W0;
        *W0 = sgd_step(w0, *Gw0, inv, up.eta, up.reg0);
    }
}

cudaError_t launch_update_ptrs(const ModelView& m, const float* gv, const float* gw,
                               const float* gw0, const double* d_scal, const int32_t* d_err,
                               UpdateParams up, int64_t feat_lo, int64_t feat_hi, cudaStream_t st,
                               int64_t* launches) {
    ++*launches;
    const int64_t nv4 = (feat_hi - feat_lo) * m.lpr;
    int64_t blocks = (nv4 + 255) / 256;
    if (blocks > (int64_t)device_sm_count() * 16) blocks = (int64_t)device_sm_count() * 16;
    if (blocks < 1) blocks = 1;
    fm_update_kernel<<<(unsigned)blocks, 256, 0, st>>>((float4*)m.v, m.w, m.w0, m.lpr, feat_lo,
                                                       feat_hi, m.k0, m.k1, (const float4*)gv, gw,
                                                       gw0, d_scal, d_err, up);
    return cudaGetLastError();
}

cudaError_t launch_update_range(const ModelView& m, const float* grad, const double* d_scal,
                                const int32_t* d_err, UpdateParams up, int64_t feat_lo,
                                int64_t feat_hi, cudaStream_t st, int64_t* launches) {
    const float* gw = grad + m.n_slots * m.kp;
    return launch_update_ptrs(m, grad, gw, gw + m.n_slots, d_scal, d_err, up, feat_lo, feat_hi, st,
                              launches);
}

cudaError_t launch_update(const ModelView& m, const float* grad, const double* d_scal,
                          const int32_t* d_err, UpdateParams up, cudaStream_t st,
                          int64_t* launches) {
    return launch_update_range(m, grad, d_scal, d_err, up, 0, m.n_slots, st, launches);
}

// ------------------------------------------------------------------------------------------
// Entry list of a batch without the model: key = (row block << key_bits) | feature,
// payload = row (all-ones data) or {row, x}.  Used once per fixed mini-batch (PARTITION sampler).
// ------------------------------------------------------------------------------------------
template <bool HAS_VAL>
__global__ void __launch_bounds__(256)
emit_entries_kernel(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ idx,
                    const float* __restrict__ val, const int32_t* __restrict__ row_ids,
                    int64_t row_lo, int64_t n_rows, const int64_t* __restrict__ out_ptr,
                    int64_t out_base, int uniform_m, int key_bits, int blk_shift, int64_t n_slots,
                    uint32_t* __restrict__ keys, uint2* __restrict__ pay) {
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
        const uint32_t kpre = (uint32_t)(pos >> blk_shift) << key_bits;
        for (int64_t j = beg + lane; j < end; j += 32) {
            const int id = __ldg(idx + j);
            const bool ok = (uint32_t)id < (uint64_t)n_slots;  // validated at load; belt and braces
            keys[obase + (j - beg)] = kpre | (ok ? (uint32_t)id : 0u);
            if (HAS_VAL)
                pay[obase + (j - beg)] = make_uint2((uint32_t)pos, ok ? __float_as_uint(__ldg(val + j)) : 0u);
            else
                reinterpret_cast<uint32_t*>(pay)[obase + (j - beg)] = (uint32_t)pos;
        }
    }
}

cudaError_t launch_emit(const BatchView& b, int key_bits, int blk_shift, int64_t n_slots,
                        uint32_t* keys, uint2* pay, int sm_count, cudaStream_t st,
                        int64_t* launches) {
    if (b.n_rows <= 0) return cudaSuccess;
    ++*launches;
    int64_t blocks = (b.n_rows + 7) / 8;
    const int64_t cap = (int64_t)sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (b.val)
        emit_entries_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(
            b.row_ptr, b.idx, b.val, b.row_ids, b.row_lo, b.n_rows, b.out_ptr, b.out_base,
            b.uniform_m, key_bits, blk_shift, n_slots, keys, pay);
    else
        emit_entries_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(
            b.row_ptr, b.idx, b.val, b.row_ids, b.row_lo, b.n_rows, b.out_ptr, b.out_base,
            b.uniform_m, key_bits, blk_shift, n_slots, keys, pay);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Small utilities.
// ------------------------------------------------------------------------------------------
__global__ void row_lens_kernel(const int64_t* __restrict__ row_ptr,
                                const int32_t* __restrict__ row_ids, int64_t n,
                                int64_t* __restrict__ lens) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t r = row_ids[i];
        lens[i] = row_ptr[r + 1] - row_ptr[r];
    }
}

cudaError_t launch_row_lens(const int64_t* row_ptr, const int32_t* row_ids, int64_t n,
                            int64_t* lens, cudaStream_t st, int64_t* launches) {
    if (n <= 0) return cudaSuccess;
    ++*launches;
    int64_t blocks = (n + 255) / 256;
    if (blocks > (int64_t)device_sm_count() * 16) blocks = (int64_t)device_sm_count() * 16;
    row_lens_kernel<<<(unsigned)blocks, 256, 0, st>>>(row_ptr, row_ids, n, lens);
    return cudaGetLastError();
}

__global__ void idx_range_kernel(const int32_t* __restrict__ idx, int64_t nnz,
                                 int32_t* __restrict__ minmax) {
    int lo = INT32_MAX, hi = INT32_MIN;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride) {
        const int v = idx[i];
        lo = min(lo, v);
        hi = max(hi, v);
    }
    for (int off = 16; off > 0; off >>= 1) {
        lo = min(lo, __shfl_xor_sync(FULL, lo, off));
        hi = max(hi, __shfl_xor_sync(FULL, hi, off));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(minmax, lo);  // integer atomics: order-independent result
        atomicMax(minmax + 1, hi);
    }
}

cudaError_t launch_idx_range(const int32_t* idx, int64_t nnz, int32_t* d_minmax, cudaStream_t st,
                             int64_t* launches) {
    if (nnz <= 0) return cudaSuccess;
    ++*launches;
    int64_t blocks = (nnz + 255) / 256;
    if (blocks > (int64_t)device_sm_count() * 16) blocks = (int64_t)device_sm_count() * 16;
    idx_range_kernel<<<(unsigned)blocks, 256, 0, st>>>(idx, nnz, d_minmax);
    return cudaGetLastError();
}

// [n_slots][k] <-> [n_slots][kp] (zero padded)
// Unpacks the compact one-hot staging format (sfm_stage_onehot): entry e = bits
// [e*id_bits, (e+1)*id_bits) of the uint32 stream.  One thread per entry: consecutive threads read
// the same or neighbouring words (coalesced through L1) and write consecutive ids.
__global__ void __launch_bounds__(256)
unpack_onehot_kernel(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ label_bits,
                     int64_t n_entries, int64_t n_rows, int id_bits, uint32_t n_slots,
                     int32_t* __restrict__ idx, float* __restrict__ label, int32_t* __restrict__ bad) {
    const uint64_t mask = id_bits >= 32 ? 0xffffffffull : ((1ull << id_bits) - 1ull);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    bool any_bad = false;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_entries; e += stride) {
        const uint64_t bit = (uint64_t)e * (uint64_t)id_bits;
        const uint64_t wd = bit >> 5;
        const uint32_t sh = (uint32_t)(bit & 31u);
        const uint64_t two = (uint64_t)__ldg(packed + wd) | ((uint64_t)__ldg(packed + wd + 1) << 32);
        uint32_t id = (uint32_t)((two >> sh) & mask);
        if (id >= n_slots) {
            any_bad = true;
            id = 0u;
        }
        idx[e] = (int32_t)id;
    }
    if (label_bits)
        for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride)
            label[r] = ((__ldg(label_bits + (r >> 5)) >> (r & 31)) & 1u) ? 1.f : 0.f;
    if (any_bad) atomicExch(bad, 1);
}

cudaError_t launch_unpack_onehot(const uint32_t* packed, const uint32_t* label_bits,
                                 int64_t n_entries, int64_t n_rows, int id_bits, int64_t n_slots,
                                 int32_t* idx, float* label, int32_t* bad, cudaStream_t st,
                                 int64_t* launches) {
    if (n_entries <= 0 && n_rows <= 0) return cudaSuccess;
    int64_t blocks = ((n_entries > n_rows ? n_entries : n_rows) + 255) / 256;
    if (blocks > (int64_t)device_sm_count() * 32) blocks = (int64_t)device_sm_count() * 32;
    unpack_onehot_kernel<<<(unsigned)blocks, 256, 0, st>>>(packed, label_bits, n_entries, n_rows,
                                                           id_bits, (uint32_t)n_slots, idx, label, bad);
    *launches += 1;
    return cudaGetLastError();
}

__global__ void pad_v_kernel(const float* __restrict__ src, float* __restrict__ dst,
                             int64_t n_slots, int k, int kp, int unpad) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t total = n_slots * (unpad ? k : kp);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        if (unpad) {
            const int64_t i = e / k;
            const int f = (int)(e % k);
            dst[e] = src[i * kp + f];
        } else {
            const int64_t i = e / kp;
            const int f = (int)(e % kp);
            dst[e] = f < k ? src[i * k + f] : 0.f;
        }
    }
}

cudaError_t launch_pad_v(const float* src, float* dst, int64_t n_slots, int k, int kp, bool unpad,
                         cudaStream_t st, int64_t* launches) {
    const int64_t total = n_slots * (unpad ? k : kp);
    if (total <= 0) return cudaSuccess;
    ++*launches;
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)device_sm_count() * 16) blocks = (int64_t)device_sm_count() * 16;
    pad_v_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, dst, n_slots, k, kp, unpad ? 1 : 0);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Synthetic CTR rows (numpy twin: sparkfm_b200/synth.py ctr_rows).  Integer
// arithmetic only, so host and device agree bit for bit.
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t mix64_dev(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

constexpr uint64_t SALT_FEATURE = 0x5fe14bd7a3c1e5b3ULL;
constexpr uint64_t SALT_A = 0x1b873593cc9e2d51ULL;
constexpr uint64_t SALT_B = 0x85ebca6bc2b2ae35ULL;
constexpr uint64_t SALT_NOISE = 0x27d4eb2f165667c5ULL;
constexpr int64_t LABEL_THRESHOLD = 110000;  // ~25 % positives (synth.py)

__global__ void __launch_bounds__(256)
synth_ctr_kernel(int64_t n_rows, int64_t row_off, int n_fields,
                 const int32_t* __restrict__ log2card, const uint32_t* __restrict__ cdf,
                 const int64_t* __restrict__ cdf_off, uint64_t seed_key, int64_t n_slots,
                 int32_t* __restrict__ idx, float* __restrict__ label,
                 int64_t* __restrict__ row_ptr) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
        const uint64_t gr = (uint64_t)(row_off + r);
        int64_t lin = 0, sb = 0, sb2 = 0;
        for (int f = 0; f < n_fields; ++f) {
            const uint64_t h = mix64_dev(seed_key + gr * (uint64_t)n_fields + (uint64_t)f);
            const uint32_t u = (uint32_t)(h >> 32);
            const uint32_t* tab = cdf + cdf_off[f];
            int lo = 0, hi = (1 << log2card[f]) - 1;  // first j with u <= tab[j]
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (u <= tab[mid]) hi = mid; else lo = mid + 1;
            }
            const uint64_t fid =
                mix64_dev((((uint64_t)f) << 40) ^ (uint64_t)lo ^ SALT_FEATURE) % (uint64_t)n_slots;
            idx[r * n_fields + f] = (int32_t)fid;
            const int64_t a = (int64_t)(mix64_dev(fid ^ SALT_A) & 0xFFFF) - 32768;
            const int64_t b = (int64_t)(mix64_dev(fid ^ SALT_B) & 0xFF) - 128;
            lin += a;
            sb += b;
            sb2 += b * b;
        }
        const int64_t noise =
            (int64_t)(mix64_dev(seed_key ^ mix64_dev(gr ^ SALT_NOISE)) & 0x7FFFF) - 262144;
        const int64_t score = lin + (sb * sb - sb2) / 2 + noise;  // (sb^2 - sb2) is always even
        label[r] = score > LABEL_THRESHOLD ? 1.f : 0.f;
        row_ptr[r] = r * n_fields;
        if (r == n_rows - 1) row_ptr[n_rows] = n_rows * n_fields;
    }
}

cudaError_t launch_synth_ctr(int64_t n_rows, int64_t row_off, int n_fields,
                             const int32_t* d_log2card, const uint32_t* d_cdf,
                             const int64_t* d_cdf_off, uint64_t seed, int64_t n_slots,
                             int32_t* idx, float* label, int64_t* row_ptr, cudaStream_t st,
                             int64_t* launches) {
    if (n_rows <= 0) return cudaSuccess;
    ++*launches;
    int64_t blocks = (n_rows + 255) / 256;
    if (blocks > (int64_t)device_sm_count() * 16) blocks = (int64_t)device_sm_count() * 16;
    synth_ctr_kernel<<<(unsigned)blocks, 256, 0, st>>>(n_rows, row_off, n_fields, d_log2card,
                                                       d_cdf, d_cdf_off, mix64_dev(seed), n_slots,
                                                       idx, label, row_ptr);
    return cudaGetLastError();
}

}  // namespace sfm
