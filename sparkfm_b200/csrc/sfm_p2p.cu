// sfm_p2p.cu -- gradient sum + SGD update fused into ONE kernel over NVLink peer memory
// (replicated data-parallel mode; replaces ncclAllReduce(68 MB) + fm_update_kernel).
//
// Every rank keeps its dense gradient [gV | gw] in a buffer the other ranks can read
// (CUDA IPC, peer access over NVLink / NVSwitch) and its model replica in a buffer the others can
// write.  Rank r owns the feature slice [r*n/G, (r+1)*n/G).  One iteration:
//
//   1. finalize writes the local gradient, then a signal kernel copies the rank's loss / count /
//      gw0 sums next to it and stores the step number into every peer's "ready" word (release,
//      system scope) -- the scalar all-reduce of the NCCL path is folded into the same exchange;
//   2. p2p_reduce_update_kernel waits until all G ready words carry this step, then for its OWN
//      slice: loads the G partial gradients with peer loads (no local caching), adds them IN RANK
//      ORDER (bitwise reproducible and identical on every rank), applies
//      theta <- theta - eta*(g/B + lambda*theta) and stores the new parameters into ALL G replicas
//      (peer stores).  The last CTA to finish stores the step number into every peer's "done" word;
//   3. a one-thread kernel waits for the G done words before the next forward may read the model.
//
// Per rank and step (G-1)/G of the gradient slice comes in and (G-1)/G of the parameter slice
// goes out over NVLink, overlapped inside one kernel -- the reduce-scatter, the update and the
// all-gather of a ring all-reduce, without intermediate buffers.  All waits have a time-out that
// raises the handle's error flag instead of hanging.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
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

constexpr int P2P_MAXG = 16;
// done_ctr words: [0] CTA completion counter, [1] sticky "a peer wait timed out" flag

struct DevPeers {
    float4* v[P2P_MAXG];
    float* w[P2P_MAXG];
    const float4* g4[P2P_MAXG];
    const float* gw[P2P_MAXG];
    const double* scal[P2P_MAXG];   // [SC_N] per-rank loss / count / gw0 / error sums of this step
    const uint32_t* tb[P2P_MAXG];   // touched-feature bitmap of this step (sparse exchange)
    uint32_t* sig[P2P_MAXG];   // [2][P2P_MAXG]: ready words, done words
};

struct P2PState {
    float* grad = nullptr;
    uint32_t* touch = nullptr;   // inside the grad allocation
    uint32_t* tball = nullptr;   // [world][words] local copy of every rank's touched bitmap (sparse exchange)
    double* scal_all = nullptr;  // [world][SC_N] local copy of the ranks' scalar blocks
    uint32_t* sig = nullptr;
    uint32_t* done_ctr = nullptr;
    void* opened[P2P_MAXG][4];
    DevPeers peers;
    uint32_t epoch = 0;
    unsigned long long timeout_ns = 0;
    unsigned long long* trace = nullptr;   // SFM_P2P_TRACE=1: [8] summed stage times (ns) of CTA 0 + step count
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// one thread: wait until sig[base + p] has reached `epoch` for every rank p
__device__ bool p2p_wait(const uint32_t* sig, int base, int world, uint32_t epoch,
                         unsigned long long timeout_ns, uint32_t* timed_out) {
    if (ld_volatile_u32(timed_out)) return false;   // sticky: never wait again after a time-out
    const unsigned long long t0 = global_ns();
    for (int p = 0; p < world; ++p) {
        while ((int32_t)(ld_volatile_u32(sig + base + p) - epoch) < 0) {
            if (global_ns() - t0 > timeout_ns) {
                st_volatile_u32(timed_out, 1u);
                return false;
            }
            __nanosleep(200);
        }
    }
    __threadfence_system();
    return true;
}

// The same wait with one polling thread per rank (the flags are G dependent L2 round trips when
// one thread reads them in turn).  Every thread of the block must call it.
__device__ bool p2p_wait_block(const uint32_t* sig, int base, int world, uint32_t epoch,
                               unsigned long long timeout_ns, uint32_t* timed_out) {
    bool ok = true;
    if ((int)threadIdx.x < world) {
        if (ld_volatile_u32(timed_out)) {
            ok = false;
        } else {
            const unsigned long long t0 = global_ns();
            while ((int32_t)(ld_volatile_u32(sig + base + threadIdx.x) - epoch) < 0) {
                if (global_ns() - t0 > timeout_ns) {
                    st_volatile_u32(timed_out, 1u);
                    ok = false;
                    break;
                }
                __nanosleep(100);
            }
        }
        __threadfence_system();
    }
    return __syncthreads_and(ok) != 0;
}

__global__ void p2p_signal_kernel(DevPeers P, int world, int rank, int phase, uint32_t epoch,
                                  const double* __restrict__ d_scal) {
    const int p = threadIdx.x;
    if (p < SC_N) const_cast<double*>(P.scal[rank])[p] = d_scal[p];
    __syncthreads();
    if (p < world) {
        __threadfence_system();
        st_volatile_u32(P.sig[p] + phase * P2P_MAXG + rank, epoch);
    }
}

__global__ void p2p_wait_kernel(const uint32_t* sig, int phase, int world, uint32_t epoch,
                                unsigned long long timeout_ns, uint32_t* timed_out) {
    p2p_wait_block(sig, phase * P2P_MAXG, world, epoch, timeout_ns, timed_out);
}

__device__ __forceinline__ float4 ld_peer4(const float4* p) {   // no caching of peer data
    float4 r;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ float ld_peer1(const float* p) {
    float r;
    asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// The G ranks' scalar blocks of this step -> tot[SC_N], added in rank order (fp64: the same on
// every rank).  One peer load per thread, all in flight together: ONE NVLink round trip instead
// of G * SC_N dependent ones (the round-1 form cost ~0.1 ms at 8 ranks).
__device__ __forceinline__ void p2p_gather_scalars(const DevPeers& P, int world, bool ok,
                                                    double (*sc)[SC_N], double* tot) {
    const int t = threadIdx.x;
    if (t < world * SC_N) {
        const int p = t / SC_N, j = t % SC_N;
        double x = 0.0;
        if (ok) asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(x) : "l"(P.scal[p] + j));
        sc[p][j] = x;
    }
    __syncthreads();
    if (t < SC_N) {
        double a = 0.0;
        for (int p = 0; p < world; ++p) a += sc[p][t];
        tot[t] = a;
    }
    __syncthreads();
}

template <int W>   // W = number of ranks rounded up to 2 / 4 / 8 / 16 (loads are issued together)
__global__ void __launch_bounds__(256)
p2p_reduce_update_kernel(DevPeers P, int world, int rank, uint32_t epoch, int64_t v4_lo,
                         int64_t v4_hi, int64_t w_lo, int64_t w_hi, int k0, int k1,
                         float* __restrict__ W0, const uint32_t* sig, double* __restrict__ d_scal,
                         const int32_t* __restrict__ err, UpdateParams up,
                         unsigned long long timeout_ns, uint32_t* done_ctr) {
    __shared__ int ok;
    __shared__ double tot[SC_N];
    __shared__ double sc[P2P_MAXG][SC_N];
    {
        const bool w_ok = p2p_wait_block(sig, 0, world, epoch, timeout_ns, done_ctr + 1);
        if (threadIdx.x == 0) ok = w_ok ? 1 : 0;
        __syncthreads();
    }
    p2p_gather_scalars(P, world, ok != 0, sc, tot);   // global loss / count / gw0 / error
    const double count = tot[SC_COUNT];
    // tot[SC_ERR] = number of ranks that saw a bad index: everybody skips the update together
    const bool active = ok && count > 0.0 && tot[SC_ERR] == 0.0;
    if (ok && blockIdx.x == 0 && threadIdx.x < SC_N) d_scal[threadIdx.x] = tot[threadIdx.x];
    if (active) {
        const float inv = (float)(1.0 / count);
        const int64_t stride = (int64_t)gridDim.x * blockDim.x;
        const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (int64_t e = v4_lo + tid; e < v4_hi; e += stride) {
            float4 x[W];
#pragma unroll
            for (int p = 0; p < W; ++p)
                x[p] = p < world ? ld_peer4(P.g4[p] + e) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 g = x[0];
#pragma unroll
            for (int p = 1; p < W; ++p) {   // rank order: fixed summation order
                g.x += x[p].x; g.y += x[p].y; g.z += x[p].z; g.w += x[p].w;
            }
            float4 v = P.v[rank][e];
            v.x = sgd_step(v.x, g.x, inv, up.eta, up.regv);
            v.y = sgd_step(v.y, g.y, inv, up.eta, up.regv);
            v.z = sgd_step(v.z, g.z, inv, up.eta, up.regv);
            v.w = sgd_step(v.w, g.w, inv, up.eta, up.regv);
#pragma unroll
            for (int p = 0; p < W; ++p)
                if (p < world) P.v[p][e] = v;
        }
        if (k1)
            for (int64_t i = w_lo + tid; i < w_hi; i += stride) {
                float g = 0.f;
#pragma unroll
                for (int p = 0; p < W; ++p) g += p < world ? ld_peer1(P.gw[p] + i) : 0.f;
                const float w = P.w[rank][i];
                const float wn = sgd_step(w, g, inv, up.eta, up.regw);
#pragma unroll
                for (int p = 0; p < W; ++p)
                    if (p < world) P.w[p][i] = wn;
            }
        if (k0 && tid == 0) {   // w0 is replicated: every rank applies the same global scalar
            const float w0 = *W0;
            *W0 = sgd_step(w0, (float)tot[SC_GW0], inv, up.eta, up.reg0);
        }
    }
    // completion: the last CTA tells every peer that this slice has been written everywhere
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd(done_ctr, 1u);
        if (ticket == gridDim.x - 1) {
            *done_ctr = 0;
            __threadfence_system();
            for (int p = 0; p < world; ++p) st_volatile_u32(P.sig[p] + P2P_MAXG + rank, epoch);
        }
    }
}

// The ranks' touched bitmaps -> one local array tball[rank][word], with coalesced 16-byte peer
// loads (a bitmap is 125 KB at 1 M features).  The update kernel then reads bitmaps locally: one
// 4-byte peer load per (warp, rank, word) -- what the first version did -- is one NVLink request
// each, and their number, not their bytes, was what an exchange at 8 ranks cost.
__global__ void __launch_bounds__(256)
p2p_bitmap_gather_kernel(DevPeers P, int world, uint32_t epoch, int words, const uint32_t* sig,
                         uint32_t* __restrict__ tball, double* __restrict__ scal_all,
                         unsigned long long timeout_ns, uint32_t* done_ctr) {
    if (!p2p_wait_block(sig, 0, world, epoch, timeout_ns, done_ctr + 1)) return;
    if (blockIdx.x == 0 && (int)threadIdx.x < world * SC_N) {   // the ranks' scalar blocks, too
        const int pp = threadIdx.x / SC_N, j = threadIdx.x % SC_N;
        const double* src = nullptr;
#pragma unroll
        for (int r = 0; r < P2P_MAXG; ++r)
            if (r == pp) src = P.scal[r];
        double x;
        asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(x) : "l"(src + j));
        scal_all[pp * SC_N + j] = x;
    }
    const int quads = words / 4;   // the bitmap is padded to a multiple of 4 words
    const int64_t total = (int64_t)world * quads;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int p = (int)(i / quads), q = (int)(i % quads);
        const uint32_t* src = nullptr;
#pragma unroll
        for (int r = 0; r < P2P_MAXG; ++r)
            if (r == p) src = P.tb[r];
        uint4 v;
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(reinterpret_cast<const uint4*>(src) + q));
        reinterpret_cast<uint4*>(tball + (size_t)p * words)[q] = v;
    }
}

// Sparse form of the same exchange: every rank publishes one bit per feature its batch touched
// (written by bkt_pull_kernel<MODE 2> next to the gradient rows of exactly those features).
//   part 1, own slice of bitmap words: for every feature some rank touched, the touching ranks'
//           rows are loaded (peer loads), added in rank order, the update is applied and the new
//           row is stored into all G replicas;
//   part 2, ALL features: a feature nobody touched only receives its L2 decay, which every rank
//           applies to its own replica -- no NVLink traffic (skipped when lambda_w = lambda_V = 0).
// Same bits as the dense kernel (an untouched rank contributes +0 there) up to the sign of a zero.
template <int W>
__global__ void __launch_bounds__(256)
p2p_sparse_update_kernel(DevPeers P, int world, int rank, uint32_t epoch, int64_t n_slots, int lsh,
                         int wd_lo, int wd_hi, int wd_total, int k0, int k1,
                         float* __restrict__ W0, const uint32_t* sig, double* __restrict__ d_scal,
                         UpdateParams up, unsigned long long timeout_ns, uint32_t* done_ctr,
                         unsigned long long* __restrict__ trace, const uint32_t* __restrict__ tball,
                         int tb_words, const double* __restrict__ scal_all) {
    __shared__ int ok;
    __shared__ double tot[SC_N];
    __shared__ double sc[P2P_MAXG][SC_N];
    const bool tr = trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    unsigned long long ts[6];
    if (tr) ts[0] = global_ns();
    {   // (the flags are already up: the bitmap kernel in front of this one waited for them)
        const bool w_ok = p2p_wait_block(sig, 0, world, epoch, timeout_ns, done_ctr + 1);
        if (threadIdx.x == 0) ok = w_ok ? 1 : 0;
    }
    if (tr) ts[1] = global_ns();
    {   // global loss / count / gw0 / error from the local copy of the ranks' scalar blocks
        const int t = threadIdx.x;
        if (t < world * SC_N) sc[t / SC_N][t % SC_N] = __ldcg(scal_all + t);
        __syncthreads();
        if (t < SC_N) {
            double acc = 0.0;
            for (int p = 0; p < world; ++p) acc += sc[p][t];   // rank order
            tot[t] = acc;
        }
        __syncthreads();
    }
    if (tr) ts[2] = global_ns();
    const double count = tot[SC_COUNT];
    const bool active = ok && count > 0.0 && tot[SC_ERR] == 0.0;
    if (ok && blockIdx.x == 0 && threadIdx.x < SC_N) d_scal[threadIdx.x] = tot[threadIdx.x];
    if (active) {
        const float inv = (float)(1.0 / count);
        const int lane = threadIdx.x & 31;
        const int lpr = 1 << lsh, fpp = 32 >> lsh;   // float4 per row, features per pass
        const int fq = lane & (lpr - 1), jl = lane >> lsh;
        const int warps = (gridDim.x * blockDim.x) >> 5;
        const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        float4* Vme = P.v[rank];
        float* Wme = P.w[rank];
        for (int wd = wd_lo + gw; wd < wd_hi; wd += warps) {
            uint32_t mine = 0;
            if (lane < world) mine = __ldcg(tball + (size_t)lane * tb_words + wd);   // local copy
            uint32_t bits[W];
            uint32_t uni = 0;
#pragma unroll
            for (int p = 0; p < W; ++p) {
                bits[p] = __shfl_sync(0xffffffffu, mine, p);
                uni |= bits[p];
            }
            if (!uni) continue;
            if (k1) {   // w: lane <-> feature of the word, so a rank's gw values / the new w values
                        // of the word travel as ONE request per rank instead of one per feature
                const int64_t f = (int64_t)wd * 32 + lane;
                const bool in = f < n_slots;
                float gs = 0.f;
#pragma unroll
                for (int p = 0; p < W; ++p)   // rank order: fixed summation order
                    gs += (in && ((bits[p] >> lane) & 1u)) ? ld_peer1(P.gw[p] + f) : 0.f;
                if (in && ((uni >> lane) & 1u)) {
                    const float wn = sgd_step(Wme[f], gs, inv, up.eta, up.regw);
#pragma unroll
                    for (int p = 0; p < W; ++p)
                        if (p < world) P.w[p][f] = wn;
                }
            }
            for (int pass = 0; pass < lpr; ++pass) {
                const int j = pass * fpp + jl;
                const int64_t f = (int64_t)wd * 32 + j;
                if (!((uni >> j) & 1u) || f >= n_slots) continue;
                const int64_t e = (f << lsh) + fq;
                float4 x[W];
#pragma unroll
                for (int p = 0; p < W; ++p)
                    x[p] = ((bits[p] >> j) & 1u) ? ld_peer4(P.g4[p] + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 g = x[0];
#pragma unroll
                for (int p = 1; p < W; ++p) {   // rank order: fixed summation order
                    g.x += x[p].x; g.y += x[p].y; g.z += x[p].z; g.w += x[p].w;
                }
                float4 v = Vme[e];
                v.x = sgd_step(v.x, g.x, inv, up.eta, up.regv);
                v.y = sgd_step(v.y, g.y, inv, up.eta, up.regv);
                v.z = sgd_step(v.z, g.z, inv, up.eta, up.regv);
                v.w = sgd_step(v.w, g.w, inv, up.eta, up.regv);
#pragma unroll
                for (int p = 0; p < W; ++p)
                    if (p < world) P.v[p][e] = v;
            }
        }
        if (tr) ts[3] = global_ns();
        if (up.regv != 0.f || (k1 && up.regw != 0.f)) {
            for (int wd = gw; wd < wd_total; wd += warps) {
                uint32_t mine = 0;
                if (lane < world) mine = __ldcg(tball + (size_t)lane * tb_words + wd);
                const uint32_t uni = __reduce_or_sync(0xffffffffu, mine);
                if (uni == 0xffffffffu) continue;
                if (k1 && up.regw != 0.f) {
                    const int64_t f = (int64_t)wd * 32 + lane;
                    if (f < n_slots && !((uni >> lane) & 1u)) Wme[f] = sgd_step(Wme[f], 0.f, inv, up.eta, up.regw);
                }
                if (up.regv == 0.f) continue;
                for (int pass = 0; pass < lpr; ++pass) {
                    const int j = pass * fpp + jl;
                    const int64_t f = (int64_t)wd * 32 + j;
                    if (((uni >> j) & 1u) || f >= n_slots) continue;
                    const int64_t e = (f << lsh) + fq;
                    float4 v = Vme[e];
                    v.x = sgd_step(v.x, 0.f, inv, up.eta, up.regv);
                    v.y = sgd_step(v.y, 0.f, inv, up.eta, up.regv);
                    v.z = sgd_step(v.z, 0.f, inv, up.eta, up.regv);
                    v.w = sgd_step(v.w, 0.f, inv, up.eta, up.regv);
                    Vme[e] = v;
                }
            }
        }
        if (k0 && blockIdx.x == 0 && threadIdx.x == 0) {
            const float w0 = *W0;
            *W0 = sgd_step(w0, (float)tot[SC_GW0], inv, up.eta, up.reg0);
        }
        if (tr) ts[4] = global_ns();
    }
    __threadfence_system();
    __syncthreads();
    if (tr && active) {   // stage times of CTA 0: wait for the ranks, scalars, touched rows, decay, fence
        ts[5] = global_ns();
        for (int j = 0; j < 5; ++j) atomicAdd(trace + j, ts[j + 1] - ts[j]);
        atomicAdd(trace + 5, 1ull);
    }
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd(done_ctr, 1u);
        if (ticket == gridDim.x - 1) {
            *done_ctr = 0;
            __threadfence_system();
            for (int p = 0; p < world; ++p) st_volatile_u32(P.sig[p] + P2P_MAXG + rank, epoch);
        }
    }
}

// Called from sfm_destroy before the model buffers are freed.  Every rank has seen every peer's
// "done" word of the last step by then (the step's wait kernel), so no peer still reads or writes
// this rank's buffers; the imported mappings are closed first, the own allocations freed after.
// Hosts are expected to destroy the handles of all ranks together, as they create them.
void p2p_teardown(sfm_handle* h) {
    P2PState* s = h->p2p;
    if (!s) return;
    if (s->trace) {
        unsigned long long t[8] = {0};
        cudaStreamSynchronize(h->stream);
        if (cudaMemcpy(t, s->trace, sizeof t, cudaMemcpyDeviceToHost) == cudaSuccess && t[5] > 0) {
            const double n = (double)t[5] * 1e3;
            fprintf(stderr, "[sfm p2p trace] rank %d/%d, %llu sparse exchanges, CTA 0 avg us: wait for ranks %.1f, "
                    "scalars %.1f, touched rows (part 1) %.1f, local decay (part 2) %.1f, fence %.1f\n",
                    h->rank, h->world, t[5], t[0] / n, t[1] / n, t[2] / n, t[3] / n, t[4] / n);
        }
        cudaFree(s->trace);
        cudaGetLastError();
    }
    for (int p = 0; p < h->world && p < P2P_MAXG; ++p)
        for (int j = 0; j < 4; ++j)
            if (s->opened[p][j]) cudaIpcCloseMemHandle(s->opened[p][j]);
    if (s->grad) cudaFree(s->grad);
    if (s->tball) cudaFree(s->tball);
    if (s->scal_all) cudaFree(s->scal_all);
    if (s->sig) cudaFree(s->sig);
    if (s->done_ctr) cudaFree(s->done_ctr);
    delete s;
    h->p2p = nullptr;
}

// Collective.  Enables the P2P path if EVERY rank could export its buffers and open all peers'.
int p2p_setup(sfm_handle* h) {
    const char* env = getenv("SFM_P2P");
    const int want = env ? atoi(env) : 1;
    const int G = h->world;
    if (G < 2 || G > P2P_MAXG || h->shard_requested) return SFM_OK;
    const ModelView& m = h->m;
    P2PState* s = new (std::nothrow) P2PState;
    if (!s) return SFM_OK;
    memset(s->opened, 0, sizeof s->opened);
    memset(&s->peers, 0, sizeof s->peers);
    const size_t glen = (size_t)m.n_slots * (m.kp + 1) + 1;
    const size_t scal_off = (glen + 3) / 4 * 4;   // floats; the [SC_N] doubles sit 16-byte aligned
    const size_t touch_off = scal_off + 2 * SC_N; // floats; then one bit per feature
    const size_t touch_words = ((size_t)(m.n_slots + 31) / 32 + 3) / 4 * 4;   // whole 16-byte quads
    bool ok = want != 0;
    cudaIpcMemHandle_t mine[4];
    memset(mine, 0, sizeof mine);
    if (ok) ok = cudaMalloc(&s->grad, sizeof(float) * (touch_off + touch_words)) == cudaSuccess;
    if (ok) s->touch = reinterpret_cast<uint32_t*>(s->grad + touch_off);
    if (ok) ok = cudaMalloc(&s->tball, sizeof(uint32_t) * touch_words * (size_t)G) == cudaSuccess;
    if (ok) ok = cudaMalloc(&s->scal_all, sizeof(double) * SC_N * P2P_MAXG) == cudaSuccess;
    if (ok) cudaMemsetAsync(s->grad + touch_off, 0, sizeof(uint32_t) * touch_words, h->stream);   // pad words stay 0
    if (ok) ok = cudaMalloc(&s->sig, sizeof(uint32_t) * 2 * P2P_MAXG) == cudaSuccess;
    if (ok) ok = cudaMalloc(&s->done_ctr, sizeof(uint32_t) * 4) == cudaSuccess;
    if (ok && getenv("SFM_P2P_TRACE")) {
        if (cudaMalloc(&s->trace, sizeof(unsigned long long) * 8) == cudaSuccess)
            cudaMemsetAsync(s->trace, 0, sizeof(unsigned long long) * 8, h->stream);
        else
            s->trace = nullptr;
        cudaGetLastError();
    }
    if (ok) {
        cudaMemsetAsync(s->sig, 0, sizeof(uint32_t) * 2 * P2P_MAXG, h->stream);
        cudaMemsetAsync(s->done_ctr, 0, sizeof(uint32_t) * 4, h->stream);
        ok = cudaIpcGetMemHandle(&mine[0], m.v) == cudaSuccess &&
             cudaIpcGetMemHandle(&mine[1], m.w) == cudaSuccess &&
             cudaIpcGetMemHandle(&mine[2], s->grad) == cudaSuccess &&
             cudaIpcGetMemHandle(&mine[3], s->sig) == cudaSuccess;
    }
    cudaGetLastError();
    // exchange {ok, 4 handles} with everybody: 4 * 64 bytes + 1 word per rank
    constexpr int HW = 4 * (int)(sizeof(cudaIpcMemHandle_t) / 4);
    constexpr int WORDS = 1 + HW + 4;   // ok | 4 handles | device uuid
    std::vector<int32_t> send(WORDS), recv((size_t)WORDS * G);
    send[0] = ok ? 1 : 0;
    memcpy(send.data() + 1, mine, sizeof mine);
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, h->device) == cudaSuccess)
            memcpy(send.data() + 1 + HW, &prop.uuid, 16);
        else
            send[0] = 0;
        cudaGetLastError();
    }
    const char* tenv = getenv("SFM_P2P_TIMEOUT_S");
    const double tsec = tenv ? atof(tenv) : 120.0;
    s->timeout_ns = (unsigned long long)((tsec > 0.001 ? tsec : 0.001) * 1e9);
    Buf dsend, drecv;
    int rc = ensure(h, dsend, sizeof(int32_t) * WORDS);
    if (rc == SFM_OK) rc = ensure(h, drecv, sizeof(int32_t) * WORDS * G);
    if (rc == SFM_OK &&
        cudaMemcpyAsync(dsend.p, send.data(), sizeof(int32_t) * WORDS, cudaMemcpyHostToDevice,
                        h->stream) != cudaSuccess)
        rc = SFM_ERR_CUDA;
    if (rc == SFM_OK)
        rc = nccl_allgather_i32(h->nccl, h->comm, (const int32_t*)dsend.p, (int32_t*)drecv.p, WORDS,
                                h->stream, &h->err);
    if (rc == SFM_OK &&
        (cudaMemcpyAsync(recv.data(), drecv.p, sizeof(int32_t) * WORDS * G, cudaMemcpyDeviceToHost,
                         h->stream) != cudaSuccess ||
         cudaStreamSynchronize(h->stream) != cudaSuccess))
        rc = SFM_ERR_CUDA;
    if (dsend.p) cudaFree(dsend.p);
    if (drecv.p) cudaFree(drecv.p);
    if (rc != SFM_OK) {   // the all-gather itself failed: nothing more can be agreed on
        h->p2p = s;
        p2p_teardown(h);
        return rc;
    }
    bool all_ok = true;
    for (int p = 0; p < G; ++p) all_ok = all_ok && recv[(size_t)p * WORDS] == 1;
    // two ranks on one GPU could starve each other's wait loops: NCCL path for those
    for (int p = 0; p < G && all_ok; ++p)
        for (int q = 0; q < p; ++q)
            if (!memcmp(&recv[(size_t)p * WORDS + 1 + HW], &recv[(size_t)q * WORDS + 1 + HW], 16))
                all_ok = false;
    bool opened_ok = all_ok;
    if (all_ok) {
        for (int p = 0; p < G && opened_ok; ++p) {
            void* ptr[4];
            if (p == h->rank) {
                ptr[0] = m.v; ptr[1] = m.w; ptr[2] = s->grad; ptr[3] = s->sig;
            } else {
                cudaIpcMemHandle_t hd[4];
                memcpy(hd, recv.data() + (size_t)p * WORDS + 1, sizeof hd);
                for (int j = 0; j < 4 && opened_ok; ++j) {
                    if (cudaIpcOpenMemHandle(&ptr[j], hd[j], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                        cudaGetLastError();
                        opened_ok = false;
                    } else {
                        s->opened[p][j] = ptr[j];
                    }
                }
            }
            if (!opened_ok) break;
            s->peers.v[p] = (float4*)ptr[0];
            s->peers.w[p] = (float*)ptr[1];
            s->peers.g4[p] = (const float4*)ptr[2];
            s->peers.gw[p] = (const float*)ptr[2] + (size_t)m.n_slots * m.kp;
            s->peers.scal[p] = (const double*)((const float*)ptr[2] + scal_off);
            s->peers.tb[p] = (const uint32_t*)((const float*)ptr[2] + touch_off);
            s->peers.sig[p] = (uint32_t*)ptr[3];
        }
    }
    // second agreement: did everybody manage to open everything?
    double flag = opened_ok ? 0.0 : 1.0;
    double* dflag = h->d_scal + 6;
    int rc2 = SFM_OK;
    if (cudaMemcpyAsync(dflag, &flag, sizeof(double), cudaMemcpyHostToDevice, h->stream) != cudaSuccess)
        rc2 = SFM_ERR_CUDA;
    if (rc2 == SFM_OK) rc2 = nccl_allreduce_f64(h->nccl, h->comm, dflag, 1, h->stream, &h->err);
    if (rc2 == SFM_OK &&
        (cudaMemcpyAsync(&flag, dflag, sizeof(double), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
         cudaStreamSynchronize(h->stream) != cudaSuccess))
        rc2 = SFM_ERR_CUDA;
    h->p2p = s;
    if (rc2 != SFM_OK || flag != 0.0) {
        p2p_teardown(h);   // fall back to the NCCL all-reduce path on every rank
        return rc2;
    }
    return SFM_OK;
}

float* p2p_grad_buffer(sfm_handle* h) { return h->p2p ? h->p2p->grad : nullptr; }
uint32_t* p2p_touch_bits(sfm_handle* h) { return h->p2p ? h->p2p->touch : nullptr; }
const int32_t* p2p_timeout_flag(sfm_handle* h) {
    return h->p2p ? (const int32_t*)(h->p2p->done_ctr + 1) : nullptr;
}

// Steps 1-3 above, queued on the compute stream after the reduce kernel wrote p2p grad.
// sparse: the reduce wrote touched rows + the touched bitmap only (bkt_pull_kernel<MODE 2>).
int p2p_reduce_update(sfm_handle* h, UpdateParams up, bool sparse) {
    P2PState* s = h->p2p;
    const ModelView& m = h->m;
    const int G = h->world, r = h->rank;
    int64_t* L = &h->stats.kernel_launches;
    const uint32_t epoch = ++s->epoch;
    p2p_signal_kernel<<<1, 32, 0, h->stream>>>(s->peers, G, r, 0, epoch, h->d_scal);
    if (sparse) {
        const int64_t words = (m.n_slots + 31) / 32;
        const int tb_words = (int)((words + 3) / 4 * 4);
        {   // the ranks' bitmaps -> local copy (coalesced peer loads)
            int64_t gb = ((int64_t)G * (tb_words / 4) + 255) / 256;
            if (gb > (int64_t)h->sm_count * 4) gb = (int64_t)h->sm_count * 4;
            if (gb < 1) gb = 1;
            p2p_bitmap_gather_kernel<<<(unsigned)gb, 256, 0, h->stream>>>(s->peers, G, epoch, tb_words, s->sig,
                                                                       s->tball, s->scal_all, s->timeout_ns,
                                                                       s->done_ctr);
            *L += 1;
        }
        const int wd_lo = (int)((int64_t)r * words / G), wd_hi = (int)((int64_t)(r + 1) * words / G);
        int lsh = 0;
        while ((1 << lsh) < m.lpr) ++lsh;
        int64_t blocks = (words + 7) / 8;   // one warp per word at most
        const int64_t cap = (int64_t)h->sm_count * 8;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
#define SP_ARGS                                                                                \
    s->peers, G, r, epoch, m.n_slots, lsh, wd_lo, wd_hi, (int)words, m.k0, m.k1, m.w0, s->sig, \
        h->d_scal, up, s->timeout_ns, s->done_ctr, s->trace, s->tball, tb_words, s->scal_all
        if (G <= 2)      p2p_sparse_update_kernel<2><<<(unsigned)blocks, 256, 0, h->stream>>>(SP_ARGS);
        else if (G <= 4) p2p_sparse_update_kernel<4><<<(unsigned)blocks, 256, 0, h->stream>>>(SP_ARGS);
        else if (G <= 8) p2p_sparse_update_kernel<8><<<(unsigned)blocks, 256, 0, h->stream>>>(SP_ARGS);
        else             p2p_sparse_update_kernel<16><<<(unsigned)blocks, 256, 0, h->stream>>>(SP_ARGS);
#undef SP_ARGS
    } else {
        const int64_t f_lo = (int64_t)r * m.n_slots / G, f_hi = (int64_t)(r + 1) * m.n_slots / G;
        int64_t blocks = ((f_hi - f_lo) * m.lpr + 255) / 256;
        const int64_t cap = (int64_t)h->sm_count * 8;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
#define RU_ARGS                                                                                \
    s->peers, G, r, epoch, f_lo * m.lpr, f_hi * m.lpr, f_lo, f_hi, m.k0, m.k1, m.w0,          \
        s->sig, h->d_scal, h->d_err, up, s->timeout_ns, s->done_ctr
        if (G <= 2)      p2p_reduce_update_kernel<2><<<(unsigned)blocks, 256, 0, h->stream>>>(RU_ARGS);
        else if (G <= 4) p2p_reduce_update_kernel<4><<<(unsigned)blocks, 256, 0, h->stream>>>(RU_ARGS);
        else if (G <= 8) p2p_reduce_update_kernel<8><<<(unsigned)blocks, 256, 0, h->stream>>>(RU_ARGS);
        else             p2p_reduce_update_kernel<16><<<(unsigned)blocks, 256, 0, h->stream>>>(RU_ARGS);
#undef RU_ARGS
    }
    p2p_wait_kernel<<<1, 32, 0, h->stream>>>(s->sig, 1, G, epoch, s->timeout_ns, s->done_ctr + 1);
    *L += 3;
    CU(cudaGetLastError());
    return SFM_OK;
}

}  // namespace sfm
