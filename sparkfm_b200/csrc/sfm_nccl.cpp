// sfm_nccl.cpp -- NCCL bound at run time with dlopen, so libsparkfm_b200.so loads on machines
// without NCCL (and, inside a torch process, shares the libnccl.so.2 torch already mapped).
// The collective replaces Spark's driver-side combination of partition results
// (Model.scala:14 `.sum()`, fm/lib/ALS.scala:153 `.reduce(_+_)`; treeAggregate in north_star).
#include <dlfcn.h>
#include <string.h>

#include <string>

#include "sfm_common.h"

namespace sfm {

// Minimal NCCL declarations (ABI-stable across NCCL 2.x).
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt32 = 2, ncclFloat32 = 7, ncclFloat64 = 8 };  // ncclDataType_t
enum { ncclSum = 0 };                      // ncclRedOp_t

struct Nccl {
    void* dl = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*CommGetAsyncError)(ncclComm_t, int*) = nullptr;   // optional
    int (*CommAbort)(ncclComm_t) = nullptr;                 // optional
};

static std::string nerr(Nccl* n, const char* what, int rc) {
    std::string s = what;
    s += ": ";
    s += n->GetErrorString ? n->GetErrorString(rc) : "nccl error";
    return s;
}

Nccl* nccl_load(std::string* err) {
    static Nccl* inst = nullptr;
    if (inst) return inst;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* dl = nullptr;
    for (const char* nm : names) {
        dl = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (dl) break;
    }
    if (!dl) {
        if (err) *err = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
        return nullptr;
    }
    Nccl* n = new Nccl;
    n->dl = dl;
#define SYM(field, name)                                                  \
    *(void**)(&n->field) = dlsym(dl, name);                               \
    if (!n->field) {                                                      \
        if (err) *err = std::string("libnccl is missing symbol ") + name; \
        delete n;                                                         \
        return nullptr;                                                   \
    }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce")
    SYM(Broadcast, "ncclBroadcast")
    SYM(AllGather, "ncclAllGather")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    *(void**)(&n->CommGetAsyncError) = dlsym(dl, "ncclCommGetAsyncError");
    *(void**)(&n->CommAbort) = dlsym(dl, "ncclCommAbort");
    inst = n;
    return n;
}

int nccl_unique_id(Nccl* n, uint8_t* id128, std::string* err) {
    ncclUniqueId id;
    const int rc = n->GetUniqueId(&id);
    if (rc != ncclSuccess) {
        if (err) *err = nerr(n, "ncclGetUniqueId", rc);
        return SFM_ERR_NCCL;
    }
    memcpy(id128, id.internal, 128);
    return SFM_OK;
}

int nccl_init(Nccl* n, void** comm, const uint8_t* id128, int rank, int world, std::string* err) {
    ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    ncclComm_t c = nullptr;
    const int rc = n->CommInitRank(&c, world, id, rank);
    if (rc != ncclSuccess) {
        if (err) *err = nerr(n, "ncclCommInitRank", rc);
        return SFM_ERR_NCCL;
    }
    *comm = c;
    return SFM_OK;
}

// Asynchronous communicator errors (a peer died, a transport failed) never surface through the
// enqueue calls; every step's read-back polls this.  On error the
// communicator is aborted so that queued collectives fail instead of hanging.
int nccl_async_error(Nccl* n, void* comm, std::string* err) {
    if (!n || !comm || !n->CommGetAsyncError) return SFM_OK;
    int async = ncclSuccess;
    const int rc = n->CommGetAsyncError((ncclComm_t)comm, &async);
    if (rc == ncclSuccess && async == ncclSuccess) return SFM_OK;
    if (err) *err = nerr(n, "ncclCommGetAsyncError", rc != ncclSuccess ? rc : async);
    return SFM_ERR_NCCL;
}

int nccl_abort(Nccl* n, void* comm) {
    if (n && comm && n->CommAbort) n->CommAbort((ncclComm_t)comm);
    return SFM_OK;
}

int nccl_destroy(Nccl* n, void* comm) {
    if (n && comm) n->CommDestroy((ncclComm_t)comm);
    return SFM_OK;
}

int nccl_allreduce_f32(Nccl* n, void* comm, float* buf, size_t count, cudaStream_t st,
                       std::string* err) {
    const int rc = n->AllReduce(buf, buf, count, ncclFloat32, ncclSum, (ncclComm_t)comm, st);
    if (rc != ncclSuccess) {
        if (err) *err = nerr(n, "ncclAllReduce(f32)", rc);
        return SFM_ERR_NCCL;
    }
    return SFM_OK;
}

int nccl_allreduce_f64(Nccl* n, void* comm, double* buf, size_t count, cudaStream_t st,
                       std::string* err) {
    const int rc = n->AllReduce(buf, buf, count, ncclFloat64, ncclSum, (ncclComm_t)comm, st);
    if (rc != ncclSuccess) {
        if (err) *err = nerr(n, "ncclAllReduce(f64)", rc);
        return SFM_ERR_NCCL;
    }
    return SFM_OK;
}

int nccl_bcast_f32(Nccl* n, void* comm, float* buf, size_t count, int root, cudaStream_t st,
                   std::string* err) {
    const int rc = n->Broadcast(buf, buf, count, ncclFloat32, root, (ncclComm_t)comm, st);
    if (rc != ncclSuccess) {
        if (err) *err = nerr(n, "ncclBroadcast", rc);
        return SFM_ERR_NCCL;
    }
    return SFM_OK;
}

int nccl_allgather_i32(Nccl* n, void* comm, const int32_t* send, int32_t* recv, size_t count,
                       cudaStream_t st, std::string* err) {
    const int rc = n->AllGather(send, recv, count, ncclInt32, (ncclComm_t)comm, st);
    if (rc != ncclSuccess) {
        if (err) *err = nerr(n, "ncclAllGather(i32)", rc);
        return SFM_ERR_NCCL;
    }
    return SFM_OK;
}

int nccl_allgather_f32(Nccl* n, void* comm, const float* send, float* recv, size_t count,
                       cudaStream_t st, std::string* err) {
    const int rc = n->AllGather(send, recv, count, ncclFloat32, (ncclComm_t)comm, st);
    if (rc != ncclSuccess) {
        if (err) *err = nerr(n, "ncclAllGather(f32)", rc);
        return SFM_ERR_NCCL;
    }
    return SFM_OK;
}

// Variable all-to-all of 4-byte elements: rank r sends send_cnt[p] * width elements starting at
// send_off[p] * width to peer p and receives recv_cnt[p] * width at recv_off[p] * width.  The self
// part is a device copy; zero-sized parts are skipped on both sides (both know the counts).
int nccl_alltoallv_4b(Nccl* n, void* comm, int rank, int world, const void* send,
                      const int64_t* send_off, const int64_t* send_cnt, void* recv,
                      const int64_t* recv_off, const int64_t* recv_cnt, int64_t width,
                      cudaStream_t st, std::string* err) {
    const char* sb = (const char*)send;
    char* rb = (char*)recv;
    if (send_cnt[rank] > 0) {
        if (send_cnt[rank] != recv_cnt[rank]) {
            if (err) *err = "alltoallv: self counts differ";
            return SFM_ERR_NCCL;
        }
        if (cudaMemcpyAsync(rb + recv_off[rank] * width * 4, sb + send_off[rank] * width * 4,
                            (size_t)(send_cnt[rank] * width * 4), cudaMemcpyDeviceToDevice, st) !=
            cudaSuccess) {
            if (err) *err = "alltoallv: self copy failed";
            return SFM_ERR_CUDA;
        }
    }
    int rc = n->GroupStart();
    for (int p = 0; p < world && rc == ncclSuccess; ++p) {
        if (p == rank) continue;
        if (send_cnt[p] > 0)
            rc = n->Send(sb + send_off[p] * width * 4, (size_t)(send_cnt[p] * width), ncclFloat32, p,
                         (ncclComm_t)comm, st);
        if (rc == ncclSuccess && recv_cnt[p] > 0)
            rc = n->Recv(rb + recv_off[p] * width * 4, (size_t)(recv_cnt[p] * width), ncclFloat32, p,
                         (ncclComm_t)comm, st);
    }
    const int rc2 = n->GroupEnd();
    if (rc == ncclSuccess) rc = rc2;
    if (rc != ncclSuccess) {
        if (err) *err = nerr(n, "ncclSend/ncclRecv", rc);
        return SFM_ERR_NCCL;
    }
    return SFM_OK;
}

int nccl_group_start(Nccl* n) { return n->GroupStart(); }
int nccl_group_end(Nccl* n) { return n->GroupEnd(); }

}  // namespace sfm
