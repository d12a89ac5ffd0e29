// tdm_comm.cpp -- the multi-GPU epilogue of the path in the C layer: gather the decoded symbol streams (packed four
// dibits per byte, TDM_OUT_PACKED) and their counts from every rank to one over NCCL / NVLink.
//
// Channels are independent recurrences, sharded by contiguous channel block, one process (or thread) per GPU; nothing
// is exchanged while demodulating.  This gather is the ONLY collective the path has
// (BASELINE.json configs[3]: "4096 channels x 4e6 samples sharded across 8xB200, NCCL gather of decoded symbols").
// It is a plain data movement with no arithmetic to fuse into: the slicer role of the demodulation kernel already
// writes the packed form, so what leaves a GPU is 0.125 B per input sample (DESIGN.md "Multi-GPU").
//
// NCCL is NOT a link-time dependency of libtdm_b200.so: libnccl.so.2 is opened at run time the first time a
// communicator is made (a host without NCCL can still use every single-GPU entry point), and only its public C API
// is used (nccl.h of the 2.x series: ncclGetUniqueId, ncclCommInitRank, ncclSend/ncclRecv in a group).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstring>
#include <mutex>
#include <new>
#include "tdm_b200.h"
#include "tdm_internal.h"

namespace {

// the slice of nccl.h this file needs (ABI-stable across NCCL 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess_ = 0 };
enum { ncclInt8_ = 0, ncclUint8_ = 1, ncclInt32_ = 2 };

struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

Nccl& nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        // an already loaded libnccl.so.2 (e.g. the one PyTorch bundles) is reused: same soname
        n.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!n.lib) { n.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL); }
        if (!n.lib) { return; }
#define TDM_SYM(field, name) n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.lib, name))
        TDM_SYM(GetUniqueId, "ncclGetUniqueId");
        TDM_SYM(CommInitRank, "ncclCommInitRank");
        TDM_SYM(CommDestroy, "ncclCommDestroy");
        TDM_SYM(GroupStart, "ncclGroupStart");
        TDM_SYM(GroupEnd, "ncclGroupEnd");
        TDM_SYM(Send, "ncclSend");
        TDM_SYM(Recv, "ncclRecv");
        TDM_SYM(GetErrorString, "ncclGetErrorString");
#undef TDM_SYM
        n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.GroupStart && n.GroupEnd && n.Send && n.Recv && n.GetErrorString;
    });
    return n;
}

int need_nccl() {
    if (!nccl().ok) { return tdm_internal_fail(TDM_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded (the multi-GPU gather needs NCCL 2.x at run time)"); }
    return TDM_OK;
}

#define TDM_NCCL(expr)                                                                                                  \
    do {                                                                                                                \
        ncclResult_t r__ = (expr);                                                                                      \
        if (r__ != ncclSuccess_) { return tdm_internal_fail(TDM_ERR_CUDA, "%s: %s", #expr, nccl().GetErrorString(r__)); } \
    } while (0)

}  // namespace

struct tdm_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    bool owned = true;
};

extern "C" {

int tdm_comm_unique_id(uint8_t id[TDM_COMM_ID_BYTES]) {
    if (!id) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_comm_unique_id: null id"); }
    int rc = need_nccl();
    if (rc != TDM_OK) { return rc; }
    ncclUniqueId u;
    TDM_NCCL(nccl().GetUniqueId(&u));
    static_assert(sizeof(u) == TDM_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    std::memcpy(id, &u, sizeof(u));
    return TDM_OK;
}

int tdm_comm_create(const uint8_t id[TDM_COMM_ID_BYTES], int32_t rank, int32_t world, int32_t device, tdm_comm** out) {
    if (!out) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_comm_create: out is null"); }
    *out = nullptr;
    if (!id || world < 1 || rank < 0 || rank >= world) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_comm_create: bad arguments"); }
    int rc = need_nccl();
    if (rc != TDM_OK) { return rc; }
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) { return tdm_internal_fail(TDM_ERR_NO_DEVICE, "tdm_comm_create: cannot select device %d", device); }
    tdm_comm* c = new (std::nothrow) tdm_comm();
    if (!c) { return tdm_internal_fail(TDM_ERR_NOMEM, "tdm_comm_create: out of host memory"); }
    c->rank = rank; c->world = world; c->device = device;
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof(u));
    const ncclResult_t r = nccl().CommInitRank(&c->comm, world, u, rank);
    if (prev >= 0 && prev != device) { cudaSetDevice(prev); }
    if (r != ncclSuccess_) { delete c; return tdm_internal_fail(TDM_ERR_CUDA, "ncclCommInitRank: %s", nccl().GetErrorString(r)); }
    *out = c;
    return TDM_OK;
}

int tdm_comm_adopt(void* nccl_comm, int32_t rank, int32_t world, int32_t device, tdm_comm** out) {
    if (!out || !nccl_comm || world < 1 || rank < 0 || rank >= world) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_comm_adopt: bad arguments"); }
    int rc = need_nccl();
    if (rc != TDM_OK) { return rc; }
    tdm_comm* c = new (std::nothrow) tdm_comm();
    if (!c) { return tdm_internal_fail(TDM_ERR_NOMEM, "tdm_comm_adopt: out of host memory"); }
    c->comm = (ncclComm_t)nccl_comm; c->rank = rank; c->world = world; c->device = device; c->owned = false;
    *out = c;
    return TDM_OK;
}

int tdm_comm_destroy(tdm_comm* c) {
    if (!c) { return TDM_OK; }
    if (c->owned && c->comm && nccl().ok) { nccl().CommDestroy(c->comm); }
    delete c;
    return TDM_OK;
}

int tdm_gather_packed(tdm_comm* c, int32_t dst, int32_t n_rows, const uint8_t* packed, int64_t packed_stride, const int32_t* counts,
                      uint8_t* packed_all, int32_t* counts_all, void* cuda_stream) {
    if (!c || !packed || !counts || n_rows <= 0 || packed_stride <= 0 || dst < 0 || dst >= c->world) {
        return tdm_internal_fail(TDM_ERR_ARG, "tdm_gather_packed: bad arguments");
    }
    if (c->rank == dst && (!packed_all || !counts_all)) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_gather_packed: the destination rank needs packed_all / counts_all"); }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t row_bytes = (size_t)n_rows * (size_t)packed_stride;
    if (c->world == 1) {
        if (packed_all != packed && cudaMemcpyAsync(packed_all, packed, row_bytes, cudaMemcpyDeviceToDevice, st) != cudaSuccess) { return tdm_internal_fail(TDM_ERR_CUDA, "tdm_gather_packed: copy failed"); }
        if (counts_all != counts && cudaMemcpyAsync(counts_all, counts, sizeof(int32_t) * (size_t)n_rows, cudaMemcpyDeviceToDevice, st) != cudaSuccess) { return tdm_internal_fail(TDM_ERR_CUDA, "tdm_gather_packed: copy failed"); }
        return TDM_OK;
    }
    // one group: every rank's rows land rank-major at the destination, which is channel-major because the shards
    // are contiguous channel blocks (equal-sized: pad the last shard)
    TDM_NCCL(nccl().GroupStart());
    if (c->rank == dst) {
        for (int r = 0; r < c->world; ++r) {
            if (r == dst) { continue; }
            TDM_NCCL(nccl().Recv(packed_all + (size_t)r * row_bytes, row_bytes, ncclUint8_, r, c->comm, st));
            TDM_NCCL(nccl().Recv(counts_all + (size_t)r * (size_t)n_rows, (size_t)n_rows, ncclInt32_, r, c->comm, st));
        }
    } else {
        TDM_NCCL(nccl().Send(packed, row_bytes, ncclUint8_, dst, c->comm, st));
        TDM_NCCL(nccl().Send(counts, (size_t)n_rows, ncclInt32_, dst, c->comm, st));
    }
    TDM_NCCL(nccl().GroupEnd());
    if (c->rank == dst) {
        if (cudaMemcpyAsync(packed_all + (size_t)dst * row_bytes, packed, row_bytes, cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(counts_all + (size_t)dst * (size_t)n_rows, counts, sizeof(int32_t) * (size_t)n_rows, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
            return tdm_internal_fail(TDM_ERR_CUDA, "tdm_gather_packed: local copy failed");
        }
    }
    return TDM_OK;
}

}  // extern "C"
