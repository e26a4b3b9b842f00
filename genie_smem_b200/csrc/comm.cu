// Multi-GPU side of the path: the ONE collective it has (north_star (4)) -- per-rank SMEM records to one rank.
//
// Two transports, both behind the C ABI:
//   * peer memory (gsm_peer_export / gsm_peer_open + gsm_smem_collect_gathered in kernels.cu): the destination buffer on
//     the gathering rank is mapped into every rank through a CUDA IPC handle, and each rank's ordered-write kernel
//     (k_gather_records, the last kernel of its step) stores its records straight into that buffer over NVLink / NVSwitch.
//     The gather is then not a pass of its own: every record is written once, to its final place.  Only the per-rank
//     record counts (8 bytes per rank) travel through NCCL (gsm_comm_allgather_u64), device to device, no host round trip.
//   * NCCL point-to-point (gsm_gather_records): one ncclGroup of ncclSend / ncclRecv with exact sizes into a
//     preallocated device buffer -- the baseline the fused path is measured against, and the path when IPC is not
//     available.
// NCCL is reached through dlopen("libnccl.so.2") (the copy torch has already loaded when the host is Python), so
// libgenie_smem.so has no link-time dependency on it and loads on machines without NCCL; only these entry points need it.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdint>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/genie_smem.h"
#include "host_common.hpp"

namespace gsm {

#define GSM_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return fail(GSM_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));       \
    } while (0)

// the slice of the NCCL API this file uses (nccl.h: stable ABI since 2.x)
struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
enum { NCCL_UINT8 = 1, NCCL_UINT64 = 5 };

struct NcclApi {
    void* so = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
};

static NcclApi g_nccl;
static std::once_flag g_nccl_once;
static std::string g_nccl_err;

static void load_nccl() {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        g_nccl.so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.so) break;
    }
    if (!g_nccl.so) { g_nccl_err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return; }
    auto sym = [](const char* n) { return dlsym(g_nccl.so, n); };
    g_nccl.GetUniqueId = (int (*)(NcclUniqueId*))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))sym("ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
    g_nccl.Send = (int (*)(const void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclSend");
    g_nccl.Recv = (int (*)(void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclRecv");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))sym("ncclAllGather");
    g_nccl.GroupStart = (int (*)())sym("ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())sym("ncclGroupEnd");
    g_nccl.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    g_nccl.GetVersion = (int (*)(int*))sym("ncclGetVersion");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.Send || !g_nccl.Recv || !g_nccl.AllGather ||
        !g_nccl.GroupStart || !g_nccl.GroupEnd || !g_nccl.GetErrorString) {
        g_nccl_err = "libnccl.so.2 lacks a required symbol";
        g_nccl.so = nullptr;
    }
}

static int need_nccl() {
    std::call_once(g_nccl_once, load_nccl);
    if (!g_nccl.so) return fail(GSM_E_INVALID, "NCCL unavailable: " + g_nccl_err);
    return GSM_OK;
}

#define GSM_NCCL(call)                                                                                   \
    do {                                                                                                 \
        int r__ = (call);                                                                                \
        if (r__ != 0) return fail(GSM_E_CUDA, std::string(#call) + ": " + g_nccl.GetErrorString(r__));   \
    } while (0)

// base address of the allocation that holds p (driver entry point resolved through the runtime: no link against libcuda)
static int alloc_base(const void* p, void** base) {
    typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    GSM_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return fail(GSM_E_CUDA, "cuMemGetAddressRange is not available from this driver");
    unsigned long long b = 0;
    size_t size = 0;
    const int r = ((range_fn)fn)(&b, &size, (unsigned long long)(uintptr_t)p);
    if (r != 0) return fail(GSM_E_CUDA, "cuMemGetAddressRange failed with CUresult " + std::to_string(r));
    *base = (void*)(uintptr_t)b;
    return GSM_OK;
}

}  // namespace gsm

using namespace gsm;

struct gsm_comm {
    NcclComm comm;
    int rank, world;
};

extern "C" {

int gsm_comm_unique_id(void* id128) {
    if (!id128) return fail(GSM_E_INVALID, "gsm_comm_unique_id: null");
    int st = need_nccl();
    if (st) return st;
    NcclUniqueId id;
    GSM_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return GSM_OK;
}

int gsm_comm_init(const void* id128, int rank, int world, gsm_comm** out) {
    if (!id128 || !out || world < 1 || rank < 0 || rank >= world) return fail(GSM_E_INVALID, "gsm_comm_init: bad rank / world / id");
    int st = need_nccl();
    if (st) return st;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(GSM_E_NODEVICE, "no CUDA device visible: libgenie_smem has no CPU fallback");
    }
    NcclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    gsm_comm* c = new gsm_comm{nullptr, rank, world};
    int r = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (r != 0) {
        delete c;
        return fail(GSM_E_CUDA, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
    }
    *out = c;
    return GSM_OK;
}

int gsm_comm_free(gsm_comm* c) {
    if (!c) return GSM_OK;
    if (g_nccl.so && c->comm) g_nccl.CommDestroy(c->comm);
    delete c;
    return GSM_OK;
}

int gsm_comm_info(const gsm_comm* c, int* rank, int* world, int* nccl_version) {
    if (!c) return fail(GSM_E_INVALID, "gsm_comm_info: null");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    if (nccl_version) { *nccl_version = 0; if (g_nccl.GetVersion) g_nccl.GetVersion(nccl_version); }
    return GSM_OK;
}

int gsm_comm_allgather_u64(gsm_comm* c, const uint64_t* send_dev, uint64_t n, uint64_t* recv_dev, void* stream) {
    if (!c || !send_dev || !recv_dev || n == 0) return fail(GSM_E_INVALID, "gsm_comm_allgather_u64: null");
    GSM_NCCL(g_nccl.AllGather(send_dev, recv_dev, (size_t)n, NCCL_UINT64, c->comm, (cudaStream_t)stream));
    return GSM_OK;
}

int gsm_gather_records(gsm_comm* c, const gsm_record* send, uint64_t n_send, const uint64_t* counts, gsm_record* recv, int dst, void* stream) {
    if (!c || !counts || dst < 0 || dst >= c->world) return fail(GSM_E_INVALID, "gsm_gather_records: null comm / counts or bad dst");
    if (counts[c->rank] != n_send) return fail(GSM_E_INVALID, "gsm_gather_records: counts[rank] != n_send");
    if (n_send && !send) return fail(GSM_E_INVALID, "gsm_gather_records: null send buffer");
    cudaStream_t s = (cudaStream_t)stream;
    if (c->rank == dst) {
        uint64_t total = 0;
        for (int r = 0; r < c->world; ++r) total += counts[r];
        if (total && !recv) return fail(GSM_E_INVALID, "gsm_gather_records: null recv buffer on dst");
        uint64_t off = 0;
        GSM_NCCL(g_nccl.GroupStart());
        for (int r = 0; r < c->world; ++r) {
            if (r != dst && counts[r]) {
                int e = g_nccl.Recv(recv + off, (size_t)counts[r] * sizeof(gsm_record), NCCL_UINT8, r, c->comm, s);
                if (e != 0) { g_nccl.GroupEnd(); return fail(GSM_E_CUDA, std::string("ncclRecv: ") + g_nccl.GetErrorString(e)); }
            }
            off += counts[r];
        }
        GSM_NCCL(g_nccl.GroupEnd());
        off = 0;
        for (int r = 0; r < dst; ++r) off += counts[r];
        if (n_send && recv + off != send)       // own shard: device-to-device on the same stream
            GSM_CUDA(cudaMemcpyAsync(recv + off, send, (size_t)n_send * sizeof(gsm_record), cudaMemcpyDeviceToDevice, s));
    } else if (n_send) {
        GSM_NCCL(g_nccl.GroupStart());
        int e = g_nccl.Send(send, (size_t)n_send * sizeof(gsm_record), NCCL_UINT8, dst, c->comm, s);
        if (e != 0) { g_nccl.GroupEnd(); return fail(GSM_E_CUDA, std::string("ncclSend: ") + g_nccl.GetErrorString(e)); }
        GSM_NCCL(g_nccl.GroupEnd());
    }
    return GSM_OK;
}

// ------------------------------------------------------------------------------------------------ peer memory (CUDA IPC)
int gsm_peer_export(const void* dev_ptr, void* handle64, uint64_t* offset) {
    if (!dev_ptr || !handle64 || !offset) return fail(GSM_E_INVALID, "gsm_peer_export: null");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaIpcMemHandle_t h;
    GSM_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
    // the handle names the whole allocation: report where dev_ptr sits inside it
    void* base = nullptr;
    const int st = alloc_base(dev_ptr, &base);
    if (st) return st;
    memcpy(handle64, &h, 64);
    *offset = (uint64_t)((const char*)dev_ptr - (const char*)base);
    return GSM_OK;
}

int gsm_peer_open(const void* handle64, uint64_t offset, void** dev_ptr) {
    if (!handle64 || !dev_ptr) return fail(GSM_E_INVALID, "gsm_peer_open: null");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* base = nullptr;
    GSM_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr = (char*)base + offset;
    return GSM_OK;
}

int gsm_peer_close(void* dev_ptr, uint64_t offset) {
    if (!dev_ptr) return GSM_OK;
    GSM_CUDA(cudaIpcCloseMemHandle((char*)dev_ptr - offset));
    return GSM_OK;
}

}  // extern "C"
