// comm.cu -- the exchange step of a multi-GPU search: NCCL all-gather over NVLink / NVSwitch.
//
// The reduction this replaces is join_all + concat + sort over the per-shard candidate lists in
// src/ivf_index.rs:249-266; here every rank holds part of the index, scans it, and the per-rank top-k
// runs are all-gathered and merged on the device (merge_runs_kernel).
//
// NCCL is bound at run time (dlopen "libnccl.so.2") when the first communicator is created: the library
// then loads on hosts without NCCL, and a host process that already carries its own NCCL (e.g. PyTorch's
// bundled copy) shares that one instead of getting a second, different version mapped next to it.
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>
#include <string>

#include "index.h"

namespace vidx {

namespace {
struct NcclApi {
    void* h = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string error;
};
NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.h) break;
        }
        if (!api.h) {
            api.error = std::string("NCCL is not available: ") + dlerror();
            return;
        }
        auto sym = [&](const char* n) {
            void* p = dlsym(api.h, n);
            if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + n;
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    });
    if (!api.error.empty()) throw ApiError(VIDX_ERR_OTHER, api.error);
    return api;
}
void nccl_check(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) throw ApiError(VIDX_ERR_OTHER, std::string(what) + ": " + nccl().GetErrorString(r));
}
}  // namespace

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    std::string version;
};

static_assert(sizeof(ncclUniqueId) == 128, "VIDX_COMM_ID_BYTES");

void comm_unique_id(uint8_t out[128]) {
    ncclUniqueId id;
    nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(out, &id, 128);
}
Comm* comm_create(int device, int rank, int world, const uint8_t idb[128]) {
    NcclApi& api = nccl();
    ncclUniqueId id;
    memcpy(&id, idb, 128);
    Comm* c = new Comm();
    c->rank = rank;
    c->world = world;
    c->device = device;
    VIDX_CUDA(cudaSetDevice(device));
    ncclResult_t r = api.CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        delete c;
        nccl_check(r, "ncclCommInitRank");
    }
    int v = 0;
    api.GetVersion(&v);
    c->version = "NCCL " + std::to_string(v / 10000) + "." + std::to_string(v / 100 % 100) + "." + std::to_string(v % 100);
    return c;
}
void comm_destroy(Comm* c) {
    if (!c) return;
    if (c->comm) nccl().CommDestroy(c->comm);
    delete c;
}
int comm_rank(const Comm* c) { return c->rank; }
int comm_world(const Comm* c) { return c->world; }
const char* comm_version(const Comm* c) { return c->version.c_str(); }
void comm_all_gather(Comm* c, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t st) {
    nccl_check(nccl().AllGather(send, recv, bytes_per_rank, ncclChar, c->comm, st), "ncclAllGather");
}

}  // namespace vidx
