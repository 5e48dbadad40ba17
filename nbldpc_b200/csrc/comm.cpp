// Multi-GPU merge of the Monte-Carlo counters: extern "C" nbl_comm_* / nbl_allreduce_counters of include/nbldpc_b200.h.
//
// The reference runs `parallel` CPU workers and adds their six counters serially on the host (main.cpp:46-51,
// Comm.cpp:484-493).  Here frames are sharded over GPUs by global frame index (nbl_sim_config.first_frame) and the
// 8 x u64 counter vector of every GPU is summed IN PLACE by one ncclAllReduce on each decoder's stream -- the only
// collective of the path.  NCCL is bound at run time (dlopen of libnccl.so.2: the copy the process already holds -- e.g.
// torch's bundled one -- else the system's), so single-GPU users of the library need no NCCL at all.
#include "internal.h"
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

NcclApi *nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    // A copy the process already holds wins (a Python process that imported torch has torch's bundled NCCL under the same
    // soname: two different NCCL versions must never meet in one process); otherwise the system's.  RTLD_LOCAL: our binding
    // goes through dlsym, nothing else should resolve against it.
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
        if (h) break;
        h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    }
    if (!h) return nullptr;
#define BIND(field, sym) api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, sym)); if (!api.field) return nullptr;
    BIND(GetUniqueId, "ncclGetUniqueId") BIND(CommInitRank, "ncclCommInitRank") BIND(CommInitAll, "ncclCommInitAll")
    BIND(CommDestroy, "ncclCommDestroy") BIND(AllReduce, "ncclAllReduce") BIND(GroupStart, "ncclGroupStart")
    BIND(GroupEnd, "ncclGroupEnd") BIND(GetErrorString, "ncclGetErrorString") BIND(GetVersion, "ncclGetVersion")
#undef BIND
    api.handle = h;
    return &api;
}

int nccl_fail(const char *what, ncclResult_t r) {
    nbl_set_error(std::string(what) + " failed: " + (nccl() ? nccl()->GetErrorString(r) : "NCCL not loaded"));
    return NBL_ERR_CUDA;
}

}  // namespace

struct nbl_comm {
    int world = 1;                      // ranks of the communicator
    std::vector<int> ranks;             // the ranks this process holds (all of them after nbl_comm_create, one after _create_rank)
    std::vector<int> devices;           // CUDA device of each local rank
    std::vector<ncclComm_t> comms;      // one NCCL communicator per local rank
};

#define NEED_NCCL(api)                                                                                         \
    NcclApi *api = nccl();                                                                                     \
    if (!api) { nbl_set_error("NCCL (libnccl.so.2) could not be loaded; multi-GPU counter merge unavailable"); return NBL_ERR_CUDA; }

extern "C" int nbl_comm_create(nbl_comm **out, int n_devices, const int *devices) {
    if (!out || n_devices < 1) { nbl_set_error("nbl_comm_create: bad argument"); return NBL_ERR_ARG; }
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess) { cudaGetLastError(); have = 0; }
    nbl_comm *c = new nbl_comm;
    c->world = n_devices;
    for (int i = 0; i < n_devices; ++i) {
        const int dv = devices ? devices[i] : i;
        if (dv < 0 || dv >= have) { delete c; nbl_set_error("nbl_comm_create: no such CUDA device (this library has no CPU fallback)"); return NBL_ERR_CUDA; }
        c->ranks.push_back(i); c->devices.push_back(dv);
    }
    if (n_devices > 1) {
        NcclApi *api = nccl();
        if (!api) { delete c; nbl_set_error("NCCL (libnccl.so.2) could not be loaded; multi-GPU counter merge unavailable"); return NBL_ERR_CUDA; }
        c->comms.assign(n_devices, nullptr);
        const ncclResult_t r = api->CommInitAll(c->comms.data(), n_devices, c->devices.data());
        if (r != ncclSuccess) { delete c; return nccl_fail("ncclCommInitAll", r); }
    }
    *out = c;
    return NBL_OK;
}

extern "C" int nbl_comm_unique_id(void *id128) {
    if (!id128) { nbl_set_error("nbl_comm_unique_id: null argument"); return NBL_ERR_ARG; }
    NEED_NCCL(api)
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    const ncclResult_t r = api->GetUniqueId(reinterpret_cast<ncclUniqueId *>(id128));
    return r == ncclSuccess ? NBL_OK : nccl_fail("ncclGetUniqueId", r);
}

extern "C" int nbl_comm_create_rank(nbl_comm **out, int world, int rank, const void *id128, int device) {
    if (!out || world < 1 || rank < 0 || rank >= world || (world > 1 && !id128)) { nbl_set_error("nbl_comm_create_rank: bad argument"); return NBL_ERR_ARG; }
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess) { cudaGetLastError(); have = 0; }
    if (device < 0 || device >= have) { nbl_set_error("nbl_comm_create_rank: no such CUDA device (this library has no CPU fallback)"); return NBL_ERR_CUDA; }
    nbl_comm *c = new nbl_comm;
    c->world = world; c->ranks.push_back(rank); c->devices.push_back(device);
    if (world > 1) {
        NcclApi *api = nccl();
        if (!api) { delete c; nbl_set_error("NCCL (libnccl.so.2) could not be loaded; multi-GPU counter merge unavailable"); return NBL_ERR_CUDA; }
        if (cudaSetDevice(device) != cudaSuccess) { delete c; nbl_set_error("cudaSetDevice failed"); return NBL_ERR_CUDA; }
        ncclUniqueId id;
        memcpy(&id, id128, sizeof id);
        c->comms.assign(1, nullptr);
        const ncclResult_t r = api->CommInitRank(&c->comms[0], world, id, rank);
        if (r != ncclSuccess) { delete c; return nccl_fail("ncclCommInitRank", r); }
    }
    *out = c;
    return NBL_OK;
}

extern "C" void nbl_comm_destroy(nbl_comm *c) {
    if (!c) return;
    NcclApi *api = nccl();
    for (size_t i = 0; i < c->comms.size(); ++i)
        if (c->comms[i] && api) { cudaSetDevice(c->devices[i]); api->CommDestroy(c->comms[i]); }
    delete c;
}

extern "C" int nbl_comm_size(const nbl_comm *c) { return c ? c->world : 0; }
extern "C" int nbl_comm_local_ranks(const nbl_comm *c) { return c ? (int)c->ranks.size() : 0; }
extern "C" int nbl_comm_rank(const nbl_comm *c, int local_index) {
    return c && local_index >= 0 && local_index < (int)c->ranks.size() ? c->ranks[local_index] : -1;
}
extern "C" int nbl_comm_device(const nbl_comm *c, int local_index) {
    return c && local_index >= 0 && local_index < (int)c->devices.size() ? c->devices[local_index] : -1;
}

extern "C" int nbl_allreduce_counters(nbl_comm *c, nbl_decoder *const *decoders, uint64_t *const *dev_counters) {
    if (!c || !decoders || !dev_counters) { nbl_set_error("nbl_allreduce_counters: null argument"); return NBL_ERR_ARG; }
    const int n = (int)c->ranks.size();
    for (int i = 0; i < n; ++i)
        if (!decoders[i] || !dev_counters[i]) { nbl_set_error("nbl_allreduce_counters: null decoder or counter pointer"); return NBL_ERR_ARG; }
    if (c->world == 1) return NBL_OK;                               // one rank: the sum over ranks is the vector itself
    NEED_NCCL(api)
    ncclResult_t r = api->GroupStart();
    if (r != ncclSuccess) return nccl_fail("ncclGroupStart", r);
    for (int i = 0; i < n; ++i) {
        if (cudaSetDevice(c->devices[i]) != cudaSuccess) { api->GroupEnd(); nbl_set_error("cudaSetDevice failed"); return NBL_ERR_CUDA; }
        r = api->AllReduce(dev_counters[i], dev_counters[i], NBL_NUM_COUNTERS, ncclUint64, ncclSum, c->comms[i],
                           (cudaStream_t)nbl_decoder_stream(decoders[i]));
        if (r != ncclSuccess) { api->GroupEnd(); return nccl_fail("ncclAllReduce", r); }
    }
    r = api->GroupEnd();
    return r == ncclSuccess ? NBL_OK : nccl_fail("ncclGroupEnd", r);
}
