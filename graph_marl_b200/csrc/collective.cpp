// collective.cpp -- the one collective of the path: the optional gradient all-reduce of the learner.
//
// Reference hook: src/main.py:1002-1004 (between `loss.backward()` and `optimizer.step()`); SURVEY 8(b) lists it as
// `gm_allreduce_grads(comm, flat f32*, n, stream)`, 8(e) as "ncclAllReduce(sum, fp32, ~0.94 M elements) once per
// training iteration".  The rollout path itself has no collective (env instances are independent).
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 the process already has -- torch's bundled copy -- or the
// system one): the library keeps no link-time dependency on it and single-GPU users never load it.  The communicator
// is this library's own (ncclCommInitRank from a 128-byte unique id that rank 0 creates and the host code
// broadcasts), so the call works from any host language, not only under torch.distributed.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace gm {

struct NcclApi {
    decltype(&ncclGetUniqueId) get_unique_id = nullptr;
    decltype(&ncclCommInitRank) comm_init_rank = nullptr;
    decltype(&ncclCommDestroy) comm_destroy = nullptr;
    decltype(&ncclAllReduce) all_reduce = nullptr;
    decltype(&ncclBroadcast) broadcast = nullptr;
    decltype(&ncclGetErrorString) error_string = nullptr;
    decltype(&ncclGetVersion) get_version = nullptr;
    bool ok = false;
};

static NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy already mapped into the process (torch's)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return api;
#define GM_SYM(field, name) api.field = (decltype(api.field))dlsym(h, name)
    GM_SYM(get_unique_id, "ncclGetUniqueId");
    GM_SYM(comm_init_rank, "ncclCommInitRank");
    GM_SYM(comm_destroy, "ncclCommDestroy");
    GM_SYM(all_reduce, "ncclAllReduce");
    GM_SYM(broadcast, "ncclBroadcast");
    GM_SYM(error_string, "ncclGetErrorString");
    GM_SYM(get_version, "ncclGetVersion");
#undef GM_SYM
    api.ok = api.get_unique_id && api.comm_init_rank && api.comm_destroy && api.all_reduce && api.broadcast && api.error_string;
    return api;
}

#define GM_NCCL(call)                                                                   \
    do {                                                                                \
        ncclResult_t r__ = (call);                                                      \
        if (r__ != ncclSuccess) {                                                       \
            gm::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, nccl().error_string(r__)); \
            return GM_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

}  // namespace gm

using namespace gm;

extern "C" {

int gm_nccl_version(void) {
    int v = 0;
    if (!nccl().ok || !nccl().get_version || nccl().get_version(&v) != ncclSuccess) return 0;
    return v;
}

int gm_nccl_unique_id(uint8_t* id128) {
    GM_CHECK_ARG(id128, "null id buffer");
    GM_CHECK_ARG(nccl().ok, "libnccl.so.2 could not be loaded");
    static_assert(sizeof(ncclUniqueId) == GM_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    GM_NCCL(nccl().get_unique_id(&id));
    memcpy(id128, &id, sizeof(id));
    return GM_OK;
}

int gm_nccl_comm_create(int32_t world_size, int32_t rank, const uint8_t* id128, void** comm) {
    GM_CHECK_ARG(id128 && comm && world_size >= 1 && rank >= 0 && rank < world_size, "bad communicator arguments");
    GM_CHECK_ARG(nccl().ok, "libnccl.so.2 could not be loaded");
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    GM_NCCL(nccl().comm_init_rank(&c, world_size, id, rank));
    *comm = (void*)c;
    return GM_OK;
}

int gm_nccl_comm_destroy(void* comm) {
    if (!comm) return GM_OK;
    GM_CHECK_ARG(nccl().ok, "libnccl.so.2 could not be loaded");
    GM_NCCL(nccl().comm_destroy((ncclComm_t)comm));
    return GM_OK;
}

int gm_allreduce_grads(void* comm, float* flat, int64_t n, int32_t average, void* stream) {
    GM_CHECK_ARG(comm && flat && n >= 0, "bad all-reduce arguments");
    GM_CHECK_ARG(nccl().ok, "libnccl.so.2 could not be loaded");
    if (n == 0) return GM_OK;
    GM_NCCL(nccl().all_reduce(flat, flat, (size_t)n, ncclFloat32, average ? ncclAvg : ncclSum, (ncclComm_t)comm, (cudaStream_t)stream));
    return GM_OK;
}

int gm_broadcast_weights(void* comm, float* flat, int64_t n, int32_t root, void* stream) {
    GM_CHECK_ARG(comm && flat && n >= 0, "bad broadcast arguments");
    GM_CHECK_ARG(nccl().ok, "libnccl.so.2 could not be loaded");
    if (n == 0) return GM_OK;
    GM_NCCL(nccl().broadcast(flat, flat, (size_t)n, ncclFloat32, root, (ncclComm_t)comm, (cudaStream_t)stream));
    return GM_OK;
}

}  // extern "C"
