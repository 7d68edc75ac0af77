// The one collective of the path behind the C ABI: all-reduce of the PDT histograms and the beam-statistics sums over the
// ranks (reference: simulations/pdt.py:30-31, simulations/beam.py:35-71 reduce one process's samples).
// NCCL is bound at run time with dlopen: the library has no link-time dependency on it, a host that never creates a
// pa_comm never needs it, and a host that already carries NCCL (torch, or the system libnccl.so.2) shares that copy.
#include "../../include/pyatm_b200.h"

#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[PA_COMM_ID_BYTES]; } ncclUniqueId;      // NCCL_UNIQUE_ID_BYTES = 128
typedef int ncclResult_t;                                               // ncclSuccess = 0
constexpr int kNcclUint64 = 5, kNcclFloat64 = 8, kNcclSum = 0;          // nccl.h: ncclDataType_t / ncclRedOp_t

struct Nccl {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

int load_nccl(const Nccl** out) {
    static Nccl api;
    static int state = 0;       // 0 = not tried, 1 = ok, -1 = failed
    if (state == 0) {
        const char* names[] = {getenv("PYATM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* name : names) {
            if (!name || !*name) continue;
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        state = -1;
        if (api.handle) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
            api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
            api.GroupStart = (decltype(api.GroupStart))dlsym(api.handle, "ncclGroupStart");
            api.GroupEnd = (decltype(api.GroupEnd))dlsym(api.handle, "ncclGroupEnd");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
            if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GroupStart && api.GroupEnd && api.GetErrorString)
                state = 1;
        }
    }
    if (state != 1) {
        pa::set_error("NCCL is not available: dlopen(libnccl.so.2) failed (%s); set PYATM_NCCL_LIB to its path", dlerror() ? dlerror() : "missing symbols");
        return PA_ERR_STATE;
    }
    *out = &api;
    return PA_OK;
}

#define PA_NCCL(api, call)                                                                                   \
    do {                                                                                                     \
        ncclResult_t r__ = (call);                                                                           \
        if (r__ != 0) {                                                                                      \
            pa::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, (api)->GetErrorString(r__));          \
            return PA_ERR_CUDA;                                                                              \
        }                                                                                                    \
    } while (0)

}  // namespace

struct pa_comm {
    ncclComm_t comm = nullptr;
    int device = 0, rank = 0, world = 1;
};

extern "C" {

int pa_comm_unique_id(unsigned char* id_out) {
    PA_REQUIRE(id_out != nullptr, "bad arguments to pa_comm_unique_id: null id buffer");
    const Nccl* api = nullptr;
    if (int rc = load_nccl(&api)) return rc;
    ncclUniqueId id;
    PA_NCCL(api, api->GetUniqueId(&id));
    memcpy(id_out, id.internal, PA_COMM_ID_BYTES);
    return PA_OK;
}

int pa_comm_create(pa_comm** out, int device, int rank, int world, const unsigned char* id_bytes) {
    PA_REQUIRE(out && id_bytes && world >= 1 && rank >= 0 && rank < world, "bad arguments to pa_comm_create");
    const Nccl* api = nullptr;
    if (int rc = load_nccl(&api)) return rc;
    PA_CUDA(cudaSetDevice(device));
    ncclUniqueId id;
    memcpy(id.internal, id_bytes, PA_COMM_ID_BYTES);
    pa_comm* c = new pa_comm();
    c->device = device;
    c->rank = rank;
    c->world = world;
    const ncclResult_t r = api->CommInitRank(&c->comm, world, id, rank);
    if (r != 0) {
        pa::set_error("ncclCommInitRank(rank %d of %d) -> %s", rank, world, api->GetErrorString(r));
        delete c;
        return PA_ERR_CUDA;
    }
    *out = c;
    return PA_OK;
}

int pa_comm_destroy(pa_comm* c) {
    PA_REQUIRE(c != nullptr, "bad arguments to pa_comm_destroy: null communicator");
    const Nccl* api = nullptr;
    if (load_nccl(&api) == PA_OK && c->comm) api->CommDestroy(c->comm);
    delete c;
    return PA_OK;
}

int pa_stats_allreduce(pa_comm* c, unsigned long long* hist_dev, size_t nbins, double* sums_dev, size_t nsums, void* stream) {
    PA_REQUIRE(c && c->comm, "bad arguments to pa_stats_allreduce: null communicator");
    PA_REQUIRE((nbins == 0 || hist_dev) && (nsums == 0 || sums_dev), "null buffer with a non-zero count");
    const Nccl* api = nullptr;
    if (int rc = load_nccl(&api)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    PA_NCCL(api, api->GroupStart());       // both reductions in one launch
    if (nbins) PA_NCCL(api, api->AllReduce(hist_dev, hist_dev, nbins, kNcclUint64, kNcclSum, c->comm, st));
    if (nsums) PA_NCCL(api, api->AllReduce(sums_dev, sums_dev, nsums, kNcclFloat64, kNcclSum, c->comm, st));
    PA_NCCL(api, api->GroupEnd());
    return PA_OK;
}

}  // extern "C"
