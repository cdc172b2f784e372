#include "nccl_shim.cuh"

#include <dlfcn.h>

#include <cstring>
#include <stdexcept>
#include <string>

namespace sage {

struct NcclError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

namespace {
// ABI subset of nccl.h (stable across NCCL 2.x): ncclUniqueId is 128 opaque bytes; ncclFloat64 = 8; ncclSum = 0.
typedef struct {
    char internal[128];
} ncclUniqueId_t;
typedef void *ncclComm_t;
typedef int (*GetUniqueId_t)(ncclUniqueId_t *);
typedef int (*CommInitRank_t)(ncclComm_t *, int, ncclUniqueId_t, int);
typedef int (*CommDestroy_t)(ncclComm_t);
typedef int (*AllReduce_t)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
typedef const char *(*GetErrorString_t)(int);

struct Api {
    void *h = nullptr;
    GetUniqueId_t GetUniqueId = nullptr;
    CommInitRank_t CommInitRank = nullptr;
    CommDestroy_t CommDestroy = nullptr;
    AllReduce_t AllReduce = nullptr;
    GetErrorString_t GetErrorString = nullptr;
};

Api &api() {
    static Api a;
    if (a.h) return a;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        a.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.h) break;
    }
    if (!a.h) throw NcclError(std::string("dlopen(libnccl.so.2) failed: ") + dlerror());
    a.GetUniqueId = (GetUniqueId_t)dlsym(a.h, "ncclGetUniqueId");
    a.CommInitRank = (CommInitRank_t)dlsym(a.h, "ncclCommInitRank");
    a.CommDestroy = (CommDestroy_t)dlsym(a.h, "ncclCommDestroy");
    a.AllReduce = (AllReduce_t)dlsym(a.h, "ncclAllReduce");
    a.GetErrorString = (GetErrorString_t)dlsym(a.h, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce) {
        // leave no half-initialised table behind: the next call must run the symbol check again, not call through a null pointer
        dlclose(a.h);
        a = Api{};
        throw NcclError("libnccl lacks required symbols");
    }
    return a;
}
void check(int rc, const char *what) {
    if (rc != 0) {
        const char *msg = api().GetErrorString ? api().GetErrorString(rc) : "?";
        throw NcclError(std::string(what) + ": " + msg);
    }
}
}  // namespace

struct NcclComm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

void nccl_unique_id(uint8_t out[128]) {
    ncclUniqueId_t id;
    check(api().GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(out, id.internal, 128);
}

NcclComm *nccl_comm_create(int rank, int world, const uint8_t idb[128]) {
    ncclUniqueId_t id;
    std::memcpy(id.internal, idb, 128);
    auto *c = new NcclComm;
    c->rank = rank, c->world = world;
    check(api().CommInitRank(&c->comm, world, id, rank), "ncclCommInitRank");
    return c;
}

void nccl_comm_destroy(NcclComm *c) {
    if (!c) return;
    if (c->comm) api().CommDestroy(c->comm);
    delete c;
}

void nccl_allreduce_sum_f64(NcclComm *c, double *dev_buf, int count, cudaStream_t stream) {
    check(api().AllReduce(dev_buf, dev_buf, (size_t)count, /*ncclFloat64*/ 8, /*ncclSum*/ 0, c->comm, stream), "ncclAllReduce");
}

}  // namespace sage
