// NCCL inside the library (SURVEY.md 8(b) "Threading", 8(e)): the statistics all-reduce of an EM iteration is issued
// by the library itself on its own stream -- no host-language callback in the hot loop -- in two set-ups:
//   * one PROCESS drives several GPUs (simples_init(ndev > 1, devs)): ncclCommInitAll, one host thread per device,
//     cells sharded inside simples_em_impute / simples_do_impute.  This is what an R session (.Call) uses.
//   * one process per GPU (torchrun): simples_comm_unique_id / simples_comm_init_rank build a rank communicator from
//     an id the host layer broadcasts; entry points called with n_total > n and no callback use it.
// libnccl is resolved at run time (dlopen), so the library has no link-time dependency on it and shares the copy
// PyTorch has already loaded when both live in one process.
#include <dlfcn.h>

#include <mutex>

#include "engine_internal.h"

namespace sbe {

namespace {

struct NcclUniqueId {
    char internal[128];
};
constexpr int kNcclFloat64 = 8, kNcclSum = 0;   // ncclDataType_t / ncclRedOp_t values (stable since NCCL 2.0)

struct NcclApi {
    void* lib = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

template <typename F>
bool sym(void* lib, const char* name, F& out) {
    out = reinterpret_cast<F>(dlsym(lib, name));
    return out != nullptr;
}

}  // namespace

int g_ndev = 0;
int g_devs[SIMPLES_MAX_DEVICES] = {0};
void* g_comms[SIMPLES_MAX_DEVICES] = {nullptr};   // single-process communicators (ncclCommInitAll)
void* g_rank_comm = nullptr;                      // this process's communicator in a one-process-per-GPU job
int g_rank_nranks = 0;

int nccl_load() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.lib) return SIMPLES_OK;
    const char* cands[4] = {getenv("SIMPLES_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
    void* lib = nullptr;
    std::string tried;
    for (const char* c : cands) {
        if (!c || !*c) continue;
        lib = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
        tried += std::string(" ") + c + " (" + dlerror() + ")";
    }
    if (!lib) return fail(SIMPLES_ECOMM, "cannot load NCCL (set SIMPLES_NCCL_LIB):" + tried);
    NcclApi a;
    a.lib = lib;
    if (!sym(lib, "ncclGetErrorString", a.GetErrorString) || !sym(lib, "ncclCommInitAll", a.CommInitAll) ||
        !sym(lib, "ncclCommInitRank", a.CommInitRank) || !sym(lib, "ncclGetUniqueId", a.GetUniqueId) ||
        !sym(lib, "ncclCommDestroy", a.CommDestroy) || !sym(lib, "ncclAllReduce", a.AllReduce) ||
        !sym(lib, "ncclGetVersion", a.GetVersion))
        return fail(SIMPLES_ECOMM, "libnccl lacks a required symbol");
    g_nccl = a;
    return SIMPLES_OK;
}

static int nccl_fail(const char* what, int rc) {
    return fail(SIMPLES_ECOMM, std::string(what) + " failed: " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
}

// simples_allreduce_fn over an NCCL communicator (user = ncclComm_t)
int nccl_allreduce_cb(void* user, double* dev_buf, size_t count, void* stream) {
    const int rc = g_nccl.AllReduce(dev_buf, dev_buf, count, kNcclFloat64, kNcclSum, user, (cudaStream_t)stream);
    return rc == 0 ? 0 : nccl_fail("ncclAllReduce", rc);
}

int comm_init_all(int ndev, const int* devs) {
    CKR(nccl_load());
    const int rc = g_nccl.CommInitAll(g_comms, ndev, devs);
    if (rc != 0) return nccl_fail("ncclCommInitAll", rc);
    return SIMPLES_OK;
}

void comm_destroy_all() {
    for (int i = 0; i < SIMPLES_MAX_DEVICES; ++i) {
        if (g_comms[i] && g_nccl.CommDestroy) g_nccl.CommDestroy(g_comms[i]);
        g_comms[i] = nullptr;
    }
    if (g_rank_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(g_rank_comm);
    g_rank_comm = nullptr;
    g_rank_nranks = 0;
}

}  // namespace sbe

extern "C" {

int simples_device_count(void) { return g_ndev; }

int simples_nccl_version(void) {
    if (nccl_load() != SIMPLES_OK) return -1;
    int v = 0;
    g_nccl.GetVersion(&v);
    return v;
}

int simples_comm_unique_id(void* id, size_t cap) {
    if (!id || cap < sizeof(NcclUniqueId)) return fail(SIMPLES_EINVAL, "simples_comm_unique_id: buffer of >= 128 bytes required");
    CKR(nccl_load());
    NcclUniqueId u;
    const int rc = g_nccl.GetUniqueId(&u);
    if (rc != 0) return nccl_fail("ncclGetUniqueId", rc);
    memcpy(id, &u, sizeof u);
    return SIMPLES_OK;
}

int simples_comm_init_rank(int nranks, int rank, const void* id) {
    if (!id || nranks < 1 || rank < 0 || rank >= nranks) return fail(SIMPLES_EINVAL, "simples_comm_init_rank: bad arguments");
    CKR(ensure_device());
    CKR(nccl_load());
    if (g_rank_comm) {
        g_nccl.CommDestroy(g_rank_comm);
        g_rank_comm = nullptr;
    }
    NcclUniqueId u;
    memcpy(&u, id, sizeof u);
    const int rc = g_nccl.CommInitRank(&g_rank_comm, nranks, u, rank);
    if (rc != 0) return nccl_fail("ncclCommInitRank", rc);
    g_rank_nranks = nranks;
    return SIMPLES_OK;
}

void simples_comm_destroy(void) {
    if (g_rank_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(g_rank_comm);
    g_rank_comm = nullptr;
    g_rank_nranks = 0;
}

int simples_comm_nranks(void) { return g_rank_nranks; }

}  // extern "C"
