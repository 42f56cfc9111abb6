// Internals shared by the translation units that implement the C ABI (engine.cu: EM_impute / do_impute,
// engine_stages.cu: the stages around them).  Not part of the public interface.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/simples_b200.h"
#include "state.h"
#include "tmap.h"

namespace sbe {
using namespace sb;

extern thread_local std::string g_err;
extern int g_device;                    // device of a one-device process (simples_init(1, ...))
extern thread_local int t_device;       // device of the calling host thread in a multi-device run (-1: g_device)
extern int g_ndev;                      // devices this process drives (0 before simples_init)
extern int g_devs[SIMPLES_MAX_DEVICES];
extern void* g_comms[SIMPLES_MAX_DEVICES];   // ncclComm_t per device (single-process multi-GPU)
extern void* g_rank_comm;               // ncclComm_t of this process in a one-process-per-GPU job, or nullptr
extern int g_rank_nranks;               // ranks of that job (0: none)
int fail(int code, const std::string& msg);
int nccl_load();
int nccl_allreduce_cb(void* comm, double* dev_buf, size_t count, void* stream);
int comm_init_all(int ndev, const int* devs);
void comm_destroy_all();

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            char buf_[512];                                                                            \
            snprintf(buf_, sizeof buf_, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,           \
                     cudaGetErrorString(e_));                                                          \
            return fail(e_ == cudaErrorMemoryAllocation ? SIMPLES_ENOMEM : SIMPLES_ECUDA, buf_);       \
        }                                                                                              \
    } while (0)

#define CKR(expr)                  \
    do {                           \
        int r_ = (expr);           \
        if (r_ != SIMPLES_OK) return r_; \
    } while (0)


inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct EvRec {
    int cls;
    cudaEvent_t a, b;
};


}  // namespace sbe

using namespace sb;
using namespace sbe;

struct simples_ctx {
    int kind = 0;   // 0 = EM, 1 = MI
    DevState s{};
    std::vector<void*> allocs;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t side = nullptr;          // side stream: cluster_prep / sstats overlap the big GEMM kernels
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool cluster_pending = false;
    // scalars of the call
    double cutoff = 0, penl = 1, sigma0 = 100, pi_alpha = 1, lower = -INFINITY, upper = INFINITY;
    int est_z = 2, max_lambda = 1, est_lam = 2, impt_it = 5, num_mc = 3;
    uint64_t seed = 0;
    uint32_t compat = SIMPLES_COMPAT_DEFAULT;
    simples_allreduce_fn allreduce = nullptr;
    void* allreduce_user = nullptr;
    bool z_null = false;
    // single-process multi-GPU run: this context holds one shard; row-sharded host matrices (z, Ef, EF, varF) are
    // addressed with the leading dimension of the FULL matrix (io_ld, 0 = the local n) through pointers that already
    // point at the shard's first row; only the first shard writes the replicated outputs (parameters, loglik)
    long long io_ld = 0;
    bool write_shared = true;
    int it = 0;             // EM iterations done
    unsigned sweep = 0;     // MC sweeps done (tape position)
    int iter_cap = 0;
    int n_alloc = 0;
    int NS_max = 1;
    long long nnz0 = 0;
    double host_mask_ms = 0.0;        // host time spent packing the Y0 <= cutoff mask (0: built on the device)
    CUtensorMap tmapY;                // Y as [n_alloc][ldG], box {16 genes, 128 cells}, SWIZZLE_128B
    CUtensorMap tmapY8;               // same, box {16 genes, 8 cells}: the per-warp tiles of the fused MC pass
    bool fused_mc = false;            // SIMPLES_FUSED_MC=1: fused impute_sweep + estep_project kernel for the MC passes (opt-in)
    double* cellsum_main = nullptr;   // [2 M0 + 1] of the main E-pass of the current iteration
    double* cellsum_last = nullptr;   // [2 M0 + 1] of the last E-pass (nz, c feed the M-step)
    double* stage = nullptr;          // staging buffer for chunked H2D / D2H
    size_t stage_elems = 0;
    unsigned long long* d_nnz = nullptr;
    // general path scratch
    double* part_g = nullptr;
    int nsplit_g = 0;
    // MI
    int mcmc = 0, burnin = 0, want_consensus = 0;
    unsigned long long* mi_gbase = nullptr;   // slot table of the compact accumulators rec1 / rec2 (state.h, SweepArgs)
    uint32_t* mi_off = nullptr;
    double *rec1 = nullptr, *rec2 = nullptr, *rEF = nullptr, *rF2 = nullptr, *rVF = nullptr, *cons = nullptr, *tot_mi = nullptr;
    // timing
    std::vector<cudaEvent_t> iter_ev;
    std::vector<double> iter_ms;
    bool profile = false;
    std::vector<EvRec> evs;
    long long kernel_launches = 0;   // kernels enqueued since the last simples_em_run began
    long long prof_launches[KC_COUNT] = {0};
    double prof_ms[KC_COUNT] = {0};

    template <typename T>
    int dalloc(T** p, size_t count, bool zero = true) {
        void* q = nullptr;
        if (count == 0) count = 1;
        CK(cudaMallocAsync(&q, count * sizeof(T), stream));   // stream-ordered pool: repeated calls reuse the blocks
        allocs.push_back(q);
        if (zero) CK(cudaMemsetAsync(q, 0, count * sizeof(T), stream));
        *p = static_cast<T*>(q);
        return SIMPLES_OK;
    }
    void prof_begin(int cls) {
        if (!profile) return;
        EvRec r;
        r.cls = cls;
        cudaEventCreate(&r.a);
        cudaEventCreate(&r.b);
        cudaEventRecord(r.a, stream);
        evs.push_back(r);
    }
    void prof_end() {
        if (!profile) return;
        cudaEventRecord(evs.back().b, stream);
    }
    void prof_collect() {
        for (auto& r : evs) {
            float ms = 0;
            cudaEventSynchronize(r.b);
            cudaEventElapsedTime(&ms, r.a, r.b);
            prof_launches[r.cls] += 1;
            prof_ms[r.cls] += ms;
            cudaEventDestroy(r.a);
            cudaEventDestroy(r.b);
        }
        evs.clear();
    }
    ~simples_ctx() {
        for (auto& r : evs) {
            cudaEventDestroy(r.a);
            cudaEventDestroy(r.b);
        }
        for (auto e : iter_ev) cudaEventDestroy(e);
        for (void* p : allocs) cudaFreeAsync(p, stream);
        if (stream) cudaStreamSynchronize(stream);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (side) cudaStreamDestroy(side);
        if (own_stream && stream) cudaStreamDestroy(stream);
    }
};

namespace sbe {

struct P {   // RAII profile scope
    simples_ctx* c;
    P(simples_ctx* c_, int cls, int nk = 1) : c(c_) {
        c->kernel_launches += nk;
        c->prof_begin(cls);
    }
    ~P() { c->prof_end(); }
};

// SIMPLES_TRACE=1: wall-clock phases of the one-shot entry points (each mark synchronises the stream)
struct Tracer {
    const char* name;
    cudaStream_t st;
    bool on;
    std::chrono::steady_clock::time_point t0;
    std::string msg;
    Tracer(const char* n, cudaStream_t s) : name(n), st(s), on(getenv("SIMPLES_TRACE") != nullptr) {
        if (on) t0 = std::chrono::steady_clock::now();
    }
    void mark(const char* what) {
        if (!on) return;
        cudaStreamSynchronize(st);
        const auto t1 = std::chrono::steady_clock::now();
        char buf[96];
        snprintf(buf, sizeof buf, " %s %.1f ms,", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        msg += buf;
        t0 = t1;
    }
    ~Tracer() {
        if (on && !msg.empty()) {
            msg.pop_back();
            fprintf(stderr, "[simples_b200] %s:%s\n", name, msg.c_str());
        }
    }
};

int ensure_device();
int allreduce(simples_ctx* c, double* buf, size_t count);
void pick_allreduce(simples_ctx* c, simples_allreduce_fn fn, void* user, long long n, long long n_total);
std::vector<double> to_rowmajor(const double* src, int rows, int cols, int ldrows, double padval);
void set_dims(sb::Dims& d, int G, int n, int M0, int K0, int MB, long long n_total, long long cell_offset);
int check_common(int G, int n, int M0, int K0, int MB);
int setup_common(simples_ctx* c, int G, int n, int M0, int K0, int MB, long long n_total, long long cell_offset,
                 const double* Y, const double* Y0, const double* pg, const double* beta, const double* sigma,
                 const double* lambda, const double* pi, const int32_t* celltype, double cutoff, bool sigma_shared);
bool sigma_is_shared(const double* sigma, int G, int M0);
int upload_z(simples_ctx* c, const double* z);
int agree_all_ranks(simples_ctx* c, long long* count);
unsigned long long host_pack_mask(const double* Y0, int G, int n, int nwords, size_t na, double cutoff, uint32_t* mask);
int e_pass(simples_ctx* c, bool update_z, bool stale_q, bool sample, bool draw_membership, bool write_W, bool write_Vg,
           double pi_const, bool have_P = false);
int run_prep(simples_ctx* c);
int run_sweep(simples_ctx* c, bool clip, double* rec1, double* rec2);
int fetch_genemajor(simples_ctx* c, const double* dev, int cols, double* host);
int fetch_transposed(simples_ctx* c, const double* dev, int rows, int cols, int ld, double* host, long long ld_host = 0);

// chunked D2H of a per-entry G x n result produced by `fill(chunk_dev_out, cell0, ncells)`.
// The staging buffer is split in two halves: the fill kernel of chunk k+1 (compute stream) overlaps the
// D2H copy of chunk k (copy stream).
template <typename F>
int fetch_big(simples_ctx* c, double* host_out, F fill) {
    const Dims& d = c->s.d;
    const size_t half_elems = c->stage_elems / 2;
    const size_t cpc = std::max<size_t>(1, half_elems / d.G);
    cudaStream_t cs;
    CK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    cudaEvent_t filled[2], copied[2];
    for (int h = 0; h < 2; ++h) {
        CK(cudaEventCreateWithFlags(&filled[h], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&copied[h], cudaEventDisableTiming));
    }
    int k = 0;
    for (size_t c0 = 0; c0 < size_t(d.n); c0 += cpc, ++k) {
        const int h = k & 1;
        const size_t nc = std::min(cpc, size_t(d.n) - c0);
        double* buf = c->stage + size_t(h) * half_elems;
        if (k >= 2) CK(cudaStreamWaitEvent(c->stream, copied[h], 0));      // half h is free again
        fill(buf, c0, nc);
        CK(cudaEventRecord(filled[h], c->stream));
        CK(cudaStreamWaitEvent(cs, filled[h], 0));
        CK(cudaMemcpyAsync(host_out + c0 * d.G, buf, nc * d.G * 8, cudaMemcpyDefault, cs));
        CK(cudaEventRecord(copied[h], cs));
    }
    CK(cudaStreamSynchronize(cs));
    CK(cudaStreamSynchronize(c->stream));
    for (int h = 0; h < 2; ++h) {
        cudaEventDestroy(filled[h]);
        cudaEventDestroy(copied[h]);
    }
    cudaStreamDestroy(cs);
    return SIMPLES_OK;
}


}  // namespace sbe
