// Host orchestration + C ABI (include/simples_b200.h) of the B200 EM imputation hot path.
// One context = one EM_impute (R/EM_impute.R:75-339) or do_impute (R/do_impute.R:29-160) call on
// this process's shard of cells.  The EM loop is enqueued on one CUDA stream without any
// host <-> device synchronisation: all data-dependent decisions (lambda eligibility, lasso active
// sets, priorB bookkeeping) happen on the device.
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

using namespace sb;

namespace {

thread_local std::string g_err;
int g_device = -1;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

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

const char* kKernelNames[KC_COUNT] = {"estep_project", "cell_post",  "impute_sweep",     "mstep_stats", "sstats", "reduce",
                                      "cluster_prep",  "gene_solve", "lambda_mu_pi",     "misc",        "allreduce"};

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct EvRec {
    int cls;
    cudaEvent_t a, b;
};

}  // namespace

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
    int it = 0;             // EM iterations done
    unsigned sweep = 0;     // MC sweeps done (tape position)
    int iter_cap = 0;
    int n_alloc = 0;
    int NS_max = 1;
    long long nnz0 = 0;
    CUtensorMap tmapY;                // Y as [n_alloc][ldG], box {16 genes, 128 cells}, SWIZZLE_128B
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

namespace {

struct P {   // RAII profile scope
    simples_ctx* c;
    P(simples_ctx* c_, int cls, int nk = 1) : c(c_) {
        c->kernel_launches += nk;
        c->prof_begin(cls);
    }
    ~P() { c->prof_end(); }
};

int ensure_device() {
    if (g_device < 0) {
        int cnt = 0;
        cudaError_t e = cudaGetDeviceCount(&cnt);
        if (e != cudaSuccess || cnt == 0)
            return fail(SIMPLES_ENODEV, std::string("no CUDA device available (no CPU fallback): ") +
                                            (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0"));
        int dev = 0;
        cudaGetDevice(&dev);
        g_device = dev;
    }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, g_device));
    if (prop.major != 10) {
        char buf[256];
        snprintf(buf, sizeof buf, "device %d (%s) is sm_%d%d; this library is built for sm_100a (B200) only", g_device,
                 prop.name, prop.major, prop.minor);
        return fail(SIMPLES_ENODEV, buf);
    }
    CK(cudaSetDevice(g_device));
    static int pool_dev = -1;
    if (pool_dev != g_device) {            // keep freed blocks cached in the default pool between calls
        cudaMemPool_t pool;
        CK(cudaDeviceGetDefaultMemPool(&pool, g_device));
        unsigned long long thr = ~0ull;
        CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        pool_dev = g_device;
    }
    return SIMPLES_OK;
}

int allreduce(simples_ctx* c, double* buf, size_t count) {
    if (!c->allreduce) return SIMPLES_OK;
    P p(c, KC_COMM, 0);
    if (c->allreduce(c->allreduce_user, buf, count, (void*)c->stream) != 0)
        return fail(SIMPLES_ECOMM, "all-reduce callback failed");
    return SIMPLES_OK;
}

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

// host col-major (rows x cols, ld = rows) -> gene-major padded [ldrows][cols]
std::vector<double> to_rowmajor(const double* src, int rows, int cols, int ldrows, double padval) {
    std::vector<double> out(size_t(ldrows) * cols, padval);
    for (int c = 0; c < cols; ++c)
        for (int r = 0; r < rows; ++r) out[size_t(r) * cols + c] = src[size_t(c) * rows + r];
    return out;
}

void set_dims(Dims& d, int G, int n, int M0, int K0, int MB, long long n_total, long long cell_offset) {
    d.G = G;
    d.ldG = round_up(G, GPAD);
    d.n = n;
    d.M0 = M0;
    d.K0 = K0;
    d.MB = MB;
    d.NS = 1;
    d.NQ = K0 + M0;
    d.NQP = round_up(d.NQ, 8);
    d.QS = d.NQP + 4;
    d.NF = K0;
    d.NFP = round_up(K0, 4);
    d.FS = round_up(K0, 8) + 4;
    d.NV = K0 + M0 + 1;
    d.NVP = round_up(d.NV, 8);
    d.VS = d.NVP + 4;
    d.n_total = n_total > 0 ? n_total : n;
    d.cell_offset = cell_offset;
}

int check_common(int G, int n, int M0, int K0, int MB) {
    if (G < 1 || n < 1) return fail(SIMPLES_EINVAL, "G and n must be >= 1");
    if (K0 < 1 || K0 > MAXK) return fail(SIMPLES_EINVAL, "K0 must be in [1, 32]");
    if (M0 < 1 || M0 > MAXM) return fail(SIMPLES_EINVAL, "M0 must be in [1, 32]");
    if (MB < 1) return fail(SIMPLES_EINVAL, "MB (ncol(pg)) must be >= 1");
    return SIMPLES_OK;
}

// Allocate and upload everything EM and MI share.
int setup_common(simples_ctx* c, int G, int n, int M0, int K0, int MB, long long n_total, long long cell_offset,
                 const double* Y, const double* Y0, const double* pg, const double* beta, const double* sigma,
                 const double* lambda, const double* pi, const int32_t* celltype, double cutoff, bool sigma_shared) {
    DevState& s = c->s;
    set_dims(s.d, G, n, M0, K0, MB, n_total, cell_offset);
    Dims& d = s.d;
    c->NS_max = sigma_shared ? 1 : M0;
    d.NS = c->NS_max;
    c->n_alloc = round_up(n, 128);
    const size_t na = c->n_alloc;
    const int ldG = d.ldG, nwords = ldG / 32;

    CKR(c->dalloc(&s.Y, na * ldG));
    CKR(c->dalloc(&s.mask, na * nwords));
    CK(make_tmap_2d(&c->tmapY, s.Y, ldG, c->n_alloc, 16, 128));
    CKR(c->dalloc(&s.celltype0, na));
    CKR(c->dalloc(&s.beta, size_t(ldG) * K0));
    CKR(c->dalloc(&s.sigma, size_t(ldG) * M0));
    CKR(c->dalloc(&s.mu, size_t(ldG) * M0));
    CKR(c->dalloc(&s.pg, size_t(ldG) * MB));
    CKR(c->dalloc(&s.gmean, size_t(ldG)));
    CKR(c->dalloc(&s.gsd, size_t(ldG)));
    CKR(c->dalloc(&s.lam, size_t(M0) * K0));
    CKR(c->dalloc(&s.pi, size_t(M0)));
    CKR(c->dalloc(&s.z, na * M0));
    CKR(c->dalloc(&s.Q, size_t(c->NS_max) * ldG * d.QS));
    CKR(c->dalloc(&s.isg, size_t(c->NS_max) * ldG));
    CKR(c->dalloc(&s.Qs, size_t(ldG) * d.FS));
    CKR(c->dalloc(&s.sdsA, size_t(ldG) * M0));
    CKR(c->dalloc(&s.isdA, size_t(ldG) * M0));
    CKR(c->dalloc(&s.musdA, size_t(ldG) * M0));
    CKR(c->dalloc(&s.igs, size_t(ldG)));
    s.n_alloc = c->n_alloc;
    CKR(c->dalloc(&s.P, size_t(c->NS_max) * na * d.NQP));
    CKR(c->dalloc(&s.S2, size_t(c->NS_max) * na));
    CKR(c->dalloc(&s.V, na * d.VS));
    CKR(c->dalloc(&s.zsum, na));
    CKR(c->dalloc(&s.Fh, na * d.FS));
    CKR(c->dalloc(&s.W, size_t(M0) * na * K0));
    CKR(c->dalloc(&s.im, na));
    CKR(c->dalloc(&s.cc.Mm, size_t(M0) * K0 * K0));
    CKR(c->dalloc(&s.cc.Mh, size_t(M0) * K0 * K0));
    CKR(c->dalloc(&s.cc.MmP, size_t(M0) * (4 * ((K0 + 3) / 4)) * (8 * ((K0 + 7) / 8) + 4)));
    CKR(c->dalloc(&s.cc.dt, size_t(M0)));
    CKR(c->dalloc(&s.cc.off, size_t(M0) * K0));
    CKR(c->dalloc(&s.cc.offq, size_t(M0) * K0));
    CKR(c->dalloc(&s.cc.cmu, size_t(M0)));
    CKR(c->dalloc(&s.cc.logpi, size_t(M0)));
    s.ncta_cell = std::min((n + 63) / 64, NUM_SMS_B200 * 8);
    CKR(c->dalloc(&s.part_cell, size_t(s.ncta_cell) * (2 * M0 + 1)));
    CKR(c->dalloc(&c->cellsum_main, size_t(2 * M0 + 1)));
    CKR(c->dalloc(&c->cellsum_last, size_t(2 * M0 + 1)));
    CKR(c->dalloc(&c->d_nnz, 1));

    // staging buffer: whole cells, at most 32 Mi doubles (256 MB)
    size_t cells_per_chunk = std::max<size_t>(1, (size_t(32) << 20) / size_t(G));
    cells_per_chunk = std::min<size_t>(cells_per_chunk, size_t(n));
    c->stage_elems = cells_per_chunk * G;
    CKR(c->dalloc(&c->stage, c->stage_elems, false));

    // ---- uploads ----
    if (ldG == G) {
        CK(cudaMemcpyAsync(s.Y, Y, size_t(n) * G * 8, cudaMemcpyDefault, c->stream));
    } else {
        CK(cudaMemcpy2DAsync(s.Y, size_t(ldG) * 8, Y, size_t(G) * 8, size_t(G) * 8, n, cudaMemcpyDefault, c->stream));
    }
    for (size_t c0 = 0; c0 < size_t(n); c0 += cells_per_chunk) {
        const size_t nc = std::min(cells_per_chunk, size_t(n) - c0);
        CK(cudaMemcpyAsync(c->stage, Y0 + c0 * G, nc * G * 8, cudaMemcpyDefault, c->stream));
        launch_build_mask(c->stage, s.mask, int(nc), int(c0), c->n_alloc, G, ldG, cutoff, c->d_nnz, c->stream);
    }
    {
        std::vector<double> hb = to_rowmajor(beta, G, K0, ldG, 0.0);
        std::vector<double> hs = to_rowmajor(sigma, G, M0, ldG, 1.0);
        std::vector<double> hp = to_rowmajor(pg, G, MB, ldG, 0.5);
        std::vector<double> hl = to_rowmajor(lambda, M0, K0, M0, 0.0);
        std::vector<int32_t> ct(n, 0);
        if (celltype) {
            for (int i = 0; i < n; ++i) {
                if (celltype[i] < 1 || celltype[i] > MB) return fail(SIMPLES_EINVAL, "celltype must be in 1..ncol(pg)");
                ct[i] = celltype[i] - 1;
            }
        }
        CK(cudaMemcpyAsync(s.beta, hb.data(), hb.size() * 8, cudaMemcpyDefault, c->stream));
        CK(cudaMemcpyAsync(s.sigma, hs.data(), hs.size() * 8, cudaMemcpyDefault, c->stream));
        CK(cudaMemcpyAsync(s.pg, hp.data(), hp.size() * 8, cudaMemcpyDefault, c->stream));
        CK(cudaMemcpyAsync(s.lam, hl.data(), hl.size() * 8, cudaMemcpyDefault, c->stream));
        CK(cudaMemcpyAsync(s.pi, pi, size_t(M0) * 8, cudaMemcpyDefault, c->stream));
        CK(cudaMemcpyAsync(s.celltype0, ct.data(), size_t(n) * 4, cudaMemcpyDefault, c->stream));
        CK(cudaStreamSynchronize(c->stream));   // host vectors go out of scope
    }
    unsigned long long h_nnz = 0;
    CK(cudaMemcpy(&h_nnz, c->d_nnz, 8, cudaMemcpyDefault));
    c->nnz0 = (long long)h_nnz;
    {   // NA/NaN/Inf in Y would poison every statistic silently (the reference's glmnet call stops on them)
        unsigned long long* d_bad = nullptr;
        CKR(c->dalloc(&d_bad, 1));
        launch_count_nonfinite(s.Y, na * ldG, d_bad, c->stream);
        unsigned long long h_bad = 0;
        CK(cudaMemcpyAsync(&h_bad, d_bad, 8, cudaMemcpyDefault, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (h_bad) return fail(SIMPLES_EINVAL, "Y contains " + std::to_string(h_bad) + " non-finite value(s)");
    }
    return SIMPLES_OK;
}

bool sigma_is_shared(const double* sigma, int G, int M0) {
    for (int m = 1; m < M0; ++m)
        for (int g = 0; g < G; ++g)
            if (sigma[size_t(m) * G + g] != sigma[g]) return false;
    return true;
}

// upload an n x M0 col-major host matrix into the [n][M0] device layout (through the staging buffer)
int upload_z(simples_ctx* c, const double* z) {
    DevState& s = c->s;
    const int n = s.d.n, M0 = s.d.M0;
    double* tmp = nullptr;
    CK(cudaMallocAsync(&tmp, size_t(n) * M0 * 8, c->stream));
    CK(cudaMemcpyAsync(tmp, z, size_t(n) * M0 * 8, cudaMemcpyDefault, c->stream));
    launch_transpose(tmp, s.z, M0, n, n, M0, c->stream);
    CK(cudaFreeAsync(tmp, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SIMPLES_OK;
}

// one E-pass: projections + cell epilogue
int e_pass(simples_ctx* c, bool update_z, bool stale_q, bool sample, bool draw_membership, bool write_W, bool write_Vg,
           double pi_const) {
    DevState& s = c->s;
    const Dims& d = s.d;
    for (int sc = 0; sc < d.NS; ++sc) {
        P p(c, KC_PROJ);
        ProjArgs pa{s.Y, s.Q + size_t(sc) * d.ldG * d.QS, s.isg + size_t(sc) * d.ldG, s.P + size_t(sc) * d.n * d.NQP,
                    s.S2 + size_t(sc) * d.n, d.n, d.ldG, d.NQP, d.QS};
        launch_proj(pa, c->tmapY, c->stream);
    }
    if (c->cluster_pending) {
        CK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
        c->cluster_pending = false;
    }
    {
        P p(c, KC_CELL);
        CellArgs ca{};
        ca.d = d;
        ca.P = s.P;
        ca.S2 = s.S2;
        ca.cc = s.cc;
        ca.z = s.z;
        ca.V = s.V;
        ca.zsum = s.zsum;
        ca.W = s.W;
        ca.Fh = s.Fh;
        ca.Vg = s.Vg;
        ca.im = s.im;
        ca.part = s.part_cell;
        ca.ncta = s.ncta_cell;
        ca.update_z = update_z;
        ca.stale_q = stale_q;
        ca.write_W = write_W;
        ca.write_Vg = write_Vg && s.Vg;
        ca.sample = sample;
        ca.draw_membership = draw_membership;
        ca.ll_const = double(d.G) * std::log(2.0 * pi_const);
        ca.seed = c->seed;
        ca.sweep = c->sweep;
        launch_cell(ca, c->stream);
    }
    CK(cudaGetLastError());
    return SIMPLES_OK;
}

int run_prep(simples_ctx* c) {
    DevState& s = c->s;
    P p(c, KC_PREP, 2);
    PrepArgs pa{s.d, s.beta, s.sigma, s.mu, s.lam, s.pi, s.gmean, s.gsd, s.Q, s.isg, s.Qs, s.sdsA, s.isdA, s.musdA, s.igs, s.cc, 0};
    if (c->profile) {                      // serial on the main stream so the per-class events mean something
        launch_prep(pa, c->stream, c->stream);
    } else {
        if (!c->side) {
            CK(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
        }
        CK(cudaEventRecord(c->ev_fork, c->stream));
        CK(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
        launch_prep(pa, c->stream, c->side);
        CK(cudaEventRecord(c->ev_join, c->side));
        c->cluster_pending = true;         // joined right before the first cell_post
    }
    CK(cudaGetLastError());
    return SIMPLES_OK;
}

int run_sweep(simples_ctx* c, bool clip, double* rec1, double* rec2) {
    DevState& s = c->s;
    P p(c, KC_SWEEP);
    SweepArgs sa{};
    sa.d = s.d;
    sa.Y = s.Y;
    sa.mask = s.mask;
    sa.Qs = s.Qs;
    sa.Fh = s.Fh;
    sa.sdsA = s.sdsA;
    sa.isdA = s.isdA;
    sa.musdA = s.musdA;
    sa.pg = s.pg;
    sa.gmean = s.gmean;
    sa.igs = s.igs;
    sa.n_alloc = s.n_alloc;
    sa.im = s.im;
    sa.celltype0 = s.celltype0;
    sa.cutoff = c->cutoff;
    sa.lower = c->lower;
    sa.upper = c->upper;
    sa.clip = clip;
    sa.seed = c->seed;
    sa.sweep = c->sweep;
    sa.rec1 = rec1;
    sa.rec2 = rec2;
    launch_sweep(sa, c->stream);
    CK(cudaGetLastError());
    return SIMPLES_OK;
}

int em_iteration(simples_ctx* c) {
    DevState& s = c->s;
    Dims& d = s.d;
    const int it = c->it + 1;
    const double pi_const = (c->compat & SIMPLES_COMPAT_PI) ? 3.14159265359 : M_PI;
    const bool general = d.NS > 1;
    const bool update_z = (it >= c->est_z) || c->z_null;
    const bool sampling = (it >= c->impt_it + 1) && c->num_mc > 0;
    const bool draw_mem = d.M0 > 1;

    CKR(run_prep(c));
    CKR(e_pass(c, update_z, false, sampling, draw_mem, !sampling, general && !sampling, pi_const));
    {
        P p(c, KC_REDUCE);
        launch_finalize_cell(s.part_cell, s.ncta_cell, d.M0, c->cellsum_main, c->stream);
    }
    if (sampling) {
        for (int mc = 0; mc < c->num_mc; ++mc) {
            CKR(run_sweep(c, true, nullptr, nullptr));
            c->sweep += 1;
            const bool last = (mc == c->num_mc - 1);
            CKR(e_pass(c, update_z, (c->compat & SIMPLES_COMPAT_STALE_DT) != 0, !last, draw_mem, last, general && last,
                       pi_const));
        }
    }
    c->z_null = false;

    // ---- sufficient statistics ----
    double* stats;
    size_t count;
    if (!general) {
        const bool overlap = !c->profile && c->side != nullptr;
        if (overlap) {      // sstats (latency-bound, few CTAs) runs beside the mstats GEMM
            CK(cudaEventRecord(c->ev_fork, c->stream));
            CK(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
        }
        {
            P p(c, KC_SSTATS);
            SstatsArgs sa{d, s.W, s.z, s.part_ss, s.nsplit_ss};
            launch_sstats(sa, overlap ? c->side : c->stream);
        }
        if (overlap) CK(cudaEventRecord(c->ev_join, c->side));
        {
            P p(c, KC_MSTATS);
            MstatsArgs ma{s.Y, s.V, s.zsum, s.part_ms, d.n, d.ldG, d.NVP, d.VS, s.splitK};
            launch_mstats(ma, c->stream);
        }
        if (overlap) CK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
        {
            P p(c, KC_REDUCE, 2);
            launch_finalize_cell(s.part_cell, s.ncta_cell, d.M0, c->cellsum_last, c->stream);
            FinalizeArgs fa{d, s.part_ms, c->cellsum_last, s.part_ss, s.splitK, 1, s.nsplit_ss, s.stats, 0, 0};
            launch_finalize_stats(fa, c->stream);
        }
        stats = s.stats;
        count = s.stats_count;
    } else {
        const int ncg = d.K0 * d.M0 + 2 * d.M0;
        const size_t nC = size_t(d.ldG) * ncg, nU2 = size_t(d.ldG) * d.M0;
        {
            P p(c, KC_MSTATS, 4);
            launch_ygemm_simple(s.Y, s.Vg, c->part_g, d.n, d.ldG, ncg, ncg, 0, c->nsplit_g, c->stream);
            launch_reduce_parts(c->part_g, s.statsg, nC, c->nsplit_g, c->stream);
            launch_ygemm_simple(s.Y, s.Vg + size_t(d.K0) * d.M0, c->part_g, d.n, d.ldG, d.M0, ncg, 1, c->nsplit_g, c->stream);
            launch_reduce_parts(c->part_g, s.statsg + nC, nU2, c->nsplit_g, c->stream);
        }
        {
            P p(c, KC_SSTATS);
            SstatsArgs sa{d, s.W, s.z, s.part_ss, s.nsplit_ss};
            launch_sstats(sa, c->stream);
        }
        {
            P p(c, KC_REDUCE, 2);
            launch_finalize_cell(s.part_cell, s.ncta_cell, d.M0, c->cellsum_last, c->stream);
            FinalizeArgs fa{d, nullptr, c->cellsum_last, s.part_ss, 0, 1, s.nsplit_ss, s.statsg, nC + nU2, 1};
            launch_finalize_stats(fa, c->stream);
        }
        stats = s.statsg;
        count = s.statsg_count;
    }
    // tot[it] comes from the main E-pass, not from the MC passes (R/EM_impute.R:144 vs :215-224)
    CK(cudaMemcpyAsync(stats + count - 1, c->cellsum_main + 2 * d.M0, 8, cudaMemcpyDeviceToDevice, c->stream));
    CKR(allreduce(c, stats, count));

    // ---- M-step ----
    {
        P p(c, KC_GENE, 3);
        MstepArgs ma{};
        ma.d = d;
        ma.stats = stats;
        ma.general = general;
        ma.beta = s.beta;
        ma.sigma = s.sigma;
        ma.mu = s.mu;
        ma.lam = s.lam;
        ma.pi = s.pi;
        ma.Mm = s.cc.Mm;
        ma.loglik_slot = s.loglik + (it - 1);
        ma.priorB = s.priorB;
        ma.part_gene = s.part_gene;
        ma.ncta_gene = s.ncta_gene;
        ma.penl = c->penl;
        ma.sigma0 = c->sigma0;
        ma.pi_alpha = c->pi_alpha;
        ma.do_lambda = c->max_lambda && it >= c->est_lam;
        ma.compat_priorb_prev = (c->compat & SIMPLES_COMPAT_PRIORB_PREV) != 0;
        launch_mstep(ma, c->stream);
    }
    CK(cudaGetLastError());
    d.NS = 1;   // sigma is cluster-shared after every M-step (R/EM_impute.R:269)
    c->it = it;
    return SIMPLES_OK;
}

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

// gene-major [ldG][cols] device -> col-major host G x cols
int fetch_genemajor(simples_ctx* c, const double* dev, int cols, double* host) {
    const Dims& d = c->s.d;
    std::vector<double> tmp(size_t(d.ldG) * cols);
    CK(cudaMemcpyAsync(tmp.data(), dev, tmp.size() * 8, cudaMemcpyDefault, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int cidx = 0; cidx < cols; ++cidx)
        for (int g = 0; g < d.G; ++g) host[size_t(cidx) * d.G + g] = tmp[size_t(g) * cols + cidx];
    return SIMPLES_OK;
}

// device [rows][cols] -> host col-major rows x cols via device transpose
int fetch_transposed(simples_ctx* c, const double* dev, int rows, int cols, int ld, double* host) {
    double* tmp = nullptr;
    CK(cudaMallocAsync(&tmp, size_t(rows) * cols * 8, c->stream));
    launch_transpose(dev, tmp, rows, cols, ld, rows, c->stream);
    CK(cudaMemcpyAsync(host, tmp, size_t(rows) * cols * 8, cudaMemcpyDefault, c->stream));
    CK(cudaFreeAsync(tmp, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SIMPLES_OK;
}

}  // namespace

// =====================================================================================================
extern "C" {

int simples_abi_version(void) { return SIMPLES_ABI_VERSION; }

const char* simples_last_error(void) { return g_err.c_str(); }

int simples_init(int ndev, const int* devs) {
    if (ndev != 1 || !devs) return fail(SIMPLES_EINVAL, "simples_init: exactly one device per process (one process per GPU)");
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt == 0) return fail(SIMPLES_ENODEV, "simples_init: no CUDA device available (no CPU fallback)");
    if (devs[0] < 0 || devs[0] >= cnt) return fail(SIMPLES_EINVAL, "simples_init: device index out of range");
    g_device = devs[0];
    return ensure_device();
}

void simples_shutdown(void) { g_device = -1; }

int simples_kernel_count(void) { return KC_COUNT; }
const char* simples_kernel_name(int k) { return (k >= 0 && k < KC_COUNT) ? kKernelNames[k] : ""; }

void simples_ctx_destroy(simples_ctx* ctx) {
    if (!ctx) return;
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    delete ctx;
}

int simples_em_create(const simples_em_args* a, simples_ctx** out) {
    if (!a || !out) return fail(SIMPLES_EINVAL, "null argument");
    *out = nullptr;
    CKR(check_common(a->G, a->n, a->M0, a->K0, a->MB));
    if (!a->Y || !a->Y0 || !a->pg || !a->beta || !a->sigma || !a->lambda || !a->pi)
        return fail(SIMPLES_EINVAL, "Y, Y0, pg, beta, sigma, lambda and pi are required");
    if (a->iter < 0 || a->num_mc < 0) return fail(SIMPLES_EINVAL, "iter and num_mc must be >= 0");
    if (a->n_total != 0 && a->n_total < a->n) return fail(SIMPLES_EINVAL, "n_total < n");
    CKR(ensure_device());

    simples_ctx* c = new simples_ctx();
    std::unique_ptr<simples_ctx> guard(c);
    c->kind = 0;
    if (a->stream) {
        c->stream = (cudaStream_t)a->stream;
    } else {
        CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    c->cutoff = a->cutoff;
    c->penl = a->penl;
    c->sigma0 = a->sigma0;
    c->pi_alpha = a->pi_alpha;
    c->lower = a->lower;
    c->upper = a->upper;
    c->est_z = a->est_z;
    c->max_lambda = a->max_lambda;
    c->est_lam = a->est_lam;
    c->impt_it = a->impt_it;
    c->num_mc = a->num_mc;
    c->seed = a->seed;
    c->compat = a->ref_compat_flags;
    c->allreduce = a->allreduce;
    c->allreduce_user = a->allreduce_user;
    c->z_null = (a->z == nullptr);

    const bool shared = sigma_is_shared(a->sigma, a->G, a->M0);
    CKR(setup_common(c, a->G, a->n, a->M0, a->K0, a->MB, a->n_total, a->cell_offset, a->Y, a->Y0, a->pg, a->beta, a->sigma,
                     a->lambda, a->pi, a->celltype, a->cutoff, shared));
    DevState& s = c->s;
    Dims& d = s.d;
    const int M0 = d.M0, K0 = d.K0, ldG = d.ldG;

    // statistics buffers
    s.splitK = mstats_pick_splitk(ldG);
    s.stats_count = size_t(ldG) * d.NVP + ldG + size_t(M0) * K0 * K0 + size_t(M0) * K0 + 2 * M0 + 1;
    CKR(c->dalloc(&s.stats, s.stats_count));
    CKR(c->dalloc(&s.part_ms, size_t(s.splitK) * ldG * (d.NVP + 1)));
    s.nsplit_ss = std::max(1, (8 * NUM_SMS_B200) / M0);   // warp-level partial slots of k_sstats
    CKR(c->dalloc(&s.part_ss, size_t(s.nsplit_ss) * M0 * (K0 * K0 + K0)));
    s.ncta_gene = mstep_gene_ctas(d.G);
    CKR(c->dalloc(&s.part_gene, size_t(s.ncta_gene)));
    CKR(c->dalloc(&s.priorB, 2));
    c->iter_cap = std::max(a->iter, 1);
    CKR(c->dalloc(&s.loglik, size_t(c->iter_cap)));
    if (!shared) {
        const int ncg = K0 * M0 + 2 * M0;
        CKR(c->dalloc(&s.Vg, size_t(c->n_alloc) * ncg));
        s.statsg_count = size_t(ldG) * ncg + size_t(ldG) * M0 + size_t(M0) * K0 * K0 + size_t(M0) * K0 + 2 * M0 + 1;
        CKR(c->dalloc(&s.statsg, s.statsg_count));
        c->nsplit_g = 8;
        CKR(c->dalloc(&c->part_g, size_t(c->nsplit_g) * ldG * ncg));
    }

    // ---- prologue (R/EM_impute.R:82-102) ----
    if (std::isfinite(a->lower) || std::isfinite(a->upper)) launch_clip(s.Y, d.n, d.G, ldG, a->lower, a->upper, c->stream);
    {
        const int nsplit = 32;
        double* part = nullptr;
        CKR(c->dalloc(&part, size_t(nsplit) * ldG));
        launch_rowsum(s.Y, part, d.n, ldG, nsplit, c->stream);
        launch_reduce_parts(part, s.gmean, ldG, nsplit, c->stream);
        CKR(allreduce(c, s.gmean, ldG));
        launch_scale(s.gmean, ldG, 1.0 / double(d.n_total), c->stream);     // gene_mean <- rowMeans(Y)
        launch_fill(s.gsd, ldG, 1.0, c->stream);                            // gene_sd <- 1
        launch_center(s.Y, s.gmean, s.gsd, d.n, ldG, 0, 0, c->stream);      // Y <- Y - rowMeans(Y)
    }
    if (a->z) CKR(upload_z(c, a->z));
    if (a->mu) {
        std::vector<double> hm = to_rowmajor(a->mu, d.G, M0, ldG, 0.0);
        CK(cudaMemcpyAsync(s.mu, hm.data(), hm.size() * 8, cudaMemcpyDefault, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    } else if (M0 > 1 && a->z) {
        // mu <- Y %*% z / n   (R/EM_impute.R:95, Q3)
        const int nsplit = 16;
        double* part = nullptr;
        CKR(c->dalloc(&part, size_t(nsplit) * ldG * M0));
        launch_ygemm_simple(s.Y, s.z, part, d.n, ldG, M0, M0, 0, nsplit, c->stream);
        launch_reduce_parts(part, s.mu, size_t(ldG) * M0, nsplit, c->stream);
        CKR(allreduce(c, s.mu, size_t(ldG) * M0));
        if (c->compat & SIMPLES_COMPAT_MU_DIV_N) {
            launch_scale(s.mu, size_t(ldG) * M0, 1.0 / double(d.n_total), c->stream);
        } else {
            return fail(SIMPLES_EINVAL, "ref_compat_flags without SIMPLES_COMPAT_MU_DIV_N requires an explicit mu");
        }
    }   // else mu = 0 (already zeroed)
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    *out = guard.release();
    return SIMPLES_OK;
}

int simples_em_set_profile(simples_ctx* c, int enable) {
    if (!c) return fail(SIMPLES_EINVAL, "null ctx");
    c->profile = enable != 0;
    return SIMPLES_OK;
}

int simples_em_run(simples_ctx* c, int iters) {
    if (!c || c->kind != 0) return fail(SIMPLES_EINVAL, "not an EM context");
    if (iters < 0) return fail(SIMPLES_EINVAL, "iters < 0");
    CKR(ensure_device());
    if (c->it + iters > c->iter_cap) {   // grow the loglik array
        double* nl = nullptr;
        const int cap = c->it + iters;
        CKR(c->dalloc(&nl, size_t(cap)));
        CK(cudaMemcpyAsync(nl, c->s.loglik, size_t(c->it) * 8, cudaMemcpyDeviceToDevice, c->stream));
        c->s.loglik = nl;
        c->iter_cap = cap;
    }
    for (auto e : c->iter_ev) cudaEventDestroy(e);
    c->iter_ev.clear();
    c->iter_ms.clear();
    c->kernel_launches = 0;
    for (int k = 0; k < KC_COUNT; ++k) {
        c->prof_launches[k] = 0;
        c->prof_ms[k] = 0;
    }
    cudaEvent_t e0;
    CK(cudaEventCreate(&e0));
    CK(cudaEventRecord(e0, c->stream));
    c->iter_ev.push_back(e0);
    for (int i = 0; i < iters; ++i) {
        CKR(em_iteration(c));
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        CK(cudaEventRecord(e, c->stream));
        c->iter_ev.push_back(e);
    }
    return SIMPLES_OK;
}

int simples_em_sync(simples_ctx* c) {
    if (!c) return fail(SIMPLES_EINVAL, "null ctx");
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    if (c->iter_ev.size() > 1 && c->iter_ms.empty()) {
        for (size_t i = 1; i < c->iter_ev.size(); ++i) {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, c->iter_ev[i - 1], c->iter_ev[i]));
            c->iter_ms.push_back(ms);
        }
    }
    c->prof_collect();
    if (getenv("SIMPLES_TRACE")) {
        int d[4];
        lasso_debug_fetch(d, true);
        if (d[0])
            fprintf(stderr, "[simples_b200] lasso: %d solves, %.1f coordinate-descent sweeps on average, max %d, %d at the cap\n",
                    d[0], double(d[1]) / d[0], d[2], d[3]);
    }
    return SIMPLES_OK;
}

int simples_em_iter_ms(simples_ctx* c, double* ms, int cap) {
    if (!c) return fail(SIMPLES_EINVAL, "null ctx");
    CKR(simples_em_sync(c));
    const int k = std::min<int>(cap, int(c->iter_ms.size()));
    for (int i = 0; i < k; ++i) ms[i] = c->iter_ms[i];
    return int(c->iter_ms.size());
}

int simples_em_profile(simples_ctx* c, int64_t* launches, double* total_ms, int cap) {
    if (!c) return fail(SIMPLES_EINVAL, "null ctx");
    CKR(simples_em_sync(c));
    for (int k = 0; k < std::min<int>(cap, KC_COUNT); ++k) {
        launches[k] = c->prof_launches[k];
        total_ms[k] = c->prof_ms[k];
    }
    return KC_COUNT;
}

int64_t simples_em_nnz0(simples_ctx* c) { return c ? c->nnz0 : -1; }
int64_t simples_em_kernel_launches(simples_ctx* c) { return c ? c->kernel_launches : -1; }

int simples_em_fetch(simples_ctx* c, simples_em_out* o) {
    if (!c || !o || c->kind != 0) return fail(SIMPLES_EINVAL, "not an EM context");
    CKR(simples_em_sync(c));
    DevState& s = c->s;
    const Dims& d = s.d;
    const int M0 = d.M0, K0 = d.K0;
    if (o->loglik && c->it > 0) CK(cudaMemcpy(o->loglik, s.loglik, size_t(c->it) * 8, cudaMemcpyDefault));
    if (o->pi) CK(cudaMemcpy(o->pi, s.pi, size_t(M0) * 8, cudaMemcpyDefault));
    if (o->mu) CKR(fetch_genemajor(c, s.mu, M0, o->mu));
    if (o->sigma) CKR(fetch_genemajor(c, s.sigma, M0, o->sigma));
    if (o->beta) CKR(fetch_genemajor(c, s.beta, K0, o->beta));
    if (o->lambda) {
        std::vector<double> t(size_t(M0) * K0);
        CK(cudaMemcpy(t.data(), s.lam, t.size() * 8, cudaMemcpyDefault));
        for (int k = 0; k < K0; ++k)
            for (int m = 0; m < M0; ++m) o->lambda[size_t(k) * M0 + m] = t[size_t(m) * K0 + k];
    }
    if (o->z) CKR(fetch_transposed(c, s.z, d.n, M0, M0, o->z));
    if (o->Ef)
        for (int m = 0; m < M0; ++m)
            CKR(fetch_transposed(c, s.W + size_t(m) * d.n * K0, d.n, K0, K0, o->Ef + size_t(m) * d.n * K0));
    if (o->Varf) CK(cudaMemcpy(o->Varf, s.cc.Mm, size_t(M0) * K0 * K0 * 8, cudaMemcpyDefault));   // symmetric
    if (o->Y)
        CKR(fetch_big(c, o->Y, [&](double* dst, size_t c0, size_t nc) {
            launch_uncenter(s.Y + c0 * d.ldG, dst, s.gmean, s.gsd, int(nc), d.G, d.ldG, c->stream);   // :335
        }));
    if (o->geneM) CK(cudaMemcpy(o->geneM, s.gmean, size_t(d.G) * 8, cudaMemcpyDefault));
    if (o->geneSd) CK(cudaMemcpy(o->geneSd, s.gsd, size_t(d.G) * 8, cudaMemcpyDefault));
    if (o->priorB) CK(cudaMemcpy(o->priorB, s.priorB, 8, cudaMemcpyDefault));
    return SIMPLES_OK;
}

int simples_em_yimp0(simples_ctx* c, double* out) {
    if (!c || !out || c->kind != 0) return fail(SIMPLES_EINVAL, "not an EM context");
    if (c->it < 1) return fail(SIMPLES_EINVAL, "simples_em_yimp0 needs at least one completed EM iteration");
    CKR(simples_em_sync(c));
    DevState& s = c->s;
    const Dims& d = s.d;
    return fetch_big(c, out, [&](double* dst, size_t c0, size_t nc) {
        launch_yimp0(d, s.mu, s.beta, s.z, s.V, s.gmean, s.gsd, dst, int(c0), int(nc), c->stream);
    });
}

int simples_bic(double loglik_last, int64_t n, int32_t G, int32_t K0, int32_t M0, double* bic, double* bic0) {
    if (n < 1 || G < 1) return fail(SIMPLES_EINVAL, "n and G must be >= 1");
    const double dn = double(n), dG = double(G);
    if (bic0) *bic0 = -2.0 * loglik_last + std::log(dn) * (2.0 * dG * M0 + dG * (K0 + 1) + K0 * (M0 - 1));      // R/SIMPLE.R:429
    if (bic)                                                                                                    // R/SIMPLE.R:430
        *bic = -2.0 * loglik_last + std::log(dn) * (2.0 * dG * M0 + dG) + std::log(dn * dG) * K0 * (M0 - 1) +
               K0 * std::log((dG * dn) / (dG + dn)) * (dG + dn);
    return SIMPLES_OK;
}

int simples_em_impute(const simples_em_args* a, simples_em_out* o) {
    using clk = std::chrono::steady_clock;
    const bool trace = getenv("SIMPLES_TRACE") != nullptr;
    const auto t0 = clk::now();
    simples_ctx* c = nullptr;
    CKR(simples_em_create(a, &c));
    const auto t1 = clk::now();
    int r = simples_em_run(c, a->iter);
    if (r == SIMPLES_OK) r = simples_em_sync(c);
    const auto t2 = clk::now();
    if (r == SIMPLES_OK && o) r = simples_em_fetch(c, o);
    if (r == SIMPLES_OK) r = simples_em_sync(c);
    const auto t3 = clk::now();
    simples_ctx_destroy(c);
    const auto t4 = clk::now();
    if (trace) {
        auto ms = [](clk::time_point x, clk::time_point y) { return std::chrono::duration<double, std::milli>(y - x).count(); };
        fprintf(stderr, "[simples_b200] em_impute: create(H2D+prologue) %.1f ms, run(%d it) %.1f ms, fetch(D2H) %.1f ms, destroy %.1f ms\n",
                ms(t0, t1), a->iter, ms(t1, t2), ms(t2, t3), ms(t3, t4));
    }
    return r;
}

// ---------------------------------------------------------------------------------------------
int simples_do_impute(const simples_mi_args* a, simples_mi_out* o) {
    if (!a || !o) return fail(SIMPLES_EINVAL, "null argument");
    CKR(check_common(a->G, a->n, a->M0, a->K0, a->MB));
    if (!a->dat || !a->Y || !a->beta || !a->lambda || !a->sigma || !a->mu || !a->pi || !a->pg)
        return fail(SIMPLES_EINVAL, "dat, Y, beta, lambda, sigma, mu, pi and pg are required");
    if (a->mcmc < 1 || a->burnin < 0) return fail(SIMPLES_EINVAL, "mcmc must be >= 1 and burnin >= 0");
    if (a->want_consensus && a->allreduce) return fail(SIMPLES_EINVAL, "consensus_cluster is single-rank only");
    CKR(ensure_device());

    std::unique_ptr<simples_ctx> guard(new simples_ctx());
    simples_ctx* c = guard.get();
    c->kind = 1;
    if (a->stream) {
        c->stream = (cudaStream_t)a->stream;
    } else {
        CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    c->cutoff = a->cutoff;
    c->seed = a->seed;
    c->compat = a->ref_compat_flags;
    c->allreduce = a->allreduce;
    c->allreduce_user = a->allreduce_user;
    c->mcmc = a->mcmc;
    c->burnin = a->burnin;
    const bool shared = sigma_is_shared(a->sigma, a->G, a->M0);
    CKR(setup_common(c, a->G, a->n, a->M0, a->K0, a->MB, a->n_total, a->cell_offset, a->Y, a->dat, a->pg, a->beta, a->sigma,
                     a->lambda, a->pi, a->celltype, a->cutoff, shared));
    DevState& s = c->s;
    const Dims& d = s.d;
    const int M0 = d.M0, K0 = d.K0, ldG = d.ldG, n = d.n, G = d.G;
    const size_t na = c->n_alloc;
    {
        std::vector<double> hm = to_rowmajor(a->mu, G, M0, ldG, 0.0);
        std::vector<double> pm(ldG, 0.0), ps(ldG, 1.0);
        if (a->pos_mean) std::copy(a->pos_mean, a->pos_mean + G, pm.begin());
        if (a->pos_sd) std::copy(a->pos_sd, a->pos_sd + G, ps.begin());
        CK(cudaMemcpyAsync(s.mu, hm.data(), hm.size() * 8, cudaMemcpyDefault, c->stream));
        CK(cudaMemcpyAsync(s.gmean, pm.data(), size_t(ldG) * 8, cudaMemcpyDefault, c->stream));
        CK(cudaMemcpyAsync(s.gsd, ps.data(), size_t(ldG) * 8, cudaMemcpyDefault, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    launch_center(s.Y, s.gmean, s.gsd, n, ldG, 0, 1, c->stream);   // Y <- (Y - pos_mean) / pos_sd  (:50)
    CKR(c->dalloc(&c->rec1, na * ldG));
    CKR(c->dalloc(&c->rec2, na * ldG));
    CKR(c->dalloc(&c->rEF, size_t(n) * K0));
    CKR(c->dalloc(&c->rF2, size_t(n) * K0));
    CKR(c->dalloc(&c->rVF, size_t(n) * K0));
    const int nit = a->mcmc + a->burnin;
    CKR(c->dalloc(&c->tot_mi, size_t(nit) * (2 * M0 + 1)));
    if (a->want_consensus && o->consensus_cluster) CKR(c->dalloc(&c->cons, size_t(n) * n));
    const double pi_const = (c->compat & SIMPLES_COMPAT_PI) ? 3.14159265359 : M_PI;

    using clk_mi = std::chrono::steady_clock;
    const bool trace_mi = getenv("SIMPLES_TRACE") != nullptr;
    if (trace_mi) cudaStreamSynchronize(c->stream);
    const auto tm0 = clk_mi::now();
    CKR(run_prep(c));   // parameters are fixed: cluster constants computed once, not per sweep
    for (int it = 1; it <= nit; ++it) {                                       // :67
        const bool rec = it > a->burnin;
        c->sweep = unsigned(it - 1);
        CKR(e_pass(c, true, false, true, true, rec, false, pi_const));        // :69-104, :107, :112
        launch_finalize_cell(s.part_cell, s.ncta_cell, M0, c->tot_mi + size_t(it - 1) * (2 * M0 + 1), c->stream);   // :89
        if (rec && c->cons) launch_consensus(s.z, c->cons, n, M0, c->stream);   // :96
        CKR(run_sweep(c, false, rec ? c->rec1 : nullptr, rec ? c->rec2 : nullptr));   // :109-136, :140-141
        if (rec) launch_mi_record_factors(d, s.W, s.z, s.cc.Mm, c->rEF, c->rF2, c->rVF, c->stream);   // :143-147
    }
    CK(cudaGetLastError());
    CKR(allreduce(c, c->tot_mi, size_t(nit) * (2 * M0 + 1)));
    CK(cudaStreamSynchronize(c->stream));
    if (trace_mi)
        fprintf(stderr, "[simples_b200] do_impute: %d sweeps on device %.1f ms (%.2f ms/sweep), nnz0 %lld\n", nit,
                std::chrono::duration<double, std::milli>(clk_mi::now() - tm0).count(),
                std::chrono::duration<double, std::milli>(clk_mi::now() - tm0).count() / nit, c->nnz0);

    // ---- outputs (R/do_impute.R:156-159) ----
    if (o->loglik) {
        std::vector<double> t(size_t(nit) * (2 * M0 + 1));
        CK(cudaMemcpy(t.data(), c->tot_mi, t.size() * 8, cudaMemcpyDefault));
        double sacc = 0.0;
        for (int it = a->burnin; it < nit; ++it) sacc += t[size_t(it) * (2 * M0 + 1) + 2 * M0];
        *o->loglik = sacc / a->mcmc;
    }
    if (o->impt)
        CKR(fetch_big(c, o->impt, [&](double* dst, size_t c0, size_t nc) {
            launch_mi_finalize(s.Y + c0 * ldG, s.mask, c->rec1 + c0 * ldG, c->rec2 + c0 * ldG, s.gmean, s.gsd, dst, nullptr,
                               int(nc), int(c0), c->n_alloc, G, ldG, a->mcmc, c->stream);
        }));
    if (o->impt_var)
        CKR(fetch_big(c, o->impt_var, [&](double* dst, size_t c0, size_t nc) {
            launch_mi_finalize(s.Y + c0 * ldG, s.mask, c->rec1 + c0 * ldG, c->rec2 + c0 * ldG, s.gmean, s.gsd, nullptr, dst,
                               int(nc), int(c0), c->n_alloc, G, ldG, a->mcmc, c->stream);
        }));
    if (o->EF || o->varF) {
        std::vector<double> ef(size_t(n) * K0), f2(size_t(n) * K0), vf(size_t(n) * K0);
        CK(cudaMemcpy(ef.data(), c->rEF, ef.size() * 8, cudaMemcpyDefault));
        CK(cudaMemcpy(f2.data(), c->rF2, f2.size() * 8, cudaMemcpyDefault));
        CK(cudaMemcpy(vf.data(), c->rVF, vf.size() * 8, cudaMemcpyDefault));
        const double mc = a->mcmc;
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < K0; ++k) {
                const size_t e = size_t(i) * K0 + k, h = size_t(k) * n + i;
                const double m1 = ef[e] / mc;
                if (o->EF) o->EF[h] = m1;
                if (o->varF) o->varF[h] = f2[e] / mc - m1 * m1 + vf[e] / mc;
            }
    }
    if (o->consensus_cluster && c->cons) {
        launch_scale(c->cons, size_t(n) * n, 1.0 / a->mcmc, c->stream);
        CK(cudaMemcpyAsync(o->consensus_cluster, c->cons, size_t(n) * n * 8, cudaMemcpyDefault, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    return SIMPLES_OK;
}

// ---------------------------------------------------------------------------------------------
// lq-gene fit between the two EM_impute calls (R/SIMPLE.R:305-411 = R/SIMPLE_B.R:332-419)
int simples_lq_fit(const simples_lq_args* a, simples_lq_out* o) {
    if (!a || !o) return fail(SIMPLES_EINVAL, "null argument");
    CKR(check_common(a->Gq, a->n, a->M0, a->K0, a->MB));
    if (!a->dat || !a->sigma || !a->pg || !a->z || !a->Ef || !a->Varf)
        return fail(SIMPLES_EINVAL, "dat, sigma, pg, z, Ef and Varf are required");
    if (a->resid_mode != 0 && a->resid_mode != 1) return fail(SIMPLES_EINVAL, "resid_mode must be 0 or 1");
    if (a->n_total != 0 && a->n_total < a->n) return fail(SIMPLES_EINVAL, "n_total < n");
    if (a->impute && !o->Y) return fail(SIMPLES_EINVAL, "impute = 1 needs out->Y");
    CKR(ensure_device());

    std::unique_ptr<simples_ctx> guard(new simples_ctx());
    simples_ctx* c = guard.get();
    c->kind = 2;
    if (a->stream) {
        c->stream = (cudaStream_t)a->stream;
    } else {
        CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    c->cutoff = a->cutoff;
    c->seed = a->seed;
    c->sweep = a->sweep;
    c->allreduce = a->allreduce;
    c->allreduce_user = a->allreduce_user;
    const int Gq = a->Gq, n = a->n, M0 = a->M0, K0 = a->K0;
    {
        std::vector<double> b0(size_t(Gq) * K0, 0.0), l1(size_t(M0) * K0, 1.0), p1(M0, 1.0 / M0);
        CKR(setup_common(c, Gq, n, M0, K0, a->MB, a->n_total, a->cell_offset, a->dat, a->dat, a->pg, b0.data(), a->sigma,
                         l1.data(), p1.data(), a->celltype, a->cutoff, true));
    }
    DevState& s = c->s;
    const Dims& d = s.d;
    const int ldG = d.ldG;
    const size_t na = c->n_alloc;
    Tracer tr("lq_fit", c->stream);
    tr.mark("H2D(dat)+mask");
    launch_fill(s.gsd, ldG, 1.0, c->stream);
    CKR(upload_z(c, a->z));
    double* varf = nullptr;
    CKR(c->dalloc(&varf, size_t(M0) * K0 * K0));
    CK(cudaMemcpyAsync(varf, a->Varf, size_t(M0) * K0 * K0 * 8, cudaMemcpyDefault, c->stream));
    {   // Ef: M0 col-major n x K0 blocks -> W [M0][n][K0]
        double* tmp = nullptr;
        CK(cudaMallocAsync(&tmp, size_t(M0) * n * K0 * 8, c->stream));
        CK(cudaMemcpyAsync(tmp, a->Ef, size_t(M0) * n * K0 * 8, cudaMemcpyDefault, c->stream));
        for (int m = 0; m < M0; ++m)
            launch_transpose(tmp + size_t(m) * n * K0, s.W + size_t(m) * n * K0, K0, n, n, K0, c->stream);
        CK(cudaFreeAsync(tmp, c->stream));
    }
    const int ncg = K0 * M0 + 2 * M0;
    CKR(c->dalloc(&s.Vg, na * ncg));
    s.statsg_count = size_t(ldG) * ncg + size_t(ldG) * M0 + size_t(M0) * K0 * K0 + size_t(M0) * K0 + 2 * M0 + 1;
    const size_t nC = size_t(ldG) * ncg, nU2 = size_t(ldG) * M0;
    CKR(c->dalloc(&s.statsg, s.statsg_count));
    c->nsplit_g = 8;
    CKR(c->dalloc(&c->part_g, size_t(c->nsplit_g) * ldG * ncg));
    s.nsplit_ss = std::max(1, (8 * NUM_SMS_B200) / M0);   // warp-level partial slots of k_sstats
    CKR(c->dalloc(&s.part_ss, size_t(s.nsplit_ss) * M0 * (K0 * K0 + K0)));

    tr.mark("H2D(z, Ef)");
    // ---- cell operands + membership draw, then the statistics of the response ----
    {
        LqCellArgs la{d, s.z, s.W, s.Vg, s.Fh, s.im, c->seed, c->sweep, M0 > 1};
        launch_lq_cell(la, c->stream);
        launch_colsum(s.z, c->cellsum_last, n, M0, c->stream);
    }
    const bool own_resp = a->resp != nullptr && a->resp != a->dat;
    double* R = s.Y;            // response matrix on the device, [n][ldG]
    if (own_resp) {
        CKR(c->dalloc(&R, na * ldG));
        if (ldG == Gq) {
            CK(cudaMemcpyAsync(R, a->resp, size_t(n) * Gq * 8, cudaMemcpyDefault, c->stream));
        } else {
            CK(cudaMemcpy2DAsync(R, size_t(ldG) * 8, a->resp, size_t(Gq) * 8, size_t(Gq) * 8, n, cudaMemcpyDefault,
                                 c->stream));
        }
    }
    // T = Y (z_m Ef_m), U = Y z, Uh = Y sqrt z: the DMMA mstep_stats kernel in blocks of <= 96 operand columns;
    // U2 = Y.^2 z through the generic kernel (M0 columns)
    const int NVPq = std::min(96, round_up(ncg, 8)), VSq = NVPq + 4, splitKq = mstats_pick_splitk(ldG);
    double *Vblk, *zs0, *pmsq, *blkq;
    CKR(c->dalloc(&Vblk, na * VSq));
    CKR(c->dalloc(&zs0, na));
    CKR(c->dalloc(&pmsq, size_t(splitKq) * ldG * (NVPq + 1)));
    CKR(c->dalloc(&blkq, size_t(ldG) * NVPq));
    auto gene_stats = [&](const double* Ymat, double* out) {
        for (int c0 = 0; c0 < ncg; c0 += NVPq) {
            launch_extract_cols(s.Vg, Vblk, n, ncg, ncg, c0, NVPq, VSq, c->stream);
            MstatsArgs ma{Ymat, Vblk, zs0, pmsq, n, ldG, NVPq, VSq, splitKq};
            launch_mstats(ma, c->stream);
            launch_reduce_parts(pmsq, blkq, size_t(ldG) * NVPq, splitKq, c->stream);
            launch_scatter_block(blkq, out, ldG, ncg, NVPq, c0, std::min(NVPq, ncg - c0), c->stream);
        }
        launch_ygemm_simple(Ymat, s.Vg + size_t(K0) * M0, c->part_g, n, ldG, M0, ncg, 1, c->nsplit_g, c->stream);
        launch_reduce_parts(c->part_g, out + nC, nU2, c->nsplit_g, c->stream);
    };
    gene_stats(R, s.statsg);
    {
        SstatsArgs sa{d, s.W, s.z, s.part_ss, s.nsplit_ss};
        launch_sstats(sa, c->stream);
        FinalizeArgs fa{d, nullptr, c->cellsum_last, s.part_ss, 0, 1, s.nsplit_ss, s.statsg, nC + nU2, 1};
        launch_finalize_stats(fa, c->stream);
    }
    CKR(allreduce(c, s.statsg, s.statsg_count));
    double* stats_d = s.statsg;
    if (own_resp && a->resid_mode) {           // residual of dat, response init_imp (R/SIMPLE.R:360-364)
        CKR(c->dalloc(&stats_d, nC + nU2 + 1));
        gene_stats(s.Y, stats_d);
        CKR(allreduce(c, stats_d, nC + nU2));
    }
    {
        LqSolveArgs qa{d, s.statsg, stats_d, varf, s.beta, s.sigma, s.mu, a->penl, a->resid_mode};
        launch_lq_solve(qa, c->stream);
    }
    CK(cudaGetLastError());
    tr.mark("statistics+regression");
    if (a->impute) {                            // R/SIMPLE.R:376-411
        PrepArgs pa{d, s.beta, s.sigma, s.mu, s.lam, s.pi, s.gmean, s.gsd, s.Q, s.isg, s.Qs, s.sdsA, s.isdA, s.musdA, s.igs, s.cc, 0};
        launch_prep(pa, c->stream, c->stream);
        CKR(run_sweep(c, false, nullptr, nullptr));
    }
    tr.mark("draw");
    if (o->mu) CKR(fetch_genemajor(c, s.mu, M0, o->mu));
    if (o->beta) CKR(fetch_genemajor(c, s.beta, K0, o->beta));
    if (o->sigma) CKR(fetch_genemajor(c, s.sigma, M0, o->sigma));
    if (a->impute && o->Y) {
        if (ldG == Gq) {
            CK(cudaMemcpyAsync(o->Y, s.Y, size_t(n) * Gq * 8, cudaMemcpyDefault, c->stream));
        } else {
            CK(cudaMemcpy2DAsync(o->Y, size_t(Gq) * 8, s.Y, size_t(ldG) * 8, size_t(Gq) * 8, n, cudaMemcpyDefault,
                                 c->stream));
        }
    }
    tr.mark("D2H");
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    return SIMPLES_OK;
}

// ---------------------------------------------------------------------------------------------
// init_impute / init_impute_bulk (R/SIMPLE.R:15-73, R/SIMPLE_B.R:23-114)
int simples_init_impute(const simples_init_args* a, simples_init_out* o) {
    if (!a || !o) return fail(SIMPLES_EINVAL, "null argument");
    if (a->G < 1 || a->n < 1) return fail(SIMPLES_EINVAL, "G and n must be >= 1");
    if (a->M0 < 1 || a->M0 > MAXM) return fail(SIMPLES_EINVAL, "M0 must be in [1, 32]");
    if (!a->Y2 || !a->clus || !o->impute) return fail(SIMPLES_EINVAL, "Y2, clus and out->impute are required");
    if (a->n_total != 0 && a->n_total < a->n) return fail(SIMPLES_EINVAL, "n_total < n");
    const int G = a->G, n = a->n, M0 = a->M0;
    std::vector<int32_t> c0(n);
    std::vector<double> cnt(M0, 0.0);
    for (int i = 0; i < n; ++i) {
        if (a->clus[i] < 1 || a->clus[i] > M0) return fail(SIMPLES_EINVAL, "clus must be in 1..M0");
        c0[i] = a->clus[i] - 1;
        cnt[c0[i]] += 1.0;
    }
    CKR(ensure_device());
    std::unique_ptr<simples_ctx> guard(new simples_ctx());
    simples_ctx* c = guard.get();
    c->kind = 3;
    if (a->stream) {
        c->stream = (cudaStream_t)a->stream;
    } else {
        CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    c->allreduce = a->allreduce;
    c->allreduce_user = a->allreduce_user;
    const size_t GM = size_t(G) * M0, nst = 5 * GM;
    double *Y2 = nullptr, *part = nullptr, *stats = nullptr, *pg1 = nullptr, *fit = nullptr;
    int32_t* clus0 = nullptr;
    const int nsplit = init_stats_splits(n);
    CKR(c->dalloc(&Y2, size_t(n) * G, false));
    CKR(c->dalloc(&clus0, size_t(n), false));
    CKR(c->dalloc(&part, size_t(nsplit) * nst));
    CKR(c->dalloc(&stats, nst + M0));
    CKR(c->dalloc(&fit, 4 * GM));
    CK(cudaMemcpyAsync(Y2, a->Y2, size_t(n) * G * 8, cudaMemcpyDefault, c->stream));
    CK(cudaMemcpyAsync(clus0, c0.data(), size_t(n) * 4, cudaMemcpyDefault, c->stream));
    CK(cudaMemcpyAsync(stats + nst, cnt.data(), size_t(M0) * 8, cudaMemcpyDefault, c->stream));
    if (a->pg1) {
        CKR(c->dalloc(&pg1, GM, false));
        CK(cudaMemcpyAsync(pg1, a->pg1, GM * 8, cudaMemcpyDefault, c->stream));
    }
    Tracer tr("init_impute", c->stream);
    tr.mark("H2D");
    launch_init_stats(Y2, clus0, n, G, M0, a->cutoff, part, nsplit, c->stream);
    launch_reduce_parts(part, stats, nst, nsplit, c->stream);
    CKR(allreduce(c, stats, nst + M0));
    InitFitArgs fa{stats, G, M0, a->cutoff, a->p_min, 0.1, 4.0, pg1, fit, fit + GM, fit + 2 * GM, fit + 3 * GM};
    tr.mark("moments");
    launch_ztrunc(fa, c->stream);
    tr.mark("fit(Brent)");
    InitDrawArgs da{Y2, Y2, clus0, n, G, fit, fit + GM, fit + 2 * GM, pg1, a->cutoff, a->seed, a->sweep, a->cell_offset};
    launch_init_draw(da, c->stream);
    CK(cudaGetLastError());
    tr.mark("draw");
    CK(cudaMemcpyAsync(o->impute, Y2, size_t(n) * G * 8, cudaMemcpyDefault, c->stream));
    if (o->pg) CK(cudaMemcpyAsync(o->pg, fit + 3 * GM, GM * 8, cudaMemcpyDefault, c->stream));
    if (o->mu) CK(cudaMemcpyAsync(o->mu, fit, GM * 8, cudaMemcpyDefault, c->stream));
    if (o->sd) CK(cudaMemcpyAsync(o->sd, fit + GM, GM * 8, cudaMemcpyDefault, c->stream));
    tr.mark("D2H");
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    return SIMPLES_OK;
}

// ---------------------------------------------------------------------------------------------
// initial clustering (R/SIMPLE.R:258-266, R/SIMPLE_B.R:293-302)
int simples_init_cluster(const simples_clus_args* a, simples_clus_out* o) {
    if (!a || !o) return fail(SIMPLES_EINVAL, "null argument");
    if (a->G < 2 || a->n < 2) return fail(SIMPLES_EINVAL, "G and n must be >= 2");
    if (a->M0 < 1 || a->M0 > MAXM) return fail(SIMPLES_EINVAL, "M0 must be in [1, 32]");
    if (a->K < 1 || a->K > 32 || a->K >= a->G) return fail(SIMPLES_EINVAL, "K must be in [1, 32] and < G");
    if (!a->X || !o->clus) return fail(SIMPLES_EINVAL, "X and out->clus are required");
    if (a->nstart < 1 || a->iter_max < 1) return fail(SIMPLES_EINVAL, "nstart and iter_max must be >= 1");
    if (a->n_total != 0 && a->n_total < a->n) return fail(SIMPLES_EINVAL, "n_total < n");
    CKR(ensure_device());
    std::unique_ptr<simples_ctx> guard(new simples_ctx());
    simples_ctx* c = guard.get();
    c->kind = 4;
    if (a->stream) {
        c->stream = (cudaStream_t)a->stream;
    } else {
        CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    c->allreduce = a->allreduce;
    c->allreduce_user = a->allreduce_user;
    cudaStream_t st = c->stream;
    const int G = a->G, n = a->n, K = a->K, M0 = a->M0, R = a->nstart;
    const int ldG = round_up(G, GPAD), na = round_up(n, 128);
    const long long n_total = a->n_total > 0 ? a->n_total : n;
    const int KP = std::min(32, std::min(K + 10, G));          // subspace width (oversampled)
    const int NQP = round_up(K, 8), QS = NQP + 4;
    const int NVP = std::min(96, round_up(G, 8)), VS = NVP + 4;

    double *Y, *gmean, *gsd, *ones, *part, *V, *zs, *pms, *blk, *C, *Q, *Z, *A, *dv, *isg, *P, *S2;
    CKR(c->dalloc(&Y, size_t(na) * ldG));
    CUtensorMap tmapY;
    CK(make_tmap_2d(&tmapY, Y, ldG, na, 16, 128));
    CKR(c->dalloc(&gmean, size_t(ldG)));
    CKR(c->dalloc(&gsd, size_t(ldG)));
    CKR(c->dalloc(&ones, size_t(na)));
    const int nsplit = 32;
    CKR(c->dalloc(&part, size_t(nsplit) * ldG));
    if (ldG == G) {
        CK(cudaMemcpyAsync(Y, a->X, size_t(n) * G * 8, cudaMemcpyDefault, st));
    } else {
        CK(cudaMemcpy2DAsync(Y, size_t(ldG) * 8, a->X, size_t(G) * 8, size_t(G) * 8, n, cudaMemcpyDefault, st));
    }
    Tracer tr("init_cluster", st);
    tr.mark("H2D");
    // ---- t(scale(t(X))): centre by the row mean, divide by the row sd (n - 1) ----
    launch_rowsum(Y, part, n, ldG, nsplit, st);
    launch_reduce_parts(part, gmean, ldG, nsplit, st);
    CKR(allreduce(c, gmean, ldG));
    launch_scale(gmean, ldG, 1.0 / double(n_total), st);
    launch_fill(gsd, ldG, 1.0, st);
    launch_center(Y, gmean, gsd, n, ldG, 0, 0, st);
    launch_fill(ones, size_t(n), 1.0, st);
    launch_ygemm_simple(Y, ones, part, n, ldG, 1, 1, 1, nsplit, st);
    launch_reduce_parts(part, gmean, ldG, nsplit, st);                      // gmean <- sum of squares (reused buffer)
    CKR(allreduce(c, gmean, ldG));
    launch_sd_from_ss(gmean, gsd, gmean, G, ldG, double(n_total - 1), st);  // gsd <- sd, gmean <- 0
    launch_center(Y, gmean, gsd, n, ldG, 0, 0, st);
    tr.mark("scale");
    // ---- Gram matrix C = Ys Ys' in column blocks through the DMMA mstep_stats kernel ----
    const int splitK = mstats_pick_splitk(ldG);
    CKR(c->dalloc(&V, size_t(na) * VS));
    CKR(c->dalloc(&zs, size_t(na)));
    CKR(c->dalloc(&pms, size_t(splitK) * ldG * (NVP + 1)));
    CKR(c->dalloc(&blk, size_t(ldG) * NVP));
    CKR(c->dalloc(&C, size_t(ldG) * ldG));
    for (int g0 = 0; g0 < G; g0 += NVP) {
        const int ncol = std::min(NVP, ldG - g0);
        launch_extract_cols(Y, V, n, ldG, G, g0, NVP, VS, st);
        MstatsArgs ma{Y, V, zs, pms, n, ldG, NVP, VS, splitK};
        launch_mstats(ma, st);
        launch_reduce_parts(pms, blk, size_t(ldG) * NVP, splitK, st);
        launch_scatter_block(blk, C, ldG, ldG, NVP, g0, ncol, st);
    }
    CKR(allreduce(c, C, size_t(ldG) * ldG));
    tr.mark("gram");
    // ---- leading K eigenvectors of C (= left singular vectors of Ys), d = sqrt(eigenvalues) ----
    CKR(c->dalloc(&Q, size_t(ldG) * KP));
    CKR(c->dalloc(&Z, size_t(ldG) * KP));
    CKR(c->dalloc(&A, size_t(ldG) * QS));
    CKR(c->dalloc(&dv, size_t(32)));
    launch_subspace(C, Q, Z, A, dv, G, ldG, KP, K, QS, 150, a->seed, st);
    tr.mark("subspace");
    // ---- cellmat = Ys' U = V diag(d)  (estep_project) ----
    CKR(c->dalloc(&isg, size_t(ldG)));
    CKR(c->dalloc(&P, size_t(na) * NQP));
    CKR(c->dalloc(&S2, size_t(na)));
    {
        ProjArgs pa{Y, A, isg, P, S2, n, ldG, NQP, QS};
        launch_proj(pa, tmapY, st);
    }
    CK(cudaGetLastError());
    tr.mark("cellmat");
    // ---- k-means: nstart restarts of Lloyd's algorithm advanced together ----
    KmArgs km{};
    km.X = P;
    km.ldx = NQP;
    km.n = n;
    km.K = K;
    km.M0 = M0;
    km.R = R;
    km.nsplit = std::max(1, std::min((n + 127) / 128, 64));
    const int nw = km.nsplit * 4;
    const size_t nsum = size_t(R) * M0 * (K + 1);
    double* sums;
    CKR(c->dalloc(&km.cent, size_t(R) * M0 * K));
    CKR(c->dalloc(&km.part, size_t(nw) * nsum));
    CKR(c->dalloc(&km.wss_part, size_t(nw) * R));
    CKR(c->dalloc(&sums, nsum + R));
    int* nlive;                                    // [iter_max] restarts still moving after each update
    CKR(c->dalloc(&km.moved, size_t(R), false));
    CKR(c->dalloc(&nlive, size_t(a->iter_max)));
    {
        std::vector<int> ones_i(R, 1);
        CK(cudaMemcpyAsync(km.moved, ones_i.data(), size_t(R) * 4, cudaMemcpyDefault, st));
        CK(cudaStreamSynchronize(st));
    }
    launch_km_init(km, a->seed, n_total, a->cell_offset, st);
    CKR(allreduce(c, km.cent, size_t(R) * M0 * K));
    int km_iters = 0;
    for (int it = 0; it < a->iter_max; ++it) {
        ++km_iters;
        launch_km_pass(km, 0, 0, nullptr, st);     // converged restarts are skipped (their sums stay valid)
        launch_reduce_parts(km.part, sums, nsum, nw, st);
        CKR(allreduce(c, sums, nsum));
        launch_km_update(km, sums, nlive + it, st);
        if ((it & 3) == 3) {                       // every 4 iterations: stop once no restart moves (Lloyd fixed point)
            int h = 1;
            CK(cudaMemcpyAsync(&h, nlive + it, 4, cudaMemcpyDefault, st));
            CK(cudaStreamSynchronize(st));
            if (h == 0) break;
        }
    }
    launch_km_pass(km, 1, 0, nullptr, st);
    launch_reduce_parts(km.wss_part, sums + nsum, size_t(R), nw, st);
    CKR(allreduce(c, sums + nsum, size_t(R)));
    std::vector<double> wss(R);
    CK(cudaMemcpyAsync(wss.data(), sums + nsum, size_t(R) * 8, cudaMemcpyDefault, st));
    CK(cudaStreamSynchronize(st));
    int best = 0;
    for (int r = 1; r < R; ++r)
        if (wss[r] < wss[best]) best = r;
    tr.mark("kmeans");
    if (tr.on) {
        std::vector<int> nl(a->iter_max, 0);
        cudaMemcpy(nl.data(), nlive, size_t(a->iter_max) * 4, cudaMemcpyDefault);
        std::string t = " [" + std::to_string(km_iters) + " Lloyd iterations; restarts still moving:";
        for (int it = 0; it < km_iters; it += 4) t += " " + std::to_string(nl[it]);
        tr.msg += t + "],";
    }
    int32_t* labels;
    CKR(c->dalloc(&labels, size_t(na)));
    launch_km_pass(km, 2, best, labels, st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(o->clus, labels, size_t(n) * 4, cudaMemcpyDefault, st));
    if (o->tot_withinss) *o->tot_withinss = wss[best];
    if (o->d) CK(cudaMemcpyAsync(o->d, dv, size_t(K) * 8, cudaMemcpyDefault, st));
    if (o->cellmat) {
        double* tmp;
        CKR(c->dalloc(&tmp, size_t(n) * K, false));
        launch_cellmat_out(P, tmp, n, K, NQP, st);
        CK(cudaMemcpyAsync(o->cellmat, tmp, size_t(n) * K * 8, cudaMemcpyDefault, st));
    }
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    return SIMPLES_OK;
}

}  // extern "C"
