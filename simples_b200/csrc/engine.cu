// Host orchestration + C ABI (include/simples_b200.h) of the B200 EM imputation hot path.
// One context = one EM_impute (R/EM_impute.R:75-339) or do_impute (R/do_impute.R:29-160) call on
// this process's shard of cells.  The EM loop is enqueued on one CUDA stream without any
// host <-> device synchronisation: all data-dependent decisions (lambda eligibility, lasso active
// sets, priorB bookkeeping) happen on the device.
#include <immintrin.h>

#include <mutex>
#include <thread>

#include "engine_internal.h"

namespace sbe {

thread_local std::string g_err;
int g_device = -1;
thread_local int t_device = -1;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

const char* kKernelNames[KC_COUNT] = {"estep_project", "cell_post",  "impute_sweep",     "mstep_stats", "sstats", "reduce",
                                      "cluster_prep",  "gene_solve", "lambda_mu_pi",     "misc",        "allreduce"};

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
        if (g_ndev == 0) {
            g_ndev = 1;
            g_devs[0] = dev;
        }
    }
    const int dev = t_device >= 0 ? t_device : g_device;
    static std::mutex mu;
    static int checked[64] = {0};          // 1 = sm_100 verified and pool configured
    CK(cudaSetDevice(dev));
    std::lock_guard<std::mutex> lk(mu);
    if (dev < 64 && checked[dev]) return SIMPLES_OK;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        char buf[256];
        snprintf(buf, sizeof buf, "device %d (%s) is sm_%d%d; this library is built for sm_100a (B200) only", dev, prop.name,
                 prop.major, prop.minor);
        return fail(SIMPLES_ENODEV, buf);
    }
    {   // keep freed blocks cached in the default pool between calls
        cudaMemPool_t pool;
        CK(cudaDeviceGetDefaultMemPool(&pool, dev));
        unsigned long long thr = ~0ull;
        CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    }
    if (dev < 64) checked[dev] = 1;
    return SIMPLES_OK;
}

// A per-rank condition (e.g. non-finite input in one shard) must fail on EVERY rank, or the others would wait in the
// next collective forever: sum the count over the ranks first.
int agree_all_ranks(simples_ctx* c, long long* count) {
    if (!c->allreduce) return SIMPLES_OK;
    double* d = nullptr;
    CKR(c->dalloc(&d, 1));
    launch_fill(d, 1, double(*count), c->stream);
    CKR(allreduce(c, d, 1));
    double h = 0;
    CK(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDefault, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *count = (long long)h;
    return SIMPLES_OK;
}

int allreduce(simples_ctx* c, double* buf, size_t count) {
    if (!c->allreduce) return SIMPLES_OK;
    P p(c, KC_COMM, 0);
    if (c->allreduce(c->allreduce_user, buf, count, (void*)c->stream) != 0)
        return fail(SIMPLES_ECOMM, "all-reduce callback failed");
    return SIMPLES_OK;
}


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
    d.NQP = std::max(16, round_up(d.NQ, 8));
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

// 0: device / managed memory, 1: pinned host memory, 2: pageable host memory
int host_pointer_kind(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return 2;
    }
    return at.type == cudaMemoryTypeUnregistered ? 2 : (at.type == cudaMemoryTypeHost ? 1 : 0);
}

int host_pack_threads() {
    const int hw = int(std::max(1u, std::thread::hardware_concurrency()));
    int T = std::min(16, std::max(1, hw / std::max(1, std::max(g_ndev, g_rank_nranks))));
    if (const char* e = getenv("SIMPLES_HOST_THREADS")) T = std::max(1, atoi(e));
    return T;
}

// Packing on the host pays when it beats sending the doubles: measured ~3 GB/s per thread against ~50 GB/s of PCIe from
// pinned memory (and the packing overlaps the upload of Y), ~12 GB/s from pageable memory (R's own matrices).
bool pack_mask_on_host(const void* Y0) {
    if (getenv("SIMPLES_DEVICE_MASK")) return false;
    const int kind = host_pointer_kind(Y0), T = host_pack_threads();
    return kind == 2 ? T >= 2 : (kind == 1 ? T >= 8 : false);
}

// mask[w][i] bit j <=> Y0[32 w + j, i] <= cutoff, for the n cells of a column-major G x n host matrix (the device layout:
// chunk-major words, row stride na).  Cell ranges are packed by host threads (SIMPLES_HOST_THREADS, default: the cores
// divided by the ranks of the job, at most 16) with SSE2 / AVX2 compares; returns the number of set bits.  Pure
// marshalling: the imputation itself has no CPU path.
static inline uint32_t pack32_sse2(const double* y, double cutoff) {
    const __m128d c = _mm_set1_pd(cutoff);
    uint32_t m = 0;
    for (int j = 0; j < 32; j += 2) m |= uint32_t(_mm_movemask_pd(_mm_cmple_pd(_mm_loadu_pd(y + j), c))) << j;
    return m;
}
__attribute__((target("avx2"))) static uint32_t pack32_avx2(const double* y, double cutoff) {
    const __m256d c = _mm256_set1_pd(cutoff);
    uint32_t m = 0;
    for (int j = 0; j < 32; j += 4) m |= uint32_t(_mm256_movemask_pd(_mm256_cmp_pd(_mm256_loadu_pd(y + j), c, _CMP_LE_OQ))) << j;
    return m;
}

unsigned long long host_pack_mask(const double* Y0, int G, int n, int nwords, size_t na, double cutoff, uint32_t* mask) {
    int T = host_pack_threads();
    T = std::min(T, std::max(1, n / 256));
    const bool avx2 = __builtin_cpu_supports("avx2");
    std::vector<unsigned long long> cnt(T, 0);
    auto work = [&](int t) {
        // whole blocks of 16 cells per thread: the 16 words of one mask row are written as one 64-byte line (a word-by-
        // word walk down a cell scatters 63 stores over rows that are 4 na bytes apart)
        const size_t nb = (size_t(n) + 15) / 16;
        const size_t i0 = std::min(size_t(n), (nb * t / T) * 16), i1 = std::min(size_t(n), (nb * (t + 1) / T) * 16);
        unsigned long long c = 0;
        for (size_t ib = i0; ib < i1; ib += 16) {
            const int bc = int(std::min<size_t>(16, i1 - ib));
            for (int w = 0; w < nwords; ++w) {
                const int g0 = w * 32, len = std::min(32, G - g0);
                uint32_t buf[16];
                for (int ii = 0; ii < bc; ++ii) {
                    const double* y = Y0 + (ib + ii) * size_t(G) + g0;
                    uint32_t m = 0;
                    if (len == 32) {
                        m = avx2 ? pack32_avx2(y, cutoff) : pack32_sse2(y, cutoff);
                    } else {
                        for (int j = 0; j < len; ++j) m |= uint32_t(y[j] <= cutoff) << j;
                    }
                    buf[ii] = m;
                    c += __builtin_popcount(m);
                }
                memcpy(mask + size_t(w) * na + ib, buf, size_t(bc) * 4);
            }
        }
        cnt[t] = c;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    unsigned long long tot = 0;
    for (auto v : cnt) tot += v;
    return tot;
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
    CK(make_tmap_2d(&c->tmapY8, s.Y, ldG, c->n_alloc, 16, 8));
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
    CKR(c->dalloc(&s.gmig, size_t(ldG) * 2));
    CKR(c->dalloc(&s.BM, size_t(ldG) * d.QS));
    CKR(c->dalloc(&s.gv, size_t(ldG) * 4));
    CKR(c->dalloc(&s.mc_counter, 1));
    CKR(c->dalloc(&s.part_prep, prep_scratch_doubles(d, c->NS_max)));
    {   // SIMPLES_FUSED_MC=1 selects the fused sweep + projection kernel (k_mcpass.cuh).  It is parity-green but measured
        // slower than the separate kernels at C3 (2.6 ms vs 1.8 + 0.36 ms per MC pass, profiles/r02_notes.md), so it is opt-in.
        const char* e = getenv("SIMPLES_FUSED_MC");
        c->fused_mc = e && atoi(e) != 0;
    }
    s.n_alloc = c->n_alloc;
    CKR(c->dalloc(&s.P, size_t(c->NS_max) * na * d.NQP));
    CKR(c->dalloc(&s.S2, size_t(c->NS_max) * na));
    CKR(c->dalloc(&s.proj_part, size_t(PROJ_MAX_SPLIT) * NUM_SMS_B200 * 128 * (d.NQP + 1)));
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
    CKR(c->dalloc(&s.cc.Vj, size_t(M0) * K0 * K0));
    CKR(c->dalloc(&s.cc.jst, size_t(M0)));
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
    // The observed matrix is needed only as the bit mask Y0 <= cutoff (26 MB at C3 against 1.6 GB of doubles):
    //   * Y0 aliases Y            -> mask from the copy of Y that is already on the device;
    //   * Y0 in HOST memory       -> packed by host threads while the DMA engine uploads Y, only the mask crosses PCIe
    //                                (halves the H2D volume of a call; with pageable R matrices it also avoids a staged copy);
    //   * Y0 in device memory     -> built chunk-wise by a kernel.
    c->host_mask_ms = 0.0;
    if (Y0 == Y) {
        for (size_t c0 = 0; c0 < size_t(n); c0 += cells_per_chunk) {
            const size_t nc = std::min(cells_per_chunk, size_t(n) - c0);
            if (ldG == G) {
                launch_build_mask(s.Y + c0 * ldG, s.mask, int(nc), int(c0), c->n_alloc, G, ldG, cutoff, c->d_nnz, c->stream);
            } else {      // the kernel reads an unpadded [cells][G] chunk
                CK(cudaMemcpy2DAsync(c->stage, size_t(G) * 8, s.Y + c0 * ldG, size_t(ldG) * 8, size_t(G) * 8, nc,
                                     cudaMemcpyDeviceToDevice, c->stream));
                launch_build_mask(c->stage, s.mask, int(nc), int(c0), c->n_alloc, G, ldG, cutoff, c->d_nnz, c->stream);
            }
        }
    } else if (pack_mask_on_host(Y0)) {
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<uint32_t> hm(size_t(nwords) * na, 0u);
        const unsigned long long cnt = host_pack_mask(Y0, G, n, nwords, na, cutoff, hm.data());
        c->host_mask_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        CK(cudaMemcpyAsync(s.mask, hm.data(), hm.size() * 4, cudaMemcpyDefault, c->stream));
        CK(cudaMemcpyAsync(c->d_nnz, &cnt, 8, cudaMemcpyDefault, c->stream));
        CK(cudaStreamSynchronize(c->stream));   // hm goes out of scope
    } else {
        for (size_t c0 = 0; c0 < size_t(n); c0 += cells_per_chunk) {
            const size_t nc = std::min(cells_per_chunk, size_t(n) - c0);
            CK(cudaMemcpyAsync(c->stage, Y0 + c0 * G, nc * G * 8, cudaMemcpyDefault, c->stream));
            launch_build_mask(c->stage, s.mask, int(nc), int(c0), c->n_alloc, G, ldG, cutoff, c->d_nnz, c->stream);
        }
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
        long long bad_all = (long long)h_bad;
        CKR(agree_all_ranks(c, &bad_all));
        if (bad_all) return fail(SIMPLES_EINVAL, "Y contains " + std::to_string(bad_all) + " non-finite value(s)");
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
    const size_t ldh = c->io_ld ? size_t(c->io_ld) : size_t(n);      // column stride of the host matrix
    CK(cudaMemcpy2DAsync(tmp, size_t(n) * 8, z, ldh * 8, size_t(n) * 8, M0, cudaMemcpyDefault, c->stream));
    launch_transpose(tmp, s.z, M0, n, n, M0, c->stream);
    CK(cudaFreeAsync(tmp, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SIMPLES_OK;
}

// one E-pass: projections + cell epilogue
int e_pass(simples_ctx* c, bool update_z, bool stale_q, bool sample, bool draw_membership, bool write_W, bool write_Vg,
           double pi_const, bool have_P) {
    DevState& s = c->s;
    const Dims& d = s.d;
    for (int sc = 0; sc < d.NS && !have_P; ++sc) {
        P p(c, KC_PROJ, proj_launches(d.n, d.ldG, s.proj_part != nullptr));
        ProjArgs pa{s.Y, s.Q + size_t(sc) * d.ldG * d.QS, s.isg + size_t(sc) * d.ldG, s.P + size_t(sc) * d.n * d.NQP,
                    s.S2 + size_t(sc) * d.n, d.n, d.ldG, d.NQP, d.QS, s.proj_part};
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
        ca.expect = c->kind == 0 && (c->compat & SIMPLES_MODE_EXPECT) != 0;
        launch_cell(ca, c->stream);
    }
    CK(cudaGetLastError());
    return SIMPLES_OK;
}

int run_prep(simples_ctx* c) {
    DevState& s = c->s;
    P p(c, KC_PREP, 2);
    PrepArgs pa{s.d, s.beta, s.sigma, s.mu, s.lam, s.pi, s.gmean, s.gsd, s.Q, s.isg, s.Qs, s.sdsA, s.isdA, s.musdA, s.gmig, s.cc, 0,
                s.BM, s.gv, s.part_prep};
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
    sa.gmig = s.gmig;
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
    sa.mi_gbase = c->mi_gbase;
    sa.mi_off = c->mi_off;
    sa.expect = c->kind == 0 && (c->compat & SIMPLES_MODE_EXPECT) != 0;
    launch_sweep(sa, c->stream);
    CK(cudaGetLastError());
    return SIMPLES_OK;
}

// fused MC pass: the sweep and the projection of the E-pass that follows it (cluster-shared sigma)
bool use_fused_mc(const simples_ctx* c) {
    return c->fused_mc && c->s.d.NS == 1 && mcpass_supported(c->s.d) && !(c->compat & SIMPLES_MODE_EXPECT);
}

int run_mcpass(simples_ctx* c, bool clip, double* rec1, double* rec2) {
    DevState& s = c->s;
    P p(c, KC_SWEEP);
    McArgs ma{};
    ma.d = s.d;
    ma.mask = s.mask;
    ma.n_alloc = s.n_alloc;
    ma.Fh = s.Fh;
    ma.im = s.im;
    ma.celltype0 = s.celltype0;
    ma.BM = s.BM;
    ma.gv = s.gv;
    ma.gmean = s.gmean;
    ma.gsd = s.gsd;
    ma.isg = s.isg;
    ma.pg = s.pg;
    ma.cutoff = c->cutoff;
    ma.lower = c->lower;
    ma.upper = c->upper;
    ma.clip = clip;
    ma.seed = c->seed;
    ma.sweep = c->sweep;
    ma.P = s.P;
    ma.S2 = s.S2;
    ma.rec1 = rec1;
    ma.rec2 = rec2;
    ma.counter = s.mc_counter;
    launch_mcpass(ma, c->tmapY8, c->stream);
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
    CKR(e_pass(c, update_z, false, sampling, draw_mem, !sampling, general && !sampling, pi_const, false));
    {
        P p(c, KC_REDUCE);
        launch_finalize_cell(s.part_cell, s.ncta_cell, d.M0, c->cellsum_main, c->stream);
    }
    if (sampling) {
        const bool fused = use_fused_mc(c);
        for (int mc = 0; mc < c->num_mc; ++mc) {
            if (fused) CKR(run_mcpass(c, true, nullptr, nullptr));
            else CKR(run_sweep(c, true, nullptr, nullptr));
            c->sweep += 1;
            const bool last = (mc == c->num_mc - 1);
            CKR(e_pass(c, update_z, (c->compat & SIMPLES_COMPAT_STALE_DT) != 0, !last, draw_mem, last, general && last,
                       pi_const, fused));
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
        P p(c, KC_GENE, 4);
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
        ma.part_mu = s.part_mu;
        ma.penl = c->penl;
        ma.sigma0 = c->sigma0;
        ma.pi_alpha = c->pi_alpha;
        ma.do_lambda = c->max_lambda && it >= c->est_lam;
        ma.compat_priorb_prev = (c->compat & SIMPLES_COMPAT_PRIORB_PREV) != 0;
        ma.debug = getenv("SIMPLES_TRACE") != nullptr;
        launch_mstep(ma, c->stream);
    }
    CK(cudaGetLastError());
    d.NS = 1;   // sigma is cluster-shared after every M-step (R/EM_impute.R:269)
    c->it = it;
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
int fetch_transposed(simples_ctx* c, const double* dev, int rows, int cols, int ld, double* host, long long ld_host) {
    double* tmp = nullptr;
    CK(cudaMallocAsync(&tmp, size_t(rows) * cols * 8, c->stream));
    launch_transpose(dev, tmp, rows, cols, ld, rows, c->stream);
    const size_t ldh = ld_host ? size_t(ld_host) : size_t(rows);
    CK(cudaMemcpy2DAsync(host, ldh * 8, tmp, size_t(rows) * 8, size_t(rows) * 8, cols, cudaMemcpyDefault, c->stream));
    CK(cudaFreeAsync(tmp, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SIMPLES_OK;
}

}  // namespace sbe

// =====================================================================================================
extern "C" {

int simples_abi_version(void) { return SIMPLES_ABI_VERSION; }

const char* simples_last_error(void) { return g_err.c_str(); }

int simples_init(int ndev, const int* devs) {
    if (ndev < 1 || ndev > SIMPLES_MAX_DEVICES || !devs)
        return fail(SIMPLES_EINVAL, "simples_init: ndev must be in 1..SIMPLES_MAX_DEVICES");
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt == 0) return fail(SIMPLES_ENODEV, "simples_init: no CUDA device available (no CPU fallback)");
    for (int i = 0; i < ndev; ++i) {
        if (devs[i] < 0 || devs[i] >= cnt) return fail(SIMPLES_EINVAL, "simples_init: device index out of range");
        for (int j = 0; j < i; ++j)
            if (devs[j] == devs[i]) return fail(SIMPLES_EINVAL, "simples_init: duplicate device");
    }
    comm_destroy_all();
    g_device = devs[0];
    g_ndev = ndev;
    for (int i = 0; i < ndev; ++i) g_devs[i] = devs[i];
    for (int i = 0; i < ndev; ++i) {       // every device must be an sm_100 part
        t_device = devs[i];
        const int r = ensure_device();
        t_device = -1;
        if (r != SIMPLES_OK) {
            g_ndev = 1;
            return r;
        }
    }
    if (ndev > 1) {
        const int r = comm_init_all(ndev, devs);
        if (r != SIMPLES_OK) {
            g_ndev = 1;
            return r;
        }
    }
    return ensure_device();
}

void simples_shutdown(void) {
    comm_destroy_all();
    g_device = -1;
    g_ndev = 0;
}

// Host-side marshalling helper (no GPU involved): the bit mask the library derives from a HOST matrix of observed values.
int64_t simples_pack_mask(const double* Y0, int32_t G, int32_t n, double cutoff, uint32_t* mask) {
    if (!Y0 || !mask || G < 1 || n < 1) return fail(SIMPLES_EINVAL, "simples_pack_mask: bad arguments");
    return (int64_t)host_pack_mask(Y0, G, n, (G + 31) / 32, size_t(n), cutoff, mask);
}

int simples_kernel_count(void) { return KC_COUNT; }
const char* simples_kernel_name(int k) { return (k >= 0 && k < KC_COUNT) ? kKernelNames[k] : ""; }

void simples_ctx_destroy(simples_ctx* ctx) {
    if (!ctx) return;
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    delete ctx;
}

}   // extern "C"

namespace sbe {
// allreduce of a context: the caller's callback, else the library's rank communicator when the cells are sharded
void pick_allreduce(simples_ctx* c, simples_allreduce_fn fn, void* user, long long n, long long n_total) {
    c->allreduce = fn;
    c->allreduce_user = user;
    if (!fn && n_total > n && g_rank_comm) {
        c->allreduce = nccl_allreduce_cb;
        c->allreduce_user = g_rank_comm;
    }
}

int em_create_impl(const simples_em_args* a, simples_ctx** out, long long io_ld) {
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
    pick_allreduce(c, a->allreduce, a->allreduce_user, a->n, a->n_total);
    c->io_ld = io_ld;
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
    CKR(c->dalloc(&s.part_mu, size_t(M0) * mu_gene_splits(ldG) * (K0 * K0 + K0)));
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
            launch_scale(s.mu, size_t(ldG) * M0, 1.0 / double(d.n_total), c->stream);      // Q3: / n
        } else {                                                                           // the intended cluster means: / nz_m
            double* nz = nullptr;
            CKR(c->dalloc(&nz, size_t(2 * M0 + 1)));
            launch_colsum(s.z, nz, d.n, M0, c->stream);
            CKR(allreduce(c, nz, size_t(M0)));
            launch_div_cols(s.mu, nz, size_t(ldG), M0, c->stream);
        }
    }   // else mu = 0 (already zeroed)
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    *out = guard.release();
    return SIMPLES_OK;
}
}   // namespace sbe

extern "C" {

int simples_em_create(const simples_em_args* a, simples_ctx** out) { return em_create_impl(a, out, 0); }

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
    const long long ldh = c->io_ld ? c->io_ld : d.n;       // rows of the full n x . host matrices
    if (c->write_shared) {                                 // replicated results: identical on every shard
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
        if (o->Varf) CK(cudaMemcpy(o->Varf, s.cc.Mm, size_t(M0) * K0 * K0 * 8, cudaMemcpyDefault));   // symmetric
        if (o->geneM) CK(cudaMemcpy(o->geneM, s.gmean, size_t(d.G) * 8, cudaMemcpyDefault));
        if (o->geneSd) CK(cudaMemcpy(o->geneSd, s.gsd, size_t(d.G) * 8, cudaMemcpyDefault));
        if (o->priorB) CK(cudaMemcpy(o->priorB, s.priorB, 8, cudaMemcpyDefault));
    }
    if (o->z) CKR(fetch_transposed(c, s.z, d.n, M0, M0, o->z, ldh));
    if (o->Ef)
        for (int m = 0; m < M0; ++m)
            CKR(fetch_transposed(c, s.W + size_t(m) * d.n * K0, d.n, K0, K0, o->Ef + size_t(m) * ldh * K0, ldh));
    if (o->Y)
        CKR(fetch_big(c, o->Y, [&](double* dst, size_t c0, size_t nc) {
            launch_uncenter(s.Y + c0 * d.ldG, dst, s.gmean, s.gsd, int(nc), d.G, d.ldG, c->stream);   // :335
        }));
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

}   // extern "C"

namespace sbe {

int em_impute_one(const simples_em_args* a, simples_em_out* o, long long io_ld, bool write_shared) {
    using clk = std::chrono::steady_clock;
    const bool trace = getenv("SIMPLES_TRACE") != nullptr;
    const auto t0 = clk::now();
    simples_ctx* c = nullptr;
    CKR(em_create_impl(a, &c, io_ld));
    c->write_shared = write_shared;
    const double mask_ms = c->host_mask_ms;
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
        fprintf(stderr, "[simples_b200] em_impute(dev %d): create(H2D+prologue) %.1f ms (host mask packing %.1f ms), run(%d it) %.1f ms, "
                "fetch(D2H) %.1f ms, destroy %.1f ms\n",
                t_device >= 0 ? t_device : g_device, ms(t0, t1), mask_ms, a->iter, ms(t1, t2), ms(t2, t3), ms(t3, t4));
    }
    return r;
}

// contiguous, balanced cell ranges (the same rule as simples_b200.dist.partition)
std::vector<long long> shard_bounds(long long n, int parts) {
    std::vector<long long> b(parts + 1, 0);
    const long long base = n / parts, rem = n % parts;
    for (int r = 0; r < parts; ++r) b[r + 1] = b[r] + base + (r < rem ? 1 : 0);
    return b;
}

// run body(rank) on one host thread per device; the first failure (lowest rank) is reported
template <typename F>
int run_on_devices(int ndev, F body) {
    std::vector<int> rc(ndev, SIMPLES_OK);
    std::vector<std::string> errs(ndev);
    std::vector<std::thread> th;
    for (int r = 0; r < ndev; ++r)
        th.emplace_back([&, r] {
            t_device = g_devs[r];
            rc[r] = body(r);
            errs[r] = g_err;
        });
    for (auto& t : th) t.join();
    for (int r = 0; r < ndev; ++r)
        if (rc[r] != SIMPLES_OK) return fail(rc[r], "device " + std::to_string(g_devs[r]) + ": " + errs[r]);
    return SIMPLES_OK;
}

bool multi_device_call(const void* cb, long long n, long long n_total, const void* stream) {
    return g_ndev > 1 && !cb && !stream && (n_total == 0 || n_total == n) && n >= g_ndev;
}

// One process, several GPUs: the cells (columns of Y / Y0, rows of z, Ef) are sharded over the devices, every shard
// runs the ordinary single-shard code on its own host thread, and the all-reduce is NCCL inside the library.
int em_impute_sharded(const simples_em_args* a, simples_em_out* o) {
    if (!a) return fail(SIMPLES_EINVAL, "null argument");
    CKR(check_common(a->G, a->n, a->M0, a->K0, a->MB));
    if (a->celltype)
        for (int i = 0; i < a->n; ++i)      // validated up front: a failure inside one shard would stall the others
            if (a->celltype[i] < 1 || a->celltype[i] > a->MB) return fail(SIMPLES_EINVAL, "celltype must be in 1..ncol(pg)");
    const int nd = g_ndev;
    const std::vector<long long> b = shard_bounds(a->n, nd);
    const size_t G = size_t(a->G);
    return run_on_devices(nd, [&](int r) -> int {
        simples_em_args la = *a;
        const long long c0 = b[r], nl = b[r + 1] - b[r];
        la.n = int(nl);
        la.Y = a->Y + size_t(c0) * G;
        la.Y0 = a->Y0 + size_t(c0) * G;
        la.z = a->z ? a->z + c0 : nullptr;
        la.celltype = a->celltype ? a->celltype + c0 : nullptr;
        la.n_total = a->n;
        la.cell_offset = c0;
        la.allreduce = nccl_allreduce_cb;
        la.allreduce_user = g_comms[r];
        simples_em_out lo{};
        if (o) {
            lo = *o;
            lo.z = o->z ? o->z + c0 : nullptr;
            lo.Ef = o->Ef ? o->Ef + c0 : nullptr;
            lo.Y = o->Y ? o->Y + size_t(c0) * G : nullptr;
        }
        return em_impute_one(&la, o ? &lo : nullptr, a->n, r == 0);
    });
}

}   // namespace sbe

extern "C" {

int simples_em_impute(const simples_em_args* a, simples_em_out* o) {
    if (a && multi_device_call((const void*)a->allreduce, a->n, a->n_total, a->stream)) return em_impute_sharded(a, o);
    return em_impute_one(a, o, 0, true);
}

}   // extern "C"

namespace sbe {
// ---------------------------------------------------------------------------------------------
int do_impute_one(const simples_mi_args* a, simples_mi_out* o, long long io_ld, bool write_shared) {
    if (!a || !o) return fail(SIMPLES_EINVAL, "null argument");
    CKR(check_common(a->G, a->n, a->M0, a->K0, a->MB));
    if (!a->dat || !a->Y || !a->beta || !a->lambda || !a->sigma || !a->mu || !a->pi || !a->pg)
        return fail(SIMPLES_EINVAL, "dat, Y, beta, lambda, sigma, mu, pi and pg are required");
    if (a->mcmc < 1 || a->burnin < 0) return fail(SIMPLES_EINVAL, "mcmc must be >= 1 and burnin >= 0");
    const bool cons_sharded = a->want_consensus && o->consensus_cluster && a->n_total > a->n;
    if (cons_sharded && (a->n_total > 2000000000LL / 8 || !(a->allreduce || g_rank_comm)))
        return fail(SIMPLES_EINVAL, "sharded consensus_cluster needs an all-reduce and n_total < 2.5e8");
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
    pick_allreduce(c, a->allreduce, a->allreduce_user, a->n, a->n_total);
    c->io_ld = io_ld;
    c->write_shared = write_shared;
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
    // accumulators of the recorded sweeps: one slot per zero entry, in the order of the sweep's own work lists
    CKR(c->dalloc(&c->mi_gbase, na / 8));
    CKR(c->dalloc(&c->mi_off, (na / 8) * size_t(ldG / 32)));
    launch_mi_index(s.mask, int(na), c->n_alloc, ldG / 32, c->mi_gbase, c->mi_off, c->stream);
    CKR(c->dalloc(&c->rec1, 2 * (size_t(c->nnz0) + 8)));     // interleaved {sum v, sum v^2}
    c->rec2 = c->rec1;
    CKR(c->dalloc(&c->rEF, size_t(n) * K0));
    CKR(c->dalloc(&c->rF2, size_t(n) * K0));
    CKR(c->dalloc(&c->rVF, size_t(n) * K0));
    const int nit = a->mcmc + a->burnin;
    CKR(c->dalloc(&c->tot_mi, size_t(nit) * (2 * M0 + 1)));
    // consensus_cluster (R/do_impute.R:61,96).  One shard: n x n.  Sharded: this shard's ROW BLOCK, n_local x n_total,
    // from the full z of every recorded sweep (each shard writes its rows into a zeroed n_total x M0 buffer and the
    // all-reduce sums the disjoint blocks).
    double* zall = nullptr;
    const long long ntot = a->n_total > 0 ? a->n_total : n;
    if (a->want_consensus && o->consensus_cluster) {
        if (cons_sharded) {
            CKR(c->dalloc(&c->cons, size_t(ntot) * n));
            CKR(c->dalloc(&zall, size_t(ntot) * M0));
        } else {
            CKR(c->dalloc(&c->cons, size_t(n) * n));
        }
    }
    const double pi_const = (c->compat & SIMPLES_COMPAT_PI) ? 3.14159265359 : M_PI;

    using clk_mi = std::chrono::steady_clock;
    const bool trace_mi = getenv("SIMPLES_TRACE") != nullptr;
    if (trace_mi) cudaStreamSynchronize(c->stream);
    const auto tm0 = clk_mi::now();
    CKR(run_prep(c));   // parameters are fixed: cluster constants computed once, not per sweep
    // (the opt-in fused MC pass is an EM-only path: it has no compact-accumulator variant and is slower, profiles/r02_notes.md 2.1)
    for (int it = 1; it <= nit; ++it) {                                       // :67
        const bool rec = it > a->burnin;
        c->sweep = unsigned(it - 1);
        CKR(e_pass(c, true, false, true, true, rec, false, pi_const, false));   // :69-104, :107, :112
        launch_finalize_cell(s.part_cell, s.ncta_cell, M0, c->tot_mi + size_t(it - 1) * (2 * M0 + 1), c->stream);   // :89
        if (rec && c->cons) {                                                   // :96
            if (zall) {
                CK(cudaMemsetAsync(zall, 0, size_t(ntot) * M0 * 8, c->stream));
                CK(cudaMemcpyAsync(zall + size_t(a->cell_offset) * M0, s.z, size_t(n) * M0 * 8, cudaMemcpyDeviceToDevice, c->stream));
                CKR(allreduce(c, zall, size_t(ntot) * M0));
                launch_consensus_rows(s.z, zall, c->cons, n, int(ntot), M0, c->stream);
            } else {
                launch_consensus(s.z, c->cons, n, M0, c->stream);
            }
        }
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
    if (o->loglik && c->write_shared) {
        std::vector<double> t(size_t(nit) * (2 * M0 + 1));
        CK(cudaMemcpy(t.data(), c->tot_mi, t.size() * 8, cudaMemcpyDefault));
        // mean(tot[-1:-burnin]) (R/do_impute.R:158).  For burnin = 0 the index is -1:0 = c(-1, 0), which still drops the
        // first sweep (Q21): the mean runs over tot[2..mcmc] and is NaN when mcmc = 1, exactly as in R.
        const int first = a->burnin > 0 ? a->burnin : 1;
        double sacc = 0.0;
        for (int it = first; it < nit; ++it) sacc += t[size_t(it) * (2 * M0 + 1) + 2 * M0];
        *o->loglik = nit > first ? sacc / double(nit - first) : std::nan("");
    }
    if (o->impt)
        CKR(fetch_big(c, o->impt, [&](double* dst, size_t c0, size_t nc) {
            launch_mi_finalize(s.Y + c0 * ldG, s.mask, c->rec1, c->rec2, s.gmean, s.gsd, dst, nullptr, int(nc), int(c0),
                               c->n_alloc, G, ldG, a->mcmc, c->mi_gbase, c->mi_off, c->stream);
        }));
    if (o->impt_var)
        CKR(fetch_big(c, o->impt_var, [&](double* dst, size_t c0, size_t nc) {
            launch_mi_finalize(s.Y + c0 * ldG, s.mask, c->rec1, c->rec2, s.gmean, s.gsd, nullptr, dst, int(nc), int(c0),
                               c->n_alloc, G, ldG, a->mcmc, c->mi_gbase, c->mi_off, c->stream);
        }));
    if (o->EF || o->varF) {
        std::vector<double> ef(size_t(n) * K0), f2(size_t(n) * K0), vf(size_t(n) * K0);
        CK(cudaMemcpy(ef.data(), c->rEF, ef.size() * 8, cudaMemcpyDefault));
        CK(cudaMemcpy(f2.data(), c->rF2, f2.size() * 8, cudaMemcpyDefault));
        CK(cudaMemcpy(vf.data(), c->rVF, vf.size() * 8, cudaMemcpyDefault));
        const double mc = a->mcmc;
        const size_t ldh = c->io_ld ? size_t(c->io_ld) : size_t(n);
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < K0; ++k) {
                const size_t e = size_t(i) * K0 + k, h = size_t(k) * ldh + i;
                const double m1 = ef[e] / mc;
                if (o->EF) o->EF[h] = m1;
                if (o->varF) o->varF[h] = f2[e] / mc - m1 * m1 + vf[e] / mc;
            }
    }
    if (o->consensus_cluster && c->cons) {
        if (zall) {      // row block: device [n_total columns][n_local rows] -> the caller's column-major matrix whose leading
                         // dimension is io_ld (the full n) in a one-process multi-GPU run, n_local for a rank's own block
            launch_scale(c->cons, size_t(ntot) * n, 1.0 / a->mcmc, c->stream);
            const size_t ldh = c->io_ld ? size_t(c->io_ld) : size_t(n);
            CK(cudaMemcpy2DAsync(o->consensus_cluster, ldh * 8, c->cons, size_t(n) * 8, size_t(n) * 8, size_t(ntot),
                                 cudaMemcpyDefault, c->stream));
        } else {
            launch_scale(c->cons, size_t(n) * n, 1.0 / a->mcmc, c->stream);
            CK(cudaMemcpyAsync(o->consensus_cluster, c->cons, size_t(n) * n * 8, cudaMemcpyDefault, c->stream));
        }
        CK(cudaStreamSynchronize(c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    return SIMPLES_OK;
}

int do_impute_sharded(const simples_mi_args* a, simples_mi_out* o) {
    if (!a || !o) return fail(SIMPLES_EINVAL, "null argument");
    CKR(check_common(a->G, a->n, a->M0, a->K0, a->MB));
    if (a->celltype)
        for (int i = 0; i < a->n; ++i)
            if (a->celltype[i] < 1 || a->celltype[i] > a->MB) return fail(SIMPLES_EINVAL, "celltype must be in 1..ncol(pg)");
    const int nd = g_ndev;
    const std::vector<long long> b = shard_bounds(a->n, nd);
    const size_t G = size_t(a->G);
    return run_on_devices(nd, [&](int r) -> int {
        simples_mi_args la = *a;
        const long long c0 = b[r], nl = b[r + 1] - b[r];
        la.n = int(nl);
        la.dat = a->dat + size_t(c0) * G;
        la.Y = a->Y + size_t(c0) * G;
        la.celltype = a->celltype ? a->celltype + c0 : nullptr;
        la.n_total = a->n;
        la.cell_offset = c0;
        la.allreduce = nccl_allreduce_cb;
        la.allreduce_user = g_comms[r];
        simples_mi_out lo = *o;
        lo.impt = o->impt ? o->impt + size_t(c0) * G : nullptr;
        lo.impt_var = o->impt_var ? o->impt_var + size_t(c0) * G : nullptr;
        lo.EF = o->EF ? o->EF + c0 : nullptr;
        lo.varF = o->varF ? o->varF + c0 : nullptr;
        lo.consensus_cluster = (a->want_consensus && o->consensus_cluster) ? o->consensus_cluster + c0 : nullptr;   // row block
        return do_impute_one(&la, &lo, a->n, r == 0);
    });
}

}   // namespace sbe

extern "C" {

int simples_do_impute(const simples_mi_args* a, simples_mi_out* o) {
    if (a && multi_device_call((const void*)a->allreduce, a->n, a->n_total, a->stream)) return do_impute_sharded(a, o);
    return do_impute_one(a, o, 0, true);
}

}  // extern "C"
