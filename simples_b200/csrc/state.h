// Device-resident state of one EM_impute / do_impute context and the launch
// interface of every kernel.  Layouts (all FP64 unless noted):
//   Y      [n][ldG]        cell-major, genes contiguous (== R's column-major G x n with padded ld)
//   mask   [ldG/32][n_alloc] u32 bitmask (chunk-major), bit g%32 of word [g/32][i] set <=> Y0[g,i] <= cutoff
//   beta   [ldG][K0]   sigma/mu [ldG][M0]   pg [ldG][MB]   gmean/gsd [ldG]   (gene-major rows)
//   lam    [M0][K0]   pi [M0]   z [n][M0]
//   Q      [NS][ldG][QS]   projection operand  [B/sigma_s | mu_m/sigma_cs(m) | 0-pad]
//   isg    [NS][ldG]       1/sigma_s
//   Qs     [ldG][FS]       sweep operand       [B*sd | 0-pad]
//   sdsA   [ldG][M0]       sqrt(sigma)*sd,  isdA = 1/sdsA,  musdA = mu*sd + mean,  gmig [ldG][2] = {gene mean, 1/sd}
//   P      [NS][n][NQP]    Y^T Q          S2 [NS][n]   sum_g Y^2/sigma_s
//   V      [n][VS]         M-step operand [Wbar (K0) | z (M0) | sum_m sqrt z_m (1) | 0-pad],  zsum [n]
//   Fh     [n][FS]         sweep operand  [f_i (K0) | 0-pad]   (the membership m_i is in im[n])
//   W      [M0][n][K0]     factor posterior means (Ef)
// ldG = G rounded up to 64; padded genes have Y = 0, mask = 0, beta = mu = 0, sigma = 1.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace sb {

constexpr int GPAD = 64;     // ldG granularity
constexpr int MAXK = 32;     // K0 <= 32
constexpr int MAXM = 32;     // M0 <= 32
constexpr int NUM_SMS_B200 = 148;
constexpr int PROJ_MAX_SPLIT = 8;   // k_proj: parts per tile of the gene-split last round

enum KernelClass {
    KC_PROJ = 0,      // estep_project   (DMMA)
    KC_CELL,          // cell_post       (responsibilities, factor means, MC cell draws)
    KC_SWEEP,         // impute_sweep    (DMMA + per-entry sampling)
    KC_MSTATS,        // mstep_stats     (DMMA split-K)
    KC_SSTATS,        // per-cluster K0xK0 second moments
    KC_REDUCE,        // deterministic split-K / partial reductions
    KC_PREP,          // cluster_prep
    KC_GENE,          // mstep_gene_solve (lasso, sigma, priorB)
    KC_LMP,           // lambda / mu / pi
    KC_MISC,          // prologue, mask build, transposes, records
    KC_COMM,          // all-reduce callback (time on stream)
    KC_COUNT
};

struct Dims {
    int G, ldG, n, M0, K0, MB;
    int NS;            // distinct sigma columns in use: 1 (cluster-shared) or M0
    int NQ, NQP, QS;   // projection cols: NQ = K0 + M0 (per sigma column), padded to 8; row stride QS = NQP + 4
    int NF, NFP, FS;   // sweep GEMM depth: NF = K0 padded to 4 (NFP); row stride FS = roundup(K0, 8) + 4 (== 4 mod 8)
    int NV, NVP, VS;   // M-step cols: NV = K0 + M0 + 1, padded to 8; row stride VS = NVP + 4
    long long n_total, cell_offset;
};

struct ClusterConsts {   // device pointers, produced by cluster_prep
    double* Mm;     // [M0][K0][K0]  (B' Sigma_m^-1 B + Lambda_m^-1)^-1
    double* MmP;    // [M0][KR][KPS] zero-padded copy of Mm for DMMA B-fragments (KR = 4 ceil(K0/4), KPS = 8 ceil(K0/8) + 4)
    double* Mh;     // [M0][K0][K0]  symmetric square root of Mm
    double* dt;     // [M0]
    double* off;    // [M0][K0]   mu_m' (B / sigma_m)
    double* offq;   // [M0][K0]   mu_m' (B / sigma_q)   q = M0-1 (stale beta2, Q1) -- used for the quadratic form in MC passes
    double* cmu;    // [M0]       sum_g mu_gm^2 / sigma_gm
    double* logpi;  // [M0]
    double* Vj;     // [M0][K0][K0] eigenvectors of Mm from the previous cluster_prep (warm start of the Jacobi sweeps)
    int* jst;       // [M0] cluster_prep calls since the last cold start (0: none yet)
};

struct DevState {
    Dims d;
    double *Y, *beta, *sigma, *mu, *pg, *gmean, *gsd, *lam, *pi, *z;
    uint32_t* mask;
    int32_t* celltype0;
    double *Q, *isg, *Qs, *sdsA, *isdA, *musdA, *gmig, *P, *S2, *V, *zsum, *Fh, *W;
    double *BM, *gv;     // fused MC pass operands: [ldG][QS] rows [B_g | mu_g | 0], [ldG][4] {isd, sds, mean, 1/sd}
    unsigned* mc_counter;
    double* proj_part;   // k_proj tail-split scratch
    int n_alloc;
    int32_t* im;
    ClusterConsts cc;
    double* Mm_snap;   // [M0][K0][K0] copy of Mm at the last E-pass (returned as Varf)
    // statistics: stats = [C (ldG x NVP) | R2 (ldG) | S (M0 K0 K0) | a (M0 K0) | nz (M0) | c (M0) | tot (1)]
    double* stats;
    size_t stats_count;
    double *part_ms;   // mstats split-K partials [splitK][ldG][NVP+1]
    double *part_cell; // cell_post partials [ncta][2*M0+1]
    double *part_ss;   // sstats partials [nsplit][M0][K0*K0+K0]
    int splitK, ncta_cell, nsplit_ss;
    // general (per-cluster sigma) path, allocated lazily
    double *Vg;        // [n][K0*M0 + 2*M0] = [z_m W_m | z | sqrt z]
    double *statsg;    // [T (M0 ldG K0) | U | Uh | U2 (ldG M0 each) | S | a | nz | c | tot]
    size_t statsg_count;
    double *loglik;    // [iter_cap]
    double *priorB;    // [2]: current, previous
    double *part_gene; // per-CTA priorB partials
    int ncta_gene;
    double *part_mu;   // k_mu_acc partial sums
    double *part_prep; // k_gram partial sums
};

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE attribute: remembered per (kernel, current device),
// thread-safe, so that one process can drive several GPUs (one host thread each) and re-initialise on another device.
void ensure_dyn_smem(const void* kernel, size_t bytes);
template <typename K>
inline void ensure_dyn_smem_k(K kernel, size_t bytes) { ensure_dyn_smem(reinterpret_cast<const void*>(kernel), bytes); }

// ---- launchers (each in its own .cu) ------------------------------------------------------------
struct ProjArgs {
    const double* Y; const double* Q; const double* isg; double* P; double* S2;
    int n, ldG, NQP, QS;
    double* part;   // scratch [PROJ_MAX_SPLIT * NUM_SMS * 128][NQP + 1] for the gene-split tail tiles (nullptr: no split)
};
void launch_proj(const ProjArgs& a, const CUtensorMap& tmapY, cudaStream_t st);   // tmapY: box {16 genes, 128 cells}
int proj_launches(int n, int ldG, bool have_scratch);

struct CellArgs {
    Dims d;
    const double *P, *S2;
    ClusterConsts cc;
    double* z;            // in/out
    double *V, *zsum, *W, *Fh, *Vg;
    int32_t* im;
    double* part;         // [ncta][2*M0+1]: nz, c, tot
    int ncta;
    int update_z, stale_q, write_W, write_Vg, sample, draw_membership;
    int expect;           // SIMPLES_MODE_EXPECT: membership = argmax z, factor = posterior mean (no draws)
    double ll_const;      // G * log(2 PI)
    unsigned long long seed;
    unsigned sweep;
};
void launch_cell(const CellArgs& a, cudaStream_t st);

struct SweepArgs {
    Dims d;
    double* Y; const uint32_t* mask; const double *Qs, *Fh, *sdsA, *isdA, *musdA, *pg, *gmig;
    int n_alloc;
    const int32_t *im, *celltype0;
    double cutoff, lower, upper;
    double hi, lo, hi_trunc;   // clamp bounds folded by launch_sweep: [lower, upper] or +-inf, min(cutoff, hi)
    int nst;                   // parameter ring depth, chosen by launch_sweep
    int clip;
    unsigned long long seed; unsigned sweep;
    uint32_t rk[20];       // Philox round keys, filled by launch_sweep
    // optional do_impute accumulators (nullptr otherwise), COMPACT: one slot per zero entry, ordered by (8-cell group,
    // 32-gene chunk, cell row, gene) = the order of the sweep's own work lists, so a warp's updates of a block fall into
    // ~113 consecutive slots (the dense G x n layout of round 1 cost a scattered read-modify-write per entry)
    double *rec1, *rec2;   // rec1: interleaved {sum v, sum v^2} per slot; rec2 is not used by k_sweep (non-null = record)
    const unsigned long long* mi_gbase;   // [ngroups]          first slot of the 8-cell group
    const uint32_t* mi_off;               // [ngroups][nwords]  first slot of the block inside its group
    int expect;            // SIMPLES_MODE_EXPECT: conditional expectation instead of a draw
};
void launch_sweep(SweepArgs a, cudaStream_t st);

// fused MC pass (k_mcpass.cuh): impute_sweep + the projection of the following E-pass; cluster-shared sigma only
struct McArgs {
    Dims d;
    const uint32_t* mask; int n_alloc;
    const double* Fh; const int32_t *im, *celltype0;
    const double* BM;      // [ldG][QS]  rows [B_g | mu_g | 0]
    const double* gv;      // [ldG][4]   {1/(sqrt(sigma) sd), sqrt(sigma) sd, gene mean, 1/sd}  (per-entry gathers)
    const double *gmean, *gsd, *isg;   // [ldG] dense per-gene vectors (isg = 1/sigma)
    const double* pg;      // [ldG][MB]
    double cutoff, lower, upper; int clip;
    unsigned long long seed; unsigned sweep;
    uint32_t rk[20];       // Philox round keys k_r = (seed_lo + r W0, seed_hi + r W1), filled by launch_mcpass
    double *P, *S2;        // outputs of the projection: [n][NQP], [n]
    double *rec1, *rec2;   // optional do_impute accumulators (layout of Y), nullptr otherwise
    unsigned* counter;     // task counter, zeroed by the launcher
};
bool mcpass_supported(const Dims& d);
void launch_mcpass(McArgs a, const CUtensorMap& tmapY8, cudaStream_t st);   // tmapY8: box {16 genes, 8 cells}

struct MstatsArgs {
    const double *Y, *V, *zsum; double* part;   // part [splitK][ldG][NVP+1]
    int n, ldG, NVP, VS, splitK;
};
void launch_mstats(const MstatsArgs& a, cudaStream_t st);
int mstats_pick_splitk(int ldG);

// generic slow GEMM for the general-sigma path and the prologue: C[g][c] = sum_i f(Y[i][g]) * V[i][c]
void launch_ygemm_simple(const double* Y, const double* V, double* part, int n, int ldG, int ncol, int vstride,
                         int square, int nsplit, cudaStream_t st);
void launch_reduce_parts(const double* part, double* out, size_t count, int nparts, cudaStream_t st);

struct SstatsArgs { Dims d; const double *W, *z; double* part; int nsplit; };
void launch_sstats(const SstatsArgs& a, cudaStream_t st);

struct FinalizeArgs {
    Dims d; const double *part_ms, *part_cell, *part_ss; int splitK, ncta_cell, nsplit_ss; double* stats;
    size_t head; int skip_head;   // general-sigma path: the first `head` entries are produced elsewhere
};
void launch_finalize_stats(const FinalizeArgs& a, cudaStream_t st);
// sums only the cell_post partials: out = [nz (M0) | c (M0) | tot]
void launch_finalize_cell(const double* part_cell, int ncta, int M0, double* out, cudaStream_t st);

struct PrepArgs {
    Dims d; const double *beta, *sigma, *mu, *lam, *pi, *gmean, *gsd;
    double *Q, *isg, *Qs, *sdsA, *isdA, *musdA, *gmig; ClusterConsts cc; int stale_q;
    double *BM, *gv;     // may be nullptr
    double* part_prep;   // k_gram partials: prep_scratch_doubles()
};
void launch_prep(const PrepArgs& a, cudaStream_t st, cudaStream_t st_cluster);
size_t prep_scratch_doubles(const Dims& d, int NS_max);

struct MstepArgs {
    Dims d;
    const double* stats;      // shared-sigma layout (or statsg when general)
    int general;
    double *beta, *sigma, *mu, *lam, *pi;
    const double* Mm;         // pre-M-step cluster matrices
    double *loglik_slot;      // loglik[it-1]
    double *priorB;           // [0]=current (written), [1]=previous (read, then updated)
    double *part_gene; int ncta_gene;
    double *part_mu;          // [M0][mu_gene_splits][K0*K0 + K0] partial sums of the mu update
    double penl, sigma0, pi_alpha;
    int do_lambda, compat_priorb_prev;
    int debug;                // count coordinate-descent sweeps (SIMPLES_TRACE)
};
void launch_mstep(const MstepArgs& a, cudaStream_t st);
int mstep_gene_ctas(int G);
int mu_gene_splits(int ldG);
void lasso_debug_fetch(int out[4], bool reset);   // {solves, total CD sweeps, max sweeps, capped} since the last reset

// low-quality-gene fit (R/SIMPLE.R:305-411)
struct LqSolveArgs {
    Dims d; const double *stats, *stats_d, *Varf;   // stats: response statistics (general layout); stats_d: those of dat
    double *beta, *sigma, *mu; double penl; int resid_mode;
};
void launch_lq_solve(const LqSolveArgs& a, cudaStream_t st);
struct LqCellArgs {
    Dims d; const double *z, *W; double *Vg, *Fh; int32_t* im;
    unsigned long long seed; unsigned sweep; int draw_membership;
};
void launch_lq_cell(const LqCellArgs& a, cudaStream_t st);
void launch_colsum(const double* z, double* out, int n, int M0, cudaStream_t st);   // out [2 M0 + 1] = [colSums(z) | 0...]

// initial fit + imputation (R/fit_ztrunc.R, R/SIMPLE.R:15-73, R/SIMPLE_B.R:23-114); Y2 is [n][G] unpadded
struct InitFitArgs {
    const double* stats;      // [M0][5][G] moments | [M0] cluster sizes
    int G, M0; double cutoff, p_min, s_lower, s_upper;
    const double* pg1;        // [M0][G] or nullptr
    double *mu, *sd, *pg_used, *pg_ret;   // [M0][G]
};
struct InitDrawArgs {
    const double* Y2; double* out; const int32_t* clus0; int n, G;
    const double *mu, *sd, *pg_used, *pg1; double cutoff;
    unsigned long long seed; unsigned sweep; long long cell_offset;
};
int init_stats_splits(int n);
void launch_init_stats(const double* Y2, const int32_t* clus0, int n, int G, int M0, double cutoff, double* part, int nsplit,
                       cudaStream_t st);
void launch_ztrunc(const InitFitArgs& a, cudaStream_t st);
void launch_init_draw(const InitDrawArgs& a, cudaStream_t st);

// initial clustering (R/SIMPLE.R:258-266): Gram blocks through mstep_stats, subspace iteration, Lloyd k-means
struct KmArgs {
    const double* X; int ldx, n, K, M0, R;   // X [n][ldx]: cellmat rows
    double* cent;       // [R][M0][K]
    double* part;       // [nsplit * 4][R][M0][K + 1] per-warp sums | counts
    double* wss_part;   // [nsplit * 4][R]
    int* moved;         // [R] 1 while the restart's centres still move
    int nsplit;
};
void launch_extract_cols(const double* Y, double* V, int n, int ldG, int G, int g0, int ncol, int VS, cudaStream_t st);
void launch_scatter_block(const double* blk, double* C, int rows, int ldc, int NVP, int g0, int ncol, cudaStream_t st);
void launch_sd_from_ss(const double* ss, double* sd, double* zero, int G, int ldG, double nm1, cudaStream_t st);
void launch_subspace(const double* C, double* Q, double* Z, double* A, double* dvals, int G, int ldG, int KP, int K, int QS,
                     int iters, unsigned long long seed, cudaStream_t st);
void launch_km_init(const KmArgs& a, unsigned long long seed, long long n_total, long long cell_offset, cudaStream_t st);
void launch_km_pass(const KmArgs& a, int mode, int rsel, int32_t* labels, cudaStream_t st);
void launch_km_update(const KmArgs& a, const double* sums, int* nlive, cudaStream_t st);
void launch_cellmat_out(const double* P, double* out, int n, int K, int ldx, cudaStream_t st);

// misc
void launch_build_mask(const double* Y0chunk, uint32_t* mask, int ncell, int cell0, int n_alloc, int G, int ldG, double cutoff,
                       unsigned long long* nnz, cudaStream_t st);
void launch_count_nonfinite(const double* p, size_t count, unsigned long long* out, cudaStream_t st);
void launch_clip(double* Y, int n, int G, int ldG, double lower, double upper, cudaStream_t st);
void launch_rowsum(const double* Y, double* part, int n, int ldG, int nsplit, cudaStream_t st);
void launch_center(double* Y, const double* gmean, const double* gsd, int n, int ldG, double scale_n, int mode, cudaStream_t st);
void launch_transpose(const double* in, double* out, int rows, int cols, int ld_in, int ld_out, cudaStream_t st);
void launch_fill(double* p, size_t count, double v, cudaStream_t st);
void launch_scale(double* p, size_t count, double s, cudaStream_t st);
void launch_div_cols(double* A, const double* d, size_t rows, int cols, cudaStream_t st);
void launch_uncenter(const double* Y, double* out, const double* gmean, const double* gsd, int n, int G, int ldG, cudaStream_t st);
void launch_mi_finalize(const double* Y, const uint32_t* mask, const double* rec1, const double* rec2, const double* gmean,
                        const double* gsd, double* impt, double* impt_var, int n, int cell0, int n_alloc, int G, int ldG, int mcmc,
                        const unsigned long long* gbase, const uint32_t* off, cudaStream_t st);
// slot table of the compact do_impute accumulators (mask is fixed for the whole call): returns nothing, fills gbase / off
void launch_mi_index(const uint32_t* mask, int n, int n_alloc, int nwords, unsigned long long* gbase, uint32_t* off,
                     cudaStream_t st);
void launch_mi_record_factors(const Dims& d, const double* W, const double* z, const double* Mm, double* rEF, double* rF2,
                              double* rVF, cudaStream_t st);
void launch_consensus(const double* z, double* cons, int n, int M0, cudaStream_t st);
void launch_consensus_rows(const double* zloc, const double* zall, double* cons, int nl, int ntot, int M0, cudaStream_t st);
void launch_yimp0(const Dims& d, const double* mu, const double* beta, const double* z, const double* V, const double* gmean,
                  const double* gsd, double* out, int cell0, int ncell, cudaStream_t st);

}  // namespace sb
