// C ABI of the stages around the hot path: the low-quality-gene fit between the two EM_impute calls,
// the initial per-gene fit + imputation, and the clustering start.  Same conventions as engine.cu.
#include "engine_internal.h"

extern "C" {

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
    pick_allreduce(c, a->allreduce, a->allreduce_user, a->n, a->n_total);
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
        PrepArgs pa{d, s.beta, s.sigma, s.mu, s.lam, s.pi, s.gmean, s.gsd, s.Q, s.isg, s.Qs, s.sdsA, s.isdA, s.musdA, s.gmig, s.cc, 0,
                    nullptr, nullptr, s.part_prep};
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
    pick_allreduce(c, a->allreduce, a->allreduce_user, a->n, a->n_total);
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
    pick_allreduce(c, a->allreduce, a->allreduce_user, a->n, a->n_total);
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
        ProjArgs pa{Y, A, isg, P, S2, n, ldG, NQP, QS, nullptr};
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
