"""Run-level drivers: host mirrors of ``SIMPLE`` (R/SIMPLE.R:200-455), ``SIMPLE_B`` (R/SIMPLE_B.R:229-455) and
``selectKM`` (R/select_KM.R:126-189) on top of the CUDA entry points.

Same argument names, defaults, control flow and return-list names as the reference; every numerical stage runs in
libsimples_b200.so (init fit + draw, EM for the high-quality genes, the low-quality-gene regression + draw, EM for
all genes, Yimp0 / BIC, multiple imputation).  The host keeps only the bookkeeping the R drivers do between those
calls (gene selection, assembling ``pg`` / ``beta`` / ``sigma``).

Keyword-only extras: ``seed`` (each stage gets its own Philox key ``seed + stage``), ``comm`` (a
``simples_b200.dist.CellShard`` when the cells are sharded over GPUs; every matrix argument is then the rank's
shard), ``consensus`` (materialise the n x n consensus matrix; default only when it is < 2 GiB and unsharded),
``resident`` (keep ``dat`` and every G x n intermediate in HBM between the stages: one host-to-device copy of
``dat``, gene gathers / scatters on the device, and the G x n results -- ``impt``, ``impt_var``, ``Yimp0`` -- come
back as column-major torch CUDA tensors; the stage entry points take device pointers as they take host pointers).
"""
from __future__ import annotations

import math

import numpy as np

from . import api
from .cluster import init_cluster


# ---- G x n matrices on the host (numpy, F order) or in HBM (torch, column-major): gene gathers / scatters ----
def _rows(X, idx):
    if api._is_dev(X):
        import torch
        return X.T[:, torch.as_tensor(np.asarray(idx), device=X.device, dtype=torch.long)].T
    return np.asfortranarray(X[idx])


def _set_rows(X, idx, V):
    if api._is_dev(X):
        import torch
        V = V if api._is_dev(V) else api.dev_matrix(V)
        X.T[:, torch.as_tensor(np.asarray(idx), device=X.device, dtype=torch.long)] = V.T
    else:
        X[idx] = V if not api._is_dev(V) else V.cpu().numpy()


def _clone(X):
    return X.T.clone().T if api._is_dev(X) else X.copy(order="F")


def _out_like(X, rows):
    """Output buffer for a rows x n result next to X (device when X is), or None (= let the API allocate on the host)."""
    return api.dev_empty(rows, X.shape[1]) if api._is_dev(X) else None


def _row_sum(X, cols=None):
    """rowSums over all cells or over a boolean selection of them, as a host vector."""
    if api._is_dev(X):
        import torch
        Xs = X if cols is None else X[:, torch.as_tensor(cols, device=X.device)]
        return Xs.sum(dim=1, dtype=torch.float64).cpu().numpy()
    return (X if cols is None else X[:, cols]).sum(axis=1)


_STAGE_PMIN, _STAGE_INIT, _STAGE_EM_HQ, _STAGE_LQ, _STAGE_EM_ALL, _STAGE_MI, _STAGE_CLUS, _STAGE_BETA = range(8)


def _as_matrix(X, resident):
    """The G x n data matrix in R's layout: F-ordered numpy, or a column-major torch CUDA tensor when resident."""
    if api._is_dev(X):
        return api._f(X)
    X = np.asfortranarray(np.asarray(X, dtype=np.float64))
    return api.dev_matrix(X) if resident else X


def _sum_over_ranks(x, comm):
    x = np.ascontiguousarray(x, dtype=np.float64)
    return x if comm is None else comm.sum_host(x)


def _n_total(n, comm):
    return n if comm is None else int(comm.n_total)


def _hq_genes(dat, cutoff, p_min, min_gene, fix_num, comm):
    """R/SIMPLE.R:243-255 = R/SIMPLE_B.R:258-266: genes expressed in at least p_min of the cells, topped up to
    min_gene by expression rate (ties keep the original order, as R's radix order does)."""
    G, n = dat.shape
    n1 = _sum_over_ranks(_row_sum(dat > cutoff), comm) / _n_total(n, comm)
    top = lambda: np.argsort(-n1, kind="stable")[:min_gene]          # order(n1, decreasing = T)[1:min_gene]
    if min_gene > G:
        raise ValueError(f"min_gene = {min_gene} > G = {G}: the reference indexes NA rows here (R/SIMPLE.R:248,253); "
                         "pass min_gene <= G")
    if fix_num:
        hq = top()
    else:
        hq = np.nonzero(n1 >= p_min)[0]
        if len(hq) < min_gene:
            hq = top()
    return hq


def _finish(dat, em, ctx, pg, celltype, M0, K0, mcmc, burnin, cutoff, seed, comm, consensus, extra):
    """R/SIMPLE.R:421-453 = R/SIMPLE_B.R:429-454: Yimp0, BIC / BIC0, multiple imputation, the return list."""
    G, n = dat.shape
    yimp0 = ctx.yimp0(out=_out_like(dat, G))
    ctx.close()
    bic, bic0 = api.bic(em["loglik"][-1], _n_total(n, comm), G, K0, M0)
    out = dict(loglik_tot=em["loglik"], priorB=em["priorB"], loglik=None, BIC=bic, BIC0=bic0, pi=em["pi"], mu=em["mu"],
               sigma=em["sigma"], beta=em["beta"], z=em["z"], Yimp0=yimp0, pg=pg)
    out["lambda"] = em["lambda"]
    if mcmc > 0:
        want = consensus if consensus is not None else (comm is None and n * n * 8 < (2 << 30))
        mi = api.do_impute(dat, em["Y"], em["beta"], em["lambda"], em["sigma"], em["mu"], em["pi"], em["geneM"],
                           em["geneSd"], celltype, mcmc=mcmc, burnin=burnin, pg=pg, cutoff=cutoff,
                           seed=seed + _STAGE_MI, comm=comm, consensus=want,
                           out=dict(impt=_out_like(dat, G), impt_var=_out_like(dat, G)) if api._is_dev(dat) else None)
        out.update(loglik=mi["loglik"], impt=mi["impt"], impt_var=mi["impt_var"], Ef=mi["EF"], Varf=mi["varF"],
                   consensus_cluster=mi["consensus_cluster"])
    else:
        out.update(impt=em["Y"], impt_var=None, Ef=em["Ef"], Varf=em["Varf"], consensus_cluster=None)
    out.update(extra)
    order = ("loglik_tot", "priorB", "loglik", "BIC", "BIC0", "pi", "mu", "sigma", "beta", "lambda", "z", "Yimp0", "pg",
             "initclus", "impt", "impt_var", "Ef", "Varf", "consensus_cluster", "p_min")
    return {k: out[k] for k in order if k in out}


def _defaults(G, n, K0, M0, beta, lambda_, sigma, mu, seed):
    """R/SIMPLE.R:210-226: pi, lambda, mu, sigma and the random start of beta."""
    pi = np.full(M0, 1.0 / M0)
    lambda_ = np.full((M0, K0), 1.0 / M0) if lambda_ is None else np.asarray(lambda_, dtype=np.float64)
    mu = np.zeros((G, M0)) if mu is None else np.array(mu, dtype=np.float64)
    sigma = np.ones((G, M0)) if sigma is None else np.array(sigma, dtype=np.float64)
    if beta is None:
        beta = np.random.default_rng(seed + _STAGE_BETA).normal(size=(G, K0)) / math.sqrt(K0) * math.sqrt(M0)
    return pi, lambda_, mu, sigma, np.array(beta, dtype=np.float64)


def _lq_and_em_all(dat, Y, pg, hq, em_hq, beta, sigma, mu, celltype, M0, K0, iter, penl, est_z, max_lambda, est_lam,
                   sigma0, pi_alpha, num_mc, cutoff, seed, comm, resp, resid_mode):
    """R/SIMPLE.R:297-419 = R/SIMPLE_B.R:323-427: copy the hq fit, regress + draw the lq genes, EM on all genes."""
    G, n = dat.shape
    beta[hq], sigma[hq], mu[hq] = em_hq["beta"], em_hq["sigma"], em_hq["mu"]
    _set_rows(Y, hq, em_hq["Y"])
    lq = np.setdiff1d(np.arange(G), hq)
    if len(lq) == 0:
        raise ValueError("no low-quality genes left (min_gene == G): the reference fails here (R/SIMPLE.R:312,368)")
    fit = api.lq_fit(_rows(dat, lq), sigma[lq], pg[lq], em_hq["z"], em_hq["Ef"], em_hq["Varf"], celltype, penl=penl,
                     cutoff=cutoff, resp=None if resp is None else _rows(resp, lq), resid_mode=resid_mode,
                     seed=seed + _STAGE_LQ, comm=comm, out_Y=_out_like(dat, len(lq)))
    mu[lq], beta[lq], sigma[lq] = fit["mu"], fit["beta"], fit["sigma"]
    np.clip(sigma, 1e-4, 9.0, out=sigma)                                      # :372-373 (whole matrix)
    _set_rows(Y, lq, fit["Y"])
    ctx = api.EMContext(Y, dat, pg, M0, K0, cutoff, beta, sigma, em_hq["lambda"], em_hq["pi"], em_hq["z"], mu=None,
                        celltype=celltype, penl=penl, est_z=est_z, max_lambda=max_lambda, est_lam=est_lam, impt_it=1,
                        sigma0=sigma0, pi_alpha=pi_alpha, num_mc=num_mc, iter=iter, seed=seed + _STAGE_EM_ALL, comm=comm)
    ctx.run(iter)
    return ctx.fetch(out=dict(Y=_out_like(dat, G)) if api._is_dev(dat) else None), ctx


def SIMPLE(dat, K0=10, M0=1, iter=10, est_lam=1, impt_it=5, penl=1, init_imp=None, sigma0=100, pi_alpha=1, beta=None,
           verbose=False, max_lambda=True, lambda_=None, sigma=None, mu=None, est_z=1, clus=None, p_min=0.4, cutoff=0.1,
           K=20, min_gene=2000, num_mc=3, fix_num=False, mcmc=50, burnin=2, *, seed=0, comm=None, consensus=None,
           resident=False):
    """``SIMPLE`` (R/SIMPLE.R:200-455).  Returns a dict with the reference's list names in the reference's order:
    loglik_tot, priorB, loglik, BIC, BIC0, pi, mu, sigma, beta, lambda, z, Yimp0, pg, initclus, impt, impt_var, Ef,
    Varf, consensus_cluster, p_min."""
    dat = _as_matrix(dat, resident)
    if init_imp is not None:
        init_imp = _as_matrix(init_imp, resident)
    G, n = dat.shape
    Y = _clone(dat)
    pi, lambda_, mu, sigma, beta = _defaults(G, n, K0, M0, beta, lambda_, sigma, mu, seed)
    if clus is not None:
        clus = np.asarray(clus).astype(np.int32)
    if p_min is None:                                                         # :229-242
        if clus is None:
            pg_all = api.init_impute(dat, 1, np.ones(n, dtype=np.int32), 0.4, cutoff, seed=seed + _STAGE_PMIN, comm=comm,
                                     out=_out_like(dat, G))[1]
            pg_all = pg_all @ np.ones((1, M0))
        else:
            pg_all = api.init_impute(dat, M0, clus, 0.4, cutoff, seed=seed + _STAGE_PMIN, comm=comm, out=_out_like(dat, G))[1]
        p_min = float(np.quantile(pg_all.ravel(), 0.25))
    pg = np.full((G, M0), float(p_min))                                       # :243
    hq = _hq_genes(dat, cutoff, p_min, min_gene, fix_num, comm)               # :246-255
    if clus is None:                                                          # :258-266
        clus = init_cluster(_rows(dat, hq), K, M0, seed=seed + _STAGE_CLUS, comm=comm)
    z = np.zeros((n, M0))
    z[np.arange(n), clus - 1] = 1.0                                           # :268-269
    dat_hq = _rows(dat, hq)
    res = api.init_impute(dat_hq, M0, clus, p_min, cutoff, seed=seed + _STAGE_INIT, comm=comm,
                          out=_out_like(dat, len(hq)))                        # :272-277 (Q19)
    start = res[0] if init_imp is None else _rows(init_imp, hq)
    em_hq = api.EM_impute(start, dat_hq, res[1], M0, K0, cutoff, 20, beta[hq], sigma[hq], lambda_, pi, z, mu=None,
                          celltype=clus, penl=penl, est_z=est_z, max_lambda=max_lambda, est_lam=est_lam, impt_it=impt_it,
                          sigma0=sigma0, pi_alpha=pi_alpha, num_mc=num_mc, seed=seed + _STAGE_EM_HQ, comm=comm,
                          out=dict(Y=_out_like(dat, len(hq))) if api._is_dev(dat) else None)   # :281-296
    del dat_hq, start
    pg[hq] = res[1]                                                           # :297
    # lq genes: response init_imp when given; residual of sg dropped when it is not (inverted branch, :354-364)
    resp = init_imp
    em, ctx = _lq_and_em_all(dat, Y, pg, hq, em_hq, beta, sigma, mu, clus, M0, K0, iter, penl, est_z, max_lambda, est_lam,
                             sigma0, pi_alpha, num_mc, cutoff, seed, comm, resp, 0 if init_imp is None else 1)
    return _finish(dat, em, ctx, pg, clus, M0, K0, mcmc, burnin, cutoff, seed, comm, consensus,
                   dict(initclus=clus, p_min=p_min))


def SIMPLE_B(dat, K0, bulk, celltype, M0=1, clus=None, K=20, b=1, iter=10, est_z=1, impt_it=3, max_lambda=False,
             est_lam=1, penl=1, sigma0=100, pi_alpha=1, beta=None, lambda_=None, sigma=None, mu=None, p_min=0.8,
             min_gene=300, cutoff=0.1, verbose=False, num_mc=3, fix_num=False, mcmc=50, burnin=2, *, seed=0, comm=None,
             consensus=None, resident=False):
    """``SIMPLE_B`` (R/SIMPLE_B.R:229-455): bulk-informed dropout rates per cell type.  Same return names as
    ``SIMPLE`` minus ``initclus`` and ``p_min``."""
    dat = _as_matrix(dat, resident)
    G, n = dat.shape
    Y = _clone(dat)
    bulk = np.asarray(bulk, dtype=np.float64).reshape(G, -1)
    celltype = np.asarray(celltype).astype(np.int32)
    pi, lambda_, mu, sigma, beta = _defaults(G, n, K0, M0, beta, lambda_, sigma, mu, seed)
    hq = _hq_genes(dat, cutoff, p_min, min_gene, fix_num, comm)               # :258-266
    MB = int(celltype.max()) if comm is None else int(comm.max_host(np.array([celltype.max()], dtype=np.float64))[0])
    pg = np.zeros((G, MB))
    for i in range(MB):                                                       # :270-282
        sel = celltype == i + 1
        cnt = _sum_over_ranks(np.array([sel.sum()]), comm)[0]
        rm = _sum_over_ranks(_row_sum(dat, sel), comm) / cnt                  # rowMeans(dat[, celltype == i])
        if b is None:                                                         # weighted no-intercept lm, :274-277
            x, w = bulk[hq, i], bulk[hq, i] ** 2
            b1 = float(np.sum(w * x * rm[hq]) / np.sum(w * x * x))
        else:
            b1 = float(b)
        pg[:, i] = (rm + p_min) / (1.0 + b1 * bulk[:, i])
    np.clip(pg, 0.1, 1.0, out=pg)                                             # :284-285
    dat_hq = _rows(dat, hq)
    res = api.init_impute_bulk(dat_hq, celltype, bulk[hq], pg[hq], cutoff, seed=seed + _STAGE_INIT, comm=comm,
                               out=_out_like(dat, len(hq)))                   # :289-290
    if clus is None:                                                          # :293-313 (on the imputed hq matrix)
        clus = init_cluster(res, K, M0, seed=seed + _STAGE_CLUS, comm=comm)
    clus = np.asarray(clus).astype(np.int32)
    z = np.zeros((n, M0))
    z[np.arange(n), clus - 1] = 1.0                                           # :315-316
    em_hq = api.EM_impute(res, dat_hq, pg[hq], M0, K0, cutoff, 30, beta[hq], sigma[hq], lambda_, pi, z, mu=None,
                          celltype=celltype, penl=penl, est_z=est_z, max_lambda=max_lambda, est_lam=est_lam,
                          impt_it=impt_it, sigma0=sigma0, pi_alpha=pi_alpha, num_mc=num_mc, seed=seed + _STAGE_EM_HQ,
                          comm=comm, out=dict(Y=_out_like(dat, len(hq))) if api._is_dev(dat) else None)   # :318-321
    del dat_hq
    em, ctx = _lq_and_em_all(dat, Y, pg, hq, em_hq, beta, sigma, mu, celltype, M0, K0, iter, penl, est_z, max_lambda,
                             est_lam, sigma0, pi_alpha, num_mc, cutoff, seed, comm, None, 1)
    return _finish(dat, em, ctx, pg, celltype, M0, K0, mcmc, burnin, cutoff, seed, comm, consensus, {})


def selectKM(dat, bulk=None, celltype=None, b=1, K0=10, M0=1, iter=10, est_lam=1, impt_it=5, penl=1, sigma0=100,
             pi_alpha=1, beta=None, verbose=False, max_lambda=True, lambda_=None, sigma=None, mu=None, est_z=1, clus=None,
             p_min=0.4, cutoff=0.1, K=20, min_gene=300, num_mc=3, fix_num=False, mcmc=50, burnin=2, rel=0.01, *, seed=0,
             comm=None, consensus=None, resident=False, replicas=None, _simple=None):
    """``selectKM`` (R/select_KM.R:126-189): BIC grid over the number of clusters and factors with the reference's
    early stopping along K.  ``bulk`` / ``celltype`` / ``b`` are accepted and ignored exactly as in the reference
    (it always calls ``SIMPLE``, :152-155).  Returns BICs (dict of dicts [M][K]), mK, mM, result.

    ``replicas`` (a ``simples_b200.dist.Replicas``): one process per GPU, every process holds the whole ``dat``; the
    M = 1 pass runs everywhere (it provides ``init_imp`` for the others, :185), the remaining values of M are dealt
    round-robin to the ranks, each runs its own early-stopped K loop, and the BIC table is gathered.  The selected
    (mK, mM) is the same as the sequential loop's (first minimum in (M, K) order); ``result`` is the full run on the
    rank that computed it and ``None`` elsewhere."""
    run = SIMPLE if _simple is None else _simple
    dat = _as_matrix(dat, resident) if _simple is None else dat               # uploaded once for the whole grid
    n = _n_total(np.shape(dat)[1], comm)
    Ms = sorted(set(np.atleast_1d(M0).astype(int).tolist()) | {1})            # :136-137
    Ks = sorted(np.atleast_1d(K0).astype(int).tolist())
    state = dict(p_min=p_min)

    def k_loop(M1, init_imp):
        """The inner loop of :147-183 for one M: BIC row, and the first minimum (crit, K, result) of the row."""
        row, bestM, prev = {}, (math.inf, None, None), math.inf
        for K1 in Ks:
            result = run(dat, K1, M1, iter=iter, est_lam=est_lam, impt_it=impt_it, penl=penl, init_imp=init_imp,
                         sigma0=sigma0, pi_alpha=pi_alpha, beta=beta, verbose=verbose, max_lambda=max_lambda,
                         lambda_=lambda_, sigma=sigma, mu=mu, est_z=est_z, clus=clus, p_min=state["p_min"], cutoff=cutoff,
                         K=K, min_gene=min_gene, num_mc=num_mc, fix_num=fix_num, mcmc=mcmc, burnin=burnin, seed=seed,
                         comm=comm, consensus=consensus, resident=resident)
            if state["p_min"] is None:
                state["p_min"] = result["p_min"]                              # :156
            crit = result["BIC0"] if n < 1000 else result["BIC"]              # :158-182
            if bestM[0] > crit:
                bestM = (crit, K1, result)
            row[K1] = crit
            if (crit - prev) > -abs(crit) * rel:
                break
            prev = crit
        return row, bestM

    BICs = {m: {k: math.nan for k in Ks} for m in Ms}
    rows, bests = {}, {}
    rows[1], bests[1] = k_loop(1, None)                                       # every replica: init_imp for M > 1 (:185)
    init_imp = bests[1][2]["impt"]
    rest = [m for m in Ms if m != 1]
    mine = rest if replicas is None else rest[replicas.rank::replicas.world_size]
    for M1 in mine:                                                           # :145
        rows[M1], bests[M1] = k_loop(M1, init_imp)
    summary = {m: (rows[m], bests[m][0], bests[m][1]) for m in rows}
    if replicas is not None:
        merged = {}
        for part in replicas.allgather(summary):
            merged.update(part)
        summary = merged
    mBIC, mK, mM = math.inf, None, None
    for m in Ms:                                                              # first minimum in (M, K) order == :159-176
        BICs[m].update(summary[m][0])
        if mBIC > summary[m][1]:
            mBIC, mK, mM = summary[m][1], summary[m][2], m
    return dict(BICs=BICs, mK=mK, mM=mM, result=bests[mM][2] if mM in bests else None)
