"""PLUMBING SELF-CHECK ONLY -- not a parity pin.  Writes a directory in the layout of tools/export_golden.R with the ORACLE
standing in for R (glmnet-compatible stopping rule, i.e. what real glmnet would return up to its own tolerance), so that
tests/test_r_golden.py -- which can only meet real R output outside this image -- is at least executed end to end: file
names, shapes, column-major order, dims.txt parsing, band computation.  Real vectors come from tools/export_golden.R."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O
from oracle import simples_oracle as SO


def wr(out, name, x):
    np.asarray(x, dtype="<f8").ravel(order="F").tofile(os.path.join(out, name + ".f64"))


def wri(out, name, x):
    np.asarray(x, dtype="<i4").ravel(order="F").tofile(os.path.join(out, name + ".i32"))


def main(out, n=120, G=160):
    os.makedirs(out, exist_ok=True)
    K0, M0, cutoff = 4, 3, 0.1
    sim = O.simulate(n=n, G=G, K=3, MC=M0, S0=8, seed=4)
    Y2, clus = sim["Y2"], sim["Z"]
    G, n = Y2.shape                     # the simulator drops genes with <= 4 non-zeros (utils/utils.R:88-90)
    dims = dict(G=G, n=n, K0=K0, M0=M0, cutoff=cutoff)
    SO.LASSO_MODE = "glmnet"            # the stand-in for R's glmnet
    try:
        # ---- A
        imp, pg, _, _ = O.init_impute_ref(Y2, M0, clus, 0.5, cutoff, seed=1)
        z = np.zeros((n, M0)); z[np.arange(n), clus - 1] = 1.0
        beta = np.random.default_rng(1).standard_normal((G, K0)) * np.sqrt(M0 / K0)
        a = dict(Y=imp, Y0=Y2, pg=pg, beta=beta, sigma=np.ones((G, M0)), **{"lambda": np.full((M0, K0), 1.0 / M0)},
                 pi=np.full(M0, 1.0 / M0), z=z)
        kw = dict(celltype=clus, est_z=1, est_lam=1, impt_it=5)
        fit = O.em_impute_fast(a["Y"], a["Y0"], a["pg"], M0, K0, cutoff, 5, a["beta"], a["sigma"], a["lambda"], a["pi"], a["z"], **kw)
        fit1 = O.em_impute_fast(a["Y"], a["Y0"], a["pg"], M0, K0, cutoff, 1, a["beta"], a["sigma"], a["lambda"], a["pi"], a["z"],
                                **dict(kw, impt_it=1))
        dims.update(A_iter=5, A_est_z=1, A_est_lam=1, A_impt_it=5)
        for k in ("Y", "Y0", "pg", "beta", "sigma", "lambda", "pi", "z"):
            wr(out, "A_in_" + k, a[k])
        wri(out, "A_in_celltype", clus)
        for k in ("loglik", "pi", "mu", "sigma", "beta", "lambda", "z", "Y", "geneM"):
            wr(out, "A_out_" + k, fit[k])
        wr(out, "A_out_priorB", [fit["priorB"]])
        wr(out, "A_out_Ef", np.concatenate(fit["Ef"], axis=1))
        wr(out, "A_out_Varf", np.concatenate(fit["Varf"], axis=1))
        for k in ("loglik", "pi", "mu", "sigma", "beta", "lambda", "z"):
            wr(out, "A1_out_" + k, fit1[k])
        # ---- B
        par, pgz = O.ztruncnorm_ref(Y2, cutoff=cutoff, p_min=0.5)
        wr(out, "B_zt_param", par); wr(out, "B_zt_pg", pgz)
        dims["B_p_min"] = 0.5
        # ---- C: hq EM + lq regression as SIMPLE() runs them
        n1 = (Y2 > cutoff).mean(axis=1)
        min_gene, p_min = G // 2, 0.5
        hq = np.nonzero(n1 >= p_min)[0]
        if hq.size < min_gene:
            hq = np.argsort(-n1, kind="stable")[:min_gene]
        lq = np.setdiff1d(np.arange(G), hq)
        imph, pgh, _, _ = O.init_impute_ref(Y2[hq], M0, clus, p_min, cutoff, seed=2)
        c1 = O.em_impute_fast(imph, Y2[hq], pgh, M0, K0, cutoff, 20, beta[hq], np.ones((hq.size, M0)),
                              np.full((M0, K0), 1.0 / M0), np.full(M0, 1.0 / M0), z, celltype=clus, est_z=1, est_lam=1, impt_it=20)
        mu_l, beta_l, sig_l = O.lq_regress_fast(Y2[lq], Y2[lq], np.ones((lq.size, M0)), c1["z"], c1["Ef"], c1["Varf"], penl=1,
                                                resid_mode=0)
        dims.update(C_Ghq=hq.size, C_Glq=lq.size, C_iter1=20, C_impt_it1=20, C_min_gene=min_gene, C_p_min=p_min)
        wri(out, "C_hq_ind", hq + 1); wri(out, "C_lq_ind", lq + 1)
        wr(out, "C1_out_z", c1["z"]); wr(out, "C1_out_Ef", np.concatenate(c1["Ef"], axis=1))
        wr(out, "C1_out_Varf", np.concatenate(c1["Varf"], axis=1))
        wr(out, "C_lq_beta", beta_l); wr(out, "C_lq_sigma", sig_l); wr(out, "C_dat", Y2)
    finally:
        SO.LASSO_MODE = "exact"
    # ---- D (no lasso involved)
    mi = O.do_impute_ref(Y2, fit["Y"], fit["beta"], fit["lambda"], fit["sigma"], fit["mu"], fit["pi"], fit["geneM"], np.ones(G),
                         clus, mcmc=1, burnin=0, pg=pg, cutoff=cutoff, seed=9, consensus=True)
    wr(out, "D_EF", mi["EF"]); wr(out, "D_varF", mi["varF"]); wr(out, "D_consensus", mi["consensus_cluster"])
    mi20 = O.do_impute_ref(Y2, fit["Y"], fit["beta"], fit["lambda"], fit["sigma"], fit["mu"], fit["pi"], fit["geneM"],
                           np.ones(G), clus, mcmc=20, burnin=2, pg=pg, cutoff=cutoff, seed=10, consensus=False)
    zmask = Y2 <= cutoff
    wr(out, "D_impt_zero_mean", (mi20["impt"] * zmask).sum(1) / np.maximum(zmask.sum(1), 1))
    wr(out, "D_loglik20", [mi20["loglik"]])
    dims.update(D_mcmc=20, D_burnin=2)
    with open(os.path.join(out, "dims.txt"), "w") as f:
        for k, v in dims.items():
            f.write(f"{k}={v!r}\n")


if __name__ == "__main__":
    main(sys.argv[1])
