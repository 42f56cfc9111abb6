"""Generate the committed fixtures.  They come from the CPU oracle (the reference R package cannot
run here: no R, no glmnet/Rsolnp/msm/mixtools), so they pin REGRESSION of oracle and CUDA path, not
parity with R -- 'parity unpinned', see DESIGN.md.  tools/export_golden.R produces the same file
layout from the real reference for anyone with R."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as O  # noqa: E402
from util import make_case  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    K0, M0, iters, seed = 5, 3, 4, 17
    st = make_case(150, 200, 3, 3, K0, M0, seed=11)
    out = O.em_impute_fast(st["Y"], st["Y0"], st["pg"], M0, K0, 0.1, iters, st["beta"], st["sigma"], st["lam"], st["pi"],
                           st["z"], celltype=st["celltype"], impt_it=1, est_z=1, est_lam=1, seed=seed)
    mi = O.do_impute_ref(st["Y0"], out["Y"], out["beta"], out["lambda"], out["sigma"], out["mu"], out["pi"], out["geneM"],
                         out["geneSd"], st["celltype"], mcmc=4, burnin=2, pg=st["pg"], seed=seed + 1)
    np.savez_compressed(
        os.path.join(HERE, "em_small.npz"), K0=K0, M0=M0, iter=iters, seed=seed, Y=st["Y"], Y0=st["Y0"], pg=st["pg"],
        beta=st["beta"], sigma=st["sigma"], lam=st["lam"], pi=st["pi"], z=st["z"], celltype=st["celltype"],
        **{"out_" + k: out[k] for k in ("loglik", "pi", "mu", "sigma", "beta", "lambda", "z", "Y", "geneM")},
        out_Ef=np.stack(out["Ef"]), out_Varf=np.stack(out["Varf"]), out_priorB=out["priorB"],
        mi_loglik=mi["loglik"], mi_impt=mi["impt"], mi_impt_var=mi["impt_var"], mi_EF=mi["EF"], mi_varF=mi["varF"])


def stages():
    """Fixtures of the stages around the hot path: initial fit + draw, clustering start, lq-gene fit,
    and a whole SIMPLE() run, all from the oracle on one small simulated data set."""
    sim = O.simulate(n=140, G=110, K=2, MC=2, S0=8, seed=21)
    Y2, Z = sim["Y2"], sim["Z"]
    imp, pg, mu, sd = O.init_impute_ref(Y2, 2, Z, 0.4, 0.1, seed=3, sweep=1)
    pg1 = np.clip(np.random.default_rng(2).uniform(0.05, 1.2, (Y2.shape[0], 2)), 0.1, 1.0)
    impb = O.init_impute_bulk_ref(Y2, Z, pg1, 0.1, seed=4, sweep=2)
    par, pgz = O.ztruncnorm_ref(Y2, cutoff=0.1, p_min=0.6, p_max=1.0)
    clus, cm, d, wss = O.init_cluster_ref(Y2[:60], 4, 2, nstart=5, iter_max=15, seed=6)
    K0, M0 = 3, 2
    st = O.bench_state(sim, K0, M0, seed=21)
    em = O.em_impute_fast(st["Y"], st["Y0"], st["pg"], M0, K0, 0.1, 3, st["beta"], st["sigma"], st["lam"], st["pi"], st["z"],
                          celltype=st["celltype"], impt_it=1, est_z=1, est_lam=1, seed=2)
    lq = np.arange(40, 90)
    sig = np.random.default_rng(5).uniform(0.4, 1.6, (len(lq), M0))
    lq_mu, lq_beta, lq_sigma = O.lq_regress_ref(Y2[lq], Y2[lq], sig, em["z"], em["Ef"], em["Varf"], 1, 1)
    lq_Y = O.lq_impute_ref(Y2[lq], lq_mu, lq_beta, lq_sigma, st["pg"][lq], em["z"], em["Ef"], st["celltype"], 0.1, 9, 3)
    run = O.simple_ref(Y2, K0=K0, M0=M0, iter=3, clus=Z, min_gene=45, mcmc=2, burnin=1, seed=5)
    np.savez_compressed(
        os.path.join(HERE, "stages_small.npz"), Y2=Y2, Z=Z, init_impute=imp, init_pg=pg, init_mu=mu, init_sd=sd, pg1=pg1,
        init_bulk=impb, zt_param=par, zt_pg=pgz, clus=clus, cellmat=cm, d=d, wss=wss,
        em_z=em["z"], em_Ef=np.stack(em["Ef"]), em_Varf=np.stack(em["Varf"]), lq_pg=st["pg"][lq], celltype=st["celltype"],
        lq_sigma_in=sig, lq_mu=lq_mu, lq_beta=lq_beta, lq_sigma=lq_sigma, lq_Y=lq_Y,
        run_loglik_tot=run["loglik_tot"], run_beta=run["beta"], run_sigma=run["sigma"], run_mu=run["mu"], run_pi=run["pi"],
        run_lambda=run["lambda"], run_z=run["z"], run_impt=run["impt"], run_Yimp0=run["Yimp0"], run_BIC=run["BIC"],
        run_BIC0=run["BIC0"], run_pg=run["pg"])


if __name__ == "__main__":
    main()
    stages()
