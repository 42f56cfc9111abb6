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


if __name__ == "__main__":
    main()
