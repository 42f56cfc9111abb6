"""Developer script: field-by-field GPU vs oracle comparison on small problems."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as O
import simples_b200 as sb


def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))


def compare(tag, got, ref, M0):
    out = []
    for k in ["loglik", "pi", "mu", "sigma", "beta", "lambda", "z", "Y", "geneM", "geneSd"]:
        out.append(f"{k}={rel(got[k], ref[k]):.2e}")
    out.append("Ef=%.2e" % max(rel(got["Ef"][m], ref["Ef"][m]) for m in range(M0)))
    out.append("Varf=%.2e" % max(rel(got["Varf"][m], ref["Varf"][m]) for m in range(M0)))
    out.append("priorB=%.2e" % (abs(got["priorB"] - ref["priorB"]) / (abs(ref["priorB"]) + 1e-300)))
    print(tag, " ".join(out), flush=True)


def case(n, G, K, MC, K0, M0, iters, impt_it, seed, nonshared=False, **kw):
    sim = O.simulate(n=n, G=G, K=K, MC=MC, S0=10, seed=seed)
    st = O.bench_state(sim, K0, M0)
    sg = st["sigma"]
    if nonshared:
        sg = sg * np.random.default_rng(0).uniform(0.5, 1.5, size=sg.shape)
    args = (st["Y"], st["Y0"], st["pg"], M0, K0, 0.1, iters, st["beta"], sg, st["lam"], st["pi"], st["z"])
    kws = dict(celltype=st["celltype"], impt_it=impt_it, est_z=1, est_lam=1, seed=7, **kw)
    t = time.time(); ref = O.em_impute_fast(*args, **kws); t_or = time.time() - t
    t = time.time(); got = sb.EM_impute(*args, **kws); t_gpu = time.time() - t
    compare(f"[n={n} G={st['Y'].shape[0]} K0={K0} M0={M0} it={iters} impt_it={impt_it} ns={nonshared}] oracle {t_or:.2f}s gpu {t_gpu:.2f}s:", got, ref, M0)
    return st, got, ref


if __name__ == "__main__":
    case(200, 150, 3, 3, 4, 3, 1, 5, 3)          # one deterministic iteration
    case(200, 150, 3, 3, 4, 3, 4, 9, 3)          # deterministic EM
    case(200, 150, 3, 3, 4, 3, 3, 1, 3)          # with MC sweeps
    case(200, 150, 3, 3, 4, 3, 3, 1, 3, nonshared=True)
    case(1000, 500, 6, 5, 10, 5, 4, 1, 4)
    st, got, ref = case(300, 260, 3, 1, 5, 1, 3, 1, 5)   # M0 = 1
    # do_impute
    r_or = O.do_impute_ref(st["Y0"], ref["Y"], ref["beta"], ref["lambda"], ref["sigma"], ref["mu"], ref["pi"], ref["geneM"],
                           ref["geneSd"], st["celltype"], mcmc=4, burnin=2, pg=st["pg"], seed=11)
    r_gpu = sb.do_impute(st["Y0"], ref["Y"], ref["beta"], ref["lambda"], ref["sigma"], ref["mu"], ref["pi"], ref["geneM"],
                         ref["geneSd"], st["celltype"], mcmc=4, burnin=2, pg=st["pg"], seed=11)
    print("do_impute:", " ".join(f"{k}={rel(r_gpu[k], r_or[k]):.2e}" for k in ["loglik", "impt", "impt_var", "EF", "varF", "consensus_cluster"]))
