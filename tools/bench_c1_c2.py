"""BASELINE configs[0] ("C1": the vignette simulation, ~300 cells x 1000 genes, K0 = 10, M0 = 3, SIMPLE()) and configs[1]
("C2": SIMPLE_B with bulk means and 7 cell types, K0 = 10; chu.RData is not in the checkout, SURVEY F7, so the stated
substitute is used: 3000 genes x 763 cells simulated with 7 clusters, bulk = per-type mean profile): whole-run wall time on
one B200 through the Python drivers over the C ABI, and the NumPy oracle's time for the same call on the host cores.
The R package itself cannot be timed here (no R)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oracle as O            # checker / CPU baseline only
import simples_b200 as sb

torch.cuda.set_device(0); sb.init(0)
out = {}

def timed(fn, reps=3):
    fn()                       # warm: module load, allocator pools
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return r, min(ts)

# ---- C1: utils/simulation.R:44,65-67 (n = 300, K = 6 blocks of 32, 3 clusters; SIMPLE(K0 = 10, M0 = 3, p_min = 0.5, min_gene = 200))
sim = O.simulate(n=300, G=1000, K=6, MC=3, S0=20, seed=1)
kw = dict(K0=10, M0=3, iter=10, clus=None, K=20, p_min=0.5, cutoff=0.01, min_gene=200, mcmc=50, burnin=2, seed=1)
res, t = timed(lambda: sb.SIMPLE(sim["Y2"], **kw))
_, tr = timed(lambda: sb.SIMPLE(sim["Y2"], resident=True, **kw))
out["C1"] = {"shape": list(sim["Y2"].shape), "SIMPLE_wall_s": round(t, 4), "SIMPLE_resident_wall_s": round(tr, 4),
             "ari_vs_truth": O.adjusted_rand_index(np.argmax(np.asarray(res["z"]), axis=1), sim["Z"]),
             "phases": "init fit, irlba-equivalent + kmeans(nstart = 300), EM 20 it on hq genes, lq fit, EM 10 it on all genes, 52 MI sweeps"}
t0 = time.perf_counter(); O.simple_ref(sim["Y2"], **dict(kw, clus=sim["Z"])); out["C1"]["oracle_cpu_s_given_clusters"] = round(time.perf_counter() - t0, 2)
_, tg = timed(lambda: sb.SIMPLE(sim["Y2"], **dict(kw, clus=sim["Z"])))
out["C1"]["SIMPLE_wall_s_given_clusters"] = round(tg, 4)

# ---- C2 substitute
sim2 = O.simulate(n=763, G=3000, K=6, MC=7, S0=20, seed=11)
bulk = np.maximum(sim2["Mu"], 0.0)
kw2 = dict(M0=7, clus=sim2["Z"], b=1, iter=10, min_gene=300, mcmc=50, burnin=2, seed=8)
res2, t2 = timed(lambda: sb.SIMPLE_B(sim2["Y2"], 10, bulk, sim2["Z"], **kw2))
out["C2_substitute"] = {"shape": list(sim2["Y2"].shape), "celltypes": 7, "SIMPLE_B_wall_s": round(t2, 4)}
t0 = time.perf_counter(); ref2 = O.simple_b_ref(sim2["Y2"], 10, bulk, sim2["Z"], **kw2); out["C2_substitute"]["oracle_cpu_s"] = round(time.perf_counter() - t0, 2)
rel = lambda a, b: float(np.abs(np.asarray(a) - np.asarray(b)).max() / (np.abs(np.asarray(b)).max() + 1e-300))
out["C2_substitute"]["rel_err_vs_oracle"] = {k: rel(res2[k], ref2[k]) for k in ("beta", "sigma", "mu", "impt", "loglik_tot")}
out["cores"] = os.cpu_count()
print(json.dumps(out))
