"""Stage timings of a SIMPLE() run at the bench workload (C3: 1e5 cells x 2000 genes, K0 = M0 = 10, 1000 hq genes):
init_impute, init_cluster, EM on hq genes, lq_fit, EM on all genes, do_impute -- each through its C entry point with
host buffers (SIMPLES_TRACE splits H2D / compute / D2H), plus the CPU oracle on a bounded sample of each stage."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import simples_b200 as sb
import bench
from simples_b200 import synth

torch.cuda.set_device(0); sb.init(0)
os.environ["SIMPLES_TRACE"] = "1"
n = int(os.environ.get("STAGE_CELLS", bench.CFG["n"]))
gp = synth.gene_params(bench.CFG["G"], bench.CFG["K"], bench.CFG["MC"], seed=1)
Y2, Z = synth.simulate_shard(gp, n, shard=0, seed=1)
Y2 = np.asfortranarray(Y2)
G = Y2.shape[0]
out = {"n": n, "G": G}
def timed(name, fn):
    t0 = time.perf_counter(); r = fn(); out[name + "_s"] = round(time.perf_counter() - t0, 4); return r
n1 = (Y2 > 0.1).mean(axis=1)
hq = np.argsort(-n1, kind="stable")[:1000]
lq = np.setdiff1d(np.arange(G), hq)
clus = np.asarray(Z).astype(np.int32)
sb.init_impute(Y2[hq][:, :1000], 10, clus[:1000], 0.4, 0.1)            # warm the pools / module
res = timed("init_impute", lambda: sb.init_impute(Y2[hq], 10, clus, 0.4, 0.1, seed=1))
cl = timed("init_cluster", lambda: sb.init_cluster(Y2[hq], 20, 10, seed=2, details=True))
import oracle as O
out["init_cluster_ari_vs_truth"] = O.adjusted_rand_index(cl["clus"], Z)
out["init_cluster_wss"] = cl["tot_withinss"]
full = timed("SIMPLE_total", lambda: sb.SIMPLE(Y2, K0=10, M0=10, iter=10, clus=clus, p_min=0.4, min_gene=1000, mcmc=10, burnin=2, seed=1, consensus=False))
resd = timed("SIMPLE_resident_total", lambda: sb.SIMPLE(Y2, K0=10, M0=10, iter=10, clus=clus, p_min=0.4, min_gene=1000, mcmc=10, burnin=2, seed=1, consensus=False, resident=True))
torch.cuda.synchronize()
out["resident_equals_host"] = bool(np.array_equal(full["beta"], resd["beta"]))
out["SIMPLE_ari_vs_truth"] = O.adjusted_rand_index(np.argmax(full["z"], axis=1), Z)
out["SIMPLE_mse_zero_entries"] = None
# CPU oracle on bounded samples (same code path as the parity tests)
ns = 2000
t0 = time.perf_counter(); O.init_impute_ref(Y2[hq][:, :ns], 10, clus[:ns], 0.4, 0.1, seed=1); out["cpu_init_impute_s_per_2000_cells"] = round(time.perf_counter() - t0, 3)
t0 = time.perf_counter(); O.init_cluster_ref(Y2[hq][:, :ns], 20, 10, nstart=10, iter_max=80, seed=2); out["cpu_init_cluster_s_per_2000_cells_10_starts"] = round(time.perf_counter() - t0, 3)
print(json.dumps(out))
