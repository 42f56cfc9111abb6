"""Timing of the clustering start at the bench shape (1000 hq genes x 1e5 cells, K = 20, M0 = 10, 300 starts)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import simples_b200 as sb, bench
from simples_b200 import synth
import oracle as O
torch.cuda.set_device(0); sb.init(0)
os.environ["SIMPLES_TRACE"] = "1"
gp = synth.gene_params(bench.CFG["G"], bench.CFG["K"], bench.CFG["MC"], seed=1)
Y2, Z = synth.simulate_shard(gp, bench.CFG["n"], shard=0, seed=1)
n1 = (Y2 > 0.1).mean(axis=1)
hq = np.argsort(-n1, kind="stable")[:1000]
X = sb.dev_matrix(Y2[hq])
for rep in range(3):
    t0 = time.perf_counter(); r = sb.init_cluster(X, 20, 10, seed=2, details=True); t = time.perf_counter() - t0
    print(json.dumps({"rep": rep, "seconds": round(t, 4), "ari": O.adjusted_rand_index(r["clus"], Z), "wss": r["tot_withinss"]}), flush=True)
