"""do_impute throughput at the C3 shape (imputed entries/s = nnz0 (mcmc + burnin) / wall)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0]]
import numpy as np, torch
import bench, simples_b200 as sb
sb.init(0)
n, G, K0, M0 = bench.CFG["n"], bench.CFG["G"], bench.CFG["K0"], bench.CFG["M0"]
st = bench.make_workload(n, 0, pinned=True)
pos = (st["Y"], st["Y0"], st["pg"], M0, K0, 0.1, 3, st["beta"], st["sigma"], st["lam"], st["pi"], st["z"])
em = sb.EM_impute(*pos, **bench.em_kwargs(st))
outs, keep = {}, []
for k in ("impt", "impt_var"):
    t = torch.empty((n, G), dtype=torch.float64, pin_memory=True); keep.append(t); outs[k] = t.numpy().T
Ypin = torch.empty((n, G), dtype=torch.float64, pin_memory=True); Ypin.numpy()[...] = em["Y"].T
args = (st["Y0"], Ypin.numpy().T, em["beta"], em["lambda"], em["sigma"], em["mu"], em["pi"], em["geneM"], em["geneSd"], st["celltype"])
kw = dict(mcmc=10, burnin=2, pg=st["pg"], cutoff=0.1, seed=3, consensus=False, out=outs)
sb.do_impute(*args, **kw)                      # warm
os.environ["SIMPLES_TRACE"] = "1"
t0 = time.perf_counter(); r = sb.do_impute(*args, **kw); t = time.perf_counter() - t0
nnz0 = int((st["Y0"] <= 0.1).sum())
print(json.dumps({"do_impute_call_seconds": t, "sweeps": 12, "imputed_entries_per_sec_e2e": nnz0 * 12 / t, "nnz0": nnz0, "loglik": r["loglik"]}))
