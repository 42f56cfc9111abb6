"""Multi-GPU invariance check (run under torchrun): cells sharded over N ranks + NCCL all-reduce must
reproduce the single-GPU result (parameters to reduction-order rounding, Y / z shards equal to the
corresponding slices), because the Philox tape is keyed by the GLOBAL cell index."""
import os, sys, json
_TESTS = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(_TESTS))
sys.path.insert(0, _TESTS)
import numpy as np
import torch, torch.distributed as dist
import simples_b200 as sb
from util import make_case, em_args, rel

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); sb.init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
K0, M0 = 6, 4
st = make_case(3001, 700, 4, 4, K0, M0, seed=31)          # same on every rank (seeded)
G, n = st["Y"].shape
kw = dict(impt_it=1, est_z=1, est_lam=1, seed=13, num_mc=2)
sh = sb.CellShard(n)
part = sb.EM_impute(sh.cols(st["Y"]), sh.cols(st["Y0"]), st["pg"], M0, K0, 0.1, 4, st["beta"], st["sigma"], st["lam"],
                    st["pi"], sh.rows(st["z"]), celltype=sh.rows(st["celltype"]), comm=sh, **kw)
mi = sb.do_impute(sh.cols(st["Y0"]), part["Y"], part["beta"], part["lambda"], part["sigma"], part["mu"], part["pi"],
                  part["geneM"], part["geneSd"], sh.rows(st["celltype"]), mcmc=3, burnin=1, pg=st["pg"], seed=5, comm=sh,
                  consensus="rows")
ok = True
msgs = []
# the same sharded call with the all-reduce inside the library (its own NCCL rank communicator, no Python callback)
shn = sb.CellShard(n, native=True)
partn = sb.EM_impute(shn.cols(st["Y"]), shn.cols(st["Y0"]), st["pg"], M0, K0, 0.1, 4, st["beta"], st["sigma"], st["lam"],
                     st["pi"], shn.rows(st["z"]), celltype=shn.rows(st["celltype"]), comm=shn, **kw)
for k in ("loglik", "pi", "mu", "sigma", "beta", "lambda", "Y", "z"):
    e = rel(partn[k], part[k]); ok &= e <= 1e-12
    if rank == 0: msgs.append(f"native.{k}={e:.1e}")
assert shn.n_allreduce == 0 and sb.load().simples_comm_nranks() == world
if rank == 0:
    full = sb.EM_impute(*em_args(st, M0, K0, 4), celltype=st["celltype"], **kw)
    fmi = sb.do_impute(st["Y0"], full["Y"], full["beta"], full["lambda"], full["sigma"], full["mu"], full["pi"],
                       full["geneM"], full["geneSd"], st["celltype"], mcmc=3, burnin=1, pg=st["pg"], seed=5, consensus=True)
    for k in ("loglik", "pi", "mu", "sigma", "beta", "lambda", "geneM"):
        e = rel(part[k], full[k]); msgs.append(f"{k}={e:.1e}"); ok &= e <= 1e-9
    a, b = sh.bounds[0], sh.bounds[1]
    for k, sl in (("Y", full["Y"][:, a:b]), ("z", full["z"][a:b])):
        e = rel(part[k], sl); msgs.append(f"{k}[shard0]={e:.1e}"); ok &= e <= 1e-8
    e = abs(mi["loglik"] - fmi["loglik"]) / abs(fmi["loglik"]); msgs.append(f"mi.loglik={e:.1e}"); ok &= e <= 1e-9
    e = rel(mi["impt"], fmi["impt"][:, a:b]); msgs.append(f"mi.impt[shard0]={e:.1e}"); ok &= e <= 1e-8
    e = rel(mi["consensus_cluster"], fmi["consensus_cluster"][a:b, :]); msgs.append(f"mi.consensus[rows of shard0]={e:.1e}"); ok &= e <= 1e-9
    print(json.dumps({"world": world, "ok": bool(ok), "allreduces_em+mi": sh.n_allreduce, "errs": msgs}), flush=True)
# ---- run-level driver, cells sharded: init fit, hq EM, lq fit, all-gene EM, MI and the clustering start ----
import oracle as O
sim = O.simulate(n=1201, G=400, K=3, MC=3, S0=12, seed=9)
shd = sb.CellShard(sim["Y2"].shape[1])
dkw = dict(K0=4, M0=3, iter=3, K=5, min_gene=150, mcmc=2, burnin=1, seed=3)
for clus in (sim["Z"], None):
    pr = sb.SIMPLE(shd.cols(sim["Y2"]), clus=None if clus is None else shd.rows(clus), comm=shd, **dkw)
    if rank == 0:
        fr = sb.SIMPLE(sim["Y2"], clus=clus, consensus=False, **dkw)
        a, b = shd.bounds[0], shd.bounds[1]
        tag = "clus" if clus is not None else "kmeans"
        for k in ("loglik_tot", "beta", "sigma", "mu", "pi", "lambda"):
            e = rel(pr[k], fr[k]); msgs.append(f"SIMPLE[{tag}].{k}={e:.1e}"); ok &= e <= 1e-7
        for k, sl in (("impt", fr["impt"][:, a:b]), ("Yimp0", fr["Yimp0"][:, a:b]), ("z", fr["z"][a:b])):
            e = rel(pr[k], sl); msgs.append(f"SIMPLE[{tag}].{k}[shard0]={e:.1e}"); ok &= e <= 1e-6
        ok &= bool(np.array_equal(pr["initclus"], fr["initclus"][a:b]))
        e = abs(pr["BIC"] - fr["BIC"]) / abs(fr["BIC"]); msgs.append(f"SIMPLE[{tag}].BIC={e:.1e}"); ok &= e <= 1e-9
# ---- selectKM with the grid dealt over replicas (every rank holds the whole data) vs the sequential loop ----
skw = dict(K0=[2, 3, 4], M0=[1, 2, 3], iter=2, min_gene=150, mcmc=0, seed=4, clus=None, K=5, resident=True)
rep = sb.selectKM(sim["Y2"], replicas=sb.Replicas(), **skw)
if rank == 0:
    seq = sb.selectKM(sim["Y2"], **skw)
    same = (rep["mK"], rep["mM"]) == (seq["mK"], seq["mM"])
    for m in seq["BICs"]:
        for k in seq["BICs"][m]:
            a, b = rep["BICs"][m][k], seq["BICs"][m][k]
            same &= bool((np.isnan(a) and np.isnan(b)) or abs(a - b) <= 1e-9 * abs(b))
    msgs.append(f"selectKM replicas == sequential: {same} (mM={seq['mM']}, mK={seq['mK']})"); ok &= same
if rank == 0:
    print(json.dumps({"world": world, "ok": bool(ok), "driver": msgs[-21:]}), flush=True)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
