"""Sweep throughput vs zero fraction at the bench shape (diagnostic): extra dropout is applied to Y0."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import simples_b200 as sb
import bench
torch.cuda.set_device(0); sb.init(0)
st = bench.make_workload(bench.CFG["n"], 0, pinned=False)
G, K0, M0 = bench.CFG["G"], bench.CFG["K0"], bench.CFG["M0"]
kw = bench.em_kwargs(st, None)
rng = np.random.default_rng(0)
res = []
for keep in (1.0, 0.5, 0.2, 0.1):
    Y0 = np.asfortranarray(np.asarray(st["Y0"]) * (rng.uniform(size=st["Y0"].shape) < keep))
    pos = (st["Y"], Y0, st["pg"], M0, K0, bench.CFG["cutoff"], st["beta"], st["sigma"], st["lam"], st["pi"], st["z"])
    ctx = sb.EMContext(*pos, iter=8, **kw)
    ctx.run(3); ctx.sync()
    ctx.set_profile(True)
    ctx.run(3); ctx.sync()
    pr = ctx.profile()
    nnz = ctx.nnz0
    sw = pr["impute_sweep"]
    res.append(dict(keep=keep, rho=nnz / (G * bench.CFG["n"]), sweep_ms=sw[1] / sw[0], entries_per_s=nnz / (sw[1] / sw[0] * 1e-3)))
    ctx.close()
    print(json.dumps(res[-1]), flush=True)
