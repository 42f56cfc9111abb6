"""Per-iteration device time of the first EM iterations at the bench workload (diagnostic)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import simples_b200 as sb
import bench
torch.cuda.set_device(0); sb.init(0)
st = bench.make_workload(bench.CFG["n"], 0, pinned=False)
G, K0, M0 = bench.CFG["G"], bench.CFG["K0"], bench.CFG["M0"]
pos = (st["Y"], st["Y0"], st["pg"], M0, K0, bench.CFG["cutoff"], st["beta"], st["sigma"], st["lam"], st["pi"], st["z"])
kw = bench.em_kwargs(st, None)
ctx = sb.EMContext(*pos, iter=12, **kw)
ctx.run(12); ctx.sync()
print("iter_ms", [round(x, 2) for x in ctx.iter_ms()])
ctx.close()
ctx = sb.EMContext(*pos, iter=4, **kw)
ctx.set_profile(True)
ctx.run(3); ctx.sync()
print("profile it1-3", {k: (v[0], round(v[1], 2)) for k, v in ctx.profile().items()})
