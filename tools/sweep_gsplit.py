"""Tuning run: legacy impute_sweep time vs the gene-split factor (SIMPLES_SWEEP_GSPLIT) at several shard sizes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import simples_b200 as sb
sb.init(0)
C = bench.CFG
out = {}
for n in [int(x) for x in (sys.argv[1:] or ["100000", "50000", "25000", "12500"])]:
    st = bench.make_workload(n, 0, pinned=False)
    pos = (st["Y"], st["Y0"], st["pg"], C["M0"], C["K0"], C["cutoff"], st["beta"], st["sigma"], st["lam"], st["pi"], st["z"])
    kw = bench.em_kwargs(st, None)
    for gs in ["auto", "1", "2", "3", "4", "5", "6", "7", "8"]:
        if gs == "auto": os.environ.pop("SIMPLES_SWEEP_GSPLIT", None)
        else: os.environ["SIMPLES_SWEEP_GSPLIT"] = gs
        ctx = sb.EMContext(*pos, iter=8, **kw)
        ctx.run(3); ctx.sync()
        ctx.set_profile(True); ctx.run(3); prof = ctx.profile(); ctx.set_profile(False); ctx.close()
        out[f"{n}:{gs}"] = round(prof["impute_sweep"][1] / prof["impute_sweep"][0], 4)
    print(n, {k.split(":")[1]: v for k, v in out.items() if k.startswith(str(n) + ":")}, flush=True)
print(json.dumps(out))
