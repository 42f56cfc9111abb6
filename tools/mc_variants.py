"""Tuning run: time the fused MC pass variants (SIMPLES_MC_CFG, experiment build) on the C3 workload in one process."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import simples_b200 as sb

def main():
    n = int(os.environ.get("CELLS", "100000"))
    sb.init(0)
    st = bench.make_workload(n, 0, pinned=False)
    C = bench.CFG
    pos = (st["Y"], st["Y0"], st["pg"], C["M0"], C["K0"], C["cutoff"], st["beta"], st["sigma"], st["lam"], st["pi"], st["z"])
    kw = bench.em_kwargs(st, None)
    out = {}
    for cfg in sys.argv[1:] or ["0"]:
        if cfg == "legacy":
            os.environ.pop("SIMPLES_FUSED_MC", None)
        else:
            os.environ["SIMPLES_FUSED_MC"] = "1"
            os.environ["SIMPLES_MC_CFG"] = cfg     # needs a build with -DSIMPLES_MC_EXPERIMENT
        ctx = sb.EMContext(*pos, iter=12, **kw)
        ctx.run(3); ctx.sync()
        ctx.run(6); ctx.sync()
        ms = ctx.iter_ms()
        ctx.set_profile(True); ctx.run(2); prof = ctx.profile(); ctx.set_profile(False)
        ll = ctx.fetch(big=False)["loglik"]
        ctx.close()
        out[cfg] = dict(ms_per_iter=sum(ms) / len(ms), sweep_ms=prof["impute_sweep"][1] / prof["impute_sweep"][0], loglik=float(ll[-1]))
        print(cfg, out[cfg], flush=True)
    print(json.dumps(out))

main()
