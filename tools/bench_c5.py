"""BASELINE configs[4] ("C5"): the selectKM grid (R/select_KM.R:126-189) on 200k cells x 2k genes over GPU replicas.

Run under torchrun (one process per GPU).  Every rank holds the whole data set device-resident; the M = 1 pass runs on
every rank (it yields init_imp for the others, :185), the other values of M are dealt round-robin, every rank runs its own
K loop with the reference's early stop (relative BIC decrease < rel, :169,180), and only the BIC table is gathered.
Prints one JSON line on rank 0: wall time, SIMPLE() runs executed (total and the longest rank), the BIC table with the
early-stop pattern (NaN = never run), the selected (M, K).
  torchrun --nproc-per-node 8 tools/bench_c5.py [--cells 200000] [--K0 5 10 15 20 25 30] [--M0 1 3 5 8 10 12 15 18 20]
"""
import argparse
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import simples_b200 as sb
from simples_b200 import synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=200_000)
    ap.add_argument("--genes", type=int, default=2000)
    ap.add_argument("--K0", type=int, nargs="+", default=[5, 10, 15, 20, 25, 30])
    ap.add_argument("--M0", type=int, nargs="+", default=[1, 3, 5, 8, 10, 12, 15, 18, 20])
    ap.add_argument("--iter", type=int, default=10)
    ap.add_argument("--mcmc", type=int, default=10)
    ap.add_argument("--rel", type=float, default=0.01)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    sb.init(local)
    rep = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        rep = sb.Replicas()
    gp = synth.gene_params(a.genes, 10, 10, seed=1)
    Y2, Z = synth.simulate_shard(gp, a.cells, shard=0, seed=1)          # the SAME data on every rank
    Y2 = np.asfortranarray(Y2)
    runs = {"n": 0, "s": 0.0, "log": []}
    from simples_b200 import driver

    def counted(dat, K1, M1, **kw):
        t0 = time.perf_counter()
        r = driver.SIMPLE(dat, K1, M1, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        runs["n"] += 1
        runs["s"] += dt
        runs["log"].append((M1, K1, round(dt, 3)))
        return r

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dat = driver._as_matrix(Y2, True)                                    # one upload for the whole grid
    out = sb.selectKM(dat, K0=a.K0, M0=a.M0, iter=a.iter, mcmc=a.mcmc, burnin=2, min_gene=1000, p_min=0.4, K=20, rel=a.rel,
                      seed=3, resident=True, replicas=rep, consensus=False, _simple=counted)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    stats = [runs]
    if world > 1:
        stats = [None] * world
        dist.all_gather_object(stats, runs)
        dist.barrier()
    if rank == 0:
        line = {"config": f"C5: selectKM grid on {a.cells} cells x {a.genes} genes, K0 in {a.K0}, M0 in {sorted(set(a.M0) | {1})}, "
                          f"iter={a.iter}, mcmc={a.mcmc}, rel={a.rel}; {world} GPU replica(s), data device-resident",
                "n_gpus": world, "wall_s": round(wall, 3),
                "simple_runs_total": int(sum(s["n"] for s in stats)), "simple_runs_longest_rank": int(max(s["n"] for s in stats)),
                "simple_seconds_per_rank": [round(s["s"], 2) for s in stats],
                "runs_per_rank": [s["log"] for s in stats],
                "grid_cells": len(a.K0) * len(set(a.M0) | {1}),
                "BIC": {str(m): {str(k): (None if (isinstance(v, float) and math.isnan(v)) else float(v)) for k, v in row.items()}
                        for m, row in out["BICs"].items()},
                "selected": {"mM": int(out["mM"]), "mK": int(out["mK"])}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


main()
