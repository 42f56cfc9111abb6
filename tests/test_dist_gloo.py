"""world_size-2 gloo tests (CPU) of the N > 1 host logic: cell partitioning, the all-reduce callback
wiring (simples_b200.dist.CellShard.allreduce_sum_ on raw pointers), and that sharded sufficient
statistics + one all-reduce reproduce the single-rank M-step (checked with the oracle)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_balanced_contiguous():
    from simples_b200 import partition
    for n, w in ((10, 3), (100000, 8), (7, 7), (5, 8)):
        b = partition(n, w)
        assert b[0] == 0 and b[-1] == n and len(b) == w + 1
        sizes = np.diff(b)
        assert sizes.min() >= 0 and sizes.max() - sizes.min() <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    from simples_b200 import CellShard
    from util import em_args, make_case
    K0, M0 = 4, 3
    st = make_case(101, 120, 3, 3, K0, M0, seed=21)
    G, n = st["Y"].shape
    sh = CellShard(n)
    assert sh.n_total == n and sh.n_local == sh.bounds[rank + 1] - sh.bounds[rank]
    # one deterministic E-pass on the local shard, then the library-style all-reduce of packed stats
    Yc = st["Y"] - st["Y"].mean(axis=1)[:, None]            # (global gene means: every rank has the full matrix here)
    mu = Yc @ st["z"] / n
    Yl, zl = sh.cols(Yc), sh.rows(st["z"])
    assert Yl.shape == (G, sh.n_local) and zl.shape == (sh.n_local, M0)
    ll, lik, sc, Ms, _, _ = O.epass(Yl, st["beta"], st["sigma"], st["lam"], mu, st["pi"])
    z = lik / lik.sum(axis=1)[:, None]
    Wm = O.factor_means(Yl, st["beta"], st["sigma"], mu, Ms)
    buf = np.ascontiguousarray(O.stats_pack(O.local_stats(Yl, z, Wm, True)))
    sh.allreduce_sum_(buf.ctypes.data, buf.size, None)      # same entry point the CUDA library calls back
    stt = O.stats_unpack(buf, G, M0, K0, True)
    out = O.mstep_from_stats(stt, n, Ms, st["beta"], st["sigma"], st["lam"], mu, M0, K0, 1, 1, True, 1, 100, 1)
    if rank == 0:
        q.put([np.asarray(o) for o in out[:5]] + [sh.n_allreduce, sh.bytes_allreduce])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_reproduces_single_rank_mstep():
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as O
    from util import make_case
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    K0, M0 = 4, 3
    st = make_case(101, 120, 3, 3, K0, M0, seed=21)
    G, n = st["Y"].shape
    Yc = st["Y"] - st["Y"].mean(axis=1)[:, None]
    mu = Yc @ st["z"] / n
    ll, lik, sc, Ms, _, _ = O.epass(Yc, st["beta"], st["sigma"], st["lam"], mu, st["pi"])
    z = lik / lik.sum(axis=1)[:, None]
    Wm = O.factor_means(Yc, st["beta"], st["sigma"], mu, Ms)
    ref = O.mstep_from_stats(O.local_stats(Yc, z, Wm, True), n, Ms, st["beta"], st["sigma"], st["lam"], mu, M0, K0, 1, 1,
                             True, 1, 100, 1)
    for a, b in zip(got[:5], ref[:5]):
        assert np.allclose(a, b, rtol=1e-9, atol=1e-12)
    assert got[5] == 1 and got[6] == 8 * (G * K0 + 2 * G * M0 + G + M0 * K0 * K0 + M0 * K0 + 2 * M0)


def _driver_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from simples_b200 import CellShard
    from simples_b200.driver import _hq_genes, _n_total, _sum_over_ranks
    rng = np.random.default_rng(5)
    dat = rng.uniform(size=(40, 31)) * (rng.uniform(size=(40, 31)) < rng.uniform(0.1, 0.9, size=(40, 1)))
    sh = CellShard(31)
    loc = sh.cols(dat)
    hq_a = _hq_genes(loc, 0.1, 0.5, 12, False, sh)            # expression rates summed over the ranks
    hq_b = _hq_genes(loc, 0.1, 0.99, 12, False, sh)           # too few pass -> top-up branch
    tot = _sum_over_ranks(np.array([loc.shape[1]]), sh)[0]
    mx = sh.max_host(np.array([float(rank + 1)]))[0]
    if rank == 0:
        q.put((hq_a, hq_b, tot, mx, _n_total(loc.shape[1], sh)))
    dist.barrier()
    dist.destroy_process_group()


def test_driver_host_logic_sharded():
    """SIMPLE()'s gene selection on sharded cells (R/SIMPLE.R:246-255) equals the unsharded selection."""
    import torch.multiprocessing as mp
    from simples_b200.driver import _hq_genes
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 7
    ps = [ctx.Process(target=_driver_worker, args=(r, 2, port, q)) for r in range(2)]
    for p_ in ps:
        p_.start()
    hq_a, hq_b, tot, mx, nt = q.get(timeout=120)
    for p_ in ps:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    rng = np.random.default_rng(5)
    dat = rng.uniform(size=(40, 31)) * (rng.uniform(size=(40, 31)) < rng.uniform(0.1, 0.9, size=(40, 1)))
    assert np.array_equal(hq_a, _hq_genes(dat, 0.1, 0.5, 12, False, None))
    assert np.array_equal(hq_b, _hq_genes(dat, 0.1, 0.99, 12, False, None)) and len(hq_b) == 12
    assert tot == 31 and mx == 2.0 and nt == 31


def _fake_simple(dat, K1, M1, **kw):
    """Stand-in for SIMPLE(): a deterministic BIC surface with a unique minimum at (M, K) = (3, 6)."""
    bic = 1000.0 + 40.0 * (M1 - 3) ** 2 + 9.0 * (K1 - 6) ** 2 - (5.0 if kw.get("init_imp") is not None else 0.0)
    return dict(BIC=bic, BIC0=bic + 1.0, impt=("impt", M1, K1), p_min=0.37, tag=(M1, K1))


def _km_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from simples_b200 import Replicas
    from simples_b200.driver import selectKM
    dat = np.zeros((5, 2000))                                   # n >= 1000 -> BIC (not BIC0) is the criterion
    out = selectKM(dat, K0=[2, 4, 6, 8, 10], M0=[1, 2, 3, 4, 5], p_min=None, replicas=Replicas(), _simple=_fake_simple)
    q.put((rank, out["BICs"], out["mK"], out["mM"], None if out["result"] is None else out["result"]["tag"]))
    dist.barrier()
    dist.destroy_process_group()


def test_select_km_replicas_equal_sequential():
    """selectKM over replicas (values of M dealt to the ranks) == the sequential loop of R/select_KM.R:145-186."""
    import torch.multiprocessing as mp
    from simples_b200.driver import selectKM
    seq = selectKM(np.zeros((5, 2000)), K0=[2, 4, 6, 8, 10], M0=[1, 2, 3, 4, 5], p_min=None, _simple=_fake_simple)
    assert (seq["mM"], seq["mK"]) == (3, 6) and seq["result"]["tag"] == (3, 6)
    assert np.isnan(seq["BICs"][3][10]) and not np.isnan(seq["BICs"][3][8])     # early stop along K (:169,180)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 13
    ps = [ctx.Process(target=_km_worker, args=(r, 2, port, q)) for r in range(2)]
    for p_ in ps:
        p_.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p_ in ps:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    owners = 0
    for rank, bics, mK, mM, tag in got:
        assert (mM, mK) == (3, 6)
        for m in seq["BICs"]:
            for k in seq["BICs"][m]:
                a, b_ = bics[m][k], seq["BICs"][m][k]
                assert (np.isnan(a) and np.isnan(b_)) or a == b_
        owners += tag == (3, 6)
    assert owners == 1                                           # M = 3 was dealt to exactly one rank
