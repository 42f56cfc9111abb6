"""CPU tests of the multi-threaded C / BLAS build of the oracle (oracle/simples_oracle_c.c + oracle/accel.py), the CPU
baseline of bench.py.  It is held to the NumPy oracle statement by statement: special functions, one MC sweep
(R/EM_impute.R:173-198), the E-pass in its fresh and stale-beta2 forms (:126-157, :202-213), the sufficient statistics,
and whole EM_impute runs.  Test infrastructure only -- PARITY UNPINNED like the oracle itself."""
import ctypes

import numpy as np
import pytest
from scipy import special

import oracle as O
from oracle import simples_oracle as SO
from oracle.accel import Accel, load
from util import assert_em_close, em_args, make_case, rel

_dp = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def acc():
    return Accel()


def _vec(lib, name, *arrs):
    out = np.empty_like(arrs[0])
    getattr(lib, name)(*[a.ctypes.data_as(_dp) for a in arrs], arrs[0].size, out.ctypes.data_as(_dp))
    return out


def test_library_exports_and_threads():
    lib = load()
    for s in ("so_abi_version", "so_max_threads", "so_sweep_entries", "so_sq_over_sigma", "so_rowsq_weighted",
              "so_qnorm", "so_pnorm", "so_trunc_z"):
        assert hasattr(lib, s), s
    assert lib.so_abi_version() == 1 and lib.so_max_threads() >= 1


def test_special_functions_match_scipy():
    lib = load()
    rng = np.random.default_rng(0)
    p = np.concatenate([rng.uniform(0, 1, 20000), 10.0 ** rng.uniform(-300, -1, 2000), 1 - 10.0 ** rng.uniform(-15, -1, 2000),
                        [0.075, 0.925, 0.5, 2.0 ** -53]])
    q = _vec(lib, "so_qnorm_v", p)
    want = special.ndtri(p)
    assert np.max(np.abs(q - want) / np.maximum(1.0, np.abs(want))) <= 5e-15
    x = np.concatenate([rng.normal(0, 3, 20000), rng.uniform(-36, 8, 5000)])
    pn = _vec(lib, "so_pnorm_v", x)
    assert np.max(np.abs(pn - special.ndtr(x)) / special.ndtr(x)) <= 1e-13
    # truncated draws, incl. the log-space branch below r = -30 (oracle.trunc_z)
    r = np.concatenate([rng.normal(0, 2, 5000), rng.uniform(-30, -5, 2000), rng.uniform(-200, -30, 2000)])
    u = rng.uniform(0, 1, r.size)
    tz = _vec(lib, "so_trunc_z_v", u, r)
    want = SO.trunc_z(u, r)
    assert np.max(np.abs(tz - want) / np.maximum(1.0, np.abs(want))) <= 2e-12
    assert np.all(tz <= r)


@pytest.mark.parametrize("M0,MB,clip", [(3, None, False), (1, None, True), (4, 3, True)])
def test_sweep_matches_numpy(acc, M0, MB, clip):
    st = make_case(150, 96, 4, 3, 5, M0, seed=11 + M0, MB=MB)
    Yc, gm, gs, mu = SO._prologue(st["Y"], st["z"], None, M0, -np.inf, np.inf, O.COMPAT_DEFAULT)
    pg = st["pg"] if st["pg"].ndim == 2 else st["pg"][:, None]
    ct0 = np.asarray(st["celltype"], dtype=np.int64) - 1
    mask = st["Y0"] <= 0.1
    gs = gs * np.random.default_rng(1).uniform(0.5, 2.0, gs.shape)          # do_impute passes pos_sd != 1
    _, _, _, Ms, _, _ = SO.epass(Yc, st["beta"], st["sigma"], st["lam"], mu, st["pi"])
    Wm = SO.factor_means(Yc, st["beta"], st["sigma"], mu, Ms)
    kw = dict(cell0=2 ** 33 + 7, draw_membership=M0 > 1, lower=0.05 if clip else -np.inf, upper=3.0 if clip else np.inf,
              clip=True)
    Ya, Yb = Yc.copy(order="F"), Yc.copy(order="F")
    ia, fa = SO.sample_sweep(Ya, mask, st["z"], Wm, Ms, mu, st["beta"], st["sigma"], pg, ct0, gm, gs, 0.1, 5, 3, **kw)
    ib, fb = acc.sample_sweep(Yb, mask, st["z"], Wm, Ms, mu, st["beta"], st["sigma"], pg, ct0, gm, gs, 0.1, 5, 3, **kw)
    assert np.array_equal(ia, ib) and rel(fb, fa) <= 1e-15
    assert np.array_equal(Ya[~mask], Yb[~mask])                              # observed entries untouched
    assert np.abs(Ya - Yb).max() <= 1e-11 * max(1.0, np.abs(Ya).max())
    # mode = "expect" goes through the NumPy statement
    Ye = Yc.copy(order="F")
    acc.sample_sweep(Ye, mask, st["z"], Wm, Ms, mu, st["beta"], st["sigma"], pg, ct0, gm, gs, 0.1, 5, 3, expect=True)
    assert np.isfinite(Ye).all()


@pytest.mark.parametrize("nonshared", [False, True])
def test_epass_and_stats_match_numpy(acc, nonshared):
    M0, K0 = 4, 6
    st = make_case(220, 128, 4, 3, K0, M0, seed=5, nonshared=nonshared)
    Yc, gm, gs, mu = SO._prologue(st["Y"], st["z"], None, M0, -np.inf, np.inf, O.COMPAT_DEFAULT)
    a = (Yc, st["beta"], st["sigma"], st["lam"], mu, st["pi"])
    ll, lik, sc, Ms, b2s, dts = SO.epass(*a)
    ll2, lik2, sc2, Ms2, b2s2, dts2 = acc.epass(*a)
    assert rel(ll2, ll) <= 1e-12 and rel(lik2, lik) <= 1e-10 and rel(sc2, sc) <= 1e-12
    Wm, Wm2 = SO.factor_means(Yc, st["beta"], st["sigma"], mu, Ms), acc.factor_means(Yc, st["beta"], st["sigma"], mu, Ms2)
    assert max(rel(x, y) for x, y in zip(Wm2, Wm)) <= 1e-11
    # MC-loop pass: stale beta2 / dt of cluster M0 (Q1, R/EM_impute.R:202-213)
    s = SO.epass(*a, M_list=Ms, beta2_fix=b2s[M0 - 1], dt_fix=dts[M0 - 1])
    s2 = acc.epass(*a, M_list=Ms, beta2_fix=b2s[M0 - 1], dt_fix=dts[M0 - 1])
    assert rel(s2[0], s[0]) <= 1e-12
    assert max(rel(x, y) for x, y in zip(acc.factor_means(Yc, st["beta"], st["sigma"], mu, Ms), Wm)) <= 1e-11
    # cluster constants the accelerated pass was not run for: the NumPy formula
    assert max(rel(x, y) for x, y in zip(acc.factor_means(Yc, st["beta"], st["sigma"], mu, list(Ms)), Wm)) <= 1e-11
    z = lik / lik.sum(axis=1)[:, None]
    shared = not nonshared
    A, B = SO.local_stats(Yc, z, Wm, shared), acc.local_stats(Yc, z, Wm, shared)
    assert A.keys() == B.keys()
    for k in A:
        assert rel(B[k], A[k]) <= 1e-12, k
    assert rel(acc.local_stats(Yc[:, 40:150], z[40:150], [w[40:150] for w in Wm], shared)["U"],
               SO.local_stats(Yc[:, 40:150], z[40:150], [w[40:150] for w in Wm], shared)["U"]) <= 1e-12


@pytest.mark.parametrize("n,G,K0,M0,MB", [(256, 320, 6, 3, None), (500, 200, 10, 1, None), (240, 160, 5, 4, 3)])
def test_em_impute_accel_matches_numpy(acc, n, G, K0, M0, MB):
    st = make_case(n, G, 4, 3, K0, M0, seed=2, MB=MB)
    kw = dict(celltype=st["celltype"], impt_it=1, est_z=1, est_lam=1, seed=3)
    ref = O.em_impute_fast(*em_args(st, M0, K0, 4), **kw)
    got = O.em_impute_fast(*em_args(st, M0, K0, 4), accel=acc, **kw)
    assert_em_close(got, ref, M0, tol_param=1e-9, tol_y=1e-9)
    # sharded statistics (the all-reduce emulation) through the accelerated loops
    got2 = O.em_impute_fast(*em_args(st, M0, K0, 4), accel=acc, shards=[0, n // 3, n], **kw)
    assert_em_close(got2, ref, M0, tol_param=1e-9, tol_y=1e-9)
