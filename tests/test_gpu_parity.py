"""GPU parity tests (run on a B200 with -m gpu): every call goes through the C ABI
(simples_b200.EM_impute / do_impute -> libsimples_b200.so) and is compared with the CPU oracle on the
same seeded inputs and the same Philox variate tape.  Tolerances are north_star's: <= 1e-8 relative
on parameters / log-likelihood, <= 1e-6 relative on imputed entries, identical cluster assignments."""
import os

import numpy as np
import pytest

import oracle as O
from util import assert_em_close, em_args, make_case, rel

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def sb(built_lib):
    import simples_b200
    simples_b200.init(0)
    return simples_b200


CASES = [
    # n,   G,   K, MC, K0, M0, iters, impt_it, nonshared
    (200, 150, 3, 3, 4, 3, 1, 5, False),     # one deterministic iteration
    (200, 150, 3, 3, 4, 3, 5, 9, False),     # deterministic EM (impt_it >= iter never samples, R/EM_impute.R:161)
    (200, 150, 3, 3, 4, 3, 4, 1, False),     # Monte-Carlo EM
    (200, 150, 3, 3, 4, 3, 3, 1, True),      # user-supplied per-cluster sigma (general path at it = 1)
    (333, 257, 3, 2, 3, 2, 4, 2, False),     # ragged: G % 64 != 0, n % 128 != 0
    (130, 64, 1, 1, 2, 1, 3, 1, False),      # M0 = 1 (no membership draw, R/EM_impute.R:165-166)
    (900, 520, 6, 5, 12, 7, 3, 1, False),    # K0 + M0 > 16
    (64, 40, 1, 2, 1, 2, 2, 1, False),       # K0 = 1, tiny
    (1500, 1000, 6, 3, 10, 3, 3, 1, False),  # config C1 shape (vignette simulation, K0 = 10, M0 = 3)
    (1200, 700, 8, 6, 20, 20, 3, 1, False),  # config C4 factor / cluster counts (K0 = M0 = 20)
    (900, 400, 8, 5, 30, 20, 2, 1, False),   # config C5 maximum (K0 = 30, M0 = 20): M_m tables read from L2
    (700, 300, 4, 4, 32, 32, 2, 1, True),    # ABI limits K0 = M0 = 32, per-cluster sigma
]


@pytest.mark.parametrize("n,G,K,MC,K0,M0,iters,impt_it,nonshared", CASES)
def test_em_impute_matches_oracle(sb, n, G, K, MC, K0, M0, iters, impt_it, nonshared):
    st = make_case(n, G, K, MC, K0, M0, seed=3, nonshared=nonshared)
    kw = dict(celltype=st["celltype"], impt_it=impt_it, est_z=1, est_lam=1, seed=7)
    ref = O.em_impute_fast(*em_args(st, M0, K0, iters), **kw)
    got = sb.EM_impute(*em_args(st, M0, K0, iters), **kw)
    assert_em_close(got, ref, M0)


def test_em_matches_literal_reference_formulation(sb):
    """Against the literal restatement (augmented design + dense bigM), not the sufficient statistics."""
    K0, M0 = 4, 3
    st = make_case(120, 100, 3, 3, K0, M0, seed=12)
    kw = dict(celltype=st["celltype"], impt_it=2, est_z=1, est_lam=1, seed=5)
    ref = O.em_impute_ref(*em_args(st, M0, K0, 4), **kw)
    got = sb.EM_impute(*em_args(st, M0, K0, 4), **kw)
    assert_em_close(got, ref, M0)


def test_em_reference_defaults_and_options(sb):
    """est_z = 2 / est_lam = 2 defaults (R/EM_impute.R:76), z = NULL, explicit mu, max_lambda = F,
    finite clip bounds, celltype with MB != M0 (SIMPLE_B style pg, R/SIMPLE_B.R:318-321)."""
    K0, M0 = 5, 3
    st = make_case(300, 200, 3, 3, K0, M0, seed=4, MB=4)
    a = em_args(st, M0, K0, 4)
    for kw in (dict(celltype=st["celltype"], impt_it=2, seed=1),                                   # defaults
               dict(celltype=st["celltype"], impt_it=1, max_lambda=False, seed=2, penl=0.3, sigma0=10, pi_alpha=2),
               dict(celltype=st["celltype"], impt_it=1, est_z=1, lower=-0.5, upper=4.0, num_mc=2, seed=3),
               dict(celltype=None, impt_it=1, est_z=3, est_lam=3, seed=4)):
        if kw["celltype"] is None:
            st2 = dict(st, pg=st["pg"][:, :1])
            a2 = em_args(st2, M0, K0, 4)
        else:
            a2 = a
        ref = O.em_impute_fast(*a2, **kw)
        got = sb.EM_impute(*a2, **kw)
        assert_em_close(got, ref, M0)
    # z = NULL => mu = 0 and z computed at it = 1 even though est_z = 2 (R/EM_impute.R:93-99,148)
    a3 = a[:11] + (None,)
    kw = dict(celltype=st["celltype"], impt_it=1, seed=5)
    assert_em_close(sb.EM_impute(*a3, **kw), O.em_impute_fast(*a3, **kw), M0)
    mu0 = np.random.default_rng(0).standard_normal((st["Y"].shape[0], M0)) * 0.1
    kw = dict(celltype=st["celltype"], impt_it=1, est_z=1, mu=mu0, seed=6)
    assert_em_close(sb.EM_impute(*a, **kw), O.em_impute_fast(*a, **kw), M0)


def test_compat_flags(sb):
    K0, M0 = 3, 3
    st = make_case(150, 100, 3, 3, K0, M0, seed=6)
    for flags in (O.COMPAT_DEFAULT & ~O.COMPAT_STALE_DT, O.COMPAT_DEFAULT & ~O.COMPAT_PI,
                  O.COMPAT_DEFAULT & ~O.COMPAT_PRIORB_PREV, O.COMPAT_DEFAULT & ~O.COMPAT_MU_DIV_N, 0):
        kw = dict(celltype=st["celltype"], impt_it=1, est_z=1, est_lam=1, seed=1)
        ref = O.em_impute_fast(*em_args(st, M0, K0, 4), compat=flags, **kw)
        got = sb.EM_impute(*em_args(st, M0, K0, 4), ref_compat=flags, **kw)
        assert_em_close(got, ref, M0)


def test_resident_context_equals_one_shot(sb):
    K0, M0 = 4, 3
    st = make_case(260, 190, 3, 3, K0, M0, seed=9)
    kw = dict(celltype=st["celltype"], impt_it=1, est_z=1, est_lam=1, seed=11)
    one = sb.EM_impute(*em_args(st, M0, K0, 5), **kw)
    with sb.EMContext(st["Y"], st["Y0"], st["pg"], M0, K0, 0.1, st["beta"], st["sigma"], st["lam"], st["pi"], st["z"],
                      iter=2, **kw) as ctx:
        ctx.run(2)
        ctx.run(3)                    # grows the loglik array, continues the tape
        two = ctx.fetch()
        assert ctx.nnz0 == int((st["Y0"] <= 0.1).sum())
        assert ctx.kernel_launches > 0 and len(ctx.iter_ms()) == 3
    for k in ("loglik", "beta", "sigma", "mu", "lambda", "pi", "z", "Y"):
        assert np.array_equal(one[k], two[k]), k            # bit-identical: deterministic reductions


def test_run_to_run_determinism(sb):
    K0, M0 = 6, 4
    st = make_case(700, 300, 4, 4, K0, M0, seed=10)
    kw = dict(celltype=st["celltype"], impt_it=1, est_z=1, est_lam=1, seed=3)
    a = sb.EM_impute(*em_args(st, M0, K0, 3), **kw)
    b = sb.EM_impute(*em_args(st, M0, K0, 3), **kw)
    for k in ("loglik", "beta", "sigma", "mu", "lambda", "z", "Y"):
        assert np.array_equal(a[k], b[k]), k
    c = sb.EM_impute(*em_args(st, M0, K0, 3), **dict(kw, seed=4))
    assert not np.array_equal(a["Y"], c["Y"])


def test_do_impute_matches_oracle(sb):
    K0, M0 = 5, 3
    st = make_case(240, 180, 3, 3, K0, M0, seed=13, MB=2)
    em = O.em_impute_fast(*em_args(st, M0, K0, 3), celltype=st["celltype"], impt_it=1, est_z=1, est_lam=1, seed=2)
    G = st["Y"].shape[0]
    sd = np.random.default_rng(1).uniform(0.7, 1.4, G)
    for pos_sd, pg in ((em["geneSd"], st["pg"]), (sd, 0.5)):
        kw = dict(mcmc=5, burnin=2, pg=pg, cutoff=0.1, seed=21)
        args = (st["Y0"], em["Y"], em["beta"], em["lambda"], em["sigma"], em["mu"], em["pi"], em["geneM"], pos_sd,
                st["celltype"])
        ref = O.do_impute_ref(*args, **kw)
        got = sb.do_impute(*args, **kw)
        assert abs(got["loglik"] - ref["loglik"]) <= 1e-8 * abs(ref["loglik"])
        assert rel(got["impt"], ref["impt"]) <= 1e-6
        assert np.abs(got["impt_var"] - ref["impt_var"]).max() <= 1e-6 * max(1.0, np.abs(ref["impt_var"]).max())
        assert rel(got["EF"], ref["EF"]) <= 1e-7 and rel(got["varF"], ref["varF"]) <= 1e-7
        assert rel(got["consensus_cluster"], ref["consensus_cluster"]) <= 1e-8
    nocons = sb.do_impute(*args, consensus=False, **kw)
    assert nocons["consensus_cluster"] is None and rel(nocons["impt"], ref["impt"]) <= 1e-6


def test_golden_fixture(sb):
    g = np.load(os.path.join(GOLD, "em_small.npz"))
    K0, M0 = int(g["K0"]), int(g["M0"])
    got = sb.EM_impute(g["Y"], g["Y0"], g["pg"], M0, K0, 0.1, int(g["iter"]), g["beta"], g["sigma"], g["lam"], g["pi"],
                       g["z"], celltype=g["celltype"], impt_it=1, est_z=1, est_lam=1, seed=int(g["seed"]))
    for k in ("loglik", "pi", "mu", "sigma", "beta", "lambda", "z"):
        assert rel(got[k], g["out_" + k]) <= 1e-8, k
    assert rel(got["Y"], g["out_Y"]) <= 1e-6
    assert rel(np.stack(got["Ef"]), g["out_Ef"]) <= 1e-7 and rel(np.stack(got["Varf"]), g["out_Varf"]) <= 1e-8
    mi = sb.do_impute(g["Y0"], g["out_Y"], g["out_beta"], g["out_lambda"], g["out_sigma"], g["out_mu"], g["out_pi"],
                      g["out_geneM"], np.ones(g["Y"].shape[0]), g["celltype"], mcmc=4, burnin=2, pg=g["pg"],
                      seed=int(g["seed"]) + 1)
    assert rel(mi["impt"], g["mi_impt"]) <= 1e-6 and abs(mi["loglik"] - float(g["mi_loglik"])) <= 1e-8 * abs(float(g["mi_loglik"]))


def test_non_finite_input_is_rejected(sb):
    K0, M0 = 3, 2
    st = make_case(80, 60, 2, 2, K0, M0, seed=3)
    Y = np.array(st["Y"]); Y[7, 11] = np.nan
    with pytest.raises(sb.SimplesError) as e:
        sb.EM_impute(Y, st["Y0"], st["pg"], M0, K0, 0.1, 1, st["beta"], st["sigma"], st["lam"], st["pi"], st["z"])
    assert e.value.code == -1 and "non-finite" in str(e.value)


def test_error_paths(sb):
    K0, M0 = 3, 2
    st = make_case(64, 40, 1, 2, K0, M0, seed=1)
    bad_ct = st["celltype"].copy()
    bad_ct[0] = 9
    with pytest.raises(sb.SimplesError) as e:
        sb.EM_impute(*em_args(st, M0, K0, 1), celltype=bad_ct)
    assert e.value.code == -1 and "celltype" in str(e.value)


def test_full_size_properties(sb):
    """BASELINE.json configs[2] (100k cells x 2k genes, K0 = M0 = 10): size-independent properties --
    observed entries untouched, imputed entries of non-recorded dropouts finite, z rows sum to 1,
    sigma within the clip, finite log-likelihood, and a 4k-cell slice of the first deterministic
    iteration agreeing with the oracle run on that slice (cells are conditionally independent)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from simples_b200 import synth
    n, G, K0, M0 = 100_000, 2000, 10, 10
    gp = synth.gene_params(G, 10, 10, seed=1)
    Y2, Z = synth.simulate_shard(gp, n, shard=0, seed=1)
    st = synth.init_state(gp, Y2, Z, K0, M0, seed=1)
    kw = dict(celltype=st["celltype"], est_z=1, est_lam=1, impt_it=1, num_mc=3, seed=99)
    with sb.EMContext(st["Y"], st["Y0"], st["pg"], M0, K0, 0.1, st["beta"], st["sigma"], st["lam"], st["pi"], st["z"],
                      iter=3, **kw) as ctx:
        ctx.run(3)
        out = ctx.fetch()
    assert np.all(np.isfinite(out["loglik"])) and np.all(np.isfinite(out["beta"]))
    obs = st["Y0"] > 0.1
    assert np.allclose(out["Y"][obs], st["Y0"][obs], rtol=0, atol=1e-9)
    assert np.all(np.isfinite(out["Y"]))
    assert np.allclose(out["z"].sum(axis=1), 1.0, atol=1e-12)
    assert out["sigma"].min() >= 1e-4 and out["sigma"].max() <= 9.0 and np.all(out["sigma"] == out["sigma"][:, :1])
    assert abs(out["pi"].sum() - 1.0) < 1e-12 and np.allclose(out["lambda"].sum(axis=0), 1.0, atol=1e-9)
    # E-pass on a slice == oracle E-pass on that slice with the returned (pre-M-step ... post) parameters:
    # re-run one deterministic iteration on 4k cells starting from the fitted parameters.
    sl = slice(0, 4000)
    a = (np.asfortranarray(out["Y"][:, sl]), np.asfortranarray(st["Y0"][:, sl]), st["pg"], M0, K0, 0.1, 1, out["beta"],
         out["sigma"], out["lambda"], out["pi"], out["z"][sl])
    kw1 = dict(celltype=st["celltype"][sl], est_z=1, est_lam=1, impt_it=5, mu=out["mu"])
    assert_em_close(sb.EM_impute(*a, **kw1), O.em_impute_fast(*a, **kw1), M0)


def test_yimp0_and_bic(sb):
    """Yimp0 and BIC / BIC0 of the drivers (R/SIMPLE.R:421-430) from the resident EM state."""
    K0, M0 = 5, 3
    st = make_case(210, 170, 3, 3, K0, M0, seed=14)
    kw = dict(celltype=st["celltype"], impt_it=1, est_z=1, est_lam=1, seed=4)
    ref = O.em_impute_fast(*em_args(st, M0, K0, 3), **kw)
    with sb.EMContext(st["Y"], st["Y0"], st["pg"], M0, K0, 0.1, st["beta"], st["sigma"], st["lam"], st["pi"], st["z"],
                      iter=3, **kw) as ctx:
        ctx.run(3)
        got = ctx.yimp0()
        ll = ctx.fetch(big=False)["loglik"]
    assert rel(got, O.yimp0_ref(ref)) <= 1e-8
    G, n = st["Y"].shape
    b, b0 = sb.bic(ll[-1], n, G, K0, M0)
    rb, rb0 = O.bic_ref(ref["loglik"][-1], n, G, K0, M0)
    assert abs(b - rb) <= 1e-10 * abs(rb) and abs(b0 - rb0) <= 1e-10 * abs(rb0)


def test_c4_shape_slice(sb):
    """BASELINE.json configs[3] shape (G = 5000 genes, K0 = M0 = 20) on a slice of cells: exercises ldG = 5056
    (79 gene tiles), NQP = 40, NVP = 48, NFP = 48 template variants against the oracle."""
    K0, M0 = 20, 20
    st = make_case(2500, 5000, 20, 20, K0, M0, seed=15)
    kw = dict(celltype=st["celltype"], impt_it=1, est_z=1, est_lam=1, seed=8, num_mc=2)
    ref = O.em_impute_fast(*em_args(st, M0, K0, 3), **kw)
    got = sb.EM_impute(*em_args(st, M0, K0, 3), **kw)
    assert_em_close(got, ref, M0)


class _SoloShard:
    """Single-rank stand-in for simples_b200.dist.CellShard: exercises n_total / cell_offset / the
    all-reduce callback plumbing of the C ABI without a process group."""

    def __init__(self, n_total, cell_offset):
        self.n_total, self.cell_offset, self.calls = n_total, cell_offset, 0

    def allreduce_sum_(self, ptr, count, stream=None):
        self.calls += 1          # sum over one rank = identity


def test_global_cell_index_64bit_and_seed_hi(sb):
    """The Philox counter carries the GLOBAL cell index (64 bit) and the key the full 64-bit seed."""
    K0, M0 = 4, 3
    st = make_case(200, 150, 3, 3, K0, M0, seed=16)
    off = 2**33 + 7
    seed = (0xA5A5A5A5 << 32) | 0x1234567
    kw = dict(celltype=st["celltype"], impt_it=1, est_z=1, est_lam=1, seed=seed)
    sh = _SoloShard(st["Y"].shape[1], off)
    got = sb.EM_impute(*em_args(st, M0, K0, 3), comm=sh, **kw)
    ref = O.em_impute_fast(*em_args(st, M0, K0, 3), cell0=off, **kw)
    assert_em_close(got, ref, M0)
    assert sh.calls == 1 + 2 + 3    # input-validation agreement, gene means, default mu, one per EM iteration
    other = O.em_impute_fast(*em_args(st, M0, K0, 3), cell0=7, **kw)
    assert rel(other["Y"], ref["Y"]) > 1e-3       # the high word really changes the stream


def test_extreme_tails_and_degenerate_dropout_prior(sb):
    """pg = 1 (no dropout possible: every zero is a truncated-normal draw), pg = 0.99, and genes whose mean sits
    40-80 sd above the cutoff (r << -30: log-space Newton branch of the truncated quantile, cold pnorm tails)."""
    K0, M0 = 4, 3
    st = make_case(260, 160, 3, 3, K0, M0, seed=18)
    pg = st["pg"].copy()
    pg[:6] = 1.0
    pg[6:12] = 0.99
    pg[12:18] = 1e-3
    Y = np.array(st["Y"])
    Y[:4] += 25.0            # huge expression with small noise => (cutoff - ms) / sds << -30 at the zero entries
    Y[20:24] -= 30.0         # far below the cutoff => r >> 8
    st = dict(st, pg=pg, Y=np.asfortranarray(Y), sigma=np.full_like(st["sigma"], 0.3))
    kw = dict(celltype=st["celltype"], impt_it=1, est_z=1, est_lam=1, seed=21, num_mc=2)
    ref = O.em_impute_fast(*em_args(st, M0, K0, 3), **kw)
    got = sb.EM_impute(*em_args(st, M0, K0, 3), **kw)
    assert np.all(np.isfinite(got["Y"])) and np.all(np.isfinite(ref["Y"]))
    assert_em_close(got, ref, M0)


def _lq_case(K0, M0, n, G, seed, MB=None):
    st = make_case(n, G, 3, 3, K0, M0, seed=seed, MB=MB)
    em = O.em_impute_fast(*em_args(st, M0, K0, 3), celltype=st["celltype"], impt_it=1, est_z=1, est_lam=1, seed=2)
    rng = np.random.default_rng(seed)
    G = st["Y"].shape[0]
    lq = np.sort(rng.choice(G, size=G // 3, replace=False))
    return st, em, lq, rng.uniform(0.4, 1.6, (len(lq), M0))


@pytest.mark.parametrize("K0,M0,n,G,MB", [(4, 3, 150, 120, None), (1, 1, 90, 80, None), (7, 2, 333, 210, 2), (32, 4, 260, 150, None)])
def test_lq_fit_matches_oracle(sb, K0, M0, n, G, MB):
    """Row N1: lq-gene regression + imputation draw between the two EM calls (R/SIMPLE.R:305-411)."""
    st, em, lq, sig = _lq_case(K0, M0, n, G, seed=31 + K0, MB=MB)
    dat = np.asarray(st["Y0"])[lq]
    init = np.asarray(st["Y"])[lq]
    for mode, resp in ((0, None), (1, None), (1, init)):
        got = sb.lq_fit(dat, sig, st["pg"][lq], em["z"], em["Ef"], em["Varf"], st["celltype"], penl=1, cutoff=0.1,
                        resp=resp, resid_mode=mode, seed=17, sweep=1000)
        ref = O.lq_regress_ref(dat, dat if resp is None else resp, sig, em["z"], em["Ef"], em["Varf"], 1, mode)
        for k, r in zip(("mu", "beta", "sigma"), ref):
            assert rel(got[k], r) <= 1e-8, (mode, k, rel(got[k], r))
        assert np.array_equal(got["beta"] == 0, ref[1] == 0)
        Yq = O.lq_impute_ref(dat, ref[0], ref[1], ref[2], st["pg"][lq], em["z"], em["Ef"], st["celltype"], 0.1, 17, 1000)
        assert rel(got["Y"], Yq) <= 1e-6
        obs = dat > 0.1
        assert np.array_equal(got["Y"][obs], dat[obs])
    only = sb.lq_fit(dat, sig, st["pg"][lq], em["z"], em["Ef"], em["Varf"], st["celltype"], impute=False)
    assert only["Y"] is None and rel(only["beta"], O.lq_regress_ref(dat, dat, sig, em["z"], em["Ef"], em["Varf"], 1, 0)[1]) <= 1e-8


def test_lq_fit_sharded_equals_single(sb):
    """Cell shards passed one after another through a summing callback reproduce the unsharded call."""
    import torch
    K0, M0 = 4, 3
    st, em, lq, sig = _lq_case(K0, M0, 200, 120, seed=77)
    dat = np.asarray(st["Y0"])[lq]
    full = sb.lq_fit(dat, sig, st["pg"][lq], em["z"], em["Ef"], em["Varf"], st["celltype"], resid_mode=1, seed=5, sweep=9)
    n = dat.shape[1]
    cuts = [0, 72, n]

    class Rec:      # pass 1 records each shard's statistics; pass 2 replaces them by the global sum
        def __init__(self, n_total, off):
            self.n_total, self.cell_offset, self.seen, self.total = n_total, off, [], None
        def allreduce_sum_(self, ptr, count, stream):
            buf = _dev_view(ptr, count, stream)
            if self.total is None:
                self.seen.append(buf.clone())
            else:
                buf.copy_(self.total[len(self.seen)])
                self.seen.append(None)
            torch.cuda.synchronize()

    def _dev_view(ptr, count, stream):
        class W:
            __cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3}
        torch.cuda.synchronize()
        return torch.as_tensor(W(), device="cuda")

    def run(comm, lo, hi):
        return sb.lq_fit(dat[:, lo:hi], sig, st["pg"][lq], em["z"][lo:hi], [e[lo:hi] for e in em["Ef"]], em["Varf"],
                         st["celltype"][lo:hi], resid_mode=1, seed=5, sweep=9, comm=comm)

    recs = [Rec(n, cuts[r]) for r in range(2)]
    for r in range(2):
        run(recs[r], cuts[r], cuts[r + 1])
    totals = [recs[0].seen[j] + recs[1].seen[j] for j in range(len(recs[0].seen))]
    outs = []
    for r in range(2):
        c = Rec(n, cuts[r])
        c.total = totals
        outs.append(run(c, cuts[r], cuts[r + 1]))
    for k in ("mu", "beta", "sigma"):
        assert rel(outs[0][k], full[k]) <= 1e-10 and np.array_equal(outs[0][k], outs[1][k])
    assert rel(np.concatenate([outs[0]["Y"], outs[1]["Y"]], axis=1), full["Y"]) <= 1e-9


def _init_data(seed, n=260, G=170, M0=3):
    sim = O.simulate(n=n, G=G, K=3, MC=M0, S0=10, seed=seed)
    Y2 = np.array(sim["Y2"])
    # rows exercising the NA / refit rules: no positives, one positive, three positives, constant positives
    Y2[0] = 0.0
    Y2[1] = 0.0; Y2[1, 5] = 2.0
    Y2[2] = 0.0; Y2[2, [3, 100, 200]] = (1.0, 2.5, 1.7)
    Y2[3] = 1.5
    return Y2, sim["Z"]


def _draw_mismatch(got, ref, tol=1e-6):
    bad = np.abs(got - ref) > tol * np.maximum(1.0, np.abs(ref))
    return int(bad.sum())


def test_init_impute_matches_oracle(sb):
    """Row N2: ztruncnorm fit + init_impute draw (R/fit_ztrunc.R:55-134, R/SIMPLE.R:15-73).  The fit is a Brent
    minimisation with the reference's tolerance sqrt(eps) ~ 1.5e-8 on sigma, so parameters agree to ~1e-6, not 1e-8."""
    Y2, clus = _init_data(5)
    for p_min in (0.4, 0.6):
        imp, pg, mu, sd = sb.init_impute(Y2, 3, clus, p_min, 0.1, seed=9, sweep=77, details=True)
        rimp, rpg, rmu, rsd = O.init_impute_ref(Y2, 3, clus, p_min, 0.1, seed=9, sweep=77)
        assert rel(sd, rsd) <= 1e-6 and rel(mu, rmu) <= 1e-5 and rel(pg, rpg) <= 1e-6
        assert np.array_equal(pg == 0.99, rpg == 0.99) and np.array_equal(pg == p_min, rpg == p_min)
        obs = Y2 > 0.1
        assert np.array_equal(imp[obs], Y2[obs]) and np.isfinite(imp).all()
        # a Bernoulli decision can flip where the fitted probabilities differ by ~1e-7: allow a handful of entries
        assert _draw_mismatch(imp, rimp, 1e-5) <= 3
    one = sb.init_impute(Y2, 1, np.ones(Y2.shape[1], dtype=int), 0.4, 0.1, seed=9)       # R/SIMPLE.R:233,277
    rone = O.init_impute_ref(Y2, 1, np.ones(Y2.shape[1], dtype=int), 0.4, 0.1, seed=9)
    assert rel(one[1], rone[1]) <= 1e-6 and _draw_mismatch(one[0], rone[0], 1e-5) <= 3


def test_init_impute_bulk_matches_oracle(sb):
    Y2, clus = _init_data(6)
    pg1 = np.clip(np.random.default_rng(0).uniform(0.05, 1.2, (Y2.shape[0], 3)), 0.1, 1.0)      # R/SIMPLE_B.R:284-285
    imp, pg, mu, sd = sb.init_impute_bulk(Y2, clus, None, pg1, 0.1, seed=4, sweep=3, details=True)
    rimp = O.init_impute_bulk_ref(Y2, clus, pg1, 0.1, seed=4, sweep=3)
    assert np.isfinite(imp).all() and np.array_equal(imp[Y2 > 0.1], Y2[Y2 > 0.1])
    assert _draw_mismatch(imp, rimp, 1e-5) <= 3
    assert (pg <= pg1 + 1e-15).all() and pg.min() >= 0.01
    # the structural-zero component N(-1, 0.5) is used: some imputed values far below any fitted mean - 4 sd
    assert (imp[Y2 <= 0.1] < 0).any()


def test_init_impute_sharded(sb):
    """Statistics are cell-additive: two shards summed == one call (draws keyed by the global cell index)."""
    import torch
    Y2, clus = _init_data(7)
    full = sb.init_impute(Y2, 3, clus, 0.4, 0.1, seed=3, details=True)
    n = Y2.shape[1]
    cuts = [0, 101, n]

    def view(ptr, count):
        class W:
            __cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3}
        torch.cuda.synchronize()
        return torch.as_tensor(W(), device="cuda")

    class Rec:
        def __init__(self, off, total=None):
            self.n_total, self.cell_offset, self.seen, self.total = n, off, [], total
        def allreduce_sum_(self, ptr, count, stream):
            buf = view(ptr, count)
            if self.total is None:
                self.seen.append(buf.clone())
            else:
                buf.copy_(self.total)
            torch.cuda.synchronize()

    recs = [Rec(cuts[r]) for r in range(2)]
    for r in range(2):
        sb.init_impute(Y2[:, cuts[r]:cuts[r + 1]], 3, clus[cuts[r]:cuts[r + 1]], 0.4, 0.1, seed=3, comm=recs[r])
    total = recs[0].seen[0] + recs[1].seen[0]
    outs = [sb.init_impute(Y2[:, cuts[r]:cuts[r + 1]], 3, clus[cuts[r]:cuts[r + 1]], 0.4, 0.1, seed=3,
                           comm=Rec(cuts[r], total), details=True) for r in range(2)]
    assert rel(outs[0][1], full[1]) <= 1e-6 and np.array_equal(outs[0][1], outs[1][1])
    assert _draw_mismatch(np.concatenate([outs[0][0], outs[1][0]], axis=1), full[0], 1e-5) <= 3


def _drv_close(got, ref, tol):
    for k in ("loglik_tot", "pi", "mu", "sigma", "beta", "lambda", "z", "Yimp0", "pg", "impt"):
        assert rel(got[k], ref[k]) <= tol, (k, rel(got[k], ref[k]))
    assert abs(got["BIC"] - ref["BIC"]) <= tol * abs(ref["BIC"]) and abs(got["BIC0"] - ref["BIC0"]) <= tol * abs(ref["BIC0"])
    assert np.array_equal(np.argmax(got["z"], axis=1), np.argmax(ref["z"], axis=1))


def test_simple_driver_matches_oracle(sb):
    """Row N3: SIMPLE() end to end (init fit, EM on hq genes, lq regression + draw, EM on all genes, Yimp0 / BIC,
    multiple imputation) against the oracle's restatement of R/SIMPLE.R:200-455.  The init fit is a Brent search
    (sigma to ~1.5e-8), and 23 EM iterations sit on top of it, so the end-to-end tolerance is 1e-5."""
    sim = O.simulate(n=160, G=120, K=2, MC=2, S0=8, seed=3)
    kw = dict(K0=3, M0=2, iter=3, clus=sim["Z"], min_gene=50, mcmc=3, burnin=1, seed=5)
    got = sb.SIMPLE(sim["Y2"], **kw)
    ref = O.simple_ref(sim["Y2"], **kw)
    assert list(got.keys()) == ["loglik_tot", "priorB", "loglik", "BIC", "BIC0", "pi", "mu", "sigma", "beta", "lambda", "z",
                                "Yimp0", "pg", "initclus", "impt", "impt_var", "Ef", "Varf", "consensus_cluster", "p_min"]
    _drv_close(got, ref, 1e-5)
    assert abs(got["loglik"] - ref["loglik"]) <= 1e-5 * abs(ref["loglik"])
    assert rel(got["impt_var"], ref["impt_var"]) <= 1e-4 and rel(got["Ef"], ref["Ef"]) <= 1e-5
    assert rel(got["consensus_cluster"], ref["consensus_cluster"]) <= 1e-5
    # mcmc = 0 branch (R/SIMPLE.R:447-453), p_min = NULL branch (:229-242), init_imp given (Q16 other branch)
    got0 = sb.SIMPLE(sim["Y2"], **dict(kw, mcmc=0, p_min=None))
    ref0 = O.simple_ref(sim["Y2"], **dict(kw, mcmc=0, p_min=None))
    assert got0["loglik"] is None and got0["impt_var"] is None and got0["consensus_cluster"] is None
    assert abs(got0["p_min"] - ref0["p_min"]) <= 1e-6
    _drv_close(got0, ref0, 1e-5)
    goti = sb.SIMPLE(sim["Y2"], **dict(kw, mcmc=0, init_imp=ref["impt"]))
    refi = O.simple_ref(sim["Y2"], **dict(kw, mcmc=0, init_imp=ref["impt"]))
    _drv_close(goti, refi, 1e-5)
    with pytest.raises(ValueError):
        sb.SIMPLE(sim["Y2"], **dict(kw, min_gene=5000))


def test_simple_b_driver_matches_oracle(sb):
    sim = O.simulate(n=160, G=120, K=2, MC=2, S0=8, seed=4)
    G = sim["Y2"].shape[0]
    bulk = np.stack([sim["Y2"][:, sim["Z"] == m + 1].mean(axis=1) * 1.3 + 0.05 for m in range(2)], axis=1)
    for b in (1, None):
        kw = dict(M0=2, clus=sim["Z"], b=b, iter=3, min_gene=50, mcmc=2, burnin=1, seed=8)
        got = sb.SIMPLE_B(sim["Y2"], 3, bulk, sim["Z"], **kw)
        ref = O.simple_b_ref(sim["Y2"], 3, bulk, sim["Z"], **kw)
        assert "initclus" not in got and "p_min" not in got
        _drv_close(got, ref, 1e-5)


def test_select_km_matches_oracle(sb):
    sim = O.simulate(n=140, G=100, K=2, MC=2, S0=8, seed=6)
    clus2 = sim["Z"]
    kw = dict(iter=2, min_gene=40, mcmc=0, seed=2)
    # the reference passes one `clus` to every M; use M0 = 1 and 2 with labels valid for both
    got = sb.selectKM(sim["Y2"], K0=[2, 3, 4], M0=[1], clus=np.ones_like(clus2), **kw)
    ref = O.select_km_ref(sim["Y2"], [2, 3, 4], [1], clus_of=lambda m: np.ones_like(clus2), **kw)
    assert got["mK"] == ref["mK"] and got["mM"] == ref["mM"] == 1
    for k in (2, 3, 4):
        a, b_ = got["BICs"][1][k], ref["BICs"][1][k]
        assert (np.isnan(a) and np.isnan(b_)) or abs(a - b_) <= 1e-5 * abs(b_)
    assert rel(got["result"]["impt"], ref["result"]["impt"]) <= 1e-5


@pytest.mark.parametrize("keep,pgv", [(0.07, 0.75), (0.0, 0.9), (0.3, 0.5)])
def test_blocks_dense_in_zeros(sb, keep, pgv):
    """Real scRNA-seq matrices are 80-95 % zeros, so a warp's 8 x 32 block is (almost) all zero entries and most of
    them are certain 'expressed but below the cutoff' draws when pg is high: the sweep's work lists fill up
    completely (this case once corrupted entries when the re-filed list ran into the unread one)."""
    K0, M0 = 4, 3
    st = make_case(200, 200, 3, 3, K0, M0, seed=23)
    rng = np.random.default_rng(3)
    st["Y0"] = np.asarray(st["Y0"]) * (rng.uniform(size=st["Y0"].shape) < keep)
    st["pg"] = np.full_like(st["pg"], pgv)
    kw = dict(celltype=st["celltype"], impt_it=1, est_z=1, est_lam=1, seed=6)
    ref = O.em_impute_fast(*em_args(st, M0, K0, 3), **kw)
    got = sb.EM_impute(*em_args(st, M0, K0, 3), **kw)
    assert_em_close(got, ref, M0)
    args = (st["Y0"], ref["Y"], ref["beta"], ref["lambda"], ref["sigma"], ref["mu"], ref["pi"], ref["geneM"], ref["geneSd"],
            st["celltype"])
    mkw = dict(mcmc=3, burnin=1, pg=st["pg"], cutoff=0.1, seed=2)
    r2, g2 = O.do_impute_ref(*args, **mkw), sb.do_impute(*args, **mkw)
    assert rel(g2["impt"], r2["impt"]) <= 1e-6 and abs(g2["loglik"] - r2["loglik"]) <= 1e-8 * abs(r2["loglik"])


def _cluster_data(seed, n=420, G=160, M=3, rank=4):
    rng = np.random.default_rng(seed)
    Z = np.sort(rng.integers(1, M + 1, n))
    cen = rng.normal(size=(G, M)) * 1.5
    B = rng.normal(size=(G, rank)) * np.array([0.6, 0.5, 0.4, 0.3])[:rank]
    X = cen[:, Z - 1] + B @ rng.normal(size=(rank, n)) + 0.7 * rng.normal(size=(G, n))
    X[5] = 2.0                                               # a constant gene: sd = 0
    return X, Z


def test_init_cluster_matches_oracle(sb):
    """Row N4: scale -> truncated SVD -> k-means (R/SIMPLE.R:258-266).  Singular values and the K-dimensional cell
    embedding agree with LAPACK's SVD; Lloyd's restarts from the same Philox-drawn centres give the same partition."""
    X, Z = _cluster_data(4)
    K, M0 = 6, 3
    got = sb.init_cluster(X, K, M0, seed=11, nstart=12, iter_max=25, details=True)
    lab, cm, d, wss = O.init_cluster_ref(X, K, M0, nstart=12, iter_max=25, seed=11)
    assert rel(got["d"], d) <= 1e-8
    # the embedding is defined up to an orthogonal map of the columns: compare the n x n Gram matrices
    assert rel(got["cellmat"] @ got["cellmat"].T, cm @ cm.T) <= 1e-6
    assert O.adjusted_rand_index(got["clus"], lab) >= 0.99 and abs(got["tot_withinss"] - wss) <= 1e-6 * wss
    assert O.adjusted_rand_index(got["clus"], Z) >= 0.95
    assert set(np.unique(got["clus"])) <= set(range(1, M0 + 1))
    one = sb.init_cluster(X, 3, 1, seed=1, nstart=2, iter_max=3)
    assert np.array_equal(one, np.ones(X.shape[1], dtype=np.int32))


def test_simple_driver_with_clustering_start(sb):
    """SIMPLE(clus = NULL): the whole pipeline including the clustering start recovers the simulated cell types."""
    sim = O.simulate(n=300, G=260, K=3, MC=3, S0=12, seed=9)
    res = sb.SIMPLE(sim["Y2"], K0=3, M0=3, iter=3, clus=None, K=5, min_gene=120, mcmc=2, burnin=1, seed=3)
    assert res["initclus"].shape == (300,) and O.adjusted_rand_index(np.argmax(res["z"], axis=1), sim["Z"]) >= 0.8
    assert np.isfinite(res["impt"]).all() and np.isfinite(res["BIC"])


def test_resident_pipeline_equals_host_pipeline(sb):
    """Device-resident run (dat uploaded once, G x n intermediates stay in HBM, device pointers through the C ABI)
    == the same run with host buffers between the stages, bit for bit."""
    import torch
    sim = O.simulate(n=200, G=150, K=2, MC=2, S0=8, seed=12)
    kw = dict(K0=3, M0=2, iter=3, clus=sim["Z"], min_gene=60, mcmc=3, burnin=1, seed=5)
    host = sb.SIMPLE(sim["Y2"], **kw)
    dev = sb.SIMPLE(sim["Y2"], resident=True, **kw)
    for k in ("loglik_tot", "pi", "mu", "sigma", "beta", "lambda", "z", "pg", "Ef", "Varf"):
        assert np.array_equal(np.asarray(host[k]), np.asarray(dev[k])), k
    for k in ("impt", "impt_var", "Yimp0"):
        assert isinstance(dev[k], torch.Tensor) and dev[k].is_cuda
        assert np.array_equal(host[k], dev[k].cpu().numpy()), k
    assert host["BIC"] == dev["BIC"] and host["loglik"] == dev["loglik"]
    bulk = np.stack([sim["Y2"][:, sim["Z"] == m + 1].mean(axis=1) * 1.3 + 0.05 for m in range(2)], axis=1)
    kb = dict(M0=2, clus=None, K=4, iter=2, min_gene=60, mcmc=0, seed=8)
    hb = sb.SIMPLE_B(sim["Y2"], 3, bulk, sim["Z"], **kb)
    db = sb.SIMPLE_B(sim["Y2"], 3, bulk, sim["Z"], resident=True, **kb)
    # rowMeans(dat[, celltype == i]) is summed by torch on the device and by numpy on the host: different orders,
    # so pg differs in the last bit and the runs agree to rounding instead of bit for bit
    assert rel(db["beta"], hb["beta"]) <= 1e-7 and rel(db["impt"].cpu().numpy(), hb["impt"]) <= 1e-7
    # device matrices straight into a stage entry point
    Yd = sb.dev_matrix(sim["Y2"])
    a = sb.init_impute(Yd, 2, sim["Z"], 0.4, 0.1, seed=1, out=sb.dev_empty(*sim["Y2"].shape))
    b = sb.init_impute(sim["Y2"], 2, sim["Z"], 0.4, 0.1, seed=1)
    assert np.array_equal(a[0].cpu().numpy(), b[0]) and np.array_equal(a[1], b[1])


def test_golden_stages(sb):
    """CUDA path against the committed fixtures of the stages around the hot path (oracle-generated)."""
    g = np.load(os.path.join(GOLD, "stages_small.npz"))
    Y2, Z = g["Y2"], g["Z"]
    imp, pg, mu, sd = sb.init_impute(Y2, 2, Z, 0.4, 0.1, seed=3, sweep=1, details=True)
    assert rel(sd, g["init_sd"]) <= 1e-6 and rel(pg, g["init_pg"]) <= 1e-6 and _draw_mismatch(imp, g["init_impute"], 1e-5) <= 3
    impb = sb.init_impute_bulk(Y2, Z, None, g["pg1"], 0.1, seed=4, sweep=2)
    assert _draw_mismatch(impb, g["init_bulk"], 1e-5) <= 3
    cl = sb.init_cluster(Y2[:60], 4, 2, seed=6, nstart=5, iter_max=15, details=True)
    assert rel(cl["d"], g["d"]) <= 1e-8 and O.adjusted_rand_index(cl["clus"], g["clus"]) >= 0.99
    lq = np.arange(40, 90)
    fit = sb.lq_fit(Y2[lq], g["lq_sigma_in"], g["lq_pg"], g["em_z"], list(g["em_Ef"]), list(g["em_Varf"]), g["celltype"],
                    resid_mode=1, seed=9, sweep=3)
    for k in ("mu", "beta", "sigma"):
        assert rel(fit[k], g["lq_" + k]) <= 1e-8, k
    assert rel(fit["Y"], g["lq_Y"]) <= 1e-6
    run = sb.SIMPLE(Y2, K0=3, M0=2, iter=3, clus=Z, min_gene=45, mcmc=2, burnin=1, seed=5)
    for k in ("loglik_tot", "beta", "sigma", "mu", "pi", "lambda", "z", "impt", "Yimp0", "pg"):
        assert rel(run[k], g["run_" + k]) <= 1e-5, k
    assert abs(run["BIC"] - float(g["run_BIC"])) <= 1e-6 * abs(float(g["run_BIC"]))


def test_driver_properties_at_scale(sb):
    """Size-independent properties of a whole SIMPLE() run at 20k cells x 600 genes (the oracle driver would take
    minutes here): recovered cell types, observed entries untouched, finite outputs, BIC consistent with the last
    log-likelihood, and the device-resident run identical to the host-buffer run."""
    from simples_b200 import synth
    gp = synth.gene_params(600, 5, 5, seed=2)
    Y2, Z = synth.simulate_shard(gp, 20000, shard=0, seed=2)
    kw = dict(K0=5, M0=5, iter=4, clus=None, K=8, p_min=0.4, min_gene=300, mcmc=4, burnin=1, seed=11, consensus=False)
    res = sb.SIMPLE(Y2, **kw)
    G, n = Y2.shape
    assert O.adjusted_rand_index(np.argmax(res["z"], axis=1), Z) >= 0.85
    obs = Y2 > 0.1
    # observed entries come back through the centring round trip (y - mean) + mean, as in the reference: 1 ulp
    assert np.allclose(res["impt"][obs], Y2[obs], rtol=1e-14, atol=1e-14)
    assert np.isfinite(res["impt"]).all() and np.isfinite(res["Yimp0"]).all()
    assert (res["impt_var"][obs] == 0).all() and (res["impt_var"] >= 0).all()
    bic, bic0 = O.bic_ref(res["loglik_tot"][-1], n, G, 5, 5)
    assert abs(res["BIC"] - bic) <= 1e-9 * abs(bic) and abs(res["BIC0"] - bic0) <= 1e-9 * abs(bic0)
    assert abs(res["pi"].sum() - 1) <= 1e-12 and np.allclose(res["z"].sum(axis=1), 1.0, atol=1e-10)
    assert (res["sigma"] >= 1e-4).all() and (res["sigma"] <= 9).all() and res["consensus_cluster"] is None
    dev = sb.SIMPLE(Y2, resident=True, **kw)
    assert np.array_equal(res["beta"], dev["beta"]) and np.array_equal(res["impt"], dev["impt"].cpu().numpy())


@pytest.mark.parametrize("K0,M0", [(1, 1), (3, 1), (2, 5)])
def test_do_impute_small_models(sb, K0, M0):
    """do_impute with M0 = 1 (the reference still calls rmultinom there, Q12) and K0 = 1."""
    st = make_case(150, 90, 2, max(M0, 1), K0, M0, seed=40 + K0 + M0)
    em = O.em_impute_fast(*em_args(st, M0, K0, 2), celltype=st["celltype"], impt_it=1, est_z=1, est_lam=1, seed=2)
    args = (st["Y0"], em["Y"], em["beta"], em["lambda"], em["sigma"], em["mu"], em["pi"], em["geneM"], em["geneSd"],
            st["celltype"])
    kw = dict(mcmc=3, burnin=1, pg=st["pg"], cutoff=0.1, seed=8)
    ref, got = O.do_impute_ref(*args, **kw), sb.do_impute(*args, **kw)
    assert rel(got["impt"], ref["impt"]) <= 1e-6 and abs(got["loglik"] - ref["loglik"]) <= 1e-8 * abs(ref["loglik"])
    assert rel(got["EF"], ref["EF"]) <= 1e-7 and rel(got["consensus_cluster"], ref["consensus_cluster"]) <= 1e-8


def test_do_impute_burnin_zero_drops_the_first_sweep(sb):
    """Q21: `mean(tot[-1:-burnin])` (R/do_impute.R:158) with burnin = 0 indexes with -1:0 = c(-1, 0), i.e. the first
    sweep is still dropped from the log-likelihood mean (and mcmc = 1 gives NaN); everything else records all sweeps."""
    K0, M0 = 3, 2
    st = make_case(120, 96, 2, 2, K0, M0, seed=51)
    em = O.em_impute_fast(*em_args(st, M0, K0, 2), celltype=st["celltype"], impt_it=1, est_z=1, est_lam=1, seed=2)
    args = (st["Y0"], em["Y"], em["beta"], em["lambda"], em["sigma"], em["mu"], em["pi"], em["geneM"], em["geneSd"],
            st["celltype"])
    kw = dict(pg=st["pg"], cutoff=0.1, seed=8)
    g0, r0 = sb.do_impute(*args, mcmc=3, burnin=0, **kw), O.do_impute_ref(*args, mcmc=3, burnin=0, **kw)
    g1 = sb.do_impute(*args, mcmc=2, burnin=1, **kw)               # same chain, first sweep as burn-in
    assert g0["loglik"] == g1["loglik"]
    assert abs(g0["loglik"] - r0["loglik"]) <= 1e-8 * abs(r0["loglik"]) and rel(g0["impt"], r0["impt"]) <= 1e-6
    assert not np.allclose(g0["impt"], g1["impt"])                  # burnin = 0 records the first sweep, burnin = 1 does not
    assert np.isnan(sb.do_impute(*args, mcmc=1, burnin=0, **kw)["loglik"])
