"""GPU parity at size (run on a B200 with -m gpu): the named configurations at sizes where the Monte-Carlo path, the
multiple-imputation sampler and the SIMPLE_B driver meet the oracle on more than a few hundred cells, and the
multi-rank NCCL path under the test runner.  Tolerances as everywhere: <= 1e-8 relative on parameters / log-likelihood,
<= 1e-6 on imputed entries, identical cluster assignments."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle as O
from util import assert_em_close, em_args, make_case, rel

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sb(built_lib):
    import simples_b200
    simples_b200.init(0)
    return simples_b200


def test_c3_shape_monte_carlo_em_20k_cells(sb):
    """C3's genes / factors / clusters (2000 x K0 = M0 = 10) on 20k cells: one deterministic iteration and two
    Monte-Carlo iterations (2 sweeps each = 4 sweeps, 5 E-passes on sampled data, 3 M-steps) against the oracle."""
    K0 = M0 = 10
    st = make_case(20000, 2000, 10, 10, K0, M0, seed=5)
    kw = dict(celltype=st["celltype"], impt_it=1, est_z=1, est_lam=1, seed=7, num_mc=2)
    ref = O.em_impute_fast(*em_args(st, M0, K0, 3), **kw)
    got = sb.EM_impute(*em_args(st, M0, K0, 3), **kw)
    assert_em_close(got, ref, M0)
    # the fused sweep + projection kernel (opt-in) computes the same MC passes
    os.environ["SIMPLES_FUSED_MC"] = "1"
    try:
        fused = sb.EM_impute(*em_args(st, M0, K0, 3), **kw)
    finally:
        os.environ.pop("SIMPLES_FUSED_MC", None)
    assert_em_close(fused, ref, M0)
    assert rel(fused["Y"], got["Y"]) <= 1e-10


def test_do_impute_5k_cells(sb):
    K0, M0 = 8, 5
    st = make_case(5000, 1000, 6, 5, K0, M0, seed=15, MB=3)
    em = O.em_impute_fast(*em_args(st, M0, K0, 2), celltype=st["celltype"], impt_it=1, est_z=1, est_lam=1, seed=2)
    args = (st["Y0"], em["Y"], em["beta"], em["lambda"], em["sigma"], em["mu"], em["pi"], em["geneM"], em["geneSd"],
            st["celltype"])
    kw = dict(mcmc=3, burnin=1, pg=st["pg"], cutoff=0.1, seed=21, consensus=False)
    ref = O.do_impute_ref(*args, **kw)
    for fused in ("0", "1"):
        os.environ["SIMPLES_FUSED_MC"] = fused
        try:
            got = sb.do_impute(*args, **kw)
        finally:
            os.environ.pop("SIMPLES_FUSED_MC", None)
        assert abs(got["loglik"] - ref["loglik"]) <= 1e-8 * abs(ref["loglik"])
        assert rel(got["impt"], ref["impt"]) <= 1e-6
        assert np.abs(got["impt_var"] - ref["impt_var"]).max() <= 1e-6 * max(1.0, np.abs(ref["impt_var"]).max())
        assert rel(got["EF"], ref["EF"]) <= 1e-7 and rel(got["varF"], ref["varF"]) <= 1e-7


def test_simple_b_c2_substitute(sb):
    """BASELINE config 2 (bundled chu hESC data, SIMPLE_B with bulk means and 7 cell types, K0 = 10): chu.RData is not in
    the checkout, so the stated substitute is used: 3000 genes x 763 cells simulated with 7 clusters
    (utils/utils.R:2-92 restated), bulk = the per-type mean profile (utils/utils.R:91), celltype = the true types."""
    sim = O.simulate(n=763, G=3000, K=6, MC=7, S0=20, seed=11)
    bulk = np.maximum(sim["Mu"], 0.0)
    kw = dict(M0=7, clus=sim["Z"], b=1, iter=3, min_gene=300, mcmc=2, burnin=1, seed=8)
    ref = O.simple_b_ref(sim["Y2"], 10, bulk, sim["Z"], **kw)
    got = sb.SIMPLE_B(sim["Y2"], 10, bulk, sim["Z"], **kw)
    assert "initclus" not in got and "p_min" not in got
    for k in ("loglik_tot", "pi", "mu", "sigma", "beta", "lambda", "z", "Yimp0", "pg", "impt"):
        assert rel(got[k], ref[k]) <= 1e-5, (k, rel(got[k], ref[k]))         # 33 EM iterations on top of a sqrt(eps) Brent fit
    assert np.array_equal(np.argmax(got["z"], 1), np.argmax(ref["z"], 1))
    assert got["pg"].shape == (3000, 7)


def _ngpu():
    import torch
    return torch.cuda.device_count()


def test_two_rank_nccl_parity():
    """Cells sharded over 2 ranks + one NCCL all-reduce per iteration == the single-GPU result (EM_impute, do_impute, the
    SIMPLE driver incl. the clustering start, selectKM over replicas).  Needs >= 2 visible GPUs."""
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    env = dict(os.environ, NCCL_DEBUG="WARN")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29617", os.path.join(ROOT, "tests", "helpers", "mgpu_check.py")],
                       capture_output=True, text=True, timeout=1500, env=env, cwd=ROOT)
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines and all(l["ok"] for l in lines), (r.stdout[-3000:], r.stderr[-3000:])


def test_single_process_multi_gpu_c_abi(sb):
    """simples_init(ndev = 2): one process drives two GPUs through the C ABI alone (host threads + in-library NCCL), the
    route an R session takes; equals the 1-GPU result to reduction-order rounding."""
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    K0, M0 = 6, 4
    st = make_case(3001, 700, 4, 4, K0, M0, seed=31)
    kw = dict(celltype=st["celltype"], impt_it=1, est_z=1, est_lam=1, seed=13, num_mc=2)
    one = sb.EM_impute(*em_args(st, M0, K0, 4), **kw)
    mi_kw = dict(mcmc=3, burnin=1, pg=st["pg"], seed=5, consensus=True)
    mi_args = (st["Y0"], one["Y"], one["beta"], one["lambda"], one["sigma"], one["mu"], one["pi"], one["geneM"],
               one["geneSd"], st["celltype"])
    mi_one = sb.do_impute(*mi_args, **mi_kw)
    sb.init([0, 1])
    try:
        two = sb.EM_impute(*em_args(st, M0, K0, 4), **kw)
        mi_two = sb.do_impute(*mi_args, **mi_kw)
    finally:
        sb.init(0)
    for k in ("loglik", "pi", "mu", "sigma", "beta", "lambda", "geneM", "z", "Y"):
        assert rel(two[k], one[k]) <= 1e-9, (k, rel(two[k], one[k]))
    for m in range(M0):
        assert rel(two["Ef"][m], one["Ef"][m]) <= 1e-8
    assert abs(mi_two["loglik"] - mi_one["loglik"]) <= 1e-9 * abs(mi_one["loglik"])
    assert rel(mi_two["impt"], mi_one["impt"]) <= 1e-8 and rel(mi_two["EF"], mi_one["EF"]) <= 1e-8
    assert rel(mi_two["consensus_cluster"], mi_one["consensus_cluster"]) <= 1e-9     # row blocks filled by the two devices


def test_expect_mode_matches_oracle_20k_cells(sb):
    """mode = "expect" (SIMPLES_MODE_EXPECT): the deterministic companion of the Monte-Carlo sweep -- most probable
    cluster, factor at its posterior mean, every zero entry at its conditional expectation.  No variate is consumed, so
    this checks the whole sweep data path (mask, compaction, DMMA conditional means, stores, the E-passes on the updated
    matrix, the M-step on it) at size independently of the Philox tape."""
    K0 = M0 = 10
    st = make_case(20000, 2000, 10, 10, K0, M0, seed=6)
    kw = dict(celltype=st["celltype"], impt_it=1, est_z=1, est_lam=1, seed=7, num_mc=2)
    ref = O.em_impute_fast(*em_args(st, M0, K0, 3), mode="expect", **kw)
    got = sb.EM_impute(*em_args(st, M0, K0, 3), mode="expect", **kw)
    assert_em_close(got, ref, M0)
    other = sb.EM_impute(*em_args(st, M0, K0, 3), mode="expect", **dict(kw, seed=99))
    assert np.array_equal(other["Y"], got["Y"]) and np.array_equal(other["beta"], got["beta"])     # tape-independent
    samp = sb.EM_impute(*em_args(st, M0, K0, 3), **kw)
    assert not np.array_equal(samp["Y"], got["Y"])


def test_r_golden_loader_plumbing_gpu(tmp_path):
    """The CUDA half of tests/test_r_golden.py executed end to end against the oracle stand-in directory (see
    tests/helpers/fake_r_golden.py): a plumbing check of the loader, not a parity pin."""
    gold = str(tmp_path / "golden")
    subprocess.run([sys.executable, os.path.join(ROOT, "tests", "helpers", "fake_r_golden.py"), gold], check=True, cwd=ROOT)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_r_golden.py"), "-q", "-m", "gpu",
                        "-p", "no:cacheprovider"], capture_output=True, text=True, cwd=ROOT, env=dict(os.environ, SIMPLES_R_GOLDEN=gold))
    assert r.returncode == 0 and "2 passed" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
