import numpy as np
import oracle as O


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))


def make_case(n, G, K, MC, K0, M0, seed, nonshared=False, MB=None):
    sim = O.simulate(n=n, G=G, K=K, MC=MC, S0=min(10, G // (2 * MC)), seed=seed)
    st = O.bench_state(sim, K0, M0, seed=seed)
    if nonshared:
        st["sigma"] = st["sigma"] * np.random.default_rng(seed).uniform(0.5, 1.5, size=st["sigma"].shape)
    if MB is not None:   # cell types independent of clusters (SIMPLE_B style, R/SIMPLE_B.R:318-321)
        rng = np.random.default_rng(seed + 5)
        st["celltype"] = rng.integers(1, MB + 1, size=st["Y"].shape[1]).astype(np.int32)
        st["celltype"][:MB] = np.arange(1, MB + 1)
        st["pg"] = rng.uniform(0.3, 0.9, size=(st["Y"].shape[0], MB))
    return st


def em_args(st, M0, K0, iters, cutoff=0.1):
    return (st["Y"], st["Y0"], st["pg"], M0, K0, cutoff, iters, st["beta"], st["sigma"], st["lam"], st["pi"], st["z"])


PARAM_KEYS = ("loglik", "pi", "mu", "sigma", "beta", "lambda", "z", "geneM", "geneSd")


def assert_em_close(got, ref, M0, tol_param=1e-8, tol_y=1e-6):
    """north_star tolerances: <= 1e-8 relative on parameters / log-lik, <= 1e-6 on imputed entries,
    identical cluster assignments."""
    for k in PARAM_KEYS:
        assert rel(got[k], ref[k]) <= tol_param, (k, rel(got[k], ref[k]))
    assert rel(got["Y"], ref["Y"]) <= tol_y, ("Y", rel(got["Y"], ref["Y"]))
    for m in range(M0):
        assert rel(got["Ef"][m], ref["Ef"][m]) <= tol_param * 10, ("Ef", m)
        assert rel(got["Varf"][m], ref["Varf"][m]) <= tol_param, ("Varf", m)
    assert abs(got["priorB"] - ref["priorB"]) <= tol_param * max(1.0, abs(ref["priorB"]))
    assert np.array_equal(np.argmax(got["z"], axis=1), np.argmax(ref["z"], axis=1))
