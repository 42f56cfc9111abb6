"""The R side of the drop-in boundary, exercised without R.

``r/simples_rcall.c`` (the ``.Call`` shim a SIMPLEs maintainer adds as ``src/simples_rcall.c``) is compiled UNMODIFIED
against ``tests/helpers/rstub`` -- a heap-object stand-in for the dozen R API entry points the shim touches, with R's
semantics (``Rf_ncols`` of a plain vector is 1, ``Rf_error`` does not return) -- and linked with libsimples_b200.so.

CPU: the shim compiles, ``R_init_SIMPLEs`` registers every routine with the argument count the R wrappers in ``r/*.R``
pass to ``.Call``, and errors raised by the library surface as R errors (no device here => SIMPLES_ENODEV).
GPU (``-m gpu``): ``.Call("C_simples_em_impute", ...)`` / ``"C_simples_do_impute"`` / ``"C_simples_init_impute"`` with R-shaped
arguments return the reference's named lists (R/EM_impute.R:337-338, R/do_impute.R:158-159) and the same numbers as the
Python mirror of the same C ABI.
"""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "helpers", "rstub")
OUT = os.path.join(STUB, "_build")
SHIM = os.path.join(OUT, "librshim.so")

REALSXP, INTSXP, VECSXP, NILSXP = 14, 13, 19, 0


def _build_shim():
    from simples_b200 import _build
    lib = _build.build()
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(STUB, "rstub.c"), os.path.join(ROOT, "r", "simples_rcall.c")]
    deps = srcs + [os.path.join(STUB, "rstub.h"), os.path.join(ROOT, "include", "simples_b200.h"), lib]
    if not os.path.exists(SHIM) or os.path.getmtime(SHIM) < max(os.path.getmtime(d) for d in deps):
        libdir = os.path.dirname(lib)
        cmd = ["gcc", "-std=gnu11", "-O1", "-shared", "-fPIC", "-I" + STUB, "-I" + os.path.join(ROOT, "include"), "-o", SHIM,
               *srcs, "-L" + libdir, "-l:libsimples_b200.so", "-Wl,-rpath," + libdir]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    return SHIM


class RStub:
    """ctypes view of the stand-in: build R-shaped arguments, .Call a registered routine, read the result back."""

    def __init__(self):
        L = ctypes.CDLL(_build_shim())
        vp, ll, ci, dp = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.POINTER(ctypes.c_double)
        sig = {"rstub_nil": (vp, []), "rstub_real": (vp, [dp, ll, ci, ci]), "rstub_int": (vp, [ctypes.POINTER(ci), ll]),
               "rstub_lgl": (vp, [ci]), "rstub_list": (vp, [ll]), "rstub_type": (ci, [vp]), "rstub_len": (ll, [vp]),
               "rstub_nrow": (ci, [vp]), "rstub_ncol": (ci, [vp]), "rstub_data": (vp, [vp]),
               "rstub_name": (ctypes.c_char_p, [vp, ll]), "rstub_registered": (ci, []),
               "rstub_routine_name": (ctypes.c_char_p, [ci]), "rstub_routine_nargs": (ci, [ci]),
               "rstub_dynamic_symbols": (ci, []), "rstub_dot_call": (vp, [ctypes.c_char_p, ci, ctypes.POINTER(vp)]),
               "rstub_last_error": (ctypes.c_char_p, []), "rstub_free_all": (None, []), "R_init_SIMPLEs": (None, [vp]),
               "SET_VECTOR_ELT": (vp, [vp, ll, vp]), "VECTOR_ELT": (vp, [vp, ll])}
        for k, (res, args) in sig.items():
            f = getattr(L, k)
            f.restype, f.argtypes = res, args
        self.L = L
        L.R_init_SIMPLEs(None)

    # ---- R values ----
    def mat(self, a):                     # double matrix, column-major like R
        a = np.asfortranarray(a, dtype=np.float64)
        return self.L.rstub_real(a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), a.size, a.shape[0], a.shape[1])

    def vec(self, a):                     # double vector / scalar
        a = np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64)
        return self.L.rstub_real(a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), a.size, a.size, -1)

    def ivec(self, a):
        a = np.ascontiguousarray(np.atleast_1d(a), dtype=np.int32)
        return self.L.rstub_int(a.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), a.size)

    def lgl(self, v):
        return self.L.rstub_lgl(int(bool(v)))

    def nil(self):
        return self.L.rstub_nil()

    def rlist(self, items):
        x = self.L.rstub_list(len(items))
        for i, it in enumerate(items):
            self.L.SET_VECTOR_ELT(x, i, it)
        return x

    def call(self, name, *args):
        arr = (ctypes.c_void_p * len(args))(*args)
        res = self.L.rstub_dot_call(name.encode(), len(args), arr)
        if not res:
            raise RuntimeError(self.L.rstub_last_error().decode())
        return res

    def py(self, x):
        """SEXP -> numpy array / dict (named list) / list / None."""
        t = self.L.rstub_type(x)
        if t == NILSXP:
            return None
        n = self.L.rstub_len(x)
        if t == VECSXP:
            items = [self.py(self.L.VECTOR_ELT(x, i)) for i in range(n)]
            nm = self.L.rstub_name(x, 0) if n else None
            return {self.L.rstub_name(x, i).decode(): it for i, it in enumerate(items)} if nm else items
        ct = ctypes.c_double if t == REALSXP else ctypes.c_int
        a = np.ctypeslib.as_array(ctypes.cast(self.L.rstub_data(x), ctypes.POINTER(ct)), shape=(max(n, 1),))[:n].copy()
        nc = self.L.rstub_ncol(x)
        return a if nc < 0 else a.reshape((self.L.rstub_nrow(x), nc), order="F")


@pytest.fixture(scope="module")
def R():
    return RStub()


def _dot_call_arity(path):
    """{routine: number of arguments after the routine name} of every .Call in an R file (comments included: the
    commented driver snippets must stay in step with the table too)."""
    src = open(path).read()
    src = re.sub(r"(?m)^\s*#\s?", "", src)
    out = {}
    for m in re.finditer(r"\.Call\(\s*\"?(C_simples_\w+)\"?\s*,", src):
        i, depth, nargs, cur = m.end(), 1, 0, False
        while depth:
            c = src[i]
            if c in "([":
                depth += 1
            elif c in ")]":
                depth -= 1
                if depth == 0 and cur:
                    nargs += 1
            elif c == "," and depth == 1:
                nargs += 1
                cur = False
                i += 1
                continue
            if not c.isspace() and depth:
                cur = True
            i += 1
        call = src[m.end():i]
        nargs -= len(re.findall(r"\bPACKAGE\s*=", call))
        out.setdefault(m.group(1), set()).add(nargs)
    return out


def test_shim_compiles_and_registers_what_the_r_wrappers_call(R):
    reg = {R.L.rstub_routine_name(i).decode(): R.L.rstub_routine_nargs(i) for i in range(R.L.rstub_registered())}
    assert reg == {"C_simples_init": 1, "C_simples_em_impute": 25, "C_simples_do_impute": 16, "C_simples_init_impute": 7,
                   "C_simples_init_cluster": 6, "C_simples_lq_fit": 12}
    assert R.L.rstub_dynamic_symbols() == 0            # R_useDynamicSymbols(dll, FALSE)
    used = {}
    for f in ("EM_impute_b200.R", "SIMPLE_b200.R"):
        for k, v in _dot_call_arity(os.path.join(ROOT, "r", f)).items():
            used.setdefault(k, set()).update(v)
    assert set(used) == set(reg), (sorted(used), sorted(reg))
    for k, counts in used.items():
        assert counts == {reg[k]}, (k, counts, reg[k])
    # every C entry point the shim binds is declared in the public header
    hdr = open(os.path.join(ROOT, "include", "simples_b200.h")).read()
    shim = open(os.path.join(ROOT, "r", "simples_rcall.c")).read()
    for fn in set(re.findall(r"\b(simples_[a-z_]+)\(", shim)):
        assert re.search(r"\b%s\s*\(" % fn, hdr), fn


def test_dot_call_argument_and_symbol_checks(R):
    with pytest.raises(RuntimeError, match="Incorrect number of arguments"):
        R.call("C_simples_init", R.ivec([0]), R.ivec([1]))
    with pytest.raises(RuntimeError, match="not in registered"):
        R.call("C_simples_nope", R.ivec([0]))
    # a wrongly typed argument is an R error, not a crash: REAL() of an integer matrix
    with pytest.raises(RuntimeError, match="REAL\\(\\) can only be applied"):
        R.call("C_simples_init_impute", R.ivec([1, 2, 3, 4]), R.ivec([1]), R.ivec([1, 1]), R.nil(), R.vec(0.6), R.vec(0.1),
               R.vec(0))


def _em_case(seed=2):
    import oracle as O
    sim = O.simulate(n=96, G=128, K=3, MC=2, S0=10, seed=seed)
    return O.bench_state(sim, 4, 2)


def _em_call_args(R, st, M0, K0, iters, seed):
    return (R.mat(st["Y"]), R.mat(st["Y0"]), R.mat(np.atleast_2d(st["pg"]).reshape(st["Y"].shape[0], -1)), R.ivec(M0),
            R.ivec(K0), R.vec(0.1), R.ivec(iters), R.mat(st["beta"]), R.mat(st["sigma"]), R.mat(st["lam"]), R.vec(st["pi"]),
            R.mat(st["z"]), R.nil(), R.ivec(st["celltype"]), R.vec(1.0), R.ivec(1), R.lgl(True), R.ivec(1), R.ivec(1),
            R.vec(100.0), R.vec(1.0), R.ivec(3), R.vec(-np.inf), R.vec(np.inf), R.vec(float(seed)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


def test_library_errors_surface_as_r_errors(R):
    if _has_gpu():
        pytest.skip("needs a host without a CUDA device (the error path tested is SIMPLES_ENODEV)")
    st = _em_case()
    with pytest.raises(RuntimeError, match=r"simples_b200: .* \(status -?\d+\)"):
        R.call("C_simples_em_impute", *_em_call_args(R, st, 2, 4, 2, 3))
    with pytest.raises(RuntimeError, match=r"simples_b200: .* \(status -?\d+\)"):
        R.call("C_simples_init", R.ivec([0]))
    R.L.rstub_free_all()


@pytest.mark.gpu
def test_dot_call_em_impute_and_do_impute_match_the_python_mirror(R):
    import simples_b200 as sb
    st = _em_case()
    M0, K0 = 2, 4
    assert R.py(R.call("C_simples_init", R.ivec([0])))[0] >= 1
    res = R.py(R.call("C_simples_em_impute", *_em_call_args(R, st, M0, K0, 4, 3)))
    assert list(res) == ["loglik", "pi", "mu", "sigma", "beta", "lambda", "z", "Ef", "Varf", "Y", "geneM", "geneSd",
                         "priorB"]                                                  # R/EM_impute.R:337-338
    ref = sb.EM_impute(st["Y"], st["Y0"], st["pg"], M0, K0, 0.1, 4, st["beta"], st["sigma"], st["lam"], st["pi"], st["z"],
                       celltype=st["celltype"], impt_it=1, est_z=1, est_lam=1, seed=3)
    G, n = st["Y"].shape
    assert res["mu"].shape == (G, M0) and res["beta"].shape == (G, K0) and res["lambda"].shape == (M0, K0)
    assert res["z"].shape == (n, M0) and res["Y"].shape == (G, n)
    assert len(res["Ef"]) == M0 and res["Ef"][0].shape == (n, K0) and res["Varf"][1].shape == (K0, K0)
    for k in ("loglik", "pi", "mu", "sigma", "beta", "lambda", "z", "Y", "geneM", "geneSd"):
        assert np.array_equal(res[k], np.asarray(ref[k]).reshape(res[k].shape)), k  # same library, same seed: bit-identical
    for m in range(M0):
        assert np.array_equal(res["Ef"][m], ref["Ef"][m]) and np.array_equal(res["Varf"][m], ref["Varf"][m])
    assert res["priorB"][0] == ref["priorB"]

    mi = R.py(R.call("C_simples_do_impute", R.mat(st["Y0"]), R.mat(res["Y"]), R.mat(res["beta"]), R.mat(res["lambda"]),
                     R.mat(res["sigma"]), R.mat(res["mu"]), R.vec(res["pi"]), R.vec(res["geneM"]), R.vec(res["geneSd"]),
                     R.ivec(st["celltype"]), R.ivec(3), R.ivec(1), R.mat(np.atleast_2d(st["pg"]).reshape(G, -1)), R.vec(0.1),
                     R.vec(5.0), R.lgl(True)))
    assert list(mi) == ["loglik", "impt", "impt_var", "EF", "varF", "consensus_cluster"]   # R/do_impute.R:158-159
    mo = sb.do_impute(st["Y0"], res["Y"], res["beta"], res["lambda"], res["sigma"], res["mu"], res["pi"], res["geneM"],
                      res["geneSd"], st["celltype"], mcmc=3, burnin=1, pg=st["pg"], seed=5)
    assert mi["loglik"][0] == mo["loglik"]
    for k in ("impt", "impt_var", "EF", "varF", "consensus_cluster"):
        assert np.array_equal(mi[k], np.asarray(mo[k])), k
    # consensus = FALSE gives NULL in the list slot
    mi2 = R.py(R.call("C_simples_do_impute", R.mat(st["Y0"]), R.mat(res["Y"]), R.mat(res["beta"]), R.mat(res["lambda"]),
                      R.mat(res["sigma"]), R.mat(res["mu"]), R.vec(res["pi"]), R.nil(), R.nil(), R.nil(), R.ivec(2), R.ivec(1),
                      R.mat(np.full((G, 1), 0.5)), R.vec(0.1), R.vec(5.0), R.lgl(False)))
    assert mi2["consensus_cluster"] is None and mi2["impt"].shape == (G, n)
    # an invalid argument comes back as an R error carrying the library's message
    bad = list(_em_call_args(R, st, M0, K0, 2, 3))
    bad[4] = R.ivec(99)                                                             # K0 > 32
    with pytest.raises(RuntimeError, match=r"simples_b200: .* \(status -?\d+\)"):
        R.call("C_simples_em_impute", *bad)
    R.L.rstub_free_all()


@pytest.mark.gpu
def test_dot_call_init_impute_matches_the_python_mirror(R):
    import simples_b200 as sb
    rng = np.random.default_rng(4)
    G, n, M0 = 60, 90, 2
    clus = rng.integers(1, M0 + 1, n).astype(np.int32)
    Y2 = np.maximum(rng.normal(1.0, 1.0, (G, n)), 0.0) * (rng.uniform(size=(G, n)) < 0.7)
    out = R.py(R.call("C_simples_init_impute", R.mat(Y2), R.ivec(M0), R.ivec(clus), R.nil(), R.vec(0.6), R.vec(0.1), R.vec(7.0)))
    assert list(out) == ["impute", "pg"] and out["impute"].shape == (G, n) and out["pg"].shape == (G, M0)
    ref = sb.init_impute(Y2, M0, clus, p_min=0.6, cutoff=0.1, seed=7)
    got_imp, got_pg = (ref["impute"], ref["pg"]) if isinstance(ref, dict) else (ref[0], ref[1])
    assert np.array_equal(out["impute"], np.asarray(got_imp)) and np.array_equal(out["pg"], np.asarray(got_pg))
    R.L.rstub_free_all()


@pytest.mark.gpu
def test_dot_call_lq_fit_and_init_cluster_match_the_python_mirror(R):
    """The call sites inside SIMPLE() / SIMPLE_B() (r/SIMPLE_b200.R): clustering start and lq-gene step."""
    import simples_b200 as sb
    st = _em_case(seed=6)
    M0, K0 = 2, 4
    em = sb.EM_impute(st["Y"], st["Y0"], st["pg"], M0, K0, 0.1, 3, st["beta"], st["sigma"], st["lam"], st["pi"], st["z"],
                      celltype=st["celltype"], impt_it=1, est_z=1, est_lam=1, seed=3)
    rng = np.random.default_rng(8)
    Gq, n = 40, st["Y"].shape[1]
    dat = np.maximum(rng.normal(0.3, 1.0, (Gq, n)), 0.0) * (rng.uniform(size=(Gq, n)) < 0.4)       # lq genes: mostly zeros
    sig = np.full((Gq, M0), 1.0)
    pg = np.full((Gq, int(np.max(st["celltype"]))), 0.6)            # one column per cell type (R/SIMPLE.R:312-373)
    for resp, mode in ((None, 0), (dat + 0.1, 1)):
        out = R.py(R.call("C_simples_lq_fit", R.mat(dat), R.nil() if resp is None else R.mat(resp), R.mat(sig), R.mat(pg),
                          R.mat(em["z"]), R.rlist([R.mat(e) for e in em["Ef"]]), R.rlist([R.mat(v) for v in em["Varf"]]),
                          R.ivec(st["celltype"]), R.vec(1.0), R.vec(0.1), R.ivec(mode), R.vec(11.0)))
        assert list(out) == ["mu", "beta", "sigma", "Y"]
        ref = sb.lq_fit(dat, sig, pg, em["z"], em["Ef"], em["Varf"], celltype=st["celltype"], penl=1, cutoff=0.1, resp=resp,
                        resid_mode=mode, seed=11)
        for k in ("mu", "beta", "sigma", "Y"):
            assert np.array_equal(out[k], np.asarray(ref[k]).reshape(out[k].shape)), (k, mode)
    X = st["Y"]
    cl = R.py(R.call("C_simples_init_cluster", R.mat(X), R.ivec(5), R.ivec(3), R.ivec(20), R.ivec(30), R.vec(2.0)))
    assert cl.dtype == np.int32 and cl.shape == (n,) and set(cl) <= {1, 2, 3}
    assert np.array_equal(cl, sb.init_cluster(X, 5, 3, seed=2, nstart=20, iter_max=30))
    R.L.rstub_free_all()
