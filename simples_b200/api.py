"""Host-side mirror of the reference's hot-path R functions.

``EM_impute`` and ``do_impute`` keep the names, argument order, defaults and
return-list layout of ``R/EM_impute.R:75-78,337-338`` and
``R/do_impute.R:29-30,158-159`` (R is not available in this image, so the
reference's host language is mirrored in Python above the C ABI; the R glue is
shipped as source under ``r/``).  Matrices follow R's orientation: ``Y`` is
``G x n`` (genes x cells), ``z`` is ``n x M0``, ``beta`` is ``G x K0`` ...

All compute happens in libsimples_b200.so (CUDA, sm_100a).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import COMPAT_DEFAULT, EmArgs, EmOut, MiArgs, MiOut, SimplesError, check  # noqa: F401


def _is_dev(a):
    """A torch CUDA tensor: the big matrices (Y, Y0, dat, X ... and their outputs) may live in HBM already; the
    library copies with cudaMemcpyDefault, so a device pointer is as good as a host pointer."""
    return bool(getattr(a, "is_cuda", False))


def dev_matrix(a):
    """Host matrix (R orientation, rows x cols) -> column-major float64 matrix on the current CUDA device."""
    import torch
    a = np.asarray(a, dtype=np.float64)
    return torch.from_numpy(np.ascontiguousarray(a.T)).cuda().T


def dev_empty(rows, cols):
    """Uninitialised column-major float64 rows x cols matrix on the current CUDA device (for ``out=`` buffers)."""
    import torch
    return torch.empty((cols, rows), dtype=torch.float64, device="cuda").T


def _f(a, shape=None, name=""):
    """float64, column-major (R layout) view/copy; device matrices pass through after a layout check."""
    if _is_dev(a):
        import torch
        if a.dtype != torch.float64 or a.dim() != 2 or not a.T.is_contiguous():
            raise ValueError(f"{name}: a device matrix must be float64 and column-major (see simples_b200.dev_matrix)")
        if shape is not None and tuple(a.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(a.shape)}")
        torch.cuda.current_stream().synchronize()       # producers on torch's stream are done before the library reads
        return a
    a = np.asarray(a, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(a.shape)}")
    return np.asfortranarray(a)


def _ptr(a):
    if a is None:
        return None
    if _is_dev(a):
        return C.c_void_p(a.data_ptr())
    return a.ctypes.data_as(C.c_void_p)


def _check_out(b, shape, order, name):
    if _is_dev(b):
        import torch
        ok = b.dtype == torch.float64 and tuple(b.shape) == tuple(shape) and (b.T.is_contiguous() if order == "F" else b.is_contiguous())
    else:
        ok = b.dtype == np.float64 and tuple(b.shape) == tuple(shape) and (b.flags.f_contiguous if order == "F" else b.flags.c_contiguous)
    if not ok:
        raise ValueError(f"out[{name!r}] must be float64 {tuple(shape)} in {order} order")
    return b


MODE_EXPECT = 16     # SIMPLES_MODE_EXPECT (include/simples_b200.h)


class _Comm:
    """Adapter: any object with ``allreduce_sum_(dev_ptr, count, stream)`` plus
    ``n_total`` / ``cell_offset`` (see simples_b200.dist.CellShard)."""

    def __init__(self, comm):
        self.comm = comm
        self.error = None

        def _cb(user, buf, count, stream):
            try:
                comm.allreduce_sum_(buf, count, stream)
                return 0
            except Exception as e:  # noqa: BLE001 - must not propagate through C frames
                self.error = e
                return 1

        # native communicator (CellShard(native=True)): NULL callback => the library all-reduces over its own NCCL comm
        self.cb = _lib.ALLREDUCE_FN() if getattr(comm, "native", False) else _lib.ALLREDUCE_FN(_cb)


def _em_args(Y, Y0, pg, M0, K0, cutoff, iter, beta, sigma, lambda_, pi, z, mu, celltype, penl, est_z,
             max_lambda, est_lam, impt_it, sigma0, pi_alpha, num_mc, lower, upper, seed, ref_compat, comm, stream):
    Y = _f(Y)
    if Y.ndim != 2:
        raise ValueError("Y must be a G x n matrix")
    G, n = Y.shape
    Y0 = _f(Y0, (G, n), "Y0")
    pg = _f(pg)
    if pg.ndim == 1:
        pg = _f(pg.reshape(G, -1))
    if pg.shape[0] != G:
        raise ValueError("pg must have G rows")
    MB = pg.shape[1]
    keep = dict(Y=Y, Y0=Y0, pg=pg,
                beta=_f(beta, (G, K0), "beta"), sigma=_f(sigma, (G, M0), "sigma"),
                lambda_=_f(lambda_, (M0, K0), "lambda"), pi=_f(np.ravel(pi), (M0,), "pi"),
                z=None if z is None else _f(z, (n, M0), "z"),
                mu=None if mu is None else _f(mu, (G, M0), "mu"),
                celltype=None if celltype is None else np.ascontiguousarray(np.asarray(celltype).astype(np.int32)))
    if keep["celltype"] is not None and keep["celltype"].shape != (n,):
        raise ValueError("celltype must have length n")
    a = EmArgs()
    a.G, a.n, a.M0, a.K0, a.MB = G, n, int(M0), int(K0), MB
    for k in ("Y", "Y0", "pg", "beta", "sigma", "lambda_", "pi", "z", "mu", "celltype"):
        setattr(a, k, _ptr(keep[k]))
    a.cutoff, a.iter = float(cutoff), int(iter)
    a.penl = float(penl)
    a.est_z, a.max_lambda, a.est_lam, a.impt_it = int(est_z), int(bool(max_lambda)), int(est_lam), int(impt_it)
    a.sigma0, a.pi_alpha, a.num_mc = float(sigma0), float(pi_alpha), int(num_mc)
    a.lower, a.upper = float(lower), float(upper)
    a.seed, a.ref_compat_flags = int(seed) & 0xFFFFFFFFFFFFFFFF, int(ref_compat)
    if comm is not None:
        cw = _Comm(comm)
        keep["_comm"] = cw
        a.n_total, a.cell_offset = int(comm.n_total), int(comm.cell_offset)
        a.allreduce = cw.cb
    else:
        a.n_total, a.cell_offset = 0, 0
    a.stream = stream
    return a, keep, (G, n, MB)


def em_out_shapes(G, n, M0, K0, iter):
    """Shapes / orders of the EM_impute output buffers (for callers that pre-allocate them, e.g. in
    pinned host memory): name -> (shape, order)."""
    return dict(loglik=((max(iter, 1),), "C"), pi=((M0,), "C"), mu=((G, M0), "F"), sigma=((G, M0), "F"),
                beta=((G, K0), "F"), lambda_=((M0, K0), "F"), z=((n, M0), "F"), Ef=((M0, K0, n), "C"),
                Varf=((M0, K0, K0), "C"), Y=((G, n), "F"), geneM=((G,), "C"), geneSd=((G,), "C"), priorB=((1,), "C"))


def _em_out(G, n, M0, K0, iter, out=None):
    bufs = {}
    for k, (shape, order) in em_out_shapes(G, n, M0, K0, iter).items():
        b = None if out is None else out.get(k)
        if b is None:
            b = np.empty(shape, order=order)
        else:
            _check_out(b, shape, order, k)
        bufs[k] = b
    o = EmOut()
    for k, v in bufs.items():
        setattr(o, k, _ptr(v))
    return o, bufs


def _em_result(bufs, M0, iter):
    # Ef buffer holds M0 column-major n x K0 matrices back to back == C-order (M0, K0, n)
    Ef = [np.asfortranarray(bufs["Ef"][m].T) for m in range(M0)]
    Varf = [bufs["Varf"][m].copy() for m in range(M0)]
    return {"loglik": bufs["loglik"][:iter], "pi": bufs["pi"], "mu": bufs["mu"], "sigma": bufs["sigma"],
            "beta": bufs["beta"], "lambda": bufs["lambda_"], "z": bufs["z"], "Ef": Ef, "Varf": Varf,
            "Y": bufs["Y"], "geneM": bufs["geneM"], "geneSd": bufs["geneSd"], "priorB": float(bufs["priorB"][0])}


def EM_impute(Y, Y0, pg, M0, K0, cutoff, iter, beta, sigma, lambda_, pi, z, mu=None, celltype=None, penl=1,
              est_z=2, max_lambda=True, est_lam=2, impt_it=5, sigma0=100, pi_alpha=1, verbose=False, num_mc=3,
              lower=-math.inf, upper=math.inf, *, seed=0, ref_compat=COMPAT_DEFAULT, comm=None, stream=None, out=None,
              mode="sample"):
    """Monte-Carlo EM imputation on B200 -- drop-in for ``EM_impute`` (R/EM_impute.R:75-339).

    Returns a dict with the reference's list names in the reference's order:
    loglik, pi, mu, sigma, beta, lambda, z, Ef, Varf, Y, geneM, geneSd, priorB.
    ``verbose`` is accepted and ignored (the reference's verbose block uses
    undefined globals, R/EM_impute.R:327-329).  Keyword-only extras: ``seed``
    (Philox key replacing R's global RNG), ``ref_compat`` (quirk bits),
    ``comm`` (a simples_b200.dist.CellShard when cells are sharded over GPUs), ``out`` (optional dict of
    pre-allocated output arrays, see ``em_out_shapes`` -- lets a caller keep results in pinned memory),
    ``mode``: ``"sample"`` (the reference's behaviour) or ``"expect"`` (deterministic companion: most probable cluster,
    factor at its posterior mean, zero entries at their conditional expectation, cf. R/SIMPLE.R:59).
    """
    lib = _lib.load()
    if mode not in ("sample", "expect"):
        raise ValueError("mode must be 'sample' or 'expect'")
    if mode == "expect":
        ref_compat = int(ref_compat) | MODE_EXPECT
    a, keep, (G, n, _) = _em_args(Y, Y0, pg, M0, K0, cutoff, iter, beta, sigma, lambda_, pi, z, mu, celltype, penl,
                                  est_z, max_lambda, est_lam, impt_it, sigma0, pi_alpha, num_mc, lower, upper, seed,
                                  ref_compat, comm, stream)
    o, bufs = _em_out(G, n, int(M0), int(K0), int(iter), out)
    rc = lib.simples_em_impute(C.byref(a), C.byref(o))
    if rc < 0 and "_comm" in keep and keep["_comm"].error is not None:
        raise keep["_comm"].error
    check(rc)
    return _em_result(bufs, int(M0), int(iter))


class EMContext:
    """Device-resident EM_impute state (simples_em_create / run / fetch): keeps
    Y, the zero mask and the parameters in HBM between calls."""

    def __init__(self, Y, Y0, pg, M0, K0, cutoff, beta, sigma, lambda_, pi, z, mu=None, celltype=None, penl=1,
                 est_z=2, max_lambda=True, est_lam=2, impt_it=5, sigma0=100, pi_alpha=1, num_mc=3,
                 lower=-math.inf, upper=math.inf, *, iter=1, seed=0, ref_compat=COMPAT_DEFAULT, comm=None, stream=None):
        self._lib = _lib.load()
        a, keep, (G, n, _) = _em_args(Y, Y0, pg, M0, K0, cutoff, iter, beta, sigma, lambda_, pi, z, mu, celltype,
                                      penl, est_z, max_lambda, est_lam, impt_it, sigma0, pi_alpha, num_mc, lower,
                                      upper, seed, ref_compat, comm, stream)
        self._keep = keep
        self.G, self.n, self.M0, self.K0 = G, n, int(M0), int(K0)
        self.iters_done = 0
        h = C.c_void_p()
        self._keep = keep
        self._check(self._lib.simples_em_create(C.byref(a), C.byref(h)))
        self._h = h

    def _check(self, rc):
        """Raise the Python exception an all-reduce callback stored (instead of the generic ECOMM), else the library's."""
        cw = getattr(self, "_keep", {}).get("_comm")
        if rc < 0 and cw is not None and cw.error is not None:
            err, cw.error = cw.error, None
            raise err
        return check(rc)

    def run(self, iters):
        self._check(self._lib.simples_em_run(self._h, int(iters)))
        self.iters_done += int(iters)

    def sync(self):
        check(self._lib.simples_em_sync(self._h))

    def iter_ms(self):
        buf = (C.c_double * 4096)()
        k = check(self._lib.simples_em_iter_ms(self._h, buf, 4096))
        return [buf[i] for i in range(min(k, 4096))]

    def set_profile(self, on):
        check(self._lib.simples_em_set_profile(self._h, int(bool(on))))

    def profile(self):
        k = self._lib.simples_kernel_count()
        la = (C.c_int64 * k)()
        ms = (C.c_double * k)()
        check(self._lib.simples_em_profile(self._h, la, ms, k))
        return {self._lib.simples_kernel_name(i).decode(): (int(la[i]), float(ms[i])) for i in range(k)}

    @property
    def nnz0(self):
        return int(self._lib.simples_em_nnz0(self._h))

    @property
    def kernel_launches(self):
        return int(self._lib.simples_em_kernel_launches(self._h))

    def fetch(self, big=True, out=None):
        o, bufs = _em_out(self.G, self.n, self.M0, self.K0, self.iters_done, out)
        if not big:
            o.Y = None
            o.Ef = None
        check(self._lib.simples_em_fetch(self._h, C.byref(o)))
        return _em_result(bufs, self.M0, self.iters_done)

    def yimp0(self, out=None):
        """``Yimp0`` of SIMPLE()/SIMPLE_B() (R/SIMPLE.R:421-426) for this shard, G x n."""
        if out is None:
            out = np.empty((self.G, self.n), order="F")
        check(self._lib.simples_em_yimp0(self._h, _ptr(out)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            self._lib.simples_ctx_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def do_impute(dat, Y, beta, lambda_, sigma, mu, pi, pos_mean=None, pos_sd=None, celltype=None, mcmc=10, burnin=2,
              verbose=False, pg=0.5, cutoff=0.1, *, seed=0, ref_compat=COMPAT_DEFAULT, consensus=True, comm=None,
              stream=None, out=None):
    """Multiple imputation at fixed parameters on B200 -- drop-in for
    ``do_impute`` (R/do_impute.R:29-160).  Returns loglik, impt, impt_var, EF,
    varF, consensus_cluster.  ``consensus=True``: the n x n matrix of R/do_impute.R:61 (also when this process
    drives several GPUs: every device fills its row block).  With ``comm`` (one process per GPU) the default is
    ``None``; ``consensus="rows"`` returns this rank's row block, n_local x n_total.  ``consensus=False``: ``None``."""
    lib = _lib.load()
    Y = _f(Y)
    G, n = Y.shape
    dat = _f(dat, (G, n), "dat")
    mu = _f(mu)
    beta = _f(beta)
    M0, K0 = mu.shape[1], beta.shape[1]                                  # :35-36
    ct = None if celltype is None else np.ascontiguousarray(np.asarray(celltype).astype(np.int32))
    MB = 1 if ct is None else int(ct.max())                              # :45
    pg = np.asarray(pg, dtype=np.float64)
    if pg.size == 1:
        pg = np.full((G, MB), float(pg))                                 # :47-48
    pg = _f(pg.reshape(G, -1))
    keep = dict(dat=dat, Y=Y, beta=_f(beta, (G, K0)), lambda_=_f(lambda_, (M0, K0), "lambda"),
                sigma=_f(sigma, (G, M0), "sigma"), mu=_f(mu, (G, M0)), pi=_f(np.ravel(pi), (M0,), "pi"),
                pos_mean=None if pos_mean is None else _f(np.ravel(pos_mean), (G,), "pos_mean"),
                pos_sd=None if pos_sd is None else _f(np.ravel(pos_sd), (G,), "pos_sd"), celltype=ct, pg=pg)
    a = MiArgs()
    a.G, a.n, a.M0, a.K0, a.MB = G, n, M0, K0, pg.shape[1]
    for k in ("dat", "Y", "beta", "lambda_", "sigma", "mu", "pi", "pos_mean", "pos_sd", "celltype", "pg"):
        setattr(a, k, _ptr(keep[k]))
    a.mcmc, a.burnin, a.cutoff = int(mcmc), int(burnin), float(cutoff)
    a.seed, a.ref_compat_flags = int(seed) & 0xFFFFFFFFFFFFFFFF, int(ref_compat)
    want_cons = (bool(consensus) and comm is None) or (comm is not None and consensus == "rows")
    a.want_consensus = int(want_cons)
    n_cols = n if comm is None else int(comm.n_total)
    cw = None
    if comm is not None:
        cw = _Comm(comm)
        a.n_total, a.cell_offset, a.allreduce = int(comm.n_total), int(comm.cell_offset), cw.cb
    a.stream = stream
    out = out or {}
    bufs = dict(loglik=np.zeros(1),
                impt=out.get("impt") if out.get("impt") is not None else np.empty((G, n), order="F"),
                impt_var=out.get("impt_var") if out.get("impt_var") is not None else np.empty((G, n), order="F"),
                EF=np.zeros((n, K0), order="F"), varF=np.zeros((n, K0), order="F"),
                consensus_cluster=np.zeros((n, n_cols), order="F") if want_cons else None)
    for k in ("impt", "impt_var"):
        _check_out(bufs[k], (G, n), "F", k)
    o = MiOut()
    for k, v in bufs.items():
        setattr(o, k, _ptr(v))
    rc = lib.simples_do_impute(C.byref(a), C.byref(o))
    if rc < 0 and cw is not None and cw.error is not None:
        raise cw.error
    check(rc)
    return {"loglik": float(bufs["loglik"][0]), "impt": bufs["impt"], "impt_var": bufs["impt_var"], "EF": bufs["EF"],
            "varF": bufs["varF"], "consensus_cluster": bufs["consensus_cluster"]}


def lq_fit(dat, sigma, pg, z, Ef, Varf, celltype=None, penl=1, cutoff=0.1, *, resp=None, resid_mode=0, impute=True,
           seed=0, sweep=0, comm=None, stream=None, out_Y=None):
    """The low-quality-gene step between the two ``EM_impute`` calls of ``SIMPLE()`` / ``SIMPLE_B()``
    (R/SIMPLE.R:305-411, R/SIMPLE_B.R:332-419) on B200: per gene M0 unpenalised intercepts + K0 lasso loadings,
    the shared variance, and one imputation draw of the entries <= cutoff with the factor at its posterior mean.
    All matrices hold the lq rows only.  ``resp`` = ``init_imp[lq_ind, ]`` when SIMPLE() got one; ``resid_mode``
    0 = SIMPLE() with ``init_imp = NULL`` (residual term dropped), 1 = residual of ``dat`` (SIMPLE() with
    ``init_imp``; SIMPLE_B()).  Returns mu, beta, sigma, Y."""
    lib = _lib.load()
    dat = _f(dat)
    Gq, n = dat.shape
    z = _f(z)
    M0 = z.shape[1]
    K0 = np.asarray(Ef[0]).shape[1]
    ct = None if celltype is None else np.ascontiguousarray(np.asarray(celltype).astype(np.int32))
    pg = _f(np.asarray(pg, dtype=np.float64).reshape(Gq, -1))
    keep = dict(dat=dat, resp=None if resp is None else _f(resp, (Gq, n), "resp"), sigma=_f(sigma, (Gq, M0), "sigma"),
                pg=pg, z=_f(z, (n, M0), "z"),
                Ef=np.ascontiguousarray(np.stack([_f(e, (n, K0), "Ef").T for e in Ef])),       # M0 col-major n x K0 blocks
                Varf=np.ascontiguousarray(np.stack([_f(v, (K0, K0), "Varf").T for v in Varf])),
                celltype=ct)
    a = _lib.LqArgs()
    a.Gq, a.n, a.M0, a.K0, a.MB = Gq, n, M0, K0, pg.shape[1]
    for k, v in keep.items():
        setattr(a, k, _ptr(v))
    a.penl, a.cutoff, a.resid_mode, a.impute = float(penl), float(cutoff), int(resid_mode), int(bool(impute))
    a.seed, a.sweep = int(seed) & 0xFFFFFFFFFFFFFFFF, int(sweep)
    cw = None
    if comm is not None:
        cw = _Comm(comm)
        a.n_total, a.cell_offset, a.allreduce = int(comm.n_total), int(comm.cell_offset), cw.cb
    a.stream = stream
    bufs = dict(mu=np.zeros((Gq, M0), order="F"), beta=np.zeros((Gq, K0), order="F"), sigma=np.zeros((Gq, M0), order="F"),
                Y=(np.empty((Gq, n), order="F") if out_Y is None else _check_out(out_Y, (Gq, n), "F", "Y")) if impute else None)
    o = _lib.LqOut()
    for k, v in bufs.items():
        setattr(o, k, _ptr(v))
    rc = lib.simples_lq_fit(C.byref(a), C.byref(o))
    if rc < 0 and cw is not None and cw.error is not None:
        raise cw.error
    check(rc)
    return bufs


def _init_call(Y2, M0, clus, pg1, p_min, cutoff, seed, sweep, comm, stream, details, out=None):
    lib = _lib.load()
    Y2 = _f(Y2)
    G, n = Y2.shape
    cl = np.ascontiguousarray(np.asarray(clus).astype(np.int32))
    if cl.shape != (n,):
        raise ValueError("clus must have length n")
    keep = dict(Y2=Y2, clus=cl, pg1=None if pg1 is None else _f(pg1, (G, M0), "pg1"))
    a = _lib.InitArgs()
    a.G, a.n, a.M0 = G, n, int(M0)
    for k, v in keep.items():
        setattr(a, k, _ptr(v))
    a.p_min, a.cutoff = float(p_min), float(cutoff)
    a.seed, a.sweep = int(seed) & 0xFFFFFFFFFFFFFFFF, int(sweep)
    cw = None
    if comm is not None:
        cw = _Comm(comm)
        a.n_total, a.cell_offset, a.allreduce = int(comm.n_total), int(comm.cell_offset), cw.cb
    a.stream = stream
    bufs = dict(impute=np.empty((G, n), order="F") if out is None else _check_out(out, (G, n), "F", "impute"),
                pg=np.empty((G, M0), order="F"),
                mu=np.empty((G, M0), order="F") if details else None, sd=np.empty((G, M0), order="F") if details else None)
    o = _lib.InitOut()
    for k, v in bufs.items():
        setattr(o, k, _ptr(v))
    rc = lib.simples_init_impute(C.byref(a), C.byref(o))
    if rc < 0 and cw is not None and cw.error is not None:
        raise cw.error
    check(rc)
    return bufs


def init_impute(Y2, M0, clus, p_min=0.6, cutoff=0.1, verbose=False, *, seed=0, sweep=0, comm=None, stream=None,
                details=False, out=None):
    """``init_impute`` (R/SIMPLE.R:15-73) on B200: per cluster and gene the zero-inflated censored-normal fit of
    ``ztruncnorm`` (R/fit_ztrunc.R:55-134), then one draw of every entry <= cutoff.  Returns ``[impute, pg]`` like the
    reference (``details=True`` appends the fitted per-cluster mean and sd)."""
    b = _init_call(Y2, M0, clus, None, p_min, cutoff, seed, sweep, comm, stream, details, out)
    return [b["impute"], b["pg"]] + ([b["mu"], b["sd"]] if details else [])


def init_impute_bulk(Y2, clus, bulk, pg1, cutoff=0.1, verbose=False, *, seed=0, sweep=0, comm=None, stream=None,
                     details=False, out=None):
    """``init_impute_bulk`` (R/SIMPLE_B.R:23-114) on B200; ``bulk`` only feeds the reference's verbose plots and is
    ignored.  Returns the imputed matrix (``details=True``: ``[impute0, pg, mu, sd]``)."""
    M0 = int(np.max(clus))
    b = _init_call(Y2, M0, clus, pg1, 0.01, cutoff, seed, sweep, comm, stream, details, out)
    return [b["impute"], b["pg"], b["mu"], b["sd"]] if details else b["impute"]


def bic(loglik_last, n, G, K0, M0):
    """(BIC, BIC0) of SIMPLE()/SIMPLE_B() (R/SIMPLE.R:429-430); ``n`` is the global number of cells."""
    b, b0 = C.c_double(), C.c_double()
    check(_lib.load().simples_bic(float(loglik_last), int(n), int(G), int(K0), int(M0), C.byref(b), C.byref(b0)))
    return b.value, b0.value
