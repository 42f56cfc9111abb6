"""NumPy restatement of SIMPLEs' EM imputation hot path (TEST INFRASTRUCTURE).

Follows, line by line, ``R/EM_impute.R:75-339`` and ``R/do_impute.R:29-160`` of
JunLiuLab/SIMPLEs (paths relative to /root/reference), including the quirks
listed in SURVEY.md section 8(c), and -- for the rows of SURVEY.md 8(f) -- the
stages around them: ``R/fit_ztrunc.R:2-134`` (ztruncnorm; optim's Brent is
restated from the published Netlib fmin), ``R/SIMPLE.R:15-73`` /
``R/SIMPLE_B.R:23-114`` (init_impute / init_impute_bulk), ``R/SIMPLE.R:258-266``
(clustering start), ``R/SIMPLE.R:305-411`` (lq-gene fit) and the drivers
``R/SIMPLE.R:200-455``, ``R/SIMPLE_B.R:229-455``, ``R/select_KM.R:126-189``.
PARITY UNPINNED (see ``oracle/__init__``).

Two M-step formulations are provided and cross-checked by the tests:

* ``em_impute_ref``  -- literal: per gene the ``(M0*n+K0) x K0`` augmented design
  ``W_aug``/``Y_aug`` of ``R/EM_impute.R:238-270`` and the dense ``G x G`` ``bigM``
  of ``R/EM_impute.R:311-314``.
* ``em_impute_fast`` -- cell-additive sufficient statistics (what the CUDA
  path computes and all-reduces across GPUs).

Third-party arithmetic on the path that is NOT in /root/reference (no version
pins, DESCRIPTION:14-24): glmnet (lasso, replaced by an exact minimiser of the
same objective), Rsolnp::solnp (replaced by a deterministic damped Newton
started from ``lam_init``), mixtools::rmvnorm (symmetric eigen square root),
msm::rtnorm / stats::rnorm / rbinom / rmultinom (replaced by inverse-CDF draws
from a Philox4x32-10 counter tape shared bit-for-bit with the CUDA kernels).
"""
from __future__ import annotations

import math

import numpy as np
from scipy import special

__all__ = [
    "PI_REF", "COMPAT_STALE_DT", "COMPAT_PI", "COMPAT_MU_DIV_N",
    "COMPAT_PRIORB_PREV", "COMPAT_DEFAULT",
    "philox4x32", "cell_uniforms", "entry_uniforms", "trunc_z", "sym_sqrt",
    "lasso_solve", "lasso_solve_batch", "lasso_glmnet_compat", "glmnet_band", "lambda_solve_k", "lambda_step",
    "cluster_terms", "epass", "factor_means", "sample_sweep",
    "local_stats", "stats_pack", "stats_unpack", "mstep_from_stats",
    "em_impute_ref", "em_impute_fast", "do_impute_ref",
    "simulate", "bench_state", "expected_impute", "yimp0_ref", "bic_ref",
    "lq_regress_ref", "lq_regress_fast", "lq_impute_ref",
    "brent_fmin", "ztruncnorm_ref", "init_impute_ref", "init_impute_bulk_ref",
    "simple_ref", "simple_b_ref", "select_km_ref", "init_cluster_ref", "adjusted_rand_index",
]

PI_REF = 3.14159265359          # R/EM_impute.R:80, R/do_impute.R:31  (Q2)

# ref_compat bits (default: all on = reproduce the reference's behaviour)
COMPAT_STALE_DT = 1             # Q1  R/EM_impute.R:206,212
COMPAT_PI = 2                   # Q2  R/EM_impute.R:80
COMPAT_MU_DIV_N = 4             # Q3  R/EM_impute.R:95
COMPAT_PRIORB_PREV = 8          # Q5  R/EM_impute.R:119,144-146,275
COMPAT_DEFAULT = 15

_U32 = np.uint64(0xFFFFFFFF)


# ----------------------------------------------------------------------------
# Counter-based variate tape (shared with csrc/rng.cuh)
# ----------------------------------------------------------------------------
def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Philox4x32-10 (Salmon et al. 2011).  All inputs broadcastable, uint32
    values carried in uint64 arrays.  Returns four uint64 arrays (< 2**32)."""
    m0 = np.uint64(0xD2511F53)
    m1 = np.uint64(0xCD9E8D57)
    w0 = np.uint64(0x9E3779B9)
    w1 = np.uint64(0xBB67AE85)
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3)])
    k0 = np.uint64(k0)
    k1 = np.uint64(k1)
    sh = np.uint64(32)
    for _ in range(rounds):
        p0 = m0 * c0
        p1 = m1 * c2
        hi0, lo0 = p0 >> sh, p0 & _U32
        hi1, lo1 = p1 >> sh, p1 & _U32
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + w0) & _U32
        k1 = (k1 + w1) & _U32
    return c0, c1, c2, c3


def _u52(hi, lo):
    """Two 32-bit words -> uniform in (0,1): ((x >> 12) + 0.5) * 2**-52."""
    x = (hi << np.uint64(32)) | lo
    return ((x >> np.uint64(12)).astype(np.float64) + 0.5) * (2.0 ** -52)


def _seed_key(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, seed >> 32


def cell_uniforms(seed, sweep, cells, nu):
    """Per-cell uniforms: stream tag 1, block j gives uniforms 2j, 2j+1.
    Uniform 0 = membership draw, uniform 1+k = factor normal k."""
    cells = np.asarray(cells, dtype=np.uint64)
    k0, k1 = _seed_key(seed)
    nb = (nu + 1) // 2
    out = np.empty((cells.shape[0], 2 * nb))
    for j in range(nb):
        x0, x1, x2, x3 = philox4x32(np.uint64(j), cells & _U32, np.uint64(sweep),
                                     ((cells >> np.uint64(32)) << np.uint64(8)) | np.uint64(1), k0, k1)
        out[:, 2 * j] = _u52(x0, x1)
        out[:, 2 * j + 1] = _u52(x2, x3)
    return out[:, :nu]


def entry_uniforms(seed, sweep, cells, genes):
    """Per-entry uniforms (u_a: Bernoulli dropout draw, u_b: value draw).
    Stream tag 0, counter (gene, cell_lo, sweep, cell_hi<<8)."""
    cells = np.asarray(cells, dtype=np.uint64)
    genes = np.asarray(genes, dtype=np.uint64)
    k0, k1 = _seed_key(seed)
    x0, x1, x2, x3 = philox4x32(genes, cells & _U32, np.uint64(sweep),
                                 (cells >> np.uint64(32)) << np.uint64(8), k0, k1)
    return _u52(x0, x1), _u52(x2, x3)


_SQRT2 = math.sqrt(2.0)
_SQRT2PI = math.sqrt(2.0 * math.pi)
TRUNC_LOG_BRANCH = -30.0


def trunc_z(u, r):
    """Inverse-CDF draw of a standard normal truncated to (-inf, r]:
    z with Phi(z) = u * Phi(r).  Replaces msm::rtnorm(upper=cutoff)
    (R/EM_impute.R:191-192, R/do_impute.R:130-131) -- same distribution, one
    uniform per draw.  For r < -30 the product underflows, so the equation is
    solved in log space by 4 Newton steps on log Phi (erfcx form)."""
    u = np.asarray(u, dtype=np.float64)
    r = np.asarray(r, dtype=np.float64)
    u, r = np.broadcast_arrays(np.atleast_1d(u), np.atleast_1d(r))
    scalar = (np.ndim(u) == 1 and u.shape == (1,))
    z = special.ndtri(u * special.ndtr(r))
    deep = r < TRUNC_LOG_BRANCH
    if np.any(deep):
        rd = r[deep]
        ud = u[deep]
        # log Phi(x) = log(erfcx(-x/sqrt2)/2) - x^2/2
        t = np.log(ud) + np.log(0.5 * special.erfcx(-rd / _SQRT2)) - 0.5 * rd * rd
        zd = -np.sqrt(rd * rd - 2.0 * np.log(ud))
        for _ in range(4):
            e = 0.5 * special.erfcx(-zd / _SQRT2)
            zd = zd - (np.log(e) - 0.5 * zd * zd - t) * e * _SQRT2PI
        z = z.copy()
        z[deep] = zd
    out = np.minimum(z, r)
    return out


def sym_sqrt(M):
    """Symmetric PSD square root (mixtools::rmvnorm uses V diag(sqrt(ev)) V^T)."""
    w, V = np.linalg.eigh(0.5 * (M + M.T))
    return (V * np.sqrt(np.maximum(w, 0.0))) @ V.T


# ----------------------------------------------------------------------------
# Lasso (replaces glmnet, R/EM_impute.R:256-261): exact minimiser of
#   0.5 b'Ab - rhs'b + kappa |b|_1
# ----------------------------------------------------------------------------
def _polish(A, rhs, kappa, beta):
    """kappa: scalar or per-coordinate vector (0 = unpenalised coordinate, always active)."""
    kv = np.broadcast_to(np.asarray(kappa, dtype=np.float64), beta.shape)
    S = (beta != 0.0) | (kv == 0.0)
    if not S.any():
        return beta if np.all(np.abs(rhs) <= kv * (1 + 1e-12)) else None
    sgn = np.sign(beta[S])
    try:
        x = np.linalg.solve(A[np.ix_(S, S)], rhs[S] - kv[S] * sgn)
    except np.linalg.LinAlgError:
        return None
    pen = kv[S] > 0
    if np.any(np.sign(x[pen]) != sgn[pen]):
        return None
    full = np.zeros_like(beta)
    full[S] = x
    grad = rhs - A @ full
    if np.any(np.abs(grad[~S]) > kv[~S] * (1 + 1e-10) + 1e-300):
        return None
    return full


def lasso_solve_batch(A, rhs, kappa, max_sweeps=20000, tol=1e-15):
    """A: (K,K) shared or (G,K,K); rhs (G,K); kappa (G,) or (G,K) (per-coordinate
    penalties, 0 = unpenalised).  Cyclic coordinate descent from 0 (glmnet's
    start), then an exact active-set polish."""
    rhs = np.asarray(rhs, dtype=np.float64)
    G, K = rhs.shape
    shared = (A.ndim == 2)
    kappa = np.asarray(kappa, dtype=np.float64)
    kappa = np.broadcast_to(kappa[:, None] if kappa.ndim == 1 else kappa, (G, K))
    beta = np.zeros((G, K))
    r = rhs.copy()
    live = np.arange(G)
    for sweep in range(max_sweeps):
        Al = A if shared else A[live]
        bl = beta[live]
        rl = r[live]
        kl = kappa[live]
        maxd = np.zeros(live.shape[0])
        for k in range(K):
            akk = Al[k, k] if shared else Al[:, k, k]
            t = rl[:, k] + akk * bl[:, k]
            bn = np.sign(t) * np.maximum(np.abs(t) - kl[:, k], 0.0) / akk
            d = bn - bl[:, k]
            col = Al[:, k] if shared else Al[:, :, k]
            rl -= d[:, None] * col
            bl[:, k] = bn
            maxd = np.maximum(maxd, np.abs(d))
        beta[live] = bl
        r[live] = rl
        scale = np.maximum(1.0, np.abs(bl).max(axis=1))
        done = maxd <= tol * scale
        if sweep >= 3 and (sweep % 4 == 3 or done.all()):
            # try the exact polish; genes that verify leave the live set
            keep = np.ones(live.shape[0], dtype=bool)
            for j, g in enumerate(live):
                Ag = A if shared else A[g]
                p = _polish(Ag, rhs[g], kappa[g], beta[g])
                if p is not None:
                    beta[g] = p
                    keep[j] = False
                elif done[j]:
                    keep[j] = False
            live = live[keep]
            if live.size == 0:
                break
    return beta


def lasso_solve(A, rhs, kappa):
    kap = np.asarray(kappa, dtype=np.float64)
    kap = kap[None] if kap.ndim == 0 else kap[None, :]
    return lasso_solve_batch(np.asarray(A), np.asarray(rhs)[None, :], kap)[0]


# ----------------------------------------------------------------------------
# glmnet-compatible stopping (TEST ONLY): what the reference's own lasso call returns
# ----------------------------------------------------------------------------
LASSO_MODE = "exact"            # "exact" | "glmnet": which solver mstep_from_stats / lq_regress_fast use
GLMNET_THRESH = 1e-7            # glmnet's default `thresh` (R/EM_impute.R:257-259 passes none)


def lasso_glmnet_compat(A, rhs, kappa, N, sum_y, sum_y2, thresh=GLMNET_THRESH, maxit=100000, info=None, vp=None):
    """The lasso of R/EM_impute.R:257-259 as glmnet itself solves it, from the Gram form.

    glmnet (>= 2.0-16, ``family="gaussian"``, ``standardize=F``, ``intercept=F``, one user
    ``lambda``, fewer than 500 variables => covariance updating) is not in /root/reference;
    this restates its published algorithm (Friedman, Hastie, Tibshirani 2010, J. Stat.
    Softw. 33(1) section 2.2, and the ``elnet1`` loop of the package's Fortran source):

    * observation weights ``1/N``; the response is scaled by its *population* standard
      deviation ``ys`` (also without an intercept) and the user ``lambda`` divided by it;
    * cyclic coordinate descent from 0 on the covariance form ``g_k = <x_k, r>``,
      ``xv_k = <x_k, x_k>``; strong-rule screening ``|g_k| > 2 lambda`` with a KKT pass over
      the excluded coordinates; after every full pass that changed something, passes over
      the active set only until converged;
    * converged when the largest ``xv_k * delta_k^2`` of a pass is below ``thresh``.

    ``A`` = X'X (K x K), ``rhs`` = X'y, ``kappa`` = N * lambda_user (the penalty on the
    0.5 * RSS scale, as in ``lasso_solve``); ``N, sum_y, sum_y2`` describe the response.
    ``vp`` = per-coordinate penalty factors AFTER glmnet's rescale to sum K (None = all 1;
    0 = unpenalised, as the cluster intercepts of R/SIMPLE.R:343-349).
    Returns the coefficients on the original scale.  ``info`` (dict) receives ``passes``.
    """
    A = np.asarray(A, dtype=np.float64)
    K = A.shape[0]
    ys = math.sqrt(max(sum_y2 / N - (sum_y / N) ** 2, 0.0))
    if not ys > 0.0:
        return np.zeros(K)
    C = A / N
    xv = np.diag(C).copy()
    g = np.asarray(rhs, dtype=np.float64) / N / ys
    alm = float(kappa) / N / ys
    vp = np.ones(K) if vp is None else np.asarray(vp, dtype=np.float64)
    a = np.zeros(K)
    ix = np.abs(g) > 2.0 * alm * vp                  # strong rule with alm0 = 0 (single user lambda)
    active = np.zeros(K, dtype=bool)
    nlp = 0

    def update(k, cols):
        ak = a[k]
        u = g[k] + ak * xv[k]
        v = abs(u) - alm * vp[k]
        nk = math.copysign(v, u) / xv[k] if v > 0.0 else 0.0
        if nk == ak:
            return 0.0
        a[k] = nk
        d = nk - ak
        g[cols] -= C[cols, k] * d
        return xv[k] * d * d

    while True:
        # full pass over the screened set
        nlp += 1
        dlx = 0.0
        for k in range(K):
            if not ix[k]:
                continue
            before = a[k]
            dk = update(k, ix)
            if a[k] != before:
                active[k] = True
            dlx = max(dlx, dk)
        if dlx < thresh:
            # KKT pass over the excluded coordinates (their gradient was not maintained)
            excl = ~ix
            if excl.any():
                gex = np.asarray(rhs, dtype=np.float64)[excl] / N / ys - C[np.ix_(excl, np.ones(K, bool))] @ a
                viol = np.abs(gex) > alm * vp[excl]
                if viol.any():
                    idx = np.flatnonzero(excl)[viol]
                    ix[idx] = True
                    g[idx] = gex[viol]
                    continue
            break
        if nlp > maxit:
            raise RuntimeError("glmnet-compat lasso: maxit reached")
        # passes over the active set only; gradients of the other screened coordinates are refreshed afterwards
        da = a.copy()
        while True:
            nlp += 1
            dlx = 0.0
            for k in np.flatnonzero(active):
                dlx = max(dlx, update(k, active))
            if dlx < thresh:
                break
            if nlp > maxit:
                raise RuntimeError("glmnet-compat lasso: maxit reached")
        da = a - da
        rest = ix & ~active
        if rest.any():
            g[rest] -= C[np.ix_(rest, active)] @ da[active]
    if info is not None:
        info["passes"] = nlp
    return a * ys


def glmnet_band(A, kappa, N, sum_y, sum_y2, beta_compat, thresh=GLMNET_THRESH):
    """Bound on |beta_exact - beta_compat| implied by glmnet's stopping rule (same support).

    At the stop every coordinate changed by less than sqrt(thresh / xv_k) (scaled units) in the last pass, so
    the stationarity residual on the support S is at most R = max_k sum_j |C_kj| sqrt(thresh / xv_j), and the
    exact minimiser on S differs by at most ||C_SS^-1||_inf * R (times ys on the original scale).
    """
    A = np.asarray(A, dtype=np.float64)
    ys = math.sqrt(max(sum_y2 / N - (sum_y / N) ** 2, 0.0))
    S = np.asarray(beta_compat) != 0.0
    if not S.any() or not ys > 0.0:
        return 0.0
    C = A[np.ix_(S, S)] / N
    step = np.sqrt(thresh / np.diag(C))
    R = float(np.max(np.abs(C) @ step))
    return ys * float(np.max(np.sum(np.abs(np.linalg.inv(C)), axis=1))) * R


# ----------------------------------------------------------------------------
# lambda step (replaces Rsolnp::solnp, R/EM_impute.R:278-304)
# ----------------------------------------------------------------------------
def lambda_solve_k(E, c, x0, T, max_it=500):
    """min sum(E/x + c log x)  s.t. sum x = T, x > 0, from x0 (sum x0 = T).
    Damped Newton on the equality-constrained problem with |f''| curvature
    and Armijo backtracking.  Mirrored operation-for-operation by
    csrc/mstep.cu::lambda_solve_k."""
    E = np.asarray(E, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    x = np.array(x0, dtype=np.float64)
    mt = x.shape[0]

    def F(v):
        s = 0.0
        for m in range(mt):
            s += E[m] / v[m] + c[m] * math.log(v[m])
        return s

    f = F(x)
    for _ in range(max_it):
        g = np.empty(mt)
        hm = np.empty(mt)
        sg = 0.0
        sh = 0.0
        for m in range(mt):
            xi = x[m]
            g[m] = -E[m] / (xi * xi) + c[m] / xi
            h = 2.0 * E[m] / (xi * xi * xi) - c[m] / (xi * xi)
            h = abs(h)
            if h < 1e-300:
                h = 1e-300
            hm[m] = h
            sg += g[m] / h
            sh += 1.0 / h
        nu = -sg / sh
        dx = np.empty(mt)
        gd = 0.0
        amax = 1.0
        dmax = 0.0
        for m in range(mt):
            dx[m] = -(g[m] + nu) / hm[m]
            gd += g[m] * dx[m]
            if dx[m] < 0.0:
                a = -0.9 * x[m] / dx[m]
                if a < amax:
                    amax = a
            if abs(dx[m]) > dmax:
                dmax = abs(dx[m])
        if dmax <= 1e-15 * T:
            break
        a = amax
        accepted = False
        for _bt in range(60):
            xn = x + a * dx
            fn = F(xn)
            if fn <= f + 1e-4 * a * gd + 1e-14 * abs(f):
                accepted = True
                break
            a *= 0.5
        if not accepted:
            break
        x = xn
        f = fn
    return x


def lambda_step(Em_full, nz, beta, lam, M0, K0):
    """R/EM_impute.R:279-303.  Em_full: (K0, M0) with
    Em[k,m] = sum_i z_im W_m[i,k]^2 + M_m[k,k] nz_m."""
    lam = lam.copy()
    ind = np.nonzero((nz > K0 + 1) & (nz > 3))[0]                 # :279
    mt = ind.shape[0]
    if mt > 1:                                                    # :281
        Em = Em_full[:, ind]                                      # (K0, mt)   :282-285
        not_ind = np.setdiff1d(np.arange(M0), ind)
        lam[not_ind, :] = 1.0 / M0                                # :288
        lam_init = (Em + 1.0).T / (nz[ind] + 1.0)[:, None]        # (mt, K0)   :289
        lam_init = lam_init / lam_init.sum(axis=0)[None, :] * mt / M0   # :290
        for k in range(K0):                                       # :292
            if np.all(beta[:, k] == 0):                           # :293
                lam[:, k] = 1.0 / M0
            else:
                lam[ind, k] = lambda_solve_k(Em[k, :], nz[ind] - 2.0, lam_init[:, k], mt / M0)
    else:
        lam = np.full((M0, K0), 1.0 / M0)                         # :302
    return lam


# ----------------------------------------------------------------------------
# E-pass pieces
# ----------------------------------------------------------------------------
def cluster_terms(beta, sigma, lam, m):
    """R/EM_impute.R:123-125 (= R/do_impute.R:70-72)."""
    beta2 = beta / sigma[:, m][:, None]
    Mm = np.linalg.inv(beta.T @ beta2 + np.diag(1.0 / lam[m]))
    sgn, logdet = np.linalg.slogdet(Mm)
    dt = np.sum(np.log(sigma[:, m])) + np.sum(np.log(lam[m])) - logdet
    return beta2, Mm, dt


def epass(Yc, beta, sigma, lam, mu, pi, pi_const=PI_REF, M_list=None, beta2_fix=None, dt_fix=None):
    """One E-pass: loglik (n,M0) incl. log pi, lik, sconst.  With
    ``beta2_fix``/``dt_fix`` it is the MC-loop pass of R/EM_impute.R:202-213
    (stale beta2, dt of cluster M0 -- Q1) reusing ``M_list``."""
    G, n = Yc.shape
    M0 = mu.shape[1]
    ll = np.empty((n, M0))
    Ms, b2s, dts = [], [], []
    for m in range(M0):
        if M_list is None:
            beta2, Mm, dt = cluster_terms(beta, sigma, lam, m)
        else:
            Mm = M_list[m]
            beta2, dt = beta2_fix, dt_fix
        R = Yc - mu[:, m][:, None]
        y = np.sum(R / sigma[:, m][:, None] * R, axis=0)                      # :128
        x2 = R.T @ beta2                                                     # :129
        y = y - np.sum((x2 @ Mm) * x2, axis=1)                               # :130
        ll[:, m] = (-y - dt - G * math.log(2 * pi_const)) / 2                # :135
        Ms.append(Mm)
        b2s.append(beta2)
        dts.append(dt)
    ll = ll + np.log(pi)[None, :]                                            # :138
    sconst = ll.max(axis=1)                                                  # :139
    lik = np.exp(ll - sconst[:, None])                                       # :140-141
    return ll, lik, sconst, Ms, b2s, dts


def factor_means(Yc, beta, sigma, mu, Ms):
    """R/EM_impute.R:155-157: Wm[[m]] = t(M_m B' ((Y-mu_m)/sigma_m)), n x K0."""
    return [(Ms[m] @ beta.T @ ((Yc - mu[:, m][:, None]) / sigma[:, m][:, None])).T
            for m in range(len(Ms))]


def sample_sweep(Yc, mask, z, Wm, Ms, mu, beta, sigma, pg, celltype0, gene_mean, gene_sd,
                 cutoff, seed, sweep, cell0=0, draw_membership=True, lower=-np.inf,
                 upper=np.inf, clip=True, expect=False):
    """One MC imputation sweep, R/EM_impute.R:163-199 / R/do_impute.R:107-136,
    with every random variate taken from the Philox tape.  Modifies Yc in
    place; returns the sampled membership (0-based) and factors."""
    G, n = Yc.shape
    M0 = len(Ms)
    K0 = beta.shape[1]
    cells = np.arange(n, dtype=np.uint64) + np.uint64(cell0)
    U = cell_uniforms(seed, sweep, cells, 1 + K0)
    if expect:
        # deterministic companion (north_star "conditional-expectation imputation"): most probable cluster,
        # factor at its posterior mean, every zero entry at its conditional expectation -- no variate is consumed
        im = np.argmax(z, axis=1)
        f = np.stack([Wm[m][i] for i, m in enumerate(im)]) if n else np.empty((0, K0))
        gi, ci = np.nonzero(mask)
        mm = im[ci]
        ms = (mu[gi, mm] + np.einsum("ek,ek->e", beta[gi], f[ci])) * gene_sd[gi] + gene_mean[gi]
        sds = np.sqrt(sigma[gi, mm]) * gene_sd[gi]
        x = expected_impute(ms, sds, pg[gi, celltype0[ci]], cutoff)
        if clip:
            x = np.maximum(np.minimum(x, upper), lower)
        Yc[gi, ci] = (x - gene_mean[gi]) / gene_sd[gi]
        return im, f
    if draw_membership:
        cum = np.cumsum(z, axis=1)
        im = np.minimum((U[:, :1] >= cum).sum(axis=1), M0 - 1)               # :164
    else:
        im = np.zeros(n, dtype=np.int64)                                     # :166
    eps = special.ndtri(U[:, 1:1 + K0])
    Mh = [sym_sqrt(Mm) for Mm in Ms]
    f = np.empty((n, K0))
    for m in range(M0):
        sel = im == m
        if sel.any():
            f[sel] = Wm[m][sel] + eps[sel] @ Mh[m]                            # :171
    gi, ci = np.nonzero(mask)                                                # :173
    mm = im[ci]
    ms = (mu[gi, mm] + np.einsum("ek,ek->e", beta[gi], f[ci])) * gene_sd[gi] + gene_mean[gi]   # :175
    sds = np.sqrt(sigma[gi, mm]) * gene_sd[gi]                               # :176
    p = pg[gi, celltype0[ci]]                                                # :177
    r = (cutoff - ms) / sds
    prob = special.ndtr(r)                                                   # :179
    prob_drop = (1 - p) / (prob * p + (1 - p))                               # :180
    ua, ub = entry_uniforms(seed, sweep, cells[ci], gi)
    drop = ua < prob_drop                                                    # :181
    x = np.where(drop, ms + sds * special.ndtri(ub),                         # :186
                 ms + sds * trunc_z(ub, r))                                  # :191-192
    if clip:
        x = np.minimum(x, upper)                                             # :195
        x = np.maximum(x, lower)                                             # :196
    Yc[gi, ci] = (x - gene_mean[gi]) / gene_sd[gi]                           # :198
    return im, f


def expected_impute(ms, sds, p, cutoff):
    """Closed-form E[x] of one zero entry (north_star's 'conditional expectation'; hinted at in R/SIMPLE.R:59):
    E[x] = pd ms + (1 - pd) (ms - sds phi(r)/Phi(r)) with pd = (1-p)/(Phi(r) p + 1-p), r = (cutoff - ms)/sds.
    Written as ms - sds p phi(r) / (p Phi(r) + 1 - p): the same number without the 0/0 of the Mills ratio in the tail."""
    r = (cutoff - ms) / sds
    prob = special.ndtr(r)
    return ms - sds * p * (np.exp(-0.5 * r * r) / _SQRT2PI) / (p * prob + (1 - p))


# ----------------------------------------------------------------------------
# Sufficient statistics (cell-additive => one all-reduce per EM iteration)
# ----------------------------------------------------------------------------
def local_stats(Yc, z, Wm, shared_sigma):
    """Statistics of SURVEY.md section 3.3 for one shard of cells."""
    G, n = Yc.shape
    M0 = z.shape[1]
    K0 = Wm[0].shape[1]
    st = {}
    st["nz"] = z.sum(axis=0)
    st["c"] = np.sqrt(z).sum(axis=0)
    st["S"] = np.stack([(Wm[m] * z[:, m][:, None]).T @ Wm[m] for m in range(M0)])
    st["a"] = np.stack([(Wm[m] * z[:, m][:, None]).sum(axis=0) for m in range(M0)])
    st["U"] = Yc @ z
    st["Uh"] = Yc @ np.sqrt(z)
    if shared_sigma:
        Wbar = sum(Wm[m] * z[:, m][:, None] for m in range(M0))
        st["Tbar"] = Yc @ Wbar                                   # sum_m T_m
        st["R2"] = (Yc * Yc) @ z.sum(axis=1)                     # sum_m U2[,m]
    else:
        st["T"] = np.stack([Yc @ (Wm[m] * z[:, m][:, None]) for m in range(M0)])
        st["U2"] = (Yc * Yc) @ z
    return st


_ORDER_SHARED = ("Tbar", "U", "Uh", "R2", "S", "a", "nz", "c")
_ORDER_GENERAL = ("T", "U", "Uh", "U2", "S", "a", "nz", "c")


def stats_pack(st):
    order = _ORDER_SHARED if "Tbar" in st else _ORDER_GENERAL
    return np.concatenate([np.ravel(st[k]) for k in order])


def stats_unpack(buf, G, M0, K0, shared_sigma):
    order = _ORDER_SHARED if shared_sigma else _ORDER_GENERAL
    shapes = {"Tbar": (G, K0), "T": (M0, G, K0), "U": (G, M0), "Uh": (G, M0), "R2": (G,),
              "U2": (G, M0), "S": (M0, K0, K0), "a": (M0, K0), "nz": (M0,), "c": (M0,)}
    st, o = {}, 0
    for k in order:
        sz = int(np.prod(shapes[k]))
        st[k] = buf[o:o + sz].reshape(shapes[k])
        o += sz
    return st


def mstep_from_stats(st, n, Ms, beta, sigma, lam, mu, M0, K0, it, penl, max_lambda, est_lam,
                     sigma0, pi_alpha):
    """M-step of R/EM_impute.R:234-318 from global statistics (n = total cells)."""
    G = beta.shape[0]
    nz = st["nz"]
    C = np.stack([st["S"][m] + nz[m] * Ms[m] for m in range(M0)])            # S_m + Vm_m  (:234)
    N = M0 * n + K0
    shared = "Tbar" in st
    D = (st["U2"] if not shared else None)
    if shared:
        sg0 = sigma[:, 0]
        A = C.sum(axis=0)                                                    # Gram * sigma_g
        bt = st["Tbar"] - mu @ st["a"]                                       # sum_m (T_m - mu a_m)
        d2 = st["R2"] + np.sum(-2 * mu * st["U"] + mu * mu * nz[None, :], axis=1)
        sumY = np.sum(st["Uh"] - mu * st["c"][None, :], axis=1) / np.sqrt(sg0)
        sumY2 = d2 / sg0
        var = (sumY2 - sumY * sumY / N) / (N - 1)                            # var(Y_aug)  (Q7)
        kappa = penl * var                                                   # :258
        if LASSO_MODE == "glmnet":      # what glmnet's own stopping rule returns (test only)
            nb = np.stack([lasso_glmnet_compat(A / sg0[g], bt[g] / sg0[g], kappa[g], N, sumY[g], sumY2[g])
                           for g in range(G)])
        else:
            nb = lasso_solve_batch(A, bt, kappa * sg0)
        quad = np.einsum("gj,jk,gk->g", nb, A, nb)
        sg = (d2 - 2 * np.sum(nb * bt, axis=1) + quad + 1) / (n + 3)         # :263-267
    else:
        D2 = D - 2 * mu * st["U"] + mu * mu * nz[None, :]                    # (G,M0)
        bm = st["T"] - mu.T[:, :, None] * st["a"][:, None, :]                # (M0,G,K0)
        A = np.einsum("mjk,gm->gjk", C, 1.0 / sigma)
        rhs = np.einsum("mgk,gm->gk", bm, 1.0 / sigma)
        sumY = np.sum((st["Uh"] - mu * st["c"][None, :]) / np.sqrt(sigma), axis=1)
        sumY2 = np.sum(D2 / sigma, axis=1)
        var = (sumY2 - sumY * sumY / N) / (N - 1)
        kappa = penl * var
        if LASSO_MODE == "glmnet":
            nb = np.stack([lasso_glmnet_compat(A[g], rhs[g], kappa[g], N, sumY[g], sumY2[g]) for g in range(G)])
        else:
            nb = lasso_solve_batch(A, rhs, kappa)
        quad = np.einsum("gj,mjk,gk->g", nb, C, nb)
        sg = (D2.sum(axis=1) - 2 * np.einsum("gk,mgk->g", nb, bm) + quad + 1) / (n + 3)
    priorB = float(np.sum(penl * var / sg * np.abs(nb).sum(axis=1)))         # :268,275
    beta = nb
    sigma = np.repeat(np.clip(sg, 1e-4, 9.0)[:, None], M0, axis=1)           # :269,273-274
    if max_lambda and it >= est_lam:                                         # :278
        Em = np.stack([np.diag(st["S"][m]) + np.diag(Ms[m]) * nz[m] for m in range(M0)], axis=1)
        lam = lambda_step(Em, nz, beta, lam, M0, K0)
    mu = mu.copy()
    for m in range(M0):                                                      # :307-315 (Woodbury, no GxG)
        s = nz[m] + sigma[:, m] / sigma0
        b2 = beta / s[:, None]
        v = st["U"][:, m]
        Cm = np.diag(sigma0 / lam[m]) + b2.T @ beta
        mu[:, m] = v / s - b2 @ np.linalg.solve(Cm, b2.T @ v)
    pi = (nz + pi_alpha - 1) / (n + pi_alpha * M0 - M0)                      # :318
    return beta, sigma, lam, mu, pi, priorB


# ----------------------------------------------------------------------------
# EM_impute
# ----------------------------------------------------------------------------
def _prologue(Y, z, mu, M0, lower, upper, compat):
    Y = np.array(Y, dtype=np.float64, order="F")
    Y = np.minimum(Y, upper)                                                 # :82
    Y = np.maximum(Y, lower)                                                 # :83
    G, n = Y.shape
    gene_mean = Y.mean(axis=1)                                               # :88
    gene_sd = np.ones(G)                                                     # :89
    Yc = Y - gene_mean[:, None]                                              # :90
    if mu is None:
        if M0 > 1 and z is not None:
            mu = Yc @ z / (n if compat & COMPAT_MU_DIV_N else 1.0)           # :95  (Q3)
            if not compat & COMPAT_MU_DIV_N:
                nzc = z.sum(axis=0)
                mu = mu / np.where(nzc != 0.0, nzc, 1.0)[None, :]               # empty cluster: column stays 0
        else:
            mu = np.zeros((G, M0))                                           # :97
    return Yc, gene_mean, gene_sd, np.array(mu, dtype=np.float64)


def _mstep_literal(Yc, z, Wm, Ms, nz, beta, sigma, lam, mu, M0, K0, it, penl, max_lambda,
                   est_lam, sigma0, pi_alpha):
    """R/EM_impute.R:234-318 with the explicit augmented design."""
    G, n = Yc.shape
    Vm = [Ms[m] * nz[m] for m in range(M0)]                                  # :234
    nbs = np.empty((G, K0))
    sgs = np.empty(G)
    prs = np.empty(G)
    sq = np.sqrt(z)
    for g in range(G):                                                       # :238
        V = sum(Vm[m] / sigma[g, m] for m in range(M0))                      # :240
        Wt = np.concatenate([Wm[m] * sq[:, m][:, None] / math.sqrt(sigma[g, m]) for m in range(M0)])  # :247
        Yt = np.concatenate([(Yc[g] - mu[g, m]) * sq[:, m] / math.sqrt(sigma[g, m]) for m in range(M0)])  # :245
        ML = np.linalg.cholesky(V).T                                         # :250 (upper)
        W_aug = np.concatenate([Wt, ML])                                     # :252
        Y_aug = np.concatenate([Yt, np.zeros(K0)])                           # :253
        kappa = penl * np.var(Y_aug, ddof=1)                                 # :258 (x N of glmnet's lambda)
        nb = lasso_solve(W_aug.T @ W_aug, W_aug.T @ Y_aug, kappa)            # :257-261
        sg = 0.0
        for m in range(M0):                                                  # :263-266
            sg += np.sum((Yc[g] - mu[g, m] - Wm[m] @ nb) ** 2 * z[:, m]) + nb @ Vm[m] @ nb
        sg = (sg + 1) / (n + 3)                                              # :267
        prs[g] = penl * np.var(Y_aug, ddof=1) / sg * np.sum(np.abs(nb))      # :268
        nbs[g] = nb
        sgs[g] = sg
    beta = nbs                                                               # :271
    sigma = np.repeat(np.clip(sgs, 1e-4, 9.0)[:, None], M0, axis=1)          # :272-274
    priorB = float(prs.sum())                                                # :275
    if max_lambda and it >= est_lam:                                         # :278
        Em = np.stack([np.sum(Wm[m] ** 2 * z[:, m][:, None], axis=0) + np.diag(Ms[m]) * nz[m]
                       for m in range(M0)], axis=1)                          # :282-285
        lam = lambda_step(Em, nz, beta, lam, M0, K0)
    mu = mu.copy()
    for m in range(M0):                                                      # :307-315
        st = nz[m] + sigma[:, m] / sigma0
        b2 = beta / st[:, None]
        bigM = np.diag(1.0 / st) - b2 @ np.linalg.solve(np.diag(sigma0 / lam[m]) + b2.T @ beta, b2.T)
        mu[:, m] = bigM @ (Yc @ z[:, m])
    pi = (nz + pi_alpha - 1) / (n + pi_alpha * M0 - M0)                      # :318
    return beta, sigma, lam, mu, pi, priorB


def _em_impute(Y, Y0, pg, M0, K0, cutoff, iter, beta, sigma, lam, pi, z, mu, celltype, penl,
               est_z, max_lambda, est_lam, impt_it, sigma0, pi_alpha, num_mc, lower, upper,
               seed, compat, literal, shards=None, trace=None, cell0=0, expect=False, accel=None):
    # accel: oracle.accel.Accel() swaps the O(n G) loops for their multi-threaded C / BLAS build (same arithmetic up to
    # rounding; the CPU baseline of bench.py), None = the NumPy statements below
    epass, factor_means, sample_sweep, local_stats = (
        (globals()[k] for k in ("epass", "factor_means", "sample_sweep", "local_stats")) if accel is None else
        (accel.epass, accel.factor_means, accel.sample_sweep, accel.local_stats))
    pi_const = PI_REF if compat & COMPAT_PI else math.pi
    z = None if z is None else np.array(z, dtype=np.float64)
    Yc, gene_mean, gene_sd, mu = _prologue(Y, z, mu, M0, lower, upper, compat)
    G, n = Yc.shape
    beta = np.array(beta, dtype=np.float64)
    sigma = np.array(sigma, dtype=np.float64)
    lam = np.array(lam, dtype=np.float64)
    pi = np.array(pi, dtype=np.float64)
    pg = np.asarray(pg, dtype=np.float64)
    if pg.ndim == 1:
        pg = pg[:, None]
    celltype0 = np.zeros(n, dtype=np.int64) if celltype is None else np.asarray(celltype, dtype=np.int64) - 1  # :101-102
    mask = np.asarray(Y0) <= cutoff
    tot = np.zeros(iter)
    priorB = 0.0                                                             # :119
    sweep = 0
    Wm, Ms = None, None
    for it in range(1, iter + 1):                                            # :120
        ll, lik, sconst, Ms, b2s, dts = epass(Yc, beta, sigma, lam, mu, pi, pi_const)
        tot[it - 1] = np.sum(np.log(lik.sum(axis=1)) + sconst)               # :144
        tot[it - 1] -= priorB if compat & COMPAT_PRIORB_PREV else 0.0        # :146
        if it >= est_z or z is None:                                         # :148
            z = lik / lik.sum(axis=1)[:, None]
        nz = z.sum(axis=0)                                                   # :152
        Wm = factor_means(Yc, beta, sigma, mu, Ms)                           # :155-157
        if it >= impt_it + 1:                                                # :161
            for _ in range(num_mc):                                          # :162
                sample_sweep(Yc, mask, z, Wm, Ms, mu, beta, sigma, pg, celltype0, gene_mean,
                             gene_sd, cutoff, seed, sweep, cell0, draw_membership=(M0 > 1),
                             lower=lower, upper=upper, clip=True, expect=expect)
                sweep += 1
                if compat & COMPAT_STALE_DT:                                 # :202-213  (Q1)
                    ll, lik, sconst, _, _, _ = epass(Yc, beta, sigma, lam, mu, pi, pi_const,
                                                     M_list=Ms, beta2_fix=b2s[M0 - 1], dt_fix=dts[M0 - 1])
                else:
                    ll, lik, sconst, _, _, _ = epass(Yc, beta, sigma, lam, mu, pi, pi_const)
                if it >= est_z or z is None:                                 # :220
                    z = lik / lik.sum(axis=1)[:, None]
                nz = z.sum(axis=0)                                           # :224
                Wm = factor_means(Yc, beta, sigma, mu, Ms)                   # :227-229
        if trace is not None:
            trace.append(dict(it=it, z=z.copy(), Wm=[w.copy() for w in Wm], Ms=[m.copy() for m in Ms],
                              Yc=Yc.copy(), tot=tot[it - 1]))
        if literal:
            beta, sigma, lam, mu, pi, priorB = _mstep_literal(
                Yc, z, Wm, Ms, nz, beta, sigma, lam, mu, M0, K0, it, penl, max_lambda, est_lam,
                sigma0, pi_alpha)
        else:
            shared = bool(np.all(sigma == sigma[:, :1]))
            bounds = [0, n] if shards is None else list(shards)
            buf = None
            for a, b in zip(bounds[:-1], bounds[1:]):
                st = local_stats(Yc[:, a:b], z[a:b], [w[a:b] for w in Wm], shared)
                buf = stats_pack(st) if buf is None else buf + stats_pack(st)   # the all-reduce
            st = stats_unpack(buf, G, M0, K0, shared)
            beta, sigma, lam, mu, pi, priorB = mstep_from_stats(
                st, n, Ms, beta, sigma, lam, mu, M0, K0, it, penl, max_lambda, est_lam, sigma0,
                pi_alpha)
    Yout = Yc * gene_sd[:, None] + gene_mean[:, None]                        # :335
    return dict(loglik=tot, pi=pi, mu=mu, sigma=sigma, beta=beta, **{"lambda": lam}, z=z, Ef=Wm,
                Varf=Ms, Y=Yout, geneM=gene_mean, geneSd=gene_sd, priorB=priorB)   # :337-338


def em_impute_ref(Y, Y0, pg, M0, K0, cutoff, iter, beta, sigma, lam, pi, z, mu=None, celltype=None,
                  penl=1, est_z=2, max_lambda=True, est_lam=2, impt_it=5, sigma0=100, pi_alpha=1,
                  num_mc=3, lower=-np.inf, upper=np.inf, seed=0, compat=COMPAT_DEFAULT, trace=None):
    """Literal restatement of EM_impute (R/EM_impute.R:75-339)."""
    return _em_impute(Y, Y0, pg, M0, K0, cutoff, iter, beta, sigma, lam, pi, z, mu, celltype, penl,
                      est_z, max_lambda, est_lam, impt_it, sigma0, pi_alpha, num_mc, lower, upper,
                      seed, compat, literal=True, trace=trace)


def em_impute_fast(Y, Y0, pg, M0, K0, cutoff, iter, beta, sigma, lam, pi, z, mu=None, celltype=None,
                   penl=1, est_z=2, max_lambda=True, est_lam=2, impt_it=5, sigma0=100, pi_alpha=1,
                   num_mc=3, lower=-np.inf, upper=np.inf, seed=0, compat=COMPAT_DEFAULT,
                   shards=None, trace=None, cell0=0, mode="sample", accel=None):
    """Same algorithm with the sufficient-statistic M-step; ``shards`` = list of
    cell boundaries emulating the multi-GPU partition + all-reduce; ``accel`` = ``oracle.accel.Accel()`` for the
    multi-threaded C / BLAS build of the O(n G) loops."""
    return _em_impute(Y, Y0, pg, M0, K0, cutoff, iter, beta, sigma, lam, pi, z, mu, celltype, penl,
                      est_z, max_lambda, est_lam, impt_it, sigma0, pi_alpha, num_mc, lower, upper,
                      seed, compat, literal=False, shards=shards, trace=trace, cell0=cell0, expect=(mode == "expect"),
                      accel=accel)


# ----------------------------------------------------------------------------
# do_impute
# ----------------------------------------------------------------------------
def do_impute_ref(dat, Y, beta, lam, sigma, mu, pi, pos_mean=None, pos_sd=None, celltype=None,
                  mcmc=10, burnin=2, pg=0.5, cutoff=0.1, seed=0, compat=COMPAT_DEFAULT,
                  consensus=True):
    """Restatement of do_impute (R/do_impute.R:29-160)."""
    pi_const = PI_REF if compat & COMPAT_PI else math.pi
    Y = np.array(Y, dtype=np.float64)
    G, n = Y.shape
    M0 = mu.shape[1]
    K0 = beta.shape[1]
    pos_mean = np.zeros(G) if pos_mean is None else np.asarray(pos_mean, dtype=np.float64)   # :38-39
    pos_sd = np.ones(G) if pos_sd is None else np.asarray(pos_sd, dtype=np.float64)          # :40-41
    celltype0 = np.zeros(n, dtype=np.int64) if celltype is None else np.asarray(celltype, dtype=np.int64) - 1
    MB = int(celltype0.max()) + 1                                            # :45
    pg = np.asarray(pg, dtype=np.float64)
    if pg.ndim == 0 or pg.size == 1:
        pg = np.full((G, MB), float(pg))                                     # :47-48
    Yc = (Y - pos_mean[:, None]) / pos_sd[:, None]                           # :50
    mask = np.asarray(dat) <= cutoff
    tot = np.zeros(mcmc + burnin)
    rec1 = np.zeros((G, n))
    rec2 = np.zeros((G, n))
    rEF = np.zeros((n, K0))
    rF2 = np.zeros((n, K0))
    rVF = np.zeros((n, K0))          # only the diagonal of record_varF is used (:156-157)
    cons = np.zeros((n, n)) if consensus else None
    for it in range(1, mcmc + burnin + 1):                                   # :67
        ll, lik, sconst, Ms, _, _ = epass(Yc, beta, sigma, lam, mu, pi, pi_const)
        tot[it - 1] = np.sum(np.log(lik.sum(axis=1)) + sconst)               # :89
        z = lik / lik.sum(axis=1)[:, None]                                   # :91
        if it > burnin and consensus:
            cons += z @ z.T                                                  # :96
        Wm = factor_means(Yc, beta, sigma, mu, Ms)                           # :102-104
        sample_sweep(Yc, mask, z, Wm, Ms, mu, beta, sigma, pg, celltype0, pos_mean, pos_sd,
                     cutoff, seed, it - 1, 0, draw_membership=True, clip=False)   # :107-136
        if it > burnin:                                                      # :139
            rec1 += Yc
            rec2 += Yc * Yc
            for m in range(M0):
                rEF += Wm[m] * z[:, m][:, None]
                rF2 += Wm[m] ** 2 * z[:, m][:, None]
                rVF += z[:, m][:, None] * np.diag(Ms[m])[None, :]
    varF = rF2 / mcmc - (rEF / mcmc) ** 2 + rVF / mcmc                       # :156-157
    # :158 mean(tot[-1:-burnin]); for burnin = 0 R's index -1:0 = c(-1, 0) still drops the first sweep (Q21)
    kept = tot[max(burnin, 1):]
    return dict(loglik=float(np.mean(kept)) if kept.size else float("nan"),
                impt=(rec1 / mcmc) * pos_sd[:, None] + pos_mean[:, None],
                impt_var=rec2 / mcmc - (rec1 / mcmc) ** 2, EF=rEF / mcmc, varF=varF,
                consensus_cluster=None if cons is None else cons / mcmc)


# ----------------------------------------------------------------------------
# Synthetic inputs (restates utils/utils.R:2-92 simulation_bulk, scaled)
# ----------------------------------------------------------------------------
def simulate(n=300, S0=20, K=6, MC=3, block_size=32, overlap=0, indepG=None, G=1000, dropr=0.3,
             seed=1, filter_genes=True):
    rng = np.random.default_rng(seed)
    Z = np.sort(rng.integers(0, MC, size=n))                                 # :8-10
    if indepG is None:
        indepG = G - (block_size + (K - 1) * (block_size - overlap))
    G = block_size + (K - 1) * (block_size - overlap) + indepG               # :14
    Sigma = rng.gamma(2.0, 0.3, size=G)                                      # :20
    Mu = np.repeat(np.exp(rng.normal(0.5, 0.5, size=G))[:, None], MC, axis=1)  # :23
    act = rng.permutation(G)[:S0 * MC]                                       # :25
    for m in range(MC):                                                      # :28-33
        ss = act[m * S0:(m + 1) * S0]
        Mu[ss, m] *= rng.choice(np.array([0.2, 0.5, 1.2, 1.5, 2.0]), size=S0)
    Gamma = np.zeros((G, K))
    for i in range(K):                                                       # :38-45
        a = (block_size - overlap) * i
        Gamma[a:a + block_size, i] = 1
    B = Gamma / 4                                                            # :47
    W = rng.standard_normal((K, n))                                          # :54
    Y = Mu[:, Z] + B @ W
    Y += rng.standard_normal((G, n)) * Sigma[:, None]                        # :55-56
    Y2 = Y.copy()
    pdrop = np.exp(-dropr * Mu.mean(axis=1) ** 2)                            # :66
    dZ = rng.random((G, n)) < pdrop[:, None]
    Y2[dZ] = 0
    Y2[Y2 < 0] = 0                                                           # :67-68
    if filter_genes:
        ind = np.nonzero((Y2 != 0).sum(axis=1) > 4)[0]                       # :70
        Y2, Y, B, Mu, Sigma = Y2[ind], Y[ind], B[ind], Mu[ind], Sigma[ind]
    return dict(Y2=np.asfortranarray(Y2), Y=Y, B=B, W=W, Mu=Mu, Sigma=Sigma, Z=Z + 1,
                bulk=Mu[:, Z].mean(axis=1))


def bench_state(sim, K0, M0, cutoff=0.1, pgv=0.6, seed=1):
    """EM-state initialisation for kernel benches/tests:
    beta ~ N(0, M0/K0), sigma=1, lambda=1/M0, pi=1/M0, z=one-hot(Z),
    pg=0.6, initial Y = Y2 with zeros replaced by truncated-normal draws."""
    rng = np.random.default_rng(seed + 1000)
    Y2 = sim["Y2"]
    G, n = Y2.shape
    beta = rng.standard_normal((G, K0)) * math.sqrt(M0 / K0)
    sigma = np.ones((G, M0))
    lam = np.full((M0, K0), 1.0 / M0)
    pi = np.full(M0, 1.0 / M0)
    clus = ((sim["Z"] - 1) % M0) + 1
    z = np.zeros((n, M0))
    z[np.arange(n), clus - 1] = 1.0
    pg = np.full((G, M0), pgv)
    mask = Y2 <= cutoff
    gm = np.array([Y2[g][~mask[g]].mean() if (~mask[g]).any() else 1.0 for g in range(G)])
    gs = np.array([Y2[g][~mask[g]].std() + 0.1 if (~mask[g]).any() else 1.0 for g in range(G)])
    gi, ci = np.nonzero(mask)
    r = (cutoff - gm[gi]) / gs[gi]
    Yi = np.array(Y2, order="F")
    Yi[gi, ci] = gm[gi] + gs[gi] * trunc_z(rng.random(gi.shape[0]) * 0.999 + 0.0005, r)
    return dict(Y=np.asfortranarray(Yi), Y0=Y2, pg=pg, beta=beta, sigma=sigma, lam=lam, pi=pi, z=z,
                celltype=clus.astype(np.int32))


# ----------------------------------------------------------------------------
# Driver-side consumers of the EM result
# ----------------------------------------------------------------------------
def yimp0_ref(res):
    """R/SIMPLE.R:421-426 (= R/SIMPLE_B.R:429-434)."""
    M0 = len(res["Ef"])
    n = res["z"].shape[0]
    G = res["beta"].shape[0]
    imp = np.zeros((n, G))
    for m in range(M0):
        imp += (res["mu"][:, m][:, None] + res["beta"] @ res["Ef"][m].T).T * res["z"][:, m][:, None]
    return imp.T * res["geneSd"][:, None] + res["geneM"][:, None]


def bic_ref(loglik_tot, n, G, K0, M0):
    """R/SIMPLE.R:429-430."""
    bic0 = -2 * loglik_tot + math.log(n) * (2 * G * M0 + G * (K0 + 1) + K0 * (M0 - 1))
    bic = (-2 * loglik_tot + math.log(n) * (2 * G * M0 + G) + math.log(n * G) * K0 * (M0 - 1)
           + K0 * math.log((G * n) / (G + n)) * (G + n))
    return bic, bic0


# ----------------------------------------------------------------------------
# N1: low-quality-gene regression and imputation draw between the two EM calls
# (R/SIMPLE.R:308-411, R/SIMPLE_B.R:335-419)
# ----------------------------------------------------------------------------
def lq_regress_ref(dat, resp, sigma, z, Ef, Varf, penl=1, resid_mode=0):
    """Literal restatement.  dat / resp / sigma are the rows of the lq genes.
    resp = regression response (dat, or init_imp when SIMPLE got one);
    resid_mode 0 = residual term of sg dropped (SIMPLE with init_imp = NULL: the
    inverted branch indexes NULL, quirk Q16, R/SIMPLE.R:354-359), 1 = residual
    from dat (SIMPLE with init_imp, R/SIMPLE.R:360-364; SIMPLE_B, R/SIMPLE_B.R:370-374)."""
    Gq, n = dat.shape
    M0 = z.shape[1]
    K0 = Ef[0].shape[1]
    nz = z.sum(axis=0)
    Vm = [Varf[m] * nz[m] for m in range(M0)]                               # R/SIMPLE.R:305-306
    sq = np.sqrt(z)
    mu = np.empty((Gq, M0))
    beta = np.empty((Gq, K0))
    sg_out = np.empty(Gq)
    for g in range(Gq):                                                      # :312
        V = sum(Vm[m] / sigma[g, m] for m in range(M0))                      # :314
        rows, ys = [], []
        for m in range(M0):                                                  # :318-330
            ys.append(resp[g] * sq[:, m] / math.sqrt(sigma[g, m]))
            Wb = Ef[m] * sq[:, m][:, None] / math.sqrt(sigma[g, m])
            Wmu = np.zeros((n, M0))
            Wmu[:, m] = sq[:, m] / math.sqrt(sigma[g, m])
            rows.append(np.concatenate([Wmu, Wb], axis=1))
        ML = np.concatenate([np.zeros((K0, M0)), np.linalg.cholesky(V).T], axis=1)   # :332
        W_aug = np.concatenate(rows + [ML])                                  # :335
        Y_aug = np.concatenate(ys + [np.zeros(K0)])                          # :337
        # glmnet(lambda = penalty K0/(M0+K0), penalty.factor = c(0.., 1..)) with the internal rescale of the
        # penalty factors to sum to M0 + K0  ==  kappa = penl / var(Y_aug) on the loadings, 0 on the intercepts
        kappa = np.concatenate([np.zeros(M0), np.full(K0, penl / np.var(Y_aug, ddof=1))])   # :341-349
        if LASSO_MODE == "glmnet":      # what glmnet's own stopping rule returns (test only)
            Nq = Y_aug.size
            lam_user = penl / Nq / np.var(Y_aug, ddof=1) * K0 / (M0 + K0)                    # :341, :345
            vpq = np.concatenate([np.zeros(M0), np.full(K0, (M0 + K0) / K0)])               # rescaled penalty.factor
            coef = lasso_glmnet_compat(W_aug.T @ W_aug, W_aug.T @ Y_aug, Nq * lam_user, Nq, Y_aug.sum(),
                                       float(Y_aug @ Y_aug), vp=vpq)
        else:
            coef = lasso_solve(W_aug.T @ W_aug, W_aug.T @ Y_aug, kappa)
        tempmu, coeff = coef[:M0], coef[M0:]                                 # :351-352
        sg = 0.0
        for m in range(M0):                                                  # :354-364
            if resid_mode:
                sg += np.sum((dat[g] - tempmu[m] - Ef[m] @ coeff) ** 2 * z[:, m])
            sg += coeff @ Vm[m] @ coeff
        mu[g], beta[g], sg_out[g] = tempmu, coeff, (sg + 1) / (n + 3)        # :365
    sigma_new = np.repeat(np.clip(sg_out, 1e-4, 9.0)[:, None], M0, axis=1)   # :371-373
    return mu, beta, sigma_new


def lq_regress_fast(dat, resp, sigma, z, Ef, Varf, penl=1, resid_mode=0):
    """Same result from cell-additive statistics; the M0 unpenalised intercepts are profiled out
    (their block of the Gram matrix is diagonal), leaving a K0-dimensional lasso per gene."""
    if LASSO_MODE == "glmnet":          # glmnet iterates the (M0 + K0)-dimensional problem: use the literal formulation
        return lq_regress_ref(dat, resp, sigma, z, Ef, Varf, penl=penl, resid_mode=resid_mode)
    Gq, n = dat.shape
    M0 = z.shape[1]
    K0 = Ef[0].shape[1]
    nz = z.sum(axis=0)
    S = np.stack([(Ef[m] * z[:, m][:, None]).T @ Ef[m] for m in range(M0)])
    a = np.stack([(Ef[m] * z[:, m][:, None]).sum(axis=0) for m in range(M0)])
    D = np.stack([S[m] + nz[m] * Varf[m] - np.outer(a[m], a[m]) / nz[m] for m in range(M0)])

    def stats(Yq):
        return (np.stack([Yq @ (Ef[m] * z[:, m][:, None]) for m in range(M0)]), Yq @ z, Yq @ np.sqrt(z), (Yq * Yq) @ z)

    T, U, Uh, U2 = stats(resp)
    N = M0 * n + K0
    sumY = np.sum(Uh / np.sqrt(sigma), axis=1)
    sumY2 = np.sum(U2 / sigma, axis=1)
    var = (sumY2 - sumY * sumY / N) / (N - 1)
    A = np.einsum("mjk,gm->gjk", D, 1.0 / sigma)
    rhs = np.einsum("mgk,gm->gk", T - U.T[:, :, None] * (a / nz[:, None])[:, None, :], 1.0 / sigma)
    beta = lasso_solve_batch(A, rhs, penl / var)
    mu = (U - beta @ a.T) / nz[None, :]
    C = np.stack([nz[m] * Varf[m] for m in range(M0)])
    sg = np.einsum("gj,mjk,gk->g", beta, C, beta)
    if resid_mode:
        Td, Ud, _, U2d = (T, U, Uh, U2) if resp is dat else stats(dat)
        for m in range(M0):
            bm = Td[m] - mu[:, m][:, None] * a[m][None, :]
            sg += (U2d[:, m] - 2 * mu[:, m] * Ud[:, m] + mu[:, m] ** 2 * nz[m] - 2 * np.sum(beta * bm, axis=1)
                   + np.einsum("gj,jk,gk->g", beta, S[m], beta))
    sigma_new = np.repeat(np.clip((sg + 1) / (n + 3), 1e-4, 9.0)[:, None], M0, axis=1)
    return mu, beta, sigma_new


def lq_impute_ref(dat, mu, beta, sigma, pg, z, Ef, celltype, cutoff=0.1, seed=0, sweep=0, cell0=0):
    """R/SIMPLE.R:376-411: one imputation draw of the lq genes' zero entries with the factor fixed at its posterior
    mean Ef[[m]][i,] (no factor noise), membership sampled from z; uncentred output."""
    Gq, n = dat.shape
    M0 = z.shape[1]
    K0 = beta.shape[1]
    Y = np.array(dat, dtype=np.float64)
    celltype0 = np.zeros(n, dtype=np.int64) if celltype is None else np.asarray(celltype, dtype=np.int64) - 1
    sample_sweep(Y, dat <= cutoff, z, Ef, [np.zeros((K0, K0))] * M0, mu, beta, sigma, np.asarray(pg), celltype0,
                 np.zeros(Gq), np.ones(Gq), cutoff, seed, sweep, cell0, draw_membership=(M0 > 1), clip=False)
    return Y


# ----------------------------------------------------------------------------
# N2: per-gene zero-inflated censored-normal fit and the initial imputation
# (R/fit_ztrunc.R:2-134, R/SIMPLE.R:15-73, R/SIMPLE_B.R:23-114)
# ----------------------------------------------------------------------------
_DBL_MAX = float(np.finfo(np.float64).max)
_BRENT_TOL = math.sqrt(np.finfo(np.float64).eps)        # optim(method="Brent") -> optimize(tol = con$reltol)


def brent_fmin(f, ax, bx, tol=_BRENT_TOL):
    """Brent's golden-section / parabolic-interpolation minimiser on [ax, bx] -- the published Netlib `fmin`
    (Forsythe, Malcolm & Moler 1977; Brent 1973) that stats::optimize documents as its algorithm and that
    optim(method = "Brent") calls (R/fit_ztrunc.R:80-83, 116-120).  Third-party arithmetic not in /root/reference:
    base R `stats` (no pinned version).  Non-finite function values are replaced by DBL_MAX as optimize does."""
    def fv_(x):
        v = f(x)
        return v if math.isfinite(v) else _DBL_MAX
    c = (3.0 - math.sqrt(5.0)) * 0.5
    eps = math.sqrt(np.finfo(np.float64).eps)              # ~ square root of the relative machine precision
    a, b = ax, bx
    v = a + c * (b - a)
    w = x = v
    d = e = 0.0
    fx = fv_(x)
    fv = fw = fx
    tol3 = tol / 3.0
    while True:
        xm = (a + b) * 0.5
        tol1 = eps * abs(x) + tol3
        t2 = tol1 * 2.0
        if abs(x - xm) <= t2 - (b - a) * 0.5:
            break
        p = q = r = 0.0
        if abs(e) > tol1:                       # fit a parabola
            r = (x - w) * (fx - fv)
            q = (x - v) * (fx - fw)
            p = (x - v) * q - (x - w) * r
            q = (q - r) * 2.0
            if q > 0.0:
                p = -p
            else:
                q = -q
            r = e
            e = d
        if abs(p) >= abs(q * 0.5 * r) or p <= q * (a - x) or p >= q * (b - x):
            e = (b - x) if x < xm else (a - x)  # golden-section step
            d = c * e
        else:                                   # parabolic step
            d = p / q
            u = x + d
            if u - a < t2 or b - u < t2:
                d = tol1 if x < xm else -tol1
        if abs(d) >= tol1:
            u = x + d
        elif d > 0.0:
            u = x + tol1
        else:
            u = x - tol1
        fu = fv_(u)
        if fu <= fx:
            if u < x:
                b = x
            else:
                a = x
            v, w, x = w, x, u
            fv, fw, fx = fw, fx, fu
        else:
            if u < x:
                a = u
            else:
                b = u
            if fu <= fw or w == x:
                v, fv = w, fw
                w, fw = u, fu
            elif fu <= fv or v == x or v == w:
                v, fv = u, fu
    return x


def _LL(sigma, xbar, Sx, delta):                                             # R/fit_ztrunc.R:2-10
    mu = xbar + (Sx - sigma * sigma) / (xbar - delta)
    return math.log(sigma) + (Sx + (xbar - mu) ** 2) / 2 / (sigma * sigma) + float(special.log_ndtr((mu - delta) / sigma))


def _LLwZ(sigma, xbar, Sx, r01, p, delta):                                   # R/fit_ztrunc.R:18-25
    mu = xbar + (Sx - sigma * sigma) / (xbar - delta)
    phi = float(special.ndtr((delta - mu) / sigma))
    with np.errstate(divide="ignore", invalid="ignore"):
        lg = float(np.log(np.float64(1 - p + p * phi)))
        return math.log(sigma) + (Sx + (xbar - mu) ** 2) / 2 / (sigma * sigma) - float(np.float64(r01) * lg)


def ztruncnorm_ref(dat, cutoff=0.1, s_upper=4.0, s_lower=0.1, p_min=0.6, p_max=1.0):
    """R/fit_ztrunc.R:55-134.  NA is NaN.  Returns (param (G, 2), pg (G,))."""
    dat = np.asarray(dat, dtype=np.float64)
    G, n = dat.shape
    pos = dat > cutoff
    n1 = pos.sum(axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        mu = np.where(pos, dat, 0.0).sum(axis=1) / n1                        # :59 rowMeans(na.rm = T)
        dev = np.where(pos, dat - mu[:, None], 0.0)
        var = (dev * dev).sum(axis=1) / (n1 - 1.0)                           # :60 rowVars (NA when n1 < 2)
        var = np.where(n1 >= 2, var, np.nan) * (n1 - 1.0) / n1               # :62
    p_max = np.broadcast_to(np.asarray(p_max, dtype=np.float64), (G,)).copy()   # :64-66
    param = np.full((G, 2), np.nan)
    for g in range(G):                                                       # :70-93
        if n1[g] < 4:
            continue
        s = brent_fmin(lambda sg: _LL(sg, mu[g], var[g], cutoff), s_lower, s_upper)
        param[g] = (mu[g] + (var[g] - s * s) / (mu[g] - cutoff), s)
    with np.errstate(divide="ignore", invalid="ignore"):
        n0 = n - n1                                                          # :96
        r01 = n0 / n1.astype(np.float64)
        p1 = n1 / float(n)
        p11 = special.ndtr((param[:, 0] - cutoff) / param[:, 1])             # :99
        pg = p1 / p11
        ind = np.nonzero((pg > p_max) | np.isnan(pg))[0]                     # :102
    pg[ind] = p_max[ind]
    ind2 = np.nonzero(pg < p_min)[0]                                         # :104
    pg[ind2] = p_min
    for g in np.concatenate([ind, ind2]):                                    # :106-131 (later rows win on duplicates)
        if np.isnan(var[g]):
            param[g] = (-1.0, 0.5)
            continue
        s = brent_fmin(lambda sg: _LLwZ(sg, mu[g], var[g], r01[g], pg[g], cutoff), s_lower, s_upper)
        param[g] = (mu[g] + (var[g] - s * s) / (mu[g] - cutoff), s)
    return param, pg


def _init_cluster_fit(Y2, sel, cutoff, p_min, p_max):
    res, pgc = ztruncnorm_ref(Y2[:, sel], cutoff=cutoff, p_min=p_min, p_max=p_max)
    mu, sd = res[:, 0].copy(), res[:, 1].copy()
    na = np.isnan(mu)
    if na.any():
        sub = Y2[np.ix_(na, sel)]
        mu[na] = sub.mean(axis=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            sd[na] = sub.std(axis=1, ddof=1) if sub.shape[1] > 1 else np.nan
    return mu, sd, pgc, na


def init_impute_ref(Y2, M0, clus, p_min=0.6, cutoff=0.1, seed=0, sweep=0, cell0=0):
    """R/SIMPLE.R:15-73 with the draws taken from the Philox tape (entry stream, gene = row of Y2).
    Returns (impute, pg, mu (G, M0), sd (G, M0))."""
    Y2 = np.asarray(Y2, dtype=np.float64)
    G, n = Y2.shape
    clus = np.asarray(clus)
    impute = Y2.copy()
    pg = np.zeros((G, M0))
    mus, sds_ = np.zeros((G, M0)), np.zeros((G, M0))
    for i in range(M0):
        sel = clus == i + 1
        mu, sd, pgc, na = _init_cluster_fit(Y2, sel, cutoff, p_min, 1.0)     # :22-26
        pg[:, i] = pgc
        pg[pg > 0.99] = 0.99                                                 # :28 (whole matrix, every pass)
        pg[na, i] = 1.0                                                      # :36
        mus[:, i], sds_[:, i] = mu, sd
        cells = np.nonzero(sel)[0]
        gi, cj = np.nonzero(Y2[:, cells] <= cutoff)                          # :43
        ci = cells[cj]
        ms, sds, p = mu[gi], sd[gi], pg[gi, i]
        with np.errstate(divide="ignore", invalid="ignore"):
            r = (cutoff - ms) / sds
            prob = special.ndtr(r)                                           # :49
            prob_drop = (1 - p) / (prob * p + (1 - p))                       # :50
            ua, ub = entry_uniforms(seed, sweep, ci.astype(np.uint64) + np.uint64(cell0), gi)
            x = np.where(ua < prob_drop, ms + sds * special.ndtri(ub), ms + sds * trunc_z(ub, r))   # :51-66
        impute[gi, ci] = x
    impute[np.isnan(impute)] = 0.0                                           # :69
    return impute, pg, mus, sds_


def init_impute_bulk_ref(Y2, clus, pg1, cutoff=0.1, seed=0, sweep=0, cell0=0):
    """R/SIMPLE_B.R:23-114 (the `bulk` argument only feeds the verbose plots).  Three-way draw per zero entry:
    dropout of an expressed gene / structural zero ~ N(-1, 0.5) / censored."""
    Y2 = np.asarray(Y2, dtype=np.float64)
    pg1 = np.asarray(pg1, dtype=np.float64)
    clus = np.asarray(clus)
    impute0 = Y2.copy()
    M0 = int(clus.max())                                                     # :28
    for i in range(M0):
        sel = clus == i + 1
        mu, sd, pgc, na = _init_cluster_fit(Y2, sel, cutoff, 0.01, pg1[:, i])   # :34
        pgc = pgc.copy()
        pgc[na] = pg1[na, i]                                                 # :45
        cells = np.nonzero(sel)[0]
        gi, cj = np.nonzero(Y2[:, cells] <= cutoff)                          # :68
        ci = cells[cj]
        m_, s_, p, q1 = mu[gi], sd[gi], pgc[gi], pg1[gi, i]
        with np.errstate(divide="ignore", invalid="ignore"):
            r = (cutoff - m_) / s_
            prob = special.ndtr(r)                                           # :71
            den = prob * p + (1 - p)
            prob_drop = (p / q1 - p) / den                                   # :75
            prob_zero = (1 - p / q1) / den                                   # :78
            ua, ub = entry_uniforms(seed, sweep, ci.astype(np.uint64) + np.uint64(cell0), gi)
            zq = special.ndtri(ub)
            x = np.where(ua < prob_drop, m_ + s_ * zq,                       # :85-86
                         np.where(ua < prob_drop + prob_zero, -1.0 + 0.5 * zq,   # :89-90
                                  m_ + s_ * trunc_z(ub, r)))                 # :92-93
        impute0[gi, ci] = x
    impute0[np.isnan(impute0)] = 0.0                                         # :108
    return impute0


# ----------------------------------------------------------------------------
# N3: the run-level drivers SIMPLE / SIMPLE_B / selectKM
# (R/SIMPLE.R:200-455, R/SIMPLE_B.R:229-455, R/select_KM.R:126-189).
# Each stage draws from its own Philox key seed + stage (same numbering as the product driver).
# ----------------------------------------------------------------------------
ST_PMIN, ST_INIT, ST_EM_HQ, ST_LQ, ST_EM_ALL, ST_MI, ST_CLUS, ST_BETA = range(8)


def _hq_genes_ref(dat, cutoff, p_min, min_gene, fix_num):
    n1 = (dat > cutoff).mean(axis=1)                                         # R/SIMPLE.R:246
    if fix_num:
        return np.argsort(-n1, kind="stable")[:min_gene]                     # :248
    hq = np.nonzero(n1 >= p_min)[0]                                          # :250
    if len(hq) < min_gene:
        hq = np.argsort(-n1, kind="stable")[:min_gene]                       # :253
    return hq


def _driver_tail(dat, Y, pg, hq, em_hq, beta, sigma, mu, celltype, M0, K0, iter, penl, est_z, max_lambda, est_lam,
                 sigma0, pi_alpha, num_mc, cutoff, seed, resp, resid_mode, mcmc, burnin, extra):
    G, n = dat.shape
    beta[hq], sigma[hq], mu[hq] = em_hq["beta"], em_hq["sigma"], em_hq["mu"]   # R/SIMPLE.R:298-300
    Y[hq] = em_hq["Y"]                                                       # :301
    lq = np.setdiff1d(np.arange(G), hq)                                      # :310
    m_, b_, s_ = lq_regress_fast(dat[lq], dat[lq] if resp is None else resp[lq], sigma[lq], em_hq["z"], em_hq["Ef"],
                                 em_hq["Varf"], penl, resid_mode)            # :312-366
    mu[lq], beta[lq], sigma[lq] = m_, b_, s_                                 # :368-371
    np.clip(sigma, 1e-4, 9.0, out=sigma)                                     # :372-373
    Y[lq] = lq_impute_ref(dat[lq], m_, b_, sigma[lq], pg[lq], em_hq["z"], em_hq["Ef"], celltype, cutoff,
                          seed + ST_LQ, 0)                                   # :376-411
    em = em_impute_fast(Y, dat, pg, M0, K0, cutoff, iter, beta, sigma, em_hq["lambda"], em_hq["pi"], em_hq["z"],
                        mu=None, celltype=celltype, penl=penl, est_z=est_z, max_lambda=max_lambda, est_lam=est_lam,
                        impt_it=1, sigma0=sigma0, pi_alpha=pi_alpha, num_mc=num_mc, seed=seed + ST_EM_ALL)   # :414-419
    bic, bic0 = bic_ref(em["loglik"][-1], n, G, K0, M0)                      # :427-430
    out = dict(loglik_tot=em["loglik"], priorB=em["priorB"], loglik=None, BIC=bic, BIC0=bic0, pi=em["pi"], mu=em["mu"],
               sigma=em["sigma"], beta=em["beta"], z=em["z"], Yimp0=yimp0_ref(em), pg=pg)
    out["lambda"] = em["lambda"]
    if mcmc > 0:                                                             # :431-446
        mi = do_impute_ref(dat, em["Y"], em["beta"], em["lambda"], em["sigma"], em["mu"], em["pi"], em["geneM"],
                           em["geneSd"], celltype, mcmc=mcmc, burnin=burnin, pg=pg, cutoff=cutoff, seed=seed + ST_MI)
        out.update(loglik=mi["loglik"], impt=mi["impt"], impt_var=mi["impt_var"], Ef=mi["EF"], Varf=mi["varF"],
                   consensus_cluster=mi["consensus_cluster"])
    else:                                                                    # :447-453
        out.update(impt=em["Y"], impt_var=None, Ef=em["Ef"], Varf=em["Varf"], consensus_cluster=None)
    out.update(extra)
    return out


def _driver_defaults(G, K0, M0, beta, lam, sigma, mu, seed):
    pi = np.full(M0, 1.0 / M0)                                               # R/SIMPLE.R:210
    lam = np.full((M0, K0), 1.0 / M0) if lam is None else np.asarray(lam, dtype=np.float64)
    mu = np.zeros((G, M0)) if mu is None else np.array(mu, dtype=np.float64)
    sigma = np.ones((G, M0)) if sigma is None else np.array(sigma, dtype=np.float64)
    if beta is None:
        beta = np.random.default_rng(seed + ST_BETA).normal(size=(G, K0)) / math.sqrt(K0) * math.sqrt(M0)   # :224-226
    return pi, lam, mu, sigma, np.array(beta, dtype=np.float64)


def simple_ref(dat, K0=10, M0=1, iter=10, est_lam=1, impt_it=5, penl=1, init_imp=None, sigma0=100, pi_alpha=1,
               beta=None, max_lambda=True, lam=None, sigma=None, mu=None, est_z=1, clus=None, p_min=0.4, cutoff=0.1,
               K=20, min_gene=2000, num_mc=3, fix_num=False, mcmc=50, burnin=2, seed=0):
    """R/SIMPLE.R:200-455 with `clus` given (the clustering start is row N4)."""
    dat = np.asarray(dat, dtype=np.float64)
    G, n = dat.shape
    Y = dat.copy()
    pi, lam, mu, sigma, beta = _driver_defaults(G, K0, M0, beta, lam, sigma, mu, seed)
    clus = np.asarray(clus)
    if p_min is None:                                                        # :229-242
        p_min = float(np.quantile(init_impute_ref(dat, M0, clus, 0.4, cutoff, seed + ST_PMIN)[1].ravel(), 0.25))
    pg = np.full((G, M0), float(p_min))                                      # :243
    hq = _hq_genes_ref(dat, cutoff, p_min, min_gene, fix_num)
    z = np.zeros((n, M0))
    z[np.arange(n), clus - 1] = 1.0                                          # :268-269
    imp, pg_hq = init_impute_ref(dat[hq], M0, clus, p_min, cutoff, seed + ST_INIT)[:2]   # :276
    start = imp if init_imp is None else np.asarray(init_imp)[hq]
    em_hq = em_impute_fast(start, dat[hq], pg_hq, M0, K0, cutoff, 20, beta[hq], sigma[hq], lam, pi, z, mu=None,
                           celltype=clus, penl=penl, est_z=est_z, max_lambda=max_lambda, est_lam=est_lam,
                           impt_it=impt_it, sigma0=sigma0, pi_alpha=pi_alpha, num_mc=num_mc, seed=seed + ST_EM_HQ)   # :281-296
    pg[hq] = pg_hq                                                           # :297
    return _driver_tail(dat, Y, pg, hq, em_hq, beta, sigma, mu, clus, M0, K0, iter, penl, est_z, max_lambda, est_lam,
                        sigma0, pi_alpha, num_mc, cutoff, seed, None if init_imp is None else np.asarray(init_imp),
                        0 if init_imp is None else 1, mcmc, burnin, dict(initclus=clus, p_min=p_min))


def simple_b_ref(dat, K0, bulk, celltype, M0=1, clus=None, b=1, iter=10, est_z=1, impt_it=3, max_lambda=False,
                 est_lam=1, penl=1, sigma0=100, pi_alpha=1, beta=None, lam=None, sigma=None, mu=None, p_min=0.8,
                 min_gene=300, cutoff=0.1, num_mc=3, fix_num=False, mcmc=50, burnin=2, seed=0):
    """R/SIMPLE_B.R:229-455 with `clus` given."""
    dat = np.asarray(dat, dtype=np.float64)
    G, n = dat.shape
    Y = dat.copy()
    bulk = np.asarray(bulk, dtype=np.float64).reshape(G, -1)
    celltype = np.asarray(celltype)
    pi, lam, mu, sigma, beta = _driver_defaults(G, K0, M0, beta, lam, sigma, mu, seed)
    hq = _hq_genes_ref(dat, cutoff, p_min, min_gene, fix_num)                # :258-266
    MB = int(celltype.max())
    pg = np.zeros((G, MB))
    for i in range(MB):                                                      # :270-282
        rm = dat[:, celltype == i + 1].mean(axis=1)
        if b is None:
            x, w = bulk[hq, i], bulk[hq, i] ** 2
            b1 = float(np.sum(w * x * rm[hq]) / np.sum(w * x * x))           # coef(lm(y ~ 0 + x, weights = x^2))
        else:
            b1 = float(b)
        pg[:, i] = (rm + p_min) / (1.0 + b1 * bulk[:, i])
    np.clip(pg, 0.1, 1.0, out=pg)                                            # :284-285
    res = init_impute_bulk_ref(dat[hq], celltype, pg[hq], cutoff, seed + ST_INIT)   # :289-290
    clus = np.asarray(clus)
    z = np.zeros((n, M0))
    z[np.arange(n), clus - 1] = 1.0
    em_hq = em_impute_fast(res, dat[hq], pg[hq], M0, K0, cutoff, 30, beta[hq], sigma[hq], lam, pi, z, mu=None,
                           celltype=celltype, penl=penl, est_z=est_z, max_lambda=max_lambda, est_lam=est_lam,
                           impt_it=impt_it, sigma0=sigma0, pi_alpha=pi_alpha, num_mc=num_mc, seed=seed + ST_EM_HQ)   # :318-321
    return _driver_tail(dat, Y, pg, hq, em_hq, beta, sigma, mu, celltype, M0, K0, iter, penl, est_z, max_lambda, est_lam,
                        sigma0, pi_alpha, num_mc, cutoff, seed, None, 1, mcmc, burnin, {})


def select_km_ref(dat, K0, M0, rel=0.01, **kw):
    """R/select_KM.R:126-189 (always through SIMPLE, Q18)."""
    n = np.asarray(dat).shape[1]
    Ms = sorted(set(np.atleast_1d(M0).astype(int).tolist()) | {1})
    Ks = sorted(np.atleast_1d(K0).astype(int).tolist())
    BICs = {m: {k: math.nan for k in Ks} for m in Ms}
    best, mBIC, mK, mM, init_imp = {}, math.inf, None, None, None
    clus_of = kw.pop("clus_of")
    for M1 in Ms:
        prev = math.inf
        for K1 in Ks:
            r = simple_ref(dat, K1, M1, init_imp=init_imp, clus=clus_of(M1), **kw)
            crit = r["BIC0"] if n < 1000 else r["BIC"]
            if mBIC > crit:
                mBIC, mK, mM = crit, K1, M1
                best[M1] = r
            BICs[M1][K1] = crit
            if (crit - prev) > -abs(crit) * rel:
                break
            prev = crit
        if M1 == 1:
            init_imp = best[1]["impt"]
    return dict(BICs=BICs, mK=mK, mM=mM, result=best[mM])


# ----------------------------------------------------------------------------
# N4: initial clustering (R/SIMPLE.R:258-266, R/SIMPLE_B.R:293-302)
# ----------------------------------------------------------------------------
def init_cluster_ref(X, K, M0, nstart=300, iter_max=80, seed=0):
    """t(scale(t(X))) -> rank-K SVD -> cellmat = V diag(d) -> k-means.  The SVD is exact (LAPACK) instead of irlba's
    Lanczos iteration; k-means is Lloyd's algorithm from `nstart` sets of M0 distinct random cells (drawn from the
    Philox cell stream: sweep = restart, cell = centre index, block = retry) instead of Hartigan-Wong with R's RNG, so
    only the partition is comparable with the reference.  Returns (clus 1-based, cellmat, d, wss)."""
    X = np.asarray(X, dtype=np.float64)
    G, n = X.shape
    sd = X.std(axis=1, ddof=1)
    Xs = (X - X.mean(axis=1, keepdims=True)) / np.where(sd > 0, sd, 1.0)[:, None]      # :259
    _, sv, Vt = np.linalg.svd(Xs, full_matrices=False)                                   # :261
    cm = Vt[:K].T * sv[:K]                                                               # :263
    best = (np.inf, None)
    for r in range(nstart):                                                              # :264 nstart
        picks = []
        for j in range(M0):
            t = 0
            while True:
                u0 = cell_uniforms(seed, r, np.array([j], dtype=np.uint64), 2 * t + 1)[0, 2 * t]
                c = min(int(u0 * n), n - 1)
                if c not in picks or t >= 64 or n <= j:
                    picks.append(c)
                    break
                t += 1
        cen = cm[picks].copy()
        for _ in range(iter_max):
            d2 = ((cm[:, None, :] - cen[None, :, :]) ** 2).sum(axis=2)
            lab = np.argmin(d2, axis=1)
            for j in range(M0):
                if np.any(lab == j):
                    cen[j] = cm[lab == j].mean(axis=0)
        d2 = ((cm[:, None, :] - cen[None, :, :]) ** 2).sum(axis=2)
        lab = np.argmin(d2, axis=1)
        wss = float(d2[np.arange(n), lab].sum())
        if wss < best[0]:
            best = (wss, lab + 1)
    return best[1], cm, sv[:K], best[0]


def adjusted_rand_index(a, b):
    """mclust::adjustedRandIndex, the reference's own agreement metric (utils/simulation.R:71-80)."""
    a = np.asarray(a)
    b = np.asarray(b)
    _, ai = np.unique(a, return_inverse=True)
    _, bi = np.unique(b, return_inverse=True)
    ct = np.zeros((ai.max() + 1, bi.max() + 1))
    np.add.at(ct, (ai, bi), 1)
    comb = lambda x: x * (x - 1) / 2.0
    sij, sa, sb_ = comb(ct).sum(), comb(ct.sum(axis=1)).sum(), comb(ct.sum(axis=0)).sum()
    exp = sa * sb_ / comb(len(a))
    mx = 0.5 * (sa + sb_)
    return float((sij - exp) / (mx - exp)) if mx != exp else 1.0
