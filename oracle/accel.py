"""Multi-threaded CPU build of the oracle (TEST / BASELINE INFRASTRUCTURE, see oracle/__init__).

``em_impute_fast(..., accel=Accel())`` runs the same EM iteration as the NumPy oracle with its three O(n G) loops replaced:

* E-pass (R/EM_impute.R:126-157): instead of materialising ``R = Y - mu_m`` per cluster, the quadratic form is expanded,
  ``sum_g (Y - mu_m)^2 / sigma_m = S2 - 2 Y'(mu_m / sigma_m) + sum mu_m^2 / sigma_m`` and
  ``x2_m = Y'(B / sigma_m) - mu_m'(B / sigma_m)``: one threaded BLAS product ``Y' [B/sigma_s | mu_m/sigma_m]`` plus the C loop
  ``so_sq_over_sigma`` -- the formulation the CUDA path uses.  The factor means
  ``Wm[[m]] = ((Y - mu_m)/sigma_m)' B M_m`` (:155-157) come from the same product.
* MC sweep (R/EM_impute.R:163-199): cell-level draws in NumPy as in ``oracle.sample_sweep``, the per-entry part in C
  (``so_sweep_entries``, OpenMP over cells, same Philox tape).
* sufficient statistics: BLAS products in NumPy, the second moments ``sum_i Y_gi^2 w_i`` in C.

Everything small (cluster constants, lasso, lambda, mu, pi) is the NumPy oracle's own code.  tests/test_oracle_c.py holds
this build to the NumPy oracle (<= 1e-10 on every returned quantity, identical Bernoulli outcomes).
"""
from __future__ import annotations

import ctypes
import math
import os

import numpy as np
from scipy import special

from . import simples_oracle as SO

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libsimples_oracle_c.so")

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_longlong)
_bp = ctypes.POINTER(ctypes.c_ubyte)
_ll = ctypes.c_longlong


def _d(a):
    return a.ctypes.data_as(_dp)


def load():
    """The C library; built on demand when a compiler is present (the GPU box receives the prebuilt file)."""
    if not os.path.exists(_LIBPATH):
        from . import build_c
        build_c.build()
    lib = ctypes.CDLL(_LIBPATH)
    lib.so_abi_version.restype = ctypes.c_int
    lib.so_max_threads.restype = ctypes.c_int
    lib.so_sweep_entries.restype = _ll
    lib.so_sweep_entries.argtypes = [_dp, _bp, _ll, _ll, ctypes.c_int, ctypes.c_int, ctypes.c_int, _ip, _dp, _dp, _dp, _dp,
                                     _dp, _ip, _dp, _dp, ctypes.c_double, ctypes.c_ulonglong, ctypes.c_uint,
                                     ctypes.c_ulonglong, ctypes.c_double, ctypes.c_double, ctypes.c_int]
    lib.so_sq_over_sigma.restype = None
    lib.so_sq_over_sigma.argtypes = [_dp, _ll, _ll, _dp, ctypes.c_int, _dp]
    lib.so_rowsq_weighted.restype = None
    lib.so_rowsq_weighted.argtypes = [_dp, _ll, _ll, _dp, _dp, _dp]
    for nm in ("so_qnorm_v", "so_pnorm_v"):
        getattr(lib, nm).restype = None
        getattr(lib, nm).argtypes = [_dp, _ll, _dp]
    lib.so_trunc_z_v.restype = None
    lib.so_trunc_z_v.argtypes = [_dp, _dp, _ll, _dp]
    assert lib.so_abi_version() == 1
    return lib


def _fcol(a):
    """G x n matrix as column-major float64 without a copy when it already is."""
    return a if (a.dtype == np.float64 and a.flags.f_contiguous) else np.asfortranarray(a, dtype=np.float64)


class Accel:
    def __init__(self):
        self.lib = load()
        self.threads = int(self.lib.so_max_threads())
        self._wm = None

    # ---- E-pass + factor means ------------------------------------------------------------------------------------
    def epass(self, Yc, beta, sigma, lam, mu, pi, pi_const=SO.PI_REF, M_list=None, beta2_fix=None, dt_fix=None):
        """oracle.epass (R/EM_impute.R:126-141, MC-loop variant :202-213) by one projection; also prepares the factor
        means of :155-157 for ``factor_means``."""
        assert Yc.flags.f_contiguous and Yc.dtype == np.float64
        G, n = Yc.shape
        M0, K0 = mu.shape[1], beta.shape[1]
        Ms, b2s, dts = [], [], []
        for m in range(M0):
            if M_list is None:
                beta2, Mm, dt = SO.cluster_terms(beta, sigma, lam, m)
            else:
                Mm, beta2, dt = M_list[m], beta2_fix, dt_fix
            Ms.append(Mm)
            b2s.append(beta2)
            dts.append(dt)
        # distinct sigma columns (one after the first M-step: sigma is cluster-shared, R/EM_impute.R:269)
        cols, sidx = [], []
        for m in range(M0):
            for j, c in enumerate(cols):
                if np.array_equal(sigma[:, c], sigma[:, m]):
                    sidx.append(j)
                    break
            else:
                cols.append(m)
                sidx.append(len(cols) - 1)
        NS = len(cols)
        inv = np.ascontiguousarray((1.0 / sigma[:, cols]).T)                      # [NS][G]
        stale = M_list is not None and beta2_fix is not None
        ops = [beta * inv[s][:, None] for s in range(NS)]                        # B / sigma_s      (Wm, and x2 when fresh)
        if stale:
            ops.append(beta2_fix)                                                # stale beta2      (x2 of the MC loop, Q1)
        ops.append(mu * inv[sidx].T)                                             # mu_m / sigma_m
        Q = np.concatenate(ops, axis=1)
        P = Yc.T @ Q                                                             # n x (NS K0 [+ K0] + M0)
        S2 = np.empty((NS, n))
        self.lib.so_sq_over_sigma(_d(Yc), G, n, _d(inv), NS, _d(S2))
        ll = np.empty((n, M0))
        Wm = []
        off_mu = P.shape[1] - M0
        for m in range(M0):
            s = sidx[m]
            bs = ops[s]
            Pw = P[:, s * K0:(s + 1) * K0] - (mu[:, m] @ bs)[None, :]            # ((Y - mu_m)/sigma_m)' B
            if stale:
                x2 = P[:, NS * K0:(NS + 1) * K0] - (mu[:, m] @ beta2_fix)[None, :]
            else:
                x2 = Pw
            y = S2[s] - 2.0 * P[:, off_mu + m] + float(np.sum(mu[:, m] * mu[:, m] * inv[s]))      # :128
            y = y - np.sum((x2 @ Ms[m]) * x2, axis=1)                            # :130
            ll[:, m] = (-y - dts[m] - G * math.log(2 * pi_const)) / 2            # :135
            Wm.append(Pw @ Ms[m])                                                # :155-157 (M_m symmetric)
        ll = ll + np.log(pi)[None, :]                                            # :138
        sconst = ll.max(axis=1)                                                  # :139
        lik = np.exp(ll - sconst[:, None])                                       # :140-141
        self._wm = (id(M_list) if M_list is not None else id(Ms), Wm)
        return ll, lik, sconst, Ms, b2s, dts

    def factor_means(self, Yc, beta, sigma, mu, Ms):
        key, Wm = self._wm if self._wm is not None else (None, None)
        if key != id(Ms):           # not the cluster constants of the last epass: the NumPy formula
            return SO.factor_means(Yc, beta, sigma, mu, Ms)
        return Wm

    # ---- MC sweep -------------------------------------------------------------------------------------------------
    def sample_sweep(self, Yc, mask, z, Wm, Ms, mu, beta, sigma, pg, celltype0, gene_mean, gene_sd, cutoff, seed, sweep,
                     cell0=0, draw_membership=True, lower=-np.inf, upper=np.inf, clip=True, expect=False):
        if expect:
            return SO.sample_sweep(Yc, mask, z, Wm, Ms, mu, beta, sigma, pg, celltype0, gene_mean, gene_sd, cutoff, seed,
                                   sweep, cell0, draw_membership, lower, upper, clip, expect=True)
        assert Yc.flags.f_contiguous and Yc.dtype == np.float64
        G, n = Yc.shape
        M0, K0 = len(Ms), beta.shape[1]
        cells = np.arange(n, dtype=np.uint64) + np.uint64(cell0)
        U = SO.cell_uniforms(seed, sweep, cells, 1 + K0)
        if draw_membership:
            cum = np.cumsum(z, axis=1)
            im = np.minimum((U[:, :1] >= cum).sum(axis=1), M0 - 1)               # :164
        else:
            im = np.zeros(n, dtype=np.int64)                                     # :166
        eps = special.ndtri(U[:, 1:1 + K0])
        Mh = [SO.sym_sqrt(Mm) for Mm in Ms]
        f = np.empty((n, K0))
        for m in range(M0):
            sel = im == m
            if sel.any():
                f[sel] = Wm[m][sel] + eps[sel] @ Mh[m]                            # :171
        mk = mask if (mask.dtype == np.bool_ and mask.flags.f_contiguous) else np.asfortranarray(mask, dtype=np.bool_)
        im64 = np.ascontiguousarray(im, dtype=np.int64)
        ct64 = np.ascontiguousarray(celltype0, dtype=np.int64)
        c = np.ascontiguousarray
        k0, k1 = SO._seed_key(seed)
        self.lib.so_sweep_entries(_d(Yc), mk.view(np.uint8).ctypes.data_as(_bp), G, n, K0, M0, pg.shape[1],
                                  im64.ctypes.data_as(_ip), _d(c(f)), _d(c(mu)), _d(c(beta)), _d(c(sigma)), _d(c(pg)),
                                  ct64.ctypes.data_as(_ip), _d(c(gene_mean)), _d(c(gene_sd)), float(cutoff),
                                  (k1 << 32) | k0, int(sweep), int(cell0), float(lower), float(upper), int(bool(clip)))
        return im, f

    # ---- sufficient statistics ------------------------------------------------------------------------------------
    def _rowsq(self, Yc, w):
        G, n = Yc.shape
        out = np.empty(G)
        scratch = np.empty(self.threads * G)
        self.lib.so_rowsq_weighted(_d(Yc), G, n, _d(np.ascontiguousarray(w, dtype=np.float64)), _d(out), _d(scratch))
        return out

    def local_stats(self, Yc, z, Wm, shared_sigma):
        """oracle.local_stats with the second moments in C."""
        Yc = _fcol(Yc)
        M0 = z.shape[1]
        st = {}
        st["nz"] = z.sum(axis=0)
        st["c"] = np.sqrt(z).sum(axis=0)
        st["S"] = np.stack([(Wm[m] * z[:, m][:, None]).T @ Wm[m] for m in range(M0)])
        st["a"] = np.stack([(Wm[m] * z[:, m][:, None]).sum(axis=0) for m in range(M0)])
        st["U"] = Yc @ z
        st["Uh"] = Yc @ np.sqrt(z)
        if shared_sigma:
            Wbar = sum(Wm[m] * z[:, m][:, None] for m in range(M0))
            st["Tbar"] = Yc @ Wbar
            st["R2"] = self._rowsq(Yc, z.sum(axis=1))
        else:
            st["T"] = np.stack([Yc @ (Wm[m] * z[:, m][:, None]) for m in range(M0)])
            st["U2"] = np.stack([self._rowsq(Yc, z[:, m]) for m in range(M0)], axis=1)
        return st
