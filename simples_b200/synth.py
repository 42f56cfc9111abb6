"""Synthetic zero-inflated log-count matrices for benchmarks and smoke runs.

Restates the reference's simulator ``simulation_bulk`` (utils/utils.R:2-92) at
the scaled shapes of SURVEY.md 8(d): disjoint 32-gene factor blocks with
loadings 1/4, cluster-specific mean shifts on S0 marker genes, gamma noise
scales, mean-dependent dropout ``exp(-0.3 mean^2)``, negatives set to 0.  Gene
-level quantities depend only on ``seed`` (identical on every rank); cell-level
quantities depend on ``(seed, shard)`` so each GPU can generate its own shard.
"""
from __future__ import annotations

import math

import numpy as np


def gene_params(G, K, MC, S0=20, block_size=32, seed=1):
    rng = np.random.default_rng([seed, 0xA11CE])
    Sigma = rng.gamma(2.0, 0.3, size=G)                                       # utils.R:20
    Mu = np.repeat(np.exp(rng.normal(0.5, 0.5, size=G))[:, None], MC, axis=1)  # :23
    act = rng.permutation(G)[:S0 * MC]                                        # :25
    for m in range(MC):                                                       # :28-33
        ss = act[m * S0:(m + 1) * S0]
        Mu[ss, m] *= rng.choice(np.array([0.2, 0.5, 1.2, 1.5, 2.0]), size=S0)
    B = np.zeros((G, K))
    for i in range(K):                                                        # :38-47 (overlap 0)
        B[i * block_size:(i + 1) * block_size, i] = 0.25
    return dict(Sigma=Sigma, Mu=Mu, B=B)


def simulate_shard(gp, n, shard=0, seed=1, dropr=0.3, dtype=np.float64):
    """Cells of one shard: returns Y2 (G x n, F-order, zeros = dropouts) and true labels Z (1-based)."""
    rng = np.random.default_rng([seed, 0xCE11, shard])
    G, MC = gp["Mu"].shape
    K = gp["B"].shape[1]
    Z = np.sort(rng.integers(0, MC, size=n))                                  # :8-10
    Yt = np.empty((n, G), dtype=dtype)                                        # cell-major == F-order G x n
    pdrop = np.exp(-dropr * gp["Mu"].mean(axis=1) ** 2)                       # :66
    step = max(1, (1 << 24) // G)
    for a in range(0, n, step):
        b = min(n, a + step)
        W = rng.standard_normal((b - a, K))
        blk = gp["Mu"][:, Z[a:b]].T + W @ gp["B"].T + rng.standard_normal((b - a, G)) * gp["Sigma"][None, :]   # :54-56
        blk[rng.random((b - a, G)) < pdrop[None, :]] = 0.0                    # :66-67
        np.maximum(blk, 0.0, out=blk)                                         # :68
        Yt[a:b] = blk
    return Yt.T, Z + 1


def init_state(gp, Y2, Z, K0, M0, cutoff=0.1, pgv=0.6, seed=1):
    """Kernel-bench EM state: beta ~ N(0, M0/K0), sigma = 1, lambda = pi = 1/M0,
    z = one-hot(Z), pg = 0.6, Y = Y2 with zeros replaced by truncated-normal draws."""
    from scipy import special
    G, n = Y2.shape
    rng = np.random.default_rng([seed, 0xB0B])          # parameters: identical on every rank
    beta = rng.standard_normal((G, K0)) * math.sqrt(M0 / K0)
    sigma = np.ones((G, M0))
    lam = np.full((M0, K0), 1.0 / M0)
    pi = np.full(M0, 1.0 / M0)
    clus = ((Z - 1) % M0) + 1
    z = np.zeros((n, M0))
    z[np.arange(n), clus - 1] = 1.0
    pg = np.full((G, M0), pgv)
    gm = gp["Mu"].mean(axis=1)
    gs = gp["Sigma"] + 0.1
    rngc = np.random.default_rng([seed, 0xD1CE, int(Z[0]) + 7 * n])
    Yt = np.array(Y2.T, order="C")                       # (n, G) copy
    step = max(1, (1 << 23) // G)
    r = (cutoff - gm) / gs
    pr = special.ndtr(r)
    for a in range(0, n, step):
        b = min(n, a + step)
        blk = Yt[a:b]
        msk = blk <= cutoff
        u = rngc.random(blk.shape) * 0.998 + 0.001
        zq = special.ndtri(u * pr[None, :])
        zq = np.where(np.isfinite(zq), zq, r[None, :] - 0.05)        # Phi(r) underflows for genes far above the cutoff
        draw = gm[None, :] + gs[None, :] * zq
        blk[msk] = draw[msk]
    return dict(Y=Yt.T, Y0=Y2, pg=pg, beta=beta, sigma=sigma, lam=lam, pi=pi, z=z, celltype=clus.astype(np.int32))
