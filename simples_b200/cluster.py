"""Initial clustering of the cells (R/SIMPLE.R:258-266, R/SIMPLE_B.R:293-302) on B200."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .api import _Comm, _f, _ptr, check


def init_cluster(X, K, M0, *, seed=0, comm=None, nstart=300, iter_max=80, stream=None, details=False):
    """``kmeans(irlba(t(scale(t(X))), nv = K)$v %*% diag(d), M0, iter.max = 80, nstart = 300)$cluster``.

    X is ``G x n`` (high-quality genes x cells).  Returns the labels (1..M0); ``details=True`` returns a dict with
    clus, cellmat (n x K), d (K) and tot_withinss.  The reference's Lanczos start and Hartigan-Wong restarts are
    random, so results agree with it as partitions, not entry by entry."""
    lib = _lib.load()
    X = _f(X)
    G, n = X.shape
    K = int(min(K, G - 1))
    a = _lib.ClusArgs()
    a.G, a.n, a.K, a.M0 = G, n, K, int(M0)
    a.X = _ptr(X)
    a.nstart, a.iter_max = int(nstart), int(iter_max)
    a.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    cw = None
    if comm is not None:
        cw = _Comm(comm)
        a.n_total, a.cell_offset, a.allreduce = int(comm.n_total), int(comm.cell_offset), cw.cb
    a.stream = stream
    clus = np.zeros(n, dtype=np.int32)
    cellmat = np.empty((n, K), order="F") if details else None
    d = np.empty(K) if details else None
    wss = np.zeros(1)
    o = _lib.ClusOut()
    o.clus, o.cellmat, o.d, o.tot_withinss = _ptr(clus), _ptr(cellmat), _ptr(d), _ptr(wss)
    rc = lib.simples_init_cluster(C.byref(a), C.byref(o))
    if rc < 0 and cw is not None and cw.error is not None:
        raise cw.error
    check(rc)
    return dict(clus=clus, cellmat=cellmat, d=d, tot_withinss=float(wss[0])) if details else clus
