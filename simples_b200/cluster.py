"""Initial clustering of the cells (R/SIMPLE.R:258-266, R/SIMPLE_B.R:293-302)."""
from __future__ import annotations


def init_cluster(X, K, M0, *, seed=0, comm=None, nstart=300, iter_max=80):
    raise NotImplementedError("initial clustering is not built yet: pass `clus`")
