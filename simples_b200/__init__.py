"""simples_b200 -- B200-native EM imputation hot path of JunLiuLab/SIMPLEs.

Public surface mirrors the reference's R functions for this path
(``EM_impute``: R/EM_impute.R:75, ``do_impute``: R/do_impute.R:29) above a
C ABI (include/simples_b200.h) implemented in hand-written sm_100a CUDA.
Importing this package never imports the test oracle and never falls back to
CPU arithmetic.
"""
from ._lib import (COMPAT_DEFAULT, COMPAT_MU_DIV_N, COMPAT_PI, COMPAT_PRIORB_PREV, COMPAT_STALE_DT, LIB_PATH,
                   SimplesError, comm_init_from_torch, init, load)
from .api import EM_impute, EMContext, bic, dev_empty, dev_matrix, do_impute, em_out_shapes, init_impute, init_impute_bulk, lq_fit
from .cluster import init_cluster
from .dist import CellShard, Replicas, partition
from .driver import SIMPLE, SIMPLE_B, selectKM

__all__ = ["SIMPLE", "SIMPLE_B", "selectKM", "EM_impute", "do_impute", "lq_fit", "init_impute", "init_impute_bulk", "init_cluster", "EMContext", "em_out_shapes", "bic", "dev_matrix", "dev_empty", "CellShard", "Replicas", "partition", "SimplesError", "load", "init", "comm_init_from_torch", "LIB_PATH",
           "COMPAT_DEFAULT", "COMPAT_STALE_DT", "COMPAT_PI", "COMPAT_MU_DIV_N", "COMPAT_PRIORB_PREV"]
