"""CPU oracle for the SIMPLEs EM imputation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product path
(``simples_b200``) never imports this package and fails loudly without its
CUDA library.

PARITY UNPINNED: the reference (JunLiuLab/SIMPLEs) ships no tests, golden
vectors or fixtures for this path, R is not installed in this image, and the
lasso / lambda / sampling arithmetic lives in un-vendored, un-pinned CRAN
packages (glmnet, Rsolnp, msm, mixtools).  The oracle therefore restates
``R/EM_impute.R`` and ``R/do_impute.R`` line by line and is pinned by
self-consistency only (literal augmented-design M-step vs sufficient
statistics, dense MVN density vs Woodbury, KKT residuals); see DESIGN.md.

``simples_oracle.py`` is the NumPy restatement; ``simples_oracle_c.c`` + ``accel.py`` are the multi-threaded C (OpenMP) /
BLAS build of its O(n G) loops -- the CPU baseline ``bench.py`` times -- held to the NumPy statements by
``tests/test_oracle_c.py``.
"""
from .simples_oracle import *  # noqa: F401,F403
