# Replacement bodies / call sites for the stages of SIMPLE() and SIMPLE_B() around EM_impute / do_impute.
# SHIPPED AS SOURCE (R is not available in the build image).  Each block cites the reference lines it replaces.

# ---- device selection (no counterpart in the reference) ---------------------------------------------------------------
# simples_init_b200(0)      one GPU (default when never called: the current device)
# simples_init_b200(0:7)    this R session drives 8 GPUs: EM_impute() / do_impute() shard the cells over them inside
#                           the library (one host thread per GPU, one NCCL all-reduce per EM iteration); SIMPLE(),
#                           SIMPLE_B() and selectKM() need no change.
simples_init_b200 <- function(devices = 0L) .Call(C_simples_init, as.integer(devices))

# ---- R/SIMPLE.R:15-73 ---------------------------------------------------------------------------------------------
init_impute <- function(Y2, M0, clus, p_min = 0.6, cutoff = 0.1, verbose = F) {
  storage.mode(Y2) <- "double"
  res <- .Call(C_simples_init_impute, Y2, as.integer(M0), as.integer(clus), NULL, as.double(p_min),
               as.double(cutoff), as.double(sample.int(.Machine$integer.max, 1)))
  list(res$impute, res$pg)
}

# ---- R/SIMPLE_B.R:23-114 (`bulk` only feeds the verbose plots) ------------------------------------------------------
init_impute_bulk <- function(Y2, clus, bulk, pg1, cutoff = 0.1, verbose = F) {
  storage.mode(Y2) <- "double"; storage.mode(pg1) <- "double"
  .Call(C_simples_init_impute, Y2, as.integer(max(clus)), as.integer(clus), pg1, 0.01, as.double(cutoff),
        as.double(sample.int(.Machine$integer.max, 1)))$impute
}

# ---- R/SIMPLE.R:258-266, R/SIMPLE_B.R:293-302: inside SIMPLE() / SIMPLE_B() ------------------------------------------
#   if (is.null(clus)) {
#     clus <- .Call(C_simples_init_cluster, dat[hq_ind, ], as.integer(K), as.integer(M0), 300L, 80L,
#                   as.double(sample.int(.Machine$integer.max, 1)))
#   }

# ---- R/SIMPLE.R:308-411, R/SIMPLE_B.R:335-419: inside SIMPLE() / SIMPLE_B() ------------------------------------------
#   lq_ind <- setdiff(1:G, hq_ind)
#   fit <- .Call(C_simples_lq_fit, dat[lq_ind, , drop = F],
#                if (is.null(init_imp)) NULL else init_imp[lq_ind, , drop = F],   # SIMPLE_B: always NULL
#                sigma[lq_ind, , drop = F], pg[lq_ind, , drop = F], z, impute_hq$Ef, impute_hq$Varf,
#                as.integer(clus),                                                # SIMPLE_B: as.integer(celltype)
#                as.double(penl), as.double(cutoff),
#                if (is.null(init_imp)) 0L else 1L,                               # SIMPLE_B: 1L (R/SIMPLE_B.R:370-374)
#                as.double(sample.int(.Machine$integer.max, 1)))
#   mu[lq_ind, ] <- fit$mu; beta[lq_ind, ] <- fit$beta; sigma[lq_ind, ] <- fit$sigma
#   sigma[sigma > 9] <- 9; sigma[sigma < 1e-04] <- 1e-04
#   Y[lq_ind, ] <- fit$Y
