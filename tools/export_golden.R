# Run with the REAL reference package (R + SIMPLEs + glmnet/Rsolnp/msm/mixtools installed) to export
# golden vectors for the deterministic part of the hot path, in the layout tests/golden expects.
# EM_impute(..., impt_it = iter) never samples (gate at R/EM_impute.R:161), so it is a pure function of
# its inputs and pins beta / sigma / lambda / mu / pi / z / loglik of the CUDA path and the oracle against
# the reference (expect ~1e-4 relative on beta: glmnet thresh = 1e-7; ~1e-6 on lambda: solnp tol).
# Usage: Rscript tools/export_golden.R out_dir
suppressMessages({library(SIMPLEs); library(foreach)})
source("utils/utils.R")
out <- commandArgs(trailingOnly = TRUE)[1]; dir.create(out, showWarnings = FALSE)
set.seed(1)
sim <- simulation_bulk(n = 300, S0 = 20, K = 6, MC = 3, block_size = 32, indepG = 1000 - 32 * 6, verbose = F, overlap = 0)
Y2 <- sim$Y2; G <- nrow(Y2); n <- ncol(Y2); K0 <- 10; M0 <- 3
clus <- sim$Z
res <- SIMPLEs:::init_impute(Y2, M0, clus, 0.5, cutoff = 0.01, verbose = F)
z <- matrix(0, n, M0); for (m in 1:M0) z[clus == m, m] <- 1
beta <- matrix(rnorm(G * K0), G, K0) / sqrt(K0) * sqrt(M0)
args <- list(Y = res[[1]], Y0 = Y2, pg = res[[2]], M0 = M0, K0 = K0, cutoff = 0.01, iter = 5, beta = beta,
             sigma = matrix(1, G, M0), lambda = matrix(1 / M0, M0, K0), pi = rep(1 / M0, M0), z = z,
             celltype = clus, est_z = 1, est_lam = 1, impt_it = 5)
fit <- do.call(SIMPLEs:::EM_impute, args)
wr <- function(name, x) writeBin(as.double(x), file.path(out, paste0(name, ".f64")), size = 8, endian = "little")
for (k in c("Y", "Y0", "pg", "beta", "sigma", "lambda", "pi", "z")) wr(paste0("in_", k), args[[k]])
writeBin(as.integer(clus), file.path(out, "in_celltype.i32"), size = 4, endian = "little")
for (k in c("loglik", "pi", "mu", "sigma", "beta", "lambda", "z", "Y", "geneM")) wr(paste0("out_", k), fit[[k]])
# ztruncnorm has no random draws at all: its (mean, sd) and pg pin the oracle's restatement of
# R/fit_ztrunc.R:55-134 (Brent search: expect ~1e-7 relative on sd, the optimiser's own tolerance) and k_ztrunc.
zt <- SIMPLEs:::ztruncnorm(Y2, cutoff = 0.01, p_min = 0.5)
wr("zt_param", zt[[1]]); wr("zt_pg", zt[[2]])
# the lq-gene regression of R/SIMPLE.R:312-366 is deterministic given (z, Ef, Varf): export its inputs and outputs
# by running the block of SIMPLE() on the first 200 genes as "hq" (copy of the loop body, not reproduced here);
# the all-gene EM_impute call with impt_it = iter then pins the whole deterministic chain.
cat(sprintf("G=%d n=%d K0=%d M0=%d iter=5 cutoff=0.01\n", G, n, K0, M0), file = file.path(out, "dims.txt"))
print(sessionInfo())
