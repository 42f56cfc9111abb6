# Export golden vectors from the REAL reference package, to pin the oracle and the CUDA path against it.
#
#   Rscript tools/export_golden.R <out_dir>          (run from a checkout of JunLiuLab/SIMPLEs: needs utils/utils.R)
#   SIMPLES_R_GOLDEN=<out_dir> python -m pytest tests/test_r_golden.py            (CPU: oracle vs R)
#   SIMPLES_R_GOLDEN=<out_dir> python -m pytest tests/test_r_golden.py -m gpu     (B200: CUDA vs R)
#
# Needs R with SIMPLEs and its dependencies (glmnet, Rsolnp, msm, mixtools, matrixStats, irlba, foreach).
# Nothing here can run in the build image (no R); the loader test is skipped when SIMPLES_R_GOLDEN is unset.
#
# What is exported, and why each piece is a pure function of its exported inputs:
#   A. EM_impute with impt_it = iter: the sampling gate `it >= impt_it + 1` (R/EM_impute.R:161) never opens, so
#      loglik / pi / mu / sigma / beta / lambda / z / Ef / Varf / Y depend on the inputs only (no RNG).
#   B. ztruncnorm (R/fit_ztrunc.R:57-134): no RNG at all.
#   C. SIMPLE(clus = <given>, impt_it = 20, mcmc = 0) with EM_impute wrapped by a recorder: the first
#      (high-quality-gene) EM call is deterministic given its recorded arguments (20 iterations, gate closed);
#      the lq-gene regression (R/SIMPLE.R:308-373) is deterministic given that call's result, and its output
#      beta[lq, ] / sigma[lq, ] is visible in the recorded arguments of the SECOND EM call (R/SIMPLE.R:414-419).
#   D. do_impute(mcmc = 1, burnin = 0): the E-pass of the only sweep runs before any draw, so EF, varF and
#      consensus_cluster are deterministic (R/do_impute.R:69-104, :143-147); impt is a random draw, exported
#      only as per-gene means over the zero entries for a distributional check.
# Layout: raw little-endian doubles `<name>.f64` (column-major, as R stores them), int32 `<name>.i32`,
# `dims.txt` with key=value lines, `sessionInfo.txt`.
suppressMessages({library(SIMPLEs); library(foreach)})
source("utils/utils.R")
out <- commandArgs(trailingOnly = TRUE)[1]
if (is.na(out)) stop("usage: Rscript tools/export_golden.R <out_dir>")
dir.create(out, showWarnings = FALSE, recursive = TRUE)
wr <- function(name, x) writeBin(as.double(x), file.path(out, paste0(name, ".f64")), size = 8, endian = "little")
wri <- function(name, x) writeBin(as.integer(x), file.path(out, paste0(name, ".i32")), size = 4, endian = "little")
dims <- c()
dim_set <- function(key, val) dims[[key]] <<- val

# ---------------------------------------------------------------- inputs (utils/utils.R:2-92, as utils/simulation.R:44)
set.seed(1)
sim <- simulation_bulk(n = 300, S0 = 20, K = 6, MC = 3, block_size = 32, indepG = 1000 - 32 * 6, verbose = F, overlap = 0)
Y2 <- sim$Y2; G <- nrow(Y2); n <- ncol(Y2); K0 <- 10; M0 <- 3; cutoff <- 0.01
clus <- sim$Z
dim_set("G", G); dim_set("n", n); dim_set("K0", K0); dim_set("M0", M0); dim_set("cutoff", cutoff)

# ---------------------------------------------------------------- A. deterministic EM_impute
res <- SIMPLEs:::init_impute(Y2, M0, clus, 0.5, cutoff = cutoff, verbose = F)
z <- matrix(0, n, M0); for (m in 1:M0) z[clus == m, m] <- 1
beta <- matrix(rnorm(G * K0), G, K0) / sqrt(K0) * sqrt(M0)
argsA <- list(Y = res[[1]], Y0 = Y2, pg = res[[2]], M0 = M0, K0 = K0, cutoff = cutoff, iter = 5, beta = beta,
              sigma = matrix(1, G, M0), lambda = matrix(1 / M0, M0, K0), pi = rep(1 / M0, M0), z = z,
              celltype = clus, est_z = 1, est_lam = 1, impt_it = 5)
fitA <- do.call(SIMPLEs:::EM_impute, argsA)
dim_set("A_iter", 5); dim_set("A_est_z", 1); dim_set("A_est_lam", 1); dim_set("A_impt_it", 5)
for (k in c("Y", "Y0", "pg", "beta", "sigma", "lambda", "pi", "z")) wr(paste0("A_in_", k), argsA[[k]])
wri("A_in_celltype", clus)
for (k in c("loglik", "pi", "mu", "sigma", "beta", "lambda", "z", "Y", "geneM", "priorB")) wr(paste0("A_out_", k), fitA[[k]])
wr("A_out_Ef", do.call(cbind, fitA$Ef))          # M0 matrices n x K0 back to back (column-major)
wr("A_out_Varf", do.call(cbind, fitA$Varf))      # M0 matrices K0 x K0 back to back
# every intermediate iteration as well (iter = 1..4 from the same start): per-iteration loglik is in A_out_loglik;
# parameters after 1 iteration isolate a single M-step (glmnet / solnp tolerances before they compound)
fit1 <- do.call(SIMPLEs:::EM_impute, modifyList(argsA, list(iter = 1, impt_it = 1)))
for (k in c("loglik", "pi", "mu", "sigma", "beta", "lambda", "z")) wr(paste0("A1_out_", k), fit1[[k]])

# ---------------------------------------------------------------- B. ztruncnorm
zt <- SIMPLEs:::ztruncnorm(Y2, cutoff = cutoff, p_min = 0.5)
wr("B_zt_param", zt[[1]]); wr("B_zt_pg", zt[[2]])
dim_set("B_p_min", 0.5)

# ---------------------------------------------------------------- C. SIMPLE() with the EM calls recorded
rec <- new.env(); rec$calls <- list()
em_orig <- SIMPLEs:::EM_impute
em_rec <- function(Y, Y0, pg, M0, K0, cutoff, iter, beta, sigma, lambda, pi, z, mu = NULL, celltype = NULL,
                   penl = 1, est_z = 2, max_lambda = T, est_lam = 2, impt_it = 5, sigma0 = 100, pi_alpha = 1,
                   verbose = F, num_mc = 3, lower = -Inf, upper = Inf) {
  a <- list(Y = Y, Y0 = Y0, pg = pg, M0 = M0, K0 = K0, cutoff = cutoff, iter = iter, beta = beta, sigma = sigma,
            lambda = lambda, pi = pi, z = z, mu = mu, celltype = celltype, penl = penl, est_z = est_z,
            max_lambda = max_lambda, est_lam = est_lam, impt_it = impt_it, sigma0 = sigma0, pi_alpha = pi_alpha,
            num_mc = num_mc, lower = lower, upper = upper)
  r <- do.call(em_orig, c(a, list(verbose = verbose)))
  rec$calls[[length(rec$calls) + 1]] <- list(args = a, out = r)
  r
}
assignInNamespace("EM_impute", em_rec, ns = "SIMPLEs")
min_gene <- 200; p_min <- 0.5
betaC <- matrix(rnorm(G * K0), G, K0) / sqrt(K0) * sqrt(M0)
fitC <- SIMPLE(Y2, K0 = K0, M0 = M0, iter = 3, est_lam = 1, impt_it = 20, penl = 1, beta = betaC, est_z = 1,
               clus = clus, p_min = p_min, cutoff = cutoff, min_gene = min_gene, num_mc = 3, mcmc = 0)
assignInNamespace("EM_impute", em_orig, ns = "SIMPLEs")
stopifnot(length(rec$calls) == 2)
c1 <- rec$calls[[1]]; c2 <- rec$calls[[2]]
n1 <- rowMeans(Y2 > cutoff)                                        # R/SIMPLE.R:246-255
hq_ind <- which(n1 >= p_min); if (length(hq_ind) < min_gene) hq_ind <- order(n1, decreasing = T)[1:min_gene]
lq_ind <- setdiff(1:G, hq_ind)
stopifnot(nrow(c1$args$Y) == length(hq_ind))
dim_set("C_Ghq", length(hq_ind)); dim_set("C_Glq", length(lq_ind)); dim_set("C_iter1", c1$args$iter)
dim_set("C_impt_it1", c1$args$impt_it); dim_set("C_min_gene", min_gene); dim_set("C_p_min", p_min)
wri("C_hq_ind", hq_ind); wri("C_lq_ind", lq_ind)
wr("C_beta0", betaC)
for (k in c("Y", "Y0", "pg", "beta", "sigma", "lambda", "pi", "z")) wr(paste0("C1_in_", k), c1$args[[k]])
for (k in c("loglik", "pi", "mu", "sigma", "beta", "lambda", "z", "Y")) wr(paste0("C1_out_", k), c1$out[[k]])
wr("C1_out_Ef", do.call(cbind, c1$out$Ef)); wr("C1_out_Varf", do.call(cbind, c1$out$Varf))
# the lq-gene regression result as handed to the second EM call (mu[lq, ] is not passed on: R/SIMPLE.R:417)
wr("C_lq_beta", c2$args$beta[lq_ind, , drop = F]); wr("C_lq_sigma", c2$args$sigma[lq_ind, , drop = F])
wr("C_dat", Y2)

# ---------------------------------------------------------------- D. do_impute, deterministic part
argsD <- list(dat = Y2, Y = fitA$Y, beta = fitA$beta, lambda = fitA$lambda, sigma = fitA$sigma, mu = fitA$mu,
              pi = fitA$pi, pos_mean = fitA$geneM, pos_sd = fitA$geneSd, celltype = clus, mcmc = 1, burnin = 0,
              pg = res[[2]], cutoff = cutoff)
set.seed(2)
fitD <- do.call(SIMPLEs:::do_impute, argsD)
wr("D_EF", fitD$EF); wr("D_varF", fitD$varF); wr("D_consensus", fitD$consensus_cluster)
# distributional summary of 20 recorded sweeps: per-gene mean of impt over the entries <= cutoff
set.seed(3)
fitD2 <- do.call(SIMPLEs:::do_impute, modifyList(argsD, list(mcmc = 20, burnin = 2)))
zmask <- Y2 <= cutoff
wr("D_impt_zero_mean", rowSums(fitD2$impt * zmask) / pmax(rowSums(zmask), 1))
wr("D_loglik20", fitD2$loglik)
dim_set("D_mcmc", 20); dim_set("D_burnin", 2)

writeLines(paste0(names(dims), "=", format(unlist(dims), digits = 17)), file.path(out, "dims.txt"))
writeLines(capture.output(print(sessionInfo())), file.path(out, "sessionInfo.txt"))
cat("golden vectors written to", out, "\n")
