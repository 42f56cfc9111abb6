# Drop-in replacement bodies for the two hot-path functions of JunLiuLab/SIMPLEs.
# SHIPPED AS SOURCE (R is not available in the build image).  Same names, argument order, defaults and
# return lists as R/EM_impute.R:75-78,337-338 and R/do_impute.R:29-30,158-159, so SIMPLE(), SIMPLE_B()
# and selectKM() (R/SIMPLE.R:283-296,414-419,433-438; R/SIMPLE_B.R:318-321,425-427,441-443;
# R/select_KM.R:152-155) call them unchanged.
# NAMESPACE gains:  useDynLib(SIMPLEs, .registration = TRUE)
#
# Multi-GPU: call simples_init_b200(devices = 0:7) once per R session (r/SIMPLE_b200.R); the library then shards the
# cells over the GPUs inside these two calls (host thread per GPU, NCCL all-reduce) -- the R code does not change.
#
# Differences a caller can observe (DESIGN.md section 8):
#  * random draws come from a Philox4x32-10 tape keyed by `seed` (default: one draw from R's RNG, so
#    set.seed() still makes runs reproducible) instead of R's Mersenne-Twister stream;
#  * `verbose` is ignored (the reference block uses undefined globals, R/EM_impute.R:327-329);
#  * do_impute() materialises the n x n consensus matrix only when consensus = TRUE.

EM_impute <- function(Y, Y0, pg, M0, K0, cutoff, iter, beta, sigma, lambda, pi, z,
                      mu = NULL, celltype = NULL, penl = 1, est_z = 2, max_lambda = T, est_lam = 2,
                      impt_it = 5, sigma0 = 100, pi_alpha = 1, verbose = F, num_mc = 3, lower = -Inf,
                      upper = Inf, seed = NULL) {
  if (is.null(seed)) seed <- floor(runif(1) * 2^52)
  Y <- as.matrix(Y); Y0 <- as.matrix(Y0)
  storage.mode(Y) <- "double"; storage.mode(Y0) <- "double"
  pg <- as.matrix(pg); storage.mode(pg) <- "double"
  beta <- as.matrix(beta); sigma <- as.matrix(sigma); lambda <- as.matrix(lambda)
  storage.mode(beta) <- "double"; storage.mode(sigma) <- "double"; storage.mode(lambda) <- "double"
  if (!is.null(z)) { z <- as.matrix(z); storage.mode(z) <- "double" }
  if (!is.null(mu)) { mu <- as.matrix(mu); storage.mode(mu) <- "double" }
  if (!is.null(celltype)) celltype <- as.integer(celltype)
  .Call("C_simples_em_impute", Y, Y0, pg, as.integer(M0), as.integer(K0), as.double(cutoff),
        as.integer(iter), beta, sigma, lambda, as.double(pi), z, mu, celltype, as.double(penl),
        as.integer(est_z), as.logical(max_lambda), as.integer(est_lam), as.integer(impt_it),
        as.double(sigma0), as.double(pi_alpha), as.integer(num_mc), as.double(lower), as.double(upper),
        as.double(seed), PACKAGE = "SIMPLEs")
}

do_impute <- function(dat, Y, beta, lambda, sigma, mu, pi, pos_mean = NULL, pos_sd = NULL,
                      celltype = NULL, mcmc = 10, burnin = 2, verbose = F, pg = 0.5, cutoff = 0.1,
                      seed = NULL, consensus = TRUE) {
  if (is.null(seed)) seed <- floor(runif(1) * 2^52)
  G <- nrow(Y)
  MB <- if (is.null(celltype)) 1L else max(celltype)                 # R/do_impute.R:43-45
  if (length(pg) == 1) pg <- matrix(pg, G, MB)                        # R/do_impute.R:47-48
  dat <- as.matrix(dat); Y <- as.matrix(Y); pg <- as.matrix(pg)
  storage.mode(dat) <- "double"; storage.mode(Y) <- "double"; storage.mode(pg) <- "double"
  dbl <- function(x) { x <- as.matrix(x); storage.mode(x) <- "double"; x }     # .Call takes REAL() of every matrix
  beta <- dbl(beta); lambda <- dbl(lambda); sigma <- dbl(sigma); mu <- dbl(mu)
  if (!is.null(pos_mean)) pos_mean <- as.double(pos_mean)
  if (!is.null(pos_sd)) pos_sd <- as.double(pos_sd)
  if (!is.null(celltype)) celltype <- as.integer(celltype)
  .Call("C_simples_do_impute", dat, Y, beta, lambda, sigma, mu,
        as.double(pi), pos_mean, pos_sd, celltype, as.integer(mcmc), as.integer(burnin), pg,
        as.double(cutoff), as.double(seed), as.logical(consensus), PACKAGE = "SIMPLEs")
}
