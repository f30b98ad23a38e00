# prcomp() stand-in for GAPIT.PCA (reference: GAPIT.PCA.R:10).  GAPIT.PCA itself is unchanged (plots, csv output,
# return list); only its `PCA.X <- prcomp(X)` line is served from the GPU:
#     PCA.X <- gapit.prcomp(X)          # instead of prcomp(X)
# $x and $sdev as prcomp returns them (component signs are LAPACK's to choose in both); $rotation (the m x k
# loadings written to GAPIT.PCA.loadings.csv when file.output=TRUE, :67) is formed on demand from the scores.
gapit.prcomp <- function(X, ncomp = min(nrow(X), ncol(X)), loadings = FALSE) {
  X <- as.matrix(X)
  r <- .Call("gapit_R_pca", X, as.integer(ncomp))
  sdev <- sqrt(r[[2]])
  out <- list(sdev = sdev, x = r[[1]], center = colMeans(X), scale = FALSE)
  colnames(out$x) <- paste("PC", seq_len(ncol(out$x)), sep = "")
  if (loadings) {   # rotation = X_c' u / s, u = x / s
    s <- sdev[seq_len(ncomp)] * sqrt(nrow(X) - 1)
    out$rotation <- crossprod(sweep(X, 2, out$center), sweep(out$x, 2, s * s, "/"))
  }
  class(out) <- "prcomp"
  out
}
