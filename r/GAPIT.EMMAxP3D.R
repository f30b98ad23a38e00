# Replacement wrapper for GAPIT.EMMAxP3D (reference: GAPIT.EMMAxP3D.R:1-1022).
# The signature is unchanged.  The call is accelerated when it is the plain P3D scan
# (Z NULL or square, SNP.P3D, fullGD, no indicators, !optOnly -- SURVEY.md section 8b);
# every other argument combination is forwarded to the reference's own body, which must
# have been saved as GAPIT.EMMAxP3D.reference before sourcing this file:
#     GAPIT.EMMAxP3D.reference <- GAPIT.EMMAxP3D ; source("r/GAPIT.EMMAxP3D.R")
`GAPIT.EMMAxP3D` <-
function(ys,xs,K=NULL,Z=NULL,X0=NULL,CVI=NULL,CV.Inheritance=NULL,GI=NULL,GP=NULL,
		file.path=NULL,file.from=NULL,file.to=NULL,file.total=1, genoFormat="Hapmap", file.fragment=NULL,byFile=FALSE,fullGD=TRUE,SNP.fraction=1,
    file.G=NULL,file.Ext.G=NULL,GTindex=NULL,file.GD=NULL, file.GM=NULL, file.Ext.GD=NULL,file.Ext.GM=NULL,
    SNP.P3D=TRUE,Timmer,Memory,optOnly=TRUE,SNP.effect="Add",SNP.impute="Middle", SNP.permutation=FALSE,
    ngrids=100,llim=-10,ulim=10,esp=1e-10,name.of.trait=NULL, Create.indicator = FALSE, Major.allele.zero = FALSE){
squareZ = is.null(Z) || (ncol(Z) == nrow(Z))
# complete phenotypes / covariates and a full-rank X0: the reference drops NA phenotypes per trait (vids, :113-114) and
# returns the REML = 0 list for singular X (GAPIT.emma.REMLE.R:16-19); those calls keep the reference body.  Several
# rows of ys (g traits, looped at :81 with the rotation repeated for each) ARE accelerated: one eigendecomposition and
# one rotation serve all of them (BASELINE configs[4]).
cleanXY = !anyNA(ys) && (is.null(X0) || (!anyNA(X0) && det(crossprod(X0,X0)) != 0))
accelerated = squareZ && SNP.P3D && fullGD && !Create.indicator && !optOnly && !is.null(xs) && cleanXY
if(!accelerated) return(GAPIT.EMMAxP3D.reference(ys=ys,xs=xs,K=K,Z=Z,X0=X0,CVI=CVI,CV.Inheritance=CV.Inheritance,GI=GI,GP=GP,
    file.path=file.path,file.from=file.from,file.to=file.to,file.total=file.total,genoFormat=genoFormat,file.fragment=file.fragment,
    byFile=byFile,fullGD=fullGD,SNP.fraction=SNP.fraction,file.G=file.G,file.Ext.G=file.Ext.G,GTindex=GTindex,file.GD=file.GD,
    file.GM=file.GM,file.Ext.GD=file.Ext.GD,file.Ext.GM=file.Ext.GM,SNP.P3D=SNP.P3D,Timmer=Timmer,Memory=Memory,optOnly=optOnly,
    SNP.effect=SNP.effect,SNP.impute=SNP.impute,SNP.permutation=SNP.permutation,ngrids=ngrids,llim=llim,ulim=ulim,esp=esp,
    name.of.trait=name.of.trait,Create.indicator=Create.indicator,Major.allele.zero=Major.allele.zero))

Timmer=GAPIT.Timmer(Timmer=Timmer,Infor="P3D Start")
Memory=GAPIT.Memory(Memory=Memory,Infor="P3D Start")
if(is.null(dim(ys)) || ncol(ys) == 1)  ys <- matrix(ys, 1, length(ys))        # :29
if(is.null(X0)) X0 <- matrix(1, ncol(ys), 1)                                  # :30
if(!is.null(K)) {if(length(K)<= 1) K = NULL}                                  # :34
g <- nrow(ys); n <- ncol(ys); q0 <- ncol(X0); q1 <- q0 + 1
Y <- t(ys)                                                                    # n x g
y <- as.numeric(ys[g,])                                                       # the scalars the loop leaves are the last trait's
xs.arg <- if(is.raw(xs)) xs else as.matrix(xs)        # a raw vector is the content of a PLINK .bed (r/GAPIT.Bed.R)
impute <- switch(SNP.impute, Minor=0L, Middle=1L, Major=2L, -1L)
glm <- is.null(K)
if(!glm && !identical(getOption("gapit.reml.route"), "eigR")){
  # :48, :224  eigen(K) is decomposed ON the GPUs (spectrum sliced over them) and stays there, sliced for the rotation,
  # behind an external pointer: eig.L$vectors (3.2 GB at n = 20,000) never comes back to R -- the REML projections,
  # BLUP/PEV and the scan all read the resident copy.  R keeps the eigenvalues.
  es <- .Call("gapit_R_eigsys", K)
  on.exit(.Call("gapit_R_eigsys_free", es$handle), add=TRUE)
  Timmer=GAPIT.Timmer(Timmer=Timmer,Infor="eig.L"); Memory=GAPIT.Memory(Memory=Memory,Infor="eig.L")
  # :58, :202  the REML likelihood is evaluated from eig.L alone (gapit_reml_eigL: the same function of delta as
  # emma.delta.REML.{LL,dLL}.wo.Z on eig.R, delta agrees to 1e-13), so the second eigendecomposition is skipped.
  Vt <- .Call("gapit_R_eigsys_project", es$handle, cbind(X0, Y))               # crossprod(eig.L$vectors, cbind(X0, Y))
  r <- .Call("gapit_R_reml_eigL_multi", es$values, Vt[,-(1:q0),drop=FALSE], Vt[,1:q0,drop=FALSE], as.integer(ngrids), llim, ulim, esp)
  Timmer=GAPIT.Timmer(Timmer=Timmer,Infor="eig.R"); Memory=GAPIT.Memory(Memory=Memory,Infor="eig.R")
  REMLs <- r[1,g]; REMLE_delta <- r[2,g]; ves <- r[3,g]; vgs <- r[4,g]
  Timmer=GAPIT.Timmer(Timmer=Timmer,Infor="REML"); Memory=GAPIT.Memory(Memory=Memory,Infor="REML")
  bp <- .Call("gapit_R_eigsys_blup_pev", es$handle, X0, y, REMLE_delta, vgs)                   # :779-839
  BLUP <- bp[[1]]; PEV <- bp[[2]]; BLUP_Plus_Mean <- bp[[3]][1] + BLUP                        # :818-819
  Timmer=GAPIT.Timmer(Timmer=Timmer,Infor="PEV"); Memory=GAPIT.Memory(Memory=Memory,Infor="PEV")
  res <- .Call("gapit_R_eigsys_scan", xs.arg, es$handle, X0, Y, as.numeric(r[2,]), as.numeric(r[4,]), impute)
} else if(!glm){
  # options(gapit.reml.route="eigR"): the reference's own route -- eig.L and eig.R as R objects (gapit_R_eigen_L,
  # gapit_R_eigen_R, gapit_R_reml), the eigenvectors uploaded again for BLUP/PEV and for the scan
  eig.L <- .Call("gapit_R_eigen_L", K)                                        # :48
  Timmer=GAPIT.Timmer(Timmer=Timmer,Infor="eig.L"); Memory=GAPIT.Memory(Memory=Memory,Infor="eig.L")
  eig.R <- try(.Call("gapit_R_eigen_R", K, X0), silent=TRUE)                  # :58
  if(inherits(eig.R, "try-error"))                                            # :70-74
    return(list(ps = NULL, REMLs = NA, stats = NULL, effect.est = NULL, dfs = NULL,maf=NULL,nobs = NULL,Timmer=Timmer,Memory=Memory,
        vgs = NA, ves = NA, BLUP = NULL, BLUP_Plus_Mean = NULL, PEV = NULL, BLUE=NULL))
  etas <- crossprod(eig.R$vectors, Y)                                         # GAPIT.emma.REMLE.R:25
  r <- sapply(1:g, function(j) .Call("gapit_R_reml", eig.R$values, as.numeric(etas[,j]), as.integer(ngrids), llim, ulim, esp))   # :202
  rm(eig.R)
  Timmer=GAPIT.Timmer(Timmer=Timmer,Infor="eig.R"); Memory=GAPIT.Memory(Memory=Memory,Infor="eig.R")
  REMLs <- r[1,g]; REMLE_delta <- r[2,g]; ves <- r[3,g]; vgs <- r[4,g]
  Timmer=GAPIT.Timmer(Timmer=Timmer,Infor="REML"); Memory=GAPIT.Memory(Memory=Memory,Infor="REML")
  bp <- .Call("gapit_R_blup_pev", eig.L$vectors, eig.L$values, X0, y, REMLE_delta, vgs)      # :779-839
  BLUP <- bp[[1]]; PEV <- bp[[2]]; BLUP_Plus_Mean <- bp[[3]][1] + BLUP                        # :818-819
  Timmer=GAPIT.Timmer(Timmer=Timmer,Infor="PEV"); Memory=GAPIT.Memory(Memory=Memory,Infor="PEV")
  res <- .Call("gapit_R_p3d_scan", xs.arg, eig.L$vectors, eig.L$values, X0, Y, as.numeric(r[2,]), as.numeric(r[4,]), FALSE, impute)
} else {
  res <- .Call("gapit_R_p3d_scan", xs.arg, NULL, NULL, X0, Y, 0, 0, TRUE, impute)
}
Timmer=GAPIT.Timmer(Timmer=Timmer,Infor="Screening SNPs"); Memory=GAPIT.Memory(Memory=Memory,Infor="Screening SNPs")
base <- res$base                                                              # (7 + q0) x g
m <- nrow(res$effect)
if(glm){ ves <- base[4,g]; REMLs <- base[5,g]; vgs <- 0; BLUP <- 0; BLUP_Plus_Mean <- NaN; PEV <- ves }   # :868-875
if(base[7,1] != 0) print("At least two of your covariates are linearly dependent. Please reconsider the covariates you are using for GWAS and GPS")   # :711
beta.cv <- base[8:(7+q0),g]
stats <- res$tvalue; normal <- !is.na(stats)
dfs <- ifelse(normal, n - q1, NA)                                             # :586
rsquare_base <- ifelse(normal, matrix(base[3,], m, g, byrow=TRUE), NA)
# :972  stderr = beta[ncol(CVI)+1]/t : lands on the marker effect only when ncol(CVI) == q0, past the end it is NA
stderr <- matrix(NA_real_, m, g)
if(!is.null(CVI) && !is.null(ncol(CVI)) && ncol(CVI) == q0) stderr <- res$effect/stats
Timmer=GAPIT.Timmer(Timmer=Timmer,Infor="GWAS done for this Trait"); Memory=GAPIT.Memory(Memory=Memory,Infor="GWAS done for this Trait")
# BLUE = XCV %*% beta.Inheritance (:841-861, :877-889)
BLUE <- NA
if(q0 == 1) { BLUE <- X0 %*% beta.cv } else if(!is.null(CVI)) {
  XCV <- as.matrix(cbind(1, data.frame(CVI[,-1]))); beta.Inheritance <- beta.cv
  if(!is.null(CV.Inheritance)){ XCV <- XCV[,1:(1+CV.Inheritance)]; beta.Inheritance <- beta.cv[1:(1+CV.Inheritance)] }
  BLUE <- try(XCV %*% beta.Inheritance, silent=TRUE); if(inherits(BLUE, "try-error")) BLUE <- NA
}
return(list(ps = res$pvalue, REMLs = -2*REMLs, stats = stats, effect.est = res$effect,
    rsquare_base = rsquare_base, rsquare = res$rsquare, dfs = dfs, df = dfs,
    tvalue = stats, stderr = stderr, maf=matrix(res$maf, m, g), nobs = matrix(n, m, g), Timmer=Timmer, Memory=Memory,
    vgs = vgs, ves = ves, BLUP = BLUP, BLUP_Plus_Mean = BLUP_Plus_Mean, PEV = PEV, BLUE=BLUE,
    logLM = res$loglm[m,g], effect.snp=res$effect, effect.cv=beta.cv))
}
