# GAPIT3-style sugar named by BASELINE.json: GAPIT(model="GLM"/"MLM", kinship.algorithm="VanRaden").
# This snapshot's GAPIT() has no `model=` argument (GAPIT.R:2-19): GLM is group.from=group.to=1 and
# MLM is group.from=group.to=<number of taxa> (GAPIT.Power.compare.R:80-108, GAPIT.Main.R:269-279 clips it).
GAPIT.reference <- GAPIT
GAPIT <- function(..., model=NULL){
  if(is.null(model)) return(GAPIT.reference(...))
  grp <- switch(model, GLM=1, MLM=1000000, stop("model must be \"GLM\" or \"MLM\""))
  GAPIT.reference(..., group.from=grp, group.to=grp)
}

# GPUs: the library opens every visible GPU of the box at its first call and shards markers across them
# (gapit_mctx_create / gapit_mgeno_pack).  gapit.gpus(k) re-opens it on the first k devices; returns the count in use.
`gapit.gpus` <- function(ngpus = getOption("gapit.gpus", 0L)) .Call("gapit_R_init", as.integer(ngpus))
