# Replacement body for GAPIT.HapMap (reference: GAPIT.HapMap.R:1-66) fed by the FILE instead of the data frame:
# the per-marker apply(..., GAPIT.Numericalization) loop (:34-38) runs on the GPU over the raw text of the table.
# Same return list (GT, GD, GI).  GAPIT.Genotype calls it as GAPIT.HapMap(G, ...) with G = read.table(file.G, head=FALSE)
# (GAPIT.Genotype.R:166); pass the file path instead (`file.G` is in scope there) or keep the data frame and the
# reference body -- this wrapper accepts both.
`GAPIT.HapMap` <-
function(G,SNP.effect="Add",SNP.impute="Middle",heading=TRUE, Create.indicator = FALSE, Major.allele.zero = FALSE){
if(!is.character(G) || length(G) != 1 || Create.indicator)
  return(GAPIT.HapMap.reference(G,SNP.effect=SNP.effect,SNP.impute=SNP.impute,heading=heading,
                                Create.indicator=Create.indicator,Major.allele.zero=Major.allele.zero))
print(paste("Converting HapMap format to numerical under model of ", SNP.impute,sep=""))
text <- readBin(G, "raw", file.size(G))
print("Perform numericalization")
effect <- switch(SNP.effect, Add=0L, Dom=1L, Left=2L, Right=3L)
impute <- switch(SNP.impute, Minor=0L, Middle=1L, Major=2L, -1L)
GD <- .Call("gapit_R_hapmap", text, heading, effect, impute, Major.allele.zero)
head11 <- read.table(G, head=FALSE, sep="\t", colClasses=c(rep("character",4), rep("NULL",7)), flush=TRUE,
                     comment.char="", quote="", fill=TRUE, nrows=ncol(GD)+heading)[,1:4]
if(heading){
  GT <- t(read.table(G, head=FALSE, sep="\t", nrows=1, comment.char="", quote="", colClasses="character")[1,-(1:11)])
  colnames(GT) <- "taxa"
  GI <- head11[-1,c(1,3,4)]
}else{
  GT <- NULL
  GI <- head11[,c(1,3,4)]
}
colnames(GI) <- c("SNP","Chromosome","Position")
print("Succesfuly finished converting HapMap")
return(list(GT=GT,GD=GD,GI=GI))
}
