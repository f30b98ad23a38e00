# PLINK binary genotypes for the accelerated path (SURVEY.md section 8f.2).  The reference reads HapMap and numeric text
# only; this reader hands the .bed to the library as it is -- 2 bits per call over PCIe, PLINK codes mapped to dosages
# on the GPU -- so a 20,000 x 1,000,000 panel (5 GB on disk) never exists as a 160 GB R matrix.
#   bed <- GAPIT.Bed("panel")                       # panel.bed / panel.bim / panel.fam
#   K   <- GAPIT.kinship.VanRaden(bed$xs)           # the replacement body accepts the raw vector
#   p3d <- GAPIT.EMMAxP3D(ys=y, xs=bed$xs, K=K, X0=X0, ...)
# count = "A1" gives copies of the first .bim allele (PLINK's --recode A convention), "A2" the second.
`GAPIT.Bed` <-
function(prefix, count="A1"){
fam=read.table(paste(prefix,".fam",sep=""), head=FALSE, colClasses="character")
bim=read.table(paste(prefix,".bim",sep=""), head=FALSE, colClasses=c("character","character","numeric","numeric","character","character"))
path=paste(prefix,".bed",sep="")
xs=readBin(path, "raw", file.size(path))
if(length(xs)<3 || xs[1]!=as.raw(0x6c) || xs[2]!=as.raw(0x1b)) stop("not a PLINK .bed file")
if(xs[3]!=as.raw(0x01)) stop("PLINK .bed in sample-major mode: re-export variant-major")
attr(xs,"n")=nrow(fam)
attr(xs,"count")=ifelse(count=="A2",2L,1L)
GI=data.frame(SNP=bim[,2],Chromosome=bim[,1],Position=bim[,4])
return(list(xs=xs, GT=fam[,2], GI=GI, alleles=bim[,5:6], n=nrow(fam), m=nrow(bim)))
}
