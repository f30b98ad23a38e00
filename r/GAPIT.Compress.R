# Replacement wrapper for GAPIT.Compress (reference: GAPIT.Compress.R:1-113), SURVEY.md section 8f.3.
# GAPIT.Main.R:386 calls it unconditionally, and dist() + hclust() on the n x n kinship (:37-38) is O(n^3) scalar R:
# at n >= 20,000 it costs more wall time than the whole GPU scan saves.  At the two settings the MLM / GLM models use
# the clustering is known without computing it:
#   GN >= n  every taxon is its own group, numbered in row order (cutree numbers clusters by first appearance);
#            the group kinship t(b) KI b / CT (:59-61) is KI itself;
#   GN == 1  one group; KG is the 1 x 1 mean of KI.
# Any other GN is forwarded to the reference body, saved beforehand as GAPIT.Compress.reference:
#     GAPIT.Compress.reference <- GAPIT.Compress ; source("r/GAPIT.Compress.R")
`GAPIT.Compress` <-
function(KI,kinship.cluster = "average",kinship.group = "Mean",GN=nrow(KI),Timmer,Memory){
n=nrow(KI)
if(!(GN>=n || GN==1)) return(GAPIT.Compress.reference(KI=KI,kinship.cluster=kinship.cluster,kinship.group=kinship.group,GN=GN,Timmer=Timmer,Memory=Memory))
Timmer=GAPIT.Timmer(Timmer=Timmer,Infor="CP start")
Memory=GAPIT.Memory(Memory=Memory,Infor="cp start")
line.names <- KI[,1]                                                           # :12
KI <- as.matrix(KI[ ,-1])                                                      # :18
if(GN>=n){
  group.membership=1:n
  KG=matrix(as.numeric(KI),n,n)       # mean of one element; Max / Min / Median (:91-96) of one element likewise
}else{
  group.membership=rep(1,n)
  KG=matrix(switch(kinship.group, Mean=mean(KI), Max=max(KI), Min=min(KI), Median=median(KI)),1,1)
}
rownames(KG)=c(1:nrow(KG))                                                     # :62-63
colnames(KG)=c(1:ncol(KG))
Timmer=GAPIT.Timmer(Timmer=Timmer,Infor="CP calculation")
Memory=GAPIT.Memory(Memory=Memory,Infor="cp calculation")
GA <- data.frame(cbind(as.character(line.names),as.numeric(group.membership) ))   # :104
return(list(GA=GA, KG=KG,Timmer=Timmer,Memory=Memory))
}
