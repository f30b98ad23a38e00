# Replacement body for GAPIT.ZmatrixCompress (reference: GAPIT.ZmatrixCompress.R:1-46), SURVEY.md section 8f.3.
# Same result; the only change is :24,:34: the reference multiplies Z by the dense 0/1 matrix DS = diag(n)[x,]
# (an n^3 dgemm at GN = n), here the columns of Z that share a group ID are summed directly (rowsum, O(n^2)).
`GAPIT.ZmatrixCompress` <-
function(Z,GAU){
effect.Z=as.matrix(Z[1,-1])                                                    # :13
effect.GAU=as.matrix(GAU[,1])
taxa=as.data.frame(Z[-1,1])
GAU0=GAU[effect.GAU%in%effect.Z,]                                              # :17
order.GAU=order(GAU0[,1])
GAU1 <- GAU0[order.GAU,]
id.1=GAU1[which(GAU1[,3]<2),4]                                                 # :21
n=max(as.numeric(as.vector(id.1)))
x=as.numeric(as.matrix(GAU1[,4]))
order.Z=order(effect.Z)                                                        # :27
Z=Z[-1,-1]
Z <- Z[,order.Z]
Z <- matrix(as.numeric(as.matrix(Z)), nrow = nrow(Z), ncol = ncol(Z))          # :33
ZG=matrix(0,nrow(Z),n)                                                         # Z %*% diag(n)[x,]  (:24,:34)
if(!anyDuplicated(x)){ ZG[,x]=Z }else{ s=rowsum(t(Z),x); ZG[,as.numeric(rownames(s))]=t(s) }
Z=data.frame(cbind(taxa,ZG))                                                   # :37
Z=Z[order(as.matrix(taxa)),]                                                   # :41
return(list(Z=Z))
}
