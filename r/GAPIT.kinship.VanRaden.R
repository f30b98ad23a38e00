# Replacement body for GAPIT.kinship.VanRaden (reference: GAPIT.kinship.VanRaden.R:1-37).
# Signature, return value (bare n x n double matrix) and the five progress prints are unchanged;
# the arithmetic is one .Call into libgapitcuda.so (int8 tcgen05 cross-product + fp64 epilogue).
# source() this file AFTER gapit_functions.txt and dyn.load("gapitcuda_R.so").
`GAPIT.kinship.VanRaden` <-
function(snps,hasInbred=TRUE) {
print("Calculating kinship with VanRaden method...")
print("substracting P...")
print("Getting X'X...")
K=.Call("gapit_R_vanraden", if(is.raw(snps)) snps else as.matrix(snps))   # raw = a PLINK .bed (r/GAPIT.Bed.R)
print("Adjusting...")
print("Calculating kinship with VanRaden method: done")
return(K)
}
