# Replacement body for GAPIT.kinship.ZHANG (reference: GAPIT.kinship.ZHANG.R:1-66).
# Same signature, return value and progress prints; the centred cross-product, the range scaling and the two
# adjustments run in libgapitcuda.so (gapit_zhang) on the exact integer cross-product.
`GAPIT.kinship.ZHANG` <-
function(snps,hasInbred=TRUE) {
print("Calculating ZHANG relationship defined by Zhiwu Zhang...")
print("substracting mean...")
print("Getting X'X...")
K=.Call("gapit_R_zhang", as.matrix(snps))
print("Adjusting...")
print("Calculating kinship with Zhang method: done")
return(K)
}
