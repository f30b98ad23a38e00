"""Exercises every genotype-ingestion kernel once at a realistic size, for `ncu --metrics gpu__time_duration.sum,
dram__bytes_read.sum,dram__bytes_write.sum` (tools/ncu_ingest_summary.py turns the csv into GB/s per kernel)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gapit_b200 as gb
from gapit_b200 import synth

n = int(os.environ.get("ING_N", 5000)); m = int(os.environ.get("ING_M", 100000))
ctx = gb.Context(0)
G = synth.make_genotypes(n, m, seed=5, threads=16)
g = gb.Genotypes(G, ctx=ctx, marker_major=True)                    # int8: colstats
K = gb.GAPIT_kinship_VanRaden(g, ctx=ctx, verbose=False)           # transpose, gram, mirror, rowsum, epilogue
g.close()
g = gb.Genotypes(gb.Packed2.from_dosages(G, marker_major=True), ctx=ctx)   # unpack2
g.close()
mf = min(m, 40000)
g = gb.Genotypes(G[:mf].astype(np.float64), ctx=ctx, marker_major=True)    # convert<double>
g.close()
g = gb.Genotypes(G[:mf].astype(np.int32), ctx=ctx, marker_major=True)      # convert<int>
g.close()
nh, mh = int(os.environ.get("ING_NH", 2000)), int(os.environ.get("ING_MH", 20000))
text = synth.make_hapmap_text(G[:mh, :nh], bit=2, missing=0.02)
hm = gb.GAPIT_HapMap(text, SNP_impute="Major", ctx=ctx, device=True)        # numericalize
print("hapmap", hm["GD"].shape, len(text), "bytes of text")
hm["geno"].close()
Kz = gb.GAPIT_kinship_ZHANG(G[:20000].T, ctx=ctx, verbose=False)            # centred gram, het counts, reduce, adjust
print("done", K.shape, Kz.shape)
