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

# ---- round 2: PLINK .bed rows, GAPIT numeric text, taxa selection ----
P2 = gb.Packed2.from_dosages(G, marker_major=True)
# a .bed is the same 2 bits per call with PLINK's code map: build one from the packed rows' dosages (A1 counting)
code = np.array([3, 2, 0, 1], dtype=np.uint8)          # dosage of A1: 0 -> 11b, 1 -> 10b, 2 -> 00b, missing -> 01b
d = G.astype(np.uint8)
nb = (n + 3) // 4
pad = np.zeros((m, nb * 4), dtype=np.uint8)
pad[:, :n] = code[d]
bed = np.concatenate([np.array([0x6C, 0x1B, 0x01], dtype=np.uint8),
                      (pad[:, 0::4] | (pad[:, 1::4] << 2) | (pad[:, 2::4] << 4) | (pad[:, 3::4] << 6)).reshape(-1)])
rows = gb.Packed2.from_bed(bed, n, count="A1")
g = gb.Genotypes(rows, ctx=ctx, impute="Middle")                             # unpack2 with the .bed recode
cs, _ = g.colsums()
assert np.array_equal(cs, G.astype(np.int64).sum(1)), ".bed dosages differ"
sel = g.select_taxa(np.random.default_rng(1).permutation(n)[: n * 3 // 4])   # select_taxa (GD[GTindex,])
sel.close()
g.close()
nt, mt = int(os.environ.get("ING_NT", 2000)), int(os.environ.get("ING_MT", 50000))
lines = ["taxa\t" + "\t".join(f"m{k}" for k in range(mt))]
digits = np.array(["0", "1", "2"])
for i in range(nt):
    lines.append(f"line_{i:05d}\t" + "\t".join(digits[G[:mt, i]]))
text = ("\n".join(lines) + "\n").encode()
gd, GT, names, _ = gb.GAPIT_numeric_text(text, SNP_impute="Middle", ctx=ctx)  # tokenise + parse + byte transpose
assert np.array_equal(gd, G[:mt, :nt].T), "numeric text dosages differ"
print("numeric text", gd.shape, len(text), "bytes of text")
