"""Synthetic GWAS inputs of the shapes BASELINE.json names.

Genotypes: ancestral allele frequencies p_k ~ U(0.05, 0.5); three
sub-populations drawn by the Balding-Nichols model (F_ST = 0.05) so the
kinship has structure; dosages x_ik ~ Binomial(2, p_k^{pop(i)}) stored int8
{0,1,2} with no missing calls (the reference imputes NA before the scan,
GAPIT.EMMAxP3D.R:20-24); about 0.1% monomorphic and 0.1% duplicated adjacent
columns are injected to exercise GAPIT.EMMAxP3D.R:499 and :512.

Phenotype after GAPIT.Phenotype.Simulation.R:39-46,92-123: additive QTN
effects ~ N(0,1) on randomly chosen markers plus N(0, sigma_e^2) residual
scaled to the requested heritability.

The genotype array is returned *marker-major* (shape (m, n), C-contiguous):
that is byte-for-byte the memory of R's column-major n x m matrix, so
``G.T`` is the n x m matrix GAPIT passes around.
Generation is chunked over markers with one PCG64 stream per chunk, so the
result does not depend on the thread count.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np

BASE_SEED = 20160301
CHUNK = 4096


def _chunk_geno(seed_seq, pop, n, mc, fst, p_mono, p_dup):
    rng = np.random.Generator(np.random.PCG64(seed_seq))
    p_anc = rng.uniform(0.05, 0.5, size=mc)
    a = p_anc * (1.0 - fst) / fst
    b = (1.0 - p_anc) * (1.0 - fst) / fst
    npop = int(pop.max()) + 1
    p_pop = rng.beta(a[None, :], b[None, :], size=(npop, mc))
    # two Bernoulli draws per call through byte thresholds (p quantised to 1/65536)
    thr = np.minimum((p_pop * 65536.0).astype(np.uint32), 65535).astype(np.uint16)
    T = thr[pop].T                                  # (mc, n)
    u = rng.integers(0, 65536, size=(2, mc, n), dtype=np.uint16)
    g = (u[0] < T).astype(np.int8)
    g += (u[1] < T)
    mono = np.nonzero(rng.random(mc) < p_mono)[0]
    if mono.size:
        g[mono, :] = rng.integers(0, 3, size=mono.size, dtype=np.int8)[:, None]
    dup = np.nonzero(rng.random(mc) < p_dup)[0]
    dup = dup[dup > 0]
    g[dup, :] = g[dup - 1, :]
    return g


def make_genotypes(n, m, seed=BASE_SEED, npop=3, fst=0.05, p_mono=0.001, p_dup=0.001, threads=8):
    """-> int8 array (m, n), marker-major."""
    ss = np.random.SeedSequence(seed)
    rng0 = np.random.Generator(np.random.PCG64(ss.spawn(1)[0]))
    pop = rng0.integers(0, npop, size=n)
    nchunks = (m + CHUNK - 1) // CHUNK
    seeds = np.random.SeedSequence(seed + 1).spawn(nchunks)
    out = np.empty((m, n), dtype=np.int8)

    def work(c):
        lo = c * CHUNK
        hi = min(m, lo + CHUNK)
        out[lo:hi] = _chunk_geno(seeds[c], pop, n, hi - lo, fst, p_mono, p_dup)

    if threads > 1 and nchunks > 1:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(work, range(nchunks)))
    else:
        for c in range(nchunks):
            work(c)
    return out


def make_phenotypes(G, ntraits=1, nqtn=10, h2=0.5, seed=BASE_SEED):
    """G: (m, n) int8 marker-major -> Y (n, ntraits) fp64."""
    m, n = G.shape
    rng = np.random.Generator(np.random.PCG64(seed + 7))
    Y = np.empty((n, ntraits), dtype=np.float64)
    for t in range(ntraits):
        qtn = rng.choice(m, size=min(nqtn, m), replace=False)
        eff = rng.standard_normal(qtn.shape[0])
        gval = eff @ G[qtn].astype(np.float64)
        vg = float(gval.var())
        if vg <= 0:
            vg = 1.0
        ve = vg * (1.0 - h2) / h2
        Y[:, t] = 10.0 + gval + rng.standard_normal(n) * np.sqrt(ve)
    return Y


def make_covariates(G, npc=0, ncols=4096):
    """Intercept (+ npc leading PCs of a marker subsample, cf. Debuge.R:36,48)."""
    m, n = G.shape
    X0 = np.ones((n, 1), dtype=np.float64)
    if npc > 0:
        sub = G[: min(m, ncols)].astype(np.float64).T
        sub = sub - sub.mean(axis=0, keepdims=True)
        u, s, _ = np.linalg.svd(sub, full_matrices=False)
        X0 = np.column_stack([X0, u[:, :npc] * s[:npc]])
    return X0


_IUPAC = {frozenset("AG"): "R", frozenset("CT"): "Y", frozenset("GT"): "K", frozenset("AC"): "M",
          frozenset("CG"): "S", frozenset("AT"): "W"}


def make_hapmap_text(G, bit=2, missing=0.02, seed=BASE_SEED, heading=True, crlf=False, odd_rows=True):
    """Dosages (m, n) -> the text of a HapMap table (bytes): 11 leading columns, tab separated, one row per marker.

    Alleles are drawn per marker; dosage 0/2 are the two homozygotes, 1 the heterozygote (``bit=2``: "AG";
    ``bit=1``: the IUPAC letter).  ``missing`` of the calls become one of the reference's missing codes
    (GAPIT.Numericalization.R:8-22).  ``odd_rows`` adds the table shapes the reference special-cases: both
    heterozygote spellings in one row (four levels -> all zero, :76), a monomorphic row and a row with no heterozygote."""
    rng = np.random.Generator(np.random.PCG64(seed + 13))
    m, n = G.shape
    miss2, miss1 = ["NN", "XX", "--", "++", "//"], ["N", "X", "-", "+", "/"]
    lines = []
    if heading:
        lines.append("\t".join(["rs#", "alleles", "chrom", "pos", "strand", "assembly#", "center", "protLSID", "assayLSID",
                                "panelLSID", "QCcode"] + [f"taxa{i:05d}" for i in range(n)]))
    for k in range(m):
        a, b = rng.choice(list("ACGT"), size=2, replace=False)
        if bit == 2:
            codes = np.array([a + a, a + b, b + b])
        else:
            codes = np.array([a, _IUPAC[frozenset((a, b))], b])
        row = codes[G[k]].astype(object)
        if odd_rows and bit == 2 and k % 17 == 5:
            het = np.nonzero(G[k] == 1)[0]
            row[het[::2]] = b + a                                   # second spelling of the heterozygote
        if odd_rows and k % 23 == 7:
            row[:] = codes[0]                                       # monomorphic
        if odd_rows and k % 29 == 11:
            row[G[k] == 1] = codes[2]                               # two levels only
        if missing > 0:
            na = np.nonzero(rng.random(n) < missing)[0]
            row[na] = rng.choice(miss2 if bit == 2 else miss1, size=na.size)
        lines.append("\t".join([f"rs{k}", f"{a}/{b}", str(1 + k % 10), str(1000 + 37 * k), "+", "NA", "NA", "NA", "NA",
                                "NA", "NA"] + list(row)))
    return (("\r\n" if crlf else "\n").join(lines) + ("\r\n" if crlf else "\n")).encode()


def hapmap_rows(text):
    """The same table as the list of string rows R's read.table(head=FALSE) would hold."""
    return [ln.rstrip("\r").split("\t") for ln in text.decode().split("\n") if ln.strip("\r")]
