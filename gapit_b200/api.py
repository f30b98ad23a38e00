"""Host-side mirror of the reference's interface for the hot path.

R is not available in this image, so the reference-facing layer is restated in
Python with the reference's own names, argument meaning and return lists
(INTEGRATION.md shows the R shim that binds the same C ABI):

    GAPIT.kinship.VanRaden(snps, hasInbred)   -> GAPIT_kinship_VanRaden
    emma.eigen.L.wo.Z(K) / emma.eigen.R.wo.Z  -> emma_eigen_L_wo_Z / emma_eigen_R_wo_Z
    GAPIT.emma.REMLE(y, X, K, ...)            -> GAPIT_emma_REMLE
    GAPIT.EMMAxP3D(ys, xs, K, Z, X0, ...)     -> GAPIT_EMMAxP3D

All arithmetic happens in libgapitcuda.so; this module only marshals buffers.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

from . import _lib
from ._lib import GapitCudaError, check  # noqa: F401


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Context:
    """One GPU (gapit_ctx).  One per process under torchrun, one per R session."""

    def __init__(self, device=0, slices=None):
        self.lib = _lib.load()
        h = C.c_void_p()
        check(self.lib.gapit_ctx_create(int(device), C.byref(h)))
        self.h = h
        self.device = int(device)
        if slices:
            check(self.lib.gapit_set_slices(self.h, int(slices)))

    def set_stream(self, cuda_stream):
        """Run on a caller-owned stream (e.g. ``torch.cuda.current_stream().cuda_stream``); None restores the
        context's own stream.  Handle 0 is the legacy default stream (what torch uses until a stream is made
        current): it is passed as cudaStreamLegacy, because NULL means "restore" in the C ABI."""
        if cuda_stream is None:
            check(self.lib.gapit_ctx_set_stream(self.h, None))
        else:
            check(self.lib.gapit_ctx_set_stream(self.h, C.c_void_p(int(cuda_stream) or 1)))

    def set_slices(self, slices):
        """Digit planes of the rotation (4..8); 0 restores the library default."""
        check(self.lib.gapit_set_slices(self.h, int(slices)))

    @property
    def slices(self):
        return int(self.lib.gapit_get_slices(self.h))

    def stats(self):
        s = _lib.Stats()
        check(self.lib.gapit_ctx_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_}

    def close(self):
        if getattr(self, "h", None):
            self.lib.gapit_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


class Packed2:
    """Host genotypes at 2 bits each (GAPIT_B2): ``data`` is (m, ceil(n/4)) uint8, marker-major, taxon i of a marker in
    byte i//4 at bits 2(i%4).. (PLINK .bed bit order); code 0/1/2 = dosage, 3 = NA.  A quarter of the int8 bytes
    over PCIe; the device expands it to the int8 operand layout."""

    def __init__(self, data, n, dtype=_lib.GAPIT_B2):
        self.data = data if (isinstance(data, np.ndarray) and data.dtype == np.uint8 and data.flags.c_contiguous) \
            else np.ascontiguousarray(data, dtype=np.uint8)
        self.n = int(n)
        assert self.data.ndim == 2 and self.data.shape[1] >= (self.n + 3) // 4
        self.m = int(self.data.shape[0])
        self.dtype = int(dtype)      # GAPIT_B2, or GAPIT_BED_A1 / GAPIT_BED_A2: the rows of a PLINK .bed as they are

    @staticmethod
    def from_bed(bed, n, count="A1"):
        """The variant rows of a PLINK .bed (bytes, bytearray, numpy uint8 array or np.memmap of the whole file) for n
        samples: header checked by the library, rows used in place (no copy, no host-side recoding); dosage = copies
        of ``count`` ("A1", PLINK's --recode A convention, or "A2")."""
        buf = np.frombuffer(bed, dtype=np.uint8) if not isinstance(bed, np.ndarray) else bed.reshape(-1)
        m, off = C.c_int64(), C.c_int64()
        check(_lib.load().gapit_bed_header(_ptr(buf), buf.shape[0], int(n), C.byref(m), C.byref(off)))
        nb = (int(n) + 3) // 4
        rows = buf[off.value:off.value + m.value * nb].reshape(m.value, nb)
        return Packed2(rows, n, dtype=_lib.GAPIT_BED_A1 if count == "A1" else _lib.GAPIT_BED_A2)

    @staticmethod
    def from_dosages(a, marker_major=False, pinned=False, threads=0):
        """(n, m) dosages (or (m, n) with ``marker_major``), values 0/1/2, NaN / NA_integer_ / negative int8 = NA.
        Packed by the library on the host's threads (gapit_pack2_host)."""
        a = np.asarray(a)
        if not marker_major:
            a = a.T
        if a.dtype == np.int8 or a.dtype == np.uint8:
            dtype = _lib.GAPIT_I8
        elif a.dtype == np.int32:
            dtype = _lib.GAPIT_I32
        elif a.dtype.kind in "iu":
            a = np.where(a < 0, np.iinfo(np.int32).min, a).astype(np.int32)
            dtype = _lib.GAPIT_I32
        else:
            a = a.astype(np.float64, copy=False)
            dtype = _lib.GAPIT_F64
        a = np.ascontiguousarray(a)
        m, n = a.shape
        nb = (n + 3) // 4
        if pinned:
            import torch
            out = torch.empty((m, nb), dtype=torch.uint8).pin_memory().numpy()
        else:
            out = np.empty((m, nb), dtype=np.uint8)
        check(_lib.load().gapit_pack2_host(_ptr(a), dtype, n, m, n, _ptr(out), int(threads)))
        return Packed2(out, n)


class Genotypes:
    """Device-resident dosage matrix (gapit_geno): int8, marker-major, plus column statistics."""

    def __init__(self, snps, ctx=None, impute=None, marker_major=False):
        """snps: n x m array (taxa x markers, the reference's orientation), or with
        ``marker_major=True`` an (m, n) C-contiguous array (= R's column-major memory)."""
        self.ctx = ctx or default_context()
        lib = self.ctx.lib
        if isinstance(snps, Packed2):
            h = C.c_void_p()
            check(lib.gapit_geno_pack(self.ctx.h, _ptr(snps.data), snps.dtype, snps.n, snps.m, snps.data.shape[1],
                                      _lib.IMPUTE[impute], C.byref(h)))
            self.h = h
            self.n, self.m = snps.n, snps.m
            return
        a = np.asarray(snps)
        if not marker_major:
            # n x m: column-major memory is what the C ABI wants (element (i,k) at i + k*ld)
            a = a.T  # (m, n) view
        if a.dtype == np.int8 or a.dtype == np.uint8:
            dtype = _lib.GAPIT_I8
        elif a.dtype == np.int32:
            dtype = _lib.GAPIT_I32
        else:
            a = a.astype(np.float64, copy=False)
            dtype = _lib.GAPIT_F64
        a = np.ascontiguousarray(a)  # (m, n) row-major == n x m column-major, ld = n
        m, n = a.shape
        h = C.c_void_p()
        check(lib.gapit_geno_pack(self.ctx.h, _ptr(a), dtype, n, m, n, _lib.IMPUTE[impute], C.byref(h)))
        self.h = h
        self.n, self.m = int(n), int(m)

    @classmethod
    def from_device(cls, data_ptr, n, m, ld=None, ctx=None):
        """int8 dosages already on the context's device (marker-major, marker k at data_ptr + k*ld), e.g.
        ``t.data_ptr()`` of a contiguous (m, n) torch.int8 CUDA tensor."""
        self = cls.__new__(cls)
        self.ctx = ctx or default_context()
        h = C.c_void_p()
        check(self.ctx.lib.gapit_geno_from_device(self.ctx.h, C.c_void_p(int(data_ptr)), int(n), int(m),
                                                  int(ld if ld is not None else n), C.byref(h)))
        self.h, self.n, self.m = h, int(n), int(m)
        return self

    @classmethod
    def wrap_device(cls, data_ptr, n, m, ctx=None, keepalive=None):
        """Zero-copy: int8 dosages already on the device in the library's own layout -- marker k at data_ptr + k*ld with
        ld = n rounded up to 128, e.g. a contiguous (m, ld) torch.int8 CUDA tensor whose first n columns hold the data.
        The memory stays the caller's; ``keepalive`` (the tensor) is held by the handle."""
        self = cls.__new__(cls)
        self.ctx = ctx or default_context()
        ld = (int(n) + 127) // 128 * 128
        h = C.c_void_p()
        check(self.ctx.lib.gapit_geno_wrap_device(self.ctx.h, C.c_void_p(int(data_ptr)), int(n), int(m), ld, C.byref(h)))
        self.h, self.n, self.m, self._keepalive = h, int(n), int(m), keepalive
        return self

    def select_taxa(self, GTindex):
        """``GD[GTindex,]`` (GAPIT.EMMAxP3D.R:315): a new handle with the taxa ``GTindex`` (0-based) in that order."""
        idx = np.ascontiguousarray(GTindex, dtype=np.int64).reshape(-1)
        out = Genotypes.__new__(Genotypes)
        h = C.c_void_p()
        check(self.ctx.lib.gapit_geno_select_taxa(self.ctx.h, self.h, _ptr(idx), idx.shape[0], C.byref(h)))
        out.ctx, out.h, out.n, out.m = self.ctx, h, int(idx.shape[0]), self.m
        return out

    def colsums(self):
        cs = np.empty(self.m, dtype=np.int32)
        cq = np.empty(self.m, dtype=np.int32)
        check(self.ctx.lib.gapit_geno_colsums(self.h, _ptr(cs), _ptr(cq)))
        return cs, cq

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.gapit_geno_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class EigenSystem:
    """Eigen-system of K resident on the device, eigenvectors sliced for the int8 rotation."""

    def __init__(self, values, vectors, ctx=None):
        self.ctx = ctx or default_context()
        self.values = np.ascontiguousarray(values, dtype=np.float64)
        v = np.asfortranarray(vectors, dtype=np.float64)
        n = self.values.shape[0]
        assert v.shape == (n, n)
        h = C.c_void_p()
        check(self.ctx.lib.gapit_eigsys_upload(self.ctx.h, _ptr(v), _ptr(self.values), n, C.byref(h)))
        self.h = h
        self.n = n

    @classmethod
    def from_device(cls, values, vectors_ptr, ctx=None):
        """Eigenvectors already on the context's device (n x n column-major, descending; e.g. after ``eigh_dev``)."""
        self = cls.__new__(cls)
        self.ctx = ctx or default_context()
        self.values = np.ascontiguousarray(values, dtype=np.float64)
        n = self.values.shape[0]
        h = C.c_void_p()
        check(self.ctx.lib.gapit_eigsys_from_device(self.ctx.h, C.c_void_p(int(vectors_ptr)), _ptr(self.values), n,
                                                    C.byref(h)))
        self.h, self.n = h, n
        return self

    @classmethod
    def from_kinship(cls, K, ctx=None):
        """eigen(K) of a host kinship decomposed and sliced on the device in one call (gapit_eigsys_from_kinship): the
        eigenvectors never visit the host; ``values`` holds the eigenvalues (descending)."""
        self = cls.__new__(cls)
        self.ctx = ctx or default_context()
        K = np.asfortranarray(K, dtype=np.float64)
        n = K.shape[0]
        assert K.shape == (n, n)
        self.values = np.empty(n)
        h = C.c_void_p()
        check(self.ctx.lib.gapit_eigsys_from_kinship(self.ctx.h, _ptr(K), n, _ptr(self.values), C.byref(h)))
        self.h, self.n = h, n
        return self

    def project(self, R):
        """V'R for an n x nc host matrix (the rotated phenotypes / covariates REML starts from)."""
        R = np.asfortranarray(np.asarray(R, dtype=np.float64).reshape(self.n, -1))
        out = np.empty_like(R, order="F")
        check(self.ctx.lib.gapit_eigsys_project(self.ctx.h, self.h, _ptr(R), R.shape[1], _ptr(out)))
        return out

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.gapit_eigsys_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------
# GAPIT.HapMap(G, SNP.effect, SNP.impute, heading, Create.indicator, Major.allele.zero)   GAPIT.HapMap.R:1-66
# GAPIT.Numericalization(x, bit, effect, impute, Create.indicator, Major.allele.zero)      GAPIT.Numericalization.R
# ---------------------------------------------------------------------------
def _numericalize(ctx, text, row_off, bit, stride, n, m, effect, impute, maz, want_gd, want_geno):
    buf = np.frombuffer(text, dtype=np.uint8)
    off = np.ascontiguousarray(row_off, dtype=np.int64)
    gd = np.empty((m, n), dtype=np.int8) if want_gd else None
    h = C.c_void_p()
    check(ctx.lib.gapit_hapmap_numericalize(ctx.h, _ptr(buf), buf.shape[0], _ptr(off), int(bit), int(stride), n, m,
                                            _lib.EFFECT[effect], _lib.IMPUTE[impute], 1 if maz else 0, _ptr(gd),
                                            C.byref(h) if want_geno else None))
    geno = None
    if want_geno:
        geno = Genotypes.__new__(Genotypes)
        geno.ctx, geno.h, geno.n, geno.m = ctx, h, int(n), int(m)
    return gd, geno


def GAPIT_Numericalization(x, bit=2, effect="Add", impute="None", Create_indicator=False, Major_allele_zero=False,
                           byRow=True, ctx=None):
    """One marker's genotype strings -> n x 1 (or 1 x n) numeric matrix, NaN = NA (GAPIT.Numericalization.R:1-112)."""
    ctx = ctx or default_context()
    codes = [str(v).encode() for v in x]
    if any(len(c) != bit for c in codes):
        raise ValueError("every genotype code must have `bit` characters")
    n = len(codes)
    gd, _ = _numericalize(ctx, b"".join(codes), [0], bit, bit, n, 1, effect, impute, Major_allele_zero, True, False)
    out = np.where(gd[0] < 0, np.nan, gd[0].astype(np.float64))
    return out.reshape(n, 1) if byRow else out.reshape(1, n)


def GAPIT_HapMap(G, SNP_effect="Add", SNP_impute="Middle", heading=True, Create_indicator=False,
                 Major_allele_zero=False, ctx=None, device=False):
    """HapMap table -> {"GT": taxa, "GD": n x m dosages, "GI": (SNP, Chromosome, Position)} (GAPIT.HapMap.R:1-66).

    ``G`` is the text of the table (bytes, or a path to the file): what the reference receives as a data frame from
    ``read.table(file, head=FALSE)``.  The per-marker loop of :34-38 runs on the GPU on the raw text.  GD is int8
    (n, m) with the reference's column-major memory; -1 marks NA (only with ``SNP_impute="None"``).
    ``device=True`` also returns the device-resident ``Genotypes`` under "geno" for the kinship / scan calls."""
    ctx = ctx or default_context()
    if Create_indicator:
        raise NotImplementedError("Create.indicator keeps the character table (GAPIT.HapMap.R:36): reference path")
    if isinstance(G, str):
        with open(G, "rb") as f:
            G = f.read()
    lib = ctx.lib
    buf = np.frombuffer(G, dtype=np.uint8)
    n, m, bit, hoff = C.c_int64(), C.c_int64(), C.c_int(), C.c_int64()
    check(lib.gapit_hapmap_index(_ptr(buf), buf.shape[0], 1 if heading else 0, C.byref(n), C.byref(m), C.byref(bit), None,
                                 0, C.byref(hoff)))
    off = np.empty(m.value, dtype=np.int64)
    check(lib.gapit_hapmap_index(_ptr(buf), buf.shape[0], 1 if heading else 0, C.byref(n), C.byref(m), C.byref(bit),
                                 _ptr(off), m.value, C.byref(hoff)))
    gd, geno = _numericalize(ctx, G, off, bit.value, bit.value + 1, n.value, m.value, SNP_effect, SNP_impute,
                             Major_allele_zero, True, device)
    GT = None
    if heading:                                                            # :14-16
        end = G.find(b"\n", hoff.value)
        GT = G[hoff.value:end if end >= 0 else len(G)].rstrip(b"\r").decode().split("\t")
    GI = []                                                                # :15,:18  columns 1, 3, 4
    for k in range(m.value):
        start = G.rfind(b"\n", 0, off[k]) + 1
        f = G[start:off[k]].split(b"\t", 4)
        GI.append((f[0].decode(), f[2].decode(), f[3].decode()))
    out = {"GT": GT, "GD": gd.T, "GI": GI}
    if device:
        out["geno"] = geno
    return out


def p3d_scan_hapmap(G, X0, Y, eigsys=None, delta=None, vg=None, glm=False, file_fragment=65536, SNP_effect="Add",
                    SNP_impute="Middle", Major_allele_zero=False, heading=True, ctx=None, GTindex=None):
    """The by-file branch of GAPIT.EMMAxP3D (GAPIT.EMMAxP3D.R:262-330) on a HapMap file: fragments of
    ``file_fragment`` markers (GAPIT.Fragment.R:14-77) are numericalized on the GPU straight from the text
    (GAPIT.HapMap.R:34-38) and scanned with the shared eigen-system and variance components; nothing but the file text
    crosses PCIe.  ``G``: bytes of the table or a path.  ``GTindex`` (0-based taxa of the file, in phenotype order) is the
    reference's ``xs = myFRG$GD[GTindex,]`` (:315), applied on the device; without it the file's taxa must already be in
    the order of ``Y``.  Returns (arrays (T, m), per-trait bases, GT, GI)."""
    ctx = ctx or default_context()
    if isinstance(G, str):
        with open(G, "rb") as f:
            G = f.read()
    lib = ctx.lib
    buf = np.frombuffer(G, dtype=np.uint8)
    n, m, bit, hoff = C.c_int64(), C.c_int64(), C.c_int(), C.c_int64()
    check(lib.gapit_hapmap_index(_ptr(buf), buf.shape[0], 1 if heading else 0, C.byref(n), C.byref(m), C.byref(bit), None,
                                 0, C.byref(hoff)))
    off = np.empty(m.value, dtype=np.int64)
    check(lib.gapit_hapmap_index(_ptr(buf), buf.shape[0], 1 if heading else 0, C.byref(n), C.byref(m), C.byref(bit),
                                 _ptr(off), m.value, C.byref(hoff)))
    n, m, bit = n.value, m.value, bit.value
    n_scan = n if GTindex is None else len(GTindex)
    Y = np.asfortranarray(np.asarray(Y, dtype=np.float64).reshape(n_scan, -1))
    T = Y.shape[1]
    names = ("effect", "tvalue", "se", "pvalue", "rsquare", "loglm")
    arrs = {k: np.empty((T, m)) for k in names}
    arrs["maf"] = np.empty(m)
    bases = None
    for k0 in range(0, m, int(file_fragment)):                            # one fragment = one GAPIT.Fragment call
        k1 = min(m, k0 + int(file_fragment))
        _, geno = _numericalize(ctx, G, off[k0:k1], bit, bit + 1, n, k1 - k0, SNP_effect, SNP_impute, Major_allele_zero,
                                False, True)
        if GTindex is not None:
            sel = geno.select_taxa(GTindex)
            geno.close()
            geno = sel
        part, bases = p3d_scan(geno, X0, Y, eigsys=eigsys, delta=delta, vg=vg, glm=glm, ctx=ctx)
        geno.close()
        for k in names:
            arrs[k][:, k0:k1] = part[k]
        arrs["maf"][k0:k1] = part["maf"]
    GT = None
    if heading:
        end = G.find(b"\n", hoff.value)
        GT = G[hoff.value:end if end >= 0 else len(G)].rstrip(b"\r").decode().split("\t")
    GI = []
    for k in range(m):
        start = G.rfind(b"\n", 0, off[k]) + 1
        f = G[start:off[k]].split(b"\t", 4)
        GI.append((f[0].decode(), f[2].decode(), f[3].decode()))
    return arrs, bases, GT, GI


# ---------------------------------------------------------------------------
# Genotype files either side of the path: PLINK .bed / .bim / .fam and GAPIT's numeric text
# (genoFormat = "EMMA": file.GD / file.GM, GAPIT.Fragment.R:94-136).  The bulk bytes go to the GPU as they are.
# ---------------------------------------------------------------------------
def read_plink(prefix, count="A1", mmap=True):
    """``prefix``.bed/.bim/.fam -> (Packed2 rows of the .bed used in place, taxa names GT, GI columns SNP / Chromosome
    / Position, alleles (A1, A2)).  .fam and .bim are small whitespace tables parsed here; the .bed is memory-mapped,
    so a scan streams it from disk in marker blocks (gapit_p3d_scan_host: upload overlapped with the scan)."""
    with open(prefix + ".fam") as f:
        fam = [ln.split() for ln in f if ln.strip()]
    with open(prefix + ".bim") as f:
        bim = [ln.split() for ln in f if ln.strip()]
    GT = [r[1] for r in fam]                                   # within-family ID
    bed = np.memmap(prefix + ".bed", dtype=np.uint8, mode="r") if mmap else np.fromfile(prefix + ".bed", dtype=np.uint8)
    rows = Packed2.from_bed(bed, len(GT), count=count)
    if rows.m != len(bim):
        raise ValueError(f"{prefix}.bed holds {rows.m} variants, {prefix}.bim {len(bim)}")
    GI = {"SNP": [r[1] for r in bim], "Chromosome": [r[0] for r in bim], "Position": [int(r[3]) for r in bim]}
    return rows, GT, GI, [(r[4], r[5]) for r in bim]


def GAPIT_numeric_text(GD, heading=True, SNP_impute=None, ctx=None, want_gd=True, want_geno=False):
    """GAPIT's numeric genotype table (``file.GD``: heading row, then one row per taxon: name, one dosage per marker) ->
    (GD as int8 (n, m) or None, taxa names GT, marker names of the heading or None, Genotypes handle or None).
    ``GD``: the bytes of the table or a path.  Tokenising and parsing happen on the GPU (gapit_numtext_numericalize)."""
    ctx = ctx or default_context()
    if isinstance(GD, str):
        with open(GD, "rb") as f:
            GD = f.read()
    buf = np.frombuffer(GD, dtype=np.uint8)
    lib = ctx.lib
    nl = C.c_int64()
    check(lib.gapit_text_index_lines(_ptr(buf), buf.shape[0], None, 0, C.byref(nl)))
    off = np.empty(nl.value + 1, dtype=np.int64)
    check(lib.gapit_text_index_lines(_ptr(buf), buf.shape[0], _ptr(off), off.shape[0], C.byref(nl)))
    n, m = C.c_int64(), C.c_int64()
    check(lib.gapit_numtext_shape(_ptr(buf), buf.shape[0], _ptr(off), nl.value, 1 if heading else 0, C.byref(n), C.byref(m)))
    n, m = n.value, m.value
    first = 1 if heading else 0
    gd = np.empty((m, n), dtype=np.int8) if want_gd else None
    h = C.c_void_p()
    check(lib.gapit_numtext_numericalize(ctx.h, _ptr(buf), buf.shape[0], _ptr(off), first, n, m, _lib.IMPUTE[SNP_impute],
                                         _ptr(gd), C.byref(h) if want_geno else None))
    geno = None
    if want_geno:
        geno = Genotypes.__new__(Genotypes)
        geno.ctx, geno.h, geno.n, geno.m = ctx, h, int(n), int(m)
    text = bytes(GD)
    names = text[off[0]:off[1]].split()[1:] if heading else None
    GT = [text[off[first + i]:off[first + i + 1]].replace(b",", b" ").split(None, 1)[0].decode() for i in range(n)]
    return (gd.T if gd is not None else None), GT, ([x.decode() for x in names] if names else None), geno


def p3d_scan_file(genotypes, X0, Y, eigsys=None, delta=None, vg=None, glm=False, SNP_impute="Middle", ctx=None,
                  genoFormat=None, count="A1", GTindex=None):
    """The by-file branch of GAPIT.EMMAxP3D (GAPIT.EMMAxP3D.R:259-331 over GAPIT.Fragment.R) for the two file formats
    beside HapMap: a PLINK prefix (``genoFormat="bed"``: the memory-mapped .bed streams through the scan in marker
    blocks, file bytes uploaded as they are, upload overlapped with the scan) or GAPIT's numeric text
    (``genoFormat="EMMA"``: parsed on the GPU, then scanned resident).  ``GTindex`` (0-based taxa of the file in
    phenotype order) applies the reference's ``GD[GTindex,]`` (:315) on the device (the .bed is then packed resident in
    marker blocks instead of streamed).  Returns (arrays (T, m), per-trait bases, GT, GI or marker names)."""
    ctx = ctx or default_context()
    if genoFormat is None:
        genoFormat = "bed" if isinstance(genotypes, str) and os.path.exists(genotypes + ".bed") else "EMMA"
    if genoFormat == "bed":
        rows, GT, GI, _ = read_plink(genotypes, count=count)
        if GTindex is None:
            arrs, bases = p3d_scan_host(rows, X0, Y, eigsys=eigsys, delta=delta, vg=vg, glm=glm, ctx=ctx, impute=SNP_impute)
            return arrs, bases, GT, GI
        Y = np.asfortranarray(np.asarray(Y, dtype=np.float64).reshape(len(GTindex), -1))
        arrs, bases = alloc_scan_out(Y.shape[1], rows.m, pinned=False), None
        frag = 2 * 37888                                          # marker blocks, as GAPIT.Fragment reads them; whole rounds
                                                                  # of the 74 CTA pairs x 512 markers of a 148-SM part
        for k0 in range(0, rows.m, frag):
            k1 = min(rows.m, k0 + frag)
            full = Genotypes(Packed2(rows.data[k0:k1], rows.n, dtype=rows.dtype), ctx=ctx, impute=SNP_impute)
            geno = full.select_taxa(GTindex)
            full.close()
            part, bases = p3d_scan(geno, X0, Y, eigsys=eigsys, delta=delta, vg=vg, glm=glm, ctx=ctx)
            geno.close()
            for k in part:
                if k == "maf":
                    arrs[k][k0:k1] = part[k]
                else:
                    arrs[k][:, k0:k1] = part[k]
        return arrs, bases, GT, GI
    _, GT, names, geno = GAPIT_numeric_text(genotypes, SNP_impute=SNP_impute, ctx=ctx, want_gd=False, want_geno=True)
    if GTindex is not None:
        sel = geno.select_taxa(GTindex)
        geno.close()
        geno = sel
    arrs, bases = p3d_scan(geno, X0, Y, eigsys=eigsys, delta=delta, vg=vg, glm=glm, ctx=ctx)
    geno.close()
    return arrs, bases, GT, names


def _as_geno(snps, ctx, impute=None):
    if isinstance(snps, Genotypes):
        return snps, False
    return Genotypes(snps, ctx=ctx, impute=impute), True


# ---------------------------------------------------------------------------
# GAPIT.kinship.VanRaden(snps, hasInbred=TRUE)      GAPIT.kinship.VanRaden.R:1-37
# ---------------------------------------------------------------------------
def GAPIT_kinship_VanRaden(snps, hasInbred=True, ctx=None, return_gram=False, verbose=True):
    """n x m dosages (array or Genotypes) -> n x n kinship (bare fp64 matrix, like the reference).

    The reference's five progress prints are kept (user-visible behaviour)."""
    ctx = ctx or default_context()
    say = print if verbose else (lambda *_: None)
    say("Calculating kinship with VanRaden method...")
    g, owned = _as_geno(snps, ctx)
    n = g.n
    K = np.empty((n, n), dtype=np.float64, order="F")
    G = np.empty((n, n), dtype=np.int32) if return_gram else None
    say("substracting P...")
    say("Getting X'X...")
    check(ctx.lib.gapit_vanraden(ctx.h, g.h, _ptr(K), _ptr(G)))
    say("Adjusting...")
    if owned:
        g.close()
    say("Calculating kinship with VanRaden method: done")
    return (K, G) if return_gram else K


# ---------------------------------------------------------------------------
# GAPIT.kinship.ZHANG(snps, hasInbred=TRUE)          GAPIT.kinship.ZHANG.R:1-66
# GAPIT.PCA(X, taxa, PC.number, file.output, PCA.total)   GAPIT.PCA.R:1-73
# ---------------------------------------------------------------------------
def GAPIT_kinship_ZHANG(snps, hasInbred=True, ctx=None, verbose=True):
    """n x m dosages (array or Genotypes) -> n x n ZHANG relationship; the reference's progress prints are kept."""
    ctx = ctx or default_context()
    say = print if verbose else (lambda *_: None)
    say("Calculating ZHANG relationship defined by Zhiwu Zhang...")
    g, owned = _as_geno(snps, ctx)
    K = np.empty((g.n, g.n), dtype=np.float64, order="F")
    say("substracting mean...")
    say("Getting X'X...")
    check(ctx.lib.gapit_zhang(ctx.h, g.h, _ptr(K)))
    say("Adjusting...")
    if owned:
        g.close()
    say("Calculating kinship with Zhang method: done")
    return K


def GAPIT_PCA(X, taxa=None, PC_number=None, file_output=False, PCA_total=0, ctx=None, verbose=True):
    """prcomp of the n x m dosage matrix (GAPIT.PCA.R:10).  Returns {"PCs": n x PC_number scores (the reference binds
    `taxa` in front, :54), "EV": all min(n, m) variances (:73), "nPCs": None}.  Plots and csv files (:17-51, :63-69)
    stay in R (file_output must be False here)."""
    if file_output:
        raise NotImplementedError("plots and csv output of GAPIT.PCA stay in R")
    ctx = ctx or default_context()
    say = print if verbose else (lambda *_: None)
    say("Calling prcomp...")
    g, owned = _as_geno(X, ctx)
    rank = min(g.n, g.m)
    k = rank if PC_number is None else int(min(PC_number, rank))
    scores = np.empty((g.n, k), dtype=np.float64, order="F")
    ev = np.empty(rank, dtype=np.float64)
    check(ctx.lib.gapit_pca(ctx.h, g.h, k, _ptr(scores), _ptr(ev)))
    if owned:
        g.close()
    say("Joining taxa...")
    return {"PCs": scores, "taxa": taxa, "EV": ev, "nPCs": None}


# ---------------------------------------------------------------------------
# emma.eigen.L.wo.Z / emma.eigen.R.wo.Z             emma.txt:58-61, 82-89
# ---------------------------------------------------------------------------
def emma_eigen_L_wo_Z(K, ctx=None):
    ctx = ctx or default_context()
    K = np.asfortranarray(K, dtype=np.float64)
    n = K.shape[0]
    values = np.empty(n)
    vectors = np.empty((n, n), order="F")
    check(ctx.lib.gapit_eigh(ctx.h, _ptr(K), n, _ptr(values), _ptr(vectors)))
    return {"values": values, "vectors": vectors}


def eigh_dev(K_ptr, n, ctx=None):
    """In-place eigendecomposition of a device-resident symmetric matrix (gapit_eigh_dev): the n x n column-major
    buffer at ``K_ptr`` becomes the eigenvectors (descending); returns the eigenvalues."""
    ctx = ctx or default_context()
    values = np.empty(int(n))
    check(ctx.lib.gapit_eigh_dev(ctx.h, C.c_void_p(int(K_ptr)), int(n), _ptr(values)))
    return values


def eigh_dev_range(K_ptr, n, first, count, out_ptr, ctx=None):
    """Eigenpairs [first, first + count) of the descending order of a device-resident symmetric matrix
    (gapit_eigh_dev_range, cuSOLVER syevdx by index): the n x count column-major vectors go to the device buffer
    ``out_ptr``, the matrix at ``K_ptr`` is destroyed; returns the ``count`` eigenvalues.  Several GPUs holding the
    same matrix each take one slice of the spectrum (gapit_b200/fullsize.py)."""
    ctx = ctx or default_context()
    values = np.empty(int(count))
    check(ctx.lib.gapit_eigh_dev_range(ctx.h, C.c_void_p(int(K_ptr)), int(n), int(first), int(count), _ptr(values),
                                       C.c_void_p(int(out_ptr))))
    return values


def reml_from_projections(values, Vty, VtX, ngrids=100, llim=-10, ulim=10, esp=1e-10):
    """gapit_reml_eigL on rotated inputs (EigenSystem.project): {"REML", "delta", "ve", "vg"}."""
    lam = np.ascontiguousarray(values, dtype=np.float64)
    Vty = np.ascontiguousarray(Vty, dtype=np.float64).reshape(-1)
    VtX = np.asfortranarray(np.asarray(VtX, dtype=np.float64).reshape(lam.shape[0], -1))
    out = _lib.RemlOut()
    check(_lib.load().gapit_reml_eigL(_ptr(lam), _ptr(Vty), _ptr(VtX), lam.shape[0], VtX.shape[1], int(ngrids),
                                      float(llim), float(ulim), float(esp), C.byref(out)))
    return {"REML": out.REML, "delta": out.delta, "ve": out.ve, "vg": out.vg}


def reml_from_projections_multi(values, VtY, VtX, ngrids=100, llim=-10, ulim=10, esp=1e-10):
    """T traits at once (VtY: n x T): list of {"REML", "delta", "ve", "vg"}; the searches run on the host's threads."""
    lam = np.ascontiguousarray(values, dtype=np.float64)
    n = lam.shape[0]
    VtY = np.asfortranarray(np.asarray(VtY, dtype=np.float64).reshape(n, -1))
    VtX = np.asfortranarray(np.asarray(VtX, dtype=np.float64).reshape(n, -1))
    T = VtY.shape[1]
    outs = (_lib.RemlOut * T)()
    check(_lib.load().gapit_reml_eigL_multi(_ptr(lam), _ptr(VtY), _ptr(VtX), n, VtX.shape[1], T, int(ngrids), float(llim),
                                            float(ulim), float(esp), outs, None))
    return [{"REML": o.REML, "delta": o.delta, "ve": o.ve, "vg": o.vg} for o in outs]


def emma_eigen_R_wo_Z(K, X, ctx=None):
    ctx = ctx or default_context()
    K = np.asfortranarray(K, dtype=np.float64)
    X = np.asfortranarray(X, dtype=np.float64)
    n, q = X.shape
    values = np.empty(n - q)
    vectors = np.empty((n, n - q), order="F")
    check(ctx.lib.gapit_eigen_R(ctx.h, _ptr(K), _ptr(X), n, q, _ptr(values), _ptr(vectors)))
    return {"values": values, "vectors": vectors}


# ---------------------------------------------------------------------------
# GAPIT.emma.REMLE(y, X, K, Z=NULL, ngrids, llim, ulim, esp, eig.L, eig.R)
#                                                    GAPIT.emma.REMLE.R:1-111
# ---------------------------------------------------------------------------
def GAPIT_emma_REMLE(y, X, K, Z=None, ngrids=100, llim=-10, ulim=10, esp=1e-10, eig_L=None, eig_R=None, ctx=None):
    if Z is not None:
        raise NotImplementedError("compressed MLM (non-NULL Z) keeps the reference's R body; see DESIGN.md")
    y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
    X = np.asfortranarray(X, dtype=np.float64)
    n, q = X.shape
    if np.linalg.det(X.T @ X) == 0:                      # GAPIT.emma.REMLE.R:15-18
        return {"REML": 0.0, "delta": 0.0, "ve": 0.0, "vg": 0.0}
    out = _lib.RemlOut()
    if eig_R is None and eig_L is not None:
        # eig.L given, eig.R not: the same likelihood from the spectrum of K alone (gapit_reml_eigL) --
        # no second eigendecomposition
        V = eig_L["vectors"]
        lam = np.ascontiguousarray(eig_L["values"], dtype=np.float64)
        Vty = np.ascontiguousarray(V.T @ y)
        VtX = np.asfortranarray(V.T @ X)
        check(_lib.load().gapit_reml_eigL(_ptr(lam), _ptr(Vty), _ptr(VtX), n, q, int(ngrids), float(llim), float(ulim),
                                          float(esp), C.byref(out)))
        return {"REML": out.REML, "delta": out.delta, "ve": out.ve, "vg": out.vg}
    if eig_R is None:                                    # :22-24
        eig_R = emma_eigen_R_wo_Z(K, X, ctx=ctx)
    lam = np.ascontiguousarray(eig_R["values"], dtype=np.float64)
    etas = np.ascontiguousarray(eig_R["vectors"].T @ y)  # :25 (n-q dot products; host)
    check(_lib.load().gapit_reml(_ptr(lam), _ptr(etas), lam.shape[0], int(ngrids), float(llim), float(ulim),
                                 float(esp), C.byref(out)))
    return {"REML": out.REML, "delta": out.delta, "ve": out.ve, "vg": out.vg}


# ---------------------------------------------------------------------------
# the marker scan proper
# ---------------------------------------------------------------------------
def _host_array(shape, pinned):
    if pinned:
        import torch
        return torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()
    return np.empty(shape, dtype=np.float64)


def alloc_scan_out(T, m, pinned=True):
    """Result arrays of a scan of m markers x T traits, for the ``out=`` argument of :func:`p3d_scan`: page-locking
    gigabytes of host memory takes seconds, so callers that scan block after block allocate once."""
    arrs = {k: _host_array((T, m), pinned) for k in ("effect", "tvalue", "se", "pvalue", "rsquare", "loglm")}
    arrs["maf"] = _host_array((m,), pinned)
    return arrs


def _check_scan_shapes(n, X0, T, delta, vg, glm):
    """The C ABI takes raw pointers: row counts and per-trait vector lengths are checked here, before it reads them."""
    if X0.ndim != 2 or X0.shape[0] != n:
        raise ValueError(f"X0 must have {n} rows (one per taxon), got shape {X0.shape}")
    if not glm:
        for name, v in (("delta", delta), ("vg", vg)):
            if v is not None and np.size(v) != T:
                raise ValueError(f"{name} must hold one value per trait ({T}), got {np.size(v)}")


def p3d_scan(geno, X0, Y, eigsys=None, delta=None, vg=None, glm=False, ctx=None, pinned=False, out=None):
    """Scan all markers of ``geno`` for T traits.  Returns (dict of (T, m) arrays, list of per-trait base dicts).
    ``pinned=True`` puts the result arrays in page-locked host memory (needs torch) so the D2H copy runs at
    full PCIe rate; ``out`` reuses the array dict of an earlier call of the same shape (page-locking gigabytes per call
    costs more than the copy)."""
    ctx = ctx or geno.ctx
    X0 = np.asfortranarray(X0, dtype=np.float64)
    Y = np.asfortranarray(np.asarray(Y, dtype=np.float64).reshape(geno.n, -1))
    n, q0 = X0.shape
    T = Y.shape[1]
    m = geno.m
    _check_scan_shapes(geno.n, X0, T, delta, vg, glm)
    names = ("effect", "tvalue", "se", "pvalue", "rsquare", "loglm")
    if out is not None and out["effect"].shape == (T, m) and out["maf"].shape == (m,):
        arrs = out
    else:
        arrs = {k: _host_array((T, m), pinned) for k in names}
        arrs["maf"] = _host_array((m,), pinned)
    so = _lib.ScanOut(**{k: arrs[k].ctypes.data_as(C.POINTER(C.c_double)) for k in arrs})
    base = (_lib.ScanBase * T)()
    d = np.ascontiguousarray(delta, dtype=np.float64).reshape(-1) if delta is not None else None
    v = np.ascontiguousarray(vg, dtype=np.float64).reshape(-1) if vg is not None else None
    check(ctx.lib.gapit_p3d_scan_dev(ctx.h, geno.h, eigsys.h if eigsys is not None else None, _ptr(X0), q0, _ptr(Y), T,
                                     _ptr(d), _ptr(v), 1 if glm else 0, C.byref(so), base))
    bases = []
    for t in range(T):
        b = base[t]
        bases.append({"logL0": b.logL0, "logLM_base": b.logLM_base, "rsquare_base": b.rsquare_base, "ves": b.ves,
                      "reml": b.reml, "df": b.df, "beta0": np.array(b.beta0[:q0]), "singular_x0": bool(b.singular_x0)})
    return arrs, bases


def _host_geno(snps, marker_major):
    """-> (array (m, n) C-contiguous, dtype code) exactly as the C ABI wants it (ld = n)."""
    a = np.asarray(snps)
    if not marker_major:
        a = a.T
    if a.dtype == np.int8 or a.dtype == np.uint8:
        dtype = _lib.GAPIT_I8
    elif a.dtype == np.int32:
        dtype = _lib.GAPIT_I32
    else:
        a = a.astype(np.float64, copy=False)
        dtype = _lib.GAPIT_F64
    return np.ascontiguousarray(a), dtype


def p3d_scan_host(snps, X0, Y, eigsys=None, delta=None, vg=None, glm=False, ctx=None, impute=None, marker_major=False,
                  pinned=False, out=None):
    """Streaming scan straight from a host genotype matrix (H2D overlapped with the scan).  Same returns as
    :func:`p3d_scan`."""
    ctx = ctx or default_context()
    if isinstance(snps, Packed2):
        a, dtype, m, n, ld = snps.data, snps.dtype, snps.m, snps.n, snps.data.shape[1]
    else:
        a, dtype = _host_geno(snps, marker_major)
        m, n = a.shape
        ld = n
    X0 = np.asfortranarray(X0, dtype=np.float64)
    Y = np.asfortranarray(np.asarray(Y, dtype=np.float64).reshape(n, -1))
    q0, T = X0.shape[1], Y.shape[1]
    _check_scan_shapes(n, X0, T, delta, vg, glm)
    if out is not None and out["effect"].shape == (T, m) and out["maf"].shape == (m,):
        arrs = out
    else:
        arrs = alloc_scan_out(T, m, pinned)
    so = _lib.ScanOut(**{k: arrs[k].ctypes.data_as(C.POINTER(C.c_double)) for k in arrs})
    base = (_lib.ScanBase * T)()
    d = np.ascontiguousarray(delta, dtype=np.float64).reshape(-1) if delta is not None else None
    v = np.ascontiguousarray(vg, dtype=np.float64).reshape(-1) if vg is not None else None
    check(ctx.lib.gapit_p3d_scan_host(ctx.h, _ptr(a), dtype, n, m, ld, _lib.IMPUTE[impute],
                                      eigsys.h if eigsys is not None else None, _ptr(X0), q0, _ptr(Y), T, _ptr(d), _ptr(v),
                                      1 if glm else 0, C.byref(so), base))
    bases = [{"logL0": b.logL0, "logLM_base": b.logLM_base, "rsquare_base": b.rsquare_base, "ves": b.ves,
              "reml": b.reml, "df": b.df, "beta0": np.array(b.beta0[:q0]), "singular_x0": bool(b.singular_x0)}
             for b in base]
    return arrs, bases


def blup_pev(eigsys, X0, y, delta, vg, ctx=None):
    """i = 0 BLUP / PEV block of GAPIT.EMMAxP3D (GAPIT.EMMAxP3D.R:779-839) in the eigenbasis of K.
    Returns (BLUP, PEV, beta0)."""
    ctx = ctx or eigsys.ctx
    X0 = np.asfortranarray(X0, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
    n, q0 = X0.shape
    blup, pev, beta = np.empty(n), np.empty(n), np.empty(q0)
    check(ctx.lib.gapit_blup_pev(ctx.h, eigsys.h, _ptr(X0), q0, _ptr(y), float(delta), float(vg), _ptr(blup), _ptr(pev),
                                 _ptr(beta)))
    return blup, pev, beta


def _timmer(Timmer, label):
    """GAPIT.Timmer.R:8-17: append (label, now, elapsed since previous stamp)."""
    now = time.time()
    if Timmer is None:
        return [(label, now, 0.0)]
    return Timmer + [(label, now, now - Timmer[-1][1])]


# ---------------------------------------------------------------------------
# GAPIT.EMMAxP3D(ys, xs, K, Z, X0, CVI, ..., SNP.P3D, Timmer, Memory, optOnly, ...)
#                                                    GAPIT.EMMAxP3D.R:1-1022
# ---------------------------------------------------------------------------
def GAPIT_EMMAxP3D(ys, xs, K=None, Z=None, X0=None, CVI=None, CV_Inheritance=None, GI=None, GP=None,
                   fullGD=True, SNP_P3D=True, Timmer=None, Memory=None, optOnly=False, SNP_effect="Add",
                   SNP_impute="Middle", ngrids=100, llim=-10, ulim=10, esp=1e-10, name_of_trait=None,
                   Create_indicator=False, ctx=None, eig_L=None, reml=None, reml_route="eigL"):
    """The accelerated branch of GAPIT.EMMAxP3D: Z NULL or square, SNP.P3D, fullGD, no indicators, !optOnly.

    Returns the reference's named list (GAPIT.EMMAxP3D.R:1018-1020) as a dict; per-marker
    entries are (m, g) arrays like the reference's m x g matrices (g traits = rows of ``ys``; the scalars vgs / ves /
    REMLs / BLUP / effect.cv are those of the last trait, as the reference's loop leaves them).  ``eig_L`` /
    ``reml`` may be injected (tests feed both sides identical spectra); they are otherwise
    computed where the reference computes them (:48, :58, :202).  ``reml_route="eigL"`` (default) evaluates the REML
    likelihood from eig.L alone (``gapit_reml_eigL``: same function of delta, measured agreement 1e-13 in delta), so
    the eig.R decomposition of :58 is skipped; ``reml_route="eigR"`` runs the reference's two-decomposition route.
    Other argument combinations are the reference's other code paths and are not provided here.
    """
    ctx = ctx or default_context()
    if Z is not None and Z.shape[0] != Z.shape[1]:
        raise NotImplementedError("compressed MLM (rectangular Z) keeps the reference's R body")
    if not SNP_P3D or not fullGD or Create_indicator or optOnly:
        raise NotImplementedError("only the P3D / fullGD / additive scan is accelerated (SURVEY.md section 8b)")
    Timmer = _timmer(Timmer, "P3D Start")
    # :29  ys is a g x n matrix of g traits (a vector is one trait).  The reference loops the traits (:81) and repeats the
    # rotation for each; here all g traits share one eigendecomposition and ONE rotation (BASELINE configs[4]).
    # Missing phenotypes (per-trait subsetting, :113-114) keep the reference body.
    ys = np.asarray(ys, dtype=np.float64)
    ys = ys.reshape(1, -1) if (ys.ndim == 1 or 1 in ys.shape) else ys
    if np.isnan(ys).any():
        raise NotImplementedError("missing phenotypes are subset per trait by the reference body (GAPIT.EMMAxP3D.R:113-114)")
    g, n = ys.shape
    Y = np.asfortranarray(ys.T)                                           # n x g
    if X0 is None:                                                        # :30
        X0 = np.ones((n, 1))
    X0 = np.asarray(X0, dtype=np.float64)
    if K is not None and np.size(K) <= 1:                                 # :34
        K = None
    q0 = X0.shape[1]
    # :20-24 (NA -> SNP.impute value).  A host matrix is streamed (upload overlapped with the scan); an
    # already-resident Genotypes handle is scanned in place.
    resident = isinstance(xs, Genotypes)
    m = xs.m if isinstance(xs, (Genotypes, Packed2)) else np.asarray(xs).shape[1]

    def run_scan(**kw):
        if resident:
            return p3d_scan(xs, X0, Y, ctx=ctx, **kw)
        return p3d_scan_host(xs, X0, Y, ctx=ctx, impute=SNP_impute, **kw)
    glm = K is None
    y_last = np.ascontiguousarray(Y[:, g - 1])
    if not glm:
        if eig_L is None and reml_route == "eigL":
            es = EigenSystem.from_kinship(K, ctx=ctx)                     # :48, and :224 folded into the kernel weights;
            Timmer = _timmer(Timmer, "eig.L")                             # the eigenvectors stay on the device
        else:
            if eig_L is None:
                eig_L = emma_eigen_L_wo_Z(K, ctx=ctx)                     # :48
            Timmer = _timmer(Timmer, "eig.L")
            es = EigenSystem(eig_L["values"], eig_L["vectors"], ctx=ctx)  # :224 folded into the kernel weights
        if reml is not None:
            remls = [reml] if isinstance(reml, dict) else list(reml)
        elif reml_route == "eigL":                                        # :202 for every trait, from eig.L alone
            proj = es.project(np.column_stack([X0, Y]))
            remls = reml_from_projections_multi(es.values, proj[:, q0:], proj[:, :q0], ngrids, llim, ulim, esp)
        else:
            try:
                eig_R = emma_eigen_R_wo_Z(K, X0, ctx=ctx)                 # :58
            except GapitCudaError:
                # :70-74 early return when eig.R fails
                es.close()
                return {"ps": None, "REMLs": np.nan, "stats": None, "effect.est": None, "dfs": None, "maf": None,
                        "nobs": None, "Timmer": Timmer, "Memory": Memory, "vgs": np.nan, "ves": np.nan,
                        "BLUP": None, "BLUP_Plus_Mean": None, "PEV": None, "BLUE": None}
            Timmer = _timmer(Timmer, "eig.R")
            remls = [GAPIT_emma_REMLE(Y[:, t], X0, K, None, ngrids, llim, ulim, esp, eig_R=eig_R, ctx=ctx) for t in range(g)]
        if len(remls) != g:
            raise ValueError(f"{g} traits need {g} REML results, got {len(remls)}")
        Timmer = _timmer(Timmer, "REML")
        Timmer = _timmer(Timmer, "U Matrix")
        # i = 0: BLUP / PEV of the marker-free model (:779-839) -- of the LAST trait, which is what the reference's
        # loop leaves in its scalars -- then the marker loop of all traits
        last = remls[-1]
        BLUP, PEV, beta0 = blup_pev(es, X0, y_last, last["delta"], last["vg"], ctx=ctx)
        Timmer = _timmer(Timmer, "PEV")
        arrs, bases = run_scan(eigsys=es, delta=[r["delta"] for r in remls], vg=[r["vg"] for r in remls])
        es.close()
        vgs, ves, REMLs = last["vg"], last["ve"], last["REML"]
    else:
        arrs, bases = run_scan(glm=True)
        vgs, ves, REMLs = 0.0, bases[-1]["ves"], bases[-1]["reml"]         # :868-872
        BLUP, PEV = 0.0, ves                                              # :873-875
        beta0 = bases[-1]["beta0"]
    Timmer = _timmer(Timmer, "Screening SNPs")
    if bases[0]["singular_x0"]:                                           # :711
        print("At least two of your covariates are linearly dependent. Please reconsider the covariates you are using for GWAS and GPS")
    mat = lambda a: np.ascontiguousarray(np.asarray(a).T)                 # (g, m) -> the reference's m x g
    stats = arrs["tvalue"]
    normal = ~np.isnan(stats)
    dfs = np.where(normal, np.array([b["df"] for b in bases])[:, None], np.nan)
    # :972 stderr = beta[ncol(CVI)+1] / t -- an index into the (q0+1)-vector of the full model that only
    # lands on the marker effect when ncol(CVI) == q0; past the end R yields NA.
    stderr = np.full((g, m), np.nan)
    if CVI is not None:
        idx = np.asarray(CVI).shape[1]                                    # 0-based position ncol(CVI)
        if idx == q0:
            stderr = arrs["effect"] / stats
        # idx < q0 would need the per-marker covariate effects the scan does not return
    rsq_base = np.where(normal, np.array([b["rsquare_base"] for b in bases])[:, None], np.nan)
    # :841-861 / :877-889  BLUE = XCV %*% beta.Inheritance with XCV = cbind(1, CVI[,-1]) (X itself for an intercept-only fit)
    BLUE = None
    if q0 == 1:
        BLUE = (X0 @ beta0).reshape(n, 1)
    elif CVI is not None:
        XCV = np.column_stack([np.ones(n), np.asarray(CVI, dtype=np.float64)[:, 1:]])
        b_inh = beta0
        if CV_Inheritance is not None:
            XCV, b_inh = XCV[:, : 1 + CV_Inheritance], beta0[: 1 + CV_Inheritance]
        if XCV.shape[1] == b_inh.shape[0]:
            BLUE = (XCV @ b_inh).reshape(n, 1)
    Timmer = _timmer(Timmer, "GWAS done for this Trait")
    return {
        "ps": mat(arrs["pvalue"]), "REMLs": -2.0 * REMLs, "stats": mat(stats),
        "effect.est": mat(arrs["effect"]), "rsquare_base": mat(rsq_base), "rsquare": mat(arrs["rsquare"]),
        "dfs": mat(dfs), "df": mat(dfs), "tvalue": mat(stats), "stderr": mat(stderr),
        "maf": np.repeat(arrs["maf"].reshape(m, 1), g, axis=1), "nobs": np.full((m, g), float(n)), "Timmer": Timmer,
        "Memory": Memory, "vgs": vgs, "ves": ves,
        # i = 0 block (:779-861): BLUP = K H^-1 (y - X beta), PEV = diag(C22), BLUE = XCV beta
        "BLUP": BLUP if glm else BLUP.reshape(n, 1),
        "BLUP_Plus_Mean": (np.nan if glm else (beta0[0] + BLUP).reshape(n, 1)),           # :818-819, :874
        "PEV": PEV if glm else PEV.reshape(n, 1), "BLUE": BLUE,
        "logLM": mat(arrs["loglm"]), "effect.snp": mat(arrs["effect"]), "effect.cv": bases[-1]["beta0"],
        "se": mat(arrs["se"]), "delta": (None if glm else [r["delta"] for r in remls] if g > 1 else remls[0]["delta"]),
    }


# ---------------------------------------------------------------------------
# Compression glue GAPIT.Main.R:386-400 runs before every GAPIT.EMMAxP3D call.
# Host-side table bookkeeping, no kernel: the point is NOT to run dist() + hclust() (O(n^3) scalar R,
# GAPIT.Compress.R:37-38) nor the dense Z %*% DS (GAPIT.ZmatrixCompress.R:34) at the two group settings the
# MLM (GN = n) and GLM (GN = 1) models use, where their results are known in closed form.
# Tables: KI = (names, n x n array); GA rows = (taxon, group); GAU rows = (taxon, group, block, ID);
# Z = (row taxa, column taxa, array).
# ---------------------------------------------------------------------------
def GAPIT_Compress(line_names, KI, kinship_cluster="average", kinship_group="Mean", GN=None):
    """GAPIT.Compress at GN >= n (every taxon its own group, numbered in row order as cutree numbers clusters by
    first appearance; KG = KI, :59-61 with b = I) and GN == 1 (one group; KG = the 1 x 1 Mean / Max / Min / Median
    of KI).  Other group counts are the compressed-MLM path, which stays with the reference's R body."""
    KI = np.asarray(KI, dtype=np.float64)
    n = KI.shape[0]
    GN = n if GN is None else int(GN)
    if GN >= n:
        return [(str(t), i + 1) for i, t in enumerate(line_names)], KI.copy()
    if GN == 1:
        f = {"Mean": np.mean, "Max": np.max, "Min": np.min, "Median": np.median}[kinship_group]
        return [(str(t), 1) for t in line_names], np.array([[f(KI)]])
    raise NotImplementedError("compressed MLM (1 < GN < n) keeps the reference's GAPIT.Compress body")


def GAPIT_Block(Z_col_taxa, GA, KG):
    """GAPIT.Block (GAPIT.Block.R:14-64): groups with a phenotyped member -> KW, the rest -> KO; GAU rows in
    merge()'s order (group key sorted as character)."""
    zset = set(str(t) for t in Z_col_taxa)
    blk = [1 if t in zset else 2 for t, _ in GA]
    grp1 = list(dict.fromkeys(g for (_, g), b in zip(GA, blk) if b == 1))
    in1 = set(grp1)
    grp2 = [g for g in dict.fromkeys(g for _, g in GA) if g not in in1]
    gid = {g: (1, i + 1) for i, g in enumerate(grp1)}
    gid.update({g: (2, i + 1) for i, g in enumerate(grp2)})
    rows = sorted(range(len(GA)), key=lambda i: blk[i])
    rows.sort(key=lambda i: str(GA[i][1]))
    GAU = [(GA[i][0], GA[i][1], (blk[i] + gid[GA[i][1]][0]) / 2.0, gid[GA[i][1]][1]) for i in rows]
    i1 = np.asarray(grp1, dtype=np.int64) - 1
    i2 = np.asarray(grp2, dtype=np.int64) - 1
    KG = np.asarray(KG)
    return GAU, KG[np.ix_(i1, i1)], KG[np.ix_(i2, i2)], KG[np.ix_(i1, i2)]


def GAPIT_ZmatrixCompress(Z_row_taxa, Z_col_taxa, Zmat, GAU):
    """GAPIT.ZmatrixCompress: Z %*% DS with DS = diag(n)[x,] (:24,:34) is a column scatter-add -- O(n^2), not the
    reference's dense n^3 product; rows sorted by taxon (:41)."""
    zset = set(str(t) for t in Z_col_taxa)
    GAU1 = sorted((r for r in GAU if r[0] in zset), key=lambda r: r[0])
    n = max(r[3] for r in GAU1 if r[2] < 2)
    x = np.asarray([r[3] for r in GAU1], dtype=np.int64) - 1
    order_Z = sorted(range(len(Z_col_taxa)), key=lambda i: str(Z_col_taxa[i]))
    Zs = np.asarray(Zmat, dtype=np.float64)[:, order_Z]
    out = np.zeros((Zs.shape[0], n))
    if len(set(x.tolist())) == len(x):
        out[:, x] = Zs
    else:
        np.add.at(out.T, x, Zs.T)
    order_rows = sorted(range(len(Z_row_taxa)), key=lambda i: str(Z_row_taxa[i]))
    return [str(Z_row_taxa[i]) for i in order_rows], out[order_rows, :]


# ---------------------------------------------------------------------------
# All GPUs of the box from one process (gapit_mctx / gapit_mgeno): what the R shim binds, mirrored for the tests.
# ---------------------------------------------------------------------------
class MultiContext:
    """gapit_mctx: ``ngpus`` devices (0 = all visible) with peer access between every pair."""

    def __init__(self, ngpus=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        check(self.lib.gapit_mctx_create(int(ngpus), C.byref(h)))
        self.h = h
        self.ngpus = int(self.lib.gapit_mctx_ngpus(h))

    def context(self, gpu=0):
        """The one-GPU context of device ``gpu`` (borrowed: for spectra / REML between the two multi-GPU calls)."""
        c = Context.__new__(Context)
        c.lib, c.h, c.device = self.lib, C.c_void_p(self.lib.gapit_mctx_ctx(self.h, int(gpu))), int(gpu)
        c.close = lambda: None          # owned by the multi-context
        return c

    def close(self):
        if getattr(self, "h", None):
            self.lib.gapit_mctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiGenotypes:
    """gapit_mgeno: the host matrix split into contiguous marker blocks, one per GPU."""

    def __init__(self, snps, mctx, impute=None, marker_major=False):
        self.mctx = mctx
        if isinstance(snps, Packed2):
            a, dtype, m, n, ld = snps.data, snps.dtype, snps.m, snps.n, snps.data.shape[1]
        else:
            a, dtype = _host_geno(snps, marker_major)
            m, n = a.shape
            ld = n
        h = C.c_void_p()
        check(mctx.lib.gapit_mgeno_pack(mctx.h, _ptr(a), dtype, n, m, ld, _lib.IMPUTE[impute], C.byref(h)))
        self.h, self.n, self.m = h, int(n), int(m)

    def close(self):
        if getattr(self, "h", None):
            self.mctx.lib.gapit_mgeno_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def multi_vanraden(mgeno, return_gram=False):
    """GAPIT.kinship.VanRaden on every GPU of the multi-context (gapit_multi_vanraden)."""
    n = mgeno.n
    K = np.empty((n, n), dtype=np.float64, order="F")
    G = np.empty((n, n), dtype=np.int32) if return_gram else None
    check(mgeno.mctx.lib.gapit_multi_vanraden(mgeno.mctx.h, mgeno.h, _ptr(K), _ptr(G)))
    return (K, G) if return_gram else K


class MultiEigenSystem:
    """gapit_meigsys: the eigen-system of K resident (and sliced) on every GPU of a multi-context.  Built from the host
    kinship (decomposed on the GPUs, spectrum sliced over them, slices exchanged over NVLink) or from host eigenvectors."""

    def __init__(self, mctx, K=None, values=None, vectors=None):
        self.mctx = mctx
        h = C.c_void_p()
        if K is not None:
            K = np.asfortranarray(K, dtype=np.float64)
            n = K.shape[0]
            self.values = np.empty(n)
            check(mctx.lib.gapit_meigsys_from_kinship(mctx.h, _ptr(K), n, _ptr(self.values), C.byref(h)))
        else:
            self.values = np.ascontiguousarray(values, dtype=np.float64)
            v = np.asfortranarray(vectors, dtype=np.float64)
            n = self.values.shape[0]
            assert v.shape == (n, n)
            check(mctx.lib.gapit_meigsys_upload(mctx.h, _ptr(v), _ptr(self.values), n, C.byref(h)))
        self.h, self.n = h, int(n)

    def part(self, gpu=0):
        """The one-GPU EigenSystem on device ``gpu`` (borrowed: projections, BLUP/PEV)."""
        e = EigenSystem.__new__(EigenSystem)
        e.ctx, e.values, e.n = self.mctx.context(gpu), self.values, self.n
        e.h = C.c_void_p(self.mctx.lib.gapit_meigsys_part(self.h, int(gpu)))
        e.close = lambda: None
        return e

    def close(self):
        if getattr(self, "h", None):
            self.mctx.lib.gapit_meigsys_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def multi_eigh(K, mctx):
    """emma.eigen.L.wo.Z (emma.txt:58-61) with the spectrum sliced over the GPUs of the multi-context (gapit_multi_eigh)."""
    K = np.asfortranarray(K, dtype=np.float64)
    n = K.shape[0]
    values = np.empty(n)
    vectors = np.empty((n, n), order="F")
    check(mctx.lib.gapit_multi_eigh(mctx.h, _ptr(K), n, _ptr(values), _ptr(vectors)))
    return {"values": values, "vectors": vectors}


def multi_p3d_scan(mgeno, X0, Y, values=None, vectors=None, delta=None, vg=None, glm=False, pinned=False, eigsys=None):
    """The P3D scan of every marker block on its GPU (gapit_multi_p3d_scan, or gapit_multi_p3d_scan_es against a
    resident :class:`MultiEigenSystem`); same returns as :func:`p3d_scan`."""
    n, m = mgeno.n, mgeno.m
    X0 = np.asfortranarray(X0, dtype=np.float64)
    Y = np.asfortranarray(np.asarray(Y, dtype=np.float64).reshape(n, -1))
    q0, T = X0.shape[1], Y.shape[1]
    _check_scan_shapes(n, X0, T, delta, vg, glm)
    names = ("effect", "tvalue", "se", "pvalue", "rsquare", "loglm")
    arrs = {k: _host_array((T, m), pinned) for k in names}
    arrs["maf"] = _host_array((m,), pinned)
    so = _lib.ScanOut(**{k: arrs[k].ctypes.data_as(C.POINTER(C.c_double)) for k in arrs})
    base = (_lib.ScanBase * T)()
    d = np.ascontiguousarray(delta, dtype=np.float64).reshape(-1) if delta is not None else None
    v = np.ascontiguousarray(vg, dtype=np.float64).reshape(-1) if vg is not None else None
    vec = np.asfortranarray(vectors, dtype=np.float64) if vectors is not None else None
    val = np.ascontiguousarray(values, dtype=np.float64) if values is not None else None
    if eigsys is not None or glm:
        check(mgeno.mctx.lib.gapit_multi_p3d_scan_es(mgeno.mctx.h, mgeno.h, eigsys.h if eigsys is not None else None, _ptr(X0),
                                                     q0, _ptr(Y), T, _ptr(d), _ptr(v), 1 if glm else 0, C.byref(so), base))
    else:
        check(mgeno.mctx.lib.gapit_multi_p3d_scan(mgeno.mctx.h, mgeno.h, _ptr(vec), _ptr(val), _ptr(X0), q0, _ptr(Y), T,
                                                  _ptr(d), _ptr(v), 0, C.byref(so), base))
    bases = [{"logL0": b.logL0, "logLM_base": b.logLM_base, "rsquare_base": b.rsquare_base, "ves": b.ves,
              "reml": b.reml, "df": b.df, "beta0": np.array(b.beta0[:q0]), "singular_x0": bool(b.singular_x0)}
             for b in base]
    return arrs, bases
