"""ctypes binding of libgapitcuda.so (include/gapitcuda.h).

The library is the product: there is no Python or CPU fallback.  Loading
fails loudly when the shared object is missing, and every compute entry point
returns an error (raised as ``GapitCudaError``) when no sm_100 GPU is present.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgapitcuda.so")

GAPIT_MAX_Q0 = 16
GAPIT_F64, GAPIT_I32, GAPIT_I8, GAPIT_B2, GAPIT_BED_A1, GAPIT_BED_A2 = 0, 1, 2, 3, 4, 5
IMPUTE = {None: -1, "None": -1, "Minor": 0, "Middle": 1, "Major": 2}
EFFECT = {"Add": 0, "Dom": 1, "Left": 2, "Right": 3}


class GapitCudaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libgapitcuda error {code}: {msg}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("h2d_ms", "pack_ms", "kernel_ms", "post_ms", "d2h_ms", "total_ms")]


class RemlOut(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("REML", "delta", "ve", "vg")]


class ScanOut(C.Structure):
    _fields_ = [(k, C.POINTER(C.c_double)) for k in ("effect", "tvalue", "se", "pvalue", "rsquare", "loglm", "maf")]


class ScanBase(C.Structure):
    _fields_ = [("logL0", C.c_double), ("logLM_base", C.c_double), ("rsquare_base", C.c_double),
                ("ves", C.c_double), ("reml", C.c_double), ("df", C.c_double),
                ("beta0", C.c_double * GAPIT_MAX_Q0), ("singular_x0", C.c_int)]


_P = C.c_void_p
_I64 = C.c_int64
_DP = C.POINTER(C.c_double)

# name -> (restype, argtypes); every symbol include/gapitcuda.h declares
SIGNATURES = {
    "gapit_last_error": (C.c_char_p, []),
    "gapit_device_count": (C.c_int, []),
    "gapit_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "gapit_ctx_destroy": (None, [_P]),
    "gapit_ctx_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "gapit_ctx_set_stream": (C.c_int, [_P, _P]),
    "gapit_geno_pack": (C.c_int, [_P, _P, C.c_int, _I64, _I64, _I64, C.c_int, C.POINTER(_P)]),
    "gapit_pack2_host": (C.c_int, [_P, C.c_int, _I64, _I64, _I64, _P, C.c_int]),
    "gapit_geno_from_device": (C.c_int, [_P, _P, _I64, _I64, _I64, C.POINTER(_P)]),
    "gapit_geno_wrap_device": (C.c_int, [_P, _P, _I64, _I64, _I64, C.POINTER(_P)]),
    "gapit_geno_select_taxa": (C.c_int, [_P, _P, _P, _I64, C.POINTER(_P)]),
    "gapit_geno_free": (None, [_P]),
    "gapit_geno_n": (_I64, [_P]),
    "gapit_geno_m": (_I64, [_P]),
    "gapit_geno_colsums": (C.c_int, [_P, _P, _P]),
    "gapit_hapmap_index": (C.c_int, [_P, _I64, C.c_int, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(C.c_int), _P, _I64,
                                     C.POINTER(_I64)]),
    "gapit_hapmap_numericalize": (C.c_int, [_P, _P, _I64, _P, C.c_int, C.c_int, _I64, _I64, C.c_int, C.c_int, C.c_int, _P,
                                            C.POINTER(_P)]),
    "gapit_vanraden": (C.c_int, [_P, _P, _P, _P]),
    "gapit_gram_ld": (_I64, [_I64]),
    "gapit_vanraden_gram": (C.c_int, [_P, _P, _P, _P]),
    "gapit_gram_plan": (C.c_int, [_I64, _I64, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gapit_vanraden_gram_fused": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "gapit_vanraden_gram_gather": (C.c_int, [_P, C.c_int64, _P, C.c_int, C.c_int]),
    "gapit_gram_owned_rows": (C.c_int, [_I64, C.c_int, C.c_int, C.POINTER(_I64), C.POINTER(_I64)]),
    "gapit_vanraden_finish": (C.c_int, [_P, _I64, _P, _P, _P, _P, _P]),
    "gapit_zhang": (C.c_int, [_P, _P, _P]),
    "gapit_pca": (C.c_int, [_P, _P, _I64, _P, _P]),
    "gapit_eigh": (C.c_int, [_P, _P, _I64, _P, _P]),
    "gapit_eigh_dev": (C.c_int, [_P, _P, _I64, _P]),
    "gapit_eigh_dev_range": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _P]),
    "gapit_eigh_dev_mg": (C.c_int, [_P, _P, _I64, _P, C.c_int]),
    "gapit_eigh_mg": (C.c_int, [_P, _P, _I64, _P, _P, C.c_int]),
    "gapit_eigen_R": (C.c_int, [_P, _P, _P, _I64, _I64, _P, _P]),
    "gapit_reml": (C.c_int, [_P, _P, _I64, C.c_int, C.c_double, C.c_double, C.c_double, C.POINTER(RemlOut)]),
    "gapit_reml_eigL": (C.c_int, [_P, _P, _P, _I64, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                  C.POINTER(RemlOut)]),
    "gapit_reml_eigL_multi": (C.c_int, [_P, _P, _P, _I64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                        C.POINTER(RemlOut), _P]),
    "gapit_p3d_scan": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, _P, C.c_int, _P, _P, C.c_int,
                                 C.POINTER(ScanOut), C.POINTER(ScanBase)]),
    "gapit_eigsys_upload": (C.c_int, [_P, _P, _P, _I64, C.POINTER(_P)]),
    "gapit_eigsys_from_device": (C.c_int, [_P, _P, _P, _I64, C.POINTER(_P)]),
    "gapit_eigsys_from_kinship": (C.c_int, [_P, _P, _I64, _P, C.POINTER(_P)]),
    "gapit_eigsys_project": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "gapit_eigsys_free": (None, [_P]),
    "gapit_p3d_scan_dev": (C.c_int, [_P, _P, _P, _P, C.c_int, _P, C.c_int, _P, _P, C.c_int,
                                     C.POINTER(ScanOut), C.POINTER(ScanBase)]),
    "gapit_p3d_scan_host": (C.c_int, [_P, _P, C.c_int, _I64, _I64, _I64, C.c_int, _P, _P, C.c_int, _P, C.c_int, _P, _P,
                                      C.c_int, C.POINTER(ScanOut), C.POINTER(ScanBase)]),
    "gapit_blup_pev": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_double, C.c_double, _P, _P, _P]),
    "gapit_set_slices": (C.c_int, [_P, C.c_int]),
    "gapit_get_slices": (C.c_int, [_P]),
    "gapit_eigsys_vectors_dev": (_P, [_P]),
    "gapit_bed_header": (C.c_int, [_P, _I64, _I64, C.POINTER(_I64), C.POINTER(_I64)]),
    "gapit_text_index_lines": (C.c_int, [_P, _I64, _P, _I64, C.POINTER(_I64)]),
    "gapit_numtext_shape": (C.c_int, [_P, _I64, _P, _I64, C.c_int, C.POINTER(_I64), C.POINTER(_I64)]),
    "gapit_numtext_numericalize": (C.c_int, [_P, _P, _I64, _P, _I64, _I64, _I64, C.c_int, _P, C.POINTER(_P)]),
    "gapit_mctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "gapit_mctx_destroy": (None, [_P]),
    "gapit_mctx_ngpus": (C.c_int, [_P]),
    "gapit_mctx_ctx": (_P, [_P, C.c_int]),
    "gapit_mgeno_pack": (C.c_int, [_P, _P, C.c_int, _I64, _I64, _I64, C.c_int, C.POINTER(_P)]),
    "gapit_mgeno_free": (None, [_P]),
    "gapit_mgeno_n": (_I64, [_P]),
    "gapit_mgeno_m": (_I64, [_P]),
    "gapit_multi_vanraden": (C.c_int, [_P, _P, _P, _P]),
    "gapit_multi_eigh": (C.c_int, [_P, _P, _I64, _P, _P]),
    "gapit_meigsys_from_kinship": (C.c_int, [_P, _P, _I64, _P, C.POINTER(_P)]),
    "gapit_meigsys_upload": (C.c_int, [_P, _P, _P, _I64, C.POINTER(_P)]),
    "gapit_meigsys_part": (_P, [_P, C.c_int]),
    "gapit_meigsys_free": (None, [_P]),
    "gapit_multi_p3d_scan_es": (C.c_int, [_P, _P, _P, _P, C.c_int, _P, C.c_int, _P, _P, C.c_int,
                                          C.POINTER(ScanOut), C.POINTER(ScanBase)]),
    "gapit_multi_p3d_scan": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, _P, C.c_int, _P, _P, C.c_int,
                                       C.POINTER(ScanOut), C.POINTER(ScanBase)]),
}

_lib = None


def load():
    """Load libgapitcuda.so (built in-tree by ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C gapit_b200/csrc).  gapit_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().gapit_last_error()
        raise GapitCudaError(rc, msg.decode() if msg else "")
