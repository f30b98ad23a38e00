"""gapit_b200 -- B200-native engine for GAPIT's VanRaden kinship + EMMAx/P3D scan.

Everything numerical runs inside ``libgapitcuda.so`` (hand-written sm_100a CUDA,
C ABI in ``include/gapitcuda.h``); this package is the thin host mirror of the
reference's R functions for that path.  There is no CPU fallback.
"""
from .api import (Context, EigenSystem, GapitCudaError, Genotypes, Packed2, GAPIT_emma_REMLE, GAPIT_EMMAxP3D,  # noqa: F401
                  GAPIT_kinship_VanRaden, GAPIT_kinship_ZHANG, GAPIT_PCA, GAPIT_HapMap, GAPIT_Numericalization, blup_pev, default_context, eigh_dev, reml_from_projections, reml_from_projections_multi, emma_eigen_L_wo_Z, emma_eigen_R_wo_Z, p3d_scan, p3d_scan_host, p3d_scan_hapmap, alloc_scan_out,
                  MultiContext, MultiGenotypes, multi_vanraden, multi_eigh, MultiEigenSystem, multi_p3d_scan, eigh_dev_range, GAPIT_Compress, GAPIT_Block,
                  GAPIT_ZmatrixCompress, read_plink, GAPIT_numeric_text, p3d_scan_file)

__all__ = ["Context", "EigenSystem", "GapitCudaError", "Genotypes", "Packed2", "GAPIT_emma_REMLE", "GAPIT_EMMAxP3D",
           "GAPIT_kinship_VanRaden", "GAPIT_kinship_ZHANG", "GAPIT_PCA", "GAPIT_HapMap", "GAPIT_Numericalization", "blup_pev", "default_context", "eigh_dev", "reml_from_projections", "reml_from_projections_multi", "emma_eigen_L_wo_Z", "emma_eigen_R_wo_Z", "p3d_scan", "p3d_scan_host", "p3d_scan_hapmap", "alloc_scan_out",
           "MultiContext", "MultiGenotypes", "multi_vanraden", "multi_eigh", "MultiEigenSystem", "multi_p3d_scan", "eigh_dev_range", "GAPIT_Compress", "GAPIT_Block",
           "GAPIT_ZmatrixCompress", "read_plink", "GAPIT_numeric_text", "p3d_scan_file"]
