/* libgapitcuda.so -- C ABI of the B200-native GAPIT hot path.
 *
 * The reference (Zhiwu-Zhang-Lab/GAPIT, "2016.03.01") has no FFI layer: the
 * boundary of its hot path is two R closures.  Each entry point below names
 * the reference lines it replaces; INTEGRATION.md shows the `.Call` shim and
 * the replacement R bodies that bind them.
 *
 *   GAPIT.kinship.VanRaden(snps)            GAPIT.kinship.VanRaden.R:1-37   -> gapit_vanraden
 *   emma.eigen.L.wo.Z / emma.eigen.R.wo.Z   emma.txt:58-61, 82-89           -> gapit_eigh, gapit_eigen_R
 *   GAPIT.emma.REMLE (Z = NULL)             GAPIT.emma.REMLE.R:21-54,96-110 -> gapit_reml
 *   GAPIT.EMMAxP3D marker loop              GAPIT.EMMAxP3D.R:397-979        -> gapit_p3d_scan
 *   GAPIT.HapMap / GAPIT.Numericalization   GAPIT.HapMap.R:19-37            -> gapit_hapmap_numericalize
 *
 * Conventions: plain pointers and sizes only.  Host buffers are caller-owned,
 * read-only unless named *_out; device memory lives behind opaque handles.
 * Every call is synchronous on return, returns GAPIT_OK (0) or a negative
 * status, and never throws or longjmps; gapit_last_error() gives the message
 * of the last failure on the calling thread.  A context is not re-entrant.
 * There is no CPU fallback: without a CUDA device every compute call fails
 * with GAPIT_ERR_NO_DEVICE.
 */
#ifndef GAPITCUDA_H
#define GAPITCUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GAPIT_OK 0
#define GAPIT_ERR_ARG (-1)       /* bad argument (shape, NULL, unsupported q0 ...) */
#define GAPIT_ERR_CUDA (-2)      /* CUDA / cuSOLVER runtime failure */
#define GAPIT_ERR_NO_DEVICE (-3) /* no usable sm_100 device */
#define GAPIT_ERR_DATA (-4)      /* genotype value outside {0,1,2} (or NA without an impute rule) */
#define GAPIT_ERR_NUMERIC (-5)   /* eigen solver did not converge, REML bracket failure ... */

#define GAPIT_MAX_Q0 16 /* covariate columns incl. intercept the scan kernels are instantiated for */

typedef struct gapit_ctx gapit_ctx;
typedef struct gapit_geno gapit_geno;

/* element type of a host genotype matrix.  GAPIT_B2: 2 bits per genotype, marker-major: marker k starts at byte
 * k*ld (ld >= ceil(n/4) BYTES), taxon i in byte i/4 at bits 2(i%4)..2(i%4)+1 (the bit order of a PLINK .bed row);
 * code 0/1/2 = dosage, 3 = NA (imputed by the gapit_impute rule, rejected with GAPIT_IMPUTE_NONE). */
/* GAPIT_BED_A1 / GAPIT_BED_A2: the rows of a PLINK .bed file as they are (same bit order; codes 00 = homozygous A1,
 * 01 = missing, 10 = heterozygous, 11 = homozygous A2), dosage = copies of A1 / of A2; recoded on the device. */
typedef enum { GAPIT_F64 = 0, GAPIT_I32 = 1, GAPIT_I8 = 2, GAPIT_B2 = 3, GAPIT_BED_A1 = 4, GAPIT_BED_A2 = 5 } gapit_dtype;
/* NA rule for GAPIT_F64 input (GAPIT.EMMAxP3D.R:20-24, SNP.impute): NaN -> value; NONE rejects NaN */
typedef enum { GAPIT_IMPUTE_NONE = -1, GAPIT_IMPUTE_MINOR = 0, GAPIT_IMPUTE_MIDDLE = 1, GAPIT_IMPUTE_MAJOR = 2 } gapit_impute;

/* wall/device timings of the last call on a context, milliseconds (CUDA events) */
typedef struct {
  double h2d_ms;    /* host->device copies */
  double pack_ms;   /* dtype conversion, column statistics, transposition */
  double kernel_ms; /* the hot kernel(s): cross-product or rotation+reduction */
  double post_ms;   /* epilogue kernels after the hot kernel */
  double d2h_ms;    /* device->host copies */
  double total_ms;
} gapit_stats;

const char* gapit_last_error(void);
int gapit_device_count(void);

/* One context = one GPU (one process per GPU under torchrun / one R session). */
int gapit_ctx_create(int device, gapit_ctx** out);
void gapit_ctx_destroy(gapit_ctx* ctx);
int gapit_ctx_stats(const gapit_ctx* ctx, gapit_stats* out);
/* Run this context's work on a caller-owned CUDA stream (cudaStream_t passed as void*; NULL restores the
 * context's own stream), so a host framework can bracket calls with its own events. */
int gapit_ctx_set_stream(gapit_ctx* ctx, void* cuda_stream);

/* Genotypes: host matrix `snps`, n taxa x m markers, element (i,k) at
 * snps[i + k*ld] (R's column-major layout, ld >= n), values 0/1/2.  Stored on
 * the device as int8, marker-major, with per-marker sum and sum of squares.
 * Replaces the `as.matrix(thisGD)` / `xs` arguments of GAPIT.Genotype.R:431 and
 * GAPIT.Main.R:427. */
int gapit_geno_pack(gapit_ctx* ctx, const void* snps, gapit_dtype dtype, int64_t n, int64_t m, int64_t ld,
                    gapit_impute impute, gapit_geno** out);
/* Host-side packer for GAPIT_B2 (runs on the host's threads; nthreads 0 = all): an R-layout F64 / I32 / I8 matrix
 * (element (i,k) at snps[i + k*ld]) -> m rows of ceil(n/4) bytes in `out`.  NaN, NA_integer_ and negative int8 become
 * code 3 (NA); any other value outside {0,1,2} is GAPIT_ERR_DATA.  A caller holding R doubles then uploads 1/32 of the
 * bytes (gapit_geno_pack / gapit_p3d_scan_host with GAPIT_B2). */
int gapit_pack2_host(const void* snps, gapit_dtype dtype, int64_t n, int64_t m, int64_t ld, uint8_t* out, int nthreads);
/* The same handle from int8 dosages already resident on this context's device (marker k at mk_dev + k*ld, ld >= n):
 * one device-to-device re-pitch plus the column statistics; values outside {0,1,2} are rejected.  The copy runs on
 * the context's stream: either share that stream with the producer of mk_dev (gapit_ctx_set_stream) or
 * synchronise the producer first. */
int gapit_geno_from_device(gapit_ctx* ctx, const int8_t* mk_dev, int64_t n, int64_t m, int64_t ld, gapit_geno** out);
/* Zero-copy form of the same: mk_dev already has the library's layout (ld = n rounded up to 128, 128-byte aligned) and
 * stays the caller's (it must outlive the handle); the padding bytes n..ld of every row are zeroed by the call. */
int gapit_geno_wrap_device(gapit_ctx* ctx, int8_t* mk_dev, int64_t n, int64_t m, int64_t ld, gapit_geno** out);
/* A new handle holding taxa index[0..n_new) of `g` (0-based, repeats allowed), every marker: the `GD[GTindex,]` of the
 * reference's by-file loop (GAPIT.EMMAxP3D.R:315), which brings the taxa of a genotype file into phenotype order. */
int gapit_geno_select_taxa(gapit_ctx* ctx, const gapit_geno* g, const int64_t* index, int64_t n_new, gapit_geno** out);
void gapit_geno_free(gapit_geno* g);
int64_t gapit_geno_n(const gapit_geno* g);
int64_t gapit_geno_m(const gapit_geno* g);
/* per-marker allele-count sums (host out, length m) */
int gapit_geno_colsums(const gapit_geno* g, int32_t* colsum_out, int32_t* colsumsq_out);

/* ---- HapMap ingestion (GAPIT.HapMap.R:19-37 looping GAPIT.Numericalization.R:1-112 per marker) --------------
 * gapit_hapmap_index (host): scans the text of a HapMap table (tab separated, 11 leading columns, one row per
 * marker, optional heading row) and returns taxa count n, marker count m, genotype width `bit` (1 or 2 characters,
 * nchar(G[2,12]) at GAPIT.HapMap.R:28) and, when row_off_out != NULL (capacity row_cap), the byte offset of the first
 * genotype character of every data row; header_off_out (may be NULL) gets the offset of the first taxa name of the
 * heading row or -1.  Call once with row_off_out = NULL to size the array.
 * gapit_hapmap_numericalize (device): the text goes to the GPU as is; genotype i of marker k is the `bit` bytes at
 * text[row_off[k] + i*stride] (stride = bit + 1 for tab-separated text, bit for a dense code matrix).  Rules of
 * GAPIT.Numericalization: missing codes (:8-22), sorted levels and counts (:25-37), Major.allele.zero (:41-62),
 * dosage by level (:76-86), impute None/Middle/Minor/Major (:92-101), effect 0 Add / 1 Dom / 2 Left / 3 Right
 * (:105-107).  Outputs: gd_out (may be NULL) m x n int8 marker-major (= R's n x m column-major GD), NA = -1;
 * geno_out (may be NULL) the device-resident genotype handle the kinship and scan calls take (fails with
 * GAPIT_ERR_DATA if NA remain). */
int gapit_hapmap_index(const char* text, int64_t bytes, int heading, int64_t* n_out, int64_t* m_out, int* bit_out,
                       int64_t* row_off_out, int64_t row_cap, int64_t* header_off_out);
int gapit_hapmap_numericalize(gapit_ctx* ctx, const char* text, int64_t bytes, const int64_t* row_off, int bit, int stride,
                              int64_t n, int64_t m, int effect, gapit_impute impute, int major_allele_zero,
                              int8_t* gd_out, gapit_geno** geno_out);

/* ---- PLINK .bed (the reference itself only reads HapMap and numeric text) ------------
 * gapit_bed_header checks the magic bytes (6C 1B) and the variant-major mode byte (01), and returns the number of
 * variants m for n samples (.fam rows) and the offset of the first variant row (3).  The rows then go to
 * gapit_geno_pack / gapit_p3d_scan_host / gapit_mgeno_pack as GAPIT_BED_A1 or GAPIT_BED_A2 with ld = ceil(n/4): the
 * file bytes cross PCIe as they are (a memory-mapped .bed streams through gapit_p3d_scan_host in marker blocks, upload
 * overlapped with the scan) and the PLINK codes are mapped to dosages by the unpack kernel.  .bim / .fam are small
 * text tables parsed by the host language (r/GAPIT.Bed.R, gapit_b200.read_plink). */
int gapit_bed_header(const void* bytes, int64_t len, int64_t n, int64_t* m_out, int64_t* data_off_out);

/* ---- numeric text genotypes, genoFormat = "EMMA" (GAPIT.Fragment.R:94-136: `file.GD`, heading row, one row per TAXON:
 * name then one dosage per marker, blank / tab / comma separated; 0 1 2 (also 1.0 ...), NA NaN . for missing) ------
 * gapit_text_index_lines (host): offsets of the line starts, off_out[nlines] = bytes; off_out = NULL sizes the array.
 * gapit_numtext_shape (host): taxa n (non-blank rows after the heading) and markers m (tokens of the first data row - 1).
 * gapit_numtext_numericalize (device): the text goes to the GPU in row blocks (upload overlapped with parsing); every
 * row is tokenised and parsed by one CTA, then one byte transpose gives the marker-major operand layout.  gd_out (may
 * be NULL): n x m int8, R's column-major layout; geno_out (may be NULL): the handle the kinship and scan calls take.
 * A value that is not a dosage, a ragged row, or NA with GAPIT_IMPUTE_NONE is GAPIT_ERR_DATA. */
int gapit_text_index_lines(const char* text, int64_t bytes, int64_t* off_out, int64_t cap, int64_t* nlines_out);
int gapit_numtext_shape(const char* text, int64_t bytes, const int64_t* line_off, int64_t nlines, int heading,
                        int64_t* n_out, int64_t* m_out);
int gapit_numtext_numericalize(gapit_ctx* ctx, const char* text, int64_t bytes, const int64_t* line_off, int64_t first_row,
                               int64_t n, int64_t m, gapit_impute impute, int8_t* gd_out, gapit_geno** geno_out);

/* ---- VanRaden kinship (GAPIT.kinship.VanRaden.R:11-32) -------------------
 * K_out: n x n fp64 column-major (symmetric).  G_out (optional, may be NULL):
 * n x n int32 row-major raw cross-product sum_k x_ik x_jk over ALL markers of
 * `g` (the exact integer the epilogue starts from). */
int gapit_vanraden(gapit_ctx* ctx, gapit_geno* g, double* K_out, int32_t* G_out);

/* Sharded form (markers split across processes): each rank computes its
 * partial cross-product into a caller-owned DEVICE buffer G_dev
 * (gapit_gram_ld(n)^2 int32, zeroed by the call) and 3 int64 partial scalars
 * {sum c, sum c^2, #all-2 columns} into sums_dev (device); the caller sums both
 * across ranks (ncclAllReduce, integer => bit-identical for any rank count)
 * and any rank finishes.  Only the LOWER triangle of G_dev (by 256 x 256 tiles) is computed and read; the finish call
 * writes K to K_dev_out (device, may be NULL) and / or K_host_out, and mirrors G_dev into a full symmetric matrix only
 * when G_host_out is asked for. */
int64_t gapit_gram_ld(int64_t n);
int gapit_vanraden_gram(gapit_ctx* ctx, gapit_geno* g, int32_t* G_dev, int64_t* sums_dev);
int gapit_vanraden_finish(gapit_ctx* ctx, int64_t n, int32_t* G_dev, const int64_t* sums_dev, double* K_dev_out,
                          double* K_host_out, int32_t* G_host_out);
/* The same partial cross-product with the exchange FUSED into the kernel's epilogue over NVLink peer memory, instead of
 * a collective after it: every 256 x 256 tile leaves the kernel as bulk tensor reduce-adds (TMA, 32 x 32 int32 boxes)
 * into the Gram buffer of the rank that owns it.  peers[r] = rank r's Gram buffer (gapit_gram_ld(n)^2 int32 each, same
 * layout everywhere, peer-mapped into this process: cudaDeviceEnablePeerAccess pointers, torch symmetric memory or
 * cudaIpc handles).  Protocol: every rank zeroes its buffer, all ranks pass a barrier, every rank calls this, all ranks
 * synchronise their streams and pass a barrier again.  Then, with root >= 0, rank `root` holds the complete lower
 * triangle (reduce to root); with root < 0 rank r holds the complete block-rows gapit_gram_owned_rows(n, world, r)
 * (reduce-scatter: each rank's NVLink ingress carries 1/world of the tiles) and the rank that goes on calls
 * gapit_vanraden_gram_gather to pull the other ranks' block-rows into its own buffer.  gapit_vanraden_finish then runs
 * on a complete buffer.  Integer adds: bit-identical to the all-reduce route for any rank count.
 * sums_dev receives this rank's 3 partial scalars as in gapit_vanraden_gram.  sum_peers (may be NULL): sum_peers[r] = 3
 * int64 on rank r, peer-mapped like the Gram buffers and zeroed with them before the opening barrier (e.g. the tail of
 * the same allocation); every rank adds its scalars into all of them, so after the closing barrier each rank's own
 * mailbox holds the totals gapit_vanraden_finish wants -- no collective for them either.  With NULL the caller sums the
 * partial scalars over the ranks itself. */
int gapit_vanraden_gram_fused(gapit_ctx* ctx, gapit_geno* g, int32_t* const* peers, int world, int rank, int root,
                              int64_t* sums_dev, int64_t* const* sum_peers);
int gapit_vanraden_gram_gather(gapit_ctx* ctx, int64_t n, int32_t* const* peers, int world, int rank);
int gapit_gram_owned_rows(int64_t n, int world, int rank, int64_t* row_lo, int64_t* row_hi);
/* How the cross-product launch of n taxa x m markers is cut on a GPU with sm_count SMs (host arithmetic, no device
 * needed): pair jobs of the lower triangle, marker splits per job (chosen so that the equal-length jobs fill whole
 * rounds of the sm_count / 2 CTA pairs), rounds = ceil(jobs * splits / pairs).  Any out pointer may be NULL. */
int gapit_gram_plan(int64_t n, int64_t m, int sm_count, int* jobs_out, int* splits_out, int* rounds_out);

/* ---- ZHANG kinship and PCA covariates from the same integer cross-product ----------------------------
 * gapit_zhang: GAPIT.kinship.ZHANG(snps) (GAPIT.kinship.ZHANG.R:1-66): centred cross-product (:21-26), range
 * scaling to [0, 1 + inbreeding] (:43-47) and the two diagonal / off-diagonal adjustments (:50-61).  K_out n x n fp64.
 * gapit_pca: prcomp(X) of GAPIT.PCA.R:10 for the n x m dosage matrix: the first `ncomp` columns of PCA.X$x
 * (scores_out, n x ncomp column-major; sign of a component is arbitrary as in LAPACK) and all min(n, m) variances
 * PCA.X$sdev^2 (ev_out, may be NULL), from the eigen-system of the centred cross-product X_c X_c'. */
int gapit_zhang(gapit_ctx* ctx, gapit_geno* g, double* K_out);
int gapit_pca(gapit_ctx* ctx, gapit_geno* g, int64_t ncomp, double* scores_out, double* ev_out);

/* ---- spectral pieces (emma.txt:58-61, 82-89) ------------------------------
 * gapit_eigh: eigen(A, symmetric=TRUE): values descending, vectors n x n
 * column-major in matching order (cuSOLVER syevd; timed separately). */
int gapit_eigh(gapit_ctx* ctx, const double* A, int64_t n, double* values_out, double* vectors_out);
/* Device-resident form: A_dev (n x n column-major, e.g. K_dev_out of gapit_vanraden_finish) is overwritten by the
 * eigenvectors in descending order of the eigenvalues (values_out: host, n) -- no n^2 host round trip. */
int gapit_eigh_dev(gapit_ctx* ctx, double* A_dev, int64_t n, double* values_out);
/* Eigenpairs [first, first + count) of the descending order only (cusolverDnXsyevdx by index range): values_out host
 * (count), vectors_dev_out device n x count column-major; A_dev (lower triangle read) is destroyed.  A range call costs
 * the tridiagonalisation whatever the range (3.9 s of syevd's 4.5 s at n = 20,000, profiles/r2_syevd_breakdown.log), so
 * giving each GPU one slice of the spectrum splits the back-transformation and the output, not the whole solve. */
int gapit_eigh_dev_range(gapit_ctx* ctx, double* A_dev, int64_t n, int64_t first, int64_t count, double* values_out,
                         double* vectors_dev_out);
/* The same on `ngpus` GPUs of this process (0 = every visible one): cuSOLVER's multi-GPU solver cusolverMgSyevd on a
 * 1-D block-cyclic copy of A dealt out over NVLink, eigenvectors gathered back into A_dev.  Needs peer access; falls
 * through to gapit_eigh_dev for ngpus < 2 or n < 2048.  The solver sits in libgapitcuda_mg.so (loaded on first use);
 * GAPIT_ERR_NO_DEVICE when it cannot be loaded into this process (libcusolverMg needs the CUDA toolkit's libcublas: a
 * host process carrying another copy, e.g. a PyTorch wheel's, cannot use it).  gapit_eigh_mg is the host-buffer form
 * (eigen(K) of the R shim) and falls back to the one-GPU solver by itself. */
int gapit_eigh_dev_mg(gapit_ctx* ctx, double* A_dev, int64_t n, double* values_out, int ngpus);
int gapit_eigh_mg(gapit_ctx* ctx, const double* A, int64_t n, double* values_out, double* vectors_out, int ngpus);
/* gapit_eigen_R: eigen(S (K+I) S) with S = I - X (X'X)^-1 X'; returns the first
 * n-q values minus 1 and the first n-q vectors (n x (n-q) column-major). */
int gapit_eigen_R(gapit_ctx* ctx, const double* K, const double* X, int64_t n, int64_t q, double* values_out,
                  double* vectors_out);

/* ---- REML variance components (GAPIT.emma.REMLE.R:21-54, 96-110) ----------
 * lambda: n-q eigenvalues from gapit_eigen_R; etas = vectors' y.  Host code. */
typedef struct {
  double REML, delta, ve, vg;
} gapit_reml_out;
int gapit_reml(const double* lambda, const double* etas, int64_t nq, int ngrids, double llim, double ulim,
               double esp, gapit_reml_out* out);
/* Same optimisation (grid :27-33, uniroot :46-54, which.max :96-108) with the four sums of
 * emma.delta.REML.{LL,dLL}.wo.Z (emma.txt:145-164) evaluated from the spectrum of K alone: lambda = the n values of
 * gapit_eigh(K), Vty = vectors' y (n), VtX = vectors' X (n x q column-major).  Mathematically the same function of
 * delta as gapit_reml on eig.R, so the second O(n^3) eigendecomposition (emma.txt:82-89) can be skipped. */
int gapit_reml_eigL(const double* lambda, const double* Vty, const double* VtX, int64_t n, int q, int ngrids,
                    double llim, double ulim, double esp, gapit_reml_out* out);
/* T traits sharing X and the spectrum (VtY: n x T column-major; out: T entries; rc_out: per-trait status or NULL):
 * the independent searches run on the host's threads. */
int gapit_reml_eigL_multi(const double* lambda, const double* VtY, const double* VtX, int64_t n, int q, int T,
                          int ngrids, double llim, double ulim, double esp, gapit_reml_out* out, int* rc_out);

/* ---- P3D / EMMAx marker scan (GAPIT.EMMAxP3D.R:397-979) -------------------
 * Inputs: eigen-system of K (vectors n x n column-major, values), covariates
 * X0 (n x q0 column-major, intercept included by the caller as GAPIT.Main.R:319
 * does), phenotypes Y (n x T column-major) with per-trait delta and vg from
 * gapit_reml.  glm != 0 runs the K = NULL branch (:759-763, :865-870, :937,
 * :961-964): vectors/values/delta/vg are ignored.
 * Outputs are caller-allocated SoA arrays of m*T doubles (trait-major blocks of
 * m), NULL to skip one.  Rows of constant markers follow :499-511 (p = 1,
 * everything else NaN). */
typedef struct {
  double* effect;   /* beta of the marker                   effect.est  :938 */
  double* tvalue;   /* beta / sqrt(iXX_ss * vg|ve)          stats       :936-937 */
  double* se;       /* sqrt(iXX_ss * vg|ve)  (clean SE; the reference's `stderr` quirk :972 is rebuilt in the shim) */
  double* pvalue;   /* 2 pt(|t|, n-q0-1, lower=FALSE)       ps          :939-940 */
  double* rsquare;  /* 1 - exp(-(2/n)(logLM - logL0))       rsquare     :967 */
  double* loglm;    /* full-model log-likelihood            logLM       :958-963 */
  double* maf;      /* min(ss/2n, 1 - ss/2n), length m      maf         :486 */
} gapit_scan_out;

/* per-trait scalars of the marker-free model (i = 0 of the loop) */
typedef struct {
  double logL0;         /* :551-553 */
  double logLM_base;    /* :918-925 */
  double rsquare_base;  /* :926 */
  double ves;           /* GLM residual variance :868 (MLM: delta*vg passthrough) */
  double reml;          /* GLM: :870 ; MLM: unused (comes from gapit_reml) */
  double df;            /* n - q0 - 1 :586 */
  double beta0[GAPIT_MAX_Q0]; /* effect.cv :913 */
  int singular_x0;      /* 1 if X0'X0 needed the pseudo-inverse (:708-712) */
} gapit_scan_base;

int gapit_p3d_scan(gapit_ctx* ctx, gapit_geno* g, const double* vectors, const double* values, const double* X0,
                   int q0, const double* Y, int T, const double* delta, const double* vg, int glm,
                   gapit_scan_out* out, gapit_scan_base* base_out /* T entries */);

/* Device-resident form for repeated scans / benchmarking: upload the
 * eigen-system once (slices V for the int8 rotation), then scan. */
typedef struct gapit_eigsys gapit_eigsys;
int gapit_eigsys_upload(gapit_ctx* ctx, const double* vectors, const double* values, int64_t n, gapit_eigsys** out);
/* the same from eigenvectors already on the device (gapit_eigh_dev), and the projections V'R (R: n x nc column-major
 * host, out: n x nc host) that gapit_reml_eigL takes (V'y, V'X) */
int gapit_eigsys_from_device(gapit_ctx* ctx, const double* vectors_dev, const double* values, int64_t n, gapit_eigsys** out);
/* eigen(K) (host K, n x n) decomposed and sliced on the device in one call: the eigenvectors never visit the host.
 * values_out: host, n, descending. */
int gapit_eigsys_from_kinship(gapit_ctx* ctx, const double* K, int64_t n, double* values_out, gapit_eigsys** out);
int gapit_eigsys_project(gapit_ctx* ctx, gapit_eigsys* e, const double* R, int nc, double* out);
void gapit_eigsys_free(gapit_eigsys* e);
/* device address of the n x n column-major eigenvectors held by the handle (valid until gapit_eigsys_free) */
const double* gapit_eigsys_vectors_dev(const gapit_eigsys* e);
int gapit_p3d_scan_dev(gapit_ctx* ctx, gapit_geno* g, gapit_eigsys* e, const double* X0, int q0, const double* Y, int T,
                       const double* delta, const double* vg, int glm, gapit_scan_out* out,
                       gapit_scan_base* base_out);

/* Host-streaming form: the same scan straight from a host genotype matrix (layout as gapit_geno_pack).
 * Marker blocks are uploaded on a copy stream while the previous block is scanned, so the call costs
 * about max(H2D, scan).  e may be NULL when glm != 0. */
int gapit_p3d_scan_host(gapit_ctx* ctx, const void* snps, gapit_dtype dtype, int64_t n, int64_t m, int64_t ld,
                        gapit_impute impute, gapit_eigsys* e, const double* X0, int q0, const double* Y, int T,
                        const double* delta, const double* vg, int glm, gapit_scan_out* out,
                        gapit_scan_base* base_out);

/* ---- i = 0 BLUP / PEV block (GAPIT.EMMAxP3D.R:779-839), evaluated in the eigenbasis of K -------------
 * BLUP = K H^-1 (y - X0 beta0), PEV = diag(C22) of the mixed-model equations, beta0 = effect.cv; all length n
 * (beta_out: q0, may be NULL).  ve = delta * vg.  The reference's solve(K) -> MASS::ginv(K) fallback is decided
 * from the spectrum (|lambda|min / |lambda|max < DBL_EPSILON => pseudo-inverse with tolerance sqrt(eps)). */
int gapit_blup_pev(gapit_ctx* ctx, gapit_eigsys* e, const double* X0, int q0, const double* y, double delta, double vg,
                   double* blup_out, double* pev_out, double* beta_out);

/* ---- all GPUs of one box from ONE process -----------------------------------
 * What an R session binds: no launcher, no torch.distributed.  gapit_mctx_create(ngpus, ...) opens `ngpus` devices
 * (0 = every visible one, at most 8) and enables peer access between every pair; it fails with GAPIT_ERR_NO_DEVICE when
 * a pair has none.  gapit_mgeno_pack splits the host genotype matrix (layout and dtypes of gapit_geno_pack) into
 * contiguous marker blocks, one per GPU, uploaded concurrently on one host thread per GPU.  gapit_multi_vanraden:
 * every GPU's cross-product kernel reduce-scatters its tiles to the block-row owners over NVLink (bulk tensor
 * reduce-adds), GPU 0 gathers the rows and runs the exact epilogue; the result is bit-identical to gapit_vanraden for
 * any GPU count.  gapit_multi_p3d_scan: the eigen-system is uploaded once and copied GPU to GPU, every GPU scans its
 * block, rows land in the caller's arrays (m markers per trait) in marker order.  Arguments as in gapit_vanraden /
 * gapit_p3d_scan.  The worker threads never call back into the host runtime. */
typedef struct gapit_mctx gapit_mctx;
typedef struct gapit_mgeno gapit_mgeno;
int gapit_mctx_create(int ngpus, gapit_mctx** out);
void gapit_mctx_destroy(gapit_mctx* mc);
int gapit_mctx_ngpus(const gapit_mctx* mc);
gapit_ctx* gapit_mctx_ctx(gapit_mctx* mc, int gpu); /* the one-GPU context of device `gpu` (spectra, REML inputs ...) */
int gapit_mgeno_pack(gapit_mctx* mc, const void* snps, gapit_dtype dtype, int64_t n, int64_t m, int64_t ld,
                     gapit_impute impute, gapit_mgeno** out);
void gapit_mgeno_free(gapit_mgeno* g);
int64_t gapit_mgeno_n(const gapit_mgeno* g);
int64_t gapit_mgeno_m(const gapit_mgeno* g);
int gapit_multi_vanraden(gapit_mctx* mc, gapit_mgeno* g, double* K_out, int32_t* G_out);
/* eigen(A) as gapit_eigh (values descending, vectors n x n column-major, both host), the spectrum sliced over the
 * GPUs of the context: each one runs gapit_eigh_dev_range for n/ngpus eigenpairs.  One GPU (or n < 2048): gapit_eigh. */
int gapit_multi_eigh(gapit_mctx* mc, const double* A, int64_t n, double* values_out, double* vectors_out);
/* The eigen-system of K resident on EVERY GPU of the context (one sliced copy per GPU), so that eigen(K), the REML
 * projections, BLUP/PEV and the scan of one GAPIT.EMMAxP3D call share it without the eigenvectors crossing PCIe:
 * gapit_meigsys_from_kinship decomposes K with the spectrum sliced over the GPUs (as gapit_multi_eigh) and exchanges the
 * slices GPU to GPU; gapit_meigsys_upload starts from host eigenvectors instead (one upload, NVLink copies);
 * gapit_meigsys_part(me, gpu) is the one-GPU handle on device `gpu` (use it with gapit_mctx_ctx(mc, gpu) for
 * gapit_eigsys_project / gapit_blup_pev); gapit_multi_p3d_scan_es scans against it (me NULL with glm != 0). */
typedef struct gapit_meigsys gapit_meigsys;
int gapit_meigsys_from_kinship(gapit_mctx* mc, const double* K, int64_t n, double* values_out, gapit_meigsys** out);
int gapit_meigsys_upload(gapit_mctx* mc, const double* vectors, const double* values, int64_t n, gapit_meigsys** out);
gapit_eigsys* gapit_meigsys_part(gapit_meigsys* me, int gpu);
void gapit_meigsys_free(gapit_meigsys* me);
int gapit_multi_p3d_scan_es(gapit_mctx* mc, gapit_mgeno* g, gapit_meigsys* me, const double* X0, int q0, const double* Y,
                            int T, const double* delta, const double* vg, int glm, gapit_scan_out* out,
                            gapit_scan_base* base_out /* T entries */);
int gapit_multi_p3d_scan(gapit_mctx* mc, gapit_mgeno* g, const double* vectors, const double* values, const double* X0,
                         int q0, const double* Y, int T, const double* delta, const double* vg, int glm,
                         gapit_scan_out* out, gapit_scan_base* base_out /* T entries */);

/* Number of int8 digit planes each eigenvector is split into for the rotation (4..8; 0 = the default).  Each plane is 8
 * bits of the eigenvector relative to its largest element: S planes reproduce the fp64 product to ~2^-(8S) per element.
 * Default 5: measured against the fp64 oracle (bench.py `parity`, tests) beta holds 5e-10 relative, SE 2e-13, -log10 p
 * 1e-11 -- 20x inside the 1e-8 / 1e-6 tolerances -- and the scan is 21 % faster than with 6 planes (which reach the
 * oracle's own rounding, 3e-11). */
#define GAPIT_DEFAULT_SLICES 5
int gapit_set_slices(gapit_ctx* ctx, int slices);
int gapit_get_slices(const gapit_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* GAPITCUDA_H */
