// EMMAx / P3D marker scan (GAPIT.EMMAxP3D.R:397-979), B200-native.
//
// The reference rotates one marker at a time by U = V diag((lambda+delta)^-1/2) (two n x n GEMVs per
// marker, :634 and :957).  Here the rotation is batched into a GEMM  W = V' X  and never written to HBM:
//
//   K3a slice_v          V (fp64) -> S int8 digit planes per eigenvector (balanced base-256 digits of
//                        V_jk / 2^e_k; "Ozaki split").  The genotype operand is exactly int8, so
//                        D_s[k,i] = sum_j digit_s(V_jk) x_ji are exact int32 and
//                        W_ki = c_k sum_s 256^s D_s[k,i] reproduces the fp64 product to ~2^-(8S-1)
//                        relative to max_j |V_jk|, deterministically (integer sums are order-free).
//   K3  rotate_reduce    tcgen05 kind::i8, M = markers (TMEM lanes), N = (k, slice) columns.  Each epilogue
//                        thread owns one marker: it rebuilds W_ki from the S int32 slices and accumulates
//                        xx = sum_k w_k W_ki^2,  xa_j = sum_k w_k A_kj W_ki,  xy = sum_k w_k b_k W_ki
//                        (w = 1/(lambda+delta), A = V'X0, b = V'y: the dot products of :653-655,:682)
//                        in registers while sweeping k, in a fixed order => bit-identical rows for
//                        identical genotype columns wherever they sit.
//   K3g glm_reduce       K = NULL branch (:759-763): the same UMMA main loop with the digit planes of [X0 | y]
//                        as B operand (U = I); streams the genotypes once through TMA, HBM-bound.
//   K4  finalize         per marker: Schur complement of [X0t xst]'[X0t xst] (:736-748), beta (:772),
//                        t (:936-937), p = 2 pt(|t|, df) (:939), logLM and R^2 (:955-967).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "umma_i8.cuh"

namespace gapit {

// geometry of the sliced eigenvector operand for S slices
struct SliceGeom {
  int S, BN, KPB;
};
static SliceGeom slice_geom(int S) {
  switch (S) {
    case 4: return {4, 256, 64};
    case 5: return {5, 240, 48};
    case 6: return {6, 240, 40};
    case 7: return {7, 224, 32};
    default: return {8, 256, 32};
  }
}

}  // namespace gapit

struct gapit_eigsys {
  gapit_ctx* ctx = nullptr;
  int64_t n = 0;
  int S = 0, BN = 0, KPB = 0;
  int64_t ldn = 0;    // bytes per row of Vs
  int64_t nkq = 0;    // number of BN-row tiles of Vs
  int8_t* Vs = nullptr;      // [(tiles_cap*BN)][ldn] row (k*S + s) = digit plane s of eigenvector k; tiles beyond nkq:
                             // digit planes of the back-rotated columns of a multi-trait scan (rewritten per call)
  int64_t tiles_cap = 0;
  double* V = nullptr;       // n x n column-major
  double* scale = nullptr;   // [nkq*KPB] c_k = max_j |V_jk| / (126 * 256^(S-1)), 0 beyond n
  std::vector<double> values;  // host, descending
  std::vector<double> h_scale;
  uint64_t serial = 0;         // unique per handle: part of the scan-plan cache key, so a plan built for a freed handle
                               // can never be matched again (the handle may outlive its context at interpreter exit)
  int device = 0;
};

namespace gapit {

// ------------------------------------------------------------------ K3a
// one block per eigenvector k: exponent from max |V_jk|, then S balanced base-256 digits per element
__global__ void slice_v_kernel(const double* __restrict__ V, int64_t n, int S, int8_t* __restrict__ Vs, int64_t ldn,
                               double* __restrict__ scale) {
  const int64_t k = blockIdx.x;
  const double* v = V + k * n;
  __shared__ double red[32];
  double mx = 0.0;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) mx = fmax(mx, fabs(v[j]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x < 32) {
    mx = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (threadIdx.x == 0) red[0] = mx;
  }
  __syncthreads();
  mx = red[0];
  // |q| <= 126 * 256^(S-1): the balanced digits below then stay within int8 (top digit <= 127).  The scale is not a
  // power of two: against 2^e >= max it keeps 1-2 more bits of every element (q is rounded to nearest either way).
  const double qmax = ldexp(126.0, 8 * (S - 1));
  const double up = (mx > 0.0) ? qmax / mx : 0.0;
  if (threadIdx.x == 0) scale[k] = (mx > 0.0) ? mx / qmax : 0.0;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
    long long q = __double2ll_rn(v[j] * up);
    for (int s = 0; s < S; ++s) {
      const long long d = ((q + 128) & 255) - 128;  // balanced digit in [-128, 127]
      Vs[(k * S + s) * ldn + j] = static_cast<int8_t>(d);
      q = (q - d) >> 8;
    }
  }
}

// ------------------------------------------------------------------ K3 pieces
struct ScanSched {
  struct Params {
    int mtiles;   // units of CL 256-marker tiles (a cluster of CL CTAs takes one unit: same eigen tiles, CL marker tiles)
    int cl;       // CTAs per cluster (1 or 2)
    int gsplit;   // splits of the eigen index (partial sums per split)
    int nkq;      // BN-row tiles of the sliced eigenvectors
    int kblocks;  // taxa axis in units of kBlockK
    int bn;
    uint64_t hint_a;  // L2 eviction hints (default: marker tile evict_last -- it is re-read for every eigen tile;
    uint64_t hint_b;  //  eigen tiles evict_normal -- streamed, shared by the CTAs running together)
    unsigned int* round_ctr;  // [rounds] zeroed before the launch, or nullptr: soft alignment of the units per round
    int round_wait;           // spin budget (clock cycles) before a unit gives up waiting for the others
  };
  __device__ uint64_t hint_a() const { return p.hint_a; }
  __device__ uint64_t hint_b() const { return p.hint_b; }
  // Every unit (CTA pair) starts round r = gid / stride (its r-th marker tile x eigen split) together with the others:
  // the ~15 pairs that stream the same window of digit planes then stay in step, so the window is fetched from DRAM
  // once and hit in L2 by the rest.  It is a soft barrier: arrival is counted, waiting is bounded by `round_wait`
  // cycles, so a unit that cannot be co-resident (shared GPU) only costs its peers that bounded wait, never a hang.
  __device__ void align_round() const {
    if (!p.round_ctr) return;
    const int total = p.mtiles * p.gsplit;
    const int r = gid / stride;
    const unsigned int expect = static_cast<unsigned int>(min(stride, total - r * stride));
    volatile unsigned int* c = p.round_ctr + r;
    atomicAdd(p.round_ctr + r, 1u);
    const long long t0 = clock64();
    while (*c < expect && clock64() - t0 < p.round_wait) __nanosleep(64);
  }
  Params p;
  int gid, stride, kq, kq_begin, kq_end, mt, ks, crank;
  __device__ ScanSched(const Params& pp, int unit, int nunits, uint32_t cta_rank)
      : p(pp), gid(unit - nunits), stride(nunits), kq(0), kq_begin(0), kq_end(0), mt(0), ks(0),
        crank(static_cast<int>(cta_rank)) {}
  __device__ bool next(Job& j) {
    if (kq >= kq_end) {
      gid += stride;
      if (gid >= p.mtiles * p.gsplit) return false;
      // split-minor, then marker tile: CTAs running together share the marker tile
      mt = gid / p.gsplit;
      ks = gid - mt * p.gsplit;
      kq_begin = (ks * p.nkq) / p.gsplit;
      kq_end = ((ks + 1) * p.nkq) / p.gsplit;
      kq = kq_begin;
    }
    j.a_row0 = (mt * p.cl + crank) * 256;
    j.b_row0 = kq * p.bn;
    j.kb_begin = 0;
    j.kb_end = p.kblocks;
    j.tag0 = kq;
    j.tag1 = ks;
    j.tag2 = 0;
    j.group_first = (kq == kq_begin);
    j.group_last = (kq == kq_end - 1);
    ++kq;
    return true;
  }
};

template <int S, int BN, int Q>
struct ScanEpi {
  static constexpr int KPB = BN / S;
  static constexpr int GC = (S == 4 || S == 8) ? 8 : (S == 5 ? 40 : (S == 6 ? 24 : 56));  // lcm(S, 8)
  static constexpr int KG = GC / S;
  static_assert(BN % GC == 0, "accumulator columns must split into whole digit groups");
  struct Params {
    const double* gconst;  // [nkq*KPB][Q+2]: w c^2 | w c A_j ... | w c b
    double* part;          // [gsplit*nhalf][Q+2][m_ld]
    int64_t m_ld;
    int64_t m;
    int pair;              // 1: merge adjacent digit planes in int32 before converting (needs 65792 n < 2^31)
  };
  // the constants of one job (KPB eigen indices) are staged in shared memory while the job's MMAs run
  static constexpr int kSmemDoubles = KPB * (Q + 2);
  static constexpr int kSmemBytes = kSmemDoubles * 8;
  Params p;
  double* cs;
  double acc[Q + 2];
  // exact int32 -> fp64: the bits 0x43300000:(x ^ 0x80000000) are the double 2^52 + 2^31 + x
  static __device__ __forceinline__ double i2d(int x) {
    return __hiloint2double(0x43300000, x ^ static_cast<int>(0x80000000u)) - 4503601774854144.0;
  }
  __device__ ScanEpi(const Params& pp, uint8_t* scratch) : p(pp), cs(reinterpret_cast<double*>(scratch)) {
#pragma unroll
    for (int i = 0; i < Q + 2; ++i) acc[i] = 0.0;
  }
  __device__ void finish() {}
  __device__ void prepare(const Job& job, int tid, int nthreads) {
    const double* src = p.gconst + int64_t(job.tag0) * kSmemDoubles;
    for (int i = tid; i < kSmemDoubles; i += nthreads) cs[i] = __ldg(src + i);
  }
  // `half` of `nhalf` epilogue warps on this (sub-tile, quarter): column groups are dealt round-robin and every
  // warp keeps its own partial sums (split index tag1 * nhalf + half in `part`)
  __device__ void run(const Job& job, uint32_t taddr, int sub, int row, int half, int nhalf) {
    if (job.group_first) {
#pragma unroll
      for (int i = 0; i < Q + 2; ++i) acc[i] = 0.0;
    }
#pragma unroll 1
    for (int g = half; g < BN / GC; g += nhalf) {
      uint32_t r[GC];
#pragma unroll
      for (int j = 0; j < GC / 8; ++j) ptx::tmem_ld_x8(taddr + g * GC + j * 8, *reinterpret_cast<uint32_t(*)[8]>(&r[j * 8]));
      ptx::tmem_ld_wait();
#pragma unroll
      for (int kk = 0; kk < KG; ++kk) {
        // u = sum_s 256^s D_s.  Adjacent digit planes are first merged in int32 (exact while
        // 257 * 256 * n < 2^31), halving the int -> fp64 conversions, which use the 2^52 magic
        // constant (one LOP + one DADD) instead of the quarter-rate I2F.F64.
        const int* d = reinterpret_cast<const int*>(&r[kk * S]);
        double u;
        if (p.pair) {
          if (S & 1) {
            u = i2d(d[S - 1]);
#pragma unroll
            for (int s = S - 3; s >= 0; s -= 2) u = fma(u, 65536.0, i2d(d[s + 1] * 256 + d[s]));
          } else {
            u = i2d(d[S - 1] * 256 + d[S - 2]);
#pragma unroll
            for (int s = S - 4; s >= 0; s -= 2) u = fma(u, 65536.0, i2d(d[s + 1] * 256 + d[s]));
          }
        } else {
          u = i2d(d[S - 1]);
#pragma unroll
          for (int s = S - 2; s >= 0; --s) u = fma(u, 256.0, i2d(d[s]));
        }
        const double* c = cs + (g * KG + kk) * (Q + 2);   // shared-memory broadcast reads
        acc[0] = fma(c[0] * u, u, acc[0]);
#pragma unroll
        for (int i = 1; i < Q + 2; ++i) acc[i] = fma(c[i], u, acc[i]);
      }
    }
    if (job.group_last) {
      const int64_t mrk = int64_t(job.a_row0) + sub * 128 + row;
      if (mrk < p.m_ld) {
        double* dst = p.part + int64_t(job.tag1 * nhalf + half) * (Q + 2) * p.m_ld + mrk;
#pragma unroll
        for (int i = 0; i < Q + 2; ++i) dst[i * p.m_ld] = acc[i];
      }
    }
  }
};

// ------------------------------------------------------------------ K3m
// Many traits sharing one eigendecomposition (BASELINE configs[4]; GAPIT.R:101-132 calls GAPIT.EMMAxP3D per trait and
// every call repeats the rotation :634).  W = V'X does not depend on the trait; only w_t = 1/(lambda + delta_t) does:
//   xx_t(i)   = sum_k w_tk W_ki^2                      needs the rotation (quadratic in x_i)
//   xa_tj(i)  = x_i' H_t^-1 X0_j,  num_t(i) = x_i' H_t^-1 (y_t - X0 beta0_t)     are LINEAR in x_i
// (H_t = K + delta_t I = V diag(1/w_t) V').  So the B operand of ONE rotation launch is the digit planes of V followed
// by the digit planes of the back-rotated columns R = [H_t^-1 X0 | H_t^-1 r_t]_t, and this epilogue only rebuilds the
// fp64 value of every accumulator column and stores it: W_ki^2 for the eigenvector rows (row k of `wbuf`), the finished
// linear sums for the R rows.  A DMMA GEMM (xx_contract, contract.cu) then forms xx_t for all traits from W^2 -- the
// rotation runs once for any number of traits and no trait-sized accumulator block has to fit in registers.
template <int S, int BN>
struct StoreEpi {
  static constexpr int KPB = BN / S;
  static constexpr int GC = (S == 4 || S == 8) ? 8 : (S == 5 ? 40 : (S == 6 ? 24 : 56));  // lcm(S, 8)
  static constexpr int KG = GC / S;
  struct Params {
    const double* kscale;  // [rows] c_k of every B row group (eigenvectors, then R columns)
    double* wbuf;          // [rows][m_ld]
    int64_t m_ld;
    int nk_sq;             // rows below this index are squared (eigenvector rows)
    int pair;              // 1: merge adjacent digit planes in int32 first
  };
  static constexpr int kSmemBytes = 0;
  Params p;
  static __device__ __forceinline__ double i2d(int x) {
    return __hiloint2double(0x43300000, x ^ static_cast<int>(0x80000000u)) - 4503601774854144.0;
  }
  __device__ StoreEpi(const Params& pp, uint8_t*) : p(pp) {}
  __device__ void prepare(const Job&, int, int) {}
  __device__ void finish() {}
  __device__ void run(const Job& job, uint32_t taddr, int sub, int row, int half, int nhalf) {
    const int64_t mrk = int64_t(job.a_row0) + sub * 128 + row;
    const bool live = mrk < p.m_ld;
#pragma unroll 1
    for (int g = half; g < BN / GC; g += nhalf) {
      uint32_t r[GC];
#pragma unroll
      for (int j = 0; j < GC / 8; ++j) ptx::tmem_ld_x8(taddr + g * GC + j * 8, *reinterpret_cast<uint32_t(*)[8]>(&r[j * 8]));
      ptx::tmem_ld_wait();
#pragma unroll
      for (int kk = 0; kk < KG; ++kk) {
        const int* d = reinterpret_cast<const int*>(&r[kk * S]);
        double u;
        if (p.pair) {
          if (S & 1) {
            u = i2d(d[S - 1]);
#pragma unroll
            for (int s = S - 3; s >= 0; s -= 2) u = fma(u, 65536.0, i2d(d[s + 1] * 256 + d[s]));
          } else {
            u = i2d(d[S - 1] * 256 + d[S - 2]);
#pragma unroll
            for (int s = S - 4; s >= 0; s -= 2) u = fma(u, 65536.0, i2d(d[s + 1] * 256 + d[s]));
          }
        } else {
          u = i2d(d[S - 1]);
#pragma unroll
          for (int s = S - 2; s >= 0; --s) u = fma(u, 256.0, i2d(d[s]));
        }
        const int k = job.tag0 * KPB + g * KG + kk;
        double v = u * __ldg(p.kscale + k);
        if (k < p.nk_sq) v *= v;
        if (live) p.wbuf[int64_t(k) * p.m_ld + mrk] = v;
      }
    }
  }
};

// ------------------------------------------------------------------ K3g
// GLM (K = NULL, U = I): xa_j = sum_i X0_ij x_i and xy = sum_i y_i x_i are the product of the marker tile
// with R = [X0 | y] (n x (q0+1)), so the same int8 UMMA main loop is used with the digit planes of R as the
// B operand (one BN-row tile, N = (q0+1) S columns).  The kernel then streams the genotypes exactly once
// through TMA at HBM rate; xx = sum_i x_i^2 is the exact integer the pack stage already holds.
struct GlmSched {
  struct Params {
    int mtiles;   // 256-marker tiles
    int kblocks;  // taxa axis in units of kBlockK
  };
  __device__ uint64_t hint_a() const { return ptx::kEvictFirst; }  // genotypes: read once
  __device__ uint64_t hint_b() const { return ptx::kEvictLast; }   // digit planes of R: re-read by every tile
  __device__ void align_round() const {}
  Params p;
  int tile, stride;
  __device__ GlmSched(const Params& pp, int unit, int nunits, uint32_t) : p(pp), tile(unit), stride(nunits) {}
  __device__ bool next(Job& j) {
    if (tile >= p.mtiles) return false;
    j.a_row0 = tile * 256;
    j.b_row0 = 0;
    j.kb_begin = 0;
    j.kb_end = p.kblocks;
    j.tag0 = j.tag1 = j.tag2 = 0;
    j.group_first = j.group_last = true;
    tile += stride;
    return true;
  }
};

// exact int32 -> fp64: the bits 0x43300000:(x ^ 0x80000000) are the double 2^52 + 2^31 + x
static __device__ __forceinline__ double int_to_double(int x) {
  return __hiloint2double(0x43300000, x ^ static_cast<int>(0x80000000u)) - 4503601774854144.0;
}

template <int S, int BN>
struct GlmEpi {
  static constexpr int NCMAX = BN / S;
  struct Params {
    const double* scale;       // [nc] c_j = max_i |R_ij| / (126 * 256^(S-1)) of column j of R
    const int32_t* colsumsq;   // [m]
    double* part;              // [qpad+2][m_ld]
    int64_t m_ld;
    int64_t m;
    int nc;                    // q0 + 1 columns of R in use
    int qpad;
  };
  static constexpr int kSmemBytes = 0;
  Params p;
  __device__ GlmEpi(const Params& pp, uint8_t*) : p(pp) {}
  __device__ void prepare(const Job&, int, int) {}
  __device__ void finish() {}
  __device__ void run(const Job& job, uint32_t taddr, int sub, int row, int /*half*/, int /*nhalf*/) {
    uint32_t r[BN];
#pragma unroll
    for (int j = 0; j < BN / 8; ++j) ptx::tmem_ld_x8(taddr + j * 8, *reinterpret_cast<uint32_t(*)[8]>(&r[j * 8]));
    ptx::tmem_ld_wait();
    const int64_t mrk = int64_t(job.a_row0) + sub * 128 + row;
    if (mrk >= p.m_ld) return;
    p.part[mrk] = (mrk < p.m) ? static_cast<double>(p.colsumsq[mrk]) : 0.0;
#pragma unroll
    for (int c = 0; c < NCMAX; ++c) {
      if (c < p.nc) {
        const int* d = reinterpret_cast<const int*>(&r[c * S]);
        double u = int_to_double(d[S - 1]);
#pragma unroll
        for (int s = S - 2; s >= 0; --s) u = fma(u, 256.0, int_to_double(d[s]));
        // columns 0..q0-1 are covariates (slots 1..q0), column q0 is the phenotype (slot qpad+1)
        const int slot = (c == p.nc - 1) ? (p.qpad + 1) : (1 + c);
        p.part[int64_t(slot) * p.m_ld + mrk] = u * __ldg(p.scale + c);
      }
    }
  }
};

// ------------------------------------------------------------------ K4
struct FinalizeConsts {
  double iXX[GAPIT_MAX_Q0 * GAPIT_MAX_Q0];  // (X0t'X0t)^-1, row-major q0 x q0
  double beta0[GAPIT_MAX_Q0];
  double var_scale;   // vg (MLM) or ves (GLM)
  double rss0;        // reduced-model weighted RSS
  double sumlog;      // sum log(lambda + delta)   (0 for GLM)
  double logL0;
  double n;
  double df;
  double lbeta;       // log Beta(df/2, 1/2)
  int q0;
  int qpad;    // covariate width the reduction kernels were instantiated for (>= q0)
};

// continued fraction of the incomplete beta function (modified Lentz)
__device__ double betacf(double a, double b, double x) {
  const double FPMIN = 1e-300, EPS = 2.5e-16;
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < FPMIN) d = FPMIN;
  d = 1.0 / d;
  double h = d;
  for (int mm = 1; mm <= 2000; ++mm) {
    const double m2 = 2.0 * mm;
    double aa = mm * (b - mm) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < FPMIN) d = FPMIN;
    c = 1.0 + aa / c;
    if (fabs(c) < FPMIN) c = FPMIN;
    d = 1.0 / d;
    h *= d * c;
    aa = -(a + mm) * (qab + mm) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < FPMIN) d = FPMIN;
    c = 1.0 + aa / c;
    if (fabs(c) < FPMIN) c = FPMIN;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < EPS) break;
  }
  return h;
}

// 1 / (1.5 + k): the term ratios of the series below without a division
__device__ const double kInvHalfPlus[64] = {0.6666666666666666, 0.4, 0.2857142857142857, 0.2222222222222222, 0.18181818181818182, 0.15384615384615385, 0.13333333333333333, 0.11764705882352941, 0.10526315789473684, 0.09523809523809523, 0.08695652173913043, 0.08, 0.07407407407407407, 0.06896551724137931, 0.06451612903225806, 0.06060606060606061, 0.05714285714285714, 0.05405405405405406, 0.05128205128205128, 0.04878048780487805, 0.046511627906976744, 0.044444444444444446, 0.0425531914893617, 0.04081632653061224, 0.0392156862745098, 0.03773584905660377, 0.03636363636363636, 0.03508771929824561, 0.03389830508474576, 0.03278688524590164, 0.031746031746031744, 0.03076923076923077, 0.029850746268656716, 0.028985507246376812, 0.028169014084507043, 0.0273972602739726, 0.02666666666666667, 0.025974025974025976, 0.02531645569620253, 0.024691358024691357, 0.024096385542168676, 0.023529411764705882, 0.022988505747126436, 0.02247191011235955, 0.02197802197802198, 0.021505376344086023, 0.021052631578947368, 0.020618556701030927, 0.020202020202020204, 0.019801980198019802, 0.019417475728155338, 0.01904761904761905, 0.018691588785046728, 0.01834862385321101, 0.018018018018018018, 0.017699115044247787, 0.017391304347826087, 0.017094017094017096, 0.01680672268907563, 0.01652892561983471, 0.016260162601626018, 0.016, 0.015748031496062992, 0.015503875968992248};

// 2 * pt(|t|, df, lower.tail = FALSE) = I_{df/(df+t^2)}(df/2, 1/2)
__device__ double two_sided_t_pvalue(double t, double df, double lbeta) {
  if (t != t) return t;
  const double t2 = t * t;
  if (isinf(t2)) return 0.0;
  const double a = 0.5 * df, b = 0.5;
  const double x = df / (df + t2);
  const double y = t2 / (df + t2);
  if (y <= 0.0) return 1.0;
  const double ln_x = -log1p(t2 / df);
  const double ln_y = log(y);
  // Nearly every marker (t^2 below ~20): the lower part 1 - p = I_y(1/2, df/2) = y^(1/2) (1-y)^(df/2) / (1/2 B) * S with
  // the hypergeometric series S = sum_k [(1/2 + df/2)_k / (3/2)_k] y^k, whose terms behave like (t^2/2)^k / (3/2)_k:
  // 20-60 multiply-adds, against ~sqrt(df) continued-fraction steps with two divisions each.  This test comes FIRST: one
  // lane on the continued fraction holds its whole warp there, and with the customary switch at t^2 ~ 3 that is 94 % of
  // the warps.  1 - (.) costs 1e-16 / p of relative accuracy -- 1e-11 at the switch-over (p ~ 1e-5), far inside the
  // 1e-6 tolerance on log10 p (log B is formed in extended precision for the same reason).
  if (a * y < 10.0) {
    const double ay = (a + b) * y;
    double term = 1.0, S = 1.0;
    int k = 0;
    for (; k < 64; ++k) {
      term *= fma(static_cast<double>(k), y, ay) * kInvHalfPlus[k];
      S += term;
      if (term < 1e-17 * S) break;
    }
    if (k < 64) return 1.0 - exp(b * ln_y + a * ln_x - lbeta - log(b)) * S;
  }
  if (x < (a + 1.0) / (a + b + 2.0)) return exp(a * ln_x + b * ln_y - lbeta - log(a)) * betacf(a, b, x);
  return 1.0 - exp(b * ln_y + a * ln_x - lbeta - log(b)) * betacf(b, a, y);
}

// where the per-marker sums of trait t live: xx[t*xx_t + g*xx_g + k], xa_j[t*xa_t + g*xa_g + j*xa_j + k],
// xy[t*xy_t + g*xy_g + k], g = 0..nsplit-1 partial sums added in index order
struct FinalizeSrc {
  const double* xx;
  const double* xa;
  const double* xy;
  int64_t xx_t, xx_g, xa_t, xa_g, xa_j, xy_t, xy_g;
  int nsplit;
};

// grid (marker blocks, traits); outputs of trait t at out + t*out_t + a*out_ld + k, a = 0..6
__global__ void finalize_kernel(FinalizeSrc src, int64_t m, const int32_t* __restrict__ colsum,
                                const int32_t* __restrict__ colsumsq, const FinalizeConsts* __restrict__ fcs,
                                double* __restrict__ out, int64_t out_t, int64_t out_ld) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= m) return;
  const int t = blockIdx.y;
  const FinalizeConsts& fc = fcs[t];
  double* effect = out + t * out_t + k;
  double* tvalue = effect + out_ld;
  double* se = effect + 2 * out_ld;
  double* pvalue = effect + 3 * out_ld;
  double* rsquare = effect + 4 * out_ld;
  double* loglm = effect + 5 * out_ld;
  const int q0 = fc.q0;
  const double NaN = __longlong_as_double(0x7ff8000000000000LL);
  const long long ss = colsum[k], sq = colsumsq[k];
  const double half_freq = 0.5 * static_cast<double>(ss) / fc.n;
  if (t == 0) effect[6 * out_ld] = fmin(half_freq, 1.0 - half_freq);
  // constant marker (min == max  <=>  n sum x^2 == (sum x)^2): GAPIT.EMMAxP3D.R:499-511
  if (static_cast<long long>(fc.n) * sq == ss * ss) {
    *effect = NaN; *tvalue = NaN; *se = NaN; *pvalue = 1.0; *rsquare = NaN; *loglm = NaN;
    return;
  }
  double xx = 0.0, xy = 0.0, xa[GAPIT_MAX_Q0];
  for (int j = 0; j < q0; ++j) xa[j] = 0.0;
  for (int g = 0; g < src.nsplit; ++g) {  // fixed order: deterministic for any grid
    xx += src.xx[t * src.xx_t + g * src.xx_g + k];
    for (int j = 0; j < q0; ++j) xa[j] += src.xa[t * src.xa_t + g * src.xa_g + j * src.xa_j + k];
    xy += src.xy[t * src.xy_t + g * src.xy_g + k];
  }
  // Schur complement of the bordered normal equations (:736-748)
  double quad = 0.0, num = xy;
  for (int i = 0; i < q0; ++i) {
    double ti = 0.0;
    for (int j = 0; j < q0; ++j) ti = fma(fc.iXX[i * q0 + j], xa[j], ti);
    quad = fma(xa[i], ti, quad);
    num = fma(-xa[i], fc.beta0[i], num);
  }
  const double B22 = xx - quad;
  const double beta = num / B22;                      // :772 last element
  const double sdev = sqrt(fc.var_scale / B22);       // sqrt(iXX[q1,q1] * vg|ve)
  const double tv = beta / sdev;                      // :936-937
  double p = two_sided_t_pvalue(tv, fc.df, fc.lbeta); // :939
  if (p != p) p = 1.0;                                // :940
  const double rss = fc.rss0 - num * beta;            // ||U'(y - X beta)||^2 (:957) via the Schur update
  const double ll = 0.5 * (-fc.n * log((2.0 * 3.14159265358979323846 / fc.n) * rss) - fc.sumlog - fc.n);  // :958
  *effect = beta;
  *tvalue = tv;
  *se = sdev;
  *pvalue = p;
  *loglm = ll;
  *rsquare = 1.0 - exp(-(2.0 / fc.n) * (ll - fc.logL0));  // :967
}

}  // namespace gapit

// =============================================================================== host side
namespace gapit {

// inverse of a small symmetric matrix; falls back to the Moore-Penrose inverse (MASS::ginv's role at
// GAPIT.EMMAxP3D.R:708-712) when it is numerically singular.  A: q x q row-major.
static void sym_inverse(int q, const double* A, double* inv, int* singular) {
  std::vector<double> L(q * q, 0.0);
  bool ok = true;
  double dmax = 0.0, dmin = INFINITY;
  for (int i = 0; i < q && ok; ++i) {
    for (int j = 0; j <= i; ++j) {
      double s = A[i * q + j];
      for (int k = 0; k < j; ++k) s -= L[i * q + k] * L[j * q + k];
      if (i == j) {
        if (!(s > 0.0)) { ok = false; break; }
        L[i * q + i] = std::sqrt(s);
        dmax = std::max(dmax, L[i * q + i]);
        dmin = std::min(dmin, L[i * q + i]);
      } else {
        L[i * q + j] = s / L[j * q + j];
      }
    }
  }
  if (ok && dmin / dmax < 1e-8) ok = false;  // cond(A) ~ (dmax/dmin)^2 beyond what solve() accepts
  if (ok) {
    // inv = L^-T L^-1
    std::vector<double> Li(q * q, 0.0);
    for (int i = 0; i < q; ++i) {
      Li[i * q + i] = 1.0 / L[i * q + i];
      for (int j = 0; j < i; ++j) {
        double s = 0.0;
        for (int k = j; k < i; ++k) s -= L[i * q + k] * Li[k * q + j];
        Li[i * q + j] = s / L[i * q + i];
      }
    }
    for (int i = 0; i < q; ++i)
      for (int j = 0; j < q; ++j) {
        double s = 0.0;
        for (int k = std::max(i, j); k < q; ++k) s += Li[k * q + i] * Li[k * q + j];
        inv[i * q + j] = s;
      }
    *singular = 0;
    return;
  }
  // Jacobi eigen-decomposition, then pseudo-inverse with MASS::ginv's tolerance sqrt(eps) * max
  *singular = 1;
  std::vector<double> M(A, A + q * q), E(q * q, 0.0);
  for (int i = 0; i < q; ++i) E[i * q + i] = 1.0;
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < q; ++i)
      for (int j = i + 1; j < q; ++j) off += M[i * q + j] * M[i * q + j];
    if (off < 1e-300) break;
    for (int p = 0; p < q; ++p)
      for (int r = p + 1; r < q; ++r) {
        if (std::fabs(M[p * q + r]) < 1e-300) continue;
        const double theta = (M[r * q + r] - M[p * q + p]) / (2.0 * M[p * q + r]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < q; ++k) {
          const double a = M[k * q + p], b = M[k * q + r];
          M[k * q + p] = c * a - s * b;
          M[k * q + r] = s * a + c * b;
        }
        for (int k = 0; k < q; ++k) {
          const double a = M[p * q + k], b = M[r * q + k];
          M[p * q + k] = c * a - s * b;
          M[r * q + k] = s * a + c * b;
        }
        for (int k = 0; k < q; ++k) {
          const double a = E[k * q + p], b = E[k * q + r];
          E[k * q + p] = c * a - s * b;
          E[k * q + r] = s * a + c * b;
        }
      }
  }
  double emax = 0.0;
  for (int i = 0; i < q; ++i) emax = std::max(emax, std::fabs(M[i * q + i]));
  const double tol = std::sqrt(2.220446049250313e-16) * emax;
  for (int i = 0; i < q; ++i)
    for (int j = 0; j < q; ++j) {
      double s = 0.0;
      for (int k = 0; k < q; ++k)
        if (M[k * q + k] > tol) s += E[i * q + k] * E[j * q + k] / M[k * q + k];
      inv[i * q + j] = s;
    }
}

static int pad_q(int q0) {
  if (q0 <= 1) return 1;
  if (q0 <= 2) return 2;
  if (q0 <= 4) return 4;
  if (q0 <= 8) return 8;
  return 16;
}

// The rotation runs as the cta_group::2 pair kernel whenever the block holds at least two 256-marker tiles, else
// (a single tile) as the single-CTA form of the same main loop.  The single-trait epilogue runs two warps per
// (sub-tile, lane quarter), each with its own partial sums: kScanHalves partial sums per eigen split whatever form
// runs, so the fp64 summation order -- hence every output bit -- does not depend on the tile count.
constexpr int kScanHalves = 2;

// digit planes may be merged pairwise in int32 while 257 * 256 * n < 2^31 (n <= 32,639); GAPIT_EPI_PAIR=0 forces the
// unmerged epilogue of larger n on a small problem (test hook)
static int epilogue_merges_planes(int64_t n) {
  int pair = (65792ll * n < 2147483647ll) ? 1 : 0;
  if (const char* ev = getenv("GAPIT_EPI_PAIR")) pair = pair && atoi(ev);
  return pair;
}

// single-trait rotate + reduce: gconst_dev [nk][Q+2], part_dev [gsplit * kScanHalves][Q+2][m_ld]
template <int S, int BN, int Q>
static int launch_scan_SQ(gapit_ctx* ctx, gapit_geno* g, gapit_eigsys* e, const double* gconst_dev, double* part_dev,
                          int64_t m_ld, int gsplit) {
  using Epi = ScanEpi<S, BN, Q>;
  CUtensorMap map_a, map_b;
  GAPIT_CHECK(make_tensor_map_u8(&map_a, g->mk, static_cast<uint64_t>(g->ldn), static_cast<uint64_t>(g->m),
                                 static_cast<uint64_t>(g->ldn), 256));
  const int cl = (m_ld / 256 >= 2) ? 2 : 1;
  GAPIT_CHECK(make_tensor_map_u8(&map_b, e->Vs, static_cast<uint64_t>(e->ldn), static_cast<uint64_t>(e->nkq * BN),
                                 static_cast<uint64_t>(e->ldn), BN / cl));
  ScanSched::Params sp;
  sp.mtiles = static_cast<int>((m_ld / 256 + cl - 1) / cl);
  sp.cl = cl;
  sp.gsplit = gsplit;
  sp.nkq = static_cast<int>(e->nkq);
  sp.kblocks = static_cast<int>(g->ldn / kBlockK);
  sp.bn = BN;
  sp.hint_a = ptx::kEvictLast;     // marker tile: re-read for every eigen tile
  sp.hint_b = ptx::kEvictNormal;   // eigen tiles: streamed, shared by the CTAs running together
  sp.round_ctr = nullptr;
  sp.round_wait = 100000;   // ~50 us
  typename Epi::Params ep{gconst_dev, part_dev, m_ld, g->m, epilogue_merges_planes(g->n)};
  const int nunits = std::min(ctx->sm_count / cl, sp.mtiles * sp.gsplit);
  constexpr int EPIW = kScanHalves;
  constexpr int kThreadsEpi = 128 + 2 * 128 * EPIW;
  if (cl == 2) {
    const size_t rounds = size_t((sp.mtiles * sp.gsplit + nunits - 1) / nunits) + 1;
    GAPIT_CHECK(ctx_ws_t(ctx, WS_SYNC, rounds, &sp.round_ctr));
    GAPIT_CUDA(cudaMemsetAsync(sp.round_ctr, 0, rounds * sizeof(unsigned int), ctx->stream));
    using PCfg = UmmaPairCfg<2, BN, 4>;
    auto kern = umma_i8_pair_kernel<2, BN, 4, EPIW, ScanSched, Epi, 1, 1>;
    GAPIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PCfg::kSmemBytes + Epi::kSmemBytes + 16));
    GAPIT_CUDA(launch_umma(kern, nunits, 2, kThreadsEpi, PCfg::kSmemBytes + Epi::kSmemBytes + 16, ctx->stream, map_a, map_b, sp, ep));
  } else {
    using Cfg = UmmaCfg<2, BN, 3>;
    auto kern = umma_i8_kernel<2, BN, 3, EPIW, ScanSched, Epi, 1, 1>;
    GAPIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes + Epi::kSmemBytes + 16));
    GAPIT_CUDA(launch_umma(kern, nunits, 1, kThreadsEpi, Cfg::kSmemBytes + Epi::kSmemBytes + 16, ctx->stream, map_a, map_b, sp, ep));
  }
  return GAPIT_OK;
}

template <int S, int BN>
static int launch_scan_S(gapit_ctx* ctx, gapit_geno* g, gapit_eigsys* e, int Q, const double* gconst_dev,
                         double* part_dev, int64_t m_ld, int gsplit) {
  switch (Q) {
    case 1: return launch_scan_SQ<S, BN, 1>(ctx, g, e, gconst_dev, part_dev, m_ld, gsplit);
    case 2: return launch_scan_SQ<S, BN, 2>(ctx, g, e, gconst_dev, part_dev, m_ld, gsplit);
    case 4: return launch_scan_SQ<S, BN, 4>(ctx, g, e, gconst_dev, part_dev, m_ld, gsplit);
    case 8: return launch_scan_SQ<S, BN, 8>(ctx, g, e, gconst_dev, part_dev, m_ld, gsplit);
    default: return launch_scan_SQ<S, BN, 16>(ctx, g, e, gconst_dev, part_dev, m_ld, gsplit);
  }
}

static int launch_scan(gapit_ctx* ctx, gapit_geno* g, gapit_eigsys* e, int Q, const double* gconst_dev, double* part_dev,
                       int64_t m_ld, int gsplit) {
  switch (e->S) {
    case 4: return launch_scan_S<4, 256>(ctx, g, e, Q, gconst_dev, part_dev, m_ld, gsplit);
    case 5: return launch_scan_S<5, 240>(ctx, g, e, Q, gconst_dev, part_dev, m_ld, gsplit);
    case 6: return launch_scan_S<6, 240>(ctx, g, e, Q, gconst_dev, part_dev, m_ld, gsplit);
    case 7: return launch_scan_S<7, 224>(ctx, g, e, Q, gconst_dev, part_dev, m_ld, gsplit);
    default: return launch_scan_S<8, 256>(ctx, g, e, Q, gconst_dev, part_dev, m_ld, gsplit);
  }
}

// The multi-trait form of the rotation: same main loop and scheduler, store epilogue (StoreEpi).  `g` is a view of one
// marker block, `tiles` = eigenvector tiles + tiles of back-rotated columns in e->Vs.
template <int S, int BN>
static int launch_store_S(gapit_ctx* ctx, gapit_geno* g, gapit_eigsys* e, int64_t tiles, int gsplit, const double* kscale,
                          double* wbuf, int64_t m_ld, int nk_sq) {
  using Epi = StoreEpi<S, BN>;
  CUtensorMap map_a, map_b;
  GAPIT_CHECK(make_tensor_map_u8(&map_a, g->mk, static_cast<uint64_t>(g->ldn), static_cast<uint64_t>(g->m),
                                 static_cast<uint64_t>(g->ldn), 256));
  const int cl = (m_ld / 256 >= 2) ? 2 : 1;
  GAPIT_CHECK(make_tensor_map_u8(&map_b, e->Vs, static_cast<uint64_t>(e->ldn), static_cast<uint64_t>(tiles * BN),
                                 static_cast<uint64_t>(e->ldn), BN / cl));
  ScanSched::Params sp;
  sp.mtiles = static_cast<int>((m_ld / 256 + cl - 1) / cl);
  sp.cl = cl;
  sp.gsplit = static_cast<int>(std::min<int64_t>(gsplit, tiles));
  sp.nkq = static_cast<int>(tiles);
  sp.kblocks = static_cast<int>(g->ldn / kBlockK);
  sp.bn = BN;
  sp.hint_a = ptx::kEvictLast;
  sp.hint_b = ptx::kEvictNormal;
  sp.round_ctr = nullptr;
  sp.round_wait = 100000;   // ~50 us
  typename Epi::Params ep{kscale, wbuf, m_ld, nk_sq, epilogue_merges_planes(g->n)};
  const int nunits = std::min(ctx->sm_count / cl, sp.mtiles * sp.gsplit);
  constexpr int EPIW = 2;
  constexpr int kThreadsEpi = 128 + 2 * 128 * EPIW;
  if (cl == 2) {
    const size_t rounds = size_t((sp.mtiles * sp.gsplit + nunits - 1) / nunits) + 1;
    GAPIT_CHECK(ctx_ws_t(ctx, WS_SYNC, rounds, &sp.round_ctr));
    GAPIT_CUDA(cudaMemsetAsync(sp.round_ctr, 0, rounds * sizeof(unsigned int), ctx->stream));
    using PCfg = UmmaPairCfg<2, BN, 4>;
    auto kern = umma_i8_pair_kernel<2, BN, 4, EPIW, ScanSched, Epi, 1, 1>;
    GAPIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PCfg::kSmemBytes + 16));
    GAPIT_CUDA(launch_umma(kern, nunits, 2, kThreadsEpi, PCfg::kSmemBytes + 16, ctx->stream, map_a, map_b, sp, ep));
  } else {
    using Cfg = UmmaCfg<2, BN, 3>;
    auto kern = umma_i8_kernel<2, BN, 3, EPIW, ScanSched, Epi, 1, 1>;
    GAPIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes + 16));
    GAPIT_CUDA(launch_umma(kern, nunits, 1, kThreadsEpi, Cfg::kSmemBytes + 16, ctx->stream, map_a, map_b, sp, ep));
  }
  return GAPIT_OK;
}

static int launch_store(gapit_ctx* ctx, gapit_geno* g, gapit_eigsys* e, int64_t tiles, int gsplit, const double* kscale,
                        double* wbuf, int64_t m_ld, int nk_sq) {
  switch (e->S) {
    case 4: return launch_store_S<4, 256>(ctx, g, e, tiles, gsplit, kscale, wbuf, m_ld, nk_sq);
    case 5: return launch_store_S<5, 240>(ctx, g, e, tiles, gsplit, kscale, wbuf, m_ld, nk_sq);
    case 6: return launch_store_S<6, 240>(ctx, g, e, tiles, gsplit, kscale, wbuf, m_ld, nk_sq);
    case 7: return launch_store_S<7, 224>(ctx, g, e, tiles, gsplit, kscale, wbuf, m_ld, nk_sq);
    default: return launch_store_S<8, 256>(ctx, g, e, tiles, gsplit, kscale, wbuf, m_ld, nk_sq);
  }
}

template <int BN>
static int launch_glm_bn(gapit_ctx* ctx, gapit_geno* g, const int8_t* Rs_dev, int64_t rs_rows, const double* scale_dev,
                         int nc, int Q, double* part_dev, int64_t m_ld) {
  constexpr int S = 6;
  using Cfg = UmmaCfg<2, BN, 4>;
  using Epi = GlmEpi<S, BN>;
  CUtensorMap map_a, map_b;
  GAPIT_CHECK(make_tensor_map_u8(&map_a, g->mk, static_cast<uint64_t>(g->ldn), static_cast<uint64_t>(g->m),
                                 static_cast<uint64_t>(g->ldn), 256));
  GAPIT_CHECK(make_tensor_map_u8(&map_b, Rs_dev, static_cast<uint64_t>(g->ldn), static_cast<uint64_t>(rs_rows),
                                 static_cast<uint64_t>(g->ldn), BN));
  GlmSched::Params sp{static_cast<int>(m_ld / 256), static_cast<int>(g->ldn / kBlockK)};
  typename Epi::Params ep{scale_dev, g->colsumsq, part_dev, m_ld, g->m, nc, Q};
  auto kern = umma_i8_kernel<2, BN, 4, 1, GlmSched, Epi, 1, 1>;
  GAPIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  const int grid = std::min(ctx->sm_count, sp.mtiles);
  GAPIT_CUDA(launch_umma(kern, grid, 1, Cfg::kThreads, Cfg::kSmemBytes, ctx->stream, map_a, map_b, sp, ep));
  return GAPIT_OK;
}

// Rs_dev: digit planes of R = [X0 | y], rows (c*6 + s), `rs_rows` rows allocated (>= BN, zero beyond nc*6)
static int launch_glm(gapit_ctx* ctx, gapit_geno* g, const int8_t* Rs_dev, int64_t rs_rows, const double* scale_dev,
                      int nc, int Q, double* part_dev, int64_t m_ld) {
  const int cols = nc * 6;
  if (cols <= 16) return launch_glm_bn<16>(ctx, g, Rs_dev, rs_rows, scale_dev, nc, Q, part_dev, m_ld);
  if (cols <= 32) return launch_glm_bn<32>(ctx, g, Rs_dev, rs_rows, scale_dev, nc, Q, part_dev, m_ld);
  if (cols <= 64) return launch_glm_bn<64>(ctx, g, Rs_dev, rs_rows, scale_dev, nc, Q, part_dev, m_ld);
  return launch_glm_bn<112>(ctx, g, Rs_dev, rs_rows, scale_dev, nc, Q, part_dev, m_ld);
}

// eigen-index splits: keep the marker tiles of the CTAs that run together (148/g tiles of 256 x ldn bytes)
// inside ~48 MB of L2.  Depends on n only, so results do not change with the marker shard or SM count.
static int choose_gsplit(int64_t ldn, int64_t nkq) {
  const double bytes = 148.0 * 256.0 * static_cast<double>(ldn);
  const double budget = 48.0e6;
  int g = static_cast<int>(std::ceil(bytes / budget));
  g = std::max(1, g);
  return static_cast<int>(std::min<int64_t>(g, nkq));
}

int launch_xx_contract(cudaStream_t st, const double* W2, int64_t m_ld, const double* Wt, int wt_ld, int nk, double* XX,
                       int64_t xx_ld, int T);
int xx_trait_tile(int T);
int launch_vt_project(cudaStream_t st, const double* V, int64_t n, const double* R, int nc, double* out);
int launch_v_times_cols(cudaStream_t st, const double* V, int64_t n, const double* Z, int nc, double* out, bool square);
int launch_colstats(gapit_ctx* ctx, cudaStream_t st, const int8_t* mk, int64_t ldn, int64_t n, int64_t m, int32_t* colsum,
                    int32_t* colsumsq, int* bad_dev);
int launch_convert(cudaStream_t st, gapit_dtype dtype, const void* stage, int64_t n, int64_t mc, int8_t* dst, int64_t ldn,
                   int impute, int* bad_dev);
int launch_unpack2(gapit_ctx* ctx, cudaStream_t st, const void* stage, int64_t n, int64_t mc, int8_t* dst, int64_t ldn,
                   int impute, int* bad_dev, int recode);

// Everything about a scan that does not depend on the markers: the marker-free model of every trait
// (i = 0 of the reference loop), the per-eigenvalue constants of the rotation epilogue (MLM) or the digit
// planes of [X0 | y] (GLM), resident on the device.
struct ScanPlan {
  int64_t n = 0, ldn = 0;
  int q0 = 0, Q = 0, T = 0, glm = 0, gsplit = 1;
  gapit_eigsys* e = nullptr;
  std::vector<FinalizeConsts> fc;   // per trait
  double* gconst = nullptr;         // MLM, T = 1: [nk][Q+2] epilogue constants of the fused rotate + reduce
  int64_t nk = 0;
  int8_t* Rs = nullptr;             // GLM: [T][128][ldn] digit planes of [X0 | y_t]
  double* Rscale = nullptr;         // GLM: [T][GAPIT_MAX_Q0+1]
  FinalizeConsts* fc_dev = nullptr; // [T]
  std::vector<gapit_scan_base> bases;
  // MLM, T > 1: one rotation shared by all traits (StoreEpi + xx_contract)
  bool shared = false;
  int64_t lin_tiles = 0;            // tiles of back-rotated columns behind the eigenvector tiles of e->Vs
  double* kscale = nullptr;         // [(nkq + lin_tiles) * KPB]
  double* wt = nullptr;             // [nk][wt_ld]: w_t(k) = 1 / (lambda_k + delta_t)
  int wt_ld = 0;
};

// one block of markers resident on the device
struct MarkerView {
  const int8_t* mk;
  int64_t m;
  const int32_t* colsum;
  const int32_t* colsumsq;
};

// make room for `tiles` BN-row tiles in e->Vs (the eigenvector tiles are preserved)
static int eigsys_reserve_tiles(gapit_eigsys* e, int64_t tiles) {
  if (tiles <= e->tiles_cap) return GAPIT_OK;
  gapit_ctx* ctx = e->ctx;
  const int64_t cap = tiles + tiles / 16 + 1;
  int8_t* nv = nullptr;
  GAPIT_CUDA(cudaMalloc(&nv, size_t(cap) * e->BN * e->ldn));
  cudaError_t err = cudaMemcpyAsync(nv, e->Vs, size_t(e->nkq) * e->BN * e->ldn, cudaMemcpyDeviceToDevice, ctx->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->stream);
  if (err != cudaSuccess) {
    cudaFree(nv);
    return fail(GAPIT_ERR_CUDA, std::string("eigsys_reserve_tiles: ") + cudaGetErrorString(err));
  }
  cudaFree(e->Vs);
  e->Vs = nv;
  e->tiles_cap = cap;
  return GAPIT_OK;
}

static int scan_plan_build(gapit_ctx* ctx, int64_t n, int64_t ldn, gapit_eigsys* e, const double* X0, int q0, const double* Y,
                           int T, const double* delta, const double* vg, int glm, ScanPlan* plan) {
  if (!ctx || !X0 || !Y || T < 1) return fail(GAPIT_ERR_ARG, "gapit_p3d_scan: NULL argument");
  if (q0 < 1 || q0 > GAPIT_MAX_Q0) return fail(GAPIT_ERR_ARG, "gapit_p3d_scan: q0 must be 1.." + std::to_string(GAPIT_MAX_Q0));
  if (!glm && (!e || !delta || !vg)) return fail(GAPIT_ERR_ARG, "gapit_p3d_scan: MLM needs an eigen-system, delta and vg");
  if (!glm && e->n != n) return fail(GAPIT_ERR_ARG, "gapit_p3d_scan: eigen-system and genotypes disagree on n");
  if (n - q0 - 1 < 1) return fail(GAPIT_ERR_ARG, "gapit_p3d_scan: no residual degrees of freedom");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  plan->n = n;
  plan->ldn = ldn;
  plan->q0 = q0;
  plan->Q = pad_q(q0);
  plan->T = T;
  plan->glm = glm;
  plan->e = e;
  plan->gsplit = glm ? 1 : choose_gsplit(ldn, e->nkq);
  plan->fc.resize(T);
  plan->bases.assign(T, gapit_scan_base{});
  gapit_scan_base* base_out = plan->bases.data();
  const int Q = plan->Q;
  const int nc = q0 + T;

  double* dR = nullptr;
  GAPIT_CHECK(ctx_ws_t(ctx, WS_R, size_t(n) * nc, &dR));
  GAPIT_CUDA(cudaMemcpyAsync(dR, X0, size_t(n) * q0 * 8, cudaMemcpyHostToDevice, ctx->stream));
  GAPIT_CUDA(cudaMemcpyAsync(dR + size_t(n) * q0, Y, size_t(n) * T * 8, cudaMemcpyHostToDevice, ctx->stream));

  std::vector<double> proj(size_t(nc) * n);  // [(q0+T)][n]: rotated covariates then rotated phenotypes (raw ones for GLM)
  if (!glm) {
    double* dProj = nullptr;
    GAPIT_CHECK(ctx_ws_t(ctx, WS_PROJ, size_t(n) * nc, &dProj));
    GAPIT_CHECK(launch_vt_project(ctx->stream, e->V, n, dR, nc, dProj));
    GAPIT_CUDA(cudaMemcpyAsync(proj.data(), dProj, size_t(n) * nc * 8, cudaMemcpyDeviceToHost, ctx->stream));
    GAPIT_CUDA(cudaStreamSynchronize(ctx->stream));
    plan->nk = e->nkq * e->KPB;
    plan->shared = (T > 1);   // several traits: one rotation for all of them (StoreEpi + xx_contract)
    if (!plan->shared) GAPIT_CHECK(ctx_ws_t(ctx, WS_GCONST, size_t(plan->nk) * (Q + 2), &plan->gconst));
  } else {
    std::copy(X0, X0 + size_t(n) * q0, proj.begin());
    std::copy(Y, Y + size_t(n) * T, proj.begin() + size_t(n) * q0);
    GAPIT_CHECK(ctx_ws_t(ctx, WS_RS, size_t(T) * 128 * ldn, &plan->Rs));
    GAPIT_CHECK(ctx_ws_t(ctx, WS_RSCALE, size_t(T) * (GAPIT_MAX_Q0 + 1), &plan->Rscale));
    GAPIT_CUDA(cudaMemsetAsync(plan->Rs, 0, size_t(T) * 128 * ldn, ctx->stream));
  }

  std::vector<double> w(n), gconst, zcols, wt;
  const int lin_nc = (q0 + 1) * T;   // back-rotated columns of the shared-rotation form: [H_t^-1 X0 | H_t^-1 r_t] per trait
  if (plan->shared) {
    const int tile = xx_trait_tile(T);
    plan->wt_ld = (T + tile - 1) / tile * tile;
    zcols.assign(size_t(lin_nc) * n, 0.0);
    wt.assign(size_t(plan->nk) * plan->wt_ld, 0.0);
  }
  if (!glm && !plan->shared) gconst.assign(size_t(plan->nk) * (Q + 2), 0.0);
  const double dn = static_cast<double>(n);
  const double two_pi = 2.0 * 3.14159265358979323846;
  for (int t = 0; t < T; ++t) {
    const double* yt = proj.data() + size_t(q0 + t) * n;
    double sumlog = 0.0;
    for (int64_t k = 0; k < n; ++k) {
      if (glm) {
        w[k] = 1.0;
      } else {
        const double ld = e->values[k] + delta[t];
        w[k] = 1.0 / ld;
        sumlog += std::log(ld);
      }
    }
    // ---- i = 0 of the reference loop: the marker-free model (:617-620,:646,:678,:707-714,:772)
    FinalizeConsts& fc = plan->fc[t];
    std::fill(fc.iXX, fc.iXX + GAPIT_MAX_Q0 * GAPIT_MAX_Q0, 0.0);
    std::fill(fc.beta0, fc.beta0 + GAPIT_MAX_Q0, 0.0);
    std::vector<double> XX(q0 * q0, 0.0), XY(q0, 0.0), inv(q0 * q0, 0.0);
    for (int i = 0; i < q0; ++i) {
      const double* xi = proj.data() + size_t(i) * n;
      for (int j = 0; j <= i; ++j) {
        const double* xj = proj.data() + size_t(j) * n;
        double s = 0.0;
        for (int64_t k = 0; k < n; ++k) s += w[k] * xi[k] * xj[k];
        XX[i * q0 + j] = XX[j * q0 + i] = s;
      }
      double s = 0.0;
      for (int64_t k = 0; k < n; ++k) s += w[k] * xi[k] * yt[k];
      XY[i] = s;
    }
    int singular = 0;
    sym_inverse(q0, XX.data(), inv.data(), &singular);
    for (int i = 0; i < q0; ++i) {
      double s = 0.0;
      for (int j = 0; j < q0; ++j) {
        fc.iXX[i * q0 + j] = inv[i * q0 + j];
        s += inv[i * q0 + j] * XY[j];
      }
      fc.beta0[i] = s;
    }
    double rss0 = 0.0;
    for (int64_t k = 0; k < n; ++k) {
      double r = yt[k];
      for (int j = 0; j < q0; ++j) r -= fc.beta0[j] * proj[size_t(j) * n + k];
      rss0 += w[k] * r * r;
    }
    // intercept-only OLS likelihood on the raw phenotype (:536-553)
    const double* yraw = Y + size_t(t) * n;
    double ybar = 0.0;
    for (int64_t k = 0; k < n; ++k) ybar += yraw[k];
    ybar /= dn;
    double ss0 = 0.0;
    for (int64_t k = 0; k < n; ++k) ss0 += (yraw[k] - ybar) * (yraw[k] - ybar);
    fc.logL0 = 0.5 * (-dn * std::log((two_pi / dn) * ss0) - dn);
    fc.rss0 = rss0;
    fc.sumlog = sumlog;
    fc.n = dn;
    fc.df = static_cast<double>(n - q0 - 1);
    // log B(df/2, 1/2) in extended precision: the two lgamma terms (~1e5 at n = 20,000) cancel to O(log n), and the
    // p-values of the bulk of the markers are 1 - (lower part), which magnifies an absolute error here by 1 / p
    fc.lbeta = static_cast<double>(lgammal(0.5L * (long double)fc.df) + lgammal(0.5L) - lgammal(0.5L * (long double)fc.df + 0.5L));
    fc.q0 = q0;
    fc.qpad = Q;
    const double ves_glm = rss0 / static_cast<double>(n - q0);
    fc.var_scale = glm ? ves_glm : vg[t];
    if (base_out) {
      gapit_scan_base& b = base_out[t];
      b.logL0 = fc.logL0;
      b.logLM_base = 0.5 * (-dn * std::log((two_pi / dn) * rss0) - sumlog - dn);
      b.rsquare_base = 1.0 - std::exp(-(2.0 / dn) * (b.logLM_base - fc.logL0));
      b.ves = glm ? ves_glm : delta[t] * vg[t];
      b.reml = glm ? (-0.5 * (dn - q0) * std::log(ves_glm) - 0.5 * dn - 0.5 * (dn - q0) * std::log(two_pi)) : 0.0;
      b.df = fc.df;
      std::fill(b.beta0, b.beta0 + GAPIT_MAX_Q0, 0.0);
      for (int j = 0; j < q0; ++j) b.beta0[j] = fc.beta0[j];
      b.singular_x0 = singular;
    }
    if (plan->shared) {
      // Z_t = diag(w_t) V'[X0 | r_t], r_t = y_t - X0 beta0_t: V Z_t = H_t^-1 [X0 | r_t].  Projecting the reduced-model
      // residual instead of y_t gives num = x'H^-1 r directly (no xy - xa'beta0 cancellation at finalize).
      double* zt = zcols.data() + size_t(t) * (q0 + 1) * n;
      for (int64_t k = 0; k < n; ++k) {
        double r = yt[k];
        for (int j = 0; j < q0; ++j) {
          zt[size_t(j) * n + k] = w[k] * proj[size_t(j) * n + k];
          r -= fc.beta0[j] * proj[size_t(j) * n + k];
        }
        zt[size_t(q0) * n + k] = w[k] * r;
        wt[size_t(k) * plan->wt_ld + t] = w[k];
      }
    } else if (!glm) {
      double* gt = gconst.data();
      for (int64_t k = 0; k < n; ++k) {
        const double c = e->h_scale[k];
        double* gk = gt + size_t(k) * (Q + 2);
        gk[0] = w[k] * c * c;
        for (int j = 0; j < q0; ++j) gk[1 + j] = w[k] * c * proj[size_t(j) * n + k];
        gk[Q + 1] = w[k] * c * yt[k];
      }
    } else {
      // R_t = [X0 | y_t] (contiguous columns) -> 6 int8 digit planes per column, same split as the eigenvectors
      double* dRg = nullptr;
      GAPIT_CHECK(ctx_ws_t(ctx, WS_RG, size_t(GAPIT_MAX_Q0 + 1) * n, &dRg));
      GAPIT_CUDA(cudaMemcpyAsync(dRg, dR, size_t(n) * q0 * 8, cudaMemcpyDeviceToDevice, ctx->stream));
      GAPIT_CUDA(cudaMemcpyAsync(dRg + size_t(q0) * n, dR + size_t(q0 + t) * n, size_t(n) * 8, cudaMemcpyDeviceToDevice,
                                 ctx->stream));
      slice_v_kernel<<<static_cast<unsigned>(q0 + 1), 256, 0, ctx->stream>>>(
          dRg, n, 6, plan->Rs + size_t(t) * 128 * ldn, ldn, plan->Rscale + size_t(t) * (GAPIT_MAX_Q0 + 1));
      GAPIT_CUDA(cudaGetLastError());
    }
  }
  std::vector<FinalizeConsts> fcs(plan->fc);
  if (plan->shared)
    for (auto& f : fcs) std::fill(f.beta0, f.beta0 + GAPIT_MAX_Q0, 0.0);   // the xy column already is x'H^-1 (y - X0 beta0)
  GAPIT_CHECK(ctx_ws_t(ctx, WS_FC, size_t(T), &plan->fc_dev));
  GAPIT_CUDA(cudaMemcpyAsync(plan->fc_dev, fcs.data(), size_t(T) * sizeof(FinalizeConsts), cudaMemcpyHostToDevice, ctx->stream));
  if (plan->shared) {
    // back-rotate: R = V Z (n x lin_nc), then its digit planes behind the eigenvector tiles of e->Vs
    double *dZ = nullptr, *dRlin = nullptr;
    GAPIT_CHECK(ctx_ws_t(ctx, WS_Z, zcols.size(), &dZ));
    GAPIT_CHECK(ctx_ws_t(ctx, WS_RG, zcols.size(), &dRlin));
    GAPIT_CHECK(ctx_ws_t(ctx, WS_WT, wt.size(), &plan->wt));
    GAPIT_CUDA(cudaMemcpyAsync(dZ, zcols.data(), zcols.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    GAPIT_CUDA(cudaMemcpyAsync(plan->wt, wt.data(), wt.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    GAPIT_CHECK(launch_v_times_cols(ctx->stream, e->V, n, dZ, lin_nc, dRlin, false));
    plan->lin_tiles = (lin_nc + e->KPB - 1) / e->KPB;
    GAPIT_CHECK(eigsys_reserve_tiles(e, e->nkq + plan->lin_tiles));
    const size_t rows = size_t(e->nkq + plan->lin_tiles) * e->KPB;
    GAPIT_CHECK(ctx_ws_t(ctx, WS_KSCALE, rows, &plan->kscale));
    GAPIT_CUDA(cudaMemsetAsync(plan->kscale, 0, rows * 8, ctx->stream));
    GAPIT_CUDA(cudaMemcpyAsync(plan->kscale, e->scale, size_t(plan->nk) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    int8_t* tail = e->Vs + size_t(e->nkq) * e->BN * e->ldn;
    GAPIT_CUDA(cudaMemsetAsync(tail, 0, size_t(plan->lin_tiles) * e->BN * e->ldn, ctx->stream));
    slice_v_kernel<<<static_cast<unsigned>(lin_nc), 256, 0, ctx->stream>>>(dRlin, n, e->S, tail, e->ldn, plan->kscale + plan->nk);
    GAPIT_CUDA(cudaGetLastError());
  } else if (!glm) {
    GAPIT_CUDA(cudaMemcpyAsync(plan->gconst, gconst.data(), gconst.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  GAPIT_CUDA(cudaStreamSynchronize(ctx->stream));  // the pageable host sources must outlive their copies
  return GAPIT_OK;
}

// 64-bit mix of a byte range (word-wise multiply-xorshift); keys the plan cache
static uint64_t mix_bytes(const void* p, size_t nbytes, uint64_t h) {
  const unsigned char* b = static_cast<const unsigned char*>(p);
  size_t i = 0;
  for (; i + 8 <= nbytes; i += 8) {
    uint64_t w;
    memcpy(&w, b + i, 8);
    h = (h ^ w) * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
  }
  uint64_t w = 0;
  if (i < nbytes) memcpy(&w, b + i, nbytes - i);
  h = (h ^ w ^ (uint64_t(nbytes) << 32)) * 0xC2B2AE3D27D4EB4Full;
  return h ^ (h >> 32);
}

// The plan of a scan depends on the traits, covariates, variance components and the eigen-system, not on the markers:
// GAPIT's by-file loop (GAPIT.EMMAxP3D.R:259-331) and any block-by-block driver call the scan again and again with the
// same ones, so the last plan stays on the context, keyed by a hash of every input it was built from.
static int scan_prepare(gapit_ctx* ctx, int64_t n, int64_t ldn, gapit_eigsys* e, const double* X0, int q0, const double* Y,
                        int T, const double* delta, const double* vg, int glm, ScanPlan** plan_out, gapit_scan_base* base_out) {
  if (!ctx || !X0 || !Y || T < 1) return fail(GAPIT_ERR_ARG, "gapit_p3d_scan: NULL argument");
  if (q0 < 1 || q0 > GAPIT_MAX_Q0) return fail(GAPIT_ERR_ARG, "gapit_p3d_scan: q0 must be 1.." + std::to_string(GAPIT_MAX_Q0));
  if (!glm && (!e || !delta || !vg)) return fail(GAPIT_ERR_ARG, "gapit_p3d_scan: MLM needs an eigen-system, delta and vg");
  uint64_t key = 0x243F6A8885A308D3ull;
  const int64_t dims[6] = {n, ldn, q0, T, glm, (!glm && e) ? int64_t(e->serial) * 16 + e->S : 0};
  key = mix_bytes(dims, sizeof(dims), key);
  key = mix_bytes(X0, size_t(n) * q0 * 8, key);
  key = mix_bytes(Y, size_t(n) * T * 8, key);
  if (!glm) {
    key = mix_bytes(delta, size_t(T) * 8, key);
    key = mix_bytes(vg, size_t(T) * 8, key);
  }
  if (key == 0) key = 1;
  if (!ctx->plan_cache || ctx->plan_key != key) {
    scan_plan_cache_drop(ctx);
    ScanPlan* p = new ScanPlan();
    const int rc = scan_plan_build(ctx, n, ldn, e, X0, q0, Y, T, delta, vg, glm, p);
    if (rc != GAPIT_OK) {
      delete p;
      return rc;
    }
    ctx->plan_cache = p;
    ctx->plan_key = key;
  }
  *plan_out = ctx->plan_cache;
  if (base_out) std::copy(ctx->plan_cache->bases.begin(), ctx->plan_cache->bases.end(), base_out);
  return GAPIT_OK;
}

// copy rows [k0, k0 + mc) of every trait's result arrays to the caller's SoA arrays (m_total markers per trait):
// one strided copy per array
static int fetch_block(cudaStream_t st, const double* out_dev, int64_t out_tstride, int64_t out_ld, int64_t k0, int64_t mc,
                       int T, int64_t m_total, const gapit_scan_out* out, int64_t host_off) {
  double* dsts[7] = {out->effect, out->tvalue, out->se, out->pvalue, out->rsquare, out->loglm, out->maf};
  for (int a = 0; a < 7; ++a) {
    if (!dsts[a]) continue;
    GAPIT_CUDA(cudaMemcpy2DAsync(dsts[a] + host_off + k0, size_t(m_total) * 8, out_dev + size_t(a) * out_ld + k0, size_t(out_tstride) * 8,
                                 size_t(mc) * 8, a == 6 ? 1 : size_t(T), cudaMemcpyDeviceToHost, st));
  }
  return GAPIT_OK;
}

static int launch_finalize(gapit_ctx* ctx, const ScanPlan& plan, const FinalizeSrc& src, const MarkerView& v, double* out,
                           int64_t out_tstride, int64_t out_ld) {
  dim3 grid(static_cast<unsigned>((v.m + 255) / 256), static_cast<unsigned>(plan.T));
  finalize_kernel<<<grid, 256, 0, ctx->stream>>>(src, v.m, v.colsum, v.colsumsq, plan.fc_dev, out, out_tstride, out_ld);
  GAPIT_CUDA(cudaGetLastError());
  return GAPIT_OK;
}

// markers per pass of the shared-rotation form: whole rounds of the persistent grid, W^2 block within ~1/3 of the free
// device memory (at most 8 GB)
static int64_t shared_block_markers(gapit_ctx* ctx, int64_t rows) {
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = size_t(8) << 30;
  size_t cap = 0;
  if (ctx->ws_ptr[WS_WBUF]) cap = ctx->ws_cap[WS_WBUF];   // already ours: counts as available
  const double budget = std::min(8.0e9, (double(free_b) + double(cap)) / 3.0);
  const int64_t unit = int64_t(ctx->sm_count / 2) * 512;
  int64_t mb = static_cast<int64_t>(budget / (double(rows) * 8.0));
  if (const char* ev = getenv("GAPIT_SHARED_BLOCK")) mb = atoll(ev);   // tests: force several blocks
  if (mb >= unit) return mb / unit * unit;
  return std::max<int64_t>(512, mb / 512 * 512);
}

// scan one block of markers for every trait of the plan; results of trait t go to
// out_dev[t * out_tstride + a * out_ld + out_off + k], a = 0..6.  With host_out the finished rows are copied to the
// caller's arrays (m_total markers per trait, rows out_off..) on the D2H stream while the next rows are computed.
static int scan_run(gapit_ctx* ctx, const ScanPlan& plan, const MarkerView& v, double* out_dev, int64_t out_tstride,
                    int64_t out_ld, int64_t out_off, cudaEvent_t after_reduce = nullptr,
                    const gapit_scan_out* host_out = nullptr, int64_t m_total = 0, int64_t host_off = 0) {
  const int Q = plan.Q;
  gapit_geno view;  // borrowed pointers only
  view.ctx = ctx;
  view.n = plan.n;
  view.ldn = plan.ldn;
  if (plan.shared) {
    gapit_eigsys* e = plan.e;
    const int64_t tiles = e->nkq + plan.lin_tiles;
    const int64_t rows = tiles * e->KPB;
    const int64_t mb = std::min(shared_block_markers(ctx, rows), round_up(v.m, 256));
    double *wbuf = nullptr, *xx = nullptr;
    GAPIT_CHECK(ctx_ws_t(ctx, WS_WBUF, size_t(rows) * mb, &wbuf));
    GAPIT_CHECK(ctx_ws_t(ctx, WS_XX, size_t(plan.T) * mb, &xx));
    for (int64_t k0 = 0; k0 < v.m; k0 += mb) {
      const int64_t mc = std::min(mb, v.m - k0);
      const int64_t m_ld = round_up(mc, 256);
      view.m = mc;
      view.mk = const_cast<int8_t*>(v.mk) + size_t(k0) * plan.ldn;
      GAPIT_CHECK(launch_store(ctx, &view, e, tiles, plan.gsplit, plan.kscale, wbuf, m_ld, static_cast<int>(plan.nk)));
      GAPIT_CHECK(launch_xx_contract(ctx->stream, wbuf, m_ld, plan.wt, plan.wt_ld, static_cast<int>(plan.nk), xx, m_ld, plan.T));
      FinalizeSrc src;
      src.xx = xx;
      src.xx_t = m_ld;
      src.xx_g = 0;
      src.xa = wbuf + size_t(plan.nk) * m_ld;
      src.xa_t = int64_t(plan.q0 + 1) * m_ld;
      src.xa_j = m_ld;
      src.xa_g = 0;
      src.xy = src.xa + size_t(plan.q0) * m_ld;
      src.xy_t = src.xa_t;
      src.xy_g = 0;
      src.nsplit = 1;
      MarkerView blk{view.mk, mc, v.colsum + k0, v.colsumsq + k0};
      GAPIT_CHECK(launch_finalize(ctx, plan, src, blk, out_dev + out_off + k0, out_tstride, out_ld));
      if (host_out) {
        GAPIT_CUDA(cudaEventRecord(ctx->ev_blk, ctx->stream));
        GAPIT_CUDA(cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_blk, 0));
        GAPIT_CHECK(fetch_block(ctx->d2h_stream, out_dev, out_tstride, out_ld, out_off + k0, mc, plan.T, m_total, host_out, host_off));
      }
    }
    if (after_reduce) cudaEventRecord(after_reduce, ctx->stream);
    return GAPIT_OK;
  }
  const int64_t m_ld = round_up(v.m, 256);
  const int nsplit = plan.glm ? 1 : plan.gsplit * kScanHalves;   // partial sums per marker
  const int64_t part_tstride = int64_t(nsplit) * (Q + 2) * m_ld;
  double* part = nullptr;
  GAPIT_CHECK(ctx_ws_t(ctx, WS_PART, size_t(plan.T) * part_tstride, &part));
  view.m = v.m;
  view.mk = const_cast<int8_t*>(v.mk);
  view.colsum = const_cast<int32_t*>(v.colsum);
  view.colsumsq = const_cast<int32_t*>(v.colsumsq);
  if (!plan.glm) {
    GAPIT_CHECK(launch_scan(ctx, &view, plan.e, Q, plan.gconst, part, m_ld, plan.gsplit));   // single trait
  } else {
    for (int t = 0; t < plan.T; ++t)
      GAPIT_CHECK(launch_glm(ctx, &view, plan.Rs + size_t(t) * 128 * plan.ldn, 128,
                             plan.Rscale + size_t(t) * (GAPIT_MAX_Q0 + 1), plan.q0 + 1, Q, part + t * part_tstride, m_ld));
  }
  if (after_reduce) cudaEventRecord(after_reduce, ctx->stream);
  FinalizeSrc src;
  src.xx = part;
  src.xa = part + m_ld;
  src.xy = part + int64_t(Q + 1) * m_ld;
  src.xx_t = src.xa_t = src.xy_t = part_tstride;
  src.xx_g = src.xa_g = src.xy_g = int64_t(Q + 2) * m_ld;
  src.xa_j = m_ld;
  src.nsplit = nsplit;
  GAPIT_CHECK(launch_finalize(ctx, plan, src, v, out_dev + out_off, out_tstride, out_ld));
  if (host_out) {
    GAPIT_CUDA(cudaEventRecord(ctx->ev_blk, ctx->stream));
    GAPIT_CUDA(cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_blk, 0));
    GAPIT_CHECK(fetch_block(ctx->d2h_stream, out_dev, out_tstride, out_ld, out_off, v.m, plan.T, m_total, host_out, host_off));
  }
  return GAPIT_OK;
}

static int check_out(const gapit_scan_out* out) {
  if (!out || !out->effect || !out->tvalue || !out->se || !out->pvalue || !out->rsquare || !out->loglm)
    return fail(GAPIT_ERR_ARG, "gapit_p3d_scan: output arrays effect/tvalue/se/pvalue/rsquare/loglm are required");
  return GAPIT_OK;
}

// every exit path of a scan leaves no work queued that still reads the caller's buffers (the header promises calls
// are synchronous on return)
struct StreamDrain {
  gapit_ctx* c;
  ~StreamDrain() {
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamSynchronize(c->d2h_stream);
  }
};

// size the per-block workspaces of scan_run once for blocks of up to max_m markers (no regrow, hence no stream
// synchronisation, in the middle of a pipelined call)
static int scan_reserve(gapit_ctx* ctx, const ScanPlan& plan, int64_t max_m) {
  void* p = nullptr;
  const int64_t m_ld = round_up(max_m, 256);
  if (plan.shared) {
    const int64_t rows = (plan.e->nkq + plan.lin_tiles) * plan.e->KPB;
    const int64_t mb = std::min(shared_block_markers(ctx, rows), m_ld);
    GAPIT_CHECK(ctx_ws(ctx, WS_WBUF, size_t(rows) * mb * 8, &p));
    GAPIT_CHECK(ctx_ws(ctx, WS_XX, size_t(plan.T) * mb * 8, &p));
  } else {
    const int nsplit = plan.glm ? 1 : plan.gsplit * kScanHalves;
    GAPIT_CHECK(ctx_ws(ctx, WS_PART, size_t(plan.T) * nsplit * (plan.Q + 2) * m_ld * 8, &p));
  }
  return GAPIT_OK;
}

// m_total / m_off: the caller's result arrays hold m_total markers per trait and this handle's markers start at m_off
// (a marker shard of a table scanned by several GPUs); m_total = 0: the arrays hold exactly g->m markers
int scan_impl(gapit_ctx* ctx, gapit_geno* g, gapit_eigsys* e, const double* X0, int q0, const double* Y, int T,
              const double* delta, const double* vg, int glm, gapit_scan_out* out, gapit_scan_base* base_out,
              int64_t m_total = 0, int64_t m_off = 0) {
  if (!g) return fail(GAPIT_ERR_ARG, "gapit_p3d_scan: NULL genotype handle");
  GAPIT_CHECK(check_out(out));
  ScanPlan* planp = nullptr;
  const bool trace = getenv("GAPIT_TRACE") != nullptr;   // host-side phase times on stderr
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms_since = [&](std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(now() - t0).count();
  };
  auto t_start = now();
  StreamDrain drain{ctx};
  GAPIT_CHECK(scan_prepare(ctx, g->n, g->ldn, e, X0, q0, Y, T, delta, vg, glm, &planp, base_out));
  const ScanPlan& plan = *planp;
  if (trace) fprintf(stderr, "[gapit] scan_prepare %.2f ms\n", ms_since(t_start));
  const int64_t m = g->m;
  double* o = nullptr;
  GAPIT_CHECK(ctx_ws_t(ctx, WS_OUT, size_t(T) * 7 * m, &o));
  GAPIT_CHECK(scan_reserve(ctx, plan, m));
  MarkerView v{g->mk, m, g->colsum, g->colsumsq};
  cudaEventRecord(ctx->ev0, ctx->stream);
  GAPIT_CHECK(scan_run(ctx, plan, v, o, 7 * m, m, 0, ctx->ev3, out, m_total > 0 ? m_total : m, m_off));
  cudaEventRecord(ctx->ev1, ctx->stream);
  GAPIT_CUDA(cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev1, 0));
  cudaEventRecord(ctx->ev2, ctx->d2h_stream);
  if (trace) fprintf(stderr, "[gapit] launches enqueued at %.2f ms\n", ms_since(t_start));
  GAPIT_CUDA(cudaStreamSynchronize(ctx->stream));
  GAPIT_CUDA(cudaStreamSynchronize(ctx->d2h_stream));
  if (trace) fprintf(stderr, "[gapit] scan done at %.2f ms\n", ms_since(t_start));
  float ms = 0.f;
  ctx->stats = gapit_stats{};
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev3);
  ctx->stats.kernel_ms = ms;   // rotate_reduce or glm_reduce; shared rotation: every block's rotation + contraction + finalize
  cudaEventElapsedTime(&ms, ctx->ev3, ctx->ev1);
  ctx->stats.post_ms = ms;     // finalize
  cudaEventElapsedTime(&ms, ctx->ev1, ctx->ev2);
  ctx->stats.d2h_ms = ms;      // result copies not hidden behind the kernels
  return GAPIT_OK;
}

// Streaming form: the host genotype matrix is uploaded in marker blocks on a copy stream while the previous
// block is scanned, so a host-to-results call costs about max(H2D, scan) instead of their sum.
static int scan_host_impl(gapit_ctx* ctx, const void* snps, gapit_dtype dtype, int64_t n, int64_t m, int64_t ld,
                          gapit_impute impute, gapit_eigsys* e, const double* X0, int q0, const double* Y, int T,
                          const double* delta, const double* vg, int glm, gapit_scan_out* out, gapit_scan_base* base_out) {
  const int64_t nb2 = (n + 3) / 4;   // bytes per marker of the 2-bit packed form
  if (!snps || n <= 0 || m <= 0 || ld < (is_packed2(dtype) ? nb2 : n))
    return fail(GAPIT_ERR_ARG, "gapit_p3d_scan_host: bad genotype argument");
  GAPIT_CHECK(check_out(out));
  const int64_t ldn = round_up(n, 128);
  ScanPlan* planp = nullptr;
  StreamDrain drain{ctx};
  GAPIT_CHECK(scan_prepare(ctx, n, ldn, e, X0, q0, Y, T, delta, vg, glm, &planp, base_out));
  const ScanPlan& plan = *planp;
  const size_t esz = dtype == GAPIT_F64 ? 8 : (dtype == GAPIT_I32 ? 4 : 1);
  const size_t row_bytes = is_packed2(dtype) ? size_t(nb2) : size_t(n) * esz;   // dense bytes of one marker on the host
  // marker block: ~320 MB of int8, at most 65535 markers for the convert grid -- and WHOLE ROUNDS of the persistent
  // grid: every launch of the tensor kernels costs ceil(jobs / units) rounds of equal jobs, a block of sm_count / 2 pair
  // tiles (x 512 markers) fills every round whatever the eigen split, so only one block per call (the first, which takes
  // the remainder) pays a partly filled round.  C2 in 65,280-marker blocks was 70 rounds against 67 resident.
  const int64_t quantum = int64_t(std::max(1, ctx->sm_count / 2)) * 512;
  int64_t chunk = (int64_t(320) << 20) / ldn;
  if (dtype != GAPIT_I8) chunk = std::min<int64_t>(chunk, 65280);
  chunk = chunk >= quantum ? chunk / quantum * quantum : std::max<int64_t>(256, chunk / 256 * 256);
  if (const char* ev = getenv("GAPIT_STREAM_CHUNK")) chunk = std::max<int64_t>(256, atoll(ev) / 256 * 256);  // tests
  chunk = std::min(chunk, round_up(m, 256));
  int8_t* mk[2];
  int32_t *cs[2], *cq[2];
  void* stage[2] = {nullptr, nullptr};
  int* bad = nullptr;
  double* o = nullptr;
  for (int b = 0; b < 2; ++b) {
    GAPIT_CHECK(ctx_ws_t(ctx, WS_CHUNK0 + b, size_t(chunk) * ldn, &mk[b]));
    GAPIT_CHECK(ctx_ws_t(ctx, WS_CS0 + b, size_t(chunk), &cs[b]));
    GAPIT_CHECK(ctx_ws_t(ctx, WS_CQ0 + b, size_t(chunk), &cq[b]));
    if (dtype != GAPIT_I8 || (ld == n && ldn != n)) GAPIT_CHECK(ctx_ws(ctx, WS_STAGE0 + b, size_t(chunk) * row_bytes + 8, &stage[b]));
    if (ldn != n && !is_packed2(dtype)) GAPIT_CUDA(cudaMemsetAsync(mk[b], 0, size_t(chunk) * ldn, ctx->stream));
  }
  GAPIT_CHECK(ctx_ws_t(ctx, WS_BAD, 4, &bad));
  GAPIT_CHECK(ctx_ws_t(ctx, WS_OUT, size_t(T) * 7 * m, &o));
  GAPIT_CHECK(scan_reserve(ctx, plan, chunk));
  GAPIT_CUDA(cudaMemsetAsync(bad, 0, 4, ctx->stream));
  cudaEventRecord(ctx->ev0, ctx->stream);
  // the copy stream must not start before the workspaces are initialised
  GAPIT_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev0, 0));
  // the first block takes the remainder (its upload cannot overlap anything; every later block is a full one)
  const int64_t first = m <= chunk ? m : ((m % chunk) ? (m % chunk) : chunk);
  const int64_t nchunks = 1 + (m - first) / chunk;
  for (int64_t c = 0; c < nchunks; ++c) {
    const int b = static_cast<int>(c & 1);
    const int64_t k0 = (c == 0) ? 0 : first + (c - 1) * chunk;
    const int64_t mc = (c == 0) ? first : std::min(chunk, m - k0);
    const char* src = static_cast<const char*>(snps) + size_t(k0) * ld * esz;
    if (c >= 2) GAPIT_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[b], 0));   // buffer b scanned
    if (dtype == GAPIT_I8 && stage[b]) {
      // dense host rows: one flat PCIe transfer, then re-pitch on the device (a strided H2D of n-byte rows is ~2x slower)
      GAPIT_CUDA(cudaMemcpyAsync(stage[b], src, size_t(mc) * n, cudaMemcpyHostToDevice, ctx->copy_stream));
      GAPIT_CUDA(cudaMemcpy2DAsync(mk[b], ldn, stage[b], n, n, mc, cudaMemcpyDeviceToDevice, ctx->copy_stream));
    } else if (dtype == GAPIT_I8) {
      GAPIT_CUDA(cudaMemcpy2DAsync(mk[b], ldn, src, ld, n, mc, cudaMemcpyHostToDevice, ctx->copy_stream));
    } else if (is_packed2(dtype)) {
      // a quarter of the int8 bytes over PCIe, expanded to the int8 operand layout on the device
      if (ld == nb2) GAPIT_CUDA(cudaMemcpyAsync(stage[b], src, size_t(mc) * nb2, cudaMemcpyHostToDevice, ctx->copy_stream));
      else GAPIT_CUDA(cudaMemcpy2DAsync(stage[b], nb2, src, ld, nb2, mc, cudaMemcpyHostToDevice, ctx->copy_stream));
      GAPIT_CHECK(launch_unpack2(ctx, ctx->copy_stream, stage[b], n, mc, mk[b], ldn, int(impute), bad, packed2_recode(dtype)));
    } else {
      GAPIT_CUDA(cudaMemcpy2DAsync(stage[b], size_t(n) * esz, src, size_t(ld) * esz, size_t(n) * esz, mc,
                                   cudaMemcpyHostToDevice, ctx->copy_stream));
      GAPIT_CHECK(launch_convert(ctx->copy_stream, dtype, stage[b], n, mc, mk[b], ldn, int(impute), bad));
    }
    GAPIT_CUDA(cudaEventRecord(ctx->ev_ready[b], ctx->copy_stream));
    GAPIT_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_ready[b], 0));
    GAPIT_CHECK(launch_colstats(ctx, ctx->stream, mk[b], ldn, n, mc, cs[b], cq[b], bad));
    MarkerView v{mk[b], mc, cs[b], cq[b]};
    GAPIT_CHECK(scan_run(ctx, plan, v, o, 7 * m, m, k0, nullptr, out, m));
    GAPIT_CUDA(cudaEventRecord(ctx->ev_free[b], ctx->stream));
  }
  cudaEventRecord(ctx->ev1, ctx->stream);
  int hbad = 0;
  GAPIT_CUDA(cudaMemcpyAsync(&hbad, bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
  GAPIT_CUDA(cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev1, 0));
  cudaEventRecord(ctx->ev2, ctx->d2h_stream);
  GAPIT_CUDA(cudaStreamSynchronize(ctx->stream));
  GAPIT_CUDA(cudaStreamSynchronize(ctx->copy_stream));
  GAPIT_CUDA(cudaStreamSynchronize(ctx->d2h_stream));
  if (hbad & 2) return fail(GAPIT_ERR_DATA, "genotype matrix holds NA and no impute rule was given");
  if (hbad & 1) return fail(GAPIT_ERR_DATA, "genotype matrix holds a value outside {0,1,2}");
  float ms = 0.f;
  ctx->stats = gapit_stats{};
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  ctx->stats.kernel_ms = ms;   // upload + pack + scan, overlapped
  cudaEventElapsedTime(&ms, ctx->ev1, ctx->ev2);
  ctx->stats.d2h_ms = ms;
  return GAPIT_OK;
}

}  // namespace gapit

namespace gapit {
void scan_plan_cache_drop(gapit_ctx* ctx) {
  if (!ctx || !ctx->plan_cache) return;
  delete ctx->plan_cache;
  ctx->plan_cache = nullptr;
  ctx->plan_key = 0;
}
}  // namespace gapit

using namespace gapit;

static int eigsys_build(gapit_ctx* ctx, const double* vectors, cudaMemcpyKind kind, const double* values, int64_t n,
                        gapit_eigsys** out) {
  if (!ctx || !vectors || !values || !out || n <= 0) return fail(GAPIT_ERR_ARG, "gapit_eigsys_upload: bad argument");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  gapit_eigsys* e = new gapit_eigsys();
  static std::atomic<uint64_t> next_serial{1};
  e->serial = next_serial.fetch_add(1);
  e->ctx = ctx;
  e->device = ctx->device;
  e->n = n;
  const SliceGeom sg = slice_geom(ctx->slices);
  e->S = sg.S;
  e->BN = sg.BN;
  e->KPB = sg.KPB;
  e->ldn = round_up(n, 128);
  e->nkq = (n + e->KPB - 1) / e->KPB;
  e->tiles_cap = e->nkq;
  e->values.assign(values, values + n);
  const int64_t nk = e->nkq * e->KPB;
  cudaError_t err = cudaMalloc(&e->V, size_t(n) * n * 8);
  if (err == cudaSuccess) err = cudaMalloc(&e->Vs, size_t(e->nkq) * e->BN * e->ldn);
  if (err == cudaSuccess) err = cudaMalloc(&e->scale, size_t(nk) * 8);
  if (err == cudaSuccess) err = cudaMemcpyAsync(e->V, vectors, size_t(n) * n * 8, kind, ctx->stream);
  if (err == cudaSuccess) err = cudaMemsetAsync(e->Vs, 0, size_t(e->nkq) * e->BN * e->ldn, ctx->stream);
  if (err == cudaSuccess) err = cudaMemsetAsync(e->scale, 0, size_t(nk) * 8, ctx->stream);
  if (err == cudaSuccess) {
    slice_v_kernel<<<static_cast<unsigned>(n), 256, 0, ctx->stream>>>(e->V, n, e->S, e->Vs, e->ldn, e->scale);
    err = cudaGetLastError();
  }
  e->h_scale.assign(size_t(nk), 0.0);
  if (err == cudaSuccess) err = cudaMemcpyAsync(e->h_scale.data(), e->scale, size_t(nk) * 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->stream);
  if (err != cudaSuccess) {
    gapit_eigsys_free(e);
    return fail(GAPIT_ERR_CUDA, std::string("gapit_eigsys_upload: ") + cudaGetErrorString(err));
  }
  *out = e;
  return GAPIT_OK;
}

extern "C" int gapit_eigsys_upload(gapit_ctx* ctx, const double* vectors, const double* values, int64_t n,
                                   gapit_eigsys** out) {
  return eigsys_build(ctx, vectors, cudaMemcpyHostToDevice, values, n, out);
}

extern "C" int gapit_eigsys_from_device(gapit_ctx* ctx, const double* vectors_dev, const double* values, int64_t n,
                                        gapit_eigsys** out) {
  return eigsys_build(ctx, vectors_dev, cudaMemcpyDeviceToDevice, values, n, out);
}

// out[c*n + k] = sum_j V[j + k n] R[j + c n]: the projections V'R the REML search and the reduced model start from
extern "C" int gapit_eigsys_project(gapit_ctx* ctx, gapit_eigsys* e, const double* R, int nc, double* out) {
  if (!ctx || !e || !R || !out || nc < 1) return fail(GAPIT_ERR_ARG, "gapit_eigsys_project: bad argument");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  const int64_t n = e->n;
  double *dR = nullptr, *dP = nullptr;
  GAPIT_CHECK(ctx_ws_t(ctx, WS_R, size_t(n) * nc, &dR));
  GAPIT_CHECK(ctx_ws_t(ctx, WS_PROJ, size_t(n) * nc, &dP));
  GAPIT_CUDA(cudaMemcpyAsync(dR, R, size_t(n) * nc * 8, cudaMemcpyHostToDevice, ctx->stream));
  GAPIT_CHECK(launch_vt_project(ctx->stream, e->V, n, dR, nc, dP));
  GAPIT_CUDA(cudaMemcpyAsync(out, dP, size_t(n) * nc * 8, cudaMemcpyDeviceToHost, ctx->stream));
  GAPIT_CUDA(cudaStreamSynchronize(ctx->stream));
  return GAPIT_OK;
}

extern "C" const double* gapit_eigsys_vectors_dev(const gapit_eigsys* e) { return e ? e->V : nullptr; }

extern "C" void gapit_eigsys_free(gapit_eigsys* e) {
  if (!e) return;
  cudaSetDevice(e->device);   // never through e->ctx: host languages may destroy the context first
  cudaFree(e->V);
  cudaFree(e->Vs);
  cudaFree(e->scale);
  delete e;
}

extern "C" int gapit_p3d_scan_dev(gapit_ctx* ctx, gapit_geno* g, gapit_eigsys* e, const double* X0, int q0, const double* Y,
                                  int T, const double* delta, const double* vg, int glm, gapit_scan_out* out,
                                  gapit_scan_base* base_out) {
  return scan_impl(ctx, g, e, X0, q0, Y, T, delta, vg, glm, out, base_out);
}

extern "C" int gapit_p3d_scan_host(gapit_ctx* ctx, const void* snps, gapit_dtype dtype, int64_t n, int64_t m, int64_t ld,
                                   gapit_impute impute, gapit_eigsys* e, const double* X0, int q0, const double* Y, int T,
                                   const double* delta, const double* vg, int glm, gapit_scan_out* out,
                                   gapit_scan_base* base_out) {
  if (!ctx) return fail(GAPIT_ERR_ARG, "gapit_p3d_scan_host: NULL ctx");
  return scan_host_impl(ctx, snps, dtype, n, m, ld, impute, e, X0, q0, Y, T, delta, vg, glm, out, base_out);
}

extern "C" int gapit_p3d_scan(gapit_ctx* ctx, gapit_geno* g, const double* vectors, const double* values, const double* X0,
                              int q0, const double* Y, int T, const double* delta, const double* vg, int glm,
                              gapit_scan_out* out, gapit_scan_base* base_out) {
  if (glm) return scan_impl(ctx, g, nullptr, X0, q0, Y, T, delta, vg, 1, out, base_out);
  if (!g) return fail(GAPIT_ERR_ARG, "gapit_p3d_scan: NULL genotype handle");
  gapit_eigsys* e = nullptr;
  GAPIT_CHECK(gapit_eigsys_upload(ctx, vectors, values, g->n, &e));
  const int rc = scan_impl(ctx, g, e, X0, q0, Y, T, delta, vg, 0, out, base_out);
  gapit_eigsys_free(e);
  return rc;
}

// =============================================================================== i = 0 BLUP / PEV block
// GAPIT.EMMAxP3D.R:779-839 in the eigenbasis of K.  With H = K + delta I, r = y - X0 beta0 and
// B = K H^-1 = V diag(lambda/(lambda+delta)) V':
//   BLUP = K U U'(y - X beta)            = B r                                   (:788-803)
//   C11  = vg solve(Xt'Xt)               = vg (X0' H^-1 X0)^-1                   (:824)
//   C21  = -K Zt'Xt C11                  = -(B X0) C11                           (:827)
//   term.1 = solve(I/ve + Kinv/vg)       = V diag(1/(1/ve + kappa_k/vg)) V'      (:828-835)
//   term.2 = C21 Xt'Zt K                 = C21 (B X0)'                           (:837)
//   PEV_i  = diag(term.1 - term.2)_i     = sum_k V_ik^2 g_k + (BX0)_i' C11 (BX0)_i   (:838-839)
// kappa_k = 1/lambda_k when solve(K) succeeds; when K is numerically singular (the usual case: centring by 2p
// makes K 1 = 0) the reference falls back to MASS::ginv(K), i.e. kappa_k = 0 for |lambda_k| <= sqrt(eps) max|lambda|.
// The O(n^3) inversions of the reference become O(n^2 q) products with V.
extern "C" int gapit_blup_pev(gapit_ctx* ctx, gapit_eigsys* e, const double* X0, int q0, const double* y, double delta,
                              double vg, double* blup_out, double* pev_out, double* beta_out) {
  if (!ctx || !e || !X0 || !y || !blup_out || !pev_out || q0 < 1 || q0 > GAPIT_MAX_Q0)
    return fail(GAPIT_ERR_ARG, "gapit_blup_pev: bad argument");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  const int64_t n = e->n;
  const double ve = delta * vg;
  const int nc = q0 + 1;
  // projections V'[X0 | y]
  double *dR = nullptr, *dProj = nullptr;
  GAPIT_CHECK(ctx_ws_t(ctx, WS_R, size_t(n) * (nc + 2), &dR));
  GAPIT_CHECK(ctx_ws_t(ctx, WS_PROJ, size_t(n) * (nc + 2), &dProj));
  GAPIT_CUDA(cudaMemcpyAsync(dR, X0, size_t(n) * q0 * 8, cudaMemcpyHostToDevice, ctx->stream));
  GAPIT_CUDA(cudaMemcpyAsync(dR + size_t(n) * q0, y, size_t(n) * 8, cudaMemcpyHostToDevice, ctx->stream));
  GAPIT_CHECK(launch_vt_project(ctx->stream, e->V, n, dR, nc, dProj));
  std::vector<double> proj(size_t(n) * nc);
  GAPIT_CUDA(cudaMemcpyAsync(proj.data(), dProj, size_t(n) * nc * 8, cudaMemcpyDeviceToHost, ctx->stream));
  GAPIT_CUDA(cudaStreamSynchronize(ctx->stream));
  // reduced model in the eigenbasis (same arithmetic as scan_prepare)
  std::vector<double> w(n), XX(q0 * q0, 0.0), XY(q0, 0.0), inv(q0 * q0), beta(q0), C11(q0 * q0);
  const double* yt = proj.data() + size_t(q0) * n;
  for (int64_t k = 0; k < n; ++k) w[k] = 1.0 / (e->values[k] + delta);
  for (int i = 0; i < q0; ++i) {
    const double* xi = proj.data() + size_t(i) * n;
    for (int j = 0; j <= i; ++j) {
      const double* xj = proj.data() + size_t(j) * n;
      double s = 0.0;
      for (int64_t k = 0; k < n; ++k) s += w[k] * xi[k] * xj[k];
      XX[i * q0 + j] = XX[j * q0 + i] = s;
    }
    double s = 0.0;
    for (int64_t k = 0; k < n; ++k) s += w[k] * xi[k] * yt[k];
    XY[i] = s;
  }
  int singular = 0;
  sym_inverse(q0, XX.data(), inv.data(), &singular);
  for (int i = 0; i < q0; ++i) {
    double s = 0.0;
    for (int j = 0; j < q0; ++j) {
      s += inv[i * q0 + j] * XY[j];
      C11[i * q0 + j] = vg * inv[i * q0 + j];
    }
    beta[i] = s;
    if (beta_out) beta_out[i] = s;
  }
  // Z columns (eigenbasis): f o V'X0 (q0 columns), f o V'r, and the term.1 spectrum g
  double lmax = 0.0, lmin = INFINITY;
  for (int64_t k = 0; k < n; ++k) {
    lmax = std::max(lmax, std::fabs(e->values[k]));
    lmin = std::min(lmin, std::fabs(e->values[k]));
  }
  const bool use_ginv = !(lmin / lmax >= 2.220446049250313e-16);   // solve(K) refuses rcond < eps (:828-829)
  const double tol = std::sqrt(2.220446049250313e-16) * lmax;       // MASS::ginv
  std::vector<double> Z(size_t(n) * (nc + 1));
  for (int64_t k = 0; k < n; ++k) {
    const double lam = e->values[k];
    const double f = lam * w[k];
    double rk = yt[k];
    for (int j = 0; j < q0; ++j) {
      Z[size_t(j) * n + k] = f * proj[size_t(j) * n + k];
      rk -= beta[j] * proj[size_t(j) * n + k];
    }
    Z[size_t(q0) * n + k] = f * rk;
    const double kappa = use_ginv ? (std::fabs(lam) > tol ? 1.0 / lam : 0.0) : 1.0 / lam;
    Z[size_t(q0 + 1) * n + k] = 1.0 / (1.0 / ve + kappa / vg);
  }
  double* dZ = dR;       // reuse
  double* dOut = dProj;  // reuse
  GAPIT_CUDA(cudaMemcpyAsync(dZ, Z.data(), Z.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  GAPIT_CHECK(launch_v_times_cols(ctx->stream, e->V, n, dZ, nc, dOut, false));
  GAPIT_CHECK(launch_v_times_cols(ctx->stream, e->V, n, dZ + size_t(nc) * n, 1, dOut + size_t(nc) * n, true));
  std::vector<double> O(size_t(n) * (nc + 1));
  GAPIT_CUDA(cudaMemcpyAsync(O.data(), dOut, O.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  GAPIT_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int64_t i = 0; i < n; ++i) {
    blup_out[i] = O[size_t(q0) * n + i];
    double quad = 0.0;
    for (int a = 0; a < q0; ++a)
      for (int b = 0; b < q0; ++b) quad += O[size_t(a) * n + i] * C11[a * q0 + b] * O[size_t(b) * n + i];
    pev_out[i] = O[size_t(q0 + 1) * n + i] + quad;
  }
  return GAPIT_OK;
}
