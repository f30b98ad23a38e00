// fp64 contractions around the rotation of the multi-trait P3D scan (BASELINE configs[4]: many traits share one
// eigendecomposition, GAPIT.R:101-132 looping GAPIT.EMMAxP3D.R:634-683 per trait).
//
//   xx_contract   XX[t][i] = sum_k Wt[k][t] * W2[k][i]      the trait-specific quadratic forms  x_i' H_t^-1 x_i  (:655)
//                 from the squared rotated markers W2 = (V'x_i)_k^2 the rotation kernel left in HBM: an
//                 [markers x n] x [n x traits] GEMM on the fp64 tensor pipe (DMMA.8x8x4), cp.async-staged.
//   v_times_cols  out = V Z (n x n times n x nc), the back-rotation H_t^-1 [X0 | r_t] = V diag(w_t) V' [X0 | r_t]
//                 and the BLUP / PEV products; n^2 nc work, off the marker path.
#include "common.cuh"

namespace gapit {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// D (8 x 8) += A (8 x 4, row) * B (4 x 8, col): lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2(l%4) + {0,1}]
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

constexpr int kXxKT = 16;      // contracted rows per ring stage
constexpr int kXxStages = 3;
constexpr int kXxAS = 132;     // padded row of the marker tile: stride = 4 (mod 16) doubles => conflict-free fragment reads
__host__ __device__ constexpr int xx_tns(int tn) { return tn + ((4 - tn % 16) + 16) % 16; }
__host__ __device__ constexpr int xx_smem_bytes(int nt8) {
  return kXxStages * kXxKT * (kXxAS + xx_tns(8 * nt8)) * 8;
}

// CTA = 128 markers x (8 NT8) traits; warp w owns markers [16 w, 16 w + 16) as two 8-row DMMA tiles.
// W2: [nk][m_ld] (markers contiguous), Wt: [nk][wt_ld] (traits contiguous, zero padded to the grid's trait tiles),
// XX: [traits][xx_ld].  Summation order over k is fixed (k ascending in chunks of 4) => bit-reproducible.
template <int NT8>
__global__ void __launch_bounds__(256, (NT8 <= 8) ? 2 : 1)
    xx_contract_kernel(const double* __restrict__ W2, int64_t m_ld, const double* __restrict__ Wt, int wt_ld, int nk,
                       double* __restrict__ XX, int64_t xx_ld, int T) {
  constexpr int TN = 8 * NT8;
  constexpr int TNS = xx_tns(TN);
  constexpr int AS = kXxAS, KT = kXxKT, STAGES = kXxStages;
  extern __shared__ double xsm[];
  double* As = xsm;
  double* Bs = xsm + STAGES * KT * AS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t i0 = int64_t(blockIdx.x) * 128;
  const int tb = blockIdx.y * TN;
  const int nkt = (nk + KT - 1) / KT;

  auto load_stage = [&](int st, int kt) {
    const int k0 = kt * KT;
    for (int c = tid; c < KT * 64; c += 256) {
      const int r = c >> 6, cc = c & 63;
      const int k = k0 + r;
      cp_async16(&As[(st * KT + r) * AS + cc * 2], W2 + int64_t(min(k, nk - 1)) * m_ld + i0 + cc * 2, k < nk ? 16 : 0);
    }
    for (int c = tid; c < KT * (TN / 2); c += 256) {
      const int r = c / (TN / 2), cc = c - r * (TN / 2);
      const int k = k0 + r;
      cp_async16(&Bs[(st * KT + r) * TNS + cc * 2], Wt + int64_t(min(k, nk - 1)) * wt_ld + tb + cc * 2, k < nk ? 16 : 0);
    }
  };

  double acc[2][NT8][2];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int j = 0; j < NT8; ++j) acc[h][j][0] = acc[h][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nkt) load_stage(s, s);
    cp_async_commit();
  }
  const int ar = lane & 3, ac = lane >> 2;
  for (int kt = 0; kt < nkt; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (kt + STAGES - 1 < nkt) load_stage((kt + STAGES - 1) % STAGES, kt + STAGES - 1);
    cp_async_commit();
    const double* a_st = As + (kt % STAGES) * KT * AS + warp * 16 + ac;
    const double* b_st = Bs + (kt % STAGES) * KT * TNS + ac;
#pragma unroll
    for (int k4 = 0; k4 < KT / 4; ++k4) {
      const double a0 = a_st[(k4 * 4 + ar) * AS];
      const double a1 = a_st[(k4 * 4 + ar) * AS + 8];
      double b[NT8];
#pragma unroll
      for (int j = 0; j < NT8; ++j) b[j] = b_st[(k4 * 4 + ar) * TNS + j * 8];
#pragma unroll
      for (int j = 0; j < NT8; ++j) {
        dmma884(acc[0][j][0], acc[0][j][1], a0, b[j]);
        dmma884(acc[1][j][0], acc[1][j][1], a1, b[j]);
      }
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int64_t mrk = i0 + warp * 16 + h * 8 + ac;
#pragma unroll
    for (int j = 0; j < NT8; ++j) {
      const int t = tb + j * 8 + 2 * ar;
      if (t < T) XX[int64_t(t) * xx_ld + mrk] = acc[h][j][0];
      if (t + 1 < T) XX[int64_t(t + 1) * xx_ld + mrk] = acc[h][j][1];
    }
  }
}

template <int NT8>
static int launch_xx_nt(cudaStream_t st, const double* W2, int64_t m_ld, const double* Wt, int wt_ld, int nk, double* XX,
                        int64_t xx_ld, int T, int ytiles) {
  auto kern = xx_contract_kernel<NT8>;
  constexpr int smem = xx_smem_bytes(NT8);
  GAPIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid(static_cast<unsigned>(m_ld / 128), static_cast<unsigned>(ytiles));
  kern<<<grid, 256, smem, st>>>(W2, m_ld, Wt, wt_ld, nk, XX, xx_ld, T);
  GAPIT_CUDA(cudaGetLastError());
  return GAPIT_OK;
}

// trait tile width (multiple of 8) for T traits: one tile when T <= 128, else tiles of 128
int xx_trait_tile(int T) {
  const int t8 = (T + 7) / 8;
  if (t8 <= 1) return 8;
  if (t8 <= 2) return 16;
  if (t8 <= 4) return 32;
  if (t8 <= 8) return 64;
  if (t8 <= 13) return 104;
  return 128;
}

// m_ld must be a multiple of 128; Wt rows hold ceil(T / tile) * tile doubles (wt_ld)
int launch_xx_contract(cudaStream_t st, const double* W2, int64_t m_ld, const double* Wt, int wt_ld, int nk, double* XX,
                       int64_t xx_ld, int T) {
  const int tile = xx_trait_tile(T);
  const int ytiles = (T + tile - 1) / tile;
  if (wt_ld < ytiles * tile || m_ld % 128 != 0) return fail(GAPIT_ERR_ARG, "launch_xx_contract: bad leading dimensions");
  switch (tile) {
    case 8: return launch_xx_nt<1>(st, W2, m_ld, Wt, wt_ld, nk, XX, xx_ld, T, ytiles);
    case 16: return launch_xx_nt<2>(st, W2, m_ld, Wt, wt_ld, nk, XX, xx_ld, T, ytiles);
    case 32: return launch_xx_nt<4>(st, W2, m_ld, Wt, wt_ld, nk, XX, xx_ld, T, ytiles);
    case 64: return launch_xx_nt<8>(st, W2, m_ld, Wt, wt_ld, nk, XX, xx_ld, T, ytiles);
    case 104: return launch_xx_nt<13>(st, W2, m_ld, Wt, wt_ld, nk, XX, xx_ld, T, ytiles);
    default: return launch_xx_nt<16>(st, W2, m_ld, Wt, wt_ld, nk, XX, xx_ld, T, ytiles);
  }
}

// ------------------------------------------------------------------ V'R
// out[c*n + k] = sum_j V[j + k n] R[j + c n]  (the projections V'[X0 | Y] the reduced model, REML and the back-rotation
// start from): the same DMMA tiling with both operands contiguous along the contracted index j.  CTA = 128 eigenvectors
// x (8 NT8) columns, ring stages of 16 taxa; 8-byte cp.async so that odd n needs no alignment.  j ascending in chunks of
// 4 => fixed summation order.
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, int src_bytes) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}

constexpr int kVtLd = 20;   // padded row of 16 taxa: stride = 4 (mod 16) doubles => conflict-free fragment reads
__host__ __device__ constexpr int vt_smem_bytes(int nt8) { return kXxStages * (128 + 8 * nt8) * kVtLd * 8; }

template <int NT8>
__global__ void __launch_bounds__(256, (NT8 <= 8) ? 2 : 1)
    vt_project_kernel(const double* __restrict__ V, int64_t n, const double* __restrict__ R, int nc, double* __restrict__ out) {
  constexpr int TN = 8 * NT8, KT = kXxKT, STAGES = kXxStages, LD = kVtLd;
  extern __shared__ double xsm[];
  double* As = xsm;                       // [STAGES][128][LD]
  double* Bs = xsm + STAGES * 128 * LD;   // [STAGES][TN][LD]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t k0 = int64_t(blockIdx.x) * 128;
  const int cb = blockIdx.y * TN;
  const int nkt = static_cast<int>((n + KT - 1) / KT);

  auto load_stage = [&](int st, int kt) {
    const int64_t j0 = int64_t(kt) * KT;
    for (int c = tid; c < 128 * KT; c += 256) {
      const int r = c >> 4, jj = c & 15;
      const bool ok = (k0 + r < n) && (j0 + jj < n);
      cp_async8(&As[(st * 128 + r) * LD + jj], V + (ok ? (k0 + r) * n + j0 + jj : 0), ok ? 8 : 0);
    }
    for (int c = tid; c < TN * KT; c += 256) {
      const int r = c >> 4, jj = c & 15;
      const bool ok = (cb + r < nc) && (j0 + jj < n);
      cp_async8(&Bs[(st * TN + r) * LD + jj], R + (ok ? int64_t(cb + r) * n + j0 + jj : 0), ok ? 8 : 0);
    }
  };

  double acc[2][NT8][2];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int j = 0; j < NT8; ++j) acc[h][j][0] = acc[h][j][1] = 0.0;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nkt) load_stage(s, s);
    cp_async_commit();
  }
  const int ar = lane & 3, ac = lane >> 2;
  for (int kt = 0; kt < nkt; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (kt + STAGES - 1 < nkt) load_stage((kt + STAGES - 1) % STAGES, kt + STAGES - 1);
    cp_async_commit();
    const double* a_st = As + ((kt % STAGES) * 128 + warp * 16 + ac) * LD + ar;
    const double* b_st = Bs + ((kt % STAGES) * TN + ac) * LD + ar;
#pragma unroll
    for (int k4 = 0; k4 < KT / 4; ++k4) {
      const double a0 = a_st[k4 * 4];
      const double a1 = a_st[8 * LD + k4 * 4];
      double b[NT8];
#pragma unroll
      for (int j = 0; j < NT8; ++j) b[j] = b_st[j * 8 * LD + k4 * 4];
#pragma unroll
      for (int j = 0; j < NT8; ++j) {
        dmma884(acc[0][j][0], acc[0][j][1], a0, b[j]);
        dmma884(acc[1][j][0], acc[1][j][1], a1, b[j]);
      }
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int64_t k = k0 + warp * 16 + h * 8 + ac;
    if (k >= n) continue;
#pragma unroll
    for (int j = 0; j < NT8; ++j) {
      const int c = cb + j * 8 + 2 * ar;
      if (c < nc) out[int64_t(c) * n + k] = acc[h][j][0];
      if (c + 1 < nc) out[int64_t(c + 1) * n + k] = acc[h][j][1];
    }
  }
}

template <int NT8>
static int launch_vt_nt(cudaStream_t st, const double* V, int64_t n, const double* R, int nc, double* out, int ytiles) {
  auto kern = vt_project_kernel<NT8>;
  constexpr int smem = vt_smem_bytes(NT8);
  GAPIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid(static_cast<unsigned>((n + 127) / 128), static_cast<unsigned>(ytiles));
  kern<<<grid, 256, smem, st>>>(V, n, R, nc, out);
  GAPIT_CUDA(cudaGetLastError());
  return GAPIT_OK;
}

// V (n x n column-major), R (n x nc column-major) -> out (n x nc column-major) = V'R; all device pointers
int launch_vt_project(cudaStream_t st, const double* V, int64_t n, const double* R, int nc, double* out) {
  const int tile = xx_trait_tile(nc);
  const int ytiles = (nc + tile - 1) / tile;
  switch (tile) {
    case 8: return launch_vt_nt<1>(st, V, n, R, nc, out, ytiles);
    case 16: return launch_vt_nt<2>(st, V, n, R, nc, out, ytiles);
    case 32: return launch_vt_nt<4>(st, V, n, R, nc, out, ytiles);
    case 64: return launch_vt_nt<8>(st, V, n, R, nc, out, ytiles);
    case 104: return launch_vt_nt<13>(st, V, n, R, nc, out, ytiles);
    default: return launch_vt_nt<16>(st, V, n, R, nc, out, ytiles);
  }
}

// ------------------------------------------------------------------ V Z
// out[j + c n] = sum_k (SQUARE ? V_jk^2 : V_jk) Z[k + c n]: one thread per row j (coalesced along j), NC columns per
// thread so V is streamed once per NC columns, Z chunks broadcast from shared memory; k ascending => fixed order.
template <int NC, bool SQUARE>
__global__ void __launch_bounds__(128) v_times_cols_kernel(const double* __restrict__ V, int64_t n,
                                                           const double* __restrict__ Z, int nc, double* __restrict__ out) {
  constexpr int KC = 64;
  __shared__ double zs[KC][NC];
  const int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int c0 = blockIdx.y * NC;
  double acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) acc[c] = 0.0;
  for (int64_t k0 = 0; k0 < n; k0 += KC) {
    __syncthreads();
    for (int i = threadIdx.x; i < KC * NC; i += blockDim.x) {
      const int kk = i % KC, c = i / KC;
      zs[kk][c] = (k0 + kk < n && c0 + c < nc) ? __ldg(Z + int64_t(c0 + c) * n + k0 + kk) : 0.0;
    }
    __syncthreads();
    if (j < n) {
      const int kend = (n - k0 < KC) ? static_cast<int>(n - k0) : KC;
      for (int kk = 0; kk < kend; ++kk) {
        double v = V[j + (k0 + kk) * n];
        if (SQUARE) v *= v;
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = fma(v, zs[kk][c], acc[c]);
      }
    }
  }
  if (j < n) {
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (c0 + c < nc) out[int64_t(c0 + c) * n + j] = acc[c];
  }
}

int launch_v_times_cols(cudaStream_t st, const double* V, int64_t n, const double* Z, int nc, double* out, bool square) {
  dim3 grid(static_cast<unsigned>((n + 127) / 128), static_cast<unsigned>((nc + 7) / 8));
  if (square) v_times_cols_kernel<8, true><<<grid, 128, 0, st>>>(V, n, Z, nc, out);
  else v_times_cols_kernel<8, false><<<grid, 128, 0, st>>>(V, n, Z, nc, out);
  GAPIT_CUDA(cudaGetLastError());
  return GAPIT_OK;
}

}  // namespace gapit
