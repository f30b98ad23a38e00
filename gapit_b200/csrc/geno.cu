// Genotype ingestion: host matrix (R column-major n x m, f64 / i32 / i8) ->
// device int8 marker-major [m][ldn] with per-marker sum and sum of squares,
// All three kernels are HBM-bound byte shuffles (128-bit loads, coalesced rows).
#include <algorithm>
#include <climits>
#include <vector>

#include "common.cuh"

namespace gapit {

// ---- f64 / i32 -> int8 with validation ------------------------------------
// src: chunk of `mc` markers, element (i,k) at src[i + k*ld]; dst row k at dst + k*ldn (128-byte aligned rows).
// One thread converts four consecutive elements and writes one 32-bit word, so a warp reads 1 KB (f64) / 512 B
// (i32) of contiguous input and writes 128 contiguous bytes; the row padding up to ldn is written as zeros.
template <typename T>
__global__ void __launch_bounds__(256) convert_kernel(const T* __restrict__ src, int64_t ld, int64_t n, int64_t mc,
                                                      int8_t* __restrict__ dst, int64_t ldn, int impute,
                                                      int* __restrict__ bad) {
  const int64_t words = ldn >> 2;
  const bool vec = ((reinterpret_cast<uintptr_t>(src) | static_cast<uintptr_t>(ld * sizeof(T))) & 15) == 0;
  int flag = 0;
  for (int64_t k = blockIdx.y; k < mc; k += gridDim.y) {
    const T* s = src + k * ld;
    uint32_t* d = reinterpret_cast<uint32_t*>(dst + k * ldn);
    for (int64_t w = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; w < words; w += int64_t(gridDim.x) * blockDim.x) {
      // classification: code 0/1/2, 3 = NA, 4 = not a dosage
      uint32_t code[4] = {0u, 0u, 0u, 0u};
      const int64_t i0 = w * 4;
      if (sizeof(T) == 4) {
        int q[4] = {0, 0, 0, 0};
        if (vec && i0 + 3 < n) {   // one 16-byte load when every row start is 16-byte aligned
          const int4 v = __ldg(reinterpret_cast<const int4*>(s + i0));
          q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w;
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (i0 + c < n) q[c] = static_cast<int>(s[i0 + c]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)   // R's NA_integer_ is INT_MIN
          code[c] = static_cast<unsigned>(q[c]) <= 2u ? static_cast<uint32_t>(q[c]) : (q[c] == INT_MIN ? 3u : 4u);
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (i0 + c < n) {
            const double v = static_cast<double>(s[i0 + c]);
            const int q = __double2int_rn(v);
            code[c] = (v != v) ? 3u                                                   // NaN / NA_real_ -> NA
                      : (static_cast<double>(q) == v && static_cast<unsigned>(q) <= 2u) ? static_cast<uint32_t>(q) : 4u;
          }
        }
      }
      uint32_t o = 0u;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v = code[c];
        if (v == 3u) {
          if (impute < 0) flag |= 2;
          v = impute < 0 ? 0u : static_cast<uint32_t>(impute);
        } else if (v == 4u) {
          flag |= 1;
          v = 0u;
        }
        o |= v << (8 * c);
      }
      d[w] = o;
    }
  }
  if (flag) atomicOr(bad, flag);
}

// ---- 2-bit packed -> int8 ---------------------------------------------------
// src: `mc` dense rows of nb = ceil(n/4) bytes (rows are NOT word aligned in general); taxon i of a marker sits in
// byte i/4 at bits 2(i%4)..2(i%4)+1 (PLINK .bed bit order), code 0/1/2 = dosage, 3 = NA.  One thread expands 4 packed
// bytes (fetched as two aligned words and a funnel shift) into 16 output bytes with bit spreads and writes them as
// one 128-bit store; the row padding up to ldn is written as zeros.  The staging buffer carries 8 bytes of slack.
__device__ __forceinline__ uint32_t spread2(uint32_t b) {   // 4 x 2-bit fields of a byte -> 4 bytes
  uint32_t v = (b | (b << 12)) & 0x000F000Fu;
  return (v | (v << 6)) & 0x03030303u;
}

// recode: 0 = GAPIT_B2 codes (0/1/2 dosage, 3 NA); 1 / 2 = the codes of a PLINK .bed row (00 hom A1, 01 missing, 10 het,
// 11 hom A2) counted as copies of A1 / of A2.  With (h, l) the two bits of a code the B2 code is
//   copies of A2: (l, h ^ l)      copies of A1: (~h, h ^ l)        (01 -> 11 = NA either way)
// applied to all 16 codes of a word at once.
__global__ void __launch_bounds__(256) unpack2_kernel(const uint8_t* __restrict__ src, int64_t nb, int64_t n, int64_t mc,
                                                      int8_t* __restrict__ dst, int64_t ldn, int impute, int recode,
                                                      int* __restrict__ bad) {
  const int groups = static_cast<int>(ldn >> 4);  // ldn is a multiple of 128
  const uint32_t na = impute < 0 ? 0u : static_cast<uint32_t>(impute);
  const uint32_t* src32 = reinterpret_cast<const uint32_t*>(src);
  int flag = 0;
  for (int64_t k = blockIdx.y; k < mc; k += gridDim.y) {
    const int64_t row_off = k * nb;
    uint4* drow = reinterpret_cast<uint4*>(dst + k * ldn);
    for (int gi = blockIdx.x * blockDim.x + threadIdx.x; gi < groups; gi += gridDim.x * blockDim.x) {
      const int64_t byte0 = int64_t(gi) * 4;
      uint32_t pk = 0u;
      if (byte0 < nb) {
        const int64_t off = row_off + byte0;
        const uint32_t lo = __ldg(src32 + (off >> 2));
        const uint32_t hi = __ldg(src32 + (off >> 2) + 1);
        pk = __funnelshift_r(lo, hi, static_cast<uint32_t>(off & 3) * 8u);
        if (recode) {
          const uint32_t l = pk & 0x55555555u, h = (pk >> 1) & 0x55555555u;
          pk = ((recode == 1 ? (~h & 0x55555555u) : l) << 1) | (h ^ l);
        }
        const int64_t left = n - byte0 * 4;              // genotypes of this row from byte0 on
        if (left < 16) pk &= (1u << (2 * static_cast<int>(left))) - 1u;   // bits past the last taxon are ignored
      }
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t v = spread2((pk >> (8 * j)) & 0xFFu);
        const uint32_t m3 = v & (v >> 1) & 0x01010101u;   // bytes holding code 3
        if (m3) {
          if (impute < 0) flag = 2;
          v = (v & ~(m3 * 3u)) | (m3 * na);
        }
        w[j] = v;
      }
      drow[gi] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  if (flag) atomicOr(bad, flag);
}

// ---- per-marker statistics: one warp per marker, 16-byte loads -------------
__global__ void colstats_kernel(const int8_t* __restrict__ mk, int64_t ldn, int64_t n, int64_t m,
                                int32_t* __restrict__ colsum, int32_t* __restrict__ colsumsq, int* __restrict__ bad) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t k = warp; k < m; k += nwarps) {
    const uint4* row = reinterpret_cast<const uint4*>(mk + k * ldn);
    int s1 = 0, s2 = 0, flag = 0;
    for (int64_t c = lane; c * 16 < n; c += 32) {
      uint4 v = __ldg(row + c);
      uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // bytes are 0,1,2 (padding bytes beyond n are zero): sum and sum of squares by dp4a
        s1 = __dp4a(static_cast<int>(w[j]), 0x01010101, s1);
        s2 = __dp4a(static_cast<int>(w[j]), static_cast<int>(w[j]), s2);
        flag |= (w[j] & 0xFCFCFCFCu) | ((w[j] >> 1) & w[j] & 0x01010101u);  // any byte > 2
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      flag |= __shfl_xor_sync(0xffffffffu, flag, o);
    }
    if (lane == 0) {
      colsum[k] = s1;
      colsumsq[k] = s2;
      if (flag) atomicOr(bad, 1);
    }
  }
}

// asynchronous pieces shared with the streaming scan (scan.cu)
int launch_colstats(gapit_ctx* ctx, cudaStream_t st, const int8_t* mk, int64_t ldn, int64_t n, int64_t m, int32_t* colsum,
                    int32_t* colsumsq, int* bad_dev) {
  const int threads = 256;
  const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((m * 32 + threads - 1) / threads, int64_t(ctx->sm_count) * 16)));
  colstats_kernel<<<blocks, threads, 0, st>>>(mk, ldn, n, m, colsum, colsumsq, bad_dev);
  GAPIT_CUDA(cudaGetLastError());
  return GAPIT_OK;
}

// stage: `mc` markers of n elements (dense, ld = n) already on the device
int launch_convert(cudaStream_t st, gapit_dtype dtype, const void* stage, int64_t n, int64_t mc, int8_t* dst, int64_t ldn,
                   int impute, int* bad_dev) {
  const int64_t words = ldn / 4;
  const int threads = words >= 256 ? 256 : 64;
  dim3 grid(static_cast<unsigned>((words + threads - 1) / threads), static_cast<unsigned>(std::min<int64_t>(mc, 65535)));
  if (dtype == GAPIT_F64)
    convert_kernel<double><<<grid, threads, 0, st>>>(static_cast<const double*>(stage), n, n, mc, dst, ldn, impute, bad_dev);
  else
    convert_kernel<int32_t><<<grid, threads, 0, st>>>(static_cast<const int32_t*>(stage), n, n, mc, dst, ldn, impute, bad_dev);
  GAPIT_CUDA(cudaGetLastError());
  return GAPIT_OK;
}

// stage: `mc` dense rows of ceil(n/4) packed bytes already on the device
int launch_unpack2(gapit_ctx* ctx, cudaStream_t st, const void* stage, int64_t n, int64_t mc, int8_t* dst, int64_t ldn,
                   int impute, int* bad_dev, int recode) {
  const int64_t groups = ldn / 16;
  int threads = 64;   // CTA width with the fewest idle threads in a row's last CTA
  for (int t : {128, 256})
    if ((groups + t - 1) / t * t <= (groups + threads - 1) / threads * threads) threads = t;
  dim3 grid(static_cast<unsigned>((groups + threads - 1) / threads), static_cast<unsigned>(std::min<int64_t>(mc, 65535)));
  unpack2_kernel<<<grid, threads, 0, st>>>(static_cast<const uint8_t*>(stage), (n + 3) / 4, n, mc, dst, ldn, impute, recode, bad_dev);
  GAPIT_CUDA(cudaGetLastError());
  return GAPIT_OK;
}

static int run_colstats(gapit_geno* g) {
  gapit_ctx* ctx = g->ctx;
  int* bad = nullptr;
  GAPIT_CHECK(ctx_ws_t(ctx, WS_BAD, 4, &bad));
  GAPIT_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream));
  GAPIT_CHECK(launch_colstats(ctx, ctx->stream, g->mk, g->ldn, g->n, g->m, g->colsum, g->colsumsq, bad));
  int hbad = 0;
  GAPIT_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  GAPIT_CUDA(cudaStreamSynchronize(ctx->stream));
  if (hbad) return fail(GAPIT_ERR_DATA, "genotype matrix holds a value outside {0,1,2}");
  return GAPIT_OK;
}

}  // namespace gapit

using namespace gapit;

extern "C" int gapit_geno_pack(gapit_ctx* ctx, const void* snps, gapit_dtype dtype, int64_t n, int64_t m, int64_t ld,
                               gapit_impute impute, gapit_geno** out) {
  if (!ctx || !snps || !out || n <= 0 || m <= 0 || ld < (is_packed2(dtype) ? (n + 3) / 4 : n))
    return fail(GAPIT_ERR_ARG, "gapit_geno_pack: bad argument");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  gapit_geno* g = new gapit_geno();
  g->ctx = ctx;
  g->device = ctx->device;
  g->n = n;
  g->m = m;
  g->ldn = round_up(n, 128);
  auto cleanup = [&](int rc) {
    gapit_geno_free(g);
    return rc;
  };
  cudaError_t e;
  if ((e = cudaMalloc(&g->mk, size_t(m) * g->ldn)) != cudaSuccess ||
      (e = cudaMalloc(&g->colsum, size_t(m) * 4)) != cudaSuccess ||
      (e = cudaMalloc(&g->colsumsq, size_t(m) * 4)) != cudaSuccess)
    return cleanup(fail(GAPIT_ERR_CUDA, std::string("gapit_geno_pack: cudaMalloc: ") + cudaGetErrorString(e)));

  cudaEventRecord(ctx->ev0, ctx->stream);
  if (dtype == GAPIT_I8) {
    if (g->ldn != n) cudaMemsetAsync(g->mk, 0, size_t(m) * g->ldn, ctx->stream);
    if (ld == n && g->ldn != n) {
      // dense host rows: flat PCIe transfers through a staging buffer, re-pitched on the device
      const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(m, (int64_t(256) << 20) / n));
      void* stage = nullptr;
      int rc2 = ctx_ws(ctx, WS_STAGE0, size_t(chunk) * n, &stage);
      if (rc2 != GAPIT_OK) return cleanup(rc2);
      e = cudaSuccess;
      for (int64_t k0 = 0; k0 < m && e == cudaSuccess; k0 += chunk) {
        const int64_t mc = std::min(chunk, m - k0);
        e = cudaMemcpyAsync(stage, static_cast<const char*>(snps) + size_t(k0) * n, size_t(mc) * n, cudaMemcpyHostToDevice,
                            ctx->stream);
        if (e == cudaSuccess)
          e = cudaMemcpy2DAsync(g->mk + k0 * g->ldn, g->ldn, stage, n, n, mc, cudaMemcpyDeviceToDevice, ctx->stream);
      }
    } else {
      e = cudaMemcpy2DAsync(g->mk, g->ldn, snps, ld, n, m, cudaMemcpyHostToDevice, ctx->stream);
    }
    if (e != cudaSuccess) return cleanup(fail(GAPIT_ERR_CUDA, std::string("H2D genotype copy: ") + cudaGetErrorString(e)));
  } else if (dtype == GAPIT_F64 || dtype == GAPIT_I32) {
    const size_t esz = (dtype == GAPIT_F64) ? 8 : 4;
    cudaMemsetAsync(g->mk, 0, size_t(m) * g->ldn, ctx->stream);
    // stream the wide host matrix through a bounded device staging buffer
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(m, 65535), (int64_t(256) << 20) / int64_t(n * esz)));
    void* stage = nullptr;
    int* bad = nullptr;
    if ((e = cudaMalloc(&stage, size_t(chunk) * n * esz)) != cudaSuccess || (e = cudaMalloc(&bad, 4)) != cudaSuccess)
      return cleanup(fail(GAPIT_ERR_CUDA, std::string("staging cudaMalloc: ") + cudaGetErrorString(e)));
    cudaMemsetAsync(bad, 0, 4, ctx->stream);
    for (int64_t k0 = 0; k0 < m && e == cudaSuccess; k0 += chunk) {
      const int64_t mc = std::min(chunk, m - k0);
      const char* src = static_cast<const char*>(snps) + size_t(k0) * ld * esz;
      e = cudaMemcpy2DAsync(stage, size_t(n) * esz, src, size_t(ld) * esz, size_t(n) * esz, mc, cudaMemcpyHostToDevice,
                            ctx->stream);
      if (e != cudaSuccess) break;
      if (launch_convert(ctx->stream, dtype, stage, n, mc, g->mk + k0 * g->ldn, g->ldn, int(impute), bad) != GAPIT_OK) {
        cudaFree(stage);
        cudaFree(bad);
        return cleanup(GAPIT_ERR_CUDA);
      }
      e = cudaGetLastError();
    }
    int hbad = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&hbad, bad, 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(stage);
    cudaFree(bad);
    if (e != cudaSuccess) return cleanup(fail(GAPIT_ERR_CUDA, std::string("genotype conversion: ") + cudaGetErrorString(e)));
    if (hbad & 2) return cleanup(fail(GAPIT_ERR_DATA, "genotype matrix holds NA and no impute rule was given"));
    if (hbad & 1) return cleanup(fail(GAPIT_ERR_DATA, "genotype matrix holds a value outside {0,1,2}"));
  } else if (is_packed2(dtype)) {
    const int64_t nb = (n + 3) / 4;
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(m, (int64_t(256) << 20) / nb));
    void* stage = nullptr;
    int* bad = nullptr;
    int rc2 = ctx_ws(ctx, WS_STAGE0, size_t(chunk) * nb + 8, &stage);
    if (rc2 == GAPIT_OK) rc2 = ctx_ws_t(ctx, WS_BAD, 4, &bad);
    if (rc2 != GAPIT_OK) return cleanup(rc2);
    e = cudaMemsetAsync(bad, 0, 4, ctx->stream);
    for (int64_t k0 = 0; k0 < m && e == cudaSuccess; k0 += chunk) {
      const int64_t mc = std::min(chunk, m - k0);
      const char* src = static_cast<const char*>(snps) + size_t(k0) * ld;
      if (ld == nb) e = cudaMemcpyAsync(stage, src, size_t(mc) * nb, cudaMemcpyHostToDevice, ctx->stream);
      else e = cudaMemcpy2DAsync(stage, nb, src, ld, nb, mc, cudaMemcpyHostToDevice, ctx->stream);
      if (e == cudaSuccess && launch_unpack2(ctx, ctx->stream, stage, n, mc, g->mk + k0 * g->ldn, g->ldn, int(impute), bad, packed2_recode(dtype)) != GAPIT_OK)
        return cleanup(GAPIT_ERR_CUDA);
    }
    int hbad = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&hbad, bad, 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return cleanup(fail(GAPIT_ERR_CUDA, std::string("packed genotype upload: ") + cudaGetErrorString(e)));
    if (hbad & 2) return cleanup(fail(GAPIT_ERR_DATA, "genotype matrix holds NA and no impute rule was given"));
  } else {
    return cleanup(fail(GAPIT_ERR_ARG, "gapit_geno_pack: unknown dtype"));
  }
  cudaEventRecord(ctx->ev1, ctx->stream);
  const int rc = run_colstats(g);
  if (rc != GAPIT_OK) return cleanup(rc);
  *out = g;
  return GAPIT_OK;
}

extern "C" int gapit_geno_from_device(gapit_ctx* ctx, const int8_t* mk_dev, int64_t n, int64_t m, int64_t ld,
                                      gapit_geno** out) {
  if (!ctx || !mk_dev || !out || n <= 0 || m <= 0 || ld < n) return fail(GAPIT_ERR_ARG, "gapit_geno_from_device: bad argument");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  gapit_geno* g = new gapit_geno();
  g->ctx = ctx;
  g->device = ctx->device;
  g->n = n;
  g->m = m;
  g->ldn = round_up(n, 128);
  cudaError_t e;
  if ((e = cudaMalloc(&g->mk, size_t(m) * g->ldn)) != cudaSuccess || (e = cudaMalloc(&g->colsum, size_t(m) * 4)) != cudaSuccess ||
      (e = cudaMalloc(&g->colsumsq, size_t(m) * 4)) != cudaSuccess) {
    gapit_geno_free(g);
    return fail(GAPIT_ERR_CUDA, std::string("gapit_geno_from_device: cudaMalloc: ") + cudaGetErrorString(e));
  }
  if (g->ldn != n) e = cudaMemsetAsync(g->mk, 0, size_t(m) * g->ldn, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpy2DAsync(g->mk, g->ldn, mk_dev, ld, n, m, cudaMemcpyDeviceToDevice, ctx->stream);
  if (e != cudaSuccess) {
    gapit_geno_free(g);
    return fail(GAPIT_ERR_CUDA, std::string("gapit_geno_from_device: ") + cudaGetErrorString(e));
  }
  const int rc = run_colstats(g);
  if (rc != GAPIT_OK) {
    gapit_geno_free(g);
    return rc;
  }
  *out = g;
  return GAPIT_OK;
}

// Zero-copy form: the caller already holds the dosages in the device layout (marker k at mk_dev + k*ld, ld = n rounded up
// to 128, 128-byte aligned) and keeps ownership; the bytes n..ld of every row are zeroed here, the statistics computed.
extern "C" int gapit_geno_wrap_device(gapit_ctx* ctx, int8_t* mk_dev, int64_t n, int64_t m, int64_t ld, gapit_geno** out) {
  if (!ctx || !mk_dev || !out || n <= 0 || m <= 0 || ld != round_up(n, 128) || (reinterpret_cast<uintptr_t>(mk_dev) & 127))
    return fail(GAPIT_ERR_ARG, "gapit_geno_wrap_device: needs ld = n rounded up to 128 and a 128-byte aligned pointer");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  gapit_geno* g = new gapit_geno();
  g->ctx = ctx;
  g->device = ctx->device;
  g->n = n;
  g->m = m;
  g->ldn = ld;
  g->mk = mk_dev;
  g->borrowed = true;
  cudaError_t e = cudaMalloc(&g->colsum, size_t(m) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&g->colsumsq, size_t(m) * 4);
  if (e == cudaSuccess && ld != n) e = cudaMemset2DAsync(mk_dev + n, size_t(ld), 0, size_t(ld - n), size_t(m), ctx->stream);
  const int rc = (e == cudaSuccess) ? run_colstats(g)
                                    : fail(GAPIT_ERR_CUDA, std::string("gapit_geno_wrap_device: ") + cudaGetErrorString(e));
  if (rc != GAPIT_OK) {
    gapit_geno_free(g);
    return rc;
  }
  *out = g;
  return GAPIT_OK;
}

// out[k][i] = in[k][index[i]]: the `GD[GTindex,]` row selection of the reference's by-file loop
// (GAPIT.EMMAxP3D.R:315), on the marker-major device matrix.  grid (taxa chunks, markers)
__global__ void select_taxa_kernel(const int8_t* __restrict__ src, int64_t ld_src, const int64_t* __restrict__ index,
                                   int64_t n_new, int64_t m, int8_t* __restrict__ dst, int64_t ld_dst) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= ld_dst) return;
  const int64_t from = i < n_new ? index[i] : -1;
  for (int64_t k = blockIdx.y; k < m; k += gridDim.y) dst[k * ld_dst + i] = from >= 0 ? src[k * ld_src + from] : int8_t(0);
}

extern "C" int gapit_geno_select_taxa(gapit_ctx* ctx, const gapit_geno* g, const int64_t* index, int64_t n_new,
                                      gapit_geno** out) {
  if (!ctx || !g || !index || !out || n_new <= 0) return fail(GAPIT_ERR_ARG, "gapit_geno_select_taxa: bad argument");
  for (int64_t i = 0; i < n_new; ++i)
    if (index[i] < 0 || index[i] >= g->n) return fail(GAPIT_ERR_ARG, "gapit_geno_select_taxa: taxon index out of range");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  gapit_geno* h = new gapit_geno();
  h->ctx = ctx;
  h->device = ctx->device;
  h->n = n_new;
  h->m = g->m;
  h->ldn = round_up(n_new, 128);
  int64_t* idx_dev = nullptr;
  cudaError_t e = cudaMalloc(&h->mk, size_t(h->m) * h->ldn);
  if (e == cudaSuccess) e = cudaMalloc(&h->colsum, size_t(h->m) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&h->colsumsq, size_t(h->m) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&idx_dev, size_t(n_new) * 8);
  if (e == cudaSuccess) e = cudaMemcpyAsync(idx_dev, index, size_t(n_new) * 8, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    dim3 grid(static_cast<unsigned>((h->ldn + 255) / 256), static_cast<unsigned>(std::min<int64_t>(h->m, 65535)));
    select_taxa_kernel<<<grid, 256, 0, ctx->stream>>>(g->mk, g->ldn, idx_dev, n_new, h->m, h->mk, h->ldn);
    e = cudaGetLastError();
  }
  int rc = (e == cudaSuccess) ? run_colstats(h) : fail(GAPIT_ERR_CUDA, std::string("gapit_geno_select_taxa: ") + cudaGetErrorString(e));
  cudaStreamSynchronize(ctx->stream);
  cudaFree(idx_dev);
  if (rc != GAPIT_OK) {
    gapit_geno_free(h);
    return rc;
  }
  *out = h;
  return GAPIT_OK;
}

extern "C" void gapit_geno_free(gapit_geno* g) {
  if (!g) return;
  cudaSetDevice(g->device);   // never through g->ctx: host languages may destroy the context first
  if (!g->borrowed) cudaFree(g->mk);
  cudaFree(g->colsum);
  cudaFree(g->colsumsq);
  delete g;
}

extern "C" int64_t gapit_geno_n(const gapit_geno* g) { return g ? g->n : 0; }
extern "C" int64_t gapit_geno_m(const gapit_geno* g) { return g ? g->m : 0; }

extern "C" int gapit_geno_colsums(const gapit_geno* g, int32_t* colsum_out, int32_t* colsumsq_out) {
  if (!g) return fail(GAPIT_ERR_ARG, "gapit_geno_colsums: NULL handle");
  GAPIT_CUDA(cudaSetDevice(g->device));
  if (colsum_out) GAPIT_CUDA(cudaMemcpy(colsum_out, g->colsum, size_t(g->m) * 4, cudaMemcpyDeviceToHost));
  if (colsumsq_out) GAPIT_CUDA(cudaMemcpy(colsumsq_out, g->colsumsq, size_t(g->m) * 4, cudaMemcpyDeviceToHost));
  return GAPIT_OK;
}
