// HapMap ingestion: the per-marker numericalization GAPIT runs before the hot path
// (GAPIT.HapMap.R:19-37 looping GAPIT.Numericalization.R:1-112 over the rows of the table), on the raw text.
//
//   gapit_hapmap_index          host: line starts, the 11th tab of every row, n, m, genotype width (1 or 2 chars)
//   numericalize_kernel         one CTA per marker: distinct genotype strings (<= 3 matter), their counts, the
//                               reference's level ordering / Major.allele.zero / impute / effect rules, then the
//                               0/1/2 dosages straight into the int8 marker-major operand layout of the scan
// Byte work: the text row of a marker is staged once in shared memory with 16-byte loads, the three passes (levels,
// counts, dosages) read it there, and 1 byte per genotype is written as 32-bit words.
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace gapit {

int launch_colstats(gapit_ctx* ctx, cudaStream_t st, const int8_t* mk, int64_t ldn, int64_t n, int64_t m, int32_t* colsum,
                    int32_t* colsumsq, int* bad_dev);

constexpr uint32_t kKeyNA = 0xFFFFFFFFu;
constexpr uint32_t kKeyEmpty = 0xFFFFFFFEu;

// genotype string -> key whose integer order is the byte order of the string; missing codes -> kKeyNA
// (GAPIT.Numericalization.R:8-22)
__device__ __forceinline__ uint32_t read_key(const uint8_t* p, int bit) {
  if (bit == 1) {
    uint32_t c = p[0];
    if (c == 'X' || c == '-' || c == '+' || c == '/' || c == 'N') return kKeyNA;
    if (c == 'K') c = 'Z';
    return c;
  }
  const uint32_t c0 = p[0], c1 = p[1];
  const uint32_t key = (c0 << 8) | c1;
  if (c0 == c1 && (c0 == 'X' || c0 == '-' || c0 == '+' || c0 == '/' || c0 == 'N')) return kKeyNA;
  return key;
}

struct NumParams {
  const uint8_t* text;      // device copy of the byte range [text_base, ...) of the file
  const int64_t* row_off;   // [m] file offset of the first genotype character of each marker
  int64_t text_base;        // file offset of text[0]
  int64_t k0, k1;           // markers of this launch
  int64_t n;
  int bit, stride;
  int effect;               // 0 Add, 1 Dom, 2 Left, 3 Right
  int impute;               // gapit_impute
  int maz;                  // Major.allele.zero
  int8_t* out;              // [m][ldo], NA = -1
  int64_t ldo;
  int* has_na;
};

// stable order of up to three counts (R's order(): ties keep their position)
__device__ void order3(const int* c, int len, bool decreasing, int* idx) {
  for (int i = 0; i < len; ++i) idx[i] = i;
  for (int i = 1; i < len; ++i)
    for (int j = i; j > 0; --j) {
      const bool swap = decreasing ? (c[idx[j]] > c[idx[j - 1]]) : (c[idx[j]] < c[idx[j - 1]]);
      if (!swap) break;
      const int t = idx[j];
      idx[j] = idx[j - 1];
      idx[j - 1] = t;
    }
}

// STAGE: the marker's text row is first copied to shared memory with 16-byte loads (rows up to kStageBytes) and the
// three passes read it from there; otherwise they read global memory (very wide rows).
constexpr int kStageBytes = 48 * 1024;

template <bool STAGE>
__global__ void __launch_bounds__(256) numericalize_kernel(const NumParams p) {
  extern __shared__ uint4 stage_raw[];
  __shared__ uint32_t wmin[8];
  __shared__ uint32_t lev[3];
  __shared__ int len_s;
  __shared__ int cnt[3];
  __shared__ int code[3];   // dosage of sorted level j
  __shared__ int na_val;    // value of a missing genotype after the impute rule (-1 = stays NA)
  const int tid = threadIdx.x;
  for (int64_t k = p.k0 + blockIdx.x; k < p.k1; k += gridDim.x) {
    const uint8_t* grow = p.text + (p.row_off[k] - p.text_base);
    const uint8_t* row = grow;
    if (STAGE) {
      const uintptr_t a = reinterpret_cast<uintptr_t>(grow);
      const int lead = static_cast<int>(a & 15);
      const uint4* g16 = reinterpret_cast<const uint4*>(a - lead);
      const int nvec = static_cast<int>((lead + (p.n - 1) * p.stride + p.bit + 15) >> 4);
      for (int v = tid; v < nvec; v += blockDim.x) stage_raw[v] = __ldg(g16 + v);
      row = reinterpret_cast<const uint8_t*>(stage_raw) + lead;
    }
    if (tid < 3) cnt[tid] = 0;
    __syncthreads();
    // ---- pass 1: the distinct genotype strings of this thread's share (only "more than three" matters beyond three)
    uint32_t k0 = kKeyEmpty, k1 = kKeyEmpty, k2 = kKeyEmpty, k3 = kKeyEmpty;   // registers, no indexed array
    for (int64_t i = tid; i < p.n; i += blockDim.x) {
      const uint32_t key = read_key(row + i * p.stride, p.bit);
      if (key == kKeyNA || key == k0 || key == k1 || key == k2 || key == k3) continue;
      if (k0 == kKeyEmpty) k0 = key;
      else if (k1 == kKeyEmpty) k1 = key;
      else if (k2 == kKeyEmpty) k2 = key;
      else if (k3 == kKeyEmpty) k3 = key;
    }
    // levels(as.factor(x)) (:25): the sorted distinct strings, found as successive block-wide minima (no atomics)
    {
      uint32_t prev = 0u;   // every key is > 0
      int len = 0;
      for (int r = 0; r < 4; ++r) {
        uint32_t cand = kKeyEmpty;
        if (k0 > prev && k0 < cand) cand = k0;
        if (k1 > prev && k1 < cand) cand = k1;
        if (k2 > prev && k2 < cand) cand = k2;
        if (k3 > prev && k3 < cand) cand = k3;
        cand = __reduce_min_sync(0xffffffffu, cand);
        if ((tid & 31) == 0) wmin[tid >> 5] = cand;
        __syncthreads();
        uint32_t mn = kKeyEmpty;
        for (int w = 0; w < (blockDim.x >> 5); ++w) mn = min(mn, wmin[w]);
        __syncthreads();
        if (mn == kKeyEmpty) break;   // uniform across the block
        if (r < 3 && tid == 0) lev[r] = mn;
        ++len;
        prev = mn;
      }
      if (tid == 0) {
        for (int i = len; i < 3; ++i) lev[i] = kKeyEmpty;
        len_s = len;
      }
    }
    __syncthreads();
    const int len = len_s;
    if (len >= 2 && len <= 3) {
      // ---- pass 2: genotype counts (:34-37)
      int c0 = 0, c1 = 0, c2 = 0;
      const uint32_t l0 = lev[0], l1 = lev[1], l2 = lev[2];
      for (int64_t i = tid; i < p.n; i += blockDim.x) {
        const uint32_t key = read_key(row + i * p.stride, p.bit);
        c0 += (key == l0);
        c1 += (key == l1);
        c2 += (key == l2);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        c0 += __shfl_xor_sync(0xffffffffu, c0, o);
        c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        c2 += __shfl_xor_sync(0xffffffffu, c2, o);
      }
      if ((tid & 31) == 0) {
        atomicAdd(&cnt[0], c0);
        atomicAdd(&cnt[1], c1);
        atomicAdd(&cnt[2], c2);
      }
      __syncthreads();
      if (tid == 0) {
        // perm[j] = sorted level that sits at position j of `lev` after the Major.allele.zero reordering (:41-62)
        int perm[3] = {0, 1, 2};
        int count[3] = {cnt[0], cnt[1], cnt[2]};
        if (p.maz) {
          if (len == 2) {
            int idx[2];
            order3(count, 2, true, idx);
            perm[0] = idx[0];
            perm[1] = idx[1];
          } else if (p.bit == 1) {     // rows 1,2 by count, row 3 stays (:45-49)
            int idx[2];
            order3(count, 2, true, idx);
            perm[0] = idx[0];
            perm[1] = idx[1];
            perm[2] = 2;
          } else {                      // rows 1,3 by count, row 2 stays (:53-57)
            const int c13[2] = {count[0], count[2]};
            int idx[2];
            order3(c13, 2, true, idx);
            perm[0] = idx[0] == 0 ? 0 : 2;
            perm[1] = 1;
            perm[2] = idx[1] == 0 ? 0 : 2;
          }
          const int cc[3] = {count[perm[0]], count[perm[1]], count[perm[2]]};
          for (int j = 0; j < 3; ++j) count[j] = cc[j];
        }
        // dosage of the level at position j (:79-86): len 2: 0,2; len 3: 0,1,2 (two characters) or 0,2,1 (one)
        int val[3];
        if (len == 2) {
          val[0] = 0;
          val[1] = 2;
          val[2] = 2;
        } else if (p.bit == 1) {
          val[0] = 0;
          val[1] = 2;
          val[2] = 1;
        } else {
          val[0] = 0;
          val[1] = 1;
          val[2] = 2;
        }
        if (p.bit == 1 && len == 3) {   // :67-71
          const int t = count[1];
          count[1] = count[2];
          count[2] = t;
        }
        int pos[3];
        order3(count, len, false, pos);   // position = order(count) (:72), 0-based here
        int nv = -1;
        if (p.impute == GAPIT_IMPUTE_MIDDLE) nv = 1;                                             // :92
        else if (p.impute == GAPIT_IMPUTE_MINOR) nv = (len == 3) ? pos[0] : 2 * pos[0];          // :95,:99
        else if (p.impute == GAPIT_IMPUTE_MAJOR) nv = (len == 3) ? pos[len - 1] : 2 * pos[len - 1];   // :96,:100
        auto eff = [&](int v) {
          if (v < 0) return v;
          if (p.effect == 1) return v == 1 ? 1 : 0;   // Dom   :105
          if (p.effect == 2) return v == 1 ? 0 : v;   // Left  :106
          if (p.effect == 3) return v == 1 ? 2 : v;   // Right :107
          return v;
        };
        for (int j = 0; j < 3; ++j) code[perm[j]] = eff(val[j]);
        na_val = eff(nv);
      }
      __syncthreads();
    }
    // ---- pass 3: dosages, four per thread, one 32-bit store (rows are padded to ldo, a multiple of 128)
    uint32_t* dst = reinterpret_cast<uint32_t*>(p.out + (k - p.k0) * p.ldo);
    const int words = static_cast<int>(p.ldo >> 2);
    bool left_na = false;
    if (len < 2 || len > 3) {   // :76  x = 0
      for (int w = tid; w < words; w += blockDim.x) dst[w] = 0u;
    } else {
      const uint32_t l0 = lev[0], l1 = lev[1];
      const int v0 = code[0], v1 = code[1], v2 = code[2], vn = na_val;
      for (int w = tid; w < words; w += blockDim.x) {
        uint32_t o = 0u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int64_t i = int64_t(w) * 4 + c;
          if (i < p.n) {
            const uint32_t key = read_key(row + i * p.stride, p.bit);
            const int v = (key == kKeyNA) ? vn : (key == l0 ? v0 : (key == l1 ? v1 : v2));
            left_na |= (v < 0);
            o |= (static_cast<uint32_t>(v) & 0xFFu) << (8 * c);
          }
        }
        dst[w] = o;
      }
    }
    if (left_na) atomicOr(p.has_na, 1);
    __syncthreads();   // shared state is reused by the next marker
  }
}

}  // namespace gapit

using namespace gapit;

// Index a HapMap text table: every data row's offset of the first genotype character (after the 11th tab).
extern "C" int gapit_hapmap_index(const char* text, int64_t bytes, int heading, int64_t* n_out, int64_t* m_out,
                                  int* bit_out, int64_t* row_off_out, int64_t row_cap, int64_t* header_off_out) {
  if (!text || bytes <= 0 || !n_out || !m_out || !bit_out) return fail(GAPIT_ERR_ARG, "gapit_hapmap_index: bad argument");
  int64_t pos = 0, m = 0, n = -1;
  int bit = 0;
  bool first = true;
  if (header_off_out) *header_off_out = -1;
  while (pos < bytes) {
    const char* nl = static_cast<const char*>(memchr(text + pos, '\n', size_t(bytes - pos)));
    int64_t end = nl ? (nl - text) : bytes;
    const int64_t next = end + 1;
    if (end > pos && text[end - 1] == '\r') --end;
    if (end > pos) {   // skip empty lines
      int tabs = 0;
      int64_t g0 = -1;
      for (int64_t i = pos; i < end; ++i)
        if (text[i] == '\t' && ++tabs == 11) {
          g0 = i + 1;
          break;
        }
      if (g0 < 0 || g0 >= end) return fail(GAPIT_ERR_DATA, "gapit_hapmap_index: a row has fewer than 12 columns");
      if (first && heading) {
        if (header_off_out) *header_off_out = g0;
      } else {
        if (n < 0) {   // genotype width and taxa count from the first data row (GAPIT.HapMap.R:28)
          int64_t i = g0;
          while (i < end && text[i] != '\t') ++i;
          bit = static_cast<int>(i - g0);
          if (bit < 1 || bit > 2) return fail(GAPIT_ERR_DATA, "gapit_hapmap_index: genotype codes must be 1 or 2 characters");
          n = (end - g0 + 1) / (bit + 1);
        }
        if (end - g0 != n * (bit + 1) - 1)
          return fail(GAPIT_ERR_DATA, "gapit_hapmap_index: row " + std::to_string(m + 1) + " does not hold " +
                                          std::to_string(n) + " genotypes of " + std::to_string(bit) + " characters");
        if (row_off_out) {
          if (m >= row_cap) return fail(GAPIT_ERR_ARG, "gapit_hapmap_index: row_off_out too small");
          row_off_out[m] = g0;
        }
        ++m;
      }
      first = false;
    }
    pos = next;
  }
  if (m == 0 || n <= 0) return fail(GAPIT_ERR_DATA, "gapit_hapmap_index: no data rows");
  *n_out = n;
  *m_out = m;
  *bit_out = bit;
  return GAPIT_OK;
}

extern "C" int gapit_hapmap_numericalize(gapit_ctx* ctx, const char* text, int64_t bytes, const int64_t* row_off, int bit,
                                         int stride, int64_t n, int64_t m, int effect, gapit_impute impute,
                                         int major_allele_zero, int8_t* gd_out, gapit_geno** geno_out) {
  if (!ctx || !text || !row_off || bytes <= 0 || n <= 0 || m <= 0 || bit < 1 || bit > 2 || stride < bit || effect < 0 ||
      effect > 3 || (!gd_out && !geno_out))
    return fail(GAPIT_ERR_ARG, "gapit_hapmap_numericalize: bad argument");
  for (int64_t k = 0; k < m; ++k)
    if (row_off[k] < 0 || row_off[k] + (n - 1) * stride + bit > bytes || (k && row_off[k] < row_off[k - 1]))
      return fail(GAPIT_ERR_ARG, "gapit_hapmap_numericalize: row offsets out of range or not ascending");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  const int64_t ldn = round_up(n, 128);
  gapit_geno* g = new gapit_geno();
  g->ctx = ctx;
  g->device = ctx->device;
  g->n = n;
  g->m = m;
  g->ldn = ldn;
  int64_t* off_dev = nullptr;
  int* flag_dev = nullptr;
  uint8_t* text_dev = nullptr;
  auto cleanup = [&](int rc) {
    cudaFree(off_dev);
    cudaFree(flag_dev);
    cudaFree(text_dev);
    gapit_geno_free(g);
    return rc;
  };
  const int64_t row_span = (n - 1) * stride + bit;
  // markers per upload: at most ~256 MB of file text (rows are ascending, so a block is one byte range)
  const int64_t budget = int64_t(256) << 20;
  cudaError_t e;
  if ((e = cudaMalloc(&g->mk, size_t(m) * ldn)) != cudaSuccess || (e = cudaMalloc(&g->colsum, size_t(m) * 4)) != cudaSuccess ||
      (e = cudaMalloc(&g->colsumsq, size_t(m) * 4)) != cudaSuccess || (e = cudaMalloc(&off_dev, size_t(m) * 8)) != cudaSuccess ||
      (e = cudaMalloc(&flag_dev, 8)) != cudaSuccess)
    return cleanup(fail(GAPIT_ERR_CUDA, std::string("gapit_hapmap_numericalize: cudaMalloc: ") + cudaGetErrorString(e)));
  int64_t max_span = 0;
  for (int64_t k0 = 0; k0 < m;) {
    int64_t k1 = k0 + 1;
    while (k1 < m && row_off[k1] + row_span - row_off[k0] <= budget) ++k1;
    max_span = std::max(max_span, row_off[k1 - 1] + row_span - row_off[k0]);
    k0 = k1;
  }
  if ((e = cudaMalloc(&text_dev, size_t(max_span) + 32)) != cudaSuccess)
    return cleanup(fail(GAPIT_ERR_CUDA, std::string("gapit_hapmap_numericalize: cudaMalloc: ") + cudaGetErrorString(e)));
  cudaEventRecord(ctx->ev0, ctx->stream);
  e = cudaMemcpyAsync(off_dev, row_off, size_t(m) * 8, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(flag_dev, 0, 8, ctx->stream);
  for (int64_t k0 = 0; k0 < m && e == cudaSuccess;) {
    int64_t k1 = k0 + 1;
    while (k1 < m && row_off[k1] + row_span - row_off[k0] <= budget) ++k1;
    const int64_t span = row_off[k1 - 1] + row_span - row_off[k0];
    e = cudaMemcpyAsync(text_dev, text + row_off[k0], size_t(span), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) break;
    NumParams p{text_dev, off_dev, row_off[k0], k0, k1, n, bit, stride, effect, int(impute), major_allele_zero,
                g->mk + k0 * ldn, ldn, flag_dev};
    // one CTA per marker; narrow tables use one-warp CTAs so that more markers are in flight per SM
    const int threads = n <= 4096 ? 32 : (n <= 16384 ? 128 : 256);
    const int blocks = static_cast<int>(std::min<int64_t>(k1 - k0, int64_t(ctx->sm_count) * (2048 / threads)));
    // the kernel also holds ~100 bytes of static shared memory: keep static + dynamic inside the 48 KB default limit
    if (row_span + 32 + 512 <= kStageBytes) {
      const size_t smem = size_t((row_span + 15 + 15) / 16) * 16;
      numericalize_kernel<true><<<blocks, threads, smem, ctx->stream>>>(p);
    } else {
      numericalize_kernel<false><<<blocks, threads, 0, ctx->stream>>>(p);
    }
    e = cudaGetLastError();
    k0 = k1;
  }
  cudaEventRecord(ctx->ev1, ctx->stream);
  int has_na = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&has_na, flag_dev, 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess && gd_out) e = cudaMemcpy2DAsync(gd_out, n, g->mk, ldn, n, m, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return cleanup(fail(GAPIT_ERR_CUDA, std::string("gapit_hapmap_numericalize: ") + cudaGetErrorString(e)));
  float ms = 0.f;
  ctx->stats = gapit_stats{};
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  ctx->stats.pack_ms = ms;
  if (!geno_out) return cleanup(GAPIT_OK);
  if (has_na) return cleanup(fail(GAPIT_ERR_DATA, "gapit_hapmap_numericalize: missing genotypes remain (impute = None); "
                                                  "a device genotype handle needs complete data"));
  int* bad = flag_dev + 1;
  int rc = launch_colstats(ctx, ctx->stream, g->mk, ldn, n, m, g->colsum, g->colsumsq, bad);
  if (rc != GAPIT_OK) return cleanup(rc);
  if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
    return cleanup(fail(GAPIT_ERR_CUDA, std::string("gapit_hapmap_numericalize: ") + cudaGetErrorString(e)));
  cudaFree(off_dev);
  cudaFree(flag_dev);
  cudaFree(text_dev);
  *geno_out = g;
  return GAPIT_OK;
}
