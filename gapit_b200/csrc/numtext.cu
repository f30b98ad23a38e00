// Numeric-text genotype ingestion: GAPIT's genoFormat = "EMMA" files (GAPIT.Fragment.R:94-136): `file.GD` is a table with
// a heading row and one row per TAXON -- the taxon name followed by one dosage (0 / 1 / 2, NA) per marker, separated by
// blanks, tabs or commas.  The reference re-reads the whole table once per fragment (read.table with colClasses,
// :100-109); here the text goes to the GPU once:
//
//   gapit_text_index_lines   host: line starts (one memchr pass)
//   numtext_parse_kernel     one CTA per taxon row: token starts by a block-wide scan over the row's bytes, every token
//                            parsed where it starts, 1 byte per call into a taxa-major staging matrix
//   byte_transpose_kernel    taxa-major -> the marker-major int8 operand layout of the kinship / scan kernels
// Byte / integer work, HBM- and instruction-bound; no tensor cores here.
#include <algorithm>
#include <cstring>
#include <string>

#include "common.cuh"

namespace gapit {

int launch_colstats(gapit_ctx* ctx, cudaStream_t st, const int8_t* mk, int64_t ldn, int64_t n, int64_t m, int32_t* colsum,
                    int32_t* colsumsq, int* bad_dev);

__device__ __forceinline__ bool is_delim(uint8_t c) { return c == ' ' || c == '\t' || c == ',' || c == '\r' || c == '\n'; }

// token at p (not a delimiter), at most `avail` readable bytes: 0 / 1 / 2 (optionally "1.0", "2.00"), NA spellings -> 3,
// anything else -> 255
__device__ uint8_t parse_dosage(const uint8_t* p, int avail) {
  int len = 0;
  while (len < avail && !is_delim(p[len])) ++len;
  if (len == 0) return 255;
  const uint8_t c0 = p[0];
  if (c0 >= '0' && c0 <= '2') {
    if (len == 1) return c0 - '0';
    if (p[1] != '.') return 255;
    for (int i = 2; i < len; ++i)
      if (p[i] != '0') return 255;
    return c0 - '0';
  }
  if (len == 2 && c0 == 'N' && p[1] == 'A') return 3;
  if (len == 3 && (c0 == 'N' || c0 == 'n') && (p[1] == 'a' || p[1] == 'A') && (p[2] == 'N' || p[2] == 'n')) return 3;
  if (len == 1 && (c0 == '.' || c0 == '-')) return 3;
  return 255;
}

constexpr int kNumChunk = 4096;   // bytes of a row staged per round (256 threads x 16)
constexpr int kNumTokMax = 24;    // longest token looked at (the staging tile overlaps the next round by this much)

// flags in *bad: 1 = token that is no dosage, 2 = NA without an impute rule, 4 = a row with the wrong number of tokens
__global__ void __launch_bounds__(256) numtext_parse_kernel(const uint8_t* __restrict__ text, const int64_t* __restrict__ line_off,
                                                            int64_t text_base, int64_t row0, int64_t nrows, int64_t m,
                                                            int skip_tokens, int impute, uint8_t* __restrict__ tx,
                                                            int64_t ldm, int* __restrict__ bad) {
  __shared__ uint8_t tile[kNumChunk + kNumTokMax + 8];
  __shared__ int warp_sum[8];
  __shared__ int carry_tokens;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int flag = 0;
  for (int64_t r = blockIdx.x; r < nrows; r += gridDim.x) {
    const int64_t b0 = line_off[row0 + r] - text_base, b1 = line_off[row0 + r + 1] - text_base;   // [b0, b1): row incl. newline
    uint8_t* out = tx + (row0 + r) * ldm;
    if (tid == 0) carry_tokens = 0;
    uint8_t prev_last = ' ';   // the byte before the row counts as a delimiter
    for (int64_t c0 = b0; c0 < b1; c0 += kNumChunk) {
      __syncthreads();
      const int len = static_cast<int>((b1 - c0 < kNumChunk + kNumTokMax) ? (b1 - c0) : int64_t(kNumChunk + kNumTokMax));
      for (int i = tid; i < kNumChunk + kNumTokMax; i += 256) tile[i] = i < len ? text[c0 + i] : uint8_t('\n');
      __syncthreads();
      const int base = tid * 16;
      const int body = static_cast<int>((b1 - c0 < kNumChunk) ? (b1 - c0) : int64_t(kNumChunk));   // token starts are counted in [0, body)
      uint32_t starts = 0;
      uint8_t p = base == 0 ? prev_last : tile[base - 1];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint8_t c = tile[base + j];
        if (base + j < body && is_delim(p) && !is_delim(c)) starts |= 1u << j;
        p = c;
      }
      const int cnt = __popc(starts);
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) warp_sum[warp] = incl;
      __syncthreads();
      int before = carry_tokens;
      for (int w = 0; w < warp; ++w) before += warp_sum[w];
      int t = before + incl - cnt;   // index of my first token in the row
      while (starts) {
        const int j = __ffs(starts) - 1;
        starts &= starts - 1;
        const int64_t k = int64_t(t) - skip_tokens;
        if (k >= 0) {
          if (k >= m) {
            flag |= 4;
          } else {
            uint8_t v = parse_dosage(tile + base + j, kNumTokMax);
            if (v == 255) {
              flag |= 1;
              v = 0;
            } else if (v == 3) {
              if (impute < 0) flag |= 2;
              v = impute < 0 ? 0 : static_cast<uint8_t>(impute);
            }
            out[k] = v;
          }
        }
        ++t;
      }
      prev_last = tile[kNumChunk - 1];
      __syncthreads();
      if (tid == 255) carry_tokens = before + incl;
    }
    __syncthreads();
    if (tid == 0 && carry_tokens != m + skip_tokens) flag |= 4;
  }
  if (flag) atomicOr(bad, flag);
}

// dst[c][r] = src[r][c] for bytes: 128 x 128 tiles through shared memory, word accesses on both sides
__global__ void __launch_bounds__(1024) byte_transpose_kernel(const uint8_t* __restrict__ src, int64_t ld_src, int64_t rows,
                                                              int64_t cols, uint8_t* __restrict__ dst, int64_t ld_dst) {
  __shared__ uint32_t tile[128][33];
  const int tx = threadIdx.x, ty = threadIdx.y;     // 32 x 32
  const int64_t c0 = int64_t(blockIdx.x) * 128, r0 = int64_t(blockIdx.y) * 128;
  uint32_t w[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int64_t r = r0 + ty * 4 + q, c = c0 + tx * 4;
    w[q] = (r < rows && c < ld_src) ? __ldg(reinterpret_cast<const uint32_t*>(src + r * ld_src + c)) : 0u;
  }
  // 4x4 byte transpose: o[c] = { w[0].c, w[1].c, w[2].c, w[3].c }
  const uint32_t t0 = __byte_perm(w[0], w[1], 0x5140), t1 = __byte_perm(w[0], w[1], 0x7362);
  const uint32_t t2 = __byte_perm(w[2], w[3], 0x5140), t3 = __byte_perm(w[2], w[3], 0x7362);
  tile[tx * 4 + 0][ty] = __byte_perm(t0, t2, 0x5410);
  tile[tx * 4 + 1][ty] = __byte_perm(t0, t2, 0x7632);
  tile[tx * 4 + 2][ty] = __byte_perm(t1, t3, 0x5410);
  tile[tx * 4 + 3][ty] = __byte_perm(t1, t3, 0x7632);
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int rr = ty + 32 * q;                      // row of the tile = one source column
    const int64_t c = c0 + rr;
    if (c < cols && r0 + tx * 4 < ld_dst) *reinterpret_cast<uint32_t*>(dst + c * ld_dst + r0 + tx * 4) = tile[rr][tx];
  }
}

}  // namespace gapit

using namespace gapit;

// offsets of the line starts of `text` (a final line without newline counts); off_out[nlines] = bytes (one past the end).
// Call with off_out = NULL to size the array (nlines + 1 entries).
extern "C" int gapit_text_index_lines(const char* text, int64_t bytes, int64_t* off_out, int64_t cap, int64_t* nlines_out) {
  if (!text || bytes <= 0 || !nlines_out) return fail(GAPIT_ERR_ARG, "gapit_text_index_lines: bad argument");
  int64_t nl = 0, pos = 0;
  while (pos < bytes) {
    if (off_out) {
      if (nl >= cap) return fail(GAPIT_ERR_ARG, "gapit_text_index_lines: offset array too small");
      off_out[nl] = pos;
    }
    ++nl;
    const void* e = memchr(text + pos, '\n', size_t(bytes - pos));
    pos = e ? (static_cast<const char*>(e) - text) + 1 : bytes;
  }
  if (off_out) {
    if (nl >= cap) return fail(GAPIT_ERR_ARG, "gapit_text_index_lines: offset array too small");
    off_out[nl] = bytes;
  }
  *nlines_out = nl;
  return GAPIT_OK;
}

// number of tokens of line [b0, b1)
static int64_t count_tokens(const char* text, int64_t b0, int64_t b1) {
  int64_t t = 0;
  bool prev = true;
  for (int64_t i = b0; i < b1; ++i) {
    const char c = text[i];
    const bool d = c == ' ' || c == '\t' || c == ',' || c == '\r' || c == '\n';
    if (prev && !d) ++t;
    prev = d;
  }
  return t;
}

extern "C" int gapit_numtext_shape(const char* text, int64_t bytes, const int64_t* line_off, int64_t nlines, int heading,
                                   int64_t* n_out, int64_t* m_out) {
  if (!text || !line_off || !n_out || !m_out || nlines < (heading ? 2 : 1))
    return fail(GAPIT_ERR_ARG, "gapit_numtext_shape: bad argument");
  int64_t last = nlines;   // trailing blank lines are not taxa
  while (last > 0 && count_tokens(text, line_off[last - 1], line_off[last]) == 0) --last;
  const int64_t first = heading ? 1 : 0;
  if (last <= first) return fail(GAPIT_ERR_DATA, "gapit_numtext_shape: no data rows");
  *n_out = last - first;
  *m_out = count_tokens(text, line_off[first], line_off[first + 1]) - 1;   // minus the taxon name
  if (*m_out < 1) return fail(GAPIT_ERR_DATA, "gapit_numtext_shape: a data row holds no genotype");
  (void)bytes;
  return GAPIT_OK;
}

// text: the whole table on the host; line_off: gapit_text_index_lines; rows [first_row, first_row + n) are taxa.
// gd_out (may be NULL): n x m int8 in R's column-major layout (= marker-major bytes); geno_out (may be NULL).
extern "C" int gapit_numtext_numericalize(gapit_ctx* ctx, const char* text, int64_t bytes, const int64_t* line_off,
                                          int64_t first_row, int64_t n, int64_t m, gapit_impute impute, int8_t* gd_out,
                                          gapit_geno** geno_out) {
  if (!ctx || !text || !line_off || n <= 0 || m <= 0 || first_row < 0 || (!gd_out && !geno_out))
    return fail(GAPIT_ERR_ARG, "gapit_numtext_numericalize: bad argument");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  const int64_t ldm = round_up(m, 128), ldn = round_up(n, 128);
  gapit_geno* g = new gapit_geno();
  g->ctx = ctx;
  g->device = ctx->device;
  g->n = n;
  g->m = m;
  g->ldn = ldn;
  uint8_t *tx = nullptr, *text_dev[2] = {nullptr, nullptr};
  int64_t* off_dev = nullptr;
  int* flag_dev = nullptr;
  auto cleanup = [&](int rc) {
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
    cudaFree(tx);
    cudaFree(text_dev[0]);
    cudaFree(text_dev[1]);
    cudaFree(off_dev);
    cudaFree(flag_dev);
    if (rc != GAPIT_OK || !geno_out) gapit_geno_free(g);
    return rc;
  };
  // row blocks of ~128 MB of text, uploaded on the copy stream while the previous block is parsed
  const int64_t budget = int64_t(128) << 20;
  int64_t max_span = 0;
  for (int64_t r0 = 0; r0 < n;) {
    int64_t r1 = r0 + 1;
    while (r1 < n && line_off[first_row + r1 + 1] - line_off[first_row + r0] <= budget) ++r1;
    max_span = std::max(max_span, line_off[first_row + r1] - line_off[first_row + r0]);
    r0 = r1;
  }
  cudaError_t e = cudaMalloc(&g->mk, size_t(m) * ldn);
  if (e == cudaSuccess) e = cudaMalloc(&g->colsum, size_t(m) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&g->colsumsq, size_t(m) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&tx, size_t(ldn) * ldm);
  if (e == cudaSuccess) e = cudaMalloc(&text_dev[0], size_t(max_span) + 64);
  if (e == cudaSuccess) e = cudaMalloc(&text_dev[1], size_t(max_span) + 64);
  if (e == cudaSuccess) e = cudaMalloc(&off_dev, size_t(n + 1) * 8);
  if (e == cudaSuccess) e = cudaMalloc(&flag_dev, 8);
  if (e != cudaSuccess) return cleanup(fail(GAPIT_ERR_CUDA, std::string("gapit_numtext_numericalize: cudaMalloc: ") + cudaGetErrorString(e)));
  cudaEventRecord(ctx->ev0, ctx->stream);
  e = cudaMemcpyAsync(off_dev, line_off + first_row, size_t(n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(flag_dev, 0, 8, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(tx, 0, size_t(ldn) * ldm, ctx->stream);
  if (e == cudaSuccess) e = cudaEventRecord(ctx->ev2, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev2, 0);
  int blk = 0;
  for (int64_t r0 = 0; r0 < n && e == cudaSuccess; ++blk) {
    int64_t r1 = r0 + 1;
    while (r1 < n && line_off[first_row + r1 + 1] - line_off[first_row + r0] <= budget) ++r1;
    const int b = blk & 1;
    const int64_t t0 = line_off[first_row + r0], t1 = line_off[first_row + r1];
    if (blk >= 2) e = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[b], 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(text_dev[b], text + t0, size_t(t1 - t0), cudaMemcpyHostToDevice, ctx->copy_stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_ready[b], ctx->copy_stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->ev_ready[b], 0);
    if (e != cudaSuccess) break;
    const int grid = static_cast<int>(std::min<int64_t>(r1 - r0, int64_t(ctx->sm_count) * 8));
    numtext_parse_kernel<<<grid, 256, 0, ctx->stream>>>(text_dev[b], off_dev, t0, r0, r1 - r0, m, 1, int(impute), tx, ldm, flag_dev);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_free[b], ctx->stream);
    r0 = r1;
  }
  if (e == cudaSuccess) {
    dim3 grid(static_cast<unsigned>(ldm / 128), static_cast<unsigned>(ldn / 128));
    byte_transpose_kernel<<<grid, dim3(32, 32), 0, ctx->stream>>>(tx, ldm, n, m, reinterpret_cast<uint8_t*>(g->mk), ldn);
    e = cudaGetLastError();
  }
  cudaEventRecord(ctx->ev1, ctx->stream);
  int flag = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, flag_dev, 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess && gd_out) e = cudaMemcpy2DAsync(gd_out, n, g->mk, ldn, n, m, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return cleanup(fail(GAPIT_ERR_CUDA, std::string("gapit_numtext_numericalize: ") + cudaGetErrorString(e)));
  if (flag & 4) return cleanup(fail(GAPIT_ERR_DATA, "numeric genotype table: a row does not hold one value per marker"));
  if (flag & 1) return cleanup(fail(GAPIT_ERR_DATA, "numeric genotype table: a value is not 0, 1, 2 or NA"));
  if (flag & 2) return cleanup(fail(GAPIT_ERR_DATA, "numeric genotype table holds NA and no impute rule was given"));
  float ms = 0.f;
  ctx->stats = gapit_stats{};
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  ctx->stats.pack_ms = ms;
  if (!geno_out) return cleanup(GAPIT_OK);
  int* bad = flag_dev + 1;
  const int rc = launch_colstats(ctx, ctx->stream, g->mk, ldn, n, m, g->colsum, g->colsumsq, bad);
  if (rc != GAPIT_OK) return cleanup(rc);
  if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
    return cleanup(fail(GAPIT_ERR_CUDA, std::string("gapit_numtext_numericalize: ") + cudaGetErrorString(e)));
  *geno_out = g;
  return cleanup(GAPIT_OK);
}

// PLINK .bed: magic 0x6C 0x1B, mode 0x01 (variant-major); m variants of ceil(n/4) bytes follow.  Returns m and the
// offset of the first variant (3); the rows go to gapit_geno_pack / gapit_p3d_scan_host / gapit_mgeno_pack as
// GAPIT_BED_A1 or GAPIT_BED_A2.
extern "C" int gapit_bed_header(const void* bytes, int64_t len, int64_t n, int64_t* m_out, int64_t* data_off_out) {
  if (!bytes || len < 3 || n <= 0 || !m_out) return fail(GAPIT_ERR_ARG, "gapit_bed_header: bad argument");
  const uint8_t* b = static_cast<const uint8_t*>(bytes);
  if (b[0] != 0x6C || b[1] != 0x1B) return fail(GAPIT_ERR_DATA, "not a PLINK .bed file (magic bytes 6C 1B missing)");
  if (b[2] != 0x01) return fail(GAPIT_ERR_DATA, "PLINK .bed in sample-major mode (third byte 00): re-export variant-major");
  const int64_t nb = (n + 3) / 4;
  if ((len - 3) % nb != 0) return fail(GAPIT_ERR_DATA, "PLINK .bed length is not a whole number of variants for this .fam");
  *m_out = (len - 3) / nb;
  if (data_off_out) *data_off_out = 3;
  return GAPIT_OK;
}
