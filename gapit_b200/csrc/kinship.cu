// VanRaden kinship (GAPIT.kinship.VanRaden.R:11-32), B200-native.
//
//   K1  gram_i8_umma      G += T T'   int8 x int8 -> int32 on tcgen05 (kind::i8), lower-triangular
//                         256 x 256 tiles, split along the marker axis; partial tiles are added with
//                         integer RED (exact, order-independent => bit-identical for any split / GPU count)
//   K2a colsum_reduce     sum c, sum c^2, #columns with c == 2n          (int64, exact)
//   K2b gram_mirror       fill the strictly-upper tiles from the lower ones
//   K2c gram_rowsum       s_i = sum_j G_ij                                (int64, exact)
//   K2d vanraden_epilogue K_ij = 2 (n^2 G'_ij - n (s'_i + s'_j) + sum c^2) / D   (one fp64 division per element)
//
// Algebra.  With c_k = sum_i x_ik, the reference's Z[k,i] = x_ik - 2 p_k = x_ik - c_k/n (:20-25) gives
//   (Z'Z)_ij = G_ij - (s_i + s_j)/n + sum_k c_k^2 / n^2,  s_i = sum_k c_k x_ik = sum_j G_ij,
//   adj = 2 sum p(1-p) = D / (2 n^2),  D = sum_k c_k (2n - c_k)   (:31)
// over the columns the reference keeps (:11-13).  Dropped columns: c_k = 0 contributes nothing anywhere;
// c_k = 2n (all dosages 2) contributes exactly 4 to every G_ij, 4n to every s_i, 4n^2 to sum c^2 and 0 to D,
// so the epilogue subtracts n2 = #(c_k == 2n) of those instead of re-packing the matrix.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "umma_i8.cuh"

namespace gapit {


// ------------------------------------------------------------------ K1 pieces
struct GramSched {
  struct Params {
    int tiles_side;  // T: number of 256-row blocks
    int cl;          // CTAs per cluster: a cluster takes CL consecutive block-rows against one block-column
    int ntri;        // jobs per marker split = sum over row groups R of min(CL (R+1), T)
    int ksplit;      // splits of the marker axis
    int kblocks;     // marker axis in units of kBlockK
    unsigned int* round_ctr;  // [rounds] zeroed before the launch, or nullptr (see ScanSched::align_round)
    int round_wait;           // spin budget in clock cycles
  };
  __device__ uint64_t hint_a() const { return ptx::kEvictNormal; }
  __device__ uint64_t hint_b() const { return ptx::kEvictNormal; }
  // All jobs have the same length; starting each unit's r-th job together keeps the tiles in flight on the same marker
  // window, so the ~24 row panels they share are fetched from DRAM once.  Soft barrier, bounded wait (no hang).
  __device__ void align_round() const {
    if (!p.round_ctr) return;
    const int job = next_job - stride;   // the job next() just handed out
    const int r = job / stride;
    const unsigned int expect = static_cast<unsigned int>(min(stride, p.ntri * p.ksplit - r * stride));
    volatile unsigned int* c = p.round_ctr + r;
    atomicAdd(p.round_ctr + r, 1u);
    const long long t0 = clock64();
    while (*c < expect && clock64() - t0 < p.round_wait) __nanosleep(64);
  }
  static constexpr int kBandRows = 12;  // block-rows per band
  Params p;
  int next_job, stride, crank;
  __device__ GramSched(const Params& pp, int unit, int nunits, uint32_t cta_rank)
      : p(pp), next_job(unit), stride(nunits), crank(static_cast<int>(cta_rank)) {}
  __device__ bool next(Job& j) {
    if (next_job >= p.ntri * p.ksplit) return false;
    // split-major order: CTAs running together share one marker window (L2 reuse of the row panels).
    // Inside a split the lower triangle is walked in bands of kBandRows block-rows, column by column, so the
    // ~148 tiles in flight cover about 12 x 12 blocks: 24 row panels in L2 feed 144 tiles (a plain row-major
    // walk would stream one private 256-row panel per tile from HBM).  Row group R = CL block-rows
    // [CL R, CL R + CL); it meets block-columns J < min(CL (R+1), T) (with CL = 2 the one tile above the
    // diagonal per group is computed redundantly and later overwritten by the mirror pass).
    const int ks = next_job / p.ntri;
    int t = next_job - ks * p.ntri;
    const int T = p.tiles_side, CL = p.cl;
    const int TR = (T + CL - 1) / CL;
    const int GB = kBandRows / CL;
    int R0 = 0, rows = 0;
    for (;;) {
      rows = min(GB, TR - R0);
      int cnt = 0;
      for (int r = 0; r < rows; ++r) cnt += min(CL * (R0 + r + 1), T);
      if (t < cnt) break;
      t -= cnt;
      R0 += rows;
    }
    int R, J;
    const int full_cols = min(CL * R0 + CL, T);   // columns met by every row group of the band
    if (t < rows * full_cols) {
      J = t / rows;
      R = R0 + t - J * rows;
    } else {                                      // tail: column J is met by row groups R >= J / CL
      t -= rows * full_cols;
      J = full_cols;
      for (;;) {
        const int c = R0 + rows - J / CL;
        if (t < c) break;
        t -= c;
        ++J;
      }
      R = J / CL + t;
    }
    const int I = R * CL + crank;
    j.a_row0 = I * 256;
    j.b_row0 = J * 256;
    j.kb_begin = static_cast<int>((static_cast<long long>(ks) * p.kblocks) / p.ksplit);
    j.kb_end = static_cast<int>((static_cast<long long>(ks + 1) * p.kblocks) / p.ksplit);
    j.tag0 = I;
    j.tag1 = J;
    j.tag2 = ks;
    j.group_first = j.group_last = true;
    next_job += stride;
    return true;
  }
};

int make_tensor_map_i32_box32(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t ld_bytes);

// Mode 3: the accumulator leaves through TMA.  Each epilogue warp owns 32 accumulator rows; per 32-column chunk it
// stages a 32 x 32 int32 box in shared memory (128-byte swizzle, conflict-free 16-byte stores) and one lane issues
// cp.reduce.async.bulk.tensor (.add) into the Gram buffer of the rank that owns the block-row -- local or peer-mapped
// over NVLink -- so the exchange travels as 4 KB bulk reductions instead of 4-byte ones, overlapped with the MMAs of
// the other CTAs.  Integer adds: bit-identical to every other route.
struct GramEpiTma {
  struct Params {
    CUtensorMap maps[8];   // [world] Gram buffer of each rank, boxes of 32 x 32 int32
    int world;
    int tiles_side;
    int root;              // >= 0: every tile goes to this rank; < 0: to the owner of its block-row
  };
  static constexpr int kSmemBytes = 8 * 4096 + 1024;   // 8 warps x one box, 1024-byte aligned for the swizzle atoms
  const Params* pp;
  uint8_t* stage;
  __device__ GramEpiTma(const Params& p_, uint8_t* scratch) : pp(&p_) {
    stage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(scratch) + 1023) & ~uintptr_t(1023));
  }
  __device__ void prepare(const Job&, int, int) {}
  __device__ void finish() {
    if ((threadIdx.x & 31) == 0) ptx::bulk_wait_group0();   // every bulk reduce of this warp has been performed
  }
  __device__ void run(const Job& job, uint32_t taddr, int sub, int row, int /*half*/, int /*nhalf*/) {
    const int lane = threadIdx.x & 31;
    const int ewarp = (threadIdx.x >> 5) - 4;          // 0..7: epilogue warp of this CTA
    uint8_t* box = stage + ewarp * 4096;
    const int owner = pp->root >= 0 ? pp->root : min(pp->world - 1, (job.tag0 * pp->world) / pp->tiles_side);
    const CUtensorMap* map = &pp->maps[owner];
    const int row0 = job.a_row0 + sub * 128 + (row & ~31);
#pragma unroll 1
    for (int c0 = 0; c0 < 256; c0 += 32) {
      uint32_t r[32];
      ptx::tmem_ld_x16(taddr + c0, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
      ptx::tmem_ld_x16(taddr + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
      if (lane == 0) ptx::bulk_wait_group_read0();     // the previous box has been read out of the staging tile
      __syncwarp();
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j)                       // row `lane`, 16-byte chunk j at its swizzled position
        *reinterpret_cast<uint4*>(box + lane * 128 + ((j ^ (lane & 7)) << 4)) =
            make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::tma_reduce_add_2d(map, box, job.b_row0 + c0, row0);
        ptx::bulk_commit_group();
      }
    }
  }
};


// ------------------------------------------------------------------ K2 pieces
__global__ void colsum_reduce_kernel(const int32_t* __restrict__ colsum, int64_t m, int64_t two_n,
                                     unsigned long long* __restrict__ sums /* [3] */) {
  long long s1 = 0, s2 = 0, n2 = 0;
  for (int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < m; k += int64_t(gridDim.x) * blockDim.x) {
    const long long c = colsum[k];
    s1 += c;
    s2 += c * c;
    n2 += (c == two_n);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    n2 += __shfl_xor_sync(0xffffffffu, n2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&sums[0], static_cast<unsigned long long>(s1));
    atomicAdd(&sums[1], static_cast<unsigned long long>(s2));
    atomicAdd(&sums[2], static_cast<unsigned long long>(n2));
  }
}

// the three scalar sums of this rank's markers added into every rank's mailbox (peer-mapped memory, system-scope atomics):
// after the exchange's closing barrier each mailbox holds the totals, no collective needed for them
struct SumPeers {
  unsigned long long* p[8];
};
__global__ void push_sums_kernel(const unsigned long long* __restrict__ local, SumPeers peers, int world) {
  const int r = threadIdx.x / 3, i = threadIdx.x % 3;
  if (r < world) atomicAdd_system(peers.p[r] + i, local[i]);
}

// upper tile (I < J in 256-blocks) <- transpose of lower tile; 32x32 sub-tiles through shared memory
__global__ void gram_mirror_kernel(int32_t* __restrict__ G, int64_t ldg) {
  __shared__ int32_t t[32][33];
  const int64_t bi = blockIdx.y, bj = blockIdx.x;  // 32-blocks: source (bi rows, bj cols), bi >= bj region
  if ((bj >> 3) >= (bi >> 3)) return;              // only strictly-lower 256-tiles are sources
  const int x = threadIdx.x, y0 = threadIdx.y;
  for (int y = y0; y < 32; y += 8) t[y][x] = G[(bi * 32 + y) * ldg + bj * 32 + x];
  __syncthreads();
  for (int y = y0; y < 32; y += 8) G[(bj * 32 + y) * ldg + bi * 32 + x] = t[x][y];
}

__global__ void gram_rowsum_kernel(const int32_t* __restrict__ G, int64_t ldg, int64_t n, long long* __restrict__ s) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const int32_t* g = G + row * ldg;
  long long acc = 0;
  for (int64_t j = lane; j < n; j += 32) acc += g[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) s[row] = acc;
}

// ---- K (column-major n x n) from the LOWER triangle of G alone (no mirror pass) --------------------------------------
// s_i = sum_j G_ij over the full symmetric row = the part of row i the lower triangle holds (columns up to the end of its
// diagonal 32 x 32 tile, which carries both triangles) + the part of column i below that tile.  ONE pass over the
// triangle: a CTA owns a strip of 32 rows and streams it in 128-column chunks (a warp per chunk, 32 independent 16-byte
// loads per lane: the strip is read at HBM rate); a lane keeps the 32 row partials in registers for the whole strip and the
// column sums of its 4 columns per chunk, which go to colpart[strip][column] -- complete for that strip, single writer.
// A second small kernel adds the column parts below each diagonal tile.  No atomics, fixed order.
// (Two strided passes with 157 CTAs took 0.24 ms of the 3.5 ms kinship step at n = 5,000 and 11 ms of 95 ms at n = 50,000.)
__global__ void __launch_bounds__(256) gram_strip_sums_kernel(const int32_t* __restrict__ G, int64_t ldg, int tiles32,
                                                              long long* __restrict__ rowsum, long long* __restrict__ colpart) {
  const int b = tiles32 - 1 - static_cast<int>(blockIdx.x);   // longest strips first
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row0 = int64_t(b) * 32;
  const int64_t col_end = row0 + 32;   // through the diagonal tile
  long long racc[32];
#pragma unroll
  for (int r = 0; r < 32; ++r) racc[r] = 0;
  const int nchunks = static_cast<int>((col_end + 127) / 128);
  for (int c = warp; c < nchunks; c += 8) {
    const int64_t j0 = int64_t(c) * 128 + 4 * lane;
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    const int32_t* src = G + row0 * ldg + j0;
#pragma unroll
    for (int r8 = 0; r8 < 32; r8 += 8) {
      int4 v[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) v[r] = __ldg(reinterpret_cast<const int4*>(src + int64_t(r8 + r) * ldg));
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (j0 + 3 >= col_end) {   // the chunk that holds the diagonal tile: columns past it are not part of the triangle
          if (j0 + 0 >= col_end) v[r].x = 0;
          if (j0 + 1 >= col_end) v[r].y = 0;
          if (j0 + 2 >= col_end) v[r].z = 0;
          if (j0 + 3 >= col_end) v[r].w = 0;
        }
        racc[r8 + r] += static_cast<long long>(v[r].x) + v[r].y + v[r].z + v[r].w;
        c0 += v[r].x;
        c1 += v[r].y;
        c2 += v[r].z;
        c3 += v[r].w;
      }
    }
    if (j0 < row0) {   // strictly below the diagonal tile (row0 is a multiple of 32, j0 of 4: all four columns or none)
      long long* dst = colpart + int64_t(b) * ldg + j0;
      reinterpret_cast<longlong2*>(dst)[0] = make_longlong2(c0, c1);
      reinterpret_cast<longlong2*>(dst)[1] = make_longlong2(c2, c3);
    }
  }
  __shared__ long long red[8][32];
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    long long v = racc[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == r) red[warp][r] = v;
  }
  __syncthreads();
  if (warp == 0) {
    long long t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][lane];
    rowsum[row0 + lane] = t;
  }
}

// s_j = rowsum_j + the column parts of the strips below j's diagonal tile; one CTA per 32 columns, 8 strips in flight
__global__ void gram_strip_combine_kernel(const long long* __restrict__ colpart, int64_t ldg, int tiles32,
                                          long long* __restrict__ s) {
  const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
  const int bj = blockIdx.x;
  const int64_t j = int64_t(bj) * 32 + x;
  long long t = 0;
  for (int b = bj + 1 + y; b < tiles32; b += 8) t += colpart[int64_t(b) * ldg + j];
  __shared__ long long red[8][32];
  red[y][x] = t;
  __syncthreads();
  if (y == 0) {
    long long v = s[j];
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][x];
    s[j] = v;
  }
}

// K_ij and K_ji from the lower 64 x 64 tile (i in block-row bi, j in block-column bj <= bi; CTAs are numbered along the
// triangle, so none is launched for the upper half; a 512 x 512 super-tile order was measured slower: 11.9 vs 9.5 ms
// at n = 50,000): exact int64 numerator, one fp64 division; the (j, i) entry is
// written as read (contiguous along j), the (i, j) entry through a shared-memory transpose.  16 independent loads per
// thread; 20 GB of fp64 written at n = 50,000 -- the kernel is bound by the HBM write rate.
__global__ void __launch_bounds__(256, 3) vanraden_epilogue_lower_kernel(const int32_t* __restrict__ G, int64_t ldg, int64_t n,
                                                                         const long long* __restrict__ s,
                                                                         const unsigned long long* __restrict__ sums,
                                                                         double* __restrict__ K) {
  // triangular index -> (bi, bj), bj <= bi
  const long long t = blockIdx.x;
  int bi = static_cast<int>((sqrt(8.0 * static_cast<double>(t) + 1.0) - 1.0) * 0.5);
  while (static_cast<long long>(bi) * (bi + 1) / 2 > t) --bi;
  while (static_cast<long long>(bi + 1) * (bi + 2) / 2 <= t) ++bi;
  const int bj = static_cast<int>(t - static_cast<long long>(bi) * (bi + 1) / 2);
  __shared__ double tile[64][65];
  __shared__ long long ai_sh[64];
  const long long sum_c = static_cast<long long>(sums[0]);
  const long long sum_c2 = static_cast<long long>(sums[1]);
  const long long n2 = static_cast<long long>(sums[2]);
  const long long nn = n;
  // remove the all-2 columns (GAPIT.kinship.VanRaden.R:11-13)
  const long long c2 = sum_c2 - n2 * 4 * nn * nn;
  const long long c1 = sum_c - n2 * 2 * nn;
  const double D = static_cast<double>(2 * nn * c1 - c2);
  // num_ij = n^2 (G_ij - 4 n2) - n (s_i + s_j) + c2 = n^2 G_ij + a_i + b_j  (exact in int64)
  const long long nn2 = nn * nn;
  const int x = threadIdx.x & 63, y0 = threadIdx.x >> 6;
  const int64_t i0 = int64_t(bi) * 64, j0 = int64_t(bj) * 64;
  if (threadIdx.x < 64)
    ai_sh[threadIdx.x] = (i0 + threadIdx.x < n) ? c2 - nn2 * 4 * n2 - nn * (s[i0 + threadIdx.x] - 4 * n2 * nn) : 0;
  const int64_t j = j0 + x;
  const long long b_j = (j < n) ? -nn * (s[j] - 4 * n2 * nn) : 0;
  int32_t g[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int64_t i = i0 + y0 + 4 * q;
    g[q] = (i < n && j < n) ? __ldg(G + i * ldg + j) : 0;
  }
  __syncthreads();
  // v = 2 num / D, correctly rounded: one multiplication by the rounded reciprocal and one residual correction (the
  // refinement step an fp64 division ends with) instead of the ~20-instruction division sequence per element
  const double Dh = 0.5 * D, r = 1.0 / Dh;
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int y = y0 + 4 * q;
    const int64_t i = i0 + y;
    double v = 0.0;
    if (i < n && j < n) {
      const double numd = static_cast<double>(nn2 * static_cast<long long>(g[q]) + ai_sh[y] + b_j);
      v = numd * r;
      v = fma(fma(-v, Dh, numd), r, v);
      K[j + i * n] = v;                                  // element (j, i): contiguous along j
    }
    tile[y][x] = v;
  }
  if (bi == bj) return;                                  // the diagonal tile holds both triangles: written above
  __syncthreads();
  const int64_t i2 = i0 + x;                             // now x runs along i
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int y = y0 + 4 * q;
    const int64_t j2 = j0 + y;
    if (i2 < n && j2 < n) K[i2 + j2 * n] = tile[x][y];   // element (i, j): contiguous along i
  }
}

// Centred cross-product C_ij = sum_k (x_ik - c_k/n)(x_jk - c_k/n) = (n^2 G_ij - n (s_i + s_j) + sum c^2) / n^2:
// the matrix GAPIT.kinship.ZHANG.R:21-26 calls K before its adjustments and the X X' of prcomp(X) (GAPIT.PCA.R:10).
// Monomorphic columns contribute exactly 0, so nothing has to be dropped.  Exact int64 numerator, one division.
__global__ void centered_gram_kernel(const int32_t* __restrict__ G, int64_t ldg, int64_t n, const long long* __restrict__ s,
                                     const unsigned long long* __restrict__ sums, double* __restrict__ K) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long nn = n;
  const long long c2 = static_cast<long long>(sums[1]);
  const long long si = s[i];
  const double d = static_cast<double>(nn) * static_cast<double>(nn);
  for (int64_t j = blockIdx.y; j < n; j += gridDim.y) {
    const long long num = nn * nn * static_cast<long long>(G[j * ldg + i]) - nn * (si + s[j]) + c2;
    K[i + j * n] = static_cast<double>(num) / d;
  }
}

// per-taxon number of heterozygous calls: het_i = sum_k [x_ik == 1] (GAPIT.kinship.ZHANG.R:14-15), straight from the
// marker-major matrix: a thread owns four taxa (one 32-bit word of every marker row), walks a slab of markers and
// counts the bytes equal to 1 in four packed 8-bit lanes (flushed before they can overflow); coalesced row reads.
__global__ void taxa_het_kernel(const int8_t* __restrict__ mk, int64_t ldn, int64_t m, int64_t rows_per_slab,
                                unsigned long long* __restrict__ het /* [ldn], zero-initialised */) {
  const int64_t w = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (w * 4 >= ldn) return;
  const int64_t k0 = int64_t(blockIdx.y) * rows_per_slab;
  const int64_t k1 = min(m, k0 + rows_per_slab);
  unsigned int c[4] = {0u, 0u, 0u, 0u};
  for (int64_t kb = k0; kb < k1; kb += 255) {
    uint32_t lanes = 0u;
    const int64_t ke = min(k1, kb + 255);
    for (int64_t k = kb; k < ke; ++k) {
      const uint32_t x = __ldg(reinterpret_cast<const uint32_t*>(mk + k * ldn) + w);
      lanes += x & ~(x >> 1) & 0x01010101u;   // byte == 1 (values are 0, 1, 2)
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] += (lanes >> (8 * j)) & 0xFFu;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (c[j]) atomicAdd(het + w * 4 + j, static_cast<unsigned long long>(c[j]));
}

// per-block partials of: min over all entries, min / max of the diagonal, max of the off-diagonal
__global__ void zhang_reduce_kernel(const double* __restrict__ K, int64_t n, double* __restrict__ part) {
  double mn = INFINITY, dmin = INFINITY, dmax = -INFINITY, omax = -INFINITY;
  const int64_t total = n * n;
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
    const double v = K[t];
    const int64_t j = t / n, i = t - j * n;
    mn = fmin(mn, v);
    if (i == j) {
      dmin = fmin(dmin, v);
      dmax = fmax(dmax, v);
    } else {
      omax = fmax(omax, v);
    }
  }
  __shared__ double red[4][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
    dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    omax = fmax(omax, __shfl_xor_sync(0xffffffffu, omax, o));
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    red[0][w] = mn;
    red[1][w] = dmin;
    red[2][w] = dmax;
    red[3][w] = omax;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) {
      red[0][0] = fmin(red[0][0], red[0][k]);
      red[1][0] = fmin(red[1][0], red[1][k]);
      red[2][0] = fmax(red[2][0], red[2][k]);
      red[3][0] = fmax(red[3][0], red[3][k]);
    }
    for (int k = 0; k < 4; ++k) part[blockIdx.x * 4 + k] = red[k][0];
  }
}

struct ZhangAdj {
  double top, floor, range;   // K <- top (K - floor) / range                       (:46)
  int adj_diag;               // Dmin < 1                                          (:50)
  double Dmin, diag_den, off_mul;   // diag <- (diag - Dmin + 1) / diag_den, off <- off * off_mul   (:52-53)
  int adj_off;                // Omax > top                                        (:58)
  double off_mul2;            // off <- off * (top / Omax)                         (:60)
};

__global__ void zhang_adjust_kernel(double* __restrict__ K, int64_t n, const ZhangAdj a) {
  const int64_t total = n * n;
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
    const int64_t j = t / n, i = t - j * n;
    double v = a.top * (K[t] - a.floor) / a.range;
    if (i == j) {
      if (a.adj_diag) v = (v - a.Dmin + 1.0) / a.diag_den;
    } else {
      if (a.adj_diag) v = v * a.off_mul;
      if (a.adj_off) v = v * a.off_mul2;
    }
    K[t] = v;
  }
}

// Splits of the marker axis.  Every job of a launch is equally long (one tile pair x kblocks / ks k-blocks), the units
// take them round by round, so the launch lasts ceil(jobs * ks / units) rounds of (kblocks / ks + a fixed cost per job:
// pipeline refill and the last accumulator drain).  Pick the ks with the shortest launch, each job keeping >= 32
// k-blocks; ties go to the smaller ks (fewer reduce-adds per tile, locally and over NVLink).  C2 (110 jobs, 74 units):
// 2 splits = 2.97 -> 3 rounds (99 % of the units busy), where "at least six waves" gave 5 splits = 7.4 -> 8 rounds (93 %).
static int choose_ksplit(int jobs, int kblocks, int units) {
  const double fixed = 4.0;
  int best = 1;
  double best_t = 0.0;
  for (int ks = 1; ks <= 128 && (ks == 1 || kblocks / ks >= 32); ++ks) {
    const double rounds = double((int64_t(jobs) * ks + units - 1) / units);
    const double t = rounds * (double(kblocks) / ks + fixed);
    if (ks == 1 || t < best_t * 0.97) {   // a finer split must pay for its extra reduce-adds
      best = ks;
      best_t = t;
    }
  }
  return best;
}

// the cut of one cross-product launch: pair jobs of the lower triangle (a cluster of 2 CTAs takes two block-rows against
// one block-column), k-blocks of the marker axis, marker splits per job
struct GramPlan {
  int tiles_side, ntri, kblocks, ksplit, units;
};
static GramPlan plan_gram(int64_t n, int64_t m, int sm_count) {
  constexpr int cl = 2;
  GramPlan p;
  p.tiles_side = static_cast<int>((n + 255) / 256);
  const int TR = (p.tiles_side + cl - 1) / cl;
  p.ntri = 0;
  for (int R = 0; R < TR; ++R) p.ntri += std::min(cl * (R + 1), p.tiles_side);
  p.kblocks = static_cast<int>(round_up(m, 128) / kBlockK);
  p.units = std::max(1, sm_count / cl);
  p.ksplit = choose_ksplit(p.ntri, p.kblocks, p.units);
  return p;
}

struct GramExchange {   // where the tiles of this launch are reduced into (world 1: the local buffer)
  int world = 1;
  int root = 0;           // >= 0: every tile goes to this rank's buffer; < 0: to the owner of its block-row
  int32_t* peers[8] = {};
};

// One form for every shape: the cta_group::2 kernel reads the marker-major matrix directly as MN-major operands (taxa
// contiguous, one marker per 128-byte row; no taxa-major copy, no transpose pass) and delivers each 256 x 256 tile as
// bulk tensor reduce-adds into the Gram buffer of the rank that owns it -- the local buffer, or a peer's over NVLink.
// A pair covers two block-rows; with a single block-row (n <= 256) the second CTA multiplies zero padding.
static int launch_gram(gapit_ctx* ctx, gapit_geno* g, int64_t ldg, const GramExchange& ex) {
  GramSched::Params sp;
  constexpr int cl = 2;
  const GramPlan gp = plan_gram(g->n, g->m, ctx->sm_count);
  sp.tiles_side = gp.tiles_side;
  if (int64_t(round_up(sp.tiles_side, cl)) * 256 > ldg) return fail(GAPIT_ERR_ARG, "launch_gram: Gram buffer too small");
  sp.cl = cl;
  sp.ntri = gp.ntri;
  sp.kblocks = gp.kblocks;
  sp.ksplit = gp.ksplit;
  const int nunits = std::min(ctx->sm_count / cl, sp.ntri * sp.ksplit);
  sp.round_ctr = nullptr;
  sp.round_wait = 100000;
  const size_t rounds = size_t((sp.ntri * sp.ksplit + nunits - 1) / nunits) + 1;
  GAPIT_CHECK(ctx_ws_t(ctx, WS_SYNC, rounds, &sp.round_ctr));
  GAPIT_CUDA(cudaMemsetAsync(sp.round_ctr, 0, rounds * sizeof(unsigned int), ctx->stream));
  // one map for both operands: [m markers][ldn bytes of taxa], boxes of 128 markers x 128 taxa
  CUtensorMap map_a;
  GAPIT_CHECK(make_tensor_map_u8(&map_a, g->mk, static_cast<uint64_t>(g->ldn), static_cast<uint64_t>(g->m),
                                 static_cast<uint64_t>(g->ldn), 128));
  using PCfg = UmmaPairCfg<2, 256, 4>;
  GramEpiTma::Params et{};
  et.world = ex.world;
  et.tiles_side = sp.tiles_side;
  et.root = ex.root;
  for (int r = 0; r < ex.world; ++r)
    GAPIT_CHECK(make_tensor_map_i32_box32(&et.maps[r], ex.peers[r], static_cast<uint64_t>(ldg), static_cast<uint64_t>(ldg),
                                          static_cast<uint64_t>(ldg) * 4));
  const size_t smem = PCfg::kSmemBytes + GramEpiTma::kSmemBytes + 16;
  auto kern = umma_i8_pair_kernel<2, 256, 4, 1, GramSched, GramEpiTma, 0, 0, true>;
  GAPIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  GAPIT_CUDA(launch_umma(kern, nunits, 2, PCfg::kThreads, smem, ctx->stream, map_a, map_a, sp, et));
  return GAPIT_OK;
}

// root pulls the block-rows the other ranks own (lower-triangle part) out of their buffers over NVLink:
// grid (row, owner-independent), 16-byte loads from the peer-mapped source
__global__ void gram_gather_rows_kernel(int32_t* __restrict__ dst, const int32_t* const* __restrict__ peers_dev, int64_t ldg,
                                        int tiles_side, int world, int self) {
  const int64_t row = blockIdx.x;
  const int I = static_cast<int>(row >> 8);
  if (I >= tiles_side) return;
  const int owner = min(world - 1, (I * world) / tiles_side);
  if (owner == self) return;
  const int64_t cols = int64_t(I + 1) * 256;   // lower triangle incl. the diagonal tile
  const uint4* s = reinterpret_cast<const uint4*>(peers_dev[owner] + row * ldg);
  uint4* d = reinterpret_cast<uint4*>(dst + row * ldg);
  for (int64_t c = threadIdx.x; c < cols / 4; c += blockDim.x) d[c] = s[c];
}

}  // namespace gapit

using namespace gapit;

extern "C" int64_t gapit_gram_ld(int64_t n) { return round_up(n, 512); }  // whole pairs of 256-row blocks

extern "C" int gapit_vanraden_gram(gapit_ctx* ctx, gapit_geno* g, int32_t* G_dev, int64_t* sums_dev) {
  if (!ctx || !g || !G_dev || !sums_dev) return fail(GAPIT_ERR_ARG, "gapit_vanraden_gram: NULL argument");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  const int64_t ldg = gapit_gram_ld(g->n);
  GAPIT_CUDA(cudaMemsetAsync(G_dev, 0, size_t(ldg) * ldg * 4, ctx->stream));
  GAPIT_CUDA(cudaMemsetAsync(sums_dev, 0, 3 * 8, ctx->stream));
  colsum_reduce_kernel<<<std::min<int64_t>(1024, (g->m + 255) / 256), 256, 0, ctx->stream>>>(
      g->colsum, g->m, 2 * g->n, reinterpret_cast<unsigned long long*>(sums_dev));
  GAPIT_CUDA(cudaGetLastError());
  cudaEventRecord(ctx->ev0, ctx->stream);
  GramExchange local;
  local.peers[0] = G_dev;
  GAPIT_CHECK(launch_gram(ctx, g, ldg, local));
  cudaEventRecord(ctx->ev1, ctx->stream);
  GAPIT_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  ctx->stats = gapit_stats{};
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  ctx->stats.kernel_ms = ms;
  return GAPIT_OK;
}

// Sharded kinship with the exchange fused into the Gram kernel's epilogue (GramEpiTma): every rank calls this with the
// same peer-mapped buffers (peers[r] = rank r's buffer), all zeroed and fenced by a barrier BEFORE anyone starts.  After
// the kernels of all ranks have completed (stream sync + barrier): root >= 0: rank `root` holds the complete lower
// triangle (reduce to root); root < 0: rank r holds the block-rows gapit_gram_owned_rows() gives it (reduce-scatter:
// every rank's NVLink ingress carries 1/world of the tiles), and gapit_vanraden_gram_gather pulls them together.
// sum_peers (may be NULL): sum_peers[r] = 3 int64 on rank r, peer-mapped and zeroed with the Gram buffers; this rank's
// scalar sums are added into every one of them, so after the closing barrier each holds the totals for gapit_vanraden_finish.
extern "C" int gapit_vanraden_gram_fused(gapit_ctx* ctx, gapit_geno* g, int32_t* const* peers, int world, int rank,
                                         int root, int64_t* sums_dev, int64_t* const* sum_peers) {
  if (!ctx || !g || !peers || !sums_dev || world < 1 || world > 8 || rank < 0 || rank >= world || root >= world)
    return fail(GAPIT_ERR_ARG, "gapit_vanraden_gram_fused: bad argument");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  const int64_t ldg = gapit_gram_ld(g->n);
  GramExchange ex;
  ex.world = world;
  ex.root = root;
  for (int r = 0; r < world; ++r) {
    if (!peers[r]) return fail(GAPIT_ERR_ARG, "gapit_vanraden_gram_fused: NULL peer buffer");
    ex.peers[r] = peers[r];
  }
  GAPIT_CUDA(cudaMemsetAsync(sums_dev, 0, 3 * 8, ctx->stream));
  colsum_reduce_kernel<<<std::min<int64_t>(1024, (g->m + 255) / 256), 256, 0, ctx->stream>>>(
      g->colsum, g->m, 2 * g->n, reinterpret_cast<unsigned long long*>(sums_dev));
  GAPIT_CUDA(cudaGetLastError());
  if (sum_peers) {
    SumPeers sp;
    for (int r = 0; r < 8; ++r) sp.p[r] = nullptr;
    for (int r = 0; r < world; ++r) {
      if (!sum_peers[r]) return fail(GAPIT_ERR_ARG, "gapit_vanraden_gram_fused: NULL sums mailbox");
      sp.p[r] = reinterpret_cast<unsigned long long*>(sum_peers[r]);
    }
    push_sums_kernel<<<1, 32, 0, ctx->stream>>>(reinterpret_cast<const unsigned long long*>(sums_dev), sp, world);
    GAPIT_CUDA(cudaGetLastError());
  }
  cudaEventRecord(ctx->ev0, ctx->stream);
  GAPIT_CHECK(launch_gram(ctx, g, ldg, ex));
  cudaEventRecord(ctx->ev1, ctx->stream);
  GAPIT_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  ctx->stats = gapit_stats{};
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  ctx->stats.kernel_ms = ms;
  return GAPIT_OK;
}

// After a reduce-scatter (root < 0 above) and the barrier that follows it: rank `rank` copies the block-rows owned by
// the other ranks out of their peer-mapped buffers into its own (lower triangle only), which then holds the complete
// sums gapit_vanraden_finish starts from.  One NVLink read of (world-1)/world of the triangle.
extern "C" int gapit_vanraden_gram_gather(gapit_ctx* ctx, int64_t n, int32_t* const* peers, int world, int rank) {
  if (!ctx || !peers || n <= 0 || world < 1 || world > 8 || rank < 0 || rank >= world)
    return fail(GAPIT_ERR_ARG, "gapit_vanraden_gram_gather: bad argument");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  if (world == 1) return GAPIT_OK;
  const int64_t ldg = gapit_gram_ld(n);
  const int tiles_side = static_cast<int>((n + 255) / 256);
  const int32_t** pd = nullptr;
  GAPIT_CHECK(ctx_ws_t(ctx, WS_SYNC, 16, &pd));
  GAPIT_CUDA(cudaMemcpyAsync(pd, peers, size_t(world) * sizeof(int32_t*), cudaMemcpyHostToDevice, ctx->stream));
  cudaEventRecord(ctx->ev0, ctx->stream);
  gram_gather_rows_kernel<<<static_cast<unsigned>(int64_t(tiles_side) * 256), 256, 0, ctx->stream>>>(
      peers[rank], pd, ldg, tiles_side, world, rank);
  GAPIT_CUDA(cudaGetLastError());
  cudaEventRecord(ctx->ev1, ctx->stream);
  GAPIT_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  ctx->stats.post_ms = ms;
  return GAPIT_OK;
}

// how a Gram launch of n taxa x m markers is cut on a GPU with `sm_count` SMs (host arithmetic only; what launch_gram uses)
extern "C" int gapit_gram_plan(int64_t n, int64_t m, int sm_count, int* jobs_out, int* splits_out, int* rounds_out) {
  if (n <= 0 || m <= 0 || sm_count < 2) return fail(GAPIT_ERR_ARG, "gapit_gram_plan: bad argument");
  const GramPlan p = plan_gram(n, m, sm_count);
  if (jobs_out) *jobs_out = p.ntri;
  if (splits_out) *splits_out = p.ksplit;
  if (rounds_out) *rounds_out = static_cast<int>((int64_t(p.ntri) * p.ksplit + p.units - 1) / p.units);
  return GAPIT_OK;
}

// block-rows [lo, hi) (units of 256 taxa rows) whose sums land on `rank` after a reduce-scatter
extern "C" int gapit_gram_owned_rows(int64_t n, int world, int rank, int64_t* row_lo, int64_t* row_hi) {
  if (n <= 0 || world < 1 || rank < 0 || rank >= world || !row_lo || !row_hi) return fail(GAPIT_ERR_ARG, "gapit_gram_owned_rows");
  const int64_t T = (n + 255) / 256;
  int64_t lo = T, hi = 0;
  for (int64_t I = 0; I < T; ++I) {
    const int64_t o = std::min<int64_t>(world - 1, (I * world) / T);
    if (o == rank) {
      lo = std::min(lo, I);
      hi = std::max(hi, I + 1);
    }
  }
  if (hi <= lo) lo = hi = 0;
  *row_lo = lo * 256;
  *row_hi = std::min<int64_t>(hi * 256, gapit_gram_ld(n));
  return GAPIT_OK;
}

extern "C" int gapit_vanraden_finish(gapit_ctx* ctx, int64_t n, int32_t* G_dev, const int64_t* sums_dev,
                                     double* K_dev_out, double* K_host_out, int32_t* G_host_out) {
  if (!ctx || !G_dev || !sums_dev || n <= 0) return fail(GAPIT_ERR_ARG, "gapit_vanraden_finish: bad argument");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  const int64_t ldg = gapit_gram_ld(n);
  long long* s = nullptr;
  double* K_dev = K_dev_out;
  long long* colpart = nullptr;
  GAPIT_CHECK(ctx_ws_t(ctx, WS_R, size_t(n) + 32, &s));                  // arena: no leak on the error paths below
  // the column sums per 32-row strip are consumed before the first K entry is written, so they live in the output
  // buffer itself (627 MB at n = 50,000; a separate allocation cost 1-12 ms inside the call); tiny n: the arena
  if (size_t((n + 31) / 32) * size_t(ldg) > size_t(n) * size_t(n) || (reinterpret_cast<uintptr_t>(K_dev_out) & 15))
    GAPIT_CHECK(ctx_ws_t(ctx, WS_PART, size_t((n + 31) / 32) * size_t(ldg), &colpart));
  if (!K_dev) GAPIT_CHECK(ctx_ws_t(ctx, WS_OUT, size_t(n) * n, &K_dev));
  if (!colpart) colpart = reinterpret_cast<long long*>(K_dev);
  {
    // everything from the lower triangle the Gram kernel (and the row gather of the sharded form) left: the row sums,
    // then both orientations of every K entry; the int32 matrix itself is mirrored only when the caller asks for it
    const unsigned t32 = static_cast<unsigned>((n + 31) / 32);
    gram_strip_sums_kernel<<<t32, 256, 0, ctx->stream>>>(G_dev, ldg, static_cast<int>(t32), s, colpart);
    gram_strip_combine_kernel<<<t32, 256, 0, ctx->stream>>>(colpart, ldg, static_cast<int>(t32), s);
    const long long t64 = (n + 63) / 64;
    vanraden_epilogue_lower_kernel<<<static_cast<unsigned>(t64 * (t64 + 1) / 2), 256, 0, ctx->stream>>>(
        G_dev, ldg, n, s, reinterpret_cast<const unsigned long long*>(sums_dev), K_dev);
    if (G_host_out) {
      dim3 gm(static_cast<unsigned>(ldg / 32), static_cast<unsigned>(ldg / 32));
      gram_mirror_kernel<<<gm, dim3(32, 8), 0, ctx->stream>>>(G_dev, ldg);
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && K_host_out)
    e = cudaMemcpyAsync(K_host_out, K_dev, size_t(n) * n * 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess && G_host_out)
    e = cudaMemcpy2DAsync(G_host_out, size_t(n) * 4, G_dev, size_t(ldg) * 4, size_t(n) * 4, n, cudaMemcpyDeviceToHost,
                          ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return fail(GAPIT_ERR_CUDA, std::string("gapit_vanraden_finish: ") + cudaGetErrorString(e));
  return GAPIT_OK;
}

extern "C" int gapit_vanraden(gapit_ctx* ctx, gapit_geno* g, double* K_out, int32_t* G_out) {
  if (!ctx || !g || !K_out) return fail(GAPIT_ERR_ARG, "gapit_vanraden: NULL argument");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  const int64_t ldg = gapit_gram_ld(g->n);
  int32_t* G_dev = nullptr;
  int64_t* sums = nullptr;
  GAPIT_CUDA(cudaMalloc(&G_dev, size_t(ldg) * ldg * 4));
  cudaError_t e = cudaMalloc(&sums, 3 * 8);
  if (e != cudaSuccess) {
    cudaFree(G_dev);
    return fail(GAPIT_ERR_CUDA, "gapit_vanraden: cudaMalloc");
  }
  int rc = gapit_vanraden_gram(ctx, g, G_dev, sums);
  if (rc == GAPIT_OK) rc = gapit_vanraden_finish(ctx, g->n, G_dev, sums, nullptr, K_out, G_out);
  cudaFree(G_dev);
  cudaFree(sums);
  return rc;
}

namespace gapit {
int syevd_dev(gapit_ctx* ctx, double* A_dev, int64_t n, double* w_dev);

// mirrored Gram + row sums + centred cross-product of every marker of `g` into K_dev (n x n column-major)
static int centered_gram_dev(gapit_ctx* ctx, gapit_geno* g, double* K_dev, int32_t** G_keep, int64_t** sums_keep) {
  const int64_t n = g->n, ldg = gapit_gram_ld(n);
  int32_t* G_dev = nullptr;
  int64_t* sums = nullptr;
  long long* s = nullptr;
  GAPIT_CUDA(cudaMalloc(&G_dev, size_t(ldg) * ldg * 4));
  if (cudaMalloc(&sums, 3 * 8) != cudaSuccess) {
    cudaFree(G_dev);
    return fail(GAPIT_ERR_CUDA, "centered_gram: cudaMalloc");
  }
  int rc = gapit_vanraden_gram(ctx, g, G_dev, sums);
  if (rc == GAPIT_OK) rc = ctx_ws_t(ctx, WS_R, size_t(n), &s);
  if (rc == GAPIT_OK) {
    dim3 grid(static_cast<unsigned>(ldg / 32), static_cast<unsigned>(ldg / 32));
    gram_mirror_kernel<<<grid, dim3(32, 8), 0, ctx->stream>>>(G_dev, ldg);
    gram_rowsum_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0, ctx->stream>>>(G_dev, ldg, n, s);
    dim3 g2(static_cast<unsigned>((n + 255) / 256), static_cast<unsigned>(std::min<int64_t>(n, 65535)));
    centered_gram_kernel<<<g2, 256, 0, ctx->stream>>>(G_dev, ldg, n, s, reinterpret_cast<const unsigned long long*>(sums), K_dev);
    if (cudaGetLastError() != cudaSuccess) rc = fail(GAPIT_ERR_CUDA, "centered_gram: launch failed");
  }
  if (rc == GAPIT_OK && G_keep) {
    *G_keep = G_dev;
    *sums_keep = sums;
    return rc;
  }
  cudaStreamSynchronize(ctx->stream);
  cudaFree(G_dev);
  cudaFree(sums);
  return rc;
}
}  // namespace gapit

// GAPIT.kinship.ZHANG(snps)  GAPIT.kinship.ZHANG.R:1-66
extern "C" int gapit_zhang(gapit_ctx* ctx, gapit_geno* g, double* K_out) {
  if (!ctx || !g || !K_out) return fail(GAPIT_ERR_ARG, "gapit_zhang: NULL argument");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  const int64_t n = g->n;
  double* K_dev = nullptr;
  long long* het = nullptr;
  double* part = nullptr;
  const int nblk = std::max(1, std::min(ctx->sm_count * 8, int((n * n + 255) / 256)));
  GAPIT_CHECK(ctx_ws_t(ctx, WS_OUT, size_t(n) * n, &K_dev));
  GAPIT_CHECK(ctx_ws_t(ctx, WS_PROJ, size_t(g->ldn), &het));
  GAPIT_CHECK(ctx_ws_t(ctx, WS_PART, size_t(nblk) * 4, &part));
  int32_t* G_dev = nullptr;
  int64_t* sums = nullptr;
  GAPIT_CHECK(centered_gram_dev(ctx, g, K_dev, &G_dev, &sums));
  {
    const int64_t words = g->ldn / 4;
    const int64_t slabs = std::max<int64_t>(1, std::min<int64_t>(1024, (int64_t(ctx->sm_count) * 8 * 128) / std::max<int64_t>(words, 1)));
    const int64_t rows_per_slab = (g->m + slabs - 1) / slabs;
    dim3 grid(static_cast<unsigned>((words + 127) / 128), static_cast<unsigned>((g->m + rows_per_slab - 1) / rows_per_slab));
    cudaMemsetAsync(het, 0, size_t(g->ldn) * 8, ctx->stream);
    taxa_het_kernel<<<grid, 128, 0, ctx->stream>>>(g->mk, g->ldn, g->m, rows_per_slab,
                                                  reinterpret_cast<unsigned long long*>(het));
  }
  zhang_reduce_kernel<<<nblk, 256, 0, ctx->stream>>>(K_dev, n, part);
  std::vector<long long> h_het(n);
  std::vector<double> h_part(size_t(nblk) * 4);
  std::vector<int32_t> h_cs(g->m);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(h_het.data(), het, size_t(n) * 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h_part.data(), part, h_part.size() * 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h_cs.data(), g->colsum, size_t(g->m) * 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(G_dev);
  cudaFree(sums);
  if (e != cudaSuccess) return fail(GAPIT_ERR_CUDA, std::string("gapit_zhang: ") + cudaGetErrorString(e));
  // :9-11 markers kept (0 < fa < 1); :14-17 inbreeding from the per-taxon heterozygosity
  int64_t nsnp = 0;
  for (int64_t k = 0; k < g->m; ++k) nsnp += (h_cs[k] > 0 && h_cs[k] < 2 * n);
  if (nsnp == 0) return fail(GAPIT_ERR_DATA, "gapit_zhang: every marker is monomorphic");
  double fmin_i = INFINITY;
  for (int64_t i = 0; i < n; ++i) fmin_i = std::min(fmin_i, static_cast<double>(h_het[i]) / (2.0 * static_cast<double>(nsnp)));
  const double inbreeding = 1.0 - fmin_i;
  double floor_ = INFINITY, DL = INFINITY, DU = -INFINITY, omax_raw = -INFINITY;
  for (int b = 0; b < nblk; ++b) {
    floor_ = std::min(floor_, h_part[b * 4 + 0]);
    DL = std::min(DL, h_part[b * 4 + 1]);
    DU = std::max(DU, h_part[b * 4 + 2]);
    omax_raw = std::max(omax_raw, h_part[b * 4 + 3]);
  }
  ZhangAdj a{};
  a.top = 1.0 + inbreeding;                                  // :45
  a.floor = floor_;
  a.range = DU - floor_;
  a.Dmin = a.top * (DL - floor_) / a.range;                  // :47
  double omax = a.top * (omax_raw - floor_) / a.range;       // the adjustments are monotone: Omax follows the raw maximum
  a.adj_diag = a.Dmin < 1.0;                                 // :50
  if (a.adj_diag) {
    a.diag_den = (a.top + 1.0 - a.Dmin) * .5;                // :52
    a.off_mul = 1.0 / a.Dmin;                                // :53
    omax = omax * a.off_mul;
  }
  a.adj_off = (n > 1) && omax > a.top;                       // :57-58
  if (a.adj_off) a.off_mul2 = a.top / omax;                  // :60
  zhang_adjust_kernel<<<nblk, 256, 0, ctx->stream>>>(K_dev, n, a);
  GAPIT_CUDA(cudaGetLastError());
  GAPIT_CUDA(cudaMemcpyAsync(K_out, K_dev, size_t(n) * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  GAPIT_CUDA(cudaStreamSynchronize(ctx->stream));
  return GAPIT_OK;
}

// prcomp(X) of GAPIT.PCA.R:10 from the centred cross-product: X_c X_c' = V diag(l) V'  =>  scores PCA.X$x = V sqrt(l),
// variances PCA.X$sdev^2 = l / (n - 1); min(n, m) components, descending.
extern "C" int gapit_pca(gapit_ctx* ctx, gapit_geno* g, int64_t ncomp, double* scores_out, double* ev_out) {
  if (!ctx || !g || ncomp < 0 || (ncomp > 0 && !scores_out)) return fail(GAPIT_ERR_ARG, "gapit_pca: bad argument");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  const int64_t n = g->n, rank = std::min(n, g->m);
  if (ncomp > rank) return fail(GAPIT_ERR_ARG, "gapit_pca: more components than min(n, m)");
  double *K_dev = nullptr, *w_dev = nullptr;
  GAPIT_CUDA(cudaMalloc(&K_dev, size_t(n) * n * 8));
  if (cudaMalloc(&w_dev, size_t(n) * 8) != cudaSuccess) {
    cudaFree(K_dev);
    return fail(GAPIT_ERR_CUDA, "gapit_pca: cudaMalloc");
  }
  int rc = centered_gram_dev(ctx, g, K_dev, nullptr, nullptr);
  if (rc == GAPIT_OK) rc = syevd_dev(ctx, K_dev, n, w_dev);   // ascending, vectors in place
  std::vector<double> w(n), col(n);
  cudaError_t e = cudaSuccess;
  if (rc == GAPIT_OK) e = cudaMemcpyAsync(w.data(), w_dev, size_t(n) * 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (rc == GAPIT_OK && e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  for (int64_t c = 0; rc == GAPIT_OK && e == cudaSuccess && c < rank; ++c) {
    const double l = std::max(w[n - 1 - c], 0.0);
    if (ev_out) ev_out[c] = l / static_cast<double>(n - 1);
    if (c < ncomp) {
      e = cudaMemcpy(col.data(), K_dev + (n - 1 - c) * n, size_t(n) * 8, cudaMemcpyDeviceToHost);
      const double sd = std::sqrt(l);
      for (int64_t i = 0; i < n; ++i) scores_out[c * n + i] = col[i] * sd;
    }
  }
  cudaFree(K_dev);
  cudaFree(w_dev);
  if (rc != GAPIT_OK) return rc;
  if (e != cudaSuccess) return fail(GAPIT_ERR_CUDA, std::string("gapit_pca: ") + cudaGetErrorString(e));
  return GAPIT_OK;
}
