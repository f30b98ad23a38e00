// Warp-specialised int8 tcgen05 main loop shared by the two hot kernels
// (VanRaden cross-product and the eigen-rotation of the P3D scan).
//
//   D[MSUB*128 x BN] (int32, TMEM) = A[MSUB*128 x K] (int8, K-major) * B[BN x K]^T (int8, K-major)
//
// One CTA per SM, persistent over a scheduler's job list:
//   warp 0      TMA producer  (one lane): 128-byte-swizzled boxes of A and B into a STAGES-deep smem ring
//   warp 1      UMMA issuer   (one lane): MSUB x 4 tcgen05.mma (M=128, N=BN, K=32) per ring slot,
//                                         tcgen05.commit releases the slot / publishes the accumulator
//   warp 2      TMEM allocator (512 columns)
//   warps 4..   epilogue (128 threads per 128-row sub-tile): tcgen05.ld of the int32 accumulator,
//               kernel-specific reduction, then hands the TMEM stage back
// With MSUB == 1 the accumulator is double-buffered in TMEM (2 x 256 columns) so the epilogue of job i
// overlaps the MMAs of job i+1; with MSUB == 2 both 256-column halves hold one 256 x BN tile
// (operand bytes per MMA drop from 96 to 64 B/clk/SM, the epilogue is exposed once per job).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "ptx.cuh"

namespace gapit {

constexpr int kBlockK = 128;  // int8 elements per smem row: exactly one 128-byte swizzle atom
constexpr int kUmmaK = 32;    // K of one kind::i8 instruction
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;  // TMEM columns reserved per 128-row accumulator

struct Job {
  int a_row0;    // first row of the A box (MSUB*128 rows of the A tensor)
  int b_row0;    // first row of the B box (BN rows of the B tensor)
  int kb_begin;  // K range in units of kBlockK
  int kb_end;
  int tag0;  // scheduler payload for the epilogue
  int tag1;
  int tag2;
  bool group_first;  // first / last job of a run whose epilogue state is carried in registers
  bool group_last;
};

template <int MSUB, int BN, int STAGES>
struct UmmaCfg {
  static_assert(MSUB == 1 || MSUB == 2, "one or two 128-row sub-tiles");
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128");
  static constexpr int kAccStages = (MSUB == 1) ? 2 : 1;
  static constexpr int kABytes = MSUB * 128 * kBlockK;
  static constexpr int kBBytes = BN * kBlockK;
  static_assert(kBBytes % 1024 == 0, "B stage must keep 1024-byte alignment for SWIZZLE_128B");
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarBytes = (2 * STAGES + 2 * kAccStages) * 8 + 16;
  static constexpr int kSmemBytes = STAGES * kStageBytes + kBarBytes + 1024;  // +1024: manual alignment slack
  static constexpr int kThreads = 128 + MSUB * 128;
};

// Sched: struct with `Params`, ctor(Params, unit, nunits, cta_rank), bool next(Job&).
// Epi  : struct with `Params`, `kSmemBytes`, ctor(Params, smem scratch), prepare(job, tid, nthreads),
//        run(job, taddr, sub, row, half, nhalf), finish() (after the last job, before the CTA may exit).
//        The parameter block is a __grid_constant__, so an epilogue may hand tensor maps inside it to TMA.
//
template <int MSUB, int BN, int STAGES, int EPIW, class Sched, class Epi, uint32_t A_SIGNED, uint32_t B_SIGNED>
__global__ void __launch_bounds__(128 + MSUB * 128 * EPIW, 1)
    umma_i8_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const typename Sched::Params sp, const __grid_constant__ typename Epi::Params ep) {
  using Cfg = UmmaCfg<MSUB, BN, STAGES>;
  const uint32_t crank = 0u;
  const int unit = blockIdx.x;
  const int nunits = gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = tfull + Cfg::kAccStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + Cfg::kAccStages);
  // epilogue scratch (Epi::kSmemBytes, 16-byte aligned) behind the barriers; the launch adds it to the dynamic size
  uint8_t* epi_smem = smem + ((STAGES * Cfg::kStageBytes + Cfg::kBarBytes + 15) & ~15);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_a);
    ptx::prefetch_tensormap(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < Cfg::kAccStages; ++i) {
      ptx::mbar_init(&tfull[i], 1);
      ptx::mbar_init(&tempty[i], MSUB * 128 * EPIW);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc(tmem_ptr, kTmemCols);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      Sched sch(sp, unit, nunits, crank);
      Job job;
      uint32_t st = 0, ph = 0;
      while (sch.next(job)) {
        for (int kb = job.kb_begin; kb < job.kb_end; ++kb) {
          ptx::mbar_wait(&empty[st], ph ^ 1);
          uint8_t* sa = smem + st * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          ptx::mbar_arrive_expect_tx(&full[st], Cfg::kStageBytes);
          ptx::tma_load_2d(&map_a, &full[st], sa, kb * kBlockK, job.a_row0, sch.hint_a());
          ptx::tma_load_2d(&map_b, &full[st], sb, kb * kBlockK, job.b_row0, sch.hint_b());
          if (++st == STAGES) {
            st = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- UMMA issuer
    constexpr uint32_t idesc = ptx::idesc_i8(128, BN, A_SIGNED, B_SIGNED);
    Sched sch(sp, unit, nunits, crank);
    Job job;
    uint32_t st = 0, ph = 0, it = 0;
    while (sch.next(job)) {
      const uint32_t acc = it % Cfg::kAccStages;
      const uint32_t accph = (it / Cfg::kAccStages) & 1;
      ptx::mbar_wait(&tempty[acc], accph ^ 1);
      ptx::tc_fence_after_sync();
      for (int kb = job.kb_begin; kb < job.kb_end; ++kb) {
        ptx::mbar_wait(&full[st], ph);
        ptx::tc_fence_after_sync();
        if (lane == 0) {
          const uint32_t a_base = ptx::smem_u32(smem + st * Cfg::kStageBytes);
          const uint32_t b_base = a_base + Cfg::kABytes;
#pragma unroll
          for (int s = 0; s < MSUB; ++s) {
            const uint32_t d_col = tmem_base + (acc * MSUB + s) * kAccStride;
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              const uint64_t da = ptx::smem_desc_k_sw128(a_base + s * 128 * kBlockK + k * kUmmaK);
              const uint64_t db = ptx::smem_desc_k_sw128(b_base + k * kUmmaK);
              ptx::umma_i8_ss(d_col, da, db, idesc, (kb > job.kb_begin || k > 0) ? 1u : 0u);
            }
          }
          ptx::umma_commit(&empty[st]);
          if (kb == job.kb_end - 1) ptx::umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++st == STAGES) {
          st = 0;
          ph ^= 1;
        }
      }
      ++it;
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue
    const int sub = ((warp - 4) >> 2) % MSUB;
    const int half = ((warp - 4) >> 2) / MSUB;  // which share of the accumulator columns this warp drains (EPIW > 1)
    const int quarter = warp & 3;  // TMEM lane quarter this warp may touch
    const int row = quarter * 32 + lane;
    Epi epi(ep, epi_smem);
    Sched sch(sp, unit, nunits, crank);
    Job job;
    uint32_t it = 0;
    constexpr int kEpiThreads = MSUB * 128 * EPIW;
    while (sch.next(job)) {
      const uint32_t acc = it % Cfg::kAccStages;
      const uint32_t accph = (it / Cfg::kAccStages) & 1;
      if (Epi::kSmemBytes > 0) {
        // stage the job's epilogue constants while its MMAs run: nobody may still read the previous job's copy
        ptx::named_bar_sync(1, kEpiThreads);
        epi.prepare(job, threadIdx.x - 128, kEpiThreads);
        ptx::named_bar_sync(1, kEpiThreads);
      }
      ptx::mbar_wait(&tfull[acc], accph);
      ptx::tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + (acc * MSUB + sub) * kAccStride;
      epi.run(job, taddr, sub, row, half, EPIW);
      ptx::tc_fence_before_sync();
      ptx::mbar_arrive(&tempty[acc]);
      ++it;
    }
    epi.finish();
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------------------------
// Pair form (tcgen05 cta_group::2).  Two CTAs on the SM pair of one TPC run as a cluster and issue ONE UMMA of
// M = 256 per 128-row sub-tile: each CTA stages its own 128 x MSUB rows of A and only HALF of the B rows
// (BN/2) in shared memory; the tensor cores of both SMs read their own A and exchange the B halves over the
// pair datapath.  Per SM this halves the shared-memory reads of B and the TMA fill of B, which is what bounds
// the single-CTA form (operand reads 96 B/clk + TMA fill 64 B/clk against 128 B/clk of shared-memory bandwidth).
//   producer (both CTAs): TMA with .cta_group::2, bytes complete on the LEADER's `full` barrier
//   UMMA issuer (leader only): tcgen05.mma.cta_group::2; tcgen05.commit multicast frees the slot in both CTAs
//                              and publishes the accumulator to both epilogues
//   epilogue (both CTAs): each drains its own 128 x MSUB TMEM lanes, then arrives on the leader's `tempty`
template <int MSUB, int BN, int STAGES>
struct UmmaPairCfg {
  static_assert(MSUB == 1 || MSUB == 2, "one or two 128-row sub-tiles per CTA");
  static_assert(BN % 32 == 0 || (BN / 2) % 8 == 0, "half of B must keep whole 8-row swizzle groups");
  static_assert(BN % 16 == 0 && BN >= 32 && BN <= 256, "UMMA N for M=256");
  static constexpr int kAccStages = (MSUB == 1) ? 2 : 1;
  static constexpr int kABytes = MSUB * 128 * kBlockK;
  static constexpr int kBBytes = (BN / 2) * kBlockK;
  static_assert(kBBytes % 1024 == 0, "B stage must keep 1024-byte alignment for SWIZZLE_128B");
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarBytes = (2 * STAGES + 2 * kAccStages) * 8 + 16;
  static constexpr int kSmemBytes = STAGES * kStageBytes + kBarBytes + 1024;
  static constexpr int kThreads = 128 + MSUB * 128;
};

// EPIW epilogue warps share each (sub-tile, lane quarter): they split the accumulator columns between them and
// halve the time the single TMEM stage is blocked (the epilogue keeps per-warp partial sums).
// MNMAJ: both operands are MN-major (the contracted index is the ROW of the global matrix): a ring slot holds, per
// 128-row operand chunk, a [kBlockK contracted rows][128 bytes] box.  job.a_row0 / b_row0 are then byte offsets along
// the rows and kb walks the row axis.  Needs BN == 256 (one 128-wide chunk of B per CTA).
template <int MSUB, int BN, int STAGES, int EPIW, class Sched, class Epi, uint32_t A_SIGNED, uint32_t B_SIGNED,
          bool MNMAJ = false>
__global__ void __launch_bounds__(128 + MSUB * 128 * EPIW, 1)
    umma_i8_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                        const typename Sched::Params sp, const __grid_constant__ typename Epi::Params ep) {
  using Cfg = UmmaPairCfg<MSUB, BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* full = bars;                          // used on the leader: both producers arrive, bytes of both CTAs
  uint64_t* empty = bars + STAGES;                // per CTA: freed by the leader's multicast commit
  uint64_t* tfull = bars + 2 * STAGES;            // per CTA: accumulator ready (multicast commit)
  uint64_t* tempty = tfull + Cfg::kAccStages;     // used on the leader: both epilogues arrive
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + Cfg::kAccStages);
  // epilogue scratch (Epi::kSmemBytes, 16-byte aligned) behind the barriers; the launch adds it to the dynamic size
  uint8_t* epi_smem = smem + ((STAGES * Cfg::kStageBytes + Cfg::kBarBytes + 15) & ~15);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = ptx::cluster_ctarank();
  const bool leader = (crank == 0);
  const int unit = blockIdx.x / 2;
  const int nunits = gridDim.x / 2;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_a);
    ptx::prefetch_tensormap(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full[i], 2);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < Cfg::kAccStages; ++i) {
      ptx::mbar_init(&tfull[i], 1);
      ptx::mbar_init(&tempty[i], 2 * MSUB * 128 * EPIW);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc_pair(tmem_ptr, kTmemCols);
  ptx::tc_fence_before_sync();
  ptx::cluster_sync();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      Sched sch(sp, unit, nunits, crank);
      Job job;
      uint32_t st = 0, ph = 0;
      while (sch.next(job)) {
        if (leader && job.group_first) sch.align_round();   // keep the pairs that share an operand window in step
        for (int kb = job.kb_begin; kb < job.kb_end; ++kb) {
          ptx::mbar_wait(&empty[st], ph ^ 1);
          uint8_t* sa = smem + st * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if (leader) ptx::mbar_arrive_expect_tx(&full[st], 2 * Cfg::kStageBytes);
          else ptx::mbar_arrive_remote(&full[st], 0);
          if (MNMAJ) {
            static_assert(!MNMAJ || BN == 256, "MN-major B: one 128-wide chunk per CTA");
#pragma unroll
            for (int s = 0; s < MSUB; ++s)
              ptx::tma_load_2d_pair(&map_a, &full[st], sa + s * 128 * kBlockK, job.a_row0 + s * 128, kb * kBlockK,
                                    sch.hint_a());
            ptx::tma_load_2d_pair(&map_b, &full[st], sb, job.b_row0 + static_cast<int>(crank) * (BN / 2), kb * kBlockK,
                                  sch.hint_b());
          } else {
            ptx::tma_load_2d_pair(&map_a, &full[st], sa, kb * kBlockK, job.a_row0, sch.hint_a());
            ptx::tma_load_2d_pair(&map_b, &full[st], sb, kb * kBlockK, job.b_row0 + static_cast<int>(crank) * (BN / 2),
                                  sch.hint_b());
          }
          if (++st == STAGES) {
            st = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- UMMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = ptx::idesc_i8(256, BN, A_SIGNED, B_SIGNED, MNMAJ ? 1u : 0u);
      Sched sch(sp, unit, nunits, crank);
      Job job;
      uint32_t st = 0, ph = 0, it = 0;
      while (sch.next(job)) {
        const uint32_t acc = it % Cfg::kAccStages;
        const uint32_t accph = (it / Cfg::kAccStages) & 1;
        ptx::mbar_wait(&tempty[acc], accph ^ 1);
        ptx::tc_fence_after_sync();
        for (int kb = job.kb_begin; kb < job.kb_end; ++kb) {
          ptx::mbar_wait(&full[st], ph);
          ptx::tc_fence_after_sync();
          if (lane == 0) {
            const uint32_t a_base = ptx::smem_u32(smem + st * Cfg::kStageBytes);
            const uint32_t b_base = a_base + Cfg::kABytes;
#pragma unroll
            for (int s = 0; s < MSUB; ++s) {
              const uint32_t d_col = tmem_base + (acc * MSUB + s) * kAccStride;
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                // K-major: 32 bytes along the 128-byte rows per step; MN-major: 32 rows of 128 bytes per step
                const uint64_t da = MNMAJ ? ptx::smem_desc_mn_sw128(a_base + s * 128 * kBlockK + k * kUmmaK * 128, 128 * kBlockK)
                                          : ptx::smem_desc_k_sw128(a_base + s * 128 * kBlockK + k * kUmmaK);
                const uint64_t db = MNMAJ ? ptx::smem_desc_mn_sw128(b_base + k * kUmmaK * 128, 128 * kBlockK)
                                          : ptx::smem_desc_k_sw128(b_base + k * kUmmaK);
                ptx::umma_i8_ss_pair(d_col, da, db, idesc, (kb > job.kb_begin || k > 0) ? 1u : 0u);
              }
            }
            ptx::umma_commit_pair_multicast(&empty[st], 3);
            if (kb == job.kb_end - 1) ptx::umma_commit_pair_multicast(&tfull[acc], 3);
          }
          __syncwarp();
          if (++st == STAGES) {
            st = 0;
            ph ^= 1;
          }
        }
        ++it;
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue (both CTAs)
    const int sub = ((warp - 4) >> 2) % MSUB;
    const int half = ((warp - 4) >> 2) / MSUB;   // which share of the accumulator columns this warp drains
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    Epi epi(ep, epi_smem);
    Sched sch(sp, unit, nunits, crank);
    Job job;
    uint32_t it = 0;
    constexpr int kEpiThreads = MSUB * 128 * EPIW;
    while (sch.next(job)) {
      const uint32_t acc = it % Cfg::kAccStages;
      const uint32_t accph = (it / Cfg::kAccStages) & 1;
      if (Epi::kSmemBytes > 0) {
        // stage the job's epilogue constants while its MMAs run: nobody may still read the previous job's copy
        ptx::named_bar_sync(1, kEpiThreads);
        epi.prepare(job, threadIdx.x - 128, kEpiThreads);
        ptx::named_bar_sync(1, kEpiThreads);
      }
      ptx::mbar_wait(&tfull[acc], accph);
      ptx::tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + (acc * MSUB + sub) * kAccStride;
      epi.run(job, taddr, sub, row, half, EPIW);
      ptx::tc_fence_before_sync();
      if (leader) ptx::mbar_arrive(&tempty[acc]);
      else ptx::mbar_arrive_remote(&tempty[acc], 0);
      ++it;
    }
    epi.finish();
  }

  ptx::tc_fence_before_sync();
  ptx::cluster_sync();
  if (warp == 2) ptx::tmem_dealloc_pair(tmem_base, kTmemCols);
}

// Host side: 2-D tensor map over a row-major uint8 matrix [rows][ld] with boxes of
// (box_rows x 128 bytes), 128-byte swizzle, out-of-bounds rows read as zero.
int make_tensor_map_u8(CUtensorMap* out, const void* base, uint64_t inner_bytes, uint64_t rows, uint64_t ld_bytes,
                       uint32_t box_rows);

// Launch helper: grid of `nunits * CL` CTAs as clusters of CL (CL == 1: a plain launch).
template <class Kernel, class... Args>
inline cudaError_t launch_umma(Kernel kern, int nunits, int cl, int threads, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(nunits * cl));
  cfg.blockDim = dim3(static_cast<unsigned>(threads));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cl);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

}  // namespace gapit
