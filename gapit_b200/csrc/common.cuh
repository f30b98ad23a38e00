// Shared host-side plumbing of libgapitcuda: error capture, the context and
// genotype handles behind the C ABI (include/gapitcuda.h).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/gapitcuda.h"

namespace gapit {

int fail(int code, const std::string& msg);

#define GAPIT_CUDA(call)                                                                          \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      return ::gapit::fail(GAPIT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__) +  \
                                               " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
    }                                                                                             \
  } while (0)

#define GAPIT_CHECK(call)          \
  do {                             \
    int rc__ = (call);             \
    if (rc__ != GAPIT_OK) return rc__; \
  } while (0)

inline int64_t round_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
// host dtypes that hold 2 bits per genotype call (rows of ceil(n/4) bytes), and the device recode each needs
inline bool is_packed2(gapit_dtype d) { return d == GAPIT_B2 || d == GAPIT_BED_A1 || d == GAPIT_BED_A2; }
inline int packed2_recode(gapit_dtype d) { return d == GAPIT_BED_A1 ? 1 : (d == GAPIT_BED_A2 ? 2 : 0); }

}  // namespace gapit

namespace gapit {
struct ScanPlan;
}

// Opaque handles of the C ABI.
struct gapit_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  cusolverDnHandle_t solver = nullptr;
  // timings of the last call, ms (CUDA events on `stream`)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  gapit_stats stats = {};
  int slices = GAPIT_DEFAULT_SLICES;  // int8 digit planes of V used by the rotation (Ozaki split)
  // grow-only device workspaces reused across calls (no cudaMalloc/cudaFree on the hot path)
  static constexpr int kWsSlots = 32;
  void* ws_ptr[kWsSlots] = {};
  size_t ws_cap[kWsSlots] = {};
  cudaStream_t copy_stream = nullptr;   // H2D of the next genotype chunk while the current one is scanned
  cudaEvent_t ev2 = nullptr, ev3 = nullptr;
  cudaEvent_t ev_ready[2] = {}, ev_free[2] = {};
  // the marker-free part of the last scan (reduced models, epilogue constants, back-rotated columns), kept so that
  // scanning block after block / fragment after fragment with the same traits does not redo it (scan.cu)
  gapit::ScanPlan* plan_cache = nullptr;
  uint64_t plan_key = 0;
  cudaStream_t d2h_stream = nullptr;   // finished result rows go to the host while the next marker block is computed
  cudaEvent_t ev_blk = nullptr;        // results of a marker block are final
};

namespace gapit {
enum WsSlot { WS_R = 0, WS_PROJ, WS_GCONST, WS_PART, WS_OUT, WS_RG, WS_RS, WS_RSCALE, WS_BAD, WS_CHUNK0, WS_CHUNK1,
              WS_CS0, WS_CS1, WS_CQ0, WS_CQ1, WS_STAGE0, WS_STAGE1, WS_SYNC, WS_FC, WS_WBUF, WS_WT, WS_XX, WS_KSCALE, WS_Z };
// device workspace of at least `bytes` in `slot` (contents are not preserved when it grows)
int ctx_ws(gapit_ctx* ctx, int slot, size_t bytes, void** out);
void scan_plan_cache_drop(gapit_ctx* ctx);   // scan.cu
template <class T>
inline int ctx_ws_t(gapit_ctx* ctx, int slot, size_t count, T** out) {
  void* p = nullptr;
  int rc = ctx_ws(ctx, slot, count * sizeof(T), &p);
  *out = static_cast<T*>(p);
  return rc;
}
}  // namespace gapit

struct gapit_geno {
  gapit_ctx* ctx = nullptr;
  int device = 0;         // for the free path: the handle may outlive its context
  bool borrowed = false;  // mk belongs to the caller (gapit_geno_wrap_device)
  int64_t n = 0;      // taxa
  int64_t m = 0;      // markers
  int64_t ldn = 0;    // bytes per marker row of the marker-major copy (multiple of 128)
  int8_t* mk = nullptr;   // marker-major  [m][ldn]: K-major operand of the rotation (contract over taxa) and MN-major
                          // operand of the kinship cross-product (contract over markers)
  int32_t* colsum = nullptr;    // [m] sum_i x_ik
  int32_t* colsumsq = nullptr;  // [m] sum_i x_ik^2
};
