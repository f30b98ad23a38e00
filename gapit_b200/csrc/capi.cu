// Context, error channel and library-level glue of the C ABI (include/gapitcuda.h).
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "umma_i8.cuh"

namespace gapit {

static thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

// cuTensorMapEncodeTiled lives in libcuda; resolve it through the runtime so the
// library links (and loads for the symbol checks) on a box without a driver.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tensor_map_u8(CUtensorMap* out, const void* base, uint64_t inner_bytes, uint64_t rows, uint64_t ld_bytes,
                       uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(GAPIT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (ld_bytes % 16 != 0 || (reinterpret_cast<uintptr_t>(base) & 127) != 0 || box_rows > 256)
    return fail(GAPIT_ERR_ARG, "make_tensor_map_u8: misaligned tensor");
  cuuint64_t dims[2] = {inner_bytes, rows};
  cuuint64_t strides[1] = {ld_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(GAPIT_ERR_CUDA, "cuTensorMapEncodeTiled failed, CUresult " + std::to_string(int(r)));
  return GAPIT_OK;
}

// 2-D tensor map over a row-major int32 matrix [rows][ld] with boxes of 32 x 32 elements (128-byte rows, 128-byte
// swizzle): the destination of the Gram epilogue's bulk reduce-adds (local or peer-mapped memory).
int make_tensor_map_i32_box32(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t ld_bytes) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(GAPIT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (ld_bytes % 16 != 0 || (reinterpret_cast<uintptr_t>(base) & 127) != 0)
    return fail(GAPIT_ERR_ARG, "make_tensor_map_i32_box32: misaligned tensor");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_bytes};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(GAPIT_ERR_CUDA, "cuTensorMapEncodeTiled (int32) failed, CUresult " + std::to_string(int(r)));
  return GAPIT_OK;
}

int ctx_ws(gapit_ctx* ctx, int slot, size_t bytes, void** out) {
  if (slot < 0 || slot >= gapit_ctx::kWsSlots) return fail(GAPIT_ERR_ARG, "ctx_ws: bad slot");
  if (bytes == 0) bytes = 16;
  if (ctx->ws_cap[slot] < bytes) {
    if (ctx->ws_ptr[slot]) {
      cudaStreamSynchronize(ctx->stream);
      cudaFree(ctx->ws_ptr[slot]);
      ctx->ws_ptr[slot] = nullptr;
      ctx->ws_cap[slot] = 0;
    }
    const size_t cap = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&ctx->ws_ptr[slot], cap);
    if (e != cudaSuccess) return fail(GAPIT_ERR_CUDA, std::string("workspace cudaMalloc: ") + cudaGetErrorString(e));
    ctx->ws_cap[slot] = cap;
  }
  *out = ctx->ws_ptr[slot];
  return GAPIT_OK;
}

}  // namespace gapit

using namespace gapit;

extern "C" const char* gapit_last_error(void) { return g_last_error.c_str(); }

extern "C" int gapit_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" int gapit_ctx_create(int device, gapit_ctx** out) {
  if (!out) return fail(GAPIT_ERR_ARG, "gapit_ctx_create: NULL out");
  int ndev = gapit_device_count();
  if (ndev <= 0) return fail(GAPIT_ERR_NO_DEVICE, "no CUDA device visible: libgapitcuda has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(GAPIT_ERR_ARG, "gapit_ctx_create: device index out of range");
  cudaDeviceProp prop;
  GAPIT_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(GAPIT_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                         "; libgapitcuda is built for sm_100a only");
  GAPIT_CUDA(cudaSetDevice(device));
  gapit_ctx* ctx = new gapit_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  cudaError_t e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
  ctx->stream = ctx->own_stream;
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev1);
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev2);
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev3);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_blk, cudaEventDisableTiming);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&ctx->ev_ready[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_free[i], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    gapit_ctx_destroy(ctx);
    return fail(GAPIT_ERR_CUDA, std::string("gapit_ctx_create: ") + cudaGetErrorString(e));
  }
  *out = ctx;
  return GAPIT_OK;
}

extern "C" void gapit_ctx_destroy(gapit_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  scan_plan_cache_drop(ctx);
  if (ctx->solver) cusolverDnDestroy(ctx->solver);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->ev2) cudaEventDestroy(ctx->ev2);
  if (ctx->ev3) cudaEventDestroy(ctx->ev3);
  if (ctx->ev_blk) cudaEventDestroy(ctx->ev_blk);
  for (int i = 0; i < 2; ++i) {
    if (ctx->ev_ready[i]) cudaEventDestroy(ctx->ev_ready[i]);
    if (ctx->ev_free[i]) cudaEventDestroy(ctx->ev_free[i]);
  }
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
  for (int i = 0; i < gapit_ctx::kWsSlots; ++i)
    if (ctx->ws_ptr[i]) cudaFree(ctx->ws_ptr[i]);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

extern "C" int gapit_ctx_stats(const gapit_ctx* ctx, gapit_stats* out) {
  if (!ctx || !out) return fail(GAPIT_ERR_ARG, "gapit_ctx_stats: NULL argument");
  *out = ctx->stats;
  return GAPIT_OK;
}

extern "C" int gapit_ctx_set_stream(gapit_ctx* ctx, void* cuda_stream) {
  if (!ctx) return fail(GAPIT_ERR_ARG, "gapit_ctx_set_stream: NULL ctx");
  ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
  if (ctx->solver) cusolverDnSetStream(ctx->solver, ctx->stream);
  return GAPIT_OK;
}

extern "C" int gapit_get_slices(const gapit_ctx* ctx) { return ctx ? ctx->slices : 0; }

extern "C" int gapit_set_slices(gapit_ctx* ctx, int slices) {
  if (!ctx) return fail(GAPIT_ERR_ARG, "gapit_set_slices: NULL ctx");
  if (slices == 0) slices = GAPIT_DEFAULT_SLICES;
  if (slices < 4 || slices > 8) return fail(GAPIT_ERR_ARG, "gapit_set_slices: slices must be 4..8");
  ctx->slices = slices;
  return GAPIT_OK;
}

// Host-side packer for the GAPIT_B2 upload format: R-layout double / integer / int8 dosages -> 2 bits per call, on the
// host's threads, so that a caller holding R doubles (8 bytes per call) ships 1/32 of the bytes over PCIe.
// NaN / NA_integer_ / negative int8 -> code 3 (NA); anything else outside {0,1,2} -> GAPIT_ERR_DATA.
#include <atomic>
#include <climits>
#include <thread>
#include <vector>

namespace {
template <typename T>
inline unsigned code_of(T v, bool* bad);
template <>
inline unsigned code_of<double>(double v, bool* bad) {
  if (v != v) return 3u;
  if (v == 0.0) return 0u;
  if (v == 1.0) return 1u;
  if (v == 2.0) return 2u;
  *bad = true;
  return 0u;
}
template <>
inline unsigned code_of<int32_t>(int32_t v, bool* bad) {
  if (v == INT_MIN) return 3u;
  if (v >= 0 && v <= 2) return static_cast<unsigned>(v);
  *bad = true;
  return 0u;
}
template <>
inline unsigned code_of<int8_t>(int8_t v, bool* bad) {
  if (v < 0) return 3u;
  if (v <= 2) return static_cast<unsigned>(v);
  *bad = true;
  return 0u;
}

// fast path: four calls per output byte, branch-free; any call that is not exactly 0, 1 or 2 sends the byte to the
// careful path (NA codes, error detection)
inline bool fast4(const double* s, unsigned* byte) {
  static const uint64_t pat[4] = {0ull, 0x3FF0000000000000ull, 0x4000000000000000ull, ~0ull};
  uint64_t b[4];
  memcpy(b, s, 32);
  unsigned out = 0u;
  uint64_t diff = 0ull;
  for (int c = 0; c < 4; ++c) {
    const unsigned hi = static_cast<unsigned>(b[c] >> 32);
    const unsigned code = ((hi >> 30) & 1u) * 2u + ((hi >> 29) & 1u);   // 0.0 -> 0, 1.0 -> 1, 2.0 -> 2
    diff |= b[c] ^ pat[code];
    out |= code << (2 * c);
  }
  *byte = out;
  return diff == 0ull;
}
inline bool fast4(const int32_t* s, unsigned* byte) {
  unsigned out = 0u, over = 0u;
  for (int c = 0; c < 4; ++c) {
    const unsigned v = static_cast<unsigned>(s[c]);
    over |= (v > 2u);
    out |= (v & 3u) << (2 * c);
  }
  *byte = out;
  return over == 0u;
}
inline bool fast4(const int8_t* s, unsigned* byte) {
  unsigned out = 0u, over = 0u;
  for (int c = 0; c < 4; ++c) {
    const unsigned v = static_cast<uint8_t>(s[c]);
    over |= (v > 2u);
    out |= (v & 3u) << (2 * c);
  }
  *byte = out;
  return over == 0u;
}

template <typename T>
bool pack_rows(const T* src, int64_t n, int64_t ld, int64_t k0, int64_t k1, uint8_t* out, int64_t nb) {
  bool bad = false;
  const int64_t full = n / 4;
  for (int64_t k = k0; k < k1; ++k) {
    const T* s = src + k * ld;
    uint8_t* o = out + k * nb;
    for (int64_t b = 0; b < nb; ++b) {
      unsigned byte = 0u;
      if (b >= full || !fast4(s + 4 * b, &byte)) {
        byte = 0u;
        for (int c = 0; c < 4 && 4 * b + c < n; ++c) byte |= code_of<T>(s[4 * b + c], &bad) << (2 * c);
      }
      o[b] = static_cast<uint8_t>(byte);
    }
  }
  return !bad;
}
}  // namespace

extern "C" int gapit_pack2_host(const void* snps, gapit_dtype dtype, int64_t n, int64_t m, int64_t ld, uint8_t* out,
                                int nthreads) {
  if (!snps || !out || n <= 0 || m <= 0 || ld < n || (dtype != GAPIT_F64 && dtype != GAPIT_I32 && dtype != GAPIT_I8))
    return fail(GAPIT_ERR_ARG, "gapit_pack2_host: bad argument");
  const int64_t nb = (n + 3) / 4;
  unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  int nt = nthreads > 0 ? nthreads : static_cast<int>(hw);
  nt = static_cast<int>(std::min<int64_t>(nt, std::max<int64_t>(1, m / 64)));
  std::atomic<int> ok{1};
  auto work = [&](int t) {
    const int64_t k0 = m * t / nt, k1 = m * (t + 1) / nt;
    bool good = true;
    if (dtype == GAPIT_F64) good = pack_rows(static_cast<const double*>(snps), n, ld, k0, k1, out, nb);
    else if (dtype == GAPIT_I32) good = pack_rows(static_cast<const int32_t*>(snps), n, ld, k0, k1, out, nb);
    else good = pack_rows(static_cast<const int8_t*>(snps), n, ld, k0, k1, out, nb);
    if (!good) ok.store(0);
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  if (!ok.load()) return fail(GAPIT_ERR_DATA, "genotype matrix holds a value outside {0,1,2}");
  return GAPIT_OK;
}
