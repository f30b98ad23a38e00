// Spectral pieces of the P3D path (emma.txt:58-61 and :82-89).  The O(n^3) work is cuSOLVER's
// syevd (a library call, timed separately from the hot-path roofline); this file only forms
// S (K+I) S without the reference's two n^3 GEMMs and puts the spectrum in R's order.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>
#include <vector>

#include <dlfcn.h>

#include "common.cuh"

namespace gapit {

static int ensure_solver(gapit_ctx* ctx) {
  if (ctx->solver) return GAPIT_OK;
  if (cusolverDnCreate(&ctx->solver) != CUSOLVER_STATUS_SUCCESS) return fail(GAPIT_ERR_CUDA, "cusolverDnCreate failed");
  if (cusolverDnSetStream(ctx->solver, ctx->stream) != CUSOLVER_STATUS_SUCCESS)
    return fail(GAPIT_ERR_CUDA, "cusolverDnSetStream failed");
  return GAPIT_OK;
}

// out[c*n + k] = sum_j A[j + k n] R[j + c n]   (A symmetric => (A R)[k][c]); one warp per k
__global__ void sym_times_cols_kernel(const double* __restrict__ A, int64_t n, const double* __restrict__ R, int nc,
                                      double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t k = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (k >= n) return;
  const double* a = A + k * n;
  for (int c = 0; c < nc; ++c) {
    double acc = 0.0;
    for (int64_t j = lane; j < n; j += 32) acc = fma(a[j], __ldg(R + c * n + j), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[c * n + k] = acc;
  }
}

// A = S M S with M = K + I, S = I - X (X'X)^-1 X', expanded as a rank-2q correction:
//   A_ij = M_ij - sum_l ( Y1_il B_jl + B_il Y1_jl - Y2_il Y1_jl ),  B = M X, Y1 = X (X'X)^-1, Y2 = Y1 (X'B)
// only the lower triangle (i >= j) is needed by syevd; the full matrix is written anyway.
__global__ void form_sks_kernel(const double* __restrict__ K, int64_t n, const double* __restrict__ Y1,
                                const double* __restrict__ B, const double* __restrict__ Y2, int q,
                                double* __restrict__ A) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int64_t j = blockIdx.y; j < n; j += gridDim.y) {
    double v = K[i + j * n] + (i == j ? 1.0 : 0.0);
    for (int l = 0; l < q; ++l) {
      const double y1i = Y1[l * n + i], y1j = Y1[l * n + j];
      v -= y1i * B[l * n + j] + B[l * n + i] * y1j - Y2[l * n + i] * y1j;
    }
    A[i + j * n] = v;
  }
}

// column c of dst <- column (n-1-c) of src, for c < ncols
__global__ void reverse_cols_kernel(const double* __restrict__ src, int64_t n, int64_t ncols, double* __restrict__ dst) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int64_t c = blockIdx.y; c < ncols; c += gridDim.y) dst[i + c * n] = src[i + (n - 1 - c) * n];
}

// column c of dst <- column (k-1-c) of src, for c < k (the first k columns, reversed)
__global__ void reverse_first_cols_kernel(const double* __restrict__ src, int64_t n, int64_t k, double* __restrict__ dst) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int64_t c = blockIdx.y; c < k; c += gridDim.y) dst[i + c * n] = src[i + (k - 1 - c) * n];
}

// column c <-> column n-1-c, in place (ascending -> descending order of the eigenvectors)
__global__ void swap_cols_kernel(double* __restrict__ A, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int64_t c = blockIdx.y; c < n / 2; c += gridDim.y) {
    const double a = A[i + c * n], b = A[i + (n - 1 - c) * n];
    A[i + c * n] = b;
    A[i + (n - 1 - c) * n] = a;
  }
}

// in-place syevd on a device matrix; eigenvalues (ascending) into w_dev.  The 64-bit API (cusolverDnXsyevd) has no
// 32-bit workspace limit on n.
int syevd_dev(gapit_ctx* ctx, double* A_dev, int64_t n, double* w_dev) {
  GAPIT_CHECK(ensure_solver(ctx));
  int* info = nullptr;
  void* work = nullptr;
  std::vector<char> hwork;
  cusolverDnParams_t params = nullptr;
  GAPIT_CUDA(cudaMalloc(&info, 4));
  size_t dbytes = 0, hbytes = 0;
  cusolverStatus_t st = cusolverDnCreateParams(&params);
  if (st == CUSOLVER_STATUS_SUCCESS)
    st = cusolverDnXsyevd_bufferSize(ctx->solver, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F,
                                     A_dev, n, CUDA_R_64F, w_dev, CUDA_R_64F, &dbytes, &hbytes);
  if (st == CUSOLVER_STATUS_SUCCESS && cudaMalloc(&work, std::max<size_t>(dbytes, 16)) != cudaSuccess)
    st = CUSOLVER_STATUS_ALLOC_FAILED;
  if (st == CUSOLVER_STATUS_SUCCESS) {
    hwork.resize(std::max<size_t>(hbytes, 16));
    st = cusolverDnXsyevd(ctx->solver, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A_dev, n,
                          CUDA_R_64F, w_dev, CUDA_R_64F, work, dbytes, hwork.data(), hbytes, info);
  }
  int hinfo = -1;
  cudaError_t e = cudaMemcpyAsync(&hinfo, info, 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(work);
  cudaFree(info);
  if (params) cusolverDnDestroyParams(params);
  if (st != CUSOLVER_STATUS_SUCCESS) return fail(GAPIT_ERR_CUDA, "cuSOLVER syevd failed, status " + std::to_string(int(st)));
  if (e != cudaSuccess) return fail(GAPIT_ERR_CUDA, std::string("eigh: ") + cudaGetErrorString(e));
  if (hinfo != 0) return fail(GAPIT_ERR_NUMERIC, "syevd did not converge, info " + std::to_string(hinfo));
  return GAPIT_OK;
}

// ascending -> descending order of the eigenvectors, in place (also used by spectral_mg.cu)
int launch_swap_cols(cudaStream_t st, double* A_dev, int64_t n) {
  if (n < 2) return GAPIT_OK;
  dim3 grid(static_cast<unsigned>((n + 255) / 256), static_cast<unsigned>(std::min<int64_t>(n / 2, 65535)));
  swap_cols_kernel<<<grid, 256, 0, st>>>(A_dev, n);
  GAPIT_CUDA(cudaGetLastError());
  return GAPIT_OK;
}

// shared tail: A_dev holds eigenvectors (ascending); emit the first `keep` in descending order
static int emit_desc(gapit_ctx* ctx, double* A_dev, double* w_dev, int64_t n, int64_t keep, double shift,
                     double* values_out, double* vectors_out) {
  double* rev = nullptr;
  GAPIT_CUDA(cudaMalloc(&rev, size_t(n) * keep * 8));
  dim3 grid(static_cast<unsigned>((n + 255) / 256), static_cast<unsigned>(std::min<int64_t>(keep, 65535)));
  reverse_cols_kernel<<<grid, 256, 0, ctx->stream>>>(A_dev, n, keep, rev);
  std::vector<double> w(n);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(w.data(), w_dev, size_t(n) * 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(vectors_out, rev, size_t(n) * keep * 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(rev);
  if (e != cudaSuccess) return fail(GAPIT_ERR_CUDA, std::string("eigh output: ") + cudaGetErrorString(e));
  for (int64_t c = 0; c < keep; ++c) values_out[c] = w[n - 1 - c] + shift;
  return GAPIT_OK;
}

}  // namespace gapit

using namespace gapit;

extern "C" int gapit_eigh(gapit_ctx* ctx, const double* A, int64_t n, double* values_out, double* vectors_out) {
  if (!ctx || !A || !values_out || !vectors_out || n <= 0) return fail(GAPIT_ERR_ARG, "gapit_eigh: bad argument");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  double *A_dev = nullptr, *w_dev = nullptr;
  GAPIT_CUDA(cudaMalloc(&A_dev, size_t(n) * n * 8));
  cudaError_t e = cudaMalloc(&w_dev, size_t(n) * 8);
  if (e == cudaSuccess) e = cudaMemcpyAsync(A_dev, A, size_t(n) * n * 8, cudaMemcpyHostToDevice, ctx->stream);
  int rc = (e == cudaSuccess) ? GAPIT_OK : fail(GAPIT_ERR_CUDA, std::string("gapit_eigh: ") + cudaGetErrorString(e));
  if (rc == GAPIT_OK) rc = syevd_dev(ctx, A_dev, n, w_dev);
  if (rc == GAPIT_OK) rc = emit_desc(ctx, A_dev, w_dev, n, n, 0.0, values_out, vectors_out);
  cudaFree(A_dev);
  cudaFree(w_dev);
  return rc;
}

extern "C" int gapit_eigh_dev(gapit_ctx* ctx, double* A_dev, int64_t n, double* values_out) {
  if (!ctx || !A_dev || !values_out || n <= 0) return fail(GAPIT_ERR_ARG, "gapit_eigh_dev: bad argument");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  double* w_dev = nullptr;
  GAPIT_CUDA(cudaMalloc(&w_dev, size_t(n) * 8));
  int rc = syevd_dev(ctx, A_dev, n, w_dev);
  std::vector<double> w(n);
  cudaError_t e = cudaSuccess;
  if (rc == GAPIT_OK) {
    if (n >= 2) {
      dim3 grid(static_cast<unsigned>((n + 255) / 256), static_cast<unsigned>(std::min<int64_t>(n / 2, 65535)));
      swap_cols_kernel<<<grid, 256, 0, ctx->stream>>>(A_dev, n);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(w.data(), w_dev, size_t(n) * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(w_dev);
  if (rc != GAPIT_OK) return rc;
  if (e != cudaSuccess) return fail(GAPIT_ERR_CUDA, std::string("gapit_eigh_dev: ") + cudaGetErrorString(e));
  for (int64_t c = 0; c < n; ++c) values_out[c] = w[n - 1 - c];
  return GAPIT_OK;
}

// Eigenpairs [first, first + count) of R's descending order, by index range (cusolverDnXsyevdx).  profiles/
// r2_syevd_breakdown.log: at n = 20,000 the tridiagonalisation is 3.9 s of syevd's 4.5 s and a range call costs 4.1 s
// whatever the range, so several GPUs each taking a slice of the spectrum (gapit_multi_eigh, gapit_b200/fullsize.py)
// share the divide-and-conquer back-transformation and the output traffic, not the tridiagonalisation.
extern "C" int gapit_eigh_dev_range(gapit_ctx* ctx, double* A_dev, int64_t n, int64_t first, int64_t count,
                                    double* values_out, double* vectors_dev_out) {
  if (!ctx || !A_dev || !values_out || !vectors_dev_out || n <= 0 || first < 0 || count < 1 || first + count > n)
    return fail(GAPIT_ERR_ARG, "gapit_eigh_dev_range: bad argument");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  GAPIT_CHECK(ensure_solver(ctx));
  const int64_t il = n - (first + count) + 1, iu = n - first;      // 1-based, ascending
  double* w_dev = nullptr;
  int* info = nullptr;
  void* work = nullptr;
  std::vector<char> hwork;
  cusolverDnParams_t params = nullptr;
  GAPIT_CUDA(cudaMalloc(&w_dev, size_t(n) * 8));
  cudaError_t e = cudaMalloc(&info, 4);
  size_t dbytes = 0, hbytes = 0;
  int64_t meig = 0;
  double vl = 0.0, vu = 0.0;
  cusolverStatus_t st = e == cudaSuccess ? cusolverDnCreateParams(&params) : CUSOLVER_STATUS_ALLOC_FAILED;
  if (st == CUSOLVER_STATUS_SUCCESS)
    st = cusolverDnXsyevdx_bufferSize(ctx->solver, params, CUSOLVER_EIG_MODE_VECTOR, CUSOLVER_EIG_RANGE_I, CUBLAS_FILL_MODE_LOWER, n,
                                      CUDA_R_64F, A_dev, n, &vl, &vu, il, iu, &meig, CUDA_R_64F, w_dev, CUDA_R_64F, &dbytes, &hbytes);
  if (st == CUSOLVER_STATUS_SUCCESS && cudaMalloc(&work, std::max<size_t>(dbytes, 16)) != cudaSuccess)
    st = CUSOLVER_STATUS_ALLOC_FAILED;
  if (st == CUSOLVER_STATUS_SUCCESS) {
    hwork.resize(std::max<size_t>(hbytes, 16));
    st = cusolverDnXsyevdx(ctx->solver, params, CUSOLVER_EIG_MODE_VECTOR, CUSOLVER_EIG_RANGE_I, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F,
                           A_dev, n, &vl, &vu, il, iu, &meig, CUDA_R_64F, w_dev, CUDA_R_64F, work, dbytes, hwork.data(), hbytes, info);
  }
  int hinfo = -1;
  std::vector<double> w(count);
  if (st == CUSOLVER_STATUS_SUCCESS && meig == count) {
    // the slice comes back ascending in the first `count` columns of A_dev and entries of w_dev
    dim3 grid(static_cast<unsigned>((n + 255) / 256), static_cast<unsigned>(std::min<int64_t>(count, 65535)));
    reverse_first_cols_kernel<<<grid, 256, 0, ctx->stream>>>(A_dev, n, count, vectors_dev_out);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(w.data(), w_dev, size_t(count) * 8, cudaMemcpyDeviceToHost, ctx->stream);
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(&hinfo, info, 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(work);
  cudaFree(info);
  cudaFree(w_dev);
  if (params) cusolverDnDestroyParams(params);
  if (st != CUSOLVER_STATUS_SUCCESS) return fail(GAPIT_ERR_CUDA, "cuSOLVER syevdx failed, status " + std::to_string(int(st)));
  if (e != cudaSuccess) return fail(GAPIT_ERR_CUDA, std::string("gapit_eigh_dev_range: ") + cudaGetErrorString(e));
  if (hinfo != 0) return fail(GAPIT_ERR_NUMERIC, "syevdx did not converge, info " + std::to_string(hinfo));
  if (meig != count) return fail(GAPIT_ERR_NUMERIC, "syevdx returned " + std::to_string(meig) + " eigenpairs, expected " + std::to_string(count));
  for (int64_t c = 0; c < count; ++c) values_out[c] = w[count - 1 - c];
  return GAPIT_OK;
}

// eigen(K) straight into a resident, sliced eigen-system: the eigenvectors never visit the host (at n = 20,000 they
// are 3.2 GB each way).  values_out: host, n, descending.
extern "C" int gapit_eigsys_from_kinship(gapit_ctx* ctx, const double* K, int64_t n, double* values_out, gapit_eigsys** out) {
  if (!ctx || !K || !values_out || !out || n <= 0) return fail(GAPIT_ERR_ARG, "gapit_eigsys_from_kinship: bad argument");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  double* A_dev = nullptr;
  GAPIT_CUDA(cudaMalloc(&A_dev, size_t(n) * n * 8));
  cudaError_t e = cudaMemcpyAsync(A_dev, K, size_t(n) * n * 8, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  int rc = e == cudaSuccess ? GAPIT_OK : fail(GAPIT_ERR_CUDA, std::string("gapit_eigsys_from_kinship: ") + cudaGetErrorString(e));
  if (rc == GAPIT_OK) rc = gapit_eigh_dev(ctx, A_dev, n, values_out);
  if (rc == GAPIT_OK) rc = gapit_eigsys_from_device(ctx, A_dev, values_out, n, out);
  cudaFree(A_dev);
  return rc;
}

// Multi-GPU eigensolver: cusolverMgSyevd lives in libgapitcuda_mg.so next to this library and is loaded on first use.
// (libcusolverMg binds the CUDA toolkit's own libcublas; a host process that already carries another libcublas -- a
// PyTorch wheel's -- cannot load it, and then this call reports that and the caller keeps the one-GPU solver.)
typedef int (*MgSyevdFn)(int, void*, double*, int64_t, double*, int, char*, int);
static MgSyevdFn mg_syevd_fn(std::string* why) {
  static MgSyevdFn fn = nullptr;
  static std::string err;
  static bool tried = false;
  if (!tried) {
    tried = true;
    Dl_info info;
    std::string dir = ".";
    if (dladdr(reinterpret_cast<void*>(&gapit_eigh_dev), &info) && info.dli_fname) {
      dir = info.dli_fname;
      const size_t slash = dir.find_last_of('/');
      dir = slash == std::string::npos ? "." : dir.substr(0, slash);
    }
    void* h = dlopen((dir + "/libgapitcuda_mg.so").c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!h) err = dlerror();
    else if (!(fn = reinterpret_cast<MgSyevdFn>(dlsym(h, "gapit_mg_syevd")))) err = "gapit_mg_syevd missing in libgapitcuda_mg.so";
  }
  if (!fn && why) *why = err;
  return fn;
}

extern "C" int gapit_eigh_dev_mg(gapit_ctx* ctx, double* A_dev, int64_t n, double* values_out, int ngpus) {
  if (!ctx || !A_dev || !values_out || n <= 0) return fail(GAPIT_ERR_ARG, "gapit_eigh_dev_mg: bad argument");
  int ndev = gapit_device_count();
  if (ngpus <= 0) ngpus = ndev;
  ngpus = std::min(ngpus, ndev);
  if (ngpus < 2 || n < 2048) return gapit_eigh_dev(ctx, A_dev, n, values_out);
  std::string why;
  MgSyevdFn fn = mg_syevd_fn(&why);
  if (!fn) return fail(GAPIT_ERR_NO_DEVICE, "multi-GPU eigensolver unavailable in this process: " + why);
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  std::vector<double> w(n);
  char msg[512] = {0};
  const int rc = fn(ctx->device, ctx->stream, A_dev, n, w.data(), ngpus, msg, sizeof(msg));
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  if (rc != 0) return fail(rc, msg);
  GAPIT_CHECK(launch_swap_cols(ctx->stream, A_dev, n));   // ascending -> R's descending order
  GAPIT_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int64_t c = 0; c < n; ++c) values_out[c] = w[n - 1 - c];
  return GAPIT_OK;
}

// host-buffer form of the above (what the R shim calls for eigen(K)): falls back to the one-GPU solver when the
// multi-GPU one cannot be used in this process
extern "C" int gapit_eigh_mg(gapit_ctx* ctx, const double* A, int64_t n, double* values_out, double* vectors_out, int ngpus) {
  if (!ctx || !A || !values_out || !vectors_out || n <= 0) return fail(GAPIT_ERR_ARG, "gapit_eigh_mg: bad argument");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  double* A_dev = nullptr;
  GAPIT_CUDA(cudaMalloc(&A_dev, size_t(n) * n * 8));
  cudaError_t e = cudaMemcpyAsync(A_dev, A, size_t(n) * n * 8, cudaMemcpyHostToDevice, ctx->stream);
  int rc = (e == cudaSuccess) ? GAPIT_OK : fail(GAPIT_ERR_CUDA, std::string("gapit_eigh_mg: ") + cudaGetErrorString(e));
  if (rc == GAPIT_OK) {
    rc = gapit_eigh_dev_mg(ctx, A_dev, n, values_out, ngpus);
    if (rc == GAPIT_ERR_NO_DEVICE) rc = gapit_eigh_dev(ctx, A_dev, n, values_out);
  }
  if (rc == GAPIT_OK) {
    e = cudaMemcpyAsync(vectors_out, A_dev, size_t(n) * n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = fail(GAPIT_ERR_CUDA, std::string("gapit_eigh_mg: ") + cudaGetErrorString(e));
  }
  cudaFree(A_dev);
  return rc;
}

extern "C" int gapit_eigen_R(gapit_ctx* ctx, const double* K, const double* X, int64_t n, int64_t q, double* values_out,
                             double* vectors_out) {
  if (!ctx || !K || !X || !values_out || !vectors_out || n <= 0 || q < 1 || q >= n || q > GAPIT_MAX_Q0)
    return fail(GAPIT_ERR_ARG, "gapit_eigen_R: bad argument");
  GAPIT_CUDA(cudaSetDevice(ctx->device));
  double *K_dev = nullptr, *A_dev = nullptr, *w_dev = nullptr, *small = nullptr;
  auto freeall = [&] {
    cudaFree(K_dev);
    cudaFree(A_dev);
    cudaFree(w_dev);
    cudaFree(small);
  };
  cudaError_t e = cudaMalloc(&K_dev, size_t(n) * n * 8);
  if (e == cudaSuccess) e = cudaMalloc(&A_dev, size_t(n) * n * 8);
  if (e == cudaSuccess) e = cudaMalloc(&w_dev, size_t(n) * 8);
  if (e == cudaSuccess) e = cudaMalloc(&small, size_t(n) * q * 8 * 4);  // X | B | Y1 | Y2
  if (e == cudaSuccess) e = cudaMemcpyAsync(K_dev, K, size_t(n) * n * 8, cudaMemcpyHostToDevice, ctx->stream);
  double* X_dev = small;
  double* B_dev = small + size_t(n) * q;
  double* Y1_dev = small + size_t(n) * q * 2;
  double* Y2_dev = small + size_t(n) * q * 3;
  if (e == cudaSuccess) e = cudaMemcpyAsync(X_dev, X, size_t(n) * q * 8, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) {
    freeall();
    return fail(GAPIT_ERR_CUDA, std::string("gapit_eigen_R: ") + cudaGetErrorString(e));
  }
  // B = (K + I) X
  sym_times_cols_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0, ctx->stream>>>(K_dev, n, X_dev, int(q), B_dev);
  std::vector<double> B(size_t(n) * q);
  e = cudaMemcpyAsync(B.data(), B_dev, size_t(n) * q * 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    freeall();
    return fail(GAPIT_ERR_CUDA, std::string("gapit_eigen_R: ") + cudaGetErrorString(e));
  }
  for (size_t i = 0; i < B.size(); ++i) B[i] += X[i];
  // small host algebra: iXX = (X'X)^-1 (Gauss-Jordan with partial pivoting, as solve() would), C = X'B
  std::vector<double> XX(q * q), iXX(q * q, 0.0), C(q * q), Y1(size_t(n) * q), Y2(size_t(n) * q);
  for (int64_t a = 0; a < q; ++a)
    for (int64_t b = 0; b < q; ++b) {
      double s = 0.0, c = 0.0;
      for (int64_t i = 0; i < n; ++i) {
        s += X[a * n + i] * X[b * n + i];
        c += X[a * n + i] * B[b * n + i];
      }
      XX[a * q + b] = s;
      C[a * q + b] = c;
    }
  {
    std::vector<double> aug(q * 2 * q, 0.0);
    for (int64_t a = 0; a < q; ++a) {
      for (int64_t b = 0; b < q; ++b) aug[a * 2 * q + b] = XX[a * q + b];
      aug[a * 2 * q + q + a] = 1.0;
    }
    for (int64_t c = 0; c < q; ++c) {
      int64_t piv = c;
      for (int64_t r = c + 1; r < q; ++r)
        if (std::fabs(aug[r * 2 * q + c]) > std::fabs(aug[piv * 2 * q + c])) piv = r;
      if (std::fabs(aug[piv * 2 * q + c]) < 1e-300) {
        freeall();
        return fail(GAPIT_ERR_NUMERIC, "gapit_eigen_R: X'X is singular");
      }
      if (piv != c)
        for (int64_t b = 0; b < 2 * q; ++b) std::swap(aug[c * 2 * q + b], aug[piv * 2 * q + b]);
      const double d = aug[c * 2 * q + c];
      for (int64_t b = 0; b < 2 * q; ++b) aug[c * 2 * q + b] /= d;
      for (int64_t r = 0; r < q; ++r) {
        if (r == c) continue;
        const double f = aug[r * 2 * q + c];
        if (f != 0.0)
          for (int64_t b = 0; b < 2 * q; ++b) aug[r * 2 * q + b] -= f * aug[c * 2 * q + b];
      }
    }
    for (int64_t a = 0; a < q; ++a)
      for (int64_t b = 0; b < q; ++b) iXX[a * q + b] = aug[a * 2 * q + q + b];
  }
  for (int64_t l = 0; l < q; ++l)
    for (int64_t i = 0; i < n; ++i) {
      double s = 0.0;
      for (int64_t a = 0; a < q; ++a) s += X[a * n + i] * iXX[a * q + l];
      Y1[l * n + i] = s;
    }
  for (int64_t l = 0; l < q; ++l)
    for (int64_t i = 0; i < n; ++i) {
      double s = 0.0;
      for (int64_t a = 0; a < q; ++a) s += Y1[a * n + i] * C[a * q + l];
      Y2[l * n + i] = s;
    }
  e = cudaMemcpyAsync(B_dev, B.data(), size_t(n) * q * 8, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(Y1_dev, Y1.data(), size_t(n) * q * 8, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(Y2_dev, Y2.data(), size_t(n) * q * 8, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    dim3 grid(static_cast<unsigned>((n + 255) / 256), static_cast<unsigned>(std::min<int64_t>(n, 65535)));
    form_sks_kernel<<<grid, 256, 0, ctx->stream>>>(K_dev, n, Y1_dev, B_dev, Y2_dev, int(q), A_dev);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // host vectors B/Y1/Y2 must outlive the copies
  int rc = (e == cudaSuccess) ? GAPIT_OK : fail(GAPIT_ERR_CUDA, std::string("gapit_eigen_R: ") + cudaGetErrorString(e));
  if (rc == GAPIT_OK) rc = syevd_dev(ctx, A_dev, n, w_dev);
  if (rc == GAPIT_OK) rc = emit_desc(ctx, A_dev, w_dev, n, n - q, -1.0, values_out, vectors_out);
  freeall();
  return rc;
}
