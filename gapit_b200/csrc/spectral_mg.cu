// eigen(K) on SEVERAL GPUs of the box: cuSOLVER's multi-GPU dense symmetric solver (cusolverMgSyevd, a library call like
// the one-GPU syevd; north_star keeps the eigendecomposition outside the hot-path roofline but it is 93 % of the
// end-to-end time of BASELINE configs[2] when seven of eight GPUs idle through it).
// K (n x n column-major on the context's device) is dealt out in column blocks, 1-D block-cyclic, over `ngpus` devices
// of this process with peer copies over NVLink, solved in place, and gathered back -- eigenvectors in R's descending
// order in A_dev, eigenvalues on the host -- so the callers of gapit_eigh_dev can switch by passing a GPU count.
#include <cusolverMg.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

// Self-contained (no symbol of libgapitcuda is used): loaded with dlopen by gapit_eigh_dev_mg (spectral.cu).
// On return A_dev (on `device`) holds the eigenvectors in ASCENDING order of the eigenvalues w_out (host, n);
// 0 = ok, otherwise a negative status with the message in err (capacity errcap).
static int failmsg(char* err, int errcap, int code, const std::string& msg) {
  if (err && errcap > 0) snprintf(err, size_t(errcap), "%s", msg.c_str());
  return code;
}

extern "C" int gapit_mg_syevd(int device, void* stream_v, double* A_dev, int64_t n, double* w_out, int ngpus, char* err,
                              int errcap) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  ngpus = std::min(ngpus <= 0 ? ndev : ngpus, ndev);
  if (ngpus < 2) return failmsg(err, errcap, -1, "gapit_mg_syevd: needs at least two GPUs");
  if (n > 2147483647ll) return failmsg(err, errcap, -1, "gapit_mg_syevd: n exceeds the solver's 32-bit index");
  // devices: the caller's own first, then the others in index order; peer access between every pair
  std::vector<int> ids{device};
  for (int d = 0; d < ndev && static_cast<int>(ids.size()) < ngpus; ++d)
    if (d != device) ids.push_back(d);
  for (int a : ids)
    for (int b : ids) {
      if (a == b) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, a, b);
      if (!can) return failmsg(err, errcap, -3, "gapit_mg_syevd: no peer access between the GPUs");
      cudaSetDevice(a);
      const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return failmsg(err, errcap, -2, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
      cudaGetLastError();
    }
  cudaSetDevice(device);
  cudaStreamSynchronize(stream);                     // A_dev is final before other devices read it
  const int64_t TA = 256;                            // column block of the 1-D block-cyclic layout
  const int64_t nblk = (n + TA - 1) / TA;
  cusolverMgHandle_t h = nullptr;
  cudaLibMgGrid_t grid = nullptr;
  cudaLibMgMatrixDesc_t desc = nullptr;
  std::vector<void*> dA(ngpus, nullptr), dW(ngpus, nullptr);
  auto release = [&](int rc) {
    for (int g = 0; g < ngpus; ++g) {
      cudaSetDevice(ids[g]);
      cudaFree(dA[g]);
      cudaFree(dW[g]);
    }
    if (desc) cusolverMgDestroyMatrixDesc(desc);
    if (grid) cusolverMgDestroyGrid(grid);
    if (h) cusolverMgDestroy(h);
    cudaSetDevice(device);
    return rc;
  };
  cusolverStatus_t st = cusolverMgCreate(&h);
  if (st == CUSOLVER_STATUS_SUCCESS) st = cusolverMgDeviceSelect(h, ngpus, ids.data());
  if (st == CUSOLVER_STATUS_SUCCESS) st = cusolverMgCreateDeviceGrid(&grid, 1, ngpus, ids.data(), CUDALIBMG_GRID_MAPPING_COL_MAJOR);
  if (st == CUSOLVER_STATUS_SUCCESS) st = cusolverMgCreateMatrixDesc(&desc, n, n, n, TA, CUDA_R_64F, grid);
  if (st != CUSOLVER_STATUS_SUCCESS) return release(failmsg(err, errcap, -2, "cusolverMg setup failed, status " + std::to_string(int(st))));
  // deal the column blocks out: block j -> device j % ngpus, local block j / ngpus
  for (int g = 0; g < ngpus; ++g) {
    const int64_t myblk = (nblk - g + ngpus - 1) / ngpus;
    cudaSetDevice(ids[g]);
    if (cudaMalloc(&dA[g], size_t(std::max<int64_t>(myblk, 1)) * TA * n * 8) != cudaSuccess)
      return release(failmsg(err, errcap, -2, "gapit_mg_syevd: cudaMalloc of a column-block panel failed"));
  }
  cudaSetDevice(device);
  for (int64_t j = 0; j < nblk; ++j) {
    const int g = static_cast<int>(j % ngpus);
    const int64_t cols = std::min(TA, n - j * TA);
    const cudaError_t e = cudaMemcpyPeerAsync(static_cast<double*>(dA[g]) + (j / ngpus) * TA * n, ids[g], A_dev + j * TA * n,
                                              device, size_t(cols) * n * 8, stream);
    if (e != cudaSuccess) return release(failmsg(err, errcap, -2, std::string("gapit_mg_syevd: scatter: ") + cudaGetErrorString(e)));
  }
  cudaStreamSynchronize(stream);
  int64_t lwork = 0;
  st = cusolverMgSyevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, static_cast<int>(n), dA.data(), 1, 1, desc,
                                  w_out, CUDA_R_64F, CUDA_R_64F, &lwork);
  if (st != CUSOLVER_STATUS_SUCCESS)
    return release(failmsg(err, errcap, -2, "cusolverMgSyevd_bufferSize failed, status " + std::to_string(int(st))));
  for (int g = 0; g < ngpus; ++g) {
    cudaSetDevice(ids[g]);
    if (cudaMalloc(&dW[g], size_t(std::max<int64_t>(lwork, 2)) * 8) != cudaSuccess)
      return release(failmsg(err, errcap, -2, "gapit_mg_syevd: cudaMalloc of the solver workspace failed"));
  }
  int info = -1;
  st = cusolverMgSyevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, static_cast<int>(n), dA.data(), 1, 1, desc, w_out,
                       CUDA_R_64F, CUDA_R_64F, dW.data(), lwork, &info);
  for (int g = 0; g < ngpus; ++g) {
    cudaSetDevice(ids[g]);
    cudaDeviceSynchronize();
  }
  if (st != CUSOLVER_STATUS_SUCCESS) return release(failmsg(err, errcap, -2, "cusolverMgSyevd failed, status " + std::to_string(int(st))));
  if (info != 0) return release(failmsg(err, errcap, -5, "cusolverMgSyevd did not converge, info " + std::to_string(info)));
  cudaSetDevice(device);
  for (int64_t j = 0; j < nblk; ++j) {
    const int g = static_cast<int>(j % ngpus);
    const int64_t cols = std::min(TA, n - j * TA);
    const cudaError_t e = cudaMemcpyPeerAsync(A_dev + j * TA * n, device, static_cast<double*>(dA[g]) + (j / ngpus) * TA * n, ids[g],
                                              size_t(cols) * n * 8, stream);
    if (e != cudaSuccess) return release(failmsg(err, errcap, -2, std::string("gapit_mg_syevd: gather: ") + cudaGetErrorString(e)));
  }
  const cudaError_t e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return release(failmsg(err, errcap, -2, std::string("gapit_mg_syevd: ") + cudaGetErrorString(e)));
  return release(0);
}
