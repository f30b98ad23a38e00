// Experiment: where does cusolverDnXsyevd spend its time at kinship sizes, and can index-range calls (Xsyevdx) split the
// eigenvector work over several GPUs?  Standalone (no libgapitcuda): builds K = Z Z'/m from random centred columns.
//   nvcc -O2 -arch=sm_100a tools/exp_syevdx.cu -lcusolver -lcublas -lcurand -o tools/_bin/exp_syevdx
//   tools/_bin/exp_syevdx [n] [m]
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <curand.h>
#include <cusolverDn.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                  \
  do {                                                                         \
    auto rc_ = (x);                                                            \
    if (rc_ != 0) {                                                            \
      std::printf("FAILED %s -> %d (line %d)\n", #x, int(rc_), __LINE__);      \
      std::exit(1);                                                            \
    }                                                                          \
  } while (0)

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

__global__ void centre_kernel(double* z, size_t count) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i < count) {   // uniform (0,1] -> dosage 0/1/2 with p = 0.3, centred by 2p
    const double u = z[i];
    const double d = u < 0.49 ? 0.0 : (u < 0.91 ? 1.0 : 2.0);
    z[i] = d - 0.6;
  }
}
__global__ void sub_eye_absmax(const double* c, int64_t k, double* out) {
  __shared__ double red[256];
  double mx = 0.0;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < k * k; i += int64_t(gridDim.x) * blockDim.x) {
    const double v = c[i] - ((i / k) == (i % k) ? 1.0 : 0.0);
    mx = fmax(mx, fabs(v));
  }
  red[threadIdx.x] = mx;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long*>(out), __double_as_longlong(red[0]));
}
__global__ void absmax(const double* c, int64_t count, double* out) {
  __shared__ double red[256];
  double mx = 0.0;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x)
    mx = fmax(mx, fabs(c[i]));
  red[threadIdx.x] = mx;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long*>(out), __double_as_longlong(red[0]));
}
__global__ void scale_cols(double* v, int64_t n, int64_t k, const double* w) {   // v[:, j] *= w[j]
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n * k; i += int64_t(gridDim.x) * blockDim.x) v[i] *= w[i / n];
}

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 8000;
  const int64_t m = argc > 2 ? atoll(argv[2]) : 2 * n;
  cusolverDnHandle_t sh;
  cublasHandle_t bh;
  cusolverDnParams_t params;
  CK(cusolverDnCreate(&sh));
  CK(cublasCreate(&bh));
  CK(cusolverDnCreateParams(&params));
  double *K, *A, *W, *red;
  CK(cudaMalloc(&K, n * n * 8));
  CK(cudaMalloc(&A, n * n * 8));
  CK(cudaMalloc(&W, n * 8));
  CK(cudaMalloc(&red, 8));
  {
    const int64_t mb = 8192;   // K accumulates block by block so Z never needs n*m doubles
    double* Z;
    CK(cudaMalloc(&Z, n * mb * 8));
    curandGenerator_t gen;
    CK(curandCreateGenerator(&gen, CURAND_RNG_PSEUDO_PHILOX4_32_10));
    CK(curandSetPseudoRandomGeneratorSeed(gen, 7));
    const double one = 1.0 / double(m);
    for (int64_t b = 0; b * mb < m; ++b) {
      CK(curandGenerateUniformDouble(gen, Z, n * mb));
      centre_kernel<<<unsigned((n * mb + 255) / 256), 256>>>(Z, size_t(n * mb));
      const double beta = b ? 1.0 : 0.0;
      CK(cublasDsyrk(bh, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, int(n), int(mb), &one, Z, int(n), &beta, K, int(n)));
    }
    CK(cudaDeviceSynchronize());
    cudaFree(Z);
    curandDestroyGenerator(gen);
  }
  auto run = [&](const char* what, cusolverEigMode_t jobz, bool ranged, int64_t il, int64_t iu) -> double {
    CK(cudaMemcpy(A, K, n * n * 8, cudaMemcpyDeviceToDevice));
    size_t wd = 0, wh = 0;
    int64_t meig = 0;
    double vl = 0, vu = 0;
    if (ranged)
      CK(cusolverDnXsyevdx_bufferSize(sh, params, jobz, CUSOLVER_EIG_RANGE_I, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, n, &vl, &vu,
                                      il, iu, &meig, CUDA_R_64F, W, CUDA_R_64F, &wd, &wh));
    else
      CK(cusolverDnXsyevd_bufferSize(sh, params, jobz, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, n, CUDA_R_64F, W, CUDA_R_64F, &wd, &wh));
    void* dbuf = nullptr;
    CK(cudaMalloc(&dbuf, wd ? wd : 8));
    std::vector<char> hbuf(wh ? wh : 8);
    int* info;
    CK(cudaMalloc(&info, 4));
    CK(cudaDeviceSynchronize());
    const double t0 = now();
    if (ranged)
      CK(cusolverDnXsyevdx(sh, params, jobz, CUSOLVER_EIG_RANGE_I, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, n, &vl, &vu, il, iu, &meig,
                           CUDA_R_64F, W, CUDA_R_64F, dbuf, wd, hbuf.data(), wh, info));
    else
      CK(cusolverDnXsyevd(sh, params, jobz, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, n, CUDA_R_64F, W, CUDA_R_64F, dbuf, wd, hbuf.data(), wh, info));
    CK(cudaDeviceSynchronize());
    const double dt = now() - t0;
    int hinfo = 0;
    CK(cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost));
    std::printf("n=%lld %-44s %.3f s  (workspace %.2f GB device, info %d, meig %lld)\n", (long long)n, what, dt, wd / 1e9, hinfo,
                (long long)meig);
    std::fflush(stdout);
    cudaFree(dbuf);
    cudaFree(info);
    return dt;
  };
  // quality of k eigenpairs left in A[:, 0:k] / W[0:k]:  |V'V - I|max and |K V - V diag(w)|max
  auto quality = [&](int64_t k) {
    double *C, *R;
    CK(cudaMalloc(&C, k * k * 8));
    CK(cudaMalloc(&R, n * k * 8));
    const double one = 1.0, zero = 0.0, mone = -1.0;
    CK(cublasDgemm(bh, CUBLAS_OP_T, CUBLAS_OP_N, int(k), int(k), int(n), &one, A, int(n), A, int(n), &zero, C, int(k)));
    double h[2] = {0, 0};
    CK(cudaMemset(red, 0, 8));
    sub_eye_absmax<<<1024, 256>>>(C, k, red);
    CK(cudaMemcpy(&h[0], red, 8, cudaMemcpyDeviceToHost));
    CK(cublasDsymm(bh, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, int(n), int(k), &one, K, int(n), A, int(n), &zero, R, int(n)));
    // R -= V diag(w): scale a copy of V
    double* V2;
    CK(cudaMalloc(&V2, n * k * 8));
    CK(cudaMemcpy(V2, A, n * k * 8, cudaMemcpyDeviceToDevice));
    scale_cols<<<1024, 256>>>(V2, n, k, W);
    CK(cublasDaxpy(bh, int(n * k), &mone, V2, 1, R, 1));
    CK(cudaMemset(red, 0, 8));
    absmax<<<1024, 256>>>(R, n * k, red);
    CK(cudaMemcpy(&h[1], red, 8, cudaMemcpyDeviceToHost));
    std::printf("    %lld eigenpairs: |V'V - I|max %.2e   |K V - V L|max %.2e\n", (long long)k, h[0], h[1]);
    std::fflush(stdout);
    cudaFree(C);
    cudaFree(R);
    cudaFree(V2);
  };
  run("Xsyevd  vectors (warm-up)", CUSOLVER_EIG_MODE_VECTOR, false, 0, 0);
  run("Xsyevd  vectors", CUSOLVER_EIG_MODE_VECTOR, false, 0, 0);
  quality(n);
  run("Xsyevd  values only (sytrd + sterf)", CUSOLVER_EIG_MODE_NOVECTOR, false, 0, 0);
  const int64_t k8 = n / 8;
  run("Xsyevdx vectors, indices 1..n/8", CUSOLVER_EIG_MODE_VECTOR, true, 1, k8);
  quality(k8);
  // keep this slice, compute the neighbouring one, check the two are orthogonal to each other
  double* Vkeep;
  CK(cudaMalloc(&Vkeep, n * k8 * 8));
  CK(cudaMemcpy(Vkeep, A, n * k8 * 8, cudaMemcpyDeviceToDevice));
  run("Xsyevdx vectors, indices n/8+1..n/4", CUSOLVER_EIG_MODE_VECTOR, true, k8 + 1, 2 * k8);
  quality(k8);
  {
    double* C;
    CK(cudaMalloc(&C, k8 * k8 * 8));
    const double one = 1.0, zero = 0.0;
    CK(cublasDgemm(bh, CUBLAS_OP_T, CUBLAS_OP_N, int(k8), int(k8), int(n), &one, Vkeep, int(n), A, int(n), &zero, C, int(k8)));
    double h = 0;
    CK(cudaMemset(red, 0, 8));
    absmax<<<1024, 256>>>(C, k8 * k8, red);
    CK(cudaMemcpy(&h, red, 8, cudaMemcpyDeviceToHost));
    std::printf("    across the two slices: |V1'V2|max %.2e\n", h);
    cudaFree(C);
  }
  run("Xsyevdx vectors, indices 7n/8+1..n", CUSOLVER_EIG_MODE_VECTOR, true, n - k8 + 1, n);
  quality(k8);
  run("Xsyevdx vectors, indices 1..n/2", CUSOLVER_EIG_MODE_VECTOR, true, 1, n / 2);
  run("Xsyevdx values only, indices 1..n/8", CUSOLVER_EIG_MODE_NOVECTOR, true, 1, k8);
  return 0;
}
