// All GPUs of one box behind the C ABI, in ONE process (`gapit_ctx_create(int ngpus, ...)`,
// genotypes sharded by marker block, section 8e).  An R session -- or any single-threaded host -- then drives 8 B200s
// through the same three calls, with no launcher and no torch.distributed:
//
//   gapit_mctx_create      one gapit_ctx per GPU, peer access enabled between every pair (NVLink / NVSwitch)
//   gapit_mgeno_pack       contiguous marker blocks, one per GPU, uploaded and packed concurrently (one host thread per GPU)
//   gapit_multi_vanraden   every GPU's Gram kernel reduce-scatters its tiles into the owners' buffers over NVLink
//                          (bulk tensor reduce-adds, kinship.cu), GPU 0 gathers the block-rows and runs the fp64 epilogue
//   gapit_multi_p3d_scan   eigenvectors uploaded once, copied GPU to GPU, sliced per GPU; every GPU scans its block;
//                          results land in the caller's arrays in marker order (the scan needs no collective)
//
// Worker threads never touch the caller's runtime (R API, Python GIL); errors come back as status + message.
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace gapit {
int scan_impl(gapit_ctx* ctx, gapit_geno* g, gapit_eigsys* e, const double* X0, int q0, const double* Y, int T,
              const double* delta, const double* vg, int glm, gapit_scan_out* out, gapit_scan_base* base_out,
              int64_t m_total, int64_t m_off);
}

struct gapit_mctx {
  std::vector<gapit_ctx*> ctx;
};

struct gapit_mgeno {
  gapit_mctx* mc = nullptr;
  int64_t n = 0, m = 0;
  std::vector<gapit_geno*> part;   // one marker block per GPU (nullptr when the block is empty)
  std::vector<int64_t> lo;         // first marker of each block; lo[world] = m
};

// eigen-system of K resident on every GPU of the context (one sliced copy per GPU)
struct gapit_meigsys {
  gapit_mctx* mc = nullptr;
  int64_t n = 0;
  std::vector<gapit_eigsys*> es;
};

namespace gapit {

// run fn(rank) on one host thread per GPU; the first failure (status, message) is returned on the calling thread
template <class F>
static int for_each_gpu(int world, F fn) {
  std::vector<int> rc(world, GAPIT_OK);
  std::vector<std::string> msg(world);
  std::vector<std::thread> th;
  th.reserve(world);
  for (int r = 0; r < world; ++r)
    th.emplace_back([&, r] {
      rc[r] = fn(r);
      if (rc[r] != GAPIT_OK) msg[r] = gapit_last_error();
    });
  for (auto& t : th) t.join();
  for (int r = 0; r < world; ++r)
    if (rc[r] != GAPIT_OK) return fail(rc[r], "GPU " + std::to_string(r) + ": " + msg[r]);
  return GAPIT_OK;
}

}  // namespace gapit

using namespace gapit;

extern "C" int gapit_mctx_create(int ngpus, gapit_mctx** out) {
  if (!out) return fail(GAPIT_ERR_ARG, "gapit_mctx_create: NULL out");
  const int ndev = gapit_device_count();
  if (ndev <= 0) return fail(GAPIT_ERR_NO_DEVICE, "no CUDA device visible: libgapitcuda has no CPU fallback");
  if (ngpus <= 0) ngpus = std::min(ndev, 8);
  if (ngpus > ndev || ngpus > 8) return fail(GAPIT_ERR_ARG, "gapit_mctx_create: more GPUs asked for than visible (or > 8)");
  gapit_mctx* mc = new gapit_mctx();
  for (int d = 0; d < ngpus; ++d) {
    gapit_ctx* c = nullptr;
    const int rc = gapit_ctx_create(d, &c);
    if (rc != GAPIT_OK) {
      gapit_mctx_destroy(mc);
      return rc;
    }
    mc->ctx.push_back(c);
  }
  for (int a = 0; a < ngpus; ++a)
    for (int b = 0; b < ngpus; ++b) {
      if (a == b) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, a, b);
      if (!can) {
        gapit_mctx_destroy(mc);
        return fail(GAPIT_ERR_NO_DEVICE, "gapit_mctx_create: GPUs " + std::to_string(a) + " and " + std::to_string(b) +
                                             " have no peer access (NVLink / NVSwitch needed for the fused exchange)");
      }
      cudaSetDevice(a);
      const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        gapit_mctx_destroy(mc);
        return fail(GAPIT_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
      }
      cudaGetLastError();
    }
  *out = mc;
  return GAPIT_OK;
}

extern "C" void gapit_mctx_destroy(gapit_mctx* mc) {
  if (!mc) return;
  for (gapit_ctx* c : mc->ctx) gapit_ctx_destroy(c);
  delete mc;
}

extern "C" int gapit_mctx_ngpus(const gapit_mctx* mc) { return mc ? static_cast<int>(mc->ctx.size()) : 0; }

extern "C" gapit_ctx* gapit_mctx_ctx(gapit_mctx* mc, int gpu) {
  return (mc && gpu >= 0 && gpu < static_cast<int>(mc->ctx.size())) ? mc->ctx[gpu] : nullptr;
}

extern "C" void gapit_mgeno_free(gapit_mgeno* g) {
  if (!g) return;
  for (gapit_geno* p : g->part) gapit_geno_free(p);
  delete g;
}

extern "C" int gapit_mgeno_pack(gapit_mctx* mc, const void* snps, gapit_dtype dtype, int64_t n, int64_t m, int64_t ld,
                                gapit_impute impute, gapit_mgeno** out) {
  if (!mc || !snps || !out || n <= 0 || m <= 0) return fail(GAPIT_ERR_ARG, "gapit_mgeno_pack: bad argument");
  const int world = static_cast<int>(mc->ctx.size());
  gapit_mgeno* g = new gapit_mgeno();
  g->mc = mc;
  g->n = n;
  g->m = m;
  g->part.assign(world, nullptr);
  g->lo.resize(world + 1);
  for (int r = 0; r <= world; ++r) g->lo[r] = (m * r) / world;   // contiguous balanced blocks: marker order = rank order
  const size_t esz = dtype == GAPIT_F64 ? 8 : (dtype == GAPIT_I32 ? 4 : 1);   // GAPIT_B2: ld counts bytes per marker
  const int rc = for_each_gpu(world, [&](int r) {
    const int64_t mr = g->lo[r + 1] - g->lo[r];
    if (mr <= 0) return GAPIT_OK;
    const char* src = static_cast<const char*>(snps) + size_t(g->lo[r]) * size_t(ld) * esz;
    return gapit_geno_pack(mc->ctx[r], src, dtype, n, mr, ld, impute, &g->part[r]);
  });
  if (rc != GAPIT_OK) {
    gapit_mgeno_free(g);
    return rc;
  }
  *out = g;
  return GAPIT_OK;
}

extern "C" int64_t gapit_mgeno_n(const gapit_mgeno* g) { return g ? g->n : 0; }
extern "C" int64_t gapit_mgeno_m(const gapit_mgeno* g) { return g ? g->m : 0; }

extern "C" int gapit_multi_vanraden(gapit_mctx* mc, gapit_mgeno* g, double* K_out, int32_t* G_out) {
  if (!mc || !g || !K_out || g->mc != mc) return fail(GAPIT_ERR_ARG, "gapit_multi_vanraden: bad argument");
  const int world = static_cast<int>(mc->ctx.size());
  const int64_t n = g->n, ldg = gapit_gram_ld(n);
  std::vector<int32_t*> G(world, nullptr);
  std::vector<int64_t*> sums(world, nullptr);
  auto release = [&](int rc) {
    for (int r = 0; r < world; ++r) {
      cudaSetDevice(mc->ctx[r]->device);
      cudaFree(G[r]);
      cudaFree(sums[r]);
    }
    return rc;
  };
  // zeroed Gram buffers everywhere before any rank's kernel may add into them (the join is the barrier)
  int rc = for_each_gpu(world, [&](int r) {
    gapit_ctx* c = mc->ctx[r];
    GAPIT_CUDA(cudaSetDevice(c->device));
    GAPIT_CUDA(cudaMalloc(&G[r], size_t(ldg) * ldg * 4));
    GAPIT_CUDA(cudaMalloc(&sums[r], 6 * 8));                 // [0..2] mailbox for the totals, [3..5] this GPU's partials
    GAPIT_CUDA(cudaMemsetAsync(G[r], 0, size_t(ldg) * ldg * 4, c->stream));
    GAPIT_CUDA(cudaMemsetAsync(sums[r], 0, 6 * 8, c->stream));
    GAPIT_CUDA(cudaStreamSynchronize(c->stream));
    return GAPIT_OK;
  });
  if (rc != GAPIT_OK) return release(rc);
  // every GPU: partial cross-product of its marker block, tiles reduce-scattered to the block-row owners over NVLink,
  // the scalar sums added into every GPU's mailbox
  rc = for_each_gpu(world, [&](int r) {
    if (!g->part[r]) return GAPIT_OK;
    return gapit_vanraden_gram_fused(mc->ctx[r], g->part[r], G.data(), world, r, world > 1 ? -1 : 0, sums[r] + 3, sums.data());
  });
  if (rc != GAPIT_OK) return release(rc);
  // GPU 0: gather the block-rows, exact epilogue
  gapit_ctx* c0 = mc->ctx[0];
  rc = gapit_vanraden_gram_gather(c0, n, G.data(), world, 0);
  if (rc != GAPIT_OK) return release(rc);
  rc = gapit_vanraden_finish(c0, n, G[0], sums[0], nullptr, K_out, G_out);
  return release(rc);
}

// eigen(K) on the GPUs of the context, left ON the devices.  world_eff GPUs (those with >= 1,024 eigenpairs each) solve
// for one slice of the spectrum each (cuSOLVER syevdx by index range; the tridiagonalisation -- 87 % of syevd at
// n = 20,000, profiles/r2_syevd_breakdown.log -- is repeated on every GPU, the rest is shared out), the slices are
// exchanged GPU to GPU over NVLink, every GPU ends with the complete V (n x n column-major, descending) in Vd[r].
static int decompose_on_gpus(gapit_mctx* mc, const double* A, int64_t n, double* values_out, std::vector<double*>& Vd) {
  const int world = static_cast<int>(mc->ctx.size());
  const int solvers = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(world, n / 1024)));
  std::vector<double*> Ad(world, nullptr);
  Vd.assign(world, nullptr);
  auto release = [&](int rc) {
    for (int r = 0; r < world; ++r) {
      cudaSetDevice(mc->ctx[r]->device);
      cudaFree(Ad[r]);
      if (rc != GAPIT_OK) {
        cudaFree(Vd[r]);
        Vd[r] = nullptr;
      }
    }
    return rc;
  };
  gapit_ctx* c0 = mc->ctx[0];
  GAPIT_CUDA(cudaSetDevice(c0->device));
  GAPIT_CUDA(cudaMalloc(&Ad[0], size_t(n) * n * 8));
  cudaError_t e0 = cudaMemcpyAsync(Ad[0], A, size_t(n) * n * 8, cudaMemcpyHostToDevice, c0->stream);
  if (e0 == cudaSuccess) e0 = cudaStreamSynchronize(c0->stream);
  if (e0 != cudaSuccess) return release(fail(GAPIT_ERR_CUDA, std::string("eigen(K) upload: ") + cudaGetErrorString(e0)));
  int rc = for_each_gpu(world, [&](int r) {
    gapit_ctx* c = mc->ctx[r];
    GAPIT_CUDA(cudaSetDevice(c->device));
    GAPIT_CUDA(cudaMalloc(&Vd[r], size_t(n) * n * 8));
    if (r > 0 && r < solvers) {
      GAPIT_CUDA(cudaMalloc(&Ad[r], size_t(n) * n * 8));
      GAPIT_CUDA(cudaMemcpyPeerAsync(Ad[r], c->device, Ad[0], c0->device, size_t(n) * n * 8, c->stream));
      GAPIT_CUDA(cudaStreamSynchronize(c->stream));
    }
    return GAPIT_OK;
  });
  if (rc != GAPIT_OK) return release(rc);                       // (the join: nobody overwrites Ad[0] before it is copied)
  auto first_of = [&](int r) { return n * r / solvers; };
  rc = for_each_gpu(solvers, [&](int r) {
    gapit_ctx* c = mc->ctx[r];
    GAPIT_CUDA(cudaSetDevice(c->device));
    if (solvers == 1) {   // one solver: plain syevd in place, then into Vd[0]
      GAPIT_CHECK(gapit_eigh_dev(c, Ad[0], n, values_out));
      GAPIT_CUDA(cudaMemcpyAsync(Vd[0], Ad[0], size_t(n) * n * 8, cudaMemcpyDeviceToDevice, c->stream));
      GAPIT_CUDA(cudaStreamSynchronize(c->stream));
      return GAPIT_OK;
    }
    const int64_t first = first_of(r), count = first_of(r + 1) - first;
    return gapit_eigh_dev_range(c, Ad[r], n, first, count, values_out + first, Vd[r] + first * n);
  });
  if (rc != GAPIT_OK) return release(rc);
  // every GPU pulls the slices it does not hold from the GPUs that solved them
  rc = for_each_gpu(world, [&](int r) {
    gapit_ctx* c = mc->ctx[r];
    GAPIT_CUDA(cudaSetDevice(c->device));
    for (int k = 1; k <= solvers; ++k) {
      const int s = (r + k) % solvers;                           // staggered sources
      if (s == r) continue;
      const int64_t first = first_of(s), count = first_of(s + 1) - first;
      GAPIT_CUDA(cudaMemcpyPeerAsync(Vd[r] + first * n, c->device, Vd[s] + first * n, mc->ctx[s]->device,
                                     size_t(n) * count * 8, c->stream));
    }
    GAPIT_CUDA(cudaStreamSynchronize(c->stream));
    return GAPIT_OK;
  });
  return release(rc);
}

extern "C" void gapit_meigsys_free(gapit_meigsys* me) {
  if (!me) return;
  for (gapit_eigsys* e : me->es) gapit_eigsys_free(e);
  delete me;
}

static int meigsys_from_device_copies(gapit_mctx* mc, std::vector<double*>& Vd, const double* values, int64_t n,
                                      gapit_meigsys** out) {
  const int world = static_cast<int>(mc->ctx.size());
  gapit_meigsys* me = new gapit_meigsys();
  me->mc = mc;
  me->n = n;
  me->es.assign(world, nullptr);
  const int rc = for_each_gpu(world, [&](int r) {
    gapit_ctx* c = mc->ctx[r];
    GAPIT_CUDA(cudaSetDevice(c->device));
    const int rc2 = gapit_eigsys_from_device(c, Vd[r], values, n, &me->es[r]);
    cudaFree(Vd[r]);
    Vd[r] = nullptr;
    return rc2;
  });
  if (rc != GAPIT_OK) {
    for (int r = 0; r < world; ++r)
      if (Vd[r]) {
        cudaSetDevice(mc->ctx[r]->device);
        cudaFree(Vd[r]);
      }
    gapit_meigsys_free(me);
    return rc;
  }
  *out = me;
  return GAPIT_OK;
}

extern "C" int gapit_meigsys_from_kinship(gapit_mctx* mc, const double* K, int64_t n, double* values_out, gapit_meigsys** out) {
  if (!mc || !K || !values_out || !out || n <= 0) return fail(GAPIT_ERR_ARG, "gapit_meigsys_from_kinship: bad argument");
  std::vector<double*> Vd;
  GAPIT_CHECK(decompose_on_gpus(mc, K, n, values_out, Vd));
  return meigsys_from_device_copies(mc, Vd, values_out, n, out);
}

// the same handle from an eigen-system the host already holds: one PCIe upload, then GPU-to-GPU copies over NVLink
extern "C" int gapit_meigsys_upload(gapit_mctx* mc, const double* vectors, const double* values, int64_t n, gapit_meigsys** out) {
  if (!mc || !vectors || !values || !out || n <= 0) return fail(GAPIT_ERR_ARG, "gapit_meigsys_upload: bad argument");
  const int world = static_cast<int>(mc->ctx.size());
  std::vector<double*> Vd(world, nullptr);
  gapit_ctx* c0 = mc->ctx[0];
  GAPIT_CUDA(cudaSetDevice(c0->device));
  GAPIT_CUDA(cudaMalloc(&Vd[0], size_t(n) * n * 8));
  cudaError_t e0 = cudaMemcpyAsync(Vd[0], vectors, size_t(n) * n * 8, cudaMemcpyHostToDevice, c0->stream);
  if (e0 == cudaSuccess) e0 = cudaStreamSynchronize(c0->stream);
  int rc = e0 == cudaSuccess ? GAPIT_OK : fail(GAPIT_ERR_CUDA, std::string("eigenvector upload: ") + cudaGetErrorString(e0));
  if (rc == GAPIT_OK)
    rc = for_each_gpu(world, [&](int r) {
      if (r == 0) return GAPIT_OK;
      gapit_ctx* c = mc->ctx[r];
      GAPIT_CUDA(cudaSetDevice(c->device));
      GAPIT_CUDA(cudaMalloc(&Vd[r], size_t(n) * n * 8));
      GAPIT_CUDA(cudaMemcpyPeerAsync(Vd[r], c->device, Vd[0], c0->device, size_t(n) * n * 8, c->stream));
      GAPIT_CUDA(cudaStreamSynchronize(c->stream));
      return GAPIT_OK;
    });
  if (rc != GAPIT_OK) {
    for (int r = 0; r < world; ++r) {
      cudaSetDevice(mc->ctx[r]->device);
      cudaFree(Vd[r]);
    }
    return rc;
  }
  return meigsys_from_device_copies(mc, Vd, values, n, out);
}

extern "C" gapit_eigsys* gapit_meigsys_part(gapit_meigsys* me, int gpu) {
  return (me && gpu >= 0 && gpu < static_cast<int>(me->es.size())) ? me->es[gpu] : nullptr;
}

// eigen(K) back on the host (values descending, vectors n x n column-major): GPU r returns its share of the columns over
// its own PCIe link
extern "C" int gapit_multi_eigh(gapit_mctx* mc, const double* A, int64_t n, double* values_out, double* vectors_out) {
  if (!mc || !A || !values_out || !vectors_out || n <= 0) return fail(GAPIT_ERR_ARG, "gapit_multi_eigh: bad argument");
  const int world = static_cast<int>(mc->ctx.size());
  if (world < 2 || n < 2048) return gapit_eigh(mc->ctx[0], A, n, values_out, vectors_out);
  std::vector<double*> Vd;
  GAPIT_CHECK(decompose_on_gpus(mc, A, n, values_out, Vd));
  const int rc = for_each_gpu(world, [&](int r) {
    gapit_ctx* c = mc->ctx[r];
    GAPIT_CUDA(cudaSetDevice(c->device));
    const int64_t first = n * r / world, count = n * (r + 1) / world - first;
    GAPIT_CUDA(cudaMemcpyAsync(vectors_out + first * n, Vd[r] + first * n, size_t(n) * count * 8, cudaMemcpyDeviceToHost, c->stream));
    GAPIT_CUDA(cudaStreamSynchronize(c->stream));
    return GAPIT_OK;
  });
  for (int r = 0; r < world; ++r) {
    cudaSetDevice(mc->ctx[r]->device);
    cudaFree(Vd[r]);
  }
  return rc;
}

// the scan of every marker block on its GPU against the resident eigen-system (NULL: the GLM branch)
extern "C" int gapit_multi_p3d_scan_es(gapit_mctx* mc, gapit_mgeno* g, gapit_meigsys* me, const double* X0, int q0,
                                       const double* Y, int T, const double* delta, const double* vg, int glm,
                                       gapit_scan_out* out, gapit_scan_base* base_out) {
  if (!mc || !g || g->mc != mc || !out) return fail(GAPIT_ERR_ARG, "gapit_multi_p3d_scan_es: bad argument");
  if (!glm && (!me || me->mc != mc || me->n != g->n))
    return fail(GAPIT_ERR_ARG, "gapit_multi_p3d_scan_es: MLM needs the eigen-system of K on this multi-context");
  const int world = static_cast<int>(mc->ctx.size());
  std::vector<std::vector<gapit_scan_base>> bases(world, std::vector<gapit_scan_base>(T));
  const int rc = for_each_gpu(world, [&](int r) {
    if (!g->part[r]) return GAPIT_OK;
    return scan_impl(mc->ctx[r], g->part[r], glm ? nullptr : me->es[r], X0, q0, Y, T, delta, vg, glm, out, bases[r].data(), g->m,
                     g->lo[r]);
  });
  if (rc == GAPIT_OK && base_out)
    for (int r = 0; r < world; ++r)
      if (g->part[r]) {   // the marker-free model is the same on every GPU
        std::copy(bases[r].begin(), bases[r].end(), base_out);
        break;
      }
  return rc;
}

extern "C" int gapit_multi_p3d_scan(gapit_mctx* mc, gapit_mgeno* g, const double* vectors, const double* values,
                                    const double* X0, int q0, const double* Y, int T, const double* delta, const double* vg,
                                    int glm, gapit_scan_out* out, gapit_scan_base* base_out) {
  if (!mc || !g || g->mc != mc || !out) return fail(GAPIT_ERR_ARG, "gapit_multi_p3d_scan: bad argument");
  if (!glm && (!vectors || !values)) return fail(GAPIT_ERR_ARG, "gapit_multi_p3d_scan: MLM needs the eigen-system of K");
  gapit_meigsys* me = nullptr;
  if (!glm) GAPIT_CHECK(gapit_meigsys_upload(mc, vectors, values, g->n, &me));
  const int rc = gapit_multi_p3d_scan_es(mc, g, me, X0, q0, Y, T, delta, vg, glm, out, base_out);
  gapit_meigsys_free(me);
  return rc;
}
