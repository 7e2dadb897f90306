// Library-level entry points of libpcc_b200.so.
#include "common.cuh"

#include <string.h>

#include <mutex>

#ifdef PCC_PDL
#include <stdlib.h>
#endif

namespace pcc {
std::atomic<uint64_t> g_launches{0};
std::atomic<uint64_t> g_routes[R_COUNT];
static const char *const kRouteNames[R_COUNT] = {"knn3w", "knn3_thread", "knn_tc2", "knn_tc1", "knn_simt", "argmin_small",
                                                 "nn_sym", "nn_asym", "nn_tc", "knn3_tc", "pm_self", "knn_bf"};
cudaError_t ws_alloc(void **ptr, size_t bytes, cudaStream_t st) {
  static std::mutex mu;
  static cudaMemPool_t pools[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaMallocAsync(ptr, bytes, st);
  cudaMemPool_t pool;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!pools[dev]) {
      cudaMemPoolProps props = {};
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = dev;
      e = cudaMemPoolCreate(&pools[dev], &props);
      if (e != cudaSuccess) {
        pools[dev] = nullptr;
        return e;
      }
      uint64_t thr = UINT64_MAX;
      cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &thr);
    }
    pool = pools[dev];
  }
  return cudaMallocFromPoolAsync(ptr, bytes, pool, st);
}

#ifdef PCC_PDL
unsigned pdl_mask() {  // read at every launch (captured graphs keep what was set at capture time): A/B runs in one process
  const char *e = getenv("PCC_PDL_MASK");
  return e ? (unsigned)strtoul(e, nullptr, 0) : (unsigned)PCC_PDL_DEFAULT_MASK;
}
#endif
}

extern "C" __attribute__((visibility("default"))) const char *pcc_version(void) { return "pcc_b200 0.1 (sm_100a)"; }

extern "C" __attribute__((visibility("default"))) uint64_t pcc_launch_count(void) { return pcc::g_launches.load(std::memory_order_relaxed); }

extern "C" __attribute__((visibility("default"))) const char *pcc_route_names(void) {
  return "knn3w,knn3_thread,knn_tc2,knn_tc1,knn_simt,argmin_small,nn_sym,nn_asym,nn_tc,knn3_tc,pm_self,knn_bf";
}

extern "C" __attribute__((visibility("default"))) int64_t pcc_route_count(const char *name) {
  if (!name) return -1;
  for (int i = 0; i < pcc::R_COUNT; ++i)
    if (strcmp(name, pcc::kRouteNames[i]) == 0) return (int64_t)pcc::g_routes[i].load(std::memory_order_relaxed);
  return -1;
}

extern "C" __attribute__((visibility("default"))) const char *pcc_status_string(int status) {
  switch (status) {
    case PCC_OK:
      return "ok";
    case PCC_EINVAL:
      return "invalid shape (pcc: shape rule of the operator violated)";
    case PCC_ENOTSUP:
      return "unsupported configuration (pcc: outside kernel limits)";
    case PCC_ELAUNCH:
      return "kernel launch rejected (cudaErrorInvalidValue)";
    default:
      return cudaGetErrorString((cudaError_t)status);
  }
}
