// Library-level entry points of libpcc_b200.so.
#include "common.cuh"

#ifdef PCC_PDL
#include <stdlib.h>
#endif

namespace pcc {
std::atomic<uint64_t> g_launches{0};
#ifdef PCC_PDL
unsigned pdl_mask() {  // read at every launch (captured graphs keep what was set at capture time): A/B runs in one process
  const char *e = getenv("PCC_PDL_MASK");
  return e ? (unsigned)strtoul(e, nullptr, 0) : (unsigned)PCC_PDL_DEFAULT_MASK;
}
#endif
}

extern "C" __attribute__((visibility("default"))) const char *pcc_version(void) { return "pcc_b200 0.1 (sm_100a)"; }

extern "C" __attribute__((visibility("default"))) uint64_t pcc_launch_count(void) { return pcc::g_launches.load(std::memory_order_relaxed); }

extern "C" __attribute__((visibility("default"))) const char *pcc_status_string(int status) {
  switch (status) {
    case PCC_OK:
      return "ok";
    case PCC_EINVAL:
      return "invalid shape (pcc: shape rule of the operator violated)";
    case PCC_ENOTSUP:
      return "unsupported configuration (pcc: outside kernel limits)";
    default:
      return cudaGetErrorString((cudaError_t)status);
  }
}
