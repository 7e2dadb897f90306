// Library-level entry points of libpcc_b200.so.
#include "common.cuh"

namespace pcc {
std::atomic<uint64_t> g_launches{0};
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
