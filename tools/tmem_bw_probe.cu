// TMEM -> register bandwidth of tcgen05.ld on sm_100a: W warps per lane quarter issue back-to-back 32x32b.x32 loads
// (4 KB per warp and instruction) of a 512-column allocation; reports bytes per clock and SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tmem_bw_probe tools/tmem_bw_probe.cu
#include "../pointcloudcounterfactual_b200/csrc/tc_ptx.cuh"

#include <stdio.h>
namespace pcc {
std::atomic<uint64_t> g_launches{0};
std::atomic<uint64_t> g_routes[R_COUNT];
cudaError_t ws_alloc(void **ptr, size_t bytes, cudaStream_t st) { return cudaMallocAsync(ptr, bytes, st); }
}  // namespace pcc
using namespace pcc;

template <int X>
__device__ __forceinline__ void ld_x(uint32_t taddr, uint32_t &acc);
template <>
__device__ __forceinline__ void ld_x<32>(uint32_t taddr, uint32_t &acc) {
  uint32_t r[32];
  tmem_ld32_issue(taddr, r);
  tmem_ld_wait_dep(r);
#pragma unroll
  for (int i = 0; i < 32; ++i) acc ^= r[i];
}
template <>
__device__ __forceinline__ void ld_x<64>(uint32_t taddr, uint32_t &acc) {  // two loads in flight, one wait
  uint32_t r[32], q[32];
  tmem_ld32_issue(taddr, r);
  tmem_ld32_issue(taddr + 32, q);
  tmem_ld_wait_dep(r);
  tmem_ld_wait_dep(q);
#pragma unroll
  for (int i = 0; i < 32; ++i) acc ^= r[i] ^ q[i];
}

// 64 columns of 16-bit data packed two per register (tcgen05.ld ... .pack::16b): 32 registers for 64 columns
template <>
__device__ __forceinline__ void ld_x<128>(uint32_t taddr, uint32_t &acc) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.pack::16b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  tmem_ld_wait_dep(r);
#pragma unroll
  for (int i = 0; i < 32; ++i) acc ^= r[i];
}

template <int X>
__global__ void __launch_bounds__(1024, 1) bw_kernel(int iters, unsigned long long *cycles, unsigned int *sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) ld_x<X>(base + (uint32_t)(((i * 8 + (warp >> 2)) * (X > 64 ? 64 : X)) & 511 & ~((X > 64 ? 64 : X) - 1)), acc);
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = (unsigned long long)(t1 - t0);
  if (acc == 0x12345678u) sink[0] = acc;
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

int main() {
  unsigned long long *cyc;
  unsigned int *sink;
  cudaMalloc(&cyc, 8);
  cudaMalloc(&sink, 4);
  const int iters = 2000;
  for (int warps : {4, 8, 16, 32}) {
    for (int x : {32, 64, 128}) {
      if (x == 32) bw_kernel<32><<<148, warps * 32>>>(iters, cyc, sink);
      else if (x == 64) bw_kernel<64><<<148, warps * 32>>>(iters, cyc, sink);
      else bw_kernel<128><<<148, warps * 32>>>(iters, cyc, sink);
      cudaError_t e = cudaDeviceSynchronize();
      unsigned long long h = 0;
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)iters * warps * 32 * (x > 64 ? 64 : x) * 4;  // x = 128: 64 columns, packed to 16 bit
      printf("%2d warps (%d per lane quarter), %s: %llu cycles, %.1f TMEM bytes/clk/SM = %.1f columns x lanes/clk/SM (%s)\n", warps,
             warps / 4, x == 32 ? "x32" : x == 64 ? "2 x32 in flight" : "x64.pack::16b", h, bytes / (double)h, bytes / 4 / (double)h, cudaGetErrorString(e));
    }
  }
  return 0;
}
