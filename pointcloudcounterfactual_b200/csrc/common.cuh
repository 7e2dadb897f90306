// Shared device helpers for libpcc_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/pcc_b200.h"

namespace pcc {

extern std::atomic<uint64_t> g_launches;  // bumped once per kernel launch (pcc_launch_count)

inline int finish_launch(int nkernels) {
  g_launches.fetch_add((uint64_t)nkernels, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

// Which kernel family an entry point dispatched to (pcc_route_count): the tests and bench.py use it to PROVE that a call
// made through the reference's unchanged call sites reached the fast kernels and not the generic SIMT path.
enum Route {
  R_KNN3W = 0,      // warp-cooperative xyz kNN (knn3w_kernel)
  R_KNN3_THREAD,    // one-thread-per-query xyz kNN (knn3_kernel) as the primary kernel
  R_KNN_TC2,        // tcgen05 feature kNN, second generation
  R_KNN_TC1,        // tcgen05 feature kNN, first generation
  R_KNN_SIMT,       // generic exact SIMT kernel (knn_kernel)
  R_ARGMIN_SMALL,   // one-thread-per-query argmin (vector quantisation)
  R_NN_SYM,         // symmetric brute-force Chamfer forward (nn_sym_kernel)
  R_NN_ASYM,        // one-direction Chamfer forward (nn_fwd_kernel)
  R_NN_TC,          // tcgen05 candidate filter + exact resolution (chamfer_tc.cu)
  R_KNN3_TC,        // xyz kNN with the tcgen05 candidate filter (knn3_tc.cu)
  R_PM_SELF,        // pcc_argkmin recognised q == r (point-major self kNN)
  R_KNN_BF,         // feature kNN, bf16-split precise scores on tcgen05 (knn_bf.cu)
  R_COUNT
};
extern std::atomic<uint64_t> g_routes[R_COUNT];
inline void note_route(Route r) { g_routes[r].fetch_add(1, std::memory_order_relaxed); }

// Stream-ordered workspace allocation from a pool PRIVATE to this library (one per device, created on first use; lib.cu).
// Freed workspaces stay cached in it (release threshold = max: a default-threshold pool returns the memory to the OS at
// every synchronisation and the next allocation costs milliseconds) -- but only this library's own scratch, never the
// device's default pool that other libraries and captured graphs allocate from.  Release with cudaFreeAsync.
cudaError_t ws_alloc(void **ptr, size_t bytes, cudaStream_t st);

// Spatial grouping of both clouds of every pair (approxmatch.cu: am_group_kernel; declared here for other translation units): perm1 (b,n) / perm2 (b,m) list each
// cloud's points so that consecutive blocks of 64 are spatially tight.  n, m <= 4096 (PCC_ENOTSUP beyond).  temp may
// be null (otherwise the approxmatch remain vectors are initialised on the way).
int am_group_launch(int b, int n, int m, const float *xyz1, const float *xyz2, unsigned short *perm1,
                    unsigned short *perm2, float *temp, float multiL, float multiR, cudaStream_t st);

// Opt-in to more than 48 KiB of dynamic shared memory.  The attribute is PER DEVICE: one process may drive several GPUs
// (the Python wrappers take tensors on any device), so the high-water mark is kept per device ordinal.
// The 48 KiB limit without opt-in counts STATIC shared memory too, so the opt-in starts well below it (at exactly 48 KiB of
// dynamic memory -- n = 3072 in the auction kernels -- a kernel with a few static words was rejected at launch).
template <typename Kernel>
inline cudaError_t smem_optin(Kernel kernel, size_t bytes, size_t (&done)[64], size_t threshold = 32 * 1024) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  size_t &d = done[dev & 63];
  if (bytes > threshold && bytes > d) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    d = bytes;
  }
  return cudaSuccess;
}

// ---- programmatic dependent launch (experimental build variant: nvcc -DPCC_PDL, tools/build_variant.sh) -------
// A kernel launched with the programmatic-stream-serialization attribute may become resident while its predecessor
// in the stream still runs; pdl_wait() -- the FIRST statement of every kernel launched through PCC_LAUNCH -- blocks
// until the predecessor has completed and its writes are visible (griddepcontrol.wait); pdl_trigger() lets the
// successor become resident.  Short kernels trigger at once (pdl_enter); the solver sweeps trigger after their main
// loop, when the SMs are draining: a successor placed while every CTA of the sweep is still resident lands on the
// SMs the sweep loaded least and ends up three deep on some of them.  Only launch latency and CTA scheduling
// overlap: no kernel touches memory before the wait.
// The PCC_PDL_MASK environment variable selects which kernel classes get the attribute (default PCC_PDL_DEFAULT_MASK).
// Without -DPCC_PDL the macro expands to the plain <<<>>> launch and pdl_enter() to nothing.
constexpr unsigned PDL_CHAMFER = 1u, PDL_EMD_SMALL = 2u, PDL_EMD_SWEEP = 4u;
#define PCC_K(...) __VA_ARGS__  // protects the commas of a template-id inside PCC_LAUNCH
#ifdef PCC_PDL
#ifndef PCC_PDL_DEFAULT_MASK
#define PCC_PDL_DEFAULT_MASK 0
#endif
unsigned pdl_mask();  // lib.cu
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline void launch_pdl(unsigned cls, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                       Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_mask() & cls) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define PCC_LAUNCH(cls, k, g, b, s, st, ...) pcc::launch_pdl(cls, k, dim3(g), dim3(b), s, st, __VA_ARGS__)
#else
__device__ __forceinline__ void pdl_wait() {}
__device__ __forceinline__ void pdl_trigger() {}
#define PCC_LAUNCH(cls, k, g, b, s, st, ...) k<<<g, b, s, st>>>(__VA_ARGS__)
#endif
__device__ __forceinline__ void pdl_enter() {
  pdl_wait();
  pdl_trigger();
}

// ---- packed fp32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2) ------------------------------------------
// A f32x2 value lives in a 64-bit register pair; .x is the low half.  Each op rounds both lanes exactly like
// the scalar .rn instruction, so results are bit-identical to scalar FADD/FMUL/FFMA.
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// Squared distance of two reference points (packed) to one query, canonical order of the reference kernels
// (nndistance.cu:22-25 as compiled): d = fma(dz,dz, fma(dx,dx, dy*dy)),  d* = ref - query.
// nq* hold the NEGATED query coordinate in both lanes.
__device__ __forceinline__ f32x2 sqdist2(f32x2 rx, f32x2 ry, f32x2 rz, f32x2 nqx, f32x2 nqy, f32x2 nqz) {
  f32x2 dx = add2(rx, nqx), dy = add2(ry, nqy), dz = add2(rz, nqz);
  return fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
}

__device__ __forceinline__ float sqdist1(float qx, float qy, float qz, float rx, float ry, float rz) {
  float dx = rx - qx, dy = ry - qy, dz = rz - qz;
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ float ex2_ftz(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rsqrt_ftz(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// knn_tc.cu: tcgen05 candidate generator + exact re-rank; PCC_ENOTSUP when the shape is outside that path
int knn_tc_launch(int b, int c, int n, int k, bool pm, const float *x, int64_t *idx, float *dist, cudaStream_t st);
// knn_bf.cu: feature kNN (C = 32 / 64, indices only) with bf16-split scores that order the candidates
int knn_bf_launch(int b, int c, int n, int k, bool pm, const float *x, int64_t *idx, float *dist, cudaStream_t st);
// knn3_tc.cu: xyz kNN through the fp16 tensor-core candidate filter; PCC_ENOTSUP outside 256 <= n <= 2048, k <= 32
int knn3_tc_launch(int b, int n, int k, bool pm, const float *x, int64_t *idx, float *dist, cudaStream_t st);

// edgeconv.cu: edge lists sorted by TARGET (packed target << 19 | source << 6 | slot, ascending inside a run;
// off[n + 1] = run starts), one list per (cloud, piece of the source range) -- see edge_sort_launch.  Four launches on
// st; the workspace is the caller's.
// interleave: the consumer is graph.cu's gather backward -- inside every block of 512 entries the q-th 16-byte group of
// all 32 sixteen-entry chunks is contiguous (edge_pos), so a warp whose lanes own consecutive chunks fetches them with
// coalesced 16-byte loads, and an entry is (edge id within the piece * 4) << 12 | target: the byte offset of the edge's
// gradient inside the piece's plane needs one shift (n <= 4096, n / pieces * k <= 2^18).
size_t edge_sort_ws_bytes(int b, int n, int k, int pieces);
bool edge_sort_ok(int b, int n, int k, int pieces);
void edge_sort_launch(int b, int n, int k, int pieces, const int64_t *idx, char *ws, bool interleave, const int **off,
                      const unsigned int **rev, int *stride, cudaStream_t st);
void edge_sort_views(int b, int n, int k, int pieces, const char *ws, const int **off, const unsigned int **rev, int *stride);
__host__ __device__ inline int edge_pos(int p) {
  const int r = p & 511, l = r >> 4, q = r & 15;
  return (p & ~511) | ((q >> 2) << 7) | (l << 2) | (q & 3);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace pcc
