// Occupancy / ILP sensitivity of the Chamfer distance mix (3 FADD2 + FMUL2 + 2 FFMA2 + FMNMX3 per 2 pairs) on B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_probe tools/pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
constexpr int ITERS = 4096;
template <int ILP, bool SCALAR, int QMODE = 0>
__global__ void k_mix(float *out, float qx, float qy, float qz) {
  if (QMODE) { qx += threadIdx.x * 1e-4f; qy += threadIdx.x * 2e-4f; qz += threadIdx.x * 3e-4f; }  // per-thread query, like the kernels
  f32x2 rx[ILP], ry[ILP], rz[ILP];
  for (int i = 0; i < ILP; ++i) { rx[i] = pack2(threadIdx.x * 1e-3f + i, 0.5f + i); ry[i] = pack2(0.1f * i, threadIdx.x * 2e-3f); rz[i] = pack2(0.3f, 0.7f * i); }
  f32x2 nqx = pack2(-qx, -qx), nqy = pack2(-qy, -qy), nqz = pack2(-qz, -qz);
  if (QMODE == 2) { asm volatile("" : "+l"(nqx), "+l"(nqy), "+l"(nqz)); }  // opaque: forces the packed-register operand form
  float best = 1e30f;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (!SCALAR) {
        f32x2 dx = add2(rx[i], nqx), dy = add2(ry[i], nqy), dz = add2(rz[i], nqz);
        f32x2 d = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
        float a, b; unpack2(d, a, b);
        best = fminf(fminf(a, b), best);
        rx[i] = d;
      } else {
        float x0, x1, y0, y1, z0, z1; unpack2(rx[i], x0, x1); unpack2(ry[i], y0, y1); unpack2(rz[i], z0, z1);
        float dx0 = x0 - qx, dy0 = y0 - qy, dz0 = z0 - qz, dx1 = x1 - qx, dy1 = y1 - qy, dz1 = z1 - qz;
        float d0 = __fmaf_rn(dz0, dz0, __fmaf_rn(dx0, dx0, __fmul_rn(dy0, dy0)));
        float d1 = __fmaf_rn(dz1, dz1, __fmaf_rn(dx1, dx1, __fmul_rn(dy1, dy1)));
        best = fminf(fminf(d0, d1), best);
        rx[i] = pack2(d0, d1);
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = best;
}
template <typename F> static float time_ms(F launch) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) launch();
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  return best;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  float *out; cudaMalloc(&out, sizeof(float) * p.multiProcessorCount * 8 * 256);
  const int cfg[][2] = {{8, 256}, {4, 256}, {4, 128}, {2, 128}, {1, 128}, {1, 64}};
  for (auto &c : cfg) {
    const int blocks = p.multiProcessorCount * c[0], threads = c[1];
    const double lanes = (double)blocks * threads;
    float t4 = time_ms([&] { k_mix<4, false><<<blocks, threads>>>(out, 0.1f, 0.2f, 0.3f); });
    float t8 = time_ms([&] { k_mix<8, false><<<blocks, threads>>>(out, 0.1f, 0.2f, 0.3f); });
    float t16 = time_ms([&] { k_mix<16, false><<<blocks, threads>>>(out, 0.1f, 0.2f, 0.3f); });
    float s8 = time_ms([&] { k_mix<8, true><<<blocks, threads>>>(out, 0.1f, 0.2f, 0.3f); });
    float q1 = time_ms([&] { k_mix<8, false, 1><<<blocks, threads>>>(out, 0.1f, 0.2f, 0.3f); });
    float q2 = time_ms([&] { k_mix<8, false, 2><<<blocks, threads>>>(out, 0.1f, 0.2f, 0.3f); });
    printf("              per-thread query: broadcast-operand form %.0f   opaque packed operands %.0f\n",
           lanes * ITERS * 16 / (q1 * 1e-3) / 1e9, lanes * ITERS * 16 / (q2 * 1e-3) / 1e9);
    printf("warps/SM %2d: packed ILP4 %.0f  ILP8 %.0f  ILP16 %.0f   scalar ILP8 %.0f  Gpairs/s\n", c[0] * c[1] / 32,
           lanes * ITERS * 8 / (t4 * 1e-3) / 1e9, lanes * ITERS * 16 / (t8 * 1e-3) / 1e9, lanes * ITERS * 32 / (t16 * 1e-3) / 1e9,
           lanes * ITERS * 16 / (s8 * 1e-3) / 1e9);
  }
  return cudaDeviceSynchronize() != cudaSuccess;
}
