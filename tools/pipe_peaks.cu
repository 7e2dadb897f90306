// Micro-benchmarks of the SIMT pipes the geometry kernels are bound by (B200, sm_100a):
//   FFMA (scalar), FFMA2 (packed f32x2), the Chamfer inner-loop mix (3 FADD2 + FMUL2 + 2 FFMA2 + FMNMX3 per 2 pairs),
//   MUFU.EX2.  Prints one JSON line; bench.py / DESIGN.md use the numbers as roofline denominators.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_peaks tools/pipe_peaks.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

constexpr int ITERS = 4096, ILP = 8;

__global__ void k_ffma(float *out, float a, float b) {
  float x[ILP];
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
  float s = 0;
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float *out, float a, float b) {
  f32x2 x[ILP];
  const f32x2 aa = pack2(a, a), bb = pack2(b, b);
  for (int i = 0; i < ILP; ++i) x[i] = pack2(threadIdx.x * 1e-3f + i, 1.f);
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = fma2(x[i], aa, bb);
  float s = 0;
  for (int i = 0; i < ILP; ++i) { float lo, hi; unpack2(x[i], lo, hi); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// the Chamfer inner loop: per step 4 reference pairs (8 points) vs one query
__global__ void k_chamfer_mix(float *out, float qx, float qy, float qz) {
  f32x2 rx[4], ry[4], rz[4];
  for (int i = 0; i < 4; ++i) { rx[i] = pack2(threadIdx.x * 1e-3f + i, 0.5f + i); ry[i] = pack2(0.1f * i, threadIdx.x * 2e-3f); rz[i] = pack2(0.3f, 0.7f * i); }
  const f32x2 nqx = pack2(-qx, -qx), nqy = pack2(-qy, -qy), nqz = pack2(-qz, -qz);
  float best = 1e30f;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f32x2 dx = add2(rx[i], nqx), dy = add2(ry[i], nqy), dz = add2(rz[i], nqz);
      f32x2 d = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
      float a, b; unpack2(d, a, b);
      best = fminf(fminf(a, b), best);
      rx[i] = d;  // keep a dependency so nothing is hoisted
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = best;
}
__global__ void k_ex2(float *out, float a) {
  float x[ILP];
  for (int i = 0; i < ILP; ++i) x[i] = -threadIdx.x * 1e-3f - i;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
  float s = a;
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the EMD solver sweep per partner and per thread (two own points packed): 3 FADD2 + FMUL2 + 2 FFMA2 (squared distance),
// FMUL2 (level), 2 MUFU.EX2, FFMA2 (ordered accumulation) -- 8 packed FMA-pipe instructions and 2 SFU instructions for 2
// pair evaluations.  ILP independent partners per step, like the 16-partner blocks of am_sweep_kernel.
__global__ void k_emd_mix(float *out, float qx, float qy, float qz, float lc) {
  f32x2 px[ILP], py[ILP], pz[ILP], acc[ILP];
  for (int i = 0; i < ILP; ++i) {
    px[i] = pack2(threadIdx.x * 1e-3f + i, 0.5f + i);
    py[i] = pack2(0.1f * i, threadIdx.x * 2e-3f);
    pz[i] = pack2(0.3f, 0.7f * i);
    acc[i] = 0ull;
  }
  const f32x2 nqx = pack2(-qx, -qx), nqy = pack2(-qy, -qy), nqz = pack2(-qz, -qz), lc2 = pack2(lc, lc), w2 = pack2(0.5f, 0.25f);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      f32x2 dx = add2(px[i], nqx), dy = add2(py[i], nqy), dz = add2(pz[i], nqz);
      f32x2 d = mul2(fma2(dz, dz, fma2(dx, dx, mul2(dy, dy))), lc2);
      float a, b; unpack2(d, a, b);
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
      const f32x2 e = pack2(a, b);
      acc[i] = fma2(e, w2, acc[i]);
      px[i] = e, py[i] = dx, pz[i] = dy;  // keep dependencies so nothing is hoisted
    }
  }
  float s = 0;
  for (int i = 0; i < ILP; ++i) { float lo, hi; unpack2(acc[i], lo, hi); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> static float time_ms(F launch) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) launch();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 8, threads = 256;
  float *out; cudaMalloc(&out, sizeof(float) * blocks * threads);
  const double lanes = (double)blocks * threads;
  float t1 = time_ms([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 1e-7f); });
  float t2 = time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 1e-7f); });
  float t3 = time_ms([&] { k_chamfer_mix<<<blocks, threads>>>(out, 0.1f, 0.2f, 0.3f); });
  float t4 = time_ms([&] { k_ex2<<<blocks, threads>>>(out, 0.f); });
  // EMD mix: plenty of warps (16 per SM sub-partition), then 2 and 1 warps per sub-partition like the sweep kernel itself
  float t5 = time_ms([&] { k_emd_mix<<<blocks, threads>>>(out, 0.1f, 0.2f, 0.3f, -1.5f); });
  float t6 = time_ms([&] { k_emd_mix<<<p.multiProcessorCount, 256>>>(out, 0.1f, 0.2f, 0.3f, -1.5f); });
  float t7 = time_ms([&] { k_emd_mix<<<p.multiProcessorCount, 128>>>(out, 0.1f, 0.2f, 0.3f, -1.5f); });
  const double emd16 = lanes * ITERS * ILP * 2 / (t5 * 1e-3), emd2 = (double)p.multiProcessorCount * 256 * ITERS * ILP * 2 / (t6 * 1e-3),
               emd1 = (double)p.multiProcessorCount * 128 * ITERS * ILP * 2 / (t7 * 1e-3);
  const double ffma_tflops = lanes * ITERS * ILP * 2 / (t1 * 1e-3) / 1e12;
  const double ffma2_tflops = lanes * ITERS * ILP * 4 / (t2 * 1e-3) / 1e12;
  const double pairs_per_s = lanes * ITERS * 8 / (t3 * 1e-3);
  const double ex2_per_s = lanes * ITERS * ILP / (t4 * 1e-3);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"ffma_tflops\": %.2f, \"ffma2_tflops\": %.2f, \"chamfer_mix_gpairs_per_s\": %.2f, "
         "\"chamfer_mix_flops_equiv_tflops\": %.2f, \"mufu_ex2_gops\": %.2f, \"emd_mix_gexp_per_s\": %.2f, "
         "\"emd_mix_gexp_per_s_2warps_per_smsp\": %.2f, \"emd_mix_gexp_per_s_1warp_per_smsp\": %.2f}\n",
         p.name, p.multiProcessorCount, ffma_tflops, ffma2_tflops, pairs_per_s / 1e9, pairs_per_s * 8 / 1e12, ex2_per_s / 1e9,
         emd16 / 1e9, emd2 / 1e9, emd1 / 1e9);
  return cudaDeviceSynchronize() != cudaSuccess;
}
