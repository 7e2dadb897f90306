// How fast can one SM sub-partition run the EMD sweep's instruction mix (8 packed FMA-pipe ops + 2 MUFU.EX2 per partner and
// thread) as a function of resident warps and of the independent partners per step?  Variants: ordered accumulation (one
// dependent FFMA2 chain, as the solver needs) and a broadcast LDS.128 per partner.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/emd_mix_probe tools/emd_mix_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float ex2(float x) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

constexpr int TILE = 2048, REPS = 8;

// MODE 0: the real loop shape: partners from shared memory (broadcast LDS.128), own points in registers, ordered accumulation
template <int U>
__global__ void __launch_bounds__(1024) k_sweep(float *out, const float *in, float lc) {
  __shared__ float4 tile[TILE];
  for (int i = threadIdx.x; i < TILE; i += blockDim.x) tile[i] = make_float4(in[i % 97], in[(i * 3) % 89], in[(i * 7) % 83], in[(i * 5) % 79]);
  __syncthreads();
  const f32x2 npx = pack2(-in[threadIdx.x % 64], -in[threadIdx.x % 61 + 1]), npy = pack2(-in[threadIdx.x % 59 + 2], -in[threadIdx.x % 53 + 3]),
              npz = pack2(-in[threadIdx.x % 47 + 4], -in[threadIdx.x % 43 + 5]), lc2 = pack2(lc, lc);
  f32x2 acc = 0ull;
  for (int rep = 0; rep < REPS; ++rep)
#pragma unroll 1
    for (int l = 0; l < TILE; l += U) {
      f32x2 E[U];
      float w[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float4 q = tile[l + u];
        w[u] = q.w;
        const f32x2 qx = pack2(q.x, q.x), qy = pack2(q.y, q.y), qz = pack2(q.z, q.z);
        const f32x2 dx = add2(qx, npx), dy = add2(qy, npy), dz = add2(qz, npz);
        float a, b;
        unpack2(mul2(fma2(dz, dz, fma2(dx, dx, mul2(dy, dy))), lc2), a, b);
        E[u] = pack2(ex2(a), ex2(b));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) acc = fma2(E[u], pack2(w[u], w[u]), acc);
    }
  float lo, hi;
  unpack2(acc, lo, hi);
  out[blockIdx.x * blockDim.x + threadIdx.x] = lo + hi;
}

// MODE 1: one own point per thread, TWO PARTNERS per packed instruction (tile holds partner pairs as (x0,x1,y0,y1),
// (z0,z1,w0,w1)); the accumulation is two ordered scalar FFMAs.  Twice the warps for the same work.
template <int U>  // partner pairs per step
__global__ void __launch_bounds__(1024) k_sweep_pp(float *out, const float *in, float lc) {
  __shared__ float4 tile[TILE];  // TILE/2 pairs x 2 float4
  for (int i = threadIdx.x; i < TILE; i += blockDim.x) tile[i] = make_float4(in[i % 97], in[(i * 3) % 89], in[(i * 7) % 83], in[(i * 5) % 79]);
  __syncthreads();
  const float px = -in[threadIdx.x % 64], py = -in[threadIdx.x % 59 + 2], pz = -in[threadIdx.x % 47 + 4];
  const f32x2 npx = pack2(px, px), npy = pack2(py, py), npz = pack2(pz, pz), lc2 = pack2(lc, lc);
  float acc = 0.f;
  for (int rep = 0; rep < REPS; ++rep)
#pragma unroll 1
    for (int l = 0; l < TILE / 2; l += U) {
      f32x2 E[U];
      f32x2 w[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float4 a4 = tile[2 * (l + u)], b4 = tile[2 * (l + u) + 1];
        w[u] = pack2(b4.z, b4.w);
        const f32x2 dx = add2(pack2(a4.x, a4.y), npx), dy = add2(pack2(a4.z, a4.w), npy), dz = add2(pack2(b4.x, b4.y), npz);
        float a, b;
        unpack2(mul2(fma2(dz, dz, fma2(dx, dx, mul2(dy, dy))), lc2), a, b);
        E[u] = pack2(ex2(a), ex2(b));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float e0, e1, w0, w1;
        unpack2(E[u], e0, e1);
        unpack2(w[u], w0, w1);
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc) : "f"(e0), "f"(w0));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc) : "f"(e1), "f"(w1));
      }
    }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F> static float time_ms(F launch) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) launch();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  float *out, *in; cudaMalloc(&out, sizeof(float) * sms * 1024); cudaMalloc(&in, 4096);
  float h[128]; for (int i = 0; i < 128; ++i) h[i] = 0.01f * i; cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  printf("EMD sweep loop (LDS.128 broadcast + ordered accumulation), Gexp/s on the whole GPU; MUFU peak 4627\n");
  for (int wps : {1, 2, 3, 4, 8}) {
    const int threads = 128 * wps;
    const double evals = (double)sms * threads * 2.0 * TILE * REPS;
    float t8 = time_ms([&] { k_sweep<8><<<sms, threads>>>(out, in, -1.5f); });
    float t16 = time_ms([&] { k_sweep<16><<<sms, threads>>>(out, in, -1.5f); });
    float t32 = time_ms([&] { k_sweep<32><<<sms, threads>>>(out, in, -1.5f); });
    printf("warps/SMSP %d:  U=8 %7.0f   U=16 %7.0f   U=32 %7.0f\n", wps, evals / t8 / 1e6, evals / t16 / 1e6, evals / t32 / 1e6);
  }
  printf("one point per thread, two partners per packed op (same work needs twice the warps)\n");
  for (int wps : {2, 3, 4, 6, 8}) {
    const int threads = 128 * wps;
    const double evals = (double)sms * threads * 1.0 * TILE * REPS;
    float t4 = time_ms([&] { k_sweep_pp<4><<<sms, threads>>>(out, in, -1.5f); });
    float t8 = time_ms([&] { k_sweep_pp<8><<<sms, threads>>>(out, in, -1.5f); });
    float t16 = time_ms([&] { k_sweep_pp<16><<<sms, threads>>>(out, in, -1.5f); });
    printf("warps/SMSP %d:  pairs/step 4 %7.0f   8 %7.0f   16 %7.0f\n", wps, evals / t4 / 1e6, evals / t8 / 1e6, evals / t16 / 1e6);
  }
  return cudaDeviceSynchronize() != cudaSuccess;
}
