// Builds the symmetric Chamfer inner loop up step by step to find what costs the distance mix its throughput (B200).
//   MODE 0  8 rows in registers x 8 columns per step from SHARED memory (6 broadcast LDS.128), row minima only
//   MODE 1  + column minima (FMNMX3 over the lane's rows)
//   MODE 2  + REDUX.MIN and ballot per column
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_probe2 tools/pipe_probe2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 sqd(f32x2 rx, f32x2 ry, f32x2 rz, f32x2 nx, f32x2 ny, f32x2 nz) {
  f32x2 dx = add2(rx, nx), dy = add2(ry, ny), dz = add2(rz, nz);
  return fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
}
constexpr int GROUPS = 256, REPS = 16, RPL = 8;
template <int MODE, bool PREFETCH = false>
__global__ void __launch_bounds__(128) k_sym(float *out, const float *in) {
  __shared__ float4 tile[GROUPS * 6];
  for (int i = threadIdx.x; i < GROUPS * 6; i += 128) tile[i] = make_float4(in[i % 97], in[(i * 3) % 89], in[(i * 7) % 83], in[(i * 5) % 79]);
  __syncthreads();
  float sx[RPL], sy[RPL], sz[RPL], best[RPL];
  for (int u = 0; u < RPL; ++u) { sx[u] = -in[threadIdx.x + u]; sy[u] = -in[threadIdx.x + 8 + u]; sz[u] = -in[threadIdx.x + 16 + u]; best[u] = 1e30f; }
  float colacc = 0.f;
  float4 nX0 = tile[0], nY0 = tile[1], nZ0 = tile[2], nX1 = tile[3], nY1 = tile[4], nZ1 = tile[5];
  for (int rep = 0; rep < REPS; ++rep)
#pragma unroll 1
    for (int g = 0; g < GROUPS; ++g) {
      float4 X0, Y0, Z0, X1, Y1, Z1;
      if (MODE == 3) {
        // opaque to the optimiser (no instruction is emitted): the distances must be recomputed every iteration
        asm volatile("" : "+f"(nX0.x), "+f"(nX0.y), "+f"(nX0.z), "+f"(nX0.w), "+f"(nY0.x), "+f"(nY0.y), "+f"(nY0.z), "+f"(nY0.w));
        asm volatile("" : "+f"(nZ0.x), "+f"(nZ0.y), "+f"(nZ0.z), "+f"(nZ0.w), "+f"(nX1.x), "+f"(nX1.y), "+f"(nX1.z), "+f"(nX1.w));
        asm volatile("" : "+f"(nY1.x), "+f"(nY1.y), "+f"(nY1.z), "+f"(nY1.w), "+f"(nZ1.x), "+f"(nZ1.y), "+f"(nZ1.z), "+f"(nZ1.w));
        X0 = nX0, Y0 = nY0, Z0 = nZ0, X1 = nX1, Y1 = nY1, Z1 = nZ1;
      } else if (PREFETCH) {
        X0 = nX0, Y0 = nY0, Z0 = nZ0, X1 = nX1, Y1 = nY1, Z1 = nZ1;
        const int gn = (g + 1) & (GROUPS - 1);
        nX0 = tile[gn * 6], nY0 = tile[gn * 6 + 1], nZ0 = tile[gn * 6 + 2], nX1 = tile[gn * 6 + 3], nY1 = tile[gn * 6 + 4], nZ1 = tile[gn * 6 + 5];
      } else {
        X0 = tile[g * 6], Y0 = tile[g * 6 + 1], Z0 = tile[g * 6 + 2], X1 = tile[g * 6 + 3], Y1 = tile[g * 6 + 4], Z1 = tile[g * 6 + 5];
      }
      float cm[8];
#pragma unroll
      for (int up = 0; up < RPL / 2; ++up) {
        float a[2][8];
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const int u = 2 * up + w;
          const f32x2 nx = pack2(sx[u], sx[u]), ny = pack2(sy[u], sy[u]), nz = pack2(sz[u], sz[u]);
          unpack2(sqd(pack2(X0.x, X0.y), pack2(Y0.x, Y0.y), pack2(Z0.x, Z0.y), nx, ny, nz), a[w][0], a[w][1]);
          unpack2(sqd(pack2(X0.z, X0.w), pack2(Y0.z, Y0.w), pack2(Z0.z, Z0.w), nx, ny, nz), a[w][2], a[w][3]);
          unpack2(sqd(pack2(X1.x, X1.y), pack2(Y1.x, Y1.y), pack2(Z1.x, Z1.y), nx, ny, nz), a[w][4], a[w][5]);
          unpack2(sqd(pack2(X1.z, X1.w), pack2(Y1.z, Y1.w), pack2(Z1.z, Z1.w), nx, ny, nz), a[w][6], a[w][7]);
          float mn = fminf(fminf(a[w][0], a[w][1]), best[u]);
          mn = fminf(fminf(a[w][2], a[w][3]), mn);
          mn = fminf(fminf(a[w][4], a[w][5]), mn);
          best[u] = fminf(fminf(a[w][6], a[w][7]), mn);
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
          for (int e = 0; e < 8; ++e) cm[e] = up ? fminf(fminf(a[0][e], a[1][e]), cm[e]) : fminf(a[0][e], a[1][e]);
        }
      }
      if (MODE == 1) {
#pragma unroll
        for (int e = 0; e < 8; ++e) colacc = fminf(colacc, cm[e]);
      }
      if (MODE == 2) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const unsigned int bits = __float_as_uint(cm[e]);
          const unsigned int mn = __reduce_min_sync(0xffffffffu, bits);
          colacc += __uint_as_float(mn) + (float)__ballot_sync(0xffffffffu, bits == mn);
        }
      }
    }
  float s = colacc;
  for (int u = 0; u < RPL; ++u) s += best[u];
  out[blockIdx.x * 128 + threadIdx.x] = s;
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
  float *out, *in; cudaMalloc(&out, sizeof(float) * p.multiProcessorCount * 8 * 128); cudaMalloc(&in, 4096);
  float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = (i * 37 % 101) * 0.01f; cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
  for (int per_sm : {1, 2, 4}) {
    const int blocks = p.multiProcessorCount * per_sm;
    const double pairs = (double)blocks * 128 * REPS * GROUPS * 8 * RPL;
    float t0 = time_ms([&] { k_sym<0><<<blocks, 128>>>(out, in); });
    float t1 = time_ms([&] { k_sym<1><<<blocks, 128>>>(out, in); });
    float t2 = time_ms([&] { k_sym<2><<<blocks, 128>>>(out, in); });
    float p0 = time_ms([&] { k_sym<0, true><<<blocks, 128>>>(out, in); });
    float p2 = time_ms([&] { k_sym<2, true><<<blocks, 128>>>(out, in); });
    float r3 = time_ms([&] { k_sym<3><<<blocks, 128>>>(out, in); });
    printf("   rows only, column operands in registers (no LDS in the loop): %.0f\n", pairs / (r3 * 1e-3) / 1e9);
    printf("   with register prefetch of the next column group: rows only %.0f   full %.0f\n", pairs / (p0 * 1e-3) / 1e9, pairs / (p2 * 1e-3) / 1e9);
    printf("CTAs/SM %d (%2d warps): rows only %.0f   + column minima %.0f   + REDUX/ballot %.0f  Gpairs/s\n", per_sm, per_sm * 4,
           pairs / (t0 * 1e-3) / 1e9, pairs / (t1 * 1e-3) / 1e9, pairs / (t2 * 1e-3) / 1e9);
  }
  return cudaDeviceSynchronize() != cudaSuccess;
}
