// Throughput of the min instructions a score epilogue can use (per SM and clock): FMNMX3 (fp32, three inputs), HMNMX2
// (packed fp16), VIMNMX3.S16x2 / VIMNMX.S16x2 (packed 16-bit integers, DPX).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/minmax_probe tools/minmax_probe.cu
#include <cuda_fp16.h>
#include <stdio.h>

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(unsigned *out, unsigned seed, int iters, unsigned long long *cyc) {
  unsigned a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + 1) + i * 0x01010101u;
  unsigned b = seed ^ 0x12345678u, c = seed + 77u;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) a[i] = __float_as_uint(fminf(fminf(__uint_as_float(a[i]), __uint_as_float(b)), __uint_as_float(c)));
      if (OP == 1) {
        __half2 h = __hmin2(*reinterpret_cast<__half2 *>(&a[i]), *reinterpret_cast<__half2 *>(&b));
        a[i] = *reinterpret_cast<unsigned *>(&h);
      }
      if (OP == 2) a[i] = __vimin3_s16x2(a[i], b, c);
      if (OP == 3) a[i] = __vmins2(a[i], b);
      if (OP == 4) a[i] = min(a[i], b);
      if (OP == 5) {
        __half2 h = __hmin2(__hmin2(*reinterpret_cast<__half2 *>(&a[i]), *reinterpret_cast<__half2 *>(&b)), *reinterpret_cast<__half2 *>(&c));
        a[i] = *reinterpret_cast<unsigned *>(&h);
      }
    }
    b += 0x00010001u;
    c ^= b;
  }
  __syncthreads();
  const long long t1 = clock64();
  unsigned r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = (unsigned long long)(t1 - t0);
}

int main() {
  unsigned *out;
  unsigned long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 8);
  const int iters = 4096;
  const char *names[6] = {"FMNMX3 (fp32 min3)", "HMNMX2 (fp16x2 min)", "VIMNMX3.S16x2 (DPX min3)", "VIMNMX.S16x2 (min)", "IMNMX.U32 (min)", "VHMNMX (fp16x2 min3)"};
  for (int op = 0; op < 6; ++op) {
    for (int rep = 0; rep < 2; ++rep) {
      if (op == 0) k<0><<<148, 1024>>>(out, 3u, iters, cyc);
      if (op == 1) k<1><<<148, 1024>>>(out, 3u, iters, cyc);
      if (op == 2) k<2><<<148, 1024>>>(out, 3u, iters, cyc);
      if (op == 3) k<3><<<148, 1024>>>(out, 3u, iters, cyc);
      if (op == 4) k<4><<<148, 1024>>>(out, 3u, iters, cyc);
      if (op == 5) k<5><<<148, 1024>>>(out, 3u, iters, cyc);
      cudaDeviceSynchronize();
    }
    unsigned long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %.1f lane-instructions per clock and SM\n", names[op], (double)iters * 8 * 1024 / (double)h);
  }
  return 0;
}
