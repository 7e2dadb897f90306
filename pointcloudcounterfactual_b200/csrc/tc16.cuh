// fp16 tensor-core helpers shared by the xyz kernels that use tcgen05 as a candidate filter (chamfer_tc.cu, knn3_tc.cu):
// 16-column fp16 operand rows in the UMMA canonical K-major no-swizzle layout, kind::f16 MMA with fp16 accumulators,
// packed 16-bit read-out of tensor memory, and the two-piece fp16 split of an fp32 value.
#pragma once

#include <cuda_fp16.h>

#include "tc_ptx.cuh"

namespace pcc {

constexpr int T16_ROWB = 32;  // bytes per operand row: 16 fp16

// K-major, no swizzle: core matrices of 8 rows x 16 B; LBO = 128 B between the two core matrices of an 8-row group along
// K, SBO = 256 B between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc_k16(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3fff) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D = A = B = F16 (format code 0), both operands K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n) {
  return ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 64 accumulator columns of 16-bit data, two per register (register u = columns 2u | 2u+1 << 16)
__device__ __forceinline__ void tmem_ld64h_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.pack::16b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

// x = h1 + h2 + r, |r| <= 2^-22 |x| (fp16 pieces by truncation / rounding; the subtraction is exact)
__device__ __forceinline__ void f16x2(float v, unsigned short &h1, unsigned short &h2) {
  const __half a = __float2half_rz(v);
  const float r = v - __half2float(a);
  const __half b = __float2half_rn(r);
  h1 = __half_as_ushort(a);
  h2 = __half_as_ushort(b);
}
__device__ __forceinline__ unsigned short hneg2(unsigned short h) {  // -2 h, exact (|h| < 128)
  return __half_as_ushort(__float2half_rn(-2.f * __half2float(__ushort_as_half(h))));
}
__device__ __forceinline__ uint32_t pk2(unsigned short lo, unsigned short hi) { return (uint32_t)lo | ((uint32_t)hi << 16); }

// byte offset of 16-byte chunk kc (0 / 1) of operand row r inside an array of rows
__device__ __forceinline__ size_t nt_off(size_t r, int kc) { return (r >> 3) * 256 + (size_t)kc * 128 + (r & 7) * 16; }

}  // namespace pcc
