// Batched fp32 GEMM on the tcgen05 tensor cores with fp32-level accuracy (3xTF32) -- sm_100a.
//
// The three contractions of the fused EdgeConv layer (src/module/layers.py:159-203 applied to the POINTS, edgeconv.py):
//     forward            uv[b] (N x 2Cout) = x[b]^T (N x C)        . ws^T (C x 2Cout)
//     input gradient     gx[b]^T (N x C)   = guv[b] (N x 2Cout)    . ws (2Cout x C)            (stored channels-first)
//     weight gradient    gw[b] (2Cout x C) = guv[b]^T (2Cout x N)  . x[b]^T (N x C)            (summed over b afterwards)
// were torch.bmm calls (cuBLAS fp32 SIMT, a third of the Cout = 256 layer).  Every one of them has an operand that is
// contiguous along the OUTPUT dimension rather than along K, and fp32 accuracy needs the operands split: a = hi + lo with
// hi = the 19 bits kind::tf32 reads, lo = a - hi (exact), D = hi.hi + hi.lo + lo.hi (the dropped lo.lo term is 2^-22
// relative).  Both are done by the CTA's loader warps on the way into shared memory: they read a 128 x 32 tile of each
// operand from global memory through arbitrary (row, k) strides -- coalesced along whichever dimension is contiguous --
// and write hi and lo tiles in the UMMA canonical K-major no-swizzle layout (core matrices of 8 rows x 16 B).  No TMA
// descriptor, no transposed copy of x or of the gradient in global memory.
//   warps 0-7   loaders: global -> (hi, lo) shared-memory tiles of A and B, 3 stages of K = 32, next chunk prefetched
//   warp  8     tcgen05.mma kind::tf32, M = 128, N = 128, K = 8: twelve per stage (4 K steps x {hi.hi, hi.lo, lo.hi}) into
//               four accumulators (the tensor core truncates when it adds into fp32: fewer updates per accumulator)
//   warps 9-12  epilogue: tcgen05.ld of the 128 x 128 fp32 accumulators, their sum, stores through the strides of D
#include "tc_ptx.cuh"

namespace pcc {

constexpr int GT_M = 128, GT_N = 128;  // CTA tile
constexpr int GT_STAGES = 3;
constexpr int GT_LOADERS = 256;        // 8 loader warps: two threads per tile row, half a K chunk each
constexpr int GT_THREADS = GT_LOADERS + 5 * 32;

// Two shapes of the same kernel: KC = 16 with ONE hi.hi accumulator (short reductions -- the forward and input-gradient
// products, K = C or 2 Cout <= 512: 96 KiB of shared memory and 256 TMEM columns, two CTAs per SM), KC = 32 with THREE
// (the weight gradient, K = points per cloud).
template <int KC>
struct GtSmem {
  unsigned char t[GT_STAGES][4][128 * KC * 4];  // per stage: A hi, A lo, B hi, B lo
  uint64_t full[GT_STAGES], empty[GT_STAGES], tfull;
  uint32_t tmem_base;
};

// K-major, no swizzle, 32-bit elements: core matrix = 8 rows x 16 B (4 values); LBO = 128 B between the core matrices of an
// 8-row group along K, SBO = KC / 4 * 128 B between 8-row groups
template <int KC>
__device__ __forceinline__ uint64_t umma_desc_kf(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3fff) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)((KC / 4 * 128) >> 4) << 32) |
         (1ull << 46);
}

// A loader thread holds KH = KC / 2 values of a 128 x KC tile (element (r, kk) = src[r * sr + kk * sk] for r < rows, kk < kc,
// zero elsewhere) under one of two mappings, chosen per operand so that a warp's load instruction is coalesced along the
// dimension that is contiguous in memory:
//   row-contiguous (sr == 1)  thread t: row t % 128, the KH values kk = (t / 128) KH + u           (lanes = consecutive rows)
//   K-contiguous   (sk == 1)  thread t: KH / 4 groups of four kk of rows (t + 256 e) / (KC / 4)    (lanes = consecutive kk)
// A 4-byte load per lane across 32 rows of a K-contiguous operand would pull a 32-byte sector per lane: 8x the traffic.
template <int KC>
__device__ __forceinline__ void gt_fetch(const float *__restrict__ src, long long sr, long long sk, int rows, int kc, bool kmaj,
                                         float (&v)[KC / 2]) {
  constexpr int KH = KC / 2, Q = KC / 4;
  if (kmaj) {
#pragma unroll
    for (int e = 0; e < KH / 4; ++e) {
      const int idx = threadIdx.x + e * 256, r = idx / Q, k4 = (idx % Q) * 4;
      const float *p = src + (long long)r * sr + k4;
      if (r < rows && k4 + 3 < kc && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        const float4 q = *reinterpret_cast<const float4 *>(p);
        v[e * 4 + 0] = q.x;
        v[e * 4 + 1] = q.y;
        v[e * 4 + 2] = q.z;
        v[e * 4 + 3] = q.w;
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) v[e * 4 + u] = (r < rows && k4 + u < kc) ? p[u] : 0.f;
      }
    }
  } else {
    const int r = threadIdx.x & 127, kh = (threadIdx.x >> 7) * KH;
    const bool live = r < rows;
#pragma unroll
    for (int u = 0; u < KH; ++u) v[u] = (live && kh + u < kc) ? src[(long long)r * sr + (long long)(kh + u) * sk] : 0.f;
  }
}
// ... split into hi (the 19 bits kind::tf32 reads) and lo = v - hi (exact) and stored into the canonical layout
template <int KC>
__device__ __forceinline__ void gt_store(const float (&v)[KC / 2], unsigned char *hi, unsigned char *lo, bool kmaj) {
  constexpr int KH = KC / 2, Q = KC / 4;
#pragma unroll
  for (int e = 0; e < KH / 4; ++e) {
    int r, k4;
    if (kmaj) {
      const int idx = threadIdx.x + e * 256;
      r = idx / Q;
      k4 = idx % Q;
    } else {
      r = threadIdx.x & 127;
      k4 = (threadIdx.x >> 7) * (KH / 4) + e;
    }
    const size_t off = (size_t)(r >> 3) * (Q * 128) + (size_t)(r & 7) * 16 + (size_t)k4 * 128;
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v[e * 4 + 0]) & 0xffffe000u);
    h.y = __uint_as_float(__float_as_uint(v[e * 4 + 1]) & 0xffffe000u);
    h.z = __uint_as_float(__float_as_uint(v[e * 4 + 2]) & 0xffffe000u);
    h.w = __uint_as_float(__float_as_uint(v[e * 4 + 3]) & 0xffffe000u);
    l = make_float4(v[e * 4 + 0] - h.x, v[e * 4 + 1] - h.y, v[e * 4 + 2] - h.z, v[e * 4 + 3] - h.w);
    *reinterpret_cast<float4 *>(hi + off) = h;
    *reinterpret_cast<float4 *>(lo + off) = l;
  }
}

// grid (ceil(m/128), ceil(n/128), batch * ksplit)
template <int KC, int NMAIN>
__global__ void __launch_bounds__(GT_THREADS, KC == 16 ? 2 : 1)
gemm_tf32x3_kernel(int m, int n, int k_total, int ksplit, const float *__restrict__ A, long long sAb, long long sAm, long long sAk,
                   const float *__restrict__ B, long long sBb, long long sBn, long long sBk, float *__restrict__ D,
                   long long sDb, long long sDm, long long sDn) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  GtSmem<KC> &S = *reinterpret_cast<GtSmem<KC> *>(smem_raw);
  constexpr int GT_KC = KC, KH = KC / 2, TCOLS = (NMAIN + 1) * GT_N <= 256 ? 256 : 512;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * GT_M, n0 = blockIdx.y * GT_N;
  const size_t bz = blockIdx.z / ksplit;
  const int kpart = blockIdx.z % ksplit;
  // K range of this CTA: a multiple of the chunk size per part; slice blockIdx.z of D receives its partial product
  const int kper = ((k_total + ksplit - 1) / ksplit + GT_KC - 1) / GT_KC * GT_KC;
  const int kbeg = min(k_total, kpart * kper);
  const int k = min(k_total, kbeg + kper) - kbeg;
  A += bz * sAb + (long long)m0 * sAm + (long long)kbeg * sAk;
  B += bz * sBb + (long long)n0 * sBn + (long long)kbeg * sBk;
  D += (size_t)blockIdx.z * sDb;
  const int nchunk = (k + GT_KC - 1) / GT_KC;

  if (threadIdx.x == 0) {
    for (int s = 0; s < GT_STAGES; ++s) {
      mbar_init(&S.full[s], GT_LOADERS);
      mbar_init(&S.empty[s], 1);
    }
    mbar_init(&S.tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == GT_LOADERS / 32) tmem_alloc(&S.tmem_base, TCOLS);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = S.tmem_base;

  if (warp < GT_LOADERS / 32) {
    // ===== loaders: every thread owns KC / 2 values of both tiles; the global loads of chunk c+1 are in flight while chunk c
    //       is split and stored =====
    const bool ka = sAk == 1 && sAm != 1, kb = sBk == 1 && sBn != 1;  // K-contiguous operands take the K-major mapping
    float va[KH], vb[KH];
    gt_fetch<KC>(A, sAm, sAk, m - m0, min(GT_KC, k), ka, va);
    gt_fetch<KC>(B, sBn, sBk, n - n0, min(GT_KC, k), kb, vb);
    for (int c = 0; c < nchunk; ++c) {
      const int s = c % GT_STAGES, par = (c / GT_STAGES) & 1;
      mbar_wait(&S.empty[s], par ^ 1);
      gt_store<KC>(va, S.t[s][0], S.t[s][1], ka);
      gt_store<KC>(vb, S.t[s][2], S.t[s][3], kb);
      if (c + 1 < nchunk) {
        const int k1 = (c + 1) * GT_KC, kc1 = min(GT_KC, k - k1);
        gt_fetch<KC>(A + (long long)k1 * sAk, sAm, sAk, m - m0, kc1, ka, va);
        gt_fetch<KC>(B + (long long)k1 * sBk, sBn, sBk, n - n0, kc1, kb, vb);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
      mbar_arrive(&S.full[s]);
    }
  } else if (warp == GT_LOADERS / 32) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t IDESC = umma_idesc_tf32(GT_M, GT_N);
      // The tensor core adds into its fp32 accumulator by TRUNCATION: the error grows linearly with the number of updates
      // (measured: 1.6e-6 relative per 100 K steps).  Four accumulators (512 TMEM columns) keep it at the level of an fp32
      // SIMT GEMM: the hi.hi products rotate over three of them, the two small cross terms share the fourth (their sum
      // is 2^-11 of the result, so its truncation does not matter); the epilogue adds the four in fp32.
      int step = 0;
      for (int c = 0; c < nchunk; ++c) {
        const int s = c % GT_STAGES, par = (c / GT_STAGES) & 1;
        mbar_wait(&S.full[s], par);
        fence_after();
        const int kc = min(GT_KC, k - c * GT_KC);
        const int ksteps = (kc + 7) >> 3;
        const uint32_t ah = smem_u32(S.t[s][0]), al = smem_u32(S.t[s][1]), bh = smem_u32(S.t[s][2]), bl = smem_u32(S.t[s][3]);
        for (int ks = 0; ks < ksteps; ++ks, ++step) {  // K = 8 values = two core matrices = 256 B further along K
          const uint32_t o = ks * 256;
          mma_tf32(tmem_base + NMAIN * GT_N, umma_desc_kf<KC>(al + o), umma_desc_kf<KC>(bh + o), IDESC, step ? 1u : 0u);
          mma_tf32(tmem_base + NMAIN * GT_N, umma_desc_kf<KC>(ah + o), umma_desc_kf<KC>(bl + o), IDESC, 1u);
          mma_tf32(tmem_base + (uint32_t)((step % NMAIN) * GT_N), umma_desc_kf<KC>(ah + o), umma_desc_kf<KC>(bh + o), IDESC,
                   step >= NMAIN ? 1u : 0u);
        }
        mma_commit(&S.empty[s]);
      }
      mma_commit(&S.tfull);
    }
  } else {
    // ===== epilogue: thread = accumulator row =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int gm = m0 + row;
    mbar_wait(&S.tfull, 0);
    fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const bool vec = sDn == 1 && (sDm & 3) == 0 && (sDb & 3) == 0 && ((reinterpret_cast<uintptr_t>(D) & 15) == 0) && ((n0 & 3) == 0);
    const int nsteps = (k + 7) >> 3;  // chunks are multiples of 8 except the last: total K steps
    const int nacc = nsteps < NMAIN ? nsteps : NMAIN;
#pragma unroll 1
    for (int ch = 0; ch < GT_N / 32; ++ch) {
      float v[32], w[32];
      tmem_ld32(taddr + (uint32_t)(NMAIN * GT_N + ch * 32), v);  // cross terms first, then the hi.hi accumulators that were used
      if (nsteps == 0) {  // an empty K part (ksplit does not divide K): its slice of D is zero
#pragma unroll
        for (int u = 0; u < 32; ++u) v[u] = 0.f;
      }
#pragma unroll 1
      for (int a2 = 0; a2 < nacc; ++a2) {
        tmem_ld32(taddr + (uint32_t)(a2 * GT_N + ch * 32), w);
#pragma unroll
        for (int u = 0; u < 32; ++u) v[u] += w[u];
      }
      const int nb = n0 + ch * 32;
      if (gm < m && nb < n) {  // (a shared-memory transpose for full-line stores of row-major D was measured: 5 % slower)
        float *drow = D + (long long)gm * sDm;
        if (vec && nb + 32 <= n) {
#pragma unroll
          for (int u = 0; u < 32; u += 4)
            *reinterpret_cast<float4 *>(drow + nb + u) = make_float4(v[u], v[u + 1], v[u + 2], v[u + 3]);
        } else {
#pragma unroll
          for (int u = 0; u < 32; ++u)
            if (nb + u < n) drow[(long long)(nb + u) * sDn] = v[u];
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == GT_LOADERS / 32) tmem_dealloc(tmem_base, TCOLS);
}

}  // namespace pcc

using namespace pcc;

extern "C" __attribute__((visibility("default"))) int pcc_gemm_tf32x3(int batch, int ksplit, int m, int n, int k, const float *A, long long sAb,
                                                                       long long sAm, long long sAk, const float *B, long long sBb,
                                                                       long long sBn, long long sBk, float *D, long long sDb,
                                                                       long long sDm, long long sDn, pcc_stream_t stream) {
  if (batch < 0 || m < 0 || n < 0 || k <= 0 || ksplit < 1) return PCC_EINVAL;
  if (batch == 0 || m == 0 || n == 0) return PCC_OK;
  if ((long long)batch * ksplit > 65535 || (n + GT_N - 1) / GT_N > 65535) return PCC_ENOTSUP;
  const dim3 grid((m + GT_M - 1) / GT_M, (n + GT_N - 1) / GT_N, batch * ksplit);
  const int kpart = (k + ksplit - 1) / ksplit;
  if (kpart <= 512) {  // short reduction: the light shape, two CTAs per SM
    static size_t attr[64];
    const size_t smem = sizeof(GtSmem<16>) + 1024;
    if (cudaError_t e = smem_optin(gemm_tf32x3_kernel<16, 1>, smem, attr); e != cudaSuccess) return (int)e;
    gemm_tf32x3_kernel<16, 1><<<grid, GT_THREADS, smem, (cudaStream_t)stream>>>(m, n, k, ksplit, A, sAb, sAm, sAk, B, sBb, sBn,
                                                                              sBk, D, sDb, sDm, sDn);
  } else {
    static size_t attr[64];
    const size_t smem = sizeof(GtSmem<32>) + 1024;
    if (cudaError_t e = smem_optin(gemm_tf32x3_kernel<32, 3>, smem, attr); e != cudaSuccess) return (int)e;
    gemm_tf32x3_kernel<32, 3><<<grid, GT_THREADS, smem, (cudaStream_t)stream>>>(m, n, k, ksplit, A, sAb, sAm, sAk, B, sBb, sBn,
                                                                              sBk, D, sDb, sDm, sDn);
  }
  return finish_launch(1);
}
