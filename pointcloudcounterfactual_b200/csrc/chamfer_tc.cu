// Chamfer nearest-neighbour search on the tcgen05 tensor cores (sm_100a) -- the production forward of pcc_nndistance /
// pcc_chamfer_reduce for clouds of 256 .. 65535 points.
//
// Replaces external/pytorch_structural_losses/src/nndistance.cu:2-128 (NmDistanceKernel, launched once per direction).
// The brute-force search evaluates B*N*M squared distances on the FP32 pipe (8 flop each).  Here the distances come out
// of the tensor cores as a CANDIDATE FILTER and only the handful of candidates per query is evaluated with the
// reference's arithmetic, so the results (distance bits and lowest-index tie rule) are identical to the SIMT kernels:
//
//   prep   nn_tc_prep_kernel     per cloud and side: translate by a sample mean of the second cloud, split every
//                                coordinate into three bf16 pieces x = b1 + b2 + b3 (exact: 3 x 8 mantissa bits) and write
//                                one 32-column bf16 operand row per point in the two roles it plays,
//                                  query role A_i = [-2a1 -2a1 -2a2 -2a2 -2a1 -2a3]_{x,y,z}  1 1 1   nA1 nA2 nA3  0..
//                                  key role   B_j = [  b1   b2   b1   b2   b3   b1]_{x,y,z} nB1 nB2 nB3  1 1 1   0..
//                                so that A_i . B_j = |a|^2 + |b|^2 - 2 a.b up to the dropped products (2^-24 relative) and
//                                the fp32 accumulation of the tensor core.  Rows are stored in the UMMA canonical
//                                K-major no-swizzle order (8-row groups of 4 core matrices), i.e. a tile of rows is one
//                                contiguous block that a plain cp.async.bulk brings into shared memory.
//   main   nn_tc_kernel          persistent CTA = one cloud, one direction, a contiguous range of 128-query tiles.  ALL keys
//                                of the cloud are brought into shared memory once (cp.async.bulk + one mbarrier per key
//                                tile) and stay there: operand traffic from L2 is one pass per CTA instead of one pass
//                                per query tile (a streaming version of this kernel was L2-bound at 4.8 TB/s, 82 us).
//                                warp 0 loads, warp 1 issues tcgen05.mma kind::f16 (bf16, M=128, N=256) into a
//                                double-buffered TMEM accumulator (512 columns), sixteen epilogue warps in four sets own
//                                one query per thread and two 32-key chunks of every key tile: tcgen05.ld, the minimum
//                                of the 32 scores (the only per-score work: one 3-input FMNMX per two scores) and a short
//                                list of the chunks whose minimum is within the error band of the running row minimum.
//                                At the end of a query tile the sets exchange their row minima and keep, per query, the
//                                chunks still within the band of the FINAL minimum (two slots per set); after the last
//                                tile every recorded chunk is evaluated exactly
//                                (d = fma(dz,dz,fma(dx,dx,dy*dy)), lexicographic minimum of (distance, key index)).
//   The candidate set provably contains the exact nearest neighbour and every key tied with it: the score of pair
//   (i,j) differs from the exact distance by at most eps_ij = NT_CEPS (|a_i|^2 + |b_j|^2) (measured: tools/nn_tc_probe),
//   and a chunk is kept when its minimum is <= row minimum + 2 max_j eps_ij.  Queries whose record overflows (massive ties),
//   whose scores are not finite (NaN / inf coordinates) or whose first key is NaN (the reference lets a NaN at k = 0
//   stick, nndistance.cu:26) are redone by an exact warp-cooperative scan with the reference's NaN semantics.
#include "tc_ptx.cuh"

namespace pcc {

constexpr int NT_M = 128;          // queries per tile = TMEM lanes
constexpr int NT_N = 256;          // keys per tile = accumulator columns per TMEM stage
constexpr int NT_ROWB = 64;        // bytes per operand row: 32 bf16
constexpr int NT_SETS = 4;         // epilogue warp sets (4 warps each: one per TMEM lane quarter)
constexpr int NT_CPS = NT_N / 32 / NT_SETS;  // 32-key chunks per tile and set
constexpr int NT_CAP = 8;          // running candidate chunks per query and set
constexpr int NT_REC = 2;          // recorded candidate chunks per query and set
constexpr int NT_THREADS = 64 + 128 * NT_SETS;  // warp 0 producer, warp 1 MMA, 16 epilogue warps
constexpr int NT_EPI = 128 * NT_SETS;
constexpr int NT_PREP_PARTS = 4;   // CTAs per cloud and side in the operand preparation
constexpr int NT_TILE_BYTES = NT_N * NT_ROWB;
constexpr int NT_MAX_KEYS = 2048;  // resident keys (8 tiles = 128 KiB) ...
constexpr int NT_MAX_Q = 2048;     // ... and queries per CTA (records)
// |score - exact distance| <= NT_CEPS * (|a|^2 + |b|^2): bf16x3 products dropped (3 * 2^-24 |a||b|), fp32 accumulation of
// 24 non-zero terms in the tensor core, rounding of the translated coordinates and of the canonical fma chain.
// tools/nn_tc_probe measures the actual maximum: 1.2e-6 over the S1 / S2 / S3 families; 2^-17 leaves a factor of 6.
constexpr float NT_CEPS = 7.6293945e-6f;

struct NtCtl {
  uint64_t kfull[NT_MAX_KEYS / NT_N], afull[2], aempty[2], tfull[2], tempty[2];
  uint32_t tmem_base;
};
struct NtSmem {
  unsigned char keys[NT_MAX_KEYS * NT_ROWB];
  unsigned char a[2][NT_M * NT_ROWB];
  float lst_v[NT_SETS][NT_CAP][NT_M];
  unsigned short lst_c[NT_SETS][NT_CAP][NT_M];
  float mrow[NT_SETS][NT_M];           // row minimum of each set (exchanged at the end of a query tile)
  unsigned char flag[NT_SETS][NT_M];   // set asks for the exact scan of this query
  unsigned short rec_c[NT_MAX_Q][NT_SETS][NT_REC];  // per query of the CTA: candidate chunks of each set
  unsigned char rec_n[NT_MAX_Q][NT_SETS];           // ... how many; 255 = exact scan
  NtCtl ctl;
};

// K-major, no swizzle: core matrices of 8 rows x 16 B, LBO = 128 B between the core matrices of one 8-row group along K,
// SBO = 512 B between 8-row groups (4 core matrices = 32 bf16 per row).
__device__ __forceinline__ uint64_t umma_desc_k32(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3fff) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D = F32, A = B = BF16, both K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- operand preparation ---------------------------------------------------------------------------------------
__device__ __forceinline__ void bf16x3(float v, uint32_t &p1, uint32_t &p2, uint32_t &p3) {  // pieces as fp32 bit patterns
  p1 = __float_as_uint(v) & 0xffff0000u;
  const float r1 = v - __uint_as_float(p1);  // exact
  p2 = __float_as_uint(r1) & 0xffff0000u;
  const float r2 = r1 - __uint_as_float(p2);  // exact, at most 8 significant bits
  p3 = __float_as_uint(r2) & 0xffff0000u;
}
__device__ __forceinline__ uint32_t pk(uint32_t lo_bits, uint32_t hi_bits) {  // two bf16 (top halves) -> one word
  return (lo_bits >> 16) | (hi_bits & 0xffff0000u);
}
__device__ __forceinline__ uint32_t neg2(uint32_t bits) {  // -2 x (exact: exponent + 1, sign flipped); zero stays zero
  const float v = __uint_as_float(bits);
  return __float_as_uint(-2.f * v);
}

// byte offset of 16-byte chunk kc of operand row r inside an array of rows
__device__ __forceinline__ size_t nt_off(size_t r, int kc) { return (r >> 3) * 512 + (size_t)kc * 128 + (r & 7) * 16; }

// grid (b, 2, NT_PREP_PARTS): side 0 = xyz1 (n points), side 1 = xyz2 (m points); every CTA takes the mean of 32 sample
// points of cloud 2 as the common translation (any translation is valid, a central one keeps the norms -- and with them
// the error band -- small) and converts its share of the rows.  ops* = [cloud][npad rows][64 B]; key4 = the original
// coordinates as float4 (the exact resolution reads them with one 16-byte load per key); nmax[side][b] must be zeroed.
__global__ void __launch_bounds__(256)
nn_tc_prep_kernel(int n, const float *__restrict__ xyz1, int m, const float *__restrict__ xyz2, int npad1, int npad2,
                  unsigned char *__restrict__ opsA1, unsigned char *__restrict__ opsB1, unsigned char *__restrict__ opsA2,
                  unsigned char *__restrict__ opsB2, float4 *__restrict__ key41, float4 *__restrict__ key42,
                  unsigned int *__restrict__ nmax /* [2][b] float bits */) {
  __shared__ float smax[8];
  __shared__ float sctr[3];
  const size_t cloud = blockIdx.x;
  const int side = blockIdx.y;
  const float *p1 = xyz1 + cloud * (size_t)n * 3, *p2 = xyz2 + cloud * (size_t)m * 3;
  const float INF = __int_as_float(0x7f800000);
  if (threadIdx.x < 32) {  // the same 32 points and the same arithmetic in every CTA of the cloud: identical centres
    const float *sp = p2 + (size_t)((long long)threadIdx.x * m / 32) * 3;
    const float x = sp[0], y = sp[1], z = sp[2];
    const bool ok = fabsf(x) < INF && fabsf(y) < INF && fabsf(z) < INF;  // NaN / inf points do not move the centre
    const float c = warp_sum(ok ? 1.f : 0.f);
    const float sx = warp_sum(ok ? x : 0.f), sy = warp_sum(ok ? y : 0.f), sz = warp_sum(ok ? z : 0.f);
    if (threadIdx.x == 0) {
      const float inv = c > 0.f ? 1.f / c : 0.f;
      sctr[0] = sx * inv;
      sctr[1] = sy * inv;
      sctr[2] = sz * inv;
    }
  }
  __syncthreads();
  const float ctr[3] = {sctr[0], sctr[1], sctr[2]};
  const int cnt = side ? m : n, npad = side ? npad2 : npad1;
  const float *p = side ? p2 : p1;
  unsigned char *oa = (side ? opsA2 : opsA1) + cloud * (size_t)npad * NT_ROWB;
  unsigned char *ob = (side ? opsB2 : opsB1) + cloud * (size_t)npad * NT_ROWB;
  float4 *k4 = (side ? key42 : key41) + cloud * (size_t)npad;
  float nm = 0.f;
  for (int r = blockIdx.z * 256 + threadIdx.x; r < npad; r += 256 * NT_PREP_PARTS) {
    uint4 a0, a1, a2, a3, b0, b1, b2, b3;
    if (r < cnt) {
      const float ox = p[(size_t)r * 3], oy = p[(size_t)r * 3 + 1], oz = p[(size_t)r * 3 + 2];
      k4[r] = make_float4(ox, oy, oz, 0.f);
      const float x = ox - ctr[0], y = oy - ctr[1], z = oz - ctr[2];
      const float nn = fmaf(z, z, fmaf(y, y, x * x));
      if (nn < INF) nm = fmaxf(nm, nn);
      uint32_t x1, x2, x3, y1, y2, y3, z1, z2, z3, n1, n2, n3;
      bf16x3(x, x1, x2, x3);
      bf16x3(y, y1, y2, y3);
      bf16x3(z, z1, z2, z3);
      bf16x3(nn, n1, n2, n3);
      const uint32_t one = 0x3f800000u;
      // column order per coordinate: (a1 b1) (a1 b2) (a2 b1) (a2 b2) (a1 b3) (a3 b1); 18..20 = 1 x nB, 21..23 = nA x 1
      const uint32_t ax[6] = {neg2(x1), neg2(x1), neg2(x2), neg2(x2), neg2(x1), neg2(x3)};
      const uint32_t ay[6] = {neg2(y1), neg2(y1), neg2(y2), neg2(y2), neg2(y1), neg2(y3)};
      const uint32_t az[6] = {neg2(z1), neg2(z1), neg2(z2), neg2(z2), neg2(z1), neg2(z3)};
      const uint32_t bx[6] = {x1, x2, x1, x2, x3, x1}, by[6] = {y1, y2, y1, y2, y3, y1}, bz[6] = {z1, z2, z1, z2, z3, z1};
      a0 = make_uint4(pk(ax[0], ax[1]), pk(ax[2], ax[3]), pk(ax[4], ax[5]), pk(ay[0], ay[1]));
      a1 = make_uint4(pk(ay[2], ay[3]), pk(ay[4], ay[5]), pk(az[0], az[1]), pk(az[2], az[3]));
      a2 = make_uint4(pk(az[4], az[5]), pk(one, one), pk(one, n1), pk(n2, n3));
      a3 = make_uint4(0u, 0u, 0u, 0u);
      b0 = make_uint4(pk(bx[0], bx[1]), pk(bx[2], bx[3]), pk(bx[4], bx[5]), pk(by[0], by[1]));
      b1 = make_uint4(pk(by[2], by[3]), pk(by[4], by[5]), pk(bz[0], bz[1]), pk(bz[2], bz[3]));
      b2 = make_uint4(pk(bz[4], bz[5]), pk(n1, n2), pk(n3, one), pk(one, one));
      b3 = make_uint4(0u, 0u, 0u, 0u);
    } else {  // padding: as a key it scores 2^127 (never a candidate), as a query its row is never written
      k4[r] = make_float4(INF, INF, INF, 0.f);
      a0 = a1 = a2 = a3 = b0 = b1 = b3 = make_uint4(0u, 0u, 0u, 0u);
      b2 = make_uint4(0u, pk(0x7f000000u, 0u), 0u, 0u);  // nB1 = 2^127
    }
    *reinterpret_cast<uint4 *>(oa + nt_off(r, 0)) = a0;
    *reinterpret_cast<uint4 *>(oa + nt_off(r, 1)) = a1;
    *reinterpret_cast<uint4 *>(oa + nt_off(r, 2)) = a2;
    *reinterpret_cast<uint4 *>(oa + nt_off(r, 3)) = a3;
    *reinterpret_cast<uint4 *>(ob + nt_off(r, 0)) = b0;
    *reinterpret_cast<uint4 *>(ob + nt_off(r, 1)) = b1;
    *reinterpret_cast<uint4 *>(ob + nt_off(r, 2)) = b2;
    *reinterpret_cast<uint4 *>(ob + nt_off(r, 3)) = b3;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nm = fmaxf(nm, __shfl_xor_sync(0xffffffffu, nm, o));
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = nm;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = smax[0];
    for (int w = 1; w < 8; ++w) v = fmaxf(v, smax[w]);
    atomicMax(&nmax[(size_t)side * gridDim.x + cloud], __float_as_uint(v));  // v >= 0: unsigned order == float order
  }
}

// exact scan of all keys for one query by a whole warp, with the reference's semantics (nndistance.cu:2-124): ascending
// keys, strict '<', a NaN distance at key 0 sticks, later NaNs are skipped
__device__ __forceinline__ void nn_exact_warp(float qx, float qy, float qz, const float *__restrict__ keys, int nr, int lane,
                                              float &bd, int &bi) {
  const float INF = __int_as_float(0x7f800000);
  float d0 = sqdist1(qx, qy, qz, keys[0], keys[1], keys[2]);
  float best = INF;
  int besti = 0x7fffffff;
  for (int j = lane; j < nr; j += 32) {
    const float d = sqdist1(qx, qy, qz, keys[(size_t)j * 3], keys[(size_t)j * 3 + 1], keys[(size_t)j * 3 + 2]);
    if (d < best) {  // ascending j per lane: the first minimum of the lane is its lowest index
      best = d;
      besti = j;
    }
  }
  const unsigned int bits = __float_as_uint(best);  // best >= 0 or +inf: unsigned order == float order
  const unsigned int mnb = __reduce_min_sync(0xffffffffu, bits);
  const unsigned int cand = bits == mnb ? (unsigned int)besti : 0x7fffffffu;
  const unsigned int mni = __reduce_min_sync(0xffffffffu, cand);
  if (d0 != d0) {
    bd = d0;
    bi = 0;
  } else if (mni == 0x7fffffffu) {  // nothing below +inf
    bd = d0;
    bi = 0;
  } else {
    bd = __uint_as_float(mnb);
    bi = (int)mni;
  }
}

__device__ __forceinline__ void nt_bar_epilogue() { asm volatile("bar.sync 1, %0;" ::"n"(NT_EPI) : "memory"); }

// rare: the list of candidate chunks is full -> keep what is still within the band of the current minimum
__device__ __noinline__ int nt_compact(float *lv, unsigned short *lc, float lim) {
  int w2 = 0;
#pragma unroll 1
  for (int r = 0; r < NT_CAP; ++r) {
    const float vv = lv[r * NT_M];
    const unsigned short cc = lc[r * NT_M];
    if (vv <= lim) {
      lv[w2 * NT_M] = vv;
      lc[w2 * NT_M] = cc;
      ++w2;
    }
  }
  return w2;
}

// grid (splits, b, 2): blockIdx.z = direction (0: queries xyz1, keys xyz2; 1: swapped); blockIdx.x owns the query tiles
// [blockIdx.x * qt_per, ...).
__global__ void __launch_bounds__(NT_THREADS, 1)
nn_tc_kernel(int n, const float *__restrict__ xyz1, int m, const float *__restrict__ xyz2, int npad1, int npad2, int qt_per,
             const unsigned char *__restrict__ opsA1, const unsigned char *__restrict__ opsB1,
             const unsigned char *__restrict__ opsA2, const unsigned char *__restrict__ opsB2,
             const float4 *__restrict__ key41, const float4 *__restrict__ key42, const float *__restrict__ nmax,
             float *__restrict__ dist1, int *__restrict__ idx1, float *__restrict__ dist2, int *__restrict__ idx2,
             unsigned int *__restrict__ stats /* may be null: [0] exact rescans, [1] resolved chunks */
#ifdef NT_DEBUG_SCORES
             ,
             float *__restrict__ dbg /* [nq][npadr] raw scores of cloud 0, direction 0 */
#endif
) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  NtSmem &S = *reinterpret_cast<NtSmem *>(smem_raw);
  const int dir = blockIdx.z;
  const int nq = dir ? m : n, nr = dir ? n : m;
  const size_t cloud = blockIdx.y;
  const int npadq = dir ? npad2 : npad1, npadr = dir ? npad1 : npad2;
  const int qt_total = (nq + NT_M - 1) / NT_M;
  const int qt0 = blockIdx.x * qt_per, qt1 = min(qt_total, qt0 + qt_per);
  if (qt0 >= qt1) return;  // uniform per CTA
  const int nqt = qt1 - qt0;
  const unsigned char *opsq = (dir ? opsA2 : opsA1) + cloud * (size_t)npadq * NT_ROWB;
  const unsigned char *opsr = (dir ? opsB1 : opsB2) + cloud * (size_t)npadr * NT_ROWB;
  const float4 *q4 = (dir ? key42 : key41) + cloud * (size_t)npadq;
  const float4 *r4 = (dir ? key41 : key42) + cloud * (size_t)npadr;
  const float *rp = (dir ? xyz1 : xyz2) + cloud * (size_t)nr * 3;
  float *dout = (dir ? dist2 : dist1) + cloud * (size_t)nq;
  int *iout = (dir ? idx2 : idx1) + cloud * (size_t)nq;
  const int ntile = npadr / NT_N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  NtCtl *ctl = &S.ctl;

  if (threadIdx.x == 0) {
    for (int t = 0; t < NT_MAX_KEYS / NT_N; ++t) mbar_init(&ctl->kfull[t], 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->afull[s], 1);
      mbar_init(&ctl->aempty[s], 1);
      mbar_init(&ctl->tfull[s], 1);
      mbar_init(&ctl->tempty[s], 4 * NT_SETS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&ctl->tmem_base, 2 * NT_N);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
#ifdef NT_DEBUG_STAMPS
  const bool probe_cta = stats && blockIdx.x == 1 && blockIdx.y == 3 && blockIdx.z == 0;
  const long long t_start = clock64();
#define NT_STAMP(slot) if (probe_cta && lane == 0) stats[slot] = (unsigned int)(clock64() - t_start)
#else
#define NT_STAMP(slot)
#endif

  if (warp == 0) {
    // ===== producer: first query tile, ALL key tiles (they stay), then the remaining query tiles through two slots =====
    if (lane == 0) {
      mbar_expect_tx(&ctl->afull[0], NT_M * NT_ROWB);
      bulk_load_1d(S.a[0], opsq + (size_t)qt0 * NT_M * NT_ROWB, NT_M * NT_ROWB, &ctl->afull[0]);
      for (int t = 0; t < ntile; ++t) {
        mbar_expect_tx(&ctl->kfull[t], NT_TILE_BYTES);
        bulk_load_1d(S.keys + (size_t)t * NT_TILE_BYTES, opsr + (size_t)t * NT_TILE_BYTES, NT_TILE_BYTES, &ctl->kfull[t]);
      }
      NT_STAMP(2);
      for (int i = 1; i < nqt; ++i) {
        const int sa = i & 1;
        mbar_wait(&ctl->aempty[sa], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&ctl->afull[sa], NT_M * NT_ROWB);
        bulk_load_1d(S.a[sa], opsq + (size_t)(qt0 + i) * NT_M * NT_ROWB, NT_M * NT_ROWB, &ctl->afull[sa]);
      }
      NT_STAMP(3);
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t IDESC = umma_idesc_bf16(NT_M, NT_N);
      int it = 0;
      for (int i = 0; i < nqt; ++i) {
        const int sa = i & 1;
        mbar_wait(&ctl->afull[sa], (i >> 1) & 1);
        const uint32_t a_addr = smem_u32(S.a[sa]);
        for (int t = 0; t < ntile; ++t, ++it) {
          const int acc = it & 1;
          if (i == 0) mbar_wait(&ctl->kfull[t], 0);
          mbar_wait(&ctl->tempty[acc], ((it >> 1) & 1) ^ 1);  // every epilogue warp has this accumulator in registers
          fence_after();
          const uint32_t b_addr = smem_u32(S.keys + (size_t)t * NT_TILE_BYTES);
          const uint32_t tmem_d = tmem_base + (uint32_t)(acc * NT_N);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)  // K = 16 bf16 per instruction = two core matrices = 256 B further along K
            mma_bf16(tmem_d, umma_desc_k32(a_addr + ks * 256), umma_desc_k32(b_addr + ks * 256), IDESC, ks ? 1u : 0u);
          mma_commit(&ctl->tfull[acc]);
          if (it == 0) NT_STAMP(4);
        }
        mma_commit(&ctl->aempty[sa]);  // the query slot may be refilled once these MMAs have read it
      }
      NT_STAMP(5);
    }
  } else {
    // ===== epilogue: one query per thread and set; the four sets split the eight 32-key chunks of every key tile =====
    const int set = (warp - 2) >> 2;
    const int quarter = warp & 3;  // this warp may touch TMEM lanes 32*quarter .. +31
    const int e = quarter * 32 + lane;
    const float INF = __int_as_float(0x7f800000);
    // band = 2 max_j eps_ij, with |a_i|^2 bounded by the maximum over the query side
    const float band = 2.f * NT_CEPS * (nmax[(size_t)dir * gridDim.y + cloud] + nmax[(size_t)(1 - dir) * gridDim.y + cloud]);
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(set * NT_CPS * 32);
    float *lv = &S.lst_v[set][0][e];
    unsigned short *lc = &S.lst_c[set][0][e];
    const uint32_t tfull_a = smem_u32(&ctl->tfull[0]), tempty_a = smem_u32(&ctl->tempty[0]);
    int it = 0;
    for (int i = 0; i < nqt; ++i) {
      float mrow = INF;
      int cnt = 0;
      bool slow = false;
      for (int t = 0; t < ntile; ++t, ++it) {
        const int acc = it & 1;
        mbar_wait_a(tfull_a + acc * 8, (it >> 1) & 1);
        fence_after();
        if (warp == 2 && it == 0) NT_STAMP(9);
        uint32_t w[NT_CPS][32];
#pragma unroll
        for (int c2 = 0; c2 < NT_CPS; ++c2) tmem_ld32_issue(tlane + (uint32_t)(acc * NT_N + c2 * 32), w[c2]);
#pragma unroll
        for (int c2 = 0; c2 < NT_CPS; ++c2) {
          if (c2 == 0) {
#pragma unroll
            for (int c3 = 0; c3 < NT_CPS; ++c3) tmem_ld_wait_dep(w[c3]);  // one wait covers every load issued above
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(tempty_a + acc * 8);  // the accumulator is in registers: the next MMA may overwrite it
          }
          const int ch = t * (NT_N / 32) + set * NT_CPS + c2;  // chunk = keys 32 ch .. 32 ch + 31
#ifdef NT_DEBUG_SCORES
          if (dbg && cloud == 0 && dir == 0 && (qt0 + i) * NT_M + e < nq)
            for (int u = 0; u < 32; ++u) dbg[(size_t)((qt0 + i) * NT_M + e) * npadr + ch * 32 + u] = __uint_as_float(w[c2][u]);
#endif
          if (ch == 0) {
            const float s0 = __uint_as_float(w[c2][0]);
            slow = slow || (s0 != s0);  // a NaN distance to key 0 sticks in the reference
          }
          float c0 = fminf(__uint_as_float(w[c2][0]), __uint_as_float(w[c2][1]));
          float c1 = fminf(__uint_as_float(w[c2][2]), __uint_as_float(w[c2][3]));
#pragma unroll
          for (int u = 4; u < 32; u += 4) {
            c0 = fminf(fminf(__uint_as_float(w[c2][u]), __uint_as_float(w[c2][u + 1])), c0);
            c1 = fminf(fminf(__uint_as_float(w[c2][u + 2]), __uint_as_float(w[c2][u + 3])), c1);
          }
          const float cm = fminf(c0, c1);
          if (cm <= mrow + band) {
            if (cnt == NT_CAP) cnt = nt_compact(lv, lc, fminf(mrow, cm) + band);
            if (cnt < NT_CAP) {
              lv[cnt * NT_M] = cm;
              lc[cnt * NT_M] = (unsigned short)ch;
              ++cnt;
            } else {
              slow = true;  // more than NT_CAP chunks tie within the band
            }
          }
          mrow = fminf(mrow, cm);
        }
      }
      // ---- end of the query tile: final row minimum over the sets, record the chunks still within its band ----
      S.mrow[set][e] = mrow;
      S.flag[set][e] = slow ? 1 : 0;
      nt_bar_epilogue();
      float mfin = S.mrow[0][e];
      unsigned int fl = S.flag[0][e];
#pragma unroll
      for (int s2 = 1; s2 < NT_SETS; ++s2) {
        mfin = fminf(mfin, S.mrow[s2][e]);
        fl |= S.flag[s2][e];
      }
      const float lim = mfin + band;
      const int ql = i * NT_M + e;
      int nv = 0;
      for (int r = 0; r < cnt; ++r)
        if (lv[r * NT_M] <= lim) {
          if (nv < NT_REC) S.rec_c[ql][set][nv] = lc[r * NT_M];
          ++nv;
        }
      S.rec_n[ql][set] = (fl || !(mfin < INF) || nv > NT_REC) ? 255 : (unsigned char)nv;
      nt_bar_epilogue();  // the exchange buffers are free for the next query tile
    }
    if (warp == 2) NT_STAMP(6);

    // ---- exact resolution.  Every warp takes 32 queries at a time; their (query, chunk) pairs are evaluated by the whole
    //      warp, one pair per step: lane l computes the distance to key 32 chunk + l (ONE coalesced 512-byte read of the
    //      float4 copies -- a thread-per-query scan touches 32 cache lines per load instruction and was 10x slower), two
    //      REDUX give (minimum distance, lowest index attaining it), the owner lane keeps the lexicographic minimum ----
    const int ew = warp - 2;  // 0 .. 15
    for (int base = ew * 32; base < nqt * NT_M; base += (NT_EPI / 32) * 32) {
      const int ql = base + lane;
      const int q = qt0 * NT_M + ql;
      const bool live = q < nq;  // base < nqt * NT_M is warp-uniform and a multiple of 32: ql is inside the records
      float bd = INF;
      int bi = 0;
      bool any = false, slow = false;
      float4 qc = make_float4(0.f, 0.f, 0.f, 0.f);
      uint32_t cn = 0;  // counts of the four sets, one byte each
      if (live) {
        qc = q4[q];
        cn = *reinterpret_cast<const uint32_t *>(&S.rec_n[ql][0]);
        slow = ((cn & 0xffu) == 255u) || (((cn >> 8) & 0xffu) == 255u) || (((cn >> 16) & 0xffu) == 255u) || ((cn >> 24) == 255u);
        if (slow) cn = 0;
      }
      unsigned int nres = 0;
#pragma unroll 1
      for (int s2 = 0; s2 < NT_SETS; ++s2) {
#pragma unroll 1
        for (int v = 0; v < NT_REC; ++v) {
          unsigned int have = __ballot_sync(0xffffffffu, (int)((cn >> (8 * s2)) & 0xffu) > v);
          nres += __popc(have);
          while (have) {
            const int src = __ffs(have) - 1;
            have &= have - 1;
            const float sx = __shfl_sync(0xffffffffu, qc.x, src), sy = __shfl_sync(0xffffffffu, qc.y, src),
                        sz = __shfl_sync(0xffffffffu, qc.z, src);
            const int j = (int)S.rec_c[base + src][s2][v] * 32 + lane;  // row j exists; padding rows hold +inf
            const float4 kk = r4[j];
            const float d = sqdist1(sx, sy, sz, kk.x, kk.y, kk.z);
            const bool ok = d == d && j < nr;  // NaN keys are skipped (key 0 was checked in the main loop)
            const unsigned int bits = ok ? __float_as_uint(d) : 0xffffffffu;  // d >= 0: unsigned order == float order
            const unsigned int mnb = __reduce_min_sync(0xffffffffu, bits);
            const unsigned int mni = __reduce_min_sync(0xffffffffu, bits == mnb ? (unsigned int)j : 0x7fffffffu);
            if (lane == src && mnb != 0xffffffffu) {
              const float dd = __uint_as_float(mnb);
              if (!any || dd < bd || (dd == bd && (int)mni < bi)) {  // the sets' chunks interleave: lexicographic
                bd = dd;
                bi = (int)mni;
                any = true;
              }
            }
          }
        }
      }
      slow = slow || !any;
      if (stats && lane == 0 && nres) atomicAdd(&stats[1], nres);
      unsigned int need = __ballot_sync(0xffffffffu, live && slow);
      if (stats && lane == 0 && need) atomicAdd(&stats[0], (unsigned int)__popc(need));
      while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        const float sx = __shfl_sync(0xffffffffu, qc.x, src), sy = __shfl_sync(0xffffffffu, qc.y, src),
                    sz = __shfl_sync(0xffffffffu, qc.z, src);
        float rd;
        int ri;
        nn_exact_warp(sx, sy, sz, rp, nr, lane, rd, ri);
        if (lane == src) {
          bd = rd;
          bi = ri;
        }
      }
      if (live) {
        dout[q] = bd;
        iout[q] = bi;
      }
    }
    if (warp == 2) NT_STAMP(8);
  }

  fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * NT_N);
}

// ---- host ----------------------------------------------------------------------------------------------------
static inline int nt_pad(int n) { return (n + NT_N - 1) / NT_N * NT_N; }

// Both directions of the Chamfer search.  PCC_ENOTSUP outside the shapes this path covers (the caller keeps the SIMT
// kernels for those).  `stats` may be null.
int nn_tc_forward(int b, int n, const float *xyz1, int m, const float *xyz2, float *dist1, int *idx1, float *dist2,
                  int *idx2, unsigned int *stats, cudaStream_t st
#ifdef NT_DEBUG_SCORES
                  ,
                  float *dbg
#endif
) {
  if (b <= 0 || b > 65535 || n < 256 || m < 256 || n > NT_MAX_KEYS || m > NT_MAX_KEYS) return PCC_ENOTSUP;
  const int npad1 = nt_pad(n), npad2 = nt_pad(m);
  const size_t rows1 = (size_t)b * npad1, rows2 = (size_t)b * npad2;
  unsigned char *ws = nullptr;
  const size_t bytes = (rows1 + rows2) * (NT_ROWB * 2 + sizeof(float4)) + sizeof(float) * 2 * b;
  cudaError_t e = ws_alloc((void **)&ws, bytes, st);
  if (e != cudaSuccess) return (int)e;
  unsigned char *a1 = ws, *b1 = a1 + rows1 * NT_ROWB, *a2 = b1 + rows1 * NT_ROWB, *b2 = a2 + rows2 * NT_ROWB;
  float4 *k1 = reinterpret_cast<float4 *>(b2 + rows2 * NT_ROWB), *k2 = k1 + rows1;
  float *nmax = reinterpret_cast<float *>(k2 + rows2);
  static size_t attr[64];
  const size_t smem = sizeof(NtSmem) + 1024;
  if (cudaError_t e2 = smem_optin(nn_tc_kernel, smem, attr); e2 != cudaSuccess) {
    cudaFreeAsync(ws, st);
    return (int)e2;
  }
  // query tiles per CTA: as many CTAs as fit one wave of the SMs (one CTA per SM: the keys take most of its shared memory)
  static int sms[64];
  int dev = 0;
  cudaGetDevice(&dev);
  int &nsm = sms[dev & 63];
  if (!nsm && (cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0)) nsm = 148;
  const int mx = n > m ? n : m;
  const int qt_total = (mx + NT_M - 1) / NT_M;
  int splits = nsm / (2 * b);
  splits = splits < 1 ? 1 : (splits > qt_total ? qt_total : splits);
  const int qt_per = (qt_total + splits - 1) / splits;
  splits = (qt_total + qt_per - 1) / qt_per;
  cudaMemsetAsync(nmax, 0, sizeof(float) * 2 * b, st);
  nn_tc_prep_kernel<<<dim3(b, 2, NT_PREP_PARTS), 256, 0, st>>>(n, xyz1, m, xyz2, npad1, npad2, a1, b1, a2, b2, k1, k2,
                                                                reinterpret_cast<unsigned int *>(nmax));
  nn_tc_kernel<<<dim3(splits, b, 2), NT_THREADS, smem, st>>>(n, xyz1, m, xyz2, npad1, npad2, qt_per, a1, b1, a2, b2, k1, k2, nmax,
                                                           dist1, idx1, dist2, idx2, stats
#ifdef NT_DEBUG_SCORES
                                                           ,
                                                           dbg
#endif
  );
  cudaFreeAsync(ws, st);
  note_route(R_NN_GRID);
  return finish_launch(2);
}

}  // namespace pcc
