// Chamfer nearest-neighbour search on the tcgen05 tensor cores (sm_100a) -- pcc_nndistance_tc, the tensor-core ALTERNATIVE
// to the SIMT forward of pcc_nndistance for clouds of 256 .. 2560 points (bit-identical results; DESIGN.md section 6
// has the measurements that decide which one is the default).
//
// Replaces external/pytorch_structural_losses/src/nndistance.cu:2-128 (NmDistanceKernel, launched once per direction).
// The brute-force search evaluates B*N*M squared distances on the FP32 pipe (8 flop each).  Here the distances come out
// of the tensor cores as a CANDIDATE FILTER and only the handful of candidates per query is evaluated with the
// reference's arithmetic, so the results (distance bits and lowest-index tie rule) are identical to the SIMT kernels.
//
// What bounds such a kernel is not the tensor pipe but reading the scores back: tcgen05.ld delivers 128 bytes of
// REGISTER data per clock and SM (tools/tmem_bw_probe), i.e. 32 fp32 scores -- the 2.7e8 scores of B=32 x 2048 x 2048 x 2
// directions would take 30 us.  With .pack::16b a load returns two 16-bit accumulator columns per register in the same
// time, so the scores are accumulated in fp16 (D = F16) from fp16 operands:
//
//   prep   nn_tc_prep_kernel     per cloud and side: translate by a sample mean of the second cloud, scale by a power of
//                                two so that the largest coordinate lands in [64, 128) (norms < 49152 fit fp16), split
//                                every coordinate into two fp16 pieces x = h1 + h2 (22 bits, residual 2^-22) and write one
//                                16-column operand row per point in the two roles it plays,
//                                  query role A_i = [-2h1 -2h1 -2h2]_{x,y,z}  1  1  gA1 gA2 0 0 0
//                                  key role   B_j = [  h1   h2   h1]_{x,y,z} gB1 gB2  1   1  0 0 0       (g = |.|^2 = g1 + g2)
//                                so that A_i . B_j = |a|^2 + |b|^2 - 2 a.b up to 2^-21 |a||b| -- ONE tcgen05.mma (K = 16)
//                                per tile, summed inside the tensor core and rounded to fp16 once, when the value is the
//                                small distance itself.  Rows are stored in the UMMA canonical K-major no-swizzle order
//                                (8-row groups of 2 core matrices): a tile of rows is one contiguous block for cp.async.bulk.
//   main   nn_tc_kernel          persistent CTA = one cloud, one direction, a contiguous range of 128-query tiles.  ALL keys
//                                of the cloud are brought into shared memory once (32 B per key) and stay there: operand
//                                traffic from L2 is one pass per CTA (a streaming version was L2-bound at 4.8 TB/s).
//                                warp 0 loads, warp 1 issues tcgen05.mma kind::f16 (M=128, N=256, K=16) into a
//                                double-buffered TMEM accumulator (512 columns), sixteen epilogue warps in four sets own
//                                one query per thread and 64 keys of every key tile: one packed tcgen05.ld, 32 HMNMX2 for
//                                the minima over the even and the odd keys (two "half chunks" of 32 keys), and a short
//                                list of the half chunks whose minimum is within the error band of the running row minimum.
//                                At the end of a query tile the sets exchange their row minima and keep, per query, the
//                                half chunks still within the band of the FINAL minimum (two slots per set); after the
//                                last tile every recorded half chunk is evaluated exactly by a whole warp
//                                (d = fma(dz,dz,fma(dx,dx,dy*dy)), lexicographic minimum of (distance, key index)).
//   The candidate set provably contains the exact nearest neighbour and every key tied with it: the score of pair (i,j)
//   differs from the exact (scaled) distance by at most eps_ij = NT_REL d_ij + NT_CEPS (|a_i|^2 + |b_j|^2) (fp16 rounding of
//   the result; dropped products and the accumulation -- measured by tools/nn_tc_probe), and a half chunk is kept when its
//   minimum is <= (1 + 2 NT_REL) row minimum + 2 max_j NT_CEPS (...).  Queries whose record overflows (massive ties), whose
//   scores are not finite (NaN / inf coordinates, distances beyond the fp16 range) or whose first key is NaN (the reference
//   lets a NaN at k = 0 stick, nndistance.cu:26) are redone by an exact warp-cooperative scan with the reference's NaN rule.
#include "tc16.cuh"

namespace pcc {

constexpr int NT_M = 128;          // queries per tile = TMEM lanes
constexpr int NT_N = 256;          // keys per tile = accumulator columns per TMEM stage
constexpr int NT_ROWB = 32;        // bytes per operand row: 16 fp16
constexpr int NT_SETS = 4;         // epilogue warp sets (4 warps each: one per TMEM lane quarter), 64 keys per set and tile
constexpr int NT_CAP = 8;          // running candidate half chunks per query and set
constexpr int NT_REC = 8;          // recorded candidate chunks per query
constexpr int NT_SLOW = 0x10000;   // flag in the record count: this query takes the exact scan
constexpr int NT_THREADS = 64 + 128 * NT_SETS;  // warp 0 producer, warp 1 MMA, 16 epilogue warps
constexpr int NT_EPI = 128 * NT_SETS;
constexpr int NT_PREP_PARTS = 4;   // CTAs per cloud and side in the operand preparation
constexpr int NT_TILE_BYTES = NT_N * NT_ROWB;
constexpr int NT_MAX_KEYS = 4096;  // resident keys (16 tiles = 128 KiB) ...
constexpr int NT_MAX_Q = 2048;     // ... and queries per CTA (records)
constexpr float NT_REL = 9.765625e-4f;    // 2^-10: the fp16 result is within one ulp whatever the rounding mode
constexpr float NT_CEPS = 3.8146973e-6f;  // 2^-18 (|a|^2 + |b|^2); tools/nn_tc_probe measures the actual maximum

struct NtCtl {
  uint64_t kfull[NT_MAX_KEYS / NT_N], k4full, afull[2], aempty[2], tfull[2], tempty[2];
  uint32_t tmem_base;
};
struct NtSmem {
  unsigned char keys[NT_MAX_KEYS * NT_ROWB];
  unsigned char a[2][NT_M * NT_ROWB];
  float lst_v[NT_SETS][NT_CAP][NT_M];
  unsigned short lst_c[NT_SETS][NT_CAP][NT_M];
  float mrow[NT_SETS][NT_M];           // row minimum of each set (exchanged at the end of a query tile)
  unsigned char flag[NT_SETS][NT_M];   // set asks for the exact scan of this query
  unsigned short rec_c[NT_MAX_Q][NT_REC];  // per query of the CTA: candidate chunks (slots handed out by atomicAdd)
  int rec_n[NT_MAX_Q];                     // ... how many were found (> NT_REC: overflow); NT_SLOW set = exact scan
  unsigned short slowq[NT_MAX_Q];                   // queries of the CTA that need the exact scan
  int nslow;
  NtCtl ctl;
};

// grid (b, 2, NT_PREP_PARTS): side 0 = xyz1 (n points), side 1 = xyz2 (m points).  Every CTA takes the mean of 32 sample
// points of cloud 2 as the common translation (any translation is valid, a central one keeps the norms -- and with them
// the error band -- small), finds the largest translated coordinate of BOTH clouds (the power-of-two scale must be common)
// and converts its share of the rows.  ops* = [cloud][npad rows][32 B]; key4 = the ORIGINAL coordinates as float4 (the
// exact resolution reads them with one 16-byte load per key); meta[cloud] = {scale^2, largest scaled norm side 0, side 1}
// (must be zeroed before the launch).
__global__ void __launch_bounds__(256)
nn_tc_prep_kernel(int n, const float *__restrict__ xyz1, int m, const float *__restrict__ xyz2, int npad1, int npad2,
                  unsigned char *__restrict__ opsA1, unsigned char *__restrict__ opsB1, unsigned char *__restrict__ opsA2,
                  unsigned char *__restrict__ opsB2, float4 *__restrict__ key41, float4 *__restrict__ key42,
                  unsigned int *__restrict__ meta /* [b][4] float bits */) {
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
  // largest |translated coordinate| over both clouds (finite values only)
  float mxa = 0.f;
  for (int i0 = threadIdx.x; i0 < n + m; i0 += 256 * 4) {
    float v[12];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * 256;
      const bool ok = i < n + m;
      const float *p = !ok ? p1 : (i < n ? p1 + (size_t)i * 3 : p2 + (size_t)(i - n) * 3);
#pragma unroll
      for (int a = 0; a < 3; ++a) v[u * 3 + a] = ok ? p[a] - ctr[a] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 12; ++u) {
      const float av = fabsf(v[u]);
      if (av < INF) mxa = fmaxf(mxa, av);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mxa = fmaxf(mxa, __shfl_xor_sync(0xffffffffu, mxa, o));
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = mxa;
  __syncthreads();
  mxa = smax[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mxa = fmaxf(mxa, smax[w]);
  __syncthreads();
  int ex = 0;
  if (mxa > 0.f) frexpf(mxa, &ex);  // mxa = f 2^ex, f in [0.5, 1)
  ex = max(-100, min(100, ex));
  const float sc = mxa > 0.f ? ldexpf(1.f, 7 - ex) : 1.f;  // largest scaled coordinate in [64, 128)

  const int cnt = side ? m : n, npad = side ? npad2 : npad1;
  const float *p = side ? p2 : p1;
  unsigned char *oa = (side ? opsA2 : opsA1) + cloud * (size_t)npad * NT_ROWB;
  unsigned char *ob = (side ? opsB2 : opsB1) + cloud * (size_t)npad * NT_ROWB;
  float4 *k4 = (side ? key42 : key41) + cloud * (size_t)npad;
  const unsigned short one = 0x3c00;
  float nm = 0.f;
  for (int r = blockIdx.z * 256 + threadIdx.x; r < npad; r += 256 * NT_PREP_PARTS) {
    uint4 a0, a1, b0, b1;
    if (r < cnt) {
      const float ox = p[(size_t)r * 3], oy = p[(size_t)r * 3 + 1], oz = p[(size_t)r * 3 + 2];
      k4[r] = make_float4(ox, oy, oz, 0.f);
      const float x = (ox - ctr[0]) * sc, y = (oy - ctr[1]) * sc, z = (oz - ctr[2]) * sc;
      const float nn = fmaf(z, z, fmaf(y, y, x * x));
      if (nn < INF) nm = fmaxf(nm, nn);
      unsigned short x1, x2, y1, y2, z1, z2, g1, g2;
      f16x2(x, x1, x2);
      f16x2(y, y1, y2);
      f16x2(z, z1, z2);
      f16x2(nn, g1, g2);
      const unsigned short nx1 = hneg2(x1), nx2 = hneg2(x2), ny1 = hneg2(y1), ny2 = hneg2(y2), nz1 = hneg2(z1), nz2 = hneg2(z2);
      // columns 0..8: (a1 b1) (a1 b2) (a2 b1) per coordinate; 9, 10: 1 x gB; 11, 12: gA x 1
      a0 = make_uint4(pk2(nx1, nx1), pk2(nx2, ny1), pk2(ny1, ny2), pk2(nz1, nz1));
      a1 = make_uint4(pk2(nz2, one), pk2(one, g1), pk2(g2, 0), 0u);
      b0 = make_uint4(pk2(x1, x2), pk2(x1, y1), pk2(y2, y1), pk2(z1, z2));
      b1 = make_uint4(pk2(z1, g1), pk2(g2, one), pk2(one, 0), 0u);
    } else {  // padding: as a key it scores >= 65504 (never a candidate), as a query its row is never written
      k4[r] = make_float4(INF, INF, INF, 0.f);
      a0 = a1 = b0 = make_uint4(0u, 0u, 0u, 0u);
      b1 = make_uint4(pk2(0, 0x7bff), 0u, 0u, 0u);  // gB1 = 65504
    }
    *reinterpret_cast<uint4 *>(oa + nt_off(r, 0)) = a0;
    *reinterpret_cast<uint4 *>(oa + nt_off(r, 1)) = a1;
    *reinterpret_cast<uint4 *>(ob + nt_off(r, 0)) = b0;
    *reinterpret_cast<uint4 *>(ob + nt_off(r, 1)) = b1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nm = fmaxf(nm, __shfl_xor_sync(0xffffffffu, nm, o));
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = nm;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = smax[0];
    for (int w = 1; w < 8; ++w) v = fmaxf(v, smax[w]);
    atomicMax(&meta[cloud * 4 + 1 + side], __float_as_uint(v));  // v >= 0: unsigned order == float order
    if (blockIdx.z == 0 && side == 0) meta[cloud * 4] = __float_as_uint(sc * sc);
  }
}

__device__ __forceinline__ void nt_bar_epilogue() { asm volatile("bar.sync 1, %0;" ::"n"(NT_EPI) : "memory"); }

// rare: the list of candidate half chunks is full -> keep what is still within the band of the current minimum
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
// [blockIdx.x * qt_per, ...).  Candidate chunk c = the 64 keys 64 c .. 64 c + 63 (one set's share of a key tile).
__global__ void __launch_bounds__(NT_THREADS, 1)
nn_tc_kernel(int n, const float *__restrict__ xyz1, int m, const float *__restrict__ xyz2, int npad1, int npad2, int qt_per,
             const unsigned char *__restrict__ opsA1, const unsigned char *__restrict__ opsB1,
             const unsigned char *__restrict__ opsA2, const unsigned char *__restrict__ opsB2,
             const float4 *__restrict__ key41, const float4 *__restrict__ key42, const float *__restrict__ meta,
             float *__restrict__ dist1, int *__restrict__ idx1, float *__restrict__ dist2, int *__restrict__ idx2,
             unsigned int *__restrict__ stats /* may be null: [0] exact rescans, [1] resolved half chunks */
#ifdef NT_DEBUG_SCORES
             ,
             float *__restrict__ dbg /* [nq][npadr] raw scores of cloud 0, direction 0, divided by scale^2 */
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
  float *dout = (dir ? dist2 : dist1) + cloud * (size_t)nq;
  int *iout = (dir ? idx2 : idx1) + cloud * (size_t)nq;
  const int ntile = npadr / NT_N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  NtCtl *ctl = &S.ctl;
  // the float4 copies of the keys (exact resolution, exact scans) ride along in shared memory when they fit behind the
  // operand rows (up to 2730 keys); larger clouds read them from L2
  const bool k4s = (size_t)npadr * (NT_ROWB + sizeof(float4)) <= sizeof(S.keys);
  const float4 *kq = k4s ? reinterpret_cast<const float4 *>(S.keys + (size_t)npadr * NT_ROWB) : r4;

  if (threadIdx.x == 0) {
    for (int t = 0; t < NT_MAX_KEYS / NT_N; ++t) mbar_init(&ctl->kfull[t], 1);
    mbar_init(&ctl->k4full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->afull[s], 1);
      mbar_init(&ctl->aempty[s], 1);
      mbar_init(&ctl->tfull[s], 1);
      mbar_init(&ctl->tempty[s], 4 * NT_SETS);
    }
    S.nslow = 0;
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
      if (k4s) {
        mbar_expect_tx(&ctl->k4full, (uint32_t)(npadr * sizeof(float4)));
        bulk_load_1d(S.keys + (size_t)npadr * NT_ROWB, r4, (uint32_t)(npadr * sizeof(float4)), &ctl->k4full);
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
    // ===== MMA issuer: one K = 16 instruction per (query tile, key tile) =====
    if (lane == 0) {
      constexpr uint32_t IDESC = umma_idesc_f16(NT_M, NT_N);
      int it = 0;
      for (int i = 0; i < nqt; ++i) {
        const int sa = i & 1;
        mbar_wait(&ctl->afull[sa], (i >> 1) & 1);
        const uint64_t a_desc = umma_desc_k16(smem_u32(S.a[sa]));
        for (int t = 0; t < ntile; ++t, ++it) {
          const int acc = it & 1;
          if (i == 0) mbar_wait(&ctl->kfull[t], 0);
          mbar_wait(&ctl->tempty[acc], ((it >> 1) & 1) ^ 1);  // every epilogue warp has this accumulator in registers
          fence_after();
#ifdef NT_DEBUG_STAMPS
          if (probe_cta && it >= 8 && it < 16) stats[32 + (it - 8)] = (unsigned int)(clock64() - t_start);
#endif
#ifdef NT_DEBUG_STAMPS
          if (!(stats && (stats[127] & 1)))
#endif
          mma_f16(tmem_base + (uint32_t)(acc * NT_N), a_desc, umma_desc_k16(smem_u32(S.keys + (size_t)t * NT_TILE_BYTES)), IDESC, 0u);
          mma_commit(&ctl->tfull[acc]);
          if (it == 0) NT_STAMP(4);
        }
        mma_commit(&ctl->aempty[sa]);  // the query slot may be refilled once these MMAs have read it
      }
      NT_STAMP(5);
    }
  } else {
    // ===== epilogue: one query per thread and set; the four sets split the 256 keys of every tile =====
    const int set = (warp - 2) >> 2;
    const int quarter = warp & 3;  // this warp may touch TMEM lanes 32*quarter .. +31
    const int e = quarter * 32 + lane;
    const float INF = __int_as_float(0x7f800000);
    const float *mt = meta + cloud * 4;
    // band: a kept half chunk has minimum <= (1 + 2 NT_REL) m + babs, babs = 2 NT_CEPS (max |a|^2 + max |b|^2) (scaled units)
    const float babs = 2.f * NT_CEPS * (mt[1] + mt[2]);
    const float brel = 1.f + 2.f * NT_REL;
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(set * 64);
    float *lv = &S.lst_v[set][0][e];
    unsigned short *lc = &S.lst_c[set][0][e];
    const uint32_t tfull_a = smem_u32(&ctl->tfull[0]), tempty_a = smem_u32(&ctl->tempty[0]);
    // The tcgen05.ld of key tile it+1 is issued before the minima of tile it are computed: the TMEM read-out (the
    // bottleneck: 512 clocks per tile at 128 B/clk) runs under the ALU work instead of in lockstep with it.
    float mrow = INF;
    int cnt = 0;
    bool slow = false;
    const int total = nqt * ntile;
    auto step = [&](uint32_t (&w)[32], uint32_t (&wn)[32], int it, int i, int t) {
      const int acc = it & 1;
      tmem_ld_wait_dep(w);  // tile `it` has landed in registers
#ifdef NT_DEBUG_STAMPS
      if (probe_cta && warp == 2 && lane == 0 && it >= 8 && it < 16) stats[48 + (it - 8)] = (unsigned int)(clock64() - t_start);
#endif
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(tempty_a + acc * 8);  // ... so the MMA after next may overwrite its accumulator
      if (it + 1 < total) {
        mbar_wait_a(tfull_a + (acc ^ 1) * 8, ((it + 1) >> 1) & 1);
        fence_after();
#ifdef NT_DEBUG_STAMPS
        if (probe_cta && warp == 2 && lane == 0 && it >= 8 && it < 16) stats[64 + (it - 8)] = (unsigned int)(clock64() - t_start);
#endif
        tmem_ld64h_issue(tlane + (uint32_t)((acc ^ 1) * NT_N), wn);
      }
#ifdef NT_DEBUG_SCORES
      if (dbg && cloud == 0 && dir == 0 && (qt0 + i) * NT_M + e < nq)
        for (int u = 0; u < 32; ++u) {
          const float2 f2 = __half22float2(*reinterpret_cast<const __half2 *>(&w[u]));
          float *o = dbg + (size_t)((qt0 + i) * NT_M + e) * npadr + t * NT_N + set * 64 + 2 * u;
          o[0] = f2.x / mt[0];
          o[1] = f2.y / mt[0];
        }
#endif
#ifdef NT_DEBUG_STAMPS
      if (stats && (stats[127] & 2)) {
        if (t == ntile - 1) {
          nt_bar_epilogue();
          S.rec_n[i * NT_M + e] = NT_SLOW;
          nt_bar_epilogue();
        }
        mrow = __uint_as_float(w[3]);
        return;
      }
#endif
      if (t == 0 && set == 0) {
        const __half s0 = __low2half(*reinterpret_cast<const __half2 *>(&w[0]));
        slow = slow || __hisnan(s0);  // a NaN distance to key 0 sticks in the reference
      }
      __half2 h0 = __hmin2(*reinterpret_cast<const __half2 *>(&w[0]), *reinterpret_cast<const __half2 *>(&w[1]));
      __half2 h1 = __hmin2(*reinterpret_cast<const __half2 *>(&w[2]), *reinterpret_cast<const __half2 *>(&w[3]));
#pragma unroll
      for (int u = 4; u < 32; u += 4) {
        h0 = __hmin2(__hmin2(*reinterpret_cast<const __half2 *>(&w[u]), *reinterpret_cast<const __half2 *>(&w[u + 1])), h0);
        h1 = __hmin2(__hmin2(*reinterpret_cast<const __half2 *>(&w[u + 2]), *reinterpret_cast<const __half2 *>(&w[u + 3])), h1);
      }
      const float2 cm2 = __half22float2(__hmin2(h0, h1));  // .x: even keys of the 64, .y: odd keys
      const float cmb = fminf(cm2.x, cm2.y);
      if (cmb <= fmaf(mrow, brel, babs)) {
        if (cnt == NT_CAP) cnt = nt_compact(lv, lc, fmaf(fminf(mrow, cmb), brel, babs));
        if (cnt < NT_CAP) {
          lv[cnt * NT_M] = cmb;
          lc[cnt * NT_M] = (unsigned short)(t * NT_SETS + set);
          ++cnt;
        } else {
          slow = true;  // more than NT_CAP chunks tie within the band
        }
      }
      mrow = fminf(mrow, cmb);
#ifdef NT_DEBUG_STAMPS
      if (probe_cta && warp == 2 && lane == 0 && it >= 8 && it < 16) stats[80 + (it - 8)] = (unsigned int)(clock64() - t_start);
#endif
      if (t == ntile - 1) {
        // ---- end of the query tile: final row minimum over the sets, record the half chunks still within its band ----
        S.mrow[set][e] = mrow;
        S.flag[set][e] = slow ? 1 : 0;
        if (set == 0) S.rec_n[i * NT_M + e] = 0;
        nt_bar_epilogue();
        float mfin = S.mrow[0][e];
        unsigned int fl = S.flag[0][e];
#pragma unroll
        for (int s2 = 1; s2 < NT_SETS; ++s2) {
          mfin = fminf(mfin, S.mrow[s2][e]);
          fl |= S.flag[s2][e];
        }
        const float lim = fmaf(mfin, brel, babs);
        const int ql = i * NT_M + e;
        if (fl || !(mfin < INF)) {
          if (set == 0) atomicOr(&S.rec_n[ql], NT_SLOW);
        } else {
          for (int r = 0; r < cnt; ++r)
            if (lv[r * NT_M] <= lim) {
              const int slot = atomicAdd(&S.rec_n[ql], 1);
              if (slot < NT_REC) S.rec_c[ql][slot] = lc[r * NT_M];
            }
        }
        nt_bar_epilogue();  // the exchange buffers are free for the next query tile
        mrow = INF;
        cnt = 0;
        slow = false;
      }
    };
    {
      uint32_t wa[32], wb[32];
      mbar_wait_a(tfull_a, 0);
      fence_after();
      if (warp == 2) NT_STAMP(9);
      tmem_ld64h_issue(tlane, wa);
      int i = 0, t = 0;
      for (int it = 0; it < total; it += 2) {
        step(wa, wb, it, i, t);
        if (++t == ntile) {
          t = 0;
          ++i;
        }
        if (it + 1 < total) {
          step(wb, wa, it + 1, i, t);
          if (++t == ntile) {
            t = 0;
            ++i;
          }
        }
      }
    }
    if (warp == 2) NT_STAMP(6);

    // ---- exact resolution.  Every warp takes 32 queries at a time; their (query, half chunk) pairs are evaluated by the
    //      whole warp, one pair per step: lane l computes the distance to key 64 (hc >> 1) + 2 l + (hc & 1) (coalesced reads
    //      of the float4 copies -- a thread-per-query scan touches 32 cache lines per load instruction and was 10x slower),
    //      two REDUX give (minimum distance, lowest index attaining it), the owner lane keeps the lexicographic minimum ----
    const int ew = warp - 2;  // 0 .. 15
    if (k4s) mbar_wait(&ctl->k4full, 0);
    for (int base = ew * 32; base < nqt * NT_M; base += (NT_EPI / 32) * 32) {
      const int ql = base + lane;
      const int q = qt0 * NT_M + ql;
      const bool live = q < nq;  // base < nqt * NT_M is warp-uniform and a multiple of 32: ql is inside the records
      float bd = INF;
      int bi = 0;
      bool any = false, slow = false;
      float4 qc = make_float4(0.f, 0.f, 0.f, 0.f);
      int np = 0;
      if (live) {
        qc = q4[q];
        np = S.rec_n[ql];
        slow = np > NT_REC;  // flagged, or more candidate chunks than slots
        if (slow) np = 0;
      }
      // the warp's (query, chunk) pairs, compacted into its 256-entry slice of the (now idle) candidate lists
      uint32_t *pairs = reinterpret_cast<uint32_t *>(&S.lst_v[0][0][0]) + ew * 256;
      int incl = np;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) incl = scan_up_add(incl, o);
      const int npairs = __shfl_sync(0xffffffffu, incl, 31);
      __syncwarp();
      for (int v = 0, pos = incl - np; v < np; ++v) pairs[pos++] = ((uint32_t)lane << 16) | S.rec_c[ql][v];
      __syncwarp();
      const unsigned int nres = (unsigned int)npairs;
      for (int p0 = 0; p0 < npairs; p0 += 4) {  // eight independent key loads in flight per lane
        uint32_t pr[4];
        float4 kk[4][2];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          pr[u] = pairs[min(p0 + u, npairs - 1)];
          const float4 *kp = kq + (int)(pr[u] & 0xffffu) * 64;  // the rows exist; padding rows hold +inf
          kk[u][0] = kp[lane];
          kk[u][1] = kp[32 + lane];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (p0 + u < npairs) {  // warp-uniform
            const int src = (int)(pr[u] >> 16), j = (int)(pr[u] & 0xffffu) * 64 + lane;
            const float sx = __shfl_sync(0xffffffffu, qc.x, src), sy = __shfl_sync(0xffffffffu, qc.y, src),
                        sz = __shfl_sync(0xffffffffu, qc.z, src);
            const float d0 = sqdist1(sx, sy, sz, kk[u][0].x, kk[u][0].y, kk[u][0].z);
            const float d1 = sqdist1(sx, sy, sz, kk[u][1].x, kk[u][1].y, kk[u][1].z);
            // NaN keys are skipped (key 0 was checked in the main loop); d >= 0: unsigned order == float order
            const unsigned int b0 = (d0 == d0 && j < nr) ? __float_as_uint(d0) : 0xffffffffu;
            const unsigned int b1 = (d1 == d1 && j + 32 < nr) ? __float_as_uint(d1) : 0xffffffffu;
            const unsigned int bits = min(b0, b1);
            const unsigned int jj = b0 <= b1 ? (unsigned int)j : (unsigned int)(j + 32);  // ties: the lower index
            const unsigned int mnb = __reduce_min_sync(0xffffffffu, bits);
            const unsigned int mni = __reduce_min_sync(0xffffffffu, bits == mnb ? jj : 0x7fffffffu);
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
      __syncwarp();  // the pair slice is rewritten by the next batch of queries
      slow = slow || !any;
      if (stats && lane == 0 && nres) atomicAdd(&stats[1], nres);
      if (live && slow) {  // rare: queued for the exact scan by all sixteen warps below
        const int slot = atomicAdd(&S.nslow, 1);
        S.slowq[slot] = (unsigned short)ql;
      } else if (live) {
        dout[q] = bd;
        iout[q] = bi;
      }
    }
    nt_bar_epilogue();
    // ---- exact scan of the queued queries with the reference's semantics (nndistance.cu:2-124: ascending keys, strict
    //      '<', a NaN distance at key 0 sticks, later NaNs are skipped): one warp per query, lanes stride over the keys ----
    const int nslow = S.nslow;
    if (stats && threadIdx.x == 64 && nslow) atomicAdd(&stats[0], (unsigned int)nslow);
    for (int sidx = ew; sidx < nslow; sidx += NT_EPI / 32) {
      const int q = qt0 * NT_M + (int)S.slowq[sidx];
      const float4 qc = q4[q];
      float best = INF;
      int besti = 0x7fffffff;
#pragma unroll 4
      for (int j = lane; j < nr; j += 32) {
        const float4 kk = kq[j];
        const float d = sqdist1(qc.x, qc.y, qc.z, kk.x, kk.y, kk.z);
        if (d < best) {  // ascending j per lane: the first minimum of the lane is its lowest index
          best = d;
          besti = j;
        }
      }
      const unsigned int bits = __float_as_uint(best);  // best >= 0 or +inf: unsigned order == float order
      const unsigned int mnb = __reduce_min_sync(0xffffffffu, bits);
      const unsigned int mni = __reduce_min_sync(0xffffffffu, bits == mnb ? (unsigned int)besti : 0x7fffffffu);
      if (lane == 0) {
        const float4 k0 = kq[0];
        const float d0 = sqdist1(qc.x, qc.y, qc.z, k0.x, k0.y, k0.z);
        const bool key0 = d0 != d0 || mni == 0x7fffffffu;  // NaN at key 0 sticks; nothing below +inf: key 0
        dout[q] = key0 ? d0 : __uint_as_float(mnb);
        iout[q] = key0 ? 0 : (int)mni;
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
  // the float4 copies of the keys must fit behind the operand rows in shared memory (npad <= 2730)
  if (b <= 0 || b > 65535 || n < 256 || m < 256 || n > 2560 || m > 2560) return PCC_ENOTSUP;
  const int npad1 = nt_pad(n), npad2 = nt_pad(m);
  const size_t rows1 = (size_t)b * npad1, rows2 = (size_t)b * npad2;
  unsigned char *ws = nullptr;
  const size_t bytes = (rows1 + rows2) * (NT_ROWB * 2 + sizeof(float4)) + sizeof(float) * 4 * b;
  cudaError_t e = ws_alloc((void **)&ws, bytes, st);
  if (e != cudaSuccess) return (int)e;
  unsigned char *a1 = ws, *b1 = a1 + rows1 * NT_ROWB, *a2 = b1 + rows1 * NT_ROWB, *b2 = a2 + rows2 * NT_ROWB;
  float4 *k1 = reinterpret_cast<float4 *>(b2 + rows2 * NT_ROWB), *k2 = k1 + rows1;
  float *meta = reinterpret_cast<float *>(k2 + rows2);
  static size_t attr[64];
  const size_t smem = sizeof(NtSmem) + 1024;
  if (cudaError_t e2 = smem_optin(nn_tc_kernel, smem, attr); e2 != cudaSuccess) {
    cudaFreeAsync(ws, st);
    return (int)e2;
  }
  // query tiles per CTA: as many CTAs as fit one wave of the SMs (one CTA per SM: the keys take most of its shared
  // memory), at most NT_MAX_Q queries each
  static int sms[64];
  int dev = 0;
  cudaGetDevice(&dev);
  int &nsm = sms[dev & 63];
  if (!nsm && (cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0)) nsm = 148;
  const int mx = n > m ? n : m;
  const int qt_total = (mx + NT_M - 1) / NT_M;
  int splits = nsm / (2 * b);
  const int min_splits = (qt_total * NT_M + NT_MAX_Q - 1) / NT_MAX_Q;
  splits = splits < min_splits ? min_splits : (splits > qt_total ? qt_total : splits);
  const int qt_per = (qt_total + splits - 1) / splits;
  splits = (qt_total + qt_per - 1) / qt_per;
  cudaMemsetAsync(meta, 0, sizeof(float) * 4 * b, st);
  nn_tc_prep_kernel<<<dim3(b, 2, NT_PREP_PARTS), 256, 0, st>>>(n, xyz1, m, xyz2, npad1, npad2, a1, b1, a2, b2, k1, k2,
                                                                reinterpret_cast<unsigned int *>(meta));
  nn_tc_kernel<<<dim3(splits, b, 2), NT_THREADS, smem, st>>>(n, xyz1, m, xyz2, npad1, npad2, qt_per, a1, b1, a2, b2, k1, k2, meta,
                                                           dist1, idx1, dist2, idx2, stats
#ifdef NT_DEBUG_SCORES
                                                           ,
                                                           dbg
#endif
  );
  cudaFreeAsync(ws, st);
  note_route(R_NN_TC);
  return finish_launch(2);
}

}  // namespace pcc

extern "C" __attribute__((visibility("default"))) int pcc_nndistance_tc(int b, int n, const float *xyz, int m, const float *xyz2,
                                                                         float *result, int *result_i, float *result2,
                                                                         int *result2_i, pcc_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return PCC_EINVAL;
  if (b == 0 || n == 0 || m == 0) return PCC_OK;
  return pcc::nn_tc_forward(b, n, xyz, m, xyz2, result, result_i, result2, result2_i, nullptr, (cudaStream_t)stream);
}
