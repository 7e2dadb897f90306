// xyz kNN graph (DGCNN first layer, decoder graph_filtering) with tcgen05 as the candidate filter -- sm_100a.
//
// Replaces the KeOps argKmin behind src/utils/neighbour_ops.py:63-82 for C = 3, 256 <= N <= 2048, k <= 32.  The
// warp-cooperative SIMT kernel (knn.cu: knn3w_kernel) spends ~570 warp instructions per query of which 96 are distances:
// the selection is the cost, and it is the same for k = 4 and k = 25.  Here the N x N squared distances come out of the
// tensor cores in fp16 (the operand rows of chamfer_tc.cu: two fp16 pieces per coordinate, ONE K = 16 tcgen05.mma per
// 128 x 128 tile, scores rounded to fp16 once) and one THREAD owns one query from the first score to the output row:
//
//   sweep 1   packed tcgen05.ld of 64 scores (two per register, .pack::16b), HMNMX2 tree -> the minima over the even and
//             over the odd keys of the 64 = two groups of 32 keys; N/32 <= 64 group minima per query.
//   tau       the k-th smallest group minimum (bitonic network in registers): at least k keys score <= tau, so the k-th
//             smallest exact distance is <= tau + eps.
//   sweep 2   the MMAs are issued again (the tensor pipe is idle otherwise); HSET2 + LOP3 turn "score <= limit" into two
//             32-bit masks per 64 scores (one instruction per score); every set bit is a candidate: its EXACT distance
//             (d = fma(dz,dz,fma(dy,dy,dx*dx)), the canonical order of the kNN path) enters the thread's own top-k SET in
//             shared memory (append, then replace-the-worst).
//   output    the warp writes its 32 contiguous output rows cooperatively; an entry's slot is its rank by (distance, index).
// Exactness: the score of pair (i,j) differs from the exact (scaled) distance by at most NT-style bounds
// eps_ij = K3T_REL d_ij + K3T_CEPS (|a_i|^2 + |b_j|^2) (fp16 rounding of the result, dropped products, accumulation; measured
// for the same operands by tools/nn_tc_probe), the limit is (1 + 2 REL) tau + 2 max eps.  A query whose limit leaves the
// fp16 range (NaN / inf coordinates, a k-th neighbour farther than the format reaches) scans every key exactly.
// Two warp sets own two query tiles at a time (double-buffered accumulators each: 4 x 128 TMEM columns); the keys of the
// cloud -- operand rows and float4 copies -- stay in shared memory for all query tiles of the CTA.
#include <cstdlib>

#include "tc16.cuh"

namespace pcc {

constexpr int K3T_M = 128;        // queries per tile = TMEM lanes
constexpr int K3T_N = 128;        // keys per MMA tile = accumulator columns per stage
constexpr int K3T_SETS = 2;       // warp sets = query tiles in flight
constexpr int K3T_THREADS = 64 + 128 * K3T_SETS;
constexpr int K3T_EPI = 128 * K3T_SETS;
constexpr int K3T_MAX_N = 2048;   // resident keys: 64 KiB of operand rows + 32 KiB of float4 copies
constexpr int K3T_MAX_K = 32;
constexpr int K3T_MAX_G = K3T_MAX_N / 32;  // group minima per query
constexpr int K3T_TILE_BYTES = K3T_N * T16_ROWB;
constexpr int K3T_CAP = 64;       // candidates per query (more: the query scans every key)
constexpr int K3T_CSTRIDE = K3T_CAP + 1;
constexpr int K3T_PREP_PARTS = 4;
constexpr float K3T_REL = 9.765625e-4f;    // 2^-10
constexpr float K3T_CEPS = 3.8146973e-6f;  // 2^-18 (|a|^2 + |b|^2)

struct K3tCtl {
  uint64_t kfull, afull[K3T_SETS], aempty[K3T_SETS], tfull[K3T_SETS][2], tempty[K3T_SETS][2];
  uint32_t tmem_base;
};
struct K3tSmem {
  unsigned char keys[K3T_MAX_N * T16_ROWB];       // B rows of the whole cloud
  float4 key4[K3T_MAX_N];                         // original coordinates
  unsigned char a[K3T_SETS][K3T_M * T16_ROWB];    // A rows of the two query tiles in flight
  // per epilogue warp (its 32 lanes move in lockstep between the phases, so the arrays of one warp may alias in time):
  //   gm  [group][lane]  fp16 group minima of sweep 1              (inside cd: dead once tau is known)
  //   cd  [lane][65]     exact distances of the lane's candidates  (row stride 65 words: conflict-free by lane AND by slot)
  //   cj  [lane][65]     their key indices
  //   scr [64]           one row's distance bits, contiguous, for the rank counting
  float cd[K3T_EPI / 32][32 * K3T_CSTRIDE];
  unsigned short cj[K3T_EPI / 32][32 * K3T_CSTRIDE];
  uint32_t scr[K3T_EPI / 32][64];
  K3tCtl ctl;
};

// grid (b, K3T_PREP_PARTS).  x: coordinate c of point i at x[cloud*3n + i*ps + c*cs] ((ps, cs) = (1, n) channels-first,
// (3, 1) point-major).  opsA / opsB = [cloud][npad rows][32 B], key4 = [cloud][npad], meta[cloud] = {scale^2, largest scaled norm}
// (zeroed before the launch).
__global__ void __launch_bounds__(256)
knn3_tc_prep_kernel(int n, int ps, int cs, const float *__restrict__ x, int npad, unsigned char *__restrict__ opsA,
                    unsigned char *__restrict__ opsB, float4 *__restrict__ key4, unsigned int *__restrict__ meta) {
  __shared__ float smax[8];
  __shared__ float sctr[3];
  const size_t cloud = blockIdx.x;
  const float *p = x + cloud * (size_t)3 * n;
  const float INF = __int_as_float(0x7f800000);
  if (threadIdx.x < 32) {  // the same 32 points and the same arithmetic in every CTA of the cloud: identical centres
    const float *sp = p + (size_t)((long long)threadIdx.x * n / 32) * ps;
    const float vx = sp[0], vy = sp[cs], vz = sp[2 * (size_t)cs];
    const bool ok = fabsf(vx) < INF && fabsf(vy) < INF && fabsf(vz) < INF;
    const float c = warp_sum(ok ? 1.f : 0.f);
    const float sx = warp_sum(ok ? vx : 0.f), sy = warp_sum(ok ? vy : 0.f), sz = warp_sum(ok ? vz : 0.f);
    if (threadIdx.x == 0) {
      const float inv = c > 0.f ? 1.f / c : 0.f;
      sctr[0] = sx * inv;
      sctr[1] = sy * inv;
      sctr[2] = sz * inv;
    }
  }
  __syncthreads();
  const float ctr[3] = {sctr[0], sctr[1], sctr[2]};
  float mxa = 0.f;
#pragma unroll 4
  for (int i = threadIdx.x; i < n; i += 256) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float av = fabsf(p[(size_t)i * ps + (size_t)a * cs] - ctr[a]);
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
  if (mxa > 0.f) frexpf(mxa, &ex);
  ex = max(-100, min(100, ex));
  const float sc = mxa > 0.f ? ldexpf(1.f, 7 - ex) : 1.f;  // largest scaled coordinate in [64, 128)

  unsigned char *oa = opsA + cloud * (size_t)npad * T16_ROWB, *ob = opsB + cloud * (size_t)npad * T16_ROWB;
  float4 *k4 = key4 + cloud * (size_t)npad;
  const unsigned short one = 0x3c00;
  float nm = 0.f;
#pragma unroll 2
  for (int r = blockIdx.y * 256 + threadIdx.x; r < npad; r += 256 * K3T_PREP_PARTS) {
    uint4 a0, a1, b0, b1;
    if (r < n) {
      const float ox = p[(size_t)r * ps], oy = p[(size_t)r * ps + cs], oz = p[(size_t)r * ps + 2 * (size_t)cs];
      k4[r] = make_float4(ox, oy, oz, 0.f);
      const float vx = (ox - ctr[0]) * sc, vy = (oy - ctr[1]) * sc, vz = (oz - ctr[2]) * sc;
      const float nn = fmaf(vz, vz, fmaf(vy, vy, vx * vx));
      if (nn < INF) nm = fmaxf(nm, nn);
      unsigned short x1, x2, y1, y2, z1, z2, g1, g2;
      f16x2(vx, x1, x2);
      f16x2(vy, y1, y2);
      f16x2(vz, z1, z2);
      f16x2(nn, g1, g2);
      const unsigned short nx1 = hneg2(x1), nx2 = hneg2(x2), ny1 = hneg2(y1), ny2 = hneg2(y2), nz1 = hneg2(z1), nz2 = hneg2(z2);
      a0 = make_uint4(pk2(nx1, nx1), pk2(nx2, ny1), pk2(ny1, ny2), pk2(nz1, nz1));
      a1 = make_uint4(pk2(nz2, one), pk2(one, g1), pk2(g2, 0), 0u);
      b0 = make_uint4(pk2(x1, x2), pk2(x1, y1), pk2(y2, y1), pk2(z1, z2));
      b1 = make_uint4(pk2(z1, g1), pk2(g2, one), pk2(one, 0), 0u);
    } else {  // padding: as a key it scores >= 65504 (never below a finite limit), as a query its row is never written
      k4[r] = make_float4(INF, INF, INF, 0.f);
      a0 = a1 = b0 = make_uint4(0u, 0u, 0u, 0u);
      b1 = make_uint4(pk2(0, 0x7bff), 0u, 0u, 0u);
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
    atomicMax(&meta[cloud * 2 + 1], __float_as_uint(v));
    if (blockIdx.y == 0) meta[cloud * 2] = __float_as_uint(sc * sc);
  }
}

template <int N>
__device__ __forceinline__ void k3t_bitonic(float (&a)[N]) {  // ascending, fully unrolled (static indices)
#pragma unroll
  for (int k2 = 2; k2 <= N; k2 <<= 1)
#pragma unroll
    for (int j = k2 >> 1; j > 0; j >>= 1)
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const int l = i ^ j;
        if (l > i) {
          const float lo = fminf(a[i], a[l]), hi = fmaxf(a[i], a[l]);
          const bool up = (i & k2) == 0;
          a[i] = up ? lo : hi;
          a[l] = up ? hi : lo;
        }
      }
}

// k-th smallest (1-based k <= ng <= NG) of the thread's group minima
template <int NG>
__device__ __forceinline__ float k3t_kth(const __half *gm, int ng, int k) {
  float a[NG];
#pragma unroll
  for (int u = 0; u < NG; ++u) a[u] = u < ng ? __half2float(gm[u * 32]) : __int_as_float(0x7f800000);
  k3t_bitonic<NG>(a);
  float tau = -__int_as_float(0x7f800000);
#pragma unroll
  for (int u = 0; u < NG; ++u) tau = u < k ? fmaxf(tau, a[u]) : tau;  // a[k-1] without dynamic register indexing
  return tau;
}

// grid (splits, b).  blockIdx.x owns the query tiles [blockIdx.x * qt_per, ...) of cloud blockIdx.y.
__global__ void __launch_bounds__(K3T_THREADS, 1)
knn3_tc_kernel(int n, int k, int npad, int qt_per, const unsigned char *__restrict__ opsA, const unsigned char *__restrict__ opsB,
               const float4 *__restrict__ key4g, const float *__restrict__ meta, int64_t *__restrict__ idx_out,
               float *__restrict__ dist_out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  K3tSmem &S = *reinterpret_cast<K3tSmem *>(smem_raw);
  const size_t cloud = blockIdx.y;
  const int qt_total = (n + K3T_M - 1) / K3T_M;
  const int qt0 = blockIdx.x * qt_per, qt1 = min(qt_total, qt0 + qt_per);
  if (qt0 >= qt1) return;  // uniform per CTA
  const int nqt = qt1 - qt0, niter = (nqt + K3T_SETS - 1) / K3T_SETS;
  const int ntile = npad / K3T_N;
  const unsigned char *opsq = opsA + cloud * (size_t)npad * T16_ROWB;
  const unsigned char *opsr = opsB + cloud * (size_t)npad * T16_ROWB;
  const float4 *k4g = key4g + cloud * (size_t)npad;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  K3tCtl *ctl = &S.ctl;

  if (threadIdx.x == 0) {
    mbar_init(&ctl->kfull, 1);
    for (int s = 0; s < K3T_SETS; ++s) {
      mbar_init(&ctl->afull[s], 1);
      mbar_init(&ctl->aempty[s], 1);
      for (int b2 = 0; b2 < 2; ++b2) {
        mbar_init(&ctl->tfull[s][b2], 1);
        mbar_init(&ctl->tempty[s][b2], 4);
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&ctl->tmem_base, 512);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ===== producer: all keys (operand rows + float4 copies) once, then the query tiles of every iteration =====
    if (lane == 0) {
      mbar_expect_tx(&ctl->kfull, (uint32_t)(npad * (T16_ROWB + sizeof(float4))));
      bulk_load_1d(S.keys, opsr, (uint32_t)(npad * T16_ROWB), &ctl->kfull);
      bulk_load_1d(S.key4, k4g, (uint32_t)(npad * sizeof(float4)), &ctl->kfull);
      for (int i = 0; i < niter; ++i)
        for (int s = 0; s < K3T_SETS; ++s) {
          const int qt = qt0 + i * K3T_SETS + s;
          if (qt >= qt1) continue;
          mbar_wait(&ctl->aempty[s], (i & 1) ^ 1);
          mbar_expect_tx(&ctl->afull[s], K3T_M * T16_ROWB);
          bulk_load_1d(S.a[s], opsq + (size_t)qt * K3T_M * T16_ROWB, K3T_M * T16_ROWB, &ctl->afull[s]);
        }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: per iteration two sweeps over the key tiles, alternating between the two query tiles =====
    if (lane == 0) {
      constexpr uint32_t IDESC = umma_idesc_f16(K3T_M, K3T_N);
      mbar_wait(&ctl->kfull, 0);
      int it[K3T_SETS] = {0, 0};  // accumulator uses per set
      for (int i = 0; i < niter; ++i) {
        const bool have1 = qt0 + i * K3T_SETS + 1 < qt1;
        mbar_wait(&ctl->afull[0], i & 1);
        if (have1) mbar_wait(&ctl->afull[1], i & 1);
        for (int t2 = 0; t2 < 2 * ntile; ++t2) {
          const int t = t2 < ntile ? t2 : t2 - ntile;
          const uint64_t b_desc = umma_desc_k16(smem_u32(S.keys + (size_t)t * K3T_TILE_BYTES));
#pragma unroll
          for (int s = 0; s < K3T_SETS; ++s) {
            if (s == 1 && !have1) continue;
            const int acc = it[s] & 1;
            mbar_wait(&ctl->tempty[s][acc], ((it[s] >> 1) & 1) ^ 1);
            fence_after();
            mma_f16(tmem_base + (uint32_t)((s * 2 + acc) * K3T_N), umma_desc_k16(smem_u32(S.a[s])), b_desc, IDESC, 0u);
            mma_commit(&ctl->tfull[s][acc]);
            ++it[s];
          }
        }
        mma_commit(&ctl->aempty[0]);
        if (have1) mma_commit(&ctl->aempty[1]);
      }
    }
  } else {
    // ===== epilogue: one query per thread; set s owns query tile qt0 + 2 i + s in iteration i =====
    const int set = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int e = quarter * 32 + lane;  // TMEM lane = row of the query tile
    const int et = set * 128 + e;
    const float INF = __int_as_float(0x7f800000);
    const float sc2 = meta[cloud * 2], nmax = meta[cloud * 2 + 1];
    const float babs = 2.f * K3T_CEPS * (nmax + nmax), brel = 1.f + 2.f * K3T_REL;
    const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(set * 2 * K3T_N);
    const uint32_t tfull_a = smem_u32(&ctl->tfull[set][0]), tempty_a = smem_u32(&ctl->tempty[set][0]);
    // two packed loads per tile; per load two groups (even / odd keys of the 64), or four groups of 16 keys (even / odd keys of
    // the registers u % 4 < 2 and of the others) while 64 group minima hold them: a tighter tau, fewer candidates
    const bool fine = ntile * 8 <= K3T_MAX_G;
    const int ng = fine ? ntile * 8 : ntile * 4;
    const int ew = warp - 2;
    __half *gm = reinterpret_cast<__half *>(&S.cd[ew][0]) + lane;  // group g at gm[g * 32]
    float *cd = &S.cd[ew][lane * K3T_CSTRIDE];
    unsigned short *cj = &S.cj[ew][lane * K3T_CSTRIDE];
    uint32_t *scr = &S.scr[ew][0];
    mbar_wait(&ctl->kfull, 0);  // the exact distances read the float4 copies
    int it = 0;
    (void)sc2;
    (void)et;
    for (int i = 0; i < niter; ++i) {
      const int qt = qt0 + i * K3T_SETS + set;
      if (qt >= qt1) break;  // uniform per warp set; the MMA issuer skips this set as well
      const int q = qt * K3T_M + e;
      const bool live = q < n;
      const float4 qc = S.key4[min(q, npad - 1)];
      // ---- sweep 1: group minima ----
      for (int t = 0; t < ntile; ++t, ++it) {
        const int acc = it & 1;
        mbar_wait_a(tfull_a + acc * 8, (it >> 1) & 1);
        fence_after();
        uint32_t w0[32], w1[32];
        tmem_ld64h_issue(tbase + (uint32_t)(acc * K3T_N), w0);
        tmem_ld64h_issue(tbase + (uint32_t)(acc * K3T_N + 64), w1);
        tmem_ld_wait_dep(w0);
        tmem_ld_wait_dep(w1);
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(tempty_a + acc * 8);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t(&w)[32] = h ? w1 : w0;
          __half2 h0 = __hmin2(*reinterpret_cast<const __half2 *>(&w[0]), *reinterpret_cast<const __half2 *>(&w[1]));
          __half2 h1 = __hmin2(*reinterpret_cast<const __half2 *>(&w[2]), *reinterpret_cast<const __half2 *>(&w[3]));
#pragma unroll
          for (int u = 4; u < 32; u += 4) {
            h0 = __hmin2(__hmin2(*reinterpret_cast<const __half2 *>(&w[u]), *reinterpret_cast<const __half2 *>(&w[u + 1])), h0);
            h1 = __hmin2(__hmin2(*reinterpret_cast<const __half2 *>(&w[u + 2]), *reinterpret_cast<const __half2 *>(&w[u + 3])), h1);
          }
          if (fine) {
            gm[(t * 8 + h * 4) * 32] = __low2half(h0);
            gm[(t * 8 + h * 4 + 1) * 32] = __high2half(h0);
            gm[(t * 8 + h * 4 + 2) * 32] = __low2half(h1);
            gm[(t * 8 + h * 4 + 3) * 32] = __high2half(h1);
          } else {
            const __half2 m2 = __hmin2(h0, h1);
            gm[(t * 4 + h * 2) * 32] = __low2half(m2);
            gm[(t * 4 + h * 2 + 1) * 32] = __high2half(m2);
          }
        }
      }
      // ---- tau = k-th smallest group minimum; limit in fp16, rounded up ----
      const float tau = ng <= 32 ? k3t_kth<32>(gm, ng, k) : k3t_kth<64>(gm, ng, k);
      const float lim = fmaf(tau, brel, babs);
      bool all = !(lim < 60000.f);  // beyond the fp16 range (or NaN): every key is a candidate, taken exactly
      const __half2 lim2 = __half2half2(__float2half_ru(all ? 0.f : lim));
      __syncwarp();  // gm is dead: the candidate arrays take its place
      // ---- sweep 2: candidate indices = keys whose score is within the limit (appended, nothing else) ----
      int cnt = 0;
      for (int t = 0; t < ntile; ++t, ++it) {
        const int acc = it & 1;
        mbar_wait_a(tfull_a + acc * 8, (it >> 1) & 1);
        fence_after();
        uint32_t w0[32], w1[32];
        tmem_ld64h_issue(tbase + (uint32_t)(acc * K3T_N), w0);
        tmem_ld64h_issue(tbase + (uint32_t)(acc * K3T_N + 64), w1);
        tmem_ld_wait_dep(w0);
        tmem_ld_wait_dep(w1);
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(tempty_a + acc * 8);
        if (!live || all) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t(&w)[32] = h ? w1 : w0;
          uint32_t m0 = 0, m1 = 0;  // bit u (u+16): even (odd) key of register u / register 16+u
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            m0 |= __hle2_mask(*reinterpret_cast<const __half2 *>(&w[u]), lim2) & (0x00010001u << u);
            m1 |= __hle2_mask(*reinterpret_cast<const __half2 *>(&w[16 + u]), lim2) & (0x00010001u << u);
          }
          const int j0 = t * K3T_N + h * 64;
          if (cnt + __popc(m0) + __popc(m1) > K3T_CAP) {
            all = true;  // massive ties: the exact scan below
            m0 = m1 = 0;
          }
          while (m0) {
            const int bit = __ffs(m0) - 1;
            m0 &= m0 - 1;
            cj[cnt++] = (unsigned short)(j0 + 2 * (bit & 15) + (bit >> 4));
          }
          while (m1) {
            const int bit = __ffs(m1) - 1;
            m1 &= m1 - 1;
            cj[cnt++] = (unsigned short)(j0 + 32 + 2 * (bit & 15) + (bit >> 4));
          }
        }
      }
      auto exact = [&](int j) {
        const float4 kk = S.key4[j];
        const float dx = qc.x - kk.x, dy = qc.y - kk.y, dz = qc.z - kk.z;
        return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
      };
      if (live && all) {
        // exact scan of every key: the k best by (distance, index) through a sorted insertion into the candidate arrays
        cnt = 0;
        for (int j = 0; j < n; ++j) {
          const float d = exact(j);
          if (!(d == d)) continue;
          if (cnt == k && !(d < cd[k - 1])) continue;  // ascending j: an equal distance never displaces an earlier key
          int p2 = cnt < k ? cnt : k - 1;
          while (p2 > 0 && d < cd[p2 - 1]) {
            cd[p2] = cd[p2 - 1];
            cj[p2] = cj[p2 - 1];
            --p2;
          }
          cd[p2] = d;
          cj[p2] = (unsigned short)j;
          if (cnt < k) ++cnt;
        }
      } else if (live) {
        // ---- exact distances of the candidates, all lanes in step; NaN distances sort last and are dropped ----
      }
      {
        const int maxc = __reduce_max_sync(0xffffffffu, (live && !all) ? cnt : 0);
        int drop = 0;
        for (int c = 0; c < maxc; ++c)
          if (live && !all && c < cnt) {
            const int j = cj[c];
            float d = exact(j);
            if (!(d == d) || j >= n) {
              d = __int_as_float(0x7fffffff);  // sorts behind everything
              ++drop;
            }
            cd[c] = d;
          }
        if (!live) cnt = 0;
        // ---- rows: the warp ranks one query at a time, (at most) two candidates per lane, by counting smaller distance
        //      bits; rank r < k is output slot r.  d >= 0: the bit patterns order like the values.  Without exact ties the
        //      strict ranks are a permutation (their sum is c (c-1) / 2); ties are broken by index in a second count. ----
        __syncwarp();
        const int q_w = qt * K3T_M + quarter * 32;  // first query of this warp
        const int rows = min(32, n - q_w);
        int64_t *ob = idx_out + ((size_t)cloud * n + q_w) * k;
        float *odb = dist_out ? dist_out + ((size_t)cloud * n + q_w) * k : nullptr;
        for (int r = 0; r < rows; ++r) {
          const int c = __shfl_sync(0xffffffffu, cnt, r);
          const int good = c - __shfl_sync(0xffffffffu, drop, r);  // candidates with a finite-or-inf (non-NaN) distance
          const float *rd = &S.cd[ew][r * K3T_CSTRIDE];
          const unsigned short *rj = &S.cj[ew][r * K3T_CSTRIDE];
          const uint32_t me0 = lane < c ? __float_as_uint(rd[lane]) : 0x7fffffffu;
          const uint32_t me1 = lane + 32 < c ? __float_as_uint(rd[lane + 32]) : 0x7fffffffu;
          const uint32_t i0 = lane < c ? rj[lane] : 0xffffu, i1 = lane + 32 < c ? rj[lane + 32] : 0xffffu;
          scr[lane] = me0;
          scr[lane + 32] = me1;
          __syncwarp();
          int rk0 = 0, rk1 = 0;
          const uint4 *sv = reinterpret_cast<const uint4 *>(scr);
          const int c4 = (c + 3) >> 2;
          if (c <= 32) {
#pragma unroll 2
            for (int u = 0; u < c4; ++u) {
              const uint4 kk = sv[u];
              rk0 += ((kk.x - me0) >> 31) + ((kk.y - me0) >> 31) + ((kk.z - me0) >> 31) + ((kk.w - me0) >> 31);
            }
          } else {
#pragma unroll 2
            for (int u = 0; u < c4; ++u) {
              const uint4 kk = sv[u];
              rk0 += ((kk.x - me0) >> 31) + ((kk.y - me0) >> 31) + ((kk.z - me0) >> 31) + ((kk.w - me0) >> 31);
              rk1 += ((kk.x - me1) >> 31) + ((kk.y - me1) >> 31) + ((kk.z - me1) >> 31) + ((kk.w - me1) >> 31);
            }
          }
          const int ssum = __reduce_add_sync(0xffffffffu, (lane < c ? rk0 : 0) + (lane + 32 < c ? rk1 : 0));
          if (ssum != c * (c - 1) / 2) {  // exact ties: break them by index
            rk0 = rk1 = 0;
            for (int u = 0; u < c; ++u) {
              const uint32_t du = scr[u], iu = rj[u];
              rk0 += (du < me0 || (du == me0 && iu < i0)) ? 1 : 0;
              rk1 += (du < me1 || (du == me1 && iu < i1)) ? 1 : 0;
            }
          }
          int64_t *o = ob + (size_t)r * k;
          float *od = odb ? odb + (size_t)r * k : nullptr;
          const int kk2 = good < k ? good : k;
          if (lane < c && rk0 < kk2) {
            o[rk0] = (int64_t)i0;
            if (od) od[rk0] = __uint_as_float(me0);
          }
          if (lane + 32 < c && rk1 < kk2) {
            o[rk1] = (int64_t)i1;
            if (od) od[rk1] = __uint_as_float(me1);
          }
          for (int t2 = kk2 + lane; t2 < k; t2 += 32) {  // fewer than k non-NaN distances only with NaN inputs
            o[t2] = 0;
            if (od) od[t2] = INF;
          }
          __syncwarp();
        }
      }
      __syncwarp();  // the arrays are reused by the next query tile
    }
  }

  fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// x (b,3,n) channels-first, or (b,n,3) point-major with pm.  PCC_ENOTSUP outside the shapes this path covers.
int knn3_tc_launch(int b, int n, int k, bool pm, const float *x, int64_t *idx, float *dist, cudaStream_t st) {
  if (b <= 0 || b > 65535 || n < 256 || n > K3T_MAX_N || k < 1 || k > K3T_MAX_K) return PCC_ENOTSUP;
  const int npad = (n + K3T_N - 1) / K3T_N * K3T_N;
  const int ngroups = npad / 128 * 8 <= K3T_MAX_G ? npad / 16 : npad / 32;  // group minima per query (see the kernel)
  if (ngroups < k) return PCC_ENOTSUP;
  // Where this path wins over knn3w_kernel (tools/knn_time.py on B200, B = 32): N = 2048 -- 107 vs 160 us at k = 25, 78 vs
  // 142 us at k = 4; at N = 1024 the two tie (47 vs 44 us at k = 20) and below that, with few clouds (fewer than ~100 query
  // tiles in flight) or with k close to N / 32 (a loose threshold) the SIMT kernel is faster.  PCC_KNN3_TC=1 forces it.
  const bool force = getenv("PCC_KNN3_TC") != nullptr;  // read per call: tests switch it
  if (!force && (n <= 1024 || ngroups < 2 * k || (long long)b * (npad / K3T_M) < 96)) return PCC_ENOTSUP;
  const size_t rows = (size_t)b * npad;
  unsigned char *ws = nullptr;
  cudaError_t e = ws_alloc((void **)&ws, rows * (2 * T16_ROWB + sizeof(float4)) + sizeof(float) * 2 * b, st);
  if (e != cudaSuccess) return (int)e;
  unsigned char *oa = ws, *ob = oa + rows * T16_ROWB;
  float4 *k4 = reinterpret_cast<float4 *>(ob + rows * T16_ROWB);
  float *meta = reinterpret_cast<float *>(k4 + rows);
  static size_t attr[64];
  const size_t smem = sizeof(K3tSmem) + 1024;
  if (cudaError_t e2 = smem_optin(knn3_tc_kernel, smem, attr); e2 != cudaSuccess) {
    cudaFreeAsync(ws, st);
    return (int)e2;
  }
  static int sms[64];
  int dev = 0;
  cudaGetDevice(&dev);
  int &nsm = sms[dev & 63];
  if (!nsm && (cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0)) nsm = 148;
  const int qt_total = (n + K3T_M - 1) / K3T_M;
  // CTAs of an even number of query tiles (two are in flight), about one wave of the SMs
  int splits = nsm / b;
  splits = splits < 1 ? 1 : splits;
  int qt_per = (qt_total + splits - 1) / splits;
  qt_per = (qt_per + 1) & ~1;
  splits = (qt_total + qt_per - 1) / qt_per;
  cudaMemsetAsync(meta, 0, sizeof(float) * 2 * b, st);
  knn3_tc_prep_kernel<<<dim3(b, K3T_PREP_PARTS), 256, 0, st>>>(n, pm ? 3 : 1, pm ? 1 : n, x, npad, oa, ob, k4,
                                                                reinterpret_cast<unsigned int *>(meta));
  knn3_tc_kernel<<<dim3(splits, b), K3T_THREADS, smem, st>>>(n, k, npad, qt_per, oa, ob, k4, meta, idx, dist);
  cudaFreeAsync(ws, st);
  return finish_launch(2);
}

}  // namespace pcc
