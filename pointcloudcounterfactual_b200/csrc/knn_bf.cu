// Feature-space kNN graph (DGCNN EdgeConv, C = 32 / 64 channels) on tcgen05 -- third generation: scores precise enough
// to ORDER the candidates, so the exact fp32 re-rank (40 % of knn_tc2's time: 26 candidates x 512 B of shared memory per
// query) is only needed inside a narrow ambiguity band.
//
// Alternative to knn_tc2 for the KeOps argKmin behind src/utils/neighbour_ops.py:77-82 when only the indices are requested
// (knn / pykeops_knn / get_graph_features).  EXPERIMENTAL, opt-in with PCC_KNN_BF=1: correct on every parity input but
// slower than knn_tc2 -- the measurements are at knn_bf_launch below and in DESIGN.md section 9.2.
//
//   operands  every value is split into two bf16 pieces, x = b1 + b2 + r with |r| <= 2^-18 |x| (bf16 has fp32's exponent
//             range: no scaling).  kind::f16 MMAs with fp32 accumulation, K = 16:
//               sweep 1   a1.b1 + ext            LOOSE score, made an upper bound by an extra column pair inside the MMA
//               sweep 2   a1.b1 + a1.b2 + a2.b1 + ext     precise score: error <= 2^-15 |x_i||x_j| + 2^-18 (n_i + n_j)
//             with a = -2 x and ext = the squared norm n_j in three bf16 pieces against ones: the accumulator IS the
//             score n_j - 2 x_i.x_j, the epilogue does no arithmetic on it.  Half the bytes of a 3xTF32 split and twice
//             its MMA rate, the same accuracy class.
//   CTA       256 queries of one cloud (two M = 128 accumulators share every key tile of 64 keys), warp 0 = bulk-copy
//             producer (tiles are pre-laid-out by the prep kernel in the canonical no-swizzle K-major layout), warp 1 =
//             MMA issuer, 8 epilogue warps, one THREAD per query from the first score to the output row.
//   sweep 1   minima over groups of 16 / 32 keys (64 per query); tau = k-th smallest: at least k keys score <= tau.
//   sweep 2   predicated append of every precise score <= tau + E (the key's column is kept in the 6 low mantissa bits).
//   final     per thread: candidates sorted by the lower end of their interval [s - w, s + w] (w = score error + the
//             rounding of the oracle's fp32 fma chain, per candidate), overlapping intervals merged into clusters.  A
//             cluster of one is ranked by its score alone; the members of larger clusters that can reach the top k get
//             the exact canonical distance (fp32 fma chain over the channels, from global memory) and are ordered by
//             (distance, index) inside their cluster.  Indices are bit-identical to the SIMT kernel / the oracle.
//   fallback  a query whose lists overflow (massive exact ties) or a cloud with non-finite values scans every key exactly.
#include <cuda_bf16.h>

#include <cstdlib>

#include "tc_ptx.cuh"

namespace pcc {

constexpr int KBF_R = 64;        // keys per tile = accumulator columns per stage
constexpr int KBF_Q = 256;       // queries per CTA (two halves of 128 = TMEM lanes)
constexpr int KBF_THREADS = 64 + KBF_Q;
constexpr int KBF_CAP = 64;      // candidates per query
constexpr int KBF_STRIDE = KBF_CAP + 1;  // odd: a lane's list never shares a bank with its neighbours'
constexpr int KBF_PCAP = 32;     // candidates per query that may need the exact distance
constexpr int KBF_MAX_N = 2048;
constexpr int KBF_MAX_TILES = KBF_MAX_N / KBF_R;
constexpr int KBF_MAX_K = 32;
// loose (one bf16 piece per operand): |2 x.y - 2 b1(x).b1(y)| <= 2 (2^-9 + (1 + 2^-9) 2^-9) |x||y|
constexpr float KBF_C1L = 0.0079f;          // 2^-7 * 1.011
constexpr float KBF_C2L = 9.5367432e-7f;    // 2^-20 (n_i + n_j): accumulation (5 truncating updates), slack x8
// precise: dropped b2.b2, the two residuals r: 6 * 2^-18 |x||y| -> 2^-15 with slack; accumulation 13 updates -> 2^-18
constexpr float KBF_C1P = 3.0517578e-5f;    // 2^-15
constexpr float KBF_C2P = 3.8146973e-6f;    // 2^-18
constexpr float KBF_TRUNC = 1.5258789e-5f;  // 2^-16 |s|: six mantissa bits of the stored score carry the column

template <int C>
struct KB {
  static constexpr int KA = 2 * C + 16;    // A row: a1 | a2 | ext, bf16 elements
  static constexpr int KB1 = C + 16;       // B row, sweep 1: b1 | ext
  static constexpr int KB2 = 2 * C + 16;   // B row, sweep 2: b1 | b2 | ext
  static constexpr int ROWA = KA * 2, ROWB1 = KB1 * 2, ROWB2 = KB2 * 2;  // bytes
  static constexpr int A_HALF = 128 * ROWA;
  static constexpr int B1_TILE = KBF_R * ROWB1, B2_TILE = KBF_R * ROWB2;
  static constexpr int KS = C / 16;        // K steps per piece
  // the operand tiles are dead after the last MMA: the final stage's index words and byte arrays take their place
  static constexpr size_t OPS = 2 * (size_t)A_HALF + 2 * B2_TILE;
  static constexpr size_t FIN = (size_t)KBF_Q * KBF_STRIDE * 4 + (size_t)KBF_Q * (KBF_CAP + 2 * KBF_PCAP);
  static constexpr size_t REGION = (OPS > FIN ? OPS : FIN + 1023) / 1024 * 1024;
  static constexpr size_t SMEM = REGION + (size_t)KBF_Q * KBF_STRIDE * 4 + (size_t)KBF_Q * (KBF_MAX_TILES + 4) + 256 + 1024;
};

// K-major, no swizzle, 16-bit elements: core matrix = 8 rows x 16 B (8 values); LBO = 128 B between core matrices along K,
// SBO = row bytes * 8 between 8-row groups.  Chunk kc (8 values) of row r of a tile sits at kbf_off(r, kc, row bytes).
__device__ __forceinline__ uint64_t kbf_desc(uint32_t smem_addr, uint32_t rowb) {
  return (uint64_t)((smem_addr >> 4) & 0x3fff) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)((rowb * 8) >> 4) << 32) | (1ull << 46);
}
__host__ __device__ inline size_t kbf_off(int r, int kc, int rowb) { return (size_t)(r >> 3) * rowb * 8 + (size_t)kc * 128 + (r & 7) * 16; }
// instruction descriptor: D = F32, A = B = BF16, both K-major, M x N
__host__ __device__ constexpr uint32_t kbf_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void kbf_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ unsigned short kbf_rn(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }
__device__ __forceinline__ float kbf_f(unsigned short h) { return __uint_as_float((uint32_t)h << 16); }
__device__ __forceinline__ unsigned short kbf_up(float v) {  // v >= 0: the smallest bf16 >= v
  uint32_t u = __float_as_uint(v);
  if (u & 0xffffu) u = (u & 0xffff0000u) + 0x10000u;
  return (unsigned short)(u >> 16);
}
__device__ __forceinline__ uint32_t kbf_pk(unsigned short lo, unsigned short hi) { return (uint32_t)lo | ((uint32_t)hi << 16); }

// grid (npad / 128, b), 128 threads: thread = point.  x (b,c,n) channels-first or (b,n,c) point-major (pm).
//   opsA  [cloud][npad/128 tiles][128 rows x ROWA]   canonical layout
//   ops1  [cloud][npad/64  tiles][64 rows x ROWB1],  ops2 [...][64 rows x ROWB2]
//   xT    (b,n,c) fp32 (written unless pm), nrm [cloud][npad] squared norms (0 for padding),
//   meta  [cloud][2] = {largest squared norm (float bits, atomicMax; zeroed before), non-finite flag}
template <int C>
__global__ void __launch_bounds__(128)
knn_bf_prep_kernel(int n, int npad, bool pm, const float *__restrict__ x, unsigned char *__restrict__ opsA,
                   unsigned char *__restrict__ ops1, unsigned char *__restrict__ ops2, float *__restrict__ xT,
                   float *__restrict__ nrm, unsigned int *__restrict__ meta) {
  using Cfg = KB<C>;
  const size_t cloud = blockIdx.y;
  const int rl = threadIdx.x, r = blockIdx.x * 128 + rl;
  const bool real = r < n;
  unsigned char *ta = opsA + (cloud * (npad / 128) + blockIdx.x) * (size_t)Cfg::A_HALF;
  const size_t t64 = cloud * (npad / KBF_R) + (size_t)(r / KBF_R);
  unsigned char *t1 = ops1 + t64 * Cfg::B1_TILE, *t2 = ops2 + t64 * Cfg::B2_TILE;
  const int r64 = r % KBF_R;
  const float *xb = x + cloud * (size_t)C * n;
  double nn = 0.0;
#pragma unroll 1
  for (int c0 = 0; c0 < C; c0 += 8) {
    unsigned short h1[8], h2[8], g1[8], g2[8];
    float v8[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float v = real ? (pm ? xb[(size_t)r * C + c0 + u] : xb[(size_t)(c0 + u) * n + r]) : 0.f;
      v8[u] = v;
      nn += (double)v * (double)v;
      h1[u] = kbf_rn(v);
      h2[u] = kbf_rn(v - kbf_f(h1[u]));
      g1[u] = kbf_rn(-2.f * kbf_f(h1[u]));  // exact
      g2[u] = kbf_rn(-2.f * kbf_f(h2[u]));
    }
    if (real && !pm) {
      float4 *o = reinterpret_cast<float4 *>(xT + (cloud * n + r) * (size_t)C + c0);
      o[0] = make_float4(v8[0], v8[1], v8[2], v8[3]);
      o[1] = make_float4(v8[4], v8[5], v8[6], v8[7]);
    }
    const uint4 b1 = make_uint4(kbf_pk(h1[0], h1[1]), kbf_pk(h1[2], h1[3]), kbf_pk(h1[4], h1[5]), kbf_pk(h1[6], h1[7]));
    const uint4 b2 = make_uint4(kbf_pk(h2[0], h2[1]), kbf_pk(h2[2], h2[3]), kbf_pk(h2[4], h2[5]), kbf_pk(h2[6], h2[7]));
    const uint4 a1 = make_uint4(kbf_pk(g1[0], g1[1]), kbf_pk(g1[2], g1[3]), kbf_pk(g1[4], g1[5]), kbf_pk(g1[6], g1[7]));
    const uint4 a2 = make_uint4(kbf_pk(g2[0], g2[1]), kbf_pk(g2[2], g2[3]), kbf_pk(g2[4], g2[5]), kbf_pk(g2[6], g2[7]));
    const int kc = c0 / 8;
    *reinterpret_cast<uint4 *>(ta + kbf_off(rl, kc, Cfg::ROWA)) = a1;
    *reinterpret_cast<uint4 *>(ta + kbf_off(rl, C / 8 + kc, Cfg::ROWA)) = a2;
    *reinterpret_cast<uint4 *>(t1 + kbf_off(r64, kc, Cfg::ROWB1)) = b1;
    *reinterpret_cast<uint4 *>(t2 + kbf_off(r64, kc, Cfg::ROWB2)) = b1;
    *reinterpret_cast<uint4 *>(t2 + kbf_off(r64, C / 8 + kc, Cfg::ROWB2)) = b2;
  }
  const float nf = (float)nn;
  const bool finite = nf < __int_as_float(0x7f800000);  // false for NaN as well
  // ext columns.  A: [1, 1, 1, up(c1L |x_i|), 1, 0...]; B sweep 1: [n1, n2, n3, up(|x_j|), up(c2L n_j), 0...] -- the loose
  // score comes out as an upper bound; B sweep 2: [n1, n2, n3, 0...]
  const unsigned short one = 0x3f80;
  unsigned short n1 = 0, n2 = 0, n3 = 0, sq = 0, cn = 0, cq = 0;
  if (real && finite) {
    n1 = kbf_rn(nf);
    const double r1 = nn - (double)kbf_f(n1);
    n2 = kbf_rn((float)r1);
    n3 = kbf_rn((float)(r1 - (double)kbf_f(n2)));
    const float s = sqrtf(nf) * 1.0000002f + 1e-30f;
    sq = kbf_up(s);
    cn = kbf_up(KBF_C2L * nf);
    cq = kbf_up(KBF_C1L * s);
  } else if (!real) {
    n1 = 0x7e00;  // padding key: score 1.7e38, never below a finite threshold
  }
  const uint4 eb0 = make_uint4(kbf_pk(n1, n2), kbf_pk(n3, sq), kbf_pk(cn, 0), 0u), z4 = make_uint4(0u, 0u, 0u, 0u);
  *reinterpret_cast<uint4 *>(t1 + kbf_off(r64, C / 8, Cfg::ROWB1)) = eb0;
  *reinterpret_cast<uint4 *>(t1 + kbf_off(r64, C / 8 + 1, Cfg::ROWB1)) = z4;
  *reinterpret_cast<uint4 *>(t2 + kbf_off(r64, 2 * C / 8, Cfg::ROWB2)) = make_uint4(kbf_pk(n1, n2), kbf_pk(n3, 0), 0u, 0u);
  *reinterpret_cast<uint4 *>(t2 + kbf_off(r64, 2 * C / 8 + 1, Cfg::ROWB2)) = z4;
  *reinterpret_cast<uint4 *>(ta + kbf_off(rl, 2 * C / 8, Cfg::ROWA)) = make_uint4(kbf_pk(one, one), kbf_pk(one, cq), kbf_pk(one, 0), 0u);
  *reinterpret_cast<uint4 *>(ta + kbf_off(rl, 2 * C / 8 + 1, Cfg::ROWA)) = z4;
  if (r < npad) nrm[cloud * npad + r] = real ? nf : 0.f;
  if (real) {
    if (finite) atomicMax(&meta[cloud * 2], __float_as_uint(nf));
    else meta[cloud * 2 + 1] = 1u;
  }
}

#ifdef KBF_STATS
__device__ unsigned long long kbf_stats[8];  // queries, exact scans, sum candidates, max candidates, sum ambiguous, overflow lists
#define KBF_STAT(i, v) atomicAdd(&kbf_stats[i], (unsigned long long)(v))
#define KBF_STATMAX(i, v) atomicMax(&kbf_stats[i], (unsigned long long)(v))
#else
#define KBF_STAT(i, v)
#define KBF_STATMAX(i, v)
#endif

struct KbfCtl {
  uint64_t afull, full[2], empty[2], tfull[2][2], tempty[2][2];
  uint32_t tmem_base;
};

template <int N>
__device__ __forceinline__ void kbf_bitonic(float (&a)[N]) {  // ascending, fully unrolled (static indices)
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
template <int NG>
__device__ __forceinline__ float kbf_kth(const float *gm, int ng, int k) {  // gm[u * KBF_Q]: this thread's group minima
  float a[NG];
#pragma unroll
  for (int u = 0; u < NG; ++u) a[u] = u < ng ? gm[u * KBF_Q] : __int_as_float(0x7f800000);
  kbf_bitonic<NG>(a);
  float tau = -__int_as_float(0x7f800000);
#pragma unroll
  for (int u = 0; u < NG; ++u) tau = u < k ? fmaxf(tau, a[u]) : tau;
  return tau;
}

// ord[rank of entry a by (value, slot)] = a for the cnt entries of a thread's list.  Rank counting with the inner loop
// unrolled over eight independent shared-memory loads (the lists have an odd stride: conflict-free); compact code -- four
// fully unrolled register versions (16 / 32 / 48 / 64 entries) were measured first and thrashed the instruction cache
// (17 % of the stall samples "no instruction").
__device__ __forceinline__ void kbf_rank_store(const float *myS, int cnt, int cmax, unsigned char *ord) {
  const float INF = __int_as_float(0x7f800000);
  const int c8 = (cmax + 7) & ~7;  // warp-uniform trip counts; entries beyond cnt read as +inf (the list has slack for it)
  for (int a = 0; a < cmax; ++a) {
    const float la = a < cnt ? myS[a] : INF;
    int r = 0;
    for (int u0 = 0; u0 < c8; u0 += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (u0 + u < cnt) ? myS[u0 + u] : INF;
#pragma unroll
      for (int u = 0; u < 8; ++u) r += (v[u] < la || (v[u] == la && u0 + u < a)) ? 1 : 0;
    }
    if (a < cnt) ord[r] = (unsigned char)a;
  }
}

template <int C>
__global__ void __launch_bounds__(KBF_THREADS, 1)
knn_bf_kernel(int n, int k, int npad, const unsigned char *__restrict__ opsA, const unsigned char *__restrict__ ops1,
              const unsigned char *__restrict__ ops2, const float *__restrict__ xT, const float *__restrict__ nrm,
              const unsigned int *__restrict__ meta, int64_t *__restrict__ idx_out) {
  using Cfg = KB<C>;
  extern __shared__ __align__(1024) unsigned char kbf_smem[];
  unsigned char *sA = kbf_smem;                                   // [2 halves][128 x ROWA]
  unsigned char *sB = sA + 2 * Cfg::A_HALF;                       // [2 stages][B2_TILE]
  float *sS = reinterpret_cast<float *>(sA + Cfg::REGION);        // [KBF_Q][KBF_STRIDE] scores / interval starts / distances
  unsigned char *sCT = reinterpret_cast<unsigned char *>(sS + KBF_Q * KBF_STRIDE);  // [KBF_Q][MAX_TILES + 4] list length after tile t
  KbfCtl *ctl = reinterpret_cast<KbfCtl *>(sCT + KBF_Q * (KBF_MAX_TILES + 4));
  // after the last MMA the operand tiles are dead: the final stage keeps its other per-thread arrays there
  uint32_t *sJ = reinterpret_cast<uint32_t *>(sA);                // [KBF_Q][KBF_STRIDE] index | bf16(w) << 16
  unsigned char *sORD = sA + (size_t)KBF_Q * KBF_STRIDE * 4;      // [KBF_Q][KBF_CAP] candidate at sorted position
  unsigned char *sPL = sORD + KBF_Q * KBF_CAP;                    // [KBF_Q][KBF_PCAP] candidates that need the exact distance
  unsigned char *sPB = sPL + KBF_Q * KBF_PCAP;                    // [KBF_Q][KBF_PCAP] first rank of their cluster

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t cloud = blockIdx.y;
  const int qt0 = blockIdx.x * 2, nqt = npad / 128;
  const bool have1 = qt0 + 1 < nqt;
  const int ntile = npad / KBF_R, niter = 2 * ntile;
  const unsigned char *ga = opsA + (cloud * nqt + qt0) * (size_t)Cfg::A_HALF;
  const unsigned char *g1 = ops1 + cloud * ntile * (size_t)Cfg::B1_TILE;
  const unsigned char *g2 = ops2 + cloud * ntile * (size_t)Cfg::B2_TILE;

  if (threadIdx.x == 0) {
    mbar_init(&ctl->afull, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
      for (int h = 0; h < 2; ++h) {
        mbar_init(&ctl->tfull[h][s], 1);
        mbar_init(&ctl->tempty[h][s], 4);
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&ctl->tmem_base, 256);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ===== producer =====
    if (lane == 0) {
      const uint32_t abytes = (have1 ? 2u : 1u) * Cfg::A_HALF;
      mbar_expect_tx(&ctl->afull, abytes);
      for (uint32_t o = 0; o < abytes; o += 32768u) bulk_load_1d(sA + o, ga + o, min(32768u, abytes - o), &ctl->afull);
      for (int i = 0; i < niter; ++i) {
        const int s = i & 1;
        mbar_wait(&ctl->empty[s], ((i >> 1) & 1) ^ 1);
        const bool second = i >= ntile;
        const int t = second ? i - ntile : i;
        const uint32_t bytes = second ? Cfg::B2_TILE : Cfg::B1_TILE;
        mbar_expect_tx(&ctl->full[s], bytes);
        bulk_load_1d(sB + s * Cfg::B2_TILE, second ? g2 + (size_t)t * Cfg::B2_TILE : g1 + (size_t)t * Cfg::B1_TILE, bytes,
                     &ctl->full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t IDESC = kbf_idesc(128, KBF_R);
      constexpr int KS = Cfg::KS;
      mbar_wait(&ctl->afull, 0);
      for (int i = 0; i < niter; ++i) {
        const int s = i & 1, par = (i >> 1) & 1;
        const bool second = i >= ntile;
        mbar_wait(&ctl->full[s], par);
        const uint32_t bb = smem_u32(sB + s * Cfg::B2_TILE);
        for (int h = 0; h < (have1 ? 2 : 1); ++h) {
          mbar_wait(&ctl->tempty[h][s], par ^ 1);
          fence_after();
          const uint32_t acc = tmem_base + (uint32_t)((h * 2 + s) * KBF_R);
          const uint32_t aa = smem_u32(sA + h * Cfg::A_HALF);
          if (!second) {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
              kbf_mma(acc, kbf_desc(aa + ks * 256, Cfg::ROWA), kbf_desc(bb + ks * 256, Cfg::ROWB1), IDESC, ks ? 1u : 0u);
            kbf_mma(acc, kbf_desc(aa + 2 * KS * 256, Cfg::ROWA), kbf_desc(bb + KS * 256, Cfg::ROWB1), IDESC, 1u);
          } else {
            // small cross terms first, the large a1.b1 and the norm last
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
              kbf_mma(acc, kbf_desc(aa + (KS + ks) * 256, Cfg::ROWA), kbf_desc(bb + ks * 256, Cfg::ROWB2), IDESC, ks ? 1u : 0u);
              kbf_mma(acc, kbf_desc(aa + ks * 256, Cfg::ROWA), kbf_desc(bb + (KS + ks) * 256, Cfg::ROWB2), IDESC, 1u);
            }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
              kbf_mma(acc, kbf_desc(aa + ks * 256, Cfg::ROWA), kbf_desc(bb + ks * 256, Cfg::ROWB2), IDESC, 1u);
            kbf_mma(acc, kbf_desc(aa + 2 * KS * 256, Cfg::ROWA), kbf_desc(bb + 2 * KS * 256, Cfg::ROWB2), IDESC, 1u);
          }
          mma_commit(&ctl->tfull[h][s]);
        }
        mma_commit(&ctl->empty[s]);
      }
    }
  } else {
    // ===== epilogue: one query per thread =====
    const int h = (warp - 2) >> 2, quarter = warp & 3;
    const int e = h * 128 + quarter * 32 + lane;  // query slot in the CTA
    const int q = qt0 * 128 + e;
    const bool live = q < n;
    const float INF = __int_as_float(0x7f800000);
    const float nmax = __uint_as_float(meta[cloud * 2]);
    bool all = meta[cloud * 2 + 1] != 0u;  // non-finite values in the cloud: exact scan
    const float nq = nrm[cloud * npad + min(q, npad - 1)];
    const float sqq = sqrtf(nq), sqm = sqrtf(nmax);
    // slack of the candidate threshold: score error + twice the rounding of the oracle's chain, with the largest key norm
    const float chainc = (float)(C + 2) * 1.1920929e-7f;  // (C + 2) 2^-23 per unit of distance
    const float emax = KBF_C1P * sqq * sqm + KBF_C2P * (nq + nmax) + 2.f * chainc * 2.f * (nq + nmax);
    float *myS = sS + e * KBF_STRIDE;
    uint32_t *myJ = sJ + e * KBF_STRIDE;
    unsigned char *myCT = sCT + e * (KBF_MAX_TILES + 4);
    float *gm = sS + e;  // group g at gm[g * KBF_Q] (aliases the lists: dead before the first append; 64 <= KBF_STRIDE groups)
    const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(h * 2 * KBF_R);
    const uint32_t tfull_a = smem_u32(&ctl->tfull[h][0]), tempty_a = smem_u32(&ctl->tempty[h][0]);
    const bool g16 = ntile <= 16;  // groups of 16 keys while 64 group minima hold them
    const int ng = g16 ? ntile * 4 : ntile * 2;
    int cnt = 0;
    float thr = INF;
    if (h == 0 || have1) {
      for (int i = 0; i < niter; ++i) {
        const int s = i & 1, par = (i >> 1) & 1;
        const bool second = i >= ntile;
        const int t = second ? i - ntile : i;
        mbar_wait_a(tfull_a + s * 8, par);
        fence_after();
        uint32_t v0[32], v1[32];
        tmem_ld32_issue(tl + (uint32_t)(s * KBF_R), v0);
        tmem_ld32_issue(tl + (uint32_t)(s * KBF_R + 32), v1);
        tmem_ld_wait_dep(v0);
        tmem_ld_wait_dep(v1);
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(tempty_a + s * 8);
        if (!second) {
          // ---- sweep 1: group minima of the loose upper bounds ----
          float m[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t(&v)[32] = g < 2 ? v0 : v1;
            const int o = (g & 1) * 16;
            float a = fminf(__uint_as_float(v[o]), __uint_as_float(v[o + 1])), b2 = fminf(__uint_as_float(v[o + 2]), __uint_as_float(v[o + 3]));
#pragma unroll
            for (int u = 4; u < 16; u += 4) {
              a = fminf(fminf(__uint_as_float(v[o + u]), __uint_as_float(v[o + u + 1])), a);
              b2 = fminf(fminf(__uint_as_float(v[o + u + 2]), __uint_as_float(v[o + u + 3])), b2);
            }
            m[g] = fminf(a, b2);
          }
          if (g16) {
#pragma unroll
            for (int g = 0; g < 4; ++g) gm[(t * 4 + g) * KBF_Q] = m[g];
          } else {
            gm[(t * 2) * KBF_Q] = fminf(m[0], m[1]);
            gm[(t * 2 + 1) * KBF_Q] = fminf(m[2], m[3]);
          }
          if (i == ntile - 1) {
            const float tau = ng <= 32 ? kbf_kth<32>(gm, ng, k) : kbf_kth<64>(gm, ng, k);
            // the per-query constant of the loose bound, then the candidate slack; a few ulps up: <= stays inclusive
            thr = tau + KBF_C2L * nq + emax;
            thr = thr + fabsf(thr) * 2.4e-7f + 1e-30f;
            if (!(thr < 1e37f)) all = true;  // fewer than k finite scores (or NaN): exact scan
            // the lists take the place of the group minima: every thread of the CTA must be done reading them
            asm volatile("bar.sync 1, %0;" ::"r"(have1 ? 256 : 128) : "memory");
          }
        } else {
          // ---- sweep 2: predicated append (score with the column in its six low mantissa bits) ----
          if (live && !all) {
            // straight-line predicated code (a branch per score diverges on almost every score: some lane of the warp
            // has a candidate); room for eight entries is checked once per eight scores
            uint32_t pa = smem_u32(myS + cnt);
            const uint32_t plim = smem_u32(myS + KBF_CAP - 8);
#pragma unroll
            for (int u8 = 0; u8 < 64; u8 += 8) {
              if (pa > plim) {
                all = true;  // more candidates than the list holds (massive ties): exact scan
                break;
              }
#pragma unroll
              for (int u = u8; u < u8 + 8; ++u) {
                const uint32_t raw = u < 32 ? v0[u] : v1[u - 32];
                const uint32_t tagged = (raw & 0xffffffc0u) | (uint32_t)u;
                asm volatile("{\n.reg .pred p;\nsetp.le.f32 p, %1, %2;\n@p st.shared.b32 [%0], %3;\n@p add.u32 %0, %0, 4;\n}"
                             : "+r"(pa)
                             : "f"(__uint_as_float(raw)), "f"(thr), "r"(tagged)
                             : "memory");
              }
            }
            cnt = (int)((pa - smem_u32(myS)) >> 2);
          }
          myCT[t] = (unsigned char)min(cnt, 255);
        }
      }
    }
    asm volatile("bar.sync 2, 256;" ::: "memory");  // every MMA has completed: the operand tiles are free
    if (live && (h == 0 || have1)) {
      const float *xq = xT + (cloud * n + q) * (size_t)C;
      const float *xc = xT + cloud * n * (size_t)C;
      auto exact = [&](int j) {  // the canonical sequential fma chain over the channels (oracle / SIMT order)
        const float4 *a4 = reinterpret_cast<const float4 *>(xq), *b4 = reinterpret_cast<const float4 *>(xc + (size_t)j * C);
        float d = 0.f;
#pragma unroll 4
        for (int c4 = 0; c4 < C / 4; ++c4) {
          const float4 a = a4[c4], b = b4[c4];
          const float t0 = a.x - b.x, t1 = a.y - b.y, t2 = a.z - b.z, t3 = a.w - b.w;
          d = fmaf(t0, t0, d);
          d = fmaf(t1, t1, d);
          d = fmaf(t2, t2, d);
          d = fmaf(t3, t3, d);
        }
        return d;
      };
      int64_t *out = idx_out + (cloud * n + q) * (size_t)k;
      unsigned char *ord = sORD + e * KBF_CAP, *pl = sPL + e * KBF_PCAP, *pb = sPB + e * KBF_PCAP;
      KBF_STAT(0, 1);
      KBF_STAT(2, cnt);
      KBF_STATMAX(3, cnt);
      if (cnt > KBF_CAP) {
        all = true;
        KBF_STAT(5, 1);
      }
      int npush = 0;
      const int mxc = __reduce_max_sync(__activemask(), all ? 0 : cnt);
      if (!all) {
        // ---- entries -> (interval start, index, half-width) ----
        int t = 0;
        for (int c = 0; c < cnt; ++c) {
          while (myCT[t] <= c) ++t;
          const uint32_t raw = __float_as_uint(myS[c]);
          const float s = __uint_as_float(raw & ~63u);
          const int j = t * KBF_R + (int)(raw & 63u);
          const float nj = nrm[cloud * npad + j];
          const float w = KBF_C1P * sqq * sqrtf(nj) + KBF_C2P * (nq + nj) + chainc * fmaxf(nq + s, 0.f) * 1.01f + KBF_TRUNC * fabsf(s);
          const unsigned short wb = kbf_up(w * 1.001f + 1e-30f);
          myS[c] = s - kbf_f(wb);
          myJ[c] = (uint32_t)j | ((uint32_t)wb << 16);
        }
        // ---- sort by interval start: rank counting, ties by slot (trip counts = the warp's longest list) ----
        kbf_rank_store(myS, cnt, mxc, ord);
        // ---- merge overlapping intervals into clusters; singletons are final ----
        int cs = 0;           // first rank of the open cluster
        float chi = -INF;     // its largest interval end
        for (int r = 0; r <= cnt; ++r) {
          float lo = INF, hi = INF;
          if (r < cnt) {
            const int a = ord[r];
            lo = myS[a];
            hi = lo + 2.f * kbf_f((unsigned short)(myJ[a] >> 16));
          }
          if (r > 0 && lo > chi) {  // the open cluster [cs, r) is complete
            if (cs < k) {
              if (r - cs == 1) {
                out[cs] = (int64_t)(myJ[ord[cs]] & 0xffffu);
              } else {
                for (int m2 = cs; m2 < r; ++m2) {
                  if (npush < KBF_PCAP) {
                    pl[npush] = ord[m2];
                    pb[npush] = (unsigned char)cs;
                  }
                  ++npush;
                }
              }
            }
            cs = r;
            chi = -INF;
            if (cs >= k) break;
          }
          chi = fmaxf(chi, hi);
        }
        if (npush > KBF_PCAP || cnt < k) all = true;
        KBF_STAT(4, npush);
      }
      if (all) KBF_STAT(1, 1);
      // ---- exact distances of the ambiguous candidates, all lanes in step ----
      const int maxp = __reduce_max_sync(__activemask(), all ? 0 : npush);
      for (int t2 = 0; t2 < maxp; ++t2)
        if (!all && t2 < npush) {
          const int a = pl[t2];
          myS[a] = exact((int)(myJ[a] & 0xffffu));
        }
      if (!all) {
        for (int t2 = 0; t2 < npush; ++t2) {
          const int a = pl[t2], base = pb[t2];
          const float da = myS[a];
          const uint32_t ja = myJ[a] & 0xffffu;
          int r = base;
          for (int u = 0; u < npush; ++u)
            if (pb[u] == base) {
              const float du = myS[pl[u]];
              const uint32_t ju = myJ[pl[u]] & 0xffffu;
              r += (du < da || (du == da && ju < ja)) ? 1 : 0;
            }
          if (r < k) out[r] = (int64_t)ja;
        }
      } else {
        // ---- exact scan of every key: the k best by (distance, index), sorted insertion (k <= KBF_CAP) ----
        int c2 = 0;
        for (int j = 0; j < n; ++j) {
          const float d = exact(j);
          if (!(d == d)) continue;
          if (c2 == k && !(d < myS[k - 1])) continue;  // ascending j: an equal distance never displaces an earlier key
          int p2 = c2 < k ? c2 : k - 1;
          while (p2 > 0 && d < myS[p2 - 1]) {
            myS[p2] = myS[p2 - 1];
            myJ[p2] = myJ[p2 - 1];
            --p2;
          }
          myS[p2] = d;
          myJ[p2] = (uint32_t)j;
          if (c2 < k) ++c2;
        }
        for (int r = 0; r < k; ++r) out[r] = r < c2 ? (int64_t)myJ[r] : 0;
      }
    }
  }

  fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

template <int C>
static int launch_bf(int b, int n, int k, int npad, bool pm, const float *x, int64_t *idx, cudaStream_t st) {
  using Cfg = KB<C>;
  const size_t nqt = npad / 128, nt = npad / KBF_R;
  const size_t a_bytes = (size_t)b * nqt * Cfg::A_HALF, b1_bytes = (size_t)b * nt * Cfg::B1_TILE, b2_bytes = (size_t)b * nt * Cfg::B2_TILE;
  const size_t xt_bytes = pm ? 0 : sizeof(float) * (size_t)b * n * C, nrm_bytes = sizeof(float) * (size_t)b * npad;
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  unsigned char *ws = nullptr;
  cudaError_t e = ws_alloc((void **)&ws, up(a_bytes) + up(b1_bytes) + up(b2_bytes) + up(xt_bytes) + up(nrm_bytes) + 256 + 8 * (size_t)b, st);
  if (e != cudaSuccess) return (int)e;
  unsigned char *oa = ws, *o1 = oa + up(a_bytes), *o2 = o1 + up(b1_bytes);
  float *xT = reinterpret_cast<float *>(o2 + up(b2_bytes));
  float *nrm = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(xT) + up(xt_bytes));
  unsigned int *meta = reinterpret_cast<unsigned int *>(reinterpret_cast<unsigned char *>(nrm) + up(nrm_bytes));
  static size_t attr[64];
  if (cudaError_t e2 = smem_optin(knn_bf_kernel<C>, Cfg::SMEM, attr); e2 != cudaSuccess) {
    cudaFreeAsync(ws, st);
    return (int)e2;
  }
  cudaMemsetAsync(meta, 0, 8 * (size_t)b, st);
  knn_bf_prep_kernel<C><<<dim3((unsigned)nqt, b), 128, 0, st>>>(n, npad, pm, x, oa, o1, o2, xT, nrm, meta);
  knn_bf_kernel<C><<<dim3((unsigned)((nqt + 1) / 2), b), KBF_THREADS, Cfg::SMEM, st>>>(n, k, npad, oa, o1, o2, pm ? x : xT, nrm, meta, idx);
  cudaFreeAsync(ws, st);
  return finish_launch(2);
}

// x (b,c,n) channels-first, or (b,n,c) point-major with pm; indices only.  PCC_ENOTSUP outside the shapes of this path.
int knn_bf_launch(int b, int c, int n, int k, bool pm, const float *x, int64_t *idx, float *dist, cudaStream_t st) {
  if (dist != nullptr || (c != 32 && c != 64) || b <= 0 || b > 65535 || n < 256 || n > KBF_MAX_N || k < 1 || k > KBF_MAX_K)
    return PCC_ENOTSUP;
  // Opt-in (PCC_KNN_BF=1, read per call: tests switch it).  Measured on B200, B = 32 (tools/knnf_time.py, tools/knn_bf_stats.py):
  // bit-identical indices on every test input, 25.3 candidates per query at C = 64, N = 1024, k = 20 of which only 0.40 need
  // the exact distance (knn_tc2 evaluates all 26 exactly) -- and still 101 us against 78 us for knn_tc2 (N = 2048, k = 25:
  // 305 vs 219 us).  Stall samples: the per-thread final stage (sort 25-40 intervals by rank counting, c^2 compares) 39 %,
  // waiting for its slowest warp 24 %, the predicated appends of sweep 2 (5 issue slots per score) 10 %, waits on the
  // MMA 12 %; with eight epilogue warps per SM the straight-line per-query code runs at 0.6 instructions per clock.
  // The re-rank it removes is cheaper than the selection it adds; knn_tc2 stays the default.
  if (getenv("PCC_KNN_BF") == nullptr) return PCC_ENOTSUP;
  const int npad = (n + 127) / 128 * 128;
  const int ntile = npad / KBF_R, ng = ntile <= 16 ? ntile * 4 : ntile * 2;
  if (ng < k || (reinterpret_cast<uintptr_t>(x) & 15) != 0) return PCC_ENOTSUP;
  return c == 32 ? launch_bf<32>(b, n, k, npad, pm, x, idx, st) : launch_bf<64>(b, n, k, npad, pm, x, idx, st);
}

}  // namespace pcc

#ifdef KBF_STATS
extern "C" __attribute__((visibility("default"))) int pcc_knn_bf_stats(unsigned long long *out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, pcc::kbf_stats, sizeof(unsigned long long) * 8);
  if (reset) {
    unsigned long long z[8] = {0};
    cudaMemcpyToSymbol(pcc::kbf_stats, z, sizeof(z));
  }
  return 0;
}
#endif
