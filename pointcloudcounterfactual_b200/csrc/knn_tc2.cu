// Feature-space kNN graph (DGCNN EdgeConv, C = 32..128 channels) on tcgen05 -- second generation.
//
// Replaces the KeOps argKmin behind src/utils/neighbour_ops.py:77-82.  Same contract as knn_tc.cu (TF32 tensor cores
// only GENERATE candidates, every candidate is re-ranked with the exact fp32 fma chain, so the indices are bit-identical
// to the SIMT kernel / the oracle), restructured so that nothing on the critical path touches global memory twice:
//
//   CTA      = 128*HALVES queries of one cloud (HALVES = 2 for C <= 64: two M=128 accumulators share every key tile).
//   warp 0   TMA producer: the query tile once, then key tiles of R keys x ALL channels (128B-swizzled K-major boxes)
//            plus the tile's squared norms (cp.async.bulk), two stages.
//   warp 1   MMA issuer: tcgen05.mma kind::tf32 M=128 x N=R x K=8 into TMEM accumulator `stage` of each half.
//   warps 2+ epilogue, one thread per query (tcgen05.ld of its accumulator row), two sweeps over the key tiles:
//     sweep 1  score = |x_j|^2 - 2 x_i.x_j; one minimum per group of 16 or 32 keys (at most 64 groups).  tau = k-th smallest group minimum
//              (bitonic network in registers) bounds the k-th smallest score from above.
//     sweep 2  the accumulators are recomputed (the tensor pipe is idle otherwise); keys whose score can be <= tau are
//              candidates.  The warp compacts its (query, key) pairs and spreads them evenly over its lanes; each
//              pair's EXACT distance is evaluated at once from the key tile that is still resident in shared memory
//              (the tile that fed the MMA) and the resident query tile.
//     finally  the ~1.5 k candidates of a query are ranked by counting, one query per warp step and one candidate per
//              lane; rank r < k is output slot r.
// The candidate set provably contains the exact top-k: the TF32 score s~ of pair (i,j) is off by at most
//     eps_ij = 1.03 * 2^-8 |x_i| |x_j| + 4e-5 (|x_i|^2 + |x_j|^2)
// (truncation of both operands -- < 2^-10 relative each, so < 2^-9 on every product, x2 for the -2 x.y term -- and
// Cauchy-Schwarz, 3 % slack: the tensor core's own accumulation error was measured at 1.6e-8 relative per MMA step;
// second term: fp32 rounding of norms / distances), a bound
// PER KEY: sweep 1 takes the k-th smallest group minimum of the UPPER bounds s~ + eps_ij (at least k keys are truly
// below it), sweep 2 keeps the keys whose LOWER bound s~ - eps_ij does not exceed it.  With one bound per cloud
// (max norm) instead, heavy-tailed features -- a dense core plus a few far outliers, what chained EdgeConv layers
// produce -- put hundreds of core points inside the band of every core query (measured: C=128, N=2048 kNN 1.1 ms on
// iid features, 7.3 ms on the third DGCNN layer's input).  A query whose candidates overflow the list (massive exact
// ties) is redone by exact brute force.
#include "tc_ptx.cuh"

namespace pcc {

template <int KB, int HALVES, int R>
struct T2 {
  static constexpr int C = KB * 32;
  static constexpr int QUERIES = 128 * HALVES;
  static constexpr int NEPI = 4 * HALVES;              // epilogue warps
  static constexpr int THREADS = 64 + 32 * NEPI;       // warp 0 = TMA, warp 1 = MMA
  static constexpr int A_BYTES = KB * QUERIES * 128;   // query tile: KB blocks of [QUERIES][32 floats]
  static constexpr int B_BYTES = KB * R * 128;         // one key stage
  // candidate slots per query: 56 where two M-tiles share the CTA (the records of 256 queries must fit next to the
  // operand tiles), 112 where one does -- fewer lists reach the prune path (a serial insertion sort) on the way
  static constexpr int CAP = HALVES == 1 ? 112 : 56;
  static constexpr int W = HALVES == 1 ? 169 : 85;     // words per query record: CAP distances + CAP/2 packed indices,
                                                       // >= 64 (sweep 1 keeps the group minima there), odd (banks)
  static constexpr int PCAP = 192;                     // (query, key) pairs per warp and tile
  static constexpr int CHUNKS = R / 32;                // 32-column TMEM loads per tile
  static constexpr int TMEM_COLS = HALVES * 2 * R;
  static constexpr size_t SMEM =
      A_BYTES + 2 * B_BYTES + 2 * 3 * R * 4 + (size_t)QUERIES * W * 4 + (size_t)NEPI * (PCAP + 64) * 4 + 256;
};

constexpr float T2_C1 = 0.00402832f;  // 2^-8 * 1.03125: truncation of both operands (< 2^-10 each), x2, Cauchy-Schwarz; 3 % slack
constexpr float T2_C2 = 4e-5f;

// transpose (B,C,N) -> (B,N,C), squared norms, and per key the three factors of the score bounds:
// bounds[(cloud*npad + j)/R tile][3][R] = { n_j (1+c2), n_j (1-c2), c1 sqrt(n_j) }, padding keys {inf, inf, 0}
template <int R>
__global__ void __launch_bounds__(256)
knn_tc2_prep_kernel(int c, int n, bool pm, const float *__restrict__ x, float *__restrict__ xT, float *__restrict__ norms,
                    float *__restrict__ bounds, int npad) {
  __shared__ float t[64][33];
  const size_t cloud = blockIdx.y;
  const int n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float *xb = x + cloud * (size_t)c * n;
  float *xo = xT + cloud * (size_t)n * c;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};  // thread (tx = channel lane, ty) accumulates points ty, ty+8, ty+16, ty+24
  // point-major input (b,n,c) -- what the reference hands to KeOps after x.transpose(2, 1).contiguous(),
  // neighbour_ops.py:79 -- is already the operand layout: only the norms are needed (same summation order)
  for (int c0 = 0; pm && c0 < c; c0 += 64) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int p = n0 + ty + 8 * r;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ch = c0 + tx + 32 * h;
        const float v = (p < n && ch < c) ? xb[(size_t)p * c + ch] : 0.f;
        acc[r] = fmaf(v, v, acc[r]);
      }
    }
  }
  for (int c0 = 0; !pm && c0 < c; c0 += 64) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {  // rows = channels c0 + ty + 8r, columns = points n0 + tx: 8 loads in flight
      const int ch = c0 + ty + 8 * r;
      t[ty + 8 * r][tx] = (ch < c && n0 + tx < n) ? xb[(size_t)ch * n + n0 + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {  // rows = points, columns = channels
      const int p = ty + 8 * r;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float v = t[tx + 32 * h][p];
        if (n0 + p < n && c0 + tx + 32 * h < c) xo[(size_t)(n0 + p) * c + c0 + tx + 32 * h] = v;
        acc[r] = fmaf(v, v, acc[r]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float s = warp_sum(acc[r]);
    const int p = n0 + ty + 8 * r;
    if (tx == 0 && p < npad) {
      const bool real = p < n;
      const float INF = __int_as_float(0x7f800000);
      norms[cloud * (size_t)npad + p] = real ? s : INF;
      float *bt = bounds + (cloud * (size_t)npad + (size_t)(p / R) * R) * 3 + (p % R);
      bt[0] = real ? s * (1.f + T2_C2) : INF;
      bt[R] = real ? s * (1.f - T2_C2) : INF;
      bt[2 * R] = real ? T2_C1 * sqrtf(s) : 0.f;
    }
  }
  // the main kernel may become resident (barriers, TMEM allocation) while the last CTAs of this one drain; it executes
  // griddepcontrol.wait before it reads anything written here.  Triggered at the END: a main-kernel CTA takes a whole SM's
  // shared memory, so an early trigger would starve the CTAs of this kernel that are still queued.
  pdl_trigger();
}

struct T2Ctl {
  uint64_t full[2], empty[2], tfull[2], tempty[2], afull;
  uint32_t tmem_base;
};

template <int N>
__device__ __forceinline__ void bitonic_sort_regs(float (&a)[N]) {  // ascending, fully unrolled (static indices)
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

template <int N>
__device__ __forceinline__ void bitonic_merge_regs(float (&a)[N]) {  // a is bitonic -> ascending
#pragma unroll
  for (int j = N >> 1; j > 0; j >>= 1)
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const int l = i ^ j;
      if (l > i) {
        const float lo = fminf(a[i], a[l]), hi = fmaxf(a[i], a[l]);
        a[i] = lo;
        a[l] = hi;
      }
    }
}

template <int KB, int HALVES, int R>
__global__ void __launch_bounds__(T2<KB, HALVES, R>::THREADS, 1)
knn_tc2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_r, int n, int k,
               int npad, const float *__restrict__ xT, const float *__restrict__ norms,
               const float *__restrict__ bounds, int64_t *__restrict__ idx_out, float *__restrict__ dist_out) {
  using Cfg = T2<KB, HALVES, R>;
  constexpr int C = Cfg::C, Q = Cfg::QUERIES, CAP = Cfg::CAP, CHUNKS = Cfg::CHUNKS, W = Cfg::W;
  extern __shared__ __align__(1024) unsigned char smem[];  // the 128B swizzle atoms need 1024-byte alignment
  unsigned char *sA = smem;                                        // [KB][Q][128 B]
  unsigned char *sB = sA + Cfg::A_BYTES;                           // [2][KB][R][128 B]
  float *rn = reinterpret_cast<float *>(sB + 2 * Cfg::B_BYTES);    // [2][3][R] per staged key: n(1+c2), n(1-c2), c1 sqrt(n)
  uint32_t *recs = reinterpret_cast<uint32_t *>(rn + 2 * 3 * R);   // [Q][W] per-query records
  uint32_t *plist = recs + Q * W;                                  // [NEPI][PCAP]
  uint32_t *scratch = plist + Cfg::NEPI * Cfg::PCAP;               // [NEPI][64], 16-byte aligned
  T2Ctl *ctl = reinterpret_cast<T2Ctl *>(scratch + Cfg::NEPI * 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cloud = blockIdx.y;
  const int q0 = blockIdx.x * Q;
  const int ntile = npad / R;
  const int niter = 2 * ntile;
  // group minima over 16 instead of 32 keys while the query record holds them (64 values, 128 with the long records)
  const bool fine = ntile * CHUNKS * 2 <= (W >= 128 ? 128 : 64);

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], Cfg::NEPI);
      mbar_init(&ctl->tfull[s], 1);
      mbar_init(&ctl->tempty[s], Cfg::NEPI);
    }
    mbar_init(&ctl->afull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&ctl->tmem_base, Cfg::TMEM_COLS);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  pdl_wait();  // everything above overlapped the prep kernel; its outputs (xT, norms, bounds) are read from here on

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_expect_tx(&ctl->afull, Cfg::A_BYTES);
      for (int kb = 0; kb < KB; ++kb) tma_load_3d(sA + kb * (Q * 128), &tmap_q, &ctl->afull, kb * 32, q0, cloud);
      for (int i = 0; i < niter; ++i) {
        const int s = i & 1, par = (i >> 1) & 1;
        const int r0 = (i % ntile) * R;
        mbar_wait(&ctl->empty[s], par ^ 1);
        mbar_expect_tx(&ctl->full[s], Cfg::B_BYTES + 3 * R * 4);
        unsigned char *st = sB + s * Cfg::B_BYTES;
        for (int kb = 0; kb < KB; ++kb) tma_load_3d(st + kb * (R * 128), &tmap_r, &ctl->full[s], kb * 32, r0, cloud);
        bulk_load_1d(rn + s * 3 * R, bounds + ((size_t)cloud * npad + r0) * 3, 3 * R * 4, &ctl->full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t IDESC = umma_idesc_tf32(128, R);
      mbar_wait(&ctl->afull, 0);
      for (int i = 0; i < niter; ++i) {
        const int s = i & 1, par = (i >> 1) & 1;
        mbar_wait(&ctl->tempty[s], par ^ 1);  // epilogue drained accumulator s
        mbar_wait(&ctl->full[s], par);
        fence_after();
        const uint32_t b_addr = smem_u32(sB + s * Cfg::B_BYTES);
#pragma unroll
        for (int h = 0; h < HALVES; ++h) {
          const uint32_t tmem_d = tmem_base + (uint32_t)((h * 2 + s) * R);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            const uint64_t da = umma_desc_sw128(smem_u32(sA) + kb * (Q * 128) + h * (128 * 128));
            const uint64_t db = umma_desc_sw128(b_addr + kb * (R * 128));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)  // UMMA_K = 8 tf32 = 32 bytes: advance the start address by 2 (x16 B)
              mma_tf32(tmem_d, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), IDESC, (kb | kk) ? 1u : 0u);
          }
        }
        mma_commit(&ctl->tfull[s]);
      }
    }
  } else {
    // ===== epilogue: one thread per query =====
    const int quarter = warp & 3;       // TMEM lanes 32*quarter .. +31 are accessible to this warp
    const int half = (warp - 2) >> 2;
    const int e0 = half * 128 + quarter * 32;  // first query slot of this warp
    const int e = e0 + lane;                   // query slot in the CTA
    const int q = q0 + e;
    const float INF = __int_as_float(0x7f800000);
    const float nq = (q < n) ? norms[(size_t)cloud * npad + q] : 0.f;
    // eps_ij = cq * (c1 sqrt(n_j)) + c2 n_j + c2 |x_q|^2: the key-side factors come precomputed (knn_tc2_prep_kernel),
    // the query side is cq = |x_q| and the constant c2 |x_q|^2, which moves into the threshold
    const float cq = sqrtf(nq), c2nq = T2_C2 * nq;
    const f32x2 m2 = pack2(-2.f, -2.f), cq2 = pack2(cq, cq), ncq2 = pack2(-cq, -cq);
    float thr = INF;
    int cnt = 0;
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 2 * R);
    uint32_t *rec = recs + e * W;                       // this query's record: distances, then packed indices
    uint32_t *myplist = plist + (warp - 2) * Cfg::PCAP;  // this warp's (query, key) pairs of the current tile
    uint32_t *myscr = scratch + (warp - 2) * 64;

    // exact squared distance of query slot eq to key row jl of the staged tile: the canonical sequential fma chain
    auto exact = [&](const unsigned char *st, int eq, int jl) {
      float d = 0.f;
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        const unsigned char *rowk = st + kb * (R * 128) + jl * 128;
        const unsigned char *rowq = sA + kb * (Q * 128) + eq * 128;
#pragma unroll
        for (int c16 = 0; c16 < 8; ++c16) {
          const float4 b4 = *reinterpret_cast<const float4 *>(rowk + ((c16 ^ (jl & 7)) << 4));
          const float4 a4 = *reinterpret_cast<const float4 *>(rowq + ((c16 ^ (eq & 7)) << 4));
          // the four differences as two packed FADD2 (bit-identical to scalar subtraction), then the scalar chain
          float t0, t1, t2, t3;
          unpack2(add2(pack2(a4.x, a4.y), pack2(-b4.x, -b4.y)), t0, t1);
          unpack2(add2(pack2(a4.z, a4.w), pack2(-b4.z, -b4.w)), t2, t3);
          d = fmaf(t0, t0, d);
          d = fmaf(t1, t1, d);
          d = fmaf(t2, t2, d);
          d = fmaf(t3, t3, d);
        }
      }
      return d;
    };
    auto store_cand = [&](int eq, int slot, float d, int j) {
      uint32_t *r = recs + eq * W;
      r[slot] = __float_as_uint(d);
      reinterpret_cast<unsigned short *>(r + CAP)[slot] = (unsigned short)j;
    };

    // keeps the k best (distance, index) entries of this query's list, in order, and tightens the threshold with the
    // k-th best exact distance so far: a key of the final top-k has a true score <= rd[k-1] - |x_q|^2, so its lower bound
    // (which dropped c2 |x_q|^2) is <= rd[k-1] - |x_q|^2 + c2 |x_q|^2
    auto prune_to_k = [&]() {
      float *rd = reinterpret_cast<float *>(rec);
      unsigned short *rix = reinterpret_cast<unsigned short *>(rec + CAP);
      for (int s1 = 1; s1 < cnt; ++s1) {
        const float vd = rd[s1];
        const unsigned short vi = rix[s1];
        int p = s1;
        while (p > 0 && (rd[p - 1] > vd || (rd[p - 1] == vd && rix[p - 1] > vi))) {
          rd[p] = rd[p - 1];
          rix[p] = rix[p - 1];
          --p;
        }
        rd[p] = vd;
        rix[p] = vi;
      }
      cnt = k;
      const float tk = rd[k - 1] - nq + c2nq;
      thr = fminf(thr, tk * (tk > 0.f ? 1.000001f : 0.999999f) + 1e-30f);
    };

    for (int i = 0; i < niter; ++i) {
      const int s = i & 1, par = (i >> 1) & 1;
      const bool second = i >= ntile;
      const int t = second ? i - ntile : i;
      mbar_wait(&ctl->full[s], par);   // acquires the TMA-written key tile and norms for this thread
      mbar_wait(&ctl->tfull[s], par);
      fence_after();
      const uint32_t taddr = tlane + (uint32_t)(s * R);
      const float *rns = rn + s * 3 * R;   // [0] upper norms, [R] lower norms, [2R] c1 sqrt(n)
      if (!second) {
        // ---- sweep 1: group minima ----
#pragma unroll
        for (int ch = 0; ch < CHUNKS; ++ch) {
          uint32_t v[32];
          tmem_ld32_issue(taddr + (uint32_t)(ch * 32), v);
          tmem_ld_wait();
          float mlo = INF, mhi = INF;  // minima of the two groups of 16 keys of this chunk
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 w = *reinterpret_cast<const float4 *>(rns + ch * 32 + 4 * g);
            const float4 sn = *reinterpret_cast<const float4 *>(rns + 2 * R + ch * 32 + 4 * g);
            // upper bound of the true score: s~ + eps_ij (without the per-query constant c2 |x_q|^2)
            // (two scores per packed instruction: same roundings as the scalar fma)
            float z0, z1, z2, z3;
            unpack2(fma2(cq2, pack2(sn.x, sn.y),
                         fma2(m2, pack2(__uint_as_float(v[4 * g + 0]), __uint_as_float(v[4 * g + 1])), pack2(w.x, w.y))), z0, z1);
            unpack2(fma2(cq2, pack2(sn.z, sn.w),
                         fma2(m2, pack2(__uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3])), pack2(w.z, w.w))), z2, z3);
            if (g < 4)
              mlo = fminf(fminf(z0, z1), fminf(fminf(z2, z3), mlo));
            else
              mhi = fminf(fminf(z0, z1), fminf(fminf(z2, z3), mhi));
          }
          if (fine) {  // groups of 16 keys (tighter tau, fewer candidates to re-rank)
            rec[(t * CHUNKS + ch) * 2] = __float_as_uint(mlo);
            rec[(t * CHUNKS + ch) * 2 + 1] = __float_as_uint(mhi);
          } else {
            rec[t * CHUNKS + ch] = __float_as_uint(fminf(mlo, mhi));
          }
        }
        fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&ctl->tempty[s]);
          mbar_arrive(&ctl->empty[s]);
        }
        if (i == ntile - 1) {
          // ---- tau = k-th smallest of the ng <= 128 group minima (at least k groups hold a key with score <= tau) ----
          const int ng = ntile * CHUNKS * (fine ? 2 : 1);
          float a[32];
#pragma unroll
          for (int u = 0; u < 32; ++u) a[u] = (u < ng) ? __uint_as_float(rec[u]) : INF;
          bitonic_sort_regs<32>(a);
          for (int blk = 32; blk < ng; blk += 32) {  // fold in the next 32 minima: keep the 32 smallest of both
            float b2[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) b2[u] = (blk + u < ng) ? __uint_as_float(rec[blk + u]) : INF;
            bitonic_sort_regs<32>(b2);
#pragma unroll
            for (int u = 0; u < 32; ++u) a[u] = fminf(a[u], b2[31 - u]);  // the 32 smallest of both, a bitonic sequence
            bitonic_merge_regs<32>(a);
          }
          float tau = -INF;
#pragma unroll
          for (int u = 0; u < 32; ++u) tau = (u < k) ? fmaxf(tau, a[u]) : tau;
          // lower bound <= upper bound of the k-th: both sides dropped c2 |x_q|^2, hence twice that.  Strict compare
          // below: move the threshold up by a few ulps so that equality is still a candidate
          thr = tau + 2.f * c2nq;
          thr = thr * (thr > 0.f ? 1.000001f : 0.999999f) + 1e-30f;
          mbar_wait(&ctl->afull, 0);  // the query tile is read by the exact re-rank
        }
      } else {
        // ---- sweep 2: candidate masks (sign bit of score - thr) ----
        unsigned int masks[CHUNKS];
        int c_l = 0;
        const f32x2 nthr2 = pack2(-thr, -thr);  // z - thr == z + (-thr) bit for bit
#pragma unroll
        for (int ch = 0; ch < CHUNKS; ++ch) {
          uint32_t v[32];
          tmem_ld32_issue(taddr + (uint32_t)(ch * 32), v);
          tmem_ld_wait();
          unsigned int mk = 0;
#pragma unroll
          for (int g = 7; g >= 0; --g) {
            const float4 w = *reinterpret_cast<const float4 *>(rns + R + ch * 32 + 4 * g);
            const float4 sn = *reinterpret_cast<const float4 *>(rns + 2 * R + ch * 32 + 4 * g);
            // lower bound of the true score, minus the threshold
            float z0, z1, z2, z3;
            unpack2(add2(fma2(ncq2, pack2(sn.x, sn.y),
                              fma2(m2, pack2(__uint_as_float(v[4 * g + 0]), __uint_as_float(v[4 * g + 1])), pack2(w.x, w.y))),
                         nthr2), z0, z1);
            unpack2(add2(fma2(ncq2, pack2(sn.z, sn.w),
                              fma2(m2, pack2(__uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3])), pack2(w.z, w.w))),
                         nthr2), z2, z3);
            mk = __funnelshift_l(__float_as_uint(z3), mk, 1);
            mk = __funnelshift_l(__float_as_uint(z2), mk, 1);
            mk = __funnelshift_l(__float_as_uint(z1), mk, 1);
            mk = __funnelshift_l(__float_as_uint(z0), mk, 1);
          }
          masks[ch] = mk;
          c_l += __popc(mk);
        }
        // a list that would overflow is first pruned to its k best entries (exact distances, so nothing is lost);
        // only a tile with more than CAP - k candidates of one query (massive ties) defeats this
        if (cnt <= CAP && cnt + c_l > CAP && cnt > k) prune_to_k();
        __syncwarp();
        // ---- (query, key) pairs of the whole warp, compacted so that the exact re-rank is spread evenly over the
        //      lanes (the candidates of one query differ a lot from tile to tile) ----
        const unsigned char *st = sB + s * Cfg::B_BYTES;
        int incl = c_l;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) incl = scan_up_add(incl, o);
        const int npairs = __shfl_sync(0xffffffffu, incl, 31);
        if (npairs <= Cfg::PCAP) {
          int pos = incl - c_l, slot = cnt;
#pragma unroll
          for (int ch = 0; ch < CHUNKS; ++ch) {
            unsigned int mk = masks[ch];
            while (mk) {
              const int bit = __ffs(mk) - 1;
              mk &= mk - 1;
              myplist[pos++] = ((unsigned int)lane << 14) | ((unsigned int)min(slot, 127) << 7) | (unsigned int)(ch * 32 + bit);
              ++slot;
            }
          }
          __syncwarp();
          for (int p2 = lane; p2 < npairs; p2 += 32) {
            const unsigned int pr = myplist[p2];
            const int jl = pr & 127, sl = (pr >> 7) & 127, eq = e0 + (int)(pr >> 14);
            const float d = exact(st, eq, jl);
            if (sl < CAP) store_cand(eq, sl, d, t * R + jl);
          }
        } else {  // more pairs than the list holds (massive ties): every thread walks its own candidates
          int slot = cnt;
#pragma unroll
          for (int ch = 0; ch < CHUNKS; ++ch) {
            unsigned int mk = masks[ch];
            while (mk) {
              const int bit = __ffs(mk) - 1;
              mk &= mk - 1;
              const int jl = ch * 32 + bit;
              if (slot < CAP) store_cand(e, slot, exact(st, e, jl), t * R + jl);
              ++slot;
            }
          }
        }
        cnt += c_l;
        fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&ctl->tempty[s]);
          mbar_arrive(&ctl->empty[s]);
        }
      }
    }

    if (CAP > 64 && cnt > 64 && cnt <= CAP) prune_to_k();  // the ranking below holds two candidates per lane
    // ---- ranking, one query of the warp at a time, one candidate per lane.  d >= 0, so the bit patterns order like the
    //      values.  Without exact ties the strict ranks are a permutation (their sum is c(c-1)/2). ----
    __syncwarp();
    for (int lq = 0; lq < 32; ++lq) {
      const int c = __shfl_sync(0xffffffffu, cnt, lq);
      const int qq = q0 + e0 + lq;
      if (qq >= n || c > CAP) continue;  // warp-uniform
      const uint32_t *r = recs + (e0 + lq) * W;
      const unsigned short *ri = reinterpret_cast<const unsigned short *>(r + CAP);
      int64_t *o = idx_out + ((size_t)cloud * n + qq) * k;
      float *od = dist_out ? dist_out + ((size_t)cloud * n + qq) * k : nullptr;
      const unsigned int me0 = lane < c ? r[lane] : 0x7fffffffu, me1 = lane + 32 < c ? r[lane + 32] : 0x7fffffffu;
      myscr[lane] = me0;
      myscr[lane + 32] = me1;
      __syncwarp();
      int rk0 = 0, rk1 = 0;
      const uint4 *sv = reinterpret_cast<const uint4 *>(myscr);
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
        const unsigned int i0 = lane < c ? ri[lane] : 0xffffu, i1 = lane + 32 < c ? ri[lane + 32] : 0xffffu;
        rk0 = rk1 = 0;
        for (int u = 0; u < c; ++u) {
          const unsigned int du = myscr[u], iu = ri[u];
          rk0 += (du < me0 || (du == me0 && iu < i0)) ? 1 : 0;
          rk1 += (du < me1 || (du == me1 && iu < i1)) ? 1 : 0;
        }
      }
      if (lane < c && rk0 < k) {
        o[rk0] = (int64_t)ri[lane];
        if (od) od[rk0] = __uint_as_float(me0);
      }
      if (lane + 32 < c && rk1 < k) {
        o[rk1] = (int64_t)ri[lane + 32];
        if (od) od[rk1] = __uint_as_float(me1);
      }
      for (int t2 = c + lane; t2 < k; t2 += 32) {  // fewer than k candidates only with NaN / inf inputs
        o[t2] = 0;
        if (od) od[t2] = INF;
      }
      __syncwarp();
    }
    if (q < n && cnt > CAP) {
      // pathological ties (e.g. duplicated clouds): exact brute force over all references, sorted insertion into the
      // query's own record
      int64_t *o = idx_out + ((size_t)cloud * n + q) * k;
      float *od = dist_out ? dist_out + ((size_t)cloud * n + q) * k : nullptr;
      float *rd = reinterpret_cast<float *>(rec);
      unsigned short *rix = reinterpret_cast<unsigned short *>(rec + CAP);
      const float4 *xqg = reinterpret_cast<const float4 *>(xT + ((size_t)cloud * n + q) * C);
      int have = 0;
      for (int j = 0; j < n; ++j) {
        const float4 *xr = reinterpret_cast<const float4 *>(xT + ((size_t)cloud * n + j) * C);
        float d = 0.f;
        for (int c4 = 0; c4 < C / 4; ++c4) {
          const float4 aq = xqg[c4], ar = xr[c4];
          const float t0 = aq.x - ar.x, t1 = aq.y - ar.y, t2 = aq.z - ar.z, t3 = aq.w - ar.w;
          d = fmaf(t0, t0, d);
          d = fmaf(t1, t1, d);
          d = fmaf(t2, t2, d);
          d = fmaf(t3, t3, d);
        }
        if (have == k && !(d < rd[k - 1])) continue;
        int p = have < k ? have : k - 1;
        while (p > 0 && d < rd[p - 1]) {
          rd[p] = rd[p - 1];
          rix[p] = rix[p - 1];
          --p;
        }
        rd[p] = d;
        rix[p] = (unsigned short)j;
        if (have < k) ++have;
      }
      for (int t2 = 0; t2 < k; ++t2) {
        o[t2] = t2 < have ? (int64_t)rix[t2] : 0;
        if (od) od[t2] = t2 < have ? rd[t2] : INF;
      }
    }
  }

  fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---- host ----------------------------------------------------------------------------------------------------
template <int KB, int HALVES, int R>
static int launch_tc2(int b, int c, int n, int k, int npad, bool pm, const float *x, float *xT_ws, float *norms, float *bounds,
                      int64_t *idx, float *dist, cudaStream_t st) {
  using Cfg = T2<KB, HALVES, R>;
  static size_t attr[64];
  if (cudaError_t e = smem_optin(knn_tc2_kernel<KB, HALVES, R>, Cfg::SMEM, attr); e != cudaSuccess) return (int)e;
  knn_tc2_prep_kernel<R><<<dim3((npad + 31) / 32, b), 256, 0, st>>>(c, n, pm, x, xT_ws, norms, bounds, npad);
  const float *xT = pm ? x : xT_ws;
  CUtensorMap mq, mr;
  int rc = tc_make_map(&mq, xT, b, n, Cfg::C, Cfg::QUERIES);
  if (rc == 0) rc = tc_make_map(&mr, xT, b, n, Cfg::C, R);
  if (rc != 0) return rc;
  dim3 grid((n + Cfg::QUERIES - 1) / Cfg::QUERIES, b);
  PCC_LAUNCH(PDL_CHAMFER, PCC_K(knn_tc2_kernel<KB, HALVES, R>), grid, Cfg::THREADS, Cfg::SMEM, st, mq, mr, n, k, npad,
             xT, (const float *)norms, (const float *)bounds, idx, dist);
  return (int)cudaGetLastError();
}

// x (b,c,n) channels-first, or (b,n,c) point-major with pm.  Returns PCC_ENOTSUP when the shape is outside this path.
int knn_tc2_launch(int b, int c, int n, int k, bool pm, const float *x, int64_t *idx, float *dist, cudaStream_t st) {
  if (c % 32 != 0 || c < 32 || c > 128 || k > 32 || n < 32 * k || n > 2048 || b > 65535) return PCC_ENOTSUP;
  if (!tc_get_encode()) return PCC_ENOTSUP;
  const int r = (c <= 64) ? 128 : 64;
  const int npad = (n + r - 1) / r * r;
  float *ws = nullptr;
  const size_t nxt = (size_t)b * n * c, nn = (size_t)b * npad;
  cudaError_t e = ws_alloc((void **)&ws, sizeof(float) * (nxt + 4 * nn), st);
  if (e != cudaSuccess) return (int)e;
  float *xT = ws, *norms = ws + nxt, *bounds = norms + nn;
  int rc;
  switch (c / 32) {
    case 1: rc = launch_tc2<1, 2, 128>(b, c, n, k, npad, pm, x, xT, norms, bounds, idx, dist, st); break;
    case 2: rc = launch_tc2<2, 2, 128>(b, c, n, k, npad, pm, x, xT, norms, bounds, idx, dist, st); break;
    case 3: rc = launch_tc2<3, 1, 64>(b, c, n, k, npad, pm, x, xT, norms, bounds, idx, dist, st); break;
    default: rc = launch_tc2<4, 1, 64>(b, c, n, k, npad, pm, x, xT, norms, bounds, idx, dist, st); break;
  }
  cudaFreeAsync(ws, st);
  if (rc == 0) g_launches.fetch_add(2, std::memory_order_relaxed);
  return rc;
}

}  // namespace pcc
