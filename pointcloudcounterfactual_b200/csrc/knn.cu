// k-nearest-neighbour graph construction (DGCNN) and general argKmin for sm_100a -- exact fp32 SIMT path.
//
// Replaces the KeOps reductions behind src/utils/neighbour_ops.py:63-82 (knn / pykeops_knn: argKmin over
// Sum(Square(x_i - x_j))) and the argmin patterns of src/train/metrics_and_losses.py:33,36 and
// src/module/quantize.py:28.  Canonical arithmetic: d(i,j) = fma chain over channels in ascending order of
// (x_i[c] - x_j[c])^2, selection ascending by (distance, index) -- stable, self included.
//
// Structure per CTA (128 queries x all references, tiles of 32 references):
//   phase 1  register-tiled distance tile (4 queries x 8 references per thread, packed FADD2/FFMA2 over
//            reference pairs), channels streamed through shared memory in chunks of 16;
//   phase 2  one thread per query: threshold filter of the 32 new distances against its current k-th best,
//            warp-compacted insertion into a sorted per-query list kept in shared memory ([slot][query] layout,
//            bank-conflict free).
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace pcc {

constexpr int KN_THREADS = 128;  // == queries per CTA
constexpr int KN_TQ = 128;
constexpr int KN_TR = 32;
constexpr int KN_CK = 16;
constexpr int KN_DS = KN_TQ + 4;  // row stride of the distance tile

struct KnnSmem {
  float xq[KN_CK][KN_TQ];
  float xr[KN_CK][KN_TR];
  float D[KN_TR][KN_DS];
  unsigned char cand[KN_TR][KN_TQ];
};

template <bool POINT_MAJOR>
__global__ void __launch_bounds__(KN_THREADS)
knn_kernel(int c, int nq, int nr, int k, const float *__restrict__ qin, const float *__restrict__ rin,
           int64_t *__restrict__ idx_out, float *__restrict__ dist_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  KnnSmem &S = *reinterpret_cast<KnnSmem *>(smem_raw);
  float *Ld = reinterpret_cast<float *>(smem_raw + sizeof(KnnSmem));  // [k][KN_TQ]
  int *Li = reinterpret_cast<int *>(Ld + (size_t)k * KN_TQ);          // [k][KN_TQ]

  const int tid = threadIdx.x;
  const size_t cloud = blockIdx.y;
  const int q0 = blockIdx.x * KN_TQ;
  const float *__restrict__ qb = qin + cloud * (size_t)c * nq;
  const float *__restrict__ rb = rin + cloud * (size_t)c * nr;

  const int ty = tid >> 2, tx = tid & 3;
  float thr = __int_as_float(0x7f800000);
  int have = 0;

  for (int r0 = 0; r0 < nr; r0 += KN_TR) {
    // ---- phase 1: D[r][q] = sum_c (xq - xr)^2 ------------------------------------------------------------
    f32x2 acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int p = 0; p < 4; ++p) acc[a][p] = 0ull;

    for (int c0 = 0; c0 < c; c0 += KN_CK) {
      const int cn = min(KN_CK, c - c0);
      __syncthreads();  // previous chunk / previous tile's phase 2 done with shared memory
      for (int e = tid; e < cn * KN_TQ; e += KN_THREADS) {
        const int ch = e / KN_TQ, i = e % KN_TQ;
        float v = 0.f;
        if (q0 + i < nq)
          v = POINT_MAJOR ? qb[(size_t)(q0 + i) * c + c0 + ch] : qb[(size_t)(c0 + ch) * nq + q0 + i];
        S.xq[ch][i] = v;
      }
      for (int e = tid; e < cn * KN_TR; e += KN_THREADS) {
        const int ch = e / KN_TR, i = e % KN_TR;
        float v = 0.f;
        if (r0 + i < nr)
          v = POINT_MAJOR ? rb[(size_t)(r0 + i) * c + c0 + ch] : rb[(size_t)(c0 + ch) * nr + r0 + i];
        S.xr[ch][i] = v;
      }
      __syncthreads();
#pragma unroll 4
      for (int ch = 0; ch < cn; ++ch) {
        const float4 qv = *reinterpret_cast<const float4 *>(&S.xq[ch][ty * 4]);
        const float4 ra = *reinterpret_cast<const float4 *>(&S.xr[ch][tx * 4]);
        const float4 rb4 = *reinterpret_cast<const float4 *>(&S.xr[ch][16 + tx * 4]);
        const f32x2 rp[4] = {pack2(ra.x, ra.y), pack2(ra.z, ra.w), pack2(rb4.x, rb4.y), pack2(rb4.z, rb4.w)};
        const float qs[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const f32x2 nq2 = pack2(-qs[a], -qs[a]);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const f32x2 d = add2(rp[p], nq2);
            acc[a][p] = fma2(d, d, acc[a][p]);
          }
        }
      }
    }
    // scatter the register tile to D[ref][query]
    {
      float v[4][8];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int p = 0; p < 4; ++p) unpack2(acc[a][p], v[a][2 * p], v[a][2 * p + 1]);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int rr = (e < 4) ? (tx * 4 + e) : (16 + tx * 4 + e - 4);
        *reinterpret_cast<float4 *>(&S.D[rr][ty * 4]) = make_float4(v[0][e], v[1][e], v[2][e], v[3][e]);
      }
    }
    __syncthreads();

    // ---- phase 2: thread `tid` owns query q0+tid --------------------------------------------------------
    int cnt = 0;
    const int rvalid = min(KN_TR, nr - r0);
#pragma unroll 8
    for (int r = 0; r < KN_TR; ++r) {
      const float d = S.D[r][tid];
      if (r < rvalid && d < thr) {
        S.cand[cnt][tid] = (unsigned char)r;
        ++cnt;
      }
    }
    const int mx = __reduce_max_sync(0xffffffffu, cnt);
    for (int i = 0; i < mx; ++i) {
      if (i < cnt) {
        const int r = S.cand[i][tid];
        const float d = S.D[r][tid];
        if (d < thr) {  // thr may have tightened since the filter
          int p = have < k ? have : k - 1;
          while (p > 0 && d < Ld[(p - 1) * KN_TQ + tid]) {
            Ld[p * KN_TQ + tid] = Ld[(p - 1) * KN_TQ + tid];
            Li[p * KN_TQ + tid] = Li[(p - 1) * KN_TQ + tid];
            --p;
          }
          Ld[p * KN_TQ + tid] = d;
          Li[p * KN_TQ + tid] = r0 + r;
          if (have < k) ++have;
          if (have == k) thr = Ld[(k - 1) * KN_TQ + tid];
        }
      }
    }
  }

  // slots that never filled (only possible with NaN / inf distances): keep the output memory-safe
  for (int t = have; t < k; ++t) {
    Ld[t * KN_TQ + tid] = __int_as_float(0x7f800000);
    Li[t * KN_TQ + tid] = 0;
  }
  __syncthreads();
  const int nvalid = min(KN_TQ, nq - q0);
  const size_t obase = (cloud * (size_t)nq + q0) * k;
  for (int e = tid; e < nvalid * k; e += KN_THREADS) {
    const int qq = e / k, t = e - qq * k;
    idx_out[obase + e] = (int64_t)Li[t * KN_TQ + qq];
    if (dist_out) dist_out[obase + e] = Ld[t * KN_TQ + qq];
  }
}

// ---- xyz (C == 3) fast path --------------------------------------------------------------------------------------
// One thread per query, two passes over the reference tile in shared memory (groups of four points: float4 X, Y, Z):
//   pass 1  distances (packed, canonical order fma(dz,dz,fma(dy,dy,dx*dx))) reduced to one minimum per group of G
//           references; the K smallest group minima are kept in a branch-free sorted register list (2 FMNMX per slot).
//           The k-th smallest group minimum tau is an upper bound of the k-th smallest distance (each of those k
//           groups holds at least one reference <= tau).
//   pass 2  distances again; references with d <= tau (about k plus a few) are appended to a per-thread buffer in
//           shared memory in ascending index order, the buffer is stably insertion-sorted by distance => ascending
//           (distance, index), self included.  A full buffer (massive exact ties) is pruned to its best k in place.
constexpr int K3_THREADS = 128;
constexpr int K3_TILE = 2048;  // references per shared-memory tile (24 KiB)
constexpr int K3_CAP = 48;     // candidate slots per query

__device__ __forceinline__ f32x2 knn_sqdist2(f32x2 rx, f32x2 ry, f32x2 rz, f32x2 nqx, f32x2 nqy, f32x2 nqz) {
  const f32x2 dx = add2(rx, nqx), dy = add2(ry, nqy), dz = add2(rz, nqz);
  return fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));  // channel order x, y, z
}

template <int K>
__global__ void __launch_bounds__(K3_THREADS)
knn3_kernel(int n, int k, int groups8, int ps, int cs, const float *__restrict__ x, int64_t *__restrict__ idx_out,
            float *__restrict__ dist_out, const int *__restrict__ only_hard) {
  if (only_hard && !only_hard[blockIdx.y]) return;  // repair pass behind knn3w_kernel: only the flagged clouds
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4 *tile = reinterpret_cast<float4 *>(smem_raw);                                   // K3_TILE/4*3 float4
  float *bufd = reinterpret_cast<float *>(smem_raw + sizeof(float4) * (K3_TILE / 4 * 3));  // [K3_CAP][K3_THREADS]
  int *bufi = reinterpret_cast<int *>(bufd + K3_CAP * K3_THREADS);

  const int tid = threadIdx.x;
  const size_t cloud = blockIdx.y;
  // coordinate c of point i sits at xr[i * ps + c * cs]: channels-first (b,3,n) has ps = 1, cs = n; point-major
  // (b,n,3) -- the layout the reference hands to KeOps, neighbour_ops.py:79 -- has ps = 3, cs = 1
  const float *__restrict__ xr = x + cloud * (size_t)3 * n;
  const int q0 = blockIdx.x * K3_THREADS;
  const int q = min(q0 + tid, n - 1);
  const float INF = __int_as_float(0x7f800000);
  const float qx = xr[(size_t)q * ps], qy = xr[(size_t)q * ps + cs], qz = xr[(size_t)q * ps + 2 * (size_t)cs];
  const f32x2 nqx = pack2(-qx, -qx), nqy = pack2(-qy, -qy), nqz = pack2(-qz, -qz);
  float *tf = reinterpret_cast<float *>(tile);

  float L[K];
#pragma unroll
  for (int i = 0; i < K; ++i) L[i] = INF;
  float tau = INF;
  int cnt = 0;

  for (int pass = 0; pass < 2; ++pass) {
    float gacc = INF;
    int gfill = 0;
    for (int base = 0; base < n; base += K3_TILE) {
      const int tcnt = min(K3_TILE, n - base);
      const int tcnt8 = (tcnt + 7) & ~7;
      if (pass == 0 || n > K3_TILE) {  // a single tile stays resident for pass 2
        __syncthreads();
        for (int i = tid; i < tcnt8; i += K3_THREADS) {
          float vx = INF, vy = INF, vz = INF;  // padding: distance +inf
          if (i < tcnt) {
            vx = xr[(size_t)(base + i) * ps];
            vy = xr[(size_t)(base + i) * ps + cs];
            vz = xr[(size_t)(base + i) * ps + 2 * (size_t)cs];
          }
          const int o = (i >> 2) * 12 + (i & 3);
          tf[o] = vx;
          tf[o + 4] = vy;
          tf[o + 8] = vz;
        }
        __syncthreads();
      }
      for (int g = 0; g < (tcnt8 >> 3); ++g) {
        const float4 X0 = tile[g * 6 + 0], Y0 = tile[g * 6 + 1], Z0 = tile[g * 6 + 2];
        const float4 X1 = tile[g * 6 + 3], Y1 = tile[g * 6 + 4], Z1 = tile[g * 6 + 5];
        float a[8];
        unpack2(knn_sqdist2(pack2(X0.x, X0.y), pack2(Y0.x, Y0.y), pack2(Z0.x, Z0.y), nqx, nqy, nqz), a[0], a[1]);
        unpack2(knn_sqdist2(pack2(X0.z, X0.w), pack2(Y0.z, Y0.w), pack2(Z0.z, Z0.w), nqx, nqy, nqz), a[2], a[3]);
        unpack2(knn_sqdist2(pack2(X1.x, X1.y), pack2(Y1.x, Y1.y), pack2(Z1.x, Z1.y), nqx, nqy, nqz), a[4], a[5]);
        unpack2(knn_sqdist2(pack2(X1.z, X1.w), pack2(Y1.z, Y1.w), pack2(Z1.z, Z1.w), nqx, nqy, nqz), a[6], a[7]);
        if (pass == 0) {
          float mn = fminf(fminf(a[0], a[1]), gacc);
          mn = fminf(fminf(a[2], a[3]), mn);
          mn = fminf(fminf(a[4], a[5]), mn);
          gacc = fminf(fminf(a[6], a[7]), mn);
          if (++gfill == groups8) {  // uniform: a group of G = 8*groups8 references is complete
            float c = gacc;
#pragma unroll
            for (int i = 0; i < K; ++i) {
              const float lo = fminf(L[i], c);
              c = fmaxf(L[i], c);
              L[i] = lo;
            }
            gacc = INF;
            gfill = 0;
          }
        } else {
          const int r0 = base + g * 8;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            if (a[e] <= tau) {
              if (cnt == K3_CAP) {  // rare: prune the buffer to its best k (stable), then accept only d < tau
                for (int s1 = 1; s1 < cnt; ++s1) {
                  const float v = bufd[s1 * K3_THREADS + tid];
                  const int vi = bufi[s1 * K3_THREADS + tid];
                  int pp = s1;
                  while (pp > 0 && bufd[(pp - 1) * K3_THREADS + tid] > v) {
                    bufd[pp * K3_THREADS + tid] = bufd[(pp - 1) * K3_THREADS + tid];
                    bufi[pp * K3_THREADS + tid] = bufi[(pp - 1) * K3_THREADS + tid];
                    --pp;
                  }
                  bufd[pp * K3_THREADS + tid] = v;
                  bufi[pp * K3_THREADS + tid] = vi;
                }
                cnt = k;
                const float kth = bufd[(k - 1) * K3_THREADS + tid];
                tau = (kth > 0.f) ? __int_as_float(__float_as_int(kth) - 1) : -1.f;  // largest float below kth
              }
              if (a[e] <= tau) {
                bufd[cnt * K3_THREADS + tid] = a[e];
                bufi[cnt * K3_THREADS + tid] = r0 + e;
                ++cnt;
              }
            }
          }
        }
      }
    }
    if (pass == 0) {
      if (gfill > 0) {  // last, partial group
        float c = gacc;
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const float lo = fminf(L[i], c);
          c = fmaxf(L[i], c);
          L[i] = lo;
        }
      }
      // tau = L[k-1] = max(L[0..k-1]) (the list is ascending); written as a predicated max so that the list is
      // never indexed dynamically (which would move it to local memory).  +inf when fewer than k groups exist.
      tau = 0.f;
#pragma unroll
      for (int i = 0; i < K; ++i) tau = (i < k) ? fmaxf(tau, L[i]) : tau;
    }
  }

  // stable insertion sort of the candidates by distance (they were appended in ascending index order)
  for (int s1 = 1; s1 < cnt; ++s1) {
    const float v = bufd[s1 * K3_THREADS + tid];
    const int vi = bufi[s1 * K3_THREADS + tid];
    int pp = s1;
    while (pp > 0 && bufd[(pp - 1) * K3_THREADS + tid] > v) {
      bufd[pp * K3_THREADS + tid] = bufd[(pp - 1) * K3_THREADS + tid];
      bufi[pp * K3_THREADS + tid] = bufi[(pp - 1) * K3_THREADS + tid];
      --pp;
    }
    bufd[pp * K3_THREADS + tid] = v;
    bufi[pp * K3_THREADS + tid] = vi;
  }
  for (int t = cnt; t < k; ++t) {  // only with NaN / inf inputs
    bufd[t * K3_THREADS + tid] = INF;
    bufi[t * K3_THREADS + tid] = 0;
  }
  __syncthreads();
  const int nvalid = min(K3_THREADS, n - q0);
  const size_t obase = (cloud * (size_t)n + q0) * k;
  for (int e = tid; e < nvalid * k; e += K3_THREADS) {
    const int qq = e / k, t = e - qq * k;
    idx_out[obase + e] = (int64_t)bufi[t * K3_THREADS + qq];
    if (dist_out) dist_out[obase + e] = bufd[t * K3_THREADS + qq];
  }
}

// ---- xyz (C == 3) warp-cooperative path -----------------------------------------------------------------------------
// The references live in REGISTERS: lane l of a warp holds the 32 points j = 1024*slice + 32*r + l (r = 0..31) of its
// slice, S warps ("slices") form a team that covers a cloud of up to 1024*S points, and the team answers one query
// at a time (query coordinates broadcast from shared memory):
//   1  every lane evaluates its 32 distances (packed FADD2/FMUL2/FFMA2, canonical order) and their minimum;
//   2  the 32*S lane minima are sorted with a shuffle bitonic network (merged across the team through shared
//      memory); tau = their k-th smallest bounds the k-th smallest distance (k lanes hold a point <= tau);
//   3  every lane flags its distances <= tau (they are still in registers: no second evaluation), a warp scan
//      compacts the ~1.5 k candidates of the team into shared memory;
//   4  each candidate's (distance bits, index) key is ranked by counting smaller keys; rank r < k is output slot r.
// No thread ever loops over the references: a query costs ~400 warp instructions instead of ~10^4 thread
// instructions.  A query whose candidate count exceeds KW_CAP (massive exact ties) is answered in place by
// knn3w_overflow: no flag array to clear and no repair launch behind the kernel.
constexpr int KW_THREADS = 128;
constexpr int KW_CAP = 128;  // candidate slots per team

__device__ __forceinline__ void team_sync(int id, int threads) {
  if (threads == 32)
    __syncwarp();
  else
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// A query with more than KW_CAP candidates (massive exact ties): selection by k rounds of "next smallest key" over the
// team's references, key = (distance bits, index) -- d >= 0, so the bit patterns order like the values; NaN and padding
// never qualify.  A round is a lane-local scan of 32 distances (recomputed from the staged cloud with the arithmetic of
// the key pass: the same bits), two warp reductions and, for S > 1, a merge through shared memory: ~350 warp
// instructions per round, slower than the ranked path but without a flag array to clear and a repair launch behind the
// kernel (3 us of the 44 us at B=32, N=1024).  Not inlined: its registers must not weigh on the fast path.
template <int S>
__device__ __noinline__ void knn3w_overflow(const float *xs, const float *ys, const float *zs, int n, int k, int slice,
                                            int lane, int team, float qx, float qy, float qz, int64_t *o, float *od,
                                            int *tot_team, unsigned int *lmin_team) {
  const float INF = __int_as_float(0x7f800000);
  const int tl = slice * 32 + lane;
  unsigned int last_d = 0u, last_j = 0u;
  bool have_last = false;
  for (int i = 0; i < k; ++i) {  // team-uniform trip count
    unsigned int bd = 0xffffffffu, bj = 0xffffffffu;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const int j = slice * 1024 + r * 32 + lane;
      const float dx = xs[j] - qx, dy = ys[j] - qy, dz = zs[j] - qz;
      const unsigned int db = __float_as_uint(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))));
      const unsigned int uj = (unsigned int)j;
      const bool ok = j < n && db <= 0x7f800000u &&                                      // a real, non-NaN distance
                      (!have_last || db > last_d || (db == last_d && uj > last_j)) &&   // beyond the previous winner
                      (db < bd || (db == bd && uj < bj));                               // below the best so far
      bd = ok ? db : bd;
      bj = ok ? uj : bj;
    }
    unsigned int md = __reduce_min_sync(0xffffffffu, bd);
    unsigned int mj = __reduce_min_sync(0xffffffffu, bd == md ? bj : 0xffffffffu);
    if (S > 1) {
      team_sync(1 + team, S * 32);  // previous readers of tot / lmin are done
      if (lane == 0) {
        tot_team[slice] = (int)md;
        lmin_team[slice] = mj;
      }
      team_sync(1 + team, S * 32);
#pragma unroll
      for (int s2 = 0; s2 < S; ++s2) {
        const unsigned int od2 = (unsigned int)tot_team[s2], oj2 = lmin_team[s2];
        if (od2 < md || (od2 == md && oj2 < mj)) {
          md = od2;
          mj = oj2;
        }
      }
    }
    const bool found = md <= 0x7f800000u;
    if (tl == 0) {
      o[i] = found ? (int64_t)mj : 0;  // fewer than k real distances only with NaN / inf inputs
      if (od) od[i] = found ? __uint_as_float(md) : INF;
    }
    last_d = md;
    last_j = mj;
    have_last = true;
    if (!found) {  // team-uniform: nothing left, the remaining slots get the same filler
      for (int t = i + 1 + tl; t < k; t += S * 32) {
        o[t] = 0;
        if (od) od[t] = INF;
      }
      break;
    }
  }
}

#ifndef PCC_KW_MINB
#define PCC_KW_MINB 3
#endif
template <int S>
__global__ void __launch_bounds__(KW_THREADS, PCC_KW_MINB)
knn3w_kernel(int n, int k, int qper, int ps, int cs, const float *__restrict__ x, int64_t *__restrict__ idx_out,
             float *__restrict__ dist_out) {
  constexpr int T = KW_THREADS / 32 / S;  // teams per CTA
  constexpr int NP = 1024 * S;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *xs = reinterpret_cast<float *>(smem_raw), *ys = xs + NP, *zs = ys + NP;
  __shared__ __align__(16) unsigned int dkey[T][KW_CAP];  // candidate distance bits (d >= 0: unsigned order == float order)
  __shared__ unsigned short cand[T][KW_CAP];              // candidate indices
  __shared__ __align__(16) unsigned int lmin[T][S * 32];  // lane minima of the team
  __shared__ int tot[T][S];
  __shared__ int tsum[T];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int team = warp / S, slice = warp % S, tl = slice * 32 + lane;
  const size_t cloud = blockIdx.y;
  const float *__restrict__ xr = x + cloud * (size_t)3 * n;
  const float INF = __int_as_float(0x7f800000);
  for (int i = tid; i < NP; i += KW_THREADS) {  // padding: +inf coordinates => +inf distance
    xs[i] = i < n ? xr[(size_t)i * ps] : INF;  // (ps, cs) = (1, n) channels-first, (3, 1) point-major
    ys[i] = i < n ? xr[(size_t)i * ps + cs] : INF;
    zs[i] = i < n ? xr[(size_t)i * ps + 2 * (size_t)cs] : INF;
  }
  if (tid < T) tsum[tid] = 0;
  __syncthreads();

  f32x2 rx[16], ry[16], rz[16];
  unsigned int valid = 0;
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const int j0 = slice * 1024 + (2 * p) * 32 + lane, j1 = j0 + 32;
    rx[p] = pack2(xs[j0], xs[j1]);
    ry[p] = pack2(ys[j0], ys[j1]);
    rz[p] = pack2(zs[j0], zs[j1]);
    valid |= (j0 < n ? 1u : 0u) << (2 * p) | (j1 < n ? 1u : 0u) << (2 * p + 1);
  }
  const unsigned int lt_mask = (1u << lane) - 1u;

  const int q_begin = blockIdx.x * qper, q_end = min(n, q_begin + qper);
  for (int q = q_begin + team; q < q_end; q += T) {
    const float qx = xs[q], qy = ys[q], qz = zs[q];
    const f32x2 nqx = pack2(-qx, -qx), nqy = pack2(-qy, -qy), nqz = pack2(-qz, -qz);
    float d[32];
#pragma unroll
    for (int p = 0; p < 16; ++p) unpack2(knn_sqdist2(rx[p], ry[p], rz[p], nqx, nqy, nqz), d[2 * p], d[2 * p + 1]);
    // lane minimum as a tree (short dependency chain)
    float m8[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) m8[g] = fminf(fminf(d[4 * g], d[4 * g + 1]), fminf(d[4 * g + 2], d[4 * g + 3]));
    const float v = fminf(fminf(fminf(m8[0], m8[1]), fminf(m8[2], m8[3])), fminf(fminf(m8[4], m8[5]), fminf(m8[6], m8[7])));
    // ---- 2: tau = k-th smallest of the team's 32*S lane minima.  Every lane counts the minima strictly below its
    //         own (all of them broadcast from shared memory: no shuffle chain); the lanes with fewer than k below
    //         them hold the k smallest values, ties included, so the maximum over those lanes is tau ----
    const unsigned int vb = __float_as_uint(v);  // v >= 0 or NaN (0x7fc00000 > +inf): unsigned order == float order
    lmin[team][tl] = vb;
    team_sync(1 + team, S * 32);
    int below = 0;
    {
      const uint4 *lm = reinterpret_cast<const uint4 *>(lmin[team]);
      int b0 = 0, b1 = 0, b2 = 0, b3 = 0;
#pragma unroll
      for (int u = 0; u < 8 * S; ++u) {
        const uint4 w = lm[u];
        b0 += (w.x - vb) >> 31;  // both < 2^31: bit 31 of the wrapped difference is set iff w < vb
        b1 += (w.y - vb) >> 31;
        b2 += (w.z - vb) >> 31;
        b3 += (w.w - vb) >> 31;
      }
      below = (b0 + b1) + (b2 + b3);
    }
    unsigned int taub = __reduce_max_sync(0xffffffffu, below < k ? vb : 0u);
    if (S > 1) {
      if (lane == 0) tot[team][slice] = (int)taub;
      team_sync(1 + team, S * 32);
#pragma unroll
      for (int s2 = 0; s2 < S; ++s2) taub = max(taub, (unsigned int)tot[team][s2]);
      team_sync(1 + team, S * 32);  // tot is reused for the candidate counts below
    }
    // ---- 3: candidates.  d <= tau  <=>  bits(d) + ~bits(tau) is negative as a signed integer; the sign bits are
    //         funnel-shifted into the mask (four independent chains) ----
    const unsigned int ntau = ~taub;
    unsigned int mk[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int r = 7; r >= 0; --r)
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) mk[c4] = __funnelshift_l(__float_as_uint(d[8 * c4 + r]) + ntau, mk[c4], 1);
    unsigned int mask = ((mk[3] << 24) | (mk[2] << 16) | (mk[1] << 8) | mk[0]) & valid;
    const int cnt = __popc(mask);
    // offsets: almost every lane has 0..3 candidates => three ballots instead of a shuffle scan
    int off, total;
    {
      const unsigned int b1 = __ballot_sync(0xffffffffu, cnt >= 1), b2 = __ballot_sync(0xffffffffu, cnt >= 2),
                         b3 = __ballot_sync(0xffffffffu, cnt >= 3), b4 = __ballot_sync(0xffffffffu, cnt >= 4);
      if (b4 == 0) {
        off = __popc(b1 & lt_mask) + __popc(b2 & lt_mask) + __popc(b3 & lt_mask);
        total = __popc(b1) + __popc(b2) + __popc(b3);
      } else {
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) incl = scan_up_add(incl, o);
        total = __shfl_sync(0xffffffffu, incl, 31);
        off = incl - cnt;
      }
    }
    if (S > 1) {
      if (lane == 0) tot[team][slice] = total;
      team_sync(1 + team, S * 32);
      total = 0;
#pragma unroll
      for (int s2 = 0; s2 < S; ++s2) {
        const int t = tot[team][s2];
        if (s2 < slice) off += t;
        total += t;
      }
    }
    if (total > KW_CAP) {  // team-uniform, rare (massive exact ties): answered in place, see knn3w_overflow
      knn3w_overflow<S>(xs, ys, zs, n, k, slice, lane, team, qx, qy, qz, idx_out + (cloud * (size_t)n + q) * k,
                        dist_out ? dist_out + (cloud * (size_t)n + q) * k : nullptr, tot[team], lmin[team]);
      team_sync(1 + team, S * 32);
      continue;
    }
    // keys: the distance is recomputed with the same arithmetic => the same bits (the register holding it cannot be
    // indexed dynamically); bank = lane, conflict free
    while (mask) {
      const int r = __ffs(mask) - 1;
      mask &= mask - 1;
      const int j = slice * 1024 + r * 32 + lane;
      const float dx = xs[j] - qx, dy = ys[j] - qy, dz = zs[j] - qz;
      dkey[team][off] = __float_as_uint(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))));
      cand[team][off] = (unsigned short)j;
      ++off;
    }
    {  // pad the last group of sixteen keys (0x7fffffff is never below a key)
      const int pi = (total & ~15) + tl;
      if (tl < 16 && pi >= total && pi < KW_CAP) dkey[team][pi] = 0x7fffffffu;
    }
    team_sync(1 + team, S * 32);
    // ---- 4: ranks ----
    int64_t *o = idx_out + (cloud * (size_t)n + q) * k;
    float *od = dist_out ? dist_out + (cloud * (size_t)n + q) * k : nullptr;
    constexpr int ROUNDS = KW_CAP / (S * 32);
    int rank[ROUNDS];
    int lsum = 0;
    const int total4 = ((total + 15) >> 4) << 2;  // uint4 groups, a multiple of four
#pragma unroll
    for (int rd = 0; rd < ROUNDS; ++rd) {
      const int t = tl + rd * S * 32;
      rank[rd] = 0;
      if (rd * S * 32 < total) {  // team-uniform
        const unsigned int me = dkey[team][min(t, total - 1)];
        int r0 = 0, r1 = 0, r2 = 0, r3 = 0;  // number of candidates strictly closer than mine
        const uint4 *kv = reinterpret_cast<const uint4 *>(dkey[team]);
#pragma unroll 1
        for (int u = 0; u < total4; u += 4) {
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const uint4 kk = kv[u + w];
            r0 += (kk.x - me) >> 31;
            r1 += (kk.y - me) >> 31;
            r2 += (kk.z - me) >> 31;
            r3 += (kk.w - me) >> 31;
          }
        }
        rank[rd] = (r0 + r1) + (r2 + r3);
        if (t < total) lsum += rank[rd];
      }
    }
    // without exact ties the strict ranks are a permutation of 0..total-1; every tied pair lowers their sum by one
    int ssum = __reduce_add_sync(0xffffffffu, lsum);
    if (S > 1) {
      if (lane == 0) atomicAdd(&tsum[team], ssum);
      team_sync(1 + team, S * 32);
      ssum = tsum[team];
    }
    if (ssum != total * (total - 1) / 2) {  // team-uniform, rare: break ties by index
#pragma unroll
      for (int rd = 0; rd < ROUNDS; ++rd) {
        const int t = tl + rd * S * 32;
        if (t < total) {
          const unsigned int me = dkey[team][t], mi = cand[team][t];
          int r = 0;
          for (int u = 0; u < total; ++u) {
            const unsigned int ku = dkey[team][u];
            r += (ku < me || (ku == me && cand[team][u] < mi)) ? 1 : 0;
          }
          rank[rd] = r;
        }
      }
    }
#pragma unroll
    for (int rd = 0; rd < ROUNDS; ++rd) {
      const int t = tl + rd * S * 32;
      if (t < total && rank[rd] < k) {
        o[rank[rd]] = (int64_t)cand[team][t];
        if (od) od[rank[rd]] = __uint_as_float(dkey[team][t]);
      }
    }
    for (int t = total + tl; t < k; t += S * 32) {  // only with NaN / inf inputs
      o[t] = 0;
      if (od) od[t] = INF;
    }
    if (S > 1 && tl == 0) tsum[team] = 0;
    team_sync(1 + team, S * 32);
  }
}

template <int S>
static int launch_knn3w_s(int b, int n, int k, int parts, bool pm, const float *x, int64_t *idx, float *dist,
                          cudaStream_t st) {
  const size_t smem = sizeof(float) * 3 * 1024 * S;
  static size_t attr[64];  // static + dynamic shared memory exceeds the 48 KiB default at S = 4: opt in at any size
  if (cudaError_t e = smem_optin(knn3w_kernel<S>, smem, attr, 0); e != cudaSuccess) return (int)e;
  const int qper = (n + parts - 1) / parts;
  knn3w_kernel<S><<<dim3((n + qper - 1) / qper, b), KW_THREADS, smem, st>>>(n, k, qper, pm ? 3 : 1, pm ? 1 : n, x, idx, dist);
  return (int)cudaGetLastError();
}

template <int K>
static int launch_knn3_k(int b, int n, int k, bool pm, const float *x, int64_t *idx, float *dist, const int *only_hard,
                         cudaStream_t st) {
  const size_t smem = sizeof(float4) * (K3_TILE / 4 * 3) + (size_t)K3_CAP * K3_THREADS * (sizeof(float) + sizeof(int));
  static size_t attr[64];
  if (cudaError_t e = smem_optin(knn3_kernel<K>, smem, attr); e != cudaSuccess) return (int)e;
  // group size G = 8 * groups8: as large as possible while leaving at least ~3k groups (tight tau, few candidates)
  int groups8 = 4;
  while (groups8 > 1 && (n / (8 * groups8)) < 3 * k) groups8 >>= 1;
  dim3 grid((n + K3_THREADS - 1) / K3_THREADS, b);
  knn3_kernel<K><<<grid, K3_THREADS, smem, st>>>(n, k, groups8, pm ? 3 : 1, pm ? 1 : n, x, idx, dist, only_hard);
  return finish_launch(1);
}

static int launch_knn3_thread(int b, int n, int k, bool pm, const float *x, int64_t *idx, float *dist, const int *only_hard,
                              cudaStream_t st) {
  if (k <= 4) return launch_knn3_k<4>(b, n, k, pm, x, idx, dist, only_hard, st);
  if (k <= 8) return launch_knn3_k<8>(b, n, k, pm, x, idx, dist, only_hard, st);
  if (k <= 16) return launch_knn3_k<16>(b, n, k, pm, x, idx, dist, only_hard, st);
  if (k <= 20) return launch_knn3_k<20>(b, n, k, pm, x, idx, dist, only_hard, st);
  if (k <= 24) return launch_knn3_k<24>(b, n, k, pm, x, idx, dist, only_hard, st);
  return launch_knn3_k<32>(b, n, k, pm, x, idx, dist, only_hard, st);
}

// Number of query ranges per cloud: fill the 148 SMs x 3 resident CTAs evenly, at least 8 queries per team
static int knn3w_parts(int b, int n, int teams) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  static const char *force = getenv("PCC_KNN3_PARTS");  // tuning hook
  if (force && atoi(force) > 0) return atoi(force);
  // score = fill of the last wave of 3 resident CTAs per SM, discounted by the per-CTA set-up (the cloud is staged in
  // shared memory and registers: worth about 4 queries per team)
  const int gmax = max(1, n / (8 * teams));
  const long long slots = 3LL * sms;
  int best = 1;
  double best_score = 0.0;
  for (int g = 1; g <= gmax && g <= 64; ++g) {
    const long long ctas = (long long)b * g;
    const double fill = (double)ctas / (double)(((ctas + slots - 1) / slots) * slots);
    const double qper = (double)n / g;
    const double score = fill * qper / (qper + 4.0 * teams);
    if (score > best_score) {
      best_score = score;
      best = g;
    }
  }
  return best;
}

static int launch_knn3(int b, int n, int k, bool pm, const float *x, int64_t *idx, float *dist, cudaStream_t st) {
  static const bool thread_only = getenv("PCC_KNN3_THREAD") != nullptr;  // test hook: one-thread-per-query kernel only
  if (thread_only || n > 4096) {
    note_route(R_KNN3_THREAD);
    return launch_knn3_thread(b, n, k, pm, x, idx, dist, nullptr, st);
  }
  note_route(R_KNN3W);
  int rc;
  if (n <= 1024) rc = launch_knn3w_s<1>(b, n, k, knn3w_parts(b, n, 4), pm, x, idx, dist, st);
  else if (n <= 2048) rc = launch_knn3w_s<2>(b, n, k, knn3w_parts(b, n, 2), pm, x, idx, dist, st);
  else rc = launch_knn3w_s<4>(b, n, k, knn3w_parts(b, n, 1), pm, x, idx, dist, st);
  if (rc == 0) g_launches.fetch_add(1, std::memory_order_relaxed);
  return rc;
}

// ---- tiny argmin (vector quantisation, src/module/quantize.py:26-28) ---------------------------------------------
// The nearest-codeword search is (B * n_codes) independent problems of ONE query against a 16-entry book of 4-dim
// codes: a CTA per problem (the general kernels) would be 99 % idle.  One thread per query instead; the references of
// a problem are a few hundred bytes read through L1.  Same arithmetic (sequential fma over channels) and the same
// lowest-index tie rule as the general path.
__global__ void __launch_bounds__(256)
argmin_small_kernel(int b, int nq, int nr, int c, const float *__restrict__ q, const float *__restrict__ r,
                    int64_t *__restrict__ idx_out, float *__restrict__ dist_out) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= (size_t)b * nq) return;
  const size_t cloud = t / nq;
  const float *qp = q + t * c;
  const float *rp = r + cloud * (size_t)nr * c;
  float best = __int_as_float(0x7f800000);
  int bi = 0;
  bool any = false;
  for (int j = 0; j < nr; ++j) {
    float d = 0.f;
    for (int ch = 0; ch < c; ++ch) {
      const float df = qp[ch] - rp[(size_t)j * c + ch];
      d = fmaf(df, df, d);
    }
    if (d < best || (!any && d == best)) {  // strict: the lowest index wins ties; +inf distances still select index 0
      best = d;
      bi = j;
      any = true;
    }
  }
  idx_out[t] = bi;
  if (dist_out) dist_out[t] = best;
}

template <bool PM>
static int launch_knn(int b, int c, int nq, int nr, int k, const float *q, const float *r, int64_t *idx,
                      float *dist, cudaStream_t st) {
  if (b < 0 || c <= 0 || nq < 0 || nr < 0 || k <= 0) return PCC_EINVAL;
  if (k > nr) return PCC_EINVAL;  // torch.topk / argKmin cannot return more neighbours than points
  if (PM && k == 1 && nr <= 64 && c <= 16 && nq <= 8 && b > 0 && nq > 0 && getenv("PCC_KNN_SIMT") == nullptr) {
    const size_t total = (size_t)b * nq;  // no 65535-cloud grid limit on this path (B * n_codes problems)
    argmin_small_kernel<<<(unsigned int)((total + 255) / 256), 256, 0, st>>>(b, nq, nr, c, q, r, idx, dist);
    note_route(R_ARGMIN_SMALL);
    return finish_launch(1);
  }
  if (k > PCC_KNN_MAX_K || b > 65535) return PCC_ENOTSUP;
  if (b == 0 || nq == 0) return PCC_OK;
  static const bool force_simt = getenv("PCC_KNN_SIMT") != nullptr;  // test hook: exact SIMT kernels only

  // Self kNN (q == r): channels-first through pcc_knn, point-major through pcc_argkmin -- the latter is what the
  // reference's UNCHANGED pykeops_knn produces (neighbour_ops.py:77-82: x.transpose(2, 1).contiguous(), then one
  // LazyTensor expression on (x, x)), so both layouts reach the same fast kernels.
  const bool self = q == r && nq == nr;
  if (PM && self) note_route(R_PM_SELF);
  if (self && c == 3 && k <= 32) {
    const bool no_tc = getenv("PCC_KNN3_SIMT") != nullptr;  // test hook (read per call): the SIMT xyz kernels only
    if (!no_tc && !force_simt) {
      const int rc = knn3_tc_launch(b, nq, k, PM, q, idx, dist, st);  // tcgen05 candidate filter (256 <= n <= 2048)
      if (rc != PCC_ENOTSUP) {
        note_route(R_KNN3_TC);
        return rc;
      }
    }
    return launch_knn3(b, nq, k, PM, q, idx, dist, st);
  }
  if (self && !force_simt && c % 32 == 0) {
    static const bool tc_v1 = getenv("PCC_KNN_TC1") != nullptr;  // test hook: first-generation tcgen05 kernel only
    int rc = tc_v1 ? PCC_ENOTSUP : knn_bf_launch(b, c, nq, k, PM, q, idx, dist, st);  // indices only, C = 32 / 64
    if (rc != PCC_ENOTSUP) {
      note_route(R_KNN_BF);
      return rc;
    }
    rc = tc_v1 ? PCC_ENOTSUP : knn_tc2_launch(b, c, nq, k, PM, q, idx, dist, st);
    if (rc != PCC_ENOTSUP) {
      note_route(R_KNN_TC2);
      return rc;
    }
    rc = knn_tc_launch(b, c, nq, k, PM, q, idx, dist, st);
    if (rc != PCC_ENOTSUP) {
      note_route(R_KNN_TC1);
      return rc;
    }
  }
  note_route(R_KNN_SIMT);
  const size_t smem = sizeof(KnnSmem) + (size_t)k * KN_TQ * (sizeof(float) + sizeof(int));
  static size_t attr[64];
  if (cudaError_t e = smem_optin(knn_kernel<PM>, smem, attr); e != cudaSuccess) return (int)e;
  dim3 grid((nq + KN_TQ - 1) / KN_TQ, b);
  knn_kernel<PM><<<grid, KN_THREADS, smem, st>>>(c, nq, nr, k, q, r, idx, dist);
  return finish_launch(1);
}

}  // namespace pcc

using namespace pcc;

extern "C" __attribute__((visibility("default"))) int pcc_knn(int b, int c, int n, int k, const float *x, int64_t *idx, float *dist, pcc_stream_t stream) {
  return launch_knn<false>(b, c, n, n, k, x, x, idx, dist, (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int pcc_argkmin(int b, int nq, int nr, int c, int k, const float *q, const float *r, int64_t *idx,
                           float *dist, pcc_stream_t stream) {
  return launch_knn<true>(b, c, nq, nr, k, q, r, idx, dist, (cudaStream_t)stream);
}
