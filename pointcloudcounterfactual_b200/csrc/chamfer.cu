// Chamfer nearest-neighbour distance, forward and backward, for sm_100a.
//
// Replaces external/pytorch_structural_losses/src/nndistance.cu (NmDistanceKernel :2-124 launched twice,
// NmDistanceGradKernel :129-148 launched twice + 2 memsets) with
//   * ONE forward launch covering both directions.  Reference points are staged in shared memory as groups of
//     four (float4 X, Y, Z), distances are evaluated two at a time with packed FADD2/FMUL2/FFMA2 in the
//     reference's rounding order, running minima use 3-input FMNMX, and the argmin is tracked per group of 8
//     reference points and resolved exactly at the end (lowest index wins ties, like the reference's strict '<').
//   * ONE deterministic backward launch (CTA per cloud and direction, counting sort of the nearest-neighbour
//     lists in shared memory, ordered segmented sums) instead of float atomics.
#include <cstdlib>

#include "common.cuh"

namespace pcc {

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
constexpr int NN_THREADS = 128;
constexpr int NN_WARPS = NN_THREADS / 32;
constexpr int NN_Q = 4;        // queries per lane: every reference loaded from shared memory is used 4 times
constexpr int NN_QT = 32 * NN_Q;  // queries per CTA (each warp covers all of them against its share of the references)
constexpr int NN_TILE = 2048;  // reference points per shared-memory tile (24 KiB)

// CTA = 128 queries x all references.  The four warps split every reference tile into quarters; lane l of each warp
// holds queries l, l+32, l+64, l+96, so one broadcast LDS.128 triple (8 references) feeds 32 distance evaluations.
// The running minimum is tracked per group of 8 references (3-input FMNMX); the exact lowest index inside the
// winning group is resolved at the end with the same arithmetic.
__global__ void __launch_bounds__(NN_THREADS, 7)
nn_fwd_kernel(int n, const float *__restrict__ xyz1, int m, const float *__restrict__ xyz2,
              float *__restrict__ dist1, int *__restrict__ idx1, float *__restrict__ dist2,
              int *__restrict__ idx2) {
  __shared__ float4 tile[NN_TILE / 4 * 3];
  __shared__ float mbest[NN_WARPS][NN_QT];
  __shared__ int mgrp[NN_WARPS][NN_QT];
  const int dir = blockIdx.z;
  const int nq = dir ? m : n, nr = dir ? n : m;
  const int q0 = blockIdx.x * NN_QT;
  if (q0 >= nq) return;  // grid.x is sized for max(n, m); uniform per CTA
  const size_t cloud = blockIdx.y;
  const float *__restrict__ qp = (dir ? xyz2 : xyz1) + cloud * (size_t)nq * 3;
  const float *__restrict__ rp = (dir ? xyz1 : xyz2) + cloud * (size_t)nr * 3;
  float *__restrict__ dout = (dir ? dist2 : dist1) + cloud * (size_t)nq;
  int *__restrict__ iout = (dir ? idx2 : idx1) + cloud * (size_t)nq;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  float best[NN_Q];
  int grp[NN_Q];
  float sx[NN_Q], sy[NN_Q], sz[NN_Q];  // negated query coordinates (broadcast into both f32x2 lanes at use)
#pragma unroll
  for (int u = 0; u < NN_Q; ++u) {
    const int j = min(q0 + u * 32 + lane, nq - 1);
    sx[u] = -qp[j * 3 + 0];
    sy[u] = -qp[j * 3 + 1];
    sz[u] = -qp[j * 3 + 2];
    best[u] = __int_as_float(0x7f800000);
    grp[u] = 0;
  }

  float *tf = reinterpret_cast<float *>(tile);
  const bool vec_ok = (reinterpret_cast<uintptr_t>(rp) & 15) == 0;
  for (int base = 0; base < nr; base += NN_TILE) {
    const int cnt = min(NN_TILE, nr - base);
    const int cnt8 = (cnt + 7) & ~7;
    __syncthreads();  // previous tile fully consumed
    const int nvec = vec_ok ? (cnt >> 2) : 0;  // groups of 4 points = 3 aligned float4 loads
    const float4 *rv = reinterpret_cast<const float4 *>(rp + (size_t)base * 3);
    for (int g = threadIdx.x; g < nvec; g += NN_THREADS) {
      const float4 a = rv[g * 3], b = rv[g * 3 + 1], c = rv[g * 3 + 2];  // x0 y0 z0 x1 | y1 z1 x2 y2 | z2 x3 y3 z3
      tile[g * 3 + 0] = make_float4(a.x, a.w, b.z, c.y);
      tile[g * 3 + 1] = make_float4(a.y, b.x, b.w, c.z);
      tile[g * 3 + 2] = make_float4(a.z, b.y, c.x, c.w);
    }
    for (int i = nvec * 4 + threadIdx.x; i < cnt8; i += NN_THREADS) {
      float x, y, z;
      if (i < cnt) {
        const float *p = rp + (size_t)(base + i) * 3;
        x = p[0];
        y = p[1];
        z = p[2];
      } else {
        x = y = z = __int_as_float(0x7f800000);  // padding: distance +inf, never selected
      }
      const int o = (i >> 2) * 12 + (i & 3);
      tf[o] = x;
      tf[o + 4] = y;
      tf[o + 8] = z;
    }
    __syncthreads();
    const int ng = cnt8 >> 3;                       // groups of 8 references in this tile
    const int per = (ng + NN_WARPS - 1) / NN_WARPS;  // contiguous share of this warp
    const int gb = warp * per, ge = min(ng, gb + per);
    const int g0 = base >> 3;
#pragma unroll 1
    for (int g = gb; g < ge; ++g) {
      const float4 X0 = tile[g * 6 + 0], Y0 = tile[g * 6 + 1], Z0 = tile[g * 6 + 2];
      const float4 X1 = tile[g * 6 + 3], Y1 = tile[g * 6 + 4], Z1 = tile[g * 6 + 5];
#pragma unroll
      for (int u = 0; u < NN_Q; ++u) {
        const f32x2 nqx_u = pack2(sx[u], sx[u]), nqy_u = pack2(sy[u], sy[u]), nqz_u = pack2(sz[u], sz[u]);
        f32x2 d01 = sqdist2(pack2(X0.x, X0.y), pack2(Y0.x, Y0.y), pack2(Z0.x, Z0.y), nqx_u, nqy_u, nqz_u);
        f32x2 d23 = sqdist2(pack2(X0.z, X0.w), pack2(Y0.z, Y0.w), pack2(Z0.z, Z0.w), nqx_u, nqy_u, nqz_u);
        f32x2 d45 = sqdist2(pack2(X1.x, X1.y), pack2(Y1.x, Y1.y), pack2(Z1.x, Z1.y), nqx_u, nqy_u, nqz_u);
        f32x2 d67 = sqdist2(pack2(X1.z, X1.w), pack2(Y1.z, Y1.w), pack2(Z1.z, Z1.w), nqx_u, nqy_u, nqz_u);
        float a0, a1, a2, a3, a4, a5, a6, a7;
        unpack2(d01, a0, a1);
        unpack2(d23, a2, a3);
        unpack2(d45, a4, a5);
        unpack2(d67, a6, a7);
        float mn = fminf(fminf(a0, a1), best[u]);
        mn = fminf(fminf(a2, a3), mn);
        mn = fminf(fminf(a4, a5), mn);
        mn = fminf(fminf(a6, a7), mn);
        grp[u] = (mn < best[u]) ? (g0 + g) : grp[u];  // strict: the first group reaching the minimum wins
        best[u] = mn;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < NN_Q; ++u) {
    mbest[warp][u * 32 + lane] = best[u];
    mgrp[warp][u * 32 + lane] = grp[u];
  }
  __syncthreads();

  // thread t finishes query q0 + t: merge the four warps (lowest group on equal minima), then resolve the exact
  // index inside the winning group of 8 (identical arithmetic => identical bits).
  const int j = q0 + (int)threadIdx.x;
  if (j >= nq) return;
  float bd = mbest[0][threadIdx.x];
  int bg = mgrp[0][threadIdx.x];
#pragma unroll
  for (int w = 1; w < NN_WARPS; ++w) {
    const float d = mbest[w][threadIdx.x];
    const int g = mgrp[w][threadIdx.x];
    if (d < bd || (d == bd && g < bg)) {
      bd = d;
      bg = g;
    }
  }
  const float qx = qp[j * 3 + 0], qy = qp[j * 3 + 1], qz = qp[j * 3 + 2];
  int bi = -1;
#pragma unroll
  for (int e = 7; e >= 0; --e) {
    const int r = bg * 8 + e;
    if (r < nr) {
      const float d = sqdist1(qx, qy, qz, rp[(size_t)r * 3], rp[(size_t)r * 3 + 1], rp[(size_t)r * 3 + 2]);
      if (d == bd) bi = r;
    }
  }
  if (bi < 0) {  // nothing compared below +inf (NaN / inf inputs): the reference keeps element 0 (nndistance.cu:26)
    bi = 0;
    bd = sqdist1(qx, qy, qz, rp[0], rp[1], rp[2]);
  }
  dout[j] = bd;
  iout[j] = bi;
}

// ------------------------------------------------------------------------------------------------------------
// forward, symmetric: every unordered pair (i in cloud 1, j in cloud 2) is evaluated ONCE and feeds both directions
// (the squared distance is bit-identical in both: the differences only change sign).  Halves the FP32-pipe work.
//   unit (one warp)   256 rows i (8 consecutive per lane, registers) x 128 columns j (shared memory, broadcast)
//     row side    running minimum per row in registers; which block of 32 columns reached it is noted once per block
//     column side per column the minimum over the lane's 8 rows (3-input FMNMX), REDUX.MIN over the warp, the lowest
//                 lane holding it (ballot): value + lane are the unit's partial result for that column
//   CTA               4 units: the same 256 rows against 4 x 128 consecutive columns
//   partial results   rows: (min, 32-column block) per (column chunk of 512, row); columns: (min, lane) per
//                     (row block, column)
//   nn_sym_finalize_kernel merges the partials (lowest chunk / block / lane wins ties = lowest index) and resolves the
//                 exact index inside the winning 32 columns / the winning lane's 8 rows with the same arithmetic.
// 4096 units at B=32 x 2048 x 2048: 6.9 per SM sub-partition, so the FP32 pipes are evenly loaded.
// ------------------------------------------------------------------------------------------------------------
constexpr int NS_THREADS = 128;
constexpr int NS_RPL = 8;                // rows per lane
constexpr int NS_ROWS = 32 * NS_RPL;     // rows per CTA
constexpr int NS_WCOLS = 128;            // columns per warp
constexpr int NS_CCOLS = 4 * NS_WCOLS;   // columns per CTA
constexpr int NS_SG = 32;                // columns per argmin block of the row side

struct NnPart {
  float v;
  int loc;
};

#ifndef PCC_NS_MINB
#define PCC_NS_MINB 4
#endif
__global__ void __launch_bounds__(NS_THREADS, PCC_NS_MINB)
nn_sym_kernel(int n, const float *__restrict__ xyz1, int m, const float *__restrict__ xyz2, int nrb, int ncc,
              NnPart *__restrict__ rowpart, NnPart *__restrict__ colpart) {
  pdl_enter();
  __shared__ float4 tile[NS_CCOLS / 4 * 3];  // groups of 4 columns: X, Y, Z
  __shared__ __align__(16) float cval[NS_CCOLS];  // per column: minimum over this CTA's rows
  __shared__ __align__(16) int cloc[NS_CCOLS];    //             lowest lane attaining it
  __shared__ float mbest[4][NS_ROWS];
  __shared__ int mgrp[4][NS_ROWS];
  const int rb = blockIdx.x / ncc, cc = blockIdx.x % ncc;
  const size_t cloud = blockIdx.y;
  const float *__restrict__ rp = xyz1 + cloud * (size_t)n * 3;  // rows
  const float *__restrict__ cp = xyz2 + cloud * (size_t)m * 3;  // columns
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float INF = __int_as_float(0x7f800000);

  float best[NS_RPL], prev[NS_RPL], sx[NS_RPL], sy[NS_RPL], sz[NS_RPL];
  int grp[NS_RPL];
#pragma unroll
  for (int u = 0; u < NS_RPL; ++u) {
    const int i = min(rb * NS_ROWS + lane * NS_RPL + u, n - 1);
    sx[u] = -rp[i * 3 + 0];
    sy[u] = -rp[i * 3 + 1];
    sz[u] = -rp[i * 3 + 2];
    best[u] = prev[u] = INF;
    grp[u] = 0;
  }
  // stage this CTA's columns (padding: +inf coordinates => +inf distance, never selected)
  const int cbase = cc * NS_CCOLS;
  const int ccnt = min(NS_CCOLS, m - cbase);
  float *tf = reinterpret_cast<float *>(tile);
  for (int i = threadIdx.x; i < NS_CCOLS; i += NS_THREADS) {
    float x = INF, y = INF, z = INF;
    if (i < ccnt) {
      const float *p = cp + (size_t)(cbase + i) * 3;
      x = p[0];
      y = p[1];
      z = p[2];
    }
    const int o = (i >> 2) * 12 + (i & 3);
    tf[o] = x;
    tf[o + 4] = y;
    tf[o + 8] = z;
  }
  __syncthreads();

  const int g0 = warp * (NS_WCOLS / 8);  // first group of 8 columns of this warp inside the CTA tile
#pragma unroll 1
  for (int g = g0; g < g0 + NS_WCOLS / 8; ++g) {
    const float4 X0 = tile[g * 6 + 0], Y0 = tile[g * 6 + 1], Z0 = tile[g * 6 + 2];
    const float4 X1 = tile[g * 6 + 3], Y1 = tile[g * 6 + 4], Z1 = tile[g * 6 + 5];
    float cm[8];
#pragma unroll
    for (int up = 0; up < NS_RPL / 2; ++up) {
      float a[2][8];
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const int u = 2 * up + w;
        const f32x2 nqx = pack2(sx[u], sx[u]), nqy = pack2(sy[u], sy[u]), nqz = pack2(sz[u], sz[u]);
        unpack2(sqdist2(pack2(X0.x, X0.y), pack2(Y0.x, Y0.y), pack2(Z0.x, Z0.y), nqx, nqy, nqz), a[w][0], a[w][1]);
        unpack2(sqdist2(pack2(X0.z, X0.w), pack2(Y0.z, Y0.w), pack2(Z0.z, Z0.w), nqx, nqy, nqz), a[w][2], a[w][3]);
        unpack2(sqdist2(pack2(X1.x, X1.y), pack2(Y1.x, Y1.y), pack2(Z1.x, Z1.y), nqx, nqy, nqz), a[w][4], a[w][5]);
        unpack2(sqdist2(pack2(X1.z, X1.w), pack2(Y1.z, Y1.w), pack2(Z1.z, Z1.w), nqx, nqy, nqz), a[w][6], a[w][7]);
        float mn = fminf(fminf(a[w][0], a[w][1]), best[u]);
        mn = fminf(fminf(a[w][2], a[w][3]), mn);
        mn = fminf(fminf(a[w][4], a[w][5]), mn);
        best[u] = fminf(fminf(a[w][6], a[w][7]), mn);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) cm[e] = up ? fminf(fminf(a[0][e], a[1][e]), cm[e]) : fminf(a[0][e], a[1][e]);
    }
    if ((g & (NS_SG / 8 - 1)) == NS_SG / 8 - 1) {  // uniform: a block of 32 columns is complete
#pragma unroll
      for (int u = 0; u < NS_RPL; ++u) {
        grp[u] = (best[u] < prev[u]) ? g : grp[u];  // strict: the first block reaching the minimum wins
        prev[u] = best[u];
      }
    }
    // column minima over the warp's 256 rows: d >= 0 (or NaN, dropped by fminf), so unsigned order == float order
    float mv[8];
    int ml[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const unsigned int bits = __float_as_uint(cm[e]);
      const unsigned int mn = __reduce_min_sync(0xffffffffu, bits);
      mv[e] = __uint_as_float(mn);
      ml[e] = (int)__ballot_sync(0xffffffffu, bits == mn);
    }
    if (lane == 0) {
      *reinterpret_cast<float4 *>(&cval[g * 8]) = make_float4(mv[0], mv[1], mv[2], mv[3]);
      *reinterpret_cast<float4 *>(&cval[g * 8 + 4]) = make_float4(mv[4], mv[5], mv[6], mv[7]);
      *reinterpret_cast<int4 *>(&cloc[g * 8]) = make_int4(ml[0], ml[1], ml[2], ml[3]);  // lane masks, decoded below
      *reinterpret_cast<int4 *>(&cloc[g * 8 + 4]) = make_int4(ml[4], ml[5], ml[6], ml[7]);
    }
  }
#pragma unroll
  for (int u = 0; u < NS_RPL; ++u) {
    mbest[warp][lane * NS_RPL + u] = best[u];
    mgrp[warp][lane * NS_RPL + u] = grp[u];
  }
  __syncthreads();
  // column partials of this row block
  NnPart *cpart = colpart + (cloud * (size_t)nrb + rb) * m;
  for (int i = threadIdx.x; i < ccnt; i += NS_THREADS) cpart[cbase + i] = NnPart{cval[i], __ffs(cloc[i]) - 1};
  // row partials of this column chunk: merge the four warps (they cover ascending column ranges)
  for (int r = threadIdx.x; r < NS_ROWS; r += NS_THREADS) {
    const int i = rb * NS_ROWS + r;
    if (i >= n) break;
    float bd = mbest[0][r];
    int bg = mgrp[0][r];
#pragma unroll
    for (int w = 1; w < 4; ++w) {
      const float d = mbest[w][r];
      if (d < bd) {
        bd = d;
        bg = mgrp[w][r];
      }
    }
    rowpart[(cloud * (size_t)ncc + cc) * n + i] = NnPart{bd, (cc * (NS_CCOLS / 8) + bg) / (NS_SG / 8)};
  }
}

// eight lanes per point of either cloud (blockIdx.z = 0: rows, cloud 1 -> nearest in cloud 2; 1: columns): the lanes
// split the partial results, agree on the winner with three shuffles, then split the exact index search.
__global__ void __launch_bounds__(256)
nn_sym_finalize_kernel(int n, const float *__restrict__ xyz1, int m, const float *__restrict__ xyz2, int nrb, int ncc,
                       const NnPart *__restrict__ rowpart, const NnPart *__restrict__ colpart,
                       float *__restrict__ dist1, int *__restrict__ idx1, float *__restrict__ dist2,
                       int *__restrict__ idx2) {
  pdl_enter();
  const size_t cloud = blockIdx.y;
  const bool cols = blockIdx.z != 0;
  const int sub = threadIdx.x & 7;
  const int npts = cols ? m : n, nother = cols ? n : m, nparts = cols ? nrb : ncc;
  const int t = min((int)((blockIdx.x * blockDim.x + threadIdx.x) >> 3), npts - 1);  // clamped: all lanes stay in the shuffles
  const bool writer = sub == 0 && (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 3) < npts;
  const float *__restrict__ pq = (cols ? xyz2 : xyz1) + cloud * (size_t)npts * 3;    // the point itself
  const float *__restrict__ pr = (cols ? xyz1 : xyz2) + cloud * (size_t)nother * 3;  // the cloud it is matched against
  const NnPart *part = (cols ? colpart : rowpart) + cloud * (size_t)nparts * npts + t;
  const int BIG = 0x7fffffff;
  float bv = __int_as_float(0x7f800000);
  int bl = 0, bp = BIG;
  for (int p = sub; p < nparts; p += 8) {
    const NnPart q = part[(size_t)p * npts];
    if (q.v < bv || bp == BIG) {  // ascending p per lane: strict keeps the lowest part on ties
      bv = q.v;
      bl = q.loc;
      bp = p;
    }
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int ol = __shfl_xor_sync(0xffffffffu, bl, o), op = __shfl_xor_sync(0xffffffffu, bp, o);
    if (ov < bv || (ov == bv && op < bp)) {
      bv = ov;
      bl = ol;
      bp = op;
    }
  }
  // d = fma(dz,dz,fma(dx,dx,dy*dy)) with d* = ref - query: the sign of the differences flips between the two
  // directions, the squares do not => the same bits as the value found by nn_sym_kernel
  const float qx = pq[t * 3], qy = pq[t * 3 + 1], qz = pq[t * 3 + 2];
  int bi = BIG;
  if (!cols) {  // the winning block of 32 columns: 4 per lane
#pragma unroll
    for (int e = 3; e >= 0; --e) {
      const int r = bl * NS_SG + sub * 4 + e;
      if (r < nother && sqdist1(qx, qy, qz, pr[(size_t)r * 3], pr[(size_t)r * 3 + 1], pr[(size_t)r * 3 + 2]) == bv) bi = r;
    }
  } else {  // the winning lane's 8 rows: one per lane
    const int r = bp * NS_ROWS + bl * NS_RPL + sub;
    if (bp != BIG && r < nother &&
        sqdist1(qx, qy, qz, pr[(size_t)r * 3], pr[(size_t)r * 3 + 1], pr[(size_t)r * 3 + 2]) == bv)
      bi = r;
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) bi = min(bi, __shfl_xor_sync(0xffffffffu, bi, o));
  if (writer) {
    if (bi == BIG) {  // nothing compared below +inf (NaN / inf inputs): the reference keeps element 0 (nndistance.cu:26)
      bi = 0;
      bv = sqdist1(qx, qy, qz, pr[0], pr[1], pr[2]);
    }
    (cols ? dist2 : dist1)[cloud * (size_t)npts + t] = bv;
    (cols ? idx2 : idx1)[cloud * (size_t)npts + t] = bi;
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward (deterministic)
// grad_T[j] = 2 g_T[j] (x_T[j] - x_S[idx_T[j]])  -  sum_{l : idx_S[l] == j} 2 g_S[l] (x_S[l] - x_T[j])
// (nndistance.cu:139-145, both launches of :152-153 folded into one expression per target point)
// ------------------------------------------------------------------------------------------------------------
constexpr int NG_THREADS = 512;
constexpr int NG_HEAVY = 32;  // in-degree above which a target is reduced cooperatively by the whole CTA

__device__ __forceinline__ int block_exclusive_scan(int v, int *warp_tot, int *total) {
  // exclusive scan of one int per thread over NG_THREADS threads
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    int t = lane < NG_THREADS / 32 ? warp_tot[lane] : 0;
    int s = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += u;
    }
    if (lane < NG_THREADS / 32) warp_tot[lane] = s - t;  // exclusive warp offsets
    if (lane == 31) *total = s;
  }
  __syncthreads();
  return warp_tot[w] + inc - v;
}

// Upstream gradients: per point (gd1, gd2), or -- for the fused reduced loss -- per cloud: gcloud[b] * s1 for every
// point of cloud 1 and gcloud[b] * s2 for every point of cloud 2 (gd1 == gd2 == nullptr).
__global__ void __launch_bounds__(NG_THREADS)
nn_grad_kernel(int n, const float *__restrict__ xyz1, int m, const float *__restrict__ xyz2,
               const float *__restrict__ gd1, const int *__restrict__ idx1, const float *__restrict__ gd2,
               const int *__restrict__ idx2, float *__restrict__ g1, float *__restrict__ g2,
               const float *__restrict__ gcloud, float s1, float s2) {
  extern __shared__ int sm[];
  __shared__ int warp_tot[32];
  __shared__ int scan_total;
  __shared__ int n_heavy;
  __shared__ float red[3][NG_THREADS / 32];

  const int dir = blockIdx.y;
  const size_t cloud = blockIdx.x;
  const int nT = dir ? m : n, nS = dir ? n : m;
  const float *__restrict__ xT = (dir ? xyz2 : xyz1) + cloud * (size_t)nT * 3;
  const float *__restrict__ xS = (dir ? xyz1 : xyz2) + cloud * (size_t)nS * 3;
  const bool per_cloud = gcloud != nullptr;
  const float *__restrict__ gT = per_cloud ? nullptr : (dir ? gd2 : gd1) + cloud * (size_t)nT;
  const float *__restrict__ gS = per_cloud ? nullptr : (dir ? gd1 : gd2) + cloud * (size_t)nS;
  const float gcT = per_cloud ? gcloud[cloud] * (dir ? s2 : s1) : 0.f, gcS = per_cloud ? gcloud[cloud] * (dir ? s1 : s2) : 0.f;
  const int *__restrict__ iT = (dir ? idx2 : idx1) + cloud * (size_t)nT;  // target -> nearest source
  const int *__restrict__ iS = (dir ? idx1 : idx2) + cloud * (size_t)nS;  // source -> nearest target
  float *__restrict__ out = (dir ? g2 : g1) + cloud * (size_t)nT * 3;

  // blockIdx.z splits the TARGETS of this (cloud, direction) into contiguous ranges, one CTA each: every CTA scans all
  // sources but keeps only those whose nearest target lies in its range (64 CTAs of serial phases are latency-bound)
  const int per = (nT + gridDim.z - 1) / gridDim.z;
  const int jb = min((int)blockIdx.z * per, nT), je = min(jb + per, nT), nR = je - jb;
  int *cursor = sm;             // nR
  int *start = sm + per;        // nR
  int *slots = sm + 2 * per;    // up to nS
  int *heavy = slots + nS;      // up to nS / (NG_HEAVY + 1) + 1 entries

  for (int j = threadIdx.x; j < nR; j += NG_THREADS) cursor[j] = 0;
  if (threadIdx.x == 0) n_heavy = 0;
  __syncthreads();
  for (int l = threadIdx.x; l < nS; l += NG_THREADS) {
    int t = min(max(iS[l], 0), nT - 1);
    if (t >= jb && t < je) atomicAdd(&cursor[t - jb], 1);
  }
  __syncthreads();
  // exclusive scan of the in-degrees: contiguous chunk per thread
  const int chunk = (nR + NG_THREADS - 1) / NG_THREADS;
  const int c0 = min((int)threadIdx.x * chunk, nR), c1 = min(c0 + chunk, nR);
  int local = 0;
  for (int j = c0; j < c1; ++j) local += cursor[j];
  int off = block_exclusive_scan(local, warp_tot, &scan_total);
  for (int j = c0; j < c1; ++j) {
    int d = cursor[j];
    start[j] = off;
    cursor[j] = off;
    off += d;
  }
  __syncthreads();
  for (int l = threadIdx.x; l < nS; l += NG_THREADS) {
    int t = min(max(iS[l], 0), nT - 1);
    if (t >= jb && t < je) slots[atomicAdd(&cursor[t - jb], 1)] = l;
  }
  __syncthreads();

  for (int jr = threadIdx.x; jr < nR; jr += NG_THREADS) {
    const int j = jb + jr;
    const int s = start[jr], e = cursor[jr], deg = e - s;
    if (deg > NG_HEAVY) {
      heavy[atomicAdd(&n_heavy, 1)] = j;
      continue;
    }
    // insertion sort of this target's (short) source list => fixed summation order
    for (int a = s + 1; a < e; ++a) {
      int v = slots[a], p = a;
      while (p > s && slots[p - 1] > v) {
        slots[p] = slots[p - 1];
        --p;
      }
      slots[p] = v;
    }
    const float tx = xT[j * 3], ty = xT[j * 3 + 1], tz = xT[j * 3 + 2];
    const int k = min(max(iT[j], 0), nS - 1);
    const float g = (per_cloud ? gcT : gT[j]) * 2.f;
    float ax = g * (tx - xS[k * 3]), ay = g * (ty - xS[k * 3 + 1]), az = g * (tz - xS[k * 3 + 2]);
    for (int a = s; a < e; ++a) {
      const int l = slots[a];
      const float gl = (per_cloud ? gcS : gS[l]) * 2.f;
      ax += -(gl * (xS[l * 3] - tx));
      ay += -(gl * (xS[l * 3 + 1] - ty));
      az += -(gl * (xS[l * 3 + 2] - tz));
    }
    out[j * 3] = ax;
    out[j * 3 + 1] = ay;
    out[j * 3 + 2] = az;
  }
  __syncthreads();

  // Targets with a long list (collapsed clouds early in training): whole-CTA ordered reduction.
  const int nh = n_heavy;
  for (int h = 0; h < nh; ++h) {
    const int j = heavy[h];
    const float tx = xT[j * 3], ty = xT[j * 3 + 1], tz = xT[j * 3 + 2];
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int l = threadIdx.x; l < nS; l += NG_THREADS) {
      if (min(max(iS[l], 0), nT - 1) == j) {
        const float gl = (per_cloud ? gcS : gS[l]) * 2.f;
        ax += -(gl * (xS[l * 3] - tx));
        ay += -(gl * (xS[l * 3 + 1] - ty));
        az += -(gl * (xS[l * 3 + 2] - tz));
      }
    }
    ax = warp_sum(ax);
    ay = warp_sum(ay);
    az = warp_sum(az);
    if ((threadIdx.x & 31) == 0) {
      red[0][threadIdx.x >> 5] = ax;
      red[1][threadIdx.x >> 5] = ay;
      red[2][threadIdx.x >> 5] = az;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const int k = min(max(iT[j], 0), nS - 1);
      const float g = (per_cloud ? gcT : gT[j]) * 2.f;
      float sx = g * (tx - xS[k * 3]), sy = g * (ty - xS[k * 3 + 1]), sz = g * (tz - xS[k * 3 + 2]);
      for (int w = 0; w < NG_THREADS / 32; ++w) {
        sx += red[0][w];
        sy += red[1][w];
        sz += red[2][w];
      }
      out[j * 3] = sx;
      out[j * 3 + 1] = sy;
      out[j * 3 + 2] = sz;
    }
    __syncthreads();
  }
}

// target ranges per (cloud, direction): about two CTAs per SM in total, at least 256 targets per range
static int nn_grad_parts(int b, int n_small) {
  int parts = 300 / (2 * (b > 0 ? b : 1));
  parts = parts < 1 ? 1 : (parts > 8 ? 8 : parts);
  while (parts > 1 && n_small / parts < 256) --parts;
  return parts;
}

// loss[b] = s1 * sum_j dist1[b][j] + s2 * sum_k dist2[b][k]: fixed summation order (deterministic)
__global__ void __launch_bounds__(256)
nn_reduce_kernel(int n, int m, const float *__restrict__ dist1, const float *__restrict__ dist2, float s1, float s2,
                 float *__restrict__ loss) {
  pdl_enter();
  __shared__ float red[2][8];
  const size_t cloud = blockIdx.x;
  float a = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) a += dist1[cloud * (size_t)n + i];
  for (int i = threadIdx.x; i < m; i += 256) c += dist2[cloud * (size_t)m + i];
  a = warp_sum(a);
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = a;
    red[1][threadIdx.x >> 5] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float sa = 0.f, sc = 0.f;
    for (int w = 0; w < 8; ++w) {
      sa += red[0][w];
      sc += red[1][w];
    }
    loss[cloud] = s2 * sc + s1 * sa;
  }
}

// Fallback for clouds whose lists do not fit in shared memory: own terms by plain stores, scatter terms by float
// atomics (same scheme as the reference, not bitwise deterministic).
__global__ void nn_grad_own_kernel(int b, int n, const float *__restrict__ xyz1, int m,
                                   const float *__restrict__ xyz2, const float *__restrict__ gd1,
                                   const int *__restrict__ idx1, float *__restrict__ g1) {
  const size_t total = (size_t)b * n;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t i = t / n;
    const int k = min(max(idx1[t], 0), m - 1);
    const float g = gd1[t] * 2.f;
    const float *p = xyz1 + t * 3, *q = xyz2 + (i * m + k) * 3;
    g1[t * 3] = g * (p[0] - q[0]);
    g1[t * 3 + 1] = g * (p[1] - q[1]);
    g1[t * 3 + 2] = g * (p[2] - q[2]);
  }
}
__global__ void nn_grad_scatter_kernel(int b, int n, const float *__restrict__ xyz1, int m,
                                       const float *__restrict__ xyz2, const float *__restrict__ gd1,
                                       const int *__restrict__ idx1, float *__restrict__ g2) {
  const size_t total = (size_t)b * n;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t i = t / n;
    const int k = min(max(idx1[t], 0), m - 1);
    const float g = gd1[t] * 2.f;
    const float *p = xyz1 + t * 3, *q = xyz2 + (i * m + k) * 3;
    float *o = g2 + (i * m + k) * 3;
    atomicAdd(o, -(g * (p[0] - q[0])));
    atomicAdd(o + 1, -(g * (p[1] - q[1])));
    atomicAdd(o + 2, -(g * (p[2] - q[2])));
  }
}

}  // namespace pcc

using namespace pcc;

// `keep`: the caller launches a dependent kernel right behind the finalize kernel (programmatic dependent launch
// needs kernel -> kernel edges) and releases the scratch buffer itself.
static int nn_forward(int b, int n, const float *xyz, int m, const float *xyz2, float *result, int *result_i,
                      float *result2, int *result2_i, pcc_stream_t stream, NnPart **keep) {
  if (keep) *keep = nullptr;
  if (b < 0 || n < 0 || m < 0) return PCC_EINVAL;
  if (b == 0 || n == 0 || m == 0) return PCC_OK;  // nothing to compare against: outputs are left untouched
  if (b > 65535) return PCC_ENOTSUP;
  cudaStream_t st = (cudaStream_t)stream;
  static const bool asym = getenv("PCC_NN_ASYM") != nullptr;  // test hook: one launch per direction pair, no sharing
  if (asym) {
    const int mx = n > m ? n : m;
    dim3 grid((mx + NN_QT - 1) / NN_QT, b, 2);
    nn_fwd_kernel<<<grid, NN_THREADS, 0, st>>>(n, xyz, m, xyz2, result, result_i, result2, result2_i);
    note_route(R_NN_ASYM);
    return finish_launch(1);
  }
  const int nrb = (n + NS_ROWS - 1) / NS_ROWS, ncc = (m + NS_CCOLS - 1) / NS_CCOLS;
  if ((long long)nrb * ncc > 2147483647LL) return PCC_ENOTSUP;
  NnPart *scratch = nullptr;
  const size_t nrow = (size_t)b * ncc * n, ncol = (size_t)b * nrb * m;
  cudaError_t e = ws_alloc((void **)&scratch, sizeof(NnPart) * (nrow + ncol), st);
  if (e != cudaSuccess) return (int)e;
  NnPart *rowpart = scratch, *colpart = scratch + nrow;
  note_route(R_NN_SYM);
  nn_sym_kernel<<<dim3(nrb * ncc, b), NS_THREADS, 0, st>>>(n, xyz, m, xyz2, nrb, ncc, rowpart, colpart);
  const int mx = n > m ? n : m;
  PCC_LAUNCH(PDL_CHAMFER, nn_sym_finalize_kernel, dim3((mx + 31) / 32, b, 2), 256, 0, st, n, xyz, m, xyz2, nrb, ncc,
             (const NnPart *)rowpart, (const NnPart *)colpart, result, result_i, result2, result2_i);
  if (keep)
    *keep = scratch;
  else
    cudaFreeAsync(scratch, st);
  return finish_launch(2);
}

extern "C" __attribute__((visibility("default"))) int pcc_nndistance(int b, int n, const float *xyz, int m, const float *xyz2, float *result,
                              int *result_i, float *result2, int *result2_i, pcc_stream_t stream) {
  return nn_forward(b, n, xyz, m, xyz2, result, result_i, result2, result2_i, stream, nullptr);
}

extern "C" __attribute__((visibility("default"))) int pcc_nndistancegrad(int b, int n, const float *xyz1, int m, const float *xyz2,
                                  const float *grad_dist1, const int *idx1, const float *grad_dist2,
                                  const int *idx2, float *grad_xyz1, float *grad_xyz2, pcc_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return PCC_EINVAL;
  if (b == 0 || (n == 0 && m == 0)) return PCC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0 || m == 0) {  // no neighbours exist: gradients are zero
    if (n) cudaMemsetAsync(grad_xyz1, 0, sizeof(float) * (size_t)b * n * 3, st);
    if (m) cudaMemsetAsync(grad_xyz2, 0, sizeof(float) * (size_t)b * m * 3, st);
    return (int)cudaGetLastError();
  }
  const size_t mx = n > m ? n : m, mn = n > m ? m : n;
  const size_t smem = sizeof(int) * (2 * mx + mx + mx / (NG_HEAVY + 1) + 8);
  if (smem <= 200 * 1024 && b <= 65535) {
    static size_t attr[64];
    if (cudaError_t e = smem_optin(nn_grad_kernel, 200 * 1024, attr); e != cudaSuccess) return (int)e;
    nn_grad_kernel<<<dim3(b, 2, nn_grad_parts(b, (int)mn)), NG_THREADS, smem, st>>>(
        n, xyz1, m, xyz2, grad_dist1, idx1, grad_dist2, idx2, grad_xyz1, grad_xyz2, nullptr, 0.f, 0.f);
    return finish_launch(1);
  }
  const int threads = 256;
  const int g1 = (int)(((size_t)b * n + threads - 1) / threads), g2 = (int)(((size_t)b * m + threads - 1) / threads);
  nn_grad_own_kernel<<<g1 < 65535 * 16 ? g1 : 65535 * 16, threads, 0, st>>>(b, n, xyz1, m, xyz2, grad_dist1, idx1, grad_xyz1);
  nn_grad_own_kernel<<<g2 < 65535 * 16 ? g2 : 65535 * 16, threads, 0, st>>>(b, m, xyz2, n, xyz1, grad_dist2, idx2, grad_xyz2);
  nn_grad_scatter_kernel<<<g1 < 65535 * 16 ? g1 : 65535 * 16, threads, 0, st>>>(b, n, xyz1, m, xyz2, grad_dist1, idx1, grad_xyz2);
  nn_grad_scatter_kernel<<<g2 < 65535 * 16 ? g2 : 65535 * 16, threads, 0, st>>>(b, m, xyz2, n, xyz1, grad_dist2, idx2, grad_xyz1);
  return finish_launch(4);
}

extern "C" __attribute__((visibility("default"))) int pcc_chamfer_reduce(int b, int n, const float *xyz1, int m, const float *xyz2, float scale1,
                                  float scale2, float *loss, float *dist1, int *idx1, float *dist2, int *idx2,
                                  pcc_stream_t stream) {
  if (b < 0 || n <= 0 || m <= 0) return PCC_EINVAL;
  if (b == 0) return PCC_OK;
  NnPart *scratch = nullptr;
  int rc = nn_forward(b, n, xyz1, m, xyz2, dist1, idx1, dist2, idx2, stream, &scratch);
  if (rc == 0) {
    PCC_LAUNCH(PDL_CHAMFER, nn_reduce_kernel, b, 256, 0, (cudaStream_t)stream, n, m, (const float *)dist1,
               (const float *)dist2, scale1, scale2, loss);
    rc = finish_launch(1);
  }
  if (scratch) cudaFreeAsync(scratch, (cudaStream_t)stream);
  return rc;
}

extern "C" __attribute__((visibility("default"))) int pcc_chamfer_reduce_grad(int b, int n, const float *xyz1, int m, const float *xyz2,
                                       const int *idx1, const int *idx2, const float *grad_loss, float scale1,
                                       float scale2, float *grad_xyz1, float *grad_xyz2, pcc_stream_t stream) {
  if (b < 0 || n <= 0 || m <= 0) return PCC_EINVAL;
  if (b == 0) return PCC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t mx = n > m ? n : m;
  const size_t smem = sizeof(int) * (2 * mx + mx + mx / (NG_HEAVY + 1) + 8);
  if (smem > 200 * 1024 || b > 65535) return PCC_ENOTSUP;  // caller falls back to per-point upstream gradients
  static size_t attr[64];
  if (cudaError_t e = smem_optin(nn_grad_kernel, 200 * 1024, attr); e != cudaSuccess) return (int)e;
  nn_grad_kernel<<<dim3(b, 2, nn_grad_parts(b, n < m ? n : m)), NG_THREADS, smem, st>>>(
      n, xyz1, m, xyz2, nullptr, idx1, nullptr, idx2, grad_xyz1, grad_xyz2, grad_loss, scale1, scale2);
  return finish_launch(1);
}
