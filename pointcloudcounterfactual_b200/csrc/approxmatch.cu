// Approximate-matching EMD (approxmatch + match cost + gradients) for sm_100a.
//
// Replaces external/pytorch_structural_losses/src/approxmatch.cu (approxmatchkernel :3-182, matchcostkernel
// :184-224, matchcostgrad{1,2}kernel :229-291).  Same algorithm -- 9 annealing levels -4^j (j = 7..-1), each made of
// three all-pairs sweeps with E(k,l) = exp(level * |x1_k - x2_l|^2) -- restructured for B200:
//
//   phase A  "solve": every sweep is one instance of the same row-sum kernel
//                S[p] = sum_q E(p,q) * w[q]
//            with a per-sweep epilogue (ratioL / ratioR+remainR / remainL update).  All clouds and all row blocks
//            run in parallel over the 148 SMs (the reference runs one 512-thread CTA per cloud), two partner
//            points per step with packed FADD2/FMUL2/FFMA2 and one MUFU.EX2 each.  The per-level scaling vectors
//            ratioL_j, ratioR_j (9 x (n+m) floats per cloud) are kept.  The sweeps of the three steepest levels skip
//            the partners whose exponential is exactly +0 ("exact-zero culling" below: same bits, fewer pairs).
//   phase B  "apply": match[l][k] = sum_j E_j(k,l) ratioL_j[k] ratioR_j[l] is evaluated ONCE per pair
//            (5 exponentials per pair: levels j+1 are obtained from level j by two squarings) and either written
//            to the (b,m,n) matrix (pcc_approxmatch; one 0.5 GiB write instead of nine read-modify-writes) or
//            consumed on the fly into cost and gradients (pcc_matchcost_fused; the matrix never exists).
#include <cstdlib>

#include "common.cuh"

namespace pcc {

constexpr int AM_LEVELS = 9;  // j = 7, 6, ..., -1  (approxmatch.cu:24; the j == -2 / level 0 sweep is dead code)
constexpr int AM_THREADS = 128;
constexpr int AM_P_DEFAULT = 2;  // points owned by one thread in a sweep (tuned on B200)
constexpr int AM_QTILE = 2048;   // partner points per shared-memory tile in the sweep kernel (32 KiB)
constexpr float AM_LOG2E = 1.4426950408889634f;  // 0f3FB8AA3B, the constant __expf multiplies by

enum { EPI_RATIO_L = 0, EPI_RATIO_R = 1, EPI_REMAIN_L = 2 };

__global__ void am_init_kernel(int n, int m, float *__restrict__ temp, float multiL, float multiR) {
  pdl_enter();
  // temp per cloud: [remainL(n) | remainR(m) | ratioL(n) | ratioR(m)]   (approxmatch.cu:4,19-21)
  float *t = temp + (size_t)blockIdx.y * (size_t)(n + m) * 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n + m; i += gridDim.x * blockDim.x)
    t[i] = i < n ? multiL : multiR;
}

__global__ void am_levels_kernel(float *levels) {
  // the reference evaluates `-powf(4.0f, j)` on the device (approxmatch.cu:25); use the same routine
  const int t = threadIdx.x;
  if (t < AM_LEVELS) levels[t] = -powf(4.0f, (float)(7 - t));
}

// One sweep of the solver, BIT-FAITHFUL to the reference kernel: for every owned point p
//     acc_p = init;  for q = 0 .. nQ-1 (ascending):  acc_p = fma(E(p,q) [* rl_p], w[q], acc_p)
// with E = ex2.approx((d2 * level) * log2e), d2 = fma(dz,dz,fma(dx,dx,dy*dy)) -- the exact operation sequence nvcc
// emits for approxmatch.cu:29-62 / :78-111 / :130-163 -- followed by that sweep's epilogue.  The iteration amplifies
// rounding differences ~1000x into match / gradients, so the summation order and every rounding are kept; the
// parallelism is over points only (each thread owns AM_P points, packed two per f32x2 lane pair, so that one
// broadcast LDS.128 of a partner feeds four exponentials and the SFU pipe, not shared memory, is the limit).
// exp2((|x_q - x_p|^2) * lc) for the AM_P points of this thread (two per f32x2) against partner (qx,qy,qz)
template <int H>
__device__ __forceinline__ void am_exp_terms(const float4 q, const f32x2 *npx, const f32x2 *npy, const f32x2 *npz,
                                             f32x2 lc2, f32x2 *E) {
  const f32x2 qx = pack2(q.x, q.x), qy = pack2(q.y, q.y), qz = pack2(q.z, q.z);
#pragma unroll
  for (int h = 0; h < H; ++h) {
    const f32x2 dx = add2(qx, npx[h]), dy = add2(qy, npy[h]), dz = add2(qz, npz[h]);
    const f32x2 d2 = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
    float a0, a1;
    unpack2(mul2(d2, lc2), a0, a1);
    E[h] = pack2(ex2_ftz(a0), ex2_ftz(a1));
  }
}

#ifndef PCC_AM_UNROLL
#define PCC_AM_UNROLL 16
#endif
template <int EPI, int AM_P, int UNROLL = PCC_AM_UNROLL>
__global__ void __launch_bounds__(AM_THREADS)
am_sweep_kernel(int nP, int nQ, const float *__restrict__ xP, const float *__restrict__ xQ,
                const float *__restrict__ wQ, size_t wQ_stride, float level, float *__restrict__ remainP,
                size_t remain_stride, float *__restrict__ ratioP, size_t ratio_stride) {
  pdl_wait();
  constexpr int H = AM_P / 2, U = UNROLL;
  __shared__ float4 tile[AM_QTILE];  // (x, y, z, w) per partner
  const size_t cloud = blockIdx.y;
  xP += cloud * (size_t)nP * 3;
  xQ += cloud * (size_t)nQ * 3;
  wQ += cloud * wQ_stride;
  float *rem = remainP + cloud * remain_stride;
  float *rat = ratioP + cloud * ratio_stride;
  // thread t of block bx owns points p0 + t + u*AM_THREADS, u = 0..AM_P-1
  const int p0 = blockIdx.x * (AM_THREADS * AM_P) + threadIdx.x;
  f32x2 npx[H], npy[H], npz[H], acc[H], rl2[H];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    const int a = min(p0 + (2 * h) * AM_THREADS, nP - 1), b = min(p0 + (2 * h + 1) * AM_THREADS, nP - 1);
    npx[h] = pack2(-xP[a * 3], -xP[b * 3]);
    npy[h] = pack2(-xP[a * 3 + 1], -xP[b * 3 + 1]);
    npz[h] = pack2(-xP[a * 3 + 2], -xP[b * 3 + 2]);
    const float init = (EPI == EPI_RATIO_L) ? 1e-9f : 0.f;  // approxmatch.cu:37 / :86 / :138
    acc[h] = pack2(init, init);
    rl2[h] = (EPI == EPI_REMAIN_L) ? pack2(rat[a], rat[b]) : 0ull;  // ratioL[k] (approxmatch.cu:146)
  }
  // level = -4^j is a power of two, so d2*level is exact and ((d2*level)*log2e) == d2*(level*log2e) bit for bit:
  // one multiply reproduces the reference's two (approxmatch.cu:55 + __expf)
  const float lc = level * AM_LOG2E;
  const f32x2 lc2 = pack2(lc, lc);

  for (int base = 0; base < nQ; base += AM_QTILE) {
    const int cnt = min(AM_QTILE, nQ - base);
    const int cntU = (cnt + U - 1) / U * U;
    __syncthreads();
    for (int i = threadIdx.x; i < cntU; i += AM_THREADS) {
      // padding partners carry weight 0: fma(E, 0, acc) == acc exactly, the summation order is untouched
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < cnt) {
        const float *q = xQ + (size_t)(base + i) * 3;
        v = make_float4(q[0], q[1], q[2], wQ[base + i]);
      }
      tile[i] = v;
    }
    __syncthreads();
    // U partners per iteration: their exponentials are independent work; the accumulation itself stays strictly
    // ordered (acc = fma(E_l, w_l, acc), l ascending -- the reference's order).
#pragma unroll 1
    for (int l = 0; l < cntU; l += U) {
      f32x2 E[U][H];
      float w[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float4 q = tile[l + u];
        w[u] = q.w;
        am_exp_terms<H>(q, npx, npy, npz, lc2, E[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const f32x2 w2 = pack2(w[u], w[u]);
#pragma unroll
        for (int h = 0; h < H; ++h) {
          f32x2 e = E[u][h];
          if (EPI == EPI_REMAIN_L) e = mul2(rl2[h], e);  // rl * E, then fma(., ratioR, suml)  (compiled :155-157)
          acc[h] = fma2(e, w2, acc[h]);
        }
      }
    }
  }
  pdl_trigger();
#pragma unroll
  for (int h = 0; h < H; ++h) {
    float s[2];
    unpack2(acc[h], s[0], s[1]);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int p = p0 + (2 * h + e) * AM_THREADS;
      if (p >= nP) continue;
      if (EPI == EPI_RATIO_L) {  // approxmatch.cu:60-61
        rat[p] = rem[p] / s[e];
      } else if (EPI == EPI_RATIO_R) {  // approxmatch.cu:104-109
        const float r = rem[p];
        const float sumr = s[e] * r;
        const float consumption = fminf(r / (sumr + 1e-9f), 1.0f);
        rat[p] = consumption * r;
        rem[p] = fmaxf(0.0f, r - sumr);
      } else {  // approxmatch.cu:161-162
        rem[p] = fmaxf(0.0f, rem[p] - s[e]);
      }
    }
  }
}

// Fused sweep: sweep 3 of level A (remainL update, approxmatch.cu:130-163) and sweep 1 of the NEXT level B
// (ratioL, :29-62) walk the same (own point k) x (all partners l) loop, so the distance is evaluated once and feeds
// two exponentials -- 10 instead of 16 FMA-pipe operations per pair for the two sweeps, one partner load instead of two.
// Each accumulator still sees exactly the reference's operation sequence, so the results stay bit-faithful.
#ifndef PCC_AM31_UNROLL
#define PCC_AM31_UNROLL 8
#endif
template <int AM_P, int UNROLL = PCC_AM31_UNROLL>
__global__ void __launch_bounds__(AM_THREADS)
am_sweep31_kernel(int nP, int nQ, const float *__restrict__ xP, const float *__restrict__ xQ,
                  const float *__restrict__ ratioR_A, size_t ratioR_stride, const float *__restrict__ remainR,
                  size_t remainR_stride, float levelA, float levelB, float *__restrict__ remainL,
                  size_t remainL_stride, const float *__restrict__ ratioL_A, float *__restrict__ ratioL_B,
                  size_t ratioL_stride) {
  pdl_wait();
  __shared__ float4 tile[AM_QTILE];  // (x, y, z, ratioR_A) per partner
  __shared__ float wB[AM_QTILE];     // remainR per partner
  const size_t cloud = blockIdx.y;
  xP += cloud * (size_t)nP * 3;
  xQ += cloud * (size_t)nQ * 3;
  ratioR_A += cloud * ratioR_stride;
  remainR += cloud * remainR_stride;
  float *rem = remainL + cloud * remainL_stride;
  const float *ratA = ratioL_A + cloud * ratioL_stride;
  float *ratB = ratioL_B + cloud * ratioL_stride;
  const int p0 = blockIdx.x * (AM_THREADS * AM_P) + threadIdx.x;
  f32x2 npx[AM_P / 2], npy[AM_P / 2], npz[AM_P / 2], acc3[AM_P / 2], acc1[AM_P / 2], rl2[AM_P / 2];
#pragma unroll
  for (int h = 0; h < AM_P / 2; ++h) {
    const int a = min(p0 + (2 * h) * AM_THREADS, nP - 1), b = min(p0 + (2 * h + 1) * AM_THREADS, nP - 1);
    npx[h] = pack2(-xP[a * 3], -xP[b * 3]);
    npy[h] = pack2(-xP[a * 3 + 1], -xP[b * 3 + 1]);
    npz[h] = pack2(-xP[a * 3 + 2], -xP[b * 3 + 2]);
    acc3[h] = 0ull;
    acc1[h] = pack2(1e-9f, 1e-9f);
    rl2[h] = pack2(ratA[a], ratA[b]);
  }
  const float lcA = levelA * AM_LOG2E, lcB = levelB * AM_LOG2E;
  const f32x2 lcA2 = pack2(lcA, lcA), lcB2 = pack2(lcB, lcB);

  for (int base = 0; base < nQ; base += AM_QTILE) {
    const int cnt = min(AM_QTILE, nQ - base);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += AM_THREADS) {
      const float *q = xQ + (size_t)(base + i) * 3;
      tile[i] = make_float4(q[0], q[1], q[2], ratioR_A[base + i]);
      wB[i] = remainR[base + i];
    }
    __syncthreads();
#pragma unroll UNROLL
    for (int l = 0; l < cnt; ++l) {
      const float4 q = tile[l];
      const float w1 = wB[l];
      const f32x2 qx = pack2(q.x, q.x), qy = pack2(q.y, q.y), qz = pack2(q.z, q.z), qw3 = pack2(q.w, q.w),
                  qw1 = pack2(w1, w1);
#pragma unroll
      for (int h = 0; h < AM_P / 2; ++h) {
        const f32x2 dx = add2(qx, npx[h]), dy = add2(qy, npy[h]), dz = add2(qz, npz[h]);
        const f32x2 d2 = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
        float a0, a1, b0, b1;
        unpack2(mul2(d2, lcA2), a0, a1);
        unpack2(mul2(d2, lcB2), b0, b1);
        const f32x2 EA = pack2(ex2_ftz(a0), ex2_ftz(a1));
        const f32x2 EB = pack2(ex2_ftz(b0), ex2_ftz(b1));
        acc3[h] = fma2(mul2(rl2[h], EA), qw3, acc3[h]);
        acc1[h] = fma2(EB, qw1, acc1[h]);
      }
    }
  }
  pdl_trigger();
#pragma unroll
  for (int h = 0; h < AM_P / 2; ++h) {
    float s3[2], s1[2];
    unpack2(acc3[h], s3[0], s3[1]);
    unpack2(acc1[h], s1[0], s1[1]);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int p = p0 + (2 * h + e) * AM_THREADS;
      if (p >= nP) continue;
      const float r = fmaxf(0.0f, rem[p] - s3[e]);  // approxmatch.cu:161-162 (level A)
      rem[p] = r;
      ratB[p] = r / s1[e];                          // approxmatch.cu:60-61   (level B)
    }
  }
}

// ---- exact-zero culling for the steep levels --------------------------------------------------------------------
// E = ex2.approx.ftz(a) is EXACTLY +0 once a = d2 * level * log2e < -126 (the result would be subnormal), and
// fma(+0 [* rl], w, acc) == acc bit for bit for finite w: a partner further than R_j = sqrt(126 / (4^j log2e)) from an
// own point contributes NOTHING to that point's sequential sum, so dropping it leaves every rounding of the
// reference's chain untouched.  At levels j = 7, 6, 5 (R = 0.073, 0.146, 0.29 on unit-sphere clouds) that is most pairs.
//   am_group_kernel      per cloud and side, a permutation that makes consecutive blocks of 64 points spatially
//                        tight (median splits by rank: counting sort by x, then by y inside every quarter, then by z
//                        inside every sixteenth).  Which thread owns which point is free: only the partner ORDER is fixed.
//   am_sweep*_cull       a warp owns one block of 64 points (two per lane).  It tests every staged partner against
//                        the block's bounding box with a conservative cutoff (130 instead of 126: the box distance
//                        is a lower bound of the computed d2 up to rounding) and appends the survivors' tile slots to
//                        its own list -- in ascending partner index, i.e. the reference's summation order -- then runs
//                        the unchanged inner loop over the list.  Lists are padded to the unroll width with a
//                        zero-weight dummy partner (fma(E, 0, acc) == acc).  Non-finite own points / own factors switch
//                        the cull off for the warp, a non-finite weight in the staged tile for the whole tile (0 * inf
//                        and NaN must keep propagating exactly as in the full sweep).
// Results are bit-identical to the full sweeps (tests/test_gpu_emd.py: culled == PCC_AM_NOCULL=1 on S1/S2/S3,
// collapsed and NaN clouds).  S1 clouds, blocks of 64: 12 % / 26 % / 56 % of the partners survive at j = 7 / 6 / 5.
constexpr int AMG_THREADS = 256;
constexpr int AMC_MAXPTS = 4096;    // clouds up to this size take the culled sweeps (u16 permutation, sort in smem)
constexpr int AMG_PPT = AMC_MAXPTS / AMG_THREADS;  // points per thread of the grouping kernel (keys in registers)
constexpr int AMC_LEVELS = 3;       // levels t = 0, 1, 2 (j = 7, 6, 5) are culled (j = 5: break-even on S1, a gain on spread clouds)
constexpr float AMC_CUT = 130.f;    // cull when boxdist2 * |level * log2e| > AMC_CUT
constexpr int AMC_PAD = 16;         // list padding granularity (>= both unroll widths) and spare tile entries

// exclusive scan of cnt[0 .. nb) in place, nb <= 2 * AMG_THREADS: two adjacent bins per thread
__device__ __forceinline__ void amg_scan(int *cnt, int nb, int *warp_tot) {
  const int b0 = 2 * threadIdx.x;
  const int c0 = b0 < nb ? cnt[b0] : 0, c1 = b0 + 1 < nb ? cnt[b0 + 1] : 0;
  const int local = c0 + c1;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  int off = inc - local;
#pragma unroll
  for (int u = 0; u < AMG_THREADS / 32; ++u) off += u < w ? warp_tot[u] : 0;
  if (b0 < nb) cnt[b0] = off;
  if (b0 + 1 < nb) cnt[b0 + 1] = off + c0;
  __syncthreads();
}

// Three counting sorts: by x over the whole cloud (256 bins), by y inside every quarter of that order (64 bins each), by z
// inside every sixteenth (32 bins each) -- median splits by RANK, so the blocks stay balanced whatever the density.  The
// order inside a bin (shared-memory atomics) is arbitrary: nothing downstream depends on which thread owns which point.
// The cloud is staged in shared memory once (every pass gathers coordinates through the current order); a thread keeps
// the keys of its positions in registers between the histogram and the scatter.  With `temp` it also initialises this
// side's remain vector (am_init_kernel's job, approxmatch.cu:19-21) so the solve starts one launch earlier.
__global__ void __launch_bounds__(AMG_THREADS)
am_group_kernel(int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                unsigned short *__restrict__ perm1, unsigned short *__restrict__ perm2, float *__restrict__ temp,
                float multiL, float multiR) {
  extern __shared__ __align__(16) float gx[];  // [cnt][3]
  __shared__ unsigned short ord[2][AMC_MAXPTS];
  __shared__ int hist[16 * 32];
  __shared__ int warp_tot[AMG_THREADS / 32];
  __shared__ float red[6][AMG_THREADS / 32];
  const size_t cloud = blockIdx.x;
  const int side = blockIdx.y;
  const int cnt = side ? m : n;
  const float *__restrict__ x = (side ? xyz2 : xyz1) + cloud * (size_t)cnt * 3;
  unsigned short *__restrict__ perm = (side ? perm2 : perm1) + cloud * (size_t)cnt;
  const float INF = __int_as_float(0x7f800000);
  if (temp) {  // temp per cloud: [remainL(n) | remainR(m) | ratioL(n) | ratioR(m)]
    float *t = temp + cloud * (size_t)(n + m) * 2 + (side ? n : 0);
    const float v = side ? multiR : multiL;
    for (int i = threadIdx.x; i < cnt; i += AMG_THREADS) t[i] = v;
  }
  // stage the coordinates; extents per axis (non-finite coordinates are ignored; they land in bin 0)
  float lo[3] = {INF, INF, INF}, hi[3] = {-INF, -INF, -INF};
  for (int i = threadIdx.x; i < cnt; i += AMG_THREADS) {
    ord[0][i] = (unsigned short)i;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v = x[i * 3 + a];
      gx[i * 3 + a] = v;
      if (fabsf(v) <= 3.0e38f) {
        lo[a] = fminf(lo[a], v);
        hi[a] = fmaxf(hi[a], v);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if ((threadIdx.x & 31) == 0) {
      red[a][threadIdx.x >> 5] = lo[a];
      red[3 + a][threadIdx.x >> 5] = hi[a];
    }
  }
  __syncthreads();
  int cur = 0;
  for (int axis = 0; axis < 3; ++axis) {
    const int nseg = axis == 0 ? 1 : (axis == 1 ? 4 : 16);
    const int nbin = axis == 0 ? 256 : (axis == 1 ? 64 : 32);
    const int seg_len = (cnt + nseg - 1) / nseg;
    const float inv_len = 1.f / (float)seg_len;  // (pos + 0.5) / seg_len is never within 1e-4 of an integer: exact floor
    for (int i = threadIdx.x; i < nseg * nbin; i += AMG_THREADS) hist[i] = 0;
    float l = red[axis][0], h = red[3 + axis][0];
#pragma unroll
    for (int w = 1; w < AMG_THREADS / 32; ++w) {
      l = fminf(l, red[axis][w]);
      h = fmaxf(h, red[3 + axis][w]);
    }
    const float scale = (h > l) ? (float)nbin / (h - l) : 0.f;
    __syncthreads();
    int key[AMG_PPT];
#pragma unroll
    for (int u = 0; u < AMG_PPT; ++u) {
      const int pos = threadIdx.x + u * AMG_THREADS;
      key[u] = -1;
      if (pos < cnt) {
        const float v = gx[ord[cur][pos] * 3 + axis];
        int bin = (fabsf(v) <= 3.0e38f) ? (int)((v - l) * scale) : 0;
        bin = min(max(bin, 0), nbin - 1);
        key[u] = (int)(((float)pos + 0.5f) * inv_len) * nbin + bin;
        atomicAdd(&hist[key[u]], 1);
      }
    }
    __syncthreads();
    amg_scan(hist, nseg * nbin, warp_tot);
#pragma unroll
    for (int u = 0; u < AMG_PPT; ++u) {
      const int pos = threadIdx.x + u * AMG_THREADS;
      if (key[u] >= 0) ord[cur ^ 1][atomicAdd(&hist[key[u]], 1)] = ord[cur][pos];
    }
    __syncthreads();
    cur ^= 1;
  }
  for (int i = threadIdx.x; i < cnt; i += AMG_THREADS) perm[i] = ord[cur][i];
}

// shared with the Chamfer forward (chamfer.cu): both clouds of every pair grouped in one launch
int am_group_launch(int b, int n, int m, const float *xyz1, const float *xyz2, unsigned short *perm1,
                    unsigned short *perm2, float *temp, float multiL, float multiR, cudaStream_t st) {
  if (n > AMC_MAXPTS || m > AMC_MAXPTS) return PCC_ENOTSUP;
  const size_t gsm = sizeof(float) * 3 * (size_t)(n > m ? n : m);
  static size_t ag[64];
  // the 48 KiB default limit counts the kernel's ~18 KiB of static shared memory too: opt in from 28 KiB of dynamic
  if (cudaError_t e = smem_optin(am_group_kernel, gsm + 20 * 1024, ag, 48 * 1024); e != cudaSuccess) return (int)e;
  am_group_kernel<<<dim3(b, 2), AMG_THREADS, gsm, st>>>(n, m, xyz1, xyz2, perm1, perm2, temp, multiL, multiR);
  return (int)cudaGetLastError();
}

// Which block of 64 grouped points warp w of CTA blockIdx.x takes.  Consecutive blocks are spatial neighbours (same
// partner density, same list length), and all CTAs of a culled sweep are resident at once, so a CTA of four neighbours
// would leave the dense regions' SMs running long after the sparse ones: instead the four warps of CTA bx take one
// block from each quarter of the order (q * gridDim.x + bx), rotated so that the CTA sharing the SM with this one
// (148 CTAs later: bx + 4 in a grid of 8 per cloud) puts the complementary quarter on every sub-partition.
__device__ __forceinline__ int am_cull_block(int warp) {
  const int bx = blockIdx.x;
  const int q = (warp + bx + 2 * (bx >> 2)) & 3;
  return q * gridDim.x + bx;
}

struct AmBox {
  float lx, hx, ly, hy, lz, hz;
  bool ok;  // every own coordinate finite: the cull may be applied
};

// bounding box of the warp's 64 own points (a, b per lane; both valid point indices)
__device__ __forceinline__ AmBox am_warp_box(float ax, float ay, float az, float bx, float by, float bz, float fa = 0.f,
                                             float fb = 0.f) {  // fa, fb: own factors that must be finite as well
  AmBox r;
  r.lx = fminf(ax, bx); r.hx = fmaxf(ax, bx);
  r.ly = fminf(ay, by); r.hy = fmaxf(ay, by);
  r.lz = fminf(az, bz); r.hz = fmaxf(az, bz);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    r.lx = fminf(r.lx, __shfl_xor_sync(0xffffffffu, r.lx, o));
    r.hx = fmaxf(r.hx, __shfl_xor_sync(0xffffffffu, r.hx, o));
    r.ly = fminf(r.ly, __shfl_xor_sync(0xffffffffu, r.ly, o));
    r.hy = fmaxf(r.hy, __shfl_xor_sync(0xffffffffu, r.hy, o));
    r.lz = fminf(r.lz, __shfl_xor_sync(0xffffffffu, r.lz, o));
    r.hz = fmaxf(r.hz, __shfl_xor_sync(0xffffffffu, r.hz, o));
  }
  const float big = 3.0e38f;
  const bool fin = fabsf(ax) <= big && fabsf(ay) <= big && fabsf(az) <= big && fabsf(bx) <= big && fabsf(by) <= big &&
                   fabsf(bz) <= big && fabsf(fa) <= big && fabsf(fb) <= big;  // false for NaN / inf
  r.ok = __all_sync(0xffffffffu, fin);
  return r;
}

// Appends, in ascending order, the byte offsets (slot * 16) of the staged partners [0, cnt) that may reach the box;
// pads the list to a multiple of AMC_PAD with the dummy slot `cnt`.  Returns the padded length.  `cull` is false when
// an own point or factor of this warp, or any staged weight, is not finite (0 * inf must stay NaN): every partner is
// kept then.  Two partners per lane and step (64 per step): their tests are independent work.
__device__ __forceinline__ int am_build_list(const float4 *tile, int cnt, const AmBox &bx, float cut, bool cull,
                                             unsigned short *list, int lane) {
  int len = 0;
  const unsigned int below = (1u << lane) - 1u;
  for (int i0 = 0; i0 < cnt; i0 += 64) {
    const int ia = i0 + lane, ib = i0 + 32 + lane;
    bool ka = false, kb = false;
    if (ia < cnt) {
      const float4 q = tile[ia];
      const float dx = fmaxf(fmaxf(bx.lx - q.x, q.x - bx.hx), 0.f);
      const float dy = fmaxf(fmaxf(bx.ly - q.y, q.y - bx.hy), 0.f);
      const float dz = fmaxf(fmaxf(bx.lz - q.z, q.z - bx.hz), 0.f);
      ka = !(fmaf(dz, dz, fmaf(dx, dx, dy * dy)) > cut) || !cull;  // NaN partner coordinates give 0 or NaN: kept
    }
    if (ib < cnt) {
      const float4 q = tile[ib];
      const float dx = fmaxf(fmaxf(bx.lx - q.x, q.x - bx.hx), 0.f);
      const float dy = fmaxf(fmaxf(bx.ly - q.y, q.y - bx.hy), 0.f);
      const float dz = fmaxf(fmaxf(bx.lz - q.z, q.z - bx.hz), 0.f);
      kb = !(fmaf(dz, dz, fmaf(dx, dx, dy * dy)) > cut) || !cull;
    }
    const unsigned int ba = __ballot_sync(0xffffffffu, ka), bb = __ballot_sync(0xffffffffu, kb);
    const int na = __popc(ba);
    if (ka) list[len + __popc(ba & below)] = (unsigned short)(ia * 16);
    if (kb) list[len + na + __popc(bb & below)] = (unsigned short)(ib * 16);
    len += na + __popc(bb);
  }
  const int lenp = (len + AMC_PAD - 1) / AMC_PAD * AMC_PAD;
  for (int i = len + lane; i < lenp; i += 32) list[i] = (unsigned short)(cnt * 16);
  __syncwarp();
  return lenp;
}

// dynamic shared memory of the culled sweeps: partner tile (+ spare entries), optional second weights, four lists
constexpr size_t AMC_TILE_BYTES = (size_t)(AM_QTILE + AMC_PAD) * 16;
constexpr size_t AMC_W_BYTES = (size_t)(AM_QTILE + AMC_PAD) * 4;
constexpr size_t AMC_LIST_BYTES = (size_t)(AM_QTILE + AMC_PAD) * 2;
constexpr size_t AMC_SMEM1 = AMC_TILE_BYTES + (AM_THREADS / 32) * AMC_LIST_BYTES;
constexpr size_t AMC_SMEM31 = AMC_TILE_BYTES + AMC_W_BYTES + (AM_THREADS / 32) * AMC_LIST_BYTES;

// am_sweep_kernel<EPI, 2, 16> over the culled partner lists; own points through `perm`.
template <int EPI>
__global__ void __launch_bounds__(AM_THREADS, 4)  // states the register budget: ptxas otherwise serialises the loop (44 regs)
am_sweep_cull_kernel(int nP, int nQ, const float *__restrict__ xP, const float *__restrict__ xQ,
                     const float *__restrict__ wQ, size_t wQ_stride, float level, float cut,
                     const unsigned short *__restrict__ permP, float *__restrict__ remainP, size_t remain_stride,
                     float *__restrict__ ratioP, size_t ratio_stride) {
  constexpr int U = 16;
  extern __shared__ __align__(16) unsigned char csm[];
  float4 *tile = reinterpret_cast<float4 *>(csm);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned short *list = reinterpret_cast<unsigned short *>(csm + AMC_TILE_BYTES + warp * AMC_LIST_BYTES);
  const size_t cloud = blockIdx.y;
  xP += cloud * (size_t)nP * 3;
  xQ += cloud * (size_t)nQ * 3;
  wQ += cloud * wQ_stride;
  permP += cloud * (size_t)nP;
  float *rem = remainP + cloud * remain_stride;
  float *rat = ratioP + cloud * ratio_stride;
  // warp w of block bx owns block am_cull_block(w) of 64 permuted slots: g0 + lane and g0 + 32 + lane
  const int g0 = am_cull_block(warp) * 64;
  const int sa = g0 + lane, sb = g0 + 32 + lane;
  const int a = permP[min(sa, nP - 1)], b = permP[min(sb, nP - 1)];
  const float ax = xP[a * 3], ay = xP[a * 3 + 1], az = xP[a * 3 + 2];
  const float bx_ = xP[b * 3], by_ = xP[b * 3 + 1], bz_ = xP[b * 3 + 2];
  const f32x2 npx = pack2(-ax, -bx_), npy = pack2(-ay, -by_), npz = pack2(-az, -bz_);
  const float init = (EPI == EPI_RATIO_L) ? 1e-9f : 0.f;
  f32x2 acc = pack2(init, init);
  const float rla = (EPI == EPI_REMAIN_L) ? rat[a] : 0.f, rlb = (EPI == EPI_REMAIN_L) ? rat[b] : 0.f;
  const f32x2 rl2 = pack2(rla, rlb);
  const float lc = level * AM_LOG2E;
  const f32x2 lc2 = pack2(lc, lc);
  const AmBox box = am_warp_box(ax, ay, az, bx_, by_, bz_, rla, rlb);  // inf * 0 must stay NaN: no cull then
  // the epilogue's operands, requested now (their latency would otherwise end the kernel)
  const float rema = rem[a], remb = rem[b];

  for (int base = 0; base < nQ; base += AM_QTILE) {
    const int cnt = min(AM_QTILE, nQ - base);
    __syncthreads();
    bool wfin = true;
#pragma unroll 4
    for (int i = threadIdx.x; i < cnt + 1; i += AM_THREADS) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);  // slot cnt: the zero-weight dummy partner
      if (i < cnt) {
        const float *q = xQ + (size_t)(base + i) * 3;
        v = make_float4(q[0], q[1], q[2], wQ[base + i]);
      }
      wfin = wfin && fabsf(v.w) <= 3.0e38f;
      tile[i] = v;
    }
    const bool cull = __syncthreads_and(wfin) && box.ok;
    const int len = g0 < nP ? am_build_list(tile, cnt, box, cut, cull, list, lane) : 0;  // warp-uniform
    const unsigned char *tb = reinterpret_cast<const unsigned char *>(tile);
#pragma unroll 1
    for (int l = 0; l < len; l += U) {
      f32x2 E[U];
      float w[U];
      const uint4 o0 = *reinterpret_cast<const uint4 *>(list + l), o1 = *reinterpret_cast<const uint4 *>(list + l + 8);
      const unsigned int ow[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
      float4 q[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned int off = (u & 1) ? (ow[u >> 1] >> 16) : (ow[u >> 1] & 0xffffu);
        q[u] = *reinterpret_cast<const float4 *>(tb + off);
      }
      asm volatile("" ::: "memory");  // all gathers are issued before the first use (the compiler sank them otherwise)
#pragma unroll
      for (int u = 0; u < U; ++u) {
        w[u] = q[u].w;
        am_exp_terms<1>(q[u], &npx, &npy, &npz, lc2, &E[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        f32x2 e = E[u];
        if (EPI == EPI_REMAIN_L) e = mul2(rl2, e);
        acc = fma2(e, pack2(w[u], w[u]), acc);
      }
    }
  }
  float s[2];
  unpack2(acc, s[0], s[1]);
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    if ((e ? sb : sa) >= nP) continue;
    const int p = e ? b : a;
    const float r = e ? remb : rema;
    if (EPI == EPI_RATIO_L) {  // approxmatch.cu:60-61
      rat[p] = r / s[e];
    } else if (EPI == EPI_RATIO_R) {  // approxmatch.cu:104-109
      const float sumr = s[e] * r;
      const float consumption = fminf(r / (sumr + 1e-9f), 1.0f);
      rat[p] = consumption * r;
      rem[p] = fmaxf(0.0f, r - sumr);
    } else {  // approxmatch.cu:161-162
      rem[p] = fmaxf(0.0f, r - s[e]);
    }
  }
}

// am_sweep31_kernel<2> over the culled partner lists (the cutoff is level B's: E_B == 0 implies E_A == 0).
__global__ void __launch_bounds__(AM_THREADS, 4)
am_sweep31_cull_kernel(int nP, int nQ, const float *__restrict__ xP, const float *__restrict__ xQ,
                       const float *__restrict__ ratioR_A, size_t ratioR_stride, const float *__restrict__ remainR,
                       size_t remainR_stride, float levelA, float levelB, float cut,
                       const unsigned short *__restrict__ permP, float *__restrict__ remainL, size_t remainL_stride,
                       const float *__restrict__ ratioL_A, float *__restrict__ ratioL_B, size_t ratioL_stride) {
  constexpr int U = 8;
  extern __shared__ __align__(16) unsigned char csm[];
  float4 *tile = reinterpret_cast<float4 *>(csm);               // (x, y, z, ratioR_A) per partner
  float *wB = reinterpret_cast<float *>(csm + AMC_TILE_BYTES);  // remainR per partner
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned short *list =
      reinterpret_cast<unsigned short *>(csm + AMC_TILE_BYTES + AMC_W_BYTES + warp * AMC_LIST_BYTES);
  const size_t cloud = blockIdx.y;
  xP += cloud * (size_t)nP * 3;
  xQ += cloud * (size_t)nQ * 3;
  ratioR_A += cloud * ratioR_stride;
  remainR += cloud * remainR_stride;
  permP += cloud * (size_t)nP;
  float *rem = remainL + cloud * remainL_stride;
  const float *ratA = ratioL_A + cloud * ratioL_stride;
  float *ratB = ratioL_B + cloud * ratioL_stride;
  const int g0 = am_cull_block(warp) * 64;
  const int sa = g0 + lane, sb = g0 + 32 + lane;
  const int a = permP[min(sa, nP - 1)], b = permP[min(sb, nP - 1)];
  const float ax = xP[a * 3], ay = xP[a * 3 + 1], az = xP[a * 3 + 2];
  const float bx_ = xP[b * 3], by_ = xP[b * 3 + 1], bz_ = xP[b * 3 + 2];
  const f32x2 npx = pack2(-ax, -bx_), npy = pack2(-ay, -by_), npz = pack2(-az, -bz_);
  f32x2 acc3 = 0ull, acc1 = pack2(1e-9f, 1e-9f);
  const float rla = ratA[a], rlb = ratA[b];
  const f32x2 rl2 = pack2(rla, rlb);
  const float lcA = levelA * AM_LOG2E, lcB = levelB * AM_LOG2E;
  const f32x2 lcA2 = pack2(lcA, lcA), lcB2 = pack2(lcB, lcB);
  const AmBox box = am_warp_box(ax, ay, az, bx_, by_, bz_, rla, rlb);
  const float rema = rem[a], remb = rem[b];  // the epilogue's operands, requested now

  for (int base = 0; base < nQ; base += AM_QTILE) {
    const int cnt = min(AM_QTILE, nQ - base);
    __syncthreads();
    bool wfin = true;
#pragma unroll 4
    for (int i = threadIdx.x; i < cnt + 1; i += AM_THREADS) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      float w1 = 0.f;
      if (i < cnt) {
        const float *q = xQ + (size_t)(base + i) * 3;
        v = make_float4(q[0], q[1], q[2], ratioR_A[base + i]);
        w1 = remainR[base + i];
      }
      wfin = wfin && fabsf(v.w) <= 3.0e38f && fabsf(w1) <= 3.0e38f;
      tile[i] = v;
      wB[i] = w1;
    }
    const bool cull = __syncthreads_and(wfin) && box.ok;
    const int len = g0 < nP ? am_build_list(tile, cnt, box, cut, cull, list, lane) : 0;  // warp-uniform
    const unsigned char *tb = reinterpret_cast<const unsigned char *>(tile);
    const unsigned char *wb = reinterpret_cast<const unsigned char *>(wB);
#pragma unroll 1
    for (int l = 0; l < len; l += U) {
      const uint4 o0 = *reinterpret_cast<const uint4 *>(list + l);
      const unsigned int ow[4] = {o0.x, o0.y, o0.z, o0.w};
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned int off = (u & 1) ? (ow[u >> 1] >> 16) : (ow[u >> 1] & 0xffffu);
        const float4 q = *reinterpret_cast<const float4 *>(tb + off);
        const float w1 = *reinterpret_cast<const float *>(wb + (off >> 2));
        const f32x2 qx = pack2(q.x, q.x), qy = pack2(q.y, q.y), qz = pack2(q.z, q.z), qw3 = pack2(q.w, q.w),
                    qw1 = pack2(w1, w1);
        const f32x2 dx = add2(qx, npx), dy = add2(qy, npy), dz = add2(qz, npz);
        const f32x2 d2 = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
        float a0, a1, b0, b1;
        unpack2(mul2(d2, lcA2), a0, a1);
        unpack2(mul2(d2, lcB2), b0, b1);
        const f32x2 EA = pack2(ex2_ftz(a0), ex2_ftz(a1));
        const f32x2 EB = pack2(ex2_ftz(b0), ex2_ftz(b1));
        acc3 = fma2(mul2(rl2, EA), qw3, acc3);
        acc1 = fma2(EB, qw1, acc1);
      }
    }
  }
  float s3[2], s1[2];
  unpack2(acc3, s3[0], s3[1]);
  unpack2(acc1, s1[0], s1[1]);
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    if ((e ? sb : sa) >= nP) continue;
    const int p = e ? b : a;
    const float r = fmaxf(0.0f, (e ? remb : rema) - s3[e]);  // approxmatch.cu:161-162 (level A)
    rem[p] = r;
    ratB[p] = r / s1[e];                                     // approxmatch.cu:60-61   (level B)
  }
}

// ---- phase B -------------------------------------------------------------------------------------------------
// match(p,q) = sum_t E_t(p,q) * a_t[p] * b_t[q],  t = 0..8  <->  j = 7..-1,  E_t = exp2(c_t * d2)
struct AmLevels {
  float lv[AM_LEVELS];  // -4^j as the device's powf returns it (am_levels_kernel)
  float lc[AM_LEVELS];  // lv * log2e: (d2 * lv) * log2e == d2 * lc bit for bit, lv being a power of two (see am_sweep_kernel)
};

constexpr int AMF_QTILE = 256;                      // partner points per tile in the phase-B kernels
constexpr int AMF_F4_PER_PAIR = (3 + AM_LEVELS) / 2;  // 6 float4 per partner pair: 3 coords + 9 factors, 2 lanes each

// evaluates the 9 level terms for two partner points at once (packed lanes = partners):
//     match = fma(rl_j * E_j, rr_j, match)   for j = 7 .. -1          (compiled approxmatch.cu:155-157)
// EXACT: every E_j by MUFU.EX2 on ((d2*level)*log2e) -- bit-identical to the reference's terms.
// !EXACT (fused cost/gradient path, nothing downstream amplifies the error): E_{j+1} = E_j^4 from 5 anchors.
template <bool EXACT>
__device__ __forceinline__ f32x2 am_match_pair(f32x2 d2, const f32x2 *bq /*9*/, const float *ap /*9*/,
                                               const AmLevels &lv) {
  float lo, hi;
  f32x2 E[AM_LEVELS];
#pragma unroll
  for (int t = 0; t < AM_LEVELS; t += (EXACT ? 1 : 2)) {  // anchors j = 7, 5, 3, 1, -1 (t = 0, 2, 4, 6, 8)
    const f32x2 a = mul2(d2, pack2(lv.lc[t], lv.lc[t]));
    unpack2(a, lo, hi);
    E[t] = pack2(ex2_ftz(lo), ex2_ftz(hi));
  }
  if (!EXACT) {
#pragma unroll
    for (int t = 2; t < AM_LEVELS; t += 2) {
      const f32x2 s = mul2(E[t], E[t]);
      E[t - 1] = mul2(s, s);
    }
  }
  f32x2 mt = 0ull;
#pragma unroll
  for (int t = 0; t < AM_LEVELS; ++t) mt = fma2(mul2(pack2(ap[t], ap[t]), E[t]), bq[t], mt);
  return mt;
}

// Loads partner points [base, base+cnt) with their 9 factors into the pair-packed tile.
__device__ __forceinline__ void amf_fill_tile(float *tf, const float *__restrict__ xQ, const float *__restrict__ fQ,
                                              size_t f_level_stride, int base, int cnt, int cnt2) {
  for (int i = threadIdx.x; i < cnt2; i += AM_THREADS) {
    float v[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) v[e] = 0.f;
    if (i < cnt) {
      const float *q = xQ + (size_t)(base + i) * 3;
      v[0] = q[0];
      v[1] = q[1];
      v[2] = q[2];
#pragma unroll
      for (int t = 0; t < AM_LEVELS; ++t) v[3 + t] = fQ[(size_t)t * f_level_stride + base + i];
    }
    float *o = tf + (size_t)(i >> 1) * (AMF_F4_PER_PAIR * 4) + (i & 1);
#pragma unroll
    for (int e = 0; e < 12; ++e) o[e * 2] = v[e];
  }
}

// MATERIALISE: thread = cloud-1 point k (coalesced along n), blockIdx.y = slab of cloud-2 points.
__global__ void __launch_bounds__(AM_THREADS)
am_materialize_kernel(int n, int m, int lslab, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                      const float *__restrict__ fL, const float *__restrict__ fR, size_t fL_level_stride,
                      size_t fR_level_stride, AmLevels sc, float *__restrict__ match) {
  __shared__ float4 tile[AMF_QTILE / 2 * AMF_F4_PER_PAIR];
  const size_t cloud = blockIdx.z;
  xyz1 += cloud * (size_t)n * 3;
  xyz2 += cloud * (size_t)m * 3;
  fL += cloud * (size_t)n;
  fR += cloud * (size_t)m;
  match += cloud * (size_t)n * m;
  const int k = blockIdx.x * AM_THREADS + threadIdx.x;
  const int kc = min(k, n - 1);
  const float px = xyz1[kc * 3], py = xyz1[kc * 3 + 1], pz = xyz1[kc * 3 + 2];
  const f32x2 npx = pack2(-px, -px), npy = pack2(-py, -py), npz = pack2(-pz, -pz);
  float ap[AM_LEVELS];
#pragma unroll
  for (int t = 0; t < AM_LEVELS; ++t) ap[t] = fL[(size_t)t * fL_level_stride + kc];
  const int l_begin = blockIdx.y * lslab, l_end = min(m, l_begin + lslab);
  for (int base = l_begin; base < l_end; base += AMF_QTILE) {
    const int cnt = min(AMF_QTILE, l_end - base), cnt2 = (cnt + 1) & ~1;
    __syncthreads();
    amf_fill_tile(reinterpret_cast<float *>(tile), xyz2, fR, fR_level_stride, base, cnt, cnt2);
    __syncthreads();
    for (int pr = 0; pr < (cnt2 >> 1); ++pr) {
      const float4 *T = tile + pr * AMF_F4_PER_PAIR;
      const float4 c0 = T[0], c1 = T[1], c2 = T[2], c3 = T[3], c4 = T[4], c5 = T[5];
      const f32x2 bq[AM_LEVELS] = {pack2(c1.z, c1.w), pack2(c2.x, c2.y), pack2(c2.z, c2.w),
                                   pack2(c3.x, c3.y), pack2(c3.z, c3.w), pack2(c4.x, c4.y),
                                   pack2(c4.z, c4.w), pack2(c5.x, c5.y), pack2(c5.z, c5.w)};
      const f32x2 d2 = sqdist2(pack2(c0.x, c0.y), pack2(c0.z, c0.w), pack2(c1.x, c1.y), npx, npy, npz);
      float m0, m1;
      unpack2(am_match_pair<true>(d2, bq, ap, sc), m0, m1);
      const int l = base + pr * 2;
      if (k < n) {
        match[(size_t)l * n + k] = m0;
        if (l + 1 < l_end) match[(size_t)(l + 1) * n + k] = m1;
      }
    }
  }
}

// COST + GRADIENT of the point set P against Q without materialising match:
//   cost_part[cloud][blockIdx.x] = sum_{p in block} sum_q match * |x_p - x_q|           (approxmatch.cu:184-224)
//   gradP[p] = sum_q match * (x_p - x_q) * rsqrt(max(|x_p - x_q|^2, 1e-20))              (approxmatch.cu:229-291)
__global__ void __launch_bounds__(AM_THREADS)
am_costgrad_kernel(int nP, int nQ, const float *__restrict__ xP, const float *__restrict__ xQ,
                   const float *__restrict__ fP, const float *__restrict__ fQ, size_t fP_level_stride,
                   size_t fQ_level_stride, AmLevels sc, float *__restrict__ cost_part, float *__restrict__ gradP) {
  pdl_wait();
  __shared__ float4 tile[AMF_QTILE / 2 * AMF_F4_PER_PAIR];
  __shared__ float red[AM_THREADS / 32];
  const size_t cloud = blockIdx.y;
  xP += cloud * (size_t)nP * 3;
  xQ += cloud * (size_t)nQ * 3;
  fP += cloud * (size_t)nP;
  fQ += cloud * (size_t)nQ;
  const int p = blockIdx.x * AM_THREADS + threadIdx.x;
  const int pc = min(p, nP - 1);
  const float px = xP[pc * 3], py = xP[pc * 3 + 1], pz = xP[pc * 3 + 2];
  const f32x2 npx = pack2(-px, -px), npy = pack2(-py, -py), npz = pack2(-pz, -pz);
  float ap[AM_LEVELS];
#pragma unroll
  for (int t = 0; t < AM_LEVELS; ++t) ap[t] = fP[(size_t)t * fP_level_stride + pc];
  f32x2 cst = 0ull, gx = 0ull, gy = 0ull, gz = 0ull;
  const f32x2 tiny = pack2(1e-20f, 1e-20f);
  (void)tiny;
  for (int base = 0; base < nQ; base += AMF_QTILE) {
    const int cnt = min(AMF_QTILE, nQ - base), cnt2 = (cnt + 1) & ~1;
    __syncthreads();
    amf_fill_tile(reinterpret_cast<float *>(tile), xQ, fQ, fQ_level_stride, base, cnt, cnt2);
    __syncthreads();
    for (int pr = 0; pr < (cnt2 >> 1); ++pr) {
      const float4 *T = tile + pr * AMF_F4_PER_PAIR;
      const float4 c0 = T[0], c1 = T[1], c2 = T[2], c3 = T[3], c4 = T[4], c5 = T[5];
      const f32x2 bq[AM_LEVELS] = {pack2(c1.z, c1.w), pack2(c2.x, c2.y), pack2(c2.z, c2.w),
                                   pack2(c3.x, c3.y), pack2(c3.z, c3.w), pack2(c4.x, c4.y),
                                   pack2(c4.z, c4.w), pack2(c5.x, c5.y), pack2(c5.z, c5.w)};
      // dq = x_q - x_p (packed over the two partners); gradient uses x_p - x_q = -dq
      const f32x2 dx = add2(pack2(c0.x, c0.y), npx), dy = add2(pack2(c0.z, c0.w), npy),
                  dz = add2(pack2(c1.x, c1.y), npz);
      const f32x2 d2 = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
      const f32x2 mt = am_match_pair<false>(d2, bq, ap, sc);
      float d0, d1;
      unpack2(d2, d0, d1);
      const f32x2 rinv = pack2(rsqrt_ftz(fmaxf(d0, 1e-20f)), rsqrt_ftz(fmaxf(d1, 1e-20f)));
      const f32x2 u = mul2(mt, rinv);       // match / dist
      cst = fma2(u, d2, cst);               // match * dist   (dist = d2 * rsqrt(d2))
      gx = fma2(u, dx, gx);                 // accumulates match * (x_q - x_p) / dist ; negated at the end
      gy = fma2(u, dy, gy);
      gz = fma2(u, dz, gz);
    }
  }
  pdl_trigger();
  float c0, c1, x0, x1, y0, y1, z0, z1;
  unpack2(cst, c0, c1);
  unpack2(gx, x0, x1);
  unpack2(gy, y0, y1);
  unpack2(gz, z0, z1);
  if (p < nP && gradP) {
    float *g = gradP + (cloud * (size_t)nP + p) * 3;
    g[0] = -(x0 + x1);
    g[1] = -(y0 + y1);
    g[2] = -(z0 + z1);
  }
  if (cost_part) {
    float c = (p < nP) ? (c0 + c1) : 0.f;
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < AM_THREADS / 32; ++w) s += red[w];
      cost_part[cloud * gridDim.x + blockIdx.x] = s;
    }
  }
}

__global__ void am_cost_reduce_kernel(int b, int parts, const float *__restrict__ cost_part, float *__restrict__ cost) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b) return;
  float s = 0.f;
  for (int p = 0; p < parts; ++p) s += cost_part[(size_t)i * parts + p];  // fixed order: deterministic
  cost[i] = s;
}

__global__ void am_export_temp_kernel(int n, int m, const float *__restrict__ fL_last, const float *__restrict__ fR_last,
                                      float *__restrict__ temp) {
  pdl_enter();
  // temp's ratioL / ratioR slots receive the last level's vectors, as the reference leaves them (approxmatch.cu:4)
  float *t = temp + (size_t)blockIdx.y * (size_t)(n + m) * 2 + (n + m);
  const float *l = fL_last + (size_t)blockIdx.y * n, *r = fR_last + (size_t)blockIdx.y * m;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n + m; i += gridDim.x * blockDim.x)
    t[i] = i < n ? l[i] : r[i - n];
}

// ---- match cost / gradient from a materialised match matrix (API parity with MatchCost / MatchCostGrad) --------
constexpr int MC_THREADS = 256;
constexpr int MC_ROWS = 16;  // cloud-2 rows per CTA

// partial[cloud][blockIdx.x] = sum over rows l of this CTA, all k:  match[l][k] * sqrt(d2)
__global__ void __launch_bounds__(MC_THREADS)
matchcost_kernel(int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                 const float *__restrict__ match, float *__restrict__ partial) {
  __shared__ float red[MC_THREADS / 32];
  const size_t cloud = blockIdx.y;
  xyz1 += cloud * (size_t)n * 3;
  xyz2 += cloud * (size_t)m * 3;
  match += cloud * (size_t)n * m;
  const int l0 = blockIdx.x * MC_ROWS, l1 = min(m, l0 + MC_ROWS);
  float acc = 0.f;
  for (int k = threadIdx.x; k < n; k += MC_THREADS) {
    const float x = xyz1[k * 3], y = xyz1[k * 3 + 1], z = xyz1[k * 3 + 2];
    for (int l = l0; l < l1; ++l) {
      const float d = sqdist1(x, y, z, xyz2[l * 3], xyz2[l * 3 + 1], xyz2[l * 3 + 2]);
      acc = fmaf(match[(size_t)l * n + k], sqrtf(d), acc);
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < MC_THREADS / 32; ++w) s += red[w];
    partial[cloud * gridDim.x + blockIdx.x] = s;
  }
}

// grad1[k] = sum_l match[l][k] (x1_k - x2_l) rsqrt(max(d2,1e-20)) : thread per k, rows streamed (coalesced over k)
__global__ void __launch_bounds__(MC_THREADS)
matchcostgrad1_kernel(int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                      const float *__restrict__ match, float *__restrict__ grad1) {
  __shared__ float4 xs[256];
  const size_t cloud = blockIdx.y;
  xyz1 += cloud * (size_t)n * 3;
  xyz2 += cloud * (size_t)m * 3;
  match += cloud * (size_t)n * m;
  const int k = blockIdx.x * MC_THREADS + threadIdx.x;
  const int kc = min(k, n - 1);
  const float x = xyz1[kc * 3], y = xyz1[kc * 3 + 1], z = xyz1[kc * 3 + 2];
  float gx = 0.f, gy = 0.f, gz = 0.f;
  for (int base = 0; base < m; base += 256) {
    const int cnt = min(256, m - base);
    __syncthreads();
    if ((int)threadIdx.x < cnt) {
      const float *q = xyz2 + (size_t)(base + threadIdx.x) * 3;
      xs[threadIdx.x] = make_float4(q[0], q[1], q[2], 0.f);
    }
    __syncthreads();
    if (k < n) {
#pragma unroll 4
      for (int l = 0; l < cnt; ++l) {
        const float4 q = xs[l];
        const float dx = x - q.x, dy = y - q.y, dz = z - q.z;
        const float d = match[(size_t)(base + l) * n + k] * rsqrtf(fmaxf(dx * dx + dy * dy + dz * dz, 1e-20f));
        gx = fmaf(dx, d, gx);
        gy = fmaf(dy, d, gy);
        gz = fmaf(dz, d, gz);
      }
    }
  }
  if (k < n) {
    float *g = grad1 + (cloud * (size_t)n + k) * 3;
    g[0] = gx;
    g[1] = gy;
    g[2] = gz;
  }
}

// grad2[l] = sum_k match[l][k] (x2_l - x1_k) rsqrt(.) : one warp per row l (coalesced over k), ordered shuffle reduce
__global__ void __launch_bounds__(MC_THREADS)
matchcostgrad2_kernel(int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                      const float *__restrict__ match, float *__restrict__ grad2) {
  const size_t cloud = blockIdx.y;
  xyz1 += cloud * (size_t)n * 3;
  xyz2 += cloud * (size_t)m * 3;
  match += cloud * (size_t)n * m;
  const int lane = threadIdx.x & 31;
  const int l = blockIdx.x * (MC_THREADS / 32) + (threadIdx.x >> 5);
  if (l >= m) return;
  const float x = xyz2[l * 3], y = xyz2[l * 3 + 1], z = xyz2[l * 3 + 2];
  const float *row = match + (size_t)l * n;
  float gx = 0.f, gy = 0.f, gz = 0.f;
  for (int k = lane; k < n; k += 32) {
    const float dx = x - xyz1[k * 3], dy = y - xyz1[k * 3 + 1], dz = z - xyz1[k * 3 + 2];
    const float d = row[k] * rsqrtf(fmaxf(dx * dx + dy * dy + dz * dz, 1e-20f));
    gx = fmaf(dx, d, gx);
    gy = fmaf(dy, d, gy);
    gz = fmaf(dz, d, gz);
  }
  gx = warp_sum(gx);
  gy = warp_sum(gy);
  gz = warp_sum(gz);
  if (lane == 0) {
    float *g = grad2 + (cloud * (size_t)m + l) * 3;
    g[0] = gx;
    g[1] = gy;
    g[2] = gz;
  }
}

// ---- host side ----------------------------------------------------------------------------------------------
struct AmWorkspace {
  float *base = nullptr;
  float *fL = nullptr;  // [9][b][n]
  float *fR = nullptr;  // [9][b][m]
  float *cost_part = nullptr;
  unsigned short *perm1 = nullptr, *perm2 = nullptr;  // spatial grouping of either cloud (culled sweeps), or null
  size_t fL_level_stride = 0, fR_level_stride = 0;
};

// PCC_AM_NOCULL=1 (read at every call: the parity test toggles it) keeps every level on the full sweeps.
static bool am_cull_enabled(int n, int m) {
  if (n > AMC_MAXPTS || m > AMC_MAXPTS) return false;
  const char *e = getenv("PCC_AM_NOCULL");
  return !(e && e[0] && e[0] != '0');
}
// number of culled levels (tuning knob for tools/emd_cull_probe.py; any value gives the same bits)
static int am_cull_levels() {
  const char *e = getenv("PCC_AM_CULL_LEVELS");
  const int v = e && e[0] ? atoi(e) : AMC_LEVELS;
  return v < 0 ? 0 : (v > AM_LEVELS ? AM_LEVELS : v);
}


// -4^j for j = 7..-1, evaluated ONCE per process by the device's own powf (the reference calls powf in the kernel,
// approxmatch.cu:25).  Falls back to the host's powf when the first call happens under stream capture.
static int am_levels(cudaStream_t st, AmLevels *out) {
  static AmLevels cached;
  static bool have = false;
  if (!have) {
    for (int t = 0; t < AM_LEVELS; ++t) cached.lv[t] = -powf(4.0f, (float)(7 - t));
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    if (cap == cudaStreamCaptureStatusNone) {
      float *d = nullptr;
      if (cudaMalloc((void **)&d, sizeof(float) * AM_LEVELS) == cudaSuccess) {
        am_levels_kernel<<<1, 32, 0, st>>>(d);
        AmLevels dev;
        if (cudaMemcpyAsync(dev.lv, d, sizeof(float) * AM_LEVELS, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
            cudaStreamSynchronize(st) == cudaSuccess) {
          cached = dev;
          have = true;
        }
        cudaFree(d);
      }
      cudaGetLastError();
    }
  }
  for (int t = 0; t < AM_LEVELS; ++t) cached.lc[t] = cached.lv[t] * AM_LOG2E;
  *out = cached;
  return 0;
}

// Phase A: runs the 27 sweeps; leaves remainL/remainR in temp and the per-level factors in ws.
static int am_solve(int b, int n, int m, const float *xyz1, const float *xyz2, float *temp, AmWorkspace &ws,
                    size_t extra_floats, const AmLevels &lv, cudaStream_t st, int *launches, bool want_temp) {
  const size_t nfl = (size_t)AM_LEVELS * b * n, nfr = (size_t)AM_LEVELS * b * m;
  const bool cull = am_cull_enabled(n, m);
  const size_t perm_floats = cull ? ((size_t)b * (n + m) + 1) / 2 : 0;  // u16 entries
  cudaError_t e = ws_alloc((void **)&ws.base, sizeof(float) * (nfl + nfr + extra_floats + perm_floats), st);
  if (e != cudaSuccess) return (int)e;
  ws.fL = ws.base;
  ws.fR = ws.base + nfl;
  ws.cost_part = ws.base + nfl + nfr;
  if (cull) {
    ws.perm1 = reinterpret_cast<unsigned short *>(ws.base + nfl + nfr + extra_floats);
    ws.perm2 = ws.perm1 + (size_t)b * n;
    static size_t a1[64], a2[64], a4[64];
    if ((e = smem_optin(am_sweep_cull_kernel<EPI_RATIO_L>, AMC_SMEM1, a1)) != cudaSuccess ||
        (e = smem_optin(am_sweep_cull_kernel<EPI_RATIO_R>, AMC_SMEM1, a2)) != cudaSuccess ||
        (e = smem_optin(am_sweep31_cull_kernel, AMC_SMEM31, a4)) != cudaSuccess)
      return (int)e;
  }
  // cutoff on the squared box distance for level t: E underflows to +0 beyond it (see the culled kernels)
  auto cut = [&](int t) { return AMC_CUT / (-(lv.lv[t] * AM_LOG2E)); };
  ws.fL_level_stride = (size_t)b * n;
  ws.fR_level_stride = (size_t)b * m;
  float multiL, multiR;  // approxmatch.cu:6-12 (integer division)
  if (n >= m) {
    multiL = 1.f;
    multiR = (float)(n / m);
  } else {
    multiL = (float)(m / n);
    multiR = 1.f;
  }
  const size_t tstride = (size_t)(n + m) * 2;
  float *remainL = temp, *remainR = temp + n;
  if (cull) {  // the grouping kernel also initialises remainL / remainR
    if (int rc = am_group_launch(b, n, m, xyz1, xyz2, ws.perm1, ws.perm2, temp, multiL, multiR, st); rc != 0) return rc;
  } else {
    PCC_LAUNCH(PDL_EMD_SMALL, am_init_kernel, dim3((n + m + 255) / 256, b), 256, 0, st, n, m, temp, multiL, multiR);
  }
  constexpr int P = AM_P_DEFAULT;
  const int per_cta = AM_THREADS * P;
  const dim3 gk((n + per_cta - 1) / per_cta, b), gl((m + per_cta - 1) / per_cta, b);
  // level t: sweep 1 (ratioL_t) -> sweep 2 (ratioR_t, remainR) -> sweep 3 (remainL); sweep 3 of level t-1 and
  // sweep 1 of level t share one fused launch.
  auto fL = [&](int t) { return ws.fL + (size_t)t * ws.fL_level_stride; };
  auto fR = [&](int t) { return ws.fR + (size_t)t * ws.fR_level_stride; };
  const int ncull = am_cull_levels();
  for (int t = 0; t < AM_LEVELS; ++t) {
    const bool cull_t = cull && t < ncull;
    if (t == 0 && cull_t) {
      am_sweep_cull_kernel<EPI_RATIO_L><<<gk, AM_THREADS, AMC_SMEM1, st>>>(n, m, xyz1, xyz2, remainR, tstride, lv.lv[0],
                                                                           cut(0), ws.perm1, remainL, tstride, fL(0),
                                                                           (size_t)n);
    } else if (t == 0) {
      PCC_LAUNCH(PDL_EMD_SWEEP, PCC_K(am_sweep_kernel<EPI_RATIO_L, P>), gk, AM_THREADS, 0, st, n, m, xyz1, xyz2, remainR,
                 tstride, lv.lv[0], remainL, tstride, fL(0), (size_t)n);
    } else if (cull_t) {
      am_sweep31_cull_kernel<<<gk, AM_THREADS, AMC_SMEM31, st>>>(n, m, xyz1, xyz2, fR(t - 1), (size_t)m, remainR, tstride,
                                                                 lv.lv[t - 1], lv.lv[t], cut(t), ws.perm1, remainL,
                                                                 tstride, fL(t - 1), fL(t), (size_t)n);
    } else {
      PCC_LAUNCH(PDL_EMD_SWEEP, am_sweep31_kernel<P>, gk, AM_THREADS, 0, st, n, m, xyz1, xyz2, fR(t - 1), (size_t)m,
                 remainR, tstride, lv.lv[t - 1], lv.lv[t], remainL, tstride, fL(t - 1), fL(t), (size_t)n);
    }
    if (cull_t)
      am_sweep_cull_kernel<EPI_RATIO_R><<<gl, AM_THREADS, AMC_SMEM1, st>>>(m, n, xyz2, xyz1, fL(t), (size_t)n, lv.lv[t],
                                                                           cut(t), ws.perm2, remainR, tstride, fR(t),
                                                                           (size_t)m);
    else
      PCC_LAUNCH(PDL_EMD_SWEEP, PCC_K(am_sweep_kernel<EPI_RATIO_R, P>), gl, AM_THREADS, 0, st, m, n, xyz2, xyz1, fL(t),
                 (size_t)n, lv.lv[t], remainR, tstride, fR(t), (size_t)m);
  }
  // The last remainL update (sweep 3 of level j = -1, approxmatch.cu:130-163) and the export of the last level's ratio
  // vectors only fill `temp`: match / cost / gradients are functions of the per-level ratio vectors alone.  The
  // fused path treats temp as scratch (its Python caller drops it, match_cost.py:25), so it skips both.
  if (want_temp) {
    PCC_LAUNCH(PDL_EMD_SWEEP, PCC_K(am_sweep_kernel<EPI_REMAIN_L, P>), gk, AM_THREADS, 0, st, n, m, xyz1, xyz2,
               fR(AM_LEVELS - 1), (size_t)m, lv.lv[AM_LEVELS - 1], remainL, tstride, fL(AM_LEVELS - 1), (size_t)n);
    PCC_LAUNCH(PDL_EMD_SMALL, am_export_temp_kernel, dim3((n + m + 255) / 256, b), 256, 0, st, n, m,
               ws.fL + (size_t)(AM_LEVELS - 1) * ws.fL_level_stride,
               ws.fR + (size_t)(AM_LEVELS - 1) * ws.fR_level_stride, temp);
    *launches += 2;
  }
  *launches += 2 * AM_LEVELS + 1;
  return (int)cudaGetLastError();
}

}  // namespace pcc

using namespace pcc;

extern "C" __attribute__((visibility("default"))) int pcc_approxmatch(int b, int n, int m, const float *xyz1, const float *xyz2, float *match, float *temp,
                               pcc_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return PCC_EINVAL;
  if (b == 0 || n == 0 || m == 0) return PCC_OK;
  if (b > 65535) return PCC_ENOTSUP;
  cudaStream_t st = (cudaStream_t)stream;
  AmWorkspace ws;
  AmLevels lv;
  am_levels(st, &lv);
  int launches = 0;
  int rc = am_solve(b, n, m, xyz1, xyz2, temp, ws, 0, lv, st, &launches, true);
  if (rc == 0) {
    const int lslab = 256;
    dim3 grid((n + AM_THREADS - 1) / AM_THREADS, (m + lslab - 1) / lslab, b);
    am_materialize_kernel<<<grid, AM_THREADS, 0, st>>>(n, m, lslab, xyz1, xyz2, ws.fL, ws.fR, ws.fL_level_stride,
                                                       ws.fR_level_stride, lv, match);
    ++launches;
    rc = (int)cudaGetLastError();
  }
  if (ws.base) cudaFreeAsync(ws.base, st);
  g_launches.fetch_add((uint64_t)launches, std::memory_order_relaxed);
  return rc;
}

extern "C" __attribute__((visibility("default"))) int pcc_matchcost_fused(int b, int n, int m, const float *xyz1, const float *xyz2, float *cost,
                                   float *grad1, float *grad2, float *temp, pcc_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return PCC_EINVAL;
  if (b == 0) return PCC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0 || m == 0) {
    if (cost) cudaMemsetAsync(cost, 0, sizeof(float) * b, st);
    if (grad1 && n) cudaMemsetAsync(grad1, 0, sizeof(float) * (size_t)b * n * 3, st);
    if (grad2 && m) cudaMemsetAsync(grad2, 0, sizeof(float) * (size_t)b * m * 3, st);
    return (int)cudaGetLastError();
  }
  if (b > 65535) return PCC_ENOTSUP;
  AmWorkspace ws;
  AmLevels sc;
  am_levels(st, &sc);
  int launches = 0;
  const int parts = (n + AM_THREADS - 1) / AM_THREADS;
  int rc = am_solve(b, n, m, xyz1, xyz2, temp, ws, (size_t)b * parts, sc, st, &launches, false);
  if (rc == 0) {
    PCC_LAUNCH(PDL_EMD_SWEEP, am_costgrad_kernel, dim3(parts, b), AM_THREADS, 0, st, n, m, xyz1, xyz2, ws.fL, ws.fR,
               ws.fL_level_stride, ws.fR_level_stride, sc, cost ? ws.cost_part : nullptr, grad1);
    ++launches;
    if (cost) {
      PCC_LAUNCH(PDL_EMD_SMALL, am_cost_reduce_kernel, (b + 127) / 128, 128, 0, st, b, parts, ws.cost_part, cost);
      ++launches;
    }
    if (grad2) {
      PCC_LAUNCH(PDL_EMD_SWEEP, am_costgrad_kernel, dim3((m + AM_THREADS - 1) / AM_THREADS, b), AM_THREADS, 0, st, m, n,
                 xyz2, xyz1, ws.fR, ws.fL, ws.fR_level_stride, ws.fL_level_stride, sc, (float *)nullptr, grad2);
      ++launches;
    }
    rc = (int)cudaGetLastError();
  }
  if (ws.base) cudaFreeAsync(ws.base, st);
  g_launches.fetch_add((uint64_t)launches, std::memory_order_relaxed);
  return rc;
}

extern "C" __attribute__((visibility("default"))) int pcc_matchcost(int b, int n, int m, const float *xyz1, const float *xyz2, const float *match,
                             float *out, pcc_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return PCC_EINVAL;
  if (b == 0) return PCC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0 || m == 0) {
    cudaMemsetAsync(out, 0, sizeof(float) * b, st);
    return (int)cudaGetLastError();
  }
  if (b > 65535) return PCC_ENOTSUP;
  const int parts = (m + MC_ROWS - 1) / MC_ROWS;
  float *partial = nullptr;
  cudaError_t e = ws_alloc((void **)&partial, sizeof(float) * (size_t)b * parts, st);
  if (e != cudaSuccess) return (int)e;
  matchcost_kernel<<<dim3(parts, b), MC_THREADS, 0, st>>>(n, m, xyz1, xyz2, match, partial);
  am_cost_reduce_kernel<<<(b + 127) / 128, 128, 0, st>>>(b, parts, partial, out);
  cudaFreeAsync(partial, st);
  return finish_launch(2);
}

extern "C" __attribute__((visibility("default"))) int pcc_matchcostgrad(int b, int n, int m, const float *xyz1, const float *xyz2, const float *match,
                                 float *grad1, float *grad2, pcc_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return PCC_EINVAL;
  if (b == 0) return PCC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0 || m == 0) {
    if (n) cudaMemsetAsync(grad1, 0, sizeof(float) * (size_t)b * n * 3, st);
    if (m) cudaMemsetAsync(grad2, 0, sizeof(float) * (size_t)b * m * 3, st);
    return (int)cudaGetLastError();
  }
  if (b > 65535) return PCC_ENOTSUP;
  matchcostgrad1_kernel<<<dim3((n + MC_THREADS - 1) / MC_THREADS, b), MC_THREADS, 0, st>>>(n, m, xyz1, xyz2, match,
                                                                                          grad1);
  matchcostgrad2_kernel<<<dim3((m + MC_THREADS / 32 - 1) / (MC_THREADS / 32), b), MC_THREADS, 0, st>>>(n, m, xyz1, xyz2,
                                                                                                        match, grad2);
  return finish_launch(2);
}

template <int P, int U = 16>
static void launch_sweep_variant(int b, int n, int m, const float *xyz1, const float *xyz2, const float *weight,
                                 const float *remain, float *ratio, float level, cudaStream_t st) {
  const int per_cta = AM_THREADS * P;
  am_sweep_kernel<EPI_RATIO_L, P, U><<<dim3((n + per_cta - 1) / per_cta, b), AM_THREADS, 0, st>>>(
      n, m, xyz1, xyz2, weight, (size_t)m, level, const_cast<float *>(remain), (size_t)n, ratio, (size_t)n);
}

extern "C" __attribute__((visibility("default"))) int pcc_approxmatch_sweep(int b, int n, int m, const float *xyz1,
                                                                            const float *xyz2, const float *weight,
                                                                            const float *remain, float *ratio,
                                                                            float level, int points_per_thread,
                                                                            pcc_stream_t stream) {
  if (b <= 0 || n <= 0 || m <= 0) return PCC_EINVAL;
  if (b > 65535) return PCC_ENOTSUP;
  cudaStream_t st = (cudaStream_t)stream;
  switch (points_per_thread) {
    case 0: launch_sweep_variant<AM_P_DEFAULT>(b, n, m, xyz1, xyz2, weight, remain, ratio, level, st); break;
    case 2: launch_sweep_variant<2>(b, n, m, xyz1, xyz2, weight, remain, ratio, level, st); break;
    case 4: launch_sweep_variant<4>(b, n, m, xyz1, xyz2, weight, remain, ratio, level, st); break;
    case 102: launch_sweep_variant<2, 8>(b, n, m, xyz1, xyz2, weight, remain, ratio, level, st); break;
    case 202: launch_sweep_variant<2, 32>(b, n, m, xyz1, xyz2, weight, remain, ratio, level, st); break;
    case 104: launch_sweep_variant<4, 8>(b, n, m, xyz1, xyz2, weight, remain, ratio, level, st); break;
    default: return PCC_ENOTSUP;
  }
  return finish_launch(1);
}
