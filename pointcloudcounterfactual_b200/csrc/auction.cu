// Auction-algorithm EMD for sm_100a: one persistent thread-block cluster (4 CTAs) per cloud runs all rounds
// (auction_cluster_kernel below; auction_kernel is the one-CTA-per-cloud version it grew out of).
//
// Replaces external/emd/src/emd_cuda.cu: the reference launches 7 kernels per round (clear, calc_unass_cnt,
// calc_unass_cnt_sum, calc_unass_idx, Bid, GetMax, Assign; :255-268) on the default stream, i.e. 350 launches for the
// recommended 50 training rounds and 70 000 for the 10 000 test rounds.  Here every cloud is owned by one
// 1024-thread CTA that keeps the target coordinates and prices in shared memory and loops over the rounds with
// __syncthreads() in between -- one launch in total, on the caller's stream.
//
// Semantics follow emd_cuda.cu:94-225 round for round (synchronous bids on start-of-round prices, one winner per
// target, forced assignment on the last round).  Where the reference is racy the result here is deterministic:
// among bids within 1e-6 of the maximum the HIGHEST source index wins (reference: last writer, :187-190); the
// compacted list of unassigned sources is in ascending order (reference: atomic order, :84-92).
#include <cstdlib>

#include "common.cuh"

namespace pcc {

constexpr int AU_THREADS = 1024;

// order-independent merge of (best value, lowest index attaining it, second-best value of the multiset)
struct Bid3 {
  float best, better;
  int idx;
};
__device__ __forceinline__ Bid3 bid_merge(Bid3 a, Bid3 b) {
  Bid3 r;
  const bool a_wins = (a.best > b.best) || (a.best == b.best && (unsigned)a.idx < (unsigned)b.idx);
  r.best = a_wins ? a.best : b.best;
  r.idx = a_wins ? a.idx : b.idx;
  r.better = fmaxf(fminf(a.best, b.best), fmaxf(a.better, b.better));
  return r;
}

__global__ void __launch_bounds__(AU_THREADS, 1)
auction_kernel(int n, const float *__restrict__ xyz1, const float *__restrict__ xyz2, float *__restrict__ dist,
               int *__restrict__ assignment, float *__restrict__ price_g, int *__restrict__ assignment_inv,
               int *__restrict__ bid, float *__restrict__ bid_inc, float *__restrict__ max_inc_g,
               int *__restrict__ unass_idx, int *__restrict__ max_idx, float eps, int iters) {
  extern __shared__ float sm[];
  float *sx = sm, *sy = sm + n, *sz = sm + 2 * n;  // target coordinates, SoA
  float *price = sm + 3 * n;
  int *max_inc = reinterpret_cast<int *>(sm + 4 * n);  // float bits; bids are positive so signed-int max == float max
  __shared__ int warp_cnt[AU_THREADS / 32];
  __shared__ int n_unass;

  const size_t cloud = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  xyz1 += cloud * (size_t)n * 3;
  xyz2 += cloud * (size_t)n * 3;
  int *asg = assignment + cloud * (size_t)n, *inv = assignment_inv + cloud * (size_t)n;
  int *bd = bid + cloud * (size_t)n, *ulist = unass_idx + cloud * (size_t)n, *mxi = max_idx + cloud * (size_t)n;
  float *binc = bid_inc + cloud * (size_t)n;

  for (int k = tid; k < n; k += AU_THREADS) {
    sx[k] = xyz2[k * 3];
    sy[k] = xyz2[k * 3 + 1];
    sz[k] = xyz2[k * 3 + 2];
    price[k] = price_g[cloud * (size_t)n + k];
    max_inc[k] = __float_as_int(max_inc_g[cloud * (size_t)n + k]);
  }
  __syncthreads();

  const int per_thread = n / AU_THREADS;  // n is a multiple of 1024 (checked on the host, emd_cuda.cu:245-248)
  for (int it = 0; it < iters; ++it) {
    const bool last = (it == iters - 1);
    // ---- ordered compaction of the unassigned sources (emd_cuda.cu:29-92) ---------------------------------
    // thread t owns sources [t*per_thread, (t+1)*per_thread): ascending order is preserved
    int mine = 0;
    for (int e = 0; e < per_thread; ++e) mine += (asg[tid * per_thread + e] == -1);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_cnt[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int t = warp_cnt[lane], s = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += u;
      }
      warp_cnt[lane] = s - t;
      if (lane == 31) n_unass = s;
    }
    __syncthreads();
    int pos = warp_cnt[warp] + incl - mine;
    for (int e = 0; e < per_thread; ++e) {
      const int j = tid * per_thread + e;
      if (asg[j] == -1) ulist[pos++] = j;
    }
    __syncthreads();
    const int U = n_unass;
    if (U == 0) break;  // uniform: nothing left to assign, later rounds are no-ops (emd_cuda.cu:103-104)

    // ---- Bid (emd_cuda.cu:94-178): `tpb` lanes cooperate on one bidder, strided over the targets -----------
    int tpb = 1;
    while (tpb < 32 && U * tpb * 2 <= AU_THREADS) tpb <<= 1;
    const int bidders_per_pass = AU_THREADS / tpb;
    for (int u0 = 0; u0 < U; u0 += bidders_per_pass) {
      const int u = u0 + tid / tpb;
      const int sub = tid % tpb;
      Bid3 r;
      r.best = -1e9f;
      r.better = -1e9f;
      r.idx = -1;
      int i = -1;
      if (u < U) {
        i = ulist[u];
        const float x1 = xyz1[i * 3], y1 = xyz1[i * 3 + 1], z1 = xyz1[i * 3 + 2];
        for (int k = sub; k < n; k += tpb) {
          const float x2 = sx[k] - x1, y2 = sy[k] - y1, z2 = sz[k] - z1;
          const float s = __fmaf_rn(z2, z2, __fmaf_rn(x2, x2, __fmul_rn(y2, y2)));
          // `3.0 - sqrtf(.) - price` with a double literal is evaluated in double by the reference (:145)
          const float d = (float)(3.0 - (double)sqrtf(s) - (double)price[k]);
          if (d > r.best) {
            r.better = r.best;
            r.best = d;
            r.idx = k;
          } else if (d > r.better) {
            r.better = d;
          }
        }
      }
      for (int o = 1; o < tpb; o <<= 1) {  // groups are aligned sub-warps
        Bid3 q;
        q.best = __shfl_xor_sync(0xffffffffu, r.best, o);
        q.better = __shfl_xor_sync(0xffffffffu, r.better, o);
        q.idx = __shfl_xor_sync(0xffffffffu, r.idx, o);
        r = bid_merge(r, q);
      }
      if (u < U && sub == 0) {
        // a source whose every value is NaN (diverged decoder output) never beats the -1e9 start: r.idx stays -1.  Bid on
        // target 0 with the minimum increment instead of indexing mig[-1] (memory safety; the loss is NaN either way)
        if (r.idx < 0) {
          r.idx = 0;
          r.best = r.better = 0.f;
        }
        const float inc = r.best - r.better + eps;
        bd[i] = r.idx;
        binc[i] = inc;
        atomicMax(&max_inc[r.idx], __float_as_int(inc));  // emd_cuda.cu:9-19,175
      }
    }
    __syncthreads();
    // ---- GetMax (emd_cuda.cu:180-193): winner = highest source index among the bids within 1e-6 of the
    //      maximum; on the forced last round every bidder takes part -----------------------------------------
    for (int u = tid; u < U; u += AU_THREADS) {
      const int i = ulist[u], t = bd[i];
      const double bi = (double)binc[i], mi = (double)__int_as_float(max_inc[t]);
      if (last || (bi - 1e-6 <= mi && mi <= bi + 1e-6)) mxi[t] = -1;  // drop the stale winner of earlier rounds
    }
    __syncthreads();
    for (int u = tid; u < U; u += AU_THREADS) {
      const int i = ulist[u], t = bd[i];
      const double bi = (double)binc[i], mi = (double)__int_as_float(max_inc[t]);
      if (last || (bi - 1e-6 <= mi && mi <= bi + 1e-6)) atomicMax(&mxi[t], i);
    }
    __syncthreads();
    // ---- Assign (emd_cuda.cu:195-214) ---------------------------------------------------------------------
    for (int u = tid; u < U; u += AU_THREADS) {
      const int i = ulist[u], t = bd[i];
      if (last || mxi[t] == i) {
        const float inc = binc[i];
        if (!last) {
          const int prev = inv[t];
          if (prev != -1) asg[prev] = -1;
          inv[t] = i;
          price[t] += inc;
        } else {  // several sources may be forced onto one target (:200): all keep it, the highest owns inv
          if (mxi[t] == i) inv[t] = i;
          atomicAdd(&price[t], inc);
        }
        asg[i] = t;
        max_inc[t] = __float_as_int(-1e9f);
      }
    }
    __syncthreads();
  }

  // ---- CalcDist (emd_cuda.cu:216-225) + state write-back -------------------------------------------------------
  for (int j = tid; j < n; j += AU_THREADS) {
    const int k = asg[j];
    float d = 0.f;
    if (k >= 0) {
      const float dx = xyz1[j * 3] - sx[k], dy = xyz1[j * 3 + 1] - sy[k], dz = xyz1[j * 3 + 2] - sz[k];
      d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
    }
    dist[cloud * (size_t)n + j] = d;
    price_g[cloud * (size_t)n + j] = price[j];
    max_inc_g[cloud * (size_t)n + j] = __int_as_float(max_inc[j]);
  }
}

// ---- the same rounds on a thread-block CLUSTER of AUC_CTAS CTAs per cloud ---------------------------------------------
// One CTA per cloud keeps 116 of the 148 SMs idle at the usual batch of 32 clouds, and the Bid scan (unassigned sources
// x all targets, with the reference's double-precision value expression) is pure arithmetic.  Here AUC_CTAS CTAs of one
// cluster share a cloud: every CTA holds the target coordinates and a per-round copy of the prices in its own shared
// memory, the bidders of a round are dealt out over all AUC_CTAS * 1024 threads, and everything the CTAs exchange
// (assignment, bids, increments, winners, prices) lives in the caller's global work buffers -- read with ld.global.cg,
// ordered by the cluster barrier.  Same arithmetic, same order-independent merges and the same tie rules as
// auction_kernel: identical results.
constexpr int AUC_CTAS = 4;

__device__ __forceinline__ void cluster_sync_all() {
  __threadfence();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned int cluster_cta_rank() {
  unsigned int r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__global__ void __cluster_dims__(AUC_CTAS, 1, 1) __launch_bounds__(AU_THREADS, 1)
auction_cluster_kernel(int n, const float *__restrict__ xyz1, const float *__restrict__ xyz2, float *dist,
                       int *assignment, float *price_g, int *assignment_inv, int *bid, float *bid_inc,
                       float *max_inc_g, int *unass_idx, int *max_idx, int *cnt_g, float eps, int iters) {
  extern __shared__ float sm[];
  float *sx = sm, *sy = sm + n, *sz = sm + 2 * n;  // target coordinates, SoA
  float *price = sm + 3 * n;                       // this round's prices (copy of price_g)
  __shared__ int warp_cnt[AU_THREADS / 32];

  const size_t cloud = blockIdx.y;
  const int rank = (int)cluster_cta_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gtid = rank * AU_THREADS + tid;
  constexpr int GT = AUC_CTAS * AU_THREADS;
  xyz1 += cloud * (size_t)n * 3;
  xyz2 += cloud * (size_t)n * 3;
  int *asg = assignment + cloud * (size_t)n, *inv = assignment_inv + cloud * (size_t)n;
  int *bd = bid + cloud * (size_t)n, *ulist = unass_idx + cloud * (size_t)n, *mxi = max_idx + cloud * (size_t)n;
  float *binc = bid_inc + cloud * (size_t)n, *pg = price_g + cloud * (size_t)n;
  int *mig = reinterpret_cast<int *>(max_inc_g + cloud * (size_t)n);  // float bits: signed-int max == float max for bids > 0
  int *cnt = cnt_g + cloud * (size_t)AUC_CTAS;                        // unassigned sources per CTA of the cluster

  for (int k = tid; k < n; k += AU_THREADS) {
    sx[k] = xyz2[k * 3];
    sy[k] = xyz2[k * 3 + 1];
    sz[k] = xyz2[k * 3 + 2];
  }
  const int per = (n + GT - 1) / GT;           // sources owned by one thread: [gtid*per, gtid*per + per)
  const int j_begin = min(gtid * per, n), j_end = min(j_begin + per, n);
  for (int it = 0; it < iters; ++it) {
    const bool last = (it == iters - 1);
    // ---- ordered compaction of the unassigned sources over the whole cluster (emd_cuda.cu:29-92) -----------------
    int mine = 0;
    for (int j = j_begin; j < j_end; ++j) mine += (__ldcg(&asg[j]) == -1);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_cnt[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int t = warp_cnt[lane], s2 = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, s2, o);
        if (lane >= o) s2 += u;
      }
      warp_cnt[lane] = s2 - t;
      if (lane == 31) cnt[rank] = s2;
    }
    for (int k = tid; k < n; k += AU_THREADS) price[k] = __ldcg(&pg[k]);  // prices at the start of the round
    cluster_sync_all();
    int base = 0, U = 0;
#pragma unroll
    for (int r = 0; r < AUC_CTAS; ++r) {
      const int c = __ldcg(&cnt[r]);
      if (r < rank) base += c;
      U += c;
    }
    if (U == 0) break;  // cluster-uniform: nothing left to assign (emd_cuda.cu:103-104)
    int pos = base + warp_cnt[warp] + incl - mine;
    for (int j = j_begin; j < j_end; ++j)
      if (__ldcg(&asg[j]) == -1) ulist[pos++] = j;
    cluster_sync_all();

    // ---- Bid (emd_cuda.cu:94-178): `tpb` lanes cooperate on one bidder, strided over the targets; the bidders of
    //      a pass are spread over all CTAs of the cluster ---------------------------------------------------------
    int tpb = 1;
    while (tpb < 32 && U * tpb * 2 <= GT) tpb <<= 1;
    const int bidders_per_pass = GT / tpb;
    for (int u0 = 0; u0 < U; u0 += bidders_per_pass) {
      const int u = u0 + gtid / tpb;
      const int sub = gtid % tpb;
      Bid3 r;
      r.best = -1e9f;
      r.better = -1e9f;
      r.idx = -1;
      int i = -1;
      if (u < U) {
        i = __ldcg(&ulist[u]);
        const float x1 = xyz1[i * 3], y1 = xyz1[i * 3 + 1], z1 = xyz1[i * 3 + 2];
        for (int k = sub; k < n; k += tpb) {
          const float x2 = sx[k] - x1, y2 = sy[k] - y1, z2 = sz[k] - z1;
          const float s2 = __fmaf_rn(z2, z2, __fmaf_rn(x2, x2, __fmul_rn(y2, y2)));
          // `3.0 - sqrtf(.) - price` with a double literal is evaluated in double by the reference (:145)
          const float d = (float)(3.0 - (double)sqrtf(s2) - (double)price[k]);
          if (d > r.best) {
            r.better = r.best;
            r.best = d;
            r.idx = k;
          } else if (d > r.better) {
            r.better = d;
          }
        }
      }
      for (int o = 1; o < tpb; o <<= 1) {  // groups are aligned sub-warps
        Bid3 q;
        q.best = __shfl_xor_sync(0xffffffffu, r.best, o);
        q.better = __shfl_xor_sync(0xffffffffu, r.better, o);
        q.idx = __shfl_xor_sync(0xffffffffu, r.idx, o);
        r = bid_merge(r, q);
      }
      if (u < U && sub == 0) {
        // a source whose every value is NaN (diverged decoder output) never beats the -1e9 start: r.idx stays -1.  Bid on
        // target 0 with the minimum increment instead of indexing mig[-1] (memory safety; the loss is NaN either way)
        if (r.idx < 0) {
          r.idx = 0;
          r.best = r.better = 0.f;
        }
        const float inc = r.best - r.better + eps;
        bd[i] = r.idx;
        binc[i] = inc;
        atomicMax(&mig[r.idx], __float_as_int(inc));  // emd_cuda.cu:9-19,175
      }
    }
    cluster_sync_all();
    // ---- GetMax (emd_cuda.cu:180-193): winner = highest source index among the bids within 1e-6 of the maximum;
    //      on the forced last round every bidder takes part ------------------------------------------------------
    for (int u = gtid; u < U; u += GT) {
      const int i = __ldcg(&ulist[u]), t = __ldcg(&bd[i]);
      const double bi = (double)__ldcg(&binc[i]), mi = (double)__int_as_float(__ldcg(&mig[t]));
      if (last || (bi - 1e-6 <= mi && mi <= bi + 1e-6)) mxi[t] = -1;  // drop the stale winner of earlier rounds
    }
    cluster_sync_all();
    for (int u = gtid; u < U; u += GT) {
      const int i = __ldcg(&ulist[u]), t = __ldcg(&bd[i]);
      const double bi = (double)__ldcg(&binc[i]), mi = (double)__int_as_float(__ldcg(&mig[t]));
      if (last || (bi - 1e-6 <= mi && mi <= bi + 1e-6)) atomicMax(&mxi[t], i);
    }
    cluster_sync_all();
    // ---- Assign (emd_cuda.cu:195-214) ---------------------------------------------------------------------------
    for (int u = gtid; u < U; u += GT) {
      const int i = __ldcg(&ulist[u]), t = __ldcg(&bd[i]);
      const int win = __ldcg(&mxi[t]);
      if (last || win == i) {
        const float inc = __ldcg(&binc[i]);
        if (!last) {
          const int prev = __ldcg(&inv[t]);
          if (prev != -1) asg[prev] = -1;
          inv[t] = i;
          pg[t] = price[t] + inc;  // one winner per target: no race
        } else {  // several sources may be forced onto one target (:200): all keep it, the highest owns inv
          if (win == i) inv[t] = i;
          atomicAdd(&pg[t], inc);
        }
        asg[i] = t;
        mig[t] = __float_as_int(-1e9f);
      }
    }
    cluster_sync_all();
  }

  // ---- CalcDist (emd_cuda.cu:216-225); prices and increments already live in the caller's buffers ------------------
  for (int j = gtid; j < n; j += GT) {
    const int k = __ldcg(&asg[j]);
    float d = 0.f;
    if (k >= 0) {
      const float dx = xyz1[j * 3] - sx[k], dy = xyz1[j * 3 + 1] - sy[k], dz = xyz1[j * 3 + 2] - sz[k];
      d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
    }
    dist[cloud * (size_t)n + j] = d;
  }
}

__global__ void auction_grad_kernel(size_t total, int n, const float *__restrict__ xyz1,
                                    const float *__restrict__ xyz2, const float *__restrict__ gdist,
                                    const int *__restrict__ idx, float *__restrict__ grad) {
  // emd_cuda.cu:283-299: grad1 = 2 g (x1 - x2[assignment])   (grad2 stays zero, emd_module.py:75-79)
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t i = t / n;
    const int k = min(max(idx[t], 0), n - 1);
    const float g = gdist[t] * 2.f;
    const float *p = xyz1 + t * 3, *q = xyz2 + (i * n + k) * 3;
    grad[t * 3] = g * (p[0] - q[0]);
    grad[t * 3 + 1] = g * (p[1] - q[1]);
    grad[t * 3 + 2] = g * (p[2] - q[2]);
  }
}

}  // namespace pcc

using namespace pcc;

extern "C" __attribute__((visibility("default"))) int pcc_emd_forward(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist,
                               int *assignment, float *price, int *assignment_inv, int *bid, float *bid_increments,
                               float *max_increments, int *unass_idx, int *unass_cnt, int *unass_cnt_sum,
                               int *cnt_tmp, int *max_idx, float eps, int iters, pcc_stream_t stream) {
  (void)unass_cnt_sum;  // of the three 512-int counters of the reference API only unass_cnt is used (cluster kernel)
  (void)cnt_tmp;
  if (n != m) return -1;          // emd_cuda.cu:235-238
  if (b > 512) return -1;         // :240-243
  if (n % 1024 != 0) return -1;   // :245-248
  if (b <= 0 || n == 0) return 1;
  const size_t smem = sizeof(float) * 5 * (size_t)n;
  if (smem > 220 * 1024) return PCC_ENOTSUP;  // n <= 11264
  static size_t attr[64];
  if (cudaError_t e = smem_optin(auction_kernel, smem, attr); e != cudaSuccess) return (int)e;
  static const bool single_cta = getenv("PCC_AUCTION_SINGLE_CTA") != nullptr;  // test hook: one CTA per cloud
  if (!single_cta && b <= 65535) {
    // clusters of AUC_CTAS CTAs per cloud; unass_cnt (512 ints, caller-zeroed) carries the per-CTA counts for up to
    // 512 / AUC_CTAS clouds, more clouds take a pool allocation
    const size_t csmem = sizeof(float) * 4 * (size_t)n;
    static size_t cattr[64];
    if (cudaError_t e = smem_optin(auction_cluster_kernel, csmem, cattr); e != cudaSuccess) return (int)e;
    int *cnt = unass_cnt;
    int *pool = nullptr;
    if (cnt == nullptr || b * AUC_CTAS > 512) {
      cudaError_t e = ws_alloc((void **)&pool, sizeof(int) * (size_t)b * AUC_CTAS, (cudaStream_t)stream);
      if (e != cudaSuccess) return (int)e;
      cnt = pool;
    }
    auction_cluster_kernel<<<dim3(AUC_CTAS, b), AU_THREADS, csmem, (cudaStream_t)stream>>>(
        n, xyz1, xyz2, dist, assignment, price, assignment_inv, bid, bid_increments, max_increments, unass_idx, max_idx,
        cnt, eps, iters);
    if (pool) cudaFreeAsync(pool, (cudaStream_t)stream);
    const int rc = finish_launch(1);
    return rc == 0 ? 1 : (rc == 1 ? PCC_ELAUNCH : rc);  // 1 means success here: keep cudaErrorInvalidValue apart
  }
  auction_kernel<<<b, AU_THREADS, smem, (cudaStream_t)stream>>>(n, xyz1, xyz2, dist, assignment, price,
                                                                 assignment_inv, bid, bid_increments, max_increments,
                                                                 unass_idx, max_idx, eps, iters);
  const int rc = finish_launch(1);
  return rc == 0 ? 1 : (rc == 1 ? PCC_ELAUNCH : rc);  // 1 means success here: keep cudaErrorInvalidValue apart
}

extern "C" __attribute__((visibility("default"))) int pcc_emd_backward(int b, int n, const float *xyz1, const float *xyz2, float *gradxyz,
                                const float *graddist, const int *idx, pcc_stream_t stream) {
  if (b <= 0 || n <= 0) return 1;
  const size_t total = (size_t)b * n;
  const int threads = 256;
  size_t blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 32) blocks = 148 * 32;
  auction_grad_kernel<<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(total, n, xyz1, xyz2, graddist, idx, gradxyz);
  const int rc = finish_launch(1);
  return rc == 0 ? 1 : (rc == 1 ? PCC_ELAUNCH : rc);  // 1 means success here: keep cudaErrorInvalidValue apart
}
