// EdgeConv front-end: neighbour gather and graph features for sm_100a, forward and backward.
//
// Replaces the torch composition behind src/utils/neighbour_ops.py:85-94 (get_neighbours: view/expand + torch.gather)
// and :113-119 (get_graph_features: gather, expand, subtract, cat, contiguous -- three (B,C,N,k) temporaries and one
// (B,2C,N,k) result, the tensors that dominate the encoder's memory traffic) with one pass that writes
// the result once:
//   mode 0  out (B,C,N,k):   out[c][i][t] = x[c][idx[i][t]]
//   mode 1  out (B,2C,N,k):  out[c][i][t] = x[c][idx[i][t]] - x[c][i],   out[C+c][i][t] = x[c][i]
// Both kernels are HBM-bound by construction: the forward writes the result exactly once (rows of x are staged in
// shared memory, 8 KB per channel at N = 2048, and gathered from there), the backward reads the upstream gradient
// exactly once.  Two backward kernels:
//   * sorted (default while one (point, neighbour) plane of the gradient, n*k floats, fits shared memory): the edges of
//     every cloud are sorted by target once per call (edgeconv.cu's edge_sort_launch), the plane of one (cloud, channel)
//     is brought into shared memory with one bulk copy, and every thread sums GS_E consecutive entries of the sorted list
//     in list order.  No atomics anywhere; the order of every addition is a function of idx alone, so the result is
//     bitwise reproducible (torch.gather's own backward is not).
//   * atomic (planes that do not fit, or n*k not a multiple of 4): scatter-add into a shared-memory row with float
//     atomics -- 511 us at C=64, N=2048, k=25, B=32, bound by the 105 M shared-memory atomics (1.4 cycles per lane).
// Measured at that shape: forward 195 us (66 % of the HBM copy peak); sorted backward 271 us (50 us of it the sort), i.e.
// 3.1 TB/s of gradient read -- 256 us when the caller runs the sort early (pcc_graph_edge_sort on a side stream under the
// forward, then pcc_graph_gather_grad_presorted); N=1024, k=20: 124 us.  Measured and dropped: (a) segmented sums with warp shuffles per 32
// edges before one atomic per run, 795 us; (b) the sorted kernel with the source range cut into four pieces and two
// alternating shared-memory stages so that copies overlap sums inside one CTA, 335 us -- the loads disappear from the
// stall profile but four times the run boundaries, eight barriers and the per-piece continuation pass cost more
// instructions (153 M against 110 M warp-level) than the overlap returns.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace pcc {

constexpr int GG_THREADS = 256;
constexpr int GG_PT = 128;   // points per CTA
constexpr int GG_CH = 4;     // channels staged per step
constexpr int GG_ITEMS = 16; // (point, neighbour) items per thread: GG_PT * 32 / GG_THREADS
constexpr int GG_MAXN = 8192;  // GG_CH rows of this many floats fit the dynamic shared memory

template <int MODE>
__global__ void __launch_bounds__(GG_THREADS)
graph_gather_kernel(int c, int n, int k, const float *__restrict__ x, const int64_t *__restrict__ idx,
                    float *__restrict__ out) {
  extern __shared__ __align__(16) float rows[];  // [GG_CH][n]
  const size_t cloud = blockIdx.y;
  const int i0 = blockIdx.x * GG_PT;
  const int items = min(GG_PT, n - i0) * k;
  const float *xb = x + cloud * (size_t)c * n;
  const int64_t *ib = idx + (cloud * (size_t)n + i0) * k;
  const int co = MODE ? 2 * c : c;
  float *ob = out + cloud * (size_t)co * n * k + (size_t)i0 * k;
  // this thread's items: neighbour index and own point, fixed for all channels
  int nbr[GG_ITEMS], own[GG_ITEMS];
#pragma unroll
  for (int it = 0; it < GG_ITEMS; ++it) {
    const int o = threadIdx.x + it * GG_THREADS;
    nbr[it] = 0;
    own[it] = 0;
    if (o < items) {
      const long long j = ib[o];
      nbr[it] = (int)min(max(j, 0LL), (long long)n - 1);  // memory-safe for invalid indices
      own[it] = i0 + o / k;
    }
  }
  for (int c0 = 0; c0 < c; c0 += GG_CH) {
    const int cn = min(GG_CH, c - c0);
    __syncthreads();
    for (int e = threadIdx.x; e < cn * n; e += GG_THREADS) rows[e] = xb[(size_t)c0 * n + e];  // rows are contiguous
    __syncthreads();
    for (int cc = 0; cc < cn; ++cc) {
      const float *row = rows + cc * n;
      float *o_top = ob + (size_t)(c0 + cc) * n * k;
      float *o_bot = ob + (size_t)(c + c0 + cc) * n * k;
#pragma unroll
      for (int it = 0; it < GG_ITEMS; ++it) {
        const int o = threadIdx.x + it * GG_THREADS;
        if (o < items) {
          const float xj = row[nbr[it]];
          if (MODE) {
            const float xi = row[own[it]];
            o_top[o] = xj - xi;
            o_bot[o] = xi;
          } else {
            o_top[o] = xj;
          }
        }
      }
    }
  }
}

// one CTA per (cloud, channel): grad_x[c][j] = sum over (i,t) with idx[i][t] == j of g_top[i][t]
//                               (+ mode 1:  sum_t (g_bot[i][t] - g_top[i][t]) for the point itself)
template <int MODE>
__global__ void __launch_bounds__(GG_THREADS)
graph_gather_grad_kernel(int c, int n, int k, const int64_t *__restrict__ idx, const float *__restrict__ gout,
                         float *__restrict__ gx) {
  extern __shared__ __align__(16) float acc[];  // [n]
  const size_t cloud = blockIdx.y;
  const int ch = blockIdx.x;
  const int co = MODE ? 2 * c : c;
  const int64_t *ib = idx + cloud * (size_t)n * k;
  const float *g_top = gout + (cloud * (size_t)co + ch) * n * k;
  const float *g_bot = gout + (cloud * (size_t)co + c + ch) * n * k;
  for (int i = threadIdx.x; i < n; i += GG_THREADS) acc[i] = 0.f;
  __syncthreads();
  // coalesced over the (point, neighbour) plane; the own term is reduced over the runs of equal point inside the warp
  // (k consecutive entries belong to one point) before it touches shared memory
  const unsigned int total = (unsigned int)n * (unsigned int)k;
  const unsigned int magic = (unsigned int)((0x100000000ULL + (unsigned int)k - 1) / (unsigned int)k);  // o / k == umulhi(o, magic) for o < 2^32 / k
  const int lane = threadIdx.x & 31;
  for (unsigned int o0 = 0; o0 < total; o0 += GG_THREADS) {
    const unsigned int o = o0 + threadIdx.x;
    const bool ok = o < total;
    int pid = -1;
    float v = 0.f;
    if (ok) {
      const long long j = ib[o];
      const float g = g_top[o];
      atomicAdd(&acc[(int)min(max(j, 0LL), (long long)n - 1)], g);
      if (MODE) {
        pid = k > 1 ? (int)__umulhi(o, magic) : (int)o;  // k == 1: the magic constant 2^32 does not fit 32 bits
        v = g_bot[o] - g;
      }
    }
    if (MODE) {
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float u = __shfl_down_sync(0xffffffffu, v, d);
        const int p = __shfl_down_sync(0xffffffffu, pid, d);
        if (lane + d < 32 && p == pid) v += u;
      }
      const int prev = __shfl_up_sync(0xffffffffu, pid, 1);
      if (ok && (lane == 0 || prev != pid)) atomicAdd(&acc[pid], v);
    }
  }
  __syncthreads();
  float *gr = gx + (cloud * (size_t)c + ch) * n;
  for (int i = threadIdx.x; i < n; i += GG_THREADS) gr[i] = acc[i];
}

// ---- deterministic backward over the target-sorted edge list ------------------------------------------------------
// One CTA per (cloud, channel).  Shared memory: one (point, neighbour) plane of the gradient (n*k floats, bulk copies),
// acc[n], and one partial per chunk of GS_E sorted entries.  The kernel is issue-bound next to the two plane loads, so
// every phase is written for few instructions per edge:
//   1. (mode 1) the bottom plane g[C+c] is loaded; thread = point sums its k slots: own[j] (registers)
//   2. the top plane g[c] replaces it; own[j] -= sum_t top[j][t]
//   3. thread = chunk of GS_E consecutive sorted entries, predicated straight-line code: every run piece is summed in
//      list order; a piece whose run STARTS in the chunk is stored to acc[target] (exactly one writer per target), the
//      piece that continues the previous chunk's run goes to pb[chunk]
//   4. thread = target: acc[j] + the continuation pieces of its run in chunk order + own[j] -> grad_x
constexpr int GS_THREADS = 1024;
constexpr int GS_E = 16;
constexpr int GS_MAXN = 4096;
constexpr int GS_OWN = GS_MAXN / GS_THREADS;  // targets per thread (the 512-thread variant is used up to half GS_MAXN)

__host__ __device__ inline size_t gs_smem_bytes(int n, int k) {
  const size_t total = (size_t)n * k;
  return 16 + sizeof(float) * (total + n + (total + GS_E - 1) / GS_E);
}

__device__ __forceinline__ void gs_load_plane(float *row, const float *src, uint32_t bytes, uint64_t *bar) {
  mbar_expect_tx(bar, bytes);
  for (uint32_t o = 0; o < bytes; o += 32768u)
    bulk_load_1d(reinterpret_cast<unsigned char *>(row) + o, reinterpret_cast<const unsigned char *>(src) + o,
                 min(32768u, bytes - o), bar);
}

template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
graph_gather_grad_sorted_kernel(int c, int n, int k, int estride, const int *__restrict__ off,
                                const unsigned int *__restrict__ rev, const float *__restrict__ gout,
                                float *__restrict__ gx) {
  extern __shared__ __align__(128) unsigned char gs_raw[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(gs_raw);
  float *row = reinterpret_cast<float *>(gs_raw + 16);
  const int total = n * k, nchunk = (total + GS_E - 1) / GS_E;
  float *acc = row + total, *pb = acc + n;
  const size_t cloud = blockIdx.y;
  const int ch = blockIdx.x, co = MODE ? 2 * c : c;
  const float *g_top = gout + (cloud * (size_t)co + ch) * total;
  const float *g_bot = gout + (cloud * (size_t)co + c + ch) * total;
  const int *offb = off + cloud * (size_t)(n + 1);
  const unsigned int *revb = rev + cloud * (size_t)estride;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    gs_load_plane(row, MODE ? g_bot : g_top, (uint32_t)total * 4u, bar);
  }
  // the first chunk's entries do not depend on the plane: fetch them under the plane load
  uint4 nx[GS_E / 4];
  unsigned int nprev = 0xffffffffu;
  auto fetch = [&](int ck) {
    if (ck < nchunk) {
      // interleaved list (edge_pos): the q-th 16 bytes of the warp's 32 chunks are contiguous
      const unsigned int *src = revb + (size_t)(ck >> 5) * 512 + (ck & 31) * 4;
#pragma unroll
      for (int q = 0; q < GS_E / 4; ++q) nx[q] = *reinterpret_cast<const uint4 *>(src + 128 * q);
      nprev = ck ? revb[edge_pos(ck * GS_E - 1)] : 0xffffffffu;  // masked where it is used: no wait on the load here
    }
  };
  fetch(threadIdx.x);
  __syncthreads();
  float own[GS_OWN];
#pragma unroll
  for (int u = 0; u < GS_OWN; ++u) own[u] = 0.f;
  auto plane_row_sum = [&](int j) {
    const float *r = row + j * k;
    float a = 0.f;
    int t = 0;
    for (; t + 5 <= k; t += 5) a = ((((a + r[t]) + r[t + 1]) + r[t + 2]) + r[t + 3]) + r[t + 4];
    for (; t < k; ++t) a += r[t];
    return a;
  };
  if (MODE) {
    mbar_wait(bar, 0);
#pragma unroll
    for (int u = 0; u < GS_OWN; ++u) {
      const int j = threadIdx.x + u * THREADS;
      if (j < n) own[u] = plane_row_sum(j);
    }
    __syncthreads();  // every read of the bottom plane is done: the async proxy may overwrite it
    if (threadIdx.x == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      gs_load_plane(row, g_top, (uint32_t)total * 4u, bar);
    }
    mbar_wait(bar, 1);
#pragma unroll
    for (int u = 0; u < GS_OWN; ++u) {
      const int j = threadIdx.x + u * THREADS;
      if (j < n) own[u] -= plane_row_sum(j);
    }
  } else {
    mbar_wait(bar, 0);
  }
  int ob[GS_OWN], oe[GS_OWN];  // run starts of this thread's targets: in flight during the chunk pass
#pragma unroll
  for (int u = 0; u < GS_OWN; ++u) {
    const int j = threadIdx.x + u * THREADS;
    ob[u] = j < n ? offb[j] : 0;
    oe[u] = j < n ? offb[j + 1] : 0;
  }
  const unsigned char *rowb = reinterpret_cast<const unsigned char *>(row);
  for (int ck = threadIdx.x; ck < nchunk; ck += THREADS) {
    unsigned int ent[GS_E];
#pragma unroll
    for (int q = 0; q < GS_E / 4; ++q)
      ent[4 * q] = nx[q].x, ent[4 * q + 1] = nx[q].y, ent[4 * q + 2] = nx[q].z, ent[4 * q + 3] = nx[q].w;
    const unsigned int prev = nprev;
    fetch(ck + THREADS);  // the next round's entries fly while this chunk is summed
    const int cnt = total - ck * GS_E;
    float *dst = (ck == 0 || ((ent[0] ^ prev) & 0xfffu) != 0) ? acc + (ent[0] & 0xfffu) : pb + ck;
    float sum = 0.f;
    if (cnt >= GS_E) {
#pragma unroll
      for (int q = 0; q < GS_E; ++q) {
        const unsigned int e = ent[q];
        sum += *reinterpret_cast<const float *>(rowb + (e >> 12));
        if (q + 1 == GS_E) {
          *dst = sum;
        } else if (((ent[q + 1] ^ e) & 0xfffu) != 0) {
          *dst = sum;
          sum = 0.f;
          dst = acc + (ent[q + 1] & 0xfffu);
        }
      }
    } else {  // the cloud's last chunk
#pragma unroll
      for (int q = 0; q < GS_E - 1; ++q) {
        if (q < cnt) {
          const unsigned int e = ent[q];
          sum += *reinterpret_cast<const float *>(rowb + (e >> 12));
          if (q + 1 == cnt) {
            *dst = sum;
          } else if (((ent[q + 1] ^ e) & 0xfffu) != 0) {
            *dst = sum;
            sum = 0.f;
            dst = acc + (ent[q + 1] & 0xfffu);
          }
        }
      }
    }
  }
  __syncthreads();
  float *gr = gx + (cloud * (size_t)c + ch) * n;
#pragma unroll
  for (int u = 0; u < GS_OWN; ++u) {
    const int j = threadIdx.x + u * THREADS;
    if (j < n) {
      const int beg = ob[u], end = oe[u];
      float a = 0.f;
      if (end > beg) {
        a = acc[j];
        for (int ck = beg / GS_E + 1; ck <= (end - 1) / GS_E; ++ck) a += pb[ck];
      }
      gr[j] = a + own[u];
    }
  }
}

template <int MODE>
static int launch_gather(int b, int c, int n, int k, const float *x, const int64_t *idx, float *out, cudaStream_t st) {
  const size_t smem = sizeof(float) * GG_CH * n;
  static size_t attr[64];
  if (cudaError_t e = smem_optin(graph_gather_kernel<MODE>, smem, attr); e != cudaSuccess) return (int)e;
  graph_gather_kernel<MODE><<<dim3((n + GG_PT - 1) / GG_PT, b), GG_THREADS, smem, st>>>(c, n, k, x, idx, out);
  return finish_launch(1);
}

// does the deterministic sorted backward cover this shape?
static bool gather_sorted_ok(int b, int n, int k) {
  static const bool force_atomic = getenv("PCC_GATHER_GRAD_ATOMIC") != nullptr;  // measurement switch
  return !force_atomic && k <= 32 && (n * k) % 4 == 0 && gs_smem_bytes(n, k) <= 227 * 1024 && n <= GS_MAXN && edge_sort_ok(b, n, k, 1);
}

template <int MODE>
static int launch_gather_grad_sorted(int b, int c, int n, int k, const int *off, const unsigned int *rev, int estride,
                                     const float *gout, float *gx, cudaStream_t st) {
  const size_t sorted_smem = gs_smem_bytes(n, k);
  // planes of up to ~100 KB: two 512-thread CTAs per SM, one sums while the other's plane loads
  const bool half = sorted_smem <= 110 * 1024 && n <= GS_MAXN / 2;
  static size_t attr[64], attr_h[64];
  if (cudaError_t e = half ? smem_optin(graph_gather_grad_sorted_kernel<MODE, 512>, sorted_smem, attr_h)
                           : smem_optin(graph_gather_grad_sorted_kernel<MODE, 1024>, sorted_smem, attr);
      e != cudaSuccess)
    return (int)e;
  if (half)
    graph_gather_grad_sorted_kernel<MODE, 512><<<dim3(c, b), 512, sorted_smem, st>>>(c, n, k, estride, off, rev, gout, gx);
  else
    graph_gather_grad_sorted_kernel<MODE, 1024><<<dim3(c, b), 1024, sorted_smem, st>>>(c, n, k, estride, off, rev, gout, gx);
  return finish_launch(1);
}

template <int MODE>
static int launch_gather_grad(int b, int c, int n, int k, const int64_t *idx, const float *gout, float *gx,
                              cudaStream_t st) {
  if (gather_sorted_ok(b, n, k) && (reinterpret_cast<uintptr_t>(gout) & 15) == 0) {
    char *ws = nullptr;
    if (cudaError_t e = ws_alloc((void **)&ws, edge_sort_ws_bytes(b, n, k, 1), st); e != cudaSuccess) return (int)e;
    const int *off;
    const unsigned int *rev;
    int estride;
    edge_sort_launch(b, n, k, 1, idx, ws, true, &off, &rev, &estride, st);
    g_launches.fetch_add(4, std::memory_order_relaxed);
    const int rc = launch_gather_grad_sorted<MODE>(b, c, n, k, off, rev, estride, gout, gx, st);
    cudaFreeAsync(ws, st);
    return rc;
  }
  const size_t smem = sizeof(float) * n;
  graph_gather_grad_kernel<MODE><<<dim3(c, b), GG_THREADS, smem, st>>>(c, n, k, idx, gout, gx);
  return finish_launch(1);
}

// ---- decoder-output smoothing (graph_filtering, src/utils/neighbour_ops.py:122-133) ----------------------------------
// out_i = (1 + sum_t w_it) x_i - sum_t w_it x_{j_it},  w_it = exp(-d_it / sigma),  d_it = |x_i - x_{j_it}| over the k-1
// nearest neighbours (column 0 of the kNN list is the point itself),  sigma = max(mean_i d_i1, 0.005) per cloud.
// The reference runs ~15 elementwise / gather / reduce kernels forward and twice that backward on (B,3,N,3) tensors;
// here one CTA per cloud keeps the cloud in shared memory and does each direction in one launch.
constexpr int GF_THREADS = 1024;
constexpr int GF_MAXK = 8;

__device__ __forceinline__ float gf_block_sum(float v, float *red) {  // fixed order: deterministic
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < GF_THREADS / 32; ++w) s += red[w];
  __syncthreads();
  return s;
}

__global__ void __launch_bounds__(GF_THREADS)
graph_filter_kernel(int n, int k, const float *__restrict__ x, const int64_t *__restrict__ idx, float *__restrict__ out,
                    float *__restrict__ mean_out) {
  extern __shared__ __align__(16) float gf_sm[];  // xs[3][n]
  __shared__ float red[GF_THREADS / 32];
  float *xs = gf_sm, *ys = xs + n, *zs = ys + n;
  const size_t cloud = blockIdx.x;
  const float *xb = x + cloud * (size_t)3 * n;
  const int64_t *ib = idx + cloud * (size_t)n * k;
  for (int i = threadIdx.x; i < 3 * n; i += GF_THREADS) gf_sm[i] = xb[i];
  __syncthreads();
  float s0 = 0.f;
  for (int i = threadIdx.x; i < n; i += GF_THREADS) {
    const int j = (int)min(max(ib[(size_t)i * k + 1], (int64_t)0), (int64_t)n - 1);
    const float dx = xs[i] - xs[j], dy = ys[i] - ys[j], dz = zs[i] - zs[j];
    s0 += sqrtf(fabsf(dx * dx + dy * dy + dz * dz));
  }
  const float mean = gf_block_sum(s0, red) / (float)n;
  const float sigma = fmaxf(mean, 0.005f);
  if (threadIdx.x == 0) mean_out[cloud] = mean;
  float *ob = out + cloud * (size_t)3 * n;
  for (int i = threadIdx.x; i < n; i += GF_THREADS) {
    const float xi = xs[i], yi = ys[i], zi = zs[i];
    float wsum = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
    for (int t = 1; t < k; ++t) {
      const int j = (int)min(max(ib[(size_t)i * k + t], (int64_t)0), (int64_t)n - 1);
      const float xj = xs[j], yj = ys[j], zj = zs[j];
      const float dx = xi - xj, dy = yi - yj, dz = zi - zj;
      const float w = expf(-(sqrtf(fabsf(dx * dx + dy * dy + dz * dz)) / sigma));
      wsum += w;
      ax += w * xj;
      ay += w * yj;
      az += w * zj;
    }
    ob[i] = (1.f + wsum) * xi - ax;
    ob[(size_t)n + i] = (1.f + wsum) * yi - ay;
    ob[(size_t)2 * n + i] = (1.f + wsum) * zi - az;
  }
}

// grad_x of the above.  Duplicate points (d = 0) contribute no distance gradient (torch yields NaN there).
__global__ void __launch_bounds__(GF_THREADS)
graph_filter_grad_kernel(int n, int k, const float *__restrict__ x, const int64_t *__restrict__ idx,
                         const float *__restrict__ mean_in, const float *__restrict__ gout, float *__restrict__ gx) {
  extern __shared__ __align__(16) float gf_sm[];  // xs[3][n], gs[3][n], acc[3][n]
  __shared__ float red[GF_THREADS / 32];
  float *xs = gf_sm, *ys = xs + n, *zs = ys + n;
  float *gxs = zs + n, *gys = gxs + n, *gzs = gys + n;
  float *acc = gzs + n;
  const size_t cloud = blockIdx.x;
  const float *xb = x + cloud * (size_t)3 * n, *gb = gout + cloud * (size_t)3 * n;
  const int64_t *ib = idx + cloud * (size_t)n * k;
  for (int i = threadIdx.x; i < 3 * n; i += GF_THREADS) {
    gf_sm[i] = xb[i];
    gxs[i] = gb[i];
    acc[i] = 0.f;
  }
  __syncthreads();
  const float mean = mean_in[cloud];
  const float sigma = fmaxf(mean, 0.005f);
  // dL/dsigma = sum_{i,t} (g_i . diff_it) w_it d_it / sigma^2   (only while the clamp is inactive)
  float ds = 0.f;
  for (int i = threadIdx.x; i < n; i += GF_THREADS) {
    for (int t = 1; t < k; ++t) {
      const int j = (int)min(max(ib[(size_t)i * k + t], (int64_t)0), (int64_t)n - 1);
      const float dx = xs[i] - xs[j], dy = ys[i] - ys[j], dz = zs[i] - zs[j];
      const float d = sqrtf(fabsf(dx * dx + dy * dy + dz * dz));
      const float w = expf(-(d / sigma));
      ds += (gxs[i] * dx + gys[i] * dy + gzs[i] * dz) * w * d;
    }
  }
  const float dLds = (mean >= 0.005f) ? gf_block_sum(ds, red) / (sigma * sigma) : (gf_block_sum(ds, red), 0.f);
  const float per_point = dLds / (float)n;  // d sigma / d d_i1
  for (int i = threadIdx.x; i < n; i += GF_THREADS) {
    const float xi = xs[i], yi = ys[i], zi = zs[i];
    const float gi0 = gxs[i], gi1 = gys[i], gi2 = gzs[i];
    float wsum = 0.f, ox = 0.f, oy = 0.f, oz = 0.f;
    for (int t = 1; t < k; ++t) {
      const int j = (int)min(max(ib[(size_t)i * k + t], (int64_t)0), (int64_t)n - 1);
      const float dx = xi - xs[j], dy = yi - ys[j], dz = zi - zs[j];
      const float d = sqrtf(fabsf(dx * dx + dy * dy + dz * dz));
      const float w = expf(-(d / sigma));
      wsum += w;
      // through the distance: dL/dd = (g . diff) * (-w / sigma) [+ the sigma path for the nearest neighbour]
      float dLdd = (gi0 * dx + gi1 * dy + gi2 * dz) * (-w / sigma);
      if (t == 1) dLdd += per_point;
      const float coef = d > 0.f ? dLdd / d : 0.f;
      ox += coef * dx;
      oy += coef * dy;
      oz += coef * dz;
      atomicAdd(&acc[j], -(gi0 * w) - coef * dx);
      atomicAdd(&acc[n + j], -(gi1 * w) - coef * dy);
      atomicAdd(&acc[2 * n + j], -(gi2 * w) - coef * dz);
    }
    atomicAdd(&acc[i], gi0 * (1.f + wsum) + ox);
    atomicAdd(&acc[n + i], gi1 * (1.f + wsum) + oy);
    atomicAdd(&acc[2 * n + i], gi2 * (1.f + wsum) + oz);
  }
  __syncthreads();
  float *gr = gx + cloud * (size_t)3 * n;
  for (int i = threadIdx.x; i < 3 * n; i += GF_THREADS) gr[i] = acc[i];
}

}  // namespace pcc

using namespace pcc;

extern "C" __attribute__((visibility("default"))) int pcc_graph_gather(int b, int c, int n, int k, const float *x, const int64_t *idx, int mode,
                                                                        float *out, pcc_stream_t stream) {
  if (b < 0 || c <= 0 || n <= 0 || k <= 0 || (mode != 0 && mode != 1)) return PCC_EINVAL;
  if (b == 0) return PCC_OK;
  if (k > 32 || n > GG_MAXN || b > 65535) return PCC_ENOTSUP;
  cudaStream_t st = (cudaStream_t)stream;
  return mode ? launch_gather<1>(b, c, n, k, x, idx, out, st) : launch_gather<0>(b, c, n, k, x, idx, out, st);
}

extern "C" __attribute__((visibility("default"))) int pcc_graph_gather_grad(int b, int c, int n, int k, const int64_t *idx, int mode,
                                                                             const float *grad_out, float *grad_x,
                                                                             pcc_stream_t stream) {
  if (b < 0 || c <= 0 || n <= 0 || k <= 0 || (mode != 0 && mode != 1)) return PCC_EINVAL;
  if (b == 0) return PCC_OK;
  if (n > GG_MAXN || b > 65535 || c > 65535) return PCC_ENOTSUP;
  cudaStream_t st = (cudaStream_t)stream;
  return mode ? launch_gather_grad<1>(b, c, n, k, idx, grad_out, grad_x, st)
              : launch_gather_grad<0>(b, c, n, k, idx, grad_out, grad_x, st);
}

// The edge sort of the backward depends on idx alone: a caller that knows a backward will follow can run it early (e.g. on a
// side stream under the forward gather) into its own buffer and hand it to pcc_graph_gather_grad_presorted.
extern "C" __attribute__((visibility("default"))) long long pcc_graph_edge_sort_bytes(int b, int n, int k) {
  if (b <= 0 || n <= 0 || k <= 0 || n > GG_MAXN || b > 65535 || !gather_sorted_ok(b, n, k)) return 0;
  return (long long)edge_sort_ws_bytes(b, n, k, 1);
}

extern "C" __attribute__((visibility("default"))) int pcc_graph_edge_sort(int b, int n, int k, const int64_t *idx, void *ws, pcc_stream_t stream) {
  if (pcc_graph_edge_sort_bytes(b, n, k) == 0) return PCC_ENOTSUP;
  const int *off;
  const unsigned int *rev;
  int estride;
  edge_sort_launch(b, n, k, 1, idx, reinterpret_cast<char *>(ws), true, &off, &rev, &estride, (cudaStream_t)stream);
  return finish_launch(4);
}

extern "C" __attribute__((visibility("default"))) int pcc_graph_gather_grad_presorted(int b, int c, int n, int k, int mode, const void *ws,
                                                                                       const float *grad_out, float *grad_x,
                                                                                       pcc_stream_t stream) {
  if (b <= 0 || c <= 0 || n <= 0 || k <= 0 || (mode != 0 && mode != 1)) return PCC_EINVAL;
  if (c > 65535 || pcc_graph_edge_sort_bytes(b, n, k) == 0 || (reinterpret_cast<uintptr_t>(grad_out) & 15) != 0) return PCC_ENOTSUP;
  const int *off;
  const unsigned int *rev;
  int estride;
  edge_sort_views(b, n, k, 1, reinterpret_cast<const char *>(ws), &off, &rev, &estride);
  cudaStream_t st = (cudaStream_t)stream;
  return mode ? launch_gather_grad_sorted<1>(b, c, n, k, off, rev, estride, grad_out, grad_x, st)
              : launch_gather_grad_sorted<0>(b, c, n, k, off, rev, estride, grad_out, grad_x, st);
}

extern "C" __attribute__((visibility("default"))) int pcc_graph_filtering(int b, int n, int k, const float *x, const int64_t *idx, float *out,
                                                                           float *mean_dist, pcc_stream_t stream) {
  if (b < 0 || n <= 0 || k < 2) return PCC_EINVAL;
  if (b == 0) return PCC_OK;
  if (k > GF_MAXK || n > 6144) return PCC_ENOTSUP;
  const size_t smem = sizeof(float) * 3 * n;
  static size_t attr[64];
  if (cudaError_t e = smem_optin(graph_filter_kernel, smem, attr); e != cudaSuccess) return (int)e;
  graph_filter_kernel<<<b, GF_THREADS, smem, (cudaStream_t)stream>>>(n, k, x, idx, out, mean_dist);
  return finish_launch(1);
}

extern "C" __attribute__((visibility("default"))) int pcc_graph_filtering_grad(int b, int n, int k, const float *x, const int64_t *idx,
                                                                                const float *mean_dist, const float *grad_out,
                                                                                float *grad_x, pcc_stream_t stream) {
  if (b < 0 || n <= 0 || k < 2) return PCC_EINVAL;
  if (b == 0) return PCC_OK;
  if (k > GF_MAXK || n > 6144) return PCC_ENOTSUP;
  const size_t smem = sizeof(float) * 9 * n;
  static size_t attr[64];
  if (cudaError_t e = smem_optin(graph_filter_grad_kernel, smem, attr); e != cudaSuccess) return (int)e;
  graph_filter_grad_kernel<<<b, GF_THREADS, smem, (cudaStream_t)stream>>>(n, k, x, idx, mean_dist, grad_out, grad_x);
  return finish_launch(1);
}
