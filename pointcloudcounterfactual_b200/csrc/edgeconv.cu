// Fused EdgeConv layer (full row): graph features -> 1x1 conv -> BatchNorm2d -> activation -> max over k,
// without the (B,2C,N,k) edge tensor and without the (B,Cout,N,k) convolution output.
//
// Replaces the sequence  get_graph_features (src/utils/neighbour_ops.py:113-119)  ->  EdgeConvLayer.forward
// (src/module/layers.py:159-203: Conv2d 1x1 without bias, BatchNorm2d, activation)  ->  x.max(dim=3)
// (src/module/encoders.py:49-54, classifier.py:55-60).  The convolution is linear, so for edge (i -> j)
//     y[o][i][j] = W[o] . [x_j - x_i ; x_i] = u[o][j] + v[o][i],   u = W1 x,  v = (W2 - W1) x,  W = [W1 | W2]
// and the host computes [u | v] for all points with ONE plain library GEMM (B,N,C) x (C,2Cout) -- 1/k of the
// reference's convolution work.  Everything after that is this file:
//   ec_reduce_kernel    one pass over the edges: per (point, channel) sum_e y, the extremum of y over the k neighbours
//                       (max where gamma >= 0, min where gamma < 0: BatchNorm is a per-channel affine map whose slope
//                       has the sign of gamma and the supported activations are non-decreasing, so
//                       max_k act(bn(y)) = act(bn(ext_k y))) and its slot; per-CTA partial sums of y and y^2 in fp64
//   ec_stats_kernel     batch statistics over all B*N*k edges (what BatchNorm2d sees), running-stat update, the
//                       per-channel affine map
//   ec_finalize_kernel  out (B,Cout,N) = act(scale * ext + shift), transposed through shared memory
// Backward (all edges get a gradient through the batch statistics, not only the arg-max edge):
//     dL/dy_e = a [e == e*] dz - c1 - c2 (y_e - mean),  a = gamma invstd, c1 = a dbeta / E, c2 = a invstd dgamma / E
//   ec_bwd_point_kernel dz = g act'(z) per (point, channel) + fp64 partials of dbeta, dgamma
//   ec_bwd_stats_kernel dbeta, dgamma, (a, c1, c2)
//   ec_hist/scan/fill/sort_kernel  edges sorted by TARGET (counting sort per cloud, each run ordered by edge id), so that
//   ec_bwd_chunk/finish_kernel  grad u_j = sum over the edges that point at j is a segmented sum over that list in a
//                       fixed order: deterministic, no float atomics (torch's backward of gather is an atomic
//                       scatter-add), and balanced whatever the in-degrees are (hubs of several hundred edges in
//                       feature space)
// All per-point tensors are point-major (B,N,Cout): one neighbour = one contiguous row, L2-resident (0.5-2 MB per cloud).
#include "common.cuh"

namespace pcc {

constexpr int EC_THREADS = 256;
constexpr int EC_PTS = 64;       // points per CTA in the edge kernels
constexpr int EC_MAX_K = 64;
constexpr int EC_MAX_N = 8192;   // (source << 8 | slot) must fit an int; the gather kernel shares the limit

enum { EC_BN_EVAL = 0, EC_BN_TRAIN = 1, EC_AFFINE = 2 };
enum { EC_ACT_NONE = 0, EC_ACT_LEAKY = 1 };

__device__ __forceinline__ int ec_clamp(long long j, int n) { return (int)min(max(j, 0LL), (long long)n - 1); }

// slot (uint8) and dz (fp32) are SLICE-major: [cloud][slice of 8 channels][point][8], so that a CTA that owns one slice
// of one cloud (the shared-memory staged kernels below) reads them as one contiguous block.  Offset of channel ch:
__host__ __device__ __forceinline__ int ec_nslice(int cout) { return (cout + 7) >> 3; }
__device__ __forceinline__ size_t ec_sl(int cloud, int n, int cout, int i, int ch) {
  return (((size_t)cloud * ec_nslice(cout) + (ch >> 3)) * n + i) * 8 + (ch & 7);
}

// ---- forward: one pass over the edges -------------------------------------------------------------------------
// thread = (point slot, quad of 4 channels); a CTA walks EC_PTS points of one cloud.
template <bool STATS>
__global__ void __launch_bounds__(EC_THREADS)
ec_reduce_kernel(int n, int k, int cout, const float *__restrict__ uv, const int64_t *__restrict__ idx,
                 const float *__restrict__ gamma, float *__restrict__ exty, float *__restrict__ sy_out,
                 unsigned char *__restrict__ slot_out, double *__restrict__ partials, int nparts) {
  __shared__ int sidx[EC_PTS * EC_MAX_K];
  __shared__ double red[2][EC_THREADS * 4];
  const int cloud = blockIdx.y;
  const int i0 = blockIdx.x * EC_PTS;
  const int npts = min(EC_PTS, n - i0);
  const int tpp = cout >> 2;               // threads per point
  const int groups = EC_THREADS / tpp;     // point slots per iteration
  const int grp = threadIdx.x / tpp, quad = threadIdx.x - grp * tpp;
  const bool active = grp < groups;
  const int64_t *ib = idx + ((size_t)cloud * n + i0) * k;
  for (int e = threadIdx.x; e < npts * k; e += EC_THREADS) sidx[e] = ec_clamp(ib[e], n);
  __syncthreads();
  const float *uvb = uv + (size_t)cloud * n * 2 * cout;
  // channels with a negative gamma need the MINIMUM over the neighbours: track max(sg * y) with sg = -1 there.  sg * y
  // = fma(u, sg, sg * v) is exact, so the result equals the direct min / max bit for bit.
  float sg[4] = {1.f, 1.f, 1.f, 1.f};
  if (active && gamma) {
    const float4 g4 = reinterpret_cast<const float4 *>(gamma)[quad];
    sg[0] = g4.x < 0.f ? -1.f : 1.f;
    sg[1] = g4.y < 0.f ? -1.f : 1.f;
    sg[2] = g4.z < 0.f ? -1.f : 1.f;
    sg[3] = g4.w < 0.f ? -1.f : 1.f;
  }
  double a1[4] = {0., 0., 0., 0.}, a2[4] = {0., 0., 0., 0.};
  if (active) {
    for (int p = grp; p < npts; p += groups) {
      const int i = i0 + p;
      const float4 v4 = *reinterpret_cast<const float4 *>(uvb + (size_t)i * 2 * cout + cout + 4 * quad);
      const float v[4] = {v4.x * sg[0], v4.y * sg[1], v4.z * sg[2], v4.w * sg[3]};
      const float NINF = __int_as_float(0xff800000);
      float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f}, ext[4] = {NINF, NINF, NINF, NINF};
      int slot[4] = {0, 0, 0, 0};
      const int *nb = sidx + p * k;
#pragma unroll 5
      for (int t = 0; t < k; ++t) {
        const float4 u4 = *reinterpret_cast<const float4 *>(uvb + (size_t)nb[t] * 2 * cout + 4 * quad);
        const float y[4] = {fmaf(u4.x, sg[0], v[0]), fmaf(u4.y, sg[1], v[1]), fmaf(u4.z, sg[2], v[2]),
                            fmaf(u4.w, sg[3], v[3])};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (STATS) {
            s1[c] += y[c];
            s2[c] = fmaf(y[c], y[c], s2[c]);
          }
          slot[c] = y[c] > ext[c] ? t : slot[c];  // strict: the first slot wins a tie
          ext[c] = fmaxf(ext[c], y[c]);
        }
      }
      const size_t o = ((size_t)cloud * n + i) * cout + 4 * quad;
      *reinterpret_cast<float4 *>(exty + o) = make_float4(ext[0] * sg[0], ext[1] * sg[1], ext[2] * sg[2], ext[3] * sg[3]);
      *reinterpret_cast<uchar4 *>(slot_out + ec_sl(cloud, n, cout, i, 4 * quad)) =
          make_uchar4((unsigned char)slot[0], (unsigned char)slot[1], (unsigned char)slot[2], (unsigned char)slot[3]);
      if (STATS) {
#pragma unroll
        for (int c = 0; c < 4; ++c) s1[c] *= sg[c];
        *reinterpret_cast<float4 *>(sy_out + o) = make_float4(s1[0], s1[1], s1[2], s1[3]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          a1[c] += (double)s1[c];
          a2[c] += (double)s2[c];
        }
      }
    }
  }
  if (STATS) {
    // per-CTA partial sums: fixed order over the point slots -> deterministic
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      red[0][threadIdx.x * 4 + c] = a1[c];
      red[1][threadIdx.x * 4 + c] = a2[c];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < 2 * cout; o += EC_THREADS) {
      const int s = o / cout, ch = o - s * cout;
      double acc = 0.;
      for (int g = 0; g < groups; ++g) acc += red[s][(g * tpp + (ch >> 2)) * 4 + (ch & 3)];
      const int part = blockIdx.y * gridDim.x + blockIdx.x;
      partials[((size_t)s * cout + ch) * nparts + part] = acc;  // [2][cout][nparts]
    }
  }
}

// Same pass with the gather served from SHARED memory: a CTA owns one slice of 8 channels of one cloud, stages that
// slice of u for all n points (32 B per point) and walks all points.  The gathered bytes (k rows per point) then come
// from shared memory instead of L2; L2 only delivers the slice once, the indices and the own v.  thread = (point slot,
// quad of the slice).  Used while the slice fits (n <= EC_STAGE_MAX_N).
constexpr int EC_STAGE_MAX_N = 2560;
constexpr int EC_STHREADS = 512;            // threads of the staged kernels
constexpr int EC_SPTS = EC_STHREADS / 2;    // points (chunks) per iteration

template <bool STATS>
__global__ void __launch_bounds__(EC_STHREADS, 2)
ec_reduce_staged_kernel(int n, int k, int cout, const float *__restrict__ uv, const int64_t *__restrict__ idx,
                        const float *__restrict__ gamma, float *__restrict__ exty, float *__restrict__ sy_out,
                        unsigned char *__restrict__ slot_out, double *__restrict__ partials, int nparts) {
  extern __shared__ __align__(16) float4 rows[];  // [n][2]; reused for the statistics reduction at the end
  const int slice = blockIdx.x, cloud = blockIdx.y;
  const int nquad = cout >> 2;
  const float *uvb = uv + (size_t)cloud * n * 2 * cout;
  const int ps = threadIdx.x >> 1, qq = threadIdx.x & 1;
  const int quad = 2 * slice + qq;
  const bool active = quad < nquad;
  // channels with a negative gamma need the MINIMUM over the neighbours: the slice is staged as sg * u, sg = -1 there, and
  // the maximum of sg * y = sg * u + sg * v is tracked (sign flips are exact: same bits as the direct min / max)
  float sg[4] = {1.f, 1.f, 1.f, 1.f};
  if (active && gamma) {
    const float4 g4 = reinterpret_cast<const float4 *>(gamma)[quad];
    sg[0] = g4.x < 0.f ? -1.f : 1.f;
    sg[1] = g4.y < 0.f ? -1.f : 1.f;
    sg[2] = g4.z < 0.f ? -1.f : 1.f;
    sg[3] = g4.w < 0.f ? -1.f : 1.f;
  }
  for (int e = threadIdx.x; e < 2 * n; e += EC_STHREADS) {  // e & 1 == qq: this thread stages its own quad
    float4 u4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) u4 = *reinterpret_cast<const float4 *>(uvb + (size_t)(e >> 1) * 2 * cout + 4 * quad);
    rows[e] = make_float4(u4.x * sg[0], u4.y * sg[1], u4.z * sg[2], u4.w * sg[3]);
  }
  double a1[4] = {0., 0., 0., 0.}, a2[4] = {0., 0., 0., 0.};
  const int64_t *ib = idx + (size_t)cloud * n * k;
  // neighbour lists of the points of one iteration: read coalesced, kept as 16-bit indices (n <= EC_STAGE_MAX_N)
  unsigned short *sidx = reinterpret_cast<unsigned short *>(rows + max(2 * n, 2 * EC_STHREADS * 4 * 8 / 16));
  const float NINF = __int_as_float(0xff800000);
  for (int it0 = 0; it0 < n; it0 += EC_SPTS) {
    __syncthreads();
    const int cnt = min(EC_SPTS, n - it0) * k;
    for (int e = threadIdx.x; e < cnt; e += EC_STHREADS)
      sidx[e] = (unsigned short)ec_clamp(ib[(size_t)it0 * k + e], n);
    __syncthreads();
    const int i = it0 + ps;
    if (i >= n || !active) continue;
    const float4 v4 = *reinterpret_cast<const float4 *>(uvb + (size_t)i * 2 * cout + cout + 4 * quad);
    const float v[4] = {v4.x * sg[0], v4.y * sg[1], v4.z * sg[2], v4.w * sg[3]};
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f}, ext[4] = {NINF, NINF, NINF, NINF};
    int slot[4] = {0, 0, 0, 0};
    const unsigned short *nb = sidx + ps * k;
    auto edge = [&](const float4 u4, int t) {
      const float y[4] = {u4.x + v[0], u4.y + v[1], u4.z + v[2], u4.w + v[3]};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (STATS) {
          s1[c] += y[c];
          s2[c] = fmaf(y[c], y[c], s2[c]);
        }
        slot[c] = y[c] > ext[c] ? t : slot[c];  // strict: the first slot wins a tie
        ext[c] = fmaxf(ext[c], y[c]);
      }
    };
    int t0 = 0;
    for (; t0 + 4 <= k; t0 += 4) {  // four rows in flight, no bounds checks inside
      float4 u8[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) u8[u] = rows[2 * (int)nb[t0 + u] + qq];
#pragma unroll
      for (int u = 0; u < 4; ++u) edge(u8[u], t0 + u);
    }
    for (; t0 < k; ++t0) edge(rows[2 * (int)nb[t0] + qq], t0);
    const size_t o = ((size_t)cloud * n + i) * cout + 4 * quad;
    *reinterpret_cast<float4 *>(exty + o) = make_float4(ext[0] * sg[0], ext[1] * sg[1], ext[2] * sg[2], ext[3] * sg[3]);
    *reinterpret_cast<uchar4 *>(slot_out + ec_sl(cloud, n, cout, i, 4 * quad)) =
        make_uchar4((unsigned char)slot[0], (unsigned char)slot[1], (unsigned char)slot[2], (unsigned char)slot[3]);
    if (STATS) {
#pragma unroll
      for (int c = 0; c < 4; ++c) s1[c] *= sg[c];
      *reinterpret_cast<float4 *>(sy_out + o) = make_float4(s1[0], s1[1], s1[2], s1[3]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        a1[c] += (double)s1[c];
        a2[c] += (double)s2[c];
      }
    }
  }
  if (STATS) {
    __syncthreads();  // every gather is done: the slice buffer becomes the reduction scratch (32 KiB are always allocated)
    double *red = reinterpret_cast<double *>(rows);  // [2][EC_STHREADS][4]
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      red[threadIdx.x * 4 + c] = a1[c];
      red[EC_STHREADS * 4 + threadIdx.x * 4 + c] = a2[c];
    }
    __syncthreads();
    if (threadIdx.x < 16) {  // (statistic, channel of the slice): fixed order over the point slots -> deterministic
      const int st = threadIdx.x >> 3, ch8 = threadIdx.x & 7;
      const int ch = 8 * slice + ch8;
      if (ch < cout) {
        double acc = 0.;
        for (int g = 0; g < EC_SPTS; ++g) acc += red[st * EC_STHREADS * 4 + (2 * g + (ch8 >> 2)) * 4 + (ch8 & 3)];
        partials[((size_t)st * cout + ch) * nparts + cloud] = acc;  // [2][cout][nparts = b]
      }
    }
  }
}

// deterministic block sum of doubles (fixed tree)
template <int T>
__device__ __forceinline__ double ec_block_sum(double v, double *sm) {
  sm[threadIdx.x] = v;
  __syncthreads();
#pragma unroll
  for (int s = T / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  const double r = sm[0];
  __syncthreads();
  return r;
}

// one CTA per channel: statistics over all E = B*N*k edges, running-stat update, per-channel affine map
__global__ void __launch_bounds__(128)
ec_stats_kernel(int cout, int nparts, double edges, int bn_mode, const double *__restrict__ partials,
                const float *__restrict__ gamma, const float *__restrict__ beta, float *__restrict__ running_mean,
                float *__restrict__ running_var, float momentum, float eps, float *__restrict__ mean_out,
                float *__restrict__ invstd_out, float *__restrict__ scale, float *__restrict__ shift) {
  __shared__ double sm[128];
  const int o = blockIdx.x;
  float mean = 0.f, invstd = 1.f;
  if (bn_mode == EC_BN_TRAIN) {
    double s1 = 0., s2 = 0.;
    for (int p = threadIdx.x; p < nparts; p += 128) {
      s1 += partials[(size_t)o * nparts + p];
      s2 += partials[((size_t)cout + o) * nparts + p];
    }
    s1 = ec_block_sum<128>(s1, sm);
    s2 = ec_block_sum<128>(s2, sm);
    const double m = s1 / edges;
    const double var = fmax(s2 / edges - m * m, 0.);  // biased, what BatchNorm normalises with
    mean = (float)m;
    invstd = (float)(1. / sqrt(var + (double)eps));
    if (threadIdx.x == 0 && running_mean && running_var) {
      const double unbiased = edges > 1. ? var * edges / (edges - 1.) : var;
      running_mean[o] = (1.f - momentum) * running_mean[o] + momentum * mean;
      running_var[o] = (1.f - momentum) * running_var[o] + momentum * (float)unbiased;
    }
  } else if (bn_mode == EC_BN_EVAL) {
    mean = running_mean[o];
    invstd = 1.f / sqrtf(running_var[o] + eps);
  }
  if (threadIdx.x == 0) {
    const float g = gamma ? gamma[o] : 1.f, b = beta ? beta[o] : 0.f;
    const float a = g * invstd;
    mean_out[o] = mean;
    invstd_out[o] = invstd;
    scale[o] = a;
    shift[o] = fmaf(-mean, a, b);
  }
}

__device__ __forceinline__ float ec_act(float z, int act, float slope) {
  return (act == EC_ACT_LEAKY && !(z > 0.f)) ? z * slope : z;
}

// out (B,Cout,N) = act(scale * exty + shift); tile of 32 points x 32 channels transposed through shared memory
__global__ void __launch_bounds__(256)
ec_finalize_kernel(int n, int cout, const float *__restrict__ exty, const float *__restrict__ scale,
                   const float *__restrict__ shift, int act, float slope, float *__restrict__ out) {
  __shared__ float tile[32][33];
  const int cloud = blockIdx.z, i0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int o = o0 + tx;
  const float a = o < cout ? scale[o] : 0.f, sh = o < cout ? shift[o] : 0.f;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r;
    if (i < n && o < cout) tile[r][tx] = ec_act(fmaf(a, exty[((size_t)cloud * n + i) * cout + o], sh), act, slope);
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int oo = o0 + r, i = i0 + tx;
    if (oo < cout && i < n) out[((size_t)cloud * cout + oo) * n + i] = tile[tx][r];
  }
}

// ---- backward ---------------------------------------------------------------------------------------------------
// dz (B,N,Cout) = grad_out (B,Cout,N)^T * act'(z), + fp64 partials of sum dz and sum dz * yhat per channel
__global__ void __launch_bounds__(256)
ec_bwd_point_kernel(int n, int cout, const float *__restrict__ gout, const float *__restrict__ exty,
                    const float *__restrict__ mean, const float *__restrict__ invstd, const float *__restrict__ gamma,
                    const float *__restrict__ beta, int act, float slope, float *__restrict__ dz,
                    double *__restrict__ partials, int nparts) {
  __shared__ float tile[32][33];
  __shared__ double red[2][8][32];
  const int cloud = blockIdx.z, i0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int oo = o0 + r, i = i0 + tx;
    tile[r][tx] = (oo < cout && i < n) ? gout[((size_t)cloud * cout + oo) * n + i] : 0.f;
  }
  __syncthreads();
  const int o = o0 + tx;
  double sb = 0., sg = 0.;
  if (o < cout) {
    const float mu = mean[o], is = invstd[o];
    const float a = (gamma ? gamma[o] : 1.f) * is, sh = fmaf(-mu, a, beta ? beta[o] : 0.f);
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const int i = i0 + r;
      if (i < n) {
        const size_t at = ((size_t)cloud * n + i) * cout + o;
        const float e = exty[at];
        const float z = fmaf(a, e, sh);
        const float d = tile[tx][r] * ((act == EC_ACT_LEAKY && !(z > 0.f)) ? slope : 1.f);
        dz[ec_sl(cloud, n, cout, i, o)] = d;
        sb += (double)d;
        sg += (double)(d * ((e - mu) * is));
      }
    }
  }
  red[0][ty][tx] = sb;
  red[1][ty][tx] = sg;
  __syncthreads();
  if (ty < 2 && o < cout) {
    double acc = 0.;
#pragma unroll
    for (int r = 0; r < 8; ++r) acc += red[ty][r][tx];
    const int part = blockIdx.z * gridDim.x + blockIdx.x;
    partials[((size_t)ty * cout + o) * nparts + part] = acc;
  }
}

// dbeta, dgamma and the three per-channel coefficients of dL/dy_e
__global__ void __launch_bounds__(128)
ec_bwd_stats_kernel(int cout, int nparts, double edges, int bn_mode, const double *__restrict__ partials,
                    const float *__restrict__ gamma, const float *__restrict__ invstd, float *__restrict__ dgamma,
                    float *__restrict__ dbeta, float *__restrict__ coef) {
  __shared__ double sm[128];
  const int o = blockIdx.x;
  double sb = 0., sg = 0.;
  for (int p = threadIdx.x; p < nparts; p += 128) {
    sb += partials[(size_t)o * nparts + p];
    sg += partials[((size_t)cout + o) * nparts + p];
  }
  sb = ec_block_sum<128>(sb, sm);
  sg = ec_block_sum<128>(sg, sm);
  if (threadIdx.x == 0) {
    if (dbeta) dbeta[o] = (float)sb;
    if (dgamma) dgamma[o] = (float)sg;
    const double is = (double)invstd[o];
    const double a = (double)(gamma ? gamma[o] : 1.f) * is;
    coef[o] = (float)a;
    coef[cout + o] = bn_mode == EC_BN_TRAIN ? (float)(a * sb / edges) : 0.f;
    coef[2 * cout + o] = bn_mode == EC_BN_TRAIN ? (float)(a * is * sg / edges) : 0.f;
  }
}

// Edges sorted by TARGET, per cloud: off (n+1) run starts, rev (n*k, per-cloud stride padded to whole chunks) packed
// (target << 19 | source << 6 | slot) -- 13 + 13 + 6 bits, n <= 8192, k <= 64 -- ascending inside each run.
constexpr int EC_PARTS = 8;  // the edges of a cloud are histogrammed / filled in EC_PARTS contiguous ranges, one CTA each

// hist2[cloud][part][j] = edges of the part's range that point at j: shared-memory atomics only
__global__ void __launch_bounds__(512)
ec_hist_kernel(int n, int total, const int64_t *__restrict__ idx, int *__restrict__ hist2) {
  extern __shared__ int sc[];  // [n]
  const int cloud = blockIdx.y, part = blockIdx.x;
  const int per = (total + EC_PARTS - 1) / EC_PARTS;
  const int e0 = min(part * per, total), e1 = min(e0 + per, total);
  for (int j = threadIdx.x; j < n; j += 512) sc[j] = 0;
  __syncthreads();
  for (int e = e0 + threadIdx.x; e < e1; e += 512) atomicAdd(&sc[ec_clamp(idx[(size_t)cloud * total + e], n)], 1);
  __syncthreads();
  int *h = hist2 + ((size_t)cloud * EC_PARTS + part) * n;
  for (int j = threadIdx.x; j < n; j += 512) h[j] = sc[j];
}

// off[j] = start of target j's run; base2[cloud][part][j] = where the part's entries for j begin inside that run
__global__ void __launch_bounds__(1024)
ec_scan_kernel(int n, int total, const int *__restrict__ hist2, int *__restrict__ off, int *__restrict__ base2) {
  __shared__ int wsum[32];
  const int cloud = blockIdx.x;
  const int *h = hist2 + (size_t)cloud * EC_PARTS * n;
  int *offb = off + (size_t)cloud * (n + 1), *bs = base2 + (size_t)cloud * EC_PARTS * n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunk = (n + 1023) / 1024;  // each thread owns a contiguous chunk
  const int j0 = threadIdx.x * chunk, j1 = min(n, j0 + chunk);
  int local = 0;
  for (int j = j0; j < j1; ++j)
#pragma unroll
    for (int p = 0; p < EC_PARTS; ++p) local += h[(size_t)p * n + j];
  int incl = local;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = wsum[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += t;
    }
    wsum[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  int run = incl - local + (warp ? wsum[warp - 1] : 0);
  for (int j = j0; j < j1; ++j) {
    offb[j] = run;
#pragma unroll
    for (int p = 0; p < EC_PARTS; ++p) {
      bs[(size_t)p * n + j] = run;
      run += h[(size_t)p * n + j];
    }
  }
  if (threadIdx.x == 0) offb[n] = total;
}

// every part places its edges with shared-memory cursors that start at base2: no global atomics
__global__ void __launch_bounds__(512)
ec_fill_kernel(int n, int k, int total, int stride, const int64_t *__restrict__ idx, const int *__restrict__ base2,
               unsigned int *__restrict__ rev_tmp) {
  extern __shared__ int sc[];  // [n] cursors
  const int cloud = blockIdx.y, part = blockIdx.x;
  const int per = (total + EC_PARTS - 1) / EC_PARTS;
  const int e0 = min(part * per, total), e1 = min(e0 + per, total);
  const int *bs = base2 + ((size_t)cloud * EC_PARTS + part) * n;
  for (int j = threadIdx.x; j < n; j += 512) sc[j] = bs[j];
  __syncthreads();
  for (int e = e0 + threadIdx.x; e < e1; e += 512) {
    const int j = ec_clamp(idx[(size_t)cloud * total + e], n);
    const int i = e / k, t = e - i * k;
    rev_tmp[(size_t)cloud * stride + atomicAdd(&sc[j], 1)] = ((unsigned int)j << 19) | ((unsigned int)i << 6) | (unsigned int)t;
  }
}

// one thread per entry: its final position by rank counting (entries are distinct).  The fill leaves every run as
// EC_PARTS consecutive segments -- the parts are ranges of ascending edge ids -- so only the entry's own segment
// [base2[part][j], base2[part+1][j]) has to be counted.  Entries of one segment are neighbours in rev_tmp, so a warp walks
// one or two segments together (broadcast loads); hub targets (in-degree of several hundred in feature space) are
// spread over many threads instead of serialising one.
__global__ void __launch_bounds__(256)
ec_sort_kernel(int n, int k, int total, int stride, const int *__restrict__ off, const int *__restrict__ base2,
               const unsigned int *__restrict__ rev_tmp, unsigned int *__restrict__ rev, bool interleave) {
  const int cloud = blockIdx.y;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= total) return;
  const unsigned int *src = rev_tmp + (size_t)cloud * stride;
  const unsigned int mine = src[p];
  const int j = (int)(mine >> 19);
  const int e = (int)((mine >> 6) & 8191u) * k + (int)(mine & 63u);  // edge id = source * k + slot
  const int per = (total + EC_PARTS - 1) / EC_PARTS;
  const int part = e / per;
  const int *bs = base2 + ((size_t)cloud * EC_PARTS + part) * n;
  const int beg = bs[j];
  const int end = part + 1 < EC_PARTS ? bs[n + j] : off[(size_t)cloud * (n + 1) + j + 1];
  int rank = 0;
  int s2 = beg;
  for (; s2 < end && (s2 & 3); ++s2) rank += (src[s2] < mine) ? 1 : 0;
  for (; s2 + 4 <= end; s2 += 4) {  // 16-byte loads: the per-cloud stride is a multiple of 32 entries
    const uint4 v = *reinterpret_cast<const uint4 *>(src + s2);
    rank += ((v.x < mine) ? 1 : 0) + ((v.y < mine) ? 1 : 0) + ((v.z < mine) ? 1 : 0) + ((v.w < mine) ? 1 : 0);
  }
  for (; s2 < end; ++s2) rank += (src[s2] < mine) ? 1 : 0;
  if (interleave)  // the gather backward's format: byte offset of g[source][slot] inside the plane << 12 | target
    rev[(size_t)cloud * stride + edge_pos(beg + rank)] = ((unsigned int)e << 14) | (unsigned int)j;
  else
    rev[(size_t)cloud * stride + beg + rank] = mine;
}

// Segmented sum over the target-sorted edge list in CHUNKS of EC_CHUNK consecutive edges -- uniform work per group
// whatever the in-degrees are.  thread = (chunk slot, quad of channels).  For every run piece inside the chunk the sums
//     gd = sum [slot(i,o) == t] dz(i,o)      gv = sum v(i,o)        over the piece's edges (i,t), in list order
// go to raw[target] when the run STARTS in this chunk, to pbuf[chunk] when it continues from the previous chunk (at most
// one such piece per chunk, the first).  ec_bwd_finish_kernel adds the pieces of a run in chunk order: deterministic.
constexpr int EC_CHUNK = 32;

template <bool TRAIN>
__global__ void __launch_bounds__(EC_THREADS)
ec_bwd_chunk_kernel(int n, int k, int cout, int nchunks, const float *__restrict__ uv,
                    const unsigned int *__restrict__ rev, const float *__restrict__ dz,
                    const unsigned char *__restrict__ slot, float *__restrict__ raw, float *__restrict__ pbuf) {
  const int cloud = blockIdx.y;
  const int tpp = cout >> 2;
  const int groups = EC_THREADS / tpp;
  const int grp = threadIdx.x / tpp, quad = threadIdx.x - grp * tpp;
  const int q = blockIdx.x * groups + grp;
  if (grp >= groups || q >= nchunks) return;
  const int total = n * k;
  const int pos0 = q * EC_CHUNK, cnt = min(EC_CHUNK, total - pos0);
  const unsigned int *revb = rev + (size_t)cloud * nchunks * EC_CHUNK + pos0;
  const size_t pbase = (size_t)cloud * n;
  int cur = (int)(revb[0] >> 19);
  bool cont = pos0 > 0 && (int)(revb[-1] >> 19) == cur;  // the first piece continues a run of the previous chunk
  float gd[4] = {0.f, 0.f, 0.f, 0.f}, gv[4] = {0.f, 0.f, 0.f, 0.f};
  auto flush = [&]() {
    float *dst = cont ? pbuf + ((size_t)cloud * nchunks + q) * 2 * cout : raw + (pbase + cur) * 2 * cout;
    *reinterpret_cast<float4 *>(dst + 4 * quad) = make_float4(gd[0], gd[1], gd[2], gd[3]);
    if (TRAIN) *reinterpret_cast<float4 *>(dst + cout + 4 * quad) = make_float4(gv[0], gv[1], gv[2], gv[3]);
  };
  for (int base = 0; base < cnt; base += 4) {
    int pk[4], tj[4];
    uchar4 s4[4];
    float4 v4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = min(base + u, cnt - 1);
      const unsigned int w = revb[e];
      tj[u] = (int)(w >> 19);
      pk[u] = (int)(((w >> 6) & 8191u) << 8 | (w & 63u));  // source << 8 | slot
      const size_t row = pbase + (pk[u] >> 8);
      s4[u] = *reinterpret_cast<const uchar4 *>(slot + ec_sl(cloud, n, cout, pk[u] >> 8, 4 * quad));
      if (TRAIN) v4[u] = *reinterpret_cast<const float4 *>(uv + row * 2 * cout + cout + 4 * quad);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (base + u < cnt) {
        if (tj[u] != cur) {
          flush();
          cur = tj[u];
          cont = false;
#pragma unroll
          for (int c = 0; c < 4; ++c) gd[c] = gv[c] = 0.f;
        }
        const int t = pk[u] & 255;
        if (s4[u].x == t || s4[u].y == t || s4[u].z == t || s4[u].w == t) {  // ~4 in k quads: dz is read where it lands
          const float4 d4 = *reinterpret_cast<const float4 *>(dz + ec_sl(cloud, n, cout, pk[u] >> 8, 4 * quad));
          gd[0] += (s4[u].x == t) ? d4.x : 0.f;
          gd[1] += (s4[u].y == t) ? d4.y : 0.f;
          gd[2] += (s4[u].z == t) ? d4.z : 0.f;
          gd[3] += (s4[u].w == t) ? d4.w : 0.f;
        }
        if (TRAIN) {
          gv[0] += v4[u].x;
          gv[1] += v4[u].y;
          gv[2] += v4[u].z;
          gv[3] += v4[u].w;
        }
      }
    }
  }
  flush();
}

// The same segmented sum with the per-edge operands served from SHARED memory: a CTA owns one slice of 8 channels of one
// cloud, stages that slice of v (32 B per point) and of slot (8 B per point) and walks all chunks of the cloud's edge
// list.  thread = (chunk slot, quad of the slice).
constexpr int EC_CTHREADS = 1024;  // the chunk kernel's slice buffers allow one CTA per SM: make it a large one

template <bool TRAIN>
__global__ void __launch_bounds__(EC_CTHREADS)
ec_bwd_chunk_staged_kernel(int n, int k, int cout, int nchunks, const float *__restrict__ uv,
                           const unsigned int *__restrict__ rev, const float *__restrict__ dz,
                           const unsigned char *__restrict__ slot, float *__restrict__ raw, float *__restrict__ pbuf) {
  extern __shared__ __align__(16) float4 rows[];  // dz [n][2] float4 | TRAIN: v [n][2] float4 | slot [n][2] uchar4
  const int slice = blockIdx.x, cloud = blockIdx.y;
  const int nquad = cout >> 2;
  const float4 *sdz = rows;
  const float4 *sv = rows + 2 * n;
  uchar4 *sslot = reinterpret_cast<uchar4 *>(rows + (TRAIN ? 4 * n : 2 * n));
  const float *uvb = uv + (size_t)cloud * n * 2 * cout;
  const uchar4 *gslot = reinterpret_cast<const uchar4 *>(slot + ec_sl(cloud, n, cout, 0, 8 * slice));
  const float4 *gdz = reinterpret_cast<const float4 *>(dz + ec_sl(cloud, n, cout, 0, 8 * slice));
  for (int e = threadIdx.x; e < 2 * n; e += EC_CTHREADS) {
    const bool ok = 2 * slice + (e & 1) < nquad;
    rows[e] = ok ? gdz[e] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (TRAIN)
      rows[2 * n + e] = ok ? *reinterpret_cast<const float4 *>(uvb + (size_t)(e >> 1) * 2 * cout + cout +
                                                               4 * (2 * slice + (e & 1)))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
    sslot[e] = ok ? gslot[e] : make_uchar4(255, 255, 255, 255);
  }
  __syncthreads();
  const int cs = threadIdx.x >> 1, qq = threadIdx.x & 1;
  const int quad = 2 * slice + qq;
  if (quad >= nquad) return;
  const int total = n * k;
  const size_t pbase = (size_t)cloud * n;
  for (int q = cs; q < nchunks; q += EC_CTHREADS / 2) {
    const int pos0 = q * EC_CHUNK, cnt = min(EC_CHUNK, total - pos0);
    const unsigned int *revb = rev + (size_t)cloud * nchunks * EC_CHUNK + pos0;  // 128-byte aligned chunk
    // the chunk's 32 entries: eight 16-byte loads issued together (the padding past the list is allocated)
    unsigned int w[EC_CHUNK];
#pragma unroll
    for (int u = 0; u < EC_CHUNK / 4; ++u) {
      const uint4 t4 = reinterpret_cast<const uint4 *>(revb)[u];
      w[4 * u] = t4.x, w[4 * u + 1] = t4.y, w[4 * u + 2] = t4.z, w[4 * u + 3] = t4.w;
    }
    int cur = (int)(w[0] >> 19);
    bool cont = pos0 > 0 && (int)(revb[-1] >> 19) == cur;  // the first piece continues a run of the previous chunk
    float gd[4] = {0.f, 0.f, 0.f, 0.f}, gv[4] = {0.f, 0.f, 0.f, 0.f};
    auto flush = [&]() {
      float *dst = cont ? pbuf + ((size_t)cloud * nchunks + q) * 2 * cout : raw + (pbase + cur) * 2 * cout;
      *reinterpret_cast<float4 *>(dst + 4 * quad) = make_float4(gd[0], gd[1], gd[2], gd[3]);
      if (TRAIN) *reinterpret_cast<float4 *>(dst + cout + 4 * quad) = make_float4(gv[0], gv[1], gv[2], gv[3]);
    };
#pragma unroll
    for (int u = 0; u < EC_CHUNK; ++u) {
      if (u < cnt) {
        const int tj = (int)(w[u] >> 19);
        if (tj != cur) {
          flush();
          cur = tj;
          cont = false;
#pragma unroll
          for (int c = 0; c < 4; ++c) gd[c] = gv[c] = 0.f;
        }
        const int i = (int)((w[u] >> 6) & 8191u);
        const unsigned int t = w[u] & 63u;
        const unsigned int s4 = reinterpret_cast<const unsigned int *>(sslot)[2 * i + qq];
        const unsigned int hit = __vcmpeq4(s4, t * 0x01010101u);  // 0xff in the bytes whose slot is t
        if (hit) {  // ~4 in k quads: dz is read where it lands
          const float4 d4 = sdz[2 * i + qq];
          gd[0] += (hit & 0x000000ffu) ? d4.x : 0.f;
          gd[1] += (hit & 0x0000ff00u) ? d4.y : 0.f;
          gd[2] += (hit & 0x00ff0000u) ? d4.z : 0.f;
          gd[3] += (hit & 0xff000000u) ? d4.w : 0.f;
        }
        if (TRAIN) {
          const float4 v4 = sv[2 * i + qq];
          gv[0] += v4.x;
          gv[1] += v4.y;
          gv[2] += v4.z;
          gv[3] += v4.w;
        }
      }
    }
    flush();
  }
}

// grad [u | v] (B,N,2Cout); thread = (point slot, quad of channels), the point is the TARGET for grad u and the SOURCE
// for grad v.
template <bool TRAIN>
__global__ void __launch_bounds__(EC_THREADS)
ec_bwd_finish_kernel(int n, int k, int cout, int nchunks, const float *__restrict__ uv, const int *__restrict__ off,
                     const float *__restrict__ raw, const float *__restrict__ pbuf, const float *__restrict__ dz,
                     const float *__restrict__ sy, const float *__restrict__ mean, const float *__restrict__ coef,
                     float *__restrict__ guv) {
  const int cloud = blockIdx.y;
  const int i0 = blockIdx.x * EC_PTS;
  const int npts = min(EC_PTS, n - i0);
  const int tpp = cout >> 2;
  const int groups = EC_THREADS / tpp;
  const int grp = threadIdx.x / tpp, quad = threadIdx.x - grp * tpp;
  if (grp >= groups) return;
  const float4 a4 = reinterpret_cast<const float4 *>(coef)[quad];
  const float4 c14 = reinterpret_cast<const float4 *>(coef + cout)[quad];
  const float4 c24 = reinterpret_cast<const float4 *>(coef + 2 * cout)[quad];
  const float4 mu4 = reinterpret_cast<const float4 *>(mean)[quad];
  const float a[4] = {a4.x, a4.y, a4.z, a4.w}, c1[4] = {c14.x, c14.y, c14.z, c14.w},
              c2[4] = {c24.x, c24.y, c24.z, c24.w}, mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w};
  const size_t pbase = (size_t)cloud * n;
  const int *offb = off + (size_t)cloud * (n + 1);
  for (int p = grp; p < npts; p += groups) {
    const int j = i0 + p;
    const int beg = offb[j], end = offb[j + 1];
    float gd[4] = {0.f, 0.f, 0.f, 0.f}, gv[4] = {0.f, 0.f, 0.f, 0.f};
    if (end > beg) {
      const float *r = raw + (pbase + j) * 2 * cout + 4 * quad;
      const float4 d4 = *reinterpret_cast<const float4 *>(r);
      gd[0] = d4.x, gd[1] = d4.y, gd[2] = d4.z, gd[3] = d4.w;
      if (TRAIN) {
        const float4 v4 = *reinterpret_cast<const float4 *>(r + cout);
        gv[0] = v4.x, gv[1] = v4.y, gv[2] = v4.z, gv[3] = v4.w;
      }
      for (int q = beg / EC_CHUNK + 1; q <= (end - 1) / EC_CHUNK; ++q) {  // pieces in later chunks, in order
        const float *pp = pbuf + ((size_t)cloud * nchunks + q) * 2 * cout + 4 * quad;
        const float4 e4 = *reinterpret_cast<const float4 *>(pp);
        gd[0] += e4.x, gd[1] += e4.y, gd[2] += e4.z, gd[3] += e4.w;
        if (TRAIN) {
          const float4 w4 = *reinterpret_cast<const float4 *>(pp + cout);
          gv[0] += w4.x, gv[1] += w4.y, gv[2] += w4.z, gv[3] += w4.w;
        }
      }
    }
    const size_t at = (pbase + j) * cout + 4 * quad;
    const float4 dj4 = *reinterpret_cast<const float4 *>(dz + ec_sl(cloud, n, cout, j, 4 * quad));
    const float dj[4] = {dj4.x, dj4.y, dj4.z, dj4.w};
    float gu[4], gvv[4];
    if (TRAIN) {
      const float deg = (float)(end - beg), kf = (float)k;
      const float4 u4 = *reinterpret_cast<const float4 *>(uv + (pbase + j) * 2 * cout + 4 * quad);
      const float4 sy4 = *reinterpret_cast<const float4 *>(sy + at);
      const float u[4] = {u4.x, u4.y, u4.z, u4.w}, syj[4] = {sy4.x, sy4.y, sy4.z, sy4.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        // sum over incoming edges of (y_e - mean) = deg (u_j - mean) + sum v_i
        gu[c] = a[c] * gd[c] - deg * c1[c] - c2[c] * fmaf(deg, u[c] - mu[c], gv[c]);
        // sum over outgoing edges of (y_e - mean) = sy_j - k mean
        gvv[c] = a[c] * dj[c] - kf * c1[c] - c2[c] * fmaf(-kf, mu[c], syj[c]);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        gu[c] = a[c] * gd[c];
        gvv[c] = a[c] * dj[c];
      }
    }
    float *g = guv + (pbase + j) * 2 * cout + 4 * quad;
    *reinterpret_cast<float4 *>(g) = make_float4(gu[0], gu[1], gu[2], gu[3]);
    *reinterpret_cast<float4 *>(g + cout) = make_float4(gvv[0], gvv[1], gvv[2], gvv[3]);
  }
}

// Target-sorted edge lists for callers outside this file (graph.cu's deterministic gather backward).  With pieces > 1
// the SOURCES of every cloud are cut into `pieces` equal ranges and every (cloud, piece) is sorted on its own (sources
// are packed relative to the piece): sub-cloud s = cloud * pieces + piece has its run starts at off + s * (n + 1) and
// its list at rev + s * stride.  The workspace holds off | hist2 | base2 | rev_tmp | rev.
static inline size_t es_up(size_t v) { return (v + 255) & ~(size_t)255; }
static inline int es_stride(int total) { return (total + 511) / 512 * 512; }  // whole interleave blocks
size_t edge_sort_ws_bytes(int b, int n, int k, int pieces) {
  const size_t subs = (size_t)b * pieces;
  return es_up(sizeof(int) * subs * (n + 1)) + 2 * es_up(sizeof(int) * subs * EC_PARTS * n) +
         2 * es_up(sizeof(int) * subs * es_stride(n / pieces * k));
}
bool edge_sort_ok(int b, int n, int k, int pieces) {
  return b > 0 && n > 0 && n <= EC_MAX_N && k > 0 && k <= EC_MAX_K && pieces > 0 && n % pieces == 0 &&
         (long long)b * pieces <= 65535;
}
// the views of a workspace edge_sort_launch has filled (same carve-up)
void edge_sort_views(int b, int n, int k, int pieces, const char *ws, const int **off_out, const unsigned int **rev_out,
                     int *stride) {
  const int subs = b * pieces, total = n / pieces * k, estride = es_stride(total);
  const size_t cnt_bytes = es_up(sizeof(int) * (size_t)subs * EC_PARTS * n), rev_bytes = es_up(sizeof(int) * (size_t)subs * estride);
  *off_out = reinterpret_cast<const int *>(ws);
  const char *w = ws + es_up(sizeof(int) * (size_t)subs * (n + 1));
  *rev_out = reinterpret_cast<const unsigned int *>(w + 2 * cnt_bytes + rev_bytes);
  *stride = estride;
}
void edge_sort_launch(int b, int n, int k, int pieces, const int64_t *idx, char *ws, bool interleave, const int **off_out,
                      const unsigned int **rev_out, int *stride, cudaStream_t st) {
  const int subs = b * pieces, total = n / pieces * k, estride = es_stride(total);
  const size_t cnt_bytes = es_up(sizeof(int) * (size_t)subs * EC_PARTS * n), rev_bytes = es_up(sizeof(int) * (size_t)subs * estride);
  int *off = reinterpret_cast<int *>(ws);
  char *w = ws + es_up(sizeof(int) * (size_t)subs * (n + 1));
  int *hist2 = reinterpret_cast<int *>(w);
  int *base2 = reinterpret_cast<int *>(w + cnt_bytes);
  unsigned int *rev_tmp = reinterpret_cast<unsigned int *>(w + 2 * cnt_bytes);
  unsigned int *rev = reinterpret_cast<unsigned int *>(w + 2 * cnt_bytes + rev_bytes);
  const dim3 egrid((total + 255) / 256, subs), pgrid(EC_PARTS, subs);
  ec_hist_kernel<<<pgrid, 512, n * sizeof(int), st>>>(n, total, idx, hist2);
  ec_scan_kernel<<<subs, 1024, 0, st>>>(n, total, hist2, off, base2);
  ec_fill_kernel<<<pgrid, 512, n * sizeof(int), st>>>(n, k, total, estride, idx, base2, rev_tmp);
  ec_sort_kernel<<<egrid, 256, 0, st>>>(n, k, total, estride, off, base2, rev_tmp, rev, interleave);
  *off_out = off;
  *rev_out = rev;
  *stride = estride;
}

static bool ec_shape_ok(int b, int n, int k, int cout) {
  return b > 0 && b <= 65535 && n > 0 && n <= EC_MAX_N && k > 0 && k <= EC_MAX_K && cout >= 4 && cout <= 1024 &&
         cout % 4 == 0;
}

}  // namespace pcc

using namespace pcc;

extern "C" __attribute__((visibility("default"))) int
pcc_edgeconv_forward(int b, int n, int k, int cout, const float *uv, const int64_t *idx, const float *gamma,
                     const float *beta, float *running_mean, float *running_var, int bn_mode, float momentum, float eps,
                     int act, float slope, float *out, float *exty, float *sy, unsigned char *slot, float *mean,
                     float *invstd, pcc_stream_t stream) {
  if (b == 0 || n == 0) return PCC_OK;
  if (!ec_shape_ok(b, n, k, cout)) return PCC_ENOTSUP;
  if (bn_mode < 0 || bn_mode > 2 || act < 0 || act > 1 || (act == EC_ACT_LEAKY && slope < 0.f)) return PCC_EINVAL;
  if (bn_mode == EC_BN_EVAL && (!running_mean || !running_var)) return PCC_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const bool staged = n <= EC_STAGE_MAX_N;
  const dim3 grid((n + EC_PTS - 1) / EC_PTS, b), sgrid(ec_nslice(cout), b);
  const int nparts = staged ? b : (int)(grid.x * grid.y);
  const size_t stage_smem = (size_t)max(n * 32, 2 * EC_STHREADS * 4 * 8) + (size_t)EC_SPTS * k * 2;
  if (staged) {
    static size_t attr1[64], attr2[64];
    cudaError_t e1 = smem_optin(ec_reduce_staged_kernel<true>, EC_STAGE_MAX_N * 32 + EC_SPTS * EC_MAX_K * 2, attr1);
    cudaError_t e2 = smem_optin(ec_reduce_staged_kernel<false>, EC_STAGE_MAX_N * 32 + EC_SPTS * EC_MAX_K * 2, attr2);
    if (e1 != cudaSuccess || e2 != cudaSuccess) return (int)(e1 != cudaSuccess ? e1 : e2);
  }
  char *ws = nullptr;
  const size_t part_bytes = sizeof(double) * 2 * cout * nparts, aff_bytes = sizeof(float) * 2 * cout;
  cudaError_t e = ws_alloc((void **)&ws, part_bytes + aff_bytes, st);
  if (e != cudaSuccess) return (int)e;
  double *partials = reinterpret_cast<double *>(ws);
  float *scale = reinterpret_cast<float *>(ws + part_bytes), *shift = scale + cout;
  if (staged && bn_mode == EC_BN_TRAIN)
    ec_reduce_staged_kernel<true><<<sgrid, EC_STHREADS, stage_smem, st>>>(n, k, cout, uv, idx, gamma, exty, sy, slot,
                                                                        partials, nparts);
  else if (staged)
    ec_reduce_staged_kernel<false><<<sgrid, EC_STHREADS, stage_smem, st>>>(n, k, cout, uv, idx, gamma, exty, sy, slot,
                                                                         partials, nparts);
  else if (bn_mode == EC_BN_TRAIN)
    ec_reduce_kernel<true><<<grid, EC_THREADS, 0, st>>>(n, k, cout, uv, idx, gamma, exty, sy, slot, partials, nparts);
  else
    ec_reduce_kernel<false><<<grid, EC_THREADS, 0, st>>>(n, k, cout, uv, idx, gamma, exty, sy, slot, partials, nparts);
  ec_stats_kernel<<<cout, 128, 0, st>>>(cout, nparts, (double)b * n * k, bn_mode, partials, gamma, beta, running_mean,
                                        running_var, momentum, eps, mean, invstd, scale, shift);
  ec_finalize_kernel<<<dim3((n + 31) / 32, (cout + 31) / 32, b), 256, 0, st>>>(n, cout, exty, scale, shift, act, slope,
                                                                              out);
  cudaFreeAsync(ws, st);
  return finish_launch(3);
}

extern "C" __attribute__((visibility("default"))) int
pcc_edgeconv_backward(int b, int n, int k, int cout, const float *uv, const int64_t *idx, const float *gamma,
                      const float *beta, const float *mean, const float *invstd, int bn_mode, int act, float slope,
                      const float *exty, const float *sy, const unsigned char *slot, const float *grad_out,
                      float *grad_uv, float *grad_gamma, float *grad_beta, pcc_stream_t stream) {
  if (b == 0 || n == 0) return PCC_OK;
  if (!ec_shape_ok(b, n, k, cout)) return PCC_ENOTSUP;
  if (bn_mode < 0 || bn_mode > 2 || act < 0 || act > 1) return PCC_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 tgrid((n + 31) / 32, (cout + 31) / 32, b);
  const int nparts = (int)(tgrid.x * b);
  const int per_cloud = n * k, nchunks = (per_cloud + EC_CHUNK - 1) / EC_CHUNK;
  const size_t total = (size_t)b * per_cloud;
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t part_bytes = up(sizeof(double) * 2 * cout * nparts), dz_bytes = up(sizeof(float) * (size_t)b * ec_nslice(cout) * n * 8),
               coef_bytes = up(sizeof(float) * 4 * cout), off_bytes = up(sizeof(int) * (size_t)b * (n + 1)),
               cnt_bytes = up(sizeof(int) * (size_t)b * EC_PARTS * n), rev_bytes = up(sizeof(int) * (size_t)b * nchunks * EC_CHUNK),
               raw_bytes = up(sizeof(float) * (size_t)b * n * 2 * cout),
               pbuf_bytes = up(sizeof(float) * (size_t)b * nchunks * 2 * cout);
  char *ws = nullptr;
  cudaError_t e = ws_alloc((void **)&ws, part_bytes + dz_bytes + coef_bytes + off_bytes + 2 * cnt_bytes +
                                                    2 * rev_bytes + raw_bytes + pbuf_bytes, st);
  if (e != cudaSuccess) return (int)e;
  char *w = ws;
  auto take = [&](size_t bytes) {
    char *r = w;
    w += bytes;
    return r;
  };
  double *partials = reinterpret_cast<double *>(take(part_bytes));
  float *dz = reinterpret_cast<float *>(take(dz_bytes));
  float *coef = reinterpret_cast<float *>(take(coef_bytes));
  int *off = reinterpret_cast<int *>(take(off_bytes));
  int *hist2 = reinterpret_cast<int *>(take(cnt_bytes));
  int *base2 = reinterpret_cast<int *>(take(cnt_bytes));
  unsigned int *rev_tmp = reinterpret_cast<unsigned int *>(take(rev_bytes));
  unsigned int *rev = reinterpret_cast<unsigned int *>(take(rev_bytes));
  const int estride = nchunks * EC_CHUNK;  // per-cloud stride of the edge lists
  float *raw = reinterpret_cast<float *>(take(raw_bytes));
  float *pbuf = reinterpret_cast<float *>(take(pbuf_bytes));
  ec_bwd_point_kernel<<<tgrid, 256, 0, st>>>(n, cout, grad_out, exty, mean, invstd, gamma, beta, act, slope, dz,
                                             partials, nparts);
  ec_bwd_stats_kernel<<<cout, 128, 0, st>>>(cout, nparts, (double)b * n * k, bn_mode, partials, gamma, invstd,
                                            grad_gamma, grad_beta, coef);
  const dim3 egrid((per_cloud + 255) / 256, b), pgrid(EC_PARTS, b);
  ec_hist_kernel<<<pgrid, 512, n * sizeof(int), st>>>(n, per_cloud, idx, hist2);
  ec_scan_kernel<<<b, 1024, 0, st>>>(n, per_cloud, hist2, off, base2);
  ec_fill_kernel<<<pgrid, 512, n * sizeof(int), st>>>(n, k, per_cloud, estride, idx, base2, rev_tmp);
  ec_sort_kernel<<<egrid, 256, 0, st>>>(n, k, per_cloud, estride, off, base2, rev_tmp, rev, false);
  const int groups = EC_THREADS / (cout >> 2);
  const dim3 cgrid((nchunks + groups - 1) / groups, b), grid((n + EC_PTS - 1) / EC_PTS, b);
  const bool staged = n <= EC_STAGE_MAX_N;
  const dim3 sgrid(ec_nslice(cout), b);
  if (staged) {
    static size_t attr1[64], attr2[64];
    cudaError_t e1 = smem_optin(ec_bwd_chunk_staged_kernel<true>, EC_STAGE_MAX_N * 72, attr1);
    cudaError_t e2 = smem_optin(ec_bwd_chunk_staged_kernel<false>, EC_STAGE_MAX_N * 40, attr2);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
      cudaFreeAsync(ws, st);
      return (int)(e1 != cudaSuccess ? e1 : e2);
    }
  }
  if (staged && bn_mode == EC_BN_TRAIN) {
    ec_bwd_chunk_staged_kernel<true><<<sgrid, EC_CTHREADS, (size_t)n * 72, st>>>(n, k, cout, nchunks, uv, rev, dz, slot,
                                                                                raw, pbuf);
    ec_bwd_finish_kernel<true><<<grid, EC_THREADS, 0, st>>>(n, k, cout, nchunks, uv, off, raw, pbuf, dz, sy, mean, coef,
                                                            grad_uv);
  } else if (staged) {
    ec_bwd_chunk_staged_kernel<false><<<sgrid, EC_CTHREADS, (size_t)n * 40, st>>>(n, k, cout, nchunks, uv, rev, dz, slot,
                                                                                raw, pbuf);
    ec_bwd_finish_kernel<false><<<grid, EC_THREADS, 0, st>>>(n, k, cout, nchunks, uv, off, raw, pbuf, dz, sy, mean,
                                                             coef, grad_uv);
  } else if (bn_mode == EC_BN_TRAIN) {
    ec_bwd_chunk_kernel<true><<<cgrid, EC_THREADS, 0, st>>>(n, k, cout, nchunks, uv, rev, dz, slot, raw, pbuf);
    ec_bwd_finish_kernel<true><<<grid, EC_THREADS, 0, st>>>(n, k, cout, nchunks, uv, off, raw, pbuf, dz, sy, mean, coef,
                                                            grad_uv);
  } else {
    ec_bwd_chunk_kernel<false><<<cgrid, EC_THREADS, 0, st>>>(n, k, cout, nchunks, uv, rev, dz, slot, raw, pbuf);
    ec_bwd_finish_kernel<false><<<grid, EC_THREADS, 0, st>>>(n, k, cout, nchunks, uv, off, raw, pbuf, dz, sy, mean,
                                                             coef, grad_uv);
  }
  cudaFreeAsync(ws, st);
  return finish_launch(8);
}
