// Standalone probe for DESIGN.md section 6 item 6 (NOT part of libpcc_b200.so, written at the end of round 1 and not yet
// run on a GPU): exact nearest neighbour of every point of cloud A in cloud B through a uniform grid, against a
// brute-force kernel with the reference's arithmetic and tie rule (nndistance.cu:2-124: d = fma(dz,dz,fma(dx,dx,dy*dy)),
// d* = ref - query, strict '<' over ascending indices).  tools/grid_sim.py --exact is the CPU model of the same search.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/grid_nn_probe tools/grid_nn_probe.cu && tools/grid_nn_probe [h]
//
// Pipeline per batch: bbox_kernel (joint bounding box of each cloud pair -> grid geometry), build_kernel (counting sort
// of each cloud by cell in shared memory: sorted (x,y,z,index) + cell offsets), search_kernel (one warp per query, in
// cell order: the 27 neighbouring cells as 9 contiguous runs, lanes stride over the candidates, (distance, index)
// minimum by two REDUX; queries whose best is not strictly inside the block bound rescan the whole cloud).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <random>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

constexpr int GMAX = 32;                       // cells per axis at most
constexpr int NCELL_MAX = GMAX * GMAX * GMAX;  // 32768 -> 128 KiB of shared-memory histogram
struct GridMeta {
  float lo[3];
  float h, inv_h;
  int dim[3];
};

__device__ __forceinline__ float sqdist1(float qx, float qy, float qz, float rx, float ry, float rz) {
  const float dx = rx - qx, dy = ry - qy, dz = rz - qz;
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// one CTA per cloud pair: joint bounding box -> grid origin, cell size (enlarged if the extent needs more than GMAX cells)
__global__ void __launch_bounds__(256) bbox_kernel(int n, const float *xyz1, int m, const float *xyz2, float h, GridMeta *meta) {
  __shared__ float smin[3][8], smax[3][8];
  const size_t cloud = blockIdx.x;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = threadIdx.x; i < n + m; i += 256) {
    const float *p = i < n ? xyz1 + (cloud * n + i) * 3 : xyz2 + (cloud * m + (i - n)) * 3;
    for (int a = 0; a < 3; ++a) {
      mn[a] = fminf(mn[a], p[a]);
      mx[a] = fmaxf(mx[a], p[a]);
    }
  }
  for (int a = 0; a < 3; ++a) {
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
    if ((threadIdx.x & 31) == 0) {
      smin[a][threadIdx.x >> 5] = mn[a];
      smax[a][threadIdx.x >> 5] = mx[a];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    GridMeta g;
    float ext = 0.f;
    for (int a = 0; a < 3; ++a) {
      float lo = smin[a][0], hi = smax[a][0];
      for (int w = 1; w < 8; ++w) {
        lo = fminf(lo, smin[a][w]);
        hi = fmaxf(hi, smax[a][w]);
      }
      g.lo[a] = lo;
      smax[a][0] = hi;
      ext = fmaxf(ext, hi - lo);
    }
    g.h = fmaxf(h, ext / (float)(GMAX - 1));
    g.inv_h = 1.0f / g.h;
    for (int a = 0; a < 3; ++a) g.dim[a] = min(GMAX, (int)floorf((smax[a][0] - g.lo[a]) * g.inv_h) + 1);
    meta[cloud] = g;
  }
}

__device__ __forceinline__ int cell_of(const GridMeta &g, float x, float y, float z, float *inner) {
  const float f[3] = {(x - g.lo[0]) * g.inv_h, (y - g.lo[1]) * g.inv_h, (z - g.lo[2]) * g.inv_h};
  int c[3];
  float in = 1.f;
  for (int a = 0; a < 3; ++a) {
    c[a] = min(g.dim[a] - 1, max(0, (int)floorf(f[a])));
    in = fminf(in, fminf(f[a] - (float)c[a], (float)(c[a] + 1) - f[a]));  // negative in a clamped cell -> forces the rescan
  }
  if (inner) *inner = in;
  return (c[0] * g.dim[1] + c[1]) * g.dim[2] + c[2];
}

// one CTA per (cloud, side): counting sort by cell.  sorted[cloud][t] = (x, y, z, original index), cell_start[cloud][ncell+1]
__global__ void __launch_bounds__(1024) build_kernel(int npts, const float *xyz, const GridMeta *meta, float4 *sorted,
                                                      int *cell_start) {
  extern __shared__ int hist[];  // [ncell + 1]
  __shared__ int wsum[32];
  const size_t cloud = blockIdx.x;
  const GridMeta g = meta[cloud];
  const int ncell = g.dim[0] * g.dim[1] * g.dim[2];
  const float *p = xyz + cloud * (size_t)npts * 3;
  for (int i = threadIdx.x; i <= ncell; i += 1024) hist[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < npts; i += 1024) atomicAdd(&hist[cell_of(g, p[i * 3], p[i * 3 + 1], p[i * 3 + 2], nullptr)], 1);
  __syncthreads();
  // exclusive scan: each thread owns a contiguous chunk of cells
  const int per = (ncell + 1023) / 1024;
  const int c0 = min(ncell, (int)threadIdx.x * per), c1 = min(ncell, c0 + per);
  int s = 0;
  for (int c = c0; c < c1; ++c) s += hist[c];
  int incl = s;
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += v;
  }
  if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = wsum[threadIdx.x];
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, o);
      if (threadIdx.x >= o) v += u;
    }
    wsum[threadIdx.x] = v;
  }
  __syncthreads();
  int run = incl - s + (threadIdx.x >= 32 ? wsum[(threadIdx.x >> 5) - 1] : 0);
  for (int c = c0; c < c1; ++c) {
    const int cnt = hist[c];
    hist[c] = run;
    run += cnt;
  }
  if (threadIdx.x == 0) hist[ncell] = npts;
  __syncthreads();
  int *cs = cell_start + cloud * (size_t)(NCELL_MAX + 1);
  for (int i = threadIdx.x; i <= ncell; i += 1024) cs[i] = hist[i];
  __syncthreads();
  float4 *out = sorted + cloud * (size_t)npts;
  for (int i = threadIdx.x; i < npts; i += 1024) {
    const float x = p[i * 3], y = p[i * 3 + 1], z = p[i * 3 + 2];
    const int pos = atomicAdd(&hist[cell_of(g, x, y, z, nullptr)], 1);  // order inside a cell is arbitrary: the search
    out[pos] = make_float4(x, y, z, __int_as_float(i));                 // takes the (distance, index) minimum
  }
}

__device__ __forceinline__ void warp_best(float &bd, int &bi) {
  const unsigned int bits = __float_as_uint(bd);  // d >= 0: unsigned order == float order (+inf included)
  const unsigned int mn = __reduce_min_sync(0xffffffffu, bits);
  const unsigned int cand = bits == mn ? (unsigned int)bi : 0x7fffffffu;
  bi = (int)__reduce_min_sync(0xffffffffu, cand);
  bd = __uint_as_float(mn);
}

constexpr int SW = 8;  // warps per CTA in the search kernel
__global__ void __launch_bounds__(SW * 32) search_kernel(int nq, const float4 *sortedQ, int nr, const float4 *sortedR,
                                                         const int *cell_startR, const GridMeta *meta, float *dist, int *idx,
                                                         unsigned int *rescans) {
  const size_t cloud = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * SW + (threadIdx.x >> 5);
  if (i >= nq) return;
  const GridMeta g = meta[cloud];
  const float4 q = sortedQ[cloud * (size_t)nq + i];
  const float4 *R = sortedR + cloud * (size_t)nr;
  const int *cs = cell_startR + cloud * (size_t)(NCELL_MAX + 1);
  float inner;
  const int cell = cell_of(g, q.x, q.y, q.z, &inner);
  const int cz = cell % g.dim[2], cy = (cell / g.dim[2]) % g.dim[1], cx = cell / (g.dim[2] * g.dim[1]);
  float bd = INFINITY;
  int bi = 0x7fffffff;
  for (int dx = -1; dx <= 1; ++dx)
    for (int dy = -1; dy <= 1; ++dy) {
      const int x = cx + dx, y = cy + dy;
      if (x < 0 || x >= g.dim[0] || y < 0 || y >= g.dim[1]) continue;  // uniform across the warp
      const int base = (x * g.dim[1] + y) * g.dim[2];
      const int s = cs[base + max(cz - 1, 0)], e = cs[base + min(cz + 1, g.dim[2] - 1) + 1];
      for (int t = s + lane; t < e; t += 32) {
        const float4 r = R[t];
        const float d = sqdist1(q.x, q.y, q.z, r.x, r.y, r.z);
        const int id = __float_as_int(r.w);
        if (d < bd || (d == bd && id < bi)) {
          bd = d;
          bi = id;
        }
      }
    }
  warp_best(bd, bi);
  // every point outside the 27-cell block is at least h (1 + inner) away; margin far above the fp32 rounding of d and
  // of the cell arithmetic; STRICT test so that an equal distance outside the block (possibly a lower index) is seen
  const float bound = g.h * (1.0f + inner);
  if (!(inner > 0.f && bd < bound * bound * (1.0f - 1e-4f))) {
    if (lane == 0) atomicAdd(rescans, 1u);
    bd = INFINITY;
    bi = 0x7fffffff;
    for (int t = lane; t < nr; t += 32) {
      const float4 r = R[t];
      const float d = sqdist1(q.x, q.y, q.z, r.x, r.y, r.z);
      const int id = __float_as_int(r.w);
      if (d < bd || (d == bd && id < bi)) {
        bd = d;
        bi = id;
      }
    }
    warp_best(bd, bi);
  }
  if (lane == 0) {
    const int qi = __float_as_int(q.w);
    dist[cloud * (size_t)nq + qi] = bd;
    idx[cloud * (size_t)nq + qi] = bi;
  }
}

// checker: the reference's scan (one thread per query, ascending references, strict '<')
__global__ void brute_kernel(int nq, const float *xq, int nr, const float *xr, float *dist, int *idx) {
  const size_t cloud = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const float *q = xq + (cloud * nq + i) * 3, *r = xr + cloud * (size_t)nr * 3;
  float bd = 0.f;
  int bi = 0;
  for (int k = 0; k < nr; ++k) {
    const float d = sqdist1(q[0], q[1], q[2], r[k * 3], r[k * 3 + 1], r[k * 3 + 2]);
    if (k == 0 || d < bd) {
      bd = d;
      bi = k;
    }
  }
  dist[cloud * (size_t)nq + i] = bd;
  idx[cloud * (size_t)nq + i] = bi;
}

static void make_clouds(int b, int n, int kind, std::vector<float> &a, std::vector<float> &c) {
  std::mt19937 gen(1234 + kind);
  std::normal_distribution<float> nd(0.f, 1.f);
  a.assign((size_t)b * n * 3, 0.f);
  c.assign((size_t)b * n * 3, 0.f);
  const float sc[3] = {1.f, 0.6f, 0.3f};
  auto normalise = [&](float *p) {
    float mean[3] = {0, 0, 0}, mx = 0.f;
    for (int i = 0; i < n; ++i)
      for (int d = 0; d < 3; ++d) mean[d] += p[i * 3 + d] / n;
    for (int i = 0; i < n; ++i) {
      float s = 0.f;
      for (int d = 0; d < 3; ++d) {
        p[i * 3 + d] -= mean[d];
        s += p[i * 3 + d] * p[i * 3 + d];
      }
      mx = std::max(mx, std::sqrt(s));
    }
    for (int i = 0; i < n * 3; ++i) p[i] /= mx;
  };
  for (int bb = 0; bb < b; ++bb) {
    float *pa = a.data() + (size_t)bb * n * 3, *pc = c.data() + (size_t)bb * n * 3;
    for (int i = 0; i < n; ++i)
      for (int d = 0; d < 3; ++d) pc[i * 3 + d] = nd(gen) * (kind == 1 ? 1.f : sc[d]);
    normalise(pc);
    if (kind == 0) {  // S1 "near": permuted reference + noise
      std::vector<int> perm(n);
      for (int i = 0; i < n; ++i) perm[i] = i;
      std::shuffle(perm.begin(), perm.end(), gen);
      for (int i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) pa[i * 3 + d] = pc[perm[i] * 3 + d] + 0.02f * nd(gen);
    } else if (kind == 1) {  // S2 "far": independent cloud
      for (int i = 0; i < n * 3; ++i) pa[i] = nd(gen);
      normalise(pa);
    } else {  // S3 "ties": 1/64 grid, drawn with replacement from 1/4 of the points
      for (int i = 0; i < n * 3; ++i) pc[i] = std::round(pc[i] * 64.f) / 64.f;
      std::uniform_int_distribution<int> pick(0, n / 4 - 1);
      for (int i = 0; i < n; ++i) {
        const int j = pick(gen), k = pick(gen);
        for (int d = 0; d < 3; ++d) pa[i * 3 + d] = pc[j * 3 + d];
        if (i >= n / 4)
          for (int d = 0; d < 3; ++d) pc[i * 3 + d] = pc[k * 3 + d];
      }
    }
  }
}

int main(int argc, char **argv) {
  const float h = argc > 1 ? (float)atof(argv[1]) : 0.08f;
  const int b = 32, n = 2048;
  float *d1, *d2, *dist[2], *bdist[2];
  int *idx[2], *bidx[2], *cstart[2];
  float4 *sorted[2];
  GridMeta *meta;
  unsigned int *rescans;
  const size_t pts = (size_t)b * n;
  CK(cudaMalloc(&d1, pts * 12));
  CK(cudaMalloc(&d2, pts * 12));
  CK(cudaMalloc(&meta, sizeof(GridMeta) * b));
  CK(cudaMalloc(&rescans, 4));
  for (int s = 0; s < 2; ++s) {
    CK(cudaMalloc(&dist[s], pts * 4));
    CK(cudaMalloc(&bdist[s], pts * 4));
    CK(cudaMalloc(&idx[s], pts * 4));
    CK(cudaMalloc(&bidx[s], pts * 4));
    CK(cudaMalloc(&sorted[s], pts * 16));
    CK(cudaMalloc(&cstart[s], sizeof(int) * (size_t)b * (NCELL_MAX + 1)));
  }
  const int build_smem = sizeof(int) * (NCELL_MAX + 1);
  CK(cudaFuncSetAttribute(build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, build_smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  auto run = [&]() {
    bbox_kernel<<<b, 256>>>(n, d1, n, d2, h, meta);
    build_kernel<<<b, 1024, build_smem>>>(n, d1, meta, sorted[0], cstart[0]);
    build_kernel<<<b, 1024, build_smem>>>(n, d2, meta, sorted[1], cstart[1]);
    search_kernel<<<dim3((n + SW - 1) / SW, b), SW * 32>>>(n, sorted[0], n, sorted[1], cstart[1], meta, dist[0], idx[0], rescans);
    search_kernel<<<dim3((n + SW - 1) / SW, b), SW * 32>>>(n, sorted[1], n, sorted[0], cstart[0], meta, dist[1], idx[1], rescans);
  };
  int bad_total = 0;
  const char *names[3] = {"S1 near", "S2 far", "S3 ties"};
  for (int kind = 0; kind < 3; ++kind) {
    std::vector<float> a, c;
    make_clouds(b, n, kind, a, c);
    CK(cudaMemcpy(d1, a.data(), pts * 12, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d2, c.data(), pts * 12, cudaMemcpyHostToDevice));
    CK(cudaMemset(rescans, 0, 4));
    run();
    brute_kernel<<<dim3((n + 127) / 128, b), 128>>>(n, d1, n, d2, bdist[0], bidx[0]);
    brute_kernel<<<dim3((n + 127) / 128, b), 128>>>(n, d2, n, d1, bdist[1], bidx[1]);
    CK(cudaDeviceSynchronize());
    unsigned int nres = 0;
    CK(cudaMemcpy(&nres, rescans, 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    std::vector<float> hd(pts), hb(pts);
    std::vector<int> hi(pts), hbi(pts);
    for (int s = 0; s < 2; ++s) {
      CK(cudaMemcpy(hd.data(), dist[s], pts * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hb.data(), bdist[s], pts * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hi.data(), idx[s], pts * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hbi.data(), bidx[s], pts * 4, cudaMemcpyDeviceToHost));
      for (size_t t = 0; t < pts; ++t) bad += (hi[t] != hbi[t]) || (hd[t] != hb[t]);
    }
    for (int w = 0; w < 3; ++w) run();
    CK(cudaEventRecord(e0));
    for (int r = 0; r < 20; ++r) run();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-8s h=%.3f  mismatches=%d of %zu  rescanned queries=%u of %zu  grid forward (bbox+2 builds+2 searches): %.1f us\n",
           names[kind], h, bad, 2 * pts, nres, 2 * pts, ms / 20 * 1e3);
    bad_total += bad;
  }
  return bad_total ? 1 : 0;
}
