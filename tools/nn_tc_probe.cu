// Standalone harness for csrc/chamfer_tc.cu (tcgen05 Chamfer search): bit-exactness against a brute-force kernel with the
// reference's arithmetic and tie rule on the S1 / S2 / S3 input families, the measured score error that NT_CEPS bounds,
// exact-rescan / resolved-chunk statistics and timing.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DNT_DEBUG_SCORES -o tools/nn_tc_probe tools/nn_tc_probe.cu
#include "../pointcloudcounterfactual_b200/csrc/chamfer_tc.cu"

#include <math.h>
#include <stdio.h>

#include <algorithm>
#include <random>
#include <vector>

namespace pcc {
std::atomic<uint64_t> g_launches{0};
std::atomic<uint64_t> g_routes[R_COUNT];
cudaError_t ws_alloc(void **ptr, size_t bytes, cudaStream_t st) { return cudaMallocAsync(ptr, bytes, st); }
}  // namespace pcc

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                        \
    }                                                                                 \
  } while (0)

__global__ void brute_kernel(int nq, const float *xq, int nr, const float *xr, float *dist, int *idx) {
  const size_t cloud = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const float *q = xq + (cloud * nq + i) * 3, *r = xr + cloud * (size_t)nr * 3;
  float bd = 0.f;
  int bi = 0;
  for (int k = 0; k < nr; ++k) {
    const float d = pcc::sqdist1(q[0], q[1], q[2], r[k * 3], r[k * 3 + 1], r[k * 3 + 2]);
    if (k == 0 || d < bd) {
      bd = d;
      bi = k;
    }
  }
  dist[cloud * (size_t)nq + i] = bd;
  idx[cloud * (size_t)nq + i] = bi;
}

static void make_clouds(int b, int n, int m, int kind, std::vector<float> &a, std::vector<float> &c) {
  std::mt19937 gen(1234 + kind);
  std::normal_distribution<float> nd(0.f, 1.f);
  a.assign((size_t)b * n * 3, 0.f);
  c.assign((size_t)b * m * 3, 0.f);
  const float sc[3] = {1.f, 0.6f, 0.3f};
  auto normalise = [&](float *p, int cnt) {
    float mean[3] = {0, 0, 0}, mx = 0.f;
    for (int i = 0; i < cnt; ++i)
      for (int d = 0; d < 3; ++d) mean[d] += p[i * 3 + d] / cnt;
    for (int i = 0; i < cnt; ++i) {
      float s = 0.f;
      for (int d = 0; d < 3; ++d) {
        p[i * 3 + d] -= mean[d];
        s += p[i * 3 + d] * p[i * 3 + d];
      }
      mx = std::max(mx, std::sqrt(s));
    }
    for (int i = 0; i < cnt * 3; ++i) p[i] /= mx;
  };
  for (int bb = 0; bb < b; ++bb) {
    float *pa = a.data() + (size_t)bb * n * 3, *pc = c.data() + (size_t)bb * m * 3;
    for (int i = 0; i < m; ++i)
      for (int d = 0; d < 3; ++d) pc[i * 3 + d] = nd(gen) * (kind == 1 ? 1.f : sc[d]);
    normalise(pc, m);
    if (kind == 0) {  // S1 "near": permuted reference + noise
      std::vector<int> perm(m);
      for (int i = 0; i < m; ++i) perm[i] = i;
      std::shuffle(perm.begin(), perm.end(), gen);
      for (int i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) pa[i * 3 + d] = pc[perm[i % m] * 3 + d] + 0.02f * nd(gen);
    } else if (kind == 1) {  // S2 "far": independent cloud
      for (int i = 0; i < n * 3; ++i) pa[i] = nd(gen);
      normalise(pa, n);
    } else if (kind == 2) {  // S3 "ties": 1/64 grid, drawn with replacement from 1/4 of the points
      for (int i = 0; i < m * 3; ++i) pc[i] = std::round(pc[i] * 64.f) / 64.f;
      std::uniform_int_distribution<int> pick(0, m / 4 - 1);
      for (int i = 0; i < n; ++i) {
        const int j = pick(gen);
        for (int d = 0; d < 3; ++d) pa[i * 3 + d] = pc[j * 3 + d];
      }
      for (int i = m / 4; i < m; ++i) {
        const int k = pick(gen);
        for (int d = 0; d < 3; ++d) pc[i * 3 + d] = pc[k * 3 + d];
      }
    } else {  // collapsed: every point identical, far from the origin
      for (int i = 0; i < n * 3; ++i) pa[i] = 5.f + (i % 3);
      for (int i = 0; i < m * 3; ++i) pc[i] = 5.f + (i % 3);
    }
  }
}

int main(int argc, char **argv) {
  const int b = argc > 1 ? atoi(argv[1]) : 32, n = argc > 2 ? atoi(argv[2]) : 2048, m = argc > 3 ? atoi(argv[3]) : 2048;
  float *d1, *d2, *dist[2], *bdist[2], *dbg;
  int *idx[2], *bidx[2];
  unsigned int *stats;
  const size_t p1 = (size_t)b * n, p2 = (size_t)b * m;
  const int npad2 = pcc::nt_pad(m);
  CK(cudaMalloc(&d1, p1 * 12));
  CK(cudaMalloc(&d2, p2 * 12));
  CK(cudaMalloc(&stats, 512));
  CK(cudaMalloc(&dbg, (size_t)n * npad2 * 4));
  const size_t cnt[2] = {p1, p2};
  for (int s = 0; s < 2; ++s) {
    CK(cudaMalloc(&dist[s], cnt[s] * 4));
    CK(cudaMalloc(&bdist[s], cnt[s] * 4));
    CK(cudaMalloc(&idx[s], cnt[s] * 4));
    CK(cudaMalloc(&bidx[s], cnt[s] * 4));
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  int bad_total = 0;
  const char *names[4] = {"S1 near", "S2 far", "S3 ties", "collapsed"};
  for (int kind = 0; kind < 4; ++kind) {
    std::vector<float> a, c;
    make_clouds(b, n, m, kind, a, c);
    CK(cudaMemcpy(d1, a.data(), p1 * 12, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d2, c.data(), p2 * 12, cudaMemcpyHostToDevice));
    CK(cudaMemset(stats, 0, 512));
    if (getenv("NT_MODE")) {
      unsigned int md = (unsigned int)atoi(getenv("NT_MODE"));
      CK(cudaMemcpy(stats + 127, &md, 4, cudaMemcpyHostToDevice));
    }
    CK(cudaMemset(dist[0], 0xff, p1 * 4));
    CK(cudaMemset(dist[1], 0xff, p2 * 4));
#ifdef NT_DEBUG_SCORES
    int rc = pcc::nn_tc_forward(b, n, d1, m, d2, dist[0], idx[0], dist[1], idx[1], stats, 0, dbg);
#else
    int rc = pcc::nn_tc_forward(b, n, d1, m, d2, dist[0], idx[0], dist[1], idx[1], stats, 0);
#endif
    if (rc != 0) {
      printf("nn_tc_forward rc=%d (%s)\n", rc, cudaGetErrorString((cudaError_t)rc));
      return 3;
    }
    brute_kernel<<<dim3((n + 127) / 128, b), 128>>>(n, d1, m, d2, bdist[0], bidx[0]);
    brute_kernel<<<dim3((m + 127) / 128, b), 128>>>(m, d2, n, d1, bdist[1], bidx[1]);
    CK(cudaDeviceSynchronize());
    unsigned int hs[128];
    CK(cudaMemcpy(hs, stats, 512, cudaMemcpyDeviceToHost));
#ifdef NT_DEBUG_STAMPS
    printf("  cycles in CTA(1,3,0): key loads issued %u, all loads issued %u, first mma committed %u, all mma %u, first tile ready %u, main loops done %u, written %u\n", hs[2], hs[3], hs[4], hs[5], hs[9], hs[6], hs[8]);
    for (int k = 0; k < 8; ++k) printf("    tile %2d: mma issued %u | epilogue: data landed %u, next tile ready %u, minima done %u\n", 8 + k, hs[32 + k], hs[48 + k], hs[64 + k], hs[80 + k]);
#endif
    int bad = 0, shown = 0;
    for (int s = 0; s < 2; ++s) {
      std::vector<float> hd(cnt[s]), hb(cnt[s]);
      std::vector<int> hi(cnt[s]), hbi(cnt[s]);
      CK(cudaMemcpy(hd.data(), dist[s], cnt[s] * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hb.data(), bdist[s], cnt[s] * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hi.data(), idx[s], cnt[s] * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hbi.data(), bidx[s], cnt[s] * 4, cudaMemcpyDeviceToHost));
      for (size_t t = 0; t < cnt[s]; ++t) {
        const bool ne = (hi[t] != hbi[t]) || (memcmp(&hd[t], &hb[t], 4) != 0);
        bad += ne;
        if (ne && shown < 5) {
          printf("  mismatch dir %d t=%zu got (%g,%d) want (%g,%d)\n", s, t, hd[t], hi[t], hb[t], hbi[t]);
          ++shown;
        }
      }
    }
    // score error of cloud 0, direction 0 against the exact distance in double, relative to |a|^2 + |b|^2 (centred)
    // score error of cloud 0, direction 0 against the exact distance in double, as a fraction of the bound the kernel
    // assumes: NT_REL d + NT_CEPS (|a|^2 + |b|^2), norms about the kernel's centre (mean of 32 sample points of cloud 2)
    double worst = 0.0, worst_abs = 0.0;
#ifdef NT_DEBUG_SCORES
    {
      std::vector<float> hs2((size_t)n * npad2);
      CK(cudaMemcpy(hs2.data(), dbg, hs2.size() * 4, cudaMemcpyDeviceToHost));
      double ctr[3] = {0, 0, 0};
      for (int t = 0; t < 32; ++t)
        for (int d = 0; d < 3; ++d) ctr[d] += c[(size_t)((long long)t * m / 32) * 3 + d] / 32.0;
      int ninf = 0;
      for (int i = 0; i < n; i += 7)
        for (int j = 0; j < m; ++j) {
          double dd = 0, na = 0, nb = 0;
          for (int d = 0; d < 3; ++d) {
            const double u = a[i * 3 + d], v = c[j * 3 + d];
            dd += (u - v) * (u - v);
            na += (u - ctr[d]) * (u - ctr[d]);
            nb += (v - ctr[d]) * (v - ctr[d]);
          }
          const float sc = hs2[(size_t)i * npad2 + j];
          if (std::isinf(sc)) {
            ++ninf;
            continue;
          }
          const double err = fabs((double)sc - dd);
          worst_abs = std::max(worst_abs, err);
          const double bound = (double)pcc::NT_REL * dd + (double)pcc::NT_CEPS * (na + nb);
          if (bound > 0) worst = std::max(worst, err / bound);
        }
      printf("  (%d sampled scores beyond the fp16 range)\n", ninf);
    }
#endif
    for (int w = 0; w < 3; ++w) pcc::nn_tc_forward(b, n, d1, m, d2, dist[0], idx[0], dist[1], idx[1], nullptr, 0
#ifdef NT_DEBUG_SCORES
                                                    ,
                                                    nullptr
#endif
    );
    CK(cudaEventRecord(e0));
    for (int r = 0; r < 20; ++r) pcc::nn_tc_forward(b, n, d1, m, d2, dist[0], idx[0], dist[1], idx[1], nullptr, 0
#ifdef NT_DEBUG_SCORES
                                                    ,
                                                    nullptr
#endif
    );
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-9s b=%d n=%d m=%d mismatches=%d of %zu  exact rescans=%u  resolved chunks/query=%.2f  score err / assumed bound: max %.3g "
           "(NT_CEPS %.3g) abs %.3g  forward (prep+search): %.1f us\n",
           names[kind], b, n, m, bad, p1 + p2, hs[0], hs[1] / (double)(p1 + p2), worst, (double)pcc::NT_CEPS, worst_abs,
           ms / 20 * 1e3);
    bad_total += bad;
  }
  return bad_total ? 1 : 0;
}
