// Feature-space kNN graph (DGCNN EdgeConv, C = 32..256 channels) on the 5th-generation tensor cores of sm_100a.
//
// Replaces the KeOps argKmin behind src/utils/neighbour_ops.py:77-82 for high-dimensional features.
//   prep      x (B,C,N) channels-first -> xT (B,N,C) point-major (K-major rows for the MMA and contiguous rows for the
//             exact re-rank), squared norms, per-cloud max norm.
//   main      per CTA: 128 query points of one cloud against all N references, two passes over the reference tiles.
//             warp 0   TMA producer: 128B-swizzled K-major tiles of xT (cp.async.bulk.tensor.3d, mbarrier complete_tx)
//             warp 1   MMA issuer: tcgen05.mma kind::tf32, M=128 x N=256 x K=8, accumulators double-buffered in TMEM
//             warps 4-7  epilogue, one thread per query: tcgen05.ld of its accumulator row, score = |x_j|^2 - 2 x_i.x_j
//                pass 1  minima over groups of 16 references -> k-th smallest group minimum T (branch-free sorted list)
//                pass 2  references with score <= T + 2*eps are candidates (eps bounds |TF32 score - exact distance|)
//                then    EXACT fp32 distances of the ~k+8 candidates in the canonical order (sequential fma over
//                        channels), stable sort by (distance, index) -> the indices are bit-identical to the exact
//                        SIMT kernel / the oracle; the tensor cores only generate candidates.
// The candidate set provably contains the exact top-k: at least k references have score <= T, their exact distances are
// <= T + eps, so the exact k-th distance is <= T + eps and every exact top-k reference has score <= T + 2 eps.
#include "tc_ptx.cuh"

namespace pcc {

constexpr int TC_THREADS = 256;
constexpr int TC_M = 128;       // queries per CTA (UMMA M)
constexpr int TC_N = 128;       // references per accumulator tile (UMMA N)
constexpr int TC_KB = 32;       // fp32 channels per K-block = one 128-byte swizzle row
constexpr int TC_STAGES = 2;
constexpr int TC_CAP = 56;      // candidate slots per query (uint16 indices + fp32 distances: 42 KiB per CTA)
constexpr int TC_A_BYTES = TC_M * TC_KB * 4;  // 16 KiB
constexpr int TC_B_BYTES = TC_N * TC_KB * 4;  // 16 KiB
constexpr int TC_TMEM_COLS = 2 * TC_N;        // two accumulators; 256 columns so that two CTAs fit one SM
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;

constexpr uint32_t TC_IDESC = umma_idesc_tf32(TC_M, TC_N);

// ---- prep: transpose + norms ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
knn_tc_prep_kernel(int c, int n, bool pm, const float *__restrict__ x, float *__restrict__ xT, float *__restrict__ norms,
                   int norm_stride, unsigned int *__restrict__ nmax_bits) {
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
    if (tx == 0 && p < n) {
      norms[cloud * (size_t)norm_stride + p] = s;
      atomicMax(&nmax_bits[cloud], __float_as_uint(s));  // norms are >= 0: unsigned order == float order
    } else if (tx == 0 && p < norm_stride) {
      norms[cloud * (size_t)norm_stride + p] = __int_as_float(0x7f800000);  // padding: never a neighbour
    }
  }
}

// ---- main kernel ---------------------------------------------------------------------------------------------
struct TcSmemCtl {
  uint64_t full[TC_STAGES], empty[TC_STAGES], tfull[2], tempty[2];
  uint32_t tmem_base;
  int qcnt[TC_M];  // candidates per query after the two passes (-1: overflow => exact brute force)
};

template <int K>
__global__ void __launch_bounds__(TC_THREADS, 2)
knn_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_r, int c, int n, int k,
              const float *__restrict__ xT, const float *__restrict__ norms, const unsigned int *__restrict__ nmax_bits,
              int64_t *__restrict__ idx_out, float *__restrict__ dist_out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *stage_base = smem;                                               // TC_STAGES * 48 KiB, 1024-aligned
  float *candd = reinterpret_cast<float *>(smem + TC_STAGES * TC_STAGE_BYTES);    // [TC_CAP][128]
  unsigned short *cand = reinterpret_cast<unsigned short *>(candd + TC_CAP * TC_M);  // [TC_CAP][128], n <= 65535
  float *rn = reinterpret_cast<float *>(cand + TC_CAP * TC_M);                    // [2][TC_N] n_j (1 +- c2), [2][TC_N] +-c1 sqrt(n_j)
  float *rs = rn + 2 * TC_N;
  TcSmemCtl *ctl = reinterpret_cast<TcSmemCtl *>(rn + 4 * TC_N);
  int *qcnt = ctl->qcnt;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cloud = blockIdx.y;
  const int q0 = blockIdx.x * TC_M;
  const int nkb = c / TC_KB;                       // K-blocks per tile
  const int ntile = (n + TC_N - 1) / TC_N;         // reference tiles per pass
  const int niter = 2 * ntile;                     // two passes

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&ctl->tfull[a], 1);
      mbar_init(&ctl->tempty[a], 4);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&ctl->tmem_base, TC_TMEM_COLS);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int it = 0; it < niter; ++it) {
        const int r0 = (it % ntile) * TC_N;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&ctl->empty[stage], phase ^ 1);
          unsigned char *sa = stage_base + stage * TC_STAGE_BYTES;
          mbar_expect_tx(&ctl->full[stage], TC_STAGE_BYTES);
          tma_load_3d(sa, &tmap_q, &ctl->full[stage], kb * TC_KB, q0, cloud);
          tma_load_3d(sa + TC_A_BYTES, &tmap_r, &ctl->full[stage], kb * TC_KB, r0, cloud);
          if (++stage == TC_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int it = 0; it < niter; ++it) {
        const int a = it & 1;
        mbar_wait(&ctl->tempty[a], ((it >> 1) & 1) ^ 1);  // epilogue drained this accumulator
        fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(a * TC_N);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&ctl->full[stage], phase);
          fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * TC_STAGE_BYTES);
          const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + TC_A_BYTES);
#pragma unroll
          for (int kk = 0; kk < TC_KB / 8; ++kk)  // UMMA_K = 8 tf32 = 32 bytes: advance the start address by 2 (x16 B)
            mma_tf32(tmem_d, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), TC_IDESC, (kb | kk) ? 1u : 0u);
          mma_commit(&ctl->empty[stage]);  // frees the smem stage when these MMAs retire
          if (++stage == TC_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        mma_commit(&ctl->tfull[a]);  // accumulator complete
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread e owns query q0 + e =====
    const int e = threadIdx.x - 128;
    const int quarter = warp & 3;  // TMEM lanes 32*quarter .. +31 are accessible to this warp
    const int q = q0 + e;
    const float INF = __int_as_float(0x7f800000);
    const float *nb = norms + (size_t)cloud * n;
    const float nq = (q < n) ? nb[q] : 0.f;
    // |score + |x_q|^2 - exact| <= eps_ij = c1 |x_q| |x_j| + c2 (|x_q|^2 + |x_j|^2): TF32 truncation of both operands (2^-9
    // relative on every product), x2 for the -2 x.y term, Cauchy-Schwarz; c1 = 2^-7.5 leaves 41 % slack, the c2 terms
    // cover fp32 rounding of norms / distances.  The bound is PER KEY (see knn_tc2.cu): pass 0 ranks upper bounds, pass
    // 1 tests lower bounds; the per-query constant c2 |x_q|^2 moves into the threshold.
    constexpr float C1 = 0.0055242717f, C2 = 4e-5f;
    const float cq = sqrtf(nq), c2nq = C2 * nq;
    float L[K];
#pragma unroll
    for (int i = 0; i < K; ++i) L[i] = INF;
    float thr = INF;
    int cnt = 0;

    for (int it = 0; it < niter; ++it) {
      const int a = it & 1;
      const int pass = it / ntile;
      const int r0 = (it % ntile) * TC_N;
      // norms of this reference tile (+inf beyond the cloud => never selected)
      asm volatile("bar.sync 1, 128;" ::: "memory");  // previous user of rn[a] is done (two iterations back)
      for (int j = e; j < TC_N; j += 128) {
        const bool real = r0 + j < n;
        const float nj = real ? nb[r0 + j] : INF;
        rn[a * TC_N + j] = real ? nj * (pass == 0 ? 1.f + C2 : 1.f - C2) : INF;
        rs[a * TC_N + j] = real ? (pass == 0 ? C1 : -C1) * sqrtf(nj) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(&ctl->tfull[a], (it >> 1) & 1);
      fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * TC_N);
#pragma unroll 1
      for (int ch = 0; ch < TC_N / 32; ++ch) {
        float v[32];
        tmem_ld32(taddr + (uint32_t)(ch * 32), v);
        const float4 *rn4 = reinterpret_cast<const float4 *>(rn + a * TC_N + ch * 32);
        const float4 *rs4 = reinterpret_cast<const float4 *>(rs + a * TC_N + ch * 32);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 w = rn4[g], sn = rs4[g];
          v[4 * g + 0] = fmaf(cq, sn.x, fmaf(-2.f, v[4 * g + 0], w.x));
          v[4 * g + 1] = fmaf(cq, sn.y, fmaf(-2.f, v[4 * g + 1], w.y));
          v[4 * g + 2] = fmaf(cq, sn.z, fmaf(-2.f, v[4 * g + 2], w.z));
          v[4 * g + 3] = fmaf(cq, sn.w, fmaf(-2.f, v[4 * g + 3], w.w));
        }
        if (pass == 0) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {  // two groups of 16 references
            float m = fminf(v[16 * h], v[16 * h + 1]);
#pragma unroll
            for (int i = 2; i < 16; i += 2) m = fminf(fminf(v[16 * h + i], v[16 * h + i + 1]), m);
            float cc = m;
#pragma unroll
            for (int i = 0; i < K; ++i) {
              const float lo = fminf(L[i], cc);
              cc = fmaxf(L[i], cc);
              L[i] = lo;
            }
          }
        } else {
          const int jb = r0 + ch * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) {  // branch-free append: always store to the next slot, advance it on a hit
            cand[min(cnt, TC_CAP - 1) * TC_M + e] = (unsigned short)(jb + i);
            cnt += (v[i] <= thr) ? 1 : 0;
          }
        }
      }
      // release the accumulator to the MMA warp
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->tempty[a]);
      if (pass == 0 && it == ntile - 1) {
        float t = -INF;
#pragma unroll
        for (int i = 0; i < K; ++i) t = (i < k) ? fmaxf(t, L[i]) : t;  // k-th smallest group minimum
        thr = t + 2.f * c2nq;
        thr += 1e-6f * fabsf(thr) + 1e-30f;  // a few ulps up: equality stays a candidate
      }
    }

    qcnt[e] = (cnt > TC_CAP - 1) ? -1 : cnt;  // the last slot is scratch for the branch-free append
  }

  fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TC_TMEM_COLS);

  // ---- exact re-rank in the canonical arithmetic (sequential fma over channels) ----------------------------------
  // Rows are staged through shared memory with coalesced loads (the pipeline stages are free now): per step the 256
  // threads load one 32-channel chunk of the 128 query rows and of two candidate rows per query (rounds 2rp, 2rp+1),
  // then thread (e, half) continues the fma chain of query e / candidate 2rp+half from its padded rows.
  {
    constexpr int RS = 36;  // padded row stride in floats: 16-byte aligned, 4 wavefronts per warp-wide LDS.128
    float *qtile = reinterpret_cast<float *>(stage_base);  // [128][RS]
    float *ctile = qtile + TC_M * RS;                       // [2][128][RS]
    const int e = threadIdx.x & 127, half = threadIdx.x >> 7;
    const int mycnt = max(qcnt[e], 0);
    __shared__ int s_max;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    if (half == 0 && q0 + e < n) atomicMax(&s_max, mycnt);
    __syncthreads();
    const int rounds = (s_max + 1) / 2;
    const float4 *xT4 = reinterpret_cast<const float4 *>(xT) + (size_t)cloud * n * (c / 4);
    const int c4n = c / 4;
    for (int rp = 0; rp < rounds; ++rp) {
      const int r = 2 * rp + half;
      float d = 0.f;
      for (int c8 = 0; c8 < c / 32; ++c8) {
        __syncthreads();
        // 3 * 128 rows * 8 float4 = 12 per thread: all loads are issued before the first store (memory-level parallelism)
        float4 val[12];
#pragma unroll
        for (int it2 = 0; it2 < 12; ++it2) {
          const int t = threadIdx.x + it2 * TC_THREADS;
          const int which = t >> 10, row = (t >> 3) & 127, f4 = t & 7;
          int src = -1;
          if (which == 0) {
            src = (q0 + row < n) ? q0 + row : -1;
          } else {
            const int rr = 2 * rp + which - 1;
            if (rr < qcnt[row]) src = cand[rr * TC_M + row];
          }
          val[it2] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (src >= 0) val[it2] = __ldg(xT4 + (size_t)src * c4n + c8 * 8 + f4);
        }
#pragma unroll
        for (int it2 = 0; it2 < 12; ++it2) {
          const int t = threadIdx.x + it2 * TC_THREADS;
          const int which = t >> 10, row = (t >> 3) & 127, f4 = t & 7;
          float *dst = (which == 0 ? qtile : ctile + (which - 1) * TC_M * RS) + row * RS + f4 * 4;
          *reinterpret_cast<float4 *>(dst) = val[it2];
        }
        __syncthreads();
        if (r < mycnt) {
          const float4 *aq = reinterpret_cast<const float4 *>(qtile + e * RS);
          const float4 *ar = reinterpret_cast<const float4 *>(ctile + half * TC_M * RS + e * RS);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float4 a = aq[u], b = ar[u];
            const float t0 = a.x - b.x, t1 = a.y - b.y, t2 = a.z - b.z, t3 = a.w - b.w;
            d = fmaf(t0, t0, d);
            d = fmaf(t1, t1, d);
            d = fmaf(t2, t2, d);
            d = fmaf(t3, t3, d);
          }
        }
      }
      if (r < mycnt) candd[r * TC_M + e] = d;
    }
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int e = threadIdx.x;
    const int q = q0 + e;
    const float INF = __int_as_float(0x7f800000);
    int cnt = qcnt[e];
    if (q < n) {
      int64_t *o = idx_out + ((size_t)cloud * n + q) * k;
      float *od = dist_out ? dist_out + ((size_t)cloud * n + q) * k : nullptr;
      if (cnt < 0) {
        // pathological ties (e.g. duplicated clouds): exact brute force over all references, sorted insertion
        const float4 *xq = reinterpret_cast<const float4 *>(xT + ((size_t)cloud * n + q) * c);
        cnt = 0;
        for (int j = 0; j < n; ++j) {
          const float4 *xr = reinterpret_cast<const float4 *>(xT + ((size_t)cloud * n + j) * c);
          float d = 0.f;
          for (int c4 = 0; c4 < c / 4; ++c4) {
            const float4 aq = xq[c4], ar = xr[c4];
            float t0 = aq.x - ar.x, t1 = aq.y - ar.y, t2 = aq.z - ar.z, t3 = aq.w - ar.w;
            d = fmaf(t0, t0, d);
            d = fmaf(t1, t1, d);
            d = fmaf(t2, t2, d);
            d = fmaf(t3, t3, d);
          }
          if (cnt == k && !(d < candd[(k - 1) * TC_M + e])) continue;
          int p = cnt < k ? cnt : k - 1;
          while (p > 0 && d < candd[(p - 1) * TC_M + e]) {
            candd[p * TC_M + e] = candd[(p - 1) * TC_M + e];
            cand[p * TC_M + e] = cand[(p - 1) * TC_M + e];
            --p;
          }
          candd[p * TC_M + e] = d;
          cand[p * TC_M + e] = (unsigned short)j;
          if (cnt < k) ++cnt;
        }
        for (int t = 0; t < k; ++t) {
          o[t] = t < cnt ? (int64_t)cand[t * TC_M + e] : 0;
          if (od) od[t] = t < cnt ? candd[t * TC_M + e] : INF;
        }
      } else {
        // rank sort: position of candidate s among (distance, append order); candidates were appended in ascending
        // index order, so this is the (distance, index) order.  No data-dependent loop, all loads independent.
        for (int t = cnt; t < k; ++t) {  // fewer than k candidates only with NaN / inf inputs
          o[t] = 0;
          if (od) od[t] = INF;
        }
        for (int s1 = 0; s1 < cnt; ++s1) {
          const float ds = candd[s1 * TC_M + e];
          int rank = 0;
#pragma unroll 8
          for (int t = 0; t < cnt; ++t) {
            const float dt = candd[t * TC_M + e];
            rank += (dt < ds || (dt == ds && t < s1)) ? 1 : 0;
          }
          if (rank < k) {
            o[rank] = (int64_t)cand[s1 * TC_M + e];
            if (od) od[rank] = ds;
          }
        }
      }
    }
  }
}

// ---- host ----------------------------------------------------------------------------------------------------
PFN_tmapEncodeTiled tc_get_encode() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_tmapEncodeTiled)p;
  }
  return fn;
}

int tc_make_map(CUtensorMap *m, const float *xT, int b, int n, int c, int box_rows) {
  PFN_tmapEncodeTiled enc = tc_get_encode();
  if (!enc) return PCC_ENOTSUP;
  cuuint64_t gdim[3] = {(cuuint64_t)c, (cuuint64_t)n, (cuuint64_t)b};
  cuuint64_t gstride[2] = {(cuuint64_t)c * 4, (cuuint64_t)n * c * 4};
  cuuint32_t box[3] = {32u, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)xT, gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : PCC_ENOTSUP;
}

template <int K>
static int launch_tc_k(const CUtensorMap &mq, const CUtensorMap &mr, int b, int c, int n, int k, const float *xT,
                       const float *norms, const unsigned int *nmax, int64_t *idx, float *dist, cudaStream_t st) {
  const size_t smem = (size_t)TC_STAGES * TC_STAGE_BYTES + (size_t)TC_CAP * TC_M * 6 + 4 * TC_N * 4 + sizeof(TcSmemCtl) + 64;
  static size_t attr[64];
  if (cudaError_t e = smem_optin(knn_tc_kernel<K>, smem, attr); e != cudaSuccess) return (int)e;
  dim3 grid((n + TC_M - 1) / TC_M, b);
  knn_tc_kernel<K><<<grid, TC_THREADS, smem, st>>>(mq, mr, c, n, k, xT, norms, nmax, idx, dist);
  return (int)cudaGetLastError();
}

// x (b,c,n) channels-first, or (b,n,c) point-major with pm.  Returns PCC_ENOTSUP when the shape is outside this path
// (caller falls back to SIMT).
int knn_tc_launch(int b, int c, int n, int k, bool pm, const float *x, int64_t *idx, float *dist, cudaStream_t st) {
  if (c % TC_KB != 0 || c < TC_KB || c > 1024 || k > 32 || k > TC_CAP / 2 || n < 16 * k || n > 65535 || b > 65535)
    return PCC_ENOTSUP;
  if (!tc_get_encode()) return PCC_ENOTSUP;
  float *ws = nullptr;
  const size_t nxt = (size_t)b * n * c, nn = (size_t)b * n;
  cudaError_t e = ws_alloc((void **)&ws, sizeof(float) * (nxt + nn) + sizeof(unsigned int) * b, st);
  if (e != cudaSuccess) return (int)e;
  const float *xT = pm ? x : ws;
  float *norms = ws + nxt;
  unsigned int *nmax = reinterpret_cast<unsigned int *>(norms + nn);
  cudaMemsetAsync(nmax, 0, sizeof(unsigned int) * b, st);
  knn_tc_prep_kernel<<<dim3((n + 31) / 32, b), 256, 0, st>>>(c, n, pm, x, ws, norms, n, nmax);
  CUtensorMap mq, mr;
  int rc = tc_make_map(&mq, xT, b, n, c, TC_M);
  if (rc == 0) rc = tc_make_map(&mr, xT, b, n, c, TC_N);
  if (rc == 0) {
    if (k <= 8) rc = launch_tc_k<8>(mq, mr, b, c, n, k, xT, norms, nmax, idx, dist, st);
    else if (k <= 16) rc = launch_tc_k<16>(mq, mr, b, c, n, k, xT, norms, nmax, idx, dist, st);
    else if (k <= 20) rc = launch_tc_k<20>(mq, mr, b, c, n, k, xT, norms, nmax, idx, dist, st);
    else if (k <= 24) rc = launch_tc_k<24>(mq, mr, b, c, n, k, xT, norms, nmax, idx, dist, st);
    else rc = launch_tc_k<32>(mq, mr, b, c, n, k, xT, norms, nmax, idx, dist, st);
  }
  cudaFreeAsync(ws, st);
  if (rc == 0) g_launches.fetch_add(2, std::memory_order_relaxed);
  return rc;
}

}  // namespace pcc
