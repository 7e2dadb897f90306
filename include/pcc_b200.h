/*
 * pcc_b200.h -- C ABI of libpcc_b200.so, the B200 (sm_100a) geometry hot path for
 * nverchev/PointCloudCounterfactual.
 *
 * Drop-in boundary: every entry point below replaces one native launcher that the reference's
 * pybind11 bindings call (paths relative to /root/reference).  Same argument lists and meaning;
 * differences from the reference launchers:
 *   - `extern "C"`, returns an int status (0 = ok, otherwise a cudaError_t value or a PCC_E* code)
 *     instead of void / throwing std::runtime_error (approxmatch.cu:303-306) / printf (emd_cuda.cu:236-248);
 *   - all pointers are DEVICE pointers on the current device unless the name ends in `_host`;
 *   - all work is enqueued on `stream` (the reference's emd kernels ignore the current stream,
 *     emd_cuda.cu:256-268, and nndistancegrad memsets on the legacy stream, nndistance.cu:150-151);
 *     no entry point synchronises the host;
 *   - no torch / ATen types anywhere in the signatures.
 */
#ifndef PCC_B200_H_
#define PCC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *pcc_stream_t; /* a cudaStream_t */

#define PCC_OK 0
#define PCC_EINVAL (-1)    /* shape rule violated (mirrors emd_cuda_forward's -1, emd_cuda.cu:235-248) */
#define PCC_ENOTSUP (-2)   /* configuration outside what the kernels support (e.g. k > PCC_KNN_MAX_K) */
#define PCC_ELAUNCH (-3)   /* pcc_emd_forward / _backward only (they return 1 for success like emd_cuda_forward): the launch
                              failed with cudaErrorInvalidValue, whose numeric value is 1 as well */
#define PCC_KNN_MAX_K 128

/* Library / build identification. */
const char *pcc_version(void);
/* Text for a status returned by any entry point (cudaGetErrorString for CUDA codes). */
const char *pcc_status_string(int status);
/* Number of kernels this library has launched since it was loaded (all entry points; for bench.py's gpu_launches). */
uint64_t pcc_launch_count(void);
/* Dispatch evidence: how often an entry point chose a kernel family since the library was loaded.  `name` is one of
 * pcc_route_names() (comma separated: knn3w, knn3_thread, knn_tc2, knn_tc1, knn_simt, argmin_small, nn_sym, nn_asym,
 * nn_tc, knn3_tc, pm_self, knn_bf); -1 for an unknown name.  tests/ and bench.py use the deltas to prove that calls made
 * through the reference's unchanged call sites (KeOps LazyTensor shim) reach the fast kernels. */
const char *pcc_route_names(void);
int64_t pcc_route_count(const char *name);

/* ---- Chamfer nearest-neighbour distance -------------------------------------------------------------
 * Replaces `void nndistance(int b,int n,const float*xyz,int m,const float*xyz2,float*result,int*result_i,
 *                           float*result2,int*result2_i,cudaStream_t)`
 *   declared external/pytorch_structural_losses/src/structural_loss.cpp:13, defined src/nndistance.cu:125-128.
 * xyz (b,n,3), xyz2 (b,m,3) fp32 contiguous.  result[b,n] = min_k |xyz[j]-xyz2[k]|^2 with the arithmetic
 * d = fma(dz,dz,fma(dx,dx,dy*dy)); result_i[b,n] = LOWEST k attaining it (int32).  result2/result2_i: roles swapped.
 * One fused launch covers both directions. */
int pcc_nndistance(int b, int n, const float *xyz, int m, const float *xyz2, float *result, int *result_i,
                   float *result2, int *result2_i, pcc_stream_t stream);

/* The same search on the tcgen05 tensor cores (csrc/chamfer_tc.cu): fp16 operand rows built from the clouds, one
 * tcgen05.mma (K = 16) per 128 x 256 tile of scores, packed 16-bit read-out of the accumulators, and exact resolution of
 * the candidate keys with the arithmetic above -- results are BIT-IDENTICAL to pcc_nndistance.  PCC_ENOTSUP unless
 * 256 <= n, m <= 2560.  Not the default: reading the scores out of tensor memory (128 B per clock and SM) bounds it at
 * about the speed of the SIMT kernel for 3-dimensional points. */
int pcc_nndistance_tc(int b, int n, const float *xyz, int m, const float *xyz2, float *result, int *result_i,
                      float *result2, int *result2_i, pcc_stream_t stream);

/* Replaces `void nndistancegrad(...)` (structural_loss.cpp:14, nndistance.cu:149-154).
 * grad_xyz1 (b,n,3) / grad_xyz2 (b,m,3) are fully written (no memset needed).  Deterministic: per target point
 * the own term is added first, then the scatter terms in ascending source index (the reference uses float
 * atomics in arbitrary order). */
int pcc_nndistancegrad(int b, int n, const float *xyz1, int m, const float *xyz2, const float *grad_dist1,
                       const int *idx1, const float *grad_dist2, const int *idx2, float *grad_xyz1,
                       float *grad_xyz2, pcc_stream_t stream);

/* Fused Chamfer loss behind `pykeops_chamfer` / `torch_chamfer` (src/train/metrics_and_losses.py:21-47):
 * loss[b] = scale1 * sum_j dist1[b,j] + scale2 * sum_k dist2[b,k]  (scale = 1/points for the mean form, 1 for the sum
 * form), with dist/idx exactly as pcc_nndistance writes them (kept for the backward).  Deterministic reduction. */
int pcc_chamfer_reduce(int b, int n, const float *xyz1, int m, const float *xyz2, float scale1, float scale2,
                       float *loss, float *dist1, int *idx1, float *dist2, int *idx2, pcc_stream_t stream);
/* Its backward: pcc_nndistancegrad with the upstream gradient grad_loss[b] * scale1 (scale2) for every point of cloud
 * 1 (2), without materialising the per-point gradient arrays.  PCC_ENOTSUP for clouds above 16k points. */
int pcc_chamfer_reduce_grad(int b, int n, const float *xyz1, int m, const float *xyz2, const int *idx1,
                            const int *idx2, const float *grad_loss, float scale1, float scale2, float *grad_xyz1,
                            float *grad_xyz2, pcc_stream_t stream);

/* ---- Approximate-matching EMD -----------------------------------------------------------------------
 * Replaces `void approxmatch(int b,int n,int m,const float*xyz1,const float*xyz2,float*match,float*temp,
 *                            cudaStream_t)` (structural_loss.cpp:10, approxmatch.cu:299-307).
 * match (b,m,n) is fully written; temp (b, 2(n+m)) is scratch = [remainL|remainR|ratioL|ratioR] per cloud. */
int pcc_approxmatch(int b, int n, int m, const float *xyz1, const float *xyz2, float *match, float *temp,
                    pcc_stream_t stream);
/* Replaces `void matchcost(...)` (structural_loss.cpp:11, approxmatch.cu:309-316): out[b] = sum match*|x1-x2|. */
int pcc_matchcost(int b, int n, int m, const float *xyz1, const float *xyz2, const float *match, float *out,
                  pcc_stream_t stream);
/* Replaces `void matchcostgrad(...)` (structural_loss.cpp:12, approxmatch.cu:318-326). */
int pcc_matchcostgrad(int b, int n, int m, const float *xyz1, const float *xyz2, const float *match,
                      float *grad1, float *grad2, pcc_stream_t stream);
/* Fused replacement for the ApproxMatch -> MatchCost -> MatchCostGrad chain behind
 * structural_losses.match_cost (structural_losses/match_cost.py:14-42): same cost and gradients, but the
 * (b,m,n) match matrix is never materialised (O(b(n+m)) memory).  grad1 / grad2 may be NULL.
 * temp: (b, 2(n+m)) floats of scratch; its contents are unspecified on return (the solver skips the last remainL
 * update and the ratio export, which only pcc_approxmatch's temp receives). */
int pcc_matchcost_fused(int b, int n, int m, const float *xyz1, const float *xyz2, float *cost, float *grad1,
                        float *grad2, float *temp, pcc_stream_t stream);

/* Measurement hook (bench.py's roofline leg): ONE sweep of the approxmatch solver -- the kernel that dominates the
 * EMD path -- exactly as pcc_approxmatch / pcc_matchcost_fused launch it for sweep 1 of a level
 * (approxmatch.cu:29-62): ratio[b,n] = remain[b,n] / (1e-9 + sum_l exp(level*|x1_k-x2_l|^2) * weight[b,l]). */
int pcc_approxmatch_sweep(int b, int n, int m, const float *xyz1, const float *xyz2, const float *weight,
                          const float *remain, float *ratio, float level, int points_per_thread /* 0 = default */,
                          pcc_stream_t stream);

/* ---- kNN graph (DGCNN) ------------------------------------------------------------------------------
 * Replaces the KeOps reduction behind `pykeops_knn` / `knn` (src/utils/neighbour_ops.py:63-82):
 * x (b,c,n) channels-first fp32 -> idx (b,n,k) int64, the k smallest squared distances per point INCLUDING the
 * point itself, ascending by (distance, index); distance = sum_c (x_i-x_j)^2 accumulated with fma over c in order.
 * dist (b,n,k) may be NULL.  Requires 1 <= k <= min(n, PCC_KNN_MAX_K). */
int pcc_knn(int b, int c, int n, int k, const float *x, int64_t *idx, float *dist, pcc_stream_t stream);

/* General argKmin between two point sets, point-major layout (the LazyTensor patterns of
 * neighbour_ops.py:35-40,81; metrics_and_losses.py:33,36; quantize.py:28):
 * q (b,nq,c), r (b,nr,c) -> idx (b,nq,k) int64 ascending by (distance, index).  k=1 is argmin. */
int pcc_argkmin(int b, int nq, int nr, int c, int k, const float *q, const float *r, int64_t *idx, float *dist,
                pcc_stream_t stream);

/* ---- EdgeConv front-end ----------------------------------------------------------------
 * Replaces the torch composition of get_neighbours (src/utils/neighbour_ops.py:85-94: expand + torch.gather) and
 * get_graph_features (:113-119: gather, expand, subtract, cat, contiguous) with one pass that writes the result once.
 * x (b,c,n) channels-first fp32, idx (b,n,k) int64 (as produced by pcc_knn or passed in by the caller).
 *   mode 0: out (b,c,n,k)   out[c][i][t] = x[c][idx[i][t]]
 *   mode 1: out (b,2c,n,k)  out[c][i][t] = x[c][idx[i][t]] - x[c][i],  out[c+C][i][t] = x[c][i]
 * Indices outside [0,n) are clamped (torch.gather raises).  PCC_ENOTSUP for k > 32 or n > 8192 (caller composes). */
int pcc_graph_gather(int b, int c, int n, int k, const float *x, const int64_t *idx, int mode, float *out,
                     pcc_stream_t stream);
/* Its backward w.r.t. x: grad_x (b,c,n) is fully written.  While one (point, neighbour) plane fits shared memory (n*k floats,
 * n*k % 4 == 0, n <= 4096, k <= 32) the edges are sorted by target and summed in list order: no atomics, bitwise reproducible;
 * other shapes scatter-add with shared-memory float atomics (summation order not fixed, like the backward of torch.gather). */
int pcc_graph_gather_grad(int b, int c, int n, int k, const int64_t *idx, int mode, const float *grad_out,
                          float *grad_x, pcc_stream_t stream);
/* The same backward with the edge sort hoisted: pcc_graph_edge_sort_bytes = size of the buffer the sort fills (0: the sorted
 * backward does not cover this shape); pcc_graph_edge_sort depends on idx alone and may run early / on another stream;
 * pcc_graph_gather_grad_presorted consumes the buffer (same results as pcc_graph_gather_grad, bit for bit). */
long long pcc_graph_edge_sort_bytes(int b, int n, int k);
int pcc_graph_edge_sort(int b, int n, int k, const int64_t *idx, void *ws, pcc_stream_t stream);
int pcc_graph_gather_grad_presorted(int b, int c, int n, int k, int mode, const void *ws, const float *grad_out,
                                    float *grad_x, pcc_stream_t stream);

/* ---- Fused EdgeConv layer ----------------------------------------------------------------
 * Replaces get_graph_features (src/utils/neighbour_ops.py:113-119) -> EdgeConvLayer.forward (src/module/layers.py:
 * 159-203: Conv2d 1x1 without bias, BatchNorm2d, activation) -> max over the k neighbours (src/module/encoders.py:
 * 49-54) without the (b,2c,n,k) and (b,cout,n,k) tensors.  The caller supplies uv (b,n,2*cout) point-major,
 * uv[.., :cout] = W1 x, uv[.., cout:] = (W2 - W1) x with W = [W1 | W2] the (cout, 2c) convolution weight (one plain
 * GEMM), and idx (b,n,k) int64.  Edge value y(i,t) = u[idx[i][t]] + v[i].
 *   bn_mode 1: batch statistics over all b*n*k edges (training); running_mean/var (may be NULL) are updated with
 *              `momentum` (unbiased variance), mean/invstd (cout) are returned for the backward
 *   bn_mode 0: running statistics (eval);   bn_mode 2: no normalisation, out = act(gamma*y + beta) (gamma/beta may be NULL)
 *   act 0: identity, 1: leaky ReLU with `slope` >= 0 (0 = ReLU)
 * out (b,cout,n) channels-first.  exty, sy (b,n,cout) fp32 and slot (b,ceil(cout/8),n,8) uint8 (private layout) are saved
 * for the backward (sy only written in bn_mode 1).  PCC_ENOTSUP: k > 64, n > 8192, cout % 4 != 0 or cout > 1024. */
int pcc_edgeconv_forward(int b, int n, int k, int cout, const float *uv, const int64_t *idx, const float *gamma,
                         const float *beta, float *running_mean, float *running_var, int bn_mode, float momentum,
                         float eps, int act, float slope, float *out, float *exty, float *sy, unsigned char *slot,
                         float *mean, float *invstd, pcc_stream_t stream);
/* Backward: grad_out (b,cout,n) -> grad_uv (b,n,2*cout), grad_gamma, grad_beta (cout; may be NULL).  In bn_mode 1 every
 * edge receives a gradient through the batch statistics.  Deterministic: the edges are sorted by target once per call
 * and summed in a fixed order (no float atomics). */
int pcc_edgeconv_backward(int b, int n, int k, int cout, const float *uv, const int64_t *idx, const float *gamma,
                          const float *beta, const float *mean, const float *invstd, int bn_mode, int act, float slope,
                          const float *exty, const float *sy, const unsigned char *slot, const float *grad_out,
                          float *grad_uv, float *grad_gamma, float *grad_beta, pcc_stream_t stream);

/* Batched fp32 GEMM on the tcgen05 tensor cores with fp32-level accuracy (3xTF32: hi.hi + hi.lo + lo.hi, error 2^-21
 * relative) -- the point contraction of the fused EdgeConv layer (the Conv2d 1x1 of src/module/layers.py:159-203 applied
 * to the points) and its two backward products, which were cuBLAS calls.
 *   D[z*ksplit + p](i,j) = sum_{l in part p of [0,k)} A[z](i,l) * B[z](j,l),   i < m, j < n, z < batch, p < ksplit
 * (ksplit = 1: the plain product; ksplit > 1 splits a long reduction over more CTAs, the caller adds the ksplit slices)
 * element (i,l) of A[z] at A[z*sAb + i*sAm + l*sAk], (j,l) of B[z] at B[z*sBb + j*sBn + l*sBk], (i,j) of D[z] at
 * D[z*sDb + i*sDm + j*sDn] (z over batch*ksplit slices; strides in elements; any of the two dimensions of an operand may be the contiguous one, a
 * batch stride of 0 shares the operand).  D is fully written. */
int pcc_gemm_tf32x3(int batch, int ksplit, int m, int n, int k, const float *A, long long sAb, long long sAm, long long sAk,
                    const float *B, long long sBb, long long sBn, long long sBk, float *D, long long sDb, long long sDm,
                    long long sDn, pcc_stream_t stream);

/* Decoder-output smoothing `graph_filtering` (src/utils/neighbour_ops.py:122-133), one launch per
 * direction.  x (b,3,n), idx (b,n,k) int64 = the kNN list of x itself (column 0 is the point), 2 <= k <= 8, n <= 6144.
 * out (b,3,n); mean_dist (b) receives the per-cloud mean nearest-neighbour distance (sigma before the 0.005 clamp),
 * which the backward takes back.  Duplicate points contribute no distance gradient (torch yields NaN there). */
int pcc_graph_filtering(int b, int n, int k, const float *x, const int64_t *idx, float *out, float *mean_dist,
                        pcc_stream_t stream);
int pcc_graph_filtering_grad(int b, int n, int k, const float *x, const int64_t *idx, const float *mean_dist,
                             const float *grad_out, float *grad_x, pcc_stream_t stream);

/* ---- Auction EMD ------------------------------------------------------------------------------------
 * Replaces `int emd_cuda_forward(at::Tensor xyz1, ..., float eps, int iters)` (external/emd/src/emd.cpp:14-21,
 * emd_cuda.cu:227-281) with the tensors passed as raw pointers in the same order.  The caller allocates and
 * initialises the work buffers exactly as emd/emd_module.py:34-45 does (assignment = assignment_inv = -1, the rest 0;
 * unass_cnt / unass_cnt_sum / cnt_tmp are 512 ints).  Returns 1 on success and -1 on a shape error, like the
 * reference (n must be a multiple of 1024, b <= 512); other values are CUDA errors.
 * One persistent thread-block cluster (4 CTAs) per cloud runs all `iters` rounds in one launch (the reference launches 7
 * kernels per round); unass_cnt carries the per-CTA counts of the cluster (a pool allocation replaces it for b > 128). */
int pcc_emd_forward(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *assignment,
                    float *price, int *assignment_inv, int *bid, float *bid_increments, float *max_increments,
                    int *unass_idx, int *unass_cnt, int *unass_cnt_sum, int *cnt_tmp, int *max_idx, float eps,
                    int iters, pcc_stream_t stream);
/* Replaces `int emd_cuda_backward(xyz1, xyz2, gradxyz, graddist, idx)` (emd.cpp:23-26, emd_cuda.cu:301-315):
 * gradxyz (b,n,3) is overwritten with 2*graddist*(xyz1 - xyz2[idx]).  Returns 1 on success. */
int pcc_emd_backward(int b, int n, const float *xyz1, const float *xyz2, float *gradxyz, const float *graddist,
                     const int *idx, pcc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PCC_B200_H_ */
