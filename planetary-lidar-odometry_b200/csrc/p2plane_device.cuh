// p2plane_device.cuh — device functions of the point-to-plane solve shared by p2plane_solve.cu (stand-alone reduce /
// solve kernels) and knn_project.cu (the same sums and the same solve fused into the projection kernels' epilogue):
// row builder and weights (src/solver.cpp:185-198, :334-364), 6x6 LDL^T, Rodrigues + polar factor (:203-217), the
// loop tail of src/laser_odometry.cpp:619-646.  Internal linkage: included by several .cu files.
#pragma once

#include <float.h>
#include <math_constants.h>

#include "plo_internal.cuh"

namespace {

__device__ __forceinline__ void ab_row(const double s[3], const double d[3], const double n[3], double a[6], double& b) {
  // src/solver.cpp:185-192
  a[0] = __dsub_rn(__dmul_rn(n[2], s[1]), __dmul_rn(n[1], s[2]));
  a[1] = __dsub_rn(__dmul_rn(n[0], s[2]), __dmul_rn(n[2], s[0]));
  a[2] = __dsub_rn(__dmul_rn(n[1], s[0]), __dmul_rn(n[0], s[1]));
  a[3] = n[0]; a[4] = n[1]; a[5] = n[2];
  b = __dadd_rn(__dadd_rn(__dmul_rn(n[0], __dsub_rn(d[0], s[0])), __dmul_rn(n[1], __dsub_rn(d[1], s[1]))),
                __dmul_rn(n[2], __dsub_rn(d[2], s[2])));
}

__device__ __forceinline__ void apply_T3(const double* __restrict__ T, const double s[3], double out[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
    out[i] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[i * 4], s[0]), __dmul_rn(T[i * 4 + 1], s[1])), __dmul_rn(T[i * 4 + 2], s[2])), T[i * 4 + 3]);
}

// point-to-plane distance of a pair under hypothesis T (src/solver.cpp:306-307, :347-348)
__device__ __forceinline__ double plane_distance(const double* __restrict__ T, const double s[3], const double d[3], const double n[3]) {
  double tp[3];
  apply_T3(T, s, tp);
  return fabs(__dadd_rn(__dadd_rn(__dmul_rn(__dsub_rn(tp[0], d[0]), n[0]), __dmul_rn(__dsub_rn(tp[1], d[1]), n[1])),
                        __dmul_rn(__dsub_rn(tp[2], d[2]), n[2])));
}

// RANSAC-final weight of src/solver.cpp:334-364 evaluated at hypothesis T (T_best of the RANSAC front,
// or the identity for PLO_W_HUBER_EXP without RANSAC); < 0 => not an inlier
__device__ __forceinline__ double huber_exp_weight(const double* __restrict__ T, const double s[3], const double d[3],
                                                   const double n[3], const DevParams& P) {
  const double dist = plane_distance(T, s, d, n);
  if (!(dist < P.ransac_dist_thr)) return -1.0;
  const double ar = exp(-dist);
  const double sq = sqrt(ar);
  return sq < P.huber_thr2 ? ar : 2.0 * P.huber_thr2 * sq - P.huber_thr2 * P.huber_thr2;
}

__device__ __forceinline__ void accumulate_pair(double acc[PLO_NSUM], const double s[3], const double d[3], const double n[3],
                                                double w) {
  double a[6], b;
  ab_row(s, d, n, a, b);
  int t = 0;
#pragma unroll
  for (int p = 0; p < 6; ++p)
#pragma unroll
    for (int q = p; q < 6; ++q) acc[t++] += w * a[p] * a[q];
#pragma unroll
  for (int p = 0; p < 6; ++p) acc[21 + p] += w * a[p] * b;
  acc[27] += w;
  acc[28] += w * b * b;
}

// ---- 6x6 solve + pose update (one thread) -------------------------------------------

__device__ void rodrigues(const double r[3], double R[9]) {
  // Eigen AngleAxisd(rot.norm(), rot.normalized()).toRotationMatrix(); a zero vector stays zero
  const double z = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  const double angle = sqrt(z);
  double ax[3] = {r[0], r[1], r[2]};
  if (z > 0.0) { ax[0] /= angle; ax[1] /= angle; ax[2] /= angle; }
  const double sn = sin(angle), cs = cos(angle);
  const double sa[3] = {sn * ax[0], sn * ax[1], sn * ax[2]};
  const double ca[3] = {(1.0 - cs) * ax[0], (1.0 - cs) * ax[1], (1.0 - cs) * ax[2]};
  double tmp;
  tmp = ca[0] * ax[1]; R[1] = tmp - sa[2]; R[3] = tmp + sa[2];
  tmp = ca[0] * ax[2]; R[2] = tmp + sa[1]; R[6] = tmp - sa[1];
  tmp = ca[1] * ax[2]; R[5] = tmp - sa[0]; R[7] = tmp + sa[0];
  R[0] = ca[0] * ax[0] + cs; R[4] = ca[1] * ax[1] + cs; R[8] = ca[2] * ax[2] + cs;
}

// orthogonal polar factor of a near-rotation (== U V^T of its SVD, src/solver.cpp:207-213):
// Newton iteration X <- (X + X^-T) / 2, quadratically convergent
__device__ void polar_orthogonalize(double R[9]) {
  for (int it = 0; it < 4; ++it) {
    const double c00 = R[4] * R[8] - R[5] * R[7], c01 = R[5] * R[6] - R[3] * R[8], c02 = R[3] * R[7] - R[4] * R[6];
    const double c10 = R[2] * R[7] - R[1] * R[8], c11 = R[0] * R[8] - R[2] * R[6], c12 = R[1] * R[6] - R[0] * R[7];
    const double c20 = R[1] * R[5] - R[2] * R[4], c21 = R[2] * R[3] - R[0] * R[5], c22 = R[0] * R[4] - R[1] * R[3];
    const double det = R[0] * c00 + R[1] * c01 + R[2] * c02;
    if (!(fabs(det) > 1e-300)) return;
    const double id = 1.0 / det;   // X^-T = cofactor matrix / det
    const double C[9] = {c00 * id, c01 * id, c02 * id, c10 * id, c11 * id, c12 * id, c20 * id, c21 * id, c22 * id};
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = 0.5 * (R[i] + C[i]);
  }
}

// diagonally pivoted LDL^T solve of H x = g; returns the number of pivots used.
// A pivot is dropped when the remaining diagonal is below (max|H_jj| * eps^2) * (cnt-k)/cnt,
// the squared form of Eigen's ColPivHouseholderQR threshold_helper test (H_jj = |col j|^2).
// x must NOT live on the caller's stack: inlined into k_solve_update, nvcc 12.9 let a local x[] share a stack slot
// with A[][] (wrong results); the callers pass a shared-memory array.
__device__ __forceinline__ int solve_ldlt6(const double H21[21], const double g[6], double count, double x[6]) {
  double A[6][6];
  int t = 0;
  for (int p = 0; p < 6; ++p)
    for (int q = p; q < 6; ++q) { A[p][q] = H21[t]; A[q][p] = H21[t]; ++t; }
  int perm[6] = {0, 1, 2, 3, 4, 5};
  double rhs[6];
  for (int i = 0; i < 6; ++i) rhs[i] = g[i];
  double hmax = 0.0;
  for (int i = 0; i < 6; ++i) hmax = fmax(hmax, A[i][i]);
  const double helper = (hmax * DBL_EPSILON) * DBL_EPSILON / fmax(count, 1.0);
  int rank = 6;
  for (int k = 0; k < 6; ++k) {
    int piv = k;
    for (int j = k + 1; j < 6; ++j) if (A[j][j] > A[piv][piv]) piv = j;
    const double dk = A[piv][piv];
    if (!(dk > 0.0) || dk < helper * (count - k)) { rank = k; break; }
    if (piv != k) {
      for (int j = 0; j < 6; ++j) { const double tmp = A[k][j]; A[k][j] = A[piv][j]; A[piv][j] = tmp; }
      for (int j = 0; j < 6; ++j) { const double tmp = A[j][k]; A[j][k] = A[j][piv]; A[j][piv] = tmp; }
      const double tr = rhs[k]; rhs[k] = rhs[piv]; rhs[piv] = tr;
      const int tp = perm[k]; perm[k] = perm[piv]; perm[piv] = tp;
    }
    for (int i = k + 1; i < 6; ++i) {
      const double lik = A[k][i] / dk;   // row k stays unscaled (A[k][i] == a_ik), column k becomes L
      for (int j = k + 1; j <= i; ++j) { A[i][j] -= lik * A[k][j]; A[j][i] = A[i][j]; }
      A[i][k] = lik;
    }
  }
  for (int i = 0; i < 6; ++i) x[i] = 0.0;
  double y[6];
  for (int i = 0; i < rank; ++i) {          // L z = rhs
    double sacc = rhs[i];
    for (int j = 0; j < i; ++j) sacc -= A[i][j] * y[j];
    y[i] = sacc;
  }
  for (int i = 0; i < rank; ++i) y[i] /= A[i][i];   // D
  for (int i = rank - 1; i >= 0; --i) {     // L^T x = y
    double sacc = y[i];
    for (int j = i + 1; j < rank; ++j) sacc -= A[j][i] * y[j];
    y[i] = sacc;
  }
  for (int i = 0; i < rank; ++i) x[perm[i]] = y[i];
  return rank;
}

// x -> deltaTrans (src/solver.cpp:203-217): Rodrigues, orthogonal polar factor, translation
__device__ void delta_from_x(const double x[6], double D[16]) {
  double R[9];
  rodrigues(x, R);
  polar_orthogonalize(R);
  const double Dl[16] = {R[0], R[1], R[2], x[3], R[3], R[4], R[5], x[4], R[6], R[7], R[8], x[5], 0.0, 0.0, 0.0, 1.0};
  for (int i = 0; i < 16; ++i) D[i] = Dl[i];
}

// tail of one loop iteration (one thread): delta, rPose = delta * rPose (src/laser_odometry.cpp:619),
// convergence test (:628-646), loop condition of the resident graph
__device__ void finish_iteration(DevState* __restrict__ st, const DevParams& P, const double x[6], int rank, int advance_loop,
                                 cudaGraphConditionalHandle cond, int use_cond) {
  st->rank = rank;
  double D[16];
  delta_from_x(x, D);
  for (int i = 0; i < 16; ++i) st->delta[i] = D[i];
  const double dd = sqrt(x[3] * x[3] + x[4] * x[4] + x[5] * x[5]);   // :628-632
  double ct = ((D[0] + D[5] + D[10]) - 1.0) / 2.0;                   // :636-638
  ct = fmin(1.0, fmax(ct, -1.0));
  const double da = acos(ct);
  st->delta_dist = dd;
  st->delta_angle = da;
  if (!advance_loop) return;
  double nP[16];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double sacc = 0.0;
      for (int k = 0; k < 4; ++k) sacc += D[i * 4 + k] * st->rPose[k * 4 + j];
      nP[i * 4 + j] = sacc;
    }
  for (int i = 0; i < 16; ++i) st->rPose[i] = nP[i];   // :619
  st->iters += 1;
  st->use_prev = 1;   // the projection just consumed left its k-th distances behind
  // small step: the temporal bound is tight, short chunks balance best; large step: only the carry
  // bound along the scan order helps, long chunks amortise the greedy bound of each chunk head
  st->chunk = (dd < 0.05 && da < 0.01) ? PLO_CHUNK_WARM : PLO_CHUNK_COLD;
  st->warm = (dd < 0.05 && da < 0.01) ? 1 : 0;   // k_project: worth widening a refresh walk for the candidate cache
  if (dd < P.delta_dist_thr && da < P.delta_angle_thr) { st->status = PLO_REG_CONVERGED; st->done = 1; }   // :643-646
  else if (st->iters >= P.iterations) { st->status = PLO_REG_MAX_ITERS; st->done = 1; }
  if (use_cond) cudaGraphSetConditional(cond, st->done ? 0 : 1);   // WHILE node: run the body again?
}

// One thread: the reduced sums (PLO_NSUM values: 21 H, 6 g, sum w, sum w b^2, pair count, six drop counters) -> loop
// state, 6x6 solve, pose update.  stage 0: weighted LS (one pass).  Trimmed LS (src/solver.cpp:74-166): stage 1 =
// first solve on all pairs, only x0 is kept (:107); stage 2 = second solve on the pairs selected by residual rank
// (:137) + loop tail.  DRPM (:499-603): stage 3 = the sums only; k_drpm_eigen / k_drpm_noise / k_drpm_finish go on.
__device__ __noinline__ void solve_from_sums(const double* s_sum, DevState* __restrict__ st, const DevParams& P, int advance_loop,
                                cudaGraphConditionalHandle cond, int use_cond, int stage) {
  const double count = s_sum[29];
  double sw = s_sum[27];
  for (int i = 0; i < 21; ++i) st->H[i] = s_sum[i];
  for (int i = 0; i < 6; ++i) st->g[i] = s_sum[21 + i];
  st->sw = sw;
  st->swbb = s_sum[28];
  if (stage != 2) {   // the statistics describe the projection, not the trimmed subset
    st->pairs = (long long)count;
    for (int i = 0; i < 6; ++i) st->dropped[i] = (long long)s_sum[30 + i];
    st->rms = count > 0.0 ? sqrt(s_sum[28] / fmax(sw, 1e-300)) : 0.0;
  }
  if (advance_loop && count < (double)P.correspond_number) {   // src/laser_odometry.cpp:570-576
    st->status = PLO_REG_TOO_FEW_PAIRS;
    st->done = 1;
    if (use_cond) cudaGraphSetConditional(cond, 0);
    return;
  }
  double H[21], g[6];
  // weights are normalised to sum 1 in the reference (src/solver.cpp:361-364); same argmin
  const double scale = (!P.ext_weights && P.weight_mode == PLO_W_HUBER_EXP && sw > 0.0) ? 1.0 / sw : 1.0;
  bool finite = true;
  for (int i = 0; i < 21; ++i) { H[i] = s_sum[i] * scale; finite = finite && isfinite(H[i]); }
  for (int i = 0; i < 6; ++i) { g[i] = s_sum[21 + i] * scale; finite = finite && isfinite(g[i]); }
  if (stage == 3) return;
  __shared__ double x[6];   // see solve_ldlt6
  const int rank = solve_ldlt6(H, g, stage == 2 ? sw : count, x);
  if (advance_loop && stage != 1 && (rank == 0 || !finite)) {
    // no pivot at all (every pair had a zero row) or non-finite sums: the reference would carry NaN / a zero step
    // through its remaining iterations (src/laser_odometry.cpp:611-616 only breaks on `false`, which WeightedLS never
    // returns); here the loop ends with the pose of the previous iteration and says so
    st->rank = rank;
    st->status = PLO_REG_SOLVE_FAILED;
    st->done = 1;
    if (use_cond) cudaGraphSetConditional(cond, 0);
    return;
  }
  if (stage == 1) {
    st->rank = rank;
    for (int i = 0; i < 6; ++i) st->x0[i] = x[i];
    return;   // the loop condition keeps its value (1): the body goes on with the selection
  }
  for (int i = 0; i < 6; ++i) st->probs[i] = 0.0;
  finish_iteration(st, P, x, rank, advance_loop, cond, use_cond);
}

// ---- block-level reduction of the normal equations, shared by k_reduce_solve and k_register_loop ------------------
// Pairs are taken in rounds of gridDim.x * blockDim.x (pair i of a round belongs to global thread i); per round every
// value is summed over the warp by a fixed shuffle tree, rounds accumulate in order in lane 0, the warps of the block
// are summed in order: the result depends on the launch geometry only.  The pairs were written by other SMs during the
// same launch (k_register_loop): L2 loads.  s_red: [WARPS][PLO_NSUM] shared scratch.  Writes partial[0..PLO_NSUM).
__device__ __forceinline__ double warp_tree_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(PLO_FULL_MASK, v, o);
  return v;
}

template <int WARPS>
__device__ __forceinline__ void reduce_pairs_block(const float4* qx, const float4* qy, const float4* qn, int n_src, const DevParams& P,
                                                   double* __restrict__ partial, double (*s_red)[PLO_NSUM]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
    for (int t = 0; t < PLO_NSUM; ++t) s_red[warp][t] = 0.0;
  const int stride = (int)(gridDim.x * blockDim.x);
  const int rounds = (n_src + stride - 1) / stride;
  for (int r = 0; r < rounds; ++r) {
    const int i = r * stride + (int)(blockIdx.x * blockDim.x + threadIdx.x);
    int status = -1;
    double a[6] = {0, 0, 0, 0, 0, 0}, b = 0.0, w = 0.0;
    if (i < n_src) {
      const float4 x = __ldcg(&qx[i]);
      status = __float_as_int(x.w);
      if (status == PLO_PT_OK) {
        const float4 y = __ldcg(&qy[i]);
        const float4 nn = __ldcg(&qn[i]);
        // getXYZ / getNormals: float32 -> double (include/common.h:51-75)
        const double s[3] = {(double)x.x, (double)x.y, (double)x.z};
        const double d[3] = {(double)y.x, (double)y.y, (double)y.z};
        const double n[3] = {(double)nn.x, (double)nn.y, (double)nn.z};
        ab_row(s, d, n, a, b);
        w = 1.0;
        if (P.weight_mode == PLO_W_HUBER_EXP) {   // weights at the identity hypothesis (src/solver.cpp:334-364)
          const double I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
          w = huber_exp_weight(I4, s, d, n, P);
          if (w < 0.0) w = 0.0;   // not an inlier: counted as a pair, absent from the sums
        }
      }
    }
    int t = 0;
#pragma unroll
    for (int p = 0; p < 6; ++p)
#pragma unroll
      for (int q = p; q < 6; ++q) {
        const double v = warp_tree_sum(w * a[p] * a[q]);
        if (lane == 0) s_red[warp][t] += v;
        ++t;
      }
#pragma unroll
    for (int p = 0; p < 6; ++p) {
      const double v = warp_tree_sum(w * a[p] * b);
      if (lane == 0) s_red[warp][21 + p] += v;
    }
    {
      const double v = warp_tree_sum(w), v2 = warp_tree_sum(w * b * b);
      if (lane == 0) { s_red[warp][27] += v; s_red[warp][28] += v2; }
    }
#pragma unroll
    for (int sd = 0; sd <= 6; ++sd) {   // pair count and the six drop counters
      const int c = __popc(__ballot_sync(PLO_FULL_MASK, status == sd));
      if (lane == 0) s_red[warp][29 + sd] += (double)c;
    }
  }
  __syncthreads();
  if (threadIdx.x < PLO_NSUM) {
    double v = 0.0;
#pragma unroll
    for (int wv = 0; wv < WARPS; ++wv) v += s_red[wv][threadIdx.x];
    partial[threadIdx.x] = v;
  }
}

// block partials -> s_sum[PLO_NSUM].  Warp w takes the partials w, w + #warps, ... (lane = value index: one coalesced
// 288-byte row per partial, all of a thread's loads in flight together -- a sequential sum over 296 L2 round trips took
// 16 us), adds them in that order, then thread t adds the warps' sums in order.  Every block that runs this on the same
// partials gets the same bits.  s_stage: [#warps][PLO_NSUM] shared scratch (may alias s_sum's neighbourhood, not s_sum).
__device__ __forceinline__ void sum_block_partials(const double* partials, int n_partials, double* s_sum, double (*s_stage)[PLO_NSUM]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int t = half * 32 + lane;
    if (t < PLO_NSUM) {
      double acc = 0.0;
      for (int b0 = warp; b0 < n_partials; b0 += 8 * nwarps) {   // eight loads in flight, added in order
        double v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int b = b0 + j * nwarps;
          v[j] = (b < n_partials) ? __ldcg(&partials[(size_t)b * PLO_NSUM + t]) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += v[j];
      }
      s_stage[warp][t] = acc;
    }
  }
  __syncthreads();
  if (threadIdx.x < PLO_NSUM) {
    double acc = 0.0;
    for (int w = 0; w < nwarps; ++w) acc += s_stage[w][threadIdx.x];
    s_sum[threadIdx.x] = acc;
  }
  __syncthreads();
}

}  // namespace
