// p2plane_device.cuh — device functions of the point-to-plane solve shared by p2plane_solve.cu (stand-alone reduce /
// solve kernels) and knn_project.cu (the same sums and the same solve fused into the projection kernels' epilogue):
// row builder and weights (src/solver.cpp:185-198, :334-364), 6x6 LDL^T, Rodrigues + polar factor (:203-217), the
// loop tail of src/laser_odometry.cpp:619-646.  Internal linkage: included by several .cu files.
#pragma once

#include <float.h>
#include <math_constants.h>

#include "plo_internal.cuh"

#ifndef PLO_WARM_DIST
#define PLO_WARM_DIST 0.05    // a step below this (m, rad): the next refresh walk stores candidate tiles
#define PLO_WARM_ANGLE 0.01
#endif

namespace {

#ifdef PLO_LOOP_TIMING
static __device__ unsigned long long g_solve_stamp[8];
#define PLO_SOLVE_STAMP(i) do { if (blockIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_solve_stamp[i] = t_; } } while (0)
#else
#define PLO_SOLVE_STAMP(i) do { } while (0)
#endif

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

// ---- 6x6 solve + pose update (one warp) ---------------------------------------------

__device__ void rodrigues(const double r[3], double R[9]) {
  // Eigen AngleAxisd(rot.norm(), rot.normalized()).toRotationMatrix(); a zero vector stays zero
  const double z = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  const double angle = sqrt(z);
  double ax[3] = {r[0], r[1], r[2]};
  if (z > 0.0) {
#pragma unroll 1
    for (int i = 0; i < 3; ++i) ax[i] /= angle;
  }
  double sn, cs;
  sincos(angle, &sn, &cs);
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
#pragma unroll 1
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

// x -> deltaTrans (src/solver.cpp:203-217): Rodrigues, orthogonal polar factor, translation
__device__ void delta_from_x(const double x[6], double D[16]) {
  double R[9];
  rodrigues(x, R);
  polar_orthogonalize(R);
  const double Dl[16] = {R[0], R[1], R[2], x[3], R[3], R[4], R[5], x[4], R[6], R[7], R[8], x[5], 0.0, 0.0, 0.0, 1.0};
  for (int i = 0; i < 16; ++i) D[i] = Dl[i];
}

// Diagonally pivoted LDL^T solve of H x = g, warp-collective (all 32 lanes of one warp call it); returns the number of
// pivots used.  A pivot is dropped when the remaining diagonal is below (max|H_jj| * eps^2) * (cnt-k)/cnt, the squared form
// of Eigen's ColPivHouseholderQR threshold_helper test (H_jj = |col j|^2).  The 6 x 7 augmented matrix [H | g] lives in shared
// memory, one or two elements per lane (element e = row e / 7, column e % 7; a lane owns e = lane and e = lane + 32), so a pivot step is
// a scan of the six diagonal entries, two shared-memory reads per element, one division in parallel on every lane and
// one multiply-subtract.  The solve runs once per ICP iteration on code the projection has evicted from the
// instruction cache, so what it costs is its SIZE (a one-thread version over a local 6 x 6 array: 8 us of 15, measured
// with PLO_LOOP_TIMING; this one 2-3) -- the loops below are deliberately not unrolled.  Both copies of a symmetric pair
// are updated with the same expression, so the matrix stays exactly symmetric.
// H21: 21 packed upper entries, g6: right-hand side (any address space), both times `scale`; x: 6 doubles, shared memory.
__device__ __forceinline__ int solve_ldlt6_warp(const double* H21, const double* g6, double scale, double count, double* x) {
  __shared__ double s_buf[2][42];   // [6][7] row-major, column 6 = right-hand side; a pivot step reads one copy, writes the other
  const int lane = threadIdx.x & 31;
  double* M = s_buf[0];
  double* N = s_buf[1];
  // this lane's two elements: (i0, j0) = element lane, (i1, j1) = element lane + 32 (lanes 0-9 only)
  const bool two = lane < 10;
  const int e1 = two ? lane + 32 : 41;
  const int i0 = lane / 7, j0 = lane % 7, i1 = e1 / 7, j1 = e1 % 7;
  {
    const int a0 = min(i0, j0), b0 = max(i0, j0), a1 = min(i1, j1), b1 = max(i1, j1);
    M[lane] = (j0 == 6 ? g6[i0] : H21[a0 * 6 - a0 * (a0 - 1) / 2 + (b0 - a0)]) * scale;
    if (two) M[lane + 32] = (j1 == 6 ? g6[i1] : H21[a1 * 6 - a1 * (a1 - 1) / 2 + (b1 - a1)]) * scale;
  }
  if (lane < 6) x[lane] = 0.0;
  // row p is read at [hi] and [lo]: M[i][j] -= (M[p][hi] / dp) * M[p][lo], the same expression for (i, j) and (j, i)
  // (M[i][p] == M[p][i] by symmetry), so the matrix stays exactly symmetric
  const int hi0 = j0 == 6 ? i0 : max(i0, j0), lo0 = j0 == 6 ? 6 : min(i0, j0);
  const int hi1 = j1 == 6 ? i1 : max(i1, j1), lo1 = j1 == 6 ? 6 : min(i1, j1);
  __syncwarp();
  double dg[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) dg[i] = M[i * 8];
  double hmax = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) hmax = fmax(hmax, dg[i]);
  const double helper = (hmax * DBL_EPSILON) * DBL_EPSILON / fmax(count, 1.0);
  unsigned done = 0u, order = 0u;   // pivots taken (bit per index), their order (3 bits per step)
  int rank = 0;
#pragma unroll 1
  for (int k = 0; k < 6; ++k) {
    // largest remaining diagonal entry, smallest index on exact ties; every lane finds the same one
    int p = -1;
    double dp = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const double v = M[i * 8];
      if (!((done >> i) & 1u) && (p < 0 || v > dp)) { p = i; dp = v; }
    }
    if (!(dp > 0.0) || dp < helper * (count - k)) break;
    done |= 1u << p;
    order |= (unsigned)p << (3 * k);
    rank = k + 1;
    // rows and columns not yet taken (the right-hand side column always)
    const double m0 = M[lane], m1 = M[e1];
    const double a0 = M[p * 7 + hi0], b0 = M[p * 7 + lo0], a1 = M[p * 7 + hi1], b1 = M[p * 7 + lo1];
    const bool live0 = !((done >> i0) & 1u) && (j0 == 6 || !((done >> j0) & 1u));
    const bool live1 = !((done >> i1) & 1u) && (j1 == 6 || !((done >> j1) & 1u));
    N[lane] = live0 ? m0 - (a0 / dp) * b0 : m0;
    if (two) N[lane + 32] = live1 ? m1 - (a1 / dp) * b1 : m1;
    __syncwarp();
    double* t = M; M = N; N = t;
  }
  // back substitution in reverse pivot order, column by column: x[p] = rhs[p] / M[p][p], then rhs[i] -= M[i][p] x[p] on the
  // lanes (i < 6) of the pivots still to come
  double rhs = lane < 6 ? M[lane * 7 + 6] : 0.0;
#pragma unroll 1
  for (int k = rank - 1; k >= 0; --k) {
    const int p = (order >> (3 * k)) & 7;
    const double xp = __shfl_sync(PLO_FULL_MASK, rhs, p) / M[p * 8];
    if (lane == p) x[p] = xp;
    if (lane < 6) rhs -= M[lane * 7 + p] * xp;
  }
  __syncwarp();
  return rank;
}

// tail of one loop iteration, warp-collective: delta, rPose = delta * rPose (src/laser_odometry.cpp:619), convergence
// test (:628-646), loop condition of the resident graph.  x: shared memory.  The polar factor and the 4 x 4 product are
// spread over lanes (one matrix element each) for the same reason as the solve: fewer instructions to fetch.
__device__ __forceinline__ void finish_iteration(DevState* __restrict__ st, const DevParams& P, const double* x, int rank, int advance_loop,
                                                 cudaGraphConditionalHandle cond, int use_cond) {
  const int lane = threadIdx.x & 31;
  __shared__ double s_D[16];
  PLO_SOLVE_STAMP(3);
  if (lane == 0) {
    st->rank = rank;
    double R[9];
    rodrigues(x, R);
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) s_D[i * 4 + j] = R[i * 3 + j];
      s_D[i * 4 + 3] = x[3 + i];
      s_D[12 + i] = 0.0;
    }
    s_D[15] = 1.0;
  }
  __syncwarp();
  {   // polar_orthogonalize(), element (i, j) on lane 3 i + j: same expressions, same bits
    const int e = lane < 9 ? lane : 0, i = e / 3, j = e % 3;
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
      const double c = s_D[i1 * 4 + j1] * s_D[i2 * 4 + j2] - s_D[i1 * 4 + j2] * s_D[i2 * 4 + j1];
      const double det = s_D[0] * __shfl_sync(PLO_FULL_MASK, c, 0) + s_D[1] * __shfl_sync(PLO_FULL_MASK, c, 1) +
                         s_D[2] * __shfl_sync(PLO_FULL_MASK, c, 2);
      if (!(fabs(det) > 1e-300)) break;
      const double r = 0.5 * (s_D[i * 4 + j] + c * (1.0 / det));
      __syncwarp();
      if (lane < 9) s_D[i * 4 + j] = r;
      __syncwarp();
    }
  }
  PLO_SOLVE_STAMP(4);
  if (lane < 16) st->delta[lane] = s_D[lane];
  double dd = 0.0, da = 0.0;
  if (lane == 0) {
    dd = sqrt(x[3] * x[3] + x[4] * x[4] + x[5] * x[5]);          // :628-632
    double ct = ((s_D[0] + s_D[5] + s_D[10]) - 1.0) / 2.0;       // :636-638
    ct = fmin(1.0, fmax(ct, -1.0));
    da = acos(ct);
    st->delta_dist = dd;
    st->delta_angle = da;
  }
  if (!advance_loop) { __syncwarp(); return; }
  double np = 0.0;
  if (lane < 16) {
    const int i = lane >> 2, j = lane & 3;
#pragma unroll
    for (int k = 0; k < 4; ++k) np += s_D[i * 4 + k] * st->rPose[k * 4 + j];
  }
  __syncwarp();
  if (lane < 16) st->rPose[lane] = np;   // :619
  if (lane == 0) {
    st->iters += 1;
    st->use_prev = 1;   // the projection just consumed left its k-th distances behind
    // small step: the temporal bound is tight, short chunks balance best; large step: only the carry
    // bound along the scan order helps, long chunks amortise the greedy bound of each chunk head
    st->chunk = (dd < 0.05 && da < 0.01) ? PLO_CHUNK_WARM : PLO_CHUNK_COLD;
    st->warm = (dd < PLO_WARM_DIST && da < PLO_WARM_ANGLE) ? 1 : 0;   // k_project: worth widening a refresh walk for the candidate cache
    if (dd < P.delta_dist_thr && da < P.delta_angle_thr) { st->status = PLO_REG_CONVERGED; st->done = 1; }   // :643-646
    else if (st->iters >= P.iterations) { st->status = PLO_REG_MAX_ITERS; st->done = 1; }
    if (use_cond) cudaGraphSetConditional(cond, st->done ? 0 : 1);   // WHILE node: run the body again?
  }
  __syncwarp();
  PLO_SOLVE_STAMP(5);
}

// One WARP (all 32 lanes call it; lane 0 does the scalar work): the reduced sums (PLO_NSUM values: 21 H, 6 g, sum w,
// sum w b^2, pair count, six drop counters) -> loop state, 6x6 solve, pose update.  stage 0: weighted LS (one pass).
// Trimmed LS (src/solver.cpp:74-166): stage 1 = first solve on all pairs, only x0 is kept (:107); stage 2 = second solve on
// the pairs selected by residual rank (:137) + loop tail.  DRPM (:499-603): stage 3 = the sums only; k_drpm_eigen /
// k_drpm_noise / k_drpm_finish go on.  s_sum and x live in shared memory.
__device__ __noinline__ void solve_from_sums(const double* s_sum, DevState* __restrict__ st, const DevParams& P, int advance_loop,
                                             cudaGraphConditionalHandle cond, int use_cond, int stage) {
  const int lane = threadIdx.x & 31;
  PLO_SOLVE_STAMP(0);
  const double count = s_sum[29];
  const double sw = s_sum[27];
  if (lane < 21) st->H[lane] = s_sum[lane];
  if (lane < 6) st->g[lane] = s_sum[21 + lane];
  if (lane == 0) {
    st->sw = sw;
    st->swbb = s_sum[28];
    if (stage != 2) {   // the statistics describe the projection, not the trimmed subset
      st->pairs = (long long)count;
      for (int i = 0; i < 6; ++i) st->dropped[i] = (long long)s_sum[30 + i];
      st->rms = count > 0.0 ? sqrt(s_sum[28] / fmax(sw, 1e-300)) : 0.0;
    }
  }
  if (advance_loop && count < (double)P.correspond_number) {   // src/laser_odometry.cpp:570-576
    if (lane == 0) {
      st->status = PLO_REG_TOO_FEW_PAIRS;
      st->done = 1;
      if (use_cond) cudaGraphSetConditional(cond, 0);
    }
    __syncwarp();
    return;
  }
  // weights are normalised to sum 1 in the reference (src/solver.cpp:361-364); same argmin
  const double scale = (!P.ext_weights && P.weight_mode == PLO_W_HUBER_EXP && sw > 0.0) ? 1.0 / sw : 1.0;
  bool finite = (lane < 27) ? isfinite(s_sum[lane] * scale) : true;
  finite = __all_sync(PLO_FULL_MASK, finite);
  if (stage == 3) { __syncwarp(); return; }
  __shared__ double x[6];
  PLO_SOLVE_STAMP(1);
  const int rank = solve_ldlt6_warp(s_sum, s_sum + 21, scale, stage == 2 ? sw : count, x);
  PLO_SOLVE_STAMP(2);
  if (advance_loop && stage != 1 && (rank == 0 || !finite)) {
    // no pivot at all (every pair had a zero row) or non-finite sums: the reference would carry NaN / a zero step
    // through its remaining iterations (src/laser_odometry.cpp:611-616 only breaks on `false`, which WeightedLS never
    // returns); here the loop ends with the pose of the previous iteration and says so
    if (lane == 0) {
      st->rank = rank;
      st->status = PLO_REG_SOLVE_FAILED;
      st->done = 1;
      if (use_cond) cudaGraphSetConditional(cond, 0);
    }
  } else if (stage == 1) {
    if (lane == 0) st->rank = rank;
    if (lane < 6) st->x0[lane] = x[lane];   // the loop condition keeps its value (1): the body goes on with the selection
  } else {
    if (lane < 6) st->probs[lane] = 0.0;
    finish_iteration(st, P, x, rank, advance_loop, cond, use_cond);
  }
  __syncwarp();
}

// ---- block-level reduction of the normal equations, shared by k_reduce_solve and k_register_loop ------------------
// Pairs are taken in rounds of gridDim.x * blockDim.x (pair i of a round belongs to global thread i); per round every
// value is summed over the warp by a fixed shuffle tree, rounds accumulate in order (one lane per value), the warps of the block
// are summed in order: the result depends on the launch geometry only.  The pairs were written by other SMs during the
// same launch (k_register_loop): L2 loads.  s_red: [WARPS][PLO_NSUM] shared scratch.  Writes partial[0..PLO_NSUM).
__device__ __forceinline__ double warp_tree_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(PLO_FULL_MASK, v, o);
  return v;
}

// The butterfly sums of SIXTEEN values at once: the same operand pairs as warp_tree_sum (xor 16, 8, 4, 2, 1), hence bitwise
// the same totals, but after each exchange a lane keeps only half of the values it held -- 8 + 4 + 2 + 1 + 1 exchanges
// instead of 16 x 5.  On return lane l holds the total of value warp_tree_slot16(l) (lanes l and l ^ 1 the same one).
__device__ __forceinline__ int warp_tree_slot16(int lane) {
  return (((lane >> 4) & 1) << 3) | (((lane >> 3) & 1) << 2) | (((lane >> 2) & 1) << 1) | ((lane >> 1) & 1);
}
__device__ __forceinline__ double warp_tree_sum16(double (&v)[16], int lane) {
#pragma unroll
  for (int width = 8; width >= 1; width >>= 1) {
    const bool up = (lane & (2 * width)) != 0;   // xor distance 2 * width: the upper partner keeps the upper half
#pragma unroll
    for (int i = 0; i < width; ++i) {
      const double send = up ? v[i] : v[i + width];
      const double keep = up ? v[i + width] : v[i];
      v[i] = keep + __shfl_xor_sync(PLO_FULL_MASK, send, 2 * width);
    }
  }
  return v[0] + __shfl_xor_sync(PLO_FULL_MASK, v[0], 1);
}

template <int WARPS>
__device__ __forceinline__ void reduce_pairs_block(const float4* qx, const float4* qy, const float4* qn, int n_src, const DevParams& P,
                                                   double* __restrict__ partial, double (*s_red)[PLO_NSUM]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
    for (int t = 0; t < PLO_NSUM; ++t) s_red[warp][t] = 0.0;
  __syncwarp();
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
    // 21 + 6 + 2 butterfly sums, sixteen at a time (warp_tree_sum16: the same totals as one warp_tree_sum per value);
    // value t: 0..20 = w a_p a_q (p <= q, row-major), 21..26 = w a_p b, 27 = w, 28 = w b^2
    const int tl = warp_tree_slot16(lane);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      double v[16];
      int t = 0;
#pragma unroll
      for (int p = 0; p < 6; ++p)
#pragma unroll
        for (int q = p; q < 6; ++q) {
          if ((t >> 4) == half) v[t & 15] = w * a[p] * a[q];
          ++t;
        }
#pragma unroll
      for (int p = 0; p < 6; ++p) {
        if ((t >> 4) == half) v[t & 15] = w * a[p] * b;
        ++t;
      }
      if (half == 1) {
        v[11] = w;           // t = 27
        v[12] = w * b * b;   // t = 28
        v[13] = 0.0; v[14] = 0.0; v[15] = 0.0;
      }
      const double tot = warp_tree_sum16(v, lane);
      if (!(lane & 1) && half * 16 + tl < 29) s_red[warp][half * 16 + tl] += tot;
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

// block partials -> s_sum[PLO_NSUM].  The block splits into groups of PLO_NSUM threads (thread t of a group = value t: one
// coalesced 288-byte row per partial); group g takes the partials g, g + #groups, ... with up to twelve loads of a thread in
// flight together (two L2 round trips for 296 partials and 512 threads -- a sequential sum over 296 round trips took 16 us,
// a lane-per-value layout with its second, 4-lane pass six round trips = 3 us), adds them in that order, then thread t
// adds the groups' sums in order.  Every block that runs this on the same partials gets the same bits.
// s_stage: [>= blockDim / PLO_NSUM][PLO_NSUM] shared scratch (may alias s_sum's neighbourhood, not s_sum).
__device__ __forceinline__ void sum_block_partials(const double* partials, int n_partials, double* s_sum, double (*s_stage)[PLO_NSUM]) {
  const int groups = (int)blockDim.x / PLO_NSUM;
  const int g = (int)threadIdx.x / PLO_NSUM, t = (int)threadIdx.x % PLO_NSUM;
  if (g < groups) {
    double acc = 0.0;
    for (int b0 = g; b0 < n_partials; b0 += 12 * groups) {
      double v[12];
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        const int b = b0 + j * groups;
        v[j] = (b < n_partials) ? __ldcg(&partials[(size_t)b * PLO_NSUM + t]) : 0.0;
      }
#pragma unroll
      for (int j = 0; j < 12; ++j) acc += v[j];
    }
    s_stage[g][t] = acc;
  }
  __syncthreads();
  if (threadIdx.x < PLO_NSUM) {
    double acc = 0.0;
    for (int gg = 0; gg < groups; ++gg) acc += s_stage[gg][threadIdx.x];
    s_sum[threadIdx.x] = acc;
  }
  __syncthreads();
}

}  // namespace
