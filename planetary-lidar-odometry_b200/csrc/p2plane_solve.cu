// p2plane_solve.cu — point-to-plane weighted least squares on the device:
// SolveMotionEstimationProblemWeightedLS (src/solver.cpp:168-220) and the driver-loop
// tail of src/laser_odometry.cpp:570-576 (too few pairs), :619 (rPose = delta * rPose),
// :628-646 (convergence test).
//
// The reference stacks A (M' x 6), b and runs a column-pivoted Householder QR.  Here the
// 21 unique entries of H = A^T W A, the 6 of g = A^T W b (+ sum w, sum w b^2, the pair
// count and the six drop counters) are reduced in fp64 — thread-strided partial sums in a
// fixed order, a fixed shuffle/shared-memory tree per block, a fixed sequential sum over
// the block partials — so results are bitwise reproducible for a given launch geometry.
// One thread then solves H x = g by diagonally-pivoted LDL^T (the same pivot order a
// column-pivoted QR of A takes), dropping pivots that are exactly degenerate the way
// Eigen's ColPivHouseholderQR::nonzeroPivots does, maps x through Rodrigues
// (AngleAxisd, src/solver.cpp:203-205) and the orthogonal polar factor (the U*V^T of
// :207-213), composes the pose and raises the device-side `done` flag.
//
// Algorithmic bytes: 36 B (x, y, n as float32) per source point read, 36 doubles per
// block written.  Roofline: HBM (negligible next to the projection kernel).
#include <float.h>
#include <math_constants.h>

#include <algorithm>

#include "plo_internal.cuh"
#include "plo_scan.cuh"

namespace {

constexpr int kReduceThreads = 256;

__device__ __forceinline__ void ab_row(const double s[3], const double d[3], const double n[3], double a[6], double& b) {
  // src/solver.cpp:185-192
  a[0] = __dsub_rn(__dmul_rn(n[2], s[1]), __dmul_rn(n[1], s[2]));
  a[1] = __dsub_rn(__dmul_rn(n[0], s[2]), __dmul_rn(n[2], s[0]));
  a[2] = __dsub_rn(__dmul_rn(n[1], s[0]), __dmul_rn(n[0], s[1]));
  a[3] = n[0]; a[4] = n[1]; a[5] = n[2];
  b = __dadd_rn(__dadd_rn(__dmul_rn(n[0], __dsub_rn(d[0], s[0])), __dmul_rn(n[1], __dsub_rn(d[1], s[1]))),
                __dmul_rn(n[2], __dsub_rn(d[2], s[2])));
}

// RANSAC-final weight of src/solver.cpp:334-364 evaluated at T_best = I; < 0 => not an inlier
__device__ __forceinline__ double huber_exp_weight(const double s[3], const double d[3], const double n[3], const DevParams& P) {
  const double dist = fabs(__dadd_rn(__dadd_rn(__dmul_rn(__dsub_rn(s[0], d[0]), n[0]), __dmul_rn(__dsub_rn(s[1], d[1]), n[1])),
                                     __dmul_rn(__dsub_rn(s[2], d[2]), n[2])));
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

__device__ __forceinline__ void block_reduce_store(double acc[PLO_NSUM], double* __restrict__ partial) {
  __shared__ double s_red[kReduceThreads / 32][PLO_NSUM];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < PLO_NSUM; ++t) {
    double v = acc[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(PLO_FULL_MASK, v, o);
    if (lane == 0) s_red[warp][t] = v;
  }
  __syncthreads();
  if (threadIdx.x < PLO_NSUM) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kReduceThreads / 32; ++w) v += s_red[w][threadIdx.x];
    partial[threadIdx.x] = v;
  }
}

// pairs of the last projection -> per-block partial sums
__global__ void __launch_bounds__(kReduceThreads) k_reduce_pairs(const float4* __restrict__ qx, const float4* __restrict__ qy,
                                                                 const float4* __restrict__ qn,
                                                                 const DevCounts* __restrict__ counts,
                                                                 const DevState* __restrict__ st, DevParams P,
                                                                 double* __restrict__ partials, int respect_done,
                                                                 const int* __restrict__ mask) {
  if (respect_done && st->done) return;
  double acc[PLO_NSUM];
#pragma unroll
  for (int t = 0; t < PLO_NSUM; ++t) acc[t] = 0.0;
  const int n_src = counts->n_source;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_src; i += gridDim.x * blockDim.x) {
    const float4 x = __ldg(&qx[i]);
    const int status = __float_as_int(x.w);
    if (status != PLO_PT_OK) { acc[30 + status - 1] += 1.0; continue; }
    const float4 y = __ldg(&qy[i]);
    const float4 nn = __ldg(&qn[i]);
    // getXYZ / getNormals: float32 -> double (include/common.h:51-75)
    const double s[3] = {(double)x.x, (double)x.y, (double)x.z};
    const double d[3] = {(double)y.x, (double)y.y, (double)y.z};
    const double n[3] = {(double)nn.x, (double)nn.y, (double)nn.z};
    acc[29] += 1.0;
    if (mask != nullptr && mask[i] == 0) continue;   // trimmed LS, second pass: pair outside the kept rank window
    double w = 1.0;
    if (P.weight_mode == PLO_W_HUBER_EXP) {
      w = huber_exp_weight(s, d, n, P);
      if (w < 0.0) continue;
    }
    accumulate_pair(acc, s, d, n, w);
  }
  block_reduce_store(acc, partials + (size_t)blockIdx.x * PLO_NSUM);
}

// reference-shaped inputs (n x 3 doubles each, optional weights) -> per-block partial sums
__global__ void __launch_bounds__(kReduceThreads) k_reduce_host_pairs(const double* __restrict__ src, const double* __restrict__ ref,
                                                                      const double* __restrict__ nrm, const double* __restrict__ w,
                                                                      long long n, double* __restrict__ partials) {
  double acc[PLO_NSUM];
#pragma unroll
  for (int t = 0; t < PLO_NSUM; ++t) acc[t] = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double s[3] = {src[3 * i], src[3 * i + 1], src[3 * i + 2]};
    const double d[3] = {ref[3 * i], ref[3 * i + 1], ref[3 * i + 2]};
    const double nn[3] = {nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]};
    acc[29] += 1.0;
    accumulate_pair(acc, s, d, nn, w ? w[i] : 1.0);
  }
  block_reduce_store(acc, partials + (size_t)blockIdx.x * PLO_NSUM);
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
__device__ int solve_ldlt6(const double H21[21], const double g[6], double count, double x[6]) {
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

__global__ void __launch_bounds__(1024) k_solve_update(const double* __restrict__ partials, int n_partials, DevState* __restrict__ st,
                                                       DevParams P, int advance_loop, cudaGraphConditionalHandle cond, int use_cond,
                                                       int stage) {
  // stage 0: weighted LS (one pass).  Trimmed LS (src/solver.cpp:74-166): stage 1 = first solve on all pairs,
  // only x0 is kept (:107); stage 2 = second solve on the pairs selected by residual rank (:137) + loop tail.
  if (advance_loop && st->done) {
    if (use_cond && threadIdx.x == 0) cudaGraphSetConditional(cond, 0);
    return;
  }
  __shared__ double s_sum[PLO_NSUM];
  {
    // value t is summed by warp (t mod 32): lane-strided partial sums in a fixed order, fixed shuffle tree
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t = warp; t < PLO_NSUM; t += 32) {
      double v = 0.0;
      for (int b = lane; b < n_partials; b += 32) v += partials[(size_t)b * PLO_NSUM + t];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(PLO_FULL_MASK, v, o);
      if (lane == 0) s_sum[t] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
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
  const double scale = (P.weight_mode == PLO_W_HUBER_EXP && sw > 0.0) ? 1.0 / sw : 1.0;
  for (int i = 0; i < 21; ++i) H[i] = s_sum[i] * scale;
  for (int i = 0; i < 6; ++i) g[i] = s_sum[21 + i] * scale;
  double x[6];
  const int rank = solve_ldlt6(H, g, stage == 2 ? sw : count, x);
  st->rank = rank;
  if (stage == 1) {
    for (int i = 0; i < 6; ++i) st->x0[i] = x[i];
    return;   // the loop condition keeps its value (1): the body goes on with the selection
  }
  double R[9];
  rodrigues(x, R);
  polar_orthogonalize(R);
  double D[16] = {R[0], R[1], R[2], x[3], R[3], R[4], R[5], x[4], R[6], R[7], R[8], x[5], 0.0, 0.0, 0.0, 1.0};
  for (int i = 0; i < 16; ++i) st->delta[i] = D[i];
  const double dd = sqrt(x[3] * x[3] + x[4] * x[4] + x[5] * x[5]);   // :628-632
  double ct = ((R[0] + R[4] + R[8]) - 1.0) / 2.0;                    // :636-638
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
  if (dd < P.delta_dist_thr && da < P.delta_angle_thr) { st->status = PLO_REG_CONVERGED; st->done = 1; }   // :643-646
  else if (st->iters >= P.iterations) { st->status = PLO_REG_MAX_ITERS; st->done = 1; }
  if (use_cond) cudaGraphSetConditional(cond, st->done ? 0 : 1);   // WHILE node: run the body again?
}

// trimmed LS: |A_i x0 - b_i| of every surviving pair as a sortable 64-bit key (src/solver.cpp:110-122);
// dropped points and the padding sort last.  Ties keep the pair order (stable sort, value = index).
__global__ void __launch_bounds__(256) k_ls_keys(const float4* __restrict__ qx, const float4* __restrict__ qy,
                                                 const float4* __restrict__ qn, const DevCounts* __restrict__ counts,
                                                 const DevState* __restrict__ st, int respect_done, int m_raw,
                                                 unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  if (respect_done && st->done) return;
  const int n_src = counts->n_source;
  double x0[6];
#pragma unroll
  for (int a = 0; a < 6; ++a) x0[a] = st->x0[a];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m_raw; i += gridDim.x * blockDim.x) {
    unsigned long long key = 0xffffffffffffffffull;
    if (i < n_src) {
      const float4 x = __ldg(&qx[i]);
      if (__float_as_int(x.w) == PLO_PT_OK) {
        const float4 y = __ldg(&qy[i]);
        const float4 nn = __ldg(&qn[i]);
        const double s[3] = {(double)x.x, (double)x.y, (double)x.z};
        const double d[3] = {(double)y.x, (double)y.y, (double)y.z};
        const double n[3] = {(double)nn.x, (double)nn.y, (double)nn.z};
        double a[6], b;
        ab_row(s, d, n, a, b);
        double r = 0.0;
#pragma unroll
        for (int t = 0; t < 6; ++t) r += a[t] * x0[t];
        key = (unsigned long long)__double_as_longlong(fabs(r - b));   // non-negative doubles order like their bits
      }
    }
    keys[i] = key;
    vals[i] = i;
  }
}

// rank window [thr*N, (1-thr)*N] of the sorted pairs -> per-source-point mask (src/solver.cpp:124-134)
__global__ void __launch_bounds__(256) k_ls_select(const int* __restrict__ vals_sorted, const DevState* __restrict__ st,
                                                   int respect_done, int m_raw, double threshold, int* __restrict__ mask) {
  if (respect_done && st->done) return;
  const long long N = st->pairs;
  const long long lower = (long long)(threshold * (double)N);
  long long upper = (long long)((1.0 - threshold) * (double)N);
  if (upper > N - 1) upper = N - 1;   // the reference reads one past the end at threshold = 0
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m_raw; j += gridDim.x * blockDim.x)
    mask[vals_sorted[j]] = (j >= lower && j <= upper) ? 1 : 0;
}

__global__ void k_init_state(DevState* st, const double* T0, int use_prev) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) {
      const double id = (i % 5 == 0) ? 1.0 : 0.0;
      st->rPose[i] = T0 ? T0[i] : id;
      st->delta[i] = id;
    }
    for (int i = 0; i < 21; ++i) st->H[i] = 0.0;
    for (int i = 0; i < 6; ++i) { st->g[i] = 0.0; st->dropped[i] = 0; st->x0[i] = 0.0; }
    st->sw = st->swbb = st->rms = st->delta_dist = st->delta_angle = 0.0;
    st->pairs = 0;
    st->iters = 0;
    st->status = 0;
    st->rank = 0;
    st->done = 0;
    st->use_prev = use_prev;
    st->chunk = use_prev ? PLO_CHUNK_MID : PLO_CHUNK_COLD;
  }
}

// ---- order-preserving compaction of the surviving pairs (plo_get_pairs, D4) -----------

__global__ void __launch_bounds__(256) k_pairs_count(const int* __restrict__ status, const DevCounts* __restrict__ counts,
                                                     int* __restrict__ blockcnt) {
  __shared__ int s_c[8];
  const int n = counts->n_source;
  const int base = blockIdx.x * kTile;
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < kTile / 256; ++j) {
    const int i = base + j * 256 + threadIdx.x;
    cnt += (i < n && status[i] == PLO_PT_OK) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(PLO_FULL_MASK, cnt, o);
  if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += s_c[w];
    blockcnt[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) k_pairs_scatter(const int* __restrict__ status, const float4* __restrict__ qx,
                                                       const float4* __restrict__ qy, const float4* __restrict__ qn,
                                                       const DevCounts* __restrict__ counts, const int* __restrict__ blockoff,
                                                       float* __restrict__ src, float* __restrict__ ref, float* __restrict__ nrm,
                                                       int* __restrict__ idx) {
  const int n = counts->n_source;
  const int base = blockIdx.x * kTile;
  bool ok[kTile / 256];
#pragma unroll
  for (int j = 0; j < kTile / 256; ++j) {
    const int i = base + j * 256 + threadIdx.x;
    ok[j] = i < n && status[i] == PLO_PT_OK;
  }
  int rank[kTile / 256];
  tile_ranks(ok, rank);
  const int off = blockoff[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kTile / 256; ++j) {
    const int i = base + j * 256 + threadIdx.x;
    if (!ok[j]) continue;
    const int o = off + rank[j];
    const float4 x = qx[i], y = qy[i], nn = qn[i];
    src[3 * o] = x.x; src[3 * o + 1] = x.y; src[3 * o + 2] = x.z;
    ref[3 * o] = y.x; ref[3 * o + 1] = y.y; ref[3 * o + 2] = y.z;
    nrm[3 * o] = nn.x; nrm[3 * o + 1] = nn.y; nrm[3 * o + 2] = nn.z;
    idx[o] = i;
  }
}

int reduce_grid(const plo_ctx* c, int64_t n) {
  const int64_t want = (n + kReduceThreads - 1) / kReduceThreads;
  return (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)plo_grid(c, 2)));
}

}  // namespace

int plo_launch_init_state(plo_ctx* c, const double* T0_host_or_null) {
  const double* dT0 = nullptr;
  if (T0_host_or_null) {
    PLO_CUDA(c, c->scratch.reserve(sizeof(double) * 16));
    // pageable -> device: the copy is staged by the runtime before the call returns
    PLO_CUDA(c, cudaMemcpyAsync(c->scratch.p, T0_host_or_null, sizeof(double) * 16, cudaMemcpyHostToDevice, c->stream));
    dT0 = c->scratch.as<double>();
  }
  k_init_state<<<1, 32, 0, c->stream>>>(c->state.as<DevState>(), dT0, c->prev_valid ? 1 : 0);
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  return PLO_OK;
}

constexpr int kLsPasses = 7;   // 64-bit keys, 10-bit digits

int plo_reserve_solver_buffers(plo_ctx* c) {
  PLO_CUDA(c, c->partials.reserve(sizeof(double) * PLO_NSUM * (size_t)plo_grid(c, 2)));
  if (c->dprm.solver == PLO_SOLVER_LS && c->m_raw > 0) {
    const size_t m = (size_t)c->m_raw;
    for (int a = 0; a < 2; ++a) {
      PLO_CUDA(c, c->ls_keys[a].reserve(sizeof(unsigned long long) * m));
      PLO_CUDA(c, c->ls_vals[a].reserve(sizeof(int) * m));
    }
    PLO_CUDA(c, c->ls_hist.reserve(sizeof(int) * plo_sort_hist_ints(c->m_raw)));
    PLO_CUDA(c, c->ls_tot.reserve(sizeof(int) * plo_sort_total_ints(kLsPasses)));
    PLO_CUDA(c, c->ls_mask.reserve(sizeof(int) * m));
  }
  return PLO_OK;
}

int plo_launch_reduce_solve(plo_ctx* c, bool advance_loop, unsigned long long cond_handle) {
  const int g = reduce_grid(c, c->m_raw);
  PLO_TRY(plo_reserve_solver_buffers(c));
  const int adv = advance_loop ? 1 : 0;
  const cudaGraphConditionalHandle cond = (cudaGraphConditionalHandle)cond_handle;
  const int use_cond = cond_handle ? 1 : 0;
  const bool trimmed = c->dprm.solver == PLO_SOLVER_LS && c->m_raw > 0;
  DevParams P = c->dprm;
  if (trimmed) P.weight_mode = PLO_W_UNIT;   // SolveMotionEstimationProblemLS is unweighted
  if (c->m_raw > 0) {
    k_reduce_pairs<<<g, kReduceThreads, 0, c->stream>>>(c->q_x.as<float4>(), c->q_y.as<float4>(), c->q_n.as<float4>(),
                                                        c->counts.as<DevCounts>(), c->state.as<DevState>(), P,
                                                        c->partials.as<double>(), adv, nullptr);
    c->launches++;
    PLO_CUDA(c, cudaGetLastError());
  }
  k_solve_update<<<1, 1024, 0, c->stream>>>(c->partials.as<double>(), c->m_raw > 0 ? g : 0, c->state.as<DevState>(), P, adv, cond,
                                            use_cond, trimmed ? 1 : 0);
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  if (!trimmed) return PLO_OK;
  // ---- trimmed LS: residual keys -> stable sort -> rank window -> second reduce + solve ----
  const int m = (int)c->m_raw;
  const int gk = (int)std::max<int64_t>(1, std::min<int64_t>((m + 255) / 256, (int64_t)plo_grid(c, 4)));
  k_ls_keys<<<gk, 256, 0, c->stream>>>(c->q_x.as<float4>(), c->q_y.as<float4>(), c->q_n.as<float4>(), c->counts.as<DevCounts>(),
                                       c->state.as<DevState>(), adv, m, c->ls_keys[0].as<unsigned long long>(),
                                       c->ls_vals[0].as<int>());
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  unsigned long long* keys[2] = {c->ls_keys[0].as<unsigned long long>(), c->ls_keys[1].as<unsigned long long>()};
  int* vals[2] = {c->ls_vals[0].as<int>(), c->ls_vals[1].as<int>()};
  int which = 0;
  PLO_TRY(plo_sort_pairs(c, keys, vals, m, kLsPasses, c->ls_hist.as<int>(), c->ls_tot.as<int>(), &which));
  k_ls_select<<<gk, 256, 0, c->stream>>>(vals[which], c->state.as<DevState>(), adv, m, c->dprm.ls_threshold, c->ls_mask.as<int>());
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  k_reduce_pairs<<<g, kReduceThreads, 0, c->stream>>>(c->q_x.as<float4>(), c->q_y.as<float4>(), c->q_n.as<float4>(),
                                                      c->counts.as<DevCounts>(), c->state.as<DevState>(), P,
                                                      c->partials.as<double>(), adv, c->ls_mask.as<int>());
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  k_solve_update<<<1, 1024, 0, c->stream>>>(c->partials.as<double>(), g, c->state.as<DevState>(), P, adv, cond, use_cond, 2);
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  return PLO_OK;
}

int plo_launch_reduce_solve_host_pairs(plo_ctx* c, const double* d_src, const double* d_ref, const double* d_nrm,
                                       const double* d_w, int64_t n) {
  const int g = reduce_grid(c, n);
  PLO_CUDA(c, c->partials.reserve(sizeof(double) * PLO_NSUM * (size_t)plo_grid(c, 2)));
  if (n > 0) {
    k_reduce_host_pairs<<<g, kReduceThreads, 0, c->stream>>>(d_src, d_ref, d_nrm, d_w, (long long)n, c->partials.as<double>());
    c->launches++;
    PLO_CUDA(c, cudaGetLastError());
  }
  DevParams P = c->dprm;
  P.weight_mode = PLO_W_UNIT;   // caller-supplied weights are used as they are
  k_solve_update<<<1, 1024, 0, c->stream>>>(c->partials.as<double>(), n > 0 ? g : 0, c->state.as<DevState>(), P, 0,
                                            (cudaGraphConditionalHandle)0, 0, 0);
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  return PLO_OK;
}

int plo_launch_compact_pairs(plo_ctx* c, float* d_src, float* d_ref, float* d_nrm, int32_t* d_idx) {
  DevCounts* dc = c->counts.as<DevCounts>();
  if (c->m_raw == 0) {
    PLO_CUDA(c, cudaMemsetAsync(&dc->n_pairs, 0, sizeof(int), c->stream));
    return PLO_OK;
  }
  const int nb = (int)((c->m_raw + kTile - 1) / kTile);
  PLO_CUDA(c, c->blockcnt.reserve(sizeof(int) * (size_t)(nb + 1)));
  k_pairs_count<<<nb, 256, 0, c->stream>>>(c->q_status.as<int>(), dc, c->blockcnt.as<int>());
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  k_scan_exclusive<<<1, 1024, 0, c->stream>>>(c->blockcnt.as<int>(), nb, &dc->n_pairs);
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  k_pairs_scatter<<<nb, 256, 0, c->stream>>>(c->q_status.as<int>(), c->q_x.as<float4>(), c->q_y.as<float4>(), c->q_n.as<float4>(),
                                             dc, c->blockcnt.as<int>(), d_src, d_ref, d_nrm, d_idx);
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  return PLO_OK;
}
