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
#include "p2plane_device.cuh"

namespace {

constexpr int kReduceThreads = 256;

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
                                                                 const int* __restrict__ mask, int inliers_unit,
                                                                 const double* __restrict__ w_ext) {
  if (respect_done && st->done) return;
  double acc[PLO_NSUM];
#pragma unroll
  for (int t = 0; t < PLO_NSUM; ++t) acc[t] = 0.0;
  const int n_src = counts->n_source;
  double Tw[16];   // hypothesis the Huber/exp weights are evaluated at
#pragma unroll
  for (int t = 0; t < 16; ++t) Tw[t] = (P.solver == PLO_SOLVER_RANSAC) ? st->Tbest[t] : ((t % 5 == 0) ? 1.0 : 0.0);
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
    if (P.ext_weights) {
      if (w_ext != nullptr) w = w_ext[i];   // caller-supplied weights (host-vector solver entry points), used as they are
    } else if (P.weight_mode == PLO_W_HUBER_EXP) {
      w = huber_exp_weight(Tw, s, d, n, P);
      if (w < 0.0) continue;
      if (inliers_unit) w = 1.0;   // RANSAC -> "LS": the inlier subset, unweighted (src/solver.cpp:366-371)
    }
    accumulate_pair(acc, s, d, n, w);
  }
  block_reduce_store(acc, partials + (size_t)blockIdx.x * PLO_NSUM);
}

// Weighted-LS loop outside k_register_loop (profiling / enqueue-all mode, graph fallback): reduce + solve + loop tail
// in ONE launch.  Every block reduces its share of the pairs (reduce_pairs_block: the partition and order of
// k_register_loop for the same geometry); the last block to finish sums the block partials (fixed order), solves and
// advances the loop state -- bitwise the result of k_register_loop.
constexpr int kFusedReduceWarps = 16;
__global__ void __launch_bounds__(kFusedReduceWarps * 32) k_reduce_solve(const float4* __restrict__ qx, const float4* __restrict__ qy,
                                                                         const float4* __restrict__ qn,
                                                                         const DevCounts* __restrict__ counts, DevState* __restrict__ st,
                                                                         DevParams P, double* __restrict__ partials,
                                                                         int* __restrict__ done_ticket, cudaGraphConditionalHandle cond,
                                                                         int use_cond) {
  if (st->done) {
    if (use_cond && blockIdx.x == 0 && threadIdx.x == 0) cudaGraphSetConditional(cond, 0);
    return;
  }
  __shared__ double s_red[kFusedReduceWarps][PLO_NSUM];
  reduce_pairs_block<kFusedReduceWarps>(qx, qy, qn, counts->n_source, P, partials + (size_t)blockIdx.x * PLO_NSUM, s_red);
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(done_ticket, 1) == (int)gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x == 0) *done_ticket = 0;
  __shared__ double s_sum[PLO_NSUM];
  sum_block_partials(partials, (int)gridDim.x, s_sum, s_red);
  if (threadIdx.x < 32) solve_from_sums(s_sum, st, P, 1, cond, use_cond, 0);   // warp 0, collectively
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

// Eigen::ColPivHouseholderQR::solve for the 3 x 6 hypothesis system of src/solver.cpp:251-273 (basic
// solution: three pivot columns solved, the other three unknowns zero) — same steps as Eigen 3.3's
// computeInPlace / _solve_impl: largest remaining column norm first, LAPACK-style norm down-dating.
__device__ __noinline__ void colpiv_qr_solve_3x6(double A[3][6], double b[3], double x[6]) {
  const int m = 3, n = 6, size = 3;
  double cnd[6], cnu[6], hcoef[3];
  int perm[6];
  double maxnorm = 0.0;
  for (int j = 0; j < n; ++j) {
    double sacc = 0.0;
    for (int i = 0; i < m; ++i) sacc += A[i][j] * A[i][j];
    cnd[j] = cnu[j] = sqrt(sacc);
    if (cnu[j] > maxnorm) maxnorm = cnu[j];
    perm[j] = j;
  }
  const double eps = DBL_EPSILON;
  const double threshold_helper = (maxnorm * eps) * (maxnorm * eps) / (double)m;
  const double downdate_thr = sqrt(eps);
  int nonzero_pivots = size;
  for (int k = 0; k < size; ++k) {
    int big = k;
    double bigv = cnu[k];
    for (int j = k + 1; j < n; ++j) if (cnu[j] > bigv) { bigv = cnu[j]; big = j; }
    if (nonzero_pivots == size && bigv * bigv < threshold_helper * (double)(m - k)) nonzero_pivots = k;
    if (big != k) {
      for (int i = 0; i < m; ++i) { const double t = A[i][k]; A[i][k] = A[i][big]; A[i][big] = t; }
      double t = cnu[k]; cnu[k] = cnu[big]; cnu[big] = t;
      t = cnd[k]; cnd[k] = cnd[big]; cnd[big] = t;
      const int ti = perm[k]; perm[k] = perm[big]; perm[big] = ti;
    }
    const double c0 = A[k][k];
    double tail_sq = 0.0;
    for (int i = k + 1; i < m; ++i) tail_sq += A[i][k] * A[i][k];
    double tau, beta;
    if (tail_sq <= DBL_MIN) {
      tau = 0.0; beta = c0;
      for (int i = k + 1; i < m; ++i) A[i][k] = 0.0;
    } else {
      beta = sqrt(c0 * c0 + tail_sq);
      if (c0 >= 0.0) beta = -beta;
      for (int i = k + 1; i < m; ++i) A[i][k] /= (c0 - beta);
      tau = (beta - c0) / beta;
    }
    hcoef[k] = tau;
    A[k][k] = beta;
    if (tau != 0.0) {
      for (int j = k + 1; j < n; ++j) {
        double sacc = A[k][j];
        for (int i = k + 1; i < m; ++i) sacc += A[i][k] * A[i][j];
        sacc *= tau;
        A[k][j] -= sacc;
        for (int i = k + 1; i < m; ++i) A[i][j] -= sacc * A[i][k];
      }
    }
    for (int j = k + 1; j < n; ++j) {
      if (cnu[j] != 0.0) {
        double temp = fabs(A[k][j]) / cnu[j];
        temp = (1.0 + temp) * (1.0 - temp);
        if (temp < 0.0) temp = 0.0;
        const double ratio = cnu[j] / cnd[j];
        const double temp2 = temp * ratio * ratio;
        if (temp2 <= downdate_thr) {
          double sacc = 0.0;
          for (int i = k + 1; i < m; ++i) sacc += A[i][j] * A[i][j];
          cnd[j] = cnu[j] = sqrt(sacc);
        } else {
          cnu[j] *= sqrt(temp);
        }
      }
    }
  }
  for (int j = 0; j < n; ++j) x[j] = 0.0;
  if (nonzero_pivots == 0) return;
  for (int k = 0; k < nonzero_pivots; ++k) {
    const double tau = hcoef[k];
    if (tau == 0.0) continue;
    double sacc = b[k];
    for (int i = k + 1; i < m; ++i) sacc += A[i][k] * b[i];
    sacc *= tau;
    b[k] -= sacc;
    for (int i = k + 1; i < m; ++i) b[i] -= sacc * A[i][k];
  }
  double y[3];
  for (int i = nonzero_pivots - 1; i >= 0; --i) {
    double sacc = b[i];
    for (int j = i + 1; j < nonzero_pivots; ++j) sacc -= A[i][j] * y[j];
    y[i] = sacc / A[i][i];
  }
  for (int i = 0; i < nonzero_pivots; ++i) x[perm[i]] = y[i];
}

// DRPM, src/solver.cpp:537-542: eigen-decomposition of the weight-normalised information matrix — cyclic Jacobi
// on a symmetric 6 x 6, eigenvalues ascending, eigenvectors in columns (the ordering contract of
// Eigen::SelfAdjointEigenSolver).  One warp: the rotation order is the serial row-cyclic one and every element
// update is the same single expression as in the oracle's sym_eigen_n (bit-identical results), but the six
// independent updates of each of a rotation's three loops run on six lanes and the matrices sit in shared memory
// (the one-thread version cost 0.75 ms per ICP iteration).
__global__ void __launch_bounds__(32) k_drpm_eigen(DevState* __restrict__ st, DevParams P, int advance_loop) {
  if (advance_loop && st->done) return;
  __shared__ double A[36], V[36];
  const int lane = threadIdx.x;
  {
    const double sw = st->sw;
    const double scale = (!P.ext_weights && P.weight_mode == PLO_W_HUBER_EXP && sw > 0.0) ? 1.0 / sw : 1.0;   // as k_solve_update
    if (lane == 0) {
      int t = 0;
      for (int a = 0; a < 6; ++a)
        for (int b = a; b < 6; ++b) { const double h = st->H[t] * scale; A[a * 6 + b] = h; A[b * 6 + a] = h; ++t; }
    }
    for (int i = lane; i < 36; i += 32) V[i] = (i % 7 == 0) ? 1.0 : 0.0;
  }
  __syncwarp();
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < 6; ++i)
      for (int j = i + 1; j < 6; ++j) off += A[i * 6 + j] * A[i * 6 + j];
    if (off == 0.0) break;   // warp-uniform: every lane read the same values
    {   // the rotated entry is not zeroed, so `off` bottoms out at rounding noise: same early stop as the oracle
      double dsum = 0.0;
      for (int i = 0; i < 6; ++i) dsum += A[i * 7] * A[i * 7];
      if (off <= 1e-30 * dsum) break;
    }
    for (int pp = 0; pp < 6; ++pp)
      for (int q = pp + 1; q < 6; ++q) {
        const double apq = A[pp * 6 + q];
        if (apq == 0.0) continue;
        const double theta = (A[q * 6 + q] - A[pp * 6 + pp]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
        __syncwarp();
        if (lane < 6) {
          const int k = lane;
          const double akp = A[k * 6 + pp], akq = A[k * 6 + q];
          A[k * 6 + pp] = cs * akp - sn * akq;
          A[k * 6 + q] = sn * akp + cs * akq;
        } else if (lane >= 8 && lane < 14) {
          const int k = lane - 8;
          const double vkp = V[k * 6 + pp], vkq = V[k * 6 + q];
          V[k * 6 + pp] = cs * vkp - sn * vkq;
          V[k * 6 + q] = sn * vkp + cs * vkq;
        }
        __syncwarp();
        if (lane < 6) {
          const int k = lane;
          const double apk = A[pp * 6 + k], aqk = A[q * 6 + k];
          A[pp * 6 + k] = cs * apk - sn * aqk;
          A[q * 6 + k] = sn * apk + cs * aqk;
        }
        __syncwarp();
      }
  }
  if (lane == 0) {
    int order[6] = {0, 1, 2, 3, 4, 5};
    for (int i = 0; i < 6; ++i)
      for (int j = i + 1; j < 6; ++j)
        if (A[order[j] * 7] < A[order[i] * 7]) { const int t = order[i]; order[i] = order[j]; order[j] = t; }
    for (int i = 0; i < 6; ++i) {
      st->ev[i] = A[order[i] * 7];
      for (int k = 0; k < 6; ++k) st->U[k * 6 + i] = V[k * 6 + order[i]];
    }
  }
}

__global__ void __launch_bounds__(256) k_solve_update(const double* __restrict__ partials, int n_partials, DevState* __restrict__ st,
                                                       DevParams P, int advance_loop, cudaGraphConditionalHandle cond, int use_cond,
                                                       int stage) {
  // stage 0: weighted LS (one pass).  Trimmed LS (src/solver.cpp:74-166): stage 1 = first solve on all pairs,
  // only x0 is kept (:107); stage 2 = second solve on the pairs selected by residual rank (:137) + loop tail.
  // DRPM (src/solver.cpp:499-603): stage 3 = eigen-decomposition of the weighted information matrix; the
  // noise estimate and the solve follow in k_drpm_noise / k_drpm_finish.
  if (advance_loop && st->done) {
    if (use_cond && threadIdx.x == 0) cudaGraphSetConditional(cond, 0);
    return;
  }
  __shared__ double s_sum[PLO_NSUM];
  __shared__ double s_stage[8][PLO_NSUM];
  sum_block_partials(partials, n_partials, s_sum, s_stage);
  if (threadIdx.x >= 32) return;
  solve_from_sums(s_sum, st, P, advance_loop, cond, use_cond, stage);   // warp 0, collectively
}

// ---- RANSAC front (src/solver.cpp:238-326) on the compacted pairs, one block --------------------

__device__ __forceinline__ unsigned long long xorshift64(unsigned long long x) {
  x ^= x << 13; x ^= x >> 7; x ^= x << 17;
  return x;
}

// block-wide argmax of (value, index): larger value wins, ties go to the smaller index; index -1 = none
__device__ __forceinline__ void block_argmax(double v, int idx, double* s_val, int* s_idx, int* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(PLO_FULL_MASK, v, o);
    const int oi = __shfl_xor_sync(PLO_FULL_MASK, idx, o);
    if (oi >= 0 && (idx < 0 || ov > v || (ov == v && oi < idx))) { v = ov; idx = oi; }
  }
  if (lane == 0) { s_val[warp] = v; s_idx[warp] = idx; }
  __syncthreads();
  if (warp == 0) {
    v = s_val[lane]; idx = s_idx[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(PLO_FULL_MASK, v, o);
      const int oi = __shfl_xor_sync(PLO_FULL_MASK, idx, o);
      if (oi >= 0 && (idx < 0 || ov > v || (ov == v && oi < idx))) { v = ov; idx = oi; }
    }
    if (lane == 0) *out = idx;
  }
  __syncthreads();
}

__device__ __forceinline__ void load3(const float* __restrict__ a, int i, double out[3]) {
  out[0] = (double)a[3 * (size_t)i]; out[1] = (double)a[3 * (size_t)i + 1]; out[2] = (double)a[3 * (size_t)i + 2];
}

__device__ __forceinline__ double dist3(const double a[3], const double b[3]) {
  const double dx = __dsub_rn(a[0], b[0]), dy = __dsub_rn(a[1], b[1]), dz = __dsub_rn(a[2], b[2]);
  return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}

// ---- RANSAC hypotheses (src/solver.cpp:229-331), one hypothesis per block, the grid strides over them -------------
// The reference draws, fits and counts hypothesis after hypothesis and stops at the first one whose inlier count exceeds
// ransac_min_inliers_percentage.  Hypothesis `it` depends on the random stream only (its first point is draw `it` of
// xorshift64, the other two follow from farthest-point sampling), so hypotheses are evaluated independently: block b takes
// it = first + b, first + b + grid, ... in increasing order, writes inlier count and transform of each into the scratch
// table and lowers `stop` to the smallest `it` that passes the test; a block gives up as soon as its `it` lies beyond
// `stop` (checked between the three passes over the pairs as well).  Every hypothesis up to the final `stop` is therefore
// evaluated, and k_ransac_pick replays the reference's sequential rule over the table: same best hypothesis, same count
// of hypotheses, whatever the grid.  Launched twice per solve: hypothesis 0 alone (in the usual case it already passes,
// and 295 other blocks reading the pair list would only slow it down), then the rest over the whole GPU -- the worst case
// (the exit never fires: 5000 hypotheses x three passes over ~130 k pairs) takes a few ms instead of seconds in one block.
struct RansacScratch {
  long long* cnt;   // [ransac_max_iterations] inliers of hypothesis it
  double* T;        // [ransac_max_iterations][16]
  int* stop;        // smallest `it` whose count passes the exit test; INT_MAX while none does
};

__global__ void __launch_bounds__(1024) k_ransac_eval(const float* __restrict__ src, const float* __restrict__ ref,
                                                      const float* __restrict__ nrm, const DevCounts* __restrict__ counts,
                                                      const DevState* __restrict__ st, DevParams P, RansacScratch rs, int it_first,
                                                      int it_end, int respect_done) {
  if (respect_done && st->done) return;
  __shared__ double s_val[32];
  __shared__ int s_idx[32];
  __shared__ long long s_cnt[32];
  __shared__ double s_T[16];
  __shared__ int s_sel[3];
  __shared__ int s_stop[3];   // `stop` as thread 0 saw it before each pass (one slot per check: no write races a read)
  const int n = counts->n_pairs;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  volatile int* vstop = rs.stop;
  if (it_first == 0 && blockIdx.x == 0 && tid == 0) *vstop = INT_MAX;   // first launch of this solve (one block)
  if (n <= 0) return;
  const int min_inliers = (int)(P.ransac_min_inliers_pct * (double)n);   // :238
  unsigned long long rng = P.ransac_seed;
  int drawn = 0;
  for (int it = it_first + (int)blockIdx.x; it < it_end; it += (int)gridDim.x) {   // :244
    // ---- farthestPointSampling(source_cloud, 3), src/common.cpp:19-82 ----
    if (tid == 0) {
      for (; drawn <= it; ++drawn) rng = xorshift64(rng);   // draw `it` of the stream
      s_sel[0] = (int)(rng % (unsigned long long)n);
      s_stop[0] = (it_first > 0 || it > 0) ? *vstop : INT_MAX;
    }
    __syncthreads();
    if (s_stop[0] < it) break;
    const int first = s_sel[0];
    double pf[3];
    load3(src, first, pf);
    double bv = -1.0;
    int bi = -1;
    for (int i = tid; i < n; i += 1024) {
      double pi[3];
      load3(src, i, pi);
      const double d = dist3(pf, pi);
      if (i != first && d > bv) { bv = d; bi = i; }
    }
    if (tid == 0) s_stop[1] = *vstop;
    block_argmax(bv, bi, s_val, s_idx, &s_sel[1]);
    if (s_stop[1] < it) break;
    int second = s_sel[1];
    const bool have_second = second >= 0;
    if (!have_second) second = first;
    double ps[3];
    load3(src, second, ps);
    bv = -1.0;
    bi = -1;
    for (int i = tid; i < n; i += 1024) {
      double pi[3];
      load3(src, i, pi);
      double m = dist3(pf, pi);   // distance to the chosen set (recomputed: no per-hypothesis array)
      if (have_second) {
        const double d = dist3(ps, pi);
        if (d < m) m = d;
      }
      if (i != first && i != second && m > bv) { bv = m; bi = i; }
    }
    block_argmax(bv, bi, s_val, s_idx, &s_sel[2]);
    if (tid == 0) {
      int ids[3] = {first, second, s_sel[2] >= 0 ? s_sel[2] : first};
      double A[3][6], b[3], x[6];
      for (int r = 0; r < 3; ++r) {   // :255-270
        double sv[3], dv[3], nv[3];
        load3(src, ids[r], sv); load3(ref, ids[r], dv); load3(nrm, ids[r], nv);
        ab_row(sv, dv, nv, A[r], b[r]);
      }
      colpiv_qr_solve_3x6(A, b, x);   // :273
      double D[16];
      delta_from_x(x, D);             // :276-298
      for (int i = 0; i < 16; ++i) s_T[i] = D[i];
      s_stop[2] = *vstop;
    }
    __syncthreads();
    if (s_stop[2] < it) break;
    // ---- inlier count, :300-314 ----
    double T[16];
    for (int i = 0; i < 16; ++i) T[i] = s_T[i];
    long long cnt = 0;
    for (int i = tid; i < n; i += 1024) {
      double sv[3], dv[3], nv[3];
      load3(src, i, sv); load3(ref, i, dv); load3(nrm, i, nv);
      cnt += (plane_distance(T, sv, dv, nv) < P.ransac_dist_thr) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(PLO_FULL_MASK, cnt, o);
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
      long long total = 0;
      for (int w = 0; w < 32; ++w) total += s_cnt[w];
      rs.cnt[it] = total;
      for (int i = 0; i < 16; ++i) rs.T[(size_t)it * 16 + i] = s_T[i];
      if (total > (long long)min_inliers) atomicMin(rs.stop, it);   // :323-325
    }
    __syncthreads();
  }
}

// the reference's sequential bookkeeping over the table (one warp): best = first hypothesis with the largest count among
// those drawn, drawing ends with the first hypothesis after which best > min_inliers (:317-325)
__global__ void __launch_bounds__(32) k_ransac_pick(const DevCounts* __restrict__ counts, DevState* __restrict__ st, DevParams P,
                                                    RansacScratch rs, int respect_done) {
  if (respect_done && st->done) return;
  const int n = counts->n_pairs;
  const int lane = threadIdx.x;
  long long best = 0;
  int best_it = -1, iters = 0;
  if (n > 0) {
    for (int base = 0; base < P.ransac_max_iterations; base += 32) {
      const int it = base + lane;
      const bool in = it < P.ransac_max_iterations;
      // entries beyond the first passing one may never have been written: read one at a time up to it
      const int stop = *rs.stop;
      long long v = (in && it <= stop) ? rs.cnt[it] : -1;
      int vi = it;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const long long ov = __shfl_xor_sync(PLO_FULL_MASK, v, o);
        const int oi = __shfl_xor_sync(PLO_FULL_MASK, vi, o);
        if (ov > v || (ov == v && oi < vi)) { v = ov; vi = oi; }
      }
      if (v > best) { best = v; best_it = vi; }
      const int last = min(min(base + 31, P.ransac_max_iterations - 1), stop);
      iters = last + 1;
      if (stop <= base + 31) break;
    }
  }
  if (lane < 16) st->Tbest[lane] = best_it >= 0 ? rs.T[(size_t)best_it * 16 + lane] : ((lane % 5 == 0) ? 1.0 : 0.0);
  if (lane == 0) {
    st->ransac_best = best;
    st->ransac_iters = iters;
  }
}

// ---- DRPM noise estimate: degeneracy::ComputeNoiseEstimate, include/degeneracy.h:14-72 ----------
constexpr int kNoiseSums = 42;   // 36 mean + 6 variance

__global__ void __launch_bounds__(kReduceThreads) k_drpm_noise(const float4* __restrict__ qx, const float4* __restrict__ qy,
                                                               const float4* __restrict__ qn,
                                                               const DevCounts* __restrict__ counts,
                                                               const DevState* __restrict__ st, DevParams P,
                                                               double* __restrict__ partials2, int respect_done,
                                                               const double* __restrict__ w_ext) {
  if (respect_done && st->done) return;
  __shared__ double s_U[36];
  __shared__ double s_T[16];
  __shared__ double s_red[kReduceThreads / 32][kNoiseSums];
  if (threadIdx.x < 36) s_U[threadIdx.x] = st->U[threadIdx.x];
  if (threadIdx.x < 16) s_T[threadIdx.x] = st->Tbest[threadIdx.x];
  __syncthreads();
  const double inv_sw = P.ext_weights ? 1.0 : (st->sw > 0.0 ? 1.0 / st->sw : 0.0);
  double acc[kNoiseSums];
#pragma unroll
  for (int t = 0; t < kNoiseSums; ++t) acc[t] = 0.0;
  const int n_src = counts->n_source;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_src; i += gridDim.x * blockDim.x) {
    const float4 x = __ldg(&qx[i]);
    if (__float_as_int(x.w) != PLO_PT_OK) continue;
    const float4 y = __ldg(&qy[i]);
    const float4 nn = __ldg(&qn[i]);
    const double p[3] = {(double)x.x, (double)x.y, (double)x.z};
    const double d[3] = {(double)y.x, (double)y.y, (double)y.z};
    const double n[3] = {(double)nn.x, (double)nn.y, (double)nn.z};
    double w = P.ext_weights ? (w_ext != nullptr ? w_ext[i] : 1.0) : huber_exp_weight(s_T, p, d, n, P);
    if (w < 0.0) continue;
    w *= inv_sw;   // weights normalised to sum 1 (src/solver.cpp:361-364); caller-supplied ones are used as they are
    // B = [[-nx, px*nx], [0, nx]] (6x6), Ncov = diag(sp2 I3, sn2 I3)   (:39-51)
    const double nx[9] = {0, -n[2], n[1], n[2], 0, -n[0], -n[1], n[0], 0};
    const double px[9] = {0, -p[2], p[1], p[2], 0, -p[0], -p[1], p[0], 0};
    double B[6][6];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c) B[r][c] = 0.0;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) sacc += px[r * 3 + k] * nx[k * 3 + c];
        B[r][c] = -nx[r * 3 + c];
        B[r][3 + c] = sacc;
        B[3 + r][3 + c] = nx[r * 3 + c];
      }
    const double sq = sqrt(w);
    const double v[6] = {sq * (px[0] * n[0] + px[1] * n[1] + px[2] * n[2]), sq * (px[3] * n[0] + px[4] * n[1] + px[5] * n[2]),
                         sq * (px[6] * n[0] + px[7] * n[1] + px[8] * n[2]), sq * n[0], sq * n[1], sq * n[2]};   // :58-60
    double a[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      double Crow[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) sacc += B[r][k] * (k < 3 ? P.drpm_sp2 : P.drpm_sn2) * B[c][k];
        Crow[c] = sacc * w;   // :53
        acc[r * 6 + c] += Crow[c];
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) {   // a_k = u_k^T C u_k, accumulated row by row
        double t = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) t += Crow[c] * s_U[c * 6 + k];
        a[k] += s_U[r * 6 + k] * t;
      }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {   // :63-69
      double bb = 0.0;
#pragma unroll
      for (int r = 0; r < 6; ++r) bb += s_U[r * 6 + k] * v[r];
      acc[36 + k] += 2.0 * a[k] * a[k] + 4.0 * a[k] * bb * bb;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < kNoiseSums; ++t) {
    double vv = acc[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vv += __shfl_xor_sync(PLO_FULL_MASK, vv, o);
    if (lane == 0) s_red[warp][t] = vv;
  }
  __syncthreads();
  if (threadIdx.x < kNoiseSums) {
    double vv = 0.0;
#pragma unroll
    for (int w2 = 0; w2 < kReduceThreads / 32; ++w2) vv += s_red[w2][threadIdx.x];
    partials2[(size_t)blockIdx.x * kNoiseSums + threadIdx.x] = vv;
  }
}

// probabilities (degeneracy::ComputeSignalToNoiseProbabilities, include/degeneracy.h:74-105), the
// probability-weighted pseudo-inverse (:107-131) or the plain weighted solve, then the loop tail
__global__ void __launch_bounds__(64) k_drpm_finish(const double* __restrict__ partials2, int n_partials, DevState* __restrict__ st,
                                                    DevParams P, int advance_loop, cudaGraphConditionalHandle cond, int use_cond) {
  if (advance_loop && st->done) {
    if (use_cond && threadIdx.x == 0) cudaGraphSetConditional(cond, 0);
    return;
  }
  __shared__ double s_sum[kNoiseSums];
  if (threadIdx.x < kNoiseSums) {
    double v = 0.0;
    for (int b = 0; b < n_partials; ++b) v += partials2[(size_t)b * kNoiseSums + threadIdx.x];   // fixed order
    s_sum[threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x >= 32) return;   // warp 0: lane 0 for the probabilities, all lanes for the solve and the loop tail
  const double sw = st->sw;
  const double scale = P.ext_weights ? 1.0 : (sw > 0.0 ? 1.0 / sw : 1.0);
  __shared__ double x[6];
  __shared__ int s_weighted;
  if (threadIdx.x == 0) {
    double H[21], g[6], Hf[36];
    for (int i = 0; i < 21; ++i) H[i] = st->H[i] * scale;
    for (int i = 0; i < 6; ++i) g[i] = st->g[i] * scale;
    int t = 0;
    for (int a = 0; a < 6; ++a)
      for (int b = a; b < 6; ++b) { Hf[a * 6 + b] = H[t]; Hf[b * 6 + a] = H[t]; ++t; }
    double probs[6], minp = CUDART_INF;
    for (int k = 0; k < 6; ++k) {
      double meas = 0.0, noise = 0.0;
      for (int r = 0; r < 6; ++r) {
        double t1 = 0.0, t2 = 0.0;
        for (int c = 0; c < 6; ++c) { t1 += Hf[r * 6 + c] * st->U[c * 6 + k]; t2 += s_sum[r * 6 + c] * st->U[c * 6 + k]; }
        meas += st->U[r * 6 + k] * t1;
        noise += st->U[r * 6 + k] * t2;
      }
      const double sd = sqrt(s_sum[36 + k]);
      const double test_point = meas / (1.0 + 10.0);   // snr_factor = 10, src/solver.cpp:547
      double pr;
      if (!(noise == noise) || !(sd == sd) || !(test_point == test_point)) pr = 0.0;
      else if (!(sd > 0.0)) pr = test_point >= noise ? 1.0 : 0.0;
      else pr = 0.5 * erfc(-(test_point - noise) / (sd * sqrt(2.0)));   // boost normal cdf
      probs[k] = pr;
      st->probs[k] = pr;
      if (pr < minp) minp = pr;
    }
    s_weighted = minp < P.drpm_threshold;
    if (s_weighted) {   // SolveWithSnrProbabilities, include/degeneracy.h:107-131
      double tt[6];
      for (int i = 0; i < 6; ++i) {
        double sacc = 0.0;
        for (int r = 0; r < 6; ++r) sacc += st->U[r * 6 + i] * g[r];
        const double dps = fabs(st->ev[i]) > 1e-10 ? probs[i] / st->ev[i] : 0.0;
        tt[i] = sacc * dps;
      }
      for (int r = 0; r < 6; ++r) {
        double sacc = 0.0;
        for (int i = 0; i < 6; ++i) sacc += st->U[r * 6 + i] * tt[i];
        x[r] = sacc;
      }
    }
  }
  __syncwarp();
  int rank = 6;
  if (!s_weighted) rank = solve_ldlt6_warp(st->H, st->g, scale, (double)st->pairs, x);   // weighted_A.colPivHouseholderQr().solve(weighted_b), :576
  finish_iteration(st, P, x, rank, advance_loop, cond, use_cond);
}

// trimmed LS: |A_i x0 - b_i| of every surviving pair as a sortable 64-bit key (src/solver.cpp:110-122);
// dropped points and the padding sort last.  Ties keep the pair order (stable sort, value = index).
__global__ void __launch_bounds__(256) k_ls_keys(const float4* __restrict__ qx, const float4* __restrict__ qy,
                                                 const float4* __restrict__ qn, const DevCounts* __restrict__ counts,
                                                 const DevState* __restrict__ st, int respect_done, int m_raw,
                                                 unsigned long long* __restrict__ keys, int* __restrict__ vals, DevParams P,
                                                 int inliers_only) {
  if (respect_done && st->done) return;
  const int n_src = counts->n_source;
  double x0[6];
#pragma unroll
  for (int a = 0; a < 6; ++a) x0[a] = st->x0[a];
  double Tw[16];   // RANSAC -> "LS": only the inliers of the best hypothesis enter the trimmed solve
#pragma unroll
  for (int t = 0; t < 16; ++t) Tw[t] = inliers_only ? st->Tbest[t] : ((t % 5 == 0) ? 1.0 : 0.0);
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
        if (!inliers_only || huber_exp_weight(Tw, s, d, n, P) >= 0.0)
          key = (unsigned long long)__double_as_longlong(fabs(r - b));   // non-negative doubles order like their bits
      }
    }
    keys[i] = key;
    vals[i] = i;
  }
}

// rank window [thr*N, (1-thr)*N] of the sorted pairs -> per-source-point mask (src/solver.cpp:124-134)
__global__ void __launch_bounds__(256) k_ls_select(const int* __restrict__ vals_sorted, const DevState* __restrict__ st,
                                                   int respect_done, int m_raw, double threshold, int* __restrict__ mask,
                                                   int n_from_weights) {
  if (respect_done && st->done) return;
  // rows of the first pass: all surviving pairs, or (RANSAC -> "LS") the inliers, whose unit weights sum to their count
  const long long N = n_from_weights ? (long long)st->sw : st->pairs;
  const long long lower = (long long)(threshold * (double)N);
  long long upper = (long long)((1.0 - threshold) * (double)N);
  if (upper > N - 1) upper = N - 1;   // the reference reads one past the end at threshold = 0
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m_raw; j += gridDim.x * blockDim.x)
    mask[vals_sorted[j]] = (j >= lower && j <= upper) ? 1 : 0;
}

__global__ void k_init_state(DevState* st, const double* T0, int use_prev, int force_warm, unsigned* loop_barrier) {
  if (threadIdx.x == 0) {
    if (loop_barrier) loop_barrier[0] = 0u;   // arrival counter of k_register_loop's grid barriers
    const int tiles_ready = use_prev ? st->tiles_ready : 0;   // tiles outlive a registration of the same clouds
    for (int i = 0; i < 16; ++i) {
      const double id = (i % 5 == 0) ? 1.0 : 0.0;
      st->rPose[i] = T0 ? T0[i] : id;
      st->delta[i] = id;
    }
    for (int i = 0; i < 21; ++i) st->H[i] = 0.0;
    for (int i = 0; i < 6; ++i) { st->g[i] = 0.0; st->dropped[i] = 0; st->x0[i] = 0.0; st->ev[i] = 0.0; st->probs[i] = 0.0; }
    for (int i = 0; i < 16; ++i) st->Tbest[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int i = 0; i < 36; ++i) st->U[i] = 0.0;
    st->ransac_best = 0;
    st->ransac_iters = 0;
    st->pad0 = 0;
    st->sw = st->swbb = st->rms = st->delta_dist = st->delta_angle = 0.0;
    st->pairs = 0;
    st->iters = 0;
    st->status = 0;
    st->rank = 0;
    st->done = 0;
    st->use_prev = use_prev;
    st->chunk = use_prev ? PLO_CHUNK_MID : PLO_CHUNK_COLD;
    st->warm = force_warm;
    st->tiles_ready = tiles_ready;
    for (int i = 0; i < 32; ++i) st->miss_hist[i] = -1;
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
  PLO_CUDA(c, c->loop_barrier.reserve(sizeof(unsigned) * 2));
  k_init_state<<<1, 32, 0, c->stream>>>(c->state.as<DevState>(), dT0, c->prev_valid ? 1 : 0, c->tune_force_warm ? 1 : 0, c->loop_barrier.as<unsigned>());
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  return PLO_OK;
}

constexpr int kLsPasses = 7;   // 64-bit keys, 10-bit digits

int plo_reserve_solver_buffers(plo_ctx* c) {
  PLO_CUDA(c, c->partials.reserve(sizeof(double) * PLO_NSUM * (size_t)plo_grid(c, 2)));
  if (!c->reduce_ticket.p) {
    PLO_CUDA(c, c->reduce_ticket.reserve(sizeof(int)));
    PLO_CUDA(c, cudaMemsetAsync(c->reduce_ticket.p, 0, sizeof(int), c->stream));
  }
  const bool trims = c->dprm.solver == PLO_SOLVER_LS || (c->dprm.solver == PLO_SOLVER_RANSAC && c->dprm.ransac_final == PLO_FINAL_LS);
  if (trims && c->m_raw > 0) {
    const size_t m = (size_t)c->m_raw;
    for (int a = 0; a < 2; ++a) {
      PLO_CUDA(c, c->ls_keys[a].reserve(sizeof(unsigned long long) * m));
      PLO_CUDA(c, c->ls_vals[a].reserve(sizeof(int) * m));
    }
    PLO_CUDA(c, c->ls_hist.reserve(sizeof(int) * plo_sort_hist_ints(c->m_raw)));
    PLO_CUDA(c, c->ls_tot.reserve(sizeof(int) * plo_sort_total_ints(kLsPasses)));
    PLO_CUDA(c, c->ls_mask.reserve(sizeof(int) * m));
  }
  if ((c->dprm.solver == PLO_SOLVER_RANSAC || c->host_drpm_only) && c->m_raw > 0) {
    const size_t m = (size_t)c->m_raw;
    PLO_CUDA(c, c->h_src.reserve(sizeof(double) * 3 * m));   // also sized for plo_solve_wls_host
    PLO_CUDA(c, c->h_ref.reserve(sizeof(double) * 3 * m));
    PLO_CUDA(c, c->h_nrm.reserve(sizeof(double) * 3 * m));
    PLO_CUDA(c, c->h_w.reserve(sizeof(double) * m));
    PLO_CUDA(c, c->blockcnt.reserve(sizeof(int) * ((m + kTile - 1) / kTile + 1)));
    PLO_CUDA(c, c->ransac_mind.reserve((sizeof(long long) + 16 * sizeof(double)) * (size_t)std::max(c->dprm.ransac_max_iterations, 1) + 64));   // RansacScratch
    PLO_CUDA(c, c->partials2.reserve(sizeof(double) * kNoiseSums * (size_t)plo_grid(c, 2)));
  }
  return PLO_OK;
}

int plo_launch_reduce_solve(plo_ctx* c, bool advance_loop, unsigned long long cond_handle) {
  const int g = reduce_grid(c, c->m_raw);
  PLO_TRY(plo_reserve_solver_buffers(c));
  const int adv = advance_loop ? 1 : 0;
  const cudaGraphConditionalHandle cond = (cudaGraphConditionalHandle)cond_handle;
  const int use_cond = cond_handle ? 1 : 0;
  const bool drpm_only = c->host_drpm_only;   // plo_solve_drpm_host: the DRPM tail alone, on caller-supplied weights
  const bool ransac = c->dprm.solver == PLO_SOLVER_RANSAC && c->m_raw > 0 && !drpm_only;
  const bool ransac_ls = ransac && c->dprm.ransac_final == PLO_FINAL_LS;   // trimmed LS on the inliers (src/solver.cpp:366-371)
  const bool trimmed = (c->dprm.solver == PLO_SOLVER_LS && c->m_raw > 0) || ransac_ls;
  const int in_unit = ransac_ls ? 1 : 0;
  DevParams P = c->dprm;
  if (trimmed) P.weight_mode = PLO_W_UNIT;       // SolveMotionEstimationProblemLS is unweighted
  if (drpm_only) { P.ext_weights = 1; P.solver = PLO_SOLVER_WLS; }
  const double* w_ext = c->host_w;
  if (ransac) P.weight_mode = PLO_W_HUBER_EXP;   // Huber/exp weights at the best hypothesis (src/solver.cpp:334-364);
                                                 // with final "LS" only their inlier test is used (in_unit)
  if (c->dprm.solver == PLO_SOLVER_RANSAC && !ransac) P.solver = PLO_SOLVER_WLS;   // empty source: nothing to sample
  if (ransac) {
    // the hypothesis sampler works on the compacted pair list (indices = positions in source_cloud)
    PLO_TRY(plo_launch_compact_pairs(c, c->h_src.as<float>(), c->h_ref.as<float>(), c->h_nrm.as<float>(), c->h_w.as<int32_t>()));
    const int maxit = P.ransac_max_iterations;
    RansacScratch rs;
    rs.cnt = c->ransac_mind.as<long long>();
    rs.T = reinterpret_cast<double*>(rs.cnt + maxit);
    rs.stop = reinterpret_cast<int*>(rs.T + (size_t)16 * maxit);
    k_ransac_eval<<<1, 1024, 0, c->stream>>>(c->h_src.as<float>(), c->h_ref.as<float>(), c->h_nrm.as<float>(), c->counts.as<DevCounts>(),
                                             c->state.as<DevState>(), P, rs, 0, maxit < 1 ? maxit : 1, adv);
    if (maxit > 1)
      k_ransac_eval<<<std::min(maxit - 1, 2 * c->sm_count), 1024, 0, c->stream>>>(c->h_src.as<float>(), c->h_ref.as<float>(), c->h_nrm.as<float>(),
                                                                                  c->counts.as<DevCounts>(), c->state.as<DevState>(), P, rs, 1, maxit, adv);
    k_ransac_pick<<<1, 32, 0, c->stream>>>(c->counts.as<DevCounts>(), c->state.as<DevState>(), P, rs, adv);
    c->launches += maxit > 1 ? 3 : 2;
    PLO_CUDA(c, cudaGetLastError());
  }
  if (advance_loop && c->tune_fuse && c->dprm.solver == PLO_SOLVER_WLS) {
    // resident weighted-LS loop: one launch (k_reduce_pairs and k_solve_update read the qx / qy / qn a projection wrote
    // in the same kernel; the projection itself may come through the read-only path here: another launch wrote it)
    k_reduce_solve<<<plo_loop_blocks(c), kFusedReduceWarps * 32, 0, c->stream>>>(c->q_x.as<float4>(), c->q_y.as<float4>(), c->q_n.as<float4>(),
                                                        c->counts.as<DevCounts>(), c->state.as<DevState>(), P,
                                                        c->partials.as<double>(), c->reduce_ticket.as<int>(), cond, use_cond);
    c->launches++;
    PLO_CUDA(c, cudaGetLastError());
    return PLO_OK;
  }
  if (c->m_raw > 0) {
    k_reduce_pairs<<<g, kReduceThreads, 0, c->stream>>>(c->q_x.as<float4>(), c->q_y.as<float4>(), c->q_n.as<float4>(),
                                                        c->counts.as<DevCounts>(), c->state.as<DevState>(), P,
                                                        c->partials.as<double>(), adv, nullptr, in_unit, w_ext);
    c->launches++;
    PLO_CUDA(c, cudaGetLastError());
  }
  const bool drpm = (ransac && c->dprm.ransac_final == PLO_FINAL_DRPM) || drpm_only;
  k_solve_update<<<1, 256, 0, c->stream>>>(c->partials.as<double>(), c->m_raw > 0 ? g : 0, c->state.as<DevState>(), P, adv, cond,
                                            use_cond, trimmed ? 1 : (drpm ? 3 : 0));
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  if (drpm) {
    k_drpm_eigen<<<1, 32, 0, c->stream>>>(c->state.as<DevState>(), P, adv);
    c->launches++;
    PLO_CUDA(c, cudaGetLastError());
    k_drpm_noise<<<g, kReduceThreads, 0, c->stream>>>(c->q_x.as<float4>(), c->q_y.as<float4>(), c->q_n.as<float4>(),
                                                      c->counts.as<DevCounts>(), c->state.as<DevState>(), P,
                                                      c->partials2.as<double>(), adv, w_ext);
    c->launches++;
    PLO_CUDA(c, cudaGetLastError());
    k_drpm_finish<<<1, 64, 0, c->stream>>>(c->partials2.as<double>(), g, c->state.as<DevState>(), P, adv, cond, use_cond);
    c->launches++;
    PLO_CUDA(c, cudaGetLastError());
    return PLO_OK;
  }
  if (!trimmed) return PLO_OK;
  // ---- trimmed LS: residual keys -> stable sort -> rank window -> second reduce + solve ----
  const int m = (int)c->m_raw;
  const int gk = (int)std::max<int64_t>(1, std::min<int64_t>((m + 255) / 256, (int64_t)plo_grid(c, 4)));
  k_ls_keys<<<gk, 256, 0, c->stream>>>(c->q_x.as<float4>(), c->q_y.as<float4>(), c->q_n.as<float4>(), c->counts.as<DevCounts>(),
                                       c->state.as<DevState>(), adv, m, c->ls_keys[0].as<unsigned long long>(),
                                       c->ls_vals[0].as<int>(), P, in_unit);
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  unsigned long long* keys[2] = {c->ls_keys[0].as<unsigned long long>(), c->ls_keys[1].as<unsigned long long>()};
  int* vals[2] = {c->ls_vals[0].as<int>(), c->ls_vals[1].as<int>()};
  int which = 0;
  PLO_TRY(plo_sort_pairs(c, keys, vals, m, kLsPasses, c->ls_hist.as<int>(), c->ls_tot.as<int>(), &which));
  k_ls_select<<<gk, 256, 0, c->stream>>>(vals[which], c->state.as<DevState>(), adv, m, c->dprm.ls_threshold, c->ls_mask.as<int>(),
                                         in_unit);
  P.weight_mode = PLO_W_UNIT;   // second pass: the mask is the selection
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  k_reduce_pairs<<<g, kReduceThreads, 0, c->stream>>>(c->q_x.as<float4>(), c->q_y.as<float4>(), c->q_n.as<float4>(),
                                                      c->counts.as<DevCounts>(), c->state.as<DevState>(), P,
                                                      c->partials.as<double>(), adv, c->ls_mask.as<int>(), 0, nullptr);
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  k_solve_update<<<1, 256, 0, c->stream>>>(c->partials.as<double>(), g, c->state.as<DevState>(), P, adv, cond, use_cond, 2);
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
  k_solve_update<<<1, 256, 0, c->stream>>>(c->partials.as<double>(), n > 0 ? g : 0, c->state.as<DevState>(), P, 0,
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
