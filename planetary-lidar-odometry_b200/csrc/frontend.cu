// frontend.cu — the front-end stage that produces the normals the matcher consumes:
// laserCloudHandler of src/scan_registration.cpp at config.json defaults — range gate (:862-863, :86-113),
// ring assignment (:938-1016), azimuth / relTime / intensity (:1018-1042), per-ring clouds (:1043, :1062-1069),
// windowed PCA normals over three rings (:158-229, :1162-1229) with the plane check (:137-156), and the
// planarity presample (:279-327, :1481-1489).  Output = filteredLaserCloud as PointXYZINormal records, resident
// on the device (it can feed plo_set_source_device / plo_map_push_device without a host round trip).
//
// Everything is float32 arithmetic in the reference; the third-party pieces that cannot be reproduced bit by bit
// (Eigen's vectorised reductions and SelfAdjointEigenSolver<Matrix3f>, FLANN's tie order, libm overload choice)
// are DEFINED as in oracle/plo_oracle_frontend.c: sums in row order without FMA, cyclic Jacobi in float, ties to
// the smaller index, angles through the double functions rounded to float.  Every kernel below mirrors that
// file operation by operation (the library is built with -fmad=false).
//
// Shape: ~12 small launches, no host synchronisation before the final count read-back.
//   k_fe_gate -> scan -> k_fe_compact      finite + range gate, order-preserving
//   k_fe_ring                              scanID, raw azimuth, first half-turn index (atomicMin)
//   k_fe_time_keys                         intensity; sort key = ring (dropped points -> sentinel)
//   one stable radix pass (index_build.cu) per-ring clouds in arrival order; ring sizes = digit totals
//   k_fe_ring_offsets, k_fe_gather         ring-ordered float4 (x, y, z, intensity)
//   k_fe_nn                                nearest point of the ring below / above: brute force over the ring from
//                                          shared-memory tiles (a ring has <= ~2100 points)
//   k_fe_pca                               one thread per point: 3 x (2w/step+1) rows, centroid, covariance, Jacobi,
//                                          plane check, normal, eigenvalues, planarity flag
//   count -> scan -> k_fe_emit             order-preserving compaction into 48-byte records
// Algorithmic bytes per input point: 12 read + 48 written + 3 x 7 x 12 gathered = 312 B; roofline: HBM (in practice
// launch-bound: 130 k points are 1.6 MB).
#include <math_constants.h>

#include <algorithm>
#include <cstddef>
#include <cstdint>

#include "plo_internal.cuh"
#include "plo_scan.cuh"

namespace {

constexpr int kMaxRings = 64;
constexpr unsigned long long kDropKey = 1023ull;   // digit of points without a ring: sorts behind every ring
constexpr double kPi = 3.14159265358979323846;

struct FeCounts {
  int m;            // points after the finite + range gate
  int first_half;   // first gated index whose azimuth passed the half turn (:1030-1033); INT_MAX if none
  int n_out;        // filteredLaserCloud size
  int total;        // points with a ring
  int st_fail, st_invalid, st_cand;
  int pad;
  int ring_off[kMaxRings + 1];
};

__device__ __forceinline__ bool fe_finite3(float x, float y, float z) { return isfinite(x) && isfinite(y) && isfinite(z); }

__global__ void k_fe_init(FeCounts* fc) {
  if (threadIdx.x == 0) { fc->m = 0; fc->first_half = 0x7fffffff; fc->n_out = 0; fc->total = 0; fc->st_fail = fc->st_invalid = fc->st_cand = 0; }
}

// :862-863 — removeNaNFromPointCloud + removeClosedPointCloud (:86-113)
__device__ __forceinline__ bool fe_gate(const char* rec, int stride, int i, float mn2, float mx2, float4& p) {
  const float* r = reinterpret_cast<const float*>(rec + (size_t)i * stride);
  p = make_float4(r[0], r[1], r[2], __int_as_float(i));
  if (!fe_finite3(p.x, p.y, p.z)) return false;
  const float d2 = p.x * p.x + p.y * p.y + p.z * p.z;
  return !(d2 < mn2 || d2 > mx2);
}

__global__ void __launch_bounds__(256) k_fe_gate(const char* __restrict__ rec, int stride, int n, float mn2, float mx2,
                                                 int* __restrict__ blockcnt) {
  __shared__ int s_c[8];
  const int base = blockIdx.x * kTile;
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < kTile / 256; ++j) {
    const int i = base + j * 256 + threadIdx.x;
    float4 p;
    cnt += (i < n && fe_gate(rec, stride, i, mn2, mx2, p)) ? 1 : 0;
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

__global__ void __launch_bounds__(256) k_fe_compact(const char* __restrict__ rec, int stride, int n, float mn2, float mx2,
                                                    const int* __restrict__ blockoff, float4* __restrict__ kp) {
  const int base = blockIdx.x * kTile;
  bool ok[kTile / 256];
  float4 p[kTile / 256];
#pragma unroll
  for (int j = 0; j < kTile / 256; ++j) {
    const int i = base + j * 256 + threadIdx.x;
    ok[j] = i < n && fe_gate(rec, stride, i, mn2, mx2, p[j]);
  }
  int rank[kTile / 256];
  tile_ranks(ok, rank);
  const int off = blockoff[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kTile / 256; ++j)
    if (ok[j]) kp[off + rank[j]] = p[j];
}

// :938-1016; -1 = dropped
__device__ __forceinline__ int fe_scan_id(float x, float y, float z, int n_scans) {
  const float range = (float)sqrt((double)(x * x + y * y));
  const float vertical_angle = (float)atan((double)(z / range));
  const float angle = (float)((double)vertical_angle * 180.0 / kPi);
  int id = 0;
  if (n_scans == 16) {
    id = (int)((double)((angle + 15.0f) / 2.0f) + 0.5);
    if (id > n_scans - 1 || id < 0) return -1;
  } else if (n_scans == 32) {
    const float tab[27] = {-25.000f, -15.639f, -11.310f, -8.843f, -7.254f, -6.148f, -5.333f, -4.667f, -4.000f,
                           -3.667f,  -3.333f,  -3.000f,  -2.667f, -2.333f, -2.000f, -1.667f, -1.333f, -1.000f,
                           -0.667f,  -0.333f,  0.000f,   0.333f,  0.667f,  1.000f,  1.333f,  1.667f,  2.333f};
    float min_diff = 3.402823466e+38f;
#pragma unroll
    for (int j = 0; j < 27; ++j) {
      const float diff = fabsf(angle - tab[j]);
      if (diff < min_diff) { min_diff = diff; id = j; }
    }
    if (id > n_scans - 1 || id < 0) return -1;
  } else {
    const float upper = 2.0f, lower = -24.33f;
    if ((double)angle >= -8.83) id = (int)((double)(upper - angle) * 3.0 + 0.5);
    else id = n_scans / 2 + (int)((-8.83 - (double)angle) * 2.0 + 0.5);
    if (angle > upper || angle < lower || id > 50 || id < 0) return -1;
  }
  return id;
}

__device__ __forceinline__ float fe_raw_ori(float x, float y) { return (float)(-atan2((double)y, (double)x)); }

// the not-yet-half-passed branch of :1020-1034: adjusted azimuth and whether this point trips the flag
__device__ __forceinline__ float fe_ori_first_half(float ori, float startOri, bool& trips) {
  if ((double)ori < (double)startOri - kPi / 2) ori = (float)((double)ori + 2 * kPi);
  else if ((double)ori > (double)startOri + kPi * 3 / 2) ori = (float)((double)ori - 2 * kPi);
  trips = (double)(ori - startOri) > kPi;
  return ori;
}

__global__ void __launch_bounds__(256) k_fe_ring(const float4* __restrict__ kp, FeCounts* __restrict__ fc, int n_scans,
                                                 int* __restrict__ ring) {
  const int m = fc->m;
  const int a = blockIdx.x * 256 + threadIdx.x;
  if (a >= m) return;
  const float4 p = kp[a];
  const int id = fe_scan_id(p.x, p.y, p.z, n_scans);
  ring[a] = id;
  if (id < 0) return;
  const float4 p0 = kp[0];
  bool trips;
  fe_ori_first_half(fe_raw_ori(p.x, p.y), fe_raw_ori(p0.x, p0.y), trips);
  if (trips) atomicMin(&fc->first_half, a);
}

// :1018-1042 — intensity = scanID + scanPeriod * relTime; sort key of the per-ring split
__global__ void __launch_bounds__(256) k_fe_time_keys(const float4* __restrict__ kp, const FeCounts* __restrict__ fc,
                                                      const int* __restrict__ ring, float scan_period, int n,
                                                      float* __restrict__ inten, unsigned long long* __restrict__ keys,
                                                      int* __restrict__ vals) {
  const int a = blockIdx.x * 256 + threadIdx.x;
  if (a >= n) return;
  const int m = fc->m;
  unsigned long long key = kDropKey;
  if (a < m) {
    const int id = ring[a];
    if (id >= 0) {
      const float4 p = kp[a], p0 = kp[0], pl = kp[m - 1];
      const float startOri = fe_raw_ori(p0.x, p0.y);
      float endOri = (float)((double)fe_raw_ori(pl.x, pl.y) + 2.0 * kPi);           // :901-903
      if ((double)(endOri - startOri) > 3.0 * kPi) endOri = (float)((double)endOri - 2.0 * kPi);   // :905-912
      else if ((double)(endOri - startOri) < kPi) endOri = (float)((double)endOri + 2.0 * kPi);
      float ori = fe_raw_ori(p.x, p.y);
      if (a <= fc->first_half) {
        bool trips;
        ori = fe_ori_first_half(ori, startOri, trips);
      } else {   // :1035-1046
        ori = (float)((double)ori + 2 * kPi);
        if ((double)ori < (double)endOri - kPi * 3 / 2) ori = (float)((double)ori + 2 * kPi);
        else if ((double)ori > (double)endOri + kPi / 2) ori = (float)((double)ori - 2 * kPi);
      }
      const float relTime = (ori - startOri) / (endOri - startOri);
      inten[a] = (float)id + scan_period * relTime;
      key = (unsigned long long)id;
    }
  }
  keys[a] = key;
  vals[a] = a;
}

// ring sizes are the digit totals of the radix pass
__global__ void k_fe_ring_offsets(const int* __restrict__ digit_total, int n_scans, FeCounts* __restrict__ fc) {
  if (threadIdx.x != 0) return;
  int run = 0;
  for (int r = 0; r < kMaxRings; ++r) {
    fc->ring_off[r] = run;
    if (r < n_scans) run += digit_total[r];
  }
  fc->ring_off[kMaxRings] = run;
  fc->total = run;
}

// laserCloudScans concatenated (:1062-1069): x, y, z, intensity + index into the caller's cloud
__global__ void __launch_bounds__(256) k_fe_gather(const int* __restrict__ vals_sorted, const float4* __restrict__ kp,
                                                   const float* __restrict__ inten, const FeCounts* __restrict__ fc,
                                                   float4* __restrict__ rp, int* __restrict__ rsrc) {
  const int o = blockIdx.x * 256 + threadIdx.x;
  if (o >= fc->total) return;
  const int a = vals_sorted[o];
  const float4 p = kp[a];
  rp[o] = make_float4(p.x, p.y, p.z, inten[a]);
  rsrc[o] = __float_as_int(p.w);
}

// findNearestPoint (:115-135): FLANN 1-NN in the ring below (s = 0) and above (s = 1); squared float distance
// ((dx*dx + dy*dy) + dz*dz), ties to the smaller index, accepted iff < knn_distance_threshold
constexpr int kNnTile = 1024;
__global__ void __launch_bounds__(256) k_fe_nn(const float4* __restrict__ rp, const FeCounts* __restrict__ fc, int n_scans,
                                               float thr, int* __restrict__ nn_below, int* __restrict__ nn_above) {
  __shared__ float4 s_p[kNnTile];
  const int i = blockIdx.y;   // ring
  if (i < 1 || i > n_scans - 2) return;
  const int o0 = fc->ring_off[i], cnt = fc->ring_off[i + 1] - o0;
  for (int s = 0; s < 2; ++s) {
    const int r = s == 0 ? i - 1 : i + 1;
    const int b0 = fc->ring_off[r], bn = fc->ring_off[r + 1] - b0;
    int* out = s == 0 ? nn_below : nn_above;
    for (int q0 = blockIdx.x * 256; q0 < cnt; q0 += gridDim.x * 256) {
      const int j = q0 + threadIdx.x;
      const bool live = j < cnt;
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) q = rp[o0 + j];
      int best = -1;
      float bd = 3.402823466e+38f;
      for (int t0 = 0; t0 < bn; t0 += kNnTile) {
        __syncthreads();
        for (int t = threadIdx.x; t < kNnTile && t0 + t < bn; t += 256) s_p[t] = rp[b0 + t0 + t];
        __syncthreads();
        const int lim = min(kNnTile, bn - t0);
        if (live) {
#pragma unroll 4
          for (int t = 0; t < lim; ++t) {
            const float4 c = s_p[t];
            const float dx = q.x - c.x, dy = q.y - c.y, dz = q.z - c.z;
            const float d = (dx * dx + dy * dy) + dz * dz;
            if (d < bd) { bd = d; best = t0 + t; }
          }
        }
      }
      if (live) out[o0 + j] = (best >= 0 && bd < thr) ? best : -1;
    }
  }
}

// cyclic Jacobi, float — the same statements as sym3_eigen_f of oracle/plo_oracle_frontend.c
__device__ void fe_sym3_eigen(float a00, float a01, float a02, float a11, float a12, float a22, float ev[3], float Vout[9]) {
  float A[3][3] = {{a00, a01, a02}, {a01, a11, a12}, {a02, a12, a22}};
  float V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 24; ++sweep) {
    const float off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    if (off == 0.0f) break;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      const float apq = A[p][q];
      if (apq == 0.0f) continue;
      const float theta = (A[q][q] - A[p][p]) / (2.0f * apq);
      const float t = (theta >= 0.0f ? 1.0f : -1.0f) / (fabsf(theta) + sqrtf(theta * theta + 1.0f));
      const float cs = 1.0f / sqrtf(t * t + 1.0f), sn = t * cs;
#pragma unroll
      for (int k = 0; k < 3; ++k) { const float akp = A[k][p], akq = A[k][q]; A[k][p] = cs * akp - sn * akq; A[k][q] = sn * akp + cs * akq; }
#pragma unroll
      for (int k = 0; k < 3; ++k) { const float apk = A[p][k], aqk = A[q][k]; A[p][k] = cs * apk - sn * aqk; A[q][k] = sn * apk + cs * aqk; }
#pragma unroll
      for (int k = 0; k < 3; ++k) { const float vkp = V[k][p], vkq = V[k][q]; V[k][p] = cs * vkp - sn * vkq; V[k][q] = sn * vkp + cs * vkq; }
    }
  }
  const float d[3] = {A[0][0], A[1][1], A[2][2]};
  int o0 = 0, o1 = 1, o2 = 2;   // bubble sort of three, ascending, stable (as the oracle's)
  if (d[o1] < d[o0]) { const int t = o0; o0 = o1; o1 = t; }
  if (d[o2] < d[o1]) { const int t = o1; o1 = o2; o2 = t; }
  if (d[o1] < d[o0]) { const int t = o0; o0 = o1; o1 = t; }
  const int ord[3] = {o0, o1, o2};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    ev[c] = d[ord[c]];
#pragma unroll
    for (int r = 0; r < 3; ++r) Vout[r * 3 + c] = V[r][ord[c]];
  }
}

struct FeParams {
  int n_scans, window, step, use_all_points;
  float plane_thr, valid_frac, planarity_thr;
};

// row t of computeNormalPCA's point matrix (:166-199): own ring, ring below, ring above
__device__ __forceinline__ float4 fe_row(const float4* __restrict__ rp, int t, int per_ring, int w, int step, int own0, int j,
                                         int below0, int nb, int above0, int na) {
  const int s = t / per_ring, d = -w + (t - s * per_ring) * step;
  const int base = s == 0 ? own0 + j : (s == 1 ? below0 + nb : above0 + na);
  return rp[base + d];
}

// status: 0 = not in the output, 1 = in the output (plane check passed), 2 = in the output, plane check failed
__global__ void __launch_bounds__(128) k_fe_pca(const float4* __restrict__ rp, const int* __restrict__ nn_below,
                                                const int* __restrict__ nn_above, FeCounts* __restrict__ fc, FeParams P,
                                                int* __restrict__ status, float4* __restrict__ nrm_out, float4* __restrict__ ev_out) {
  __shared__ int s_off[kMaxRings + 1];
  for (int r = threadIdx.x; r <= kMaxRings; r += blockDim.x) s_off[r] = fc->ring_off[r];
  __syncthreads();
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= s_off[kMaxRings]) return;
  int i = 0;
  while (i + 1 < kMaxRings && s_off[i + 1] <= o) ++i;
  const int j = o - s_off[i];
  int st = 0;
  float4 nout = make_float4(0.f, 0.f, 0.f, 0.f), eout = make_float4(0.f, 0.f, 0.f, 0.f);
  const int cnt = s_off[i + 1] - s_off[i];
  bool eligible = i >= 1 && i <= P.n_scans - 2;   // :1163
  if (eligible) {
    const int cb = s_off[i] - s_off[i - 1], ca = s_off[i + 2] - s_off[i + 1];
    eligible = !(cnt - 11 < 6 || cb - 11 < 6 || ca - 11 < 6) && j >= 5 && j < cnt - 5;   // :1167-1171
    if (eligible) {
      const int w = P.window, step = P.step;
      const int per_ring = (2 * w) / step + 1, num = 3 * per_ring;
      // :166-199 — rows in range; any missing row is a failure (count < num, :181)
      const int nb = nn_below[o], na = nn_above[o];
      bool full = nb >= 0 && na >= 0;
      if (full) {
        for (int d = -w; d <= w; d += step)
          full = full && (j + d >= 0 && j + d < cnt) && (nb + d >= 0 && nb + d < cb) && (na + d >= 0 && na + d < ca);
      }
      if (!full) {
        atomicAdd(&fc->st_fail, 1);
      } else {
        const int own0 = s_off[i], below0 = s_off[i - 1], above0 = s_off[i + 1];
        float cx = 0.f, cy = 0.f, cz = 0.f;
        for (int t = 0; t < num; ++t) {
          const float4 r = fe_row(rp, t, per_ring, w, step, own0, j, below0, nb, above0, na);
          cx += r.x; cy += r.y; cz += r.z;
        }
        cx /= (float)num; cy /= (float)num; cz /= (float)num;
        float c00 = 0.f, c01 = 0.f, c02 = 0.f, c11 = 0.f, c12 = 0.f, c22 = 0.f;
        for (int t = 0; t < num; ++t) {
          const float4 r = fe_row(rp, t, per_ring, w, step, own0, j, below0, nb, above0, na);
          const float dx = r.x - cx, dy = r.y - cy, dz = r.z - cz;
          c00 += dx * dx; c01 += dx * dy; c02 += dx * dz; c11 += dy * dy; c12 += dy * dz; c22 += dz * dz;
        }
        const float inv = (float)(num - 1);
        float ev[3], V[9];
        fe_sym3_eigen(c00 / inv, c01 / inv, c02 / inv, c11 / inv, c12 / inv, c22 / inv, ev, V);
        int valid = 0;   // :137-156
        for (int t = 0; t < num; ++t) {
          const float4 r = fe_row(rp, t, per_ring, w, step, own0, j, below0, nb, above0, na);
          const float dist = fabsf((V[0] * (r.x - cx) + V[3] * (r.y - cy)) + V[6] * (r.z - cz));
          if (dist < P.plane_thr) valid++;
        }
        const bool plane_ok = (float)valid >= P.valid_frac * (float)num;
        float l1, l2, l3, nx, ny, nz;
        bool emit = true;
        if (plane_ok) { l1 = ev[2]; l2 = ev[1]; l3 = ev[0]; nx = V[0]; ny = V[3]; nz = V[6]; }
        else {
          l1 = l2 = l3 = -1.f;
          nx = V[2]; ny = V[5]; nz = V[8];   // un-swapped column 2 (:1198 after the early return of :214-218)
          emit = P.use_all_points != 0;      // :1186-1195
          if (emit) atomicAdd(&fc->st_invalid, 1);
        }
        if (emit) {
          const float nn = sqrtf((nx * nx + ny * ny) + nz * nz);
          if (nn > 0.f) { nx /= nn; ny /= nn; nz /= nn; }
          if (nz < 0.f) { nx = -nx; ny = -ny; nz = -nz; }   // :1201-1203
          const float planarity = (l2 - l3) / l1;           // :302
          const bool cand = plane_ok && planarity > P.planarity_thr;   // :322-326, :1481-1489
          if (cand) atomicAdd(&fc->st_cand, 1);
          st = plane_ok ? 1 : 2;
          nout = make_float4(nx, ny, nz, cand ? 1.f : 0.f);
          eout = make_float4(l1, l2, l3, 0.f);
        }
      }
    }
  }
  status[o] = st;
  nrm_out[o] = nout;
  ev_out[o] = eout;
}

__global__ void __launch_bounds__(256) k_fe_count_out(const int* __restrict__ status, const FeCounts* __restrict__ fc,
                                                      int* __restrict__ blockcnt) {
  __shared__ int s_c[8];
  const int total = fc->total;
  const int base = blockIdx.x * kTile;
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < kTile / 256; ++j) {
    const int o = base + j * 256 + threadIdx.x;
    cnt += (o < total && status[o] != 0) ? 1 : 0;
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

// filteredLaserCloud (:1205-1227): PointXYZINormal = x y z 1 | nx ny nz 0 | intensity curvature 0 0
__global__ void __launch_bounds__(256) k_fe_emit(const int* __restrict__ status, const float4* __restrict__ rp,
                                                 const int* __restrict__ rsrc, const float4* __restrict__ nrm,
                                                 const float4* __restrict__ evs, const FeCounts* __restrict__ fc,
                                                 const int* __restrict__ blockoff, float4* __restrict__ rec, float* __restrict__ ev3,
                                                 unsigned char* __restrict__ cand, int* __restrict__ src_index) {
  const int total = fc->total;
  const int base = blockIdx.x * kTile;
  bool ok[kTile / 256];
#pragma unroll
  for (int j = 0; j < kTile / 256; ++j) {
    const int o = base + j * 256 + threadIdx.x;
    ok[j] = o < total && status[o] != 0;
  }
  int rank[kTile / 256];
  tile_ranks(ok, rank);
  const int off = blockoff[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kTile / 256; ++j) {
    if (!ok[j]) continue;
    const int o = base + j * 256 + threadIdx.x;
    const int d = off + rank[j];
    const float4 p = rp[o], n = nrm[o], e = evs[o];
    rec[3 * (size_t)d] = make_float4(p.x, p.y, p.z, 1.0f);
    rec[3 * (size_t)d + 1] = make_float4(n.x, n.y, n.z, 0.f);
    rec[3 * (size_t)d + 2] = make_float4(p.w, 0.f, 0.f, 0.f);
    ev3[3 * (size_t)d] = e.x; ev3[3 * (size_t)d + 1] = e.y; ev3[3 * (size_t)d + 2] = e.z;
    cand[d] = n.w != 0.f ? 1 : 0;
    src_index[d] = rsrc[o];
  }
}

}  // namespace

#define FE_LAUNCH_CHECK(c)             \
  do {                                 \
    (c)->launches++;                   \
    PLO_CUDA((c), cudaGetLastError()); \
  } while (0)

int plo_frontend_run(plo_ctx* c, const void* dev_records, int64_t n, int32_t stride, const plo_frontend_params* fp) {
  c->fe_n_in = n;
  c->fe_valid = false;
  PLO_CUDA(c, c->fe_counts.reserve(sizeof(FeCounts)));
  FeCounts* fc = c->fe_counts.as<FeCounts>();
  cudaStream_t s = c->stream;
  k_fe_init<<<1, 32, 0, s>>>(fc);
  FE_LAUNCH_CHECK(c);
  if (n == 0) { c->fe_valid = true; return PLO_OK; }
  if (n > (int64_t)1 << 28) return plo_fail(c, PLO_ERR_UNSUPPORTED, "plo_frontend: cloud larger than 2^28 points");
  const int nb = (int)((n + kTile - 1) / kTile);
  const int nt = (int)((n + 255) / 256);
  PLO_CUDA(c, c->fe_blockcnt.reserve(sizeof(int) * (size_t)(nb + 1)));
  PLO_CUDA(c, c->fe_kp.reserve(sizeof(float4) * n));
  PLO_CUDA(c, c->fe_ring.reserve(sizeof(int) * n));
  PLO_CUDA(c, c->fe_inten.reserve(sizeof(float) * n));
  PLO_CUDA(c, c->fe_rp.reserve(sizeof(float4) * n));
  PLO_CUDA(c, c->fe_rsrc.reserve(sizeof(int) * n));
  PLO_CUDA(c, c->fe_nn[0].reserve(sizeof(int) * n));
  PLO_CUDA(c, c->fe_nn[1].reserve(sizeof(int) * n));
  PLO_CUDA(c, c->fe_status.reserve(sizeof(int) * n));
  PLO_CUDA(c, c->fe_nrm.reserve(sizeof(float4) * n));
  PLO_CUDA(c, c->fe_ev.reserve(sizeof(float4) * n));
  PLO_CUDA(c, c->fe_rec.reserve(sizeof(float4) * 3 * n));
  PLO_CUDA(c, c->fe_ev3.reserve(sizeof(float) * 3 * n));
  PLO_CUDA(c, c->fe_cand.reserve((size_t)n));
  PLO_CUDA(c, c->fe_src.reserve(sizeof(int) * n));
  for (int a = 0; a < 2; ++a) {
    PLO_CUDA(c, c->fe_keys[a].reserve(sizeof(unsigned long long) * n));
    PLO_CUDA(c, c->fe_vals[a].reserve(sizeof(int) * n));
  }
  PLO_CUDA(c, c->fe_hist.reserve(sizeof(int) * plo_sort_hist_ints(n)));
  PLO_CUDA(c, c->fe_tot.reserve(sizeof(int) * plo_sort_total_ints(1)));
  const float mn2 = fp->min_range * fp->min_range, mx2 = fp->max_range * fp->max_range;
  const char* rec = static_cast<const char*>(dev_records);
  k_fe_gate<<<nb, 256, 0, s>>>(rec, stride, (int)n, mn2, mx2, c->fe_blockcnt.as<int>());
  FE_LAUNCH_CHECK(c);
  k_scan_exclusive<<<1, 1024, 0, s>>>(c->fe_blockcnt.as<int>(), nb, &fc->m);
  FE_LAUNCH_CHECK(c);
  k_fe_compact<<<nb, 256, 0, s>>>(rec, stride, (int)n, mn2, mx2, c->fe_blockcnt.as<int>(), c->fe_kp.as<float4>());
  FE_LAUNCH_CHECK(c);
  k_fe_ring<<<nt, 256, 0, s>>>(c->fe_kp.as<float4>(), fc, fp->n_scans, c->fe_ring.as<int>());
  FE_LAUNCH_CHECK(c);
  k_fe_time_keys<<<nt, 256, 0, s>>>(c->fe_kp.as<float4>(), fc, c->fe_ring.as<int>(), fp->scan_period, (int)n,
                                     c->fe_inten.as<float>(), c->fe_keys[0].as<unsigned long long>(), c->fe_vals[0].as<int>());
  FE_LAUNCH_CHECK(c);
  unsigned long long* kk[2] = {c->fe_keys[0].as<unsigned long long>(), c->fe_keys[1].as<unsigned long long>()};
  int* vv[2] = {c->fe_vals[0].as<int>(), c->fe_vals[1].as<int>()};
  int which = 0;
  PLO_TRY(plo_sort_pairs(c, kk, vv, n, 1, c->fe_hist.as<int>(), c->fe_tot.as<int>(), &which, 0));
  k_fe_ring_offsets<<<1, 32, 0, s>>>(c->fe_tot.as<int>(), fp->n_scans, fc);
  FE_LAUNCH_CHECK(c);
  k_fe_gather<<<nt, 256, 0, s>>>(vv[which], c->fe_kp.as<float4>(), c->fe_inten.as<float>(), fc, c->fe_rp.as<float4>(),
                                  c->fe_rsrc.as<int>());
  FE_LAUNCH_CHECK(c);
  {
    const dim3 grid((unsigned)std::max(1, std::min(nt, 16)), (unsigned)fp->n_scans);
    k_fe_nn<<<grid, 256, 0, s>>>(c->fe_rp.as<float4>(), fc, fp->n_scans, fp->knn_distance_threshold, c->fe_nn[0].as<int>(),
                                 c->fe_nn[1].as<int>());
    FE_LAUNCH_CHECK(c);
  }
  FeParams P;
  P.n_scans = fp->n_scans; P.window = fp->window_size; P.step = fp->iter_step; P.use_all_points = fp->use_all_points;
  P.plane_thr = fp->plane_distance_threshold; P.valid_frac = fp->valid_points_threshold; P.planarity_thr = fp->planarity_threshold;
  k_fe_pca<<<(int)((n + 127) / 128), 128, 0, s>>>(c->fe_rp.as<float4>(), c->fe_nn[0].as<int>(), c->fe_nn[1].as<int>(), fc, P,
                                                   c->fe_status.as<int>(), c->fe_nrm.as<float4>(), c->fe_ev.as<float4>());
  FE_LAUNCH_CHECK(c);
  k_fe_count_out<<<nb, 256, 0, s>>>(c->fe_status.as<int>(), fc, c->fe_blockcnt.as<int>());
  FE_LAUNCH_CHECK(c);
  k_scan_exclusive<<<1, 1024, 0, s>>>(c->fe_blockcnt.as<int>(), nb, &fc->n_out);
  FE_LAUNCH_CHECK(c);
  k_fe_emit<<<nb, 256, 0, s>>>(c->fe_status.as<int>(), c->fe_rp.as<float4>(), c->fe_rsrc.as<int>(), c->fe_nrm.as<float4>(),
                                c->fe_ev.as<float4>(), fc, c->fe_blockcnt.as<int>(), c->fe_rec.as<float4>(), c->fe_ev3.as<float>(),
                                c->fe_cand.as<unsigned char>(), c->fe_src.as<int>());
  FE_LAUNCH_CHECK(c);
  c->fe_valid = true;
  return PLO_OK;
}

// one synchronisation: counts of the last plo_frontend_run
int plo_frontend_fetch_counts(plo_ctx* c, int64_t out7[7]) {
  FeCounts h;
  PLO_CUDA(c, cudaMemcpyAsync(&h, c->fe_counts.p, offsetof(FeCounts, ring_off), cudaMemcpyDeviceToHost, c->stream));
  PLO_CUDA(c, cudaStreamSynchronize(c->stream));
  out7[0] = h.n_out; out7[1] = h.m; out7[2] = h.total; out7[3] = h.st_fail; out7[4] = h.st_invalid; out7[5] = h.st_cand;
  out7[6] = h.first_half;
  return PLO_OK;
}
