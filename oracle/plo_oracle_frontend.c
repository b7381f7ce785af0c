/* plo_oracle_frontend.c — CPU restatement of the reference FRONT-END stage that produces the normals the
 * matcher consumes: laserCloudHandler of src/scan_registration.cpp, default
 * configuration (format "pointcloud", method "pca", neighbor_scan "kdtree", presample
 * "geometric_features").  TEST INFRASTRUCTURE ONLY, like the rest of oracle/.  PARITY UNPINNED: the
 * reference cannot be built here (PCL, FLANN, Eigen, ROS absent) and ships no fixtures.
 *
 * Restated steps (file:line into src/scan_registration.cpp):
 *   :862-863  removeNaNFromPointCloud + removeClosedPointCloud (:86-113): finite xyz, min^2 <= |p|^2 <= max^2 (float)
 *   :900-914  startOri / endOri from the first / last kept point
 *   :938-1016 scanID from the vertical angle (16 / 32 / 64 rings), points outside the ring range dropped
 *   :1018-1042 azimuth unwrapping with the halfPassed flag, relTime, intensity = scanID + scanPeriod * relTime
 *   :1043, :1062-1069 per-ring clouds in arrival order, concatenated ring by ring
 *   :1162-1229 the PCA pass: rings 1 .. N-2 whose three rings have >= 17 points, points j in [5, size-5)
 *   :158-229  computeNormalPCA: 2w/step+1 points of the own ring around j, the same window around the nearest point
 *             (FLANN 1-NN, squared float distance < knn_distance_threshold, :115-135) of the ring below and of the ring
 *             above; fewer than 3*(2w/step+1) points => failure (point skipped); float32 centroid, covariance / (n-1),
 *             eigen-decomposition; :137-156 plane check (>= valid_points_threshold * n points within
 *             distance_threshold of the plane) else lambda = -1
 *   :1196-1227 output point: xyz, intensity, normal = unit eigenvector of the smallest eigenvalue flipped to +z —
 *             for a point that FAILED the plane check and is kept (use_all_points) the reference reads column 2 of
 *             the UN-swapped eigenvector matrix, i.e. the LARGEST eigenvalue's vector; restated as is
 *   :279-327  computeGeometricFeatures: planarity = (l2 - l3) / l1 > planarity_threshold => presample candidate;
 *             :1481-1489 plane-check failures are removed from the candidates
 *
 * Third-party arithmetic that is not reproducible bit by bit (absent libraries, vectorised reductions): Eigen's
 * colwise().mean(), the covariance product and SelfAdjointEigenSolver<Matrix3f>, FLANN's traversal order on exact
 * ties, libm's float/double overload choice for sqrt / atan / atan2.  Defined here: sums in row order, float
 * arithmetic without FMA; cyclic Jacobi in float; ties to the smaller index; angles through the double libm
 * functions, rounded to float. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "plo_oracle.h"

static const float kVlp32Angles[27] = {-25.000f, -15.639f, -11.310f, -8.843f, -7.254f, -6.148f, -5.333f, -4.667f, -4.000f,
                                       -3.667f,  -3.333f,  -3.000f,  -2.667f, -2.333f, -2.000f, -1.667f, -1.333f, -1.000f,
                                       -0.667f,  -0.333f,  0.000f,   0.333f,  0.667f,  1.000f,  1.333f,  1.667f,  2.333f};

/* :938-1016; returns -1 when the point is dropped */
static int scan_id(float x, float y, float z, int n_scans) {
  const float range = (float)sqrt((double)(x * x + y * y));
  const float vertical_angle = (float)atan((double)(z / range));
  const float angle = (float)((double)vertical_angle * 180.0 / M_PI);
  int id = 0;
  if (n_scans == 16) {
    id = (int)((double)((angle + 15.0f) / 2.0f) + 0.5);
    if (id > n_scans - 1 || id < 0) return -1;
  } else if (n_scans == 32) {
    float min_diff = 3.402823466e+38f;
    for (int j = 0; j < 27; ++j) {
      const float diff = fabsf(angle - kVlp32Angles[j]);
      if (diff < min_diff) { min_diff = diff; id = j; }
    }
    if (id > n_scans - 1 || id < 0) return -1;
  } else if (n_scans == 64) {
    const float upper = 2.0f, lower = -24.33f;
    if ((double)angle >= -8.83) id = (int)((double)(upper - angle) * 3.0 + 0.5);
    else id = n_scans / 2 + (int)((-8.83 - (double)angle) * 2.0 + 0.5);
    if (angle > upper || angle < lower || id > 50 || id < 0) return -1;
  } else {
    return -1;
  }
  return id;
}

/* cyclic Jacobi, float: eigenvalues ascending in ev[], unit eigenvectors in the columns of V (row-major 3x3) */
static void sym3_eigen_f(float a00, float a01, float a02, float a11, float a12, float a22, float ev[3], float Vout[9]) {
  float A[3][3] = {{a00, a01, a02}, {a01, a11, a12}, {a02, a12, a22}};
  float V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 24; ++sweep) {
    const float off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    if (off == 0.0f) break;
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      const float apq = A[p][q];
      if (apq == 0.0f) continue;
      const float theta = (A[q][q] - A[p][p]) / (2.0f * apq);
      const float t = (theta >= 0.0f ? 1.0f : -1.0f) / (fabsf(theta) + sqrtf(theta * theta + 1.0f));
      const float cs = 1.0f / sqrtf(t * t + 1.0f), sn = t * cs;
      for (int k = 0; k < 3; ++k) { const float akp = A[k][p], akq = A[k][q]; A[k][p] = cs * akp - sn * akq; A[k][q] = sn * akp + cs * akq; }
      for (int k = 0; k < 3; ++k) { const float apk = A[p][k], aqk = A[q][k]; A[p][k] = cs * apk - sn * aqk; A[q][k] = sn * apk + cs * aqk; }
      for (int k = 0; k < 3; ++k) { const float vkp = V[k][p], vkq = V[k][q]; V[k][p] = cs * vkp - sn * vkq; V[k][q] = sn * vkp + cs * vkq; }
    }
  }
  int o[3] = {0, 1, 2};
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2 - a; ++b)
      if (A[o[b + 1]][o[b + 1]] < A[o[b]][o[b]]) { const int t = o[b]; o[b] = o[b + 1]; o[b + 1] = t; }
  for (int c = 0; c < 3; ++c) {
    ev[c] = A[o[c]][o[c]];
    for (int r = 0; r < 3; ++r) Vout[r * 3 + c] = V[r][o[c]];
  }
}

void orc_frontend_default_params(orc_frontend_params* p) {
  memset(p, 0, sizeof(*p));
  p->n_scans = 64;
  p->min_range = 0.5f;          /* MINIMUM_RANGE, :62 */
  p->max_range = 120.0f;        /* MAXIMUM_RANGE, :63 */
  p->scan_period = 0.1f;        /* :55 */
  p->window_size = 3;           /* config.json:9 */
  p->iter_step = 1;             /* :10 */
  p->knn_distance_threshold = 10.0f; /* :11 */
  p->plane_distance_threshold = 0.02f;  /* :14 */
  p->valid_points_threshold = 0.8f;     /* :15 */
  p->use_all_points = 1;        /* :78 */
  p->planarity_threshold = 0.05f; /* :39 */
}

typedef struct { const float* p; int n; } ring_t; /* p: xyz triples */

/* FLANN 1-NN restated: smallest ((dx*dx + dy*dy) + dz*dz) in float, ties to the smaller index (:115-135) */
static int nearest_in_ring(const ring_t* r, const float q[3], float thr, int* out) {
  int best = -1;
  float bd = 3.402823466e+38f;
  for (int i = 0; i < r->n; ++i) {
    const float dx = q[0] - r->p[3 * i], dy = q[1] - r->p[3 * i + 1], dz = q[2] - r->p[3 * i + 2];
    const float d = (dx * dx + dy * dy) + dz * dz;
    if (d < bd) { bd = d; best = i; }
  }
  if (best >= 0 && bd < thr) { *out = best; return 1; }
  return 0;
}

int64_t orc_frontend(const void* pts, int64_t n, int32_t stride, const orc_frontend_params* P, float* out_records12,
                     float* out_eigenvalues, uint8_t* out_candidate, int32_t* out_src_index, int64_t* stats4) {
  const char* base = (const char*)pts;
  const int NS = P->n_scans;
  /* ---- :862-863 ---- */
  int32_t* kept = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n ? n : 1));
  int64_t m = 0;
  const float mn2 = P->min_range * P->min_range, mx2 = P->max_range * P->max_range;
  for (int64_t i = 0; i < n; ++i) {
    const float* p = (const float*)(base + i * stride);
    if (!(isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]))) continue;
    const float d2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
    if (d2 < mn2 || d2 > mx2) continue;
    kept[m++] = (int32_t)i;
  }
  int64_t n_out = 0;
  int64_t st_fail = 0, st_invalid = 0, st_cand = 0, st_ringed = 0;
  if (m > 0 && (NS == 16 || NS == 32 || NS == 64)) {
    /* ---- :900-914 ---- */
    const float* pf = (const float*)(base + (int64_t)kept[0] * stride);
    const float* pl = (const float*)(base + (int64_t)kept[m - 1] * stride);
    const float startOri = (float)(-atan2((double)pf[1], (double)pf[0]));
    float endOri = (float)((double)(float)(-atan2((double)pl[1], (double)pl[0])) + 2.0 * M_PI);
    if ((double)(endOri - startOri) > 3.0 * M_PI) endOri = (float)((double)endOri - 2.0 * M_PI);
    else if ((double)(endOri - startOri) < M_PI) endOri = (float)((double)endOri + 2.0 * M_PI);
    /* ---- :938-1043 ---- */
    int32_t* ring = (int32_t*)malloc(sizeof(int32_t) * (size_t)m);
    float* inten = (float*)malloc(sizeof(float) * (size_t)m);
    int* cnt = (int*)calloc((size_t)NS, sizeof(int));
    int halfPassed = 0;
    for (int64_t a = 0; a < m; ++a) {
      const float* p = (const float*)(base + (int64_t)kept[a] * stride);
      const int id = scan_id(p[0], p[1], p[2], NS);
      ring[a] = id;
      inten[a] = 0.f;
      if (id < 0) continue;
      float ori = (float)(-atan2((double)p[1], (double)p[0]));
      if (!halfPassed) {
        if ((double)ori < (double)startOri - M_PI / 2) ori = (float)((double)ori + 2 * M_PI);
        else if ((double)ori > (double)startOri + M_PI * 3 / 2) ori = (float)((double)ori - 2 * M_PI);
        if ((double)(ori - startOri) > M_PI) halfPassed = 1;
      } else {
        ori = (float)((double)ori + 2 * M_PI);
        if ((double)ori < (double)endOri - M_PI * 3 / 2) ori = (float)((double)ori + 2 * M_PI);
        else if ((double)ori > (double)endOri + M_PI / 2) ori = (float)((double)ori - 2 * M_PI);
      }
      const float relTime = (ori - startOri) / (endOri - startOri);
      inten[a] = (float)id + P->scan_period * relTime;
      cnt[id]++;
      st_ringed++;
    }
    /* ---- per-ring clouds (:1043) ---- */
    int* off = (int*)malloc(sizeof(int) * (size_t)(NS + 1));
    off[0] = 0;
    for (int r = 0; r < NS; ++r) off[r + 1] = off[r] + cnt[r];
    const int total = off[NS];
    float* rp = (float*)malloc(sizeof(float) * 3 * (size_t)(total ? total : 1));
    float* ri = (float*)malloc(sizeof(float) * (size_t)(total ? total : 1));
    int32_t* rs = (int32_t*)malloc(sizeof(int32_t) * (size_t)(total ? total : 1));
    int* fill = (int*)calloc((size_t)NS, sizeof(int));
    for (int64_t a = 0; a < m; ++a) {
      const int id = ring[a];
      if (id < 0) continue;
      const float* p = (const float*)(base + (int64_t)kept[a] * stride);
      const int o = off[id] + fill[id]++;
      rp[3 * o] = p[0]; rp[3 * o + 1] = p[1]; rp[3 * o + 2] = p[2];
      ri[o] = inten[a];
      rs[o] = kept[a];
    }
    /* ---- the PCA pass (:1162-1229) ---- */
    const int w = P->window_size, step = P->iter_step > 0 ? P->iter_step : 1;
    const int per_ring = (int)(2 * w / step) + 1;
    const int num = 3 * per_ring;
    for (int i = 1; i < NS - 1; ++i) {
      if (cnt[i] == 0) continue;
      if (cnt[i] - 11 < 6 || cnt[i - 1] - 11 < 6 || cnt[i + 1] - 11 < 6) continue; /* scanEndInd - scanStartInd < 6 */
      const ring_t own = {rp + 3 * off[i], cnt[i]}, below = {rp + 3 * off[i - 1], cnt[i - 1]}, above = {rp + 3 * off[i + 1], cnt[i + 1]};
      for (int j = 5; j < cnt[i] - 5; ++j) {
        /* gather in the reference's row order: own ring, ring i-1, ring i+1 */
        const float* rows[192];
        int count = 0;
        for (int d = -w; d <= w; d += step)
          if (j + d >= 0 && j + d < own.n) rows[count++] = own.p + 3 * (j + d);
        const ring_t* nb[2] = {&below, &above};
        for (int s = 0; s < 2; ++s) {
          int ni = j;
          if (nearest_in_ring(nb[s], own.p + 3 * j, P->knn_distance_threshold, &ni))
            for (int d = -w; d <= w; d += step)
              if (ni + d >= 0 && ni + d < nb[s]->n) rows[count++] = nb[s]->p + 3 * (ni + d);
        }
        if (count < num) { st_fail++; continue; } /* :181-184, :1178-1182 */
        float cx = 0.f, cy = 0.f, cz = 0.f;
        for (int t = 0; t < count; ++t) { cx += rows[t][0]; cy += rows[t][1]; cz += rows[t][2]; }
        cx /= (float)count; cy /= (float)count; cz /= (float)count;
        float c00 = 0.f, c01 = 0.f, c02 = 0.f, c11 = 0.f, c12 = 0.f, c22 = 0.f;
        for (int t = 0; t < count; ++t) {
          const float dx = rows[t][0] - cx, dy = rows[t][1] - cy, dz = rows[t][2] - cz;
          c00 += dx * dx; c01 += dx * dy; c02 += dx * dz; c11 += dy * dy; c12 += dy * dz; c22 += dz * dz;
        }
        const float inv = (float)(count - 1);
        float ev[3], V[9];
        sym3_eigen_f(c00 / inv, c01 / inv, c02 / inv, c11 / inv, c12 / inv, c22 / inv, ev, V);
        /* :137-156 with normal = eigenvector of the smallest eigenvalue */
        int valid = 0;
        for (int t = 0; t < count; ++t) {
          const float dist = fabsf((V[0] * (rows[t][0] - cx) + V[3] * (rows[t][1] - cy)) + V[6] * (rows[t][2] - cz));
          if (dist < P->plane_distance_threshold) valid++;
        }
        const int plane_ok = (float)valid >= P->valid_points_threshold * (float)count;
        float l1, l2, l3, nx, ny, nz;
        if (plane_ok) { l1 = ev[2]; l2 = ev[1]; l3 = ev[0]; nx = V[0]; ny = V[3]; nz = V[6]; }
        else {
          if (!P->use_all_points) continue; /* :1192-1195 */
          l1 = l2 = l3 = -1.f;
          nx = V[2]; ny = V[5]; nz = V[8]; /* un-swapped column 2: the largest eigenvalue's vector (see header) */
          st_invalid++;
        }
        const float nn = sqrtf((nx * nx + ny * ny) + nz * nz);
        if (nn > 0.f) { nx /= nn; ny /= nn; nz /= nn; }
        if (nz < 0.f) { nx = -nx; ny = -ny; nz = -nz; } /* :1201-1203 */
        const int g = off[i] + j;
        if (out_records12) {
          float* o = out_records12 + 12 * n_out;
          memset(o, 0, sizeof(float) * 12);
          o[0] = rp[3 * g]; o[1] = rp[3 * g + 1]; o[2] = rp[3 * g + 2]; o[3] = 1.0f; /* PointXYZINormal: data[3] = 1 */
          o[4] = nx; o[5] = ny; o[6] = nz;
          o[8] = ri[g]; /* intensity */
          o[9] = 0.f;   /* curvature (only written by the "curvature" presample method) */
        }
        if (out_eigenvalues) { out_eigenvalues[3 * n_out] = l1; out_eigenvalues[3 * n_out + 1] = l2; out_eigenvalues[3 * n_out + 2] = l3; }
        const float planarity = (l2 - l3) / l1;
        const int cand = plane_ok && (planarity > P->planarity_threshold);
        if (out_candidate) out_candidate[n_out] = (uint8_t)cand;
        if (out_src_index) out_src_index[n_out] = rs[g];
        st_cand += cand;
        n_out++;
      }
    }
    free(ring); free(inten); free(cnt); free(off); free(rp); free(ri); free(rs); free(fill);
  }
  free(kept);
  if (stats4) { stats4[0] = st_ringed; stats4[1] = st_fail; stats4[2] = st_invalid; stats4[3] = st_cand; }
  return n_out;
}
