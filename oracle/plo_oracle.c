/*
 * plo_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 * See plo_oracle.h for the contract, the "parity unpinned" note and the list
 * of documented deviations.  Compile with -ffp-contract=off: the distance
 * arithmetic must not be contracted into FMAs (bit-exact neighbour sets).
 *
 * Each function names the reference lines it restates (paths relative to the
 * reference root, e.g. src/imls_icp.cpp:301-483).
 */
#include "plo_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------ */
/* context                                                                   */
/* ------------------------------------------------------------------------ */

typedef struct {
  double lo[3], hi[3];
  int32_t left, right; /* children, -1 for leaf */
  int32_t start, count;
} kd_node;

struct orc_ctx {
  orc_params p;
  int nthreads;
  /* target */
  int64_t n;       /* finite points */
  double* tp;      /* n x 3, original (stripped) order: m_targetKDTreeDataBase */
  float* tnf;      /* n x 3 float normals as delivered */
  double* tn;      /* n x 3 normals used by the matcher (lazily built) */
  int tn_valid;
  /* kd-tree over tp */
  kd_node* nodes;
  int32_t n_nodes, cap_nodes;
  double* kp;      /* n x 3 in tree order */
  int32_t* kidx;   /* tree order -> original index */
  double build_seconds;
  /* source */
  int64_t m;
  float* sp; /* m x 3 */
  float* sn; /* m x 3 */
};

static double now_seconds(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void orc_default_params(orc_params* p) {
  /* config.json:84-150 */
  memset(p, 0, sizeof(*p));
  p->iterations = 30;
  p->h = 1.0;
  p->r = 3.0;
  p->r_normal = 1.0;
  p->is_get_normals = 1;
  p->search_number_normal = 10;
  p->search_number = 20;
  p->normal_angle_constraint = 1;
  p->angle_diff_threshold = 30.0;
  p->transform_normal = 0;
  p->correspond_number = 6;
  p->delta_dist_threshold = 0.001;
  p->delta_angle_threshold = 0.0001745353;
  p->solver = ORC_SOLVER_WLS;
  p->weight_mode = ORC_W_UNIT;
  p->ransac_distance_threshold = 0.8;
  p->huber_threshold = 0.648;
  p->ls_threshold = 0.02;
  p->ransac_max_iterations = 5000;
  p->ransac_min_inliers_percentage = 0.95;
  p->ransac_final = ORC_FINAL_DRPM;
  p->drpm_threshold = 0.05;
  p->drpm_stdev_points = 0.02;
  p->drpm_stdev_normals = 0.05;
  p->ransac_seed = 1;
}

orc_ctx* orc_create(void) {
  orc_ctx* c = (orc_ctx*)calloc(1, sizeof(orc_ctx));
  orc_default_params(&c->p);
  c->nthreads = 0;
  return c;
}

static void free_target(orc_ctx* c) {
  free(c->tp); free(c->tnf); free(c->tn); free(c->nodes); free(c->kp); free(c->kidx);
  c->tp = NULL; c->tnf = NULL; c->tn = NULL; c->nodes = NULL; c->kp = NULL; c->kidx = NULL;
  c->n = 0; c->n_nodes = c->cap_nodes = 0; c->tn_valid = 0;
}

void orc_destroy(orc_ctx* c) {
  if (!c) return;
  free_target(c);
  free(c->sp); free(c->sn);
  free(c);
}

void orc_set_params(orc_ctx* c, const orc_params* p) {
  c->p = *p;
  c->tn_valid = 0; /* PCA normals depend on r_normal / search_number_normal */
}

void orc_set_threads(orc_ctx* c, int nthreads) { c->nthreads = nthreads; }

int orc_get_threads(const orc_ctx* c) {
#ifdef _OPENMP
  return c->nthreads > 0 ? c->nthreads : omp_get_max_threads();
#else
  (void)c;
  return 1;
#endif
}

int64_t orc_target_size(const orc_ctx* c) { return c->n; }
int64_t orc_source_size(const orc_ctx* c) { return c->m; }
double orc_last_build_seconds(const orc_ctx* c) { return c->build_seconds; }

/* ------------------------------------------------------------------------ */
/* kd-tree (stand-in for Nabo::NNSearchD::createKDTreeLinearHeap,            */
/* src/imls_icp.cpp:101).  Any exact structure yields the same result set;   */
/* the result is defined by orc_knn's contract, not by the traversal.        */
/* ------------------------------------------------------------------------ */

#define KD_LEAF 12

static inline double dist2_3(const double* a, const double* b) {
  /* libnabo leaf loop: dist += diff*diff in dimension order, double */
  const double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  return (dx * dx + dy * dy) + dz * dz;
}

static inline double box_dist2(const kd_node* nd, const double* q) {
  double e[3];
  for (int a = 0; a < 3; ++a) {
    double lo = nd->lo[a] - q[a], hi = q[a] - nd->hi[a];
    double v = lo > hi ? lo : hi;
    e[a] = v > 0.0 ? v : 0.0;
  }
  return (e[0] * e[0] + e[1] * e[1]) + e[2] * e[2];
}

static void kd_select(double* kp, int32_t* kidx, int64_t lo, int64_t hi, int64_t nth, int axis) {
  /* quickselect on [lo,hi) so that element nth is in sorted position by (coord, idx) */
  while (hi - lo > 1) {
    int64_t mid = lo + (hi - lo) / 2;
    /* median of three */
    int64_t a = lo, b = mid, cidx = hi - 1;
#define KEYLT(i, j) (kp[3 * (i) + axis] < kp[3 * (j) + axis] || \
                     (kp[3 * (i) + axis] == kp[3 * (j) + axis] && kidx[i] < kidx[j]))
    int64_t piv;
    if (KEYLT(a, b)) { piv = KEYLT(b, cidx) ? b : (KEYLT(a, cidx) ? cidx : a); }
    else { piv = KEYLT(a, cidx) ? a : (KEYLT(b, cidx) ? cidx : b); }
    double pv = kp[3 * piv + axis];
    int32_t pi = kidx[piv];
    int64_t i = lo, j = hi - 1;
    while (i <= j) {
      while (kp[3 * i + axis] < pv || (kp[3 * i + axis] == pv && kidx[i] < pi)) ++i;
      while (kp[3 * j + axis] > pv || (kp[3 * j + axis] == pv && kidx[j] > pi)) --j;
      if (i <= j) {
        double t0 = kp[3 * i], t1 = kp[3 * i + 1], t2 = kp[3 * i + 2];
        kp[3 * i] = kp[3 * j]; kp[3 * i + 1] = kp[3 * j + 1]; kp[3 * i + 2] = kp[3 * j + 2];
        kp[3 * j] = t0; kp[3 * j + 1] = t1; kp[3 * j + 2] = t2;
        int32_t ti = kidx[i]; kidx[i] = kidx[j]; kidx[j] = ti;
        ++i; --j;
      }
    }
    if (nth <= j) hi = j + 1;
    else if (nth >= i) lo = i;
    else return;
#undef KEYLT
  }
}

static int32_t kd_new_node(orc_ctx* c) {
  if (c->n_nodes == c->cap_nodes) {
    c->cap_nodes = c->cap_nodes ? c->cap_nodes * 2 : 1024;
    c->nodes = (kd_node*)realloc(c->nodes, sizeof(kd_node) * (size_t)c->cap_nodes);
  }
  return c->n_nodes++;
}

static int32_t kd_build(orc_ctx* c, int64_t lo, int64_t hi) {
  int32_t id = kd_new_node(c);
  double blo[3] = {INFINITY, INFINITY, INFINITY}, bhi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int64_t i = lo; i < hi; ++i)
    for (int a = 0; a < 3; ++a) {
      double v = c->kp[3 * i + a];
      if (v < blo[a]) blo[a] = v;
      if (v > bhi[a]) bhi[a] = v;
    }
  {
    kd_node* nd = &c->nodes[id];
    memcpy(nd->lo, blo, sizeof(blo));
    memcpy(nd->hi, bhi, sizeof(bhi));
    nd->start = (int32_t)lo;
    nd->count = (int32_t)(hi - lo);
    nd->left = nd->right = -1;
  }
  if (hi - lo <= KD_LEAF) return id;
  int axis = 0;
  double ext = bhi[0] - blo[0];
  if (bhi[1] - blo[1] > ext) { ext = bhi[1] - blo[1]; axis = 1; }
  if (bhi[2] - blo[2] > ext) { ext = bhi[2] - blo[2]; axis = 2; }
  int64_t mid = lo + (hi - lo) / 2;
  kd_select(c->kp, c->kidx, lo, hi, mid, axis);
  int32_t l = kd_build(c, lo, mid);
  int32_t r = kd_build(c, mid, hi);
  c->nodes[id].left = l;
  c->nodes[id].right = r;
  return id;
}

typedef struct {
  int k, cnt;
  double r2;
  int allow_self;
  int32_t* idx;
  double* d2;
} topk;

static inline void topk_init(topk* t, int k, double r, int allow_self, int32_t* idx, double* d2) {
  t->k = k; t->cnt = 0; t->r2 = r * r; t->allow_self = allow_self; t->idx = idx; t->d2 = d2;
  for (int i = 0; i < k; ++i) { idx[i] = -1; d2[i] = INFINITY; }
}

static inline void topk_offer(topk* t, double d, int32_t id) {
  /* libnabo acceptance test (kdtree_cpu recurseKnn leaf loop) with the D3 tie rule:
   * d <= maxRadius2 && better than current k-th && (allowSelfMatch || d > epsilon) */
  if (!(d <= t->r2)) return;
  if (!t->allow_self && !(d > DBL_EPSILON)) return;
  int k = t->k;
  if (t->cnt == k) {
    double wd = t->d2[k - 1];
    int32_t wi = t->idx[k - 1];
    if (!(d < wd || (d == wd && id < wi))) return;
  }
  int pos = t->cnt < k ? t->cnt : k - 1;
  while (pos > 0 && (t->d2[pos - 1] > d || (t->d2[pos - 1] == d && t->idx[pos - 1] > id))) {
    t->d2[pos] = t->d2[pos - 1];
    t->idx[pos] = t->idx[pos - 1];
    --pos;
  }
  t->d2[pos] = d;
  t->idx[pos] = id;
  if (t->cnt < k) t->cnt++;
}

static void kd_search(const orc_ctx* c, int32_t node, const double* q, topk* t) {
  const kd_node* nd = &c->nodes[node];
  if (nd->left < 0) {
    const double* p = c->kp + 3 * (int64_t)nd->start;
    for (int32_t i = 0; i < nd->count; ++i, p += 3) topk_offer(t, dist2_3(q, p), c->kidx[nd->start + i]);
    return;
  }
  double dl = box_dist2(&c->nodes[nd->left], q);
  double dr = box_dist2(&c->nodes[nd->right], q);
  int32_t first = nd->left, second = nd->right;
  double df = dl, ds = dr;
  if (dr < dl) { first = nd->right; second = nd->left; df = dr; ds = dl; }
  /* box distance is a floating-point lower bound of every member's d2 (same
   * operation order, monotone rounding) => strict '>' pruning is exact even
   * with ties */
  double bound = t->cnt == t->k ? t->d2[t->k - 1] : INFINITY;
  if (bound > t->r2) bound = t->r2;
  if (!(df > bound)) kd_search(c, first, q, t);
  bound = t->cnt == t->k ? t->d2[t->k - 1] : INFINITY;
  if (bound > t->r2) bound = t->r2;
  if (!(ds > bound)) kd_search(c, second, q, t);
}

int orc_knn(const orc_ctx* c, const double q[3], int k, double r, int allow_self,
            int32_t* idx, double* d2) {
  topk t;
  topk_init(&t, k, r, allow_self, idx, d2);
  if (c->n > 0 && q[0] == q[0] && q[1] == q[1] && q[2] == q[2]) kd_search(c, 0, q, &t);
  return t.cnt;
}

int orc_knn_brute(const orc_ctx* c, const double q[3], int k, double r, int allow_self,
                  int32_t* idx, double* d2) {
  topk t;
  topk_init(&t, k, r, allow_self, idx, d2);
  for (int64_t i = 0; i < c->n; ++i) topk_offer(&t, dist2_3(q, c->tp + 3 * i), (int32_t)i);
  return t.cnt;
}

/* ------------------------------------------------------------------------ */
/* clouds                                                                    */
/* ------------------------------------------------------------------------ */

static inline int finite3f(const float* p) { return isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]); }

/* TransformToEnd, src/laser_odometry.cpp:88-114 (see plo_oracle.h) */
void orc_transform_to_end(void* pts, int64_t n, int32_t stride, const double T[16], int transform_normal) {
  char* base = (char*)pts;
  for (int64_t i = 0; i < n; ++i) {
    float* p = (float*)(base + i * stride);
    const double dx = (double)p[0] - T[3], dy = (double)p[1] - T[7], dz = (double)p[2] - T[11];
    p[0] = (float)((T[0] * dx + T[4] * dy) + T[8] * dz); /* row a of R^T = column a of R */
    p[1] = (float)((T[1] * dx + T[5] * dy) + T[9] * dz);
    p[2] = (float)((T[2] * dx + T[6] * dy) + T[10] * dz);
    if (transform_normal) {
      float* nn = (float*)(base + i * stride + 16);
      const double a = (double)nn[0], b = (double)nn[1], c = (double)nn[2];
      nn[0] = (float)((T[0] * a + T[4] * b) + T[8] * c);
      nn[1] = (float)((T[1] * a + T[5] * b) + T[9] * c);
      nn[2] = (float)((T[2] * a + T[6] * b) + T[10] * c);
    }
  }
}

/* src/imls_icp.cpp:80-103 (+ RemoveNANandINFData :58-72; pcl::isFinite tests xyz only) */
int64_t orc_set_target(orc_ctx* c, const void* pts, int64_t n, int32_t stride) {
  free_target(c);
  const char* base = (const char*)pts;
  int64_t kept = 0;
  for (int64_t i = 0; i < n; ++i) kept += finite3f((const float*)(base + i * stride));
  c->n = kept;
  c->tp = (double*)malloc(sizeof(double) * 3 * (size_t)(kept ? kept : 1));
  c->tnf = (float*)malloc(sizeof(float) * 3 * (size_t)(kept ? kept : 1));
  c->tn = (double*)malloc(sizeof(double) * 3 * (size_t)(kept ? kept : 1));
  c->kp = (double*)malloc(sizeof(double) * 3 * (size_t)(kept ? kept : 1));
  c->kidx = (int32_t*)malloc(sizeof(int32_t) * (size_t)(kept ? kept : 1));
  int64_t j = 0;
  for (int64_t i = 0; i < n; ++i) {
    const float* p = (const float*)(base + i * stride);
    if (!finite3f(p)) continue;
    const float* nn = (const float*)(base + i * stride + 16);
    for (int a = 0; a < 3; ++a) {
      c->tp[3 * j + a] = (double)p[a]; /* :96-98 */
      c->tnf[3 * j + a] = nn[a];
    }
    ++j;
  }
  double t0 = now_seconds();
  memcpy(c->kp, c->tp, sizeof(double) * 3 * (size_t)kept);
  for (int64_t i = 0; i < kept; ++i) c->kidx[i] = (int32_t)i;
  c->n_nodes = 0;
  if (kept > 0) kd_build(c, 0, kept);
  c->build_seconds = now_seconds() - t0;
  c->tn_valid = 0;
  return kept;
}

/* src/imls_icp.cpp:74-78 */
int64_t orc_set_source(orc_ctx* c, const void* pts, int64_t n, int32_t stride) {
  free(c->sp); free(c->sn);
  const char* base = (const char*)pts;
  int64_t kept = 0;
  for (int64_t i = 0; i < n; ++i) kept += finite3f((const float*)(base + i * stride));
  c->m = kept;
  c->sp = (float*)malloc(sizeof(float) * 3 * (size_t)(kept ? kept : 1));
  c->sn = (float*)malloc(sizeof(float) * 3 * (size_t)(kept ? kept : 1));
  int64_t j = 0;
  for (int64_t i = 0; i < n; ++i) {
    const float* p = (const float*)(base + i * stride);
    if (!finite3f(p)) continue;
    const float* nn = (const float*)(base + i * stride + 16);
    for (int a = 0; a < 3; ++a) { c->sp[3 * j + a] = p[a]; c->sn[3 * j + a] = nn[a]; }
    ++j;
  }
  return kept;
}

/* ------------------------------------------------------------------------ */
/* small dense linear algebra (Eigen restatements)                           */
/* ------------------------------------------------------------------------ */

/* cyclic Jacobi for a symmetric n x n matrix (n<=6); eigenvalues ascending,
 * eigenvectors in columns — same ordering contract as
 * Eigen::SelfAdjointEigenSolver (src/imls_icp.cpp:776, src/solver.cpp:540) */
static void sym_eigen_n(const double* Ain, int n, double* evals, double* evecs) {
  double A[36], V[36];
  for (int i = 0; i < n * n; ++i) A[i] = Ain[i];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j];
    if (off == 0.0) break;
    if (n == 6) { /* the rotated entry is not zeroed, so `off` bottoms out at rounding noise instead of 0:
                     stop once it is below 1e-30 of the diagonal (5-6 sweeps instead of all 64) */
      double dsum = 0.0;
      for (int i = 0; i < n; ++i) dsum += A[i * n + i] * A[i * n + i];
      if (off <= 1e-30 * dsum) break;
    }
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        double apq = A[p * n + q];
        if (apq == 0.0) continue;
        double app = A[p * n + p], aqq = A[q * n + q];
        double theta = (aqq - app) / (2.0 * apq);
        double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
        for (int k = 0; k < n; ++k) {
          double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = cs * akp - sn * akq;
          A[k * n + q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < n; ++k) {
          double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = cs * apk - sn * aqk;
          A[q * n + k] = sn * apk + cs * aqk;
        }
        for (int k = 0; k < n; ++k) {
          double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = cs * vkp - sn * vkq;
          V[k * n + q] = sn * vkp + cs * vkq;
        }
      }
  }
  int order[6];
  for (int i = 0; i < n; ++i) order[i] = i;
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j)
      if (A[order[j] * n + order[j]] < A[order[i] * n + order[i]]) { int t = order[i]; order[i] = order[j]; order[j] = t; }
  for (int i = 0; i < n; ++i) {
    evals[i] = A[order[i] * n + order[i]];
    for (int k = 0; k < n; ++k) evecs[k * n + i] = V[k * n + order[i]];
  }
}

void orc_sym3_eigen(const double A[9], double evals[3], double evecs[9]) { sym_eigen_n(A, 3, evals, evecs); }
void orc_sym6_eigen(const double A[36], double evals[6], double evecs[36]) { sym_eigen_n(A, 6, evals, evecs); }

/* IMLSICPMatcher::ComputeNormal, src/imls_icp.cpp:753-794.  D2: oriented +z. */
void orc_compute_normal(const double* pts3, int n, double normal[3]) {
  double mu[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i) { mu[0] += pts3[3 * i]; mu[1] += pts3[3 * i + 1]; mu[2] += pts3[3 * i + 2]; } /* :758-763 */
  mu[0] /= n; mu[1] /= n; mu[2] /= n;
  double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n; ++i) { /* :766-771, population covariance */
    double d[3] = {pts3[3 * i] - mu[0], pts3[3 * i + 1] - mu[1], pts3[3 * i + 2] - mu[2]};
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) cov[a * 3 + b] += d[a] * d[b];
  }
  for (int a = 0; a < 9; ++a) cov[a] /= n;
  double ev[3], V[9];
  orc_sym3_eigen(cov, ev, V); /* :776-778: eigenvector of the smallest eigenvalue */
  double v[3] = {V[0], V[3], V[6]};
  double nrm = sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
  if (nrm > 0) { v[0] /= nrm; v[1] /= nrm; v[2] /= nrm; } /* :791 */
  if (v[2] < 0) { v[0] = -v[0]; v[1] = -v[1]; v[2] = -v[2]; } /* D2, scan_registration.cpp:1196-1200 */
  normal[0] = v[0]; normal[1] = v[1]; normal[2] = v[2];
}

/* Eigen::ColPivHouseholderQR::compute + solve restated (Eigen 3.3
 * ColPivHouseholderQR.h computeInPlace/_solve_impl), used at
 * src/solver.cpp:107,137,200,273,576.  A: m x n row-major (destroyed). */
int orc_colpiv_qr_solve(double* A, double* b, int64_t m, int n, double* x, int* rank_out) {
  int size = (int)(m < n ? m : n);
  double cnd[8], cnu[8], hcoef[8];
  int perm[8];
  if (n > 8) return 0;
  double maxnorm = 0.0;
  for (int j = 0; j < n; ++j) {
    double s = 0.0;
    for (int64_t i = 0; i < m; ++i) s += A[i * n + j] * A[i * n + j];
    cnd[j] = cnu[j] = sqrt(s);
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
    double big_sq = bigv * bigv;
    if (nonzero_pivots == size && big_sq < threshold_helper * (double)(m - k)) nonzero_pivots = k;
    if (big != k) {
      for (int64_t i = 0; i < m; ++i) { double t = A[i * n + k]; A[i * n + k] = A[i * n + big]; A[i * n + big] = t; }
      double t = cnu[k]; cnu[k] = cnu[big]; cnu[big] = t;
      t = cnd[k]; cnd[k] = cnd[big]; cnd[big] = t;
      int ti = perm[k]; perm[k] = perm[big]; perm[big] = ti;
    }
    /* makeHouseholderInPlace on A[k:,k] */
    double c0 = A[(int64_t)k * n + k];
    double tail_sq = 0.0;
    for (int64_t i = k + 1; i < m; ++i) tail_sq += A[i * n + k] * A[i * n + k];
    double tau, beta;
    if (tail_sq <= DBL_MIN) {
      tau = 0.0; beta = c0;
      for (int64_t i = k + 1; i < m; ++i) A[i * n + k] = 0.0;
    } else {
      beta = sqrt(c0 * c0 + tail_sq);
      if (c0 >= 0.0) beta = -beta;
      for (int64_t i = k + 1; i < m; ++i) A[i * n + k] /= (c0 - beta);
      tau = (beta - c0) / beta;
    }
    hcoef[k] = tau;
    A[(int64_t)k * n + k] = beta;
    /* apply H = I - tau v v^T (v0 = 1) to trailing columns and to b */
    if (tau != 0.0) {
      for (int j = k + 1; j < n; ++j) {
        double s = A[(int64_t)k * n + j];
        for (int64_t i = k + 1; i < m; ++i) s += A[i * n + k] * A[i * n + j];
        s *= tau;
        A[(int64_t)k * n + j] -= s;
        for (int64_t i = k + 1; i < m; ++i) A[i * n + j] -= s * A[i * n + k];
      }
    }
    /* norm downdate (LAPACK dlaqp2 style, as Eigen) */
    for (int j = k + 1; j < n; ++j) {
      if (cnu[j] != 0.0) {
        double temp = fabs(A[(int64_t)k * n + j]) / cnu[j];
        temp = (1.0 + temp) * (1.0 - temp);
        if (temp < 0.0) temp = 0.0;
        double ratio = cnu[j] / cnd[j];
        double temp2 = temp * ratio * ratio;
        if (temp2 <= downdate_thr) {
          double s = 0.0;
          for (int64_t i = k + 1; i < m; ++i) s += A[i * n + j] * A[i * n + j];
          cnd[j] = cnu[j] = sqrt(s);
        } else {
          cnu[j] *= sqrt(temp);
        }
      }
    }
  }
  /* _solve_impl: c = Q^T b using the first nonzero_pivots reflectors */
  for (int j = 0; j < n; ++j) x[j] = 0.0;
  if (rank_out) *rank_out = nonzero_pivots;
  if (nonzero_pivots == 0) return 1;
  for (int k = 0; k < nonzero_pivots; ++k) {
    double tau = hcoef[k];
    if (tau == 0.0) continue;
    double s = b[k];
    for (int64_t i = k + 1; i < m; ++i) s += A[i * n + k] * b[i];
    s *= tau;
    b[k] -= s;
    for (int64_t i = k + 1; i < m; ++i) b[i] -= s * A[i * n + k];
  }
  double y[8];
  for (int i = nonzero_pivots - 1; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < nonzero_pivots; ++j) s -= A[(int64_t)i * n + j] * y[j];
    y[i] = s / A[(int64_t)i * n + i];
  }
  for (int i = 0; i < nonzero_pivots; ++i) x[perm[i]] = y[i];
  return 1;
}

/* Eigen::AngleAxisd(rot.norm(), rot.normalized()).toRotationMatrix(),
 * src/solver.cpp:203-205 (Eigen AngleAxis.h toRotationMatrix; normalized()
 * leaves a zero vector unchanged).  R row-major. */
void orc_angle_axis(const double rot[3], double R[9]) {
  double z = (rot[0] * rot[0] + rot[1] * rot[1]) + rot[2] * rot[2];
  double angle = sqrt(z);
  double ax[3] = {rot[0], rot[1], rot[2]};
  if (z > 0.0) { ax[0] /= angle; ax[1] /= angle; ax[2] /= angle; }
  double s = sin(angle), cc = cos(angle);
  double sa[3] = {s * ax[0], s * ax[1], s * ax[2]};
  double ca[3] = {(1.0 - cc) * ax[0], (1.0 - cc) * ax[1], (1.0 - cc) * ax[2]};
  double tmp;
  tmp = ca[0] * ax[1]; R[0 * 3 + 1] = tmp - sa[2]; R[1 * 3 + 0] = tmp + sa[2];
  tmp = ca[0] * ax[2]; R[0 * 3 + 2] = tmp + sa[1]; R[2 * 3 + 0] = tmp - sa[1];
  tmp = ca[1] * ax[2]; R[1 * 3 + 2] = tmp - sa[0]; R[2 * 3 + 1] = tmp + sa[0];
  R[0] = ca[0] * ax[0] + cc; R[4] = ca[1] * ax[1] + cc; R[8] = ca[2] * ax[2] + cc;
}

static double det3(const double* M) {
  return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

/* JacobiSVD(R, FullU|FullV); R = U V^T; det fix — src/solver.cpp:207-213.
 * One-sided (Hestenes) Jacobi; singular values sorted descending so that
 * U.col(2) is the direction Eigen would flip. */
void orc_polar_uvt(const double Rin[9], double out[9]) {
  double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  memcpy(A, Rin, sizeof(A));
  for (int sweep = 0; sweep < 64; ++sweep) {
    int rotated = 0;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double a = 0, b = 0, g = 0;
        for (int i = 0; i < 3; ++i) { a += A[i * 3 + p] * A[i * 3 + p]; b += A[i * 3 + q] * A[i * 3 + q]; g += A[i * 3 + p] * A[i * 3 + q]; }
        if (fabs(g) <= 1e-300 || fabs(g) <= DBL_EPSILON * 0.25 * sqrt(a * b)) continue;
        rotated = 1;
        double zeta = (b - a) / (2.0 * g);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
        for (int i = 0; i < 3; ++i) {
          double ap = A[i * 3 + p], aq = A[i * 3 + q];
          A[i * 3 + p] = cs * ap - sn * aq; A[i * 3 + q] = sn * ap + cs * aq;
          double vp = V[i * 3 + p], vq = V[i * 3 + q];
          V[i * 3 + p] = cs * vp - sn * vq; V[i * 3 + q] = sn * vp + cs * vq;
        }
      }
    if (!rotated) break;
  }
  double sv[3];
  int ord[3] = {0, 1, 2};
  for (int j = 0; j < 3; ++j) sv[j] = sqrt(A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j]);
  for (int i = 0; i < 3; ++i)
    for (int j = i + 1; j < 3; ++j)
      if (sv[ord[j]] > sv[ord[i]]) { int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
  double U[9], Vs[9];
  for (int j = 0; j < 3; ++j) {
    int s = ord[j];
    for (int i = 0; i < 3; ++i) {
      U[i * 3 + j] = sv[s] > 0 ? A[i * 3 + s] / sv[s] : (i == j ? 1.0 : 0.0);
      Vs[i * 3 + j] = V[i * 3 + s];
    }
  }
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += U[i * 3 + k] * Vs[j * 3 + k];
        out[i * 3 + j] = s;
      }
    if (pass == 0 && det3(out) < 0) { U[2] = -U[2]; U[5] = -U[5]; U[8] = -U[8]; } /* :209-213 */
    else break;
  }
}

/* ------------------------------------------------------------------------ */
/* solvers                                                                   */
/* ------------------------------------------------------------------------ */

static inline void ab_row(const double* s, const double* d, const double* n, double a[6], double* b) {
  /* src/solver.cpp:180-193 (identical copies :89-104, :255-273, :515-528) */
  a[0] = n[2] * s[1] - n[1] * s[2];
  a[1] = n[0] * s[2] - n[2] * s[0];
  a[2] = n[1] * s[0] - n[0] * s[1];
  a[3] = n[0]; a[4] = n[1]; a[5] = n[2];
  *b = (n[0] * (d[0] - s[0]) + n[1] * (d[1] - s[1])) + n[2] * (d[2] - s[2]);
}

static void delta_from_x(const double x[6], double delta[16]) {
  /* src/solver.cpp:203-217 */
  double R0[9], R[9];
  orc_angle_axis(x, R0);
  orc_polar_uvt(R0, R);
  for (int i = 0; i < 16; ++i) delta[i] = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) delta[i * 4 + j] = R[i * 3 + j];
  delta[3] = x[3]; delta[7] = x[4]; delta[11] = x[5]; delta[15] = 1.0;
}

/* SolveMotionEstimationProblemWeightedLS, src/solver.cpp:168-220 */
int orc_solve_wls(const double* src, const double* ref, const double* nrm,
                  const double* w, int64_t n, double delta[16]) {
  double* A = (double*)malloc(sizeof(double) * 6 * (size_t)(n ? n : 1));
  double* b = (double*)malloc(sizeof(double) * (size_t)(n ? n : 1));
  for (int64_t i = 0; i < n; ++i) {
    ab_row(src + 3 * i, ref + 3 * i, nrm + 3 * i, A + 6 * i, b + i);
    if (w) { /* :196-198 */
      double sw = sqrt(w[i]);
      for (int a = 0; a < 6; ++a) A[6 * i + a] = sw * A[6 * i + a];
      b[i] = sw * b[i];
    }
  }
  double x[6];
  orc_colpiv_qr_solve(A, b, n, 6, x, NULL); /* :200 */
  delta_from_x(x, delta);
  free(A); free(b);
  return 1;
}

void orc_normal_equations(const double* src, const double* ref, const double* nrm,
                          const double* w, int64_t n, double H21[21], double g6[6],
                          double* sw, double* swbb) {
  long double H[21], g[6], s_w = 0, s_bb = 0;
  for (int i = 0; i < 21; ++i) H[i] = 0;
  for (int i = 0; i < 6; ++i) g[i] = 0;
  for (int64_t i = 0; i < n; ++i) {
    double a[6], b;
    ab_row(src + 3 * i, ref + 3 * i, nrm + 3 * i, a, &b);
    double wi = w ? w[i] : 1.0;
    int t = 0;
    for (int p = 0; p < 6; ++p)
      for (int q = p; q < 6; ++q) H[t++] += (long double)wi * a[p] * a[q];
    for (int p = 0; p < 6; ++p) g[p] += (long double)wi * a[p] * b;
    s_w += wi; s_bb += (long double)wi * b * b;
  }
  for (int i = 0; i < 21; ++i) H21[i] = (double)H[i];
  for (int i = 0; i < 6; ++i) g6[i] = (double)g[i];
  *sw = (double)s_w; *swbb = (double)s_bb;
}

typedef struct { double key; int64_t idx; } key_idx;
static int cmp_key_idx(const void* a, const void* b) {
  const key_idx* x = (const key_idx*)a; const key_idx* y = (const key_idx*)b;
  if (x->key < y->key) return -1;
  if (x->key > y->key) return 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}

/* SolveMotionEstimationProblemLS, src/solver.cpp:74-166 (2%/98% |residual| trim).
 * std::sort there is unstable; ties are broken by index here. */
int orc_solve_ls(const double* src, const double* ref, const double* nrm,
                 int64_t n, double threshold, double delta[16]) {
  size_t nn = (size_t)(n ? n : 1);
  double* A = (double*)malloc(sizeof(double) * 6 * nn);
  double* b = (double*)malloc(sizeof(double) * nn);
  double* A2 = (double*)malloc(sizeof(double) * 6 * nn);
  double* b2 = (double*)malloc(sizeof(double) * nn);
  for (int64_t i = 0; i < n; ++i) ab_row(src + 3 * i, ref + 3 * i, nrm + 3 * i, A + 6 * i, b + i);
  memcpy(A2, A, sizeof(double) * 6 * nn);
  memcpy(b2, b, sizeof(double) * nn);
  double x[6];
  orc_colpiv_qr_solve(A2, b2, n, 6, x, NULL); /* :107 */
  key_idx* ki = (key_idx*)malloc(sizeof(key_idx) * nn);
  for (int64_t i = 0; i < n; ++i) { /* :110 residuals = A x - b */
    double r = 0;
    for (int a = 0; a < 6; ++a) r += A[6 * i + a] * x[a];
    ki[i].key = fabs(r - b[i]);
    ki[i].idx = i;
  }
  qsort(ki, (size_t)n, sizeof(key_idx), cmp_key_idx); /* :118-122 */
  int64_t lower = (int64_t)(threshold * (double)n);          /* :124 */
  int64_t upper = (int64_t)((1.0 - threshold) * (double)n);  /* :125 */
  if (upper > n - 1) upper = n - 1; /* reference would read out of bounds at threshold=0 */
  int64_t cnt = upper - lower + 1;
  if (cnt < 0) cnt = 0;
  for (int64_t i = 0; i < cnt; ++i) { /* :131-134 */
    memcpy(A2 + 6 * i, A + 6 * ki[lower + i].idx, sizeof(double) * 6);
    b2[i] = b[ki[lower + i].idx];
  }
  orc_colpiv_qr_solve(A2, b2, cnt, 6, x, NULL); /* :137 */
  delta_from_x(x, delta);
  free(A); free(b); free(A2); free(b2); free(ki);
  return 1;
}

static inline void apply_T(const double T[16], const double* s, double out[3]) {
  for (int i = 0; i < 3; ++i) out[i] = ((T[i * 4] * s[0] + T[i * 4 + 1] * s[1]) + T[i * 4 + 2] * s[2]) + T[i * 4 + 3];
}

/* RANSAC-final weights, src/solver.cpp:334-364 */
int64_t orc_ransac_weights(const double* src, const double* ref, const double* nrm, int64_t n,
                           const double Tbest[16], double distance_threshold,
                           double huber_threshold, int32_t* inlier_idx, double* w) {
  double thr2 = huber_threshold * distance_threshold; /* :339 */
  int64_t cnt = 0;
  double sum = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    double tp[3];
    apply_T(Tbest, src + 3 * i, tp);
    const double* d = ref + 3 * i; const double* nn = nrm + 3 * i;
    double distance = fabs(((tp[0] - d[0]) * nn[0] + (tp[1] - d[1]) * nn[1]) + (tp[2] - d[2]) * nn[2]); /* :348 */
    if (distance < distance_threshold) {
      double ar = exp(-fabs(distance)); /* :350 */
      double wi = sqrt(ar) < thr2 ? ar : 2 * thr2 * sqrt(ar) - thr2 * thr2; /* :351-355 */
      inlier_idx[cnt] = (int32_t)i;
      w[cnt] = wi;
      sum += wi;
      ++cnt;
    }
  }
  if (sum > 0) for (int64_t i = 0; i < cnt; ++i) w[i] /= sum; /* :361-364 */
  return cnt;
}

static double normal_cdf(double mean, double sd, double x) {
  /* boost::math::cdf(normal_distribution(mean, sd), x), include/degeneracy.h:94-95 */
  if (!(sd > 0)) return x >= mean ? 1.0 : 0.0;
  return 0.5 * erfc(-(x - mean) / (sd * sqrt(2.0)));
}

/* SolveMotionEstimationProblemDRPM, src/solver.cpp:499-603 with
 * degeneracy::ComputeNoiseEstimate / ComputeSignalToNoiseProbabilities /
 * SolveWithSnrProbabilities, include/degeneracy.h:14-131 */
int orc_solve_drpm(const double* src, const double* ref, const double* nrm,
                   const double* w, int64_t n, double threshold, double stdev_points,
                   double stdev_normals, double delta[16], double probs[6]) {
  size_t nn = (size_t)(n ? n : 1);
  double* A = (double*)malloc(sizeof(double) * 6 * nn);
  double* b = (double*)malloc(sizeof(double) * nn);
  double H[36], rhs[6];
  memset(H, 0, sizeof(H)); memset(rhs, 0, sizeof(rhs));
  for (int64_t i = 0; i < n; ++i) {
    ab_row(src + 3 * i, ref + 3 * i, nrm + 3 * i, A + 6 * i, b + i);
    double sw = sqrt(w ? w[i] : 1.0); /* :531-533 */
    for (int a = 0; a < 6; ++a) A[6 * i + a] *= sw;
    b[i] *= sw;
    for (int p = 0; p < 6; ++p) {
      for (int q = 0; q < 6; ++q) H[p * 6 + q] += A[6 * i + p] * A[6 * i + q]; /* :537 */
      rhs[p] += A[6 * i + p] * b[i];
    }
  }
  double ev[6], U[36];
  orc_sym6_eigen(H, ev, U); /* :540-542 */
  /* ComputeNoiseEstimate, degeneracy.h:14-72 */
  double mean[36], var[6];
  memset(mean, 0, sizeof(mean)); memset(var, 0, sizeof(var));
  const double sp2 = stdev_points * stdev_points, sn2 = stdev_normals * stdev_normals;
  for (int64_t i = 0; i < n; ++i) {
    const double* pt = src + 3 * i; const double* nm = nrm + 3 * i;
    double wi = w ? w[i] : 1.0;
    double nx[9] = {0, -nm[2], nm[1], nm[2], 0, -nm[0], -nm[1], nm[0], 0};
    double px[9] = {0, -pt[2], pt[1], pt[2], 0, -pt[0], -pt[1], pt[0], 0};
    double B[36];
    memset(B, 0, sizeof(B));
    double pxnx[9];
    for (int r = 0; r < 3; ++r)
      for (int cc = 0; cc < 3; ++cc) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += px[r * 3 + k] * nx[k * 3 + cc];
        pxnx[r * 3 + cc] = s;
      }
    for (int r = 0; r < 3; ++r)
      for (int cc = 0; cc < 3; ++cc) {
        B[r * 6 + cc] = -nx[r * 3 + cc];          /* :44 */
        B[r * 6 + 3 + cc] = pxnx[r * 3 + cc];     /* :45 */
        B[(3 + r) * 6 + 3 + cc] = nx[r * 3 + cc]; /* :46 */
      }
    /* N = diag(sp2 I3, sn2 I3) (isotropic, solver.cpp:486-497,536) */
    double C[36];
    for (int r = 0; r < 6; ++r)
      for (int cc = 0; cc < 6; ++cc) {
        double s = 0;
        for (int k = 0; k < 6; ++k) s += B[r * 6 + k] * (k < 3 ? sp2 : sn2) * B[cc * 6 + k];
        C[r * 6 + cc] = s * wi; /* :53 */
        mean[r * 6 + cc] += C[r * 6 + cc];
      }
    double sq = sqrt(wi);
    double v[6] = {sq * (px[0] * nm[0] + px[1] * nm[1] + px[2] * nm[2]),
                   sq * (px[3] * nm[0] + px[4] * nm[1] + px[5] * nm[2]),
                   sq * (px[6] * nm[0] + px[7] * nm[1] + px[8] * nm[2]),
                   sq * nm[0], sq * nm[1], sq * nm[2]}; /* :58-60 */
    for (int k = 0; k < 6; ++k) { /* :63-69 */
      double a = 0, bb = 0;
      for (int r = 0; r < 6; ++r) {
        double t = 0;
        for (int cc = 0; cc < 6; ++cc) t += C[r * 6 + cc] * U[cc * 6 + k];
        a += U[r * 6 + k] * t;
        bb += U[r * 6 + k] * v[r];
      }
      var[k] += 2 * a * a + 4 * a * bb * bb;
    }
  }
  /* ComputeSignalToNoiseProbabilities, degeneracy.h:74-105, snr_factor = 10 (solver.cpp:547) */
  double minp = INFINITY;
  for (int k = 0; k < 6; ++k) {
    double meas = 0, noise = 0;
    for (int r = 0; r < 6; ++r) {
      double t1 = 0, t2 = 0;
      for (int cc = 0; cc < 6; ++cc) { t1 += H[r * 6 + cc] * U[cc * 6 + k]; t2 += mean[r * 6 + cc] * U[cc * 6 + k]; }
      meas += U[r * 6 + k] * t1; noise += U[r * 6 + k] * t2;
    }
    double sd = sqrt(var[k]);
    double test_point = meas / (1.0 + 10.0);
    double pr = (isnan(noise) || isnan(sd) || isnan(test_point)) ? 0.0 : normal_cdf(noise, sd, test_point);
    probs[k] = pr;
    if (pr < minp) minp = pr;
  }
  double x[6];
  if (minp < threshold) { /* :566-573, SolveWithSnrProbabilities degeneracy.h:107-131 */
    double dps[6];
    for (int i = 0; i < 6; ++i) dps[i] = fabs(ev[i]) > 1e-10 ? probs[i] / ev[i] : 0.0;
    double t[6];
    for (int i = 0; i < 6; ++i) { double s = 0; for (int r = 0; r < 6; ++r) s += U[r * 6 + i] * rhs[r]; t[i] = s * dps[i]; }
    for (int r = 0; r < 6; ++r) { double s = 0; for (int i = 0; i < 6; ++i) s += U[r * 6 + i] * t[i]; x[r] = s; }
  } else {
    orc_colpiv_qr_solve(A, b, n, 6, x, NULL); /* :576 */
  }
  delta_from_x(x, delta); /* :580-600 */
  free(A); free(b);
  return 1;
}

static uint64_t xs64(uint64_t* s) { /* replaces unseeded rand() of src/common.cpp:49 (D5) */
  uint64_t x = *s;
  x ^= x << 13; x ^= x >> 7; x ^= x << 17;
  return *s = x;
}

/* farthestPointSampling(cloud, 3), src/common.cpp:19-82 */
static void fps3(const double* pts, int64_t n, uint64_t* rng, double* mind, int32_t out[3]) {
  int32_t first = (int32_t)(xs64(rng) % (uint64_t)n);
  out[0] = first;
  for (int64_t i = 0; i < n; ++i) mind[i] = sqrt(dist2_3(pts + 3 * first, pts + 3 * i));
  for (int s = 1; s < 3; ++s) {
    double maxd = -1.0; int32_t far = -1;
    for (int64_t i = 0; i < n; ++i) {
      int used = 0;
      for (int t = 0; t < s; ++t) used |= (out[t] == i);
      if (!used && mind[i] > maxd) { maxd = mind[i]; far = (int32_t)i; }
    }
    out[s] = far;
    if (far < 0) { out[s] = out[0]; continue; }
    for (int64_t i = 0; i < n; ++i) {
      double d = sqrt(dist2_3(pts + 3 * far, pts + 3 * i));
      if (d < mind[i]) mind[i] = d;
    }
  }
}

/* SolveMotionEstimationProblemRANSAC, src/solver.cpp:222-385 */
int orc_solve_ransac(const double* src, const double* ref, const double* nrm, int64_t n,
                     const orc_params* p, double delta[16]) {
  if (n <= 0) return 0;
  const int min_inliers = (int)(p->ransac_min_inliers_percentage * (double)n); /* :238 */
  int best = 0;
  double Tbest[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  uint64_t rng = p->ransac_seed ? p->ransac_seed : 1;
  double* mind = (double*)malloc(sizeof(double) * (size_t)n);
  for (int it = 0; it < p->ransac_max_iterations; ++it) { /* :244 */
    int32_t ids[3];
    fps3(src, n, &rng, mind, ids); /* :246-247 */
    double A[18], b[3], x[6], T[16];
    for (int i = 0; i < 3; ++i) ab_row(src + 3 * ids[i], ref + 3 * ids[i], nrm + 3 * ids[i], A + 6 * i, b + i);
    orc_colpiv_qr_solve(A, b, 3, 6, x, NULL); /* :273 */
    delta_from_x(x, T); /* :276-298 */
    int cnt = 0;
    for (int64_t i = 0; i < n; ++i) { /* :300-314 */
      double tp[3];
      apply_T(T, src + 3 * i, tp);
      const double* d = ref + 3 * i; const double* nn = nrm + 3 * i;
      double dist = fabs(((tp[0] - d[0]) * nn[0] + (tp[1] - d[1]) * nn[1]) + (tp[2] - d[2]) * nn[2]);
      if (dist < p->ransac_distance_threshold) ++cnt;
    }
    if (cnt > best) { best = cnt; memcpy(Tbest, T, sizeof(T)); } /* :317-320 */
    if (best > min_inliers) break; /* :323-325 */
  }
  free(mind);
  int32_t* iidx = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  double* w = (double*)malloc(sizeof(double) * (size_t)n);
  int64_t ni = orc_ransac_weights(src, ref, nrm, n, Tbest, p->ransac_distance_threshold, p->huber_threshold, iidx, w);
  double* s2 = (double*)malloc(sizeof(double) * 3 * (size_t)(ni ? ni : 1));
  double* r2 = (double*)malloc(sizeof(double) * 3 * (size_t)(ni ? ni : 1));
  double* n2 = (double*)malloc(sizeof(double) * 3 * (size_t)(ni ? ni : 1));
  for (int64_t i = 0; i < ni; ++i) {
    memcpy(s2 + 3 * i, src + 3 * iidx[i], 24); memcpy(r2 + 3 * i, ref + 3 * iidx[i], 24); memcpy(n2 + 3 * i, nrm + 3 * iidx[i], 24);
  }
  int ok;
  double probs[6];
  if (p->ransac_final == ORC_FINAL_LS) ok = orc_solve_ls(s2, r2, n2, ni, p->ls_threshold, delta);      /* :368-371 */
  else if (p->ransac_final == ORC_FINAL_WLS) ok = orc_solve_wls(s2, r2, n2, w, ni, delta);             /* :372-375 */
  else ok = orc_solve_drpm(s2, r2, n2, w, ni, p->drpm_threshold, p->drpm_stdev_points, p->drpm_stdev_normals, delta, probs); /* :376-379 */
  free(iidx); free(w); free(s2); free(r2); free(n2);
  return ok;
}

/* ------------------------------------------------------------------------ */
/* matcher                                                                   */
/* ------------------------------------------------------------------------ */

static inline int finite3d(const double* v) { return isfinite(v[0]) && isfinite(v[1]) && isfinite(v[2]); }

/* the normals the matcher reads: delivered ones (src/imls_icp.cpp:406,632) or
 * PCA over the search_number_normal nearest within r_normal (:411-433,:647-669
 * -> :753-794).  The PCA normal depends only on the target point, so it is
 * evaluated once per target point.  D1: computed iff all
 * search_number_normal slots are filled. */
static void ensure_target_normals(orc_ctx* c) {
  if (c->tn_valid) return;
  const int64_t n = c->n;
  if (c->p.is_get_normals) {
    for (int64_t i = 0; i < 3 * n; ++i) c->tn[i] = (double)c->tnf[i];
  } else {
    const int k = c->p.search_number_normal;
    const double rn = c->p.r_normal;
    int nt = orc_get_threads(c);
    (void)nt;
#pragma omp parallel for schedule(dynamic, 256) num_threads(nt)
    for (int64_t i = 0; i < n; ++i) {
      int32_t idx[64]; double d2[64]; double nb[64 * 3];
      int kk = k > 64 ? 64 : k;
      /* flags = SORT_RESULTS only => no self match (:414-416) */
      int cnt = orc_knn(c, c->tp + 3 * i, kk, rn, 0, idx, d2);
      if (cnt < kk) {
        c->tn[3 * i] = c->tn[3 * i + 1] = c->tn[3 * i + 2] = INFINITY; /* :418-421 */
      } else {
        for (int j = 0; j < kk; ++j) memcpy(nb + 3 * j, c->tp + 3 * (int64_t)idx[j], 24);
        orc_compute_normal(nb, kk, c->tn + 3 * i);
      }
    }
  }
  c->tn_valid = 1;
}

void orc_get_target_normals(const orc_ctx* cc, double* out) {
  orc_ctx* c = (orc_ctx*)cc;
  ensure_target_normals(c);
  memcpy(out, c->tn, sizeof(double) * 3 * (size_t)c->n);
}

static inline double angle_deg(const double* a, const double* b) {
  /* src/imls_icp.cpp:444-445 / :683-684 */
  double dot = (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
  double na = sqrt((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]);
  double nb = sqrt((b[0] * b[0] + b[1] * b[1]) + b[2] * b[2]);
  double cos_angle = dot / (na * nb);
  return acos(cos_angle) * 180.0 / M_PI;
}

/* IMLSICPMatcher::ImplicitMLSFunction, src/imls_icp.cpp:301-483 (default branch) */
static int imls_height(const orc_ctx* c, const double x[3], const double xn[3], double* height,
                       int32_t* nn_idx, double* nn_d2) {
  const int k = c->p.search_number;
  /* :372-375 knn k, eps 0, SORT_RESULTS|ALLOW_SELF_MATCH, maxRadius r */
  orc_knn(c, x, k, c->p.r, 1, nn_idx, nn_d2);
  double kp[64 * 3], kn[64 * 3];
  int cnt = 0;
  for (int i = 0; i < k; ++i) { /* :380 */
    if (!(nn_d2[i] < INFINITY) || isnan(nn_d2[i])) continue; /* :383-385 */
    const double* p = c->tp + 3 * (int64_t)nn_idx[i];
    if (!finite3d(p)) continue; /* :396-400 */
    const double* nn = c->tn + 3 * (int64_t)nn_idx[i]; /* :406 / :411-433 */
    if (!finite3d(nn)) continue; /* :436-440 */
    if (c->p.normal_angle_constraint) { /* :442-451; NaN compares false => kept */
      if (angle_deg(xn, nn) > c->p.angle_diff_threshold) continue;
    }
    memcpy(kp + 3 * cnt, p, 24);
    memcpy(kn + 3 * cnt, nn, 24);
    ++cnt;
  }
  if (cnt < 3) return 0; /* :463-466 */
  double h_max = sqrt(nn_d2[cnt - 1]) / 3; /* :468 — index into the UNFILTERED list */
  double wsum = 0.0, psum = 0.0;
  for (int i = 0; i < cnt; ++i) { /* :470-478 */
    double dx = x[0] - kp[3 * i], dy = x[1] - kp[3 * i + 1], dz = x[2] - kp[3 * i + 2];
    double diff_norm = (dx * dx + dy * dy) + dz * dz;
    double weight = exp(-diff_norm / h_max / h_max);
    double proj = ((weight * dx) * kn[3 * i] + (weight * dy) * kn[3 * i + 1]) + (weight * dz) * kn[3 * i + 2];
    wsum += weight;
    psum += proj;
  }
  *height = psum / (wsum + 1e-5); /* :480 */
  return 1;
}

int64_t orc_project(orc_ctx* c, const double T[16],
                    float* src_xyz, float* ref_xyz, float* ref_n, int32_t* src_idx,
                    int64_t counters[6],
                    int32_t* status_out, double* height_out,
                    int32_t* nn1_idx_out, double* nn1_d2_out,
                    int32_t* nn_idx_out, double* nn_d2_out) {
  ensure_target_normals(c);
  const int64_t m = c->m;
  const int k = c->p.search_number;
  size_t mm = (size_t)(m ? m : 1);
  int32_t* status = (int32_t*)malloc(sizeof(int32_t) * mm);
  float* xs = (float*)malloc(sizeof(float) * 3 * mm);
  float* ys = (float*)malloc(sizeof(float) * 3 * mm);
  float* ns = (float*)malloc(sizeof(float) * 3 * mm);
  const int want_nn = nn_idx_out != NULL || nn_d2_out != NULL;
  int nt = orc_get_threads(c);
  (void)nt;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
  for (int64_t i = 0; i < m; ++i) {
    /* src/laser_odometry.cpp:527-549: p' = rPose*[p;1] in double, stored float */
    const float* sp = c->sp + 3 * i;
    const float* snf = c->sn + 3 * i;
    double pd[3] = {(double)sp[0], (double)sp[1], (double)sp[2]};
    double td[3];
    apply_T(T, pd, td);
    float xf[3] = {(float)td[0], (float)td[1], (float)td[2]};
    float nf[3] = {snf[0], snf[1], snf[2]};
    if (c->p.transform_normal) { /* :541-548 */
      double nd[3] = {(double)snf[0], (double)snf[1], (double)snf[2]};
      for (int a = 0; a < 3; ++a) nf[a] = (float)((T[a * 4] * nd[0] + T[a * 4 + 1] * nd[1]) + T[a * 4 + 2] * nd[2]);
    }
    double x[3] = {(double)xf[0], (double)xf[1], (double)xf[2]};   /* imls_icp.cpp:556 */
    double xn[3] = {(double)nf[0], (double)nf[1], (double)nf[2]};  /* :557 */
    xs[3 * i] = xf[0]; xs[3 * i + 1] = xf[1]; xs[3 * i + 2] = xf[2];
    int32_t st = ORC_OK;
    double height = NAN;
    int32_t i1 = -1; double d1 = INFINITY;
    int32_t nidx[64]; double nd2[64];
    int have_nn = 0;
    double nearNormal[3] = {NAN, NAN, NAN};
    do {
      orc_knn(c, x, 1, c->p.r, 0, &i1, &d1); /* :601-609 */
      if (i1 < 0 || i1 >= c->n) { st = ORC_DROP_NO_NORMAL; break; } /* :612-617 */
      if (d1 > c->p.h * c->p.h) { st = ORC_DROP_TOO_FAR; break; }  /* :620-625 */
      memcpy(nearNormal, c->tn + 3 * (int64_t)i1, 24);             /* :630-633 / :645-669 */
      if (!finite3d(nearNormal)) { st = ORC_DROP_INVALID_NORMAL; break; } /* :673-679 */
      if (c->p.normal_angle_constraint && angle_deg(xn, nearNormal) > c->p.angle_diff_threshold) {
        st = ORC_DROP_NORMAL_CONSTRAINT; break; /* :681-692 */
      }
      have_nn = 1;
      if (!imls_height(c, x, xn, &height, nidx, nd2)) { st = ORC_DROP_MLS_FAIL; break; } /* :696-701 */
      if (isnan(height) || isinf(height)) { st = ORC_DROP_NAN_INF_HEIGHT; break; }      /* :703-717 */
      /* :719-731 */
      ys[3 * i] = (float)(x[0] - height * nearNormal[0]);
      ys[3 * i + 1] = (float)(x[1] - height * nearNormal[1]);
      ys[3 * i + 2] = (float)(x[2] - height * nearNormal[2]);
      ns[3 * i] = (float)nearNormal[0]; ns[3 * i + 1] = (float)nearNormal[1]; ns[3 * i + 2] = (float)nearNormal[2];
    } while (0);
    status[i] = st;
    if (status_out) status_out[i] = st;
    if (height_out) height_out[i] = height;
    if (nn1_idx_out) nn1_idx_out[i] = i1;
    if (nn1_d2_out) nn1_d2_out[i] = d1;
    if (want_nn) {
      if (!have_nn) orc_knn(c, x, k, c->p.r, 1, nidx, nd2);
      if (nn_idx_out) memcpy(nn_idx_out + (int64_t)k * i, nidx, sizeof(int32_t) * (size_t)k);
      if (nn_d2_out) memcpy(nn_d2_out + (int64_t)k * i, nd2, sizeof(double) * (size_t)k);
    }
  }
  /* D4: order-preserving compaction instead of vector::erase */
  int64_t np = 0;
  if (counters) for (int a = 0; a < 6; ++a) counters[a] = 0;
  for (int64_t i = 0; i < m; ++i) {
    if (status[i] == ORC_OK) {
      if (src_xyz) memcpy(src_xyz + 3 * np, xs + 3 * i, 12);
      if (ref_xyz) memcpy(ref_xyz + 3 * np, ys + 3 * i, 12);
      if (ref_n) memcpy(ref_n + 3 * np, ns + 3 * i, 12);
      if (src_idx) src_idx[np] = (int32_t)i;
      ++np;
    } else if (counters) {
      counters[status[i] - 1]++;
    }
  }
  free(status); free(xs); free(ys); free(ns);
  return np;
}

/* ------------------------------------------------------------------------ */
/* driver loop, src/laser_odometry.cpp:484-485,524-647                       */
/* ------------------------------------------------------------------------ */

int orc_register(orc_ctx* c, const double T0[16], double T[16], orc_reg_stats* stats,
                 int64_t* per_iter_pairs) {
  const int64_t m = c->m;
  size_t mm = (size_t)(m ? m : 1);
  float* sx = (float*)malloc(12 * mm); float* rx = (float*)malloc(12 * mm); float* rn = (float*)malloc(12 * mm);
  double* sd = (double*)malloc(24 * mm); double* rd = (double*)malloc(24 * mm); double* nd = (double*)malloc(24 * mm);
  double* w = (double*)malloc(8 * mm);
  int32_t* iidx = (int32_t*)malloc(4 * mm);
  double rPose[16];
  memcpy(rPose, T0, sizeof(rPose)); /* :484-485 (identity in the reference) */
  int status = ORC_REG_MAX_ITERS;
  int it = 0;
  int64_t np = 0;
  int64_t counters[6] = {0, 0, 0, 0, 0, 0};
  double rms = 0.0;
  for (it = 0; it < c->p.iterations; ++it) { /* :524 */
    np = orc_project(c, rPose, sx, rx, rn, NULL, counters, NULL, NULL, NULL, NULL, NULL, NULL); /* :527-559 */
    if (per_iter_pairs) per_iter_pairs[it] = np;
    if (np < c->p.correspond_number) { status = ORC_REG_TOO_FEW_PAIRS; break; } /* :570-576 */
    for (int64_t i = 0; i < 3 * np; ++i) { sd[i] = (double)sx[i]; rd[i] = (double)rx[i]; nd[i] = (double)rn[i]; } /* :595-599, common.h:51-75 */
    {
      double s = 0;
      for (int64_t i = 0; i < np; ++i) {
        double b = (nd[3 * i] * (rd[3 * i] - sd[3 * i]) + nd[3 * i + 1] * (rd[3 * i + 1] - sd[3 * i + 1])) + nd[3 * i + 2] * (rd[3 * i + 2] - sd[3 * i + 2]);
        s += b * b;
      }
      rms = sqrt(s / (double)np);
    }
    double delta[16];
    int ok; /* :609, dispatcher :173-275 */
    if (c->p.solver == ORC_SOLVER_LS) ok = orc_solve_ls(sd, rd, nd, np, c->p.ls_threshold, delta);
    else if (c->p.solver == ORC_SOLVER_RANSAC) ok = orc_solve_ransac(sd, rd, nd, np, &c->p, delta);
    else if (c->p.weight_mode == ORC_W_HUBER_EXP) {
      /* RANSAC-final weights at T_best = I then WeightedLS on the inliers */
      const double I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
      int64_t ni = orc_ransac_weights(sd, rd, nd, np, I4, c->p.ransac_distance_threshold, c->p.huber_threshold, iidx, w);
      for (int64_t i = 0; i < ni; ++i) { /* in-place gather is safe: iidx[i] >= i */
        memmove(sd + 3 * i, sd + 3 * iidx[i], 24); memmove(rd + 3 * i, rd + 3 * iidx[i], 24); memmove(nd + 3 * i, nd + 3 * iidx[i], 24);
      }
      ok = orc_solve_wls(sd, rd, nd, w, ni, delta);
    } else ok = orc_solve_wls(sd, rd, nd, NULL, np, delta);
    if (!ok) { status = ORC_REG_SOLVE_FAILED; break; } /* :611-616 */
    double nP[16]; /* :619 rPose = deltaTrans * rPose */
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        double s = 0;
        for (int kk = 0; kk < 4; ++kk) s += delta[i * 4 + kk] * rPose[kk * 4 + j];
        nP[i * 4 + j] = s;
      }
    memcpy(rPose, nP, sizeof(nP));
    double dd = sqrt(delta[3] * delta[3] + delta[7] * delta[7] + delta[11] * delta[11]); /* :628-632 */
    double ct = ((delta[0] + delta[5] + delta[10]) - 1.0) / 2.0; /* :636 */
    ct = ct > 1.0 ? 1.0 : (ct < -1.0 ? -1.0 : ct);
    double da = acos(ct);
    if (dd < c->p.delta_dist_threshold && da < c->p.delta_angle_threshold) { status = ORC_REG_CONVERGED; ++it; break; } /* :643-646 */
  }
  memcpy(T, rPose, sizeof(rPose));
  if (stats) {
    stats->status = status;
    stats->iters = it;
    stats->pairs = np;
    stats->rms = rms;
    memcpy(stats->counters, counters, sizeof(counters));
  }
  free(sx); free(rx); free(rn); free(sd); free(rd); free(nd); free(w); free(iidx);
  return status;
}
