/*
 * plo_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the reference's IMLS-ICP scan-to-map hot path
 * (spirit-man/Planetary-LiDAR-Odometry).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (libplo_cuda.so) never links, includes or calls anything here.
 *
 * PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors
 *  and cannot be compiled in this image (Eigen, libnabo, PCL,
 * ROS absent — SURVEY.md §8c).  The oracle is pinned instead by (1) an
 * independent numpy/scipy restatement (oracle/py/imls_ref.py), (2) brute-force
 * kNN, (3) analytic known-answer cases — see tests/test_oracle_*.py.
 *
 * Third-party arithmetic restated here (none vendored, none version-pinned by
 * the reference, CMakeLists.txt:26-27,40-43):
 *   libnabo  Nabo::NNSearchD::knn         -> orc_knn()          (semantics below)
 *   Eigen    ColPivHouseholderQR::solve   -> orc_colpiv_qr_solve()
 *   Eigen    AngleAxisd::toRotationMatrix -> orc_angle_axis()
 *   Eigen    JacobiSVD (U*V^T)            -> orc_polar_uvt()
 *   Eigen    SelfAdjointEigenSolver<3x3>  -> orc_sym3_eigen()
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference root).
 *
 * Documented deviations, all inert at config.json defaults
 * on tie-free data: D1 knn count, D2 normal sign (+z), D3 ties by index,
 * D4 stable compaction instead of erase, D5 WeightedLS first-class,
 * D6 no disk I/O inside the loop.
 */
#ifndef PLO_ORACLE_H
#define PLO_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status per source point; values 1..6 follow the order of the six drop
 * counters of src/imls_icp.cpp:506-511 */
enum {
  ORC_OK = 0,
  ORC_DROP_NO_NORMAL = 1,         /* :612-617 */
  ORC_DROP_TOO_FAR = 2,           /* :620-625 */
  ORC_DROP_INVALID_NORMAL = 3,    /* :673-679 */
  ORC_DROP_NORMAL_CONSTRAINT = 4, /* :681-692 */
  ORC_DROP_MLS_FAIL = 5,          /* :696-701 */
  ORC_DROP_NAN_INF_HEIGHT = 6     /* :703-717 */
};

enum { ORC_W_UNIT = 0, ORC_W_HUBER_EXP = 1 };
enum { ORC_SOLVER_WLS = 0, ORC_SOLVER_LS = 1, ORC_SOLVER_RANSAC = 2 };
enum { ORC_FINAL_LS = 0, ORC_FINAL_WLS = 1, ORC_FINAL_DRPM = 2 };
enum {
  ORC_REG_CONVERGED = 1,
  ORC_REG_MAX_ITERS = 2,
  ORC_REG_TOO_FEW_PAIRS = 3,
  ORC_REG_SOLVE_FAILED = 4
};

/* mirrors IMLSICPMatcher::setParameters (src/imls_icp.cpp:146-168) plus the
 * driver-loop keys of config.json (laser_odometry.*) */
typedef struct {
  int32_t iterations;              /* solve_method.iterations            (30) */
  double h;                        /* IMLS.h                               (1) */
  double r;                        /* IMLS.r                               (3) */
  double r_normal;                 /* get_normals.r_normal                 (1) */
  int32_t is_get_normals;          /* get_normals.enabled               (true) */
  int32_t search_number_normal;    /* get_normals.search_number_normal    (10) */
  int32_t search_number;           /* "IMLS function".search_number       (20) */
  int32_t normal_angle_constraint; /* normal_angle_constraint.enabled   (true) */
  double angle_diff_threshold;     /* ...angle_diff_threshold             (30) */
  int32_t transform_normal;        /* laser_odometry.transform_normal  (false) */
  int32_t correspond_number;       /* matching_method.correspond_number    (6) */
  double delta_dist_threshold;     /* solve_method.delta_dist_threshold (1e-3) */
  double delta_angle_threshold;    /* ...delta_angle_threshold  (1.745353e-4) */
  int32_t solver;                  /* ORC_SOLVER_*                             */
  int32_t weight_mode;             /* ORC_W_* (WLS only)                       */
  double ransac_distance_threshold;/* RANSAC.distance_threshold          (0.8) */
  double huber_threshold;          /* RANSAC.huber_threshold           (0.648) */
  double ls_threshold;             /* LS.threshold                      (0.02) */
  int32_t ransac_max_iterations;   /* RANSAC.max_iterations             (5000) */
  double ransac_min_inliers_percentage; /*                              (0.95) */
  int32_t ransac_final;            /* ORC_FINAL_*                       (DRPM) */
  double drpm_threshold, drpm_stdev_points, drpm_stdev_normals; /* .05 .02 .05 */
  uint64_t ransac_seed;            /* replaces unseeded rand(), common.cpp:49  */
} orc_params;

void orc_default_params(orc_params* p);

typedef struct orc_ctx orc_ctx;

orc_ctx* orc_create(void);
void orc_destroy(orc_ctx* c);
void orc_set_params(orc_ctx* c, const orc_params* p);
void orc_set_threads(orc_ctx* c, int nthreads); /* <=0: omp_get_max_threads() */
int orc_get_threads(const orc_ctx* c);

/* IMLSICPMatcher::setTargetPointCloud, src/imls_icp.cpp:80-103 (+ :58-72).
 * pts: records of `stride` bytes, float32 xyz at byte 0, float32 normal at
 * byte 16 (pcl::PointXYZINormal, include/common.h:17).  Returns the number of
 * points kept after the non-finite-xyz strip; neighbour indices refer to the
 * stripped cloud (the reference erases in place).  kept_index (nullable, cap n)
 * receives original index of each kept point. */
int64_t orc_set_target(orc_ctx* c, const void* pts, int64_t n, int32_t stride);
int64_t orc_set_source(orc_ctx* c, const void* pts, int64_t n, int32_t stride);
/* TransformToEnd, src/laser_odometry.cpp:88-114 — the step the reference left commented out in
 * accumulateTargetCloud (:118-124): bring a cloud expressed in the previous frame into the current frame,
 * given rPose = [R t] of the registration just done (x_prev = R x_cur + t):  p' = R^-1 (p - t), and
 * n' = R^-1 n when transform_normal.  Double arithmetic, float32 store, in place.  R^-1 is taken as R^T
 * (the reference goes through a quaternion, q_last_curr.inverse().toRotationMatrix()). */
void orc_transform_to_end(void* pts, int64_t n, int32_t stride, const double T[16], int transform_normal);
int64_t orc_target_size(const orc_ctx* c);
int64_t orc_source_size(const orc_ctx* c);
/* target normals actually used (PCA when !is_get_normals): n x 3 doubles */
void orc_get_target_normals(const orc_ctx* c, double* out);

/* libnabo knn restated (call sites src/imls_icp.cpp:372-375,605-607,414-416):
 * accept iff d2 <= r*r and (allow_self || d2 > DBL_EPSILON); k best by
 * (d2, index) ascending; unfilled slots idx=-1, d2=+inf.
 * d2 = ((dx*dx + dy*dy) + dz*dz) in double, no FMA.  Returns #filled. */
int orc_knn(const orc_ctx* c, const double q[3], int k, double r, int allow_self,
            int32_t* idx, double* d2);
/* same contract, O(N) scan — ground truth for the tree */
int orc_knn_brute(const orc_ctx* c, const double q[3], int k, double r, int allow_self,
                  int32_t* idx, double* d2);

/* IMLSICPMatcher::ComputeNormal, src/imls_icp.cpp:753-794 (D2: +z oriented) */
void orc_compute_normal(const double* pts3, int n, double normal[3]);

/* One call of the per-iteration transform (src/laser_odometry.cpp:527-549)
 * followed by IMLSICPMatcher::ProjSourcePtToSurface (src/imls_icp.cpp:496-745)
 * with ImplicitMLSFunction (:301-483) inlined.  T: row-major 4x4.
 * Compacted outputs (capacity = source size): src_xyz/ref_xyz/ref_n are float32
 * triplets exactly as the reference stores them; src_idx = index into the
 * (stripped) source.  Per-query hooks (nullable): status[M], height[M],
 * nn1_idx[M], nn1_d2[M], nn_idx[M*k], nn_d2[M*k] (k = search_number; computed
 * for every query when requested, also for the ones dropped before IMLS). */
int64_t orc_project(orc_ctx* c, const double T[16],
                    float* src_xyz, float* ref_xyz, float* ref_n, int32_t* src_idx,
                    int64_t counters[6],
                    int32_t* status, double* height,
                    int32_t* nn1_idx, double* nn1_d2,
                    int32_t* nn_idx, double* nn_d2);

/* solvers: src/solver.cpp:168-220 (WeightedLS), :74-166 (LS),
 * :222-385 (RANSAC), :499-603 (DRPM).  src/ref/nrm: n x 3 doubles; w: n or NULL
 * (unit).  delta: row-major 4x4.  Return 1 (true) / 0 (false). */
int orc_solve_wls(const double* src, const double* ref, const double* nrm,
                  const double* w, int64_t n, double delta[16]);
int orc_solve_ls(const double* src, const double* ref, const double* nrm,
                 int64_t n, double threshold, double delta[16]);
/* weights of src/solver.cpp:334-364 at T_best (row-major 4x4); returns number
 * of inliers; inlier_idx[n], w[n] (normalised to sum 1) */
int64_t orc_ransac_weights(const double* src, const double* ref, const double* nrm, int64_t n,
                           const double Tbest[16], double distance_threshold,
                           double huber_threshold, int32_t* inlier_idx, double* w);
int orc_solve_drpm(const double* src, const double* ref, const double* nrm,
                   const double* w, int64_t n, double threshold, double stdev_points,
                   double stdev_normals, double delta[16], double probs[6]);
int orc_solve_ransac(const double* src, const double* ref, const double* nrm, int64_t n,
                     const orc_params* p, double delta[16]);

/* A,b normal-equation sums the GPU reduces: H (21 upper-triangular entries,
 * row-major i<=j), g (6), sum w, sum w*b*b, count — for stage parity */
void orc_normal_equations(const double* src, const double* ref, const double* nrm,
                          const double* w, int64_t n, double H21[21], double g6[6],
                          double* sw, double* swbb);

/* small dense pieces, exposed for unit tests */
int orc_colpiv_qr_solve(double* A /* m x n row-major, destroyed */, double* b /* m, destroyed */,
                        int64_t m, int n, double* x, int* rank);
void orc_angle_axis(const double rot[3], double R[9]);
void orc_polar_uvt(const double R[9], double out[9]);
void orc_sym3_eigen(const double A[9], double evals[3], double evecs[9] /* columns */);
void orc_sym6_eigen(const double A[36], double evals[6], double evecs[36] /* columns */);

/* driver loop, src/laser_odometry.cpp:484-485,524-647: rPose starts at T0
 * (the reference uses identity), returns ORC_REG_*.  stats (nullable):
 * iters_done, pairs of last projection, rms of b of last projection.
 * per_iter_pairs (nullable, cap = iterations). */
typedef struct {
  int32_t status;
  int32_t iters;
  int64_t pairs;
  double rms;
  int64_t counters[6]; /* of the last projection */
} orc_reg_stats;
int orc_register(orc_ctx* c, const double T0[16], double T[16], orc_reg_stats* stats,
                 int64_t* per_iter_pairs);

/* timing helpers for bench.py's cpu_baseline leg */
double orc_last_build_seconds(const orc_ctx* c);

#ifdef __cplusplus
}
#endif
/* ---- front-end: ring assignment + windowed PCA normals + planarity presample ----
 * laserCloudHandler of src/scan_registration.cpp at config.json defaults; see plo_oracle_frontend.c for
 * the file:line map and the defined third-party arithmetic. */
typedef struct orc_frontend_params {
  int32_t n_scans;               /* 16 / 32 / 64 (launch parameter scan_line) */
  float min_range, max_range;    /* MINIMUM_RANGE / MAXIMUM_RANGE, src/scan_registration.cpp:62-63 */
  float scan_period;             /* :55 */
  int32_t window_size, iter_step;            /* compute_normal_method.pca */
  float knn_distance_threshold;              /* squared metres (FLANN returns squared distances) */
  float plane_distance_threshold, valid_points_threshold;   /* pca.plane_constraint */
  int32_t use_all_points;                    /* model.use_all_points */
  float planarity_threshold;                 /* presample_method.geometric_features */
} orc_frontend_params;
void orc_frontend_default_params(orc_frontend_params* p);
/* pts: n records of `stride` bytes, float32 xyz at byte 0 (the raw /velodyne_points cloud).
 * out_records12: filteredLaserCloud as 48-byte PointXYZINormal records (12 floats: x y z 1 | nx ny nz 0 |
 * intensity curvature 0 0), capacity n; out_eigenvalues: l1 >= l2 >= l3 per output point (-1 -1 -1 for a
 * failed plane check); out_candidate: planarity presample flag; out_src_index: index into pts.
 * stats4: points with a ring, PCA failures (skipped), plane-check failures (kept), candidates.
 * Any output may be NULL.  Returns the number of output points. */
int64_t orc_frontend(const void* pts, int64_t n, int32_t stride, const orc_frontend_params* P, float* out_records12,
                     float* out_eigenvalues, uint8_t* out_candidate, int32_t* out_src_index, int64_t* stats4);

#endif
