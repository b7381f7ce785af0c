/*
 * plo_c_api.h — C ABI of the B200-native IMLS-ICP scan-to-map registration path.
 *
 * This is the drop-in boundary: plain C types, opaque handle, no
 * exceptions, no torch types.  Each entry point names the reference interface it
 * replaces (paths relative to the reference root).  The reference has no FFI today
 * (string dispatch on config.json inside src/laser_odometry.cpp:487-568,606); the
 * binding a maintainer would add is shown in INTEGRATION.md, and
 * include/plo/imls_icp_cuda.h is the C++ adapter that mirrors class IMLSICPMatcher.
 *
 * Conventions
 *   - 4x4 transforms are row-major double[16].
 *   - Point records: `stride` bytes per point, float32 xyz at byte 0, float32 normal
 *     at byte 16 (pcl::PointXYZINormal = PointType, include/common.h:17; stride 48).
 *   - Points whose x/y/z is non-finite are stripped on upload, order preserved
 *     (RemoveNANandINFData, src/imls_icp.cpp:58-72); every index this API returns
 *     refers to the stripped cloud, exactly like the reference after its erase().
 *   - All functions return PLO_OK (0) or a negative plo_error; plo_last_error() gives
 *     the message.  A context is bound to one device and one stream, and is not
 *     thread-safe (the reference's matcher is not re-entrant either); use one context
 *     per host thread.  All device memory is owned by the context.
 *   - There is no CPU fallback: without a CUDA device plo_create fails.
 */
#ifndef PLO_C_API_H
#define PLO_C_API_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PLO_API __attribute__((visibility("default")))
#else
#define PLO_API
#endif

typedef struct plo_ctx plo_ctx;

typedef enum plo_error {
  PLO_OK = 0,
  PLO_ERR_INVALID_ARG = -1,
  PLO_ERR_CUDA = -2,
  PLO_ERR_NO_DEVICE = -3,
  PLO_ERR_UNSUPPORTED = -4,
  PLO_ERR_STATE = -5
} plo_error;

/* per-source-point outcome; 1..6 are the six drop counters of
 * src/imls_icp.cpp:506-511,736-744 in their order of declaration */
typedef enum plo_point_status {
  PLO_PT_OK = 0,
  PLO_PT_NO_NORMAL = 1,          /* :612-617  no 1-NN within r            */
  PLO_PT_TOO_FAR = 2,            /* :620-625  d2(1-NN) > h*h              */
  PLO_PT_INVALID_NORMAL = 3,     /* :673-679                              */
  PLO_PT_NORMAL_CONSTRAINT = 4,  /* :681-692                              */
  PLO_PT_MLS_FAIL = 5,           /* :696-701  (< 3 usable IMLS neighbours) */
  PLO_PT_NAN_INF_HEIGHT = 6      /* :703-717                              */
} plo_point_status;

/* outcome of plo_register: how the loop of src/laser_odometry.cpp:524-647 ended */
typedef enum plo_reg_status {
  PLO_REG_CONVERGED = 1,     /* :643-646 */
  PLO_REG_MAX_ITERS = 2,     /* :524 loop bound */
  PLO_REG_TOO_FEW_PAIRS = 3, /* :570-576 */
  PLO_REG_SOLVE_FAILED = 4   /* :611-616 (no pivot at all: every pair had a zero row) */
} plo_reg_status;

typedef enum plo_solver {
  PLO_SOLVER_WLS = 0, /* SolveMotionEstimationProblemWeightedLS, src/solver.cpp:168-220 (default)      */
  PLO_SOLVER_LS = 1,  /* SolveMotionEstimationProblemLS, src/solver.cpp:74-166: LS, then a second LS on  */
                      /* the pairs whose |residual| rank lies in [thr*N, (1-thr)*N] (2 % / 98 % trim)    */
  PLO_SOLVER_RANSAC = 2 /* SolveMotionEstimationProblemRANSAC, src/solver.cpp:222-385 — the config.json   */
                      /* default chain: FPS-3 hypotheses, inlier count, Huber/exp weights at the best     */
                      /* hypothesis, then ransac_final (LS, Weighted LS or DRPM :499-603).  The reference's  */
                      /* unseeded rand() (src/common.cpp:49) is replaced by xorshift64(ransac_seed).        */
} plo_solver;

typedef enum plo_ransac_final {
  PLO_FINAL_LS = 0,   /* "LS": trimmed LS (src/solver.cpp:74-166) on the inliers of the best hypothesis, :366-371 */
  PLO_FINAL_WLS = 1,  /* "Weighted LS" */
  PLO_FINAL_DRPM = 2  /* "DRPM" (config.json default) */
} plo_ransac_final;

typedef enum plo_weight_mode {
  PLO_W_UNIT = 0,     /* plain point-to-plane LS (weights = 1)                         */
  PLO_W_HUBER_EXP = 1 /* RANSAC-final weights of src/solver.cpp:334-364 at T_best = I  */
} plo_weight_mode;

/* IMLSICPMatcher::setParameters (src/imls_icp.cpp:146-168; fields
 * include/imls_icp.h:115-146) + the driver-loop keys of config.json.  The tensor-voting
 * and projected-distance switches are not represented: those branches are out of scope
 *  and requesting them through the adapter is PLO_ERR_UNSUPPORTED. */
typedef struct plo_params {
  int32_t iterations;              /* solve_method.iterations                         30 */
  double h;                        /* IMLS.h: 1-NN gate (d2 <= h*h)                    1 */
  double r;                        /* IMLS.r: search radius                            3 */
  double r_normal;                 /* get_normals.r_normal                             1 */
  int32_t is_get_normals;          /* get_normals.enabled: 1 = use delivered normals   1 */
  int32_t search_number_normal;    /* get_normals.search_number_normal (<= 32)        10 */
  int32_t search_number;           /* "IMLS function".search_number (<= 32)           20 */
  int32_t normal_angle_constraint; /* normal_angle_constraint.enabled                  1 */
  double angle_diff_threshold;     /* normal_angle_constraint.angle_diff_threshold    30 */
  int32_t transform_normal;        /* laser_odometry.transform_normal                  0 */
  int32_t correspond_number;       /* matching_method.correspond_number                6 */
  double delta_dist_threshold;     /* solve_method.delta_dist_threshold            1e-3 */
  double delta_angle_threshold;    /* solve_method.delta_angle_threshold   1.745353e-4 */
  int32_t weight_mode;             /* plo_weight_mode                                  0 */
  double ransac_distance_threshold;/* RANSAC.distance_threshold (PLO_W_HUBER_EXP)    0.8 */
  double huber_threshold;          /* RANSAC.huber_threshold    (PLO_W_HUBER_EXP)  0.648 */
  int32_t solver;                  /* plo_solver                                       0 */
  double ls_threshold;             /* solve_method.LS.threshold (PLO_SOLVER_LS)     0.02 */
  int32_t ransac_max_iterations;   /* RANSAC.max_iterations                         5000 */
  double ransac_min_inliers_percentage; /* RANSAC.min_inliers_percentage            0.95 */
  int32_t ransac_final;            /* plo_ransac_final (RANSAC.final_solve_method)  DRPM */
  double drpm_threshold;           /* RANSAC.DRPM_threshold                         0.05 */
  double drpm_stdev_points;        /* RANSAC.DRPM_stdev_points                      0.02 */
  double drpm_stdev_normals;       /* RANSAC.DRPM_stdev_normals                     0.05 */
  uint64_t ransac_seed;            /* seed of the hypothesis sampler (non-zero)        1 */
} plo_params;

typedef struct plo_proj_stats {
  int64_t n_source;   /* source points after the non-finite strip                       */
  int64_t n_pairs;    /* surviving correspondences ("USED POINTS FINAL")                */
  int64_t dropped[6]; /* the six counters of src/imls_icp.cpp:736-744                   */
} plo_proj_stats;

typedef struct plo_reg_stats {
  int32_t status;      /* plo_reg_status                                                */
  int32_t iters;       /* solves performed                                              */
  int64_t pairs;       /* pairs of the last projection                                  */
  double rms;          /* sqrt(mean(b_i^2)) of the last projection, b_i = n.(y - x)     */
  int64_t dropped[6];  /* counters of the last projection                                */
  double delta_dist;   /* |t| of the last delta  (src/laser_odometry.cpp:628-632)       */
  double delta_angle;  /* angle of the last delta (:636-638)                            */
  int32_t rank;        /* pivots used by the last 6x6 solve                             */
  int32_t reserved;
} plo_reg_stats;

/* ---- lifetime ------------------------------------------------------------------- */
PLO_API int plo_create(int device, plo_ctx** out);
PLO_API void plo_destroy(plo_ctx* ctx);
/* message of the last failing call on ctx (ctx may be NULL: last plo_create failure) */
PLO_API const char* plo_last_error(const plo_ctx* ctx);
PLO_API int plo_version(void);
/* run everything on a caller-owned CUDA stream (cudaStream_t passed as void*) */
PLO_API int plo_set_stream(plo_ctx* ctx, void* cuda_stream);
PLO_API int plo_synchronize(plo_ctx* ctx);
/* Device-resident inputs (plo_*_device, on_device batches) are read on the context's stream.  If another stream
 * produced them, record an event there and pass it here first: the context's stream then waits for it (the Python
 * Context does this for torch tensors).  Host inputs: the bytes are copied asynchronously when the buffer is pinned --
 * leave it untouched until a call that returns results (plo_register, plo_get_*, plo_synchronize) has returned. */
PLO_API int plo_stream_wait_event(plo_ctx* ctx, void* cuda_event);

/* ---- parameters: IMLSICPMatcher::setParameters, src/imls_icp.cpp:146-168 --------- */
PLO_API void plo_default_params(plo_params* p);
PLO_API int plo_set_params(plo_ctx* ctx, const plo_params* p);

/* ---- clouds ----------------------------------------------------------------------
 * plo_set_target  == IMLSICPMatcher::setTargetPointCloud, src/imls_icp.cpp:80-103:
 *   strips non-finite points, builds the spatial index in HBM (replaces
 *   Nabo::NNSearchD::createKDTreeLinearHeap, :101).
 * plo_set_source  == IMLSICPMatcher::setSourcePointCloud, src/imls_icp.cpp:74-78.
 * The *_device variants take pointers to records already resident in device memory
 * (same layout); the records are consumed before the call's work completes on the
 * context's stream, the caller keeps ownership. */
PLO_API int plo_set_target(plo_ctx* ctx, const void* host_pts, int64_t n, int32_t stride_bytes);
PLO_API int plo_set_source(plo_ctx* ctx, const void* host_pts, int64_t n, int32_t stride_bytes);
PLO_API int plo_set_target_device(plo_ctx* ctx, const void* dev_pts, int64_t n, int32_t stride_bytes);
PLO_API int plo_set_source_device(plo_ctx* ctx, const void* dev_pts, int64_t n, int32_t stride_bytes);
/* sizes after the strip (synchronises the stream) */
PLO_API int64_t plo_target_size(plo_ctx* ctx);
PLO_API int64_t plo_source_size(plo_ctx* ctx);

/* ---- one matcher pass -------------------------------------------------------------
 * plo_project == the per-iteration source transform of src/laser_odometry.cpp:527-549
 * followed by IMLSICPMatcher::ProjSourcePtToSurface (src/imls_icp.cpp:496-745) with
 * ImplicitMLSFunction (:301-483) fused in.  T = current rPose.  Results stay on the
 * device; stats may be NULL (no synchronisation then).  With `hooks` != 0 the per-query
 * parity data of plo_get_neighbors / plo_get_query_results is recorded as well. */
PLO_API int plo_project(plo_ctx* ctx, const double T[16], int32_t hooks, plo_proj_stats* stats);
/* correspondences of the last plo_project, compacted in source order — the contents of
 * in_cloud / out_cloud after ProjSourcePtToSurface returns (src/imls_icp.cpp:719-731).
 * float32 triplets; src_idx = index of each pair's point in the stripped source
 * (replaces the in-place erase).  Any output pointer may be NULL. */
PLO_API int plo_get_pairs(plo_ctx* ctx, float* src_xyz, float* ref_xyz, float* ref_n,
                          int32_t* src_idx, int64_t cap, int64_t* n);
/* parity hooks (need hooks != 0 in the last plo_project): k = search_number.
 * nn_idx[M*k] / nn_d2[M*k]: the knn of src/imls_icp.cpp:372-375 (-1 / +inf padded);
 * nn1_idx[M] / nn1_d2[M]: the 1-NN of :601-609.  Any pointer may be NULL. */
PLO_API int plo_get_neighbors(plo_ctx* ctx, int32_t* nn_idx, double* nn_d2,
                              int32_t* nn1_idx, double* nn1_d2);
/* traversal statistics of the k-NN search per query (hooks != 0): stats3[3*M] =
 * leaves scanned, internal nodes expanded, list insertions — for the roofline analysis */
PLO_API int plo_get_search_stats(plo_ctx* ctx, int32_t* stats3);
/* status[M] (plo_point_status), height[M] = I(x) of :480 (NaN where not computed) */
PLO_API int plo_get_query_results(plo_ctx* ctx, int32_t* status, double* height);
/* normals the matcher uses for the target, n x 3 doubles in stripped-cloud order:
 * the delivered ones, or ComputeNormal (src/imls_icp.cpp:753-794) when
 * is_get_normals == 0 */
PLO_API int plo_get_target_normals(plo_ctx* ctx, double* out);

/* bool IMLSICPMatcher::ImplicitMLSFunction(PointType& x, double& height) (include/imls_icp.h:75-76,
 * src/imls_icp.cpp:301-483) for a batch: xyz_normal6 = n x {x, y, z, nx, ny, nz} float32 (the point as the matcher
 * sees it, i.e. already transformed; its normal feeds the angle filter).  ok[i] = the function's return value
 * (>= 3 usable neighbours); height[i] is NaN where ok[i] == 0.  No 1-NN gates: those belong to ProjSourcePtToSurface. */
PLO_API int plo_imls_height(plo_ctx* ctx, const float* xyz_normal6, int64_t n, double* height, int32_t* ok);
/* Eigen::Vector3d IMLSICPMatcher::ComputeNormal(std::vector<Eigen::Vector3d>&) (include/imls_icp.h:84,
 * src/imls_icp.cpp:753-794): mean, population covariance, unit eigenvector of the smallest eigenvalue; like the
 * reference, no sign disambiguation (the matcher's own PCA pass orients +z, deviation D2). */
PLO_API int plo_compute_normal(plo_ctx* ctx, const double* pts3, int64_t n, double normal[3]);

/* ---- the other reference-shaped solver entry points: host vectors in, 4x4 out ----------------------------
 * src / ref / nrm are n x 3 doubles (std::vector<Eigen::Vector3d>::data()), delta is row-major.  The pairs are staged
 * in the float32 arrays the device solvers work on, so every coordinate must be float32-representable -- what
 * getXYZ / getNormals (include/common.h:51-75) produce at the reference's call site (src/laser_odometry.cpp:595-599);
 * anything else is PLO_ERR_UNSUPPORTED (plo_solve_wls_host takes arbitrary doubles).  These calls use the context's
 * per-query arrays: the last projection's pairs and temporal state are gone afterwards, the clouds stay.
 *
 * SolveMotionEstimationProblemLS (include/solver.h:84-90, src/solver.cpp:74-166): `threshold` = trim fraction. */
PLO_API int plo_solve_ls_host(plo_ctx* ctx, const double* src, const double* ref, const double* nrm, int64_t n,
                              double threshold, double delta[16], int32_t* rank);
/* SolveMotionEstimationProblemRANSAC (include/solver.h:100-114, src/solver.cpp:222-385).  `ransac` carries the
 * arguments of the reference's signature: ransac_max_iterations, ransac_distance_threshold,
 * ransac_min_inliers_percentage, huber_threshold, ransac_final (final_solve_method), ls_threshold, drpm_threshold,
 * drpm_stdev_points, drpm_stdev_normals, and ransac_seed (the reference's unseeded rand()); its other fields are
 * ignored.  probs: the six DRPM non-degeneracy probabilities (zero unless final = DRPM). */
PLO_API int plo_solve_ransac_host(plo_ctx* ctx, const double* src, const double* ref, const double* nrm, int64_t n,
                                  const plo_params* ransac, double delta[16], double probs[6], int64_t* inliers,
                                  int32_t* hypotheses);
/* SolveMotionEstimationProblemDRPM (include/solver.h:129-139, src/solver.cpp:499-603, include/degeneracy.h:14-131):
 * weights w (n doubles, NULL = unit) are used as they are. */
PLO_API int plo_solve_drpm_host(plo_ctx* ctx, const double* src, const double* ref, const double* nrm, const double* w,
                                int64_t n, double threshold, double stdev_points, double stdev_normals, double delta[16],
                                double probs[6]);

/* ---- solver -----------------------------------------------------------------------
 * plo_solve_wls == SolveMotionEstimationProblemWeightedLS (src/solver.cpp:168-220,
 * include/solver.h:92-98) on the device-resident pairs of the last plo_project with the
 * context's weight_mode.  delta = deltaTrans (row-major).  Returns PLO_OK also when the
 * system is rank-deficient (the reference always returns true); *rank (nullable) tells. */
PLO_API int plo_solve_wls(plo_ctx* ctx, double delta[16], int32_t* rank);
/* plo_solve_ls == SolveMotionEstimationProblemLS (src/solver.cpp:74-166, include/solver.h:84-90) on the
 * device-resident pairs of the last plo_project, trim fraction = params.ls_threshold.  Ties in the
 * |residual| order are broken by pair index (std::sort there is unstable). */
PLO_API int plo_solve_ls(plo_ctx* ctx, double delta[16], int32_t* rank);
/* plo_solve_ransac == SolveMotionEstimationProblemRANSAC (src/solver.cpp:222-385) with the context's
 * RANSAC / DRPM parameters on the device-resident pairs of the last plo_project.  probs[6] (nullable)
 * receives DRPM's non-degeneracy probabilities (zeros for the Weighted-LS tail); inliers / hypotheses
 * (nullable) the best hypothesis' inlier count and the number of hypotheses evaluated. */
PLO_API int plo_solve_ransac(plo_ctx* ctx, double delta[16], double probs[6], int64_t* inliers, int32_t* hypotheses);
/* same solver, reference-shaped inputs: host arrays of n x 3 doubles + n weights
 * (NULL = unit) — the argument list of SolveMotionEstimationProblemWeightedLS */
PLO_API int plo_solve_wls_host(plo_ctx* ctx, const double* src, const double* ref,
                               const double* nrm, const double* w, int64_t n,
                               double delta[16], int32_t* rank);
/* parity hook: the reduced normal equations of the last solve: H (21 upper-triangular
 * entries, row-major), g (6), sum of weights, sum w*b*b, pair count */
PLO_API int plo_get_normal_equations(plo_ctx* ctx, double H21[21], double g6[6],
                                     double* sw, double* swbb, int64_t* count);

/* ---- resident registration loop ---------------------------------------------------
 * plo_register == the ICP loop of src/laser_odometry.cpp:484-485,524-647 (and the shape
 * of IMLSICPMatcher::Match, src/imls_icp.cpp:804-919): project -> reduce -> 6x6 solve ->
 * rPose = delta * rPose -> convergence test, all on the device, one host read-back at
 * the end.  T0 = initial rPose (the reference uses identity; NULL = identity). */
PLO_API int plo_register(plo_ctx* ctx, const double T0[16], double T[16], plo_reg_stats* stats);

/* ---- batched mode -----------------------------------------------------------------
 * Registers `count` independent (source, target) pairs one after the other on the
 * context's stream with a single synchronisation at the end: pair i = host (or device,
 * if on_device != 0) record arrays sources[i] (n_src[i]) against targets[i] (n_tgt[i]).
 * T_out[16*count], stats_out[count] (nullable).  This is the unit that is sharded
 * across GPUs (one context per GPU). */
PLO_API int plo_register_batch(plo_ctx* ctx, int32_t count,
                               const void* const* sources, const int64_t* n_src,
                               const void* const* targets, const int64_t* n_tgt,
                               int32_t stride_bytes, int32_t on_device,
                               double* T_out, plo_reg_stats* stats_out);

/* ---- device-resident local map --------------------------------
 * accumulateTargetCloud (src/laser_odometry.cpp:116-136) with the TransformToEnd step (:88-114) that the
 * reference left commented out (:118-124), so that a queue of more than one frame is geometrically
 * consistent.  plo_map_push: (1) every queued frame moves from the previous frame's coordinates into the
 * new frame's, p' = R^T (p - t) in double with a float32 store (normals too when transform_normals != 0),
 * where [R t] = T_last_curr is the rPose of the registration just done (x_prev = R x_cur + t; NULL =
 * identity), or — pose_from_last_register != 0 — the pose still resident on the device after
 * plo_register, which keeps register -> push -> register free of host round trips; (2) frames beyond
 * max_queue (config.json laser_odometry.max_queue_size) drop out, oldest first; (3) the new frame is
 * appended; (4) the result becomes the target (index rebuilt on the device, as plo_set_target).
 * Only the new frame ever crosses the bus.  plo_map_reset empties the queue. */
PLO_API int plo_map_reset(plo_ctx* ctx);
PLO_API int plo_map_push(plo_ctx* ctx, const void* host_pts, int64_t n, int32_t stride_bytes, const double T_last_curr[16],
                         int32_t pose_from_last_register, int32_t max_queue, int32_t transform_normals);
PLO_API int plo_map_push_device(plo_ctx* ctx, const void* dev_pts, int64_t n, int32_t stride_bytes,
                                const double T_last_curr[16], int32_t pose_from_last_register, int32_t max_queue,
                                int32_t transform_normals);
/* queued frames / points (before the non-finite strip) */
PLO_API int plo_map_info(plo_ctx* ctx, int32_t* frames, int64_t* points);
/* the map as 32-byte records {x,y,z,0,nx,ny,nz,0} (floats), oldest frame first; cap in records */
PLO_API int plo_map_get(plo_ctx* ctx, float* records8, int64_t cap);

/* ---- front-end normals + presample ------------------------------
 * The stage of laserCloudHandler (src/scan_registration.cpp) that turns the raw /velodyne_points cloud into
 * /laser_cloud_filtered — the cloud with normals that plo_set_target / plo_set_source consume — at the
 * config.json defaults (format "pointcloud", method "pca", neighbor_scan "kdtree", presample
 * "geometric_features"): range gate (:862-863), ring assignment (:938-1016), intensity = ring + scanPeriod *
 * relTime (:1018-1042), per-ring clouds (:1043), windowed PCA over three rings (computeNormalPCA :158-229 in the
 * loop of :1162-1229) with the plane check (:137-156), planarity presample (computeGeometricFeatures :279-327,
 * :1481-1489).  The sampling stage behind it (:1493 samplePointCloud -> /laser_cloud_flat) is out of scope.
 * Float32 arithmetic as in the reference; the un-vendored pieces (Eigen reductions / eigen-solver, FLANN ties,
 * libm overloads) are defined as in oracle/plo_oracle_frontend.c.  The result stays on the device. */
typedef struct plo_frontend_params {
  int32_t n_scans;                 /* launch parameter scan_line: 16, 32 or 64                  64 */
  float min_range, max_range;      /* MINIMUM_RANGE / MAXIMUM_RANGE (:62-63)             0.5 / 120 */
  float scan_period;               /* :55                                                       0.1 */
  int32_t window_size, iter_step;  /* compute_normal_method.pca                               3 / 1 */
  float knn_distance_threshold;    /* pca.knn_distance_threshold (SQUARED metres, as FLANN's)    10 */
  float plane_distance_threshold;  /* pca.plane_constraint.distance_threshold                  0.02 */
  float valid_points_threshold;    /* pca.plane_constraint.valid_points_threshold               0.8 */
  int32_t use_all_points;          /* model.use_all_points                                        1 */
  float planarity_threshold;       /* presample_method.geometric_features.planarity_threshold  0.05 */
} plo_frontend_params;

typedef struct plo_frontend_stats {
  int64_t n_out;           /* points of filteredLaserCloud                                   */
  int64_t gated;           /* points after removeNaN + range gate                            */
  int64_t ringed;          /* of those, points with a ring ("points size", :1060)            */
  int64_t pca_failures;    /* "pca failure points size" (:1228): skipped                     */
  int64_t plane_failures;  /* "plane check failure points size" (:1229): kept, never presampled */
  int64_t candidates;      /* "Presampled points size" (:1491)                               */
} plo_frontend_stats;

PLO_API void plo_frontend_default_params(plo_frontend_params* p);
/* pts: n records of stride_bytes, float32 xyz at byte 0.  stats (nullable) costs the one synchronisation. */
PLO_API int plo_frontend(plo_ctx* ctx, const void* host_pts, int64_t n, int32_t stride_bytes, const plo_frontend_params* p,
                         plo_frontend_stats* stats);
PLO_API int plo_frontend_device(plo_ctx* ctx, const void* dev_pts, int64_t n, int32_t stride_bytes,
                                const plo_frontend_params* p, plo_frontend_stats* stats);
/* results of the last run; every pointer nullable; cap in points.  records12: 48-byte PointXYZINormal records
 * (x y z 1 | nx ny nz 0 | intensity curvature 0 0); eigenvalues3: l1 >= l2 >= l3 (-1 -1 -1 after a failed plane
 * check); candidate: planarity presample flag; src_index: index into the input cloud. */
PLO_API int plo_frontend_get(plo_ctx* ctx, float* records12, float* eigenvalues3, uint8_t* candidate, int32_t* src_index,
                             int64_t cap);
/* device pointer to the 48-byte records of the last run and their count (for plo_set_source_device /
 * plo_set_target_device / plo_map_push_device: the cloud never leaves the GPU); valid until the next run */
PLO_API int plo_frontend_device_records(plo_ctx* ctx, const void** dev_records, int64_t* n);

/* ---- introspection (bench / tests) ------------------------------------------------ */
/* kernels launched by this context since creation (bench.py's gpu_launches claim) */
PLO_API int64_t plo_launch_count(const plo_ctx* ctx);
/* device milliseconds of the last index build / last registration loop, measured with
 * CUDA events on the context's stream (0 if not measured yet) */
PLO_API int plo_last_timings(plo_ctx* ctx, float* ms_index_build, float* ms_register);
/* tuning / test knobs (no reference counterpart), also read from the environment once, at plo_create:
 *   "chunk"        consecutive source points a warp takes per tree-walk chunk (0 = device-side policy; PLO_CHUNK)
 *   "group"        queries per warp and round on the candidate-tile path (<= 32; 0 = sized from the cloud)
 *   "no_graph"     enqueue every iteration instead of the loop kernel / conditional CUDA graph (ncu cannot profile
 *                  kernel nodes of such graphs; PLO_NO_GRAPH)
 *   "loop_kernel"  0: the weighted-LS loop as a CUDA graph of k_project + k_reduce_solve launches instead of ONE
 *                  k_register_loop launch (PLO_LOOP_KERNEL); "fuse" 0: stand-alone reduce / solve kernels (PLO_FUSE)
 *   "force_warm"   stepped projections start in the settled regime: candidate tiles are written and consumed -- lets
 *                  the parity tests exercise the tile path through plo_project
 * All forms give bitwise the same poses.  PLO_NO_CHAIN=1 in the environment: ordinary stream launches in the index
 * build instead of programmatically dependent ones. */
PLO_API int plo_set_tuning(plo_ctx* ctx, const char* name, int32_t value);
/* per-kernel CUDA-event timing inside plo_register (off by default): when enabled every projection launch of the loop
 * (enqueue-all form) is bracketed by an event pair on the context's stream.  plo_last_kernel_timings returns the mean
 * device ms of the launches that did work in the last plo_register and how many those were. */
PLO_API int plo_set_profiling(plo_ctx* ctx, int32_t enabled);
PLO_API int plo_last_kernel_timings(plo_ctx* ctx, float* ms_project_mean, int32_t* n_project);
/* the same launches one by one (ICP iteration i of the last plo_register -> ms_each[i], at most cap
 * and at most 64 entries): shows how the projection gets cheaper as the pose settles */
PLO_API int plo_last_project_times(plo_ctx* ctx, float* ms_each, int32_t cap, int32_t* n_project);
/* per projection of the last plo_register: how many queries the tile path could not answer from their candidate
 * tile and handed to the tree walk (-1: the settled kernel did not run in that projection) */
PLO_API int plo_last_tile_misses(plo_ctx* ctx, int32_t* misses, int32_t cap, int32_t* n_project);
/* debug builds (-DPLO_LOOP_TIMING) only: phase-boundary timestamps of the resident loop kernel */
PLO_API int plo_debug_loop_stamps(plo_ctx* ctx, unsigned long long* out128);
/* per-kernel timing of one projection pass for the roofline: runs the projection kernel
 * `reps` times at the current pose and returns the mean device ms per launch */
PLO_API int plo_time_project_kernel(plo_ctx* ctx, const double T[16], int32_t reps, float* ms_mean);

#ifdef __cplusplus
}
#endif
#endif /* PLO_C_API_H */
