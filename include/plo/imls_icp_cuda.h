/*
 * imls_icp_cuda.h — C++ adapter over plo_c_api.h that mirrors the reference's matcher and
 * solver surface, so that src/laser_odometry.cpp can switch back-ends by changing two type /
 * function names:
 *
 *   class IMLSICPMatcherCUDA            <->  class IMLSICPMatcher   (include/imls_icp.h:45-147)
 *   SolveMotionEstimationProblemWeightedLS_CUDA
 *                                       <->  SolveMotionEstimationProblemWeightedLS
 *                                            (include/solver.h:92-98, src/solver.cpp:168-220)
 *   SolveMotionEstimationProblem{LS,RANSAC,DRPM}_CUDA
 *                                       <->  SolveMotionEstimationProblem{LS,RANSAC,DRPM}
 *                                            (include/solver.h:84-90, :100-114, :129-139)
 *
 * Header-only, C++11, no dependency beyond the C ABI: it is templated on the point / vector /
 * matrix types so that it compiles here without PCL / Eigen and, in the reference tree, with
 * pcl::PointCloud<pcl::PointXYZINormal> and Eigen::Vector3d / Eigen::Matrix4d.
 *
 * Requirements on the template arguments
 *   PointT  : trivially copyable record, float x,y,z at byte 0 and float normal_x/y/z at byte 16
 *             (pcl::PointXYZINormal, 48 bytes — include/common.h:17)
 *   CloudT  : has `.points` (contiguous std::vector<PointT>-like: data(), size(), resize(), clear())
 *   Vec3T   : operator()(int) or operator[](int) giving double components (Eigen::Vector3d)
 *   Mat4T   : operator()(row, col) assignable double (Eigen::Matrix4d)
 *
 * Behavioural differences from the reference that the boundary forces:
 *   - ProjSourcePtToSurface does not erase from a caller-owned shared cloud behind the caller's
 *     back: it rewrites `in_cloud` with the surviving (transformed) points, exactly what the
 *     reference's in-place erase leaves there, and fills `out_cloud`; `srcIndices()` tells which
 *     source points survived.
 *   - The per-iteration source transform (src/laser_odometry.cpp:527-549) is done on the device:
 *     pass the current rPose instead of a pre-transformed copy of the cloud.
 *   - Errors throw std::runtime_error with plo_last_error() (the reference only prints).
 */
#ifndef PLO_IMLS_ICP_CUDA_H
#define PLO_IMLS_ICP_CUDA_H

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "plo/plo_c_api.h"

namespace plo {

inline void check(plo_ctx* ctx, int rc, const char* what) {
  if (rc != PLO_OK) throw std::runtime_error(std::string(what) + ": " + plo_last_error(ctx));
}

template <typename PointT>
class IMLSICPMatcherCUDA {
  static_assert(sizeof(PointT) >= 28, "PointT must hold xyz at byte 0 and a normal at byte 16");

 public:
  explicit IMLSICPMatcherCUDA(int device = 0) : ctx_(nullptr) {
    int rc = plo_create(device, &ctx_);
    if (rc != PLO_OK) throw std::runtime_error(std::string("plo_create: ") + plo_last_error(nullptr));
    plo_default_params(&params_);
  }
  ~IMLSICPMatcherCUDA() { plo_destroy(ctx_); }
  IMLSICPMatcherCUDA(const IMLSICPMatcherCUDA&) = delete;
  IMLSICPMatcherCUDA& operator=(const IMLSICPMatcherCUDA&) = delete;

  /* include/imls_icp.h:56, src/imls_icp.cpp:74-78 */
  template <typename CloudPtr>
  void setSourcePointCloud(const CloudPtr& cloud) {
    check(ctx_, plo_set_source(ctx_, cloud->points.data(), (int64_t)cloud->points.size(), (int32_t)sizeof(PointT)),
          "setSourcePointCloud");
  }
  /* include/imls_icp.h:58, src/imls_icp.cpp:80-103 */
  template <typename CloudPtr>
  void setTargetPointCloud(const CloudPtr& cloud) {
    check(ctx_, plo_set_target(ctx_, cloud->points.data(), (int64_t)cloud->points.size(), (int32_t)sizeof(PointT)),
          "setTargetPointCloud");
  }

  /* include/imls_icp.h:62-66, src/imls_icp.cpp:146-168 — same 16 arguments, same order */
  void setParameters(int _iter, double _h, double _r, double _r_normal, double /*_r_proj*/, bool _useTensorVoting,
                     bool _isGetNormals, bool _useProjectedDistance, int /*_tensor_k*/, double /*_tensor_sigma*/,
                     double /*_tensor_distance_threshold*/, int _search_number_normal, int _search_number,
                     bool _normal_angle_constraint, double _angle_diff_threshold, const std::string& /*_output_dir*/) {
    if (_useTensorVoting) throw std::runtime_error("use_tensor_voting: out of scope of the CUDA path");
    if (_useProjectedDistance) throw std::runtime_error("use_projected_distance: out of scope of the CUDA path");
    params_.iterations = _iter;
    params_.h = _h;
    params_.r = _r;
    params_.r_normal = _r_normal;
    params_.is_get_normals = _isGetNormals ? 1 : 0;
    params_.search_number_normal = _search_number_normal;
    params_.search_number = _search_number;
    params_.normal_angle_constraint = _normal_angle_constraint ? 1 : 0;
    params_.angle_diff_threshold = _angle_diff_threshold;
    check(ctx_, plo_set_params(ctx_, &params_), "setParameters");
  }
  /* driver-loop keys of config.json that the reference reads inside its loop
   * (src/laser_odometry.cpp:570,606,640-641, laser_odometry.transform_normal :458) */
  void setLoopParameters(bool transform_normal, int correspond_number, double delta_dist_threshold,
                         double delta_angle_threshold, int weight_mode = PLO_W_UNIT) {
    params_.transform_normal = transform_normal ? 1 : 0;
    params_.correspond_number = correspond_number;
    params_.delta_dist_threshold = delta_dist_threshold;
    params_.delta_angle_threshold = delta_angle_threshold;
    params_.weight_mode = weight_mode;
    check(ctx_, plo_set_params(ctx_, &params_), "setLoopParameters");
  }

  /* solve_method.method / solve_method.RANSAC.* of config.json (src/laser_odometry.cpp:606 -> :173-275,
   * :196-244): picks the solver the resident loop (Match) and solveMotionEstimationProblem() below use.
   * Accepted: "WeightedLS_CUDA" / "Weighted LS", "LS" / "LS_CUDA" (trimmed, src/solver.cpp:74-166),
   * "RANSAC" with final_solve_method "LS", "Weighted LS" or "DRPM" (src/solver.cpp:222-385, :366-384, :499-603).
   * ls_threshold = solve_method.LS.threshold (:190); ransac_ls_threshold = solve_method.RANSAC.LS_threshold, the trim
   * fraction of the final "LS" inside RANSAC (:205) -- a key of its own in config.json; < 0: same as ls_threshold.
   * Anything else throws (the reference prints "Invalid SOLVE_METHOD!", :271). */
  void setSolveMethod(const std::string& solve_method, double ls_threshold = 0.02, double ransac_distance_threshold = 0.8,
                      double huber_threshold = 0.648, const std::string& final_solve_method = "DRPM",
                      int ransac_max_iterations = 5000, double ransac_min_inliers_percentage = 0.95,
                      double drpm_threshold = 0.05, double drpm_stdev_points = 0.02, double drpm_stdev_normals = 0.05,
                      double ransac_ls_threshold = -1.0) {
    if (solve_method == "WeightedLS_CUDA" || solve_method == "Weighted LS") {
      params_.solver = PLO_SOLVER_WLS;
    } else if (solve_method == "LS" || solve_method == "LS_CUDA") {
      params_.solver = PLO_SOLVER_LS;
      params_.ls_threshold = ls_threshold;
    } else if (solve_method == "RANSAC") {
      params_.solver = PLO_SOLVER_RANSAC;
      if (final_solve_method == "LS") {
        params_.ransac_final = PLO_FINAL_LS;
        params_.ls_threshold = ransac_ls_threshold >= 0.0 ? ransac_ls_threshold : ls_threshold;
      }
      else if (final_solve_method == "Weighted LS") params_.ransac_final = PLO_FINAL_WLS;
      else if (final_solve_method == "DRPM") params_.ransac_final = PLO_FINAL_DRPM;
      else throw std::runtime_error("plo: unknown RANSAC final_solve_method \"" + final_solve_method + "\"");
      params_.ransac_distance_threshold = ransac_distance_threshold;
      params_.huber_threshold = huber_threshold;
      params_.ransac_max_iterations = ransac_max_iterations;
      params_.ransac_min_inliers_percentage = ransac_min_inliers_percentage;
      params_.drpm_threshold = drpm_threshold;
      params_.drpm_stdev_points = drpm_stdev_points;
      params_.drpm_stdev_normals = drpm_stdev_normals;
    } else {
      throw std::runtime_error("plo: Invalid SOLVE_METHOD! (" + solve_method + ")");
    }
    check(ctx_, plo_set_params(ctx_, &params_), "setSolveMethod");
  }

  /* solveMotionEstimationProblem() (src/laser_odometry.cpp:173-275) for the selected method, on the
   * device-resident pairs of the last ProjSourcePtToSurface (no vector round trip). */
  template <typename Mat4T>
  bool solveMotionEstimationProblem(Mat4T& deltaTrans) {
    double D[16];
    if (params_.solver == PLO_SOLVER_RANSAC) check(ctx_, plo_solve_ransac(ctx_, D, last_probs_, nullptr, nullptr), "solve (RANSAC)");
    else if (params_.solver == PLO_SOLVER_LS) check(ctx_, plo_solve_ls(ctx_, D, nullptr), "solve (LS)");
    else check(ctx_, plo_solve_wls(ctx_, D, nullptr), "solve (Weighted LS)");
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) deltaTrans(r, c) = D[r * 4 + c];
    return true;
  }
  const double* lastDrpmProbabilities() const { return last_probs_; }

  /* bool ImplicitMLSFunction(PointType& x, double& height), include/imls_icp.h:75-76, src/imls_icp.cpp:301-483:
   * x = the (already transformed) point with its normal; false = fewer than 3 usable neighbours (:463-466). */
  bool ImplicitMLSFunction(PointT& x, double& height) {
    const float* f = reinterpret_cast<const float*>(&x);
    const float in[6] = {f[0], f[1], f[2], f[4], f[5], f[6]};
    int32_t ok = 0;
    check(ctx_, plo_imls_height(ctx_, in, 1, &height, &ok), "ImplicitMLSFunction");
    return ok != 0;
  }

  /* Eigen::Vector3d ComputeNormal(std::vector<Eigen::Vector3d>& nearPoints), include/imls_icp.h:84,
   * src/imls_icp.cpp:753-794. */
  template <typename Vec3T>
  Vec3T ComputeNormal(std::vector<Vec3T>& nearPoints) {
    std::vector<double> p(3 * nearPoints.size());
    for (size_t i = 0; i < nearPoints.size(); ++i)
      for (int k = 0; k < 3; ++k) p[3 * i + k] = nearPoints[i][k];
    double n[3];
    check(ctx_, plo_compute_normal(ctx_, p.data(), (int64_t)nearPoints.size(), n), "ComputeNormal");
    Vec3T out;
    for (int k = 0; k < 3; ++k) out[k] = n[k];
    return out;
  }

  /* include/imls_icp.h:79-82, src/imls_icp.cpp:496-745 (+ the transform of
   * src/laser_odometry.cpp:527-549).  in_cloud <- surviving transformed source points,
   * out_cloud <- projected points y with the matched normal; counters as printed at :736-744. */
  template <typename CloudPtr, typename Mat4T>
  void ProjSourcePtToSurface(const Mat4T& rPose, CloudPtr& in_cloud, CloudPtr& out_cloud, const std::string& /*timestamp*/,
                             const int& /*i*/) {
    double T[16];
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) T[r * 4 + c] = rPose(r, c);
    check(ctx_, plo_project(ctx_, T, 0, &last_proj_), "ProjSourcePtToSurface");
    const int64_t cap = last_proj_.n_source > 0 ? last_proj_.n_source : 1;
    sx_.resize(3 * cap); rx_.resize(3 * cap); rn_.resize(3 * cap); si_.resize(cap);
    int64_t n = 0;
    check(ctx_, plo_get_pairs(ctx_, sx_.data(), rx_.data(), rn_.data(), si_.data(), cap, &n), "plo_get_pairs");
    si_.resize(n);
    in_cloud->points.resize((size_t)n);
    out_cloud->points.resize((size_t)n);
    for (int64_t j = 0; j < n; ++j) {
      PointT a, b;
      std::memset(&a, 0, sizeof(PointT));
      std::memset(&b, 0, sizeof(PointT));
      float* af = reinterpret_cast<float*>(&a);
      float* bf = reinterpret_cast<float*>(&b);
      for (int k = 0; k < 3; ++k) { af[k] = sx_[3 * j + k]; bf[k] = rx_[3 * j + k]; bf[4 + k] = rn_[3 * j + k]; }
      af[3] = bf[3] = 1.0f;
      in_cloud->points[(size_t)j] = a;
      out_cloud->points[(size_t)j] = b;
    }
  }

  /* include/imls_icp.h:86-88, src/imls_icp.cpp:804-919 — the loop itself is the driver's
   * (src/laser_odometry.cpp:524-647) and runs resident on the device. */
  template <typename Mat4T>
  bool Match(Mat4T& finalPose, Mat4T& covariance, const std::string& /*timestamp*/) {
    double T[16];
    check(ctx_, plo_register(ctx_, nullptr, T, &last_reg_), "Match");
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) {
        finalPose(r, c) = T[r * 4 + c];
        covariance(r, c) = (r == c) ? 1.0 : 0.0;   /* src/imls_icp.cpp:811 */
      }
    return last_reg_.status == PLO_REG_CONVERGED || last_reg_.status == PLO_REG_MAX_ITERS;
  }

  /* accumulateTargetCloud (src/laser_odometry.cpp:116-136) with the TransformToEnd step the reference left
   * commented out (:118-124): the queue lives on the device, every queued frame moves into the new frame's
   * coordinates with the pose still resident from the last Match(), only `newCloud` crosses the bus, and the
   * result is the target of the next frame (no setTargetPointCloud call).  first_frame: nothing registered yet. */
  template <typename CloudPtr>
  void accumulateTargetCloud(const CloudPtr& newCloud, size_t max_queue_size, bool transform_normal, bool first_frame) {
    check(ctx_, plo_map_push(ctx_, newCloud->points.data(), (int64_t)newCloud->points.size(), (int32_t)sizeof(PointT), nullptr,
                             first_frame ? 0 : 1, (int32_t)max_queue_size, transform_normal ? 1 : 0),
          "accumulateTargetCloud");
  }

  const std::vector<int32_t>& srcIndices() const { return si_; }
  const plo_proj_stats& lastProjection() const { return last_proj_; }
  const plo_reg_stats& lastRegistration() const { return last_reg_; }
  plo_ctx* context() { return ctx_; }

 private:
  plo_ctx* ctx_;
  plo_params params_;
  plo_proj_stats last_proj_{};
  plo_reg_stats last_reg_{};
  double last_probs_[6] = {0, 0, 0, 0, 0, 0};
  std::vector<float> sx_, rx_, rn_;
  std::vector<int32_t> si_;
};

/* include/solver.h:92-98 / src/solver.cpp:168-220 with the same argument list:
 * (source_cloud, ref_cloud, ref_normals, deltaTrans, weights, timestamp).  `weights` may be
 * empty (unit weights).  Always returns true, like the reference (:219). */
template <typename Vec3T, typename Mat4T, typename WeightsT>
bool SolveMotionEstimationProblemWeightedLS_CUDA(plo_ctx* ctx, const std::vector<Vec3T>& source_cloud,
                                                 const std::vector<Vec3T>& ref_cloud, const std::vector<Vec3T>& ref_normals,
                                                 Mat4T& deltaTrans, const WeightsT& weights, const std::string& /*timestamp*/) {
  const size_t n = source_cloud.size();
  std::vector<double> s(3 * n), d(3 * n), nn(3 * n), w;
  for (size_t i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) {
      s[3 * i + k] = source_cloud[i][k];
      d[3 * i + k] = ref_cloud[i][k];
      nn[3 * i + k] = ref_normals[i][k];
    }
  if ((size_t)weights.size() == n && n > 0) {
    w.resize(n);
    for (size_t i = 0; i < n; ++i) w[i] = weights[i];
  }
  double D[16];
  check(ctx, plo_solve_wls_host(ctx, s.data(), d.data(), nn.data(), w.empty() ? nullptr : w.data(), (int64_t)n, D, nullptr),
        "SolveMotionEstimationProblemWeightedLS_CUDA");
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) deltaTrans(r, c) = D[r * 4 + c];
  return true;
}

namespace detail {
template <typename Vec3T>
inline void flatten3(const std::vector<Vec3T>& a, const std::vector<Vec3T>& b, const std::vector<Vec3T>& c, std::vector<double>& s,
                     std::vector<double>& d, std::vector<double>& n) {
  const size_t m = a.size();
  s.resize(3 * m); d.resize(3 * m); n.resize(3 * m);
  for (size_t i = 0; i < m; ++i)
    for (int k = 0; k < 3; ++k) { s[3 * i + k] = a[i][k]; d[3 * i + k] = b[i][k]; n[3 * i + k] = c[i][k]; }
}
template <typename Mat4T>
inline void unflatten16(const double D[16], Mat4T& M) {
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) M(r, c) = D[r * 4 + c];
}
inline void check_rc(plo_ctx* ctx, int rc, const char* what) {
  if (rc != PLO_OK) throw std::runtime_error(std::string("plo: ") + what + ": " + plo_last_error(ctx));
}
}  // namespace detail

/* SolveMotionEstimationProblemLS, include/solver.h:84-90 — same argument list
 * (source_cloud, ref_cloud, ref_normals, deltaTrans, timestamp, threshold). */
template <typename Vec3T, typename Mat4T>
bool SolveMotionEstimationProblemLS_CUDA(plo_ctx* ctx, const std::vector<Vec3T>& source_cloud, const std::vector<Vec3T>& ref_cloud,
                                         const std::vector<Vec3T>& ref_normals, Mat4T& deltaTrans, const std::string& /*timestamp*/,
                                         const double threshold) {
  std::vector<double> s, d, n;
  detail::flatten3(source_cloud, ref_cloud, ref_normals, s, d, n);
  double D[16];
  detail::check_rc(ctx, plo_solve_ls_host(ctx, s.data(), d.data(), n.data(), (int64_t)source_cloud.size(), threshold, D, nullptr),
                   "SolveMotionEstimationProblemLS_CUDA");
  detail::unflatten16(D, deltaTrans);
  return true;
}

/* SolveMotionEstimationProblemRANSAC, include/solver.h:100-114 — same argument list (source_cloud, ref_cloud, ref_normals,
 * deltaTrans, timestamp, max_iterations, distance_threshold, min_inliers_percentage, huber_threshold, final_solve_method,
 * ls_threshold, drpm_threshold, drpm_stdev_points, drpm_stdev_normals) + the seed that replaces the reference's unseeded
 * rand().  Returns false for an unknown final_solve_method, like the reference (src/solver.cpp:377-384). */
template <typename Vec3T, typename Mat4T>
bool SolveMotionEstimationProblemRANSAC_CUDA(plo_ctx* ctx, const std::vector<Vec3T>& source_cloud, const std::vector<Vec3T>& ref_cloud,
                                             const std::vector<Vec3T>& ref_normals, Mat4T& deltaTrans, const std::string& /*timestamp*/,
                                             const int max_iterations, const double distance_threshold,
                                             const double min_inliers_percentage, const double huber_threshold,
                                             const std::string final_solve_method, const double ls_threshold,
                                             const double drpm_threshold, const double drpm_stdev_points,
                                             const double drpm_stdev_normals, const unsigned long long seed = 1ull) {
  plo_params p;
  plo_default_params(&p);
  if (final_solve_method == "LS") p.ransac_final = PLO_FINAL_LS;
  else if (final_solve_method == "Weighted LS") p.ransac_final = PLO_FINAL_WLS;
  else if (final_solve_method == "DRPM") p.ransac_final = PLO_FINAL_DRPM;
  else return false;
  p.ransac_max_iterations = max_iterations;
  p.ransac_distance_threshold = distance_threshold;
  p.ransac_min_inliers_percentage = min_inliers_percentage;
  p.huber_threshold = huber_threshold;
  p.ls_threshold = ls_threshold;
  p.drpm_threshold = drpm_threshold;
  p.drpm_stdev_points = drpm_stdev_points;
  p.drpm_stdev_normals = drpm_stdev_normals;
  p.ransac_seed = seed;
  std::vector<double> s, d, n;
  detail::flatten3(source_cloud, ref_cloud, ref_normals, s, d, n);
  double D[16];
  detail::check_rc(ctx, plo_solve_ransac_host(ctx, s.data(), d.data(), n.data(), (int64_t)source_cloud.size(), &p, D, nullptr, nullptr, nullptr),
                   "SolveMotionEstimationProblemRANSAC_CUDA");
  detail::unflatten16(D, deltaTrans);
  return true;
}

/* SolveMotionEstimationProblemDRPM, include/solver.h:129-139 — same argument list (source_cloud, ref_cloud, ref_normals,
 * deltaTrans, weights, timestamp, threshold, stdev_points, stdev_normals); `weights` may be empty (unit weights). */
template <typename Vec3T, typename Mat4T, typename WeightsT>
bool SolveMotionEstimationProblemDRPM_CUDA(plo_ctx* ctx, const std::vector<Vec3T>& source_cloud, const std::vector<Vec3T>& ref_cloud,
                                           const std::vector<Vec3T>& ref_normals, Mat4T& deltaTrans, const WeightsT& weights,
                                           const std::string& /*timestamp*/, const double threshold, const double stdev_points,
                                           const double stdev_normals) {
  std::vector<double> s, d, n, w;
  detail::flatten3(source_cloud, ref_cloud, ref_normals, s, d, n);
  if ((size_t)weights.size() == source_cloud.size() && !source_cloud.empty()) {
    w.resize(source_cloud.size());
    for (size_t i = 0; i < w.size(); ++i) w[i] = weights[i];
  }
  double D[16];
  detail::check_rc(ctx, plo_solve_drpm_host(ctx, s.data(), d.data(), n.data(), w.empty() ? nullptr : w.data(),
                                            (int64_t)source_cloud.size(), threshold, stdev_points, stdev_normals, D, nullptr),
                   "SolveMotionEstimationProblemDRPM_CUDA");
  detail::unflatten16(D, deltaTrans);
  return true;
}

}  // namespace plo

#endif /* PLO_IMLS_ICP_CUDA_H */
