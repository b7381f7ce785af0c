// odometry_driver.cpp — C++ host side above the C ABI: the L3 driver loop of the reference
// (processData, src/laser_odometry.cpp:416-683) without ROS, written against the adapter
// include/plo/imls_icp_cuda.h exactly the way the reference's loop is written against
// IMLSICPMatcher + solveMotionEstimationProblem.  It exists to show (and test) that the
// drop-in really drops in from C++: the stand-in point / vector / matrix types below have the
// interface subset of pcl::PointXYZINormal, Eigen::Vector3d and Eigen::Matrix4d the loop uses.
//
// usage: odometry_driver <config.json|-> <resident|stepped> <out_poses.txt> <frame0.bin> <frame1.bin> ...
//   frame files: raw float32 records, 48 bytes per point (PointXYZINormal layout)
//   poses: savePoseToFile format (src/saver.cpp:46-54): timestamp tx ty tz qx qy qz qw
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "plo/imls_icp_cuda.h"

// ---- stand-ins for the PCL / Eigen types of the reference ------------------------------
struct PointXYZINormal {   // pcl::PointXYZINormal, include/common.h:17
  float x, y, z, pad0;
  float normal_x, normal_y, normal_z, pad1;
  float intensity, curvature, pad2, pad3;
};
static_assert(sizeof(PointXYZINormal) == 48, "PointXYZINormal must be 48 bytes");
typedef PointXYZINormal PointType;
struct PointCloud {
  std::vector<PointType> points;
  size_t size() const { return points.size(); }
};
typedef std::shared_ptr<PointCloud> CloudPtr;
struct Vector3d {
  double v[3];
  double operator[](int i) const { return v[i]; }
};
struct Matrix4d {
  double m[16];
  Matrix4d() { setIdentity(); }
  void setIdentity() { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.0 : 0.0; }
  double& operator()(int r, int c) { return m[r * 4 + c]; }
  double operator()(int r, int c) const { return m[r * 4 + c]; }
  Matrix4d operator*(const Matrix4d& o) const {
    Matrix4d r;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        double s = 0;
        for (int k = 0; k < 4; ++k) s += m[i * 4 + k] * o.m[k * 4 + j];
        r.m[i * 4 + j] = s;
      }
    return r;
  }
};

// ---- the few config.json keys the loop reads (a flat scan, no JSON library) --------------
struct Config {
  std::string text;
  bool load(const std::string& path) {
    std::ifstream f(path);
    if (!f) return false;
    std::stringstream ss;
    ss << f.rdbuf();
    text = ss.str();
    return true;
  }
  // value of the first occurrence of "key" after position of "scope" (enough for config.json's layout)
  std::string raw(const std::string& scope, const std::string& key, const std::string& dflt) const {
    size_t p = 0;
    if (!scope.empty()) {   // the occurrence of "scope" that opens an object ("scope": {), not a string value
      const std::string q = "\"" + scope + "\"";
      for (p = text.find(q); p != std::string::npos; p = text.find(q, p + 1)) {
        size_t c = text.find_first_not_of(" \t\r\n", p + q.size());
        if (c == std::string::npos || text[c] != ':') continue;
        c = text.find_first_not_of(" \t\r\n", c + 1);
        if (c != std::string::npos && text[c] == '{') break;
      }
    }
    if (p == std::string::npos) return dflt;
    p = text.find("\"" + key + "\"", p);
    if (p == std::string::npos) return dflt;
    p = text.find(':', p);
    if (p == std::string::npos) return dflt;
    size_t e = text.find_first_of(",}\n", p + 1);
    std::string v = text.substr(p + 1, e - p - 1);
    size_t a = v.find_first_not_of(" \t\""), b = v.find_last_not_of(" \t\"\r");
    return a == std::string::npos ? dflt : v.substr(a, b - a + 1);
  }
  double num(const std::string& scope, const std::string& key, double d) const {
    std::string v = raw(scope, key, "");
    return v.empty() ? d : atof(v.c_str());
  }
  bool flag(const std::string& scope, const std::string& key, bool d) const {
    std::string v = raw(scope, key, "");
    return v.empty() ? d : (v == "true");
  }
};

static CloudPtr loadCloud(const std::string& path) {
  CloudPtr c(new PointCloud);
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) throw std::runtime_error("cannot open " + path);
  const std::streamsize bytes = f.tellg();
  f.seekg(0);
  c->points.resize((size_t)bytes / sizeof(PointType));
  f.read(reinterpret_cast<char*>(c->points.data()), (std::streamsize)(c->points.size() * sizeof(PointType)));
  return c;
}

// savePoseToFile, src/saver.cpp:46-54
static void savePose(std::ofstream& f, const Matrix4d& T, double ts) {
  const double tr = T(0, 0) + T(1, 1) + T(2, 2);
  double qw = std::sqrt(std::max(0.0, 1.0 + tr)) / 2.0, qx, qy, qz;
  if (qw > 1e-8) {
    qx = (T(2, 1) - T(1, 2)) / (4 * qw); qy = (T(0, 2) - T(2, 0)) / (4 * qw); qz = (T(1, 0) - T(0, 1)) / (4 * qw);
  } else {
    qx = std::sqrt(std::max(0.0, 1.0 + T(0, 0) - T(1, 1) - T(2, 2))) / 2.0;
    qy = std::sqrt(std::max(0.0, 1.0 - T(0, 0) + T(1, 1) - T(2, 2))) / 2.0;
    qz = std::sqrt(std::max(0.0, 1.0 - T(0, 0) - T(1, 1) + T(2, 2))) / 2.0;
  }
  char buf[256];
  snprintf(buf, sizeof(buf), "%.6f %.6f %.6f %.6f %.6f %.6f %.6f %.6f\n", ts, T(0, 3), T(1, 3), T(2, 3), qx, qy, qz, qw);
  f << buf;
}

int main(int argc, char** argv) {
  if (argc < 6) {
    std::cerr << "usage: odometry_driver <config.json|-> <resident|stepped> <out_poses.txt> <frame0.bin> <frame1.bin> ...\n";
    return 2;
  }
  try {
    Config cfg;
    if (std::string(argv[1]) != "-" && !cfg.load(argv[1])) throw std::runtime_error("cannot read config");
    const bool resident = std::string(argv[2]) == "resident";
    // src/laser_odometry.cpp:487-507 — the reference's own key lookups, done once
    const std::string matching_method = cfg.raw("matching_method", "method", "IMLS");
    if (matching_method != "IMLS" && matching_method != "IMLS_CUDA") throw std::runtime_error("Invalid MATCHING_METHOD!");
    const std::string solve_method = cfg.raw("solve_method", "method", "WeightedLS_CUDA");
    const int iterations = (int)cfg.num("solve_method", "iterations", 30);
    const double h = cfg.num("IMLS", "h", 1), r = cfg.num("IMLS", "r", 3);
    const bool is_get_normals = cfg.flag("get_normals", "enabled", true);
    const double r_normal = cfg.num("get_normals", "r_normal", 1);
    const int search_number_normal = (int)cfg.num("get_normals", "search_number_normal", 10);
    const int search_number = (int)cfg.num("IMLS function", "search_number", 20);
    const bool normal_angle_constraint = cfg.flag("normal_angle_constraint", "enabled", true);
    const double angle_diff_threshold = cfg.num("normal_angle_constraint", "angle_diff_threshold", 30);
    const int correspond_number = (int)cfg.num("matching_method", "correspond_number", 6);
    const double delta_dist_threshold = cfg.num("solve_method", "delta_dist_threshold", 0.001);
    const double delta_angle_threshold = cfg.num("solve_method", "delta_angle_threshold", 0.0001745353);
    const bool transform_normal = cfg.flag("laser_odometry", "transform_normal", false);

    plo::IMLSICPMatcherCUDA<PointType> matcher;   // :489 (the reference rebuilds it per frame; one context is reused here)
    matcher.setParameters(iterations, h, r, r_normal, 0.8, false, is_get_normals, false, 50, 0.2, 0.6, search_number_normal,
                          search_number, normal_angle_constraint, angle_diff_threshold, "");   // :514-518
    matcher.setLoopParameters(transform_normal, correspond_number, delta_dist_threshold, delta_angle_threshold);
    // :606 -> :173-275, parameters of :196-244; unknown / out-of-scope strings throw
    matcher.setSolveMethod(solve_method, cfg.num("LS", "threshold", 0.02), cfg.num("RANSAC", "distance_threshold", 0.8),
                           cfg.num("RANSAC", "huber_threshold", 0.648), cfg.raw("RANSAC", "final_solve_method", "DRPM"),
                           (int)cfg.num("RANSAC", "max_iterations", 5000), cfg.num("RANSAC", "min_inliers_percentage", 0.95),
                           cfg.num("RANSAC", "DRPM_threshold", 0.05), cfg.num("RANSAC", "DRPM_stdev_points", 0.02),
                           cfg.num("RANSAC", "DRPM_stdev_normals", 0.05),
                           cfg.num("RANSAC", "LS_threshold", cfg.num("LS", "threshold", 0.02)));
    // solveMotionEstimationProblem(), src/laser_odometry.cpp:173-275, for the stepped mode: the reference's own
    // dispatch on the config string, each branch with the reference's argument list on std::vector<Vector3d>
    auto solveMotionEstimationProblem = [&](std::vector<Vector3d>& in_cloud_vec, std::vector<Vector3d>& ref_cloud_vec,
                                            std::vector<Vector3d>& ref_normal, Matrix4d& deltaTrans) -> bool {
      plo_ctx* ctx = matcher.context();
      if (solve_method == "WeightedLS_CUDA" || solve_method == "Weighted LS") {
        std::vector<double> no_weights;
        return plo::SolveMotionEstimationProblemWeightedLS_CUDA(ctx, in_cloud_vec, ref_cloud_vec, ref_normal, deltaTrans, no_weights, "");
      }
      if (solve_method == "LS" || solve_method == "LS_CUDA")                                           // :190-197
        return plo::SolveMotionEstimationProblemLS_CUDA(ctx, in_cloud_vec, ref_cloud_vec, ref_normal, deltaTrans, "",
                                                        cfg.num("LS", "threshold", 0.02));
      if (solve_method == "RANSAC")                                                                    // :199-226
        return plo::SolveMotionEstimationProblemRANSAC_CUDA(
            ctx, in_cloud_vec, ref_cloud_vec, ref_normal, deltaTrans, "", (int)cfg.num("RANSAC", "max_iterations", 5000),
            cfg.num("RANSAC", "distance_threshold", 0.8), cfg.num("RANSAC", "min_inliers_percentage", 0.95),
            cfg.num("RANSAC", "huber_threshold", 0.648), cfg.raw("RANSAC", "final_solve_method", "DRPM"),
            cfg.num("RANSAC", "LS_threshold", cfg.num("LS", "threshold", 0.02)), cfg.num("RANSAC", "DRPM_threshold", 0.05),
            cfg.num("RANSAC", "DRPM_stdev_points", 0.02), cfg.num("RANSAC", "DRPM_stdev_normals", 0.05));
      std::cerr << "Invalid SOLVE_METHOD!" << std::endl;                                               // :271
      return false;
    };

    std::ofstream poses(argv[3]);
    Matrix4d prevLaserPose;   // :48-57 globals
    CloudPtr accumulatedTargetCloud;
    int frameCount = 0;
    for (int a = 4; a < argc; ++a) {
      CloudPtr filteredLaserCloud = loadCloud(argv[a]);
      CloudPtr flatCloud = filteredLaserCloud;   // source = the full cloud (BASELINE configs)
      if (frameCount != 0) {                                     // :478
        Matrix4d rPose;                                          // :484-485
        matcher.setSourcePointCloud(flatCloud);                  // :509
        matcher.setTargetPointCloud(accumulatedTargetCloud);     // :510
        int iters = 0;
        if (resident) {
          Matrix4d cov;
          matcher.Match(rPose, cov, "");
          iters = matcher.lastRegistration().iters;
        } else {
          for (int i = 0; i < iterations; i++) {                 // :524
            CloudPtr in_cloud(new PointCloud), ref_cloud(new PointCloud);
            matcher.ProjSourcePtToSurface(rPose, in_cloud, ref_cloud, "", i);   // :527-559
            if ((int)in_cloud->size() < correspond_number || (int)ref_cloud->size() < correspond_number) break;   // :570-576
            std::vector<Vector3d> in_cloud_vec, ref_cloud_vec, ref_normal;      // getXYZ / getNormals, :595-599
            for (const auto& p : in_cloud->points) in_cloud_vec.push_back({{p.x, p.y, p.z}});
            for (const auto& p : ref_cloud->points) {
              ref_cloud_vec.push_back({{p.x, p.y, p.z}});
              ref_normal.push_back({{p.normal_x, p.normal_y, p.normal_z}});
            }
            Matrix4d deltaTrans;
            bool flag = solveMotionEstimationProblem(in_cloud_vec, ref_cloud_vec, ref_normal, deltaTrans);   // :609
            if (!flag) break;                                                   // :611-616
            rPose = deltaTrans * rPose;                                         // :619
            ++iters;
            const double deltaDist = std::sqrt(deltaTrans(0, 3) * deltaTrans(0, 3) + deltaTrans(1, 3) * deltaTrans(1, 3) +
                                               deltaTrans(2, 3) * deltaTrans(2, 3));                                   // :628-632
            double cos_theta = (deltaTrans(0, 0) + deltaTrans(1, 1) + deltaTrans(2, 2) - 1.0) / 2.0;                  // :636
            cos_theta = std::min(1.0, std::max(cos_theta, -1.0));
            if (deltaDist < delta_dist_threshold && std::acos(cos_theta) < delta_angle_threshold) break;              // :643-646
          }
        }
        Matrix4d nowPose = prevLaserPose * rPose;                // :652
        prevLaserPose = nowPose;
        std::cout << "frame " << frameCount << " iterations " << iters << std::endl;
      }
      savePose(poses, prevLaserPose, frameCount / 10.0);         // :658
      accumulatedTargetCloud = filteredLaserCloud;               // :668-670, max_queue_size = 1
      frameCount++;
    }
  } catch (const std::exception& e) {
    std::cerr << "odometry_driver: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
