// ref_cluster.cpp -- ORACLE/_ref (TEST INFRASTRUCTURE ONLY).
//
// PoseEstimator<PointT>::clusterPoses (/root/reference/src/perception/src/PoseEstimator.cpp:106-233) and
// Utils::rotationGeodesicDistance (Utils.cpp:29-32) restated statement for statement ON TOP OF THE REFERENCE TREE'S OWN
// Eigen (src/OpenGR_4pcs/3rdparty/Eigen, 3.3.90): block(), eulerAngles(2,1,0), the 3x3 product, trace() and norm() below
// are Eigen's, only the surrounding control flow is retyped (PoseEstimator.cpp itself needs PCL / yaml-cpp and cannot be
// compiled here).  The perception package builds against the SYSTEM Eigen of its host (unpinned); this pins the
// arithmetic to the one Eigen that ships with the reference.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include <Eigen/Dense>
#include <Eigen/Geometry>

namespace {
struct Hypo { Eigen::Matrix4f pose; float lcp; int id; int src; };
float rotationGeodesicDistance(const Eigen::Matrix3f &R1, const Eigen::Matrix3f &R2) { return std::acos(((R1 * R2).trace() - 1) / 2.0); }
}  // namespace

extern "C" int hop_ref_cluster_poses(const float *poses, const float *scores, const int32_t *ids, int n, float angle_diff, float dist_diff,
                                     const double *symmetry_deg, int32_t *keep_out) {
  if (n <= 0) return 0;
  std::vector<Hypo> hypos(n);
  for (int k = 0; k < n; ++k) {
    hypos[k].pose = Eigen::Map<const Eigen::Matrix4f>(poses + 16 * k);
    hypos[k].lcp = scores[k]; hypos[k].id = ids ? ids[k] : k; hypos[k].src = k;
  }
  std::sort(hypos.begin(), hypos.end(), [](const Hypo &p1, const Hypo &p2) {
    if (p1.lcp > p2.lcp) return true;
    if (p1.lcp < p2.lcp) return false;
    if (p1.id < p2.id) return true;
    if (p1.id > p2.id) return false;
    return false;
  });
  std::vector<Hypo> hypo_tmp = hypos;
  hypos.clear();
  hypos.push_back(hypo_tmp[0]);
  const float radian_thres = angle_diff / 180.0 * M_PI;
  const float x_symmetry = symmetry_deg[0] / 180 * M_PI;
  const float y_symmetry = symmetry_deg[1] / 180 * M_PI;
  const float z_symmetry = symmetry_deg[2] / 180 * M_PI;
  for (size_t i = 1; i < hypo_tmp.size(); i++) {
    bool isnew = true;
    Eigen::Matrix4f cur_pose = hypo_tmp[i].pose;
    for (auto cluster : hypos) {
      Eigen::Vector3f t0 = cluster.pose.block(0, 3, 3, 1);
      Eigen::Vector3f t1 = cur_pose.block(0, 3, 3, 1);
      if ((t0 - t1).norm() >= dist_diff) continue;
      Eigen::Matrix3f R0 = cluster.pose.block(0, 0, 3, 3);
      Eigen::Vector3f rpy = R0.eulerAngles(2, 1, 0);
      float r0 = rpy(2), p0 = rpy(1), y0 = rpy(0);
      Eigen::Matrix3f R1 = cur_pose.block(0, 0, 3, 3);
      Eigen::Vector3f rpy1 = R1.eulerAngles(2, 1, 0);
      float r1 = rpy1(2), p1 = rpy1(1), y1 = rpy1(0);
      float roll_diff = std::abs(r0 - r1), pitch_diff = std::abs(p0 - p1), yaw_diff = std::abs(y0 - y1);
      if (x_symmetry == 0) roll_diff = 0; else if (x_symmetry > 0) roll_diff = std::min(roll_diff, static_cast<float>(x_symmetry) - roll_diff);
      if (y_symmetry == 0) pitch_diff = 0; else if (y_symmetry > 0) pitch_diff = std::min(pitch_diff, static_cast<float>(y_symmetry) - pitch_diff);
      if (z_symmetry == 0) yaw_diff = 0; else if (z_symmetry > 0) yaw_diff = std::min(yaw_diff, static_cast<float>(z_symmetry) - yaw_diff);
      if (pitch_diff <= radian_thres && roll_diff <= radian_thres && yaw_diff <= radian_thres) { isnew = false; break; }
      float rot_diff = rotationGeodesicDistance(R0, R1);
      if (rot_diff <= radian_thres) { isnew = false; break; }
    }
    if (isnew) hypos.push_back(hypo_tmp[i]);
  }
  for (size_t k = 0; k < hypos.size(); ++k) keep_out[k] = hypos[k].src;
  return (int)hypos.size();
}

// R.eulerAngles(2,1,0) of a column-major 4x4's rotation block (for unit tests of the product's restatement)
extern "C" void hop_ref_euler_zyx(const float *pose, float *rpy) {
  Eigen::Matrix4f P = Eigen::Map<const Eigen::Matrix4f>(pose);
  Eigen::Matrix3f R = P.block(0, 0, 3, 3);
  Eigen::Vector3f e = R.eulerAngles(2, 1, 0);
  rpy[0] = e(2); rpy[1] = e(1); rpy[2] = e(0);
}
