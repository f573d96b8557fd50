// k_cluster.cu -- HOST: greedy pose clustering (non-maximum suppression with object symmetry).
//
//   replaces PoseEstimator<PointT>::clusterPoses       src/perception/src/PoseEstimator.cpp:106-233
//            Utils::rotationGeodesicDistance            src/perception/src/Utils.cpp:29-32
//
// Sequential by definition (a hypothesis is kept when no EARLIER kept cluster is close), so it stays on the host like in
// the reference; the Euler angles of every pose are extracted once instead of once per comparison.  Arithmetic follows
// Eigen 3.3's MatrixBase::eulerAngles(2,1,0) and fixed-size 3-term reductions (t0 + (t1 + t2)) so that threshold decisions
// match the reference's; rotationGeodesicDistance keeps the reference's trace(R1 * R2) (not R1^T R2).
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "../../include/hop_c_api.h"

namespace {

inline float m(const float *P, int r, int c) { return P[4 * c + r]; }  // column-major 4x4

struct Euler { float r, p, y; };

// R.eulerAngles(2,1,0): odd = 1, i = 2, j = 1, k = 0 (Eigen/src/Geometry/EulerAngles.h)
Euler euler_zyx(const float *P) {
  float res0 = std::atan2(m(P, 1, 0), m(P, 0, 0));
  const float a = m(P, 2, 2), b = m(P, 2, 1);
  const float c2 = std::sqrt(a * a + b * b);
  float res1;
  if (res0 < 0.f) {
    if (res0 > 0.f) res0 -= float(M_PI); else res0 += float(M_PI);
    res1 = std::atan2(-m(P, 2, 0), -c2);
  } else {
    res1 = std::atan2(-m(P, 2, 0), c2);
  }
  const float s1 = std::sin(res0), c1 = std::cos(res0);
  const float res2 = std::atan2(s1 * m(P, 0, 2) - c1 * m(P, 1, 2), c1 * m(P, 1, 1) - s1 * m(P, 0, 1));
  return Euler{res2, res1, res0};  // rpy(2), rpy(1), rpy(0)
}

inline float geodesic(const float *A, const float *B) {  // acos((trace(R1 * R2) - 1) / 2)
  float tr[3];
  for (int i = 0; i < 3; ++i) tr[i] = m(A, i, 0) * m(B, 0, i) + (m(A, i, 1) * m(B, 1, i) + m(A, i, 2) * m(B, 2, i));
  const float trace = tr[0] + (tr[1] + tr[2]);
  return (float)std::acos((trace - 1) / 2.0);
}

}  // namespace

extern "C" int hop_cluster_poses(const float *poses, const float *scores, const int32_t *ids, int n, float angle_diff_deg, float dist_diff,
                                 const float *symmetry_deg, int32_t *keep_out, int32_t *n_keep) {
  if (n < 0 || !n_keep || (n > 0 && (!poses || !scores || !keep_out)) || !symmetry_deg) return HOP_EINVAL;
  *n_keep = 0;
  if (n == 0) return HOP_OK;
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  auto id_of = [&](int k) { return ids ? ids[k] : k; };
  std::sort(order.begin(), order.end(), [&](int a, int b) {  // HypoCompare (:113-121): score descending, id ascending
    if (scores[a] > scores[b]) return true;
    if (scores[a] < scores[b]) return false;
    if (id_of(a) < id_of(b)) return true;
    return false;
  });
  const float radian_thres = angle_diff_deg / 180.0 * M_PI;
  const float sym[3] = {(float)((double)symmetry_deg[0] / 180 * M_PI), (float)((double)symmetry_deg[1] / 180 * M_PI),
                        (float)((double)symmetry_deg[2] / 180 * M_PI)};
  std::vector<Euler> eul(n);
  for (int k = 0; k < n; ++k) eul[k] = euler_zyx(poses + 16 * (size_t)k);
  auto fold = [](float diff, float s) {
    if (s == 0) return 0.f;
    if (s > 0) return std::min(diff, s - diff);
    return diff;
  };
  std::vector<int> clusters;
  clusters.push_back(order[0]);
  for (int i = 1; i < n; ++i) {
    const int cur = order[i];
    const float *P1 = poses + 16 * (size_t)cur;
    bool isnew = true;
    for (int c : clusters) {
      const float *P0 = poses + 16 * (size_t)c;
      const float dx = m(P0, 0, 3) - m(P1, 0, 3), dy = m(P0, 1, 3) - m(P1, 1, 3), dz = m(P0, 2, 3) - m(P1, 2, 3);
      if (std::sqrt(dx * dx + (dy * dy + dz * dz)) >= dist_diff) continue;
      const float roll_diff = fold(std::abs(eul[c].r - eul[cur].r), sym[0]);
      const float pitch_diff = fold(std::abs(eul[c].p - eul[cur].p), sym[1]);
      const float yaw_diff = fold(std::abs(eul[c].y - eul[cur].y), sym[2]);
      if (pitch_diff <= radian_thres && roll_diff <= radian_thres && yaw_diff <= radian_thres) { isnew = false; break; }
      if (geodesic(P0, P1) <= radian_thres) { isnew = false; break; }
    }
    if (isnew) clusters.push_back(cur);
  }
  for (size_t k = 0; k < clusters.size(); ++k) keep_out[k] = clusters[k];
  *n_keep = (int32_t)clusters.size();
  return HOP_OK;
}
