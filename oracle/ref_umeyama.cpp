// ref_umeyama.cpp -- ORACLE/_ref (TEST INFRASTRUCTURE ONLY).
// The Kabsch step of PCL's default ICP (TransformationEstimationSVD, use_umeyama = true: pcl::umeyama is a copy of
// Eigen::umeyama, float, no scaling), which Utils::runICP(pclSegment, pclModel, offsetTransform, max_corres_dist)
// (/root/reference/src/perception/src/Utils.cpp:135-164) runs per iteration -- driven through the reference tree's own
// Eigen (src/OpenGR_4pcs/3rdparty/Eigen/Eigen/Geometry, Umeyama.h).  Pins oracle/hop_oracle.c: hop_oracle_kabsch.
#include <Eigen/Core>
#include <Eigen/Geometry>

extern "C" void hop_ref_umeyama(const float *src, const float *dst, int n, float *T_colmajor) {
  Eigen::Matrix<float, 3, Eigen::Dynamic> S(3, n), D(3, n);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) { S(k, i) = src[3 * i + k]; D(k, i) = dst[3 * i + k]; }
  Eigen::Matrix4f T = Eigen::umeyama(S, D, false);
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) T_colmajor[4 * c + r] = T(r, c);
}
