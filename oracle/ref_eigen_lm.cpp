// ref_eigen_lm.cpp -- ORACLE/_ref (TEST INFRASTRUCTURE ONLY).
//
// Drives the reference tree's OWN Levenberg-Marquardt implementation
// (/root/reference/src/OpenGR_4pcs/3rdparty/Eigen/unsupported/Eigen/NonLinearOptimization, NumericalDiff)
// exactly the way PCL 1.9's TransformationEstimationLM does for the point-to-plane ICP used by
// Utils::runICP (/root/reference/src/perception/src/Utils.cpp:188-229):
//     Eigen::NumericalDiff<Functor> num_diff(functor);
//     Eigen::LevenbergMarquardt<Eigen::NumericalDiff<Functor>, float> lm(num_diff);  lm.minimize(x);   x0 = 0 in R^6
// Compiled by oracle/Makefile against the Eigen headers where they lie under /root/reference; only the
// resulting shared object lands in oracle/_ref/.  Used by tests/ to pin hop_oracle.c's C restatement of LM.
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <unsupported/Eigen/NonLinearOptimization>
#include <unsupported/Eigen/NumericalDiff>

namespace {
struct P2PlaneFunctor {
  typedef float Scalar;
  enum { InputsAtCompileTime = Eigen::Dynamic, ValuesAtCompileTime = Eigen::Dynamic };
  typedef Eigen::Matrix<float, Eigen::Dynamic, 1> InputType;
  typedef Eigen::Matrix<float, Eigen::Dynamic, 1> ValueType;
  typedef Eigen::Matrix<float, Eigen::Dynamic, Eigen::Dynamic> JacobianType;
  const float *src, *tgt, *nrm;
  int m;
  int inputs() const { return 6; }
  int values() const { return m; }
  int operator()(const InputType &x, ValueType &fvec) const {
    // WarpPointRigid6D::setParam
    Eigen::Matrix4f T = Eigen::Matrix4f::Zero();
    T(0, 3) = x[0]; T(1, 3) = x[1]; T(2, 3) = x[2]; T(3, 3) = 1;
    Eigen::Quaternionf q(0, x[3], x[4], x[5]);
    q.w() = std::sqrt(1 - q.dot(q));
    q.normalize();
    T.topLeftCorner<3, 3>() = q.toRotationMatrix();
    for (int i = 0; i < m; ++i) {
      Eigen::Vector4f p(src[3 * i], src[3 * i + 1], src[3 * i + 2], 1.f);
      Eigen::Vector4f w = T * p; w[3] = 0;
      Eigen::Vector4f t(tgt[3 * i], tgt[3 * i + 1], tgt[3 * i + 2], 0);
      Eigen::Vector4f n(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2], 0);
      fvec[i] = (w - t).dot(n);  // TransformationEstimationPointToPlane::computeDistance
    }
    return 0;
  }
};
}  // namespace

extern "C" int hop_ref_lm_point_to_plane(const float *src, const float *tgt, const float *nrm, int m, float *x, int *nfev_out) {
  P2PlaneFunctor f; f.src = src; f.tgt = tgt; f.nrm = nrm; f.m = m;
  Eigen::NumericalDiff<P2PlaneFunctor> num_diff(f);
  Eigen::LevenbergMarquardt<Eigen::NumericalDiff<P2PlaneFunctor>, float> lm(num_diff);
  Eigen::VectorXf xv(6);
  for (int i = 0; i < 6; ++i) xv[i] = x[i];
  int info = (int)lm.minimize(xv);
  for (int i = 0; i < 6; ++i) x[i] = xv[i];
  if (nfev_out) *nfev_out = (int)lm.nfev;
  return info;
}
