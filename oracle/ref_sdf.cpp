// ref_sdf.cpp -- test infrastructure: the reference's OWN signed-distance code (its vendored libigl, src/perception/include/igl,
// compiled where it lies against the reference tree's Eigen) behind a C entry point.  This is what SDFchecker::
// getSignedDistanceMinMaxWithRegistered (src/perception/src/SDFchecker.cpp:115-134) calls:
//     igl::signed_distance(pts, V, F, SIGNED_DISTANCE_TYPE_PSEUDONORMAL, -FLT_MAX, FLT_MAX, S, I, C, N)
// Built into oracle/_ref/libhop_ref.so by oracle/build_ref.sh; only tests/ and the bench's CPU legs may call it.
#include <igl/signed_distance.h>

#include <Eigen/Core>
#include <limits>

extern "C" int hop_ref_signed_distance(const float *pts, int n, const float *V, int nv, const int *F, int nf, float *S, int *I, float *C, float *N) {
  Eigen::MatrixXf P(n, 3), Vm(nv, 3);
  Eigen::MatrixXi Fm(nf, 3);
  for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) P(i, k) = pts[3 * i + k];
  for (int i = 0; i < nv; ++i) for (int k = 0; k < 3; ++k) Vm(i, k) = V[3 * i + k];
  for (int i = 0; i < nf; ++i) for (int k = 0; k < 3; ++k) Fm(i, k) = F[3 * i + k];
  Eigen::VectorXf Sv, Iv;
  Eigen::MatrixXf Cm, Nm;
  igl::signed_distance(P, Vm, Fm, igl::SIGNED_DISTANCE_TYPE_PSEUDONORMAL, -std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), Sv, Iv, Cm, Nm);
  for (int i = 0; i < n; ++i) {
    S[i] = Sv(i);
    if (I) I[i] = (int)Iv(i);
    for (int k = 0; k < 3; ++k) { if (C) C[3 * i + k] = Cm(i, k); if (N) N[3 * i + k] = Nm(i, k); }
  }
  return 0;
}
