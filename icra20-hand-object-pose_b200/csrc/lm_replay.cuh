// lm_replay.cuh -- the inner solver of K4: PCL's TransformationEstimationPointToPlane (LM), replayed on moments.
//
// What the reference runs per ICP iteration (Utils.cpp:188-229 -> pcl::IterativeClosestPoint with
// TransformationEstimationPointToPlane<.., float>, PCL 1.9):
//     Eigen::LevenbergMarquardt<Eigen::NumericalDiff<Functor>, float> lm;  lm.minimize(x),  x0 = 0 in R^6
// i.e. MINPACK's lmdif in float: forward-difference Jacobian with h = sqrt(eps)*|x_j| (sqrt(eps) at 0), column-pivoted
// QR, lmpar trust region (factor 100), ftol = xtol = sqrt(eps), gtol = 0, maxfev = 400, on the residuals
//     f_k(x) = n_k . (W(x) s_k - t_k),   W = WarpPointRigid6D:  t = x[0..2],  q = (sqrt(1 - |x[3..5]|^2), x[3..5]).normalized()
// The minimiser it returns is NOT the minimum of the objective: it stops as soon as a step changes the sum of squares
// by less than sqrt(eps) relative, which along the weakly constrained directions of a partial view leaves the pose
// millimetres / degrees short of where an exact minimiser goes.  ICP trajectories fork there, so parity with the
// reference needs its solver's own steps, not a better one.
//
// Every residual is linear in y(x) = [vec(R(x)) - vec(I); t(x); 1] (13 numbers):  f_k = a_k . y,  a_k = [n (x) s; n; n.(s - t)].
// With A = sum_k a_k a_k^T (the 13x13 moment matrix K4 accumulates once per ICP iteration) every quantity lmdif looks at
// is a function of A and y:
//     |f(x)|^2 = y^T A y                              J = V D,  D_j = (y(x + h_j e_j) - y(x)) / h_j   (same float W(x), same h)
//     J^T J = D^T A D,   J^T f = D^T A y              R, Q^T f of the QR of J  =  Cholesky factor of J^T J, R^-T J^T f
// so the whole LM run -- every trial step, gain ratio, trust-region update and stopping test, in MINPACK's order -- is
// replayed on 91 numbers without revisiting the correspondences.  What is not reproduced is the rounding of the m
// individual float residuals (zero-mean, ~1e-6 relative on the norms; the reference's own result moves by as much when
// it is compiled with other flags).
//
// Host + device: the CUDA kernel calls lm_replay_solve from one warp (all lanes redundantly, the state is uniform); the CPU
// tests compile the same header to pin it against the reference tree's own Eigen LM (oracle/_ref) on fixed
// correspondence sets (tests/test_lm_replay.py).  This header is product code; it includes nothing from oracle/.
#pragma once
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define LMR_HD __host__ __device__ __forceinline__
#else
#define LMR_HD inline
#endif

namespace lmr {

constexpr int N = 6;
constexpr int NY = 13;

// unfused float arithmetic: the reference's Eigen code is compiled for baseline x86-64 (no FMA contraction)
LMR_HD float mul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b; return r;
#endif
}
LMR_HD float add(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b; return r;
#endif
}
LMR_HD float sub(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  volatile float r = a - b; return r;
#endif
}
LMR_HD float fdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
LMR_HD float fsqrt(float a) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  return sqrtf(a);
#endif
}

// pcl::registration::WarpPointRigid6D::setParam + Eigen::Quaternionf::toRotationMatrix, as y = [vec(R) - vec(I); t; 1]
LMR_HD void warp_y(const float *x, float *y) {
  float qx = x[3], qy = x[4], qz = x[5];
  float qw = fsqrt(sub(1.f, add(add(mul(qx, qx), mul(qy, qy)), mul(qz, qz))));
  const float nn = fsqrt(add(add(add(mul(qw, qw), mul(qx, qx)), mul(qy, qy)), mul(qz, qz)));
  qw = fdiv(qw, nn); qx = fdiv(qx, nn); qy = fdiv(qy, nn); qz = fdiv(qz, nn);
  const float tx = mul(2.f, qx), ty = mul(2.f, qy), tz = mul(2.f, qz);
  const float twx = mul(tx, qw), twy = mul(ty, qw), twz = mul(tz, qw);
  const float txx = mul(tx, qx), txy = mul(ty, qx), txz = mul(tz, qx);
  const float tyy = mul(ty, qy), tyz = mul(tz, qy), tzz = mul(tz, qz);
  // R - I: the diagonal is rounded to a float next to 1 first (that is the number the reference multiplies with)
  y[0] = sub(sub(1.f, add(tyy, tzz)), 1.f); y[1] = sub(txy, twz);                  y[2] = add(txz, twy);
  y[3] = add(txy, twz);                  y[4] = sub(sub(1.f, add(txx, tzz)), 1.f); y[5] = sub(tyz, twx);
  y[6] = sub(txz, twy);                  y[7] = add(tyz, twx);                  y[8] = sub(sub(1.f, add(txx, tyy)), 1.f);
  y[9] = x[0]; y[10] = x[1]; y[11] = x[2]; y[12] = 1.f;
}

// 2-norm of a short float vector the way Eigen's stableNorm / blueNorm deliver it (correctly scaled, double inside)
LMR_HD float norm6(const float *v) {
  double s = 0.0;
  for (int i = 0; i < N; ++i) s += (double)v[i] * (double)v[i];
  return (float)sqrt(s);
}

// ---- the quantities lmdif reads from the residual vector and the Jacobian, from the moments -------------------------
// Host: plain loops over a full 13x13 double matrix.  Device: the 91 float sums stay in shared memory, lane i < 13 of the
// calling warp owns row i, the replicated results meet through shuffles (all 32 lanes must call, with uniform arguments).
#if defined(__CUDACC__)
struct MomentsDev {
  const float *sums;   // upper triangle of the 13x13 moment matrix, row-major (shared memory)
  int lane;

  __device__ __forceinline__ double at(int i, int j) const {
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    return (double)sums[lo * 13 - (lo * (lo - 1)) / 2 + (hi - lo)];
  }
};
// g = A y (replicated), returns y^T A y
__device__ __forceinline__ double quad(const MomentsDev &A, const float *y, double *g) {
  const int li = A.lane < NY ? A.lane : NY - 1;
  double gi = 0.0;
#pragma unroll
  for (int j = 0; j < NY; ++j) gi = fma(A.at(li, j), (double)y[j], gi);
  double f2 = 0.0;
#pragma unroll
  for (int j = 0; j < NY; ++j) { g[j] = __shfl_sync(0xffffffffu, gi, j); f2 = fma(g[j], (double)y[j], f2); }
  return f2;
}
// Forward-difference Jacobian as moments: G = J^T J (upper triangle valid), b = J^T f.  Lane j < 6 evaluates W(x + h_j e_j);
// a translation column of D has the single entry (fl(x_j + h) - x_j) / h, a rotation column the nine entries of dR / h.
__device__ __forceinline__ void gram(const MomentsDev &A, const float *x, const float *y, const double *g, float h_eps, double (*G)[N], double *b, float *h_out) {
  const unsigned FULL = 0xffffffffu;
  const int lane = A.lane;
  float xx[N], hj = 1.f;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    float h = h_eps * fabsf(x[k]);
    if (h == 0.f) h = h_eps;
    xx[k] = lane == k ? add(x[k], h) : x[k];
    if (lane == k) hj = h;
  }
  float yj[NY];
  warp_y(xx, yj);
  float dy[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) dy[i] = sub(yj[i], y[i]);
  // broadcast the columns; `mine[k]` = this lane's row entry of column k of D*h (lane = row index of y)
  float dr[3][9], dt[3], mine[N], hs[N];
#pragma unroll
  for (int k = 0; k < N; ++k) { hs[k] = __shfl_sync(FULL, hj, k); mine[k] = 0.f; h_out[k] = hs[k]; }
#pragma unroll
  for (int c = 0; c < 3; ++c) { dt[c] = __shfl_sync(FULL, dy[9 + c], c); if (lane == 9 + c) mine[c] = dt[c]; }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int k = 0; k < 9; ++k) { dr[r][k] = __shfl_sync(FULL, dy[k], 3 + r); if (lane == k) mine[3 + r] = dr[r][k]; }
  // this lane's row of A D (unscaled): u[j] = sum_l A[lane][l] (D h)[l][j]
  const int li = lane < 12 ? lane : 11;
  double u[N], arow[12];
#pragma unroll
  for (int l = 0; l < 12; ++l) arow[l] = A.at(li, l);
#pragma unroll
  for (int c = 0; c < 3; ++c) u[c] = arow[9 + c] * (double)dt[c];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) s = fma(arow[k], (double)dr[r][k], s);
    u[3 + r] = s;
  }
  double inv_h[N];
#pragma unroll
  for (int k = 0; k < N; ++k) inv_h[k] = 1.0 / (double)hs[k];
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int k = j; k < N; ++k) {
      double p = lane < 12 ? (double)mine[j] * u[k] : 0.0;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) p += __shfl_xor_sync(FULL, p, o);   // rows 0..11 live in lanes 0..15
      p = __shfl_sync(FULL, p, 0);
      G[j][k] = p * inv_h[j] * inv_h[k];
    }
#pragma unroll
  for (int c = 0; c < 3; ++c) b[c] = (double)dt[c] * g[9 + c] * inv_h[c];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) s = fma((double)dr[r][k], g[k], s);
    b[3 + r] = s * inv_h[3 + r];
  }
}
#endif
struct Moments {        // host: symmetric 13x13, full storage, double (filled from float sums like the kernel's)
  double a[NY][NY];
  double at(int i, int j) const { return a[i][j]; }

};
inline double quad(const Moments &A, const float *y, double *g) {
  double f2 = 0.0;
  for (int i = 0; i < NY; ++i) {
    double s = 0.0;
    for (int j = 0; j < NY; ++j) s += A.a[i][j] * (double)y[j];
    g[i] = s;
    f2 += s * (double)y[i];
  }
  return f2;
}
inline void gram(const Moments &A, const float *x, const float *y, const double *g, float h_eps, double (*G)[N], double *b, float *h_out) {
  double D[N][NY];
  float xx[N];
  for (int k = 0; k < N; ++k) xx[k] = x[k];
  for (int j = 0; j < N; ++j) {
    float h = h_eps * fabsf(x[j]);
    if (h == 0.f) h = h_eps;
    xx[j] = add(x[j], h);
    h_out[j] = h;
    float yj[NY];
    warp_y(xx, yj);
    xx[j] = x[j];
    for (int i = 0; i < NY; ++i) D[j][i] = (double)sub(yj[i], y[i]) / (double)h;
  }
  for (int j = 0; j < N; ++j) {
    double AD[NY];
    for (int i = 0; i < NY; ++i) {
      double s = 0.0;
      for (int k = 0; k < NY - 1; ++k) s += A.a[i][k] * D[j][k];   // D[j][12] = 0
      AD[i] = s;
    }
    for (int k = j; k < N; ++k) {
      double s = 0.0;
      for (int i = 0; i < NY - 1; ++i) s += D[k][i] * AD[i];
      G[j][k] = s; G[k][j] = s;
    }
    double s = 0.0;
    for (int i = 0; i < NY - 1; ++i) s += D[j][i] * g[i];
    b[j] = s;
  }
}

#if defined(__CUDA_ARCH__)
#define LMR_UNROLL _Pragma("unroll")
#else
#define LMR_UNROLL
#endif

// MINPACK qrsolv on the 6x6 upper-triangular r (identity column order): least squares of [R; D] z = [qtb; 0].  r's strict
// lower triangle receives the transposed factor S (lmpar's Newton correction reads it), sdiag its diagonal.
LMR_HD void qrsolv(float (*r)[N], const float *diag, const float *qtb, float *x, float *sdiag) {
  float wa[N];
  LMR_UNROLL
  for (int j = 0; j < N; ++j) {
    LMR_UNROLL
    for (int i = j; i < N; ++i) r[i][j] = r[j][i];
    x[j] = r[j][j];
    wa[j] = qtb[j];
  }
  LMR_UNROLL
  for (int j = 0; j < N; ++j) {
    if (diag[j] != 0.f) {
      LMR_UNROLL
      for (int k = j; k < N; ++k) sdiag[k] = 0.f;
      sdiag[j] = diag[j];
      float qtbpj = 0.f;
      LMR_UNROLL
      for (int k = j; k < N; ++k) {
        if (sdiag[k] == 0.f) continue;
        float sn, cs;  // Givens rotation eliminating sdiag[k] against r[k][k] (Eigen's makeGivens differs only in rounding)
        if (fabsf(r[k][k]) < fabsf(sdiag[k])) { const float ct = r[k][k] / sdiag[k]; sn = 0.5f / sqrtf(0.25f + 0.25f * ct * ct); cs = sn * ct; }
        else { const float tn = sdiag[k] / r[k][k]; cs = 0.5f / sqrtf(0.25f + 0.25f * tn * tn); sn = cs * tn; }
        r[k][k] = cs * r[k][k] + sn * sdiag[k];
        const float tmp = cs * wa[k] + sn * qtbpj;
        qtbpj = -sn * wa[k] + cs * qtbpj;
        wa[k] = tmp;
        LMR_UNROLL
        for (int i = k + 1; i < N; ++i) {
          const float t2 = cs * r[i][k] + sn * sdiag[i];
          sdiag[i] = -sn * r[i][k] + cs * sdiag[i];
          r[i][k] = t2;
        }
      }
    }
    sdiag[j] = r[j][j];
    r[j][j] = x[j];
  }
  int nsing = N;
  LMR_UNROLL
  for (int j = 0; j < N; ++j) {
    if (sdiag[j] == 0.f && nsing == N) nsing = j;
    if (nsing < N) wa[j] = 0.f;
  }
  LMR_UNROLL
  for (int j = N - 1; j >= 0; --j) {
    if (j < nsing) {
      float sum = 0.f;
      LMR_UNROLL
      for (int i = j + 1; i < N; ++i) sum += r[i][j] * wa[i];   // wa[i] = 0 beyond nsing
      wa[j] = (wa[j] - sum) / sdiag[j];
    }
  }
  LMR_UNROLL
  for (int j = 0; j < N; ++j) x[j] = wa[j];
}

// MINPACK lmpar: the LM parameter par with | |D x| - delta | <= 0.1 delta (or par = 0 when the Gauss-Newton step fits)
LMR_HD void lmpar(float (*r)[N], const float *diag, const float *qtb, float delta, float &par, float *x) {
  const float dwarf = FLT_MIN;
  float wa1[N], wa2[N], sdiag[N];
  int nsing = N;
  LMR_UNROLL
  for (int j = 0; j < N; ++j) {
    wa1[j] = qtb[j];
    if (r[j][j] == 0.f && nsing == N) nsing = j;
    if (nsing < N) wa1[j] = 0.f;
  }
  LMR_UNROLL
  for (int j = N - 1; j >= 0; --j) {
    if (j < nsing) {
      wa1[j] /= r[j][j];
      const float t = wa1[j];
      LMR_UNROLL
      for (int i = 0; i < j; ++i) wa1[i] -= r[i][j] * t;
    }
  }
  LMR_UNROLL
  for (int j = 0; j < N; ++j) { x[j] = wa1[j]; wa2[j] = diag[j] * x[j]; }
  int iter = 0;
  float dxnorm = norm6(wa2);
  float fp = dxnorm - delta;
  if (fp <= 0.1f * delta) { par = 0.f; return; }
  float parl = 0.f;
  if (nsing >= N) {
    LMR_UNROLL
    for (int j = 0; j < N; ++j) wa1[j] = diag[j] * (wa2[j] / dxnorm);
    LMR_UNROLL
    for (int j = 0; j < N; ++j) {
      float sum = 0.f;
      LMR_UNROLL
      for (int i = 0; i < j; ++i) sum += r[i][j] * wa1[i];
      wa1[j] = (wa1[j] - sum) / r[j][j];
    }
    const float t = norm6(wa1);
    parl = fp / delta / t / t;
  }
  LMR_UNROLL
  for (int j = 0; j < N; ++j) {
    float sum = 0.f;
    LMR_UNROLL
    for (int i = 0; i <= j; ++i) sum += r[i][j] * qtb[i];
    wa1[j] = sum / diag[j];
  }
  const float gnorm = norm6(wa1);
  float paru = gnorm / delta;
  if (paru == 0.f) paru = dwarf / fminf(delta, 0.1f);
  par = fmaxf(par, parl);
  par = fminf(par, paru);
  if (par == 0.f) par = gnorm / dxnorm;
  for (;;) {
    ++iter;
    if (par == 0.f) par = fmaxf(dwarf, 0.001f * paru);
    const float sq = sqrtf(par);
    LMR_UNROLL
    for (int j = 0; j < N; ++j) wa1[j] = sq * diag[j];
    float rr[N][N];
    LMR_UNROLL
    for (int i = 0; i < N; ++i)
      LMR_UNROLL
      for (int j = 0; j < N; ++j) rr[i][j] = r[i][j];
    qrsolv(rr, wa1, qtb, x, sdiag);
    LMR_UNROLL
    for (int j = 0; j < N; ++j) wa2[j] = diag[j] * x[j];
    dxnorm = norm6(wa2);
    float temp = fp;
    fp = dxnorm - delta;
    if (fabsf(fp) <= 0.1f * delta || (parl == 0.f && fp <= temp && temp < 0.f) || iter == 10) break;
    LMR_UNROLL
    for (int j = 0; j < N; ++j) wa1[j] = diag[j] * (wa2[j] / dxnorm);
    LMR_UNROLL
    for (int j = 0; j < N; ++j) {
      wa1[j] /= sdiag[j];
      const float t = wa1[j];
      LMR_UNROLL
      for (int i = j + 1; i < N; ++i) wa1[i] -= rr[i][j] * t;
    }
    temp = norm6(wa1);
    const float parc = fp / delta / temp / temp;
    if (fp > 0.f) parl = fmaxf(parl, par);
    if (fp < 0.f) paru = fminf(paru, par);
    par = fmaxf(parl, par + parc);
  }
}

// When does the reference's LM run away?  The translation columns of its Jacobian are the target normals (times 1 + O(1e-5) of
// forward-difference rounding), so J^T J restricted to the translations is the scatter matrix of the normals, N = sum n n^T,
// whatever x is.  If the correspondences' normals do not span three dimensions (all points on one face of a box, or on two faces:
// the sets a coarse hypothesis of a polyhedral object produces), N is singular in exact arithmetic, the float QR of the reference
// sees a pivot that is pure rounding noise (~1e-5 against column norms of ~10), divides the equally arbitrary right-hand side by
// it and accepts a slide of METRES along the free direction -- the residuals do not change along it.  The next ICP iteration finds
// no correspondences, reg.hasConverged() is false and Utils::runICP returns the identity: the hypothesis keeps its pose
// (Utils.cpp:218-225).  That outcome is deterministic even though the slide itself is noise; a solver working on exact moments
// would instead refine the constrained directions and return a different pose.  A free ROTATION (cylinder about its axis) is
// harmless in the reference: the quaternion parametrisation is bounded, steps with |q| > 1 give NaN residuals and are rejected.
// Returns true when N is singular at the precision of float-accumulated moments: relative pivot^2 < 1e-5 (their rounding leaves up
// to ~2e-6; one stray correspondence in a thousand on a third face already gives 1e-3).
template <typename MomentsT>
LMR_HD bool translation_unconstrained(const MomentsT &A) {
  double n00 = A.at(9, 9), n01 = A.at(9, 10), n02 = A.at(9, 11), n11 = A.at(10, 10), n12 = A.at(10, 11), n22 = A.at(11, 11);
  const double tol = 1e-5;
  if (!(n00 > 0.0)) return true;
  const double l01 = n01 / n00, l02 = n02 / n00;           // LDL^T, no square roots
  const double d1 = n11 - l01 * n01;
  if (!(d1 > tol * n11)) return true;
  const double m12 = n12 - l01 * n02;
  const double d2 = n22 - l02 * n02 - m12 * m12 / d1;
  return !(d2 > tol * n22);
}

// The LM run.  x (6) must be zero on entry (PCL starts from the identity); returns Eigen's LevenbergMarquardtSpace status.
// Device: called by all 32 lanes of one warp with identical arguments.
template <typename MomentsT>
LMR_HD int lm_replay_solve(const MomentsT &A, float *x, int *nfev_out) {
  const float eps = FLT_EPSILON;
  const float h_eps = 3.4526698e-4f, ftol = h_eps, xtol = h_eps, gtol = 0.f, factor = 100.f;   // sqrt(FLT_EPSILON)
  const int maxfev = 400;
  float y[NY], diag[N], qtf[N], wa1[N], wa2[N], wa3[N];
  double g[NY];
  int nfev = 1, iter = 1, status = 0;
  float par = 0.f, delta = 0.f, xnorm = 0.f;
  warp_y(x, y);
  double f2 = quad(A, y, g);
  float fnorm = (float)sqrt(f2 > 0.0 ? f2 : 0.0);
  for (;;) {
    double G[N][N], b[N];
    float hs[N];
    gram(A, x, y, g, h_eps, G, b, hs);
    nfev += N + 1;   // NumericalDiff::df (Forward) re-evaluates f(x) first: n + 1 evaluations
    LMR_UNROLL
    for (int j = 0; j < N; ++j) wa2[j] = (float)sqrt(G[j][j] > 0.0 ? G[j][j] : 0.0);
    // ---- R and Q^T f of the QR of J = Cholesky of G, R^-T J^T f.  The reference pivots its Householder QR on the column
    //      norms; the order only changes the rounding of what follows (the step is a function of J^T J), and a Cholesky
    //      factorisation in double needs no pivoting for accuracy, so the columns keep their order. ----
    float r[N][N];
    {
      double Lr[N][N];
      LMR_UNROLL
      for (int j = 0; j < N; ++j) {
        double d = G[j][j];
        LMR_UNROLL
        for (int i = 0; i < j; ++i) d -= Lr[i][j] * Lr[i][j];
        const double dd = d > 0.0 ? sqrt(d) : 0.0;
        const double inv = dd > 0.0 ? 1.0 / dd : 0.0;
        Lr[j][j] = dd;
        LMR_UNROLL
        for (int k = j + 1; k < N; ++k) {
          double s = G[j][k];
          LMR_UNROLL
          for (int i = 0; i < j; ++i) s -= Lr[i][j] * Lr[i][k];
          Lr[j][k] = s * inv;
        }
        double s = b[j];
        LMR_UNROLL
        for (int i = 0; i < j; ++i) s -= Lr[i][j] * (double)qtf[i];
        qtf[j] = (float)(s * inv);
      }
      LMR_UNROLL
      for (int i = 0; i < N; ++i)
        LMR_UNROLL
        for (int j = 0; j < N; ++j) r[i][j] = j >= i ? (float)Lr[i][j] : 0.f;
    }
    if (iter == 1) {
      LMR_UNROLL
      for (int j = 0; j < N; ++j) { diag[j] = (wa2[j] == 0.f) ? 1.f : wa2[j]; wa3[j] = diag[j] * x[j]; }
      xnorm = norm6(wa3);
      delta = factor * xnorm;
      if (delta == 0.f) delta = factor;
    }
    float gnorm = 0.f;
    if (fnorm != 0.f) {
      LMR_UNROLL
      for (int j = 0; j < N; ++j)
        if (wa2[j] != 0.f) {
          float s = 0.f;
          LMR_UNROLL
          for (int i = 0; i <= j; ++i) s += r[i][j] * (qtf[i] / fnorm);
          gnorm = fmaxf(gnorm, fabsf(s / wa2[j]));
        }
    }
    if (gnorm <= gtol) { status = 4; break; }
    LMR_UNROLL
    for (int j = 0; j < N; ++j) diag[j] = fmaxf(diag[j], wa2[j]);
    float ratio;
    bool done = false;
    do {
      lmpar(r, diag, qtf, delta, par, wa1);
      LMR_UNROLL
      for (int j = 0; j < N; ++j) { wa1[j] = -wa1[j]; wa2[j] = x[j] + wa1[j]; wa3[j] = diag[j] * wa1[j]; }
      const float pnorm = norm6(wa3);
      if (iter == 1) delta = fminf(delta, pnorm);
      float y1[NY];
      double g1[NY];
      warp_y(wa2, y1);
      const double f21 = quad(A, y1, g1);
      ++nfev;
      const float fnorm1 = (float)sqrt(f21 > 0.0 ? f21 : 0.0);
      float actred = -1.f;
      if (0.1f * fnorm1 < fnorm) { const float q = fnorm1 / fnorm; actred = 1.f - q * q; }
      LMR_UNROLL
      for (int i = 0; i < N; ++i) {
        float s = 0.f;
        LMR_UNROLL
        for (int j = i; j < N; ++j) s += r[i][j] * wa1[j];
        wa3[i] = s;
      }
      float t1 = norm6(wa3) / fnorm; t1 *= t1;
      float t2 = sqrtf(par) * pnorm / fnorm; t2 *= t2;
      const float prered = t1 + t2 / 0.5f;
      const float dirder = -(t1 + t2);
      ratio = 0.f;
      if (prered != 0.f) ratio = actred / prered;
      if (ratio <= 0.25f) {
        float temp = 0.5f;
        if (actred < 0.f) temp = 0.5f * dirder / (dirder + 0.5f * actred);
        if (0.1f * fnorm1 >= fnorm || temp < 0.1f) temp = 0.1f;
        delta = temp * fminf(delta, pnorm / 0.1f);
        par /= temp;
      } else if (!(par != 0.f && ratio < 0.75f)) {
        delta = pnorm / 0.5f;
        par = 0.5f * par;
      }
      if (ratio >= 1e-4f) {
        LMR_UNROLL
        for (int j = 0; j < N; ++j) { x[j] = wa2[j]; wa2[j] = diag[j] * x[j]; }
        LMR_UNROLL
        for (int i = 0; i < NY; ++i) { y[i] = y1[i]; g[i] = g1[i]; }
        xnorm = norm6(wa2);
        fnorm = fnorm1;
        ++iter;
      }
      const bool small_f = fabsf(actred) <= ftol && prered <= ftol && 0.5f * ratio <= 1.f;
      if (small_f && delta <= xtol * xnorm) { status = 3; done = true; break; }
      if (small_f) { status = 1; done = true; break; }
      if (delta <= xtol * xnorm) { status = 2; done = true; break; }
      if (nfev >= maxfev) { status = 5; done = true; break; }
      if (fabsf(actred) <= eps && prered <= eps && 0.5f * ratio <= 1.f) { status = 6; done = true; break; }
      if (delta <= eps * xnorm) { status = 7; done = true; break; }
      if (gnorm <= eps) { status = 8; done = true; break; }
    } while (ratio < 1e-4f);
    if (done) break;
  }
  if (nfev_out) *nfev_out = nfev;
  return status;
}

}  // namespace lmr
