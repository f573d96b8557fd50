/* hop_oracle_frame.c -- ORACLE (TEST INFRASTRUCTURE ONLY): CPU restatements of the two PCL normal estimators the reference's
 * main_realdata_auto runs around the hand branch.  PCL is not installed and not vendored: both follow PCL 1.9.1 from memory of
 *   features/include/pcl/features/impl/integral_image_normal.hpp, integral_image2D.hpp   (IntegralImageNormalEstimation)
 *   surface/include/pcl/surface/impl/mls.hpp, common/include/pcl/common/impl/eigen.hpp   (MovingLeastSquares, eigen33)
 * -> PARITY UNPINNED against PCL; the tests pin them against closed forms (planes, spheres) and the GPU against them.
 *
 * Utils::calNormalIntegralImage(cloud, -1, 0.02, 10, true)   Utils.cpp:294-329, called main_realdata_auto.cpp:61:
 *   method -1 -> SIMPLE_3D_GRADIENT, max_depth_change_factor 0.02, normal_smoothing_size 10, depth dependent smoothing,
 *   viewpoint (0,0,0), border policy IGNORE, rectangle = whole cloud.
 * Utils::calNormalMLS(cloud, 0.003)                           Utils.cpp:268-292, called main_realdata_auto.cpp:160:
 *   MovingLeastSquares, polynomial order 2, radius search, compute normals, projection SIMPLE, no upsampling.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * IntegralImageNormalEstimation::computeFeature + computeFeatureFull + computePointNormal (SIMPLE_3D_GRADIENT)
 * xyz: organized cloud, row-major h x w x 3 (invalid pixels are (0,0,0) in the reference: Utils.cpp:103-110, finite).
 * nrm: h x w x 3, NaN where PCL writes its bad point.
 * ---------------------------------------------------------------------------------------------- */
void hop_oracle_integral_image_normals(const float *xyz, int w, int h, float max_depth_change_factor, float normal_smoothing_size,
                                       float *nrm) {
  const size_t n = (size_t)w * h;
  const float bad = NAN;
  unsigned char *change = (unsigned char *)malloc(n);
  float *dist = (float *)malloc(sizeof(float) * n);
  memset(change, 255, n);
  for (int ri = 0; ri < h - 1; ++ri)
    for (int ci = 0; ci < w - 1; ++ci) {
      const size_t index = (size_t)ri * w + ci;
      const float depth = xyz[3 * index + 2], depthR = xyz[3 * (index + 1) + 2], depthD = xyz[3 * (index + w) + 2];
      const float thr = (max_depth_change_factor * (fabsf(depth) + 1.0f) * 2.0f);
      if (fabs(depth - depthR) > thr || !isfinite(depth) || !isfinite(depthR)) { change[index] = 0; change[index + 1] = 0; }
      if (fabs(depth - depthD) > thr || !isfinite(depth) || !isfinite(depthD)) { change[index] = 0; change[index + w] = 0; }
    }
  for (size_t i = 0; i < n; ++i) dist[i] = change[i] == 0 ? 0.0f : (float)(w + h);
  /* first pass (note ci runs to w - 1: previous_row[ci + 1] then reads the first element of the current row, as in PCL) */
  for (int ri = 1; ri < h; ++ri) {
    float *prev = dist + (size_t)(ri - 1) * w, *cur = dist + (size_t)ri * w;
    for (int ci = 1; ci < w; ++ci) {
      const float upLeft = prev[ci - 1] + 1.4f, up = prev[ci] + 1.0f, upRight = prev[ci + 1] + 1.4f, left = cur[ci - 1] + 1.0f;
      const float center = cur[ci];
      const float m1 = upLeft < up ? upLeft : up, m2 = left < upRight ? left : upRight;
      const float mv = m1 < m2 ? m1 : m2;
      if (mv < center) cur[ci] = mv;
    }
  }
  /* second pass (ci runs down to 0: next_row[ci - 1] then reads the last element of the current row) */
  for (int ri = h - 2; ri >= 0; --ri) {
    float *next = dist + (size_t)(ri + 1) * w, *cur = dist + (size_t)ri * w;
    for (int ci = w - 2; ci >= 0; --ci) {
      const float lowerLeft = next[ci - 1] + 1.4f, lower = next[ci] + 1.0f, lowerRight = next[ci + 1] + 1.4f, right = cur[ci + 1] + 1.0f;
      const float center = cur[ci];
      const float m1 = lowerLeft < lower ? lowerLeft : lower, m2 = right < lowerRight ? right : lowerRight;
      const float mv = m1 < m2 ? m1 : m2;
      if (mv < center) cur[ci] = mv;
    }
  }
  /* IntegralImage2D<float, 3>: first-order sums in double, (w + 1) x (h + 1); non-finite elements are skipped and counted */
  const int W1 = w + 1;
  double *I = (double *)calloc((size_t)W1 * (h + 1) * 3, sizeof(double));
  unsigned *F = (unsigned *)calloc((size_t)W1 * (h + 1), sizeof(unsigned));
  for (int r = 0; r < h; ++r) {
    double so_far[3] = {0, 0, 0};
    unsigned fin = 0;
    for (int c = 0; c < w; ++c) {
      const float *p = xyz + 3 * ((size_t)r * w + c);
      if (isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2])) { so_far[0] += p[0]; so_far[1] += p[1]; so_far[2] += p[2]; ++fin; }
      for (int k = 0; k < 3; ++k) I[3 * ((size_t)(r + 1) * W1 + c + 1) + k] = I[3 * ((size_t)r * W1 + c + 1) + k] + so_far[k];
      F[(size_t)(r + 1) * W1 + c + 1] = F[(size_t)r * W1 + c + 1] + fin;
    }
  }
#define II_SUM(out, sx, sy, ww, hh)                                                                        \
  do {                                                                                                     \
    const size_t ul = (size_t)(sy) * W1 + (sx), ur = ul + (ww), ll = (size_t)((sy) + (hh)) * W1 + (sx), lr = ll + (ww); \
    for (int k_ = 0; k_ < 3; ++k_) (out)[k_] = I[3 * lr + k_] + I[3 * ul + k_] - I[3 * ur + k_] - I[3 * ll + k_];       \
  } while (0)
  for (size_t i = 0; i < 3 * n; ++i) nrm[i] = bad;
  const int border = (int)normal_smoothing_size;
  for (int ri = border; ri < h - border; ++ri)
    for (int ci = border; ci < w - border; ++ci) {
      const size_t index = (size_t)ri * w + ci;
      const float depth = xyz[3 * index + 2];
      if (!isfinite(depth)) continue;
      const float a = dist[index], b = normal_smoothing_size + depth / 10.0f;
      const float smoothing = a < b ? a : b;
      if (!(smoothing > 2.0f)) continue;
      const int rw = (int)smoothing, rh = (int)smoothing, rw2 = rw / 2, rh2 = rh / 2;
      {
        const int sx = ci - rw2, sy = ri - rh2;
        const size_t ul = (size_t)sy * W1 + sx, ur = ul + rw, ll = (size_t)(sy + rh) * W1 + sx, lr = ll + rw;
        const unsigned cnt = F[lr] + F[ul] - F[ur] - F[ll];
        if (cnt == 0) continue;
      }
      double a1[3], a2[3], gx[3], gy[3];
      II_SUM(a1, ci + rw2, ri - rh2, 1, rh); II_SUM(a2, ci - rw2, ri - rh2, 1, rh);
      for (int k = 0; k < 3; ++k) gx[k] = a1[k] - a2[k];
      II_SUM(a1, ci - rw2, ri + rh2, rw, 1); II_SUM(a2, ci - rw2, ri - rh2, rw, 1);
      for (int k = 0; k < 3; ++k) gy[k] = a1[k] - a2[k];
      double nv[3] = {gy[1] * gx[2] - gy[2] * gx[1], gy[2] * gx[0] - gy[0] * gx[2], gy[0] * gx[1] - gy[1] * gx[0]};
      const double len = nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2];
      if (len == 0.0) continue;
      const double s = sqrt(len);
      float nx = (float)(nv[0] / s), ny = (float)(nv[1] / s), nz = (float)(nv[2] / s);
      /* pcl::flipNormalTowardsViewpoint (viewpoint 0,0,0) */
      const float vx = 0.f - xyz[3 * index], vy = 0.f - xyz[3 * index + 1], vz = 0.f - xyz[3 * index + 2];
      const float cos_theta = (vx * nx + vy * ny + vz * nz);
      if (cos_theta < 0) { nx *= -1; ny *= -1; nz *= -1; }
      nrm[3 * index] = nx; nrm[3 * index + 1] = ny; nrm[3 * index + 2] = nz;
    }
#undef II_SUM
  free(change); free(dist); free(I); free(F);
}

/* ------------------------------------------------------------------------------------------------
 * pcl::eigen33 (mat, eigenvalue, eigenvector): smallest eigenvalue of a symmetric 3x3 (double) and its eigenvector, closed form
 * (computeRoots) + the largest cross product of two rows of (A - lambda I).   common/impl/eigen.hpp
 * ---------------------------------------------------------------------------------------------- */
static void pcl_compute_roots2(double b, double c, double roots[3]) {
  roots[0] = 0.0;
  double d = b * b - 4.0 * c;
  if (d < 0.0) d = 0.0;
  const double sd = sqrt(d);
  roots[2] = 0.5 * (b + sd);
  roots[1] = 0.5 * (b - sd);
}

static void pcl_compute_roots(const double m[3][3], double roots[3]) {
  const double c0 = m[0][0] * m[1][1] * m[2][2] + 2.0 * m[0][1] * m[0][2] * m[1][2] - m[0][0] * m[1][2] * m[1][2] -
                    m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1];
  const double c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2] + m[1][1] * m[2][2] - m[1][2] * m[1][2];
  const double c2 = m[0][0] + m[1][1] + m[2][2];
  if (fabs(c0) < DBL_EPSILON) { pcl_compute_roots2(c2, c1, roots); return; }
  const double s_inv3 = 1.0 / 3.0, s_sqrt3 = sqrt(3.0);
  const double c2_over_3 = c2 * s_inv3;
  double a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.0) a_over_3 = 0.0;
  const double half_b = 0.5 * (c0 + c2_over_3 * (2.0 * c2_over_3 * c2_over_3 - c1));
  double q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.0) q = 0.0;
  const double rho = sqrt(-a_over_3);
  const double theta = atan2(sqrt(-q), half_b) * s_inv3;
  const double cos_theta = cos(theta), sin_theta = sin(theta);
  roots[0] = c2_over_3 + 2.0 * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  /* sort in increasing order */
  if (roots[0] >= roots[1]) { double t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
  if (roots[1] >= roots[2]) {
    double t = roots[1]; roots[1] = roots[2]; roots[2] = t;
    if (roots[0] >= roots[1]) { t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
  }
  if (roots[0] <= 0.0) pcl_compute_roots2(c2, c1, roots); /* eigenvalues of a covariance matrix are non-negative */
}

static void pcl_eigen33_smallest(const double A[3][3], double *eigenvalue, double ev[3]) {
  double scale = 0.0;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) if (fabs(A[i][j]) > scale) scale = fabs(A[i][j]);
  if (scale <= DBL_MIN) scale = 1.0;
  double S[3][3], roots[3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) S[i][j] = A[i][j] / scale;
  pcl_compute_roots(S, roots);
  *eigenvalue = roots[0] * scale;
  for (int i = 0; i < 3; ++i) S[i][i] -= roots[0];
  double v1[3] = {S[0][1] * S[1][2] - S[0][2] * S[1][1], S[0][2] * S[1][0] - S[0][0] * S[1][2], S[0][0] * S[1][1] - S[0][1] * S[1][0]};
  double v2[3] = {S[0][1] * S[2][2] - S[0][2] * S[2][1], S[0][2] * S[2][0] - S[0][0] * S[2][2], S[0][0] * S[2][1] - S[0][1] * S[2][0]};
  double v3[3] = {S[1][1] * S[2][2] - S[1][2] * S[2][1], S[1][2] * S[2][0] - S[1][0] * S[2][2], S[1][0] * S[2][1] - S[1][1] * S[2][0]};
  const double l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2], l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2],
               l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
  const double *v; double l;
  if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; } else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; } else { v = v3; l = l3; }
  const double s = sqrt(l);
  for (int k = 0; k < 3; ++k) ev[k] = v[k] / s;
}

/* Eigen::Vector3d::unitOrthogonal() */
static void unit_orthogonal(const double n[3], double out[3]) {
  /* !isMuchSmallerThan(x, z) || !isMuchSmallerThan(y, z)  with the default precision 1e-12 */
  if (!(fabs(n[0]) <= 1e-12 * fabs(n[2])) || !(fabs(n[1]) <= 1e-12 * fabs(n[2]))) {
    const double inv = 1.0 / sqrt(n[0] * n[0] + n[1] * n[1]);
    out[0] = -n[1] * inv; out[1] = n[0] * inv; out[2] = 0.0;
  } else {
    const double inv = 1.0 / sqrt(n[1] * n[1] + n[2] * n[2]);
    out[0] = 0.0; out[1] = -n[2] * inv; out[2] = n[1] * inv;
  }
}

/* Cholesky solve of the 6x6 normal equations (Eigen LLT) */
static int llt_solve6(double A[6][6], double b[6]) {
  double L[6][6] = {{0}};
  for (int j = 0; j < 6; ++j) {
    double d = A[j][j];
    for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
    if (!(d > 0.0)) return 0;
    L[j][j] = sqrt(d);
    for (int i = j + 1; i < 6; ++i) {
      double s = A[i][j];
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      L[i][j] = s / L[j][j];
    }
  }
  for (int i = 0; i < 6; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[i][k] * b[k]; b[i] = s / L[i][i]; }
  for (int i = 5; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < 6; ++k) s -= L[k][i] * b[k]; b[i] = s / L[i][i]; }
  return 1;
}

/* MovingLeastSquares::process for every point of a cloud (neighbours by brute force within `radius`, d^2 <= r^2 like FLANN's
 * radius search; order of the neighbours does not enter the result beyond summation order: ascending index here).
 * valid[i] = 1 when the point has >= 3 neighbours (mls.getCorrespondingIndices()); xyz_out / nrm_out: the projected point and the
 * normal of the fitted surface there (NOT oriented: PCL's MLS does not flip them). */
void hop_oracle_mls(const float *xyz, int n, float radius, float *xyz_out, float *nrm_out, int *valid) {
  const double r2 = (double)radius * (double)radius;
  const float r2f = radius * radius;
#pragma omp parallel
  {
    int *nn = (int *)malloc(sizeof(int) * (size_t)(n + 1));
#pragma omp for schedule(dynamic, 64)
    for (int i = 0; i < n; ++i) {
      const float *q = xyz + 3 * i;
      int m = 0;
      for (int j = 0; j < n; ++j) {
        const float dx = xyz[3 * j] - q[0], dy = xyz[3 * j + 1] - q[1], dz = xyz[3 * j + 2] - q[2];
        if (dx * dx + dy * dy + dz * dz <= r2f) nn[m++] = j;
      }
      valid[i] = m >= 3;
      for (int k = 0; k < 3; ++k) { xyz_out[3 * i + k] = q[k]; nrm_out[3 * i + k] = NAN; }
      if (m < 3) continue;
      /* computeMeanAndCovarianceMatrix (double accumulators, single pass: accu of x, y, z, xx, xy, xz, yy, yz, zz) */
      double acc[9] = {0};
      for (int t = 0; t < m; ++t) {
        const float *p = xyz + 3 * nn[t];
        acc[0] += (double)p[0] * p[0]; acc[1] += (double)p[0] * p[1]; acc[2] += (double)p[0] * p[2];
        acc[3] += (double)p[1] * p[1]; acc[4] += (double)p[1] * p[2]; acc[5] += (double)p[2] * p[2];
        acc[6] += p[0]; acc[7] += p[1]; acc[8] += p[2];
      }
      for (int k = 0; k < 9; ++k) acc[k] /= (double)m;
      double C[3][3];
      C[0][0] = acc[0] - acc[6] * acc[6]; C[0][1] = acc[1] - acc[6] * acc[7]; C[0][2] = acc[2] - acc[6] * acc[8];
      C[1][1] = acc[3] - acc[7] * acc[7]; C[1][2] = acc[4] - acc[7] * acc[8]; C[2][2] = acc[5] - acc[8] * acc[8];
      C[1][0] = C[0][1]; C[2][0] = C[0][2]; C[2][1] = C[1][2];
      double ev, pn[3];
      pcl_eigen33_smallest(C, &ev, pn);
      if (!isfinite(pn[0]) || !isfinite(pn[1]) || !isfinite(pn[2])) continue; /* invalid plane: point unchanged, normal zero -> reported as NaN */
      const double d4 = -(pn[0] * acc[6] + pn[1] * acc[7] + pn[2] * acc[8]);
      const double qd[3] = {q[0], q[1], q[2]};
      const double distance = qd[0] * pn[0] + qd[1] * pn[1] + qd[2] * pn[2] + d4;
      double mean[3] = {qd[0] - distance * pn[0], qd[1] - distance * pn[1], qd[2] - distance * pn[2]};
      double va[3], ua[3];
      unit_orthogonal(pn, va);
      ua[0] = pn[1] * va[2] - pn[2] * va[1]; ua[1] = pn[2] * va[0] - pn[0] * va[2]; ua[2] = pn[0] * va[1] - pn[1] * va[0];
      double c[6] = {0, 0, 0, 0, 0, 0};
      int have_poly = 0;
      if (m >= 6) {
        double PtWP[6][6] = {{0}}, rhs[6] = {0};
        for (int t = 0; t < m; ++t) {
          const float *p = xyz + 3 * nn[t];
          const double de[3] = {p[0] - mean[0], p[1] - mean[1], p[2] - mean[2]};
          const double wgt = exp(-(de[0] * de[0] + de[1] * de[1] + de[2] * de[2]) / r2);
          const double u = de[0] * ua[0] + de[1] * ua[1] + de[2] * ua[2], v = de[0] * va[0] + de[1] * va[1] + de[2] * va[2];
          const double f = de[0] * pn[0] + de[1] * pn[1] + de[2] * pn[2];
          const double P[6] = {1.0, v, v * v, u, u * v, u * u}; /* (ui, vi): (0,0) (0,1) (0,2) (1,0) (1,1) (2,0) */
          for (int a = 0; a < 6; ++a) { rhs[a] += P[a] * wgt * f; for (int b = 0; b < 6; ++b) PtWP[a][b] += P[a] * wgt * P[b]; }
        }
        if (llt_solve6(PtWP, rhs)) { memcpy(c, rhs, sizeof(c)); have_poly = isfinite(c[0]); }
      }
      double pt[3], nv[3];
      if (have_poly) { /* projectPointSimpleToPolynomialSurface at (u, v) = (0, 0): w = c0, dz/du = c[3], dz/dv = c[1] */
        for (int k = 0; k < 3; ++k) { pt[k] = mean[k] + c[0] * pn[k]; nv[k] = pn[k] - c[3] * ua[k] - c[1] * va[k]; }
        const double l = sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]);
        for (int k = 0; k < 3; ++k) nv[k] /= l;
      } else { /* projectPointToMLSPlane */
        for (int k = 0; k < 3; ++k) { pt[k] = mean[k]; nv[k] = pn[k]; }
      }
      for (int k = 0; k < 3; ++k) { xyz_out[3 * i + k] = (float)pt[k]; nrm_out[3 * i + k] = (float)nv[k]; }
    }
    free(nn);
  }
}
