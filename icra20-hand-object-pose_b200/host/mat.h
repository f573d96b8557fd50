// mat.h -- the few fixed-size float matrices the host needs (the reference uses Eigen; the only Eigen in this image is the
// reference's vendored copy, which product code must not depend on).  Column-major 4x4 = Eigen::Matrix4f::data() layout.
#pragma once
#include <cmath>
#include <cstring>
#include <iomanip>
#include <ostream>

struct Mat3f {
  float m[9];  // column-major
  float &operator()(int r, int c) { return m[3 * c + r]; }
  float operator()(int r, int c) const { return m[3 * c + r]; }
};

struct Mat4f {
  float m[16];  // column-major
  Mat4f() { setIdentity(); }
  void setIdentity() { std::memset(m, 0, sizeof(m)); m[0] = m[5] = m[10] = m[15] = 1.f; }
  float &operator()(int r, int c) { return m[4 * c + r]; }
  float operator()(int r, int c) const { return m[4 * c + r]; }
  const float *data() const { return m; }
  float *data() { return m; }
  Mat4f operator*(const Mat4f &b) const {
    Mat4f r;
    for (int c = 0; c < 4; ++c)
      for (int i = 0; i < 4; ++i) {
        float s = 0.f;
        for (int k = 0; k < 4; ++k) s += (*this)(i, k) * b(k, c);
        r(i, c) = s;
      }
    return r;
  }
  // general inverse (Gauss-Jordan in double, partial pivoting)
  Mat4f inverse() const {
    double a[4][8];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) { a[r][c] = (*this)(r, c); a[r][c + 4] = r == c; }
    for (int i = 0; i < 4; ++i) {
      int p = i;
      for (int r = i + 1; r < 4; ++r) if (std::fabs(a[r][i]) > std::fabs(a[p][i])) p = r;
      if (p != i) for (int c = 0; c < 8; ++c) std::swap(a[i][c], a[p][c]);
      const double inv = 1.0 / a[i][i];
      for (int c = 0; c < 8; ++c) a[i][c] *= inv;
      for (int r = 0; r < 4; ++r) if (r != i) { const double f = a[r][i]; if (f != 0.0) for (int c = 0; c < 8; ++c) a[r][c] -= f * a[i][c]; }
    }
    Mat4f out;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out(r, c) = (float)a[r][c + 4];
    return out;
  }
};

inline std::ostream &operator<<(std::ostream &os, const Mat4f &M) {  // row by row, like Eigen's default IOFormat
  for (int r = 0; r < 4; ++r) {
    for (int c = 0; c < 4; ++c) os << (c ? " " : "") << std::setw(12) << M(r, c);
    if (r < 3) os << "\n";
  }
  return os;
}

// Eigen::Quaternionf(w, x, y, z).normalized().toRotationMatrix()
inline void quat_to_rot(float w, float x, float y, float z, Mat4f &T) {
  const float n = std::sqrt(w * w + x * x + y * y + z * z);
  w /= n; x /= n; y /= n; z /= n;
  const float tx = 2 * x, ty = 2 * y, tz = 2 * z, twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x,
              tyy = ty * y, tyz = tz * y, tzz = tz * z;
  T(0, 0) = 1 - (tyy + tzz); T(0, 1) = txy - twz; T(0, 2) = txz + twy;
  T(1, 0) = txy + twz; T(1, 1) = 1 - (txx + tzz); T(1, 2) = tyz - twx;
  T(2, 0) = txz - twy; T(2, 1) = tyz + twx; T(2, 2) = 1 - (txx + tyy);
}
