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

#if defined(LMR_STATS) && !defined(__CUDA_ARCH__)
struct Stats { long long solves, outer, trials, lmpar_cold, qrsolv; };   // host-side work counters (tools/lm_replay_eval.py)
inline Stats &stats() { static Stats s = {0, 0, 0, 0, 0}; return s; }
#define LMR_COUNT(f) (++lmr::stats().f)
#else
#define LMR_COUNT(f) ((void)0)
#endif

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
// Division and square root.  fdiv / fsqrt: correctly rounded, used where the reference's bits matter (W(x): the forward
// differences y(x + h e_j) - y(x) are a few ulps of R's entries when h = sqrt(eps)|x_j| is tiny).  qdiv / qsqrt: the MINPACK
// algebra on the 6x6 factors, where the replay differs from the reference's Householder QR in the last bits anyway: on the device
// a reciprocal / reciprocal square root of the special function unit (<= 2 ulp) instead of the 10-25 instruction IEEE sequences --
// the solve is one long dependent chain per hypothesis, its instruction count is its latency.
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
LMR_HD float qdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
  return a * r;
#else
  return a / b;
#endif
}
LMR_HD float qsqrt(float a) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
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
#if defined(__CUDACC__)
__device__ __noinline__ float norm6_dev(float a0, float a1, float a2, float a3, float a4, float a5) {
  double s = (double)a0 * (double)a0;
  s += (double)a1 * (double)a1; s += (double)a2 * (double)a2; s += (double)a3 * (double)a3;
  s += (double)a4 * (double)a4; s += (double)a5 * (double)a5;
  return qsqrt((float)s);
}
#endif
LMR_HD float norm6(const float *v) {
#if defined(__CUDA_ARCH__)
  return norm6_dev(v[0], v[1], v[2], v[3], v[4], v[5]);   // one copy of the code for the ~20 call sites
#else
  double s = 0.0;
  for (int i = 0; i < N; ++i) s += (double)v[i] * (double)v[i];
  return sqrtf((float)s);
#endif
}

// ---- the quantities lmdif reads from the residual vector and the Jacobian, from the moments -------------------------
// Host: plain loops over a full 13x13 double matrix.  Device: the 91 float sums stay in shared memory, lane i < 13 of the
// calling warp owns row i, the replicated results meet through shuffles (all 32 lanes must call, with uniform arguments).
#if defined(__CUDACC__)
// Per-warp scratch in shared memory (LMR_SCRATCH_BYTES, 16-byte aligned).  The solver's code must stay small: unrolled and
// shuffle-based it was ~10 k instructions (160 KB), more than the SM's instruction cache -- every warp then waits ~90 cycles per
// instruction on instruction fetch (ncu: stalled_no_instruction).  Exchanging the small vectors through shared memory instead of
// 64-bit shuffles and keeping the helpers out of line brings the hot loop to a few thousand instructions.
struct LmrScratch {
  double g[16];        // A y
  double U[12][6];     // rows of A (D h)
  double G[28];        // the 21 entries of the upper triangle of J^T J, then (lm_replay_warp.cuh) the 6 entries of J^T f
  float A[NY * NY];    // the moment matrix, full symmetric storage
  float D[N][12];      // columns of D h: y(x + h_j e_j) - y(x)
  float M[12][N];      // the same, transposed (row i = entry i of every column)
  float h[8];          // the steps h_j
  float R[N][N];       // lm_replay_warp.cuh: the float Cholesky factor (upper triangle, zeros below)
  double nb[2][8];     // lm_replay_warp.cuh: the squares of a distributed 6-vector on their way to its norm (double buffered)
  double Gc[8][N];     // lm_replay_warp.cuh: [J^T J | J^T f | 0] by columns, zeros below the diagonal (column k = the 6 numbers lane k owns)
};
#define LMR_SCRATCH_BYTES ((int)sizeof(lmr::LmrScratch))

struct MomentsDev {
  const float *sums;   // upper triangle of the 13x13 moment matrix, row-major (shared memory)
  int lane;
  LmrScratch *scr;     // this warp's scratch
  __device__ __forceinline__ double at(int i, int j) const { return (double)scr->A[i * NY + j]; }
};
// expand the packed upper triangle into full storage (all 32 lanes)
__device__ __forceinline__ void moments_prepare(const MomentsDev &A) {
  for (int e = A.lane; e < NY * NY; e += 32) {
    const int i = e / NY, j = e - i * NY;
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    A.scr->A[e] = A.sums[lo * 13 - (lo * (lo - 1)) / 2 + (hi - lo)];
  }
  __syncwarp();
}
// 1 / h in double: float reciprocal + one Newton step (relative error ~1e-14)
__device__ __forceinline__ double rcp_h(float h) {
  const double r = (double)__frcp_rn(h);
  return r * (2.0 - (double)h * r);
}
// g = A y (replicated), returns y^T A y
__device__ __noinline__ double quad(const MomentsDev &A, const float *y, double *g) {
  LmrScratch &S = *A.scr;
  const int li = A.lane < NY ? A.lane : NY - 1;
  double gi = 0.0;
#pragma unroll
  for (int j = 0; j < NY; ++j) gi = fma((double)S.A[li * NY + j], (double)y[j], gi);
  __syncwarp();
  if (A.lane < NY) S.g[A.lane] = gi;
  __syncwarp();
  double f2 = 0.0;
#pragma unroll
  for (int j = 0; j < NY; ++j) { g[j] = S.g[j]; f2 = fma(g[j], (double)y[j], f2); }
  return f2;
}
// Forward-difference Jacobian as moments: G = J^T J (upper triangle valid), b = J^T f.  Lane j < 6 evaluates W(x + h_j e_j);
// a translation column of D has the single entry (fl(x_j + h) - x_j) / h, a rotation column the nine entries of dR / h.
__device__ __noinline__ void gram(const MomentsDev &A, const float *x, const float *y, const double *g, float h_eps, double (*G)[N], double *b, float *h_out) {
  LmrScratch &S = *A.scr;
  const int lane = A.lane;
  float xx[N], hj = 1.f;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    float h = h_eps * fabsf(x[k]);
    if (h == 0.f) h = h_eps;
    xx[k] = lane == k ? add(x[k], h) : x[k];
    if (lane == k) hj = h;
    h_out[k] = h;
  }
  float yj[NY];
  warp_y(xx, yj);
  __syncwarp();
  if (lane < N) {
#pragma unroll
    for (int i = 0; i < 12; ++i) { const float d = sub(yj[i], y[i]); S.D[lane][i] = d; S.M[i][lane] = d; }
    S.h[lane] = hj;
  }
  __syncwarp();
  // this lane's row of A (D h): U[lane][j] = sum_l A[lane][l] (D h)[l][j]; a translation column has one entry, a rotation column nine
  if (lane < 12) {
#pragma unroll
    for (int c = 0; c < 3; ++c) S.U[lane][c] = (double)S.A[lane * NY + 9 + c] * (double)S.D[c][9 + c];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 9; ++k) s = fma((double)S.A[lane * NY + k], (double)S.D[3 + r][k], s);
      S.U[lane][3 + r] = s;
    }
  }
  __syncwarp();
  if (lane < 21) {   // entry (j, k), j <= k, of the upper triangle, row-major
    int j = 0, l = lane;
    while (l >= N - j) { l -= N - j; ++j; }
    const int k = j + l;
    double acc = 0.0;
#pragma unroll 1
    for (int i = 0; i < 12; ++i) acc = fma((double)S.M[i][j], S.U[i][k], acc);
    S.G[lane] = acc * (rcp_h(S.h[j]) * rcp_h(S.h[k]));
  }
  __syncwarp();
  {
    int e = 0;
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
      for (int k = j; k < N; ++k) G[j][k] = S.G[e++];
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) b[c] = (double)S.D[c][9 + c] * g[9 + c] * rcp_h(h_out[c]);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) s = fma((double)S.D[3 + r][k], g[k], s);
    b[3 + r] = s * rcp_h(h_out[3 + r]);
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

// Hot path (every LM step): small fully unrolled loops, everything in registers.  Cold path (lmpar's iteration on the LM
// parameter, entered only when the Gauss-Newton step leaves the trust region): rolled loops in a function of its own -- unrolled it
// is ~15 k instructions, more than the instruction cache holds, and every solve then stalls on instruction fetch.
#if defined(__CUDA_ARCH__)
#define LMR_UNROLL _Pragma("unroll")
#define LMR_ROLL _Pragma("unroll 1")
#define LMR_COLD __device__ __noinline__
#else
#define LMR_UNROLL
#define LMR_ROLL
#define LMR_COLD inline
#endif

// MINPACK qrsolv on the 6x6 upper-triangular r (identity column order): least squares of [R; D] z = [qtb; 0].  r's strict
// lower triangle receives the transposed factor S (lmpar's Newton correction reads it), sdiag its diagonal.
LMR_COLD void qrsolv(float (*r)[N], const float *diag, const float *qtb, float *x, float *sdiag) {
  LMR_COUNT(qrsolv);
  float wa[N];
  LMR_ROLL
  for (int j = 0; j < N; ++j) {
    LMR_ROLL
    for (int i = j; i < N; ++i) r[i][j] = r[j][i];
    x[j] = r[j][j];
    wa[j] = qtb[j];
  }
  LMR_ROLL
  for (int j = 0; j < N; ++j) {
    if (diag[j] != 0.f) {
      LMR_ROLL
      for (int k = j; k < N; ++k) sdiag[k] = 0.f;
      sdiag[j] = diag[j];
      float qtbpj = 0.f;
      LMR_ROLL
      for (int k = j; k < N; ++k) {
        if (sdiag[k] == 0.f) continue;
        float sn, cs;  // Givens rotation eliminating sdiag[k] against r[k][k] (Eigen's makeGivens differs only in rounding)
        if (fabsf(r[k][k]) < fabsf(sdiag[k])) { const float ct = qdiv(r[k][k], sdiag[k]); sn = qdiv(0.5f, qsqrt(0.25f + 0.25f * ct * ct)); cs = sn * ct; }
        else { const float tn = qdiv(sdiag[k], r[k][k]); cs = qdiv(0.5f, qsqrt(0.25f + 0.25f * tn * tn)); sn = cs * tn; }
        r[k][k] = cs * r[k][k] + sn * sdiag[k];
        const float tmp = cs * wa[k] + sn * qtbpj;
        qtbpj = -sn * wa[k] + cs * qtbpj;
        wa[k] = tmp;
        LMR_ROLL
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
  LMR_ROLL
  for (int j = 0; j < N; ++j) {
    if (sdiag[j] == 0.f && nsing == N) nsing = j;
    if (nsing < N) wa[j] = 0.f;
  }
  LMR_ROLL
  for (int j = N - 1; j >= 0; --j) {
    if (j < nsing) {
      float sum = 0.f;
      LMR_ROLL
      for (int i = j + 1; i < N; ++i) sum += r[i][j] * wa[i];   // wa[i] = 0 beyond nsing
      wa[j] = qdiv(wa[j] - sum, sdiag[j]);
    }
  }
  LMR_ROLL
  for (int j = 0; j < N; ++j) x[j] = wa[j];
}

// MINPACK lmpar: the LM parameter par with | |D x| - delta | <= 0.1 delta (or par = 0 when the Gauss-Newton step fits)
LMR_COLD void lmpar_iterate(float (*r)[N], const float *diag, const float *qtb, float delta, float &par, float *x) {
  const float dwarf = FLT_MIN;
  float wa1[N], wa2[N], sdiag[N];
  int nsing = N;
  LMR_ROLL
  for (int j = 0; j < N; ++j) {
    wa1[j] = qtb[j];
    if (r[j][j] == 0.f && nsing == N) nsing = j;
    if (nsing < N) wa1[j] = 0.f;
  }
  LMR_ROLL
  for (int j = N - 1; j >= 0; --j) {
    if (j < nsing) {
      wa1[j] = qdiv(wa1[j], r[j][j]);
      const float t = wa1[j];
      LMR_ROLL
      for (int i = 0; i < j; ++i) wa1[i] -= r[i][j] * t;
    }
  }
  LMR_ROLL
  for (int j = 0; j < N; ++j) { x[j] = wa1[j]; wa2[j] = diag[j] * x[j]; }
  int iter = 0;
  float dxnorm = norm6(wa2);
  float fp = dxnorm - delta;
  if (fp <= 0.1f * delta) { par = 0.f; return; }
  LMR_COUNT(lmpar_cold);
  float parl = 0.f;
  if (nsing >= N) {
    LMR_ROLL
    for (int j = 0; j < N; ++j) wa1[j] = diag[j] * qdiv(wa2[j], dxnorm);
    LMR_ROLL
    for (int j = 0; j < N; ++j) {
      float sum = 0.f;
      LMR_ROLL
      for (int i = 0; i < j; ++i) sum += r[i][j] * wa1[i];
      wa1[j] = qdiv(wa1[j] - sum, r[j][j]);
    }
    const float t = norm6(wa1);
    parl = qdiv(qdiv(qdiv(fp, delta), t), t);
  }
  LMR_ROLL
  for (int j = 0; j < N; ++j) {
    float sum = 0.f;
    LMR_ROLL
    for (int i = 0; i <= j; ++i) sum += r[i][j] * qtb[i];
    wa1[j] = qdiv(sum, diag[j]);
  }
  const float gnorm = norm6(wa1);
  float paru = qdiv(gnorm, delta);
  if (paru == 0.f) paru = qdiv(dwarf, fminf(delta, 0.1f));
  par = fmaxf(par, parl);
  par = fminf(par, paru);
  if (par == 0.f) par = qdiv(gnorm, dxnorm);
  for (;;) {
    ++iter;
    if (par == 0.f) par = fmaxf(dwarf, 0.001f * paru);
    const float sq = qsqrt(par);
    LMR_ROLL
    for (int j = 0; j < N; ++j) wa1[j] = sq * diag[j];
    float rr[N][N];
    LMR_ROLL
    for (int i = 0; i < N; ++i)
      LMR_ROLL
      for (int j = 0; j < N; ++j) rr[i][j] = r[i][j];
    qrsolv(rr, wa1, qtb, x, sdiag);
    LMR_ROLL
    for (int j = 0; j < N; ++j) wa2[j] = diag[j] * x[j];
    dxnorm = norm6(wa2);
    float temp = fp;
    fp = dxnorm - delta;
    if (fabsf(fp) <= 0.1f * delta || (parl == 0.f && fp <= temp && temp < 0.f) || iter == 10) break;
    LMR_ROLL
    for (int j = 0; j < N; ++j) wa1[j] = diag[j] * qdiv(wa2[j], dxnorm);
    LMR_ROLL
    for (int j = 0; j < N; ++j) {
      wa1[j] = qdiv(wa1[j], sdiag[j]);
      const float t = wa1[j];
      LMR_ROLL
      for (int i = j + 1; i < N; ++i) wa1[i] -= rr[i][j] * t;
    }
    temp = norm6(wa1);
    const float parc = qdiv(qdiv(qdiv(fp, delta), temp), temp);
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

// lmpar, hot part: the Gauss-Newton step and the test that it fits the trust region (then par = 0: the common case).  Otherwise
// the whole of MINPACK's lmpar runs in the cold function, on copies (so that r stays in registers here).
LMR_HD void lmpar(const float (*r)[N], const float *diag, const float *qtb, float delta, float &par, float *x) {
  float wa1[N], wa2[N];
  bool full_rank = true;
  LMR_UNROLL
  for (int j = 0; j < N; ++j) { wa1[j] = qtb[j]; full_rank = full_rank && r[j][j] != 0.f; }
  if (full_rank) {
    LMR_UNROLL
    for (int j = N - 1; j >= 0; --j) {
      wa1[j] = qdiv(wa1[j], r[j][j]);
      const float t = wa1[j];
      LMR_UNROLL
      for (int i = 0; i < j; ++i) wa1[i] -= r[i][j] * t;
    }
    LMR_UNROLL
    for (int j = 0; j < N; ++j) wa2[j] = diag[j] * wa1[j];
    const float fp = norm6(wa2) - delta;
    if (fp <= 0.1f * delta) {
      LMR_UNROLL
      for (int j = 0; j < N; ++j) x[j] = wa1[j];
      par = 0.f;
      return;
    }
  }
  float rc[N][N], dc[N], qc[N], xc[N];
  LMR_UNROLL
  for (int i = 0; i < N; ++i) {
    dc[i] = diag[i]; qc[i] = qtb[i];
    LMR_UNROLL
    for (int j = 0; j < N; ++j) rc[i][j] = r[i][j];
  }
  lmpar_iterate(rc, dc, qc, delta, par, xc);
  LMR_UNROLL
  for (int j = 0; j < N; ++j) x[j] = xc[j];
}

// sqrt and reciprocal of a Cholesky pivot (0 for a non-positive one); out of line on the device: one copy for the six pivots
#if defined(__CUDACC__)
__device__ __forceinline__ void chol_pivot_dev(double d, double *dd, double *inv) {
  const double ri = d > 0.0 ? rsqrt(d) : 0.0;   // (2 ulp; the factor only has to be good to float precision)
  *inv = ri; *dd = d > 0.0 ? d * ri : 0.0;
}
#endif
LMR_HD void chol_pivot(double d, double &dd, double &inv) {
#if defined(__CUDA_ARCH__)
  chol_pivot_dev(d, &dd, &inv);
#else
  dd = d > 0.0 ? sqrt(d) : 0.0;
  inv = dd > 0.0 ? 1.0 / dd : 0.0;
#endif
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
  float fnorm = qsqrt(f2 > 0.0 ? (float)f2 : 0.f);
  LMR_COUNT(solves);
  for (;;) {
    LMR_COUNT(outer);
    double G[N][N], b[N];
    float hs[N];
    gram(A, x, y, g, h_eps, G, b, hs);
    nfev += N + 1;   // NumericalDiff::df (Forward) re-evaluates f(x) first: n + 1 evaluations
    LMR_UNROLL
    for (int j = 0; j < N; ++j) wa2[j] = qsqrt(G[j][j] > 0.0 ? (float)G[j][j] : 0.f);
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
        double dd, inv;
        chol_pivot(d, dd, inv);
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
          for (int i = 0; i <= j; ++i) s += r[i][j] * qdiv(qtf[i], fnorm);
          gnorm = fmaxf(gnorm, fabsf(qdiv(s, wa2[j])));
        }
    }
    if (gnorm <= gtol) { status = 4; break; }
    LMR_UNROLL
    for (int j = 0; j < N; ++j) diag[j] = fmaxf(diag[j], wa2[j]);
    float ratio;
    bool done = false;
    do {
      LMR_COUNT(trials);
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
      const float fnorm1 = qsqrt(f21 > 0.0 ? (float)f21 : 0.f);
      float actred = -1.f;
      if (0.1f * fnorm1 < fnorm) { const float q = qdiv(fnorm1, fnorm); actred = 1.f - q * q; }
      LMR_UNROLL
      for (int i = 0; i < N; ++i) {
        float s = 0.f;
        LMR_UNROLL
        for (int j = i; j < N; ++j) s += r[i][j] * wa1[j];
        wa3[i] = s;
      }
      float t1 = qdiv(norm6(wa3), fnorm); t1 *= t1;
      float t2 = qdiv(qsqrt(par) * pnorm, fnorm); t2 *= t2;
      const float prered = t1 + qdiv(t2, 0.5f);
      const float dirder = -(t1 + t2);
      ratio = 0.f;
      if (prered != 0.f) ratio = qdiv(actred, prered);
      if (ratio <= 0.25f) {
        float temp = 0.5f;
        if (actred < 0.f) temp = qdiv(0.5f * dirder, dirder + 0.5f * actred);
        if (0.1f * fnorm1 >= fnorm || temp < 0.1f) temp = 0.1f;
        delta = temp * fminf(delta, qdiv(pnorm, 0.1f));
        par = qdiv(par, temp);
      } else if (!(par != 0.f && ratio < 0.75f)) {
        delta = qdiv(pnorm, 0.5f);
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
