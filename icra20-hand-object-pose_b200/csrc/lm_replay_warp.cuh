// lm_replay_warp.cuh -- the LM replay of lm_replay.cuh, laid out over the lanes of the calling warp (device only).
//
// lm_replay_solve (lm_replay.cuh) is the scalar program: on the device every lane would execute all of it redundantly, ~2.2 k
// dependent instructions per outer LM iteration with the 6x6 double matrices spilling at the fused kernel's register cap --
// ~26 k cycles per outer iteration (ncu source page, profiles/r02_*).  The solve is a latency chain, and on small batches the
// slowest hypothesis' chain IS the kernel time.  Here the 6-vectors and the 6x6 factors are DISTRIBUTED:
//   lane j (mod 8)  owns parameter j: its diag entry, column j of J^T J through the Cholesky factorisation (lane 6: the column
//                   J^T f, which the same elimination turns into Q^T f), row j of R for the back substitution;
//   lane i (mod 16) owns row i of the 13x13 moment matrix (shared memory) for y^T A y.
// Scalars of MINPACK's control flow (fnorm, par, delta, ratio ...) are replicated.  Reductions either meet in shared memory and are
// added by every lane in the same order, or are xor butterflies (bitwise the same in all lanes: a + b == b + a at every level), so the
// warp never diverges on them.  The float operations of MINPACK's algebra keep the order of the scalar program; sums carried in double
// (norms, Gram products) associate differently, which moves a float result by at most its last bit.  lmpar's iteration on the LM
// parameter (the Gauss-Newton step leaves the trust region: 3 % of the trial steps, but most of those of the hardest hypotheses, where
// the replicated rolled version cost 15-30 k cycles per qrsolv) is distributed the same way, and its inner solve is the Cholesky
// factorisation of J^T J + par D^2 instead of qrsolv's chain of Givens rotations (lmpar_iterate_w); only a rank-deficient R falls
// back to lm_replay.cuh's scalar lmpar_iterate, fed from the factor in shared memory.
//
// The scalar program stays the specification: tests/test_gpu_lm.py solves the same moment matrices with this file (through
// hop_debug_lm_solve) and with the host build of lm_replay.cuh, which tests/test_lm_replay.py pins against the reference tree's
// own Eigen LM.
#pragma once
#include "lm_replay.cuh"

#if defined(__CUDACC__)
namespace lmr {

__device__ __forceinline__ double bfly_sum16(double v) {
#pragma unroll
  for (int o = 8; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double bfly_sum8(double v) {
#pragma unroll
  for (int o = 4; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float bfly_max8(float v) {
#pragma unroll
  for (int o = 4; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// entry l of a replicated short vector (l is lane dependent: a select chain, no local memory)
template <typename T, int LEN>
__device__ __forceinline__ T pick(const T (&v)[LEN], int l) {
  T r = (T)0;
#pragma unroll
  for (int j = 0; j < LEN; ++j) r = (l == j) ? v[j] : r;
  return r;
}
// 2-norm of a 6-vector whose entry j sits in lane j (mod 8).  The squares meet in shared memory and every lane adds them in index
// order (the scalar program's norm6; bitwise the same in all lanes): 13 instructions against 30 for a butterfly of 64-bit shuffles.
// Two buffers in turn: the barrier of call n + 1 separates the reads of call n from the writes of call n + 2.
struct NormBuf {
  double (*nb)[8];
  int tog;
};
__device__ __forceinline__ float norm6_w(float v, int lane, NormBuf &B) {
  double *b = B.nb[B.tog];
  B.tog ^= 1;
  if (lane < N) b[lane] = (double)v * (double)v;
  __syncwarp();
  double s = b[0];
#pragma unroll
  for (int i = 1; i < N; ++i) s += b[i];
  return qsqrt((float)s);
}
// y^T A y, lane (mod 16) on row l16 of A (shared memory; the loads do not depend on y and issue ahead of the chain);
// g_own = (A y)_row, 0 on the lanes without a row
__device__ __forceinline__ double quad_w(const float *Arow, const float (&y)[NY], int l16, double &g_own) {
  double s0 = 0.0, s1 = 0.0;
#pragma unroll
  for (int j = 0; j + 1 < NY; j += 2) {
    s0 = fma((double)Arow[j], (double)y[j], s0);
    s1 = fma((double)Arow[j + 1], (double)y[j + 1], s1);
  }
  s0 = fma((double)Arow[NY - 1], (double)y[NY - 1], s0);
  g_own = l16 < NY ? s0 + s1 : 0.0;
  return bfly_sum16(g_own * (double)pick(y, l16));
}

// Right-looking Cholesky of a 6x6 matrix with one right-hand side, column l8 per lane (lanes 0..5: columns of M, upper part; lane 6: the
// right-hand side b; lane 7: zeros).  In: c[i] = M(i, l8) for i <= l8 (0 below) / b(i); dgl = the diagonal of M, replicated.
// Out: c[i] = S(i, l8) for i <= l8 with M = S^T S / (S^-T b)(i), rounded to float on the way like the scalar program's Q^T f.
// Row i of the factor, right of the pivot, travels through S.U[i][.] (which is idle after the Gram step) and stays there: the row
// owners read their row of S from it afterwards.  Every lane tracks the whole diagonal of the block still to eliminate, so the
// pivot needs no broadcast.  A non-positive pivot zeroes its row (chol_pivot).
__device__ __forceinline__ void chol_w(LmrScratch &S, double (&c)[N], double (&dgl)[N], int lane, int l8) {
  __syncwarp();   // (the rows of the previous factor in S.U have been read)
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double dd, inv;
    chol_pivot_dev(dgl[i], &dd, &inv);
    double rik = l8 == i ? dd : c[i] * inv;
    if (l8 == N) rik = (double)(float)rik;
    c[i] = rik;
    if (i + 1 < N) {
      if (lane > i && lane < N) S.U[i][lane] = rik;
      __syncwarp();
#pragma unroll
      for (int m = i + 1; m < N; ++m) {
        const double rim = S.U[i][m];
        if (l8 >= m) c[m] = fma(-rim, rik, c[m]);
        dgl[m] = fma(-rim, rim, dgl[m]);
      }
    }
  }
}
// after chol_w: column l8 of the float factor (rc), its row l8 (rr, from S.U), S^-T b replicated (q, through S.h)
__device__ __forceinline__ void chol_fetch(LmrScratch &S, const double (&c)[N], float (&rc)[N], float (&rr)[N], float (&q)[N], int lane, int l8) {
  const bool own6 = l8 < N;
#pragma unroll
  for (int i = 0; i < N; ++i) rc[i] = (own6 && i <= l8) ? (float)c[i] : 0.f;
  if (lane == N) {
#pragma unroll
    for (int i = 0; i < N; ++i) S.h[i] = (float)c[i];   // (the steps h_j are not read again before the next Gram step rewrites them)
  }
  __syncwarp();
  const float rdiag = pick(rc, l8);
#pragma unroll
  for (int k = 0; k < N; ++k) {
    q[k] = S.h[k];
    rr[k] = (own6 && k > l8) ? (float)S.U[l8 < N ? l8 : 0][k] : (k == l8 ? rdiag : 0.f);
  }
  __syncwarp();   // (S.h may be rewritten by the next factorisation)
}
// R z = q by back substitution, row l8 of R per lane (rr), rinv = 1 / R(l8, l8); z replicated.  MINPACK's order: z(k) is final
// before it is subtracted from the rows above.
__device__ __forceinline__ void backsolve_w(const float (&rr)[N], float rinv, const float (&q)[N], float (&z)[N], int l8) {
  float acc = pick(q, l8);
#pragma unroll
  for (int k = N - 1; k >= 0; --k) {
    z[k] = __shfl_sync(0xffffffffu, acc * rinv, k);
    if (l8 < k) acc = fmaf(-rr[k], z[k], acc);
  }
}

// MINPACK lmpar beyond the Gauss-Newton step (full-rank R): the iteration on the LM parameter, distributed like the rest.
// The scalar program's qrsolv eliminates the rows sqrt(par) D of [R; sqrt(par) D] with 21 Givens rotations, one after the other, to get
// the factor S of R^T R + par D^2 and the step.  R^T R is J^T J and R^T Q^T f is J^T f, both still in shared memory from the Gram step: S is
// the Cholesky factor of J^T J + par D^2 and the step solves S^T S x = J^T f -- the factorisation this file already has, ~800 cycles
// instead of ~6 k for the rotation chain (in double, so at least as close to the exact step as the reference's float rotations).
// rc = column l8 of R, rd = R(l8, l8), dg = diag(l8), qtf replicated; step (replicated) holds the Gauss-Newton step on entry and the LM
// step on exit; dxnorm = |D step| of the Gauss-Newton step; g_col = this lane's column of [J^T J | J^T f] (S.Gc).
__device__ __forceinline__ void lmpar_iterate_w(LmrScratch &S, const double *g_col, const float (&rc)[N], float rd, float dg, const float (&qtf)[N], float delta,
                                                float dxnorm, float &par, float (&step)[N], int lane, NormBuf &NB) {
  const unsigned FULL = 0xffffffffu;
  const int l8 = lane & 7;
  const bool own6 = l8 < N;
  const float dwarf = FLT_MIN;
  float fp = dxnorm - delta;
  float wa2 = dg * pick(step, l8);
  // parl: |R^-T D^2 x / |D x||^-2 scaled (Newton step on phi at par = 0)
  float parl;
  {
    const float w = dg * qdiv(wa2, dxnorm);
    float sum = 0.f, fin_own = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const float fin = qdiv(w - sum, rd);                 // final on lane i: its sum over the rows above is complete
      const float vi = __shfl_sync(FULL, fin, i);
      if (l8 == i) fin_own = fin;
      if (l8 > i) sum = fmaf(rc[i], vi, sum);
    }
    const float t = norm6_w(fin_own, lane, NB);
    parl = qdiv(qdiv(qdiv(fp, delta), t), t);
  }
  float gs = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i)
    if (i <= l8) gs = fmaf(rc[i], qtf[i], gs);
  const float gnorm = norm6_w(own6 ? qdiv(gs, dg) : 0.f, lane, NB);
  float paru = qdiv(gnorm, delta);
  if (paru == 0.f) paru = qdiv(dwarf, fminf(delta, 0.1f));
  par = fmaxf(par, parl);
  par = fminf(par, paru);
  if (par == 0.f) par = qdiv(gnorm, dxnorm);
  double d2[N];   // diag^2, replicated
#pragma unroll
  for (int i = 0; i < N; ++i) { const double d = (double)__shfl_sync(FULL, dg, i); d2[i] = d * d; }
  int iter = 0;
  for (;;) {
    ++iter;
    if (par == 0.f) par = fmaxf(dwarf, 0.001f * paru);
    double c[N], dgl[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double add = (double)par * d2[i];
      dgl[i] = S.Gc[i][i] + add;
      c[i] = g_col[i] + ((own6 && i == l8) ? add : 0.0);
    }
    chol_w(S, c, dgl, lane, l8);
    float sc[N], sr[N], sq[N];
    chol_fetch(S, c, sc, sr, sq, lane, l8);
    const float sdiag = pick(sc, l8);
    backsolve_w(sr, qdiv(1.f, sdiag), sq, step, l8);
    wa2 = dg * pick(step, l8);
    dxnorm = norm6_w(wa2, lane, NB);
    const float temp0 = fp;
    fp = dxnorm - delta;
    if (fabsf(fp) <= 0.1f * delta || (parl == 0.f && fp <= temp0 && temp0 < 0.f) || iter == 10) break;
    float w = dg * qdiv(wa2, dxnorm), fin_own = 0.f;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const float fin = qdiv(w, sdiag);                    // final on lane j
      const float t = __shfl_sync(FULL, fin, j);
      if (l8 == j) fin_own = fin;
      if (l8 > j) w = fmaf(-sc[j], t, w);
    }
    const float temp = norm6_w(fin_own, lane, NB);
    const float parc = qdiv(qdiv(qdiv(fp, delta), temp), temp);
    if (fp > 0.f) parl = fmaxf(parl, par);
    if (fp < 0.f) paru = fminf(paru, par);
    par = fmaxf(parl, par + parc);
  }
}

// The LM run from x = 0.  All 32 lanes call with the same arguments; A.scr->A must hold the expanded moments (moments_prepare).
// x_out (6, replicated) receives the minimiser the reference's LM stops at; returns Eigen's LevenbergMarquardtSpace status.
__device__ __noinline__ int lm_replay_solve_warp(const MomentsDev &A, float *x_out, int *nfev_out) {
  LmrScratch &S = *A.scr;
  const unsigned FULL = 0xffffffffu;
  const int lane = A.lane, l8 = lane & 7, l16 = lane & 15;
  const bool own6 = l8 < N;
  NormBuf NB{S.nb, 0};
  const float eps = FLT_EPSILON;
  const float h_eps = 3.4526698e-4f, ftol = h_eps, xtol = h_eps, gtol = 0.f, factor = 100.f;   // sqrt(FLT_EPSILON)
  const int maxfev = 400;
  const float *Ar = S.A + (l16 < NY ? l16 : NY - 1) * NY;   // this lane's row of the moment matrix
  // this lane's entry of [J^T J | J^T f] in the Gram step: (gj, gk) of the upper triangle for lanes 0..20, J^T f entry gj for 21..26
  int gj = 0, gk = 0;
  if (lane < 21) {
    int l = lane;
    while (l >= N - gj) { l -= N - gj; ++gj; }
    gk = gj + l;
  } else if (lane < 27) gj = gk = lane - 21;
  double *const g_dst = &S.Gc[gk + (lane >= 21 ? N - gk : 0)][gj];   // where this lane's entry goes: column gk row gj / column 6 row gj
  const double *const g_col = S.Gc[l8];                              // the column this lane owns in the factorisation
  for (int e = lane; e < 8 * N; e += 32) (&S.Gc[0][0])[e] = 0.0;    // (only the upper triangle is ever written again)
  __syncwarp();
  float x[N] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, y[NY];
  float dg = 0.f;   // diag[l8]
  double g_own;
  int nfev = 1, iter = 1, status = 0;
  float par = 0.f, delta = 0.f, xnorm = 0.f;
  warp_y(x, y);
  const double f2 = quad_w(Ar, y, l16, g_own);
  float fnorm = qsqrt(f2 > 0.0 ? (float)f2 : 0.f);
  for (;;) {
    // ---- forward-difference Jacobian as moments.  Lane j < 6 evaluates W(x + h_j e_j); a translation column of D has the single
    //      entry (fl(x_j + h) - x_j) / h, a rotation column the nine entries of dR / h ----
    {
      float hj = h_eps * fabsf(pick(x, l8));   // lanes 6, 7: x = 0 there, they evaluate W(x + h e_none) = W(x) for nothing
      if (hj == 0.f) hj = h_eps;
      const float xp = add(pick(x, l8), hj);
      float xx[N];
#pragma unroll
      for (int k = 0; k < N; ++k) xx[k] = l8 == k ? xp : x[k];
      float yj[NY];
      warp_y(xx, yj);
      __syncwarp();
      if (lane < N) {
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          const float d = sub(yj[i], y[i]);
          S.D[lane][i] = d;
          S.M[i][lane] = d;
        }
        S.h[lane] = hj;
      }
      if (lane < NY) S.g[lane] = g_own;
      __syncwarp();
      if (lane < 12) {   // row `lane` of A (D h)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) S.U[lane][cc] = (double)Ar[9 + cc] * (double)S.D[cc][9 + cc];
        double u0 = 0.0, u1 = 0.0, u2 = 0.0;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const double a = (double)Ar[k];
          u0 = fma(a, (double)S.D[3][k], u0);
          u1 = fma(a, (double)S.D[4][k], u1);
          u2 = fma(a, (double)S.D[5][k], u2);
        }
        S.U[lane][3] = u0; S.U[lane][4] = u1; S.U[lane][5] = u2;
      }
      __syncwarp();
      if (lane < 27) {
        double a0 = 0.0, a1 = 0.0, scale;
        if (lane < 21) {   // (J^T J)(gj, gk) = (D h)_gj . A (D h)_gk / (h_gj h_gk)
#pragma unroll
          for (int i = 0; i < 12; i += 2) {
            a0 = fma((double)S.M[i][gj], S.U[i][gk], a0);
            a1 = fma((double)S.M[i + 1][gj], S.U[i + 1][gk], a1);
          }
          scale = rcp_h(S.h[gj]) * rcp_h(S.h[gk]);
        } else {           // (J^T f)(gj) = (D h)_gj . (A y) / h_gj   (the entries a column does not have are exact zeros)
#pragma unroll
          for (int i = 0; i < 12; i += 2) {
            a0 = fma((double)S.D[gj][i], S.g[i], a0);
            a1 = fma((double)S.D[gj][i + 1], S.g[i + 1], a1);
          }
          scale = rcp_h(S.h[gj]);
        }
        *g_dst = (a0 + a1) * scale;
      }
      __syncwarp();
    }
    nfev += N + 1;   // NumericalDiff::df (Forward) re-evaluates f(x) first: n + 1 evaluations
    // ---- column l8 of [J^T J | J^T f] -> column l8 of [R | Q^T f].  R and Q^T f of the QR of J = Cholesky factor of J^T J and
    //      R^-T J^T f (lm_replay.cuh); right-looking elimination, row i of the factor travels through shared memory ----
    double c[N], dgl[N];   // dgl: the diagonal of the block still to eliminate, replicated (every lane sees every row of the factor go by)
#pragma unroll
    for (int i = 0; i < N; ++i) { c[i] = g_col[i]; dgl[i] = S.Gc[i][i]; }
    const double gdiag = pick(c, l8);   // (J^T J)(l8, l8); 0 on lanes 6, 7
    const float wa2 = qsqrt(gdiag > 0.0 ? (float)gdiag : 0.f);   // column norm of J
    chol_w(S, c, dgl, lane, l8);
    float rc[N], qtf[N], rr[N];   // column l8 of r, Q^T f (replicated), row l8 of r
    chol_fetch(S, c, rc, rr, qtf, lane, l8);
    const float rdiag = pick(rc, l8);
    if (iter == 1) {
      dg = own6 ? (wa2 == 0.f ? 1.f : wa2) : 0.f;
      xnorm = norm6_w(dg * pick(x, l8), lane, NB);
      delta = factor * xnorm;
      if (delta == 0.f) delta = factor;
    }
    float gnorm = 0.f;
    if (fnorm != 0.f) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < N; ++i)
        if (i <= l8) s = fmaf(rc[i], qdiv(qtf[i], fnorm), s);
      gnorm = bfly_max8((own6 && wa2 != 0.f) ? fabsf(qdiv(s, wa2)) : 0.f);
    }
    if (gnorm <= gtol) { status = 4; break; }
    dg = fmaxf(dg, wa2);
    const float rinv = qdiv(1.f, rdiag);
    float ratio;
    bool done = false;
    do {
      // ---- lmpar: the Gauss-Newton step and the test that it fits the trust region (then par = 0: the common case) ----
      float step[N];
      if (__all_sync(FULL, !own6 || rdiag != 0.f)) {
        backsolve_w(rr, rinv, qtf, step, l8);
        const float dxnorm = norm6_w(dg * pick(step, l8), lane, NB);
        if (dxnorm - delta <= 0.1f * delta) par = 0.f;
        else lmpar_iterate_w(S, g_col, rc, rdiag, dg, qtf, delta, dxnorm, par, step, lane, NB);
      } else {   // rank-deficient R (a zero pivot): the scalar program, from the factor in shared memory
        __syncwarp();
        if (lane < N) {
#pragma unroll
          for (int i = 0; i < N; ++i) S.R[i][lane] = rc[i];
        }
        __syncwarp();
        float rc2[N][N], dc[N], qc[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
          dc[i] = __shfl_sync(FULL, dg, i);
          qc[i] = qtf[i];
#pragma unroll
          for (int j = 0; j < N; ++j) rc2[i][j] = S.R[i][j];
        }
        lmpar_iterate(rc2, dc, qc, delta, par, step);
      }
      float x1[N];   // p = -step
#pragma unroll
      for (int j = 0; j < N; ++j) x1[j] = x[j] + -step[j];
      const float pnorm = norm6_w(dg * -pick(step, l8), lane, NB);
      if (iter == 1) delta = fminf(delta, pnorm);
      float y1[NY];
      double g1_own;
      warp_y(x1, y1);
      const double f21 = quad_w(Ar, y1, l16, g1_own);
      ++nfev;
      const float fnorm1 = qsqrt(f21 > 0.0 ? (float)f21 : 0.f);
      float actred = -1.f;
      if (0.1f * fnorm1 < fnorm) { const float q = qdiv(fnorm1, fnorm); actred = 1.f - q * q; }
      float rp = 0.f;   // (R p)(l8)
#pragma unroll
      for (int j = 0; j < N; ++j)
        if (j >= l8) rp = fmaf(rr[j], -step[j], rp);
      float t1 = qdiv(norm6_w(rp, lane, NB), fnorm); t1 *= t1;
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
#pragma unroll
        for (int j = 0; j < N; ++j) x[j] = x1[j];
#pragma unroll
        for (int i = 0; i < NY; ++i) y[i] = y1[i];
        g_own = g1_own;
        xnorm = norm6_w(dg * pick(x, l8), lane, NB);
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
#pragma unroll
  for (int j = 0; j < N; ++j) x_out[j] = x[j];
  if (nfev_out) *nfev_out = nfev;
  return status;
}

}  // namespace lmr
#endif  // __CUDACC__
