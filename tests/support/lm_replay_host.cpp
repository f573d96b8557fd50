// lm_replay_host.cpp -- TEST INFRASTRUCTURE: compiles the product's solver header (csrc/lm_replay.cuh) for the host so the
// CPU test-suite can pin it against the reference tree's own Eigen LM (oracle/_ref) without a GPU.
// Same signature as the oracle's LM backend hook (hop_oracle_set_lm_backend), so a whole oracle ICP can be run with it.
#include <cstdlib>
#include <cstring>
#define LMR_STATS 1
#include "../../icra20-hand-object-pose_b200/csrc/lm_replay.cuh"

static int g_acc_mode = 0;  // 0: float, 256 interleaved partial sums (like the kernel's lanes) 1: float sequential 2: double

extern "C" void hop_lmr_set_acc_mode(int m) { g_acc_mode = m; }

extern "C" void hop_lmr_moments(const float *src, const float *tgt, const float *nrm, int m, double *A_out /*13x13*/) {
  const int NP = g_acc_mode == 0 ? 256 : 1;
  double *accd = (double *)calloc((size_t)NP * 169, sizeof(double));
  float *accf = (float *)calloc((size_t)NP * 169, sizeof(float));
  for (int k = 0; k < m; ++k) {
    const float *s = src + 3 * k, *t = tgt + 3 * k, *n = nrm + 3 * k;
    float v[13];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) v[3 * i + j] = n[i] * s[j];
    v[9] = n[0]; v[10] = n[1]; v[11] = n[2];
    v[12] = n[0] * (s[0] - t[0]) + n[1] * (s[1] - t[1]) + n[2] * (s[2] - t[2]);
    const int p = k % NP;
    for (int i = 0; i < 13; ++i)
      for (int j = i; j < 13; ++j) {
        if (g_acc_mode == 2) accd[p * 169 + 13 * i + j] += (double)v[i] * (double)v[j];
        else accf[p * 169 + 13 * i + j] = fmaf(v[i], v[j], accf[p * 169 + 13 * i + j]);
      }
  }
  for (int i = 0; i < 13; ++i)
    for (int j = i; j < 13; ++j) {
      double s = 0.0;
      if (g_acc_mode == 2) s = accd[13 * i + j];
      else {  // pairwise tree over the partial sums, in float like the shuffle reduction
        float buf[256];
        for (int p = 0; p < NP; ++p) buf[p] = accf[p * 169 + 13 * i + j];
        for (int w = NP / 2; w >= 1; w /= 2)
          for (int p = 0; p < w; ++p) buf[p] += buf[p + w];
        s = buf[0];
      }
      A_out[13 * i + j] = s; A_out[13 * j + i] = s;
    }
  free(accd); free(accf);
}

extern "C" int hop_lmr_solve_moments(const double *A13, float *x, int *nfev) {
  lmr::Moments A;
  memcpy(A.a, A13, sizeof(A.a));
  if (lmr::translation_unconstrained(A)) {   // the kernel ends the ICP of this hypothesis "not converged"; emulate the slide for the oracle's loop
    x[0] = x[1] = x[2] = 1000.f;
    if (nfev) *nfev = 0;
    return -1;
  }
  return lmr::lm_replay_solve(A, x, nfev);
}

extern "C" int hop_lmr_point_to_plane(const float *src, const float *tgt, const float *nrm, int m, float *x, int *nfev) {
  if (m < lmr::N) { if (nfev) *nfev = 0; return 0; }
  double A13[169];
  hop_lmr_moments(src, tgt, nrm, m, A13);
  return hop_lmr_solve_moments(A13, x, nfev);
}

// work counters of the solves since the last reset: {solves, outer iterations, trial steps, lmpar calls that left the Gauss-Newton
// step (MINPACK's iteration on the LM parameter), qrsolv calls}
extern "C" void hop_lmr_stats(long long *out5, int reset) {
  lmr::Stats &s = lmr::stats();
  out5[0] = s.solves; out5[1] = s.outer; out5[2] = s.trials; out5[3] = s.lmpar_cold; out5[4] = s.qrsolv;
  if (reset) s = lmr::Stats{0, 0, 0, 0, 0};
}
