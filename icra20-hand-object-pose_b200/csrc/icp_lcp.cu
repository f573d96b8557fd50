// icp_lcp.cu -- K4 (per-hypothesis ICP refinement) and K5 (per-hypothesis LCP scoring) for sm_100a.
//
//   K4 replaces Utils::runICP as called by PoseEstimator::refineByICP  (Utils.cpp:188-229, PoseEstimator.cpp:257-273)
//   K5 replaces Utils::computeLCP as called by PoseEstimator::selectBest (Utils.cpp:372-444, PoseEstimator.cpp:474-498)
//
// Design (B200-first, not a translation):
//   * The reference transforms the MODEL by every hypothesis and rebuilds a kd-tree of it.  Here the model and its
//     nearest-neighbour grid never move; the SCENE is carried into the model frame by X = pose^-1 and ICP iterates
//     on X.  The point-to-plane objective is frame invariant, and the refined pose is simply X_final^-1
//     ( = T_icp^-1 * pose of PoseEstimator.cpp:267 ).
//   * The work is gather bound (one voxel cell + a short candidate list + one normal per scene point, all L2
//     resident), so it is laid out FLAT: one thread per (hypothesis, scene point) at full occupancy, instead of a
//     warp that owns a hypothesis for the whole ICP (which leaves the SMs latency bound at small batch sizes and
//     pays the slowest hypothesis' tail).  Each ICP iteration is two launches:
//        icp_correspond_kernel  grid (scene tiles, hypotheses): nearest neighbour + both rejectors, writes one
//                               32-byte correspondence record per scene point (coalesced, two float4 planes);
//        icp_solve_kernel       one warp team per hypothesis: streams the records, accumulates the moments of the
//                               point-to-plane objective in registers, warp-shuffle reduction, cooperative small
//                               solve, PCL's convergence rule, state update.
//     Converged hypotheses retire at once: their tiles exit at the first instruction of later iterations.
//   * K5 is one flat launch (thread per hypothesis x scene point) + a fixed-order reduction of the tile partials.
//   * No tensor cores: these are gathers and small reductions, not dense contractions.
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <utility>

#include "hop_common.cuh"
#include "lm_replay.cuh"

namespace {

constexpr int TILE = HOP_TILE_PTS;  // scene points per CTA of the flat kernels (clouds are padded to a multiple)

// ------------------------------------------------------------------------------------------------------------
// small dense algebra used by the per-iteration solve (uniform across the warp)
// ------------------------------------------------------------------------------------------------------------
// solve (H + lambda*diag(H)) x = -g for symmetric 6x6 H (full storage); returns false when not positive definite
__device__ __forceinline__ bool chol_solve6(const float *Hs, const float *g, float lambda, float *x) {
  float L[6][6], inv[6];  // inv[j] = 1 / L[j][j]  (MUFU.RSQ: no division or square root on the dependent chain)
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      float s = Hs[6 * i + j];
      if (i == j) s += lambda * Hs[6 * i + i];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      if (i == j) {
        if (!(s > 0.f)) return false;
        inv[i] = rsqrtf(s);
        L[i][i] = s * inv[i];
      } else {
        L[i][j] = s * inv[j];
      }
    }
  }
  float y[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    float s = -g[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= L[i][k] * y[k];
    y[i] = s * inv[i];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    float s = y[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) s -= L[k][i] * x[k];
    x[i] = s * inv[i];
  }
  return true;
}

// R = exp([w]x)
__device__ __forceinline__ void so3_exp(float wx, float wy, float wz, float *R) {
  float th2 = wx * wx + wy * wy + wz * wz;
  float a, b;
  if (th2 < 1e-8f) { a = 1.f - th2 * (1.f / 6.f); b = 0.5f - th2 * (1.f / 24.f); }
  else { float th = sqrtf(th2); float s, c; sincosf(th, &s, &c); a = s / th; b = (1.f - c) / th2; }
  R[0] = 1.f - b * (wy * wy + wz * wz); R[1] = -a * wz + b * wx * wy;        R[2] = a * wy + b * wx * wz;
  R[3] = a * wz + b * wx * wy;         R[4] = 1.f - b * (wx * wx + wz * wz); R[5] = -a * wx + b * wy * wz;
  R[6] = -a * wy + b * wx * wz;        R[7] = a * wx + b * wy * wz;         R[8] = 1.f - b * (wx * wx + wy * wy);
}

// ------------------------------------------------------------------------------------------------------------
// accumulation policies for the point-to-plane step
//   residual of a correspondence under an increment (dR, dt) applied to the already-moved point p:
//       r = n . (dR p + dt - m) = u . [vec(dR - I); dt] + c ,   u = [n (x) p ; n] (12),  c = n . (p - m)
//   SOLVER 0 ("exact"): accumulate the 13x13 moment matrix of a = [u; c]  -> the full nonlinear objective
//       f(dR,dt) = y^T A y is then known in closed form and is minimised to convergence (what PCL's LM does with
//       its 400-evaluation budget) without revisiting the points.
//   SOLVER 1 ("gn"): accumulate J^T J (21) and J^T c (6), J = [p x n ; n]: one Gauss-Newton step per iteration.
// ------------------------------------------------------------------------------------------------------------
template <int SOLVER> struct Acc;

template <> struct Acc<0> {
  static constexpr int NA = 91;          // upper triangle of 13x13
  static constexpr int NACC = NA + 2;    // + sum d^2, count
  float a[NACC];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int k = 0; k < NACC; ++k) a[k] = 0.f;
  }
  __device__ __forceinline__ void add(float3 p, float3 n, float c, float d2) {
    float v[13];
    v[0] = n.x * p.x; v[1] = n.x * p.y; v[2] = n.x * p.z;
    v[3] = n.y * p.x; v[4] = n.y * p.y; v[5] = n.y * p.z;
    v[6] = n.z * p.x; v[7] = n.z * p.y; v[8] = n.z * p.z;
    v[9] = n.x; v[10] = n.y; v[11] = n.z;
    v[12] = c;
    int k = 0;
#pragma unroll
    for (int i = 0; i < 13; ++i)
#pragma unroll
      for (int j = i; j < 13; ++j) { a[k] = fmaf(v[i], v[j], a[k]); ++k; }
    a[NA] += d2;
    a[NA + 1] += 1.f;
  }
};

template <> struct Acc<1> {
  static constexpr int NA = 28;          // 21 (J^T J upper) + 6 (J^T c) + 1 (c^2)
  static constexpr int NACC = NA + 2;
  float a[NACC];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int k = 0; k < NACC; ++k) a[k] = 0.f;
  }
  __device__ __forceinline__ void add(float3 p, float3 n, float c, float d2) {
    float J[6];
    J[0] = p.y * n.z - p.z * n.y; J[1] = p.z * n.x - p.x * n.z; J[2] = p.x * n.y - p.y * n.x;
    J[3] = n.x; J[4] = n.y; J[5] = n.z;
    int k = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = i; j < 6; ++j) { a[k] = fmaf(J[i], J[j], a[k]); ++k; }
#pragma unroll
    for (int i = 0; i < 6; ++i) a[21 + i] = fmaf(J[i], c, a[21 + i]);
    a[27] = fmaf(c, c, a[27]);
    a[NA] += d2;
    a[NA + 1] += 1.f;
  }
};

constexpr int NACC_PAD = 96;
constexpr int WORK = 96;   // per warp: team totals

__device__ __forceinline__ int tri13(int i, int j) {  // index of (i,j), i<=j, in the row-major upper triangle
  return i * 13 - (i * (i - 1)) / 2 + (j - i);
}

// Exact minimiser of f(dR,dt) = y^T A y, y = [vec(dR - I); dt; 1], by damped Gauss-Newton on SE(3) from the
// identity, stopping like MINPACK's lmder does under PCL (relative reduction of the sum of squares <= sqrt(eps)).
// sums: the 91 reduced moments (shared memory).  Register resident: lane i < 13 owns row i of A, (R,t) and every
// small matrix are replicated in all lanes, rows meet through warp shuffles (no shared-memory round trips, no
// local memory).  All lanes return the same (R,t).
__device__ __forceinline__ void solve_exact(const float *sums, float *W, int lane, float *R, float *t) {
  (void)W;
  const unsigned FULL = 0xffffffffu;
  float Arow[13];
  {
    const int li = lane < 13 ? lane : 12;
#pragma unroll
    for (int j = 0; j < 13; ++j) {
      const int lo = li < j ? li : j, hi = li < j ? j : li;
      const float v = sums[tri13(lo, hi)];
      Arow[j] = lane < 13 ? v : 0.f;
    }
  }
  R[0] = 1.f; R[1] = 0.f; R[2] = 0.f; R[3] = 0.f; R[4] = 1.f; R[5] = 0.f; R[6] = 0.f; R[7] = 0.f; R[8] = 1.f;
  t[0] = t[1] = t[2] = 0.f;
  // H_tt = A[9+c][9+d] never changes
  float Htt[6];
  Htt[0] = __shfl_sync(FULL, Arow[9], 9);  Htt[1] = __shfl_sync(FULL, Arow[10], 9); Htt[2] = __shfl_sync(FULL, Arow[11], 9);
  Htt[3] = __shfl_sync(FULL, Arow[10], 10); Htt[4] = __shfl_sync(FULL, Arow[11], 10); Htt[5] = __shfl_sync(FULL, Arow[11], 11);
  // gy = A y; at the identity y = e_12, so gy is the last column of A and f = A[12][12]
  float gy[13];
#pragma unroll
  for (int j = 0; j < 13; ++j) gy[j] = __shfl_sync(FULL, Arow[12], j);
  float f = gy[12];
  float lambda = 0.f;
  const float ftol = 3.4526698e-4f;  // sqrt(FLT_EPSILON)
  int rejects = 0;
  for (int inner = 0; inner < 12; ++inner) {
    if (!(f > 0.f)) break;
    // Jacobian of y w.r.t. (w, tau): d vec(R)/dw_k = vec([e_k]x R), d t/d tau = I.   B = A J, this lane's row:
    float B[6];
    B[0] = B[1] = B[2] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      B[0] += Arow[6 + c] * R[3 + c] - Arow[3 + c] * R[6 + c];   // [e_0]x R : row1 = -R row2, row2 = R row1
      B[1] += Arow[0 + c] * R[6 + c] - Arow[6 + c] * R[0 + c];   // [e_1]x R : row0 = R row2, row2 = -R row0
      B[2] += Arow[3 + c] * R[0 + c] - Arow[0 + c] * R[3 + c];   // [e_2]x R : row0 = -R row1, row1 = R row0
    }
    B[3] = Arow[9]; B[4] = Arow[10]; B[5] = Arow[11];
    // H = J^T B (rows 0..11 of J), g = J^T gy
    float Hl[36], gl[6], dx[6];
    {
      float Bw[9][3];  // rows 0..8 of the first three columns of B, gathered from their lanes
#pragma unroll
      for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int l = 0; l < 3; ++l) Bw[i][l] = __shfl_sync(FULL, B[l], i);
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        float h0 = 0.f, h1 = 0.f, h2 = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          h0 += Bw[6 + c][l] * R[3 + c] - Bw[3 + c][l] * R[6 + c];
          h1 += Bw[0 + c][l] * R[6 + c] - Bw[6 + c][l] * R[0 + c];
          h2 += Bw[3 + c][l] * R[0 + c] - Bw[0 + c][l] * R[3 + c];
        }
        Hl[0 * 6 + l] = h0; Hl[1 * 6 + l] = h1; Hl[2 * 6 + l] = h2;
      }
      // H_wt[k][c] = row (9+c) of B, column k (A is symmetric)
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int k = 0; k < 3; ++k) { const float v = __shfl_sync(FULL, B[k], 9 + c); Hl[k * 6 + 3 + c] = v; Hl[(3 + c) * 6 + k] = v; }
      Hl[3 * 6 + 3] = Htt[0]; Hl[3 * 6 + 4] = Htt[1]; Hl[3 * 6 + 5] = Htt[2];
      Hl[4 * 6 + 3] = Htt[1]; Hl[4 * 6 + 4] = Htt[3]; Hl[4 * 6 + 5] = Htt[4];
      Hl[5 * 6 + 3] = Htt[2]; Hl[5 * 6 + 4] = Htt[4]; Hl[5 * 6 + 5] = Htt[5];
      gl[0] = gl[1] = gl[2] = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        gl[0] += gy[6 + c] * R[3 + c] - gy[3 + c] * R[6 + c];
        gl[1] += gy[0 + c] * R[6 + c] - gy[6 + c] * R[0 + c];
        gl[2] += gy[3 + c] * R[0 + c] - gy[0 + c] * R[3 + c];
      }
      gl[3] = gy[9]; gl[4] = gy[10]; gl[5] = gy[11];
    }
    bool ok = chol_solve6(Hl, gl, lambda, dx);
    while (!ok && rejects < 8) {  // rank deficient (e.g. a plane): regularise
      lambda = fmaxf(lambda * 10.f, 1e-6f);
      ++rejects;
      ok = chol_solve6(Hl, gl, lambda, dx);
    }
    if (!ok) break;
    // candidate
    float dR[9], Rn[9], tn[3];
    so3_exp(dx[0], dx[1], dx[2], dR);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) Rn[3 * i + j] = dR[3 * i] * R[j] + dR[3 * i + 1] * R[3 + j] + dR[3 * i + 2] * R[6 + j];
    tn[0] = t[0] + dx[3]; tn[1] = t[1] + dx[4]; tn[2] = t[2] + dx[5];
    float yn[13];
#pragma unroll
    for (int e = 0; e < 9; ++e) yn[e] = Rn[e] - ((e % 4 == 0) ? 1.f : 0.f);
    yn[9] = tn[0]; yn[10] = tn[1]; yn[11] = tn[2]; yn[12] = 1.f;
    float gi = 0.f;
#pragma unroll
    for (int m = 0; m < 13; ++m) gi = fmaf(Arow[m], yn[m], gi);
    float gn[13];
    float fn = 0.f;
#pragma unroll
    for (int j = 0; j < 13; ++j) { gn[j] = __shfl_sync(FULL, gi, j); fn = fmaf(gn[j], yn[j], fn); }
    const float rel = (f - fn) / f;
    const bool accepted = fn < f;
    if (accepted) {
#pragma unroll
      for (int j = 0; j < 13; ++j) gy[j] = gn[j];
#pragma unroll
      for (int e = 0; e < 9; ++e) R[e] = Rn[e];
      t[0] = tn[0]; t[1] = tn[1]; t[2] = tn[2];
      f = fn;
      lambda *= 0.1f;
      if (lambda < 1e-7f) lambda = 0.f;
    }
    // lmder's test: |actual reduction| <= ftol (a step that no longer changes the sum of squares, up or down, ends it)
    if (fabsf(rel) <= ftol) break;
    if (!accepted) {
      if (++rejects > 8) break;
      lambda = fmaxf(lambda * 10.f, 1e-4f);
    }
  }
}

// One Gauss-Newton step from the 28 reduced sums
__device__ void solve_gn(const float *sums, float *R, float *t) {
  float Hl[36], gl[6], dx[6];
  int k = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = i; j < 6; ++j) { float v = sums[k++]; Hl[6 * i + j] = v; Hl[6 * j + i] = v; }
#pragma unroll
  for (int i = 0; i < 6; ++i) gl[i] = sums[21 + i];
  float lambda = 0.f;
  bool ok = chol_solve6(Hl, gl, lambda, dx);
  for (int r = 0; !ok && r < 8; ++r) { lambda = fmaxf(lambda * 10.f, 1e-6f); ok = chol_solve6(Hl, gl, lambda, dx); }
  if (!ok) { dx[0] = dx[1] = dx[2] = dx[3] = dx[4] = dx[5] = 0.f; }
  so3_exp(dx[0], dx[1], dx[2], R);
  t[0] = dx[3]; t[1] = dx[4]; t[2] = dx[5];
}


// The reference's own solver (PCL's TransformationEstimationPointToPlane = float lmdif), replayed on the moments: see
// lm_replay.cuh.  Called by all lanes of one warp; returns the increment W(x) in all lanes.
__device__ __noinline__ void solve_lm_replay(const float *sums, int lane, float *R, float *t) {
  lmr::MomentsDev A{sums, lane};
  float x[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  lmr::lm_replay_solve(A, x, nullptr);
  float y[13];
  lmr::warp_y(x, y);
#pragma unroll
  for (int e = 0; e < 9; ++e) R[e] = (e % 4 == 0) ? 1.f + y[e] : y[e];   // exact: y = fl(R) - I
  t[0] = x[0]; t[1] = x[1]; t[2] = x[2];
}

// SOLVER: 0 = replay of the reference's LM (parity path, default), 1 = one Gauss-Newton step, 2 = exact minimiser
// Returns false when the reference's LM would run away along an unconstrained translation (lm_replay.cuh,
// translation_unconstrained): the caller ends the ICP of the hypothesis "not converged", pose unchanged.
template <int SOLVER>
__device__ __forceinline__ bool solve_increment(const float *sums, int lane, float *R, float *t) {
  if constexpr (SOLVER == 0) {
    if (lmr::translation_unconstrained(lmr::MomentsDev{sums, lane})) return false;
    solve_lm_replay(sums, lane, R, t);
  } else if constexpr (SOLVER == 1) solve_gn(sums, R, t);
  else solve_exact(sums, nullptr, lane, R, t);
  return true;
}

// ------------------------------------------------------------------------------------------------------------
// per-hypothesis ICP state (global memory, 128 bytes)
// ------------------------------------------------------------------------------------------------------------
struct __align__(16) IcpState {
  float X[12];       // scene -> model frame (row-major R | t), the iterate
  float inc[12];     // previous increment (PCL keeps transformation_ when LM early-returns)
  double prev_mse;
  int iters;
  int status;        // 0 active, 1 finished + converged, 2 finished, not converged
  int pad[2];
};
static_assert(sizeof(IcpState) == 128, "IcpState layout");

__device__ __forceinline__ Rigid state_load(const float *x) {
  Rigid T;
#pragma unroll
  for (int e = 0; e < 9; ++e) T.r[e] = __ldg(x + e);
  T.t[0] = __ldg(x + 9); T.t[1] = __ldg(x + 10); T.t[2] = __ldg(x + 11);
  return T;
}

// state of every hypothesis of the batch + the batch's first active list (all of them) and iteration counters
__global__ void icp_init_kernel(const float *__restrict__ poses, int H, IcpState *__restrict__ st, int *__restrict__ list0,
                                int *__restrict__ counters, int n_counters) {
  int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h < n_counters) counters[h] = h == 0 ? H : 0;
  if (h >= H) return;
  Rigid P = rigid_load_colmajor(poses + 16 * (size_t)h);
  Rigid X = rigid_inverse(P);
  IcpState s;
#pragma unroll
  for (int e = 0; e < 9; ++e) { s.X[e] = X.r[e]; s.inc[e] = (e % 4 == 0) ? 1.f : 0.f; }
#pragma unroll
  for (int e = 0; e < 3; ++e) { s.X[9 + e] = X.t[e]; s.inc[9 + e] = 0.f; }
  s.prev_mse = DBL_MAX; s.iters = 0; s.status = 0; s.pad[0] = s.pad[1] = 0;
  st[h] = s;
  list0[h] = h;
}

struct CorrArgs {
  CloudDev scene;
  const float4 *model_nv;
  NNGridDev grid;
  const IcpState *state;   // already offset to the batch
  const int *list;         // active hypotheses of this iteration (batch-local ids)
  const int *n_active;
  int n_tiles;
  float cos_thr;           // smallest float whose double value exceeds cos(angle)
  float max_d2;
  float4 *rec0, *rec1;     // [position in list][n_padded]: (p.xyz, n.(p-m)) and (n.xyz, d^2 | -1 when rejected)
};

// K4a: correspondence estimation + rejection.  CorrespondenceEstimation (exact 1-NN, d^2 <= max_dist^2) and
// CorrespondenceRejectorSurfaceNormal (rotated source normal . target normal > cos(angle)).
// Persistent CTAs stride over the (active hypothesis, scene tile) work items; one thread per scene point.
__global__ void __launch_bounds__(TILE) icp_correspond_kernel(CorrArgs a) {
  const int n_work = __ldg(a.n_active) * a.n_tiles;
  for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
    const int pos = w / a.n_tiles, tile = w - pos * a.n_tiles;
    const IcpState *st = a.state + __ldg(&a.list[pos]);
    const Rigid X = state_load(st->X);
    const int i = tile * TILE + threadIdx.x;
    const float4 sp = __ldg(&a.scene.pw[i]);
    const float3 p = rigid_apply(X, sp.x, sp.y, sp.z);
    float bd; float4 bp;
    const int j = nn_query(a.grid, p.x, p.y, p.z, bd, bp);
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = make_float4(0.f, 0.f, 0.f, -1.f);
    if (j >= 0 && bd <= a.max_d2) {
      const float4 sn = __ldg(&a.scene.nv[i]);
      const float4 mn = __ldg(&a.model_nv[j]);
      const float3 ns = rigid_rotate(X, sn.x, sn.y, sn.z);
      const float dot = ns.x * mn.x + ns.y * mn.y + ns.z * mn.z;
      if (dot >= a.cos_thr) {
        r0 = make_float4(p.x, p.y, p.z, mn.x * (p.x - bp.x) + mn.y * (p.y - bp.y) + mn.z * (p.z - bp.z));
        r1 = make_float4(mn.x, mn.y, mn.z, bd);
      }
    }
    const size_t o = (size_t)pos * a.scene.n_padded + i;
    a.rec0[o] = r0;
    a.rec1[o] = r1;
  }
}

struct SolveArgs {
  const float4 *rec0, *rec1;
  int n_padded;
  IcpState *state;     // offset to the batch
  float *poses;        // offset to the batch: Hb x 16, written when a hypothesis finishes converged
  int32_t *iters_out, *conv_out;  // offset to the batch (may be null)
  const int *list;       // active hypotheses of this iteration
  const int *n_active;
  int *next_list;        // survivors are appended here (order is irrelevant to the results)
  int *next_count;
  int max_iter;
  double abs_mse_eps;
};

// K4b: TransformationEstimationPointToPlane (LM) + DefaultConvergenceCriteria for one hypothesis per warp team.
template <int NW, int TEAM, int SOLVER>
__global__ void __launch_bounds__(NW * 32, 1) icp_solve_kernel(SolveArgs a) {
  constexpr int NT = NW / TEAM;
  using AccT = Acc<SOLVER == 1 ? 1 : 0>;
  __shared__ __align__(16) float s_part[NW][NACC_PAD];
  __shared__ __align__(16) float s_work[NW][WORK];
  __shared__ float s_red[NW][32 * 33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int team = warp / TEAM, tw = warp % TEAM;
  const int pos = blockIdx.x * NT + team;
  const bool active = pos < __ldg(a.n_active);
  const int h = active ? __ldg(&a.list[pos]) : 0;
  IcpState *st = a.state + h;

  AccT acc;
  acc.clear();
  if (active) {
    const float4 *r0 = a.rec0 + (size_t)pos * a.n_padded;
    const float4 *r1 = a.rec1 + (size_t)pos * a.n_padded;
    constexpr int STEP = 32 * TEAM, U = 4;  // n_padded is a multiple of 256 = 8 * 32: whole batches for TEAM <= 2
    for (int i0 = tw * 32 + lane; i0 < a.n_padded; i0 += STEP * U) {
      float4 q0[U], q1[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {  // all loads of the batch in flight before the first use
        const int i = i0 + u * STEP;
        const bool in = i < a.n_padded;
        q1[u] = in ? __ldcs(&r1[i]) : make_float4(0.f, 0.f, 0.f, -1.f);
        q0[u] = in ? __ldcs(&r0[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (q1[u].w >= 0.f) acc.add(make_float3(q0[u].x, q0[u].y, q0[u].z), make_float3(q1[u].x, q1[u].y, q1[u].z), q0[u].w, q1[u].w);
    }
  }
  // ---- reduce inside the warp through a padded shared-memory transpose (32 accumulators at a time): lane l ends
  //      with the totals of accumulators l, 32 + l, 64 + l; then across the team ----
  float *my_part = s_part[warp];
  {
    float *buf = s_red[warp];
#pragma unroll
    for (int c = 0; c < (AccT::NACC + 31) / 32; ++c) {
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (32 * c + k < AccT::NACC) buf[k * 33 + lane] = acc.a[32 * c + k];
      __syncwarp();
      if (32 * c + lane < AccT::NACC) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) s += buf[lane * 33 + j];
        my_part[32 * c + lane] = s;
      }
      __syncwarp();
    }
  }
  float *W = s_work[warp];
  const float *sums = my_part;
  if (TEAM > 1) {
    __syncthreads();
    if (tw != 0) return;  // the team's first warp sums the partial rows in a fixed order and solves
    float *tot = W;
    for (int k = lane; k < AccT::NACC; k += 32) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < TEAM; ++q) s += s_part[team * TEAM + q][k];
      tot[k] = s;
    }
    sums = tot;
  }
  __syncwarp();
  if (!active) return;

  const float cnt_f = sums[AccT::NA + 1];
  const float sumd2 = sums[AccT::NA];
  const int cnt = (int)(cnt_f + 0.5f);
  int iters = st->iters;
  bool converged = false, finished = false;
  Rigid X = state_load(st->X);
  bool runaway = false;
  Rigid inc;
  if (cnt >= 6) runaway = !solve_increment<SOLVER>(sums, lane, inc.r, inc.t);
  if (cnt < 3) {
    finished = true;  // "Not enough correspondences": hasConverged() false -> identity -> pose unchanged
  } else if (runaway) {
    // the reference's LM slides the scene metres off the model here; its next iteration finds no correspondences
    ++iters;
    finished = true;
    if (tw == 0 && lane == 0) st->iters = iters;
  } else {
    if (cnt >= 6) {
    } else if (cnt >= 4) {
      // Eigen LM: m < n -> ImproperInputParameters, x stays 0 -> identity increment
#pragma unroll
      for (int e = 0; e < 9; ++e) inc.r[e] = (e % 4 == 0) ? 1.f : 0.f;
      inc.t[0] = inc.t[1] = inc.t[2] = 0.f;
    } else {
      inc = state_load(st->inc);  // PCL's LM returns early with < 4 correspondences, transformation_ keeps its old value
    }
    X = rigid_compose(inc, X);
    ++iters;
    double mse = st->prev_mse;
    if (iters >= a.max_iter) converged = true;
    else {
      double cos_angle = 0.5 * ((double)inc.r[0] + (double)inc.r[4] + (double)inc.r[8] - 1.0);
      double tsq = (double)inc.t[0] * inc.t[0] + (double)inc.t[1] * inc.t[1] + (double)inc.t[2] * inc.t[2];
      if (cos_angle >= 1.0 && tsq <= 0.0) converged = true;
      else {
        mse = (double)sumd2 / (double)cnt;
        if (fabs(mse - st->prev_mse) < a.abs_mse_eps) converged = true;
      }
    }
    finished = converged;
    if (tw == 0 && lane == 0) {
#pragma unroll
      for (int e = 0; e < 9; ++e) { st->X[e] = X.r[e]; st->inc[e] = inc.r[e]; }
#pragma unroll
      for (int e = 0; e < 3; ++e) { st->X[9 + e] = X.t[e]; st->inc[9 + e] = inc.t[e]; }
      st->prev_mse = mse;
      st->iters = iters;
    }
  }
  if (!finished && tw == 0 && lane == 0) a.next_list[atomicAdd(a.next_count, 1)] = h;
  if (finished && tw == 0 && lane == 0) {
    st->status = converged ? 1 : 2;
    if (converged) {
      Rigid P = rigid_inverse(X);
      rigid_store_colmajor(P, a.poses + 16 * (size_t)h);
    }
    if (a.iters_out) a.iters_out[h] = iters;
    if (a.conv_out) a.conv_out[h] = converged ? 1 : 0;
  }
}


// ------------------------------------------------------------------------------------------------------------
// K4 fused: the whole ICP of a hypothesis inside one CTA, no correspondence records in global memory.
//   Persistent CTAs pull hypotheses from a queue.  Per ICP iteration and per chunk of FUSED_CHUNK scene points:
//     phase A  every thread finds the correspondences of its points (same code path as icp_correspond_kernel) and leaves
//              the 32-byte records in SHARED memory;
//     phase B  the 8 warps split the 13x13 moment matrix four ways (24 of its 91 + 2 entries each) and the chunk two
//              ways, so a thread carries 24 accumulators across all chunks instead of 93 -- registers stay low enough
//              for three CTAs per SM, which is what hides the gather latency of phase A;
//   then one shuffle reduction, the small solve + PCL's convergence rule on warp 0, and the new transform goes back to all
//   threads through shared memory.  A hypothesis that converges frees its CTA for the next one at once; there is no
//   per-iteration launch, no inter-CTA dependency and nothing but the pose is written to HBM.
// ------------------------------------------------------------------------------------------------------------
constexpr int SLICE = 24;

__host__ __device__ constexpr int tri_row(int k) { int i = 0; while (k >= 13 - i) { k -= 13 - i; ++i; } return i; }
__host__ __device__ constexpr int tri_col(int k) { int i = 0; while (k >= 13 - i) { k -= 13 - i; ++i; } return i + k; }

template <int S, int E>
__device__ __forceinline__ void acc_one(float (&acc)[SLICE], const float (&v)[13], float d2) {
  constexpr int k = SLICE * S + E;
  if constexpr (k < 91) acc[E] = fmaf(v[tri_row(k)], v[tri_col(k)], acc[E]);
  else if constexpr (k == 91) acc[E] += d2;
  else if constexpr (k == 92) acc[E] += 1.f;
}
template <int S, int... E>
__device__ __forceinline__ void acc_slice(float (&acc)[SLICE], const float (&v)[13], float d2, std::integer_sequence<int, E...>) {
  (acc_one<S, E>(acc, v, d2), ...);
}

template <int S>
__device__ __forceinline__ void accumulate_chunk(float (&acc)[SLICE], const float4 *rec0, const float4 *rec1, int begin, int end, int lane) {
  for (int i = begin + lane; i < end; i += 32) {
    const float4 q1 = rec1[i];
    if (!(q1.w >= 0.f)) continue;
    const float4 q0 = rec0[i];
    float v[13];
    v[0] = q1.x * q0.x; v[1] = q1.x * q0.y; v[2] = q1.x * q0.z;
    v[3] = q1.y * q0.x; v[4] = q1.y * q0.y; v[5] = q1.y * q0.z;
    v[6] = q1.z * q0.x; v[7] = q1.z * q0.y; v[8] = q1.z * q0.z;
    v[9] = q1.x; v[10] = q1.y; v[11] = q1.z;
    v[12] = q0.w;
    acc_slice<S>(acc, v, q1.w, std::make_integer_sequence<int, SLICE>());
  }
}

struct FusedArgs {
  CloudDev scene;
  const float4 *model_nv;
  NNGridDev grid;
  float *poses;
  int H;
  int32_t *iters_out, *conv_out;
  int *counter;
  float cos_thr, max_d2;
  int max_iter;
  double abs_mse_eps;
  long long *prof;   // null, or 6 cycle counters (HOP_FUSED_PROFILE=1)
};

// THREADS per CTA (a multiple of 128: four moment slices x THREADS/128 parts of the chunk), CHUNK scene points whose
// records sit in shared memory at a time, PROF = with the cycle accounting, MINB resident CTAs per SM.
template <int THREADS, int CHUNK, bool PROF, int MINB, int SOLVER>
__global__ void __launch_bounds__(THREADS, MINB) icp_fused_kernel(FusedArgs a) {
  constexpr int PARTS = THREADS / 128;
  extern __shared__ __align__(16) float4 fused_smem[];
  float4 *rec0 = fused_smem, *rec1 = fused_smem + CHUNK;
  __shared__ __align__(16) float s_sums[PARTS][96];
  __shared__ __align__(16) float s_tot[96];
  __shared__ float s_X[12];
  __shared__ int s_ctl[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slice = warp & 3, part = warp >> 2;
  // optional cycle accounting (thread 0 of every CTA): [0] phase A  [1] wait at the barrier after A  [2] phase B + barrier
  // [3] reduction + solve + broadcast  [4] passes  [5] whole CTA life time
  long long t_acc[6] = {0, 0, 0, 0, 0, 0}, t_mark = 0;
  const long long t_start = PROF ? clock64() : 0;
  for (;;) {
    if (tid == 0) s_ctl[0] = atomicAdd(a.counter, 1);
    __syncthreads();
    const int h = s_ctl[0];
    if (h >= a.H) break;
    Rigid X = rigid_inverse(rigid_load_colmajor(a.poses + 16 * (size_t)h));
    // convergence state (meaningful on warp 0; uniform across its lanes)
    Rigid inc_prev;
#pragma unroll
    for (int e = 0; e < 9; ++e) inc_prev.r[e] = (e % 4 == 0) ? 1.f : 0.f;
    inc_prev.t[0] = inc_prev.t[1] = inc_prev.t[2] = 0.f;
    double prev_mse = DBL_MAX;
    int iters = 0;
    bool finished = false, converged = false;
    while (!finished) {
      float acc[SLICE];
#pragma unroll
      for (int e = 0; e < SLICE; ++e) acc[e] = 0.f;
      for (int c0 = 0; c0 < a.scene.n_padded; c0 += CHUNK) {
        const int cnt = min(CHUNK, a.scene.n_padded - c0);
        // ---- phase A: correspondences of this chunk -> shared memory ----
        if (PROF && tid == 0) t_mark = clock64();
        for (int i = tid; i < cnt; i += THREADS) {
          const float4 sp = __ldg(&a.scene.pw[c0 + i]);
          const float3 p = rigid_apply(X, sp.x, sp.y, sp.z);
          float bd; float4 bp;
          const int j = nn_query(a.grid, p.x, p.y, p.z, bd, bp);
          float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = make_float4(0.f, 0.f, 0.f, -1.f);
          if (j >= 0 && bd <= a.max_d2) {
            const float4 sn = __ldg(&a.scene.nv[c0 + i]);
            const float4 mn = __ldg(&a.model_nv[j]);
            const float3 ns = rigid_rotate(X, sn.x, sn.y, sn.z);
            const float dot = ns.x * mn.x + ns.y * mn.y + ns.z * mn.z;
            if (dot >= a.cos_thr) {
              r0 = make_float4(p.x, p.y, p.z, mn.x * (p.x - bp.x) + mn.y * (p.y - bp.y) + mn.z * (p.z - bp.z));
              r1 = make_float4(mn.x, mn.y, mn.z, bd);
            }
          }
          rec0[i] = r0;
          rec1[i] = r1;
        }
        if (PROF && tid == 0) { const long long t = clock64(); t_acc[0] += t - t_mark; t_mark = t; }
        __syncthreads();
        if (PROF && tid == 0) { const long long t = clock64(); t_acc[1] += t - t_mark; t_mark = t; }
        // ---- phase B: this warp's slice of the moments over its part of the chunk ----
        const int per = (cnt + PARTS - 1) / PARTS;
        const int hb = min(part * per, cnt), he = min(hb + per, cnt);
        switch (slice) {
          case 0: accumulate_chunk<0>(acc, rec0, rec1, hb, he, lane); break;
          case 1: accumulate_chunk<1>(acc, rec0, rec1, hb, he, lane); break;
          case 2: accumulate_chunk<2>(acc, rec0, rec1, hb, he, lane); break;
          default: accumulate_chunk<3>(acc, rec0, rec1, hb, he, lane); break;
        }
        __syncthreads();
        if (PROF && tid == 0) { const long long t = clock64(); t_acc[2] += t - t_mark; t_mark = t; }
      }
      // ---- reduce: lanes -> warp totals -> the parts of the chunk ----
      float mine = 0.f;
#pragma unroll
      for (int e = 0; e < SLICE; ++e) {
        const float tot = warp_sum(acc[e]);
        if (lane == e) mine = tot;
      }
      if (lane < SLICE) s_sums[part][SLICE * slice + lane] = mine;
      __syncthreads();
      if (warp == 0) {
        for (int k = lane; k < 96; k += 32) {
          float s = s_sums[0][k];
#pragma unroll
          for (int q = 1; q < PARTS; ++q) s += s_sums[q][k];
          s_tot[k] = s;
        }
        __syncwarp();
        const float *sums = s_tot;
        const float cnt_f = sums[92], sumd2 = sums[91];
        const int cnt = (int)(cnt_f + 0.5f);
        bool runaway = false;
        Rigid inc;
        if (cnt >= 6) runaway = !solve_increment<SOLVER>(sums, lane, inc.r, inc.t);
        if (cnt < 3) {
          finished = true;  // "Not enough correspondences": hasConverged() false -> pose unchanged
        } else if (runaway) {
          // the reference's LM slides the scene metres off the model here; its next iteration finds no correspondences
          ++iters;
          finished = true;
        } else {
          if (cnt >= 6) {}
          else if (cnt >= 4) {  // Eigen LM: m < n -> ImproperInputParameters, x stays 0 -> identity increment
#pragma unroll
            for (int e = 0; e < 9; ++e) inc.r[e] = (e % 4 == 0) ? 1.f : 0.f;
            inc.t[0] = inc.t[1] = inc.t[2] = 0.f;
          } else inc = inc_prev;  // PCL's LM returns early with < 4 correspondences, transformation_ keeps its old value
          X = rigid_compose(inc, X);
          inc_prev = inc;
          ++iters;
          if (iters >= a.max_iter) converged = true;
          else {
            const double cos_angle = 0.5 * ((double)inc.r[0] + (double)inc.r[4] + (double)inc.r[8] - 1.0);
            const double tsq = (double)inc.t[0] * inc.t[0] + (double)inc.t[1] * inc.t[1] + (double)inc.t[2] * inc.t[2];
            if (cos_angle >= 1.0 && tsq <= 0.0) converged = true;
            else {
              const double mse = (double)sumd2 / (double)cnt;
              if (fabs(mse - prev_mse) < a.abs_mse_eps) converged = true;
              prev_mse = mse;
            }
          }
          finished = converged;
        }
        if (lane == 0) {
#pragma unroll
          for (int e = 0; e < 9; ++e) s_X[e] = X.r[e];
          s_X[9] = X.t[0]; s_X[10] = X.t[1]; s_X[11] = X.t[2];
          s_ctl[1] = (finished ? 1 : 0) | (converged ? 2 : 0);
        }
      }
      __syncthreads();
      if (PROF && tid == 0) { const long long t = clock64(); t_acc[3] += t - t_mark; t_mark = t; t_acc[4] += 1; }
#pragma unroll
      for (int e = 0; e < 9; ++e) X.r[e] = s_X[e];
      X.t[0] = s_X[9]; X.t[1] = s_X[10]; X.t[2] = s_X[11];
      finished = (s_ctl[1] & 1) != 0;
      converged = (s_ctl[1] & 2) != 0;
    }
    if (tid == 0) {
      if (converged) { const Rigid P = rigid_inverse(X); rigid_store_colmajor(P, a.poses + 16 * (size_t)h); }
      if (a.iters_out) a.iters_out[h] = iters;
      if (a.conv_out) a.conv_out[h] = converged ? 1 : 0;
    }
    __syncthreads();  // s_ctl / s_X are reused by the next hypothesis
  }
  if (PROF && tid == 0) {
    t_acc[5] = clock64() - t_start;
#pragma unroll
    for (int k = 0; k < 6; ++k) atomicAdd((unsigned long long *)a.prof + k, (unsigned long long)t_acc[k]);
  }
}

template <int THREADS, int CHUNK, bool PROF, int MINB, int SOLVER>
static cudaError_t launch_fused(hop_ctx *ctx, const FusedArgs &f, int H) {
  const size_t smem = 2 * (size_t)CHUNK * sizeof(float4);
  cudaError_t e = ctx->func_smem_optin(icp_fused_kernel<THREADS, CHUNK, PROF, MINB, SOLVER>, smem);
  if (e != cudaSuccess) return e;
  const int grid_ctas = (int)std::min<long>((long)H, (long)ctx->sm_count * MINB);
  icp_fused_kernel<THREADS, CHUNK, PROF, MINB, SOLVER><<<grid_ctas, THREADS, smem, ctx->stream>>>(f);
  return cudaGetLastError();
}

// CTA shape by batch size.  Small batches: 256-thread CTAs (a hypothesis finishes sooner, shorter tail); large batches: 128-thread
// CTAs (the serial small solve idles 3 warps instead of 7), 8 of them per SM at 64 registers -- the few spilled bytes cost less
// than the extra warps hide (15.39 -> 14.78 ms at the headline size; small batches lose with it).  Record chunk = 4 points per
// thread: the rest of the 256 KB stays L1.
template <int SOLVER>
static cudaError_t launch_fused_solver(hop_ctx *ctx, const FusedArgs &f, int H, bool small_ctas, bool prof_on) {
  if (prof_on) return small_ctas ? launch_fused<128, 512, true, 6, SOLVER>(ctx, f, H) : launch_fused<256, 1024, true, 3, SOLVER>(ctx, f, H);
  return small_ctas ? launch_fused<128, 512, false, 8, SOLVER>(ctx, f, H) : launch_fused<256, 1024, false, 3, SOLVER>(ctx, f, H);
}

// ------------------------------------------------------------------------------------------------------------
// K5
// ------------------------------------------------------------------------------------------------------------
struct LcpArgs {
  CloudDev scene;               // iteration order (Morton)
  const float4 *scene_nv_idx;   // normals addressed by the scene grid's (original) point index
  const float4 *model_nv;
  NNGridDev mgrid;   // model grid (radius >= dist)
  NNGridDev sgrid;   // scene grid (radius >= dist), for the reciprocal term
  const float *poses;
  int H;
  float dist, inv_dist, dist2, cos_thr;
  int use_normal, use_dot, use_recip, use_weights;
  int n_tiles;
  float *partial;    // [H][splits]
  float *scores;
};

// CTA = (hypothesis, split): one thread per scene point of every `splits`-th tile; the pose algebra is done once per
// CTA; the CTA's sum goes to partial[h][split]
template <int MINB>
__global__ void __launch_bounds__(TILE, MINB) lcp_score_kernel(LcpArgs a) {
  __shared__ float s_w[TILE / 32];
  const int h = blockIdx.y;
  const Rigid T = rigid_load_colmajor(a.poses + 16 * (size_t)h);
  const Rigid Ti = rigid_inverse(T);
  float score = 0.f;
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int i = tile * TILE + threadIdx.x;
    const float4 sp = __ldg(&a.scene.pw[i]);
    const float3 p = rigid_apply(Ti, sp.x, sp.y, sp.z);   // scene point in the model frame
    float bd; float4 bp;
    const int j = nn_query(a.mgrid, p.x, p.y, p.z, bd, bp);
    if (j >= 0 && bd < a.dist2) {                         // Utils.cpp:388 (strict)
      const float w = a.use_weights ? sp.w : 1.f;
      const float4 mn = __ldg(&a.model_nv[j]);
      // transformed model normal, normalised (rotation keeps the norm: use the stored 1/|n|)
      const float3 mr = rigid_rotate(T, mn.x * mn.w, mn.y * mn.w, mn.z * mn.w);
      if (!a.use_normal) score += w;
      else {
        const float4 sn = __ldg(&a.scene.nv[i]);
        const float dot = (sn.x * mr.x + sn.y * mr.y + sn.z * mr.z) * sn.w;
        if (dot > a.cos_thr) score += a.use_dot ? dot * (1.f - sqrtf(bd) * a.inv_dist) * w : w;
      }
      if (a.use_recip) {
        // nearest scene point of the (transformed) model neighbour; it lies within `dist` because scene point i
        // itself does, so the radius-limited scene grid is exact here
        const float3 q = rigid_apply(T, bp.x, bp.y, bp.z);
        float ed; float4 ep;
        const int k = nn_query(a.sgrid, q.x, q.y, q.z, ed, ep);
        if (k >= 0) {
          if (!a.use_normal) score += w;
          else {
            const float4 s2 = __ldg(&a.scene_nv_idx[k]);
            const float dot = (s2.x * mr.x + s2.y * mr.y + s2.z * mr.z) * s2.w;
            if (dot > a.cos_thr) score += a.use_dot ? dot * (1.f - sqrtf(ed) * a.inv_dist) * w : w;
          }
        }
      }
    }
  }
  score = warp_sum(score);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = score;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < TILE / 32; ++w) s += s_w[w];
    a.partial[(size_t)h * gridDim.x + blockIdx.x] = s;
  }
}

// fixed-order sum of the tile partials (deterministic bits run to run)
__global__ void lcp_reduce_kernel(const float *__restrict__ partial, int n_tiles, int H, float *__restrict__ scores) {
  int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  float s = 0.f;
  for (int t = 0; t < n_tiles; ++t) s += partial[(size_t)h * n_tiles + t];
  scores[h] = s;
}

// smallest float f with (double)f > thr  (PCL compares the float score against a double threshold with '>')
static float float_above(double thr) {
  float f = (float)thr;
  while ((double)f > thr) f = nextafterf(f, -INFINITY);
  while (!((double)f > thr)) f = nextafterf(f, INFINITY);
  return f;
}

static int pick_team(int H, int sm_count, int nw, int requested) {
  if (requested == 1 || requested == 2 || requested == 4 || requested == 8) return requested;
  // enough independent hypotheses to give every warp slot its own?  otherwise split hypotheses across warps
  const long slots = (long)sm_count * nw;
  int team = 1;
  while (team < 8 && (long)H * team < slots) team *= 2;
  return team;
}

constexpr int SOLVE_NW = 8;

template <int TEAM>
static void launch_solve(hop_ctx *ctx, const SolveArgs &s, int Hb, int solver) {
  constexpr int NT = SOLVE_NW / TEAM;
  const int grid = (Hb + NT - 1) / NT;
  if (solver == 1) icp_solve_kernel<SOLVE_NW, TEAM, 1><<<grid, SOLVE_NW * 32, 0, ctx->stream>>>(s);
  else if (solver == 2) icp_solve_kernel<SOLVE_NW, TEAM, 2><<<grid, SOLVE_NW * 32, 0, ctx->stream>>>(s);
  else icp_solve_kernel<SOLVE_NW, TEAM, 0><<<grid, SOLVE_NW * 32, 0, ctx->stream>>>(s);
}

}  // namespace

int hop_launch_icp(hop_ctx *ctx, const CloudDev &scene, const CloudDev &model, const NNGridDev &grid, float *d_poses, int H,
                   const hop_icp_params &p, int32_t *d_iters, int32_t *d_conv) {
  if (H <= 0) return HOP_OK;
  if (p.mode != 0) { ctx->err = "hop_icp_refine: mode 1 (point-to-point) not built yet"; return HOP_EINVAL; }
  const int max_iter = p.max_iter < 1 ? 1 : p.max_iter;
  if (p.solver < 0 || p.solver > 2) { ctx->err = "hop_icp_refine: solver must be 0 (reference LM), 1 (Gauss-Newton) or 2 (exact minimiser)"; return HOP_EINVAL; }
  if (p.solver != 1 && p.pipeline != 1) {
    // fused pipeline (default): one persistent launch for the whole batch
    FusedArgs f;
    f.scene = scene; f.model_nv = model.nv; f.grid = grid; f.poses = d_poses; f.H = H; f.iters_out = d_iters; f.conv_out = d_conv;
    f.counter = ctx->d_counter;
    f.cos_thr = float_above(cos((double)p.angle_deg / 180.0 * M_PI));
    f.max_d2 = p.max_dist * p.max_dist; f.max_iter = max_iter; f.abs_mse_eps = p.abs_mse_eps;
    HOP_CUDA(ctx, cudaMemsetAsync(ctx->d_counter, 0, sizeof(int), ctx->stream));
    const int variant = ctx->tune.fused_variant;
    const bool prof_on = ctx->tune.fused_profile;
    if (prof_on && !ctx->d_fused_prof) { HOP_CUDA(ctx, cudaMalloc(&ctx->d_fused_prof, 6 * sizeof(long long))); }
    if (prof_on) HOP_CUDA(ctx, cudaMemsetAsync(ctx->d_fused_prof, 0, 6 * sizeof(long long), ctx->stream));
    f.prof = prof_on ? ctx->d_fused_prof : nullptr;
    {
      ProfScope ps(ctx, HOP_PROF_ICP_FUSED);
      const bool small_ctas = variant == 2 || (variant == 0 && (long)H >= 24L * ctx->sm_count);
      HOP_CUDA(ctx, p.solver == 0 ? launch_fused_solver<0>(ctx, f, H, small_ctas, prof_on) : launch_fused_solver<2>(ctx, f, H, small_ctas, prof_on));
    }
    if (prof_on) {
      long long hp[6];
      HOP_CUDA(ctx, cudaMemcpyAsync(hp, ctx->d_fused_prof, sizeof(hp), cudaMemcpyDeviceToHost, ctx->stream));
      HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      const double tot = (double)std::max<long long>(hp[5], 1);
      fprintf(stderr, "[hop fused profile] H=%d passes=%lld  A %.1f%%  barrier %.1f%%  B %.1f%%  solve %.1f%%  other %.1f%%  (CTA cycles %.3g)\n", H, hp[4],
              100.0 * hp[0] / tot, 100.0 * hp[1] / tot, 100.0 * hp[2] / tot, 100.0 * hp[3] / tot, 100.0 * (tot - hp[0] - hp[1] - hp[2] - hp[3]) / tot, tot);
    }
    ctx->launches += 1;
    HOP_CUDA(ctx, cudaGetLastError());
    return HOP_OK;
  }
  const int n_tiles = scene.n_padded / TILE;
  // hypotheses per batch: bounded by the correspondence-record buffer (32 B per hypothesis x scene point)
  const size_t rec_per_h = (size_t)scene.n_padded * 32;
  const size_t rec_budget = (size_t)1 << 30;
  const int Hb_max = (int)std::min<size_t>((size_t)H, std::max<size_t>(1, rec_budget / rec_per_h));
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t state_bytes = up(sizeof(IcpState) * (size_t)Hb_max), list_bytes = up(sizeof(int) * (size_t)Hb_max);
  const size_t cnt_bytes = up(sizeof(int) * (size_t)(max_iter + 1));
  char *base = (char *)ctx->ensure_work(state_bytes + 2 * list_bytes + cnt_bytes + rec_per_h * Hb_max);
  if (!base) { ctx->err = "hop_icp_refine: work buffer allocation failed"; return HOP_ENOMEM; }
  IcpState *state = (IcpState *)base;
  int *lists[2] = {(int *)(base + state_bytes), (int *)(base + state_bytes + list_bytes)};
  int *counters = (int *)(base + state_bytes + 2 * list_bytes);
  float4 *rec0 = (float4 *)(base + state_bytes + 2 * list_bytes + cnt_bytes);
  float4 *rec1 = rec0 + (size_t)Hb_max * scene.n_padded;

  for (int h0 = 0; h0 < H; h0 += Hb_max) {
    const int Hb = std::min(Hb_max, H - h0);
    const int n_init = std::max(Hb, max_iter + 1);
    icp_init_kernel<<<(n_init + 127) / 128, 128, 0, ctx->stream>>>(d_poses + 16 * (size_t)h0, Hb, state, lists[0], counters, max_iter + 1);
    ctx->launches += 1;
    CorrArgs c;
    c.scene = scene; c.model_nv = model.nv; c.grid = grid; c.state = state; c.n_tiles = n_tiles;
    c.cos_thr = float_above(cos((double)p.angle_deg / 180.0 * M_PI));
    c.max_d2 = p.max_dist * p.max_dist;
    c.rec0 = rec0; c.rec1 = rec1;
    SolveArgs s;
    s.rec0 = rec0; s.rec1 = rec1; s.n_padded = scene.n_padded; s.state = state; s.poses = d_poses + 16 * (size_t)h0;
    s.iters_out = d_iters ? d_iters + h0 : nullptr; s.conv_out = d_conv ? d_conv + h0 : nullptr;
    s.max_iter = max_iter; s.abs_mse_eps = p.abs_mse_eps;
    const int team = pick_team(Hb, ctx->sm_count, SOLVE_NW, p.team_warps);
    const int corr_grid = (int)std::min<long>((long)n_tiles * Hb, (long)ctx->sm_count * 16);
    for (int it = 0; it < max_iter; ++it) {
      c.list = lists[it & 1]; c.n_active = counters + it;
      s.list = c.list; s.n_active = c.n_active; s.next_list = lists[(it + 1) & 1]; s.next_count = counters + it + 1;
      { ProfScope ps(ctx, HOP_PROF_ICP_CORRESPOND); icp_correspond_kernel<<<corr_grid, TILE, 0, ctx->stream>>>(c); }
      ProfScope ps(ctx, HOP_PROF_ICP_SOLVE);
      switch (team) {
        case 1: launch_solve<1>(ctx, s, Hb, p.solver); break;
        case 2: launch_solve<2>(ctx, s, Hb, p.solver); break;
        case 4: launch_solve<4>(ctx, s, Hb, p.solver); break;
        default: launch_solve<8>(ctx, s, Hb, p.solver); break;
      }
      ctx->launches += 2;
    }
  }
  HOP_CUDA(ctx, cudaGetLastError());
  return HOP_OK;
}

int hop_launch_lcp(hop_ctx *ctx, const CloudDev &scene, const float4 *scene_nv_by_index, const CloudDev &model, const NNGridDev &model_grid,
                   const NNGridDev &scene_grid, const float *d_poses, int H, const hop_lcp_params &p, int use_weights,
                   float *d_scores) {
  if (H <= 0) return HOP_OK;
  const int n_tiles = scene.n_padded / TILE;
  LcpArgs a;
  a.scene = scene; a.scene_nv_idx = scene_nv_by_index; a.model_nv = model.nv; a.mgrid = model_grid; a.sgrid = scene_grid; a.H = H;
  a.dist = p.dist; a.inv_dist = 1.f / p.dist; a.dist2 = p.dist * p.dist;
  a.cos_thr = (float)cos((double)p.angle_deg / 180.0 * M_PI);
  a.use_normal = p.use_normal; a.use_dot = p.use_dot_score; a.use_recip = p.use_reciprocal; a.use_weights = use_weights;
  const int Hb_max = 65535;
  // CTAs per hypothesis: enough to fill the machine when the batch is small, one when it is large
  int splits = (int)((8L * ctx->sm_count + H - 1) / H);
  splits = std::max(1, std::min(splits, n_tiles));
  a.n_tiles = n_tiles;
  float *partial = (float *)ctx->ensure_work(sizeof(float) * (size_t)std::min(H, Hb_max) * splits);
  if (!partial) { ctx->err = "hop_lcp_score: work buffer allocation failed"; return HOP_ENOMEM; }
  a.partial = partial;
  for (int h0 = 0; h0 < H; h0 += Hb_max) {
    const int Hb = std::min(Hb_max, H - h0);
    a.poses = d_poses + 16 * (size_t)h0; a.scores = d_scores + h0;
    ProfScope ps(ctx, HOP_PROF_LCP_SCORE);
    const int lcp_variant = ctx->tune.lcp_variant;
    if (lcp_variant == 1) lcp_score_kernel<5><<<dim3(splits, Hb), TILE, 0, ctx->stream>>>(a);
    else if (lcp_variant == 2) lcp_score_kernel<6><<<dim3(splits, Hb), TILE, 0, ctx->stream>>>(a);
    else if (lcp_variant == 4) lcp_score_kernel<4><<<dim3(splits, Hb), TILE, 0, ctx->stream>>>(a);
    else lcp_score_kernel<8><<<dim3(splits, Hb), TILE, 0, ctx->stream>>>(a);   // full occupancy: the kernel waits on gathers (measured 4 -> 8 CTAs/SM: -14 %)
    lcp_reduce_kernel<<<(Hb + 127) / 128, 128, 0, ctx->stream>>>(partial, splits, Hb, d_scores + h0);
    ctx->launches += 2;
  }
  HOP_CUDA(ctx, cudaGetLastError());
  return HOP_OK;
}
