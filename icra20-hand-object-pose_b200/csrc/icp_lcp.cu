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
//   * The work is gather bound (one voxel cell + a short candidate list + one normal per scene point, all L2 resident) and
//     the per-iteration solve is the reference's SERIAL float LM (lm_replay.cuh), tens of thousands of dependent cycles that
//     want ~250 registers.  Two shapes, chosen per call (hop_icp_params.pipeline):
//       0 (default) persistent fused: the whole ICP of a hypothesis inside one CTA (icp_fused_kernel, ONE launch per batch);
//          a CTA carries up to one hypothesis per warp, so the solves of a CTA run in parallel on its warps and a hypothesis
//          whose LM takes 300 evaluations (2 % do; the mean is 31) delays nobody but itself.
//       1 iteration-synchronous: per ICP iteration
//            icp_moments_kernel  persistent CTAs stride over (active hypothesis, group of scene chunks) work items at full
//                                occupancy: correspondences -> 32-byte records in SHARED memory -> the 13x13 moments of the
//                                point-to-plane residual; 93 partial sums per item go to HBM (384 B; never the records);
//            icp_solve_kernel    one warp per active hypothesis with the register file to itself: fixed-order sum of the
//                                partials, the LM replay, PCL's convergence rule, state update, survivors -> next list.
//          Re-balances the machine every iteration, but every launch waits for its slowest LM run (measured: 3.7 vs 2.3 ms at
//          2 k x 10 k x 1024, 31.8 vs 28.2 ms at 10 k x 10 k x 16 384).  The point-to-point mode (mode 1, a closed-form solve)
//          runs on this pipeline; for mode 0 it is the cross-check of the fused kernel (test_icp_pipelines_agree).
//     Converged hypotheses retire at once in both.
//   * K5 is one flat launch (thread per hypothesis x scene point) + a fixed-order reduction of the tile partials.
//   * No tensor cores: these are gathers and small reductions, not dense contractions.
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <utility>

#include "hop_common.cuh"
#include "lm_replay_warp.cuh"

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

__device__ __forceinline__ int tri13(int i, int j) {  // index of (i,j), i<=j, in the row-major upper triangle
  return i * 13 - (i * (i - 1)) / 2 + (j - i);
}

// Exact minimiser of f(dR,dt) = y^T A y, y = [vec(dR - I); dt; 1], by damped Gauss-Newton on SE(3) from the
// identity, stopping like MINPACK's lmder does under PCL (relative reduction of the sum of squares <= sqrt(eps)).
// sums: the 91 reduced moments (shared memory).  Register resident: lane i < 13 owns row i of A, (R,t) and every
// small matrix are replicated in all lanes, rows meet through warp shuffles (no shared-memory round trips, no
// local memory).  All lanes return the same (R,t).
__device__ __forceinline__ void solve_exact(const float *sums, int lane, float *R, float *t, int max_inner) {
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
  for (int inner = 0; inner < max_inner; ++inner) {
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

// The reference's own solver (PCL's TransformationEstimationPointToPlane = float lmdif), replayed on the moments: see
// lm_replay.cuh.  Called by all lanes of one warp; returns the increment W(x) in all lanes.
__device__ __noinline__ void solve_lm_replay(const float *sums, int lane, lmr::LmrScratch *scr, float *R, float *t) {
  lmr::MomentsDev A{sums, lane, scr};
  float x[6];
  lmr::lm_replay_solve_warp(A, x, nullptr);
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
__device__ __forceinline__ bool solve_increment(const float *sums, int lane, lmr::LmrScratch *scr, float *R, float *t) {
  if constexpr (SOLVER == 0) {
    const lmr::MomentsDev A{sums, lane, scr};
    lmr::moments_prepare(A);
    if (lmr::translation_unconstrained(A)) return false;
    solve_lm_replay(sums, lane, scr, R, t);
  } else solve_exact(sums, lane, R, t, SOLVER == 1 ? 1 : 12);   // (a Gauss-Newton step = the first step of the exact minimiser)
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

// The 91 moments + sum d^2 + count = 93 sums are split into NSL slices of SZ = 96 / NSL accumulators, one slice per warp: every warp
// reads ALL records of a chunk and accumulates only its slice.  A lane therefore adds the records lane, lane + 32, ... of the whole
// scene in order, whatever the CTA width (4 warps x 24 or 8 warps x 12): the bits of the sums -- and with them every accept / stop
// decision of the replayed LM -- do not depend on the CTA shape the batch size picks, so a shard of a batch (strong scaling) refines
// to the same bits as the whole batch.
__host__ __device__ constexpr int tri_row(int k) { int i = 0; while (k >= 13 - i) { k -= 13 - i; ++i; } return i; }
__host__ __device__ constexpr int tri_col(int k) { int i = 0; while (k >= 13 - i) { k -= 13 - i; ++i; } return i + k; }

template <int SZ, int S, int E>
__device__ __forceinline__ void acc_one(float (&acc)[SZ], const float (&v)[13], float d2) {
  constexpr int k = SZ * S + E;
  if constexpr (k < 91) acc[E] = fmaf(v[tri_row(k)], v[tri_col(k)], acc[E]);
  else if constexpr (k == 91) acc[E] += d2;
  else if constexpr (k == 92) acc[E] += 1.f;
}
template <int SZ, int S, int... E>
__device__ __forceinline__ void acc_slice(float (&acc)[SZ], const float (&v)[13], float d2, std::integer_sequence<int, E...>) {
  (acc_one<SZ, S, E>(acc, v, d2), ...);
}

// phase B: this warp's slice of the moments over records [begin, end) in shared memory
template <int SZ, int S>
__device__ __forceinline__ void accumulate_chunk(float (&acc)[SZ], const float4 *rec0, const float4 *rec1, int begin, int end, int lane) {
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
    acc_slice<SZ, S>(acc, v, q1.w, std::make_integer_sequence<int, SZ>());
  }
}
template <int NSL>
__device__ __forceinline__ void accumulate_slice(float (&acc)[96 / NSL], int slice, const float4 *rec0, const float4 *rec1, int begin, int end, int lane) {
  constexpr int SZ = 96 / NSL;
  switch (slice) {
    case 0: accumulate_chunk<SZ, 0>(acc, rec0, rec1, begin, end, lane); break;
    case 1: accumulate_chunk<SZ, 1>(acc, rec0, rec1, begin, end, lane); break;
    case 2: accumulate_chunk<SZ, 2>(acc, rec0, rec1, begin, end, lane); break;
    case 3: accumulate_chunk<SZ, 3>(acc, rec0, rec1, begin, end, lane); break;
    default:
      if constexpr (NSL > 4) {
        switch (slice) {
          case 4: accumulate_chunk<SZ, 4>(acc, rec0, rec1, begin, end, lane); break;
          case 5: accumulate_chunk<SZ, 5>(acc, rec0, rec1, begin, end, lane); break;
          case 6: accumulate_chunk<SZ, 6>(acc, rec0, rec1, begin, end, lane); break;
          default: accumulate_chunk<SZ, 7>(acc, rec0, rec1, begin, end, lane); break;
        }
      }
      break;
  }
}

// phase A: CorrespondenceEstimation (exact 1-NN, d^2 <= max_dist^2) + CorrespondenceRejectorSurfaceNormal (rotated source
// normal . target normal > cos(angle)) of scene points [c0, c0 + cnt) -> 32-byte records in shared memory:
// rec0 = (p.xyz, n.(p - m)), rec1 = (n.xyz, d^2 | -1 when there is no correspondence)
template <int THREADS>
__device__ __forceinline__ void correspond_chunk(const CloudDev &scene, const float4 *__restrict__ model_nv, const NNGridDev &grid,
                                                 const Rigid &X, float cos_thr, float max_d2, int c0, int cnt, float4 *rec0,
                                                 float4 *rec1, int tid) {
  // (one point at a time per thread.  Keeping the same stage of 2 or 4 points in flight -- points, cells, candidate position k of every
  //  list, winners' normals -- was tried for the small batches that cannot hide the gathers' latency with other warps: bit-identical
  //  records, but 17.1 -> 25.9 ms at the headline size and 1.51 -> 1.70 ms at C2: the lists advance in lockstep to the longest one and the
  //  extra live points spill at 64 registers.  gpurun_out/r02p_*)
  for (int i = tid; i < cnt; i += THREADS) {
    const float4 sp = __ldg(&scene.pw[c0 + i]);
    const float3 p = rigid_apply(X, sp.x, sp.y, sp.z);
    float bd; float4 bp;
    const int j = nn_query(grid, p.x, p.y, p.z, bd, bp);
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = make_float4(0.f, 0.f, 0.f, -1.f);
    if (j >= 0 && bd <= max_d2) {
      const float4 sn = __ldg(&scene.nv[c0 + i]);
      const float4 mn = __ldg(&model_nv[j]);
      const float3 ns = rigid_rotate(X, sn.x, sn.y, sn.z);
      const float dot = ns.x * mn.x + ns.y * mn.y + ns.z * mn.z;
      if (dot >= cos_thr) {
        r0 = make_float4(p.x, p.y, p.z, mn.x * (p.x - bp.x) + mn.y * (p.y - bp.y) + mn.z * (p.z - bp.z));
        r1 = make_float4(mn.x, mn.y, mn.z, bd);
      }
    }
    rec0[i] = r0;
    rec1[i] = r1;
  }
}

// ------------------------------------------------------------------------------------------------------------
// K4, pipeline 0: one ICP iteration of every active hypothesis = icp_moments_kernel + icp_solve_kernel
// ------------------------------------------------------------------------------------------------------------
struct MomArgs {
  CloudDev scene;
  const float4 *model_nv;
  NNGridDev grid;
  const IcpState *state;   // already offset to the batch
  const int *list;         // active hypotheses of this iteration (batch-local ids)
  const int *n_active;
  int n_groups, group_pts; // a work item = (position in the list, group of group_pts scene points)
  float cos_thr;           // smallest float whose double value exceeds cos(angle)
  float max_d2;
  float *partial;          // [position in list][n_groups][96]: the item's 91 moments, sum d^2, count (mode 1: 17 Kabsch sums)
  NNGridDev sgrid;         // mode 1: the scene's own grid (reciprocal correspondences)
};

template <int THREADS, int CHUNK, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) icp_moments_kernel(MomArgs a) {
  constexpr int NSL = THREADS / 32, SZ = 96 / NSL;
  static_assert(NSL == 4 || NSL == 8, "one slice of the moments per warp");
  __shared__ __align__(16) float4 rec0[CHUNK], rec1[CHUNK];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_work = __ldg(a.n_active) * a.n_groups;
  for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
    const int pos = w / a.n_groups, g = w - pos * a.n_groups;
    const Rigid X = state_load(a.state[__ldg(&a.list[pos])].X);
    const int begin = g * a.group_pts, end = min(a.scene.n_padded, begin + a.group_pts);
    float acc[SZ];
#pragma unroll
    for (int e = 0; e < SZ; ++e) acc[e] = 0.f;
    for (int c0 = begin; c0 < end; c0 += CHUNK) {
      const int cnt = min(CHUNK, end - c0);
      correspond_chunk<THREADS>(a.scene, a.model_nv, a.grid, X, a.cos_thr, a.max_d2, c0, cnt, rec0, rec1, tid);
      __syncthreads();
      accumulate_slice<NSL>(acc, warp, rec0, rec1, 0, cnt, lane);
      __syncthreads();
    }
    float mine = 0.f;
#pragma unroll
    for (int e = 0; e < SZ; ++e) {
      const float tot = warp_sum(acc[e]);
      if (lane == e) mine = tot;
    }
    if (lane < SZ) a.partial[(size_t)w * 96 + SZ * warp + lane] = mine;
  }
}

// ---- mode 1: Utils::runICP(segment, model, T, max_corres_dist) (Utils.cpp:135-164) = PCL's default point-to-point ICP with
// reciprocal correspondences and TransformationEstimationSVD (pcl::umeyama without scaling) ------------------------------------
constexpr int KAB = 18;   // sum p (3), sum m (3), sum p m^T (9, row = p), sum d^2 as a double (2 words), count

// CorrespondenceEstimation::determineReciprocalCorrespondences: scene point i -> nearest model point j (d^2 <= max^2) -> the
// nearest SCENE point of j must be i again (and within max).  The scene does not move either: the reverse query is made with
// X^-1 m_j against the scene's own grid (a rigid motion keeps distances); "is i" = the same coordinates (two scene points with
// identical coordinates are both accepted, PCL keeps the one whose index the kd-tree returns).
template <int THREADS>
__device__ __forceinline__ void correspond_chunk_reciprocal(const CloudDev &scene, const NNGridDev &grid, const NNGridDev &sgrid,
                                                            const Rigid &X, const Rigid &Xi, float max_d2, int c0, int cnt, float4 *rec0,
                                                            float4 *rec1, int tid) {
  for (int i = tid; i < cnt; i += THREADS) {
    const float4 sp = __ldg(&scene.pw[c0 + i]);
    const float3 p = rigid_apply(X, sp.x, sp.y, sp.z);
    float bd; float4 bp;
    const int j = nn_query(grid, p.x, p.y, p.z, bd, bp);
    float4 r0 = make_float4(0.f, 0.f, 0.f, -1.f), r1 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j >= 0 && bd <= max_d2) {
      const float3 q = rigid_apply(Xi, bp.x, bp.y, bp.z);
      float ed; float4 ep;
      const int k = nn_query(sgrid, q.x, q.y, q.z, ed, ep);
      if (k >= 0 && ed <= max_d2 && ep.x == sp.x && ep.y == sp.y && ep.z == sp.z) {
        r0 = make_float4(p.x, p.y, p.z, bd);
        r1 = make_float4(bp.x, bp.y, bp.z, 0.f);
      }
    }
    rec0[i] = r0;
    rec1[i] = r1;
  }
}

template <int THREADS, int CHUNK, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) icp_kabsch_sums_kernel(MomArgs a) {
  constexpr int NW = THREADS / 32;
  __shared__ __align__(16) float4 rec0[CHUNK], rec1[CHUNK];
  __shared__ float s_sums[NW][KAB];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_work = __ldg(a.n_active) * a.n_groups;
  for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
    const int pos = w / a.n_groups, g = w - pos * a.n_groups;
    const Rigid X = state_load(a.state[__ldg(&a.list[pos])].X);
    const Rigid Xi = rigid_inverse(X);
    const int begin = g * a.group_pts, end = min(a.scene.n_padded, begin + a.group_pts);
    float acc[16];
    double acc_d2 = 0.0;   // PCL sums the squared distances of the correspondences into a double (DefaultConvergenceCriteria::calculateMSE)
#pragma unroll             // and compares successive means against 1e-12: a float sum would decide the stop by its own rounding
    for (int e = 0; e < 16; ++e) acc[e] = 0.f;
    for (int c0 = begin; c0 < end; c0 += CHUNK) {
      const int cnt = min(CHUNK, end - c0);
      correspond_chunk_reciprocal<THREADS>(a.scene, a.grid, a.sgrid, X, Xi, a.max_d2, c0, cnt, rec0, rec1, tid);
      __syncthreads();
      for (int i = tid; i < cnt; i += THREADS) {
        const float4 q0 = rec0[i];
        if (!(q0.w >= 0.f)) continue;
        const float4 q1 = rec1[i];
        acc[0] += q0.x; acc[1] += q0.y; acc[2] += q0.z;
        acc[3] += q1.x; acc[4] += q1.y; acc[5] += q1.z;
        acc[6] = fmaf(q0.x, q1.x, acc[6]); acc[7] = fmaf(q0.x, q1.y, acc[7]); acc[8] = fmaf(q0.x, q1.z, acc[8]);
        acc[9] = fmaf(q0.y, q1.x, acc[9]); acc[10] = fmaf(q0.y, q1.y, acc[10]); acc[11] = fmaf(q0.y, q1.z, acc[11]);
        acc[12] = fmaf(q0.z, q1.x, acc[12]); acc[13] = fmaf(q0.z, q1.y, acc[13]); acc[14] = fmaf(q0.z, q1.z, acc[14]);
        acc_d2 += (double)q0.w; acc[15] += 1.f;
      }
      __syncthreads();
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const float tot = warp_sum(acc[e]);
      if (lane == 0) s_sums[warp][e < 15 ? e : 17] = tot;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc_d2 += __shfl_xor_sync(0xffffffffu, acc_d2, o);
    if (lane == 0) { s_sums[warp][15] = __int_as_float(__double2hiint(acc_d2)); s_sums[warp][16] = __int_as_float(__double2loint(acc_d2)); }
    __syncthreads();
    float *out = a.partial + (size_t)w * 96;
    if (tid < 96) {
      float t = 0.f;
      if (tid < KAB && tid != 15 && tid != 16) {
#pragma unroll
        for (int q = 0; q < NW; ++q) t += s_sums[q][tid];
      } else if (tid == 15 || tid == 16) {
        double d = 0.0;
#pragma unroll
        for (int q = 0; q < NW; ++q) d += __hiloint2double(__float_as_int(s_sums[q][15]), __float_as_int(s_sums[q][16]));
        t = tid == 15 ? __int_as_float(__double2hiint(d)) : __int_as_float(__double2loint(d));
      }
      out[tid] = t;
    }
    __syncthreads();
  }
}

// eigen decomposition of a symmetric 3x3 (cyclic Jacobi, double): A -> diagonal, V = eigenvectors in columns
__device__ __forceinline__ void sym3_jacobi(double (&A)[3][3], double (&V)[3][3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    if (off < 1e-300) break;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
#pragma unroll
        for (int k = 0; k < 3; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - sn * akq; A[k][q] = sn * akp + c * akq; }
#pragma unroll
        for (int k = 0; k < 3; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - sn * aqk; A[q][k] = sn * apk + c * aqk; }
#pragma unroll
        for (int k = 0; k < 3; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - sn * vkq; V[k][q] = sn * vkp + c * vkq; }
      }
  }
}

// Kabsch / Umeyama (no scaling) from the 17 sums: R, t minimising sum |R p + t - m|^2.  S = sum (m - cm)(p - cp)^T = U D V^T,
// R = U diag(1, 1, det) V^T; the third columns of U and V are built as cross products, which folds the reflection case in.
// Returns false for a degenerate configuration (rank < 2): the caller keeps the identity increment.
__device__ __forceinline__ bool solve_kabsch(const float *sums, float *R, float *t) {
  const double n = (double)sums[17];
  double cp[3], cm[3], S[3][3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { cp[k] = (double)sums[k] / n; cm[k] = (double)sums[3 + k] / n; }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) S[r][c] = (double)sums[6 + 3 * c + r] - n * cm[r] * cp[c];   // sums[6 + 3 a + b] = sum p_a m_b
  double A[3][3], V[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) A[r][c] = S[0][r] * S[0][c] + S[1][r] * S[1][c] + S[2][r] * S[2][c];
  sym3_jacobi(A, V);
  // the two largest eigenvalues (columns i0, i1)
  const double e0 = A[0][0], e1 = A[1][1], e2 = A[2][2];
  int i0 = 0, i1 = 1;
  if (e0 <= e1 && e0 <= e2) { i0 = 1; i1 = 2; } else if (e1 <= e0 && e1 <= e2) { i0 = 0; i1 = 2; }
  const double d0 = i0 == 0 ? e0 : e1, d1 = i1 == 1 ? e1 : e2;
  const double big = fmax(d0, d1), small = fmin(d0, d1);
  if (!(big > 0.0) || !(small > 1e-24 * big)) return false;
  double v0[3], v1[3], u0[3], u1[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { v0[k] = i0 == 0 ? V[k][0] : V[k][1]; v1[k] = i1 == 1 ? V[k][1] : V[k][2]; }
  const double is0 = 1.0 / sqrt(d0), is1 = 1.0 / sqrt(d1);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    u0[r] = (S[r][0] * v0[0] + S[r][1] * v0[1] + S[r][2] * v0[2]) * is0;
    u1[r] = (S[r][0] * v1[0] + S[r][1] * v1[1] + S[r][2] * v1[2]) * is1;
  }
  const double v2[3] = {v0[1] * v1[2] - v0[2] * v1[1], v0[2] * v1[0] - v0[0] * v1[2], v0[0] * v1[1] - v0[1] * v1[0]};
  const double u2[3] = {u0[1] * u1[2] - u0[2] * u1[1], u0[2] * u1[0] - u0[0] * u1[2], u0[0] * u1[1] - u0[1] * u1[0]};
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double Rr[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { Rr[c] = u0[r] * v0[c] + u1[r] * v1[c] + u2[r] * v2[c]; R[3 * r + c] = (float)Rr[c]; }
    t[r] = (float)(cm[r] - (Rr[0] * cp[0] + Rr[1] * cp[1] + Rr[2] * cp[2]));
  }
  return true;
}

struct SolveArgs {
  const float *partial;  // [position in list][n_groups][96]
  int n_groups;
  IcpState *state;       // offset to the batch
  float *poses;          // offset to the batch: Hb x 16, written when a hypothesis finishes converged
  int32_t *iters_out, *conv_out;  // offset to the batch (may be null)
  const int *list;       // active hypotheses of this iteration
  const int *n_active;
  int *next_list;        // survivors are appended here (order is irrelevant to the results)
  int *next_count;
  int max_iter;
  double abs_mse_eps;
};

// TransformationEstimationPointToPlane (the reference's LM) + DefaultConvergenceCriteria: one warp per active hypothesis, no
// register cap (the LM replay keeps its 6x6 matrices, the double Gram matrix and its Cholesky factor in registers).
constexpr int SOLVE_WARPS = 4;
template <int SOLVER>
__global__ void __launch_bounds__(SOLVE_WARPS * 32) icp_solve_kernel(SolveArgs a) {
  __shared__ __align__(16) float s_tot[SOLVE_WARPS][96];
  __shared__ __align__(16) lmr::LmrScratch s_scr[SOLVER == 0 ? SOLVE_WARPS : 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pos = blockIdx.x * SOLVE_WARPS + warp;
  if (pos >= __ldg(a.n_active)) return;   // (no block-wide barrier below)
  const int h = __ldg(&a.list[pos]);
  IcpState *st = a.state + h;
  float *sums = s_tot[warp];
  {
    const float *p = a.partial + (size_t)pos * a.n_groups * 96;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
    double dsum = 0.0;
    for (int g = 0; g < a.n_groups; ++g) {   // fixed order: bits do not depend on which CTA produced which partial
      t0 += __ldcs(p + g * 96 + lane); t1 += __ldcs(p + g * 96 + 32 + lane); t2 += __ldcs(p + g * 96 + 64 + lane);
      if (SOLVER == 3) dsum += __hiloint2double(__float_as_int(__ldcs(p + g * 96 + 15)), __float_as_int(__ldcs(p + g * 96 + 16)));
    }
    sums[lane] = t0; sums[32 + lane] = t1; sums[64 + lane] = t2;
    if (SOLVER == 3) { __syncwarp(); if (lane == 0) { sums[15] = __int_as_float(__double2hiint(dsum)); sums[16] = __int_as_float(__double2loint(dsum)); } }
  }
  __syncwarp();
  const float cnt_f = SOLVER == 3 ? sums[17] : sums[92];
  const double sumd2 = SOLVER == 3 ? __hiloint2double(__float_as_int(sums[15]), __float_as_int(sums[16])) : (double)sums[91];
  const int cnt = (int)(cnt_f + 0.5f);
  int iters = st->iters;
  bool converged = false, finished = false, runaway = false;
  Rigid X = state_load(st->X);
  Rigid inc;
  if constexpr (SOLVER == 3) {
    // TransformationEstimationSVD works from 3 correspondences up; a degenerate set leaves the identity
    if (cnt >= 3 && !solve_kabsch(sums, inc.r, inc.t)) {
#pragma unroll
      for (int e = 0; e < 9; ++e) inc.r[e] = (e % 4 == 0) ? 1.f : 0.f;
      inc.t[0] = inc.t[1] = inc.t[2] = 0.f;
    }
  } else if (cnt >= 6) runaway = !solve_increment<SOLVER>(sums, lane, &s_scr[SOLVER == 0 ? warp : 0], inc.r, inc.t);
  if (SOLVER != 3 && cnt >= 6 && !runaway) {
    bool finite = true;   // |q| > 1 inside the reference's LM gives NaN residuals; a step accepted there ends as a NaN transform
#pragma unroll
    for (int e = 0; e < 9; ++e) finite = finite && isfinite(inc.r[e]);
    runaway = !(finite && isfinite(inc.t[0]) && isfinite(inc.t[1]) && isfinite(inc.t[2]));
  }
  if (cnt < 3) {
    finished = true;  // "Not enough correspondences": hasConverged() false -> identity -> pose unchanged
  } else if (runaway) {
    // the reference's LM slides the scene metres off the model here (or returns a NaN transform); its next iteration finds no
    // correspondences: not converged, pose unchanged
    ++iters;
    finished = true;
    if (lane == 0) st->iters = iters;
  } else {
    if (SOLVER == 3 || cnt >= 6) {
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
        mse = sumd2 / (double)cnt;
        if (fabs(mse - st->prev_mse) < a.abs_mse_eps) converged = true;
      }
    }
    finished = converged;
    if (lane == 0) {
#pragma unroll
      for (int e = 0; e < 9; ++e) { st->X[e] = X.r[e]; st->inc[e] = inc.r[e]; }
#pragma unroll
      for (int e = 0; e < 3; ++e) { st->X[9 + e] = X.t[e]; st->inc[9 + e] = inc.t[e]; }
      st->prev_mse = mse;
      st->iters = iters;
    }
  }
  if (!finished && lane == 0) a.next_list[atomicAdd(a.next_count, 1)] = h;
  if (finished && lane == 0) {
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
  int slots_max;     // hypotheses a CTA carries at a time (<= its warps): the batch's fair share per CTA
};

// One hypothesis' state between its ICP iterations (shared memory; the slot is owned by warp `slot` during the solve)
struct __align__(16) FusedSlot {
  float X[12];        // scene -> model frame, the iterate
  float inc[12];      // previous increment (PCL keeps transformation_ when LM early-returns)
  double prev_mse;
  int h;              // hypothesis of the batch in this slot, -1 = free
  int iters;
};

// THREADS per CTA (a multiple of 128: four moment slices x THREADS/128 parts of the chunk), CHUNK scene points whose
// records sit in shared memory at a time, PROF = with the cycle accounting, MINB resident CTAs per SM.
//
// A CTA carries up to one hypothesis per warp ("slots").  One round = for every occupied slot the cooperative passes over the
// scene (phase A correspondences + phase B moments, all warps on one slot at a time: the same gather efficiency as one
// hypothesis per CTA), then ALL solves of the round at once, warp w on slot w.  The per-iteration solve is the reference's serial
// LM (lm_replay.cuh: tens of thousands of dependent cycles); with one hypothesis per CTA the other warps idle through it (59 %
// of the CTA's cycles at 2 k scene points, 19 % at 10 k), with slots it costs 1/THREADS*32 of that per hypothesis.  A slot that
// finishes takes the next hypothesis of the queue at the start of the next round.  slots_max (host, launch_fused) never exceeds the
// fair share of the batch per CTA, so a small batch still spreads over every SM instead of filling the slots of the first CTAs.
template <int THREADS, int CHUNK, bool PROF, int MINB, int SOLVER>
__global__ void __launch_bounds__(THREADS, MINB) icp_fused_kernel(FusedArgs a) {
  constexpr int NW = THREADS / 32, SZ = 96 / NW;
  static_assert(NW == 4 || NW == 8, "one slice of the moments per warp");
  extern __shared__ __align__(16) float4 fused_smem[];
  float4 *rec0 = fused_smem, *rec1 = fused_smem + CHUNK;
  __shared__ __align__(16) float s_tot[NW][96];
  __shared__ FusedSlot s_slot[NW];
  // the LM replay's per-warp scratch lives in the record buffers: they are idle while the round's solves run, and shared memory
  // taken from L1 costs phase A its hit rate (2.3 KB x warps x 8 CTAs per SM: 18.8 -> 21.2 ms at the headline size)
  static_assert(sizeof(lmr::LmrScratch) * NW <= 2 * CHUNK * sizeof(float4), "LM scratch does not fit the record buffers");
  lmr::LmrScratch *s_scr = reinterpret_cast<lmr::LmrScratch *>(fused_smem);
  __shared__ int s_ctl[2];   // [0] occupied slots this round  [1] queue exhausted
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_slots = min(NW, max(1, a.slots_max));
  // optional cycle accounting (thread 0 of every CTA): [0] phase A  [1] wait at the barrier after A  [2] phase B + barrier
  // [3] reduction + solve + broadcast  [4] passes  [5] whole CTA life time
  long long t_acc[6] = {0, 0, 0, 0, 0, 0}, t_mark = 0;
  const long long t_start = PROF ? clock64() : 0;
  if (tid < NW) s_slot[tid].h = -1;
  if (tid == 0) s_ctl[1] = 0;
  __syncthreads();
  for (;;) {
    // ---- refill: every free slot takes the next hypothesis of the queue ----
    if (warp == 0) {
      int mine = lane < n_slots ? s_slot[lane].h : 0;
      const int exhausted = s_ctl[1];
      __syncwarp();   // (every lane has read the flag before any lane sets it)
      if (lane < n_slots && mine < 0 && !exhausted) {
        const int h = atomicAdd(a.counter, 1);
        if (h < a.H) {
          FusedSlot &S = s_slot[lane];
          const Rigid X = rigid_inverse(rigid_load_colmajor(a.poses + 16 * (size_t)h));
#pragma unroll
          for (int e = 0; e < 9; ++e) { S.X[e] = X.r[e]; S.inc[e] = (e % 4 == 0) ? 1.f : 0.f; }
#pragma unroll
          for (int e = 0; e < 3; ++e) { S.X[9 + e] = X.t[e]; S.inc[9 + e] = 0.f; }
          S.prev_mse = DBL_MAX; S.iters = 0; S.h = h;
          mine = h;
        } else s_ctl[1] = 1;   // (every writer stores 1)
      }
      const unsigned occ = __ballot_sync(0xffffffffu, lane < n_slots && mine >= 0);
      if (lane == 0) s_ctl[0] = __popc(occ);
    }
    __syncthreads();
    if (s_ctl[0] == 0) break;
    // ---- the passes over the scene, one occupied slot at a time, all warps together ----
    for (int g = 0; g < n_slots; ++g) {
      if (s_slot[g].h < 0) continue;   // uniform
      Rigid X;
#pragma unroll
      for (int e = 0; e < 9; ++e) X.r[e] = s_slot[g].X[e];
      X.t[0] = s_slot[g].X[9]; X.t[1] = s_slot[g].X[10]; X.t[2] = s_slot[g].X[11];
      float acc[SZ];
#pragma unroll
      for (int e = 0; e < SZ; ++e) acc[e] = 0.f;
      for (int c0 = 0; c0 < a.scene.n_padded; c0 += CHUNK) {
        const int cnt = min(CHUNK, a.scene.n_padded - c0);
        // ---- phase A: correspondences of this chunk -> shared memory ----
        if (PROF && tid == 0) t_mark = clock64();
        correspond_chunk<THREADS>(a.scene, a.model_nv, a.grid, X, a.cos_thr, a.max_d2, c0, cnt, rec0, rec1, tid);
        if (PROF && tid == 0) { const long long t = clock64(); t_acc[0] += t - t_mark; t_mark = t; }
        __syncthreads();
        if (PROF && tid == 0) { const long long t = clock64(); t_acc[1] += t - t_mark; t_mark = t; }
        // ---- phase B: this warp's slice of the moments over the whole chunk ----
        accumulate_slice<NW>(acc, warp, rec0, rec1, 0, cnt, lane);
        __syncthreads();
        if (PROF && tid == 0) { const long long t = clock64(); t_acc[2] += t - t_mark; t_mark = t; }
      }
      // ---- reduce: lanes -> this slot's 93 sums ----
      float mine = 0.f;
#pragma unroll
      for (int e = 0; e < SZ; ++e) {
        const float tot = warp_sum(acc[e]);
        if (lane == e) mine = tot;
      }
      if (lane < SZ) s_tot[g][SZ * warp + lane] = mine;
      if (PROF && tid == 0) { const long long t = clock64(); t_acc[3] += t - t_mark; t_mark = t; t_acc[4] += 1; }
    }
    __syncthreads();
    // ---- the solves of this round, warp w on slot w ----
    if (PROF && tid == 0) t_mark = clock64();
    if (warp < n_slots && s_slot[warp].h >= 0) {
      FusedSlot &S = s_slot[warp];
      const float *sums = s_tot[warp];
      const float cnt_f = sums[92], sumd2 = sums[91];
      const int cnt = (int)(cnt_f + 0.5f);
      Rigid X;
#pragma unroll
      for (int e = 0; e < 9; ++e) X.r[e] = S.X[e];
      X.t[0] = S.X[9]; X.t[1] = S.X[10]; X.t[2] = S.X[11];
      int iters = S.iters;
      bool finished = false, converged = false, runaway = false;
      Rigid inc;
      if (cnt >= 6) runaway = !solve_increment<SOLVER>(sums, lane, &s_scr[SOLVER == 0 ? warp : 0], inc.r, inc.t);
      if (cnt >= 6 && !runaway) {
        bool finite = true;   // |q| > 1 inside the reference's LM gives NaN residuals; a step accepted there ends as a NaN transform
#pragma unroll
        for (int e = 0; e < 9; ++e) finite = finite && isfinite(inc.r[e]);
        finite = finite && isfinite(inc.t[0]) && isfinite(inc.t[1]) && isfinite(inc.t[2]);
        runaway = !finite;
      }
      if (cnt < 3) {
        finished = true;  // "Not enough correspondences": hasConverged() false -> pose unchanged
      } else if (runaway) {
        // the reference's LM slides the scene metres off the model here (or returns a NaN transform); its next iteration finds
        // no correspondences: not converged, pose unchanged
        ++iters;
        finished = true;
      } else {
        if (cnt >= 6) {}
        else if (cnt >= 4) {  // Eigen LM: m < n -> ImproperInputParameters, x stays 0 -> identity increment
#pragma unroll
          for (int e = 0; e < 9; ++e) inc.r[e] = (e % 4 == 0) ? 1.f : 0.f;
          inc.t[0] = inc.t[1] = inc.t[2] = 0.f;
        } else {              // PCL's LM returns early with < 4 correspondences, transformation_ keeps its old value
#pragma unroll
          for (int e = 0; e < 9; ++e) inc.r[e] = S.inc[e];
          inc.t[0] = S.inc[9]; inc.t[1] = S.inc[10]; inc.t[2] = S.inc[11];
        }
        X = rigid_compose(inc, X);
        ++iters;
        double prev_mse = S.prev_mse;
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
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int e = 0; e < 9; ++e) { S.X[e] = X.r[e]; S.inc[e] = inc.r[e]; }
#pragma unroll
          for (int e = 0; e < 3; ++e) { S.X[9 + e] = X.t[e]; S.inc[9 + e] = inc.t[e]; }
          S.prev_mse = prev_mse;
        }
      }
      __syncwarp();   // (the lanes' reads of the slot above come before lane 0 rewrites it, whichever branch was taken)
      if (lane == 0) {
        S.iters = iters;
        if (finished) {
          const int h = S.h;
          if (converged) { const Rigid P = rigid_inverse(X); rigid_store_colmajor(P, a.poses + 16 * (size_t)h); }
          if (a.iters_out) a.iters_out[h] = iters;
          if (a.conv_out) a.conv_out[h] = converged ? 1 : 0;
          S.h = -1;
        }
      }
    }
    __syncthreads();
    if (PROF && tid == 0) { const long long t = clock64(); t_acc[3] += t - t_mark; }
  }
  if (PROF && tid == 0) {
    t_acc[5] = clock64() - t_start;
#pragma unroll
    for (int k = 0; k < 6; ++k) atomicAdd((unsigned long long *)a.prof + k, (unsigned long long)t_acc[k]);
  }
}

template <int THREADS, int CHUNK, bool PROF, int MINB, int SOLVER>
static cudaError_t launch_fused(hop_ctx *ctx, FusedArgs f, int H) {
  const size_t smem = 2 * (size_t)CHUNK * sizeof(float4);
  cudaError_t e = ctx->func_smem_optin(icp_fused_kernel<THREADS, CHUNK, PROF, MINB, SOLVER>, smem);
  if (e != cudaSuccess) return e;
  const int grid_ctas = (int)std::min<long>((long)H, (long)ctx->sm_count * MINB);
  // Slots per CTA.  Since the solve became short against a pass over the scene (lm_replay_warp.cuh) the parallel solves of a round buy
  // less than its barrier costs (every slot waits for the round's longest solve): 128-thread CTAs carry two hypotheses only when each has
  // more than four to do, and a 256-thread CTA with three or fewer takes them one after the other (C2: 0.96 -> 0.85 ms, headline 16.2 -> 16.0;
  // profiles/r02_experiments.txt, runs r02s, r03s)
  const int fair = (H + grid_ctas - 1) / grid_ctas;
  const int slots = THREADS == 128 ? (fair <= 4 ? 1 : 2) : (fair <= 3 ? 1 : fair);
  f.slots_max = ctx->tune.fused_slots > 0 ? ctx->tune.fused_slots : slots;
  icp_fused_kernel<THREADS, CHUNK, PROF, MINB, SOLVER><<<grid_ctas, THREADS, smem, ctx->stream>>>(f);
  return cudaGetLastError();
}

// CTA shape by batch size.  Small batches: 256-thread CTAs (a hypothesis finishes sooner, shorter tail); large batches: 128-thread
// CTAs (the serial small solve idles 3 warps instead of 7), 8 of them per SM at 64 registers -- the few spilled bytes cost less
// than the extra warps hide (15.39 -> 14.78 ms at the headline size; small batches lose with it).  Record chunk = 4 points per
// thread: the rest of the 256 KB stays L1.
template <int SOLVER>
static cudaError_t launch_fused_solver(hop_ctx *ctx, const FusedArgs &f, int H, bool small_ctas, bool prof_on) {
  if constexpr (SOLVER == 0) {   // (the cycle accounting exists for the parity solver only)
    if (prof_on) return small_ctas ? launch_fused<128, 512, true, 6, SOLVER>(ctx, f, H) : launch_fused<256, 1024, true, 3, SOLVER>(ctx, f, H);
  }
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

// CTA = (hypothesis, split): one thread per scene point of every `splits`-th tile; the pose algebra is done once per CTA.  Every
// TILE's sum is reduced in a fixed order and stored on its own (partial[h][tile]); lcp_reduce_kernel then adds the tiles in order:
// the bits of a hypothesis' score depend on the scene only, not on the batch size (which picks `splits`) or its position in it --
// a sharded batch (strong scaling) must merge to the winners of the whole batch, score for score.
template <int MINB>
__global__ void __launch_bounds__(TILE, MINB) lcp_score_kernel(LcpArgs a) {
  __shared__ float s_w[TILE / 32];
  const int h = blockIdx.y;
  const Rigid T = rigid_load_colmajor(a.poses + 16 * (size_t)h);
  const Rigid Ti = rigid_inverse(T);
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    float score = 0.f;
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
    score = warp_sum(score);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = score;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < TILE / 32; ++w) s += s_w[w];
      a.partial[(size_t)h * a.n_tiles + tile] = s;
    }
    __syncthreads();
  }
}

// fixed-order sum of the tile partials (deterministic bits run to run and batch to batch)
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

template <int SOLVER>
static void launch_solve(hop_ctx *ctx, const SolveArgs &s, int Hb) {
  icp_solve_kernel<SOLVER><<<(Hb + SOLVE_WARPS - 1) / SOLVE_WARPS, SOLVE_WARPS * 32, 0, ctx->stream>>>(s);
}


// diagnostics: the LM replay on caller-supplied moments, one warp per problem (tests/test_gpu_lm.py pins lm_replay_warp.cuh against
// the host build of the scalar program)
__global__ void __launch_bounds__(32) lm_debug_kernel(const float *__restrict__ sums, int n, float *__restrict__ x_out, int32_t *__restrict__ nfev_out,
                                                       int32_t *__restrict__ status_out, long long *__restrict__ cycles_out) {
  __shared__ __align__(16) float s_sums[96];
  __shared__ __align__(16) lmr::LmrScratch s_scr;
  const int lane = threadIdx.x;
  for (int p = blockIdx.x; p < n; p += gridDim.x) {
    __syncwarp();
    for (int e = lane; e < 96; e += 32) s_sums[e] = sums[(size_t)p * 96 + e];
    __syncwarp();
    const lmr::MomentsDev A{s_sums, lane, &s_scr};
    lmr::moments_prepare(A);
    float x[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int nfev = 0, status = -1;
    const long long t0 = clock64();
    if (!lmr::translation_unconstrained(A)) status = lmr::lm_replay_solve_warp(A, x, &nfev);
    const long long t1 = clock64();
    if (lane == 0) {
      if (cycles_out) cycles_out[p] = t1 - t0;
#pragma unroll
      for (int j = 0; j < 6; ++j) x_out[(size_t)p * 6 + j] = x[j];
      nfev_out[p] = nfev;
      status_out[p] = status;
    }
  }
}

}  // namespace

int hop_launch_icp(hop_ctx *ctx, const CloudDev &scene, const CloudDev &model, const NNGridDev &grid, const NNGridDev *scene_grid,
                   float *d_poses, int H, const hop_icp_params &p, int32_t *d_iters, int32_t *d_conv) {
  if (H <= 0) return HOP_OK;
  if (p.mode != 0 && p.mode != 1) { ctx->err = "hop_icp_refine: mode must be 0 (point-to-plane) or 1 (point-to-point)"; return HOP_EINVAL; }
  if (p.mode == 1 && !scene_grid) { ctx->err = "hop_icp_refine: mode 1 needs the scene grid"; return HOP_EINVAL; }
  if (p.solver < 0 || p.solver > 2) { ctx->err = "hop_icp_refine: solver must be 0 (reference LM), 1 (Gauss-Newton) or 2 (exact minimiser)"; return HOP_EINVAL; }
  const int max_iter = p.max_iter < 1 ? 1 : p.max_iter;
  const float cos_thr = float_above(cos((double)p.angle_deg / 180.0 * M_PI));
  const float max_d2 = p.max_dist * p.max_dist;
  if (p.pipeline != 1 && p.mode == 0) {
    // persistent fused pipeline (default): one launch for the whole batch
    FusedArgs f;
    f.scene = scene; f.model_nv = model.nv; f.grid = grid; f.poses = d_poses; f.H = H; f.iters_out = d_iters; f.conv_out = d_conv;
    f.counter = ctx->d_counter;
    f.cos_thr = cos_thr; f.max_d2 = max_d2; f.max_iter = max_iter; f.abs_mse_eps = p.abs_mse_eps;
    HOP_CUDA(ctx, cudaMemsetAsync(ctx->d_counter, 0, sizeof(int), ctx->stream));
    const int variant = ctx->tune.fused_variant;
    const bool prof_on = ctx->tune.fused_profile;
    if (prof_on && !ctx->d_fused_prof) { HOP_CUDA(ctx, cudaMalloc(&ctx->d_fused_prof, 6 * sizeof(long long))); }
    if (prof_on) HOP_CUDA(ctx, cudaMemsetAsync(ctx->d_fused_prof, 0, 6 * sizeof(long long), ctx->stream));
    f.prof = prof_on ? ctx->d_fused_prof : nullptr;
    {
      ProfScope ps(ctx, HOP_PROF_ICP_FUSED);
      // by the batch's work, hypotheses x scene points (24 x SMs hypotheses at the 2 k-point scene the rule was measured on; a 10 k-point
      // scene reaches it at a fifth of that: one rank's 2048-hypothesis share under strong scaling runs 4 % faster with the small CTAs)
      const bool small_ctas = variant == 2 || (variant == 0 && (long long)H * scene.n_padded >= 24LL * ctx->sm_count * 2048LL);
      HOP_CUDA(ctx, p.solver == 0   ? launch_fused_solver<0>(ctx, f, H, small_ctas, prof_on)
                    : p.solver == 1 ? launch_fused_solver<1>(ctx, f, H, small_ctas, prof_on)
                                    : launch_fused_solver<2>(ctx, f, H, small_ctas, prof_on));
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
  // ---- iteration-synchronous pipeline (pipeline 1, and mode 1): sums + solve, two launches per ICP iteration ----
  constexpr int MOM_THREADS = 128, MOM_CHUNK = 512, MOM_MINB = 8;
  const int n_chunks = (scene.n_padded + MOM_CHUNK - 1) / MOM_CHUNK;
  // a work item = one hypothesis x a group of chunks: small enough that a handful of active hypotheses still spread over the
  // machine (latency of the late iterations), large enough that the partial sums stay a small fraction of the traffic
  const int group_chunks = ctx->tune.mom_group_chunks > 0 ? ctx->tune.mom_group_chunks : std::max(2, (n_chunks + 15) / 16);
  const int n_groups = (n_chunks + group_chunks - 1) / group_chunks;
  const int group_pts = group_chunks * MOM_CHUNK;
  const size_t part_per_h = (size_t)n_groups * 96 * sizeof(float);
  const size_t part_budget = (size_t)256 << 20;
  const int Hb_max = (int)std::min<size_t>((size_t)H, std::max<size_t>(1, part_budget / part_per_h));
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t state_bytes = up(sizeof(IcpState) * (size_t)Hb_max), list_bytes = up(sizeof(int) * (size_t)Hb_max);
  const size_t cnt_bytes = up(sizeof(int) * (size_t)(max_iter + 1));
  char *base = (char *)ctx->ensure_work(state_bytes + 2 * list_bytes + cnt_bytes + part_per_h * Hb_max);
  if (!base) { ctx->err = "hop_icp_refine: work buffer allocation failed"; return HOP_ENOMEM; }
  IcpState *state = (IcpState *)base;
  int *lists[2] = {(int *)(base + state_bytes), (int *)(base + state_bytes + list_bytes)};
  int *counters = (int *)(base + state_bytes + 2 * list_bytes);
  float *partial = (float *)(base + state_bytes + 2 * list_bytes + cnt_bytes);

  for (int h0 = 0; h0 < H; h0 += Hb_max) {
    const int Hb = std::min(Hb_max, H - h0);
    const int n_init = std::max(Hb, max_iter + 1);
    icp_init_kernel<<<(n_init + 127) / 128, 128, 0, ctx->stream>>>(d_poses + 16 * (size_t)h0, Hb, state, lists[0], counters, max_iter + 1);
    ctx->launches += 1;
    MomArgs c;
    c.scene = scene; c.model_nv = model.nv; c.grid = grid; c.state = state; c.n_groups = n_groups; c.group_pts = group_pts;
    c.cos_thr = cos_thr; c.max_d2 = max_d2; c.partial = partial;
    if (scene_grid) c.sgrid = *scene_grid; else c.sgrid = grid;
    SolveArgs s;
    s.partial = partial; s.n_groups = n_groups; s.state = state; s.poses = d_poses + 16 * (size_t)h0;
    s.iters_out = d_iters ? d_iters + h0 : nullptr; s.conv_out = d_conv ? d_conv + h0 : nullptr;
    s.max_iter = max_iter; s.abs_mse_eps = p.abs_mse_eps;
    const int mom_grid = (int)std::min<long>((long)n_groups * Hb, (long)ctx->sm_count * MOM_MINB);
    for (int it = 0; it < max_iter; ++it) {
      c.list = lists[it & 1]; c.n_active = counters + it;
      s.list = c.list; s.n_active = c.n_active; s.next_list = lists[(it + 1) & 1]; s.next_count = counters + it + 1;
      {
        ProfScope ps(ctx, HOP_PROF_ICP_CORRESPOND);
        if (p.mode == 1) icp_kabsch_sums_kernel<MOM_THREADS, MOM_CHUNK, MOM_MINB><<<mom_grid, MOM_THREADS, 0, ctx->stream>>>(c);
        else icp_moments_kernel<MOM_THREADS, MOM_CHUNK, MOM_MINB><<<mom_grid, MOM_THREADS, 0, ctx->stream>>>(c);
      }
      ProfScope ps(ctx, HOP_PROF_ICP_SOLVE);
      if (p.mode == 1) launch_solve<3>(ctx, s, Hb);
      else if (p.solver == 0) launch_solve<0>(ctx, s, Hb);
      else if (p.solver == 1) launch_solve<1>(ctx, s, Hb);
      else launch_solve<2>(ctx, s, Hb);
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
  float *partial = (float *)ctx->ensure_work(sizeof(float) * (size_t)std::min(H, Hb_max) * n_tiles);
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
    lcp_reduce_kernel<<<(Hb + 127) / 128, 128, 0, ctx->stream>>>(partial, n_tiles, Hb, d_scores + h0);
    ctx->launches += 2;
  }
  HOP_CUDA(ctx, cudaGetLastError());
  return HOP_OK;
}

int hop_debug_lm_solve_launch(hop_ctx *ctx, const float *d_sums, int n, float *d_x, int32_t *d_nfev, int32_t *d_status, long long *d_cycles) {
  if (n <= 0) return HOP_OK;
  lm_debug_kernel<<<std::min(n, ctx->sm_count * 16), 32, 0, ctx->stream>>>(d_sums, n, d_x, d_nfev, d_status, d_cycles);
  ctx->launches += 1;
  HOP_CUDA(ctx, cudaGetLastError());
  return HOP_OK;
}
