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
//   * One team of TEAM warps owns one hypothesis (TEAM=1: one warp per hypothesis).  All teams of a CTA read the
//     same scene, which a dedicated producer warp streams through shared memory in 256-point tiles with 1-D bulk
//     (TMA) copies completed on mbarriers; small scenes stay resident in shared memory across iterations.
//   * Per iteration every lane accumulates the moments of its correspondences in registers; a warp-shuffle butterfly
//     reduces them and the small solve runs cooperatively in the warp out of shared memory.  No tensor cores: these
//     are gather-bound small reductions.
//   * Hypotheses are pulled from a global atomic queue at iteration boundaries, so converged hypotheses free their
//     team immediately (ICP stops after 2-4 of the 10 allowed iterations for most hypotheses).
#include <cfloat>
#include <cmath>

#include "hop_common.cuh"

namespace {

constexpr int TILE = HOP_TILE_PTS;
constexpr int TILE_BYTES = TILE * 16;  // per stream (positions / normals)

// ------------------------------------------------------------------------------------------------------------
// shared-memory carve-up
// ------------------------------------------------------------------------------------------------------------
struct SmemLayout {
  float4 *tileP, *tileN;
  uint64_t *full, *empty;
  int *hyp;      // [2][NT]
  float *part;   // [NW][NACC_PAD]   per-warp reduced sums
  float *work;   // [NW][WORK]       per-warp solver workspace
};

template <int NW, int NT, int NACC_PAD, int WORK>
__device__ __forceinline__ SmemLayout carve(unsigned char *smem, int stages) {
  SmemLayout L;
  L.tileP = reinterpret_cast<float4 *>(smem);
  L.tileN = L.tileP + (size_t)stages * TILE;
  unsigned char *p = reinterpret_cast<unsigned char *>(L.tileN + (size_t)stages * TILE);
  L.full = reinterpret_cast<uint64_t *>(p); p += sizeof(uint64_t) * stages;
  L.empty = reinterpret_cast<uint64_t *>(p); p += sizeof(uint64_t) * stages;
  L.hyp = reinterpret_cast<int *>(p); p += sizeof(int) * 2 * NT;
  p = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15));
  L.part = reinterpret_cast<float *>(p); p += sizeof(float) * NW * NACC_PAD;
  L.work = reinterpret_cast<float *>(p);
  return L;
}
static size_t smem_bytes(int stages, int NW, int NT, int NACC_PAD, int WORK) {
  return (size_t)stages * TILE * 32 + 16 * (size_t)stages + sizeof(int) * 2 * NT + 16 + sizeof(float) * NW * (size_t)(NACC_PAD + WORK);
}

// ------------------------------------------------------------------------------------------------------------
// small dense algebra used by the per-iteration solve (uniform across the warp)
// ------------------------------------------------------------------------------------------------------------
// solve (H + lambda*diag(H)) x = -g for symmetric 6x6 H (full storage); returns false when not positive definite
__device__ __forceinline__ bool chol_solve6(const float *Hs, const float *g, float lambda, float *x) {
  float L[6][6];
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
        L[i][i] = sqrtf(s);
      } else {
        L[i][j] = s / L[j][j];
      }
    }
  }
  float y[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    float s = -g[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= L[i][k] * y[k];
    y[i] = s / L[i][i];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    float s = y[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) s -= L[k][i] * x[k];
    x[i] = s / L[i][i];
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
  __device__ __forceinline__ void add(float3 p, float3 m, float3 n, float d2) {
    float v[13];
    v[0] = n.x * p.x; v[1] = n.x * p.y; v[2] = n.x * p.z;
    v[3] = n.y * p.x; v[4] = n.y * p.y; v[5] = n.y * p.z;
    v[6] = n.z * p.x; v[7] = n.z * p.y; v[8] = n.z * p.z;
    v[9] = n.x; v[10] = n.y; v[11] = n.z;
    v[12] = n.x * (p.x - m.x) + n.y * (p.y - m.y) + n.z * (p.z - m.z);
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
  __device__ __forceinline__ void add(float3 p, float3 m, float3 n, float d2) {
    float J[6];
    J[0] = p.y * n.z - p.z * n.y; J[1] = p.z * n.x - p.x * n.z; J[2] = p.x * n.y - p.y * n.x;
    J[3] = n.x; J[4] = n.y; J[5] = n.z;
    float c = n.x * (p.x - m.x) + n.y * (p.y - m.y) + n.z * (p.z - m.z);
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
constexpr int WORK = 512;  // per warp: [0,416) solver (A 169, y, gy, J 78, B 78, H 36, g 6)  [416,512) team totals

__device__ __forceinline__ int tri13(int i, int j) {  // index of (i,j), i<=j, in the row-major upper triangle
  return i * 13 - (i * (i - 1)) / 2 + (j - i);
}

// Exact minimiser of f(dR,dt) = y^T A y, y = [vec(dR - I); dt; 1], by damped Gauss-Newton on SE(3) from the
// identity, stopping like MINPACK's lmder does under PCL (relative reduction of the sum of squares <= sqrt(eps)).
// sums: the 91 reduced moments (shared memory, this warp's row).  W: this warp's workspace.  All lanes return the
// same (R,t).
__device__ void solve_exact(const float *sums, float *W, int lane, float *R, float *t) {
  float *A = W;            // 13x13
  float *y = W + 176;      // 13
  float *gy = W + 192;     // 13
  float *Jm = W + 208;     // 13x6 (rows 12.. zero)
  float *B = W + 288;      // 13x6
  float *Hm = W + 368;     // 6x6
  float *gv = W + 404;     // 6
  for (int e = lane; e < 169; e += 32) {
    int i = e / 13, j = e % 13;
    A[e] = sums[i <= j ? tri13(i, j) : tri13(j, i)];
  }
  R[0] = 1.f; R[1] = 0.f; R[2] = 0.f; R[3] = 0.f; R[4] = 1.f; R[5] = 0.f; R[6] = 0.f; R[7] = 0.f; R[8] = 1.f;
  t[0] = t[1] = t[2] = 0.f;
  __syncwarp();
  float f = A[168];
  float lambda = 0.f;
  const float ftol = 3.4526698e-4f;  // sqrt(FLT_EPSILON)
  // gy = A y at the identity is the last column of A
  if (lane < 13) { y[lane] = (lane == 12) ? 1.f : 0.f; gy[lane] = A[13 * lane + 12]; }
  __syncwarp();
  int rejects = 0;
  for (int inner = 0; inner < 12; ++inner) {
    if (!(f > 0.f)) break;
    // Jacobian of y w.r.t. (w, tau):  d vec(R)/dw_k = vec(e_k x R(:,j)),  d t/d tau = I
    for (int e = lane; e < 78; e += 32) {
      int row = e / 6, k = e % 6;
      float v = 0.f;
      if (row < 9) {
        if (k < 3) {
          int i = row / 3, j = row % 3;  // entry (i,j) of [e_k]x R : sum_l eps(i,k,l) R(l,j)
          int l1 = (k + 1) % 3, l2 = (k + 2) % 3;  // e_k x v = (.. ) : (e_k x v)_{l2} = v_{l1}, (e_k x v)_{l1} = -v_{l2}
          if (i == l2) v = R[3 * l1 + j];
          else if (i == l1) v = -R[3 * l2 + j];
        }
      } else if (row < 12) {
        v = (k == row - 9 + 3) ? 1.f : 0.f;
      }
      Jm[e] = v;
    }
    __syncwarp();
    for (int e = lane; e < 78; e += 32) {
      int i = e / 6, k = e % 6;
      float s = 0.f;
#pragma unroll
      for (int m = 0; m < 12; ++m) s = fmaf(A[13 * i + m], Jm[6 * m + k], s);
      B[e] = s;
    }
    __syncwarp();
    for (int e = lane; e < 42; e += 32) {
      if (e < 36) {
        int a = e / 6, b = e % 6;
        float s = 0.f;
#pragma unroll
        for (int m = 0; m < 12; ++m) s = fmaf(Jm[6 * m + a], B[6 * m + b], s);
        Hm[e] = s;
      } else {
        int a = e - 36;
        float s = 0.f;
#pragma unroll
        for (int m = 0; m < 12; ++m) s = fmaf(Jm[6 * m + a], gy[m], s);
        gv[a] = s;
      }
    }
    __syncwarp();
    float Hl[36], gl[6], dx[6];
#pragma unroll
    for (int e = 0; e < 36; ++e) Hl[e] = Hm[e];
#pragma unroll
    for (int e = 0; e < 6; ++e) gl[e] = gv[e];
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
    __syncwarp();
    if (lane < 13) {
      float v;
      if (lane < 9) v = Rn[lane] - ((lane == 0 || lane == 4 || lane == 8) ? 1.f : 0.f);
      else if (lane < 12) v = tn[lane - 9];
      else v = 1.f;
      y[lane] = v;
    }
    __syncwarp();
    float gi = 0.f, yi = 0.f;
    if (lane < 13) {
#pragma unroll
      for (int m = 0; m < 13; ++m) gi = fmaf(A[13 * lane + m], y[m], gi);
      yi = y[lane];
    }
    float fn = warp_sum(gi * yi);
    if (fn < f) {
      __syncwarp();
      if (lane < 13) gy[lane] = gi;
#pragma unroll
      for (int e = 0; e < 9; ++e) R[e] = Rn[e];
      t[0] = tn[0]; t[1] = tn[1]; t[2] = tn[2];
      float rel = (f - fn) / f;
      f = fn;
      lambda *= 0.1f;
      if (lambda < 1e-7f) lambda = 0.f;
      __syncwarp();
      if (rel <= ftol) break;
    } else {
      if (++rejects > 8) break;
      lambda = fmaxf(lambda * 10.f, 1e-4f);
      __syncwarp();
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

struct IcpArgs {
  CloudDev scene;
  const float4 *model_nv;
  NNGridDev grid;
  float *poses;      // H x 16, in/out
  int H;
  int max_iter;
  float cos_thr;     // smallest float whose double value exceeds cos(angle)
  float max_d2;
  double abs_mse_eps;
  int *counter;
  int32_t *iters_out, *conv_out;
  int stages;
};

// ------------------------------------------------------------------------------------------------------------
// K4
// ------------------------------------------------------------------------------------------------------------
template <int NW, int TEAM, int SOLVER>
__global__ void __launch_bounds__((NW + 1) * 32, 1) icp_refine_kernel(IcpArgs a) {
  constexpr int NT = NW / TEAM;
  using AccT = Acc<SOLVER>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemLayout S = carve<NW, NT, NACC_PAD, WORK>(smem_raw, a.stages);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool producer = warp == NW;
  const int team = producer ? 0 : warp / TEAM;
  const int tw = warp % TEAM;
  const int n_tiles = a.scene.n_padded / TILE;
  const bool resident = n_tiles <= a.stages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], NW); }
    mbar_fence_init();
  }
  if (!producer && tw == 0 && lane == 0) {
    int h = atomicAdd(a.counter, 1);
    S.hyp[team] = h < a.H ? h : -1;
  }
  __syncthreads();

  int hyp = -1;
  Rigid X, inc_prev;
  int iters = 0;
  double prev_mse = DBL_MAX;
  bool fresh = true;
  uint32_t it = 0;  // running tile counter (ring position / phase)
  int pass = 0;

  for (;; ++pass) {
    const int *hq = S.hyp + (pass & 1) * NT;
    bool any = false;
#pragma unroll
    for (int k = 0; k < NT; ++k) any |= hq[k] >= 0;
    if (!any) break;
    const bool load_now = !resident || pass == 0;

    if (producer) {
      if (lane == 0 && load_now) {
        for (int tI = 0; tI < n_tiles; ++tI, ++it) {
          const int slot = it % a.stages;
          const uint32_t ph = (it / a.stages) & 1u;
          mbar_wait(&S.empty[slot], ph ^ 1u);
          mbar_arrive_expect_tx(&S.full[slot], 2 * TILE_BYTES);
          tma_load_1d(S.tileP + (size_t)slot * TILE, a.scene.pw + (size_t)tI * TILE, TILE_BYTES, &S.full[slot]);
          tma_load_1d(S.tileN + (size_t)slot * TILE, a.scene.nv + (size_t)tI * TILE, TILE_BYTES, &S.full[slot]);
        }
      }
      __syncwarp();
      if (TEAM > 1) __syncthreads();  // (B)
      __syncthreads();                // (A)
      continue;
    }

    // ---- consumer: (re)initialise state when a new hypothesis was assigned ----
    const int h = hq[team];
    if (fresh && h >= 0) {
      Rigid P = rigid_load_colmajor(a.poses + 16 * (size_t)h);
      X = rigid_inverse(P);
#pragma unroll
      for (int e = 0; e < 9; ++e) inc_prev.r[e] = (e % 4 == 0) ? 1.f : 0.f;
      inc_prev.t[0] = inc_prev.t[1] = inc_prev.t[2] = 0.f;
      iters = 0; prev_mse = DBL_MAX; fresh = false;
    }
    hyp = h;

    AccT acc;
    acc.clear();
    for (int tI = 0; tI < n_tiles; ++tI, ++it) {
      const int slot = resident ? tI : (int)(it % a.stages);
      if (load_now) mbar_wait(&S.full[slot], resident ? 0u : ((it / a.stages) & 1u));
      if (hyp >= 0) {
        const float4 *tp = S.tileP + (size_t)slot * TILE;
        const float4 *tn = S.tileN + (size_t)slot * TILE;
#pragma unroll 2
        for (int i = tw * 32 + lane; i < TILE; i += 32 * TEAM) {
          float4 sp = tp[i];
          float3 p = rigid_apply(X, sp.x, sp.y, sp.z);
          float bd; float4 bp;
          int j = nn_query(a.grid, p.x, p.y, p.z, bd, bp);
          if (j >= 0 && bd <= a.max_d2) {
            float4 sn = tn[i];
            float4 mn = __ldg(&a.model_nv[j]);
            float3 ns = rigid_rotate(X, sn.x, sn.y, sn.z);
            float dot = ns.x * mn.x + ns.y * mn.y + ns.z * mn.z;
            if (dot >= a.cos_thr) acc.add(p, make_float3(bp.x, bp.y, bp.z), make_float3(mn.x, mn.y, mn.z), bd);
          }
        }
      }
      if (!resident) { __syncwarp(); if (lane == 0) mbar_arrive(&S.empty[slot]); }
    }

    // ---- reduce: butterfly inside the warp, then across the team through shared memory ----
    float *my_part = S.part + (size_t)warp * NACC_PAD;
#pragma unroll
    for (int k = 0; k < AccT::NACC; ++k) {
      float v = warp_sum(acc.a[k]);
      if (lane == (k & 31)) my_part[k] = v;
    }
    float *W = S.work + (size_t)warp * WORK;
    const float *sums = my_part;
    if (TEAM > 1) {
      __syncthreads();  // (B) all partial rows of the team are written
      // every warp of the team forms the same team totals (identical order -> identical bits) in its own workspace
      float *tot = W + 416;
      for (int k = lane; k < AccT::NACC; k += 32) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < TEAM; ++q) s += S.part[(size_t)(team * TEAM + q) * NACC_PAD + k];
        tot[k] = s;
      }
      sums = tot;
    }
    __syncwarp();

    bool finished = false;
    if (hyp >= 0) {
      const float cnt_f = sums[AccT::NA + 1];
      const float sumd2 = sums[AccT::NA];
      const int cnt = (int)(cnt_f + 0.5f);
      bool converged = false;
      if (cnt < 3) {
        finished = true;  // "Not enough correspondences": hasConverged() false -> identity -> pose unchanged
      } else {
        Rigid inc;
        if (cnt >= 6) {
          if (SOLVER == 0) solve_exact(sums, W, lane, inc.r, inc.t);
          else solve_gn(sums, inc.r, inc.t);
        } else if (cnt >= 4) {
          // Eigen LM: m < n -> ImproperInputParameters, x stays 0 -> identity increment
#pragma unroll
          for (int e = 0; e < 9; ++e) inc.r[e] = (e % 4 == 0) ? 1.f : 0.f;
          inc.t[0] = inc.t[1] = inc.t[2] = 0.f;
        } else {
          inc = inc_prev;  // PCL's LM returns early with < 4 correspondences, transformation_ keeps its old value
        }
        X = rigid_compose(inc, X);
        inc_prev = inc;
        ++iters;
        if (iters >= a.max_iter) converged = true;
        else {
          double cos_angle = 0.5 * ((double)inc.r[0] + (double)inc.r[4] + (double)inc.r[8] - 1.0);
          double tsq = (double)inc.t[0] * inc.t[0] + (double)inc.t[1] * inc.t[1] + (double)inc.t[2] * inc.t[2];
          if (cos_angle >= 1.0 && tsq <= 0.0) converged = true;
          else {
            double mse = (double)sumd2 / (double)cnt;
            if (fabs(mse - prev_mse) < a.abs_mse_eps) converged = true;
            prev_mse = mse;
          }
        }
        finished = converged;
      }
      if (finished && tw == 0 && lane == 0) {
        if (converged) {
          Rigid P = rigid_inverse(X);
          rigid_store_colmajor(P, a.poses + 16 * (size_t)hyp);
        }
        if (a.iters_out) a.iters_out[hyp] = iters;
        if (a.conv_out) a.conv_out[hyp] = converged ? 1 : 0;
      }
    }
    // next assignment for this team (written to the other half of the double-buffered table)
    if (tw == 0 && lane == 0) {
      int nxt = hyp;
      if (hyp < 0) nxt = -1;
      else if (finished) { int q = atomicAdd(a.counter, 1); nxt = q < a.H ? q : -1; }
      S.hyp[((pass + 1) & 1) * NT + team] = nxt;
    }
    if (finished) fresh = true;
    __syncthreads();  // (A)
  }
}

// ------------------------------------------------------------------------------------------------------------
// K5
// ------------------------------------------------------------------------------------------------------------
struct LcpArgs {
  CloudDev scene;
  const float4 *model_nv;
  NNGridDev mgrid;   // model grid (radius >= dist)
  NNGridDev sgrid;   // scene grid (radius >= dist), for the reciprocal term
  const float *poses;
  int H;
  float dist, inv_dist, dist2, cos_thr;
  int use_normal, use_dot, use_recip, use_weights;
  int *counter;
  float *scores;
  int stages;
};

template <int NW, int TEAM>
__global__ void __launch_bounds__((NW + 1) * 32, 1) lcp_score_kernel(LcpArgs a) {
  constexpr int NT = NW / TEAM;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemLayout S = carve<NW, NT, 8, 8>(smem_raw, a.stages);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool producer = warp == NW;
  const int team = producer ? 0 : warp / TEAM;
  const int tw = warp % TEAM;
  const int n_tiles = a.scene.n_padded / TILE;
  const bool resident = n_tiles <= a.stages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], NW); }
    mbar_fence_init();
  }
  if (!producer && tw == 0 && lane == 0) {
    int h = atomicAdd(a.counter, 1);
    S.hyp[team] = h < a.H ? h : -1;
  }
  __syncthreads();

  uint32_t it = 0;
  for (int pass = 0;; ++pass) {
    const int *hq = S.hyp + (pass & 1) * NT;
    bool any = false;
#pragma unroll
    for (int k = 0; k < NT; ++k) any |= hq[k] >= 0;
    if (!any) break;
    const bool load_now = !resident || pass == 0;
    if (producer) {
      if (lane == 0 && load_now) {
        for (int tI = 0; tI < n_tiles; ++tI, ++it) {
          const int slot = it % a.stages;
          const uint32_t ph = (it / a.stages) & 1u;
          mbar_wait(&S.empty[slot], ph ^ 1u);
          mbar_arrive_expect_tx(&S.full[slot], 2 * TILE_BYTES);
          tma_load_1d(S.tileP + (size_t)slot * TILE, a.scene.pw + (size_t)tI * TILE, TILE_BYTES, &S.full[slot]);
          tma_load_1d(S.tileN + (size_t)slot * TILE, a.scene.nv + (size_t)tI * TILE, TILE_BYTES, &S.full[slot]);
        }
      }
      __syncwarp();
      if (TEAM > 1) __syncthreads();
      __syncthreads();
      continue;
    }
    const int hyp = hq[team];
    Rigid T, Ti;
    if (hyp >= 0) { T = rigid_load_colmajor(a.poses + 16 * (size_t)hyp); Ti = rigid_inverse(T); }
    float score = 0.f;
    for (int tI = 0; tI < n_tiles; ++tI, ++it) {
      const int slot = resident ? tI : (int)(it % a.stages);
      if (load_now) mbar_wait(&S.full[slot], resident ? 0u : ((it / a.stages) & 1u));
      if (hyp >= 0) {
        const float4 *tp = S.tileP + (size_t)slot * TILE;
        const float4 *tn = S.tileN + (size_t)slot * TILE;
#pragma unroll 2
        for (int i = tw * 32 + lane; i < TILE; i += 32 * TEAM) {
          float4 sp = tp[i];
          float3 p = rigid_apply(Ti, sp.x, sp.y, sp.z);   // scene point in the model frame
          float bd; float4 bp;
          int j = nn_query(a.mgrid, p.x, p.y, p.z, bd, bp);
          if (j >= 0 && bd < a.dist2) {                    // Utils.cpp:388 (strict)
            const float w = a.use_weights ? sp.w : 1.f;
            float4 mn = __ldg(&a.model_nv[j]);
            // transformed model normal, normalised (rotation keeps the norm: use the stored 1/|n|)
            float3 mr = rigid_rotate(T, mn.x * mn.w, mn.y * mn.w, mn.z * mn.w);
            if (!a.use_normal) score += w;
            else {
              float4 sn = tn[i];
              float dot = (sn.x * mr.x + sn.y * mr.y + sn.z * mr.z) * sn.w;
              if (dot > a.cos_thr) score += a.use_dot ? dot * (1.f - sqrtf(bd) * a.inv_dist) * w : w;
            }
            if (a.use_recip) {
              // nearest scene point of the (transformed) model neighbour; it lies within `dist` because scene
              // point i itself does, so the radius-limited scene grid is exact here
              float3 q = rigid_apply(T, bp.x, bp.y, bp.z);
              float ed; float4 ep;
              int k = nn_query(a.sgrid, q.x, q.y, q.z, ed, ep);
              if (k >= 0) {
                if (!a.use_normal) score += w;
                else {
                  float4 s2 = __ldg(&a.scene.nv[k]);
                  float dot = (s2.x * mr.x + s2.y * mr.y + s2.z * mr.z) * s2.w;
                  if (dot > a.cos_thr) score += a.use_dot ? dot * (1.f - sqrtf(ed) * a.inv_dist) * w : w;
                }
              }
            }
          }
        }
      }
      if (!resident) { __syncwarp(); if (lane == 0) mbar_arrive(&S.empty[slot]); }
    }
    score = warp_sum(score);
    if (TEAM > 1) {
      if (lane == 0) S.part[warp * 8] = score;
      __syncthreads();
      if (tw == 0 && lane == 0) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < TEAM; ++q) s += S.part[(team * TEAM + q) * 8];
        score = s;
      }
    }
    if (tw == 0 && lane == 0) {
      int nxt = -1;
      if (hyp >= 0) {
        a.scores[hyp] = score;
        int q = atomicAdd(a.counter, 1);
        nxt = q < a.H ? q : -1;
      }
      S.hyp[((pass + 1) & 1) * NT + team] = nxt;
    }
    __syncthreads();
  }
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

static int pick_stages(int n_tiles, size_t fixed_bytes, size_t limit) {
  // whole scene resident when it fits, else a 4-deep ring
  size_t avail = limit > fixed_bytes ? limit - fixed_bytes : 0;
  int max_stages = (int)(avail / (TILE * 32 + 16));
  if (n_tiles <= max_stages) return n_tiles < 1 ? 1 : n_tiles;
  return max_stages < 4 ? max_stages : 4;
}

template <int NW, int TEAM, int SOLVER>
static int launch_icp_t(hop_ctx *ctx, IcpArgs &a, int grid) {
  constexpr int NT = NW / TEAM;
  const int n_tiles = a.scene.n_padded / TILE;
  const size_t fixed = smem_bytes(0, NW, NT, NACC_PAD, WORK) + 256;
  a.stages = pick_stages(n_tiles, fixed, 200 * 1024);
  if (a.stages < 1) { ctx->err = "icp: no shared memory for tiles"; return HOP_EINVAL; }
  const size_t smem = smem_bytes(a.stages, NW, NT, NACC_PAD, WORK) + 128;
  HOP_CUDA(ctx, cudaFuncSetAttribute(icp_refine_kernel<NW, TEAM, SOLVER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  icp_refine_kernel<NW, TEAM, SOLVER><<<grid, (NW + 1) * 32, smem, ctx->stream>>>(a);
  ctx->launches += 1;
  HOP_CUDA(ctx, cudaGetLastError());
  return HOP_OK;
}

template <int NW, int TEAM>
static int launch_lcp_t(hop_ctx *ctx, LcpArgs &a, int grid) {
  constexpr int NT = NW / TEAM;
  const int n_tiles = a.scene.n_padded / TILE;
  const size_t fixed = smem_bytes(0, NW, NT, 8, 8) + 256;
  a.stages = pick_stages(n_tiles, fixed, 200 * 1024);
  const size_t smem = smem_bytes(a.stages, NW, NT, 8, 8) + 128;
  HOP_CUDA(ctx, cudaFuncSetAttribute(lcp_score_kernel<NW, TEAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lcp_score_kernel<NW, TEAM><<<grid, (NW + 1) * 32, smem, ctx->stream>>>(a);
  ctx->launches += 1;
  HOP_CUDA(ctx, cudaGetLastError());
  return HOP_OK;
}

}  // namespace

constexpr int ICP_NW = 8;
constexpr int LCP_NW = 16;

int hop_launch_icp(hop_ctx *ctx, const CloudDev &scene, const CloudDev &model, const NNGridDev &grid, float *d_poses, int H,
                   const hop_icp_params &p, int32_t *d_iters, int32_t *d_conv) {
  if (H <= 0) return HOP_OK;
  if (p.mode != 0) { ctx->err = "hop_icp_refine: mode 1 (point-to-point) not built yet"; return HOP_EINVAL; }
  IcpArgs a;
  a.scene = scene; a.model_nv = model.nv; a.grid = grid; a.poses = d_poses; a.H = H;
  a.max_iter = p.max_iter < 1 ? 1 : p.max_iter;
  a.cos_thr = float_above(cos((double)p.angle_deg / 180.0 * M_PI));
  a.max_d2 = p.max_dist * p.max_dist;
  a.abs_mse_eps = p.abs_mse_eps;
  a.counter = ctx->d_counter;
  a.iters_out = d_iters; a.conv_out = d_conv; a.stages = 0;
  HOP_CUDA(ctx, cudaMemsetAsync(ctx->d_counter, 0, sizeof(int), ctx->stream));
  const int team = pick_team(H, ctx->sm_count, ICP_NW, p.team_warps);
  const int teams_per_cta = ICP_NW / team;
  int grid_dim = (H + teams_per_cta - 1) / teams_per_cta;
  if (grid_dim > ctx->sm_count) grid_dim = ctx->sm_count;
#define HOP_ICP_CASE(T)                                                                   \
  case T:                                                                                 \
    return p.solver == 1 ? launch_icp_t<ICP_NW, T, 1>(ctx, a, grid_dim) : launch_icp_t<ICP_NW, T, 0>(ctx, a, grid_dim);
  switch (team) {
    HOP_ICP_CASE(1)
    HOP_ICP_CASE(2)
    HOP_ICP_CASE(4)
    HOP_ICP_CASE(8)
  }
#undef HOP_ICP_CASE
  return HOP_EINVAL;
}

int hop_launch_lcp(hop_ctx *ctx, const CloudDev &scene, const CloudDev &model, const NNGridDev &model_grid,
                   const NNGridDev &scene_grid, const float *d_poses, int H, const hop_lcp_params &p, int use_weights,
                   float *d_scores) {
  if (H <= 0) return HOP_OK;
  LcpArgs a;
  a.scene = scene; a.model_nv = model.nv; a.mgrid = model_grid; a.sgrid = scene_grid; a.poses = d_poses; a.H = H;
  a.dist = p.dist; a.inv_dist = 1.f / p.dist; a.dist2 = p.dist * p.dist;
  a.cos_thr = (float)cos((double)p.angle_deg / 180.0 * M_PI);
  a.use_normal = p.use_normal; a.use_dot = p.use_dot_score; a.use_recip = p.use_reciprocal; a.use_weights = use_weights;
  a.counter = ctx->d_counter + 1; a.scores = d_scores; a.stages = 0;
  HOP_CUDA(ctx, cudaMemsetAsync(ctx->d_counter + 1, 0, sizeof(int), ctx->stream));
  const int team = pick_team(H, ctx->sm_count, LCP_NW, p.team_warps);
  const int teams_per_cta = LCP_NW / team;
  int grid_dim = (H + teams_per_cta - 1) / teams_per_cta;
  if (grid_dim > ctx->sm_count) grid_dim = ctx->sm_count;
  switch (team) {
    case 1: return launch_lcp_t<LCP_NW, 1>(ctx, a, grid_dim);
    case 2: return launch_lcp_t<LCP_NW, 2>(ctx, a, grid_dim);
    case 4: return launch_lcp_t<LCP_NW, 4>(ctx, a, grid_dim);
    case 8: return launch_lcp_t<LCP_NW, 8>(ctx, a, grid_dim);
  }
  return HOP_EINVAL;
}
