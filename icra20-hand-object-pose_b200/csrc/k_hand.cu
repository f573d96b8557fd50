// k_hand.cu -- K1: the hand-state overlap objective over a dense grid of joint angles, for sm_100a.
//
//   replaces objFuncPSO                      (src/perception/src/Hand.cpp:10-178)
//            the swarm that evaluates it      (src/perception/include/unconstrained/pso.hpp:146-351: 16 particles x 4 evaluations,
//                                              each OpenMP thread deep-copying the clouds AND the FLANN tree, pso.hpp:83-108,232,286)
//
// Two launches.  hand_match_kernel: flat over (state, finger point) -- the finger cloud is carried into the hand-base frame by
// the state's transform and looked up in ONE exact nearest-neighbour grid of the scene shared by every state; the match
// count of a state is an integer sum, so the order of the atomics does not matter.  (The first version walked the finger
// points inside the state's warp: ten dependent gather rounds on the critical path of every state that passes the gap
// test, 95 us for 4096 states; flat, the same queries fill the machine.)  hand_overlap_kernel: one warp per hand state, 8
// states per CTA; the no-swivel scene is staged into shared memory by 1-D bulk (TMA) copies, double buffered on
// mbarriers, and every warp of the CTA reads it from there, moving it into ITS finger frame.  Integer results (match count, outer
// count) come from ballots; the float sums use a fixed shuffle order, so costs are deterministic run to run.
// The arithmetic follows the reference's types and operation order (see oracle/hop_oracle_hand.c): unfused float
// transforms (PCL 1.9), `num_match += 1 + X[0]` in double rounded to float per match, double penalties.
#include "hop_common.cuh"

namespace {

constexpr int HAND_WARPS = 8;
constexpr int HAND_CHUNK = 1024;  // no-swivel scene points per shared-memory stage (16 KB; two stages: 5 CTAs per SM)

struct HandArgs {
  hop_finger_params p;
  const float4 *f_pw, *f_nv; int nf;          // finger cloud
  NNGridDev grid;                              // scene_hand (radius >= dist_thres)
  const float4 *lk_nv; int n_lk;               // normals read with the neighbour's index
  const float4 *w_pw; int nw, nw_padded;       // no-swivel scene
  const double *thetas; const float *half_cs; int S;
  float cos_thr, thr2;
  double *cost;
  int *matches;                                // per state: finger points with an accepted neighbour (hand_match_kernel)
};

struct M4f { float m[16]; };  // column-major

__device__ __forceinline__ float m4(const float *m, int r, int c) { return m[c * 4 + r]; }

// Eigen fixed 4x4 float product, SSE packet order: ((a0 b0 + a1 b1) + a2 b2) + a3 b3, unfused
__device__ __forceinline__ void m4_mul(const float *A, const float *B, float *C) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int r = 0; r < 4; ++r)
      C[j * 4 + r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m4(A, r, 0), m4(B, 0, j)), __fmul_rn(m4(A, r, 1), m4(B, 1, j))),
                                         __fmul_rn(m4(A, r, 2), m4(B, 2, j))),
                               __fmul_rn(m4(A, r, 3), m4(B, 3, j)));
}
// row r of M * (x,y,z,1): PCL 1.9 transformPointCloudWithNormals order, unfused
__device__ __forceinline__ float row_pt(const float *M, int r, float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m4(M, r, 0), x), __fmul_rn(m4(M, r, 1), y)), __fmul_rn(m4(M, r, 2), z)), m4(M, r, 3));
}
__device__ __forceinline__ float row_vec(const float *M, int r, float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(m4(M, r, 0), x), __fmul_rn(m4(M, r, 1), y)), __fmul_rn(m4(M, r, 2), z));
}
// rows 1 and 2 (y, z) of the inverse of a rigid-up-to-rounding transform: adjugate in double, rounded once
__device__ __forceinline__ void inverse_rows_yz(const float *T, float *ry, float *rz) {
  double a[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) a[3 * r + c] = (double)m4(T, r, c);
  auto mm = [](double x, double y, double z, double w) { return __dsub_rn(__dmul_rn(x, y), __dmul_rn(z, w)); };
  const double c00 = mm(a[4], a[8], a[5], a[7]), c01 = mm(a[5], a[6], a[3], a[8]), c02 = mm(a[3], a[7], a[4], a[6]);
  const double det = __dadd_rn(__dadd_rn(__dmul_rn(a[0], c00), __dmul_rn(a[1], c01)), __dmul_rn(a[2], c02));
  const double id = __ddiv_rn(1.0, det);
  double I[6];
  I[0] = __dmul_rn(c01, id); I[1] = __dmul_rn(mm(a[0], a[8], a[2], a[6]), id); I[2] = __dmul_rn(mm(a[2], a[3], a[0], a[5]), id);
  I[3] = __dmul_rn(c02, id); I[4] = __dmul_rn(mm(a[1], a[6], a[0], a[7]), id); I[5] = __dmul_rn(mm(a[0], a[4], a[1], a[3]), id);
  const double t0 = (double)m4(T, 0, 3), t1 = (double)m4(T, 1, 3), t2 = (double)m4(T, 2, 3);
#pragma unroll
  for (int c = 0; c < 3; ++c) { ry[c] = __double2float_rn(I[c]); rz[c] = __double2float_rn(I[3 + c]); }
  ry[3] = __double2float_rn(-__dadd_rn(__dadd_rn(__dmul_rn(I[0], t0), __dmul_rn(I[1], t1)), __dmul_rn(I[2], t2)));
  rz[3] = __double2float_rn(-__dadd_rn(__dadd_rn(__dmul_rn(I[3], t0), __dmul_rn(I[4], t1)), __dmul_rn(I[5], t2)));
}

// the state's finger-link transform in the hand-base frame and the gripper-gap test (Hand.cpp:15-64).
// branch: 0 gap penalty (score is final), 4 none yet
__device__ __forceinline__ void hand_state_setup(const HandArgs &a, int s, bool valid, double &X, float *cur, int &branch, float &score) {
  X = valid ? a.thetas[s] : 0.0;
  float cw, sx;
  if (a.half_cs) { cw = valid ? a.half_cs[2 * s] : 1.f; sx = valid ? a.half_cs[2 * s + 1] : 0.f; }
  else { const float h = __fmul_rn(0.5f, (float)X); cw = cosf(h); sx = sinf(h); }
  {
    // tf_self = Quaternionf(w = cos(a/2), x = sin(a/2)).toRotationMatrix() (Hand.cpp:15-21)
    const float tx = __fmul_rn(2.f, sx), twx = __fmul_rn(tx, cw), txx = __fmul_rn(tx, sx);
    float tf[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) tf[k] = 0.f;
    tf[0] = 1.f; tf[5] = __fsub_rn(1.f, txx); tf[9] = __fsub_rn(0.f, twx); tf[6] = __fadd_rn(0.f, twx); tf[10] = __fsub_rn(1.f, txx); tf[15] = 1.f;
    m4_mul(a.p.model2handbase, tf, cur);
  }
  branch = 4;
  score = 0.f;
  {
    float tip1y;
    if (a.p.palm_side) {
      float o2h[16];
      m4_mul(cur, a.p.finger_out2parent, o2h);
      tip1y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m4(o2h, 1, 0), a.p.tip1_local[0]), __fmul_rn(m4(o2h, 1, 1), a.p.tip1_local[1])),
                                  __fmul_rn(m4(o2h, 1, 2), a.p.tip1_local[2])), __fmul_rn(m4(o2h, 1, 3), 1.f));
    } else {
      tip1y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m4(cur, 1, 0), a.p.tip1_local[0]), __fmul_rn(m4(cur, 1, 1), a.p.tip1_local[1])),
                                  __fmul_rn(m4(cur, 1, 2), a.p.tip1_local[2])), __fmul_rn(m4(cur, 1, 3), 1.f));
    }
    const float tip2y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m4(cur, 1, 0), a.p.tip2_local[0]), __fmul_rn(m4(cur, 1, 1), a.p.tip2_local[1])),
                                            __fmul_rn(m4(cur, 1, 2), a.p.tip2_local[2])), __fmul_rn(m4(cur, 1, 3), 1.f));
    float gd1, gd2;
    if (a.p.right_side) { gd1 = __fsub_rn(tip1y, a.p.pair_tip1_y); gd2 = __fsub_rn(tip2y, a.p.pair_tip2_y); }
    else { gd1 = __fadd_rn(-tip1y, a.p.pair_tip1_y); gd2 = __fadd_rn(-tip2y, a.p.pair_tip2_y); }
    const float G = a.p.gripper_min_dist;
    if (gd1 < G || gd2 < G) {
      const float pen = __double2float_rn(__dadd_rn(1e3, __dmul_rn(1e3, (double)fabsf(__fsub_rn(G, gd1)))));
      score = __fsub_rn(0.f, pen);
      branch = 0;
    }
  }
}

constexpr int MATCH_THREADS = 128;

// matches[s] = number of finger points whose nearest scene neighbour passes the distance and normal gates (Hand.cpp:84-125)
__global__ void __launch_bounds__(MATCH_THREADS) hand_match_kernel(const __grid_constant__ HandArgs a) {
  const int s = blockIdx.x;
  double X; float cur[16]; int branch; float score;
  hand_state_setup(a, s, true, X, cur, branch, score);
  if (branch == 0) return;   // (uniform across the CTA)
  const int i = blockIdx.y * MATCH_THREADS + threadIdx.x;
  bool hit = false;
  if (i < a.nf) {
    const float4 fp = __ldg(&a.f_pw[i]);
    const float px = row_pt(cur, 0, fp.x, fp.y, fp.z), py = row_pt(cur, 1, fp.x, fp.y, fp.z), pz = row_pt(cur, 2, fp.x, fp.y, fp.z);
    float bd; float4 bp;
    const int j = nn_query(a.grid, px, py, pz, bd, bp);
    if (j >= 0 && bd <= a.thr2) {
      if (!a.p.check_normal) hit = true;
      else if (j < a.n_lk) {
        const float4 n2 = __ldg(&a.lk_nv[j]);
        if (n2.x == 0.f && n2.y == 0.f && n2.z == 0.f) hit = true;
        else if (isfinite(n2.x) && isfinite(n2.y) && isfinite(n2.z)) {
          const float4 fn = __ldg(&a.f_nv[i]);
          const float n1x = row_vec(cur, 0, fn.x, fn.y, fn.z), n1y = row_vec(cur, 1, fn.x, fn.y, fn.z), n1z = row_vec(cur, 2, fn.x, fn.y, fn.z);
          const float dot = __fadd_rn(__fmul_rn(n1x, n2.x), __fadd_rn(__fmul_rn(n1y, n2.y), __fmul_rn(n1z, n2.z)));
          hit = dot >= a.cos_thr;
        }
      }
    }
  }
  const int cnt = __syncthreads_count(hit);
  if (threadIdx.x == 0 && cnt) atomicAdd(&a.matches[s], cnt);
}

__global__ void __launch_bounds__(HAND_WARPS * 32) hand_overlap_kernel(const __grid_constant__ HandArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4 *const stage0 = reinterpret_cast<float4 *>(smem_raw);  // two stages of HAND_CHUNK points
  __shared__ __align__(8) uint64_t full[2];
  __shared__ float s_hist[HOP_MAX_FINGER_BINS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_chunks = (a.nw_padded + HAND_CHUNK - 1) / HAND_CHUNK;
  if (threadIdx.x == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init(); }
  if (threadIdx.x < HOP_MAX_FINGER_BINS) s_hist[threadIdx.x] = a.p.hist_min_y[threadIdx.x];
  __syncthreads();
  auto issue = [&](int c) {  // one elected thread: bulk copy of chunk c into its stage
    const int cnt = min(HAND_CHUNK, a.nw_padded - c * HAND_CHUNK);
    mbar_arrive_expect_tx(&full[c & 1], (uint32_t)cnt * 16u);
    tma_load_1d(stage0 + (c & 1) * HAND_CHUNK, a.w_pw + (size_t)c * HAND_CHUNK, (uint32_t)cnt * 16u, &full[c & 1]);
  };
  if (threadIdx.x == 0 && n_chunks > 0) issue(0);

  // ---- per-state set-up (uniform across the warp) ----
  const int s = blockIdx.x * HAND_WARPS + warp;
  const bool valid = s < a.S;
  double X;
  float cur[16];
  int branch;  // 0 gap, 1 no match, 2 hard outer, 3 exp, 4 none
  float score;
  hand_state_setup(a, s, valid, X, cur, branch, score);

  // ---- matches: counted by hand_match_kernel ----
  if (valid && branch != 0) {
    const int matches = a.matches[s];
    // num_match += 1 + X[0]  (float accumulator, double increment), once per match
    float nm = 0.f;
    const double inc = __dadd_rn(1.0, X);
    for (int k = 0; k < matches; ++k) nm = __double2float_rn(__dadd_rn((double)nm, inc));
    score = nm;
    if (nm == 0.f) { score = __double2float_rn(__dadd_rn(-100.0, X)); branch = 1; }
  }

  // ---- outer-point penalty: the no-swivel scene, staged through shared memory, in this state's finger frame ----
  const bool need_outer = valid && branch == 4;
  float ry[4] = {0, 0, 0, 0}, rz[4] = {0, 0, 0, 0};
  if (need_outer) inverse_rows_yz(cur, ry, rz);
  float osum = 0.f;
  int ocnt = 0;
  const float min_z = a.p.min_z, stride_z = a.p.stride_z;
  const int last_bin = a.p.num_division - 1;
  for (int c = 0; c < n_chunks; ++c) {
    if (threadIdx.x == 0 && c + 1 < n_chunks) issue(c + 1);  // its stage was released by the __syncthreads of chunk c-1
    mbar_wait(&full[c & 1], (uint32_t)((c >> 1) & 1));
    if (need_outer) {
      const float4 *pts = stage0 + (c & 1) * HAND_CHUNK;
      const int cnt = min(HAND_CHUNK, a.nw - c * HAND_CHUNK);  // real points only (the tail of the cloud is padding)
      for (int i = lane; i < cnt; i += 32) {
        const float4 q = pts[i];
        const float py = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(ry[0], q.x), __fmul_rn(ry[1], q.y)), __fmul_rn(ry[2], q.z)), ry[3]);
        const float pz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(rz[0], q.x), __fmul_rn(rz[1], q.y)), __fmul_rn(rz[2], q.z)), rz[3]);
        int bin = __float2int_rz(__fdiv_rn(fmaxf(__fsub_rn(pz, min_z), 0.0f), stride_z));  // FingerProperty::getBinAlongZ
        bin = min(max(bin, 0), last_bin);
        const float face = s_hist[bin];
        if (py < face) { osum = __fadd_rn(osum, fabsf(__fsub_rn(py, face))); ++ocnt; }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { osum = __fadd_rn(osum, __shfl_xor_sync(0xffffffffu, osum, o)); ocnt += __shfl_xor_sync(0xffffffffu, ocnt, o); }

  if (valid && lane == 0) {
    if (branch == 4) {
      const float avg = __fdiv_rn(osum, (float)ocnt);  // 0/0 = NaN: no penalty
      const float d = a.p.outter_pt_dist, w = a.p.outter_pt_dist_weight;
      if (ocnt >= a.p.max_outter_pts || (double)avg >= 0.005) {
        const float pen = __double2float_rn(__dadd_rn(1e3, (double)__fmul_rn(w, fmaxf(__fsub_rn(avg, d), 0.0f))));
        score = __fsub_rn(score, pen);
      } else if (__fsub_rn(avg, d) > 0.f) {
        const float pen = __fmul_rn(w, expf(__fmul_rn(avg, 1000.f)));
        score = __fsub_rn(score, pen);
      }
    }
    a.cost[s] = -(double)score;
  }
}

// arg-min of S doubles, ties -> lowest index, NaN never wins
__global__ void __launch_bounds__(1024, 1) argmin_kernel(const double *__restrict__ cost, int S, int32_t *__restrict__ best) {
  __shared__ double s_v[32];
  __shared__ int s_i[32];
  double bv = INFINITY;
  int bi = -1;
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    double v = cost[i];
    if (!(v == v)) v = INFINITY;
    if (bi < 0 || v < bv) { bv = v; bi = i; }
  }
  auto better = [](double ov, int oi, double v, int i) { return oi >= 0 && (i < 0 || ov < v || (ov == v && oi < i)); };
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = bv; s_i[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    bv = s_v[threadIdx.x]; bi = s_i[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (threadIdx.x == 0) *best = bi;
  }
}

}  // namespace

extern "C" int hop_hand_overlap_dev(hop_ctx *ctx, hop_cloud *finger, hop_cloud *scene_hand, hop_cloud *scene_normals,
                                    hop_cloud *scene_noswivel, const hop_finger_params *params, const double *d_thetas,
                                    const float *d_half_cs, int S, double *d_cost, int32_t *d_best) {
  HOP_ENTER(ctx);
  if (!ctx || !finger || !scene_hand || !scene_noswivel || !params || S < 0) { if (ctx) ctx->err = "hop_hand_overlap: bad arguments"; return HOP_EINVAL; }
  if (S == 0) return HOP_OK;
  if (!d_thetas || !d_cost) { ctx->err = "hop_hand_overlap: null state/cost buffer"; return HOP_EINVAL; }
  if (params->num_division < 1 || params->num_division > HOP_MAX_FINGER_BINS) { ctx->err = "hop_hand_overlap: num_division out of range"; return HOP_EINVAL; }
  if (!(params->dist_thres > 0.f)) { ctx->err = "hop_hand_overlap: dist_thres must be > 0"; return HOP_EINVAL; }
  // the reference asserts a non-empty scene (Hand.cpp:322); an empty finger cloud or scene simply never matches
  HandArgs a;
  a.p = *params;
  a.f_pw = finger->d_pw; a.f_nv = finger->d_nv; a.nf = finger->n;
  if (scene_hand->n > 0) {
    NNGridHost *G = nullptr;
    int rc = hop_get_nn_grid(ctx, scene_hand, params->dist_thres, 0.f, &G);
    if (rc != HOP_OK) return rc;
    a.grid = G->dev;
  } else {
    a.grid = NNGridDev{0.f, 0.f, 0.f, 1.f, 0, 0, 0, params->dist_thres, nullptr, nullptr};  // every query falls outside
  }
  hop_cloud *lk = scene_normals ? scene_normals : scene_hand;
  a.lk_nv = lk->d_nv; a.n_lk = lk->n;
  a.w_pw = scene_noswivel->d_pw; a.nw = scene_noswivel->n; a.nw_padded = scene_noswivel->n > 0 ? scene_noswivel->n_padded : 0;
  a.thetas = d_thetas; a.half_cs = d_half_cs; a.S = S;
  a.cos_thr = (float)std::cos((double)params->normal_angle_deg / 180.0 * M_PI);
  a.thr2 = params->dist_thres * params->dist_thres;
  a.cost = d_cost;
  const size_t smem = 2 * (size_t)HAND_CHUNK * sizeof(float4);
  HOP_CUDA(ctx, ctx->func_smem_optin(hand_overlap_kernel, smem));
  a.matches = (int *)ctx->ensure_work(sizeof(int) * (size_t)S);
  if (!a.matches) { ctx->err = "hop_hand_overlap: work buffer allocation failed"; return HOP_ENOMEM; }
  {
    ProfScope ps(ctx, HOP_PROF_HAND);
    HOP_CUDA(ctx, cudaMemsetAsync(a.matches, 0, sizeof(int) * (size_t)S, ctx->stream));
    if (a.nf > 0) {
      hand_match_kernel<<<dim3(S, (a.nf + MATCH_THREADS - 1) / MATCH_THREADS), MATCH_THREADS, 0, ctx->stream>>>(a);
      ctx->launches += 1;
    }
    hand_overlap_kernel<<<(S + HAND_WARPS - 1) / HAND_WARPS, HAND_WARPS * 32, smem, ctx->stream>>>(a);
    ctx->launches += 1;
    if (d_best) { argmin_kernel<<<1, 1024, 0, ctx->stream>>>(d_cost, S, d_best); ctx->launches += 1; }
  }
  HOP_CUDA(ctx, cudaGetLastError());
  return HOP_OK;
}

extern "C" int hop_hand_overlap(hop_ctx *ctx, hop_cloud *finger, hop_cloud *scene_hand, hop_cloud *scene_normals, hop_cloud *scene_noswivel,
                                const hop_finger_params *params, const double *thetas, int S, double *cost_out, int32_t *best_out) {
  HOP_ENTER(ctx);
  if (!ctx || S < 0 || (S > 0 && (!thetas || !cost_out))) return HOP_EINVAL;
  if (best_out) *best_out = -1;
  if (S == 0) return HOP_OK;
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t tb = up(sizeof(double) * (size_t)S), hb = up(sizeof(float) * 2 * (size_t)S);
  char *d = (char *)ctx->ensure_io(2 * tb + hb + 256);
  if (!d) { ctx->err = "hop_hand_overlap: staging allocation failed"; return HOP_ENOMEM; }
  double *d_thetas = (double *)d, *d_cost = (double *)(d + tb);
  float *d_half = (float *)(d + 2 * tb);
  int32_t *d_best = (int32_t *)(d + 2 * tb + hb);
  // the half-angle sine/cosine come from the host's libm, like Eigen's AngleAxisf -> Quaternionf on the reference's host
  char *h = (char *)ctx->ensure_pinned(tb + hb);
  if (!h) { ctx->err = "hop_hand_overlap: pinned staging allocation failed"; return HOP_ENOMEM; }
  HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  double *h_thetas = (double *)h; float *h_half = (float *)(h + tb);
  for (int s = 0; s < S; ++s) {
    h_thetas[s] = thetas[s];
    const float half = 0.5f * (float)thetas[s];
    h_half[2 * s] = std::cos(half); h_half[2 * s + 1] = std::sin(half);
  }
  HOP_CUDA(ctx, cudaMemcpyAsync(d_thetas, h_thetas, sizeof(double) * (size_t)S, cudaMemcpyHostToDevice, ctx->stream));
  HOP_CUDA(ctx, cudaMemcpyAsync(d_half, h_half, sizeof(float) * 2 * (size_t)S, cudaMemcpyHostToDevice, ctx->stream));
  int rc = hop_hand_overlap_dev(ctx, finger, scene_hand, scene_normals, scene_noswivel, params, d_thetas, d_half, S, d_cost,
                                best_out ? d_best : nullptr);
  if (rc != HOP_OK) return rc;
  HOP_CUDA(ctx, cudaMemcpyAsync(cost_out, d_cost, sizeof(double) * (size_t)S, cudaMemcpyDeviceToHost, ctx->stream));
  if (best_out) HOP_CUDA(ctx, cudaMemcpyAsync(best_out, d_best, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HOP_OK;
}

// ---- HandT42::adjustHandHeight (Hand.cpp:999-1051; main_realdata_auto.cpp:141) ---------------------------------------------------
// The 13 trial heights are 13 more "hand states": one thread per (height, hand point) finds the point's exact nearest scene
// neighbour through the scene's grid, counts it when it lies within 5 mm and the (un-normalised, Eigen-ordered) normal dot
// product reaches cos 45 deg; the host keeps the first height with the most matches.
namespace {
struct HeightArgs {
  const float4 *h_pw, *h_nv; int nh;      // hand->_hand_cloud, hand-base frame
  NNGridDev grid; const float4 *s_nv;     // the hand-region scene in the hand-base frame
  const float *heights; int n_heights;
  double cos_thr;                         // std::cos(45 / 180.0 * M_PI) from the host's libm, compared in double like the reference
  int *counts;
};
__global__ void __launch_bounds__(128) hand_height_kernel(HeightArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
  if (i >= a.nh) return;
  const float4 p = a.h_pw[i];
  const float z = __fadd_rn(p.z, a.heights[t]);   // pcl::transformPointCloudWithNormals with offset(2,3) = height: x, y and the normal are unchanged
  float bd; float4 bp;
  const int j = nn_query(a.grid, p.x, p.y, z, bd, bp);
  if (j < 0 || bd > (float)(0.005 * 0.005)) return;
  const float4 n = a.h_nv[i], m = a.s_nv[j];
  const float dot = __fadd_rn(__fmul_rn(n.x, m.x), __fadd_rn(__fmul_rn(n.y, m.y), __fmul_rn(n.z, m.z)));   // Eigen: x0 y0 + (x1 y1 + x2 y2)
  if ((double)dot >= a.cos_thr) atomicAdd(a.counts + t, 1);
}
}  // namespace

extern "C" int hop_adjust_hand_height(hop_ctx *ctx, hop_cloud *hand_cloud, hop_cloud *scene_handbase, const float *heights, int n_heights,
                                      int32_t *match_counts, int32_t *best_index) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  if (!hand_cloud || !scene_handbase || !heights || n_heights < 1 || n_heights > 4096 || !best_index) { ctx->err = "hop_adjust_hand_height: bad arguments"; return HOP_EINVAL; }
  std::vector<int32_t> counts(n_heights, 0);
  if (hand_cloud->n > 0 && scene_handbase->n > 0) {
    NNGridHost *G = nullptr;
    int rc = hop_get_nn_grid(ctx, scene_handbase, 0.005f * 1.0001f, 0.f, &G);
    if (rc != HOP_OK) return rc;
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    char *d = (char *)ctx->ensure_io(2 * up(sizeof(float) * (size_t)n_heights));
    if (!d) { ctx->err = "hop_adjust_hand_height: staging allocation failed"; return HOP_ENOMEM; }
    float *d_h = (float *)d; int *d_c = (int *)(d + up(sizeof(float) * (size_t)n_heights));
    HOP_CUDA(ctx, cudaMemcpyAsync(d_h, heights, sizeof(float) * (size_t)n_heights, cudaMemcpyHostToDevice, ctx->stream));
    HOP_CUDA(ctx, cudaMemsetAsync(d_c, 0, sizeof(int) * (size_t)n_heights, ctx->stream));
    HeightArgs a;
    a.h_pw = hand_cloud->d_pw; a.h_nv = hand_cloud->d_nv; a.nh = hand_cloud->n; a.grid = G->dev; a.s_nv = scene_handbase->d_nv;
    a.heights = d_h; a.n_heights = n_heights; a.counts = d_c; a.cos_thr = std::cos(45 / 180.0 * M_PI);
    {
      ProfScope ps(ctx, HOP_PROF_HAND);
      hand_height_kernel<<<dim3((a.nh + 127) / 128, n_heights), 128, 0, ctx->stream>>>(a);
      ctx->launches += 1;
    }
    HOP_CUDA(ctx, cudaGetLastError());
    HOP_CUDA(ctx, cudaMemcpyAsync(counts.data(), d_c, sizeof(int32_t) * (size_t)n_heights, cudaMemcpyDeviceToHost, ctx->stream));
    HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  // Hand.cpp:1040-1045: the first height whose count exceeds every earlier one; none above zero -> best_height stays 0 (index -1 here)
  int best = -1, max_match = 0;
  for (int t = 0; t < n_heights; ++t) if (counts[t] > max_match) { max_match = counts[t]; best = t; }
  *best_index = best;
  if (match_counts) for (int t = 0; t < n_heights; ++t) match_counts[t] = counts[t];
  return HOP_OK;
}
