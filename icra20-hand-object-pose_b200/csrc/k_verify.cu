// k_verify.cu -- K3: per congruent quadrilateral rigid transform + LCP verification (Super4PCS inner loop) for sm_100a.
//
//   replaces gr::CongruentSetExplorationBase::TryCongruentSet  (congruentSetExplorationBase.hpp:221-340)
//            gr::MatchBase::ComputeRigidTransformation          (matchBase.hpp:230-377)
//            gr::CongruentSetExplorationBase::Verify            (congruentSetExplorationBase.hpp:346-435)
//            + KdTree::doQueryRestrictedClosestIndex            (accelerators/kdtree.h:342-404)
//
// One warp per congruent quadrilateral (the reference: one OpenMP task per quadrilateral).  Every lane rebuilds the
// 3-point frame transform (a few dozen flops, cheaper than broadcasting it), the lanes then stride over the sampled
// model points Q, move each into the scene frame and ask the scene's exact nearest-neighbour grid whether a scene point
// lies within delta; ballots count the hits.  Quadrilaterals of ALL trials of a frame go through one launch; the
// survivors (lcp > 0) are compacted in (trial, quadrilateral) order -- the reference's single-thread emission order --
// with a stable prefix-sum selection, never with atomics.
#include <cub/device/device_select.cuh>

#include "hop_common.cuh"

namespace {

struct VerifyArgs {
  const float4 *P;        // centred scene (cloud pw stream)
  NNGridDev grid;         // on the centred scene, radius >= delta
  const float4 *Q;        // centred sampled model points (x, y, z, -)
  int nQ;
  const int4 *bases;      // per trial: 4 indices into P
  const int4 *quads;      // per quadrilateral: 4 indices into Q
  const int *quad_trial;  // per quadrilateral: its trial
  int M;
  float delta, delta2;
  float cP[3], cQ[3];     // centroids removed from P and Q (MatchBase::init, matchBase.hpp:425-432)
  float *poses;           // M x 16, global frame, column-major
  float *lcp;             // M
  int *valid;             // M: 1 when lcp > 0 (the hypothesis is emitted)
};

// The reference's (rotation^T rotation).isIdentity(1e-6) gate sits at float rounding level, so the transform is
// rebuilt with Eigen's exact operation order and NO fused multiply-adds: explicit round-to-nearest intrinsics, and
// 3-term reductions associated as t0 + (t1 + t2) like Eigen's redux_novec_unroller does for fixed size 3.
__device__ __forceinline__ float3 f3(float4 a) { return make_float3(a.x, a.y, a.z); }
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float sum3(float t0, float t1, float t2) { return __fadd_rn(t0, __fadd_rn(t1, t2)); }
__device__ __forceinline__ float3 sub(float3 a, float3 b) { return make_float3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return sum3(mul(a.x, b.x), mul(a.y, b.y), mul(a.z, b.z)); }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
  return make_float3(__fsub_rn(mul(a.y, b.z), mul(a.z, b.y)), __fsub_rn(mul(a.z, b.x), mul(a.x, b.z)), __fsub_rn(mul(a.x, b.y), mul(a.y, b.x)));
}
__device__ __forceinline__ float3 centroid3(float3 a, float3 b, float3 c) {
  return make_float3(__fdiv_rn(__fadd_rn(__fadd_rn(a.x, b.x), c.x), 3.f), __fdiv_rn(__fadd_rn(__fadd_rn(a.y, b.y), c.y), 3.f),
                     __fdiv_rn(__fadd_rn(__fadd_rn(a.z, b.z), c.z), 3.f));
}

// Gram-Schmidt frame of three points (matchBase.hpp:283-299): rows e1, e2, e3.  false when degenerate.
__device__ __forceinline__ bool frame3(float3 a0, float3 a1, float3 a2, float3 &e1, float3 &e2, float3 &e3) {
  e1 = sub(a1, a0);
  float n = dot3(e1, e1);
  if (n == 0.f) return false;
  n = __fsqrt_rn(n);
  e1 = make_float3(__fdiv_rn(e1.x, n), __fdiv_rn(e1.y, n), __fdiv_rn(e1.z, n));   // Eigen normalize(): v /= norm
  const float3 d = sub(a2, a0);
  const float k = dot3(d, e1);
  e2 = make_float3(__fsub_rn(d.x, mul(k, e1.x)), __fsub_rn(d.y, mul(k, e1.y)), __fsub_rn(d.z, mul(k, e1.z)));
  n = dot3(e2, e2);
  if (n == 0.f) return false;
  n = __fsqrt_rn(n);
  e2 = make_float3(__fdiv_rn(e2.x, n), __fdiv_rn(e2.y, n), __fdiv_rn(e2.z, n));
  e3 = cross3(e1, e2);
  return true;
}

__global__ void __launch_bounds__(256) verify_lcp_kernel(VerifyArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= a.M) return;
  const int4 qi = __ldg(&a.quads[warp]);
  const int4 bi = __ldg(&a.bases[__ldg(&a.quad_trial[warp])]);
  const float3 p0 = f3(__ldg(&a.P[bi.x])), p1 = f3(__ldg(&a.P[bi.y])), p2 = f3(__ldg(&a.P[bi.z]));
  const float3 q0 = f3(__ldg(&a.Q[qi.x])), q1 = f3(__ldg(&a.Q[qi.y])), q2 = f3(__ldg(&a.Q[qi.z]));
  // centroids of the first three points (congruentSetExplorationBase.hpp:243, 268-271)
  const float3 c1 = centroid3(p0, p1, p2), c2 = centroid3(q0, q1, q2);

  float3 fp1, fp2, fp3, fq1, fq2, fq3;
  bool ok = frame3(p0, p1, p2, fp1, fp2, fp3) && frame3(q0, q1, q2, fq1, fq2, fq3);
  float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  float rms = 3.0e38f;
  if (ok) {
    // rotation = rotate_p^T * rotate_q, rows of rotate_* are the frame vectors (matchBase.hpp:304-315)
    const float P_[9] = {fp1.x, fp1.y, fp1.z, fp2.x, fp2.y, fp2.z, fp3.x, fp3.y, fp3.z};
    const float Q_[9] = {fq1.x, fq1.y, fq1.z, fq2.x, fq2.y, fq2.z, fq3.x, fq3.y, fq3.z};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) R[3 * i + j] = sum3(mul(P_[i], Q_[j]), mul(P_[3 + i], Q_[3 + j]), mul(P_[6 + i], Q_[6 + j]));
    // (rotation^T rotation).isIdentity(1e-6): off-diagonals negligible against 1, diagonal ~ 1   (:322-325)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float g = sum3(mul(R[i], R[j]), mul(R[3 + i], R[3 + j]), mul(R[6 + i], R[6 + j]));
        if (i == j) { if (!(fabsf(__fsub_rn(g, 1.f)) <= mul(1e-6f, fminf(fabsf(g), 1.f)))) ok = false; }
        else if (!(fabsf(g) <= 1e-6f)) ok = false;
      }
  }
  if (ok) {
    // rms = sum_{i<3} | R (q_i - c2) - (p_i - c1) | / ref.size() with ref.size() == 4   (:349-360)
    const float3 qs[3] = {q0, q1, q2}, ps[3] = {p0, p1, p2};
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float3 f = sub(qs[i], c2);
      const float3 t = make_float3(sum3(mul(R[0], f.x), mul(R[1], f.y), mul(R[2], f.z)), sum3(mul(R[3], f.x), mul(R[4], f.y), mul(R[5], f.z)),
                                   sum3(mul(R[6], f.x), mul(R[7], f.y), mul(R[8], f.z)));
      const float3 d = make_float3(__fadd_rn(__fsub_rn(t.x, ps[i].x), c1.x), __fadd_rn(__fsub_rn(t.y, ps[i].y), c1.y),
                                   __fadd_rn(__fsub_rn(t.z, ps[i].z), c1.z));
      s = __fadd_rn(s, __fsqrt_rn(dot3(d, d)));
    }
    rms = __fdiv_rn(s, 4.f);
  }
  // transform = translate(c1) * rotate(R) * translate(-c2)   (:373-378), in the centred frames
  const float tx = __fsub_rn(c1.x, sum3(mul(R[0], c2.x), mul(R[1], c2.y), mul(R[2], c2.z)));
  const float ty = __fsub_rn(c1.y, sum3(mul(R[3], c2.x), mul(R[4], c2.y), mul(R[5], c2.z)));
  const float tz = __fsub_rn(c1.z, sum3(mul(R[6], c2.x), mul(R[7], c2.y), mul(R[8], c2.z)));

  int hits = 0;
  const bool go = ok && rms >= 0.f && rms < a.delta;  // congruentSetExplorationBase.hpp:291-295 (distance_factor 1)
  if (go) {
    for (int i = lane; i < a.nQ; i += 32) {
      const float4 q = __ldg(&a.Q[i]);
      const float x = R[0] * q.x + R[1] * q.y + R[2] * q.z + tx;
      const float y = R[3] * q.x + R[4] * q.y + R[5] * q.z + ty;
      const float z = R[6] * q.x + R[7] * q.y + R[8] * q.z + tz;
      float bd; float4 bp;
      const int j = nn_query(a.grid, x, y, z, bd, bp);
      hits += (j >= 0 && bd <= a.delta2) ? 1 : 0;       // kdtree.h:367 (sqdist <= cl_dist)
    }
  }
  hits = __reduce_add_sync(0xffffffffu, hits);
  if (lane == 0) {
    const float lcp = go ? (float)hits / (float)a.nQ : 0.f;  // Verify returns good_points / |Q| (:434)
    a.lcp[warp] = lcp;
    a.valid[warp] = lcp > 0.f ? 1 : 0;
    // getGlobalTransform (:313-321): same linear part, translation c1 + cP - R (c2 + cQ)
    const float gx = c2.x + a.cQ[0], gy = c2.y + a.cQ[1], gz = c2.z + a.cQ[2];
    float *m = a.poses + 16 * (size_t)warp;
#pragma unroll
    for (int j = 0; j < 3; ++j) { m[4 * j] = R[j]; m[4 * j + 1] = R[3 + j]; m[4 * j + 2] = R[6 + j]; m[4 * j + 3] = 0.f; }
    m[12] = c1.x + a.cP[0] - (R[0] * gx + R[1] * gy + R[2] * gz);
    m[13] = c1.y + a.cP[1] - (R[3] * gx + R[4] * gy + R[5] * gz);
    m[14] = c1.z + a.cP[2] - (R[6] * gx + R[7] * gy + R[8] * gz);
    m[15] = 1.f;
  }
}

struct Pose64 { float4 a, b, c, d; };

__global__ void repack_q_kernel(const float *__restrict__ xyz, int n, float4 *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.f);
}

}  // namespace

// device entry: everything already in HBM.  d_Q: nQ x 3 floats; d_bases: T x 4; d_quads: M x 4; d_quad_trial: M.
// outputs (device): poses M x 16, lcp M, valid M; when d_n_valid != null the survivors are also compacted IN PLACE to
// the front of poses / lcp (stable) and their count written to *d_n_valid.
extern "C" int hop_verify_lcp_dev(hop_ctx *ctx, hop_cloud *P_centered, const float *d_Q, int nQ, const int32_t *d_bases, int T,
                                  const int32_t *d_quads, const int32_t *d_quad_trial, int M, const float *centroid_P,
                                  const float *centroid_Q, float delta, float *d_poses, float *d_lcp, int32_t *d_valid,
                                  int32_t *d_n_valid) {
  HOP_ENTER(ctx);
  if (!ctx || !P_centered || M < 0 || nQ < 0 || T < 0) return HOP_EINVAL;
  if (M == 0) { if (d_n_valid) HOP_CUDA(ctx, cudaMemsetAsync(d_n_valid, 0, sizeof(int32_t), ctx->stream)); return HOP_OK; }
  if (!d_Q || !d_bases || !d_quads || !d_quad_trial || !d_poses || !d_lcp || !d_valid || !centroid_P || !centroid_Q || !(delta > 0.f) ||
      nQ == 0 || P_centered->n <= 0) { ctx->err = "hop_verify_lcp: bad arguments"; return HOP_EINVAL; }
  NNGridHost *G = nullptr;
  int rc = hop_get_nn_grid(ctx, P_centered, delta, 0.f, &G);
  if (rc != HOP_OK) return rc;
  // scratch: Q as float4 | compacted copies | cub temp
  size_t cub_a = 0, cub_b = 0;
  cub::DeviceSelect::Flagged(nullptr, cub_a, (Pose64 *)nullptr, (int *)nullptr, (Pose64 *)nullptr, (int *)nullptr, M, ctx->stream);
  cub::DeviceSelect::Flagged(nullptr, cub_b, (float *)nullptr, (int *)nullptr, (float *)nullptr, (int *)nullptr, M, ctx->stream);
  const size_t cub_bytes = std::max(cub_a, cub_b);
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t q_bytes = up(sizeof(float4) * (size_t)nQ), pose_bytes = up(64 * (size_t)M), lcp_bytes = up(4 * (size_t)M);
  char *w = (char *)ctx->ensure_work(q_bytes + pose_bytes + lcp_bytes + up(cub_bytes) + 256);
  if (!w) { ctx->err = "hop_verify_lcp: work buffer allocation failed"; return HOP_ENOMEM; }
  float4 *d_Q4 = (float4 *)w;
  Pose64 *d_pose_tmp = (Pose64 *)(w + q_bytes);
  float *d_lcp_tmp = (float *)(w + q_bytes + pose_bytes);
  void *d_cub = w + q_bytes + pose_bytes + lcp_bytes;
  int *d_cnt = (int *)(w + q_bytes + pose_bytes + lcp_bytes + up(cub_bytes));

  repack_q_kernel<<<(nQ + 127) / 128, 128, 0, ctx->stream>>>(d_Q, nQ, d_Q4);
  VerifyArgs a;
  a.P = P_centered->d_pw; a.grid = G->dev; a.Q = d_Q4; a.nQ = nQ;
  a.bases = (const int4 *)d_bases; a.quads = (const int4 *)d_quads; a.quad_trial = d_quad_trial; a.M = M;
  a.delta = delta; a.delta2 = delta * delta;
  for (int k = 0; k < 3; ++k) { a.cP[k] = centroid_P[k]; a.cQ[k] = centroid_Q[k]; }
  a.poses = d_poses; a.lcp = d_lcp; a.valid = d_valid;
  {
    ProfScope ps(ctx, HOP_PROF_VERIFY);
    verify_lcp_kernel<<<(M + 7) / 8, 256, 0, ctx->stream>>>(a);
  }
  ctx->launches += 2;
  if (d_n_valid) {
    size_t tb = cub_bytes;
    cub::DeviceSelect::Flagged(d_cub, tb, (const Pose64 *)d_poses, (const int *)d_valid, d_pose_tmp, d_cnt, M, ctx->stream);
    tb = cub_bytes;
    cub::DeviceSelect::Flagged(d_cub, tb, (const float *)d_lcp, (const int *)d_valid, d_lcp_tmp, d_cnt, M, ctx->stream);
    HOP_CUDA(ctx, cudaMemcpyAsync(d_poses, d_pose_tmp, 64 * (size_t)M, cudaMemcpyDeviceToDevice, ctx->stream));
    HOP_CUDA(ctx, cudaMemcpyAsync(d_lcp, d_lcp_tmp, 4 * (size_t)M, cudaMemcpyDeviceToDevice, ctx->stream));
    HOP_CUDA(ctx, cudaMemcpyAsync(d_n_valid, d_cnt, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->launches += 4;
  }
  HOP_CUDA(ctx, cudaGetLastError());
  return HOP_OK;
}

// host entry: host pointers in, blocks until poses / lcp / valid (per quadrilateral, uncompacted) and, when
// hyp_poses / hyp_lcp are given, the compacted hypothesis list (capacity M) + *n_hyp are in the caller's buffers.
extern "C" int hop_verify_lcp(hop_ctx *ctx, hop_cloud *P_centered, const float *Q_xyz, int nQ, const int32_t *bases, int T,
                              const int32_t *quads, const int32_t *quad_trial, int M, const float *centroid_P, const float *centroid_Q,
                              float delta, float *poses, float *lcp, int32_t *valid, float *hyp_poses, float *hyp_lcp, int32_t *n_hyp) {
  HOP_ENTER(ctx);
  if (!ctx || M < 0) return HOP_EINVAL;
  if (n_hyp) *n_hyp = 0;
  if (M == 0) return HOP_OK;
  if (!Q_xyz || !bases || !quads || !quad_trial || nQ <= 0 || T <= 0) { ctx->err = "hop_verify_lcp: bad arguments"; return HOP_EINVAL; }
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t bq = up(12 * (size_t)nQ), bb = up(16 * (size_t)T), bqu = up(16 * (size_t)M), bt = up(4 * (size_t)M);
  const size_t bp = up(64 * (size_t)M), bl = up(4 * (size_t)M), bv = up(4 * (size_t)M);
  char *d = (char *)ctx->ensure_io(bq + bb + bqu + bt + 2 * bp + 2 * bl + bv + 256);
  if (!d) { ctx->err = "hop_verify_lcp: staging allocation failed"; return HOP_ENOMEM; }
  float *d_Q = (float *)d; int32_t *d_bases = (int32_t *)(d + bq); int32_t *d_quads = (int32_t *)(d + bq + bb);
  int32_t *d_qt = (int32_t *)(d + bq + bb + bqu);
  float *d_poses = (float *)(d + bq + bb + bqu + bt); float *d_lcp = (float *)((char *)d_poses + bp);
  int32_t *d_valid = (int32_t *)((char *)d_lcp + bl);
  float *d_hp = (float *)((char *)d_valid + bv); float *d_hl = (float *)((char *)d_hp + bp); int32_t *d_n = (int32_t *)((char *)d_hl + bl);
  HOP_CUDA(ctx, cudaMemcpyAsync(d_Q, Q_xyz, 12 * (size_t)nQ, cudaMemcpyHostToDevice, ctx->stream));
  HOP_CUDA(ctx, cudaMemcpyAsync(d_bases, bases, 16 * (size_t)T, cudaMemcpyHostToDevice, ctx->stream));
  HOP_CUDA(ctx, cudaMemcpyAsync(d_quads, quads, 16 * (size_t)M, cudaMemcpyHostToDevice, ctx->stream));
  HOP_CUDA(ctx, cudaMemcpyAsync(d_qt, quad_trial, 4 * (size_t)M, cudaMemcpyHostToDevice, ctx->stream));
  int rc = hop_verify_lcp_dev(ctx, P_centered, d_Q, nQ, d_bases, T, d_quads, d_qt, M, centroid_P, centroid_Q, delta, d_poses, d_lcp, d_valid, nullptr);
  if (rc != HOP_OK) return rc;
  if (poses) HOP_CUDA(ctx, cudaMemcpyAsync(poses, d_poses, 64 * (size_t)M, cudaMemcpyDeviceToHost, ctx->stream));
  if (lcp) HOP_CUDA(ctx, cudaMemcpyAsync(lcp, d_lcp, 4 * (size_t)M, cudaMemcpyDeviceToHost, ctx->stream));
  if (valid) HOP_CUDA(ctx, cudaMemcpyAsync(valid, d_valid, 4 * (size_t)M, cudaMemcpyDeviceToHost, ctx->stream));
  int32_t n = 0;
  if (hyp_poses || hyp_lcp || n_hyp) {
    // stable compaction on the device into second buffers (the per-quadrilateral arrays stay intact)
    size_t cub_a = 0, cub_b = 0;
    cub::DeviceSelect::Flagged(nullptr, cub_a, (Pose64 *)nullptr, (int *)nullptr, (Pose64 *)nullptr, (int *)nullptr, M, ctx->stream);
    cub::DeviceSelect::Flagged(nullptr, cub_b, (float *)nullptr, (int *)nullptr, (float *)nullptr, (int *)nullptr, M, ctx->stream);
    size_t cub_bytes = std::max(cub_a, cub_b);
    char *w = (char *)ctx->ensure_scratch(cub_bytes + 256);
    if (!w) { ctx->err = "hop_verify_lcp: scratch allocation failed"; return HOP_ENOMEM; }
    size_t tb = cub_bytes;
    cub::DeviceSelect::Flagged(w, tb, (const Pose64 *)d_poses, (const int *)d_valid, (Pose64 *)d_hp, d_n, M, ctx->stream);
    tb = cub_bytes;
    cub::DeviceSelect::Flagged(w, tb, (const float *)d_lcp, (const int *)d_valid, d_hl, d_n, M, ctx->stream);
    ctx->launches += 2;
    HOP_CUDA(ctx, cudaMemcpyAsync(&n, d_n, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (hyp_poses && n > 0) HOP_CUDA(ctx, cudaMemcpyAsync(hyp_poses, d_hp, 64 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (hyp_lcp && n > 0) HOP_CUDA(ctx, cudaMemcpyAsync(hyp_lcp, d_hl, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (n_hyp) *n_hyp = n;
  }
  HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HOP_OK;
}
