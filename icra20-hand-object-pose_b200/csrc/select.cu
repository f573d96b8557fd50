// select.cu -- winners: top-K hypotheses by LCP score, written as PoseHypo-shaped records.
//   replaces the argmax inside PoseEstimator::selectBest (PoseEstimator.cpp:491-496) and the "keep the best 100"
//   cap of refineByICP (:241).  Order: score descending, ties -> lower id (the order of PoseEstimator.cpp:113-121).
// The record buffer may be the send slot of the winners' all-gather, so nothing sits between scoring and the
// collective.
#include "hop_common.cuh"

namespace {

// K rounds of a block-wide arg-max over (score, -id); H is a few 1e3..1e5 and K <= 128, so one CTA is enough and
// keeps the whole selection in one launch.  Selected entries are masked through a bitmap in shared/global scratch.
__global__ void __launch_bounds__(1024, 1) topk_kernel(const float *__restrict__ poses, const float *__restrict__ scores, int H,
                                                      int K, int32_t id_offset, int32_t frame, hop_pose_rec *__restrict__ out,
                                                      unsigned int *__restrict__ taken) {
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  __shared__ int s_best;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < (H + 31) / 32; i += blockDim.x) taken[i] = 0u;
  __syncthreads();
  for (int k = 0; k < K; ++k) {
    float bv = -INFINITY;
    int bi = -1;
    for (int i = tid; i < H; i += blockDim.x) {
      if (taken[i >> 5] & (1u << (i & 31))) continue;
      float v = scores[i];
      if (!(v == v)) v = -INFINITY;  // NaN scores never win
      if (bi < 0 || v > bv) { bv = v; bi = i; }  // ascending i: first (lowest id) wins ties
    }
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_val[warp] = bv; s_idx[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      bv = lane < (blockDim.x >> 5) ? s_val[lane] : -INFINITY;
      bi = lane < (blockDim.x >> 5) ? s_idx[lane] : -1;
      for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
      }
      if (lane == 0) {
        s_best = bi;
        if (bi >= 0) taken[bi >> 5] |= 1u << (bi & 31);
      }
    }
    __syncthreads();
    const int b = s_best;
    if (tid < 20) {
      float *rec = reinterpret_cast<float *>(out + k);
      if (tid < 16) rec[tid] = b >= 0 ? poses[16 * (size_t)b + tid] : ((tid % 5 == 0) ? 1.f : 0.f);
      else if (tid == 16) rec[16] = b >= 0 ? scores[b] : -INFINITY;
      else if (tid == 17) reinterpret_cast<int32_t *>(rec)[17] = b >= 0 ? b + id_offset : -1;
      else if (tid == 18) reinterpret_cast<int32_t *>(rec)[18] = frame;
      else reinterpret_cast<int32_t *>(rec)[19] = 0;
    }
    __syncthreads();
  }
}

}  // namespace

int hop_launch_topk(hop_ctx *ctx, const float *d_poses, const float *d_scores, int H, int K, int32_t id_offset, int32_t frame,
                    hop_pose_rec *d_out) {
  if (K <= 0) return HOP_OK;
  if (H < 0) return HOP_EINVAL;
  size_t words = (size_t)(H + 31) / 32 + 1;
  // the bitmap lives at the tail of the context's small counter block when it fits, else in scratch
  unsigned int *taken = (unsigned int *)ctx->ensure_scratch(words * sizeof(unsigned int));
  if (!taken) { ctx->err = "topk: scratch allocation failed"; return HOP_ENOMEM; }
  ProfScope ps(ctx, HOP_PROF_TOPK);
  topk_kernel<<<1, 1024, 0, ctx->stream>>>(d_poses, d_scores, H, K, id_offset, frame, d_out, taken);
  ctx->launches += 1;
  HOP_CUDA(ctx, cudaGetLastError());
  return HOP_OK;
}
