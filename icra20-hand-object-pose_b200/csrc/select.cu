// select.cu -- winners: top-K hypotheses by LCP score, written as PoseHypo-shaped records.
//   replaces the argmax inside PoseEstimator::selectBest (PoseEstimator.cpp:491-496) and the "keep the best 100"
//   cap of refineByICP (:241).  Order: score descending, ties -> lower id (the order of PoseEstimator.cpp:113-121).
// The record buffer may be the send slot of the winners' all-gather, so nothing sits between scoring and the
// collective.
#include "hop_common.cuh"

namespace {

// K rounds of a block-wide arg-max over (score, -id); H is a few 1e3..1e5 and K <= 128, so one CTA is enough and
// keeps the whole selection in one launch.  The scores are staged once in shared memory (SMEM = true, H up to ~50 k):
// a selected entry is masked by overwriting its copy with a value below every score, and the K rounds never go back
// to global memory.  Larger batches scan the global array and mask through a bitmap in scratch.
template <bool SMEM>
__global__ void __launch_bounds__(1024, 1) topk_kernel(const float *__restrict__ poses, const float *__restrict__ scores, int H,
                                                      int K, int32_t id_offset, int32_t frame, hop_pose_rec *__restrict__ out,
                                                      unsigned int *__restrict__ taken) {
  extern __shared__ float s_scores[];
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  __shared__ int s_best;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (SMEM) {
    for (int i = tid; i < H; i += blockDim.x) {
      float v = scores[i];
      if (!(v == v)) v = -INFINITY;  // NaN scores never win
      s_scores[i] = v;
    }
  } else {
    for (int i = tid; i < (H + 31) / 32; i += blockDim.x) taken[i] = 0u;
  }
  __syncthreads();
  for (int k = 0; k < K; ++k) {
    float bv = -INFINITY;
    int bi = -1;
    for (int i = tid; i < H; i += blockDim.x) {
      float v;
      if (SMEM) {
        v = s_scores[i];
        if (__float_as_uint(v) == 0xffc00001u) continue;  // taken (a NaN pattern no staged score can have)
      } else {
        if (taken[i >> 5] & (1u << (i & 31))) continue;
        v = scores[i];
        if (!(v == v)) v = -INFINITY;
      }
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
        if (bi >= 0) {
          if (SMEM) s_scores[bi] = __uint_as_float(0xffc00001u);
          else taken[bi >> 5] |= 1u << (bi & 31);
        }
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
    if (!SMEM) __syncthreads();  // (SMEM: the next round's barrier orders s_best; records of different rounds are disjoint)
  }
}

// The common case (the scores fit in shared memory, K <= 128) without a block barrier per winner: the 32 warps first extract, each
// on its own, the top K of the elements they own (K warp-wide arg-max rounds over the staged scores), then warp 0 merges the 32
// sorted candidate lists (lane l holds the head of warp l's list).  Scores are staged as order-preserving 32-bit keys (0 = taken /
// exhausted, below every real key) so that a warp-wide arg-max is two redux.sync instructions -- the largest key, then the lowest
// index holding it -- instead of a ten-shuffle butterfly.  Same total order (score descending, ties -> lower id; NaN = -inf;
// -0 = +0), same records.
__device__ __forceinline__ unsigned int topk_key(float v) {
  if (!(v == v)) v = -INFINITY;   // NaN scores never win
  if (v == 0.f) v = 0.f;          // -0 and +0 tie (as in a float comparison)
  const unsigned int u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // >= 0x007fffff (-inf): never 0
}
// every lane brings its best (key, index) (key 0 = nothing); returns the warp's winner index or -1, its key in `key`
__device__ __forceinline__ int topk_warp_argmax(unsigned int &key, int idx) {
  const unsigned int m = __reduce_max_sync(0xffffffffu, key);
  const int w = __reduce_min_sync(0xffffffffu, (key == m) ? idx : 0x7fffffff);
  key = m;
  return m != 0u ? w : -1;
}
__global__ void __launch_bounds__(1024, 1) topk_lists_kernel(const float *__restrict__ poses, const float *__restrict__ scores, int H,
                                                            int K, int32_t id_offset, int32_t frame, hop_pose_rec *__restrict__ out) {
  extern __shared__ unsigned int s_dyn[];
  __shared__ int s_win[128];
  const int KS = K | 1;                                      // odd list stride: the merge's 32 heads fall into 32 banks
  unsigned int *s_key = s_dyn;                               // H, padded to a multiple of 32
  unsigned int *c_key = s_dyn + ((H + 31) & ~31);            // 32 lists x KS
  int *c_idx = reinterpret_cast<int *>(c_key + 32 * KS);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < H; i += 1024) s_key[i] = topk_key(scores[i]);   // a warp only ever touches the elements it stages here
  __syncwarp();
  for (int k = 0; k < K; ++k) {
    unsigned int bk = 0u;
    int bi = 0x7fffffff;
    for (int i = tid; i < H; i += 1024) {
      const unsigned int v = s_key[i];
      if (v > bk) { bk = v; bi = i; }  // ascending i: first (lowest id) wins ties
    }
    const int w = topk_warp_argmax(bk, bi);
    if (lane == 0) {
      c_key[warp * KS + k] = bk; c_idx[warp * KS + k] = w;
      if (w >= 0) s_key[w] = 0u;
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 0) {
    int head = 0;
    for (int k = 0; k < K; ++k) {
      unsigned int bk = head < K ? c_key[lane * KS + head] : 0u;
      const int mine = (head < K && bk != 0u) ? c_idx[lane * KS + head] : 0x7fffffff;
      const int w = topk_warp_argmax(bk, mine);
      if (w >= 0 && mine == w) ++head;
      if (lane == 0) s_win[k] = w;
    }
  }
  __syncthreads();
  for (int t = tid; t < 20 * K; t += 1024) {
    const int k = t / 20, f = t - 20 * k, b = s_win[k];
    float *rec = reinterpret_cast<float *>(out + k);
    if (f < 16) rec[f] = b >= 0 ? poses[16 * (size_t)b + f] : ((f % 5 == 0) ? 1.f : 0.f);
    else if (f == 16) rec[16] = b >= 0 ? scores[b] : -INFINITY;
    else if (f == 17) reinterpret_cast<int32_t *>(rec)[17] = b >= 0 ? b + id_offset : -1;
    else if (f == 18) reinterpret_cast<int32_t *>(rec)[18] = frame;
    else reinterpret_cast<int32_t *>(rec)[19] = 0;
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
  const size_t smem = sizeof(float) * (size_t)H;
  const size_t smem_lists = sizeof(float) * (size_t)((H + 31) & ~31) + 8 * (size_t)32 * (size_t)(K | 1);
  const bool rounds_only = ctx->tune.topk_rounds;   // the one-barrier-pair-per-winner kernel (A/B knob)
  if (K <= 128 && smem_lists <= 200 * 1024 && !rounds_only) {
    HOP_CUDA(ctx, ctx->func_smem_optin(topk_lists_kernel, 200 * 1024));
    topk_lists_kernel<<<1, 1024, smem_lists, ctx->stream>>>(d_poses, d_scores, H, K, id_offset, frame, d_out);
  } else if (smem <= 200 * 1024) {
    HOP_CUDA(ctx, ctx->func_smem_optin(topk_kernel<true>, 200 * 1024));
    topk_kernel<true><<<1, 1024, smem, ctx->stream>>>(d_poses, d_scores, H, K, id_offset, frame, d_out, taken);
  } else {
    topk_kernel<false><<<1, 1024, 0, ctx->stream>>>(d_poses, d_scores, H, K, id_offset, frame, d_out, taken);
  }
  ctx->launches += 1;
  HOP_CUDA(ctx, cudaGetLastError());
  return HOP_OK;
}
