#include "hop_common.cuh"
#include "s4pcs.h"
extern "C" int hop_super4pcs_run(hop_ctx *ctx, hop_s4pcs_plan *plan, float *hyp_poses, float *hyp_lcp, int capacity, int32_t *n_hyp) {
  if (ctx) ctx->err = "hop_super4pcs_run: under construction";
  return HOP_EINVAL;
}
