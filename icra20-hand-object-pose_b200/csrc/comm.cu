// comm.cu -- the one collective of the path under the C ABI: an all-gather of the per-rank winner records.
//
// SURVEY 8e / north star: hypotheses are independent, so a frame's batch is partitioned over the ranks (one hop_ctx per GPU: one
// process per GPU, or one host thread per context) and the only exchange is ONE ncclAllGather of K x 80 bytes per rank at the
// end; every rank then holds the same world x K records and finishes selectBest / clusterPoses identically.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"), not at link time: a process that already carries an NCCL (torch bundles
// its own) keeps exactly one copy, and a single-GPU user of libhop needs no NCCL at all.  Only the five entry points below are
// used; their prototypes come from <nccl.h> (types only).
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "hop_common.cuh"

namespace {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

NcclApi *nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) { api.err = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : ""); return; }
    auto sym = [&](const char *n) { void *p = dlsym(api.lib, n); if (!p) api.err = std::string("NCCL symbol missing: ") + n; return p; };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  });
  return &api;
}

int nccl_fail(hop_ctx *ctx, NcclApi *N, const char *what, ncclResult_t r) {
  if (ctx) ctx->err = std::string(what) + ": " + (N->GetErrorString ? N->GetErrorString(r) : "NCCL error");
  return HOP_ECUDA;
}

}  // namespace

struct hop_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  cudaStream_t stream = nullptr;     // the collective's own stream: the all-gather of step k overlaps the compute of step k + 1
  cudaEvent_t ev_ready = nullptr;    // compute -> comm: the send slot is written
  cudaEvent_t ev_done = nullptr;     // comm -> compute / host: the receive buffer is complete
  void *d_send = nullptr, *d_recv = nullptr; size_t send_bytes = 0, recv_bytes = 0;   // staging of the host-buffer entry point
};

extern "C" {

int hop_comm_unique_id(void *id_out_128) {
  if (!id_out_128) return HOP_EINVAL;
  NcclApi *N = nccl_api();
  if (!N->lib || !N->err.empty()) return HOP_ENODEV;
  static_assert(sizeof(ncclUniqueId) == HOP_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  if (N->GetUniqueId(&id) != ncclSuccess) return HOP_ECUDA;
  std::memcpy(id_out_128, &id, sizeof(id));
  return HOP_OK;
}

int hop_comm_init(hop_ctx *ctx, const void *id_128, int rank, int world) {
  HOP_ENTER(ctx);
  if (!ctx || !id_128 || world < 1 || rank < 0 || rank >= world) { if (ctx) ctx->err = "hop_comm_init: bad arguments"; return HOP_EINVAL; }
  if (ctx->comm) { ctx->err = "hop_comm_init: the context already has a communicator"; return HOP_EINVAL; }
  NcclApi *N = nccl_api();
  if (!N->lib || !N->err.empty()) { ctx->err = "hop_comm_init: " + N->err; return HOP_ENODEV; }
  hop_comm *c = new hop_comm();
  c->rank = rank; c->world = world;
  ncclUniqueId id;
  std::memcpy(&id, id_128, sizeof(id));
  ncclResult_t r = N->CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) { delete c; return nccl_fail(ctx, N, "ncclCommInitRank", r); }
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming) != cudaSuccess) {
    ctx->err = "hop_comm_init: stream / event creation failed";
    N->CommDestroy(c->comm);
    delete c;
    return HOP_ECUDA;
  }
  ctx->comm = c;
  return HOP_OK;
}

int hop_comm_destroy(hop_ctx *ctx) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  hop_comm *c = ctx->comm;
  if (!c) return HOP_OK;
  cudaStreamSynchronize(c->stream);
  NcclApi *N = nccl_api();
  if (c->comm && N->CommDestroy) N->CommDestroy(c->comm);
  cudaEventDestroy(c->ev_ready); cudaEventDestroy(c->ev_done);
  cudaStreamDestroy(c->stream);
  cudaFree(c->d_send); cudaFree(c->d_recv);
  delete c;
  ctx->comm = nullptr;
  return HOP_OK;
}

int hop_comm_rank(const hop_ctx *ctx, int *rank, int *world) {
  if (!ctx) return HOP_EINVAL;
  if (rank) *rank = ctx->comm ? ctx->comm->rank : 0;
  if (world) *world = ctx->comm ? ctx->comm->world : 1;
  return HOP_OK;
}

// d_send: K records of this rank (device; typically what hop_select_topk_dev / hop_refine_score_select_dev just wrote),
// d_recv: world x K records.  overlap = 0: on the context's stream (stream-ordered like every _dev entry point).
// overlap = 1: on the communicator's own stream, ordered after what the context's stream has enqueued so far; the context's
// stream is NOT blocked -- the next frame's kernels run while the records travel; hop_gather_wait orders later work after it.
int hop_gather_winners_dev(hop_ctx *ctx, const hop_pose_rec *d_send, int K, hop_pose_rec *d_recv, int overlap) {
  HOP_ENTER(ctx);
  if (!ctx || K < 0 || (K > 0 && (!d_send || !d_recv))) { if (ctx) ctx->err = "hop_gather_winners: bad arguments"; return HOP_EINVAL; }
  if (K == 0) return HOP_OK;
  hop_comm *c = ctx->comm;
  const size_t bytes = sizeof(hop_pose_rec) * (size_t)K;
  if (!c || c->world == 1) {   // a single rank: the gather is a copy
    if ((const void *)d_send != (const void *)d_recv) HOP_CUDA(ctx, cudaMemcpyAsync(d_recv, d_send, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return HOP_OK;
  }
  NcclApi *N = nccl_api();
  cudaStream_t s = ctx->stream;
  if (overlap) {
    HOP_CUDA(ctx, cudaEventRecord(c->ev_ready, ctx->stream));
    HOP_CUDA(ctx, cudaStreamWaitEvent(c->stream, c->ev_ready, 0));
    s = c->stream;
  }
  ncclResult_t r = N->AllGather(d_send, d_recv, bytes, ncclChar, c->comm, s);
  if (r != ncclSuccess) return nccl_fail(ctx, N, "ncclAllGather", r);
  if (overlap) HOP_CUDA(ctx, cudaEventRecord(c->ev_done, c->stream));
  return HOP_OK;
}

// after an overlapped gather: block_host = 0 makes the context's stream wait for it (device-side), 1 also blocks the host
int hop_gather_wait(hop_ctx *ctx, int block_host) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  hop_comm *c = ctx->comm;
  if (!c || c->world == 1) return block_host ? hop_sync(ctx) : HOP_OK;
  HOP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, c->ev_done, 0));
  if (block_host) HOP_CUDA(ctx, cudaStreamSynchronize(c->stream));
  return HOP_OK;
}

// host buffers: local K records up, all-gather, world x K records down, one synchronisation
int hop_gather_winners(hop_ctx *ctx, const hop_pose_rec *local, int K, hop_pose_rec *all) {
  HOP_ENTER(ctx);
  if (!ctx || K < 0 || (K > 0 && (!local || !all))) { if (ctx) ctx->err = "hop_gather_winners: bad arguments"; return HOP_EINVAL; }
  if (K == 0) return HOP_OK;
  hop_comm *c = ctx->comm;
  const int world = c ? c->world : 1;
  const size_t sb = sizeof(hop_pose_rec) * (size_t)K, rb = sb * (size_t)world;
  if (!c) { std::memcpy(all, local, sb); return HOP_OK; }
  if (sb > c->send_bytes) { cudaFree(c->d_send); c->d_send = nullptr; HOP_CUDA(ctx, cudaMalloc(&c->d_send, sb)); c->send_bytes = sb; }
  if (rb > c->recv_bytes) { cudaFree(c->d_recv); c->d_recv = nullptr; HOP_CUDA(ctx, cudaMalloc(&c->d_recv, rb)); c->recv_bytes = rb; }
  HOP_CUDA(ctx, cudaMemcpyAsync(c->d_send, local, sb, cudaMemcpyHostToDevice, ctx->stream));
  int rc = hop_gather_winners_dev(ctx, (const hop_pose_rec *)c->d_send, K, (hop_pose_rec *)c->d_recv, 0);
  if (rc != HOP_OK) return rc;
  HOP_CUDA(ctx, cudaMemcpyAsync(all, c->d_recv, rb, cudaMemcpyDeviceToHost, ctx->stream));
  HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HOP_OK;
}

}  // extern "C"
