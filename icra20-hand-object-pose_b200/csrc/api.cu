// api.cu -- the C ABI of libhop.so (include/hop_c_api.h): context, clouds, host/device entry points.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "hop_common.cuh"

static std::string g_create_error;

void *hop_ctx::ensure_scratch(size_t bytes) {
  if (bytes <= scratch_bytes) return d_scratch;
  if (d_scratch) { cudaStreamSynchronize(stream); cudaFree(d_scratch); d_scratch = nullptr; scratch_bytes = 0; }
  size_t cap = std::max(bytes + bytes / 4, (size_t)1 << 20);
  if (cudaMalloc(&d_scratch, cap) != cudaSuccess) { d_scratch = nullptr; return nullptr; }
  scratch_bytes = cap;
  return d_scratch;
}

cudaEvent_t hop_ctx::prof_event() {
  if (!event_pool.empty()) { cudaEvent_t e = event_pool.back(); event_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

void *hop_ctx::ensure_io(size_t bytes) {
  if (bytes <= io_bytes) return d_io;
  if (d_io) { cudaStreamSynchronize(stream); cudaFree(d_io); d_io = nullptr; io_bytes = 0; }
  size_t cap = std::max(bytes + bytes / 4, (size_t)1 << 20);
  if (cudaMalloc(&d_io, cap) != cudaSuccess) { d_io = nullptr; return nullptr; }
  io_bytes = cap;
  return d_io;
}

void *hop_ctx::ensure_work(size_t bytes) {
  if (bytes <= work_bytes) return d_work;
  if (d_work) { cudaStreamSynchronize(stream); cudaFree(d_work); d_work = nullptr; work_bytes = 0; }
  size_t cap = std::max(bytes + bytes / 8, (size_t)1 << 20);
  if (cudaMalloc(&d_work, cap) != cudaSuccess) { d_work = nullptr; return nullptr; }
  work_bytes = cap;
  return d_work;
}

void *hop_ctx::ensure_pinned(size_t bytes) {
  if (bytes <= pinned_bytes) return h_pinned;
  if (h_pinned) { cudaStreamSynchronize(stream); cudaFreeHost(h_pinned); h_pinned = nullptr; pinned_bytes = 0; }
  size_t cap = std::max(bytes + bytes / 4, (size_t)1 << 20);
  if (cudaHostAlloc(&h_pinned, cap, cudaHostAllocDefault) != cudaSuccess) { h_pinned = nullptr; return nullptr; }
  pinned_bytes = cap;
  return h_pinned;
}

namespace {

// raw staging (xyz | nrm | prob, each contiguous) -> the two padded float4 streams
__global__ void repack_cloud_kernel(const float *__restrict__ xyz, const float *__restrict__ nrm, const float *__restrict__ prob,
                                    int n, int n_padded, float4 *__restrict__ pw, float4 *__restrict__ nv) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_padded) return;
  if (i >= n) {
    pw[i] = make_float4(HOP_SENTINEL, HOP_SENTINEL, HOP_SENTINEL, 0.f);
    nv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  float w = prob ? prob[i] : 1.f;
  pw[i] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], w);
  float nx = 0.f, ny = 0.f, nz = 0.f, inv = 0.f;
  if (nrm) {
    nx = nrm[3 * i]; ny = nrm[3 * i + 1]; nz = nrm[3 * i + 2];
    float s = nx * nx + ny * ny + nz * nz;
    inv = s > 0.f ? 1.f / sqrtf(s) : 0.f;  // NaN normals: s is NaN -> inv 0, and the NaN stays in (nx,ny,nz)
  }
  nv[i] = make_float4(nx, ny, nz, inv);
}

int cloud_fill(hop_ctx *ctx, hop_cloud *c, const float *xyz, const float *nrm, const float *prob, int n) {
  if (n < 0 || (n > 0 && !xyz)) { ctx->err = "hop_cloud: bad arguments"; return HOP_EINVAL; }
  hop_cloud_join_pending(ctx, c);
  const int n_padded = std::max(HOP_TILE_PTS, (n + HOP_TILE_PTS - 1) / HOP_TILE_PTS * HOP_TILE_PTS);
  if (n_padded > c->capacity) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(c->d_pw); cudaFree(c->d_nv); cudaFree(c->d_stage);
    c->d_pw = c->d_nv = nullptr; c->d_stage = nullptr;
    HOP_CUDA(ctx, cudaMalloc(&c->d_pw, sizeof(float4) * (size_t)n_padded));
    HOP_CUDA(ctx, cudaMalloc(&c->d_nv, sizeof(float4) * (size_t)n_padded));
    HOP_CUDA(ctx, cudaMalloc(&c->d_stage, sizeof(float) * 7 * (size_t)n_padded));
    c->capacity = n_padded;
  }
  c->n = n; c->n_padded = n_padded; c->version++;
  // one pass over the caller's arrays: copy into pinned staging and take the bounding box (finite points only)
  float *h = (float *)ctx->ensure_pinned(sizeof(float) * 7 * (size_t)std::max(n, 1));
  if (!h) { ctx->err = "hop_cloud: pinned staging allocation failed"; return HOP_ENOMEM; }
  // the staging buffer may still feed an earlier async copy of this context
  HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  std::memcpy(h, xyz, sizeof(float) * 3 * (size_t)n);
  for (int i = 0; i < n; ++i) {
    const float *p = xyz + 3 * (size_t)i;
    if (std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]) && std::fabs(p[0]) < 1e20f && std::fabs(p[1]) < 1e20f &&
        std::fabs(p[2]) < 1e20f)
      for (int d = 0; d < 3; ++d) { mn[d] = std::min(mn[d], p[d]); mx[d] = std::max(mx[d], p[d]); }
  }
  if (!(mn[0] <= mx[0])) { for (int d = 0; d < 3; ++d) mn[d] = mx[d] = 0.f; }
  for (int d = 0; d < 3; ++d) { c->bbox_min[d] = mn[d]; c->bbox_max[d] = mx[d]; }
  float *hn = h + 3 * (size_t)n, *hp = h + 6 * (size_t)n;
  if (nrm) std::memcpy(hn, nrm, sizeof(float) * 3 * (size_t)n);
  if (prob) std::memcpy(hp, prob, sizeof(float) * (size_t)n);
  float *dx = c->d_stage, *dn = c->d_stage + 3 * (size_t)n, *dp = c->d_stage + 6 * (size_t)n;
  if (n > 0) {
    // xyz|nrm|prob are contiguous in both staging buffers: one copy when all three are present
    if (nrm && prob) HOP_CUDA(ctx, cudaMemcpyAsync(dx, h, sizeof(float) * 7 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    else {
      HOP_CUDA(ctx, cudaMemcpyAsync(dx, h, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
      if (nrm) HOP_CUDA(ctx, cudaMemcpyAsync(dn, hn, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
      if (prob) HOP_CUDA(ctx, cudaMemcpyAsync(dp, hp, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    }
  }
  repack_cloud_kernel<<<(n_padded + 255) / 256, 256, 0, ctx->stream>>>(dx, nrm ? dn : nullptr, prob ? dp : nullptr, n, n_padded,
                                                                      c->d_pw, c->d_nv);
  ctx->launches += 1;
  HOP_CUDA(ctx, cudaGetLastError());
  return HOP_OK;
}

}  // namespace

// capacity for n points, no contents (the caller fills d_pw / d_nv on the device and sets the bounding box)
int hop_cloud_reserve(hop_ctx *ctx, hop_cloud *c, int n) {
  if (n < 0) return HOP_EINVAL;
  hop_cloud_join_pending(ctx, c);
  const int n_padded = std::max(HOP_TILE_PTS, (n + HOP_TILE_PTS - 1) / HOP_TILE_PTS * HOP_TILE_PTS);
  if (n_padded > c->capacity) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(c->d_pw); cudaFree(c->d_nv); cudaFree(c->d_stage);
    c->d_pw = c->d_nv = nullptr; c->d_stage = nullptr;
    HOP_CUDA(ctx, cudaMalloc(&c->d_pw, sizeof(float4) * (size_t)n_padded));
    HOP_CUDA(ctx, cudaMalloc(&c->d_nv, sizeof(float4) * (size_t)n_padded));
    HOP_CUDA(ctx, cudaMalloc(&c->d_stage, sizeof(float) * 7 * (size_t)n_padded));
    c->capacity = n_padded;
  }
  c->n = n; c->n_padded = n_padded; c->version++;
  return HOP_OK;
}

extern "C" {

int hop_create(int device, hop_ctx **out) {
  if (!out) return HOP_EINVAL;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    g_create_error = std::string("hop_create: no CUDA device (") + cudaGetErrorString(e) + "); libhop has no CPU path";
    return HOP_ENODEV;
  }
  if (device < 0 || device >= count) { g_create_error = "hop_create: device index out of range"; return HOP_EINVAL; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { g_create_error = "hop_create: cudaGetDeviceProperties failed"; return HOP_ECUDA; }
  if (prop.major < 10) {
    g_create_error = "hop_create: libhop is built for sm_100a (B200) only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
    return HOP_ENODEV;
  }
  // the context's resources are created with its device current; the caller's current device is restored on every return path
  struct Restore { int prev = -1; ~Restore() { if (prev >= 0) cudaSetDevice(prev); } } restore;
  if (cudaGetDevice(&restore.prev) != cudaSuccess) restore.prev = -1;
  if (cudaSetDevice(device) != cudaSuccess) { g_create_error = "hop_create: cudaSetDevice failed"; return HOP_ECUDA; }
  hop_ctx *ctx = new hop_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  {  // tuning knobs: read once here, never in a launch path
    const char *v;
    if ((v = getenv("HOP_FUSED_VARIANT"))) ctx->tune.fused_variant = atoi(v);
    ctx->tune.fused_profile = getenv("HOP_FUSED_PROFILE") != nullptr;
    if ((v = getenv("HOP_FUSED_SLOTS"))) ctx->tune.fused_slots = atoi(v);
    if ((v = getenv("HOP_MOM_GROUP"))) ctx->tune.mom_group_chunks = atoi(v);
    if ((v = getenv("HOP_LCP_VARIANT"))) ctx->tune.lcp_variant = atoi(v);
    if ((v = getenv("HOP_VOXEL_MAX_FRAC"))) ctx->tune.voxel_max_frac = (float)atof(v);
    if ((v = getenv("HOP_VOXEL_SCALE"))) ctx->tune.voxel_scale = (float)atof(v);
    ctx->tune.topk_rounds = getenv("HOP_TOPK_ROUNDS") != nullptr;
    ctx->tune.trace = getenv("HOP_TRACE") != nullptr;
    ctx->tune.plan_debug = getenv("HOP_PLAN_DEBUG") != nullptr;
    ctx->tune.cluster_blocks = getenv("HOP_CLUSTER_BLOCKS") != nullptr;
  }
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc(&ctx->d_counter, 64 * sizeof(int)) != cudaSuccess) {
    g_create_error = "hop_create: stream/counter allocation failed";
    delete ctx;
    return HOP_ECUDA;
  }
  cudaMemset(ctx->d_counter, 0, 64 * sizeof(int));
  {  // per-call scratch comes from the stream-ordered pool: keep freed blocks instead of returning them to the driver
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  *out = ctx;
  return HOP_OK;
}

void hop_destroy(hop_ctx *ctx) {
  HopDeviceGuard guard(ctx);
  if (!ctx) return;
  cudaStreamSynchronize(ctx->stream);
  if (ctx->tune.trace && !ctx->trace.empty()) {
    std::vector<std::pair<std::string, HopTraceEntry>> rows(ctx->trace.begin(), ctx->trace.end());
    std::sort(rows.begin(), rows.end(), [](const auto &a, const auto &b) { return a.second.ms > b.second.ms; });
    fprintf(stderr, "[hop trace] device %d: host wall time per C-ABI entry point (inclusive)\n", ctx->device);
    for (const auto &r : rows) fprintf(stderr, "[hop trace] %-36s %8lld calls %10.3f ms %9.4f ms/call\n", r.first.c_str(), r.second.calls, r.second.ms, r.second.ms / (double)r.second.calls);
  }
  hop_comm_destroy(ctx);
  if (ctx->s4_scene) { hop_cloud_free(ctx, ctx->s4_scene); ctx->s4_scene = nullptr; }
  if (ctx->side) { cudaStreamSynchronize(ctx->side); cudaStreamDestroy(ctx->side); }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  cudaFree(ctx->d_side_scratch);
  cudaFree(ctx->d_scratch);
  cudaFree(ctx->d_work);
  cudaFree(ctx->d_io);
  cudaFreeHost(ctx->h_pinned);
  cudaFree(ctx->d_counter);
  cudaFree(ctx->d_fused_prof);
  for (ProfSpan &s : ctx->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
  for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char *hop_last_error(const hop_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int hop_set_stream(hop_ctx *ctx, void *cuda_stream) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->side) HOP_CUDA(ctx, cudaStreamSynchronize(ctx->side));
  if (cuda_stream) {
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
  } else if (!ctx->own_stream) {
    HOP_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  return HOP_OK;
}

int hop_sync(hop_ctx *ctx) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HOP_OK;
}

int64_t hop_launch_count(const hop_ctx *ctx) { return ctx ? ctx->launches : 0; }

static int profile_collect(hop_ctx *ctx) {
  HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (ProfSpan &s : ctx->spans) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) { ctx->prof_ms[s.kind] += ms; ctx->prof_n[s.kind] += 1; }
    ctx->event_pool.push_back(s.a); ctx->event_pool.push_back(s.b);
  }
  ctx->spans.clear();
  return HOP_OK;
}

int hop_profile_enable(hop_ctx *ctx, int on) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  int rc = profile_collect(ctx);
  if (rc != HOP_OK) return rc;
  for (int k = 0; k < HOP_PROF_KINDS; ++k) { ctx->prof_ms[k] = 0.0; ctx->prof_n[k] = 0; }
  ctx->profiling = on != 0;
  return HOP_OK;
}

int hop_profile_read(hop_ctx *ctx, int kind, double *total_ms, int64_t *spans) {
  HOP_ENTER(ctx);
  if (!ctx || kind < 0 || kind >= HOP_PROF_KINDS) return HOP_EINVAL;
  int rc = profile_collect(ctx);
  if (rc != HOP_OK) return rc;
  if (total_ms) *total_ms = ctx->prof_ms[kind];
  if (spans) *spans = ctx->prof_n[kind];
  return HOP_OK;
}

void hop_default_icp_params(hop_icp_params *p) {
  if (!p) return;
  p->max_iter = 10; p->angle_deg = 45.f; p->max_dist = 0.01f; p->abs_mse_eps = 1e-6; p->mode = 0; p->solver = 0;
  p->team_warps = 0; p->pipeline = 0;
}
void hop_default_lcp_params(hop_lcp_params *p) {
  if (!p) return;
  p->dist = 0.001f; p->angle_deg = 10.f; p->use_normal = 1; p->use_dot_score = 1; p->use_reciprocal = 1; p->team_warps = 0;
}

int hop_malloc(hop_ctx *ctx, size_t bytes, void **dev_ptr) {
  HOP_ENTER(ctx);
  if (!ctx || !dev_ptr) return HOP_EINVAL;
  HOP_CUDA(ctx, cudaMalloc(dev_ptr, bytes ? bytes : 1));
  return HOP_OK;
}
int hop_free(hop_ctx *ctx, void *dev_ptr) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  HOP_CUDA(ctx, cudaFree(dev_ptr));
  return HOP_OK;
}
int hop_host_alloc(hop_ctx *ctx, size_t bytes, void **host_ptr) {
  HOP_ENTER(ctx);
  if (!ctx || !host_ptr) return HOP_EINVAL;
  HOP_CUDA(ctx, cudaHostAlloc(host_ptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return HOP_OK;
}
int hop_host_free(hop_ctx *ctx, void *host_ptr) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  HOP_CUDA(ctx, cudaFreeHost(host_ptr));
  return HOP_OK;
}
int hop_memcpy_h2d(hop_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  HOP_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return HOP_OK;
}
int hop_memcpy_d2h(hop_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  HOP_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return HOP_OK;
}

int hop_cloud_upload(hop_ctx *ctx, const float *xyz, const float *nrm, const float *prob, int n, hop_cloud **out) {
  HOP_ENTER(ctx);
  if (!ctx || !out) return HOP_EINVAL;
  hop_cloud *c = new hop_cloud();
  int rc = cloud_fill(ctx, c, xyz, nrm, prob, n);
  if (rc != HOP_OK) { hop_cloud_free(ctx, c); *out = nullptr; return rc; }
  *out = c;
  return HOP_OK;
}

int hop_cloud_update(hop_ctx *ctx, hop_cloud *cloud, const float *xyz, const float *nrm, const float *prob, int n) {
  HOP_ENTER(ctx);
  if (!ctx || !cloud) return HOP_EINVAL;
  return cloud_fill(ctx, cloud, xyz, nrm, prob, n);
}

int hop_cloud_free(hop_ctx *ctx, hop_cloud *cloud) {
  HOP_ENTER(ctx);
  if (!cloud) return HOP_OK;
  if (ctx) cudaStreamSynchronize(ctx->stream);
  if (ctx && ctx->side) cudaStreamSynchronize(ctx->side);
  for (NNGridHost *g : cloud->grids) hop_free_nn_grid(g);
  cudaFree(cloud->d_pw); cudaFree(cloud->d_nv); cudaFree(cloud->d_stage); cudaFree(cloud->d_pw_q); cudaFree(cloud->d_nv_q);
  delete cloud;
  return HOP_OK;
}

int hop_cloud_size(const hop_cloud *cloud) { return cloud ? cloud->n : 0; }

// ---- K4 -------------------------------------------------------------------------------------------------------
int hop_icp_refine_dev(hop_ctx *ctx, hop_cloud *scene, hop_cloud *model, float *d_poses_inout, int H, const hop_icp_params *params,
                       int32_t *d_iters_out, int32_t *d_converged_out) {
  HOP_ENTER(ctx);
  if (!ctx || !scene || !model || !params || H < 0 || (H > 0 && !d_poses_inout)) { if (ctx) ctx->err = "hop_icp_refine: bad arguments"; return HOP_EINVAL; }
  if (H == 0) return HOP_OK;
  if (!(params->max_dist > 0.f)) { ctx->err = "hop_icp_refine: max_dist must be > 0"; return HOP_EINVAL; }
  if (model->n <= 0) { ctx->err = "hop_icp_refine: empty model cloud"; return HOP_EINVAL; }
  NNGridHost *G = nullptr;
  int rc = hop_get_nn_grid(ctx, model, params->max_dist, 0.f, &G);
  if (rc != HOP_OK) return rc;
  NNGridHost *Gs = nullptr;
  if (params->mode == 1) {   // reciprocal correspondences: the scene's own grid (the reverse neighbour is within max_dist as well)
    if (scene->n <= 0) { ctx->err = "hop_icp_refine: empty scene cloud"; return HOP_EINVAL; }
    rc = hop_get_nn_grid(ctx, scene, params->max_dist * 1.01f, 0.f, &Gs);
    if (rc != HOP_OK) return rc;
  }
  rc = hop_cloud_query_order(ctx, scene);   // the scene is only iterated: walk it along a Morton curve
  if (rc != HOP_OK) return rc;
  return hop_launch_icp(ctx, scene->dev_query(), model->dev(), G->dev, Gs ? &Gs->dev : nullptr, d_poses_inout, H, *params, d_iters_out,
                        d_converged_out);
}

int hop_icp_refine(hop_ctx *ctx, hop_cloud *scene, hop_cloud *model, float *poses_inout, int H, const hop_icp_params *params,
                   int32_t *iters_out, int32_t *converged_out) {
  HOP_ENTER(ctx);
  if (!ctx || (H > 0 && !poses_inout)) return HOP_EINVAL;
  if (H <= 0) return H == 0 ? HOP_OK : HOP_EINVAL;
  const size_t pb = sizeof(float) * 16 * (size_t)H, ib = sizeof(int32_t) * (size_t)H;
  // device staging for this call: poses | iters | conv   (kept separate from the grid-build scratch)
  float *d_poses = (float *)ctx->ensure_io(pb + 2 * ib);
  if (!d_poses) { ctx->err = "hop_icp_refine: staging allocation failed"; return HOP_ENOMEM; }
  int32_t *d_it = (int32_t *)((char *)d_poses + pb), *d_cv = d_it + H;
  HOP_CUDA(ctx, cudaMemcpyAsync(d_poses, poses_inout, pb, cudaMemcpyHostToDevice, ctx->stream));
  int rc = hop_icp_refine_dev(ctx, scene, model, d_poses, H, params, d_it, d_cv);
  if (rc == HOP_OK) {
    cudaMemcpyAsync(poses_inout, d_poses, pb, cudaMemcpyDeviceToHost, ctx->stream);
    if (iters_out) cudaMemcpyAsync(iters_out, d_it, ib, cudaMemcpyDeviceToHost, ctx->stream);
    if (converged_out) cudaMemcpyAsync(converged_out, d_cv, ib, cudaMemcpyDeviceToHost, ctx->stream);
  }
  if (rc != HOP_OK) return rc;
  HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HOP_OK;
}

// diagnostics: the inner solver of K4 alone (host buffers)
int hop_debug_lm_solve(hop_ctx *ctx, const float *sums, int n, float *x_out, int32_t *nfev_out, int32_t *status_out, int64_t *cycles_out) {
  HOP_ENTER(ctx);
  if (!ctx || n < 0 || (n > 0 && (!sums || !x_out || !nfev_out || !status_out))) { if (ctx) ctx->err = "hop_debug_lm_solve: bad arguments"; return HOP_EINVAL; }
  if (n == 0) return HOP_OK;
  const size_t sb = sizeof(float) * 96 * (size_t)n, xb = sizeof(float) * 6 * (size_t)n, ib = sizeof(int32_t) * (size_t)n, cb = sizeof(long long) * (size_t)n;
  char *d = (char *)ctx->ensure_io(cb + sb + xb + 2 * ib);
  if (!d) { ctx->err = "hop_debug_lm_solve: staging allocation failed"; return HOP_ENOMEM; }
  long long *d_cy = (long long *)d;
  float *d_sums = (float *)(d + cb), *d_x = (float *)(d + cb + sb);
  int32_t *d_nf = (int32_t *)(d + cb + sb + xb), *d_st = d_nf + n;
  HOP_CUDA(ctx, cudaMemcpyAsync(d_sums, sums, sb, cudaMemcpyHostToDevice, ctx->stream));
  const int rc = hop_debug_lm_solve_launch(ctx, d_sums, n, d_x, d_nf, d_st, cycles_out ? d_cy : nullptr);
  if (rc != HOP_OK) return rc;
  HOP_CUDA(ctx, cudaMemcpyAsync(x_out, d_x, xb, cudaMemcpyDeviceToHost, ctx->stream));
  HOP_CUDA(ctx, cudaMemcpyAsync(nfev_out, d_nf, ib, cudaMemcpyDeviceToHost, ctx->stream));
  HOP_CUDA(ctx, cudaMemcpyAsync(status_out, d_st, ib, cudaMemcpyDeviceToHost, ctx->stream));
  if (cycles_out) HOP_CUDA(ctx, cudaMemcpyAsync(cycles_out, d_cy, cb, cudaMemcpyDeviceToHost, ctx->stream));
  HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HOP_OK;
}

// ---- K5 -------------------------------------------------------------------------------------------------------
// The scene grid of the scoring pass (the reciprocal term's nearest scene point): radius = dist with a little head room for rounding,
// voxel edge by the size of the batch that will query it.  Its build is per-frame work, (radius / edge)^3 scatter operations per scene
// point; its lists are walked once per hypothesis and matched scene point.  From a few thousand hypotheses up the finer grid pays
// (10 k x 10 k x 16 384: build 0.2 -> 4 ms on the second stream, hidden behind the ICP; lcp_score_kernel 4.3 -> 3.5 ms); a frame's 100
// hypotheses, or 4096 on a 2 k-point scene, are better off with the cheap build (gpurun_out/r03l, r03o).
static const int kLcpFineGridBatch = 8192;
static float lcp_scene_radius(const hop_lcp_params *p) { return p->dist * 1.01f; }
static float lcp_scene_voxel(hop_ctx *ctx, hop_cloud *scene, const hop_lcp_params *p, int batch) {
  return batch >= kLcpFineGridBatch ? hop_auto_voxel(ctx, scene, lcp_scene_radius(p), 0.5f) : 0.f;   // 0 = the cloud's own default
}

int hop_lcp_prepare_scene_async(hop_ctx *ctx, hop_cloud *scene, const hop_lcp_params *params, int expected_hypotheses) {
  HOP_ENTER(ctx);
  if (!ctx || !scene || !params) { if (ctx) ctx->err = "hop_lcp_prepare_scene_async: bad arguments"; return HOP_EINVAL; }
  if (!(params->dist > 0.f)) { ctx->err = "hop_lcp_prepare_scene_async: dist must be > 0"; return HOP_EINVAL; }
  if (scene->n <= 0) return HOP_OK;
  return hop_cloud_prepare_nn_async(ctx, scene, lcp_scene_radius(params), lcp_scene_voxel(ctx, scene, params, expected_hypotheses));
}

int hop_lcp_score_dev(hop_ctx *ctx, hop_cloud *scene, hop_cloud *model, const float *d_poses, int H, const hop_lcp_params *params,
                      int use_weights, float *d_scores_out) {
  HOP_ENTER(ctx);
  if (!ctx || !scene || !model || !params || H < 0 || (H > 0 && (!d_poses || !d_scores_out))) { if (ctx) ctx->err = "hop_lcp_score: bad arguments"; return HOP_EINVAL; }
  if (H == 0) return HOP_OK;
  if (!(params->dist > 0.f)) { ctx->err = "hop_lcp_score: dist must be > 0"; return HOP_EINVAL; }
  if (model->n <= 0 || scene->n <= 0) {  // empty clouds score 0 (the reference loop body never runs)
    HOP_CUDA(ctx, cudaMemsetAsync(d_scores_out, 0, sizeof(float) * (size_t)H, ctx->stream));
    return HOP_OK;
  }
  NNGridHost *Gm = nullptr, *Gs = nullptr;
  int rc = hop_get_nn_grid(ctx, model, params->dist, 0.f, &Gm);
  if (rc != HOP_OK) return rc;
  // the reciprocal neighbour is at most `dist` away (see lcp_score_kernel).  A grid prefetched for this radius (hop_lcp_prepare_scene_async)
  // is taken as it is, whatever its voxel; otherwise it is built here, sized for this batch
  rc = hop_get_nn_grid_any(ctx, scene, lcp_scene_radius(params), lcp_scene_voxel(ctx, scene, params, H), &Gs);
  if (rc != HOP_OK) return rc;
  rc = hop_cloud_query_order(ctx, scene);
  if (rc != HOP_OK) return rc;
  // iterate the Morton-ordered copy; the reciprocal neighbour's normal is looked up by ORIGINAL index (scene grid)
  return hop_launch_lcp(ctx, scene->dev_query(), scene->d_nv, model->dev(), Gm->dev, Gs->dev, d_poses, H, *params, use_weights, d_scores_out);
}

int hop_lcp_score(hop_ctx *ctx, hop_cloud *scene, hop_cloud *model, const float *poses, int H, const hop_lcp_params *params,
                  int use_weights, float *scores_out) {
  HOP_ENTER(ctx);
  if (!ctx || (H > 0 && (!poses || !scores_out))) return HOP_EINVAL;
  if (H <= 0) return H == 0 ? HOP_OK : HOP_EINVAL;
  const size_t pb = sizeof(float) * 16 * (size_t)H, sb = sizeof(float) * (size_t)H;
  float *d_poses = (float *)ctx->ensure_io(pb + sb);
  if (!d_poses) { ctx->err = "hop_lcp_score: staging allocation failed"; return HOP_ENOMEM; }
  float *d_scores = (float *)((char *)d_poses + pb);
  HOP_CUDA(ctx, cudaMemcpyAsync(d_poses, poses, pb, cudaMemcpyHostToDevice, ctx->stream));
  int rc = hop_lcp_score_dev(ctx, scene, model, d_poses, H, params, use_weights, d_scores);
  if (rc == HOP_OK) cudaMemcpyAsync(scores_out, d_scores, sb, cudaMemcpyDeviceToHost, ctx->stream);
  if (rc != HOP_OK) return rc;
  HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HOP_OK;
}

// ---- winners ---------------------------------------------------------------------------------------------------
int hop_select_topk_dev(hop_ctx *ctx, const float *d_poses, const float *d_scores, int H, int K, int32_t id_offset, int32_t frame,
                        hop_pose_rec *d_out) {
  HOP_ENTER(ctx);
  if (!ctx || H < 0 || K < 0 || (K > 0 && !d_out) || (H > 0 && (!d_poses || !d_scores))) return HOP_EINVAL;
  return hop_launch_topk(ctx, d_poses, d_scores, H, K, id_offset, frame, d_out);
}

int hop_select_topk(hop_ctx *ctx, const float *poses, const float *scores, int H, int K, int32_t id_offset, int32_t frame,
                    hop_pose_rec *out) {
  HOP_ENTER(ctx);
  if (!ctx || H < 0 || K < 0 || (K > 0 && !out) || (H > 0 && (!poses || !scores))) return HOP_EINVAL;
  if (K == 0) return HOP_OK;
  const size_t pb = sizeof(float) * 16 * (size_t)H, sb = sizeof(float) * (size_t)H, rb = sizeof(hop_pose_rec) * (size_t)K;
  char *d = (char *)ctx->ensure_io(pb + sb + rb + 64);
  if (!d) { ctx->err = "hop_select_topk: staging allocation failed"; return HOP_ENOMEM; }
  float *d_poses = (float *)d, *d_scores = (float *)(d + pb);
  hop_pose_rec *d_out = (hop_pose_rec *)(d + ((pb + sb + 15) / 16) * 16);
  if (H > 0) {
    HOP_CUDA(ctx, cudaMemcpyAsync(d_poses, poses, pb, cudaMemcpyHostToDevice, ctx->stream));
    HOP_CUDA(ctx, cudaMemcpyAsync(d_scores, scores, sb, cudaMemcpyHostToDevice, ctx->stream));
  }
  int rc = hop_launch_topk(ctx, d_poses, d_scores, H, K, id_offset, frame, d_out);
  if (rc == HOP_OK) cudaMemcpyAsync(out, d_out, rb, cudaMemcpyDeviceToHost, ctx->stream);
  if (rc != HOP_OK) return rc;
  HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HOP_OK;
}

// ---- refineByICP + selectBest in one call ------------------------------------------------------------------------------------
int hop_refine_score_select_dev(hop_ctx *ctx, hop_cloud *scene, hop_cloud *model_icp, hop_cloud *model_lcp, float *d_poses_inout, int H,
                                const hop_icp_params *icp, const hop_lcp_params *lcp, int use_weights, int K, int32_t id_offset, int32_t frame,
                                int32_t *d_iters, int32_t *d_conv, float *d_scores, hop_pose_rec *d_winners) {
  HOP_ENTER(ctx);
  if (!ctx || !scene || !model_icp || !icp || !lcp || H < 0 || K < 0 || (H > 0 && (!d_poses_inout || !d_scores)) || (K > 0 && !d_winners)) {
    if (ctx) ctx->err = "hop_refine_score_select: bad arguments";
    return HOP_EINVAL;
  }
  if (!model_lcp) model_lcp = model_icp;
  int rc = HOP_OK;
  if (H > 0) {
    // the scene grid of the scoring pass depends on the frame only: built on the second stream while the ICP runs
    if (scene->n > 0 && model_lcp->n > 0 && lcp->dist > 0.f) {
      rc = hop_cloud_prepare_nn_async(ctx, scene, lcp_scene_radius(lcp), lcp_scene_voxel(ctx, scene, lcp, H));
      if (rc != HOP_OK) return rc;
    }
    rc = hop_icp_refine_dev(ctx, scene, model_icp, d_poses_inout, H, icp, d_iters, d_conv);
    if (rc != HOP_OK) return rc;
    rc = hop_lcp_score_dev(ctx, scene, model_lcp, d_poses_inout, H, lcp, use_weights, d_scores);
    if (rc != HOP_OK) return rc;
  }
  if (K > 0) rc = hop_launch_topk(ctx, d_poses_inout, d_scores, H, K, id_offset, frame, d_winners);
  return rc;
}

int hop_refine_score_select(hop_ctx *ctx, hop_cloud *scene, hop_cloud *model_icp, hop_cloud *model_lcp, const float *poses_in, int H,
                            const hop_icp_params *icp, const hop_lcp_params *lcp, int use_weights, int K, float *poses_out,
                            float *scores_out, int32_t *iters_out, int32_t *converged_out, hop_pose_rec *winners) {
  HOP_ENTER(ctx);
  if (!ctx || H < 0 || K < 0 || (H > 0 && !poses_in) || (K > 0 && !winners)) { if (ctx) ctx->err = "hop_refine_score_select: bad arguments"; return HOP_EINVAL; }
  if (H == 0 && K == 0) return HOP_OK;
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t pb = up(sizeof(float) * 16 * (size_t)H), ib = up(sizeof(int32_t) * (size_t)H), sb = up(sizeof(float) * (size_t)H);
  const size_t rb = up(sizeof(hop_pose_rec) * (size_t)K);
  char *d = (char *)ctx->ensure_io(pb + 2 * ib + sb + rb + 256);
  if (!d) { ctx->err = "hop_refine_score_select: staging allocation failed"; return HOP_ENOMEM; }
  float *d_poses = (float *)d;
  int32_t *d_it = (int32_t *)(d + pb), *d_cv = (int32_t *)(d + pb + ib);
  float *d_scores = (float *)(d + pb + 2 * ib);
  hop_pose_rec *d_rec = (hop_pose_rec *)(d + pb + 2 * ib + sb);
  if (H > 0) HOP_CUDA(ctx, cudaMemcpyAsync(d_poses, poses_in, sizeof(float) * 16 * (size_t)H, cudaMemcpyHostToDevice, ctx->stream));
  int rc = hop_refine_score_select_dev(ctx, scene, model_icp, model_lcp, d_poses, H, icp, lcp, use_weights, K, 0, 0, d_it, d_cv, d_scores, d_rec);
  if (rc != HOP_OK) return rc;
  if (H > 0) {
    if (poses_out) HOP_CUDA(ctx, cudaMemcpyAsync(poses_out, d_poses, sizeof(float) * 16 * (size_t)H, cudaMemcpyDeviceToHost, ctx->stream));
    if (scores_out) HOP_CUDA(ctx, cudaMemcpyAsync(scores_out, d_scores, sizeof(float) * (size_t)H, cudaMemcpyDeviceToHost, ctx->stream));
    if (iters_out) HOP_CUDA(ctx, cudaMemcpyAsync(iters_out, d_it, sizeof(int32_t) * (size_t)H, cudaMemcpyDeviceToHost, ctx->stream));
    if (converged_out) HOP_CUDA(ctx, cudaMemcpyAsync(converged_out, d_cv, sizeof(int32_t) * (size_t)H, cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (K > 0) HOP_CUDA(ctx, cudaMemcpyAsync(winners, d_rec, sizeof(hop_pose_rec) * (size_t)K, cudaMemcpyDeviceToHost, ctx->stream));
  HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HOP_OK;
}

}  // extern "C"
