// hop_common.cuh -- shared device/host definitions for libhop (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include <chrono>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/hop_c_api.h"

#ifndef __CUDA_ARCH__
#define HOP_HOST_ONLY 1
#endif

// ------------------------------------------------------------------------------------------------------------
// device data layout
//   cloud  : pw[i] = (x, y, z, prob)   nv[i] = (nx, ny, nz, 1/|n| or 0)     both padded to a multiple of
//            HOP_TILE_PTS with sentinel points (x = HOP_SENTINEL) so tiles are always whole 16-byte-aligned
//            bulk copies.
//   nn grid: dense voxel grid over the cloud's bounding box inflated by the query radius.  cell[v] = (offset,
//            count) into cand[]; cand[k] = (x, y, z, int_as_float(point index)).  The list of a voxel holds every
//            point that can be the nearest neighbour (within `radius`) of ANY query falling in that voxel, so a
//            query is: one 8-byte gather + a short linear scan.  Exact, not approximate.
// ------------------------------------------------------------------------------------------------------------
#define HOP_TILE_PTS 256
#define HOP_SENTINEL 1.0e30f

struct NNGridDev {
  float ox, oy, oz;  // min corner
  float inv_e;       // 1 / voxel edge
  int nx, ny, nz;    // multiples of 4: voxels are stored in 4x4x4 tiles (see nn_vox_index)
  float radius;      // queries are exact for neighbours within this distance
  const uint2 *cell;
  const float4 *cand;
};

struct CloudDev {
  const float4 *pw;
  const float4 *nv;
  int n;         // real points
  int n_padded;  // multiple of HOP_TILE_PTS
};

struct Rigid {  // y = R x + t, row-major R
  float r[9];
  float t[3];
};

#ifdef __CUDACC__
// ---- small math ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 rigid_apply(const Rigid &T, float x, float y, float z) {
  return make_float3(fmaf(T.r[0], x, fmaf(T.r[1], y, fmaf(T.r[2], z, T.t[0]))),
                     fmaf(T.r[3], x, fmaf(T.r[4], y, fmaf(T.r[5], z, T.t[1]))),
                     fmaf(T.r[6], x, fmaf(T.r[7], y, fmaf(T.r[8], z, T.t[2]))));
}
__device__ __forceinline__ float3 rigid_rotate(const Rigid &T, float x, float y, float z) {
  return make_float3(fmaf(T.r[0], x, fmaf(T.r[1], y, T.r[2] * z)), fmaf(T.r[3], x, fmaf(T.r[4], y, T.r[5] * z)),
                     fmaf(T.r[6], x, fmaf(T.r[7], y, T.r[8] * z)));
}
// general inverse of an affine map with (nearly) orthonormal R: uses the adjugate so a slightly non-orthonormal
// input is inverted, not transposed.
__device__ __forceinline__ Rigid rigid_inverse(const Rigid &T) {
  const float *a = T.r;
  float c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
  float det = a[0] * c00 + a[1] * c01 + a[2] * c02;
  float id = 1.0f / det;
  Rigid I;
  I.r[0] = c00 * id; I.r[1] = (a[2] * a[7] - a[1] * a[8]) * id; I.r[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  I.r[3] = c01 * id; I.r[4] = (a[0] * a[8] - a[2] * a[6]) * id; I.r[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  I.r[6] = c02 * id; I.r[7] = (a[1] * a[6] - a[0] * a[7]) * id; I.r[8] = (a[0] * a[4] - a[1] * a[3]) * id;
  I.t[0] = -(I.r[0] * T.t[0] + I.r[1] * T.t[1] + I.r[2] * T.t[2]);
  I.t[1] = -(I.r[3] * T.t[0] + I.r[4] * T.t[1] + I.r[5] * T.t[2]);
  I.t[2] = -(I.r[6] * T.t[0] + I.r[7] * T.t[1] + I.r[8] * T.t[2]);
  return I;
}
// C = A o B  (apply B first)
__device__ __forceinline__ Rigid rigid_compose(const Rigid &A, const Rigid &B) {
  Rigid C;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) C.r[3 * i + j] = A.r[3 * i] * B.r[j] + A.r[3 * i + 1] * B.r[3 + j] + A.r[3 * i + 2] * B.r[6 + j];
    C.t[i] = A.r[3 * i] * B.t[0] + A.r[3 * i + 1] * B.t[1] + A.r[3 * i + 2] * B.t[2] + A.t[i];
  }
  return C;
}
// column-major 4x4 (Eigen) <-> Rigid
__device__ __forceinline__ Rigid rigid_load_colmajor(const float *m) {
  Rigid T;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) T.r[3 * i + j] = m[4 * j + i];
    T.t[i] = m[12 + i];
  }
  return T;
}
__device__ __forceinline__ void rigid_store_colmajor(const Rigid &T, float *m) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
#pragma unroll
    for (int i = 0; i < 3; ++i) m[4 * j + i] = T.r[3 * i + j];
    m[4 * j + 3] = 0.f;
  }
  m[12] = T.t[0]; m[13] = T.t[1]; m[14] = T.t[2]; m[15] = 1.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- exact nearest neighbour through the voxel grid ------------------------------------------------------------
// Storage order of the voxels: 4x4x4 tiles, tile-linear.  The 64 cells of a tile are 512 contiguous bytes and (the
// candidate lists being allocated by a scan in this order) their lists are contiguous too, so the 32 lanes of a warp that
// walks a Morton-sorted cloud gather from a handful of cache lines instead of one line per lane.
__host__ __device__ __forceinline__ size_t nn_vox_index(int ix, int iy, int iz, int nx, int ny) {
  const size_t tile = ((size_t)(iz >> 2) * (size_t)(ny >> 2) + (size_t)(iy >> 2)) * (size_t)(nx >> 2) + (size_t)(ix >> 2);
  return tile * 64 + (size_t)(((iz & 3) << 4) | ((iy & 3) << 2) | (ix & 3));
}
__host__ __device__ __forceinline__ void nn_vox_coords(size_t v, int nx, int ny, int &ix, int &iy, int &iz) {
  const size_t tile = v >> 6;
  const int l = (int)(v & 63), tnx = nx >> 2, tny = ny >> 2;
  ix = (int)(tile % tnx) * 4 + (l & 3);
  iy = (int)((tile / tnx) % tny) * 4 + ((l >> 2) & 3);
  iz = (int)(tile / ((size_t)tnx * tny)) * 4 + (l >> 4);
}
// returns the point index (or -1) and its squared distance; exact for neighbours within g.radius.
__device__ __forceinline__ int nn_query(const NNGridDev &g, float px, float py, float pz, float &best_d2, float4 &best_pt) {
  float fx = (px - g.ox) * g.inv_e, fy = (py - g.oy) * g.inv_e, fz = (pz - g.oz) * g.inv_e;
  int ix = __float2int_rd(fx), iy = __float2int_rd(fy), iz = __float2int_rd(fz);
  best_d2 = 3.0e38f;
  int best = -1;
  // (the negated comparison also rejects NaN coordinates)
  if (!(fx >= 0.f && fy >= 0.f && fz >= 0.f) || ix >= g.nx || iy >= g.ny || iz >= g.nz) return -1;
  uint2 c = __ldg(&g.cell[nn_vox_index(ix, iy, iz, g.nx, g.ny)]);
  const float4 *lst = g.cand + c.x;
  for (uint32_t k = 0; k < c.y; ++k) {
    float4 q = __ldg(&lst[k]);
    float dx = q.x - px, dy = q.y - py, dz = q.z - pz;
    // unfused, left to right: the bits of FLANN's / the oracle's dx*dx + dy*dy + dz*dz (index parity needs equal d2)
    float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    if (d2 < best_d2) { best_d2 = d2; best = __float_as_int(q.w); best_pt = q; }
  }
  return best;
}

// ---- mbarrier + 1-D bulk (TMA) copies --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// global -> shared bulk copy, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
#endif  // __CUDACC__

// ------------------------------------------------------------------------------------------------------------
// host-side objects
// ------------------------------------------------------------------------------------------------------------
struct NNGridHost {
  float radius = 0.f, voxel = 0.f;
  NNGridDev dev{};
  uint2 *d_cell = nullptr;
  float4 *d_cand = nullptr;
  unsigned int *d_info = nullptr;  // device: [0] max list length [1] entries [2] overflow flag
  int64_t n_vox = 0, n_cand = 0, cap_vox = 0, cap_cand = 0;
  int max_list = 0;
  cudaEvent_t ready = nullptr;     // recorded on the side stream by hop_cloud_prepare_nn_async
  bool pending = false;            // built on the side stream; the first consumer on the main stream waits for `ready`
};

struct hop_cloud {
  int n = 0, n_padded = 0, capacity = 0;
  float4 *d_pw = nullptr, *d_nv = nullptr;
  float *d_stage = nullptr;  // raw xyz|nrm|prob staging (capacity * 7 floats)
  float bbox_min[3] = {0, 0, 0}, bbox_max[3] = {0, 0, 0};
  uint64_t version = 0;  // bumped by hop_cloud_update; grids are rebuilt when stale
  bool is_static = false;  // hop_cloud_hint_static: contents stay (a model): its grids are built once, so they may be finer
  std::vector<NNGridHost *> grids;
  std::vector<uint64_t> grid_version;
  // query-order copy: the same points sorted along a Morton curve, so that the 32 lanes of a warp iterating the cloud fall
  // into the same few voxels of the structure they query (built lazily by hop_cloud_query_order, nn_grid.cu)
  float4 *d_pw_q = nullptr, *d_nv_q = nullptr;
  int q_capacity = 0;
  uint64_t q_version = ~0ull;
  CloudDev dev() const { return CloudDev{d_pw, d_nv, n, n_padded}; }
  CloudDev dev_query() const { return CloudDev{d_pw_q, d_nv_q, n, n_padded}; }
};

// per-kernel timing with CUDA events on the launching stream (hop_profile_*): bench.py's roofline source
struct ProfSpan { int kind; cudaEvent_t a, b; };

// tuning knobs (environment, read ONCE in hop_create; never in a launch path)
struct HopTuning {
  int fused_variant = 0;      // HOP_FUSED_VARIANT: 1 = 256-thread CTAs, 2 = 128-thread CTAs, 0 = by batch size
  bool fused_profile = false; // HOP_FUSED_PROFILE: per-phase cycle accounting of icp_fused_kernel
  int fused_slots = 0;        // HOP_FUSED_SLOTS: hypotheses per CTA at a time (0 = the batch's fair share, capped by the CTA's warps)
  int mom_group_chunks = 0;   // HOP_MOM_GROUP: 512-point chunks per work item of icp_moments_kernel (0 = auto)
  int lcp_variant = 0;        // HOP_LCP_VARIANT: resident CTAs per SM of lcp_score_kernel
  float voxel_scale = 0.7f;   // HOP_VOXEL_SCALE: multiplies the point-spacing estimate of the NN grids' automatic voxel edge (nn_grid.cu: auto_voxel)
  float voxel_max_frac = 1.f; // HOP_VOXEL_MAX_FRAC
  bool topk_rounds = false;   // HOP_TOPK_ROUNDS: the one-barrier-pair-per-winner kernel (A/B knob)
  bool cluster_blocks = false; // HOP_CLUSTER_BLOCKS: hop_cluster_poses_gpu always through the blocked kernels (A/B knob; default: the bit matrix up to 4096 hypotheses)
  bool plan_debug = false;    // HOP_PLAN_DEBUG: the Super4PCS planner prints its trial loop's time
  bool trace = false;         // HOP_TRACE: host wall time and call count of every C-ABI entry point, printed by hop_destroy
};

struct hop_comm;   // comm.cu: the NCCL communicator of a context (null for a single-GPU context)

struct HopTraceEntry { long long calls = 0; double ms = 0.0; };

struct hop_ctx {
  int device = 0;
  hop_comm *comm = nullptr;
  HopTuning tune;
  std::map<std::string, HopTraceEntry> trace;   // HOP_TRACE: per entry point (inclusive: an entry point that calls another counts both)
  // cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute of a function: remembered per context (= per device),
  // not per process, so a second context on another GPU opts its kernels in too
  std::unordered_map<const void *, size_t> func_smem;
  template <typename F> cudaError_t func_smem_optin(F *func, size_t bytes) {
    size_t &have = func_smem[(const void *)func];
    if (bytes <= have) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) have = bytes;
    return e;
  }
  long long *d_fused_prof = nullptr;   // HOP_FUSED_PROFILE counters (this context's device)
  bool profiling = false;
  std::vector<ProfSpan> spans;
  std::vector<cudaEvent_t> event_pool;
  double prof_ms[HOP_PROF_KINDS] = {0};
  int64_t prof_n[HOP_PROF_KINDS] = {0};
  cudaEvent_t prof_event();
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  int64_t launches = 0;
  std::string err;
  // scratch
  void *d_scratch = nullptr; size_t scratch_bytes = 0;
  void *h_pinned = nullptr; size_t pinned_bytes = 0;
  int *d_counter = nullptr;  // work-queue heads (a few ints)
  void *d_work = nullptr; size_t work_bytes = 0;  // ICP state + correspondence records / LCP partials
  void *ensure_scratch(size_t bytes);
  void *d_io = nullptr; size_t io_bytes = 0;      // device staging of the host-buffer entry points (poses, scores, ...)
  void *ensure_work(size_t bytes);
  void *ensure_io(size_t bytes);
  void *ensure_pinned(size_t bytes);
  // hop_cloud_prepare_nn_async: a second stream (with its own scratch) on which a frame's scene grid is built while the main
  // stream runs the stages that do not need it
  cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr;
  void *d_side_scratch = nullptr; size_t side_scratch_bytes = 0;
  hop_cloud *s4_scene = nullptr;   // hop_super4pcs_run: the centred scene of the current frame (buffers reused across frames)
};

// times everything enqueued on the context's stream during its lifetime as one span of `kind`
struct ProfScope {
  hop_ctx *ctx; int idx = -1;
  ProfScope(hop_ctx *c, int kind) : ctx(c) {
    if (!c->profiling) return;
    ProfSpan s{kind, c->prof_event(), c->prof_event()};
    cudaEventRecord(s.a, c->stream);
    c->spans.push_back(s);
    idx = (int)c->spans.size() - 1;
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(ctx->spans[idx].b, ctx->stream); }
};

// Every extern "C" entry point that takes a context runs with the context's device current, whatever the caller (or torch)
// left selected, and restores the caller's device on return.
struct HopDeviceGuard {
  int prev = -1; bool switched = false;
  explicit HopDeviceGuard(const hop_ctx *c) {
    if (!c) return;
    if (cudaGetDevice(&prev) == cudaSuccess && prev != c->device) switched = cudaSetDevice(c->device) == cudaSuccess;
  }
  ~HopDeviceGuard() { if (switched) cudaSetDevice(prev); }
  HopDeviceGuard(const HopDeviceGuard &) = delete;
  HopDeviceGuard &operator=(const HopDeviceGuard &) = delete;
};
// host-side tracing of the C ABI (HOP_TRACE=1): where a frame's wall time goes between the kernels
struct HopTraceScope {
  hop_ctx *c; const char *fn; std::chrono::steady_clock::time_point t0;
  HopTraceScope(const hop_ctx *ctx, const char *f) : c(ctx && ctx->tune.trace ? const_cast<hop_ctx *>(ctx) : nullptr), fn(f) {
    if (c) t0 = std::chrono::steady_clock::now();
  }
  ~HopTraceScope() {
    if (!c) return;
    HopTraceEntry &e = c->trace[fn];
    e.calls += 1;
    e.ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  HopTraceScope(const HopTraceScope &) = delete;
  HopTraceScope &operator=(const HopTraceScope &) = delete;
};
#define HOP_ENTER(ctx) HopDeviceGuard hop_device_guard_(ctx); HopTraceScope hop_trace_scope_(ctx, __func__)

#define HOP_CUDA(ctx, call)                                                                            \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) {                                                                           \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                 \
      return HOP_ECUDA;                                                                                \
    }                                                                                                  \
  } while (0)

// the main stream waits for the cloud's grids still being built on the side stream (before the cloud's contents change)
inline void hop_cloud_join_pending(hop_ctx *ctx, hop_cloud *c) {
  for (NNGridHost *G : c->grids)
    if (G->pending) { cudaStreamWaitEvent(ctx->stream, G->ready, 0); G->pending = false; }
}

// implemented in nn_grid.cu
int hop_build_nn_grid(hop_ctx *ctx, hop_cloud *cloud, float radius, float voxel, NNGridHost **out);
int hop_get_nn_grid(hop_ctx *ctx, hop_cloud *cloud, float radius, float voxel, NNGridHost **out);
int hop_get_nn_grid_any(hop_ctx *ctx, hop_cloud *cloud, float radius, float voxel_if_built, NNGridHost **out);
void hop_free_nn_grid(NNGridHost *g);
int hop_cloud_query_order(hop_ctx *ctx, hop_cloud *cloud);  // (re)builds d_pw_q / d_nv_q when stale
float hop_auto_voxel(const hop_ctx *ctx, const hop_cloud *cloud, float radius, float max_frac);
// implemented in icp_lcp.cu
int hop_launch_icp(hop_ctx *ctx, const CloudDev &scene, const CloudDev &model, const NNGridDev &grid, const NNGridDev *scene_grid,
                   float *d_poses, int H, const hop_icp_params &p, int32_t *d_iters, int32_t *d_conv);
int hop_launch_lcp(hop_ctx *ctx, const CloudDev &scene, const float4 *scene_nv_by_index, const CloudDev &model, const NNGridDev &model_grid,
                   const NNGridDev &scene_grid, const float *d_poses, int H, const hop_lcp_params &p, int use_weights,
                   float *d_scores);
int hop_debug_lm_solve_launch(hop_ctx *ctx, const float *d_sums, int n, float *d_x, int32_t *d_nfev, int32_t *d_status, long long *d_cycles);
// implemented in select.cu
int hop_launch_topk(hop_ctx *ctx, const float *d_poses, const float *d_scores, int H, int K, int32_t id_offset, int32_t frame,
                    hop_pose_rec *d_out);
