// k_render.cu -- render-based rejection of pose hypotheses: a software rasteriser in place of the reference's
// OpenGL context, and the whole wrong-ratio comparison of PoseEstimator::rejectByRender for a batch of hypotheses.
//
//   replaces PoseEstimator::rejectByRender              src/perception/src/PoseEstimator.cpp:345-463
//            Renderer::addObject / doRender              src/perception/src/Renderer.cpp:42-81
//            pcl::simulation's GL camera (depth_sim)     src/depth_sim/src/range_likelihood.cpp:391-475, simulation_io.cpp:411-440,486-505
//
// The reference renders hand + object once per hypothesis with OpenGL (serially: one GL context) and then compares 480 x 640
// pixels per hypothesis under OpenMP.  Here: the hand is rasterised ONCE per frame into a depth image, together with the
// per-pixel difference to the real depth image and its running (sequential, float) sum; per hypothesis only the object is
// rasterised, into a tile the size of its own bounding box (one thread per (hypothesis, triangle), nearest fragment by
// atomicMin on the float bits); one thread per hypothesis then replays the reference's row-major float sums from the first
// row its object touches -- everything above is the shared prefix -- so the wrong ratio carries the same rounding as the
// reference's loop (its 3e5-term float sums lose the small differences once the sum is large; a tidier sum would rank the
// hypotheses differently).  Coverage is decided in integer arithmetic on a 1/256 sub-pixel grid and depth in unfused double,
// operation for operation what oracle/hop_oracle_render.c does: the two agree bit for bit.
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <numeric>
#include <vector>

#include "hop_common.cuh"

struct hop_render_scene {
  hop_render_params p;
  int n_px = 0;
  float *d_real = nullptr;       // the real depth image, metres
  float *d_zhand = nullptr;      // nearest hand surface per pixel (Z, metres; FLT_MAX = none)
  float *d_base_diff = nullptr;  // per-pixel difference of the hand-only render to the real image
  float *d_prefix = nullptr;     // n_px + 1: running float sum of d_base_diff in row-major order (prefix[i] = sum of pixels < i)
};

namespace {

constexpr int SUBPX = 256;
constexpr unsigned int Z_EMPTY = 0x7f7fffffu;   // FLT_MAX

struct Tile { int x0, y0, w, h; long long off; };

struct RasterArgs {
  hop_render_params p;
  const float *V; const int32_t *F; int nf;
  const float *poses;      // H x 16 column-major, or null (vertices already in the camera frame)
  const Tile *tiles;       // per hypothesis, or null = the whole image at offset 0
  unsigned int *zbuf;
  int H;
};

__device__ __forceinline__ long long snap(float s) { return (long long)floor(__dadd_rn(__dmul_rn((double)s, (double)SUBPX), 0.5)); }

// camera-frame vertex k of face f under hypothesis h; false when it is not in front of the near plane
__device__ __forceinline__ bool project(const RasterArgs &a, const float *T, const float *v, float &Z, long long &SX, long long &SY) {
  float x = v[0], y = v[1], z = v[2];
  if (T) {
    const float tx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], x), __fmul_rn(T[4], y)), __fmul_rn(T[8], z)), T[12]);
    const float ty = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[1], x), __fmul_rn(T[5], y)), __fmul_rn(T[9], z)), T[13]);
    const float tz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[2], x), __fmul_rn(T[6], y)), __fmul_rn(T[10], z)), T[14]);
    x = tx; y = ty; z = tz;
  }
  Z = z;
  if (!(z > a.p.z_near)) return false;
  const float sx = __fadd_rn(__fmul_rn(a.p.fx, __fdiv_rn(x, z)), a.p.cx);
  const float sy = __fadd_rn(__fmul_rn(a.p.fy, __fdiv_rn(y, z)), __fsub_rn((float)a.p.height, a.p.cy));
  SX = snap(sx); SY = snap(sy);
  return true;
}

// one thread per (hypothesis, triangle)
__global__ void raster_kernel(RasterArgs a) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)a.H * a.nf) return;
  const int h = (int)(gid / a.nf), f = (int)(gid % a.nf);
  const float *T = a.poses ? a.poses + 16 * (size_t)h : nullptr;
  Tile t;
  if (a.tiles) t = a.tiles[h]; else { t.x0 = 0; t.y0 = 0; t.w = a.p.width; t.h = a.p.height; t.off = 0; }
  if (t.w <= 0 || t.h <= 0) return;
  float Z[3]; long long SX[3], SY[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) if (!project(a, T, a.V + 3 * (size_t)a.F[3 * (size_t)f + k], Z[k], SX[k], SY[k])) return;
  long long area = (SX[1] - SX[0]) * (SY[2] - SY[0]) - (SY[1] - SY[0]) * (SX[2] - SX[0]);
  if (area == 0) return;
  if (area < 0) {   // both windings are drawn: swap vertices 1 and 2
    area = -area;
    long long s = SX[1]; SX[1] = SX[2]; SX[2] = s; s = SY[1]; SY[1] = SY[2]; SY[2] = s;
    const float z = Z[1]; Z[1] = Z[2]; Z[2] = z;
  }
  const long long mnx = min(SX[0], min(SX[1], SX[2])), mxx = max(SX[0], max(SX[1], SX[2]));
  const long long mny = min(SY[0], min(SY[1], SY[2])), mxy = max(SY[0], max(SY[1], SY[2]));
  long long x0 = mnx - SUBPX / 2 < 0 ? 0 : (mnx - SUBPX / 2 + SUBPX - 1) / SUBPX, x1 = (mxx - SUBPX / 2) / SUBPX;
  long long y0 = mny - SUBPX / 2 < 0 ? 0 : (mny - SUBPX / 2 + SUBPX - 1) / SUBPX, y1 = (mxy - SUBPX / 2) / SUBPX;
  x0 = max(x0, (long long)t.x0); y0 = max(y0, (long long)t.y0);
  x1 = min(x1, (long long)(t.x0 + t.w - 1)); y1 = min(y1, (long long)(t.y0 + t.h - 1));
  const double iz0 = __drcp_rn((double)Z[0]), iz1 = __drcp_rn((double)Z[1]), iz2 = __drcp_rn((double)Z[2]);
  const double darea = (double)area;
  for (long long y = y0; y <= y1; ++y)
    for (long long x = x0; x <= x1; ++x) {
      const long long px = x * SUBPX + SUBPX / 2, py = y * SUBPX + SUBPX / 2;
      const long long e0 = (SX[2] - SX[1]) * (py - SY[1]) - (SY[2] - SY[1]) * (px - SX[1]);
      const long long e1 = (SX[0] - SX[2]) * (py - SY[2]) - (SY[0] - SY[2]) * (px - SX[2]);
      const long long e2 = (SX[1] - SX[0]) * (py - SY[0]) - (SY[1] - SY[0]) * (px - SX[0]);
      if (e0 < 0 || e1 < 0 || e2 < 0) continue;
      const double iz = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn((double)e0, iz0), __dmul_rn((double)e1, iz1)), __dmul_rn((double)e2, iz2)), darea);
      const float z = (float)__drcp_rn(iz);
      if (!(z > a.p.z_near && z < a.p.z_far)) continue;
      atomicMin(a.zbuf + t.off + (size_t)(y - t.y0) * t.w + (x - t.x0), __float_as_uint(z));
    }
}

__device__ __forceinline__ float sim_of(float z, float z_far) {
  if (!(z < FLT_MAX)) return z_far;
  float s = __fdiv_rn((float)(int)roundf(__fmul_rn(1000.f, z)), 1000.0f);
  if (s > 2.0f) s = 2.0f;
  if (s < 0.1f) s = 0.1f;
  return s;
}
__device__ __forceinline__ float diff_of(float sim, float real) {   // PoseEstimator.cpp:410-423; the literals are doubles
  if ((double)real <= 0.1 || (double)real >= 2.0) return 2.0f;
  if ((double)sim <= 0.1 || (double)sim >= 2.0) return 2.0f;
  return fabsf(__fsub_rn(sim, real));
}

__global__ void fill_kernel(unsigned int *z, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) z[i] = Z_EMPTY;
}

__global__ void base_diff_kernel(const float *zhand, const float *real, int n, float z_far, float *base_diff) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) base_diff[i] = diff_of(sim_of(zhand[i], z_far), real[i]);
}

// the reference's running sum is sequential by definition: one thread (once per frame; the loads do not depend on the sum)
__global__ void prefix_kernel(const float *base_diff, int n, float *prefix) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  float s = 0.f;
  prefix[0] = 0.f;
  for (int i = 0; i < n; ++i) { s = __fadd_rn(s, base_diff[i]); prefix[i + 1] = s; }
}

// bounding tile of the object under each hypothesis (pixels whose centre any projected vertex can reach, clamped to the image)
struct BboxArgs { hop_render_params p; const float *V; int nv; const float *poses; int H; Tile *tiles; long long *area; };
__global__ void __launch_bounds__(128) bbox_kernel(BboxArgs a) {
  __shared__ long long s_mn[2][4], s_mx[2][4];
  const int h = blockIdx.x;
  RasterArgs ra; ra.p = a.p;
  const float *T = a.poses + 16 * (size_t)h;
  long long mnx = LLONG_MAX, mny = LLONG_MAX, mxx = LLONG_MIN, mxy = LLONG_MIN;
  for (int v = threadIdx.x; v < a.nv; v += blockDim.x) {
    float Z; long long SX, SY;
    if (!project(ra, T, a.V + 3 * (size_t)v, Z, SX, SY)) continue;
    mnx = min(mnx, SX); mxx = max(mxx, SX); mny = min(mny, SY); mxy = max(mxy, SY);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
    mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o)); mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { s_mn[0][w] = mnx; s_mn[1][w] = mny; s_mx[0][w] = mxx; s_mx[1][w] = mxy; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 4; ++k) { mnx = min(mnx, s_mn[0][k]); mny = min(mny, s_mn[1][k]); mxx = max(mxx, s_mx[0][k]); mxy = max(mxy, s_mx[1][k]); }
    Tile t; t.x0 = t.y0 = 0; t.w = t.h = 0; t.off = 0;
    if (mnx <= mxx) {
      const long long W = a.p.width, Hh = a.p.height;
      long long x0 = mnx - SUBPX / 2 < 0 ? 0 : (mnx - SUBPX / 2 + SUBPX - 1) / SUBPX, x1 = (mxx - SUBPX / 2) / SUBPX;
      long long y0 = mny - SUBPX / 2 < 0 ? 0 : (mny - SUBPX / 2 + SUBPX - 1) / SUBPX, y1 = (mxy - SUBPX / 2) / SUBPX;
      x1 = min(x1, W - 1); y1 = min(y1, Hh - 1);
      if (mxx - SUBPX / 2 >= 0 && mxy - SUBPX / 2 >= 0 && x0 <= x1 && y0 <= y1) { t.x0 = (int)x0; t.y0 = (int)y0; t.w = (int)(x1 - x0 + 1); t.h = (int)(y1 - y0 + 1); }
    }
    a.tiles[h] = t;
    a.area[h] = (long long)t.w * t.h;
  }
}

__global__ void tile_offsets_kernel(Tile *tiles, const long long *off, int H) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h < H) tiles[h].off = off[h];
}

// one thread per hypothesis replays the reference's comparison loop (PoseEstimator.cpp:399-443) from the first row of its tile
struct WalkArgs {
  hop_render_params p;
  const float *real, *zhand, *base_diff, *prefix;
  const Tile *tiles; const unsigned int *zbuf;
  int H;
  float *wrong_ratio;
};
__global__ void __launch_bounds__(64) walk_kernel(WalkArgs a) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= a.H) return;
  const Tile t = a.tiles[h];
  const int W = a.p.width, Hh = a.p.height, n = W * Hh;
  float roi = 0.f, bg;
  int roi_cnt = 0;
  if (t.w <= 0 || t.h <= 0) bg = a.prefix[n];
  else {
    bg = a.prefix[t.y0 * W];
    const unsigned int *zt = a.zbuf + t.off;
    for (int y = t.y0; y < t.y0 + t.h; ++y) {
      const int row = y * W;
      for (int x = 0; x < t.x0; ++x) bg = __fadd_rn(bg, a.base_diff[row + x]);
      for (int x = t.x0; x < t.x0 + t.w; ++x) {
        const float zo = __uint_as_float(zt[(size_t)(y - t.y0) * t.w + (x - t.x0)]);
        if (zo < a.zhand[row + x]) { roi = __fadd_rn(roi, diff_of(sim_of(zo, a.p.z_far), a.real[row + x])); ++roi_cnt; }
        else bg = __fadd_rn(bg, a.base_diff[row + x]);
      }
      for (int x = t.x0 + t.w; x < W; ++x) bg = __fadd_rn(bg, a.base_diff[row + x]);
    }
    for (int i = (t.y0 + t.h) * W; i < n; ++i) bg = __fadd_rn(bg, a.base_diff[i]);
  }
  const int bg_cnt = n - roi_cnt;
  // float diff_total = roi_weight * roi_diff / roi_cnt + bg_diff / bg_cnt  (0 / 0 = NaN when the object owns no pixel)
  a.wrong_ratio[h] = __fadd_rn(__fdiv_rn(__fmul_rn(a.p.roi_weight, roi), (float)roi_cnt), __fdiv_rn(bg, (float)bg_cnt));
}

__global__ void compose_kernel(const unsigned int *zobj, const float *zhand, int n, float z_far, float *depth, unsigned char *mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float zo = __uint_as_float(zobj[i]);
  const bool ob = zo < zhand[i];
  depth[i] = sim_of(ob ? zo : zhand[i], z_far);
  if (mask) mask[i] = ob ? 1 : 0;
}

int check_params(hop_ctx *ctx, const hop_render_params *p) {
  if (!p || p->width <= 0 || p->height <= 0 || (long long)p->width * p->height > (1ll << 26) || !(p->fx > 0.f) || !(p->fy > 0.f) || !(p->z_near > 0.f) ||
      !(p->z_far > p->z_near)) { ctx->err = "hop_render: bad camera parameters"; return HOP_EINVAL; }
  return HOP_OK;
}

}  // namespace

extern "C" void hop_default_render_params(hop_render_params *p) {
  if (!p) return;
  p->fx = p->fy = 616.596f; p->cx = 307.628f; p->cy = 239.687f;   // config_autodataset.yaml:2
  p->width = 640; p->height = 480;                                 // PoseEstimator.cpp:349
  p->z_near = 0.1f; p->z_far = 2.0f;                               // simulation_io.cpp:425-426
  p->roi_weight = 2.0f; p->keep_ratio = 0.3f;                      // config_autodataset.yaml:121-122
}

extern "C" int hop_render_scene_create(hop_ctx *ctx, const hop_render_params *params, const float *depth_m, const float *hand_V, int hand_nv,
                                       const int32_t *hand_F, int hand_nf, hop_render_scene **out) {
  if (!ctx) return HOP_EINVAL;
  if (!out || !depth_m || hand_nv < 0 || hand_nf < 0 || (hand_nf > 0 && (!hand_V || !hand_F))) { ctx->err = "hop_render_scene_create: bad arguments"; return HOP_EINVAL; }
  int rc = check_params(ctx, params);
  if (rc != HOP_OK) return rc;
  for (int k = 0; k < 3 * hand_nf; ++k) if (hand_F[k] < 0 || hand_F[k] >= hand_nv) { ctx->err = "hop_render_scene_create: face index out of range"; return HOP_EINVAL; }
  hop_render_scene *s = new hop_render_scene();
  s->p = *params; s->n_px = params->width * params->height;
  const size_t nb = sizeof(float) * (size_t)s->n_px;
  cudaStream_t st = ctx->stream;
  float *d_V = nullptr; int32_t *d_F = nullptr;
  auto fail = [&](int code, const char *msg) { ctx->err = msg; cudaFree(d_V); cudaFree(d_F); cudaFree(s->d_real); cudaFree(s->d_zhand); cudaFree(s->d_base_diff); cudaFree(s->d_prefix); delete s; return code; };
  if (cudaMalloc(&s->d_real, nb) != cudaSuccess || cudaMalloc(&s->d_zhand, nb) != cudaSuccess || cudaMalloc(&s->d_base_diff, nb) != cudaSuccess ||
      cudaMalloc(&s->d_prefix, nb + sizeof(float)) != cudaSuccess) return fail(HOP_ENOMEM, "hop_render_scene_create: allocation failed");
  if (cudaMemcpyAsync(s->d_real, depth_m, nb, cudaMemcpyHostToDevice, st) != cudaSuccess) return fail(HOP_ECUDA, "hop_render_scene_create: copy failed");
  fill_kernel<<<296, 256, 0, st>>>((unsigned int *)s->d_zhand, s->n_px);
  ctx->launches += 1;
  if (hand_nf > 0) {
    if (cudaMalloc(&d_V, sizeof(float) * 3 * (size_t)hand_nv) != cudaSuccess || cudaMalloc(&d_F, sizeof(int32_t) * 3 * (size_t)hand_nf) != cudaSuccess)
      return fail(HOP_ENOMEM, "hop_render_scene_create: allocation failed");
    cudaMemcpyAsync(d_V, hand_V, sizeof(float) * 3 * (size_t)hand_nv, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_F, hand_F, sizeof(int32_t) * 3 * (size_t)hand_nf, cudaMemcpyHostToDevice, st);
    RasterArgs ra; ra.p = s->p; ra.V = d_V; ra.F = d_F; ra.nf = hand_nf; ra.poses = nullptr; ra.tiles = nullptr; ra.zbuf = (unsigned int *)s->d_zhand; ra.H = 1;
    raster_kernel<<<(hand_nf + 127) / 128, 128, 0, st>>>(ra);
    ctx->launches += 1;
  }
  base_diff_kernel<<<(s->n_px + 255) / 256, 256, 0, st>>>(s->d_zhand, s->d_real, s->n_px, s->p.z_far, s->d_base_diff);
  prefix_kernel<<<1, 32, 0, st>>>(s->d_base_diff, s->n_px, s->d_prefix);
  ctx->launches += 2;
  const cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(d_V); cudaFree(d_F); d_V = nullptr; d_F = nullptr;
  if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) return fail(HOP_ECUDA, "hop_render_scene_create: kernel failed");
  *out = s;
  return HOP_OK;
}

extern "C" int hop_render_scene_destroy(hop_ctx *ctx, hop_render_scene *s) {
  if (!s) return HOP_OK;
  if (ctx) cudaStreamSynchronize(ctx->stream);
  cudaFree(s->d_real); cudaFree(s->d_zhand); cudaFree(s->d_base_diff); cudaFree(s->d_prefix);
  delete s;
  return HOP_OK;
}

namespace {
// uploads the object mesh into ctx scratch: returns device V, F
int upload_object(hop_ctx *ctx, const float *V, int nv, const int32_t *F, int nf, float **d_V, int32_t **d_F) {
  if (!V || !F || nv < 3 || nf < 1) { ctx->err = "hop_render: bad object mesh"; return HOP_EINVAL; }
  for (int k = 0; k < 3 * nf; ++k) if (F[k] < 0 || F[k] >= nv) { ctx->err = "hop_render: face index out of range"; return HOP_EINVAL; }
  const size_t vb = (sizeof(float) * 3 * (size_t)nv + 255) / 256 * 256, fb = sizeof(int32_t) * 3 * (size_t)nf;
  char *d = (char *)ctx->ensure_scratch(vb + fb);
  if (!d) { ctx->err = "hop_render: scratch allocation failed"; return HOP_ENOMEM; }
  *d_V = (float *)d; *d_F = (int32_t *)(d + vb);
  HOP_CUDA(ctx, cudaMemcpyAsync(*d_V, V, sizeof(float) * 3 * (size_t)nv, cudaMemcpyHostToDevice, ctx->stream));
  HOP_CUDA(ctx, cudaMemcpyAsync(*d_F, F, fb, cudaMemcpyHostToDevice, ctx->stream));
  return HOP_OK;
}
}  // namespace

extern "C" int hop_render_depth(hop_ctx *ctx, const hop_render_scene *scene, const float *obj_V, int obj_nv, const int32_t *obj_F, int obj_nf,
                                const float *pose, float *depth, uint8_t *mask) {
  if (!ctx) return HOP_EINVAL;
  if (!scene || !pose || !depth) { ctx->err = "hop_render_depth: bad arguments"; return HOP_EINVAL; }
  float *d_V; int32_t *d_F;
  int rc = upload_object(ctx, obj_V, obj_nv, obj_F, obj_nf, &d_V, &d_F);
  if (rc != HOP_OK) return rc;
  cudaStream_t st = ctx->stream;
  const int n = scene->n_px;
  const size_t zb = (sizeof(unsigned int) * (size_t)n + 255) / 256 * 256, db = (sizeof(float) * (size_t)n + 255) / 256 * 256;
  char *d = (char *)ctx->ensure_io(256 + zb + db + n);
  if (!d) { ctx->err = "hop_render_depth: staging allocation failed"; return HOP_ENOMEM; }
  float *d_pose = (float *)d; unsigned int *d_z = (unsigned int *)(d + 256); float *d_depth = (float *)(d + 256 + zb); unsigned char *d_mask = (unsigned char *)(d + 256 + zb + db);
  HOP_CUDA(ctx, cudaMemcpyAsync(d_pose, pose, 64, cudaMemcpyHostToDevice, st));
  fill_kernel<<<296, 256, 0, st>>>(d_z, n);
  RasterArgs ra; ra.p = scene->p; ra.V = d_V; ra.F = d_F; ra.nf = obj_nf; ra.poses = d_pose; ra.tiles = nullptr; ra.zbuf = d_z; ra.H = 1;
  raster_kernel<<<(obj_nf + 127) / 128, 128, 0, st>>>(ra);
  compose_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_z, scene->d_zhand, n, scene->p.z_far, d_depth, d_mask);
  ctx->launches += 3;
  HOP_CUDA(ctx, cudaGetLastError());
  HOP_CUDA(ctx, cudaMemcpyAsync(depth, d_depth, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, st));
  if (mask) HOP_CUDA(ctx, cudaMemcpyAsync(mask, d_mask, (size_t)n, cudaMemcpyDeviceToHost, st));
  HOP_CUDA(ctx, cudaStreamSynchronize(st));
  return HOP_OK;
}

extern "C" int hop_reject_by_render(hop_ctx *ctx, const hop_render_scene *scene, const float *obj_V, int obj_nv, const int32_t *obj_F, int obj_nf,
                                    const float *poses, int H, float *wrong_ratio, int32_t *order, int32_t *n_keep) {
  if (!ctx) return HOP_EINVAL;
  if (!scene || H < 0 || (H > 0 && (!poses || !wrong_ratio))) { ctx->err = "hop_reject_by_render: bad arguments"; return HOP_EINVAL; }
  if (n_keep) *n_keep = 0;
  if (H == 0) return HOP_OK;
  float *d_V; int32_t *d_F;
  int rc = upload_object(ctx, obj_V, obj_nv, obj_F, obj_nf, &d_V, &d_F);
  if (rc != HOP_OK) return rc;
  cudaStream_t st = ctx->stream;
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t pb = up(sizeof(float) * 16 * (size_t)H), tb = up(sizeof(Tile) * (size_t)H), ab = up(sizeof(long long) * (size_t)H), wb = up(sizeof(float) * (size_t)H);
  char *d = (char *)ctx->ensure_io(pb + tb + 2 * ab + wb);
  if (!d) { ctx->err = "hop_reject_by_render: staging allocation failed"; return HOP_ENOMEM; }
  float *d_poses = (float *)d; Tile *d_tiles = (Tile *)(d + pb); long long *d_area = (long long *)(d + pb + tb), *d_off = (long long *)(d + pb + tb + ab);
  float *d_wr = (float *)(d + pb + tb + 2 * ab);
  HOP_CUDA(ctx, cudaMemcpyAsync(d_poses, poses, sizeof(float) * 16 * (size_t)H, cudaMemcpyHostToDevice, st));
  ProfScope ps(ctx, HOP_PROF_RENDER);
  BboxArgs ba; ba.p = scene->p; ba.V = d_V; ba.nv = obj_nv; ba.poses = d_poses; ba.H = H; ba.tiles = d_tiles; ba.area = d_area;
  bbox_kernel<<<H, 128, 0, st>>>(ba);
  ctx->launches += 1;
  // tile offsets: a host scan of H numbers (the arena is sized from their sum)
  std::vector<long long> area(H), off(H);
  HOP_CUDA(ctx, cudaMemcpyAsync(area.data(), d_area, sizeof(long long) * (size_t)H, cudaMemcpyDeviceToHost, st));
  HOP_CUDA(ctx, cudaStreamSynchronize(st));
  long long total = 0;
  for (int h = 0; h < H; ++h) { off[h] = total; total += area[h]; }
  unsigned int *d_z = (unsigned int *)ctx->ensure_work(sizeof(unsigned int) * (size_t)std::max<long long>(total, 1));
  if (!d_z) { ctx->err = "hop_reject_by_render: tile arena allocation failed"; return HOP_ENOMEM; }
  HOP_CUDA(ctx, cudaMemcpyAsync(d_off, off.data(), sizeof(long long) * (size_t)H, cudaMemcpyHostToDevice, st));
  tile_offsets_kernel<<<(H + 127) / 128, 128, 0, st>>>(d_tiles, d_off, H);
  fill_kernel<<<592, 256, 0, st>>>(d_z, total);
  RasterArgs ra; ra.p = scene->p; ra.V = d_V; ra.F = d_F; ra.nf = obj_nf; ra.poses = d_poses; ra.tiles = d_tiles; ra.zbuf = d_z; ra.H = H;
  const long long work = (long long)H * obj_nf;
  raster_kernel<<<(unsigned int)((work + 127) / 128), 128, 0, st>>>(ra);
  WalkArgs wa; wa.p = scene->p; wa.real = scene->d_real; wa.zhand = scene->d_zhand; wa.base_diff = scene->d_base_diff; wa.prefix = scene->d_prefix;
  wa.tiles = d_tiles; wa.zbuf = d_z; wa.H = H; wa.wrong_ratio = d_wr;
  walk_kernel<<<(H + 63) / 64, 64, 0, st>>>(wa);
  ctx->launches += 4;
  HOP_CUDA(ctx, cudaGetLastError());
  HOP_CUDA(ctx, cudaMemcpyAsync(wrong_ratio, d_wr, sizeof(float) * (size_t)H, cudaMemcpyDeviceToHost, st));
  HOP_CUDA(ctx, cudaStreamSynchronize(st));
  if (order) {
    // the reference pops a priority queue ordered by _wrong_ratio (ascending); ties and NaN are unspecified there: index order, NaN last
    int keep = (int)(scene->p.keep_ratio * H);
    keep = std::min(std::max(keep, 10), H);
    std::vector<int32_t> idx(H);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) {
      const float wa2 = wrong_ratio[a], wb2 = wrong_ratio[b];
      if (wa2 != wa2) return false;
      if (wb2 != wb2) return true;
      return wa2 < wb2;
    });
    std::copy(idx.begin(), idx.begin() + keep, order);
    if (n_keep) *n_keep = keep;
  }
  return HOP_OK;
}
