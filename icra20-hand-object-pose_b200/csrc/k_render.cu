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
// hypotheses differently).  Coverage is decided in integer arithmetic on a 1/256 sub-pixel grid and depth as a fused double dot product + one float reciprocal,
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
  bool prefix_ready = false;     // the running sum is built on first use by a batch large enough to need it
  float *d_prefix = nullptr;     // n_px + 1 slots; [y * width] = running float sum of d_base_diff over the pixels before row y, [n_px] = the total
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

// 4-byte asynchronous global -> shared copy (LDGSTS): the producers put a whole row in flight at once, no register staging
__device__ __forceinline__ void cp_async4(float *dst_smem, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

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
  // 1/Z is affine in window space: per-triangle weights w_k = (1/Z_k) / area, per pixel one fused dot product and one float reciprocal
  const double darea = (double)area;
  const double w0 = __ddiv_rn(__drcp_rn((double)Z[0]), darea), w1 = __ddiv_rn(__drcp_rn((double)Z[1]), darea), w2 = __ddiv_rn(__drcp_rn((double)Z[2]), darea);
  if (x0 > x1 || y0 > y1) return;
  // edge functions at the first pixel centre, then stepped: +1 pixel in x adds -(dy) * SUBPX, +1 pixel in y adds (dx) * SUBPX (exact integers)
  const long long px0 = x0 * SUBPX + SUBPX / 2, py0 = y0 * SUBPX + SUBPX / 2;
  const long long a0 = -(SY[2] - SY[1]) * SUBPX, b0 = (SX[2] - SX[1]) * SUBPX;
  const long long a1 = -(SY[0] - SY[2]) * SUBPX, b1 = (SX[0] - SX[2]) * SUBPX;
  const long long a2 = -(SY[1] - SY[0]) * SUBPX, b2 = (SX[1] - SX[0]) * SUBPX;
  long long r0 = (SX[2] - SX[1]) * (py0 - SY[1]) - (SY[2] - SY[1]) * (px0 - SX[1]);
  long long r1 = (SX[0] - SX[2]) * (py0 - SY[2]) - (SY[0] - SY[2]) * (px0 - SX[2]);
  long long r2 = (SX[1] - SX[0]) * (py0 - SY[0]) - (SY[1] - SY[0]) * (px0 - SX[0]);
  for (long long y = y0; y <= y1; ++y, r0 += b0, r1 += b1, r2 += b2) {
    long long e0 = r0, e1 = r1, e2 = r2;
    for (long long x = x0; x <= x1; ++x, e0 += a0, e1 += a1, e2 += a2) {
      if (e0 < 0 || e1 < 0 || e2 < 0) continue;
      const double iz = __fma_rn((double)e2, w2, __fma_rn((double)e1, w1, __dmul_rn((double)e0, w0)));
      const float z = __fdiv_rn(1.0f, (float)iz);
      if (!(z > a.p.z_near && z < a.p.z_far)) continue;
      atomicMin(a.zbuf + t.off + (size_t)(y - t.y0) * t.w + (x - t.x0), __float_as_uint(z));
    }
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

// The reference's running sum is sequential by definition: thread 0 adds, row by row out of shared memory, while warps 1-3 fetch
// the next row (once per frame).  Only the sums at the row starts (and the total) are ever used.
constexpr int PREFIX_THREADS = 128;
__global__ void __launch_bounds__(PREFIX_THREADS) prefix_kernel(const float *base_diff, int W, int Hh, float *prefix) {
  extern __shared__ __align__(16) float prefix_smem[];   // 2 rows, padded to a multiple of 32 with zeros (x + 0 = x)
  const int Wp = (W + 31) & ~31;
  const int tid = threadIdx.x;
  auto produce = [&](int y) {
    if (y >= Hh) return;
    float *dst = prefix_smem + (y & 1) * Wp;
    for (int x = tid - 32; x < Wp; x += PREFIX_THREADS - 32) { if (x < W) cp_async4(dst + x, base_diff + (size_t)y * W + x); else dst[x] = 0.f; }
    cp_async_wait_all();
  };
  if (tid >= 32) produce(0);
  float s = 0.f;
  for (int y = 0; y < Hh; ++y) {
    __syncthreads();
    if (tid >= 32) { produce(y + 1); continue; }
    if (tid != 0) continue;
    prefix[(size_t)y * W] = s;
    const float4 *r4 = reinterpret_cast<const float4 *>(prefix_smem + (y & 1) * Wp);
    for (int x = 0; x < Wp; x += 32) {
      float4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = r4[(x >> 2) + k];
#pragma unroll
      for (int k = 0; k < 8; ++k) { s = __fadd_rn(s, v[k].x); s = __fadd_rn(s, v[k].y); s = __fadd_rn(s, v[k].z); s = __fadd_rn(s, v[k].w); }
    }
  }
  if (tid == 0) prefix[(size_t)W * Hh] = s;
}

// bounding tile of the object under each hypothesis (pixels whose centre any projected vertex can reach, clamped to the image)
struct BboxArgs { hop_render_params p; const float *V; int nv; const float *poses; int H; Tile *tiles; };
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
  }
}

// per tile pixel, in place: the float bits of Z become the pixel's contribution when the object owns it (its difference to the
// real image, >= 0) or -1 when it does not (empty, or behind the hand) -- everything the walk needs besides the shared image
struct ResolveArgs { hop_render_params p; const float *real, *zhand; const Tile *tiles; unsigned int *zbuf; int H; };
__global__ void __launch_bounds__(256) resolve_kernel(ResolveArgs a) {
  const int h = blockIdx.y;
  const Tile t = a.tiles[h];
  const long long n = (long long)t.w * t.h;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = t.x0 + (int)(i % t.w), y = t.y0 + (int)(i / t.w);
    const float zo = __uint_as_float(a.zbuf[t.off + i]);
    float v = -1.f;
    if (zo < FLT_MAX && zo < a.zhand[(size_t)y * a.p.width + x]) v = diff_of(sim_of(zo, a.p.z_far), a.real[(size_t)y * a.p.width + x]);
    a.zbuf[t.off + i] = __float_as_uint(v);
  }
}

// The reference's comparison loop (PoseEstimator.cpp:399-443) for 32 hypotheses per CTA.  Lane l of warp 0 carries hypothesis l's
// two running float sums add by add in row-major order (the background sum starts from the shared running sum at the first row the
// hypothesis' tile touches); the 32 lanes walk the image in lock step from the first row any of them needs.  Warps 1-3 are
// producers: while warp 0 adds row y out of shared memory they fetch row y + 1 (the shared difference row and every lane's tile
// segment) into the other buffer, so the sequential adds never wait on global memory.
struct WalkArgs {
  hop_render_params p;
  const float *base_diff, *prefix;
  const Tile *tiles; const unsigned int *zbuf;   // resolved tiles
  const int *perm; int n;                        // the hypotheses of this launch (grouped by tile width, ordered by first row)
  float *wrong_ratio;
};
constexpr int WALK_THREADS = 128;
__global__ void __launch_bounds__(WALK_THREADS) walk_kernel(WalkArgs a, int tile_stride, int lanes /* hypotheses per CTA, <= 32 */,
                                                            int from_top /* 1: no shared running sum, every lane starts at row 0 with 0 */) {
  extern __shared__ __align__(16) float walk_smem[];
  __shared__ Tile s_t[32];
  __shared__ int s_first, s_ux0, s_ux1;
  const int W = a.p.width, Hh = a.p.height, n = W * Hh;
  const int Wp = (W + 31) & ~31;                   // rows padded with zeros to a multiple of 32 (x + 0 = x)
  float *s_base = walk_smem;                       // 2 x Wp
  float *s_tile = walk_smem + 2 * Wp;              // 2 x 32 x tile_stride
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    const int slot = blockIdx.x * lanes + lane;
    Tile t; t.x0 = t.y0 = 0; t.w = t.h = 0; t.off = 0;
    if (lane < lanes && slot < a.n) t = a.tiles[a.perm[slot]];
    if (t.w <= 0 || t.h <= 0) { t.w = t.h = 0; t.y0 = Hh; }
    s_t[lane] = t;
    int first = t.y0, ux0 = t.w > 0 ? t.x0 : Wp, ux1 = t.w > 0 ? t.x0 + t.w : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
      ux0 = min(ux0, __shfl_xor_sync(0xffffffffu, ux0, o)); ux1 = max(ux1, __shfl_xor_sync(0xffffffffu, ux1, o));
    }
    if (lane == 0) { s_first = from_top ? 0 : first; s_ux0 = ux0 & ~31; s_ux1 = min(Wp, (ux1 + 31) & ~31); }   // the columns any tile of this CTA touches
  }
  __syncthreads();
  const int first = s_first, ux0 = s_ux0, ux1 = s_ux1;
  // producer: row y -> buffer (y & 1)
  auto produce = [&](int y) {
    if (y >= Hh) return;
    float *db = s_base + (y & 1) * Wp, *dt = s_tile + (size_t)(y & 1) * lanes * tile_stride;
    const float *src = a.base_diff + (size_t)y * W;
    for (int x = tid - 32; x < Wp; x += WALK_THREADS - 32) { if (x < W) cp_async4(db + x, src + x); else db[x] = 0.f; }
    for (int l = 0; l < lanes; ++l) {
      const Tile t = s_t[l];
      if (y < t.y0 || y >= t.y0 + t.h) continue;
      const unsigned int *zr = a.zbuf + t.off + (size_t)(y - t.y0) * t.w;
      for (int x = tid - 32; x < t.w; x += WALK_THREADS - 32) cp_async4(dt + l * tile_stride + x, zr + x);
    }
    cp_async_wait_all();
  };
  if (warp != 0) produce(first);
  const Tile t = s_t[lane];
  float roi = 0.f, bg = 0.f;
  int roi_cnt = 0;
  for (int y = first; y < Hh; ++y) {
    __syncthreads();                               // row y is in its buffer; the other buffer is free
    if (warp != 0) { produce(y + 1); continue; }
    const float *sb = s_base + (y & 1) * Wp;
    const float *st = s_tile + (size_t)(y & 1) * lanes * tile_stride + (lane < lanes ? lane : 0) * tile_stride;   // this lane's tile segment of row y
    if (!from_top && y == t.y0) bg = a.prefix[(size_t)y * W];
    const bool live_lane = lane < lanes && blockIdx.x * lanes + lane < a.n;
    const bool started = from_top ? live_lane : y >= t.y0, in_rows = y >= t.y0 && y < t.y0 + t.h;
    // the shared differences of columns [c0, c1) (multiples of 32), added in order by every lane that has started
    auto plain = [&](int c0, int c1) {
      if (!started) return;
      const float4 *r4 = reinterpret_cast<const float4 *>(sb);
      for (int x = c0; x < c1; x += 32) {
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = r4[(x >> 2) + k];
#pragma unroll
        for (int k = 0; k < 8; ++k) { bg = __fadd_rn(bg, v[k].x); bg = __fadd_rn(bg, v[k].y); bg = __fadd_rn(bg, v[k].z); bg = __fadd_rn(bg, v[k].w); }
      }
    };
    // x + 0.f = x exactly for the non-negative sums carried here, so "not mine" adds a zero: two independent add chains per pixel,
    // every shared-memory read of a group of 32 (8) pixels issued ahead of its adds
    if (!__any_sync(0xffffffffu, in_rows)) {
      plain(0, Wp);                              // the row is padded to a multiple of 32 with zeros
    } else {
      // only the columns some tile of this CTA touches need the per-pixel choice between the object's and the shared difference
      plain(0, ux0);
      const int xa = in_rows ? t.x0 : Wp, xb = in_rows ? t.x0 + t.w : Wp;
      for (int x0 = ux0; x0 < ux1; x0 += 8) {
        const float4 b0 = *reinterpret_cast<const float4 *>(sb + x0), b1 = *reinterpret_cast<const float4 *>(sb + x0 + 4);
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { const int x = x0 + k; v[k] = (x >= xa && x < xb) ? st[x - xa] : -1.f; }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const bool hit = !(v[k] < 0.f);   // (a NaN difference -- NaN in the real image -- still belongs to the object)
          roi = __fadd_rn(roi, hit ? v[k] : 0.f);
          roi_cnt += hit;
          bg = __fadd_rn(bg, (started && !hit) ? b[k] : 0.f);
        }
      }
      plain(ux1, Wp);
    }
  }
  if (warp != 0) return;
  const int slot = blockIdx.x * lanes + lane;
  if (lane >= lanes || slot >= a.n) return;
  const int h = a.perm[slot];
  if (!from_top && t.w == 0) bg = a.prefix[n];
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
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  if (!out || !depth_m || hand_nv < 0 || hand_nf < 0 || (hand_nf > 0 && (!hand_V || !hand_F))) { ctx->err = "hop_render_scene_create: bad arguments"; return HOP_EINVAL; }
  int rc = check_params(ctx, params);
  if (rc != HOP_OK) return rc;
  for (int k = 0; k < 3 * hand_nf; ++k) if (hand_F[k] < 0 || hand_F[k] >= hand_nv) { ctx->err = "hop_render_scene_create: face index out of range"; return HOP_EINVAL; }
  hop_render_scene *s = new hop_render_scene();
  s->p = *params; s->n_px = params->width * params->height;
  const size_t nb = sizeof(float) * (size_t)s->n_px;
  cudaStream_t st = ctx->stream;
  // per-frame object: every buffer comes from the context's stream-ordered pool (cudaMalloc / cudaFree cost milliseconds per frame)
  float *d_V = nullptr; int32_t *d_F = nullptr;
  auto release = [&]() {
    if (d_V) cudaFreeAsync(d_V, st); if (d_F) cudaFreeAsync(d_F, st);
    d_V = nullptr; d_F = nullptr;
  };
  auto fail = [&](int code, const char *msg) {
    ctx->err = msg; release();
    if (s->d_real) cudaFreeAsync(s->d_real, st); if (s->d_zhand) cudaFreeAsync(s->d_zhand, st);
    if (s->d_base_diff) cudaFreeAsync(s->d_base_diff, st); if (s->d_prefix) cudaFreeAsync(s->d_prefix, st);
    delete s; return code;
  };
  if (cudaMallocAsync(&s->d_real, nb, st) != cudaSuccess || cudaMallocAsync(&s->d_zhand, nb, st) != cudaSuccess ||
      cudaMallocAsync(&s->d_base_diff, nb, st) != cudaSuccess || cudaMallocAsync(&s->d_prefix, nb + sizeof(float), st) != cudaSuccess)
    return fail(HOP_ENOMEM, "hop_render_scene_create: allocation failed");
  if (cudaMemcpyAsync(s->d_real, depth_m, nb, cudaMemcpyHostToDevice, st) != cudaSuccess) return fail(HOP_ECUDA, "hop_render_scene_create: copy failed");
  fill_kernel<<<296, 256, 0, st>>>((unsigned int *)s->d_zhand, s->n_px);
  ctx->launches += 1;
  if (hand_nf > 0) {
    if (cudaMallocAsync(&d_V, sizeof(float) * 3 * (size_t)hand_nv, st) != cudaSuccess || cudaMallocAsync(&d_F, sizeof(int32_t) * 3 * (size_t)hand_nf, st) != cudaSuccess)
      return fail(HOP_ENOMEM, "hop_render_scene_create: allocation failed");
    cudaMemcpyAsync(d_V, hand_V, sizeof(float) * 3 * (size_t)hand_nv, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_F, hand_F, sizeof(int32_t) * 3 * (size_t)hand_nf, cudaMemcpyHostToDevice, st);
    RasterArgs ra; ra.p = s->p; ra.V = d_V; ra.F = d_F; ra.nf = hand_nf; ra.poses = nullptr; ra.tiles = nullptr; ra.zbuf = (unsigned int *)s->d_zhand; ra.H = 1;
    raster_kernel<<<(hand_nf + 127) / 128, 128, 0, st>>>(ra);
    ctx->launches += 1;
  }
  base_diff_kernel<<<(s->n_px + 255) / 256, 256, 0, st>>>(s->d_zhand, s->d_real, s->n_px, s->p.z_far, s->d_base_diff);
  ctx->launches += 1;
  const cudaError_t e = cudaStreamSynchronize(st);   // the host arrays (depth image, hand mesh) may be reused by the caller after the call
  release();
  if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) return fail(HOP_ECUDA, "hop_render_scene_create: kernel failed");
  *out = s;
  return HOP_OK;
}

extern "C" int hop_render_scene_destroy(hop_ctx *ctx, hop_render_scene *s) {
  HOP_ENTER(ctx);
  if (!s) return HOP_OK;
  cudaStream_t st = ctx ? ctx->stream : (cudaStream_t)0;   // stream-ordered: the frees queue behind the scene's last use on the context's stream
  cudaFreeAsync(s->d_real, st); cudaFreeAsync(s->d_zhand, st); cudaFreeAsync(s->d_base_diff, st); cudaFreeAsync(s->d_prefix, st);
  delete s;
  return HOP_OK;
}

namespace {
// uploads the object mesh into ctx scratch: returns device V, F
int upload_object(hop_ctx *ctx, const float *V, int nv, const int32_t *F, int nf, float **d_V, int32_t **d_F, size_t extra = 0, size_t *used = nullptr) {
  if (!V || !F || nv < 3 || nf < 1) { ctx->err = "hop_render: bad object mesh"; return HOP_EINVAL; }
  for (int k = 0; k < 3 * nf; ++k) if (F[k] < 0 || F[k] >= nv) { ctx->err = "hop_render: face index out of range"; return HOP_EINVAL; }
  const size_t vb = (sizeof(float) * 3 * (size_t)nv + 255) / 256 * 256, fb = sizeof(int32_t) * 3 * (size_t)nf;
  const size_t fbp = (fb + 255) / 256 * 256;
  char *d = (char *)ctx->ensure_scratch(vb + fbp + extra);
  if (used) *used = vb + fbp;
  if (!d) { ctx->err = "hop_render: scratch allocation failed"; return HOP_ENOMEM; }
  *d_V = (float *)d; *d_F = (int32_t *)(d + vb);
  HOP_CUDA(ctx, cudaMemcpyAsync(*d_V, V, sizeof(float) * 3 * (size_t)nv, cudaMemcpyHostToDevice, ctx->stream));
  HOP_CUDA(ctx, cudaMemcpyAsync(*d_F, F, fb, cudaMemcpyHostToDevice, ctx->stream));
  return HOP_OK;
}
}  // namespace

extern "C" int hop_render_depth(hop_ctx *ctx, const hop_render_scene *scene, const float *obj_V, int obj_nv, const int32_t *obj_F, int obj_nf,
                                const float *pose, float *depth, uint8_t *mask) {
  HOP_ENTER(ctx);
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
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  if (!scene || H < 0 || (H > 0 && (!poses || !wrong_ratio))) { ctx->err = "hop_reject_by_render: bad arguments"; return HOP_EINVAL; }
  if (n_keep) *n_keep = 0;
  if (H == 0) return HOP_OK;
  float *d_V; int32_t *d_F;
  size_t vb_obj = 0;
  int rc = upload_object(ctx, obj_V, obj_nv, obj_F, obj_nf, &d_V, &d_F, sizeof(int) * (size_t)H, &vb_obj);
  if (rc != HOP_OK) return rc;
  cudaStream_t st = ctx->stream;
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t pb = up(sizeof(float) * 16 * (size_t)H), tb = up(sizeof(Tile) * (size_t)H), wb = up(sizeof(float) * (size_t)H);
  char *d = (char *)ctx->ensure_io(pb + tb + wb);
  if (!d) { ctx->err = "hop_reject_by_render: staging allocation failed"; return HOP_ENOMEM; }
  float *d_poses = (float *)d; Tile *d_tiles = (Tile *)(d + pb);
  float *d_wr = (float *)(d + pb + tb);
  HOP_CUDA(ctx, cudaMemcpyAsync(d_poses, poses, sizeof(float) * 16 * (size_t)H, cudaMemcpyHostToDevice, st));
  ProfScope ps(ctx, HOP_PROF_RENDER);
  BboxArgs ba; ba.p = scene->p; ba.V = d_V; ba.nv = obj_nv; ba.poses = d_poses; ba.H = H; ba.tiles = d_tiles;
  bbox_kernel<<<H, 128, 0, st>>>(ba);
  ctx->launches += 1;
  // tile offsets: a host scan over the H tiles (the arena is sized from the sum of their areas)
  std::vector<Tile> tiles(H);
  HOP_CUDA(ctx, cudaMemcpyAsync(tiles.data(), d_tiles, sizeof(Tile) * (size_t)H, cudaMemcpyDeviceToHost, st));
  HOP_CUDA(ctx, cudaStreamSynchronize(st));
  long long total = 0, max_area = 1;
  int max_w = 1;
  for (int h = 0; h < H; ++h) {
    tiles[h].off = total;
    const long long ar = (long long)tiles[h].w * tiles[h].h;
    total += ar; max_area = std::max(max_area, ar); max_w = std::max(max_w, tiles[h].w);
  }
  unsigned int *d_z = (unsigned int *)ctx->ensure_work(sizeof(unsigned int) * (size_t)std::max<long long>(total, 1));
  if (!d_z) { ctx->err = "hop_reject_by_render: tile arena allocation failed"; return HOP_ENOMEM; }
  HOP_CUDA(ctx, cudaMemcpyAsync(d_tiles, tiles.data(), sizeof(Tile) * (size_t)H, cudaMemcpyHostToDevice, st));
  fill_kernel<<<592, 256, 0, st>>>(d_z, total);
  RasterArgs ra; ra.p = scene->p; ra.V = d_V; ra.F = d_F; ra.nf = obj_nf; ra.poses = d_poses; ra.tiles = d_tiles; ra.zbuf = d_z; ra.H = H;
  const long long work = (long long)H * obj_nf;
  raster_kernel<<<(unsigned int)((work + 127) / 128), 128, 0, st>>>(ra);
  {
    ResolveArgs rs; rs.p = scene->p; rs.real = scene->d_real; rs.zhand = scene->d_zhand; rs.tiles = d_tiles; rs.zbuf = d_z; rs.H = H;
    const int bx = (int)std::max<long long>(1, std::min<long long>(64, (max_area + 2047) / 2048));
    for (int h0 = 0; h0 < H; h0 += 65535) {
      ResolveArgs r2 = rs; r2.tiles += h0; r2.H = std::min(65535, H - h0);
      resolve_kernel<<<dim3(bx, r2.H), 256, 0, st>>>(r2);
    }
  }
  {
    // The shared-memory stride of a walk launch is its widest tile: one stray hypothesis that fills the image would cost every CTA
    // its occupancy, so the hypotheses are split into narrow tiles (<= 191 pixels: 4 CTAs per SM) and the rest, each group ordered by
    // first row so that the 32 lanes of a CTA start their walk together.
    std::vector<int> perm(H);
    std::iota(perm.begin(), perm.end(), 0);
    const int narrow = 191;
    std::stable_sort(perm.begin(), perm.end(), [&](int x, int y) {
      const bool wx = tiles[x].w > narrow, wy = tiles[y].w > narrow;
      if (wx != wy) return wy;
      const int fx = tiles[x].w > 0 ? tiles[x].y0 : scene->p.height, fy = tiles[y].w > 0 ? tiles[y].y0 : scene->p.height;
      return fx < fy;
    });
    int n_narrow = 0;
    for (int h = 0; h < H; ++h) n_narrow += tiles[h].w <= narrow;
    int *d_perm = (int *)ctx->ensure_scratch(vb_obj + sizeof(int) * (size_t)H) ;
    if (!d_perm) { ctx->err = "hop_reject_by_render: scratch allocation failed"; return HOP_ENOMEM; }
    d_perm = (int *)((char *)d_perm + vb_obj);
    HOP_CUDA(ctx, cudaMemcpyAsync(d_perm, perm.data(), sizeof(int) * (size_t)H, cudaMemcpyHostToDevice, st));
    const int Wp = (scene->p.width + 31) & ~31;
    // A handful of hypotheses (the reference's <= 100 after clustering): every CTA simply starts at row 0 -- cheaper than the
    // 0.7 ms sequential running sum of the whole frame, which pays off from a few CTAs' worth of hypotheses up.
    const int from_top = (H <= 64 && !scene->prefix_ready) ? 1 : 0;
    if (!from_top && !scene->prefix_ready) {
      hop_render_scene *sc = const_cast<hop_render_scene *>(scene);
      prefix_kernel<<<1, PREFIX_THREADS, sizeof(float) * 2 * (size_t)Wp, st>>>(sc->d_base_diff, sc->p.width, sc->p.height, sc->d_prefix);
      ctx->launches += 1;
      sc->prefix_ready = true;
    }
    for (int part = 0; part < 2; ++part) {
      const int begin = part == 0 ? 0 : n_narrow, count = part == 0 ? n_narrow : H - n_narrow;
      if (count <= 0) continue;
      const int stride = (part == 0 ? std::min(max_w, narrow) : max_w) | 1;   // odd: the 32 lanes' segments start in different banks
      WalkArgs wa; wa.p = scene->p; wa.base_diff = scene->d_base_diff; wa.prefix = scene->d_prefix;
      wa.tiles = d_tiles; wa.zbuf = d_z; wa.perm = d_perm + begin; wa.n = count; wa.wrong_ratio = d_wr;
      // 32 hypotheses per CTA when their tile rows fit the shared memory twice over, fewer for very wide tiles / images
      const size_t budget = 200 * 1024, base_bytes = sizeof(float) * 2 * (size_t)Wp, per_lane = sizeof(float) * 2 * (size_t)stride;
      if (base_bytes + per_lane > budget) { ctx->err = "hop_reject_by_render: image too wide for the walk kernel's shared memory"; return HOP_EINVAL; }
      const int lanes = (int)std::min<size_t>(32, (budget - base_bytes) / per_lane);
      const size_t smem = base_bytes + per_lane * (size_t)lanes;
      HOP_CUDA(ctx, ctx->func_smem_optin(walk_kernel, smem));
      walk_kernel<<<(count + lanes - 1) / lanes, WALK_THREADS, smem, st>>>(wa, stride, lanes, from_top);
      ctx->launches += 1;
    }
  }
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
