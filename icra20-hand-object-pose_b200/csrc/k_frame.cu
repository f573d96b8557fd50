// k_frame.cu -- the per-frame front end on the device: depth image -> the object-segment cloud the hot path consumes.
//
//   replaces the pre-processing of main_realdata_auto        src/perception/src/app/main_realdata_auto.cpp:54-96,144-181
//            Utils::readDepthImage / convert3dOrganized        src/perception/src/Utils.cpp:36-55,78-115
//            Utils::downsamplePointCloud (pcl::VoxelGrid)      src/perception/src/Utils.cpp:333-340
//            pcl::PassThrough x4, pcl::transformPointCloudWithNormals, the radius normals + flipNormalTowardsViewpoint
//   (needed so that a stream of depth frames -- C4: 128 of them -- reaches K2..K5 without a host detour.)
//
// Stage order and arithmetic are the host restatement's (icra20-hand-object-pose_b200/host/cloud.cpp, frameToObjectSegment),
// float operation by float operation (no FMA contraction), so everything up to the normals is BIT-identical to it:
//   1. back-projection of the valid pixels (0.1 m < z < 2 m), kept in raster order           (flags + cub::DeviceSelect)
//   2. VoxelGrid(leaf_dense): key = PCL's linear leaf index, stable radix sort, one thread per leaf sums ITS points in
//      their original order (= the order PCL's sorted index vector visits them), centroid = sum / count
//   3. camera -> hand-base frame, crop box, back to the camera frame (the float round trip included)
//   4. normals: PCA over the neighbours within normal_radius through a dense cell grid (cell = radius), covariance and
//      Jacobi eigen-solve in double like the host; the neighbour SET is identical (same float distance test), the double
//      sums run in a different order, so normals agree to ~1e-7, not to the bit
//   5. VoxelGrid(leaf_object) with normals (normalised mean normal), NaN removal, flip towards the camera, confidence = 1
// The result is written straight into a hop_cloud (padded float4 streams + bounding box): no host copy of the cloud exists.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include <cfloat>
#include <cmath>
#include <vector>

#include "hop_common.cuh"

int hop_cloud_reserve(hop_ctx *ctx, hop_cloud *c, int n);  // api.cu

namespace {

struct FBuf {   // stream-ordered scratch (see hop_create: the pool keeps freed blocks)
  void *p = nullptr;
  cudaStream_t st;
  explicit FBuf(cudaStream_t s) : st(s) {}
  FBuf(const FBuf &) = delete;
  FBuf &operator=(const FBuf &) = delete;
  ~FBuf() { if (p) cudaFreeAsync(p, st); }
  template <typename T> T *as() { return (T *)p; }
  cudaError_t alloc(size_t bytes) {
    if (p) cudaFreeAsync(p, st);
    p = nullptr;
    return cudaMallocAsync(&p, bytes < 16 ? 16 : bytes, st);
  }
};

__device__ __forceinline__ void atomic_min_f(float *a, float v) {
  if (v >= 0.f) atomicMin((int *)a, __float_as_int(v)); else atomicMax((unsigned int *)a, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float *a, float v) {
  if (v >= 0.f) atomicMax((int *)a, __float_as_int(v)); else atomicMin((unsigned int *)a, __float_as_uint(v));
}
__device__ __forceinline__ bool finite3(float x, float y, float z) { return isfinite(x) && isfinite(y) && isfinite(z); }

// Utils::readDepthImage + convert3dOrganized + PassThrough(z, 0.1, 2.0): valid pixel -> (x, y, z, confidence 1)
// bounds[0..2] = min, [3..5] = max over the points a stage KEEPS, accumulated by the stage's own kernel (all 32 lanes call): the next
// stage's voxel geometry then comes back with the kept count in one round trip instead of a bounds kernel and a second one
__device__ __forceinline__ void warp_bounds(bool keep, float x, float y, float z, float *__restrict__ bounds) {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  if (keep && finite3(x, y, z)) { mn[0] = mx[0] = x; mn[1] = mx[1] = y; mn[2] = mx[2] = z; }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
      mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (mn[k] != FLT_MAX) atomic_min_f(&bounds[k], mn[k]);
      if (mx[k] != -FLT_MAX) atomic_max_f(&bounds[3 + k], mx[k]);
    }
  }
}

__global__ void backproject_kernel(const uint16_t *__restrict__ depth_mm, int w, int h, float fx, float fy, float cx, float cy,
                                   float4 *__restrict__ pts, unsigned char *__restrict__ flag, float *__restrict__ bounds) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = i < w * h;
  bool ok = false;
  float4 p = make_float4(0.f, 0.f, 0.f, 1.f);
  if (in) {
    const int u = i / w, v = i - u * w;
    float d = __double2float_rn((double)(float)depth_mm[i] * 0.001);   // (float)depthShort * SR300_DEPTH_UNIT: the unit is a double literal (Utils.cpp:44)
    if ((double)d > 2.0 || (double)d < 0.1) d = 0.f;
    ok = (double)d > 0.1 && (double)d < 2.0;
    if (ok) {
      p.x = __fdiv_rn(__fmul_rn(__fsub_rn((float)v, cx), d), fx);
      p.y = __fdiv_rn(__fmul_rn(__fsub_rn((float)u, cy), d), fy);
      p.z = d;
    }
    pts[i] = p;
    flag[i] = ok ? 1 : 0;
  }
  if (bounds) warp_bounds(ok, p.x, p.y, p.z, bounds);
}

// bounds[0..2] = min, [3..5] = max over finite points
__global__ void bounds_kernel(const float4 *__restrict__ pts, int n, float *__restrict__ bounds) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  if (i < n) {
    const float4 p = pts[i];
    if (finite3(p.x, p.y, p.z)) { mn[0] = mx[0] = p.x; mn[1] = mx[1] = p.y; mn[2] = mx[2] = p.z; }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
      mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (mn[k] != FLT_MAX) atomic_min_f(&bounds[k], mn[k]);
      if (mx[k] != -FLT_MAX) atomic_max_f(&bounds[3 + k], mx[k]);
    }
  }
}

struct LeafGeom { float inv; long long minb[3], divb[3]; };

// PCL's leaf index: ijk = floor(p * inv_leaf) - min_b; idx = i + j * div_x + k * div_x * div_y
__global__ void leaf_key_kernel(const float4 *__restrict__ pts, int n, LeafGeom g, unsigned long long *__restrict__ key, int *__restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  const long long a = (long long)floorf(__fmul_rn(p.x, g.inv)) - g.minb[0], b = (long long)floorf(__fmul_rn(p.y, g.inv)) - g.minb[1],
                  c = (long long)floorf(__fmul_rn(p.z, g.inv)) - g.minb[2];
  key[i] = (unsigned long long)(a + b * g.divb[0] + c * g.divb[0] * g.divb[1]);
  idx[i] = i;
}

__global__ void head_flag_kernel(const unsigned long long *__restrict__ key, int n, unsigned char *__restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flag[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
}

// one thread per leaf: sequential float sums over the leaf's points in their original order (what PCL does)
__global__ void leaf_centroid_kernel(const float4 *__restrict__ pts, const float4 *__restrict__ nrm, const int *__restrict__ idx,
                                     const int *__restrict__ starts, int m, int n, float4 *__restrict__ out_pts, float4 *__restrict__ out_nrm) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= m) return;
  const int a = starts[s], b = s + 1 < m ? starts[s + 1] : n;
  float sx = 0.f, sy = 0.f, sz = 0.f, sc = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
  for (int j = a; j < b; ++j) {
    const int i = idx[j];
    const float4 p = pts[i];
    sx = __fadd_rn(sx, p.x); sy = __fadd_rn(sy, p.y); sz = __fadd_rn(sz, p.z); sc = __fadd_rn(sc, p.w);
    if (nrm) { const float4 q = nrm[i]; nx = __fadd_rn(nx, q.x); ny = __fadd_rn(ny, q.y); nz = __fadd_rn(nz, q.z); }
  }
  const float cnt = (float)(b - a);
  out_pts[s] = make_float4(__fdiv_rn(sx, cnt), __fdiv_rn(sy, cnt), __fdiv_rn(sz, cnt), __fdiv_rn(sc, cnt));
  if (nrm) {
    const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz)));
    if (len > 0.f) { nx = __fdiv_rn(nx, len); ny = __fdiv_rn(ny, len); nz = __fdiv_rn(nz, len); }
    out_nrm[s] = make_float4(nx, ny, nz, 0.f);
  }
}

struct Xf { float m[12]; };  // rows 0..2 of a 4x4, row-major

__device__ __forceinline__ float xf_row(const Xf &T, int r, float x, float y, float z) {   // ((T0 x + T1 y) + T2 z) + T3
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T.m[4 * r], x), __fmul_rn(T.m[4 * r + 1], y)), __fmul_rn(T.m[4 * r + 2], z)), T.m[4 * r + 3]);
}

// camera -> hand base, PassThrough z / x / y (lo <= v <= hi, finite), hand base -> camera
__global__ void crop_kernel(const float4 *__restrict__ pts, int n, Xf T, Xf Ti, float3 lo, float3 hi, float4 *__restrict__ out,
                            unsigned char *__restrict__ flag, float *__restrict__ bounds) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = false;
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n) {
    const float4 p = pts[i];
    const float x = xf_row(T, 0, p.x, p.y, p.z), y = xf_row(T, 1, p.x, p.y, p.z), z = xf_row(T, 2, p.x, p.y, p.z);
    ok = finite3(x, y, z) && !(z < lo.z || z > hi.z) && !(x < lo.x || x > hi.x) && !(y < lo.y || y > hi.y);
    q = make_float4(xf_row(Ti, 0, x, y, z), xf_row(Ti, 1, x, y, z), xf_row(Ti, 2, x, y, z), p.w);
    out[i] = q;
    flag[i] = ok ? 1 : 0;
  }
  if (bounds) warp_bounds(ok, q.x, q.y, q.z, bounds);
}

struct CellGeom { float inv; int mn[3], dim[3]; };

__device__ __forceinline__ void cell_of(const CellGeom &g, float4 p, int &a, int &b, int &c) {
  a = (int)floorf(__fmul_rn(p.x, g.inv)) - g.mn[0];
  b = (int)floorf(__fmul_rn(p.y, g.inv)) - g.mn[1];
  c = (int)floorf(__fmul_rn(p.z, g.inv)) - g.mn[2];
}

__global__ void cell_key_kernel(const float4 *__restrict__ pts, int n, CellGeom g, unsigned int *__restrict__ key, int *__restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int a, b, c;
  cell_of(g, pts[i], a, b, c);
  key[i] = (unsigned int)((c * g.dim[1] + b) * g.dim[0] + a);
  idx[i] = i;
}

__global__ void cell_range_kernel(const unsigned int *__restrict__ key, const int *__restrict__ idx, const float4 *__restrict__ pts, int n,
                                  int *__restrict__ cell_start, int *__restrict__ cell_end, float4 *__restrict__ sorted) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned int k = key[i];
  if (i == 0 || key[i - 1] != k) cell_start[k] = i;
  if (i == n - 1 || key[i + 1] != k) cell_end[k] = i + 1;
  sorted[i] = pts[idx[i]];
}

// PCA normal over the neighbours within the radius, flipped towards the viewpoint (host: estimateNormals)
__global__ void normals_kernel(const float4 *__restrict__ pts, int n, const float4 *__restrict__ sorted, const int *__restrict__ cell_start,
                               const int *__restrict__ cell_end, CellGeom g, float r2, float3 vp, float4 *__restrict__ nrm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  int ca, cb, cc;
  cell_of(g, p, ca, cb, cc);
  double s[3] = {0, 0, 0}, ss[6] = {0, 0, 0, 0, 0, 0};
  int cnt = 0;
  for (int dc = -1; dc <= 1; ++dc)
    for (int db = -1; db <= 1; ++db)
      for (int da = -1; da <= 1; ++da) {
        const int a = ca + da, b = cb + db, c = cc + dc;
        if ((unsigned)a >= (unsigned)g.dim[0] || (unsigned)b >= (unsigned)g.dim[1] || (unsigned)c >= (unsigned)g.dim[2]) continue;
        const int cell = (c * g.dim[1] + b) * g.dim[0] + a;
        for (int j = cell_start[cell]; j < cell_end[cell]; ++j) {
          const float4 q = sorted[j];
          const float dx = __fsub_rn(q.x, p.x), dy = __fsub_rn(q.y, p.y), dz = __fsub_rn(q.z, p.z);
          if (__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)) > r2) continue;
          s[0] += dx; s[1] += dy; s[2] += dz;
          ss[0] += __fmul_rn(dx, dx); ss[1] += __fmul_rn(dx, dy); ss[2] += __fmul_rn(dx, dz);
          ss[3] += __fmul_rn(dy, dy); ss[4] += __fmul_rn(dy, dz); ss[5] += __fmul_rn(dz, dz);
          ++cnt;
        }
      }
  const float nan = __int_as_float(0x7fc00000);
  if (cnt < 3) { nrm[i] = make_float4(nan, nan, nan, 0.f); return; }
  const double m0 = s[0] / cnt, m1 = s[1] / cnt, m2 = s[2] / cnt;
  double C[3][3] = {{ss[0] / cnt - m0 * m0, ss[1] / cnt - m0 * m1, ss[2] / cnt - m0 * m2},
                    {0, ss[3] / cnt - m1 * m1, ss[4] / cnt - m1 * m2}, {0, 0, ss[5] / cnt - m2 * m2}};
  C[1][0] = C[0][1]; C[2][0] = C[0][2]; C[2][1] = C[1][2];
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 12; ++sweep)
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = a + 1; b < 3; ++b) {
        if (fabs(C[a][b]) < 1e-30) continue;
        const double th = 0.5 * atan2(2 * C[a][b], C[b][b] - C[a][a]);
        double sn, cs;
        sincos(th, &sn, &cs);
#pragma unroll
        for (int k = 0; k < 3; ++k) { const double x = C[k][a], y = C[k][b]; C[k][a] = cs * x - sn * y; C[k][b] = sn * x + cs * y; }
#pragma unroll
        for (int k = 0; k < 3; ++k) { const double x = C[a][k], y = C[b][k]; C[a][k] = cs * x - sn * y; C[b][k] = sn * x + cs * y; }
#pragma unroll
        for (int k = 0; k < 3; ++k) { const double x = V[k][a], y = V[k][b]; V[k][a] = cs * x - sn * y; V[k][b] = sn * x + cs * y; }
      }
  int mi = 0;
  if (C[1][1] < C[0][0]) mi = 1;
  if (C[2][2] < (mi == 0 ? C[0][0] : C[1][1])) mi = 2;
  float nx = (float)(mi == 0 ? V[0][0] : mi == 1 ? V[0][1] : V[0][2]);
  float ny = (float)(mi == 0 ? V[1][0] : mi == 1 ? V[1][1] : V[1][2]);
  float nz = (float)(mi == 0 ? V[2][0] : mi == 1 ? V[2][1] : V[2][2]);
  if (__fadd_rn(__fadd_rn(__fmul_rn(__fsub_rn(vp.x, p.x), nx), __fmul_rn(__fsub_rn(vp.y, p.y), ny)), __fmul_rn(__fsub_rn(vp.z, p.z), nz)) < 0.f) {
    nx = -nx; ny = -ny; nz = -nz;
  }
  nrm[i] = make_float4(nx, ny, nz, 0.f);
}

// removeAllNaN + pcl::flipNormalTowardsViewpoint(origin) + confidence 1 (main_realdata_auto.cpp:160-177)
__global__ void finish_kernel(const float4 *__restrict__ pts, const float4 *__restrict__ nrm, int n, float4 *__restrict__ out_pts,
                              float4 *__restrict__ out_nrm, unsigned char *__restrict__ flag, float *__restrict__ bounds) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = false;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n) {
    p = pts[i];
    float4 q = nrm[i];
    ok = finite3(p.x, p.y, p.z) && finite3(q.x, q.y, q.z);
    if (__fsub_rn(__fsub_rn(__fmul_rn(-p.x, q.x), __fmul_rn(p.y, q.y)), __fmul_rn(p.z, q.z)) < 0.f) { q.x = -q.x; q.y = -q.y; q.z = -q.z; }
    out_pts[i] = make_float4(p.x, p.y, p.z, 1.f);
    out_nrm[i] = q;
    flag[i] = ok ? 1 : 0;
  }
  if (bounds) warp_bounds(ok, p.x, p.y, p.z, bounds);
}

// compacted (xyz, conf) / normal arrays -> the cloud's padded streams
__global__ void to_cloud_kernel(const float4 *__restrict__ pts, const float4 *__restrict__ nrm, int n, int n_padded, float4 *__restrict__ pw,
                                float4 *__restrict__ nv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_padded) return;
  if (i >= n) { pw[i] = make_float4(HOP_SENTINEL, HOP_SENTINEL, HOP_SENTINEL, 0.f); nv[i] = make_float4(0.f, 0.f, 0.f, 0.f); return; }
  pw[i] = pts[i];
  const float4 q = nrm[i];
  const float s = q.x * q.x + q.y * q.y + q.z * q.z;
  nv[i] = make_float4(q.x, q.y, q.z, s > 0.f ? 1.f / sqrtf(s) : 0.f);
}

#define FR_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return HOP_ECUDA; } } while (0)

inline int blocks(int n) { return (n + 255) / 256; }

// order-preserving compaction of up to two parallel float4 arrays; returns the kept count (synchronises)
// d_bounds (may be null): six floats the producing kernel accumulated (warp_bounds) -> bounds_out[0..2] = min, [3..5] = max, with the count
int compact2(hop_ctx *ctx, const float4 *a, const float4 *b, const unsigned char *flag, int n, float4 *oa, float4 *ob, int *kept,
             const float *d_bounds = nullptr, float *bounds_out = nullptr) {
  cudaStream_t st = ctx->stream;
  *kept = 0;
  if (n <= 0) return HOP_OK;
  FBuf cnt(st), tmp(st);
  FR_CUDA(cnt.alloc(sizeof(int)));
  size_t bytes = 0;
  cub::DeviceSelect::Flagged(nullptr, bytes, a, flag, oa, cnt.as<int>(), n, st);
  FR_CUDA(tmp.alloc(bytes));
  cub::DeviceSelect::Flagged(tmp.p, bytes, a, flag, oa, cnt.as<int>(), n, st);
  if (b) cub::DeviceSelect::Flagged(tmp.p, bytes, b, flag, ob, cnt.as<int>(), n, st);
  ctx->launches += b ? 2 : 1;
  FR_CUDA(cudaMemcpyAsync(kept, cnt.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  if (d_bounds && bounds_out) FR_CUDA(cudaMemcpyAsync(bounds_out, d_bounds, 6 * sizeof(float), cudaMemcpyDeviceToHost, st));
  FR_CUDA(cudaStreamSynchronize(st));
  return HOP_OK;
}

int cloud_bounds(hop_ctx *ctx, const float4 *pts, int n, float *mn, float *mx) {
  cudaStream_t st = ctx->stream;
  FBuf b(st);
  FR_CUDA(b.alloc(6 * sizeof(float)));
  const float init[6] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
  FR_CUDA(cudaMemcpyAsync(b.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
  bounds_kernel<<<blocks(n), 256, 0, st>>>(pts, n, b.as<float>());
  ctx->launches += 1;
  float out[6];
  FR_CUDA(cudaMemcpyAsync(out, b.p, sizeof(out), cudaMemcpyDeviceToHost, st));
  FR_CUDA(cudaStreamSynchronize(st));
  for (int k = 0; k < 3; ++k) { mn[k] = out[k]; mx[k] = out[3 + k]; }
  return HOP_OK;
}

// pcl::VoxelGrid on device arrays; out arrays must hold n entries
int voxel_grid(hop_ctx *ctx, const float4 *pts, const float4 *nrm, int n, float leaf, float4 *out_pts, float4 *out_nrm, int *m_out,
               const float *known_bounds = nullptr /* min xyz, max xyz of pts when the caller already has them */) {
  cudaStream_t st = ctx->stream;
  *m_out = 0;
  if (n <= 0) return HOP_OK;
  float mn[3], mx[3];
  if (known_bounds) { for (int k = 0; k < 3; ++k) { mn[k] = known_bounds[k]; mx[k] = known_bounds[3 + k]; } }
  else {
    int rc = cloud_bounds(ctx, pts, n, mn, mx);
    if (rc != HOP_OK) return rc;
  }
  LeafGeom g;
  g.inv = 1.0f / leaf;
  for (int k = 0; k < 3; ++k) { g.minb[k] = (long long)std::floor(mn[k] * g.inv); g.divb[k] = (long long)std::floor(mx[k] * g.inv) - g.minb[k] + 1; }
  const double total = (double)g.divb[0] * (double)g.divb[1] * (double)g.divb[2];
  if (!(total < 9.0e18)) { ctx->err = "voxel grid: leaf too small for the cloud's extent"; return HOP_EINVAL; }
  int end_bit = 1;
  while (end_bit < 64 && (double)(1ull << end_bit) <= total) ++end_bit;
  FBuf k0(st), k1(st), i0(st), i1(st), fl(st), starts(st), cnt(st), tmp(st);
  FR_CUDA(k0.alloc(sizeof(unsigned long long) * (size_t)n)); FR_CUDA(k1.alloc(sizeof(unsigned long long) * (size_t)n));
  FR_CUDA(i0.alloc(sizeof(int) * (size_t)n)); FR_CUDA(i1.alloc(sizeof(int) * (size_t)n));
  FR_CUDA(fl.alloc((size_t)n)); FR_CUDA(starts.alloc(sizeof(int) * (size_t)n)); FR_CUDA(cnt.alloc(sizeof(int)));
  leaf_key_kernel<<<blocks(n), 256, 0, st>>>(pts, n, g, k0.as<unsigned long long>(), i0.as<int>());
  size_t sort_bytes = 0, sel_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, k0.as<unsigned long long>(), k1.as<unsigned long long>(), i0.as<int>(), i1.as<int>(), n, 0, end_bit, st);
  cub::CountingInputIterator<int> counting(0);
  cub::DeviceSelect::Flagged(nullptr, sel_bytes, counting, fl.as<unsigned char>(), starts.as<int>(), cnt.as<int>(), n, st);
  FR_CUDA(tmp.alloc(std::max(sort_bytes, sel_bytes)));
  cub::DeviceRadixSort::SortPairs(tmp.p, sort_bytes, k0.as<unsigned long long>(), k1.as<unsigned long long>(), i0.as<int>(), i1.as<int>(), n, 0, end_bit, st);
  head_flag_kernel<<<blocks(n), 256, 0, st>>>(k1.as<unsigned long long>(), n, fl.as<unsigned char>());
  cub::DeviceSelect::Flagged(tmp.p, sel_bytes, counting, fl.as<unsigned char>(), starts.as<int>(), cnt.as<int>(), n, st);
  ctx->launches += 4;
  int m = 0;
  FR_CUDA(cudaMemcpyAsync(&m, cnt.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  FR_CUDA(cudaStreamSynchronize(st));
  leaf_centroid_kernel<<<blocks(m), 256, 0, st>>>(pts, nrm, i1.as<int>(), starts.as<int>(), m, n, out_pts, out_nrm);
  ctx->launches += 1;
  FR_CUDA(cudaGetLastError());
  *m_out = m;
  return HOP_OK;
}

Xf rows_of(const float *colmajor) {
  Xf T;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) T.m[4 * r + c] = colmajor[4 * c + r];
  return T;
}

}  // namespace

extern "C" void hop_default_frame_params(hop_frame_params *p) {
  if (!p) return;
  p->fx = p->fy = 616.6f; p->cx = 307.6f; p->cy = 239.7f;
  p->leaf_dense = 0.001f; p->normal_radius = 0.003f; p->leaf_object = 0.003f;
  for (int k = 0; k < 16; ++k) p->cam_in_handbase[k] = p->handbase_in_cam[k] = (k % 5 == 0) ? 1.f : 0.f;
  p->box_min[0] = -0.25f; p->box_max[0] = -0.07f;   // main_realdata_auto.cpp: PassThrough x / y / z in the hand-base frame
  p->box_min[1] = -0.2f; p->box_max[1] = 0.2f;
  p->box_min[2] = -0.12f; p->box_max[2] = 0.05f;
  p->viewpoint[0] = p->viewpoint[1] = p->viewpoint[2] = 0.f;
}

extern "C" int hop_frame_to_scene(hop_ctx *ctx, const uint16_t *depth_mm, int width, int height, const hop_frame_params *fp, hop_cloud **scene,
                                  int32_t *stage_counts) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  if (!depth_mm || width <= 0 || height <= 0 || !fp || !scene) { ctx->err = "hop_frame_to_scene: bad arguments"; return HOP_EINVAL; }
  if (!(fp->leaf_dense > 0.f) || !(fp->leaf_object > 0.f) || !(fp->normal_radius > 0.f)) { ctx->err = "hop_frame_to_scene: leaf sizes and radius must be > 0"; return HOP_EINVAL; }
  ProfScope ps(ctx, HOP_PROF_FRAME);
  cudaStream_t st = ctx->stream;
  const int npx = width * height;
  int32_t counts[5] = {0, 0, 0, 0, 0};   // valid pixels, dense leaves, cropped, object leaves, final
  FBuf d_depth(st), A(st), B(st), NA(st), NB(st), fl(st);
  FR_CUDA(d_depth.alloc(sizeof(uint16_t) * (size_t)npx));
  FR_CUDA(A.alloc(sizeof(float4) * (size_t)npx)); FR_CUDA(B.alloc(sizeof(float4) * (size_t)npx)); FR_CUDA(fl.alloc((size_t)npx));
  FR_CUDA(cudaMemcpyAsync(d_depth.p, depth_mm, sizeof(uint16_t) * (size_t)npx, cudaMemcpyHostToDevice, st));
  // the bounding boxes of the three kept sets (valid pixels, cropped points, final points) are accumulated by the kernels that decide
  // what is kept and come back with the counts: three round trips fewer per frame than a bounds pass before every consumer
  FBuf bnd(st);
  FR_CUDA(bnd.alloc(18 * sizeof(float)));
  static const float kBoundsInit[18] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX,
                                        FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
  FR_CUDA(cudaMemcpyAsync(bnd.p, kBoundsInit, sizeof(kBoundsInit), cudaMemcpyHostToDevice, st));
  float *d_bnd = bnd.as<float>();
  float bnd_valid[6], bnd_crop[6], bnd_final[6];
  // 1. back-projection, raster order
  backproject_kernel<<<blocks(npx), 256, 0, st>>>(d_depth.as<uint16_t>(), width, height, fp->fx, fp->fy, fp->cx, fp->cy, A.as<float4>(), fl.as<unsigned char>(), d_bnd);
  ctx->launches += 1;
  int n = 0, rc;
  if ((rc = compact2(ctx, A.as<float4>(), nullptr, fl.as<unsigned char>(), npx, B.as<float4>(), nullptr, &n, d_bnd, bnd_valid)) != HOP_OK) return rc;
  counts[0] = n;
  // 2. dense voxel grid
  int m = 0;
  if ((rc = voxel_grid(ctx, B.as<float4>(), nullptr, n, fp->leaf_dense, A.as<float4>(), nullptr, &m, bnd_valid)) != HOP_OK) return rc;
  counts[1] = m;
  // 3. crop in the hand-base frame
  int nc = 0;
  if (m > 0) {
    crop_kernel<<<blocks(m), 256, 0, st>>>(A.as<float4>(), m, rows_of(fp->cam_in_handbase), rows_of(fp->handbase_in_cam),
                                           make_float3(fp->box_min[0], fp->box_min[1], fp->box_min[2]), make_float3(fp->box_max[0], fp->box_max[1], fp->box_max[2]),
                                           B.as<float4>(), fl.as<unsigned char>(), d_bnd + 6);
    ctx->launches += 1;
    if ((rc = compact2(ctx, B.as<float4>(), nullptr, fl.as<unsigned char>(), m, A.as<float4>(), nullptr, &nc, d_bnd + 6, bnd_crop)) != HOP_OK) return rc;
  }
  counts[2] = nc;
  // 4. normals over normal_radius
  int mo = 0, nf = 0;
  FR_CUDA(NA.alloc(sizeof(float4) * (size_t)std::max(nc, 1))); FR_CUDA(NB.alloc(sizeof(float4) * (size_t)std::max(nc, 1)));
  if (nc > 0) {
    const float *mn = bnd_crop, *mx = bnd_crop + 3;
    CellGeom g;
    g.inv = 1.f / fp->normal_radius;
    double ncell = 1;
    for (int k = 0; k < 3; ++k) {
      g.mn[k] = (int)std::floor(mn[k] * g.inv);
      g.dim[k] = (int)std::floor(mx[k] * g.inv) - g.mn[k] + 1;
      ncell *= g.dim[k];
    }
    if (!(ncell < 2.5e8)) { ctx->err = "hop_frame_to_scene: normal_radius too small for the cropped extent"; return HOP_EINVAL; }
    const int nce = (int)ncell;
    FBuf k0(st), k1(st), i0(st), i1(st), cs(st), ce(st), sorted(st), tmp(st);
    FR_CUDA(k0.alloc(sizeof(unsigned int) * (size_t)nc)); FR_CUDA(k1.alloc(sizeof(unsigned int) * (size_t)nc));
    FR_CUDA(i0.alloc(sizeof(int) * (size_t)nc)); FR_CUDA(i1.alloc(sizeof(int) * (size_t)nc));
    FR_CUDA(cs.alloc(sizeof(int) * (size_t)nce)); FR_CUDA(ce.alloc(sizeof(int) * (size_t)nce)); FR_CUDA(sorted.alloc(sizeof(float4) * (size_t)nc));
    FR_CUDA(cudaMemsetAsync(cs.p, 0, sizeof(int) * (size_t)nce, st)); FR_CUDA(cudaMemsetAsync(ce.p, 0, sizeof(int) * (size_t)nce, st));
    cell_key_kernel<<<blocks(nc), 256, 0, st>>>(A.as<float4>(), nc, g, k0.as<unsigned int>(), i0.as<int>());
    size_t sort_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, k0.as<unsigned int>(), k1.as<unsigned int>(), i0.as<int>(), i1.as<int>(), nc, 0, 32, st);
    FR_CUDA(tmp.alloc(sort_bytes));
    cub::DeviceRadixSort::SortPairs(tmp.p, sort_bytes, k0.as<unsigned int>(), k1.as<unsigned int>(), i0.as<int>(), i1.as<int>(), nc, 0, 32, st);
    cell_range_kernel<<<blocks(nc), 256, 0, st>>>(k1.as<unsigned int>(), i1.as<int>(), A.as<float4>(), nc, cs.as<int>(), ce.as<int>(), sorted.as<float4>());
    normals_kernel<<<(nc + 127) / 128, 128, 0, st>>>(A.as<float4>(), nc, sorted.as<float4>(), cs.as<int>(), ce.as<int>(), g, fp->normal_radius * fp->normal_radius,
                                                     make_float3(fp->viewpoint[0], fp->viewpoint[1], fp->viewpoint[2]), NA.as<float4>());
    ctx->launches += 4;
    // 5. object voxel grid with normals, NaN removal, flip, confidence
    if ((rc = voxel_grid(ctx, A.as<float4>(), NA.as<float4>(), nc, fp->leaf_object, B.as<float4>(), NB.as<float4>(), &mo, bnd_crop)) != HOP_OK) return rc;
    if (mo > 0) {
      finish_kernel<<<blocks(mo), 256, 0, st>>>(B.as<float4>(), NB.as<float4>(), mo, A.as<float4>(), NA.as<float4>(), fl.as<unsigned char>(), d_bnd + 12);
      ctx->launches += 1;
      if ((rc = compact2(ctx, A.as<float4>(), NA.as<float4>(), fl.as<unsigned char>(), mo, B.as<float4>(), NB.as<float4>(), &nf, d_bnd + 12, bnd_final)) != HOP_OK) return rc;
    }
  }
  counts[3] = mo; counts[4] = nf;
  if (stage_counts) for (int k = 0; k < 5; ++k) stage_counts[k] = counts[k];
  // the cloud
  hop_cloud *c = *scene ? *scene : new hop_cloud();
  rc = hop_cloud_reserve(ctx, c, nf);
  if (rc != HOP_OK) { if (!*scene) hop_cloud_free(ctx, c); return rc; }
  to_cloud_kernel<<<blocks(c->n_padded), 256, 0, st>>>(B.as<float4>(), NB.as<float4>(), nf, c->n_padded, c->d_pw, c->d_nv);
  ctx->launches += 1;
  for (int k = 0; k < 3; ++k) { c->bbox_min[k] = nf > 0 ? bnd_final[k] : 0.f; c->bbox_max[k] = nf > 0 ? bnd_final[3 + k] : 0.f; }
  FR_CUDA(cudaGetLastError());
  *scene = c;
  return HOP_OK;
}

// ---- cloud filters of Hand::setCurScene (Hand.cpp:279-334): VoxelGrid, transform, PassThrough, RadiusOutlierRemoval x2,
//      StatisticalOutlierRemoval -- each returns a new device cloud, points in input order -------------------------------------------
namespace {
Xf xf_from(const float *colmajor) {
  Xf T;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) T.m[4 * r + c] = colmajor[4 * c + r];
  return T;
}

// compacted (xyz, w) / normal arrays -> *out (created when null), with its bounding box
int make_cloud(hop_ctx *ctx, const float4 *pts, const float4 *nrm, int n, hop_cloud **out) {
  hop_cloud *c = *out ? *out : new hop_cloud();
  int rc = hop_cloud_reserve(ctx, c, n);
  if (rc != HOP_OK) { if (!*out) hop_cloud_free(ctx, c); return rc; }
  to_cloud_kernel<<<blocks(c->n_padded), 256, 0, ctx->stream>>>(pts, nrm, n, c->n_padded, c->d_pw, c->d_nv);
  ctx->launches += 1;
  float mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
  if (n > 0 && (rc = cloud_bounds(ctx, pts, n, mn, mx)) != HOP_OK) { if (!*out) hop_cloud_free(ctx, c); return rc; }
  for (int k = 0; k < 3; ++k) { c->bbox_min[k] = mn[k]; c->bbox_max[k] = mx[k]; }
  FR_CUDA(cudaGetLastError());
  *out = c;
  return HOP_OK;
}

__global__ void transform_kernel(const float4 *__restrict__ pw, const float4 *__restrict__ nv, int n, Xf T, float4 *__restrict__ opw, float4 *__restrict__ onv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pw[i], q = nv[i];
  opw[i] = make_float4(xf_row(T, 0, p.x, p.y, p.z), xf_row(T, 1, p.x, p.y, p.z), xf_row(T, 2, p.x, p.y, p.z), p.w);
  onv[i] = make_float4(__fadd_rn(__fadd_rn(__fmul_rn(T.m[0], q.x), __fmul_rn(T.m[1], q.y)), __fmul_rn(T.m[2], q.z)),
                       __fadd_rn(__fadd_rn(__fmul_rn(T.m[4], q.x), __fmul_rn(T.m[5], q.y)), __fmul_rn(T.m[6], q.z)),
                       __fadd_rn(__fadd_rn(__fmul_rn(T.m[8], q.x), __fmul_rn(T.m[9], q.y)), __fmul_rn(T.m[10], q.z)), 0.f);
}

__global__ void pass_flag_kernel(const float4 *__restrict__ pw, int n, int axis, float lo, float hi, unsigned char *__restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pw[i];
  const float v = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
  flag[i] = (finite3(p.x, p.y, p.z) && !(v < lo || v > hi)) ? 1 : 0;   // pcl::PassThrough keeps lo <= v <= hi
}

// pcl::RadiusOutlierRemoval (dense input): a point stays when its (min_pts + 1)-th nearest neighbour, itself included, lies within
// the radius, i.e. when at least min_pts + 1 points have d^2 <= r^2.  One thread per point, all lanes scan the cloud together.
__global__ void __launch_bounds__(128) ror_kernel(const float4 *__restrict__ pw, int n, float r2, int need, unsigned char *__restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  const float4 p = live ? pw[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  int cnt = 0;
  for (int j = 0; j < n; ++j) {
    const float4 c = __ldg(pw + j);
    const float dx = c.x - p.x, dy = c.y - p.y, dz = c.z - p.z;
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    cnt += d2 <= r2;
  }
  if (live) flag[i] = (finite3(p.x, p.y, p.z) && cnt >= need) ? 1 : 0;
}

// pcl::StatisticalOutlierRemoval, first pass: mean distance to the mean_k nearest neighbours (the nearest hit, normally the point
// itself, is skipped), distances summed in ascending order in double.  KMAX bounds mean_k + 1.
constexpr int SOR_KMAX = 65;
__global__ void __launch_bounds__(128) sor_mean_kernel(const float4 *__restrict__ pw, int n, int k1 /* mean_k + 1 */, float *__restrict__ mean_dist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pw[i];
  float best[SOR_KMAX];
  for (int k = 0; k < k1; ++k) best[k] = FLT_MAX;
  for (int j = 0; j < n; ++j) {
    const float4 c = __ldg(pw + j);
    const float dx = c.x - p.x, dy = c.y - p.y, dz = c.z - p.z;
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    if (!(d2 < best[k1 - 1])) continue;
    int k = k1 - 1;
    while (k > 0 && best[k - 1] > d2) { best[k] = best[k - 1]; --k; }
    best[k] = d2;
  }
  double sum = 0.0;
  const int have = min(k1, n);
  for (int k = 1; k < have; ++k) sum += (double)__fsqrt_rn(best[k]);
  mean_dist[i] = (float)(sum / (double)(k1 - 1));
}

__global__ void sor_flag_kernel(const float4 *__restrict__ pw, const float *__restrict__ mean_dist, int n, double thr, unsigned char *__restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pw[i];
  flag[i] = (finite3(p.x, p.y, p.z) && (double)mean_dist[i] <= thr) ? 1 : 0;
}

// Hand::handbaseICP's hand-base region (Hand.cpp:704-728): drops the finger connection discs (15 mm around (y1, z1) and (y2, z2) in
// the y-z plane) and the strip between the two finger axes within 10 mm of z1 (the reference tests z1 twice)
__global__ void handbase_region_kernel(const float4 *__restrict__ pw, int n, float y1, float z1, float y2, float z2, unsigned char *__restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pw[i];
  bool keep = true;
  const float a1 = __fsub_rn(p.z, z1), b1 = __fsub_rn(p.y, y1), a2 = __fsub_rn(p.z, z2), b2 = __fsub_rn(p.y, y2);
  if ((double)__fadd_rn(__fmul_rn(a1, a1), __fmul_rn(b1, b1)) <= 0.015 * 0.015) keep = false;
  else if ((double)__fadd_rn(__fmul_rn(a2, a2), __fmul_rn(b2, b2)) <= 0.015 * 0.015) keep = false;
  else if (((p.y >= y1 && p.y <= y2) || (p.y >= y2 && p.y <= y1)) && (double)fabsf(a1) <= 0.01) keep = false;
  flag[i] = keep ? 1 : 0;
}

// keep the flagged points of `in` (order preserved) as a new cloud
int keep_flagged(hop_ctx *ctx, const hop_cloud *in, const unsigned char *flag, hop_cloud **out) {
  cudaStream_t st = ctx->stream;
  FBuf B(st), NB(st);
  int kept = 0, rc;
  if (in->n > 0) {
    FR_CUDA(B.alloc(sizeof(float4) * (size_t)in->n)); FR_CUDA(NB.alloc(sizeof(float4) * (size_t)in->n));
    if ((rc = compact2(ctx, in->d_pw, in->d_nv, flag, in->n, B.as<float4>(), NB.as<float4>(), &kept)) != HOP_OK) return rc;
  }
  return make_cloud(ctx, B.as<float4>(), NB.as<float4>(), kept, out);
}
}  // namespace

extern "C" int hop_cloud_voxel_grid(hop_ctx *ctx, const hop_cloud *in, float leaf, hop_cloud **out) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  if (!in || !out || *out == in || !(leaf > 0.f)) { ctx->err = "hop_cloud_voxel_grid: bad arguments"; return HOP_EINVAL; }
  cudaStream_t st = ctx->stream;
  FBuf B(st), NB(st);
  int m = 0, rc;
  if (in->n > 0) {
    FR_CUDA(B.alloc(sizeof(float4) * (size_t)in->n)); FR_CUDA(NB.alloc(sizeof(float4) * (size_t)in->n));
    ProfScope ps(ctx, HOP_PROF_FRAME);
    if ((rc = voxel_grid(ctx, in->d_pw, in->d_nv, in->n, leaf, B.as<float4>(), NB.as<float4>(), &m)) != HOP_OK) return rc;
  }
  return make_cloud(ctx, B.as<float4>(), NB.as<float4>(), m, out);
}

extern "C" int hop_cloud_transform(hop_ctx *ctx, const hop_cloud *in, const float *T, hop_cloud **out) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  if (!in || !out || *out == in || !T) { ctx->err = "hop_cloud_transform: bad arguments"; return HOP_EINVAL; }
  cudaStream_t st = ctx->stream;
  FBuf B(st), NB(st);
  if (in->n > 0) {
    FR_CUDA(B.alloc(sizeof(float4) * (size_t)in->n)); FR_CUDA(NB.alloc(sizeof(float4) * (size_t)in->n));
    transform_kernel<<<blocks(in->n), 256, 0, st>>>(in->d_pw, in->d_nv, in->n, xf_from(T), B.as<float4>(), NB.as<float4>());
    ctx->launches += 1;
  }
  return make_cloud(ctx, B.as<float4>(), NB.as<float4>(), in->n, out);
}

extern "C" int hop_cloud_pass_through(hop_ctx *ctx, const hop_cloud *in, int axis, float lo, float hi, hop_cloud **out) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  if (!in || !out || *out == in || axis < 0 || axis > 2) { ctx->err = "hop_cloud_pass_through: bad arguments"; return HOP_EINVAL; }
  FBuf fl(ctx->stream);
  if (in->n > 0) {
    FR_CUDA(fl.alloc((size_t)in->n));
    pass_flag_kernel<<<blocks(in->n), 256, 0, ctx->stream>>>(in->d_pw, in->n, axis, lo, hi, fl.as<unsigned char>());
    ctx->launches += 1;
  }
  return keep_flagged(ctx, in, fl.as<unsigned char>(), out);
}

extern "C" int hop_cloud_handbase_region(hop_ctx *ctx, const hop_cloud *in, float y1, float z1, float y2, float z2, hop_cloud **out) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  if (!in || !out || *out == in) { ctx->err = "hop_cloud_handbase_region: bad arguments"; return HOP_EINVAL; }
  FBuf fl(ctx->stream);
  if (in->n > 0) {
    FR_CUDA(fl.alloc((size_t)in->n));
    handbase_region_kernel<<<blocks(in->n), 256, 0, ctx->stream>>>(in->d_pw, in->n, y1, z1, y2, z2, fl.as<unsigned char>());
    ctx->launches += 1;
  }
  return keep_flagged(ctx, in, fl.as<unsigned char>(), out);
}

extern "C" int hop_cloud_radius_outlier_removal(hop_ctx *ctx, const hop_cloud *in, float radius, int min_neighbors, hop_cloud **out) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  if (!in || !out || *out == in || !(radius > 0.f) || min_neighbors < 0) { ctx->err = "hop_cloud_radius_outlier_removal: bad arguments"; return HOP_EINVAL; }
  FBuf fl(ctx->stream);
  if (in->n > 0) {
    FR_CUDA(fl.alloc((size_t)in->n));
    ProfScope ps(ctx, HOP_PROF_FRAME);
    ror_kernel<<<(in->n + 127) / 128, 128, 0, ctx->stream>>>(in->d_pw, in->n, (float)((double)radius * (double)radius), min_neighbors + 1, fl.as<unsigned char>());
    ctx->launches += 1;
  }
  return keep_flagged(ctx, in, fl.as<unsigned char>(), out);
}

extern "C" int hop_cloud_statistical_outlier_removal(hop_ctx *ctx, const hop_cloud *in, int mean_k, float stddev_mul, hop_cloud **out) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  if (!in || !out || *out == in || mean_k < 1 || mean_k + 1 > SOR_KMAX) { ctx->err = "hop_cloud_statistical_outlier_removal: bad arguments (mean_k 1..64)"; return HOP_EINVAL; }
  cudaStream_t st = ctx->stream;
  FBuf fl(st), md(st);
  const int n = in->n;
  if (n > 0) {
    FR_CUDA(fl.alloc((size_t)n)); FR_CUDA(md.alloc(sizeof(float) * (size_t)n));
    {
      ProfScope ps(ctx, HOP_PROF_FRAME);
      sor_mean_kernel<<<(n + 127) / 128, 128, 0, st>>>(in->d_pw, n, mean_k + 1, md.as<float>());
      ctx->launches += 1;
    }
    // the global statistics are a sequential double sum over the points in order (statistical_outlier_removal.hpp): on the host
    std::vector<float> h(n);
    FR_CUDA(cudaMemcpyAsync(h.data(), md.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, st));
    FR_CUDA(cudaStreamSynchronize(st));
    double sum = 0, sq_sum = 0;
    for (int i = 0; i < n; ++i) { sum += h[i]; sq_sum += (double)h[i] * h[i]; }
    const double mean = sum / (double)n;
    const double variance = (sq_sum - sum * sum / (double)n) / ((double)n - 1);
    const double thr = mean + (double)stddev_mul * std::sqrt(variance);
    sor_flag_kernel<<<blocks(n), 256, 0, st>>>(in->d_pw, md.as<float>(), n, thr, fl.as<unsigned char>());
    ctx->launches += 1;
  }
  return keep_flagged(ctx, in, fl.as<unsigned char>(), out);
}

// ---- HandT42::removeSurroundingPointsAndAssignProbability (Hand.cpp:781-888) ------------------------------------------------------
namespace {
constexpr int MAX_LINKS = 16;
struct HandRemovalArgs {
  const float4 *pw, *nv; int n;                 // the scene, camera frame
  const float4 *link[MAX_LINKS]; int link_n[MAX_LINKS]; float link_thr[MAX_LINKS]; int n_links;
  Xf cam_in_hb, hb_in_cam, hb_in_f12, hb_in_f22;
  float min_z;
  float4 *out_pw, *out_nv; unsigned char *flag;
};

__device__ __forceinline__ float xf_rot(const Xf &T, int r, float x, float y, float z) {   // (T0 x + T1 y) + T2 z
  return __fadd_rn(__fadd_rn(__fmul_rn(T.m[4 * r], x), __fmul_rn(T.m[4 * r + 1], y)), __fmul_rn(T.m[4 * r + 2], z));
}

// one thread per scene point; every link cloud is scanned by all lanes at the same address (broadcast loads): the clouds are a
// few thousand points, a brute-force exact nearest neighbour at ANY distance (the confidence needs it) is a few microseconds
__global__ void __launch_bounds__(128) hand_removal_kernel(HandRemovalArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const float4 p = a.pw[i], q = a.nv[i];
  const float x = xf_row(a.cam_in_hb, 0, p.x, p.y, p.z), y = xf_row(a.cam_in_hb, 1, p.x, p.y, p.z), z = xf_row(a.cam_in_hb, 2, p.x, p.y, p.z);
  const float nx = xf_rot(a.cam_in_hb, 0, q.x, q.y, q.z), ny = xf_rot(a.cam_in_hb, 1, q.x, q.y, q.z), nz = xf_rot(a.cam_in_hb, 2, q.x, q.y, q.z);
  bool near = false;
  float min_dist = 1.0f;
  for (int k = 0; k < a.n_links && !near; ++k) {
    const int m = a.link_n[k];
    if (m <= 0) continue;
    const float4 *lp = a.link[k];
    float bd = 3.4e38f; int bi = 0;
    for (int j = 0; j < m; ++j) {
      const float4 c = __ldg(lp + j);
      const float dx = c.x - x, dy = c.y - y, dz = c.z - z;
      const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      if (d2 < bd) { bd = d2; bi = j; }
    }
    min_dist = fminf(min_dist, __fsqrt_rn(bd));
    if (bd <= a.link_thr[k]) { near = true; break; }
    const float4 c = __ldg(lp + bi);
    const float px = x - c.x, py = y - c.y;
    const float planar = __fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py));
    if (planar <= a.link_thr[k] && (double)fabsf(z - c.z) <= 0.005) { near = true; break; }
  }
  bool ok = !near;
  if (ok) {
    const float y1 = xf_row(a.hb_in_f12, 1, x, y, z), z1 = xf_row(a.hb_in_f12, 2, x, y, z);
    const float y2 = xf_row(a.hb_in_f22, 1, x, y, z), z2 = xf_row(a.hb_in_f22, 2, x, y, z);
    if ((y1 < 0.f && z1 >= a.min_z) || (y2 < 0.f && z2 >= a.min_z)) ok = false;
  }
  const float conf = 1.f - expf(__fmul_rn(-231.04906018664843f, min_dist));
  a.out_pw[i] = make_float4(xf_row(a.hb_in_cam, 0, x, y, z), xf_row(a.hb_in_cam, 1, x, y, z), xf_row(a.hb_in_cam, 2, x, y, z), conf);
  a.out_nv[i] = make_float4(xf_rot(a.hb_in_cam, 0, nx, ny, nz), xf_rot(a.hb_in_cam, 1, nx, ny, nz), xf_rot(a.hb_in_cam, 2, nx, ny, nz), 0.f);
  a.flag[i] = ok ? 1 : 0;
}

Xf xf_of(const float *colmajor) {
  Xf T;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) T.m[4 * r + c] = colmajor[4 * c + r];
  return T;
}
}  // namespace

extern "C" int hop_remove_hand_points(hop_ctx *ctx, const hop_cloud *scene, const hop_cloud *const *links, const int32_t *link_kind, int n_links,
                                      const hop_hand_removal_params *params, hop_cloud **out) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  if (!scene || !params || !out || n_links < 0 || n_links > MAX_LINKS || (n_links > 0 && (!links || !link_kind))) {
    ctx->err = "hop_remove_hand_points: bad arguments (at most 16 links)"; return HOP_EINVAL;
  }
  cudaStream_t st = ctx->stream;
  const int n = scene->n;
  int kept = 0, rc;
  FBuf A(st), NA(st), B(st), NB(st), fl(st);
  if (n > 0) {
    FR_CUDA(A.alloc(sizeof(float4) * (size_t)n)); FR_CUDA(NA.alloc(sizeof(float4) * (size_t)n));
    FR_CUDA(B.alloc(sizeof(float4) * (size_t)n)); FR_CUDA(NB.alloc(sizeof(float4) * (size_t)n)); FR_CUDA(fl.alloc((size_t)n));
    HandRemovalArgs a;
    a.pw = scene->d_pw; a.nv = scene->d_nv; a.n = n; a.n_links = n_links;
    for (int k = 0; k < n_links; ++k) {
      a.link[k] = links[k] ? links[k]->d_pw : nullptr;
      a.link_n[k] = links[k] ? links[k]->n : 0;
      a.link_thr[k] = link_kind[k] == 1 ? (float)(0.005 * 0.005) : (link_kind[k] == 2 ? (float)(0.02 * 0.02) : params->dist_thres_sq);
    }
    a.cam_in_hb = xf_of(params->cam_in_handbase); a.hb_in_cam = xf_of(params->handbase_in_cam);
    a.hb_in_f12 = xf_of(params->handbase_in_finger_1_2); a.hb_in_f22 = xf_of(params->handbase_in_finger_2_2);
    a.min_z = params->min_z;
    a.out_pw = A.as<float4>(); a.out_nv = NA.as<float4>(); a.flag = fl.as<unsigned char>();
    {
      ProfScope ps(ctx, HOP_PROF_FRAME);
      hand_removal_kernel<<<(n + 127) / 128, 128, 0, st>>>(a);
      ctx->launches += 1;
    }
    FR_CUDA(cudaGetLastError());
    if ((rc = compact2(ctx, A.as<float4>(), NA.as<float4>(), fl.as<unsigned char>(), n, B.as<float4>(), NB.as<float4>(), &kept)) != HOP_OK) return rc;
  }
  hop_cloud *c = *out ? *out : new hop_cloud();
  rc = hop_cloud_reserve(ctx, c, kept);
  if (rc != HOP_OK) { if (!*out) hop_cloud_free(ctx, c); return rc; }
  to_cloud_kernel<<<blocks(c->n_padded), 256, 0, st>>>(B.as<float4>(), NB.as<float4>(), kept, c->n_padded, c->d_pw, c->d_nv);
  ctx->launches += 1;
  float mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
  if (kept > 0 && (rc = cloud_bounds(ctx, B.as<float4>(), kept, mn, mx)) != HOP_OK) { if (!*out) hop_cloud_free(ctx, c); return rc; }
  for (int k = 0; k < 3; ++k) { c->bbox_min[k] = mn[k]; c->bbox_max[k] = mx[k]; }
  FR_CUDA(cudaGetLastError());
  *out = c;
  return HOP_OK;
}

extern "C" int hop_cloud_download(hop_ctx *ctx, const hop_cloud *cloud, float *xyz, float *nrm, float *prob) {
  HOP_ENTER(ctx);
  if (!ctx || !cloud) return HOP_EINVAL;
  const int n = cloud->n;
  if (n <= 0) return HOP_OK;
  std::vector<float4> pw(n), nv(n);
  FR_CUDA(cudaMemcpyAsync(pw.data(), cloud->d_pw, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  FR_CUDA(cudaMemcpyAsync(nv.data(), cloud->d_nv, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  FR_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n; ++i) {
    if (xyz) { xyz[3 * i] = pw[i].x; xyz[3 * i + 1] = pw[i].y; xyz[3 * i + 2] = pw[i].z; }
    if (prob) prob[i] = pw[i].w;
    if (nrm) { nrm[3 * i] = nv[i].x; nrm[3 * i + 1] = nv[i].y; nrm[3 * i + 2] = nv[i].z; }
  }
  return HOP_OK;
}


// ====================================================================================================================
// The two PCL normal estimators of main_realdata_auto's hand branch
//   Utils::calNormalIntegralImage(scene_rgb, -1, 0.02, 10, true)   Utils.cpp:294-329, main_realdata_auto.cpp:61
//       pcl::IntegralImageNormalEstimation, SIMPLE_3D_GRADIENT, depth dependent smoothing, on the ORGANIZED frame
//   Utils::calNormalMLS(object1, 0.003)                             Utils.cpp:268-292, main_realdata_auto.cpp:160
//       pcl::MovingLeastSquares, order 2: projects every point onto its fitted surface and takes the normal there
// Restated from PCL 1.9.1 (not installed: parity unpinned against PCL; oracle/hop_oracle_frame.c is the checker).
// ====================================================================================================================
namespace {

// Utils::readDepthImage + convert3dOrganizedRGB: every pixel, invalid ones are (0, 0, 0) (Utils.cpp:103-110)
__global__ void organized_kernel(const uint16_t *__restrict__ depth_mm, int w, int h, float fx, float fy, float cx, float cy,
                                 float4 *__restrict__ pts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * h) return;
  const int u = i / w, v = i - u * w;
  float d = __double2float_rn((double)(float)depth_mm[i] * 0.001);
  if ((double)d > 2.0 || (double)d < 0.1) d = 0.f;
  float4 p = make_float4(0.f, 0.f, 0.f, 1.f);
  if ((double)d > 0.1 && (double)d < 2.0) {
    p.x = __fdiv_rn(__fmul_rn(__fsub_rn((float)v, cx), d), fx);
    p.y = __fdiv_rn(__fmul_rn(__fsub_rn((float)u, cy), d), fy);
    p.z = d;
  }
  pts[i] = p;
}

// depth-change map: a pixel and its right / lower neighbour are marked when their depths differ by more than
// max_depth_change_factor (|z| + 1) 2 or one of them is not finite (all writes store 0: order free)
__global__ void depth_change_kernel(const float4 *__restrict__ pts, int w, int h, float factor, unsigned char *__restrict__ change) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * h) return;
  const int ri = i / w, ci = i - ri * w;
  if (ri >= h - 1 || ci >= w - 1) return;
  const float depth = pts[i].z, depthR = pts[i + 1].z, depthD = pts[i + w].z;
  const float thr = __fmul_rn(__fmul_rn(factor, __fadd_rn(fabsf(depth), 1.0f)), 2.0f);
  if (fabs((double)__fsub_rn(depth, depthR)) > (double)thr || !isfinite(depth) || !isfinite(depthR)) { change[i] = 0; change[i + 1] = 0; }
  if (fabs((double)__fsub_rn(depth, depthD)) > (double)thr || !isfinite(depth) || !isfinite(depthD)) { change[i] = 0; change[i + w] = 0; }
}

__global__ void dist_init_kernel(const unsigned char *__restrict__ change, int n, float far, float *__restrict__ dist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dist[i] = change[i] == 0 ? 0.0f : far;
}

// PCL's two-pass chamfer distance (weights 1 / 1.4), exactly: the passes are sequential recurrences (a cell needs its left
// neighbour of the same row and three cells of the previous row) and the float sums of 1.0f and 1.4f round, so the order of the
// additions is part of the result ((int)distance later picks the smoothing rectangle).  One CTA walks the skewed wavefront
// t = 2 row + column: thread = row, one column per step, every dependency was produced at least one step (one barrier) earlier.
// FORWARD: rows 1 .. h-1, columns 1 .. w-1 (at the last column PCL reads previous_row[w] = the first cell of the current row).
// backward: rows h-2 .. 0, columns w-2 .. 0 (at column 0 it reads next_row[-1] = the last cell of the current row).
template <bool FORWARD>
__global__ void __launch_bounds__(1024, 1) chamfer_kernel(float *__restrict__ dist, int w, int h) {
  const int r = threadIdx.x;                 // FORWARD: row 1 + r; backward: row h - 2 - r
  const int rows = h - 1;
  const int steps = 2 * (rows - 1) + (w - 1);
  for (int base = 0; base < rows; base += blockDim.x) {   // (more rows than threads: bands, each band after the previous one)
    const int rr = base + r;
    const bool have = rr < rows;
    const int row = FORWARD ? 1 + rr : h - 2 - rr;
    float *cur = dist + (size_t)row * w;
    const float *adj = FORWARD ? cur - w : cur + w;        // the previous (forward) / next (backward) row
    const int band_rows = min((int)blockDim.x, rows - base);
    const int band_steps = 2 * (band_rows - 1) + (w - 1);
    for (int t = 0; t < band_steps; ++t) {
      const int k = t - 2 * r;                              // this row's column counter at step t
      if (have && k >= 0 && k < w - 1) {
        if (FORWARD) {
          const int ci = 1 + k;
          const float upLeft = __fadd_rn(adj[ci - 1], 1.4f), up = __fadd_rn(adj[ci], 1.0f), upRight = __fadd_rn(adj[ci + 1], 1.4f);
          const float left = __fadd_rn(cur[ci - 1], 1.0f), center = cur[ci];
          const float mv = fminf(fminf(upLeft, up), fminf(left, upRight));
          if (mv < center) cur[ci] = mv;
        } else {
          const int ci = w - 2 - k;
          const float lowerLeft = __fadd_rn(adj[ci - 1], 1.4f), lower = __fadd_rn(adj[ci], 1.0f), lowerRight = __fadd_rn(adj[ci + 1], 1.4f);
          const float right = __fadd_rn(cur[ci + 1], 1.0f), center = cur[ci];
          const float mv = fminf(fminf(lowerLeft, lower), fminf(right, lowerRight));
          if (mv < center) cur[ci] = mv;
        }
      }
      __syncthreads();
    }
  }
  (void)steps;
}

// IntegralImage2D<float, 3>: first-order sums in double.  Row pass: one thread per image row, sequential prefix (PCL's so_far).
__global__ void ii_rows_kernel(const float4 *__restrict__ pts, int w, int h, double *__restrict__ rowsum /* h x w x 3 */) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= h) return;
  double s0 = 0, s1 = 0, s2 = 0;
  for (int c = 0; c < w; ++c) {
    const float4 p = pts[(size_t)r * w + c];
    if (finite3(p.x, p.y, p.z)) { s0 += p.x; s1 += p.y; s2 += p.z; }
    double *o = rowsum + 3 * ((size_t)r * w + c);
    o[0] = s0; o[1] = s1; o[2] = s2;
  }
}
// Column pass: I[r + 1][c + 1] = I[r][c + 1] + so_far(r, c), one thread per (column, channel)
__global__ void ii_cols_kernel(const double *__restrict__ rowsum, int w, int h, double *__restrict__ I /* (h + 1) x (w + 1) x 3 */) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * (w + 1)) return;
  const int c1 = t / 3, k = t - 3 * c1;   // column of the integral image
  double acc = 0.0;
  I[3 * (size_t)c1 + k] = 0.0;
  for (int r = 0; r < h; ++r) {
    if (c1 > 0) acc += rowsum[3 * ((size_t)r * w + c1 - 1) + k];
    I[3 * ((size_t)(r + 1) * (w + 1) + c1) + k] = c1 > 0 ? acc : 0.0;
  }
}

__device__ __forceinline__ void ii_sum(const double *__restrict__ I, int W1, int sx, int sy, int ww, int hh, double *out) {
  const size_t ul = (size_t)sy * W1 + sx, ur = ul + ww, ll = (size_t)(sy + hh) * W1 + sx, lr = ll + ww;
#pragma unroll
  for (int k = 0; k < 3; ++k) out[k] = I[3 * lr + k] + I[3 * ul + k] - I[3 * ur + k] - I[3 * ll + k];
}

// computeFeatureFull + computePointNormal (SIMPLE_3D_GRADIENT), border policy IGNORE, viewpoint (0, 0, 0)
__global__ void ii_normal_kernel(const float4 *__restrict__ pts, const float *__restrict__ dist, const double *__restrict__ I, int w, int h,
                                 float smoothing_size, float4 *__restrict__ nrm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * h) return;
  const float nan = __int_as_float(0x7fc00000);
  float4 out = make_float4(nan, nan, nan, 0.f);
  const int ri = i / w, ci = i - ri * w, border = (int)smoothing_size;
  const float4 p = pts[i];
  if (ri >= border && ri < h - border && ci >= border && ci < w - border && isfinite(p.z)) {
    const float smoothing = fminf(dist[i], __fadd_rn(smoothing_size, __fdiv_rn(p.z, 10.0f)));
    if (smoothing > 2.0f) {
      const int rw = (int)smoothing, rh = rw, rw2 = rw / 2, rh2 = rh / 2, W1 = w + 1;
      double a1[3], a2[3], gx[3], gy[3];
      ii_sum(I, W1, ci + rw2, ri - rh2, 1, rh, a1); ii_sum(I, W1, ci - rw2, ri - rh2, 1, rh, a2);
#pragma unroll
      for (int k = 0; k < 3; ++k) gx[k] = a1[k] - a2[k];
      ii_sum(I, W1, ci - rw2, ri + rh2, rw, 1, a1); ii_sum(I, W1, ci - rw2, ri - rh2, rw, 1, a2);
#pragma unroll
      for (int k = 0; k < 3; ++k) gy[k] = a1[k] - a2[k];
      const double n0 = __dmul_rn(gy[1], gx[2]) - __dmul_rn(gy[2], gx[1]), n1 = __dmul_rn(gy[2], gx[0]) - __dmul_rn(gy[0], gx[2]),
                   n2 = __dmul_rn(gy[0], gx[1]) - __dmul_rn(gy[1], gx[0]);
      const double len = __dadd_rn(__dadd_rn(__dmul_rn(n0, n0), __dmul_rn(n1, n1)), __dmul_rn(n2, n2));
      if (len != 0.0) {
        const double s = sqrt(len);
        float nx = (float)(n0 / s), ny = (float)(n1 / s), nz = (float)(n2 / s);
        const float vx = __fsub_rn(0.f, p.x), vy = __fsub_rn(0.f, p.y), vz = __fsub_rn(0.f, p.z);
        const float cos_theta = __fadd_rn(__fadd_rn(__fmul_rn(vx, nx), __fmul_rn(vy, ny)), __fmul_rn(vz, nz));
        if (cos_theta < 0) { nx = -nx; ny = -ny; nz = -nz; }
        out = make_float4(nx, ny, nz, 1.f);
      }
    }
  }
  nrm[i] = out;
}

__global__ void valid_z_flag_kernel(const float4 *__restrict__ pts, int n, unsigned char *__restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const float z = pts[i].z; flag[i] = ((double)z >= 0.1 && (double)z <= 2.0) ? 1 : 0; }   // PassThrough z in [0.1, 2.0]
}

// ---- MovingLeastSquares, order 2 -----------------------------------------------------------------------------------
__device__ void eigen33_smallest(const double (&A)[3][3], double (&ev)[3]) {   // pcl::eigen33 (common/impl/eigen.hpp)
  double scale = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) scale = fmax(scale, fabs(A[i][j]));
  if (scale <= DBL_MIN) scale = 1.0;
  double S[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) S[i][j] = A[i][j] / scale;
  double roots[3];
  const double c0 = S[0][0] * S[1][1] * S[2][2] + 2.0 * S[0][1] * S[0][2] * S[1][2] - S[0][0] * S[1][2] * S[1][2] - S[1][1] * S[0][2] * S[0][2] -
                    S[2][2] * S[0][1] * S[0][1];
  const double c1 = S[0][0] * S[1][1] - S[0][1] * S[0][1] + S[0][0] * S[2][2] - S[0][2] * S[0][2] + S[1][1] * S[2][2] - S[1][2] * S[1][2];
  const double c2 = S[0][0] + S[1][1] + S[2][2];
  bool quadratic = fabs(c0) < DBL_EPSILON;
  if (!quadratic) {
    const double s_inv3 = 1.0 / 3.0, s_sqrt3 = sqrt(3.0);
    const double c2_over_3 = c2 * s_inv3;
    double a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.0) a_over_3 = 0.0;
    const double half_b = 0.5 * (c0 + c2_over_3 * (2.0 * c2_over_3 * c2_over_3 - c1));
    double q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.0) q = 0.0;
    const double rho = sqrt(-a_over_3);
    const double theta = atan2(sqrt(-q), half_b) * s_inv3;
    const double ct = cos(theta), st = sin(theta);
    roots[0] = c2_over_3 + 2.0 * rho * ct;
    roots[1] = c2_over_3 - rho * (ct + s_sqrt3 * st);
    roots[2] = c2_over_3 - rho * (ct - s_sqrt3 * st);
    if (roots[0] >= roots[1]) { const double t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
    if (roots[1] >= roots[2]) {
      double t = roots[1]; roots[1] = roots[2]; roots[2] = t;
      if (roots[0] >= roots[1]) { t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
    }
    if (roots[0] <= 0.0) quadratic = true;
  }
  if (quadratic) {
    roots[0] = 0.0;
    double d = c2 * c2 - 4.0 * c1;
    if (d < 0.0) d = 0.0;
    const double sd = sqrt(d);
    roots[2] = 0.5 * (c2 + sd); roots[1] = 0.5 * (c2 - sd);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) S[i][i] -= roots[0];
  const double v1[3] = {S[0][1] * S[1][2] - S[0][2] * S[1][1], S[0][2] * S[1][0] - S[0][0] * S[1][2], S[0][0] * S[1][1] - S[0][1] * S[1][0]};
  const double v2[3] = {S[0][1] * S[2][2] - S[0][2] * S[2][1], S[0][2] * S[2][0] - S[0][0] * S[2][2], S[0][0] * S[2][1] - S[0][1] * S[2][0]};
  const double v3[3] = {S[1][1] * S[2][2] - S[1][2] * S[2][1], S[1][2] * S[2][0] - S[1][0] * S[2][2], S[1][0] * S[2][1] - S[1][1] * S[2][0]};
  const double l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2], l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2],
               l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
  const double *v = v3; double l = l3;
  if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; } else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; }
  const double s = sqrt(l);
#pragma unroll
  for (int k = 0; k < 3; ++k) ev[k] = v[k] / s;
}

// visits every point within the radius of p (float distance test, FLANN's operation order), neighbours in cell order
template <typename F>
__device__ __forceinline__ void for_neighbours(const float4 p, const float4 *__restrict__ sorted, const int *__restrict__ cell_start,
                                               const int *__restrict__ cell_end, const CellGeom &g, float r2, F f) {
  int ca, cb, cc;
  cell_of(g, p, ca, cb, cc);
  for (int dc = -1; dc <= 1; ++dc)
    for (int db = -1; db <= 1; ++db)
      for (int da = -1; da <= 1; ++da) {
        const int a = ca + da, b = cb + db, c = cc + dc;
        if ((unsigned)a >= (unsigned)g.dim[0] || (unsigned)b >= (unsigned)g.dim[1] || (unsigned)c >= (unsigned)g.dim[2]) continue;
        const int cell = (c * g.dim[1] + b) * g.dim[0] + a;
        for (int j = cell_start[cell]; j < cell_end[cell]; ++j) {
          const float4 q = sorted[j];
          const float dx = __fsub_rn(q.x, p.x), dy = __fsub_rn(q.y, p.y), dz = __fsub_rn(q.z, p.z);
          if (__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)) <= r2) f(q);
        }
      }
}

__global__ void mls_kernel(const float4 *__restrict__ pts, int n, const float4 *__restrict__ sorted, const int *__restrict__ cell_start,
                           const int *__restrict__ cell_end, CellGeom g, float radius, float4 *__restrict__ out_p, float4 *__restrict__ out_n,
                           unsigned char *__restrict__ valid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  const float r2f = __fmul_rn(radius, radius);
  const double r2 = (double)radius * (double)radius;
  const float nan = __int_as_float(0x7fc00000);
  // computeMeanAndCovarianceMatrix: one pass, double accumulators of x, y, z and the six products
  double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  int m = 0;
  for_neighbours(p, sorted, cell_start, cell_end, g, r2f, [&](const float4 q) {
    acc[0] += (double)q.x * q.x; acc[1] += (double)q.x * q.y; acc[2] += (double)q.x * q.z;
    acc[3] += (double)q.y * q.y; acc[4] += (double)q.y * q.z; acc[5] += (double)q.z * q.z;
    acc[6] += q.x; acc[7] += q.y; acc[8] += q.z;
    ++m;
  });
  valid[i] = m >= 3 ? 1 : 0;
  out_p[i] = p;
  out_n[i] = make_float4(nan, nan, nan, 0.f);
  if (m < 3) return;
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] /= (double)m;
  double C[3][3];
  C[0][0] = acc[0] - acc[6] * acc[6]; C[0][1] = acc[1] - acc[6] * acc[7]; C[0][2] = acc[2] - acc[6] * acc[8];
  C[1][1] = acc[3] - acc[7] * acc[7]; C[1][2] = acc[4] - acc[7] * acc[8]; C[2][2] = acc[5] - acc[8] * acc[8];
  C[1][0] = C[0][1]; C[2][0] = C[0][2]; C[2][1] = C[1][2];
  double pn[3];
  eigen33_smallest(C, pn);
  if (!isfinite(pn[0]) || !isfinite(pn[1]) || !isfinite(pn[2])) return;
  const double d4 = -(pn[0] * acc[6] + pn[1] * acc[7] + pn[2] * acc[8]);
  const double distance = (double)p.x * pn[0] + (double)p.y * pn[1] + (double)p.z * pn[2] + d4;
  const double mean[3] = {(double)p.x - distance * pn[0], (double)p.y - distance * pn[1], (double)p.z - distance * pn[2]};
  double va[3], ua[3];
  if (!(fabs(pn[0]) <= 1e-12 * fabs(pn[2])) || !(fabs(pn[1]) <= 1e-12 * fabs(pn[2]))) {   // Eigen unitOrthogonal()
    const double inv = 1.0 / sqrt(pn[0] * pn[0] + pn[1] * pn[1]);
    va[0] = -pn[1] * inv; va[1] = pn[0] * inv; va[2] = 0.0;
  } else {
    const double inv = 1.0 / sqrt(pn[1] * pn[1] + pn[2] * pn[2]);
    va[0] = 0.0; va[1] = -pn[2] * inv; va[2] = pn[1] * inv;
  }
  ua[0] = pn[1] * va[2] - pn[2] * va[1]; ua[1] = pn[2] * va[0] - pn[0] * va[2]; ua[2] = pn[0] * va[1] - pn[1] * va[0];
  double c6[6] = {0, 0, 0, 0, 0, 0};
  bool have_poly = false;
  if (m >= 6) {
    double A[21], rhs[6];   // upper triangle of P W P^T, row-major; P W f
#pragma unroll
    for (int k = 0; k < 21; ++k) A[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) rhs[k] = 0.0;
    for_neighbours(p, sorted, cell_start, cell_end, g, r2f, [&](const float4 q) {
      const double de[3] = {(double)q.x - mean[0], (double)q.y - mean[1], (double)q.z - mean[2]};
      const double wgt = exp(-(de[0] * de[0] + de[1] * de[1] + de[2] * de[2]) / r2);
      const double u = de[0] * ua[0] + de[1] * ua[1] + de[2] * ua[2], v = de[0] * va[0] + de[1] * va[1] + de[2] * va[2];
      const double f = de[0] * pn[0] + de[1] * pn[1] + de[2] * pn[2];
      const double P[6] = {1.0, v, v * v, u, u * v, u * u};   // (ui, vi): (0,0) (0,1) (0,2) (1,0) (1,1) (2,0)
      int e = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        rhs[a] += P[a] * wgt * f;
#pragma unroll
        for (int b = a; b < 6; ++b) A[e++] += P[a] * wgt * P[b];
      }
    });
    // LLT solve
    double L[6][6];
    bool ok = true;
    {
      auto at = [&](int a, int b) { const int lo = a < b ? a : b, hi = a < b ? b : a; return A[lo * 6 - lo * (lo - 1) / 2 + (hi - lo)]; };
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        double d = at(j, j);
#pragma unroll
        for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
        if (!(d > 0.0)) { ok = false; d = 1.0; }
        L[j][j] = sqrt(d);
#pragma unroll
        for (int r = j + 1; r < 6; ++r) {
          double s = at(r, j);
#pragma unroll
          for (int k = 0; k < j; ++k) s -= L[r][k] * L[j][k];
          L[r][j] = s / L[j][j];
        }
      }
    }
    if (ok) {
#pragma unroll
      for (int r = 0; r < 6; ++r) { double s = rhs[r]; for (int k = 0; k < r; ++k) s -= L[r][k] * rhs[k]; rhs[r] = s / L[r][r]; }
#pragma unroll
      for (int r = 5; r >= 0; --r) { double s = rhs[r]; for (int k = r + 1; k < 6; ++k) s -= L[k][r] * rhs[k]; rhs[r] = s / L[r][r]; }
#pragma unroll
      for (int k = 0; k < 6; ++k) c6[k] = rhs[k];
      have_poly = isfinite(c6[0]);
    }
  }
  double pt[3], nv[3];
  if (have_poly) {   // projectPointSimpleToPolynomialSurface at (u, v) = (0, 0)
#pragma unroll
    for (int k = 0; k < 3; ++k) { pt[k] = mean[k] + c6[0] * pn[k]; nv[k] = pn[k] - c6[3] * ua[k] - c6[1] * va[k]; }
    const double l = sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) nv[k] /= l;
  } else {
#pragma unroll
    for (int k = 0; k < 3; ++k) { pt[k] = mean[k]; nv[k] = pn[k]; }
  }
  out_p[i] = make_float4((float)pt[0], (float)pt[1], (float)pt[2], p.w);
  out_n[i] = make_float4((float)nv[0], (float)nv[1], (float)nv[2], 1.f);
}

}  // namespace

// Utils::readDepthImage + convert3dOrganizedRGB + calNormalIntegralImage(-1, 0.02, 10, true) + PassThrough(z, 0.1, 2.0)
// (main_realdata_auto.cpp:54-70): the frame's valid pixels in raster order with their integral-image normals (NaN where PCL
// writes none) -- the cloud `scene_organized` / `scene_rgb` of the reference before its voxel grid.
extern "C" int hop_frame_organized(hop_ctx *ctx, const uint16_t *depth_mm, int width, int height, const hop_frame_params *fp, float max_depth_change_factor,
                                   float normal_smoothing_size, hop_cloud **out) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  if (!depth_mm || width < 3 || height < 3 || !fp || !out || !(normal_smoothing_size >= 1.f)) { ctx->err = "hop_frame_organized: bad arguments"; return HOP_EINVAL; }
  ProfScope ps(ctx, HOP_PROF_FRAME);
  cudaStream_t st = ctx->stream;
  const int npx = width * height;
  FBuf d_depth(st), P(st), N(st), ch(st), dist(st), rows(st), I(st), fl(st), CP(st), CN(st);
  FR_CUDA(d_depth.alloc(sizeof(uint16_t) * (size_t)npx)); FR_CUDA(P.alloc(sizeof(float4) * (size_t)npx)); FR_CUDA(N.alloc(sizeof(float4) * (size_t)npx));
  FR_CUDA(ch.alloc((size_t)npx)); FR_CUDA(dist.alloc(sizeof(float) * (size_t)npx)); FR_CUDA(fl.alloc((size_t)npx));
  FR_CUDA(rows.alloc(sizeof(double) * 3 * (size_t)npx)); FR_CUDA(I.alloc(sizeof(double) * 3 * (size_t)(width + 1) * (height + 1)));
  FR_CUDA(CP.alloc(sizeof(float4) * (size_t)npx)); FR_CUDA(CN.alloc(sizeof(float4) * (size_t)npx));
  FR_CUDA(cudaMemcpyAsync(d_depth.p, depth_mm, sizeof(uint16_t) * (size_t)npx, cudaMemcpyHostToDevice, st));
  organized_kernel<<<blocks(npx), 256, 0, st>>>(d_depth.as<uint16_t>(), width, height, fp->fx, fp->fy, fp->cx, fp->cy, P.as<float4>());
  FR_CUDA(cudaMemsetAsync(ch.p, 255, (size_t)npx, st));
  depth_change_kernel<<<blocks(npx), 256, 0, st>>>(P.as<float4>(), width, height, max_depth_change_factor, ch.as<unsigned char>());
  dist_init_kernel<<<blocks(npx), 256, 0, st>>>(ch.as<unsigned char>(), npx, (float)(width + height), dist.as<float>());
  const int cham_threads = std::min(1024, ((height - 1) + 31) / 32 * 32);
  chamfer_kernel<true><<<1, cham_threads, 0, st>>>(dist.as<float>(), width, height);
  chamfer_kernel<false><<<1, cham_threads, 0, st>>>(dist.as<float>(), width, height);
  ii_rows_kernel<<<(height + 63) / 64, 64, 0, st>>>(P.as<float4>(), width, height, rows.as<double>());
  ii_cols_kernel<<<(3 * (width + 1) + 63) / 64, 64, 0, st>>>(rows.as<double>(), width, height, I.as<double>());
  ii_normal_kernel<<<blocks(npx), 256, 0, st>>>(P.as<float4>(), dist.as<float>(), I.as<double>(), width, height, normal_smoothing_size, N.as<float4>());
  valid_z_flag_kernel<<<blocks(npx), 256, 0, st>>>(P.as<float4>(), npx, fl.as<unsigned char>());
  ctx->launches += 9;
  int n = 0, rc;
  if ((rc = compact2(ctx, P.as<float4>(), N.as<float4>(), fl.as<unsigned char>(), npx, CP.as<float4>(), CN.as<float4>(), &n)) != HOP_OK) return rc;
  return make_cloud(ctx, CP.as<float4>(), CN.as<float4>(), n, out);
}

// Utils::calNormalMLS (Utils.cpp:268-292): pcl::MovingLeastSquares (order 2, radius search, normals): the points with at least
// three neighbours, projected onto their fitted surfaces, with the surface normals there (not oriented); the weight channel
// (confidence) travels with the points.  Points keep their input order.
extern "C" int hop_cloud_mls(hop_ctx *ctx, const hop_cloud *in, float radius, hop_cloud **out) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  if (!in || !out || *out == in || !(radius > 0.f)) { ctx->err = "hop_cloud_mls: bad arguments"; return HOP_EINVAL; }
  cudaStream_t st = ctx->stream;
  const int n = in->n;
  FBuf OP(st), ON(st), fl(st), CP(st), CN(st);
  int kept = 0, rc;
  if (n > 0) {
    ProfScope ps(ctx, HOP_PROF_FRAME);
    float mn[3], mx[3];
    if ((rc = cloud_bounds(ctx, in->d_pw, n, mn, mx)) != HOP_OK) return rc;
    CellGeom g;
    g.inv = 1.f / radius;
    double ncell = 1;
    for (int k = 0; k < 3; ++k) {
      g.mn[k] = (int)std::floor(mn[k] * g.inv);
      g.dim[k] = (int)std::floor(mx[k] * g.inv) - g.mn[k] + 1;
      ncell *= g.dim[k];
    }
    if (!(ncell < 2.5e8)) { ctx->err = "hop_cloud_mls: radius too small for the cloud's extent"; return HOP_EINVAL; }
    const int nce = (int)ncell;
    FBuf k0(st), k1(st), i0(st), i1(st), cs(st), ce(st), sorted(st), tmp(st);
    FR_CUDA(k0.alloc(sizeof(unsigned int) * (size_t)n)); FR_CUDA(k1.alloc(sizeof(unsigned int) * (size_t)n));
    FR_CUDA(i0.alloc(sizeof(int) * (size_t)n)); FR_CUDA(i1.alloc(sizeof(int) * (size_t)n));
    FR_CUDA(cs.alloc(sizeof(int) * (size_t)nce)); FR_CUDA(ce.alloc(sizeof(int) * (size_t)nce)); FR_CUDA(sorted.alloc(sizeof(float4) * (size_t)n));
    FR_CUDA(OP.alloc(sizeof(float4) * (size_t)n)); FR_CUDA(ON.alloc(sizeof(float4) * (size_t)n)); FR_CUDA(fl.alloc((size_t)n));
    FR_CUDA(CP.alloc(sizeof(float4) * (size_t)n)); FR_CUDA(CN.alloc(sizeof(float4) * (size_t)n));
    FR_CUDA(cudaMemsetAsync(cs.p, 0, sizeof(int) * (size_t)nce, st)); FR_CUDA(cudaMemsetAsync(ce.p, 0, sizeof(int) * (size_t)nce, st));
    cell_key_kernel<<<blocks(n), 256, 0, st>>>(in->d_pw, n, g, k0.as<unsigned int>(), i0.as<int>());
    size_t sort_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, k0.as<unsigned int>(), k1.as<unsigned int>(), i0.as<int>(), i1.as<int>(), n, 0, 32, st);
    FR_CUDA(tmp.alloc(sort_bytes));
    cub::DeviceRadixSort::SortPairs(tmp.p, sort_bytes, k0.as<unsigned int>(), k1.as<unsigned int>(), i0.as<int>(), i1.as<int>(), n, 0, 32, st);
    cell_range_kernel<<<blocks(n), 256, 0, st>>>(k1.as<unsigned int>(), i1.as<int>(), in->d_pw, n, cs.as<int>(), ce.as<int>(), sorted.as<float4>());
    mls_kernel<<<(n + 63) / 64, 64, 0, st>>>(in->d_pw, n, sorted.as<float4>(), cs.as<int>(), ce.as<int>(), g, radius, OP.as<float4>(), ON.as<float4>(), fl.as<unsigned char>());
    ctx->launches += 4;
    if ((rc = compact2(ctx, OP.as<float4>(), ON.as<float4>(), fl.as<unsigned char>(), n, CP.as<float4>(), CN.as<float4>(), &kept)) != HOP_OK) return rc;
  }
  return make_cloud(ctx, CP.as<float4>(), CN.as<float4>(), kept, out);
}
