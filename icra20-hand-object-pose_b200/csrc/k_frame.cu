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
__global__ void backproject_kernel(const uint16_t *__restrict__ depth_mm, int w, int h, float fx, float fy, float cx, float cy,
                                   float4 *__restrict__ pts, unsigned char *__restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * h) return;
  const int u = i / w, v = i - u * w;
  float d = __fmul_rn((float)depth_mm[i], 0.001f);
  if ((double)d > 2.0 || (double)d < 0.1) d = 0.f;
  const bool ok = (double)d > 0.1 && (double)d < 2.0;
  float4 p = make_float4(0.f, 0.f, 0.f, 1.f);
  if (ok) {
    p.x = __fdiv_rn(__fmul_rn(__fsub_rn((float)v, cx), d), fx);
    p.y = __fdiv_rn(__fmul_rn(__fsub_rn((float)u, cy), d), fy);
    p.z = d;
  }
  pts[i] = p;
  flag[i] = ok ? 1 : 0;
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
                            unsigned char *__restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  const float x = xf_row(T, 0, p.x, p.y, p.z), y = xf_row(T, 1, p.x, p.y, p.z), z = xf_row(T, 2, p.x, p.y, p.z);
  const bool ok = finite3(x, y, z) && !(z < lo.z || z > hi.z) && !(x < lo.x || x > hi.x) && !(y < lo.y || y > hi.y);
  out[i] = make_float4(xf_row(Ti, 0, x, y, z), xf_row(Ti, 1, x, y, z), xf_row(Ti, 2, x, y, z), p.w);
  flag[i] = ok ? 1 : 0;
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
                              float4 *__restrict__ out_nrm, unsigned char *__restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  float4 q = nrm[i];
  const bool ok = finite3(p.x, p.y, p.z) && finite3(q.x, q.y, q.z);
  if (__fsub_rn(__fsub_rn(__fmul_rn(-p.x, q.x), __fmul_rn(p.y, q.y)), __fmul_rn(p.z, q.z)) < 0.f) { q.x = -q.x; q.y = -q.y; q.z = -q.z; }
  out_pts[i] = make_float4(p.x, p.y, p.z, 1.f);
  out_nrm[i] = q;
  flag[i] = ok ? 1 : 0;
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
int compact2(hop_ctx *ctx, const float4 *a, const float4 *b, const unsigned char *flag, int n, float4 *oa, float4 *ob, int *kept) {
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
int voxel_grid(hop_ctx *ctx, const float4 *pts, const float4 *nrm, int n, float leaf, float4 *out_pts, float4 *out_nrm, int *m_out) {
  cudaStream_t st = ctx->stream;
  *m_out = 0;
  if (n <= 0) return HOP_OK;
  float mn[3], mx[3];
  int rc = cloud_bounds(ctx, pts, n, mn, mx);
  if (rc != HOP_OK) return rc;
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
  // 1. back-projection, raster order
  backproject_kernel<<<blocks(npx), 256, 0, st>>>(d_depth.as<uint16_t>(), width, height, fp->fx, fp->fy, fp->cx, fp->cy, A.as<float4>(), fl.as<unsigned char>());
  ctx->launches += 1;
  int n = 0, rc;
  if ((rc = compact2(ctx, A.as<float4>(), nullptr, fl.as<unsigned char>(), npx, B.as<float4>(), nullptr, &n)) != HOP_OK) return rc;
  counts[0] = n;
  // 2. dense voxel grid
  int m = 0;
  if ((rc = voxel_grid(ctx, B.as<float4>(), nullptr, n, fp->leaf_dense, A.as<float4>(), nullptr, &m)) != HOP_OK) return rc;
  counts[1] = m;
  // 3. crop in the hand-base frame
  int nc = 0;
  if (m > 0) {
    crop_kernel<<<blocks(m), 256, 0, st>>>(A.as<float4>(), m, rows_of(fp->cam_in_handbase), rows_of(fp->handbase_in_cam),
                                           make_float3(fp->box_min[0], fp->box_min[1], fp->box_min[2]), make_float3(fp->box_max[0], fp->box_max[1], fp->box_max[2]),
                                           B.as<float4>(), fl.as<unsigned char>());
    ctx->launches += 1;
    if ((rc = compact2(ctx, B.as<float4>(), nullptr, fl.as<unsigned char>(), m, A.as<float4>(), nullptr, &nc)) != HOP_OK) return rc;
  }
  counts[2] = nc;
  // 4. normals over normal_radius
  int mo = 0, nf = 0;
  FR_CUDA(NA.alloc(sizeof(float4) * (size_t)std::max(nc, 1))); FR_CUDA(NB.alloc(sizeof(float4) * (size_t)std::max(nc, 1)));
  if (nc > 0) {
    float mn[3], mx[3];
    if ((rc = cloud_bounds(ctx, A.as<float4>(), nc, mn, mx)) != HOP_OK) return rc;
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
    if ((rc = voxel_grid(ctx, A.as<float4>(), NA.as<float4>(), nc, fp->leaf_object, B.as<float4>(), NB.as<float4>(), &mo)) != HOP_OK) return rc;
    if (mo > 0) {
      finish_kernel<<<blocks(mo), 256, 0, st>>>(B.as<float4>(), NB.as<float4>(), mo, A.as<float4>(), NA.as<float4>(), fl.as<unsigned char>());
      ctx->launches += 1;
      if ((rc = compact2(ctx, A.as<float4>(), NA.as<float4>(), fl.as<unsigned char>(), mo, B.as<float4>(), NB.as<float4>(), &nf)) != HOP_OK) return rc;
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
  float mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
  if (nf > 0 && (rc = cloud_bounds(ctx, B.as<float4>(), nf, mn, mx)) != HOP_OK) { if (!*scene) hop_cloud_free(ctx, c); return rc; }
  for (int k = 0; k < 3; ++k) { c->bbox_min[k] = mn[k]; c->bbox_max[k] = mx[k]; }
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
