// k_sdf.cu -- physics pruning of pose hypotheses: signed distance of point sets to triangle meshes and the
// whole accept / reject decision of PoseEstimator::rejectByCollisionOrNonTouching for a batch of hypotheses in one launch.
//
//   replaces PoseEstimator::rejectByCollisionOrNonTouching        src/perception/src/PoseEstimator.cpp:524-735
//            SDFchecker::getSignedDistanceMinMaxWithRegistered     src/perception/src/SDFchecker.cpp:115-134
//            = igl::signed_distance(..., SIGNED_DISTANCE_TYPE_PSEUDONORMAL, ...)   src/perception/include/igl/signed_distance.cpp:90-101,160-205
//              closest point  igl/point_simplex_squared_distance.cpp:44-108 (Ericson's region walk)
//              sign           igl/pseudonormal_test.cpp:24-130
//            SDFchecker::transformMesh per hypothesis              SDFchecker.cpp:80-86
//
// The reference moves the object MESH into every hypothesis' frame (and back), lets igl rebuild every normal and an AABB tree,
// and queries a few hundred points, inside an OpenMP loop over hypotheses.  Here no mesh ever moves: query points are carried
// into the mesh frame by the inverse placement (rigid: same distances, same signs), the normals are built once at upload, and
// one CTA per hypothesis walks the reference's decision sequence with CTA-uniform early exits.  The meshes on this path are
// convex hulls / object meshes of 10^2..10^3 faces that stay in L1: a thread scans all faces for its point, culled by a
// per-face bounding sphere (points are visited in Morton order so the lanes of a warp cull alike).
// float like the reference (Eigen::MatrixXf); which pseudonormal signs the distance follows igl's rules literally
// (barycentric classification above MIN_DOUBLE_AREA, exact vertex / edge tests below it).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <map>
#include <vector>

#include "hop_common.cuh"

struct hop_mesh {
  int nf = 0, nv = 0;
  float4 *d_hot = nullptr;    // 4 float4 per face: a, b, c, bounding sphere (centre, radius)
  float4 *d_cold = nullptr;   // 7 float4 per face: face normal (w = 1 when doublearea > 1e-4), vertex normals of a/b/c,
                              //                    normals of the edges opposite a/b/c
};

namespace {

constexpr int SDF_THREADS = 128;
constexpr int HOT_F4 = 4, COLD_F4 = 7;

struct MeshDev { const float4 *hot, *cold; int nf; };

__device__ __forceinline__ float3 f3(float4 v) { return make_float3(v.x, v.y, v.z); }
__device__ __forceinline__ float3 sub3(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 madd3(float3 a, float3 d, float t) { return make_float3(a.x + d.x * t, a.y + d.y * t, a.z + d.z * t); }

// closest point of triangle (a,b,c) to p
__device__ __forceinline__ float3 closest_on_triangle(float3 p, float3 a, float3 b, float3 c) {
  const float3 ab = sub3(b, a), ac = sub3(c, a), ap = sub3(p, a);
  const float d1 = dot3(ab, ap), d2 = dot3(ac, ap);
  if (d1 <= 0.f && d2 <= 0.f) return a;
  const float3 bp = sub3(p, b);
  const float d3 = dot3(ab, bp), d4 = dot3(ac, bp);
  if (d3 >= 0.f && d4 <= d3) return b;
  const float vc = d1 * d4 - d3 * d2;
  if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f && (a.x != b.x || a.y != b.y || a.z != b.z)) return madd3(a, ab, d1 / (d1 - d3));
  const float3 cp = sub3(p, c);
  const float d5 = dot3(ab, cp), d6 = dot3(ac, cp);
  if (d6 >= 0.f && d5 <= d6) return c;
  const float vb = d5 * d2 - d1 * d6;
  if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) return madd3(a, ac, d2 / (d2 - d6));
  const float va = d3 * d6 - d5 * d4;
  if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) return madd3(b, sub3(c, b), (d4 - d3) / ((d4 - d3) + (d5 - d6)));
  const float denom = 1.f / (va + vb + vc);
  return madd3(madd3(a, ab, vb * denom), ac, vc * denom);
}

// faces [f0, f1) stepping by `step`: the closest face (first among exact ties) and its squared distance
__device__ __forceinline__ void nearest_face(const MeshDev &m, float3 q, int f0, int step, float &best, int &bf) {
  float reach = FLT_MAX;   // sqrt(best), kept for the sphere test
  for (int f = f0; f < m.nf; f += step) {
    const float4 sp = __ldg(m.hot + (size_t)f * HOT_F4 + 3);
    const float3 dc = sub3(q, f3(sp));
    const float t = reach + sp.w;
    if (dot3(dc, dc) > t * t * 1.000001f) continue;   // the whole face is farther than the best so far
    const float3 a = f3(__ldg(m.hot + (size_t)f * HOT_F4)), b = f3(__ldg(m.hot + (size_t)f * HOT_F4 + 1)), c = f3(__ldg(m.hot + (size_t)f * HOT_F4 + 2));
    const float3 d = sub3(q, closest_on_triangle(q, a, b, c));
    const float d2 = dot3(d, d);
    if (d2 < best) { best = d2; bf = f; reach = sqrtf(d2); }
  }
}

// igl::pseudonormal_test on face f: +1 / -1
__device__ float pseudonormal_sign(const MeshDev &m, int f, float3 q) {
  const float3 A = f3(m.hot[(size_t)f * HOT_F4]), B = f3(m.hot[(size_t)f * HOT_F4 + 1]), C = f3(m.hot[(size_t)f * HOT_F4 + 2]);
  const float4 *cold = m.cold + (size_t)f * COLD_F4;
  const float4 fn = cold[0];
  const float3 c = closest_on_triangle(q, A, B, C);
  int pick = 0;   // 0 face, 1..3 vertex, 4..6 edge opposite corner
  const double eps = 1e-12;
  if (fn.w != 0.f) {
    const float3 v0 = sub3(B, A), v1 = sub3(C, A), v2 = sub3(c, A);
    const float d00 = dot3(v0, v0), d01 = dot3(v0, v1), d11 = dot3(v1, v1), d20 = dot3(v2, v0), d21 = dot3(v2, v1);
    const float denom = d00 * d11 - d01 * d01;
    float b[3];
    b[1] = (d11 * d20 - d01 * d21) / denom;
    b[2] = (d00 * d21 - d01 * d20) / denom;
    b[0] = 1.0f - (b[1] + b[2]);
    const int type = ((double)b[0] <= eps) + ((double)b[1] <= eps) + ((double)b[2] <= eps);
    if (type == 2) pick = (double)b[0] > eps ? 1 : ((double)b[1] > eps ? 2 : 3);
    else if (type == 1) pick = (double)b[0] <= eps ? 4 : ((double)b[1] <= eps ? 5 : 6);
  } else {
    const float3 P[3] = {A, B, C};
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      const float3 d = sub3(c, P[v]);
      if (pick == 0 && (double)sqrtf(dot3(d, d)) < eps) pick = 1 + v;
    }
#pragma unroll
    for (int e = 0; e < 3; ++e) {
      if (pick != 0) continue;
      const float3 s = P[(e + 1) % 3], d = P[(e + 2) % 3];
      const float3 dms = sub3(d, s), smp = sub3(s, c);
      const double t = -(double)dot3(dms, smp) / (double)dot3(dms, dms);
      float3 r = make_float3(c.x - (float)((1 - t) * (double)s.x + t * (double)d.x), c.y - (float)((1 - t) * (double)s.y + t * (double)d.y),
                             c.z - (float)((1 - t) * (double)s.z + t * (double)d.z));
      if (t < 0) r = sub3(c, s); else if (t > 1) r = sub3(c, d);
      if (sqrt((double)dot3(r, r)) < eps) pick = 4 + e;
    }
  }
  const float3 n = f3(cold[pick]);
  return dot3(sub3(q, c), n) >= 0.f ? 1.f : -1.f;
}

__device__ __forceinline__ float sdf_point(const MeshDev &m, float3 q, int *face = nullptr) {
  float best = FLT_MAX; int bf = -1;
  nearest_face(m, q, 0, 1, best, bf);
  if (face) *face = bf;
  if (bf < 0) return FLT_MAX;
  // signed_distance.cpp:127-128,158: with SDFchecker's (-FLT_MAX, FLT_MAX) bounds a point exactly ON the mesh is "out of
  // bounds" (sqrd <= low_sqr_d = 0) and gets NaN; the min / max / inside reductions below skip it (fminf, s < 0)
  if (best == 0.f) return __int_as_float(0x7fc00000);
  return pseudonormal_sign(m, bf, q) * sqrtf(best);
}

__device__ __forceinline__ float3 xform(const float *T /*3x4 row-major*/, float3 p) {
  return make_float3(T[0] * p.x + T[1] * p.y + T[2] * p.z + T[3], T[4] * p.x + T[5] * p.y + T[6] * p.z + T[7],
                     T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11]);
}

__device__ __forceinline__ void atomic_min_float(float *a, float v) {
  if (v >= 0.f) atomicMin((int *)a, __float_as_int(v)); else atomicMax((unsigned int *)a, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float *a, float v) {
  if (v >= 0.f) atomicMax((int *)a, __float_as_int(v)); else atomicMin((unsigned int *)a, __float_as_uint(v));
}

// ---- generic query: H placements x n points ---------------------------------------------------------------------------
struct SdfArgs {
  MeshDev mesh;
  const float4 *pts; int n;
  const float *xf;    // H x 12: rows of the point transform (row-major 3x4), or null = identity
  float *S; int32_t *I;   // H x n or null
  float *mn, *mx;     // H (initialised to +/- FLT_MAX)
  int *inside;        // H (zeroed)
};

__global__ void __launch_bounds__(SDF_THREADS) sdf_kernel(SdfArgs a) {
  const int h = blockIdx.y, i = blockIdx.x * SDF_THREADS + threadIdx.x;
  const bool live = i < a.n;
  float s = 0.f;
  if (live) {
    const float4 p = a.pts[i];
    float3 q = make_float3(p.x, p.y, p.z);
    if (a.xf) q = xform(a.xf + 12 * (size_t)h, q);
    int face;
    s = sdf_point(a.mesh, q, &face);
    if (a.S) a.S[(size_t)h * a.n + i] = s;
    if (a.I) a.I[(size_t)h * a.n + i] = face;
  }
  const bool num = live && s == s;   // NaN (a point on the mesh) takes no part in min / max
  float lo = num ? s : FLT_MAX, hi = num ? s : -FLT_MAX;
  int in = (live && s < 0.f) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    in += __shfl_xor_sync(0xffffffffu, in, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (lo != FLT_MAX) atomic_min_float(&a.mn[h], lo);
    if (hi != -FLT_MAX) atomic_max_float(&a.mx[h], hi);
    if (in) atomicAdd(&a.inside[h], in);
  }
}

__global__ void sdf_init_kernel(float *mn, float *mx, int *inside, int H) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  mn[h] = FLT_MAX; mx[h] = -FLT_MAX; inside[h] = 0;
}

// ---- the reject decision: one CTA per hypothesis ----------------------------------------------------------------------
struct CollisionArgs {
  MeshDev object, finger_mesh[4];
  const float4 *finger_pts[4]; int finger_n[4];
  const float4 *scene; int ns;
  const float4 *hand; int nh;
  const float4 *model; int nm;
  const float *poses;   // H x 16 column-major (model -> camera)
  hop_collision_params p;
  int32_t *keep, *reason;
  float *diag;          // H x 10 or null
};

__device__ __forceinline__ unsigned long long pack_key(float d2, int idx) { return ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned int)idx; }

// min over the CTA of a 64-bit key; every thread gets the result
__device__ __forceinline__ unsigned long long block_min_key(unsigned long long k, unsigned long long *s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, k, o); k = t < k ? t : k; }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = k;
  __syncthreads();
  unsigned long long r = s_red[0];
#pragma unroll
  for (int w = 1; w < SDF_THREADS / 32; ++w) r = s_red[w] < r ? s_red[w] : r;
  return r;
}

// nearest cloud point to c (first among ties): index and squared distance
__device__ __forceinline__ int block_nn(const float4 *pts, int n, float3 c, float &d2out, unsigned long long *s_red) {
  unsigned long long k = ~0ull;
  for (int i = threadIdx.x; i < n; i += SDF_THREADS) {
    const float4 p = __ldg(pts + i);
    const float3 d = sub3(f3(p), c);
    const unsigned long long t = pack_key(dot3(d, d), i);
    k = t < k ? t : k;
  }
  k = block_min_key(k, s_red);
  d2out = __uint_as_float((unsigned int)(k >> 32));
  return (int)(unsigned int)(k & 0xffffffffu);
}

// signed distance of ONE point, the faces spread over the CTA
__device__ __forceinline__ float block_sdf_point(const MeshDev &m, float3 q, unsigned long long *s_red) {
  float best = FLT_MAX; int bf = 0x7fffffff;
  nearest_face(m, q, threadIdx.x, SDF_THREADS, best, bf);
  const unsigned long long k = block_min_key(pack_key(best, bf), s_red);
  const int f = (int)(unsigned int)(k & 0xffffffffu);
  const float d2 = __uint_as_float((unsigned int)(k >> 32));
  if (d2 == 0.f) return __int_as_float(0x7fc00000);   // on the mesh: NaN, as in sdf_point
  return pseudonormal_sign(m, f, q) * sqrtf(d2);
}

__global__ void __launch_bounds__(SDF_THREADS) collision_kernel(CollisionArgs a) {
  __shared__ float s_M[12], s_Minv[12], s_ctr[3];
  __shared__ unsigned long long s_red[SDF_THREADS / 32];
  __shared__ float s_min[8];
  __shared__ int s_inside[4];
  const int h = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    // model -> hand base = cam2handbase * pose (float, like Eigen); its inverse carries hand-base points into the mesh frame
    const float *A = a.p.cam2handbase, *B = a.poses + 16 * (size_t)h;
    float M[16];
    for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) {
      float s = 0.f;
      for (int k = 0; k < 4; ++k) s += A[r + 4 * k] * B[k + 4 * c];
      M[r + 4 * c] = s;
    }
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) s_M[4 * r + c] = M[r + 4 * c];
    // affine inverse through the adjugate of the 3x3 block
    const float m00 = M[0], m01 = M[4], m02 = M[8], m10 = M[1], m11 = M[5], m12 = M[9], m20 = M[2], m21 = M[6], m22 = M[10];
    const float c00 = m11 * m22 - m12 * m21, c01 = m02 * m21 - m01 * m22, c02 = m01 * m12 - m02 * m11;
    const float c10 = m12 * m20 - m10 * m22, c11 = m00 * m22 - m02 * m20, c12 = m02 * m10 - m00 * m12;
    const float c20 = m10 * m21 - m11 * m20, c21 = m01 * m20 - m00 * m21, c22 = m00 * m11 - m01 * m10;
    const float id = 1.f / (m00 * c00 + m01 * c10 + m02 * c20);
    const float I[9] = {c00 * id, c01 * id, c02 * id, c10 * id, c11 * id, c12 * id, c20 * id, c21 * id, c22 * id};
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) s_Minv[4 * r + c] = I[3 * r + c];
      s_Minv[4 * r + 3] = -(I[3 * r] * M[12] + I[3 * r + 1] * M[13] + I[3 * r + 2] * M[14]);
    }
    const float3 ctr = xform(s_M, make_float3(a.p.model_center[0], a.p.model_center[1], a.p.model_center[2]));
    s_ctr[0] = ctr.x; s_ctr[1] = ctr.y; s_ctr[2] = ctr.z;
  }
  if (tid < 8) s_min[tid] = FLT_MAX;
  if (tid < 4) s_inside[tid] = 0;
  __syncthreads();
  const float3 ctr = make_float3(s_ctr[0], s_ctr[1], s_ctr[2]);
  float *dg = a.diag ? a.diag + 10 * (size_t)h : nullptr;
  if (dg && tid < 10) dg[tid] = FLT_MAX;
  int why = 0;

  // 1. the scene point nearest to the object's centre must not lie deep inside the object (PoseEstimator.cpp:596-615)
  if (a.ns > 0) {
    float d2;
    const int i = block_nn(a.scene, a.ns, ctr, d2, s_red);
    const float s = block_sdf_point(a.object, xform(s_Minv, f3(__ldg(a.scene + i))), s_red);
    if (dg && tid == 0) dg[0] = s;
    if (s <= a.p.inside_ob_dist) why = 1;
  }
  // 2. quick check with the hand point nearest to the object's centre (:618-641)
  if (!why && a.nh > 0) {
    float d2;
    const int i = block_nn(a.hand, a.nh, ctr, d2, s_red);
    if (sqrtf(d2) < a.p.ob_diameter / 2) {
      const float s = block_sdf_point(a.object, xform(s_Minv, f3(__ldg(a.hand + i))), s_red);
      if (dg && tid == 0) dg[1] = s;
      if (s < a.p.collision_dist) why = 2;
    }
  }
  // 3. finger clouds against the object (:645-668): min signed distance per finger link
  if (!why) {
    int off[5]; off[0] = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool use = a.finger_n[k] > 0 && (k < 2 ? a.p.finger_status[0] : a.p.finger_status[2]);
      off[k + 1] = off[k] + (use ? a.finger_n[k] : 0);
    }
    float mn[4] = {FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX};
    for (int j = tid; j < off[4]; j += SDF_THREADS) {
      const int k = j < off[1] ? 0 : (j < off[2] ? 1 : (j < off[3] ? 2 : 3));
      const float s = sdf_point(a.object, xform(s_Minv, f3(__ldg(a.finger_pts[k] + (j - off[k])))));
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) if (kk == k) mn[kk] = fminf(mn[kk], s);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v = mn[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
      if ((tid & 31) == 0 && v != FLT_MAX) atomic_min_float(&s_min[k], v);
    }
    __syncthreads();
    bool nt[4] = {false, false, false, false};
    for (int k = 0; k < 4 && !why; ++k) {
      if (off[k + 1] == off[k]) continue;
      const float v = s_min[k];
      if (dg && tid == 0) dg[2 + k] = v;
      if (v <= a.p.collision_dist) why = 3;
      else if (v > a.p.non_touch_dist && a.p.finger_status[k]) nt[k] = true;
    }
    // 4. one whole side not touching (:675-680)
    if (!why && ((nt[0] && nt[1]) || (nt[2] && nt[3]))) why = 4;
  }
  // 5./6. the object's own points against every finger mesh (:683-723)
  if (!why && a.nm > 0) {
    float mn[4] = {FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX};
    int in[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (a.finger_mesh[k].nf <= 0) continue;
      for (int i = tid; i < a.nm; i += SDF_THREADS) {
        const float s = sdf_point(a.finger_mesh[k], xform(s_M, f3(__ldg(a.model + i))));
        mn[k] = fminf(mn[k], s);
        in[k] += s < 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v = mn[k]; int c = in[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o)); c += __shfl_xor_sync(0xffffffffu, c, o); }
      if ((tid & 31) == 0) { if (v != FLT_MAX) atomic_min_float(&s_min[4 + k], v); if (c) atomicAdd(&s_inside[k], c); }
    }
    __syncthreads();
    for (int k = 0; k < 4 && !why; ++k) {
      if (a.finger_mesh[k].nf <= 0) continue;
      const float v = s_min[4 + k];
      if (dg && tid == 0) dg[6 + k] = v;
      if (v < a.p.collision_finger_dist) why = 5;
      else if ((float)(s_inside[k] / a.nm) > a.p.collision_finger_volume_ratio) why = 6;   // integer division, as in the reference
    }
  }
  if (tid == 0) { a.keep[h] = why == 0; if (a.reason) a.reason[h] = why; }
}

// ---- host: igl's normals, built once per mesh ----------------------------------------------------------------------------
struct V3 { float x, y, z; };
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float dotf(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// igl::doublearea for three corners: float edge lengths, Kahan's Heron formula in double
inline double double_area(V3 A, V3 B, V3 C) {
  double l[3] = {(double)std::sqrt(dotf(B - C, B - C)), (double)std::sqrt(dotf(C - A, C - A)), (double)std::sqrt(dotf(A - B, A - B))};
  std::sort(l, l + 3, [](double u, double v) { return u > v; });
  const double arg = (l[0] + (l[1] + l[2])) * (l[2] - (l[0] - l[1])) * (l[2] + (l[0] - l[1])) * (l[0] + (l[1] - l[2]));
  return 2.0 * 0.25 * std::sqrt(arg);
}

MeshDev mesh_dev(const hop_mesh *m) { return m ? MeshDev{m->d_hot, m->d_cold, m->nf} : MeshDev{nullptr, nullptr, 0}; }

}  // namespace

extern "C" int hop_mesh_upload(hop_ctx *ctx, const float *V, int nv, const int32_t *F, int nf, hop_mesh **out) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  if (!out || !V || !F || nv < 3 || nf < 1) { ctx->err = "hop_mesh_upload: bad arguments"; return HOP_EINVAL; }
  for (int k = 0; k < 3 * nf; ++k) if (F[k] < 0 || F[k] >= nv) { ctx->err = "hop_mesh_upload: face index out of range"; return HOP_EINVAL; }
  auto vtx = [&](int i) { return V3{V[3 * i], V[3 * i + 1], V[3 * i + 2]}; };
  // per_face_normals; per_vertex_normals with angle weights (internal_angles_using_squared_edge_lengths); per_edge_normals,
  // uniform and left un-normalised -- signed_distance.cpp:97-100
  std::vector<V3> FN(nf), VN(nv, V3{0, 0, 0});
  std::map<std::pair<int, int>, V3> EN;
  for (int f = 0; f < nf; ++f) {
    const int id[3] = {F[3 * f], F[3 * f + 1], F[3 * f + 2]};
    const V3 p[3] = {vtx(id[0]), vtx(id[1]), vtx(id[2])};
    const V3 v1 = p[1] - p[0], v2 = p[2] - p[0];
    V3 n = {v1.y * v2.z - v1.z * v2.y, v1.z * v2.x - v1.x * v2.z, v1.x * v2.y - v1.y * v2.x};
    const float r = std::sqrt(dotf(n, n));
    FN[f] = r == 0.f ? V3{0, 0, 0} : V3{n.x / r, n.y / r, n.z / r};
    float L[3];
    for (int d = 0; d < 3; ++d) { const V3 e = p[(d + 1) % 3] - p[(d + 2) % 3]; L[d] = dotf(e, e); }
    for (int d = 0; d < 3; ++d) {
      const float s1 = L[d], s2 = L[(d + 1) % 3], s3 = L[(d + 2) % 3];
      const float w = (float)std::acos((double)(s3 + s2 - s1) / (2. * std::sqrt((double)(s3 * s2))));
      V3 &vn = VN[id[d]];
      vn = {vn.x + w * FN[f].x, vn.y + w * FN[f].y, vn.z + w * FN[f].z};
      const int u = id[(d + 1) % 3], q = id[(d + 2) % 3];   // the edge opposite corner d
      V3 &e = EN[{std::min(u, q), std::max(u, q)}];
      e = {e.x + FN[f].x, e.y + FN[f].y, e.z + FN[f].z};
    }
  }
  for (auto &vn : VN) { const float r = std::sqrt(dotf(vn, vn)); if (r > 0.f) vn = {vn.x / r, vn.y / r, vn.z / r}; }
  std::vector<float4> hot((size_t)nf * HOT_F4), cold((size_t)nf * COLD_F4);
  auto put = [](V3 v, float w) { return make_float4(v.x, v.y, v.z, w); };
  for (int f = 0; f < nf; ++f) {
    const int id[3] = {F[3 * f], F[3 * f + 1], F[3 * f + 2]};
    const V3 p[3] = {vtx(id[0]), vtx(id[1]), vtx(id[2])};
    float4 *hh = &hot[(size_t)f * HOT_F4], *cc = &cold[(size_t)f * COLD_F4];
    const V3 ctr = {(p[0].x + p[1].x + p[2].x) / 3.f, (p[0].y + p[1].y + p[2].y) / 3.f, (p[0].z + p[1].z + p[2].z) / 3.f};
    float rad = 0.f;
    for (int k = 0; k < 3; ++k) { hh[k] = put(p[k], 0.f); rad = std::max(rad, std::sqrt(dotf(p[k] - ctr, p[k] - ctr))); }
    hh[3] = put(ctr, rad * 1.00001f);
    cc[0] = put(FN[f], double_area(p[0], p[1], p[2]) > 1e-4 ? 1.f : 0.f);
    for (int k = 0; k < 3; ++k) {
      cc[1 + k] = put(VN[id[k]], 0.f);
      const int u = id[(k + 1) % 3], q = id[(k + 2) % 3];
      cc[4 + k] = put(EN[{std::min(u, q), std::max(u, q)}], 0.f);
    }
  }
  hop_mesh *m = new hop_mesh();
  m->nf = nf; m->nv = nv;
  if (cudaMalloc(&m->d_hot, sizeof(float4) * hot.size()) != cudaSuccess || cudaMalloc(&m->d_cold, sizeof(float4) * cold.size()) != cudaSuccess) {
    cudaFree(m->d_hot); delete m; ctx->err = "hop_mesh_upload: allocation failed"; return HOP_ENOMEM;
  }
  HOP_CUDA(ctx, cudaMemcpyAsync(m->d_hot, hot.data(), sizeof(float4) * hot.size(), cudaMemcpyHostToDevice, ctx->stream));
  HOP_CUDA(ctx, cudaMemcpyAsync(m->d_cold, cold.data(), sizeof(float4) * cold.size(), cudaMemcpyHostToDevice, ctx->stream));
  HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *out = m;
  return HOP_OK;
}

extern "C" int hop_mesh_free(hop_ctx *ctx, hop_mesh *mesh) {
  HOP_ENTER(ctx);
  if (!mesh) return HOP_OK;
  if (ctx) cudaStreamSynchronize(ctx->stream);
  cudaFree(mesh->d_hot); cudaFree(mesh->d_cold);
  delete mesh;
  return HOP_OK;
}

extern "C" int hop_sdf_query(hop_ctx *ctx, const hop_mesh *mesh, const float *pts, int n, const float *point_transforms, int H, float *S,
                             int32_t *I, float *min_out, float *max_out, int32_t *n_inside) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  if (!mesh || n < 0 || H < 0 || (n > 0 && !pts)) { ctx->err = "hop_sdf_query: bad arguments"; return HOP_EINVAL; }
  if (H == 0) return HOP_OK;
  if (n == 0) {   // Eigen's minCoeff of an empty vector is undefined in the reference; here: the identities of min / max
    for (int h = 0; h < H; ++h) { if (min_out) min_out[h] = FLT_MAX; if (max_out) max_out[h] = -FLT_MAX; if (n_inside) n_inside[h] = 0; }
    return HOP_OK;
  }
  cudaStream_t st = ctx->stream;
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t pb = up(sizeof(float4) * (size_t)n), xb = up(sizeof(float) * 12 * (size_t)H), sb = up(sizeof(float) * (size_t)H * n), hb = up(sizeof(float) * (size_t)H);
  char *d = (char *)ctx->ensure_io(pb + xb + 2 * sb + 3 * hb);
  if (!d) { ctx->err = "hop_sdf_query: staging allocation failed"; return HOP_ENOMEM; }
  float4 *d_pts = (float4 *)d;
  float *d_xf = (float *)(d + pb), *d_S = (float *)(d + pb + xb);
  int32_t *d_I = (int32_t *)(d + pb + xb + sb);
  float *d_mn = (float *)(d + pb + xb + 2 * sb), *d_mx = (float *)(d + pb + xb + 2 * sb + hb);
  int *d_in = (int *)(d + pb + xb + 2 * sb + 2 * hb);
  std::vector<float4> hp(n);
  for (int i = 0; i < n; ++i) hp[i] = make_float4(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], 0.f);
  HOP_CUDA(ctx, cudaMemcpyAsync(d_pts, hp.data(), sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, st));
  std::vector<float> hx;
  if (point_transforms) {
    hx.resize(12 * (size_t)H);
    for (int h = 0; h < H; ++h) for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) hx[12 * (size_t)h + 4 * r + c] = point_transforms[16 * (size_t)h + 4 * c + r];
    HOP_CUDA(ctx, cudaMemcpyAsync(d_xf, hx.data(), sizeof(float) * hx.size(), cudaMemcpyHostToDevice, st));
  }
  sdf_init_kernel<<<(H + 127) / 128, 128, 0, st>>>(d_mn, d_mx, d_in, H);
  ctx->launches += 1;
  SdfArgs a;
  a.mesh = mesh_dev(mesh); a.pts = d_pts; a.n = n; a.xf = point_transforms ? d_xf : nullptr;
  a.S = S ? d_S : nullptr; a.I = I ? d_I : nullptr; a.mn = d_mn; a.mx = d_mx; a.inside = d_in;
  for (int h0 = 0; h0 < H; h0 += 65535) {
    const int Hb = std::min(65535, H - h0);
    SdfArgs b = a;
    if (b.xf) b.xf += 12 * (size_t)h0;
    if (b.S) b.S += (size_t)h0 * n;
    if (b.I) b.I += (size_t)h0 * n;
    b.mn += h0; b.mx += h0; b.inside += h0;
    ProfScope ps(ctx, HOP_PROF_SDF);
    sdf_kernel<<<dim3((n + SDF_THREADS - 1) / SDF_THREADS, Hb), SDF_THREADS, 0, st>>>(b);
    ctx->launches += 1;
  }
  HOP_CUDA(ctx, cudaGetLastError());
  if (S) HOP_CUDA(ctx, cudaMemcpyAsync(S, d_S, sizeof(float) * (size_t)H * n, cudaMemcpyDeviceToHost, st));
  if (I) HOP_CUDA(ctx, cudaMemcpyAsync(I, d_I, sizeof(int32_t) * (size_t)H * n, cudaMemcpyDeviceToHost, st));
  if (min_out) HOP_CUDA(ctx, cudaMemcpyAsync(min_out, d_mn, sizeof(float) * (size_t)H, cudaMemcpyDeviceToHost, st));
  if (max_out) HOP_CUDA(ctx, cudaMemcpyAsync(max_out, d_mx, sizeof(float) * (size_t)H, cudaMemcpyDeviceToHost, st));
  if (n_inside) HOP_CUDA(ctx, cudaMemcpyAsync(n_inside, d_in, sizeof(int32_t) * (size_t)H, cudaMemcpyDeviceToHost, st));
  HOP_CUDA(ctx, cudaStreamSynchronize(st));
  return HOP_OK;
}

extern "C" int hop_reject_by_collision_dev(hop_ctx *ctx, const hop_mesh *object, const hop_mesh *const *finger_meshes, hop_cloud *const *finger_clouds,
                                           hop_cloud *scene_without_hand, hop_cloud *hand_cloud, hop_cloud *model, const float *d_poses, int H,
                                           const hop_collision_params *params, int32_t *d_keep, int32_t *d_reason, float *d_diag) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  if (!object || !params || H < 0 || (H > 0 && (!d_poses || !d_keep))) { ctx->err = "hop_reject_by_collision: bad arguments"; return HOP_EINVAL; }
  if (H == 0) return HOP_OK;
  CollisionArgs a;
  a.object = mesh_dev(object);
  for (int k = 0; k < 4; ++k) {
    a.finger_mesh[k] = mesh_dev(finger_meshes ? finger_meshes[k] : nullptr);
    hop_cloud *c = finger_clouds ? finger_clouds[k] : nullptr;
    a.finger_pts[k] = nullptr; a.finger_n[k] = 0;
    if (c && c->n > 0) {   // the min over a cloud does not depend on the order: visit it along the Morton curve
      const int rc = hop_cloud_query_order(ctx, c);
      if (rc != HOP_OK) return rc;
      a.finger_pts[k] = c->d_pw_q; a.finger_n[k] = c->n;
    }
  }
  a.scene = scene_without_hand ? scene_without_hand->d_pw : nullptr; a.ns = scene_without_hand ? scene_without_hand->n : 0;
  a.hand = hand_cloud ? hand_cloud->d_pw : nullptr; a.nh = hand_cloud ? hand_cloud->n : 0;
  a.model = nullptr; a.nm = 0;
  if (model && model->n > 0) {
    const int rc = hop_cloud_query_order(ctx, model);
    if (rc != HOP_OK) return rc;
    a.model = model->d_pw_q; a.nm = model->n;
  }
  a.poses = d_poses; a.p = *params; a.keep = d_keep; a.reason = d_reason; a.diag = d_diag;
  {
    ProfScope ps(ctx, HOP_PROF_SDF);
    collision_kernel<<<H, SDF_THREADS, 0, ctx->stream>>>(a);
    ctx->launches += 1;
  }
  HOP_CUDA(ctx, cudaGetLastError());
  return HOP_OK;
}

extern "C" int hop_reject_by_collision(hop_ctx *ctx, const hop_mesh *object, const hop_mesh *const *finger_meshes, hop_cloud *const *finger_clouds,
                                       hop_cloud *scene_without_hand, hop_cloud *hand_cloud, hop_cloud *model, const float *poses, int H,
                                       const hop_collision_params *params, int32_t *keep, int32_t *reason, float *diag) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  if (H < 0 || (H > 0 && (!poses || !keep))) { ctx->err = "hop_reject_by_collision: bad arguments"; return HOP_EINVAL; }
  if (H == 0) return HOP_OK;
  cudaStream_t st = ctx->stream;
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t pb = up(sizeof(float) * 16 * (size_t)H), kb = up(sizeof(int32_t) * (size_t)H), db = up(sizeof(float) * 10 * (size_t)H);
  char *d = (char *)ctx->ensure_io(pb + 2 * kb + db);
  if (!d) { ctx->err = "hop_reject_by_collision: staging allocation failed"; return HOP_ENOMEM; }
  float *d_poses = (float *)d;
  int32_t *d_keep = (int32_t *)(d + pb), *d_reason = (int32_t *)(d + pb + kb);
  float *d_diag = (float *)(d + pb + 2 * kb);
  HOP_CUDA(ctx, cudaMemcpyAsync(d_poses, poses, sizeof(float) * 16 * (size_t)H, cudaMemcpyHostToDevice, st));
  const int rc = hop_reject_by_collision_dev(ctx, object, finger_meshes, finger_clouds, scene_without_hand, hand_cloud, model, d_poses, H, params,
                                             d_keep, reason ? d_reason : nullptr, diag ? d_diag : nullptr);
  if (rc != HOP_OK) return rc;
  HOP_CUDA(ctx, cudaMemcpyAsync(keep, d_keep, sizeof(int32_t) * (size_t)H, cudaMemcpyDeviceToHost, st));
  if (reason) HOP_CUDA(ctx, cudaMemcpyAsync(reason, d_reason, sizeof(int32_t) * (size_t)H, cudaMemcpyDeviceToHost, st));
  if (diag) HOP_CUDA(ctx, cudaMemcpyAsync(diag, d_diag, sizeof(float) * 10 * (size_t)H, cudaMemcpyDeviceToHost, st));
  HOP_CUDA(ctx, cudaStreamSynchronize(st));
  return HOP_OK;
}
