// k_s4pcs.cu -- K2a / K2b: pair extraction and congruent-set search of Super4PCS for ALL trials of a frame at once, then
// K3 (k_verify.cu) on the result: the device part of PoseEstimator::runSuper4pcs.
//
//   K2a  FunctorSuper4PCS::ExtractPairs            src/OpenGR_4pcs/src/gr/algorithms/FunctorSuper4pcs.h:79-116
//        (IntersectionFunctor accelerators/pairExtraction/intersectionFunctor.h:105-240 is a conservative pre-filter: what
//         survives is exactly PairCreationFunctor::process pairCreationFunctor.h:189-214 = the distance gate
//         + pairPPFisGood / AdaptivePointFilter PointPairFilter.h:17-38,88-172)
//   K2b  FunctorSuper4PCS::FindCongruentQuadrilaterals   FunctorSuper4pcs.h:131-293
//        (IndexedNormalSet accelerators/normalset.hpp:113-259: position cell within one ring AND direction bin on the
//         rasterised cone AND the invariant points within sqrt(delta))
//
// The reference walks an octree and allocates a 343-bin angular grid per cell for every trial, serially.  Here both
// steps are plain data-parallel tests: K2a = one thread per unordered pair of Q (all trials x both base segments in one
// launch), K2b = one warp per first-set pair scanning the second set; matches are counted, offsets come from a prefix
// sum, and a second pass writes the quadrilaterals -- so the output order is deterministic: (trial, first-pair, second-
// pair) with pairs in (i, j) order.  (The reference's order inside a trial follows its octree traversal; the SET of pairs
// and of quadrilaterals per trial is identical, tests pin both against the compiled reference.)
// Everything that needs libm (acos/sin/cos of per-trial quantities) is computed on the host like the reference does; the
// device uses IEEE-exact +,-,*,/,sqrt in Eigen's evaluation order, so cells and bins are bit-identical.  The one
// per-pair transcendental (acos in pairPPFisGood) is evaluated in double and rounded to float.
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include <algorithm>
#include <memory>
#include <cmath>

#include "hop_common.cuh"
#include "s4pcs.h"

namespace {

constexpr int NG = 7;                 // direction grid cells per dimension (IndexedNormalSet<Point,3,7,Scalar>)
constexpr int MASK_WORDS = 11;        // 343 bins
constexpr int MAX_RING = 64;          // ring samples per query (the reference's formula gives at most 56)

struct ExtractParams {                // one per (trial, base segment)
  float pair_distance, eps;           // |b0 - b1|, delta
  float b_n0, b_n1, n0_n1;            // the base pair's three PPF angles in degrees (pairPPFisGood)
  float pna, norm_threshold;          // pair_normals_angle; 0.5 * max_normal_difference in radians, < 0 = off
  int active;
};

struct TrialParams {
  float inv1, inv2;
  int ring_n;                         // nbSample
  int active;
  float ring[MAX_RING][2];            // (sin(alpha) cos(theta_a), sin(alpha) sin(theta_a)) from the host's libm
  float cos_alpha;
};

struct GridParams {
  float inv_eps_div;                  // _epsilon (positions are divided by it)
  int eg_size;
  float nepsilon;                     // 1/7 + 1e-5
  float delta;                        // compared against a SQUARED norm, as the reference does (FunctorSuper4pcs.h:277)
};

struct __align__(16) PairRec1 {      // first-set pair: what addElement stores + its invariant point on Q
  int a, b;                           // Q indices
  int cell;                           // packed (cx, cy, cz), 10 bits each
  int nid;                            // direction bin
  float ix, iy, iz; int pad;
};
struct __align__(16) PairRec2 {      // second-set pair: the query
  int a, b;
  int cell;
  int pad;
  float qx, qy, qz; int pad2;
  uint32_t mask[MASK_WORDS]; uint32_t pad3;
};

__device__ __forceinline__ float sum3(float t0, float t1, float t2) { return __fadd_rn(t0, __fadd_rn(t1, t2)); }
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
  return sum3(__fmul_rn(ax, bx), __fmul_rn(ay, by), __fmul_rn(az, bz));
}
__device__ __forceinline__ void normalize3(float &x, float &y, float &z) {  // Eigen normalized(): v / sqrt(v.v) when v.v > 0
  const float s = dot3(x, y, z, x, y, z);
  if (s > 0.f) { const float n = __fsqrt_rn(s); x = __fdiv_rn(x, n); y = __fdiv_rn(y, n); z = __fdiv_rn(z, n); }
}
__device__ __forceinline__ float acosf_cr(float x) { return __double2float_rn(acos((double)x)); }
__device__ __forceinline__ float to_deg(float rad) { return __double2float_rn(__dmul_rn(__ddiv_rn((double)rad, M_PI), 180.0)); }

__device__ __forceinline__ void unrank_pair(long long k, int &i, int &j) {  // k = i (i - 1) / 2 + j, i > j >= 0
  long long r = (long long)((1.0 + sqrt(1.0 + 8.0 * (double)k)) * 0.5);
  while (r * (r - 1) / 2 > k) --r;
  while ((r + 1) * r / 2 <= k) ++r;
  i = (int)r; j = (int)(k - r * (r - 1) / 2);
}

// K2a: flags[e][k] = 1 when the unordered pair k of Q passes PairCreationFunctor::process for extraction e
__global__ void __launch_bounds__(256) extract_pairs_kernel(const float4 *__restrict__ Qp, const float4 *__restrict__ Qn, long long NP,
                                                            const ExtractParams *__restrict__ ex, unsigned char *__restrict__ flags,
                                                            int *__restrict__ counts) {
  const int e = blockIdx.y;
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= NP) return;
  const ExtractParams x = ex[e];
  bool keep = false;
  if (x.active) {
    int i, j;
    unrank_pair(k, i, j);
    const float4 p = __ldg(&Qp[j]), q = __ldg(&Qp[i]);           // process(i, j), i > j: p = Q[j], q = Q[i]
    float dx = __fsub_rn(q.x, p.x), dy = __fsub_rn(q.y, p.y), dz = __fsub_rn(q.z, p.z);
    const float dist = __fsqrt_rn(dot3(dx, dy, dz, dx, dy, dz));
    keep = !(fabsf(__fsub_rn(dist, x.pair_distance)) > x.eps);
    // pairPPFisGood(p, q, b0, b1)
    if (keep && (double)fabsf(__fsub_rn(dist, x.pair_distance)) > 5e-3) keep = false;  // length2 == pair_distance (same expression)
    if (keep) {
      const float4 np = __ldg(&Qn[j]), nq = __ldg(&Qn[i]);
      normalize3(dx, dy, dz);                                     // pq
      const float pq_np = to_deg(acosf_cr(fabsf(dot3(dx, dy, dz, np.x, np.y, np.z))));
      const float pq_nq = to_deg(acosf_cr(fabsf(dot3(dx, dy, dz, nq.x, nq.y, nq.z))));
      const float np_nq = to_deg(acosf_cr(dot3(np.x, np.y, np.z, nq.x, nq.y, nq.z)));
      if (fabsf(__fsub_rn(pq_np, x.b_n0)) > 30.f || fabsf(__fsub_rn(pq_nq, x.b_n1)) > 30.f || fabsf(__fsub_rn(np_nq, x.n0_n1)) > 30.f) keep = false;
      // AdaptivePointFilter's normal-difference gate (off in the reference's configuration: max_normal_difference = -1)
      if (keep && x.norm_threshold > 0.f && dot3(nq.x, nq.y, nq.z, nq.x, nq.y, nq.z) > 0.f && dot3(np.x, np.y, np.z, np.x, np.y, np.z) > 0.f) {
        const float mx = __fsub_rn(nq.x, np.x), my = __fsub_rn(nq.y, np.y), mz = __fsub_rn(nq.z, np.z);
        const float sx = __fadd_rn(nq.x, np.x), sy = __fadd_rn(nq.y, np.y), sz = __fadd_rn(nq.z, np.z);
        const double first = (double)__fsqrt_rn(dot3(mx, my, mz, mx, my, mz)), second = (double)__fsqrt_rn(dot3(sx, sy, sz, sx, sy, sz));
        const float nd = __double2float_rn(fmin(fabs(__dsub_rn(first, (double)x.pna)), fabs(__dsub_rn(second, (double)x.pna))));
        if (nd > x.norm_threshold) keep = false;
      }
    }
  }
  flags[(size_t)e * NP + k] = keep ? 1 : 0;
  // per-extraction totals (order comes from the flag compaction, not from this atomic)
  const unsigned m = __ballot_sync(__activemask(), keep);
  if (keep && (threadIdx.x & 31) == (__ffs(m) - 1)) atomicAdd(&counts[e], __popc(m));
}

__device__ __forceinline__ int pack_cell(int cx, int cy, int cz) { return (cx & 1023) | ((cy & 1023) << 10) | ((cz & 1023) << 20); }
__device__ __forceinline__ bool cells_adjacent(int a, int b, int eg) {
  // `a` is a stored pair's cell, `b` the query's: the pair was inserted into the in-bounds cells of its one-ring
  const int ax = a & 1023, ay = (a >> 10) & 1023, az = (a >> 20) & 1023, bx = b & 1023, by = (b >> 10) & 1023, bz = (b >> 20) & 1023;
  return abs(ax - bx) <= 1 && abs(ay - by) <= 1 && abs(az - bz) <= 1 && bx < eg && by < eg && bz < eg;
}

// K2b prep: one thread per selected unordered pair; writes its two ordered pairs (i, j), (j, i) as PairRec1 or PairRec2
__global__ void __launch_bounds__(128) prepare_pairs_kernel(const long long *__restrict__ sel, int n_sel, long long NP, const int *__restrict__ ex_begin,
                                                            const float4 *__restrict__ Qp, const float4 *__restrict__ Qu,
                                                            const TrialParams *__restrict__ tp, GridParams g, PairRec1 *__restrict__ r1,
                                                            PairRec2 *__restrict__ r2, const int *__restrict__ rec_base) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_sel) return;
  const long long flat = sel[s];
  const int e = (int)(flat / NP);
  int i, j;
  unrank_pair(flat - (long long)e * NP, i, j);
  const int t = e >> 1, second = e & 1;
  const TrialParams *T = tp + t;
  const int local = s - ex_begin[e];                 // rank of this unordered pair inside its extraction
  for (int o = 0; o < 2; ++o) {                      // pairs.emplace_back(i, j); pairs.emplace_back(j, i)
    const int a = o == 0 ? i : j, b = o == 0 ? j : i;
    const float4 u1 = __ldg(&Qu[a]), u2 = __ldg(&Qu[b]);
    const float4 w1 = __ldg(&Qp[a]), w2 = __ldg(&Qp[b]);
    float nx = __fsub_rn(u2.x, u1.x), ny = __fsub_rn(u2.y, u1.y), nz = __fsub_rn(u2.z, u1.z);
    const float inv = second ? T->inv2 : T->inv1;
    // invariant point in the unit cube and its cell
    const float ex_ = __fadd_rn(u1.x, __fmul_rn(inv, nx)), ey = __fadd_rn(u1.y, __fmul_rn(inv, ny)), ez = __fadd_rn(u1.z, __fmul_rn(inv, nz));
    const int cell = pack_cell(__float2int_rz(__fdiv_rn(ex_, g.inv_eps_div)), __float2int_rz(__fdiv_rn(ey, g.inv_eps_div)),
                               __float2int_rz(__fdiv_rn(ez, g.inv_eps_div)));
    normalize3(nx, ny, nz);
    const int idx = rec_base[e] + 2 * local + o;
    if (!second) {
      // invPoint = pp1 + (pp2 - pp1) * invariant1 on Q
      PairRec1 r;
      r.a = a; r.b = b; r.cell = cell; r.pad = 0;
      const int bx = __float2int_rz(__fdiv_rn(__fadd_rn(__fdiv_rn(nx, 2.f), 0.5f), g.nepsilon));
      const int by = __float2int_rz(__fdiv_rn(__fadd_rn(__fdiv_rn(ny, 2.f), 0.5f), g.nepsilon));
      const int bz = __float2int_rz(__fdiv_rn(__fadd_rn(__fdiv_rn(nz, 2.f), 0.5f), g.nepsilon));
      r.nid = bx + NG * by + NG * NG * bz;
      r.ix = __fadd_rn(w1.x, __fmul_rn(__fsub_rn(w2.x, w1.x), inv));
      r.iy = __fadd_rn(w1.y, __fmul_rn(__fsub_rn(w2.y, w1.y), inv));
      r.iz = __fadd_rn(w1.z, __fmul_rn(__fsub_rn(w2.z, w1.z), inv));
      r1[idx] = r;
    } else {
      PairRec2 r;
      r.a = a; r.b = b; r.cell = cell; r.pad = r.pad2 = 0; r.pad3 = 0;
      r.qx = __fadd_rn(w1.x, __fmul_rn(inv, __fsub_rn(w2.x, w1.x)));
      r.qy = __fadd_rn(w1.y, __fmul_rn(inv, __fsub_rn(w2.y, w1.y)));
      r.qz = __fadd_rn(w1.z, __fmul_rn(inv, __fsub_rn(w2.z, w1.z)));
#pragma unroll
      for (int w = 0; w < MASK_WORDS; ++w) r.mask[w] = 0u;
      // getNeighbors(query, queryn, alpha): the direction bins met by the cone of half-angle alpha around queryn
      // q.setFromTwoVectors((0,0,1), n)
      float v1x = nx, v1y = ny, v1z = nz;
      normalize3(v1x, v1y, v1z);
      const float c = sum3(__fmul_rn(v1x, 0.f), __fmul_rn(v1y, 0.f), __fmul_rn(v1z, 1.f));
      float qx, qy, qz, qw;
      if (c < __fadd_rn(-1.f, 1e-5f)) {
        // nearly opposite to +z: Eigen takes the axis from an SVD; any unit axis orthogonal to z gives the same cone,
        // sampled at another phase (a measure-zero configuration: queryn within 0.26 degrees of -z)
        const float cc = fmaxf(c, -1.f), w2 = __fmul_rn(__fadd_rn(1.f, cc), 0.5f);
        qw = __fsqrt_rn(w2); const float sv = __fsqrt_rn(__fsub_rn(1.f, w2)); qx = sv; qy = 0.f; qz = 0.f;
      } else {
        const float ax = __fsub_rn(__fmul_rn(0.f, v1z), __fmul_rn(1.f, v1y)), ay = __fsub_rn(__fmul_rn(1.f, v1x), __fmul_rn(0.f, v1z)),
                    az = __fsub_rn(__fmul_rn(0.f, v1y), __fmul_rn(0.f, v1x));
        const float sq = __fsqrt_rn(__fmul_rn(__fadd_rn(1.f, c), 2.f)), invs = __fdiv_rn(1.f, sq);
        qx = __fmul_rn(ax, invs); qy = __fmul_rn(ay, invs); qz = __fmul_rn(az, invs); qw = __fmul_rn(sq, 0.5f);
      }
      for (int s2 = 0; s2 < T->ring_n; ++s2) {
        const float vx = T->ring[s2][0], vy = T->ring[s2][1], vz = T->cos_alpha;
        // Quaternion::_transformVector: uv = vec x v; uv += uv; v + w * uv + vec x uv
        float ux = __fsub_rn(__fmul_rn(qy, vz), __fmul_rn(qz, vy)), uy = __fsub_rn(__fmul_rn(qz, vx), __fmul_rn(qx, vz)),
              uz = __fsub_rn(__fmul_rn(qx, vy), __fmul_rn(qy, vx));
        ux = __fadd_rn(ux, ux); uy = __fadd_rn(uy, uy); uz = __fadd_rn(uz, uz);
        float dx = __fadd_rn(__fadd_rn(vx, __fmul_rn(qw, ux)), __fsub_rn(__fmul_rn(qy, uz), __fmul_rn(qz, uy)));
        float dy = __fadd_rn(__fadd_rn(vy, __fmul_rn(qw, uy)), __fsub_rn(__fmul_rn(qz, ux), __fmul_rn(qx, uz)));
        float dz = __fadd_rn(__fadd_rn(vz, __fmul_rn(qw, uz)), __fsub_rn(__fmul_rn(qx, uy), __fmul_rn(qy, ux)));
        normalize3(dx, dy, dz);
        const int bx = __float2int_rz(__fdiv_rn(__fadd_rn(__fdiv_rn(dx, 2.f), 0.5f), g.nepsilon));
        const int by = __float2int_rz(__fdiv_rn(__fadd_rn(__fdiv_rn(dy, 2.f), 0.5f), g.nepsilon));
        const int bz = __float2int_rz(__fdiv_rn(__fadd_rn(__fdiv_rn(dz, 2.f), 0.5f), g.nepsilon));
        const int id = bx + NG * by + NG * NG * bz;
        if (id >= 0 && id < NG * NG * NG) r.mask[id >> 5] |= 1u << (id & 31);
      }
      r2[idx] = r;
    }
  }
}

struct JoinArgs {
  const PairRec1 *r1; const PairRec2 *r2;
  const int *r1_trial;      // trial of every first-set record
  const int *r2_begin;      // per trial: range of its second-set records
  const int *r2_end;
  int n1;
  int eg_size; float delta;
  int *counts;              // FILL == false: matches per first-set record
  const int *offsets;       // FILL == true: exclusive prefix sum of counts
  int4 *quads; int *quad_trial;
};

// K2b: one warp per first-set pair `id`, scanning its trial's second-set pairs `i` in order
template <bool FILL>
__global__ void __launch_bounds__(256) congruent_join_kernel(JoinArgs a) {
  const int id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (id >= a.n1) return;
  const PairRec1 p = a.r1[id];
  const int t = a.r1_trial[id];
  const int b = a.r2_begin[t], e = a.r2_end[t];
  int total = 0;
  const int base = FILL ? a.offsets[id] : 0;
  for (int i0 = b; i0 < e; i0 += 32) {
    const int i = i0 + lane;
    bool hit = false;
    int qa = 0, qb = 0;
    if (i < e) {
      const PairRec2 *q = a.r2 + i;
      const int cell = __ldg(&q->cell);
      if (cells_adjacent(p.cell, cell, a.eg_size) && p.nid >= 0 && p.nid < NG * NG * NG && ((__ldg(&q->mask[p.nid >> 5]) >> (p.nid & 31)) & 1u)) {
        const float dx = __fsub_rn(__ldg(&q->qx), p.ix), dy = __fsub_rn(__ldg(&q->qy), p.iy), dz = __fsub_rn(__ldg(&q->qz), p.iz);
        hit = dot3(dx, dy, dz, dx, dy, dz) <= a.delta;   // (queryQ - invPoint).squaredNorm() <= distance_threshold2
        if (FILL && hit) { qa = __ldg(&q->a); qb = __ldg(&q->b); }
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (FILL && hit) {
      const int at = base + total + __popc(m & ((1u << lane) - 1u));
      a.quads[at] = make_int4(p.a, p.b, qa, qb);
      a.quad_trial[at] = t;
    }
    total += __popc(m);
  }
  if (!FILL && lane == 0) a.counts[id] = total;
}

__global__ void fill_int_ranges_kernel(int *out, const int *begin, const int *end, int n_ranges) {
  // out[k] = r for every k in [begin[r], end[r])
  const int r = blockIdx.y;
  if (r >= n_ranges) return;
  for (int k = begin[r] + blockIdx.x * blockDim.x + threadIdx.x; k < end[r]; k += gridDim.x * blockDim.x) out[k] = r;
}

// stream-ordered scratch of one call: cudaMallocAsync / cudaFreeAsync on the context's stream, served from the device's
// default memory pool (hop_create raises its release threshold, so after the first frame no call reaches the driver's allocator)
struct DevBuf {
  void *p = nullptr;
  cudaStream_t st;
  explicit DevBuf(cudaStream_t s) : st(s) {}
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { if (p) cudaFreeAsync(p, st); }
  template <typename T> T *as() { return (T *)p; }
  cudaError_t alloc(size_t bytes) {
    if (p) cudaFreeAsync(p, st);
    p = nullptr;
    return cudaMallocAsync(&p, std::max<size_t>(bytes, 16), st);
  }
};

}  // namespace

extern "C" int hop_super4pcs_run(hop_ctx *ctx, hop_s4pcs_plan *plan, float *hyp_poses, float *hyp_lcp, int capacity, int32_t *n_hyp) {
  HOP_ENTER(ctx);
  if (!ctx || !plan || capacity < 0 || (capacity > 0 && (!hyp_poses || !hyp_lcp))) { if (ctx) ctx->err = "hop_super4pcs_run: bad arguments"; return HOP_EINVAL; }
  if (n_hyp) *n_hyp = 0;
  plan->trial_ranges.clear(); plan->pairs.clear(); plan->quads.clear(); plan->trials_executed = 0;
  const int T = (int)plan->trials.size(), nQ = (int)plan->Q.size(), nP = (int)plan->P.size();
  if (T == 0 || nQ < 2 || nP < 1) return HOP_OK;
  bool any = false;
  for (const S4Trial &t : plan->trials) any = any || t.base_ok;
  if (!any) return HOP_OK;
  cudaStream_t st = ctx->stream;
  const float delta = plan->opt.delta;

  // ---- per-trial / per-extraction parameters: everything that needs libm is evaluated here, on the host ----
  std::vector<ExtractParams> ex(2 * T);
  std::vector<TrialParams> tp(T);
  auto dot = [](const float *a, const float *b) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); };
  for (int t = 0; t < T; ++t) {
    const S4Trial &tr = plan->trials[t];
    std::memset(&tp[t], 0, sizeof(TrialParams));
    for (int s = 0; s < 2; ++s) {
      ExtractParams &x = ex[2 * t + s];
      std::memset(&x, 0, sizeof(x));
      x.active = tr.base_ok;
      if (!tr.base_ok) continue;
      const S4Pt &b0 = tr.b[2 * s], &b1 = tr.b[2 * s + 1];
      x.pair_distance = s == 0 ? tr.dist1 : tr.dist2;
      x.eps = delta;  // distance_factor (1.0) * delta
      float d[3] = {b1.p[0] - b0.p[0], b1.p[1] - b0.p[1], b1.p[2] - b0.p[2]};
      const float z = dot(d, d);
      if (z > 0.f) { const float n = std::sqrt(z); d[0] /= n; d[1] /= n; d[2] /= n; }
      x.b_n0 = std::acos(std::abs(dot(d, b0.n))) / M_PI * 180;   // pairPPFisGood (PointPairFilter.h:27-32)
      x.b_n1 = std::acos(std::abs(dot(d, b1.n))) / M_PI * 180;
      x.n0_n1 = std::acos(dot(b0.n, b1.n)) / M_PI * 180;
      x.pna = s == 0 ? tr.nangle1 : tr.nangle2;
      x.norm_threshold = plan->opt.max_normal_difference > 0.f ? (float)(0.5 * plan->opt.max_normal_difference * M_PI / 180.0) : -1.f;
    }
    TrialParams &p = tp[t];
    p.active = tr.base_ok;
    if (!tr.base_ok) continue;
    p.inv1 = tr.inv1; p.inv2 = tr.inv2; p.cos_alpha = tr.alpha;
    // IndexedNormalSet::getNeighbors (normalset.hpp:207-216)
    const float alpha = std::acos(tr.alpha);
    const float perimeter = 2.f * M_PI * std::atan(alpha);
    const unsigned int nb = 2 * std::ceil(perimeter * float(NG) / 2.f);
    const float angle_step = 2.f * M_PI / float(nb);
    const float sin_alpha = std::sin(alpha);
    p.ring_n = (int)std::min<unsigned int>(nb, MAX_RING);
    for (int a2 = 0; a2 < p.ring_n; ++a2) {
      const float theta = float(a2) * angle_step;
      p.ring[a2][0] = sin_alpha * std::cos(theta);
      p.ring[a2][1] = sin_alpha * std::sin(theta);
    }
  }
  GridParams g;
  {
    // IndexedNormalSet(eps) with eps = getNormalizedEpsilon(delta) = delta / ratio (normalset.h:116-126)
    const float eps = delta / plan->ratio;
    const int depth = -std::log2(eps);
    g.eg_size = std::pow(2, depth);
    if (g.eg_size < 1) g.eg_size = 1;
    if (g.eg_size > 1023) { ctx->err = "hop_super4pcs_run: delta too small for the 10-bit cell index"; return HOP_EINVAL; }
    g.inv_eps_div = 1.f / g.eg_size;
    g.nepsilon = float(1.) / float(NG) + 0.00001;
    g.delta = delta;
  }

  std::unique_ptr<HopTraceScope> trace(new HopTraceScope(ctx, "  s4run: uploads + K2a + select + sync"));
  // ---- the scene of the verification step (K3) goes up first and its NN grid is built on the context's second stream while K2a / K2b run.
  //      The centred scene lives in a cloud the context keeps from frame to frame: its buffers and the buffers of its NN grid are reused (a
  //      fresh cloud per call costs a dozen cudaMalloc / cudaFree round trips, 1.5 ms of a 2 ms call) ----
  {
    std::vector<float> Pxyz(3 * (size_t)nP), Pn(3 * (size_t)nP);
    for (int i = 0; i < nP; ++i) for (int k = 0; k < 3; ++k) { Pxyz[3 * i + k] = plan->P[i].p[k]; Pn[3 * i + k] = plan->P[i].n[k]; }
    int rc0 = ctx->s4_scene ? hop_cloud_update(ctx, ctx->s4_scene, Pxyz.data(), Pn.data(), nullptr, nP)
                            : hop_cloud_upload(ctx, Pxyz.data(), Pn.data(), nullptr, nP, &ctx->s4_scene);
    if (rc0 != HOP_OK) return rc0;
    rc0 = hop_cloud_prepare_nn_async(ctx, ctx->s4_scene, delta, 0.f);
    if (rc0 != HOP_OK) return rc0;
  }
  // ---- uploads ----
  const long long NP = (long long)nQ * (nQ - 1) / 2;
  DevBuf bQp(st), bQn(st), bQu(st), bEx(st), bTp(st), bFlags(st), bCnt(st), bSel(st), bCub(st);
  std::vector<float4> hQp(nQ), hQn(nQ), hQu(nQ);
  for (int i = 0; i < nQ; ++i) {
    hQp[i] = make_float4(plan->Q[i].p[0], plan->Q[i].p[1], plan->Q[i].p[2], 0.f);
    hQn[i] = make_float4(plan->Q[i].n[0], plan->Q[i].n[1], plan->Q[i].n[2], 0.f);
    hQu[i] = make_float4(plan->Qunit[3 * i], plan->Qunit[3 * i + 1], plan->Qunit[3 * i + 2], 0.f);
  }
  HOP_CUDA(ctx, bQp.alloc(sizeof(float4) * nQ)); HOP_CUDA(ctx, bQn.alloc(sizeof(float4) * nQ)); HOP_CUDA(ctx, bQu.alloc(sizeof(float4) * nQ));
  HOP_CUDA(ctx, bEx.alloc(sizeof(ExtractParams) * 2 * T)); HOP_CUDA(ctx, bTp.alloc(sizeof(TrialParams) * T));
  HOP_CUDA(ctx, bFlags.alloc((size_t)NP * 2 * T)); HOP_CUDA(ctx, bCnt.alloc(sizeof(int) * (2 * T + 4)));
  HOP_CUDA(ctx, cudaMemcpyAsync(bQp.p, hQp.data(), sizeof(float4) * nQ, cudaMemcpyHostToDevice, st));
  HOP_CUDA(ctx, cudaMemcpyAsync(bQn.p, hQn.data(), sizeof(float4) * nQ, cudaMemcpyHostToDevice, st));
  HOP_CUDA(ctx, cudaMemcpyAsync(bQu.p, hQu.data(), sizeof(float4) * nQ, cudaMemcpyHostToDevice, st));
  HOP_CUDA(ctx, cudaMemcpyAsync(bEx.p, ex.data(), sizeof(ExtractParams) * 2 * T, cudaMemcpyHostToDevice, st));
  HOP_CUDA(ctx, cudaMemcpyAsync(bTp.p, tp.data(), sizeof(TrialParams) * T, cudaMemcpyHostToDevice, st));
  HOP_CUDA(ctx, cudaMemsetAsync(bCnt.p, 0, sizeof(int) * (2 * T + 4), st));

  // ---- K2a ----
  std::unique_ptr<ProfScope> prof(new ProfScope(ctx, HOP_PROF_S4_PAIRS));
  extract_pairs_kernel<<<dim3((unsigned)((NP + 255) / 256), 2 * T), 256, 0, st>>>(bQp.as<float4>(), bQn.as<float4>(), NP, bEx.as<ExtractParams>(),
                                                                                  bFlags.as<unsigned char>(), bCnt.as<int>());
  ctx->launches += 1;
  // ordered list of the selected (extraction, pair) indices = ascending flat index
  const long long n_flat = NP * 2 * T;
  if (n_flat > 0x7fffffffll) { ctx->err = "hop_super4pcs_run: too many candidate pairs for one launch"; return HOP_EINVAL; }
  HOP_CUDA(ctx, bSel.alloc(sizeof(long long) * (size_t)n_flat));
  size_t cub_bytes = 0;
  cub::CountingInputIterator<long long> counting(0);
  int *d_nsel = bCnt.as<int>() + 2 * T;
  cub::DeviceSelect::Flagged(nullptr, cub_bytes, counting, bFlags.as<unsigned char>(), bSel.as<long long>(), d_nsel, (int)n_flat, st);
  HOP_CUDA(ctx, bCub.alloc(cub_bytes));
  cub::DeviceSelect::Flagged(bCub.p, cub_bytes, counting, bFlags.as<unsigned char>(), bSel.as<long long>(), d_nsel, (int)n_flat, st);
  ctx->launches += 1;
  prof.reset();
  std::vector<int> cnt(2 * T + 4);
  HOP_CUDA(ctx, cudaMemcpyAsync(cnt.data(), bCnt.p, sizeof(int) * (2 * T + 4), cudaMemcpyDeviceToHost, st));
  HOP_CUDA(ctx, cudaStreamSynchronize(st));
  const int n_sel = cnt[2 * T];
  trace.reset(new HopTraceScope(ctx, "  s4run: layout + K2b count + scan + sync"));

  // layout of the ordered-pair records: first-set records of all trials, then second-set records of all trials
  std::vector<int> ex_begin(2 * T), rec_base(2 * T), r2_begin(T), r2_end(T), r1_begin(T), r1_end(T);
  int acc = 0, n1 = 0, n2 = 0;
  for (int e = 0; e < 2 * T; ++e) { ex_begin[e] = acc; acc += cnt[e]; }
  for (int t = 0; t < T; ++t) { r1_begin[t] = n1; rec_base[2 * t] = n1; n1 += 2 * cnt[2 * t]; r1_end[t] = n1; }
  for (int t = 0; t < T; ++t) { r2_begin[t] = n2; rec_base[2 * t + 1] = n2; n2 += 2 * cnt[2 * t + 1]; r2_end[t] = n2; }
  if (acc != n_sel) { ctx->err = "hop_super4pcs_run: pair count mismatch"; return HOP_ECUDA; }
  const bool keep = plan->opt.keep_intermediates != 0;
  if (keep) plan->trial_ranges.assign(6 * T, 0);

  int M = 0, T_exec = 0;
  std::vector<int> quad_begin(T + 1, 0);
  DevBuf bR1(st), bR2(st), bMeta(st), bR1Trial(st), bCounts(st), bOffsets(st), bQuads(st), bQuadTrial(st);
  if (n1 > 0 && n2 > 0) {
    HOP_CUDA(ctx, bR1.alloc(sizeof(PairRec1) * (size_t)n1)); HOP_CUDA(ctx, bR2.alloc(sizeof(PairRec2) * (size_t)n2));
    // meta: ex_begin | rec_base | r1_begin | r1_end | r2_begin | r2_end
    std::vector<int> meta;
    meta.insert(meta.end(), ex_begin.begin(), ex_begin.end()); meta.insert(meta.end(), rec_base.begin(), rec_base.end());
    meta.insert(meta.end(), r1_begin.begin(), r1_begin.end()); meta.insert(meta.end(), r1_end.begin(), r1_end.end());
    meta.insert(meta.end(), r2_begin.begin(), r2_begin.end()); meta.insert(meta.end(), r2_end.begin(), r2_end.end());
    HOP_CUDA(ctx, bMeta.alloc(sizeof(int) * meta.size()));
    HOP_CUDA(ctx, cudaMemcpyAsync(bMeta.p, meta.data(), sizeof(int) * meta.size(), cudaMemcpyHostToDevice, st));
    const int *d_ex_begin = bMeta.as<int>(), *d_rec_base = d_ex_begin + 2 * T, *d_r1_begin = d_rec_base + 2 * T, *d_r1_end = d_r1_begin + T,
              *d_r2_begin = d_r1_end + T, *d_r2_end = d_r2_begin + T;
    prof.reset(new ProfScope(ctx, HOP_PROF_S4_JOIN));
    prepare_pairs_kernel<<<(n_sel + 127) / 128, 128, 0, st>>>(bSel.as<long long>(), n_sel, NP, d_ex_begin, bQp.as<float4>(), bQu.as<float4>(),
                                                              bTp.as<TrialParams>(), g, bR1.as<PairRec1>(), bR2.as<PairRec2>(), d_rec_base);
    HOP_CUDA(ctx, bR1Trial.alloc(sizeof(int) * (size_t)n1));
    fill_int_ranges_kernel<<<dim3(8, T), 128, 0, st>>>(bR1Trial.as<int>(), d_r1_begin, d_r1_end, T);
    HOP_CUDA(ctx, bCounts.alloc(sizeof(int) * ((size_t)n1 + 1))); HOP_CUDA(ctx, bOffsets.alloc(sizeof(int) * ((size_t)n1 + 1)));
    HOP_CUDA(ctx, cudaMemsetAsync(bCounts.p, 0, sizeof(int) * ((size_t)n1 + 1), st));
    JoinArgs ja;
    ja.r1 = bR1.as<PairRec1>(); ja.r2 = bR2.as<PairRec2>(); ja.r1_trial = bR1Trial.as<int>(); ja.r2_begin = d_r2_begin; ja.r2_end = d_r2_end;
    ja.n1 = n1; ja.eg_size = g.eg_size; ja.delta = delta; ja.counts = bCounts.as<int>(); ja.offsets = nullptr; ja.quads = nullptr; ja.quad_trial = nullptr;
    congruent_join_kernel<false><<<(n1 + 7) / 8, 256, 0, st>>>(ja);
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, bCounts.as<int>(), bOffsets.as<int>(), n1 + 1, st);
    if (scan_bytes > cub_bytes) { HOP_CUDA(ctx, bCub.alloc(scan_bytes)); cub_bytes = scan_bytes; }
    cub::DeviceScan::ExclusiveSum(bCub.p, scan_bytes, bCounts.as<int>(), bOffsets.as<int>(), n1 + 1, st);
    ctx->launches += 4;
    prof.reset();
    // quadrilaterals per trial -> which trials Perform_N_steps would have executed
    std::vector<int> offs(n1 + 1);
    HOP_CUDA(ctx, cudaMemcpyAsync(offs.data(), bOffsets.p, sizeof(int) * ((size_t)n1 + 1), cudaMemcpyDeviceToHost, st));
    HOP_CUDA(ctx, cudaStreamSynchronize(st));
    trace.reset(new HopTraceScope(ctx, "  s4run: K2b fill"));
    int successes = 0;
    T_exec = T;
    for (int t = 0; t < T; ++t) {
      quad_begin[t] = offs[r1_begin[t]];
      const int nq_t = offs[r1_end[t]] - offs[r1_begin[t]];
      // generateCongruents succeeds when both pair sets and the congruent set are non-empty (match4pcsBase.hpp:266-280)
      if (plan->trials[t].base_ok && cnt[2 * t] > 0 && cnt[2 * t + 1] > 0 && nq_t > 0) ++successes;
      if (successes >= plan->opt.success_quadrilaterals) { T_exec = t + 1; break; }   // Perform_N_steps (:186)
    }
    for (int t = 0; t < T_exec; ++t) quad_begin[t] = offs[r1_begin[t]];
    for (int t = T_exec; t <= T; ++t) quad_begin[t] = offs[r1_end[T_exec - 1]];
    M = offs[r1_end[T_exec - 1]];
    if (M > 0) {
      HOP_CUDA(ctx, bQuads.alloc(sizeof(int4) * (size_t)M)); HOP_CUDA(ctx, bQuadTrial.alloc(sizeof(int) * (size_t)M));
      ja.n1 = r1_end[T_exec - 1]; ja.offsets = bOffsets.as<int>(); ja.quads = bQuads.as<int4>(); ja.quad_trial = bQuadTrial.as<int>();
      {
        ProfScope ps(ctx, HOP_PROF_S4_JOIN);
        congruent_join_kernel<true><<<(ja.n1 + 7) / 8, 256, 0, st>>>(ja);
      }
      ctx->launches += 1;
    }
  } else {
    // no pair in any trial: every planned trial runs and fails
    T_exec = T;
  }
  plan->trials_executed = T_exec;

  if (keep) {
    // pairs (first sets then second sets, reference layout (i,j),(j,i)) and quadrilaterals back to the host for inspection
    std::vector<PairRec1> h1(n1); std::vector<PairRec2> h2(n2);
    if (n1 > 0 && n2 > 0) {
      HOP_CUDA(ctx, cudaMemcpyAsync(h1.data(), bR1.p, sizeof(PairRec1) * (size_t)n1, cudaMemcpyDeviceToHost, st));
      HOP_CUDA(ctx, cudaMemcpyAsync(h2.data(), bR2.p, sizeof(PairRec2) * (size_t)n2, cudaMemcpyDeviceToHost, st));
      HOP_CUDA(ctx, cudaStreamSynchronize(st));
    } else { n1 = n2 = 0; }
    plan->pairs.resize(2 * (size_t)(n1 + n2));
    for (int k = 0; k < n1; ++k) { plan->pairs[2 * k] = h1[k].a; plan->pairs[2 * k + 1] = h1[k].b; }
    for (int k = 0; k < n2; ++k) { plan->pairs[2 * (n1 + k)] = h2[k].a; plan->pairs[2 * (n1 + k) + 1] = h2[k].b; }
    for (int t = 0; t < T; ++t) {
      int *r = plan->trial_ranges.data() + 6 * t;
      r[0] = r1_begin[t]; r[1] = r1_end[t]; r[2] = n1 + r2_begin[t]; r[3] = n1 + r2_end[t];
      r[4] = t < T_exec ? quad_begin[t] : M; r[5] = t < T_exec ? quad_begin[t + 1] : M;
    }
    plan->quads.resize(4 * (size_t)M);
    if (M > 0) {
      HOP_CUDA(ctx, cudaMemcpyAsync(plan->quads.data(), bQuads.p, sizeof(int4) * (size_t)M, cudaMemcpyDeviceToHost, st));
      HOP_CUDA(ctx, cudaStreamSynchronize(st));
    }
  }
  if (M == 0) return HOP_OK;

  trace.reset(new HopTraceScope(ctx, "  s4run: K3 (cloud update, grid, verify, download)"));
  // ---- K3 on every quadrilateral of the executed trials ----
  std::vector<float> Qc(3 * (size_t)nQ);
  for (int i = 0; i < nQ; ++i) for (int k = 0; k < 3; ++k) Qc[3 * i + k] = plan->Q[i].p[k];
  hop_cloud *Pcloud = ctx->s4_scene;
  int rc = HOP_OK;
  std::vector<int32_t> bases(4 * (size_t)T);
  for (int t = 0; t < T; ++t) for (int k = 0; k < 4; ++k) bases[4 * t + k] = plan->trials[t].base[k];
  DevBuf bBases(st), bQc(st), bPoses(st), bLcp(st), bValid(st), bN(st);
  cudaError_t ce = cudaSuccess;
  if ((ce = bBases.alloc(sizeof(int32_t) * 4 * T)) != cudaSuccess || (ce = bQc.alloc(sizeof(float) * 3 * nQ)) != cudaSuccess ||
      (ce = bPoses.alloc(64 * (size_t)M)) != cudaSuccess || (ce = bLcp.alloc(4 * (size_t)M)) != cudaSuccess ||
      (ce = bValid.alloc(4 * (size_t)M)) != cudaSuccess || (ce = bN.alloc(16)) != cudaSuccess) {
    ctx->err = std::string("hop_super4pcs_run: ") + cudaGetErrorString(ce);
    return HOP_ENOMEM;
  }
  cudaMemsetAsync(bPoses.p, 0, 64 * (size_t)M, st); cudaMemsetAsync(bLcp.p, 0, 4 * (size_t)M, st);   // (the records past the emitted count travel too)
  cudaMemcpyAsync(bBases.p, bases.data(), sizeof(int32_t) * 4 * T, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(bQc.p, Qc.data(), sizeof(float) * 3 * nQ, cudaMemcpyHostToDevice, st);
  rc = hop_verify_lcp_dev(ctx, Pcloud, bQc.as<float>(), nQ, bBases.as<int32_t>(), T, (const int32_t *)bQuads.p, bQuadTrial.as<int32_t>(), M,
                          plan->centroid_P, plan->centroid_Q, delta, bPoses.as<float>(), bLcp.as<float>(), bValid.as<int32_t>(), bN.as<int32_t>());
  int32_t n = 0;
  if (rc == HOP_OK) {
    // one round trip: the count and, with it, as many records as the caller has room for (n <= M is known here; the records past n
    // are whatever the buffers held and are not reported)
    const int k = std::min<int>(M, capacity);
    cudaMemcpyAsync(&n, bN.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    if (k > 0) {
      cudaMemcpyAsync(hyp_poses, bPoses.p, 64 * (size_t)k, cudaMemcpyDeviceToHost, st);
      cudaMemcpyAsync(hyp_lcp, bLcp.p, 4 * (size_t)k, cudaMemcpyDeviceToHost, st);
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) rc = HOP_ECUDA;
    if (rc == HOP_OK && n_hyp) *n_hyp = n;
  }
  if (rc == HOP_ECUDA) ctx->err = std::string("hop_super4pcs_run: ") + cudaGetErrorString(cudaGetLastError());
  return rc;
}
