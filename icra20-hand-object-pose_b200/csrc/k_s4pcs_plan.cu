// k_s4pcs_plan.cu -- the planner of Super4PCS: everything that consumes the reference's random streams, planned for all
// trials before the congruent-set kernels start.  The random draws are serial and stay on the host; what they consult -- "does the
// PPF key of this pair of scene points exist in the model's table?" (gr::computePPF + the std::map lookup, matchBase.hpp:27-68,
// 134,159,201) -- is data parallel and comes from the device when a context is given (hop_s4pcs_plan_create_gpu): one launch fills a
// bit matrix of ALL pairs of scene points (up to 6144 points; row by row on demand above), the host planner then only reads bits.
// The keys truncate acos(.) / pi * 180 to an integer: a pair whose angle lies within 2e-4 degrees of an integer (or whose cosine is
// within 1e-6 of +-1) is flagged by the kernel and re-evaluated on the host with the reference's libm, so the pools -- and with them
// the replayed random streams -- stay bit-identical (tests/test_s4pcs_plan.py, tests/test_gpu_s4pcs.py).
// hop_ppf_table_build does the same for the model's table itself (computePPF.cpp:56-107: all pairs of the 5 mm model).
//
//   MatchBase::init                 src/OpenGR_4pcs/src/gr/algorithms/matchBase.hpp:382-462
//   UniformDistSampler              src/OpenGR_4pcs/src/gr/sampling.h:67-145
//   PairCreationFunctor::synch3DContent   .../pairCreationFunctor.h:129-161
//   computePPF / ppfClosestBin      .../matchBase.hpp:27-68
//   SelectRandomTriangle            .../matchBase.hpp:111-212
//   SelectQuadrilateral / TryQuadrilateral / distSegmentToSegment   .../match4pcsBase.hpp:50-189,287-354
//
// The reference draws from two std::mt19937 streams through std::shuffle, operator() % n and
// std::discrete_distribution<>; this file uses the same standard-library classes, so built with the same libstdc++ it
// replays the same sequence.  Float expressions are written in the order Eigen evaluates them on the reference's build
// (no FMA, fixed-size 3-term reductions as t0 + (t1 + t2)), so bases and invariants come out bit-identical to the
// compiled reference (tests/test_s4pcs_plan.py pins this against oracle/_ref).
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <limits>
#include <random>
#include <unordered_set>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include <unordered_map>

#include "hop_common.cuh"
#include "s4pcs.h"

namespace {

typedef std::array<float, 3> V3;

inline V3 sub(const float *a, const float *b) { return V3{a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
inline float dot(const V3 &a, const V3 &b) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); }
inline float sqnorm(const V3 &a) { return dot(a, a); }
inline float norm(const V3 &a) { return std::sqrt(sqnorm(a)); }
inline V3 normalized(const V3 &a) {  // Eigen: z = squaredNorm(); z > 0 ? a / sqrt(z) : a
  const float z = sqnorm(a);
  if (z > 0.f) { const float s = std::sqrt(z); return V3{a[0] / s, a[1] / s, a[2] / s}; }
  return a;
}

inline int ppf_closest_bin(int value, int discretization) {
  int lower_limit = value - (value % discretization);
  int upper_limit = lower_limit + discretization;
  int dist_from_lower = value - lower_limit;
  int dist_from_upper = upper_limit - value;
  return (dist_from_lower < dist_from_upper) ? lower_limit : upper_limit;
}

// gr::computePPF (matchBase.hpp:44-68)
void compute_ppf(const S4Pt &a, const S4Pt &b, int *ppf) {
  V3 n1 = normalized(V3{a.n[0], a.n[1], a.n[2]}), n2 = normalized(V3{b.n[0], b.n[1], b.n[2]});
  n1 = normalized(n1); n2 = normalized(n2);  // `.normalized()` then `.normalize()`
  const int dist = static_cast<int>(norm(sub(a.p, b.p)) * 1000);
  const V3 d = normalized(sub(b.p, a.p));
  const int n1_p1p2 = static_cast<int>(std::acos(dot(n1, d)) / M_PI * 180);
  const int n2_p1p2 = static_cast<int>(std::acos(dot(n2, d)) / M_PI * 180);
  const int n1_n2 = static_cast<int>(std::acos(dot(n1, n2)) / M_PI * 180);
  ppf[0] = ppf_closest_bin(dist, 5);
  ppf[1] = ppf_closest_bin(n1_p1p2, 10);
  ppf[2] = ppf_closest_bin(n2_p1p2, 10);
  ppf[3] = ppf_closest_bin(n1_n2, 10);
}

inline uint64_t pack_key(const int *k) {
  return ((uint64_t)(uint16_t)k[0] << 48) | ((uint64_t)(uint16_t)k[1] << 32) | ((uint64_t)(uint16_t)k[2] << 16) | (uint64_t)(uint16_t)k[3];
}


// ---- device side: PPF keys of point pairs ---------------------------------------------------------------------------------
__device__ __forceinline__ float d_dot3(const float *a, const float *b) {   // Eigen's 3-term reduction: t0 + (t1 + t2), no FMA
  return __fadd_rn(__fmul_rn(a[0], b[0]), __fadd_rn(__fmul_rn(a[1], b[1]), __fmul_rn(a[2], b[2])));
}
__device__ __forceinline__ void d_normalize(float *a) {
  const float z = d_dot3(a, a);
  if (z > 0.f) { const float sq = __fsqrt_rn(z); a[0] = __fdiv_rn(a[0], sq); a[1] = __fdiv_rn(a[1], sq); a[2] = __fdiv_rn(a[2], sq); }
}
__device__ __forceinline__ int d_closest_bin(int value, int discretization) {
  const int lower = value - (value % discretization), upper = lower + discretization;
  return (value - lower < upper - value) ? lower : upper;
}
// angle bin of acos(x) / pi * 180 the way the host computes it (float acosf, double division); `amb` is raised when the host's libm
// could land on the other side of an integer
__device__ __forceinline__ int d_angle_deg(float x, bool &amb) {
  if (!(fabsf(x) <= 1.f - 1e-6f)) { amb = true; return 0; }
  const float af = (float)acos((double)x);
  const double deg = (double)af / M_PI * 180.0;
  const double fr = deg - floor(deg);
  if (fr < 2e-4 || fr > 1.0 - 2e-4) amb = true;
  return (int)deg;
}
// gr::computePPF for the pair (a, b); returns the packed key (or ~0 when a field leaves 16 bits)
__device__ __forceinline__ uint64_t d_ppf_key(const float4 pa, const float4 na, const float4 pb, const float4 nb, bool &amb) {
  float n1[3] = {na.x, na.y, na.z}, n2[3] = {nb.x, nb.y, nb.z};
  d_normalize(n1); d_normalize(n1); d_normalize(n2); d_normalize(n2);   // compute_ppf normalises twice (the points' normals once already)
  const float ab[3] = {__fsub_rn(pa.x, pb.x), __fsub_rn(pa.y, pb.y), __fsub_rn(pa.z, pb.z)};
  const int dist = (int)__fmul_rn(__fsqrt_rn(d_dot3(ab, ab)), 1000.f);
  float d[3] = {__fsub_rn(pb.x, pa.x), __fsub_rn(pb.y, pa.y), __fsub_rn(pb.z, pa.z)};
  d_normalize(d);
  const int a1 = d_angle_deg(d_dot3(n1, d), amb), a2 = d_angle_deg(d_dot3(n2, d), amb), a3 = d_angle_deg(d_dot3(n1, n2), amb);
  const int k[4] = {d_closest_bin(dist, 5), d_closest_bin(a1, 10), d_closest_bin(a2, 10), d_closest_bin(a3, 10)};
  if (k[0] < 0 || k[0] > 65535 || k[1] < 0 || k[1] > 65535 || k[2] < 0 || k[2] > 65535 || k[3] < 0 || k[3] > 65535) return ~0ull;
  return ((uint64_t)k[0] << 48) | ((uint64_t)k[1] << 32) | ((uint64_t)k[2] << 16) | (uint64_t)k[3];
}
__device__ __forceinline__ bool d_key_in_table(const uint64_t *__restrict__ keys, int n, uint64_t key) {
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] < key) lo = mid + 1; else hi = mid; }
  return lo < n && keys[lo] == key;
}
// rows[r] x every point j: member bit = key(P[rows[r]], P[j]) is in the table, amb bit = the host must re-evaluate the pair
__global__ void ppf_rows_kernel(const float4 *__restrict__ pos, const float4 *__restrict__ nrm, int n, const int *__restrict__ rows, int n_rows,
                                const uint64_t *__restrict__ keys, int n_keys, uint32_t *__restrict__ member, uint32_t *__restrict__ ambig, int words) {
  const int r = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = rows ? rows[r] : r;
  bool amb = false, mem = false;
  if (j < n && j != i) {
    const uint64_t key = d_ppf_key(pos[i], nrm[i], pos[j], nrm[j], amb);
    mem = key != ~0ull && d_key_in_table(keys, n_keys, key);
  }
  const unsigned mb = __ballot_sync(0xffffffffu, mem), ab = __ballot_sync(0xffffffffu, amb);
  if ((threadIdx.x & 31) == 0 && (j >> 5) < words) { member[(size_t)r * words + (j >> 5)] = mb; ambig[(size_t)r * words + (j >> 5)] = ab; }
}
// all pairs i < j of a cloud: packed key (or ~0 for an ambiguous / out-of-range pair, which the host re-evaluates)
__global__ void ppf_pairs_kernel(const float4 *__restrict__ pos, const float4 *__restrict__ nrm, int n, uint64_t *__restrict__ keys, unsigned char *__restrict__ ambig) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n * n) return;
  const int i = (int)(t / n), j = (int)(t - (size_t)i * n);
  uint64_t key = ~0ull;
  bool amb = false;
  if (i < j) { key = d_ppf_key(pos[i], nrm[i], pos[j], nrm[j], amb); if (amb) key = ~0ull; }
  keys[t] = key;
  ambig[t] = (i < j && amb) ? 1 : 0;
}

// "is the PPF key of (P[a], P[b]) in the table?" answered from device-computed bits; ambiguous pairs fall back to the host formula
struct PpfBits {
  hop_ctx *ctx = nullptr;
  int n = 0, words = 0, n_keys = 0;
  bool full = false;
  float4 *d_pos = nullptr, *d_nrm = nullptr;
  uint64_t *d_keys = nullptr;
  uint32_t *d_member = nullptr, *d_ambig = nullptr;
  int *d_rows = nullptr;
  std::vector<uint32_t> member, ambig;                       // full mode: n x words (backing store when the matrix was made on the host)
  const uint32_t *member_p = nullptr, *ambig_p = nullptr;    // full mode: the matrices (the context's pinned buffer after a device build)
  std::unordered_map<int, std::pair<std::vector<uint32_t>, std::vector<uint32_t>>> rows;   // row mode: fetched on demand
  int launches = 0;
  // stream-ordered allocations from the context's pool: a plan is made once per frame
  ~PpfBits() {
    if (!ctx) return;
    HopDeviceGuard g(ctx);
    void *bufs[6] = {d_pos, d_nrm, d_keys, d_member, d_ambig, d_rows};
    for (void *b : bufs) if (b) cudaFreeAsync(b, ctx->stream);
  }
  int init(hop_ctx *c, const std::vector<S4Pt> &P, const std::vector<uint64_t> &sorted_keys) {
    ctx = c; n = (int)P.size(); words = (n + 31) / 32; n_keys = (int)sorted_keys.size();
    if (n == 0) return HOP_OK;
    std::vector<float4> hp(n), hn(n);
    for (int i = 0; i < n; ++i) { hp[i] = make_float4(P[i].p[0], P[i].p[1], P[i].p[2], 0.f); hn[i] = make_float4(P[i].n[0], P[i].n[1], P[i].n[2], 0.f); }
    HOP_CUDA(ctx, cudaMallocAsync(&d_pos, sizeof(float4) * (size_t)n, ctx->stream)); HOP_CUDA(ctx, cudaMallocAsync(&d_nrm, sizeof(float4) * (size_t)n, ctx->stream));
    HOP_CUDA(ctx, cudaMallocAsync(&d_keys, sizeof(uint64_t) * (size_t)std::max(n_keys, 1), ctx->stream));
    HOP_CUDA(ctx, cudaMemcpyAsync(d_pos, hp.data(), sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    HOP_CUDA(ctx, cudaMemcpyAsync(d_nrm, hn.data(), sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    if (n_keys) HOP_CUDA(ctx, cudaMemcpyAsync(d_keys, sorted_keys.data(), sizeof(uint64_t) * (size_t)n_keys, cudaMemcpyHostToDevice, ctx->stream));
    full = n <= 6144;
    const size_t rows_alloc = full ? (size_t)n : 64;
    HOP_CUDA(ctx, cudaMallocAsync(&d_member, sizeof(uint32_t) * rows_alloc * words, ctx->stream)); HOP_CUDA(ctx, cudaMallocAsync(&d_ambig, sizeof(uint32_t) * rows_alloc * words, ctx->stream));
    HOP_CUDA(ctx, cudaMallocAsync(&d_rows, sizeof(int) * 64, ctx->stream));
    if (full) {
      ppf_rows_kernel<<<dim3((n + 255) / 256, n), 256, 0, ctx->stream>>>(d_pos, d_nrm, n, nullptr, n, d_keys, n_keys, d_member, d_ambig, words);
      ctx->launches += 1; ++launches;
      // into pinned memory: a pageable destination makes each copy a staged, synchronous one (1 MB at 2 k points)
      const size_t cnt = (size_t)n * words;
      uint32_t *pin = (uint32_t *)ctx->ensure_pinned(2 * sizeof(uint32_t) * cnt);
      if (!pin) { ctx->err = "hop_s4pcs_plan_create_gpu: pinned buffer allocation failed"; return HOP_ENOMEM; }
      HOP_CUDA(ctx, cudaMemcpyAsync(pin, d_member, sizeof(uint32_t) * cnt, cudaMemcpyDeviceToHost, ctx->stream));
      HOP_CUDA(ctx, cudaMemcpyAsync(pin + cnt, d_ambig, sizeof(uint32_t) * cnt, cudaMemcpyDeviceToHost, ctx->stream));
      member_p = pin; ambig_p = pin + cnt;
    }
    HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return HOP_OK;
  }
  // the row of point a (row mode: one launch + copy per first use)
  bool row(int a, const uint32_t *&m, const uint32_t *&am) {
    if (full) { m = member_p + (size_t)a * words; am = ambig_p + (size_t)a * words; return true; }
    auto it = rows.find(a);
    if (it == rows.end()) {
      if (cudaMemcpyAsync(d_rows, &a, sizeof(int), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return false;
      ppf_rows_kernel<<<dim3((n + 255) / 256, 1), 256, 0, ctx->stream>>>(d_pos, d_nrm, n, d_rows, 1, d_keys, n_keys, d_member, d_ambig, words);
      ctx->launches += 1; ++launches;
      std::pair<std::vector<uint32_t>, std::vector<uint32_t>> r;
      r.first.resize(words); r.second.resize(words);
      if (cudaMemcpyAsync(r.first.data(), d_member, sizeof(uint32_t) * words, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
          cudaMemcpyAsync(r.second.data(), d_ambig, sizeof(uint32_t) * words, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
          cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        return false;
      it = rows.emplace(a, std::move(r)).first;
    }
    m = it->second.first.data(); am = it->second.second.data();
    return true;
  }
};

struct Planner {
  hop_s4pcs_plan &pl;
  std::unordered_set<uint64_t> keys;       // host path: a lookup per pair, millions per plan
  std::vector<uint64_t> sorted_keys;       // device path: the table the membership kernel searches; the host only looks up the flagged pairs
  std::vector<float> point_probs;   // _point_probs (anneals across trials)
  std::mt19937 random_generator;    // randomGenerator_
  std::mt19937 point_index_engine;  // _point_index_engine, seeded 0 (matchBase.hpp:76)
  float max_base_diameter = -1.f;
  std::vector<unsigned char> member;   // scratch of select_random_triangle

  explicit Planner(hop_s4pcs_plan &p) : pl(p), random_generator(p.opt.random_seed ? p.opt.random_seed : std::mt19937::default_seed), point_index_engine(0) {}

  PpfBits *bits = nullptr;   // device-computed membership (hop_s4pcs_plan_create_gpu), null = evaluate on the host
  bool has_ppf(const S4Pt &a, const S4Pt &b) const {
    int k[4];
    compute_ppf(a, b, k);
    if (k[0] < 0 || k[0] > 65535 || k[1] < 0 || k[1] > 65535 || k[2] < 0 || k[2] > 65535 || k[3] < 0 || k[3] > 65535) return false;
    const uint64_t key = pack_key(k);
    return sorted_keys.empty() ? keys.count(key) != 0 : std::binary_search(sorted_keys.begin(), sorted_keys.end(), key);
  }
  // membership of the pair of scene points (a, b): a bit of the device matrix, the host formula for flagged pairs
  bool has_ppf_idx(int a, int b) const {
    if (bits && a != b) {
      const uint32_t *m, *am;
      if (bits->full || bits->rows.count(a)) {
        if (bits->row(a, m, am)) {
          if (!((am[b >> 5] >> (b & 31)) & 1u)) return (m[b >> 5] >> (b & 31)) & 1u;
        }
      }
    }
    return has_ppf(pl.P[a], pl.P[b]);
  }
  // the whole row of point a at once (the O(N) scans): member[i] for every i
  void ppf_row(int a, std::vector<unsigned char> &out, const std::vector<int> *subset) const {
    const std::vector<S4Pt> &P = pl.P;
    const int n = subset ? (int)subset->size() : (int)P.size();
    out.assign(n, 0);
    const uint32_t *m = nullptr, *am = nullptr;
    const bool dev = bits && bits->row(a, m, am);
#pragma omp parallel for schedule(static) if (!dev && n >= 512)
    for (int t = 0; t < n; ++t) {
      const int i = subset ? (*subset)[t] : t;
      if (i == a) continue;
      if (dev && !((am[i >> 5] >> (i & 31)) & 1u)) out[t] = (m[i >> 5] >> (i & 31)) & 1u;
      else out[t] = has_ppf(P[a], P[i]) ? 1 : 0;
    }
  }

  // MatchBase::init
  void init(const float *P_xyz, const float *P_nrm, const float *P_prob, int nP, const float *Q_xyz, const float *Q_nrm, int nQ) {
    auto fill = [](const float *xyz, const float *nrm, int i) {  // fillPointSet + Point3D::set_normal (normalises once)
      S4Pt q;
      for (int k = 0; k < 3; ++k) q.p[k] = xyz[3 * i + k];
      V3 n = normalized(V3{nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]});
      for (int k = 0; k < 3; ++k) q.n[k] = n[k];
      return q;
    };
    pl.P.resize(nP);
    pl.P_prob.resize(nP);
    for (int i = 0; i < nP; ++i) { pl.P[i] = fill(P_xyz, P_nrm, i); pl.P_prob[i] = P_prob ? P_prob[i] : 1.f; }
    pl.Q.clear(); pl.q_ids.clear();
    if ((size_t)nQ > (size_t)pl.opt.sample_size) {
      // UniformDistSampler: the first point met in every voxel of edge delta, in input order
      struct VoxHash { size_t operator()(const std::array<int, 3> &c) const { return (size_t)(100000007ull * (uint64_t)c[0] + 161803409ull * (uint64_t)c[1] + 423606823ull * (uint64_t)c[2]); } };
      std::unordered_set<std::array<int, 3>, VoxHash> seen;
      const float scale = 1.0f / pl.opt.delta;
      std::vector<int> uniform;
      for (int i = 0; i < nQ; ++i) {
        std::array<int, 3> c{int(std::floor(Q_xyz[3 * i] * scale)), int(std::floor(Q_xyz[3 * i + 1] * scale)), int(std::floor(Q_xyz[3 * i + 2] * scale))};
        if (seen.insert(c).second) uniform.push_back(i);
      }
      std::shuffle(uniform.begin(), uniform.end(), random_generator);
      const size_t nb = std::min(uniform.size(), (size_t)pl.opt.sample_size);
      for (size_t k = 0; k < nb; ++k) { pl.Q.push_back(fill(Q_xyz, Q_nrm, uniform[k])); pl.q_ids.push_back(uniform[k]); }
    } else {
      for (int i = 0; i < nQ; ++i) { pl.Q.push_back(fill(Q_xyz, Q_nrm, i)); pl.q_ids.push_back(i); }
    }
    point_probs = pl.P_prob;
    auto center = [](std::vector<S4Pt> &c, float *centroid) {
      centroid[0] = centroid[1] = centroid[2] = 0.f;
      for (const S4Pt &q : c) for (int k = 0; k < 3; ++k) centroid[k] += q.p[k];
      const float n = (float)c.size();
      for (int k = 0; k < 3; ++k) centroid[k] /= n;
      for (S4Pt &q : c) for (int k = 0; k < 3; ++k) q.p[k] -= centroid[k];
    };
    center(pl.P, pl.centroid_P);
    center(pl.Q, pl.centroid_Q);
    // "diameter of P": 1000 random pairs -- of Q (matchBase.hpp:439-448)
    pl.diameter = 0.f;
    const size_t nq = pl.Q.size();
    if (nq > 0)
      for (int i = 0; i < 1000; ++i) {
        int at = random_generator() % nq;
        int bt = random_generator() % nq;
        float l = norm(sub(pl.Q[bt].p, pl.Q[at].p));
        if (l > pl.diameter) pl.diameter = l;
      }
    max_base_diameter = pl.diameter;
    // PairCreationFunctor::synch3DContent: Q in the unit cube
    float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
    float mx[3] = {std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest()};
    for (const S4Pt &q : pl.Q) for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], q.p[k]); mx[k] = std::max(mx[k], q.p[k]); }
    float diag = 0.f;
    for (int k = 0; k < 3; ++k) { pl.gcenter[k] = (mn[k] + mx[k]) / 2.f; diag = k == 0 ? mx[k] - mn[k] : std::max(diag, mx[k] - mn[k]); }
    pl.ratio = (float)((double)diag + 0.001);
    pl.Qunit.resize(3 * nq);
    for (size_t i = 0; i < nq; ++i) for (int k = 0; k < 3; ++k) pl.Qunit[3 * i + k] = (pl.Q[i].p[k] - pl.gcenter[k]) / pl.ratio + 0.5f;
  }

  // One draw of std::discrete_distribution<int>(w, w + n) from `engine`, without building the distribution (libstdc++ 13,
  // bits/random.tcc: param_type::_M_initialize normalises every weight by the sequential double sum, partial_sum gives the
  // cumulative table with its last entry forced to 1, operator() draws generate_canonical<double, 53> and returns the
  // upper_bound).  Same doubles in the same order, so the same index; the scan stops at the answer instead of finishing the
  // table, and nothing is allocated.  The reference constructs a distribution per draw (matchBase.hpp:120-140): at ~1000
  // points and 30 trials that construction was most of the planner's host time.  tests/test_s4pcs_plan.py pins the bases bit for bit
  // against the reference compiled here, i.e. against the real std::discrete_distribution.
  template <typename W>
  static double weight_sum(const W *w, int n) {   // std::accumulate(prob.begin(), prob.end(), 0.0)
    double sum = 0.0;
    for (int i = 0; i < n; ++i) sum += (double)w[i];
    return sum;
  }
  template <typename W>
  static int draw_discrete(const W *w, int n, std::mt19937 &engine) { return draw_discrete(w, n, weight_sum(w, n), engine); }
  // (sum = weight_sum(w, n): callers that draw twice from unchanged weights add them up once)
  template <typename W>
  static int draw_discrete(const W *w, int n, double sum, std::mt19937 &engine) {
    if (n < 2) return 0;   // (an empty or one-entry table: libstdc++ returns 0 without touching the engine)
    const double p = std::generate_canonical<double, std::numeric_limits<double>::digits>(engine);
    double acc = 0.0;
    for (int k = 0; k + 1 < n; ++k) {
      acc = k == 0 ? (double)w[0] / sum : acc + (double)w[k] / sum;
      if (acc > p) return k;
    }
    return n - 1;
  }

  // SelectRandomTriangle (matchBase.hpp:111-212).  sample_pool is returned as the pool for the 4th point.
  std::vector<float> tri_probs;   // scratch of select_random_triangle (kept across calls: no allocation per trial)
  std::vector<int> tri_backup;
  bool select_random_triangle(int &base1, int &base2, int &base3, std::vector<int> &sample_pool) {
    const std::vector<S4Pt> &P = pl.P;
    const int number_of_points = (int)P.size();
    base1 = base2 = base3 = -1;
    const int first_point = draw_discrete(point_probs.data(), number_of_points, point_index_engine);
    point_probs[first_point] *= pl.opt.dispersion;
    sample_pool.clear();
    std::vector<float> &probs = tri_probs;
    probs.clear();
    // (the membership tests -- three acos and a set lookup per point -- are the planner's O(N) cost: a bit of the device's matrix
    //  when there is one, else evaluated on all host threads with the same libm, gathered in index order, so the pool is the one
    //  a sequential loop builds)
    const uint32_t *m1 = nullptr, *a1 = nullptr;
    if (bits && bits->row(first_point, m1, a1)) {
      // straight from the device's bit row: 32 points per word, the host formula only for the flagged (ambiguous) pairs
      for (int w0 = 0; w0 < number_of_points; w0 += 32) {
        const uint32_t amb = a1[w0 >> 5];
        uint32_t mem = m1[w0 >> 5] & ~amb;
        for (uint32_t rest = amb; rest; rest &= rest - 1) {
          const int i = w0 + __builtin_ctz(rest);
          if (i < number_of_points && i != first_point && has_ppf(P[first_point], P[i])) mem |= 1u << (i - w0);
        }
        if (first_point >= w0 && first_point < w0 + 32) mem &= ~(1u << (first_point - w0));
        for (; mem; mem &= mem - 1) {
          const int i = w0 + __builtin_ctz(mem);
          if (i >= number_of_points) break;
          sample_pool.push_back(i); probs.push_back(point_probs[i]);
        }
      }
    } else {
      ppf_row(first_point, member, nullptr);
      for (int i = 0; i < number_of_points; ++i)
        if (member[i]) { sample_pool.push_back(i); probs.push_back(point_probs[i]); }
    }
    if (sample_pool.size() < 3) return false;
    const float sq_max_base_diameter = max_base_diameter * max_base_diameter;
    const int n_pool = (int)sample_pool.size();
    double pool_sum = weight_sum(probs.data(), n_pool);   // (the weights only change after a pair passed the PPF test)
    for (int i = 0; (size_t)i < sample_pool.size() * sample_pool.size() / 4; ++i) {
      const int second_point = draw_discrete(probs.data(), n_pool, pool_sum, point_index_engine);
      const int third_point = draw_discrete(probs.data(), n_pool, pool_sum, point_index_engine);
      if (second_point == third_point) continue;
      if (!has_ppf_idx(sample_pool[second_point], sample_pool[third_point])) continue;
      probs[second_point] *= pl.opt.dispersion;
      probs[third_point] *= pl.opt.dispersion;
      pool_sum = weight_sum(probs.data(), n_pool);
      const V3 u = sub(P[sample_pool[second_point]].p, P[first_point].p);
      const V3 w = sub(P[sample_pool[third_point]].p, P[first_point].p);
      const float how_wide = dot(normalized(u), normalized(w));
      if (std::abs(how_wide) <= std::cos(45 * M_PI / 180.0) && sqnorm(u) < sq_max_base_diameter && sqnorm(w) < sq_max_base_diameter) {
        base1 = first_point; base2 = sample_pool[second_point]; base3 = sample_pool[third_point];
        break;
      }
    }
    if (base2 == -1 || base3 == -1) return false;
    std::vector<int> &backup = tri_backup;
    backup.swap(sample_pool);
    sample_pool.clear();
    const int nb = (int)backup.size();
    // the reference stores the POOL INDEX i here, not the point id backup[i] (matchBase.hpp:203), and later uses it as
    // a point id (match4pcsBase.hpp:159): reproduced
    const uint32_t *m2 = nullptr, *a2 = nullptr, *m3 = nullptr, *a3 = nullptr;
    if (bits && bits->row(base2, m2, a2) && bits->row(base3, m3, a3)) {
      for (int i = 0; i < nb; ++i) {
        const int q = backup[i], w = q >> 5, b = q & 31;
        if (q == base2 || q == base3 || q == base1) continue;
        const bool in2 = ((a2[w] >> b) & 1u) ? has_ppf(P[base2], P[q]) : ((m2[w] >> b) & 1u) != 0;
        if (!in2) continue;
        const bool in3 = ((a3[w] >> b) & 1u) ? has_ppf(P[base3], P[q]) : ((m3[w] >> b) & 1u) != 0;
        if (in3) sample_pool.push_back(i);
      }
    } else {
      std::vector<unsigned char> m2, m3;
      ppf_row(base2, m2, &backup);
      ppf_row(base3, m3, &backup);
      for (int i = 0; i < nb; ++i) {
        if (backup[i] == base2 || backup[i] == base3 || backup[i] == base1) continue;
        if (m2[i] && m3[i]) sample_pool.push_back(i);
      }
    }
    if (sample_pool.size() < 1) return false;
    return base1 != -1 && base2 != -1 && base3 != -1;
  }

  static float dist_segment_to_segment(const float *p1, const float *p2, const float *q1, const float *q2, float &invariant1, float &invariant2) {
    static const float kSmallNumber = 0.0001;
    const V3 u = sub(p2, p1), v = sub(q2, q1), w = sub(p1, q1);
    const float a = dot(u, u), b = dot(u, v), c = dot(v, v), d = dot(u, w), e = dot(v, w);
    const float f = a * c - b * b;
    float s1 = 0.0, s2 = f, t1 = 0.0, t2 = f;
    if (f < kSmallNumber) { s1 = 0.0; s2 = 1.0; t1 = e; t2 = c; }
    else {
      s1 = (b * e - c * d);
      t1 = (a * e - b * d);
      if (s1 < 0.0) { s1 = 0.0; t1 = e; t2 = c; }
      else if (s1 > s2) { s1 = s2; t1 = e + b; t2 = c; }
    }
    if (t1 < 0.0) {
      t1 = 0.0;
      if (-d < 0.0) s1 = 0.0;
      else if (-d > a) s1 = s2;
      else { s1 = -d; s2 = a; }
    } else if (t1 > t2) {
      t1 = t2;
      if ((-d + b) < 0.0) s1 = 0;
      else if ((-d + b) > a) s1 = s2;
      else { s1 = (-d + b); s2 = a; }
    }
    invariant1 = (std::abs(s1) < kSmallNumber ? 0.0 : s1 / s2);
    invariant2 = (std::abs(t1) < kSmallNumber ? 0.0 : t1 / t2);
    const V3 r{(w[0] + invariant1 * u[0]) - invariant2 * v[0], (w[1] + invariant1 * u[1]) - invariant2 * v[1], (w[2] + invariant1 * u[2]) - invariant2 * v[2]};
    return norm(r);
  }

  // TryQuadrilateral (match4pcsBase.hpp:50-101): pick the pairing of the four points whose segments pass closest
  static bool try_quadrilateral(S4Pt *base3d, float &invariant1, float &invariant2, int *id) {
    float min_distance = std::numeric_limits<float>::max();
    int best1 = -1, best2 = -1, best3 = -1, best4 = -1;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        if (i == j) continue;
        int k = 0; while (k == i || k == j) k++;
        int l = 0; while (l == i || l == j || l == k) l++;
        float li1, li2;
        const float segment_distance = dist_segment_to_segment(base3d[i].p, base3d[j].p, base3d[k].p, base3d[l].p, li1, li2);
        if (segment_distance < min_distance) { min_distance = segment_distance; best1 = i; best2 = j; best3 = k; best4 = l; invariant1 = li1; invariant2 = li2; }
      }
    if (best1 < 0 || best2 < 0 || best3 < 0 || best4 < 0) return false;
    const S4Pt tmp[4] = {base3d[0], base3d[1], base3d[2], base3d[3]};
    base3d[0] = tmp[best1]; base3d[1] = tmp[best2]; base3d[2] = tmp[best3]; base3d[3] = tmp[best4];
    const int tid[4] = {id[0], id[1], id[2], id[3]};
    id[0] = tid[best1]; id[1] = tid[best2]; id[2] = tid[best3]; id[3] = tid[best4];
    return true;
  }

  // SelectQuadrilateral (match4pcsBase.hpp:107-189)
  bool select_quadrilateral(S4Trial &t) {
    const std::vector<S4Pt> &P = pl.P;
    const float kBaseTooSmall = 0.2;
    int current_trial = 0;
    int base1, base2, base3, base4;
    while (current_trial < 1000) {
      current_trial++;
      std::vector<int> sample_pool;
      if (!select_random_triangle(base1, base2, base3, sample_pool)) continue;
      S4Pt b3d[4];
      b3d[0] = P[base1]; b3d[1] = P[base2]; b3d[2] = P[base3];
      const double x1 = b3d[0].p[0], y1 = b3d[0].p[1], z1 = b3d[0].p[2];
      const double x2 = b3d[1].p[0], y2 = b3d[1].p[1], z2 = b3d[1].p[2];
      const double x3 = b3d[2].p[0], y3 = b3d[2].p[1], z3 = b3d[2].p[2];
      const float denom = (-x3 * y2 * z1 + x2 * y3 * z1 + x3 * y1 * z2 - x1 * y3 * z2 - x2 * y1 * z3 + x1 * y2 * z3);
      if (denom != 0) {
        const float A = (-y2 * z1 + y3 * z1 + y1 * z2 - y3 * z2 - y1 * z3 + y2 * z3) / denom;
        const float B = (x2 * z1 - x3 * z1 - x1 * z2 + x3 * z2 + x1 * z3 - x2 * z3) / denom;
        const float C = (-x2 * y1 + x3 * y1 + x1 * y2 - x3 * y2 - x1 * y3 + x2 * y3) / denom;
        base4 = -1;
        float best_distance = std::numeric_limits<float>::max();
        const float too_small = std::pow(max_base_diameter * kBaseTooSmall, 2);
        // the pool point nearest to the plane of the triangle, the FIRST of equals (the reference's strict '<' in pool order): on all host
        // threads for the pools of large scenes (18 of 27 ms per plan at 50 k points), every thread on an ascending block of the pool
        const int n_pool = (int)sample_pool.size();
        int best_i = -1;
#pragma omp parallel if (n_pool >= 4096)
        {
          int bi = -1;
          float bd = std::numeric_limits<float>::max();
#pragma omp for schedule(static) nowait
          for (int i = 0; i < n_pool; ++i) {
            const S4Pt &p = P[sample_pool[i]];
            if (sqnorm(sub(p.p, b3d[0].p)) >= too_small && sqnorm(sub(p.p, b3d[1].p)) >= too_small && sqnorm(sub(p.p, b3d[2].p)) >= too_small) {
              const float distance = std::abs(A * p.p[0] + B * p.p[1] + C * p.p[2] - 1.0);
              if (distance < bd) { bd = distance; bi = i; }
            }
          }
#pragma omp critical(hop_plan_base4)
          if (bi >= 0 && (bd < best_distance || (bd == best_distance && bi < best_i))) { best_distance = bd; best_i = bi; }
        }
        if (best_i >= 0) base4 = int(sample_pool[best_i]);
        if (base4 != -1) {
          b3d[3] = P[base4];
          int id[4] = {base1, base2, base3, base4};
          float inv1 = 0.f, inv2 = 0.f;
          if (try_quadrilateral(b3d, inv1, inv2, id)) {
            t.base_ok = 1;
            for (int k = 0; k < 4; ++k) { t.base[k] = id[k]; t.b[k] = b3d[k]; }
            t.inv1 = inv1; t.inv2 = inv2;
            return true;
          }
        }
      }
    }
    return false;
  }

  void plan_trials() {
    const int T = pl.opt.max_trials > 0 ? pl.opt.max_trials : 30;
    pl.trials.assign(T, S4Trial());
    if (pl.P.size() < 4 || pl.Q.size() < 4) return;
    for (int t = 0; t < T; ++t) {
      S4Trial &tr = pl.trials[t];
      if (!select_quadrilateral(tr)) continue;
      // generateCongruents (match4pcsBase.hpp:243-256)
      tr.dist1 = norm(sub(tr.b[0].p, tr.b[1].p));
      tr.dist2 = norm(sub(tr.b[2].p, tr.b[3].p));
      tr.nangle1 = norm(sub(tr.b[0].n, tr.b[1].n));
      tr.nangle2 = norm(sub(tr.b[2].n, tr.b[3].n));
      // FindCongruentQuadrilaterals (FunctorSuper4pcs.h:161-163)
      tr.alpha = dot(normalized(sub(tr.b[1].p, tr.b[0].p)), normalized(sub(tr.b[3].p, tr.b[2].p)));
    }
  }
};

}  // namespace

extern "C" {

void hop_default_s4pcs_options(hop_s4pcs_options *o) {
  if (!o) return;
  o->sample_size = 100; o->overlap = 0.2f; o->delta = 0.003f; o->dispersion = 0.5f; o->success_quadrilaterals = 10;
  o->max_normal_difference = -1.f; o->max_color_distance = -1.f; o->max_trials = 0; o->random_seed = 0; o->keep_intermediates = 0;
}

void hop_compute_ppf(const float *p1, const float *n1, const float *p2, const float *n2, int32_t *key) {
  S4Pt a, b;
  V3 na = normalized(V3{n1[0], n1[1], n1[2]}), nb = normalized(V3{n2[0], n2[1], n2[2]});  // Point3D::set_normal
  for (int k = 0; k < 3; ++k) { a.p[k] = p1[k]; b.p[k] = p2[k]; a.n[k] = na[k]; b.n[k] = nb[k]; }
  int k4[4];
  compute_ppf(a, b, k4);
  for (int k = 0; k < 4; ++k) key[k] = k4[k];
}

static int plan_create(hop_ctx *ctx, const float *P_xyz, const float *P_nrm, const float *P_prob, int nP, const float *Q_xyz, const float *Q_nrm,
                       int nQ, const int32_t *ppf_keys, int n_keys, const hop_s4pcs_options *opt, hop_s4pcs_plan **out) {
  if (!out) return HOP_EINVAL;
  *out = nullptr;
  if (!opt || nP < 0 || nQ < 0 || n_keys < 0 || (nP > 0 && (!P_xyz || !P_nrm)) || (nQ > 0 && (!Q_xyz || !Q_nrm)) || (n_keys > 0 && !ppf_keys) ||
      !(opt->delta > 0.f) || opt->sample_size < 1)
    return HOP_EINVAL;
  hop_s4pcs_plan *pl = new hop_s4pcs_plan();
  pl->opt = *opt;
  Planner planner(*pl);
  {
    HopTraceScope ts(ctx, "  plan: key set");
    const bool device_path = ctx != nullptr || getenv("HOP_PLAN_HOSTBITS") != nullptr;
    if (device_path) planner.sorted_keys.reserve((size_t)n_keys);
    for (int i = 0; i < n_keys; ++i) {
      const int k[4] = {ppf_keys[4 * i], ppf_keys[4 * i + 1], ppf_keys[4 * i + 2], ppf_keys[4 * i + 3]};
      if (k[0] >= 0 && k[0] <= 65535 && k[1] >= 0 && k[1] <= 65535 && k[2] >= 0 && k[2] <= 65535 && k[3] >= 0 && k[3] <= 65535) {
        if (device_path) planner.sorted_keys.push_back(pack_key(k)); else planner.keys.insert(pack_key(k));
      }
    }
    if (device_path) {
      std::sort(planner.sorted_keys.begin(), planner.sorted_keys.end());
      planner.sorted_keys.erase(std::unique(planner.sorted_keys.begin(), planner.sorted_keys.end()), planner.sorted_keys.end());
    }
  }
  {
    HopTraceScope ts(ctx, "  plan: MatchBase::init");
    planner.init(P_xyz, P_nrm, P_prob, nP, Q_xyz, Q_nrm, nQ);
  }
  PpfBits bits;
  if (ctx && pl->P.size() >= 4 && pl->Q.size() >= 4) {
    HopTraceScope ts(ctx, "  plan: PPF membership on the device");
    const int rc = bits.init(ctx, pl->P, planner.sorted_keys);
    if (rc != HOP_OK) { delete pl; return rc; }
    planner.bits = &bits;
  } else if (!ctx && getenv("HOP_PLAN_HOSTBITS")) {
    // debugging aid (no GPU needed): the membership matrix the device would deliver, computed with the host formula, so that the
    // trial loop runs the way it does behind hop_s4pcs_plan_create_gpu (HOP_PLAN_DEBUG prints its time)
    const int n = (int)pl->P.size();
    bits.n = n; bits.words = (n + 31) / 32; bits.full = true;
    bits.member.assign((size_t)n * bits.words, 0u); bits.ambig.assign((size_t)n * bits.words, 0u);
    for (int a = 0; a < n; ++a)
      for (int b = 0; b < n; ++b)
        if (a != b && planner.has_ppf(pl->P[a], pl->P[b])) bits.member[(size_t)a * bits.words + (b >> 5)] |= 1u << (b & 31);
    bits.member_p = bits.member.data(); bits.ambig_p = bits.ambig.data();
    planner.bits = &bits;
  }
  {
    HopTraceScope ts(ctx, "  plan: trials (RNG replay)");
    const auto t0 = std::chrono::steady_clock::now();
    planner.plan_trials();
    if (ctx ? ctx->tune.plan_debug : getenv("HOP_PLAN_DEBUG") != nullptr) fprintf(stderr, "[plan debug] nP %zu trials_ms %.3f\n", pl->P.size(), std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  }
  *out = pl;
  return HOP_OK;
}

// diagnostics: `draws` indices from weights w[0..n) with the planner's table-free replica of std::discrete_distribution<int>, engine
// std::mt19937(seed) -- tests/test_s4pcs_plan.py compares the stream with the standard library's own
int hop_debug_draw_discrete(const float *w, int n, uint32_t seed, int draws, int32_t *out) {
  if (n < 0 || draws < 0 || (n > 0 && !w) || (draws > 0 && !out)) return HOP_EINVAL;
  std::mt19937 engine(seed);
  for (int k = 0; k < draws; ++k) out[k] = Planner::draw_discrete(w, n, engine);
  return HOP_OK;
}

int hop_s4pcs_plan_create(const float *P_xyz, const float *P_nrm, const float *P_prob, int nP, const float *Q_xyz, const float *Q_nrm,
                          int nQ, const int32_t *ppf_keys, int n_keys, const hop_s4pcs_options *opt, hop_s4pcs_plan **out) {
  return plan_create(nullptr, P_xyz, P_nrm, P_prob, nP, Q_xyz, Q_nrm, nQ, ppf_keys, n_keys, opt, out);
}

int hop_s4pcs_plan_create_gpu(hop_ctx *ctx, const float *P_xyz, const float *P_nrm, const float *P_prob, int nP, const float *Q_xyz,
                              const float *Q_nrm, int nQ, const int32_t *ppf_keys, int n_keys, const hop_s4pcs_options *opt, hop_s4pcs_plan **out) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  return plan_create(ctx, P_xyz, P_nrm, P_prob, nP, Q_xyz, Q_nrm, nQ, ppf_keys, n_keys, opt, out);
}

// computePPF.cpp:56-107: the table of the model = the distinct keys of all its point pairs, sorted.  keys_out: capacity x 4 ints
// (may be NULL to ask for the count); *n_keys = the number of distinct keys.
int hop_ppf_table_build(hop_ctx *ctx, const float *xyz, const float *nrm, int n, int32_t *keys_out, int capacity, int32_t *n_keys) {
  HOP_ENTER(ctx);
  if (!ctx || n < 0 || (n > 0 && (!xyz || !nrm)) || !n_keys || capacity < 0 || (capacity > 0 && !keys_out)) { if (ctx) ctx->err = "hop_ppf_table_build: bad arguments"; return HOP_EINVAL; }
  *n_keys = 0;
  if (n < 2) return HOP_OK;
  if (n > 20000) { ctx->err = "hop_ppf_table_build: more than 20000 model points (the table is built from the 5 mm model)"; return HOP_EINVAL; }
  cudaStream_t st = ctx->stream;
  const size_t np = (size_t)n * n;
  std::vector<float4> hp(n), hn(n);
  for (int i = 0; i < n; ++i) {   // Point3D::set_normal normalises once
    V3 nn = normalized(V3{nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]});
    hp[i] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.f); hn[i] = make_float4(nn[0], nn[1], nn[2], 0.f);
  }
  float4 *d_pos = nullptr, *d_nrm = nullptr; uint64_t *d_k = nullptr, *d_s = nullptr, *d_u = nullptr; unsigned char *d_amb = nullptr; int *d_cnt = nullptr; void *d_tmp = nullptr;
  auto release = [&]() { cudaFree(d_pos); cudaFree(d_nrm); cudaFree(d_k); cudaFree(d_s); cudaFree(d_u); cudaFree(d_amb); cudaFree(d_cnt); cudaFree(d_tmp); };
#define PT_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); release(); return HOP_ECUDA; } } while (0)
  PT_CUDA(cudaMalloc(&d_pos, sizeof(float4) * (size_t)n)); PT_CUDA(cudaMalloc(&d_nrm, sizeof(float4) * (size_t)n));
  PT_CUDA(cudaMalloc(&d_k, sizeof(uint64_t) * np)); PT_CUDA(cudaMalloc(&d_s, sizeof(uint64_t) * np)); PT_CUDA(cudaMalloc(&d_u, sizeof(uint64_t) * np));
  PT_CUDA(cudaMalloc(&d_amb, np)); PT_CUDA(cudaMalloc(&d_cnt, sizeof(int)));
  PT_CUDA(cudaMemcpyAsync(d_pos, hp.data(), sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, st));
  PT_CUDA(cudaMemcpyAsync(d_nrm, hn.data(), sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, st));
  ppf_pairs_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(d_pos, d_nrm, n, d_k, d_amb);
  size_t b1 = 0, b2 = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, b1, d_k, d_s, (int)np, 0, 64, st);
  cub::DeviceSelect::Unique(nullptr, b2, d_s, d_u, d_cnt, (int)np, st);
  PT_CUDA(cudaMalloc(&d_tmp, std::max(b1, b2)));
  cub::DeviceRadixSort::SortKeys(d_tmp, b1, d_k, d_s, (int)np, 0, 64, st);
  cub::DeviceSelect::Unique(d_tmp, b2, d_s, d_u, d_cnt, (int)np, st);
  ctx->launches += 3;
  int cnt = 0;
  std::vector<unsigned char> amb(np);
  PT_CUDA(cudaMemcpyAsync(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
  PT_CUDA(cudaMemcpyAsync(amb.data(), d_amb, np, cudaMemcpyDeviceToHost, st));
  PT_CUDA(cudaStreamSynchronize(st));
  std::vector<uint64_t> uniq(cnt);
  if (cnt) PT_CUDA(cudaMemcpy(uniq.data(), d_u, sizeof(uint64_t) * (size_t)cnt, cudaMemcpyDeviceToHost));
  release();
#undef PT_CUDA
  if (!uniq.empty() && uniq.back() == ~0ull) uniq.pop_back();   // the marker of skipped (i >= j), ambiguous and out-of-range pairs
  // the flagged pairs with the reference's libm
  bool added = false;
  for (size_t t = 0; t < np; ++t)
    if (amb[t]) {
      const int i = (int)(t / n), j = (int)(t - (size_t)i * n);
      S4Pt a, b;
      for (int k = 0; k < 3; ++k) { a.p[k] = (&hp[i].x)[k]; b.p[k] = (&hp[j].x)[k]; a.n[k] = (&hn[i].x)[k]; b.n[k] = (&hn[j].x)[k]; }
      int k4[4];
      compute_ppf(a, b, k4);
      if (k4[0] < 0 || k4[0] > 65535 || k4[1] < 0 || k4[1] > 65535 || k4[2] < 0 || k4[2] > 65535 || k4[3] < 0 || k4[3] > 65535) continue;
      uniq.push_back(pack_key(k4));
      added = true;
    }
  if (added) { std::sort(uniq.begin(), uniq.end()); uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end()); }
  *n_keys = (int32_t)uniq.size();
  for (size_t t = 0; t < uniq.size() && (int)t < capacity; ++t) {
    keys_out[4 * t] = (int32_t)(uniq[t] >> 48); keys_out[4 * t + 1] = (int32_t)((uniq[t] >> 32) & 0xffff);
    keys_out[4 * t + 2] = (int32_t)((uniq[t] >> 16) & 0xffff); keys_out[4 * t + 3] = (int32_t)(uniq[t] & 0xffff);
  }
  return HOP_OK;
}

void hop_s4pcs_plan_destroy(hop_s4pcs_plan *plan) { delete plan; }

int hop_s4pcs_plan_sizes(const hop_s4pcs_plan *plan, int32_t *sizes) {
  if (!plan || !sizes) return HOP_EINVAL;
  sizes[0] = (int32_t)plan->P.size(); sizes[1] = (int32_t)plan->Q.size(); sizes[2] = (int32_t)plan->trials.size();
  sizes[3] = (int32_t)(plan->pairs.size() / 2); sizes[4] = (int32_t)(plan->quads.size() / 4); sizes[5] = plan->trials_executed;
  return HOP_OK;
}

int hop_s4pcs_plan_get(const hop_s4pcs_plan *plan, float *Pc, float *Qc, int32_t *q_ids, float *centroids, float *misc, int32_t *trial_i,
                       float *trial_f) {
  if (!plan) return HOP_EINVAL;
  if (Pc) for (size_t i = 0; i < plan->P.size(); ++i) for (int k = 0; k < 3; ++k) Pc[3 * i + k] = plan->P[i].p[k];
  if (Qc) for (size_t i = 0; i < plan->Q.size(); ++i) for (int k = 0; k < 3; ++k) Qc[3 * i + k] = plan->Q[i].p[k];
  if (q_ids) std::memcpy(q_ids, plan->q_ids.data(), sizeof(int32_t) * plan->q_ids.size());
  if (centroids) for (int k = 0; k < 3; ++k) { centroids[k] = plan->centroid_P[k]; centroids[3 + k] = plan->centroid_Q[k]; }
  if (misc) { misc[0] = plan->diameter; misc[1] = plan->ratio; }
  for (size_t t = 0; t < plan->trials.size(); ++t) {
    const S4Trial &tr = plan->trials[t];
    if (trial_i) { trial_i[5 * t] = tr.base_ok; for (int k = 0; k < 4; ++k) trial_i[5 * t + 1 + k] = tr.base[k]; }
    if (trial_f) { trial_f[4 * t] = tr.inv1; trial_f[4 * t + 1] = tr.inv2; trial_f[4 * t + 2] = tr.dist1; trial_f[4 * t + 3] = tr.dist2; }
  }
  return HOP_OK;
}

int hop_s4pcs_plan_intermediates(const hop_s4pcs_plan *plan, int32_t *trial_ranges, int32_t *pairs, int32_t *quads) {
  if (!plan) return HOP_EINVAL;
  if (trial_ranges) std::memcpy(trial_ranges, plan->trial_ranges.data(), sizeof(int32_t) * plan->trial_ranges.size());
  if (pairs) std::memcpy(pairs, plan->pairs.data(), sizeof(int32_t) * plan->pairs.size());
  if (quads) std::memcpy(quads, plan->quads.data(), sizeof(int32_t) * plan->quads.size());
  return HOP_OK;
}

}  // extern "C"
