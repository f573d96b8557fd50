// k_cluster.cu -- greedy pose clustering (non-maximum suppression with object symmetry): host version and device version.
//
//   replaces PoseEstimator<PointT>::clusterPoses       src/perception/src/PoseEstimator.cpp:106-233
//            Utils::rotationGeodesicDistance            src/perception/src/Utils.cpp:29-32
//
// Sequential by definition (a hypothesis is kept when no EARLIER kept cluster is close).  hop_cluster_poses runs it on the
// host like the reference (the Euler angles of every pose are extracted once instead of once per comparison);
// hop_cluster_poses_gpu (bottom of the file) makes the same decisions on the device for the batch sizes where the
// O(hypotheses x clusters) host loop becomes the serial bottleneck of the frame.  Arithmetic follows
// Eigen 3.3's MatrixBase::eulerAngles(2,1,0) and fixed-size 3-term reductions (t0 + (t1 + t2)) so that threshold decisions
// match the reference's; rotationGeodesicDistance keeps the reference's trace(R1 * R2) (not R1^T R2).
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "hop_common.cuh"

namespace {

inline float m(const float *P, int r, int c) { return P[4 * c + r]; }  // column-major 4x4

struct Euler { float r, p, y; };

// R.eulerAngles(2,1,0): odd = 1, i = 2, j = 1, k = 0 (Eigen/src/Geometry/EulerAngles.h)
Euler euler_zyx(const float *P) {
  float res0 = std::atan2(m(P, 1, 0), m(P, 0, 0));
  const float a = m(P, 2, 2), b = m(P, 2, 1);
  const float c2 = std::sqrt(a * a + b * b);
  float res1;
  if (res0 < 0.f) {
    if (res0 > 0.f) res0 -= float(M_PI); else res0 += float(M_PI);
    res1 = std::atan2(-m(P, 2, 0), -c2);
  } else {
    res1 = std::atan2(-m(P, 2, 0), c2);
  }
  const float s1 = std::sin(res0), c1 = std::cos(res0);
  const float res2 = std::atan2(s1 * m(P, 0, 2) - c1 * m(P, 1, 2), c1 * m(P, 1, 1) - s1 * m(P, 0, 1));
  return Euler{res2, res1, res0};  // rpy(2), rpy(1), rpy(0)
}

inline float geodesic(const float *A, const float *B) {  // acos((trace(R1 * R2) - 1) / 2)
  float tr[3];
  for (int i = 0; i < 3; ++i) tr[i] = m(A, i, 0) * m(B, 0, i) + (m(A, i, 1) * m(B, 1, i) + m(A, i, 2) * m(B, 2, i));
  const float trace = tr[0] + (tr[1] + tr[2]);
  return (float)std::acos((trace - 1) / 2.0);
}

}  // namespace

extern "C" int hop_cluster_poses(const float *poses, const float *scores, const int32_t *ids, int n, float angle_diff_deg, float dist_diff,
                                 const float *symmetry_deg, int32_t *keep_out, int32_t *n_keep) {
  if (n < 0 || !n_keep || (n > 0 && (!poses || !scores || !keep_out)) || !symmetry_deg) return HOP_EINVAL;
  *n_keep = 0;
  if (n == 0) return HOP_OK;
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  auto id_of = [&](int k) { return ids ? ids[k] : k; };
  std::sort(order.begin(), order.end(), [&](int a, int b) {  // HypoCompare (:113-121): score descending, id ascending
    if (scores[a] > scores[b]) return true;
    if (scores[a] < scores[b]) return false;
    if (id_of(a) < id_of(b)) return true;
    return false;
  });
  const float radian_thres = angle_diff_deg / 180.0 * M_PI;
  const float sym[3] = {(float)((double)symmetry_deg[0] / 180 * M_PI), (float)((double)symmetry_deg[1] / 180 * M_PI),
                        (float)((double)symmetry_deg[2] / 180 * M_PI)};
  std::vector<Euler> eul(n);
  for (int k = 0; k < n; ++k) eul[k] = euler_zyx(poses + 16 * (size_t)k);
  auto fold = [](float diff, float s) {
    if (s == 0) return 0.f;
    if (s > 0) return std::min(diff, s - diff);
    return diff;
  };
  std::vector<int> clusters;
  clusters.push_back(order[0]);
  for (int i = 1; i < n; ++i) {
    const int cur = order[i];
    const float *P1 = poses + 16 * (size_t)cur;
    bool isnew = true;
    for (int c : clusters) {
      const float *P0 = poses + 16 * (size_t)c;
      const float dx = m(P0, 0, 3) - m(P1, 0, 3), dy = m(P0, 1, 3) - m(P1, 1, 3), dz = m(P0, 2, 3) - m(P1, 2, 3);
      if (std::sqrt(dx * dx + (dy * dy + dz * dz)) >= dist_diff) continue;
      const float roll_diff = fold(std::abs(eul[c].r - eul[cur].r), sym[0]);
      const float pitch_diff = fold(std::abs(eul[c].p - eul[cur].p), sym[1]);
      const float yaw_diff = fold(std::abs(eul[c].y - eul[cur].y), sym[2]);
      if (pitch_diff <= radian_thres && roll_diff <= radian_thres && yaw_diff <= radian_thres) { isnew = false; break; }
      if (geodesic(P0, P1) <= radian_thres) { isnew = false; break; }
    }
    if (isnew) clusters.push_back(cur);
  }
  for (size_t k = 0; k < clusters.size(); ++k) keep_out[k] = clusters[k];
  *n_keep = (int32_t)clusters.size();
  return HOP_OK;
}


// ------------------------------------------------------------------------------------------------------------
// Device version.  The hypotheses are sorted and their Euler angles extracted on the host exactly as above (libm's atan2 /
// sin / cos decide thresholds, and the sort is a few ms even at 65 k); the O(n x clusters) comparisons run on the device
// in sorted blocks of CL_BLOCK hypotheses, the classic two-phase exact greedy suppression:
//   cluster_vs_keepers_kernel : every block member against every cluster kept so far (all SMs);
//   cluster_block_kernel      : the survivors of the block against each other -- a 1024 x 1024 bit matrix in shared memory
//                               built by 1024 threads, then one warp walks it in order and appends the new clusters.
// The pair test repeats the host's float operations one by one (no FMA contraction); the geodesic test
// acos((trace - 1) / 2) <= thr is monotone in its argument, so the host turns it once per call into the smallest double
// x* with (float)acos(x*) <= thr and the device compares against x*: the same decision without a device acos.
// ------------------------------------------------------------------------------------------------------------
namespace {

constexpr int CL_BLOCK = 1024;
constexpr int CL_ROW = 33;     // words per bit-matrix row (32 + 1: conflict-free column walks)
constexpr int CL_KCHUNK = 256; // keepers per CTA of cluster_vs_keepers_kernel

struct ClParams {
  float dist_diff, radian_thres, sym[3];
  double x_star;   // geodesic gate: close when x_star <= x <= 1
};

struct ClFeat { float4 a, b, c, d; };  // a = (tx,ty,tz,roll)  b = (pitch,yaw,R00,R01)  c = (R02,R10,R11,R12)  d = (R20,R21,R22,-)

__device__ __forceinline__ float cl_fold(float diff, float s) {
  if (s == 0.f) return 0.f;
  if (s > 0.f) { const float o = __fsub_rn(s, diff); return o < diff ? o : diff; }   // std::min(diff, s - diff)
  return diff;
}

// A = the earlier (kept) hypothesis, B = the later one: the roles matter for the rounding of trace(R_A * R_B)
__device__ __forceinline__ bool cl_close(const ClFeat &A, const ClFeat &B, const ClParams &p) {
  const float dx = __fsub_rn(A.a.x, B.a.x), dy = __fsub_rn(A.a.y, B.a.y), dz = __fsub_rn(A.a.z, B.a.z);
  const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dz, dz))));
  if (d >= p.dist_diff) return false;
  const float rd = cl_fold(fabsf(__fsub_rn(A.a.w, B.a.w)), p.sym[0]);
  const float pd = cl_fold(fabsf(__fsub_rn(A.b.x, B.b.x)), p.sym[1]);
  const float yd = cl_fold(fabsf(__fsub_rn(A.b.y, B.b.y)), p.sym[2]);
  if (pd <= p.radian_thres && rd <= p.radian_thres && yd <= p.radian_thres) return true;
  // tr_i = A(i,0) B(0,i) + (A(i,1) B(1,i) + A(i,2) B(2,i))
  const float t0 = __fadd_rn(__fmul_rn(A.b.z, B.b.z), __fadd_rn(__fmul_rn(A.b.w, B.c.y), __fmul_rn(A.c.x, B.d.x)));
  const float t1 = __fadd_rn(__fmul_rn(A.c.y, B.b.w), __fadd_rn(__fmul_rn(A.c.z, B.c.z), __fmul_rn(A.c.w, B.d.y)));
  const float t2 = __fadd_rn(__fmul_rn(A.d.x, B.c.x), __fadd_rn(__fmul_rn(A.d.y, B.c.w), __fmul_rn(A.d.z, B.d.z)));
  const float trace = __fadd_rn(t0, __fadd_rn(t1, t2));
  const double x = __ddiv_rn((double)__fsub_rn(trace, 1.f), 2.0);
  return x >= p.x_star && x <= 1.0;   // (x > 1 or NaN: acos is NaN and the host's comparison is false)
}

__global__ void __launch_bounds__(CL_KCHUNK) cluster_vs_keepers_kernel(const ClFeat *__restrict__ feat, int blk_begin, int blk_cnt,
                                                                       const int *__restrict__ keepers, const int *__restrict__ n_keep,
                                                                       unsigned char *__restrict__ suppressed, ClParams p) {
  __shared__ ClFeat s_k[CL_KCHUNK];
  const int K = *n_keep, k0 = blockIdx.x * CL_KCHUNK;
  if (k0 >= K) return;
  const int nk = min(CL_KCHUNK, K - k0);
  if ((int)threadIdx.x < nk) s_k[threadIdx.x] = feat[keepers[k0 + threadIdx.x]];
  __syncthreads();
  for (int j = threadIdx.x; j < blk_cnt; j += CL_KCHUNK) {
    if (suppressed[j]) continue;
    const ClFeat B = feat[blk_begin + j];
    for (int k = 0; k < nk; ++k)
      if (cl_close(s_k[k], B, p)) { suppressed[j] = 1; break; }
  }
}

__global__ void __launch_bounds__(CL_BLOCK, 1) cluster_block_kernel(const ClFeat *__restrict__ feat, int blk_begin, int blk_cnt, int *__restrict__ keepers,
                                                                    int *__restrict__ n_keep, unsigned char *__restrict__ suppressed, ClParams p) {
  extern __shared__ __align__(16) unsigned char cl_smem[];
  ClFeat *s_f = reinterpret_cast<ClFeat *>(cl_smem);
  unsigned int *rows = reinterpret_cast<unsigned int *>(cl_smem + sizeof(ClFeat) * CL_BLOCK);
  __shared__ unsigned char s_sup[CL_BLOCK];
  const int i = threadIdx.x;
  const bool in = i < blk_cnt;
  if (in) s_f[i] = feat[blk_begin + i];
  s_sup[i] = in ? suppressed[i] : 1;
  __syncthreads();
  // row i: which later members of the block hypothesis i suppresses if it is kept
  const bool alive = !s_sup[i];
  ClFeat A;
  if (alive) A = s_f[i];
  for (int w = 0; w < 32; ++w) {
    unsigned int bits = 0u;
    if (alive && 32 * w + 31 > i) {
      for (int b = 0; b < 32; ++b) {
        const int j = 32 * w + b;
        if (j > i && j < blk_cnt && !s_sup[j] && cl_close(A, s_f[j], p)) bits |= 1u << b;
      }
    }
    rows[i * CL_ROW + w] = bits;
  }
  __syncthreads();
  if (i < 32) {
    // the greedy walk: lane l carries word l of the "already suppressed" mask
    unsigned int removed = 0u;
    for (int b = 0; b < 32; ++b) removed |= (unsigned int)(s_sup[32 * i + b] ? 1u : 0u) << b;
    const int K0 = *n_keep;
    int nk = 0;
    for (int k = 0; k < blk_cnt; ++k) {
      const unsigned int word = __shfl_sync(0xffffffffu, removed, k >> 5);
      if (!((word >> (k & 31)) & 1u)) {
        if (i == 0) keepers[K0 + nk] = blk_begin + k;
        ++nk;
        removed |= rows[k * CL_ROW + i];
      }
    }
    if (i == 0) *n_keep = K0 + nk;
  }
}

// Batches of a few thousand hypotheses (a frame's Super4PCS output): the whole upper-triangular "i suppresses j" relation as a bit
// matrix, one thread per (row, word), every SM busy for a few microseconds; the greedy walk over it -- sequential by definition, a few
// thousand steps of "if not removed: keep, OR the row in" -- runs on the host on the downloaded matrix.  Same pair test, same order of
// decisions as the blocked kernels above, which a single CTA walks at ~200 us per 1024 hypotheses.
constexpr int CL_MATRIX_MAX = 4096;   // 4096 x 128 words = 2 MB over PCIe

__global__ void __launch_bounds__(128) cluster_matrix_kernel(const ClFeat *__restrict__ feat, int n, int words, unsigned int *__restrict__ rows, ClParams p) {
  const int i = blockIdx.y;
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= words) return;
  unsigned int bits = 0u;
  if (32 * w + 31 > i) {
    const ClFeat A = feat[i];
    for (int b = 0; b < 32; ++b) {
      const int j = 32 * w + b;
      if (j > i && j < n && cl_close(A, feat[j], p)) bits |= 1u << b;
    }
  }
  rows[(size_t)i * words + w] = bits;
}

// smallest double x in [-1, 1] with (float)acos(x) <= thr (2 when there is none)
double geodesic_gate(float thr) {
  auto ok = [&](double x) { return (float)std::acos(x) <= thr; };
  if (!ok(1.0)) return 2.0;
  if (ok(-1.0)) return -1.0;
  double lo = -1.0, hi = 1.0;   // ok(lo) false, ok(hi) true
  for (int it = 0; it < 200; ++it) {
    const double mid = lo + (hi - lo) / 2;
    if (!(mid > lo && mid < hi)) break;
    if (ok(mid)) hi = mid; else lo = mid;
  }
  return hi;
}

}  // namespace

extern "C" int hop_cluster_poses_gpu(hop_ctx *ctx, const float *poses, const float *scores, const int32_t *ids, int n, float angle_diff_deg,
                                     float dist_diff, const float *symmetry_deg, int32_t *keep_out, int32_t *n_keep) {
  HOP_ENTER(ctx);
  if (!ctx) return HOP_EINVAL;
  if (n < 0 || !n_keep || (n > 0 && (!poses || !scores || !keep_out)) || !symmetry_deg) { ctx->err = "hop_cluster_poses_gpu: bad arguments"; return HOP_EINVAL; }
  *n_keep = 0;
  if (n == 0) return HOP_OK;
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  auto id_of = [&](int k) { return ids ? ids[k] : k; };
  std::sort(order.begin(), order.end(), [&](int a, int b) {
    if (scores[a] > scores[b]) return true;
    if (scores[a] < scores[b]) return false;
    if (id_of(a) < id_of(b)) return true;
    return false;
  });
  ClParams p;
  p.dist_diff = dist_diff;
  p.radian_thres = angle_diff_deg / 180.0 * M_PI;
  for (int k = 0; k < 3; ++k) p.sym[k] = (float)((double)symmetry_deg[k] / 180 * M_PI);
  p.x_star = geodesic_gate(p.radian_thres);
  std::vector<ClFeat> feat(n);
  // (three atan2, a sin / cos and a sqrt per hypothesis with the host's libm: 0.4 ms of a 0.5 ms call at 2.8 k hypotheses on one thread)
#pragma omp parallel for schedule(static) if (n >= 512)
  for (int k = 0; k < n; ++k) {
    const float *P = poses + 16 * (size_t)order[k];
    const Euler e = euler_zyx(P);
    feat[k].a = make_float4(m(P, 0, 3), m(P, 1, 3), m(P, 2, 3), e.r);
    feat[k].b = make_float4(e.p, e.y, m(P, 0, 0), m(P, 0, 1));
    feat[k].c = make_float4(m(P, 0, 2), m(P, 1, 0), m(P, 1, 1), m(P, 1, 2));
    feat[k].d = make_float4(m(P, 2, 0), m(P, 2, 1), m(P, 2, 2), 0.f);
  }
  cudaStream_t st = ctx->stream;
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t feat_bytes = up(sizeof(ClFeat) * (size_t)n), keep_bytes = up(sizeof(int) * ((size_t)n + 1)), sup_bytes = up(CL_BLOCK);
  char *base = (char *)ctx->ensure_work(feat_bytes + keep_bytes + sup_bytes);
  if (!base) { ctx->err = "hop_cluster_poses_gpu: work buffer allocation failed"; return HOP_ENOMEM; }
  ClFeat *d_feat = (ClFeat *)base;
  int *d_keep = (int *)(base + feat_bytes), *d_nkeep = d_keep + n;
  unsigned char *d_sup = (unsigned char *)(base + feat_bytes + keep_bytes);
  HOP_CUDA(ctx, cudaMemcpyAsync(d_feat, feat.data(), sizeof(ClFeat) * (size_t)n, cudaMemcpyHostToDevice, st));
  if (n <= CL_MATRIX_MAX && !ctx->tune.cluster_blocks) {
    const int words = (n + 31) / 32;
    const size_t mat_bytes = sizeof(unsigned int) * (size_t)n * words;
    unsigned int *d_mat = (unsigned int *)ctx->ensure_scratch(mat_bytes);
    unsigned int *h_mat = (unsigned int *)ctx->ensure_pinned(mat_bytes);
    if (!d_mat || !h_mat) { ctx->err = "hop_cluster_poses_gpu: matrix allocation failed"; return HOP_ENOMEM; }
    {
      ProfScope ps(ctx, HOP_PROF_CLUSTER);
      cluster_matrix_kernel<<<dim3((words + 127) / 128, n), 128, 0, st>>>(d_feat, n, words, d_mat, p);
      ctx->launches += 1;
    }
    HOP_CUDA(ctx, cudaGetLastError());
    HOP_CUDA(ctx, cudaMemcpyAsync(h_mat, d_mat, mat_bytes, cudaMemcpyDeviceToHost, st));
    HOP_CUDA(ctx, cudaStreamSynchronize(st));
    std::vector<unsigned int> removed(words, 0u);
    int nk = 0;
    for (int k = 0; k < n; ++k) {
      if ((removed[k >> 5] >> (k & 31)) & 1u) continue;
      keep_out[nk++] = order[k];
      const unsigned int *row = h_mat + (size_t)k * words;
      for (int w = k >> 5; w < words; ++w) removed[w] |= row[w];
    }
    *n_keep = nk;
    return HOP_OK;
  }
  HOP_CUDA(ctx, cudaMemsetAsync(d_nkeep, 0, sizeof(int), st));
  const size_t smem = sizeof(ClFeat) * CL_BLOCK + sizeof(unsigned int) * CL_BLOCK * CL_ROW;
  HOP_CUDA(ctx, ctx->func_smem_optin(cluster_block_kernel, smem));
  {
    ProfScope ps(ctx, HOP_PROF_CLUSTER);
    for (int b0 = 0; b0 < n; b0 += CL_BLOCK) {
      const int cnt = std::min(CL_BLOCK, n - b0);
      HOP_CUDA(ctx, cudaMemsetAsync(d_sup, 0, CL_BLOCK, st));
      if (b0 > 0) {
        cluster_vs_keepers_kernel<<<(b0 + CL_KCHUNK - 1) / CL_KCHUNK, CL_KCHUNK, 0, st>>>(d_feat, b0, cnt, d_keep, d_nkeep, d_sup, p);
        ctx->launches += 1;
      }
      cluster_block_kernel<<<1, CL_BLOCK, smem, st>>>(d_feat, b0, cnt, d_keep, d_nkeep, d_sup, p);
      ctx->launches += 1;
    }
  }
  HOP_CUDA(ctx, cudaGetLastError());
  int nk = 0;
  HOP_CUDA(ctx, cudaMemcpyAsync(&nk, d_nkeep, sizeof(int), cudaMemcpyDeviceToHost, st));
  HOP_CUDA(ctx, cudaStreamSynchronize(st));
  std::vector<int> kept(nk);
  if (nk > 0) {
    HOP_CUDA(ctx, cudaMemcpyAsync(kept.data(), d_keep, sizeof(int) * (size_t)nk, cudaMemcpyDeviceToHost, st));
    HOP_CUDA(ctx, cudaStreamSynchronize(st));
  }
  for (int k = 0; k < nk; ++k) keep_out[k] = order[kept[k]];
  *n_keep = nk;
  return HOP_OK;
}
