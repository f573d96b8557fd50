/*
 * hop_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, never shipped, never on the product path).
 *
 * Plain-C restatement of the reference's per-hypothesis ICP refinement and LCP scoring:
 *   Utils::runICP<PointT>        /root/reference/src/perception/src/Utils.cpp:188-229
 *   Utils::computeLCP<PointT>    /root/reference/src/perception/src/Utils.cpp:372-444
 *   PoseEstimator::refineByICP   /root/reference/src/perception/src/PoseEstimator.cpp:235-275
 *   PoseEstimator::selectBest    /root/reference/src/perception/src/PoseEstimator.cpp:465-502
 *
 * The arithmetic of runICP lives in PCL 1.9 (find_package(PCL 1.9), src/perception/CMakeLists.txt:18), which
 * is NOT vendored in /root/reference and is not installed here.  Its published algorithm is restated below:
 *   pcl::IterativeClosestPoint::computeTransformation   (registration/impl/icp.hpp)
 *   pcl::registration::CorrespondenceEstimation         (exact 1-NN, d^2 <= max_dist^2)
 *   pcl::registration::CorrespondenceRejectorSurfaceNormal (n_src . n_tgt > cos(angle), source normals rotated)
 *   pcl::registration::TransformationEstimationPointToPlane -> TransformationEstimationLM
 *        = Eigen::LevenbergMarquardt<Eigen::NumericalDiff<functor>, float> over WarpPointRigid6D
 *   pcl::registration::DefaultConvergenceCriteria       (abs MSE 1e-6, max iterations)
 * The Levenberg-Marquardt driver (MINPACK lmder/lmpar/qrsolv as carried by Eigen's unsupported
 * NonLinearOptimization module) IS present in the reference tree (src/OpenGR_4pcs/3rdparty/Eigen/unsupported) and
 * oracle/_ref (built by oracle/Makefile from those sources where they lie) is used by tests/ to pin the lm_*
 * functions below step for step.
 *
 * PARITY PINNING: the reference holds no golden vectors for this path and PCL cannot be run
 * here, so runICP/computeLCP are "parity unpinned" against PCL itself; the LM solver is pinned against the
 * reference tree's own Eigen, the NN search against brute force, and LCP against scipy cKDTree in tests/.
 *
 * All cloud arithmetic is float (PCL Scalar=float), MSE in double, like the reference.
 * Matrices crossing this API are 4x4 COLUMN-major floats (Eigen::Matrix4f::data()).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define M4(m, r, c) ((m)[(c) * 4 + (r)])

/* ------------------------------------------------------------------------------------------------
 * Exact kd-tree (stand-in for pcl::KdTreeFLANN = FLANN KDTreeSingleIndex, leaf 15, exact search).
 * Any exact 1-NN structure returns the same neighbour up to exact distance ties.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int left, right; /* children (node ids) or -1 */
  int lo, hi;      /* index range [lo,hi) into perm for leaves */
  int dim;
  float split;
  float bmin[3], bmax[3];
} kd_node;

typedef struct {
  const float *xyz; /* N x 3, not owned */
  int n;
  int *perm;
  kd_node *nodes;
  int n_nodes, cap_nodes;
} kd_tree;

#define KD_LEAF 15

static void kd_swap(int *a, int *b) { int t = *a; *a = *b; *b = t; }

static void kd_select(const float *xyz, int *perm, int lo, int hi, int k, int dim) {
  /* quickselect so that perm[k] holds the k-th smallest along dim within [lo,hi) */
  while (hi - lo > 1) {
    int mid = lo + (hi - lo) / 2;
    float pv = xyz[3 * perm[mid] + dim];
    kd_swap(&perm[mid], &perm[hi - 1]);
    int st = lo;
    for (int i = lo; i < hi - 1; ++i)
      if (xyz[3 * perm[i] + dim] < pv) { kd_swap(&perm[i], &perm[st]); ++st; }
    kd_swap(&perm[st], &perm[hi - 1]);
    if (st == k) return;
    if (k < st) hi = st; else lo = st + 1;
  }
}

static int kd_new_node(kd_tree *t) {
  if (t->n_nodes == t->cap_nodes) {
    t->cap_nodes = t->cap_nodes ? 2 * t->cap_nodes : 64;
    t->nodes = (kd_node *)realloc(t->nodes, sizeof(kd_node) * (size_t)t->cap_nodes);
  }
  return t->n_nodes++;
}

static int kd_build_rec(kd_tree *t, int lo, int hi) {
  int id = kd_new_node(t);
  kd_node nd;
  nd.left = nd.right = -1; nd.lo = lo; nd.hi = hi; nd.dim = 0; nd.split = 0.f;
  for (int d = 0; d < 3; ++d) { nd.bmin[d] = FLT_MAX; nd.bmax[d] = -FLT_MAX; }
  for (int i = lo; i < hi; ++i)
    for (int d = 0; d < 3; ++d) {
      float v = t->xyz[3 * t->perm[i] + d];
      if (v < nd.bmin[d]) nd.bmin[d] = v;
      if (v > nd.bmax[d]) nd.bmax[d] = v;
    }
  if (hi - lo > KD_LEAF) {
    int dim = 0; float ext = nd.bmax[0] - nd.bmin[0];
    for (int d = 1; d < 3; ++d) if (nd.bmax[d] - nd.bmin[d] > ext) { ext = nd.bmax[d] - nd.bmin[d]; dim = d; }
    if (ext > 0.f) {
      int mid = lo + (hi - lo) / 2;
      kd_select(t->xyz, t->perm, lo, hi, mid, dim);
      nd.dim = dim; nd.split = t->xyz[3 * t->perm[mid] + dim];
      t->nodes[id] = nd;
      int l = kd_build_rec(t, lo, mid);
      int r = kd_build_rec(t, mid, hi);
      nd.left = l; nd.right = r;
    }
  }
  t->nodes[id] = nd;
  return id;
}

static kd_tree *kd_build(const float *xyz, int n) {
  kd_tree *t = (kd_tree *)calloc(1, sizeof(kd_tree));
  t->xyz = xyz; t->n = n;
  t->perm = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i) t->perm[i] = i;
  if (n > 0) kd_build_rec(t, 0, n);
  return t;
}

static void kd_free(kd_tree *t) { if (!t) return; free(t->perm); free(t->nodes); free(t); }

static inline float kd_box_d2(const kd_node *nd, const float q[3]) {
  float s = 0.f;
  for (int d = 0; d < 3; ++d) {
    float v = 0.f;
    if (q[d] < nd->bmin[d]) v = nd->bmin[d] - q[d]; else if (q[d] > nd->bmax[d]) v = q[d] - nd->bmax[d];
    s += v * v;
  }
  return s;
}

static void kd_search(const kd_tree *t, int id, const float q[3], int *best, float *best_d2) {
  const kd_node *nd = &t->nodes[id];
  if (nd->left < 0) {
    for (int i = nd->lo; i < nd->hi; ++i) {
      int p = t->perm[i];
      float dx = t->xyz[3 * p] - q[0], dy = t->xyz[3 * p + 1] - q[1], dz = t->xyz[3 * p + 2] - q[2];
      float d2 = dx * dx + dy * dy + dz * dz;
      if (d2 < *best_d2 || (d2 == *best_d2 && p < *best)) { *best_d2 = d2; *best = p; }
    }
    return;
  }
  int first = q[nd->dim] < nd->split ? nd->left : nd->right;
  int second = first == nd->left ? nd->right : nd->left;
  if (kd_box_d2(&t->nodes[first], q) <= *best_d2) kd_search(t, first, q, best, best_d2);
  if (kd_box_d2(&t->nodes[second], q) <= *best_d2) kd_search(t, second, q, best, best_d2);
}

/* nearestKSearch(pt, 1, ...): returns index or -1 when the tree is empty; *d2 = squared distance (float) */
static int kd_nn(const kd_tree *t, const float q[3], float *d2) {
  if (t->n <= 0) return -1;
  int best = -1; float bd = FLT_MAX;
  kd_search(t, 0, q, &best, &bd);
  *d2 = bd;
  return best;
}

/* exported for tests: brute force and kd-tree 1-NN of every query (validates kd_nn) */
void hop_oracle_nn(const float *pts, int n, const float *q, int nq, int use_kdtree, int *idx, float *d2) {
  kd_tree *t = use_kdtree ? kd_build(pts, n) : NULL;
  for (int k = 0; k < nq; ++k) {
    if (use_kdtree) { idx[k] = kd_nn(t, q + 3 * k, d2 + k); continue; }
    int b = -1; float bd = FLT_MAX;
    for (int i = 0; i < n; ++i) {
      float dx = pts[3 * i] - q[3 * k], dy = pts[3 * i + 1] - q[3 * k + 1], dz = pts[3 * i + 2] - q[3 * k + 2];
      float d = dx * dx + dy * dy + dz * dz;
      if (d < bd) { bd = d; b = i; }
    }
    idx[k] = b; d2[k] = bd;
  }
  kd_free(t);
}

/* ------------------------------------------------------------------------------------------------
 * pcl::transformPointCloudWithNormals (float 4x4): p' = R p + t, n' = R n.
 * Call sites: PoseEstimator.cpp:263 (model -> model_4pcs), :487 (model001 -> transformed_model).
 * ---------------------------------------------------------------------------------------------- */
void hop_oracle_transform_cloud(const float *T, const float *xyz, const float *nrm, int n, float *oxyz, float *onrm) {
  for (int i = 0; i < n; ++i) {
    float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    oxyz[3 * i + 0] = M4(T, 0, 0) * x + M4(T, 0, 1) * y + M4(T, 0, 2) * z + M4(T, 0, 3);
    oxyz[3 * i + 1] = M4(T, 1, 0) * x + M4(T, 1, 1) * y + M4(T, 1, 2) * z + M4(T, 1, 3);
    oxyz[3 * i + 2] = M4(T, 2, 0) * x + M4(T, 2, 1) * y + M4(T, 2, 2) * z + M4(T, 2, 3);
    if (nrm) {
      float a = nrm[3 * i], b = nrm[3 * i + 1], c = nrm[3 * i + 2];
      onrm[3 * i + 0] = M4(T, 0, 0) * a + M4(T, 0, 1) * b + M4(T, 0, 2) * c;
      onrm[3 * i + 1] = M4(T, 1, 0) * a + M4(T, 1, 1) * b + M4(T, 1, 2) * c;
      onrm[3 * i + 2] = M4(T, 2, 0) * a + M4(T, 2, 1) * b + M4(T, 2, 2) * c;
    }
  }
}

static void m4_identity(float *T) { memset(T, 0, 16 * sizeof(float)); T[0] = T[5] = T[10] = T[15] = 1.f; }

static void m4_mul(const float *A, const float *B, float *C) { /* C = A*B, column-major, float like Eigen */
  float R[16];
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      float s = 0.f;
      for (int k = 0; k < 4; ++k) s += M4(A, r, k) * M4(B, k, c);
      M4(R, r, c) = s;
    }
  memcpy(C, R, sizeof(R));
}

/* general 4x4 inverse in double (Eigen::Matrix4f::inverse() is a general cofactor inverse; PoseEstimator.cpp:267) */
static int m4_inverse(const float *Tf, float *out) {
  double a[4][8];
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) { a[r][c] = M4(Tf, r, c); a[r][c + 4] = (r == c); }
  for (int i = 0; i < 4; ++i) {
    int p = i; for (int r = i + 1; r < 4; ++r) if (fabs(a[r][i]) > fabs(a[p][i])) p = r;
    if (a[p][i] == 0.0) return -1;
    if (p != i) for (int c = 0; c < 8; ++c) { double t = a[i][c]; a[i][c] = a[p][c]; a[p][c] = t; }
    double inv = 1.0 / a[i][i];
    for (int c = 0; c < 8; ++c) a[i][c] *= inv;
    for (int r = 0; r < 4; ++r) if (r != i) { double f = a[r][i]; if (f != 0.0) for (int c = 0; c < 8; ++c) a[r][c] -= f * a[i][c]; }
  }
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) M4(out, r, c) = (float)a[r][c + 4];
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Utils::computeLCP  (Utils.cpp:372-444).  scene/model: N x 3 xyz + N x 3 normals, model ALREADY transformed.
 * ---------------------------------------------------------------------------------------------- */
static inline void v3_normalize(float *v) { /* Eigen normalize(): v /= sqrt(v.v) when norm > 0 */
  float z = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  if (z > 0.f) { float s = sqrtf(z); v[0] /= s; v[1] /= s; v[2] /= s; }
}

float hop_oracle_compute_lcp(const float *s_xyz, const float *s_nrm, int ns, const float *weights,
                             const float *m_xyz, const float *m_nrm, int nm, float dist_thres, float angle_thres,
                             int use_normal, int use_dot_score, int use_reciprocal) {
  float cp = 0.f;
  kd_tree *kd_model = kd_build(m_xyz, nm);  /* Utils.cpp:376-377 */
  kd_tree *kd_scene = kd_build(s_xyz, ns);  /* Utils.cpp:378-379 */
  const float cos_thres = (float)cos(angle_thres / 180.0 * M_PI); /* :384 */
  for (int i = 0; i < ns; ++i) {
    float d2 = 0.f;
    int j = kd_nn(kd_model, s_xyz + 3 * i, &d2);
    if (j < 0 || !(d2 < dist_thres * dist_thres)) continue; /* :388 */
    const float w = weights ? weights[i] : 1.f;
    if (!use_normal) cp += w;
    else {
      float n1[3] = {s_nrm[3 * i], s_nrm[3 * i + 1], s_nrm[3 * i + 2]};
      float n2[3] = {m_nrm[3 * j], m_nrm[3 * j + 1], m_nrm[3 * j + 2]};
      v3_normalize(n1); v3_normalize(n2);
      float dot = n1[0] * n2[0] + n1[1] * n2[1] + n1[2] * n2[2];
      if (dot > cos_thres) {
        if (!use_dot_score) cp += w;
        else cp += dot * (1.f - sqrtf(d2) / dist_thres) * w; /* :408 */
      }
    }
    if (use_reciprocal) { /* :412-440 : model neighbour back into the scene, no radius */
      float e2 = 0.f;
      int k = kd_nn(kd_scene, m_xyz + 3 * j, &e2);
      if (k < 0) continue;
      if (!use_normal) cp += w;
      else {
        float n1[3] = {m_nrm[3 * j], m_nrm[3 * j + 1], m_nrm[3 * j + 2]};
        float n2[3] = {s_nrm[3 * k], s_nrm[3 * k + 1], s_nrm[3 * k + 2]};
        v3_normalize(n1); v3_normalize(n2);
        float dot = n1[0] * n2[0] + n1[1] * n2[1] + n1[2] * n2[2];
        if (dot > cos_thres) {
          if (!use_dot_score) cp += w;
          else cp += dot * (1.f - sqrtf(e2) / dist_thres) * w; /* :433 (weight of scene point i, as written) */
        }
      }
    }
  }
  kd_free(kd_model); kd_free(kd_scene);
  return cp;
}

/* ------------------------------------------------------------------------------------------------
 * Levenberg-Marquardt (MINPACK lmder as driven by Eigen::LevenbergMarquardt<NumericalDiff<F>,float>::minimize
 * with default Parameters: factor 100, maxfev 400, ftol = xtol = sqrt(eps), gtol 0, epsfcn 0) on the
 * point-to-plane residual of pcl::registration::TransformationEstimationPointToPlane:
 *     f_k(x) = (W(x) s_k - t_k) . n_k ,   W = WarpPointRigid6D: t = x[0..2], q = (sqrt(1-|x[3..5]|^2), x[3..5]).normalized
 * Numerical differentiation: forward, h = sqrt(eps)*|x_j| (h = sqrt(eps) when 0)   (Eigen NumericalDiff, Forward).
 * ---------------------------------------------------------------------------------------------- */
#define LM_N 6
static const float LM_EPS = FLT_EPSILON;

typedef struct {
  const float *src; /* m x 3 (already-transformed source points) */
  const float *tgt; /* m x 3 */
  const float *nrm; /* m x 3 target normals */
  int m;
} lm_problem;

void hop_oracle_warp6d(const float *x, float *T) { /* pcl::registration::WarpPointRigid6D::setParam */
  float qx = x[3], qy = x[4], qz = x[5];
  float qw = sqrtf(1.f - (qx * qx + qy * qy + qz * qz)); /* q.w() = sqrt(1 - q.dot(q)) with w=0 beforehand */
  float nn = sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
  qw /= nn; qx /= nn; qy /= nn; qz /= nn;
  /* Eigen::Quaternion::toRotationMatrix */
  float tx = 2.f * qx, ty = 2.f * qy, tz = 2.f * qz;
  float twx = tx * qw, twy = ty * qw, twz = tz * qw;
  float txx = tx * qx, txy = ty * qx, txz = tz * qx;
  float tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  memset(T, 0, 16 * sizeof(float));
  M4(T, 0, 0) = 1.f - (tyy + tzz); M4(T, 0, 1) = txy - twz;         M4(T, 0, 2) = txz + twy;
  M4(T, 1, 0) = txy + twz;         M4(T, 1, 1) = 1.f - (txx + tzz); M4(T, 1, 2) = tyz - twx;
  M4(T, 2, 0) = txz - twy;         M4(T, 2, 1) = tyz + twx;         M4(T, 2, 2) = 1.f - (txx + tyy);
  M4(T, 0, 3) = x[0]; M4(T, 1, 3) = x[1]; M4(T, 2, 3) = x[2]; M4(T, 3, 3) = 1.f;
}

static void lm_residuals(const lm_problem *P, const float *x, float *f) {
  float T[16];
  hop_oracle_warp6d(x, T);
  for (int k = 0; k < P->m; ++k) {
    const float *s = P->src + 3 * k, *t = P->tgt + 3 * k, *n = P->nrm + 3 * k;
    float wx = M4(T, 0, 0) * s[0] + M4(T, 0, 1) * s[1] + M4(T, 0, 2) * s[2] + M4(T, 0, 3);
    float wy = M4(T, 1, 0) * s[0] + M4(T, 1, 1) * s[1] + M4(T, 1, 2) * s[2] + M4(T, 1, 3);
    float wz = M4(T, 2, 0) * s[0] + M4(T, 2, 1) * s[1] + M4(T, 2, 2) * s[2] + M4(T, 2, 3);
    f[k] = (wx - t[0]) * n[0] + (wy - t[1]) * n[1] + (wz - t[2]) * n[2];
  }
}

static float lm_norm(const float *v, int n) { /* stableNorm/blueNorm: scale-safe 2-norm */
  double s = 0.0; float mx = 0.f;
  for (int i = 0; i < n; ++i) { float a = fabsf(v[i]); if (a > mx) mx = a; }
  if (mx == 0.f || !isfinite(mx)) return mx;
  for (int i = 0; i < n; ++i) { double a = v[i] / mx; s += a * a; }
  return (float)(mx * sqrt(s));
}

/* Householder QR with column pivoting of the m x 6 Jacobian (column-major, ld = m).  On exit: R in the upper
 * triangle of the first 6 rows, qtf = first 6 entries of Q^T f, perm[j] = original column at position j. */
static void lm_qr(float *J, int m, float *f_work, int *perm, float *qtf) {
  float cn[LM_N];
  for (int j = 0; j < LM_N; ++j) { perm[j] = j; cn[j] = lm_norm(J + (size_t)j * m, m); }
  for (int j = 0; j < LM_N && j < m; ++j) {
    int kmax = j; float best = -1.f;
    for (int k = j; k < LM_N; ++k) { float nr = lm_norm(J + (size_t)k * m + j, m - j); if (nr > best) { best = nr; kmax = k; } }
    if (kmax != j) {
      for (int i = 0; i < m; ++i) { float t = J[(size_t)j * m + i]; J[(size_t)j * m + i] = J[(size_t)kmax * m + i]; J[(size_t)kmax * m + i] = t; }
      int tp = perm[j]; perm[j] = perm[kmax]; perm[kmax] = tp;
    }
    float *col = J + (size_t)j * m;
    float alpha = lm_norm(col + j, m - j);
    if (alpha == 0.f) continue;
    if (col[j] > 0.f) alpha = -alpha; /* R(j,j) = -sign(x0)*|x| */
    /* v = x - alpha e1 ; H = I - 2 v v^T/(v^T v) */
    float v0 = col[j] - alpha;
    double vtv = (double)v0 * v0;
    for (int i = j + 1; i < m; ++i) vtv += (double)col[i] * col[i];
    if (vtv == 0.0) continue;
    for (int k = j + 1; k < LM_N; ++k) {
      float *ck = J + (size_t)k * m;
      double dot = (double)v0 * ck[j];
      for (int i = j + 1; i < m; ++i) dot += (double)col[i] * ck[i];
      float sc = (float)(2.0 * dot / vtv);
      ck[j] -= sc * v0;
      for (int i = j + 1; i < m; ++i) ck[i] -= sc * col[i];
    }
    {
      double dot = (double)v0 * f_work[j];
      for (int i = j + 1; i < m; ++i) dot += (double)col[i] * f_work[i];
      float sc = (float)(2.0 * dot / vtv);
      f_work[j] -= sc * v0;
      for (int i = j + 1; i < m; ++i) f_work[i] -= sc * col[i];
    }
    col[j] = alpha;
    for (int i = j + 1; i < m; ++i) col[i] = 0.f;
  }
  (void)cn;
  for (int j = 0; j < LM_N; ++j) qtf[j] = j < m ? f_work[j] : 0.f;
}

/* MINPACK qrsolv on the 6x6 R (row-major r[i][j], upper) : solve min |R P^T x - qtb|^2 + |D x|^2 */
static void lm_qrsolv(float r[LM_N][LM_N], const int *ipvt, const float *diag, const float *qtb, float *x, float *sdiag) {
  float s[LM_N][LM_N], wa[LM_N];
  for (int i = 0; i < LM_N; ++i) for (int j = 0; j < LM_N; ++j) s[i][j] = (j >= i) ? r[i][j] : 0.f;
  /* copy upper to lower as MINPACK does (work on transposed storage) */
  for (int j = 0; j < LM_N; ++j) { for (int i = j; i < LM_N; ++i) s[i][j] = r[j][i]; x[j] = r[j][j]; wa[j] = qtb[j]; }
  for (int j = 0; j < LM_N; ++j) {
    int l = ipvt[j];
    if (diag[l] != 0.f) {
      for (int k = j; k < LM_N; ++k) sdiag[k] = 0.f;
      sdiag[j] = diag[l];
      float qtbpj = 0.f;
      for (int k = j; k < LM_N; ++k) {
        if (sdiag[k] == 0.f) continue;
        float sn, cs;
        if (fabsf(s[k][k]) < fabsf(sdiag[k])) { float ct = s[k][k] / sdiag[k]; sn = 0.5f / sqrtf(0.25f + 0.25f * ct * ct); cs = sn * ct; }
        else { float tn = sdiag[k] / s[k][k]; cs = 0.5f / sqrtf(0.25f + 0.25f * tn * tn); sn = cs * tn; }
        s[k][k] = cs * s[k][k] + sn * sdiag[k];
        float tmp = cs * wa[k] + sn * qtbpj;
        qtbpj = -sn * wa[k] + cs * qtbpj;
        wa[k] = tmp;
        for (int i = k + 1; i < LM_N; ++i) {
          float t2 = cs * s[i][k] + sn * sdiag[i];
          sdiag[i] = -sn * s[i][k] + cs * sdiag[i];
          s[i][k] = t2;
        }
      }
    }
    sdiag[j] = s[j][j];
    s[j][j] = x[j];
  }
  int nsing = LM_N;
  for (int j = 0; j < LM_N; ++j) { if (sdiag[j] == 0.f && nsing == LM_N) nsing = j; if (nsing < LM_N) wa[j] = 0.f; }
  for (int k = 0; k < nsing; ++k) {
    int j = nsing - 1 - k;
    float sum = 0.f;
    for (int i = j + 1; i < nsing; ++i) sum += s[i][j] * wa[i];
    wa[j] = (wa[j] - sum) / sdiag[j];
  }
  for (int j = 0; j < LM_N; ++j) x[ipvt[j]] = wa[j];
  /* hand the lower-triangular factor back for lmpar's Newton correction */
  for (int i = 0; i < LM_N; ++i) for (int j = 0; j < i; ++j) r[i][j] = s[i][j];
}

static void lm_lmpar(float r[LM_N][LM_N], const int *ipvt, const float *diag, const float *qtb, float delta, float *par, float *x) {
  const float dwarf = FLT_MIN;
  float wa1[LM_N], wa2[LM_N], sdiag[LM_N];
  int nsing = LM_N;
  for (int j = 0; j < LM_N; ++j) { wa1[j] = qtb[j]; if (r[j][j] == 0.f && nsing == LM_N) nsing = j; if (nsing < LM_N) wa1[j] = 0.f; }
  for (int k = 0; k < nsing; ++k) {
    int j = nsing - 1 - k;
    wa1[j] /= r[j][j];
    float t = wa1[j];
    for (int i = 0; i < j; ++i) wa1[i] -= r[i][j] * t;
  }
  for (int j = 0; j < LM_N; ++j) x[ipvt[j]] = wa1[j];
  int iter = 0;
  for (int j = 0; j < LM_N; ++j) wa2[j] = diag[j] * x[j];
  float dxnorm = lm_norm(wa2, LM_N);
  float fp = dxnorm - delta;
  if (fp <= 0.1f * delta) { *par = 0.f; return; }
  float parl = 0.f;
  if (nsing >= LM_N) {
    for (int j = 0; j < LM_N; ++j) { int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
    for (int j = 0; j < LM_N; ++j) {
      float sum = 0.f;
      for (int i = 0; i < j; ++i) sum += r[i][j] * wa1[i];
      wa1[j] = (wa1[j] - sum) / r[j][j];
    }
    float t = lm_norm(wa1, LM_N);
    parl = fp / delta / t / t;
  }
  for (int j = 0; j < LM_N; ++j) {
    float sum = 0.f;
    for (int i = 0; i <= j; ++i) sum += r[i][j] * qtb[i];
    wa1[j] = sum / diag[ipvt[j]];
  }
  float gnorm = lm_norm(wa1, LM_N);
  float paru = gnorm / delta;
  if (paru == 0.f) paru = dwarf / fminf(delta, 0.1f);
  *par = fmaxf(*par, parl); *par = fminf(*par, paru);
  if (*par == 0.f) *par = gnorm / dxnorm;
  for (;;) {
    ++iter;
    if (*par == 0.f) *par = fmaxf(dwarf, 0.001f * paru);
    float sq = sqrtf(*par);
    for (int j = 0; j < LM_N; ++j) wa1[j] = sq * diag[j];
    float rr[LM_N][LM_N];
    memcpy(rr, r, sizeof(rr));
    lm_qrsolv(rr, ipvt, wa1, qtb, x, sdiag);
    for (int j = 0; j < LM_N; ++j) wa2[j] = diag[j] * x[j];
    dxnorm = lm_norm(wa2, LM_N);
    float temp = fp;
    fp = dxnorm - delta;
    if (fabsf(fp) <= 0.1f * delta || (parl == 0.f && fp <= temp && temp < 0.f) || iter == 10) break;
    for (int j = 0; j < LM_N; ++j) { int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
    for (int j = 0; j < LM_N; ++j) {
      wa1[j] /= sdiag[j];
      float t = wa1[j];
      for (int i = j + 1; i < LM_N; ++i) wa1[i] -= rr[i][j] * t;
    }
    temp = lm_norm(wa1, LM_N);
    float parc = fp / delta / temp / temp;
    if (fp > 0.f) parl = fmaxf(parl, *par);
    if (fp < 0.f) paru = fminf(paru, *par);
    *par = fmaxf(parl, *par + parc);
  }
  if (iter == 0) *par = 0.f;
}

/* returns Eigen's LevenbergMarquardtSpace::Status code; x (6) in/out (PCL starts from 0) */
int hop_oracle_lm_point_to_plane(const float *src, const float *tgt, const float *nrm, int m, float *x, int *nfev_out) {
  if (m < LM_N) { if (nfev_out) *nfev_out = 0; return 0; } /* minimizeInit: m < n -> ImproperInputParameters, x untouched */
  lm_problem P = {src, tgt, nrm, m};
  const float factor = 100.f, ftol = sqrtf(LM_EPS), xtol = sqrtf(LM_EPS), gtol = 0.f;
  const int maxfev = 400;
  const float h_eps = sqrtf(LM_EPS); /* sqrt(max(epsfcn=0, eps)) */
  float *fvec = (float *)malloc(sizeof(float) * (size_t)m * 3);
  float *wa4 = fvec + m, *fw = fvec + 2 * (size_t)m;
  float *J = (float *)malloc(sizeof(float) * (size_t)m * LM_N);
  float diag[LM_N], qtf[LM_N], wa1[LM_N], wa2[LM_N], wa3[LM_N];
  int perm[LM_N];
  int nfev = 0, iter = 1, status = 0;
  float par = 0.f, delta = 0.f, xnorm = 0.f, fnorm, gnorm;
  lm_residuals(&P, x, fvec); nfev = 1;
  fnorm = lm_norm(fvec, m);
  for (;;) {
    /* jacobian by forward differences (Eigen::NumericalDiff::df); its n evaluations count as nfev */
    for (int j = 0; j < LM_N; ++j) {
      float h = h_eps * fabsf(x[j]);
      if (h == 0.f) h = h_eps;
      float xs = x[j];
      x[j] += h;
      lm_residuals(&P, x, wa4);
      x[j] = xs;
      for (int i = 0; i < m; ++i) J[(size_t)j * m + i] = (wa4[i] - fvec[i]) / h;
    }
    nfev += LM_N + 1; /* NumericalDiff::df (Forward) re-evaluates f(x) first: n+1 evaluations */
    for (int j = 0; j < LM_N; ++j) wa2[j] = lm_norm(J + (size_t)j * m, m);
    memcpy(fw, fvec, sizeof(float) * (size_t)m);
    lm_qr(J, m, fw, perm, qtf);
    float r[LM_N][LM_N];
    for (int i = 0; i < LM_N; ++i) for (int j = 0; j < LM_N; ++j) r[i][j] = (j >= i && i < m) ? J[(size_t)j * m + i] : 0.f;
    if (iter == 1) {
      for (int j = 0; j < LM_N; ++j) diag[j] = (wa2[j] == 0.f) ? 1.f : wa2[j];
      for (int j = 0; j < LM_N; ++j) wa3[j] = diag[j] * x[j];
      xnorm = lm_norm(wa3, LM_N);
      delta = factor * xnorm;
      if (delta == 0.f) delta = factor;
    }
    gnorm = 0.f;
    if (fnorm != 0.f)
      for (int j = 0; j < LM_N; ++j)
        if (wa2[perm[j]] != 0.f) {
          float s = 0.f;
          for (int i = 0; i <= j; ++i) s += r[i][j] * (qtf[i] / fnorm);
          gnorm = fmaxf(gnorm, fabsf(s / wa2[perm[j]]));
        }
    if (gnorm <= gtol) { status = 4; break; } /* CosinusTooSmall */
    for (int j = 0; j < LM_N; ++j) diag[j] = fmaxf(diag[j], wa2[j]);
    float ratio;
    int done = 0;
    do {
      lm_lmpar(r, perm, diag, qtf, delta, &par, wa1);
      for (int j = 0; j < LM_N; ++j) { wa1[j] = -wa1[j]; wa2[j] = x[j] + wa1[j]; wa3[j] = diag[j] * wa1[j]; }
      float pnorm = lm_norm(wa3, LM_N);
      if (iter == 1) delta = fminf(delta, pnorm);
      lm_residuals(&P, wa2, wa4); ++nfev;
      float fnorm1 = lm_norm(wa4, m);
      float actred = -1.f;
      if (0.1f * fnorm1 < fnorm) { float q = fnorm1 / fnorm; actred = 1.f - q * q; }
      /* wa3 = R * P^T * p */
      for (int i = 0; i < LM_N; ++i) { float s = 0.f; for (int j = i; j < LM_N; ++j) s += r[i][j] * wa1[perm[j]]; wa3[i] = s; }
      float t1 = lm_norm(wa3, LM_N) / fnorm; t1 *= t1;
      float t2 = sqrtf(par) * pnorm / fnorm; t2 *= t2;
      float prered = t1 + t2 / 0.5f;
      float dirder = -(t1 + t2);
      ratio = 0.f;
      if (prered != 0.f) ratio = actred / prered;
      if (ratio <= 0.25f) {
        float temp = 0.5f;
        if (actred < 0.f) temp = 0.5f * dirder / (dirder + 0.5f * actred);
        if (0.1f * fnorm1 >= fnorm || temp < 0.1f) temp = 0.1f;
        delta = temp * fminf(delta, pnorm / 0.1f);
        par /= temp;
      } else if (!(par != 0.f && ratio < 0.75f)) {
        delta = pnorm / 0.5f;
        par = 0.5f * par;
      }
      if (ratio >= 1e-4f) {
        for (int j = 0; j < LM_N; ++j) { x[j] = wa2[j]; wa2[j] = diag[j] * x[j]; }
        memcpy(fvec, wa4, sizeof(float) * (size_t)m);
        xnorm = lm_norm(wa2, LM_N);
        fnorm = fnorm1;
        ++iter;
      }
      int small_f = fabsf(actred) <= ftol && prered <= ftol && 0.5f * ratio <= 1.f;
      if (small_f && delta <= xtol * xnorm) { status = 3; done = 1; break; }
      if (small_f) { status = 1; done = 1; break; }
      if (delta <= xtol * xnorm) { status = 2; done = 1; break; }
      if (nfev >= maxfev) { status = 5; done = 1; break; }
      if (fabsf(actred) <= LM_EPS && prered <= LM_EPS && 0.5f * ratio <= 1.f) { status = 6; done = 1; break; }
      if (delta <= LM_EPS * xnorm) { status = 7; done = 1; break; }
      if (gnorm <= LM_EPS) { status = 8; done = 1; break; }
    } while (ratio < 1e-4f);
    if (done) break;
  }
  if (nfev_out) *nfev_out = nfev;
  free(fvec); free(J);
  return status;
}

/* optional hook: tests may route the LM step through oracle/_ref (the reference tree's own Eigen LM) */
typedef int (*hop_lm_fn)(const float *, const float *, const float *, int, float *, int *);
static hop_lm_fn g_lm_backend = hop_oracle_lm_point_to_plane;
void hop_oracle_set_lm_backend(hop_lm_fn fn) { g_lm_backend = fn ? fn : hop_oracle_lm_point_to_plane; }

/* ------------------------------------------------------------------------------------------------
 * Utils::runICP<PointT>(src, tgt, T, max_iter, rejection_angle, max_corres_dist)   Utils.cpp:188-229
 * = pcl::IterativeClosestPoint<PointNormal,PointNormal> with CorrespondenceEstimation,
 *   CorrespondenceRejectorSurfaceNormal, TransformationEstimationPointToPlane (PCL 1.9 semantics, see header).
 * Output: T (4x4 col-major, source->target) = final transformation, identity when "not converged".
 * Returns the number of ICP iterations executed; *converged_out mirrors reg.hasConverged().
 * abs_mse_eps: 1e-6 in the reference (Utils.cpp:208).
 * ---------------------------------------------------------------------------------------------- */
int hop_oracle_run_icp(const float *src_xyz_in, const float *src_nrm_in, int ns_in, const float *tgt_xyz_in,
                       const float *tgt_nrm_in, int nt_in, float *T_out, int max_iter, float rejection_angle,
                       float max_corres_dist, double abs_mse_eps, int *converged_out) {
  /* removeNaNNormalsFromPointCloud (Utils.cpp:198-199) */
  float *sx = (float *)malloc(sizeof(float) * 3 * (size_t)(ns_in + 1)), *sn = (float *)malloc(sizeof(float) * 3 * (size_t)(ns_in + 1));
  float *tx = (float *)malloc(sizeof(float) * 3 * (size_t)(nt_in + 1)), *tn = (float *)malloc(sizeof(float) * 3 * (size_t)(nt_in + 1));
  int ns = 0, nt = 0;
  for (int i = 0; i < ns_in; ++i)
    if (isfinite(src_nrm_in[3 * i]) && isfinite(src_nrm_in[3 * i + 1]) && isfinite(src_nrm_in[3 * i + 2])) {
      memcpy(sx + 3 * ns, src_xyz_in + 3 * i, 12); memcpy(sn + 3 * ns, src_nrm_in + 3 * i, 12); ++ns;
    }
  for (int i = 0; i < nt_in; ++i)
    if (isfinite(tgt_nrm_in[3 * i]) && isfinite(tgt_nrm_in[3 * i + 1]) && isfinite(tgt_nrm_in[3 * i + 2])) {
      memcpy(tx + 3 * nt, tgt_xyz_in + 3 * i, 12); memcpy(tn + 3 * nt, tgt_nrm_in + 3 * i, 12); ++nt;
    }
  kd_tree *tree = kd_build(tx, nt); /* target tree, rebuilt per call exactly like the reference */
  const double rej_thr = cos(rejection_angle / 180.0 * M_PI); /* Utils.cpp:205 (double threshold_) */
  const float max_d2 = max_corres_dist * max_corres_dist;
  float final_T[16], T[16];
  m4_identity(final_T); m4_identity(T);
  float *cs = (float *)malloc(sizeof(float) * 9 * (size_t)(ns + 1));
  float *ct = cs + 3 * (size_t)ns, *cn = cs + 6 * (size_t)ns;
  double prev_mse = DBL_MAX;
  int iters = 0, converged = 0;
  if (max_iter < 1) max_iter = 1;
  for (;;) {
    int m = 0; double mse = 0.0;
    for (int i = 0; i < ns; ++i) {
      float d2;
      int j = kd_nn(tree, sx + 3 * i, &d2);
      if (j < 0 || d2 > max_d2) continue;                       /* CorrespondenceEstimation */
      float dot = sn[3 * i] * tn[3 * j] + sn[3 * i + 1] * tn[3 * j + 1] + sn[3 * i + 2] * tn[3 * j + 2];
      if (!((double)dot > rej_thr)) continue;                   /* CorrespondenceRejectorSurfaceNormal */
      memcpy(cs + 3 * m, sx + 3 * i, 12); memcpy(ct + 3 * m, tx + 3 * j, 12); memcpy(cn + 3 * m, tn + 3 * j, 12);
      mse += d2; ++m;
    }
    if (m < 3) { converged = 0; break; }                        /* min_number_correspondences_ = 3 */
    if (m >= 4) {                                               /* TransformationEstimationLM needs >= 4 */
      float x[6] = {0, 0, 0, 0, 0, 0};
      g_lm_backend(cs, ct, cn, m, x, NULL);
      hop_oracle_warp6d(x, T);
    } /* else: transformation_ keeps its previous value (PCL early-returns without touching it) */
    hop_oracle_transform_cloud(T, sx, sn, ns, sx, sn);          /* transformCloud, normals rotated */
    m4_mul(T, final_T, final_T);
    ++iters;
    /* DefaultConvergenceCriteria::hasConverged */
    if (iters >= max_iter) { converged = 1; break; }
    double cos_angle = 0.5 * ((double)M4(T, 0, 0) + (double)M4(T, 1, 1) + (double)M4(T, 2, 2) - 1.0);
    double tsq = (double)M4(T, 0, 3) * M4(T, 0, 3) + (double)M4(T, 1, 3) * M4(T, 1, 3) + (double)M4(T, 2, 3) * M4(T, 2, 3);
    if (cos_angle >= 1.0 && tsq <= 0.0) { converged = 1; break; } /* transformation_epsilon 0, similar-iters 0 */
    mse /= (double)m;
    if (fabs(mse - prev_mse) < abs_mse_eps) { converged = 1; break; }
    /* relative MSE threshold is overwritten with -DBL_MAX by ICP (euclidean_fitness_epsilon_): never fires */
    prev_mse = mse;
  }
  if (converged) memcpy(T_out, final_T, sizeof(final_T)); else m4_identity(T_out); /* Utils.cpp:218-225 */
  if (converged_out) *converged_out = converged;
  kd_free(tree); free(sx); free(sn); free(tx); free(tn); free(cs);
  return iters;
}

/* ------------------------------------------------------------------------------------------------
 * PoseEstimator::refineByICP body for a batch of hypotheses (PoseEstimator.cpp:257-273), OpenMP over
 * hypotheses with schedule(dynamic) like the reference.  poses: H x 16 (model2scene, col-major) in/out.
 * ---------------------------------------------------------------------------------------------- */
void hop_oracle_refine_by_icp(const float *s_xyz, const float *s_nrm, int ns, const float *m_xyz, const float *m_nrm, int nm,
                              float *poses, int H, int max_iter, float angle_thres, float dist_thres, double abs_mse_eps,
                              int nthreads, int *iters_out, int *converged_out) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    float *mx = (float *)malloc(sizeof(float) * 3 * (size_t)(nm + 1)), *mn = (float *)malloc(sizeof(float) * 3 * (size_t)(nm + 1));
#pragma omp for schedule(dynamic)
    for (int i = 0; i < H; ++i) {
      float *pose = poses + 16 * (size_t)i;
      float T[16], Tinv[16], out[16];
      hop_oracle_transform_cloud(pose, m_xyz, m_nrm, nm, mx, mn);            /* :263 */
      int conv = 0;
      int it = hop_oracle_run_icp(s_xyz, s_nrm, ns, mx, mn, nm, T, max_iter, angle_thres, dist_thres, abs_mse_eps, &conv); /* :266 */
      if (m4_inverse(T, Tinv) == 0) { m4_mul(Tinv, pose, out); memcpy(pose, out, sizeof(out)); } /* :267-269 */
      if (iters_out) iters_out[i] = it;
      if (converged_out) converged_out[i] = conv;
    }
    free(mx); free(mn);
  }
}

/* ------------------------------------------------------------------------------------------------
 * Utils::runICP(pclSegment, pclModel, offsetTransform, max_corres_dist)            Utils.cpp:135-164
 * = pcl::IterativeClosestPoint<PointXYZRGBNormal, PointXYZRGBNormal> with its defaults (PCL 1.9):
 *   setUseReciprocalCorrespondences(true)  -> CorrespondenceEstimation::determineReciprocalCorrespondences: source point i
 *       -> nearest target j (d^2 <= max^2) -> nearest SOURCE point of target j must be i (and within max)
 *   TransformationEstimationSVD (use_umeyama = true): pcl::umeyama = Eigen::umeyama without scaling, float
 *   setMaximumIterations(100); DefaultConvergenceCriteria: absolute MSE 1e-12 (default), relative MSE overwritten with
 *   -DBL_MAX, rotation threshold 1.0 / translation threshold 0 (transformation_epsilon_ = 0)
 *   setRANSACIterations(100) has no effect without a CorrespondenceRejectorSampleConsensus (none is added).
 * The Kabsch step: R = U diag(1,1,+-1) V^T of the cross-covariance, t = c_tgt - R c_src (tests pin it against the reference
 * tree's own Eigen::umeyama, oracle/_ref).  PARITY UNPINNED against PCL itself (not installed).
 * Output as hop_oracle_run_icp: T = final transformation (identity when not converged); returns the iterations.
 * ---------------------------------------------------------------------------------------------- */
static void sym3_jacobi(double A[3][3], double V[3][3]) { /* eigen decomposition of a symmetric 3x3: A -> diag, V = vectors */
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = i == j;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    if (off < 1e-300) break;
    for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
      if (A[p][q] == 0.0) continue;
      double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
      double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
      for (int k = 0; k < 3; ++k) { double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - sn * akq; A[k][q] = sn * akp + c * akq; }
      for (int k = 0; k < 3; ++k) { double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - sn * aqk; A[q][k] = sn * apk + c * aqk; }
      for (int k = 0; k < 3; ++k) { double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - sn * vkq; V[k][q] = sn * vkp + c * vkq; }
    }
  }
}

/* rigid transform (4x4 col-major) minimising sum |R s + t - d|^2 over n pairs; returns 0 when the configuration is degenerate */
int hop_oracle_kabsch(const float *src, const float *dst, int n, float *T) {
  double cs[3] = {0, 0, 0}, cd[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) { cs[k] += src[3 * i + k]; cd[k] += dst[3 * i + k]; }
  for (int k = 0; k < 3; ++k) { cs[k] /= n; cd[k] /= n; }
  double S[3][3] = {{0}}; /* sigma = sum (d - cd)(s - cs)^T */
  for (int i = 0; i < n; ++i)
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) S[r][c] += (dst[3 * i + r] - cd[r]) * (src[3 * i + c] - cs[c]);
  /* S = U D V^T:  S^T S = V D^2 V^T */
  double A[3][3], V[3][3];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { A[r][c] = 0; for (int k = 0; k < 3; ++k) A[r][c] += S[k][r] * S[k][c]; }
  sym3_jacobi(A, V);
  int ord[3] = {0, 1, 2}; /* singular values descending */
  for (int a = 0; a < 2; ++a) for (int b = a + 1; b < 3; ++b) if (A[ord[b]][ord[b]] > A[ord[a]][ord[a]]) { int t = ord[a]; ord[a] = ord[b]; ord[b] = t; }
  double Vs[3][3], U[3][3], sv[3];
  for (int c = 0; c < 3; ++c) { sv[c] = sqrt(fmax(A[ord[c]][ord[c]], 0.0)); for (int r = 0; r < 3; ++r) Vs[r][c] = V[r][ord[c]]; }
  if (!(sv[1] > 1e-12 * sv[0]) || !(sv[0] > 0)) return 0; /* rank < 2: no unique rotation */
  for (int c = 0; c < 2; ++c) for (int r = 0; r < 3; ++r) { double u = 0; for (int k = 0; k < 3; ++k) u += S[r][k] * Vs[k][c]; U[r][c] = u / sv[c]; }
  /* third columns by cross products so that det U = det V = +1, then the reflection case flips the sign of the last term */
  double v2[3] = {Vs[1][0] * Vs[2][1] - Vs[2][0] * Vs[1][1], Vs[2][0] * Vs[0][1] - Vs[0][0] * Vs[2][1], Vs[0][0] * Vs[1][1] - Vs[1][0] * Vs[0][1]};
  double u2[3] = {U[1][0] * U[2][1] - U[2][0] * U[1][1], U[2][0] * U[0][1] - U[0][0] * U[2][1], U[0][0] * U[1][1] - U[1][0] * U[0][1]};
  /* with det U = det V = +1 the optimal rotation is U V^T with all three terms positive (the sign of det(S) is absorbed) */
  double R[3][3];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[r][c] = U[r][0] * Vs[c][0] + U[r][1] * Vs[c][1] + u2[r] * v2[c];
  memset(T, 0, 16 * sizeof(float));
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) M4(T, r, c) = (float)R[r][c];
    M4(T, r, 3) = (float)(cd[r] - (R[r][0] * cs[0] + R[r][1] * cs[1] + R[r][2] * cs[2]));
  }
  M4(T, 3, 3) = 1.f;
  return 1;
}

int hop_oracle_run_icp_p2p(const float *src_xyz, int ns, const float *tgt_xyz, int nt, float *T_out, int max_iter,
                           float max_corres_dist, double abs_mse_eps, int *converged_out) {
  float *sx = (float *)malloc(sizeof(float) * 3 * (size_t)(ns + 1));
  memcpy(sx, src_xyz, sizeof(float) * 3 * (size_t)ns);
  kd_tree *tree = kd_build(tgt_xyz, nt);
  const float max_d2 = max_corres_dist * max_corres_dist;
  float final_T[16], T[16];
  m4_identity(final_T); m4_identity(T);
  float *cs = (float *)malloc(sizeof(float) * 6 * (size_t)(ns + 1)), *ct = cs + 3 * (size_t)ns;
  double prev_mse = DBL_MAX;
  int iters = 0, converged = 0;
  if (max_iter < 1) max_iter = 1;
  for (;;) {
    kd_tree *rtree = kd_build(sx, ns); /* tree_reciprocal_: the (moved) source cloud */
    int m = 0; double mse = 0.0;
    for (int i = 0; i < ns; ++i) {
      float d2, e2;
      int j = kd_nn(tree, sx + 3 * i, &d2);
      if (j < 0 || d2 > max_d2) continue;
      int k = kd_nn(rtree, tgt_xyz + 3 * j, &e2);
      if (e2 > max_d2 || k != i) continue;
      memcpy(cs + 3 * m, sx + 3 * i, 12); memcpy(ct + 3 * m, tgt_xyz + 3 * j, 12);
      mse += d2; ++m;
    }
    kd_free(rtree);
    if (m < 3) { converged = 0; break; }
    if (!hop_oracle_kabsch(cs, ct, m, T)) m4_identity(T);
    hop_oracle_transform_cloud(T, sx, NULL, ns, sx, NULL);
    m4_mul(T, final_T, final_T);
    ++iters;
    if (iters >= max_iter) { converged = 1; break; }
    double cos_angle = 0.5 * ((double)M4(T, 0, 0) + (double)M4(T, 1, 1) + (double)M4(T, 2, 2) - 1.0);
    double tsq = (double)M4(T, 0, 3) * M4(T, 0, 3) + (double)M4(T, 1, 3) * M4(T, 1, 3) + (double)M4(T, 2, 3) * M4(T, 2, 3);
    if (cos_angle >= 1.0 && tsq <= 0.0) { converged = 1; break; }
    mse /= (double)m;
    if (fabs(mse - prev_mse) < abs_mse_eps) { converged = 1; break; }
    prev_mse = mse;
  }
  if (converged) memcpy(T_out, final_T, sizeof(final_T)); else m4_identity(T_out);
  if (converged_out) *converged_out = converged;
  kd_free(tree); free(sx); free(cs);
  return iters;
}

/* a batch of hypotheses like hop_oracle_refine_by_icp (the model is moved by the hypothesis, the scene is the source) */
void hop_oracle_refine_by_icp_p2p(const float *s_xyz, int ns, const float *m_xyz, int nm, float *poses, int H, int max_iter,
                                  float dist_thres, double abs_mse_eps, int nthreads, int *iters_out, int *converged_out) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    float *mx = (float *)malloc(sizeof(float) * 3 * (size_t)(nm + 1));
#pragma omp for schedule(dynamic)
    for (int i = 0; i < H; ++i) {
      float *pose = poses + 16 * (size_t)i;
      float T[16], Tinv[16], out[16];
      hop_oracle_transform_cloud(pose, m_xyz, NULL, nm, mx, NULL);
      int conv = 0;
      int it = hop_oracle_run_icp_p2p(s_xyz, ns, mx, nm, T, max_iter, dist_thres, abs_mse_eps, &conv);
      if (m4_inverse(T, Tinv) == 0) { m4_mul(Tinv, pose, out); memcpy(pose, out, sizeof(out)); }
      if (iters_out) iters_out[i] = it;
      if (converged_out) converged_out[i] = conv;
    }
    free(mx);
  }
}

/* PoseEstimator::selectBest (PoseEstimator.cpp:474-498): LCP of every hypothesis, returns argmax (first best) */
int hop_oracle_select_best(const float *s_xyz, const float *s_nrm, int ns, const float *weights, const float *m_xyz,
                           const float *m_nrm, int nm, const float *poses, int H, float lcp_dist, float normal_angle,
                           int nthreads, float *scores_out) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    float *mx = (float *)malloc(sizeof(float) * 3 * (size_t)(nm + 1)), *mn = (float *)malloc(sizeof(float) * 3 * (size_t)(nm + 1));
#pragma omp for schedule(dynamic)
    for (int i = 0; i < H; ++i) {
      hop_oracle_transform_cloud(poses + 16 * (size_t)i, m_xyz, m_nrm, nm, mx, mn); /* :487 */
      scores_out[i] = hop_oracle_compute_lcp(s_xyz, s_nrm, ns, weights, mx, mn, nm, lcp_dist, normal_angle, 1, 1, 1); /* :488 */
    }
    free(mx); free(mn);
  }
  int best = 0; float best_lcp = 0.f; /* best_hypo = _pose_hypos[0]; best_lcp = 0 (:468-469); '>' keeps the first */
  for (int i = 0; i < H; ++i) if (scores_out[i] > best_lcp) { best_lcp = scores_out[i]; best = i; }
  return best;
}

int hop_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * Super4PCS congruent-set verification (K3 oracle).  Restates, for a batch of congruent quadrilaterals:
 *   gr::MatchBase::ComputeRigidTransformation     src/OpenGR_4pcs/src/gr/algorithms/matchBase.hpp:230-377
 *   gr::CongruentSetExplorationBase::TryCongruentSet  .../congruentSetExplorationBase.hpp:221-340
 *   gr::CongruentSetExplorationBase::Verify           .../congruentSetExplorationBase.hpp:346-435
 *   gr::KdTree::doQueryRestrictedClosestIndex         src/OpenGR_4pcs/src/gr/accelerators/kdtree.h:342-404
 * PINNED: tests/ compare it with the reference's own matcher compiled in oracle/_ref (ref_opengr.cpp) on the
 * very quadrilaterals that matcher generated.
 * Pc / Qc: CENTRED clouds (MatchBase::init removes the centroids, matchBase.hpp:425-432).
 * ---------------------------------------------------------------------------------------------- */
/* Eigen evaluates fixed-size 3-term reductions (dot, squaredNorm, 3x3 coefficient products) as t0 + (t1 + t2)
 * (redux_novec_unroller splits [0,1) | [1,3)); the (R^T R).isIdentity(1e-6) gate sits at float rounding level, so the
 * restatement keeps that association everywhere. */
static inline float vq_sum3(float t0, float t1, float t2) { return t0 + (t1 + t2); }

static int vq_frame(const float *a0, const float *a1, const float *a2, float *e /* 3x3 rows */) {
  float v1[3] = {a1[0] - a0[0], a1[1] - a0[1], a1[2] - a0[2]};
  float n = vq_sum3(v1[0] * v1[0], v1[1] * v1[1], v1[2] * v1[2]);
  if (n == 0.f) return 0;
  n = sqrtf(n); v1[0] /= n; v1[1] /= n; v1[2] /= n;
  float d[3] = {a2[0] - a0[0], a2[1] - a0[1], a2[2] - a0[2]};
  float k = vq_sum3(d[0] * v1[0], d[1] * v1[1], d[2] * v1[2]);
  float v2[3] = {d[0] - k * v1[0], d[1] - k * v1[1], d[2] - k * v1[2]};
  n = vq_sum3(v2[0] * v2[0], v2[1] * v2[1], v2[2] * v2[2]);
  if (n == 0.f) return 0;
  n = sqrtf(n); v2[0] /= n; v2[1] /= n; v2[2] /= n;
  e[0] = v1[0]; e[1] = v1[1]; e[2] = v1[2]; e[3] = v2[0]; e[4] = v2[1]; e[5] = v2[2];
  e[6] = v1[1] * v2[2] - v1[2] * v2[1]; e[7] = v1[2] * v2[0] - v1[0] * v2[2]; e[8] = v1[0] * v2[1] - v1[1] * v2[0];
  return 1;
}

int hop_oracle_verify_quads(const float *Pc, int nP, const float *Qc, int nQ, const int32_t *bases, const int32_t *quads,
                            const int32_t *quad_trial, int M, const float *cP, const float *cQ, float delta, float *poses,
                            float *lcp, int32_t *valid, int nthreads) {
  kd_tree *tree = kd_build(Pc, nP);
  const float sq_eps = delta * delta;
  int emitted = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic) reduction(+ : emitted)
  for (int m = 0; m < M; ++m) {
    const int32_t *b = bases + 4 * quad_trial[m], *q = quads + 4 * m;
    const float *p0 = Pc + 3 * b[0], *p1 = Pc + 3 * b[1], *p2 = Pc + 3 * b[2];
    const float *q0 = Qc + 3 * q[0], *q1 = Qc + 3 * q[1], *q2 = Qc + 3 * q[2];
    float c1[3], c2[3];
    for (int k = 0; k < 3; ++k) { c1[k] = ((p0[k] + p1[k]) + p2[k]) / 3.f; c2[k] = ((q0[k] + q1[k]) + q2[k]) / 3.f; }
    float Fp[9], Fq[9], R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    int ok = vq_frame(p0, p1, p2, Fp) && vq_frame(q0, q1, q2, Fq);
    float rms = FLT_MAX;
    if (ok) {
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[3 * i + j] = vq_sum3(Fp[i] * Fq[j], Fp[3 + i] * Fq[3 + j], Fp[6 + i] * Fq[6 + j]);
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { /* (R^T R).isIdentity(1e-6) */
        float g = vq_sum3(R[i] * R[j], R[3 + i] * R[3 + j], R[6 + i] * R[6 + j]);
        if (i == j) { if (!(fabsf(g - 1.f) <= 1e-6f * fminf(fabsf(g), 1.f))) ok = 0; }
        else if (!(fabsf(g) <= 1e-6f)) ok = 0;
      }
    }
    if (ok) {
      const float *qs[3] = {q0, q1, q2}, *ps[3] = {p0, p1, p2};
      float s = 0.f;
      for (int i = 0; i < 3; ++i) {
        float f[3] = {qs[i][0] - c2[0], qs[i][1] - c2[1], qs[i][2] - c2[2]}, d[3];
        for (int r = 0; r < 3; ++r) d[r] = (vq_sum3(R[3 * r] * f[0], R[3 * r + 1] * f[1], R[3 * r + 2] * f[2]) - ps[i][r]) + c1[r];
        s += sqrtf(vq_sum3(d[0] * d[0], d[1] * d[1], d[2] * d[2]));
      }
      rms = s / 4.f; /* divides by ref.size() == 4 (matchBase.hpp:357) */
    }
    float t[3];
    for (int r = 0; r < 3; ++r) t[r] = c1[r] - vq_sum3(R[3 * r] * c2[0], R[3 * r + 1] * c2[1], R[3 * r + 2] * c2[2]);
    float score = 0.f;
    if (ok && rms >= 0.f && rms < delta) {
      unsigned good = 0;
      for (int i = 0; i < nQ; ++i) {
        const float *x = Qc + 3 * i;
        float y[3];
        for (int r = 0; r < 3; ++r) y[r] = R[3 * r] * x[0] + R[3 * r + 1] * x[1] + R[3 * r + 2] * x[2] + t[r];
        float d2;
        int j = kd_nn(tree, y, &d2);
        if (j >= 0 && d2 <= sq_eps) ++good;
      }
      score = (float)good / (float)nQ;
    }
    lcp[m] = score;
    valid[m] = score > 0.f;
    emitted += valid[m];
    float *o = poses + 16 * (size_t)m;
    memset(o, 0, 16 * sizeof(float));
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) M4(o, r, c) = R[3 * r + c];
      float g[3] = {c2[0] + cQ[0], c2[1] + cQ[1], c2[2] + cQ[2]};
      M4(o, r, 3) = c1[r] + cP[r] - (R[3 * r] * g[0] + R[3 * r + 1] * g[1] + R[3 * r + 2] * g[2]);
    }
    M4(o, 3, 3) = 1.f;
  }
  kd_free(tree);
  return emitted;
}
