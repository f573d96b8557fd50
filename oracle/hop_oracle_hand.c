/*
 * hop_oracle_hand.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, never shipped, never on the product path).
 *
 * Plain-C restatement of the hand-state overlap objective and of the swarm schedule that drives it:
 *   objFuncPSO                      /root/reference/src/perception/src/Hand.cpp:10-178
 *   FingerProperty (z histogram)    /root/reference/src/perception/src/Hand.cpp:182-250
 *   optim::pso_int (schedule only)  /root/reference/src/perception/include/unconstrained/pso.hpp:146-351
 *
 * PARITY UNPINNED: objFuncPSO needs PCL (transformPointCloudWithNormals, KdTreeFLANN), Armadillo and yaml-cpp, none
 * of which are in /root/reference or installed here, and the reference holds no golden vectors for it.
 * The restatement follows the source line by line, including its types: float score, `num_match += 1 + X[0]` in
 * double then rounded to float per match, double penalties rounded to float, PCL 1.9's unfused left-to-right
 * point transform, Eigen's AngleAxis -> Quaternion -> Matrix3f chain for tf_self, and the quirk that the neighbour's
 * NORMAL is read from `scene_hand_region` with an index that came from the kd-tree of
 * `scene_hand_region_removed_noise` (Hand.cpp:94 vs Hand.cpp:327-329).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/hop_c_api.h"

#define M4(m, r, c) ((m)[(c) * 4 + (r)])

/* exact 1-NN by brute force (any exact structure returns the same neighbour up to exact ties; lowest index wins) */
static int bf_nn(const float *pts, int n, const float *q, float *d2) {
  int b = -1; float bd = FLT_MAX;
  for (int i = 0; i < n; ++i) {
    float dx = pts[3 * i] - q[0], dy = pts[3 * i + 1] - q[1], dz = pts[3 * i + 2] - q[2];
    float d = dx * dx + dy * dy + dz * dz;
    if (d < bd) { bd = d; b = i; }
  }
  *d2 = bd;
  return b;
}

/* Eigen fixed-size 4x4 float product (SSE packet path): c(r,j) = ((a(r,0) b(0,j) + a(r,1) b(1,j)) + a(r,2) b(2,j)) + a(r,3) b(3,j) */
static void hand_m4_mul(const float *A, const float *B, float *C) {
  float R[16];
  for (int j = 0; j < 4; ++j)
    for (int r = 0; r < 4; ++r)
      M4(R, r, j) = ((M4(A, r, 0) * M4(B, 0, j) + M4(A, r, 1) * M4(B, 1, j)) + M4(A, r, 2) * M4(B, 2, j)) + M4(A, r, 3) * M4(B, 3, j);
  memcpy(C, R, sizeof(R));
}

/* tf_self of Hand.cpp:15-21: R = AngleAxisf(0,Z) * AngleAxisf(0,Y) * AngleAxisf(X[0],X) goes through Eigen quaternions
 * (AngleAxis * AngleAxis is a Quaternion product); the identity factors are exact, the matrix is
 * Quaternion(w = cos(a/2), x = sin(a/2)).toRotationMatrix(). */
void hop_oracle_hand_tf_self(double X, float *tf /* 16, column-major */) {
  const float a = (float)X;
  const float w = cosf(0.5f * a), x = sinf(0.5f * a);
  const float tx = 2.f * x, twx = tx * w, txx = tx * x;
  memset(tf, 0, 16 * sizeof(float));
  M4(tf, 0, 0) = 1.f;
  M4(tf, 1, 1) = 1.f - txx; M4(tf, 1, 2) = 0.f - twx;
  M4(tf, 2, 1) = 0.f + twx; M4(tf, 2, 2) = 1.f - txx;
  M4(tf, 3, 3) = 1.f;
}

/* inverse of the (rigid up to rounding) float transform: adjugate of the 3x3 block and -Rinv t in double, rounded to
 * float once.  (Eigen::Matrix4f::inverse() is a general cofactor inverse in float SSE; its bits are not reproducible
 * without Eigen, the double evaluation is within half an ulp of the exact inverse.) */
void hop_oracle_hand_inverse(const float *T, float *out) {
  double a[9];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) a[3 * r + c] = (double)M4(T, r, c);
  const double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
  const double det = (a[0] * c00 + a[1] * c01) + a[2] * c02;
  const double id = 1.0 / det;
  double I[9];
  I[0] = c00 * id; I[1] = (a[2] * a[7] - a[1] * a[8]) * id; I[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  I[3] = c01 * id; I[4] = (a[0] * a[8] - a[2] * a[6]) * id; I[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  I[6] = c02 * id; I[7] = (a[1] * a[6] - a[0] * a[7]) * id; I[8] = (a[0] * a[4] - a[1] * a[3]) * id;
  const double t0 = (double)M4(T, 0, 3), t1 = (double)M4(T, 1, 3), t2 = (double)M4(T, 2, 3);
  memset(out, 0, 16 * sizeof(float));
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) M4(out, r, c) = (float)I[3 * r + c];
    M4(out, r, 3) = (float)(-((I[3 * r] * t0 + I[3 * r + 1] * t1) + I[3 * r + 2] * t2));
  }
  M4(out, 3, 3) = 1.f;
}

/* FingerProperty::getBinAlongZ (Hand.cpp:239-245) */
static int hand_bin(const hop_finger_params *p, float z) {
  int bin = (int)(fmaxf(z - p->min_z, 0.0f) / p->stride_z);
  if (bin < 0) bin = 0;
  if (bin > p->num_division - 1) bin = p->num_division - 1;
  return bin;
}

/* FingerProperty::FingerProperty (Hand.cpp:184-236): bounding box + per-z-bin min/max; fills the fields of
 * hop_finger_params that come from it (min_z, stride_z, num_division, hist_min_y) and returns the box. */
void hop_oracle_finger_property(const float *xyz, int n, int num_division, hop_finger_params *p, float *bbox /* min xyz, max xyz */) {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) { if (xyz[3 * i + k] < mn[k]) mn[k] = xyz[3 * i + k]; if (xyz[3 * i + k] > mx[k]) mx[k] = xyz[3 * i + k]; }
  p->num_division = num_division;
  p->min_z = mn[2];
  p->stride_z = (mx[2] - mn[2]) / num_division;
  float hist[6][HOP_MAX_FINGER_BINS];
  int changed[HOP_MAX_FINGER_BINS];
  for (int b = 0; b < num_division; ++b) {
    changed[b] = 0;
    for (int k = 0; k < 3; ++k) { hist[k][b] = FLT_MAX; hist[3 + k][b] = -FLT_MAX; }
  }
  for (int i = 0; i < n; ++i) {
    const float *q = xyz + 3 * i;
    int b = hand_bin(p, q[2]);
    for (int k = 0; k < 3; ++k) { hist[k][b] = fminf(hist[k][b], q[k]); hist[3 + k][b] = fmaxf(hist[3 + k][b], q[k]); }
    changed[b] = 1;
  }
  for (int i = 0; i < num_division; ++i) { /* untouched bin: copy the next touched one */
    if (changed[i]) continue;
    for (int j = i + 1; j < num_division; ++j)
      if (changed[j]) { for (int k = 0; k < 6; ++k) hist[k][i] = hist[k][j]; changed[i] = 1; break; }
  }
  if (!changed[num_division - 1])
    for (int i = num_division - 2; i >= 0; --i)
      if (changed[i]) { for (int k = 0; k < 6; ++k) hist[k][num_division - 1] = hist[k][i]; changed[num_division - 1] = 1; break; }
  for (int b = 0; b < num_division; ++b) p->hist_min_y[b] = hist[1][b];
  if (bbox) { for (int k = 0; k < 3; ++k) { bbox[k] = mn[k]; bbox[3 + k] = mx[k]; } }
}

/* objFuncPSO for one state X (radians).
 *   f_xyz/f_nrm   : args->model (the finger cloud, its own frame)
 *   nn_xyz        : scene_hand_region_removed_noise (hand-base frame) -- what kdtree_scene indexes (Hand.cpp:327-328)
 *   lk_nrm        : normals of args->scene_hand_region (Hand.cpp:326), read with the kd-tree's index (Hand.cpp:94)
 *   w_xyz         : args->scene_remove_swivel (hand-base frame)
 * detail (may be NULL): [0] matches [1] num_outer [2] outer_dist_sum [3] branch (0 gap, 1 no match, 2 hard outer, 3 exp, 4 none) */
double hop_oracle_obj_func_pso(double X, const hop_finger_params *p, const float *f_xyz, const float *f_nrm, int nf,
                               const float *nn_xyz, int n_nn, const float *lk_nrm, const float *w_xyz, int nw, double *detail) {
  float score = 0.f;
  float tf_self[16], cur[16];
  hop_oracle_hand_tf_self(X, tf_self);
  hand_m4_mul(p->model2handbase, tf_self, cur); /* cur_model2handbase (:22) */

  /* gripper-gap penalty (:24-64) */
  float tip1[4], tip2[4], v[4];
  if (p->palm_side) {
    float out2hb[16];
    hand_m4_mul(cur, p->finger_out2parent, out2hb); /* (model2handbase * tf_self) * finger_out2parent (:30) */
    v[0] = p->tip1_local[0]; v[1] = p->tip1_local[1]; v[2] = p->tip1_local[2]; v[3] = 1.f;
    for (int r = 0; r < 4; ++r) tip1[r] = ((M4(out2hb, r, 0) * v[0] + M4(out2hb, r, 1) * v[1]) + M4(out2hb, r, 2) * v[2]) + M4(out2hb, r, 3) * v[3];
  } else {
    v[0] = p->tip1_local[0]; v[1] = p->tip1_local[1]; v[2] = p->tip1_local[2]; v[3] = 1.f;
    for (int r = 0; r < 4; ++r) tip1[r] = ((M4(cur, r, 0) * v[0] + M4(cur, r, 1) * v[1]) + M4(cur, r, 2) * v[2]) + M4(cur, r, 3) * v[3];
  }
  v[0] = p->tip2_local[0]; v[1] = p->tip2_local[1]; v[2] = p->tip2_local[2]; v[3] = 1.f;
  for (int r = 0; r < 4; ++r) tip2[r] = ((M4(cur, r, 0) * v[0] + M4(cur, r, 1) * v[1]) + M4(cur, r, 2) * v[2]) + M4(cur, r, 3) * v[3];
  float gd1, gd2;
  if (p->right_side) { gd1 = tip1[1] - p->pair_tip1_y; gd2 = tip2[1] - p->pair_tip2_y; }
  else { gd1 = -tip1[1] + p->pair_tip1_y; gd2 = -tip2[1] + p->pair_tip2_y; }
  const float G = p->gripper_min_dist;
  if (gd1 < G || gd2 < G) {
    float pen = (float)(1e3 + 1e3 * (double)fabsf(G - gd1));
    score -= pen;
    if (detail) { detail[0] = 0; detail[1] = 0; detail[2] = 0; detail[3] = 0; }
    return -(double)score;
  }

  /* matches (:67-128) */
  float num_match = 0.f;
  int matches = 0;
  const float cos_thr = (float)cos((double)p->normal_angle_deg / 180.0 * M_PI); /* float NORMAL_ANGLE_THRES = std::cos(float/180.0*M_PI) */
  const float thr2 = p->dist_thres * p->dist_thres;
  for (int i = 0; i < nf; ++i) {
    const float x = f_xyz[3 * i], y = f_xyz[3 * i + 1], z = f_xyz[3 * i + 2];
    float pt[3], n1[3];
    for (int r = 0; r < 3; ++r) pt[r] = M4(cur, r, 0) * x + M4(cur, r, 1) * y + M4(cur, r, 2) * z + M4(cur, r, 3);
    const float a = f_nrm[3 * i], b = f_nrm[3 * i + 1], c = f_nrm[3 * i + 2];
    for (int r = 0; r < 3; ++r) n1[r] = M4(cur, r, 0) * a + M4(cur, r, 1) * b + M4(cur, r, 2) * c;
    float d2;
    const int j = bf_nn(nn_xyz, n_nn, pt, &d2);
    if (j < 0 || !(d2 <= thr2)) continue;
    int hit = 0;
    if (!p->check_normal) hit = 1;
    else {
      const float *n2 = lk_nrm + 3 * j;
      if (n2[0] == 0 && n2[1] == 0 && n2[2] == 0) hit = 1;
      else if (isfinite(n2[0]) && isfinite(n2[1]) && isfinite(n2[2])) {
        const float dot = n1[0] * n2[0] + (n1[1] * n2[1] + n1[2] * n2[2]); /* Eigen 3-term redux: t0 + (t1 + t2) */
        if (dot >= cos_thr) hit = 1;
      }
    }
    if (hit) { num_match = (float)((double)num_match + (1 + X)); ++matches; }
  }
  score += num_match;
  if (num_match == 0) { /* (:135-139) */
    score = (float)(-100 + X);
    if (detail) { detail[0] = 0; detail[1] = 0; detail[2] = 0; detail[3] = 1; }
    return -(double)score;
  }

  /* outer-point penalty (:129-173) */
  float inv[16];
  hop_oracle_hand_inverse(cur, inv);
  float outer_sum = 0.f;
  int num_outer = 0;
  for (int i = 0; i < nw; ++i) {
    const float x = w_xyz[3 * i], y = w_xyz[3 * i + 1], z = w_xyz[3 * i + 2];
    const float py = M4(inv, 1, 0) * x + M4(inv, 1, 1) * y + M4(inv, 1, 2) * z + M4(inv, 1, 3);
    const float pz = M4(inv, 2, 0) * x + M4(inv, 2, 1) * y + M4(inv, 2, 2) * z + M4(inv, 2, 3);
    const int bin = hand_bin(p, pz);
    const float face = p->hist_min_y[bin];
    if (py >= face) continue;
    outer_sum += fabsf(py - face);
    ++num_outer;
  }
  const float avg = outer_sum / num_outer; /* 0/0 = NaN: every comparison below is false */
  int branch = 4;
  if (num_outer >= p->max_outter_pts || avg >= 0.005) {
    float pen = (float)(1e3 + (double)(p->outter_pt_dist_weight * fmaxf(avg - p->outter_pt_dist, 0.0f)));
    score -= pen; branch = 2;
  } else if (num_outer >= 0 && avg - p->outter_pt_dist > 0) {
    float pen = p->outter_pt_dist_weight * expf(avg * 1000);
    score -= pen; branch = 3;
  }
  if (detail) { detail[0] = matches; detail[1] = num_outer; detail[2] = outer_sum; detail[3] = branch; }
  return -(double)score;
}

/* a grid of S states (the dense replacement of the swarm): cost[s] = objFuncPSO(thetas[s]) */
void hop_oracle_hand_overlap(const hop_finger_params *p, const float *f_xyz, const float *f_nrm, int nf, const float *nn_xyz,
                             int n_nn, const float *lk_nrm, const float *w_xyz, int nw, const double *thetas, int S,
                             double *cost, double *detail /* S x 4 or NULL */, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic)
  for (int s = 0; s < S; ++s)
    cost[s] = hop_oracle_obj_func_pso(thetas[s], p, f_xyz, f_nrm, nf, nn_xyz, n_nn, lk_nrm, w_xyz, nw, detail ? detail + 4 * s : NULL);
}

/* optim::pso_int's schedule (pso.hpp:146-351) with the reference's settings (n_pop + 1 centre particle, inertia
 * method 1, velocity method 1, bounds mapped linearly to [0,1]); the random numbers are supplied by the caller
 * (Armadillo's generator is not reproducible here): r_init[n_pop], and per generation r_cog[n_pop+1], r_soc[n_pop+1].
 * Used by tests to show the dense grid's optimum is never worse than what the swarm finds.  Returns the best cost. */
double hop_oracle_pso_1d(double lb, double ub, int n_pop, int n_gen, double c_cog, double c_soc, double w0, const double *r_init,
                         const double *r_cog, const double *r_soc, double (*f)(double, void *), void *user, double *best_x) {
  const int n = n_pop + 1;
  double *P = malloc(sizeof(double) * n), *V = calloc(n, sizeof(double)), *bp = malloc(sizeof(double) * n), *bv = malloc(sizeof(double) * n);
  for (int i = 0; i < n_pop; ++i) P[i] = r_init[i];     /* pso_initial_lb/ub map to [0,1] under invLinearTransform */
  P[n_pop] = 0.5;                                       /* centre particle (:226-229) */
  double gbest = DBL_MAX, gx = 0.5;
  for (int i = 0; i < n; ++i) {
    bv[i] = f(lb + P[i] * (ub - lb), user); bp[i] = P[i];
    if (bv[i] < gbest) { gbest = bv[i]; gx = P[i]; }
  }
  double w = w0;
  const double par_w_max = 0.99, par_w_min = 0.10;       /* optim_structs.hpp:131-132 */
  for (int it = 0; it < n_gen; ++it) {
    for (int i = 0; i < n; ++i) {
      V[i] = w * V[i] + c_cog * r_cog[it * n + i] * (bp[i] - P[i]) + c_soc * r_soc[it * n + i] * (gx - P[i]);
      P[i] += V[i];
      if (P[i] < 0) P[i] = 0; if (P[i] > 1) P[i] = 1;  /* :283 */
    }
    for (int i = 0; i < n; ++i) {
      double v = f(lb + P[i] * (ub - lb), user);
      if (v < bv[i]) { bv[i] = v; bp[i] = P[i]; }
    }
    for (int i = 0; i < n; ++i) if (bv[i] < gbest) { gbest = bv[i]; gx = bp[i]; }
    w = par_w_min + (par_w_max - par_w_min) * (it + 1) / n_gen; /* inertia method 1 (:329-330) */
  }
  *best_x = lb + gx * (ub - lb);
  free(P); free(V); free(bp); free(bv);
  return gbest;
}

/* ---- HandT42::removeSurroundingPointsAndAssignProbability (src/perception/src/Hand.cpp:781-888) ----------------------------
 * TEST INFRASTRUCTURE ONLY.  The scene (camera frame, with normals) goes into the hand-base frame; a point within `dist_thres`
 * (squared; 5 mm for the proximal finger links, 20 mm for base / swivels) of its nearest neighbour in ANY link cloud, or
 * within that planar (x, y) distance and 5 mm in z of that neighbour, is dropped; the others get the confidence
 * 1 - exp(-lambda * min_dist), lambda = 231.049..., min_dist = the smallest nearest-neighbour distance over the links visited
 * (starting at 1.0).  Then points on the outer side of either distal link (y < 0 and z >= min_z in the link's own frame) are
 * dropped and the rest goes back to the camera frame.  Links are visited in std::map order of their names: the caller
 * passes them sorted.  kind[k]: 0 = dist_thres, 1 = finger_1_1 / finger_2_1 (5 mm), 2 = base / swivel_1 / swivel_2 (20 mm).
 * Output in input order (the reference's order is the OpenMP critical-section order).  Returns the number of points kept.
 * PARITY UNPINNED against PCL (FLANN nearest neighbour: exact, ties -> lowest index here). */
typedef struct hop_oracle_hand_removal_params { /* same layout as hop_hand_removal_params (include/hop_c_api.h) */
  float cam_in_handbase[16];   /* column-major: _handbase_in_cam.inverse() */
  float handbase_in_cam[16];
  float handbase_in_finger_1_2[16], handbase_in_finger_2_2[16]; /* getTFHandBase(name).inverse() */
  float min_z;                 /* _finger_properties["finger_1_2"]._min_z */
  float dist_thres_sq;         /* near_hand_dist^2 */
} hop_oracle_hand_removal_params;

static inline float hr_row(const float *T, int r, float x, float y, float z) { return ((T[r] * x + T[r + 4] * y) + T[r + 8] * z) + T[r + 12]; }
static inline float hr_rot(const float *T, int r, float x, float y, float z) { return (T[r] * x + T[r + 4] * y) + T[r + 8] * z; }

int hop_oracle_remove_hand_points(const float *xyz, const float *nrm, int n, const float *link_xyz, const int32_t *link_n, const int32_t *link_kind,
                                  int n_links, const hop_oracle_hand_removal_params *p, float *out_xyz, float *out_nrm, float *out_conf) {
  const float lambda = 231.04906018664843f;
  int *keep = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  float *hb = (float *)malloc(sizeof(float) * 7 * (size_t)(n > 0 ? n : 1));
  if (!keep || !hb) { free(keep); free(hb); return -1; }
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    const float x = hr_row(p->cam_in_handbase, 0, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    const float y = hr_row(p->cam_in_handbase, 1, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    const float z = hr_row(p->cam_in_handbase, 2, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    float *h = hb + 7 * (size_t)i;
    h[0] = x; h[1] = y; h[2] = z;
    for (int r = 0; r < 3; ++r) h[3 + r] = hr_rot(p->cam_in_handbase, r, nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]);
    int near = 0;
    float min_dist = 1.0f;
    const float *lp = link_xyz;
    for (int k = 0; k < n_links && !near; lp += 3 * (size_t)link_n[k], ++k) {
      const float thr = link_kind[k] == 1 ? (float)(0.005 * 0.005) : (link_kind[k] == 2 ? (float)(0.02 * 0.02) : p->dist_thres_sq);
      if (link_n[k] <= 0) continue;                       /* nearestKSearch returns 0 on an empty cloud */
      int bi = 0; float bd = FLT_MAX;
      for (int j = 0; j < link_n[k]; ++j) {
        const float dx = lp[3 * j] - x, dy = lp[3 * j + 1] - y, dz = lp[3 * j + 2] - z;
        const float d2 = (dx * dx + dy * dy) + dz * dz;
        if (d2 < bd) { bd = d2; bi = j; }
      }
      const float d = sqrtf(bd);
      if (d < min_dist) min_dist = d;
      if (bd <= thr) { near = 1; break; }
      const float px = x - lp[3 * bi], py = y - lp[3 * bi + 1];
      const float planar = px * px + py * py;
      if (planar <= thr && (double)fabsf(z - lp[3 * bi + 2]) <= 0.005) { near = 1; break; }
    }
    h[6] = 1 - expf(-lambda * min_dist);
    int ok = !near;
    if (ok) { /* outer side of the distal links */
      const float y1 = hr_row(p->handbase_in_finger_1_2, 1, x, y, z), z1 = hr_row(p->handbase_in_finger_1_2, 2, x, y, z);
      const float y2 = hr_row(p->handbase_in_finger_2_2, 1, x, y, z), z2 = hr_row(p->handbase_in_finger_2_2, 2, x, y, z);
      if ((y1 < 0 && z1 >= p->min_z) || (y2 < 0 && z2 >= p->min_z)) ok = 0;
    }
    keep[i] = ok;
  }
  int m = 0;
  for (int i = 0; i < n; ++i) {
    if (!keep[i]) continue;
    const float *h = hb + 7 * (size_t)i;
    for (int r = 0; r < 3; ++r) {
      out_xyz[3 * m + r] = hr_row(p->handbase_in_cam, r, h[0], h[1], h[2]);
      out_nrm[3 * m + r] = hr_rot(p->handbase_in_cam, r, h[3], h[4], h[5]);
    }
    out_conf[m] = h[6];
    ++m;
  }
  free(keep); free(hb);
  return m;
}

/* ---- HandT42::adjustHandHeight (Hand.cpp:999-1051) -- TEST INFRASTRUCTURE ONLY ------------------------------------------------
 * counts[t] = #{hand points p : NN of (p.x, p.y, p.z + heights[t]) in the scene (hand-base frame) has d^2 <= 0.005^2 and
 * n_p . n_nn >= cos(45 deg)}; returns the first index with the highest non-zero count, or -1.  Exact NN, ties -> lowest index. */
int hop_oracle_adjust_hand_height(const float *hand_xyz, const float *hand_nrm, int nh, const float *scene_xyz, const float *scene_nrm, int ns,
                                  const float *heights, int n_heights, int32_t *counts) {
  int best = -1, max_match = 0;
  for (int t = 0; t < n_heights; ++t) {
    int c = 0;
#pragma omp parallel for reduction(+ : c) schedule(static)
    for (int i = 0; i < nh; ++i) {
      const float x = hand_xyz[3 * i], y = hand_xyz[3 * i + 1], z = hand_xyz[3 * i + 2] + heights[t];
      int bi = -1; float bd = FLT_MAX;
      for (int j = 0; j < ns; ++j) {
        const float dx = scene_xyz[3 * j] - x, dy = scene_xyz[3 * j + 1] - y, dz = scene_xyz[3 * j + 2] - z;
        const float d2 = (dx * dx + dy * dy) + dz * dz;
        if (d2 < bd) { bd = d2; bi = j; }
      }
      if (bi < 0 || bd > (float)(0.005 * 0.005)) continue;
      const float *n = hand_nrm + 3 * i, *m = scene_nrm + 3 * bi;
      const float dot = n[0] * m[0] + (n[1] * m[1] + n[2] * m[2]);   /* Eigen's 3-vector dot: x0 y0 + (x1 y1 + x2 y2) */
      if ((double)dot >= cos(45 / 180.0 * M_PI)) ++c;
    }
    counts[t] = c;
    if (c > max_match) { max_match = c; best = t; }
  }
  return best;
}
