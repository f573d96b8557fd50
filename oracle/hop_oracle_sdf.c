/* hop_oracle_sdf.c -- TEST INFRASTRUCTURE ONLY (CPU oracle; never on the product path).
 *
 * Plain-C restatement of the physics pruning step of the reference:
 *   - igl::signed_distance(..., SIGNED_DISTANCE_TYPE_PSEUDONORMAL, ...)      src/perception/include/igl/signed_distance.cpp:90-101,160-205
 *       closest point: Ericson's region walk                                  igl/point_simplex_squared_distance.cpp:44-108
 *       sign: igl::pseudonormal_test                                          igl/pseudonormal_test.cpp:24-130
 *       normals: per_face_normals.cpp:19-35, per_vertex_normals.cpp:39-110 (angle weights, internal_angles.cpp:68-86),
 *                per_edge_normals.cpp:21-75 (uniform, NOT normalised)
 *     The reference finds the closest face with an AABB tree; this restatement scans every face and keeps the first (lowest
 *     index) among exact ties -- same distance, possibly another face of an exact tie.
 *   - SDFchecker::transformVertices / transformMesh                           src/perception/src/SDFchecker.cpp:22-33,80-86
 *   - PoseEstimator::rejectByCollisionOrNonTouching                           src/perception/src/PoseEstimator.cpp:524-735
 *     followed per hypothesis in the reference's order: the mesh is moved INTO the hypothesis' frame (float, like
 *     transformVertices) and every normal is rebuilt, as igl does on every call.
 * Pinned against the reference's own libigl compiled from where it lies (oracle/ref_sdf.cpp -> oracle/_ref/libhop_ref.so)
 * by tests/test_sdf_oracle.py.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int nf;
  float *tri;  /* nf x 9: a, b, c */
  float *fn;   /* nf x 3 */
  float *vn;   /* nf x 9: vertex normal of corner 0/1/2 */
  float *en;   /* nf x 9: normal of the edge OPPOSITE corner 0/1/2 (EMAP(F.rows()*x+f)) */
  unsigned char *big; /* doublearea > MIN_DOUBLE_AREA */
} sdf_mesh;

static void sdf_mesh_free(sdf_mesh *m) { free(m->tri); free(m->fn); free(m->vn); free(m->en); free(m->big); }

typedef struct { int u, w, f, c; } edge_rec;
static int edge_cmp(const void *pa, const void *pb) {
  const edge_rec *a = (const edge_rec *)pa, *b = (const edge_rec *)pb;
  if (a->u != b->u) return a->u < b->u ? -1 : 1;
  if (a->w != b->w) return a->w < b->w ? -1 : 1;
  if (a->c != b->c) return a->c < b->c ? -1 : 1; /* igl accumulates in (f + c*m) order: c-major... the loop is f-major */
  return a->f < b->f ? -1 : (a->f > b->f);
}

static inline float dot3f(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* igl::doublearea(A,B,C) for 3-D corners: float edge lengths, Kahan's Heron in double (doublearea.cpp:83-115,145-199) */
static double doublearea3(const float *A, const float *B, const float *C) {
  float e0[3], e1[3], e2[3];
  for (int k = 0; k < 3; ++k) { e0[k] = B[k] - C[k]; e1[k] = C[k] - A[k]; e2[k] = A[k] - B[k]; }
  double l[3] = {(double)sqrtf(dot3f(e0, e0)), (double)sqrtf(dot3f(e1, e1)), (double)sqrtf(dot3f(e2, e2))};
  for (int i = 0; i < 3; ++i) for (int j = i + 1; j < 3; ++j) if (l[j] > l[i]) { double t = l[i]; l[i] = l[j]; l[j] = t; } /* descending */
  const double arg = (l[0] + (l[1] + l[2])) * (l[2] - (l[0] - l[1])) * (l[2] + (l[0] - l[1])) * (l[0] + (l[1] - l[2]));
  return 2.0 * 0.25 * sqrt(arg);
}

/* builds every normal igl::signed_distance builds for the PSEUDONORMAL type, from float vertices V (nv x 3) */
static int sdf_mesh_build(sdf_mesh *m, const float *V, int nv, const int32_t *F, int nf) {
  m->nf = nf;
  m->tri = (float *)malloc(sizeof(float) * 9 * (size_t)nf);
  m->fn = (float *)malloc(sizeof(float) * 3 * (size_t)nf);
  m->vn = (float *)malloc(sizeof(float) * 9 * (size_t)nf);
  m->en = (float *)malloc(sizeof(float) * 9 * (size_t)nf);
  m->big = (unsigned char *)malloc((size_t)nf);
  float *VN = (float *)calloc((size_t)nv * 3, sizeof(float));
  edge_rec *er = (edge_rec *)malloc(sizeof(edge_rec) * 3 * (size_t)nf);
  if (!m->tri || !m->fn || !m->vn || !m->en || !m->big || !VN || !er) return -1;
  for (int f = 0; f < nf; ++f) {
    const float *p[3] = {V + 3 * F[3 * f], V + 3 * F[3 * f + 1], V + 3 * F[3 * f + 2]};
    memcpy(m->tri + 9 * f, p[0], 12); memcpy(m->tri + 9 * f + 3, p[1], 12); memcpy(m->tri + 9 * f + 6, p[2], 12);
    float v1[3], v2[3], n[3];
    for (int k = 0; k < 3; ++k) { v1[k] = p[1][k] - p[0][k]; v2[k] = p[2][k] - p[0][k]; }
    n[0] = v1[1] * v2[2] - v1[2] * v2[1]; n[1] = v1[2] * v2[0] - v1[0] * v2[2]; n[2] = v1[0] * v2[1] - v1[1] * v2[0];
    const float r = sqrtf(dot3f(n, n));
    for (int k = 0; k < 3; ++k) m->fn[3 * f + k] = r == 0.f ? 0.f : n[k] / r;
    m->big[f] = doublearea3(p[0], p[1], p[2]) > 1e-4;
    /* squared edge lengths: column d = edge opposite corner d; angle at corner d = acos((s3+s2-s1)/(2 sqrt(s3 s2))) */
    float L[3];
    for (int d = 0; d < 3; ++d) { float e[3]; for (int k = 0; k < 3; ++k) e[k] = p[(d + 1) % 3][k] - p[(d + 2) % 3][k]; L[d] = dot3f(e, e); }
    for (int d = 0; d < 3; ++d) {
      const float s1 = L[d], s2 = L[(d + 1) % 3], s3 = L[(d + 2) % 3];
      const float w = (float)acos((double)(s3 + s2 - s1) / (2. * sqrt((double)(s3 * s2))));
      for (int k = 0; k < 3; ++k) VN[3 * F[3 * f + d] + k] += w * m->fn[3 * f + k];
      const int u = F[3 * f + (d + 1) % 3], w2 = F[3 * f + (d + 2) % 3];
      er[3 * f + d].u = u < w2 ? u : w2; er[3 * f + d].w = u < w2 ? w2 : u; er[3 * f + d].f = f; er[3 * f + d].c = d;
    }
  }
  for (int v = 0; v < nv; ++v) { /* N.rowwise().normalize() */
    const float r = sqrtf(dot3f(VN + 3 * v, VN + 3 * v));
    if (r > 0.f) for (int k = 0; k < 3; ++k) VN[3 * v + k] /= r;
  }
  for (int f = 0; f < nf; ++f) for (int d = 0; d < 3; ++d) memcpy(m->vn + 9 * f + 3 * d, VN + 3 * F[3 * f + d], 12);
  qsort(er, 3 * (size_t)nf, sizeof(edge_rec), edge_cmp);
  for (size_t i = 0; i < 3 * (size_t)nf;) {
    size_t j = i;
    float s[3] = {0.f, 0.f, 0.f};
    while (j < 3 * (size_t)nf && er[j].u == er[i].u && er[j].w == er[i].w) { for (int k = 0; k < 3; ++k) s[k] += m->fn[3 * er[j].f + k]; ++j; }
    for (size_t q = i; q < j; ++q) memcpy(m->en + 9 * er[q].f + 3 * er[q].c, s, 12);
    i = j;
  }
  free(VN); free(er);
  return 0;
}

/* Ericson, as igl::point_simplex_squared_distance states it; returns the squared distance, writes the closest point */
static float closest_on_triangle(const float *p, const float *a, const float *b, const float *c, float *out) {
  float ab[3], ac[3], ap[3], bp[3], cp[3];
  for (int k = 0; k < 3; ++k) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; ap[k] = p[k] - a[k]; }
  const float d1 = dot3f(ab, ap), d2 = dot3f(ac, ap);
  int done = 0;
  if (d1 <= 0.f && d2 <= 0.f) { memcpy(out, a, 12); done = 1; }
  float d3 = 0, d4 = 0, d5 = 0, d6 = 0, vc = 0, vb = 0, va = 0;
  if (!done) {
    for (int k = 0; k < 3; ++k) bp[k] = p[k] - b[k];
    d3 = dot3f(ab, bp); d4 = dot3f(ac, bp);
    if (d3 >= 0.f && d4 <= d3) { memcpy(out, b, 12); done = 1; }
  }
  if (!done) {
    vc = d1 * d4 - d3 * d2;
    if ((a[0] != b[0] || a[1] != b[1] || a[2] != b[2]) && vc <= 0.f && d1 >= 0.f && d3 <= 0.f) {
      const float v = d1 / (d1 - d3);
      for (int k = 0; k < 3; ++k) out[k] = a[k] + v * ab[k];
      done = 1;
    }
  }
  if (!done) {
    for (int k = 0; k < 3; ++k) cp[k] = p[k] - c[k];
    d5 = dot3f(ab, cp); d6 = dot3f(ac, cp);
    if (d6 >= 0.f && d5 <= d6) { memcpy(out, c, 12); done = 1; }
  }
  if (!done) {
    vb = d5 * d2 - d1 * d6;
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) {
      const float w = d2 / (d2 - d6);
      for (int k = 0; k < 3; ++k) out[k] = a[k] + w * ac[k];
      done = 1;
    }
  }
  if (!done) {
    va = d3 * d6 - d5 * d4;
    if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
      const float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
      for (int k = 0; k < 3; ++k) out[k] = b[k] + w * (c[k] - b[k]);
      done = 1;
    }
  }
  if (!done) {
    const float denom = 1.0f / (va + vb + vc);
    const float v = vb * denom, w = vc * denom;
    for (int k = 0; k < 3; ++k) out[k] = a[k] + ab[k] * v + ac[k] * w;
  }
  float d[3] = {p[0] - out[0], p[1] - out[1], p[2] - out[2]};
  return dot3f(d, d);
}

/* igl::pseudonormal_test for face f with closest point c of query q: returns +1 / -1, writes the normal used */
static float pseudonormal_sign(const sdf_mesh *m, int f, const float *q, const float *c, float *n_out) {
  const float *A = m->tri + 9 * f, *B = A + 3, *C = A + 6;
  const float *n = m->fn + 3 * f;
  const double eps = 1e-12;
  if (m->big[f]) {
    float v0[3], v1[3], v2[3];
    for (int k = 0; k < 3; ++k) { v0[k] = B[k] - A[k]; v1[k] = C[k] - A[k]; v2[k] = c[k] - A[k]; }
    const float d00 = dot3f(v0, v0), d01 = dot3f(v0, v1), d11 = dot3f(v1, v1), d20 = dot3f(v2, v0), d21 = dot3f(v2, v1);
    const float denom = d00 * d11 - d01 * d01;
    float b[3];
    b[1] = (d11 * d20 - d01 * d21) / denom;
    b[2] = (d00 * d21 - d01 * d20) / denom;
    b[0] = 1.0f - (b[1] + b[2]);
    const int type = ((double)b[0] <= eps) + ((double)b[1] <= eps) + ((double)b[2] <= eps);
    if (type == 2) { for (int x = 0; x < 3; ++x) if ((double)b[x] > eps) { n = m->vn + 9 * f + 3 * x; break; } }
    else if (type == 1) { for (int x = 0; x < 3; ++x) if ((double)b[x] <= eps) { n = m->en + 9 * f + 3 * x; break; } }
  } else {
    int found = 0;
    for (int v = 0; v < 3 && !found; ++v) {
      const float d[3] = {c[0] - A[3 * v], c[1] - A[3 * v + 1], c[2] - A[3 * v + 2]};
      if ((double)sqrtf(dot3f(d, d)) < eps) { found = 1; n = m->vn + 9 * f + 3 * v; }
    }
    for (int e = 0; e < 3 && !found; ++e) {
      const float *s = A + 3 * ((e + 1) % 3), *d = A + 3 * ((e + 2) % 3);
      float dms[3], smp[3];
      for (int k = 0; k < 3; ++k) { dms[k] = d[k] - s[k]; smp[k] = s[k] - c[k]; }
      const double vsq = (double)dot3f(dms, dms);
      double t = -(double)dot3f(dms, smp) / vsq;
      float r[3];
      for (int k = 0; k < 3; ++k) r[k] = c[k] - (float)((1 - t) * (double)s[k] + t * (double)d[k]);
      double sq = (double)dot3f(r, r);
      if (t < 0) { for (int k = 0; k < 3; ++k) r[k] = c[k] - s[k]; sq = (double)dot3f(r, r); }
      else if (t > 1) { for (int k = 0; k < 3; ++k) r[k] = c[k] - d[k]; sq = (double)dot3f(r, r); }
      if (sqrt(sq) < eps) { n = m->en + 9 * f + 3 * e; found = 1; }
    }
  }
  const float qc[3] = {q[0] - c[0], q[1] - c[1], q[2] - c[2]};
  if (n_out) memcpy(n_out, n, 12);
  return dot3f(qc, n) >= 0.f ? 1.f : -1.f;
}

/* test support: set to 1 when some face within a relative tie_tol (3e-7, a few ulp) of the smallest squared distance (a float tie that another
 * summation order, an FMA, or igl's AABB traversal order resolves the other way) would sign the distance differently -- the
 * reference's own answer is then a coin toss (e.g. a far point whose foot lies within ~1e-6 m of a sharp edge of small faces:
 * the face fallback of pseudonormal_test.cpp:117-121 signs it with whichever face came first) */
static float tie_tol = 3e-7f, dot_tol = 1e-6f;
void hop_oracle_sdf_amb_tolerances(float tie, float dot) { tie_tol = tie; dot_tol = dot; }
static float sdf_point_amb(const sdf_mesh *m, const float *q, float s, float best, int *amb) {
  for (int f = 0; f < m->nf; ++f) {
    float c[3];
    const float d2 = closest_on_triangle(q, m->tri + 9 * f, m->tri + 9 * f + 3, m->tri + 9 * f + 6, c);
    if (d2 > best * (1.f + tie_tol) + 1e-14f) continue;
    /* every normal igl could pick for this foot: the face's, and those of the vertices / edges the foot lies on (within 1e-6 m) */
    const float qc[3] = {q[0] - c[0], q[1] - c[1], q[2] - c[2]};
    const float len = sqrtf(dot3f(qc, qc));
    const float *cand[7]; int nc = 0;
    cand[nc++] = m->fn + 3 * f;
    for (int v = 0; v < 3; ++v) {
      const float *P = m->tri + 9 * f + 3 * v;
      const float d[3] = {c[0] - P[0], c[1] - P[1], c[2] - P[2]};
      if (sqrtf(dot3f(d, d)) < 1e-6f) cand[nc++] = m->vn + 9 * f + 3 * v;
    }
    for (int e = 0; e < 3; ++e) {
      const float *S = m->tri + 9 * f + 3 * ((e + 1) % 3), *D = m->tri + 9 * f + 3 * ((e + 2) % 3);
      float dms[3], sc[3];
      for (int k = 0; k < 3; ++k) { dms[k] = D[k] - S[k]; sc[k] = c[k] - S[k]; }
      float t = dot3f(dms, sc) / dot3f(dms, dms);
      t = t < 0.f ? 0.f : (t > 1.f ? 1.f : t);
      const float r[3] = {sc[0] - t * dms[0], sc[1] - t * dms[1], sc[2] - t * dms[2]};
      if (sqrtf(dot3f(r, r)) < 1e-6f) cand[nc++] = m->en + 9 * f + 3 * e;
    }
    for (int k = 0; k < nc; ++k) {
      const float nl = sqrtf(dot3f(cand[k], cand[k]));
      const float dt = dot3f(qc, cand[k]);
      if (dt * s < 0.f || fabsf(dt) < dot_tol * len * nl) { *amb = 1; return s; }
    }
  }
  return s;
}

static float sdf_point(const sdf_mesh *m, const float *q, int *face, float *cp);
static float sdf_point_flag(const sdf_mesh *m, const float *q, int *amb) {
  const float s = sdf_point(m, q, NULL, NULL);
  if (amb && s == s && fabsf(s) > 1e-4f) sdf_point_amb(m, q, s, s * s, amb);   /* next to the surface a flipped sign moves no threshold test */
  return s;
}

static float sdf_point(const sdf_mesh *m, const float *q, int *face, float *cp) {
  float best = FLT_MAX, bc[3] = {0, 0, 0};
  int bf = -1;
  for (int f = 0; f < m->nf; ++f) {
    float c[3];
    const float d2 = closest_on_triangle(q, m->tri + 9 * f, m->tri + 9 * f + 3, m->tri + 9 * f + 6, c);
    if (d2 < best) { best = d2; bf = f; memcpy(bc, c, 12); }
  }
  if (face) *face = bf;
  if (cp) memcpy(cp, bc, 12);
  if (bf < 0) return FLT_MAX;
  /* signed_distance.cpp:127-128,158: low_sqr_d = 0 for the (-FLT_MAX, FLT_MAX) bounds SDFchecker passes, and
   * "sqrd <= low_sqr_d" sends a point lying exactly on the mesh out of bounds: S = NaN (min / max / counts skip it here) */
  if (best == 0.f) return NAN;
  return pseudonormal_sign(m, bf, q, bc, NULL) * sqrtf(best);
}

/* S[i] = signed distance of pts[i] to the mesh; I (closest face) and Cp (closest point) may be NULL */
int hop_oracle_signed_distance(const float *pts, int n, const float *V, int nv, const int32_t *F, int nf, float *S, int32_t *I, float *Cp) {
  sdf_mesh m;
  if (sdf_mesh_build(&m, V, nv, F, nf)) return -1;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    int f; float c[3];
    S[i] = sdf_point(&m, pts + 3 * i, &f, c);
    if (I) I[i] = f;
    if (Cp) memcpy(Cp + 3 * i, c, 12);
  }
  sdf_mesh_free(&m);
  return 0;
}

/* ---- PoseEstimator::rejectByCollisionOrNonTouching ------------------------------------------------------------------ */
typedef struct hop_oracle_collision_params { /* same layout as hop_collision_params (include/hop_c_api.h) */
  float cam2handbase[16];  /* column-major: hand->_handbase_in_cam.inverse() */
  float model_center[3];   /* _model_center_init */
  float ob_diameter;
  float collision_dist;    /* min(-_smallest_dim * collision_thres, -0.007) */
  float inside_ob_dist;    /* min(-_smallest_dim / 5, -0.01) */
  float non_touch_dist;
  float collision_finger_dist;          /* -cfg["collision_finger_dist"] */
  float collision_finger_volume_ratio;
  int32_t finger_status[4];             /* _component_status of finger_1_1, finger_1_2, finger_2_1, finger_2_2 */
} hop_oracle_collision_params;

static void m4_mulf(const float *A, const float *B, float *C) {
  for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) {
    float s = 0.f;
    for (int k = 0; k < 4; ++k) s += A[r + 4 * k] * B[k + 4 * c];
    C[r + 4 * c] = s;
  }
}
static void xform_pts(const float *T, const float *in, int n, float *out) { /* SDFchecker::transformVertices: pose * [v;1], float */
  for (int i = 0; i < n; ++i) {
    const float x = in[3 * i], y = in[3 * i + 1], z = in[3 * i + 2];
    for (int r = 0; r < 3; ++r) out[3 * i + r] = T[r] * x + T[r + 4] * y + T[r + 8] * z + T[r + 12];
  }
}
static int nn_brute(const float *pts, int n, const float *q, float *d2out) {
  int best = -1; float bd = FLT_MAX;
  for (int i = 0; i < n; ++i) {
    const float d[3] = {pts[3 * i] - q[0], pts[3 * i + 1] - q[1], pts[3 * i + 2] - q[2]};
    const float d2 = dot3f(d, d);
    if (d2 < bd) { bd = d2; best = i; }
  }
  if (d2out) *d2out = bd;
  return best;
}

/* Finger arrays are concatenated in the order finger_1_1, finger_1_2, finger_2_1, finger_2_2 (std::map order of both
 * finger_cloud_eigens and _component_status); a finger with n = 0 points is "not in the map" (component disabled), a finger
 * mesh with 0 faces is not registered.  Finger clouds and finger meshes are in the hand-base frame already.
 * keep[h] = 1 when hypothesis h survives; reason[h]: 0 kept, 1 scene point inside the object, 2 hand point inside the
 * object, 3 finger cloud penetrates, 4 one side does not touch, 5 object penetrates a finger mesh, 6 object inside a finger.
 * diag (may be NULL): H x 10 floats = sdf of the scene point, of the hand point, min over each finger cloud (4), min of the model
 * over each finger mesh (4); FLT_MAX where the step was not reached. */
/* mode 0: the distance as found; mode 1 / 2 (test support): a point whose sign is a coin toss (sdf_point_amb) counts as
 * outside / inside -- a hypothesis whose decision is the same in all three modes does not depend on any coin toss */
static float sdf_eval(const sdf_mesh *m, const float *q, int mode) {
  if (mode == 0) return sdf_point(m, q, NULL, NULL);
  int amb = 0;
  const float s = sdf_point_flag(m, q, &amb);
  return amb ? (mode == 1 ? fabsf(s) : -fabsf(s)) : s;
}

typedef struct {
  const float *scene_xyz, *hand_xyz, *model_xyz, *finger_pts;
  const int32_t *finger_n;
  int ns, nh, nm;
  const int *poff;
  const sdf_mesh *fm;
  const hop_oracle_collision_params *p;
} reject_ctx;

static int reject_one(const reject_ctx *c, const sdf_mesh *om, const float *M, const float *P, int mode, float *dg) {
  const hop_oracle_collision_params *p = c->p;
  float ctr[3];
  xform_pts(M, p->model_center, 1, ctr);
  if (c->ns > 0) { /* nearest scene point to the object centre inside the object? (PoseEstimator.cpp:596-615) */
    const int i = nn_brute(c->scene_xyz, c->ns, ctr, NULL);
    const float s = sdf_eval(om, c->scene_xyz + 3 * i, mode);
    if (dg) dg[0] = s;
    if (s <= p->inside_ob_dist) return 1;
  }
  if (c->nh > 0) { /* quick check: the hand point nearest to the object centre (:618-641) */
    float d2;
    const int i = nn_brute(c->hand_xyz, c->nh, ctr, &d2);
    if (sqrtf(d2) < p->ob_diameter / 2) {
      const float s = sdf_eval(om, c->hand_xyz + 3 * i, mode);
      if (dg) dg[1] = s;
      if (s < p->collision_dist) return 2;
    }
  }
  int non_touch[4] = {0, 0, 0, 0};
  for (int k = 0; k < 4; ++k) { /* finger clouds against the object (:645-668) */
    if (c->finger_n[k] <= 0) continue;
    if (!p->finger_status[0] && k < 2) continue;
    if (!p->finger_status[2] && k >= 2) continue;
    float mn = FLT_MAX;
    for (int i = 0; i < c->finger_n[k]; ++i) { const float s = sdf_eval(om, c->finger_pts + 3 * (c->poff[k] + i), mode); if (s < mn) mn = s; }
    if (dg) dg[2 + k] = mn;
    if (mn <= p->collision_dist) return 3;
    if (mn > p->non_touch_dist && p->finger_status[k]) non_touch[k] = 1;
  }
  if ((non_touch[0] && non_touch[1]) || (non_touch[2] && non_touch[3])) return 4; /* :675-680 */
  for (int k = 0; k < 4; ++k) { /* the object's points against every finger mesh (:683-723) */
    if (c->fm[k].nf <= 0 || c->nm <= 0) continue;
    float mn = FLT_MAX; int inside = 0;
    for (int i = 0; i < c->nm; ++i) { const float s = sdf_eval(&c->fm[k], P + 3 * i, mode); if (s < mn) mn = s; if (s < 0.f) ++inside; }
    if (dg) dg[6 + k] = mn;
    if (mn < p->collision_finger_dist) return 5;
    if ((float)(inside / c->nm) > p->collision_finger_volume_ratio) return 6; /* integer division, as in the reference */
  }
  return 0;
}

int hop_oracle_reject_by_collision_amb(const float *objV, int onv, const int32_t *objF, int onf, const float *fingerV, const int32_t *fnv,
                                       const int32_t *fingerF, const int32_t *fnf, const float *finger_pts, const int32_t *finger_n,
                                       const float *scene_xyz, int ns, const float *hand_xyz, int nh, const float *model_xyz, int nm,
                                       const float *poses, int H, const hop_oracle_collision_params *p, int32_t *keep, int32_t *reason,
                                       float *diag, int32_t *ambiguous /* H or NULL: 1 = the decision hangs on a coin toss (sdf_eval) */) {
  sdf_mesh fm[4];
  int voff[5] = {0}, foff[5] = {0}, poff[5] = {0};
  for (int k = 0; k < 4; ++k) { voff[k + 1] = voff[k] + fnv[k]; foff[k + 1] = foff[k] + fnf[k]; poff[k + 1] = poff[k] + finger_n[k]; }
  for (int k = 0; k < 4; ++k) { fm[k].nf = 0; if (fnf[k] > 0 && sdf_mesh_build(&fm[k], fingerV + 3 * voff[k], fnv[k], fingerF + 3 * foff[k], fnf[k])) return -1; }
  const reject_ctx c = {scene_xyz, hand_xyz, model_xyz, finger_pts, finger_n, ns, nh, nm, poff, fm, p};
  int rc = 0;
#pragma omp parallel for schedule(dynamic)
  for (int h = 0; h < H; ++h) {
    float M[16];
    m4_mulf(p->cam2handbase, poses + 16 * (size_t)h, M);
    float *Vt = (float *)malloc(sizeof(float) * 3 * (size_t)onv);
    float *P = (float *)malloc(sizeof(float) * 3 * (size_t)(nm > 0 ? nm : 1));
    float *dg = diag ? diag + 10 * (size_t)h : NULL;
    if (dg) for (int k = 0; k < 10; ++k) dg[k] = FLT_MAX;
    sdf_mesh om;
    if (!Vt || !P) { rc = -1; free(Vt); free(P); continue; }
    xform_pts(M, objV, onv, Vt);   /* SDFchecker::transformMesh("object", model2handbase) */
    if (sdf_mesh_build(&om, Vt, onv, objF, onf)) { rc = -1; free(Vt); free(P); continue; }
    xform_pts(M, model_xyz, nm, P);
    const int why = reject_one(&c, &om, M, P, 0, dg);
    keep[h] = why == 0;
    if (reason) reason[h] = why;
    if (ambiguous) ambiguous[h] = reject_one(&c, &om, M, P, 1, NULL) != why || reject_one(&c, &om, M, P, 2, NULL) != why;
    sdf_mesh_free(&om); free(Vt); free(P);
  }
  for (int k = 0; k < 4; ++k) if (fm[k].nf > 0) sdf_mesh_free(&fm[k]);
  return rc;
}

int hop_oracle_reject_by_collision(const float *objV, int onv, const int32_t *objF, int onf, const float *fingerV, const int32_t *fnv,
                                   const int32_t *fingerF, const int32_t *fnf, const float *finger_pts, const int32_t *finger_n,
                                   const float *scene_xyz, int ns, const float *hand_xyz, int nh, const float *model_xyz, int nm,
                                   const float *poses, int H, const hop_oracle_collision_params *p, int32_t *keep, int32_t *reason,
                                   float *diag) {
  return hop_oracle_reject_by_collision_amb(objV, onv, objF, onf, fingerV, fnv, fingerF, fnf, finger_pts, finger_n, scene_xyz, ns, hand_xyz, nh,
                                            model_xyz, nm, poses, H, p, keep, reason, diag, NULL);
}

/* per-point ambiguity flags of hop_oracle_signed_distance (test support, see sdf_point_amb) */
int hop_oracle_signed_distance_amb(const float *pts, int n, const float *V, int nv, const int32_t *F, int nf, float *S, int32_t *amb) {
  sdf_mesh m;
  if (sdf_mesh_build(&m, V, nv, F, nf)) return -1;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) { int a = 0; S[i] = sdf_point_flag(&m, pts + 3 * i, &a); amb[i] = a; }
  sdf_mesh_free(&m);
  return 0;
}
