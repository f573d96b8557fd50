/* hop_oracle_render.c -- TEST INFRASTRUCTURE ONLY (CPU oracle; never on the product path).
 *
 * Plain-C restatement of the render-based rejection of the reference:
 *   PoseEstimator::rejectByRender                       src/perception/src/PoseEstimator.cpp:345-463
 *   Renderer::addObject / doRender                      src/perception/src/Renderer.cpp:42-81
 *   the OpenGL camera of pcl::simulation (depth_sim):   src/depth_sim/src/range_likelihood.cpp:391-426 (projection from the
 *     intrinsics, z_near 0.1, z_far 2.0), :428-475 (CV camera -> GL axes), src/depth_sim/src/simulation_io.cpp:411-440
 *     (depth buffer -> millimetres, rounded, image flipped up-down), :486-505 (colour buffer, flipped)
 * PARITY UNPINNED: the reference rasterises with OpenGL (no GL context here); this file states the same camera
 * and the same per-pixel rules with a software rasteriser whose coverage is decided in integer arithmetic, so that the CUDA
 * path can be held to it bit for bit.  Pixels on triangle borders may differ from a particular GL implementation.
 *
 * The camera: a CV-frame point (X, Y, Z) lands at window x = fx X/Z + cx, window y (bottom-up) = cy - fy Y/Z; GL samples pixel
 * centres, and both read-backs flip the image, so image pixel (x, y) samples sx = x + 0.5, sy = y + 0.5 of
 *     sx = fx X/Z + cx,   sy = fy Y/Z + (height - cy).
 * Depth: 1/Z is affine over a triangle in window space (evaluated in double with fma, then one float reciprocal); the nearest fragment wins (GL_LESS, 32-bit depth); the object is drawn
 * after the hand, so it owns a pixel only where it is strictly nearer.  sim = round(1000 Z) mm / 1000, clamped to [0.1, 2.0];
 * background = the cleared depth = z_far = 2.0.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct hop_oracle_render_params { /* same layout as hop_render_params (include/hop_c_api.h) */
  float fx, fy, cx, cy;
  int32_t width, height;
  float z_near, z_far;       /* 0.1, 2.0 */
  float roi_weight;          /* render_roi_weight */
  float keep_ratio;          /* render_keep_hypo */
} hop_oracle_render_params;

#define SUB 256 /* sub-pixel grid of the coverage test */

/* vertices (n x 3, already in the camera frame when T == NULL, else moved by the column-major 4x4 T like
 * Utils::transformPolygonMesh) rasterised into zbuf (float Z per pixel, +inf where empty): nearest fragment wins */
static void raster(const hop_oracle_render_params *p, const float *V, const int32_t *F, int nf, const float *T, float *zbuf) {
  const int W = p->width, H = p->height;
  for (int f = 0; f < nf; ++f) {
    float X[3], Y[3], Z[3];
    long long SX[3], SY[3];
    int ok = 1;
    for (int k = 0; k < 3; ++k) {
      const float *v = V + 3 * F[3 * f + k];
      float x = v[0], y = v[1], z = v[2];
      if (T) {
        const float tx = ((T[0] * x + T[4] * y) + T[8] * z) + T[12];
        const float ty = ((T[1] * x + T[5] * y) + T[9] * z) + T[13];
        const float tz = ((T[2] * x + T[6] * y) + T[10] * z) + T[14];
        x = tx; y = ty; z = tz;
      }
      X[k] = x; Y[k] = y; Z[k] = z;
      if (!(z > p->z_near)) { ok = 0; break; }          /* no near-plane clipping: such a triangle is dropped */
      const float sx = p->fx * (x / z) + p->cx;
      const float sy = p->fy * (y / z) + ((float)H - p->cy);
      SX[k] = (long long)floor((double)sx * SUB + 0.5);
      SY[k] = (long long)floor((double)sy * SUB + 0.5);
    }
    if (!ok) continue;
    long long area = (SX[1] - SX[0]) * (SY[2] - SY[0]) - (SY[1] - SY[0]) * (SX[2] - SX[0]);
    if (area == 0) continue;
    int i0 = 0, i1 = 1, i2 = 2;
    if (area < 0) { i1 = 2; i2 = 1; area = -area; }     /* both windings are drawn (no culling) */
    long long mnx = SX[0], mxx = SX[0], mny = SY[0], mxy = SY[0];
    for (int k = 1; k < 3; ++k) { if (SX[k] < mnx) mnx = SX[k]; if (SX[k] > mxx) mxx = SX[k]; if (SY[k] < mny) mny = SY[k]; if (SY[k] > mxy) mxy = SY[k]; }
    /* pixel x is sampled at x*SUB + SUB/2 */
    long long x0 = (mnx - SUB / 2 + SUB - 1) / SUB, x1 = (mxx - SUB / 2) / SUB, y0 = (mny - SUB / 2 + SUB - 1) / SUB, y1 = (mxy - SUB / 2) / SUB;
    if (mnx - SUB / 2 < 0) x0 = 0;
    if (mny - SUB / 2 < 0) y0 = 0;
    if (x0 < 0) x0 = 0;
    if (y0 < 0) y0 = 0;
    if (x1 > W - 1) x1 = W - 1;
    if (y1 > H - 1) y1 = H - 1;
    /* 1/Z is affine in window space: per-triangle weights w_k = (1/Z_k) / area, per pixel one fused dot product and one float reciprocal */
    const double w0 = (1.0 / (double)Z[i0]) / (double)area, w1 = (1.0 / (double)Z[i1]) / (double)area, w2 = (1.0 / (double)Z[i2]) / (double)area;
    for (long long y = y0; y <= y1; ++y)
      for (long long x = x0; x <= x1; ++x) {
        const long long px = x * SUB + SUB / 2, py = y * SUB + SUB / 2;
        const long long e0 = (SX[i2] - SX[i1]) * (py - SY[i1]) - (SY[i2] - SY[i1]) * (px - SX[i1]);   /* weight of vertex i0 */
        const long long e1 = (SX[i0] - SX[i2]) * (py - SY[i2]) - (SY[i0] - SY[i2]) * (px - SX[i2]);
        const long long e2 = (SX[i1] - SX[i0]) * (py - SY[i0]) - (SY[i1] - SY[i0]) * (px - SX[i0]);
        if (e0 < 0 || e1 < 0 || e2 < 0) continue;
        const double iz = fma((double)e2, w2, fma((double)e1, w1, (double)e0 * w0));
        const float z = 1.0f / (float)iz;
        if (!(z > p->z_near && z < p->z_far)) continue;   /* clipped by the near / far planes */
        float *dst = zbuf + (size_t)y * W + x;
        if (z < *dst) *dst = z;
      }
  }
}

static float sim_of(float z, const hop_oracle_render_params *p) { /* simulation_io.cpp:429 + Renderer.cpp:70-73 */
  if (!(z < FLT_MAX)) return p->z_far;                    /* cleared depth: exactly z_far = 2.0 */
  float s = (float)(int)roundf(1000.f * z) / 1000.0f;
  if (s > 2.0f) s = 2.0f;
  if (s < 0.1f) s = 0.1f;
  return s;
}

static float diff_of(float sim, float real) { /* PoseEstimator.cpp:410-423; the literals are doubles */
  if ((double)real <= 0.1 || (double)real >= 2.0) return 2.0f;
  if ((double)sim <= 0.1 || (double)sim >= 2.0) return 2.0f;
  return fabsf(sim - real);
}

/* one hypothesis: the simulated depth image (metres) and the object mask (1 where the object is the nearest surface);
 * hand_V / hand_F: every enabled hand mesh already in the camera frame (may be empty) */
int hop_oracle_render_depth(const hop_oracle_render_params *p, const float *hand_V, int hand_nv, const int32_t *hand_F, int hand_nf,
                            const float *obj_V, int obj_nv, const int32_t *obj_F, int obj_nf, const float *pose, float *depth, uint8_t *mask) {
  const size_t n = (size_t)p->width * p->height;
  float *zh = (float *)malloc(sizeof(float) * n), *zo = (float *)malloc(sizeof(float) * n);
  if (!zh || !zo) { free(zh); free(zo); return -1; }
  for (size_t i = 0; i < n; ++i) zh[i] = zo[i] = FLT_MAX;
  (void)hand_nv; (void)obj_nv;
  if (hand_nf > 0) raster(p, hand_V, hand_F, hand_nf, NULL, zh);
  raster(p, obj_V, obj_F, obj_nf, pose, zo);
  for (size_t i = 0; i < n; ++i) {
    const int ob = zo[i] < zh[i];
    depth[i] = sim_of(ob ? zo[i] : zh[i], p);
    if (mask) mask[i] = (uint8_t)ob;
  }
  free(zh); free(zo);
  return 0;
}

/* PoseEstimator::rejectByRender: wrong_ratio[h] for every hypothesis; order = the hypotheses kept, in the order the reference's
 * priority queue pops them (ascending wrong ratio; ties and NaN -- an object that owns no pixel divides 0 by 0 -- are left
 * unspecified by std::priority_queue: here lower index first, NaN last); n_keep = min(max(int(keep_ratio * H), 10), H) */
int hop_oracle_reject_by_render(const hop_oracle_render_params *p, const float *depth_m, const float *hand_V, int hand_nv, const int32_t *hand_F,
                                int hand_nf, const float *obj_V, int obj_nv, const int32_t *obj_F, int obj_nf, const float *poses, int H,
                                float *wrong_ratio, int32_t *order, int32_t *n_keep) {
  const size_t n = (size_t)p->width * p->height;
  float *zh = (float *)malloc(sizeof(float) * n);
  if (!zh) return -1;
  for (size_t i = 0; i < n; ++i) zh[i] = FLT_MAX;
  (void)hand_nv; (void)obj_nv;
  if (hand_nf > 0) raster(p, hand_V, hand_F, hand_nf, NULL, zh);
  int rc = 0;
#pragma omp parallel for schedule(dynamic)
  for (int h = 0; h < H; ++h) {
    float *zo = (float *)malloc(sizeof(float) * n);
    if (!zo) { rc = -1; continue; }
    for (size_t i = 0; i < n; ++i) zo[i] = FLT_MAX;
    raster(p, obj_V, obj_F, obj_nf, poses + 16 * (size_t)h, zo);
    float roi_diff = 0, bg_diff = 0;
    int roi_cnt = 0, bg_cnt = 0;
    for (size_t i = 0; i < n; ++i) {                      /* row-major, sequential float sums like the reference */
      const int ob = zo[i] < zh[i];
      const float diff = diff_of(sim_of(ob ? zo[i] : zh[i], p), depth_m[i]);
      if (ob) { roi_diff += diff; roi_cnt++; } else { bg_diff += diff; bg_cnt++; }
    }
    wrong_ratio[h] = p->roi_weight * roi_diff / roi_cnt + bg_diff / bg_cnt;
    free(zo);
  }
  free(zh);
  int keep = (int)(p->keep_ratio * H);
  if (keep < 10) keep = 10;
  if (keep > H) keep = H;
  /* stable selection by (wrong ratio asc, NaN last, index asc) */
  int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)(H > 0 ? H : 1));
  if (!idx) return -1;
  for (int h = 0; h < H; ++h) idx[h] = h;
  for (int a = 1; a < H; ++a) { /* insertion sort: test sizes */
    const int32_t k = idx[a];
    const float wk = wrong_ratio[k];
    int b = a - 1;
    while (b >= 0) {
      const float wb = wrong_ratio[idx[b]];
      const int after = (wb != wb) ? (wk == wk) : (wk == wk && wk < wb);   /* k goes before idx[b]? */
      if (!after) break;
      idx[b + 1] = idx[b]; --b;
    }
    idx[b + 1] = k;
  }
  for (int i = 0; i < keep; ++i) order[i] = idx[i];
  *n_keep = keep;
  free(idx);
  return rc;
}
