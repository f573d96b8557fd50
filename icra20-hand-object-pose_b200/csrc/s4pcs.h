// s4pcs.h -- the host-side plan of one Super4PCS registration (shared by k_s4pcs_plan.cu and k_s4pcs.cu).
#pragma once
#include <stdint.h>

#include <vector>

#include "../../include/hop_c_api.h"

struct S4Pt { float p[3]; float n[3]; };

struct S4Trial {
  int base_ok = 0;
  int base[4] = {0, 0, 0, 0};  // indices into P, reordered by TryQuadrilateral
  S4Pt b[4];                   // base_3D_ in that order (centred positions, unit normals)
  float inv1 = 0.f, inv2 = 0.f;
  float dist1 = 0.f, dist2 = 0.f;       // |b0 - b1|, |b2 - b3|
  float nangle1 = 0.f, nangle2 = 0.f;   // |n0 - n1|, |n2 - n3|
  float alpha = 0.f;                    // cosine between the two base segments
};

struct hop_s4pcs_plan {
  hop_s4pcs_options opt;
  std::vector<S4Pt> P, Q;          // centred; Q is the sampled model
  std::vector<float> P_prob;
  std::vector<int32_t> q_ids;      // index of each sampled Q point in the caller's Q
  std::vector<float> Qunit;        // 3 floats per sampled Q point: PairCreationFunctor::points (unit cube)
  float centroid_P[3] = {0, 0, 0}, centroid_Q[3] = {0, 0, 0};
  float gcenter[3] = {0, 0, 0}, ratio = 1.f;
  float diameter = 0.f;
  std::vector<S4Trial> trials;
  // results kept by hop_super4pcs_run when opt.keep_intermediates
  std::vector<int32_t> trial_ranges, pairs, quads;
  int trials_executed = 0;
};
