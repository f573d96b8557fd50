// PoseHypo.h -- class PoseHypo of the reference (src/perception/include/PoseHypo.h:7-27): the record every stage passes on.
#pragma once
#include "mat.h"

class PoseHypo {
 public:
  PoseHypo() {}
  explicit PoseHypo(int id) : _id(id) {}
  PoseHypo(const Mat4f &pose, int id, float lcp_score) : _pose(pose), _lcp_score(lcp_score), _id(id) {}
  Mat4f _pose;
  float _wrong_ratio = 1.f;
  float _lcp_score = 0.f;
  int _id = -1;
};
