// PoseEstimator.h -- the pipeline stage object of the reference (src/perception/include/PoseEstimator.h:11-49) for the
// methods on the hot path; every method body is a call into libhop's C ABI (include/hop_c_api.h).
//   runSuper4pcs   PoseEstimator.cpp:62-100     clusterPoses   :106-233
//   refineByICP    :235-275                     selectBest     :465-502
// rejectByCollisionOrNonTouching / rejectByRender are outside the scope of this build: not declared.
#pragma once
#include <vector>

#include "ConfigParser.h"
#include "PoseHypo.h"
#include "cloud.h"
#include "hop_c_api.h"

class PoseEstimator {
 public:
  PoseEstimator(ConfigParser *cfg1, const Cloud &model, const Cloud &model001, hop_ctx *ctx);
  ~PoseEstimator();
  // object_segment: the hand-free object cloud with per-point confidence (setCurScene keeps confidence >= thres, :41-45)
  void setCurScene(const Cloud &object_segment);
  bool runSuper4pcs(const std::vector<int32_t> &ppf_keys /* n x 4: the keys of the reference's ppfs map */);
  void clusterPoses(float angle_diff, float dist_diff, bool assign_id);
  void refineByICP();
  void selectBest(PoseHypo &best_hypo);

  std::vector<PoseHypo> _pose_hypos;
  Cloud _scene_high_confidence;

 private:
  ConfigParser *cfg;
  hop_ctx *ctx;
  Cloud _model, _model001;
  hop_cloud *d_scene = nullptr, *d_model = nullptr, *d_model001 = nullptr;
  void check(int rc, const char *what);
};
