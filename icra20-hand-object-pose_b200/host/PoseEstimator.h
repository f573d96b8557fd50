// PoseEstimator.h -- the pipeline stage object of the reference (src/perception/include/PoseEstimator.h:11-49) for the
// methods on the hot path; every method body is a call into libhop's C ABI (include/hop_c_api.h).
//   runSuper4pcs   PoseEstimator.cpp:62-100     clusterPoses   :106-233
//   refineByICP    :235-275                     selectBest     :465-502
//   registerMesh / registerHandMesh  :506-521   rejectByCollisionOrNonTouching  :524-735  (physics pruning)
//   rejectByRender  :345-463  (render-based rejection, SURVEY.md 8f rank 4: a software rasteriser in place of the GL context)
#pragma once
#include <map>
#include <string>
#include <vector>

#include "ConfigParser.h"
#include "PoseHypo.h"
#include "cloud.h"
#include "hop_c_api.h"

// What rejectByCollisionOrNonTouching reads from the reference's HandT42 (Hand.h:29-92), as plain data in the hand-base frame.
// Link names: finger_1_1, finger_1_2, finger_2_1, finger_2_2.
struct HandState {
  std::map<std::string, bool> _component_status;   // hand->_component_status
  std::map<std::string, Cloud> finger_clouds;      // hand->_clouds[name] transformed by getTFHandBase(name) (PoseEstimator.cpp:541-551)
  Cloud _hand_cloud;                               // hand->_hand_cloud
  Mat4f _handbase_in_cam;
  // hand->_meshes of the enabled links, already moved by _handbase_in_cam * getTFHandBase(name) and concatenated
  // (what Renderer::addObject receives, PoseEstimator.cpp:362-383)
  std::vector<float> meshes_in_cam_V;
  std::vector<int32_t> meshes_in_cam_F;
};

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
  // SDFchecker::registerMesh through PoseEstimator::registerMesh / registerHandMesh: name = "object" or a finger link; the
  // vertices are moved by `pose` once (a finger link: its getTFHandBase) and stay on the device
  bool registerMesh(const std::string &mesh_dir, const std::string &name, const Mat4f &pose);
  void registerMesh(const std::vector<float> &V, const std::vector<int32_t> &F, const std::string &name, const Mat4f &pose);
  // cloud_withouthand_raw: the scene without the hand, camera frame (setCurScene's _cloud_withouthand_raw)
  void rejectByCollisionOrNonTouching(const HandState &hand, const Cloud &cloud_withouthand_raw);
  // depth_meters: the frame's depth image (_depth_meters, height x width, metres); the object mesh is the registered "object"
  void rejectByRender(float projection_thres, const HandState &hand, const std::vector<float> &depth_meters, int width, int height);

  std::vector<PoseHypo> _pose_hypos;
  Cloud _scene_high_confidence;

 private:
  ConfigParser *cfg;
  hop_ctx *ctx;
  Cloud _model, _model001;
  hop_cloud *d_scene = nullptr, *d_model = nullptr, *d_model001 = nullptr;
  // what refineByICP's single device visit brought back for selectBest: refined poses (n x 16) and their LCP scores
  std::vector<float> _scored_poses, _scored_lcp;
  std::vector<float> _s4_poses, _s4_lcp;   // runSuper4pcs' receive buffers, kept from frame to frame (1.3 MB: zero-filling fresh pages was ~0.1 ms a frame)
  float _scored_lcp_dist = 0.f, _scored_lcp_angle = 0.f;
  std::map<std::string, hop_mesh *> _meshes;
  std::vector<float> _obj_mesh_V;      // _obj_mesh (model frame), kept for the renderer
  std::vector<int32_t> _obj_mesh_F;
  void check(int rc, const char *what);
};
