// Hand.h -- the hand model of the reference (src/perception/include/Hand.h:29-92: class Hand / HandT42) for the methods on and next
// to the hot path; every heavy step is a call into libhop's C ABI (include/hop_c_api.h).
//   setCurScene              Hand.cpp:279-334  -> hop_cloud_* filters (+ handbaseICP :677-763 -> hop_icp_refine)
//   matchOneComponentPSO     Hand.cpp:603-672  -> hop_hand_overlap over a dense grid of joint angles (K1)
//   adjustHandHeight         Hand.cpp:999-1051 -> hop_adjust_hand_height
//   makeHandCloud            Hand.cpp:537-558
//   removeSurroundingPointsAndAssignProbability  Hand.cpp:781-888 -> hop_remove_hand_points
//   getTFHandBase            Hand.cpp:505-523;  FingerProperty  Hand.cpp:182-250
// The reference builds the kinematic tree from a URDF (Hand.cpp:375-502, pugixml) that does not ship with it; here the links are
// added explicitly (addComponent: cloud in the link frame, pose in the parent link) -- the same data parseURDF hands to addComponent.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "ConfigParser.h"
#include "PoseEstimator.h"   // HandState
#include "cloud.h"
#include "hop_c_api.h"

class FingerProperty {   // Hand.cpp:182-250: bounding box + per-z-bin extents of a link cloud
 public:
  FingerProperty() {}
  FingerProperty(const Cloud &model, int num_division);
  int getBinAlongZ(float z) const;
  int _num_division = 0;
  float _min_x = 0, _min_y = 0, _min_z = 0, _max_x = 0, _max_y = 0, _max_z = 0, _stride_z = 0;
  std::vector<float> _hist_alongz;   // 6 x num_division, row-major: min x/y/z, max x/y/z per bin
};

class Hand {
 public:
  Hand(ConfigParser *cfg1, hop_ctx *ctx);
  ~Hand();
  // Hand::addComponent (Hand.cpp:526-535): the link cloud is downsampled to 5 mm like the reference
  void addComponent(const std::string &name, const std::string &parent_name, const Cloud &cloud, const Mat4f &tf_in_parent);
  // Hand::parseURDF (Hand.cpp:375-502): every <link> (except the rails) becomes a component -- cloud from Hand.<name>.cloud scaled by
  // the visual mesh scale and moved by the visual origin, pose in the parent from the <joint> whose child it is.  Convex meshes
  // (Hand.<name>.convex_mesh, same scale and origin) are collected for PoseEstimator::registerHandMesh.  False when a file is missing.
  bool parseURDF(const std::string &urdf_path, std::string *err = nullptr);
  struct LinkMesh { std::vector<float> V; std::vector<int32_t> F; };
  std::map<std::string, LinkMesh> _convex_meshes, _meshes;   // link frame
  void getTFHandBase(std::string cur_name, Mat4f &tf_in_handbase) const;
  // scene_organized: the whole frame's cloud (camera frame, normals) for handbaseICP; scene_hand_region: the cropped hand region
  void setCurScene(const Cloud &scene_organized, const Cloud &scene_hand_region, const Mat4f &handbase_in_cam);
  void handbaseICP(const Cloud &scene_organized);
  bool matchOneComponentPSO(std::string model_name, float min_angle, float max_angle, bool use_normal, float dist_thres, float normal_angle_thres,
                            float least_match);
  void adjustHandHeight();
  void makeHandCloud();
  void removeSurroundingPointsAndAssignProbability(const Cloud &scene, Cloud &scene_out, float dist_thres);
  HandState state() const;   // what PoseEstimator::rejectBy* read

  std::map<std::string, bool> _component_status;
  std::map<std::string, Mat4f> _tf_self, _tf_in_parent;
  std::map<std::string, std::string> _parent_names;
  std::map<std::string, Cloud> _clouds;
  std::map<std::string, FingerProperty> _finger_properties;
  Mat4f _handbase_in_cam;
  Cloud _hand_cloud;                          // hand-base frame
  std::map<std::string, Cloud> _link_clouds_in_handbase;   // what the reference's _kdtrees index (makeHandCloud)
  int n_states = 4096;                        // joint angles evaluated per search (the reference's swarm visits 64)
  double objval = 0;

 private:
  ConfigParser *cfg;
  hop_ctx *ctx;
  std::map<std::string, hop_cloud *> d_clouds;              // link clouds, link frame
  hop_cloud *d_scene_hand_region = nullptr, *d_removed_noise = nullptr, *d_remove_swivel = nullptr;   // hand-base frame (_pso_args)
  void check(int rc, const char *what) const;
  void free_scene();
};
