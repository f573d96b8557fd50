// ConfigParser.h -- same public members and YAML schema as the reference's ConfigParser
// (src/perception/include/ConfigParser.h:7-33, src/perception/src/ConfigParser.cpp:30-137; config_autodataset.yaml), read with
// the built-in YAML-subset reader instead of yaml-cpp + `rosparam load` (neither exists here).  Optional new keys live
// under a new top-level map `b200:` (gpus, hand_grid, max_hypotheses) so the reference's own config files load unchanged.
#pragma once
#include <string>

#include "mat.h"
#include "mini_yaml.h"

class ConfigParser {
 public:
  explicit ConfigParser(std::string cfg_file);
  void parseYMLFile(std::string filepath);

  miniyaml::Node yml;
  Mat3f cam_intrinsic;
  Mat4f cam_in_world, cam1_in_leftarm, palm_in_baselink, leftarm_in_base, handbase_in_palm, endeffector2global;
  std::string rgb_path, depth_path, object_model_path, object_mesh_path, cam_in_world_file;
  float leaf_size = 0, radius = 0, min_number = 0;
  float super4pcs_sample_size = 100, super4pcs_overlap = 0.2f, super4pcs_delta = 0.003f, super4pcs_max_normal_difference = -1,
        super4pcs_max_color_distance = -1, super4pcs_max_time_seconds = 1;
  float pose_estimator_wrong_ratio = 1, pose_estimator_high_confidence_thres = 0.8f;
  float gripper_min_dist = 0;
  // b200: extensions (all optional)
  int b200_device = 0, b200_hand_grid = 4096, b200_max_hypotheses = 20000;
};
