#include <cstdlib>
#include "ConfigParser.h"

#include <cstdio>
#include <iostream>
#include <vector>

#include "cloud.h"

ConfigParser::ConfigParser(std::string cfg_file) {
  yml = miniyaml::LoadFile(cfg_file);
  parseYMLFile(cfg_file);
}

static bool fill44(const std::vector<float> &d, Mat4f &M) {
  if (d.size() < 16) return false;
  for (int i = 0; i < 16; ++i) M(i / 4, i % 4) = d[i];
  return true;
}

void ConfigParser::parseYMLFile(std::string) {
  {
    std::vector<float> K = yml["cam_K"].as<std::vector<float>>();
    if (K.size() < 9) { printf("config: cam_K needs 9 numbers (got %d)\n", (int)K.size()); exit(1); }
    for (int i = 0; i < 9; i++) cam_intrinsic(i / 3, i % 3) = K[i];
  }
  endeffector2global.setIdentity();
  endeffector2global(0, 0) = 0; endeffector2global(0, 1) = 1; endeffector2global(1, 0) = -1; endeffector2global(1, 1) = 0;

  // cam_in_world is simulation-only; the reference exits when the file is missing (ConfigParser.cpp:61-67), which makes its
  // shipped config unusable without the simulation data -- tolerated here (identity + a note)
  cam_in_world.setIdentity();
  if (yml.has("cam_in_world")) {
    cam_in_world_file = yml["cam_in_world"].as<std::string>();
    std::vector<float> d;
    if (!parsePoseTxt(cam_in_world_file, d) || !fill44(d, cam_in_world)) printf("cam_in_world not available, using identity\n");
  }
  {
    std::vector<float> data = yml["cam1_in_leftarm"].as<std::vector<float>>();   // xyz, q(xyzw)
    if (data.size() < 7) { printf("config: cam1_in_leftarm needs 7 numbers, xyz + quaternion xyzw (got %d)\n", (int)data.size()); exit(1); }
    cam1_in_leftarm.setIdentity();
    cam1_in_leftarm(0, 3) = data[0]; cam1_in_leftarm(1, 3) = data[1]; cam1_in_leftarm(2, 3) = data[2];
    quat_to_rot(data[6], data[3], data[4], data[5], cam1_in_leftarm);
  }
  {
    std::vector<float> data;
    if (!parsePoseTxt(yml["palm_in_baselink"].as<std::string>(), data) || !fill44(data, palm_in_baselink)) { printf("palm_in_baselink unreadable\n"); exit(1); }
  }
  {
    std::vector<float> data;
    if (!parsePoseTxt(yml["leftarm_in_base"].as<std::string>(), data) || !fill44(data, leftarm_in_base)) { printf("leftarm_in_base unreadable\n"); exit(1); }
  }
  {
    std::vector<float> data = yml["handbase_in_palm"].as<std::vector<float>>();
    if (!fill44(data, handbase_in_palm)) { printf("handbase_in_palm needs 16 values\n"); exit(1); }
  }
  rgb_path = yml["rgb_path"].as<std::string>(std::string());
  depth_path = yml["depth_path"].as<std::string>();
  object_model_path = yml["object_model_path"].as<std::string>();
  object_mesh_path = yml["object_mesh_path"].as<std::string>(std::string());

  leaf_size = yml["down_sample"]["leaf_size"].as<float>(0.005f);
  radius = yml["remove_noise"]["radius"].as<float>(0.f);
  min_number = yml["remove_noise"]["min_number"].as<float>(0.f);
  super4pcs_sample_size = yml["super4pcs_sample_size"].as<float>(100.f);
  super4pcs_overlap = yml["super4pcs_overlap"].as<float>(0.2f);
  super4pcs_delta = yml["super4pcs_delta"].as<float>(0.003f);
  super4pcs_max_normal_difference = yml["super4pcs_max_normal_difference"].as<float>(-1.f);
  super4pcs_max_color_distance = yml["super4pcs_max_color_distance"].as<float>(-1.f);
  super4pcs_max_time_seconds = yml["super4pcs_max_time_seconds"].as<float>(1.f);
  pose_estimator_wrong_ratio = yml["pose_estimator_wrong_ratio"].as<float>(1.f);
  pose_estimator_high_confidence_thres = yml["pose_estimator_high_confidence_thres"].as<float>(0.8f);

  const miniyaml::Node &b = yml["b200"];
  b200_device = b["device"].as<int>(0);
  b200_hand_grid = b["hand_grid"].as<int>(4096);
  b200_max_hypotheses = b["max_hypotheses"].as<int>(20000);
}
