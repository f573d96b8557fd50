// hand_demo -- drives the C++ Hand class (Hand.h) the way main_realdata_auto.cpp:100-148 drives HandT42, on inputs read from a
// directory (the hand's URDF / link clouds do not ship with the reference; the test-suite writes synthetic ones):
//     hand_demo <config.yaml> <dir>
//   <dir>/links.txt            one link per line: name parent cloud.ply t00 t01 ... t33 (tf_in_parent, row-major); not read when the
//                              config names urdf_path + Hand.<link>.cloud (then Hand::parseURDF builds the hand like the reference)
//   <dir>/scene_organized.ply  the frame's cloud (camera frame, normals)       <dir>/scene_hand_region.ply  the cropped hand region
//   <dir>/handbase_in_cam.txt  4x4
// Prints the four link searches, the corrected hand base and the hand-point removal; writes <dir>/object1.ply (x y z nx ny nz confidence).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

#include "ConfigParser.h"
#include "Hand.h"
#include "cloud.h"

int main(int argc, char **argv) {
  if (argc < 3) { std::cout << "usage: hand_demo <config.yaml> <dir>" << std::endl; return 1; }
  ConfigParser cfg(argv[1]);
  const std::string dir = argv[2];
  hop_ctx *ctx = nullptr;
  if (hop_create(cfg.b200_device, &ctx) != HOP_OK) { fprintf(stderr, "%s\n", hop_last_error(nullptr)); return 1; }
  {
    Hand hand(&cfg, ctx);
    cfg.gripper_min_dist = cfg.yml["gripper_min_dist"].as<float>(0.03f);
    const std::string urdf_path = cfg.yml["urdf_path"].as<std::string>(std::string());
    if (!urdf_path.empty()) {   // the reference's way (Hand::Hand -> parseURDF, Hand.cpp:375-502): links, clouds and meshes named by the config
      std::string err;
      if (!hand.parseURDF(urdf_path, &err)) { printf("parseURDF: %s\n", err.c_str()); return 1; }
    }
    std::ifstream lf(dir + "/links.txt");
    std::string line;
    while (urdf_path.empty() && std::getline(lf, line)) {
      std::istringstream ss(line);
      std::string name, parent, ply;
      if (!(ss >> name >> parent >> ply)) continue;
      Mat4f T;
      for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) { float v; ss >> v; T(r, c) = v; }
      Cloud cl; std::string err;
      if (!loadPLYFile(dir + "/" + ply, cl, &err)) { printf("cannot load %s: %s\n", ply.c_str(), err.c_str()); return 1; }
      hand.addComponent(name, parent, cl, T);
    }
    Cloud organized, region;
    std::string err;
    if (!loadPLYFile(dir + "/scene_organized.ply", organized, &err) || !loadPLYFile(dir + "/scene_hand_region.ply", region, &err)) { printf("%s\n", err.c_str()); return 1; }
    std::vector<float> t;
    if (!parsePoseTxt(dir + "/handbase_in_cam.txt", t) || t.size() < 16) { printf("bad handbase_in_cam.txt\n"); return 1; }
    Mat4f hic; for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) hic(r, c) = t[4 * r + c];
    hand.setCurScene(organized, region, hic);
    std::cout << "handbase matched " << (hand._component_status["handbase"] ? 1 : 0) << "\n";
    const float f1_match = cfg.yml["hand_match"]["finger1_min_match"].as<float>(5.f), f2_match = cfg.yml["hand_match"]["finger2_min_match"].as<float>(5.f);
    const float f1_dist = cfg.yml["hand_match"]["finger1_dist_thres"].as<float>(0.005f), f2_dist = cfg.yml["hand_match"]["finger2_dist_thres"].as<float>(0.005f);
    const float f1_ang = cfg.yml["hand_match"]["finger1_normal_angle"].as<float>(60.f), f2_ang = cfg.yml["hand_match"]["finger2_normal_angle"].as<float>(60.f);
    const Mat4f cam_in_handbase = hic.inverse();
    const char *order_r[4] = {"finger_2_1", "finger_2_2", "finger_1_1", "finger_1_2"}, *order_l[4] = {"finger_1_1", "finger_1_2", "finger_2_1", "finger_2_2"};
    const char **order = cam_in_handbase(1, 3) > 0 ? order_r : order_l;   // main_realdata_auto.cpp:114-139
    for (int f = 0; f < 2; ++f) {
      const bool m = hand.matchOneComponentPSO(order[2 * f], 0, 120, false, f1_dist, f1_ang, f1_match);
      printf("match %s %d objval %.9g\n", order[2 * f], m ? 1 : 0, hand.objval);
      if (m) {
        const bool m2 = hand.matchOneComponentPSO(order[2 * f + 1], 0, 90, true, f2_dist, f2_ang, f2_match);
        printf("match %s %d objval %.9g\n", order[2 * f + 1], m2 ? 1 : 0, hand.objval);
      }
    }
    for (const char *n : order_l) std::cout << "tf_self " << n << " " << hand._tf_self[n](1, 1) << " " << hand._tf_self[n](2, 1) << "\n";
    hand.adjustHandHeight();
    hand.makeHandCloud();
    std::cout << "handbase_in_cam\n" << hand._handbase_in_cam << "\n";
    Cloud object1;
    const float near_dist = cfg.yml["near_hand_dist"].as<float>(0.003f);
    hand.removeSurroundingPointsAndAssignProbability(region, object1, near_dist * near_dist);
    std::cout << "hand cloud " << hand._hand_cloud.size() << " region " << region.size() << " object1 " << object1.size() << "\n";
    savePLYFile(dir + "/object1.ply", object1);
  }
  hop_destroy(ctx);
  return 0;
}
