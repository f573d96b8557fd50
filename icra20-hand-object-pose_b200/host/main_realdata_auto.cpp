// main_realdata_auto -- the reference's entry point (src/perception/src/app/main_realdata_auto.cpp:13-222) on top of libhop:
//     main_realdata_auto <config.yaml>
// Same configuration schema, same stage order, same outputs (out_dir/{best.obj, scene_normals.ply, model2scene.txt}; stdout
// "best tf:").  What differs from the reference, by necessity of what ships with it:
//   * the hand model (urdf_path, Hand.<link>.{mesh,cloud}) is an external download; when it is absent the hand-state search and
//     the hand-point removal are skipped with a note and every cropped scene point keeps confidence 1 (the device side of
//     that search, hop_hand_overlap, is exercised by the test-suite on synthetic links);
//   * ppf_path: a table in libhop's portable format is loaded when present, otherwise it is built from the model on the fly
//     (what the reference's computePPF app does offline);
//   * rejectByCollisionOrNonTouching runs (main_realdata_auto.cpp:199) when object_mesh_path is readable; without the hand model only
//     its first test (a scene point deep inside the placed object) has inputs, the finger tests find no enabled link;
//     rejectByRender (:200) runs right after it on the software rasteriser (object only when there is no hand model).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

#include "ConfigParser.h"
#include "PoseEstimator.h"
#include "cloud.h"
#include "ppf_table.h"

static bool file_exists(const std::string &p) { std::ifstream f(p); return (bool)f; }

int main(int argc, char **argv) {
  if (argc < 2) { std::cout << "usage: main_realdata_auto <config.yaml>" << std::endl; return 1; }
  const std::string config_dir = argv[1];
  std::cout << "Using config file: " << config_dir << std::endl;
  ConfigParser cfg(config_dir);

  Cloud model, model001;
  {
    Cloud raw;
    std::string err;
    if (!loadPLYFile(cfg.object_model_path, raw, &err) || raw.size() == 0) { printf("cannot load object model: %s\n", err.c_str()); return 1; }
    if (!raw.has_normals()) { printf("object model has no normals\n"); return 1; }
    downsamplePointCloud(raw, model001, 0.001f);
    downsamplePointCloud(model001, model, 0.005f);
  }
  {
    float mn[3], mx[3];
    getMinMax3D(model001, mn, mx);
    cfg.gripper_min_dist = 0.8 * std::min(std::min(std::abs(mn[0] - mx[0]), std::abs(mn[1] - mx[1])), std::abs(mn[2] - mx[2]));
  }
  std::vector<int32_t> ppfs;
  const std::string ppf_path = cfg.yml["ppf_path"].as<std::string>(std::string());
  if (!ppf_path.empty() && loadPPFTable(ppf_path, ppfs)) printf("loaded %d PPF keys from %s\n", (int)(ppfs.size() / 4), ppf_path.c_str());
  else { ppfs = buildPPFTable(model); printf("built %d PPF keys from the model (%d points)\n", (int)(ppfs.size() / 4), (int)model.size()); }

  // We treat Motoman left arm as world (main_realdata_auto.cpp:47-52)
  const Mat4f handbase_in_leftarm = cfg.leftarm_in_base.inverse() * cfg.palm_in_baselink * cfg.handbase_in_palm;
  const Mat4f handbase_in_cam = cfg.cam1_in_leftarm.inverse() * handbase_in_leftarm;

  hop_ctx *ctx = nullptr;
  if (hop_create(cfg.b200_device, &ctx) != HOP_OK) { fprintf(stderr, "%s\n", hop_last_error(nullptr)); return 1; }

  // the frame's front end on the device (hop_frame_to_scene = cloud.cpp's frameToObjectSegment, main_realdata_auto.cpp:54-96,
  // 144-181): depth PNG -> back-projection -> 1 mm voxels -> hand-base crop -> normals over 3 mm -> 3 mm voxels -> normals
  // towards the camera
  std::vector<uint16_t> depth_mm;
  int w = 0, h = 0;
  {
    std::string err;
    if (!readPNG16(cfg.depth_path, depth_mm, w, h, &err)) { printf("readDepthImage: %s\n", err.c_str()); hop_destroy(ctx); return 1; }
  }
  std::vector<float> depth_meters((size_t)w * h);   // Utils::readDepthImage (Utils.cpp:36-55): metres, 0 outside 0.1 .. 2 m
  for (size_t i = 0; i < depth_meters.size(); ++i) {
    const float d = (float)((float)depth_mm[i] * 0.001);   // (float)depthShort * SR300_DEPTH_UNIT: the unit is a double literal (Utils.h:107)
    depth_meters[i] = (d > 2.0 || d < 0.1) ? 0.f : d;
  }
  const Mat4f cam_in_handbase = handbase_in_cam.inverse();
  const Mat4f cam_in_handbase_inv = cam_in_handbase.inverse();
  hop_frame_params fp;
  hop_default_frame_params(&fp);
  fp.fx = cfg.cam_intrinsic(0, 0); fp.fy = cfg.cam_intrinsic(1, 1); fp.cx = cfg.cam_intrinsic(0, 2); fp.cy = cfg.cam_intrinsic(1, 2);
  std::memcpy(fp.cam_in_handbase, cam_in_handbase.data(), 64);
  std::memcpy(fp.handbase_in_cam, cam_in_handbase_inv.data(), 64);
  const std::string urdf = cfg.yml["urdf_path"].as<std::string>(std::string());
  if (urdf.empty() || !file_exists(urdf))
    printf("hand model not available (urdf_path): hand-state search and hand-point removal skipped, confidence = 1\n");

  // optional second argument: process the frame that many times and report the stage times of the last pass (steady state:
  // the first pass of a process also pays module loading and the first allocations)
  const int repeat = argc >= 3 ? std::max(1, atoi(argv[2])) : 1;
  {   // (the estimator's device clouds must be gone before the context)
  PoseEstimator est(&cfg, model, model001, ctx);
  const std::string mesh_path = cfg.yml["object_mesh_path"].as<std::string>(std::string());
  const bool use_physics = cfg.yml["pose_estimator_use_physics"].as<bool>(true) && !mesh_path.empty() && file_exists(mesh_path) &&
                           est.registerMesh(mesh_path, "object", Mat4f());   // main_realdata_auto.cpp:39
  if (!use_physics) printf("physics pruning off (pose_estimator_use_physics / object_mesh_path)\n");
  const std::string out_dir = cfg.yml["out_dir"].as<std::string>();
  for (int pass = 0; pass < repeat; ++pass) {
    typedef std::chrono::steady_clock Clock;
    auto ms = [](Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const Clock::time_point t0 = Clock::now();
    hop_cloud *d_segment = nullptr;
    int32_t counts[5];
    if (hop_frame_to_scene(ctx, depth_mm.data(), w, h, &fp, &d_segment, counts) != HOP_OK) { fprintf(stderr, "%s\n", hop_last_error(ctx)); exit(1); }
    printf("scene in the hand region: %d points\n", (int)counts[2]);
    if (counts[2] == 0) { printf("empty hand region\n"); exit(1); }
    // host copy of the object segment: the Super4PCS planner replays the reference's RNG on the host, and scene_normals.ply is written from it
    Cloud object_segment;
    {
      const int n = hop_cloud_size(d_segment);
      object_segment.xyz.resize(3 * (size_t)n); object_segment.nrm.resize(3 * (size_t)n); object_segment.conf.resize(n);
      if (hop_cloud_download(ctx, d_segment, object_segment.xyz.data(), object_segment.nrm.data(), object_segment.conf.data()) != HOP_OK) {
        fprintf(stderr, "%s\n", hop_last_error(ctx)); exit(1);
      }
      hop_cloud_free(ctx, d_segment);
    }
    est.setCurScene(object_segment);
    const Clock::time_point t1 = Clock::now();
    const bool succeed = est.runSuper4pcs(ppfs);
    if (!succeed) {
      printf("No pose found...\n");
      savePoseTxt(out_dir + "/model2scene.txt", Mat4f());
      exit(1);
    }
    const Clock::time_point t2 = Clock::now();
    est.clusterPoses(30, 0.015, true);
    const Clock::time_point t3 = Clock::now();
    est.refineByICP();
    const Clock::time_point t4 = Clock::now();
    est.clusterPoses(5, 0.003, false);
    const Clock::time_point t4a = Clock::now();
    Clock::time_point t4b = t4a, t4c = t4a;
    if (use_physics) {   // main_realdata_auto.cpp:199; no hand model -> no enabled link, no hand cloud
      HandState hand;
      hand._handbase_in_cam = handbase_in_cam;
      est.rejectByCollisionOrNonTouching(hand, object_segment);
      if (est._pose_hypos.empty()) { printf("No pose found...\n"); savePoseTxt(out_dir + "/model2scene.txt", Mat4f()); exit(1); }
      t4b = Clock::now();
      est.rejectByRender(cfg.yml["pose_estimator_wrong_ratio"].as<float>(0.f), hand, depth_meters, w, h);   // main_realdata_auto.cpp:200
      t4c = Clock::now();
    }
    PoseHypo best(-1);
    est.selectBest(best);
    const Clock::time_point t5 = Clock::now();
    printf("timing_ms front_end %.3f super4pcs %.3f cluster %.3f icp %.3f cluster_select %.3f physics %.3f render %.3f total %.3f\n", ms(t0, t1), ms(t1, t2),
           ms(t2, t3), ms(t3, t4), ms(t4, t4a) + ms(t4c, t5), ms(t4a, t4b), ms(t4b, t4c), ms(t0, t5));
    if (pass + 1 < repeat) continue;
    const Mat4f model2scene = best._pose;
    std::cout << "best tf:\n" << model2scene << "\n\n";
    Cloud model_viz;
    transformPointCloudWithNormals(model001, model_viz, model2scene);
    saveOBJVertices(out_dir + "/best.obj", model_viz);
    savePLYFile(out_dir + "/scene_normals.ply", object_segment);
    savePoseTxt(out_dir + "/model2scene.txt", model2scene);
  }
  }
  hop_destroy(ctx);
  return 0;
}
