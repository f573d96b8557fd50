// main_realdata_auto -- the reference's entry point (src/perception/src/app/main_realdata_auto.cpp:13-222) on top of libhop:
//     main_realdata_auto <config.yaml>
// Same configuration schema, same stage order, same outputs (out_dir/{best.obj, scene_normals.ply, hand.ply, model2scene.txt}; stdout
// "best tf:").  With a hand model (urdf_path + Hand.<link>.cloud readable) the whole of main_realdata_auto.cpp:54-181 runs:
//     organized cloud + integral-image normals (hop_frame_organized) -> 1 mm voxels -> hand-base crop
//     -> HandT42: setCurScene (handbaseICP), matchOneComponentPSO x4 (K1), adjustHandHeight, makeHandCloud
//     -> removeSurroundingPointsAndAssignProbability -> MLS normals (hop_cloud_mls) -> 3 mm voxels -> confidences from object1.
// What differs from the reference, by necessity of what ships with it or is installed here:
//   * without a hand model (the URDF / link clouds are an external download) the hand branch is skipped with a note: the device front
//     end hop_frame_to_scene crops the scene, takes radius-PCA normals (the MLS step belongs to the hand branch) and every point keeps
//     confidence 1;
//   * the two PCL normal estimators are restated from PCL 1.9.1 (tests/test_gpu_normals.py: bit-exact / 1e-5 against
//     oracle/hop_oracle_frame.c, which itself cannot be pinned against PCL here);
//   * best.obj: the reference transforms the object MESH (pcl::io::loadOBJFile + saveOBJFile); here the transformed 1 mm model cloud is
//     written as "v" lines when object_mesh_path is unreadable, the transformed mesh (vertices + faces) when it is;
//   * ppf_path: a table in libhop's portable format is loaded when present, otherwise it is built from the model on the fly
//     (what the reference's computePPF app does offline; its Boost binary archive is platform specific and not shipped);
//   * rejectByCollisionOrNonTouching / rejectByRender (main_realdata_auto.cpp:199-200) run when object_mesh_path is readable; the
//     finger tests need Hand.<link>.convex_mesh / .mesh, without them only the scene-point test has inputs and the render is object only.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

#include "ConfigParser.h"
#include "Hand.h"
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
  // We treat Motoman left arm as world (main_realdata_auto.cpp:47-52)
  const Mat4f handbase_in_leftarm = cfg.leftarm_in_base.inverse() * cfg.palm_in_baselink * cfg.handbase_in_palm;
  const Mat4f handbase_in_cam = cfg.cam1_in_leftarm.inverse() * handbase_in_leftarm;

  hop_ctx *ctx = nullptr;
  if (hop_create(cfg.b200_device, &ctx) != HOP_OK) { fprintf(stderr, "%s\n", hop_last_error(nullptr)); return 1; }

  // the model's PPF table (main_realdata_auto.cpp:28-31 reads the Boost archive computePPF wrote; not shipped, platform specific)
  std::vector<int32_t> ppfs;
  const std::string ppf_path = cfg.yml["ppf_path"].as<std::string>(std::string());
  if (!ppf_path.empty() && loadPPFTable(ppf_path, ppfs)) printf("loaded %d PPF keys from %s\n", (int)(ppfs.size() / 4), ppf_path.c_str());
  else {   // what computePPF.cpp:56-107 does offline: all point pairs of the 5 mm model (all-pairs kernel + device sort / unique)
    int32_t nk = 0;
    if (hop_ppf_table_build(ctx, model.xyz.data(), model.nrm.data(), (int)model.size(), nullptr, 0, &nk) != HOP_OK) { fprintf(stderr, "%s\n", hop_last_error(ctx)); return 1; }
    ppfs.resize(4 * (size_t)nk);
    if (hop_ppf_table_build(ctx, model.xyz.data(), model.nrm.data(), (int)model.size(), ppfs.data(), nk, &nk) != HOP_OK) { fprintf(stderr, "%s\n", hop_last_error(ctx)); return 1; }
    printf("built %d PPF keys from the model (%d points)\n", (int)nk, (int)model.size());
  }

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
  const bool have_hand = !urdf.empty() && file_exists(urdf);
  if (!have_hand) printf("hand model not available (urdf_path): hand-state search and hand-point removal skipped, confidence = 1\n");

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
    Cloud object_segment, cloud_withouthand_raw, hand_cloud_cam;
    HandState hand_state;
    hand_state._handbase_in_cam = handbase_in_cam;
    auto dl = [&](hop_cloud *d, Cloud &c) {
      const int n = hop_cloud_size(d);
      c.xyz.resize(3 * (size_t)n); c.nrm.resize(3 * (size_t)n); c.conf.resize(n);
      if (hop_cloud_download(ctx, d, c.xyz.data(), c.nrm.data(), c.conf.data()) != HOP_OK) { fprintf(stderr, "%s\n", hop_last_error(ctx)); exit(1); }
    };
    auto ok = [&](int rc, const char *what) { if (rc != HOP_OK) { fprintf(stderr, "%s: %s\n", what, hop_last_error(ctx)); exit(1); } };
    if (have_hand) {
      // ---- main_realdata_auto.cpp:54-96: organized cloud with integral-image normals, 1 mm voxels, crop in the hand-base frame ----
      hop_cloud *d_org = nullptr, *v = nullptr, *t = nullptr, *a = nullptr, *b = nullptr, *c = nullptr, *d_region = nullptr;
      ok(hop_frame_organized(ctx, depth_mm.data(), w, h, &fp, 0.02f, 10.f, &d_org), "hop_frame_organized");
      ok(hop_cloud_voxel_grid(ctx, d_org, 0.001f, &v), "voxel grid 1 mm");
      ok(hop_cloud_transform(ctx, v, cam_in_handbase.data(), &t), "transform");
      ok(hop_cloud_pass_through(ctx, t, 2, -0.12f, 0.05f, &a), "pass z");
      ok(hop_cloud_pass_through(ctx, a, 0, -0.25f, -0.07f, &b), "pass x");
      ok(hop_cloud_pass_through(ctx, b, 1, -0.2f, 0.2f, &c), "pass y");
      ok(hop_cloud_transform(ctx, c, cam_in_handbase_inv.data(), &d_region), "transform back");
      Cloud scene_organized, scene_rgb;
      dl(d_org, scene_organized); dl(d_region, scene_rgb);
      for (hop_cloud *x : {d_org, v, t, a, b, c, d_region}) hop_cloud_free(ctx, x);
      printf("scene in the hand region: %d points\n", (int)scene_rgb.size());
      if (scene_rgb.size() == 0) { printf("empty hand region\n"); exit(1); }
      // ---- :99-148: the hand ----
      Hand hand(&cfg, ctx);
      std::string err;
      if (!hand.parseURDF(urdf, &err)) { printf("parseURDF: %s\n", err.c_str()); exit(1); }
      hand.setCurScene(scene_organized, scene_rgb, handbase_in_cam);
      hand.makeHandCloud();
      const miniyaml::Node &hm = cfg.yml["hand_match"];
      const float f1_match = hm["finger1_min_match"].as<float>(5.f), f2_match = hm["finger2_min_match"].as<float>(5.f);
      const float f1_dist = hm["finger1_dist_thres"].as<float>(0.005f), f2_dist = hm["finger2_dist_thres"].as<float>(0.005f);
      const float f1_ang = hm["finger1_normal_angle"].as<float>(60.f), f2_ang = hm["finger2_normal_angle"].as<float>(60.f);
      const char *order_r[4] = {"finger_2_1", "finger_2_2", "finger_1_1", "finger_1_2"}, *order_l[4] = {"finger_1_1", "finger_1_2", "finger_2_1", "finger_2_2"};
      const char **order = cam_in_handbase(1, 3) > 0 ? order_r : order_l;   // cam on the right side of the hand first (:114-139)
      for (int f = 0; f < 2; ++f)
        if (hand.matchOneComponentPSO(order[2 * f], 0, 120, false, f1_dist, f1_ang, f1_match))
          hand.matchOneComponentPSO(order[2 * f + 1], 0, 90, true, f2_dist, f2_ang, f2_match);
      for (const char *nme : order_l) printf("tf_self %s %.9g %.9g\n", nme, hand._tf_self[nme](1, 1), hand._tf_self[nme](2, 1));
      hand.adjustHandHeight();
      hand.makeHandCloud();
      // ---- :144-181: hand-point removal with confidences, MLS normals, 3 mm voxels, confidences from the nearest point of object1 ----
      Cloud object1;
      const float near_dist = cfg.yml["near_hand_dist"].as<float>(0.003f);
      hand.removeSurroundingPointsAndAssignProbability(scene_rgb, object1, near_dist * near_dist);
      cloud_withouthand_raw = object1;
      if (object1.size() == 0) { printf("no scene point left after the hand removal\n"); exit(1); }
      hop_cloud *d_obj1 = nullptr, *d_mls = nullptr, *d_seg = nullptr;
      ok(hop_cloud_upload(ctx, object1.xyz.data(), object1.nrm.data(), object1.conf.data(), (int)object1.size(), &d_obj1), "upload object1");
      ok(hop_cloud_mls(ctx, d_obj1, 0.003f, &d_mls), "hop_cloud_mls");
      ok(hop_cloud_voxel_grid(ctx, d_mls, 0.003f, &d_seg), "voxel grid 3 mm");
      Cloud seg;
      dl(d_seg, seg);
      removeAllNaNFromPointCloud(seg);
      std::vector<int32_t> nn_idx(seg.size());
      std::vector<float> nn_d2(seg.size());
      if (seg.size()) ok(hop_cloud_nn_query(ctx, d_mls, 0.006f, seg.xyz.data(), (int)seg.size(), nn_idx.data(), nn_d2.data()), "hop_cloud_nn_query");
      Cloud mls_host;
      dl(d_mls, mls_host);
      for (size_t i = 0; i < seg.size(); ++i) {
        float *n = &seg.nrm[3 * i];
        const float *q = &seg.xyz[3 * i];
        if ((0.f - q[0]) * n[0] + (0.f - q[1]) * n[1] + (0.f - q[2]) * n[2] < 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }   // flipNormalTowardsViewpoint(0, 0, 0)
        seg.conf[i] = nn_idx[i] >= 0 ? mls_host.conf[nn_idx[i]] : 0.f;                                                          // :166-177
      }
      for (hop_cloud *x : {d_obj1, d_mls, d_seg}) hop_cloud_free(ctx, x);
      object_segment = seg;
      hand_state = hand.state();
      transformPointCloudWithNormals(hand._hand_cloud, hand_cloud_cam, hand._handbase_in_cam);   // hand.ply (:215-217)
      printf("hand removal: %d -> %d points, object segment %d points\n", (int)scene_rgb.size(), (int)object1.size(), (int)object_segment.size());
    } else {
    hop_cloud *d_segment = nullptr;
    int32_t counts[5];
    if (hop_frame_to_scene(ctx, depth_mm.data(), w, h, &fp, &d_segment, counts) != HOP_OK) { fprintf(stderr, "%s\n", hop_last_error(ctx)); exit(1); }
    printf("scene in the hand region: %d points\n", (int)counts[2]);
    if (counts[2] == 0) { printf("empty hand region\n"); exit(1); }
    // host copy of the object segment: the Super4PCS planner replays the reference's RNG on the host, and scene_normals.ply is written from it
    dl(d_segment, object_segment);
    hop_cloud_free(ctx, d_segment);
    cloud_withouthand_raw = object_segment;
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
      const HandState &hand = hand_state;
      est.rejectByCollisionOrNonTouching(hand, cloud_withouthand_raw);
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
    {   // best.obj (:209-213): the object mesh moved by the pose; the 1 mm model cloud as bare vertices when no mesh is readable
      std::vector<float> mV; std::vector<int32_t> mF; std::string merr;
      if (!mesh_path.empty() && file_exists(mesh_path) && loadOBJMesh(mesh_path, mV, mF, &merr)) {
        for (size_t i = 0; i + 2 < mV.size(); i += 3) {
          const float x = mV[i], y = mV[i + 1], z = mV[i + 2];
          for (int r = 0; r < 3; ++r) mV[i + r] = model2scene(r, 0) * x + model2scene(r, 1) * y + model2scene(r, 2) * z + model2scene(r, 3);
        }
        saveOBJMesh(out_dir + "/best.obj", mV, mF);
      } else {
        Cloud model_viz;
        transformPointCloudWithNormals(model001, model_viz, model2scene);
        saveOBJVertices(out_dir + "/best.obj", model_viz);
      }
    }
    savePLYFile(out_dir + "/scene_normals.ply", object_segment);
    if (have_hand) savePLYFile(out_dir + "/hand.ply", hand_cloud_cam);
    savePoseTxt(out_dir + "/model2scene.txt", model2scene);
  }
  }
  hop_destroy(ctx);
  return 0;
}
