#include "PoseEstimator.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

void PoseEstimator::check(int rc, const char *what) {
  if (rc == HOP_OK) return;
  fprintf(stderr, "%s failed (%d): %s\n", what, rc, hop_last_error(ctx));
  exit(1);
}

PoseEstimator::PoseEstimator(ConfigParser *cfg1, const Cloud &model, const Cloud &model001, hop_ctx *c) : cfg(cfg1), ctx(c), _model(model), _model001(model001) {
  check(hop_cloud_upload(ctx, _model.xyz.data(), _model.nrm.data(), nullptr, (int)_model.size(), &d_model), "upload model");
  check(hop_cloud_upload(ctx, _model001.xyz.data(), _model001.nrm.data(), nullptr, (int)_model001.size(), &d_model001), "upload model001");
  check(hop_cloud_hint_static(ctx, d_model, 1), "hint model"); check(hop_cloud_hint_static(ctx, d_model001, 1), "hint model001");   // loaded once: finer NN grids
}

PoseEstimator::~PoseEstimator() {
  hop_cloud_free(ctx, d_scene); hop_cloud_free(ctx, d_model); hop_cloud_free(ctx, d_model001);
  for (auto &m : _meshes) hop_mesh_free(ctx, m.second);
}

// ---- physics pruning (PoseEstimator.cpp:506-735) ------------------------------------------------------------------------
void PoseEstimator::registerMesh(const std::vector<float> &V, const std::vector<int32_t> &F, const std::string &name, const Mat4f &pose) {
  std::vector<float> Vt(V.size());
  for (size_t i = 0; i + 2 < V.size(); i += 3)   // SDFchecker::transformVertices: pose * [v; 1], float
    for (int r = 0; r < 3; ++r) Vt[i + r] = pose(r, 0) * V[i] + pose(r, 1) * V[i + 1] + pose(r, 2) * V[i + 2] + pose(r, 3);
  hop_mesh *m = nullptr;
  check(hop_mesh_upload(ctx, Vt.data(), (int)(Vt.size() / 3), F.data(), (int)(F.size() / 3), &m), "hop_mesh_upload");
  auto it = _meshes.find(name);
  if (it != _meshes.end()) hop_mesh_free(ctx, it->second);
  _meshes[name] = m;
  if (name == "object") { _obj_mesh_V = Vt; _obj_mesh_F = F; }
}

void PoseEstimator::rejectByRender(float /*projection_thres: unused by the reference too*/, const HandState &hand, const std::vector<float> &depth_meters,
                                   int width, int height) {
  printf("before projection check, #hypo=%d\n", (int)_pose_hypos.size());
  if (_pose_hypos.empty()) return;
  if (_obj_mesh_F.empty()) { printf("rejectByRender: no object mesh registered, skipped\n"); return; }
  hop_render_params p;
  hop_default_render_params(&p);
  p.fx = cfg->cam_intrinsic(0, 0); p.fy = cfg->cam_intrinsic(1, 1); p.cx = cfg->cam_intrinsic(0, 2); p.cy = cfg->cam_intrinsic(1, 2);
  p.width = width; p.height = height;
  p.roi_weight = cfg->yml["render_roi_weight"].as<float>(2.0f);
  p.keep_ratio = cfg->yml["render_keep_hypo"].as<float>(0.3f);
  hop_render_scene *scene = nullptr;
  check(hop_render_scene_create(ctx, &p, depth_meters.data(), hand.meshes_in_cam_V.data(), (int)(hand.meshes_in_cam_V.size() / 3),
                                hand.meshes_in_cam_F.data(), (int)(hand.meshes_in_cam_F.size() / 3), &scene), "hop_render_scene_create");
  const size_t n = _pose_hypos.size();
  std::vector<float> poses(16 * n), wrong(n);
  std::vector<int32_t> order(n);
  int32_t n_keep = 0;
  for (size_t i = 0; i < n; ++i) std::memcpy(&poses[16 * i], _pose_hypos[i]._pose.data(), 64);
  check(hop_reject_by_render(ctx, scene, _obj_mesh_V.data(), (int)(_obj_mesh_V.size() / 3), _obj_mesh_F.data(), (int)(_obj_mesh_F.size() / 3), poses.data(),
                             (int)n, wrong.data(), order.data(), &n_keep), "hop_reject_by_render");
  hop_render_scene_destroy(ctx, scene);
  std::vector<PoseHypo> kept;
  for (int i = 0; i < n_keep; ++i) { PoseHypo h = _pose_hypos[order[i]]; h._wrong_ratio = wrong[order[i]]; kept.push_back(h); }
  _pose_hypos.swap(kept);
  printf("after projection check, #hypo=%d\n", (int)_pose_hypos.size());
}

bool PoseEstimator::registerMesh(const std::string &mesh_dir, const std::string &name, const Mat4f &pose) {
  std::vector<float> V;
  std::vector<int32_t> F;
  std::string err;
  if (!loadOBJMesh(mesh_dir, V, F, &err)) { printf("registerMesh(%s): %s\n", name.c_str(), err.c_str()); return false; }
  registerMesh(V, F, name, pose);
  return true;
}

void PoseEstimator::rejectByCollisionOrNonTouching(const HandState &hand, const Cloud &cloud_withouthand_raw) {
  if (cfg->yml["pose_estimator_use_physics"].as<bool>(true) == false) { printf("Not using physics\n"); return; }
  if (_pose_hypos.empty()) return;
  auto obj = _meshes.find("object");
  if (obj == _meshes.end()) { printf("rejectByCollisionOrNonTouching: no object mesh registered, skipped\n"); return; }
  static const char *names[4] = {"finger_1_1", "finger_1_2", "finger_2_1", "finger_2_2"};
  hop_collision_params p;
  const Mat4f cam2handbase = hand._handbase_in_cam.inverse();
  std::memcpy(p.cam2handbase, cam2handbase.data(), 64);
  {   // the constructor's _model_center_init, _smallest_dim, _ob_diameter (PoseEstimator.cpp:12-20): from model001
    float mn[3], mx[3];
    getMinMax3D(_model001, mn, mx);
    double c[3] = {0, 0, 0};
    for (size_t i = 0; i < _model001.size(); ++i) for (int k = 0; k < 3; ++k) c[k] += _model001.xyz[3 * i + k];
    for (int k = 0; k < 3; ++k) p.model_center[k] = (float)(c[k] / std::max<size_t>(_model001.size(), 1));
    const float ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
    const float smallest = std::min(std::min(ex, ey), ez);
    p.ob_diameter = std::sqrt(ex * ex + ey * ey + ez * ez);
    p.collision_dist = std::min(-smallest * cfg->yml["collision_thres"].as<float>(0.4f), -0.007f);
    p.inside_ob_dist = std::min(-smallest / 5, -0.01f);
  }
  p.non_touch_dist = cfg->yml["non_touch_dist"].as<float>(0.01f);
  p.collision_finger_dist = -cfg->yml["collision_finger_dist"].as<float>(0.012f);
  p.collision_finger_volume_ratio = cfg->yml["collision_finger_volume_ratio"].as<float>(0.25f);
  const hop_mesh *fm[4];
  hop_cloud *fc[4];
  for (int k = 0; k < 4; ++k) {
    auto st = hand._component_status.find(names[k]);
    p.finger_status[k] = st != hand._component_status.end() && st->second;
    auto m = _meshes.find(names[k]);
    fm[k] = m == _meshes.end() ? nullptr : m->second;
    fc[k] = nullptr;
    auto c = hand.finger_clouds.find(names[k]);
    if (p.finger_status[k] && c != hand.finger_clouds.end() && c->second.size() > 0)
      check(hop_cloud_upload(ctx, c->second.xyz.data(), nullptr, nullptr, (int)c->second.size(), &fc[k]), "upload finger cloud");
  }
  // the scene without the hand: into the hand-base frame, then the 5 mm VoxelGrid (PoseEstimator.cpp:556-559)
  hop_cloud *d_wo = nullptr, *d_hand = nullptr;
  if (cloud_withouthand_raw.size() > 0) {
    Cloud hb, ds;
    if (cloud_withouthand_raw.has_normals()) transformPointCloudWithNormals(cloud_withouthand_raw, hb, cam2handbase);
    else {
      hb = cloud_withouthand_raw;
      for (size_t i = 0; i < hb.size(); ++i) {
        const float x = hb.xyz[3 * i], y = hb.xyz[3 * i + 1], z = hb.xyz[3 * i + 2];
        for (int r = 0; r < 3; ++r) hb.xyz[3 * i + r] = cam2handbase(r, 0) * x + cam2handbase(r, 1) * y + cam2handbase(r, 2) * z + cam2handbase(r, 3);
      }
    }
    downsamplePointCloud(hb, ds, 0.005f);
    if (ds.size() > 0) check(hop_cloud_upload(ctx, ds.xyz.data(), nullptr, nullptr, (int)ds.size(), &d_wo), "upload scene without hand");
  }
  if (hand._hand_cloud.size() > 0)
    check(hop_cloud_upload(ctx, hand._hand_cloud.xyz.data(), nullptr, nullptr, (int)hand._hand_cloud.size(), &d_hand), "upload hand cloud");
  const size_t n = _pose_hypos.size();
  std::vector<float> poses(16 * n);
  std::vector<int32_t> keep(n), reason(n);
  for (size_t i = 0; i < n; ++i) std::memcpy(&poses[16 * i], _pose_hypos[i]._pose.data(), 64);
  printf("collision_dist=%f, non_touch_dist=%f\n", p.collision_dist, p.non_touch_dist);
  check(hop_reject_by_collision(ctx, obj->second, fm, fc, d_wo, d_hand, d_model, poses.data(), (int)n, &p, keep.data(), reason.data(), nullptr),
        "hop_reject_by_collision");
  std::vector<PoseHypo> kept;
  int why[7] = {0, 0, 0, 0, 0, 0, 0};
  for (size_t i = 0; i < n; ++i) { why[reason[i]]++; if (keep[i]) kept.push_back(_pose_hypos[i]); }   // in order (the reference's survivors come out in OpenMP order)
  printf("physics pruning: %d of %d hypotheses kept (scene-in-object %d, hand-in-object %d, finger collision %d, side not touching %d, object-in-finger %d/%d)\n",
         why[0], (int)n, why[1], why[2], why[3], why[4], why[5], why[6]);
  _pose_hypos.swap(kept);
  for (int k = 0; k < 4; ++k) hop_cloud_free(ctx, fc[k]);
  hop_cloud_free(ctx, d_wo); hop_cloud_free(ctx, d_hand);
}

void PoseEstimator::setCurScene(const Cloud &object_segment) {
  const float thres = cfg->pose_estimator_high_confidence_thres;
  _scene_high_confidence.clear();
  for (size_t i = 0; i < object_segment.size(); ++i)
    if (object_segment.conf[i] >= thres) _scene_high_confidence.push(&object_segment.xyz[3 * i], &object_segment.nrm[3 * i], object_segment.conf[i]);
  const Cloud &s = _scene_high_confidence;
  if (!d_scene) check(hop_cloud_upload(ctx, s.xyz.data(), s.nrm.data(), s.conf.data(), (int)s.size(), &d_scene), "upload scene");
  else check(hop_cloud_update(ctx, d_scene, s.xyz.data(), s.nrm.data(), s.conf.data(), (int)s.size()), "update scene");
}

bool PoseEstimator::runSuper4pcs(const std::vector<int32_t> &ppf_keys) {
  hop_s4pcs_options o;
  hop_default_s4pcs_options(&o);
  o.sample_size = (int)cfg->yml["super4pcs_sample_size"].as<float>(100.f);
  o.overlap = cfg->yml["super4pcs_overlap"].as<float>(0.2f);
  o.delta = cfg->yml["super4pcs_delta"].as<float>(0.003f);
  o.dispersion = cfg->yml["super4pcs_dispersion"].as<float>(0.5f);
  o.success_quadrilaterals = cfg->yml["super4pcs_success_quadrilaterals"].as<int>(10);
  o.max_normal_difference = cfg->super4pcs_max_normal_difference;
  o.max_color_distance = cfg->super4pcs_max_color_distance;
  const Cloud &P = _scene_high_confidence;
  hop_s4pcs_plan *plan = nullptr;
  // (the planner's PPF-membership scans run on the device; the random draws it replays stay on the host)
  int rc = hop_s4pcs_plan_create_gpu(ctx, P.xyz.data(), P.nrm.data(), P.conf.data(), (int)P.size(), _model.xyz.data(), _model.nrm.data(), (int)_model.size(),
                                 ppf_keys.data(), (int)(ppf_keys.size() / 4), &o, &plan);
  if (rc != HOP_OK) { fprintf(stderr, "hop_s4pcs_plan_create failed (%d)\n", rc); return false; }
  const int cap = cfg->b200_max_hypotheses;   // the reference reserves 20000 (super4pcs.h:134)
  std::vector<float> &poses = _s4_poses, &lcp = _s4_lcp;
  if (poses.size() < 16 * (size_t)cap) { poses.resize(16 * (size_t)cap); lcp.resize((size_t)cap); }
  int32_t n = 0;
  rc = hop_super4pcs_run(ctx, plan, poses.data(), lcp.data(), cap, &n);
  if (rc == HOP_OK && n > cap) {
    // the reference only reserve()s 20000 and keeps every hypothesis (super4pcs.h:134): fetch them all -- dropping the tail in
    // (trial, quadrilateral) emission order could lose the best hypotheses of the later trials before clusterPoses sorts
    printf("runSuper4pcs: %d hypotheses exceed b200_max_hypotheses = %d, fetching all of them\n", (int)n, cap);
    poses.resize(16 * (size_t)n); lcp.resize((size_t)n);
    const int all = n;
    rc = hop_super4pcs_run(ctx, plan, poses.data(), lcp.data(), all, &n);
    n = std::min<int32_t>(n, all);
  } else n = std::min<int32_t>(n, cap);
  hop_s4pcs_plan_destroy(plan);
  check(rc, "hop_super4pcs_run");
  _pose_hypos.clear();
  for (int i = 0; i < n; ++i) {
    Mat4f T;
    std::memcpy(T.data(), &poses[16 * (size_t)i], 64);
    _pose_hypos.push_back(PoseHypo(T, i, lcp[i]));
  }
  return !_pose_hypos.empty();
}

void PoseEstimator::clusterPoses(float angle_diff, float dist_diff, bool assign_id) {
  printf("num original candidates = %d\n", (int)_pose_hypos.size());
  const int n = (int)_pose_hypos.size();
  if (n == 0) return;
  const std::string model_name = cfg->yml["model_name"].as<std::string>();
  const miniyaml::Node &s = cfg->yml["object_symmetry"][model_name];
  const float sym[3] = {s["x"].as<float>(360.f), s["y"].as<float>(360.f), s["z"].as<float>(360.f)};
  printf("object symmetry: x=%f, y=%f, z=%f\n", sym[0] / 180 * M_PI, sym[1] / 180 * M_PI, sym[2] / 180 * M_PI);
  std::vector<float> poses(16 * (size_t)n), scores(n);
  std::vector<int32_t> ids(n), keep(n);
  for (int i = 0; i < n; ++i) { std::memcpy(&poses[16 * (size_t)i], _pose_hypos[i]._pose.data(), 64); scores[i] = _pose_hypos[i]._lcp_score; ids[i] = _pose_hypos[i]._id; }
  int32_t nk = 0;
  // the device version decides identically; it pays from a few thousand hypotheses up
  if (n >= 2048) check(hop_cluster_poses_gpu(ctx, poses.data(), scores.data(), ids.data(), n, angle_diff, dist_diff, sym, keep.data(), &nk), "hop_cluster_poses_gpu");
  else check(hop_cluster_poses(poses.data(), scores.data(), ids.data(), n, angle_diff, dist_diff, sym, keep.data(), &nk), "hop_cluster_poses");
  std::vector<PoseHypo> out;
  for (int k = 0; k < nk; ++k) out.push_back(_pose_hypos[keep[k]]);
  _pose_hypos.swap(out);
  if (assign_id) for (size_t i = 0; i < _pose_hypos.size(); i++) _pose_hypos[i]._id = (int)i;
  printf("num of pose clusters: %d\n", (int)_pose_hypos.size());
}

void PoseEstimator::refineByICP() {
  const size_t n = std::min<size_t>(_pose_hypos.size(), 100);   // PoseEstimator.cpp:241
  _pose_hypos.resize(n);
  _scored_poses.clear(); _scored_lcp.clear();
  if (n == 0) return;
  std::vector<float> poses(16 * n), scores(n);
  for (size_t i = 0; i < n; ++i) std::memcpy(&poses[16 * i], _pose_hypos[i]._pose.data(), 64);
  hop_icp_params p;
  hop_default_icp_params(&p);   // 10 iterations, abs MSE 1e-6 (Utils.cpp:207-208; PoseEstimator.cpp:266)
  p.angle_deg = cfg->yml["icp_angle_thres"].as<float>(45.f);
  p.max_dist = cfg->yml["icp_dist_thres"].as<float>(0.01f);
  hop_lcp_params q;
  hop_default_lcp_params(&q);
  q.dist = cfg->yml["lcp"]["dist"].as<float>(0.001f);
  q.angle_deg = cfg->yml["lcp"]["normal_angle"].as<float>(10.f);
  // One visit to the device: the hypotheses go up once, are refined (K4) and scored (K5, what selectBest will ask for: the score
  // of a hypothesis does not depend on which others survive the pruning in between), and come back with one synchronisation.
  check(hop_refine_score_select(ctx, d_scene, d_model, d_model001, poses.data(), (int)n, &p, &q, 0, 0, poses.data(), scores.data(), nullptr, nullptr,
                                nullptr), "hop_refine_score_select");
  for (size_t i = 0; i < n; ++i) std::memcpy(_pose_hypos[i]._pose.data(), &poses[16 * i], 64);
  _scored_poses.swap(poses); _scored_lcp.swap(scores);
  _scored_lcp_dist = q.dist; _scored_lcp_angle = q.angle_deg;
}

void PoseEstimator::selectBest(PoseHypo &best_hypo) {
  const size_t n = _pose_hypos.size();
  if (n == 0) return;
  hop_lcp_params p;
  hop_default_lcp_params(&p);
  p.dist = cfg->yml["lcp"]["dist"].as<float>(0.001f);
  p.angle_deg = cfg->yml["lcp"]["normal_angle"].as<float>(10.f);
  std::vector<float> scores(n);
  // scores refineByICP already brought back: valid for a hypothesis whose pose is bit-identical to a refined one
  std::vector<size_t> todo;
  const bool cache_ok = !_scored_lcp.empty() && _scored_lcp_dist == p.dist && _scored_lcp_angle == p.angle_deg;
  for (size_t i = 0; i < n; ++i) {
    bool hit = false;
    if (cache_ok) {
      const size_t id = (size_t)_pose_hypos[i]._id;
      size_t k = id < _scored_lcp.size() && std::memcmp(&_scored_poses[16 * id], _pose_hypos[i]._pose.data(), 64) == 0 ? id : _scored_lcp.size();
      for (size_t j = 0; k == _scored_lcp.size() && j < _scored_lcp.size(); ++j)
        if (std::memcmp(&_scored_poses[16 * j], _pose_hypos[i]._pose.data(), 64) == 0) k = j;
      if (k < _scored_lcp.size()) { scores[i] = _scored_lcp[k]; hit = true; }
    }
    if (!hit) todo.push_back(i);
  }
  if (!todo.empty()) {
    std::vector<float> poses(16 * todo.size()), sc(todo.size());
    for (size_t t = 0; t < todo.size(); ++t) std::memcpy(&poses[16 * t], _pose_hypos[todo[t]]._pose.data(), 64);
    check(hop_lcp_score(ctx, d_scene, d_model001, poses.data(), (int)todo.size(), &p, 0, sc.data()), "hop_lcp_score");
    for (size_t t = 0; t < todo.size(); ++t) scores[todo[t]] = sc[t];
  }
  float best_lcp = 0;
  best_hypo = _pose_hypos[0];
  for (size_t i = 0; i < n; ++i) {
    _pose_hypos[i]._lcp_score = scores[i];
    if (scores[i] > best_lcp) { best_lcp = scores[i]; best_hypo = _pose_hypos[i]; }
  }
  printf("best hypo id=%d, lcp=%f\n", best_hypo._id, best_lcp);
}
