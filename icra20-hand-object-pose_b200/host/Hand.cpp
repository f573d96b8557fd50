// Hand.cpp -- see Hand.h.  Host glue only: kinematics, thresholds and the reference's accept / reject rules; the clouds live on the device.
#include "Hand.h"

#include "urdf.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

// ---- FingerProperty (Hand.cpp:184-250) --------------------------------------------------------------------------------------
FingerProperty::FingerProperty(const Cloud &model, int num_division) : _num_division(num_division) {
  float mn[3], mx[3];
  getMinMax3D(model, mn, mx);
  _min_x = mn[0]; _min_y = mn[1]; _min_z = mn[2]; _max_x = mx[0]; _max_y = mx[1]; _max_z = mx[2];
  _stride_z = (_max_z - _min_z) / num_division;
  _hist_alongz.assign(6 * (size_t)num_division, 0.f);
  for (int b = 0; b < num_division; ++b)
    for (int r = 0; r < 6; ++r) _hist_alongz[r * num_division + b] = r < 3 ? FLT_MAX : -FLT_MAX;
  std::vector<char> changed(num_division, 0);
  for (size_t i = 0; i < model.size(); ++i) {
    const int b = getBinAlongZ(model.xyz[3 * i + 2]);
    for (int k = 0; k < 3; ++k) {
      float &lo = _hist_alongz[k * num_division + b], &hi = _hist_alongz[(3 + k) * num_division + b];
      lo = std::min(lo, model.xyz[3 * i + k]); hi = std::max(hi, model.xyz[3 * i + k]);
    }
    changed[b] = 1;
  }
  for (int i = 0; i < num_division; ++i) {   // an untouched bin takes the next touched one (:213-225)
    if (changed[i]) continue;
    for (int j = i + 1; j < num_division; ++j)
      if (changed[j]) { for (int r = 0; r < 6; ++r) _hist_alongz[r * num_division + i] = _hist_alongz[r * num_division + j]; changed[i] = 1; break; }
  }
  if (!changed[num_division - 1])
    for (int i = num_division - 2; i >= 0; --i)
      if (changed[i]) { for (int r = 0; r < 6; ++r) _hist_alongz[r * num_division + num_division - 1] = _hist_alongz[r * num_division + i]; break; }
}

int FingerProperty::getBinAlongZ(float z) const {
  int b = (int)(std::max(z - _min_z, 0.f) / _stride_z);
  return std::max(0, std::min(b, _num_division - 1));
}

// ---- Hand ---------------------------------------------------------------------------------------------------------------------
Hand::Hand(ConfigParser *cfg1, hop_ctx *c) : cfg(cfg1), ctx(c) {}

void Hand::free_scene() {
  hop_cloud_free(ctx, d_scene_hand_region); hop_cloud_free(ctx, d_removed_noise); hop_cloud_free(ctx, d_remove_swivel);
  d_scene_hand_region = d_removed_noise = d_remove_swivel = nullptr;
}

Hand::~Hand() {
  free_scene();
  for (auto &h : d_clouds) hop_cloud_free(ctx, h.second);
}

void Hand::check(int rc, const char *what) const {
  if (rc != HOP_OK) { fprintf(stderr, "%s: %s\n", what, hop_last_error(ctx)); exit(1); }
}

void Hand::addComponent(const std::string &name, const std::string &parent_name, const Cloud &cloud, const Mat4f &tf_in_parent) {
  Cloud ds;
  downsamplePointCloud(cloud, ds, 0.005f);
  _clouds[name] = ds;
  _parent_names[name] = parent_name;
  _tf_in_parent[name] = tf_in_parent;
  _tf_self[name] = Mat4f();
  _component_status[name] = false;
  if (name.find("finger") != std::string::npos) _finger_properties[name] = FingerProperty(ds, 10);   // Hand.cpp:269
  hop_cloud *d = nullptr;
  check(hop_cloud_upload(ctx, ds.xyz.data(), ds.has_normals() ? ds.nrm.data() : nullptr, nullptr, (int)ds.size(), &d), "upload link cloud");
  auto it = d_clouds.find(name);
  if (it != d_clouds.end()) hop_cloud_free(ctx, it->second);
  d_clouds[name] = d;
}

bool Hand::parseURDF(const std::string &urdf_path, std::string *err) {
  std::vector<UrdfLink> links;
  std::string e;
  if (!parseUrdfLinks(urdf_path, links, &e)) { if (err) *err = e; return false; }
  for (const UrdfLink &L : links) {
    const std::string &name = L.name;
    Cloud cloud;
    const std::string cloud_path = cfg->yml["Hand"][name]["cloud"].as<std::string>(std::string());
    if (cloud_path.empty() || !loadPLYFile(cloud_path, cloud, &e) || cloud.size() == 0) { if (err) *err = "Hand." + name + ".cloud: " + (cloud_path.empty() ? "not configured" : e); return false; }
    for (size_t i = 0; i < cloud.size(); ++i) for (int k = 0; k < 3; ++k) cloud.xyz[3 * i + k] *= L.scale[k];
    Cloud moved;
    transformPointCloudWithNormals(cloud, moved, L.tf_init);   // "component init pose must be applied at beginning according to URDF"
    for (const char *kind : {"mesh", "convex_mesh"}) {
      const std::string mp = cfg->yml["Hand"][name][kind].as<std::string>(std::string());
      LinkMesh lm;
      if (mp.empty() || !loadOBJMesh(mp, lm.V, lm.F, nullptr)) continue;   // meshes feed the physics / render stages only
      for (size_t i = 0; i + 2 < lm.V.size(); i += 3) {
        const float x = lm.V[i] * L.scale[0], y = lm.V[i + 1] * L.scale[1], z = lm.V[i + 2] * L.scale[2];
        for (int r = 0; r < 3; ++r) lm.V[i + r] = L.tf_init(r, 0) * x + L.tf_init(r, 1) * y + L.tf_init(r, 2) * z + L.tf_init(r, 3);
      }
      (std::string(kind) == "mesh" ? _meshes : _convex_meshes)[name] = lm;
    }
    std::printf("adding component name:%s\n", name.c_str());
    addComponent(name, L.parent, moved, L.tf_in_parent);
  }
  return !_clouds.empty();
}

void Hand::getTFHandBase(std::string cur_name, Mat4f &tf_in_handbase) const {
  tf_in_handbase = Mat4f();
  while (cur_name != "base_link") {
    auto s = _tf_self.find(cur_name);
    if (s == _tf_self.end()) { std::printf("cur_name does not exist!!!\n"); exit(1); }
    tf_in_handbase = _tf_in_parent.at(cur_name) * s->second * tf_in_handbase;
    cur_name = _parent_names.at(cur_name);
  }
}

static void geodesic_and_pitch(const Mat4f &T, float &rot_diff_deg, float &pitch) {
  // Utils::rotationGeodesicDistance(I, R) (Utils.cpp:29-32) and Eigen's eulerAngles(2,1,0)[1]
  const float tr = T(0, 0) + T(1, 1) + T(2, 2);
  rot_diff_deg = (float)(std::acos(std::max(-1.0, std::min(1.0, ((double)tr - 1.0) / 2.0))) / M_PI * 180.0);
  float a0 = std::atan2(T(1, 0), T(0, 0));
  const float c2 = std::sqrt(T(2, 2) * T(2, 2) + T(2, 1) * T(2, 1));
  pitch = a0 < 0.f ? std::atan2(-T(2, 0), -c2) : std::atan2(-T(2, 0), c2);
}

void Hand::handbaseICP(const Cloud &scene_organized) {
  auto base = d_clouds.find("base_link");
  if (base == d_clouds.end() || scene_organized.size() == 0) return;
  hop_cloud *d_org = nullptr, *ds = nullptr, *hb = nullptr, *px = nullptr, *pz = nullptr, *region = nullptr;
  check(hop_cloud_upload(ctx, scene_organized.xyz.data(), scene_organized.nrm.data(), nullptr, (int)scene_organized.size(), &d_org), "upload scene");
  const Mat4f cam_in_handbase = _handbase_in_cam.inverse();
  check(hop_cloud_voxel_grid(ctx, d_org, 0.005f, &ds), "voxel grid");
  check(hop_cloud_transform(ctx, ds, cam_in_handbase.data(), &hb), "transform");
  check(hop_cloud_pass_through(ctx, hb, 0, -0.07f, 0.03f, &px), "pass x");
  check(hop_cloud_pass_through(ctx, px, 2, -0.18f, 0.01f, &pz), "pass z");
  const Mat4f &f1 = _tf_in_parent.count("finger_1_1") ? _tf_in_parent["finger_1_1"] : Mat4f(), &f2 = _tf_in_parent.count("finger_2_1") ? _tf_in_parent["finger_2_1"] : Mat4f();
  check(hop_cloud_handbase_region(ctx, pz, f1(1, 3), f1(2, 3), f2(1, 3), f2(2, 3), &region), "handbase region");
  Mat4f offset;   // cam2handbase_offset
  if (hop_cloud_size(region) > 0) {
    hop_icp_params p;
    hop_default_icp_params(&p);
    p.max_iter = 50; p.angle_deg = 30.f; p.max_dist = 0.03f;   // Utils::runICP(scene_handbase, handbase, T, 50, 30, 0.03, 1e-4)
    Mat4f pose;    // hop_icp_refine moves the TARGET's pose: pose <- T^-1 * pose; from the identity it returns T^-1
    check(hop_icp_refine(ctx, region, base->second, pose.data(), 1, &p, nullptr, nullptr), "hop_icp_refine (handbase)");
    offset = pose.inverse();
  }
  for (hop_cloud *c : {d_org, ds, hb, px, pz, region}) hop_cloud_free(ctx, c);
  const float translation = std::sqrt(offset(0, 3) * offset(0, 3) + offset(1, 3) * offset(1, 3) + offset(2, 3) * offset(2, 3));
  if (translation >= 0.05) { printf("cam2handbase_offset set to Identity, translation=%f\n", translation); offset = Mat4f(); }
  float rot_diff, pitch;
  geodesic_and_pitch(offset, rot_diff, pitch);
  pitch = std::min(std::abs(pitch), std::abs((float)M_PI - pitch));
  pitch = std::min(std::abs(pitch), std::abs((float)M_PI + pitch));
  if (rot_diff >= 10 || std::abs(pitch) >= 10 / 180.0 * M_PI) { offset = Mat4f(); printf("cam2handbase_offset set to Identity"); }
  bool identity = true;
  { const Mat4f I; for (int k = 0; k < 16; ++k) identity = identity && offset.m[k] == I.m[k]; }
  if (!identity) _component_status["handbase"] = true;
  _handbase_in_cam = _handbase_in_cam * offset.inverse();
}

void Hand::setCurScene(const Cloud &scene_organized, const Cloud &scene_hand_region, const Mat4f &handbase_in_cam) {
  _handbase_in_cam = handbase_in_cam;
  _component_status["handbase"] = false;
  handbaseICP(scene_organized);
  free_scene();
  hop_cloud *region = nullptr, *ds = nullptr, *r1 = nullptr, *r2 = nullptr;
  check(hop_cloud_upload(ctx, scene_hand_region.xyz.data(), scene_hand_region.nrm.data(), nullptr, (int)scene_hand_region.size(), &region), "upload hand region");
  const Mat4f cih = _handbase_in_cam.inverse();
  check(hop_cloud_voxel_grid(ctx, region, 0.003f, &ds), "voxel grid");
  check(hop_cloud_transform(ctx, ds, cih.data(), &d_scene_hand_region), "transform");
  check(hop_cloud_radius_outlier_removal(ctx, d_scene_hand_region, 0.02f, 30, &r1), "radius outlier removal");
  check(hop_cloud_radius_outlier_removal(ctx, r1, 0.04f, 100, &r2), "radius outlier removal");
  check(hop_cloud_statistical_outlier_removal(ctx, r2, 20, 2.0f, &d_removed_noise), "statistical outlier removal");
  check(hop_cloud_pass_through(ctx, d_removed_noise, 0, -0.25f, -0.1f, &d_remove_swivel), "pass through");
  for (hop_cloud *c : {region, ds, r1, r2}) hop_cloud_free(ctx, c);
  if (hop_cloud_size(d_remove_swivel) == 0) { printf("scene_remove_swivel is empty\n"); }   // the reference asserts (Hand.cpp:322)
}

bool Hand::matchOneComponentPSO(std::string model_name, float min_angle, float max_angle, bool /*use_normal: read from hand_match.check_normal like objFuncPSO*/,
                                float dist_thres, float normal_angle_thres, float least_match) {
  static const std::map<std::string, std::string> pair_names = {{"finger_1_1", "finger_2_1"}, {"finger_2_1", "finger_1_1"}, {"finger_1_2", "finger_2_2"}, {"finger_2_2", "finger_1_2"}};
  auto failed = [&]() { printf("%s PSO matching failed\n", model_name.c_str()); _tf_self[model_name] = Mat4f(); _component_status[model_name] = false; return false; };
  if (!d_removed_noise || !pair_names.count(model_name) || !d_clouds.count(model_name)) return failed();
  const std::string pair_name = pair_names.at(model_name);
  const bool palm_side = model_name == "finger_1_1" || model_name == "finger_2_1";
  hop_finger_params p;
  std::memset(&p, 0, sizeof(p));
  auto tip = [&](const FingerProperty &fp, bool max_z, const Mat4f &T, float *out4) {   // T * (min_x, max_y, min_z | max_z, 1)
    const float v[3] = {fp._min_x, fp._max_y, max_z ? fp._max_z : fp._min_z};
    for (int r = 0; r < 3; ++r) out4[r] = T(r, 0) * v[0] + T(r, 1) * v[1] + T(r, 2) * v[2] + T(r, 3);
  };
  Mat4f pair_in_base;
  float t1[3], t2[3];
  const FingerProperty &mine = _finger_properties[model_name];
  Mat4f out2parent;
  if (palm_side) {   // the opposite finger taken as a whole straight finger (Hand.cpp:620-633)
    const std::string pair_out = pair_name == "finger_1_1" ? "finger_1_2" : "finger_2_2";
    getTFHandBase(pair_out, pair_in_base); tip(_finger_properties[pair_out], false, pair_in_base, t1);
    getTFHandBase(pair_name, pair_in_base); tip(_finger_properties[pair_name], false, pair_in_base, t2);
    const std::string out_name = model_name == "finger_1_1" ? "finger_1_2" : "finger_2_2";
    out2parent = _tf_in_parent[out_name];
    const FingerProperty &outp = _finger_properties[out_name];
    p.tip1_local[0] = outp._min_x; p.tip1_local[1] = outp._max_y; p.tip1_local[2] = outp._min_z;
    p.tip2_local[0] = mine._min_x; p.tip2_local[1] = mine._max_y; p.tip2_local[2] = mine._min_z;
  } else {
    getTFHandBase(pair_name, pair_in_base);
    tip(_finger_properties[pair_name], false, pair_in_base, t1);
    tip(_finger_properties[pair_name], true, pair_in_base, t2);
    p.tip1_local[0] = mine._min_x; p.tip1_local[1] = mine._max_y; p.tip1_local[2] = mine._min_z;
    p.tip2_local[0] = mine._min_x; p.tip2_local[1] = mine._max_y; p.tip2_local[2] = mine._max_z;
  }
  p.pair_tip1_y = t1[1]; p.pair_tip2_y = t2[1];
  Mat4f model2handbase;
  getTFHandBase(model_name, model2handbase);
  std::memcpy(p.model2handbase, model2handbase.data(), 64);
  std::memcpy(p.finger_out2parent, out2parent.data(), 64);
  p.palm_side = palm_side; p.right_side = model_name == "finger_2_1" || model_name == "finger_2_2";
  p.gripper_min_dist = (float)cfg->gripper_min_dist;
  p.dist_thres = dist_thres; p.normal_angle_deg = normal_angle_thres;
  p.check_normal = cfg->yml["hand_match"]["check_normal"].as<bool>(true) ? 1 : 0;
  p.num_division = mine._num_division; p.min_z = mine._min_z; p.stride_z = mine._stride_z;
  for (int b = 0; b < mine._num_division && b < HOP_MAX_FINGER_BINS; ++b) p.hist_min_y[b] = mine._hist_alongz[1 * mine._num_division + b];
  p.max_outter_pts = cfg->yml["hand_match"]["max_outter_pts"].as<int>(300);
  p.outter_pt_dist = cfg->yml["hand_match"]["outter_pt_dist"].as<float>(0.002f);
  p.outter_pt_dist_weight = cfg->yml["hand_match"]["outter_pt_dist_weight"].as<float>(1.f);
  // the admissible interval as a dense grid (the swarm's bounds: min_angle*M_PI/180 .. max_angle*M_PI/180, float * double)
  const double lo = min_angle * M_PI / 180, hi = max_angle * M_PI / 180;
  std::vector<double> thetas(n_states), cost(n_states);
  for (int s = 0; s < n_states; ++s) thetas[s] = lo + (hi - lo) * ((double)s / std::max(n_states - 1, 1));
  int32_t best = -1;
  check(hop_hand_overlap(ctx, d_clouds[model_name], d_removed_noise, d_scene_hand_region, d_remove_swivel, &p, thetas.data(), n_states, cost.data(), &best),
        "hop_hand_overlap");
  objval = best >= 0 ? cost[best] : 0;
  if (best < 0 || -objval <= least_match) return failed();
  const float angle = (float)thetas[best];
  Mat4f R;
  R(1, 1) = std::cos(angle); R(1, 2) = -std::sin(angle); R(2, 1) = std::sin(angle); R(2, 2) = std::cos(angle);
  _tf_self[model_name] = R;
  _component_status[model_name] = true;
  printf("%s PSO final angle=%f, match_score=%f\n", model_name.c_str(), angle, -objval);
  return true;
}

void Hand::makeHandCloud() {
  _hand_cloud.clear();
  _link_clouds_in_handbase.clear();
  for (auto &h : _clouds) {
    Mat4f model2handbase;
    getTFHandBase(h.first, model2handbase);
    Cloud tmp;
    transformPointCloudWithNormals(h.second, tmp, model2handbase);
    for (size_t i = 0; i < tmp.size(); ++i) _hand_cloud.push(&tmp.xyz[3 * i], tmp.has_normals() ? &tmp.nrm[3 * i] : nullptr, 1.f);
    _link_clouds_in_handbase[h.first] = tmp;
  }
}

void Hand::adjustHandHeight() {
  makeHandCloud();
  if (_component_status["handbase"]) return;
  if (!d_scene_hand_region || _hand_cloud.size() == 0) return;
  static const float heights[13] = {-0.03f, -0.025f, -0.02f, -0.015f, -0.01f, -0.005f, 0, 0.005f, 0.01f, 0.015f, 0.02f, 0.025f, 0.03f};
  hop_cloud *d_hand = nullptr;
  check(hop_cloud_upload(ctx, _hand_cloud.xyz.data(), _hand_cloud.nrm.data(), nullptr, (int)_hand_cloud.size(), &d_hand), "upload hand cloud");
  int32_t best = -1;
  check(hop_adjust_hand_height(ctx, d_hand, d_scene_hand_region, heights, 13, nullptr, &best), "hop_adjust_hand_height");
  hop_cloud_free(ctx, d_hand);
  if (best >= 0) { Mat4f off; off(2, 3) = heights[best]; _handbase_in_cam = _handbase_in_cam * off; }
}

void Hand::removeSurroundingPointsAndAssignProbability(const Cloud &scene, Cloud &scene_out, float dist_thres) {
  scene_out.clear();
  if (scene.size() == 0) return;
  std::vector<hop_cloud *> links;
  std::vector<int32_t> kinds;
  for (auto &h : _link_clouds_in_handbase) {   // std::map order, like _kdtrees
    hop_cloud *d = nullptr;
    check(hop_cloud_upload(ctx, h.second.xyz.data(), nullptr, nullptr, (int)h.second.size(), &d), "upload link");
    links.push_back(d);
    kinds.push_back(h.first == "finger_2_1" || h.first == "finger_1_1" ? 1 : (h.first == "base" || h.first == "swivel_1" || h.first == "swivel_2" ? 2 : 0));
  }
  hop_hand_removal_params p;
  const Mat4f cih = _handbase_in_cam.inverse();
  Mat4f f12, f22;
  if (_tf_self.count("finger_1_2")) getTFHandBase("finger_1_2", f12);
  if (_tf_self.count("finger_2_2")) getTFHandBase("finger_2_2", f22);
  const Mat4f f12i = f12.inverse(), f22i = f22.inverse();
  std::memcpy(p.cam_in_handbase, cih.data(), 64); std::memcpy(p.handbase_in_cam, _handbase_in_cam.data(), 64);
  std::memcpy(p.handbase_in_finger_1_2, f12i.data(), 64); std::memcpy(p.handbase_in_finger_2_2, f22i.data(), 64);
  p.min_z = _finger_properties.count("finger_1_2") ? _finger_properties["finger_1_2"]._min_z : FLT_MAX;
  p.dist_thres_sq = dist_thres;
  hop_cloud *d_scene = nullptr, *d_out = nullptr;
  check(hop_cloud_upload(ctx, scene.xyz.data(), scene.nrm.data(), nullptr, (int)scene.size(), &d_scene), "upload scene");
  check(hop_remove_hand_points(ctx, d_scene, links.data(), kinds.data(), (int)links.size(), &p, &d_out), "hop_remove_hand_points");
  const int n = hop_cloud_size(d_out);
  scene_out.xyz.resize(3 * (size_t)n); scene_out.nrm.resize(3 * (size_t)n); scene_out.conf.resize(n);
  if (n > 0) check(hop_cloud_download(ctx, d_out, scene_out.xyz.data(), scene_out.nrm.data(), scene_out.conf.data()), "download");
  for (hop_cloud *c : links) hop_cloud_free(ctx, c);
  hop_cloud_free(ctx, d_scene); hop_cloud_free(ctx, d_out);
}

HandState Hand::state() const {
  HandState s;
  s._component_status = _component_status;
  s._hand_cloud = _hand_cloud;
  s._handbase_in_cam = _handbase_in_cam;
  for (auto &h : _link_clouds_in_handbase)
    if (h.first.find("finger") != std::string::npos) s.finger_clouds[h.first] = h.second;
  return s;
}
