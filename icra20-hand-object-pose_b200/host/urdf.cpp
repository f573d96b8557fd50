#include "urdf.h"

#include <cmath>
#include <cstdlib>

#include "mini_xml.h"

namespace {
std::vector<float> floats_of(const std::string &s, size_t n, float fill) {   // Utils::delimitString into n floats
  std::vector<float> v;
  const char *p = s.c_str();
  char *e = nullptr;
  for (;;) { const float x = std::strtof(p, &e); if (e == p) break; v.push_back(x); p = e; }
  v.resize(n, fill);
  return v;
}
Mat4f pose_of(const std::vector<float> &rpy, const std::vector<float> &xyz) {
  const float cr = std::cos(rpy[0]), sr = std::sin(rpy[0]), cp = std::cos(rpy[1]), sp = std::sin(rpy[1]), cy = std::cos(rpy[2]), sy = std::sin(rpy[2]);
  Mat4f T;
  T(0, 0) = cy * cp; T(0, 1) = cy * sp * sr - sy * cr; T(0, 2) = cy * sp * cr + sy * sr;
  T(1, 0) = sy * cp; T(1, 1) = sy * sp * sr + cy * cr; T(1, 2) = sy * sp * cr - cy * sr;
  T(2, 0) = -sp;     T(2, 1) = cp * sr;                T(2, 2) = cp * cr;
  T(0, 3) = xyz[0]; T(1, 3) = xyz[1]; T(2, 3) = xyz[2];
  return T;
}
}  // namespace

bool parseUrdfLinks(const std::string &path, std::vector<UrdfLink> &links, std::string *err) {
  XmlNode doc;
  links.clear();
  if (!loadXmlFile(path, doc, err)) return false;
  const XmlNode &robot = doc.child("robot");
  if (robot.empty()) { if (err) *err = "no <robot> element in " + path; return false; }
  for (const XmlNode *link : robot.all("link")) {
    UrdfLink L;
    L.name = link->get("name");
    if (L.name.find("rail") != std::string::npos) continue;
    const XmlNode &visual = link->child("visual");
    L.tf_init = pose_of(floats_of(visual.child("origin").get("rpy"), 3, 0.f), floats_of(visual.child("origin").get("xyz"), 3, 0.f));
    for (const XmlNode *joint : robot.all("joint")) {
      if (joint->child("child").get("link") != L.name) continue;
      L.parent = joint->child("parent").get("link");
      L.tf_in_parent = pose_of(floats_of(joint->child("origin").get("rpy"), 3, 0.f), floats_of(joint->child("origin").get("xyz"), 3, 0.f));
      break;
    }
    const std::vector<float> sc = floats_of(visual.child("geometry").child("mesh").get("scale"), 3, 1.f);
    for (int k = 0; k < 3; ++k) L.scale[k] = sc[k];
    links.push_back(L);
  }
  return true;
}
