// urdf.h -- what Hand::parseURDF (src/perception/src/Hand.cpp:375-502) reads from the hand's URDF: per <link> (rails skipped) the
// visual origin (the link's initial pose), the visual mesh scale, and from the <joint> whose <child> it is the parent link and the
// pose in the parent.  Rotations are Rz(yaw) * Ry(pitch) * Rx(roll) of the rpy attribute, like the reference's AngleAxisf product.
#pragma once
#include <string>
#include <vector>

#include "mat.h"

struct UrdfLink {
  std::string name, parent;
  Mat4f tf_init, tf_in_parent;
  float scale[3] = {1.f, 1.f, 1.f};
};

bool parseUrdfLinks(const std::string &path, std::vector<UrdfLink> &links, std::string *err = nullptr);
