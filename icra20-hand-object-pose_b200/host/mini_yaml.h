// mini_yaml.h -- the subset of YAML that config_autodataset.yaml uses, behind a yaml-cpp-like interface
// (cfg.yml["hand_match"]["pso"]["n_pop"].as<int>()): block mappings nested by indentation, plain / quoted scalars,
// flow sequences "[a, b, c]" that may continue over several lines, and '#' comments.  yaml-cpp itself is not installed.
#pragma once
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace miniyaml {

class Node {
 public:
  enum Kind { Null, Scalar, Sequence, Map };
  Kind kind = Null;
  std::string scalar;
  std::vector<Node> seq;
  std::vector<std::pair<std::string, Node>> map;  // file order kept

  bool IsDefined() const { return kind != Null; }
  bool IsMap() const { return kind == Map; }
  bool IsSequence() const { return kind == Sequence; }
  size_t size() const { return kind == Sequence ? seq.size() : (kind == Map ? map.size() : 0); }
  bool has(const std::string &key) const;
  const Node &operator[](const std::string &key) const;  // undefined node when missing (as<T>() on it throws)
  const Node &operator[](size_t i) const { return seq.at(i); }
  template <class T> T as() const;
  template <class T> T as(const T &fallback) const { return kind == Scalar ? as<T>() : fallback; }
  std::vector<std::string> keys() const;
};

Node LoadFile(const std::string &path);   // throws std::runtime_error
Node Load(const std::string &text);

template <> inline std::string Node::as<std::string>() const {
  if (kind != Scalar) throw std::runtime_error("miniyaml: not a scalar");
  return scalar;
}
template <> inline bool Node::as<bool>() const {
  const std::string s = as<std::string>();
  if (s == "true" || s == "True" || s == "TRUE" || s == "yes" || s == "on") return true;
  if (s == "false" || s == "False" || s == "FALSE" || s == "no" || s == "off") return false;
  throw std::runtime_error("miniyaml: bad bool '" + s + "'");
}
template <class T> inline T Node::as() const {
  const std::string s = as<std::string>();
  std::istringstream is(s);
  T v;
  is >> v;
  if (is.fail()) throw std::runtime_error("miniyaml: cannot convert '" + s + "'");
  return v;
}
template <> inline std::vector<float> Node::as<std::vector<float>>() const {
  if (kind != Sequence) throw std::runtime_error("miniyaml: not a sequence");
  std::vector<float> out;
  for (const Node &n : seq) out.push_back(n.as<float>());
  return out;
}

}  // namespace miniyaml
