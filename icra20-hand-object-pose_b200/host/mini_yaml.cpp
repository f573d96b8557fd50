#include "mini_yaml.h"

#include <fstream>

namespace miniyaml {

namespace {
const Node kUndefined;

std::string trim(const std::string &s) {
  size_t b = s.find_first_not_of(" \t\r\n"), e = s.find_last_not_of(" \t\r\n");
  return b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
}

// strips a trailing comment (a '#' at line start or preceded by whitespace, outside quotes)
std::string strip_comment(const std::string &line) {
  bool sq = false, dq = false;
  for (size_t i = 0; i < line.size(); ++i) {
    const char c = line[i];
    if (c == '\'' && !dq) sq = !sq;
    else if (c == '"' && !sq) dq = !dq;
    else if (c == '#' && !sq && !dq && (i == 0 || line[i - 1] == ' ' || line[i - 1] == '\t')) return line.substr(0, i);
  }
  return line;
}

std::string unquote(const std::string &s) {
  if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) return s.substr(1, s.size() - 2);
  return s;
}

Node parse_flow_seq(const std::string &text) {  // "[a, b, [c, d]]" without the outer brackets handled by the caller
  Node n;
  n.kind = Node::Sequence;
  std::string cur;
  int depth = 0;
  auto flush = [&]() {
    const std::string t = trim(cur);
    cur.clear();
    if (t.empty()) return;
    Node e;
    if (t.front() == '[' && t.back() == ']') e = parse_flow_seq(t.substr(1, t.size() - 2));
    else { e.kind = Node::Scalar; e.scalar = unquote(t); }
    n.seq.push_back(e);
  };
  for (char c : text) {
    if (c == '[') ++depth;
    if (c == ']') --depth;
    if (c == ',' && depth == 0) flush(); else cur.push_back(c);
  }
  flush();
  return n;
}

struct Line { int indent; std::string text; };

int bracket_balance(const std::string &s) {
  int b = 0;
  for (char c : s) { if (c == '[') ++b; if (c == ']') --b; }
  return b;
}

Node parse_block(const std::vector<Line> &lines, size_t &i, int indent) {
  Node node;
  node.kind = Node::Map;
  while (i < lines.size()) {
    const Line &ln = lines[i];
    if (ln.indent < indent) break;
    if (ln.indent > indent) throw std::runtime_error("miniyaml: unexpected indentation at '" + ln.text + "'");
    const size_t colon = ln.text.find(':');
    if (colon == std::string::npos) throw std::runtime_error("miniyaml: expected 'key: value' at '" + ln.text + "'");
    const std::string key = unquote(trim(ln.text.substr(0, colon)));
    std::string rest = trim(ln.text.substr(colon + 1));
    ++i;
    Node value;
    if (rest.empty()) {
      if (i < lines.size() && lines[i].indent > indent) value = parse_block(lines, i, lines[i].indent);
    } else if (rest.front() == '[') {
      int bal = bracket_balance(rest);
      while (bal > 0 && i < lines.size()) { rest += " " + lines[i].text; bal = bracket_balance(rest); ++i; }  // multi-line flow sequence
      rest = trim(rest);
      if (rest.back() != ']') throw std::runtime_error("miniyaml: unterminated sequence for key '" + key + "'");
      value = parse_flow_seq(rest.substr(1, rest.size() - 2));
    } else {
      value.kind = Node::Scalar;
      value.scalar = unquote(rest);
    }
    node.map.emplace_back(key, value);
  }
  return node;
}
}  // namespace

bool Node::has(const std::string &key) const {
  for (const auto &kv : map) if (kv.first == key) return true;
  return false;
}
const Node &Node::operator[](const std::string &key) const {
  for (const auto &kv : map) if (kv.first == key) return kv.second;
  return kUndefined;
}
std::vector<std::string> Node::keys() const {
  std::vector<std::string> k;
  for (const auto &kv : map) k.push_back(kv.first);
  return k;
}

Node Load(const std::string &text) {
  std::vector<Line> lines;
  std::istringstream is(text);
  std::string raw;
  while (std::getline(is, raw)) {
    std::string s = strip_comment(raw);
    if (trim(s).empty()) continue;
    size_t ind = 0;
    while (ind < s.size() && s[ind] == ' ') ++ind;
    lines.push_back(Line{(int)ind, trim(s)});
  }
  size_t i = 0;
  if (lines.empty()) return Node();
  // continuation lines of a multi-line flow sequence may sit at any indentation: they are consumed by their key's parser
  return parse_block(lines, i, lines[0].indent);
}

Node LoadFile(const std::string &path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("miniyaml: cannot open " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  return Load(ss.str());
}

}  // namespace miniyaml
