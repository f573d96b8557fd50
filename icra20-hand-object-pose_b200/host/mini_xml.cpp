#include "mini_xml.h"

#include <cctype>
#include <fstream>
#include <sstream>

const XmlNode &XmlNode::child(const std::string &tag) const {
  static const XmlNode none;
  for (const XmlNode &c : children) if (c.name == tag) return c;
  return none;
}
std::vector<const XmlNode *> XmlNode::all(const std::string &tag) const {
  std::vector<const XmlNode *> out;
  for (const XmlNode &c : children) if (c.name == tag) out.push_back(&c);
  return out;
}

namespace {
struct Parser {
  const std::string &s;
  size_t i = 0;
  std::string err;
  explicit Parser(const std::string &t) : s(t) {}
  void skip_ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; }
  bool starts(const char *p) const { return s.compare(i, std::char_traits<char>::length(p), p) == 0; }
  bool skip_misc() {   // whitespace, text, comments, <?...?>, <!DOCTYPE ...>
    for (;;) {
      while (i < s.size() && s[i] != '<') ++i;
      if (i >= s.size()) return true;
      if (starts("<!--")) { const size_t e = s.find("-->", i + 4); if (e == std::string::npos) { err = "unterminated comment"; return false; } i = e + 3; continue; }
      if (starts("<?")) { const size_t e = s.find("?>", i + 2); if (e == std::string::npos) { err = "unterminated declaration"; return false; } i = e + 2; continue; }
      if (starts("<!")) { const size_t e = s.find('>', i + 2); if (e == std::string::npos) { err = "unterminated <!"; return false; } i = e + 1; continue; }
      return true;
    }
  }
  std::string ident() { const size_t b = i; while (i < s.size() && (std::isalnum((unsigned char)s[i]) || s[i] == '_' || s[i] == ':' || s[i] == '-' || s[i] == '.')) ++i; return s.substr(b, i - b); }
  bool element(XmlNode &out) {   // at '<' of an opening tag
    ++i;
    out.name = ident();
    if (out.name.empty()) { err = "tag name expected"; return false; }
    for (;;) {
      skip_ws();
      if (i >= s.size()) { err = "unterminated tag <" + out.name; return false; }
      if (s[i] == '/') { if (i + 1 < s.size() && s[i + 1] == '>') { i += 2; return true; } err = "stray '/' in <" + out.name; return false; }
      if (s[i] == '>') { ++i; break; }
      const std::string key = ident();
      skip_ws();
      if (key.empty() || i >= s.size() || s[i] != '=') { err = "attribute expected in <" + out.name; return false; }
      ++i; skip_ws();
      if (i >= s.size() || (s[i] != '"' && s[i] != '\'')) { err = "quoted value expected in <" + out.name; return false; }
      const char q = s[i++];
      const size_t e = s.find(q, i);
      if (e == std::string::npos) { err = "unterminated attribute value in <" + out.name; return false; }
      out.attr[key] = s.substr(i, e - i);
      i = e + 1;
    }
    for (;;) {   // children until the closing tag
      if (!skip_misc()) return false;
      if (i >= s.size()) { err = "missing </" + out.name + ">"; return false; }
      if (starts("</")) {
        i += 2;
        const std::string close = ident();
        skip_ws();
        if (close != out.name || i >= s.size() || s[i] != '>') { err = "mismatched </" + close + "> for <" + out.name + ">"; return false; }
        ++i;
        return true;
      }
      XmlNode c;
      if (!element(c)) return false;
      out.children.push_back(std::move(c));
    }
  }
};
}  // namespace

bool parseXml(const std::string &text, XmlNode &root, std::string *err) {
  Parser p(text);
  root = XmlNode();
  for (;;) {
    if (!p.skip_misc()) { if (err) *err = p.err; return false; }
    if (p.i >= text.size()) return true;
    if (p.starts("</")) { if (err) *err = "unexpected closing tag"; return false; }
    XmlNode c;
    if (!p.element(c)) { if (err) *err = p.err; return false; }
    root.children.push_back(std::move(c));
  }
}

bool loadXmlFile(const std::string &path, XmlNode &root, std::string *err) {
  std::ifstream f(path);
  if (!f) { if (err) *err = "cannot open " + path; return false; }
  std::stringstream ss;
  ss << f.rdbuf();
  return parseXml(ss.str(), root, err);
}
