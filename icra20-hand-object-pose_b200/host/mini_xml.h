// mini_xml.h -- the XML subset a URDF needs (elements, attributes, nesting, comments, declarations, self-closing tags), in place of
// the pugixml copy the reference vendors (src/perception/include/pugixml.hpp).  Text content and entities are ignored: parseURDF
// (Hand.cpp:375-502) only reads attributes.
#pragma once
#include <map>
#include <string>
#include <vector>

struct XmlNode {
  std::string name;
  std::map<std::string, std::string> attr;
  std::vector<XmlNode> children;
  // first child with that tag, or an empty node (like pugi::xml_node's null handle: lookups on it yield nothing)
  const XmlNode &child(const std::string &tag) const;
  std::vector<const XmlNode *> all(const std::string &tag) const;
  bool has(const std::string &a) const { return attr.count(a) != 0; }
  std::string get(const std::string &a) const { auto it = attr.find(a); return it == attr.end() ? std::string() : it->second; }
  bool empty() const { return name.empty(); }
};

// parses `text`; returns false (with a message) on malformed input.  root.children holds the top-level elements.
bool parseXml(const std::string &text, XmlNode &root, std::string *err = nullptr);
bool loadXmlFile(const std::string &path, XmlNode &root, std::string *err = nullptr);
