// Flat "key: value" reader standing in for yaml-cpp (absent from this image); enough for the reference's
// gpuhc_settings.yaml, which is a flat map (problems/trifocal_2op1p_30x30/gpuhc_settings.yaml:5-34).
#ifndef HCB200_REFSHIM_YAML_H
#define HCB200_REFSHIM_YAML_H
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <iostream>
namespace YAML {
class Node {
  std::shared_ptr<std::map<std::string, std::string>> m_;
  std::string v_;
public:
  Node() : m_(std::make_shared<std::map<std::string, std::string>>()) {}
  void set(const std::string& k, const std::string& v) { (*m_)[k] = v; }
  Node operator[](const std::string& k) const {
    Node n; auto it = m_->find(k);
    if (it == m_->end()) throw std::runtime_error("yaml key not found: " + k);
    n.v_ = it->second; return n;
  }
  template <typename T> T as() const { std::istringstream s(v_); T t; s >> t; return t; }
  friend std::ostream& operator<<(std::ostream& o, const Node& n) { for (auto& kv : *n.m_) o << kv.first << ": " << kv.second << "\n"; return o; }
};
template <> inline std::string Node::as<std::string>() const { return v_; }
template <> inline bool Node::as<bool>() const { return v_ == "true" || v_ == "True" || v_ == "1"; }
inline Node LoadFile(const std::string& path) {
  std::ifstream f(path); if (!f) throw std::runtime_error("cannot open " + path);
  Node n; std::string ln;
  while (std::getline(f, ln)) {
    auto h = ln.find('#'); if (h != std::string::npos) ln = ln.substr(0, h);
    auto c = ln.find(':'); if (c == std::string::npos || ln[0] == '%') continue;
    auto trim = [](std::string s) { size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r"); return a == std::string::npos ? std::string() : s.substr(a, b - a + 1); };
    n.set(trim(ln.substr(0, c)), trim(ln.substr(c + 1)));
  }
  return n;
}
}
#endif
