// Flat "key: value" settings reader with the slice of yaml-cpp's API the reference uses on gpuhc_settings.yaml
// (YAML::LoadFile, node["key"].as<T>(), operator<<; reference GPU_HC_Solver.cpp:46-66, cmd/magmaHC-main.cpp:239-251).
// yaml-cpp is not available in this image; the settings file is a flat map, so nothing else is needed.
#ifndef HCB200_HOST_YAML_LITE_HPP
#define HCB200_HOST_YAML_LITE_HPP
#include <fstream>
#include <map>
#include <memory>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace YAML {

class Node {
public:
  Node() : entries_(std::make_shared<Map>()), order_(std::make_shared<std::vector<std::string>>()) {}

  Node operator[](const std::string& key) const {
    auto it = entries_->find(key);
    if (it == entries_->end()) throw std::runtime_error("settings key '" + key + "' is missing");
    Node leaf;
    leaf.scalar_ = it->second;
    leaf.is_scalar_ = true;
    return leaf;
  }
  bool has(const std::string& key) const { return entries_->count(key) != 0; }
  void set(const std::string& key, const std::string& value) {
    if (!entries_->count(key)) order_->push_back(key);
    (*entries_)[key] = value;
  }
  template <typename T> T as() const {
    std::istringstream in(scalar_);
    T out{};
    in >> out;
    if (in.fail()) throw std::runtime_error("cannot convert settings value '" + scalar_ + "'");
    return out;
  }
  template <typename T> T as_or(const std::string& key, T fallback) const { return has(key) ? (*this)[key].template as<T>() : fallback; }

  friend std::ostream& operator<<(std::ostream& os, const Node& n) {
    for (const auto& k : *n.order_) os << k << ": " << n.entries_->at(k) << "\n";
    return os;
  }

private:
  using Map = std::map<std::string, std::string>;
  std::shared_ptr<Map> entries_;
  std::shared_ptr<std::vector<std::string>> order_;
  std::string scalar_;
  bool is_scalar_ = false;
};

template <> inline std::string Node::as<std::string>() const { return scalar_; }
template <> inline bool Node::as<bool>() const {
  return scalar_ == "true" || scalar_ == "True" || scalar_ == "TRUE" || scalar_ == "1" || scalar_ == "yes";
}

inline std::string strip(const std::string& s) {
  const char* ws = " \t\r\n";
  const size_t a = s.find_first_not_of(ws);
  if (a == std::string::npos) return std::string();
  return s.substr(a, s.find_last_not_of(ws) - a + 1);
}

inline Node LoadFile(const std::string& path) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("cannot open settings file " + path);
  Node root;
  std::string line;
  while (std::getline(in, line)) {
    const size_t hash = line.find('#');
    if (hash != std::string::npos) line.erase(hash);
    if (line.empty() || line[0] == '%') continue;
    const size_t colon = line.find(':');
    if (colon == std::string::npos) continue;
    const std::string key = strip(line.substr(0, colon));
    if (!key.empty()) root.set(key, strip(line.substr(colon + 1)));
  }
  return root;
}

}  // namespace YAML
#endif
