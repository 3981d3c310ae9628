// ros/ros.h — minimal stand-in for roscpp so that the SSC / Utility class surface and the reference's
// unchanged src/main.cpp (main.cpp:3-15) build in an image without ROS.  With a real ROS install, drop
// host/compat from the include path: nothing in host/include or host/src depends on this shim's internals.
//
// The parameter server is backed by the YAML file the launch file would have loaded
// (launch/run_semantickitti.launch:6: <rosparam file=".../config/semantickitti.yaml" command="load"/>):
//   ros::init looks for  `_params:=<file.yaml>` / `--params <file.yaml>` / `--launch <file.launch>` in argv,
//   then the environment variable UFO_PARAMS.  Missing keys fall back to the defaults given to param().
#pragma once
#include <cstdio>
#include <cstdlib>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace ros {

namespace param_server {
std::map<std::string, std::string>& table();           // "section/key" -> raw scalar or "[a, b, c]"
bool load_yaml(const std::string& path);               // two-level YAML subset used by config/*.yaml
bool load_launch(const std::string& path);             // picks the <rosparam file="..."> attribute
void set(const std::string& key, const std::string& value);
}  // namespace param_server

void init(int& argc, char** argv, const std::string& name);
inline bool ok() { return true; }
inline void spin() {}      // nothing is subscribed in the reference either (main.cpp:12 idles)
inline void shutdown() {}

namespace console {
namespace levels {
enum Level { Debug, Info, Warn, Error, Fatal };
}
bool set_logger_level(const std::string& name, levels::Level level);
levels::Level current_level();
}  // namespace console

class NodeHandle {
 public:
  template <typename T>
  bool param(const std::string& key, T& value, const T& fallback) const {
    auto& t = param_server::table();
    auto it = t.find(key);
    if (it == t.end()) {
      value = fallback;
      return false;
    }
    return parse(it->second, value) || ((value = fallback), false);
  }

 private:
  static bool parse(const std::string& s, std::string& v) {
    v = s;
    return true;
  }
  static bool parse(const std::string& s, bool& v) {
    v = (s == "true" || s == "True" || s == "1");
    return true;
  }
  template <typename T>
  static bool parse(const std::string& s, T& v) {
    std::istringstream is(s);
    return (bool)(is >> v);
  }
  template <typename T>
  static bool parse(const std::string& s, std::vector<T>& v) {
    v.clear();
    std::string body = s;
    for (char& c : body)
      if (c == '[' || c == ']' || c == ',') c = ' ';
    std::istringstream is(body);
    T x;
    while (is >> x) v.push_back(x);
    return true;
  }
};

}  // namespace ros

#define ROSCONSOLE_DEFAULT_NAME "ros.ufo"
#define ROS_LOG_AT__(lvl, tag, ...)                          \
  do {                                                       \
    if ((int)ros::console::current_level() <= (int)(lvl)) { \
      std::fprintf(stderr, "[%s] ", tag);                    \
      std::fprintf(stderr, __VA_ARGS__);                     \
      std::fprintf(stderr, "\n");                            \
    }                                                        \
  } while (0)
#define ROS_DEBUG(...) ROS_LOG_AT__(ros::console::levels::Debug, "DEBUG", __VA_ARGS__)
#define ROS_INFO(...) ROS_LOG_AT__(ros::console::levels::Info, " INFO", __VA_ARGS__)
#define ROS_WARN(...) ROS_LOG_AT__(ros::console::levels::Warn, " WARN", __VA_ARGS__)
#define ROS_ERROR(...) ROS_LOG_AT__(ros::console::levels::Error, "ERROR", __VA_ARGS__)
#define ROS_BREAK() std::abort()
