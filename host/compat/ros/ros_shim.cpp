// ros_shim.cpp — parameter server + init of the roscpp stand-in (see ros.h).
#include <cstring>
#include <fstream>
#include <iostream>

#include "ros.h"

namespace ros {
namespace param_server {

std::map<std::string, std::string>& table() {
  static std::map<std::string, std::string> t;
  return t;
}

void set(const std::string& key, const std::string& value) { table()[key] = value; }

static std::string strip_comment(const std::string& line) {
  bool in_quote = false;
  for (size_t i = 0; i < line.size(); ++i) {
    if (line[i] == '"') in_quote = !in_quote;
    if (line[i] == '#' && !in_quote) return line.substr(0, i);
  }
  return line;
}
static std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
static std::string unquote(const std::string& s) {
  if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) return s.substr(1, s.size() - 2);
  return s;
}

// The subset of YAML the reference's config files use: top-level "section:" lines, indented
// "key: value" lines, '#' comments, quoted strings and flow sequences that may span lines.
bool load_yaml(const std::string& path) {
  std::ifstream in(path);
  if (!in) return false;
  std::string line, section;
  while (std::getline(in, line)) {
    std::string raw = strip_comment(line);
    if (trim(raw).empty()) continue;
    const bool indented = raw[0] == ' ' || raw[0] == '\t';
    std::string t = trim(raw);
    size_t colon = t.find(':');
    if (colon == std::string::npos) continue;
    std::string key = trim(t.substr(0, colon)), value = trim(t.substr(colon + 1));
    if (!indented && value.empty()) {
      section = key;
      continue;
    }
    if (!value.empty() && value[0] == '[')
      while (value.find(']') == std::string::npos && std::getline(in, line)) value += " " + trim(strip_comment(line));
    set((indented && !section.empty() ? section + "/" : std::string()) + key, unquote(value));
  }
  return true;
}

bool load_launch(const std::string& path) {
  std::ifstream in(path);
  if (!in) return false;
  std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  size_t p = text.find("<rosparam");
  if (p == std::string::npos) return false;
  size_t f = text.find("file=\"", p);
  if (f == std::string::npos) return false;
  f += 6;
  size_t e = text.find('"', f);
  if (e == std::string::npos) return false;
  return load_yaml(text.substr(f, e - f));
}

}  // namespace param_server

void init(int& argc, char** argv, const std::string& name) {
  (void)name;
  bool loaded = false;
  for (int i = 1; i < argc && !loaded; ++i) {
    std::string a = argv[i];
    if (a.rfind("_params:=", 0) == 0)
      loaded = param_server::load_yaml(a.substr(9));
    else if (a == "--params" && i + 1 < argc)
      loaded = param_server::load_yaml(argv[i + 1]);
    else if (a == "--launch" && i + 1 < argc)
      loaded = param_server::load_launch(argv[i + 1]);
  }
  if (!loaded) {
    const char* env = std::getenv("UFO_PARAMS");
    if (env && *env) loaded = param_server::load_yaml(env);
  }
  if (!loaded) std::fprintf(stderr, "[ WARN] no parameter file given (_params:=file.yaml, --launch file.launch or UFO_PARAMS): using Utility defaults\n");
  // `key:=value` overrides, as rosrun would remap private parameters
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    size_t p = a.find(":=");
    if (p != std::string::npos && a.rfind("_params:=", 0) != 0 && a[0] == '_') param_server::set(a.substr(1, p - 1), a.substr(p + 2));
  }
}

namespace console {
static levels::Level g_level = levels::Info;
bool set_logger_level(const std::string&, levels::Level level) {
  g_level = level;
  return true;
}
levels::Level current_level() { return g_level; }
}  // namespace console

}  // namespace ros
