// scvod_timing.cu — optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline block).
#include "scvod_kernel_common.cuh"

namespace scvod {

// ------------------------------------------------------------------------------------------------
// optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline block)
// ------------------------------------------------------------------------------------------------
namespace {
struct TimedLaunch {
  int name_id;
  cudaEvent_t e0, e1;
};
bool g_timing = false;
std::vector<std::string> g_names;
std::vector<double> g_ms;
std::vector<long long> g_cnt;
std::vector<TimedLaunch> g_pending;
std::vector<cudaEvent_t> g_event_pool;
std::mutex g_timing_mu;  // contexts on different host threads share the table

cudaEvent_t get_event() {
  if (!g_event_pool.empty()) {
    cudaEvent_t e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
int name_id(const char* name) {
  for (size_t i = 0; i < g_names.size(); ++i)
    if (g_names[i] == name) return (int)i;
  g_names.push_back(name);
  g_ms.push_back(0.0);
  g_cnt.push_back(0);
  return (int)g_names.size() - 1;
}
}  // namespace

LaunchTimer::LaunchTimer(const char* name, void* stream) : st(stream), id(0), e0(nullptr), e1(nullptr), on(g_timing) {
  if (on) {
    {
      std::lock_guard<std::mutex> lk(g_timing_mu);
      id = name_id(name);
      e0 = (void*)get_event();
      e1 = (void*)get_event();
    }
    cudaEventRecord((cudaEvent_t)e0, (cudaStream_t)st);
  }
}
LaunchTimer::~LaunchTimer() {
  if (on) {
    cudaEventRecord((cudaEvent_t)e1, (cudaStream_t)st);
    TimedLaunch t;
    t.name_id = id;
    t.e0 = (cudaEvent_t)e0;
    t.e1 = (cudaEvent_t)e1;
    std::lock_guard<std::mutex> lk(g_timing_mu);
    g_pending.push_back(t);
  }
}

void timing_enable(bool on) { g_timing = on; }
void timing_reset() {
  timing_collect();
  for (auto& v : g_ms) v = 0;
  for (auto& v : g_cnt) v = 0;
}
void timing_collect() {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  for (auto& t : g_pending) {
    cudaEventSynchronize(t.e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, t.e0, t.e1);
    g_ms[t.name_id] += ms;
    g_cnt[t.name_id] += 1;
    g_event_pool.push_back(t.e0);
    g_event_pool.push_back(t.e1);
  }
  g_pending.clear();
}
std::string timing_report() {
  timing_collect();
  std::string out;
  char line[256];
  for (size_t i = 0; i < g_names.size(); ++i) {
    snprintf(line, sizeof(line), "%s %.6f %lld\n", g_names[i].c_str(), g_ms[i], g_cnt[i]);
    out += line;
  }
  return out;
}

}  // namespace scvod
