// tictoc.h — wall-clock stopwatch with the interface of the reference's TicToc (include/tictoc.h:16-65):
// tic() restarts, toc() returns elapsed milliseconds.  (Device-side stage times come from CUDA events:
// scvod_kernel_timing / scvod_kernel_timing_report in include/scvod.h.)
#pragma once
#include <chrono>

class TicToc {
 public:
  TicToc() { tic(); }
  explicit TicToc(bool display) : display_(display) { tic(); }
  void tic() { t0_ = std::chrono::steady_clock::now(); }
  double toc() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0_).count(); }

 private:
  std::chrono::steady_clock::time_point t0_;
  bool display_ = false;
};
