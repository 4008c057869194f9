/* TEST INFRASTRUCTURE ONLY (oracle/_ref/libfilterref.so): stands in for the reference's src/common/progress.hpp, whose
 * ProgressMeter starts a reporter thread that prints to std::cerr. The mapping filters only call increment(). */
#pragma once
#include <atomic>
#include <cstdint>
#include <memory>
#include <string>
namespace progress_meter {
class ProgressMeter {
 public:
  std::atomic<bool> is_finished{false};
  ProgressMeter() {}
  ProgressMeter(uint64_t, const std::string&, bool = true) {}
  void increment(const uint64_t&) {}
  void finish() {}
  void reset_timer() {}
};
}
