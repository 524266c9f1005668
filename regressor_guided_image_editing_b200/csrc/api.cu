// Error plumbing + version for the C ABI (include/rgie.h).
#include <atomic>
#include "common.cuh"
#include "rgie.h"

namespace rgie {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(const std::string& msg) {
  g_last_error = msg;
  return RGIE_ERR;
}
static std::atomic<long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace rgie

extern "C" {
long rgie_launch_count(void) { return rgie::g_launches.load(); }
int rgie_version(void) { return RGIE_ABI_VERSION; }
const char* rgie_last_error(void) { return rgie::g_last_error.c_str(); }
}
