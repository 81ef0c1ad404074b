// Error state, launch counter and version of the C ABI (include/hamgnn_b200.h).
#include <atomic>
#include <cstdarg>

#include "hgb_common.cuh"

namespace hgb {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace hgb

extern "C" {
int hgb_abi_version(void) { return HGB_ABI_VERSION; }
const char* hgb_last_error(void) { return hgb::g_err; }
int64_t hgb_launch_count(void) { return (int64_t)hgb::g_launches.load(std::memory_order_relaxed); }
}
