// Error state, launch counter and version of the C ABI (include/hamgnn_b200.h).
#include <atomic>
#include <cstdarg>
#include <mutex>
#include <vector>

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

// ---- per-kernel device timing (bench.py's live roofline lines): CUDA events on the launching stream around every kernel
// launch of the library while enabled; hgb_timing_collect() synchronises them and sums per kernel id.
static bool g_timing = false;
struct TimeRec { cudaEvent_t a, b; int id; };
static std::vector<TimeRec> g_recs;
static std::mutex g_tmu;

bool timing_on() { return g_timing; }
void timing_begin(int id, void* stream) {
  if (!g_timing) return;
  TimeRec r;
  r.id = id;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, (cudaStream_t)stream);
  std::lock_guard<std::mutex> lk(g_tmu);
  g_recs.push_back(r);
}
void timing_end(int id, void* stream) {
  if (!g_timing) return;
  std::lock_guard<std::mutex> lk(g_tmu);
  for (size_t i = g_recs.size(); i-- > 0;)
    if (g_recs[i].id == id) { cudaEventRecord(g_recs[i].b, (cudaStream_t)stream); break; }
}
}  // namespace hgb

extern "C" {
int hgb_abi_version(void) { return HGB_ABI_VERSION; }
const char* hgb_last_error(void) { return hgb::g_err; }
int64_t hgb_launch_count(void) { return (int64_t)hgb::g_launches.load(std::memory_order_relaxed); }

int hgb_timing_enable(int32_t on) {
  hgb::g_timing = on != 0;
  return 0;
}
// ms_out[id] += elapsed milliseconds, count_out[id] += launches, for id < HGB_N_KERNEL_IDS; clears the records
int hgb_timing_collect(float* ms_out, int64_t* count_out) {
  std::lock_guard<std::mutex> lk(hgb::g_tmu);
  for (auto& r : hgb::g_recs) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess && r.id >= 0 &&
        r.id < HGB_N_KERNEL_IDS) {
      if (ms_out) ms_out[r.id] += ms;
      if (count_out) count_out[r.id] += 1;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  hgb::g_recs.clear();
  (void)cudaGetLastError();
  return 0;
}
}
