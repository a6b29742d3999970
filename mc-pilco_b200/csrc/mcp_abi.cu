// Error/launch bookkeeping and the small informational entry points of the C ABI.
#include <stdarg.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "mcp_common.cuh"

namespace mcp {
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// ---- per-launch timing of the dominant kernel (enabled only by bench.py) ----
struct ProfState {
  std::mutex mu;
  bool on = false;
  std::vector<cudaEvent_t> ev;  // pairs: begin, end
  size_t used = 0;              // events handed out since the last read
  double flops = 0.0;
  bool open = false;
};
static ProfState g_prof;
constexpr size_t PROF_MAX_EVENTS = 1 << 17;
void prof_begin(cudaStream_t st) {
  if (!g_prof.on) return;
  std::lock_guard<std::mutex> lk(g_prof.mu);
  if (g_prof.used + 2 > PROF_MAX_EVENTS) return;
  while (g_prof.ev.size() < g_prof.used + 2) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    g_prof.ev.push_back(e);
  }
  cudaEventRecord(g_prof.ev[g_prof.used], st);
  g_prof.open = true;
}
void prof_end(cudaStream_t st, double flops) {
  if (!g_prof.on) return;
  std::lock_guard<std::mutex> lk(g_prof.mu);
  if (!g_prof.open) return;
  cudaEventRecord(g_prof.ev[g_prof.used + 1], st);
  g_prof.used += 2;
  g_prof.flops += flops;
  g_prof.open = false;
}
}  // namespace mcp

extern "C" __attribute__((visibility("default"))) int mcpilco_abi_version(void) { return MCP_ABI_VERSION; }
extern "C" __attribute__((visibility("default"))) const char* mcpilco_last_error(void) { return mcp::g_err; }
extern "C" __attribute__((visibility("default"))) uint64_t mcpilco_launch_count(int reset) {
  return reset ? mcp::g_launches.exchange(0) : mcp::g_launches.load();
}
extern "C" __attribute__((visibility("default"))) int mcpilco_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  MCP_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  MCP_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return MCP_OK;
}
extern "C" __attribute__((visibility("default"))) int mcpilco_set_device(int device) {
  MCP_CUDA(cudaSetDevice(device));
  return MCP_OK;
}
// sizeof of every ABI struct, in header order, so that a foreign-language binding can verify its layout
extern "C" __attribute__((visibility("default"))) int mcpilco_struct_sizes(size_t* out, int n) {
  const size_t s[] = {sizeof(McpGpSpec), sizeof(McpGp),    sizeof(McpModel),   sizeof(McpPolicy),     sizeof(McpCost),
                      sizeof(McpMeas),   sizeof(McpNoise), sizeof(McpRollout), sizeof(McpRolloutGrad)};
  const int k = (int)(sizeof(s) / sizeof(s[0]));
  for (int i = 0; i < n && i < k; i++) out[i] = s[i];
  return k;
}

extern "C" __attribute__((visibility("default"))) int mcpilco_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(mcp::g_prof.mu);
  mcp::g_prof.on = on != 0;
  mcp::g_prof.used = 0;
  mcp::g_prof.flops = 0.0;
  mcp::g_prof.open = false;
  return MCP_OK;
}
extern "C" __attribute__((visibility("default"))) int mcpilco_prof_read(double* total_ms, uint64_t* launches, double* flops) {
  std::lock_guard<std::mutex> lk(mcp::g_prof.mu);
  double ms = 0.0;
  for (size_t i = 0; i + 1 < mcp::g_prof.used; i += 2) {
    MCP_CUDA(cudaEventSynchronize(mcp::g_prof.ev[i + 1]));
    float t = 0.f;
    MCP_CUDA(cudaEventElapsedTime(&t, mcp::g_prof.ev[i], mcp::g_prof.ev[i + 1]));
    ms += t;
  }
  if (total_ms) *total_ms = ms;
  if (launches) *launches = mcp::g_prof.used / 2;
  if (flops) *flops = mcp::g_prof.flops;
  mcp::g_prof.used = 0;
  mcp::g_prof.flops = 0.0;
  return MCP_OK;
}
