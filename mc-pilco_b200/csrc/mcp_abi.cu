// Error/launch bookkeeping and the small informational entry points of the C ABI.
#include <stdarg.h>

#include <atomic>

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
