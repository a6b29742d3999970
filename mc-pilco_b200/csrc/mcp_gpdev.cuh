// Device-side table of fitted GPs: lets ONE launch serve all E outputs of a model (blockIdx.z / .y = output) instead of one launch
// per output with the specification passed by value.  Used by the fused small-shape path and the batched per-step posterior.
#pragma once
#include "mcp_common.cuh"

namespace mcp {

struct McpGpDev {
  McpGpSpec spec;
  int N, ld;
  const double* Xtr;
  const double* alpha;
  const double* Kinv;
  double var_scale;
};

// fill `host_tab[0..E)` from the rollout's GP descriptors
static inline void gpdev_fill(McpGpDev* host_tab, const McpGp* gps, int E) {
  for (int e = 0; e < E; e++) {
    host_tab[e].spec = gps[e].spec;
    host_tab[e].N = gps[e].N;
    host_tab[e].ld = gps[e].ld_kinv;
    host_tab[e].Xtr = gps[e].Xtr;
    host_tab[e].alpha = gps[e].alpha;
    host_tab[e].Kinv = gps[e].Kinv;
    host_tab[e].var_scale = gps[e].var_scale;
  }
}

// Upload of the table: one tiny kernel per GP that takes the descriptor BY VALUE (a kernel parameter is copied when the launch is
// recorded), instead of an asynchronous host -> device copy from a host stack array — the copy would not survive stream capture into
// a CUDA graph (the graph would re-read a dead stack frame on replay), the kernel does.
__global__ void gpdev_store_kernel(const __grid_constant__ McpGpDev g, McpGpDev* __restrict__ dst);
static inline cudaError_t gpdev_upload(McpGpDev* dev_tab, const McpGp* gps, int E, cudaStream_t st) {
  McpGpDev host_tab[MCP_MAX_E];
  gpdev_fill(host_tab, gps, E);
  for (int e = 0; e < E; e++) gpdev_store_kernel<<<1, 32, 0, st>>>(host_tab[e], dev_tab + e);
  return cudaGetLastError();
}

// Programmatic dependent launch (sm_90+): a kernel launched with programmatic stream serialisation may start while its predecessor
// drains; it must wait here before touching anything the predecessor wrote.  A no-op without the launch attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static cudaError_t launch_chain(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

static inline bool pdl_enabled() {
  static const bool on = getenv("MCPILCO_NO_PDL") == nullptr;
  return on;
}

// batched V_e = K*_e Kinv_e for all outputs in one launch (mcp_small.cu)
int launch_small_gemm(const McpGpDev* tab, int M, int nmax, int E, const double* Ks, double* V, int ldk, size_t gp_stride, bool pdl,
                      cudaStream_t st);

}  // namespace mcp
