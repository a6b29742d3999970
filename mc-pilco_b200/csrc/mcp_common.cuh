// Common device/host helpers for libmcpilco_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <mutex>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mcpilco_b200.h"

namespace mcp {

// ---- host-side error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// bench.py's roofline hook: bracket a launch of the dominant kernel with events (no-ops unless enabled)
void prof_begin(cudaStream_t st);
void prof_end(cudaStream_t st, double flops);

#define MCP_CHECK_ARG(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      ::mcp::set_error(__VA_ARGS__);      \
      return MCP_E_ARG;                   \
    }                                     \
  } while (0)

#define MCP_CUDA(expr)                                                                    \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::mcp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MCP_E_CUDA;                                                                  \
    }                                                                                     \
  } while (0)

#define MCP_LAUNCH_CHECK()                  \
  do {                                      \
    ::mcp::count_launch();                  \
    MCP_CUDA(cudaGetLastError());           \
  } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Per-device one-time setup (cudaFuncSetAttribute is per function AND per device).  One process-wide mutex serialises every lazily
// initialised piece of host state (kernel attributes, side streams, driver entry points): a second thread must not see "done" before the
// first one has finished the setup.
constexpr int MCP_MAX_DEVICES = 64;
inline std::mutex& init_mutex() {
  static std::mutex m;
  return m;
}
template <class Kern>
static inline cudaError_t ensure_dynamic_smem(bool (&done)[MCP_MAX_DEVICES], Kern kern, int bytes) {
  std::lock_guard<std::mutex> lock(init_mutex());
  int d = 0;
  cudaError_t e = cudaGetDevice(&d);
  if (e != cudaSuccess) return e;
  if (d >= 0 && d < MCP_MAX_DEVICES && done[d]) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && d >= 0 && d < MCP_MAX_DEVICES) done[d] = true;
  return e;
}
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- device primitives --------------------------------------------------------------------------
// FP64 tensor-core MMA, native shape on sm_100a (SASS: DMMA.8x8x4).
//   A 8x4 row-major: lane holds A[lane>>2][lane&3];  B 4x8 col-major: lane holds B[lane&3][lane>>2];
//   C 8x8: lane holds C[lane>>2][2*(lane&3) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// 16-byte async copy global->shared with zero fill of the bytes past src_bytes (0, 8 or 16).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes));
}
// 8-byte async copy global->shared; src_bytes = 0 zero-fills
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src, int src_bytes) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- Philox4x32-10 (counter-based RNG; Salmon et al. 2011) ----------------------------------------
struct Philox4 {
  uint32_t v[4];
};
__device__ __forceinline__ Philox4 philox4x32_10(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  Philox4 r{{c0, c1, c2, c3}};
#pragma unroll
  for (int i = 0; i < 10; i++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, r.v[0]), lo0 = 0xD2511F53u * r.v[0];
    uint32_t hi1 = __umulhi(0xCD9E8D57u, r.v[2]), lo1 = 0xCD9E8D57u * r.v[2];
    Philox4 n{{hi1 ^ r.v[1] ^ k0, lo1, hi0 ^ r.v[3] ^ k1, lo0}};
    r = n;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return r;
}
// uniform in (0,1) from 2x32 bits (53-bit mantissa), never 0
__device__ __forceinline__ double u01(uint32_t a, uint32_t b) {
  uint64_t x = ((uint64_t)a << 21) ^ (uint64_t)(b >> 11) ^ ((uint64_t)(b & 0x7ffu) << 42);
  x &= ((1ull << 53) - 1);
  return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
}
// two standard normals by Box-Muller
__device__ __forceinline__ void box_muller(const Philox4& p, double& n0, double& n1) {
  double u = u01(p.v[0], p.v[1]), w = u01(p.v[2], p.v[3]);
  double r = sqrt(-2.0 * log(u));
  double s, c;
  sincospi(2.0 * w, &s, &c);
  n0 = r * c;
  n1 = r * s;
}
// streams of the rollout: which random object a counter addresses
enum { RNG_EPS = 0, RNG_MASK = 1, RNG_MEAS = 2, RNG_X0 = 3 };
// standard normal #j of (particle, t, stream)
__device__ __forceinline__ double rng_normal(uint64_t seed, uint64_t pid, int t, int stream, int j) {
  Philox4 p = philox4x32_10(seed, (uint32_t)pid, (uint32_t)(pid >> 32), (uint32_t)t | ((uint32_t)stream << 24), (uint32_t)(j >> 1));
  double a, b;
  box_muller(p, a, b);
  return (j & 1) ? b : a;
}
// dropout keep-mask for basis b of (particle, t): keep with probability 1-p
__device__ __forceinline__ bool rng_keep(uint64_t seed, uint64_t pid, int t, int b, double p) {
  Philox4 q = philox4x32_10(seed, (uint32_t)pid, (uint32_t)(pid >> 32), (uint32_t)t | ((uint32_t)RNG_MASK << 24), (uint32_t)(b >> 2));
  double u = (double)q.v[b & 3] * (1.0 / 4294967296.0);
  return u >= p;
}

}  // namespace mcp
