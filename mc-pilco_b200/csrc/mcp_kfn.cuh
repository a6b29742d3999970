// Covariance function of one GP (SE + sum of multiplicative polynomial kernels) and its derivative
// with respect to the first argument.  Closed form of the reference's Sum_Independent_GP tree:
//   gpr_lib/GP_prior/GP_prior.py:314-347, Stationary_GP.py:65-109,162-181, Sparse_GP.py:391-453,613-668.
// Convention (set by the host when it fills McpGpSpec): unused factor slots f >= poly_deg[p] hold
// w2[j] = 0 and offset = 1 (a neutral factor), so device code never branches on the degree.
#pragma once
#include "mcp_common.cuh"

namespace mcp {

template <int DT>
struct KFn {
  // load a row of D doubles into a zero-padded register array
  static __device__ __forceinline__ void load(double (&x)[DT], const double* __restrict__ p, int D) {
#pragma unroll
    for (int j = 0; j < DT; j++) x[j] = (j < D) ? p[j] : 0.0;
  }

  static __device__ __forceinline__ double lin(const McpGpSpec& s, int p, int f, const double (&x)[DT], const double (&y)[DT]) {
    double a = s.poly_w2[p][f][MCP_MAX_D];
#pragma unroll
    for (int j = 0; j < DT; j++) a = fma(s.poly_w2[p][f][j] * x[j], y[j], a);
    return a;
  }

  // k(x, y)
  static __device__ __forceinline__ double k(const McpGpSpec& s, const double (&x)[DT], const double (&y)[DT]) {
    double kv = 0.0;
    if (s.has_se) {
      double d2 = 0.0;
#pragma unroll
      for (int j = 0; j < DT; j++) {
        double t = (x[j] - y[j]) * s.inv_ls[j];
        d2 = fma(t, t, d2);
      }
      kv = s.lambda * exp(-d2);
    }
#pragma unroll
    for (int p = 0; p < MCP_MAX_POLY; p++) {
      if (p < s.n_poly) {
        double pr = 1.0;
#pragma unroll
        for (int f = 0; f < MCP_MAX_DEG; f++) pr *= lin(s, p, f, x, y);
        kv += pr;
      }
    }
    return kv;
  }

  // k(x, y) and dk/dx
  static __device__ __forceinline__ void k_grad(const McpGpSpec& s, const double (&x)[DT], const double (&y)[DT], double& kv,
                                                double (&dk)[DT]) {
    kv = 0.0;
#pragma unroll
    for (int j = 0; j < DT; j++) dk[j] = 0.0;
    if (s.has_se) {
      double d2 = 0.0;
#pragma unroll
      for (int j = 0; j < DT; j++) {
        double t = (x[j] - y[j]) * s.inv_ls[j];
        d2 = fma(t, t, d2);
      }
      double kse = s.lambda * exp(-d2);
      kv = kse;
      double m2k = -2.0 * kse;
#pragma unroll
      for (int j = 0; j < DT; j++) dk[j] = m2k * (s.inv_ls[j] * s.inv_ls[j]) * (x[j] - y[j]);
    }
#pragma unroll
    for (int p = 0; p < MCP_MAX_POLY; p++) {
      if (p < s.n_poly) {
        double L[MCP_MAX_DEG];
#pragma unroll
        for (int f = 0; f < MCP_MAX_DEG; f++) L[f] = lin(s, p, f, x, y);
        double pr = 1.0;
#pragma unroll
        for (int f = 0; f < MCP_MAX_DEG; f++) pr *= L[f];
        kv += pr;
#pragma unroll
        for (int f = 0; f < MCP_MAX_DEG; f++) {
          double c = 1.0;
#pragma unroll
          for (int g = 0; g < MCP_MAX_DEG; g++)
            if (g != f) c *= L[g];
#pragma unroll
          for (int j = 0; j < DT; j++) dk[j] = fma(c * s.poly_w2[p][f][j], y[j], dk[j]);
        }
      }
    }
  }

  // k(x, x) without noise and its gradient
  static __device__ __forceinline__ void kdiag_grad(const McpGpSpec& s, const double (&x)[DT], double& kd, double (&dkd)[DT]) {
    kd = s.has_se ? s.lambda : 0.0;
#pragma unroll
    for (int j = 0; j < DT; j++) dkd[j] = 0.0;
#pragma unroll
    for (int p = 0; p < MCP_MAX_POLY; p++) {
      if (p < s.n_poly) {
        double L[MCP_MAX_DEG];
#pragma unroll
        for (int f = 0; f < MCP_MAX_DEG; f++) L[f] = lin(s, p, f, x, x);
        double pr = 1.0;
#pragma unroll
        for (int f = 0; f < MCP_MAX_DEG; f++) pr *= L[f];
        kd += pr;
#pragma unroll
        for (int f = 0; f < MCP_MAX_DEG; f++) {
          double c = 2.0;
#pragma unroll
          for (int g = 0; g < MCP_MAX_DEG; g++)
            if (g != f) c *= L[g];
#pragma unroll
          for (int j = 0; j < DT; j++) dkd[j] = fma(c * s.poly_w2[p][f][j], x[j], dkd[j]);
        }
      }
    }
  }
  static __device__ __forceinline__ double kdiag(const McpGpSpec& s, const double (&x)[DT]) {
    double kd = s.has_se ? s.lambda : 0.0;
#pragma unroll
    for (int p = 0; p < MCP_MAX_POLY; p++) {
      if (p < s.n_poly) {
        double pr = 1.0;
#pragma unroll
        for (int f = 0; f < MCP_MAX_DEG; f++) pr *= lin(s, p, f, x, x);
        kd += pr;
      }
    }
    return kd;
  }
};

// dispatch a functor templated on the padded input dimension
#define MCP_DISPATCH_D(D, ...)                         \
  do {                                                 \
    if ((D) <= 4) { constexpr int DT = 4; __VA_ARGS__; }        \
    else if ((D) <= 6) { constexpr int DT = 6; __VA_ARGS__; }   \
    else if ((D) <= 8) { constexpr int DT = 8; __VA_ARGS__; }   \
    else if ((D) <= 12) { constexpr int DT = 12; __VA_ARGS__; } \
    else if ((D) <= 16) { constexpr int DT = 16; __VA_ARGS__; } \
    else if ((D) <= 24) { constexpr int DT = 24; __VA_ARGS__; } \
    else { constexpr int DT = 32; __VA_ARGS__; }                \
  } while (0)

}  // namespace mcp
