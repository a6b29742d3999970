// Device helpers shared by the rollout kernels (per-step path, fused small-shape path, backward).
#pragma once
#include "mcp_common.cuh"

namespace mcp {

// ------------------------------------------------------------------------------------------------
// small device helpers shared by forward and backward
// ------------------------------------------------------------------------------------------------
// policy feature j of the state seen by the policy (already divided by scale_factor)
__device__ __forceinline__ double policy_feature(const McpPolicy& p, const double* __restrict__ x, int t, int j) {
  double f;
  if (p.kind == 1) {
    if (j < p.n_na) f = x[p.na_idx[j]];
    else if (j < p.n_na + p.n_a) f = cos(x[p.a_idx[j - p.n_na]]);
    else f = sin(x[p.a_idx[j - p.n_na - p.n_a]]);
  } else if (p.kind == 2) {
    f = (j < p.Ds) ? x[j] : (p.target_traj[(size_t)t * p.Ds + (j - p.Ds)] - x[j - p.Ds]);
  } else {
    f = x[j];
  }
  return f * p.inv_scale[j];
}

// Philox key of this rollout: the baked seed plus (optionally) a device word the host bumps between replays of a captured graph
__device__ __forceinline__ uint64_t noise_seed(const McpNoise& nz) { return nz.seed_dev ? nz.seed + __ldg(nz.seed_dev) : nz.seed; }

__device__ __forceinline__ bool dropout_active(const McpPolicy& p, const McpNoise& nz) { return p.use_drop && nz.p_dropout > 0.0; }

// tm addresses the injected mask tensor, t the Philox counter (they differ only for the stand-alone policy call)
__device__ __forceinline__ bool keep_unit(const McpNoise& nz, int M, int nb, int tm, int t, int m, int b) {
  if (nz.masks) return nz.masks[((size_t)tm * M + m) * nb + b] != 0;
  return rng_keep(noise_seed(nz), nz.particle_offset + (uint64_t)m, t, b, nz.p_dropout);
}

__device__ __forceinline__ double cost_value(const McpCost& c, const double* __restrict__ x, int t, int Ds) {
  if (c.kind == 1) {
    double a = (fabs(x[c.idx[0]]) - c.target[0]) * c.inv_ls[0], b = (x[c.idx[1]] - c.target[1]) * c.inv_ls[1];
    return 1.0 - exp(-(a * a) - b * b);
  }
  double d = 0.0;
  for (int i = 0; i < c.n_idx; i++) {
    double tg = (c.kind == 2) ? c.target_traj[(size_t)t * Ds + c.idx[i]] : c.target[i];
    double r = (x[c.idx[i]] - tg) * c.inv_ls[i];
    d = fma(r, r, d);
  }
  return (c.kind == 4) ? d : 1.0 - exp(-d);
}

// lam[j] += w * d cost / d x_j
__device__ __forceinline__ void cost_grad_add(const McpCost& c, const double* __restrict__ x, int t, int Ds, double w, double* lam) {
  if (c.kind == 1) {
    double th = x[c.idx[0]];
    double a = (fabs(th) - c.target[0]) * c.inv_ls[0], b = (x[c.idx[1]] - c.target[1]) * c.inv_ls[1];
    double e = exp(-(a * a) - b * b);
    double sg = (th > 0.0) ? 1.0 : ((th < 0.0) ? -1.0 : 0.0);
    lam[c.idx[0]] += w * e * 2.0 * a * c.inv_ls[0] * sg;
    lam[c.idx[1]] += w * e * 2.0 * b * c.inv_ls[1];
    return;
  }
  double d = 0.0;
  for (int i = 0; i < c.n_idx; i++) {
    double tg = (c.kind == 2) ? c.target_traj[(size_t)t * Ds + c.idx[i]] : c.target[i];
    double r = (x[c.idx[i]] - tg) * c.inv_ls[i];
    d = fma(r, r, d);
  }
  double e = (c.kind == 4) ? 1.0 : exp(-d);
  for (int i = 0; i < c.n_idx; i++) {
    double tg = (c.kind == 2) ? c.target_traj[(size_t)t * Ds + c.idx[i]] : c.target[i];
    lam[c.idx[i]] += w * e * 2.0 * (x[c.idx[i]] - tg) * c.inv_ls[i] * c.inv_ls[i];
  }
}


}  // namespace mcp
