// Negative marginal log likelihood of one GP and its gradient with respect to the kernel hyper-parameters — the training
// objective of GP_prior.fit_model (gpr_lib/GP_prior/GP_prior.py:179-230) with Marginal_log_likelihood
// (gpr_lib/Likelihood/Gaussian_likelihood.py:12-24):
//
//     L = 0.5 * ( (y - m)^T K^-1 (y - m) + log det K )          (the reference drops the N log 2 pi constant)
//
// The reference differentiates through torch.cholesky / torch.inverse with autograd.  Here the factorisation is the
// precompute of the rollout path (blocked Cholesky, triangular inverse, K^-1, alpha) and the gradient is analytic:
//
//     dL/dtheta = 0.5 * sum_ij G_ij dK_ij/dtheta ,   G = K^-1 - alpha alpha^T ,      dL/dm = -sum_i alpha_i
//
// evaluated by one pass over the N x N pairs per parameter group (SE parameters; one group per polynomial factor), with
// per-block partial sums reduced in a fixed order (bit-stable).  Gradients are returned with respect to the FIELDS of
// McpGpSpec (inv_ls, lambda, poly_w2, sigma_n2, mean0); the host layer applies the chain rule of its own parametrisation.
#include "mcp_kfn.cuh"

namespace mcp {

constexpr int NL_ROWS = 64;

// mode -1: SE group  -> acc[0..DT) d/d inv_ls, acc[DT] d/d lambda, acc[DT+1] d/d sigma_n2 (trace of G)
// mode p * MCP_MAX_DEG + f: polynomial factor (p, f) -> acc[0..DT) d/d w2[p][f][j], acc[DT] d/d offset
template <int DT>
__global__ void __launch_bounds__(256) nlml_grad_kernel(const __grid_constant__ McpGpSpec s, const double* __restrict__ X, int N,
                                                        const double* __restrict__ alpha, const double* __restrict__ Kinv, int ld,
                                                        int mode, double* __restrict__ partials) {
  __shared__ double sx[NL_ROWS][DT];
  __shared__ double sa[NL_ROWS];
  __shared__ double red[8][DT + 2];
  const int D = s.D, i0 = blockIdx.y * NL_ROWS, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int el = tid; el < NL_ROWS * DT; el += 256) {
    const int r = el / DT, j = el - r * DT;
    sx[r][j] = (i0 + r < N && j < D) ? X[(size_t)(i0 + r) * D + j] : 0.0;
  }
  if (tid < NL_ROWS) sa[tid] = (i0 + tid < N) ? alpha[i0 + tid] : 0.0;
  __syncthreads();
  const int c = blockIdx.x * 256 + tid;
  double acc[DT + 2];
#pragma unroll
  for (int k = 0; k < DT + 2; k++) acc[k] = 0.0;
  if (c < N) {
    double y[DT];
    KFn<DT>::load(y, X + (size_t)c * D, D);
    const double ac = alpha[c];
    const int rows = min(NL_ROWS, N - i0);
    const int p = mode >= 0 ? mode / MCP_MAX_DEG : 0, f = mode >= 0 ? mode % MCP_MAX_DEG : 0;
    for (int r = 0; r < rows; r++) {
      const double G = Kinv[(size_t)(i0 + r) * ld + c] - sa[r] * ac;
      if (mode < 0) {
        double d2 = 0.0, dd[DT];
#pragma unroll
        for (int j = 0; j < DT; j++) {
          const double t = (sx[r][j] - y[j]);
          dd[j] = t * t * s.inv_ls[j];
          d2 = fma(dd[j], s.inv_ls[j], d2);
        }
        const double e = s.has_se ? exp(-d2) : 0.0, ge = G * e;
#pragma unroll
        for (int j = 0; j < DT; j++) acc[j] = fma(ge, dd[j], acc[j]);  // x (-2 lambda) at the end
        acc[DT] += ge;
        if (i0 + r == c) acc[DT + 1] += G;
      } else {
        double cf = 1.0;
#pragma unroll
        for (int g2 = 0; g2 < MCP_MAX_DEG; g2++) {
          double L = s.poly_w2[p][g2][MCP_MAX_D];
#pragma unroll
          for (int j = 0; j < DT; j++) L = fma(s.poly_w2[p][g2][j] * sx[r][j], y[j], L);
          if (g2 != f) cf *= L;
        }
        const double gc = G * cf;
#pragma unroll
        for (int j = 0; j < DT; j++) acc[j] = fma(gc, sx[r][j] * y[j], acc[j]);
        acc[DT] += gc;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < DT + 2; k++) {
    const double v = warp_sum(acc[k]);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (tid < DT + 2) {
    double v = 0.0;
    for (int w = 0; w < 8; w++) v += red[w][tid];
    partials[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (DT + 2) + tid] = v;
  }
}

// out layout (doubles): [0] nlml, [1] d/dlambda, [2] d/dmean0, [3] d/dsigma_n2, [4 .. 4+MAX_D) d/dinv_ls,
// [4+MAX_D + ((p*MAX_DEG + f) * (MAX_D+1)) + j] d/dpoly_w2[p][f][j]
__global__ void nlml_finish_group_kernel(const __grid_constant__ McpGpSpec s, const double* __restrict__ partials, int nblocks, int width,
                                         int mode, double* __restrict__ out) {
  const int k = threadIdx.x;
  if (k >= width) return;
  double v = 0.0;
  for (int b = 0; b < nblocks; b++) v += partials[(size_t)b * width + k];
  const int DT = width - 2;
  if (mode < 0) {
    if (k < DT) { if (k < s.D) out[4 + k] = 0.5 * (-2.0 * s.lambda) * v; }
    else if (k == DT) out[1] = 0.5 * v;
    else out[3] = 0.5 * v;
  } else {
    const int base = 4 + MCP_MAX_D + mode * (MCP_MAX_D + 1);
    if (k < DT) { if (k < s.D) out[base + k] = 0.5 * v; }
    else if (k == DT) out[base + MCP_MAX_D] = 0.5 * v;
  }
}

// nlml = 0.5 (sum_i (y_i - m) alpha_i + 2 sum_i log L_ii);  d/dmean0 = -sum_i alpha_i
__global__ void __launch_bounds__(256) nlml_value_kernel(const double* __restrict__ y, const double* __restrict__ alpha,
                                                         const double* __restrict__ Lfac, int ld, int N, double mean0,
                                                         double* __restrict__ out) {
  __shared__ double red[3][8];
  double a = 0.0, b = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < N; i += 256) {
    a = fma(y[i] - mean0, alpha[i], a);
    b += log(Lfac[(size_t)i * ld + i]);
    c += alpha[i];
  }
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; red[2][threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sa = 0.0, sb = 0.0, sc = 0.0;
    for (int w = 0; w < 8; w++) { sa += red[0][w]; sb += red[1][w]; sc += red[2][w]; }
    out[0] = 0.5 * (sa + 2.0 * sb);
    out[2] = -sc;
  }
}

}  // namespace mcp

using namespace mcp;

extern "C" __attribute__((visibility("default"))) int mcpilco_gp_nlml_grad_size(void) {
  return 4 + MCP_MAX_D + MCP_MAX_POLY * MCP_MAX_DEG * (MCP_MAX_D + 1);
}

extern "C" __attribute__((visibility("default"))) size_t mcpilco_gp_nlml_workspace_bytes(int N) {
  size_t n = (size_t)(N > 0 ? N : 1), ld = n + (n & 1);
  size_t blocks = (size_t)cdiv((int)n, 256) * cdiv((int)n, NL_ROWS);
  return mcpilco_gp_precompute_workspace_bytes(N) + (2 * n * ld + n + blocks * (MCP_MAX_D + 2)) * sizeof(double) + 1024;
}

extern "C" __attribute__((visibility("default"))) int mcpilco_gp_nlml(const McpGpSpec* spec, const double* X, const double* y, int N,
                                                                       double* out, void* workspace, size_t workspace_bytes, void* stream) {
  MCP_CHECK_ARG(spec && X && y && out && N >= 1, "gp_nlml: bad arguments");
  MCP_CHECK_ARG(workspace && workspace_bytes >= mcpilco_gp_nlml_workspace_bytes(N), "gp_nlml: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)N, ld = n + (n & 1);
  double* p = (double*)align_up((size_t)workspace, 256);
  double* Kinv = p; p += n * ld;
  double* Lfac = p; p += n * ld;
  double* alpha = p; p += align_up(n, 32);
  double* partials = p; p += align_up((size_t)cdiv(N, 256) * cdiv(N, NL_ROWS) * (MCP_MAX_D + 2), 32);
  void* pre_ws = p;
  size_t pre_bytes = workspace_bytes - ((char*)pre_ws - (char*)workspace);
  MCP_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * mcpilco_gp_nlml_grad_size(), st));
  if (int e = mcpilco_gp_precompute(spec, X, y, N, alpha, Kinv, (int)ld, Lfac, nullptr, pre_ws, pre_bytes, stream)) return e;
  nlml_value_kernel<<<1, 256, 0, st>>>(y, alpha, Lfac, (int)ld, N, spec->mean0, out);
  MCP_LAUNCH_CHECK();
  dim3 grid(cdiv(N, 256), cdiv(N, NL_ROWS));
  const int nblocks = grid.x * grid.y;
  auto group = [&](int mode) -> int {
    int width = 0;
    MCP_DISPATCH_D(spec->D, (width = DT + 2, nlml_grad_kernel<DT><<<grid, 256, 0, st>>>(*spec, X, N, alpha, Kinv, (int)ld, mode, partials)));
    MCP_LAUNCH_CHECK();
    nlml_finish_group_kernel<<<1, 64, 0, st>>>(*spec, partials, nblocks, width, mode, out);
    MCP_LAUNCH_CHECK();
    return MCP_OK;
  };
  if (int e = group(-1)) return e;
  for (int pp = 0; pp < spec->n_poly; pp++)
    for (int f = 0; f < spec->poly_deg[pp]; f++)
      if (int e = group(pp * MCP_MAX_DEG + f)) return e;
  return MCP_OK;
}
