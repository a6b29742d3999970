// Greedy subset-of-data selection on the device — GP_prior.get_SOD (gpr_lib/GP_prior/GP_prior.py:232-257), SURVEY.md §8 f3.
// The reference refits the GP on the current subset from scratch for every candidate (O(N |S|^3) with a host round trip each);
// here one CTA keeps the Cholesky factor of K_S + sn2 I and grows it by one row per accepted point:
//     v = L^-1 k(X_S, x_i),   var_i = k(x_i, x_i) - v^T v,   accept if sqrt(var_i) > threshold  (NaN compares false, as in torch),
//     on accept:  L <- [[L, 0], [v^T, sqrt(var_i + sn2)]].
// Same quantity as the reference's predictive variance without noise (GP_prior.py:152 through get_estimate), O(N |S|^2) in total and
// no host synchronisation inside the loop.  L is kept transposed (Lt[s][r] = L[r][s]) so the column sweep of the forward
// substitution reads contiguous memory.
#include "mcp_kfn.cuh"

namespace mcp {

template <int DT>
__global__ void __launch_bounds__(512) sod_select_kernel(const __grid_constant__ McpGpSpec s, const double* __restrict__ X, int N,
                                                         const int* __restrict__ order, double threshold, int* __restrict__ idx_out,
                                                         int* __restrict__ count_out, double* __restrict__ Lt, int cap,
                                                         double* __restrict__ XS, double* __restrict__ vbuf) {
  __shared__ double s_red[16];
  __shared__ double s_pivot;
  __shared__ int s_count;
  const int tid = threadIdx.x, nt = blockDim.x, D = s.D;
  // seed: the first point of the order
  if (tid == 0) {
    const int i0 = order ? order[0] : 0;
    double x[DT];
    KFn<DT>::load(x, X + (size_t)i0 * D, D);
    Lt[0] = sqrt(KFn<DT>::kdiag(s, x) + s.sigma_n2);
    idx_out[0] = i0;
    s_count = 1;
  }
  if (tid < D) XS[tid] = X[(size_t)(order ? order[0] : 0) * D + tid];
  __syncthreads();
  for (int c = 1; c < N; c++) {
    const int i = order ? order[c] : c, n = s_count;
    double xi[DT];
    KFn<DT>::load(xi, X + (size_t)i * D, D);
    // k(X_S, x_i)
    for (int r = tid; r < n; r += nt) {
      double y[DT];
      KFn<DT>::load(y, XS + (size_t)r * D, D);
      vbuf[r] = KFn<DT>::k(s, y, xi);
    }
    __syncthreads();
    // forward substitution, column sweep: v_s = b_s / L_ss; b_r -= L_rs v_s for r > s
    double vv = 0.0;
    for (int sidx = 0; sidx < n; sidx++) {
      if (tid == 0) s_pivot = vbuf[sidx] / Lt[(size_t)sidx * cap + sidx];
      __syncthreads();
      const double vs = s_pivot;
      if (tid == 0) { vbuf[sidx] = vs; }
      const double* col = Lt + (size_t)sidx * cap;
      for (int r = sidx + 1 + tid; r < n; r += nt) vbuf[r] = fma(-col[r], vs, vbuf[r]);
      if (tid == 0) vv = fma(vs, vs, vv);
      __syncthreads();
    }
    if (tid == 0) {
      const double var = KFn<DT>::kdiag(s, xi) - vv;
      if (sqrt(var) > threshold) {
        for (int r = 0; r < n; r++) Lt[(size_t)r * cap + n] = vbuf[r];
        Lt[(size_t)n * cap + n] = sqrt(var + s.sigma_n2);
        idx_out[n] = i;
        s_count = n + 1;
        s_pivot = 1.0;
      } else {
        s_pivot = 0.0;
      }
    }
    __syncthreads();
    if (s_pivot != 0.0 && tid < D) XS[(size_t)n * D + tid] = X[(size_t)i * D + tid];
    __syncthreads();
  }
  if (tid == 0) *count_out = s_count;
}

}  // namespace mcp

using namespace mcp;

extern "C" __attribute__((visibility("default"))) size_t mcpilco_gp_sod_workspace_bytes(int N) {
  const size_t n = (size_t)(N > 0 ? N : 1);
  return (n * n + n * MCP_MAX_D + n) * sizeof(double) + 1024;
}

extern "C" __attribute__((visibility("default"))) int mcpilco_gp_sod_select(const McpGpSpec* spec, const double* X, int N, const int* order,
                                                                             double threshold, int* idx_out, int* count_out, void* workspace,
                                                                             size_t workspace_bytes, void* stream) {
  MCP_CHECK_ARG(spec && X && idx_out && count_out && N >= 1, "gp_sod_select: bad arguments");
  MCP_CHECK_ARG(spec->D >= 1 && spec->D <= MCP_MAX_D, "gp_sod_select: gp input dim %d outside [1,%d]", spec->D, MCP_MAX_D);
  MCP_CHECK_ARG(workspace && workspace_bytes >= mcpilco_gp_sod_workspace_bytes(N), "gp_sod_select: workspace too small");
  double* Lt = (double*)align_up((size_t)workspace, 256);
  double* XS = Lt + (size_t)N * N;
  double* vbuf = XS + (size_t)N * MCP_MAX_D;
  MCP_DISPATCH_D(spec->D, (sod_select_kernel<DT><<<1, 512, 0, (cudaStream_t)stream>>>(*spec, X, N, order, threshold, idx_out, count_out, Lt, N, XS, vbuf)));
  MCP_LAUNCH_CHECK();
  return MCP_OK;
}
