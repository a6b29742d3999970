// GP side of the hot path: covariance evaluation, the per-model-update precompute (blocked Cholesky,
// triangular inverse, K^-1, alpha) and the posterior mean/variance with their input Jacobians.
// Reference behaviour being reproduced: gpr_lib/GP_prior/GP_prior.py:91-155 and the kernel classes
// cited in mcp_kfn.cuh; driven by model_learning/Model_learning.py:163-208,265-336.
#include <cooperative_groups.h>

#include "mcp_dgemm.cuh"
#include "mcp_gpdev.cuh"
#include "mcp_kfn.cuh"

namespace mcp {

// ------------------------------------------------------------------------------------------------
// covariance matrices
// ------------------------------------------------------------------------------------------------
// K[i][j] = k(X1_i, X2_j) for i < n1, j < n2; columns n2..ncols_out-1 are written as zero (row padding for
// the 16-byte tile loads of the GEMM).  add_noise adds sigma_n2 on the diagonal (X2 == X1 case).
template <int DT>
__device__ __forceinline__ void cov_body(const McpGpSpec& s, const double* __restrict__ X1, int n1, const double* __restrict__ X2, int n2,
                                         int add_noise, double* __restrict__ K, int ldk, int ncols_out) {
  int j = blockIdx.x * 32 + (threadIdx.x & 31);
  int i0 = blockIdx.y * 32 + (threadIdx.x >> 5) * 4;
  if (j >= ncols_out) return;
  double y[DT];
  if (j < n2) KFn<DT>::load(y, X2 + (size_t)j * s.D, s.D);
#pragma unroll
  for (int r = 0; r < 4; r++) {
    int i = i0 + r;
    if (i >= n1) break;
    double v = 0.0;
    if (j < n2) {
      double x[DT];
      KFn<DT>::load(x, X1 + (size_t)i * s.D, s.D);
      v = KFn<DT>::k(s, x, y);
      if (add_noise && i == j) v += s.sigma_n2;
    }
    K[(size_t)i * ldk + j] = v;
  }
}

template <int DT>
__global__ void __launch_bounds__(256) cov_kernel(const __grid_constant__ McpGpSpec s, const double* __restrict__ X1, int n1,
                                                  const double* __restrict__ X2, int n2, int add_noise, double* __restrict__ K,
                                                  int ldk, int ncols_out) {
  cov_body<DT>(s, X1, n1, X2, n2, add_noise, K, ldk, ncols_out);
}

// K* tiles for WIDE gp inputs (8 < D <= 32: SE + at most one linear term, the UR5 model).  The generic kernel keeps x[DT], y[DT] in
// registers and walks the whole McpGpSpec per entry (~100 loads); here one thread owns a training point (column), the block a strip of
// 8 particles (16 measured 2 us per UR5 step slower: too few CTAs): the training tile is staged coalesced, the particle rows as (x, w1 x) pairs read by broadcast, and the loop runs
// dimension-outer over the row accumulators.  Arithmetic per entry is exactly KFn::k's, so the result is bit-identical to cov_kernel.
constexpr int CW_THREADS = 128, CW_ROWS = 8;
static bool wide_reduce_ok(const McpGpSpec& s);
template <int DT>
__device__ __forceinline__ void cov_wide_body(const McpGpSpec& s, const double* __restrict__ X1, int n1, const double* __restrict__ X2, int n2,
                                              double* __restrict__ K, int ldk, int ncols_out) {
  __shared__ double sy[CW_THREADS][DT + 1];
  __shared__ double2 sxp[CW_ROWS][DT];
  __shared__ double sil[DT];
  const int D = s.D, tid = threadIdx.x, nb0 = blockIdx.x * CW_THREADS, n = nb0 + tid, i0 = blockIdx.y * CW_ROWS;
  const bool np1 = s.n_poly == 1;
  const int cnt = min(CW_THREADS, n2 - nb0);
  for (int el = tid; el < cnt * D; el += CW_THREADS) {  // asynchronous copy: the L2 round trips of the strip overlap instead of queueing
    const int r = el / D, j = el - r * D;
    cp_async8(&sy[r][j], X2 + (size_t)nb0 * D + el, 8);
  }
  cp_async_commit();
  {
    constexpr int NI = (CW_ROWS * DT + CW_THREADS - 1) / CW_THREADS;
    double xv[NI];
#pragma unroll
    for (int u = 0; u < NI; u++) {  // loads first, stores after
      const int el = tid + u * CW_THREADS, r = el / DT, j = el - r * DT;
      xv[u] = (el < CW_ROWS * DT && i0 + r < n1 && j < D) ? X1[(size_t)(i0 + r) * D + j] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < NI; u++) {
      const int el = tid + u * CW_THREADS, r = el / DT, j = el - r * DT;
      if (el < CW_ROWS * DT) {
        const double w = (np1 && j < D) ? s.poly_w2[0][0][j] : 0.0;
        sxp[r][j] = make_double2(xv[u], w * xv[u]);
      }
    }
  }
  if (tid < DT) sil[tid] = tid < D ? s.inv_ls[tid] : 0.0;
  cp_async_wait<0>();
  __syncthreads();
  if (n >= ncols_out) return;
  double d2[CW_ROWS], a[CW_ROWS];
  const double off = np1 ? s.poly_w2[0][0][MCP_MAX_D] : 0.0;
#pragma unroll
  for (int r = 0; r < CW_ROWS; r++) { d2[r] = 0.0; a[r] = off; }
  if (n < n2) {
    for (int j = 0; j < D; j++) {
      const double y = sy[tid][j], il = sil[j];
#pragma unroll
      for (int r = 0; r < CW_ROWS; r++) {
        const double2 xp = sxp[r][j];
        const double t = (xp.x - y) * il;
        d2[r] = fma(t, t, d2[r]);
        a[r] = fma(xp.y, y, a[r]);
      }
    }
  }
  const double lam = s.lambda;
  const bool se = s.has_se != 0;
#pragma unroll
  for (int r = 0; r < CW_ROWS; r++) {
    const int i = i0 + r;
    if (i >= n1) break;
    double kv = 0.0;
    if (n < n2) {
      if (se) kv = lam * exp(-d2[r]);
      if (np1) kv += a[r];
    }
    K[(size_t)i * ldk + n] = kv;
  }
}

template <int DT>
__global__ void __launch_bounds__(CW_THREADS) cov_wide_kernel(const __grid_constant__ McpGpSpec s, const double* __restrict__ X1, int n1,
                                                              const double* __restrict__ X2, int n2, double* __restrict__ K, int ldk,
                                                              int ncols_out) {
  cov_wide_body<DT>(s, X1, n1, X2, n2, K, ldk, ncols_out);
}

template <int DT>
__global__ void __launch_bounds__(CW_THREADS) cov_wide_batched_kernel(const McpGpDev* __restrict__ gps, const double* __restrict__ Xs, int M,
                                                                      double* __restrict__ Ks, int ldk, size_t gp_stride) {
  const McpGpDev& g = gps[blockIdx.z];
  cov_wide_body<DT>(g.spec, Xs, M, g.Xtr, g.N, Ks + blockIdx.z * gp_stride, ldk, ldk);
}

// K* tile kernel for the same kernel shapes as the fast reduce (SE + Volterra polynomial, D <= 8): one thread per training
// point (its inputs, pre-scaled, stay in registers), 64 particles per block staged in shared memory with their
// particle-scaled weights (all lanes read the same particle: broadcast loads), four rows in flight for ILP on exp().
constexpr int COV_ROWS = 64;
template <int DT, int NP>
__global__ void __launch_bounds__(256) cov_fast_kernel(const __grid_constant__ McpGpSpec s, const double* __restrict__ X1, int n1,
                                                       const double* __restrict__ X2, int n2, double* __restrict__ K, int ldk,
                                                       int ncols_out) {
  __shared__ double sx[COV_ROWS][4][DT];  // per particle: x*ils, w1*x, w2a*x, w2b*x
  const int D = s.D, i0 = blockIdx.y * COV_ROWS, tid = threadIdx.x;
  for (int el = tid; el < COV_ROWS * DT; el += 256) {
    const int r = el / DT, j = el - r * DT;
    const double xv = (i0 + r < n1 && j < D) ? X1[(size_t)(i0 + r) * D + j] : 0.0;
    sx[r][0][j] = xv * s.inv_ls[j];
    sx[r][1][j] = NP >= 1 ? xv * s.poly_w2[0][0][j] : 0.0;
    sx[r][2][j] = NP >= 2 ? xv * s.poly_w2[1][0][j] : 0.0;
    sx[r][3][j] = NP >= 2 ? xv * s.poly_w2[1][1][j] : 0.0;
  }
  __syncthreads();
  const int c = blockIdx.x * 256 + tid;
  if (c >= ncols_out) return;
  double y[DT], ys[DT];
#pragma unroll
  for (int j = 0; j < DT; j++) {
    y[j] = (c < n2 && j < D) ? X2[(size_t)c * D + j] : 0.0;
    ys[j] = y[j] * s.inv_ls[j];
  }
  const int rows = min(COV_ROWS, n1 - i0);
  double* out = K + (size_t)i0 * ldk + c;
  for (int r0 = 0; r0 < rows; r0 += 4) {
    double kv[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int r = min(r0 + u, COV_ROWS - 1);
      double d2 = 0.0;
#pragma unroll
      for (int j = 0; j < DT; j++) {
        const double t = sx[r][0][j] - ys[j];
        d2 = fma(t, t, d2);
      }
      double k = s.lambda * exp(-d2);
      if (NP >= 1) {
        double L1 = s.poly_w2[0][0][MCP_MAX_D];
#pragma unroll
        for (int j = 0; j < DT; j++) L1 = fma(sx[r][1][j], y[j], L1);
        k += L1;
      }
      if (NP >= 2) {
        double La = s.poly_w2[1][0][MCP_MAX_D], Lb = s.poly_w2[1][1][MCP_MAX_D];
#pragma unroll
        for (int j = 0; j < DT; j++) {
          La = fma(sx[r][2][j], y[j], La);
          Lb = fma(sx[r][3][j], y[j], Lb);
        }
        k = fma(La, Lb, k);
      }
      kv[u] = (c < n2) ? k : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (r0 + u < rows) out[(size_t)(r0 + u) * ldk] = kv[u];
  }
}

// K* tile kernel of the opt-in INT8 contraction: the same arithmetic per entry as cov_fast_kernel, but a thread owns FOUR consecutive
// training points so that, besides the fp64 K* row (the reduce streams it), it emits the row's balanced base-256 digit planes as packed
// 32-bit stores — the layout ozaki_slice_kernel writes (planes[row][seg][p][k], p = 0 most significant; mcp_ozaki.cu) — and K* is never
// read back for slicing.  The row scale 2^e must be known before the first entry: it comes from the Cauchy-Schwarz bound
// |k(x, x_n)| <= sqrt(k(x,x) max_n k(x_n,x_n)) instead of the exact row maximum (a few leading bits of 8 S are given away when the
// bound is loose; the representation stays exact to 2^(e - 8 S)).
template <int DT, int NP>
__global__ void __launch_bounds__(256) cov_slice_kernel(const __grid_constant__ McpGpSpec s, const double* __restrict__ X1, int n1,
                                                        const double* __restrict__ X2, int n2, double kdiag_max, double* __restrict__ K, int ldk,
                                                        int S, int nseg, int Ks, int Ksp, int8_t* __restrict__ planes,
                                                        int32_t* __restrict__ expo) {
  __shared__ double sx[COV_ROWS][4][DT];  // per particle: x*ils, w1*x, w2a*x, w2b*x
  __shared__ double sscale[COV_ROWS];     // 2^(8 S - e_row)
  const int D = s.D, i0 = blockIdx.y * COV_ROWS, tid = threadIdx.x;
  for (int el = tid; el < COV_ROWS * DT; el += 256) {
    const int r = el / DT, j = el - r * DT;
    const double xv = (i0 + r < n1 && j < D) ? X1[(size_t)(i0 + r) * D + j] : 0.0;
    sx[r][0][j] = xv * s.inv_ls[j];
    sx[r][1][j] = NP >= 1 ? xv * s.poly_w2[0][0][j] : 0.0;
    sx[r][2][j] = NP >= 2 ? xv * s.poly_w2[1][0][j] : 0.0;
    sx[r][3][j] = NP >= 2 ? xv * s.poly_w2[1][1][j] : 0.0;
  }
  if (tid < COV_ROWS) {
    int ex = 0;
    if (i0 + tid < n1) {
      double x[DT];
      KFn<DT>::load(x, X1 + (size_t)(i0 + tid) * D, D);
      const double bound = sqrt(KFn<DT>::kdiag(s, x) * kdiag_max) * (1.0 + 1e-9);
      ex = (bound > 0.0 && isfinite(bound)) ? max(ilogb(bound) + 3, -900) : 0;  // |k| 2^-e < 1/4 (see ozaki_slice_kernel)
      if (blockIdx.x == 0) expo[i0 + tid] = ex;
    }
    sscale[tid] = scalbn(1.0, 8 * S - ex);
  }
  __syncthreads();
  const int kk = (blockIdx.x * 256 + tid) * 4;  // first of this thread's four plane columns (Ksp is a multiple of 128)
  if (kk >= nseg * Ksp) return;
  const int seg = kk / Ksp, kq = kk - seg * Ksp, c0 = seg * Ks + kq;
  double y[4][DT], ys[4][DT];
  bool ok[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    ok[q] = kq + q < Ks && c0 + q < n2;
#pragma unroll
    for (int j = 0; j < DT; j++) {
      y[q][j] = (ok[q] && j < D) ? X2[(size_t)(c0 + q) * D + j] : 0.0;
      ys[q][j] = y[q][j] * s.inv_ls[j];
    }
  }
  const int rows = min(COV_ROWS, n1 - i0);
  const unsigned long long bias = S >= 8 ? 0x8080808080808080ull : ((1ull << (8 * S)) - 1ull) / 255ull * 128ull;
  const size_t row_bytes = (size_t)nseg * S * Ksp;
  for (int r = 0; r < rows; r++) {
    double kv[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      double d2 = 0.0;
#pragma unroll
      for (int j = 0; j < DT; j++) {
        const double t = sx[r][0][j] - ys[q][j];
        d2 = fma(t, t, d2);
      }
      double k = s.lambda * exp(-d2);
      if (NP >= 1) {
        double L1 = s.poly_w2[0][0][MCP_MAX_D];
#pragma unroll
        for (int j = 0; j < DT; j++) L1 = fma(sx[r][1][j], y[q][j], L1);
        k += L1;
      }
      if (NP >= 2) {
        double La = s.poly_w2[1][0][MCP_MAX_D], Lb = s.poly_w2[1][1][MCP_MAX_D];
#pragma unroll
        for (int j = 0; j < DT; j++) {
          La = fma(sx[r][2][j], y[q][j], La);
          Lb = fma(sx[r][3][j], y[q][j], Lb);
        }
        k = fma(La, Lb, k);
      }
      kv[q] = ok[q] ? k : 0.0;
    }
    // fp64 row (ldk is a multiple of 16 doubles and c0 of 4: 32-byte aligned); columns past n2 inside the padded row are zeros
    if (c0 + 3 < ldk) {
      double* out = K + (size_t)(i0 + r) * ldk + c0;
      *reinterpret_cast<double2*>(out) = make_double2(kv[0], kv[1]);
      *reinterpret_cast<double2*>(out + 2) = make_double2(kv[2], kv[3]);
    }
    const double sc = sscale[r];
    unsigned long long G[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const long long F = isfinite(kv[q]) ? __double2ll_rn(kv[q] * sc) : 0;  // |F| < 2^(8S-2)
      G[q] = ((unsigned long long)F + bias) ^ bias;
    }
    int8_t* prow = planes + (size_t)(i0 + r) * row_bytes + (size_t)seg * S * Ksp + kq;
    for (int p = 0; p < S; p++) {
      const int sh8 = 8 * (S - 1 - p);
      unsigned packed = 0;
#pragma unroll
      for (int q = 0; q < 4; q++) packed |= ((unsigned)(G[q] >> sh8) & 255u) << (8 * q);
      *reinterpret_cast<unsigned*>(prow + (size_t)p * Ksp) = packed;
    }
  }
}

template <int DT>
__global__ void __launch_bounds__(256) kdiag_kernel(const __grid_constant__ McpGpSpec s, const double* __restrict__ X, int n,
                                                    double* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x[DT];
  KFn<DT>::load(x, X + (size_t)i * s.D, s.D);
  out[i] = KFn<DT>::kdiag(s, x);
}

// the fast reduce covers: D <= 8, an SE term, and polynomial terms in Volterra order (term p of degree p + 1), at most two
static bool fast_reduce_ok(const McpGpSpec& s) {
  if (s.D > 8 || !s.has_se || s.n_poly > 2 || (s.D > 6 && s.n_poly == 2)) return false;  // (8, 2) would spill registers
  for (int p = 0; p < s.n_poly; p++)
    if (s.poly_deg[p] != p + 1) return false;
  return true;
}


static int launch_cov(const McpGpSpec& s, const double* X1, int n1, const double* X2, int n2, int add_noise, double* K, int ldk,
                      int ncols_out, cudaStream_t st) {
  if (n1 <= 0 || ncols_out <= 0) return MCP_OK;
  if (!add_noise && fast_reduce_ok(s)) {  // cross-covariances, in particular the K* tiles of the posterior
    dim3 gridf(cdiv(ncols_out, 256), cdiv(n1, COV_ROWS));
#define MCP_FAST_COV(DT_, NP_) cov_fast_kernel<DT_, NP_><<<gridf, 256, 0, st>>>(s, X1, n1, X2, n2, K, ldk, ncols_out)
    const int np_ = s.n_poly;
    if (s.D <= 4) { if (np_ == 0) MCP_FAST_COV(4, 0); else if (np_ == 1) MCP_FAST_COV(4, 1); else MCP_FAST_COV(4, 2); }
    else if (s.D <= 6) { if (np_ == 0) MCP_FAST_COV(6, 0); else if (np_ == 1) MCP_FAST_COV(6, 1); else MCP_FAST_COV(6, 2); }
    else { if (np_ == 0) MCP_FAST_COV(8, 0); else MCP_FAST_COV(8, 1); }
#undef MCP_FAST_COV
    MCP_LAUNCH_CHECK();
    return MCP_OK;
  }
  if (!add_noise && wide_reduce_ok(s)) {  // wide inputs (UR5): same values as cov_kernel, an order of magnitude fewer loads
    dim3 gridw(cdiv(ncols_out, CW_THREADS), cdiv(n1, CW_ROWS));
    if (s.D <= 16) cov_wide_kernel<16><<<gridw, CW_THREADS, 0, st>>>(s, X1, n1, X2, n2, K, ldk, ncols_out);
    else if (s.D <= 24) cov_wide_kernel<24><<<gridw, CW_THREADS, 0, st>>>(s, X1, n1, X2, n2, K, ldk, ncols_out);
    else cov_wide_kernel<32><<<gridw, CW_THREADS, 0, st>>>(s, X1, n1, X2, n2, K, ldk, ncols_out);
    MCP_LAUNCH_CHECK();
    return MCP_OK;
  }
  dim3 grid(cdiv(ncols_out, 32), cdiv(n1, 32));
  MCP_DISPATCH_D(s.D, (cov_kernel<DT><<<grid, 256, 0, st>>>(s, X1, n1, X2, n2, add_noise, K, ldk, ncols_out)));
  MCP_LAUNCH_CHECK();
  return MCP_OK;
}

// K* rows and their digit planes in one pass (INT8 variant; kernel shapes of the fast reduce only)
void ozaki_geometry(int K, int S, int* nseg, int* Ks, int* Ksp);
static int launch_cov_slice(const McpGpSpec& s, const double* X1, int n1, const double* X2, int n2, double kdiag_max, double* K, int ldk, int S,
                            int8_t* planes, int32_t* expo, cudaStream_t st) {
  if (n1 <= 0) return MCP_OK;
  int nseg, Ks, Ksp;
  ozaki_geometry(n2, S, &nseg, &Ks, &Ksp);
  MCP_CHECK_ARG(ldk % 4 == 0 && ((uintptr_t)K % 32) == 0 && ((uintptr_t)planes % 4) == 0, "cov_slice: K* scratch must be 32-byte aligned");
  MCP_CHECK_ARG(nseg * Ks >= ldk, "cov_slice: plane columns do not cover the K* row");  // nseg Ks is N rounded up to 128 per segment >= ld16(N)
  dim3 grid(cdiv(nseg * Ksp, 1024), cdiv(n1, COV_ROWS));
#define MCP_COV_SLICE(DT_, NP_) cov_slice_kernel<DT_, NP_><<<grid, 256, 0, st>>>(s, X1, n1, X2, n2, kdiag_max, K, ldk, S, nseg, Ks, Ksp, planes, expo)
  const int np_ = s.n_poly;
  if (s.D <= 4) { if (np_ == 0) MCP_COV_SLICE(4, 0); else if (np_ == 1) MCP_COV_SLICE(4, 1); else MCP_COV_SLICE(4, 2); }
  else if (s.D <= 6) { if (np_ == 0) MCP_COV_SLICE(6, 0); else if (np_ == 1) MCP_COV_SLICE(6, 1); else MCP_COV_SLICE(6, 2); }
  else { if (np_ == 0) MCP_COV_SLICE(8, 0); else MCP_COV_SLICE(8, 1); }
#undef MCP_COV_SLICE
  MCP_LAUNCH_CHECK();
  return MCP_OK;
}

static int check_spec(const McpGpSpec* s) {
  MCP_CHECK_ARG(s != nullptr, "null gp spec");
  MCP_CHECK_ARG(s->D >= 1 && s->D <= MCP_MAX_D, "gp input dim %d outside [1,%d]", s->D, MCP_MAX_D);
  MCP_CHECK_ARG(s->n_poly >= 0 && s->n_poly <= MCP_MAX_POLY, "n_poly %d outside [0,%d]", s->n_poly, MCP_MAX_POLY);
  for (int p = 0; p < s->n_poly; p++)
    MCP_CHECK_ARG(s->poly_deg[p] >= 1 && s->poly_deg[p] <= MCP_MAX_DEG, "poly_deg[%d]=%d outside [1,%d]", p, s->poly_deg[p], MCP_MAX_DEG);
  MCP_CHECK_ARG(s->has_se || s->n_poly > 0, "empty kernel");
  return MCP_OK;
}

// ------------------------------------------------------------------------------------------------
// precompute: recursive blocked Cholesky (64 x 64 leaves) with the triangular inverse carried along
// ------------------------------------------------------------------------------------------------
constexpr int NB = 64;

// Factor one 64x64 diagonal block in place (lower), write inv(L_kk) (lower, dense 64x64) to `invL` and its transpose to `invLt`
// (both with leading dimension ldi).  Non-SPD input yields NaN (sqrt of a negative pivot) which propagates, like every
// numerical failure here.  Panel-blocked Cholesky (see below) and four threads per column of the inverse (forward substitution),
// everything in shared memory.
constexpr int POTRF_SMEM = 2 * NB * (NB + 1) * (int)sizeof(double);
__global__ void __launch_bounds__(256) potrf_block_kernel(double* __restrict__ Akk, int lda, double* __restrict__ invL,
                                                          double* __restrict__ invLt, int ldi) {
  extern __shared__ __align__(16) double potrf_smem[];
  double (*s)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(potrf_smem);
  double (*x)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(potrf_smem + NB * (NB + 1));
  const int tid = threadIdx.x, i = tid >> 2, p = tid & 3;
  {  // all 16 loads of a thread in flight before the first store (a store waiting on its load would serialise the L2 round trips)
    double v[NB * NB / 256];
#pragma unroll
    for (int u = 0; u < NB * NB / 256; u++) {
      const int e = tid + u * 256;
      v[u] = Akk[(size_t)(e / NB) * lda + (e % NB)];
    }
#pragma unroll
    for (int u = 0; u < NB * NB / 256; u++) {
      const int e = tid + u * 256;
      s[e / NB][e % NB] = v[u];
    }
  }
  __syncthreads();
  __shared__ double rdiag[NB];  // 1 / l_jj
  // Cholesky in panels of 16 columns: warp 0 factors the 16 x 16 diagonal block warp-synchronously (two lanes per row, no block
  // barrier), one thread per row solves the panel below it, all threads apply the rank-16 update to the trailing lower triangle.
  // 3 block barriers per panel (12 in all) instead of 2 per column (128): the leaf was barrier- and latency-bound.
  constexpr int PW = 16;
  const int warp = tid >> 5, lane = tid & 31;
  for (int c0 = 0; c0 < NB; c0 += PW) {
    if (warp == 0) {
      // the 16 x 16 diagonal block in registers, lane r (and its mirror r + 16) holding row r; right-looking: per column one pivot
      // broadcast, then 15 - j independent shuffle + FMA pairs — no shared-memory round trip on the column-to-column chain
      const int r = lane & (PW - 1);
      double a[PW];
#pragma unroll
      for (int k = 0; k < PW; k++) a[k] = s[c0 + r][c0 + k];
#pragma unroll
      for (int j = 0; j < PW; j++) {
        const double piv = sqrt(__shfl_sync(0xffffffffu, a[j], j));  // NaN for a non-positive pivot, which propagates
        const double rp = 1.0 / piv;
        const double lrj = (r == j) ? piv : a[j] * rp;
        a[j] = lrj;
        if (lane == j) rdiag[c0 + j] = rp;
#pragma unroll
        for (int k = j + 1; k < PW; k++) a[k] = fma(-lrj, __shfl_sync(0xffffffffu, lrj, k), a[k]);  // meaningful for r >= k
      }
      if (lane < PW) {
#pragma unroll
        for (int k = 0; k < PW; k++)
          if (k <= r) s[c0 + r][c0 + k] = a[k];
      }
    }
    __syncthreads();
    const int R = c0 + PW, nrows = NB - R;
    if (tid < nrows) {  // row R + tid of the panel:  l_rj = (a_rj - sum_{k<j} l_rk l_jk) / l_jj
      const int r = R + tid;
      double l[PW];
#pragma unroll
      for (int j = 0; j < PW; j++) {
        double a = s[r][c0 + j];
#pragma unroll
        for (int k = 0; k < j; k++) a = fma(-l[k], s[c0 + j][c0 + k], a);
        l[j] = a * rdiag[c0 + j];
      }
#pragma unroll
      for (int j = 0; j < PW; j++) s[r][c0 + j] = l[j];
    }
    __syncthreads();
    for (int ii = tid >> 4; ii < nrows; ii += 16) {  // trailing lower triangle:  a_ik -= sum_j l_ij l_kj
      for (int kk = tid & 15; kk <= ii; kk += 16) {
        double a = s[R + ii][R + kk];
#pragma unroll
        for (int j = 0; j < PW; j++) a = fma(-s[R + ii][c0 + j], s[R + kk][c0 + j], a);
        s[R + ii][R + kk] = a;
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < NB * NB; e += 256) {
    int r = e / NB, k = e % NB;
    Akk[(size_t)r * lda + k] = (k <= r) ? s[r][k] : 0.0;
  }
  // inverse X = L^-1 in 16 x 16 blocks.  Diagonal blocks first, all four at once (warp b, lane c: column c of X_bb by forward
  // substitution in registers, x_r = (delta_rc - sum_{c<=k<r} l_rk x_k) / l_rr; the l_rk are broadcast reads).  Then block row by block
  // row  X_ij = -X_ii (sum_{k=j}^{i-1} L_ik X_kj): two small products with all threads, no serial chain inside.
  for (int e = tid; e < NB * (NB + 1); e += 256) (&x[0][0])[e] = 0.0;
  __syncthreads();
  if (warp < NB / PW && lane < PW) {
    const int b0 = warp * PW, c = lane;
    double xc[PW];
#pragma unroll
    for (int r = 0; r < PW; r++) {
      double acc = (r == c) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < r; k++) acc = fma(-s[b0 + r][b0 + k], (k >= c) ? xc[k] : 0.0, acc);
      xc[r] = (r < c) ? 0.0 : acc * rdiag[b0 + r];
    }
#pragma unroll
    for (int r = 0; r < PW; r++) x[b0 + r][b0 + c] = xc[r];
  }
  __syncthreads();
  double (*tmp)[NB + 1] = s;  // T_ij goes where the (already exported) strict upper triangle of s is: rows j-block, cols i-block — see below
  for (int bi = 1; bi < NB / PW; bi++) {
    const int i0 = bi * PW;
    // T[r][c] = sum_{k = j0}^{i0 - 1} L[i0 + r][k] X[k][c]   for every column c < i0 (its block column j0 = c / 16 * 16); X[k][c] = 0 for k < c
    for (int e = tid; e < PW * i0; e += 256) {
      const int r = e / i0, c = e - r * i0;
      double acc = 0.0;
      for (int k = c; k < i0; k++) acc = fma(s[i0 + r][k], x[k][c], acc);
      tmp[c][i0 + r] = acc;  // stored transposed into the upper triangle of s (c < i0 <= i0 + r): free space, no clash with L
    }
    __syncthreads();
    // X[i0 + r][c] = -sum_{q <= r} X_ii[r][q] T[q][c]
    for (int e = tid; e < PW * i0; e += 256) {
      const int r = e / i0, c = e - r * i0;
      double acc = 0.0;
#pragma unroll
      for (int q = 0; q < PW; q++) acc = fma(x[i0 + r][i0 + q], (q <= r) ? tmp[c][i0 + q] : 0.0, acc);
      x[i0 + r][c] = -acc;
    }
    __syncthreads();
  }
  __syncthreads();
  for (int e = tid; e < NB * NB; e += 256) {
    int r = e / NB, k = e % NB;
    invL[(size_t)r * ldi + k] = x[r][k];
    invLt[(size_t)r * ldi + k] = x[k][r];
  }
}

// dst[j][i] = src[i][j] for a rows x cols block (32 x 32 tiles through shared memory)
__global__ void __launch_bounds__(256) transpose_kernel(const double* __restrict__ src, int lds, double* __restrict__ dst, int ldd,
                                                        int rows, int cols) {
  __shared__ double t[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8)
    if (r0 + r < rows && c0 + tx < cols) t[r][tx] = src[(size_t)(r0 + r) * lds + c0 + tx];
  __syncthreads();
  for (int c = ty; c < 32; c += 8)
    if (c0 + c < cols && r0 + tx < rows) dst[(size_t)(c0 + c) * ldd + r0 + tx] = t[tx][c];
}

__global__ void pad_identity_kernel(double* K, int ld, int n, int np) {
  int i = n + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < np) K[(size_t)i * ld + i] = 1.0;
}

// dst = lower triangle of src, zeros above
__global__ void lower_copy_kernel(const double* __restrict__ src, int lds, double* __restrict__ dst, int ldd, int n) {
  int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
  if (i < n && j < n) dst[(size_t)i * ldd + j] = (j <= i) ? src[(size_t)i * lds + j] : 0.0;
}

// copy the lower triangle of src (computed) into a full symmetric dst [n x n]
__global__ void symmetrize_kernel(const double* __restrict__ src, int lds, double* __restrict__ dst, int ldd, int n) {
  int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
  if (i < n && j < n) dst[(size_t)i * ldd + j] = (j <= i) ? src[(size_t)i * lds + j] : src[(size_t)j * lds + i];
}

__global__ void copy2d_kernel(const double* __restrict__ src, int lds, double* __restrict__ dst, int ldd, int rows, int cols) {
  int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
  if (i < rows && j < cols) dst[(size_t)i * ldd + j] = src[(size_t)i * lds + j];
}

// alpha = Kinv (y - mean0); one warp per row.  GP_prior.py:133
__global__ void alpha_kernel(const double* __restrict__ Kinv, int ld, const double* __restrict__ y, double mean0, int n,
                             double* __restrict__ alpha) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  double a = 0.0;
  for (int k = lane; k < n; k += 32) a = fma(Kinv[(size_t)row * ld + k], y[k] - mean0, a);
  a = warp_sum(a);
  if (lane == 0) alpha[row] = a;
}

}  // namespace mcp

using namespace mcp;

extern "C" __attribute__((visibility("default"))) size_t mcpilco_gp_precompute_workspace_bytes(int N) {
  size_t np = align_up((size_t)(N > 0 ? N : 1), NB);
  return 3 * np * np * sizeof(double) + 256;
}

extern "C" __attribute__((visibility("default"))) int mcpilco_gp_covariance(const McpGpSpec* spec, const double* X1, int n1, const double* X2, int n2, int add_noise,
                                     double* K, int ldk, void* stream) {
  if (int e = check_spec(spec)) return e;
  MCP_CHECK_ARG(X1 && K && n1 >= 0, "gp_covariance: null pointer or negative size");
  if (!X2) { X2 = X1; n2 = n1; } else add_noise = 0;
  MCP_CHECK_ARG(ldk >= n2, "gp_covariance: ldk %d < n2 %d", ldk, n2);
  return launch_cov(*spec, X1, n1, X2, n2, add_noise, K, ldk, n2, (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int mcpilco_gp_diag_covariance(const McpGpSpec* spec, const double* X, int n, double* diag, void* stream) {
  if (int e = check_spec(spec)) return e;
  if (n <= 0) return MCP_OK;
  MCP_CHECK_ARG(X && diag, "gp_diag_covariance: null pointer");
  MCP_DISPATCH_D(spec->D, (kdiag_kernel<DT><<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(*spec, X, n, diag)));
  MCP_LAUNCH_CHECK();
  return MCP_OK;
}

extern "C" __attribute__((visibility("default"))) int mcpilco_gp_precompute(const McpGpSpec* spec, const double* Xtr, const double* y, int N, double* alpha, double* Kinv,
                                     int ld, double* Lfac, double* Linv, void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_spec(spec)) return e;
  MCP_CHECK_ARG(N >= 1 && Xtr && y && alpha && Kinv && ld >= N, "gp_precompute: bad arguments (N=%d ld=%d)", N, ld);
  MCP_CHECK_ARG(workspace && workspace_bytes >= mcpilco_gp_precompute_workspace_bytes(N), "gp_precompute: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int np = (int)align_up((size_t)N, NB), nblk = np / NB;
  double* Kp = (double*)align_up((size_t)workspace, 256);
  double* I = Kp + (size_t)np * np;  // L^-1 (lower)
  double* W = I + (size_t)np * np;   // L^-T (upper) = I^T

  // 1. K = k(X,X) + sigma_n2 I, padded to a multiple of the block size with an identity tail
  MCP_CUDA(cudaMemsetAsync(Kp, 0, sizeof(double) * 3 * (size_t)np * np, st));
  if (int e = launch_cov(*spec, Xtr, N, Xtr, N, 1, Kp, np, N, st)) return e;
  if (np > N) { pad_identity_kernel<<<cdiv(np - N, 128), 128, 0, st>>>(Kp, np, N, np); MCP_LAUNCH_CHECK(); }

  // 2 + 3. recursive Cholesky with the triangular inverse carried along, so that every update is a large NT product:
  //   [A11 .; A21 A22]:  (L11, I11) <- node(A11);  L21 = A21 I11^T;  A22 -= L21 L21^T;  (L22, I22) <- node(A22);
  //   I21 = -I22 (L21 I11),  formed as  T^T = I11^T L21^T  (scratch: the unused upper-right block of Kp)  and  I21 = -I22 (T^T)^T.
  // Leaves are 64 x 64 blocks factored and inverted by one CTA.  L overwrites the lower triangle of Kp.
  struct Rec {
    double *Kp, *I, *W; int np; cudaStream_t st;
    static int gemm(int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double beta, double* C, int ldc,
                    int tri, int kflags, cudaStream_t st) {
      if ((size_t)cdiv(M, 128) * cdiv(N, 128) >= 120 && dgemm_tma_usable(A, lda, B, ldb, C, ldc))
        return dgemm_nt_tma_trim(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, tri, kflags, st);
      return dgemm_nt(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, tri, kflags, st);
    }
    int node(int b0, int nb) const {
      const size_t o = (size_t)b0 * NB * np + (size_t)b0 * NB;   // offset of the node's diagonal origin
      if (nb == 1) {
        potrf_block_kernel<<<1, 256, POTRF_SMEM, st>>>(Kp + o, np, I + o, W + o, np);
        MCP_LAUNCH_CHECK();
        return MCP_OK;
      }
      const int nb1 = nb / 2, nb2 = nb - nb1, h1 = nb1 * NB, h2 = nb2 * NB;
      if (int e = node(b0, nb1)) return e;
      double* A21 = Kp + o + (size_t)h1 * np;      // [h2 x h1]
      double* A22 = A21 + h1;                      // [h2 x h2]
      double* U12 = Kp + o + h1;                   // [h1 x h2] scratch (stale upper triangle)
      double* I21 = I + o + (size_t)h1 * np;       // [h2 x h1]
      if (int e = gemm(h2, h1, h1, 1.0, A21, np, I + o, np, 0.0, I21, np, 0, KF_B_LOWER, st)) return e;
      copy2d_kernel<<<dim3(cdiv(h1, 32), cdiv(h2, 8)), dim3(32, 8), 0, st>>>(I21, np, A21, np, h2, h1);
      MCP_LAUNCH_CHECK();
      if (int e = gemm(h2, h2, h1, -1.0, A21, np, A21, np, 1.0, A22, np, 1, 0, st)) return e;
      if (int e = node(b0 + nb1, nb2)) return e;
      if (int e = gemm(h1, h2, h1, 1.0, W + o, np, A21, np, 0.0, U12, np, 0, KF_A_UPPER, st)) return e;
      if (int e = gemm(h2, h1, h2, -1.0, I + o + (size_t)h1 * np + h1, np, U12, np, 0.0, I21, np, 0, KF_A_LOWER, st)) return e;
      transpose_kernel<<<dim3(cdiv(h1, 32), cdiv(h2, 32)), 256, 0, st>>>(I21, np, W + o + h1, np, h2, h1);
      MCP_LAUNCH_CHECK();
      return MCP_OK;
    }
  };
  static bool potrf_configured[MCP_MAX_DEVICES] = {};
  MCP_CUDA(ensure_dynamic_smem(potrf_configured, potrf_block_kernel, POTRF_SMEM));
  if (int e = Rec{Kp, I, W, np, st}.node(0, nblk)) return e;
  if (Lfac) {  // export L (lower; the upper part of Kp holds scratch, mask it)
    lower_copy_kernel<<<dim3(cdiv(N, 32), cdiv(N, 8)), dim3(32, 8), 0, st>>>(Kp, np, Lfac, ld, N);
    MCP_LAUNCH_CHECK();
  }
  if (Linv) {  // export R = L^-1 (lower): what forward-only posteriors contract with instead of the full Kinv
    lower_copy_kernel<<<dim3(cdiv(N, 32), cdiv(N, 8)), dim3(32, 8), 0, st>>>(I, np, Linv, ld, N);
    MCP_LAUNCH_CHECK();
    if (ld > N) MCP_CUDA(cudaMemset2DAsync(Linv + N, sizeof(double) * ld, 0, sizeof(double) * (ld - N), N, st));
  }

  // 4. K^-1 = W W^T (lower tiles, contraction trimmed to k >= max(row blocks)), mirrored into the output
  if (int e = Rec::gemm(np, np, np, 1.0, W, np, W, np, 0.0, Kp, np, 1, KF_A_UPPER | KF_B_UPPER, st)) return e;
  symmetrize_kernel<<<dim3(cdiv(N, 32), cdiv(N, 8)), dim3(32, 8), 0, st>>>(Kp, np, Kinv, ld, N);
  MCP_LAUNCH_CHECK();
  if (ld > N) {  // keep the row padding finite (zero) for the tile loaders
    MCP_CUDA(cudaMemset2DAsync(Kinv + N, sizeof(double) * ld, 0, sizeof(double) * (ld - N), N, st));
  }
  // 5. alpha = K^-1 (y - m)
  alpha_kernel<<<cdiv(N, 8), 256, 0, st>>>(Kinv, ld, y, spec->mean0, N, alpha);
  MCP_LAUNCH_CHECK();
  return MCP_OK;
}

// ------------------------------------------------------------------------------------------------
// posterior mean / variance and their Jacobians w.r.t. the test input (a3 + the per-step part of BPTT)
// ------------------------------------------------------------------------------------------------
namespace mcp {

// One warp per particle.  Given V = K* K^-1 (row m), recompute k(x_m, xtr_n) and dk/dx on the fly and reduce
//   mean = mean0 + sum_n alpha_n k_n            q = sum_n V_n k_n            var = scale (k** - q)
//   dmean/dx = sum_n alpha_n dk_n/dx            dvar/dx = scale (dk**/dx - 2 sum_n V_n dk_n/dx)
// (d(k^T Kinv k)/dx = 2 (Kinv k)^T dk/dx since Kinv is symmetric.)
template <int DT, bool JAC>
__global__ void __launch_bounds__(256) posterior_reduce_kernel(const __grid_constant__ McpGpSpec s, const double* __restrict__ Xs,
                                                               int M, const double* __restrict__ Xtr, const double* __restrict__ alpha,
                                                               int N, const double* __restrict__ V, int ldv, double var_scale,
                                                               int E, int e, double* __restrict__ mean, double* __restrict__ var,
                                                               double* __restrict__ jmean, double* __restrict__ jvar) {
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= M) return;
  double x[DT];
  KFn<DT>::load(x, Xs + (size_t)m * s.D, s.D);
  double mu = 0.0, q = 0.0, gm[DT], gq[DT];
#pragma unroll
  for (int j = 0; j < DT; j++) gm[j] = gq[j] = 0.0;
  const double* v = V + (size_t)m * ldv;
  for (int n = lane; n < N; n += 32) {
    double y[DT];
    KFn<DT>::load(y, Xtr + (size_t)n * s.D, s.D);
    double a = alpha[n], vn = v[n];
    if (JAC) {
      double kv, dk[DT];
      KFn<DT>::k_grad(s, x, y, kv, dk);
      mu = fma(a, kv, mu);
      q = fma(vn, kv, q);
#pragma unroll
      for (int j = 0; j < DT; j++) {
        gm[j] = fma(a, dk[j], gm[j]);
        gq[j] = fma(vn, dk[j], gq[j]);
      }
    } else {
      double kv = KFn<DT>::k(s, x, y);
      mu = fma(a, kv, mu);
      q = fma(vn, kv, q);
    }
  }
  mu = warp_sum(mu);
  q = warp_sum(q);
  if (JAC) {
#pragma unroll
    for (int j = 0; j < DT; j++) {
      gm[j] = warp_sum(gm[j]);
      gq[j] = warp_sum(gq[j]);
    }
  }
  if (lane == 0) {
    double kd, dkd[DT];
    KFn<DT>::kdiag_grad(s, x, kd, dkd);
    mean[(size_t)m * E + e] = s.mean0 + mu;
    var[(size_t)m * E + e] = var_scale * (kd - q);
    if (JAC) {
#pragma unroll
      for (int j = 0; j < DT; j++) {
        if (j < s.D) {
          jmean[((size_t)m * E + e) * s.D + j] = gm[j];
          jvar[((size_t)m * E + e) * s.D + j] = var_scale * (dkd[j] - 2.0 * gq[j]);
        }
      }
    }
  }
}


// Forward-only posterior from the triangular factor: row m of Ks = k(x_m, Xtr) and of W = Ks L^-T given,
//   mean = mean0 + sum_n alpha_n Ks_n,   var = scale (k** - sum_n W_n^2)      (k*^T Kinv k* = |L^-1 k*|^2).
// One warp per particle; streams the two rows once (HBM / L2 bound: 16 N bytes per particle).
template <int DT>
__global__ void __launch_bounds__(256) posterior_tri_reduce_kernel(const __grid_constant__ McpGpSpec s, const double* __restrict__ Xs, int M,
                                                                   const double* __restrict__ alpha, int N, const double* __restrict__ Ks,
                                                                   const double* __restrict__ W, int ldk, double var_scale, int E, int e,
                                                                   double* __restrict__ mean, double* __restrict__ var) {
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= M) return;
  const double* k = Ks + (size_t)m * ldk;
  const double* w = W + (size_t)m * ldk;
  double mu = 0.0, q = 0.0;
  int n = lane * 2;
  for (; n + 1 < N; n += 64) {  // ldk is a multiple of 16 and the scratch base is 256-byte aligned: 16-byte row loads
    const double2 kk = *reinterpret_cast<const double2*>(k + n), ww = *reinterpret_cast<const double2*>(w + n);
    const double2 aa = make_double2(alpha[n], alpha[n + 1]);
    mu = fma(aa.x, kk.x, mu);
    mu = fma(aa.y, kk.y, mu);
    q = fma(ww.x, ww.x, q);
    q = fma(ww.y, ww.y, q);
  }
  if (n < N) {
    mu = fma(alpha[n], k[n], mu);
    q = fma(w[n], w[n], q);
  }
  mu = warp_sum(mu);
  q = warp_sum(q);
  if (lane == 0) {
    double x[DT];
    KFn<DT>::load(x, Xs + (size_t)m * s.D, s.D);
    mean[(size_t)m * E + e] = s.mean0 + mu;
    var[(size_t)m * E + e] = var_scale * (KFn<DT>::kdiag(s, x) - q);
  }
}


// Restructured reduce for the common kernel shapes (SE + Volterra polynomial of degree <= 2, D <= 8): the gradient sums are
// factored so that the particle-dependent coefficients leave the n-loop,
//   sum_n a_n dk_n/dx_j = -2 ils_j^2 (x_j E0 - E1_j) + sum_{p,f} w_pfj C_pfj,
//   E0 = sum_n a_n e_n,  E1_j = sum_n a_n e_n y_nj,  C_pfj = sum_n a_n c_pf,n y_nj,  c_pf,n = prod_{g != f} L_pg,n,
// for the two weight channels a = alpha (mean) and a = V (variance).  The kernel VALUE k_n is not evaluated again: the K* row the
// contraction consumed is streamed in beside the V row, the (cheap) polynomial factors are rebuilt from the particle-scaled weights
// and the squared-exponential part is e_n = k_n - poly_n — no distance, no exp.  Per (particle, training point) that leaves ~4D + 5
// FMAs per channel plus ~3D for the polynomial factors.
// NP = number of polynomial terms; term p has degree p + 1 (get_Volterra_MPK_GP, Sparse_GP.py:671-737).
constexpr int RED_TILE = 256;     // training points staged per tile
constexpr int RED_THREADS = 128;  // 4 particles per block; two blocks share an SM, so one block's barriers and cp.async waits overlap
                                  // with the other's FP64 work
constexpr int RED_PB = RED_THREADS / 32;
// row stride of the transposed training-input tile: staging writes (thread <-> (point, dimension), dimension fastest) then spread over
// the shared-memory banks instead of piling D-fold on one (16.9 M bank conflicts per launch with a stride of 256)
template <int DT>
constexpr int red_yld() { return RED_TILE + (DT <= 4 ? 4 : DT <= 6 ? 3 : 2); }
template <int DT>
constexpr size_t red_smem_bytes() {
  return sizeof(double) * (size_t)(2 * 2 * RED_PB * RED_TILE + 2 * RED_TILE + 2 * DT * red_yld<DT>());
}
template <int DT, int NP>
__global__ void __launch_bounds__(RED_THREADS, 2) posterior_reduce_fast_kernel(const __grid_constant__ McpGpSpec s, const double* __restrict__ Xs,
                                                                    int M, const double* __restrict__ Xtr,
                                                                    const double* __restrict__ alpha, int N,
                                                                    const double* __restrict__ Ks, const double* __restrict__ V, int ldv,
                                                                    double var_scale, int E, int e, double* __restrict__ mean,
                                                                    double* __restrict__ var, double* __restrict__ jmean,
                                                                    double* __restrict__ jvar) {
  // double-buffered by cp.async: per warp (= particle) its K* and V rows; shared by the block: alpha and the training inputs
  // (transposed, sY[buf][j][i]: conflict-free for lane <-> point)
  extern __shared__ __align__(16) double red_smem[];
  constexpr int YLD = red_yld<DT>();
  double(*sK)[RED_PB][RED_TILE] = reinterpret_cast<double(*)[RED_PB][RED_TILE]>(red_smem);
  double(*sV)[RED_PB][RED_TILE] = reinterpret_cast<double(*)[RED_PB][RED_TILE]>(red_smem + 2 * RED_PB * RED_TILE);
  double(*sA)[RED_TILE] = reinterpret_cast<double(*)[RED_TILE]>(red_smem + 4 * RED_PB * RED_TILE);
  double(*sY)[DT][YLD] = reinterpret_cast<double(*)[DT][YLD]>(red_smem + 4 * RED_PB * RED_TILE + 2 * RED_TILE);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int m = min(blockIdx.x * RED_PB + warp, M - 1);  // surplus warps shadow the last particle (they still help staging)
  const bool owner = blockIdx.x * RED_PB + warp < M;
  const int D = s.D;
  const double* krow = Ks + (size_t)m * ldv;
  const double* vrow = V + (size_t)m * ldv;
  auto stage = [&](int t, int buf) {
    const int n0 = t * RED_TILE;
#pragma unroll
    for (int c = 0; c < RED_TILE / RED_THREADS; c++) {  // thread <-> training point: its D inputs into the D transposed rows, no index division
      const int i = tid + c * RED_THREADS;
      const bool ok = n0 + i < N;
      const double* src = Xtr + (size_t)(ok ? n0 + i : 0) * D;
#pragma unroll
      for (int j = 0; j < DT; j++)
        if (j < D) cp_async8(&sY[buf][j][i], src + j, ok ? 8 : 0);
      cp_async8(&sA[buf][i], alpha + (ok ? n0 + i : 0), ok ? 8 : 0);
    }
#pragma unroll
    for (int c = 0; c < RED_TILE / 64; c++) {  // rows are 16-byte aligned (ld a multiple of 16 doubles); entries past N arrive as zeros
      const int i = 2 * (lane + 32 * c), left = N - (n0 + i);
      const int bytes = left >= 2 ? 16 : (left == 1 ? 8 : 0);
      const int src = left > 0 ? n0 + i : 0;
      cp_async16(&sK[buf][warp][i], krow + src, bytes);
      cp_async16(&sV[buf][warp][i], vrow + src, bytes);
    }
    cp_async_commit();
  };
  double x[DT];
  KFn<DT>::load(x, Xs + (size_t)m * D, D);
  // particle-scaled polynomial weights: L = o + sum_j (w_j x_j) y_j
  double xw1[DT], xw2a[DT], xw2b[DT];
#pragma unroll
  for (int j = 0; j < DT; j++) {
    xw1[j] = NP >= 1 ? s.poly_w2[0][0][j] * x[j] : 0.0;
    xw2a[j] = NP >= 2 ? s.poly_w2[1][0][j] * x[j] : 0.0;
    xw2b[j] = NP >= 2 ? s.poly_w2[1][1][j] * x[j] : 0.0;
  }
  double mu = 0.0, q = 0.0, E0a = 0.0, E0v = 0.0;
  double E1a[DT], E1v[DT], C1a[DT], C1v[DT], C2a0[DT], C2v0[DT], C2a1[DT], C2v1[DT];
#pragma unroll
  for (int j = 0; j < DT; j++) E1a[j] = E1v[j] = C1a[j] = C1v[j] = C2a0[j] = C2v0[j] = C2a1[j] = C2v1[j] = 0.0;
  const int T = (N + RED_TILE - 1) / RED_TILE;
  constexpr int PER = RED_TILE / 32;
  stage(0, 0);
  for (int t = 0; t < T; t++) {
    const int buf = t & 1;
    if (t + 1 < T) {
      stage(t + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < PER; it++) {
      const int i = it * 32 + lane;
      double y[DT];
#pragma unroll
      for (int j = 0; j < DT; j++) y[j] = (j < D) ? sY[buf][j][i] : 0.0;
      const double a = sA[buf][i], vn = sV[buf][warp][i], kv = sK[buf][warp][i];  // all zero past N: padded points contribute nothing
      double poly = 0.0, L2a = 0.0, L2b = 0.0;
      if (NP >= 1) {
        double L1 = s.poly_w2[0][0][MCP_MAX_D];
#pragma unroll
        for (int j = 0; j < DT; j++) L1 = fma(xw1[j], y[j], L1);
        poly = L1;
      }
      if (NP >= 2) {
        L2a = s.poly_w2[1][0][MCP_MAX_D];
        L2b = s.poly_w2[1][1][MCP_MAX_D];
#pragma unroll
        for (int j = 0; j < DT; j++) {
          L2a = fma(xw2a[j], y[j], L2a);
          L2b = fma(xw2b[j], y[j], L2b);
        }
        poly = fma(L2a, L2b, poly);
      }
      const double ev = kv - poly;  // the squared-exponential part of k_n (exactly k_n when there is no polynomial term)
      mu = fma(a, kv, mu);
      q = fma(vn, kv, q);
      const double ta = a * ev, tv = vn * ev;
      E0a += ta;
      E0v += tv;
      const double ua0 = a * L2b, uv0 = vn * L2b, ua1 = a * L2a, uv1 = vn * L2a;
#pragma unroll
      for (int j = 0; j < DT; j++) {
        E1a[j] = fma(ta, y[j], E1a[j]);
        E1v[j] = fma(tv, y[j], E1v[j]);
        if (NP >= 1) {
          C1a[j] = fma(a, y[j], C1a[j]);
          C1v[j] = fma(vn, y[j], C1v[j]);
        }
        if (NP >= 2) {
          C2a0[j] = fma(ua0, y[j], C2a0[j]);
          C2v0[j] = fma(uv0, y[j], C2v0[j]);
          C2a1[j] = fma(ua1, y[j], C2a1[j]);
          C2v1[j] = fma(uv1, y[j], C2v1[j]);
        }
      }
    }
    __syncthreads();  // the buffer read here is refilled by the next iteration's stage()
  }
  mu = warp_sum(mu);
  q = warp_sum(q);
  E0a = warp_sum(E0a);
  E0v = warp_sum(E0v);
  double gm[DT], gq[DT];
#pragma unroll
  for (int j = 0; j < DT; j++) {
    const double il2 = -2.0 * s.inv_ls[j] * s.inv_ls[j];
    double ga = il2 * (x[j] * E0a - warp_sum(E1a[j])), gv = il2 * (x[j] * E0v - warp_sum(E1v[j]));
    if (NP >= 1) {
      ga = fma(s.poly_w2[0][0][j], warp_sum(C1a[j]), ga);
      gv = fma(s.poly_w2[0][0][j], warp_sum(C1v[j]), gv);
    }
    if (NP >= 2) {
      ga = fma(s.poly_w2[1][0][j], warp_sum(C2a0[j]), ga);
      gv = fma(s.poly_w2[1][0][j], warp_sum(C2v0[j]), gv);
      ga = fma(s.poly_w2[1][1][j], warp_sum(C2a1[j]), ga);
      gv = fma(s.poly_w2[1][1][j], warp_sum(C2v1[j]), gv);
    }
    gm[j] = ga;
    gq[j] = gv;
  }
  if (lane == 0 && owner) {
    double kd, dkd[DT];
    KFn<DT>::kdiag_grad(s, x, kd, dkd);
    mean[(size_t)m * E + e] = s.mean0 + mu;
    var[(size_t)m * E + e] = var_scale * (kd - q);
#pragma unroll
    for (int j = 0; j < DT; j++) {
      if (j < D) {
        jmean[((size_t)m * E + e) * D + j] = gm[j];
        jvar[((size_t)m * E + e) * D + j] = var_scale * (dkd[j] - 2.0 * gq[j]);
      }
    }
  }
}

// Reduce for WIDE gp inputs (8 < D <= 32; the UR5 model has D = 24) with an SE term and at most one linear term.  Everything is small
// here (UR5: 200 particles x 6 outputs x 400 training points), so the kernel is built to be short rather than wide:
//   * the kernel VALUE is not evaluated again: the K* row the contraction consumed is read beside the V row, the linear term
//     L1[p][n] = o + sum_j (w_j x_pj) y_nj of the CTA's 8 particles comes from one small DMMA product per tile, and the
//     squared-exponential part is e_n = k_n - L1_n (no distance, no exp);
//   * the gradient sums are ONE 8 x 8 x 4 DMMA chain per warp and feature tile: rows = the four weight channels (a e, v e, a, v) of the
//     warp's two particles, columns = features [y_0 .. y_{D-1} | 1, k_A, k_B], contraction over the tile's 64 training points:
//       sum_n a_n dk_n/dx_j = -2 ils_j^2 (x_j E0 - E1_j) + w1_j C1_j,  E1_j = [a e] x y_j,  E0 = [a e] x 1,  C1_j = [a] x y_j,
//     and mean / q fall out of the same product as [a] x k, [v] x k;
//   * the training points are split over the CTAs of a thread-block cluster (gridDim.z = cluster size = 1, 2 or 4, chosen from N alone):
//     each CTA reduces every nseg-th tile, the partial sums meet in rank 0 through distributed shared memory in a fixed order
//     (bit-stable), rank 0 finalises.  UR5: 25 x 6 x 4 = 600 CTAs of 128 threads, one wave.
constexpr int WIDE_TILE = 64;
constexpr int WIDE_WPC = 4;              // warps per CTA
constexpr int WIDE_PPC = 2 * WIDE_WPC;   // particles per CTA (two per warp: the two halves of the DMMA's eight rows)
constexpr int WIDE_LDT = WIDE_TILE + 4;  // row stride = 4 (mod 16) doubles: conflict-free DMMA fragment loads, lane <-> point reads stride 1
constexpr int WIDE_LDX = MCP_MAX_D + 4;
constexpr int WIDE_LDG = MCP_MAX_D + 8;  // feature columns of the result tile: y tiles, then the special tile [1, k_A, k_B, 0 ...]
static inline int wide_segments(int N) {
  const int tiles = cdiv(N, WIDE_TILE);
  return tiles >= 4 ? 4 : (tiles >= 2 ? 2 : 1);
}
static inline size_t wide_smem_bytes(int D) {
  const int Dp8 = (D + 7) & ~7;
  return sizeof(double) * ((size_t)WIDE_WPC * 10 * WIDE_LDT + WIDE_PPC * WIDE_LDT + WIDE_PPC * WIDE_LDX + WIDE_PPC * MCP_MAX_D + WIDE_TILE +
                           MCP_MAX_D + (size_t)Dp8 * WIDE_LDT);
}
__device__ __forceinline__ void reduce_wide_body(const McpGpSpec& s, const double* __restrict__ Xs, int M, const double* __restrict__ Xtr,
                                                 const double* __restrict__ alpha, int N, const double* __restrict__ Ks,
                                                 const double* __restrict__ V, int ldv, double var_scale, int E, int e,
                                                 double* __restrict__ mean, double* __restrict__ var, double* __restrict__ jmean,
                                                 double* __restrict__ jvar) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int nseg = (int)cluster.num_blocks(), seg = (int)cluster.block_rank();
  // dynamic shared memory (wide_smem_bytes(D): 44 KB at D = 24, so that five CTAs — all 600 of the UR5 step — share an SM)
  extern __shared__ __align__(16) double wide_smem[];
  const int Dp8 = (s.D + 7) & ~7;
  double(*sW)[10][WIDE_LDT] = reinterpret_cast<double(*)[10][WIDE_LDT]>(wide_smem);  // per warp: 8 weight rows (4 channels x 2 particles), the two K* rows
  double(*sL1)[WIDE_LDT] = reinterpret_cast<double(*)[WIDE_LDT]>(wide_smem + WIDE_WPC * 10 * WIDE_LDT);  // linear kernel term [particle][point]
  double(*sXw)[WIDE_LDX] = reinterpret_cast<double(*)[WIDE_LDX]>(&sL1[WIDE_PPC][0]);                     // w1_j x_pj
  double(*sX)[MCP_MAX_D] = reinterpret_cast<double(*)[MCP_MAX_D]>(&sXw[WIDE_PPC][0]);
  double* sA = &sX[WIDE_PPC][0];
  double* sIl = sA + WIDE_TILE;
  double(*sYt)[WIDE_LDT] = reinterpret_cast<double(*)[WIDE_LDT]>(sIl + MCP_MAX_D);  // [Dp8][LDT] training inputs of the tile, transposed
  static_assert(8 * WIDE_LDG <= 10 * WIDE_LDT, "the result tile reuses the warp's weight rows");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x, gq = lane >> 2, q = lane & 3;
  const int D = s.D, np1 = s.n_poly;  // np1 in {0, 1}
  const double off = np1 ? s.poly_w2[0][0][MCP_MAX_D] : 0.0;
  const bool se = s.has_se != 0;
  const int nfy = (D + 7) >> 3;       // feature tiles of training inputs; tile nfy is the special one
  const int p0 = blockIdx.x * WIDE_PPC;
  {  // all loads first, then the stores: a store that waits for its load would serialise the L2 round trips (in-order issue)
    constexpr int NI = (WIDE_PPC * WIDE_LDX + WIDE_WPC * 32 - 1) / (WIDE_WPC * 32);
    double xv[NI];
#pragma unroll
    for (int u = 0; u < NI; u++) {
      const int el = tid + u * WIDE_WPC * 32, p = el / WIDE_LDX, j = el - p * WIDE_LDX;
      xv[u] = (el < WIDE_PPC * WIDE_LDX && j < D && p0 + p < M) ? Xs[(size_t)(p0 + p) * D + j] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < NI; u++) {
      const int el = tid + u * WIDE_WPC * 32, p = el / WIDE_LDX, j = el - p * WIDE_LDX;
      if (el < WIDE_PPC * WIDE_LDX) {
        sXw[p][j] = (np1 && j < D) ? s.poly_w2[0][0][j] * xv[u] : 0.0;
        if (j < MCP_MAX_D) sX[p][j] = xv[u];
      }
    }
  }
  if (tid < MCP_MAX_D) sIl[tid] = tid < D ? s.inv_ls[tid] : 0.0;
  for (int el = tid; el < (Dp8 - D) * WIDE_LDT; el += WIDE_WPC * 32) sYt[D + el / WIDE_LDT][el % WIDE_LDT] = 0.0;  // rows past D stay zero
  double acc[5][2];
#pragma unroll
  for (int f = 0; f < 5; f++) acc[f][0] = acc[f][1] = 0.0;
  const int mA = min(p0 + 2 * warp, M - 1), mB = min(p0 + 2 * warp + 1, M - 1);  // surplus particles shadow the last one
  const double *kA = Ks + (size_t)mA * ldv, *kB = Ks + (size_t)mB * ldv, *vA = V + (size_t)mA * ldv, *vB = V + (size_t)mB * ldv;
  for (int n0 = seg * WIDE_TILE; n0 < N; n0 += nseg * WIDE_TILE) {
    __syncthreads();  // previous tile fully consumed (first pass: sXw / sX / sIl and the zero rows visible)
    const int cnt = min(WIDE_TILE, N - n0);
    for (int el = tid; el < WIDE_TILE * D; el += WIDE_WPC * 32) {  // asynchronous, transposing copy: no register round trip
      const int i = el / D, j = el - i * D;
      cp_async8(&sYt[j][i], Xtr + (i < cnt ? (size_t)n0 * D + el : 0), i < cnt ? 8 : 0);
    }
    if (tid < WIDE_TILE) cp_async8(&sA[tid], alpha + (tid < cnt ? n0 + tid : 0), tid < cnt ? 8 : 0);
    cp_async_commit();
    // this warp's share of the weights' inputs: K* and V of its two particles at the tile's points (two points per lane)
    double kk[2][2], vv[2][2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int n = n0 + h * 32 + lane;
      const bool ok = n < N;
      kk[0][h] = ok ? kA[n] : 0.0;
      kk[1][h] = ok ? kB[n] : 0.0;
      vv[0][h] = ok ? vA[n] : 0.0;
      vv[1][h] = ok ? vB[n] : 0.0;
    }
    cp_async_wait<0>();
    __syncthreads();
    if (np1) {  // L1 of all 8 particles at this warp's 16 points: C[p][n] = sum_j xw[p][j] y[n][j]
#pragma unroll
      for (int nb = 0; nb < 2; nb++) {
        const int c0 = (2 * warp + nb) * 8;
        double l0 = 0.0, l1 = 0.0;
        for (int k4 = 0; k4 < D; k4 += 4) dmma884(l0, l1, sXw[gq][k4 + q], sYt[k4 + q][c0 + gq]);
        sL1[gq][c0 + 2 * q] = l0 + off;
        sL1[gq][c0 + 2 * q + 1] = l1 + off;
      }
      __syncthreads();
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int i = h * 32 + lane;
      const double a = sA[i];
#pragma unroll
      for (int P = 0; P < 2; P++) {
        const double kv = kk[P][h], vn = vv[P][h];
        const double ev = se ? (np1 ? kv - sL1[2 * warp + P][i] : kv) : 0.0;
        sW[warp][4 * P + 0][i] = a * ev;
        sW[warp][4 * P + 1][i] = vn * ev;
        sW[warp][4 * P + 2][i] = a;
        sW[warp][4 * P + 3][i] = vn;
        sW[warp][8 + P][i] = kv;
      }
    }
    __syncwarp();
    for (int k4 = 0; k4 < WIDE_TILE; k4 += 4) {
      const double wa = sW[warp][gq][k4 + q];
#pragma unroll
      for (int f = 0; f < 4; f++)
        if (f < nfy) dmma884(acc[f][0], acc[f][1], wa, sYt[8 * f + gq][k4 + q]);
      const double sp = gq == 0 ? 1.0 : (gq == 1 ? sW[warp][8][k4 + q] : (gq == 2 ? sW[warp][9][k4 + q] : 0.0));
      dmma884(acc[4][0], acc[4][1], wa, sp);
    }
    __syncwarp();
  }
  // result tile of this warp -> shared memory [row][feature] (over its weight rows)
  double* sG = &sW[warp][0][0];
#pragma unroll
  for (int f = 0; f < 4; f++) {
    if (f < nfy) {
      sG[gq * WIDE_LDG + 8 * f + 2 * q] = acc[f][0];
      sG[gq * WIDE_LDG + 8 * f + 2 * q + 1] = acc[f][1];
    }
  }
  sG[gq * WIDE_LDG + 8 * nfy + 2 * q] = acc[4][0];
  sG[gq * WIDE_LDG + 8 * nfy + 2 * q + 1] = acc[4][1];
  if (nseg > 1) {
    cluster.sync();
    if (seg == 0) {  // fixed order: own partial, then ranks 1, 2, 3
      constexpr int NV = 8 * WIDE_LDG / 32;  // values per lane
      for (int r = 1; r < nseg; r++) {
        const double* rG = cluster.map_shared_rank(sG, r);
        double t[NV];
#pragma unroll
        for (int u = 0; u < NV; u++) t[u] = rG[lane + 32 * u];  // the remote loads in flight together
#pragma unroll
        for (int u = 0; u < NV; u++) sG[lane + 32 * u] += t[u];
      }
    }
    cluster.sync();  // the other ranks' shared memory stays alive until rank 0 has read it
    if (seg != 0) return;
  }
  __syncwarp();
  const int spc = 8 * nfy;  // first column of the special tile
#pragma unroll
  for (int P = 0; P < 2; P++) {
    const int p = 2 * warp + P, m = p0 + p;
    if (m >= M) continue;
    const double* G = sG + 4 * P * WIDE_LDG;  // rows: a e, v e, a, v
    const double E0a = G[0 * WIDE_LDG + spc], E0v = G[1 * WIDE_LDG + spc];
    if (lane < D) {
      const int j = lane;
      const double il2 = -2.0 * sIl[j] * sIl[j], xj = sX[p][j], w1 = np1 ? s.poly_w2[0][0][j] : 0.0;
      const double ga = il2 * (xj * E0a - G[0 * WIDE_LDG + j]) + w1 * G[2 * WIDE_LDG + j];
      const double gv = il2 * (xj * E0v - G[1 * WIDE_LDG + j]) + w1 * G[3 * WIDE_LDG + j];
      const double dkd = np1 ? 2.0 * w1 * xj : 0.0;  // d k(x,x) / dx_j for SE + linear: 2 w1_j x_j
      jmean[((size_t)m * E + e) * D + j] = ga;
      jvar[((size_t)m * E + e) * D + j] = var_scale * (dkd - 2.0 * gv);
    }
    if (lane == 0) {
      double kd = se ? s.lambda : 0.0;
      if (np1) {
        double L = off;
        for (int j = 0; j < D; j++) L = fma(s.poly_w2[0][0][j] * sX[p][j], sX[p][j], L);
        kd += L;
      }
      mean[(size_t)m * E + e] = s.mean0 + G[2 * WIDE_LDG + spc + 1 + P];
      var[(size_t)m * E + e] = var_scale * (kd - G[3 * WIDE_LDG + spc + 1 + P]);
    }
  }
}

__global__ void __launch_bounds__(WIDE_WPC * 32) posterior_reduce_wide_kernel(const __grid_constant__ McpGpSpec s, const double* __restrict__ Xs,
                                                                    int M, const double* __restrict__ Xtr,
                                                                    const double* __restrict__ alpha, int N, const double* __restrict__ Ks,
                                                                    const double* __restrict__ V, int ldv, double var_scale, int E,
                                                                    int e, double* __restrict__ mean, double* __restrict__ var,
                                                                    double* __restrict__ jmean, double* __restrict__ jvar) {
  reduce_wide_body(s, Xs, M, Xtr, alpha, N, Ks, V, ldv, var_scale, E, e, mean, var, jmean, jvar);
}

// the same reduce for ALL outputs in one launch (blockIdx.y = output); a link of a programmatic-dependent-launch chain
__global__ void __launch_bounds__(WIDE_WPC * 32) posterior_reduce_wide_batched_kernel(const McpGpDev* __restrict__ gps, const double* __restrict__ Xs,
                                                                            int M, const double* __restrict__ Ks, const double* __restrict__ V,
                                                                            int ldv, size_t gp_stride, int E, double* __restrict__ mean,
                                                                            double* __restrict__ var, double* __restrict__ jmean,
                                                                            double* __restrict__ jvar) {
  pdl_wait();
  const int e = blockIdx.y;
  const McpGpDev& g = gps[e];
  reduce_wide_body(g.spec, Xs, M, g.Xtr, g.alpha, g.N, Ks + e * gp_stride, V + e * gp_stride, ldv, g.var_scale, E, e, mean, var, jmean, jvar);
}

// launch with a runtime cluster size along z (and, optionally, programmatic stream serialisation)
template <typename... KArgs, typename... Args>
static cudaError_t launch_cluster_z(bool pdl, int nseg, size_t smem, void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = (unsigned)nseg;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

static bool wide_reduce_ok(const McpGpSpec& s) {
  return s.D > 8 && s.has_se && (s.n_poly == 0 || (s.n_poly == 1 && s.poly_deg[0] == 1));
}

static inline int ld16(int n) { return (n + 15) / 16 * 16; }

// opt-in INT8 tensor-core contraction (mcp_ozaki.cu)
size_t ozaki_scratch_bytes(int mc, int N, int S);
size_t ozaki_plane_bytes(int rows, int K, int S);
int ozaki_mma(const int8_t* Ap, const int32_t* Ae, const int8_t* Bp, const int32_t* Be, int M, int N, int S, int nseg, int Ksp, double* V, int ldv,
              cudaStream_t st);
int ozaki_contract(const double* A, int lda, int mc, int N, int S, const int8_t* Bplanes, const int32_t* Bexp, double* V, int ldv, void* scratch,
                   size_t scratch_bytes, cudaStream_t st);

// ---- batched posterior: all E outputs of a step in three launches (K* rows, contraction, reduce) on one stream ----------------
// For small rollouts with wide gp inputs (UR5: D = 24, E = 6, 200 particles) the per-output chains on side streams cost ~36 API calls
// per step and the rollout was host-bound (135 us per step); batched over the output index it is 3 launches per step.  Same device
// code as the per-output kernels, so the results are bit-identical.
bool gp_posterior_batched_ok(const McpGp* gps, int E, bool jac) {
  if (!jac || E < 2 || getenv("MCPILCO_NO_BATCHED_STEP") != nullptr) return false;
  for (int e = 0; e < E; e++) {
    const McpGp& g = gps[e];
    if (!wide_reduce_ok(g.spec) || g.ozaki_slices != 0 || g.spec.D != gps[0].spec.D) return false;
    if (wide_segments(g.N) != wide_segments(gps[0].N)) return false;  // one cluster size per launch, and the same split as the per-output path
    if (g.N < 1 || !g.Xtr || !g.alpha || !g.Kinv || g.ld_kinv < g.N || g.ld_kinv % 2 != 0 || ((uintptr_t)g.Kinv % 16) != 0) return false;
  }
  return true;
}

size_t gp_posterior_batched_doubles(int M, int E, int nmax) {
  return 2 * (size_t)E * M * ld16(nmax) + (sizeof(McpGpDev) * (size_t)E + 7) / 8 + 64;
}

// one-time setup per rollout: upload the table to the front of `scratch`; returns the device table and the K* / V base
int gp_posterior_batched_setup(const McpGp* gps, int E, int M, int nmax, double* scratch, size_t scratch_doubles, const McpGpDev** tab_out,
                               double** ks_out, cudaStream_t st) {
  MCP_CHECK_ARG(scratch_doubles >= gp_posterior_batched_doubles(M, E, nmax), "rollout (batched step): workspace too small");
  McpGpDev* tab = reinterpret_cast<McpGpDev*>(scratch);
  MCP_CUDA(gpdev_upload(tab, gps, E, st));
  double* Ks = scratch + (sizeof(McpGpDev) * (size_t)E + 7) / 8 + 8;
  *ks_out = reinterpret_cast<double*>(align_up((size_t)Ks, 256));
  *tab_out = tab;
  return MCP_OK;
}

int gp_posterior_batched(const McpGpDev* tab, int E, int D, int nmax, const double* Xs, int M, double* mean, double* var, double* jmean,
                         double* jvar, double* Ks, cudaStream_t st) {
  const int ldk = ld16(nmax);
  const size_t gp_stride = (size_t)M * ldk;
  double* V = Ks + (size_t)E * gp_stride;
  const bool pdl = pdl_enabled();
  dim3 cgrid(cdiv(ldk, CW_THREADS), cdiv(M, CW_ROWS), E);
  if (D <= 16) cov_wide_batched_kernel<16><<<cgrid, CW_THREADS, 0, st>>>(tab, Xs, M, Ks, ldk, gp_stride);
  else if (D <= 24) cov_wide_batched_kernel<24><<<cgrid, CW_THREADS, 0, st>>>(tab, Xs, M, Ks, ldk, gp_stride);
  else cov_wide_batched_kernel<32><<<cgrid, CW_THREADS, 0, st>>>(tab, Xs, M, Ks, ldk, gp_stride);
  MCP_LAUNCH_CHECK();
  if (int err = launch_small_gemm(tab, M, nmax, E, Ks, V, ldk, gp_stride, pdl, st)) return err;
  const int nseg = wide_segments(nmax);  // gp_posterior_batched_ok: every output has this segment count
  MCP_CUDA(launch_cluster_z(pdl, nseg, wide_smem_bytes(D), posterior_reduce_wide_batched_kernel, dim3(cdiv(M, WIDE_PPC), E, nseg), dim3(WIDE_WPC * 32), st, tab, Xs, M, Ks, V,
                            ldk, gp_stride, E, mean, var, jmean, jvar));
  count_launch();
  return MCP_OK;
}

// posterior of ONE GP for a chunk of particles through scratch [2 x Mc x ld16(N)]
int gp_posterior_chunk(const McpGp& g, int E, int e, const double* Xs, int M, double* mean, double* var, double* jmean,
                       double* jvar, double* scratch, size_t scratch_doubles, cudaStream_t st) {
  const int N = g.N, ldk = ld16(N);
  MCP_CHECK_ARG(N >= 1 && g.Xtr && g.alpha && g.Kinv, "gp %d: null training data", e);
  MCP_CHECK_ARG(g.ld_kinv >= N && g.ld_kinv % 2 == 0 && ((uintptr_t)g.Kinv % 16) == 0,
                "gp %d: Kinv must be 16-byte aligned with an even leading dimension >= N (ld=%d N=%d)", e, g.ld_kinv, N);
  const int oz = g.ozaki_slices;
  MCP_CHECK_ARG(oz == 0 || (g.kinv_planes && g.kinv_exp), "gp %d: ozaki_slices set without digit planes", e);
  MCP_CHECK_ARG(g.Linv == nullptr || (g.ld_linv >= N && g.ld_linv % 2 == 0 && ((uintptr_t)g.Linv % 16) == 0),
                "gp %d: Linv must be 16-byte aligned with an even leading dimension >= N", e);
  // doubles of scratch per particle: K* and V rows, plus (INT8 variant) digit planes, exponent and the int32 product planes
  size_t per = 2 * (size_t)ldk;
  if (oz) per += (ozaki_scratch_bytes(1024, N, oz) / 1024 + 7) / 8 + 1;
  MCP_CHECK_ARG(scratch_doubles >= per + (oz ? 16384 : 0), "posterior workspace too small for N=%d", N);
  const size_t usable = scratch_doubles - (oz ? 16384 : 0);
  int Mc = (int)((usable / per) < (size_t)M ? (usable / per) : (size_t)M);
  const bool jac = jmean != nullptr && jvar != nullptr;
  for (int m0 = 0; m0 < M; m0 += Mc) {
    int mc = (M - m0 < Mc) ? (M - m0) : Mc;
    double* Ks = scratch;
    double* V = scratch + (size_t)Mc * ldk;
    const double* xs = Xs + (size_t)m0 * g.spec.D;
    // INT8 variant with a kernel shape the fused K* kernel knows: the digit planes of K* come out of the K* kernel itself
    const bool oz_fused = oz && fast_reduce_ok(g.spec) && g.kdiag_max > 0.0;
    int8_t* oz_planes = nullptr;
    int32_t* oz_exp = nullptr;
    if (oz_fused) {
      char* pz = (char*)align_up((size_t)(scratch + 2 * (size_t)Mc * ldk), 256);
      oz_planes = (int8_t*)pz;
      oz_exp = (int32_t*)(pz + align_up(ozaki_plane_bytes(mc, N, oz), 256));
      if (int err = launch_cov_slice(g.spec, xs, mc, g.Xtr, N, g.kdiag_max, Ks, ldk, oz, oz_planes, oz_exp, st)) return err;
    } else {
      if (int err = launch_cov(g.spec, xs, mc, g.Xtr, N, 0, Ks, ldk, ldk, st)) return err;
    }
    if (!jac && !oz && g.Linv != nullptr && dgemm_tma_usable(Ks, ldk, g.Linv, g.ld_linv, V, ldk)) {
      // forward-only: w = K* L^-T over the triangle (column tile n0 needs k < n0 + tile), var = k** - |w|^2: N^2 instead of 2 N^2 flops
      prof_begin(st);
      if ((size_t)cdiv(mc, 128) * cdiv(N, 128) >= 96) {
        if (int err = dgemm_nt_tma_trim(mc, N, N, 1.0, Ks, ldk, g.Linv, g.ld_linv, 0.0, V, ldk, 0, KF_B_LOWER, st)) return err;
      } else {
        if (int err = dgemm_nt(mc, N, N, 1.0, Ks, ldk, g.Linv, g.ld_linv, 0.0, V, ldk, 0, KF_B_LOWER, st)) return err;
      }
      prof_end(st, (double)mc * (double)N * (double)N);
      MCP_DISPATCH_D(g.spec.D, (posterior_tri_reduce_kernel<DT><<<cdiv(mc, 8), 256, 0, st>>>(g.spec, xs, mc, g.alpha, N, Ks, V, ldk, g.var_scale, E, e,
                                                                                              mean + (size_t)m0 * E, var + (size_t)m0 * E)));
      MCP_LAUNCH_CHECK();
      continue;
    }
    prof_begin(st);
    if (oz_fused) {
      int nseg_, Ks_, Ksp_;
      ozaki_geometry(N, oz, &nseg_, &Ks_, &Ksp_);
      if (int err = ozaki_mma(oz_planes, oz_exp, g.kinv_planes, g.kinv_exp, mc, N, oz, nseg_, Ksp_, V, ldk, st)) return err;
    } else if (oz) {
      void* osc = (void*)(scratch + 2 * (size_t)Mc * ldk);
      const size_t osb = (scratch_doubles - 2 * (size_t)Mc * ldk) * sizeof(double);
      if (int err = ozaki_contract(Ks, ldk, mc, N, oz, g.kinv_planes, g.kinv_exp, V, ldk, osc, osb, st)) return err;
    } else
    // the TMA kernel's 128 x 128 tiles need enough of them to fill the GPU; below that the small-tile cp.async kernel wins
    if ((size_t)cdiv(mc, 128) * cdiv(N, 128) >= 96 && dgemm_tma_usable(Ks, ldk, g.Kinv, g.ld_kinv, V, ldk)) {
      if (int err = dgemm_nt_tma(mc, N, N, 1.0, Ks, ldk, g.Kinv, g.ld_kinv, V, ldk, st)) return err;
    } else {
      if (int err = dgemm_nt(mc, N, N, 1.0, Ks, ldk, g.Kinv, g.ld_kinv, 0.0, V, ldk, 0, 0, st)) return err;
    }
    prof_end(st, 2.0 * (double)mc * (double)N * (double)N);
    dim3 grid(cdiv(mc, 8));
    double* jm = jac ? jmean + (size_t)m0 * E * g.spec.D : nullptr;
    double* jv = jac ? jvar + (size_t)m0 * E * g.spec.D : nullptr;
    if (jac && fast_reduce_ok(g.spec)) {
#define MCP_FAST_REDUCE(DT_, NP_)                                                                                                   \
  do {                                                                                                                              \
    static bool cfg_[MCP_MAX_DEVICES] = {};                                                                                         \
    MCP_CUDA(ensure_dynamic_smem(cfg_, posterior_reduce_fast_kernel<DT_, NP_>, (int)red_smem_bytes<DT_>()));                        \
    posterior_reduce_fast_kernel<DT_, NP_><<<cdiv(mc, RED_PB), RED_THREADS, red_smem_bytes<DT_>(), st>>>(                           \
        g.spec, xs, mc, g.Xtr, g.alpha, N, Ks, V, ldk, g.var_scale, E, e, mean + (size_t)m0 * E, var + (size_t)m0 * E, jm, jv);     \
  } while (0)
      const int np_ = g.spec.n_poly;
      if (g.spec.D <= 4) { if (np_ == 0) MCP_FAST_REDUCE(4, 0); else if (np_ == 1) MCP_FAST_REDUCE(4, 1); else MCP_FAST_REDUCE(4, 2); }
      else if (g.spec.D <= 6) { if (np_ == 0) MCP_FAST_REDUCE(6, 0); else if (np_ == 1) MCP_FAST_REDUCE(6, 1); else MCP_FAST_REDUCE(6, 2); }
      else { if (np_ == 0) MCP_FAST_REDUCE(8, 0); else if (np_ == 1) MCP_FAST_REDUCE(8, 1); else MCP_FAST_REDUCE(8, 2); }
#undef MCP_FAST_REDUCE
    } else if (jac && wide_reduce_ok(g.spec)) {
      const int nseg = wide_segments(N);
      MCP_CUDA(launch_cluster_z(false, nseg, wide_smem_bytes(g.spec.D), posterior_reduce_wide_kernel, dim3(cdiv(mc, WIDE_PPC), 1, nseg), dim3(WIDE_WPC * 32), st, g.spec, xs, mc, g.Xtr,
                                g.alpha, N, Ks, V, ldk, g.var_scale, E, e, mean + (size_t)m0 * E, var + (size_t)m0 * E, jm, jv));
    } else if (jac) {
      MCP_DISPATCH_D(g.spec.D, (posterior_reduce_kernel<DT, true><<<grid, 256, 0, st>>>(
                                   g.spec, xs, mc, g.Xtr, g.alpha, N, V, ldk, g.var_scale, E, e, mean + (size_t)m0 * E,
                                   var + (size_t)m0 * E, jm, jv)));
    } else {
      MCP_DISPATCH_D(g.spec.D, (posterior_reduce_kernel<DT, false><<<grid, 256, 0, st>>>(
                                   g.spec, xs, mc, g.Xtr, g.alpha, N, V, ldk, g.var_scale, E, e, mean + (size_t)m0 * E,
                                   var + (size_t)m0 * E, nullptr, nullptr)));
    }
    MCP_LAUNCH_CHECK();
  }
  return MCP_OK;
}

}  // namespace mcp

extern "C" __attribute__((visibility("default"))) size_t mcpilco_gp_predict_workspace_bytes(int M, int Nmax) {
  // scratch for K* and V = K* K^-1 of one particle chunk (plus, when the opt-in INT8 contraction is possible for this N, its digit
  // planes and int32 product planes); capped at 2 GiB, at least 128 particles
  const int N = Nmax > 0 ? Nmax : 1;
  size_t per = 2 * (size_t)mcp::ld16(N) * sizeof(double);
  size_t fixed = 256;
  per += mcp::ozaki_scratch_bytes(1024, N, 8) / 1024 + 8;
  fixed += 16384 * sizeof(double);
  const size_t m = (size_t)(M > 0 ? M : 1);
  size_t want = per * m;
  const size_t cap = (size_t)2 << 30, floor_ = per * 128;
  if (want > cap) want = cap;
  if (want < floor_) want = floor_ < per * m ? floor_ : per * m;
  return want + fixed;
}

extern "C" __attribute__((visibility("default"))) int mcpilco_gp_predict(const McpGp* gps, int E, const double* Xs, int M, double* mean, double* var, double* jmean,
                                  double* jvar, void* workspace, size_t workspace_bytes, void* stream) {
  MCP_CHECK_ARG(gps && E >= 1 && E <= MCP_MAX_E, "gp_predict: bad E=%d", E);
  MCP_CHECK_ARG(M >= 0 && (M == 0 || (Xs && mean && var)), "gp_predict: null pointer");
  MCP_CHECK_ARG((jmean == nullptr) == (jvar == nullptr), "gp_predict: jmean and jvar must be given together");
  if (M == 0) return MCP_OK;
  MCP_CHECK_ARG(workspace && workspace_bytes >= 512, "gp_predict: workspace missing");
  double* scratch = (double*)align_up((size_t)workspace, 256);
  size_t doubles = (workspace_bytes - ((char*)scratch - (char*)workspace)) / sizeof(double);
  for (int e = 0; e < E; e++) {
    if (int err = check_spec(&gps[e].spec)) return err;
    MCP_CHECK_ARG(gps[e].spec.D == gps[0].spec.D, "gp_predict: all GPs must share the input dimension");
    if (int err = gp_posterior_chunk(gps[e], E, e, Xs, M, mean, var, jmean, jvar, scratch, doubles, (cudaStream_t)stream)) return err;
  }
  return MCP_OK;
}
